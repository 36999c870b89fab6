// Joint (text+image) non-causal attention forward for sm_100a, head_dim 128 — two query tiles per CTA ping-ponging on 128-key tiles
// (round 2).  Replaces `dispatch_attention_fn` -> F.scaled_dot_product_attention on the concatenated [txt; img] sequence
// (/root/reference/src/qflux/models/transformer_qwenimage.py:322-345, transformer_flux.py:149-156).
//
// Why: clock64 timelines of the round-1 kernel (64-key tiles, two CTAs per SM) and of a first two-tile version with Q in tensor memory and
// 64-key tiles show the same thing: a query tile's critical loop per key tile is  S ready -> softmax (64 exponentials per thread, ~940 clk)
// -> P V issue -> next S issue -> S ready (~1150 clk of issue + commit/wake latencies that do not shrink with the tile).  Two tiles
// interleave, so a 64-key step costs ~2100 clk however the tensor-memory budget is spent.  Here the fixed part is amortised over 128 keys
// and the softmax of a tile is split over TWO warps per lane quadrant (thread = query row x 64 key columns, row max combined through shared
// memory), so it still takes ~950 clk per step:
//   * one CTA owns 256 queries as two 128-row tiles g = 0, 1; tensor memory: S_0, S_1 (128 columns each, P_g overwrites S_g as bf16),
//     O_0, O_1 (128 each) = 512 columns; Q_g, and K / V rings of two 128-key tiles, in shared memory (192 KB);
//   * issue order per key tile j:  S_0(j) | P V_1(j-1) | S_1(j) | P V_0(j)  — the softmax of tile g runs under the other tile's two
//     contractions; P V takes P from tensor memory (TS mode); O_g is rescaled lazily (only when the row max grows by more than 2^8);
// MEASURED (B=4, H=24, S=2400, round 2): 0.341 ms against 0.286 ms for the round-1 64-key kernel (attn_fwd64_kernel) and 0.215 ms for cuDNN —
// NOT the default (opt-in: QFX_ATTN_FWD2=1).  The timeline (tools/attn_timeline_fwd.py) shows why: a (256 q x 128 k) step needs 32768
// exponentials = 2048 clk of the SM's MUFU pipe and ~1600 issue slots per scheduler for the two softmax passes, the two tiles' softmax phases
// overlap only half of the time, and between them each tile idles ~1600 clk in  P ready -> P V issue -> next S issue -> S ready  (commit /
// mbarrier wake-up latencies): 3650 clk per step against a 2048-2600 clk bound.  Moving half of the exponentials to the FMA pipe or spinning
// on test_wait made it slower (the passes are issue bound: 0.39 - 0.47 ms).  A first version with Q in tensor memory and 64-key tiles:
// 0.351 ms.  Kept as the starting point for a version whose S tiles are double-buffered (needs O in fewer tensor-memory columns).
//   * CTAs are ordered query-tile-major so the half-empty last tile of every head (2400 = 9 x 256 + 96: its second 128-row tile holds no
//     query and is skipped) is scheduled last and fills the tail wave.
#include <stdlib.h>
#include <string.h>

#include "../../include/qfx.h"
#include "host_common.h"
#include "sm100.cuh"

namespace qfx {

struct AttnFwd2Params {
  CUtensorMap tmQ, tmK, tmV;  // [B*H, S, 128] bf16, box {64, 128, 1}
  bf16* out0;                 // rows with joint position s < split  -> out0[(b*rows0 + s) * ld0 + h*128 ..]
  bf16* out1;                 // rows with s >= split               -> out1[(b*rows1 + s - split) * ld1 + h*128 ..]
  int64_t ld0, ld1;
  int rows0, rows1, split;
  float* lse;          // [B*H, S] log2-domain logsumexp:  m + log2(l)
  const int* kv_len;   // [B] end of the valid joint sequence per sample (NULL: S)
  const int* txt_len;  // [B] valid text tokens per sample (NULL: split): keys in [txt_len[b], split) are masked too
  int S, H, BH;
  float scale_log2;
  long long* dbg;  // optional clock64 stamps of CTA 0: [key tile][16] (tools/attn_timeline_fwd.py)
};

constexpr int F2_QT = 128 * 128 * 2;  // one 128-row tile of Q, K or V: 32 KB (two 64-column swizzle atoms of 16 KB)
constexpr int F2_ATOM = 128 * 128;    // 16 KB
constexpr int F2_NST = 2;             // K / V ring depth (128-key tiles)
constexpr int F2_RED = 2 * 2 * 128 * 4;                                 // row-max exchange: float [2 tiles][2 halves][128 rows]
constexpr int F2_SMEM = (2 + 2 * F2_NST) * F2_QT + F2_RED + 256;       // 198,912 B
constexpr int F2_THREADS = 640;  // warp 0 TMA, 1 MMA, 2-3 idle, 4-11 softmax of tile 0, 12-19 softmax of tile 1
                                 // (quadrant = warp & 3, key-column half = ((warp - 4) >> 2) & 1)

__device__ __forceinline__ void f2_mask_scores(uint32_t* r, int col0, int valid, int gap0, int gap1) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = col0 + i;
    if (!(k < valid) || (k >= gap0 && k < gap1)) r[i] = 0xff800000u;  // -inf
  }
}

__global__ void __launch_bounds__(F2_THREADS, 1) attn_fwd2q_kernel(const __grid_constant__ AttnFwd2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  auto sQ = [&](int g) { return smem_base + g * F2_QT; };
  auto sK = [&](int st) { return smem_base + (2 + st) * F2_QT; };
  auto sV = [&](int st) { return smem_base + (2 + F2_NST + st) * F2_QT; };
  const uint32_t red_base = smem_base + (2 + 2 * F2_NST) * F2_QT;
  const uint32_t bar_base = red_base + F2_RED;
  const uint32_t q_full = bar_base;
  auto k_full = [&](int st) { return bar_base + 8u * (1 + st); };
  auto k_empty = [&](int st) { return bar_base + 8u * (3 + st); };
  auto v_full = [&](int st) { return bar_base + 8u * (5 + st); };
  auto v_empty = [&](int st) { return bar_base + 8u * (7 + st); };
  auto s_full = [&](int g) { return bar_base + 8u * (9 + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (11 + g); };
  auto pv_done = [&](int g) { return bar_base + 8u * (13 + g); };
  auto o_full = [&](int g) { return bar_base + 8u * (15 + g); };
  const uint32_t tmem_slot = bar_base + 8u * 17;
  float* red_gen = reinterpret_cast<float*>(smem_raw + (red_base - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x / P.BH;  // query-tile-major CTA order: the light last tiles run last
  const int bh = blockIdx.x - qt * P.BH;
  const int b = bh / P.H, h = bh - b * P.H;
  const int q0 = qt * 256;
  const bool g1_on = q0 + 128 < P.S;  // the second 128-row tile holds at least one query
  const int n_groups = g1_on ? 2 : 1;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  const int n_tiles = (kv_len + 127) / 128;
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == 0;
#define F2_DBG(j, e) do { if (dbg_on && (j) < 40) P.dbg[(j) * 16 + (e)] = clock64(); } while (0)

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int st = 0; st < F2_NST; ++st) {
      mbar_init(k_full(st), 1);
      mbar_init(k_empty(st), 1);
      mbar_init(v_full(st), 1);
      mbar_init(v_empty(st), 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 8);
      mbar_init(pv_done(g), 1);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));
  auto tS = [&](int g) { return tmem_base + g * 128; };
  auto tO = [&](int g) { return tmem_base + 256 + g * 128; };

  if (warp == 0) {
    // ================================================================= TMA producers: lane 0 = Q tiles + K ring, lane 1 = V ring
    if (lane == 0) {
      tma_prefetch_desc(&P.tmQ);
      mbar_expect_tx(q_full, n_groups * F2_QT);
      for (int g = 0; g < n_groups; ++g) {
        tma_load_3d(sQ(g), &P.tmQ, q_full, 0, q0 + g * 128, bh);
        tma_load_3d(sQ(g) + F2_ATOM, &P.tmQ, q_full, 64, q0 + g * 128, bh);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % F2_NST;
        if (j >= F2_NST) mbar_wait(k_empty(st), ((j / F2_NST) - 1) & 1);
        mbar_expect_tx(k_full(st), F2_QT);
        tma_load_3d(sK(st), &P.tmK, k_full(st), 0, j * 128, bh);
        tma_load_3d(sK(st) + F2_ATOM, &P.tmK, k_full(st), 64, j * 128, bh);
      }
    } else if (lane == 1) {
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % F2_NST;
        if (j >= F2_NST) mbar_wait(v_empty(st), ((j / F2_NST) - 1) & 1);
        mbar_expect_tx(v_full(st), F2_QT);
        tma_load_3d(sV(st), &P.tmV, v_full(st), 0, j * 128, bh);
        tma_load_3d(sV(st) + F2_ATOM, &P.tmV, v_full(st), 64, j * 128, bh);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (whole warp, elected lane issues)
    constexpr uint32_t id_s = idesc_bf16(128, 128, 0, 0);   // S_g = Q_g K_j^T : A, B K-major
    constexpr uint32_t id_pv = idesc_bf16(128, 128, 0, 1);  // O_g += P_g V_j  : A in tensor memory, B MN-major [128 keys x 128 d]
    const uint32_t kQ[2] = {sdesc_lo(sQ(0), 16), sdesc_lo(sQ(1), 16)};
    auto issue_s = [&](int g, int j) {
      const uint32_t kK = sdesc_lo(sK(j % F2_NST), 16);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t inc = (uint32_t)((k >> 2) * 1024 + (k & 3) * 2);
        umma_ss_w(tS(g), kQ[g] + inc, kK + inc, id_s, k != 0);
      }
      umma_commit_w(s_full(g));
    };
    auto issue_pv = [&](int g, int j) {
      const uint32_t mV = sdesc_lo(sV(j % F2_NST), 16384);
      mbar_wait(p_full(g), j & 1);
      mbar_wait(v_full(j % F2_NST), (j / F2_NST) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_ts_w(tO(g), tS(g) + k * 8, mV + (uint32_t)(k * 128), id_pv, (j | k) != 0);
      umma_commit_w(pv_done(g));
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(k_full(j % F2_NST), (j / F2_NST) & 1);
      tc_fence_after();
      if (lane == 0) F2_DBG(j, 0);
      issue_s(0, j);  // S_0 is free: P V_0(j-1), the last reader of P_0(j-1), was issued before (MMAs complete in order)
      if (g1_on) {
        if (lane == 0) F2_DBG(j, 1);
        if (j > 0) {
          issue_pv(1, j - 1);
          umma_commit_w(v_empty((j - 1) % F2_NST));  // V_{j-1} has served both tiles
        }
        if (lane == 0) F2_DBG(j, 2);
        issue_s(1, j);
      }
      umma_commit_w(k_empty(j % F2_NST));  // K_j has served both tiles
      if (lane == 0) F2_DBG(j, 3);
      issue_pv(0, j);
      if (lane == 0) F2_DBG(j, 4);
      if (!g1_on) umma_commit_w(v_empty(j % F2_NST));
    }
    if (g1_on) {
      issue_pv(1, n_tiles - 1);
      umma_commit_w(v_empty((n_tiles - 1) % F2_NST));
    }
    umma_commit_w(o_full(0));
    umma_commit_w(o_full(1));
  } else if (warp >= 4) {
    // ================================================================= softmax warps: group g = query tile; two warps per lane quadrant,
    // thread = query row x 64 of the 128 key columns (and 64 of the 128 head-dim columns of O)
    const int g = (warp - 4) >> 3, quad = warp & 3, half = ((warp - 4) >> 2) & 1;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    if (g < n_groups) {
      const uint32_t tSg = tS(g) + lane_off + half * 64, tOg = tO(g) + lane_off + half * 64;
      float* red = red_gen + g * 256;  // [2 halves][128 rows]
      auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + g * 4 + quad) : "memory"); };
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < n_tiles; ++j) {
        const int valid = kv_len - j * 128;
        const int gap0 = txt_len - j * 128, gap1 = P.split - j * 128;
        const bool full_tile = valid >= 128 && (gap0 >= gap1 || gap0 >= 128 || gap1 <= 0);
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        if (quad == 0 && half == 0 && lane == 0) F2_DBG(j, 8 + 4 * g);
        // two passes over the 64 score columns (tensor-memory reads cost ~50 clk; holding all 64 scores AND the packed P spilled)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t r[32];
          tmem_ld32(tSg + cc * 32, r);
          tmem_ld_wait();
          if (!full_tile) f2_mask_scores(r, half * 64 + cc * 32, valid, gap0, gap1);
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // combine the row max of the two column halves
        red[half * 128 + row] = mx;
        tc_fence_before();
        pair_sync();
        tc_fence_after();
        mx = fmaxf(mx, red[(half ^ 1) * 128 + row]);
        const float m_new = fmaxf(m_used, mx * P.scale_log2);  // a fully padded tile: mx = -inf, m_new = m_used (tile 0 has a valid key)
        if (__any_sync(0xffffffffu, (j == 0) || (m_new > m_used + 8.f))) {  // warp-uniform lazy rescale (same rows in both halves -> same decision)
          if (j > 0) {
            mbar_wait(pv_done(g), (j - 1) & 1);  // all earlier P V of this tile have landed in O_g
            tc_fence_after();
            const float alpha = exp2f(m_used - m_new);
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
              uint32_t o[32];
              tmem_ld32(tOg + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(tOg + c, o);
            }
            tmem_st_wait();
            l *= alpha;
          }
          m_used = m_new;
        }
        uint32_t pk[32];
        float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t r[32];
          tmem_ld32(tSg + cc * 32, r);
          tmem_ld_wait();
          if (!full_tile) f2_mask_scores(r, half * 64 + cc * 32, valid, gap0, gap1);  // exp2(-inf) = 0 for masked keys
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = exp2f(__uint_as_float(r[2 * i]) * P.scale_log2 - m_used);
            const float p1 = exp2f(__uint_as_float(r[2 * i + 1]) * P.scale_log2 - m_used);
            pk[cc * 16 + i] = pack_bf16(p0, p1);
            ls4[i & 3] += p0 + p1;
          }
        }
        l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
        tc_fence_before();
        // the partner has finished reading S_g (this warp's bf16 P columns [32h, 32h+32) lie over score columns the OTHER half reads when
        // h = 1) and its red[] slot of this tile
        pair_sync();
        tc_fence_after();
        tmem_st32(tS(g) + lane_off + half * 32, pk);  // P_g(j): 64 keys -> 32 packed columns
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
        if (quad == 0 && half == 0 && lane == 0) F2_DBG(j, 9 + 4 * g);
      }
      // ----------------------------------------------------------------- epilogue: combine the partial row sums, O / l -> bf16
      red[half * 128 + row] = l;
      pair_sync();
      l += red[(half ^ 1) * 128 + row];
      mbar_wait(o_full(g), 0);
      tc_fence_after();
      const int sq = q0 + g * 128 + row;
      const bool ok = sq < P.S;
      const float inv = 1.f / l;
      bf16* dst = nullptr;
      if (ok) {
        dst = (sq < P.split ? P.out0 + ((int64_t)b * P.rows0 + sq) * P.ld0 + h * 128
                            : P.out1 + ((int64_t)b * P.rows1 + (sq - P.split)) * P.ld1 + h * 128) + half * 64;
        if (P.lse && half == 0) P.lse[(int64_t)bh * P.S + sq] = m_used + log2f(l);
      }
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t o[32];
        tmem_ld32(tOg + c, o);
        tmem_ld_wait();
        if (ok) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c);
#pragma unroll
          for (int v = 0; v < 4; ++v)
            d4[v] = make_uint4(pack_bf16(__uint_as_float(o[8 * v]) * inv, __uint_as_float(o[8 * v + 1]) * inv),
                               pack_bf16(__uint_as_float(o[8 * v + 2]) * inv, __uint_as_float(o[8 * v + 3]) * inv),
                               pack_bf16(__uint_as_float(o[8 * v + 4]) * inv, __uint_as_float(o[8 * v + 5]) * inv),
                               pack_bf16(__uint_as_float(o[8 * v + 6]) * inv, __uint_as_float(o[8 * v + 7]) * inv));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int make_qkv_tmap(CUtensorMap* m, const void* base, int BH, int S);
}  // namespace qfx
extern long long* g_qfx_attn_bwd_dbg;  // shared debugging hook (qfx_attn_bwd_set_debug)
namespace qfx {

int attn_fwd_two_tiles(const void* Q, const void* K, const void* V, void* out0, int64_t ld0, int rows0, void* out1, int64_t ld1,
                       int rows1, int split, float* lse, const int* kv_len, const int* txt_len, int B, int H, int S, float softmax_scale,
                       cudaStream_t stream) {
  AttnFwd2Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = make_qkv_tmap(&P.tmQ, Q, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmK, K, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmV, V, B * H, S))) return rc;
  P.out0 = (bf16*)out0; P.out1 = (bf16*)out1;
  P.ld0 = ld0; P.ld1 = ld1; P.rows0 = rows0; P.rows1 = rows1; P.split = split;
  P.lse = lse; P.kv_len = kv_len; P.txt_len = txt_len; P.S = S; P.H = H; P.BH = B * H;
  P.scale_log2 = softmax_scale * 1.4426950408889634f;
  P.dbg = g_qfx_attn_bwd_dbg;
  static bool attr_done = false;
  if (!attr_done) {
    QFX_CUDA(cudaFuncSetAttribute(attn_fwd2q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    attr_done = true;
  }
  const int n_qt = (S + 255) / 256;
  attn_fwd2q_kernel<<<n_qt * B * H, F2_THREADS, F2_SMEM, stream>>>(P);
  QFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qfx
