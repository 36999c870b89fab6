// Joint attention backward for sm_100a (head_dim 128): dQ, dK, dV from Q, K, V, dO, logsumexp and delta = rowsum(dO*O).
// This is the autograd backward of F.scaled_dot_product_attention as the reference runs it
// (/root/reference/src/qflux/models/transformer_qwenimage.py:329-337 under loss.backward(), base_trainer.py:528).
//
// One CTA per (128-key tile j, batch*head); it loops over the 128-query tiles i.  All five contractions run on
// tcgen05 with accumulators in TMEM (512 columns: S|dQ, dP, dV, dK):
//     S   = Q_i K_j^T          A = Q_i  (K-major)   B = K_j (K-major)
//     dP  = dO_i V_j^T         A = dO_i (K-major)   B = V_j (K-major)
//     P   = exp2(S*c - L_i) ;  dS = P o (dP - delta_i) * scale        (compute warps: thread == query row)
//     dV += P^T  dO_i          A = P   read MN-major (transposed)   B = dO_i read MN-major
//     dK += dS^T Q_i           A = dS  read MN-major                B = Q_i  read MN-major
//     dQ_i = dS K_j            A = dS  (K-major)                    B = K_j  read MN-major   -> reuses the S columns
// The same 128B-swizzled shared-memory tile serves as K-major or MN-major operand — only the descriptor changes — so
// nothing is transposed in memory.  dQ_i is staged fp32 in the (free) P/dS buffers and added to the global fp32
// accumulator with one TMA reduce-add (cp.reduce.async.bulk.tensor) per 32-column slab instead of 16K scalar atomics.
#include <stdlib.h>
#include <string.h>

#include "../../include/qfx.h"
#include "host_common.h"
#include "sm100.cuh"

namespace qfx {

constexpr int BW_TILE = 128 * 128 * 2;  // 32 KB bf16 tile
constexpr int BW_ATOM = 128 * 128;      // 16 KB: one 64-wide swizzle atom column of 128 rows

struct AttnBwdParams {
  CUtensorMap tmQ, tmK, tmV, tmdO;  // bf16 [B*H, S, 128], box {64,128,1}
  CUtensorMap tmdQ;                 // fp32 [B*H, S, 128], box {32,128,1}  (reduce-add target)
  bf16* dK;
  bf16* dV;             // [B*H, S, 128]
  const float* lse;     // [B*H, S] log2 domain
  const float* delta;   // [B*H, S]
  const int* kv_len;    // [B] or NULL: keys >= kv_len[b] masked
  const int* txt_len;   // [B] or NULL: keys in [txt_len[b], split) masked (text padding)
  int split;
  long long* dbg;       // optional: clock64 stamps of CTA (1, 0) — [iteration][16] (tools/attn_timeline.py)
  int S, H;
  float scale, scale_log2;
};

constexpr int BW_SMEM = 7 * BW_TILE + 256;  // K, V, Q x2, dO, P, dS (+ barriers); 1024-aligned base, no slack

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ void mask_scores_bwd(uint32_t* r, int col0, int valid, int gap0, int gap1) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = col0 + i;
    if (!(k < valid) || (k >= gap0 && k < gap1)) r[i] = 0xff800000u;  // -inf
  }
}

// Pipeline per query tile i (key tile fixed per CTA):
//   MMA     : S = Q_i K^T, dP = dO_i V^T                 -> s_full
//   compute : P = exp2(S c - L)  -> smem                    -> p_full      (thread = query row x 64 key columns)
//   MMA     : dV += P^T dO_i                                -> do_empty    (dO_{i+1} starts loading)
//   compute : dS = P (dP - delta) scale -> smem (over P, after dv_done)  -> ds_full
//   MMA     : dK += dS^T Q_i  -> q_empty[i&1] ;  dQ_i = dS K (into the S columns)  -> dq_full
//   compute : dQ_i TMEM -> fp32 slabs parked in the consumed Q_i slot + the dS buffer -> dq_empty (S_{i+1}/dP_{i+1} may issue)
//             -> TMA reduce-add -> stage_free / qstage_free   (P is not aliased: the next P pass overlaps the reduce)
// Q is double-buffered and dO is released right after dV, so the loads of tile i+1 hide behind the math of tile i.
__global__ void __launch_bounds__(320, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t sK = smem_base, sV = smem_base + BW_TILE, sdO = smem_base + 4 * BW_TILE;
  auto sQ = [&](int st) { return smem_base + (2 + st) * BW_TILE; };
  const uint32_t sP = smem_base + 5 * BW_TILE, sdS = smem_base + 6 * BW_TILE;
  const uint32_t bar_base = smem_base + 7 * BW_TILE;
  const uint32_t kv_full = bar_base, do_full = bar_base + 8, do_empty = bar_base + 16, s_full = bar_base + 24;
  const uint32_t p_full = bar_base + 32, ds_full = bar_base + 40, dq_full = bar_base + 48, dq_empty = bar_base + 56;
  const uint32_t stage_free = bar_base + 64, acc_full = bar_base + 72;
  auto q_full = [&](int st) { return bar_base + 80 + 8u * st; };
  auto q_empty = [&](int st) { return bar_base + 96 + 8u * st; };
  auto qstage_free = [&](int st) { return bar_base + 112 + 8u * st; };  // dQ reduce finished reading the slabs parked in Q slot st
  const uint32_t dq_staged = bar_base + 128;  // all 8 compute warps have parked their dQ slabs in shared memory
  const uint32_t dv_done = bar_base + 136;    // dV_i has consumed P: the P buffer may be overwritten with dS
  const uint32_t tmem_slot = bar_base + 144;
  uint8_t* smem_gen = smem_raw;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int b = bh / P.H;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int n_q = (P.S + 127) / 128;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  // a fully masked key tile (beyond kv_len, or entirely inside the text padding) contributes nothing: write zeros and leave
  const bool active = kv0 < kv_len && !(kv0 >= txt_len && kv0 + 128 <= P.split);
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0;
#define DBG(i, e) do { if (dbg_on) P.dbg[(i) * 16 + (e)] = clock64(); } while (0)

  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    mbar_init(stage_free, 1);
    mbar_init(dq_staged, 8);
    mbar_init(dv_done, 1);
    mbar_init(acc_full, 1);
    for (int st = 0; st < 2; ++st) {
      mbar_init(q_full(st), 1);
      mbar_init(q_empty(st), 1);
      mbar_init(qstage_free(st), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 0) {
    // ================================================================= TMA producers: lane 0 = K, V, Q tiles; lane 1 = dO tiles
    if (lane == 0 && active) {
      mbar_expect_tx(kv_full, 2 * BW_TILE);
      tma_load_3d(sK, &P.tmK, kv_full, 0, kv0, bh);
      tma_load_3d(sK + BW_ATOM, &P.tmK, kv_full, 64, kv0, bh);
      tma_load_3d(sV, &P.tmV, kv_full, 0, kv0, bh);
      tma_load_3d(sV + BW_ATOM, &P.tmV, kv_full, 64, kv0, bh);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        if (i >= 2) {
          mbar_wait(q_empty(st), ((i - 2) >> 1) & 1);      // dK_{i-2} has consumed Q_{i-2}
          mbar_wait(qstage_free(st), ((i - 2) >> 1) & 1);  // ... and the dQ_{i-2} slabs parked in this slot have been reduced out
        }
        mbar_expect_tx(q_full(st), BW_TILE);
        tma_load_3d(sQ(st), &P.tmQ, q_full(st), 0, i * 128, bh);
        tma_load_3d(sQ(st) + BW_ATOM, &P.tmQ, q_full(st), 64, i * 128, bh);
      }
    } else if (lane == 1 && active) {
      for (int i = 0; i < n_q; ++i) {
        if (i >= 1) mbar_wait(do_empty, (i - 1) & 1);
        mbar_expect_tx(do_full, BW_TILE);
        tma_load_3d(sdO, &P.tmdO, do_full, 0, i * 128, bh);
        tma_load_3d(sdO + BW_ATOM, &P.tmdO, do_full, 64, i * 128, bh);
      }
    } else if (lane == 2 && active) {
      // dQ reducer: a thread of its own, so that waiting for the TMA reduce-add to drain shared memory never stalls a compute warp
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(dq_staged, i & 1);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
          tma_reduce_add_3d(&P.tmdQ, sl < 2 ? sQ(i & 1) + sl * BW_ATOM : sdS + (sl - 2) * BW_ATOM, sl * 32, i * 128, bh);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        DBG(i, 14);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(stage_free);
        mbar_arrive(qstage_free(i & 1));
        DBG(i, 15);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all reduce-adds have been performed before the CTA exits
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0 && active) {
      constexpr uint32_t id_kk = idesc_bf16(128, 128, 0, 0);  // A K-major, B K-major
      constexpr uint32_t id_mm = idesc_bf16(128, 128, 1, 1);  // A MN-major, B MN-major
      constexpr uint32_t id_km = idesc_bf16(128, 128, 0, 1);  // A K-major, B MN-major
      auto kmaj = [](uint32_t base, int k) { return sdesc_sw128(base + (k >> 2) * BW_ATOM + (k & 3) * 32, 16, 1024); };
      auto mnmaj = [](uint32_t base, int k) { return sdesc_sw128(base + k * 2048, BW_ATOM, 1024); };
      mbar_wait(kv_full, 0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        const uint32_t q = sQ(st);
        mbar_wait(q_full(st), (i >> 1) & 1);
        mbar_wait(do_full, i & 1);
        DBG(i, 0);
        if (i > 0) mbar_wait(dq_empty, (i - 1) & 1);  // S columns (shared with dQ) drained
        tc_fence_after();
        DBG(i, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tS, kmaj(q, k), kmaj(sK, k), id_kk, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdP, kmaj(sdO, k), kmaj(sV, k), id_kk, k != 0);
        umma_commit(s_full);
        mbar_wait(p_full, i & 1);
        tc_fence_after();
        DBG(i, 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdV, mnmaj(sP, k), mnmaj(sdO, k), id_mm, (i | k) != 0);
        umma_commit(do_empty);  // dO_i no longer needed
        umma_commit(dv_done);   // ... and neither is P: dS is written over it (the old dS buffer only stages dQ now)
        mbar_wait(ds_full, i & 1);
        tc_fence_after();
        DBG(i, 3);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdK, mnmaj(sP, k), mnmaj(q, k), id_mm, (i | k) != 0);
        umma_commit(q_empty(st));  // Q_i no longer needed
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tS, kmaj(sP, k), mnmaj(sK, k), id_km, k != 0);
        umma_commit(dq_full);
      }
      umma_commit(acc_full);
    }
  } else {
    // ================================================================= compute warps: 2 per TMEM lane quadrant,
    // thread = query row x 64 of the 128 key columns (and 64 of the 128 head-dim columns when draining dQ / dK / dV)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int c0 = half * 64;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int tid = threadIdx.x - 64;  // 0..255
    if (active) {
      const int valid = kv_len - kv0;  // key columns >= valid are masked
      const int gap0 = txt_len - kv0, gap1 = P.split - kv0;  // tile-local text-padding columns [gap0, gap1)
      const bool full_tile = valid >= 128 && (gap0 >= gap1 || gap0 >= 128 || gap1 <= 0);
      const uint32_t off_row = (uint32_t)row * 128 + half * BW_ATOM;  // this thread's 64 columns = one swizzle-atom row
      for (int i = 0; i < n_q; ++i) {
        const int q = i * 128 + row;
        const bool q_ok = q < P.S;
        const float L = q_ok ? P.lse[(int64_t)bh * P.S + q] : INFINITY;
        const float dls = (q_ok ? P.delta[(int64_t)bh * P.S + q] : 0.f) * P.scale;
        if (tid == 0) DBG(i, 8);
        mbar_wait(s_full, i & 1);
        tc_fence_after();
        if (tid == 0) DBG(i, 9);
        float p[64];
        // ---- pass 1: P
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t rs[32];
          tmem_ld32(tS + lane_off + c0 + cc * 32, rs);
          tmem_ld_wait();
          uint32_t pk[16];
          if (!full_tile) mask_scores_bwd(rs, c0 + cc * 32, valid, gap0, gap1);  // rare: masked keys -> -inf -> P = 0
#pragma unroll
          for (int e = 0; e < 32; ++e) p[cc * 32 + e] = exp2f(__uint_as_float(rs[e]) * P.scale_log2 - L);
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(p[cc * 32 + 2 * e], p[cc * 32 + 2 * e + 1]);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t chunk = (uint32_t)(cc * 4 + v) ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off_row + chunk * 16), "r"(pk[4 * v]),
                         "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (tid == 0) DBG(i, 11);
        // ---- pass 2: dS = P * (dP * scale - delta * scale), written over P once dV_i has consumed it
        if (tid == 0) DBG(i, 10);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t rp[32];
          tmem_ld32(tdP + lane_off + c0 + cc * 32, rp);
          tmem_ld_wait();
          uint32_t dk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e)
            dk[e] = pack_bf16(p[cc * 32 + 2 * e] * (__uint_as_float(rp[2 * e]) * P.scale - dls),
                              p[cc * 32 + 2 * e + 1] * (__uint_as_float(rp[2 * e + 1]) * P.scale - dls));
          if (cc == 0) mbar_wait(dv_done, i & 1);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t chunk = (uint32_t)(cc * 4 + v) ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off_row + chunk * 16), "r"(dk[4 * v]),
                         "r"(dk[4 * v + 1]), "r"(dk[4 * v + 2]), "r"(dk[4 * v + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
        if (tid == 0) DBG(i, 12);

        // ---- drain dQ_i: TMEM -> fp32 swizzled slabs in the (now idle) P|dS buffers -> TMA reduce-add to global
        mbar_wait(dq_full, i & 1);
        tc_fence_after();
        if (tid == 0) DBG(i, 13);
        if (i > 0) mbar_wait(stage_free, (i - 1) & 1);  // the staging buffer of dQ_{i-1} has been reduced out (long ago)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t r[32];
          const int c = c0 + cc * 32;
          tmem_ld32(tS + lane_off + c, r);
          tmem_ld_wait();
          // [128 rows x 32 fp32] slabs, 128 B rows: head-dim columns 0..63 are parked in the (consumed) Q_i slot, 64..127 in a staging
          // buffer of their own — nothing the next tile writes is aliased, so its P / dS passes never wait for this reduce
          const uint32_t slab = (c < 64 ? sQ(i & 1) + (c >> 5) * BW_ATOM : sdS + ((c - 64) >> 5) * BW_ATOM) + (uint32_t)row * 128;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint32_t chunk = (uint32_t)v ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + chunk * 16), "r"(r[4 * v]), "r"(r[4 * v + 1]),
                         "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                         : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);  // the S columns are free: S_{i+1} / dP_{i+1} may issue while the reduce drains
        if (tid == 0) DBG(i, 7);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_staged);  // the reducer thread (warp 0, lane 2) takes it from here
      }
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    // ----------------------------------------------------------------- epilogue: dK, dV tiles
    const int kv = kv0 + row;
    if (kv0 < P.S) {
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        bf16* dst = (which ? P.dK : P.dV) + ((int64_t)bh * P.S + kv) * 128;
        const uint32_t t = (which ? tdK : tdV) + lane_off;
#pragma unroll 1
        for (int c = c0; c < c0 + 64; c += 32) {
          uint32_t r[32];
          if (active) {
            tmem_ld32(t + c, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = 0;
          }
          if (kv < P.S) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + c);
#pragma unroll
            for (int v = 0; v < 4; ++v)
              d4[v] = make_uint4(pack_bf16(__uint_as_float(r[8 * v]), __uint_as_float(r[8 * v + 1])),
                                 pack_bf16(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3])),
                                 pack_bf16(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5])),
                                 pack_bf16(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7])));
          }
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

using namespace qfx;

long long* g_qfx_attn_bwd_dbg = nullptr;
/* debugging aid: device buffer of >= 16 * n_query_tiles int64 that receives clock64 stamps of CTA (1,0) (NULL disables) */
extern "C" void qfx_attn_bwd_set_debug(long long* buf) { g_qfx_attn_bwd_dbg = buf; }

/* All of Q, K, V, dO, dK, dV: [B, H, S, 128] bf16; dQ_accum: [B, H, S, 128] fp32, MUST be zeroed by the caller (it is the
 * target of TMA reduce-adds from every key tile).  lse: log2-domain logsumexp from qfx_attn_fwd; delta = rowsum(dO*O). */
namespace qfx {
int attn_bwd_pipelined(const void* Q, const void* K, const void* V, const void* dO, const float* lse, const float* delta, float* dQ,
                       void* dK, void* dV, const int* kv_len, const int* txt_len, int split, int B, int H, int S, float softmax_scale,
                       cudaStream_t stream);  // attention_bwd3.cu
}

extern "C" int qfx_attn_bwd(const void* Q, const void* K, const void* V, const void* dO, const float* lse, const float* delta,
                            float* dQ_accum, void* dK, void* dV, const int* kv_len, const int* txt_len, int split, int B, int H,
                            int S, float softmax_scale, void* stream) {
  extern long long* g_qfx_attn_bwd_dbg;
  QFX_CHECK_ARG(B > 0 && H > 0 && S > 0 && dQ_accum && dK && dV && lse && delta, "qfx_attn_bwd: bad arguments");
  // default: the software-pipelined transposed kernel (attention_bwd3.cu).  A/B switch: QFX_ATTN_BWD1=1 runs the round-1 serial-chain
  // kernel below (the same-box comparison in profiles/r02_bench_b_*.json).  The 64-query transposed variant of round 1 lives on as
  // tools/experiments/attention_bwd_transposed_64q.cu — it never beat either and is no longer part of the library.
  static const bool serial = getenv("QFX_ATTN_BWD1") != nullptr;
  if (!serial)
    return qfx::attn_bwd_pipelined(Q, K, V, dO, lse, delta, dQ_accum, dK, dV, kv_len, txt_len, split, B, H, S, softmax_scale,
                                   (cudaStream_t)stream);
  AttnBwdParams P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = make_qkv_tmap(&P.tmQ, Q, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmK, K, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmV, V, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmdO, dO, B * H, S))) return rc;
  {
    uint64_t dims[3] = {128, (uint64_t)S, (uint64_t)(B * H)};
    uint64_t strides[2] = {128 * 4, (uint64_t)S * 128 * 4};
    uint32_t box[3] = {32, 128, 1};
    if ((rc = make_tmap_f32(&P.tmdQ, dQ_accum, 3, dims, strides, box))) return rc;
  }
  P.dK = (bf16*)dK; P.dV = (bf16*)dV; P.lse = lse; P.delta = delta; P.kv_len = kv_len; P.txt_len = txt_len; P.split = txt_len ? split : 0; P.S = S; P.H = H;
  P.dbg = g_qfx_attn_bwd_dbg;
  P.scale = softmax_scale;
  P.scale_log2 = softmax_scale * 1.4426950408889634f;
  static bool attr_done = false;
  if (!attr_done) {
    QFX_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    attr_done = true;
  }
  dim3 grid((S + 127) / 128, B * H);
  attn_bwd_kernel<<<grid, 320, BW_SMEM, (cudaStream_t)stream>>>(P);
  QFX_CUDA(cudaGetLastError());
  return 0;
}
