// Joint attention backward for sm_100a (head_dim 128), transposed single-kernel formulation — selected by QFX_ATTN_BWD2=1
// (parity-green on every attn_bwd case and the whole GPU suite; 0.98 ms against 1.01 ms for attention_bwd.cu at B=4, H=24, S=2400 —
// a tie in the full step, so attention_bwd.cu stays the default for round 1; see DESIGN.md 4.2 for what the variants showed).
// This is the autograd backward of F.scaled_dot_product_attention as the reference runs it
// (/root/reference/src/qflux/models/transformer_qwenimage.py:329-337 under loss.backward(), base_trainer.py:528).
//
// One CTA per (128-key tile, batch*head) loops over 64-QUERY tiles i with the scores TRANSPOSED in tensor memory, which is what
// lets everything fit and overlap (the first backward, attention_bwd.cu, keeps S|dQ, dP, dV, dK of a 128 x 128 pair in all 512
// TMEM columns and therefore runs its five contractions and three register passes as one serial chain, 6600 clk per pair):
//     S^T  = K Q_i^T, dP^T = V dO_i^T      [128 keys x 64 q] each, DOUBLE-buffered (4 x 64 columns): tile i+1 is issued before tile i's
//                                          accumulations, so its MMAs run under tile i's exponentials
//     P^T  = exp2(S^T c - L_q), dS^T = P^T o (dP^T - delta_q) c      8 compute warps, thread = key row x 32 query columns; the per-query
//                                          L / delta are per-COLUMN here and come from a small shared-memory ring (stager warp)
//     dV  += P^T dO_i, dK += dS^T Q_i      accumulators [128 keys x 128 d] (2 x 128 columns) over the whole query loop
//     dQ_i^T = K^T dS^T                    [128 d x 64 q] into the S^T columns the P pass has just consumed; 4 drain warps (thread =
//                                          head-dim index) add it to the fp32 dQ with red.global.add.f32 — a warp instruction covers one
//                                          full 128-byte line of a query row (the [q x d] formulation needed smem staging + TMA reduce)
// K is read K-major for S^T and MN-major (as K^T) for dQ^T; Q_i / dO_i K-major for the scores and MN-major for dK / dV — only the
// UMMA descriptors differ.  Keys >= kv_len[b] and text padding [txt_len[b], split) are masked per key ROW.
#include <stdlib.h>
#include <string.h>

#include "../../include/qfx.h"
#include "host_common.h"
#include "sm100.cuh"

namespace qfx {

struct AttnBwd2Params {
  CUtensorMap tmK128, tmV128;   // box {64, 128, 1}: the CTA's K / V tile
  CUtensorMap tmQ64, tmdO64;    // box {64, 64, 1}:  64-query tiles of Q / dO
  float* dQ;                    // [B*H, S, 128] fp32, zeroed by the caller: every key tile adds its partial dQ
  bf16* dK;
  bf16* dV;
  const float* lse;    // log2 domain
  const float* delta;
  const int* kv_len;
  const int* txt_len;
  int split, S, H;
  float scale, scale_log2;
};

constexpr int DQ_T128 = 128 * 128 * 2;  // 32 KB: [128 rows x 128] bf16 tile (two 64-column swizzle atoms of 16 KB)
constexpr int DQ_T64 = 64 * 128 * 2;    // 16 KB: [64 rows x 128]

constexpr int KV_STAGES = 3;
constexpr int KV_LD = 4;  // L / delta ring depth: the stager runs up to 3 query tiles ahead (its global loads take ~1 tile period)
constexpr int KV_SMEM = 2 * DQ_T128 + KV_STAGES * 2 * DQ_T64 + 2 * 128 * 128 + KV_LD * 512 + 256;  // K, V, (Q_i, dO_i) x3, P^T, dS^T, L/delta ring
constexpr int KV_THREADS = 512;  // warp 0 TMA, 1 MMA, 2 L/delta stager, 3 idle, 4-11 compute (quad = warp & 3, column half = (warp - 4) >> 2),
                                 // 12-15 dQ drain (quad = warp & 3)

__global__ void __launch_bounds__(KV_THREADS, 1) attn_bwd_t_kernel(const __grid_constant__ AttnBwd2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t sK = smem_base, sV = smem_base + DQ_T128;
  auto sQ = [&](int st) { return smem_base + 2 * DQ_T128 + st * 2 * DQ_T64; };
  auto sdO = [&](int st) { return smem_base + 2 * DQ_T128 + st * 2 * DQ_T64 + DQ_T64; };
  const uint32_t sPt = smem_base + 2 * DQ_T128 + KV_STAGES * 2 * DQ_T64, sdSt = sPt + 128 * 128;
  const uint32_t sLD = sdSt + 128 * 128;  // float [KV_LD buffers][2 (L, delta*scale)][64]
  const uint32_t bar_base = sLD + KV_LD * 512;
  const uint32_t kv_full = bar_base;
  auto q_full = [&](int st) { return bar_base + 8u * (1 + st); };
  auto q_empty = [&](int st) { return bar_base + 8u * (4 + st); };
  auto s_full = [&](int u) { return bar_base + 8u * (7 + u); };     // S^T / dP^T buffer u written by the tensor core
  auto ld_full = [&](int u) { return bar_base + 8u * (9 + u); };    // L / delta of the tile in ring slot u staged in smem
  auto ld_empty = [&](int u) { return bar_base + 8u * (13 + u); };  // ... and read by all 8 compute warps
  const uint32_t p_full = bar_base + 8u * 17, ds_full = bar_base + 8u * 18, dv_done = bar_base + 8u * 19, dk_done = bar_base + 8u * 20;
  const uint32_t acc_full = bar_base + 8u * 21;
  auto dq_full = [&](int u) { return bar_base + 8u * (22 + u); };     // dQ^T of the tile in buffer u has been written by the tensor core
  auto dq_drained = [&](int u) { return bar_base + 8u * (24 + u); };  // ... and read out of tensor memory by the 4 drain warps
  const uint32_t tmem_slot = bar_base + 8u * 26;
  float* ld_gen = reinterpret_cast<float*>(smem_raw + (sLD - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int b = bh / P.H;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  const int n_q = (P.S + 63) / 64;
  const bool active = kv0 < kv_len && !(kv0 >= txt_len && kv0 + 128 <= P.split);

  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int st = 0; st < KV_STAGES; ++st) {
      mbar_init(q_full(st), 1);
      mbar_init(q_empty(st), 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(s_full(u), 1);
      mbar_init(dq_full(u), 1);
      mbar_init(dq_drained(u), 4);
    }
    for (int u = 0; u < KV_LD; ++u) {
      mbar_init(ld_full(u), 1);
      mbar_init(ld_empty(u), 8);
    }
    mbar_init(p_full, 8);
    mbar_init(ds_full, 8);
    mbar_init(dv_done, 1);
    mbar_init(dk_done, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));
  auto tS = [&](int u) { return tmem_base + u * 128; };        // S^T buffer u: 64 columns
  auto tdP = [&](int u) { return tmem_base + u * 128 + 64; };  // dP^T buffer u
  const uint32_t tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0 && active) {
      mbar_expect_tx(kv_full, 2 * DQ_T128);
      tma_load_3d(sK, &P.tmK128, kv_full, 0, kv0, bh);
      tma_load_3d(sK + 16384, &P.tmK128, kv_full, 64, kv0, bh);
      tma_load_3d(sV, &P.tmV128, kv_full, 0, kv0, bh);
      tma_load_3d(sV + 16384, &P.tmV128, kv_full, 64, kv0, bh);
      for (int i = 0; i < n_q; ++i) {
        const int st = i % KV_STAGES;
        if (i >= KV_STAGES) mbar_wait(q_empty(st), ((i / KV_STAGES) - 1) & 1);
        mbar_expect_tx(q_full(st), 2 * DQ_T64);
        tma_load_3d(sQ(st), &P.tmQ64, q_full(st), 0, i * 64, bh);
        tma_load_3d(sQ(st) + 8192, &P.tmQ64, q_full(st), 64, i * 64, bh);
        tma_load_3d(sdO(st), &P.tmdO64, q_full(st), 0, i * 64, bh);
        tma_load_3d(sdO(st) + 8192, &P.tmdO64, q_full(st), 64, i * 64, bh);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0 && active) {
      constexpr uint32_t id_s = idesc_bf16(128, 64, 0, 0);   // S^T = K Q_i^T, dP^T = V dO_i^T : A K-major [128 keys], B K-major [64 queries]
      constexpr uint32_t id_a = idesc_bf16(128, 128, 0, 1);  // dV += P^T dO_i, dK += dS^T Q_i : A K-major [128 x 64], B MN-major [64 x 128]
      constexpr uint32_t id_q = idesc_bf16(128, 64, 1, 1);   // dQ_i^T = K^T dS^T : A = K read MN-major [128 d x 128 keys], B = dS^T MN-major [128 keys x 64 q]
      auto issue_s = [&](int i) {
        const int st = i % KV_STAGES, u = i & 1;
        mbar_wait(q_full(st), (i / KV_STAGES) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tS(u), sdesc_sw128(sK + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    sdesc_sw128(sQ(st) + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), id_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tdP(u), sdesc_sw128(sV + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    sdesc_sw128(sdO(st) + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), id_s, k != 0);
        umma_commit(s_full(u));
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i % KV_STAGES;
        // the scores of tile i+1 are issued BEFORE the accumulations of tile i (they run under tile i's exponentials);
        // TMEM buffer (i+1)&1 was drained by tile i-1, whose ds_full this thread waited for in the previous iteration
        if (i + 1 < n_q) {
          if (i >= 1) mbar_wait(dq_drained((i + 1) & 1), ((i - 1) >> 1) & 1);  // dQ^T_{i-1} sat in the S^T columns of that buffer
          issue_s(i + 1);
        }
        mbar_wait(p_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdV, sdesc_sw128(sPt + k * 32, 16, 1024), sdesc_sw128(sdO(st) + k * 2048, 8192, 1024), id_a, (i | k) != 0);
        umma_commit(dv_done);
        mbar_wait(ds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdK, sdesc_sw128(sdSt + k * 32, 16, 1024), sdesc_sw128(sQ(st) + k * 2048, 8192, 1024), id_a, (i | k) != 0);
        // dQ_i^T [128 d x 64 q] = K^T dS^T over the 128 keys, into the (consumed) S^T columns of this tile's buffer
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tS(i & 1), sdesc_sw128(sK + k * 2048, 16384, 1024), sdesc_sw128(sdSt + k * 2048, 8192, 1024), id_q, k != 0);
        umma_commit(dk_done);  // the dS^T buffer has been read by both contractions
        umma_commit(dq_full(i & 1));
        umma_commit(q_empty(st));
      }
      umma_commit(acc_full);
    }
  } else if (warp == 2) {
    // ================================================================= L / delta stager: the compute threads own KEY rows here, so the
    // per-query L and delta are per-COLUMN values, broadcast-read from smem
    if (active) {
      for (int i = 0; i < n_q; ++i) {
        const int u = i % KV_LD;
        if (i >= KV_LD) mbar_wait(ld_empty(u), ((i / KV_LD) - 1) & 1);
        for (int e = lane; e < 64; e += 32) {
          const int q = i * 64 + e;
          const bool ok = q < P.S;
          ld_gen[u * 128 + e] = ok ? P.lse[(int64_t)bh * P.S + q] : INFINITY;
          ld_gen[u * 128 + 64 + e] = ok ? P.delta[(int64_t)bh * P.S + q] * P.scale : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ld_full(u));  // release: orders the smem writes above before the consumers' acquire
      }
    }
  } else if (warp >= 12) {
    // ================================================================= dQ drain: thread = head-dim index d (TMEM lane), 64 query columns.
    // Lanes of a warp are 32 consecutive d of one query row -> every red.global.add.f32 instruction covers one full 128-byte line.
    if (active) {
      const int quad = warp & 3;
      const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
      float* base = P.dQ + (int64_t)bh * P.S * 128 + quad * 32 + lane;
      for (int i = 0; i < n_q; ++i) {
        const int u = i & 1;
        mbar_wait(dq_full(u), (i >> 1) & 1);
        tc_fence_after();
        uint32_t r[64];
        tmem_ld32(tS(u) + lane_off, r);
        tmem_ld32(tS(u) + lane_off + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_drained(u));
        const int q0 = i * 64;
        const int nq = min(64, P.S - q0);
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < nq) atomicAdd(base + (int64_t)(q0 + j) * 128, __uint_as_float(r[j]));
      }
    }
  } else if (warp >= 4) {
    // ================================================================= compute: thread = key row x 32 query columns
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = ((uint32_t)(quad * 32) << 16) + half * 32;
    const int key = kv0 + row;
    const bool key_ok = key < kv_len && !(key >= txt_len && key < P.split);
    const bool all_keys_ok = kv0 + 128 <= kv_len && !(txt_len < P.split && kv0 + 128 > txt_len && kv0 < P.split);
    if (active) {
      for (int i = 0; i < n_q; ++i) {
        const int u = i & 1, ul = i % KV_LD;
        mbar_wait(ld_full(ul), (i / KV_LD) & 1);
        const float4* L4 = reinterpret_cast<const float4*>(ld_gen + ul * 128 + half * 32);
        const float4* D4 = reinterpret_cast<const float4*>(ld_gen + ul * 128 + 64 + half * 32);
        mbar_wait(s_full(u), (i >> 1) & 1);
        tc_fence_after();
        float p[32];
        uint32_t pk[16];
        {
          uint32_t rs[32];
          tmem_ld32(tS(u) + lane_off, rs);
          tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 l = L4[v];
            p[4 * v] = exp2f(__uint_as_float(rs[4 * v]) * P.scale_log2 - l.x);  // L = +inf for queries past S -> 0
            p[4 * v + 1] = exp2f(__uint_as_float(rs[4 * v + 1]) * P.scale_log2 - l.y);
            p[4 * v + 2] = exp2f(__uint_as_float(rs[4 * v + 2]) * P.scale_log2 - l.z);
            p[4 * v + 3] = exp2f(__uint_as_float(rs[4 * v + 3]) * P.scale_log2 - l.w);
          }
          if (!all_keys_ok) {
#pragma unroll
            for (int e = 0; e < 32; ++e) p[e] = key_ok ? p[e] : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(p[2 * e], p[2 * e + 1]);
        }
        if (i > 0) mbar_wait(dv_done, (i - 1) & 1);  // P^T buffer consumed by the dV MMA of tile i-1
        {
          const uint32_t prow = sPt + row * 128;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t chunk = (uint32_t)(half * 4 + v) ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + chunk * 16), "r"(pk[4 * v]), "r"(pk[4 * v + 1]),
                         "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        {
          uint32_t rp[32];
          tmem_ld32(tdP(u) + lane_off, rp);
          tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 d = D4[v];
            const float a0 = p[4 * v] * (__uint_as_float(rp[4 * v]) * P.scale - d.x);
            const float a1 = p[4 * v + 1] * (__uint_as_float(rp[4 * v + 1]) * P.scale - d.y);
            const float a2 = p[4 * v + 2] * (__uint_as_float(rp[4 * v + 2]) * P.scale - d.z);
            const float a3 = p[4 * v + 3] * (__uint_as_float(rp[4 * v + 3]) * P.scale - d.w);
            pk[2 * v] = pack_bf16(a0, a1);
            pk[2 * v + 1] = pack_bf16(a2, a3);
          }
        }
        if (i > 0) mbar_wait(dk_done, (i - 1) & 1);  // dS^T buffer consumed by the dK MMA of tile i-1
        {
          const uint32_t drow = sdSt + row * 128;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t chunk = (uint32_t)(half * 4 + v) ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(drow + chunk * 16), "r"(pk[4 * v]), "r"(pk[4 * v + 1]),
                         "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(ds_full);
          mbar_arrive(ld_empty(ul));
        }
      }
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    // epilogue: warps of column half 0 store dV, half 1 store dK (both [128 keys x 128] fp32 accumulators -> bf16)
    {
      bf16* dst = (half ? P.dK : P.dV) + ((int64_t)bh * P.S + key) * 128;
      const uint32_t t = (half ? tdK : tdV) + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        if (active) {  // CTA-uniform; tcgen05.ld is warp-collective, so the row bound is applied to the store only
          tmem_ld32(t + c, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = 0;
        }
        if (key < P.S) {
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
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int tmap3(CUtensorMap* m, const void* base, int BH, int S, uint32_t rows) {
  uint64_t dims[3] = {128, (uint64_t)S, (uint64_t)BH};
  uint64_t strides[2] = {128 * 2, (uint64_t)S * 128 * 2};
  uint32_t box[3] = {64, rows, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

int attn_bwd_transposed(const void* Q, const void* K, const void* V, const void* dO, const float* lse, const float* delta, float* dQ,
                        void* dK, void* dV, const int* kv_len, const int* txt_len, int split, int B, int H, int S, float softmax_scale,
                        cudaStream_t stream) {
  AttnBwd2Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = tmap3(&P.tmK128, K, B * H, S, 128)) || (rc = tmap3(&P.tmV128, V, B * H, S, 128)) || (rc = tmap3(&P.tmQ64, Q, B * H, S, 64)) ||
      (rc = tmap3(&P.tmdO64, dO, B * H, S, 64)))
    return rc;
  P.dQ = dQ; P.dK = (bf16*)dK; P.dV = (bf16*)dV; P.lse = lse; P.delta = delta; P.kv_len = kv_len; P.txt_len = txt_len;
  P.split = txt_len ? split : 0; P.S = S; P.H = H; P.scale = softmax_scale; P.scale_log2 = softmax_scale * 1.4426950408889634f;
  static bool attr_done = false;
  if (!attr_done) {
    QFX_CUDA(cudaFuncSetAttribute(attn_bwd_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM));
    attr_done = true;
  }
  dim3 grid((S + 127) / 128, B * H);
  attn_bwd_t_kernel<<<grid, KV_THREADS, KV_SMEM, stream>>>(P);
  QFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qfx
