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
  int S, H;
  float scale, scale_log2;
};

constexpr int BW_SMEM = 6 * BW_TILE + 1024 + 256;  // K, V, Q, dO, P, dS

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(320, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = smem_base + BW_TILE, sQ = smem_base + 2 * BW_TILE, sdO = smem_base + 3 * BW_TILE;
  const uint32_t sP = smem_base + 4 * BW_TILE, sdS = smem_base + 5 * BW_TILE;
  const uint32_t bar_base = smem_base + 6 * BW_TILE;
  const uint32_t kv_full = bar_base, qdo_full = bar_base + 8, qdo_empty = bar_base + 16, s_full = bar_base + 24;
  const uint32_t pds_full = bar_base + 32, dq_full = bar_base + 40, dq_empty = bar_base + 48, acc_full = bar_base + 56;
  const uint32_t tmem_slot = bar_base + 64;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int b = bh / P.H;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int n_q = (P.S + 127) / 128;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  // a fully masked key tile (beyond kv_len, or entirely inside the text padding) contributes nothing: write zeros and leave
  const bool active = kv0 < kv_len && !(kv0 >= txt_len && kv0 + 128 <= P.split);

  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    mbar_init(qdo_full, 1);
    mbar_init(qdo_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0 && active) {
      mbar_expect_tx(kv_full, 2 * BW_TILE);
      tma_load_3d(sK, &P.tmK, kv_full, 0, kv0, bh);
      tma_load_3d(sK + BW_ATOM, &P.tmK, kv_full, 64, kv0, bh);
      tma_load_3d(sV, &P.tmV, kv_full, 0, kv0, bh);
      tma_load_3d(sV + BW_ATOM, &P.tmV, kv_full, 64, kv0, bh);
      for (int i = 0; i < n_q; ++i) {
        if (i > 0) mbar_wait(qdo_empty, (i - 1) & 1);
        mbar_expect_tx(qdo_full, 2 * BW_TILE);
        tma_load_3d(sQ, &P.tmQ, qdo_full, 0, i * 128, bh);
        tma_load_3d(sQ + BW_ATOM, &P.tmQ, qdo_full, 64, i * 128, bh);
        tma_load_3d(sdO, &P.tmdO, qdo_full, 0, i * 128, bh);
        tma_load_3d(sdO + BW_ATOM, &P.tmdO, qdo_full, 64, i * 128, bh);
      }
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
        mbar_wait(qdo_full, i & 1);
        if (i > 0) mbar_wait(dq_empty, (i - 1) & 1);  // S columns (shared with dQ) drained, P/dS smem free
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tS, kmaj(sQ, k), kmaj(sK, k), id_kk, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdP, kmaj(sdO, k), kmaj(sV, k), id_kk, k != 0);
        umma_commit(s_full);
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdV, mnmaj(sP, k), mnmaj(sdO, k), id_mm, (i | k) != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdK, mnmaj(sdS, k), mnmaj(sQ, k), id_mm, (i | k) != 0);
        umma_commit(qdo_empty);  // Q_i / dO_i no longer needed: the next tile may load while dQ is computed and drained
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tS, kmaj(sdS, k), mnmaj(sK, k), id_km, k != 0);
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
      for (int i = 0; i < n_q; ++i) {
        const int q = i * 128 + row;
        const bool q_ok = q < P.S;
        const float L = q_ok ? P.lse[(int64_t)bh * P.S + q] : INFINITY;
        const float dl = q_ok ? P.delta[(int64_t)bh * P.S + q] : 0.f;
        mbar_wait(s_full, i & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = c0; c < c0 + 64; c += 32) {
          uint32_t rs[32], rp[32];
          tmem_ld32(tS + lane_off + c, rs);
          tmem_ld32(tdP + lane_off + c, rp);
          tmem_ld_wait();
          uint32_t pk[16], dk[16];
          const float dls = dl * P.scale;
          if (full_tile) {  // warp-uniform fast path: no key masking
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float p0 = exp2f(__uint_as_float(rs[2 * e]) * P.scale_log2 - L);
              const float p1 = exp2f(__uint_as_float(rs[2 * e + 1]) * P.scale_log2 - L);
              pk[e] = pack_bf16(p0, p1);
              dk[e] = pack_bf16(p0 * (__uint_as_float(rp[2 * e]) * P.scale - dls), p1 * (__uint_as_float(rp[2 * e + 1]) * P.scale - dls));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int k0 = c + 2 * e, k1 = k0 + 1;
              const float p0 = (k0 < valid && !(k0 >= gap0 && k0 < gap1)) ? exp2f(__uint_as_float(rs[2 * e]) * P.scale_log2 - L) : 0.f;
              const float p1 = (k1 < valid && !(k1 >= gap0 && k1 < gap1)) ? exp2f(__uint_as_float(rs[2 * e + 1]) * P.scale_log2 - L) : 0.f;
              pk[e] = pack_bf16(p0, p1);
              dk[e] = pack_bf16(p0 * (__uint_as_float(rp[2 * e]) * P.scale - dls), p1 * (__uint_as_float(rp[2 * e + 1]) * P.scale - dls));
            }
          }
          const uint32_t off = (uint32_t)row * 128 + (c >> 6) * BW_ATOM;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t chunk = (uint32_t)(((c & 63) >> 3) + v) ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off + chunk * 16), "r"(pk[4 * v]),
                         "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS + off + chunk * 16), "r"(dk[4 * v]),
                         "r"(dk[4 * v + 1]), "r"(dk[4 * v + 2]), "r"(dk[4 * v + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);

        // ---- drain dQ_i: TMEM -> fp32 swizzled slabs in the (now idle) P|dS buffers -> TMA reduce-add to global
        mbar_wait(dq_full, i & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = c0; c < c0 + 64; c += 32) {
          uint32_t r[32];
          tmem_ld32(tS + lane_off + c, r);
          tmem_ld_wait();
          const uint32_t slab = sP + (c >> 5) * BW_ATOM + (uint32_t)row * 128;  // [128 rows x 32 fp32], 128 B rows
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint32_t chunk = (uint32_t)v ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + chunk * 16), "r"(r[4 * v]), "r"(r[4 * v + 1]),
                         "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) tma_reduce_add_3d(&P.tmdQ, sP + sl * BW_ATOM, sl * 32, i * 128, bh);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(dq_empty);
        }
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

/* All of Q, K, V, dO, dK, dV: [B, H, S, 128] bf16; dQ_accum: [B, H, S, 128] fp32, MUST be zeroed by the caller (it is the
 * target of TMA reduce-adds from every key tile).  lse: log2-domain logsumexp from qfx_attn_fwd; delta = rowsum(dO*O). */
extern "C" int qfx_attn_bwd(const void* Q, const void* K, const void* V, const void* dO, const float* lse, const float* delta,
                            float* dQ_accum, void* dK, void* dV, const int* kv_len, const int* txt_len, int split, int B, int H,
                            int S, float softmax_scale, void* stream) {
  QFX_CHECK_ARG(B > 0 && H > 0 && S > 0 && dQ_accum && dK && dV && lse && delta, "qfx_attn_bwd: bad arguments");
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
