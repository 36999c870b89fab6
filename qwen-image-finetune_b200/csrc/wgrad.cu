// LoRA weight gradients on tcgen05:  G += A^T B  reduced over the token dimension M (split across CTAs).
//   dB[out, r] = dY^T (s X A^T)        dA[r, in] = (s dY B)^T X         (autograd of peft's LoRA Linear,
//   injected at /root/reference/src/qflux/trainer/base_trainer.py:929-941)
// Both operands are row-major [M, cols] activations, i.e. the reduction dimension is the OUTER one — exactly the
// MN-major UMMA operand layout — so TMA tiles of the tensors as they lie in HBM feed the tensor core directly:
//   A tile [64 rows(m) x 128 cols] = 2 swizzle atoms   -> MMA "M" = 128 A-columns
//   B tile [64 rows(m) x 64*NG cols] = NG atoms        -> MMA "N" = 64*NG (LoRA rank padded to 64 per group)
// mode 0 (dA of a fused q|k|v site, or any single site): every A-column tile pairs with all NG groups of B.
// mode 1 (dB of a fused site): A columns [g*Dg, (g+1)*Dg) pair with B group g only.
// The [128 x 64*NG] fp32 accumulator lives in TMEM; the epilogue adds the first r columns of each group into the flat
// fp32 LoRA-gradient buffer with red.global.add.f32 (a few thousand per CTA).
#include <string.h>

#include "../../include/qfx.h"
#include "host_common.h"
#include "sm100.cuh"

namespace qfx {

constexpr int WG_MAXG = 3;

struct WgradParams {
  CUtensorMap tmA, tmB;
  float* G[WG_MAXG];
  int64_t gs_i, gs_j;
  int M, Na, ng, mode, Dg, r, rows_per_split;
};

template <int NG>
__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradParams P) {
  constexpr int A_BYTES = 2 * 8192, B_BYTES = NG * 8192, STAGE = A_BYTES + B_BYTES, STAGES = 4;
  constexpr int TCOLS = NG == 1 ? 64 : NG == 2 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int i0 = blockIdx.x * 128;                       // first A column of this tile
  const int m_begin = blockIdx.y * P.rows_per_split;
  const int m_end = min(P.M, m_begin + P.rows_per_split);
  const int nkb = (m_end - m_begin + 63) / 64;
  const int grp = P.mode == 1 ? i0 / P.Dg : 0;           // mode 1: the single B group this tile pairs with
  const int b_col0 = P.mode == 1 ? grp * 64 : 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&P.tmA);
      tma_prefetch_desc(&P.tmB);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(empty_bar(s), ((kb / STAGES) - 1) & 1);
        const uint32_t a_dst = smem_base + s * STAGE, b_dst = a_dst + A_BYTES;
        mbar_expect_tx(full_bar(s), STAGE);
        const int m0 = m_begin + kb * 64;
        tma_load_2d(a_dst, &P.tmA, full_bar(s), i0, m0);
        tma_load_2d(a_dst + 8192, &P.tmA, full_bar(s), i0 + 64, m0);
#pragma unroll
        for (int g = 0; g < NG; ++g) tma_load_2d(b_dst + g * 8192, &P.tmB, full_bar(s), b_col0 + g * 64, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128, 64 * NG, 1, 1);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(full_bar(s), (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * STAGE, b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, sdesc_sw128(a_addr + k * 2048, 8192, 1024), sdesc_sw128(b_addr + k * 2048, 8192, 1024), idesc,
                    (kb | k) != 0);
        umma_commit(empty_bar(s));
      }
      umma_commit(acc_bar);
    }
  } else {
    const int quad = warp & 3;
    const int i = i0 + quad * 32 + lane;  // A column == accumulator row
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        if (c >= P.r) break;
        uint32_t v[32];
        tmem_ld32(t_row + g * 64 + c, v);
        tmem_ld_wait();
        if (i < P.Na && nkb > 0) {
          const int gg = P.mode == 1 ? grp : g;
          const int64_t ii = P.mode == 1 ? (int64_t)(i - grp * P.Dg) : (int64_t)i;
          float* dst = P.G[gg] + ii * P.gs_i;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c + e < P.r) atomicAdd(dst + (int64_t)(c + e) * P.gs_j, __uint_as_float(v[e]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

template <int NG>
static int launch_wgrad(const WgradParams& P, dim3 grid, cudaStream_t st) {
  constexpr int SMEM = 4 * (2 * 8192 + NG * 8192) + 1024 + 256;
  static bool done = false;
  if (!done) {
    QFX_CUDA(cudaFuncSetAttribute(wgrad_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    done = true;
  }
  wgrad_kernel<NG><<<grid, 192, SMEM, st>>>(P);
  QFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qfx

using namespace qfx;

/* A [M, Na] (lda), B [M, 64*G] (ldb), both bf16 row-major.  mode 0: for every group g < G and j < r:
 *   Gp[g][i*gs_i + j*gs_j] += sum_m A[m,i] * B[m, g*64 + j]                                  (i < Na)
 * mode 1 (Na = G*Dg): Gp[g][(i - g*Dg)*gs_i + j*gs_j] += sum_m A[m,i] * B[m, g*64 + j]   for g = i / Dg.        */
extern "C" int qfx_lora_wgrad_tc(const void* A, int64_t lda, int Na, const void* B, int64_t ldb, int G, int M, int mode, int Dg,
                                 float* const* Gp, int64_t gs_i, int64_t gs_j, int r, void* stream) {
  QFX_CHECK_ARG(G >= 1 && G <= WG_MAXG && r >= 1 && r <= 64 && M > 0 && Na % 128 == 0, "qfx_lora_wgrad_tc: bad shape");
  QFX_CHECK_ARG(mode == 0 || (mode == 1 && Dg % 128 == 0 && Na == G * Dg), "qfx_lora_wgrad_tc: bad mode/Dg");
  QFX_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "qfx_lora_wgrad_tc: leading dims must be multiples of 8");
  WgradParams P;
  memset(&P, 0, sizeof(P));
  {
    uint64_t dims[2] = {(uint64_t)Na, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {64, 64};
    int rc = make_tmap_bf16(&P.tmA, A, 2, dims, str, box);
    if (rc) return rc;
    uint64_t dimsb[2] = {(uint64_t)(64 * G), (uint64_t)M};
    uint64_t strb[1] = {(uint64_t)ldb * 2};
    rc = make_tmap_bf16(&P.tmB, B, 2, dimsb, strb, box);
    if (rc) return rc;
  }
  for (int g = 0; g < G; ++g) P.G[g] = Gp[g];
  P.gs_i = gs_i; P.gs_j = gs_j; P.M = M; P.Na = Na; P.ng = G; P.mode = mode; P.Dg = Dg; P.r = r;
  const int tiles = Na / 128;
  int splits = (2 * num_sms() + tiles - 1) / tiles;
  const int max_splits = (M + 255) / 256;  // at least 4 k-blocks per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows = ((M + splits - 1) / splits + 63) / 64 * 64;
  P.rows_per_split = rows;
  dim3 grid(tiles, (M + rows - 1) / rows);
  const int ng = mode == 1 ? 1 : G;
  cudaStream_t st = (cudaStream_t)stream;
  switch (ng) {
    case 1: return launch_wgrad<1>(P, grid, st);
    case 2: return launch_wgrad<2>(P, grid, st);
    default: return launch_wgrad<3>(P, grid, st);
  }
}
