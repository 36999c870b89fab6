// Joint attention backward for sm_100a (head_dim 128) — software-pipelined transposed formulation (round 2 default).
// Autograd backward of F.scaled_dot_product_attention as the reference runs it
// (/root/reference/src/qflux/models/transformer_qwenimage.py:329-337 under loss.backward(), base_trainer.py:528).
//
// Why a third kernel: ncu on the round-1 kernels (profiles/r01_ncu_full_attention_bwd.md) shows the tensor pipe 33 % busy with the
// tensor-core shared-memory pipe saturated whenever it is busy, and the compute warps parked on mbarriers the rest of the time: both
// earlier variants move ~0.5 MB of shared-memory operand traffic per 128 x 128 tile pair (every contraction SS-mode, P and dS through
// shared memory, dQ staged for a TMA reduce) and run the five contractions and the two register passes as a serial chain.  Here:
//   * scores are TRANSPOSED in tensor memory (lane = key row), 128 keys x 128 queries per step:
//         S^T  = K Q_i^T               SS   A = K (K-major)            B = Q_i (K-major)                     -> tS   [128 cols]
//         dP^T = V dO_i^T              SS   A = V                      B = dO_i                              -> tdP  [128 cols]
//         P^T  = exp2(S^T c - L_q)     8 compute warps, thread = key row x 64 query columns; bf16 P^T is written back over the S^T
//                                      columns it came from and feeds the next contraction FROM TENSOR MEMORY (no smem round trip)
//         dV  += P^T dO_i              TS   A = P^T (tensor memory)    B = dO_i read MN-major                -> tdV  [128 cols]
//         dS^T = P^T o (dP^T - delta_q) c  -> bf16 in shared memory (it is a B operand below, which tensor memory cannot supply)
//         dK  += dS^T Q_i              SS   A = dS^T (K-major)         B = Q_i read MN-major                 -> tdK  [128 cols]
//         dQ_i^T = K^T dS^T            SS   A = K read MN-major        B = dS^T read MN-major                -> over tdP (consumed)
//     dQ_i^T has lane = head-dim index, so 4 drain warps add it to the fp32 dQ with red.global.add.f32 whose 32 lanes cover one full
//     128-byte line of a query row: no shared-memory staging, no TMA reduce.  Operand traffic: 288 KB + 32 KB of dS^T stores per pair.
//   * the single MMA thread issues in the order  S_i | dQ_{i-1}, dK_{i-1} | dP_i | dV_i :  the P pass of tile i runs under
//     dQ/dK of tile i-1 and dP of tile i, the dS pass of tile i under dV_i and S_{i+1} — S^T/P^T and dP^T/dQ^T each need only ONE
//     tensor-memory buffer for that (512 columns: dK, dV, dP^T|dQ^T, S^T|P^T), which is what leaves room for 128-query tiles.
// Keys >= kv_len[b] and text padding [txt_len[b], split) are masked per key ROW; queries past S get L = +inf (P = 0).
#include <stdlib.h>
#include <string.h>

#include "../../include/qfx.h"
#include "host_common.h"
#include "sm100.cuh"

namespace qfx {

struct AttnBwd3Params {
  CUtensorMap tmQ, tmK, tmV, tmdO;  // bf16 [B*H, S, 128], box {64, 128, 1}
  float* dQ;                        // [B*H, S, 128] fp32, zeroed by the caller: every key tile adds its partial dQ
  bf16* dK;
  bf16* dV;
  const float* lse;    // log2 domain
  const float* delta;
  const int* kv_len;
  const int* txt_len;
  int split, S, H;
  float scale, scale_log2;
  int rev;         // 1: heads are visited in descending order, i.e. most recently zero-filled dQ lines (still in L2) first (A/B)
  int rotate;      // 1: rotate the query-tile order per key tile (default), 0: all key tiles walk query tiles 0..n-1 (A/B)
  long long* dbg;  // optional clock64 stamps of CTA (1, 0): [iteration][16] (tools/attn_timeline3.py)
};

// warp-converged issue with split descriptors (sm100.cuh: sdesc_lo / umma_ss_w / umma_ts_w)
__device__ __forceinline__ uint32_t b3_lo_kmaj(uint32_t base) { return sdesc_lo(base, 16); }
__device__ __forceinline__ uint32_t b3_lo_mnmaj(uint32_t base) { return sdesc_lo(base, 16384); }
__host__ __device__ constexpr uint32_t b3_inc_kmaj(int k) { return (uint32_t)((k >> 2) * 1024 + (k & 3) * 2); }  // (k>>2) * 16 KB atom + (k&3) * 32 B, in 16-B units
__host__ __device__ constexpr uint32_t b3_inc_mnmaj(int k) { return (uint32_t)(k * 128); }                        // 16 rows of 128 B
__device__ __forceinline__ void b3_mma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) { umma_ss_w(d, a_lo, b_lo, idesc, acc); }
__device__ __forceinline__ void b3_mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) { umma_ts_w(d, a_tmem, b_lo, idesc, acc); }

__device__ __forceinline__ uint32_t b3_mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

constexpr int B3_TILE = 128 * 128 * 2;  // 32 KB: [128 rows x 128] bf16, two 64-column swizzle atoms of 16 KB
constexpr int B3_ATOM = 128 * 128;      // 16 KB
constexpr int B3_RING = 2;              // L / delta ring depth (tile i -> slot i & 1)
constexpr int B3_SMEM = 7 * B3_TILE + B3_RING * 1024 + 256;  // K, V, Q x2, dO x2, dS^T, L/delta ring, barriers  (231,680 B)
constexpr int B3_THREADS = 640;  // warp 0 TMA, 1 MMA, 2 L/delta stager, 3 idle, 4-11 compute (quad = warp & 3, column half = (warp - 4) >> 2),
                                 // 12-19 dQ drain (quad = warp & 3, column half = (warp - 12) >> 2).  EIGHT drain warps: a warp sustains one
                                 // 128-byte red.global per ~40 clk (tools/experiments/red_bw.cu: 4 warps/SM reach 3.7 TB/s chip-wide, 16 warps
                                 // 5.9), and a tile pair produces 512 of them — with four warps the drain set the pace of the whole kernel

template <bool WARP_ISSUE>
__global__ void __launch_bounds__(B3_THREADS, 1) attn_bwd3_kernel(const __grid_constant__ AttnBwd3Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t sK = smem_base, sV = smem_base + B3_TILE;
  auto sQ = [&](int st) { return smem_base + (2 + st) * B3_TILE; };
  auto sdO = [&](int st) { return smem_base + (4 + st) * B3_TILE; };
  const uint32_t sdS = smem_base + 6 * B3_TILE;
  const uint32_t sLD = smem_base + 7 * B3_TILE;  // float [B3_RING][2 (L, delta * scale)][128]
  const uint32_t bar_base = sLD + B3_RING * 1024;
  const uint32_t kv_full = bar_base;
  auto q_full = [&](int st) { return bar_base + 8u * (1 + st); };
  auto q_empty = [&](int st) { return bar_base + 8u * (3 + st); };
  auto do_full = [&](int st) { return bar_base + 8u * (5 + st); };
  auto do_empty = [&](int st) { return bar_base + 8u * (7 + st); };
  auto ld_full = [&](int u) { return bar_base + 8u * (9 + u); };
  auto ld_empty = [&](int u) { return bar_base + 8u * (11 + u); };
  const uint32_t s_full = bar_base + 8u * 13, dp_full = bar_base + 8u * 14, p_full = bar_base + 8u * 15, ds_full = bar_base + 8u * 16;
  const uint32_t ds_free = bar_base + 8u * 17, dq_full = bar_base + 8u * 18, dq_drained = bar_base + 8u * 19, acc_full = bar_base + 8u * 20;
  const uint32_t tmem_slot = bar_base + 8u * 21;
  float* ld_gen = reinterpret_cast<float*>(smem_raw + (sLD - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = P.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int b = bh / P.H;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  // query tiles that hold only padding rows of a pad-to-max batch (positions >= kv_len[b]) carry dO = 0: nothing to add, skip them
  const int n_q = ((kv_len < P.S ? kv_len : P.S) + 127) / 128;
  // a fully masked key tile (beyond kv_len, or entirely inside the text padding) contributes nothing: write zeros and leave
  const bool active = kv0 < kv_len && !(kv0 >= txt_len && kv0 + 128 <= P.split);
  // Query tiles are visited in a per-key-tile ROTATED order: the 19 key-tile CTAs of a head run concurrently, and with a common order all
  // of them would add into the same dQ rows at the same time (the L2 atomic unit serialises per address)
  const int rot = P.rotate ? (int)(blockIdx.x % (unsigned)n_q) : 0;
  auto qt = [&](int i) { int t = i + rot; return t >= n_q ? t - n_q : t; };
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0;
#define B3_DBG(i, e) do { if (dbg_on) P.dbg[(i) * 16 + (e)] = clock64(); } while (0)

  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int st = 0; st < 2; ++st) {
      mbar_init(q_full(st), 1);
      mbar_init(q_empty(st), 1);
      mbar_init(do_full(st), 1);
      mbar_init(do_empty(st), 1);
      mbar_init(ld_full(st), 1);
      mbar_init(ld_empty(st), 8);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_full, 8);
    mbar_init(ds_full, 8);
    mbar_init(ds_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_drained, 8);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const uint32_t tdK = tmem_base, tdV = tmem_base + 128, tdP = tmem_base + 256, tS = tmem_base + 384;

  if (warp == 0) {
    // ================================================================= TMA producers: lane 0 = K, V, Q_i ; lane 1 = dO_i
    if (lane == 0 && active) {
      mbar_expect_tx(kv_full, 2 * B3_TILE);
      tma_load_3d(sK, &P.tmK, kv_full, 0, kv0, bh);
      tma_load_3d(sK + B3_ATOM, &P.tmK, kv_full, 64, kv0, bh);
      tma_load_3d(sV, &P.tmV, kv_full, 0, kv0, bh);
      tma_load_3d(sV + B3_ATOM, &P.tmV, kv_full, 64, kv0, bh);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        if (i >= 2) mbar_wait(q_empty(st), ((i - 2) >> 1) & 1);  // dK_{i-2} has consumed Q_{i-2}
        mbar_expect_tx(q_full(st), B3_TILE);
        tma_load_3d(sQ(st), &P.tmQ, q_full(st), 0, qt(i) * 128, bh);
        tma_load_3d(sQ(st) + B3_ATOM, &P.tmQ, q_full(st), 64, qt(i) * 128, bh);
      }
    } else if (lane == 1 && active) {
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        if (i >= 2) mbar_wait(do_empty(st), ((i - 2) >> 1) & 1);  // dV_{i-2} has consumed dO_{i-2}
        mbar_expect_tx(do_full(st), B3_TILE);
        tma_load_3d(sdO(st), &P.tmdO, do_full(st), 0, qt(i) * 128, bh);
        tma_load_3d(sdO(st) + B3_ATOM, &P.tmdO, do_full(st), 64, qt(i) * 128, bh);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (WARP_ISSUE) {
      // the whole warp runs the issue loop with identical values; an elected lane issues each tcgen05 instruction
      if (active) {
        constexpr uint32_t id_kk = idesc_bf16(128, 128, 0, 0), id_km = idesc_bf16(128, 128, 0, 1), id_mm = idesc_bf16(128, 128, 1, 1);
        const uint32_t kK = b3_lo_kmaj(sK), kV = b3_lo_kmaj(sV), mK = b3_lo_mnmaj(sK), kdS = b3_lo_kmaj(sdS), mdS = b3_lo_mnmaj(sdS);
        auto issue_s = [&](int i) {
          const uint32_t kQ = b3_lo_kmaj(sQ(i & 1));
          mbar_wait(q_full(i & 1), (i >> 1) & 1);
          tc_fence_after();
          if (lane == 0) B3_DBG(i, 0);
#pragma unroll
          for (int k = 0; k < 8; ++k) b3_mma_ss(tS, kK + b3_inc_kmaj(k), kQ + b3_inc_kmaj(k), id_kk, k != 0);
          umma_commit_w(s_full);
          if (lane == 0) B3_DBG(i, 1);
        };
        auto issue_dp = [&](int i) {
          const uint32_t kdO = b3_lo_kmaj(sdO(i & 1));
          mbar_wait(do_full(i & 1), (i >> 1) & 1);
          if (i > 0) mbar_wait(dq_drained, (i - 1) & 1);
          tc_fence_after();
          if (lane == 0) B3_DBG(i, 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) b3_mma_ss(tdP, kV + b3_inc_kmaj(k), kdO + b3_inc_kmaj(k), id_kk, k != 0);
          umma_commit_w(dp_full);
          if (lane == 0) B3_DBG(i, 5);
        };
        auto issue_dv = [&](int i) {
          const uint32_t mdO = b3_lo_mnmaj(sdO(i & 1));
          mbar_wait(p_full, i & 1);
          tc_fence_after();
          if (lane == 0) B3_DBG(i, 6);
#pragma unroll
          for (int k = 0; k < 8; ++k) b3_mma_ts(tdV, tS + k * 8, mdO + b3_inc_mnmaj(k), id_km, (i | k) != 0);
          umma_commit_w(do_empty(i & 1));
          if (lane == 0) B3_DBG(i, 7);
        };
        // (Measured and rejected: issuing dP_{i+1} between two halves of dK_i shortens the dS-pass dependency loop by one dK/2, but Q_i is
        // then released one dP later and the ~1700-clk TMA latency of Q_{i+2} lands on the S_{i+2} issue: 0.757 -> 0.773 ms.)
        auto issue_dq_dk = [&](int i) {
          const uint32_t mQ = b3_lo_mnmaj(sQ(i & 1));
          mbar_wait(ds_full, i & 1);
          tc_fence_after();
          if (lane == 0) B3_DBG(i, 2);
#pragma unroll
          for (int k = 0; k < 8; ++k) b3_mma_ss(tdP, mK + b3_inc_mnmaj(k), mdS + b3_inc_mnmaj(k), id_mm, k != 0);
          umma_commit_w(dq_full);
#pragma unroll
          for (int k = 0; k < 8; ++k) b3_mma_ss(tdK, kdS + b3_inc_kmaj(k), mQ + b3_inc_mnmaj(k), id_km, (i | k) != 0);
          umma_commit_w(q_empty(i & 1));
          umma_commit_w(ds_free);
          if (lane == 0) B3_DBG(i, 3);
        };
        mbar_wait(kv_full, 0);
        issue_s(0);
        issue_dp(0);
        issue_dv(0);
        for (int i = 1; i < n_q; ++i) {
          issue_s(i);          // runs under the dS pass of tile i-1
          issue_dq_dk(i - 1);  // run under the P pass of tile i
          issue_dp(i);
          issue_dv(i);
        }
        issue_dq_dk(n_q - 1);
        umma_commit_w(acc_full);
      }
    } else if (lane == 0 && active) {
      constexpr uint32_t id_kk = idesc_bf16(128, 128, 0, 0);  // A K-major, B K-major
      constexpr uint32_t id_km = idesc_bf16(128, 128, 0, 1);  // A K-major (smem or tensor memory), B MN-major
      constexpr uint32_t id_mm = idesc_bf16(128, 128, 1, 1);  // A MN-major, B MN-major
      auto kmaj = [](uint32_t base, int k) { return sdesc_sw128(base + (k >> 2) * B3_ATOM + (k & 3) * 32, 16, 1024); };
      auto mnmaj = [](uint32_t base, int k) { return sdesc_sw128(base + k * 2048, B3_ATOM, 1024); };
      auto issue_s = [&](int i) {  // S^T_i = K Q_i^T   (tS is free: dV_{i-1}, the last reader of P^T_{i-1}, was issued earlier — MMAs complete in order)
        mbar_wait(q_full(i & 1), (i >> 1) & 1);
        tc_fence_after();
        B3_DBG(i, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tS, kmaj(sK, k), kmaj(sQ(i & 1), k), id_kk, k != 0);
        umma_commit(s_full);
        B3_DBG(i, 1);
      };
      auto issue_dp = [&](int i) {  // dP^T_i = V dO_i^T  (tdP holds dQ^T_{i-1} until the drain warps have read it)
        mbar_wait(do_full(i & 1), (i >> 1) & 1);
        if (i > 0) mbar_wait(dq_drained, (i - 1) & 1);
        tc_fence_after();
        B3_DBG(i, 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdP, kmaj(sV, k), kmaj(sdO(i & 1), k), id_kk, k != 0);
        umma_commit(dp_full);
        B3_DBG(i, 5);
      };
      auto issue_dv = [&](int i) {  // dV += P^T_i dO_i : A = bf16 P^T in the first 64 columns of tS (8 columns per 16 queries)
        mbar_wait(p_full, i & 1);
        tc_fence_after();
        B3_DBG(i, 6);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(tdV, tS + k * 8, mnmaj(sdO(i & 1), k), id_km, (i | k) != 0);
        umma_commit(do_empty(i & 1));
        B3_DBG(i, 7);
      };
      auto issue_dq_dk = [&](int i) {  // dQ^T_i = K^T dS^T_i (over the consumed dP^T_i) ; dK += dS^T_i Q_i
        mbar_wait(ds_full, i & 1);
        tc_fence_after();
        B3_DBG(i, 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdP, mnmaj(sK, k), mnmaj(sdS, k), id_mm, k != 0);
        umma_commit(dq_full);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tdK, kmaj(sdS, k), mnmaj(sQ(i & 1), k), id_km, (i | k) != 0);
        umma_commit(q_empty(i & 1));
        umma_commit(ds_free);
        B3_DBG(i, 3);
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      issue_dp(0);
      issue_dv(0);
      for (int i = 1; i < n_q; ++i) {
        issue_s(i);          // runs under the dS pass of tile i-1
        issue_dq_dk(i - 1);  // run under the P pass of tile i
        issue_dp(i);
        issue_dv(i);
      }
      issue_dq_dk(n_q - 1);
      umma_commit(acc_full);
    }
  } else if (warp == 2) {
    // ================================================================= L / delta stager: the compute threads own KEY rows, so the
    // per-query L and delta are per-COLUMN values, broadcast-read from a small shared-memory ring
    if (active) {
      for (int i = 0; i < n_q; ++i) {
        const int u = i & 1;
        if (i >= 2) mbar_wait(ld_empty(u), ((i - 2) >> 1) & 1);
        for (int e = lane; e < 128; e += 32) {
          const int q = qt(i) * 128 + e;
          const bool ok = q < P.S;
          ld_gen[u * 256 + e] = ok ? P.lse[(int64_t)bh * P.S + q] : INFINITY;
          ld_gen[u * 256 + 128 + e] = ok ? P.delta[(int64_t)bh * P.S + q] * P.scale : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ld_full(u));  // release: orders the smem writes above before the consumers' acquire
      }
    }
  } else if (warp >= 12) {
    // ================================================================= dQ drain: thread = head-dim index d (TMEM lane), 64 query columns.
    // Lanes of a warp are 32 consecutive d of one query row -> every red.global.add.f32 instruction covers one full 128-byte line.
    // The tensor-memory read takes ~100 clk and releases tdP at once; the 64 reds then trickle out under the next tile's math.
    if (active) {
      const int quad = warp & 3, ch = (warp - 12) >> 2;
      const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
      float* base = P.dQ + (int64_t)bh * P.S * 128 + quad * 32 + lane;
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(dq_full, i & 1);
        tc_fence_after();
        if (warp == 12 && lane == 0) B3_DBG(i, 12);
        const int q0 = qt(i) * 128 + ch * 64;
        const int nq = min(64, P.S - q0);
        uint32_t r[64];
        tmem_ld32(tdP + lane_off + ch * 64, r);
        tmem_ld32(tdP + lane_off + ch * 64 + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_drained);  // the tensor core may overwrite tdP with dP^T_{i+1}
        if (warp == 12 && lane == 0) B3_DBG(i, 13);
        float* dst = base + (int64_t)q0 * 128;
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < nq && P.rotate != 2) atomicAdd(dst + j * 128, __uint_as_float(r[j]));
      }
    }
  } else if (warp >= 4) {
    // ================================================================= compute: thread = key row x 64 query columns
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int key = kv0 + row;
    const bool key_ok = key < kv_len && !(key >= txt_len && key < P.split);
    if (active) {
      const uint32_t ds_row = sdS + half * B3_ATOM + row * 128;  // this thread's 64 dS^T columns = one 128-byte swizzle-atom row
      for (int i = 0; i < n_q; ++i) {
        const int u = i & 1;
        mbar_wait(ld_full(u), (i >> 1) & 1);
        const float4* L4 = reinterpret_cast<const float4*>(ld_gen + u * 256 + half * 64);
        const float4* D4 = reinterpret_cast<const float4*>(ld_gen + u * 256 + 128 + half * 64);
        // ---- P pass: P^T = exp2(S^T c - L_q), bf16, back into tensor memory
        mbar_wait(s_full, i & 1);
        tc_fence_after();
        if (warp == 4 && lane == 0) B3_DBG(i, 8);
        uint32_t pk[32];
        {
          // both 32-column loads are in flight before the first exponential (one exposed tensor-memory latency per pass instead of two)
          uint32_t rs[64];
          tmem_ld32(tS + lane_off + half * 64, rs);
          tmem_ld32(tS + lane_off + half * 64 + 32, rs + 32);
          tmem_ld_wait();
          if (warp == 4 && lane == 0) B3_DBG(i, 14);
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            const float4 l = L4[v];  // L = +inf for queries past S -> P = 0
            pk[2 * v] = pack_bf16(exp2f(__uint_as_float(rs[4 * v]) * P.scale_log2 - l.x), exp2f(__uint_as_float(rs[4 * v + 1]) * P.scale_log2 - l.y));
            pk[2 * v + 1] = pack_bf16(exp2f(__uint_as_float(rs[4 * v + 2]) * P.scale_log2 - l.z), exp2f(__uint_as_float(rs[4 * v + 3]) * P.scale_log2 - l.w));
          }
          if (!key_ok) {  // masked key ROW (last key tile / text padding only): no per-element select in the common path
#pragma unroll
            for (int e = 0; e < 32; ++e) pk[e] = 0u;
          }
        }
        // the packed P^T columns of column-half 1 ([32, 64)) lie over S^T columns that column-half 0 reads: both warps of the pair
        // must have their S^T values in registers before either stores
        tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        tc_fence_after();
        tmem_st32(tS + lane_off + half * 32, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (warp == 4 && lane == 0) B3_DBG(i, 9);
        // ---- dS pass: dS^T = P^T o (dP^T c - delta_q c), bf16, to shared memory (B operand of dQ^T, A operand of dK)
        mbar_wait(dp_full, i & 1);
        tc_fence_after();
        if (warp == 4 && lane == 0) B3_DBG(i, 10);
        if (i > 0) mbar_wait(ds_free, (i - 1) & 1);  // dQ^T_{i-1} and dK_{i-1} have read the dS^T buffer (long ago: they ran under the P pass)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t rp[32];
          tmem_ld32(tdP + lane_off + half * 64 + cc * 32, rp);
          tmem_ld_wait();
          if (cc == 0 && warp == 4 && lane == 0) B3_DBG(i, 15);
          uint32_t ds[16];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 d = D4[cc * 8 + v];
            const uint32_t pa = pk[cc * 16 + 2 * v], pb = pk[cc * 16 + 2 * v + 1];
            // dS = P o (dP c - delta c): the bracket in fp32, rounded to bf16 and multiplied by the bf16 P as a packed pair (one HMUL2 for
            // two elements instead of unpack + 2 FMUL + pack; dS is stored in bf16 either way)
            ds[2 * v] = b3_mul_bf16x2(pa, pack_bf16(__uint_as_float(rp[4 * v]) * P.scale - d.x, __uint_as_float(rp[4 * v + 1]) * P.scale - d.y));
            ds[2 * v + 1] = b3_mul_bf16x2(pb, pack_bf16(__uint_as_float(rp[4 * v + 2]) * P.scale - d.z, __uint_as_float(rp[4 * v + 3]) * P.scale - d.w));
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t chunk = (uint32_t)(cc * 4 + v) ^ (uint32_t)(row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + chunk * 16), "r"(ds[4 * v]), "r"(ds[4 * v + 1]),
                         "r"(ds[4 * v + 2]), "r"(ds[4 * v + 3])
                         : "memory");
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(ds_full);
          mbar_arrive(ld_empty(u));
        }
        if (warp == 4 && lane == 0) B3_DBG(i, 11);
      }
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    // epilogue: warps of column half 0 store dV, half 1 store dK (both [128 keys x 128] fp32 accumulators -> bf16)
    if (kv0 < P.S) {
      bf16* dst = (half ? P.dK : P.dV) + ((int64_t)bh * P.S + key) * 128;
      const uint32_t t = (half ? tdK : tdV) + lane_off;
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

int make_qkv_tmap(CUtensorMap* m, const void* base, int BH, int S);
}  // namespace qfx
extern long long* g_qfx_attn_bwd_dbg;
namespace qfx {

int attn_bwd_pipelined(const void* Q, const void* K, const void* V, const void* dO, const float* lse, const float* delta, float* dQ,
                       void* dK, void* dV, const int* kv_len, const int* txt_len, int split, int B, int H, int S, float softmax_scale,
                       cudaStream_t stream) {
  AttnBwd3Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = make_qkv_tmap(&P.tmQ, Q, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmK, K, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmV, V, B * H, S))) return rc;
  if ((rc = make_qkv_tmap(&P.tmdO, dO, B * H, S))) return rc;
  P.dQ = dQ; P.dK = (bf16*)dK; P.dV = (bf16*)dV; P.lse = lse; P.delta = delta; P.kv_len = kv_len; P.txt_len = txt_len;
  P.split = txt_len ? split : 0; P.S = S; P.H = H;
  P.scale = softmax_scale;
  P.scale_log2 = softmax_scale * 1.4426950408889634f;
  P.dbg = g_qfx_attn_bwd_dbg;
  static const bool no_rotate = getenv("QFX_ATTN_BWD3_NO_ROTATE") != nullptr;
  P.rotate = no_rotate ? 0 : 1;
  if (getenv("QFX_ATTN_BWD3_NO_RED")) P.rotate = 2;  // EXPERIMENT ONLY (wrong dQ): how much of the kernel time is the dQ red traffic?
  P.rev = getenv("QFX_ATTN_BWD3_REV") ? atoi(getenv("QFX_ATTN_BWD3_REV")) : 0;
  static const bool lane_issue = getenv("QFX_ATTN_BWD3_LANE_ISSUE") != nullptr;  // A/B: single-lane issue region (lane-serialising loops)
  static bool attr_done = false;
  if (!attr_done) {
    QFX_CUDA(cudaFuncSetAttribute(attn_bwd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B3_SMEM));
    QFX_CUDA(cudaFuncSetAttribute(attn_bwd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B3_SMEM));
    attr_done = true;
  }
  dim3 grid((S + 127) / 128, B * H);
  if (lane_issue) attn_bwd3_kernel<false><<<grid, B3_THREADS, B3_SMEM, stream>>>(P);
  else attn_bwd3_kernel<true><<<grid, B3_THREADS, B3_SMEM, stream>>>(P);
  QFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qfx
