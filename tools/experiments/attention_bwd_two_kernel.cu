// EXPERIMENT, NOT BUILT INTO libqfx_b200.so (kept for the measurement it records; see DESIGN.md "attention backward: what was tried").
// Joint attention backward, two-kernel formulation for sm_100a (head_dim 128).
// Measured on B200 at B=4, H=24, S=2400 (round 1): parity-green on every attn_bwd_* case, but dq_kernel 605 us + dkv_kernel 665 us
// = 1.33 ms against 1.01 ms for the single-kernel backward (csrc/attention_bwd.cu).  Both designs are bound by shared-memory
// bandwidth (SS-mode tcgen05 operands at 128 B/clk/SM), and recomputing S / dP in two kernels with 64-wide tiles reads MORE
// operand bytes per tile pair (7 contractions, N=64 instructions re-read the 128-row A operand twice as often) than it saves
// by dropping the dQ reduce-add.  tools/microbench.cu holds the micro-benchmarks (tcgen05.ld ~900 B/clk/SM, MUFU 15.7 ex2/clk/SM)
// that ruled out TMEM bandwidth and exponentials as the limiter.
// Timeline of the dK/dV kernel (attn_timeline_two_kernel.py, clock64 stamps): the 8 compute warps need ~2000 clk per 128-key x
// 64-query tile (P pass ~750 = MUFU bound for 8192 exponentials, dS pass ~300, and ~900 of fixed cost: four mbarrier polls,
// two st.shared + fence.proxy.async + arrive sequences) while the tensor work of a tile is ~1050 clk: with all compute warps in
// lock-step on one tile the fixed cost is amortised over only 32 elements per thread.  The fix would be two compute groups
// ping-ponging on alternate tiles (16 warps, P^T/dS^T buffers per group); not pursued in round 1.
// To rebuild: copy into qwen-image-finetune_b200/csrc/, declare qfx::attn_bwd_two_kernel in attention_bwd.cu and dispatch to it.
//
// The single-kernel backward (attention_bwd.cu) accumulates dQ across key tiles with 2.2 GB of L2 reduce-adds per call and
// keeps every stage of a (query tile, key tile) pair on one serial chain (~6600 clk per pair against 2560 clk of MMA).
// Splitting it removes the atomics and lets each kernel overlap tensor work with the exp / dS math:
//
//  dq_kernel   one CTA per 128-query tile, loops over 64-key tiles, TWO CTAs per SM (112 KB smem, 256 TMEM columns):
//                 S = Q K_j^T, dP = dO V_j^T -> P, dS (thread = query row) -> dQ += dS K_j  accumulated in TMEM, stored once.
//  dkv_kernel  one CTA per 128-key tile, loops over 64-query tiles; the transposed scores live in TMEM double-buffered
//                 S^T = K Q_i^T, dP^T = V dO_i^T (thread = key row) -> P^T, dS^T -> dV += P^T dO_i, dK += dS^T Q_i,
//              so the MMAs of tile i+1 run under the exponentials of tile i.
// S and dP are recomputed in both kernels (7 instead of 5 contractions per pair) — cheaper than the serialisation they remove.
#include <stdlib.h>
#include <string.h>

#include "../../include/qfx.h"
#include "host_common.h"
#include "sm100.cuh"

extern long long* g_qfx_attn_bwd_dbg;  // attention_bwd.cu (qfx_attn_bwd_set_debug)

namespace qfx {

struct AttnBwd2Params {
  CUtensorMap tmQ128, tmdO128;  // box {64, 128, 1}: 128-row tiles (dq_kernel Q / dO)
  CUtensorMap tmK64, tmV64;     // box {64, 64, 1}:  64-row tiles (dq_kernel K / V)
  CUtensorMap tmK128, tmV128;   // box {64, 128, 1}: dkv_kernel K / V
  CUtensorMap tmQ64, tmdO64;    // box {64, 64, 1}:  dkv_kernel Q / dO
  float* dQ;                    // [B*H, S, 128] fp32 (plain stores; no zero-initialisation required)
  bf16* dK;
  bf16* dV;
  const float* lse;    // log2 domain
  const float* delta;
  const int* kv_len;
  const int* txt_len;
  int split, S, H;
  float scale, scale_log2;
  long long* dbg;  // optional clock64 stamps of dkv CTA (1, 0): [query tile][16]  (qfx_attn_bwd_set_debug)
};

__device__ __forceinline__ void mask32(uint32_t* r, int col0, int valid, int gap0, int gap1) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = col0 + i;
    if (!(k < valid) || (k >= gap0 && k < gap1)) r[i] = 0xff800000u;  // -inf -> P = 0
  }
}

// ====================================================================================================================== dQ
constexpr int DQ_T128 = 128 * 128 * 2;  // 32 KB
constexpr int DQ_T64 = 64 * 128 * 2;    // 16 KB
constexpr int DQ_SMEM = 2 * DQ_T128 + 2 * DQ_T64 + 128 * 128 + 256;  // Q, dO, K, V, dS[128 x 64] + barriers = 112.25 KB

__global__ void __launch_bounds__(192, 2) attn_bwd_dq_kernel(const __grid_constant__ AttnBwd2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t sQ = smem_base, sdO = smem_base + DQ_T128, sK = smem_base + 2 * DQ_T128, sV = sK + DQ_T64, sdS = sV + DQ_T64;
  const uint32_t bar_base = sdS + 128 * 128;
  const uint32_t qdo_full = bar_base, kv_full = bar_base + 8, kv_empty = bar_base + 16, sdp_full = bar_base + 24;
  const uint32_t ds_full = bar_base + 32, acc_full = bar_base + 40, tmem_slot = bar_base + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int b = bh / P.H;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  const int n_tiles = (kv_len + 63) / 64;

  if (warp == 1 && lane == 0) {
    mbar_init(qdo_full, 1);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(ds_full, 4);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const uint32_t tS = tmem_base, tdP = tmem_base + 64, tdQ = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qdo_full, 2 * DQ_T128);
      tma_load_3d(sQ, &P.tmQ128, qdo_full, 0, q0, bh);
      tma_load_3d(sQ + 16384, &P.tmQ128, qdo_full, 64, q0, bh);
      tma_load_3d(sdO, &P.tmdO128, qdo_full, 0, q0, bh);
      tma_load_3d(sdO + 16384, &P.tmdO128, qdo_full, 64, q0, bh);
      for (int j = 0; j < n_tiles; ++j) {
        if (j > 0) mbar_wait(kv_empty, (j - 1) & 1);  // dQ MMA of tile j-1 has consumed K_{j-1} (and S/dP used K/V earlier)
        mbar_expect_tx(kv_full, 2 * DQ_T64);
        tma_load_3d(sK, &P.tmK64, kv_full, 0, j * 64, bh);
        tma_load_3d(sK + 8192, &P.tmK64, kv_full, 64, j * 64, bh);
        tma_load_3d(sV, &P.tmV64, kv_full, 0, j * 64, bh);
        tma_load_3d(sV + 8192, &P.tmV64, kv_full, 64, j * 64, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_s = idesc_bf16(128, 64, 0, 0);    // S / dP: A K-major (Q, dO), B K-major (K_j, V_j rows = keys)
      constexpr uint32_t id_q = idesc_bf16(128, 128, 0, 1);   // dQ: A = dS K-major, B = K_j MN-major (d contiguous)
      mbar_wait(qdo_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(kv_full, j & 1);
        tc_fence_after();  // (S / dP columns are free: ds_full(j-1) was waited before the dQ MMA of tile j-1)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tS, sdesc_sw128(sQ + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    sdesc_sw128(sK + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), id_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tdP, sdesc_sw128(sdO + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    sdesc_sw128(sV + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), id_s, k != 0);
        umma_commit(sdp_full);
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdQ, sdesc_sw128(sdS + k * 32, 16, 1024), sdesc_sw128(sK + k * 2048, 8192, 1024), id_q, (j | k) != 0);
        umma_commit(kv_empty);
      }
      umma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int q = q0 + row;
    const bool q_ok = q < P.S;
    const float L = q_ok ? P.lse[(int64_t)bh * P.S + q] : INFINITY;
    const float dls = (q_ok ? P.delta[(int64_t)bh * P.S + q] : 0.f) * P.scale;
    for (int j = 0; j < n_tiles; ++j) {
      const int valid = kv_len - j * 64;
      const int gap0 = txt_len - j * 64, gap1 = P.split - j * 64;
      const bool full_tile = valid >= 64 && (gap0 >= gap1 || gap0 >= 64 || gap1 <= 0);
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t dk[32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t rs[32], rp[32];
        tmem_ld32(tS + lane_off + cc * 32, rs);
        tmem_ld32(tdP + lane_off + cc * 32, rp);
        tmem_ld_wait();
        if (!full_tile) mask32(rs, cc * 32, valid, gap0, gap1);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = exp2f(__uint_as_float(rs[2 * e]) * P.scale_log2 - L);
          const float p1 = exp2f(__uint_as_float(rs[2 * e + 1]) * P.scale_log2 - L);
          dk[cc * 16 + e] = pack_bf16(p0 * (__uint_as_float(rp[2 * e]) * P.scale - dls), p1 * (__uint_as_float(rp[2 * e + 1]) * P.scale - dls));
        }
      }
      // the dS buffer is free: kv_empty(j-1) (== dQ MMA of tile j-1 done) precedes sdp_full(j) on the in-order tensor pipe
      const uint32_t ds_row = sdS + row * 128;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const uint32_t chunk = (uint32_t)v ^ (uint32_t)(row & 7);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + chunk * 16), "r"(dk[4 * v]), "r"(dk[4 * v + 1]),
                     "r"(dk[4 * v + 2]), "r"(dk[4 * v + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* dst = P.dQ + ((int64_t)bh * P.S + q) * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      tmem_ld32(tdQ + lane_off + c, r);
      tmem_ld_wait();
      if (n_tiles == 0) {  // no key tile ever accumulated: the TMEM columns were never written
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0;
      }
      if (q_ok) {
        float4* d4 = reinterpret_cast<float4*>(dst + c);
#pragma unroll
        for (int v = 0; v < 8; ++v)
          d4[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                              __uint_as_float(r[4 * v + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ====================================================================================================================== dK, dV
constexpr int KV_STAGES = 3;
constexpr int KV_LD = 4;  // L / delta ring depth: the stager runs up to 3 query tiles ahead (its global loads take ~1 tile period)
constexpr int KV_SMEM = 2 * DQ_T128 + KV_STAGES * 2 * DQ_T64 + 2 * 128 * 128 + KV_LD * 512 + 256;  // K, V, (Q_i, dO_i) x3, P^T, dS^T, L/delta ring
constexpr int KV_THREADS = 384;  // warp 0 TMA, 1 MMA, 2 L/delta stager, 3 idle, 4-11 compute (quad = warp & 3, column half = (warp - 4) >> 2)

__global__ void __launch_bounds__(KV_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnBwd2Params P) {
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
  const uint32_t acc_full = bar_base + 8u * 21, tmem_slot = bar_base + 8u * 22;
  float* ld_gen = reinterpret_cast<float*>(smem_raw + (sLD - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int b = bh / P.H;
  const int kv_len = P.kv_len ? P.kv_len[b] : P.S;
  const int txt_len = P.txt_len ? P.txt_len[b] : P.split;
  const int n_q = (P.S + 63) / 64;
  const bool active = kv0 < kv_len && !(kv0 >= txt_len && kv0 + 128 <= P.split);
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0;
#define DBG2(i, e) do { if (dbg_on) P.dbg[(i) * 16 + (e)] = clock64(); } while (0)

  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int st = 0; st < KV_STAGES; ++st) {
      mbar_init(q_full(st), 1);
      mbar_init(q_empty(st), 1);
    }
    for (int u = 0; u < 2; ++u) mbar_init(s_full(u), 1);
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
      auto issue_s = [&](int i) {
        const int st = i % KV_STAGES, u = i & 1;
        mbar_wait(q_full(st), (i / KV_STAGES) & 1);
        tc_fence_after();
        DBG2(i, 0);
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
        if (i + 1 < n_q) issue_s(i + 1);
        mbar_wait(p_full, i & 1);
        tc_fence_after();
        DBG2(i, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdV, sdesc_sw128(sPt + k * 32, 16, 1024), sdesc_sw128(sdO(st) + k * 2048, 8192, 1024), id_a, (i | k) != 0);
        umma_commit(dv_done);
        mbar_wait(ds_full, i & 1);
        tc_fence_after();
        DBG2(i, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdK, sdesc_sw128(sdSt + k * 32, 16, 1024), sdesc_sw128(sQ(st) + k * 2048, 8192, 1024), id_a, (i | k) != 0);
        umma_commit(dk_done);
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
        if (threadIdx.x == 128) DBG2(i, 4);
        mbar_wait(ld_full(ul), (i / KV_LD) & 1);
        if (threadIdx.x == 128) DBG2(i, 5);
        const float4* L4 = reinterpret_cast<const float4*>(ld_gen + ul * 128 + half * 32);
        const float4* D4 = reinterpret_cast<const float4*>(ld_gen + ul * 128 + 64 + half * 32);
        mbar_wait(s_full(u), (i >> 1) & 1);
        tc_fence_after();
        if (threadIdx.x == 128) DBG2(i, 6);
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
        if (threadIdx.x == 128) DBG2(i, 7);
        if (i > 0) mbar_wait(dv_done, (i - 1) & 1);  // P^T buffer consumed by the dV MMA of tile i-1
        if (threadIdx.x == 128) DBG2(i, 8);
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
        if (threadIdx.x == 128) DBG2(i, 9);
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
        if (threadIdx.x == 128) DBG2(i, 10);
        if (i > 0) mbar_wait(dk_done, (i - 1) & 1);  // dS^T buffer consumed by the dK MMA of tile i-1
        if (threadIdx.x == 128) DBG2(i, 11);
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
        if (threadIdx.x == 128) DBG2(i, 12);
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

int attn_bwd_two_kernel(const void* Q, const void* K, const void* V, const void* dO, const float* lse, const float* delta,
                        float* dQ, void* dK, void* dV, const int* kv_len, const int* txt_len, int split, int B, int H, int S,
                        float softmax_scale, cudaStream_t stream) {
  AttnBwd2Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = tmap3(&P.tmQ128, Q, B * H, S, 128)) || (rc = tmap3(&P.tmdO128, dO, B * H, S, 128)) || (rc = tmap3(&P.tmK64, K, B * H, S, 64)) ||
      (rc = tmap3(&P.tmV64, V, B * H, S, 64)) || (rc = tmap3(&P.tmK128, K, B * H, S, 128)) || (rc = tmap3(&P.tmV128, V, B * H, S, 128)) ||
      (rc = tmap3(&P.tmQ64, Q, B * H, S, 64)) || (rc = tmap3(&P.tmdO64, dO, B * H, S, 64)))
    return rc;
  P.dQ = dQ; P.dK = (bf16*)dK; P.dV = (bf16*)dV; P.lse = lse; P.delta = delta; P.kv_len = kv_len; P.txt_len = txt_len;
  P.split = txt_len ? split : 0; P.S = S; P.H = H; P.scale = softmax_scale; P.scale_log2 = softmax_scale * 1.4426950408889634f;
  P.dbg = ::g_qfx_attn_bwd_dbg;
  static bool attr_done = false;
  if (!attr_done) {
    QFX_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    QFX_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM));
    attr_done = true;
  }
  dim3 grid((S + 127) / 128, B * H);
  static const char* only = getenv("QFX_ATTN_BWD2_ONLY");  // debugging aid: "dq" or "dkv" launches just that kernel
  if (!only || only[1] == 'k') attn_bwd_dkv_kernel<<<grid, KV_THREADS, KV_SMEM, stream>>>(P);
  QFX_CUDA(cudaGetLastError());
  if (!only || only[1] == 'q') attn_bwd_dq_kernel<<<grid, 192, DQ_SMEM, stream>>>(P);
  QFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qfx
