// Fused softmax self-attention on tcgen05 / TMEM for the 224^2 ViT-S/16 shape (S = 197 tokens, head_dim 64; any
// S <= 256).  Replaces timm Attention: q @ k^T * scale -> softmax -> @ v (same math MOD:52-64).
//
// One CTA = one (image, head).  Q (<= 2 tiles of 128 rows), K and V of the head are brought in by TMA straight out of
// the qkv GEMM's output layout [NB][S][3][H][64] (3-D tensor map: column, token, image - tokens past S zero-fill).
// Per 128-query tile:
//   S  = Q K^T          one accumulator of NK = ceil16(S) fp32 columns in TMEM (4 UMMAs 128 x NK x 16)
//   P  = softmax(S)     4 warps, thread = query row: two passes over the row in TMEM (max; exp2 + sum); P is written
//                       back over S as packed 16-bit (tcgen05.st) - scores never leave the SM
//   O  = P V            A operand = P from TMEM, B = V tile read MN-major (NK/16 UMMAs 128 x 64 x 16)
//   epilogue            O / rowsum -> 16-bit -> swizzled smem (the spent Q tile) -> TMA store (rows past S clipped)
// 256 TMEM columns and ~100 KB of smem per CTA: two CTAs per SM overlap each other's load / MMA / softmax phases.
#include "common.cuh"
#include "mfvit_internal.h"

#include <stdlib.h>

namespace mfv {

constexpr int FA_THREADS = 160;  // warps 0..3: softmax + epilogue (TMEM lane quarter = warp), warp 4: TMA + MMA issue
constexpr uint32_t FA_TMEM_COLS = 256;
constexpr uint32_t FA_O_COL = 128;  // O accumulator columns [128, 192): past P (<= 128 columns), inside the spent S
constexpr float FA_LOG2E = 1.4426950408889634f;
constexpr float FA_LN2 = 0.6931471805599453f;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fb_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fa_tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(FA_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                   float* __restrict__ lse, int S, int H, int NK, float scale_log2, int has_o2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // 2 tiles x [128][64] 16-bit, 128B-swizzled; tile mt is reused as O staging
  uint8_t* sK = sQ + 32768;            // [NK][64]
  uint8_t* sV = sK + NK * 128;         // [NK][64]
  uint8_t* sO2 = sV + NK * 128;        // 16 KB staging of the optional bf16 copy of O
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO2 + 16384);
  uint64_t* bar_qk = bars;             // Q + K landed
  uint64_t* bar_v = bars + 1;          // V landed
  uint64_t* bar_s = bars + 2;          // S accumulator complete
  uint64_t* bar_p = bars + 3;          // P written to TMEM (4 warp arrivals)
  uint64_t* bar_o = bars + 4;          // O accumulator complete
  uint64_t* bar_free = bars + 5;       // O read out: TMEM may be overwritten by the next tile (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh % H;
  const int n_mt = (S + 127) >> 7;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmO);
      mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_o, 1);
      mbar_init(bar_free, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  if (warp == 4) {
    {  // whole warp in uniform control flow, one elected lane issues (see elect_one() in common.cuh)
      griddep_wait();
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_qk, (uint32_t)(n_mt * 16384 + NK * 128));
        for (int mt = 0; mt < n_mt; ++mt) tma_load_3d(sQ + mt * 16384, &tmQ, bar_qk, h * 64, mt * 128, b);
        tma_load_3d(sK, &tmKV, bar_qk, (H + h) * 64, 0, b);
        mbar_arrive_expect_tx(bar_v, (uint32_t)(NK * 128));
        tma_load_3d(sV, &tmKV, bar_v, (2 * H + h) * 64, 0, b);
      }
      __syncwarp();
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc_s = make_idesc2(fmt, fmt, 128, (uint32_t)NK, 0, 0);
      const uint32_t idesc_o = make_idesc2(fmt, fmt, 128, 64, 0, 1);
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      for (int mt = 0; mt < n_mt; ++mt) {
        if (mt > 0) {
          mbar_wait(bar_free, (uint32_t)((mt - 1) & 1));
          tc_fence_after();
        }
        const uint32_t qa = smem_u32(sQ + mt * 16384), ka = smem_u32(sK);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, make_smem_desc_sw128(qa + k * 32, 0u, 1024u), make_smem_desc_sw128(ka + k * 32, 0u, 1024u),
                      idesc_s, k > 0 ? 1u : 0u);
          umma_commit(bar_s);
        }
        __syncwarp();
        mbar_wait(bar_p, (uint32_t)(mt & 1));
        tc_fence_after();
        if (mt == 0) {
          mbar_wait(bar_v, 0);
          tc_fence_after();
        }
        const uint32_t va = smem_u32(sV);
        if (elect_one()) {
          for (int k = 0; k < NK / 16; ++k)
            umma_f16_ts(tmem_base + FA_O_COL, tmem_base + (uint32_t)(k * 8),
                        make_smem_desc_sw128(va + k * 2048, 8192u, 1024u), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(bar_o);
        }
        __syncwarp();
      }
    }
  } else {
    griddep_wait();
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int mt = 0; mt < n_mt; ++mt) {
      const uint32_t ph = (uint32_t)(mt & 1);
      mbar_wait(bar_s, ph);
      tc_fence_after();
      // Both passes stream the row out of TMEM in 32-column chunks with the next chunk's tcgen05.ld already in flight
      // (two register buffers), so the TMEM read latency is paid once per pass instead of once per chunk.
      const int nch = (NK + 31) >> 5;
      auto ld_chunk = [&](int c, uint32_t (&v)[32]) {
        if (c * 32 + 32 <= NK) tmem_ld32(trow + (uint32_t)(c * 32), v); else tmem_ld16(trow + (uint32_t)(c * 32), v);
      };
      // ---- pass 1: row maximum (columns >= S are padding)
      float mx = -INFINITY;
      {
        uint32_t va[32], vb[32];
        ld_chunk(0, va);
        auto red = [&](int c, const uint32_t (&v)[32]) {
          const int c0 = c * 32;
          if (c0 + 32 <= S) {
#pragma unroll
            for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(v[k]));
          } else {
            const bool full = c0 + 32 <= NK;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (c0 + k < S && (full || k < 16)) mx = fmaxf(mx, __uint_as_float(v[k]));
          }
        };
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nch) ld_chunk(c + 1, vb);
          red(c, va);
          if (c + 1 < nch) {
            tmem_ld_wait();
            if (c + 2 < nch) ld_chunk(c + 2, va);
            red(c + 1, vb);
          }
        }
      }
      const float mc = mx * scale_log2;
      // ---- pass 2: p = exp2(s * c - max * c), row sum, packed 16-bit P written back over S
      float sum = 0.f;
      {
        uint32_t va[32], vb[32];
        ld_chunk(0, va);
        auto proc = [&](int c, const uint32_t (&v)[32]) {
          const int c0 = c * 32;
          const bool full = c0 + 32 <= NK;
          float pr[32];
          if (c0 + 32 <= S) {
#pragma unroll
            for (int k = 0; k < 32; ++k) pr[k] = fast_exp2(fmaf(__uint_as_float(v[k]), scale_log2, -mc));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              pr[k] = (c0 + k < S && (full || k < 16)) ? fast_exp2(fmaf(__uint_as_float(v[k]), scale_log2, -mc)) : 0.f;
          }
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            sum += pr[2 * k] + pr[2 * k + 1];
            pk[k] = F16 ? pack_f16(pr[2 * k], pr[2 * k + 1]) : pack_bf16(pr[2 * k], pr[2 * k + 1]);
          }
          // P chunk c lands in columns [16c, 16c+16): always below the S columns still to be read (>= 32(c+1)), and the
          // chunk c+1 load in flight reads [32(c+1), 32(c+2)) - no overlap either
          if (full) tmem_st16(trow + (uint32_t)(c0 >> 1), pk); else tmem_st8(trow + (uint32_t)(c0 >> 1), pk);
        };
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nch) ld_chunk(c + 1, vb);
          proc(c, va);
          if (c + 1 < nch) {
            tmem_ld_wait();
            if (c + 2 < nch) ld_chunk(c + 2, va);
            proc(c + 1, vb);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      const int row = mt * 128 + warp * 32 + lane;
      if (row < S) lse[(long long)bh * S + row] = (mc + log2f(sum)) * FA_LN2;
      const float inv = 1.f / sum;
      // ---- epilogue: O / sum -> 16-bit -> swizzled smem -> TMA store
      mbar_wait(bar_o, ph);
      tc_fence_after();
      float o[64];
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + FA_O_COL, v0);
        tmem_ld32(trow + FA_O_COL + 32u, v1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) { o[k] = __uint_as_float(v0[k]) * inv; o[32 + k] = __uint_as_float(v1[k]) * inv; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
      if (mt * 128 + warp * 32 < S) {  // warp-uniform: slices fully past S store nothing
        uint8_t* st = sQ + mt * 16384 + warp * 4096;  // this tile's Q rows are spent (S MMA complete)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 w;
          if (F16)
            w = make_uint4(pack_f16(o[8 * j], o[8 * j + 1]), pack_f16(o[8 * j + 2], o[8 * j + 3]),
                           pack_f16(o[8 * j + 4], o[8 * j + 5]), pack_f16(o[8 * j + 6], o[8 * j + 7]));
          else
            w = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                           pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
          *reinterpret_cast<uint4*>(st + lane * 128 + ((j ^ (lane & 7)) << 4)) = w;
        }
        uint8_t* st2 = sO2 + warp * 4096;
        if (has_o2) {
          if (mt > 0) {  // the previous tile's copy store must have finished reading the staging buffer
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(st2 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                           pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          fa_tma_store_3d(&tmO, st, h * 64, mt * 128 + warp * 32, b);
          if (has_o2) fa_tma_store_3d(&tmO2, st2, h * 64, mt * 128 + warp * 32, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FA_TMEM_COLS);
  }
}


// =====================================================================================================================
// Backward.  One CTA = one (image, head); bf16 operands (fp16 q/k/v of the fp16-forward mode are converted in place in
// shared memory), fp32 accumulation in TMEM.  With q-tiles i and key-blocks j of 128 (S <= 224 -> 2 x 2 steps):
//   S_ij  = Q_i K_j^T            dP_ij = dO_i V_j^T                       (TMEM columns   0..127 / 128..255)
//   P = exp2(S c - lse)          dS = P (dP - delta) scale                16 warps: thread = (row, 32-column segment);
//                                                                         no row reductions are needed in the backward
//   P, dS -> bf16 -> one swizzled smem tile each, read both K-major (dQ) and MN-major (dK, dV: the transposed use)
//   dV_j += P^T dO_i             dK_j += dS^T Q_i       dQ_i += dS K_j    (TMEM 448..511 / 384..447 / 256..383)
// The MMA warp runs one step ahead: S / dP of step n+1 are issued as soon as the compute warps have read step n out of
// TMEM, so the tensor pipe works while P / dS of step n are being formed.  delta = rowsum(dO * O) is computed in the
// prologue.  Gradients leave through 64B-swizzled staging + TMA stores straight into dqkv [NB][S][3][H][64].
constexpr int FB_CWARPS = 16;
constexpr int FB_THREADS = 32 * (FB_CWARPS + 1);
constexpr uint32_t FB_S_COL = 0, FB_DP_COL = 128, FB_DQ_COL = 256, FB_DK_COL = 384, FB_DV_COL = 448;

template <bool QKV_F16>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmG, const __nv_bfloat16* __restrict__ o,
                   const float* __restrict__ lse, int S, int H, int NK, float scale, unsigned long long* prof) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int TILE = NK * 128;           // [NK][64] bf16, 128B-swizzled (multiple of 2 KB)
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TILE;
  uint8_t* sK = sdO + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;             // [2 key halves of 64][128 q][128 B]
  uint8_t* sdS = sP + 32768;
  uint8_t* sStage = sdS + 32768;       // 16 warps x 2 KB
  float* sDelta = reinterpret_cast<float*>(sStage + FB_CWARPS * 2048);
  float* sLse = sDelta + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLse + 256);
  uint64_t* bar_ld = bars;             // [2]: Q,K | V,dO landed
  uint64_t* bar_conv = bars + 2;       // fp16 mode: Q, K converted to bf16 (16 warp arrivals)
  uint64_t* bar_sdp = bars + 3;        // S, dP accumulators complete
  uint64_t* bar_sdp_free = bars + 4;   // S, dP read out of TMEM (16)
  uint64_t* bar_pds = bars + 5;        // P, dS tiles written (16)
  uint64_t* bar_mma = bars + 6;        // dV, dK, dQ MMAs of the step complete (P / dS tiles free, accumulators valid)
  uint64_t* bar_acc_free = bars + 7;   // dK_j, dV_j read out (16)
  uint64_t* bar_conv2 = bars + 8;      // fp16 mode: V converted to bf16 (16 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh % H;
  const int n_t = (S + 127) >> 7;      // q-tiles == key-blocks
  const int n_steps = n_t * n_t;
  auto rows_of = [&](int t) { return min(128, S - 128 * t); };
  auto r16 = [](int x) { return (x + 15) & ~15; };

  if (warp == FB_CWARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      tma_prefetch_desc(&tmG);
      mbar_init(&bar_ld[0], 1); mbar_init(&bar_ld[1], 1);
      mbar_init(bar_conv, FB_CWARPS); mbar_init(bar_conv2, FB_CWARPS); mbar_init(bar_sdp, 1);
      mbar_init(bar_sdp_free, FB_CWARPS);
      mbar_init(bar_pds, FB_CWARPS); mbar_init(bar_mma, 1); mbar_init(bar_acc_free, FB_CWARPS);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  if (warp == FB_CWARPS) {
    // ------------------------------------------------------------------------------------------ TMA + MMA issue
    // (whole warp in uniform control flow, one elected lane issues: see elect_one() in common.cuh)
    {
      griddep_wait();
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_ld[0], (uint32_t)(2 * TILE));
        tma_load_3d(sQ, &tmQKV, &bar_ld[0], h * 64, 0, b);
        tma_load_3d(sK, &tmQKV, &bar_ld[0], (H + h) * 64, 0, b);
        mbar_arrive_expect_tx(&bar_ld[1], (uint32_t)(2 * TILE));
        tma_load_3d(sV, &tmQKV, &bar_ld[1], (2 * H + h) * 64, 0, b);
        tma_load_3d(sdO, &tmDO, &bar_ld[1], h * 64, 0, b);
      }
      __syncwarp();
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
      auto issue_s = [&](int n) {
        const int j = n / n_t, i = n % n_t;
        const uint32_t idesc = make_idesc2(1u, 1u, 128, (uint32_t)r16(rows_of(j)), 0, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + FB_S_COL, make_smem_desc_sw128(aQ + i * 16384 + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(aK + j * 16384 + k * 32, 0u, 1024u), idesc, k > 0 ? 1u : 0u);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int n) {
        const int j = n / n_t, i = n % n_t;
        const uint32_t idesc = make_idesc2(1u, 1u, 128, (uint32_t)r16(rows_of(j)), 0, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + FB_DP_COL, make_smem_desc_sw128(adO + i * 16384 + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(aV + j * 16384 + k * 32, 0u, 1024u), idesc, k > 0 ? 1u : 0u);
          umma_commit(bar_sdp);
        }
        __syncwarp();
      };
      // S and dP of a step alternate k-step by k-step: consecutive tcgen05.mma on the SAME accumulator columns leave a
      // bubble of ~45 clk between them (tests/gpu_ring_probe3.py), two accumulators taken in turn hide it
      auto issue_sdp = [&](int n) {
        const int j = n / n_t, i = n % n_t;
        const uint32_t idesc = make_idesc2(1u, 1u, 128, (uint32_t)r16(rows_of(j)), 0, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem_base + FB_S_COL, make_smem_desc_sw128(aQ + i * 16384 + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(aK + j * 16384 + k * 32, 0u, 1024u), idesc, k > 0 ? 1u : 0u);
            umma_bf16(tmem_base + FB_DP_COL, make_smem_desc_sw128(adO + i * 16384 + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(aV + j * 16384 + k * 32, 0u, 1024u), idesc, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_sdp);
        }
        __syncwarp();
      };
      // the first S = Q K^T starts as soon as Q and K are in shared memory (and converted, in fp16 mode) - it does not
      // wait for V, dO or the delta prologue of the compute warps
      mbar_wait(&bar_ld[0], 0);
      if (QKV_F16) mbar_wait(bar_conv, 0);
      tc_fence_after();
      issue_s(0);
      mbar_wait(&bar_ld[1], 0);
      if (QKV_F16) mbar_wait(bar_conv2, 0);
      tc_fence_after();
      issue_dp(0);
      const uint32_t idesc_t = make_idesc2(1u, 1u, 128, 64, 1, 1);  // A = P / dS read MN-major (transposed use)
      const uint32_t idesc_q = make_idesc2(1u, 1u, 128, 64, 0, 1);  // A = dS K-major
      for (int n = 0; n < n_steps; ++n) {
        const int j = n / n_t, i = n % n_t;
        if (n + 1 < n_steps) {
          mbar_wait(bar_sdp_free, (uint32_t)(n & 1));
          tc_fence_after();
          issue_sdp(n + 1);
        }
        mbar_wait(bar_pds, (uint32_t)(n & 1));
        tc_fence_after();
        if (i == 0 && j > 0) {
          mbar_wait(bar_acc_free, (uint32_t)((j - 1) & 1));
          tc_fence_after();
        }
        const int kq = r16(rows_of(i)) / 16;  // reduction over the q rows of tile i
        const int kk = r16(rows_of(j)) / 16;  // reduction over the keys of block j
        if (elect_one()) {
          // dV, dK and dQ k-steps in turn (three accumulators: no back-to-back UMMAs on the same columns)
          const int kmax = kq > kk ? kq : kk;
          for (int k = 0; k < kmax; ++k) {
            if (k < kq) {
              const uint64_t da_p = make_smem_desc_sw128(aP + k * 2048, 16384u, 1024u);
              const uint64_t da_s = make_smem_desc_sw128(adS + k * 2048, 16384u, 1024u);
              const uint64_t db_do = make_smem_desc_sw128(adO + i * 16384 + k * 2048, 8192u, 1024u);
              const uint64_t db_q = make_smem_desc_sw128(aQ + i * 16384 + k * 2048, 8192u, 1024u);
              const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
              umma_bf16(tmem_base + FB_DV_COL, da_p, db_do, idesc_t, acc);
              umma_bf16(tmem_base + FB_DK_COL, da_s, db_q, idesc_t, acc);
            }
            if (k < kk) {
              const uint64_t da = make_smem_desc_sw128(adS + (k >> 2) * 16384 + (k & 3) * 32, 0u, 1024u);
              const uint64_t db = make_smem_desc_sw128(aK + j * 16384 + k * 2048, 8192u, 1024u);
              umma_bf16(tmem_base + FB_DQ_COL + (uint32_t)(i * 64), da, db, idesc_q, (j > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_mma);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ compute warps
    griddep_wait();
    // phase clocks (measurement aid, MFVIT_ATTN_PROF): 0 prologue (loads, conversion, delta), 1 wait for S / dP,
    // 2 TMEM read + P / dS math, 3 wait for the previous step's MMAs, 4 P / dS tiles to smem, 5 dK / dV read-out,
    // 6 dQ read-out and drain
    long long prof_t = prof ? clock64() : 0;
    unsigned long long prof_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define FB_MARK(slot)                                    \
  if (prof) {                                            \
    const long long now_ = clock64();                    \
    prof_c[slot] += (unsigned long long)(now_ - prof_t); \
    prof_t = now_;                                       \
  }
    const int tid = threadIdx.x;  // 0..511
    const int q = warp & 3;       // TMEM lane quarter
    const int cseg = warp >> 2;   // 32-column segment
    // the O row of the delta prologue comes straight from global: issue those loads before waiting for the TMA tiles
    const int drow = tid >> 1, dhalf = tid & 1;
    uint4 ovec[4] = {};
    float lse_row = INFINITY;
    if (drow < S) {
      const uint4* og = reinterpret_cast<const uint4*>(o + (((long long)b * S + drow) * H + h) * 64 + dhalf * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) ovec[c] = __ldg(og + c);
      lse_row = lse[(long long)bh * S + drow] * FA_LOG2E;
    }
    auto to_bf16 = [&](uint8_t* base) {  // fp16 -> bf16 in place (element-wise, so the swizzle is irrelevant)
      for (int c = tid; c < TILE / 16; c += FB_CWARPS * 32) {
        uint4* pp = reinterpret_cast<uint4*>(base + c * 16);
        uint4 w = *pp;
        uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f2 = unpack_f16(ww[e]);
          ww[e] = pack_bf16(f2.x, f2.y);
        }
        *pp = w;
      }
    };
    mbar_wait(&bar_ld[0], 0);
    if (QKV_F16) {
      to_bf16(sQ);
      to_bf16(sK);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv);  // S = Q K^T may start
    }
    mbar_wait(&bar_ld[1], 0);
    if (QKV_F16) {
      to_bf16(sV);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv2);  // dP = dO V^T may start
    }
    {  // delta = rowsum(dO * O), lse in the log2 domain (+inf for padded rows -> P = dS = 0 there)
      const int row = drow, half = dhalf;
      float acc = 0.f;
      if (row < S) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 dv = *reinterpret_cast<const uint4*>(sdO + row * 128 + (((half * 4 + c) ^ (row & 7)) << 4));
          const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ovec[c]);
          const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = unpack_bf16(ow[e]), d = unpack_bf16(dw[e]);
            acc = fmaf(a.x, d.x, acc);
            acc = fmaf(a.y, d.y, acc);
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0) {
        sDelta[row] = acc;
        sLse[row] = lse_row;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(FB_CWARPS * 32) : "memory");  // sDelta / sLse visible to every compute warp

    const float scale_log2 = scale * FA_LOG2E;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* stg = sStage + warp * 2048;
    auto stage_store = [&](const float (&f)[32], int which, int colseg, int row0) {
      // 32 rows x 32 bf16 columns -> 64B-swizzled staging -> TMA store into dqkv[.., which, h, colseg*32 ..]
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(pack_bf16(f[8 * c], f[8 * c + 1]), pack_bf16(f[8 * c + 2], f[8 * c + 3]),
                       pack_bf16(f[8 * c + 4], f[8 * c + 5]), pack_bf16(f[8 * c + 6], f[8 * c + 7]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        fa_tma_store_3d(&tmG, stg, (which * H + h) * 64 + colseg * 32, row0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    FB_MARK(0)
    for (int n = 0; n < n_steps; ++n) {
      const int j = n / n_t, i = n % n_t;
      const int qrows = rows_of(i), nk = r16(rows_of(j));
      const bool active = (q * 32 < qrows) && (cseg * 32 < nk);
      const int r = q * 32 + lane;
      mbar_wait(bar_sdp, (uint32_t)(n & 1));
      tc_fence_after();
      FB_MARK(1)
      uint32_t pk[16], dk[16];
      if (active) {
        const float l2 = sLse[i * 128 + r], dl = sDelta[i * 128 + r];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // two 16-column halves keep the live register set small
          uint32_t sv[16], dv[16];
          fb_tmem_ld16(tlane + FB_S_COL + (uint32_t)(cseg * 32 + hh * 16), sv);
          fb_tmem_ld16(tlane + FB_DP_COL + (uint32_t)(cseg * 32 + hh * 16), dv);
          tmem_ld_wait();
          const int key0 = j * 128 + cseg * 32 + hh * 16;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * e]), scale_log2, -l2));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * e + 1]), scale_log2, -l2));
            if (key0 + 2 * e >= S) p0 = 0.f;
            if (key0 + 2 * e + 1 >= S) p1 = 0.f;
            const float d0 = p0 * (__uint_as_float(dv[2 * e]) - dl) * scale;
            const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - dl) * scale;
            pk[hh * 8 + e] = pack_bf16(p0, p1);
            dk[hh * 8 + e] = pack_bf16(d0, d1);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sdp_free);
      FB_MARK(2)
      if (n > 0) mbar_wait(bar_mma, (uint32_t)((n - 1) & 1));  // the previous step's MMAs have finished reading P / dS
      FB_MARK(3)
      if (active) {
        const uint32_t off = (uint32_t)((cseg >> 1) * 16384 + r * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t sw = off + ((((cseg & 1) * 4 + c) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(sP + sw) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(sdS + sw) = make_uint4(dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pds);
      FB_MARK(4)
      if (i == n_t - 1) {  // dK_j, dV_j are complete once this step's MMAs retire
        mbar_wait(bar_mma, (uint32_t)(n & 1));
        tc_fence_after();
        const bool rows_ok = q * 32 < rows_of(j);  // lanes = keys of block j
        float f[32];
        if (rows_ok) {
          uint32_t v[32];
          tmem_ld32(tlane + (cseg < 2 ? FB_DK_COL : FB_DV_COL) + (uint32_t)((cseg & 1) * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_free);
        if (rows_ok) stage_store(f, cseg < 2 ? 1 : 2, cseg & 1, j * 128 + q * 32);
        FB_MARK(5)
      }
    }
    {  // dQ: tile = cseg >> 1 (lanes = its queries), columns (cseg & 1) * 32
      const int i = cseg >> 1;
      if (i < n_t && q * 32 < rows_of(i)) {
        uint32_t v[32];
        float f[32];
        tc_fence_after();
        tmem_ld32(tlane + FB_DQ_COL + (uint32_t)(i * 64 + (cseg & 1) * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
        stage_store(f, 0, cseg & 1, i * 128 + q * 32);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
    FB_MARK(6)
    if (prof && lane == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(prof + warp * 8 + k, prof_c[k]);
      if (warp == 0) atomicAdd(prof + FB_CWARPS * 8, 1ULL);
    }
#undef FB_MARK
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FB_CWARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// Backward for long sequences (S > 224, e.g. 577 tokens at 384^2).  One CTA = one (image, head, key block j of 128 keys);
// it walks the query tiles i, with K_j / V_j resident and Q_i / dO_i streamed through a two-slot TMA ring:
//   S = Q_i K_j^T, dP = dO_i V_j^T -> P, dS (as above) -> dV_j += P^T dO_i, dK_j += dS^T Q_i (accumulated over i in TMEM)
//   dQ_i (this key block's share) = dS K_j -> fp32 TMA reduce-add into dq_ws [NB][S][H*64]; a small kernel turns the
//   summed dQ into bf16 afterwards.  delta = rowsum(dO * O) comes from attn_delta_kernel.
// TMEM: S | dP | dQ | dK | dV = 128 + 128 + 64 + 64 + 64 columns.  Warps 0..15 compute, 16 = MMA issue, 17 = TMA producer.
constexpr int FM_THREADS = 32 * (FB_CWARPS + 2);
constexpr uint32_t FM_S_COL = 0, FM_DP_COL = 128, FM_DQ_COL = 256, FM_DK_COL = 320, FM_DV_COL = 384;

__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                  long long rows /* NB*S */, int S, int H) {
  // one thread per (row, head): 64 bf16 of O and dO
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * H) return;
  const long long row = t / H;
  const int h = (int)(t % H);
  const uint4* a = reinterpret_cast<const uint4*>(o + (row * H + h) * 64);
  const uint4* b = reinterpret_cast<const uint4*>(d_o + (row * H + h) * 64);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 x = __ldg(a + c), y = __ldg(b + c);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x);
    const uint32_t* yw = reinterpret_cast<const uint32_t*>(&y);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 p = unpack_bf16(xw[e]), q = unpack_bf16(yw[e]);
      acc = fmaf(p.x, q.x, acc);
      acc = fmaf(p.y, q.y, acc);
    }
  }
  const long long nb = row / S;
  const int srow = (int)(row % S);
  delta[(nb * H + h) * S + srow] = acc;  // [NB][H][S], the layout of lse
}

// dqkv[nb][s][0][h][:] = bf16(dq_ws[nb][s][h][:])
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ dq_ws, __nv_bfloat16* __restrict__ dqkv, long long rows, int H) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 8 elements
  const long long per_row = (long long)H * 8;
  if (t >= rows * per_row) return;
  const long long row = t / per_row;
  const int c8 = (int)(t % per_row);
  const float4 a = reinterpret_cast<const float4*>(dq_ws + row * H * 64)[c8 * 2];
  const float4 b = reinterpret_cast<const float4*>(dq_ws + row * H * 64)[c8 * 2 + 1];
  reinterpret_cast<uint4*>(dqkv + row * 3 * H * 64)[c8] =
      make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

template <bool QKV_F16>
__global__ void __launch_bounds__(FM_THREADS, 1)
attn_bwd_tc_mb_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmDQ,
                      const float* __restrict__ lse, const float* __restrict__ delta, int S, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                  // [128][64] bf16, 128B-swizzled (keys j*128 ..)
  uint8_t* sV = sK + 16384;
  uint8_t* sQ = sV + 16384;            // 2 slots x [128][64]
  uint8_t* sdO = sQ + 2 * 16384;       // 2 slots
  uint8_t* sP = sdO + 2 * 16384;       // [2 key halves of 64][128 q][128 B]
  uint8_t* sdS = sP + 32768;
  uint8_t* sStage = sdS + 32768;       // 16 warps x 2 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + FB_CWARPS * 2048);
  uint64_t* bar_kv = bars;             // K_j, V_j landed
  uint64_t* bar_q = bars + 1;          // [2] Q_i, dO_i landed in slot
  uint64_t* bar_qfree = bars + 3;      // [2] slot consumed by the MMAs of its step
  uint64_t* bar_kvconv = bars + 5;     // fp16 mode: K, V converted (16)
  uint64_t* bar_qconv = bars + 6;      // [2] fp16 mode: Q_i converted (16)
  uint64_t* bar_sdp = bars + 8;        // S, dP complete
  uint64_t* bar_sdp_free = bars + 9;   // S, dP read out (16)
  uint64_t* bar_pds = bars + 10;       // P, dS written (16)
  uint64_t* bar_mma = bars + 11;       // dV, dK, dQ MMAs of the step complete
  uint64_t* bar_dq_free = bars + 12;   // dQ read out of TMEM (16)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int bh = blockIdx.x, j = blockIdx.y;
  const int b = bh / H, h = bh % H;
  const int n_t = (S + 127) >> 7;
  auto rows_of = [&](int t) { return min(128, S - 128 * t); };
  auto r16 = [](int x) { return (x + 15) & ~15; };
  const int nk = r16(rows_of(j));      // keys of this block, padded to the UMMA N / K granularity

  if (warp == FB_CWARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmDQ);
      mbar_init(bar_kv, 1);
      for (int s = 0; s < 2; ++s) { mbar_init(&bar_q[s], 1); mbar_init(&bar_qfree[s], 1); mbar_init(&bar_qconv[s], FB_CWARPS); }
      mbar_init(bar_kvconv, FB_CWARPS); mbar_init(bar_sdp, 1); mbar_init(bar_sdp_free, FB_CWARPS);
      mbar_init(bar_pds, FB_CWARPS); mbar_init(bar_mma, 1); mbar_init(bar_dq_free, FB_CWARPS);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  if (warp == FB_CWARPS + 1) {
    // ------------------------------------------------------------------------------------------ TMA producer
    // (whole warp in uniform control flow, one elected lane issues: see elect_one() in common.cuh)
    griddep_wait();
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_kv, 2 * 16384);
      tma_load_3d(sK, &tmQKV, bar_kv, (H + h) * 64, j * 128, b);
      tma_load_3d(sV, &tmQKV, bar_kv, (2 * H + h) * 64, j * 128, b);
    }
    __syncwarp();
    for (int i = 0; i < n_t; ++i) {
      const int s = i & 1;
      if (i >= 2) mbar_wait(&bar_qfree[s], (uint32_t)(((i >> 1) - 1) & 1));
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_q[s], 2 * 16384);
        tma_load_3d(sQ + s * 16384, &tmQKV, &bar_q[s], h * 64, i * 128, b);
        tma_load_3d(sdO + s * 16384, &tmDO, &bar_q[s], h * 64, i * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == FB_CWARPS) {
    // ------------------------------------------------------------------------------------------ MMA issue
    {
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO);
      const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
      const uint32_t idesc_s = make_idesc2(1u, 1u, 128, (uint32_t)nk, 0, 0);
      const uint32_t idesc_t = make_idesc2(1u, 1u, 128, 64, 1, 1);  // A = P / dS read MN-major
      const uint32_t idesc_q = make_idesc2(1u, 1u, 128, 64, 0, 1);  // A = dS K-major
      auto wait_q = [&](int i) {
        const int s = i & 1;
        mbar_wait(&bar_q[s], (uint32_t)((i >> 1) & 1));
        if (QKV_F16) mbar_wait(&bar_qconv[s], (uint32_t)((i >> 1) & 1));
        tc_fence_after();
      };
      auto issue_sdp = [&](int i) {
        const uint32_t q = aQ + (i & 1) * 16384, d = adO + (i & 1) * 16384;
        if (elect_one()) {  // S and dP k-steps in turn: no back-to-back UMMAs on the same accumulator columns
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem_base + FM_S_COL, make_smem_desc_sw128(q + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(aK + k * 32, 0u, 1024u), idesc_s, k > 0 ? 1u : 0u);
            umma_bf16(tmem_base + FM_DP_COL, make_smem_desc_sw128(d + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(aV + k * 32, 0u, 1024u), idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_sdp);
        }
        __syncwarp();
      };
      mbar_wait(bar_kv, 0);
      if (QKV_F16) mbar_wait(bar_kvconv, 0);
      wait_q(0);
      issue_sdp(0);
      for (int i = 0; i < n_t; ++i) {
        if (i + 1 < n_t) {
          mbar_wait(bar_sdp_free, (uint32_t)(i & 1));
          wait_q(i + 1);
          issue_sdp(i + 1);
        }
        mbar_wait(bar_pds, (uint32_t)(i & 1));
        if (i > 0) mbar_wait(bar_dq_free, (uint32_t)((i - 1) & 1));  // dQ of the previous tile has been read out
        tc_fence_after();
        const uint32_t q = aQ + (i & 1) * 16384, d = adO + (i & 1) * 16384;
        const int kq = r16(rows_of(i)) / 16;
        if (elect_one()) {
          const int kk = nk / 16, kmax = kq > kk ? kq : kk;
          for (int k = 0; k < kmax; ++k) {  // dV, dK and dQ k-steps in turn
            if (k < kq) {
              const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
              umma_bf16(tmem_base + FM_DV_COL, make_smem_desc_sw128(aP + k * 2048, 16384u, 1024u),
                        make_smem_desc_sw128(d + k * 2048, 8192u, 1024u), idesc_t, acc);
              umma_bf16(tmem_base + FM_DK_COL, make_smem_desc_sw128(adS + k * 2048, 16384u, 1024u),
                        make_smem_desc_sw128(q + k * 2048, 8192u, 1024u), idesc_t, acc);
            }
            if (k < kk)
              umma_bf16(tmem_base + FM_DQ_COL, make_smem_desc_sw128(adS + (k >> 2) * 16384 + (k & 3) * 32, 0u, 1024u),
                        make_smem_desc_sw128(aK + k * 2048, 8192u, 1024u), idesc_q, k > 0 ? 1u : 0u);
          }
          umma_commit(&bar_qfree[i & 1]);  // the slot's Q / dO tiles are no longer needed
          umma_commit(bar_mma);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ compute warps
    griddep_wait();
    const int tid = threadIdx.x;  // 0..511
    const int q = warp & 3, cseg = warp >> 2;
    auto to_bf16 = [&](uint8_t* base) {
      for (int c = tid; c < 16384 / 16; c += FB_CWARPS * 32) {
        uint4* pp = reinterpret_cast<uint4*>(base + c * 16);
        uint4 w = *pp;
        uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f2 = unpack_f16(ww[e]);
          ww[e] = pack_bf16(f2.x, f2.y);
        }
        *pp = w;
      }
    };
    if (QKV_F16) {
      mbar_wait(bar_kv, 0);
      to_bf16(sK);
      to_bf16(sV);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_kvconv);
    }
    const float scale_log2 = scale * FA_LOG2E;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* stg = sStage + warp * 2048;
    const bool col_active = cseg * 32 < nk;
    float dqv[16];            // this warp's 16 dQ columns of the previous tile, waiting for their reduce-add
    int dq_row0 = -1;
    auto flush_dq = [&]() {   // fp32 tile 32 rows x 16 columns -> staging -> TMA reduce-add into dq_ws
      if (dq_row0 < 0) return;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
            make_float4(dqv[4 * c], dqv[4 * c + 1], dqv[4 * c + 2], dqv[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmDQ)),
                     "r"(smem_u32(stg)), "r"(h * 64 + cseg * 16), "r"(dq_row0), "r"(b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      dq_row0 = -1;
    };
    auto read_dq = [&](int i) {  // after bar_mma(i): this warp's columns [cseg*16, +16) of dQ_i, rows q*32 ..
      tc_fence_after();
      const bool ok = q * 32 < rows_of(i);
      if (ok) {
        uint32_t v[16];
        fb_tmem_ld16(tlane + FM_DQ_COL + (uint32_t)(cseg * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) dqv[e] = __uint_as_float(v[e]);
        dq_row0 = i * 128 + q * 32;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dq_free);
    };
    for (int i = 0; i < n_t; ++i) {
      const int qrows = rows_of(i);
      const bool active = (q * 32 < qrows) && col_active;
      const int r = q * 32 + lane;
      auto convert_q = [&](int t) {  // fp16 mode: Q tile of step t to bf16 in place (dO is bf16 already)
        const int ss = t & 1;
        mbar_wait(&bar_q[ss], (uint32_t)((t >> 1) & 1));
        to_bf16(sQ + ss * 16384);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_qconv[ss]);
      };
      if (QKV_F16 && i == 0) convert_q(0);
      float l2 = INFINITY, dl = 0.f;
      if (i * 128 + r < S) {
        l2 = lse[(long long)bh * S + i * 128 + r] * FA_LOG2E;
        dl = delta[(long long)bh * S + i * 128 + r];
      }
      mbar_wait(bar_sdp, (uint32_t)(i & 1));
      tc_fence_after();
      uint32_t pk[16], dk[16];
      if (active) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t sv[16], dv[16];
          fb_tmem_ld16(tlane + FM_S_COL + (uint32_t)(cseg * 32 + hh * 16), sv);
          fb_tmem_ld16(tlane + FM_DP_COL + (uint32_t)(cseg * 32 + hh * 16), dv);
          tmem_ld_wait();
          const int key0 = j * 128 + cseg * 32 + hh * 16;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * e]), scale_log2, -l2));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * e + 1]), scale_log2, -l2));
            if (key0 + 2 * e >= S) p0 = 0.f;
            if (key0 + 2 * e + 1 >= S) p1 = 0.f;
            const float d0 = p0 * (__uint_as_float(dv[2 * e]) - dl) * scale;
            const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - dl) * scale;
            pk[hh * 8 + e] = pack_bf16(p0, p1);
            dk[hh * 8 + e] = pack_bf16(d0, d1);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sdp_free);
      // the next tile's Q is converted now, one step ahead, so the MMA warp can issue S / dP of step i+1 while this
      // step's P / dS are still being written (its slot was released by the MMAs of step i-1)
      if (QKV_F16 && i + 1 < n_t) convert_q(i + 1);
      if (i > 0) {  // the previous step's MMAs: P / dS tiles free again, dQ_{i-1} complete
        mbar_wait(bar_mma, (uint32_t)((i - 1) & 1));
        read_dq(i - 1);
      }
      if (active) {
        const uint32_t off = (uint32_t)((cseg >> 1) * 16384 + r * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t sw = off + ((((cseg & 1) * 4 + c) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(sP + sw) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(sdS + sw) = make_uint4(dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pds);
      flush_dq();  // dQ_{i-1}: off the critical path (the MMAs of step i are already unblocked)
    }
    // ---- tail: last tile's dQ, then dK_j / dV_j (lanes = keys of this block)
    mbar_wait(bar_mma, (uint32_t)((n_t - 1) & 1));
    read_dq(n_t - 1);
    flush_dq();
    {
      tc_fence_after();
      if (q * 32 < rows_of(j)) {
        uint32_t v[32];
        float f[32];
        tmem_ld32(tlane + (cseg < 2 ? FM_DK_COL : FM_DV_COL) + (uint32_t)((cseg & 1) * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(pack_bf16(f[8 * c], f[8 * c + 1]), pack_bf16(f[8 * c + 2], f[8 * c + 3]),
                         pack_bf16(f[8 * c + 4], f[8 * c + 5]), pack_bf16(f[8 * c + 6], f[8 * c + 7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          fa_tma_store_3d(&tmG, stg, ((cseg < 2 ? 1 : 2) * H + h) * 64 + (cseg & 1) * 32, j * 128 + q * 32, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FB_CWARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// Forward for long sequences (S > 256, e.g. 577 tokens at 384^2): flash-style.  One CTA = one (image, head, 128-query
// tile); K / V stream through a 4-slot TMA ring in blocks of 64 keys; S is double-buffered in TMEM (2 x 64 columns) so the
// Q K^T of block kb+1 runs under the softmax of block kb; the running row maximum m and sum l live in registers (thread =
// query row) and O in TMEM is rescaled only when a row's maximum moved (tcgen05.ld / st, warp-uniform skip otherwise).
// Warps 0..3 softmax + epilogue, 4 MMA issue, 5 TMA producer.  256 TMEM columns, ~97 KB smem: two CTAs per SM.
constexpr int FL_THREADS = 192;
constexpr int FL_RING = 4;
constexpr uint32_t FL_O_COL = 128;  // S buffers at columns 0 and 64 (P aliases the first 32 columns of each), O at 128..191

template <bool F16>
__global__ void __launch_bounds__(FL_THREADS, 2)
attn_fwd_tc_mb_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                      float* __restrict__ lse, int S, int H, float scale_log2, int has_o2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // [128][64], reused as O staging
  uint8_t* sRing = sQ + 16384;               // FL_RING x (K block [64][64] 8 KB | V block 8 KB)
  uint8_t* sO2 = sRing + FL_RING * 16384;    // 16 KB staging of the optional bf16 copy of O
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO2 + 16384);
  uint64_t* bar_q = bars;                    // Q landed
  uint64_t* ring_full = bars + 1;            // [FL_RING]
  uint64_t* ring_free = bars + 1 + FL_RING;  // [FL_RING]
  uint64_t* bar_s = bars + 1 + 2 * FL_RING;  // [2] S buffer complete
  uint64_t* bar_p = bar_s + 2;               // P written / O rescaled (4 warp arrivals)
  // P.V of block kb complete: bar_o[kb & 1].  TWO barriers, because a softmax warp does not wait for every block's P.V
  // (lazy rescaling below) and a parity wait cannot tell phase k from phase k + 2: with one barrier per block parity the
  // phase a warp asks for is always the barrier's current or last one.
  uint64_t* bar_o = bar_p + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 2);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int bh = blockIdx.x, mt = blockIdx.y;
  const int b = bh / H, h = bh % H;
  const int n_kb = (S + 63) >> 6;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmO);
      mbar_init(bar_q, 1);
      for (int i = 0; i < FL_RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_free[i], 1); }
      mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1); mbar_init(bar_p, 4); mbar_init(&bar_o[0], 1); mbar_init(&bar_o[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  if (warp == 5) {
    // ------------------------------------------------------------------------------------------ TMA producer
    // (whole warp in uniform control flow, one elected lane issues: see elect_one() in common.cuh)
    griddep_wait();
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, 16384);
      tma_load_3d(sQ, &tmQ, bar_q, h * 64, mt * 128, b);
    }
    __syncwarp();
    for (int kb = 0; kb < n_kb; ++kb) {
      const int sl = kb % FL_RING;
      if (kb >= FL_RING) mbar_wait(&ring_free[sl], (uint32_t)(((kb / FL_RING) - 1) & 1));
      if (elect_one()) {
        mbar_arrive_expect_tx(&ring_full[sl], 16384);
        tma_load_3d(sRing + sl * 16384, &tmKV, &ring_full[sl], (H + h) * 64, kb * 64, b);
        tma_load_3d(sRing + sl * 16384 + 8192, &tmKV, &ring_full[sl], (2 * H + h) * 64, kb * 64, b);
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------------------------------ MMA issue
    {
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t qa = smem_u32(sQ);
      auto keys_of = [&](int kb) { return min(64, S - 64 * kb); };
      auto issue_s = [&](int kb) {
        const int sl = kb % FL_RING;
        mbar_wait(&ring_full[sl], (uint32_t)((kb / FL_RING) & 1));
        tc_fence_after();
        const uint32_t ka = smem_u32(sRing + sl * 16384);
        const uint32_t idesc = make_idesc2(fmt, fmt, 128, (uint32_t)((keys_of(kb) + 15) & ~15), 0, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + (uint32_t)((kb & 1) * 64), make_smem_desc_sw128(qa + k * 32, 0u, 1024u),
                      make_smem_desc_sw128(ka + k * 32, 0u, 1024u), idesc, k > 0 ? 1u : 0u);
          umma_commit(&bar_s[kb & 1]);
        }
        __syncwarp();
      };
      const uint32_t idesc_o = make_idesc2(fmt, fmt, 128, 64, 0, 1);
      mbar_wait(bar_q, 0);
      issue_s(0);
      if (n_kb > 1) issue_s(1);
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(bar_p, (uint32_t)(kb & 1));
        tc_fence_after();
        const uint32_t va = smem_u32(sRing + (kb % FL_RING) * 16384 + 8192);
        const int ksteps = ((keys_of(kb) + 15) & ~15) / 16;
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            umma_f16_ts(tmem_base + FL_O_COL, tmem_base + (uint32_t)((kb & 1) * 64 + k * 8),
                        make_smem_desc_sw128(va + k * 2048, 8192u, 1024u), idesc_o, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&ring_free[kb % FL_RING]);  // K (used by S) and V of this block are spent
          umma_commit(&bar_o[kb & 1]);
        }
        __syncwarp();
        if (kb + 2 < n_kb) issue_s(kb + 2);     // reuses S buffer kb & 1: ordered behind the P.V just issued
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ softmax warps
    griddep_wait();
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;  // running maximum (raw score units) and sum for this thread's query row
    for (int kb = 0; kb < n_kb; ++kb) {
      const int keys = min(64, S - 64 * kb);
      const int nk = (keys + 15) & ~15;
      const uint32_t sbuf = trow + (uint32_t)((kb & 1) * 64);
      mbar_wait(&bar_s[kb & 1], (uint32_t)((kb >> 1) & 1));
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(sbuf, v0);
      if (nk > 32) tmem_ld32(sbuf + 32u, v1);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < keys) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(v0[k]));
      if (nk > 32) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (32 + k < keys) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(v1[k]));
      }
      // Lazy reference maximum: m_run (the value the exponents are taken against) moves only when this block's maximum
      // exceeds it by more than 8 in the log2 domain, so P stays <= 2^8 (exact enough in fp16 / bf16, fp32 accumulation)
      // and O - whose rescaling needs the previous P.V to have RETIRED - is touched, and that MMA waited for, only then.
      // The row sum l and the lse use the same reference, so the result is the same softmax.
      const float m_cand = fmaxf(m_run, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      const bool move = (m_cand - m_run) * scale_log2 > 8.0f;  // always on the first block (m_run = -inf)
      const float m_new = move ? m_cand : m_run;
      const float mc = m_new * scale_log2;
      const float alpha = move ? fast_exp2(m_run * scale_log2 - mc) : 1.0f;  // 0 on the first block
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float p0 = (2 * k < keys) ? fast_exp2(fmaf(__uint_as_float(v0[2 * k]), scale_log2, -mc)) : 0.f;
        const float p1 = (2 * k + 1 < keys) ? fast_exp2(fmaf(__uint_as_float(v0[2 * k + 1]), scale_log2, -mc)) : 0.f;
        s4[k & 3] += p0 + p1;
        pk[k] = F16 ? pack_f16(p0, p1) : pack_bf16(p0, p1);
      }
      if (nk > 32) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = (32 + 2 * k < keys) ? fast_exp2(fmaf(__uint_as_float(v1[2 * k]), scale_log2, -mc)) : 0.f;
          const float p1 = (33 + 2 * k < keys) ? fast_exp2(fmaf(__uint_as_float(v1[2 * k + 1]), scale_log2, -mc)) : 0.f;
          s4[k & 3] += p0 + p1;
          pk[16 + k] = F16 ? pack_f16(p0, p1) : pack_bf16(p0, p1);
        }
      } else {
#pragma unroll
        for (int k = 16; k < 32; ++k) pk[k] = 0u;
      }
      l_run = l_run * alpha + ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      m_run = m_new;
      // P (packed 16-bit) over the first 32 columns of this S buffer
      {
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { lo[k] = pk[k]; hi[k] = pk[16 + k]; }
        tmem_st16(sbuf, lo);
        if (nk > 32) tmem_st16(sbuf + 16u, hi);
      }
      if (kb > 0 && __any_sync(0xffffffffu, move)) {
        // O so far belongs to the old reference: wait for the previous P.V, rescale the rows whose reference moved
        mbar_wait(&bar_o[(kb - 1) & 1], (uint32_t)(((kb - 1) >> 1) & 1));
        tc_fence_after();
        {
          uint32_t o0[32], o1[32];
          tmem_ld32(trow + FL_O_COL, o0);
          tmem_ld32(trow + FL_O_COL + 32u, o1);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            o0[k] = __float_as_uint(__uint_as_float(o0[k]) * alpha);
            o1[k] = __float_as_uint(__uint_as_float(o1[k]) * alpha);
          }
          uint32_t t[16];
#pragma unroll
          for (int part = 0; part < 4; ++part) {
#pragma unroll
            for (int k = 0; k < 16; ++k) t[k] = part < 2 ? o0[(part & 1) * 16 + k] : o1[(part & 1) * 16 + k];
            tmem_st16(trow + FL_O_COL + (uint32_t)(part * 16), t);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue
    const int row = mt * 128 + warp * 32 + lane;
    if (row < S) lse[(long long)bh * S + row] = (m_run * scale_log2 + log2f(l_run)) * FA_LN2;
    const float inv = 1.f / l_run;
    mbar_wait(&bar_o[(n_kb - 1) & 1], (uint32_t)(((n_kb - 1) >> 1) & 1));
    tc_fence_after();
    float o[64];
    {
      uint32_t v0[32], v1[32];
      tmem_ld32(trow + FL_O_COL, v0);
      tmem_ld32(trow + FL_O_COL + 32u, v1);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) { o[k] = __uint_as_float(v0[k]) * inv; o[32 + k] = __uint_as_float(v1[k]) * inv; }
    }
    if (mt * 128 + warp * 32 < S) {
      uint8_t* st = sQ + warp * 4096;  // Q is spent: every S MMA has completed (the last bar_o follows them)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 w;
        if (F16)
          w = make_uint4(pack_f16(o[8 * j], o[8 * j + 1]), pack_f16(o[8 * j + 2], o[8 * j + 3]),
                         pack_f16(o[8 * j + 4], o[8 * j + 5]), pack_f16(o[8 * j + 6], o[8 * j + 7]));
        else
          w = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                         pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
        *reinterpret_cast<uint4*>(st + lane * 128 + ((j ^ (lane & 7)) << 4)) = w;
      }
      uint8_t* st2 = sO2 + warp * 4096;
      if (has_o2) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(st2 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                         pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        fa_tma_store_3d(&tmO, st, h * 64, mt * 128 + warp * 32, b);
        if (has_o2) fa_tma_store_3d(&tmO2, st2, h * 64, mt * 128 + warp * 32, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================================
// Forward for S <= 256, persistent: one CTA per SM walks over (image, head) items; the two 128-query tiles of an item
// are softmaxed by two independent warpgroups (warps 0..3 / 4..7), each with its own 256-column TMEM slot, so the Q K^T /
// P V MMAs of one tile run under the exp of the other.  K, V and both Q tiles of the NEXT item are prefetched by TMA into
// the second smem buffer while this one is being worked on; K / V are read once per head (the flash-style kernel reads
// them once per query tile).  The MMA warp is an event loop over the two slots (whichever tile has its operands ready
// is issued), the softmax is the single-shot two-pass form (scores never leave TMEM; P packed over S; O inside the spent
// S columns).  Warps whose 32 query rows all lie past S (the last quarter of the second tile at S = 197) skip the math.
constexpr int FP_THREADS = 320;     // warps 0..3: tile 0, 4..7: tile 1, 8: MMA issue, 9: TMA producer
constexpr uint32_t FP_SLOT = 256;   // TMEM columns per tile slot
constexpr uint32_t FP_O_COL = 128;

template <bool F16>
__global__ void __launch_bounds__(FP_THREADS, 1)
attn_fwd_tc_p2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                      float* __restrict__ lse, int S, int H, int NK, int items, float scale_log2, int has_o2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KV = NK * 128;                 // bytes of K (or V) of a head: [NK][64] 16-bit, 128B-swizzled
  const int BUF = 32768 + 2 * KV;          // Q tile 0 | Q tile 1 | K | V
  uint8_t* sO2 = smem + 2 * BUF;           // 8 warps x 4 KB staging of the optional bf16 copy of O
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO2 + 32768);
  uint64_t* qk_full = bars;                // [2] Q (both tiles) + K landed
  uint64_t* v_full = bars + 2;             // [2]
  uint64_t* buf_free = bars + 4;           // [2] every softmax warp is done with the buffer (O staged out of the Q
                                           //     tiles included) and both tiles' MMAs have retired
  uint64_t* s_full = bars + 6;             // [2 slots]
  uint64_t* p_full = bars + 8;             // [2] 4 warp arrivals
  uint64_t* o_full = bars + 10;            // [2]
  uint64_t* o_free = bars + 12;            // [2] 4 warp arrivals: O read out, the slot may take the next S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (S + 127) >> 7;
  const int n_local = items > (int)blockIdx.x ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmO2);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&qk_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&buf_free[i], (uint32_t)(5 * n_mt));
        mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  griddep_wait();

  if (warp == 9) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int k = 0; k < n_local; ++k) {
        const int buf = k & 1;
        const int bh = (int)blockIdx.x + k * (int)gridDim.x;
        const int b = bh / H, h = bh % H;
        if (k >= 2) mbar_wait(&buf_free[buf], (uint32_t)(((k >> 1) - 1) & 1));
        uint8_t* base = smem + buf * BUF;
        mbar_arrive_expect_tx(&qk_full[buf], (uint32_t)(n_mt * 16384 + KV));
        for (int mt = 0; mt < n_mt; ++mt) tma_load_3d(base + mt * 16384, &tmQ, &qk_full[buf], h * 64, mt * 128, b);
        tma_load_3d(base + 32768, &tmKV, &qk_full[buf], (H + h) * 64, 0, b);
        mbar_arrive_expect_tx(&v_full[buf], (uint32_t)KV);
        tma_load_3d(base + 32768 + KV, &tmKV, &v_full[buf], (2 * H + h) * 64, 0, b);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------------------------------ MMA issue: event loop
    if (lane == 0) {
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc_s = make_idesc2(fmt, fmt, 128, (uint32_t)NK, 0, 0);
      const uint32_t idesc_o = make_idesc2(fmt, fmt, 128, 64, 0, 1);
      int ks[2] = {0, n_mt > 1 ? 0 : n_local};
      int st[2] = {0, 0};
      const long long t_start = clock64();
      while (ks[0] < n_local || ks[1] < n_local) {
        if (clock64() - t_start > 4000000000LL) {  // a protocol bug must trap instead of hanging the GPU box
          printf("mfvit: attention MMA loop timeout (block %d)\n", blockIdx.x);
          __trap();
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int k = ks[t];
          if (k >= n_local) continue;
          const int buf = k & 1;
          const uint32_t use_ph = (uint32_t)((k >> 1) & 1);
          uint8_t* base = smem + buf * BUF;
          const uint32_t slot = tmem_base + (uint32_t)t * FP_SLOT;
          if (st[t] == 0) {
            if (!mbar_test_wait(&qk_full[buf], use_ph)) continue;
            if (k > 0 && !mbar_test_wait(&o_free[t], (uint32_t)((k - 1) & 1))) continue;
            tc_fence_after();
            const uint32_t qa = smem_u32(base + t * 16384), ka = smem_u32(base + 32768);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(slot, make_smem_desc_sw128(qa + kk * 32, 0u, 1024u), make_smem_desc_sw128(ka + kk * 32, 0u, 1024u),
                        idesc_s, kk > 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
            st[t] = 1;
          } else {
            if (!mbar_test_wait(&p_full[t], (uint32_t)(k & 1))) continue;
            if (!mbar_test_wait(&v_full[buf], use_ph)) continue;
            tc_fence_after();
            const uint32_t va = smem_u32(base + 32768 + KV);
            for (int kk = 0; kk < NK / 16; ++kk)
              umma_f16_ts(slot + FP_O_COL, slot + (uint32_t)(kk * 8), make_smem_desc_sw128(va + kk * 2048, 8192u, 1024u),
                          idesc_o, kk > 0 ? 1u : 0u);
            umma_commit(&o_full[t]);
            umma_commit(&buf_free[buf]);
            st[t] = 0;
            ks[t] = k + 1;
          }
        }
      }
    }
  } else if ((warp >> 2) < n_mt) {
    // ------------------------------------------------------------------------------------------ softmax + epilogue
    const int t = warp >> 2, w = warp & 3;
    const uint32_t trow = tmem_base + (uint32_t)t * FP_SLOT + ((uint32_t)(w * 32) << 16);
    const bool skip = t * 128 + w * 32 >= S;  // warp-uniform: no valid query row in this warp's quarter
    uint8_t* st2 = sO2 + warp * 4096;
    const int nch = (NK + 31) >> 5;
    auto ld_chunk = [&](int c, uint32_t (&v)[32]) {
      if (c * 32 + 32 <= NK) tmem_ld32(trow + (uint32_t)(c * 32), v); else tmem_ld16(trow + (uint32_t)(c * 32), v);
    };
    for (int k = 0; k < n_local; ++k) {
      const int buf = k & 1;
      const uint32_t ph = (uint32_t)(k & 1);
      const int bh = (int)blockIdx.x + k * (int)gridDim.x;
      const int b = bh / H, h = bh % H;
      mbar_wait(&s_full[t], ph);
      tc_fence_after();
      if (k > 0) {
        // the stores of the previous item have long finished reading their staging (its Q tiles, sO2): hand that buffer
        // back to the producer
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&buf_free[buf ^ 1]);
        }
        __syncwarp();
      }
      float inv = 0.f;
      if (!skip) {
        // ---- pass 1: row maximum (columns >= S are padding)
        float mx = -INFINITY;
        {
          uint32_t va[32], vb[32];
          ld_chunk(0, va);
          auto red = [&](int c, const uint32_t (&v)[32]) {
            const int c0 = c * 32;
            if (c0 + 32 <= S) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else {
              const bool full = c0 + 32 <= NK;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < S && (full || j < 16)) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
          };
          for (int c = 0; c < nch; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nch) ld_chunk(c + 1, vb);
            red(c, va);
            if (c + 1 < nch) {
              tmem_ld_wait();
              if (c + 2 < nch) ld_chunk(c + 2, va);
              red(c + 1, vb);
            }
          }
        }
        const float mc = mx * scale_log2;
        // ---- pass 2: p = exp2(s * c - max * c), row sum, packed 16-bit P written back over S
        float sum = 0.f;
        {
          uint32_t va[32], vb[32];
          ld_chunk(0, va);
          auto proc = [&](int c, const uint32_t (&v)[32]) {
            const int c0 = c * 32;
            const bool full = c0 + 32 <= NK;
            float pr[32];
            if (c0 + 32 <= S) {
#pragma unroll
              for (int j = 0; j < 32; ++j) pr[j] = fast_exp2(fmaf(__uint_as_float(v[j]), scale_log2, -mc));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                pr[j] = (c0 + j < S && (full || j < 16)) ? fast_exp2(fmaf(__uint_as_float(v[j]), scale_log2, -mc)) : 0.f;
            }
            uint32_t pk[16];
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              s4[j & 3] += pr[2 * j] + pr[2 * j + 1];
              pk[j] = F16 ? pack_f16(pr[2 * j], pr[2 * j + 1]) : pack_bf16(pr[2 * j], pr[2 * j + 1]);
            }
            sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
            if (full) tmem_st16(trow + (uint32_t)(c0 >> 1), pk); else tmem_st8(trow + (uint32_t)(c0 >> 1), pk);
          };
          for (int c = 0; c < nch; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nch) ld_chunk(c + 1, vb);
            proc(c, va);
            if (c + 1 < nch) {
              tmem_ld_wait();
              if (c + 2 < nch) ld_chunk(c + 2, va);
              proc(c + 1, vb);
            }
          }
        }
        tmem_st_wait();
        const int row = t * 128 + w * 32 + lane;
        if (row < S) lse[(long long)bh * S + row] = (mc + log2f(sum)) * FA_LN2;
        inv = 1.f / sum;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      // ---- epilogue: O / sum -> 16-bit -> swizzled smem (this tile's spent Q rows) -> TMA store
      mbar_wait(&o_full[t], ph);
      tc_fence_after();
      if (skip) {
        if (lane == 0) mbar_arrive(&o_free[t]);
        continue;
      }
      float o[64];
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + FP_O_COL, v0);
        tmem_ld32(trow + FP_O_COL + 32u, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(v0[j]) * inv; o[32 + j] = __uint_as_float(v1[j]) * inv; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      uint8_t* stg = smem + buf * BUF + t * 16384 + w * 4096;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 wv;
        if (F16)
          wv = make_uint4(pack_f16(o[8 * j], o[8 * j + 1]), pack_f16(o[8 * j + 2], o[8 * j + 3]),
                          pack_f16(o[8 * j + 4], o[8 * j + 5]), pack_f16(o[8 * j + 6], o[8 * j + 7]));
        else
          wv = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                          pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = wv;
      }
      if (has_o2) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(st2 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                         pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        fa_tma_store_3d(&tmO, stg, h * 64, t * 128 + w * 32, b);
        if (has_o2) fa_tma_store_3d(&tmO2, st2, h * 64, t * 128 + w * 32, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int fa_encode(CUtensorMap* map, const void* base, int is_f16, long long inner, long long rows, long long images,
                     int box_rows, int box_cols = 64) {
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)images};
  cuuint64_t strides[2] = {(cuuint64_t)inner * 2, (cuuint64_t)(inner * rows) * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15)) return MFV_ERR_ALIGN;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MFV_ERR_INIT;
  CUresult r = enc(map, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFV_OK : MFV_ERR_ARG;
}

// Host entry used by mfv_attn_fwd (attn.cu) for D == 64 and S <= 256.
int attn_fwd_tc(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse, long long NB,
                long long S, long long H, float scale, cudaStream_t st) {
  if (o_is_f16 != qkv_is_f16) return MFV_ERR_ARG;  // P / O follow the operand format
  const int NK = ((int)S + 15) & ~15;
  CUtensorMap tmQ, tmKV, tmO, tmO2;
  int rc;
  if ((rc = fa_encode(&tmQ, qkv, qkv_is_f16, 3 * H * 64, S, NB, 128))) return rc;
  if ((rc = fa_encode(&tmKV, qkv, qkv_is_f16, 3 * H * 64, S, NB, NK))) return rc;
  if ((rc = fa_encode(&tmO, o, o_is_f16, H * 64, S, NB, 32))) return rc;
  tmO2 = tmO;
  if (o_bf16_copy && (rc = fa_encode(&tmO2, o_bf16_copy, 0, H * 64, S, NB, 32))) return rc;
  const size_t smem = 1024 + 32768 + 2 * (size_t)NK * 128 + 16384 + 64;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(NB * H));
  cfg.blockDim = dim3(FA_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const float sl2 = scale * FA_LOG2E;
  const int has_o2 = o_bf16_copy != nullptr;
  if (qkv_is_f16) {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_tc_kernel<true>, tmQ, tmKV, tmO, tmO2, lse, (int)S, (int)H, NK, sl2,
                                      has_o2));
  } else {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_tc_kernel<false>, tmQ, tmKV, tmO, tmO2, lse, (int)S, (int)H, NK, sl2,
                                      has_o2));
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// Host entry for S <= 256: the persistent two-warpgroup kernel (one CTA per SM).
int attn_fwd_tc_p2(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse, long long NB,
                   long long S, long long H, float scale, cudaStream_t st) {
  if (o_is_f16 != qkv_is_f16) return MFV_ERR_ARG;
  const int NK = ((int)S + 15) & ~15;
  CUtensorMap tmQ, tmKV, tmO, tmO2;
  int rc;
  if ((rc = fa_encode(&tmQ, qkv, qkv_is_f16, 3 * H * 64, S, NB, 128))) return rc;
  if ((rc = fa_encode(&tmKV, qkv, qkv_is_f16, 3 * H * 64, S, NB, NK))) return rc;
  if ((rc = fa_encode(&tmO, o, o_is_f16, H * 64, S, NB, 32))) return rc;
  tmO2 = tmO;
  if (o_bf16_copy && (rc = fa_encode(&tmO2, o_bf16_copy, 0, H * 64, S, NB, 32))) return rc;
  const size_t smem = 1024 + 2 * (32768 + 2 * (size_t)NK * 128) + 32768 + 256;
  const long long items = NB * H;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min<long long>(items, num_sms()));
  cfg.blockDim = dim3(FP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const float sl2 = scale * FA_LOG2E;
  const int has_o2 = o_bf16_copy != nullptr;
  if (qkv_is_f16) {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_p2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_tc_p2_kernel<true>, tmQ, tmKV, tmO, tmO2, lse, (int)S, (int)H, NK,
                                      (int)items, sl2, has_o2));
  } else {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_p2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_tc_p2_kernel<false>, tmQ, tmKV, tmO, tmO2, lse, (int)S, (int)H, NK,
                                      (int)items, sl2, has_o2));
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// Host entry used by mfv_attn_bwd (attn.cu) for D == 64 and S <= 224.  o and d_o are bf16.
int attn_bwd_tc(const void* qkv, int qkv_is_f16, const void* o, const void* d_o, const float* lse, void* dqkv,
                long long NB, long long S, long long H, float scale, cudaStream_t st) {
  const int NK = ((int)S + 15) & ~15;
  CUtensorMap tmQKV, tmDO, tmG;
  int rc;
  if ((rc = fa_encode(&tmQKV, qkv, qkv_is_f16, 3 * H * 64, S, NB, NK))) return rc;
  if ((rc = fa_encode(&tmDO, d_o, 0, H * 64, S, NB, NK))) return rc;
  if ((rc = fa_encode(&tmG, dqkv, 0, 3 * H * 64, S, NB, 32, 32))) return rc;
  const size_t smem = 1024 + 4 * (size_t)NK * 128 + 65536 + FB_CWARPS * 2048 + 2048 + 128;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(NB * H));
  cfg.blockDim = dim3(FB_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const __nv_bfloat16* op = reinterpret_cast<const __nv_bfloat16*>(o);
  // measurement aid: MFVIT_ATTN_PROF = device address (decimal) of 16 x 8 + 1 uint64 counters (tests/gpu_attn_prof.py)
  unsigned long long* prof = nullptr;
  if (const char* e = getenv("MFVIT_ATTN_PROF")) prof = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 10));
  if (qkv_is_f16) {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_bwd_tc_kernel<true>, tmQKV, tmDO, tmG, op, lse, (int)S, (int)H, NK, scale, prof));
  } else {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_bwd_tc_kernel<false>, tmQKV, tmDO, tmG, op, lse, (int)S, (int)H, NK, scale, prof));
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// Long sequences (S > 224): key-block CTAs, dQ summed in fp32 through dq_ws (f32 [NB][S][H*64], zeroed here) and
// converted afterwards; delta is the caller's f32 [NB][H][S] scratch.
int attn_bwd_tc_mb(const void* qkv, int qkv_is_f16, const void* o, const void* d_o, const float* lse, float* delta,
                   float* dq_ws, void* dqkv, long long NB, long long S, long long H, float scale, cudaStream_t st) {
  CUtensorMap tmQKV, tmDO, tmG, tmDQ;
  int rc;
  if ((rc = fa_encode(&tmQKV, qkv, qkv_is_f16, 3 * H * 64, S, NB, 128))) return rc;
  if ((rc = fa_encode(&tmDO, d_o, 0, H * 64, S, NB, 128))) return rc;
  if ((rc = fa_encode(&tmG, dqkv, 0, 3 * H * 64, S, NB, 32, 32))) return rc;
  {  // fp32 [NB][S][H*64], box 16 columns (64 B) x 32 rows, 64B swizzle
    cuuint64_t dims[3] = {(cuuint64_t)(H * 64), (cuuint64_t)S, (cuuint64_t)NB};
    cuuint64_t strides[2] = {(cuuint64_t)(H * 64) * 4, (cuuint64_t)(H * 64 * S) * 4};
    cuuint32_t box[3] = {16, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (reinterpret_cast<uintptr_t>(dq_ws) & 15) return MFV_ERR_ALIGN;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return MFV_ERR_INIT;
    if (enc(&tmDQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dq_ws, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MFV_ERR_ARG;
  }
  const long long rows = NB * S;
  if ((rc = mfv_fill_f32(dq_ws, 0.f, rows * H * 64, st))) return rc;
  attn_delta_kernel<<<(unsigned)((rows * H + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), delta, rows, (int)S, (int)H);
  MFV_LAUNCH_CHECK();
  const size_t smem = 1024 + 6 * 16384 + 65536 + FB_CWARPS * 2048 + 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(NB * H), (unsigned)((S + 127) / 128));
  cfg.blockDim = dim3(FM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (qkv_is_f16) {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_mb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_bwd_tc_mb_kernel<true>, tmQKV, tmDO, tmG, tmDQ, lse, (const float*)delta,
                                      (int)S, (int)H, scale));
  } else {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_mb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_bwd_tc_mb_kernel<false>, tmQKV, tmDO, tmG, tmDQ, lse, (const float*)delta,
                                      (int)S, (int)H, scale));
  }
  MFV_LAUNCH_CHECK();
  attn_dq_convert_kernel<<<(unsigned)((rows * H * 8 + 255) / 256), 256, 0, st>>>(
      dq_ws, reinterpret_cast<__nv_bfloat16*>(dqkv), rows, (int)H);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// Long sequences (S > 256): flash-style kernel, one CTA per (image, head, 128-query tile).
int attn_fwd_tc_mb(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse, long long NB,
                   long long S, long long H, float scale, cudaStream_t st) {
  if (o_is_f16 != qkv_is_f16) return MFV_ERR_ARG;
  CUtensorMap tmQ, tmKV, tmO, tmO2;
  int rc;
  if ((rc = fa_encode(&tmQ, qkv, qkv_is_f16, 3 * H * 64, S, NB, 128))) return rc;
  if ((rc = fa_encode(&tmKV, qkv, qkv_is_f16, 3 * H * 64, S, NB, 64))) return rc;
  if ((rc = fa_encode(&tmO, o, o_is_f16, H * 64, S, NB, 32))) return rc;
  tmO2 = tmO;
  if (o_bf16_copy && (rc = fa_encode(&tmO2, o_bf16_copy, 0, H * 64, S, NB, 32))) return rc;
  const size_t smem = 1024 + 16384 + FL_RING * 16384 + 16384 + 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(NB * H), (unsigned)((S + 127) / 128));
  cfg.blockDim = dim3(FL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const float sl2 = scale * FA_LOG2E;
  const int has_o2 = o_bf16_copy != nullptr;
  if (qkv_is_f16) {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_mb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_tc_mb_kernel<true>, tmQ, tmKV, tmO, tmO2, lse, (int)S, (int)H, sl2, has_o2));
  } else {
    static bool set = false;
    if (!set) {
      MFV_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_mb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      set = true;
    }
    MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_tc_mb_kernel<false>, tmQ, tmKV, tmO, tmO2, lse, (int)S, (int)H, sl2, has_o2));
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

}  // namespace mfv
