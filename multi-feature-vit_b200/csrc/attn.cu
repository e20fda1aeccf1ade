// Fused softmax self-attention, forward and backward, for the ViT-S/16 block: S = 197 (224^2) or 577
// (384^2) tokens, head_dim 64 (or 32).  The whole K/V (forward, dQ pass) or Q/dO (dK/dV pass) of one head stays
// resident in shared memory; each warp owns 16 rows and keeps its softmax statistics in registers (warp-level online
// softmax, quad shuffles only).  Scores never touch HBM; the backward recomputes them from the saved log-sum-exp.
// Tensor-core path: mma.sync.m16n8k16 bf16 with ldmatrix from XOR-swizzled smem (the attention FLOPs are <8 % of the
// step; the dense contractions live in gemm.cu on tcgen05).
#include <stdlib.h>
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int ATT_WARPS = 8;
constexpr int ATT_ROWS = ATT_WARPS * 16;  // 128 rows per CTA pass (one 16-row MMA slab per warp)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma16816_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ void mma_any(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16) mma16816_f16(c, a, b0, b1); else mma16816(c, a, b0, b1);
}

// Shared-memory tile [rows][D] bf16, 16-byte chunks XOR-swizzled so ldmatrix rows hit distinct banks.
template <int D>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  if constexpr (D == 64) {
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
  } else {  // D == 32: two rows per 128-byte line
    return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
  }
}

// Load `nrows` rows (starting at token s0) of one [S][.] slice into a swizzled tile; rows >= S are zero-filled.
template <int D>
__device__ __forceinline__ void load_tile(uint8_t* tile, const __nv_bfloat16* base, long long row_stride, int s0,
                                          int nrows, int S) {
  constexpr int CH = D / 8;
  for (int i = threadIdx.x; i < nrows * CH; i += ATT_WARPS * 32) {
    const int r = i / CH, c = i % CH;
    const int s = s0 + r;
    const bool ok = s < S;
    cp_async16(tile + tile_off<D>(r, c), base + (long long)(ok ? s : 0) * row_stride + c * 8, ok);
  }
}

// acc[j] (16 x 8 per n-tile j) += A(16 x D, register fragments) * Bt, with Bt rows = "n" index taken from a
// [n][D] row-major tile (ldmatrix, no transpose).  Used for Q K^T, dO V^T, K Q^T, V dO^T.
template <int D, bool F16 = false, int NP = 4>  // NP: number of 16-row groups of the tile actually multiplied (tail trimming)
__device__ __forceinline__ void mma_a_bt(float (&acc)[8][4], const uint32_t (&afrag)[D / 16][4], const uint8_t* tile,
                                         int n_row0, int lane) {
  const uint32_t tbase = smem_u32(tile);
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
    for (int jp = 0; jp < NP; ++jp) {
      const int row = n_row0 + jp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
      const int chunk = 2 * kk + ((lane >> 3) & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tbase + tile_off<D>(row, chunk), b0, b1, b2, b3);
      mma_any<F16>(acc[2 * jp], afrag[kk], b0, b1);
      mma_any<F16>(acc[2 * jp + 1], afrag[kk], b2, b3);
    }
  }
}

// out[j] (16 x 8 per d-tile j, D/8 tiles) += P(16 x 64, from accumulator registers, rounded to bf16) * B, with B rows =
// reduction index taken from a [k][D] row-major tile (ldmatrix.trans).  Used for P V, dS K, P^T dO, dS^T Q.
template <int D, bool F16 = false, int NP = 4>  // NP: number of 16-deep reduction steps actually taken
__device__ __forceinline__ void mma_p_b(float (&out)[D / 8][4], const float (&p)[8][4], const uint8_t* tile,
                                        int k_row0, int lane) {
  const uint32_t tbase = smem_u32(tile);
#pragma unroll
  for (int kk = 0; kk < NP; ++kk) {
    uint32_t a[4];
    if constexpr (F16) {
      a[0] = pack_f16(p[2 * kk][0], p[2 * kk][1]);
      a[1] = pack_f16(p[2 * kk][2], p[2 * kk][3]);
      a[2] = pack_f16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
      a[3] = pack_f16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    } else {
      a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
      a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
      a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
      a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    }
#pragma unroll
    for (int jp = 0; jp < D / 16; ++jp) {
      const int row = k_row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int chunk = 2 * jp + ((lane >> 4) & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tbase + tile_off<D>(row, chunk), b0, b1, b2, b3);
      mma_any<F16>(out[2 * jp], a, b0, b1);
      mma_any<F16>(out[2 * jp + 1], a, b2, b3);
    }
  }
}

// In-place fp16 -> bf16 conversion of a shared-memory tile (elementwise, so the swizzle is irrelevant).  The forward
// keeps q,k,v in fp16 (finer mantissa); the backward runs in bf16 (gradient range) and converts after the load.
__device__ __forceinline__ void tile_f16_to_bf16(uint8_t* tile, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += ATT_WARPS * 32 * 16) {
    uint4 v = *reinterpret_cast<uint4*>(tile + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_f16(w[j]);
      w[j] = pack_bf16(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(tile + i) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// A fragments (16 rows x D) of this warp's rows from a swizzled tile.
template <int D>
__device__ __forceinline__ void load_afrag(uint32_t (&afrag)[D / 16][4], const uint8_t* tile, int row0, int lane) {
  const uint32_t tbase = smem_u32(tile);
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const int row = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int chunk = 2 * kk + (lane >> 4);
    ldsm_x4(tbase + tile_off<D>(row, chunk), afrag[kk][0], afrag[kk][1], afrag[kk][2], afrag[kk][3]);
  }
}

// Write this warp's 16 x D fp32 fragments as bf16 rows (row stride in elements); rows >= S are skipped.
template <int D, bool F16 = false>
__device__ __forceinline__ void store_rows(const float (&acc)[D / 8][4], __nv_bfloat16* base, long long row_stride,
                                           int s_row0, int S, int lane, float mul0, float mul1) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = s_row0 + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < D / 8; ++j) {
    if (r0 < S)
      *reinterpret_cast<uint32_t*>(base + (long long)r0 * row_stride + j * 8 + 2 * t) =
          F16 ? pack_f16(acc[j][0] * mul0, acc[j][1] * mul0) : pack_bf16(acc[j][0] * mul0, acc[j][1] * mul0);
    if (r1 < S)
      *reinterpret_cast<uint32_t*>(base + (long long)r1 * row_stride + j * 8 + 2 * t) =
          F16 ? pack_f16(acc[j][2] * mul1, acc[j][3] * mul1) : pack_bf16(acc[j][2] * mul1, acc[j][3] * mul1);
  }
}

// ------------------------------------------------------------------------------------------------ forward
// grid = (ceil(S/q_rows), NB*H).  smem: K[Spad][D], V[Spad][D], Q[q_rows][D]; q_rows is a multiple of 128 and the CTA
// walks it in passes of 128 rows (8 warps x 16), so K/V of a head are fetched once (S=197) or ceil(S/128) times.
template <int D, bool F16>
__global__ void __launch_bounds__(ATT_WARPS * 32, 2)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ o, int o_is_f16,
                __nv_bfloat16* __restrict__ o_bf, float* __restrict__ lse, int S, int H, float scale_log2, int q_rows) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Spad = (S + 63) & ~63;
  uint8_t* sK = smem;
  uint8_t* sV = sK + Spad * D * 2;
  uint8_t* sQ = sV + Spad * D * 2;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rs = 3LL * H * D;
  const __nv_bfloat16* qb = qkv + (long long)b * S * rs + h * D;
  load_tile<D>(sQ, qb, rs, blockIdx.x * q_rows, q_rows, S);
  load_tile<D>(sK, qb + H * D, rs, 0, Spad, S);
  load_tile<D>(sV, qb + 2 * H * D, rs, 0, Spad, S);
  cp_async_wait_all();
  __syncthreads();

  for (int pass = 0; pass < q_rows / ATT_ROWS; ++pass) {
  const int q0 = blockIdx.x * q_rows + pass * ATT_ROWS;
  if (q0 + warp * 16 >= S) break;  // warp-uniform; no block-level sync below
  uint32_t qf[D / 16][4];
  load_afrag<D>(qf, sQ, pass * ATT_ROWS + warp * 16, lane);
  float oacc[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int t = lane & 3;

  for (int kb = 0; kb < S; kb += 64) {
    const int np = min(4, (S - kb + 15) >> 4);  // 16-key groups holding valid keys in this block (tail: S=197 -> 1)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    if (np == 4) mma_a_bt<D, F16, 4>(s, qf, sK, kb, lane);
    else if (np == 3) mma_a_bt<D, F16, 3>(s, qf, sK, kb, lane);
    else if (np == 2) mma_a_bt<D, F16, 2>(s, qf, sK, kb, lane);
    else mma_a_bt<D, F16, 1>(s, qf, sK, kb, lane);
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kb + j * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = (key + (e & 1) < S) ? s[j][e] * scale_log2 : -INFINITY;
        s[j][e] = v;
      }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float a0 = exp2f(m0 - mx0), a1 = exp2f(m1 - mx1);  // first block: exp2(-inf) = 0
    m0 = mx0; m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] - m0); s[j][1] = exp2f(s[j][1] - m0);
      s[j][2] = exp2f(s[j][2] - m1); s[j][3] = exp2f(s[j][3] - m1);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      oacc[j][0] *= a0; oacc[j][1] *= a0; oacc[j][2] *= a1; oacc[j][3] *= a1;
    }
    if (np == 4) mma_p_b<D, F16, 4>(oacc, s, sV, kb, lane);
    else if (np == 3) mma_p_b<D, F16, 3>(oacc, s, sV, kb, lane);
    else if (np == 2) mma_p_b<D, F16, 2>(oacc, s, sV, kb, lane);
    else mma_p_b<D, F16, 1>(oacc, s, sV, kb, lane);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int row0 = q0 + warp * 16;
  __nv_bfloat16* ob = o + (long long)b * S * H * D + h * D;
  if (o_is_f16) store_rows<D, true>(oacc, ob, (long long)H * D, row0, S, lane, 1.f / l0, 1.f / l1);
  else store_rows<D, false>(oacc, ob, (long long)H * D, row0, S, lane, 1.f / l0, 1.f / l1);
  if (o_bf) store_rows<D, false>(oacc, o_bf + (long long)b * S * H * D + h * D, (long long)H * D, row0, S, lane, 1.f / l0, 1.f / l1);
  if (t == 0) {
    const int g = lane >> 2;
    float* lb = lse + (long long)bh * S;
    if (row0 + g < S) lb[row0 + g] = (m0 + log2f(l0)) * LN2;
    if (row0 + g + 8 < S) lb[row0 + g + 8] = (m1 + log2f(l1)) * LN2;
  }
  }  // pass
}

// ------------------------------------------------------------------------------------------------ backward: dQ
// grid = (ceil(S/128), NB*H).  smem: K[Spad][D], V[Spad][D], Q[128][D], dO[128][D].  Also emits delta = rowsum(dO*O).
template <int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o,
                   const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse, float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, int S, int H, float scale, int qkv_is_f16) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Spad = (S + 63) & ~63;
  uint8_t* sK = smem;
  uint8_t* sV = sK + Spad * D * 2;
  uint8_t* sQ = sV + Spad * D * 2;
  uint8_t* sdO = sQ + ATT_ROWS * D * 2;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * ATT_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rs = 3LL * H * D, ors = (long long)H * D;
  const __nv_bfloat16* qb = qkv + (long long)b * S * rs + h * D;
  const __nv_bfloat16* dob = d_o + (long long)b * S * ors + h * D;
  const __nv_bfloat16* ob = o + (long long)b * S * ors + h * D;
  load_tile<D>(sQ, qb, rs, q0, ATT_ROWS, S);
  load_tile<D>(sdO, dob, ors, q0, ATT_ROWS, S);
  load_tile<D>(sK, qb + H * D, rs, 0, Spad, S);
  load_tile<D>(sV, qb + 2 * H * D, rs, 0, Spad, S);

  // delta for this warp's 16 rows: 2 lanes per row, D/2 elements each (straight from global)
  const int row0 = q0 + warp * 16;
  float dl;
  {
    const int r = row0 + (lane >> 1);
    float acc = 0.f;
    if (r < S) {
      const uint4* po = reinterpret_cast<const uint4*>(ob + (long long)r * ors + (lane & 1) * (D / 2));
      const uint4* pd = reinterpret_cast<const uint4*>(dob + (long long)r * ors + (lane & 1) * (D / 2));
#pragma unroll
      for (int i = 0; i < D / 16; ++i) {
        const uint4 a = __ldg(po + i), c = __ldg(pd + i);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16(aw[j]), y = unpack_bf16(cw[j]);
          acc += x.x * y.x + x.y * y.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    dl = acc;  // lanes 2r, 2r+1 hold delta of row r
    if ((lane & 1) == 0 && r < S) delta[(long long)bh * S + r] = acc;
  }
  const int g = lane >> 2, t = lane & 3;
  const float dl0 = __shfl_sync(0xffffffffu, dl, 2 * g), dl1 = __shfl_sync(0xffffffffu, dl, 2 * (g + 8));
  const float* lb = lse + (long long)bh * S;
  const float lse0 = (row0 + g < S) ? lb[row0 + g] * LOG2E : INFINITY;
  const float lse1 = (row0 + g + 8 < S) ? lb[row0 + g + 8] * LOG2E : INFINITY;
  const float scale_log2 = scale * LOG2E;

  cp_async_wait_all();
  __syncthreads();
  if (qkv_is_f16) {
    tile_f16_to_bf16(sK, (2 * Spad + ATT_ROWS) * D * 2);  // sK, sV, sQ are contiguous
    __syncthreads();
  }
  uint32_t qf[D / 16][4], dof[D / 16][4];
  load_afrag<D>(qf, sQ, warp * 16, lane);
  load_afrag<D>(dof, sdO, warp * 16, lane);
  float dq[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;

  for (int kb = 0; kb < S; kb += 64) {
    const int np = min(4, (S - kb + 15) >> 4);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
    if (np == 4) { mma_a_bt<D, false, 4>(s, qf, sK, kb, lane); mma_a_bt<D, false, 4>(dp, dof, sV, kb, lane); }
    else if (np == 3) { mma_a_bt<D, false, 3>(s, qf, sK, kb, lane); mma_a_bt<D, false, 3>(dp, dof, sV, kb, lane); }
    else if (np == 2) { mma_a_bt<D, false, 2>(s, qf, sK, kb, lane); mma_a_bt<D, false, 2>(dp, dof, sV, kb, lane); }
    else { mma_a_bt<D, false, 1>(s, qf, sK, kb, lane); mma_a_bt<D, false, 1>(dp, dof, sV, kb, lane); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kb + j * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = key + (e & 1) < S;
        const float p = ok ? exp2f(s[j][e] * scale_log2 - (e < 2 ? lse0 : lse1)) : 0.f;
        s[j][e] = p * (dp[j][e] - (e < 2 ? dl0 : dl1)) * scale;  // dS
      }
    }
    if (np == 4) mma_p_b<D, false, 4>(dq, s, sK, kb, lane);
    else if (np == 3) mma_p_b<D, false, 3>(dq, s, sK, kb, lane);
    else if (np == 2) mma_p_b<D, false, 2>(dq, s, sK, kb, lane);
    else mma_p_b<D, false, 1>(dq, s, sK, kb, lane);
  }
  __nv_bfloat16* dqb = dqkv + (long long)b * S * rs + h * D;
  store_rows<D>(dq, dqb, rs, row0, S, lane, 1.f, 1.f);
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
// grid = (ceil(S/128), NB*H) over key tiles.  smem: Q[Spad][D], dO[Spad][D], K[128][D], V[128][D], lse2[Spad], delta[Spad].
template <int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_o,
                    const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                    int S, int H, float scale, int qkv_is_f16) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Spad = (S + 63) & ~63;
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + Spad * D * 2;
  uint8_t* sK = sdO + Spad * D * 2;
  uint8_t* sV = sK + ATT_ROWS * D * 2;
  float* sLse = reinterpret_cast<float*>(sV + ATT_ROWS * D * 2);
  float* sDel = sLse + Spad;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int k0 = blockIdx.x * ATT_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rs = 3LL * H * D, ors = (long long)H * D;
  const __nv_bfloat16* qb = qkv + (long long)b * S * rs + h * D;
  const __nv_bfloat16* dob = d_o + (long long)b * S * ors + h * D;
  load_tile<D>(sK, qb + H * D, rs, k0, ATT_ROWS, S);
  load_tile<D>(sV, qb + 2 * H * D, rs, k0, ATT_ROWS, S);
  load_tile<D>(sQ, qb, rs, 0, Spad, S);
  load_tile<D>(sdO, dob, ors, 0, Spad, S);
  for (int i = threadIdx.x; i < Spad; i += ATT_WARPS * 32) {
    sLse[i] = i < S ? lse[(long long)bh * S + i] * LOG2E : INFINITY;  // +inf -> p = 0 for padded queries
    sDel[i] = i < S ? delta[(long long)bh * S + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  if (qkv_is_f16) {
    tile_f16_to_bf16(sQ, Spad * D * 2);
    tile_f16_to_bf16(sK, 2 * ATT_ROWS * D * 2);  // sK, sV contiguous
    __syncthreads();
  }

  uint32_t kf[D / 16][4], vf[D / 16][4];
  load_afrag<D>(kf, sK, warp * 16, lane);
  load_afrag<D>(vf, sV, warp * 16, lane);
  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  const int t = lane & 3;
  const float scale_log2 = scale * LOG2E;

  for (int qb0 = 0; qb0 < S; qb0 += 64) {
    const int np = min(4, (S - qb0 + 15) >> 4);  // 16-query groups with valid queries (padded queries have p = 0)
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
    }
    // S^T = K Q^T (rows: keys, cols: queries), dP^T = V dO^T
    if (np == 4) { mma_a_bt<D, false, 4>(st, kf, sQ, qb0, lane); mma_a_bt<D, false, 4>(dpt, vf, sdO, qb0, lane); }
    else if (np == 3) { mma_a_bt<D, false, 3>(st, kf, sQ, qb0, lane); mma_a_bt<D, false, 3>(dpt, vf, sdO, qb0, lane); }
    else if (np == 2) { mma_a_bt<D, false, 2>(st, kf, sQ, qb0, lane); mma_a_bt<D, false, 2>(dpt, vf, sdO, qb0, lane); }
    else { mma_a_bt<D, false, 1>(st, kf, sQ, qb0, lane); mma_a_bt<D, false, 1>(dpt, vf, sdO, qb0, lane); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int qi = qb0 + j * 8 + 2 * t;
      const float ls0 = sLse[qi], ls1 = sLse[qi + 1], de0 = sDel[qi], de1 = sDel[qi + 1];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(st[j][e] * scale_log2 - ((e & 1) ? ls1 : ls0));
        st[j][e] = p;                                                   // P^T
        dpt[j][e] = p * (dpt[j][e] - ((e & 1) ? de1 : de0)) * scale;     // dS^T
      }
    }
    // dV += P^T dO ; dK += dS^T Q
    if (np == 4) { mma_p_b<D, false, 4>(dv, st, sdO, qb0, lane); mma_p_b<D, false, 4>(dk, dpt, sQ, qb0, lane); }
    else if (np == 3) { mma_p_b<D, false, 3>(dv, st, sdO, qb0, lane); mma_p_b<D, false, 3>(dk, dpt, sQ, qb0, lane); }
    else if (np == 2) { mma_p_b<D, false, 2>(dv, st, sdO, qb0, lane); mma_p_b<D, false, 2>(dk, dpt, sQ, qb0, lane); }
    else { mma_p_b<D, false, 1>(dv, st, sdO, qb0, lane); mma_p_b<D, false, 1>(dk, dpt, sQ, qb0, lane); }
  }
  __nv_bfloat16* dkb = dqkv + (long long)b * S * rs + H * D + h * D;
  __nv_bfloat16* dvb = dqkv + (long long)b * S * rs + 2 * H * D + h * D;
  store_rows<D>(dk, dkb, rs, k0 + warp * 16, S, lane, 1.f, 1.f);
  store_rows<D>(dv, dvb, rs, k0 + warp * 16, S, lane, 1.f, 1.f);
}

template <typename Kern>
static int set_smem(Kern kern, size_t bytes) {
  if (bytes > 227 * 1024) return MFV_ERR_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? MFV_OK : (int)e;
}

}  // namespace mfv

extern "C" int mfv_attn_fwd(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse,
                            int64_t NB, int64_t S, int64_t H, int64_t D, float scale, void* stream) {
  using namespace mfv;
  if (NB <= 0 || S <= 0 || H <= 0 || (D != 64 && D != 32)) return MFV_ERR_SHAPE;
  if (D == 64 && (o_is_f16 != 0) == (qkv_is_f16 != 0) && !legacy_attention()) {
    // The flash-style kernel (64-key blocks, S double-buffered in TMEM, one TMEM read per score) is also the faster
    // one at S = 197 (23.5 vs 28.7 us in fp16 mode); MFVIT_ATTN_FLASH=0 selects the single-shot kernel for S <= 256.
    static int flash_all = -1, p2 = -1;
    if (flash_all < 0) {
      const char* e = getenv("MFVIT_ATTN_FLASH");
      flash_all = (e && e[0] == '0') ? 0 : 1;
      const char* e2 = getenv("MFVIT_ATTN_P2");
      p2 = (e2 && e2[0] == '1') ? 1 : 0;
    }
    // MFVIT_ATTN_P2=1: the persistent single-shot kernel for S <= 256 (both query tiles of a head in flight on one SM,
    // K / V read once per head).  Correct (tests/gpu_opcheck.py attn) but not faster: 22.1 us (bf16) / 29.3 us (fp16 +
    // bf16 copy) against 22.4 / 23.4 us for the flash-style kernel at 64 images x 6 heads - its two passes read every
    // score out of TMEM twice, and TMEM reads (64 B/clk/SM) cost as much as the exponentials (profiles/r02_summary.md).
    if (S <= 256 && p2)
      return attn_fwd_tc_p2(qkv, qkv_is_f16, o, o_is_f16, o_bf16_copy, lse, NB, S, H, scale,
                            reinterpret_cast<cudaStream_t>(stream));
    if (S <= 256 && !flash_all)
      return attn_fwd_tc(qkv, qkv_is_f16, o, o_is_f16, o_bf16_copy, lse, NB, S, H, scale,
                         reinterpret_cast<cudaStream_t>(stream));
    return attn_fwd_tc_mb(qkv, qkv_is_f16, o, o_is_f16, o_bf16_copy, lse, NB, S, H, scale,
                          reinterpret_cast<cudaStream_t>(stream));
  }
  if (NB * H > 65535) return MFV_ERR_SHAPE;
  const int Spad = ((int)S + 63) & ~63;
  // whole query range in one CTA when K, V and Q of a head fit twice per SM (S=197: 96 KB); else 128-row chunks
  const int Sq = ((int)S + ATT_ROWS - 1) / ATT_ROWS * ATT_ROWS;
  int q_rows = ((size_t)(2 * Spad + Sq) * D * 2 <= 110 * 1024) ? Sq : ATT_ROWS;
  const size_t smem = (size_t)(2 * Spad + q_rows) * D * 2;
  dim3 grid((unsigned)((S + q_rows - 1) / q_rows), (unsigned)(NB * H));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(o);
  __nv_bfloat16* obf = reinterpret_cast<__nv_bfloat16*>(o_bf16_copy);
  int rc;
#define MFV_ATT_FWD(DD, FF)                                                                                          \
  do {                                                                                                               \
    if ((rc = set_smem(attn_fwd_kernel<DD, FF>, smem))) return rc;                                                   \
    attn_fwd_kernel<DD, FF><<<grid, ATT_WARPS * 32, smem, st>>>(q, op, o_is_f16, obf, lse, (int)S, (int)H,           \
                                                                 scale * LOG2E, q_rows);                              \
  } while (0)
  if (D == 64) { if (qkv_is_f16) MFV_ATT_FWD(64, true); else MFV_ATT_FWD(64, false); }
  else { if (qkv_is_f16) MFV_ATT_FWD(32, true); else MFV_ATT_FWD(32, false); }
#undef MFV_ATT_FWD
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" size_t mfv_attn_bwd_workspace_bytes(int64_t NB, int64_t S, int64_t H, int64_t D) {
  return (D == 64 && S > 224) ? (size_t)(NB * S * H * D) * sizeof(float) : 0;
}

extern "C" int mfv_attn_bwd_ws(const void* qkv, int qkv_is_f16, const void* o, const void* d_o, const float* lse,
                               float* delta, void* dqkv, float* workspace, int64_t NB, int64_t S, int64_t H, int64_t D,
                               float scale, void* stream) {
  using namespace mfv;
  if (NB <= 0 || S <= 0 || H <= 0 || (D != 64 && D != 32)) return MFV_ERR_SHAPE;
  if (workspace && D == 64 && S > 224 && !legacy_attention())
    return attn_bwd_tc_mb(qkv, qkv_is_f16, o, d_o, lse, delta, workspace, dqkv, NB, S, H, scale,
                          reinterpret_cast<cudaStream_t>(stream));
  return mfv_attn_bwd(qkv, qkv_is_f16, o, d_o, lse, delta, dqkv, NB, S, H, D, scale, stream);
}

extern "C" int mfv_attn_bwd(const void* qkv, int qkv_is_f16, const void* o, const void* d_o, const float* lse,
                            float* delta, void* dqkv, int64_t NB, int64_t S, int64_t H, int64_t D, float scale,
                            void* stream) {
  using namespace mfv;
  if (NB <= 0 || S <= 0 || H <= 0 || (D != 64 && D != 32)) return MFV_ERR_SHAPE;
  if (D == 64 && S <= 224 && !legacy_attention())
    return attn_bwd_tc(qkv, qkv_is_f16, o, d_o, lse, dqkv, NB, S, H, scale, reinterpret_cast<cudaStream_t>(stream));
  if (NB * H > 65535) return MFV_ERR_SHAPE;
  const int Spad = ((int)S + 63) & ~63;
  const size_t smem_dq = (size_t)(2 * Spad + 2 * ATT_ROWS) * D * 2;
  const size_t smem_dkv = smem_dq + (size_t)Spad * 8;
  dim3 grid((unsigned)((S + ATT_ROWS - 1) / ATT_ROWS), (unsigned)(NB * H));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* op = reinterpret_cast<const __nv_bfloat16*>(o);
  const __nv_bfloat16* dop = reinterpret_cast<const __nv_bfloat16*>(d_o);
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
  int rc;
  if (D == 64) {
    if ((rc = set_smem(attn_bwd_dq_kernel<64>, smem_dq))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<64>, smem_dkv))) return rc;
    attn_bwd_dq_kernel<64><<<grid, ATT_WARPS * 32, smem_dq, st>>>(q, op, dop, lse, delta, dq, (int)S, (int)H, scale, qkv_is_f16);
    MFV_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<64><<<grid, ATT_WARPS * 32, smem_dkv, st>>>(q, dop, lse, delta, dq, (int)S, (int)H, scale, qkv_is_f16);
  } else {
    if ((rc = set_smem(attn_bwd_dq_kernel<32>, smem_dq))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<32>, smem_dkv))) return rc;
    attn_bwd_dq_kernel<32><<<grid, ATT_WARPS * 32, smem_dq, st>>>(q, op, dop, lse, delta, dq, (int)S, (int)H, scale, qkv_is_f16);
    MFV_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<32><<<grid, ATT_WARPS * 32, smem_dkv, st>>>(q, dop, lse, delta, dq, (int)S, (int)H, scale, qkv_is_f16);
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}
