// MoCo v2-style InfoNCE: L2-normalise q,k; l_pos = q.k; l_neg = q @ queue; logits = cat/T; fused
// log-sum-exp for the cross-entropy with label 0; backward to q; transposed enqueue of the gathered keys.
// The queue ([D][K] fp32, 64 MiB at K=65536, D=256) is streamed exactly once per pass, fully coalesced along K; the
// reference's queue.clone() (BLD:185) and torch.cat copy are gone.  Arithmetic is plain fp32 FMA (bit-comparable to
// the fp32 reference up to summation order).
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int NCE_TM = 128;  // samples per CTA tile (whole per-GPU batch at N <= 128)
constexpr int NCE_TN = 64;   // keys per CTA tile
constexpr int NCE_TK = 32;   // reduction chunk
constexpr int NCE_THREADS = 256;

// F.normalize(x, dim=1) for q and k, plus l_pos/T into logits[:,0].  One warp per row.
__global__ void nce_normalize_kernel(const float* __restrict__ q_raw, const float* __restrict__ k_raw,
                                     float* __restrict__ qn, float* __restrict__ kn, float* __restrict__ logits, int N,
                                     int D, long long ld_logits, float invT, __half* __restrict__ qs16 = nullptr) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  float sq = 0.f, sk = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float a = q_raw[(size_t)r * D + c], b = k_raw[(size_t)r * D + c];
    sq += a * a;
    sk += b * b;
  }
  sq = warp_sum(sq);
  sk = warp_sum(sk);
  const float dq = fmaxf(sqrtf(sq), 1e-12f), dk = fmaxf(sqrtf(sk), 1e-12f);
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float a = q_raw[(size_t)r * D + c] / dq, b = k_raw[(size_t)r * D + c] / dk;
    qn[(size_t)r * D + c] = a;
    kn[(size_t)r * D + c] = b;
    if (qs16) qs16[(size_t)r * D + c] = __float2half_rn(a * invT);  // tensor-core path: A operand = qn / T
    dot += a * b;
  }
  dot = warp_sum(dot);
  if (lane == 0) logits[(size_t)r * ld_logits] = dot * invT;
}

// logits[n][1 + j] = (qn[n] . queue[:, j]) / T for a 128 x 64 tile; per-row partial (max, sum exp) of the tile.
// grid = (K/64, ceil(N/128)).  Register tile 8 x 4 per thread.
__global__ void __launch_bounds__(NCE_THREADS)
nce_logits_kernel(const float* __restrict__ qn, const float* __restrict__ queue, float* __restrict__ logits,
                  float* __restrict__ part_max, float* __restrict__ part_sum, int N, int D, int K, float invT) {
  __shared__ float sA[NCE_TK][NCE_TM + 4];  // qn^T chunk
  __shared__ float sB[NCE_TK][NCE_TN];      // queue chunk
  __shared__ float sRed[NCE_TM][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  const int j0 = blockIdx.x * NCE_TN;
  const int n0 = blockIdx.y * NCE_TM;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int c0 = 0; c0 < D; c0 += NCE_TK) {
    // A: 128 rows x 32 cols -> transposed into sA[c][n]
    for (int i = threadIdx.x; i < NCE_TM * NCE_TK; i += NCE_THREADS) {
      const int n = i / NCE_TK, c = i % NCE_TK;
      sA[c][n] = (n0 + n < N && c0 + c < D) ? qn[(size_t)(n0 + n) * D + c0 + c] : 0.f;
    }
    for (int i = threadIdx.x; i < NCE_TK * NCE_TN / 4; i += NCE_THREADS) {
      const int c = i / (NCE_TN / 4), j4 = i % (NCE_TN / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + c < D) v = __ldg(reinterpret_cast<const float4*>(queue + (size_t)(c0 + c) * K + j0) + j4);
      *reinterpret_cast<float4*>(&sB[c][j4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < NCE_TK; ++c) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[c][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[c][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[c][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] += av[i] * b.x; acc[i][1] += av[i] * b.y; acc[i][2] += av[i] * b.z; acc[i][3] += av[i] * b.w;
      }
    }
    __syncthreads();
  }
  // scale, store, per-row tile statistics
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + ty * 8 + i;
    float4 v = make_float4(acc[i][0] * invT, acc[i][1] * invT, acc[i][2] * invT, acc[i][3] * invT);
    if (n < N) {
      float* dst = logits + (size_t)n * (K + 1) + 1 + j0 + tx * 4;  // +1: column 0 is l_pos (row is 4B-aligned only)
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    sRed[ty * 8 + i][tx] = mx;
  }
  __syncthreads();
  float rowmax = -INFINITY;
  if (threadIdx.x < NCE_TM) {
#pragma unroll
    for (int t = 0; t < 16; ++t) rowmax = fmaxf(rowmax, sRed[threadIdx.x][t]);
  }
  __syncthreads();
  if (threadIdx.x < NCE_TM) sRed[threadIdx.x][16] = rowmax;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float m = sRed[ty * 8 + i][16];
    sRed[ty * 8 + i][tx] = __expf(acc[i][0] * invT - m) + __expf(acc[i][1] * invT - m) + __expf(acc[i][2] * invT - m) +
                           __expf(acc[i][3] * invT - m);
  }
  __syncthreads();
  if (threadIdx.x < NCE_TM && n0 + threadIdx.x < N) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += sRed[threadIdx.x][t];
    const size_t idx = (size_t)(n0 + threadIdx.x) * gridDim.x + blockIdx.x;
    part_max[idx] = sRed[threadIdx.x][16];
    part_sum[idx] = s;
  }
}

// lse[n] = logsumexp over [l_pos/T, all tiles]; loss = mean(lse - logits[:,0]).  Single block.
__global__ void __launch_bounds__(256)
nce_finalize_kernel(const float* __restrict__ logits, const float* __restrict__ part_max,
                    const float* __restrict__ part_sum, float* __restrict__ lse, float* __restrict__ loss, int N,
                    int tiles, long long ld_logits) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float local = 0.f;
  for (int n = warp; n < N; n += 8) {
    const float lp = logits[(size_t)n * ld_logits];
    float m = lp;
    for (int t = lane; t < tiles; t += 32) m = fmaxf(m, part_max[(size_t)n * tiles + t]);
    m = warp_max(m);
    float s = 0.f;
    for (int t = lane; t < tiles; t += 32) s += part_sum[(size_t)n * tiles + t] * __expf(part_max[(size_t)n * tiles + t] - m);
    s = warp_sum(s) + __expf(lp - m);
    const float l = m + logf(s);
    if (lane == 0) { lse[n] = l; local += l - lp; }
  }
  if (lane == 0) red[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    loss[0] = s / (float)N;
  }
}

// dqn[n][c] += sum_j g[n][j] * queue[c][j] over this CTA's key range, g = d(logits[:,1+j]) / T.
// grid = (ctas, ceil(N/128)); each CTA walks key tiles ctas apart; 8 x 16 register tile per thread (128 x 256 out).
__global__ void __launch_bounds__(NCE_THREADS)
nce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ lse, const float* __restrict__ dlogits_ext,
               const float* __restrict__ queue, const float* __restrict__ ov, int ov_start, int ov_n,
               float* __restrict__ dqn_accum, int N, int D, int K, float invT, float gcoef) {
  constexpr int TJ = 16;
  __shared__ float sP[TJ][NCE_TM + 4];  // g^T chunk: [j][n]
  __shared__ float sQ[TJ][256 + 4];     // queue^T chunk: [j][c]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n0 = blockIdx.y * NCE_TM;
  float acc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[i][k] = 0.f;
  for (int j0 = blockIdx.x * TJ; j0 < K; j0 += gridDim.x * TJ) {
    for (int i = threadIdx.x; i < NCE_TM * TJ; i += NCE_THREADS) {
      const int n = i / TJ, j = i % TJ;
      float g = 0.f;
      if (n0 + n < N) {
        const size_t li = (size_t)(n0 + n) * (K + 1) + 1 + j0 + j;
        g = dlogits_ext ? dlogits_ext[li] * invT : __expf(logits[li] - lse[n0 + n]) * gcoef * invT;
      }
      sP[j][n] = g;
    }
    for (int i = threadIdx.x; i < D * TJ; i += NCE_THREADS) {
      const int c = i / TJ, j = i % TJ;
      const int jj = j0 + j - ov_start;  // columns overwritten by the enqueue since the forward: use the saved copy
      sQ[j][c] = (ov && jj >= 0 && jj < ov_n) ? ov[(size_t)c * ov_n + jj] : queue[(size_t)c * K + j0 + j];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < TJ; ++j) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sP[j][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sP[j][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 b = *reinterpret_cast<const float4*>(&sQ[j][tx * 4 + 64 * k]);
        bv[4 * k] = b.x; bv[4 * k + 1] = b.y; bv[4 * k + 2] = b.z; bv[4 * k + 3] = b.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[i][k] += av[i] * bv[k];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + ty * 8 + i;
    if (n >= N) continue;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int c = tx * 4 + 64 * (k >> 2) + (k & 3);
      if (c < D) atomicAdd(dqn_accum + (size_t)n * D + c, acc[i][k]);
    }
  }
}

// dq_raw = (dqn_total - qn * (qn . dqn_total)) / max(||q||, eps), dqn_total = dqn_accum + g_pos * kn.  One warp per row.
__global__ void nce_bwd_finish_kernel(const float* __restrict__ q_raw, const float* __restrict__ qn,
                                      const float* __restrict__ kn, const float* __restrict__ logits,
                                      const float* __restrict__ lse, const float* __restrict__ dlogits_ext,
                                      float* __restrict__ dq /* in: dqn_accum, out: dq_raw */, int N, int D,
                                      long long ld_logits, float invT, float gcoef,
                                      const float* __restrict__ gs_dev = nullptr /* dqn_accum is scaled by *gs_dev */) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  const size_t l0 = (size_t)r * ld_logits;
  const float descale = gs_dev ? 1.0f / *gs_dev : 1.0f;
  const float gpos = (dlogits_ext ? dlogits_ext[l0] : (__expf(logits[l0] - lse[r]) - 1.f) * gcoef) * invT;
  float nrm = 0.f, dot = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float a = q_raw[(size_t)r * D + c];
    nrm += a * a;
    const float g = dq[(size_t)r * D + c] * descale + gpos * kn[(size_t)r * D + c];
    dot += g * qn[(size_t)r * D + c];
  }
  nrm = fmaxf(sqrtf(warp_sum(nrm)), 1e-12f);
  dot = warp_sum(dot);
  for (int c = lane; c < D; c += 32) {
    const float g = dq[(size_t)r * D + c] * descale + gpos * kn[(size_t)r * D + c];
    dq[(size_t)r * D + c] = (g - qn[(size_t)r * D + c] * dot) / nrm;
  }
}

// queue[c][ptr + i] = keys[i][c]   (32 x 32 smem transpose, coalesced on both sides)
__global__ void enqueue_kernel(const float* __restrict__ keys, float* __restrict__ queue, int n, int D, int K,
                               int ptr) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && c < D) ? keys[(size_t)i * D + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, i = i0 + threadIdx.x;
    if (c < D && i < n) queue[(size_t)c * K + ptr + i] = tile[threadIdx.x][r];
  }
}


// ------------------------------------------------------------------------------------------------ tensor-core path
// l_neg = (qn / T) @ queue runs on the tcgen05 GEMM (gemm.cu) with fp16 operands - qn/T and an fp16 shadow of the queue
// (entries in [-1, 1]: 2^-12 absolute rounding, logits within ~5e-4 of the fp32 result at T = 0.2) - and fp32
// accumulation, writing straight into the logits buffer.  What remains here is streaming: the per-row log-sum-exp
// partials, the fp16 gradient operand of the backward GEMM, the queue shadow maintenance.

// part_max / part_sum of 1024-column chunks of the l_neg block.  grid = (K / 1024, N), 256 threads x float4.
__global__ void __launch_bounds__(256)
nce_lse_partials_kernel(const float* __restrict__ lneg, long long ld, float* __restrict__ part_max,
                        float* __restrict__ part_sum) {
  __shared__ float red[8];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const float4 v = reinterpret_cast<const float4*>(lneg + (size_t)n * ld + (size_t)chunk * 1024)[threadIdx.x];
  float m = warp_max(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = warp_sum(__expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m));
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    part_max[(size_t)n * gridDim.x + chunk] = m;
    part_sum[(size_t)n * gridDim.x + chunk] = t;
  }
}

// dl16[n][j] = fp16(gs * g[n][j]), g = d(loss)/d(l_neg[n][j]) (fused CE: softmax * gcoef / T, or an external gradient / T);
// columns [ov_start, ov_start + ov_n) were overwritten by the enqueue since the forward and are zeroed here (their
// exact fp32 contribution is added by nce_override_kernel).  gs keeps the tiny per-key probabilities (~1e-7) inside the
// fp16 normal range; *gs_dev is written by nce_scale_kernel.  grid = (K / 1024, N).
__global__ void __launch_bounds__(256)
nce_dl16_kernel(const float* __restrict__ lneg, const float* __restrict__ dext, long long ld,
                const float* __restrict__ lse, const float* __restrict__ gs_dev, __half* __restrict__ dl16, int K,
                int ov_start, int ov_n, float invT, float gcoef) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * 1024 + threadIdx.x * 4;
  const float gs = *gs_dev;
  float4 g;
  if (dext) {
    const float4 e = *reinterpret_cast<const float4*>(dext + (size_t)n * ld + j);
    g = make_float4(e.x * invT, e.y * invT, e.z * invT, e.w * invT);
  } else {
    const float4 v = *reinterpret_cast<const float4*>(lneg + (size_t)n * ld + j);
    const float l = lse[n], c = gcoef * invT;
    g = make_float4(__expf(v.x - l) * c, __expf(v.y - l) * c, __expf(v.z - l) * c, __expf(v.w - l) * c);
  }
  float gv[4] = {g.x * gs, g.y * gs, g.z * gs, g.w * gs};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (j + k >= ov_start && j + k < ov_start + ov_n) gv[k] = 0.f;
  reinterpret_cast<uint2*>(dl16 + (size_t)n * K + j)[0] = make_uint2(pack_f16(gv[0], gv[1]), pack_f16(gv[2], gv[3]));
}

// *gs_dev = target / max(|g|): host-known bound (fused CE: probabilities <= 1) or a device max-reduction of dext
__global__ void nce_scale_kernel(float* gs_dev, float bound, const float* __restrict__ absmax_dev, float invT) {
  float b = bound;
  if (absmax_dev) b = fmaxf(*absmax_dev * invT, 1e-30f);
  *gs_dev = exp2f(floorf(log2f(16384.0f / b)));
}
__global__ void __launch_bounds__(256)
nce_absmax_kernel(const float* __restrict__ dext, long long ld, int K, float* __restrict__ out) {
  const int n = blockIdx.y;
  float m = 0.f;
  for (int j = blockIdx.x * 1024 + threadIdx.x * 4; j < K; j += gridDim.x * 1024) {
    const float4 e = *reinterpret_cast<const float4*>(dext + (size_t)n * ld + j);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(e.x), fabsf(e.y)), fmaxf(fabsf(e.z), fabsf(e.w))));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));  // m >= 0: int order == float order
}

// dq_accum[n][c] += gs * sum_jj g[n][ov_start + jj] * ov[c][jj]  (the pre-enqueue contents of the overwritten columns,
// exact fp32).  grid = N CTAs of 256 threads (thread = c).
__global__ void __launch_bounds__(256)
nce_override_kernel(const float* __restrict__ lneg, const float* __restrict__ dext, long long ld,
                    const float* __restrict__ lse, const float* __restrict__ gs_dev, const float* __restrict__ ov,
                    int ov_start, int ov_n, float* __restrict__ dq_accum, int D, float invT, float gcoef) {
  __shared__ float sG[32];
  __shared__ float sO[256][33];
  const int n = blockIdx.x, c = threadIdx.x;
  const float gs = *gs_dev;
  float acc = 0.f;
  for (int j0 = 0; j0 < ov_n; j0 += 32) {
    if (threadIdx.x < 32) {
      const int jj = j0 + threadIdx.x;
      float g = 0.f;
      if (jj < ov_n) {
        const size_t li = (size_t)n * ld + ov_start + jj;
        g = dext ? dext[li] * invT : __expf(lneg[li] - lse[n]) * gcoef * invT;
      }
      sG[threadIdx.x] = g * gs;
    }
    for (int i = threadIdx.x; i < D * 32; i += 256) {
      const int cc = i >> 5, jj = j0 + (i & 31);
      sO[cc][i & 31] = (jj < ov_n) ? ov[(size_t)cc * ov_n + jj] : 0.f;
    }
    __syncthreads();
    if (c < D) {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc += sG[k] * sO[c][k];
    }
    __syncthreads();
  }
  if (c < D) dq_accum[(size_t)n * D + c] += acc;
}

// queue16[c][j] = fp16(queue[c][j]) for j in [col0, col0 + ncols); grid-stride over D x ncols
__global__ void nce_queue16_kernel(const float* __restrict__ queue, __half* __restrict__ queue16, int D, int K, int col0,
                                   int ncols) {
  const long long total = (long long)D * ncols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / ncols), j = col0 + (int)(i % ncols);
    queue16[(size_t)c * K + j] = __float2half_rn(queue[(size_t)c * K + j]);
  }
}

}  // namespace mfv

using namespace mfv;

// scratch for the per-tile softmax statistics lives at the tail of `logits`' sibling buffer: the caller passes `lse`
// with room for N * (1 + 2 * K/64) floats (lse first, then part_max, part_sum).
extern "C" int mfv_infonce_fwd(const float* q_raw, const float* k_raw, const float* queue, float* qn, float* kn,
                               float* logits, float* lse, float* loss, int64_t N, int64_t D, int64_t K, float T,
                               void* stream) {
  if (N <= 0 || D <= 0 || K <= 0 || K % NCE_TN || D % 4 || D > 256) return MFV_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float invT = 1.0f / T;
  const int tiles = (int)(K / NCE_TN);
  float* part_max = lse + N;
  float* part_sum = part_max + (size_t)N * tiles;
  nce_normalize_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(q_raw, k_raw, qn, kn, logits, (int)N, (int)D, K + 1, invT);
  MFV_LAUNCH_CHECK();
  nce_logits_kernel<<<dim3((unsigned)tiles, (unsigned)((N + NCE_TM - 1) / NCE_TM)), NCE_THREADS, 0, st>>>(
      qn, queue, logits, part_max, part_sum, (int)N, (int)D, (int)K, invT);
  MFV_LAUNCH_CHECK();
  nce_finalize_kernel<<<1, 256, 0, st>>>(logits, part_max, part_sum, lse, loss, (int)N, tiles, K + 1);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_infonce_bwd(const float* q_raw, const float* qn, const float* kn, const float* queue,
                               const float* logits, const float* lse, const float* dlogits_ext,
                               const float* queue_override, int64_t ov_start, int64_t ov_n, float gscale,
                               float* dq_raw, int64_t N, int64_t D, int64_t K, float T, void* stream) {
  if (N <= 0 || D != 256 || K <= 0 || K % 32) return MFV_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float invT = 1.0f / T;
  const float gcoef = gscale / (float)N;  // mean reduction of the CE
  int rc = mfv_fill_f32(dq_raw, 0.f, N * D, stream);
  if (rc) return rc;
  const unsigned ctas = (unsigned)num_sms();
  nce_bwd_kernel<<<dim3(ctas, (unsigned)((N + NCE_TM - 1) / NCE_TM)), NCE_THREADS, 0, st>>>(
      logits, lse, dlogits_ext, queue, queue_override, (int)ov_start, (int)ov_n, dq_raw, (int)N, (int)D, (int)K, invT,
      gcoef);
  MFV_LAUNCH_CHECK();
  nce_bwd_finish_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(q_raw, qn, kn, logits, lse, dlogits_ext, dq_raw,
                                                                 (int)N, (int)D, K + 1, invT, gcoef);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_enqueue_keys(const float* keys, float* queue, int64_t n, int64_t D, int64_t K, int64_t ptr,
                                void* stream) {
  if (n <= 0 || D <= 0 || ptr < 0 || ptr + n > K) return MFV_ERR_SHAPE;
  enqueue_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((D + 31) / 32)), dim3(32, 8), 0,
                   reinterpret_cast<cudaStream_t>(stream)>>>(keys, queue, (int)n, (int)D, (int)K, (int)ptr);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// ---- tensor-core InfoNCE.  Buffer conventions (all caller-owned):
//   queue16  fp16 [D][K]  shadow of `queue`, maintained with mfv_queue16_update
//   logits   f32  [N][ld], ld = K + 8: column 7 = l_pos / T, columns 8.. = l_neg / T (16-byte aligned block, written by
//            the GEMM's TMA stores); the tensor handed to the loss is the strided view logits[:, 7:]
//   qs16     fp16 [N][D]  scratch (qn / T)
//   lse      f32  [N * (1 + 2 * K / 1024)]: lse first, then the per-chunk partials
extern "C" int mfv_infonce_tc_fwd(const float* q_raw, const float* k_raw, const void* queue16, float* qn, float* kn,
                                  void* qs16, float* logits, int64_t ld_logits, float* lse, float* loss, int64_t N,
                                  int64_t D, int64_t K, float T, void* stream) {
  if (N <= 0 || D <= 0 || K <= 0 || K % 1024 || D % 64 || ld_logits < K + 8 || ld_logits % 4) return MFV_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float invT = 1.0f / T;
  const int chunks = (int)(K / 1024);
  float* part_max = lse + N;
  float* part_sum = part_max + (size_t)N * chunks;
  nce_normalize_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(q_raw, k_raw, qn, kn, logits + 7, (int)N, (int)D, ld_logits,
                                                                 invT, reinterpret_cast<__half*>(qs16));
  MFV_LAUNCH_CHECK();
  mfv_gemm_args a = {};
  a.A = qs16; a.B = queue16; a.C = logits + 8;
  a.M = N; a.N = K; a.K = D; a.G = 1;
  a.lda = D; a.ldb = K; a.ldc = ld_logits;
  a.b_mn_major = 1;  // queue is [D][K]: the key index is contiguous
  a.epilogue = MFV_EPI_F32;
  a.dtype_flags = 3;  // fp16 operands
  a.block_n = 256; a.cta_group = 1;
  int rc = mfv_gemm(&a, stream);
  if (rc) return rc;
  nce_lse_partials_kernel<<<dim3((unsigned)chunks, (unsigned)N), 256, 0, st>>>(logits + 8, ld_logits, part_max, part_sum);
  MFV_LAUNCH_CHECK();
  nce_finalize_kernel<<<1, 256, 0, st>>>(logits + 7, part_max, part_sum, lse, loss, (int)N, chunks, ld_logits);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// dl16: fp16 [N][K] scratch; scal: f32 [2] scratch.  dlogits_ext (optional) has the layout of `logits`.
extern "C" int mfv_infonce_tc_bwd(const float* q_raw, const float* qn, const float* kn, const void* queue16,
                                  const float* logits, int64_t ld_logits, const float* lse, const float* dlogits_ext,
                                  const float* queue_override, int64_t ov_start, int64_t ov_n, float gscale, void* dl16,
                                  float* scal, float* dq_raw, int64_t N, int64_t D, int64_t K, float T, void* stream) {
  if (N <= 0 || D != 256 || K <= 0 || K % 1024 || ld_logits < K + 8) return MFV_ERR_SHAPE;
  if (!queue_override) ov_n = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float invT = 1.0f / T;
  const float gcoef = gscale / (float)N;  // mean reduction of the CE
  int rc = mfv_fill_f32(dq_raw, 0.f, N * D, stream);
  if (rc) return rc;
  const float* dneg = dlogits_ext ? dlogits_ext + 8 : nullptr;
  if (dlogits_ext) {
    rc = mfv_fill_f32(scal, 0.f, 4, stream);  // scal[1] = running |max|
    if (rc) return rc;
    nce_absmax_kernel<<<dim3(16, (unsigned)N), 256, 0, st>>>(dneg, ld_logits, (int)K, scal + 1);
    MFV_LAUNCH_CHECK();
  }
  nce_scale_kernel<<<1, 1, 0, st>>>(scal, fmaxf(fabsf(gcoef) * invT, 1e-30f), dlogits_ext ? scal + 1 : nullptr, invT);
  MFV_LAUNCH_CHECK();
  nce_dl16_kernel<<<dim3((unsigned)(K / 1024), (unsigned)N), 256, 0, st>>>(
      logits + 8, dneg, ld_logits, lse, scal, reinterpret_cast<__half*>(dl16), (int)K, (int)ov_start, (int)ov_n, invT, gcoef);
  MFV_LAUNCH_CHECK();
  mfv_gemm_args a = {};
  a.A = dl16; a.B = queue16; a.C = dq_raw;
  a.M = N; a.N = D; a.K = K; a.G = 1;
  a.lda = K; a.ldb = K; a.ldc = D;
  a.epilogue = MFV_EPI_ATOMIC_F32;
  a.dtype_flags = 3;
  a.block_n = 256; a.cta_group = 1;
  a.splits = num_sms() - 20;  // one 128 x 256 tile per split: ~8 k-blocks of 64 keys each
  rc = mfv_gemm(&a, stream);
  if (rc) return rc;
  if (ov_n > 0) {
    nce_override_kernel<<<(unsigned)N, 256, 0, st>>>(logits + 8, dneg, ld_logits, lse, scal, queue_override,
                                                      (int)ov_start, (int)ov_n, dq_raw, (int)D, invT, gcoef);
    MFV_LAUNCH_CHECK();
  }
  nce_bwd_finish_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(q_raw, qn, kn, logits + 7, lse,
                                                                 dlogits_ext ? dlogits_ext + 7 : nullptr, dq_raw, (int)N,
                                                                 (int)D, ld_logits, invT, gcoef, scal);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_queue16_update(const float* queue, void* queue16, int64_t D, int64_t K, int64_t col0, int64_t ncols,
                                  void* stream) {
  if (D <= 0 || K <= 0 || col0 < 0 || ncols <= 0 || col0 + ncols > K) return MFV_ERR_SHAPE;
  const long long total = D * ncols;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 16u * (unsigned)num_sms()) grid = 16u * (unsigned)num_sms();
  nce_queue16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      queue, reinterpret_cast<__half*>(queue16), (int)D, (int)K, (int)col0, (int)ncols);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}
