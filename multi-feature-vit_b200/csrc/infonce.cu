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
                                     int D, long long ld_logits, float invT) {
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
                                      long long ld_logits, float invT, float gcoef) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  const size_t l0 = (size_t)r * ld_logits;
  const float gpos = (dlogits_ext ? dlogits_ext[l0] : (__expf(logits[l0] - lse[r]) - 1.f) * gcoef) * invT;
  float nrm = 0.f, dot = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float a = q_raw[(size_t)r * D + c];
    nrm += a * a;
    const float g = dq[(size_t)r * D + c] + gpos * kn[(size_t)r * D + c];
    dot += g * qn[(size_t)r * D + c];
  }
  nrm = fmaxf(sqrtf(warp_sum(nrm)), 1e-12f);
  dot = warp_sum(dot);
  for (int c = lane; c < D; c += 32) {
    const float g = dq[(size_t)r * D + c] + gpos * kn[(size_t)r * D + c];
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
