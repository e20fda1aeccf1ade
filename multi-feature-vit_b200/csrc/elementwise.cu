// HBM-bound vectorised kernels: bf16 shadow cast, fill, EMA (bit-exact), bias-gradient column sums, patchify /
// token assembly, small classification heads, small-class cross-entropy, fused SGD / Adam steps.
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

// ----------------------------------------------------------------------------------------------- cast / fill
// one read of the fp32 master, up to two 16-bit shadows (bf16 for backward GEMMs, fp16 for forward GEMMs)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     __half* __restrict__ dst_h, long long n) {
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    if (dst)
      reinterpret_cast<uint4*>(dst)[i] =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    if (dst_h)
      reinterpret_cast<uint4*>(dst_h)[i] =
          make_uint4(pack_f16(a.x, a.y), pack_f16(a.z, a.w), pack_f16(b.x, b.y), pack_f16(b.z, b.w));
  }
  for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (dst) dst[i] = __float2bfloat16_rn(src[i]);
    if (dst_h) dst_h[i] = __float2half_rn(src[i]);
  }
}

// bf16 -> f32 (gradients all-reduced in bf16 come back into the fp32 buffer the optimizer reads)
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[i];
    const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
  for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __bfloat162float(src[i]);
}

__global__ void fill_f32_kernel(float* __restrict__ dst, float v, long long n) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 v4 = make_float4(v, v, v, v);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    reinterpret_cast<float4*>(dst)[i] = v4;
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

// ----------------------------------------------------------------------------------------------- EMA (BLD:83-89)
// k = k*m + q*(1-m): __fmul_rn / __fadd_rn forbid FMA contraction so the result is bit-identical to the eager
// three-kernel sequence (mul, mul, add).  grid.y = chunk index.
__global__ void ema_kernel(const mfv_ema_chunk* __restrict__ chunks, float m, float omm) {
  const mfv_ema_chunk ch = chunks[blockIdx.y];
  float* __restrict__ k = ch.k;
  const float* __restrict__ q = ch.q;
  const long long n = ch.n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
  long long done = 0;
  if (aligned) {
    const long long n4 = n >> 2;
    // 2 independent 128-bit loads per tensor in flight per thread
    for (long long i = tid; i < n4; i += 2 * stride) {
      const long long j = i + stride;
      const float4 k0 = reinterpret_cast<const float4*>(k)[i];
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(q) + i);
      float4 k1 = k0, q1 = q0;
      if (j < n4) {
        k1 = reinterpret_cast<const float4*>(k)[j];
        q1 = __ldg(reinterpret_cast<const float4*>(q) + j);
      }
      float4 r;
      r.x = __fadd_rn(__fmul_rn(k0.x, m), __fmul_rn(q0.x, omm));
      r.y = __fadd_rn(__fmul_rn(k0.y, m), __fmul_rn(q0.y, omm));
      r.z = __fadd_rn(__fmul_rn(k0.z, m), __fmul_rn(q0.z, omm));
      r.w = __fadd_rn(__fmul_rn(k0.w, m), __fmul_rn(q0.w, omm));
      reinterpret_cast<float4*>(k)[i] = r;
      if (j < n4) {
        r.x = __fadd_rn(__fmul_rn(k1.x, m), __fmul_rn(q1.x, omm));
        r.y = __fadd_rn(__fmul_rn(k1.y, m), __fmul_rn(q1.y, omm));
        r.z = __fadd_rn(__fmul_rn(k1.z, m), __fmul_rn(q1.z, omm));
        r.w = __fadd_rn(__fmul_rn(k1.w, m), __fmul_rn(q1.w, omm));
        reinterpret_cast<float4*>(k)[j] = r;
      }
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += stride) k[i] = __fadd_rn(__fmul_rn(k[i], m), __fmul_rn(q[i], omm));
}

// ----------------------------------------------------------------------------------------------- column sums
// x bf16 [G][rows][C]; out[g][c] += sum_r x.  block = 8 warps; a warp covers 256 columns (8 per lane).
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int C,
                   long long out_gstride) {
  __shared__ float red[8][256 + 8];
  griddep_wait();    // PDL launch: the producer of x may still be draining
  griddep_launch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const long long g = blockIdx.z;
  const long long rows_per_chunk = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, rows);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < C) {
    const __nv_bfloat16* base = x + (g * rows) * C + col;
    auto add = [&](const uint4& v) {
      const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    };
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {  // four independent 16-byte loads in flight per lane (the loop is latency-bound)
      const uint4 v0 = *reinterpret_cast<const uint4*>(base + r * C);
      const uint4 v1 = *reinterpret_cast<const uint4*>(base + (r + 8) * C);
      const uint4 v2 = *reinterpret_cast<const uint4*>(base + (r + 16) * C);
      const uint4 v3 = *reinterpret_cast<const uint4*>(base + (r + 24) * C);
      add(v0); add(v1); add(v2); add(v3);
    }
    for (; r < r1; r += 8) add(*reinterpret_cast<const uint4*>(base + r * C));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(out + g * out_gstride + blockIdx.x * 256 + c, s);
  }
}

// ----------------------------------------------------------------------------------------------- patch embedding glue
// img f32 [GB][3][HW][HW] -> patches bf16 [GB*np][768], k = c*256 + i*16 + j.  One thread per 8 pixels of a row.
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ patches, int is_f16,
                                __nv_bfloat16* __restrict__ patches_bf, long long GB, int HW) {
  const int chunks = HW / 8;
  const int pw = HW / 16;
  const long long total = GB * 3LL * HW * chunks;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int x8 = (int)(r % chunks); r /= chunks;
    const int y = (int)(r % HW); r /= HW;
    const int c = (int)(r % 3);
    const long long gb = r / 3;
    const float4* src = reinterpret_cast<const float4*>(img + ((gb * 3 + c) * HW + y) * (long long)HW + x8 * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const long long p = gb * (long long)(pw * pw) + (y / 16) * pw + (x8 / 2);
    const int k = c * 256 + (y % 16) * 16 + (x8 % 2) * 8;
    const uint4 bfv = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    *reinterpret_cast<uint4*>(patches + p * 768 + k) =
        is_f16 ? make_uint4(pack_f16(a.x, a.y), pack_f16(a.z, a.w), pack_f16(b.x, b.y), pack_f16(b.z, b.w)) : bfv;
    if (patches_bf) *reinterpret_cast<uint4*>(patches_bf + p * 768 + k) = bfv;
  }
}

// x[g][b][0] = cls + pos[0]; x[g][b][1+p] = acc[g][b*np+p] + pos[1+p]   (bias already added by the GEMM epilogue)
__global__ void embed_finish_kernel(const float* __restrict__ acc, const float* __restrict__ cls,
                                    const float* __restrict__ pos, float* __restrict__ x, long long B, int np, int C,
                                    long long p_gstride) {
  const int c4n = C / 4;
  const long long g = blockIdx.y;
  const long long total = B * (np + 1) * c4n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % c4n);
    const long long row = t / c4n;
    const int tok = (int)(row % (np + 1));
    const long long b = row / (np + 1);
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + g * p_gstride + (long long)tok * C) + c4);
    float4 v;
    if (tok == 0)
      v = __ldg(reinterpret_cast<const float4*>(cls + g * p_gstride) + c4);
    else
      v = reinterpret_cast<const float4*>(acc + ((g * B + b) * np + (tok - 1)) * C)[c4];
    reinterpret_cast<float4*>(x + (g * B * (np + 1) + row) * C)[c4] =
        make_float4(v.x + pe.x, v.y + pe.y, v.z + pe.z, v.w + pe.w);
  }
}

// dacc[g][b*np+p] = bf16(dx[g][b][1+p]);  dcls[g] += sum_b dx[g][b][0]   (dbias comes from mfv_colsum_bf16(dacc))
__global__ void embed_finish_bwd_kernel(const float* __restrict__ dx, __nv_bfloat16* __restrict__ dacc,
                                        float* __restrict__ dcls, long long B, int np, int C, long long p_gstride) {
  const int c4n = C / 4;
  const long long g = blockIdx.y;
  const long long total = B * np * c4n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % c4n);
    const long long row = t / c4n;
    const int p = (int)(row % np);
    const long long b = row / np;
    const float4 v = reinterpret_cast<const float4*>(dx + ((g * B + b) * (np + 1) + 1 + p) * C)[c4];
    reinterpret_cast<uint2*>(dacc + ((g * B + b) * np + p) * C)[c4] =
        make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  if (dcls && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (long long b = 0; b < B; ++b) s += dx[((g * B + b) * (np + 1)) * C + c];
      dcls[g * p_gstride + c] += s;
    }
  }
}

// ----------------------------------------------------------------------------------------------- small heads
// y[r][n] = b[n] + x[r] . w[n]; one warp per row.
__global__ void linear_small_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                        const float* __restrict__ b, float* __restrict__ y, long long rows, int C,
                                        int N) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  for (int n = 0; n < N; ++n) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[r * ldx + c] * __ldg(w + (long long)n * C + c);
    s = warp_sum(s);
    if (lane == 0) y[r * N + n] = s + (b ? b[n] : 0.f);
  }
}
// single block: dx[r][c] = sum_n dy[r][n] w[n][c];  dw[n][c] += sum_r dy[r][n] x[r][c];  db[n] += sum_r dy[r][n]
__global__ void linear_small_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                        const float* __restrict__ dy, float* __restrict__ dx, long long lddx,
                                        float* __restrict__ dw, float* __restrict__ db, long long rows, int C, int N) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float wc[32], acc[32];
    for (int n = 0; n < N; ++n) { wc[n] = w[(long long)n * C + c]; acc[n] = 0.f; }
    for (long long r = 0; r < rows; ++r) {
      const float xv = x[r * ldx + c];
      float s = 0.f;
      for (int n = 0; n < N; ++n) {
        const float d = dy[r * N + n];
        s += d * wc[n];
        acc[n] += d * xv;
      }
      if (dx) dx[r * lddx + c] = s;
    }
    if (dw) for (int n = 0; n < N; ++n) dw[(long long)n * C + c] += acc[n];
  }
  if (db && blockIdx.x == 0 && threadIdx.x < N) {
    float s = 0.f;
    for (long long r = 0; r < rows; ++r) s += dy[r * N + threadIdx.x];
    db[threadIdx.x] += s;
  }
}

// ----------------------------------------------------------------------------------------------- CE (MAIN_CA:868-873)
__global__ void __launch_bounds__(256)
ce_small_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                const long long* __restrict__ target, float* __restrict__ loss, float* __restrict__ dlogits,
                long long rows, int NC) {
  __shared__ float red[8];
  float local = 0.f;
  const float inv_rows = 1.0f / (float)rows;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    float z[32];
    float mx = -INFINITY;
    for (int n = 0; n < NC; ++n) {
      float v = a[r * NC + n];
      if (b) v += b[r * NC + n];
      if (c) v += c[r * NC + n];
      z[n] = v;
      mx = fmaxf(mx, v);
    }
    float se = 0.f;
    for (int n = 0; n < NC; ++n) se += expf(z[n] - mx);
    const float lse = mx + logf(se);
    const long long tl = target[r];
    if (tl < 0 || tl >= NC) {  // torch's CE device-asserts here; never index z[] with it: the loss turns NaN instead
      local = __int_as_float(0x7fc00000);
      if (dlogits)
        for (int n = 0; n < NC; ++n) dlogits[r * NC + n] = __int_as_float(0x7fc00000);
      continue;
    }
    const int t = (int)tl;
    local += lse - z[t];
    if (dlogits)
      for (int n = 0; n < NC; ++n) dlogits[r * NC + n] = (expf(z[n] - lse) - (n == t ? 1.f : 0.f)) * inv_rows;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    loss[0] = s * inv_rows;
  }
}

// ----------------------------------------------------------------------------------------------- optimiser steps
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                           __nv_bfloat16* __restrict__ shadow, __half* __restrict__ shadow_h, long long n, float lr,
                           float mom, float wd, int first, const float* __restrict__ lr_dev) {
  if (lr_dev) lr = __ldg(lr_dev);  // learning rate kept on the device: schedules work under CUDA-graph replay
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w};
    float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float bb[4] = {0.f, 0.f, 0.f, 0.f};
    if (buf && !first) {
      const float4 bv = reinterpret_cast<float4*>(buf)[i];
      bb[0] = bv.x; bb[1] = bv.y; bb[2] = bv.z; bb[3] = bv.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d = gg[j] + wd * pp[j];
      if (buf) { bb[j] = first ? d : mom * bb[j] + d; d = bb[j]; }
      pp[j] -= lr * d;
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    if (buf) reinterpret_cast<float4*>(buf)[i] = make_float4(bb[0], bb[1], bb[2], bb[3]);
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16(pp[0], pp[1]), pack_bf16(pp[2], pp[3]));
    if (shadow_h) reinterpret_cast<uint2*>(shadow_h)[i] = make_uint2(pack_f16(pp[0], pp[1]), pack_f16(pp[2], pp[3]));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = g[i] + wd * p[i];
    if (buf) { const float b = first ? d : mom * buf[i] + d; buf[i] = b; d = b; }
    const float np_ = p[i] - lr * d;
    p[i] = np_;
    if (shadow) shadow[i] = __float2bfloat16_rn(np_);
    if (shadow_h) shadow_h[i] = __float2half_rn(np_);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m1,
                            float* __restrict__ m2, __nv_bfloat16* __restrict__ shadow, __half* __restrict__ shadow_h,
                            long long n, float lr, float b1, float b2, float eps, float wd, int decoupled, float bc1,
                            float bc2_sqrt, const float* __restrict__ lr_dev, const long long* __restrict__ step_dev) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (lr_dev) lr = __ldg(lr_dev);
  if (step_dev) {  // bias corrections from the device-resident step count (1-based), as torch computes them per step
    const float st = (float)__ldg(step_dev);
    bc1 = 1.0f - powf(b1, st);
    bc2_sqrt = sqrtf(1.0f - powf(b2, st));
  }
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pv = p[i];
    float gv = g[i];
    if (decoupled) pv *= (1.f - lr * wd); else gv += wd * pv;
    const float a = m1[i] + (1.f - b1) * (gv - m1[i]);          // exp_avg.lerp_(grad, 1-beta1)
    const float v = b2 * m2[i] + (1.f - b2) * gv * gv;
    m1[i] = a;
    m2[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    pv -= step_size * (a / denom);
    p[i] = pv;
    if (shadow) shadow[i] = __float2bfloat16_rn(pv);
    if (shadow_h) shadow_h[i] = __float2half_rn(pv);
  }
}

static inline unsigned grid_for(long long work_items, int threads, int max_blocks) {
  long long b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (unsigned)b;
}

}  // namespace mfv

using namespace mfv;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mfv_cast_shadow(const float* src, void* dst_bf16, void* dst_f16, int64_t n, void* stream) {
  if (n <= 0) return MFV_OK;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_bf16) | reinterpret_cast<uintptr_t>(dst_f16)) & 15)
    return MFV_ERR_ALIGN;
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), reinterpret_cast<__half*>(dst_f16), n);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_cast_bf16_f32(const void* src_bf16, float* dst, int64_t n, void* stream) {
  if (n <= 0) return MFV_OK;
  if ((reinterpret_cast<uintptr_t>(src_bf16) | reinterpret_cast<uintptr_t>(dst)) & 15) return MFV_ERR_ALIGN;
  cast_bf16_f32_kernel<<<grid_for(n / 8 + 1, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src_bf16), dst, n);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_fill_f32(float* dst, float value, int64_t n, void* stream) {
  if (n <= 0) return MFV_OK;
  if (reinterpret_cast<uintptr_t>(dst) & 15) return MFV_ERR_ALIGN;
  fill_f32_kernel<<<grid_for(n / 4 + 1, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(dst, value, n);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_ema_update(const mfv_ema_chunk* chunks_dev, int64_t n_chunks, int64_t max_chunk_elems, float m,
                              float one_minus_m, void* stream) {
  if (n_chunks <= 0) return MFV_OK;
  if (n_chunks > 65535) return MFV_ERR_SHAPE;
  // 8 floats per thread per iteration; enough blocks in x to cover the biggest chunk with ~8 CTAs per SM in total
  unsigned gx = grid_for(max_chunk_elems / 8 + 1, 256, 8 * num_sms());
  if (n_chunks > 1) {
    const unsigned cap = (unsigned)((16LL * num_sms() + n_chunks - 1) / n_chunks);
    if (gx > cap) gx = cap < 1 ? 1 : cap;
  }
  ema_kernel<<<dim3(gx, (unsigned)n_chunks), 256, 0, STREAM(stream)>>>(chunks_dev, m, one_minus_m);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_colsum_bf16(const void* x, float* out, int64_t G, int64_t rows, int64_t C, int64_t out_gstride,
                               void* stream) {
  if (G <= 0 || rows <= 0 || C <= 0 || C % 8) return MFV_ERR_SHAPE;
  const unsigned gx = (unsigned)((C + 255) / 256);
  long long gy = (4LL * num_sms()) / (gx * G);
  if (gy < 1) gy = 1;
  if (gy > (rows + 63) / 64) gy = (rows + 63) / 64;
  MFV_CUDA_CHECK(launch_pdl(colsum_bf16_kernel, dim3(gx, (unsigned)gy, (unsigned)G), dim3(256), 0, STREAM(stream),
                            reinterpret_cast<const __nv_bfloat16*>(x), out, (long long)rows, (int)C,
                            (long long)out_gstride));
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_patchify(const float* img, void* patches, int is_f16, void* patches_bf16_copy, int64_t GB,
                            int64_t HW, void* stream) {
  if (GB <= 0 || HW <= 0 || HW % 16) return MFV_ERR_SHAPE;
  const long long total = GB * 3 * HW * (HW / 8);
  patchify_kernel<<<grid_for(total, 256, 32 * num_sms()), 256, 0, STREAM(stream)>>>(
      img, reinterpret_cast<__nv_bfloat16*>(patches), is_f16, reinterpret_cast<__nv_bfloat16*>(patches_bf16_copy), GB,
      (int)HW);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_embed_finish(const float* acc, const float* bias, const float* cls, const float* pos, float* x,
                                int64_t G, int64_t B, int64_t np, int64_t C, int64_t p_gstride, void* stream) {
  (void)bias;  // bias is folded into the GEMM epilogue; kept in the signature for ABI stability
  if (G <= 0 || B <= 0 || C % 4) return MFV_ERR_SHAPE;
  const long long total = B * (np + 1) * (C / 4);
  embed_finish_kernel<<<dim3(grid_for(total, 256, 16 * num_sms()), (unsigned)G), 256, 0, STREAM(stream)>>>(
      acc, cls, pos, x, B, (int)np, (int)C, p_gstride);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_embed_finish_bwd(const float* dx, void* dacc_bf16, float* dbias, float* dcls, int64_t G, int64_t B,
                                    int64_t np, int64_t C, int64_t p_gstride, void* stream) {
  if (G <= 0 || B <= 0 || C % 4) return MFV_ERR_SHAPE;
  const long long total = B * np * (C / 4);
  embed_finish_bwd_kernel<<<dim3(grid_for(total, 256, 16 * num_sms()), (unsigned)G), 256, 0, STREAM(stream)>>>(
      dx, reinterpret_cast<__nv_bfloat16*>(dacc_bf16), dcls, B, (int)np, (int)C, p_gstride);
  MFV_LAUNCH_CHECK();
  if (dbias) return mfv_colsum_bf16(dacc_bf16, dbias, G, B * np, C, p_gstride, stream);
  return MFV_OK;
}

extern "C" int mfv_linear_small_fwd(const float* x, int64_t ldx, const float* w, const float* b, float* y,
                                    int64_t rows, int64_t C, int64_t N, void* stream) {
  if (rows <= 0 || N <= 0 || N > 32) return MFV_ERR_SHAPE;
  linear_small_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, STREAM(stream)>>>(x, ldx, w, b, y, rows, (int)C, (int)N);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_linear_small_bwd(const float* x, int64_t ldx, const float* w, const float* dy, float* dx,
                                    int64_t lddx, float* dw, float* db, int64_t rows, int64_t C, int64_t N,
                                    void* stream) {
  if (rows <= 0 || N <= 0 || N > 32) return MFV_ERR_SHAPE;
  linear_small_bwd_kernel<<<(unsigned)((C + 127) / 128), 128, 0, STREAM(stream)>>>(x, ldx, w, dy, dx, lddx, dw, db, rows,
                                                                                  (int)C, (int)N);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_ce_small(const float* a, const float* b, const float* c, const int64_t* target, float* loss,
                            float* dlogits, int64_t rows, int64_t NC, void* stream) {
  if (rows <= 0 || NC <= 0 || NC > 32) return MFV_ERR_SHAPE;
  ce_small_kernel<<<1, 256, 0, STREAM(stream)>>>(a, b, c, reinterpret_cast<const long long*>(target), loss, dlogits,
                                                 rows, (int)NC);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_sgd_step(float* p, const float* g, float* buf, void* shadow_bf16, void* shadow_f16, int64_t n,
                            float lr, float momentum, float weight_decay, int first_step, void* stream) {
  if (n <= 0) return MFV_OK;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(buf)) & 15)
    return MFV_ERR_ALIGN;
  sgd_kernel<<<grid_for(n / 4 + 1, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(
      p, g, buf, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), reinterpret_cast<__half*>(shadow_f16), n, lr, momentum,
      weight_decay, first_step, nullptr);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_sgd_step_dev(float* p, const float* g, float* buf, void* shadow_bf16, void* shadow_f16, int64_t n,
                                const float* lr_dev, float momentum, float weight_decay, int first_step, void* stream) {
  if (n <= 0) return MFV_OK;
  if (!lr_dev) return MFV_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(buf)) & 15)
    return MFV_ERR_ALIGN;
  sgd_kernel<<<grid_for(n / 4 + 1, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(
      p, g, buf, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), reinterpret_cast<__half*>(shadow_f16), n, 0.f, momentum,
      weight_decay, first_step, lr_dev);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                             void* shadow_f16, int64_t n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int decoupled_wd, int64_t step, void* stream) {
  if (n <= 0) return MFV_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(
      p, g, exp_avg, exp_avg_sq, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), reinterpret_cast<__half*>(shadow_f16), n,
      lr, beta1, beta2, eps, weight_decay, decoupled_wd, bc1, sqrtf(bc2), nullptr, nullptr);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_adam_step_dev(float* p, const float* g, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                                 void* shadow_f16, int64_t n, const float* lr_dev, float beta1, float beta2, float eps,
                                 float weight_decay, int decoupled_wd, const int64_t* step_dev, void* stream) {
  if (n <= 0) return MFV_OK;
  if (!lr_dev || !step_dev) return MFV_ERR_ARG;
  adam_kernel<<<grid_for(n, 256, 16 * num_sms()), 256, 0, STREAM(stream)>>>(
      p, g, exp_avg, exp_avg_sq, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), reinterpret_cast<__half*>(shadow_f16), n,
      0.f, beta1, beta2, eps, weight_decay, decoupled_wd, 1.f, 1.f, lr_dev, reinterpret_cast<const long long*>(step_dev));
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}
