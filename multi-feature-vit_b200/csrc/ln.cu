// LayerNorm forward / backward over the fp32 residual stream.  HBM-bound: one warp per token row,
// 128-bit loads, the row lives in registers between the statistics pass and the normalise pass (single read of x).
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int LN_WARPS = 8;

template <int NV>  // NV float4 per lane; C = NV * 128
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y16, int y16_is_f16, __nv_bfloat16* __restrict__ y16b,
              float* __restrict__ y32, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              long long rows_per_group, long long total_rows, long long gb_gstride, float eps) {
  constexpr int C = NV * 128;
  griddep_wait();    // PDL launch: the producer of x may still be draining
  griddep_launch();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  const long long g = row / rows_per_group;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4* gm = reinterpret_cast<const float4*>(gamma + g * gb_gstride);
  const float4* bt = reinterpret_cast<const float4*>(beta + g * gb_gstride);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 gg = __ldg(gm + lane + 32 * i), bb = __ldg(bt + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * gg.x + bb.x;
    o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
    o.z = (v[i].z - mean) * rstd * gg.z + bb.z;
    o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
    if (y16) {
      uint2 pk = y16_is_f16 ? make_uint2(pack_f16(o.x, o.y), pack_f16(o.z, o.w))
                            : make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      reinterpret_cast<uint2*>(y16 + row * C)[lane + 32 * i] = pk;
    }
    if (y16b)  // bf16 copy kept for the backward GEMMs when the forward operand is fp16
      reinterpret_cast<uint2*>(y16b + row * C)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    if (y32) reinterpret_cast<float4*>(y32 + row * C)[lane + 32 * i] = o;
  }
}

// grid = (blocks_per_group, G).  Each warp strides over the rows of its group; per-lane partial dgamma/dbeta stay in
// registers and are combined through shared memory, then one red.global.add per column per block.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ dres,
              const __nv_bfloat16* __restrict__ dres16, const float* __restrict__ x, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ gamma, float* __restrict__ dx32, __nv_bfloat16* __restrict__ dx16,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum,
              long long rows_per_group, long long gb_gstride) {
  constexpr int C = NV * 128;
  __shared__ float red[LN_WARPS][C + 4];
  griddep_wait();    // PDL launch: the producers of dy / dres may still be draining
  griddep_launch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long g = blockIdx.y;
  const float4* gm = reinterpret_cast<const float4*>(gamma + g * gb_gstride);
  float4 gam[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gam[i] = __ldg(gm + lane + 32 * i);
  float4 dg[NV], db[NV], ds[NV];  // ds: column sums of dx = bias gradient of the Linear that produced x's residual branch
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (long long r = (long long)blockIdx.x * LN_WARPS + warp; r < rows_per_group; r += (long long)gridDim.x * LN_WARPS) {
    const long long row = g * rows_per_group + r;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[NV], d[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = reinterpret_cast<const float4*>(x + row * C)[lane + 32 * i];
      if (dy16) {
        const uint2 pk = reinterpret_cast<const uint2*>(dy16 + row * C)[lane + 32 * i];
        const float2 a = unpack_bf16(pk.x), b = unpack_bf16(pk.y);
        d[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d[i] = reinterpret_cast<const float4*>(dy32 + row * C)[lane + 32 * i];
      }
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      dg[i].x += d[i].x * xh[i].x; dg[i].y += d[i].y * xh[i].y; dg[i].z += d[i].z * xh[i].z; dg[i].w += d[i].w * xh[i].w;
      db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
      d[i].x *= gam[i].x; d[i].y *= gam[i].y; d[i].z *= gam[i].z; d[i].w *= gam[i].w;
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
    const float c1 = warp_sum(s1) * (1.0f / C), c2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (d[i].x - c1 - xh[i].x * c2);
      o.y = rstd * (d[i].y - c1 - xh[i].y * c2);
      o.z = rstd * (d[i].z - c1 - xh[i].z * c2);
      o.w = rstd * (d[i].w - c1 - xh[i].w * c2);
      if (dres) {
        const float4 rr = reinterpret_cast<const float4*>(dres + row * C)[lane + 32 * i];
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      } else if (dres16) {
        const uint2 pk = reinterpret_cast<const uint2*>(dres16 + row * C)[lane + 32 * i];
        const float2 a = unpack_bf16(pk.x), b = unpack_bf16(pk.y);
        o.x += a.x; o.y += a.y; o.z += b.x; o.w += b.y;
      }
      ds[i].x += o.x; ds[i].y += o.y; ds[i].z += o.z; ds[i].w += o.w;
      if (dx32) reinterpret_cast<float4*>(dx32 + row * C)[lane + 32 * i] = o;
      if (dx16)
        reinterpret_cast<uint2*>(dx16 + row * C)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
  if (!dgamma && !dbeta && !dxsum) return;
  // block reduction of the column partials: three rounds through smem (dgamma, dbeta, sum of dx)
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    float* out = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 t = pass == 0 ? dg[i] : (pass == 1 ? db[i] : ds[i]);
      *reinterpret_cast<float4*>(&red[warp][(lane + 32 * i) * 4]) = t;
    }
    __syncthreads();
    if (out) {
      for (int c = threadIdx.x; c < C; c += LN_WARPS * 32) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) acc += red[w][c];
        atomicAdd(out + g * gb_gstride + c, acc);
      }
    }
  }
}

// The backward's common case - dy in bf16, the residual-path gradient in bf16 (or absent) - with TWO rows per warp
// iteration and every load of both rows issued before the first reduction.  The generic kernel above is latency-bound
// (a warp owns ~5 rows of a 32-pair step, each a serial chain load -> two warp reductions -> store: 54 % of the HBM peak
// in-step with one row in flight); 16-bit inputs stay packed in registers until they are used.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
ln_bwd16_kernel(const __nv_bfloat16* __restrict__ dy16, const __nv_bfloat16* __restrict__ dres16,
                const float* __restrict__ x, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                const float* __restrict__ gamma, float* __restrict__ dx32, __nv_bfloat16* __restrict__ dx16,
                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum,
                long long rows_per_group, long long gb_gstride) {
  constexpr int C = NV * 128;
  constexpr int R = 2;  // rows in flight per warp
  __shared__ float red[LN_WARPS][C + 4];
  griddep_wait();    // PDL launch: the producers of dy / dres may still be draining
  griddep_launch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long g = blockIdx.y;
  const float4* gm = reinterpret_cast<const float4*>(gamma + g * gb_gstride);
  float4 dg[NV], db[NV], ds[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const long long stride = (long long)gridDim.x * LN_WARPS;
  for (long long r0 = (long long)blockIdx.x * LN_WARPS + warp; r0 < rows_per_group; r0 += R * stride) {
    float4 xv[R][NV];
    uint2 dq[R][NV], rq[R][NV];
    float mean[R], rstd[R];
    bool live[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long long r = r0 + k * stride;
      live[k] = r < rows_per_group;
      const long long row = g * rows_per_group + (live[k] ? r : r0);
      mean[k] = mean_in[row];
      rstd[k] = rstd_in[row];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xv[k][i] = reinterpret_cast<const float4*>(x + row * C)[lane + 32 * i];
        dq[k][i] = reinterpret_cast<const uint2*>(dy16 + row * C)[lane + 32 * i];
        rq[k][i] = dres16 ? reinterpret_cast<const uint2*>(dres16 + row * C)[lane + 32 * i] : make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (!live[k]) continue;  // warp-uniform
      const long long row = g * rows_per_group + r0 + k * stride;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4& v = xv[k][i];  // becomes x_hat in place
        v = make_float4((v.x - mean[k]) * rstd[k], (v.y - mean[k]) * rstd[k], (v.z - mean[k]) * rstd[k],
                        (v.w - mean[k]) * rstd[k]);
        const float2 a = unpack_bf16(dq[k][i].x), b = unpack_bf16(dq[k][i].y);
        const float4 gmi = __ldg(gm + lane + 32 * i);
        dg[i].x += a.x * v.x; dg[i].y += a.y * v.y; dg[i].z += b.x * v.z; dg[i].w += b.y * v.w;
        db[i].x += a.x; db[i].y += a.y; db[i].z += b.x; db[i].w += b.y;
        const float4 dd = make_float4(a.x * gmi.x, a.y * gmi.y, b.x * gmi.z, b.y * gmi.w);
        s1 += (dd.x + dd.y) + (dd.z + dd.w);
        s2 += (dd.x * v.x + dd.y * v.y) + (dd.z * v.z + dd.w * v.w);
      }
      const float c1 = warp_sum(s1) * (1.0f / C), c2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 v = xv[k][i];
        const float2 a = unpack_bf16(dq[k][i].x), b = unpack_bf16(dq[k][i].y);
        const float2 ra = unpack_bf16(rq[k][i].x), rb = unpack_bf16(rq[k][i].y);
        const float4 gmi = __ldg(gm + lane + 32 * i);
        float4 o;
        o.x = rstd[k] * (a.x * gmi.x - c1 - v.x * c2) + ra.x;
        o.y = rstd[k] * (a.y * gmi.y - c1 - v.y * c2) + ra.y;
        o.z = rstd[k] * (b.x * gmi.z - c1 - v.z * c2) + rb.x;
        o.w = rstd[k] * (b.y * gmi.w - c1 - v.w * c2) + rb.y;
        ds[i].x += o.x; ds[i].y += o.y; ds[i].z += o.z; ds[i].w += o.w;
        if (dx32) reinterpret_cast<float4*>(dx32 + row * C)[lane + 32 * i] = o;
        if (dx16)
          reinterpret_cast<uint2*>(dx16 + row * C)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
  if (!dgamma && !dbeta && !dxsum) return;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    float* out = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 t = pass == 0 ? dg[i] : (pass == 1 ? db[i] : ds[i]);
      *reinterpret_cast<float4*>(&red[warp][(lane + 32 * i) * 4]) = t;
    }
    __syncthreads();
    if (out) {
      for (int c = threadIdx.x; c < C; c += LN_WARPS * 32) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) acc += red[w][c];
        atomicAdd(out + g * gb_gstride + c, acc);
      }
    }
  }
}

}  // namespace mfv

extern "C" int mfv_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y16v, int y16_is_f16,
                                 void* y_bf16_copy, float* y_f32, float* mean, float* rstd, int64_t G, int64_t rows,
                                 int64_t C, int64_t gb_gstride, float eps, void* stream) {
  using namespace mfv;
  if (G <= 0 || rows <= 0) return MFV_ERR_SHAPE;
  const long long total = G * rows;
  const unsigned grid = (unsigned)((total + LN_WARPS - 1) / LN_WARPS);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* y16 = reinterpret_cast<__nv_bfloat16*>(y16v);
  __nv_bfloat16* y16b = reinterpret_cast<__nv_bfloat16*>(y_bf16_copy);
  switch (C) {
    case 256: MFV_CUDA_CHECK(launch_pdl(ln_fwd_kernel<2>, dim3(grid), dim3(LN_WARPS * 32), 0, s, x, gamma, beta, y16, y16_is_f16, y16b, y_f32, mean, rstd, rows, total, gb_gstride, eps)); break;
    case 384: MFV_CUDA_CHECK(launch_pdl(ln_fwd_kernel<3>, dim3(grid), dim3(LN_WARPS * 32), 0, s, x, gamma, beta, y16, y16_is_f16, y16b, y_f32, mean, rstd, rows, total, gb_gstride, eps)); break;
    case 768: MFV_CUDA_CHECK(launch_pdl(ln_fwd_kernel<6>, dim3(grid), dim3(LN_WARPS * 32), 0, s, x, gamma, beta, y16, y16_is_f16, y16b, y_f32, mean, rstd, rows, total, gb_gstride, eps)); break;
    default: return MFV_ERR_SHAPE;
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* dres, const void* dres_bf16,
                                 const float* x, const float* mean, const float* rstd, const float* gamma, float* dx_f32,
                                 void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int64_t G, int64_t rows,
                                 int64_t C, int64_t gb_gstride, void* stream) {
  using namespace mfv;
  if (G <= 0 || rows <= 0) return MFV_ERR_SHAPE;
  if (!dy_bf16 && !dy_f32) return MFV_ERR_ARG;
  if (dres && dres_bf16) return MFV_ERR_ARG;
  const __nv_bfloat16* dres16 = reinterpret_cast<const __nv_bfloat16*>(dres_bf16);
  long long bx = (rows + LN_WARPS - 1) / LN_WARPS;
  const long long cap = (2LL * num_sms() + G - 1) / G;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)G);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dy16 = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  __nv_bfloat16* dx16 = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  if (dy16 && !dres) {  // 16-bit inputs: two rows in flight per warp
    switch (C) {
      case 256: MFV_CUDA_CHECK(launch_pdl(ln_bwd16_kernel<2>, grid, dim3(LN_WARPS * 32), 0, s, dy16, dres16, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, dx_colsum, rows, gb_gstride)); break;
      case 384: MFV_CUDA_CHECK(launch_pdl(ln_bwd16_kernel<3>, grid, dim3(LN_WARPS * 32), 0, s, dy16, dres16, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, dx_colsum, rows, gb_gstride)); break;
      case 768: MFV_CUDA_CHECK(launch_pdl(ln_bwd16_kernel<6>, grid, dim3(LN_WARPS * 32), 0, s, dy16, dres16, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, dx_colsum, rows, gb_gstride)); break;
      default: return MFV_ERR_SHAPE;
    }
    MFV_LAUNCH_CHECK();
    return MFV_OK;
  }
  switch (C) {
    case 256: MFV_CUDA_CHECK(launch_pdl(ln_bwd_kernel<2>, grid, dim3(LN_WARPS * 32), 0, s, dy16, dy_f32, dres, dres16, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, dx_colsum, rows, gb_gstride)); break;
    case 384: MFV_CUDA_CHECK(launch_pdl(ln_bwd_kernel<3>, grid, dim3(LN_WARPS * 32), 0, s, dy16, dy_f32, dres, dres16, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, dx_colsum, rows, gb_gstride)); break;
    case 768: MFV_CUDA_CHECK(launch_pdl(ln_bwd_kernel<6>, grid, dim3(LN_WARPS * 32), 0, s, dy16, dy_f32, dres, dres16, x, mean, rstd, gamma, dx_f32, dx16, dgamma, dbeta, dx_colsum, rows, gb_gstride)); break;
    default: return MFV_ERR_SHAPE;
  }
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}
