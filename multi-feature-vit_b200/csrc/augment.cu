// Device side of the paired input pipeline and the on-device epoch metrics (row 2).
//
// augment_u8_kernel: one pass from the uint8 image the host decoded and resized (H x W x 3, as cv2 / PIL hand it to the
// reference's transform, loader.py:122-129) to the normalised float32 NCHW batch the encoders consume.  Per sample:
// horizontal flip -> nearest-neighbour rotation about the image centre (Pillow's 16.16 fixed-point inverse map, so the
// sampled pixel is the one PIL picks) -> crop window -> x/255 -> (x - mean)/std, i.e. image_transform.py:50-84's
// RandomHorizontalFlip / RandomRotation / RandomCrop (or CenterCrop) / ToTensor / Normalize.  HBM-bound byte work:
// 3 B read + 12 B written per pixel, 4 pixels per thread, float4 stores; the two IEEE divisions per value are taken
// from a 3 x 256 table in shared memory, which keeps the result bit-identical to the eager float32 sequence.
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int AUG_PARAMS = 12;  // per sample: flip, rotate, a0..a5, top, left, 2 reserved (include/mfvit.h)

__global__ void __launch_bounds__(256)
augment_u8_kernel(const uint8_t* __restrict__ src, const int* __restrict__ params, const float* __restrict__ mean,
                  const float* __restrict__ stdv, float* __restrict__ out, int Hs, int Ws, int crop) {
  __shared__ float lut[3 * 256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8;
    // ToTensor: float(v).div(255); Normalize: sub_(mean).div_(std) - correctly rounded ops, no contraction
    lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), mean[c]), stdv[c]);
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int* pr = params + (long long)b * AUG_PARAMS;
  const int flip = pr[0], rot = pr[1];
  const int a0 = pr[2], a1 = pr[3], a2 = pr[4], a3 = pr[5], a4 = pr[6], a5 = pr[7];
  const int top = pr[8], left = pr[9];
  const int quads = crop >> 2;
  const long long plane = (long long)crop * crop;
  const uint8_t* img = src + (long long)b * Hs * Ws * 3;
  float* o = out + (long long)b * 3 * plane;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads * crop; q += gridDim.x * blockDim.x) {
    const int oy = q / quads, ox = (q - oy * quads) << 2;
    float v[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int x = ox + k + left, y = oy + top;
      if (rot) {
        const int xx = a2 + y * a1 + x * a0, yy = a5 + y * a4 + x * a3;  // pixel-centre offsets are folded into a2 / a5
        x = xx >> 16;
        y = yy >> 16;
      }
      // outside the source image: fill value 0 (rotation corners); also keeps a bad crop window in the device-side
      // parameters from reading out of bounds
      const bool ok = x >= 0 && x < Ws && y >= 0 && y < Hs;
      if (flip) x = Ws - 1 - x;
      const uint8_t* px = img + ((long long)y * Ws + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c][k] = lut[c * 256 + (ok ? (int)__ldg(px + c) : 0)];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(o + c * plane + (long long)oy * crop + ox) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

// MAIN_CA:868-899 without the per-iteration .item() / .cpu(): running loss, argmax hits and the raw summed logits /
// labels / predictions of the epoch stay on the device; the host reads them once per epoch.  One CTA, one row per thread.
__global__ void __launch_bounds__(256)
epoch_metrics_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                     const long long* __restrict__ target, const float* __restrict__ loss, int rows, int NC,
                     double* __restrict__ loss_sum, long long* __restrict__ counters, long long capacity,
                     float* __restrict__ vals, int* __restrict__ preds, int* __restrict__ gts) {
  __shared__ int hits[8];
  const long long base = counters[0];
  int local = 0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    float best = -INFINITY;
    int arg = 0;
    const long long slot = base + r;
    for (int n = 0; n < NC; ++n) {
      float v = a[r * NC + n];
      if (b) v += b[r * NC + n];
      if (c) v += c[r * NC + n];
      if (v > best) { best = v; arg = n; }  // torch.max: first maximal index
      if (slot < capacity) vals[slot * NC + n] = v;
    }
    const int t = (int)target[r];
    if (slot < capacity) { preds[slot] = arg; gts[slot] = t; }
    local += (arg == t);
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0) hits[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int h = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) h += hits[w];
    counters[0] = base + rows;
    counters[1] += h;
    const long long over = base + rows - capacity;
    if (over > 0) counters[2] += over < rows ? over : rows;
    loss_sum[0] += (double)loss[0] * (double)rows;  // running_loss += loss.item() * images.size(0)
  }
}

}  // namespace mfv

using namespace mfv;

extern "C" int mfv_augment_u8(const void* src_u8, const int32_t* params, const float* mean3, const float* std3,
                              float* out, int64_t B, int64_t Hs, int64_t Ws, int64_t crop, void* stream) {
  if (B <= 0) return MFV_OK;
  if (!src_u8 || !params || !mean3 || !std3 || !out) return MFV_ERR_ARG;
  if (crop <= 0 || crop % 4 || crop > Hs || crop > Ws || Hs >= 32768 || Ws >= 32768 || B > 65535) return MFV_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(out) & 15) return MFV_ERR_ALIGN;
  const long long quads = crop / 4 * crop;
  long long gx = (quads + 255) / 256;
  if (gx > 8LL * num_sms()) gx = 8LL * num_sms();
  augment_u8_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(src_u8), params, mean3, std3, out, (int)Hs, (int)Ws, (int)crop);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

extern "C" int mfv_epoch_metrics(const float* a, const float* b, const float* c, const int64_t* target,
                                 const float* loss, int64_t rows, int64_t NC, double* loss_sum, int64_t* counters,
                                 int64_t capacity, float* vals, int32_t* preds, int32_t* gts, void* stream) {
  if (rows <= 0) return MFV_OK;
  if (!a || !target || !loss || !loss_sum || !counters || !vals || !preds || !gts) return MFV_ERR_ARG;
  if (NC <= 0 || NC > 32 || rows > (1 << 24) || capacity < 0) return MFV_ERR_SHAPE;
  epoch_metrics_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      a, b, c, reinterpret_cast<const long long*>(target), loss, (int)rows, (int)NC, loss_sum,
      reinterpret_cast<long long*>(counters), (long long)capacity, vals, preds, gts);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}
