// Patch embedding as an im2col-free GEMM (timm PatchEmbed = Conv2d(3, C, k=16, s=16) + flatten + transpose,
// commented restatement FUS:197-221; token assembly crossvit.py:130-146).
//
//     x[g][b][1 + p][n] = sum_k pixel(b, p, k) * W[g][n][k] + bias[g][n] + pos[g][1 + p][n]        x[g][b][0][:] = cls + pos[0]
//
// with p = (ph, pw) the patch, k = (c, i, j) the position inside it.  No patch matrix is ever written to HBM: the A operand
// is fetched by TMA straight out of the NCHW fp32 image through a 5-D tensor map whose dimensions are ordered
// (j, i, pw, ph, c*b) - strides need not be monotonic - so that a box of (16 j, 4 i, gw pw, PH ph) lands in shared
// memory as dense [patch][64 k] fp32 rows.  The 16 epilogue warps, idle during the mainloop, round that tile to the
// 16-bit operand format (fp16, or bf16 in the all-bf16 forward mode) into the 128B-swizzled K-major layout tcgen05
// reads; the weight tile comes by TMA from the 16-bit shadow.  (Feeding the fp32 tile to kind::tf32 directly also
// works - the 64B-swizzled variant of this map was measured - but tf32 TRUNCATES its operands: 2.6x the rounding error
// of fp16 on the one GEMM every token depends on, logits error 1.1-1.7e-3 -> 1.5-3.1e-3 against a 2e-3 bound.)
// One CTA per band of PH patch rows of an image (PH = 128 / gw rounded down: 9 of the 14 rows at 224^2, 5 of 24 at 384^2)
// and all 384 output channels (UMMA 128 x 256 + 128 x 128), so the epilogue finishes whole token rows: + bias + position
// embedding, written directly as the fp32 residual stream of block 0 - the former patchify -> GEMM -> embed_finish chain
// of three launches and two HBM round trips.
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int PE_THREADS = 32 * 18;       // warp 0: TMA producer, warp 1: MMA issuer, warps 2..17: converters + epilogue
constexpr int PE_RAW_BYTES = 128 * 256;   // [128 patches][4 i][16 j] fp32, dense
constexpr int PE_A_BYTES = 128 * 128;     // [128 patches][64 k] 16-bit, 128B-swizzled
constexpr int PE_B_BYTES = 384 * 128;     // [384 n][64 k] 16-bit, 128B-swizzled
constexpr int PE_RING = 2;
constexpr int PE_SMEM = PE_RING * (PE_RAW_BYTES + PE_A_BYTES + PE_B_BYTES) + 256 + 1024;
constexpr int PE_CWARPS = 16;

struct PatchParams {
  int B, G, gw, gh, PH, tiles_per_img, C, f16;
  long long S;                 // tokens per image (1 + gw * gh)
  long long P;                 // elements between the groups' parameter blocks
  const float* bias;           // [G][C]
  const float* cls;            // [G][C]
  const float* pos;            // [G][S][C]
  float* x;                    // [G][B][S][C]
};

// mbarrier.arrive that cannot be issued before `dep` has been produced (a register data dependency)
__device__ __forceinline__ void mbar_arrive_dep(uint64_t* bar, uint32_t dep) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];  // after %1" ::"r"(smem_u32(bar)), "r"(dep) : "memory");
}

__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmImg0, const __grid_constant__ CUtensorMap tmImg1,
                   const __grid_constant__ CUtensorMap tmW, const PatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_raw = smem;
  uint8_t* s_a = s_raw + PE_RING * PE_RAW_BYTES;
  uint8_t* s_b = s_a + PE_RING * PE_A_BYTES;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_b + PE_RING * PE_B_BYTES);
  uint64_t* raw_empty = raw_full + PE_RING;  // 16 converter-warp arrivals
  uint64_t* a_full = raw_empty + PE_RING;    // 16 converter-warp arrivals
  uint64_t* a_empty = a_full + PE_RING;      // MMA commit
  uint64_t* b_full = a_empty + PE_RING;
  uint64_t* b_empty = b_full + PE_RING;      // MMA commit
  uint64_t* tfull_bar = b_empty + PE_RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  int item = blockIdx.x;                                 // (group, image, band of PH patch rows)
  const int t = item % p.tiles_per_img; item /= p.tiles_per_img;
  const int b = item % p.B;
  const int g = item / p.B;
  const int ph0 = t * p.PH;
  const int rows = p.gw * min(p.PH, p.gh - ph0);
  constexpr int KB = 12;                                  // 768 / 64 k-blocks: (channel, four pixel rows of the patch)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(g ? &tmImg1 : &tmImg0);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < PE_RING; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], PE_CWARPS);
      mbar_init(&a_full[s], PE_CWARPS);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  if (warp == 0) {
    {  // whole warp in uniform control flow, one elected lane issues (see elect_one() in common.cuh)
      griddep_wait();
      const CUtensorMap* tmImg = g ? &tmImg1 : &tmImg0;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb & 1;
        const uint32_t par = (uint32_t)((kb >> 1) & 1);
        mbar_wait(&raw_empty[s], par ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&raw_full[s], (uint32_t)(p.gw * p.PH * 256));
          // pixels (j 0..15, rows 4*(kb%4) .. +3 of the patch, every patch column, PH patch rows, channel kb/4 of image b)
          tma_load_5d(s_raw + s * PE_RAW_BYTES, tmImg, &raw_full[s], 0, (kb & 3) * 4, 0, ph0, b * 3 + (kb >> 2));
        }
        __syncwarp();
        mbar_wait(&b_empty[s], par ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[s], (uint32_t)PE_B_BYTES);
#pragma unroll
          for (int j = 0; j < 6; ++j)
            tma_load_3d(s_b + s * PE_B_BYTES + j * 8192, &tmW, &b_full[s], kb * 64, j * 64, g);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t fmt = p.f16 ? 0u : 1u;
      const uint32_t idesc = make_idesc2(fmt, fmt, 128, 256, 0, 0);
      const uint32_t idesc2 = make_idesc2(fmt, fmt, 128, 128, 0, 0);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb & 1;
        const uint32_t par = (uint32_t)((kb >> 1) & 1);
        mbar_wait(&b_full[s], par);
        mbar_wait(&a_full[s], par);
        tc_fence_after();
        const uint32_t sa = smem_u32(s_a + s * PE_A_BYTES);
        const uint32_t sb = smem_u32(s_b + s * PE_B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * 32, 0u, 1024u);
            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 0u, 1024u);
            const uint64_t db2 = make_smem_desc_sw128(sb + 32768u + k * 32, 0u, 1024u);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_bf16(tmem_base, da, db, idesc, acc);
            umma_bf16(tmem_base + 256u, da, db2, idesc2, acc);
          }
          umma_commit(&a_empty[s]);
          umma_commit(&b_empty[s]);
          if (kb == KB - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ converter + epilogue warps
    griddep_wait();
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;
    const int C = p.C;
    const float* bias = p.bias + (long long)g * p.P;
    const float* pos = p.pos + (long long)g * p.P;
    float* xim = p.x + ((long long)g * p.B + b) * p.S * C;
    if (t == 0 && ew < 3) {  // class token row: cls + pos[0]
      const float* cls = p.cls + (long long)g * p.P;
      const int c4 = (ew * 32 + lane) * 4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(cls + c4)), e = __ldg(reinterpret_cast<const float4*>(pos + c4));
      *reinterpret_cast<float4*>(xim + c4) = make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w);
    }
    // ---- operand conversion: thread = (patch row pr, pixel row di of the k-block): 16 fp32 -> 16 x 16-bit
    const int ct = ew * 32 + lane;       // 0..511
    const int pr = ct >> 2, di = ct & 3;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb & 1;
      const uint32_t par = (uint32_t)((kb >> 1) & 1);
      mbar_wait(&raw_full[s], par);
      const float4* src = reinterpret_cast<const float4*>(s_raw + s * PE_RAW_BYTES + pr * 256 + di * 64);
      const float4 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
      uint4 o0, o1;
      if (p.f16) {
        o0 = make_uint4(pack_f16(v0.x, v0.y), pack_f16(v0.z, v0.w), pack_f16(v1.x, v1.y), pack_f16(v1.z, v1.w));
        o1 = make_uint4(pack_f16(v2.x, v2.y), pack_f16(v2.z, v2.w), pack_f16(v3.x, v3.y), pack_f16(v3.z, v3.w));
      } else {
        o0 = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
        o1 = make_uint4(pack_bf16(v2.x, v2.y), pack_bf16(v2.z, v2.w), pack_bf16(v3.x, v3.y), pack_bf16(v3.z, v3.w));
      }
      // Release the raw slot only once the loads have RETURNED: mbarrier.arrive is not ordered behind shared-memory loads
      // that were merely issued (it ran ahead of the LDS and let the next TMA box overwrite pixels still being read -
      // intermittent wrong operands).  The arrive therefore carries a register dependency on the converted values.
      __syncwarp();
      if (lane == 0) mbar_arrive_dep(&raw_empty[s], o0.x ^ o0.w ^ o1.x ^ o1.w);
      mbar_wait(&a_empty[s], par ^ 1);   // the MMAs that read this operand slot two k-blocks ago have retired
      uint8_t* row = s_a + s * PE_A_BYTES + pr * 128;
      *reinterpret_cast<uint4*>(row + (((2 * di) ^ (pr & 7)) << 4)) = o0;      // 128B swizzle: 16-byte chunk ^ (row & 7)
      *reinterpret_cast<uint4*>(row + (((2 * di + 1) ^ (pr & 7)) << 4)) = o1;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
    }
    // ---- epilogue
    const int lr = q * 32 + lane;                                // row of this band
    const bool valid = lr < rows;
    const long long tok = 1 + (long long)t * p.gw * p.PH + lr;  // token index of the patch
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if (q * 32 < rows) {  // warp-uniform: some row of this quarter is real
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int n0 = (h + 4 * i) * 16;
        uint32_t v[32];
        tmem_ld16(trow + (uint32_t)n0, v);
        tmem_ld_wait();
        if (valid) {
          const float* prow = pos + tok * C + n0;
          float* xr = xim + tok * C + n0;
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + k));
            const float4 e4 = __ldg(reinterpret_cast<const float4*>(prow + k));
            *reinterpret_cast<float4*>(xr + k) =
                make_float4(__uint_as_float(v[k]) + b4.x + e4.x, __uint_as_float(v[k + 1]) + b4.y + e4.y,
                            __uint_as_float(v[k + 2]) + b4.z + e4.z, __uint_as_float(v[k + 3]) + b4.w + e4.w);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int encode_image_map(CUtensorMap* map, const float* img, long long B, int HW, int PH, int probe_mode = 0) {
  (void)probe_mode;
  const int gw = HW / 16;
  cuuint64_t dims[5] = {16, 16, (cuuint64_t)gw, (cuuint64_t)gw, (cuuint64_t)(3 * B)};
  cuuint64_t strides[4] = {(cuuint64_t)HW * 4, 64, (cuuint64_t)16 * HW * 4, (cuuint64_t)HW * HW * 4};
  cuuint32_t box[5] = {16, 4, (cuuint32_t)gw, (cuuint32_t)PH, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (reinterpret_cast<uintptr_t>(img) & 15) return MFV_ERR_ALIGN;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MFV_ERR_INIT;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(img), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFV_OK : MFV_ERR_ARG;
}

// 16-bit weight shadow [G][C][768] (K-major): box of 64 k x 64 rows, 128B swizzle
static int encode_weight_map(CUtensorMap* map, const void* w16, long long gstride, int G, int C, int is_f16) {
  cuuint64_t dims[3] = {768, (cuuint64_t)C, (cuuint64_t)G};
  cuuint64_t strides[2] = {768 * 2, (cuuint64_t)(G > 1 ? gstride : (long long)C * 768) * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(w16) & 15) || (strides[1] & 15)) return MFV_ERR_ALIGN;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MFV_ERR_INIT;
  CUresult r = enc(map, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(w16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFV_OK : MFV_ERR_ARG;
}

// Test aid (tests/gpu_opcheck.py): one 5-D box of the image map copied raw out of shared memory, so the layout the
// converter warps assume - dense [patch][i][j] rows of 256 B - can be checked against the pixels.
__global__ void __launch_bounds__(128) patch_tma_probe_kernel(const __grid_constant__ CUtensorMap tmImg, float* out, int kb,
                                                              int ph0, int cb, int bytes, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  for (int i = threadIdx.x; i < 8192; i += 128) reinterpret_cast<float*>(smem)[i] = -777.0f;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)bytes);
    tma_load_5d(smem, &tmImg, bar, 0, (kb & 3) * 4, 0, ph0, cb);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 8192; i += 128) out[i] = reinterpret_cast<float*>(smem)[i];
}

}  // namespace mfv

extern "C" int mfv_debug_patch_tma_probe(const float* img, float* out8192, int64_t B, int64_t HW, int64_t kb, int64_t ph0,
                                         int64_t cb, void* stream) {
  using namespace mfv;
  const int gw = (int)(HW / 16), PH = 128 / gw;
  CUtensorMap tm;
  int rc = encode_image_map(&tm, img, B, (int)HW, PH);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    MFV_CUDA_CHECK(cudaFuncSetAttribute(patch_tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64 + 1024));
    attr = true;
  }
  patch_tma_probe_kernel<<<1, 128, 32768 + 64 + 1024, reinterpret_cast<cudaStream_t>(stream)>>>(tm, out8192, (int)kb, (int)ph0,
                                                                                              (int)cb, gw * PH * 256, 0);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

// img0 / img1: f32 [B][3][HW][HW] of group 0 / 1 (img1 ignored when G == 1); w16: 16-bit shadow of the conv weight
// [G][C][768] (fp16 when w_is_f16, else bf16), bias / cls f32 [G][C], pos f32 [G][S][C] (all inside the groups' parameter
// blocks, `p_gstride` elements apart); x f32 [G][B][S][C] out.
extern "C" int mfv_patch_embed_tma(const float* img0, const float* img1, const void* w16, int w_is_f16, const float* bias,
                                   const float* cls, const float* pos, float* x, int64_t G, int64_t B, int64_t HW, int64_t C,
                                   int64_t p_gstride, void* stream) {
  using namespace mfv;
  if (G < 1 || G > 2 || B <= 0 || HW <= 0 || HW % 16 || C != 384) return MFV_ERR_SHAPE;
  if (!img0 || (G == 2 && !img1) || !w16 || !bias || !cls || !pos || !x) return MFV_ERR_ARG;
  const int gw = (int)(HW / 16);
  if (gw > 128) return MFV_ERR_SHAPE;
  PatchParams p;
  p.B = (int)B; p.G = (int)G; p.gw = gw; p.gh = gw; p.PH = 128 / gw; p.C = (int)C; p.f16 = w_is_f16 ? 1 : 0;
  p.tiles_per_img = (gw + p.PH - 1) / p.PH;
  p.S = 1 + (long long)gw * gw;
  p.P = p_gstride;
  p.bias = bias; p.cls = cls; p.pos = pos; p.x = x;
  CUtensorMap tm0, tm1, tmW;
  int rc = encode_image_map(&tm0, img0, B, (int)HW, p.PH);
  if (rc) return rc;
  tm1 = tm0;
  if (G == 2) {
    rc = encode_image_map(&tm1, img1, B, (int)HW, p.PH);
    if (rc) return rc;
  }
  rc = encode_weight_map(&tmW, w16, p_gstride, (int)G, (int)C, p.f16);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    MFV_CUDA_CHECK(cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(G * B * p.tiles_per_img));
  cfg.blockDim = dim3(PE_THREADS);
  cfg.dynamicSmemBytes = PE_SMEM;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, patch_embed_kernel, tm0, tm1, tmW, p));
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}
