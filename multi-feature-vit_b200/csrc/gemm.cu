// tcgen05/TMEM bf16 GEMM for every dense contraction of the ViT-S/16 block (SURVEY K3/K5/K6 and their dgrad/wgrad):
//
//     C[g][m][n] = epilogue( sum_k A[g][m][k] * B[g][n][k] )          g < G groups (the two MF-ViT branches)
//
// * warp-specialised, persistent: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer (+TMEM owner),
//   warps 2..5 = epilogue (tcgen05.ld -> registers -> global).  Accumulators are double-buffered in TMEM so the
//   epilogue of tile i overlaps the MMAs of tile i+1 (K is only 6 k-blocks for K=384).
// * operands are staged by TMA into 128B-swizzled shared memory through 3-D tensor maps (inner, rows, group); rows
//   or reduction elements past the tensor end are zero-filled by TMA, so no padding of M=B*197 tokens is needed.
// * each operand may be K-major (reduction dim contiguous: activations / nn.Linear weights in forward) or MN-major
//   (reduction dim strided: W in dgrad, dY and X in wgrad) - selected by the UMMA descriptors, no transposed copies.
// * split-K with fp32 red.global.add epilogue for the weight gradients (reduction over all tokens).
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 4;
constexpr int GEMM_THREADS = 32 * (2 + NUM_EPI_WARPS);

struct GemmParams {
  int M, N, K, G;
  int tiles_m, tiles_n, splits, kb_total, kb_per_split;
  int a_mn, b_mn;  // 1 = MN-major operand
  int a_f16, b_f16, out_f16;  // 1 = IEEE fp16 instead of bf16 (operands may be mixed)
  int epi;
  long long ldc, c_gstride;        // elements
  long long aux_ld, aux_gstride;   // residual (fp32) or pre-activation u (bf16)
  long long bias_gstride;
  void* C;
  void* C2;
  void* C3;
  const float* bias;
  const void* aux;
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 for manual alignment
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + S::STAGES * S::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tfull_bar = empty_bar + S::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], NUM_EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_group = p.tiles_m * p.tiles_n * p.splits;
  const int total_tiles = tiles_per_group * p.G;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int r = t;
        const int n_tile = r % p.tiles_n; r /= p.tiles_n;
        const int split = r % p.splits;   r /= p.splits;
        const int m_tile = r % p.tiles_m;
        const int g = r / p.tiles_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_3d(sa, &tmA, &full_bar[stage], kb * BK, m_tile * BM, g);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_3d(sa + j * 8192, &tmA, &full_bar[stage], m_tile * BM + j * 64, kb * BK, g);
          }
          if (!p.b_mn) {
            tma_load_3d(sb, &tmB, &full_bar[stage], kb * BK, n_tile * BN, g);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(sb + j * 8192, &tmB, &full_bar[stage], n_tile * BN + j * 64, kb * BK, g);
          }
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc2(p.a_f16 ? 0u : 1u, p.b_f16 ? 0u : 1u, BM, BN, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
      const uint32_t a_kadv = p.a_mn ? (UMMA_K * 128u) : (UMMA_K * 2u);
      const uint32_t b_kadv = p.b_mn ? (UMMA_K * 128u) : (UMMA_K * 2u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int r = t / p.tiles_n;
        const int split = r % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * a_kadv, a_lbo, 1024u);
            const uint64_t db = make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      int r = t;
      const int n_tile = r % p.tiles_n; r /= p.tiles_n;
      r /= p.splits;
      const int m_tile = r % p.tiles_m;
      const int g = r / p.tiles_m;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int m = m_tile * BM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const long long crow = (long long)g * p.c_gstride + (long long)m * p.ldc;
      const long long arow = (long long)g * p.aux_gstride + (long long)m * p.aux_ld;
      const float* bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_tile * BN + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + i));
            f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
          }
        }
        if (!row_ok) continue;
        switch (p.epi) {
          case MFV_EPI_BF16: {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + crow + n0);
            if (p.out_f16) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4(pack_f16(f[8 * i], f[8 * i + 1]), pack_f16(f[8 * i + 2], f[8 * i + 3]),
                                    pack_f16(f[8 * i + 4], f[8 * i + 5]), pack_f16(f[8 * i + 6], f[8 * i + 7]));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4(pack_bf16(f[8 * i], f[8 * i + 1]), pack_bf16(f[8 * i + 2], f[8 * i + 3]),
                                    pack_bf16(f[8 * i + 4], f[8 * i + 5]), pack_bf16(f[8 * i + 6], f[8 * i + 7]));
            }
          } break;
          case MFV_EPI_GELU: {  // C = u (pre-activation, saved for backward), C2 = gelu(u)
            uint4* du = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + crow + n0);
            uint4* dg = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C2) + crow + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              du[i] = make_uint4(pack_bf16(f[8 * i], f[8 * i + 1]), pack_bf16(f[8 * i + 2], f[8 * i + 3]),
                                 pack_bf16(f[8 * i + 4], f[8 * i + 5]), pack_bf16(f[8 * i + 6], f[8 * i + 7]));
              float gl[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) gl[j] = gelu_erf(f[8 * i + j]);
              if (p.C3)
                reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C3) + crow + n0)[i] =
                    make_uint4(pack_bf16(gl[0], gl[1]), pack_bf16(gl[2], gl[3]), pack_bf16(gl[4], gl[5]),
                               pack_bf16(gl[6], gl[7]));
              dg[i] = p.out_f16 ? make_uint4(pack_f16(gl[0], gl[1]), pack_f16(gl[2], gl[3]), pack_f16(gl[4], gl[5]),
                                             pack_f16(gl[6], gl[7]))
                                : make_uint4(pack_bf16(gl[0], gl[1]), pack_bf16(gl[2], gl[3]), pack_bf16(gl[4], gl[5]),
                                             pack_bf16(gl[6], gl[7]));
            }
          } break;
          case MFV_EPI_RESID_F32: {  // C(fp32) = acc + bias + aux(fp32 residual stream)
            const float4* res = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + arow + n0);
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + crow + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 rr = res[i];
              dst[i] = make_float4(f[4 * i] + rr.x, f[4 * i + 1] + rr.y, f[4 * i + 2] + rr.z, f[4 * i + 3] + rr.w);
            }
          } break;
          case MFV_EPI_DGELU: {  // C(bf16) = acc * gelu'(u), u = aux (bf16)
            const uint4* up = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + arow + n0);
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + crow + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 uu = up[i];
              const uint32_t uw[4] = {uu.x, uu.y, uu.z, uu.w};
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 u2 = unpack_bf16(uw[j]);
                o[j] = pack_bf16(f[8 * i + 2 * j] * gelu_erf_grad(u2.x), f[8 * i + 2 * j + 1] * gelu_erf_grad(u2.y));
              }
              dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
            }
          } break;
          case MFV_EPI_F32: {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + crow + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } break;
          case MFV_EPI_ATOMIC_F32: {  // split-K weight gradients: accumulate into the fp32 grad buffer
            float* dst = reinterpret_cast<float*>(p.C) + crow + n0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "f"(f[4 * i]),
                           "f"(f[4 * i + 1]), "f"(f[4 * i + 2]), "f"(f[4 * i + 3])
                           : "memory");
          } break;
          default: break;
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
static int encode_operand_map(CUtensorMap* map, const void* base, int mn_major, long long rows_mn, long long k,
                              long long ld, long long gstride, int groups, int box_mn, int is_f16) {
  // K-major : dims (k, rows_mn, G), strides (1, ld, gstride), box (64, box_mn, 1)
  // MN-major: dims (rows_mn, k, G), strides (1, ld, gstride), box (64, 64, 1)
  cuuint64_t dims[3];
  cuuint64_t strides[2];
  cuuint32_t box[3];
  cuuint32_t estr[3] = {1, 1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t)k; dims[1] = (cuuint64_t)rows_mn; dims[2] = (cuuint64_t)groups;
    box[0] = BK; box[1] = (cuuint32_t)box_mn; box[2] = 1;
  } else {
    dims[0] = (cuuint64_t)rows_mn; dims[1] = (cuuint64_t)k; dims[2] = (cuuint64_t)groups;
    box[0] = 64; box[1] = BK; box[2] = 1;
  }
  strides[0] = (cuuint64_t)ld * 2;
  strides[1] = (cuuint64_t)(groups > 1 ? gstride : (long long)dims[1] * ld) * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return MFV_ERR_ALIGN;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MFV_ERR_INIT;
  CUresult r = enc(map, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFV_OK : MFV_ERR_ARG;
}

template <int BN>
static int launch_gemm(const mfv_gemm_args* a, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    MFV_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  GemmParams p;
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K; p.G = (int)a->G;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  p.kb_total = (p.K + BK - 1) / BK;
  int splits = a->splits > 0 ? a->splits : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits > 1 && a->epilogue != MFV_EPI_ATOMIC_F32) return MFV_ERR_ARG;
  p.a_mn = a->a_mn_major; p.b_mn = a->b_mn_major; p.epi = a->epilogue;
  p.a_f16 = a->dtype_flags & 1; p.b_f16 = (a->dtype_flags >> 1) & 1; p.out_f16 = (a->dtype_flags >> 2) & 1;
  p.ldc = a->ldc; p.c_gstride = a->c_gstride;
  p.aux_ld = a->aux_ld; p.aux_gstride = a->aux_gstride; p.bias_gstride = a->bias_gstride;
  p.C = a->C; p.C2 = a->C2; p.C3 = a->C3; p.bias = (const float*)a->bias; p.aux = a->aux;

  CUtensorMap tmA, tmB;
  int rc = encode_operand_map(&tmA, a->A, a->a_mn_major, a->M, a->K, a->lda, a->a_gstride, p.G, BM, a->dtype_flags & 1);
  if (rc) return rc;
  rc = encode_operand_map(&tmB, a->B, a->b_mn_major, a->N, a->K, a->ldb, a->b_gstride, p.G, BN, (a->dtype_flags >> 1) & 1);
  if (rc) return rc;

  const int total = p.tiles_m * p.tiles_n * p.splits * p.G;
  int grid = total < num_sms() ? total : num_sms();
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tmA, tmB, p);
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

}  // namespace mfv

extern "C" int mfv_gemm(const mfv_gemm_args* a, void* stream) {
  using namespace mfv;
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0 || a->G <= 0) return MFV_ERR_SHAPE;
  if (a->N % 32 != 0) return MFV_ERR_SHAPE;
  if (a->ldc % 8 != 0) return MFV_ERR_ALIGN;
  if ((a->epilogue == MFV_EPI_RESID_F32 || a->epilogue == MFV_EPI_DGELU) && !a->aux) return MFV_ERR_ARG;
  if (a->epilogue == MFV_EPI_GELU && !a->C2) return MFV_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bn = a->block_n;
  if (bn == 0) {
    // pick the widest tile that still gives at least ~2 waves of CTAs; N=384 prefers 128 (3 exact tiles)
    const long long tm = (a->M + BM - 1) / BM;
    const long long sp = a->splits > 0 ? a->splits : 1;
    bn = 128;
    if (a->N % 256 == 0 && tm * (a->N / 256) * a->G * sp >= 2LL * num_sms()) bn = 256;
    if (tm * ((a->N + 127) / 128) * a->G * sp < num_sms() && a->N % 64 == 0) bn = 64;
  }
  switch (bn) {
    case 64: return launch_gemm<64>(a, s);
    case 128: return launch_gemm<128>(a, s);
    case 256: return launch_gemm<256>(a, s);
    default: return MFV_ERR_ARG;
  }
}
