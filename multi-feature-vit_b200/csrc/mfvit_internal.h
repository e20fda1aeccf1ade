// Internal glue shared by the .cu files (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/mfvit.h"

namespace mfv {
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();
int num_sms();
bool pdl_enabled();
bool fuse_ln_enabled();  // MFVIT_FUSE_LN=0: standalone LayerNorm launches instead of the fused proj / fc2 epilogue
bool dx32_stream_enabled();  // MFVIT_DX32=1: fp32 residual-gradient stream through the LayerNorm backward (default: bf16)
bool patch_tma_enabled();  // MFVIT_PATCH_TMA=0: patchify -> GEMM -> embed_finish instead of the im2col-free TMA kernel
bool gemm_multicast_enabled();  // MFVIT_GEMM_MC (default 0): B multicast between two pairs in the 384-wide GEMMs
int streamk_mask();           // MFVIT_STREAMK bit mask (runtime.cu)
int streamk_min_kb();         // MFVIT_STREAMK_MINKB: shortest reduction (64-wide k-blocks per tile) stream-K is used for
float* streamk_workspace();   // partial-accumulator slots of the stream-K GEMMs (nullptr before mfv_init)
unsigned* streamk_flags();
int rows96_mode();  // MFVIT_ROWS96: 96 rows per CTA in the 384-wide pair tiles: 1 = forward residual GEMMs, 2 = + bf16 dgrads
// Side stream of mfv_vit_backward: weight-gradient GEMMs and bias column sums are off the critical path (nothing in
// the backward consumes them), so they run beside the dgrad / attention / LayerNorm chain and fill the SMs those
// kernels leave idle (dgrad grids are 100 CTAs on 148 SMs at 32 pairs).  MFVIT_SIDE_STREAM=0 serialises everything.
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork[2];  // main -> side: inputs of the MLP-half / attention-half weight gradients are ready
  cudaEvent_t done[2];  // side -> main: those weight gradients have finished reading the reusable buffers
  cudaEvent_t fusion_fork, fusion_done;  // mfv_fusion_bwd_deferred / mfv_fusion_bwd_join (fusion.cu)
};
SideStream* side_stream();  // nullptr when disabled or creation failed
bool legacy_attention();  // MFVIT_ATTN=legacy forces the mma.sync attention kernels (A/B measurements)
// tcgen05 attention (attn_tc.cu); head_dim 64, S <= 256
int attn_fwd_tc(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse, long long NB,
                long long S, long long H, float scale, cudaStream_t st);
int attn_fwd_tc_p2(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse, long long NB,
                long long S, long long H, float scale, cudaStream_t st);
int attn_fwd_tc_mb(const void* qkv, int qkv_is_f16, void* o, int o_is_f16, void* o_bf16_copy, float* lse, long long NB,
                   long long S, long long H, float scale, cudaStream_t st);  // S > 256
int attn_bwd_tc(const void* qkv, int qkv_is_f16, const void* o, const void* d_o, const float* lse, void* dqkv,
                long long NB, long long S, long long H, float scale, cudaStream_t st);  // S <= 224
int attn_bwd_tc_mb(const void* qkv, int qkv_is_f16, const void* o, const void* d_o, const float* lse, float* delta,
                   float* dq_ws, void* dqkv, long long NB, long long S, long long H, float scale, cudaStream_t st);  // MFVIT_PDL=0 disables programmatic dependent launch (default on)

// Launch with programmatic dependent launch (PDL) enabled: the kernel may start while its predecessor in the stream is
// still draining, and MUST call griddep_wait() (common.cuh) before its first global memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Optional per-kernel-class device timing of the encoder executor (CUDA events on the launching stream).
enum ProfLabel {
  PROF_PATCHIFY = 0, PROF_GEMM_FWD, PROF_LN_FWD, PROF_ATTN_FWD, PROF_EMBED, PROF_GEMM_DGRAD, PROF_GEMM_WGRAD,
  PROF_COLSUM, PROF_LN_BWD, PROF_ATTN_BWD, PROF_EMBED_BWD, PROF_NUM_LABELS
};
bool prof_enabled();
void prof_begin(int label, cudaStream_t st);
void prof_end(int label, cudaStream_t st);
struct ProfScope {
  int label; cudaStream_t st; bool on;
  ProfScope(int l, cudaStream_t s) : label(l), st(s), on(prof_enabled()) { if (on) prof_begin(label, st); }
  ~ProfScope() { if (on) prof_end(label, st); }
};
}  // namespace mfv
