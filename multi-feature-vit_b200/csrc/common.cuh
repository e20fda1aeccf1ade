// Shared device helpers for the mfvit sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// vector load/store helpers and warp reductions.  Everything here is inline PTX for sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MFV_OK 0
#define MFV_ERR_SHAPE (-1)
#define MFV_ERR_ALIGN (-2)
#define MFV_ERR_ARCH (-3)
#define MFV_ERR_INIT (-4)
#define MFV_ERR_ARG (-5)

namespace mfv {
extern unsigned long long g_launch_count;
void note_error(const char* file, int line, const char* expr);  // remembered for mfv_last_error_where()
}
#define MFV_CUDA_CHECK(expr)                         \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) {                         \
      mfv::note_error(__FILE__, __LINE__, #expr);    \
      return (int)_e;                                \
    }                                                \
  } while (0)

// every kernel launch in the library is followed by this macro: error check + launch accounting (mfv_launch_count)
#define MFV_LAUNCH_CHECK()                           \
  do {                                               \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) {                         \
      mfv::note_error(__FILE__, __LINE__, "launch"); \
      return (int)_e;                                \
    }                                                \
    __atomic_add_fetch(&mfv::g_launch_count, 1ULL, __ATOMIC_RELAXED); /* the loader launches from its own thread */ \
  } while (0)

namespace mfv {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// Warp index as a value the compiler knows to be warp-uniform, and a one-lane election.  A role loop written as
// `if (lane == 0) { ... tcgen05.mma / TMA ... }` is divergent code to the compiler: every operand of UTCHMMA / UTMALDG /
// UTCBAR must be in a UNIFORM register, so each such instruction is wrapped in an ELECT + R2UR.BROADCAST + BRA.U.ANY
// "waterfall" loop - measured 180-310 clk per tcgen05.mma k-step in the GEMM mainloop (tests/gpu_ring_probe3.py), more
// than the 64-128 clk the tensor pipe needs for it.  With the whole warp running the loop (uniform control flow, uniform
// values) and only the asynchronous instruction itself under elect_one(), descriptors and coordinates live in uniform
// registers and the instruction issues directly.
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a while before it answers; an event loop over several
// barriers wants the answer now)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (visible error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mfvit: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ------------------------------------------------------------------ programmatic dependent launch
// Blocks until every kernel this grid depends on has completed and its memory is visible.  A no-op when the kernel was
// launched without cudaLaunchAttributeProgrammaticStreamSerialization.  EVERY global access of a kernel launched with that
// attribute must come after this call.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel in the stream (if launched with the PDL attribute) start its CTAs as soon as SM resources free
// up; it still blocks in griddep_wait() until this grid has completed.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane+i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: thread i of the warp writes TMEM lane (base_lane+i), 16 (8) consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// named barrier: `count` threads (a multiple of 32) of the CTA meet at barrier `id` (1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ------------------------------------------------------------------ clusters / cta_group::2 (CTA pairs)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// same without the release fence (.release.cluster costs a MEMBAR.ALL + ERRBAR per arrive): for hand-offs whose data
// is ordered by other means (TMEM reads: tcgen05.wait::ld + tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
// TMA load issued by either CTA of a pair; the transaction bytes are signalled on the LEADER CTA's mbarrier
// (peer bit 24 of the shared::cluster address cleared, as cute::SM100_TMA_2SM_LOAD does)
__device__ __forceinline__ void tma_load_3d_cg2(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// The same load delivered to the same smem offset of every CTA in cta_mask (cluster ranks); each destination CTA's share
// of the bytes is signalled on the barrier of ITS pair's leader (cute::SM100_TMA_2SM_LOAD_MULTICAST)
__device__ __forceinline__ void tma_load_3d_cg2_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                   int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], "
      "[%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the pair; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives on the mbarrier at this offset in every CTA selected by cta_mask
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout), SWIZZLE_128B, version 1.
//   bits [0,14)  start address >> 4     bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A/B format fmt (1 = bf16, 2 = tf32).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                  uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ------------------------------------------------------------------ misc math / memory helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_f16(uint32_t v) {
  __half2 t = *reinterpret_cast<__half2*>(&v);
  return __half22float2(t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
// Exact (erf) GELU, as torch.nn.GELU() / timm Mlp, evaluated as x * Phi(x) with the normal CDF written as a
// logistic of an odd polynomial:  Phi(x) = 1 / (1 + 2^(x * Q(x^2))),  Q = degree-4 minimax fit of -log2(e) * logit(Phi(x)) / x
// over |x| <= 7 (monotone beyond, so no clamping).  Max |x * dPhi| = 3.4e-6 and |dgelu'| = 6.9e-6 in fp32 arithmetic
// (tools/fit_gelu.py) - 70x below the fp16 rounding of the stored result - in 10 instructions: 6 FMA/MUL + 2 MUFU
// (ex2, rcp) instead of ~20 for an erf polynomial; the GELU epilogue is issue-bound, so this is what sets its speed.
__device__ __forceinline__ float normal_cdf_fast(float x, float x2) {
  float q = -3.2289765385939972e-06f;
  q = fmaf(q, x2, 8.82378953974694e-05f);
  q = fmaf(q, x2, 0.0003602751239668578f);
  q = fmaf(q, x2, -0.10522668808698654f);
  q = fmaf(q, x2, -2.3020453453063965f);
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(q * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));  // t = +inf (x << 0) -> 0
  return r;
}
__device__ __forceinline__ float gelu_erf(float x) { return x * normal_cdf_fast(x, x * x); }
// gelu(x) and gelu'(x) = Phi(x) + x * phi(x) from one CDF evaluation (fc2-dgrad epilogue: dhid and the recomputed GELU)
__device__ __forceinline__ void gelu_erf_both(float x, float& g, float& dg) {
  const float x2 = x * x;
  const float cdf = normal_cdf_fast(x, x2);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x2 * -0.72134752044448170368f));  // exp(-x^2/2)
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float g, dg;
  gelu_erf_both(x, g, dg);
  return dg;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): two elements per issue slot.  The GEMM epilogues
// are instruction-issue bound, so the GELU polynomials run on pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 splat2(float c) { return pack2(c, c); }
// Phi(x) for a pair (same polynomial as normal_cdf_fast): 7 packed instructions + 4 MUFU per pair
__device__ __forceinline__ f32x2 normal_cdf_fast2(f32x2 x, f32x2 x2) {
  f32x2 q = fma2(splat2(-3.2289765385939972e-06f), x2, splat2(8.82378953974694e-05f));
  q = fma2(q, x2, splat2(0.0003602751239668578f));
  q = fma2(q, x2, splat2(-0.10522668808698654f));
  q = fma2(q, x2, splat2(-2.3020453453063965f));
  float a0, a1;
  unpack2(mul2(q, x), a0, a1);
  float t0, t1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  unpack2(add2(pack2(t0, t1), splat2(1.0f)), t0, t1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
  return pack2(r0, r1);
}
__device__ __forceinline__ void gelu_erf2(float& a, float& b) {
  const f32x2 x = pack2(a, b);
  unpack2(mul2(x, normal_cdf_fast2(x, mul2(x, x))), a, b);
}
// Forward GELU of a pair with ONE MUFU per pair: Phi(x) = 1/2 + 1/2 tanh(x H(x^2)) (the logistic polynomial above, halved:
// the same H as gelu_tanh_both2) with the tanh taken on both elements at once by tanh.approx.f16x2.  The GELU epilogue of
// the fc1 GEMM is bound by the XU pipe in its math phase (4 warps per scheduler x 64 MUFU x 8 clk per 32-column piece:
// tests/gpu_epi_prof.py); this form issues a quarter of the MUFUs.  Error budget against the exact GELU: the f16 argument
// (|dPhi| <= 5.5e-5) and the f16 tanh (half an ulp below 1: |dPhi| <= 1.2e-4), i.e. |dg| <= 1.7e-4 |x| - below half an ulp
// of the fp16 value the result is stored as (2.4e-4 |g| .. 4.9e-4 |g|) wherever |g| > |x| / 2, i.e. for every x > 0.
__device__ __forceinline__ void gelu_tanh_h2(float& a, float& b) {
  const f32x2 x = pack2(a, b);
  const f32x2 x2 = mul2(x, x);
  f32x2 hh = fma2(splat2(1.1190779787284555e-06f), x2, splat2(-3.058092624996789e-05f));
  hh = fma2(hh, x2, splat2(-0.00012486183550208807f));
  hh = fma2(hh, x2, splat2(0.03646879270672798f));
  hh = fma2(hh, x2, splat2(0.7978281378746033f));
  float z0, z1;
  unpack2(mul2(hh, x), z0, z1);
  uint32_t zh, th;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(zh) : "f"(z1), "f"(z0));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(zh));
  const float2 t = unpack_f16(th);
  unpack2(mul2(x, fma2(splat2(0.5f), pack2(t.x, t.y), splat2(0.5f))), a, b);
}

// gelu and gelu' of a pair
__device__ __forceinline__ void gelu_erf_both2(float xa, float xb, f32x2& g, f32x2& dg) {
  const f32x2 x = pack2(xa, xb);
  const f32x2 x2 = mul2(x, x);
  const f32x2 cdf = normal_cdf_fast2(x, x2);
  float e0, e1;
  unpack2(mul2(x2, splat2(-0.72134752044448170368f)), e0, e1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(e1));
  g = mul2(x, cdf);
  dg = fma2(mul2(x, splat2(0.39894228040143267794f)), pack2(e0, e1), cdf);
}

// Backward-only variant with ONE MUFU per element: Phi(x) = 1/2 + 1/2 tanh(x H(x^2)) (the same logistic polynomial, halved)
// and phi(x) = Phi'(x) = 1/2 (1 - t^2) d/dx[x H(x^2)] evaluated as a second polynomial - no ex2 / rcp.  The GELU' epilogue
// is bound by the XU pipe (MUFU and F2FP share it at 8 cycles per warp instruction: 3 MUFU + 1 F2FP-equivalent per
// element = 21 us for the fc2 dgrad); this form needs 1 MUFU.  tanh.approx has 2^-11 relative error: |dg| error <= 3e-3
// and |g| error <= 2.4e-4 |x|, both below the bf16 rounding of the stored values - not used in the forward.
__device__ __forceinline__ void gelu_tanh_both2(float xa, float xb, f32x2& g, f32x2& dg) {
  const f32x2 x = pack2(xa, xb);
  const f32x2 x2 = mul2(x, x);
  f32x2 hh = fma2(splat2(1.1190779787284555e-06f), x2, splat2(-3.058092624996789e-05f));
  hh = fma2(hh, x2, splat2(-0.00012486183550208807f));
  hh = fma2(hh, x2, splat2(0.03646879270672798f));
  hh = fma2(hh, x2, splat2(0.7978281378746033f));
  float a0, a1, t0, t1;
  unpack2(mul2(hh, x), a0, a1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  const f32x2 t = pack2(t0, t1);
  f32x2 dd = fma2(splat2(5.035850790591212e-06f), x2, splat2(-0.0001070332364179194f));
  dd = fma2(dd, x2, splat2(-0.00031215461785905063f));
  dd = fma2(dd, x2, splat2(0.05470318719744682f));
  dd = fma2(dd, x2, splat2(0.39891406893730164f));
  const f32x2 sig = fma2(splat2(0.5f), t, splat2(0.5f));
  const f32x2 om = fma2(mul2(t, splat2(-1.0f)), t, splat2(1.0f));  // 1 - t^2
  g = mul2(x, sig);
  dg = fma2(mul2(x, om), dd, sig);
}

// same, with independent A / B element formats (0 = fp16, 1 = bf16): kind::f16 accepts mixed 16-bit operands
__host__ __device__ constexpr uint32_t make_idesc2(uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N,
                                                   uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

}  // namespace mfv
