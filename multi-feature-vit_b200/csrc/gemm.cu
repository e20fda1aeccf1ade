// tcgen05/TMEM 16-bit GEMM for every dense contraction of the ViT-S/16 block (SURVEY K3/K5/K6 and their dgrad/wgrad):
//
//     C[g][m][n] = epilogue( sum_k A[g][m][k] * B[g][n][k] )          g < G groups (the two MF-ViT branches)
//
// * warp-specialised, persistent: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer (+TMEM owner),
//   warps 2..5 = epilogue.  Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
//   tile i+1 (K is only 6 k-blocks for K=384).
// * operands are staged by TMA into 128B-swizzled shared memory through 3-D tensor maps (inner, rows, group); rows
//   or reduction elements past the tensor end are zero-filled by TMA, so no padding of M=B*197 tokens is needed.
// * each operand may be K-major (reduction dim contiguous: activations / nn.Linear weights in forward) or MN-major
//   (reduction dim strided: W in dgrad, dY and X in wgrad) - selected by the UMMA descriptors, no transposed copies.
// * epilogue: tcgen05.ld (lane = row) -> registers -> 128B-swizzled smem staging -> TMA bulk-tensor store, one
//   32-row box per epilogue warp, so every global access of the epilogue is a full-line asynchronous bulk copy
//   (round-1 ncu: per-thread row stores left the kernel latency-bound at 6-20 % tensor-pipe utilisation).  The fp32
//   residual / pre-GELU operand of the fused epilogues arrives the same way (TMA load, prefetched one chunk ahead);
//   split-K weight gradients leave through TMA reduce-add (cp.reduce.async.bulk.tensor .add.f32).
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;             // four warps per TMEM lane quarter: each owns every fourth column piece
constexpr int GEMM_THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int EPI_BUF = 2048;                 // one staging slot: 32 rows x 64 B = one 64B-swizzled TMA box
// slots in each epilogue warp's private ring: 3, or 4 for the fc2-dgrad epilogue (two stores per piece: with four slots
// the pre-GELU tiles of BOTH pieces of a tile are requested at tile start - with three, the second one waited for the
// previous tile's last store and its L2 latency was exposed, 12 % of the stall samples in profiles/r01_summary.md)
constexpr int epi_nbuf(int epi) { return epi == MFV_EPI_DGELU ? 4 : 3; }
constexpr bool GELU_H2 = false;               // forward GELU through tanh.approx.f16x2 (one MUFU per pair), see common.cuh
constexpr int MAX_STAGES = 6;                 // barrier slots reserved per operand ring
constexpr int AUX_BARS = 2 * NUM_EPI_WARPS;   // aux (residual / pre-GELU) TMA loads: 2 in flight per epilogue warp

struct GemmParams {
  int M, N, K, G;
  int tiles_m, tiles_n, splits, kb_total, kb_per_split;
  int a_mn, b_mn;             // 1 = MN-major operand
  int a_f16, b_f16, out_f16;  // 1 = IEEE fp16 instead of bf16 (A and B must agree: mixed 16-bit operands trap)
  int epi;
  int has_c2, has_c3;
  int dbg_skip_epilogue;  // measurement aid (dtype_flags bits 8..9): see launch_gemm
  int dbg_mma;            // measurement aid (dtype_flags bits 16..19, epilogue skipped): 1 = no TMA loads (MMAs on stale smem),
                          // 2 / 3 = + only the N=256 / only the N=128 UMMA of a 384-wide tile, 4 = + two N=192 UMMAs
  unsigned long long* dbg_prof;  // measurement aid (dtype_flags bit 20; the buffer rides in row_sum): per epilogue warp,
                          // clock64 sums of 8 phases - [CTA][16 epilogue warps + the UMMA issuer][8], see PROF_MARK
  int dbg_stages;         // measurement aid (dtype_flags bits 12..15, only with the epilogue skipped): operand ring depth;
                          // stages past S::STAGES lie over the unused epilogue rings
  int rows_cta;           // rows of the output tile each CTA owns: 128, or 96 (K-major A only; see launch_gemm_epi)
  float* row_sum;         // BN == 384, fp32 reduce-add epilogue: += row sums of A (bias gradient of a wgrad GEMM)
  // Second problem of a paired launch (mfv_gemm_wgrad_pair: two weight-gradient GEMMs that share the reduction length,
  // split count, groups and operand formats run as ONE grid; tiles [0, tiles0) belong to problem 0, the rest to problem
  // 1, whose operand / output maps ride in the tmC2 / tmC3 / tmAux slots).  tiles0 = INT_MAX for ordinary launches.
  int tiles0, M1, N1, tiles_m1, tiles_n1;
  float* row_sum1;
  long long bias_gstride;
  const float* bias;
  // MFV_EPI_RESID_LN: LayerNorm of the finished row, fused into the epilogue (gamma / beta share bias_gstride)
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  int ln_out_f32;
  // Stream-K (sk_units > 0): the (tile, k-block) space is cut into equal ranges of sk_units k-blocks per CTA pair
  // instead of whole tiles, so a 50-tile GEMM keeps all 74 pairs busy for 0.68 tile times (and a 100-tile one for 1.35
  // instead of 2).  Only the FIRST fragment of a pair's range can start inside a tile (k-blocks [x, ..) with x > 0): the
  // pair dumps that partial accumulator into its workspace slot and raises its counter, before anything else it does.
  // Only the LAST fragment can be the head of a tile it does not finish (k-blocks [0, y)): the pair owns that tile - it
  // adds the partials of the pairs behind it that cover the rest of the tile, then runs the normal epilogue.
  int sk_units;
  float* sk_ws;        // [pairs * CG][16 warps][BN / 4 * 32] fp32
  unsigned* sk_flags;  // [pairs * CG][2]: partial written (16 warp arrivals) / partial consumed (16)
};

// CG = 1: one CTA computes a 128 x BN tile.  CG = 2: a CTA pair (cta_group::2) computes 256 x BN; each CTA stages its own
// 128 rows of A and HALF of the B tile, so the operand bytes per UMMA cycle are halved (128 x 128 single-CTA tiles: measured
// 52 % of the tensor peak).
// BN = 384 (pairs only): the 256 x 384 output tile of a CTA pair is two UMMAs per k-step (N = 256 and N = 128) into one
// 384-column accumulator - for the N = 384 GEMMs (proj, fc2, every dgrad, qkv/fc1 wgrad) the A operand is then read from
// L2 exactly once instead of three times.  With the UMMAs issued from uniform control flow (see elect_one() in common.cuh)
// the mainloops are UMMA-bound: 820 clk per k-block for the 256 x 384 tile (768 at the peak), 550 for 256 x 256 (512), with
// or without the operand loads, K- or MN-major, 2 to 6 stages (tests/gpu_ring_probe*.py).
template <int BN, int CG, int NBUF, int EPI>
struct GemmSmem {
  static_assert(BN != 384 || CG == 2, "384-wide tiles need a CTA pair");
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NACC = (2 * BN <= 512) ? 2 : 1;  // TMEM accumulators: double-buffered when two fit in 512 columns
  static constexpr int BAR_BYTES = 512;
  static constexpr int ONES_BYTES = (BN == 384 && EPI == MFV_EPI_F32) ? 2048 : 0;  // [16 k][64 n] tile of 1.0 for the row-sum UMMA
  static constexpr int LN_BYTES = (EPI == MFV_EPI_RESID_LN) ? NUM_EPI_WARPS * 32 * 8 : 0;  // (mean, M2) of each warp's share of a row
  // bias (and LayerNorm gamma, beta) of the tile's 384 columns, staged once per tile: with 227 KB of the SM's 228 KB
  // configured as shared memory there is no L1 left, and every __ldg of these vectors in the epilogue was an L2 round
  // trip (~700 clk per piece: tests/gpu_epi_prof.py)
  static constexpr int PAR_BYTES = (BN != 384) ? 0 : (EPI == MFV_EPI_RESID_LN) ? 3 * 384 * 4 : (EPI == MFV_EPI_RESID_F32) ? 384 * 4 : 0;
  static constexpr int EPI_BYTES = NUM_EPI_WARPS * NBUF * EPI_BUF;
  // operand stages: what is left of the 227 KB after the epilogue rings (3 slots: 6/5/4/4/3/2 stages for 16/24/32/32/
  // 40/48 KB stages, 4 slots: one fewer from 32 KB up), at most 6
  static constexpr int AVAIL = 232448 - 1024 - BAR_BYTES - ONES_BYTES - LN_BYTES - PAR_BYTES - EPI_BYTES;
  static constexpr int STAGES = (AVAIL / STAGE_BYTES > 6) ? 6 : AVAIL / STAGE_BYTES;
  static_assert(STAGES >= 2, "not enough shared memory for a double-buffered mainloop");
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + ONES_BYTES + LN_BYTES + PAR_BYTES + 1024;  // +1024: alignment
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1,
                                                  int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// at most n (clamped to 0..3) of this thread's newest bulk groups may still be reading their shared-memory source
__device__ __forceinline__ void bulk_wait_read_n(int n) {
  if (n <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (n == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else if (n == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
}

// 16-byte chunk j (0..3) of row r in a 32 x 64 B staging slot laid out with the TMA 64B swizzle (chunk index XOR
// address bits [7,9) = (r >> 1) & 3); a warp's 32 x 16 B store then covers every bank exactly 4 times (conflict-free)
__device__ __forceinline__ uint32_t stage_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

struct TileInfo { int prob, n_tile, split, m_tile, g; };
__device__ __forceinline__ TileInfo decode_tile(const GemmParams& p, int t) {
  TileInfo ti;
  ti.prob = 0;
  int tm = p.tiles_m, tn = p.tiles_n;
  if (t >= p.tiles0) { t -= p.tiles0; ti.prob = 1; tm = p.tiles_m1; tn = p.tiles_n1; }
  ti.n_tile = t % tn; t /= tn;
  ti.split = t % p.splits; t /= p.splits;
  ti.m_tile = t % tm;
  ti.g = t / tm;
  return ti;
}

// One unit of work of a CTA pair: k-blocks [kb0, kb1) of tile t.  role 0 = the whole reduction range of the tile (or of
// its split), 1 = stream-K tail (dump the partial), 2 = stream-K head (add the next pair's partial, then the epilogue).
struct Frag { int t, kb0, kb1, role, contrib; };  // contrib: pairs behind this one that hold the rest of the tile (role 2)
struct FragCursor {
  int u, u1, step;
  __device__ __forceinline__ void init(const GemmParams& p, int cta_id, int num_ctas, int total_tiles) {
    if (p.sk_units > 0) {
      u = cta_id * p.sk_units;
      u1 = min(u + p.sk_units, total_tiles * p.kb_total);
      step = 0;
    } else {
      u = cta_id; u1 = total_tiles; step = num_ctas;
    }
  }
  __device__ __forceinline__ bool next(const GemmParams& p, Frag& f) {
    if (u >= u1) return false;
    if (p.sk_units > 0) {
      f.t = u / p.kb_total;
      f.kb0 = u - f.t * p.kb_total;
      f.kb1 = min(p.kb_total, f.kb0 + (u1 - u));
      f.role = f.kb0 > 0 ? 1 : (f.kb1 < p.kb_total ? 2 : 0);
      u += f.kb1 - f.kb0;
      // role 2: this fragment ends at u == u1 (the end of this pair's range); the tile ends at (t + 1) * kb_total
      f.contrib = f.role == 2 ? ((f.t + 1) * p.kb_total - 1) / p.sk_units - (u1 - 1) / p.sk_units : 0;
    } else {
      f.t = u; f.role = 0; f.kb0 = -1; f.kb1 = -1; f.contrib = 0;  // k range from the tile's split index
      u += step;
    }
    return true;
  }
};

// EPI is a template parameter (MFV_EPI_ATOMIC_F32 shares the MFV_EPI_F32 instance): the epilogue is the issue-bound
// part of these kernels, and a specialised instruction stream keeps it small (I-cache) and spill-free.
// MC = 1 (BN = 384 pair tiles only): clusters of FOUR CTAs = two pairs that work on two adjacent row tiles of the same
// column tile in lockstep and share its B operand - each CTA loads half of its pair's B half and TMA-multicasts it to
// the CTA of the same pair rank in the other pair, so the weights are read from L2 once per two row tiles.  (These
// mainloops are bound by operand bytes out of L2: profiles/r02_summary.md.)  Stage release then needs both pairs: every
// MMA commit arrives on the empty barrier of all four CTAs, which count two arrivals.
template <int BN, int CG, int EPI, int MC = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                 const __grid_constant__ CUtensorMap tmC3, const __grid_constant__ CUtensorMap tmAux,
                 const GemmParams p) {
  constexpr int EPI_NBUF = epi_nbuf(EPI);
  using S = GemmSmem<BN, CG, EPI_NBUF, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + S::STAGES * S::STAGE_BYTES;
  uint8_t* bar_base = epi_base + S::EPI_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint8_t* ones_tile = bar_base + S::BAR_BYTES;  // BN == 384 only
  float2* ln_part = reinterpret_cast<float2*>(ones_tile + S::ONES_BYTES);  // MFV_EPI_RESID_LN only: [16 warps][32 lanes]
  [[maybe_unused]] float* par = reinterpret_cast<float*>(ones_tile + S::ONES_BYTES + S::LN_BYTES);  // bias | gamma | beta
  uint64_t* aux_bar = tempty_bar + 2;  // [NUM_EPI_WARPS][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + AUX_BARS);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int nst = p.dbg_stages ? p.dbg_stages : S::STAGES;
  static_assert(MC == 0 || (BN == 384 && CG == 2), "B multicast is built for the 384-wide pair tiles");
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;  // rank in the cluster: 0..1, or 0..3 with MC
  const uint32_t rank = crank & 1u;                           // 0 = leader of the pair (issues the MMAs)
  const uint32_t pq = MC ? (crank >> 1) : 0u;                 // which pair of the cluster
  // persistent schedule runs over clusters (with MC a cluster takes a pair of row tiles per step)
  const int cta_id = blockIdx.x / (CG * (MC ? 2 : 1)), num_ctas = gridDim.x / (CG * (MC ? 2 : 1));
  constexpr int NACC = S::NACC;
  constexpr uint32_t TMEM_COLS = (NACC * BN <= 32) ? 32 : (NACC * BN <= 64) ? 64 : (NACC * BN <= 128) ? 128 : (NACC * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], NUM_EPI_WARPS * CG);  // one arrival per epilogue warp of the pair
    }
    for (int s = 0; s < AUX_BARS; ++s) mbar_init(&aux_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_cg2(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  if (S::ONES_BYTES > 0 && warp == 2 && (p.row_sum || p.row_sum1)) {  // bf16 1.0 everywhere: the swizzle is irrelevant for a constant tile
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<uint4*>(ones_tile)[lane + 32 * i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();  // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  const int total_tiles = p.tiles_m * p.tiles_n * p.splits * p.G +
                          (p.tiles0 == 0x7fffffff ? 0 : p.tiles_m1 * p.tiles_n1 * p.splits * p.G);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      griddep_wait();  // PDL: operands may still be in flight in the previous kernel
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
        if (CG == 2) tma_load_3d_cg2(dst, map, bar, c0, c1, c2); else tma_load_3d(dst, map, bar, c0, c1, c2);
      };
      FragCursor cur;
      cur.init(p, cta_id, num_ctas, total_tiles);
      Frag fr;
      while (cur.next(p, fr)) {
        const TileInfo ti = decode_tile(p, fr.t);
        const int n_tile = ti.n_tile, split = ti.split, g = ti.g;
        const CUtensorMap* mapA = ti.prob ? &tmC2 : &tmA;
        const CUtensorMap* mapB = ti.prob ? &tmC3 : &tmB;
        const int kb0 = fr.kb0 >= 0 ? fr.kb0 : split * p.kb_per_split;
        const int kb1 = fr.kb0 >= 0 ? fr.kb1 : min(kb0 + p.kb_per_split, p.kb_total);
        const int m_tile = MC ? ti.m_tile * 2 + (int)pq : ti.m_tile;  // MC: p.tiles_m counts PAIRS of row tiles
        const int m0 = (m_tile * CG + (int)rank) * p.rows_cta;       // this CTA's rows of A (past M: TMA zero-fills)
        const int n0 = n_tile * BN + (int)rank * (BN / CG);          // this CTA's share of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if (p.dbg_mma && p.dbg_mma < 13) {
            if (rank == 0 && elect_one()) mbar_arrive(&full_bar[stage]);
            __syncwarp();
            if (++stage == nst) { stage = 0; phase ^= 1; }
            continue;
          }
          if (elect_one()) {
            // the leader's barrier collects the bytes of both CTAs' loads
            // measurement aid: dbg_mma 13 = B loads only, 14 = A loads only (which operand's delivery costs what)
            const bool ld_a = p.dbg_mma != 13, ld_b = p.dbg_mma != 14;
            if (rank == 0)
              mbar_arrive_expect_tx(&full_bar[stage], ((ld_b ? S::STAGE_BYTES - S::A_BYTES : 0) + (ld_a ? p.rows_cta * BK * 2 : 0)) * CG);
            if (!ld_a) {
            } else if (!p.a_mn) {
              load(sa, mapA, &full_bar[stage], kb * BK, m0, g);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) load(sa + j * 8192, mapA, &full_bar[stage], m0 + j * 64, kb * BK, g);
            }
            if (!ld_b) {
            } else if (BN == 384) {
              // pair tile = UMMA N=256 (each CTA supplies rows [rank*128, +128) of it) + UMMA N=128 (rows 256 + rank*64)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int nn = n_tile * BN + (j < 2 ? (int)rank * 128 + j * 64 : 256 + (int)rank * 64);
                if constexpr (MC) {
                  // pair 0 fetches boxes 0 and 1, pair 1 box 2; every box goes to this CTA and to the CTA of the same
                  // pair rank in the other pair (cluster ranks rank and rank + 2)
                  if ((j < 2) != (pq == 0)) continue;
                  const uint16_t mask = (uint16_t)(0x5u << rank);
                  if (!p.b_mn) tma_load_3d_cg2_mc(sb + j * 8192, mapB, &full_bar[stage], kb * BK, nn, g, mask);
                  else tma_load_3d_cg2_mc(sb + j * 8192, mapB, &full_bar[stage], nn, kb * BK, g, mask);
                } else {
                  if (!p.b_mn) load(sb + j * 8192, mapB, &full_bar[stage], kb * BK, nn, g);
                  else load(sb + j * 8192, mapB, &full_bar[stage], nn, kb * BK, g);
                }
              }
            } else if (!p.b_mn) {
              load(sb, mapB, &full_bar[stage], kb * BK, n0, g);
            } else {
#pragma unroll
              for (int j = 0; j < BN / CG / 64; ++j) load(sb + j * 8192, mapB, &full_bar[stage], n0 + j * 64, kb * BK, g);
            }
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp of the leader CTA runs the
    // loop; one elected lane issues the tcgen05 instructions)
    if (rank == 0) {
      // measurement aid: UMMA shapes / TMEM columns / B offsets of the two UMMAs of a 384-wide k-step (p.dbg_mma)
      int pn1 = 256, pn2 = 128;
      uint32_t pcol2 = 256u, pboff2 = 16384u;
      switch (p.dbg_mma) {
        case 2: pn2 = 0; break;
        case 3: pn1 = 0; break;
        case 4: pn1 = 192; pn2 = 192; pcol2 = 192u; pboff2 = 12288u; break;
        case 5: pn1 = 0; pcol2 = 0u; break;
        case 6: pn1 = 128; pn2 = 0; break;
        case 7: pn1 = 0; pboff2 = 0u; break;
        case 8: pn1 = 0; pn2 = 256; pcol2 = 0u; pboff2 = 0u; break;
        case 9: pn1 = 128; pn2 = 128; pcol2 = 128u; pboff2 = 8192u; break;
        case 10: pcol2 = 0u; break;
        case 11: pn1 = 0; pn2 = 64; break;
        case 12: pn1 = 0; pn2 = 256; pcol2 = 256u; pboff2 = 0u; break;
        default: break;
      }
      const uint32_t idesc = make_idesc2(p.a_f16 ? 0u : 1u, p.b_f16 ? 0u : 1u, BM * CG,
                                         BN == 384 ? (uint32_t)(pn1 ? pn1 : 256) : BN, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
      const uint32_t idesc2 = make_idesc2(p.a_f16 ? 0u : 1u, p.b_f16 ? 0u : 1u, BM * CG, (uint32_t)(pn2 ? pn2 : 128),
                                          (uint32_t)p.a_mn, (uint32_t)p.b_mn);  // second UMMA of a 384-wide tile
      const uint32_t idesc3 = make_idesc2(p.a_f16 ? 0u : 1u, p.b_f16 ? 0u : 1u, BM * CG, 16, (uint32_t)p.a_mn, 1u);
      const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
      const uint32_t a_kadv = p.a_mn ? (UMMA_K * 128u) : (UMMA_K * 2u);
      const uint32_t b_kadv = p.b_mn ? (UMMA_K * 128u) : (UMMA_K * 2u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      // measurement aid (dtype_flags bit 20): slot 16 of the CTA's phase clocks = the issuer: [0] waiting for a free
      // accumulator, [1] in the k-loop (waits for operands + issue), [2] tiles
      unsigned long long mprof[3] = {0, 0, 0};
      FragCursor cur;
      cur.init(p, cta_id, num_ctas, total_tiles);
      Frag fr;
      for (; cur.next(p, fr); ++it) {
        const long long mt0 = p.dbg_prof ? clock64() : 0;
        const TileInfo ti = decode_tile(p, fr.t);
        const int split = ti.split;
        const bool want_rs = S::ONES_BYTES > 0 && (ti.prob ? p.row_sum1 : p.row_sum) != nullptr;
        const int kb0 = fr.kb0 >= 0 ? fr.kb0 : split * p.kb_per_split;
        const int kb1 = fr.kb0 >= 0 ? fr.kb1 : min(kb0 + p.kb_per_split, p.kb_total);
        const int as = it % NACC;
        const uint32_t aphase = (it / NACC) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const long long mt1 = p.dbg_prof ? clock64() : 0;
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = make_smem_desc_sw128(sa + k * a_kadv, a_lbo, 1024u);
              const uint64_t db = make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u);
              if (BN == 384 && pn1 == 0) {
              } else if (CG == 2) umma_bf16_cg2(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              if (BN == 384 && pn2 != 0) {
                const uint64_t db2 = make_smem_desc_sw128(sb + pboff2 + k * b_kadv, b_lbo, 1024u);
                umma_bf16_cg2(tmem_d + pcol2, da, db2, idesc2, (kb > kb0 || k > 0) ? 1u : 0u);
                if (want_rs)  // columns 384..399 += A . ones: every column is the row sum of A over this k-step
                  umma_bf16_cg2(tmem_d + 384u, da, make_smem_desc_sw128(smem_u32(ones_tile), 8192u, 1024u), idesc3,
                                (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
            // frees the smem stage in BOTH CTAs (their producers wait on their own empty barrier)
            if (CG == 2) umma_commit_cg2(&empty_bar[stage], MC ? 0xF : 0x3); else umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) {
              if (CG == 2) umma_commit_cg2(&tfull_bar[as], (uint16_t)(0x3u << (2 * pq))); else umma_commit(&tfull_bar[as]);
            }
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        if (p.dbg_prof) {
          const long long mt2 = clock64();
          mprof[0] += (unsigned long long)(mt1 - mt0); mprof[1] += (unsigned long long)(mt2 - mt1); mprof[2] += 1;
        }
      }
      if (p.dbg_prof && lane == 0) {
        for (int k = 0; k < 3; ++k) p.dbg_prof[((size_t)blockIdx.x * (NUM_EPI_WARPS + 1) + NUM_EPI_WARPS) * 8 + k] = mprof[k];
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // 16 warps.  Warp (q, h): q = TMEM lane quarter it may access (rows q*32 + lane of the tile), h = 0..3 selects the
    // column pieces it owns (piece index = h mod 4).  A piece is 32 rows x 64 B of the primary output (32 16-bit or 16
    // fp32 columns) = one 64B-swizzled TMA box of 2 KB.  Every warp works alone - its own ring of EPI_NBUF staging
    // slots and its own bulk-store groups; there is no CTA-level barrier in the epilogue - so four
    // warps per scheduler hide each other's TMEM / MUFU / TMA latencies (8 warps left the GELU epilogue latency-bound at
    // 17 % issue utilisation, profiles/r01_ncu_gemm_fc1_v2.md).
    // Ring protocol: every slot use ends in exactly one committed bulk group (lane 0), so slot (use % EPI_NBUF) is free
    // again once at most EPI_NBUF-1 newer groups are still reading (cp.async.bulk.wait_group.read).
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;
    // phase clocks (measurement aid): 0 accumulator wait, 1 TMEM read, 2 slot / aux wait, 3 convert + staging stores,
    // 4 fence + bulk store issue, 5 epilogue math, 6 LayerNorm row barrier, 7 final drain
    unsigned long long* const prof = p.dbg_prof;
    long long prof_t = 0;
    unsigned long long prof_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_MARK(slot)                              \
  if (prof) {                                        \
    const long long now_ = clock64();                \
    prof_c[slot] += (unsigned long long)(now_ - prof_t); \
    prof_t = now_;                                   \
  }
    uint8_t* ring = epi_base + ew * (EPI_NBUF * EPI_BUF);
    uint32_t use = 0;  // ring uses so far
    uint64_t* abar = aux_bar + 2 * ew;
    uint32_t aux_phase = 0u;  // bit j = parity of abar[j]
    constexpr int epi = EPI;
    constexpr bool has_aux = (epi == MFV_EPI_RESID_F32 || epi == MFV_EPI_DGELU || epi == MFV_EPI_RESID_LN);
    constexpr bool out32 = (epi == MFV_EPI_RESID_F32 || epi == MFV_EPI_F32 || epi == MFV_EPI_RESID_LN);
    static_assert(epi != MFV_EPI_RESID_LN || BN == 384, "the fused LayerNorm needs a tile that owns whole rows");
    constexpr int PW = out32 ? 16 : 32;  // columns per piece
    constexpr int npieces = BN / PW;
    auto slot_ptr = [&](uint32_t u) { return ring + (u % EPI_NBUF) * EPI_BUF; };
    // make the next slot writable: the bulk store that used it EPI_NBUF uses ago has finished reading it
    auto acquire_slot = [&]() {
      PROF_MARK(5)
      if (elect_one()) bulk_wait_read_n(EPI_NBUF - 1);
      __syncwarp();
      PROF_MARK(2)
    };
    auto publish = [&](const CUtensorMap* map, const uint8_t* src, int c0, int r0, int gg, bool reduce) {
      PROF_MARK(3)
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {  // the elected lane is the same one every time: bulk groups are per thread
        if (!(p.dbg_skip_epilogue & 2)) {
          if (reduce) tma_reduce_add_3d(map, src, c0, r0, gg); else tma_store_3d(map, src, c0, r0, gg);
        }
        bulk_commit();
      }
      ++use;
      PROF_MARK(4)
    };
    auto release_accumulator = [&](int as) {  // all tcgen05.ld of this warp have completed (wait::ld is warp-wide)
      tc_fence_before();
      __syncwarp();
      if (elect_one()) {
        if (CG == 2) mbar_arrive_cluster_relaxed(&tempty_bar[as], 2 * pq); else mbar_arrive_relaxed(&tempty_bar[as]);
      }
    };
    auto pack16 = [&](float a, float b) { return p.out_f16 ? pack_f16(a, b) : pack_bf16(a, b); };
    griddep_wait();  // PDL: everything above overlapped the previous kernel's tail
    if (prof) prof_t = clock64();
    int it = 0;
    [[maybe_unused]] int par_g = -1;  // group whose bias / gamma / beta the parameter block holds
    constexpr bool bias_by_shfl = (S::PAR_BYTES == 0) && (PW == 32) && (npieces <= 8);
    FragCursor cur;
    cur.init(p, cta_id, num_ctas, total_tiles);
    Frag fr;
    // stream-K workspace of this warp: [slot = pair * CG + rank][16 warps][BN / 4 * 32 fp32], piece i at i * 32 * PW,
    // lane's PW values contiguous (a warp reads / writes 32 x 64 B or 32 x 128 B in one piece: fully coalesced)
    constexpr int SK_WARP_FLOATS = BN / 4 * 32;
    auto sk_slot = [&](int pair) {
      return p.sk_ws + ((size_t)(pair * CG + (int)rank) * NUM_EPI_WARPS + ew) * SK_WARP_FLOATS;
    };
    auto sk_flag = [&](int pair) { return p.sk_flags + (size_t)(pair * CG + (int)rank) * 2; };
    for (; cur.next(p, fr); ++it) {
      const TileInfo ti = decode_tile(p, fr.t);
      const int n_tile = ti.n_tile, m_tile = MC ? ti.m_tile * 2 + (int)pq : ti.m_tile, g = ti.g;
      const int pM = ti.prob ? p.M1 : p.M, pN = ti.prob ? p.N1 : p.N;
      const CUtensorMap* mapC = ti.prob ? &tmAux : &tmC;   // paired launches only exist for the fp32 reduce-add epilogue
      float* const row_sum = ti.prob ? p.row_sum1 : p.row_sum;
      const int as = it % NACC;
      const uint32_t aphase = (it / NACC) & 1;
      const int row0 = (m_tile * CG + (int)rank) * p.rows_cta + q * 32;  // first row of this warp's 32-row slice
      const int ncol0 = n_tile * BN;
      const float* bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
      // pieces owned by this warp that hold real output (slices past M / N are skipped; TMA clips partial ones)
      int n_my = 0;
      if (row0 < pM && q * 32 < p.rows_cta)  // 96-row CTAs: TMEM lanes 96..127 hold rows the neighbour tile owns
        for (int c = h; c < npieces && ncol0 + c * PW < pN; c += 4) ++n_my;
      // The aux operand (fp32 residual / bf16 pre-GELU u) of piece i arrives by TMA in the ring slot where the result
      // of piece i is then computed in place.  AUX_AHEAD pieces are in flight: the load of piece i+AUX_AHEAD is issued
      // right after the first store of piece i, into the slot whose previous store is then the second-newest bulk group
      // (wait_group.read 1 - never a wait on the store just issued).  upc = ring slots used per piece.
      const int upc = (epi == MFV_EPI_DGELU && p.has_c2) ? 2 : 1;
      const int aux_ahead = (upc == 1 || EPI_NBUF >= 4) ? 2 : 1;
      const uint32_t use0 = use;
      auto issue_aux = [&](int i, int pending_ok) {  // the elected lane only
        bulk_wait_read_n(pending_ok);
        mbar_arrive_expect_tx(&abar[i & 1], EPI_BUF);
        tma_load_3d(slot_ptr(use0 + (uint32_t)(i * upc)), &tmAux, &abar[i & 1], ncol0 + (h + 4 * i) * PW, row0, g);
      };
      if constexpr (has_aux) {
        if (fr.role != 1) {
          if (elect_one())
            for (int j = 0; j < aux_ahead && j < n_my; ++j) issue_aux(j, (int)EPI_NBUF - 1 - j * upc);
          __syncwarp();
        }
      }
      // Bias (gamma, beta) of this tile, fetched while its MMAs are still running.  384-wide tiles: the whole vectors
      // into the shared parameter block (all 16 warps; re-staged only when the group changes).  Narrower tiles with
      // 32-column pieces: lane L keeps bias[piece column L] of the warp's (at most two) pieces in a register and the
      // piece loop broadcasts it by shuffle.  Either way no epilogue thread waits on L2 for them.
      [[maybe_unused]] float breg0 = 0.f, breg1 = 0.f;
      if constexpr (S::PAR_BYTES > 0) {
        if (bias && fr.role != 1 && g != par_g) {
          if (par_g >= 0) named_bar_sync(5u, 32u * NUM_EPI_WARPS);  // every warp has finished with the previous group's values
          for (int idx = ew * 32 + lane; idx < 384; idx += 32 * NUM_EPI_WARPS) {
            par[idx] = __ldg(bias + idx);
            if constexpr (EPI == MFV_EPI_RESID_LN) {
              par[384 + idx] = __ldg(p.ln_gamma + (long long)g * p.bias_gstride + idx);
              par[768 + idx] = __ldg(p.ln_beta + (long long)g * p.bias_gstride + idx);
            }
          }
          named_bar_sync(5u, 32u * NUM_EPI_WARPS);
          par_g = g;
        }
      } else if constexpr (bias_by_shfl) {
        if (bias && fr.role != 1) {
          if (n_my > 0) breg0 = __ldg(bias + ncol0 + h * PW + lane);
          if (n_my > 1) breg1 = __ldg(bias + ncol0 + (h + 4) * PW + lane);
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      PROF_MARK(0)
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (fr.role == 1) {
        // ---- stream-K tail: this pair's partial sums of a tile the previous pair finishes.  Dump, raise the counter.
        if (!(p.dbg_skip_epilogue & 1) && !(p.dbg_skip_epilogue & 8)) {
          float* ws = sk_slot(cta_id) + lane * PW;
#pragma unroll 1
          for (int i = 0; i < n_my; ++i) {
            const int c = h + 4 * i;
            uint32_t v[32];
            if constexpr (!out32) tmem_ld32(trow + (uint32_t)(c * PW), v); else tmem_ld16(trow + (uint32_t)(c * PW), v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < PW; k += 4)
              __stcg(reinterpret_cast<float4*>(ws + i * 32 * PW + k),
                     make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]),
                                 __uint_as_float(v[k + 3])));
          }
        }
        release_accumulator(as);
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(sk_flag(cta_id), 1u);
        continue;
      }
      const float* sk_part = nullptr;
      if (fr.role == 2) {
        // ---- stream-K head: the rest of this tile was the FIRST thing the pairs behind this one did; their partials are
        // (almost always) there already.  All 16 warps of a contributing CTA have arrived when its counter reads 16.
        if (lane == 0) {
          const long long t0 = clock64();
          for (int j = 1; j <= fr.contrib; ++j) {
            unsigned* fl = sk_flag(cta_id + j);
            while (*reinterpret_cast<volatile unsigned*>(fl) < (unsigned)NUM_EPI_WARPS) {
              if (clock64() - t0 > 4000000000LL) {
                printf("mfvit: stream-K partial never arrived (block %d warp %d)\n", blockIdx.x, warp);
                __trap();
              }
            }
          }
          __threadfence();
        }
        __syncwarp();
        sk_part = sk_slot(cta_id + 1) + lane * PW;
      }
      // every warp of a head fragment counts itself out, whatever path it leaves by; the last one re-arms the slot
      auto sk_consumed = [&]() {
        if (fr.role != 2) return;
        __syncwarp();
        if (lane == 0) {
          for (int j = 1; j <= fr.contrib; ++j) {
            unsigned* fl = sk_flag(cta_id + j);
            if (atomicAdd(fl + 1, 1u) == (unsigned)NUM_EPI_WARPS - 1u) {
              fl[1] = 0u;
              __threadfence();
              fl[0] = 0u;
            }
          }
        }
      };
      if (n_my == 0) {
        release_accumulator(as);
        sk_consumed();
        continue;
      }
      if (p.dbg_skip_epilogue & 1) {  // measurement aid: drain the aux loads already issued, touch nothing else
        sk_consumed();
        release_accumulator(as);
        if constexpr (has_aux) {
          for (int j = 0; j < aux_ahead && j < n_my; ++j) {
            mbar_wait(&abar[j & 1], (aux_phase >> (j & 1)) & 1u);
            aux_phase ^= 1u << (j & 1);
          }
        }
        continue;
      }
      if (BN == 384 && epi == MFV_EPI_F32 && row_sum && h == 0 && n_tile == 0) {
        uint32_t rs;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(rs) : "r"(trow + 384u) : "memory");
        tmem_ld_wait();
        if (row0 + lane < pM) atomicAdd(row_sum + (long long)g * p.bias_gstride + row0 + lane, __uint_as_float(rs));
      }
      [[maybe_unused]] float ln_m = 0.f, ln_s = 0.f;  // MFV_EPI_RESID_LN: mean and sum of squared deviations so far
#pragma unroll 1
      for (int i = 0; i < n_my; ++i) {
        const int c = h + 4 * i;
        const int n0 = ncol0 + c * PW;
        float f[32];
        {
          uint32_t v[32];
          if constexpr (!out32) {
            tmem_ld32(trow + (uint32_t)(c * PW), v);
          } else {
            tmem_ld16(trow + (uint32_t)(c * PW), v);
#pragma unroll
            for (int k = 16; k < 32; ++k) v[k] = 0u;
          }
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
        }
        PROF_MARK(1)
        if (sk_part) {  // stream-K head: + the partial sums of the pairs that hold the rest of this tile
          for (int j = 0; j < ((p.dbg_skip_epilogue & 4) ? 0 : fr.contrib); ++j) {
            const float* wp = sk_part + (size_t)j * (CG * NUM_EPI_WARPS * SK_WARP_FLOATS) + i * 32 * PW;
#pragma unroll
            for (int k = 0; k < PW; k += 4) {
              const float4 w4 = __ldcg(reinterpret_cast<const float4*>(wp + k));
              f[k] += w4.x; f[k + 1] += w4.y; f[k + 2] += w4.z; f[k + 3] += w4.w;
            }
          }
          if (i == n_my - 1) sk_consumed();
        }
        if (epi != MFV_EPI_RESID_LN && i == n_my - 1)
          release_accumulator(as);  // last TMEM read of the tile: the MMA warp may reuse the buffer
        if (bias) {
          if constexpr (S::PAR_BYTES > 0) {
            const float* bp = par + c * PW;  // the same address in every lane: one broadcast wavefront per load
#pragma unroll
            for (int k = 0; k < PW; k += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bp + k);
              f[k] += b4.x; f[k + 1] += b4.y; f[k + 2] += b4.z; f[k + 3] += b4.w;
            }
          } else if constexpr (bias_by_shfl) {
            const float bsel = (i == 0) ? breg0 : breg1;
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] += __shfl_sync(0xffffffffu, bsel, k);
          } else {
            const float* bp = bias + n0;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              if (k < PW) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + k));
                f[k] += b4.x; f[k + 1] += b4.y; f[k + 2] += b4.z; f[k + 3] += b4.w;
              }
            }
          }
        }
        uint8_t* st0 = slot_ptr(use);
        PROF_MARK(5)
        if constexpr (has_aux) {
          mbar_wait(&abar[i & 1], (aux_phase >> (i & 1)) & 1u);
          aux_phase ^= 1u << (i & 1);
        } else {
          acquire_slot();
        }
        PROF_MARK(2)
        if constexpr (epi == MFV_EPI_BF16) {
          {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(st0 + stage_off(lane, j)) =
                  make_uint4(pack16(f[8 * j], f[8 * j + 1]), pack16(f[8 * j + 2], f[8 * j + 3]),
                             pack16(f[8 * j + 4], f[8 * j + 5]), pack16(f[8 * j + 6], f[8 * j + 7]));
            publish(&tmC, st0, n0, row0, g, false);
          }
        } else if constexpr (epi == MFV_EPI_GELU) {
          {  // C = u (bf16, saved for backward), C2 = gelu(u) (fp16|bf16), C3 = optional bf16 copy of C2
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(st0 + stage_off(lane, j)) =
                  make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                             pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            publish(&tmC, st0, n0, row0, g, false);
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              if (GELU_H2) gelu_tanh_h2(f[k], f[k + 1]); else gelu_erf2(f[k], f[k + 1]);
            }
            uint8_t* st1 = slot_ptr(use);
            acquire_slot();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(st1 + stage_off(lane, j)) =
                  make_uint4(pack16(f[8 * j], f[8 * j + 1]), pack16(f[8 * j + 2], f[8 * j + 3]),
                             pack16(f[8 * j + 4], f[8 * j + 5]), pack16(f[8 * j + 6], f[8 * j + 7]));
            publish(&tmC2, st1, n0, row0, g, false);
            if (p.has_c3) {
              uint8_t* st2 = slot_ptr(use);
              acquire_slot();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(st2 + stage_off(lane, j)) =
                    make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                               pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
              publish(&tmC3, st2, n0, row0, g, false);
            }
          }
        } else if constexpr (epi == MFV_EPI_RESID_F32) {
          {  // C(fp32) = acc + bias + aux(fp32 residual stream), computed in place
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* pp = reinterpret_cast<float4*>(st0 + stage_off(lane, j));
              const float4 rr = *pp;
              *pp = make_float4(f[4 * j] + rr.x, f[4 * j + 1] + rr.y, f[4 * j + 2] + rr.z, f[4 * j + 3] + rr.w);
            }
            publish(&tmC, st0, n0, row0, g, false);
            if (i + aux_ahead < n_my) {
              if (elect_one()) issue_aux(i + aux_ahead, 1);
              __syncwarp();
            }
          }
        } else if constexpr (epi == MFV_EPI_RESID_LN) {
          {  // pass 1 of the fused LayerNorm: x = acc + bias + residual -> C (fp32) as above, x kept in TMEM over the
             // accumulator, and the running (mean, M2) of this thread's columns (pairwise / Chan update: no cancellation)
            float y[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* pp = reinterpret_cast<float4*>(st0 + stage_off(lane, j));
              const float4 rr = *pp;
              y[4 * j] = f[4 * j] + rr.x; y[4 * j + 1] = f[4 * j + 1] + rr.y;
              y[4 * j + 2] = f[4 * j + 2] + rr.z; y[4 * j + 3] = f[4 * j + 3] + rr.w;
              *pp = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
            }
            publish(&tmC, st0, n0, row0, g, false);
            if (i + aux_ahead < n_my) {
              if (elect_one()) issue_aux(i + aux_ahead, 1);
              __syncwarp();
            }
            {
              uint32_t yb[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) yb[k] = __float_as_uint(y[k]);
              tmem_st16(trow + (uint32_t)(c * PW), yb);
            }
            float s1 = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) s1 += y[k];
            const float pm = s1 * (1.0f / 16.0f);
            float m2 = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) { const float d = y[k] - pm; m2 = fmaf(d, d, m2); }
            if (i == 0) {
              ln_m = pm; ln_s = m2;
            } else {
              const float na = 16.0f * (float)i, nn = na + 16.0f, dl = pm - ln_m;
              ln_m = fmaf(dl, 16.0f / nn, ln_m);
              ln_s += fmaf(dl * dl, na * 16.0f / nn, m2);
            }
          }
        } else if constexpr (epi == MFV_EPI_DGELU) {
          {  // C(bf16) = acc * gelu'(u) in place over u = aux (bf16); C2 (optional) = gelu(u) bf16
            uint32_t gk[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4* pp = reinterpret_cast<uint4*>(st0 + stage_off(lane, j));
              const uint4 uu = *pp;
              const uint32_t uw[4] = {uu.x, uu.y, uu.z, uu.w};
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 u2 = unpack_bf16(uw[k]);
                f32x2 gg, dd;
                gelu_tanh_both2(u2.x, u2.y, gg, dd);
                float g0, g1, d0, d1;
                unpack2(gg, g0, g1);
                unpack2(mul2(pack2(f[8 * j + 2 * k], f[8 * j + 2 * k + 1]), dd), d0, d1);
                o[k] = pack_bf16(d0, d1);
                gk[4 * j + k] = pack_bf16(g0, g1);
              }
              *pp = make_uint4(o[0], o[1], o[2], o[3]);
            }
            publish(&tmC, st0, n0, row0, g, false);
            if (p.has_c2) {
              uint8_t* st1 = slot_ptr(use);
              acquire_slot();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(st1 + stage_off(lane, j)) =
                    make_uint4(gk[4 * j], gk[4 * j + 1], gk[4 * j + 2], gk[4 * j + 3]);
              publish(&tmC2, st1, n0, row0, g, false);
            }
            // the slot piece i + aux_ahead lands in was last read by a store that is at least the second-newest group
            if (i + aux_ahead < n_my) {
              if (elect_one()) issue_aux(i + aux_ahead, 1);
              __syncwarp();
            }
          }
        } else {
          {  // MFV_EPI_F32 (also serves MFV_EPI_ATOMIC_F32: p.epi selects the reduce-add store): raw fp32 tile
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(st0 + stage_off(lane, j)) =
                  make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            publish(mapC, st0, n0, row0, g, p.epi == MFV_EPI_ATOMIC_F32);
          }
        }
      }
      if constexpr (epi == MFV_EPI_RESID_LN) {
        // ---- row statistics: the four warps of this TMEM lane quarter own 96 columns each of the same 32 rows
        tmem_st_wait();
        ln_part[ew * 32 + lane] = make_float2(ln_m, ln_s);
        tc_fence_before();
        PROF_MARK(5)
        named_bar_sync(1u + (uint32_t)q, 128u);  // the four epilogue warps of TMEM lane quarter q: ew = (ew & 3) + 4 h
        tc_fence_after();
        PROF_MARK(6)
        float mean = 0.f, m2 = 0.f;
        float2 part[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          part[k] = ln_part[((ew & 3) + 4 * k) * 32 + lane];
          mean += part[k].x;
          m2 += part[k].y;
        }
        mean *= 0.25f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = part[k].x - mean; m2 = fmaf(96.0f * d, d, m2); }
        const float rstd = rsqrtf(m2 * (1.0f / 384.0f) + p.ln_eps);
        if (h == 0 && row0 + lane < p.M) {
          p.ln_mean[(long long)g * p.M + row0 + lane] = mean;
          p.ln_rstd[(long long)g * p.M + row0 + lane] = rstd;
        }
        const float* gam = par + 384;  // staged at the start of the tile (BN == N == 384: ncol0 == 0)
        const float* bet = par + 768;
        // ---- pass 2: x back out of TMEM, normalised, to C2 (and C3)
        if (!p.ln_out_f32) {
#pragma unroll 1
          for (int j = 0; j < 3; ++j) {  // 32-column pieces h, h + 4, h + 8 of the 12
            const int c2 = h + 4 * j;
            uint32_t v[32];
            PROF_MARK(5)
            tmem_ld32(trow + (uint32_t)(c2 * 32), v);
            tmem_ld_wait();
            PROF_MARK(1)
            if (j == 2) release_accumulator(as);
            float o[32];
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              const float4 g4 = *reinterpret_cast<const float4*>(gam + c2 * 32 + k);
              const float4 b4 = *reinterpret_cast<const float4*>(bet + c2 * 32 + k);
              o[k] = fmaf((__uint_as_float(v[k]) - mean) * rstd, g4.x, b4.x);
              o[k + 1] = fmaf((__uint_as_float(v[k + 1]) - mean) * rstd, g4.y, b4.y);
              o[k + 2] = fmaf((__uint_as_float(v[k + 2]) - mean) * rstd, g4.z, b4.z);
              o[k + 3] = fmaf((__uint_as_float(v[k + 3]) - mean) * rstd, g4.w, b4.w);
            }
            uint8_t* s2 = slot_ptr(use);
            acquire_slot();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              *reinterpret_cast<uint4*>(s2 + stage_off(lane, jj)) =
                  make_uint4(pack16(o[8 * jj], o[8 * jj + 1]), pack16(o[8 * jj + 2], o[8 * jj + 3]),
                             pack16(o[8 * jj + 4], o[8 * jj + 5]), pack16(o[8 * jj + 6], o[8 * jj + 7]));
            publish(&tmC2, s2, ncol0 + c2 * 32, row0, g, false);
            if (p.has_c3) {
              uint8_t* s3 = slot_ptr(use);
              acquire_slot();
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                *reinterpret_cast<uint4*>(s3 + stage_off(lane, jj)) =
                    make_uint4(pack_bf16(o[8 * jj], o[8 * jj + 1]), pack_bf16(o[8 * jj + 2], o[8 * jj + 3]),
                               pack_bf16(o[8 * jj + 4], o[8 * jj + 5]), pack_bf16(o[8 * jj + 6], o[8 * jj + 7]));
              publish(&tmC3, s3, ncol0 + c2 * 32, row0, g, false);
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < n_my; ++i) {  // fp32 output (final norm -> tokens): the 16-column pieces of pass 1
            const int c = h + 4 * i;
            uint32_t v[32];
            tmem_ld16(trow + (uint32_t)(c * 16), v);
            tmem_ld_wait();
            if (i == n_my - 1) release_accumulator(as);
            uint8_t* s2 = slot_ptr(use);
            acquire_slot();
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const float4 g4 = *reinterpret_cast<const float4*>(gam + c * 16 + k);
              const float4 b4 = *reinterpret_cast<const float4*>(bet + c * 16 + k);
              *reinterpret_cast<float4*>(s2 + stage_off(lane, k >> 2)) =
                  make_float4(fmaf((__uint_as_float(v[k]) - mean) * rstd, g4.x, b4.x),
                              fmaf((__uint_as_float(v[k + 1]) - mean) * rstd, g4.y, b4.y),
                              fmaf((__uint_as_float(v[k + 2]) - mean) * rstd, g4.z, b4.z),
                              fmaf((__uint_as_float(v[k + 3]) - mean) * rstd, g4.w, b4.w));
            }
            publish(&tmC2, s2, ncol0 + c * 16, row0, g, false);
          }
        }
      }
    }
    PROF_MARK(5)
    if (elect_one()) bulk_wait0();  // all global writes of this warp are complete before the CTA exits
    __syncwarp();
    PROF_MARK(7)
    if (prof && lane == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) prof[((size_t)blockIdx.x * (NUM_EPI_WARPS + 1) + ew) * 8 + k] = prof_c[k];
    }
#undef PROF_MARK
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();  // the peer's smem / TMEM stay alive until every MMA retired
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
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
  CUresult r = enc(map, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFV_OK : MFV_ERR_ARG;
}

// Row-major [G][rows][cols] tensor written / read by the epilogue: box = (64 B of columns, 32 rows, 1), 64B swizzle.
static int encode_tile_map(CUtensorMap* map, const void* base, int elem_bytes, int is_f16, long long rows,
                           long long cols, long long ld, long long gstride, int groups) {
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem_bytes,
                           (cuuint64_t)(groups > 1 ? gstride : rows * ld) * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)(64 / elem_bytes), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return MFV_ERR_ALIGN;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return MFV_ERR_INIT;
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFV_OK : MFV_ERR_ARG;
}

template <int BN, int CG, int EPI>
static int launch_gemm_epi(const mfv_gemm_args* a, cudaStream_t stream) {
  using S = GemmSmem<BN, CG, epi_nbuf(EPI), EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MFV_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, CG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::TOTAL));
    attr_set = true;
  }
  GemmParams p;
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K; p.G = (int)a->G;
  p.tiles_n = (p.N + BN - 1) / BN;
  // Rows per CTA.  The N = 384 GEMMs of a 32- or 64-pair step are one or two waves of 256 x 384 pair tiles that leave a
  // third of the SMs idle (50 tiles on 74 pairs).  With 192-row pair tiles each CTA stages / stores only 96 rows (TMEM
  // lanes 96..127 of the M = 256 UMMA compute rows of the next tile and are ignored): 66 tiles on 74 pairs, and the
  // isolated GEMMs get 2-9 % faster (tests/gpu_rows96_bench.py).  The step does not: in the backward the idle SMs are
  // where the side-stream weight-gradient GEMMs run (filling them cost 1.3 % of the step), and in the forward the gain
  // disappears behind PDL overlap.  So it stays opt-in: rows_per_cta = 96, or MFVIT_ROWS96=1 for the forward GEMMs.
  p.rows_cta = BM;
  if (BN == 384 && CG == 2 && !a->a_mn_major && a->epilogue != MFV_EPI_ATOMIC_F32) {
    const long long pairs = num_sms() / 2;
    const long long t128 = ((p.M + 255) / 256) * p.tiles_n * p.G, t96 = ((p.M + 191) / 192) * p.tiles_n * p.G;
    const bool fits = (t96 + pairs - 1) / pairs <= (t128 + pairs - 1) / pairs;
    // MFVIT_ROWS96: 1 = forward residual GEMMs (proj / fc2, with or without the fused LayerNorm), 2 = also the bf16 dgrads
    const int r96 = rows96_mode();
    const bool fwd_res = (EPI == MFV_EPI_RESID_F32 || EPI == MFV_EPI_RESID_LN);
    if (a->rows_per_cta == 96 || (a->rows_per_cta == 0 && fits && ((fwd_res && r96 >= 1) || (EPI == MFV_EPI_BF16 && r96 >= 2))))
      p.rows_cta = 96;
  }
  p.tiles_m = (p.M + p.rows_cta * CG - 1) / (p.rows_cta * CG);
  p.kb_total = (p.K + BK - 1) / BK;
  int splits = a->splits > 0 ? a->splits : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits > 1 && a->epilogue != MFV_EPI_ATOMIC_F32) return MFV_ERR_ARG;
  p.a_mn = a->a_mn_major; p.b_mn = a->b_mn_major; p.epi = a->epilogue;
  p.a_f16 = a->dtype_flags & 1; p.b_f16 = (a->dtype_flags >> 1) & 1; p.out_f16 = (a->dtype_flags >> 2) & 1;
  if (p.a_f16 != p.b_f16) return MFV_ERR_ARG;  // tcgen05 kind::f16 traps on mixed fp16 x bf16 operands
  p.has_c2 = a->C2 != nullptr;
  p.has_c3 = a->C3 != nullptr;
  p.dbg_skip_epilogue = (a->dtype_flags >> 8) & 15;  // bit0: skip everything, bit1: skip the bulk stores
  p.dbg_stages = 0;
  p.dbg_mma = (p.dbg_skip_epilogue & 1) ? ((a->dtype_flags >> 16) & 15) : 0;
  if ((p.dbg_skip_epilogue & 1) && ((a->dtype_flags >> 12) & 15)) {
    const int want = (a->dtype_flags >> 12) & 15;
    const int room = (S::STAGES * S::STAGE_BYTES + S::EPI_BYTES) / S::STAGE_BYTES;
    p.dbg_stages = want < 2 ? 2 : (want > room ? room : (want > MAX_STAGES ? MAX_STAGES : want));
  }
  p.bias_gstride = a->bias_gstride;
  p.bias = (const float*)a->bias;
  p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_mean = a->ln_mean; p.ln_rstd = a->ln_rstd;
  p.ln_eps = a->ln_eps; p.ln_out_f32 = a->ln_out_f32;
  if (EPI == MFV_EPI_RESID_LN) {
    if (BN != 384 || a->N != 384 || !a->ln_gamma || !a->ln_beta || !a->ln_mean || !a->ln_rstd || !a->C2 || !a->bias)
      return MFV_ERR_ARG;
    if (a->ln_out_f32 && a->C3) return MFV_ERR_ARG;
  }
  p.row_sum = nullptr;
  p.dbg_prof = nullptr;
  if (((a->dtype_flags >> 20) & 1) && a->row_sum && a->epilogue != MFV_EPI_ATOMIC_F32) {
    p.dbg_prof = reinterpret_cast<unsigned long long*>(a->row_sum);
  } else if (a->row_sum) {
    if (BN != 384 || a->epilogue != MFV_EPI_ATOMIC_F32 || a->N % 384 != 0) return MFV_ERR_ARG;
    p.row_sum = a->row_sum;
  }
  p.tiles0 = 0x7fffffff; p.M1 = p.N1 = p.tiles_m1 = p.tiles_n1 = 0; p.row_sum1 = nullptr;

  CUtensorMap tmA, tmB, tmC, tmC2, tmC3, tmAux;
  int rc = encode_operand_map(&tmA, a->A, a->a_mn_major, a->M, a->K, a->lda, a->a_gstride, p.G,
                              a->a_mn_major ? BM : p.rows_cta, p.a_f16);
  if (rc) return rc;
  rc = encode_operand_map(&tmB, a->B, a->b_mn_major, a->N, a->K, a->ldb, a->b_gstride, p.G, BN == 384 ? 64 : BN / CG,
                          p.b_f16);
  if (rc) return rc;
  const int e = a->epilogue;
  const bool c32 = (e == MFV_EPI_RESID_F32 || e == MFV_EPI_F32 || e == MFV_EPI_ATOMIC_F32 || e == MFV_EPI_RESID_LN);
  // C: u of the GELU epilogue is always bf16; other 16-bit outputs follow out_f16
  rc = encode_tile_map(&tmC, a->C, c32 ? 4 : 2, (e == MFV_EPI_BF16) ? p.out_f16 : 0, a->M, a->N, a->ldc, a->c_gstride,
                       p.G);
  if (rc) return rc;
  tmC2 = tmC; tmC3 = tmC; tmAux = tmC;
  if (e == MFV_EPI_DGELU && a->C2) {  // gelu(u) for the fc2 weight gradient, always bf16
    rc = encode_tile_map(&tmC2, a->C2, 2, 0, a->M, a->N, a->ldc, a->c_gstride, p.G);
    if (rc) return rc;
  }
  if (e == MFV_EPI_GELU) {
    rc = encode_tile_map(&tmC2, a->C2, 2, p.out_f16, a->M, a->N, a->ldc, a->c_gstride, p.G);
    if (rc) return rc;
    if (a->C3) {
      rc = encode_tile_map(&tmC3, a->C3, 2, 0, a->M, a->N, a->ldc, a->c_gstride, p.G);
      if (rc) return rc;
    }
  }
  if (e == MFV_EPI_RESID_LN) {  // C2 = LayerNorm(x): 16-bit (out_f16 selects fp16) or f32; C3 = optional bf16 copy
    rc = encode_tile_map(&tmC2, a->C2, a->ln_out_f32 ? 4 : 2, p.out_f16, a->M, a->N, a->ldc, a->c_gstride, p.G);
    if (rc) return rc;
    if (a->C3) {
      rc = encode_tile_map(&tmC3, a->C3, 2, 0, a->M, a->N, a->ldc, a->c_gstride, p.G);
      if (rc) return rc;
    }
  }
  if (e == MFV_EPI_RESID_F32 || e == MFV_EPI_DGELU || e == MFV_EPI_RESID_LN) {
    rc = encode_tile_map(&tmAux, a->aux, e == MFV_EPI_DGELU ? 2 : 4, 0, a->M, a->N, a->aux_ld, a->aux_gstride, p.G);
    if (rc) return rc;
  }

  // B multicast between two pairs (MC kernel): the 384-wide bf16 / residual / residual + LayerNorm GEMMs with K-major A
  constexpr bool mc_family = BN == 384 && CG == 2 && (EPI == MFV_EPI_BF16 || EPI == MFV_EPI_RESID_F32 || EPI == MFV_EPI_RESID_LN);
  bool use_mc = false;
  if constexpr (mc_family) {
    use_mc = gemm_multicast_enabled() && !a->a_mn_major && p.splits == 1 && p.rows_cta == BM && p.tiles_m >= 2 &&
             !p.dbg_skip_epilogue;
    if (use_mc) p.tiles_m = (p.tiles_m + 1) / 2;  // the schedule walks PAIRS of row tiles
  }
  const int total = p.tiles_m * p.tiles_n * p.splits * p.G;
  int clusters = num_sms() / CG;
  if (total < clusters) clusters = total;
  // Stream-K: where whole tiles leave pairs idle (fewer tiles than pairs, or a ragged last round), for the epilogues that
  // write each output element once, for the families MFVIT_STREAMK selects, and only when the reduction is long enough
  // (>= streamk_min_kb k-blocks per tile) for the shorter mainloop to pay for the partial dump + add.
  p.sk_units = 0; p.sk_ws = nullptr; p.sk_flags = nullptr;
  {
    const int mask = streamk_mask();
    const bool family = (BN == 384 && (mask & 1) && (EPI == MFV_EPI_BF16 || EPI == MFV_EPI_RESID_F32 || EPI == MFV_EPI_RESID_LN)) ||
                        (BN == 256 && (mask & 2) && EPI == MFV_EPI_BF16) || (BN == 256 && (mask & 4) && EPI == MFV_EPI_GELU) ||
                        (BN == 256 && (mask & 8) && EPI == MFV_EPI_DGELU);
    const int pairs = num_sms() / CG;
    if (CG == 2 && family && !use_mc && p.splits == 1 && a->epilogue != MFV_EPI_ATOMIC_F32 && total % pairs != 0 &&
        p.kb_total >= streamk_min_kb() && streamk_workspace()) {
      const long long units = (long long)total * p.kb_total;
      const int per = (int)((units + pairs - 1) / pairs);
      if (per >= 2) {
        p.sk_units = per;
        p.sk_ws = streamk_workspace();
        p.sk_flags = streamk_flags();
        clusters = (int)((units + per - 1) / per);  // every pair that has a range (<= pairs)
      }
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CG));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {  // programmatic dependent launch: prologue overlaps the previous kernel's tail (griddep_wait)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CG == 2) {  // single-CTA tiles are plain (non-cluster) launches
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if constexpr (mc_family) {
    if (use_mc) {
      static bool mc_attr_set = false;
      static int mc_max_clusters = 0;
      attr[na - 1].val.clusterDim.x = 4;  // the cluster attribute is the last one written above (CG == 2)
      if (!mc_attr_set) {
        MFV_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, CG, EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            S::TOTAL));
        cfg.gridDim = dim3((unsigned)(num_sms() / 4 * 4));
        MFV_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&mc_max_clusters, gemm_bf16_kernel<BN, CG, EPI, 1>, &cfg));
        if (mc_max_clusters < 1) mc_max_clusters = 1;
        mc_attr_set = true;
      }
      int cl4 = mc_max_clusters;  // clusters of four CTAs that can be resident at once (GPC boundaries cost a few SMs)
      if (cl4 > num_sms() / 4) cl4 = num_sms() / 4;
      if (total < cl4) cl4 = total;
      cfg.gridDim = dim3((unsigned)(cl4 * 4));
      MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, CG, EPI, 1>, tmA, tmB, tmC, tmC2, tmC3, tmAux, p));
      MFV_LAUNCH_CHECK();
      return MFV_OK;
    }
  }
  MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, CG, EPI>, tmA, tmB, tmC, tmC2, tmC3, tmAux, p));
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

template <int BN, int CG>
static int launch_gemm(const mfv_gemm_args* a, cudaStream_t stream) {
  switch (a->epilogue) {
    case MFV_EPI_BF16: return launch_gemm_epi<BN, CG, MFV_EPI_BF16>(a, stream);
    case MFV_EPI_GELU: return launch_gemm_epi<BN, CG, MFV_EPI_GELU>(a, stream);
    case MFV_EPI_RESID_F32: return launch_gemm_epi<BN, CG, MFV_EPI_RESID_F32>(a, stream);
    case MFV_EPI_DGELU: return launch_gemm_epi<BN, CG, MFV_EPI_DGELU>(a, stream);
    case MFV_EPI_RESID_LN:
      if constexpr (BN == 384) return launch_gemm_epi<BN, CG, MFV_EPI_RESID_LN>(a, stream);
      else return MFV_ERR_ARG;
    case MFV_EPI_F32:
    case MFV_EPI_ATOMIC_F32: return launch_gemm_epi<BN, CG, MFV_EPI_F32>(a, stream);
    default: return MFV_ERR_ARG;
  }
}

// Two weight-gradient GEMMs as ONE grid of 256 x 384 pair tiles (fp32 reduce-add epilogue): same reduction length K,
// split count, groups and operand formats; problem 1's operand / output maps ride in the tmC2 / tmC3 / tmAux slots.
static int launch_wgrad_pair(const mfv_gemm_args* a, const mfv_gemm_args* b, cudaStream_t stream) {
  constexpr int BN = 384, CG = 2;
  using S = GemmSmem<BN, CG, epi_nbuf(MFV_EPI_F32), MFV_EPI_F32>;
  static bool attr_set = false;
  if (!attr_set) {
    MFV_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, CG, MFV_EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::TOTAL));
    attr_set = true;
  }
  GemmParams p = {};
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K; p.G = (int)a->G;
  p.M1 = (int)b->M; p.N1 = (int)b->N;
  p.rows_cta = BM;
  p.tiles_n = p.N / BN; p.tiles_m = (p.M + BM * CG - 1) / (BM * CG);
  p.tiles_n1 = p.N1 / BN; p.tiles_m1 = (p.M1 + BM * CG - 1) / (BM * CG);
  p.kb_total = (p.K + BK - 1) / BK;
  int splits = a->splits > 0 ? a->splits : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.a_mn = a->a_mn_major; p.b_mn = a->b_mn_major; p.epi = MFV_EPI_ATOMIC_F32;
  p.a_f16 = a->dtype_flags & 1; p.b_f16 = (a->dtype_flags >> 1) & 1; p.out_f16 = 0;
  if (p.a_f16 != p.b_f16) return MFV_ERR_ARG;
  p.bias_gstride = a->bias_gstride;
  p.row_sum = a->row_sum; p.row_sum1 = b->row_sum;
  p.tiles0 = p.tiles_m * p.tiles_n * p.splits * p.G;
  CUtensorMap tmA, tmB, tmC, tmA1, tmB1, tmC1;
  int rc = encode_operand_map(&tmA, a->A, a->a_mn_major, a->M, a->K, a->lda, a->a_gstride, p.G, BM, p.a_f16);
  if (rc) return rc;
  rc = encode_operand_map(&tmB, a->B, a->b_mn_major, a->N, a->K, a->ldb, a->b_gstride, p.G, 64, p.b_f16);
  if (rc) return rc;
  rc = encode_tile_map(&tmC, a->C, 4, 0, a->M, a->N, a->ldc, a->c_gstride, p.G);
  if (rc) return rc;
  rc = encode_operand_map(&tmA1, b->A, b->a_mn_major, b->M, b->K, b->lda, b->a_gstride, p.G, BM, p.a_f16);
  if (rc) return rc;
  rc = encode_operand_map(&tmB1, b->B, b->b_mn_major, b->N, b->K, b->ldb, b->b_gstride, p.G, 64, p.b_f16);
  if (rc) return rc;
  rc = encode_tile_map(&tmC1, b->C, 4, 0, b->M, b->N, b->ldc, b->c_gstride, p.G);
  if (rc) return rc;
  const int total = p.tiles0 + p.tiles_m1 * p.tiles_n1 * p.splits * p.G;
  int clusters = num_sms() / CG;
  if (total < clusters) clusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CG));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  MFV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, CG, MFV_EPI_F32>, tmA, tmB, tmC, tmA1, tmB1, tmC1, p));
  MFV_LAUNCH_CHECK();
  return MFV_OK;
}

}  // namespace mfv

extern "C" int mfv_gemm(const mfv_gemm_args* a, void* stream) {
  using namespace mfv;
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0 || a->G <= 0) return MFV_ERR_SHAPE;
  if (a->N % 32 != 0) return MFV_ERR_SHAPE;
  if (a->ldc % 8 != 0) return MFV_ERR_ALIGN;
  if ((a->epilogue == MFV_EPI_RESID_F32 || a->epilogue == MFV_EPI_DGELU || a->epilogue == MFV_EPI_RESID_LN) && !a->aux)
    return MFV_ERR_ARG;
  if (a->epilogue == MFV_EPI_GELU && !a->C2) return MFV_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // Tile selection (0 = auto).  The widest tile that fits wins (fewest operand re-reads, fewest epilogue pieces): CTA pairs
  // (256 rows) whenever M allows, 384 columns when N is exactly 384 (A read once), else 256 columns for N >= 512.
  int bn = a->block_n;
  int cg = a->cta_group;
  const bool pair_ok = a->M > 128 && cg != 1;
  if (a->epilogue == MFV_EPI_RESID_LN) {  // whole rows in one tile: 384-wide pair tiles only
    if (a->N != 384 || a->M <= 128) return MFV_ERR_SHAPE;
    return launch_gemm<384, 2>(a, s);
  }
  if (bn == 0) {
    if (a->N == 384 && pair_ok && a->epilogue != MFV_EPI_GELU) bn = 384;
    else if (a->N >= 512 && pair_ok) bn = 256;
    else {
      const long long tm = (a->M + BM - 1) / BM;
      const long long sp = a->splits > 0 ? a->splits : 1;
      bn = 128;
      if (tm * ((a->N + 127) / 128) * a->G * sp < num_sms() && a->N % 64 == 0) bn = 64;
    }
  }
  if (cg == 0) cg = pair_ok ? 2 : 1;
  if (bn == 384) {
    if (cg != 2 || a->N % 384 != 0) return MFV_ERR_ARG;
    return launch_gemm<384, 2>(a, s);
  }
  if (cg == 2) {
    switch (bn) {
      case 128: return launch_gemm<128, 2>(a, s);
      case 256: return launch_gemm<256, 2>(a, s);
      default: cg = 1; break;
    }
  }
  switch (bn) {
    case 64: return launch_gemm<64, 1>(a, s);
    case 128: return launch_gemm<128, 1>(a, s);
    case 256: return launch_gemm<256, 1>(a, s);
    default: return MFV_ERR_ARG;
  }
}

// Two split-K weight-gradient GEMMs (MFV_EPI_ATOMIC_F32) in ONE launch: the four weight gradients of a block are two
// launches instead of four, each long enough to fill the machine without slicing the reduction into short pieces.  Both
// problems must have N % 384 == 0 (256 x 384 pair tiles), M > 128, the same K, G, splits, operand majorness and formats,
// and the same bias_gstride for their row_sum outputs.
extern "C" int mfv_gemm_wgrad_pair(const mfv_gemm_args* a, const mfv_gemm_args* b, void* stream) {
  using namespace mfv;
  if (!a || !b) return MFV_ERR_ARG;
  for (const mfv_gemm_args* x : {a, b}) {
    if (x->M <= 128 || x->N <= 0 || x->N % 384 != 0 || x->K <= 0 || x->G <= 0) return MFV_ERR_SHAPE;
    if (x->epilogue != MFV_EPI_ATOMIC_F32 || x->ldc % 8 != 0) return MFV_ERR_ARG;
  }
  if (a->K != b->K || a->G != b->G || a->splits != b->splits || a->a_mn_major != b->a_mn_major ||
      a->b_mn_major != b->b_mn_major || (a->dtype_flags & 3) != (b->dtype_flags & 3) || a->bias_gstride != b->bias_gstride)
    return MFV_ERR_ARG;
  return launch_wgrad_pair(a, b, reinterpret_cast<cudaStream_t>(stream));
}
