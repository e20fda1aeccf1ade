// Native runtime for the ViT-S/16 encoder: one call enqueues the whole forward (or backward) of G branches on a
// stream.  No kernels here - this file is the host-side executor that sequences gemm.cu / ln.cu / attn.cu /
// elementwise.cu launches over caller-owned buffers (see mfv_vit_plan in include/mfvit.h).
//
// Forward per block l (residual stream x kept in fp32, GEMM operands bf16):
//   xn1 = LN(x_in)            -> qkv = xn1 Wqkv^T + b        -> o = softmax(q k^T / sqrt(d)) v
//   x_mid = x_in + o Wp^T + b -> xn2 = LN(x_mid)             -> u = xn2 W1^T + b, g = gelu(u)
//   x_out = x_mid + g W2^T + b
// Backward mirrors it: for every Linear a split-K wgrad (A,B MN-major), a dgrad (B MN-major), a bias column-sum; the
// LN backward kernels carry the fp32 residual-stream gradient and emit the bf16 copy the next GEMMs consume.
#include "common.cuh"
#include "mfvit_internal.h"

#include <stdlib.h>

namespace mfv {

struct PlanView {
  const mfv_vit_plan* p;
  long long M, MC, Mh, M3;
  explicit PlanView(const mfv_vit_plan* pl) : p(pl) {
    M = pl->B * pl->S;
    MC = pl->G * M * pl->C;
    Mh = pl->G * M * pl->hidden;
    M3 = pl->G * M * 3 * pl->C;
  }
  int slot(int i) const { return p->save_for_backward ? i : 0; }
  float* x(int i) const { return p->x + (long long)(p->save_for_backward ? i : (i & 1)) * MC; }
  float* mean(int i) const { return p->stats + (long long)slot(i) * 2 * p->G * M; }
  float* rstd(int i) const { return mean(i) + p->G * M; }
  __nv_bfloat16* xn(int i) const { return reinterpret_cast<__nv_bfloat16*>(p->xn) + (long long)slot(i) * MC; }
  __nv_bfloat16* qkv(int l) const { return reinterpret_cast<__nv_bfloat16*>(p->qkv) + (long long)slot(l) * M3; }
  __nv_bfloat16* ao(int l) const { return reinterpret_cast<__nv_bfloat16*>(p->attn_o) + (long long)slot(l) * MC; }
  float* lse(int l) const { return p->lse + (long long)slot(l) * p->G * p->B * p->H * p->S; }
  __nv_bfloat16* u(int l) const { return reinterpret_cast<__nv_bfloat16*>(p->u) + (long long)slot(l) * Mh; }
  __nv_bfloat16* g(int l) const { return reinterpret_cast<__nv_bfloat16*>(p->gact) + (long long)slot(l) * Mh; }
  long long boff(int l, long long rel) const { return p->off_block0 + (long long)l * p->block_stride + rel; }
  const float* w32(long long off) const { return p->master + off; }
  const __nv_bfloat16* w16(long long off) const { return reinterpret_cast<const __nv_bfloat16*>(p->shadow) + off; }
  // forward GEMM weight operand: fp16 shadow in fwd_f16 mode
  const __nv_bfloat16* wf(long long off) const {
    return reinterpret_cast<const __nv_bfloat16*>(p->fwd_f16 ? p->shadow16 : p->shadow) + off;
  }
  bool dual() const { return p->fwd_f16 && p->save_for_backward; }
  // bf16 copies for the backward GEMMs (same slot indexing); alias the forward buffers in pure-bf16 mode
  __nv_bfloat16* xn_b(int i) const {
    return p->fwd_f16 ? reinterpret_cast<__nv_bfloat16*>(p->xn_bf) + (long long)slot(i) * MC : xn(i);
  }
  __nv_bfloat16* ao_b(int l) const {
    return p->fwd_f16 ? reinterpret_cast<__nv_bfloat16*>(p->attn_o_bf) + (long long)slot(l) * MC : ao(l);
  }
  // fp16-forward mode: the bf16 GELU output for the fc2 weight gradient is either stored per block by the fc1 epilogue
  // (plan.gact_bf_per_block, the default of mfvit.engine: 4.543 -> 4.528 ms per step for 27 MB per pair) or recomputed by
  // the fc2-dgrad epilogue from u (which it reads anyway) into one reusable [G][M][hidden] buffer just before the wgrad
  __nv_bfloat16* g_b(int l) const {
    if (!p->fwd_f16) return g(l);
    return reinterpret_cast<__nv_bfloat16*>(p->gact_bf) + (p->gact_bf_per_block ? (long long)slot(l) * Mh : 0);
  }
  const void* patches_b() const { return p->fwd_f16 ? p->patches_bf : p->patches; }
  float* gr(long long off) const { return p->grad + off; }
};

// y = x W^T (+bias): A = activations [G][M][K] bf16 K-major, B = weight [N][K] bf16 K-major
static int linear_fwd(const PlanView& v, const void* A, long long K, long long w_off, long long b_off, long long N,
                      int epi, void* C, void* C2, void* C3, const void* aux, long long aux_ld, cudaStream_t st) {
  mfv_gemm_args a = {};
  a.A = A; a.B = v.wf(w_off); a.C = C; a.C2 = C2; a.C3 = C3;
  // fp16 operands in fwd_f16 mode; 16-bit outputs consumed by the forward (GELU out, qkv) are fp16 too
  a.dtype_flags = v.p->fwd_f16 ? (3 | 4) : 0;
  a.bias = b_off >= 0 ? v.w32(b_off) : nullptr;
  a.aux = aux;
  a.M = v.M; a.N = N; a.K = K; a.G = v.p->G;
  a.lda = K; a.ldb = K; a.ldc = N;
  a.a_gstride = v.M * K; a.b_gstride = v.p->P; a.c_gstride = v.M * N;
  a.aux_ld = aux_ld; a.aux_gstride = v.M * aux_ld; a.bias_gstride = v.p->P;
  a.epilogue = epi;
  ProfScope ps(PROF_GEMM_FWD, st);
  return mfv_gemm(&a, st);
}
// x_new = x_old + y W^T + bias AND the LayerNorm of x_new in the same epilogue (MFV_EPI_RESID_LN, C == 384): the operand
// of the next Linear (xn, + bf16 copy), or the fp32 tokens when this is the last Linear in front of the final norm
static int linear_fwd_ln(const PlanView& v, const void* A, long long K, long long w_off, long long b_off, float* x_new,
                         const float* x_old, long long ln_w_off, long long ln_b_off, int stat_idx, void* xn16,
                         void* xn_copy, float* out_f32, cudaStream_t st) {
  const long long C = v.p->C;
  mfv_gemm_args a = {};
  a.A = A; a.B = v.wf(w_off); a.C = x_new; a.aux = x_old;
  a.C2 = out_f32 ? (void*)out_f32 : xn16; a.C3 = out_f32 ? nullptr : xn_copy;
  a.dtype_flags = v.p->fwd_f16 ? (3 | 4) : 0;
  a.bias = v.w32(b_off);
  a.M = v.M; a.N = C; a.K = K; a.G = v.p->G;
  a.lda = K; a.ldb = K; a.ldc = C;
  a.a_gstride = v.M * K; a.b_gstride = v.p->P; a.c_gstride = v.M * C;
  a.aux_ld = C; a.aux_gstride = v.M * C; a.bias_gstride = v.p->P;
  a.epilogue = MFV_EPI_RESID_LN;
  a.ln_gamma = v.w32(ln_w_off); a.ln_beta = v.w32(ln_b_off);
  a.ln_mean = v.mean(stat_idx); a.ln_rstd = v.rstd(stat_idx);
  a.ln_eps = 1e-6f; a.ln_out_f32 = out_f32 ? 1 : 0;
  ProfScope ps(PROF_GEMM_FWD, st);
  return mfv_gemm(&a, st);
}
// dx = dy W: A = dy [G][M][N] K-major (reduction over N), B = W [N][K] read MN-major
static int linear_dgrad(const PlanView& v, const void* dY, long long N, long long w_off, long long K, int epi, void* C,
                        void* C2, const void* aux, long long aux_ld, cudaStream_t st) {
  mfv_gemm_args a = {};
  a.A = dY; a.B = v.w16(w_off); a.C = C; a.C2 = C2; a.aux = aux;
  a.M = v.M; a.N = K; a.K = N; a.G = v.p->G;
  a.lda = N; a.ldb = K; a.ldc = K;
  a.a_gstride = v.M * N; a.b_gstride = v.p->P; a.c_gstride = v.M * K;
  a.aux_ld = aux_ld; a.aux_gstride = v.M * aux_ld;
  a.b_mn_major = 1;
  a.epilogue = epi;
  ProfScope ps(PROF_GEMM_DGRAD, st);
  return mfv_gemm(&a, st);
}
// dW[N][K] += dy^T x (reduction over `rows` tokens), db[N] += colsum(dy)
static int linear_wgrad(const PlanView& v, const void* dY, long long N, const void* X, long long K, long long rows,
                        long long w_off, long long b_off, cudaStream_t st) {
  mfv_gemm_args a = {};
  a.A = dY; a.B = X; a.C = v.gr(w_off);
  a.M = N; a.N = K; a.K = rows; a.G = v.p->G;
  a.lda = N; a.ldb = K; a.ldc = K;
  a.a_gstride = rows * N; a.b_gstride = rows * K; a.c_gstride = v.p->P;
  a.a_mn_major = 1; a.b_mn_major = 1;
  a.epilogue = MFV_EPI_ATOMIC_F32;
  // tile = what mfv_gemm's auto selection will pick: 256 x 384 pair tiles when the input width is 384 (qkv / fc1 /
  // proj), else 256 x 256 (fc2: 384 x 1536 output)
  const long long bn = (K == 384) ? 384 : (K >= 512 ? 256 : 128);
  const long long bm = (N > 128) ? 256 : 128;
  // split-K: choose the split count whose tile count fills whole waves of the persistent grid best, while keeping
  // at least 8 k-blocks (512 tokens) per split so the fp32 reduce-add traffic stays small against the mainloop
  const long long tiles = ((N + bm - 1) / bm) * ((K + bn - 1) / bn) * v.p->G;
  const long long kb = (rows + 63) / 64;
  const long long sms = num_sms() / (bm == 256 ? 2 : 1);  // schedulable units: CTA pairs or single CTAs
  long long splits = 1;
  double best = -1.0;
  for (long long sp = 1; sp <= 32 && sp <= kb; ++sp) {
    if (kb / sp < 8 && sp > 1) break;
    const long long per = (kb + sp - 1) / sp;
    const long long eff_sp = (kb + per - 1) / per;
    const long long units = tiles * eff_sp;
    const long long waves = (units + sms - 1) / sms;
    const double eff = (double)units / (double)(waves * sms) - 0.002 * (double)sp;
    if (waves <= 3 && eff > best) { best = eff; splits = sp; }
  }
  a.splits = (int)splits;
  // bias gradient: folded into the 384-wide pair tile as one extra N=16 UMMA against a tile of ones (no extra pass
  // over dY); other tilings fall back to the column-sum kernel
  const bool fold_bias = b_off >= 0 && bn == 384 && bm == 256;
  if (fold_bias) {
    a.row_sum = v.gr(b_off);
    a.bias_gstride = v.p->P;
  }
  int rc;
  {
    ProfScope ps(PROF_GEMM_WGRAD, st);
    rc = mfv_gemm(&a, st);
  }
  if (rc) return rc;
  if (b_off >= 0 && !fold_bias) {
    ProfScope ps(PROF_COLSUM, st);
    rc = mfv_colsum_bf16(dY, v.gr(b_off), v.p->G, rows, N, v.p->P, st);
  }
  return rc;
}

// Two weight gradients that become ready together (fc2 + fc1 after the fc2 dgrad, proj + qkv after the attention backward)
// as ONE launch of 256 x 384 pair tiles (mfv_gemm_wgrad_pair): half the launches, and enough tiles per launch that the
// split count - chosen here so the units fill whole waves of the 74 CTA pairs - leaves long reduction slices.
// Measured at 32 pairs: the weight-gradient class time drops from 1.24 to 0.93 ms per step, but the STEP gets slower
// (4.577 -> 4.631 ms): the four short kernels slip into the SMs the 100-CTA dgrad kernels leave idle, the two long ones
// hold all 148 SMs and push the critical chain back.  Opt-in therefore (MFVIT_WGRAD_PAIR=1).
struct WgradSpec { const void* dY; long long N; const void* X; long long K; long long w_off, b_off; };
static bool wgrad_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MFVIT_WGRAD_PAIR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
static int linear_wgrad_pair(const PlanView& v, const WgradSpec& s0, const WgradSpec& s1, long long rows, cudaStream_t st) {
  const bool ok = wgrad_pair_enabled() && s0.K % 384 == 0 && s1.K % 384 == 0 && s0.N > 128 && s1.N > 128;
  if (!ok) {
    int rc = linear_wgrad(v, s0.dY, s0.N, s0.X, s0.K, rows, s0.w_off, s0.b_off, st);
    if (rc) return rc;
    return linear_wgrad(v, s1.dY, s1.N, s1.X, s1.K, rows, s1.w_off, s1.b_off, st);
  }
  mfv_gemm_args a[2] = {};
  long long tiles = 0;
  const WgradSpec* sp[2] = {&s0, &s1};
  for (int i = 0; i < 2; ++i) {
    const WgradSpec& s = *sp[i];
    a[i].A = s.dY; a[i].B = s.X; a[i].C = v.gr(s.w_off);
    a[i].M = s.N; a[i].N = s.K; a[i].K = rows; a[i].G = v.p->G;
    a[i].lda = s.N; a[i].ldb = s.K; a[i].ldc = s.K;
    a[i].a_gstride = rows * s.N; a[i].b_gstride = rows * s.K; a[i].c_gstride = v.p->P;
    a[i].a_mn_major = 1; a[i].b_mn_major = 1;
    a[i].epilogue = MFV_EPI_ATOMIC_F32;
    a[i].bias_gstride = v.p->P;
    if (s.b_off >= 0) a[i].row_sum = v.gr(s.b_off);  // bias gradient folded into the tile (one N=16 UMMA against ones)
    tiles += ((s.N + 255) / 256) * (s.K / 384) * v.p->G;
  }
  const long long kb = (rows + 63) / 64;
  const long long pairs = num_sms() / 2;
  long long splits = 1;
  double best = -1.0;
  for (long long spl = 1; spl <= 32 && spl <= kb; ++spl) {
    if (kb / spl < 8 && spl > 1) break;
    const long long per = (kb + spl - 1) / spl;
    const long long eff_sp = (kb + per - 1) / per;
    const long long units = tiles * eff_sp;
    const long long waves = (units + pairs - 1) / pairs;
    const double eff = (double)units / (double)(waves * pairs) - 0.002 * (double)spl;
    if (waves <= 3 && eff > best) { best = eff; splits = spl; }
  }
  a[0].splits = a[1].splits = (int)splits;
  ProfScope ps(PROF_GEMM_WGRAD, st);
  return mfv_gemm_wgrad_pair(&a[0], &a[1], st);
}

#define RC(expr)            \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)
// same, with the launch(es) inside timed under a profiling label
#define RCP(label, expr)              \
  do {                                \
    ProfScope _ps(label, st);         \
    int _rc = (expr);                 \
    if (_rc) return _rc;              \
  } while (0)

static int check_plan(const mfv_vit_plan* p) {
  if (!p) return MFV_ERR_ARG;
  if (p->G < 1 || p->G > 2 || p->B < 1 || p->depth < 1) return MFV_ERR_SHAPE;
  if (p->C % 128 || p->hidden % 32 || p->C % p->H) return MFV_ERR_SHAPE;
  if (p->S != p->np + 1 || p->np != (p->img / 16) * (p->img / 16)) return MFV_ERR_SHAPE;
  const long long d = p->C / p->H;
  if (d != 64 && d != 32) return MFV_ERR_SHAPE;
  if (p->P % 8) return MFV_ERR_ALIGN;
  if (p->fwd_f16 && !p->shadow16) return MFV_ERR_ARG;
  if (p->fwd_f16 && p->save_for_backward && (!p->patches_bf || !p->xn_bf || !p->attn_o_bf || !p->gact_bf))
    return MFV_ERR_ARG;
  return MFV_OK;
}

}  // namespace mfv

extern "C" int mfv_vit_forward(const mfv_vit_plan* p, void* stream) {
  using namespace mfv;
  RC(check_plan(p));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const PlanView v(p);
  const long long G = p->G, C = p->C, Hd = p->hidden, M = v.M;
  const long long D = C / p->H;
  const float scale = 1.0f / sqrtf((float)D);
  const long long rows_pe = p->B * p->np;
  // patch embedding.  Default (C == 384): ONE im2col-free GEMM - TMA tiles straight from the NCHW fp32 images, bias +
  // position embedding + class token in the epilogue (patch_embed.cu).  Otherwise patchify (per group: the images are
  // separate caller tensors) -> GEMM(+bias) -> +cls, +pos.
  const bool pe_tma = patch_tma_enabled() && C == 384 && p->img / 16 <= 128;
  if (pe_tma) {
    RCP(PROF_GEMM_FWD, mfv_patch_embed_tma(p->images[0], G > 1 ? p->images[1] : nullptr, v.wf(p->off_pe_w), p->fwd_f16,
                                           v.w32(p->off_pe_b), v.w32(p->off_cls), v.w32(p->off_pos), v.x(0), G, p->B, p->img,
                                           C, p->P, st));
  } else {
    for (int g = 0; g < G; ++g)
      RCP(PROF_PATCHIFY, mfv_patchify(p->images[g], reinterpret_cast<__nv_bfloat16*>(p->patches) + (long long)g * rows_pe * 768,
                      p->fwd_f16,
                      v.dual() ? reinterpret_cast<__nv_bfloat16*>(p->patches_bf) + (long long)g * rows_pe * 768 : nullptr,
                      p->B, p->img, st));
    {
      mfv_gemm_args a = {};
      a.A = p->patches; a.B = v.wf(p->off_pe_w); a.C = p->acc; a.bias = v.w32(p->off_pe_b);
      a.dtype_flags = p->fwd_f16 ? 3 : 0;
      a.M = rows_pe; a.N = C; a.K = 768; a.G = G;
      a.lda = 768; a.ldb = 768; a.ldc = C;
      a.a_gstride = rows_pe * 768; a.b_gstride = p->P; a.c_gstride = rows_pe * C; a.bias_gstride = p->P;
      a.epilogue = MFV_EPI_F32;
      RCP(PROF_GEMM_FWD, mfv_gemm(&a, st));
    }
    RCP(PROF_EMBED, mfv_embed_finish(p->acc, nullptr, v.w32(p->off_cls), v.w32(p->off_pos), v.x(0), G, p->B, p->np, C, p->P, st));
  }

  // LayerNorm placement: with C == 384 the proj / fc2 GEMMs own whole rows (256 x 384 pair tiles) and normalise what
  // they just wrote (MFV_EPI_RESID_LN) - only the first LayerNorm of block 0 (input: the embedding) is a launch of its
  // own.  Other widths (and MFVIT_FUSE_LN=0) keep the standalone kernels.
  const bool fuse_ln = (C == 384) && M > 128 && fuse_ln_enabled();
  const int last = 2 * (int)p->depth;
  for (int l = 0; l < p->depth; ++l) {
    float* x_in = v.x(2 * l);
    float* x_mid = v.x(2 * l + 1);
    float* x_out = v.x(2 * l + 2);
    const bool dual = v.dual();
    const bool final_block = (l == (int)p->depth - 1);
    if (l == 0 || !fuse_ln)
      RCP(PROF_LN_FWD, mfv_layernorm_fwd(x_in, v.w32(v.boff(l, p->r_ln1_w)), v.w32(v.boff(l, p->r_ln1_b)), v.xn(2 * l), p->fwd_f16,
                           dual ? v.xn_b(2 * l) : nullptr, nullptr, v.mean(2 * l), v.rstd(2 * l), G, M, C, p->P, 1e-6f,
                           st));
    RC(linear_fwd(v, v.xn(2 * l), C, v.boff(l, p->r_qkv_w), v.boff(l, p->r_qkv_b), 3 * C, MFV_EPI_BF16, v.qkv(l),
                  nullptr, nullptr, nullptr, 0, st));
    RCP(PROF_ATTN_FWD, mfv_attn_fwd(v.qkv(l), p->fwd_f16, v.ao(l), p->fwd_f16, dual ? v.ao_b(l) : nullptr, v.lse(l), G * p->B, p->S, p->H, D, scale,
                    st));
    if (fuse_ln) {
      RC(linear_fwd_ln(v, v.ao(l), C, v.boff(l, p->r_proj_w), v.boff(l, p->r_proj_b), x_mid, x_in,
                       v.boff(l, p->r_ln2_w), v.boff(l, p->r_ln2_b), 2 * l + 1, v.xn(2 * l + 1),
                       dual ? v.xn_b(2 * l + 1) : nullptr, nullptr, st));
    } else {
      RC(linear_fwd(v, v.ao(l), C, v.boff(l, p->r_proj_w), v.boff(l, p->r_proj_b), C, MFV_EPI_RESID_F32, x_mid, nullptr,
                    nullptr, x_in, C, st));
      RCP(PROF_LN_FWD, mfv_layernorm_fwd(x_mid, v.w32(v.boff(l, p->r_ln2_w)), v.w32(v.boff(l, p->r_ln2_b)), v.xn(2 * l + 1), p->fwd_f16,
                           dual ? v.xn_b(2 * l + 1) : nullptr, nullptr, v.mean(2 * l + 1), v.rstd(2 * l + 1), G, M, C,
                           p->P, 1e-6f, st));
    }
    RC(linear_fwd(v, v.xn(2 * l + 1), C, v.boff(l, p->r_fc1_w), v.boff(l, p->r_fc1_b), Hd, MFV_EPI_GELU, v.u(l),
                  v.g(l), (dual && p->gact_bf_per_block) ? v.g_b(l) : nullptr, nullptr, 0, st));
    if (fuse_ln && !final_block) {  // fc2 also writes LN1 of the next block
      RC(linear_fwd_ln(v, v.g(l), Hd, v.boff(l, p->r_fc2_w), v.boff(l, p->r_fc2_b), x_out, x_mid,
                       v.boff(l + 1, p->r_ln1_w), v.boff(l + 1, p->r_ln1_b), 2 * l + 2, v.xn(2 * l + 2),
                       dual ? v.xn_b(2 * l + 2) : nullptr, nullptr, st));
    } else if (fuse_ln) {           // ... or the final norm: fp32 tokens
      RC(linear_fwd_ln(v, v.g(l), Hd, v.boff(l, p->r_fc2_w), v.boff(l, p->r_fc2_b), x_out, x_mid, p->off_norm_w,
                       p->off_norm_b, last, nullptr, nullptr, p->tokens, st));
    } else {
      RC(linear_fwd(v, v.g(l), Hd, v.boff(l, p->r_fc2_w), v.boff(l, p->r_fc2_b), C, MFV_EPI_RESID_F32, x_out, nullptr,
                    nullptr, x_mid, C, st));
    }
  }
  if (!fuse_ln)
    RCP(PROF_LN_FWD, mfv_layernorm_fwd(v.x(last), v.w32(p->off_norm_w), v.w32(p->off_norm_b), nullptr, 0, nullptr, p->tokens,
                         v.mean(last), v.rstd(last), G, M, C, p->P, 1e-6f, st));
  return MFV_OK;
}

extern "C" int mfv_vit_backward(const mfv_vit_plan* p, void* stream) {
  if (!p) return MFV_ERR_ARG;
  return mfv_vit_backward_range(p, stream, (int)p->depth - 1, 0, MFV_BWD_HEAD | MFV_BWD_TAIL);
}

extern "C" int mfv_vit_backward_range(const mfv_vit_plan* p, void* stream, int block_hi, int block_lo, int flags) {
  using namespace mfv;
  RC(check_plan(p));
  if (!p->save_for_backward || !p->grad || !p->dtokens) return MFV_ERR_ARG;
  if (block_lo < 0 || block_hi >= p->depth || block_lo > block_hi) return MFV_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const PlanView v(p);
  const long long G = p->G, C = p->C, Hd = p->hidden, M = v.M;
  const long long D = C / p->H;
  const float scale = 1.0f / sqrtf((float)D);
  const int last = 2 * (int)p->depth;
  // Two lanes: `st` carries the chain every later kernel depends on (LayerNorm backward, dgrads, attention backward);
  // `sw` carries the weight-gradient GEMMs + bias column sums, which only feed the optimizer.  fork[h] marks their
  // inputs ready, done[h] marks the reusable buffers they read (dhid, gact_bf, dqkv, dx16) free again.
  SideStream* ss = side_stream();
  cudaStream_t sw = ss ? ss->stream : st;
  bool pending[2] = {false, false};  // done[h] recorded and not yet waited on by the main lane
  auto fork = [&](int h) -> int {
    if (!ss) return MFV_OK;
    MFV_CUDA_CHECK(cudaEventRecord(ss->fork[h], st));
    MFV_CUDA_CHECK(cudaStreamWaitEvent(sw, ss->fork[h], 0));
    return MFV_OK;
  };
  auto side_done = [&](int h) -> int {
    if (!ss) return MFV_OK;
    MFV_CUDA_CHECK(cudaEventRecord(ss->done[h], sw));
    pending[h] = true;
    return MFV_OK;
  };
  auto join = [&](int h) -> int {
    if (!ss || !pending[h]) return MFV_OK;
    MFV_CUDA_CHECK(cudaStreamWaitEvent(st, ss->done[h], 0));
    pending[h] = false;
    return MFV_OK;
  };
  int cur = 0;  // dx[cur] / dx16[cur] hold the gradient of the residual stream (always 0 at a block boundary: two flips per block)
  // Residual-gradient stream.  Default: bf16 only - the copy the dgrad / wgrad GEMMs read anyway is also what the next
  // LayerNorm backward adds its result to, so the fp32 read + write of the stream (44 % of the LayerNorm-backward
  // traffic) disappears; 24 roundings to bf16 along the depth cost ~0.5 % relative error on the earliest gradients
  // (cosine 0.9999+).  MFVIT_DX32=1 keeps the fp32 stream of round 1.  The very last LayerNorm backward (block 0) still
  // writes fp32: the embedding backward reads it.
  const bool dx32 = dx32_stream_enabled();
  auto res32 = [&](int c) -> const float* { return dx32 ? p->dx[c] : nullptr; };
  auto res16 = [&](int c) -> const void* { return dx32 ? nullptr : p->dx16[c]; };
  // final norm
  // each LN backward also emits colsum(dx) = bias gradient of the Linear feeding that residual add (fc2 / proj)
  // The bf16 patch matrix the conv weight-gradient GEMM reads (the im2col-free forward never wrote one) depends on the
  // input images only: it is built on the weight-gradient lane at the START of the backward, not in front of that GEMM
  // at its end, where it sat on the chain the optimizer waits for (2 x 9 us at 32 pairs).  Every later launch on `sw`
  // and the joins in front of the tail order it before the GEMM.
  const bool patch_rebuild = !p->stop_grad_conv1 && patch_tma_enabled() && C == 384 && p->img / 16 <= 128;
  if ((flags & MFV_BWD_HEAD) && patch_rebuild && ss) {
    RC(fork(1));
    const long long rows_pe0 = p->B * p->np;
    for (int g = 0; g < G; ++g) {
      if (!p->images[g]) return MFV_ERR_ARG;
      RCP(PROF_PATCHIFY, mfv_patchify(p->images[g],
                                      reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(v.patches_b())) + (long long)g * rows_pe0 * 768,
                                      0, nullptr, p->B, p->img, sw));
    }
    RC(side_done(1));
  }
  if (flags & MFV_BWD_HEAD)
    RCP(PROF_LN_BWD, mfv_layernorm_bwd(nullptr, p->dtokens, nullptr, nullptr, v.x(last), v.mean(last), v.rstd(last), v.w32(p->off_norm_w),
                         dx32 ? p->dx[cur] : nullptr, p->dx16[cur], v.gr(p->off_norm_w), v.gr(p->off_norm_b),
                         v.gr(v.boff((int)p->depth - 1, p->r_fc2_b)), G, M, C, p->P, st));
  for (int l = block_hi; l >= block_lo; --l) {
    // ---- MLP half: x_out = x_mid + fc2(gelu(fc1(LN2(x_mid))))
    RC(join(0));  // the previous block's MLP-half weight gradients still read dhid / gact_bf
    RC(linear_dgrad(v, p->dx16[cur], C, v.boff(l, p->r_fc2_w), Hd, MFV_EPI_DGELU, p->dhid,
                    (p->fwd_f16 && !p->gact_bf_per_block) ? v.g_b(l) : nullptr, v.u(l), Hd, st));
    RC(fork(0));
    {  // fc2 (bias: LN backward above) + fc1 weight gradients, one launch
      const WgradSpec w_fc2 = {p->dx16[cur], C, v.g_b(l), Hd, v.boff(l, p->r_fc2_w), -1};
      const WgradSpec w_fc1 = {p->dhid, Hd, v.xn_b(2 * l + 1), C, v.boff(l, p->r_fc1_w), v.boff(l, p->r_fc1_b)};
      RC(linear_wgrad_pair(v, w_fc2, w_fc1, M, sw));
    }
    RC(side_done(0));
    RC(linear_dgrad(v, p->dhid, Hd, v.boff(l, p->r_fc1_w), C, MFV_EPI_BF16, p->dxn, nullptr, nullptr, 0, st));
    RC(join(1));  // the previous block's attention-half weight gradients still read dx16[cur ^ 1]
    RCP(PROF_LN_BWD, mfv_layernorm_bwd(p->dxn, nullptr, res32(cur), res16(cur), v.x(2 * l + 1), v.mean(2 * l + 1), v.rstd(2 * l + 1),
                         v.w32(v.boff(l, p->r_ln2_w)), dx32 ? p->dx[cur ^ 1] : nullptr, p->dx16[cur ^ 1], v.gr(v.boff(l, p->r_ln2_w)),
                         v.gr(v.boff(l, p->r_ln2_b)), v.gr(v.boff(l, p->r_proj_b)), G, M, C, p->P, st));
    cur ^= 1;
    // ---- attention half: x_mid = x_in + proj(attn(qkv(LN1(x_in))))
    RC(linear_dgrad(v, p->dx16[cur], C, v.boff(l, p->r_proj_w), C, MFV_EPI_BF16, p->d_o, nullptr, nullptr, 0, st));
    RCP(PROF_ATTN_BWD, mfv_attn_bwd_ws(v.qkv(l), p->fwd_f16, v.ao_b(l), p->d_o, v.lse(l), p->delta, p->dqkv, p->attn_ws,
                                       G * p->B, p->S, p->H, D, scale, st));
    RC(fork(1));
    {  // proj (bias: LN2 backward above) + qkv weight gradients, one launch
      const WgradSpec w_proj = {p->dx16[cur], C, v.ao_b(l), C, v.boff(l, p->r_proj_w), -1};
      const WgradSpec w_qkv = {p->dqkv, 3 * C, v.xn_b(2 * l), C, v.boff(l, p->r_qkv_w), v.boff(l, p->r_qkv_b)};
      RC(linear_wgrad_pair(v, w_proj, w_qkv, M, sw));
    }
    RC(side_done(1));
    RC(linear_dgrad(v, p->dqkv, 3 * C, v.boff(l, p->r_qkv_w), C, MFV_EPI_BF16, p->dxn, nullptr, nullptr, 0, st));
    RC(join(0));  // fc2's weight gradient of this block reads dx16[cur ^ 1], which the LayerNorm backward below rewrites
    // fp32 output only where somebody reads it: the fp32-stream mode, or block 0 (the embedding backward)
    RCP(PROF_LN_BWD, mfv_layernorm_bwd(p->dxn, nullptr, res32(cur), res16(cur), v.x(2 * l), v.mean(2 * l), v.rstd(2 * l),
                         v.w32(v.boff(l, p->r_ln1_w)), (dx32 || l == 0) ? p->dx[cur ^ 1] : nullptr, p->dx16[cur ^ 1], v.gr(v.boff(l, p->r_ln1_w)),
                         v.gr(v.boff(l, p->r_ln1_b)), l > 0 ? v.gr(v.boff(l - 1, p->r_fc2_b)) : nullptr, G, M, C, p->P,
                         st));
    cur ^= 1;
  }
  RC(join(0));
  RC(join(1));  // every gradient of blocks block_hi..block_lo is complete in stream order when this call returns
  if (!(flags & MFV_BWD_TAIL)) return MFV_OK;
  // ---- embedding: cls gradient, conv bias / weight gradient (pos_embed is a fixed table)
  const long long rows_pe = p->B * p->np;
  RCP(PROF_EMBED_BWD, mfv_embed_finish_bwd(p->dx[cur], p->dacc, p->stop_grad_conv1 ? nullptr : v.gr(p->off_pe_b), v.gr(p->off_cls), G,
                          p->B, p->np, C, p->P, st));
  if (!p->stop_grad_conv1) {
    if (patch_rebuild && !ss) {
      // no side stream (MFVIT_SIDE_STREAM=0): the patch matrix is built here, in front of its only reader
      for (int g = 0; g < G; ++g) {
        if (!p->images[g]) return MFV_ERR_ARG;
        RCP(PROF_PATCHIFY, mfv_patchify(p->images[g],
                                        reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(v.patches_b())) + (long long)g * rows_pe * 768,
                                        0, nullptr, p->B, p->img, st));
      }
    }
    RC(linear_wgrad(v, p->dacc, C, v.patches_b(), 768, rows_pe, p->off_pe_w, -1, st));
  }
  return MFV_OK;
}
