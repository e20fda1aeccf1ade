// Library runtime: device binding, driver entry points, error strings.
#include "common.cuh"
#include "mfvit_internal.h"

#include <stdlib.h>
#include <string>
#include <vector>

namespace mfv {
unsigned long long g_launch_count = 0;
static thread_local char g_err_where[256] = "";
void note_error(const char* file, int line, const char* expr) {
  const char* base = file;
  for (const char* c = file; *c; ++c) if (*c == '/') base = c + 1;
  snprintf(g_err_where, sizeof(g_err_where), "%s:%d: %s", base, line, expr);
}
static bool g_prof_on = false;
struct ProfRec { int label; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<ProfRec> g_prof_pool;
static ProfRec g_prof_open[PROF_NUM_LABELS];
bool prof_enabled() { return g_prof_on; }
void prof_begin(int label, cudaStream_t st) {
  ProfRec r;
  if (!g_prof_pool.empty()) { r = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
  r.label = label;
  cudaEventRecord(r.a, st);
  g_prof_open[label] = r;
}
void prof_end(int label, cudaStream_t st) {
  ProfRec r = g_prof_open[label];
  cudaEventRecord(r.b, st);
  g_prof_recs.push_back(r);
}
static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 0;
static int g_device = -1;

PFN_encodeTiled get_encode_tiled() { return g_encode; }
// SMs the persistent kernels size their grids for.  "reserve_sms" (MFVIT_RESERVE_SMS, set by the data-parallel trainer)
// leaves that many SMs free for the NCCL all-reduce kernels that run beside the backward: a one-CTA-per-SM GEMM grid
// otherwise makes every collective CTA wait for a tile to finish.
static int g_opt_reserve = -1;
int num_sms() {
  if (g_opt_reserve < 0) {
    const char* e = getenv("MFVIT_RESERVE_SMS");
    g_opt_reserve = e ? atoi(e) : 0;
  }
  const int n = g_num_sms > 0 ? g_num_sms : 148;
  const int r = g_opt_reserve < 0 ? 0 : (g_opt_reserve > n / 2 ? n / 2 : g_opt_reserve);
  return (n - r) & ~1;  // even: CTA pairs
}
// Runtime switches (defaults from the environment, overridable through mfv_set_option): -1 = not read yet
static int g_opt_pdl = -1, g_opt_side = -1, g_opt_legacy_attn = -1, g_opt_rows96 = -1, g_opt_fuse_ln = -1, g_opt_dx32 = -1, g_opt_patch_tma = -1;
static int env_flag(const char* name, int dflt, char off_char) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] == off_char ? !dflt : dflt;
}
SideStream* side_stream() {
  static SideStream ss;
  static int state = 0;  // 0 = untried, 1 = ready, -1 = creation failed
  if (g_opt_side < 0) g_opt_side = env_flag("MFVIT_SIDE_STREAM", 1, '0');
  if (!g_opt_side || prof_enabled()) return nullptr;  // per-class timings want serialised launches
  if (state == 0) {
    state = -1;
    if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) == cudaSuccess) {
      bool ok = true;
      for (int i = 0; i < 2; ++i) {
        ok = ok && cudaEventCreateWithFlags(&ss.fork[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ss.done[i], cudaEventDisableTiming) == cudaSuccess;
      }
      ok = ok && cudaEventCreateWithFlags(&ss.fusion_fork, cudaEventDisableTiming) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&ss.fusion_done, cudaEventDisableTiming) == cudaSuccess;
      if (ok) state = 1;
    }
  }
  return state == 1 ? &ss : nullptr;
}
bool legacy_attention() {
  if (g_opt_legacy_attn < 0) g_opt_legacy_attn = env_flag("MFVIT_ATTN", 0, 'l');
  return g_opt_legacy_attn == 1;
}
int rows96_mode() {
  if (g_opt_rows96 < 0) {
    const char* e = getenv("MFVIT_ROWS96");
    g_opt_rows96 = (e && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : 0;
  }
  return g_opt_rows96;
}
bool fuse_ln_enabled() {
  if (g_opt_fuse_ln < 0) g_opt_fuse_ln = env_flag("MFVIT_FUSE_LN", 1, '0');
  return g_opt_fuse_ln == 1;
}
bool dx32_stream_enabled() {
  if (g_opt_dx32 < 0) g_opt_dx32 = env_flag("MFVIT_DX32", 0, '1');
  return g_opt_dx32 == 1;
}
bool patch_tma_enabled() {
  if (g_opt_patch_tma < 0) g_opt_patch_tma = env_flag("MFVIT_PATCH_TMA", 1, '0');
  return g_opt_patch_tma == 1;
}
bool pdl_enabled() {
  if (g_opt_pdl < 0) g_opt_pdl = env_flag("MFVIT_PDL", 1, '0');
  return g_opt_pdl == 1;
}
// Stream-K (gemm.cu): bit mask of the GEMM families whose tile lists are cut into equal k-block ranges per CTA pair.
// 1 = 384-wide pair tiles (proj / fc2 forward, every dgrad), 2 = 256-wide bf16 (qkv), 4 = GELU (fc1), 8 = GELU' (fc2 dgrad)
// Default 0.  Measured (profiles/r02_summary.md, fc2-shaped bf16 GEMM, graph-timed): the mainloop is bound by operand
// bytes out of L2 (~6-7 TB/s whatever the number of busy pairs), so spreading 50 tiles over 74 pairs does not shorten it
// (17.5 -> 17.3 us) and 100 tiles only by 15 % (30.9 -> 26.3 us), while the partial dump + add through L2 costs 10 + 23 us
// as written (20.2 -> 52.9 us in total).  Step: 4.70 ms off, 5.80 ms with mask 1.
static int g_opt_streamk = -1;
int streamk_mask() {
  if (g_opt_streamk < 0) {
    const char* e = getenv("MFVIT_STREAMK");
    g_opt_streamk = e ? atoi(e) : 0;
  }
  return g_opt_streamk;
}
// B-operand multicast between two CTA pairs in the 384-wide GEMMs (gemm.cu, MC kernels): MFVIT_GEMM_MC=1.  Off by
// default: it takes 13 % off the L2 traffic of an fc2-shaped GEMM (152 -> 132 MB, ncu lts__t_bytes) and nothing off its
// duration (25.9 us cold either way; step 4.71 vs 4.74 ms) - these mainloops are UMMA-bound, not bound by L2 bytes.
static int g_opt_gemm_mc = -1;
bool gemm_multicast_enabled() {
  if (g_opt_gemm_mc < 0) g_opt_gemm_mc = env_flag("MFVIT_GEMM_MC", 0, '1');
  return g_opt_gemm_mc == 1;
}
static int g_opt_streamk_min_kb = -1;
int streamk_min_kb() {
  if (g_opt_streamk_min_kb < 0) {
    const char* e = getenv("MFVIT_STREAMK_MINKB");
    g_opt_streamk_min_kb = e ? atoi(e) : 12;
  }
  return g_opt_streamk_min_kb;
}
// Partial-accumulator workspace of the stream-K GEMMs: one slot per CTA (fp32 [16 warps][BN / 4 x 32 values], BN <= 384)
// + two counters per CTA.  Allocated when the option is switched on (mfv_init with MFVIT_STREAMK set, or mfv_set_option -
// never inside a stream capture), zeroed once; the kernels leave the counters at zero.
static float* g_sk_ws = nullptr;
static unsigned* g_sk_flags = nullptr;
float* streamk_workspace() { return g_sk_ws; }
unsigned* streamk_flags() { return g_sk_flags; }
static int streamk_alloc(int sms) {
  if (g_sk_ws) return MFV_OK;
  const size_t per_cta = (size_t)16 * (384 / 4 * 32) * sizeof(float);
  MFV_CUDA_CHECK(cudaMalloc(&g_sk_ws, per_cta * (size_t)sms));
  MFV_CUDA_CHECK(cudaMalloc(&g_sk_flags, sizeof(unsigned) * 2 * (size_t)sms));
  MFV_CUDA_CHECK(cudaMemset(g_sk_flags, 0, sizeof(unsigned) * 2 * (size_t)sms));
  return MFV_OK;
}
}  // namespace mfv

// Source location and expression of the last CUDA runtime failure seen by this thread ("" if none).
extern "C" const char* mfv_last_error_where(void) { return mfv::g_err_where; }

// Runtime switches for A/B measurements and tests: "pdl" (programmatic dependent launch, default 1), "side_stream"
// (weight gradients on a second stream, default 1), "legacy_attention" (mma.sync attention kernels, default 0),
// "rows96" (192-row pair tiles for the forward N = 384 GEMMs when they fill the SMs better, default 0), "fuse_ln"
// (LayerNorm inside the proj / fc2 epilogues, default 1).
extern "C" int mfv_set_option(const char* key, int value) {
  using namespace mfv;
  if (!key) return MFV_ERR_ARG;
  const std::string k(key);
  if (k == "pdl") g_opt_pdl = value ? 1 : 0;
  else if (k == "side_stream") g_opt_side = value ? 1 : 0;
  else if (k == "legacy_attention") g_opt_legacy_attn = value ? 1 : 0;
  else if (k == "rows96") g_opt_rows96 = value < 0 ? 0 : value;
  else if (k == "fuse_ln") g_opt_fuse_ln = value ? 1 : 0;
  else if (k == "dx32") g_opt_dx32 = value ? 1 : 0;
  else if (k == "patch_tma") g_opt_patch_tma = value ? 1 : 0;
  else if (k == "reserve_sms") g_opt_reserve = value < 0 ? 0 : value;
  else if (k == "gemm_mc") g_opt_gemm_mc = value ? 1 : 0;
  else if (k == "streamk") {  // not inside a stream capture: switching it on allocates the partial-sum workspace
    g_opt_streamk = value < 0 ? 0 : value;
    if (g_opt_streamk && g_num_sms > 0) {
      const int rc = streamk_alloc(g_num_sms);
      if (rc) return rc;
    }
  }
  else if (k == "streamk_min_kb") g_opt_streamk_min_kb = value < 1 ? 1 : value;
  else return MFV_ERR_ARG;
  return MFV_OK;
}

extern "C" int mfv_abi_version(void) { return MFV_ABI_VERSION; }

extern "C" uint64_t mfv_launch_count(void) { return mfv::g_launch_count; }

extern "C" int mfv_prof_enable(int on) {
  mfv::g_prof_on = on != 0;
  return MFV_OK;
}

extern "C" int mfv_prof_num_labels(void) { return mfv::PROF_NUM_LABELS; }

extern "C" const char* mfv_prof_label_name(int label) {
  static const char* names[] = {"patchify", "gemm_fwd", "ln_fwd", "attn_fwd", "embed", "gemm_dgrad", "gemm_wgrad",
                                "colsum", "ln_bwd", "attn_bwd", "embed_bwd"};
  return (label >= 0 && label < mfv::PROF_NUM_LABELS) ? names[label] : "?";
}

// Synchronises the device, adds the elapsed time of every recorded scope to ms[label] / count[label], resets.
extern "C" int mfv_prof_read(float* ms, int* count, int nlabels) {
  using namespace mfv;
  MFV_CUDA_CHECK(cudaDeviceSynchronize());
  for (const ProfRec& r : g_prof_recs) {
    float t = 0.f;
    MFV_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.label < nlabels) { ms[r.label] += t; count[r.label] += 1; }
    g_prof_pool.push_back(r);
  }
  g_prof_recs.clear();
  return MFV_OK;
}

extern "C" int mfv_num_sms(void) { return mfv::num_sms(); }

extern "C" int mfv_init(int device) {
  using namespace mfv;
  cudaDeviceProp prop;
  MFV_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return MFV_ERR_ARCH;  // sm_100a only: no fallback path exists
  // One device per process (the deployment model: one process per GPU).  The side stream, its events, the opt-in
  // shared-memory attributes and the SM count are process-wide state bound to the first device.
  if (g_device >= 0 && g_device != device) {
    note_error(__FILE__, __LINE__, "mfv_init called for a second device in one process");
    return MFV_ERR_ARG;
  }
  MFV_CUDA_CHECK(cudaSetDevice(device));
  g_num_sms = prop.multiProcessorCount;
  g_device = device;
  if (streamk_mask() != 0) {  // opt-in experiment: its 29 MB workspace exists only when it is switched on
    const int rc = streamk_alloc(g_num_sms);
    if (rc) return rc;
  }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MFV_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) return MFV_ERR_INIT;
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  return MFV_OK;
}

extern "C" const char* mfv_strerror(int code) {
  switch (code) {
    case MFV_OK: return "ok";
    case MFV_ERR_SHAPE: return "mfvit: unsupported shape";
    case MFV_ERR_ALIGN: return "mfvit: pointer or stride not 16-byte aligned";
    case MFV_ERR_ARCH: return "mfvit: device is not sm_100 (no fallback path)";
    case MFV_ERR_INIT: return "mfvit: mfv_init not called or driver entry point missing";
    case MFV_ERR_ARG: return "mfvit: invalid argument";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "mfvit: unknown error";
}
