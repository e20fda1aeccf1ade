// Library runtime: device binding, driver entry points, error strings.
#include "common.cuh"
#include "mfvit_internal.h"

namespace mfv {
static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 0;
static int g_device = -1;

PFN_encodeTiled get_encode_tiled() { return g_encode; }
int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }
}  // namespace mfv

extern "C" int mfv_abi_version(void) { return MFV_ABI_VERSION; }

extern "C" int mfv_num_sms(void) { return mfv::num_sms(); }

extern "C" int mfv_init(int device) {
  using namespace mfv;
  cudaDeviceProp prop;
  MFV_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return MFV_ERR_ARCH;  // sm_100a only: no fallback path exists
  MFV_CUDA_CHECK(cudaSetDevice(device));
  g_num_sms = prop.multiProcessorCount;
  g_device = device;
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
