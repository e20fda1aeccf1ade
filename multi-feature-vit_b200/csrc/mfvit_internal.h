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
}  // namespace mfv
