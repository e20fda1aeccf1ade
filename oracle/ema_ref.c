/* C restatement of the MoCo momentum update (BLD:83-89): k = k*m + q*(1-m) as three separately rounded fp32 operations
 * (mul, mul, add), which is what the eager PyTorch sequence computes.  Compile with -ffp-contract=off so the compiler
 * cannot fuse the multiply-add.  TEST INFRASTRUCTURE ONLY (bit-exactness oracle for mfv_ema_update). */
#include <stddef.h>
void ema_ref(float* k, const float* q, size_t n, float m, float one_minus_m) {
  for (size_t i = 0; i < n; ++i) {
    volatile float a = k[i] * m;
    volatile float b = q[i] * one_minus_m;
    k[i] = a + b;
  }
}
