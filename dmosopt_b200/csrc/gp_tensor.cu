// tcgen05 split-precision path of the GP posterior variance (placeholder until the kernel lands).
#include "gp.cuh"

int gp_predict_tensor(dmo_ctx* ctx, dmo_gp* gp, const double* dXn, int64_t P, double* d_mean, double* d_var) {
  (void)gp;
  (void)dXn;
  (void)P;
  (void)d_mean;
  (void)d_var;
  return dmo_fail(ctx, DMO_ERR_UNSUPPORTED, "gp_predict: the tensor path is not built into this library");
}
