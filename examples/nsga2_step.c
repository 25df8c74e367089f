/* Plain-C host for the resident generation step: what a non-Python controller would write against
 * include/dmosopt_b200.h.  Build:  gcc -std=c99 -Iinclude examples/nsga2_step.c -Ldmosopt_b200 -ldmosopt_b200
 *                                   -Wl,-rpath,$PWD/dmosopt_b200 -lm -o nsga2_step
 * The posterior state (X_train, alpha, Cholesky factors, kernel hyper-parameters) comes from whatever fitted the GP
 * (dmosopt fits with scikit-learn, dmosopt/model.py:1227-1251); here a toy one-point-per-objective model is used so
 * that the program is self-contained. */
#include <stdio.h>
#include <stdlib.h>

#include "dmosopt_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int st_ = (call);                                                            \
    if (st_ != 0) {                                                              \
      fprintf(stderr, "%s failed: %s\n", #call, ctx ? dmo_last_error(ctx) : "no context"); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(void) {
  enum { POP = 1024, D = 4, M = 2, N = 1 };
  dmo_ctx* ctx = NULL;
  CHECK(dmo_create(0, &ctx));

  /* a GP with a single training point per objective: K = c + noise, L = sqrt(K), alpha = y_n / K */
  double x_train[N * D] = {0.5, 0.5, 0.5, 0.5};
  double alpha[M * N] = {0.0, 0.0}, factor[M * N * N] = {1.0005, 1.0005};
  double constant[M] = {1.0, 1.0}, length_scale[M * D], noise[M] = {1e-3, 1e-3};
  double y_mean[M] = {0.0, 1.0}, y_std[M] = {1.0, 1.0}, xlb[D], xub[D], di_c[D], di_m[D];
  for (int j = 0; j < D; ++j) { xlb[j] = 0.0; xub[j] = 1.0; di_c[j] = 1.0; di_m[j] = 20.0; }
  for (int j = 0; j < M * D; ++j) length_scale[j] = 0.5;
  dmo_gp* gp = NULL;
  CHECK(dmo_gp_create(ctx, N, D, M, DMO_KERNEL_MATERN52, x_train, alpha, factor, 0, constant, length_scale, noise, y_mean,
                      y_std, xlb, xub, &gp));

  /* population resident on the device */
  double *h_x = malloc(sizeof(double) * POP * D), *h_y = malloc(sizeof(double) * POP * M);
  int32_t* h_r = calloc(POP, sizeof(int32_t));
  for (int i = 0; i < POP * D; ++i) h_x[i] = (double)rand() / RAND_MAX;
  for (int i = 0; i < POP; ++i) { h_y[i * M] = h_x[i * D]; h_y[i * M + 1] = 1.0 - h_x[i * D]; }
  void *d_x = NULL, *d_y = NULL, *d_r = NULL;
  CHECK(dmo_device_alloc(ctx, &d_x, sizeof(double) * POP * D));
  CHECK(dmo_device_alloc(ctx, &d_y, sizeof(double) * POP * M));
  CHECK(dmo_device_alloc(ctx, &d_r, sizeof(int32_t) * POP));
  CHECK(dmo_memcpy(ctx, d_x, h_x, sizeof(double) * POP * D));
  CHECK(dmo_memcpy(ctx, d_y, h_y, sizeof(double) * POP * M));
  CHECK(dmo_rank_nd(ctx, (const double*)d_y, POP, M, (int32_t*)d_r));

  const double ref[M] = {2.0, 2.0};
  for (int gen = 0; gen < 5; ++gen) {
    int64_t n_children = 0;
    double hv = 0.0;
    CHECK(dmo_nsga2_step(ctx, gp, (double*)d_x, (double*)d_y, (int32_t*)d_r, POP, D, M, 0.9, 0.1, 1.0 / D, di_c, di_m, xlb, xub,
                         /*seed*/ 42u, /*stream*/ 2u * gen + 1u, DMO_GP_AUTO, DMO_METRIC_CROWDING, /*variance*/ 1, /*float32 state*/ 1, ref,
                         &n_children, &hv));
    printf("generation %d: %lld children, hypervolume %.6f\n", gen, (long long)n_children, hv);
  }
  dmo_device_free(ctx, d_x);
  dmo_device_free(ctx, d_y);
  dmo_device_free(ctx, d_r);
  dmo_gp_destroy(ctx, gp);
  dmo_destroy(ctx);
  free(h_x);
  free(h_y);
  free(h_r);
  return 0;
}
