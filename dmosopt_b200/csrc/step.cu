// Fused resident NSGA-II surrogate generation ("fused dmo_generation_step").
//
// One C call = one pass of MOASMO.optimize's loop body (dmosopt/MOASMO.py:105-116) for the NSGA-II plugin with a GP
// surrogate, population resident in HBM:
//   tournament (NSGA2.py:116-140) -> variation loop (NSGA2.py:142-178) -> GP posterior mean [+ variance]
//   (model.py:1254-1275) -> children stacked over parents, rank + stable truncation (NSGA2.py:205-214, MOEA.py:398-423)
//   -> float32 rounding of the stored objectives (NSGA2.py:228-230) -> optional hypervolume of the survivors.
// It is a composition of the entry points of this library on device buffers (no host round trips except the offspring
// count and the hypervolume value); bench.py's `value` leg is this call.
#include "common.cuh"
#include "gp.cuh"

extern "C" {
int dmo_nsga2_step(dmo_ctx* ctx, dmo_gp* gp, double* pop_x, double* pop_y, int32_t* rank, int64_t pop, int d, int M,
                   double crossover_prob, double mutation_prob, double mutation_rate, const double* di_crossover,
                   const double* di_mutation, const double* xlb, const double* xub, uint64_t seed, uint64_t stream_id,
                   int precision, int distance_metric, int with_variance, int round_to_f32, const double* hv_ref, int64_t* n_children,
                   double* hv_out) {
  if (!ctx) return DMO_ERR_ARG;
  DMO_CUDA(cudaSetDevice(ctx->device));
  DMO_REQUIRE(gp && pop_x && pop_y && rank && pop >= 2 && d >= 1 && M >= 1, "nsga2_step: bad arguments");
  DMO_REQUIRE(distance_metric == DMO_METRIC_NONE || distance_metric == DMO_METRIC_CROWDING || distance_metric == DMO_METRIC_EUCLIDEAN,
              "nsga2_step: unknown distance metric %d", distance_metric);
  DMO_REQUIRE(dmo_is_device_ptr(pop_x) && dmo_is_device_ptr(pop_y) && dmo_is_device_ptr(rank),
              "nsga2_step: the population (pop_x, pop_y, rank) must be resident on the device");
  int64_t poolsize = pop / 2;  // int(round(popsize / 2.0)), NSGA2.py:64: Python rounds halves to even
  if ((pop & 1) && (poolsize & 1)) poolsize += 1;
  const int64_t cap = pop + 1;             // the variation loop emits pop-1 .. pop+1 children (NSGA2.py:142)
  DevBuf<int64_t> pool, perm;
  DevBuf<double> Xs, Ys, var;
  DevBuf<int32_t> kind;
  DMO_TRY(pool.alloc(ctx, poolsize));
  DMO_TRY(perm.alloc(ctx, pop));
  DMO_TRY(Xs.alloc(ctx, (size_t)(cap + pop) * d));
  DMO_TRY(Ys.alloc(ctx, (size_t)(cap + pop) * M));
  DMO_TRY(kind.alloc(ctx, cap));
  if (with_variance) DMO_TRY(var.alloc(ctx, (size_t)cap * M));
  int rc = dmo_tournament(ctx, rank, nullptr, pop, poolsize, seed, stream_id, pool.p, nullptr);
  if (rc != DMO_OK) return rc;
  int64_t P = 0;
  rc = dmo_nsga2_generate(ctx, pop_x, pop, d, pool.p, poolsize, pop, crossover_prob, mutation_prob, mutation_rate,
                          di_crossover, di_mutation, xlb, xub, seed, stream_id + 1, Xs.p, kind.p, &P, nullptr);
  if (rc != DMO_OK) return rc;
  if (n_children) *n_children = P;
  rc = dmo_gp_predict(ctx, gp, Xs.p, P, Ys.p, with_variance ? var.p : nullptr, precision);
  if (rc != DMO_OK) return rc;
  // parents under the children (np.vstack((x_gen, population_parm)), NSGA2.py:205-206)
  DMO_CUDA(cudaMemcpyAsync(Xs.p + (size_t)P * d, pop_x, (size_t)pop * d * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  DMO_CUDA(cudaMemcpyAsync(Ys.p + (size_t)P * M, pop_y, (size_t)pop * M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  rc = dmo_remove_worst(ctx, Xs.p, Ys.p, P + pop, d, M, distance_metric, nullptr, 0, pop, pop_x, pop_y, rank, perm.p);
  if (rc != DMO_OK) return rc;
  if (round_to_f32) {
    rc = dmo_round_f32(ctx, pop_y, pop * M);
    if (rc != DMO_OK) return rc;
  }
  if (hv_ref && hv_out) {
    // the survivors carry their ranks within the merged set: rows of rank > 0 cannot add volume (hv.cu)
    DMO_REQUIRE(M <= 16, "nsga2_step: too many objectives for the hypervolume");
    double h_ref[16];
    DMO_CUDA(cudaMemcpy(h_ref, hv_ref, M * sizeof(double), cudaMemcpyDefault));
    rc = hypervolume_device_ranked(ctx, pop_y, pop, M, h_ref, rank, hv_out);
    if (rc != DMO_OK) return rc;
  }
  return DMO_OK;
}
}
