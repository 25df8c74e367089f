// Vectorised multi-objective benchmark functions.
// The reference evaluates them one row at a time in Python (dmosopt/benchmarks/moo_benchmarks.py:21-375; ZDT1 / ZDT3 are
// the example objectives, examples/example_dmosopt_zdt1.py:9-20, examples/example_dmosopt_zdt3.py:9-21), which becomes
// the bottleneck of every non-surrogate run (MOASMO.py:57-58, 107-108).  Here: one thread per row, float64, the
// reference's formulas term by term (sums in index order; results agree to ~1e-15 relative, NumPy sums pairwise).
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int BM_MAXOBJ = 16;
constexpr double PI = 3.14159265358979323846;

enum { BM_ZDT1 = 0, BM_ZDT3 = 1, BM_DTLZ1 = 10, BM_DTLZ2 = 11, BM_DTLZ3 = 12, BM_DTLZ4 = 13, BM_DTLZ5 = 14, BM_DTLZ7 = 16, BM_WFG4 = 24 };

// f_i = scale * prod_{j < M-i-1} c(x_j) * (i > 0 ? s(x_{M-i-1}) : 1): the product form shared by DTLZ1-5 and the WFG shapes
template <class C, class S>
__device__ __forceinline__ void product_form(const double* v, int M, double scale, C c, S s, double* f) {
  for (int i = 0; i < M; ++i) {
    double fi = scale;
    for (int j = 0; j < M - i - 1; ++j) fi *= c(v[j]);
    if (i > 0) fi *= s(v[M - i - 1]);
    f[i] = fi;
  }
}

__global__ void benchmark_kernel(int problem, const double* __restrict__ X, int64_t n, int d, int M, double alpha,
                                 double* __restrict__ Y) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double* x = X + r * d;
  double f[BM_MAXOBJ];
  const int k = d - M + 1;  // DTLZ: the last k variables drive g
  switch (problem) {
    case BM_ZDT1:
    case BM_ZDT3: {
      double s = 0.0;
      for (int j = 1; j < d; ++j) s += x[j];
      const double g = 1.0 + 9.0 / (d - 1) * s;
      f[0] = x[0];
      if (problem == BM_ZDT1)
        f[1] = g * (1.0 - sqrt(x[0] / g));
      else
        f[1] = g * (1.0 - sqrt(x[0] / g)) - (x[0] / g) * sin(10.0 * PI * x[0]);  // as the reference's example writes it (g h - j)
      break;
    }
    case BM_DTLZ1:
    case BM_DTLZ3: {
      double s = 0.0;
      for (int j = d - k; j < d; ++j) {
        const double t = x[j] - 0.5;
        s += t * t - cos(20.0 * PI * t);
      }
      const double g = 100.0 * (k + s);
      if (problem == BM_DTLZ1)
        product_form(x, M, 0.5 * (1.0 + g), [](double v) { return v; }, [](double v) { return 1.0 - v; }, f);
      else
        product_form(x, M, 1.0 + g, [](double v) { return cos(v * PI / 2); }, [](double v) { return sin(v * PI / 2); }, f);
      break;
    }
    case BM_DTLZ2:
    case BM_DTLZ4:
    case BM_DTLZ5: {
      double g = 0.0;
      for (int j = d - k; j < d; ++j) {
        const double t = x[j] - 0.5;
        g += t * t;
      }
      if (problem == BM_DTLZ2) {
        product_form(x, M, 1.0 + g, [](double v) { return cos(v * PI / 2); }, [](double v) { return sin(v * PI / 2); }, f);
      } else if (problem == BM_DTLZ4) {
        product_form(x, M, 1.0 + g, [alpha](double v) { return cos(pow(v, alpha) * PI / 2); },
                     [alpha](double v) { return sin(pow(v, alpha) * PI / 2); }, f);
      } else {
        double th[BM_MAXOBJ];
        th[0] = x[0] * PI / 2;
        for (int i = 1; i < M - 1; ++i) th[i] = (1.0 + 2.0 * g * x[i]) / (2.0 * (1.0 + g)) * PI / 2;
        product_form(th, M, 1.0 + g, [](double v) { return cos(v); }, [](double v) { return sin(v); }, f);
      }
      break;
    }
    case BM_DTLZ7: {
      double s = 0.0;
      for (int j = d - k; j < d; ++j) s += x[j];
      const double g = 1.0 + 9.0 * (s / k);
      double h = 0.0;
      for (int i = 0; i < M - 1; ++i) {
        f[i] = x[i];
        h += f[i] / (1.0 + g) * (1.0 + sin(3.0 * PI * f[i]));
      }
      f[M - 1] = (1.0 + g) * ((double)M - h);
      break;
    }
    case BM_WFG4: {
      // y = x / (2 i); t1 = y + 0.35 - 0.15 cos(10 pi y - 5); shape vector = means over ll = d - (M - 1) wide windows
      const int kk = M - 1, ll = d - kk;
      double xv[BM_MAXOBJ];
      for (int i = 0; i < M; ++i) {
        const int lo = (i < M - 1) ? i * ll : d - ll;
        const int hi = (i < M - 1) ? ((i + 1) * ll < d ? (i + 1) * ll : d) : d;
        double s = 0.0;
        for (int j = lo; j < hi; ++j) {
          const double y = x[j] / (2.0 * (j + 1));
          s += y + 0.35 - 0.15 * cos(10.0 * PI * y - 5.0);
        }
        xv[i] = hi > lo ? s / (hi - lo) : CUDART_NAN;  // np.mean of an empty slice is nan in the reference as well
      }
      product_form(xv, M, 1.0, [](double v) { return 1.0 - cos(v * PI / 2); }, [](double v) { return 1.0 - sin(v * PI / 2); }, f);
      for (int i = 0; i < M; ++i) f[i] *= (double)(i + 2);  // * (1 + arange(1, M + 1))
      break;
    }
    default:
      for (int i = 0; i < M; ++i) f[i] = CUDART_NAN;
  }
  for (int i = 0; i < M; ++i) Y[r * M + i] = f[i];
}

}  // namespace

extern "C" {

int dmo_benchmark_eval(dmo_ctx* ctx, int problem, const double* X, int64_t n, int n_var, int n_obj, double alpha, double* Y) {
  if (!ctx) return DMO_ERR_ARG;
  DMO_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return DMO_OK;
  DMO_REQUIRE(n > 0 && X && Y && n_var >= 2 && n_obj >= 2 && n_obj <= BM_MAXOBJ, "benchmark_eval: bad arguments");
  const bool zdt = problem == BM_ZDT1 || problem == BM_ZDT3;
  const bool known = zdt || problem == BM_DTLZ1 || problem == BM_DTLZ2 || problem == BM_DTLZ3 || problem == BM_DTLZ4 ||
                     problem == BM_DTLZ5 || problem == BM_DTLZ7 || problem == BM_WFG4;
  if (!known) return dmo_fail(ctx, DMO_ERR_UNSUPPORTED, "benchmark_eval: unknown problem id %d", problem);
  DMO_REQUIRE(!zdt || n_obj == 2, "benchmark_eval: ZDT problems have two objectives");
  DMO_REQUIRE(zdt || n_var >= n_obj, "benchmark_eval: n_var must be at least n_obj");
  In<double> x;
  Out<double> y;
  DMO_TRY(x.init(ctx, X, (size_t)n * n_var));
  DMO_TRY(y.init(ctx, Y, (size_t)n * n_obj));
  DMO_LAUNCH(benchmark_kernel, (unsigned)ceil_div(n, 128), 128, 0, problem, x.d, n, n_var, n_obj, alpha, y.d);
  DMO_CHECK_LAUNCH();
  DMO_TRY(y.finish(ctx));
  DMO_CUDA(cudaStreamSynchronize(ctx->stream));
  return DMO_OK;
}

}  // extern "C"
