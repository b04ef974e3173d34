// Host build of the device optimiser (csrc/lbfgsb_dense.cuh) for CPU-side testing against
// scipy.  Test infrastructure only: the product never loads this library.
#include "lbfgsb_dense.cuh"

typedef double (*lb_fn_t)(const double* x, double* g);

extern "C" int lbfgsb_host_minimize(int n, const double* x0, const double* l, const double* u, lb_fn_t fn,
                                    double* x_out, double* f_out, int* nfev, int* nit, int* status,
                                    double* trace, int trace_cap) {
  LbfgsbState s;
  lbfgsb_init(s, n, x0, l, u);
  double g[LB_NMAX];
  int evals = 0;
  for (;;) {
    double f = fn(s.x, g);
    if (trace && evals < trace_cap) {
      for (int i = 0; i < n; i++) trace[evals * (n + 1) + i] = s.x[i];
      trace[evals * (n + 1) + n] = f;
    }
    evals++;
    if (lbfgsb_feed(s, f, g) == LB_DONE) break;
    if (evals > 100000) break;
  }
  for (int i = 0; i < n; i++) x_out[i] = s.x[i];
  *f_out = s.f_last;
  *nfev = evals;
  *nit = s.iter;
  *status = s.status;
  return 0;
}
