// accuracy check of exp_nonpos / sqrt_nonneg against libm on the device
#include <cstdio>
#include <cmath>
#include "../../efficient-tlbo_b200/csrc/tlbo_common.cuh"
void tlbo_set_error(const char*) {}
__global__ void k(const double* x, double* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { o[2 * i] = exp_nonpos(-x[i]) ; o[2 * i + 1] = sqrt_nonneg(x[i]); }
}
int main() {
  const int n = 1 << 20;
  double* hx = new double[n]; double* ho = new double[2 * n];
  srand(1);
  for (int i = 0; i < n; i++) { double u = rand() / (double)RAND_MAX; hx[i] = (i % 3 == 0) ? u * 700 : (i % 3 == 1 ? u * 30 : u * 1e-3); }
  hx[0] = 0.0; hx[1] = 708.5; hx[2] = 1e-300; hx[3] = 745.0;
  double *dx, *d_o; cudaMalloc(&dx, n * 8); cudaMalloc(&d_o, 2 * n * 8);
  cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dx, d_o, n); cudaMemcpy(ho, d_o, 2 * n * 8, cudaMemcpyDeviceToHost);
  double me = 0, ms = 0;
  for (int i = 0; i < n; i++) {
    double e = exp(-hx[i]), s = sqrt(hx[i]);
    if (hx[i] < 708.0) { double r = fabs(ho[2 * i] - e) / e; if (r > me) me = r; }
    if (s > 0) { double r = fabs(ho[2 * i + 1] - s) / s; if (r > ms) ms = r; }
  }
  printf("max rel err exp %.3e sqrt %.3e ; exp(-0)=%g exp(-708.5)=%g exp(-745)=%g sqrt(0)=%g sqrt(1e-300)=%g [%s]\n", me, ms, ho[0], ho[2], ho[6],
         ho[1], ho[5], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
