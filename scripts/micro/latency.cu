// Single-warp latency probes on sm_100a (calibration for the table-walk kernel): cycles per dependent step.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, int iters, double a, double lo, double hi, int kind) {
  double w = out[threadIdx.x];
  long long t0 = clock64();
  if (kind == 0) { for (int i = 0; i < iters; i++) w = w * a; }
  else if (kind == 1) { for (int i = 0; i < iters; i++) { double t = w * a; t = (t < lo) ? lo : t; w = (hi < t) ? hi : t; } }
  else if (kind == 2) { for (int i = 0; i < iters; i++) { double t = w * a; w = (t >= lo) ? t : lo; } }
  else if (kind == 3) { for (int i = 0; i < iters; i++) { double t = w * a; if (t == hi) { t = t * 0.5 + lo; } w = t; } }   // data-dependent branch, never taken
  else if (kind == 4) { for (int i = 0; i < iters; i++) w = fmin(fmax(w * a, lo), hi); }
  else if (kind == 5) { for (int i = 0; i < iters; i++) w = w + a; }
  else if (kind == 6) { unsigned long long x = __double_as_longlong(w); for (int i = 0; i < iters; i++) { x = x * 3 + 1; } w = __longlong_as_double(x >> 12 | 0x3ff0000000000000ull); }
  long long t1 = clock64();
  out[threadIdx.x] = w;
  if (threadIdx.x == 0) cyc[kind] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 16 * 8);
  double h[32]; for (int i = 0; i < 32; i++) h[i] = 0.5 + i * 1e-3;
  const char* names[] = {"DMUL", "DMUL+clamp2 (setp/sel)", "DMUL+max (setp/sel)", "DMUL+untaken data-dependent branch", "DMUL+fmax+fmin", "DADD", "IMAD64"};
  int iters = 1 << 16;
  for (int k = 0; k < 7; k++) {
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    probe<<<1, 32>>>(d, c, 256, 0.99999, 1e-4, 0.999, k);
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    probe<<<1, 32>>>(d, c, iters, 0.99999, 1e-4, 0.999, k);
    long long hc[16]; cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
    printf("%-40s %.2f cycles/step\n", names[k], (double)hc[k] / iters);
  }
  return 0;
}
