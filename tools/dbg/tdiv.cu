#include <cstdio>
#include "../../lala-pc_b200/csrc/pir_device.cuh"
using namespace lpc;
__global__ void k(int* out) {
  Itv r1(-5, 0), r2(-5, 1), r3(-1, -1);
  deduce_regs<true>(D_TDIV, r1, r2, r3);
  out[0] = r1.lb; out[1] = r1.ub; out[2] = r2.lb; out[3] = r2.ub; out[4] = r3.lb; out[5] = r3.ub;
  Itv a(-1, 0), c(-1, -1);
  Itv n = num_tdiv(a, c);
  out[6] = n.lb; out[7] = n.ub;
  Itv f = num_fdiv(Itv(1, 0), c);
  out[8] = f.lb; out[9] = f.ub;
  Itv g = num_cdiv(Itv(-1, -1), c);
  out[10] = g.lb; out[11] = g.ub;
}
int main() {
  int* d; cudaMalloc(&d, 64); k<<<1, 1>>>(d); int h[12]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("deduce: x=[%d,%d] y=[%d,%d] z=[%d,%d]  num_tdiv=[%d,%d] num_fdiv(1,0)=[%d,%d] num_cdiv(-1,-1)=[%d,%d]\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11]);
  Itv a(-1, 0), c(-1, -1);
  Itv n = num_tdiv(a, c);
  printf("host num_tdiv=[%d,%d]\n", n.lb, n.ub);
  return 0;
}
