// div_miscompile.cu — repro of the ptxas 12.9 / sm_100a miscompile that keeps lala-pc_b200/csrc/pir_div.cu at -Xptxas -O0.
//
// The SAME source (the division propagators of pir_div.cuh, __host__ __device__ here) is evaluated on the host and on
// the device for every interval triple of [-R, R]^3 and each of the four division operators; one step of
// deduce_div_rules is pure register arithmetic, so host and device must agree bit for bit. Build and run (tools/repro/run.sh):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr div_miscompile.cu -o div_O3            (fails)
//   nvcc -O3 -Xptxas -O0 ...                                                            -o div_O0            (passes)
//   nvcc -O3 -DLPC_DIV_OPAQUE_NEG ...                                                   -o div_O3_opaque     
// Exit code 0 = no mismatch. Prints the first mismatching case per operator.
#include <cstdio>
#include <vector>
#define LPC_HOST_HARNESS
#include "../../lala-pc_b200/csrc/pir_div.cuh"
using namespace lpc;

struct Case { int v[6]; };
__global__ void k_div(int op, const Case* in, Case* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Itv r1(in[i].v[0], in[i].v[1]), r2(in[i].v[2], in[i].v[3]), r3(in[i].v[4], in[i].v[5]);
  deduce_div_rules(op, r1, r2, r3);
  out[i] = Case{{r1.lb, r1.ub, r2.lb, r2.ub, r3.lb, r3.ub}};
}

int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 6;
  std::vector<Case> cases;
  for(int xl = -R; xl <= R; ++xl) for(int xu = xl; xu <= R; ++xu)
  for(int yl = -R; yl <= R; ++yl) for(int yu = yl; yu <= R; ++yu)
  for(int zl = -R; zl <= R; ++zl) for(int zu = zl; zu <= R; ++zu) cases.push_back(Case{{xl, xu, yl, yu, zl, zu}});
  const int n = (int)cases.size();
  Case *d_in, *d_out;
  if(cudaMalloc(&d_in, n * sizeof(Case)) != cudaSuccess || cudaMalloc(&d_out, n * sizeof(Case)) != cudaSuccess) { printf("no device\n"); return 2; }
  cudaMemcpy(d_in, cases.data(), n * sizeof(Case), cudaMemcpyHostToDevice);
  std::vector<Case> got(n);
  int bad_total = 0;
  const int ops[4] = {D_TDIV, D_FDIV, D_CDIV, D_EDIV};
  const char* names[4] = {"TDIV", "FDIV", "CDIV", "EDIV"};
  for(int k = 0; k < 4; ++k) {
    k_div<<<(n + 255) / 256, 256>>>(ops[k], d_in, d_out, n);
    if(cudaMemcpy(got.data(), d_out, n * sizeof(Case), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("kernel failed\n"); return 2; }
    int bad = 0;
    for(int i = 0; i < n; ++i) {
      Itv r1(cases[i].v[0], cases[i].v[1]), r2(cases[i].v[2], cases[i].v[3]), r3(cases[i].v[4], cases[i].v[5]);
      deduce_div_rules(ops[k], r1, r2, r3);
      const int want[6] = {r1.lb, r1.ub, r2.lb, r2.ub, r3.lb, r3.ub};
      bool same = true;
      for(int j = 0; j < 6; ++j) same &= want[j] == got[i].v[j];
      if(!same && bad++ == 0)
        printf("%s x=[%d,%d] y=[%d,%d] z=[%d,%d]: host x=[%d,%d] y=[%d,%d] z=[%d,%d]  device x=[%d,%d] y=[%d,%d] z=[%d,%d]\n", names[k],
               cases[i].v[0], cases[i].v[1], cases[i].v[2], cases[i].v[3], cases[i].v[4], cases[i].v[5], want[0], want[1], want[2], want[3],
               want[4], want[5], got[i].v[0], got[i].v[1], got[i].v[2], got[i].v[3], got[i].v[4], got[i].v[5]);
    }
    printf("%s: %d of %d cases differ\n", names[k], bad, n);
    bad_total += bad;
  }
  return bad_total ? 1 : 0;
}
