// pir_div.cu — out-of-line definition of the division propagators; built with -DLPC_DIV_FIX=3 (see pir_div.cuh).
#include "pir_div.cuh"

namespace lpc {
__device__ __noinline__ void deduce_div(int op, Itv& r1, Itv& r2, Itv& r3) { deduce_div_rules(op, r1, r2, r3); }
} // namespace lpc
