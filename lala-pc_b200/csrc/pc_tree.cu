// pc_tree.cu — out-of-line instance of the tree interpreter (pc_tree.cuh) over the global-memory store accessor. Its own
// translation unit, compiled with -maxrregcount (Makefile): the interpreter is the cold path of the PC kernels and is
// called through the device ABI, so its register count must fit under every caller's launch bounds.
#include "pc_tree.cuh"

namespace lpc {
__device__ __noinline__ int pc_tree_deduce_global(GlobalAcc& a, const int* words) { return pc_tree_deduce_impl(a, words); }
__device__ __noinline__ bool pc_tree_ask_global(const GlobalAcc& a, const int* words) { return pc_tree_ask_impl(a, words); }
__device__ __noinline__ int pc_tree_deduce_global_bits(GlobalBitAcc& a, const int* words) { return pc_tree_deduce_impl(a, words); }
__device__ __noinline__ bool pc_tree_ask_global_bits(const GlobalBitAcc& a, const int* words) { return pc_tree_ask_impl(a, words); }
} // namespace lpc
