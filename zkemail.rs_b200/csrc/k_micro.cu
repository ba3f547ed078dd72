#include "kernels.h"
#include "microbench.cuh"
namespace zkb {
void launch_int_peak(int kind, unsigned grid, unsigned block, uint32_t* out, uint32_t seed, int iters, cudaStream_t s) {
  switch (kind) {
    case 0: int_peak_kernel<0><<<grid, block, 0, s>>>(out, seed, iters); break;
    case 1: int_peak_kernel<1><<<grid, block, 0, s>>>(out, seed, iters); break;
    case 2: int_peak_kernel<2><<<grid, block, 0, s>>>(out, seed, iters); break;
    default: int_peak_kernel<3><<<grid, block, 0, s>>>(out, seed, iters); break;
  }
}
}  // namespace zkb
