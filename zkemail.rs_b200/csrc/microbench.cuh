// microbench.cuh — register-only integer-pipe issue-rate probes.  MEASURED_PEAKS.json carries no
// integer peak, so the roofline denominators for the SHA-256 (ALU pipe:
// IADD3/LOP3/SHF) and RSA (FMA pipe: IMAD.WIDE) kernels are measured live with these.
#pragma once
#include "common.cuh"

namespace zkb {

// kind 0: IMAD.WIDE.U32 with carry chains (8 independent accumulators per thread)
// kind 1: IADD3   kind 2: LOP3   kind 3: SHF (funnel shift)
template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a * (i + 1) + b;
  for (int it = 0; it < iters; it++) {
    if (KIND == 0) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        // 8 wide MACs in two carry chains of 4 (mirrors the RSA inner loop)
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
            "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
            "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
            "mad.lo.cc.u32 %8, %17, %16, %8;\n\tmadc.hi.cc.u32 %9, %17, %16, %9;\n\t"
            "madc.lo.cc.u32 %10, %17, %16, %10;\n\tmadc.hi.cc.u32 %11, %17, %16, %11;\n\t"
            "madc.lo.cc.u32 %12, %17, %16, %12;\n\tmadc.hi.cc.u32 %13, %17, %16, %13;\n\t"
            "madc.lo.cc.u32 %14, %17, %16, %14;\n\tmadc.hi.u32 %15, %17, %16, %15;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
              "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
              "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
            : "r"(a), "r"(b));
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (KIND == 1) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
          if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
          if (KIND == 3) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(a));
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc ^= r[i];
  if (acc == 0x12345678u) out[0] = acc;  // keep the work alive
}

}  // namespace zkb
