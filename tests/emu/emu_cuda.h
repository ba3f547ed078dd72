// emu_cuda.h — TEST INFRASTRUCTURE: a minimal host emulation of the CUDA execution model so that
// the product's kernel sources (zkemail.rs_b200/csrc/*.cuh) can be compiled with g++ and run on the
// CPU test box, which has no GPU.  Every CUDA thread of a block is one std::thread; warp
// collectives (__shfl*_sync, __ballot_sync) exchange values through a per-warp mailbox guarded by
// a std::barrier, so lanes really execute the same source lines the GPU runs, in lock step at the
// collectives.  The PTX carry-chain macros of rsa.cuh are emulated on a per-thread carry flag.
// This checks the kernels' arithmetic and control flow before GPU time is spent; the GPU parity
// tests (-m gpu) remain the authority.
#pragma once
#define ZKB_HOST_EMU 1
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
typedef int cudaError_t;
typedef void* cudaStream_t;

namespace emu {
struct Warp {
  std::barrier<> bar;
  uint32_t box[2][32];
  explicit Warp(int n) : bar(n) {}
};
struct Block {
  std::barrier<> bar;
  explicit Block(int n) : bar(n) {}
};
struct Ctx {
  dim3 tid, bid, bdim, gdim;
  Warp* warp = nullptr;
  Block* block = nullptr;
  int lane = 0;
  int phase = 0;
  uint32_t cf = 0;  // PTX condition-code carry flag
};
inline thread_local Ctx ctx;

// Runs `fn` for every thread of a 1-D grid of 1-D blocks (block size a multiple of 32).
inline void launch(unsigned grid, unsigned block, const std::function<void()>& fn) {
  for (unsigned b = 0; b < grid; b++) {
    unsigned nw = (block + 31) / 32;
    std::vector<std::unique_ptr<Warp>> warps;
    for (unsigned w = 0; w < nw; w++) {
      unsigned lanes = std::min(32u, block - 32 * w);
      warps.emplace_back(new Warp((int)lanes));
    }
    Block blk((int)block);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; t++) {
      th.emplace_back([&, t]() {
        ctx = Ctx();
        ctx.tid = dim3(t); ctx.bid = dim3(b); ctx.bdim = dim3(block); ctx.gdim = dim3(grid);
        ctx.warp = warps[t / 32].get(); ctx.block = &blk; ctx.lane = (int)(t % 32);
        fn();
        // a thread that returns early must not dead-lock the others at later collectives
        ctx.warp->bar.arrive_and_drop();
        ctx.block->bar.arrive_and_drop();
      });
    }
    for (auto& x : th) x.join();
  }
}
inline uint32_t exchange(uint32_t v, int src_lane) {
  Ctx& c = ctx;
  int ph = c.phase;
  c.phase ^= 1;
  c.warp->box[ph][c.lane] = v;
  c.warp->bar.arrive_and_wait();
  // a lane that already exited leaves a stale value; CUDA calls that undefined, kernels avoid it
  return c.warp->box[ph][src_lane & 31];
}
}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)

static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src, int width = 32) {
  int lane = emu::ctx.lane;
  int base = lane & ~(width - 1);
  return emu::exchange(v, base + (src & (width - 1)));
}
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d, int width = 32) {
  int lane = emu::ctx.lane;
  int base = lane & ~(width - 1);
  int src = lane + (int)d;
  if (src >= base + width) src = lane;
  return emu::exchange(v, src);
}
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d, int width = 32) {
  int lane = emu::ctx.lane;
  int base = lane & ~(width - 1);
  int src = lane - (int)d;
  if (src < base) src = lane;
  return emu::exchange(v, src);
}
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m, int width = 32) {
  (void)width;
  return emu::exchange(v, emu::ctx.lane ^ m);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::Ctx& c = emu::ctx;
  int ph = c.phase;
  c.phase ^= 1;
  c.warp->box[ph][c.lane] = pred ? 1u : 0u;
  c.warp->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (c.warp->box[ph][i] & 1u) << i;
  return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  emu::Ctx& c = emu::ctx;
  int ph = c.phase;
  c.phase ^= 1;
  c.warp->box[ph][c.lane] = v;
  c.warp->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= c.warp->box[ph][i];
  return r;
}
static inline void __syncthreads() { emu::ctx.block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx.warp->bar.arrive_and_wait(); emu::ctx.phase ^= 0; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, unsigned s) {
  s &= 31;
  return s ? (uint32_t)((((uint64_t)hi << 32) | lo) >> s) : lo;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, unsigned s) {
  s &= 31;
  return s ? (uint32_t)(((((uint64_t)hi << 32) | lo) << s) >> 32) : hi;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) {
  return __atomic_fetch_or(p, v, __ATOMIC_RELAXED);
}
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

// ---- PTX multiply-add / add / sub with carry: semantics of mad.lo.cc / madc.hi.cc etc. ----
namespace emu {
inline void madw(uint32_t& d0, uint32_t& d1, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1, bool cin) {
  unsigned __int128 v = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)c1 << 32) | c0) + (cin ? ctx.cf : 0);
  d0 = (uint32_t)v; d1 = (uint32_t)(v >> 32); ctx.cf = (uint32_t)(v >> 64) & 1;
}
inline void addc(uint32_t& d, uint32_t a, bool cin, bool cout) {
  uint64_t v = (uint64_t)d + a + (cin ? ctx.cf : 0);
  d = (uint32_t)v;
  if (cout) ctx.cf = (uint32_t)(v >> 32);
}
inline void subc(uint32_t& d, uint32_t a, bool cin, bool cout) {
  // PTX sub.cc: CC.CF = borrow-out; subc: d = a - b - CC.CF
  uint64_t v = (uint64_t)d - a - (cin ? ctx.cf : 0);
  d = (uint32_t)v;
  if (cout) ctx.cf = (uint32_t)(v >> 63) & 1;
}
}  // namespace emu
