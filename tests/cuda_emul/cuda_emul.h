// TEST INFRASTRUCTURE ONLY. A minimal host emulation of the CUDA execution model, enough to run the simple (non-tensor-core)
// kernels of transfuser_b200/csrc *unchanged* on the CPU in the build container, where there is no GPU: one OS thread per CUDA
// thread of a block, blocks executed one after another, __syncthreads() = a pthread barrier, __shared__ = a static.
// tests/test_kernel_emulation.py pastes the kernel part of a .cu file (everything above its C-ABI entry points) after this
// header, compiles it with g++ and compares the results with the oracle. It checks indexing / control flow / arithmetic
// order of the kernel source; it says nothing about performance, memory-model races or the real device compiler.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <functional>
#include <thread>
#include <vector>

struct uint3_ { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint2 { uint32_t x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
// bf16 storage type with round-to-nearest-even conversion (NaN kept quiet), as cuda_bf16.h's __float2bfloat16_rn
struct __nv_bfloat16 { uint16_t bits; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{(uint16_t)0x7fff};
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }

static thread_local uint3_ threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
static pthread_barrier_t emul_barrier;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static const
#define __shared__ static
#define INFINITY_F (__builtin_inff())

static inline void __syncthreads() { pthread_barrier_wait(&emul_barrier); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v) {
  uint32_t* ip = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), want;
  float f;
  do {
    memcpy(&f, &old, 4);
    f += v;
    memcpy(&want, &f, 4);
  } while (!__atomic_compare_exchange_n(ip, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &old, 4);
  return f;
}
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
using std::floor;
using std::fma;
using std::fmax;
using std::fmin;

static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline double atomicAdd(double* p, double v) {
  uint64_t* ip = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), want;
  double f;
  do {
    memcpy(&f, &old, 8);
    f += v;
    memcpy(&want, &f, 8);
  } while (!__atomic_compare_exchange_n(ip, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &old, 8);
  return f;
}
// common.cuh's block-wide reductions (warp shuffles on the device): here through a static scratch and barriers. The
// summation order differs from the device's tree, so only use them where rounding of the reduction is not under test.
static float emul_red[1024];
static inline unsigned emul_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
static inline float block_sum(float v, float*) {
  const unsigned n = blockDim.x * blockDim.y * blockDim.z;
  __syncthreads();
  emul_red[emul_tid()] = v;
  __syncthreads();
  float s = 0.f;
  for (unsigned i = 0; i < n; ++i) s += emul_red[i];
  __syncthreads();
  return s;
}
static inline float block_max(float v, float*) {
  const unsigned n = blockDim.x * blockDim.y * blockDim.z;
  __syncthreads();
  emul_red[emul_tid()] = v;
  __syncthreads();
  float s = -__builtin_inff();
  for (unsigned i = 0; i < n; ++i) s = emul_red[i] > s ? emul_red[i] : s;
  __syncthreads();
  return s;
}

template <typename F>
static void emul_launch(dim3 grid, dim3 block, F kernel) {
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        pthread_barrier_init(&emul_barrier, nullptr, nthreads);
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          ts.emplace_back([=]() {
            threadIdx = uint3_{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = uint3_{bx, by, bz};
            blockDim = block;
            gridDim = grid;
            kernel();
          });
        for (auto& th : ts) th.join();
        pthread_barrier_destroy(&emul_barrier);
      }
}
