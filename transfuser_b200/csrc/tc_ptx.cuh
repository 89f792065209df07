// Inline-PTX wrappers for the Blackwell (sm_100a) async machinery: mbarrier, TMA (cp.async.bulk.tensor),
// TMEM allocation, tcgen05.mma / commit / ld, and UMMA shared-memory / instruction descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(addr), "r"(parity) : "memory");
}

// ---------------- TMA ----------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------- TMEM / tcgen05 ----------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; kind::tf32 (fp32 storage, 10-bit mantissa used) or kind::f16 (bf16/fp16).
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (lane_base + t), columns [col, col+16).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Same, without the wait: issue several loads back to back, then tmem_ld_wait() once.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 vector reduction into global memory (sm_90+): four consecutive floats, 16-byte aligned, one L2 transaction per lane instead of 4
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Epilogue helper: one warp moves its 32 accumulator rows x 32 columns TMEM -> registers -> padded smem (row stride 36 floats:
// conflict-free 16-byte accesses both ways) -> global memory with COALESCED row segments (4 rows per instruction, 128 bytes of fp32
// or 64 bytes of bf16 per row), applying o = alpha*acc + bias (ReLU). `stage` is this warp's private 32 x 36 float buffer.
// `row_off(row)` returns the ELEMENT offset of accumulator row `row` (0..31) in the output (a multiple of 4) or -1 for rows that
// must not be written; columns >= cols_valid are not written. The result goes to out32 (fp32) or, when out16 is given, to out16
// (bf16, round-to-nearest-even) at the same element offsets. atomic: the fp32 values are ADDED to out32 (split-K partial tiles).
// stat_sum / stat_sq (optional, fp32 accumulators in SHARED memory, this chunk's column 0): per-column sum and sum of squares of
// the values written (rows with offset -1 contribute nothing) — the training-mode BatchNorm statistics of a conv / GEMM output
// come out of its epilogue instead of a separate read pass over the tensor. The CTA accumulates over all of its tiles and flushes
// once (stat_flush) with one fp64 global atomic per column: a first version that issued global atomics per warp and chunk
// serialised thousands of same-address atomics on the 72-channel, 70 400-row stage-1 GEMMs and cost more than the pass it removed.
template <class RowOff>
__device__ __forceinline__ void epilogue_chunk32(uint32_t taddr, float* stage, RowOff row_off, int cols_valid, const float* bias,
                                                 float alpha, int relu, int lane, float* out32, __nv_bfloat16* out16 = nullptr,
                                                 float* stat_sum = nullptr, float* stat_sq = nullptr, int atomic = 0) {
  uint32_t r[32];
  tmem_ld16_nowait(taddr, r);
  tmem_ld16_nowait(taddr + 16, r + 16);
  tmem_ld_wait();
  float* srow = stage + lane * 36;
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    *reinterpret_cast<float4*>(srow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                       __uint_as_float(r[j + 3]));
  __syncwarp();
  const int c4 = (lane & 7) * 4, rsub = lane >> 3;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) {
    bv.x = c4 + 0 < cols_valid ? bias[c4 + 0] : 0.f;
    bv.y = c4 + 1 < cols_valid ? bias[c4 + 1] : 0.f;
    bv.z = c4 + 2 < cols_valid ? bias[c4 + 2] : 0.f;
    bv.w = c4 + 3 < cols_valid ? bias[c4 + 3] : 0.f;
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + rsub;
    float4 v = *reinterpret_cast<const float4*>(stage + row * 36 + c4);
    v.x = fmaf(alpha, v.x, bv.x); v.y = fmaf(alpha, v.y, bv.y); v.z = fmaf(alpha, v.z, bv.z); v.w = fmaf(alpha, v.w, bv.w);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    const int64_t off = row_off(row);
    if (off >= 0) {
      if (stat_sum) {
        s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
        s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]); s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
      }
      if (out16) {
        __nv_bfloat16* p = out16 + off + c4;
        if (c4 + 3 < cols_valid) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&lo);
          o.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(p) = o;
        } else {
          if (c4 + 0 < cols_valid) p[0] = __float2bfloat16_rn(v.x);
          if (c4 + 1 < cols_valid) p[1] = __float2bfloat16_rn(v.y);
          if (c4 + 2 < cols_valid) p[2] = __float2bfloat16_rn(v.z);
        }
      } else {
        float* p = out32 + off + c4;
        if (atomic) {                 // split-K partial tile: one 16-byte vector reduction per lane (coalesced 128-byte row segments)
          if (c4 + 3 < cols_valid) {
            red_add_v4(p, v);
          } else {
            if (c4 + 0 < cols_valid) atomicAdd(p + 0, v.x);
            if (c4 + 1 < cols_valid) atomicAdd(p + 1, v.y);
            if (c4 + 2 < cols_valid) atomicAdd(p + 2, v.z);
          }
        } else if (c4 + 3 < cols_valid) {
          *reinterpret_cast<float4*>(p) = v;
        } else {
          if (c4 + 0 < cols_valid) p[0] = v.x;
          if (c4 + 1 < cols_valid) p[1] = v.y;
          if (c4 + 2 < cols_valid) p[2] = v.z;
        }
      }
    }
  }
  if (stat_sum) {   // warp-uniform: lanes with equal (lane & 7) hold the same 4 columns for different rows
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], 8);  s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], 8);
      s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], 16); s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c4 + q < cols_valid) { atomicAdd(stat_sum + c4 + q, s1[q]); atomicAdd(stat_sq + c4 + q, s2[q]); }
    }
  }
  __syncwarp();
}

// The four epilogue warps (128 threads, named barrier 1) clear / flush the CTA's per-column statistics accumulators s_stat[2][n]
// (shared) into stats[2][n] (global fp64: [sum | sum of squares]). t = thread index among the 128 epilogue threads.
__device__ __forceinline__ void stat_clear(float* s_stat, int n, int t) {
  for (int i = t; i < 2 * n; i += 128) s_stat[i] = 0.f;
  asm volatile("bar.sync 1, 128;" ::: "memory");
}
__device__ __forceinline__ void stat_flush(const float* s_stat, double* stats, int n, int t) {
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int i = t; i < 2 * n; i += 128) {
    const float v = s_stat[i];
    if (v != 0.f) atomicAdd(stats + i, (double)v);
  }
}

// ---------------- descriptors ----------------
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (Blackwell).
//   K-major operand : rows of 128 B (the K extent of one stage), 8-row groups 1024 B apart -> SBO = 1024, LBO = 16 (ignored)
//   MN-major operand: 128 B of MN-contiguous elements per K row, 8-row K groups 1024 B apart (SBO), next 128 B MN chunk at LBO
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor (upper 32 bits of idescE): fp32 accumulate, A/B format (0 f16, 1 bf16, 2 tf32), majors, N>>3, M>>4.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t m, uint32_t n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace tc
