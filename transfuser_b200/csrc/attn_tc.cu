// Fused multi-head self-attention on the tcgen05 tensor cores (forward and backward) for the GPT fusion blocks
// (reference: SelfAttention.forward, transfuser.py:510-527: q k^T / sqrt(hs) -> softmax -> attn_drop -> @ v; T = 174 tokens,
// 4 heads, head sizes 18 / 54 / 144 / 378). Nothing of size T x T ever reaches HBM: scores live in TMEM, probabilities go
// through shared memory as the bf16 A operand of the second product, the softmax is one thread per row on the TMEM lanes.
//
// One CTA = one 128-row tile of one (batch, head). Three modes share the kernel body:
//   FWD   rows = queries : S = Q K^T (TMEM) -> P = softmax(scale S), Pd = dropout(P) (bf16, smem) -> O = Pd V ; saves the
//                          row log-sum-exp for the backward pass
//   BWD_Q rows = queries : S = Q K^T, dPd = dO V^T (both TMEM) -> dS = scale P (dPd drop - D) (smem) -> dQ = dS K
//   BWD_K rows = keys    : S^T = K Q^T, dPd^T = V dO^T (TMEM) -> Pd^T (smem) -> dV = Pd^T dO ; dS^T (smem) -> dK = dS^T Q
// (P is recomputed from the saved log-sum-exp; D_i = sum_d dO_id O_id comes from a small pre-pass.)
// Operands are staged by all threads into 128B-swizzled shared memory (the head slices of the packed [B*T, 3C] q|k|v buffer
// start at 4-byte-aligned columns, which rules out TMA boxes) and described to tcgen05.mma with the same K-major / MN-major
// shared-memory descriptors as gemm_tc.cu. fp32 or bf16 inputs (converted while staging), fp32 accumulation.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRowsX = 128;                 // tile rows (UMMA M)
constexpr int kRowsY = 192;                 // padded rows of the column operand (T <= 192)
constexpr int kXBytes = kRowsX * 128;       // one 64-column bf16 chunk of the row operand
constexpr int kYBytes = kRowsY * 128;       // one 64-column bf16 chunk of the column operand
constexpr int kStageBytes = 2 * (kXBytes + kYBytes);   // X1 | Y1 | X2 | Y2 (the second pair is unused in FWD)
constexpr int kMatBytes = 3 * kXBytes;      // the 128 x 192 bf16 probability / score-gradient matrix (3 K chunks of 64)
constexpr int kZBuf = 2 * kYBytes;          // two 64-wide chunks of the MN-major B operand of the second product
constexpr int kSmemStages = 2 * kStageBytes;                     // 163840: two streaming stages; later three Z buffers
constexpr int kSmemTotal = kSmemStages + kMatBytes + 2048 + 1024;  // + barriers / row statistics + alignment slack
constexpr uint32_t kAcc1 = 0, kAcc2 = 192, kOut = 384;           // TMEM columns

enum { MODE_FWD = 0, MODE_BWD_Q = 1, MODE_BWD_K = 2 };

struct Params {
  const void* qkv;      // [B*T, 3C] q | k | v
  const void* dy;       // [B*T, C]  dO (backward)
  float* out32_a;       // FWD: y [B*T, C]; BWD: dqkv [B*T, 3C]   (fp32, may be null)
  __nv_bfloat16* out16_a;  // same, bf16 (may be null)
  float* lse;           // [B, nh, T]
  const float* dsum;    // [B, nh, T]  D (backward)
  const uint64_t* seed_dev;
  uint64_t seed_off;
  int B, T, nh, hs, C;
  int in_bf16, dy_bf16;
  float scale, p_drop;
  int mode;             // MODE_FWD, or 1 = backward (BWD_K CTAs first, then BWD_Q)
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cheap counter-based keep/drop decision for the attention dropout (forward and backward of THIS kernel regenerate it)
__device__ __forceinline__ float drop_scale(uint32_t seed_lo, uint32_t seed_hi, uint32_t idx, float p, float keep_scale) {
  uint32_t x = idx * 0x9E3779B1u + seed_lo;
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= seed_hi;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  const float u = (float)(x >> 8) * (1.0f / 16777216.0f);
  return u < p ? 0.f : keep_scale;
}

// Stages rows [row0, row0 + rows_pad) x columns [col0, col0 + 64) of a row-major matrix (leading dimension ld elements) as one
// 64-column bf16 chunk in the SWIZZLE_128B layout: row r at (r / 8) * 1024 + (r % 8) * 128, 16-byte unit u at (u ^ (r % 8)) * 16.
// Rows >= rows_valid and columns >= cols_valid are zero-filled. cols_valid and col0 are even.
template <typename TIn>
__device__ __forceinline__ void stage_chunk(uint8_t* dst, const TIn* __restrict__ src, int64_t ld, int64_t row0, int rows_valid,
                                            int rows_pad, int col0, int cols_valid, int tid, int nthreads) {
  for (int idx = tid; idx < rows_pad * 8; idx += nthreads) {
    const int r = idx >> 3, u = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int c = u * 8;
    if (r < rows_valid && c < cols_valid) {
      const TIn* p = src + (row0 + r) * ld + col0 + c;
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + 2 * j < cols_valid) {
          if (sizeof(TIn) == 4) {
            const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p) + 2 * j);
            __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
            w[j] = *reinterpret_cast<uint32_t*>(&b);
          } else {
            w[j] = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p) + 2 * j);
          }
        } else {
          w[j] = 0u;
        }
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(dst + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)) = v;
  }
}

// bf16 source: the same chunk through cp.async (global -> shared without registers, many copies in flight per thread; the copies
// of one call are NOT waited for here — cp_async_wait_all() before the proxy fence). VEC = bytes per copy (16 / 8 / 4): the largest
// power of two that divides the source addresses (the head slices of the packed q|k|v rows are only 4-byte aligned in general).
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VEC>
__device__ __forceinline__ void stage_chunk_async(uint8_t* dst, const __nv_bfloat16* __restrict__ src, int64_t ld, int64_t row0,
                                                  int rows_valid, int rows_pad, int col0, int cols_valid, int tid, int nthreads) {
  constexpr int PIECES = 16 / VEC, EPP = VEC / 2;      // copies per 16-byte unit, elements per copy
  for (int idx = tid; idx < rows_pad * 8; idx += nthreads) {
    const int r = idx >> 3, u = idx & 7;
    uint8_t* d = dst + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4);
    const int c = u * 8;
    if (r < rows_valid && c < cols_valid) {
      const __nv_bfloat16* p = src + (row0 + r) * ld + col0 + c;
#pragma unroll
      for (int j = 0; j < PIECES; ++j) {
        if (c + (j + 1) * EPP <= cols_valid) {
          const uint32_t da = tc::smem_u32(d + j * VEC);
          if (VEC == 16)     asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(p + j * EPP) : "memory");
          else if (VEC == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(da), "l"(p + j * EPP) : "memory");
          else               asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(da), "l"(p + j * EPP) : "memory");
        } else {                                       // partially valid copy (VEC > 4 only): element pairs, the rest zero
#pragma unroll
          for (int e = 0; e < EPP; e += 2) {
            uint32_t w = 0u;
            if (c + j * EPP + e < cols_valid) w = *reinterpret_cast<const uint32_t*>(p + j * EPP + e);
            *reinterpret_cast<uint32_t*>(d + j * VEC + e * 2) = w;
          }
        }
      }
    } else {
      *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// Issues the staging of one chunk; bf16 sources are asynchronous (see above), fp32 sources are converted synchronously.
__device__ __forceinline__ void stage_any(uint8_t* dst, const void* src, int is_bf16, int64_t ld, int64_t row0, int rows_valid,
                                          int rows_pad, int col0, int cols_valid, int tid, int nthreads) {
  if (is_bf16) {
    const __nv_bfloat16* s16 = (const __nv_bfloat16*)src;
    const uintptr_t a = reinterpret_cast<uintptr_t>(s16 + col0) | (uintptr_t)(ld * 2);
    if ((a & 15) == 0)     stage_chunk_async<16>(dst, s16, ld, row0, rows_valid, rows_pad, col0, cols_valid, tid, nthreads);
    else if ((a & 7) == 0) stage_chunk_async<8>(dst, s16, ld, row0, rows_valid, rows_pad, col0, cols_valid, tid, nthreads);
    else                   stage_chunk_async<4>(dst, s16, ld, row0, rows_valid, rows_pad, col0, cols_valid, tid, nthreads);
  } else {
    stage_chunk<float>(dst, (const float*)src, ld, row0, rows_valid, rows_pad, col0, cols_valid, tid, nthreads);
  }
}

// byte offset of element (row r, column j) of the 128 x 192 K-major bf16 matrix (3 chunks of 64 columns)
__device__ __forceinline__ uint32_t mat_off(int r, int j) {
  const int c = j >> 6, u = (j & 63) >> 3, e = j & 7;
  return (uint32_t)(c * kXBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4) + e * 2);
}

__global__ void __launch_bounds__(kThreads, 1) attn_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* mat = smem + kSmemStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemStages + kMatBytes);   // [0,1]: stage s consumed; [2]: second product done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* stat_a = reinterpret_cast<float*>(smem + kSmemStages + kMatBytes + 64);  // [192] lse of the column tokens (BWD_K)
  float* stat_b = stat_a + kRowsY;                                                 // [192] D   of the column tokens (BWD_K)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, nh = p.nh, hs = p.hs, C = p.C;
  const int tiles = (T + kRowsX - 1) / kRowsX;
  int blk = blockIdx.x, mode = MODE_FWD;
  if (p.mode != MODE_FWD) {
    const int half = p.B * nh * tiles;
    mode = blk < half ? MODE_BWD_K : MODE_BWD_Q;
    if (blk >= half) blk -= half;
  }
  const int tile = blk % tiles, bh = blk / tiles, h = bh % nh, b = bh / nh;
  const int m0 = tile * kRowsX;
  const int rows_valid = min(kRowsX, T - m0);
  const int ncols = (T + 15) & ~15;                     // UMMA N of the score products (and K extent of the second product)
  const int64_t tok0 = (int64_t)b * T;                  // first token row of this batch element
  const int64_t ld3 = 3 * (int64_t)C;
  const bool bwd = mode != MODE_FWD;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  uint32_t ph_stage[2] = {0u, 0u}, ph_out = 0u;

  // column offsets (elements) of the operands inside their matrices
  const int colq = h * hs, colk = C + h * hs, colv = 2 * C + h * hs, coly = h * hs;
  // first product(s): acc1 = X1 Y1^T (rows x ncols), acc2 = X2 Y2^T, streamed over 64-wide chunks of the head dimension
  //   FWD / BWD_Q: X1 = Q tile, Y1 = K;  BWD_Q also X2 = dO tile, Y2 = V.   BWD_K: X1 = K tile, Y1 = Q, X2 = V tile, Y2 = dO.
  const int x1col = mode == MODE_BWD_K ? colk : colq, y1col = mode == MODE_BWD_K ? colq : colk;
  const int nchunks = (hs + 63) / 64;
  auto issue_loads = [&](int c) {
    uint8_t* st = smem + (c & 1) * kStageBytes;
    const int cv = min(64, hs - c * 64);
    stage_any(st, p.qkv, p.in_bf16, ld3, tok0 + m0, rows_valid, kRowsX, x1col + c * 64, cv, tid, kThreads);
    stage_any(st + kXBytes, p.qkv, p.in_bf16, ld3, tok0, T, kRowsY, y1col + c * 64, cv, tid, kThreads);
    if (mode == MODE_BWD_Q) {
      stage_any(st + kXBytes + kYBytes, p.dy, p.dy_bf16, C, tok0 + m0, rows_valid, kRowsX, coly + c * 64, cv, tid, kThreads);
      stage_any(st + 2 * kXBytes + kYBytes, p.qkv, p.in_bf16, ld3, tok0, T, kRowsY, colv + c * 64, cv, tid, kThreads);
    } else if (mode == MODE_BWD_K) {
      stage_any(st + kXBytes + kYBytes, p.qkv, p.in_bf16, ld3, tok0 + m0, rows_valid, kRowsX, colv + c * 64, cv, tid, kThreads);
      stage_any(st + 2 * kXBytes + kYBytes, p.dy, p.dy_bf16, C, tok0, T, kRowsY, coly + c * 64, cv, tid, kThreads);
    }
    cp_async_commit();
  };
  issue_loads(0);
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    if (c + 1 < nchunks) {
      // the copies of chunk c + 1 go out before chunk c is consumed: its stage was last read by the MMAs of chunk c - 1
      if (c >= 1) { tc::mbar_wait(&bars[s ^ 1], ph_stage[s ^ 1]); ph_stage[s ^ 1] ^= 1u; }
      issue_loads(c + 1);
      cp_async_wait_group<1>();                        // everything but the newest group (chunk c + 1) has landed
    } else {
      cp_async_wait_all();
    }
    uint8_t* st = smem + s * kStageBytes;
    const int cv = min(64, hs - c * 64);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::make_idesc(1u, 0u, 0u, kRowsX, (uint32_t)ncols);
      const uint32_t sx1 = tc::smem_u32(st), sy1 = sx1 + kXBytes, sx2 = sy1 + kYBytes, sy2 = sx2 + kXBytes;
      const int ksteps = (cv + 15) / 16;
      for (int k = 0; k < ksteps; ++k) {
        const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
        tc::umma_f16(tmem + kAcc1, tc::make_smem_desc(sx1 + k * 32, 16, 1024), tc::make_smem_desc(sy1 + k * 32, 16, 1024), idesc, acc);
        if (bwd)
          tc::umma_f16(tmem + kAcc2, tc::make_smem_desc(sx2 + k * 32, 16, 1024), tc::make_smem_desc(sy2 + k * 32, 16, 1024), idesc, acc);
      }
      tc::umma_commit(&bars[s]);
    }
  }
  // drain: every outstanding stage commit (at most two) is waited for exactly once, oldest first
  // (iteration c waited for chunk c - 1 whenever it prefetched chunk c + 1: chunks 0 .. nchunks - 3 are done with)
  if (nchunks >= 2) { const int s = nchunks & 1; tc::mbar_wait(&bars[s], ph_stage[s]); ph_stage[s] ^= 1u; }
  { const int s = (nchunks - 1) & 1; tc::mbar_wait(&bars[s], ph_stage[s]); ph_stage[s] ^= 1u; }
  tc::fence_after_sync();

  const uint32_t seed_lo = (uint32_t)((p.seed_dev ? *p.seed_dev : 0ull) + p.seed_off);
  const uint32_t seed_hi = (uint32_t)(((p.seed_dev ? *p.seed_dev : 0ull) + p.seed_off) >> 32);
  const float keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  const float sl2 = p.scale * 1.4426950408889634f;
  const int r = tid & 127;                               // this thread's tile row = its TMEM lane (warps 0-3; warps 4-7 mirror)
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const bool row_ok = r < rows_valid;
  const int tok = m0 + r;                                // token index of this row inside the batch element
  float inv_sum = 1.f;

  if (mode == MODE_BWD_K) {                              // per-column statistics of all query tokens
    for (int i = tid; i < kRowsY; i += kThreads) {
      stat_a[i] = i < T ? p.lse[(int64_t)bh * T + i] * 1.4426950408889634f : 0.f;
      stat_b[i] = i < T ? p.dsum[(int64_t)bh * T + i] : 0.f;
    }
    __syncthreads();
  }

  // second product(s): out[rows x hs] = M Z, M (bf16, smem, K-major, K = ncols tokens) from the element-wise stage, Z MN-major.
  //   FWD: M = Pd, Z = V -> y.   BWD_Q: M = dS, Z = K -> dQ.   BWD_K: round 0 M = Pd^T, Z = dO -> dV; round 1 M = dS^T, Z = Q -> dK.
  const int nrounds = mode == MODE_BWD_K ? 2 : 1;
  const int npass = (hs + 127) / 128;
  for (int round = 0; round < nrounds; ++round) {
    const void* zsrc; int z_bf16, zcol; int64_t zld;
    int ocol;                                            // output column (in out*_a, leading dimension old)
    int64_t old;
    if (mode == MODE_FWD)        { zsrc = p.qkv; z_bf16 = p.in_bf16; zld = ld3; zcol = colv; ocol = coly; old = C; }
    else if (mode == MODE_BWD_Q) { zsrc = p.qkv; z_bf16 = p.in_bf16; zld = ld3; zcol = colk; ocol = colq; old = ld3; }
    else if (round == 0)         { zsrc = p.dy;  z_bf16 = p.dy_bf16; zld = C;   zcol = coly; ocol = colv; old = ld3; }
    else                         { zsrc = p.qkv; z_bf16 = p.in_bf16; zld = ld3; zcol = colq; ocol = colk; old = ld3; }

    // ---- element-wise stage (warps 0-3: one thread per row, straight out of TMEM) while warps 4-7 stage the first Z buffer ----
    if (warp < 4) {
      if (mode == MODE_FWD) {
        float mx = -INFINITY;
        for (int j0 = 0; j0 < ncols; j0 += 16) {
          float v[16];
          tc::tmem_ld16(tmem + lane_addr + kAcc1 + j0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j0 + j < T) mx = fmaxf(mx, v[j]);
        }
        float sum = 0.f;
        const uint32_t ibase = (uint32_t)(((int64_t)bh * T + tok) * T);
        for (int j0 = 0; j0 < ncols; j0 += 16) {
          float v[16];
          tc::tmem_ld16(tmem + lane_addr + kAcc1 + j0, v);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float e0 = 0.f, e1 = 0.f;
            if (j0 + j < T)     { e0 = exp2f((v[j] - mx) * sl2);     sum += e0; if (p.p_drop > 0.f) e0 *= drop_scale(seed_lo, seed_hi, ibase + j0 + j, p.p_drop, keep); }
            if (j0 + j + 1 < T) { e1 = exp2f((v[j + 1] - mx) * sl2); sum += e1; if (p.p_drop > 0.f) e1 *= drop_scale(seed_lo, seed_hi, ibase + j0 + j + 1, p.p_drop, keep); }
            __nv_bfloat162 bb = __floats2bfloat162_rn(e0, e1);
            w[j >> 1] = *reinterpret_cast<uint32_t*>(&bb);
          }
          *reinterpret_cast<uint4*>(mat + mat_off(r, j0)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(mat + mat_off(r, j0 + 8)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
        inv_sum = 1.f / sum;
        if (row_ok && p.lse) p.lse[(int64_t)bh * T + tok] = mx * p.scale + logf(sum);
      } else if (mode == MODE_BWD_Q) {
        const float lse2 = row_ok ? p.lse[(int64_t)bh * T + tok] * 1.4426950408889634f : 0.f;
        const float dsum = row_ok ? p.dsum[(int64_t)bh * T + tok] : 0.f;
        const uint32_t ibase = (uint32_t)(((int64_t)bh * T + tok) * T);
        for (int j0 = 0; j0 < ncols; j0 += 16) {
          float sv[16], gv[16];
          tc::tmem_ld16(tmem + lane_addr + kAcc1 + j0, sv);
          tc::tmem_ld16(tmem + lane_addr + kAcc2 + j0, gv);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float d0 = 0.f, d1 = 0.f;
            if (j0 + j < T) {
              const float pr = exp2f(sv[j] * sl2 - lse2);
              const float g = p.p_drop > 0.f ? gv[j] * drop_scale(seed_lo, seed_hi, ibase + j0 + j, p.p_drop, keep) : gv[j];
              d0 = p.scale * pr * (g - dsum);
            }
            if (j0 + j + 1 < T) {
              const float pr = exp2f(sv[j + 1] * sl2 - lse2);
              const float g = p.p_drop > 0.f ? gv[j + 1] * drop_scale(seed_lo, seed_hi, ibase + j0 + j + 1, p.p_drop, keep) : gv[j + 1];
              d1 = p.scale * pr * (g - dsum);
            }
            __nv_bfloat162 bb = __floats2bfloat162_rn(d0, d1);
            w[j >> 1] = *reinterpret_cast<uint32_t*>(&bb);
          }
          *reinterpret_cast<uint4*>(mat + mat_off(r, j0)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(mat + mat_off(r, j0 + 8)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      } else {
        // rows = keys (token tok), columns = queries i: S^T and dPd^T
        for (int j0 = 0; j0 < ncols; j0 += 16) {
          float sv[16], gv[16];
          tc::tmem_ld16(tmem + lane_addr + kAcc1 + j0, sv);
          tc::tmem_ld16(tmem + lane_addr + kAcc2 + j0, gv);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float o[2] = {0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = j0 + j + e;                  // query token
              if (i < T) {
                const float pr = exp2f(sv[j + e] * sl2 - stat_a[i]);
                const float ds = p.p_drop > 0.f ? drop_scale(seed_lo, seed_hi, (uint32_t)(((int64_t)bh * T + i) * T + tok), p.p_drop, keep) : 1.f;
                o[e] = round == 0 ? pr * ds : p.scale * pr * (gv[j + e] * ds - stat_b[i]);
              }
            }
            __nv_bfloat162 bb = __floats2bfloat162_rn(o[0], o[1]);
            w[j >> 1] = *reinterpret_cast<uint32_t*>(&bb);
          }
          *reinterpret_cast<uint4*>(mat + mat_off(r, j0)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(mat + mat_off(r, j0 + 8)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc::fence_before_sync();
    } else {
      // Z buffer 0 = head-dimension columns [0, 128) as two 64-wide chunks of kRowsY token rows (MN-major B operand)
      for (int q = 0; q < 2; ++q) {
        const int cv = max(0, min(64, hs - q * 64));
        if (cv > 0) stage_any(smem + q * kYBytes, zsrc, z_bf16, zld, tok0, T, kRowsY, zcol + q * 64, cv, tid - 128, 128);
      }
      cp_async_wait_all();
    }
    fence_async_smem();
    __syncthreads();

    for (int ps = 0; ps < npass; ++ps) {
      uint8_t* zb = smem + (ps % 3) * kZBuf;
      const int dvalid = min(128, hs - ps * 128);         // head-dimension columns of this pass
      const int nmma = dvalid > 64 ? 128 : 64;
      if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc(1u, 0u, 1u, kRowsX, (uint32_t)nmma);
        const uint32_t sm = tc::smem_u32(mat), sz = tc::smem_u32(zb);
        const int ksteps = ncols / 16;
        for (int k = 0; k < ksteps; ++k)
          tc::umma_f16(tmem + kOut, tc::make_smem_desc(sm + (k >> 2) * kXBytes + (k & 3) * 32, 16, 1024),
                       tc::make_smem_desc(sz + k * 2048, kYBytes, 1024), idesc, k > 0 ? 1u : 0u);
        tc::umma_commit(&bars[2]);
      }
      if (ps + 1 < npass) {                               // prefetch the next pass's Z while the tensor core works
        uint8_t* zn = smem + ((ps + 1) % 3) * kZBuf;
        for (int q = 0; q < 2; ++q) {
          const int cv = max(0, min(64, hs - (ps + 1) * 128 - q * 64));
          if (cv > 0) stage_any(zn + q * kYBytes, zsrc, z_bf16, zld, tok0, T, kRowsY, zcol + (ps + 1) * 128 + q * 64, cv, tid, kThreads);
        }
        cp_async_wait_all();
        fence_async_smem();
      }
      tc::mbar_wait(&bars[2], ph_out); ph_out ^= 1u;
      tc::fence_after_sync();
      if (warp < 4) {
        const float mul = mode == MODE_FWD ? inv_sum : 1.f;
        float* o32 = p.out32_a ? p.out32_a + (tok0 + tok) * old + ocol + ps * 128 : nullptr;
        __nv_bfloat16* o16 = p.out16_a ? p.out16_a + (tok0 + tok) * old + ocol + ps * 128 : nullptr;
        for (int d0 = 0; d0 < dvalid; d0 += 16) {
          float v[16];
          tc::tmem_ld16(tmem + lane_addr + kOut + d0, v);
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              if (d0 + j < dvalid) {                      // (dvalid is even)
                const float a = v[j] * mul, bq = v[j + 1] * mul;
                if (o32) *reinterpret_cast<float2*>(o32 + d0 + j) = make_float2(a, bq);
                if (o16) *reinterpret_cast<__nv_bfloat162*>(o16 + d0 + j) = __floats2bfloat162_rn(a, bq);
              }
            }
          }
        }
        tc::fence_before_sync();
      }
      __syncthreads();
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

// D[b, h, i] = sum_d dO[b*T + i, h*hs + d] * O[b*T + i, h*hs + d]: one warp per (token row, head); the same pass writes the bf16
// copy of dO that the backward kernel stages with cp.async (hs even).
template <typename TY>
__global__ void __launch_bounds__(256) attn_dsum_kernel(const float* __restrict__ dy, const TY* __restrict__ y, float* __restrict__ dsum,
                                                        __nv_bfloat16* __restrict__ dy16, int B, int T, int nh, int hs) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * T * nh) return;
  const int h = (int)(w % nh);
  const int64_t row = w / nh;
  const int b = (int)(row / T), i = (int)(row % T);
  const int64_t base = row * (int64_t)nh * hs + h * hs;
  float s = 0.f;
  for (int d = lane * 2; d < hs; d += 64) {
    const float2 g = *reinterpret_cast<const float2*>(dy + base + d);
    float y0, y1;
    if (sizeof(TY) == 4) {
      const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(y) + base + d);
      y0 = v.x; y1 = v.y;
    } else {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(y) + base + d);
      y0 = __low2float(v); y1 = __high2float(v);
    }
    s = fmaf(g.x, y0, fmaf(g.y, y1, s));
    if (dy16) *reinterpret_cast<__nv_bfloat162*>(dy16 + base + d) = __floats2bfloat162_rn(g.x, g.y);
  }
  s = warp_sum(s);
  if (lane == 0) dsum[((int64_t)b * nh + h) * T + i] = s;
}

int launch_attn(const Params& p, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal) != cudaSuccess) {
      tfb_set_last_error("cudaFuncSetAttribute(smem) failed for attn_tc_kernel");
      return TFB_ERR_DRIVER;
    }
    attr_done = true;
  }
  const int tiles = (p.T + kRowsX - 1) / kRowsX;
  const int grid = p.B * p.nh * tiles * (p.mode == MODE_FWD ? 1 : 2);
  attn_tc_kernel<<<grid, kThreads, kSmemTotal, stream>>>(p);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

}  // namespace

// Fused attention forward on the packed q|k|v buffer: y[b*T + i, h*hs : (h+1)*hs] = dropout(softmax(scale q k^T)) v for every
// (b, h). qkv: [B*T, 3*nh*hs] fp32 (qkv_bf16 = 0) or bf16 (1). y32 / y16: [B*T, nh*hs] outputs (either may be null).
// lse: [B, nh, T] row log-sum-exp of the scaled scores, saved for tfb_attn_bwd_tc. Dropout: keep-probability 1 - p_drop, mask =
// hash(*seed_dev + seed_off, ((b*nh + h)*T + i)*T + j), regenerated in backward. T <= 192, hs even.
TFB_API int tfb_attn_fwd_tc(const void* qkv, int qkv_bf16, int B, int T, int nh, int hs, float* y32, void* y16, float* lse, float scale,
                            float p_drop, const uint64_t* seed_dev, uint64_t seed_off, cudaStream_t stream) {
  TFB_REQUIRE(qkv && (y32 || y16) && lse && B > 0 && T > 0 && T <= kRowsY && nh > 0 && hs > 0 && hs % 2 == 0);
  TFB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && p_drop >= 0.f && p_drop < 1.f);
  Params p{};
  p.qkv = qkv; p.in_bf16 = qkv_bf16; p.out32_a = y32; p.out16_a = (__nv_bfloat16*)y16; p.lse = lse;
  p.seed_dev = seed_dev; p.seed_off = seed_off; p.B = B; p.T = T; p.nh = nh; p.hs = hs; p.C = nh * hs;
  p.scale = scale; p.p_drop = p_drop; p.mode = MODE_FWD;
  return launch_attn(p, stream);
}

// Backward of tfb_attn_fwd_tc: dqkv[B*T, 3C] (fp32 and / or bf16; every element written) from dy[B*T, C] (fp32), the forward's
// q|k|v input, its output y (fp32, or bf16 when y_bf16: for D = rowsum(dy * y)) and lse. Scratch: dsum [B, nh, T] fp32 and
// dy16 [B*T, C] bf16 (the bf16 copy of dy the kernel stages; written by the pre-pass).
TFB_API int tfb_attn_bwd_tc(const void* qkv, int qkv_bf16, const float* dy, const void* y, int y_bf16, const float* lse, float* dsum,
                            void* dy16, int B, int T, int nh, int hs, float* dqkv32, void* dqkv16, float scale, float p_drop,
                            const uint64_t* seed_dev, uint64_t seed_off, cudaStream_t stream) {
  TFB_REQUIRE(qkv && dy && y && lse && dsum && dy16 && (dqkv32 || dqkv16) && B > 0 && T > 0 && T <= kRowsY && nh > 0 && hs > 0 && hs % 2 == 0);
  TFB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(dy16) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0);
  const int64_t warps = (int64_t)B * T * nh;
  if (y_bf16) attn_dsum_kernel<__nv_bfloat16><<<(int)((warps + 7) / 8), 256, 0, stream>>>(dy, (const __nv_bfloat16*)y, dsum, (__nv_bfloat16*)dy16, B, T, nh, hs);
  else        attn_dsum_kernel<float><<<(int)((warps + 7) / 8), 256, 0, stream>>>(dy, (const float*)y, dsum, (__nv_bfloat16*)dy16, B, T, nh, hs);
  TFB_CHECK_LAUNCH();
  Params p{};
  p.qkv = qkv; p.in_bf16 = qkv_bf16; p.dy = dy16; p.dy_bf16 = 1; p.out32_a = dqkv32; p.out16_a = (__nv_bfloat16*)dqkv16;
  p.lse = const_cast<float*>(lse); p.dsum = dsum;
  p.seed_dev = seed_dev; p.seed_off = seed_off; p.B = B; p.T = T; p.nh = nh; p.hs = hs; p.C = nh * hs;
  p.scale = scale; p.p_drop = p_drop; p.mode = 1;
  return launch_attn(p, stream);
}
