// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX), and the 3xTF32 operand split.
//
// Operand layout used by every tensor-core kernel here: K-major, no swizzle.  A tile of ROWS rows
// is stored as 16-byte k-chunks (4 tf32 each):  byte offset of (row r, chunk kc) =
//     kc * LBO + r * 16        with LBO = ROWS * 16 + 16,
// i.e. core matrices of 8 rows x 16 B are contiguous (SBO = 128 B between 8-row groups) and the
// 16 B of padding per chunk spreads a warp's 128-bit stores over all bank groups.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival + `bytes` of pending async-copy traffic (completed by cp.async.bulk ... complete_tx)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// TMA bulk copy (no tensor map): `bytes` contiguous bytes global -> shared, 16-byte aligned on both
// sides, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- proxies / fences ---------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() {  // generic-proxy smem writes -> async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
// called by ONE full warp; ncols: power of two >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// 32 lanes x 32 columns: thread i of the warp receives row (lane base + i), columns c .. c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM, 32 lanes x 16 columns: thread i of the warp writes row (lane base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
// registers -> TMEM, 32 lanes x 8 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- UMMA descriptors -----------------------------------------------------------------------------
// K-major, no swizzle (see the layout at the top).  All byte quantities are multiples of 16.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);  // version 1 (sm_100), no swizzle
  return ((uint64_t)hi << 32) | lo;
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, accumulate flag as a compile-time constant (no predicate setup in the issue loop), and the
// descriptor of k-step `ks` derived from the k-step-0 descriptor by one add: the start-address
// field (bits 0-13, units of 16 B) advances by 2 chunks = 2 * LBO bytes per tf32 k-step of 8.
template <bool ACC>
__device__ __forceinline__ void mma_tf32_c(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc) {
  if (ACC) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
  }
}
// A operand in TMEM (M = 128 lanes x 8 columns of tf32 per instruction), B in shared memory
template <bool ACC>
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc,
                                            uint32_t idesc) {
  if (ACC) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc)
        : "memory");
  }
}
// The same, issued by ONE elected lane of a fully converged warp: the whole warp runs the issue
// loop (warp-uniform control flow), so the compiler emits elect + UTCHMMA back to back instead of
// the per-instruction convergence loop it wraps around an MMA inside `if (lane == 0)`.
template <bool ACC>
__device__ __forceinline__ void mma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a,
                                                  uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}
// both operands in shared memory, elected lane of a converged warp; runtime accumulate flag
__device__ __forceinline__ void mma_tf32_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A in TMEM, B in shared memory, elected lane of a converged warp, runtime accumulate flag
__device__ __forceinline__ void mma_tf32_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__host__ __device__ constexpr uint64_t desc_kstep(int lbo_bytes) { return (uint64_t)((2 * lbo_bytes) >> 4); }

// all MMAs issued so far by this thread arrive on `bar` when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi keeps
// FP32-level accuracy on the tensor cores (error ~ 2^-21 relative per product).
// Round to TF32 (10-bit mantissa), nearest with ties away from zero — what cvt.rna.tf32.f32 returns
// for every finite input, done on the bit pattern: add half a TF32 ulp, clear the 13 low bits (a
// carry out of the mantissa bumps the exponent, as rounding should).  cvt.rna.tf32.f32 itself is
// not a single SASS instruction on sm_100 (~5 with its NaN/Inf handling); this is 2 integer ops,
// and the operand split below runs once per element of every tensor-core operand tile.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
  hi.x = tf32_rn(x.x); hi.y = tf32_rn(x.y); hi.z = tf32_rn(x.z); hi.w = tf32_rn(x.w);
  // the remainder is left as it is: the tensor core reads only the TF32 bits of an operand, i.e.
  // truncates it, which costs at most 2^-10 of a term that is itself below 2^-11 of x — the same
  // order as the lo x lo product the 3xTF32 scheme drops anyway — and saves two of the five ALU
  // operations the split spends per element
  lo.x = x.x - hi.x; lo.y = x.y - hi.y;
  lo.z = x.z - hi.z; lo.w = x.w - hi.w;
}

// Coalesced row-major store of a warp's 32 x 32 accumulator chunk: thread `lane` holds row `lane`
// (32 consecutive columns in v[]).  The chunk is transposed through a per-warp shared-memory
// buffer (32 x 33 floats) so that every store instruction writes 4 rows x 128 contiguous bytes.
// `rowptr(r)` returns the output pointer of the chunk's row r at its first column (or nullptr).
template <class RowPtr>
__device__ __forceinline__ void store_chunk_coalesced(float* wbuf, const float (&v)[32], int lane,
                                                      int ncols_valid, bool vec_ok, RowPtr rowptr) {
#pragma unroll
  for (int j = 0; j < 32; ++j) wbuf[lane * 33 + j] = v[j];
  __syncwarp();
  const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    float* dst = rowptr(r);
    const float a = wbuf[r * 33 + c4], b = wbuf[r * 33 + c4 + 1], c = wbuf[r * 33 + c4 + 2],
                d = wbuf[r * 33 + c4 + 3];
    if (dst) {
      if (vec_ok && c4 + 3 < ncols_valid) {
        *reinterpret_cast<float4*>(dst + c4) = make_float4(a, b, c, d);
      } else {
        if (c4 + 0 < ncols_valid) dst[c4 + 0] = a;
        if (c4 + 1 < ncols_valid) dst[c4 + 1] = b;
        if (c4 + 2 < ncols_valid) dst[c4 + 2] = c;
        if (c4 + 3 < ncols_valid) dst[c4 + 3] = d;
      }
    }
  }
  __syncwarp();
}

// byte size of one operand tile (ROWS rows x KC4 16-byte chunks) in the layout above
__host__ __device__ constexpr int tile_lbo(int rows) { return rows * 16 + 16; }
__host__ __device__ constexpr int tile_bytes(int rows, int kc4) { return kc4 * tile_lbo(rows); }

}  // namespace tc
