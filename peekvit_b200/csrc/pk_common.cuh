// peekvit_b200 — shared device/host helpers for the sm_100a kernels.
// Inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM) and
// the small math helpers every kernel uses.  sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#define PK_OK 0
#define PK_ERR_INVALID (-1)
#define PK_ERR_CUDA (-2)
#define PK_ERR_UNSUPPORTED (-3)
#define PK_ERR_DEVICE_TIMEOUT (-4)

namespace pk {

// ---------------------------------------------------------------- host-side error plumbing
void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define PK_CHECK_CUDA(expr)                                   \
  do {                                                        \
    int _pk_rc = ::pk::check_cuda((expr), #expr);             \
    if (_pk_rc != PK_OK) return _pk_rc;                       \
  } while (0)

#define PK_REQUIRE(cond, ...)                                 \
  do {                                                        \
    if (!(cond)) {                                            \
      ::pk::set_last_error(__VA_ARGS__);                      \
      return PK_ERR_INVALID;                                  \
    }                                                         \
  } while (0)

int num_sms();

// Device-side watchdog flag (one word of device memory owned by the library): a bounded
// mbarrier wait that expires stores a code there instead of hanging the GPU (a hung box is a
// strike on the shared pool).  Read and reset through pk_device_flag().
unsigned int* device_flag_ptr();

#if defined(__CUDACC__)
// ---------------------------------------------------------------- generic device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// One lane of a converged warp (warp-uniform predicate input, divergent result).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ int warp_min_i32(int v) { return __reduce_min_sync(0xffffffffu, v); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// hi = bf16(a), lo = bf16(a - hi) for a pair of values: the two-term split of the bf16x2 arithmetic mode
__device__ __forceinline__ void split2_pack(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a0, a1);
  lo = pack_bf16(a0 - __uint_as_float(hi << 16), a1 - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Exact-GELU (erf form, reference blocks.py:82 F.gelu default) with the Abramowitz–Stegun
// 7.1.26 rational erf (|abs err| <= 1.5e-7, far below the bf16 rounding of the result).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erfc_z = poly * __expf(-z * z);          // erfc(z), z >= 0
  const float cdf = x >= 0.f ? 1.0f - 0.5f * erfc_z : 0.5f * erfc_z;
  return x * cdf;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// try_wait with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes or `ns`
// nanoseconds pass, instead of coming back at once and re-issuing the probe.  Measured on the ragged attention kernel (ncu,
// profiles/r02): the plain form returned immediately, and the softmax warps of the idle TMEM region spent a quarter of the
// SM's issued instructions polling s_full -- on the very schedulers the other region's softmax warps were running on.
#ifndef PK_MBAR_SUSPEND_NS
#define PK_MBAR_SUSPEND_NS 2000
#endif
__device__ __forceinline__ uint32_t mbar_try_wait_suspend(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(PK_MBAR_SUSPEND_NS)
      : "memory");
  return done;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: returns false (and raises the device flag) after ~2 s instead of hanging.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, unsigned int* flag, unsigned int code) {
  if (mbar_try_wait(bar, parity)) return true;
  unsigned long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait_suspend(bar, parity)) return true;
#ifdef PK_MBAR_SLEEP_NS
    if (it > 2) __nanosleep(PK_MBAR_SLEEP_NS);
#endif
    if ((it & 0x3ffu) == 0) {
      if (*(volatile unsigned int*)flag != 0u) return false;
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) {
        atomicCAS(flag, 0u, code);
        return false;
      }
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// Pull a 2-D tile into L2 ahead of its load (no shared memory, no completion tracking).
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1)
               : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-bounds rows/columns are clipped.
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
// 2-D tiled reduction shared -> global: global[tile] += smem[tile] (f32 add performed at L2).
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- packed fp32x2 math (FFMA2 on sm_100)
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Exact-GELU of two values at once on the packed FFMA2 pipe with ONE MUFU op per value (the GEMM epilogue
// shares the MUFU with nothing else, but at 4 lanes/clk/SMSP two ops per element were 4096 cycles per
// 128x256 tile).  Abramowitz–Stegun 7.1.28:  erfc(z) = 1 / (1 + a1 z + ... + a6 z^6)^16,  |err| <= 3e-7, z >= 0.
//   h = 0.5*erfc(|x|/sqrt2);  gelu(x) = x * (0.5 + sign(x) * (0.5 - h))
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
  const uint64_t ax = f2_pack(fabsf(x0), fabsf(x1));
  const uint64_t z = f2_mul(ax, f2_pack(0.70710678118654752f, 0.70710678118654752f));
  uint64_t s = f2_fma(f2_pack(0.0000430638f, 0.0000430638f), z, f2_pack(0.0002765672f, 0.0002765672f));
  s = f2_fma(s, z, f2_pack(0.0001520143f, 0.0001520143f));
  s = f2_fma(s, z, f2_pack(0.0092705272f, 0.0092705272f));
  s = f2_fma(s, z, f2_pack(0.0422820123f, 0.0422820123f));
  s = f2_fma(s, z, f2_pack(0.0705230784f, 0.0705230784f));
  s = f2_fma(s, z, f2_pack(1.0f, 1.0f));
  s = f2_mul(s, s);
  s = f2_mul(s, s);
  s = f2_mul(s, s);
  s = f2_mul(s, s);                                  // (..)^16; overflows to +inf for |x| > ~25 -> erfc = 0
  float s0, s1;
  f2_unpack(s, s0, s1);
  const uint64_t r = f2_pack(rcp_approx(s0), rcp_approx(s1));                                      // erfc(z)
  const uint64_t g = f2_fma(r, f2_pack(-0.5f, -0.5f), f2_pack(0.5f, 0.5f));                        // 0.5 - h  (>= 0)
  // x*(0.5 + sign(x)*g) = 0.5*x + |x|*g : no sign transfer needed
  const uint64_t o = f2_fma(ax, g, f2_mul(f2_pack(x0, x1), f2_pack(0.5f, 0.5f)));
  f2_unpack(o, x0, x1);
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> f32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread t of the warp receives lane (base_lane + t), 32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (rows of 64 bf16 = 128 B, 8-row
// swizzle atoms 1024 B apart).  Field layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;               // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;       // SBO = 1024 B between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;               // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;               // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, shape M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// same with A = B = IEEE half (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int b_mn_major = 0) {
  return (1u << 4) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
#endif  // __CUDACC__

// ---------------------------------------------------------------- host: TMA descriptors
// 2-D row-major bf16 tensor (rows x cols, leading dimension ld elements) with a
// (box_rows x 64)-element box and 128-byte swizzle.  Cached by key.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols);
// General form: elem_bytes 2 (bf16) or 4 (f32); swizzle_bytes 128 or 64 (= box_cols * elem_bytes).
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes);

}  // namespace pk
