// Blackwell (sm_100a) primitives shared by every STAGE kernel: mbarrier, 1-D bulk copies (TMA engine,
// SASS UBLKCP), TMEM allocation, tcgen05.mma issue/commit, tcgen05.ld, and the fp32 -> (hi, lo) fp16
// split that gives the tensor-core GEMMs fp32-grade accuracy (3 MMAs per k-step: Ah*Wh + Al*Wh + Ah*Wl).
//
// Operand layout used everywhere ("chunked K-major, no swizzle"): for an operand with R rows (M or N)
// and K columns of fp16, element (r, k) lives at byte
//        (k / 8) * LBO + r * 16 + (k % 8) * 2
// i.e. a core matrix is 8 rows x 16 bytes = 128 contiguous bytes, 8-row groups are 128 B apart (SBO = 128)
// and the two K-halves of one K=16 MMA are LBO bytes apart.  LBO = R*16 (+16 pad to spread banks when
// threads of one warp store different k-chunks).  One MMA (K=16) advances the start address by 2*LBO.
// The same image read as an MN-major operand (N = the 8-element chunk axis, K = the row axis) uses
// LBO_mn = 128, SBO_mn = LBO and advances by 256 B per K=16.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stg {

// ------------------------------------------------------------------------------------------------
// device-side error flag: every bounded wait that times out records a code here instead of hanging
// ------------------------------------------------------------------------------------------------
__device__ unsigned int g_dev_error = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out: with the default
// hint a waiting thread comes back every few dozen cycles and its polling loop competes with the working warps for issue
// slots (measured: 30 % of all executed instructions of the encoder kernel), so every wait passes an explicit, long hint.
#define STG_MBAR_SUSPEND_NS 20000u
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(STG_MBAR_SUSPEND_NS)
      : "memory");
  return ok;
}
// Bounded wait (4 s of wall clock).  On timeout the error flag is raised and the caller carries on with garbage (the host
// checks the flag) instead of hanging the GPU.  The %globaltimer read between two attempts doubles as a short back-off:
// A/B on B200 (same box, configs[1]): this loop 14.16 ms/step, a bare attempt loop 14.67, __nanosleep(20..100) between
// attempts 14.61; the value of the suspend hint (500 ns .. 1 ms) makes no difference, its presence does; an additional
// exponential __nanosleep back-off (32 .. 2048 ns) changes nothing either.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned code = 1) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) {
      atomicMax(&g_dev_error, 0x100u + code);
      return;
    }
  }
}

// The same bounded wait for waiters that are far from any critical path (a warp that waits most of a tile for its next
// hand-over): sleeps between attempts, so its polling does not compete with the working warps for issue slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned code = 1, unsigned sleep_ns = 256) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) {
      atomicMax(&g_dev_error, 0x100u + code);
      return;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// proxy fences + bulk copy (global -> shared, completes on an mbarrier)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() {  // generic-proxy smem writes -> visible to async proxy (MMA/TMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMEM
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {  // one full warp
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "power of 2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 32 consecutive columns: thread t of the warp gets row (lane base + t), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%"
      "29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 32-bit, 16 consecutive columns, registers -> TMEM (thread t of the warp writes row lane-base + t)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 32 lanes x 32-bit, 8 consecutive columns (= 16 packed fp16: one K=16 slice of a TMEM-resident A operand)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(a.x), "r"(a.y),
               "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle, sm_100 version field = 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// instruction descriptor: kind::f16, A/B fp16, D fp32, M=128, given N; b_mn = 1 -> B is MN-major
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n, uint32_t a_mn = 0, uint32_t b_mn = 0) {
  return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// One lane of a CONVERGED warp (the role warps run their control flow with all 32 lanes and elect the issuer): under a
// `lane == 0` branch the compiler cannot keep descriptors in uniform registers and wraps every tcgen05 instruction in a
// lane-by-lane waterfall loop (~130 cycles per MMA measured, twice the tensor time of a 128x128x16 MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is K-major in tensor memory, row r of the tile in lane r, two fp16 per 32-bit column
// (one K=16 step reads 8 columns starting at a_tmem)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// split-fp16 helpers
// ------------------------------------------------------------------------------------------------
struct HL8 {  // 8 consecutive k values: 16 B of hi and 16 B of lo
  uint4 hi, lo;
};
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  // packed conversions (F2FP.F16.F32.PACK_AB) instead of two scalar F2F each
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Packed fp32 arithmetic (FFMA2 / FADD2, sm_100): one instruction for two lanes of a float2.  The 3-register FFMA issues
// every other cycle per scheduler, so the element-wise phases of these kernels are bound by fma-pipe instruction count.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float2 max2(float2 a, float b) { return make_float2(fmaxf(a.x, b), fmaxf(a.y, b)); }
// split of a float2: hi = fp16x2(v), lo = fp16x2(v - float(hi)) with the subtraction as one packed FMA.  (A truncating
// split -- hi = v masked to 11 significant bits on the ALU pipe -- was A/B'd on B200: 0.8 % slower end to end.)
__device__ __forceinline__ void split2p(float2 v, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __float22half2_rn(v);
  const float2 f = __half22float2(h);
  const __half2 l = __float22half2_rn(fma2(f, make_float2(-1.f, -1.f), v));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ HL8 split8(const float* v) {
  HL8 o;
  split2p(make_float2(v[0], v[1]), o.hi.x, o.lo.x);
  split2p(make_float2(v[2], v[3]), o.hi.y, o.lo.y);
  split2p(make_float2(v[4], v[5]), o.hi.z, o.lo.z);
  split2p(make_float2(v[6], v[7]), o.hi.w, o.lo.w);
  return o;
}

// issue the 3-term product for one K=16 step.  a_hi/a_lo/b_hi/b_lo are descriptors already advanced to this step.
__device__ __forceinline__ void umma_split3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                            uint64_t b_lo, uint32_t idesc, uint32_t accumulate) {
  umma_f16(d_tmem, a_lo, b_hi, idesc, accumulate);
  umma_f16(d_tmem, a_hi, b_lo, idesc, 1u);
  umma_f16(d_tmem, a_hi, b_hi, idesc, 1u);
}

// 2^x on the SFU (MUFU.EX2, max relative error 2^-22); arguments are <= 0 in the softmax
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 256-bit read-only global load (sm_100): 32 contiguous, 32-byte aligned bytes in one request
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
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

}  // namespace stg
