// UMMA convention probe: one CTA, operands supplied as raw shared-memory images by the host, descriptors
// built from host-supplied fields.  Used by tests/test_umma_probe.py to pin the layout conventions
// (K-major / MN-major no-swizzle images, LBO/SBO roles, TMEM lane/column mapping, bulk-copy path) that the
// production kernels rely on.
#pragma once
#include "umma.cuh"

namespace stg {

struct ProbeArgs {
  const uint8_t* a_img;  // bytes copied verbatim into smem at offset 0
  const uint8_t* b_img;  // bytes copied verbatim into smem at offset a_bytes (1024-aligned)
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, a_step;  // descriptor fields (bytes) and start-address advance per K=16 step
  uint32_t b_lbo, b_sbo, b_step;
  uint32_t idesc;
  uint32_t k_steps;
  uint32_t n;          // MMA N (multiple of 16, <= 256)
  uint32_t use_bulk;   // 1: B image arrives through cp.async.bulk + mbarrier, 0: generic stores
  uint32_t a_tmem;     // 1: A comes from TENSOR MEMORY: a_img is row-major fp16 [128][16*k_steps]; thread r packs its row two
                       //    k values per 32-bit column and stores it with tcgen05.st at columns [256 + 8*ks, +8)
  float* d_out;        // [128][n]
  uint32_t repeat;     // the k_steps MMAs are issued this many times (accumulating)
  long long* cycles_out;
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_tx;
  __shared__ uint32_t tmem_slot;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023u) & ~1023u);

  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tx, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
  if (p.use_bulk) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_tx, p.b_bytes);
      bulk_g2s(sb, p.b_img, p.b_bytes, &bar_tx);
    }
  } else {
    for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16)
      *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(p.b_img + i);
  }
  if (p.a_tmem) {
    const uint32_t r = warp * 32 + lane;
    for (uint32_t ks = 0; ks < p.k_steps; ++ks) {
      const uint4* src = reinterpret_cast<const uint4*>(p.a_img + (static_cast<size_t>(r) * p.k_steps + ks) * 32);
      const uint4 q0 = src[0], q1 = src[1];
      tmem_st8(tbase + ((warp * 32u) << 16) + 256 + ks * 8, q0, q1);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  fence_async_smem();
  __syncthreads();

  if (tid == 0) {
    if (p.use_bulk) mbar_wait(&bar_tx, 0, 2);
    tc_fence_after();
    long long t_start = clock64();
    if (p.repeat > 1 && p.k_steps == 8) {
      // throughput mode: descriptors precomputed, 8 MMAs per unrolled loop body (what a lean issue loop can reach)
      uint64_t ad[8], bd[8];
#pragma unroll
      for (uint32_t ks = 0; ks < 8; ++ks) {
        ad[ks] = make_smem_desc(smem_u32(sa) + ks * p.a_step, p.a_lbo, p.a_sbo);
        bd[ks] = make_smem_desc(smem_u32(sb) + ks * p.b_step, p.b_lbo, p.b_sbo);
      }
      const uint32_t ta = tbase + 256, idesc = p.idesc;
      t_start = clock64();
      if (p.a_tmem) {
#pragma unroll 1
        for (uint32_t rep = 0; rep < p.repeat; ++rep) {
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) umma_f16_ts(tbase, ta + ks * 8, bd[ks], idesc, 1u);
        }
      } else {
#pragma unroll 1
        for (uint32_t rep = 0; rep < p.repeat; ++rep) {
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) umma_f16(tbase, ad[ks], bd[ks], idesc, 1u);
        }
      }
    } else {
      for (uint32_t rep = 0; rep < p.repeat; ++rep) {
        for (uint32_t ks = 0; ks < p.k_steps; ++ks) {
          uint64_t ad = make_smem_desc(smem_u32(sa) + ks * p.a_step, p.a_lbo, p.a_sbo);
          uint64_t bd = make_smem_desc(smem_u32(sb) + ks * p.b_step, p.b_lbo, p.b_sbo);
          if (p.a_tmem) umma_f16_ts(tbase, tbase + 256 + ks * 8, bd, p.idesc, (ks > 0 || rep > 0) ? 1u : 0u);
          else umma_f16(tbase, ad, bd, p.idesc, (ks > 0 || rep > 0) ? 1u : 0u);
        }
      }
    }
    umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0, 3);
    if (p.cycles_out != nullptr) *p.cycles_out = clock64() - t_start;

  }
  mbar_wait(&bar_mma, 0, 3);
  tc_fence_after();

  const uint32_t row = warp * 32 + lane;
  for (uint32_t c = 0; c < p.n; c += 16) {
    float v[16];
    tmem_ld16(tbase + ((warp * 32u) << 16) + c, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d_out[row * p.n + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace stg
