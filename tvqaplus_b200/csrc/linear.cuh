// Row-wise Linear on tcgen05:  y = LN_out( relu?( LN_in( l2norm?(x) ) W^T + b ) )
//
// Persistent CTAs, one 128-row tile at a time.  Warp roles (416 threads):
//   warps 0-3   epilogue: TMEM -> registers (thread per row), bias / ReLU / LayerNorm, store y
//   warps 4-11  producers: per-row statistics pre-pass, then per K-chunk load fp32 x (128-bit, row-coalesced), apply
//               the prologue, split into fp16 (hi, lo) and store the A operand image; the first producer thread also
//               launches the bulk copy of the matching weight chunk (one cp.async.bulk per stage)
//   warp  12    MMA issuer (one lane): 3 tcgen05.mma per K=16 step (split-fp16); tcgen05.commit frees the stage
// Pipelines: full[s]/empty[s] per smem stage, acc_full[b]/acc_empty[b] per TMEM accumulator buffer (two buffers when
// 2*NPAD fits in 512 columns, so the epilogue of tile i overlaps the MMAs of tile i+1).
//
// Algorithmic work per row: 2*K*N FLOP (x3 executed for the split), reads K*4 B, writes N*4 B.
#pragma once
#include "pack.cuh"
#include "umma.cuh"

namespace stg {

template <int NPAD, int KC, int STAGES>
struct LinearCfg {
  static constexpr int K8 = KC / 8;
  static constexpr uint32_t A_PAD = (KC == 32) ? 32u : 16u;  // spreads the 16-B stores of one quarter-warp over all banks
  static constexpr uint32_t LBO_A = 128u * 16u + A_PAD;
  static constexpr uint32_t A_PART = K8 * LBO_A;
  static constexpr uint32_t A_STAGE = 2u * A_PART;
  static constexpr uint32_t LBO_B = NPAD * 16u;
  static constexpr uint32_t B_PART = K8 * LBO_B;
  static constexpr uint32_t B_STAGE = 2u * B_PART;
  static constexpr uint32_t STAGE_BYTES = (A_STAGE + B_STAGE + 127u) & ~127u;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024u;
  static constexpr int ACC_BUFS = (2 * NPAD <= 512 && NPAD <= 256) ? 2 : 1;
  static constexpr uint32_t ACC_STRIDE = NPAD <= 128 ? 128u : 256u;
  static constexpr uint32_t TMEM_COLS = (ACC_BUFS == 2) ? 2u * ACC_STRIDE : 512u;
  static constexpr int N0 = NPAD <= 256 ? NPAD : ((NPAD / 2 + 15) / 16 * 16);  // first MMA N; second = NPAD - N0
  static constexpr int THREADS = 416;
};

struct LinearDev {
  const float* x;
  const uint8_t* w;
  const float* bias;
  const float* ln_in_g;
  const float* ln_in_b;
  const float* ln_out_g;
  const float* ln_out_b;
  float* y;
  long long m;
  int k, n, n_chunks;
  int l2norm_in, relu;
};

template <int NPAD, int KC, int STAGES>
__global__ void __launch_bounds__(416, 1) linear_kernel(LinearDev p) {
  using Cfg = LinearCfg<NPAD, KC, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float row_inv[128], row_mean[128], row_rstd[128];

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 256);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  const long long n_tiles = (p.m + 127) / 128;
  const bool has_pro = (p.ln_in_g != nullptr) || p.l2norm_in;

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue: thread per row
    uint32_t tile_it = 0;
    const bool ln = (p.ln_out_g != nullptr);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t b = tile_it % Cfg::ACC_BUFS, use = tile_it / Cfg::ACC_BUFS;
      mbar_wait(&acc_full[b], use & 1u, 11);
      tc_fence_after();
      const int r = warp * 32 + lane;
      const long long gr = tile * 128 + r;
      const uint32_t trow = tbase + b * Cfg::ACC_STRIDE + ((warp * 32u) << 16);
      float mean = 0.f, rstd = 1.f;
      if (ln) {
        float s = 0.f;
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            if (c < p.n) {
              float t = v[j] + __ldg(p.bias + c);
              if (p.relu) t = fmaxf(t, 0.f);
              s += t;
            }
          }
        }
        mean = s / static_cast<float>(p.n);
        float q = 0.f;
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            if (c < p.n) {
              float t = v[j] + __ldg(p.bias + c);
              if (p.relu) t = fmaxf(t, 0.f);
              t -= mean;
              q += t * t;
            }
          }
        }
        rstd = 1.f / sqrtf(q / static_cast<float>(p.n) + 1e-5f);
      }
      for (int c0 = 0; c0 < NPAD; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j;
          if (c < p.n) {
            float t = v[j] + __ldg(p.bias + c);
            if (p.relu) t = fmaxf(t, 0.f);
            if (ln) t = (t - mean) * rstd * __ldg(p.ln_out_g + c) + __ldg(p.ln_out_b + c);
            v[j] = t;
          }
        }
        if (gr < p.m) {
          float* dst = p.y + gr * p.n + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (c0 + j + 4 <= p.n) {
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int jj = j; jj < j + 4; ++jj)
                if (c0 + jj < p.n) dst[jj] = v[jj];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ producers
    const uint32_t ptid = tid - 128, pwarp = warp - 4;
    uint32_t it = 0;  // global chunk counter (stage / phase tracking)
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row0 = tile * 128;
      if (has_pro) {
        asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's last chunk has consumed the old statistics
        // per-row statistics, exact two-pass (mean, then centred second moment), one warp per row
        for (int rr = 0; rr < 16; ++rr) {
          const int r = pwarp * 16 + rr;
          const long long gr = row0 + r;
          float inv = 1.f, mean = 0.f, rstd = 1.f;
          if (gr < p.m) {
            const float4* xr = reinterpret_cast<const float4*>(p.x + gr * p.k);
            const int nv = p.k >> 2;
            if (p.l2norm_in) {
              float ss = 0.f;
              for (int i = lane; i < nv; i += 32) {
                float4 v = __ldg(xr + i);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
              }
              ss = warp_sum(ss);
              inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            }
            if (p.ln_in_g != nullptr) {
              float s = 0.f;
              for (int i = lane; i < nv; i += 32) {
                float4 v = __ldg(xr + i);
                s += (v.x * inv + v.y * inv) + (v.z * inv + v.w * inv);
              }
              mean = warp_sum(s) / static_cast<float>(p.k);
              float q = 0.f;
              for (int i = lane; i < nv; i += 32) {
                float4 v = __ldg(xr + i);
                float a = v.x * inv - mean, b = v.y * inv - mean, c = v.z * inv - mean, d = v.w * inv - mean;
                q += (a * a + b * b) + (c * c + d * d);
              }
              rstd = 1.f / sqrtf(warp_sum(q) / static_cast<float>(p.k) + 1e-5f);
            }
          }
          if (lane == 0) {
            row_inv[r] = inv;
            row_mean[r] = mean;
            row_rstd[r] = rstd;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }

      for (int c = 0; c < p.n_chunks; ++c, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        mbar_wait(&empty_bar[s], (use & 1u) ^ 1u, 10);
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_STAGE;
        if (ptid == 0) {
          mbar_arrive_expect_tx(&full_bar[s], Cfg::B_STAGE);
          bulk_g2s(sB, p.w + static_cast<size_t>(c) * Cfg::B_STAGE, Cfg::B_STAGE, &full_bar[s]);
        }
        for (int item = ptid; item < 128 * Cfg::K8; item += 256) {
          const int r = item / Cfg::K8, k8 = item % Cfg::K8;
          const int k0 = c * KC + k8 * 8;
          const long long gr = row0 + r;
          float v[8];
          if (gr < p.m && k0 + 8 <= p.k) {
            const float4* src = reinterpret_cast<const float4*>(p.x + gr * p.k + k0);
            float4 a = __ldg(src), b = __ldg(src + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (gr < p.m && k0 + j < p.k) ? __ldg(p.x + gr * p.k + k0 + j) : 0.f;
          }
          if (has_pro) {
            const float inv = row_inv[r], mean = row_mean[r], rstd = row_rstd[r];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t = v[j] * inv;
              if (p.ln_in_g != nullptr) {
                const int kk = k0 + j;
                t = (kk < p.k) ? (t - mean) * rstd * __ldg(p.ln_in_g + kk) + __ldg(p.ln_in_b + kk) : 0.f;
              }
              v[j] = t;
            }
          }
          HL8 hl = split8(v);
          uint8_t* dst = sA + k8 * Cfg::LBO_A + r * 16;
          *reinterpret_cast<uint4*>(dst) = hl.hi;
          *reinterpret_cast<uint4*>(dst + Cfg::A_PART) = hl.lo;
        }
        fence_async_smem();
        if (ptid != 0) mbar_arrive(&full_bar[s]);
      }
    }
  } else if (lane == 0) {
    // -------------------------------------------------------------------- MMA issuer
    uint32_t it = 0, tile_it = 0;
    const uint32_t idesc0 = make_idesc_f16(Cfg::N0);
    const uint32_t idesc1 = make_idesc_f16(NPAD - Cfg::N0 > 0 ? NPAD - Cfg::N0 : 16);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t b = tile_it % Cfg::ACC_BUFS, ause = tile_it / Cfg::ACC_BUFS;
      mbar_wait(&acc_empty[b], (ause & 1u) ^ 1u, 12);
      tc_fence_after();
      const uint32_t dcol = tbase + b * Cfg::ACC_STRIDE;
      for (int c = 0; c < p.n_chunks; ++c, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        mbar_wait(&full_bar[s], use & 1u, 13);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + Cfg::A_STAGE;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          const uint64_t a_hi = make_smem_desc(sA + ks * 2 * Cfg::LBO_A, Cfg::LBO_A, 128);
          const uint64_t a_lo = make_smem_desc(sA + Cfg::A_PART + ks * 2 * Cfg::LBO_A, Cfg::LBO_A, 128);
          const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
          {
            const uint64_t b_hi = make_smem_desc(sB + ks * 2 * Cfg::LBO_B, Cfg::LBO_B, 128);
            const uint64_t b_lo = make_smem_desc(sB + Cfg::B_PART + ks * 2 * Cfg::LBO_B, Cfg::LBO_B, 128);
            umma_split3(dcol, a_hi, a_lo, b_hi, b_lo, idesc0, acc);
          }
          if (NPAD > Cfg::N0) {
            const uint64_t b_hi = make_smem_desc(sB + ks * 2 * Cfg::LBO_B + Cfg::N0 * 16, Cfg::LBO_B, 128);
            const uint64_t b_lo = make_smem_desc(sB + Cfg::B_PART + ks * 2 * Cfg::LBO_B + Cfg::N0 * 16, Cfg::LBO_B, 128);
            umma_split3(dcol + Cfg::N0, a_hi, a_lo, b_hi, b_lo, idesc1, acc);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_full[b]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<Cfg::TMEM_COLS>(tbase);
}

}  // namespace stg
