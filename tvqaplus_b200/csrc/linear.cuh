// Row-wise Linear on tcgen05:  y = LN_out( relu?( LN_in( l2norm?(x) ) W^T + b ) )
//
// The input-side normalisations are folded into the GEMM so that the A operand is the RAW row, streamed from HBM exactly
// once:
//      LN_in(x^) W^T + b = rstd * inv * ((x - c) (W*gamma)^T - (mean(x) - c) * cs) + cb ,   x^ = x * inv,
//      inv = 1/max(|x|_2, 1e-12) (or 1),   cs = (W*gamma) 1,  cb = W beta + b      (stage_pack_weight_ln)
// c = first element of the row: the tensor cores see the SHIFTED row, so a constant row yields exactly beta and a large
// common offset costs neither GEMM nor variance precision.  mean / rstd come from sum(x - c) and sum((x - c)^2), which
// the producers accumulate while they convert the row.
//
// Persistent CTAs, one 128-row tile at a time.  Warp roles (416 threads):
//   warps 0-3   epilogue: TMEM -> registers (thread per row): folded LayerNorm / bias / ReLU / exact two-pass LN_out,
//               then a 32x32 transpose through shared memory so that every global store covers whole 128-byte lines
//   warps 4-11  producers: each thread owns the same (row, 8-column slot) pairs for every K chunk; its global loads run
//               PREF chunks ahead in registers (HBM latency), values are split into fp16 (hi, lo) and stored as the
//               A operand image; the first producer thread also launches the bulk copy of the matching weight chunk
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
  static constexpr int TS = 36;                                  // fp32 row stride of a warp's 32x32 transpose buffer
  static constexpr uint32_t TR_BYTES = 4u * 32u * TS * 4u;       // 18432: one buffer per epilogue warp
  static constexpr uint32_t VEC_BYTES = 4u * NPAD * 4u;          // cs | cb (or bias) | ln_out gamma | ln_out beta
  static constexpr uint32_t OFF_TR = STAGES * STAGE_BYTES;
  static constexpr uint32_t OFF_VEC = OFF_TR + TR_BYTES;
  static constexpr uint32_t SMEM_BYTES = OFF_VEC + VEC_BYTES + 1024u;
  static constexpr int ACC_BUFS = (2 * NPAD <= 512 && NPAD <= 256) ? 2 : 1;
  static constexpr uint32_t ACC_STRIDE = NPAD <= 128 ? 128u : 256u;
  static constexpr uint32_t TMEM_COLS = (ACC_BUFS == 2) ? 2u * ACC_STRIDE : 512u;
  static constexpr int N0 = NPAD <= 256 ? NPAD : ((NPAD / 2 + 15) / 16 * 16);  // first MMA N; second = NPAD - N0
  static constexpr int THREADS = 416;
  static constexpr int ITEMS = 128 * K8 / 256;  // (row, 8-column slot) pairs per producer thread per chunk
  static constexpr int PREF = (KC == 32) ? 4 : 2;  // chunks of global loads in flight per producer thread
};

struct LinearDev {
  const float* x;
  const uint8_t* w;
  const float* bias;      // used when cs == nullptr
  const float* cs;        // LN_in folded: column sums of W*gamma  (nullptr: no LayerNorm in front)
  const float* cb;        // LN_in folded: W beta + b
  const float* ln_out_g;
  const float* ln_out_b;
  float* y;
  long long m;
  int k, n, n_chunks;
  int l2norm_in, relu;
  int out_blocked;        // 1: y is written as [m / rows_per_block][n / 4][rows_per_block] float4 (see cqa.cuh, Za)
  int rows_per_block;
};

template <int NPAD, int KC, int STAGES>
__global__ void __launch_bounds__(416, 1) linear_kernel(LinearDev p) {
  using Cfg = LinearCfg<NPAD, KC, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float row_s1[2][128], row_s2[2][128], row_c[2][128];  // per-row sum(x-c), sum((x-c)^2), c; double-buffered over tiles

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible
  float* vec = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
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
  for (int i = tid; i < NPAD; i += Cfg::THREADS) {  // epilogue vectors, zero beyond n
    const bool in = i < p.n;
    vec[i] = (in && p.cs != nullptr) ? p.cs[i] : 0.f;
    vec[NPAD + i] = in ? (p.cs != nullptr ? p.cb[i] : p.bias[i]) : 0.f;
    vec[2 * NPAD + i] = (in && p.ln_out_g != nullptr) ? p.ln_out_g[i] : 1.f;
    vec[3 * NPAD + i] = (in && p.ln_out_b != nullptr) ? p.ln_out_b[i] : 0.f;
  }
  if (warp == 12) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  const long long n_tiles = (p.m + 127) / 128;
  const bool need_stats = (p.cs != nullptr) || p.l2norm_in;

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue: thread per row
    uint32_t tile_it = 0;
    const bool ln = (p.ln_out_g != nullptr);
    float* tr = reinterpret_cast<float*>(smem + Cfg::OFF_TR) + warp * 32 * Cfg::TS;
    const float inv_k = 1.f / static_cast<float>(p.k), inv_n = 1.f / static_cast<float>(p.n);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t b = tile_it % Cfg::ACC_BUFS, use = tile_it / Cfg::ACC_BUFS;
      mbar_wait(&acc_full[b], use & 1u, 11);
      tc_fence_after();
      const int r = warp * 32 + lane;
      const long long row0 = tile * 128 + warp * 32;  // first global row of this warp
      const uint32_t trow = tbase + b * Cfg::ACC_STRIDE + ((warp * 32u) << 16);
      // input-side scalars of my row:  pre[c] = ga * acc[c] + gm * cs[c] + cb[c]   (acc is the GEMM of the SHIFTED row x - c)
      float ga = 1.f, gm = 0.f;
      if (need_stats) {
        const uint32_t sb = tile_it & 1u;
        const float s1 = row_s1[sb][r], s2 = row_s2[sb][r], c = row_c[sb][r];
        float inv = 1.f;
        if (p.l2norm_in) {  // sum(x^2) = s2 + 2 c s1 + K c^2
          const float ss = fmaxf(s2 + 2.f * c * s1 + static_cast<float>(p.k) * c * c, 0.f);
          inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        }
        ga = inv;
        if (p.cs != nullptr) {
          const float d = s1 * inv_k;                                   // mean(x) - c
          const float var = fmaxf(s2 * inv_k - d * d, 0.f) * inv * inv;  // variance of x^ (shift invariant)
          const float rstd = 1.f / sqrtf(var + 1e-5f);
          ga = rstd * inv;
          gm = -rstd * d * inv;
        }
      }
      float mean = 0.f, rstd = 1.f;
      if (ln) {  // exact two-pass statistics of the activated row
        float s = 0.f;
        for (int c0 = 0; c0 < NPAD; c0 += 32) {
          float v[32];
          tmem_ld16(trow + c0, v);
          if (c0 + 16 < NPAD) tmem_ld16(trow + c0 + 16, v + 16);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c < p.n) {
              float t = fmaf(ga, v[j], fmaf(gm, vec[c], vec[NPAD + c]));
              if (p.relu) t = fmaxf(t, 0.f);
              s += t;
            }
          }
        }
        mean = s * inv_n;
        float q = 0.f;
        for (int c0 = 0; c0 < NPAD; c0 += 32) {
          float v[32];
          tmem_ld16(trow + c0, v);
          if (c0 + 16 < NPAD) tmem_ld16(trow + c0 + 16, v + 16);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c < p.n) {
              float t = fmaf(ga, v[j], fmaf(gm, vec[c], vec[NPAD + c]));
              if (p.relu) t = fmaxf(t, 0.f);
              t -= mean;
              q = fmaf(t, t, q);
            }
          }
        }
        rstd = 1.f / sqrtf(q * inv_n + 1e-5f);
      }
      for (int c0 = 0; c0 < NPAD; c0 += 32) {
        float v[32];
        tmem_ld16(trow + c0, v);
        if (c0 + 16 < NPAD) tmem_ld16(trow + c0 + 16, v + 16);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = c0 + j;
          float t = 0.f;
          if (c < NPAD) {
            t = fmaf(ga, v[j], fmaf(gm, vec[c], vec[NPAD + c]));
            if (p.relu) t = fmaxf(t, 0.f);
            if (ln) t = fmaf((t - mean) * rstd, vec[2 * NPAD + c], vec[3 * NPAD + c]);
          }
          v[j] = t;
        }
        // 32 rows x 32 columns through the warp's transpose buffer: afterwards lane = (row % 4 group, 4-column slot)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(tr + lane * Cfg::TS + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int c4 = lane & 7, rsub = lane >> 3;
        const int col = c0 + c4 * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rsub;
          const long long gr = row0 + rr;
          const float4 o = *reinterpret_cast<const float4*>(tr + rr * Cfg::TS + c4 * 4);
          if (gr < p.m && col < p.n) {
            if (p.out_blocked) {
              const long long blk = gr / p.rows_per_block;
              const int rl = static_cast<int>(gr - blk * p.rows_per_block);
              *reinterpret_cast<float4*>(p.y + ((blk * (p.n >> 2) + (col >> 2)) * p.rows_per_block + rl) * 4) = o;
            } else if ((p.n & 3) == 0) {
              *reinterpret_cast<float4*>(p.y + gr * p.n + col) = o;
            } else {
              const float ov[4] = {o.x, o.y, o.z, o.w};
              for (int e = 0; e < 4; ++e)
                if (col + e < p.n) p.y[gr * p.n + col + e] = ov[e];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ producers
    const uint32_t ptid = tid - 128;
    constexpr int K8 = Cfg::K8, NI = Cfg::ITEMS, PREF = Cfg::PREF;
    const int k8 = ptid % K8;
    const int rbase = ptid / K8;                 // my rows of a tile: rbase + i * (256 / K8)
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long total = my_tiles * p.n_chunks;  // flat (tile, chunk) sequence of this CTA

    float4 buf[PREF][NI][2];
    long long ptl = 0;  // (tile, chunk) of the next prefetch
    int pc = 0;
    auto issue = [&](float4 (&dst)[NI][2]) {
      const long long row0 = (blockIdx.x + ptl * gridDim.x) * 128;
      const int k0 = pc * KC + k8 * 8;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const long long gr = row0 + rbase + i * (256 / K8);
        const float4* src = reinterpret_cast<const float4*>(p.x + gr * p.k + k0);
        const bool ok = ptl < my_tiles && gr < p.m;
        dst[i][0] = (ok && k0 + 4 <= p.k) ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[i][1] = (ok && k0 + 8 <= p.k) ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++pc == p.n_chunks) {
        pc = 0;
        ++ptl;
      }
    };
#pragma unroll
    for (int d = 0; d < PREF; ++d) issue(buf[d]);

    uint32_t it = 0;  // flat chunk counter (stage / phase tracking)
    long long tl = 0;
    int c = 0;
    float cshift[NI], s1[NI], s2[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) cshift[i] = s1[i] = s2[i] = 0.f;
    for (long long g0 = 0; g0 < total; g0 += PREF) {
#pragma unroll
      for (int d = 0; d < PREF; ++d) {
        if (g0 + d < total) {  // flat chunk g0 + d = (tile tl, chunk c) sits in register buffer d
          const uint32_t s = it % STAGES, use = it / STAGES;
          uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_STAGE;
          const int kleft = p.k - (c * KC + k8 * 8);  // valid columns of my 8-column slot in this chunk
          HL8 hl[NI];
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            float v[8] = {buf[d][i][0].x, buf[d][i][0].y, buf[d][i][0].z, buf[d][i][0].w,
                          buf[d][i][1].x, buf[d][i][1].y, buf[d][i][1].z, buf[d][i][1].w};
            if (need_stats) {
              if (c == 0) {
                // with LN_in the row is shifted by its first element (held by the k8 == 0 lane) BEFORE it meets the tensor
                // cores: a constant row then gives exactly beta and a large common offset costs no GEMM precision
                const float first = __shfl_sync(0xffffffffu, v[0], lane & ~(K8 - 1));
                cshift[i] = (p.cs != nullptr) ? first : 0.f;
                s1[i] = 0.f;
                s2[i] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[j] = (j < kleft) ? v[j] - cshift[i] : 0.f;
                s1[i] += v[j];
                s2[i] = fmaf(v[j], v[j], s2[i]);
              }
            }
            hl[i] = split8(v);
          }
          issue(buf[d]);  // refill this register buffer: the loads fly while the next chunks are converted
          mbar_wait(&empty_bar[s], (use & 1u) ^ 1u, 10);
          if (ptid == 0) {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::B_STAGE);
            bulk_g2s(sB, p.w + static_cast<size_t>(c) * Cfg::B_STAGE, Cfg::B_STAGE, &full_bar[s]);
          }
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            uint8_t* dst = sA + k8 * Cfg::LBO_A + (rbase + i * (256 / K8)) * 16;
            *reinterpret_cast<uint4*>(dst) = hl[i].hi;
            *reinterpret_cast<uint4*>(dst + Cfg::A_PART) = hl[i].lo;
          }
          if (need_stats && c == p.n_chunks - 1) {  // the row statistics travel with the last chunk of the tile
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              float a = s1[i], q = s2[i];
#pragma unroll
              for (int o = 1; o < K8; o <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
              }
              if (k8 == 0) {
                const int r = rbase + i * (256 / K8);
                row_s1[tl & 1][r] = a;
                row_s2[tl & 1][r] = q;
                row_c[tl & 1][r] = cshift[i];
              }
            }
          }
          fence_async_smem();
          if (ptid != 0) mbar_arrive(&full_bar[s]);
          ++it;
          if (++c == p.n_chunks) {
            c = 0;
            ++tl;
          }
        }
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
