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
// Persistent CTAs, one 128-row tile at a time.  Warp roles (416 threads; 544 with a single accumulator, see EPI_WARPS):
//   warps 9-..  epilogue: TMEM -> registers (thread per row): folded LayerNorm / bias / ReLU / exact two-pass LN_out,
//               then a 32x32 transpose through shared memory so that every global store covers whole 128-byte lines
//   warps 0-7   producers: each thread owns the same (row, 8-column slot) pairs for every K chunk; its global loads run
//               PREF chunks ahead in registers (HBM latency), values are split into fp16 (hi, lo) and stored as the
//               A operand image; the first producer thread also launches the bulk copy of the matching weight chunk
//   warp  8     MMA issuer (converged warp, one elected lane): 3 tcgen05.mma per K=16 step (split-fp16); tcgen05.commit
//               frees the stage
// Pipelines: full[s]/empty[s] per smem stage, acc_full[b]/acc_empty[b] per TMEM accumulator buffer (two buffers when
// 2*NPAD fits in 512 columns, so the epilogue of tile i overlaps the MMAs of tile i+1).
//
// Algorithmic work per row: 2*K*N FLOP (x3 executed for the split), reads K*4 B, writes N*4 B.
#pragma once
#include "pack.cuh"
#include "umma.cuh"

namespace stg {

#define LIN_STAMP(tile_idx, idx)                                                           \
  do {                                                                                     \
    if (p.timeline != nullptr && blockIdx.x == 0 && (tile_idx) < 32)                        \
      p.timeline[(tile_idx) * 8 + (idx)] = clock64();                                       \
  } while (0)

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
  static constexpr uint32_t TR_BYTES = 8u * 32u * TS * 4u;       // 36864: one 4.6 KB buffer per epilogue warp
  static constexpr uint32_t VEC_BYTES = 4u * NPAD * 4u;          // cs | cb (or bias) | ln_out gamma | ln_out beta
  static constexpr uint32_t OFF_TR = STAGES * STAGE_BYTES;
  static constexpr uint32_t OFF_VEC = OFF_TR + TR_BYTES;
  static constexpr uint32_t SMEM_BYTES = OFF_VEC + VEC_BYTES + 1024u;
  static constexpr int ACC_BUFS = (2 * NPAD <= 512 && NPAD <= 256) ? 2 : 1;
  static constexpr uint32_t ACC_STRIDE = NPAD <= 128 ? 128u : 256u;
  static constexpr uint32_t TMEM_COLS = (ACC_BUFS == 2) ? 2u * ACC_STRIDE : 512u;
  static constexpr int N0 = NPAD <= 256 ? NPAD : ((NPAD / 2 + 15) / 16 * 16);  // first MMA N; second = NPAD - N0
  // With a single accumulator (N > 256) the epilogue cannot hide behind the next tile's MMAs, so it gets a second set of
  // four warps (13-16) that takes the upper half of the 32-column blocks.
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = (EPI_WARPS == 8) ? 544 : 416;
  static constexpr int ITEMS = 128 * K8 / 256;  // (row, 8-column slot) pairs per producer thread per chunk
  static constexpr int PREF = (KC == 32) ? 2 : 1;  // chunks of global loads in flight per producer thread
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
  long long* timeline;    // optional debug: [32 tiles][8] SM clock stamps of block 0 (tools/linear_timeline.py)
  int out_blocked;        // 1: y is written as [m / rows_per_block][n / 4][rows_per_block] float4 (see cqa.cuh, Za)
  int rows_per_block;
  long long ldy;          // row stride of y in floats (>= n; a column slice of a wider tensor)
  const float* x_rowmax;  // optional [m]: max |x| of every row; the row is multiplied by 2^-ceil(log2(max)) in front of the fp16
                          // split and its result scaled back - gradient rows sit far below fp16's normal range and differ by
                          // orders of magnitude from row to row
};

// exact power-of-two scale that brings a row's largest magnitude into (0.5, 1]
__device__ __forceinline__ float pow2_scale(float m) {
  return (m > 0.f && m < 3.0e38f) ? exp2f(-ceilf(log2f(m))) : 1.f;
}

template <int NPAD, int KC, int STAGES>
__global__ void __launch_bounds__((LinearCfg<NPAD, KC, STAGES>::THREADS), 1) linear_kernel(LinearDev p) {
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
      mbar_init(&full_bar[s], 257);  // 256 producers behind their own stores + the weight copy's expect_tx arrival
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], Cfg::EPI_WARPS * 32);
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
  if (warp == 8) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  const long long n_tiles = (p.m + 127) / 128;
  const bool need_stats = (p.cs != nullptr) || p.l2norm_in;

  if (warp >= 9) {
    // ------------------------------------------------------------------ epilogue: thread per row; with eight epilogue warps
    // the second four (TMEM lane quadrant = warp % 4 like everyone) emit the upper half of the column blocks.  The
    // epilogue warps carry the HIGHEST warp ids: the schedulers favour high ids, and with the low ids these warps took
    // 18 k cycles to get through an EMPTY epilogue next to the producers (measured), which gates the next tile's MMAs.
    const uint32_t ew = warp - 9;  // 0..7
    const uint32_t qw = warp & 3u;                   // lane quadrant = the 32 rows this warp handles
    uint32_t tile_it = 0;
    const bool ln = (p.ln_out_g != nullptr);
    float* tr = reinterpret_cast<float*>(smem + Cfg::OFF_TR) + ew * 32 * Cfg::TS;
    const float inv_k = 1.f / static_cast<float>(p.k), inv_n = 1.f / static_cast<float>(p.n);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t b = tile_it % Cfg::ACC_BUFS, use = tile_it / Cfg::ACC_BUFS;
      mbar_wait(&acc_full[b], use & 1u, 11);
      tc_fence_after();
      if (ew == 0 && lane == 0) LIN_STAMP(tile_it, 4);
      const int r = qw * 32 + lane;
      const long long row0 = tile * 128 + qw * 32;  // first global row of this warp
      const uint32_t trow = tbase + b * Cfg::ACC_STRIDE + ((qw * 32u) << 16);
      // input-side scalars of my row:  pre[c] = ga * acc[c] + gm * cs[c] + cb[c]   (acc is the GEMM of the SHIFTED row x - c)
      float ga = 1.f, gm = 0.f;
      if (p.x_rowmax != nullptr && row0 + lane < p.m) ga = 1.f / pow2_scale(__ldg(p.x_rowmax + row0 + lane));
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
      // activated value of column c:  t = act(ga * acc + gm * cs[c] + cb[c]); 32 columns per TMEM round trip, the next
      // block's tcgen05.ld is always in flight while the current one is processed
      // (the block loops below are deliberately NOT unrolled: fully unrolled they are ~250 KB of code that is executed once
      //  per tile and streams through the instruction cache)
      constexpr int NBLK = (NPAD + 31) / 32;
      // blocks this warp emits (the LayerNorm statistics pass always covers the whole row)
      constexpr int HALF_BLK = (NBLK + 1) / 2;
      const int blk_lo = (Cfg::EPI_WARPS == 8 && ew >= 4) ? HALF_BLK : 0;
      const int blk_hi = (Cfg::EPI_WARPS == 8 && ew < 4) ? HALF_BLK : NBLK;
      auto load_blk = [&](int blk, float (&v)[32]) {
        tmem_ld16(trow + blk * 32, v);
        if (blk * 32 + 16 < NPAD) tmem_ld16(trow + blk * 32 + 16, v + 16);
      };
      auto activate = [&](int blk, float (&v)[32]) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int c = blk * 32 + j4 * 4;
          if (c < NPAD) {
            const float4 cs = *reinterpret_cast<const float4*>(vec + c);
            const float4 cb = *reinterpret_cast<const float4*>(vec + NPAD + c);
            v[j4 * 4 + 0] = fmaf(ga, v[j4 * 4 + 0], fmaf(gm, cs.x, cb.x));
            v[j4 * 4 + 1] = fmaf(ga, v[j4 * 4 + 1], fmaf(gm, cs.y, cb.y));
            v[j4 * 4 + 2] = fmaf(ga, v[j4 * 4 + 2], fmaf(gm, cs.z, cb.z));
            v[j4 * 4 + 3] = fmaf(ga, v[j4 * 4 + 3], fmaf(gm, cs.w, cb.w));
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[j4 * 4 + e] = fmaxf(v[j4 * 4 + e], 0.f);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[j4 * 4 + e] = 0.f;
          }
        }
      };
      float mean = 0.f, rstd = 1.f;
      if (ln) {
        // LN_out statistics in one pass, shifted by the row's first activated value (rows reaching here are O(1) ReLU /
        // linear outputs; the shift keeps E[(t-t0)^2] - mean^2 exact for rows with a common offset)
        float s = 0.f, q = 0.f, t0 = 0.f;
        float va[32], vb[32];
        auto accumulate = [&](int blk, float (&v)[32]) {
          activate(blk, v);
          if (blk == 0) t0 = v[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (blk * 32 + j < p.n) {
              const float t = v[j] - t0;
              s += t;
              q = fmaf(t, t, q);
            }
          }
        };
        if constexpr (Cfg::EPI_WARPS == 8) {  // two warps per scheduler hide the TMEM latency: one buffer, fewer registers
#pragma unroll 1
          for (int blk = 0; blk < NBLK; ++blk) {
            load_blk(blk, va);
            tmem_wait_ld();
            accumulate(blk, va);
          }
        } else {
          load_blk(0, va);
#pragma unroll 1
          for (int blk = 0; blk < NBLK; blk += 2) {
            tmem_wait_ld();
            if (blk + 1 < NBLK) load_blk(blk + 1, vb);
            accumulate(blk, va);
            if (blk + 1 < NBLK) {
              tmem_wait_ld();
              if (blk + 2 < NBLK) load_blk(blk + 2, va);
              accumulate(blk + 1, vb);
            }
          }
        }
        const float d = s * inv_n;
        mean = t0 + d;
        rstd = 1.f / sqrtf(fmaxf(q * inv_n - d * d, 0.f) + 1e-5f);
      }
      {
        // after the 32x32 transpose lane = (row % 4 group, 4-column slot); its 8 rows' output offsets are fixed per tile
        const int c4 = lane & 7, rsub = lane >> 3;
        const bool vec_ok = p.out_blocked || ((p.n & 3) == 0 && (p.ldy & 3) == 0);
        long long off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long gr = row0 + i * 4 + rsub;
          if (gr >= p.m) {
            off[i] = -1;
          } else if (p.out_blocked) {
            const long long bk = gr / p.rows_per_block;
            off[i] = (bk * (p.n >> 2) * p.rows_per_block + (gr - bk * p.rows_per_block)) * 4;
          } else {
            off[i] = gr * p.ldy;
          }
        }
        const int cmul = p.out_blocked ? p.rows_per_block * 4 : 4;
        const float nm = -mean * rstd;
        if (ew == 0 && lane == 0) LIN_STAMP(tile_it, 6);
        float va[32], vb[32];
        auto emit = [&](int blk, float (&v)[32]) {
          activate(blk, v);
          if (ln) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const int c = blk * 32 + j4 * 4;
              if (c < NPAD) {
                const float4 g = *reinterpret_cast<const float4*>(vec + 2 * NPAD + c);
                const float4 bb = *reinterpret_cast<const float4*>(vec + 3 * NPAD + c);
                v[j4 * 4 + 0] = fmaf(fmaf(v[j4 * 4 + 0], rstd, nm), g.x, bb.x);
                v[j4 * 4 + 1] = fmaf(fmaf(v[j4 * 4 + 1], rstd, nm), g.y, bb.y);
                v[j4 * 4 + 2] = fmaf(fmaf(v[j4 * 4 + 2], rstd, nm), g.z, bb.z);
                v[j4 * 4 + 3] = fmaf(fmaf(v[j4 * 4 + 3], rstd, nm), g.w, bb.w);
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(tr + lane * Cfg::TS + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int col = blk * 32 + c4 * 4;
          if (col < p.n) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 o = *reinterpret_cast<const float4*>(tr + (i * 4 + rsub) * Cfg::TS + c4 * 4);
              if (off[i] >= 0) {
                if (vec_ok) {
                  *reinterpret_cast<float4*>(p.y + off[i] + static_cast<long long>(col >> 2) * cmul) = o;
                } else {
                  const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll 1
                  for (int e = 0; e < 4; ++e)
                    if (col + e < p.n) p.y[off[i] + col + e] = ov[e];
                }
              }
            }
          }
        };
        if constexpr (Cfg::EPI_WARPS == 8) {
#pragma unroll 1
          for (int blk = blk_lo; blk < blk_hi; ++blk) {
            load_blk(blk, va);
            tmem_wait_ld();
            emit(blk, va);
          }
        } else {
          load_blk(blk_lo, va);
#pragma unroll 1
          for (int blk = blk_lo; blk < blk_hi; blk += 2) {
            tmem_wait_ld();
            if (blk + 1 < blk_hi) load_blk(blk + 1, vb);
            emit(blk, va);
            if (blk + 1 < blk_hi) {
              tmem_wait_ld();
              if (blk + 2 < blk_hi) load_blk(blk + 2, va);
              emit(blk + 1, vb);
            }
          }
        }
      }
      if (ew == 0 && lane == 0) LIN_STAMP(tile_it, 7);
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);
      if (ew == 0 && lane == 0) LIN_STAMP(tile_it, 5);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ producers (warps 0-7)
    const uint32_t ptid = tid;
    constexpr int K8 = Cfg::K8, NI = Cfg::ITEMS, PREF = Cfg::PREF;
    const int k8 = ptid % K8;
    const int rbase = ptid / K8;                 // my rows of a tile: rbase + i * (256 / K8)
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long total = my_tiles * p.n_chunks;  // flat (tile, chunk) sequence of this CTA

    float4 buf[PREF][NI][2];
    float xsr[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) xsr[i] = 1.f;
    const bool wide = (p.k & 7) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 31u) == 0;  // rows and slots are 32-byte aligned
    long long ptl = 0;  // (tile, chunk) of the next prefetch
    int pc = 0;
    // my rows of the prefetch tile: slot-0 address of each (nullptr = row beyond m / tile beyond my last), set when the
    // prefetch moves on to a new tile -- the per-chunk address is then one add
    const float* rowp[NI];
    auto set_rows = [&]() {
      const long long row0 = (blockIdx.x + ptl * gridDim.x) * 128;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const long long gr = row0 + rbase + i * (256 / K8);
        rowp[i] = (ptl < my_tiles && gr < p.m) ? p.x + gr * p.k + k8 * 8 : nullptr;
      }
    };
    set_rows();
    auto issue = [&](float4 (&dst)[NI][2]) {
      const int k0 = pc * KC + k8 * 8;
      const bool in4 = k0 + 4 <= p.k, in8 = k0 + 8 <= p.k;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float4* src = reinterpret_cast<const float4*>(rowp[i] + pc * KC);
        dst[i][0] = dst[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rowp[i] != nullptr) {
          if (wide) {  // one 256-bit request per 8-column slot: every 32-byte sector is fetched once, by one instruction
            if (in8) ldg256(src, dst[i][0], dst[i][1]);
          } else {
            if (in4) dst[i][0] = __ldg(src);
            if (in8) dst[i][1] = __ldg(src + 1);
          }
        }
      }
      if (++pc == p.n_chunks) {
        pc = 0;
        ++ptl;
        set_rows();
      }
    };
#pragma unroll
    for (int d = 0; d < PREF; ++d) issue(buf[d]);

    uint32_t it = 0;  // flat chunk counter (stage / phase tracking)
    long long tl = 0;
    int c = 0;
    float cshift[NI];
    float2 s1[NI], s2[NI];  // packed partial sums (even, odd columns)
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      cshift[i] = 0.f;
      s1[i] = s2[i] = make_float2(0.f, 0.f);
    }
    for (long long g0 = 0; g0 < total; g0 += PREF) {
#pragma unroll
      for (int d = 0; d < PREF; ++d) {
        if (g0 + d < total) {  // flat chunk g0 + d = (tile tl, chunk c) sits in register buffer d
          const uint32_t s = it % STAGES, use = it / STAGES;
          uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_STAGE;
          const int kleft = p.k - (c * KC + k8 * 8);  // valid columns of my 8-column slot in this chunk
          if (ptid == 0 && c == 0) LIN_STAMP(tl, 0);
          const bool full = (c + 1) * KC <= p.k;  // every slot of this chunk lies inside K: nothing to mask
          // KC = 64 (four items per thread): wait for the stage first, then every item goes from registers straight to
          // shared memory; KC = 32: convert first so that the conversion hides the wait (16 registers parked)
          constexpr bool WAIT_FIRST = (NI > 2);
          auto claim_stage = [&]() {
            mbar_wait(&empty_bar[s], (use & 1u) ^ 1u, 10);
            if (ptid == 0) {
              mbar_arrive_expect_tx(&full_bar[s], Cfg::B_STAGE);
              bulk_g2s(sB, p.w + static_cast<size_t>(c) * Cfg::B_STAGE, Cfg::B_STAGE, &full_bar[s]);
            }
          };
          if constexpr (WAIT_FIRST) claim_stage();
          HL8 hl[WAIT_FIRST ? 1 : NI];
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            float2 v0 = lo2(buf[d][i][0]), v1 = hi2(buf[d][i][0]), v2 = lo2(buf[d][i][1]), v3 = hi2(buf[d][i][1]);
            if (p.x_rowmax != nullptr) {
              if (c == 0) {
                const long long gr = (blockIdx.x + tl * gridDim.x) * 128 + rbase + i * (256 / K8);
                xsr[i] = gr < p.m ? pow2_scale(__ldg(p.x_rowmax + gr)) : 1.f;
              }
              const float2 xs2 = dup2(xsr[i]);
              v0 = mul2(v0, xs2); v1 = mul2(v1, xs2); v2 = mul2(v2, xs2); v3 = mul2(v3, xs2);
            }
            if (need_stats) {
              if (c == 0) {
                // with LN_in the row is shifted by its first element (held by the k8 == 0 lane) BEFORE it meets the tensor
                // cores: a constant row then gives exactly beta and a large common offset costs no GEMM precision
                const float first = __shfl_sync(0xffffffffu, v0.x, lane & ~(K8 - 1));
                cshift[i] = (p.cs != nullptr) ? first : 0.f;
                s1[i] = s2[i] = make_float2(0.f, 0.f);
              }
              const float2 nc = dup2(-cshift[i]);
              v0 = add2(v0, nc); v1 = add2(v1, nc); v2 = add2(v2, nc); v3 = add2(v3, nc);
              if (!full) {  // the chunk that crosses K: columns >= K contribute nothing (the shift made them non-zero)
                v0.x = 0 < kleft ? v0.x : 0.f; v0.y = 1 < kleft ? v0.y : 0.f; v1.x = 2 < kleft ? v1.x : 0.f; v1.y = 3 < kleft ? v1.y : 0.f;
                v2.x = 4 < kleft ? v2.x : 0.f; v2.y = 5 < kleft ? v2.y : 0.f; v3.x = 6 < kleft ? v3.x : 0.f; v3.y = 7 < kleft ? v3.y : 0.f;
              }
              s1[i] = add2(s1[i], add2(add2(v0, v1), add2(v2, v3)));
              s2[i] = fma2(v0, v0, fma2(v1, v1, fma2(v2, v2, fma2(v3, v3, s2[i]))));
            }
            HL8& h8 = hl[WAIT_FIRST ? 0 : i];
            split2p(v0, h8.hi.x, h8.lo.x);
            split2p(v1, h8.hi.y, h8.lo.y);
            split2p(v2, h8.hi.z, h8.lo.z);
            split2p(v3, h8.hi.w, h8.lo.w);
            if constexpr (WAIT_FIRST) {
              uint8_t* dst = sA + k8 * Cfg::LBO_A + (rbase + i * (256 / K8)) * 16;
              *reinterpret_cast<uint4*>(dst) = h8.hi;
              *reinterpret_cast<uint4*>(dst + Cfg::A_PART) = h8.lo;
            }
          }
          issue(buf[d]);  // refill this register buffer: the loads fly while the next chunks are converted
          if constexpr (!WAIT_FIRST) {
            claim_stage();
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              uint8_t* dst = sA + k8 * Cfg::LBO_A + (rbase + i * (256 / K8)) * 16;
              *reinterpret_cast<uint4*>(dst) = hl[WAIT_FIRST ? 0 : i].hi;
              *reinterpret_cast<uint4*>(dst + Cfg::A_PART) = hl[WAIT_FIRST ? 0 : i].lo;
            }
          }
          if (need_stats && c == p.n_chunks - 1) {  // the row statistics travel with the last chunk of the tile
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              float a = s1[i].x + s1[i].y, q = s2[i].x + s2[i].y;
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
          mbar_arrive(&full_bar[s]);  // (thread 0 too: its expect_tx arrival above came before its own A rows were stored)
          ++it;
          if (++c == p.n_chunks) {
            if (ptid == 0) LIN_STAMP(tl, 1);
            c = 0;
            ++tl;
          }
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- MMA issuer (warp 8): the whole warp runs the control flow
    // (waits included), one elected lane issues -- in converged code the descriptors stay in uniform registers, under a
    // `lane == 0` branch every tcgen05.mma sits in a lane-by-lane waterfall loop (~130 cycles per MMA measured)
    uint32_t it = 0, tile_it = 0;
    const uint32_t idesc0 = make_idesc_f16(Cfg::N0);
    const uint32_t idesc1 = make_idesc_f16(NPAD - Cfg::N0 > 0 ? NPAD - Cfg::N0 : 16);
    const uint32_t smem0 = smem_u32(smem);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t b = tile_it % Cfg::ACC_BUFS, ause = tile_it / Cfg::ACC_BUFS;
      mbar_wait(&acc_empty[b], (ause & 1u) ^ 1u, 12);
      tc_fence_after();
      if (lane == 0) LIN_STAMP(tile_it, 2);
      const uint32_t dcol = tbase + b * Cfg::ACC_STRIDE;
      for (int c = 0; c < p.n_chunks; ++c, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        mbar_wait(&full_bar[s], use & 1u, 13);
        tc_fence_after();
        const uint32_t sA = smem0 + s * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + Cfg::A_STAGE;
        if (elect_one()) {
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
          if (c == p.n_chunks - 1) umma_commit(&acc_full[b]);
        }
      }
      if (lane == 0) LIN_STAMP(tile_it, 3);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<Cfg::TMEM_COLS>(tbase);
}

}  // namespace stg
