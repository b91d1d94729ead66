// EncoderBlock on one SM-resident tile:  x + pe -> n_conv x [LN -> depthwise conv -> pointwise GEMM -> ReLU -> +res] -> LN
// (model/encoder.py:29-52).  A tile is 128 rows = floor(128/L) whole sequences; the fp32 residual stream X never leaves
// shared memory, the pointwise 128x128 GEMM runs on tcgen05 (split-fp16, accumulator in TMEM).
//
// v3: TWO CTAs per SM.  The latency chain of a tile (produce A -> MMA -> read accumulator, twice) cannot be shortened, so
// the kernel is built to fit twice on an SM and let the hardware interleave two tiles:
//   * the A operand of the pointwise GEMM lives in TENSOR MEMORY (tcgen05.st, two fp16 per 32-bit column, hi and lo parts
//     64 columns each) instead of a 66 KB shared-memory image; with the 128-column accumulator a CTA owns 256 of the 512
//     TMEM columns;
//   * the pointwise weights stream through a 2 x 16 KB ring of K=32 chunks (cp.async.bulk from L2) instead of a resident
//     64 KB image;
//   * shared memory per CTA = X (66 KB) + ring (32 KB) + per-channel parameter tables (<= 13 KB)  ->  ~113 KB.
// Thread = (row, column half) for every phase (the TMEM lane a thread may touch fixes its row), so there is no warp
// shuffle anywhere; the two halves of a row meet through the 16 padding bytes of the row in X.
//   phase 0   raw rows + position encoding -> X, LayerNorm statistics of the row
//   conv      out[r,c] = db + g*sum_j w[c,j]*(x[r+j-P,c]*rstd_j - mean_j*rstd_j) + beta*sum_j w[c,j]   (taps outside the row's
//             own sequence dropped), i.e. the LayerNorm is applied tap by tap from per-row scalars; split to fp16 (hi, lo),
//             16 columns at a time straight into the TMEM A operand
//   GEMM      MMA thread: 4 ring chunks x 2 K-steps x 3 MMAs (split product, A from TMEM, B from the ring)
//   epilogue  X += relu(D + b), statistics of the NEXT LayerNorm from the values being written
//   tail      closing LayerNorm -> store, or mask_logits + max over the sequence (model/stage.py:503)
// Warp roles (320 threads): warps 0-7 compute, warp 8 = MMA issue, warp 9 = weight-ring copies.
//
// Packed tiling (round 2, masked max-pool output only).  With one 100-row sequence per 128-row tile 22 % of every phase
// worked on rows nobody reads.  In packed mode the rows of ALL sequences are treated as one long array that is cut into
// tiles of 128 input rows advancing by OUT = 128 - 2*halo rows (halo = n_conv * (k/2): the receptive field of the conv
// stack): every tile recomputes `halo` rows on either side, the rows [halo, 128 - halo) are exact, and a sequence that
// straddles a cut gets its column maxima from two tiles through an order-independent atomic max.  5 sequences of 100
// rows take 4.17 tiles instead of 5.
//
// Padded tokens are processed like real ones (the reference never masks before the convs, SURVEY.md §7 quirk i); the
// depthwise conv sees zeros outside [0, L) of its own sequence only.
//
// Algorithmic work per token: n_conv*(2*k*128 + 2*128*128) FLOP; reads 512 B, writes 512 B (mode 0) or 512/L B.
#pragma once
#include "umma.cuh"

namespace stg {

struct EncCfg {
  static constexpr int XS = 132;                       // padded fp32 row stride (floats): conflict-free row-per-lane float4 access;
                                                       // the 4 padding floats of a row carry the half-row partial statistics
  static constexpr uint32_t X_BYTES = 128u * XS * 4u;  // 67584
  static constexpr uint32_t LBO_B = 128u * 16u;        // 2048
  static constexpr uint32_t W_HALF = 4u * LBO_B;       // 8192: hi (or lo) part of a K=32 chunk
  static constexpr uint32_t W_SLOT = 2u * W_HALF;      // 16384
  static constexpr int W_SLOTS = 2;
  static constexpr uint32_t W_SRC_CHUNK = 32768;       // stage_pack_weight(kc=64): [chunk][hi 8 k-groups | lo 8 k-groups]
  static constexpr uint32_t OFF_W = X_BYTES;
  static constexpr uint32_t OFF_P = OFF_W + W_SLOTS * W_SLOT;  // 100352
  static constexpr int P_SLOTS = 2;                    // parameter tables of two convs are resident
  static constexpr int p_floats(int ks) { return (ks + 5) * 128; }  // w[ks][128], sum_j w, ln g, ln b, dw bias, pw bias
  static constexpr uint32_t smem_bytes(int ks) { return OFF_P + (P_SLOTS * p_floats(ks) + 2 * 128) * 4u; }
  static constexpr int THREADS = 320;
  static constexpr int COMPUTE = 256;
  static constexpr uint32_t TMEM_COLS = 256;           // D 0..127 | A_hi 128..191 | A_lo 192..255
};

struct EncDev {
  const float* x;
  const float* mask;
  const float* pe;
  const float* ln_g[STAGE_MAX_CONV];
  const float* ln_b[STAGE_MAX_CONV];
  const float* dw_w[STAGE_MAX_CONV];
  const float* dw_b[STAGE_MAX_CONV];
  const uint8_t* pw[STAGE_MAX_CONV];
  const float* pw_b[STAGE_MAX_CONV];
  const float* fin_g;
  const float* fin_b;
  float* y;
  long long n_seq;
  int L, n_conv, out_mode, flags;
  int S;  // row stride of the blocked input layout (STAGE_ENC_IN_BLOCKED)
  int packed;  // 1: tiles cut the concatenated rows of all sequences every 128 - 2*halo rows (out_mode 1 only; see below)
  long long* timeline;  // optional debug: [32 tiles][16] SM clock stamps of thread 0 of CTA 0 (tools/encoder_timeline.py)
  // out_mode 2: geometry of the CQA context blob the sequence is written into (filled from CqaCfg by the host code)
  uint32_t blob_bytes, blob_lbo, blob_part, blob_inv_off, blob_mask_off;
};

#define STG_NEG (-1e10f)
#define ENC_STAMP(idx)                                                                   \
  do {                                                                                   \
    if (p.timeline != nullptr && blockIdx.x == 0 && tid == 0 && tl_tile < 32)             \
      p.timeline[tl_tile * 16 + (idx)] = clock64();                                       \
  } while (0)
#define ENC_BAR() asm volatile("bar.sync 1, 256;" ::: "memory")

template <int KS, bool PACKED>
__global__ void __launch_bounds__(320, 2) encoder_kernel(EncDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t a_full, mma_done, w_full[EncCfg::W_SLOTS], w_empty[EncCfg::W_SLOTS];
  __shared__ uint32_t tmem_slot;
  __shared__ float2 row_stat[128];  // (rstd, -mean*rstd) of the LayerNorm that reads X next; max-pool partials in the tail

  float* X = reinterpret_cast<float*>(smem);
  uint8_t* sW = smem + EncCfg::OFF_W;
  float* PT = reinterpret_cast<float*>(smem + EncCfg::OFF_P);
  constexpr int PF = (KS + 5) * 128;  // == EncCfg::p_floats(KS)
  float* FIN = PT + EncCfg::P_SLOTS * PF;  // closing LayerNorm gamma | beta
  constexpr int XS = EncCfg::XS;
  constexpr int P = KS / 2;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int L = p.L;
  const int G = 128 / L;  // sequences per tile
  const int halo = p.n_conv * (KS / 2), OUT = 128 - 2 * halo;  // packed mode: rows of a tile whose whole receptive field is inside it
  const int total_rows = PACKED ? static_cast<int>(p.n_seq * L) : 0;  // (host-checked: fits 31 bits)
  const long long n_tiles = PACKED ? (total_rows + OUT - 1) / OUT : (p.n_seq + G - 1) / G;
  const long long n_my = static_cast<long long>(blockIdx.x) < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (tid == 0) {
    mbar_init(&a_full, EncCfg::COMPUTE);
    mbar_init(&mma_done, 1);
    for (int i = 0; i < EncCfg::W_SLOTS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<EncCfg::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  if (warp < 8) {
    // =================================================================== compute warps: thread = (row, column half)
    const int r = (warp & 3) * 32 + lane, h = warp >> 2;
    const uint32_t lane_off = ((warp & 3) * 32u) << 16;
    int sq = r / L, pos = r - sq * L;  // my row's sequence within the tile and position within the sequence (per tile when packed)
    // tap j of the depthwise conv lies inside my row's own sequence.  Rows behind the last sequence of a tile belong to no
    // sequence; nothing reads what they produce, so they run as interior rows (every tap on, neighbours clamped to the tile)
    uint32_t vbits = 0;
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      const int pp = pos + j - P;
      if (sq >= G || (pp >= 0 && pp < L)) vbits |= 1u << j;
    }
    bool interior = vbits == (1u << KS) - 1u;
    float* xrow = X + r * XS;
    float* xh = xrow + h * 64;
    const bool blocked = (p.flags & STAGE_ENC_IN_BLOCKED) != 0;
    const bool use_pe = !(p.flags & STAGE_ENC_NO_PE);
    const bool fin = !(p.flags & STAGE_ENC_NO_FINAL_LN);
    const bool resident = p.n_conv <= EncCfg::P_SLOTS;
    uint32_t done_cnt = 0;
    int tl_tile = 0;

    // per-channel tables of conv ci -> parameter slot s (threads 0..127, one channel each)
    auto load_params = [&](int ci, int s) {
      if (tid < 128) {
        float* T = PT + s * PF;
        const int c = tid;
        float sw = 0.f;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const float w = __ldg(p.dw_w[ci] + c * KS + j);
          T[j * 128 + c] = w;
          sw += w;
        }
        T[KS * 128 + c] = sw;
        T[(KS + 1) * 128 + c] = p.ln_g[ci] ? __ldg(p.ln_g[ci] + c) : 1.f;
        T[(KS + 2) * 128 + c] = p.ln_b[ci] ? __ldg(p.ln_b[ci] + c) : 0.f;
        T[(KS + 3) * 128 + c] = __ldg(p.dw_b[ci] + c);
        T[(KS + 4) * 128 + c] = __ldg(p.pw_b[ci] + c);
      }
    };
    if (resident)
      for (int ci = 0; ci < p.n_conv; ++ci) load_params(ci, ci);
    if (tid < 128) {
      FIN[tid] = fin ? __ldg(p.fin_g + tid) : 1.f;
      FIN[128 + tid] = fin ? __ldg(p.fin_b + tid) : 0.f;
    }
    ENC_BAR();

    // statistics of my row from the two half-row partials parked in the row's padding: (rstd, -mean*rstd)
    auto combine = [&]() -> float2 {
      const float4 q = *reinterpret_cast<const float4*>(xrow + 128);
      const float mean = (q.x + q.z) * (1.f / 128.f);
      const float ex2 = (q.y + q.w) * (1.f / 128.f);
      const float rs = 1.f / sqrtf(fmaxf(ex2 - mean * mean, 0.f) + 1e-5f);
      return make_float2(rs, -mean * rs);
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      long long seq0 = tile * G;  // first sequence of the tile; packed: the sequence of my row
      const int n_here = static_cast<int>((p.n_seq - seq0) < G ? (p.n_seq - seq0) : G);
      int n_valid = n_here * L;
      bool live = r < n_valid;
      const int g0 = PACKED ? static_cast<int>(tile) * OUT - halo : 0;  // packed: global row of tile row 0
      if constexpr (PACKED) {
        const int g = g0 + r;
        live = g >= 0 && g < total_rows;
        const int sg = live ? g / L : 0;
        seq0 = sg;
        sq = 0;  // (address arithmetic below uses seq0 + sq)
        pos = live ? g - sg * L : 0;
        vbits = 0;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const int pp = pos + j - P;
          if (!live || (pp >= 0 && pp < L)) vbits |= 1u << j;
        }
        interior = vbits == (1u << KS) - 1u;
        const int left = total_rows - g0;
        n_valid = left < 128 ? left : 128;  // rows of the tile that exist (row-major loads)
      }
      ENC_STAMP(0);
      if (PACKED && tid < 3 && tile + gridDim.x < n_tiles) {
        // packed: the next tile of this CTA touches at most three sequences; their blocks / rows start their way into L2 now
        const int gn = static_cast<int>(tile + gridDim.x) * OUT - halo;
        const int sn = (gn < 0 ? 0 : gn) / L + static_cast<int>(tid);
        if (static_cast<long long>(sn) < p.n_seq && sn * L < gn + 128) {
          const float* nx = p.x + static_cast<long long>(sn) * (blocked ? 128ll * p.S : 128ll * L);
          const uint32_t nbytes = blocked ? 512u * p.S : 512u * L;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"(nbytes) : "memory");
        }
      }
      if (!PACKED && tid == 0 && tile + gridDim.x < n_tiles) {  // the next tile of this CTA starts its way into L2 now (contiguous in both layouts)
        const long long nseq0 = (tile + gridDim.x) * G;
        const long long nn = (p.n_seq - nseq0) < G ? (p.n_seq - nseq0) : G;
        const float* nx = p.x + nseq0 * (blocked ? 128ll * p.S : 128ll * L);
        const uint32_t nbytes = static_cast<uint32_t>(nn) * (blocked ? 512u * p.S : 512u * L);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"(nbytes) : "memory");
      }
      const float m_row = (live && p.mask != nullptr) ? __ldg(p.mask + (PACKED ? static_cast<long long>(g0 + r) : seq0 * L + r)) : 0.f;

      // ---- phase 0: raw rows (+ position encoding) -> X, first LayerNorm statistics
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);  // packed partial sums: (even, odd) float2 lanes
      if (blocked) {
        // float4 (row, c4) of sequence i at (i*32 + c4)*S + row, the position table at c4*L + pos: lanes = rows read runs
        const float4* src = reinterpret_cast<const float4*>(p.x) + ((seq0 + sq) * 32 + h * 16) * p.S + pos;
        const float4* pe4 = reinterpret_cast<const float4*>(p.pe) + (h * 16) * L + pos;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float4 v[8], e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = live ? __ldg(src + static_cast<size_t>(b * 8 + j) * p.S) : make_float4(0.f, 0.f, 0.f, 0.f);
            e[j] = (live && use_pe) ? __ldg(pe4 + (b * 8 + j) * L) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 a = add2(lo2(v[j]), lo2(e[j])), c = add2(hi2(v[j]), hi2(e[j]));
            *reinterpret_cast<float4*>(xh + (b * 8 + j) * 4) = make_float4(a.x, a.y, c.x, c.y);
            s1 = add2(s1, add2(a, c));
            s2 = fma2(a, a, fma2(c, c, s2));
          }
        }
      } else {
        // row-major input: warp per row so that a row is one 512-byte request, 4 rows in flight per warp
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float4 v[4], e[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = static_cast<int>(warp) + 8 * (b * 4 + i);
            const long long grow = PACKED ? static_cast<long long>(g0 + rr) : seq0 * L + rr;  // global row (all sequences concatenated)
            const bool on = PACKED ? (grow >= 0 && grow < total_rows) : rr < n_valid;
            v[i] = on ? __ldg(reinterpret_cast<const float4*>(p.x + grow * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            e[i] = (on && use_pe) ? __ldg(reinterpret_cast<const float4*>(p.pe + (PACKED ? static_cast<int>(grow) % L : rr - (rr / L) * L) * 128) + lane)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = static_cast<int>(warp) + 8 * (b * 4 + i);
            *reinterpret_cast<float4*>(X + rr * XS + lane * 4) =
                make_float4(v[i].x + e[i].x, v[i].y + e[i].y, v[i].z + e[i].z, v[i].w + e[i].w);
          }
        }
        ENC_BAR();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 o = *reinterpret_cast<const float4*>(xh + j * 4);
          s1 = add2(s1, add2(lo2(o), hi2(o)));
          s2 = fma2(lo2(o), lo2(o), fma2(hi2(o), hi2(o), s2));
        }
      }
      *reinterpret_cast<float2*>(xrow + 128 + 2 * h) = make_float2(s1.x + s1.y, s2.x + s2.y);
      ENC_BAR();
      float2 my_stat = combine();
      if (h == 0) row_stat[r] = my_stat;
      ENC_BAR();
      ENC_STAMP(1);

      for (int ci = 0; ci < p.n_conv; ++ci) {
        const int ps = resident ? ci : (ci & 1);
        if (!resident) {  // more convs than parameter slots: (re)load this conv's tables
          load_params(ci, ps);
          ENC_BAR();
        }
        const float* T = PT + ps * PF;
        // ---- conv: LayerNorm applied tap by tap from the per-row scalars, k-tap FIR, split -> TMEM A operand
        {
          const bool idn = (p.ln_g[ci] == nullptr);  // no LayerNorm in front of this conv
          float2 rs[KS], nm[KS];  // (rstd, rstd), (-mean*rstd, -mean*rstd) of the tap rows; 0 for a tap outside the sequence
          int xo[KS];
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const int q = min(max(r + j - P, 0), 127);
            const bool on = (vbits >> j) & 1u;
            const float2 st = idn ? make_float2(1.f, 0.f) : row_stat[q];
            rs[j] = dup2(on ? st.x : 0.f);
            nm[j] = dup2(on ? st.y : 0.f);
            xo[j] = q * XS + h * 64;
          }
          const float* Th = T + h * 64;
#pragma unroll 1
          for (int cg = 0; cg < 4; ++cg) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int col = cg * 16 + c4 * 4;
              float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
              for (int j = 0; j < KS; ++j) {
                const float4 xv = *reinterpret_cast<const float4*>(X + xo[j] + col);
                const float4 w = *reinterpret_cast<const float4*>(Th + j * 128 + col);
                a0 = fma2(lo2(w), fma2(lo2(xv), rs[j], nm[j]), a0);
                a1 = fma2(hi2(w), fma2(hi2(xv), rs[j], nm[j]), a1);
              }
              const float4 sw4 = *reinterpret_cast<const float4*>(Th + KS * 128 + col);
              float2 sw0 = lo2(sw4), sw1 = hi2(sw4);
              if (!interior) {  // a row near a sequence end: beta reaches the output only through the taps that exist
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                  if (!((vbits >> j) & 1u)) {  // at most P taps per row, and only in the warps that hold a sequence end
                    const float4 w = *reinterpret_cast<const float4*>(Th + j * 128 + col);
                    sw0 = fma2(lo2(w), make_float2(-1.f, -1.f), sw0);
                    sw1 = fma2(hi2(w), make_float2(-1.f, -1.f), sw1);
                  }
                }
              }
              const float4 g4 = *reinterpret_cast<const float4*>(Th + (KS + 1) * 128 + col);
              const float4 b4 = *reinterpret_cast<const float4*>(Th + (KS + 2) * 128 + col);
              const float4 d4 = *reinterpret_cast<const float4*>(Th + (KS + 3) * 128 + col);
              split2p(fma2(lo2(g4), a0, fma2(lo2(b4), sw0, lo2(d4))), hi[c4 * 2], lo[c4 * 2]);
              split2p(fma2(hi2(g4), a1, fma2(hi2(b4), sw1, hi2(d4))), hi[c4 * 2 + 1], lo[c4 * 2 + 1]);
            }
            const uint32_t ta = tbase + lane_off + 128 + h * 32 + cg * 8;
            tmem_st8(ta, make_uint4(hi[0], hi[1], hi[2], hi[3]), make_uint4(hi[4], hi[5], hi[6], hi[7]));
            tmem_st8(ta + 64, make_uint4(lo[0], lo[1], lo[2], lo[3]), make_uint4(lo[4], lo[5], lo[6], lo[7]));
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&a_full);
        if (ci < 2) ENC_STAMP(2 + ci * 4);

        // ---- pointwise GEMM on the MMA warp
        mbar_wait(&mma_done, done_cnt & 1u, 21);
        ++done_cnt;
        tc_fence_after();
        if (ci < 2) ENC_STAMP(3 + ci * 4);

        // ---- epilogue  X += relu(D + b)  + statistics of the new row
        {
          const float* pwb = T + (KS + 4) * 128 + h * 64;
          const bool relu = !(p.flags & STAGE_ENC_NO_RELU), resid = !(p.flags & STAGE_ENC_NO_RESIDUAL);
          s1 = make_float2(0.f, 0.f);
          s2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld16(tbase + lane_off + h * 64 + half * 32, v);
            tmem_ld16(tbase + lane_off + h * 64 + half * 32 + 16, v + 16);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(pwb + half * 32 + j);
              const float4 xv = *reinterpret_cast<float4*>(xh + half * 32 + j);
              float2 d0 = add2(make_float2(v[j], v[j + 1]), lo2(bb)), d1 = add2(make_float2(v[j + 2], v[j + 3]), hi2(bb));
              if (relu) { d0 = max2(d0, 0.f); d1 = max2(d1, 0.f); }
              if (resid) { d0 = add2(d0, lo2(xv)); d1 = add2(d1, hi2(xv)); }
              *reinterpret_cast<float4*>(xh + half * 32 + j) = make_float4(d0.x, d0.y, d1.x, d1.y);
              s1 = add2(s1, add2(d0, d1));
              s2 = fma2(d0, d0, fma2(d1, d1, s2));
            }
          }
          *reinterpret_cast<float2*>(xrow + 128 + 2 * h) = make_float2(s1.x + s1.y, s2.x + s2.y);
        }
        tc_fence_before();
        ENC_BAR();
        my_stat = combine();
        if (h == 0) row_stat[r] = my_stat;
        ENC_BAR();
        if (ci < 2) ENC_STAMP(4 + ci * 4);
      }

      // ---- closing LayerNorm (in place; mask_logits for the max-pool), then coalesced store / column-wise max
      {
        const float2 rsn = dup2(fin ? my_stat.x : 1.f), nmn = dup2(fin ? my_stat.y : 0.f);
        const float2 mm = dup2(m_row), off = dup2((1.f - m_row) * STG_NEG);  // mask_logits: x*m + (1-m)*(-1e10)
        const float* fg = FIN + h * 64;
        const float* fb = FIN + 128 + h * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(xh + j);
          const float4 g4 = *reinterpret_cast<const float4*>(fg + j), b4 = *reinterpret_cast<const float4*>(fb + j);
          float2 o0 = fma2(fma2(lo2(v), rsn, nmn), lo2(g4), lo2(b4)), o1 = fma2(fma2(hi2(v), rsn, nmn), hi2(g4), hi2(b4));
          if (p.out_mode == 1) { o0 = fma2(o0, mm, off); o1 = fma2(o1, mm, off); }
          *reinterpret_cast<float4*>(xh + j) = make_float4(o0.x, o0.y, o1.x, o1.y);
        }
      }
      ENC_BAR();
      ENC_STAMP(12);
      if (p.out_mode == 2) {
        // one sequence per tile (host-checked): row r = token r of frame `tile`.  The closing LayerNorm output (sitting in
        // X now) becomes the split-fp16 image of the CQA context blob: 16 bytes per (row, 8-column group), lanes = rows, so
        // a warp writes 512-byte runs; rows [L, round_up(L, 16)) are the zero padding of the image.
        const int ls_pad = (L + 15) & ~15;
        uint8_t* blob = reinterpret_cast<uint8_t*>(p.y) + static_cast<size_t>(tile) * p.blob_bytes;
        float2 ss = make_float2(0.f, 0.f);
        if (r < ls_pad) {
#pragma unroll
          for (int kg = 0; kg < 8; ++kg) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (live) {
              a = *reinterpret_cast<const float4*>(xh + kg * 8);
              b = *reinterpret_cast<const float4*>(xh + kg * 8 + 4);
            }
            ss = fma2(lo2(a), lo2(a), fma2(hi2(a), hi2(a), fma2(lo2(b), lo2(b), fma2(hi2(b), hi2(b), ss))));
            uint4 hi, lo;
            split2p(lo2(a), hi.x, lo.x);
            split2p(hi2(a), hi.y, lo.y);
            split2p(lo2(b), hi.z, lo.z);
            split2p(hi2(b), hi.w, lo.w);
            uint8_t* dst = blob + (h * 8 + kg) * p.blob_lbo + r * 16;
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + p.blob_part) = lo;
          }
        }
        // the two half-row sums of squares meet in the row's padding (free again: the statistics were consumed above)
        reinterpret_cast<float*>(xrow + 128)[h] = ss.x + ss.y;
        ENC_BAR();
        if (h == 0 && r < ls_pad) {
          const float tot = xrow[128] + xrow[129];
          // F.normalize: x / max(||x||, eps); a masked row stores 0 so that the CQA kernel needs one load, not two
          reinterpret_cast<float*>(blob + p.blob_inv_off)[r] = (live && m_row != 0.f) ? 1.f / fmaxf(sqrtf(tot), 1e-12f) : 0.f;
          reinterpret_cast<float*>(blob + p.blob_mask_off)[r] = live ? m_row : 0.f;
        }
      } else if (p.out_mode == 0) {
        for (int rr = warp; rr < n_valid; rr += 8)
          *reinterpret_cast<float4*>(p.y + (seq0 * L + rr) * 128 + lane * 4) = *reinterpret_cast<const float4*>(X + rr * XS + lane * 4);
      } else {
        const int c = tid & 127, part = tid >> 7;
        float* pm = reinterpret_cast<float*>(row_stat);  // [2][128]; row_stat is dead until the next tile's phase 0
        if constexpr (PACKED) {
          // exact rows of this tile: [halo, 128 - halo) (and inside the data); every sequence segment in there sends its
          // column maxima to the sequence's output row (max is exact and order independent; y starts at -inf)
          const int ga = g0 + halo < 0 ? 0 : g0 + halo;
          const int gb = g0 + 128 - halo < total_rows ? g0 + 128 - halo : total_rows;
          for (int sg = ga / L; sg * L < gb; ++sg) {
            const int a0 = (sg * L > ga ? sg * L : ga) - g0, b0 = ((sg + 1) * L < gb ? (sg + 1) * L : gb) - g0;
            float mx = -3.0e38f;
#pragma unroll 4
            for (int rr = a0 + part; rr < b0; rr += 2) mx = fmaxf(mx, X[rr * XS + c]);
            if (a0 + part < b0) {
              float* dst = p.y + static_cast<long long>(sg) * 128 + c;
              if (mx >= 0.f) atomicMax(reinterpret_cast<int*>(dst), __float_as_int(mx));
              else atomicMin(reinterpret_cast<unsigned int*>(dst), __float_as_uint(mx));
            }
          }
        } else if (G == 1) {  // one long sequence: the two thread halves take half the rows each
          const int half_rows = (L + 1) >> 1;
          const int lo = part * half_rows, hi = min(L, lo + half_rows);
          float mx = -3.0e38f;
#pragma unroll 4
          for (int rr = lo; rr < hi; ++rr) mx = fmaxf(mx, X[rr * XS + c]);
          pm[part * 128 + c] = mx;
          ENC_BAR();
          if (part == 0 && n_here > 0) p.y[seq0 * 128 + c] = fmaxf(pm[c], pm[128 + c]);
        } else {       // several short sequences: they alternate between the thread halves
          for (int g = part; g < n_here; g += 2) {
            float mx = -3.0e38f;
            for (int rr = g * L; rr < g * L + L; ++rr) mx = fmaxf(mx, X[rr * XS + c]);
            p.y[(seq0 + g) * 128 + c] = mx;
          }
        }
      }
      ENC_STAMP(13);
      ENC_BAR();  // X / row_stat are rewritten by the next tile's phase 0
      ENC_STAMP(14);
      ++tl_tile;
    }
  } else if (warp == 8) {
    // =================================================================== MMA issue: the whole warp runs the control flow, one
    // elected lane issues (converged code keeps the descriptors in uniform registers; under `lane == 0` every tcgen05.mma
    // would sit in a lane-by-lane waterfall loop)
    const uint32_t idesc = make_idesc_f16(128);
    const uint32_t w0 = smem_u32(sW);
    uint32_t q = 0, ph = 0;
    for (long long t = 0; t < n_my; ++t) {
      for (int ci = 0; ci < p.n_conv; ++ci) {
        mbar_wait(&a_full, ph & 1u, 22);  // all 256 producers have stored (and fenced) their part of A; D has been read
        ++ph;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c, ++q) {
          const uint32_t slot = q & 1u;
          mbar_wait(&w_full[slot], (q >> 1) & 1u, 23);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
              const uint32_t ks = c * 2 + k2;
              const uint32_t a_hi = tbase + 128 + ks * 8, a_lo = tbase + 192 + ks * 8;
              const uint32_t wa = w0 + slot * EncCfg::W_SLOT + k2 * 2 * EncCfg::LBO_B;
              const uint64_t b_hi = make_smem_desc(wa, EncCfg::LBO_B, 128);
              const uint64_t b_lo = make_smem_desc(wa + EncCfg::W_HALF, EncCfg::LBO_B, 128);
              umma_f16_ts(tbase, a_lo, b_hi, idesc, ks > 0 ? 1u : 0u);
              umma_f16_ts(tbase, a_hi, b_lo, idesc, 1u);
              umma_f16_ts(tbase, a_hi, b_hi, idesc, 1u);
            }
            umma_commit(&w_empty[slot]);
            if (c == 3) umma_commit(&mma_done);
          }
        }
      }
    }
  } else if (warp == 9 && lane == 0) {
    // =================================================================== weight ring (one thread): chunk q = (tile, conv, K=32 chunk)
    const long long total = n_my * p.n_conv * 4;
    for (long long q = 0; q < total; ++q) {
      const uint32_t slot = static_cast<uint32_t>(q & 1), use = static_cast<uint32_t>(q >> 1);
      if (use > 0) mbar_wait(&w_empty[slot], (use - 1u) & 1u, 24);  // the MMAs that read the slot's previous chunk have completed
      const int ci = static_cast<int>((q >> 2) % p.n_conv), c = static_cast<int>(q & 3);
      const uint8_t* src = p.pw[ci] + (c >> 1) * EncCfg::W_SRC_CHUNK + (c & 1) * EncCfg::W_HALF;
      uint8_t* dst = sW + slot * EncCfg::W_SLOT;
      mbar_arrive_expect_tx(&w_full[slot], EncCfg::W_SLOT);
      bulk_g2s(dst, src, EncCfg::W_HALF, &w_full[slot]);                                         // hi part: k-groups 4*(c&1)..+4
      bulk_g2s(dst + EncCfg::W_HALF, src + EncCfg::W_SRC_CHUNK / 2, EncCfg::W_HALF, &w_full[slot]);  // lo part
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<EncCfg::TMEM_COLS>(tbase);
}

}  // namespace stg
