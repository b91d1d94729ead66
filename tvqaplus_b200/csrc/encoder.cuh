// EncoderBlock on one SM-resident tile:  x + pe -> n_conv x [LN -> depthwise conv -> pointwise GEMM -> ReLU -> +res] -> LN
// (model/encoder.py:29-52).  A tile is 128 rows = floor(128/L) whole sequences; the fp32 residual stream never leaves
// shared memory, the pointwise 128x128 GEMM runs on tcgen05 (split-fp16, accumulator in TMEM).
//
// 512 threads, one CTA per SM, persistent over tiles:
//   phase 0   the tile's raw rows were prefetched by ONE cp.async.bulk (they are contiguous in HBM) into the region
//             that later holds the A image; warp-per-row: + position encoding -> X (fp32, smem) and the row's LayerNorm
//             statistics from the same registers
//   conv      warp = 8 consecutive rows, lane = 4 channels: a register sliding window of LayerNorm-ed rows feeds the
//             k-tap depthwise FIR (each row is normalised once, 2 FMA per element); result split to fp16 (hi, lo) and
//             stored as the tcgen05 A operand image
//   GEMM      thread 0: 8 K-steps x 3 MMAs (split product), weights of the conv arrive by one bulk copy that was
//             launched under the previous epilogue
//   epilogue  thread = (row, column quarter): X += relu(D + b); the NEXT LayerNorm's statistics are accumulated from the
//             values being written (no extra pass over X)
//   tail      closing LayerNorm -> store, or mask_logits + max over the sequence (model/stage.py:503)
//
// Padded tokens are processed like real ones (the reference never masks before the convs, SURVEY.md §7 quirk i); the
// depthwise conv sees zeros outside [0, L) of its own sequence only.
//
// Algorithmic work per token: n_conv*(2*k*128 + 2*128*128) FLOP; reads 512 B, writes 512 B (mode 0) or 512/L B.
#pragma once
#include "umma.cuh"

namespace stg {

struct EncCfg {
  static constexpr int XS = 132;                       // padded fp32 row stride (floats): conflict-free float4 row access
  static constexpr uint32_t X_BYTES = 128u * XS * 4u;  // 67584
  static constexpr uint32_t LBO_A = 128u * 16u + 16u;
  static constexpr uint32_t A_PART = 16u * LBO_A;      // 33024
  static constexpr uint32_t A_BYTES = 2u * A_PART;     // 66048 >= 128 raw rows (65536)
  static constexpr uint32_t LBO_B = 128u * 16u;
  static constexpr uint32_t W_CHUNK = 2u * 8u * LBO_B;  // one kc=64 chunk: hi | lo
  static constexpr uint32_t W_BYTES = 2u * W_CHUNK;     // 65536
  static constexpr uint32_t OFF_A = (X_BYTES + 127u) & ~127u;
  static constexpr uint32_t OFF_W = OFF_A + ((A_BYTES + 127u) & ~127u);
  static constexpr uint32_t SMEM_BYTES = OFF_W + W_BYTES + 1024u;
  static constexpr int THREADS = 512;
  static constexpr int ROWS_PER_WARP = 8;
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
  long long* timeline;  // optional debug: [32 tiles][16] SM clock stamps of thread 0 of CTA 0 (tools/encoder_timeline.py)
};

#define STG_NEG (-1e10f)
#define ENC_STAMP(idx)                                                                   \
  do {                                                                                   \
    if (p.timeline != nullptr && blockIdx.x == 0 && tid == 0 && tl_tile < 32)             \
      p.timeline[tl_tile * 16 + (idx)] = clock64();                                       \
  } while (0)

template <int KS>
__global__ void __launch_bounds__(512, 1) encoder_kernel(EncDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t w_bar, mma_bar, x_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 row_stat[128];   // (mean * rstd, rstd) of the LayerNorm that reads X next
  __shared__ float row_mask[128];
  __shared__ int row_sid[128];       // sequence index of a tile row (r / L)
  __shared__ float xq[2][4][128];    // cross-quarter exchange of the epilogue: [sum | sum of squares][quarter][row]

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible
  float* X = reinterpret_cast<float*>(smem);
  uint8_t* sA = smem + EncCfg::OFF_A;
  uint8_t* sW = smem + EncCfg::OFF_W;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int P = KS / 2;
  constexpr int RW = EncCfg::ROWS_PER_WARP;
  constexpr int WIN = RW + 2 * P;

  const int L = p.L;
  const int G = 128 / L;  // sequences per tile
  const long long n_tiles = (p.n_seq + G - 1) / G;

  if (tid == 0) {
    mbar_init(&w_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_init(&x_bar, 1);
    mbar_fence_init();
  }
  if (tid < 128) row_sid[tid] = static_cast<int>(tid) / L;
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = make_idesc_f16(128);
  uint32_t w_count = 0, mma_count = 0, x_count = 0;
  int tl_tile = 0;

  auto rows_of = [&](long long tile) -> int {
    const long long seq0 = tile * G;
    const long long n_here = (p.n_seq - seq0) < G ? (p.n_seq - seq0) : G;
    return static_cast<int>(n_here) * L;
  };
  const bool blocked = (p.flags & STAGE_ENC_IN_BLOCKED) != 0;
  const long long seq_floats = blocked ? 128ll * p.S : 128ll * L;  // floats one sequence occupies in x
  auto fetch_tile = [&](long long tile) {  // thread 0: ONE bulk copy brings the whole tile (contiguous in either layout)
    const uint32_t bytes = static_cast<uint32_t>(rows_of(tile) / L) * static_cast<uint32_t>(seq_floats) * 4u;
    mbar_arrive_expect_tx(&x_bar, bytes);
    bulk_g2s(sA, p.x + tile * G * seq_floats, bytes, &x_bar);
  };
  if (tid == 0 && static_cast<long long>(blockIdx.x) < n_tiles) {
    fetch_tile(blockIdx.x);
    mbar_arrive_expect_tx(&w_bar, EncCfg::W_BYTES);
    bulk_g2s(sW, p.pw[0], EncCfg::W_BYTES, &w_bar);
  }

  // position-encoding rows (the same for every tile) and masks of my 8 rows of a tile: plain L2 loads (shared memory
  // takes the whole L1 carve-out), requested one phase ahead so that their latency never sits in phase 0
  float4 pe_r[RW];
  float m_r[RW];
  auto prefetch_rows = [&](long long tile) {
    const int nv = rows_of(tile);
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      const int r = warp * RW + rr;
      pe_r[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
      m_r[rr] = 0.f;
      if (r < nv) {
        const int pos = r - (r / L) * L;
        if (!(p.flags & STAGE_ENC_NO_PE) && !blocked) pe_r[rr] = __ldg(reinterpret_cast<const float4*>(p.pe + pos * 128) + lane);
        if (p.mask != nullptr) m_r[rr] = __ldg(p.mask + tile * G * L + r);
      }
    }
  };
  prefetch_rows(blockIdx.x);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long seq0 = tile * G;
    const int n_valid = rows_of(tile);
    const int n_here = n_valid / L;

    // ---- phase 0: raw rows (smem, prefetched) + pe -> X, first LayerNorm statistics.  warp per row, lane = 4 channels
    //      (the position rows and masks of this tile were requested from L2 under the previous tile's tail)
    ENC_STAMP(0);
    mbar_wait(&x_bar, x_count & 1u, 22);
    ++x_count;
    ENC_STAMP(1);
    if (blocked) {
      // blocked input: thread = (row, column quarter) like the epilogue; raw float4 (row, c4) at (g*32 + c4)*S + pos, the
      // position table at c4*L + pos -> lanes (= rows) read consecutive addresses in shared and global memory alike
      const int r = (warp & 3) * 32 + lane, qd = warp >> 2;
      const int g = row_sid[r], pos = r - g * L;
      const float4* raw = reinterpret_cast<const float4*>(sA) + (g * 32 + qd * 8) * p.S + pos;
      const float4* pe4 = reinterpret_cast<const float4*>(p.pe) + (qd * 8) * L + pos;
      float* xr = X + r * EncCfg::XS + qd * 32;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_valid) {
          v = raw[j * p.S];
          if (!(p.flags & STAGE_ENC_NO_PE)) {
            const float4 e = __ldg(pe4 + j * L);
            v = make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
          }
        }
        *reinterpret_cast<float4*>(xr + j * 4) = v;
        s1 += (v.x + v.y) + (v.z + v.w);
        s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
      }
      xq[0][qd][r] = s1;
      xq[1][qd][r] = s2;
      if (qd == 0) row_mask[r] = (r < n_valid && p.mask != nullptr) ? __ldg(p.mask + seq0 * L + r) : 0.f;
      __syncthreads();
      if (tid < 128) {
        const float mean = ((xq[0][0][tid] + xq[0][1][tid]) + (xq[0][2][tid] + xq[0][3][tid])) * (1.f / 128.f);
        const float ex2 = ((xq[1][0][tid] + xq[1][1][tid]) + (xq[1][2][tid] + xq[1][3][tid])) * (1.f / 128.f);
        const float rs = 1.f / sqrtf(fmaxf(ex2 - mean * mean, 0.f) + 1e-5f);
        row_stat[tid] = make_float2(mean * rs, rs);
      }
    } else {
      const float4* raw = reinterpret_cast<const float4*>(sA);
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) {
        const int r = warp * RW + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_valid) {
          v = raw[r * 32 + lane];
          const float4 e = pe_r[rr];
          v = make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
        }
        *reinterpret_cast<float4*>(X + r * EncCfg::XS + lane * 4) = v;
        const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f);
        if (lane == 0) {
          const float rs = 1.f / sqrtf(var + 1e-5f);
          row_stat[r] = make_float2(mean * rs, rs);
          row_mask[r] = m_r[rr];
        }
      }
    }
    __syncthreads();
    ENC_STAMP(2);

    for (int ci = 0; ci < p.n_conv; ++ci) {
      // ---- conv: LayerNorm applied once per row into a register window, k-tap FIR, split -> A image
      {
        const bool idn = (p.ln_g[ci] == nullptr);  // no LayerNorm in front of this conv
        const float4 g4 = idn ? make_float4(1.f, 1.f, 1.f, 1.f) : __ldg(reinterpret_cast<const float4*>(p.ln_g[ci]) + lane);
        const float4 b4 = idn ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(p.ln_b[ci]) + lane);
        const float4 db = __ldg(reinterpret_cast<const float4*>(p.dw_b[ci]) + lane);
        float wt[4][KS];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < KS; ++j) wt[c][j] = __ldg(p.dw_w[ci] + (lane * 4 + c) * KS + j);
        const int r0 = warp * RW;
        float4 win[WIN];
        int sid[WIN];
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          const int q = r0 - P + i;
          win[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          sid[i] = -1;
          if (q >= 0 && q < 128) {
            const float4 v = *reinterpret_cast<const float4*>(X + q * EncCfg::XS + lane * 4);
            const float2 st = idn ? make_float2(0.f, 1.f) : row_stat[q];
            win[i] = make_float4(fmaf(fmaf(v.x, st.y, -st.x), g4.x, b4.x), fmaf(fmaf(v.y, st.y, -st.x), g4.y, b4.y),
                                 fmaf(fmaf(v.z, st.y, -st.x), g4.z, b4.z), fmaf(fmaf(v.w, st.y, -st.x), g4.w, b4.w));
            sid[i] = row_sid[q];
          }
        }
#pragma unroll
        for (int o = 0; o < RW; ++o) {
          const int r = r0 + o;
          const int me = sid[o + P];
          float o0 = db.x, o1 = db.y, o2 = db.z, o3 = db.w;
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            if (sid[o + j] == me) {  // warp-uniform: the tap lies inside this row's own sequence
              o0 = fmaf(wt[0][j], win[o + j].x, o0);
              o1 = fmaf(wt[1][j], win[o + j].y, o1);
              o2 = fmaf(wt[2][j], win[o + j].z, o2);
              o3 = fmaf(wt[3][j], win[o + j].w, o3);
            }
          }
          uint2 hi, lo;
          split2(o0, o1, hi.x, lo.x);
          split2(o2, o3, hi.y, lo.y);
          uint8_t* dst = sA + (lane >> 1) * EncCfg::LBO_A + r * 16 + (lane & 1) * 8;
          *reinterpret_cast<uint2*>(dst) = hi;
          *reinterpret_cast<uint2*>(dst + EncCfg::A_PART) = lo;
        }
      }
      ENC_STAMP(3 + ci * 5);
      fence_async_smem();
      __syncthreads();
      ENC_STAMP(4 + ci * 5);

      // ---- pointwise GEMM  D[128 x 128] = A[128 x 128] . W^T   (8 K-steps x 3 MMAs)
      if (tid == 0) {
        mbar_wait(&w_bar, w_count & 1u, 20);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t wa = w0 + (ks >> 2) * EncCfg::W_CHUNK + (ks & 3) * 2 * EncCfg::LBO_B;
          const uint64_t a_hi = make_smem_desc(a0 + ks * 2 * EncCfg::LBO_A, EncCfg::LBO_A, 128);
          const uint64_t a_lo = make_smem_desc(a0 + EncCfg::A_PART + ks * 2 * EncCfg::LBO_A, EncCfg::LBO_A, 128);
          const uint64_t b_hi = make_smem_desc(wa, EncCfg::LBO_B, 128);
          const uint64_t b_lo = make_smem_desc(wa + EncCfg::W_CHUNK / 2, EncCfg::LBO_B, 128);
          umma_split3(tbase, a_hi, a_lo, b_hi, b_lo, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&mma_bar);
      }
      ++w_count;
      mbar_wait(&mma_bar, mma_count & 1u, 21);
      ++mma_count;
      tc_fence_after();
      ENC_STAMP(5 + ci * 5);
      if (tid == 0) {
        // weights of the next conv (or conv 0 of this CTA's next tile) stream in under the epilogue; after the last conv
        // the A region is free as well and receives the next tile's raw rows
        const int nxt = ci + 1;
        const bool more = tile + gridDim.x < n_tiles;
        if (nxt < p.n_conv || more) {
          mbar_arrive_expect_tx(&w_bar, EncCfg::W_BYTES);
          bulk_g2s(sW, p.pw[nxt < p.n_conv ? nxt : 0], EncCfg::W_BYTES, &w_bar);
        }
        if (nxt == p.n_conv && more) fetch_tile(tile + gridDim.x);
      }

      // ---- epilogue  X += relu(D + b)   (thread = row x column quarter) + statistics of the new row
      {
        const int r = (warp & 3) * 32 + lane, qd = warp >> 2;
        const int cb = qd * 32;
        const uint32_t trow = tbase + (((warp & 3) * 32u) << 16) + cb;
        float* xr = X + r * EncCfg::XS + cb;
        float v[32];
        tmem_ld16(trow, v);
        tmem_ld16(trow + 16, v + 16);
        tmem_wait_ld();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.pw_b[ci] + cb + j));
          float4 xv = *reinterpret_cast<float4*>(xr + j);
          float4 dv = make_float4(v[j] + bb.x, v[j + 1] + bb.y, v[j + 2] + bb.z, v[j + 3] + bb.w);
          if (!(p.flags & STAGE_ENC_NO_RELU)) dv = make_float4(fmaxf(dv.x, 0.f), fmaxf(dv.y, 0.f), fmaxf(dv.z, 0.f), fmaxf(dv.w, 0.f));
          if (p.flags & STAGE_ENC_NO_RESIDUAL) xv = dv;
          else xv = make_float4(xv.x + dv.x, xv.y + dv.y, xv.z + dv.z, xv.w + dv.w);
          *reinterpret_cast<float4*>(xr + j) = xv;
          s1 += (xv.x + xv.y) + (xv.z + xv.w);
          s2 = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, fmaf(xv.w, xv.w, s2))));
        }
        xq[0][qd][r] = s1;
        xq[1][qd][r] = s2;
      }
      ENC_STAMP(6 + ci * 5);
      tc_fence_before();
      __syncthreads();
      if (tid < 128) {
        const float mean = ((xq[0][0][tid] + xq[0][1][tid]) + (xq[0][2][tid] + xq[0][3][tid])) * (1.f / 128.f);
        const float ex2 = ((xq[1][0][tid] + xq[1][1][tid]) + (xq[1][2][tid] + xq[1][3][tid])) * (1.f / 128.f);
        const float rs = 1.f / sqrtf(fmaxf(ex2 - mean * mean, 0.f) + 1e-5f);
        row_stat[tid] = make_float2(mean * rs, rs);
      }
      __syncthreads();
      ENC_STAMP(7 + ci * 5);
    }
    prefetch_rows(tile + gridDim.x);  // (a tile past the end has no rows: every load is predicated off)

    // ---- closing LayerNorm (+ store, or mask_logits written back to X for the max-pool).  warp per row
    {
      const bool fin = !(p.flags & STAGE_ENC_NO_FINAL_LN);
      const float4 g4 = fin ? __ldg(reinterpret_cast<const float4*>(p.fin_g) + lane) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 b4 = fin ? __ldg(reinterpret_cast<const float4*>(p.fin_b) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) {
        const int r = warp * RW + rr;
        const float4 v = *reinterpret_cast<const float4*>(X + r * EncCfg::XS + lane * 4);
        const float2 st = fin ? row_stat[r] : make_float2(0.f, 1.f);
        float4 o = make_float4(fmaf(fmaf(v.x, st.y, -st.x), g4.x, b4.x), fmaf(fmaf(v.y, st.y, -st.x), g4.y, b4.y),
                               fmaf(fmaf(v.z, st.y, -st.x), g4.z, b4.z), fmaf(fmaf(v.w, st.y, -st.x), g4.w, b4.w));
        if (p.out_mode == 0) {
          if (r < n_valid) *reinterpret_cast<float4*>(p.y + (seq0 * L + r) * 128 + lane * 4) = o;
        } else {
          const float m = row_mask[r], off = (1.f - m) * STG_NEG;  // mask_logits: x*m + (1-m)*(-1e10)
          o = make_float4(o.x * m + off, o.y * m + off, o.z * m + off, o.w * m + off);
          *reinterpret_cast<float4*>(X + r * EncCfg::XS + lane * 4) = o;
        }
      }
    }
    if (p.out_mode != 0) {
      __syncthreads();
      // column-wise max over each sequence: thread = (channel, quarter of the sequence's rows), quarters merged through xq
      const int c = tid & 127, part = tid >> 7;
      const int span = (L + 3) / 4;
      for (int g = 0; g < n_here; ++g) {
        const int lo = g * L + part * span;
        const int hi = min(g * L + L, lo + span);
        float mx = -3.0e38f;
        for (int r = lo; r < hi; ++r) mx = fmaxf(mx, X[r * EncCfg::XS + c]);
        xq[g & 1][part][c] = mx;
        __syncthreads();
        if (part == 0) {
          p.y[(seq0 + g) * 128 + c] = fmaxf(fmaxf(xq[g & 1][0][c], xq[g & 1][1][c]), fmaxf(xq[g & 1][2][c], xq[g & 1][3][c]));
        }
      }
    }
    ENC_STAMP(14);
    __syncthreads();
    ENC_STAMP(15);
    ++tl_tile;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}

}  // namespace stg
