// EncoderBlock on one SM-resident tile:  x + pe -> n_conv x [LN -> depthwise conv -> pointwise GEMM -> ReLU -> +res] -> LN
// (model/encoder.py:29-52).  A tile is 128 rows = floor(128/L) whole sequences; the fp32 residual stream never leaves
// shared memory, the pointwise 128x128 GEMM runs on tcgen05 (split-fp16, accumulator in TMEM), the weights arrive by
// one cp.async.bulk per conv and the next conv's weights are prefetched while the epilogue runs.
//
// Padded tokens are processed like real ones (the reference never masks before the convs, SURVEY.md §7 quirk i); the
// depthwise conv sees zeros outside [0, L) of its own sequence only.
//
// out_mode 0: y[n_seq, L, 128];  out_mode 1: y[n_seq, 128] = max_l mask_logits(LN(x), mask)  (model/stage.py:503).
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
  static constexpr uint32_t A_BYTES = 2u * A_PART;
  static constexpr uint32_t LBO_B = 128u * 16u;
  static constexpr uint32_t W_CHUNK = 2u * 8u * LBO_B;  // one kc=64 chunk: hi | lo
  static constexpr uint32_t W_BYTES = 2u * W_CHUNK;     // 65536
  static constexpr uint32_t OFF_A = (X_BYTES + 127u) & ~127u;
  static constexpr uint32_t OFF_W = OFF_A + ((A_BYTES + 127u) & ~127u);
  static constexpr uint32_t SMEM_BYTES = OFF_W + W_BYTES + 1024u;
  static constexpr int THREADS = 256;
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
};

#define STG_NEG (-1e10f)

template <int KS>
__global__ void __launch_bounds__(256, 1) encoder_kernel(EncDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t w_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float row_mean[128], row_rstd[128], row_mask[128];

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible
  float* X = reinterpret_cast<float*>(smem);
  uint8_t* sA = smem + EncCfg::OFF_A;
  uint8_t* sW = smem + EncCfg::OFF_W;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int P = KS / 2;

  if (tid == 0) {
    mbar_init(&w_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = make_idesc_f16(128);

  const int L = p.L;
  const int G = 128 / L;  // sequences per tile
  const long long n_tiles = (p.n_seq + G - 1) / G;
  uint32_t w_count = 0, mma_count = 0;

  if (tid == 0 && static_cast<long long>(blockIdx.x) < n_tiles) {
    mbar_arrive_expect_tx(&w_bar, EncCfg::W_BYTES);
    bulk_g2s(sW, p.pw[0], EncCfg::W_BYTES, &w_bar);
  }

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long seq0 = tile * G;
    const int n_here = static_cast<int>((p.n_seq - seq0) < G ? (p.n_seq - seq0) : G);
    const int n_valid = n_here * L;

    // ---- phase 0: x + pe -> X (warp per row, 128-bit coalesced)
    for (int rr = 0; rr < 16; ++rr) {
      const int r = warp * 16 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float m = 0.f;
      if (r < n_valid) {
        const int g = r / L, pos = r - g * L;
        const long long grow = (seq0 + g) * L + pos;
        float4 a = __ldg(reinterpret_cast<const float4*>(p.x + grow * 128) + lane);
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!(p.flags & STAGE_ENC_NO_PE)) e = __ldg(reinterpret_cast<const float4*>(p.pe + pos * 128) + lane);
        v = make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w);
        if (p.mask != nullptr) m = __ldg(p.mask + grow);
      }
      *reinterpret_cast<float4*>(X + r * EncCfg::XS + lane * 4) = v;
      if (lane == 0) row_mask[r] = m;
    }
    __syncthreads();

    for (int ci = 0; ci < p.n_conv; ++ci) {
      // ---- phase 1: LayerNorm statistics, exact two-pass, warp per row
      for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        float4 v = *reinterpret_cast<const float4*>(X + r * EncCfg::XS + lane * 4);
        const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f);
        if (lane == 0) {
          row_mean[r] = mean;
          row_rstd[r] = 1.f / sqrtf(var + 1e-5f);
        }
      }
      __syncthreads();

      // ---- phase 2: depthwise conv over LN(x) (LN applied on the fly) -> split-fp16 A image.  lane = 4 channels
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
        for (int rr = 0; rr < 16; ++rr) {
          const int r = warp * 16 + rr;
          const int s0 = (r / L) * L;
          float o0 = db.x, o1 = db.y, o2 = db.z, o3 = db.w;
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const int q = r + j - P;
            if (q >= s0 && q < s0 + L && q < 128) {
              const float4 v = *reinterpret_cast<const float4*>(X + q * EncCfg::XS + lane * 4);
              const float mu = idn ? 0.f : row_mean[q], rs = idn ? 1.f : row_rstd[q];
              o0 = fmaf(wt[0][j], (v.x - mu) * rs * g4.x + b4.x, o0);
              o1 = fmaf(wt[1][j], (v.y - mu) * rs * g4.y + b4.y, o1);
              o2 = fmaf(wt[2][j], (v.z - mu) * rs * g4.z + b4.z, o2);
              o3 = fmaf(wt[3][j], (v.w - mu) * rs * g4.w + b4.w, o3);
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
      fence_async_smem();
      __syncthreads();

      // ---- phase 3: pointwise GEMM  D[128 x 128] = A[128 x 128] . W^T   (8 K-steps x 3 MMAs)
      if (tid == 0) {
        mbar_wait(&w_bar, w_count & 1u, 20);
        ++w_count;
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
      mbar_wait(&mma_bar, mma_count & 1u, 21);
      ++mma_count;
      tc_fence_after();
      if (tid == 0) {  // weights for the next conv (or conv 0 of this CTA's next tile) stream in under the epilogue
        const int nxt = ci + 1;
        if (nxt < p.n_conv || tile + gridDim.x < n_tiles) {
          mbar_arrive_expect_tx(&w_bar, EncCfg::W_BYTES);
          bulk_g2s(sW, p.pw[nxt < p.n_conv ? nxt : 0], EncCfg::W_BYTES, &w_bar);
        }
      }

      // ---- phase 4: epilogue  X += relu(D + b)   (thread = row, warp/4 picks the column half)
      {
        const int r = (warp & 3) * 32 + lane;
        const int cb = (warp >> 2) * 64;
        const uint32_t trow = tbase + (((warp & 3) * 32u) << 16) + cb;
        float* xr = X + r * EncCfg::XS + cb;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.pw_b[ci] + cb + c0 + j));
            float4 xv = *reinterpret_cast<float4*>(xr + c0 + j);
            float4 dv = make_float4(v[j] + bb.x, v[j + 1] + bb.y, v[j + 2] + bb.z, v[j + 3] + bb.w);
            if (!(p.flags & STAGE_ENC_NO_RELU)) dv = make_float4(fmaxf(dv.x, 0.f), fmaxf(dv.y, 0.f), fmaxf(dv.z, 0.f), fmaxf(dv.w, 0.f));
            if (p.flags & STAGE_ENC_NO_RESIDUAL) xv = dv;
            else xv = make_float4(xv.x + dv.x, xv.y + dv.y, xv.z + dv.z, xv.w + dv.w);
            *reinterpret_cast<float4*>(xr + c0 + j) = xv;
          }
        }
      }
      tc_fence_before();
      __syncthreads();
    }

    // ---- final LayerNorm (+ store or masked max-pool)
    {
      const bool fin = !(p.flags & STAGE_ENC_NO_FINAL_LN);
      const float4 g4 = fin ? __ldg(reinterpret_cast<const float4*>(p.fin_g) + lane) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 b4 = fin ? __ldg(reinterpret_cast<const float4*>(p.fin_b) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        float4 v = *reinterpret_cast<const float4*>(X + r * EncCfg::XS + lane * 4);
        const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f);
        const float rs = 1.f / sqrtf(var + 1e-5f);
        float4 o = make_float4(a * rs * g4.x + b4.x, b * rs * g4.y + b4.y, c * rs * g4.z + b4.z, d * rs * g4.w + b4.w);
        if (!fin) o = v;
        if (p.out_mode == 0) {
          if (r < n_valid) {
            const int g = r / L, pos = r - g * L;
            *reinterpret_cast<float4*>(p.y + ((seq0 + g) * L + pos) * 128 + lane * 4) = o;
          }
        } else {
          const float m = row_mask[r], off = (1.f - m) * STG_NEG;  // mask_logits: x*m + (1-m)*(-1e10)
          o = make_float4(o.x * m + off, o.y * m + off, o.z * m + off, o.w * m + off);
          *reinterpret_cast<float4*>(X + r * EncCfg::XS + lane * 4) = o;
        }
      }
    }
    if (p.out_mode != 0) {
      __syncthreads();
      const int c = tid & 127;
      for (int g = tid >> 7; g < n_here; g += 2) {
        float mx = X[(g * L) * EncCfg::XS + c];
        for (int pos = 1; pos < L; ++pos) mx = fmaxf(mx, X[(g * L + pos) * EncCfg::XS + c]);
        p.y[(seq0 + g) * 128 + c] = mx;
      }
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}

}  // namespace stg
