// Multi-head self-attention sub-layer of an EncoderBlock on one SM-resident tile (model/self_attention.py:19-71,
// model/encoder.py:46-52):
//
//   xn = LN_in(x)?                                Q,K,V = xn W{q,k,v}^T + b          (3 GEMMs 128x128x128)
//   S_h = Q_h K_h^T / sqrt(d_k)                   query-row mask: rows with mask == 0 are filled with -1e9, i.e. attend
//   P_h = softmax(S_h)                            uniformly over all L positions; valid rows see padded keys too
//   O   = concat_h(P_h V_h) ;  y = O Wo^T + bo (+ x)? ;  y = LN_fin(y)? ;  optional masked max-pool over L
//
// A tile is 128 rows = floor(128/L) whole sequences (attention is block-diagonal inside the tile: the softmax of a row
// only covers the columns of its own sequence, every other P entry is 0).  All six GEMM families run on tcgen05 with
// the split-fp16 3-term product, accumulators in TMEM:
//   TMEM (512 columns): T0 = Q -> S_0 -> O,  T1 = K -> S_1 -> out-projection,  T2 = V -> S_2,  T3 = S_3
//   shared memory: three 66 KB operand-image regions
//        R1: xn image -> Q image -> P_h image -> O image        (A operands, thread/warp-per-row layout)
//        R2: Wq -> Wv -> K image -> Wo                          (B operands)
//        R3: Wk -> V image (read MN-major for P.V) -> fp32 staging for the max-pool
// 256 threads; warp 0 runs the issue control flow and elects the lane that issues every MMA and bulk copy, epilogues are
// thread = (row, column half).
//
// Algorithmic work per token: 4*2*128*128 + 4*128*L FLOP.
#pragma once
#include "umma.cuh"

namespace stg {

struct MhaCfg {
  static constexpr uint32_t LBO = 128u * 16u + 16u;  // 2064
  static constexpr uint32_t PART = 16u * LBO;        // 33024
  static constexpr uint32_t IMG = 2u * PART;         // 66048
  static constexpr uint32_t LBO_W = 128u * 16u;      // packed weight (kc=64): 2 chunks of [hi 8 slabs | lo 8 slabs]
  static constexpr uint32_t W_CHUNK = 2u * 8u * LBO_W;
  static constexpr uint32_t W_BYTES = 2u * W_CHUNK;  // 65536
  static constexpr uint32_t SMEM_BYTES = 3u * IMG + 1024u;
  static constexpr int THREADS = 256;
  static constexpr int PS = 129;  // fp32 staging row stride of the max-pool (128*129*4 = IMG)
  static_assert(W_BYTES <= IMG && 128u * PS * 4u <= IMG, "weights / pool staging overlay an image region");
};

struct MhaDev {
  const float* x;       // [n_seq, L, 128]
  const float* mask;    // [n_seq, L] or null
  const float* ln_g;    // LayerNorm in front (attn_layer_norm) or null
  const float* ln_b;
  const uint8_t* w[4];  // packed linears.0..3 (kc = 64)
  const float* b[4];
  const float* fin_g;   // closing LayerNorm or null
  const float* fin_b;
  float* y;             // out_mode 0: [n_seq, L, 128]; 1: [n_seq, 128]
  long long n_seq;
  int L, n_heads, residual, out_mode;
};

// TMEM [row, c0 .. c0+64) + bias -> split-fp16 image rows (thread-per-row layout)
__device__ __forceinline__ void mha_acc_to_image(uint32_t taddr, const float* bias, uint8_t* img, int row, int c0) {
#pragma unroll
  for (int cc = 0; cc < 64; cc += 16) {
    float v[16];
    tmem_ld16(taddr + cc, v);
    tmem_wait_ld();
    if (bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + cc + j));
        v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
      }
    }
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
      HL8 hl = split8(v + h8 * 8);
      uint8_t* dst = img + ((c0 + cc) / 8 + h8) * MhaCfg::LBO + row * 16;
      *reinterpret_cast<uint4*>(dst) = hl.hi;
      *reinterpret_cast<uint4*>(dst + MhaCfg::PART) = hl.lo;
    }
  }
}

__global__ void __launch_bounds__(256, 1) mha_kernel(MhaDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t w2_bar, w3_bar, q_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float row_mask[128];
  __shared__ float xq[2][2][128];  // cross-half exchange: [quantity][half][row]

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* R1 = smem;
  uint8_t* R2 = smem + MhaCfg::IMG;
  uint8_t* R3 = smem + 2u * MhaCfg::IMG;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&w2_bar, 1);
    mbar_init(&w3_bar, 1);
    mbar_init(&q_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t T0 = tbase, T1 = tbase + 128, T2 = tbase + 256;
  const uint32_t idesc128 = make_idesc_f16(128);

  const int L = p.L, nh = p.n_heads, dk = 128 / nh;
  const int G = 128 / L;
  const long long n_tiles = (p.n_seq + G - 1) / G;
  const uint32_t r1 = smem_u32(R1), r2 = smem_u32(R2), r3 = smem_u32(R3);
  uint32_t w2_cnt = 0, w3_cnt = 0, q_cnt = 0, mma_cnt = 0;

  // epilogue geometry: thread = (row, column half); a warp may only touch the TMEM lane quadrant warp % 4
  const int row = (warp & 3) * 32 + lane, half = warp >> 2, c0 = half * 64;
  const uint32_t lane_off = ((warp & 3) * 32u) << 16;

  // D[T] = A(image, K-major) . B(packed weight)^T, 8 K-steps
  auto issue_weight_gemm = [&](uint32_t d_tmem, uint32_t a_img, uint32_t w_img) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t wa = w_img + (ks >> 2) * MhaCfg::W_CHUNK + (ks & 3) * 2 * MhaCfg::LBO_W;
      const uint64_t a_hi = make_smem_desc(a_img + ks * 2 * MhaCfg::LBO, MhaCfg::LBO, 128);
      const uint64_t a_lo = make_smem_desc(a_img + MhaCfg::PART + ks * 2 * MhaCfg::LBO, MhaCfg::LBO, 128);
      const uint64_t b_hi = make_smem_desc(wa, MhaCfg::LBO_W, 128);
      const uint64_t b_lo = make_smem_desc(wa + MhaCfg::W_CHUNK / 2, MhaCfg::LBO_W, 128);
      umma_split3(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc128, ks > 0 ? 1u : 0u);
    }
  };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long seq0 = tile * G;
    const int n_here = static_cast<int>((p.n_seq - seq0) < G ? (p.n_seq - seq0) : G);
    const int n_valid = n_here * L;

    if (tid == 0) {  // R2 / R3 are free: the previous tile's out-projection and P.V MMAs have completed
      mbar_arrive_expect_tx(&w2_bar, MhaCfg::W_BYTES);
      bulk_g2s(R2, p.w[0], MhaCfg::W_BYTES, &w2_bar);
      mbar_arrive_expect_tx(&w3_bar, MhaCfg::W_BYTES);
      bulk_g2s(R3, p.w[1], MhaCfg::W_BYTES, &w3_bar);
    }

    // ---- phase 0: rows -> (LayerNorm) -> split-fp16 image R1   (warp per row, lane = 4 channels)
    {
      const bool has_ln = p.ln_g != nullptr;
      const float4 g4 = has_ln ? __ldg(reinterpret_cast<const float4*>(p.ln_g) + lane) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 b4 = has_ln ? __ldg(reinterpret_cast<const float4*>(p.ln_b) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float m = 1.f;
        if (r < n_valid) {
          const long long grow = seq0 * L + r;
          v = __ldg(reinterpret_cast<const float4*>(p.x + grow * 128) + lane);
          if (p.mask != nullptr) m = __ldg(p.mask + grow);
        }
        if (has_ln) {
          const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
          const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
          const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f);
          const float rs = 1.f / sqrtf(var + 1e-5f);
          v = make_float4(a * rs * g4.x + b4.x, b * rs * g4.y + b4.y, c * rs * g4.z + b4.z, d * rs * g4.w + b4.w);
        }
        uint2 hi, lo;
        split2(v.x, v.y, hi.x, lo.x);
        split2(v.z, v.w, hi.y, lo.y);
        uint8_t* dst = R1 + (lane >> 1) * MhaCfg::LBO + r * 16 + (lane & 1) * 8;
        *reinterpret_cast<uint2*>(dst) = hi;
        *reinterpret_cast<uint2*>(dst + MhaCfg::PART) = lo;
        if (lane == 0) row_mask[r] = m;
      }
    }
    fence_async_smem();
    __syncthreads();

    // ---- phase 1: Q -> T0, K -> T1, V -> T2   (w2_cnt / w3_cnt / q_cnt are tracked by warp 0 only)
    if (warp == 0) {  // the whole warp runs the waits, one elected lane issues (no lane-by-lane waterfall around the MMAs)
      tc_fence_after();
      mbar_wait(&w2_bar, w2_cnt & 1u, 60);
      ++w2_cnt;
      if (elect_one()) {
        issue_weight_gemm(T0, r1, r2);
        umma_commit(&q_bar);
      }
      mbar_wait(&w3_bar, w3_cnt & 1u, 61);
      ++w3_cnt;
      if (elect_one()) issue_weight_gemm(T1, r1, r3);
      mbar_wait(&q_bar, q_cnt & 1u, 62);  // Q done: Wq may be replaced by Wv
      ++q_cnt;
      if (elect_one()) {
        mbar_arrive_expect_tx(&w2_bar, MhaCfg::W_BYTES);
        bulk_g2s(R2, p.w[2], MhaCfg::W_BYTES, &w2_bar);
      }
      mbar_wait(&w2_bar, w2_cnt & 1u, 63);
      ++w2_cnt;
      if (elect_one()) {
        issue_weight_gemm(T2, r1, r2);
        umma_commit(&mma_bar);
      }
    }
    mbar_wait(&mma_bar, mma_cnt & 1u, 64);  // all three projections complete
    ++mma_cnt;
    tc_fence_after();

    // ---- phase 2: accumulators + bias -> Q image (R1), K image (R2), V image (R3)
    mha_acc_to_image(T0 + lane_off + c0, p.b[0], R1, row, c0);
    mha_acc_to_image(T1 + lane_off + c0, p.b[1], R2, row, c0);
    mha_acc_to_image(T2 + lane_off + c0, p.b[2], R3, row, c0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- phase 3: S_h = Q_h K_h^T for every head -> T_h
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
      for (int h = 0; h < nh; ++h) {
        for (int ks = 0; ks < dk / 16; ++ks) {
          const uint32_t off = (h * (dk / 8) + ks * 2) * MhaCfg::LBO;
          const uint64_t a_hi = make_smem_desc(r1 + off, MhaCfg::LBO, 128);
          const uint64_t a_lo = make_smem_desc(r1 + MhaCfg::PART + off, MhaCfg::LBO, 128);
          const uint64_t b_hi = make_smem_desc(r2 + off, MhaCfg::LBO, 128);
          const uint64_t b_lo = make_smem_desc(r2 + MhaCfg::PART + off, MhaCfg::LBO, 128);
          umma_split3(tbase + h * 128, a_hi, a_lo, b_hi, b_lo, idesc128, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(&mma_bar);
      }
    }
    mbar_wait(&mma_bar, mma_cnt & 1u, 65);
    ++mma_cnt;
    tc_fence_after();
    if (tid == 0) {  // K image is dead: the out-projection weight streams in under the softmax
      mbar_arrive_expect_tx(&w2_bar, MhaCfg::W_BYTES);
      bulk_g2s(R2, p.w[3], MhaCfg::W_BYTES, &w2_bar);
    }

    // ---- phase 4: per head softmax -> P image (R1) -> O[:, h*dk ..] += P_h V_h
    const int s0 = (row / L) * L;                     // first row/column of my sequence inside the tile
    const bool live = row < n_valid;
    const float qm = row_mask[row];
    const float sc2 = 1.4426950408889634f / sqrtf(static_cast<float>(dk));  // log2(e)/sqrt(d_k): softmax runs in base 2
    const uint32_t idesc_pv = make_idesc_f16(static_cast<uint32_t>(dk), 0, 1);
    for (int h = 0; h < nh; ++h) {
      float s[64];
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) tmem_ld16(tbase + h * 128 + lane_off + c0 + cc, s + cc);
      tmem_wait_ld();
      float mx = -3.0e38f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int c = c0 + j;
        s[j] *= sc2;
        if (c >= s0 && c < s0 + L) mx = fmaxf(mx, s[j]);
      }
      xq[0][half][row] = mx;
      __syncthreads();
      mx = fmaxf(xq[0][0][row], xq[0][1][row]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int c = c0 + j;
        float e = 0.f;
        if (c >= s0 && c < s0 + L) { e = fast_exp2(s[j] - mx); sum += e; }
        s[j] = e;
      }
      xq[1][half][row] = sum;
      __syncthreads();
      const float inv = 1.f / (xq[1][0][row] + xq[1][1][row]);
      const float uni = 1.f / static_cast<float>(L);  // softmax of a constant (-1e9) row
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int c = c0 + j;
        float pj = s[j] * inv;
        if (qm == 0.f) pj = (c >= s0 && c < s0 + L) ? uni : 0.f;
        s[j] = live ? pj : 0.f;
      }
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        HL8 hl = split8(s + c8 * 8);
        uint8_t* dst = R1 + (c0 / 8 + c8) * MhaCfg::LBO + row * 16;
        *reinterpret_cast<uint4*>(dst) = hl.hi;
        *reinterpret_cast<uint4*>(dst + MhaCfg::PART) = hl.lo;
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();  // P_h complete, S_h consumed by every thread
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
        // V image read MN-major: N = the dk channels of head h (8-channel groups are LBO apart), K = key rows (16 rows = 256 B)
        const uint32_t vh = r3 + h * (dk / 8) * MhaCfg::LBO;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a_hi = make_smem_desc(r1 + ks * 2 * MhaCfg::LBO, MhaCfg::LBO, 128);
          const uint64_t a_lo = make_smem_desc(r1 + MhaCfg::PART + ks * 2 * MhaCfg::LBO, MhaCfg::LBO, 128);
          const uint64_t b_hi = make_smem_desc(vh + ks * 256, 128, MhaCfg::LBO);
          const uint64_t b_lo = make_smem_desc(vh + MhaCfg::PART + ks * 256, 128, MhaCfg::LBO);
          umma_split3(T0 + h * dk, a_hi, a_lo, b_hi, b_lo, idesc_pv, ks > 0 ? 1u : 0u);
        }
        umma_commit(&mma_bar);
        }
      }
      mbar_wait(&mma_bar, mma_cnt & 1u, 66);  // P buffer is reused by the next head
      ++mma_cnt;
      tc_fence_after();
    }

    // ---- phase 5: O -> image (R1) -> out-projection -> T1
    mha_acc_to_image(T0 + lane_off + c0, nullptr, R1, row, c0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      mbar_wait(&w2_bar, w2_cnt & 1u, 67);
      ++w2_cnt;
      if (elect_one()) {
        issue_weight_gemm(T1, r1, r2);
        umma_commit(&mma_bar);
      }
    }
    mbar_wait(&mma_bar, mma_cnt & 1u, 68);
    ++mma_cnt;
    tc_fence_after();

    // ---- phase 6: + bias (+ residual) -> (LayerNorm) -> y or masked max-pool
    {
      float o[64];
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) tmem_ld16(T1 + lane_off + c0 + cc, o + cc);
      tmem_wait_ld();
      const long long grow = seq0 * L + row;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b[3] + c0 + j));
        float4 xr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.residual && live) xr = __ldg(reinterpret_cast<const float4*>(p.x + grow * 128 + c0 + j));
        o[j] += bb.x + xr.x; o[j + 1] += bb.y + xr.y; o[j + 2] += bb.z + xr.z; o[j + 3] += bb.w + xr.w;
      }
      if (p.fin_g != nullptr) {  // exact two-pass LayerNorm across the two column halves
        float s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) s1 += o[j];
        xq[0][half][row] = s1;
        __syncthreads();
        const float mean = (xq[0][0][row] + xq[0][1][row]) * (1.f / 128.f);
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { o[j] -= mean; s2 = fmaf(o[j], o[j], s2); }
        xq[1][half][row] = s2;
        __syncthreads();
        const float rs = 1.f / sqrtf((xq[1][0][row] + xq[1][1][row]) * (1.f / 128.f) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 64; ++j) o[j] = o[j] * rs * __ldg(p.fin_g + c0 + j) + __ldg(p.fin_b + c0 + j);
      }
      if (p.out_mode == 0) {
        if (live) {
#pragma unroll
          for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(p.y + grow * 128 + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
      } else {  // mask_logits + max over the sequence (model/stage.py:503); staging in R3 (V is dead)
        float* stage = reinterpret_cast<float*>(R3);
        const float m = qm, off = (1.f - m) * (-1e10f);
#pragma unroll
        for (int j = 0; j < 64; ++j) stage[row * MhaCfg::PS + c0 + j] = o[j] * m + off;
        __syncthreads();
        const int c = tid & 127;
        for (int g = tid >> 7; g < n_here; g += 2) {
          float mx = stage[(g * L) * MhaCfg::PS + c];
          for (int pos = 1; pos < L; ++pos) mx = fmaxf(mx, stage[(g * L + pos) * MhaCfg::PS + c]);
          p.y[(seq0 + g) * 128 + c] = mx;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace stg
