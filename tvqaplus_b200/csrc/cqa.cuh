// Fused QA<->context attention + fusion MLPs, one persistent CTA per (example, frame), looping over the 5 answers.
//
//   S   = (a . ctx^T) * inv|a| * inv|ctx|  - 1e10*(1 - mask)      StructuredAttention.similarity   (cqa.py:95-100)
//   P   = softmax(scale * S) * mask ;  U = P . ctx                 StructuredAttention.forward      (cqa.py:59-61,81)
//   mix = relu(LN384([a, U, a*U]) Wc^T + bc)     (both streams)    qa_ctx_attention / c2q_down_proj (stage.py:381-386)
//   out = LN128(relu(LN384([ms, mv, ms*mv]) Wf^T + bf))            concat_fc                        (stage.py:107-113)
//
// Everything between the fp32 encoder outputs and the statement tile stays on chip:
//   shared memory: split-fp16 images of the frame's ctx tiles (read K-major for S and MN-major for U), the answer
//                  tile, P, and the A / W pipelines of the three 384->128 GEMMs (weights streamed by cp.async.bulk)
//   TMEM (512 columns): T0 = S_sub -> c2q(sub) accumulator, T1 = S_vid -> c2q(vid), T2 = U_sub -> concat accumulator,
//                  T3 = U_vid
// Warp roles (288 threads): warps 0-7 compute (thread = row, warp/4 = column half), warp 8 = MMA issue + weight copies.
//
// Algorithmic work per (answer, frame) tile with Lq rows: 4*128*Lq*(Ls+R) + 3*(2*384*128)*Lq FLOP (x3 executed).
#pragma once
#include "umma.cuh"

namespace stg {

struct CqaCfg {
  static constexpr int LS_MAX = 112, R_MAX = 48;
  static constexpr uint32_t LBO_CS = LS_MAX * 16u + 16u;   // 1808
  static constexpr uint32_t CS_PART = 16u * LBO_CS;        // 28928
  static constexpr uint32_t LBO_CV = R_MAX * 16u + 16u;    // 784
  static constexpr uint32_t CV_PART = 16u * LBO_CV;        // 12544
  static constexpr uint32_t LBO_A = 128u * 16u + 16u;      // 2064: every thread-per-row / warp-per-row A image
  static constexpr uint32_t AQ_PART = 16u * LBO_A;         // 33024
  static constexpr uint32_t PS_PART = (LS_MAX / 8) * LBO_A;  // 28896
  static constexpr uint32_t PV_PART = (R_MAX / 8) * LBO_A;   // 12384
  static constexpr uint32_t AST_PART = 4u * LBO_A;         // 8256: one 32-column A chunk
  static constexpr uint32_t AST_BYTES = 2u * AST_PART;     // 16512
  static constexpr uint32_t W_SLOT = 2u * 4u * 2048u;      // 16384: one kc=32 weight chunk (hi | lo)
  static constexpr int W_SLOTS = 4;
  // regions
  static constexpr uint32_t OFF_CS = 0;
  static constexpr uint32_t OFF_CV = OFF_CS + 2u * CS_PART;            // 57856
  static constexpr uint32_t OFF_AQ = (OFF_CV + 2u * CV_PART + 127u) & ~127u;  // a image; later P_vid + A stages
  static constexpr uint32_t AQ_BYTES = 2u * AQ_PART;                   // 66048
  static constexpr uint32_t OFF_PV = OFF_AQ;                           // P_vid (after S is done)
  static constexpr uint32_t OFF_AST = OFF_AQ + ((2u * PV_PART + 127u) & ~127u);  // 2 A stages behind P_vid
  static constexpr uint32_t OFF_B = (OFF_AQ + AQ_BYTES + 127u) & ~127u;  // P_sub; later the weight ring
  static constexpr uint32_t B_BYTES = W_SLOTS * W_SLOT;                // 65536 >= 2*PS_PART (57792)
  static constexpr uint32_t SMEM_BYTES = OFF_B + B_BYTES + 1024u;
  static constexpr int THREADS = 288;
  static_assert(OFF_AST + 2u * AST_BYTES <= OFF_AQ + AQ_BYTES, "A stages must fit behind P_vid");
  static_assert(2u * PS_PART <= B_BYTES, "P_sub must fit in region B");
};

struct CqaDev {
  int n, t, lq, ls, r, ls_pad, r_pad;
  float scale;
  const float* a_embed; const float* a_mask;
  const float* sub_embed; const float* sub_mask;
  const float* vid_embed; const float* vid_mask;
  const float* c2q_g; const float* c2q_b; const uint8_t* c2q_w; const float* c2q_bias;
  const float* cat_g; const float* cat_b; const uint8_t* cat_w; const float* cat_bias;
  const float* out_g; const float* out_b;
  float* statement; float* statement_mask;
  float* vid_raw_s; float* sub_raw_s; float* vid_norm_s; float* sub_norm_s;
  float* att_vid; float* att_sub; float* mixed_vid; float* mixed_sub;
};

// ---- build a split-fp16 K-major image of `rows` x 128 fp32 rows (warp per row, lane = 4 channels); rows >= n_real are 0.
// Also records 1/max(||row||, 1e-12) and the row mask.
__device__ __forceinline__ void build_rows_image(const float* __restrict__ src, const float* __restrict__ msk, int n_real,
                                                 int n_pad, uint8_t* img, uint32_t lbo, uint32_t part, float* inv_out,
                                                 float* mask_out, int warp, int lane) {
  for (int j = warp; j < n_pad; j += 8) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float m = 0.f;
    if (j < n_real) {
      v = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(j) * 128) + lane);
      m = __ldg(msk + j);
    }
    const float ss = warp_sum((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
    uint2 hi, lo;
    split2(v.x, v.y, hi.x, lo.x);
    split2(v.z, v.w, hi.y, lo.y);
    uint8_t* dst = img + (lane >> 1) * lbo + j * 16 + (lane & 1) * 8;
    *reinterpret_cast<uint2*>(dst) = hi;
    *reinterpret_cast<uint2*>(dst + part) = lo;
    if (lane == 0) {
      inv_out[j] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      mask_out[j] = m;
    }
  }
}

// ---- masked, scaled softmax of one S row held in TMEM -> P image (A operand of U = P.ctx).  NCH = 16-column chunks.
template <int NCH>
__device__ __forceinline__ void softmax_row(uint32_t trow, int n_real, int n_pad, float inv_a, float mask_a,
                                            const float* inv_c, const float* mask_c, float scale, uint8_t* p_img,
                                            uint32_t p_part, int row, float* raw_out, float* norm_out) {
  float s[NCH * 16];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 16 < n_pad) tmem_ld16(trow + c * 16, s + c * 16);
  }
  tmem_wait_ld();
  float mx = -3.0e38f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 16 < n_pad) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c * 16 + j;
        if (col < n_real) {
          const float m = mask_a * mask_c[col];
          float v = s[col] * inv_a * inv_c[col];
          v = v - 1e10f * (1.f - m);  // exactly the reference's S - 1e10*(1 - S_mask)
          if (raw_out != nullptr) raw_out[col] = v;
          v *= scale;
          s[col] = v;
          mx = fmaxf(mx, v);
        }
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 16 < n_pad) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c * 16 + j;
        if (col < n_real) {
          const float e = expf(s[col] - mx);
          s[col] = e;
          sum += e;
        } else {
          s[col] = 0.f;
        }
      }
    }
  }
  const float inv_sum = 1.f / sum;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 16 < n_pad) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c * 16 + j;
        if (col < n_real) {
          const float pj = s[col] * inv_sum * (mask_a * mask_c[col]);
          s[col] = pj;
          if (norm_out != nullptr) norm_out[col] = pj;
        }
      }
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        HL8 hl = split8(s + c * 16 + h8 * 8);
        uint8_t* dst = p_img + (c * 2 + h8) * CqaCfg::LBO_A + row * 16;
        *reinterpret_cast<uint4*>(dst) = hl.hi;
        *reinterpret_cast<uint4*>(dst + p_part) = hl.lo;
      }
    }
  }
}

// ---- one 16-column slice of the two GEMM-input operands p, q  (concat row = [p, q, p*q])
template <bool CONCAT>
__device__ __forceinline__ void load_pq16(int cc, const float (&a_reg)[64], uint32_t tP, uint32_t tQ,
                                          const float* __restrict__ bias, int col0, float* p, float* q) {
  if (CONCAT) {
    tmem_ld16(tP + cc * 16, p);
    tmem_ld16(tQ + cc * 16, q);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float b = __ldg(bias + col0 + cc * 16 + j);
      p[j] = fmaxf(p[j] + b, 0.f);
      q[j] = fmaxf(q[j] + b, 0.f);
    }
  } else {
    tmem_ld16(tQ + cc * 16, q);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = a_reg[cc * 16 + j];
  }
}

// ---- LayerNorm(384) over [p, q, p*q] + split-fp16 A-chunk production for one 384->128 GEMM.
// Thread = (row, half): owns columns [64*half, 64*half+64) of each 128-wide segment, i.e. 6 of the 12 K=32 chunks.
template <bool CONCAT>
__device__ __forceinline__ void produce_gemm(const float (&a_reg)[64], uint32_t tP, uint32_t tQ, const float* bias,
                                             const float* __restrict__ ln_g, const float* __restrict__ ln_b, int row,
                                             int half, float (*xch)[128], uint8_t* a_stage, uint64_t* a_full,
                                             uint64_t* a_empty, uint32_t& a_use, float* dbg_p, float* dbg_q) {
  const int col0 = half * 64;
  float s = 0.f;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    float p[16], q[16];
    load_pq16<CONCAT>(cc, a_reg, tP, tQ, bias, col0, p, q);
#pragma unroll
    for (int j = 0; j < 16; ++j) s += (p[j] + q[j]) + p[j] * q[j];
    if (dbg_q != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(dbg_q + col0 + cc * 16 + j) = make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]);
    }
    if (CONCAT && dbg_p != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(dbg_p + col0 + cc * 16 + j) = make_float4(p[j], p[j + 1], p[j + 2], p[j + 3]);
    }
  }
  xch[half][row] = s;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float mean = (xch[0][row] + xch[1][row]) * (1.f / 384.f);
  float qq = 0.f;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    float p[16], q[16];
    load_pq16<CONCAT>(cc, a_reg, tP, tQ, bias, col0, p, q);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float d0 = p[j] - mean, d1 = q[j] - mean, d2 = p[j] * q[j] - mean;
      qq += (d0 * d0 + d1 * d1) + d2 * d2;
    }
  }
  xch[2 + half][row] = qq;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float rstd = 1.f / sqrtf((xch[2][row] + xch[3][row]) * (1.f / 384.f) + 1e-5f);

#pragma unroll
  for (int pi = 0; pi < 6; ++pi) {
    const int seg = pi >> 1, sub = pi & 1;
    mbar_wait(a_empty, (a_use & 1u) ^ 1u, 30);
    ++a_use;
    const int kbase = seg * 128 + col0 + sub * 32;  // first concat column of this chunk
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      float p[16], q[16], v[16];
      load_pq16<CONCAT>(sub * 2 + c2, a_reg, tP, tQ, bias, col0, p, q);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = seg == 0 ? p[j] : (seg == 1 ? q[j] : p[j] * q[j]);
        const int k = kbase + c2 * 16 + j;
        v[j] = (x - mean) * rstd * __ldg(ln_g + k) + __ldg(ln_b + k);
      }
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        HL8 hl = split8(v + h8 * 8);
        uint8_t* dst = a_stage + (c2 * 2 + h8) * CqaCfg::LBO_A + row * 16;
        *reinterpret_cast<uint4*>(dst) = hl.hi;
        *reinterpret_cast<uint4*>(dst + CqaCfg::AST_PART) = hl.lo;
      }
    }
    fence_async_smem();
    mbar_arrive(a_full);
  }
}

__global__ void __launch_bounds__(288, 1) cqa_kernel(CqaDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t op_ready, mma_done, a_full[2], a_empty[2], w_full[CqaCfg::W_SLOTS], w_empty[CqaCfg::W_SLOTS];
  __shared__ uint32_t tmem_slot;
  __shared__ float inv_cs[CqaCfg::LS_MAX], mask_cs[CqaCfg::LS_MAX], inv_cv[CqaCfg::R_MAX], mask_cv[CqaCfg::R_MAX];
  __shared__ float inv_a[128], mask_a[128];
  __shared__ float xch[4][128];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sCS = smem + CqaCfg::OFF_CS;
  uint8_t* sCV = smem + CqaCfg::OFF_CV;
  uint8_t* sAQ = smem + CqaCfg::OFF_AQ;
  uint8_t* sPV = smem + CqaCfg::OFF_PV;
  uint8_t* sAST = smem + CqaCfg::OFF_AST;
  uint8_t* sB = smem + CqaCfg::OFF_B;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&op_ready, 256);
    mbar_init(&mma_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < CqaCfg::W_SLOTS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t T0 = tbase, T1 = tbase + 128, T2 = tbase + 256, T3 = tbase + 384;

  const int n_frames = p.n * p.t;

  if (warp < 8) {
    // =================================================================== compute warps
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const uint32_t lane_off = ((warp & 3) * 32u) << 16;
    uint32_t done_cnt = 0, a_use = 0;
    for (int frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
      const int n_i = frame / p.t, t_i = frame - n_i * p.t;
      // ---- frame setup: ctx images (shared by the 5 answers)
      build_rows_image(p.sub_embed + static_cast<size_t>(frame) * p.ls * 128, p.sub_mask + static_cast<size_t>(frame) * p.ls,
                       p.ls, p.ls_pad, sCS, CqaCfg::LBO_CS, CqaCfg::CS_PART, inv_cs, mask_cs, warp, lane);
      build_rows_image(p.vid_embed + static_cast<size_t>(frame) * p.r * 128, p.vid_mask + static_cast<size_t>(frame) * p.r,
                       p.r, p.r_pad, sCV, CqaCfg::LBO_CV, CqaCfg::CV_PART, inv_cv, mask_cv, warp, lane);
      for (int a_i = 0; a_i < 5; ++a_i) {
        const size_t na = static_cast<size_t>(n_i) * 5 + a_i;
        const size_t tile = (na * p.t + t_i);  // index of this (n, a, t) tile
        // ---- A: answer tile image (rows >= lq are zero)
        build_rows_image(p.a_embed + na * p.lq * 128, p.a_mask + na * p.lq, p.lq, 128, sAQ, CqaCfg::LBO_A,
                         CqaCfg::AQ_PART, inv_a, mask_a, warp, lane);
        fence_async_smem();
        mbar_arrive(&op_ready);  // E1
        // ---- B: softmax (warps 0-3: sub stream, warps 4-7: vid stream), thread = row
        mbar_wait(&mma_done, done_cnt & 1u, 31);  // D1: S ready
        ++done_cnt;
        tc_fence_after();
        {
          const float ia = inv_a[row], ma = mask_a[row];
          const bool live = row < p.lq;
          if (half == 0) {
            float* raw = (p.sub_raw_s && live) ? p.sub_raw_s + (tile * p.lq + row) * p.ls : nullptr;
            float* nrm = (p.sub_norm_s && live) ? p.sub_norm_s + (tile * p.lq + row) * p.ls : nullptr;
            softmax_row<7>(T0 + lane_off, p.ls, p.ls_pad, ia, ma, inv_cs, mask_cs, p.scale, sB, CqaCfg::PS_PART, row, raw, nrm);
          } else {
            float* raw = (p.vid_raw_s && live) ? p.vid_raw_s + (tile * p.lq + row) * p.r : nullptr;
            float* nrm = (p.vid_norm_s && live) ? p.vid_norm_s + (tile * p.lq + row) * p.r : nullptr;
            softmax_row<3>(T1 + lane_off, p.r, p.r_pad, ia, ma, inv_cv, mask_cv, p.scale, sPV, CqaCfg::PV_PART, row, raw, nrm);
          }
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(&op_ready);  // E2
        // ---- my 64 answer values (columns [64*half, +64) of row), kept in registers for both c2q GEMMs
        float a_reg[64];
        {
          const float4* src = reinterpret_cast<const float4*>(p.a_embed + (na * p.lq + row) * 128 + half * 64);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float4 v = (row < p.lq) ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            a_reg[j * 4] = v.x; a_reg[j * 4 + 1] = v.y; a_reg[j * 4 + 2] = v.z; a_reg[j * 4 + 3] = v.w;
          }
        }
        // ---- C: c2q for both streams
        mbar_wait(&mma_done, done_cnt & 1u, 32);  // D2: U ready
        ++done_cnt;
        tc_fence_after();
        const uint32_t coff = lane_off + half * 64;
        const bool live = row < p.lq;
        const size_t orow = (tile * p.lq + row) * 128;
        produce_gemm<false>(a_reg, 0, T2 + coff, nullptr, p.c2q_g, p.c2q_b, row, half, xch, sAST + half * CqaCfg::AST_BYTES,
                            &a_full[half], &a_empty[half], a_use, nullptr, (p.att_sub && live) ? p.att_sub + orow : nullptr);
        produce_gemm<false>(a_reg, 0, T3 + coff, nullptr, p.c2q_g, p.c2q_b, row, half, xch, sAST + half * CqaCfg::AST_BYTES,
                            &a_full[half], &a_empty[half], a_use, nullptr, (p.att_vid && live) ? p.att_vid + orow : nullptr);
        // ---- D: concat_fc input from the two c2q accumulators
        tc_fence_before();
        mbar_wait(&mma_done, done_cnt & 1u, 33);  // D3: both c2q accumulators complete
        ++done_cnt;
        tc_fence_after();
        produce_gemm<true>(a_reg, T0 + coff, T1 + coff, p.c2q_bias, p.cat_g, p.cat_b, row, half, xch,
                           sAST + half * CqaCfg::AST_BYTES, &a_full[half], &a_empty[half], a_use,
                           (p.mixed_sub && live) ? p.mixed_sub + orow : nullptr, (p.mixed_vid && live) ? p.mixed_vid + orow : nullptr);
        // ---- E: relu(acc + b) -> LN128 -> statement
        tc_fence_before();
        mbar_wait(&mma_done, done_cnt & 1u, 34);  // D4: concat accumulator complete
        ++done_cnt;
        tc_fence_after();
        {
          float s = 0.f;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float v[16];
            tmem_ld16(T2 + coff + cc * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) s += fmaxf(v[j] + __ldg(p.cat_bias + half * 64 + cc * 16 + j), 0.f);
          }
          xch[half][row] = s;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float mean = (xch[0][row] + xch[1][row]) * (1.f / 128.f);
          float qq = 0.f;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float v[16];
            tmem_ld16(T2 + coff + cc * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float d = fmaxf(v[j] + __ldg(p.cat_bias + half * 64 + cc * 16 + j), 0.f) - mean;
              qq += d * d;
            }
          }
          xch[2 + half][row] = qq;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float rstd = 1.f / sqrtf((xch[2][row] + xch[3][row]) * (1.f / 128.f) + 1e-5f);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float v[16];
            tmem_ld16(T2 + coff + cc * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = half * 64 + cc * 16 + j;
              v[j] = (fmaxf(v[j] + __ldg(p.cat_bias + c), 0.f) - mean) * rstd * __ldg(p.out_g + c) + __ldg(p.out_b + c);
            }
            if (live) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(p.statement + orow + half * 64 + cc * 16 + j) =
                    make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
          if (live && half == 0) {
            float cnt = 0.f;
            for (int j = 0; j < p.r; ++j) cnt += mask_a[row] * mask_cv[j];  // S_mask.sum(-1) != 0  (stage.py:386)
            p.statement_mask[tile * p.lq + row] = (cnt != 0.f) ? 1.f : 0.f;
          }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");  // xch / inv_a / mask_a are rewritten by the next answer
      }
    }
  } else if (lane == 0) {
    // =================================================================== MMA issue + weight ring (one thread)
    uint32_t ready_cnt = 0, done_cnt = 0, q_cnt = 0, a_cnt[2] = {0, 0};
    const uint32_t id_ssub = make_idesc_f16(p.ls_pad), id_svid = make_idesc_f16(p.r_pad);
    const uint32_t id_u = make_idesc_f16(128, 0, 1), id_g = make_idesc_f16(128);
    const uint32_t aq = smem_u32(sAQ), cs = smem_u32(sCS), cv = smem_u32(sCV), ps = smem_u32(sB), pv = smem_u32(sPV);
    for (int frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
      for (int a_i = 0; a_i < 5; ++a_i) {
        // ---- S = a . ctx^T for both streams
        mbar_wait(&op_ready, ready_cnt & 1u, 40);  // E1
        ++ready_cnt;
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a_hi = make_smem_desc(aq + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
          const uint64_t a_lo = make_smem_desc(aq + CqaCfg::AQ_PART + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
          const uint64_t s_hi = make_smem_desc(cs + ks * 2 * CqaCfg::LBO_CS, CqaCfg::LBO_CS, 128);
          const uint64_t s_lo = make_smem_desc(cs + CqaCfg::CS_PART + ks * 2 * CqaCfg::LBO_CS, CqaCfg::LBO_CS, 128);
          umma_split3(T0, a_hi, a_lo, s_hi, s_lo, id_ssub, ks > 0 ? 1u : 0u);
          const uint64_t v_hi = make_smem_desc(cv + ks * 2 * CqaCfg::LBO_CV, CqaCfg::LBO_CV, 128);
          const uint64_t v_lo = make_smem_desc(cv + CqaCfg::CV_PART + ks * 2 * CqaCfg::LBO_CV, CqaCfg::LBO_CV, 128);
          umma_split3(T1, a_hi, a_lo, v_hi, v_lo, id_svid, ks > 0 ? 1u : 0u);
        }
        umma_commit(&mma_done);  // D1
        ++done_cnt;
        // ---- U = P . ctx  (ctx image read MN-major: LBO = 128 between 8-row k groups, SBO = image LBO between n groups)
        mbar_wait(&op_ready, ready_cnt & 1u, 41);  // E2
        ++ready_cnt;
        tc_fence_after();
        for (int ks = 0; ks < p.ls_pad / 16; ++ks) {
          const uint64_t a_hi = make_smem_desc(ps + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
          const uint64_t a_lo = make_smem_desc(ps + CqaCfg::PS_PART + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
          const uint64_t b_hi = make_smem_desc(cs + ks * 256, 128, CqaCfg::LBO_CS);
          const uint64_t b_lo = make_smem_desc(cs + CqaCfg::CS_PART + ks * 256, 128, CqaCfg::LBO_CS);
          umma_split3(T2, a_hi, a_lo, b_hi, b_lo, id_u, ks > 0 ? 1u : 0u);
        }
        for (int ks = 0; ks < p.r_pad / 16; ++ks) {
          const uint64_t a_hi = make_smem_desc(pv + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
          const uint64_t a_lo = make_smem_desc(pv + CqaCfg::PV_PART + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
          const uint64_t b_hi = make_smem_desc(cv + ks * 256, 128, CqaCfg::LBO_CV);
          const uint64_t b_lo = make_smem_desc(cv + CqaCfg::CV_PART + ks * 256, 128, CqaCfg::LBO_CV);
          umma_split3(T3, a_hi, a_lo, b_hi, b_lo, id_u, ks > 0 ? 1u : 0u);
        }
        umma_commit(&mma_done);  // D2
        mbar_wait(&mma_done, done_cnt & 1u, 42);  // region B (P_sub) is free once U has completed
        ++done_cnt;
        // ---- three 384->128 GEMMs, 12 K=32 chunks each; chunk order alternates between the two column halves
        auto w_src = [&](int q) -> const uint8_t* {
          const int g = q / 12, ql = q % 12, pi = ql >> 1, h = ql & 1;
          const int c = (pi >> 1) * 4 + h * 2 + (pi & 1);  // K chunk of the 384-wide weight
          return (g < 2 ? p.c2q_w : p.cat_w) + static_cast<size_t>(c) * CqaCfg::W_SLOT;
        };
        for (int q = 0; q < CqaCfg::W_SLOTS - 1; ++q) {  // prime the ring
          const uint32_t slot = (q_cnt + q) % CqaCfg::W_SLOTS, use = (q_cnt + q) / CqaCfg::W_SLOTS;
          mbar_wait(&w_empty[slot], (use & 1u) ^ 1u, 43);
          mbar_arrive_expect_tx(&w_full[slot], CqaCfg::W_SLOT);
          bulk_g2s(sB + slot * CqaCfg::W_SLOT, w_src(q), CqaCfg::W_SLOT, &w_full[slot]);
        }
        for (int q = 0; q < 36; ++q) {
          const int g = q / 12, ql = q % 12, h = ql & 1;
          const uint32_t slot = (q_cnt + q) % CqaCfg::W_SLOTS, wuse = (q_cnt + q) / CqaCfg::W_SLOTS;
          mbar_wait(&a_full[h], a_cnt[h] & 1u, 45);
          ++a_cnt[h];
          mbar_wait(&w_full[slot], wuse & 1u, 46);
          tc_fence_after();
          const uint32_t ast = smem_u32(sAST + h * CqaCfg::AST_BYTES), wsl = smem_u32(sB + slot * CqaCfg::W_SLOT);
          const uint32_t dcol = g == 0 ? T0 : (g == 1 ? T1 : T2);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t a_hi = make_smem_desc(ast + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
            const uint64_t a_lo = make_smem_desc(ast + CqaCfg::AST_PART + ks * 2 * CqaCfg::LBO_A, CqaCfg::LBO_A, 128);
            const uint64_t b_hi = make_smem_desc(wsl + ks * 2 * 2048, 2048, 128);
            const uint64_t b_lo = make_smem_desc(wsl + CqaCfg::W_SLOT / 2 + ks * 2 * 2048, 2048, 128);
            umma_split3(dcol, a_hi, a_lo, b_hi, b_lo, id_g, (ql > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&a_empty[h]);
          umma_commit(&w_empty[slot]);
          if (q == 23 || q == 35) {
            umma_commit(&mma_done);  // D3 / D4
            ++done_cnt;
          }
          if (q + CqaCfg::W_SLOTS - 1 < 36) {  // refill the slot chunk q-1 used (chunk q keeps the tensor pipe busy meanwhile)
            const uint32_t qq = q_cnt + q + CqaCfg::W_SLOTS - 1;
            const uint32_t rslot = qq % CqaCfg::W_SLOTS, ruse = qq / CqaCfg::W_SLOTS;
            mbar_wait(&w_empty[rslot], (ruse & 1u) ^ 1u, 44);
            mbar_arrive_expect_tx(&w_full[rslot], CqaCfg::W_SLOT);
            bulk_g2s(sB + rslot * CqaCfg::W_SLOT, w_src(q + CqaCfg::W_SLOTS - 1), CqaCfg::W_SLOT, &w_full[rslot]);
          }
        }
        q_cnt += 36;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tbase);
}

}  // namespace stg
