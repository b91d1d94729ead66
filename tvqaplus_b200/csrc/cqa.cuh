// Fused QA<->context attention + fusion MLPs, one persistent CTA per (example, frame), looping over the 128-row tiles
// into which the 5 x Lq answer rows of the example are packed (4 tiles instead of 5 at Lq = 100).
//
//   S   = (a . ctx^T) * inv|a| * inv|ctx|  - 1e10*(1 - mask)      StructuredAttention.similarity   (cqa.py:95-100)
//   P   = softmax(scale * S) * mask ;  U = P . ctx                 StructuredAttention.forward      (cqa.py:59-61,81)
//   mix = relu(LN384([a, U, a*U]) Wc^T + bc)     (both streams)    qa_ctx_attention / c2q_down_proj (stage.py:381-386)
//   out = LN128(relu(LN384([ms, mv, ms*mv]) Wf^T + bf))            concat_fc                        (stage.py:107-113)
//
// LayerNorm(384) is folded into the GEMM that follows it:
//        LN(x) W^T + b = rstd * (x (W*gamma)^T - mean * cs) + cb ,  cs = (W*gamma) 1,  cb = W beta + b
// so the A operand is the RAW [p, q, p*q] row (one pass over TMEM, statistics accumulated on the fly), and the
// frame-invariant third of c2q, Za = a (Wc*gamma)[:, :128]^T, is computed once per (example, answer), loaded under the
// U MMAs and written into both c2q accumulators with tcgen05.st (executed FLOPs of c2q drop by a third; algorithmic FLOPs
// are still quoted in full).  The softmax is normalised late: P' = exp2(scale*log2e*S - shift) goes to the tensor cores
// un-normalised and U = P'.ctx is divided by the row sum when it is read back, so the softmax phase needs no barrier.
//
// v4 (round 2).  What paced v3 (measured, profiles/r2a/cqa_tl.txt): with both operands in shared memory an M=128, N=128
// MMA reads 128 B per clock — all the shared-memory bandwidth an SM has — so the producers' stores, the weight fills and
// the MMAs starved one another (12 MMAs of a K=64 stage took 1.15 k cycles instead of 768), and the issuing warp then
// waited another ~0.6 k per stage for a weight slot that could only be refilled after its previous MMAs had completed.
// The tile was also one strictly serial chain (softmax -> U -> c2q -> concat -> epilogue).  Now:
//   * A operands are converted IN PLACE in tensor memory: a thread reads its 16 fp32 accumulator cells (U, or the c2q
//     accumulator after activation), splits them into fp16 (hi, lo) pairs and writes the 8 + 8 packed columns back over
//     the very cells it read (tcgen05.st).  The next GEMM takes them as a TS-form A operand (A from tensor memory), so
//     10 of the 14 K=64 stages of a tile never touch shared memory on the A side, need no ring, no a_empty wait and no
//     async-proxy fence.  Only the a*U segments of c2q (no free TMEM columns while U and both accumulators are live)
//     still go through the two shared-memory stages.
//   * The concat product segment ms*mv is rebuilt from the packed (hi + lo) cells of T0 / T1 and written to T3.
//   * Four epilogue warps (thread = row) own LN128 + the statement store; the compute warps hand over two scalars per
//     row and go straight to the next tile, whose S MMAs are issued right behind the concat MMAs (the answer image has
//     been able to land since the last shared-memory A stage was consumed, i.e. half a tile earlier).
//   * c2q(sub) and c2q(vid) complete on separate barriers, so the ms segment is produced while c2q(vid) drains.
//
// Everything between the fp32 encoder outputs and the statement tile stays on chip:
//   shared memory: split-fp16 images of the frame's ctx tiles (one cp.async.bulk per frame; read K-major for S and
//                  MN-major for U), the answer image (one bulk copy per tile), P', two K=64 A stages (a*U) and a 4-slot
//                  ring of K=32 weight chunks for the three 384->128 GEMMs
//   TMEM (512 columns): T0 = S_sub -> c2q(sub) accumulator -> ms operand,  T1 = S_vid -> c2q(vid) -> mv operand,
//                  T2 = U_sub -> U_sub operand -> concat accumulator,  T3 = U_vid -> U_vid operand -> ms*mv operand
// Warp roles (768 threads, register budgets set per role with setmaxnreg): warp 0 = MMA issue (never waits for a
// completion; it carries the lowest warp id because the schedulers favour older warps and the issuer must not starve
// behind busy compute warps), warp 1 = weight ring copies, warp 2 = ctx / answer image copies, warp 3 idle, warps 4-19
// compute (thread = row x column quarter, 96 registers: everything is produced 16 columns at a time; shared memory takes
// the whole L1 carve-out, a spill would be an L2 round trip), warps 20-23 = epilogue.
// The 5 x Lq answer rows of an example are packed into 128-row tiles (4 tiles instead of 5 at Lq = 100); per-thread
// global accesses (fp32 answers, Za, and optionally the statement) use a blocked [4-column group][row] layout so that a
// warp's 32 rows touch whole lines.
//
// Algorithmic work per (answer, frame) with Lq rows: 4*128*Lq*(Ls+R) + 3*(2*384*128)*Lq FLOP.
#pragma once
#include "umma.cuh"

namespace stg {

struct CqaCfg {
  static constexpr int LS_MAX = 112, R_MAX = 48;
  static constexpr uint32_t LBO_CS = LS_MAX * 16u + 16u;     // 1808
  static constexpr uint32_t CS_PART = 16u * LBO_CS;          // 28928
  static constexpr uint32_t LBO_CV = R_MAX * 16u + 16u;      // 784
  static constexpr uint32_t CV_PART = 16u * LBO_CV;          // 12544
  static constexpr uint32_t LBO_A = 128u * 16u + 16u;        // 2064: every thread-per-row / warp-per-row A image
  static constexpr uint32_t AQ_PART = 16u * LBO_A;           // 33024
  static constexpr uint32_t PS_PART = (LS_MAX / 8) * LBO_A;  // 28896
  static constexpr uint32_t PV_PART = (R_MAX / 8) * LBO_A;   // 12384
  static constexpr uint32_t AST_PART = 8u * LBO_A;           // 16512: one K=64 A stage (16 columns from each column quarter)
  static constexpr uint32_t AST_BYTES = 2u * AST_PART;       // 33024: hi | lo
  static constexpr uint32_t W_SLOT = 2u * 4u * 2048u;        // 16384: one kc=32 weight chunk (hi | lo)
  static constexpr int W_SLOTS = 4;
  static constexpr int AQ_W_SLOTS = 4;  // four more weight slots inside the AQ region between the last a*U stage and the next answer image
  // per-frame ctx blob (global and shared layout are identical)
  static constexpr uint32_t CTX_CS = 0;                                  // sub image hi | lo
  static constexpr uint32_t CTX_CS_INV = 2u * CS_PART;                   // 57856: inv norms [112]
  static constexpr uint32_t CTX_CS_MASK = CTX_CS_INV + LS_MAX * 4u;      // masks [112]
  static constexpr uint32_t CTX_CV = CTX_CS_MASK + LS_MAX * 4u;          // 58752: vid image hi | lo
  static constexpr uint32_t CTX_CV_INV = CTX_CV + 2u * CV_PART;          // 83840
  static constexpr uint32_t CTX_CV_MASK = CTX_CV_INV + R_MAX * 4u;
  static constexpr uint32_t CTX_BYTES = CTX_CV_MASK + R_MAX * 4u;        // 84224
  // per-(example, answer) blob: image hi | lo, then inv|a|, mask, sum(a), sum(a^2) per row
  static constexpr uint32_t AQ_BYTES = 2u * AQ_PART;                     // 66048
  static constexpr uint32_t ABLOB_ROWDATA = AQ_BYTES;
  static constexpr uint32_t ABLOB_BYTES = AQ_BYTES + 4u * 128u * 4u;     // 68096
  // shared-memory regions
  static constexpr uint32_t OFF_CTX = 0;
  static constexpr uint32_t OFF_AQ = (CTX_BYTES + 127u) & ~127u;         // a image; later P_vid, then the 4 A stages
  static constexpr uint32_t OFF_B = (OFF_AQ + AQ_BYTES + 127u) & ~127u;  // P_sub; later the weight ring
  static constexpr uint32_t B_BYTES = W_SLOTS * W_SLOT;                  // 65536 >= 2*PS_PART (57792)
  static constexpr uint32_t SMEM_BYTES = OFF_B + B_BYTES + 1024u;
  static constexpr int THREADS = 768;        // MMA, weight copies, ctx/answer copies, idle | 16 compute warps | 4 epilogue warps
  static constexpr int CHUNKS = 8 + 12;  // K=32 weight chunks per tile: c2q (each used by both streams), concat
  static_assert(2u * AST_BYTES <= AQ_BYTES, "the two A stages overlay the answer image");
  static_assert(2u * PV_PART <= AQ_BYTES && 2u * PS_PART <= B_BYTES, "P images overlay AQ / the weight ring");
  static_assert(CTX_BYTES % 16 == 0 && ABLOB_BYTES % 16 == 0 && CTX_CV % 16 == 0, "bulk copies / descriptors need 16 B");
  static_assert(CHUNKS == 20 && W_SLOTS == 4 && AQ_W_SLOTS == 4, "the slot schedule below is written for 8 + 12 chunks, 4 + 4 slots");
  static_assert(AQ_W_SLOTS * W_SLOT <= AQ_BYTES, "the extra weight slots live inside the AQ region");
  // Slot schedule of a tile's 20 weight chunks: c2q chunks 0..7 and concat chunks 8..11, 16..19 go through the 4-slot ring
  // in region B (slot j % 4: four uses per slot and tile); concat chunks 12..15 (segment mv) sit in the AQ region, which is
  // idle between the last a*U stage (D3) and the next answer image.  All eight concat chunks of segments ms / mv are thus
  // requested at D3, while the compute warps are still producing the first concat operand.
  __host__ __device__ static constexpr bool w_in_aq(int j) { return j >= 12 && j < 16; }
  __host__ __device__ static constexpr int w_use(int j) { return j < 4 ? 0 : (j < 8 ? 1 : (j < 12 ? 2 : 3)); }
};

struct CqaDev {
  int n, t, lq, ls, r, ls_pad, r_pad;
  int n_pt;                   // 128-row tiles per (example, frame): the 5*lq answer rows of an example are packed
  float scale;
  const float* a_blk;         // fp32 answers, blocked [n][32][5*lq] float4 (for the a*U products; lanes = rows read 512 B runs)
  const uint8_t* a_blob;      // [n*n_pt] ABLOB_BYTES: packed answer rows [128*j, 128*j + 128) of example n
  const uint8_t* ctx_blob;    // [n*t] CTX_BYTES
  const float* za;            // a (Wc*gamma)[:, :128]^T in the same blocked layout [n][32][5*lq] float4
  const uint8_t* c2q_w;       // packed (Wc*gamma)[:, 128:384], kc=32: 8 chunks
  const float* c2q_cs; const float* c2q_cb;
  const uint8_t* cat_w;       // packed (Wf*gamma), kc=32: 12 chunks
  const float* cat_cs; const float* cat_cb;
  const float* out_g; const float* out_b;
  float* statement; float* statement_mask;
  int stmt_stride;            // 0: row-major statement; else S of the blocked tile layout
  float* vid_raw_s; float* sub_raw_s; float* vid_norm_s; float* sub_norm_s;
  float* att_vid; float* att_sub; float* mixed_vid; float* mixed_sub;
  long long* timeline;  // optional [8][64][16] clock64 stamps of block 0, first 64 tiles (tools/cqa_timeline.py)
};

// ------------------------------------------------------------------------------------------------------------------
// prep: fp32 rows -> split-fp16 operand images + per-row data (one warp per row, lane = 4 channels)
// ------------------------------------------------------------------------------------------------------------------
struct CqaPrepDev {
  int n, t, lq, ls, r, ls_pad, r_pad, n_pt;
  const float* a_embed; const float* a_mask;
  const float* sub_embed; const float* sub_mask;
  const float* vid_embed; const float* vid_mask;
  uint8_t* a_blob; uint8_t* ctx_blob;
  float* a_blk;  // fp32 copy of a_embed in the blocked layout [example][4-column group (32)][packed row (5*lq)] float4
  int sub_prepared;  // the sub-title parts of the ctx blobs already exist (written by the input encoder): skip them
};

__global__ void __launch_bounds__(256) cqa_prep_kernel(CqaPrepDev p) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_a = static_cast<long long>(p.n) * p.n_pt * 128;
  const long long n_s = p.sub_prepared ? 0 : static_cast<long long>(p.n) * p.t * p.ls_pad;
  const long long n_v = static_cast<long long>(p.n) * p.t * p.r_pad;
  if (warp >= n_a + n_s + n_v) return;
  const float* src = nullptr;
  float4* blk_dst = nullptr;
  float m = 0.f;
  uint8_t* img;
  uint32_t lbo, part;
  float *inv_out, *mask_out, *sum_out = nullptr, *sq_out = nullptr;
  int j;
  if (warp < n_a) {
    const long long nj = warp / 128;  // (example, packed tile)
    j = static_cast<int>(warp % 128);
    const long long ex = nj / p.n_pt;
    const int prow = static_cast<int>(nj - ex * p.n_pt) * 128 + j;  // packed row of the example: answer = prow / lq
    uint8_t* blob = p.a_blob + nj * CqaCfg::ABLOB_BYTES;
    img = blob; lbo = CqaCfg::LBO_A; part = CqaCfg::AQ_PART;
    float* rd = reinterpret_cast<float*>(blob + CqaCfg::ABLOB_ROWDATA);
    inv_out = rd; mask_out = rd + 128; sum_out = rd + 256; sq_out = rd + 384;
    if (prow < 5 * p.lq) {  // a_embed / a_mask are [n,5,lq,..] = [n, 5*lq, ..]
      src = p.a_embed + (ex * 5 * p.lq + prow) * 128;
      m = __ldg(p.a_mask + ex * 5 * p.lq + prow);
      blk_dst = reinterpret_cast<float4*>(p.a_blk) + (ex * 32 + lane) * (5 * p.lq) + prow;
    }
  } else if (warp < n_a + n_s) {
    const long long w = warp - n_a, f = w / p.ls_pad;
    j = static_cast<int>(w % p.ls_pad);
    uint8_t* blob = p.ctx_blob + f * CqaCfg::CTX_BYTES;
    img = blob + CqaCfg::CTX_CS; lbo = CqaCfg::LBO_CS; part = CqaCfg::CS_PART;
    inv_out = reinterpret_cast<float*>(blob + CqaCfg::CTX_CS_INV);
    mask_out = reinterpret_cast<float*>(blob + CqaCfg::CTX_CS_MASK);
    if (j < p.ls) { src = p.sub_embed + (f * p.ls + j) * 128; m = __ldg(p.sub_mask + f * p.ls + j); }
  } else {
    const long long w = warp - n_a - n_s, f = w / p.r_pad;
    j = static_cast<int>(w % p.r_pad);
    uint8_t* blob = p.ctx_blob + f * CqaCfg::CTX_BYTES;
    img = blob + CqaCfg::CTX_CV; lbo = CqaCfg::LBO_CV; part = CqaCfg::CV_PART;
    inv_out = reinterpret_cast<float*>(blob + CqaCfg::CTX_CV_INV);
    mask_out = reinterpret_cast<float*>(blob + CqaCfg::CTX_CV_MASK);
    if (j < p.r) { src = p.vid_embed + (f * p.r + j) * 128; m = __ldg(p.vid_mask + f * p.r + j); }
  }
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (src != nullptr) v = __ldg(reinterpret_cast<const float4*>(src) + lane);
  if (blk_dst != nullptr) *blk_dst = v;
  const float ss = warp_sum((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
  const float sm = warp_sum((v.x + v.y) + (v.z + v.w));
  uint2 hi, lo;
  split2(v.x, v.y, hi.x, lo.x);
  split2(v.z, v.w, hi.y, lo.y);
  uint8_t* dst = img + (lane >> 1) * lbo + j * 16 + (lane & 1) * 8;
  *reinterpret_cast<uint2*>(dst) = hi;
  *reinterpret_cast<uint2*>(dst + part) = lo;
  if (lane == 0) {
    // F.normalize: x / max(||x||, eps); a masked row / column stores 0 so the kernel needs one load, not two
    inv_out[j] = (m != 0.f) ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 0.f;
    mask_out[j] = m;
    if (sum_out != nullptr) { sum_out[j] = sm; sq_out[j] = ss; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// compute-side helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reg_alloc_compute() { asm volatile("setmaxnreg.inc.sync.aligned.u32 96;"); }
__device__ __forceinline__ void reg_release_issue() { asm volatile("setmaxnreg.dec.sync.aligned.u32 56;"); }
__device__ __forceinline__ void reg_release_epilogue() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); }

// 16 fp32 values of this thread's row -> fp16 (hi, lo) pairs written over the 16 TMEM cells they came from: columns
// [c, c+8) hold the 16 hi halves (two per 32-bit column, k ascending), [c+8, c+16) the lo halves.  A TS-form MMA reads
// them as one K=16 step of the A operand (hi at c, lo at c + 8).
__device__ __forceinline__ void put_operand16(uint32_t taddr, const float* v) {
  const HL8 h0 = split8(v), h1 = split8(v + 8);
  tmem_st8(taddr, h0.hi, h1.hi);
  tmem_st8(taddr + 8, h0.lo, h1.lo);
}
// the inverse: 16 packed cells -> 16 floats (hi + lo reproduces the value the split saw to 2^-22)
__device__ __forceinline__ void get_operand16(uint32_t taddr, float* v) {
  float raw[16];
  tmem_ld16(taddr, raw);
  tmem_wait_ld();
  const uint32_t* r = reinterpret_cast<const uint32_t*>(raw);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&r[j]));
    const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&r[8 + j]));
    const float2 s = add2(h, l);
    v[2 * j] = s.x;
    v[2 * j + 1] = s.y;
  }
}

// phase stamps exist only in the TIMELINE instantiation (tools/cqa_timeline.py); the production kernel carries none
#define CQA_STAMP(who, idx)                                                                         \
  do {                                                                                              \
    if constexpr (TIMELINE) {                                                                       \
      if (p.timeline != nullptr && blockIdx.x == 0 && tl_tile < 64 && (who) < 8)                    \
        p.timeline[((who) * 64 + tl_tile) * 16 + (idx)] = clock64();                                \
    }                                                                                               \
  } while (0)

template <bool TIMELINE>
__global__ void __launch_bounds__(768, 1) cqa_kernel(CqaDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // Completion events of a tile, each with its OWN barrier (a thread that wakes up late from a suspended wait must never
  // find its barrier two phases further on): mma_done[0] D1 S, [1] D2v U_vid, [2] D3 c2q, [3] D2s U_sub, [4] D4 concat.
  __shared__ uint64_t op_ready, mma_done[5], ctx_full, aq_full, ctx_free, aq_free, b_free, a_full[2], a_empty[2],
      useg_full[2], cat_full[3], ep_ready, t2_free, w_full[CqaCfg::W_SLOTS], w_empty[CqaCfg::W_SLOTS],
      wq_full[CqaCfg::AQ_W_SLOTS], aq_w_free;
  __shared__ uint32_t tmem_slot;
  __shared__ float xq[4][4][128];  // cross-quarter exchange: [quantity][quarter][row]
  __shared__ float2 ep_stat[128];  // (rstd, -rstd*mean) of the concat LayerNorm(384) row, compute -> epilogue warps
  __shared__ __align__(16) float pvec[6][128];  // c2q cs, c2q cb, cat cs, cat cb, out gamma, out beta

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible
  uint8_t* sCTX = smem + CqaCfg::OFF_CTX;
  uint8_t* sCS = sCTX + CqaCfg::CTX_CS;
  uint8_t* sCV = sCTX + CqaCfg::CTX_CV;
  const float* inv_cs = reinterpret_cast<const float*>(sCTX + CqaCfg::CTX_CS_INV);
  const float* mask_cs = reinterpret_cast<const float*>(sCTX + CqaCfg::CTX_CS_MASK);
  const float* inv_cv = reinterpret_cast<const float*>(sCTX + CqaCfg::CTX_CV_INV);
  const float* mask_cv = reinterpret_cast<const float*>(sCTX + CqaCfg::CTX_CV_MASK);
  uint8_t* sAQ = smem + CqaCfg::OFF_AQ;  // a image -> P_vid -> the two a*U stages
  uint8_t* sB = smem + CqaCfg::OFF_B;    // P_sub -> weight ring
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&op_ready, 512);
    for (int i = 0; i < 5; ++i) mbar_init(&mma_done[i], 1);
    mbar_init(&ctx_full, 1);
    mbar_init(&aq_full, 1);
    mbar_init(&ctx_free, 1);
    mbar_init(&aq_free, 1);
    mbar_init(&b_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 512);
      mbar_init(&a_empty[i], 1);
      mbar_init(&useg_full[i], 512);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&cat_full[i], 512);
    mbar_init(&ep_ready, 512);
    mbar_init(&t2_free, 128);
    for (int i = 0; i < CqaCfg::W_SLOTS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < CqaCfg::AQ_W_SLOTS; ++i) mbar_init(&wq_full[i], 1);
    mbar_init(&aq_w_free, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 128; i += CqaCfg::THREADS) {  // per-column vectors of the folded GEMM epilogues
    pvec[0][i] = p.c2q_cs[i]; pvec[1][i] = p.c2q_cb[i]; pvec[2][i] = p.cat_cs[i];
    pvec[3][i] = p.cat_cb[i]; pvec[4][i] = p.out_g[i]; pvec[5][i] = p.out_b[i];
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t T0 = tbase, T1 = tbase + 128, T2 = tbase + 256, T3 = tbase + 384;
  const int n_frames = p.n * p.t;
  const int my_frames = static_cast<int>(blockIdx.x) < n_frames ? (n_frames - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  const int my_tiles = my_frames * p.n_pt;

  if (warp >= 4 && warp < 20) {
    reg_alloc_compute();
    // =================================================================== compute warps
    // thread = (row, quarter): row = 32*(warp & 3) + lane (the TMEM lane quadrant a warp may touch is warp % 4),
    // quarter = warp / 4 owns columns [32*quarter, +32) of every 128-wide segment = exactly one K=32 chunk per segment
    const int row = (warp & 3) * 32 + lane, qd = (warp >> 2) - 1;
    const uint32_t lane_off = ((warp & 3) * 32u) << 16;
    const uint32_t coff = lane_off + qd * 32;
    const uint32_t ctid = tid - 128;  // index among the compute threads
    uint32_t done_cnt = 0, a_use = 0, frame_cnt = 0;
    // timeline build: thread 0 stamps rows 0 (phases) and 2 (a*U stages); thread 480 (warp 15) rows 4 / 6; thread 160
    // (warp 5) row 7 (phases only)
    [[maybe_unused]] int tl_tile = (ctid == 0 || ctid == 160 || ctid == 480) ? 0 : 64;
    [[maybe_unused]] const int tl_ph = ctid == 0 ? 0 : (ctid == 480 ? 4 : 7), tl_st = ctid == 0 ? 2 : (ctid == 480 ? 6 : 8);
    float any_v = 0.f, any_s = 0.f;  // does this frame have a valid region / subtitle token?  (S_mask.sum(-1) != 0, stage.py:386)

    // The a*U segment of a c2q GEMM = two K=64 stages in shared memory; stage h holds columns [32*quarter + 16*h, +16) of
    // every quarter (the weight columns of that segment are packed in the same order).  CQA_SEG_PUT publishes my 16 columns
    // of stage h: wait until the MMAs of the previous segment have drained the stage, split, store, fence, arrive.
#define CQA_SEG_PUT(h, v, code)                                                              \
  {                                                                                          \
    const HL8 hl0 = split8((v)), hl1 = split8((v) + 8);                                      \
    uint8_t* dst = sAQ + (h) * CqaCfg::AST_BYTES + (qd * 2) * CqaCfg::LBO_A + row * 16;      \
    if (st_seq < 8) CQA_STAMP(tl_st, st_seq * 2);                                            \
    mbar_wait(&a_empty[h], (a_use & 1u) ^ 1u, code);                                         \
    *reinterpret_cast<uint4*>(dst) = hl0.hi;                                                 \
    *reinterpret_cast<uint4*>(dst + CqaCfg::AST_PART) = hl0.lo;                              \
    *reinterpret_cast<uint4*>(dst + CqaCfg::LBO_A) = hl1.hi;                                 \
    *reinterpret_cast<uint4*>(dst + CqaCfg::LBO_A + CqaCfg::AST_PART) = hl1.lo;              \
    fence_async_smem();                                                                      \
    mbar_arrive(&a_full[h]);                                                                 \
    if (st_seq < 8) CQA_STAMP(tl_st, st_seq * 2 + 1);                                        \
    ++st_seq;                                                                                \
  }

    for (int frame = blockIdx.x; frame < n_frames; frame += gridDim.x, ++frame_cnt) {
      const int n_i = frame / p.t, t_i = frame - n_i * p.t;
      for (int a_i = 0; a_i < p.n_pt; ++a_i) {
        // my row of this tile: packed row prow of example n_i = row (prow % lq) of answer (prow / lq)
        const int prow = a_i * 128 + row;
        const bool live = prow < 5 * p.lq;
        const int ans = live ? prow / p.lq : 0;
        const size_t nj = static_cast<size_t>(n_i) * p.n_pt + a_i;
        // my first float4 of a_blk / za (blocked [example][4-column group][packed row]): group stride = 5*lq float4
        const uint32_t arow = (static_cast<uint32_t>(n_i) * 32u + qd * 8u) * (5u * p.lq) + (live ? prow : 0);
        const uint32_t srow = ((static_cast<uint32_t>(n_i) * 5u + ans) * p.t + t_i) * p.lq + (prow - ans * p.lq);  // output row
#define orow (static_cast<size_t>(srow) * 128 + qd * 32) /* my 32 columns of the output row */
        CQA_STAMP(tl_ph, 0);
        [[maybe_unused]] int st_seq = 0;  // a*U stage counter of this tile (timeline build)
        const float* rd = reinterpret_cast<const float*>(p.a_blob + nj * CqaCfg::ABLOB_BYTES + CqaCfg::ABLOB_ROWDATA);
        const float ia = __ldg(rd + row), ma = __ldg(rd + 128 + row);
        if (a_i == 0) {
          mbar_wait(&ctx_full, frame_cnt & 1u, 30);  // ctx blob (images + inv norms + masks) has landed
          any_v = 0.f;
          any_s = 0.f;
          for (int j = 0; j < p.r; ++j) any_v += mask_cv[j];
          for (int j = 0; j < p.ls; ++j) any_s += mask_cs[j];
        }

        // ---- softmax of both streams, normalisation deferred: P' = exp2(scale*log2e*S - shift) goes to the tensor cores
        //      un-normalised and U = P'.ctx is divided by the row sum when it is read back (the partial sums of the four
        //      quarters meet in shared memory under the U MMAs, so this phase has no block barrier at all).
        //      The 7 sub + 3 vid score chunks (16 columns each) are dealt to the quarters as
        //      q0: sub 0,1,2   q1: sub 3,4,5   q2: sub 6, vid 0   q3: vid 1, 2
        CQA_STAMP(tl_ph, 1);
        mbar_wait(&mma_done[0], done_cnt & 1u, 31);  // D1: S ready
        tc_fence_after();
        CQA_STAMP(tl_ph, 2);
        {
          float* raw_s = (p.sub_raw_s && live) ? p.sub_raw_s + static_cast<size_t>(srow) * p.ls : nullptr;
          float* raw_v = (p.vid_raw_s && live) ? p.vid_raw_s + static_cast<size_t>(srow) * p.r : nullptr;
          float* nrm_s = (p.sub_norm_s && live) ? p.sub_norm_s + static_cast<size_t>(srow) * p.ls : nullptr;
          float* nrm_v = (p.vid_norm_s && live) ? p.vid_norm_s + static_cast<size_t>(srow) * p.r : nullptr;
#define CQA_SLOT(k)                                                                   \
  const bool vid = (qd == 3) || (qd == 2 && (k) == 1);                                 \
  const int ch = qd < 2 ? 3 * qd + (k) : (qd == 2 ? ((k) == 0 ? 6 : 0) : (k) + 1);      \
  const int n_real = vid ? p.r : p.ls;                                                 \
  const bool on = (qd < 2 || (k) < 2) && ch * 16 < (vid ? p.r_pad : p.ls_pad);
          // |cos| <= 1, so sc2 = scale*log2(e) bounds every live score: shifting by sc2 - 15 keeps P' inside fp16 range
          // ((0, 2^15]) without knowing the row maximum as long as sc2 <= 15 (a row whose best cosine is c peaks at
          // 2^(15 - sc2*(1-c)) >= 2^-15).  Larger scales (the reference's docstring suggests 100) take the exact-maximum
          // path: the row peaks at exactly 2^15 and whatever underflows fp16 is < 2^-29 of the row sum.
          const float sc2 = p.scale * 1.4426950408889634f;
          float shift_s = sc2 - 15.f, shift_v = sc2 - 15.f;
          if (sc2 > 15.f) {
            float mxs = -3.0e38f, mxv = -3.0e38f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              CQA_SLOT(k)
              (void)n_real;
              if (on) {
                float sv[16];
                tmem_ld16((vid ? T1 : T0) + lane_off + ch * 16, sv);
                tmem_wait_ld();
                const float* ic = (vid ? inv_cv : inv_cs) + ch * 16;
                float mx = -3.0e38f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float w = ia * ic[j];
                  if (w != 0.f) mx = fmaxf(mx, sv[j] * w * sc2);
                }
                if (vid) mxv = fmaxf(mxv, mx); else mxs = fmaxf(mxs, mx);
              }
            }
            xq[0][qd][row] = mxs;
            xq[1][qd][row] = mxv;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // a row without a live column (padded answer token, frame without sub-title / regions) keeps the static shift:
            // its arguments stay at -1e10*sc2, exp2 gives exactly 0 and U = 0 (with the -3e38 sentinel the shift would turn
            // them into +3e38 -> inf -> NaN in the split)
            const float ms_ = fmaxf(fmaxf(xq[0][0][row], xq[0][1][row]), fmaxf(xq[0][2][row], xq[0][3][row]));
            const float mv_ = fmaxf(fmaxf(xq[1][0][row], xq[1][1][row]), fmaxf(xq[1][2][row], xq[1][3][row]));
            shift_s = (ms_ > -1e30f ? ms_ : sc2) - 15.f;
            shift_v = (mv_ > -1e30f ? mv_ : sc2) - 15.f;
            asm volatile("bar.sync 1, 512;" ::: "memory");
          }
          float sms = 0.f, smv = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            CQA_SLOT(k)
            if (on) {
              float sv[16];
              tmem_ld16((vid ? T1 : T0) + lane_off + ch * 16, sv);
              tmem_wait_ld();
              const float4* ic4 = reinterpret_cast<const float4*>((vid ? inv_cv : inv_cs) + ch * 16);
              float* raw = vid ? raw_v : raw_s;
              float* nrm = vid ? nrm_v : nrm_s;
              const float shift = vid ? shift_v : shift_s;
              float2 sum2 = make_float2(0.f, 0.f);
              const float2 ia2 = dup2(ia), sc22 = dup2(sc2), nsh2 = dup2(-shift);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 c4 = ic4[j4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = j4 * 4 + e * 2;
                  // S - 1e10*(1 - S_mask) of the reference: |S| <= 1 is absorbed by -1e10 in fp32, so a select is exact.
                  // Padded columns (>= n_real) carry inverse norm 0 and behave like masked ones (weight exactly 0).
                  const float2 w = mul2(ia2, e == 0 ? lo2(c4) : hi2(c4));
                  float2 v = mul2(make_float2(sv[j], sv[j + 1]), w);
                  v.x = (w.x != 0.f) ? v.x : -1e10f;
                  v.y = (w.y != 0.f) ? v.y : -1e10f;
                  if (raw != nullptr) {
                    if (ch * 16 + j < n_real) raw[ch * 16 + j] = v.x;
                    if (ch * 16 + j + 1 < n_real) raw[ch * 16 + j + 1] = v.y;
                  }
                  const float2 arg = fma2(v, sc22, nsh2);  // masked / padded: exp2(-1.4e11) = 0
                  const float2 ex = make_float2(fast_exp2(arg.x), fast_exp2(arg.y));
                  sum2 = add2(sum2, ex);
                  sv[j] = ex.x;
                  sv[j + 1] = ex.y;
                }
              }
              const float sum = sum2.x + sum2.y;
              if (nrm != nullptr) {  // optional output: normalised after the row sums are known (see below)
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (ch * 16 + j < n_real) nrm[ch * 16 + j] = sv[j];
              }
              if (vid) smv += sum; else sms += sum;
              uint8_t* img = vid ? sAQ : sB;                       // P_vid overlays the answer image, P_sub the weight ring
              const uint32_t part = vid ? CqaCfg::PV_PART : CqaCfg::PS_PART;
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                HL8 hl = split8(sv + h8 * 8);
                uint8_t* dst = img + (ch * 2 + h8) * CqaCfg::LBO_A + row * 16;
                *reinterpret_cast<uint4*>(dst) = hl.hi;
                *reinterpret_cast<uint4*>(dst + part) = hl.lo;
              }
            }
          }
          xq[0][qd][row] = sms;  // (xq[0..1] were last read before this tile's predecessor passed its final barrier)
          xq[1][qd][row] = smv;
        }
        CQA_STAMP(tl_ph, 3);
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(&op_ready);  // E2: P' images and the partial row sums are in place, S has been consumed
        CQA_STAMP(tl_ph, 4);
        // ---- my 32 answer values (columns [32*quarter, +32) of the row); the loads fly while U = P.ctx runs
        float a_reg[32];
        {
          const float4* src = reinterpret_cast<const float4*>(p.a_blk) + arow;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = live ? __ldg(src + j * (5 * p.lq)) : make_float4(0.f, 0.f, 0.f, 0.f);
            a_reg[j * 4] = v.x; a_reg[j * 4 + 1] = v.y; a_reg[j * 4 + 2] = v.z; a_reg[j * 4 + 3] = v.w;
          }
        }
        // ---- the frame-invariant a.Wa term of c2q (same for both streams): its loads fly under the U MMAs, then it is
        //      written into both c2q accumulators (my own cells of T0 / T1; S there is dead once U has completed) and the
        //      c2q MMAs accumulate on top of it
        {
          float za[32];
          const float4* zsrc = reinterpret_cast<const float4*>(p.za) + arow;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = live ? __ldg(zsrc + j * (5 * p.lq)) : make_float4(0.f, 0.f, 0.f, 0.f);
            za[j * 4] = v.x; za[j * 4 + 1] = v.y; za[j * 4 + 2] = v.z; za[j * 4 + 3] = v.w;
          }
          mbar_wait(&mma_done[1], done_cnt & 1u, 32);  // D2v: U_vid ready (issued first: its 6 MMAs hide under nothing else)
          tc_fence_after();
          CQA_STAMP(tl_ph, 5);
          tmem_st16(T0 + coff, za);
          tmem_st16(T0 + coff + 16, za + 16);
          tmem_st16(T1 + coff, za);
          tmem_st16(T1 + coff + 16, za + 16);
        }
        // U completed => every thread has passed E2 => all partial row sums are in xq[0..1].  Masked entries of a live row
        // are exactly 0; a fully masked row must become 0 after the "* S_mask" of the reference -> its 1/sum is replaced by 0
        const float is = (ma != 0.f && any_s != 0.f) ? 1.f / ((xq[0][0][row] + xq[0][1][row]) + (xq[0][2][row] + xq[0][3][row])) : 0.f;
        const float iv = (ma != 0.f && any_v != 0.f) ? 1.f / ((xq[1][0][row] + xq[1][1][row]) + (xq[1][2][row] + xq[1][3][row])) : 0.f;
        if (p.sub_norm_s != nullptr || p.vid_norm_s != nullptr) {  // optional output: normalise what this thread stored above
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            CQA_SLOT(k)
            float* nrm = vid ? p.vid_norm_s : p.sub_norm_s;
            if (on && live && nrm != nullptr) {
              nrm += static_cast<size_t>(srow) * n_real;
              for (int j = 0; j < 16; ++j)
                if (ch * 16 + j < n_real) nrm[ch * 16 + j] *= (vid ? iv : is);
            }
          }
        }
#undef CQA_SLOT
        // ---- c2q A operands of both streams, 16 columns at a time: the U segment goes back IN PLACE over the cells it came
        //      from (TS-form operand, natural K order), the a*U segment through the two shared-memory stages; LayerNorm(384)
        //      statistics on the fly.  (The stores of Za above are ordered before the first c2q MMA by the fence that
        //      precedes my first arrival below.)
        // Both streams are produced side by side, half by half, because the c2q weights are the same for both: every weight
        // chunk that lands in the ring is used twice (sub, then vid).  Stage g of the shared-memory ring belongs to stream g.
        float2 s1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        // The vid stream goes first in every half: U_vid (6 MMAs) completes ~1.4 k cycles before U_sub (21 MMAs), so its
        // first operands are produced while the tensor pipe still works on U_sub.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            const int g = 1 - gi;
            if (h == 0 && g == 0) {
              mbar_wait(&mma_done[3], done_cnt & 1u, 32);  // D2s: U_sub ready
              tc_fence_after();
            }
            const uint32_t tU = (g == 0 ? T2 : T3) + coff;
            float* att = g == 0 ? p.att_sub : p.att_vid;
            const float2 isum2 = dup2(g == 0 ? is : iv);
            float u[16];
            tmem_ld16(tU + h * 16, u);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 t = mul2(make_float2(u[j], u[j + 1]), isum2);
              s1[g] = add2(s1[g], t);
              s2[g] = fma2(t, t, s2[g]);
              u[j] = t.x;
              u[j + 1] = t.y;
            }
            if (att != nullptr && live) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(att + orow + h * 16 + j) = make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]);
            }
            put_operand16(tU + h * 16, u);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 t = mul2(make_float2(u[j], u[j + 1]), make_float2(a_reg[h * 16 + j], a_reg[h * 16 + j + 1]));
              s1[g] = add2(s1[g], t);
              s2[g] = fma2(t, t, s2[g]);
              u[j] = t.x;
              u[j + 1] = t.y;
            }
            tmem_wait_st();
            tc_fence_before();  // Za (first pass) and my operand cells are in tensor memory before I arrive anywhere
            CQA_SEG_PUT(g, u, 33)
          }
          ++a_use;
          mbar_arrive(&useg_full[h]);  // my cells of this half of both U operands are in place
        }
        float st_s1[2], st_s2[2];
        st_s1[0] = s1[0].x + s1[0].y; st_s2[0] = s2[0].x + s2[0].y;
        st_s1[1] = s1[1].x + s1[1].y; st_s2[1] = s2[1].x + s2[1].y;
        CQA_STAMP(tl_ph, 6);
        float r_s, r_v, nm_s, nm_v;  // rstd and -rstd*mean of the two LayerNorm(384) rows
        {
          const float asum = __ldg(rd + 256 + row), asq = __ldg(rd + 384 + row);  // sum(a), sum(a^2) of my row
          xq[0][qd][row] = st_s1[0]; xq[1][qd][row] = st_s2[0];
          xq[2][qd][row] = st_s1[1]; xq[3][qd][row] = st_s2[1];
          asm volatile("bar.sync 1, 512;" ::: "memory");
          const float t1 = asum + ((xq[0][0][row] + xq[0][1][row]) + (xq[0][2][row] + xq[0][3][row]));
          const float t2 = asq + ((xq[1][0][row] + xq[1][1][row]) + (xq[1][2][row] + xq[1][3][row]));
          const float mu_s = t1 * (1.f / 384.f);
          r_s = 1.f / sqrtf(fmaxf(t2 * (1.f / 384.f) - mu_s * mu_s, 0.f) + 1e-5f);
          nm_s = -r_s * mu_s;
          const float v1 = asum + ((xq[2][0][row] + xq[2][1][row]) + (xq[2][2][row] + xq[2][3][row]));
          const float v2 = asq + ((xq[3][0][row] + xq[3][1][row]) + (xq[3][2][row] + xq[3][3][row]));
          const float mu_v = v1 * (1.f / 384.f);
          r_v = 1.f / sqrtf(fmaxf(v2 * (1.f / 384.f) - mu_v * mu_v, 0.f) + 1e-5f);
          nm_v = -r_v * mu_v;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");  // everyone has read xq before it is reused below
        CQA_STAMP(tl_ph, 7);

        // ---- concat_fc A operands: ms, mv = relu(rstd*(acc - mean*cs) + cb) go back in place over their accumulators
        //      (T0 / T1) as TS-form operands; the product segment is rebuilt from those packed cells and written to T3
        //      (U_vid's operand there is dead once c2q(vid) has completed).
        float2 c1 = make_float2(0.f, 0.f), c2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t tC = (g == 0 ? T0 : T1) + coff;
          const float2 rr = dup2(g == 0 ? r_s : r_v), nm = dup2(g == 0 ? nm_s : nm_v);
          float* mixed = g == 0 ? p.mixed_sub : p.mixed_vid;
          if (g == 0) {
            mbar_wait(&mma_done[2], done_cnt & 1u, 35);  // D3: both c2q accumulators are complete
            tc_fence_after();
          }
          CQA_STAMP(tl_ph, 8 + 2 * g);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float m[16];
            tmem_ld16(tC + h * 16, m);
            tmem_wait_ld();
            const float4* cs4 = reinterpret_cast<const float4*>(&pvec[0][qd * 32 + h * 16]);
            const float4* cb4 = reinterpret_cast<const float4*>(&pvec[1][qd * 32 + h * 16]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 cs = cs4[j4], cb = cb4[j4];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = j4 * 4 + e * 2;
                const float2 t = max2(fma2(rr, make_float2(m[j], m[j + 1]), fma2(nm, e == 0 ? lo2(cs) : hi2(cs), e == 0 ? lo2(cb) : hi2(cb))), 0.f);
                c1 = add2(c1, t);
                c2 = fma2(t, t, c2);
                m[j] = t.x;
                m[j + 1] = t.y;
              }
            }
            if (live && mixed != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(mixed + orow + h * 16 + j) = make_float4(m[j], m[j + 1], m[j + 2], m[j + 3]);
            }
            put_operand16(tC + h * 16, m);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&cat_full[g]);
          CQA_STAMP(tl_ph, 9 + 2 * g);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float m[16], w[16];
          get_operand16(T0 + coff + h * 16, m);
          get_operand16(T1 + coff + h * 16, w);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 t = mul2(make_float2(m[j], m[j + 1]), make_float2(w[j], w[j + 1]));
            c1 = add2(c1, t);
            c2 = fma2(t, t, c2);
            m[j] = t.x;
            m[j + 1] = t.y;
          }
          put_operand16(T3 + coff + h * 16, m);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&cat_full[2]);
        CQA_STAMP(tl_ph, 12);
        // ---- LayerNorm(384) statistics of the concat row; then, while the MMA warp issues the next tile's S behind the
        //      concat MMAs, the first half of the epilogue: relu(rstd*(acc - mean*cs) + cb) goes back in place over the
        //      accumulator together with the LayerNorm(128) statistics; the epilogue warps normalise and store.
        xq[0][qd][row] = c1.x + c1.y;
        xq[1][qd][row] = c2.x + c2.y;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float mu_c = ((xq[0][0][row] + xq[0][1][row]) + (xq[0][2][row] + xq[0][3][row])) * (1.f / 384.f);
        const float r_c = 1.f / sqrtf(fmaxf(((xq[1][0][row] + xq[1][1][row]) + (xq[1][2][row] + xq[1][3][row])) * (1.f / 384.f) - mu_c * mu_c, 0.f) + 1e-5f);
        const float nm_c = -r_c * mu_c;
        mbar_wait(&mma_done[4], done_cnt & 1u, 39);  // D4: concat accumulator complete
        ++done_cnt;  // one phase of every mma_done barrier per tile
        tc_fence_after();
        CQA_STAMP(tl_ph, 13);
        {
          float2 o1 = make_float2(0.f, 0.f), o2 = make_float2(0.f, 0.f);
          const float2 r_c2 = dup2(r_c), nm_c2 = dup2(nm_c);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float o[16];
            tmem_ld16(T2 + coff + h * 16, o);
            tmem_wait_ld();
            const float4* cs4 = reinterpret_cast<const float4*>(&pvec[2][qd * 32 + h * 16]);
            const float4* cb4 = reinterpret_cast<const float4*>(&pvec[3][qd * 32 + h * 16]);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 cs = cs4[j4], cb = cb4[j4];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = j4 * 4 + e * 2;
                const float2 t = max2(fma2(r_c2, make_float2(o[j], o[j + 1]), fma2(nm_c2, e == 0 ? lo2(cs) : hi2(cs), e == 0 ? lo2(cb) : hi2(cb))), 0.f);
                o1 = add2(o1, t);
                o2 = fma2(t, t, o2);
                o[j] = t.x;
                o[j + 1] = t.y;
              }
            }
            tmem_st16(T2 + coff + h * 16, o);
          }
          xq[2][qd][row] = o1.x + o1.y;
          xq[3][qd][row] = o2.x + o2.y;
          tmem_wait_st();
          tc_fence_before();
          asm volatile("bar.sync 1, 512;" ::: "memory");  // (also: everyone has read xq[0..1] before the next tile rewrites them)
          if (qd == 0) {
            const float mean = ((xq[2][0][row] + xq[2][1][row]) + (xq[2][2][row] + xq[2][3][row])) * (1.f / 128.f);
            const float rstd = 1.f / sqrtf(fmaxf(((xq[3][0][row] + xq[3][1][row]) + (xq[3][2][row] + xq[3][3][row])) * (1.f / 128.f) - mean * mean, 0.f) + 1e-5f);
            ep_stat[row] = make_float2(rstd, -mean * rstd);
            // (the ctx region may already be receiving the next frame here, hence any_v was taken at frame start)
            if (live) p.statement_mask[srow] = (ma * any_v != 0.f) ? 1.f : 0.f;
          }
          mbar_arrive(&ep_ready);  // my activated cells are in tensor memory (and, for quarter 0, the row statistics in place)
        }
        CQA_STAMP(tl_ph, 14);
        if constexpr (TIMELINE) ++tl_tile;
#undef orow
      }
    }
#undef CQA_SEG_PUT
  } else if (warp < 4) {
    reg_release_issue();
    if (warp == 0) {
      // ================================================================= MMA issue.  The whole warp runs the control flow
      // (waits included) and ONE elected lane issues: only in converged code can the compiler keep descriptors in uniform
      // registers; under a `lane == 0` branch every tcgen05 instruction sits in a lane-by-lane waterfall loop.  The issuer
      // never waits for a completion.  The per-tile program is fully unrolled straight-line code: every shared-memory
      // descriptor is a base descriptor (built once, below) plus a compile-time constant, every weight-ring slot index is a
      // constant (20 chunks per tile, 4 slots) — measured: with descriptors rebuilt from run-time indices the issuing warp
      // needed 300-400 cycles of dependent uniform-datapath instructions per K=16 step, twice the tensor time of the step.
      [[maybe_unused]] int tl_tile = (lane == 0) ? 0 : 64;
      [[maybe_unused]] long long w_wait = 0, o_wait = 0;  // timeline build: cycles spent waiting for weights / for operands
#define CQA_TIMED(acc, stmt)                        \
  do {                                              \
    if constexpr (TIMELINE) {                       \
      const long long t_ = clock64();               \
      stmt;                                         \
      acc += clock64() - t_;                        \
    } else {                                        \
      stmt;                                         \
    }                                               \
  } while (0)
      const uint32_t id_ssub = make_idesc_f16(p.ls_pad), id_svid = make_idesc_f16(p.r_pad);
      const uint32_t id_u = make_idesc_f16(128, 0, 1), id_g = make_idesc_f16(128);
      // base descriptors; desc(addr + off) = desc(addr) + (off >> 4) while the address stays inside shared memory
      const uint64_t dA = make_smem_desc(smem_u32(sAQ), CqaCfg::LBO_A, 128);    // answer image / P_vid / a*U stages (K-major, 128 rows)
      const uint64_t dP = make_smem_desc(smem_u32(sB), CqaCfg::LBO_A, 128);     // P_sub
      const uint64_t dW = make_smem_desc(smem_u32(sB), 2048, 128);              // weight ring (K-major, 128 rows)
      const uint64_t dAW = make_smem_desc(smem_u32(sAQ), 2048, 128);            // the four extra weight slots in the AQ region
      const uint64_t dCS = make_smem_desc(smem_u32(sCS), CqaCfg::LBO_CS, 128);  // sub-title tile, K-major (S)
      const uint64_t dCV = make_smem_desc(smem_u32(sCV), CqaCfg::LBO_CV, 128);  // region tile, K-major (S)
      const uint64_t dCSt = make_smem_desc(smem_u32(sCS), 128, CqaCfg::LBO_CS); // sub-title tile, MN-major (U)
      const uint64_t dCVt = make_smem_desc(smem_u32(sCV), 128, CqaCfg::LBO_CV); // region tile, MN-major (U)
#define DOFF(d, bytes) ((d) + static_cast<uint64_t>((bytes) >> 4))
      // S = a . ctx^T for both streams of tile k (issued right behind the previous tile's concat MMAs)
      auto issue_s = [&](int k) {
        const int a_i = k % p.n_pt, f = k / p.n_pt;
        if (a_i == 0) mbar_wait(&ctx_full, f & 1u, 40);
        mbar_wait(&aq_full, k & 1u, 41);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) {
            const uint64_t a_hi = DOFF(dA, ks * 2 * CqaCfg::LBO_A), a_lo = DOFF(dA, CqaCfg::AQ_PART + ks * 2 * CqaCfg::LBO_A);
            umma_split3(T0, a_hi, a_lo, DOFF(dCS, ks * 2 * CqaCfg::LBO_CS), DOFF(dCS, CqaCfg::CS_PART + ks * 2 * CqaCfg::LBO_CS), id_ssub, ks > 0 ? 1u : 0u);
            umma_split3(T1, a_hi, a_lo, DOFF(dCV, ks * 2 * CqaCfg::LBO_CV), DOFF(dCV, CqaCfg::CV_PART + ks * 2 * CqaCfg::LBO_CV), id_svid, ks > 0 ? 1u : 0u);
          }
          umma_commit(&mma_done[0]);  // D1
        }
      };
      if (my_tiles > 0) issue_s(0);
      for (int k = 0; k < my_tiles; ++k) {
        const int a_i = k % p.n_pt;
        const uint32_t wuse = static_cast<uint32_t>(k) * 4u;  // uses of every region-B ring slot before this tile
        // local chunk j: region-B slot j % 4 (its (wuse + w_use(j))-th use) or, for j = 12..15, AQ slot j - 12 (once per tile)
#define W_WAIT(j)                                                                                                       \
  CQA_TIMED(w_wait, (CqaCfg::w_in_aq(j) ? mbar_wait(&wq_full[((j) - 12) & 3], k & 1u, 46)                                \
                                        : mbar_wait(&w_full[(j) % CqaCfg::W_SLOTS], (wuse + CqaCfg::w_use(j)) & 1u, 46)))
#define W_DESC(j, off) \
  (CqaCfg::w_in_aq(j) ? DOFF(dAW, (((j) - 12) & 3) * CqaCfg::W_SLOT + (off)) : DOFF(dW, ((j) % CqaCfg::W_SLOTS) * CqaCfg::W_SLOT + (off)))
        CQA_STAMP(1, 0);
        // ---- U = P' . ctx  (ctx image read MN-major: LBO = 128 between 8-row k groups, SBO = image LBO between n groups).
        //      (Issuing the K=16 steps one by one while the softmax is still producing later chunks was measured and lost:
        //      the per-chunk fences and arrivals cost the compute warps more than the overlap returned.)
        mbar_wait(&op_ready, k & 1u, 42);  // E2
        tc_fence_after();
        CQA_STAMP(1, 1);
        if (elect_one()) {  // U_vid first (T3: the previous tile's product operand there is dead, its MMAs were issued earlier)
#pragma unroll
          for (uint32_t ks = 0; ks < CqaCfg::R_MAX / 16; ++ks)
            if (static_cast<int>(ks * 16) < p.r_pad)
              umma_split3(T3, DOFF(dA, ks * 2 * CqaCfg::LBO_A), DOFF(dA, CqaCfg::PV_PART + ks * 2 * CqaCfg::LBO_A),
                          DOFF(dCVt, ks * 256), DOFF(dCVt, CqaCfg::CV_PART + ks * 256), id_u, ks > 0 ? 1u : 0u);
          umma_commit(&mma_done[1]);  // D2v
        }
        if (k > 0) {
          CQA_TIMED(o_wait, mbar_wait(&t2_free, (k - 1) & 1u, 43));  // the epilogue warps have read the previous concat accumulator (T2)
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (uint32_t ks = 0; ks < CqaCfg::LS_MAX / 16; ++ks)
            if (static_cast<int>(ks * 16) < p.ls_pad)
              umma_split3(T2, DOFF(dP, ks * 2 * CqaCfg::LBO_A), DOFF(dP, CqaCfg::PS_PART + ks * 2 * CqaCfg::LBO_A),
                          DOFF(dCSt, ks * 256), DOFF(dCSt, CqaCfg::CS_PART + ks * 256), id_u, ks > 0 ? 1u : 0u);
          umma_commit(&mma_done[3]);  // D2s
          umma_commit(&b_free);       // P_sub (region B) may now become the weight ring
          if (a_i == p.n_pt - 1) umma_commit(&ctx_free);  // last use of this frame's ctx tiles
        }
        CQA_STAMP(1, 2);
        // ---- c2q: per half h of the K range, the a*U stages of both streams (shared memory, K=64, two weight chunks used by
        //      sub then vid) and then that half of the U operands (tensor memory, two weight chunks, again sub then vid);
        //      accumulators T0 / T1 start from the stored a.Wa term.  Local chunks: h0: 0,1 (a*U) 2,3 (U); h1: 4,5 / 6,7.
#pragma unroll
        for (uint32_t h = 0; h < 2; ++h) {
#pragma unroll
          for (uint32_t gi = 0; gi < 2; ++gi) {
            const uint32_t g = 1u - gi;  // vid first, like the producers
            CQA_TIMED(o_wait, mbar_wait(&a_full[g], h, 45));  // two phases per tile: h = 0, 1
            if (gi == 0) {
              W_WAIT(4 * h);
              W_WAIT(4 * h + 1);
            }
            tc_fence_after();
            if (h == 0) CQA_STAMP(3, gi * 2);
            if (elect_one()) {
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks) {
                const uint32_t j = 4 * h + (ks >> 1);
                umma_split3(g == 0 ? T0 : T1, DOFF(dA, g * CqaCfg::AST_BYTES + ks * 2 * CqaCfg::LBO_A),
                            DOFF(dA, g * CqaCfg::AST_BYTES + CqaCfg::AST_PART + ks * 2 * CqaCfg::LBO_A),
                            W_DESC(j, (ks & 1) * 2 * 2048), W_DESC(j, CqaCfg::W_SLOT / 2 + (ks & 1) * 2 * 2048), id_g, 1u);
              }
              umma_commit(&a_empty[g]);
              if (gi == 1) {
                umma_commit(&w_empty[(4 * h) % CqaCfg::W_SLOTS]);
                umma_commit(&w_empty[(4 * h + 1) % CqaCfg::W_SLOTS]);
              }
            }
            if (h == 0) CQA_STAMP(3, gi * 2 + 1);
          }
          CQA_TIMED(o_wait, mbar_wait(&useg_full[h], k & 1u, 47));
          tc_fence_after();
          if (h == 0) CQA_STAMP(3, 4);
          // packed K-step (h, q) of the U segment = tensor-memory columns [32q + 16h, +16) of T2 / T3
#pragma unroll
          for (uint32_t c = 0; c < 2; ++c) {
            const uint32_t j = 4 * h + 2 + c;
            W_WAIT(j);  // (bulk-copied shared memory: the mbarrier orders it, no tcgen05 fence — a fence here stalls the issue)
            if (elect_one()) {
#pragma unroll
              for (uint32_t g = 0; g < 2; ++g) {
#pragma unroll
                for (uint32_t s2 = 0; s2 < 2; ++s2) {
                  const uint32_t acol = (g == 0 ? T2 : T3) + 32u * (2u * c + s2) + 16u * h;
                  const uint64_t b_hi = W_DESC(j, s2 * 2 * 2048), b_lo = W_DESC(j, CqaCfg::W_SLOT / 2 + s2 * 2 * 2048);
                  umma_f16_ts(g == 0 ? T0 : T1, acol + 8, b_hi, id_g, 1u);
                  umma_f16_ts(g == 0 ? T0 : T1, acol, b_lo, id_g, 1u);
                  umma_f16_ts(g == 0 ? T0 : T1, acol, b_hi, id_g, 1u);
                }
              }
              umma_commit(&w_empty[j % CqaCfg::W_SLOTS]);
            }
          }
        }
        if (elect_one()) {
          umma_commit(&mma_done[2]);  // D3: both c2q accumulators
          umma_commit(&aq_free);      // both shared-memory stages are drained: the AQ region can take weight chunks 12..15
        }
        CQA_STAMP(1, 3);
        CQA_STAMP(1, 4);
        // ---- concat: ms (T0), mv (T1), ms*mv (T3) against four weight chunks each (local chunks 8..19), accumulator T2
#pragma unroll
        for (uint32_t sgm = 0; sgm < 3; ++sgm) {
          CQA_TIMED(o_wait, mbar_wait(&cat_full[sgm], k & 1u, 48));
          tc_fence_after();
          CQA_STAMP(3, 5 + sgm);
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {
            const uint32_t j = 8 + 4 * sgm + c;
            W_WAIT(j);  // (bulk-copied shared memory: the mbarrier orders it, no tcgen05 fence — a fence here stalls the issue)
            if (elect_one()) {
#pragma unroll
              for (uint32_t s2 = 0; s2 < 2; ++s2) {
                const uint32_t acol = (sgm == 0 ? T0 : (sgm == 1 ? T1 : T3)) + 32u * c + 16u * s2;
                const uint64_t b_hi = W_DESC(j, s2 * 2 * 2048), b_lo = W_DESC(j, CqaCfg::W_SLOT / 2 + s2 * 2 * 2048);
                umma_f16_ts(T2, acol + 8, b_hi, id_g, (sgm == 0 && c == 0 && s2 == 0) ? 0u : 1u);
                umma_f16_ts(T2, acol, b_lo, id_g, 1u);
                umma_f16_ts(T2, acol, b_hi, id_g, 1u);
              }
              if (!CqaCfg::w_in_aq(j)) umma_commit(&w_empty[j % CqaCfg::W_SLOTS]);
              if (j == 15) umma_commit(&aq_w_free);  // the AQ region's weight slots are drained: the next answer image may land
            }
          }
        }
        if (elect_one()) umma_commit(&mma_done[4]);  // D4
        CQA_STAMP(1, 5);
        if (k + 1 < my_tiles) issue_s(k + 1);
        CQA_STAMP(1, 6);
        if constexpr (TIMELINE) {
          if (p.timeline != nullptr && blockIdx.x == 0 && tl_tile < 64) {
            p.timeline[(1 * 64 + tl_tile) * 16 + 14] = w_wait;
            p.timeline[(1 * 64 + tl_tile) * 16 + 15] = o_wait;
          }
          w_wait = 0;
          o_wait = 0;
          ++tl_tile;
        }
#undef W_WAIT
#undef W_DESC
      }
#undef DOFF
    } else if (warp == 1 && lane == 0) {
      // ================================================================= weight ring (one thread)
      auto w_src = [&](int q) -> const uint8_t* {
        return q < 8 ? p.c2q_w + static_cast<size_t>(q) * CqaCfg::W_SLOT : p.cat_w + static_cast<size_t>(q - 8) * CqaCfg::W_SLOT;
      };
      for (int k = 0; k < my_tiles; ++k) {
        mbar_wait_relaxed(&b_free, k & 1u, 52, 128);  // U has completed: region B turns from P_sub into the weight ring
        for (int q = 0; q < CqaCfg::CHUNKS; ++q) {
          if (CqaCfg::w_in_aq(q)) {
            if (q == 12) mbar_wait(&aq_free, k & 1u, 54);  // D3: the a*U stages in the AQ region are drained
            const int sl = q - 12;
            mbar_arrive_expect_tx(&wq_full[sl], CqaCfg::W_SLOT);
            bulk_g2s(sAQ + sl * CqaCfg::W_SLOT, w_src(q), CqaCfg::W_SLOT, &wq_full[sl]);
          } else {
            const uint32_t slot = q % CqaCfg::W_SLOTS, use = static_cast<uint32_t>(k) * 4u + CqaCfg::w_use(q);
            mbar_wait(&w_empty[slot], (use & 1u) ^ 1u, 53);
            mbar_arrive_expect_tx(&w_full[slot], CqaCfg::W_SLOT);
            bulk_g2s(sB + slot * CqaCfg::W_SLOT, w_src(q), CqaCfg::W_SLOT, &w_full[slot]);
          }
        }
      }
    } else if (warp == 2 && lane == 0) {
      // ================================================================= ctx blobs and answer images (one thread)
      int k = 0;
      uint32_t frame_cnt = 0;
      for (int frame = blockIdx.x; frame < n_frames; frame += gridDim.x, ++frame_cnt) {
        const int n_i = frame / p.t;
        if (frame_cnt > 0) mbar_wait_relaxed(&ctx_free, (frame_cnt - 1u) & 1u, 50);  // the previous frame's last U has completed
        mbar_arrive_expect_tx(&ctx_full, CqaCfg::CTX_BYTES);
        bulk_g2s(sCTX, p.ctx_blob + static_cast<size_t>(frame) * CqaCfg::CTX_BYTES, CqaCfg::CTX_BYTES, &ctx_full);
        for (int a_i = 0; a_i < p.n_pt; ++a_i, ++k) {
          const size_t nj = static_cast<size_t>(n_i) * p.n_pt + a_i;
          if (k > 0) mbar_wait(&aq_w_free, (k - 1) & 1u, 51);  // the previous tile's a*U stages and AQ weight slots are drained
          mbar_arrive_expect_tx(&aq_full, CqaCfg::AQ_BYTES);
          bulk_g2s(sAQ, p.a_blob + nj * CqaCfg::ABLOB_BYTES, CqaCfg::AQ_BYTES, &aq_full);
        }
      }
    }
  } else {
    reg_release_epilogue();
    // =================================================================== epilogue warps: thread = row, all 128 columns.
    //   The compute warps have left relu(...) in place over the concat accumulator and the LayerNorm(128) statistics of
    //   the row in ep_stat; here: normalise, scale, store, and release the accumulator to the next tile's U as soon as the
    //   last cells have been read.
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = ((warp & 3) * 32u) << 16;
    [[maybe_unused]] int tl_tile = (tid == 640) ? 0 : 64;
    int k = 0;
    for (int frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
      const int n_i = frame / p.t, t_i = frame - n_i * p.t;
      for (int a_i = 0; a_i < p.n_pt; ++a_i, ++k) {
        const int prow = a_i * 128 + row;
        const bool live = prow < 5 * p.lq;
        const int ans = live ? prow / p.lq : 0;
        const uint32_t srow = ((static_cast<uint32_t>(n_i) * 5u + ans) * p.t + t_i) * p.lq + (prow - ans * p.lq);
        // row-major: the row is 512 contiguous bytes (a warp touches 32 lines per store);
        // blocked: float4 c4 of row r sits at (tile*32 + c4)*S + r, so the 32 rows of a warp write 512-byte runs
        float4* dst = reinterpret_cast<float4*>(p.statement) + static_cast<size_t>(srow) * 32;
        size_t dstep = 1;
        if (p.stmt_stride != 0) {
          const uint32_t trow = srow / p.lq;  // (example, answer, frame) tile index
          dst = reinterpret_cast<float4*>(p.statement) + static_cast<size_t>(trow) * 32 * p.stmt_stride + (srow - trow * p.lq);
          dstep = p.stmt_stride;
        }
        mbar_wait_relaxed(&ep_ready, k & 1u, 60);  // (most of a tile away: sleep between polls)
        tc_fence_after();
        CQA_STAMP(5, 0);
        const float2 st = ep_stat[row];
        const float2 rstd2 = dup2(st.x), nmr2 = dup2(st.y);
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float o[16];
          tmem_ld16(T2 + lane_off + c * 16, o);
          tmem_wait_ld();
          if (c == 7) {
            tc_fence_before();
            mbar_arrive(&t2_free);  // the accumulator may be overwritten by the next tile's U
            CQA_STAMP(5, 2);
          }
          const float4* g4p = reinterpret_cast<const float4*>(&pvec[4][c * 16]);
          const float4* b4p = reinterpret_cast<const float4*>(&pvec[5][c * 16]);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 g = g4p[j4], b = b4p[j4];
            const float2 y0 = fma2(fma2(make_float2(o[j4 * 4], o[j4 * 4 + 1]), rstd2, nmr2), lo2(g), lo2(b));
            const float2 y1 = fma2(fma2(make_float2(o[j4 * 4 + 2], o[j4 * 4 + 3]), rstd2, nmr2), hi2(g), hi2(b));
            if (live) dst[(c * 4 + j4) * dstep] = make_float4(y0.x, y0.y, y1.x, y1.y);
          }
        }
        CQA_STAMP(5, 3);
        if constexpr (TIMELINE) ++tl_tile;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace stg
