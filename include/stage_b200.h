/* stage_b200.h — C ABI of libstage_b200.so: the B200 (sm_100a) implementation of the STAGE forward hot path
 * of jayleicn/TVQAplus.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; all tensors are fp32, row-major, contiguous,
 *    16-byte aligned; masks are fp32 {0,1}
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, never synchronises,
 *    never allocates or frees device memory and keeps no pointer after it returns
 *  - return value 0 = success; non-zero = STAGE_E_* code, text via stage_last_error() (thread-local)
 *  - inputs are const, outputs are fully overwritten
 *
 * Each entry point names the reference interface (file:line in jayleicn/TVQAplus) it stands in for.
 */
#ifndef STAGE_B200_H_
#define STAGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAGE_ABI_VERSION 4
#define STAGE_HSZ 128          /* opt.hsz the kernels are specialised for (config.py:103) */
#define STAGE_MAX_SEQ 128      /* longest token/region axis one encoder tile holds */
#define STAGE_MAX_CONV 4       /* n_conv per encoder block (config.py:108,114 default 2) */
#define STAGE_MAX_KSIZE 15     /* depthwise kernel size (config.py:109,115 default 7 / 5) */

enum {
  STAGE_OK = 0,
  STAGE_E_BADARG = 1,      /* null / misaligned pointer, non-positive size */
  STAGE_E_UNSUPPORTED = 2, /* shape outside what the kernels are specialised for */
  STAGE_E_LAUNCH = 3,      /* cudaGetLastError() after launch */
  STAGE_E_DEVICE = 4       /* a kernel raised the device-side error flag (bounded wait timed out) */
};

int stage_abi_version(void);
const char* stage_last_error(void);
/* Reads and clears the device-side error flag (synchronises the stream). 0 = clean. */
int stage_check_device_error(void* stream);
/* Asynchronous form for callers that must not synchronise: enqueues a copy of the error flag into `host_flag` (one
 * uint32 of PINNED host memory owned by the caller) behind everything already on `stream`.  The caller reads it once an
 * event recorded after this call has completed; non-zero = a bounded wait timed out in an earlier kernel and its outputs
 * are garbage.  The flag is not cleared. */
int stage_post_device_error(unsigned int* host_flag, void* stream);

/* ---- weight packing ------------------------------------------------------------------------------------------
 * nn.Linear / 1x1 Conv1d weights W[N,K] (row-major fp32, as stored in the reference state_dict, e.g.
 * model/stage.py:88,110,117,136 and model/cnn.py:27) are split into fp16 (hi, lo) pairs and laid out as the
 * shared-memory operand image the tcgen05 kernels bulk-copy.  `kc` is the K-chunk (32 or 64) the consuming kernel
 * streams.  Re-run after every optimizer step (key the cache on the parameter version). */
size_t stage_packed_weight_bytes(int n, int k, int kc);
int stage_pack_weight(const float* w, int n, int k, int kc, void* packed, void* stream);

/* LayerNorm folded into the Linear that follows it (LN(x) W^T + b = rstd * (x (W*gamma)^T - mean * cs) + cb):
 * packs W*gamma like stage_pack_weight and writes cs[n] = sum_k W[n,k] gamma[k], cb[n] = sum_k W[n,k] beta[k] + bias[n]. */
int stage_pack_weight_ln(const float* w, int n, int k, int kc, const float* ln_g, const float* ln_b, const float* bias,
                         void* packed, float* cs, float* cb, void* stream);

/* ---- row-wise Linear with fused prologue / epilogue -------------------------------------------------------------
 * y = LN_out( act( LN_in( l2norm?(x) ) W^T + b ) ), each piece optional.  Stands in for the nn.Sequential stacks
 * bert_word_encoding_fc / vid_fc / input_embedding (model/stage.py:85-91,98-104,115-120) and F.normalize at
 * model/stage.py:256, and for LinearWrapper (model/stage.py:15-32).  n <= 304, k % 4 == 0.
 * The raw rows are streamed from HBM once; LN_in / l2norm are applied in the epilogue from row statistics gathered on
 * the way (so a weight used with LN_in must come from stage_pack_weight_ln, together with its cs / cb vectors). */
typedef struct {
  int64_t m;             /* rows */
  int k, n;              /* in / out features */
  int kc;                /* K-chunk the weight was packed with */
  int l2norm_in;         /* 1: x <- x / max(||x||_2, 1e-12) first */
  int relu;              /* 1: ReLU after bias */
  const float* x;        /* [m,k] */
  const void* w_packed;  /* stage_pack_weight(W[n,k]) or, with LN_in, stage_pack_weight_ln */
  const float* bias;     /* [n]; ignored when ln_in_cs is given (cb contains it) */
  const float* ln_in_cs; /* [n] from stage_pack_weight_ln, or NULL = no LayerNorm in front */
  const float* ln_in_cb;
  const float* ln_out_g; /* [n] or NULL */
  const float* ln_out_b;
  float* y;              /* [m,n] */
  long long* timeline;   /* optional debug: [32][8] SM clock stamps of CTA 0 (tools/linear_timeline.py), else NULL */
  int64_t ldy;           /* row stride of y in floats; 0 = n (y may be a column slice of a wider tensor) */
  const float* x_rowmax; /* optional [m] = max |x| of every row (stage_row_absmax): each row is scaled by an exact power of two
                            into fp16's normal range in front of the split and its result scaled back - for gradient tensors
                            (dX = dY W in the backward), whose rows sit far below 6e-5 and orders of magnitude apart.
                            Not with LN_in / l2norm_in. */
} StageLinearArgs;
int stage_linear_fwd(const StageLinearArgs* a, void* stream);

/* ---- EncoderBlock, conv part (model/encoder.py:9-44,52, model/cnn.py:6-47, model/position_encoding.py:33-43) -------
 * x + pe -> n_conv x [LN -> depthwise conv(k, zero pad k/2) -> pointwise 128x128 -> ReLU -> + residual] -> LN.
 * With self-attention (n_heads > 0) the caller passes STAGE_ENC_NO_FINAL_LN here and finishes the block with
 * stage_mha_fwd (LN -> MHA -> + residual -> LN).
 * out_mode 2 (STAGE_ENC_OUT_CQA_SUB): the closing LayerNorm output of sequence i goes straight into the sub-title part of
 * context blob i of a stage_cqa_fused_fwd workspace (split-fp16 operand image, inverse L2 norms, masks; y = the blob
 * base stage_cqa_ctx_blobs(workspace, ..)), so that stage_cqa_fused_fwd(sub_prepared = 1) needs neither the fp32 sub-title
 * embeddings nor a conversion pass over them.  Needs a mask and 64 < L <= 112 (one sequence per tile).
 * out_mode 0: y[n_seq,L,128].  out_mode 1: masked max-pool over L (mask_logits + max, model/stage.py:503):
 * y[n_seq,128].  L <= 128. */
#define STAGE_ENC_NO_PE 1        /* skip the position-encoding add */
#define STAGE_ENC_NO_RESIDUAL 2  /* x <- act(conv(x)) instead of x <- x + act(conv(LN x)) */
#define STAGE_ENC_NO_FINAL_LN 4  /* skip the closing LayerNorm */
#define STAGE_ENC_NO_RELU 8      /* pointwise conv without ReLU (DepthwiseSeparableConv(relu=False), model/cnn.py:45-46) */
#define STAGE_ENC_IN_BLOCKED 16  /* x (and pe) come in the blocked tile layout below instead of row-major */
/* Blocked tile layout (internal fast path between stage_cqa_fused_fwd and the cls encoder): sequence i of length L is a
 * block of 32 x S float4, element (row, 4-column group c4) at float4 index (i*32 + c4)*S + row, S = stage_blocked_stride(L)
 * (>= L, == 1 mod 8 so that both a row-per-lane writer and reader are bank-conflict free).  With STAGE_ENC_IN_BLOCKED the
 * position table is passed in the same form: float4 index c4*L + pos. */
int stage_blocked_stride(int L);
/* a conv whose ln_g/ln_b are NULL takes its input without LayerNorm (plain DepthwiseSeparableConv, model/cnn.py:37-47;
 * with them and NO_PE|NO_RESIDUAL|NO_FINAL_LN it is ConvLinear, model/stage.py:35-52) */
typedef struct {
  int64_t n_seq;
  int L;
  int n_conv, ksize;
  int out_mode;
  int flags;           /* STAGE_ENC_* bits; 0 = the full (attention-free) EncoderBlock */
  const float* x;      /* [n_seq,L,128] */
  const float* mask;   /* [n_seq,L] (needed for out_mode 1) */
  const float* pe;     /* [>=L,128] position-encoding buffer (NULL with STAGE_ENC_NO_PE) */
  const float* ln_g[STAGE_MAX_CONV];
  const float* ln_b[STAGE_MAX_CONV];
  const float* dw_w[STAGE_MAX_CONV];   /* depthwise weight [128,1,k] */
  const float* dw_b[STAGE_MAX_CONV];
  const void* pw_packed[STAGE_MAX_CONV]; /* stage_pack_weight(pointwise [128,128], kc=64) */
  const float* pw_b[STAGE_MAX_CONV];
  const float* final_ln_g;
  const float* final_ln_b;
  float* y;
  long long* timeline; /* optional debug: [32][16] SM clock stamps of CTA 0 (tools/encoder_timeline.py), else NULL */
} StageEncoderArgs;
int stage_encoder_block_fwd(const StageEncoderArgs* a, void* stream);

/* ---- multi-head self-attention (model/self_attention.py:19-71) with the EncoderBlock glue around it
 *      (model/encoder.py:46-52) -------------------------------------------------------------------------------------
 * y = LN_fin?( (x +)? Wo . concat_h softmax(Q_h K_h^T / sqrt(d_k)) V_h ),  Q,K,V = LN_in?(x) W^T + b.
 * Query-row mask exactly like the reference (self_attention.py:43,66-67): rows with mask == 0 attend uniformly,
 * valid rows also see padded keys.  n_heads in {1, 2, 4}; L <= 128.  out_mode as for the encoder. */
typedef struct {
  int64_t n_seq;
  int L;
  int n_heads;
  int residual;            /* 1: y = x + attention(LN_in(x)) */
  int out_mode;
  const float* x;          /* [n_seq,L,128] */
  const float* mask;       /* [n_seq,L] or NULL (no masking) */
  const float* ln_in_g;    /* attn_layer_norm or NULL */
  const float* ln_in_b;
  const void* w_packed[4]; /* stage_pack_weight(linears.0..3 weight [128,128], kc=64) */
  const float* b[4];
  const float* final_ln_g; /* closing LayerNorm or NULL */
  const float* final_ln_b;
  float* y;
} StageMhaArgs;
int stage_mha_fwd(const StageMhaArgs* a, void* stream);

/* ---- fused QA<->context attention + fusion MLPs ------------------------------------------------------------------
 * StructuredAttention (model/context_query_attention.py:35-101) for both streams, qa_ctx_attention +
 * c2q_down_projection (model/stage.py:365-387,133-138) and concat_fc (model/stage.py:107-113,276-279):
 * one CTA per (example, frame), looping over the 5 answers; S, P, U, both mixed tensors stay in TMEM / smem.
 * Lq <= 128, Ls <= 112, R <= 48.
 *
 * Weights are prepared once per parameter version with stage_cqa_prepare_weights (LayerNorm(384) folded into the
 * Linear that follows it, split-fp16 images); the forward needs a caller-owned workspace for the operand images. */
size_t stage_cqa_weight_blob_bytes(void);
int stage_cqa_prepare_weights(const float* c2q_ln_g, const float* c2q_ln_b, const float* c2q_w /*[128,384]*/,
                              const float* c2q_b, const float* cat_ln_g, const float* cat_ln_b,
                              const float* cat_w /*[128,384]*/, const float* cat_b, void* wblob, void* stream);
size_t stage_cqa_workspace_bytes(int n, int t, int lq);
typedef struct {
  int n, t, lq, ls, r;
  float scale;                 /* opt.scale (config.py:67) */
  const float* a_embed;        /* [n,5,lq,128] */
  const float* a_mask;         /* [n,5,lq] */
  const float* sub_embed;      /* [n,t,ls,128] */
  const float* sub_mask;       /* [n,t,ls] */
  const float* vid_embed;      /* [n,t,r,128] */
  const float* vid_mask;       /* [n,t,r] */
  const void* wblob;           /* stage_cqa_prepare_weights output */
  const float* cat_out_ln_g;   /* [128] concat_fc.4 */
  const float* cat_out_ln_b;
  void* workspace;             /* stage_cqa_workspace_bytes(n, t, lq) bytes, 128-byte aligned */
  int statement_blocked;       /* 0: statement is [n,5,t,lq,128]; 1: blocked tile layout, n*5*t blocks (see stage_blocked_stride) */
  float* statement;            /* [n,5,t,lq,128] or n*5*t*32*S float4 */
  float* statement_mask;       /* [n,5,t,lq]  (the vid mixed_mask, model/stage.py:280-282) */
  float* vid_raw_s;            /* optional [n,5,t,lq,r] (other_outputs["vid_raw_s"], model/stage.py:273) */
  float* sub_raw_s;            /* optional [n,5,t,lq,ls] */
  float* vid_norm_s;           /* optional [n,5,t,lq,r]  masked softmax weights */
  float* sub_norm_s;           /* optional [n,5,t,lq,ls] */
  float* att_vid;              /* optional [n,5,t,lq,128] U = P.ctx (StructuredAttention's A) */
  float* att_sub;              /* optional */
  float* mixed_vid;            /* optional [n,5,t,lq,128] c2q output */
  float* mixed_sub;            /* optional */
  long long* timeline;         /* optional debug: [8][64][16] SM clock stamps of CTA 0 (see tools/cqa_timeline.py) */
  int sub_prepared;            /* 1: the sub-title parts of the context blobs were written by stage_encoder_block_fwd(out_mode 2);
                                  sub_embed may then be NULL */
} StageCqaArgs;
int stage_cqa_fused_fwd(const StageCqaArgs* a, void* stream);
/* first context blob (one per (example, frame), stage_cqa_ctx_blob_bytes() apart) inside a workspace of
 * stage_cqa_workspace_bytes(n, t, lq) bytes: the y argument of stage_encoder_block_fwd(out_mode 2) */
void* stage_cqa_ctx_blobs(void* workspace, int n, int t, int lq);
size_t stage_cqa_ctx_blob_bytes(void);

/* ---- span + answer heads -------------------------------------------------------------------------------------------
 * residual_temporal_predictor, mask, softmax over frames, best (st<=ed) span, span-local / global masked max,
 * classifier (model/stage.py:469-537, 439-467; model/model_utils.py:78-121), eval branch, t_iter layers >= 1
 * skipped (they never reach the outputs, model/stage.py:515-518). */
typedef struct {
  int n, t;
  int add_local, extra_span_length;
  const float* max_statement;      /* [n*5,t,128] */
  const float* statement_mask;     /* [n,5,t,lq]: max_statement_mask = (sum over lq != 0) */
  int lq;
  const float* ts_mask;            /* [n,t] */
  const float* proj_ln_g; const float* proj_ln_b; const float* proj_w; const float* proj_b; /* cls_projection_layers.0 */
  const float* st_ln_g; const float* st_ln_b; const float* st_w; const float* st_b;       /* temporal_scoring_st_layers.0 */
  const float* ed_ln_g; const float* ed_ln_b; const float* ed_w; const float* ed_b;
  const float* cls_ln_g; const float* cls_ln_b; const float* cls_w; const float* cls_b;   /* classifier */
  float* t_scores;     /* [n,5,t,2] masked raw logits */
  float* t_probs;      /* [n,5,t,2] softmax over t */
  int32_t* spans;      /* [n*5,2] predicted (st, ed) inclusive */
  float* span_conf;    /* [n*5] p_st*p_ed of the chosen span */
  float* answer;       /* [n,5] */
  float* workspace;    /* n*5*t*128 floats (encoded max_statement) */
  /* optional (all NULL = off): get_ts_loss (model/stage.py:539-555) fused into the same launch — the cross-entropy terms
   * of the ground-truth answer's row and their sum over the batch, (st + ed) / 2 — and what the train-mode proposal pass
   * below reads */
  const int64_t* target;   /* [n] ground-truth answer index */
  const int64_t* ts_st;    /* [n] ground-truth start frame */
  const int64_t* ts_ed;    /* [n] ground-truth end frame (inclusive) */
  float* ts_terms;         /* [n,2] scratch: -log softmax(scores)[label] for start / end */
  float* ts_loss;          /* [1] temporal loss */
  float* frame_mask;       /* [n*5,t] max_statement_mask */
  float* gfeat;            /* [n*5,128] masked max over all frames of the encoded statement */
} StageHeadArgs;
int stage_span_head_fwd(const StageHeadArgs* a, void* stream);

/* ---- train branch of get_proposals + classifier (model/stage.py:407-438,536), run after stage_span_head_fwd -----------
 * Candidate row 2*i is example i's ground-truth span, row 2*i + 1 its predicted span (the best st <= ed pair of the
 * ground-truth answer's row), valid only if confidence >= ce_prob_thd and temporal IoU with the ground truth >= 0.5
 * (model_utils.py:37-51).  Valid rows hold the five answer scores of that proposal; the caller compacts the valid rows
 * (in order) into the reference's (N_new, 5) output and (N_new,) targets. */
typedef struct {
  int n, t;
  int extra_span_length;
  float ce_prob_thd;
  const float* encoded;     /* [n,5,t,128] workspace of stage_span_head_fwd */
  const float* frame_mask;  /* [n*5,t] */
  const float* gfeat;       /* [n*5,128] */
  const int32_t* spans;     /* [n*5,2] */
  const float* span_conf;   /* [n*5] */
  const int64_t* target; const int64_t* ts_st; const int64_t* ts_ed;  /* [n] each */
  const float* cls_ln_g; const float* cls_ln_b; const float* cls_w; const float* cls_b;   /* classifier (256 -> 1) */
  float* answer_rows;       /* [2n,5] */
  int32_t* row_valid;       /* [2n] */
  int64_t* row_target;      /* [2n] */
} StageProposalArgs;
int stage_train_proposals_fwd(const StageProposalArgs* a, void* stream);

/* ---- best span search on its own (find_max_triples topN=1, model/model_utils.py:92-121) -------------------------- */
int stage_find_max_span(const float* p_st, const float* p_ed, int rows, int t, int32_t* spans, float* conf,
                        void* stream);

/* ---- span of the predicted answer at inference time (find_max_pair, inference.py:13-35) ---------------------------
 * p1 / p2: start / end probabilities, element (row, i) at p[row*row_stride + i*elem_stride] (so the (N,T,2) slice of
 * t_scores for the predicted answers is passed as p1 = base, p2 = base + 1, row_stride = 2T, elem_stride = 2).
 * spans[rows,2] = (st, ed) with st <= ed, val[rows] = p1[st]*p2[ed] in double, exactly the reference's Python-float loop
 * (first maximum wins; (0, 1) / 0.0 when every product is 0). */
int stage_find_max_pair(const float* p1, const float* p2, int rows, int t, int64_t row_stride, int elem_stride,
                        int32_t* spans, double* val, void* stream);

/* ---- attention ranking loss (get_att_loss, model/stage.py:731-746) ---------------------------------------------------
 * scores = vid_raw_s (N,5,T,Lq,R) of stage_cqa_fused_fwd (masked entries are -1e10); pos_off / neg_off [m] = flat element
 * offsets into it of the sampled (positive, negative) region pairs (the host builds them from att_labels exactly as
 * model/stage.py:613-694 does).  loss_type 0 = "lse": sum log1p(exp(alpha*(s_neg - s_pos))); 1 = "hinge": sum
 * max(0, margin + s_neg - s_pos).  partials: caller-owned scratch of (m + 255) / 256 floats; loss[1]. */
int stage_att_loss_fwd(const float* scores, const int64_t* pos_off, const int64_t* neg_off, int64_t m, int loss_type,
                       float alpha, float margin, float* partials, float* loss, void* stream);

/* ---- pad_sequences_2d on the device (collate / prepare_inputs, tvqa_dataset.py:515-553,631-688) -------------------
 * The host uploads only the VALID rows of a ragged feature tensor, concatenated (`rows`, d floats each, d % 4 == 0), plus
 * per-sequence row offsets and lengths; out[n_seq, l_pad, d] gets sequence s's first min(len[s], l_pad) rows and zeros
 * behind them (pads must be zero: the reference's outputs depend on pad values), mask[n_seq, l_pad] = (j < len[s]).
 * The clipping to l_pad is prepare_inputs' `[:, :, :max_*_l]`. */
int stage_pad_rows(const float* rows, const int64_t* off, const int32_t* len, int64_t n_seq, int l_pad, int d, float* out,
                   float* mask, void* stream);

/* ---- mask_logits (model/model_utils.py:14-15): y = x*m + (1-m)*(-1e10), m broadcast over the last `inner` elems ---- */
int stage_mask_logits(const float* x, const float* mask, int64_t outer, int inner, float* y, void* stream);

/* ==== training primitives (train_b200.cu): forward-with-saved-state and backward of every operator of the path ========
 * The reference trains through PyTorch autograd (main.py:53-66: loss.backward(), clip_grad_norm_, Adam.step()).  Here the
 * backward is a sequence of these entries driven by tvqaplus_b200/train.py from inside one autograd.Function, exact fp32
 * (CUDA cores).  Row order of the attended tensors in the training engine: (n, a, i, t) - see train.py.
 * Gradients of parameters are ACCUMULATED (+=, atomics) into caller-zeroed buffers: weights shared by the three input
 * streams collect their three contributions in place. */

/* C[b0,b1] (op)= alpha * A[b0,b1] * B[b0,b1] (+ bias[n]) (ReLU).  A(m,k) at a + m*a_sm + k*a_sk (one of the two strides
 * must be 1), B(k,n) at b + k*b_sk + n*b_sn (one must be 1), C row-major with leading dimension ldc; all strides in
 * elements.  mode 0: C = ..., 1: C += ..., 2: atomicAdd (required when split_k > 1 or when batches share C).
 * Stands in for F.linear / torch.matmul and their autograd formulas (dX = dY W, dW = dY^T X). */
typedef struct {
  const float* a; const float* b; float* c; const float* bias;
  int32_t m, n, k, nb0, nb1;
  int64_t a_sm, a_sk, b_sk, b_sn, ldc;
  int64_t a_b0, a_b1, b_b0, b_b1, c_b0, c_b1;
  int32_t split_k, mode, relu;
  float alpha;
} StageGemmArgs;
int stage_gemm(const StageGemmArgs* a, void* stream);

/* nn.LayerNorm(d), eps 1e-5 (d <= 768); saves mean / rstd per row for the backward */
int stage_ln_fwd(const float* x, const float* g, const float* b, float* y, float* mean, float* rstd, int64_t rows, int d, void* stream);
/* dx = [add +] dLN/dx (dy) [masked by x > 0 when relu_x: x is a ReLU output and the gradient wanted is the ReLU input's];
 * dg += sum dy*xhat, db += sum dy.  dx may be NULL (parameter gradients only); add may alias dx. */
int stage_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* g, const float* add, float* dx,
                 float* dg, float* db, int64_t rows, int d, int relu_x, void* stream);
/* out[r] = max_c |x[r, c]| */
int stage_row_absmax(const float* x, int64_t rows, int d, float* out, void* stream);
/* out[c] += sum_r x[r*ld + c] (bias gradients) */
int stage_colsum(const float* x, int64_t rows, int d, int64_t ld, float* out, void* stream);
int stage_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
/* y = a + b; period > 0: b has `period` elements and repeats (PositionEncoding, model/position_encoding.py:33-43) */
int stage_add(const float* a, const float* b, float* y, int64_t n, int64_t period, void* stream);
/* x *= (*scalar) * mult: chain rule with a device-resident upstream gradient, no host read */
int stage_scale(float* x, int64_t n, const float* scalar, float mult, void* stream);
/* nn.Dropout / F.dropout in train mode (inverted dropout): y[i] = keep(i) ? x[i] / (1 - p) : 0 with
 * keep(i) = 16 bits of hash(seed_dev[0], seed_dev[1], site, i / 4) >= p * 2^16 (seed_dev: two int32 on the device - the run's seed and the
 * forward-call counter; `site` names the dropout module).  Stateless: the backward applies the same call to the gradient.
 * y may alias x.  The distribution is the reference's, the random stream is not torch's. */
int stage_dropout(const float* x, float* y, int64_t n, float p, const int32_t* seed_dev, uint32_t site, void* stream);
/* (groups, I, T, d) -> (groups, T, I, d) */
int stage_permute_it(const float* x, float* y, int64_t groups, int i_dim, int t_dim, int d, void* stream);
/* depthwise Conv1d(groups = d, padding = k/2) of DepthwiseSeparableConv (model/cnn.py:25-28), d = 128, odd k <= 7 */
int stage_dwconv_fwd(const float* x, const float* w, const float* bias, float* y, int64_t n_seq, int l, int d, int k, void* stream);
int stage_dwconv_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t n_seq, int l, int d, int k, void* stream);
/* F.normalize(p=2, dim=-1) of rows, d % 4 == 0 (model/context_query_attention.py:54-55, model/stage.py:256); inv may be NULL */
int stage_l2norm_fwd(const float* x, float* y, float* inv, int64_t rows, int d, void* stream);
int stage_l2norm_bwd(const float* dy, const float* y, const float* inv, float* dx, int64_t rows, int d, int accumulate, void* stream);
/* s[n, rows_per_n, t, lc] raw cosines -> masked in place (raw - 1e10 (1 - cm*qm)); p = softmax(scale*s) * cm*qm
 * (model/context_query_attention.py:57-66,100).  c_mask [n*rows_per_n], q_mask [n, t, lc]. */
int stage_attn_softmax_fwd(float* s, const float* c_mask, const float* q_mask, float* p, int64_t n, int rows_per_n, int t, int lc, float scale, void* stream);
/* dp <- scale * p * (dp - sum_j dp*p) */
int stage_attn_softmax_bwd(float* dp, const float* p, int64_t rows, int lc, float scale, void* stream);
/* out[r] = [a[r / a_div], b[r], a*b], 128 + 128 + 128 (torch.cat of model/stage.py:376 and :276) and its backward */
int stage_cat3_fwd(const float* a, const float* b, float* out, int64_t rows, int a_div, void* stream);
int stage_cat3_bwd(const float* dcat, const float* a, const float* b, float* da, float* db, int64_t a_rows, int a_div, int accumulate_a, void* stream);
/* max over the rows of each sequence of mask_logits(x, mask) (model/stage.py:503), 128 channels, with the argmax */
int stage_maxpool_fwd(const float* x, const float* mask, float* y, int32_t* arg, int64_t n_seq, int l, void* stream);
int stage_maxpool_bwd(const float* dy, const float* mask, const int32_t* arg, float* dx, int64_t n_seq, int l, void* stream);
/* get_ts_loss (model/stage.py:539-555) on the masked scores [n,5,t,2]: terms[n,2] = CE_st, CE_ed of the GT answer's rows,
 * probs[2,n,t] = their softmax over frames (input of the span search), dscores[n,5,t,2] (caller-zeroed) = gradient of
 * 0.5*(CE_st + CE_ed) w.r.t. the UNMASKED scores (already times the frame mask). */
int stage_ts_loss_grad(const float* t_scores, const float* ts_mask, const int64_t* target, const int64_t* st, const int64_t* ed, int n, int t,
                       float* terms, float* probs, float* dscores, void* stream);
/* train branch of get_proposals (model/stage.py:407-438): rows 2n = GT span, 2n+1 = predicted span (valid if confident and
 * IoU >= 0.5); feat[2n,5,256] = [span-local max, global max] of mask_logits(x[n,5,t,128], frame_mask[n,5,t]) + argmax */
int stage_proposal_pool_fwd(const float* x, const float* frame_mask, const int64_t* target, const int64_t* st, const int64_t* ed,
                            const int32_t* spans, const float* conf, int n, int t, int extra, float prob_thd, float* feat, int32_t* arg,
                            int32_t* valid, int64_t* row_target, void* stream);
int stage_proposal_pool_bwd(const float* dfeat, const int32_t* arg, const int32_t* valid, const float* frame_mask, int n, int t, float* dx, void* stream);
/* gradient of stage_att_loss_fwd scattered (atomicAdd) into dscores; upstream = device scalar */
int stage_att_loss_bwd(const float* scores, const int64_t* pos_off, const int64_t* neg_off, int64_t m, int loss_type, float alpha, float margin,
                       const float* upstream, float* dscores, void* stream);
/* cls loss of main.py:57-59 over the VALID proposal rows [n_rows,5]: out[0] = sum CE * (n_examples / n_valid), out[1] =
 * n_valid, out[2] = correct rows; drows[n_rows,5] = its gradient (0 on invalid rows).  No host read, no compaction. */
int stage_answer_ce(const float* rows, const int32_t* valid, const int64_t* row_target, int n_rows, int n_examples, float* out, float* drows, void* stream);
/* out[0] += sum g^2 (caller zeroes out): the global norm of clip_grad_norm_ (main.py:65) */
int stage_grad_sumsq(const float* g, int64_t n, float* out, void* stream);
/* clip (coef = min(1, max_norm / (sqrt(*sumsq)*grad_scale + 1e-6)); max_norm <= 0: off) + torch.optim.Adam with L2 weight
 * decay on flat buffers (main.py:209-212); bc1 = 1 - beta1^step, bc2 = 1 - beta2^step - or, with step_dev != NULL, computed
 * from the device-resident step count (so the launch can sit in a CUDA graph).  g is left holding the clipped grad. */
int stage_adam_step(float* p, float* g, float* m, float* v, int64_t n, const float* sumsq, float max_norm, float lr, float beta1, float beta2,
                    float eps, float wd, float bc1, float bc2, float grad_scale, const int32_t* step_dev, void* stream);
int stage_counter_inc(int32_t* counter, void* stream);

/* ---- UMMA layout probe (test-only; pins the descriptor conventions the kernels rely on) ------------------------------ */
typedef struct {
  const uint8_t* a_img; const uint8_t* b_img;
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;
  uint32_t a_mn, b_mn;
  uint32_t k_steps, n, use_bulk;
  float* d_out; /* [128,n] */
  uint32_t repeat;      /* > 1: the k_steps MMAs are issued `repeat` times back to back (throughput measurement) */
  long long* cycles_out; /* optional: SM cycles from the first MMA issue to the completion of the last */
} StageProbeArgs;
int stage_umma_probe(const StageProbeArgs* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STAGE_B200_H_ */
