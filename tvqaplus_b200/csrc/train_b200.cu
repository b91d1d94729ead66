// libstage_b200.so, second translation unit: the C-ABI entry points of the training primitives (train.cuh): every operator of
// the STAGE path forward-with-saved-state and backward, the losses' gradients, and the clip + Adam step on flat buffers.
// Same rules as stage_b200.cu: plain pointers and sizes, asynchronous on the caller's stream, caller-owned memory.
#include "common.cuh"
#include "train.cuh"

using namespace stg;
using namespace stg::trn;

namespace {
inline int grid_for(long long work_items, int per_block) {  // streaming kernels: a multiple of the SM count, grid-stride inside
  const long long want = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

int stage_gemm(const StageGemmArgs* a, void* stream) {
  STG_REQUIRE(a && a->a && a->b && a->c, STAGE_E_BADARG, "gemm: null pointer");
  STG_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0 && a->nb0 > 0 && a->nb1 > 0, STAGE_E_BADARG, "gemm: m=%d n=%d k=%d batch=%dx%d", a->m, a->n, a->k, a->nb0, a->nb1);
  STG_REQUIRE(a->a_sm == 1 || a->a_sk == 1, STAGE_E_UNSUPPORTED, "gemm: A needs one unit stride (a_sm=%lld a_sk=%lld)", (long long)a->a_sm, (long long)a->a_sk);
  STG_REQUIRE(a->b_sk == 1 || a->b_sn == 1, STAGE_E_UNSUPPORTED, "gemm: B needs one unit stride (b_sk=%lld b_sn=%lld)", (long long)a->b_sk, (long long)a->b_sn);
  STG_REQUIRE(a->mode >= 0 && a->mode <= 2, STAGE_E_BADARG, "gemm: mode=%d", a->mode);
  const int split = a->split_k < 1 ? 1 : a->split_k;
  STG_REQUIRE(split == 1 || (a->mode == 2 && !a->relu), STAGE_E_BADARG, "gemm: split_k needs mode 2 (atomic) and no ReLU");
  const long long nz = static_cast<long long>(a->nb0) * a->nb1 * split;
  STG_REQUIRE(nz <= 65535, STAGE_E_UNSUPPORTED, "gemm: batch*split_k=%lld > 65535", nz);
  GemmDev d;
  d.A = a->a; d.B = a->b; d.C = a->c; d.bias = a->bias;
  d.m = a->m; d.n = a->n; d.k = a->k;
  d.a_sm = a->a_sm; d.a_sk = a->a_sk; d.b_sk = a->b_sk; d.b_sn = a->b_sn; d.ldc = a->ldc;
  d.nb1 = a->nb1;
  d.a_b0 = a->a_b0; d.a_b1 = a->a_b1; d.b_b0 = a->b_b0; d.b_b1 = a->b_b1; d.c_b0 = a->c_b0; d.c_b1 = a->c_b1;
  d.split_k = split; d.mode = a->mode; d.relu = a->relu; d.alpha = a->alpha;
  const bool a_kc = a->a_sk == 1 && a->k > 1 ? true : (a->a_sm == 1 ? false : true);
  const bool b_kc = a->b_sk == 1 && a->k > 1 ? true : (a->b_sn == 1 ? false : true);
  auto m4 = [](long long v) { return (v & 3) == 0; };
  d.vec_a = aligned16(a->a) && m4(a_kc ? a->a_sm : a->a_sk) && m4(a->a_b0) && m4(a->a_b1);
  d.vec_b = aligned16(a->b) && m4(b_kc ? a->b_sn : a->b_sk) && m4(a->b_b0) && m4(a->b_b1);
  d.vec_c = aligned16(a->c) && m4(a->ldc) && m4(a->c_b0) && m4(a->c_b1);
  dim3 grid((a->m + GBM - 1) / GBM, (a->n + GBN - 1) / GBN, static_cast<unsigned>(nz));
  STG_REQUIRE(grid.y <= 65535, STAGE_E_UNSUPPORTED, "gemm: n=%d too wide", a->n);
  cudaStream_t s = S(stream);
  if (a_kc && b_kc) sgemm_kernel<true, true><<<grid, 256, 0, s>>>(d);
  else if (a_kc && !b_kc) sgemm_kernel<true, false><<<grid, 256, 0, s>>>(d);
  else if (!a_kc && b_kc) sgemm_kernel<false, true><<<grid, 256, 0, s>>>(d);
  else sgemm_kernel<false, false><<<grid, 256, 0, s>>>(d);
  return check_launch("gemm");
}

int stage_ln_fwd(const float* x, const float* g, const float* b, float* y, float* mean, float* rstd, int64_t rows, int d, void* stream) {
  STG_REQUIRE(x && g && b && y && mean && rstd, STAGE_E_BADARG, "ln_fwd: null pointer");
  STG_REQUIRE(rows > 0 && d > 0 && d <= 768, STAGE_E_UNSUPPORTED, "ln_fwd: rows=%lld d=%d (d <= 768)", (long long)rows, d);
  const int grid = grid_for(rows, 8);
  cudaStream_t s = S(stream);
  const bool vec = d % 128 == 0 && aligned16(x) && aligned16(y) && aligned16(g) && aligned16(b);
  auto X = reinterpret_cast<const float4*>(x), G = reinterpret_cast<const float4*>(g), B = reinterpret_cast<const float4*>(b);
  auto Y = reinterpret_cast<float4*>(y);
  if (vec && d == 128) ln_fwd_vec_kernel<1><<<grid, 256, 0, s>>>(X, G, B, Y, mean, rstd, rows);
  else if (vec && d == 256) ln_fwd_vec_kernel<2><<<grid, 256, 0, s>>>(X, G, B, Y, mean, rstd, rows);
  else if (vec && d == 384) ln_fwd_vec_kernel<3><<<grid, 256, 0, s>>>(X, G, B, Y, mean, rstd, rows);
  else if (vec && d == 768) ln_fwd_vec_kernel<6><<<grid, 256, 0, s>>>(X, G, B, Y, mean, rstd, rows);
  else if (d <= 320) ln_fwd_kernel<10><<<grid, 256, 0, s>>>(x, g, b, y, mean, rstd, rows, d);
  else ln_fwd_kernel<LN_MAXPL><<<grid, 256, 0, s>>>(x, g, b, y, mean, rstd, rows, d);
  return check_launch("ln_fwd");
}
int stage_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* g, const float* add, float* dx,
                 float* dg, float* db, int64_t rows, int d, int relu_x, void* stream) {
  STG_REQUIRE(dy && x && mean && rstd && g && dg && db, STAGE_E_BADARG, "ln_bwd: null pointer");
  STG_REQUIRE(rows > 0 && d > 0 && d <= 768, STAGE_E_UNSUPPORTED, "ln_bwd: rows=%lld d=%d (d <= 768)", (long long)rows, d);
  const long long want = (rows + 63) / 64;  // at least 8 rows per warp: the column sums end in atomics
  const int grid = static_cast<int>(want < 1 ? 1 : (want < num_sms() * 8 ? want : num_sms() * 8));
  cudaStream_t s = S(stream);
  const bool vec = d % 128 == 0 && aligned16(x) && aligned16(dy) && aligned16(g) && (!add || aligned16(add)) && (!dx || aligned16(dx));
  auto DY = reinterpret_cast<const float4*>(dy), X = reinterpret_cast<const float4*>(x), G = reinterpret_cast<const float4*>(g);
  auto AD = reinterpret_cast<const float4*>(add);
  auto DX = reinterpret_cast<float4*>(dx);
  if (vec && d == 128) ln_bwd_vec_kernel<1><<<grid, 256, 0, s>>>(DY, X, mean, rstd, G, AD, DX, dg, db, rows, relu_x);
  else if (vec && d == 256) ln_bwd_vec_kernel<2><<<grid, 256, 0, s>>>(DY, X, mean, rstd, G, AD, DX, dg, db, rows, relu_x);
  else if (vec && d == 384) ln_bwd_vec_kernel<3><<<grid, 256, 0, s>>>(DY, X, mean, rstd, G, AD, DX, dg, db, rows, relu_x);
  else if (vec && d == 768) ln_bwd_vec_kernel<6><<<grid, 256, 0, s>>>(DY, X, mean, rstd, G, AD, DX, dg, db, rows, relu_x);
  else if (d <= 320) ln_bwd_kernel<10><<<grid, 256, 0, s>>>(dy, x, mean, rstd, g, add, dx, dg, db, rows, d, relu_x);
  else ln_bwd_kernel<LN_MAXPL><<<grid, 256, 0, s>>>(dy, x, mean, rstd, g, add, dx, dg, db, rows, d, relu_x);
  return check_launch("ln_bwd");
}
int stage_row_absmax(const float* x, int64_t rows, int d, float* out, void* stream) {
  STG_REQUIRE(x && out && rows > 0 && d > 0, STAGE_E_BADARG, "row_absmax: bad argument");
  row_absmax_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(x, rows, d, out);
  return check_launch("row_absmax");
}
int stage_colsum(const float* x, int64_t rows, int d, int64_t ld, float* out, void* stream) {
  STG_REQUIRE(x && out, STAGE_E_BADARG, "colsum: null pointer");
  STG_REQUIRE(rows > 0 && d > 0 && d <= 768 && ld >= d, STAGE_E_UNSUPPORTED, "colsum: rows=%lld d=%d ld=%lld", (long long)rows, d, (long long)ld);
  const long long want = (rows + 63) / 64;
  const int grid = static_cast<int>(want < 1 ? 1 : (want < num_sms() * 4 ? want : num_sms() * 4));
  colsum_kernel<<<grid, 256, 0, S(stream)>>>(x, rows, d, ld, out);
  return check_launch("colsum");
}
int stage_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  STG_REQUIRE(dy && y && dx && n > 0, STAGE_E_BADARG, "relu_bwd: bad argument");
  relu_bwd_kernel<<<grid_for(n, 1024), 256, 0, S(stream)>>>(dy, y, dx, n);
  return check_launch("relu_bwd");
}
int stage_add(const float* a, const float* b, float* y, int64_t n, int64_t period, void* stream) {
  STG_REQUIRE(a && b && y && n > 0 && period >= 0, STAGE_E_BADARG, "add: bad argument");
  add_kernel<<<grid_for(n, 1024), 256, 0, S(stream)>>>(a, b, y, n, period);
  return check_launch("add");
}
int stage_scale(float* x, int64_t n, const float* scalar, float mult, void* stream) {
  STG_REQUIRE(x && scalar && n > 0, STAGE_E_BADARG, "scale: bad argument");
  scale_kernel<<<grid_for(n, 1024), 256, 0, S(stream)>>>(x, n, scalar, mult);
  return check_launch("scale");
}
int stage_dropout(const float* x, float* y, int64_t n, float p, const int32_t* seed_dev, uint32_t site, void* stream) {
  STG_REQUIRE(x && y && seed_dev && n > 0 && p >= 0.f && p < 1.f, STAGE_E_BADARG, "dropout: bad argument (0 <= p < 1)");
  dropout_kernel<<<grid_for((n + 3) / 4, 1024), 256, 0, S(stream)>>>(x, y, n, p, seed_dev, site, aligned16(x) && aligned16(y));
  return check_launch("dropout");
}
int stage_permute_it(const float* x, float* y, int64_t groups, int i_dim, int t_dim, int d, void* stream) {
  STG_REQUIRE(x && y && x != y, STAGE_E_BADARG, "permute: null or aliased pointers");
  STG_REQUIRE(groups > 0 && i_dim > 0 && t_dim > 0 && d > 0 && d % 4 == 0 && aligned16(x) && aligned16(y), STAGE_E_BADARG, "permute: bad shape or alignment");
  permute_kernel<<<grid_for(groups * i_dim * t_dim * (d / 4), 1024), 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), groups, i_dim, t_dim, d / 4);
  return check_launch("permute");
}
int stage_dwconv_fwd(const float* x, const float* w, const float* bias, float* y, int64_t n_seq, int l, int d, int k, void* stream) {
  STG_REQUIRE(x && w && bias && y && x != y, STAGE_E_BADARG, "dwconv_fwd: null or aliased pointers");
  STG_REQUIRE(n_seq > 0 && l > 0 && d == 128 && k >= 1 && k <= 7 && (k & 1), STAGE_E_UNSUPPORTED, "dwconv_fwd: d=%d k=%d (d = 128, odd k <= 7)", d, k);
  STG_REQUIRE(aligned16(x) && aligned16(y) && aligned16(bias), STAGE_E_BADARG, "dwconv_fwd: alignment");
  dwconv_fwd_kernel<<<grid_for(n_seq * l, 8), 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(x), w, bias, reinterpret_cast<float4*>(y), n_seq, l, k);
  return check_launch("dwconv_fwd");
}
int stage_dwconv_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t n_seq, int l, int d, int k, void* stream) {
  STG_REQUIRE(dy && x && w && dx && dw && db && dx != dy, STAGE_E_BADARG, "dwconv_bwd: null or aliased pointers");
  STG_REQUIRE(n_seq > 0 && l > 0 && d == 128 && k >= 1 && k <= 7 && (k & 1), STAGE_E_UNSUPPORTED, "dwconv_bwd: d=%d k=%d (d = 128, odd k <= 7)", d, k);
  STG_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx), STAGE_E_BADARG, "dwconv_bwd: alignment");
  const long long want = (n_seq * l + 63) / 64;
  const int grid = static_cast<int>(want < 1 ? 1 : (want < num_sms() * 4 ? want : num_sms() * 4));
  dwconv_bwd_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), w, reinterpret_cast<float4*>(dx), dw, db, n_seq, l, k);
  return check_launch("dwconv_bwd");
}
int stage_l2norm_fwd(const float* x, float* y, float* inv, int64_t rows, int d, void* stream) {
  STG_REQUIRE(x && y && rows > 0 && d > 0 && d % 4 == 0 && aligned16(x) && aligned16(y), STAGE_E_BADARG, "l2norm_fwd: bad argument (d %% 4 == 0)");
  l2norm_fwd_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), inv, rows, d / 4);
  return check_launch("l2norm_fwd");
}
int stage_l2norm_bwd(const float* dy, const float* y, const float* inv, float* dx, int64_t rows, int d, int accumulate, void* stream) {
  STG_REQUIRE(dy && y && inv && dx && rows > 0 && d > 0 && d % 4 == 0 && aligned16(dy) && aligned16(y) && aligned16(dx), STAGE_E_BADARG, "l2norm_bwd: bad argument (d %% 4 == 0)");
  l2norm_bwd_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y), inv, reinterpret_cast<float4*>(dx), rows, d / 4, accumulate);
  return check_launch("l2norm_bwd");
}
int stage_attn_softmax_fwd(float* s, const float* c_mask, const float* q_mask, float* p, int64_t n, int rows_per_n, int t, int lc, float scale, void* stream) {
  STG_REQUIRE(s && c_mask && q_mask && p, STAGE_E_BADARG, "attn_softmax_fwd: null pointer");
  STG_REQUIRE(n > 0 && rows_per_n > 0 && t > 0 && lc > 0 && lc <= 128, STAGE_E_UNSUPPORTED, "attn_softmax_fwd: lc=%d (<= 128)", lc);
  const long long ct = static_cast<long long>(rows_per_n) * t, rows = n * ct;
  attn_softmax_fwd_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(s, c_mask, q_mask, p, rows, t, ct, lc, scale);
  return check_launch("attn_softmax_fwd");
}
int stage_attn_softmax_bwd(float* dp, const float* p, int64_t rows, int lc, float scale, void* stream) {
  STG_REQUIRE(dp && p && rows > 0 && lc > 0 && lc <= 128, STAGE_E_BADARG, "attn_softmax_bwd: bad argument");
  attn_softmax_bwd_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(dp, p, rows, lc, scale);
  return check_launch("attn_softmax_bwd");
}
int stage_cat3_fwd(const float* a, const float* b, float* out, int64_t rows, int a_div, void* stream) {
  STG_REQUIRE(a && b && out && rows > 0 && a_div > 0 && aligned16(a) && aligned16(b) && aligned16(out), STAGE_E_BADARG, "cat3_fwd: bad argument");
  cat3_fwd_kernel<<<grid_for(rows, 8), 256, 0, S(stream)>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out), rows, a_div);
  return check_launch("cat3_fwd");
}
int stage_cat3_bwd(const float* dcat, const float* a, const float* b, float* da, float* db, int64_t a_rows, int a_div, int accumulate_a, void* stream) {
  STG_REQUIRE(dcat && a && b && da && db && a_rows > 0 && a_rows < (1ll << 31) && a_div > 0, STAGE_E_BADARG, "cat3_bwd: bad argument");
  cat3_bwd_kernel<<<static_cast<unsigned>(a_rows), 128, 0, S(stream)>>>(dcat, a, b, da, db, a_div, accumulate_a);
  return check_launch("cat3_bwd");
}
int stage_maxpool_fwd(const float* x, const float* mask, float* y, int32_t* arg, int64_t n_seq, int l, void* stream) {
  STG_REQUIRE(x && mask && y && arg && n_seq > 0 && n_seq < (1ll << 31) && l > 0, STAGE_E_BADARG, "maxpool_fwd: bad argument");
  maxpool_fwd_kernel<<<static_cast<unsigned>(n_seq), 128, 0, S(stream)>>>(x, mask, y, arg, l);
  return check_launch("maxpool_fwd");
}
int stage_maxpool_bwd(const float* dy, const float* mask, const int32_t* arg, float* dx, int64_t n_seq, int l, void* stream) {
  STG_REQUIRE(dy && mask && arg && dx && n_seq > 0 && n_seq < (1ll << 31) && l > 0, STAGE_E_BADARG, "maxpool_bwd: bad argument");
  maxpool_bwd_kernel<<<static_cast<unsigned>(n_seq), 128, 0, S(stream)>>>(dy, mask, arg, dx, l);
  return check_launch("maxpool_bwd");
}
int stage_ts_loss_grad(const float* t_scores, const float* ts_mask, const int64_t* target, const int64_t* st, const int64_t* ed, int n, int t,
                       float* terms, float* probs, float* dscores, void* stream) {
  STG_REQUIRE(t_scores && ts_mask && target && st && ed && terms && probs && dscores, STAGE_E_BADARG, "ts_loss_grad: null pointer");
  STG_REQUIRE(n > 0 && t > 0, STAGE_E_BADARG, "ts_loss_grad: n=%d t=%d", n, t);
  ts_loss_kernel<<<n, 256, 0, S(stream)>>>(t_scores, ts_mask, reinterpret_cast<const long long*>(target), reinterpret_cast<const long long*>(st),
                                           reinterpret_cast<const long long*>(ed), t, terms, probs, dscores);
  return check_launch("ts_loss_grad");
}
int stage_proposal_pool_fwd(const float* x, const float* frame_mask, const int64_t* target, const int64_t* st, const int64_t* ed,
                            const int32_t* spans, const float* conf, int n, int t, int extra, float prob_thd, float* feat, int32_t* arg,
                            int32_t* valid, int64_t* row_target, void* stream) {
  STG_REQUIRE(x && frame_mask && target && st && ed && spans && conf && feat && arg && valid && row_target, STAGE_E_BADARG, "proposal_pool_fwd: null pointer");
  STG_REQUIRE(n > 0 && t > 0 && extra >= 0, STAGE_E_BADARG, "proposal_pool_fwd: n=%d t=%d", n, t);
  proposal_pool_fwd_kernel<<<2 * n, 128, 0, S(stream)>>>(x, frame_mask, reinterpret_cast<const long long*>(target), reinterpret_cast<const long long*>(st),
                                                         reinterpret_cast<const long long*>(ed), spans, conf, t, extra, prob_thd, feat, arg, valid,
                                                         reinterpret_cast<long long*>(row_target));
  return check_launch("proposal_pool_fwd");
}
int stage_proposal_pool_bwd(const float* dfeat, const int32_t* arg, const int32_t* valid, const float* frame_mask, int n, int t, float* dx, void* stream) {
  STG_REQUIRE(dfeat && arg && valid && frame_mask && dx && n > 0 && t > 0, STAGE_E_BADARG, "proposal_pool_bwd: bad argument");
  proposal_pool_bwd_kernel<<<2 * n * 5, 256, 0, S(stream)>>>(dfeat, arg, valid, frame_mask, t, dx);
  return check_launch("proposal_pool_bwd");
}
int stage_att_loss_bwd(const float* scores, const int64_t* pos_off, const int64_t* neg_off, int64_t m, int loss_type, float alpha, float margin,
                       const float* upstream, float* dscores, void* stream) {
  STG_REQUIRE(scores && pos_off && neg_off && upstream && dscores, STAGE_E_BADARG, "att_loss_bwd: null pointer");
  STG_REQUIRE(m > 0 && (loss_type == 0 || loss_type == 1), STAGE_E_BADARG, "att_loss_bwd: m=%lld loss_type=%d", (long long)m, loss_type);
  att_loss_bwd_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, S(stream)>>>(scores, reinterpret_cast<const long long*>(pos_off),
                                                                                    reinterpret_cast<const long long*>(neg_off), m, loss_type, alpha, margin, upstream, dscores);
  return check_launch("att_loss_bwd");
}
int stage_answer_ce(const float* rows, const int32_t* valid, const int64_t* row_target, int n_rows, int n_examples, float* out, float* drows, void* stream) {
  STG_REQUIRE(rows && valid && row_target && out && drows && n_rows > 0 && n_examples > 0, STAGE_E_BADARG, "answer_ce: bad argument");
  answer_ce_kernel<<<1, 256, 0, S(stream)>>>(rows, valid, reinterpret_cast<const long long*>(row_target), n_rows, n_examples, out, drows);
  return check_launch("answer_ce");
}
int stage_grad_sumsq(const float* g, int64_t n, float* out, void* stream) {
  STG_REQUIRE(g && out && n > 0, STAGE_E_BADARG, "grad_sumsq: bad argument");
  sumsq_kernel<<<grid_for(n, 2048), 256, 0, S(stream)>>>(g, n, out);
  return check_launch("grad_sumsq");
}
int stage_adam_step(float* p, float* g, float* m, float* v, int64_t n, const float* sumsq, float max_norm, float lr, float beta1, float beta2,
                    float eps, float wd, float bc1, float bc2, float grad_scale, const int32_t* step_dev, void* stream) {
  STG_REQUIRE(p && g && m && v && n > 0 && (max_norm <= 0.f || sumsq), STAGE_E_BADARG, "adam_step: bad argument");
  adam_kernel<<<grid_for(n, 1024), 256, 0, S(stream)>>>(p, g, m, v, n, sumsq, max_norm, lr, beta1, beta2, eps, wd, bc1, bc2, grad_scale, step_dev);
  return check_launch("adam_step");
}
int stage_counter_inc(int32_t* counter, void* stream) {
  STG_REQUIRE(counter, STAGE_E_BADARG, "counter_inc: null pointer");
  counter_inc_kernel<<<1, 1, 0, S(stream)>>>(counter);
  return check_launch("counter_inc");
}

}  // extern "C"
