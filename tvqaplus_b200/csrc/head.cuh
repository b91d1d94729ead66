// Span + answer heads (fp32 CUDA-core kernel, warp-shuffle reductions): one CTA per (example, answer) row.
//   y_t   = x_t + relu(Lin128(LN(x_t)))                              residual_temporal_predictor  (stage.py:478)
//   st/ed = Lin1(LN(y_t)) ; masked with ts_label_mask                (stage.py:479-481,521)
//   p     = softmax over frames ; best (st <= ed) = argmax p_st*p_ed (model_utils.py:92-121, O(T) prefix-max scan)
//   local = masked max over frames [st-E, ed+1+E), global = masked max over all frames          (stage.py:449-466)
//   answer = Lin1(LN([local, global]))                               classifier                   (stage.py:536)
// The whole head is ~0.05 GFLOP per example; what matters is that it removes the reference's device->host sync
// (numpy argsort of T*T products per row) and reproduces the -1e10 / LayerNorm corner cases exactly.
//
// Tie rule of the span search: among equal fp32 products the larger flat index st*T+ed wins (the reference takes the
// last element of an ascending argsort; numpy's default sort is not stable, so ties are not defined there).
#pragma once
#include "umma.cuh"

namespace stg {

struct HeadDev {
  int n, t, lq, add_local, extra;
  const float* max_statement;   // [n*5, t, 128]
  const float* statement_mask;  // [n,5,t,lq]
  const float* ts_mask;         // [n,t]
  const float *proj_g, *proj_b, *proj_w, *proj_bias;
  const float *st_g, *st_b, *st_w, *st_bias;
  const float *ed_g, *ed_b, *ed_w, *ed_bias;
  const float *cls_g, *cls_b, *cls_w, *cls_bias;
  float* t_scores; float* t_probs; int* spans; float* span_conf; float* answer; float* enc;
  // optional: temporal loss terms of the ground-truth answer rows (get_ts_loss, stage.py:539-555) and what the train-mode
  // proposal kernel needs (frame mask, global feature)
  const long long* target; const long long* ts_st; const long long* ts_ed;
  float* ts_terms;     // [n,2]: -log softmax(scores[:, c])[label_c] of the ground-truth answer row
  float* frame_mask;   // [n*5, t]
  float* gfeat;        // [n*5, 128]
};

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red, int tid) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// LayerNorm of one 128-vector held as 4 values per lane
__device__ __forceinline__ float4 ln128(float4 v, float4 g, float4 b) {
  const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
  const float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
  const float var = warp_sum((a * a + c * c) + (d * d + e * e)) * (1.f / 128.f);
  const float rs = 1.f / sqrtf(var + 1e-5f);
  return make_float4(a * rs * g.x + b.x, c * rs * g.y + b.y, d * rs * g.z + b.z, e * rs * g.w + b.w);
}

// the same LayerNorm split into its statistics (two-pass: a constant -1e10 row must give exactly beta) and its affine part,
// for rows that feed two LayerNorms with different gamma / beta
__device__ __forceinline__ float4 ln128_center(float4 v, float& rs) {
  const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
  const float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
  const float var = warp_sum((a * a + c * c) + (d * d + e * e)) * (1.f / 128.f);
  rs = 1.f / sqrtf(var + 1e-5f);
  return make_float4(a, c, d, e);
}
__device__ __forceinline__ float4 ln128_affine(float4 c, float rs, float4 g, float4 b) {
  return make_float4(c.x * rs * g.x + b.x, c.y * rs * g.y + b.y, c.z * rs * g.z + b.z, c.w * rs * g.w + b.w);
}

__global__ void __launch_bounds__(256, 2) head_kernel(HeadDev p) {
  extern __shared__ __align__(16) float hs[];
  float* Wt = hs;                      // [128 k][128 c] transposed projection weight
  float* lnv = Wt + 128 * 128;         // [8 warps][2 frames][128] LN(x) staging
  float* sc = lnv + 16 * 128;           // [t][2] masked scores, then probabilities
  float* fm = sc + 2 * p.t;            // [t] frame mask (max_statement_mask)
  float* feat = fm + p.t;              // [256] classifier input
  __shared__ float red[8];
  __shared__ int span_sh[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = blockIdx.x;          // n*5 + a
  const int n_i = row / 5;
  const int T = p.t;

  for (int i = tid; i < 128 * 128; i += 256) {
    const int c = i >> 7, k = i & 127;
    Wt[k * 128 + c] = __ldg(p.proj_w + i);
  }
  // frame mask = "any token of the statement is valid" (max_statement_mask, stage.py:505): the row's T*lq mask values are
  // one contiguous block; every thread keeps eight independent loads in flight and raises the frame's flag
  for (int t = tid; t < T; t += 256) fm[t] = 0.f;
  __syncthreads();
  {
    const float* mrow = p.statement_mask + static_cast<size_t>(row) * T * p.lq;
    const int total = T * p.lq;
    for (int base = tid; base < total; base += 256 * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (base + u * 256 < total) ? __ldg(mrow + base + u * 256) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (v[u] != 0.f) fm[(base + u * 256) / p.lq] = 1.f;  // benign race: every writer stores the same value
    }
  }
  __syncthreads();
  if (p.frame_mask != nullptr)
    for (int t = tid; t < T; t += 256) p.frame_mask[static_cast<size_t>(row) * T + t] = fm[t];

  const float4 pg = __ldg(reinterpret_cast<const float4*>(p.proj_g) + lane), pb = __ldg(reinterpret_cast<const float4*>(p.proj_b) + lane);
  const float4 pbias = __ldg(reinterpret_cast<const float4*>(p.proj_bias) + lane);
  const float4 sg = __ldg(reinterpret_cast<const float4*>(p.st_g) + lane), sb = __ldg(reinterpret_cast<const float4*>(p.st_b) + lane);
  const float4 sw = __ldg(reinterpret_cast<const float4*>(p.st_w) + lane);
  const float4 eg = __ldg(reinterpret_cast<const float4*>(p.ed_g) + lane), eb = __ldg(reinterpret_cast<const float4*>(p.ed_b) + lane);
  const float4 ew = __ldg(reinterpret_cast<const float4*>(p.ed_w) + lane);
  const float st_bias = __ldg(p.st_bias), ed_bias = __ldg(p.ed_bias);

  // ---- per frame: residual projection + st/ed logits (warp per PAIR of frames, lane = 4 channels: one weight row read
  //      from shared memory serves both frames and the eight accumulator chains hide the FMA latency)
  float4 gmax = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);  // masked max over this warp's frames (global feature)
  for (int t0 = warp * 2; t0 < T; t0 += 16) {
    const bool two = t0 + 1 < T;
    const size_t off0 = (static_cast<size_t>(row) * T + t0) * 128, off1 = off0 + 128;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(p.max_statement + off0) + lane);
    const float4 x1 = two ? __ldg(reinterpret_cast<const float4*>(p.max_statement + off1) + lane) : x0;
    const float4 l0 = ln128(x0, pg, pb), l1 = ln128(x1, pg, pb);
    float* lv = lnv + warp * 256;
    __syncwarp();
    *reinterpret_cast<float4*>(lv + lane * 4) = l0;
    *reinterpret_cast<float4*>(lv + 128 + lane * 4) = l1;
    __syncwarp();
    float4 acc0 = pbias, acc1 = pbias;
#pragma unroll 4
    for (int k = 0; k < 128; k += 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(lv + k), a1 = *reinterpret_cast<const float4*>(lv + 128 + k);
      const float av0[4] = {a0.x, a0.y, a0.z, a0.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 w = *reinterpret_cast<const float4*>(Wt + (k + e) * 128 + lane * 4);
        acc0.x = fmaf(av0[e], w.x, acc0.x); acc0.y = fmaf(av0[e], w.y, acc0.y); acc0.z = fmaf(av0[e], w.z, acc0.z); acc0.w = fmaf(av0[e], w.w, acc0.w);
        acc1.x = fmaf(av1[e], w.x, acc1.x); acc1.y = fmaf(av1[e], w.y, acc1.y); acc1.z = fmaf(av1[e], w.z, acc1.z); acc1.w = fmaf(av1[e], w.w, acc1.w);
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (f == 1 && !two) break;
      const int t = t0 + f;
      const float4 x = f == 0 ? x0 : x1, acc = f == 0 ? acc0 : acc1;
      const size_t off = f == 0 ? off0 : off1;
      const float4 y = make_float4(x.x + fmaxf(acc.x, 0.f), x.y + fmaxf(acc.y, 0.f), x.z + fmaxf(acc.z, 0.f), x.w + fmaxf(acc.w, 0.f));
      *reinterpret_cast<float4*>(p.enc + off + lane * 4) = y;
      {
        const float m = fm[t], o = (1.f - m) * STG_NEG;  // mask_logits, then max over frames (stage.py:456-466)
        gmax = make_float4(fmaxf(gmax.x, y.x * m + o), fmaxf(gmax.y, y.y * m + o), fmaxf(gmax.z, y.z * m + o), fmaxf(gmax.w, y.w * m + o));
      }
      float rs_y;
      const float4 yc = ln128_center(y, rs_y);  // one set of statistics serves the start and the end scorer
      const float4 ls = ln128_affine(yc, rs_y, sg, sb), le = ln128_affine(yc, rs_y, eg, eb);
      const float s_st = warp_sum((ls.x * sw.x + ls.y * sw.y) + (ls.z * sw.z + ls.w * sw.w)) + st_bias;
      const float s_ed = warp_sum((le.x * ew.x + le.y * ew.y) + (le.z * ew.z + le.w * ew.w)) + ed_bias;
      if (lane == 0) {
        const float m = __ldg(p.ts_mask + static_cast<size_t>(n_i) * T + t);
        const float off_m = (1.f - m) * STG_NEG;
        const float a0 = s_st * m + off_m, a1 = s_ed * m + off_m;  // mask_logits
        sc[2 * t] = a0; sc[2 * t + 1] = a1;
        p.t_scores[(static_cast<size_t>(row) * T + t) * 2] = a0;
        p.t_scores[(static_cast<size_t>(row) * T + t) * 2 + 1] = a1;
      }
    }
  }
  __syncwarp();
  *reinterpret_cast<float4*>(lnv + warp * 256 + lane * 4) = gmax;  // the staging rows are free now: [warp][128] partial maxima
  __syncthreads();

  // ---- softmax over frames for st (c=0) and ed (c=1); on the ground-truth answer's row also the cross-entropy term of
  //      get_ts_loss (stage.py:539-555): -log_softmax(scores)[label] = mx + log(sum exp(s - mx)) - s[label]
  const bool gt_row = p.ts_terms != nullptr && (row - n_i * 5) == static_cast<int>(__ldg(p.target + n_i));
  for (int c = 0; c < 2; ++c) {
    float mx = -3.0e38f, raw_label = 0.f;
    for (int t = tid; t < T; t += 256) mx = fmaxf(mx, sc[2 * t + c]);
    // (read before the reduction's barriers: behind them the other threads start turning the scores into exponentials)
    if (gt_row && tid == 0) raw_label = sc[2 * static_cast<int>(__ldg((c == 0 ? p.ts_st : p.ts_ed) + n_i)) + c];
    mx = block_reduce(mx, true, red, tid);
    float s = 0.f;
    for (int t = tid; t < T; t += 256) {
      const float e = expf(sc[2 * t + c] - mx);
      sc[2 * t + c] = e;
      s += e;
    }
    s = block_reduce(s, false, red, tid);
    if (gt_row && tid == 0) p.ts_terms[2 * n_i + c] = (mx + logf(s)) - raw_label;
    for (int t = tid; t < T; t += 256) {
      const float pr = sc[2 * t + c] / s;
      sc[2 * t + c] = pr;
      p.t_probs[(static_cast<size_t>(row) * T + t) * 2 + c] = pr;
    }
  }
  __syncthreads();

  // ---- best span: max over st <= ed of p_st*p_ed via prefix max (ties: larger st, then larger ed)
  if (tid == 0) {
    float pm = -1.f, best = -1.f;
    int pi = 0, bs = 0, be = 0;
    for (int j = 0; j < T; ++j) {
      const float a = sc[2 * j];
      if (a >= pm) { pm = a; pi = j; }
      const float v = pm * sc[2 * j + 1];
      if (v > best || (v == best && (pi > bs || (pi == bs && j > be)))) { best = v; bs = pi; be = j; }
    }
    span_sh[0] = bs; span_sh[1] = be;
    p.spans[2 * row] = bs; p.spans[2 * row + 1] = be;
    p.span_conf[row] = best;
  }
  __syncthreads();

  // ---- global masked max: the eight per-warp partial maxima; span-local masked max: frames [st-E, ed+1+E) re-read from
  //      the workspace this CTA wrote (L2), four loads in flight per thread
  {
    const int c = tid & 127;
    if (tid >= 128) {
      float mx = lnv[c];
#pragma unroll
      for (int w = 1; w < 8; ++w) mx = fmaxf(mx, lnv[w * 256 + c]);
      feat[p.add_local ? tid : c] = mx;
      if (p.gfeat != nullptr) p.gfeat[static_cast<size_t>(row) * 128 + c] = mx;
    } else if (p.add_local) {
      int lo = span_sh[0] - p.extra, hi = span_sh[1] + 1 + p.extra;
      if (lo < 0) lo = 0;
      if (hi > T) hi = T;
      const float* e = p.enc + static_cast<size_t>(row) * T * 128 + c;
      float mx = -3.0e38f;
      int t = lo;
      for (; t + 4 <= hi; t += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcg(e + static_cast<size_t>(t + u) * 128);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float m = fm[t + u];
          mx = fmaxf(mx, v[u] * m + (1.f - m) * STG_NEG);
        }
      }
      for (; t < hi; ++t) {
        const float m = fm[t];
        mx = fmaxf(mx, __ldcg(e + static_cast<size_t>(t) * 128) * m + (1.f - m) * STG_NEG);
      }
      feat[tid] = mx;
    }
  }
  __syncthreads();

  // ---- classifier: LN(feat) . w + b
  {
    const int F = p.add_local ? 256 : 128;
    const float v = tid < F ? feat[tid] : 0.f;
    const float mean = block_reduce(v, false, red, tid) / static_cast<float>(F);
    const float d = tid < F ? v - mean : 0.f;
    const float var = block_reduce(d * d, false, red, tid) / static_cast<float>(F);
    const float rs = 1.f / sqrtf(var + 1e-5f);
    const float term = tid < F ? (d * rs * __ldg(p.cls_g + tid) + __ldg(p.cls_b + tid)) * __ldg(p.cls_w + tid) : 0.f;
    const float tot = block_reduce(term, false, red, tid);
    if (tid == 0) p.answer[row] = tot + __ldg(p.cls_bias);
  }
}

// (st + ed) / 2 summed over the examples, in a fixed order (one thread: n is a batch size)
__global__ void ts_loss_sum_kernel(const float* __restrict__ terms, int n, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float st = 0.f, ed = 0.f;
    for (int i = 0; i < n; ++i) { st += terms[2 * i]; ed += terms[2 * i + 1]; }
    *loss = (st + ed) / 2.f;
  }
}

// Train branch of get_proposals + classifier (stage.py:407-438,536).  Candidate row r = 2*example + kind: kind 0 = the
// ground-truth span [st, ed+1), kind 1 = the predicted span of the ground-truth answer's row (head_kernel's prefix-max
// search on that row, as find_max_triples(topN=1, prob_thd) does), kept only if its confidence >= ce_prob_thd and its
// temporal IoU with the ground truth >= 0.5 (compute_temporal_iou: intersection over the convex hull, model_utils.py:37-51).
// Every kept row scores all five answers: [masked max over the span's frames (+- extra), masked max over all frames] ->
// LayerNorm(256) -> Linear.  One CTA per candidate row.
struct ProposalDev {
  int n, t, extra;
  float ce_prob_thd;
  const float* enc;          // [n,5,t,128] encoded max_statement (head_kernel workspace)
  const float* frame_mask;   // [n*5,t]
  const float* gfeat;        // [n*5,128]
  const int* spans;          // [n*5,2] inclusive (st, ed) of every row
  const float* span_conf;    // [n*5]
  const long long* target; const long long* ts_st; const long long* ts_ed;
  const float *cls_g, *cls_b, *cls_w, *cls_bias;
  float* answer_rows;        // [2n,5]
  int* row_valid;            // [2n]
  long long* row_target;     // [2n]
};

__global__ void __launch_bounds__(256) proposal_cls_kernel(ProposalDev p) {
  __shared__ float feat[256];
  __shared__ float red[8];
  __shared__ int sh[3];
  const int tid = threadIdx.x, r = blockIdx.x, n_i = r >> 1, kind = r & 1, T = p.t;
  if (tid == 0) {
    const int tgt = static_cast<int>(__ldg(p.target + n_i));
    const int g0 = static_cast<int>(__ldg(p.ts_st + n_i)), g1 = static_cast<int>(__ldg(p.ts_ed + n_i)) + 1;
    int s0 = g0, s1 = g1, ok = 1;
    if (kind == 1) {
      const int row = n_i * 5 + tgt;
      s0 = p.spans[2 * row];
      s1 = p.spans[2 * row + 1] + 1;
      const int inter = max(0, min(s1, g1) - max(s0, g0)), uni = max(s1, g1) - min(s0, g0);
      // float(conf) >= prob_thd and 1.0*inter/union >= 0.5 of the reference, both evaluated exactly
      ok = (static_cast<double>(p.span_conf[row]) >= static_cast<double>(p.ce_prob_thd)) && uni != 0 && 2 * inter >= uni;
    }
    sh[0] = max(0, s0 - p.extra);       // expand_span + python slice clipping (model_utils.py:78-89)
    sh[1] = min(T, s1 + p.extra);
    sh[2] = ok;
    p.row_valid[r] = ok;
    p.row_target[r] = tgt;
  }
  __syncthreads();
  const int lo = sh[0], hi = sh[1];
  if (!sh[2]) {
    if (tid < 5) p.answer_rows[r * 5 + tid] = 0.f;
    return;
  }
  const int c = tid & 127;
  for (int a = 0; a < 5; ++a) {
    const int row = n_i * 5 + a;
    if (tid >= 128) {
      feat[tid] = __ldg(p.gfeat + static_cast<size_t>(row) * 128 + c);
    } else {
      const float* e = p.enc + static_cast<size_t>(row) * T * 128 + c;
      const float* fm = p.frame_mask + static_cast<size_t>(row) * T;
      float mx = -3.0e38f;  // an empty slice cannot occur: st <= ed
      for (int t = lo; t < hi; ++t) {
        const float m = __ldg(fm + t);
        mx = fmaxf(mx, __ldg(e + static_cast<size_t>(t) * 128) * m + (1.f - m) * STG_NEG);
      }
      feat[tid] = mx;
    }
    __syncthreads();
    const float v = feat[tid];
    const float mean = block_reduce(v, false, red, tid) / 256.f;
    const float d = v - mean;
    const float var = block_reduce(d * d, false, red, tid) / 256.f;
    const float rs = 1.f / sqrtf(var + 1e-5f);
    const float term = (d * rs * __ldg(p.cls_g + tid) + __ldg(p.cls_b + tid)) * __ldg(p.cls_w + tid);
    const float tot = block_reduce(term, false, red, tid);
    if (tid == 0) p.answer_rows[r * 5 + a] = tot + __ldg(p.cls_bias);
    __syncthreads();
  }
}

__global__ void find_max_span_kernel(const float* __restrict__ p_st, const float* __restrict__ p_ed, int rows, int T,
                                     int* __restrict__ spans, float* __restrict__ conf) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float pm = -1.f, best = -1.f;
  int pi = 0, bs = 0, be = 0;
  for (int j = 0; j < T; ++j) {
    const float a = p_st[static_cast<size_t>(r) * T + j];
    if (a >= pm) { pm = a; pi = j; }
    const float v = pm * p_ed[static_cast<size_t>(r) * T + j];
    if (v > best || (v == best && (pi > bs || (pi == bs && j > be)))) { best = v; bs = pi; be = j; }
  }
  spans[2 * r] = bs; spans[2 * r + 1] = be;
  conf[r] = best;
}

// inference.py:13-35 find_max_pair: running FIRST argmax of p1 (strict '<'), best p1[k1]*p2[i] under strict '>' starting
// from value 0 / span (0, 1).  The reference multiplies Python floats, i.e. the exact double products of the fp32
// probabilities, so the comparison runs in double here as well (an fp32 x fp32 product is exact in fp64).
__global__ void find_max_pair_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int rows, int T,
                                     long long row_stride, int elem_stride, int* __restrict__ spans,
                                     double* __restrict__ val) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* a = p1 + static_cast<size_t>(r) * row_stride;
  const float* b = p2 + static_cast<size_t>(r) * row_stride;
  double best = 0.0;
  int k1 = 0, bs = 0, be = 1;
  float v1 = T > 0 ? a[0] : 0.f;
  for (int i = 0; i < T; ++i) {
    const float ai = a[static_cast<size_t>(i) * elem_stride];
    if (v1 < ai) { k1 = i; v1 = ai; }
    const double v = static_cast<double>(v1) * static_cast<double>(b[static_cast<size_t>(i) * elem_stride]);
    if (v > best) { best = v; bs = k1; be = i; }
  }
  spans[2 * r] = bs; spans[2 * r + 1] = be;
  val[r] = best;
}

// Attention ranking loss (get_att_loss, stage.py:731-746): for every sampled (positive, negative) pair of region scores
//   lse:   log1p(exp(alpha * (s_neg - s_pos)))          hinge:   max(0, margin + s_neg - s_pos)
// summed over the pairs.  The pair lists are built once on the host from att_labels (they depend on the labels and the
// host RNG only, not on any device value); here: gather, loss, block sums; a second one-thread pass adds the block sums in
// order so that the result does not depend on the launch's timing.
__global__ void __launch_bounds__(256) att_loss_kernel(const float* __restrict__ scores, const long long* __restrict__ pos_off,
                                                       const long long* __restrict__ neg_off, long long m, int hinge, float alpha,
                                                       float margin, float* __restrict__ partials) {
  __shared__ float red[8];
  const int tid = threadIdx.x;
  const long long i = static_cast<long long>(blockIdx.x) * 256 + tid;
  float v = 0.f;
  if (i < m) {
    const float d = __ldg(scores + __ldg(neg_off + i)) - __ldg(scores + __ldg(pos_off + i));
    v = hinge ? fmaxf(margin + d, 0.f) : log1pf(expf(alpha * d));
  }
  const float tot = block_reduce(v, false, red, tid);
  if (tid == 0) partials[blockIdx.x] = tot;
}
__global__ void sum_in_order_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += partials[i];
    *out = s;
  }
}

// pad_sequences_2d on the device (tvqa_dataset.py:515-553): the valid rows of all sequences arrive concatenated
// (`rows`, D floats each, sequence s at row offset off[s]); sequence s becomes rows [s*l_pad, (s+1)*l_pad) of `out`, its
// first min(len[s], l_pad) rows copied, the rest zero-filled, and mask[s, j] = j < len[s].  One warp per output row.
__global__ void pad_rows_kernel(const float* __restrict__ rows, const long long* __restrict__ off, const int* __restrict__ len,
                                long long n_seq, int l_pad, int d4, float* __restrict__ out, float* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const long long n_rows = n_seq * l_pad;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_rows;
       r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const long long s = r / l_pad;
    const int j = static_cast<int>(r - s * l_pad);
    const bool on = j < __ldg(len + s);
    float4* dst = reinterpret_cast<float4*>(out) + r * d4;
    if (on) {
      const float4* src = reinterpret_cast<const float4*>(rows) + (__ldg(off + s) + j) * d4;
      for (int c = lane; c < d4; c += 32) dst[c] = __ldg(src + c);
    } else {
      for (int c = lane; c < d4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (lane == 0) mask[r] = on ? 1.f : 0.f;
  }
}

__global__ void mask_logits_kernel(const float* __restrict__ x, const float* __restrict__ m, long long outer, int inner,
                                   float* __restrict__ y) {
  const long long total = outer * inner;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float mm = m[i / inner];
    y[i] = x[i] * mm + (1.f - mm) * STG_NEG;
  }
}

}  // namespace stg
