// Training primitives: forward-with-saved-statistics and backward kernels for every operator of the STAGE path
// (reference model/stage.py:192-555, model/encoder.py:29-52, model/cnn.py:42-47, model/context_query_attention.py:35-101;
// the reference gets these from PyTorch autograd, main.py:61).  Exact fp32 on the CUDA cores: the gradient GEMMs reduce over
// millions of rows and are checked against autograd at 1e-4, so they do not use the split-fp16 tensor-core scheme of the
// inference kernels; the strided, batched GEMM below is the one workhorse, everything else is a streaming kernel.
// Row order of the attended tensors in the training engine is (n, a, i, t) - question word before frame - so that every
// contraction of the context-query attention and every reduction over frames is a plain strided GEMM (train.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stg {
namespace trn {

#define TRN_NEG (-1e10f)

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// C[b] (op)= alpha * A[b] * B[b] (+ bias) (ReLU): A(m,k) at A + m*a_sm + k*a_sk, B(k,n) at B + k*b_sk + n*b_sn, C row-major
// with leading dimension ldc; two batch indices with their own strides per operand; split over K with atomic accumulation.
// ------------------------------------------------------------------------------------------------------------------
struct GemmDev {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int m, n, k;
  long long a_sm, a_sk, b_sk, b_sn, ldc;
  int nb1;
  long long a_b0, a_b1, b_b0, b_b1, c_b0, c_b1;
  int split_k, mode, relu, vec_a, vec_b, vec_c;  // mode 0 store, 1 C +=, 2 atomicAdd
  float alpha;
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GLD = 132;

// KC: the k index is the contiguous one in memory (row stride s_r); otherwise the row index is (k stride s_k)
template <bool KC>
__device__ __forceinline__ void g_load(const float* __restrict__ base, long long s_r, long long s_k, int r0, int k0, int R, int K1,
                                       int tid, bool vec, float (&v)[8]) {
  if (KC) {
    const int r = r0 + (tid >> 1), kk = k0 + ((tid & 1) << 3);
    const float* p = base + static_cast<long long>(r) * s_r + kk;
    if (vec && r < R && kk + 8 <= K1) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (r < R && kk + i < K1) ? __ldg(p + i) : 0.f;
    }
  } else {
    const int kk = k0 + (tid >> 4), r = r0 + ((tid & 15) << 3);
    const float* p = base + static_cast<long long>(kk) * s_k + r;
    if (vec && kk < K1 && r + 8 <= R) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (kk < K1 && r + i < R) ? __ldg(p + i) : 0.f;
    }
  }
}
template <bool KC>
__device__ __forceinline__ void g_store(float* __restrict__ s, int tid, const float (&v)[8]) {  // s: [GBK][GLD]
  if (KC) {
    const int r = tid >> 1, kk = (tid & 1) << 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[(kk + i) * GLD + r] = v[i];
  } else {
    const int kk = tid >> 4, r = (tid & 15) << 3;
    *reinterpret_cast<float4*>(s + kk * GLD + r) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(s + kk * GLD + r + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(GemmDev d) {
  __shared__ __align__(16) float As[2][GBK * GLD];
  __shared__ __align__(16) float Bs[2][GBK * GLD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int z = blockIdx.z, split = z % d.split_k, b = z / d.split_k, b0 = b / d.nb1, b1 = b - b0 * d.nb1;
  const float* A = d.A + b0 * d.a_b0 + b1 * d.a_b1;
  const float* B = d.B + b0 * d.b_b0 + b1 * d.b_b1;
  float* C = d.C + b0 * d.c_b0 + b1 * d.c_b1;
  int kper = (d.k + d.split_k - 1) / d.split_k;
  kper = (kper + GBK - 1) / GBK * GBK;
  const int kbeg = split * kper, kend = min(d.k, kbeg + kper);
  if (kbeg >= kend) return;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  const int nt = (kend - kbeg + GBK - 1) / GBK;
  g_load<A_KC>(A, d.a_sm, d.a_sk, m0, kbeg, d.m, kend, tid, d.vec_a, ra);
  g_load<B_KC>(B, d.b_sn, d.b_sk, n0, kbeg, d.n, kend, tid, d.vec_b, rb);
  g_store<A_KC>(As[0], tid, ra);
  g_store<B_KC>(Bs[0], tid, rb);
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    if (t + 1 < nt) {
      g_load<A_KC>(A, d.a_sm, d.a_sk, m0, kbeg + (t + 1) * GBK, d.m, kend, tid, d.vec_a, ra);
      g_load<B_KC>(B, d.b_sn, d.b_sk, n0, kbeg + (t + 1) * GBK, d.n, kend, tid, d.vec_b, rb);
    }
    const float* as = As[t & 1];
    const float* bs = Bs[t & 1];
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(as + kk * GLD + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(as + kk * GLD + 64 + ty * 4);
      const float4 c0 = *reinterpret_cast<const float4*>(bs + kk * GLD + tx * 4);
      const float4 c1 = *reinterpret_cast<const float4*>(bs + kk * GLD + 64 + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nt) {
      g_store<A_KC>(As[(t + 1) & 1], tid, ra);
      g_store<B_KC>(Bs[(t + 1) & 1], tid, rb);
    }
    __syncthreads();
  }
  const bool lead = split == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (r >= d.m) continue;
    float* crow = C + static_cast<long long>(r) * d.ldc;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = n0 + h * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = d.alpha * acc[i][h * 4 + j];
        if (d.bias != nullptr && lead && c + j < d.n) x += __ldg(d.bias + c + j);
        if (d.relu) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      if (d.mode == 0 && d.vec_c && c + 4 <= d.n) {
        *reinterpret_cast<float4*>(crow + c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c + j >= d.n) continue;
          if (d.mode == 0) crow[c + j] = v[j];
          else if (d.mode == 1) crow[c + j] += v[j];
          else atomicAdd(crow + c + j, v[j]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, biased variance, two-pass): one warp per row, up to 768 columns (24 per lane).
// ------------------------------------------------------------------------------------------------------------------
constexpr int LN_MAXPL = 24;

// ---- d % 128 == 0 (128, 256, 384, 768): NV float4 per lane, lane l holds columns (i*32 + l)*4 .. +3
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ g, const float4* __restrict__ b,
                                                         float4* __restrict__ y, float* __restrict__ mean_o, float* __restrict__ rstd_o, long long rows) {
  constexpr int D4 = NV * 32;
  const int lane = threadIdx.x & 31;
  const float inv_d = 1.f / (D4 * 4);
  float4 gv[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { gv[i] = __ldg(g + i * 32 + lane); bv[i] = __ldg(b + i * 32 + lane); }
  const long long w0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += nw) {
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = __ldg(x + r * D4 + i * 32 + lane);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = wsum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = 1.f / sqrtf(wsum(q) * inv_d + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      y[r * D4 + i * 32 + lane] = make_float4(v[i].x * rstd * gv[i].x + bv[i].x, v[i].y * rstd * gv[i].y + bv[i].y,
                                              v[i].z * rstd * gv[i].z + bv[i].z, v[i].w * rstd * gv[i].w + bv[i].w);
    if (lane == 0) { mean_o[r] = mean; rstd_o[r] = rstd; }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float* __restrict__ mean_i,
                                                         const float* __restrict__ rstd_i, const float4* __restrict__ g, const float4* add,
                                                         float4* dx, float* __restrict__ dg, float* __restrict__ db, long long rows, int relu_x) {
  constexpr int D4 = NV * 32;
  __shared__ float sg[D4 * 4], sb[D4 * 4];
  const int tid = threadIdx.x, lane = tid & 31;
  const float inv_d = 1.f / (D4 * 4);
  for (int c = tid; c < D4 * 4; c += 256) { sg[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
  float4 gv[NV], ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gv[i] = __ldg(g + i * 32 + lane);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long w0 = (static_cast<long long>(blockIdx.x) * blockDim.x + tid) >> 5, nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += nw) {
    const float mean = __ldg(mean_i + r), rstd = __ldg(rstd_i + r);
    float4 xh[NV], dg_[NV];
    float s1 = 0.f, s2 = 0.f;
    unsigned pm = 0u;  // x > 0 per held column (the ReLU mask when x is a ReLU output)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = __ldg(x + r * D4 + i * 32 + lane), dv = __ldg(dy + r * D4 + i * 32 + lane);
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      dg_[i] = make_float4(dv.x * gv[i].x, dv.y * gv[i].y, dv.z * gv[i].z, dv.w * gv[i].w);
      s1 += (dg_[i].x + dg_[i].y) + (dg_[i].z + dg_[i].w);
      s2 += (dg_[i].x * xh[i].x + dg_[i].y * xh[i].y) + (dg_[i].z * xh[i].z + dg_[i].w * xh[i].w);
      ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
      ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
      pm |= ((xv.x > 0.f ? 1u : 0u) | (xv.y > 0.f ? 2u : 0u) | (xv.z > 0.f ? 4u : 0u) | (xv.w > 0.f ? 8u : 0u)) << (4 * i);
    }
    if (dx != nullptr) {
      const float c1 = wsum(s1) * inv_d, c2 = wsum(s2) * inv_d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 o = make_float4(rstd * (dg_[i].x - c1 - xh[i].x * c2), rstd * (dg_[i].y - c1 - xh[i].y * c2),
                               rstd * (dg_[i].z - c1 - xh[i].z * c2), rstd * (dg_[i].w - c1 - xh[i].w * c2));
        if (relu_x) {
          const unsigned b4 = pm >> (4 * i);
          o.x = (b4 & 1u) ? o.x : 0.f; o.y = (b4 & 2u) ? o.y : 0.f; o.z = (b4 & 4u) ? o.z : 0.f; o.w = (b4 & 8u) ? o.w : 0.f;
        }
        if (add != nullptr) {
          const float4 a = add[r * D4 + i * 32 + lane];
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        dx[r * D4 + i * 32 + lane] = o;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    atomicAdd(sg + c, ag[i].x); atomicAdd(sg + c + 1, ag[i].y); atomicAdd(sg + c + 2, ag[i].z); atomicAdd(sg + c + 3, ag[i].w);
    atomicAdd(sb + c, ab[i].x); atomicAdd(sb + c + 1, ab[i].y); atomicAdd(sb + c + 2, ab[i].z); atomicAdd(sb + c + 3, ab[i].w);
  }
  __syncthreads();
  for (int c = tid; c < D4 * 4; c += 256) { atomicAdd(dg + c, sg[c]); atomicAdd(db + c, sb[c]); }
}

// ---- any d <= 768 (300): scalar columns, PL per lane
template <int PL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                     float* __restrict__ y, float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                     long long rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += nw) {
    const float* xr = x + r * d;
    float v[PL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < d ? __ldg(xr + c) : 0.f;
      s += v[i];
    }
    const float mean = wsum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int c = lane + 32 * i;
      const float e = c < d ? v[i] - mean : 0.f;
      q += e * e;
    }
    const float rstd = 1.f / sqrtf(wsum(q) / d + 1e-5f);
    float* yr = y + r * d;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int c = lane + 32 * i;
      if (c < d) yr[c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
    if (lane == 0) { mean_o[r] = mean; rstd_o[r] = rstd; }
  }
}

// dx = [add +] rstd*(dy*g - mean(dy*g) - xhat*mean(dy*g*xhat)) [* (x > 0)], dg += sum dy*xhat, db += sum dy.
// relu_x: x is the output of a ReLU whose input gradient is wanted (the mask is x > 0).
template <int PL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_i,
                                                     const float* __restrict__ rstd_i, const float* __restrict__ g, const float* add,
                                                     float* dx, float* __restrict__ dg, float* __restrict__ db, long long rows, int d,
                                                     int relu_x) {
  __shared__ float sg[PL * 32], sb[PL * 32];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int c = tid; c < d; c += 256) { sg[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
  float ag[PL], ab[PL], gv[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    ag[i] = 0.f; ab[i] = 0.f;
    const int c = lane + 32 * i;
    gv[i] = c < d ? __ldg(g + c) : 0.f;
  }
  const long long w0 = (static_cast<long long>(blockIdx.x) * blockDim.x + tid) >> 5, nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += nw) {
    const float mean = __ldg(mean_i + r), rstd = __ldg(rstd_i + r);
    float xh[PL], dg_[PL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int c = lane + 32 * i;
      const float xv = c < d ? __ldg(x + r * d + c) : mean, dv = c < d ? __ldg(dy + r * d + c) : 0.f;
      xh[i] = (xv - mean) * rstd;
      dg_[i] = dv * gv[i];
      s1 += dg_[i];
      s2 += dg_[i] * xh[i];
      ag[i] += dv * xh[i];
      ab[i] += dv;
    }
    if (dx != nullptr) {
      const float c1 = wsum(s1) / d, c2 = wsum(s2) / d;
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        const int c = lane + 32 * i;
        if (c < d) {
          float o = rstd * (dg_[i] - c1 - xh[i] * c2);
          if (relu_x && !(__ldg(x + r * d + c) > 0.f)) o = 0.f;
          if (add != nullptr) o += add[r * d + c];
          dx[r * d + c] = o;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int c = lane + 32 * i;
    if (c < d) { atomicAdd(sg + c, ag[i]); atomicAdd(sb + c, ab[i]); }
  }
  __syncthreads();
  for (int c = tid; c < d; c += 256) { atomicAdd(dg + c, sg[c]); atomicAdd(db + c, sb[c]); }
}

// out[r] = max_c |x[r, c]|: one warp per row
__global__ void __launch_bounds__(256) row_absmax_kernel(const float* __restrict__ x, long long rows, int d, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    float m = 0.f;
    for (int c = lane; c < d; c += 32) m = fmaxf(m, fabsf(__ldg(x + r * d + c)));
    m = wmax(m);
    if (lane == 0) out[r] = m;
  }
}

// out[c] += sum over rows of x[r, c] (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int d, long long ld, float* __restrict__ out) {
  __shared__ float sb[768];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int c = tid; c < d; c += 256) sb[c] = 0.f;
  __syncthreads();
  float ab[LN_MAXPL];
#pragma unroll
  for (int i = 0; i < LN_MAXPL; ++i) ab[i] = 0.f;
  const long long w0 = (static_cast<long long>(blockIdx.x) * blockDim.x + tid) >> 5, nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += nw) {
#pragma unroll
    for (int i = 0; i < LN_MAXPL; ++i) {
      const int c = lane + 32 * i;
      if (c < d) ab[i] += __ldg(x + r * ld + c);
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAXPL; ++i) {
    const int c = lane + 32 * i;
    if (c < d) atomicAdd(sb + c, ab[i]);
  }
  __syncthreads();
  for (int c = tid; c < d; c += 256) atomicAdd(out + c, sb[c]);
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------------------------
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// y = a + b  (b broadcast with period `period` elements when period > 0: the position table)
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n, long long period) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = a[i] + b[period > 0 ? i % period : i];
}
// x *= *scalar (a device-resident scalar: the upstream gradient of a loss)
__global__ void scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ scalar, float mult) {
  const float s = __ldg(scalar) * mult;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    x[i] *= s;
}
// inverted dropout with a counter-based generator: element i of site `site` in call `seed[1]` of a run seeded `seed[0]` is
// kept when hash(seed, site, i) >= p * 2^32 and scaled by 1 / (1 - p).  The backward applies the same kernel to the gradient
// (no mask is stored); the seed lives on the device so that the launch can be replayed from a CUDA graph.
// one 64-bit hash serves four consecutive elements (16 random bits each: p is resolved to 1 / 65536)
__global__ void __launch_bounds__(256) dropout_kernel(const float* x, float* y, long long n, float p, const int* __restrict__ seed, unsigned site, int vec) {
  const unsigned long long key = (static_cast<unsigned long long>(static_cast<unsigned>(__ldg(seed))) << 32 | static_cast<unsigned>(__ldg(seed + 1))) *
                                     0xD6E8FEB86659FD93ull + static_cast<unsigned long long>(site) * 0x9E3779B97F4A7C15ull;
  const unsigned thr = static_cast<unsigned>(fminf(p, 0.99998f) * 65536.f);
  const float sc = 1.f / (1.f - p);
  const long long groups = (n + 3) >> 2;
  for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += static_cast<long long>(gridDim.x) * blockDim.x) {
    unsigned long long z = key + static_cast<unsigned long long>(g + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const unsigned lo = static_cast<unsigned>(z), hi = static_cast<unsigned>(z >> 32);
    const float k0 = (lo & 0xFFFFu) >= thr ? sc : 0.f, k1 = (lo >> 16) >= thr ? sc : 0.f;
    const float k2 = (hi & 0xFFFFu) >= thr ? sc : 0.f, k3 = (hi >> 16) >= thr ? sc : 0.f;
    if (vec && 4 * g + 4 <= n) {
      const float4 v = reinterpret_cast<const float4*>(x)[g];
      reinterpret_cast<float4*>(y)[g] = make_float4(v.x * k0, v.y * k1, v.z * k2, v.w * k3);
    } else {
      const float kk[4] = {k0, k1, k2, k3};
      for (int e = 0; e < 4 && 4 * g + e < n; ++e) y[4 * g + e] = x[4 * g + e] * kk[e];
    }
  }
}

// (g, I, T, d) -> (g, T, I, d)
__global__ void permute_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long groups, int I, int T, int d4) {
  const long long total = groups * I * T * d4;
  for (long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(o % d4);
    long long r = o / d4;
    const int i = static_cast<int>(r % I);
    r /= I;
    const int t = static_cast<int>(r % T);
    const long long gq = r / T;
    y[o] = __ldg(x + ((gq * I + i) * T + t) * d4 + c);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// depthwise FIR along the sequence (cross-correlation, zero padding k/2; model/cnn.py:25-28), d = 128 channels
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float4* __restrict__ y, long long n_seq, int L, int k) {
  const int c4 = threadIdx.x & 31, half = k / 2;
  float wv[4][7];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 7; ++j) wv[q][j] = j < k ? __ldg(w + (c4 * 4 + q) * k + j) : 0.f;
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  const long long rows = n_seq * L;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int l = static_cast<int>(r % L);
    float4 a = bv;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int ll = l + j - half;
      if (j < k && ll >= 0 && ll < L) {
        const float4 v = __ldg(x + (r + j - half) * 32 + c4);
        a.x = fmaf(wv[0][j], v.x, a.x); a.y = fmaf(wv[1][j], v.y, a.y); a.z = fmaf(wv[2][j], v.z, a.z); a.w = fmaf(wv[3][j], v.w, a.w);
      }
    }
    y[r * 32 + c4] = a;
  }
}
// dx[l] = sum_j w[j] dy[l - j + k/2]; dw[j] += sum_l dy[l] x[l + j - k/2]; db += sum dy
__global__ void __launch_bounds__(256) dwconv_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float* __restrict__ w,
                                                         float4* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                         long long n_seq, int L, int k) {
  __shared__ float sw[128 * 7], sbi[128];
  const int tid = threadIdx.x, c4 = tid & 31, half = k / 2;
  for (int i = tid; i < 128 * 7; i += 256) sw[i] = 0.f;
  if (tid < 128) sbi[tid] = 0.f;
  __syncthreads();
  float wv[4][7], aw[4][7], ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 7; ++j) { wv[q][j] = j < k ? __ldg(w + (c4 * 4 + q) * k + j) : 0.f; aw[q][j] = 0.f; }
  const long long rows = n_seq * L;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + tid) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int l = static_cast<int>(r % L);
    const float4 g = __ldg(dy + r * 32 + c4);
    ab[0] += g.x; ab[1] += g.y; ab[2] += g.z; ab[3] += g.w;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j >= k) continue;
      const int lx = l + j - half, ld_ = l - j + half;
      if (lx >= 0 && lx < L) {
        const float4 v = __ldg(x + (r + j - half) * 32 + c4);
        aw[0][j] = fmaf(g.x, v.x, aw[0][j]); aw[1][j] = fmaf(g.y, v.y, aw[1][j]); aw[2][j] = fmaf(g.z, v.z, aw[2][j]); aw[3][j] = fmaf(g.w, v.w, aw[3][j]);
      }
      if (ld_ >= 0 && ld_ < L) {
        const float4 v = __ldg(dy + (r - j + half) * 32 + c4);
        a.x = fmaf(wv[0][j], v.x, a.x); a.y = fmaf(wv[1][j], v.y, a.y); a.z = fmaf(wv[2][j], v.z, a.z); a.w = fmaf(wv[3][j], v.w, a.w);
      }
    }
    dx[r * 32 + c4] = a;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    atomicAdd(sbi + c4 * 4 + q, ab[q]);
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < k) atomicAdd(sw + (c4 * 4 + q) * 7 + j, aw[q][j]);
  }
  __syncthreads();
  for (int i = tid; i < 128 * k; i += 256) atomicAdd(dw + i, sw[(i / k) * 7 + i % k]);
  if (tid < 128) atomicAdd(db + tid, sbi[tid]);
}

// ------------------------------------------------------------------------------------------------------------------
// L2 normalisation of rows (F.normalize, eps 1e-12; d % 4 == 0): y = x * inv, inv = 1 / max(|x|, eps)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, float* __restrict__ inv_o, long long rows, int d4) {
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    float q = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 v = __ldg(x + r * d4 + c);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    const float inv = 1.f / fmaxf(sqrtf(wsum(q)), 1e-12f);
    for (int c = lane; c < d4; c += 32) {
      const float4 v = __ldg(x + r * d4 + c);
      y[r * d4 + c] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
    if (lane == 0 && inv_o != nullptr) inv_o[r] = inv;
  }
}
// dx (+)= inv * (dy - y (y . dy))
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, const float* __restrict__ inv_i,
                                                         float4* dx, long long rows, int d4, int accumulate) {
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    float dot = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 g = __ldg(dy + r * d4 + c), v = __ldg(y + r * d4 + c);
      dot += (g.x * v.x + g.y * v.y) + (g.z * v.z + g.w * v.w);
    }
    dot = wsum(dot);
    const float inv = __ldg(inv_i + r);
    for (int c = lane; c < d4; c += 32) {
      const float4 g = __ldg(dy + r * d4 + c), v = __ldg(y + r * d4 + c);
      float4 o = make_float4(inv * (g.x - v.x * dot), inv * (g.y - v.y * dot), inv * (g.z - v.z * dot), inv * (g.w - v.w * dot));
      if (accumulate) { const float4 p = dx[r * d4 + c]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
      dx[r * d4 + c] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// masked softmax of the cosine scores (context_query_attention.py:59-66,100): row (n, a, i, t), columns j < lc.
// s (in place): raw cosine -> raw - 1e10 (1 - cm*qm); p = softmax(scale * s) * cm*qm.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(float* __restrict__ s, const float* __restrict__ c_mask, const float* __restrict__ q_mask,
                                                               float* __restrict__ p, long long rows, int T, long long ct_rows, int lc, float scale) {
  // c_mask row = r / T; q_mask row = (n * T + t) with n = r / ct_rows (ct_rows = 5 * Lq * T)
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const float cm = __ldg(c_mask + r / T);
    const long long qrow = (r / ct_rows) * T + r % T;
    float z[4], m[4];
    float mx = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      if (j < lc) {
        m[i] = cm * __ldg(q_mask + qrow * lc + j);
        const float v = s[r * lc + j] - 1e10f * (1.f - m[i]);
        s[r * lc + j] = v;
        z[i] = v * scale;
        mx = fmaxf(mx, z[i]);
      } else { m[i] = 0.f; z[i] = -3.0e38f; }
    }
    mx = wmax(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      z[i] = j < lc ? expf(z[i] - mx) : 0.f;
      sum += z[i];
    }
    sum = wsum(sum);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      if (j < lc) p[r * lc + j] = z[i] / sum * m[i];
    }
  }
}
// ds = scale * p * (dp - sum_j dp*p), written over dp
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(float* __restrict__ dp, const float* __restrict__ p, long long rows, int lc, float scale) {
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    float g[4], pv[4];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      g[i] = j < lc ? dp[r * lc + j] : 0.f;
      pv[i] = j < lc ? __ldg(p + r * lc + j) : 0.f;
      dot += g[i] * pv[i];
    }
    dot = wsum(dot);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = lane + 32 * i;
      if (j < lc) dp[r * lc + j] = scale * pv[i] * (g[i] - dot);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// out = [a, b, a*b] (128 + 128 + 128); a's row is r / a_div (a_div = T: the answer row repeated over frames, stage.py:374)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cat3_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                                                       long long rows, int a_div) {
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const float4 u = __ldg(a + (r / a_div) * 32 + lane), v = __ldg(b + r * 32 + lane);
    out[r * 96 + lane] = u;
    out[r * 96 + 32 + lane] = v;
    out[r * 96 + 64 + lane] = make_float4(u.x * v.x, u.y * v.y, u.z * v.z, u.w * v.w);
  }
}
// one block (128 threads, one channel each) per row of a: da (+)= sum over its a_div rows of d1 + d3*b; db = d2 + d3*a
__global__ void __launch_bounds__(128) cat3_bwd_kernel(const float* __restrict__ dcat, const float* __restrict__ a, const float* __restrict__ b,
                                                       float* da, float* __restrict__ dbo, int a_div, int accumulate_a) {
  const int c = threadIdx.x;
  const long long ar = blockIdx.x;
  const float av = __ldg(a + ar * 128 + c);
  float acc = 0.f;
  for (int t = 0; t < a_div; ++t) {
    const long long r = ar * a_div + t;
    const float d1 = __ldg(dcat + r * 384 + c), d2 = __ldg(dcat + r * 384 + 128 + c), d3 = __ldg(dcat + r * 384 + 256 + c);
    const float bv = __ldg(b + r * 128 + c);
    acc += d1 + d3 * bv;
    dbo[r * 128 + c] = d2 + d3 * av;
  }
  da[ar * 128 + c] = accumulate_a ? da[ar * 128 + c] + acc : acc;
}

// ------------------------------------------------------------------------------------------------------------------
// masked max over the rows of a sequence (stage.py:503 and mask_logits, model_utils.py:14-15), with the argmax
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) maxpool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y,
                                                          int* __restrict__ arg, int L) {
  const int c = threadIdx.x;
  const long long s = blockIdx.x;
  float best = -3.0e38f;
  int bi = 0;
  for (int i = 0; i < L; ++i) {
    const float m = __ldg(mask + s * L + i);
    const float v = __ldg(x + (s * L + i) * 128 + c) * m + (1.f - m) * TRN_NEG;
    if (v > best) { best = v; bi = i; }
  }
  y[s * 128 + c] = best;
  arg[s * 128 + c] = bi;
}
__global__ void __launch_bounds__(128) maxpool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ mask, const int* __restrict__ arg,
                                                          float* __restrict__ dx, int L) {
  const int c = threadIdx.x;
  const long long s = blockIdx.x;
  const int a = __ldg(arg + s * 128 + c);
  const float g = __ldg(dy + s * 128 + c);
  for (int i = 0; i < L; ++i) dx[(s * L + i) * 128 + c] = i == a ? g * __ldg(mask + s * L + i) : 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
// head: temporal cross-entropy (stage.py:539-555), span proposals and their pooled features (stage.py:407-438)
// ------------------------------------------------------------------------------------------------------------------
// one block per example: rows of the GT answer; loss terms, softmax probabilities (for the span search) and the gradient of
// 0.5*(CE_st + CE_ed) w.r.t. the MASKED scores, already multiplied by the frame mask (mask_logits backward).
__global__ void __launch_bounds__(256) ts_loss_kernel(const float* __restrict__ t_scores, const float* __restrict__ ts_mask,
                                                      const long long* __restrict__ target, const long long* __restrict__ st,
                                                      const long long* __restrict__ ed, int T, float* __restrict__ terms,
                                                      float* __restrict__ probs, float* __restrict__ dscores) {
  __shared__ float red[8];
  const int tid = threadIdx.x, n = blockIdx.x;
  const long long row = static_cast<long long>(n) * 5 + __ldg(target + n);
  const float* sc = t_scores + row * T * 2;
  for (int c = 0; c < 2; ++c) {
    float mx = -3.0e38f;
    for (int t = tid; t < T; t += 256) mx = fmaxf(mx, sc[2 * t + c]);
    mx = wmax(mx);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    float s = 0.f;
    for (int t = tid; t < T; t += 256) s += expf(sc[2 * t + c] - mx);
    s = wsum(s);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    s = red[0];
    for (int i = 1; i < 8; ++i) s += red[i];
    const int lab = static_cast<int>(__ldg((c == 0 ? st : ed) + n));
    if (tid == 0) terms[2 * n + c] = (mx + logf(s)) - sc[2 * lab + c];
    for (int t = tid; t < T; t += 256) {
      const float pr = expf(sc[2 * t + c] - mx) / s;
      probs[(static_cast<long long>(c) * gridDim.x + n) * T + t] = pr;
      dscores[(row * T + t) * 2 + c] = 0.5f * (pr - (t == lab ? 1.f : 0.f)) * __ldg(ts_mask + static_cast<long long>(n) * T + t);
    }
  }
}

// proposals of example n: kind 0 = the GT span, kind 1 = the predicted span if confident and IoU >= 0.5 (stage.py:407-438).
// feat[r, a, 0:128] = max over the (expanded) span's frames, [128:256] = max over all frames, of mask_logits(x, frame mask).
__global__ void __launch_bounds__(128) proposal_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ fmask,
                                                                const long long* __restrict__ target, const long long* __restrict__ st,
                                                                const long long* __restrict__ ed, const int* __restrict__ spans,
                                                                const float* __restrict__ conf, int T, int extra, float prob_thd,
                                                                float* __restrict__ feat, int* __restrict__ arg, int* __restrict__ valid,
                                                                long long* __restrict__ row_target) {
  __shared__ int sh[3];
  const int c = threadIdx.x, r = blockIdx.x, n = r >> 1, kind = r & 1;
  if (c == 0) {
    const int g0 = static_cast<int>(__ldg(st + n)), g1 = static_cast<int>(__ldg(ed + n)) + 1;
    int s0 = g0, s1 = g1, ok = 1;
    if (kind == 1) {
      s0 = spans[2 * n];
      s1 = spans[2 * n + 1] + 1;
      const int inter = max(0, min(s1, g1) - max(s0, g0)), uni = max(s1, g1) - min(s0, g0);
      ok = (static_cast<double>(conf[n]) >= static_cast<double>(prob_thd)) && uni != 0 && 2 * inter >= uni;
    }
    sh[0] = max(0, s0 - extra);
    sh[1] = min(T, s1 + extra);
    sh[2] = ok;
    valid[r] = ok;
    row_target[r] = __ldg(target + n);
  }
  __syncthreads();
  const int lo = sh[0], hi = sh[1];
  for (int a = 0; a < 5; ++a) {
    const long long row = static_cast<long long>(n) * 5 + a;
    float bl = -3.0e38f, bg = -3.0e38f;
    int il = lo, ig = 0;
    for (int t = 0; t < T; ++t) {
      const float m = __ldg(fmask + row * T + t);
      const float v = __ldg(x + (row * T + t) * 128 + c) * m + (1.f - m) * TRN_NEG;
      if (v > bg) { bg = v; ig = t; }
      if (t >= lo && t < hi && v > bl) { bl = v; il = t; }
    }
    const long long o = (static_cast<long long>(r) * 5 + a) * 256;
    feat[o + c] = bl; feat[o + 128 + c] = bg;
    arg[o + c] = il; arg[o + 128 + c] = ig;
  }
}
// dx[(n, a, arg), c] += dfeat * frame mask, valid proposals only
__global__ void __launch_bounds__(256) proposal_pool_bwd_kernel(const float* __restrict__ dfeat, const int* __restrict__ arg,
                                                                const int* __restrict__ valid, const float* __restrict__ fmask, int T,
                                                                float* __restrict__ dx) {
  const int j = threadIdx.x, r = blockIdx.x / 5, a = blockIdx.x % 5, n = r >> 1;
  if (!valid[r]) return;
  const long long o = (static_cast<long long>(r) * 5 + a) * 256 + j, row = static_cast<long long>(n) * 5 + a;
  const int t = arg[o];
  atomicAdd(dx + (row * T + t) * 128 + (j & 127), dfeat[o] * __ldg(fmask + row * T + t));
}

// gradient of the ranking loss (stage.py:741-746) scattered into the score tensor: for lse, d/dneg = +sig(alpha*d)*alpha,
// d/dpos = the negative; for the hinge, +-1 where margin + d > 0.  `up` = device scalar upstream gradient.
__global__ void att_loss_bwd_kernel(const float* __restrict__ scores, const long long* __restrict__ pos_off, const long long* __restrict__ neg_off,
                                    long long m, int hinge, float alpha, float margin, const float* __restrict__ up, float* __restrict__ dscores) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long po = __ldg(pos_off + i), no = __ldg(neg_off + i);
  const float d = __ldg(scores + no) - __ldg(scores + po);
  float g;
  if (hinge) g = (margin + d > 0.f) ? 1.f : 0.f;
  else g = alpha / (1.f + expf(-alpha * d));
  g *= __ldg(up);
  atomicAdd(dscores + no, g);
  atomicAdd(dscores + po, -g);
}

// cls loss of main.py:57-59 on the proposal rows: sum over valid rows of CE(answer_row, target) * (n / n_valid), its gradient,
// and the number of correct valid rows.  One block.  out[0] = loss, out[1] = n_valid, out[2] = n_correct.
__global__ void __launch_bounds__(256) answer_ce_kernel(const float* __restrict__ rows, const int* __restrict__ valid,
                                                        const long long* __restrict__ target, int n_rows, int n_examples,
                                                        float* __restrict__ out, float* __restrict__ drows) {
  __shared__ float red[3][8];
  const int tid = threadIdx.x;
  float loss = 0.f, cnt = 0.f, ok = 0.f;
  for (int r = tid; r < n_rows; r += 256) {
    if (!valid[r]) continue;
    float v[5], mx = -3.0e38f;
    int am = 0;
    for (int a = 0; a < 5; ++a) { v[a] = rows[r * 5 + a]; if (v[a] > mx) { mx = v[a]; am = a; } }
    float s = 0.f;
    for (int a = 0; a < 5; ++a) s += expf(v[a] - mx);
    const int t = static_cast<int>(target[r]);
    loss += (mx + logf(s)) - v[t];
    cnt += 1.f;
    ok += am == t ? 1.f : 0.f;
  }
  loss = wsum(loss); cnt = wsum(cnt); ok = wsum(ok);
  if ((tid & 31) == 0) { red[0][tid >> 5] = loss; red[1][tid >> 5] = cnt; red[2][tid >> 5] = ok; }
  __syncthreads();
  float tl = 0.f, tc = 0.f, to = 0.f;
  for (int i = 0; i < 8; ++i) { tl += red[0][i]; tc += red[1][i]; to += red[2][i]; }
  const float w = tc > 0.f ? static_cast<float>(n_examples) / tc : 0.f;
  if (tid == 0) { out[0] = tl * w; out[1] = tc; out[2] = to; }
  for (int r = tid; r < n_rows; r += 256) {
    float v[5], mx = -3.0e38f;
    for (int a = 0; a < 5; ++a) { v[a] = rows[r * 5 + a]; mx = fmaxf(mx, v[a]); }
    float s = 0.f;
    for (int a = 0; a < 5; ++a) { v[a] = expf(v[a] - mx); s += v[a]; }
    const int t = static_cast<int>(target[r]);
    for (int a = 0; a < 5; ++a) drows[r * 5 + a] = valid[r] ? w * (v[a] / s - (a == t ? 1.f : 0.f)) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// optimiser: global gradient norm, then clip + Adam with L2 weight decay on the flat buffers (main.py:63-66,208-212)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) s += g[i] * g[i];
  s = wsum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}
__global__ void counter_inc_kernel(int* c) { *c += 1; }
// clip coefficient = min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_); torch.optim.Adam with weight_decay
// (grad += wd * p), bias corrections passed in as bc1 = 1 - beta1^t, bc2 = 1 - beta2^t.
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            const float* __restrict__ sumsq, float max_norm, float lr, float beta1, float beta2, float eps, float wd,
                            float bc1, float bc2, float grad_scale, const int* __restrict__ step_dev) {
  if (step_dev != nullptr) {  // step count kept on the device: the launch can be replayed from a CUDA graph
    const float t = static_cast<float>(__ldg(step_dev));
    bc1 = 1.f - powf(beta1, t);
    bc2 = 1.f - powf(beta2, t);
  }
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float nrm = sqrtf(__ldg(sumsq)) * grad_scale;
    coef *= fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  const float step = lr / bc1, rs2 = 1.f / sqrtf(bc2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gc = g[i] * coef;
    g[i] = gc;
    const float gg = gc + wd * p[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gg, vi = beta2 * v[i] + (1.f - beta2) * gg * gg;
    m[i] = mi; v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * rs2 + eps);
  }
}

}  // namespace trn
}  // namespace stg
