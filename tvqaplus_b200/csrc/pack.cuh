// Weight packing: W[N,K] fp32 (nn.Linear / 1x1 conv layout) -> split-fp16 operand image streamed by the GEMM kernels.
//
// Image = n_chunks blocks of `chunk_bytes`; block c holds K columns [c*KC, (c+1)*KC) as
//     [hi part | lo part], each part = (KC/8) k8-slabs of LBO_B = NPAD*16 bytes; inside a slab row n owns 16 bytes
//     (8 consecutive k).  NPAD is the consuming kernel's N tile (128 for N <= 128, 304 for N <= 304); rows >= N and
//     columns >= K are zero.
// A block is exactly the shared-memory image of one pipeline stage, so a stage load is ONE cp.async.bulk.
#pragma once
#include "umma.cuh"

namespace stg {

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct PackGeom {
  int n, k, kc, npad, kpad, n_chunks;
  uint32_t lbo, part_bytes, chunk_bytes;
};
__host__ __device__ inline PackGeom pack_geom(int n, int k, int kc) {
  PackGeom g;
  g.n = n; g.k = k; g.kc = kc;
  g.npad = n <= 128 ? 128 : 304;
  g.kpad = round_up(k, kc);
  g.n_chunks = g.kpad / kc;
  g.lbo = static_cast<uint32_t>(g.npad) * 16u;
  g.part_bytes = static_cast<uint32_t>(kc / 8) * g.lbo;
  g.chunk_bytes = 2u * g.part_bytes;
  return g;
}

// w: row-major [n, ld]; packs columns [k0, k0 + g.k) optionally scaled per column by col_scale[k0 + k] (LayerNorm gamma
// folded into the weight, see cqa.cuh).  perm128 = 1 reorders the columns inside every 128-column segment to the order
// in which the fused CQA kernel stages its A operand: packed position h*64 + q*16 + j  <-  source column q*32 + h*16 + j.
__global__ void pack_weight_kernel(const float* __restrict__ w, int ld, int k0, const float* __restrict__ col_scale,
                                   PackGeom g, uint8_t* __restrict__ out, int perm128 = 0) {
  // one thread per (k8 slab over the whole padded K, row n)
  const int total = (g.kpad / 8) * g.npad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % g.npad, k8 = idx / g.npad;
    const int kk = k8 * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int src = kk + j;
      if (perm128) {
        const int pos = src & 127;
        src = (src & ~127) + ((pos & 63) >> 4) * 32 + (pos >> 6) * 16 + (pos & 15);
      }
      float x = (n < g.n && src < g.k) ? w[static_cast<size_t>(n) * ld + k0 + src] : 0.f;
      if (col_scale != nullptr && src < g.k) x *= col_scale[k0 + src];
      v[j] = x;
    }
    HL8 hl = split8(v);
    const int c = kk / g.kc, s = (kk % g.kc) / 8;
    uint8_t* base = out + static_cast<size_t>(c) * g.chunk_bytes + static_cast<size_t>(s) * g.lbo + static_cast<size_t>(n) * 16;
    *reinterpret_cast<uint4*>(base) = hl.hi;
    *reinterpret_cast<uint4*>(base + g.part_bytes) = hl.lo;
  }
}

// One half (h = 0 / 1) of a 128-column segment [k0, k0 + 128) in the fused CQA kernel's K order: packed position
// q*16 + j of the half  <-  source column q*32 + h*16 + j (q = 0..3).  g = pack_geom(n, 64, 32): two K=32 chunks.
__global__ void pack_weight_half_kernel(const float* __restrict__ w, int ld, int k0, const float* __restrict__ col_scale,
                                        PackGeom g, int h, uint8_t* __restrict__ out) {
  const int total = (g.kpad / 8) * g.npad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % g.npad, k8 = idx / g.npad;
    const int kk = k8 * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pos = kk + j;  // 0..63 inside the half
      const int src = (pos >> 4) * 32 + h * 16 + (pos & 15);
      float x = n < g.n ? w[static_cast<size_t>(n) * ld + k0 + src] : 0.f;
      if (col_scale != nullptr) x *= col_scale[k0 + src];
      v[j] = x;
    }
    HL8 hl = split8(v);
    const int c = kk / g.kc, s = (kk % g.kc) / 8;
    uint8_t* base = out + static_cast<size_t>(c) * g.chunk_bytes + static_cast<size_t>(s) * g.lbo + static_cast<size_t>(n) * 16;
    *reinterpret_cast<uint4*>(base) = hl.hi;
    *reinterpret_cast<uint4*>(base + g.part_bytes) = hl.lo;
  }
}

// LayerNorm folded into the Linear that follows it:  LN(x) W^T + b = rstd * (x (W*gamma)^T - mean * cs) + cb  with
//   cs[n] = sum_k W[n,k] gamma[k],  cb[n] = sum_k W[n,k] beta[k] + b[n].          One warp per output feature n.
__global__ void fold_layernorm_kernel(const float* __restrict__ w, int n, int k, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                      float* __restrict__ cs, float* __restrict__ cb) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < k; i += 32) {
    const float x = w[static_cast<size_t>(row) * k + i];
    a = fmaf(x, gamma[i], a);
    b = fmaf(x, beta[i], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    cs[row] = a;
    cb[row] = b + bias[row];
  }
}

}  // namespace stg
