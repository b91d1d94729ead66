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

__global__ void pack_weight_kernel(const float* __restrict__ w, PackGeom g, uint8_t* __restrict__ out) {
  // one thread per (k8 slab over the whole padded K, row n)
  const int total = (g.kpad / 8) * g.npad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % g.npad, k8 = idx / g.npad;
    const int k0 = k8 * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (n < g.n && k0 + j < g.k) ? w[static_cast<size_t>(n) * g.k + k0 + j] : 0.f;
    HL8 hl = split8(v);
    const int c = k0 / g.kc, s = (k0 % g.kc) / 8;
    uint8_t* base = out + static_cast<size_t>(c) * g.chunk_bytes + static_cast<size_t>(s) * g.lbo + static_cast<size_t>(n) * 16;
    *reinterpret_cast<uint4*>(base) = hl.hi;
    *reinterpret_cast<uint4*>(base + g.part_bytes) = hl.lo;
  }
}

}  // namespace stg
