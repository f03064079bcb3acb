// reduce.cuh — deterministic reduction of small fp32 partial tables (shared by the backward finalisers).
#pragma once
#include "common.cuh"

namespace dsg {

// Fixed-order (deterministic) reduction of small fp32 partial tables, shared by the d gamma / d beta finaliser and the
// column-sum finaliser.  Element (sample i, part k, channel c, component q) lives at
//   src[i * sample_stride + k * part_stride + c * COMPS + q].
// Block = 32 channels x 8 slices; a slice owns the parts k = slice, slice + 8, ...; all loads of a round are
// independent (the naive one-thread-per-channel loop was a chain of n * parts dependent L2 latencies).
template <int COMPS>
__device__ __forceinline__ void reduce_rows_body(const float* __restrict__ src, int n, int parts, int C,
                                                 int64_t sample_stride, int64_t part_stride, float* __restrict__ per_n,
                                                 int per_n_stride, int per_n_off, const float* __restrict__ inv_scale,
                                                 float* __restrict__ out0, float* __restrict__ out0b,
                                                 float* __restrict__ out1, int block_x, float* smem_f) {
  constexpr int S = 8, TS = 8;  // slices, samples per round
  float (*s1)[S][32][COMPS] = reinterpret_cast<float (*)[S][32][COMPS]>(smem_f);
  float (*s2)[32][COMPS] = reinterpret_cast<float (*)[32][COMPS]>(smem_f + TS * S * 32 * COMPS);
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = block_x * 32 + lane;
  const bool ok = c < C;
  float tot[COMPS];
#pragma unroll
  for (int q = 0; q < COMPS; ++q) tot[q] = 0.f;
  for (int i0 = 0; i0 < n; i0 += TS) {
    float v[TS][COMPS];
#pragma unroll
    for (int s = 0; s < TS; ++s)
#pragma unroll
      for (int q = 0; q < COMPS; ++q) v[s][q] = 0.f;
    if (ok) {
      for (int k = slice; k < parts; k += S) {
#pragma unroll
        for (int s = 0; s < TS; ++s) {
          if (i0 + s < n) {
            const float* p = src + (int64_t)(i0 + s) * sample_stride + (int64_t)k * part_stride + (int64_t)c * COMPS;
#pragma unroll
            for (int q = 0; q < COMPS; ++q) v[s][q] += p[q];
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < TS; ++s)
#pragma unroll
      for (int q = 0; q < COMPS; ++q) s1[s][slice][lane][q] = v[s][q];
    __syncthreads();
    {  // thread (lane, slice) finishes sample i0 + slice: the 8 slices in a fixed order
      const int s = slice;
#pragma unroll
      for (int q = 0; q < COMPS; ++q) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) t += s1[s][j][lane][q];
        s2[s][lane][q] = t;
        if (q == 0 && per_n && ok && i0 + s < n) per_n[(int64_t)(i0 + s) * per_n_stride + per_n_off + c] = t;
      }
    }
    __syncthreads();
    if (slice == 0) {
#pragma unroll
      for (int s = 0; s < TS; ++s)
#pragma unroll
        for (int q = 0; q < COMPS; ++q) tot[q] += s2[s][lane][q];   // samples past n contributed zeros
    }
  }
  if (slice == 0 && ok) {
    const float sc = inv_scale ? *inv_scale : 1.0f;
    if (out0) out0[c] = tot[0] * sc;
    if (out0b) out0b[c] = tot[0] * sc;
    if (COMPS > 1 && out1) out1[c] = tot[COMPS - 1] * sc;
  }
}

constexpr int RR_SMEM_FLOATS = 2 * (8 * 8 * 32 + 8 * 32);   // s1 + s2 at COMPS = 2

template <int COMPS>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ src, int n, int parts, int C,
                                                          int64_t sample_stride, int64_t part_stride,
                                                          float* __restrict__ per_n, int per_n_stride, int per_n_off,
                                                          const float* __restrict__ inv_scale, float* __restrict__ out0,
                                                          float* __restrict__ out0b, float* __restrict__ out1) {
  __shared__ float smem_f[RR_SMEM_FLOATS];
  pdl_sync();
  reduce_rows_body<COMPS>(src, n, parts, C, sample_stride, part_stride, per_n, per_n_stride, per_n_off, inv_scale, out0,
                          out0b, out1, (int)blockIdx.x, smem_f);
}

}  // namespace dsg
