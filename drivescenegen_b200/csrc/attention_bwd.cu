// attention_bwd.cu — backward of the mid-block self-attention core (softmax(q k^T / sqrt(d)) v per head), head_dim 8.
// Replaces the autograd backward of F.scaled_dot_product_attention inside diffusers 0.20.0 AttnProcessor2_0
// (models/attention_processor.py), reached from `accelerator.backward(loss)`
// (DriveSceneGen/pipeline/training_pipeline.py:86).
//
// Flash-style recomputation on CUDA cores, fp32 math, three passes over the (query, key) pairs of one (sample, head):
//   prep : lse_i = logsumexp_j(s_ij),  delta_i = <dO_i, O_i>                     (thread = query row)
//   dkv  : p_ij = exp(s_ij - lse_i);  dV_j += p_ij dO_i;  dS_ij = p_ij (<dO_i, V_j> - delta_i) / sqrt(d);
//          dK_j += dS_ij Q_i                                                     (thread = key row)
//   dq   : dQ_i += dS_ij K_j                                                      (thread = query row)
// At head_dim 8 the contraction depth is far below a tensor-core tile; the op is 1-2 % of the backward FLOPs.
// Each thread owns one row and accumulates in registers in a fixed order: deterministic, no atomics.
#include "common.cuh"

namespace dsg {

constexpr int AB_D = 8;
constexpr int AB_ROWS = 128;   // rows per CTA = threads
constexpr int AB_TILE = 64;    // "other side" rows staged in shared memory per step

__device__ __forceinline__ void load_row8(const __half* p, float* f) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  unpack8(r, f);
}

// grid (ceil(T / 128), heads, n)
__global__ void __launch_bounds__(AB_ROWS) attn_bwd_prep_kernel(const __half* __restrict__ qkv,
                                                               const __half* __restrict__ o,
                                                               const __half* __restrict__ dout, float* __restrict__ lse,
                                                               float* __restrict__ delta, int T, int heads,
                                                               float scale_log2) {
  const int C = heads * AB_D;
  const int n = blockIdx.z, hd = blockIdx.y;
  const int i = blockIdx.x * AB_ROWS + threadIdx.x;
  const bool valid = i < T;
  const __half* base = qkv + (int64_t)n * T * 3 * C;
  __shared__ float sk[AB_TILE][AB_D];
  float q[AB_D];
#pragma unroll
  for (int d = 0; d < AB_D; ++d) q[d] = 0.f;
  if (valid) load_row8(base + (int64_t)i * 3 * C + hd * AB_D, q);
#pragma unroll
  for (int d = 0; d < AB_D; ++d) q[d] *= scale_log2;  // scores in log2 units
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < T; j0 += AB_TILE) {
    __syncthreads();
    if (threadIdx.x < AB_TILE) {
      float kf[AB_D];
#pragma unroll
      for (int d = 0; d < AB_D; ++d) kf[d] = 0.f;
      if (j0 + (int)threadIdx.x < T) load_row8(base + (int64_t)(j0 + threadIdx.x) * 3 * C + C + hd * AB_D, kf);
#pragma unroll
      for (int d = 0; d < AB_D; ++d) sk[threadIdx.x][d] = kf[d];
    }
    __syncthreads();
    const int jn = min(AB_TILE, T - j0);
    float s[AB_TILE];
    float tm = m;
#pragma unroll
    for (int j = 0; j < AB_TILE; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < AB_D; ++d) a = fmaf(q[d], sk[j][d], a);
      s[j] = j < jn ? a : -INFINITY;
      tm = fmaxf(tm, s[j]);
    }
    l *= exp2f(m - tm);
#pragma unroll
    for (int j = 0; j < AB_TILE; ++j) l += exp2f(s[j] - tm);
    m = tm;
  }
  if (valid) {
    const int64_t r = ((int64_t)n * heads + hd) * T + i;
    lse[r] = m + log2f(l);  // log2 units
    float of[AB_D], df[AB_D];
    load_row8(o + ((int64_t)n * T + i) * C + hd * AB_D, of);
    load_row8(dout + ((int64_t)n * T + i) * C + hd * AB_D, df);
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < AB_D; ++d) dl = fmaf(of[d], df[d], dl);
    delta[r] = dl;
  }
}

// thread = key row j; loops over all queries.  Writes dK and dV slices of dqkv.
__global__ void __launch_bounds__(AB_ROWS) attn_bwd_dkv_kernel(const __half* __restrict__ qkv,
                                                              const __half* __restrict__ dout,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ delta,
                                                              __half* __restrict__ dqkv, int T, int heads,
                                                              float scale, float scale_log2) {
  const int C = heads * AB_D;
  const int n = blockIdx.z, hd = blockIdx.y;
  const int j = blockIdx.x * AB_ROWS + threadIdx.x;
  const bool valid = j < T;
  const __half* base = qkv + (int64_t)n * T * 3 * C;
  __shared__ float sq[AB_TILE][AB_D], sdo[AB_TILE][AB_D], sl[AB_TILE], sdl[AB_TILE];
  float k[AB_D], v[AB_D], dk[AB_D], dv[AB_D];
#pragma unroll
  for (int d = 0; d < AB_D; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
  if (valid) {
    load_row8(base + (int64_t)j * 3 * C + C + hd * AB_D, k);
    load_row8(base + (int64_t)j * 3 * C + 2 * C + hd * AB_D, v);
  }
  float ks[AB_D];
#pragma unroll
  for (int d = 0; d < AB_D; ++d) ks[d] = k[d] * scale_log2;
  const int64_t rbase = ((int64_t)n * heads + hd) * T;
  for (int i0 = 0; i0 < T; i0 += AB_TILE) {
    __syncthreads();
    if (threadIdx.x < AB_TILE) {
      const int i = i0 + threadIdx.x;
      float qf[AB_D], df[AB_D];
#pragma unroll
      for (int d = 0; d < AB_D; ++d) { qf[d] = 0.f; df[d] = 0.f; }
      float li = INFINITY, di = 0.f;  // exp2(s - inf) = 0 for rows past the end
      if (i < T) {
        load_row8(base + (int64_t)i * 3 * C + hd * AB_D, qf);
        load_row8(dout + ((int64_t)n * T + i) * C + hd * AB_D, df);
        li = lse[rbase + i];
        di = delta[rbase + i];
      }
#pragma unroll
      for (int d = 0; d < AB_D; ++d) { sq[threadIdx.x][d] = qf[d]; sdo[threadIdx.x][d] = df[d]; }
      sl[threadIdx.x] = li;
      sdl[threadIdx.x] = di;
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < AB_TILE; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < AB_D; ++d) {
        s = fmaf(sq[i][d], ks[d], s);
        dp = fmaf(sdo[i][d], v[d], dp);
      }
      const float p = exp2f(s - sl[i]);
      const float ds = p * (dp - sdl[i]) * scale;
#pragma unroll
      for (int d = 0; d < AB_D; ++d) {
        dv[d] = fmaf(p, sdo[i][d], dv[d]);
        dk[d] = fmaf(ds, sq[i][d], dk[d]);
      }
    }
  }
  if (valid) {
    __half* ob = dqkv + ((int64_t)n * T + j) * 3 * C;
    stg_v4(ob + C + hd * AB_D, pack8(dk));
    stg_v4(ob + 2 * C + hd * AB_D, pack8(dv));
  }
}

// thread = query row i; loops over all keys.  Writes the dQ slice of dqkv.
__global__ void __launch_bounds__(AB_ROWS) attn_bwd_dq_kernel(const __half* __restrict__ qkv,
                                                             const __half* __restrict__ dout,
                                                             const float* __restrict__ lse,
                                                             const float* __restrict__ delta,
                                                             __half* __restrict__ dqkv, int T, int heads, float scale,
                                                             float scale_log2) {
  const int C = heads * AB_D;
  const int n = blockIdx.z, hd = blockIdx.y;
  const int i = blockIdx.x * AB_ROWS + threadIdx.x;
  const bool valid = i < T;
  const __half* base = qkv + (int64_t)n * T * 3 * C;
  __shared__ float sk[AB_TILE][AB_D], sv[AB_TILE][AB_D];
  float q[AB_D], df[AB_D], dq[AB_D];
#pragma unroll
  for (int d = 0; d < AB_D; ++d) { q[d] = 0.f; df[d] = 0.f; dq[d] = 0.f; }
  float li = 0.f, di = 0.f;
  const int64_t rbase = ((int64_t)n * heads + hd) * T;
  if (valid) {
    load_row8(base + (int64_t)i * 3 * C + hd * AB_D, q);
    load_row8(dout + ((int64_t)n * T + i) * C + hd * AB_D, df);
    li = lse[rbase + i];
    di = delta[rbase + i];
  }
  float qs[AB_D];
#pragma unroll
  for (int d = 0; d < AB_D; ++d) qs[d] = q[d] * scale_log2;
  for (int j0 = 0; j0 < T; j0 += AB_TILE) {
    __syncthreads();
    if (threadIdx.x < AB_TILE) {
      const int j = j0 + threadIdx.x;
      float kf[AB_D], vf[AB_D];
#pragma unroll
      for (int d = 0; d < AB_D; ++d) { kf[d] = 0.f; vf[d] = 0.f; }
      if (j < T) {
        load_row8(base + (int64_t)j * 3 * C + C + hd * AB_D, kf);
        load_row8(base + (int64_t)j * 3 * C + 2 * C + hd * AB_D, vf);
      }
#pragma unroll
      for (int d = 0; d < AB_D; ++d) { sk[threadIdx.x][d] = kf[d]; sv[threadIdx.x][d] = vf[d]; }
    }
    __syncthreads();
    const int jn = min(AB_TILE, T - j0);
#pragma unroll 4
    for (int j = 0; j < AB_TILE; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < AB_D; ++d) {
        s = fmaf(qs[d], sk[j][d], s);
        dp = fmaf(df[d], sv[j][d], dp);
      }
      const float p = j < jn ? exp2f(s - li) : 0.f;
      const float ds = p * (dp - di) * scale;
#pragma unroll
      for (int d = 0; d < AB_D; ++d) dq[d] = fmaf(ds, sk[j][d], dq[d]);
    }
  }
  if (valid) stg_v4(dqkv + ((int64_t)n * T + i) * 3 * C + hd * AB_D, pack8(dq));
}

}  // namespace dsg

using namespace dsg;

namespace dsg {
int launch_attention_bwd_tc(const __half* qkv, const __half* o, const __half* dout, const float* lse, __half* dqkv, int n,
                            int tokens, int heads, int head_dim, cudaStream_t st);  // attention_bwd_tc.cu
}

extern "C" int dsg_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, float* ws,
                                 const float* lse, int32_t n, int32_t tokens, int32_t heads, int32_t head_dim,
                                 void* stream) {
  DSG_CHECK_ARG(qkv && out && dout && dqkv && (ws || lse), "dsg_attention_bwd: null pointer");
  DSG_CHECK_ARG(head_dim == AB_D, "dsg_attention_bwd: head_dim %d not supported (only 8)", head_dim);
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && tokens > 0 && heads > 0 && heads <= 65535, "dsg_attention_bwd: bad sizes");
  DSG_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dqkv) % 16) == 0,
                "dsg_attention_bwd: unaligned pointer");
  if (n == 0) return DSG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (lse) {  // tcgen05 path: needs the forward's log-sum-exp (dsg_attention_train)
    const int rc = launch_attention_bwd_tc((const __half*)qkv, (const __half*)out, (const __half*)dout, lse,
                                           (__half*)dqkv, n, tokens, heads, head_dim, st);
    if (rc <= 0) return rc;
    DSG_CHECK_ARG(ws != nullptr, "dsg_attention_bwd: shape outside the tcgen05 kernel and no workspace given");
  }
  const float scale = 1.0f / sqrtf((float)head_dim);
  const float scale_log2 = scale * 1.4426950408889634f;
  float* lse_ws = ws;
  float* delta = ws + (int64_t)n * heads * tokens;
  const dim3 grid(ceil_div(tokens, AB_ROWS), heads, n);
  attn_bwd_prep_kernel<<<grid, AB_ROWS, 0, st>>>((const __half*)qkv, (const __half*)out, (const __half*)dout, lse_ws,
                                                 delta, tokens, heads, scale_log2);
  DSG_CUDA_LAUNCH_CHECK("dsg_attention_bwd/prep");
  attn_bwd_dkv_kernel<<<grid, AB_ROWS, 0, st>>>((const __half*)qkv, (const __half*)dout, lse_ws, delta, (__half*)dqkv,
                                                tokens, heads, scale, scale_log2);
  DSG_CUDA_LAUNCH_CHECK("dsg_attention_bwd/dkv");
  attn_bwd_dq_kernel<<<grid, AB_ROWS, 0, st>>>((const __half*)qkv, (const __half*)dout, lse_ws, delta, (__half*)dqkv,
                                               tokens, heads, scale, scale_log2);
  DSG_CUDA_LAUNCH_CHECK("dsg_attention_bwd/dq");
  return DSG_OK;
}
