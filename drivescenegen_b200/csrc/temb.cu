// temb.cu — timestep embedding path in two launches (upstream: ~50 tiny cuBLAS launches per step).
// Replaces diffusers 0.20.0 models/embeddings.py get_timestep_embedding + TimestepEmbedding and every
// ResnetBlock2D.time_emb_proj(SiLU(emb)).  fp32 weights, fp32 math.
#include "common.cuh"

namespace dsg {

constexpr int TE_THREADS = 256;

// grid = batch, one thread per output feature.  emb_ws[b][:] = SiLU(linear_2(SiLU(linear_1(sinusoid(t[b]))))).
// Weights come TRANSPOSED ([in][out]) so that consecutive threads read consecutive addresses and every thread runs
// an independent, unrollable dot product (the [out][in] + warp-reduction form was a chain of dependent latencies).
__global__ void __launch_bounds__(TE_THREADS) temb_mlp_kernel(const float* __restrict__ t,
                                                              const float* __restrict__ freqs, int half, int flip,
                                                              const float* __restrict__ w1t, const float* __restrict__ b1,
                                                              const float* __restrict__ w2t, const float* __restrict__ b2,
                                                              int hidden, float* __restrict__ emb_ws,
                                                              float* __restrict__ saved) {
  // saved (training only, may be NULL): four dense arrays [batch][2*half] sinusoid, then [batch][hidden] each of
  // pre1, SiLU(pre1), pre2
  extern __shared__ float sm[];
  float* e = sm;               // [2*half]
  float* h1 = sm + 2 * half;   // [hidden]
  const int b = blockIdx.x, in_dim = 2 * half;
  pdl_sync();
  const float tv = t[b];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float arg = __fmul_rn(tv, freqs[i]);
    const float s = sinf(arg), c = cosf(arg);
    // upstream: cat([sin, cos]); flip_sin_to_cos swaps the halves
    e[flip ? half + i : i] = s;
    e[flip ? i : half + i] = c;
  }
  __syncthreads();
  const int64_t nb = gridDim.x;
  float* sv_e = saved ? saved + (int64_t)b * in_dim : nullptr;
  float* sv_h = saved ? saved + nb * in_dim + (int64_t)b * hidden : nullptr;  // + k * nb * hidden for array k
  if (saved)
    for (int i = threadIdx.x; i < in_dim; i += blockDim.x) sv_e[i] = e[i];
  for (int r = threadIdx.x; r < hidden; r += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 4 <= in_dim; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(w1t[(int64_t)(k + u) * hidden + r], e[k + u], acc[u]);
    }
    for (; k < in_dim; ++k) acc[0] = fmaf(w1t[(int64_t)k * hidden + r], e[k], acc[0]);
    const float y = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + b1[r];
    h1[r] = y / (1.0f + expf(-y));
    if (saved) { sv_h[r] = y; sv_h[nb * hidden + r] = h1[r]; }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < hidden; r += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 4 <= hidden; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(w2t[(int64_t)(k + u) * hidden + r], h1[k + u], acc[u]);
    }
    for (; k < hidden; ++k) acc[0] = fmaf(w2t[(int64_t)k * hidden + r], h1[k], acc[0]);
    const float y = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + b2[r];
    emb_ws[(int64_t)b * hidden + r] = y / (1.0f + expf(-y));
    if (saved) sv_h[2 * nb * hidden + r] = y;
  }
}

// one warp per projection row; out[b][r] = bp[r] + <wp[r], semb[b]>
constexpr int TP_BT = 16;  // batch tile staged in shared memory
__global__ void __launch_bounds__(TE_THREADS) temb_proj_kernel(const float* __restrict__ semb,
                                                               const float* __restrict__ wp,
                                                               const float* __restrict__ bp, int hidden,
                                                               int proj_total, float* __restrict__ out, int batch) {
  extern __shared__ float sm[];  // [TP_BT][hidden]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int r = blockIdx.x * nwarp + warp;
  pdl_sync();
  for (int b0 = 0; b0 < batch; b0 += TP_BT) {
    const int nb = min(TP_BT, batch - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * hidden; i += blockDim.x) sm[i] = semb[(int64_t)b0 * hidden + i];
    __syncthreads();
    if (r < proj_total) {
      float acc[TP_BT];
#pragma unroll
      for (int j = 0; j < TP_BT; ++j) acc[j] = 0.f;
      for (int k = lane; k < hidden; k += 32) {
        const float w = wp[(int64_t)r * hidden + k];
#pragma unroll
        for (int j = 0; j < TP_BT; ++j)
          if (j < nb) acc[j] = fmaf(w, sm[j * hidden + k], acc[j]);
      }
      const float bias = bp[r];
#pragma unroll
      for (int j = 0; j < TP_BT; ++j) {
        if (j < nb) {
          const float v = warp_sum(acc[j]);
          if (lane == 0) out[(int64_t)(b0 + j) * proj_total + r] = v + bias;
        }
      }
    }
  }
}

}  // namespace dsg

using namespace dsg;

extern "C" int dsg_time_embed_ex(const float* t, const float* freqs, int32_t half, int32_t flip_sin_to_cos,
                                 const float* w1t, const float* b1, const float* w2t, const float* b2, int32_t hidden,
                                 const float* wp, const float* bp, int32_t proj_total, float* emb_ws, float* out,
                                 int32_t batch, float* saved, void* stream) {
  DSG_CHECK_ARG(t && freqs && w1t && b1 && w2t && b2 && wp && bp && emb_ws && out, "dsg_time_embed: null pointer");
  DSG_CHECK_ARG(half > 0 && hidden > 0 && hidden <= 2048 && proj_total > 0 && batch >= 0,
                "dsg_time_embed: bad sizes");
  if (batch == 0) return DSG_OK;
  const size_t sm1 = (size_t)(2 * half + hidden) * sizeof(float);
  launch_k(temb_mlp_kernel, dim3(batch), dim3(TE_THREADS), sm1, (cudaStream_t)stream, t, freqs, half, flip_sin_to_cos,
           w1t, b1, w2t, b2, hidden, emb_ws, saved);
  DSG_CUDA_LAUNCH_CHECK("dsg_time_embed/mlp");
  const size_t sm2 = (size_t)TP_BT * hidden * sizeof(float);
  if (sm2 > 48 * 1024)
    cudaFuncSetAttribute(temb_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  const int rows_per_cta = TE_THREADS / 32;
  launch_k(temb_proj_kernel, dim3(ceil_div(proj_total, rows_per_cta)), dim3(TE_THREADS), sm2, (cudaStream_t)stream,
           (const float*)emb_ws, wp, bp, hidden, proj_total, out, batch);
  DSG_CUDA_LAUNCH_CHECK("dsg_time_embed/proj");
  return DSG_OK;
}

extern "C" int dsg_time_embed(const float* t, const float* freqs, int32_t half, int32_t flip_sin_to_cos,
                              const float* w1t, const float* b1, const float* w2t, const float* b2, int32_t hidden,
                              const float* wp, const float* bp, int32_t proj_total, float* emb_ws, float* out,
                              int32_t batch, void* stream) {
  return dsg_time_embed_ex(t, freqs, half, flip_sin_to_cos, w1t, b1, w2t, b2, hidden, wp, bp, proj_total, emb_ws, out,
                           batch, nullptr, stream);
}
