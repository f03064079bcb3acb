// train_small.cu — the small fp32 pieces of the training step:
//   * backward of the timestep-embedding path (Timesteps -> Linear -> SiLU -> Linear -> SiLU -> per-block Linear):
//     generic small-batch linear dgrad / wgrad kernels (M = batch <= a few dozen rows; latency-bound, CUDA cores),
//   * gradient global-norm + unscale + AdamW over ONE flat parameter / gradient buffer (replaces
//     accelerator.clip_grad_norm_ + torch.optim.AdamW.step of DriveSceneGen/pipeline/training_pipeline.py:88-89,
//     DriveSceneGen/scripts/train.py:66): HBM-bound, 4 reads + 3 writes per parameter in a single pass.
#include "common.cuh"

namespace dsg {

__device__ __forceinline__ float silu_grad_exact(float y) {
  const float sg = 1.0f / (1.0f + expf(-y));
  return sg * (1.0f + y * (1.0f - sg));
}

// dx[n][k] = (sum_r dy[n][dy_off + r] * w[r][k]) * (pre ? silu'(pre[n][k]) : 1);  w is [rows][cols] (torch [out][in]).
// One block per (16 columns, sample): 16 row lanes stride over the rows (5824 for the time_emb_proj stack — one thread per
// output walked them serially), then the lanes are summed in a fixed order (deterministic).
constexpr int LD_KT = 16, LD_RL = 16;
__global__ void __launch_bounds__(LD_KT * LD_RL) lin_dgrad_small_kernel(const float* __restrict__ dy, int ldy, int dy_off,
                                                                        const float* __restrict__ w, int rows, int cols,
                                                                        const float* __restrict__ pre,
                                                                        float* __restrict__ dx, int batch) {
  __shared__ float part[LD_RL][LD_KT + 1];
  const int kx = threadIdx.x % LD_KT, ry = threadIdx.x / LD_KT;
  const int k = blockIdx.x * LD_KT + kx;
  const int n = blockIdx.y;
  const float* d = dy + (int64_t)n * ldy + dy_off;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < cols) {
    int r = ry;
    for (; r + 3 * LD_RL < rows; r += 4 * LD_RL) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(d[r + u * LD_RL], w[(int64_t)(r + u * LD_RL) * cols + k], acc[u]);
    }
    for (; r < rows; r += LD_RL) acc[0] = fmaf(d[r], w[(int64_t)r * cols + k], acc[0]);
  }
  part[ry][kx] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (ry == 0 && k < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < LD_RL; ++i) v += part[i][kx];
    if (pre) v *= silu_grad_exact(pre[(int64_t)n * cols + k]);
    dx[(int64_t)n * cols + k] = v;
  }
}

// dw[r][k] = s * sum_n dy[n][dy_off + r] * x[n][k];  db[r] = s * sum_n dy[n][dy_off + r]   (x optional-activated)
__global__ void __launch_bounds__(256) lin_wgrad_small_kernel(const float* __restrict__ dy, int ldy, int dy_off,
                                                              const float* __restrict__ x, int ldx, int rows, int cols,
                                                              int batch, const float* __restrict__ inv_scale,
                                                              float* __restrict__ dw, float* __restrict__ db) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (r >= rows) return;
  const float s = inv_scale ? inv_scale[0] : 1.0f;
  if (k < cols) {
    float acc = 0.f;
    for (int n = 0; n < batch; ++n) acc = fmaf(dy[(int64_t)n * ldy + dy_off + r], x[(int64_t)n * ldx + k], acc);
    dw[(int64_t)r * cols + k] = acc * s;
  }
  if (db && blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int n = 0; n < batch; ++n) acc += dy[(int64_t)n * ldy + dy_off + r];
    db[r] = acc * s;
  }
}

// ---- flat-buffer optimizer
// partial[b] = sum of squares of g over block b's slice (fixed order inside a block => deterministic); flags nonfinite
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t numel,
                                                            double* __restrict__ partial) {
  double acc = 0.0;
  float facc = 0.f;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    facc = fmaf(v, v, facc);
    if (++cnt == 64) { acc += (double)facc; facc = 0.f; cnt = 0; }
  }
  acc += (double)facc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sm[i];
    partial[blockIdx.x] = t;
  }
}
// out[0] = total L2 norm of (g * inv_loss_scale); out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) folded with
// inv_loss_scale (multiply RAW gradients by out[1]); out[2] = 1 if the norm is not finite (skip the step), else 0
__global__ void __launch_bounds__(256) grad_norm_finalize_kernel(const double* __restrict__ partial, int parts,
                                                                 float inv_loss_scale, float max_norm,
                                                                 float* __restrict__ out) {
  // fixed assignment of partials to threads + fixed-order tree: deterministic, and not a chain of dependent loads
  __shared__ double sm[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < parts; i += 256) t += partial[i];
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  t = sm[0];
  const float norm = (float)sqrt(t) * inv_loss_scale;
  const bool finite = norm <= 3.0e38f;  // false for inf and nan
  float coef = inv_loss_scale;
  if (finite && max_norm > 0.f) {
    const float c = max_norm / (norm + 1e-6f);
    if (c < 1.0f) coef *= c;
  }
  out[0] = norm;
  out[1] = coef;
  out[2] = finite ? 0.f : 1.f;
}

// torch.optim.AdamW (decoupled weight decay, bias correction, no amsgrad), single tensor form over flat buffers:
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// g is multiplied by ctl[1] first (unscale + clip); the whole update is skipped when ctl[2] != 0 (inf/nan gradients).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t numel,
                                                    float lr, float b1, float b2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, const float* __restrict__ ctl) {
  const float gcoef = ctl ? ctl[1] : 1.0f;
  if (ctl && ctl[2] != 0.f) return;
  const float step_size = lr / bc1;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < numel;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    if (i + 4 <= numel) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gg = gp[j] * gcoef;
        pp[j] *= 1.0f - lr * wd;
        mp[j] = b1 * mp[j] + (1.0f - b1) * gg;
        vp[j] = b2 * vp[j] + (1.0f - b2) * gg * gg;
        const float denom = sqrtf(vp[j]) / bc2_sqrt + eps;
        pp[j] -= step_size * (mp[j] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t k = i; k < numel; ++k) {
        const float gg = g[k] * gcoef;
        float pk = p[k] * (1.0f - lr * wd);
        const float mk = b1 * m[k] + (1.0f - b1) * gg;
        const float vk = b2 * v[k] + (1.0f - b2) * gg * gg;
        pk -= step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
        p[k] = pk; m[k] = mk; v[k] = vk;
      }
    }
  }
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_lin_dgrad_small(const float* dy, int32_t ldy, int32_t dy_off, const float* w, int32_t rows, int32_t cols,
                        const float* pre, float* dx, int32_t batch, void* stream) {
  DSG_CHECK_ARG(dy && w && dx && rows > 0 && cols > 0 && batch >= 0 && batch <= 65535, "dsg_lin_dgrad_small: bad args");
  if (batch == 0) return DSG_OK;
  lin_dgrad_small_kernel<<<dim3(ceil_div(cols, LD_KT), batch), LD_KT * LD_RL, 0, (cudaStream_t)stream>>>(
      dy, ldy, dy_off, w, rows, cols, pre, dx, batch);
  DSG_CUDA_LAUNCH_CHECK("dsg_lin_dgrad_small");
  return DSG_OK;
}

int dsg_lin_wgrad_small(const float* dy, int32_t ldy, int32_t dy_off, const float* x, int32_t ldx, int32_t rows,
                        int32_t cols, int32_t batch, const float* inv_scale, float* dw, float* db, void* stream) {
  DSG_CHECK_ARG(dy && x && dw && rows > 0 && rows <= 65535 && cols > 0 && batch >= 0, "dsg_lin_wgrad_small: bad args");
  lin_wgrad_small_kernel<<<dim3(ceil_div(cols, 256), rows), 256, 0, (cudaStream_t)stream>>>(
      dy, ldy, dy_off, x, ldx, rows, cols, batch, inv_scale, dw, db);
  DSG_CUDA_LAUNCH_CHECK("dsg_lin_wgrad_small");
  return DSG_OK;
}

int dsg_grad_norm(const float* g, int64_t numel, double* partial, int32_t parts, float inv_loss_scale, float max_norm,
                  float* out3, void* stream) {
  DSG_CHECK_ARG(g && partial && out3 && numel >= 0 && parts >= 1 && parts <= 65535, "dsg_grad_norm: bad args");
  sumsq_partial_kernel<<<parts, 256, 0, (cudaStream_t)stream>>>(g, numel, partial);
  DSG_CUDA_LAUNCH_CHECK("dsg_grad_norm/partial");
  grad_norm_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial, parts, inv_loss_scale, max_norm, out3);
  DSG_CUDA_LAUNCH_CHECK("dsg_grad_norm/finalize");
  return DSG_OK;
}

int dsg_adamw_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int32_t step, const float* ctl, void* stream) {
  DSG_CHECK_ARG(p && g && m && v && numel >= 0 && step >= 1, "dsg_adamw_step: bad args");
  DSG_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0,
                "dsg_adamw_step: buffers must be 16-byte aligned");
  if (numel == 0) return DSG_OK;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int64_t blocks = ceil_div64(numel, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, numel, lr, beta1, beta2, eps,
                                                                   weight_decay, bc1, bc2_sqrt, ctl);
  DSG_CUDA_LAUNCH_CHECK("dsg_adamw_step");
  return DSG_OK;
}
}
