// elementwise.cu — HBM-bound fp32 kernels: DDPM/DDIM scheduler.step, add_noise, latent -> image.
// Upstream semantics: diffusers 0.20.0 schedulers/scheduling_ddpm.py::step, scheduling_ddim.py::step,
// DDPMScheduler.add_noise, pipelines/ddpm/pipeline_ddpm.py post-processing.
// Reference call sites: DriveSceneGen/pipeline/training_pipeline.py:26-32,80; DriveSceneGen/scripts/generation.py:14-24.
#include "common.cuh"

namespace dsg {

// Every product/sum is an explicitly rounded intrinsic: ptxas must not contract mul+add into FMA, so that the
// result is bit-identical to the op-by-op fp32 evaluation the reference performs with ATen on the CPU.
__device__ __forceinline__ float ddpm_one(float e, float x, float z, const float* c) {
  // pred_original_sample = (sample - beta_prod_t**0.5 * model_output) / alpha_prod_t**0.5
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c[0], e)), c[1]);
  // clamp(-r, r); torch.clamp propagates NaN (fminf / fmaxf alone would turn a NaN prediction into +-r)
  x0 = (x0 != x0) ? x0 : fminf(fmaxf(x0, -c[5]), c[5]);
  // pred_prev_sample = coeff_x0 * x0 + coeff_xt * sample
  float p = __fadd_rn(__fmul_rn(c[2], x0), __fmul_rn(c[3], x));
  // + sigma * noise (t > 0)
  if (c[6] != 0.0f) p = __fadd_rn(p, __fmul_rn(c[4], z));
  return p;
}
__device__ __forceinline__ float ddim_one(float e, float x, float z, const float* c) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c[0], e)), c[1]);
  x0 = (x0 != x0) ? x0 : fminf(fmaxf(x0, -c[5]), c[5]);
  // prev = abar_prev**0.5 * x0 + (1 - abar_prev - std^2)**0.5 * eps
  float p = __fadd_rn(__fmul_rn(c[2], x0), __fmul_rn(c[3], e));
  if (c[6] != 0.0f) p = __fadd_rn(p, __fmul_rn(c[4], z));
  return p;
}

template <bool DDIM>
__global__ void __launch_bounds__(256) sched_step_kernel(const float* __restrict__ eps, const float* __restrict__ x,
                                                         const float* __restrict__ z, float* __restrict__ out,
                                                         int64_t numel, const float* __restrict__ table,
                                                         const int32_t* __restrict__ row_dev, int32_t row) {
  __shared__ float c[8];
  if (threadIdx.x < 8) {
    int r = row_dev ? *row_dev : row;
    c[threadIdx.x] = table[(int64_t)r * 8 + threadIdx.x];
  }
  __syncthreads();
  float cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cc[i] = c[i];
  const bool noisy = (cc[6] != 0.0f) && (z != nullptr);
  if (!noisy) cc[6] = 0.0f;
  const int64_t nvec = numel >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 e4 = __ldg(reinterpret_cast<const float4*>(eps) + i);
    float4 x4 = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 z4 = noisy ? __ldg(reinterpret_cast<const float4*>(z) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    if (DDIM) {
      o.x = ddim_one(e4.x, x4.x, z4.x, cc); o.y = ddim_one(e4.y, x4.y, z4.y, cc);
      o.z = ddim_one(e4.z, x4.z, z4.z, cc); o.w = ddim_one(e4.w, x4.w, z4.w, cc);
    } else {
      o.x = ddpm_one(e4.x, x4.x, z4.x, cc); o.y = ddpm_one(e4.y, x4.y, z4.y, cc);
      o.z = ddpm_one(e4.z, x4.z, z4.z, cc); o.w = ddpm_one(e4.w, x4.w, z4.w, cc);
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
  // tail (numel not a multiple of 4)
  for (int64_t i = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    float zz = noisy ? z[i] : 0.f;
    out[i] = DDIM ? ddim_one(eps[i], x[i], zz, cc) : ddpm_one(eps[i], x[i], zz, cc);
  }
}

__global__ void __launch_bounds__(256) add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ nz,
                                                        const int64_t* __restrict__ t,
                                                        const float* __restrict__ sqrt_ac,
                                                        const float* __restrict__ sqrt_1mac, float* __restrict__ out,
                                                        int64_t per_sample, int32_t table_len) {
  const int n = blockIdx.y;
  const int64_t tt = t[n];
  // a timestep outside the table raises IndexError upstream; here the sample's output is poisoned with NaN instead of
  // reading out of bounds (the host wrapper range-checks timesteps it can see without a device sync)
  const bool ok = tt >= 0 && tt < table_len;
  const float a = ok ? sqrt_ac[tt] : __int_as_float(0x7fc00000), b = ok ? sqrt_1mac[tt] : a;
  const float* xs = x0 + (int64_t)n * per_sample;
  const float* ns = nz + (int64_t)n * per_sample;
  float* os = out + (int64_t)n * per_sample;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((per_sample & 3) == 0) {
    for (int64_t i = start; i < (per_sample >> 2); i += stride) {
      float4 xv = __ldg(reinterpret_cast<const float4*>(xs) + i);
      float4 nv = __ldg(reinterpret_cast<const float4*>(ns) + i);
      float4 o;
      o.x = __fadd_rn(__fmul_rn(a, xv.x), __fmul_rn(b, nv.x));
      o.y = __fadd_rn(__fmul_rn(a, xv.y), __fmul_rn(b, nv.y));
      o.z = __fadd_rn(__fmul_rn(a, xv.z), __fmul_rn(b, nv.z));
      o.w = __fadd_rn(__fmul_rn(a, xv.w), __fmul_rn(b, nv.w));
      reinterpret_cast<float4*>(os)[i] = o;
    }
  } else {
    for (int64_t i = start; i < per_sample; i += stride)
      os[i] = __fadd_rn(__fmul_rn(a, xs[i]), __fmul_rn(b, ns[i]));
  }
}

// One block: the next timestep of the sampling schedule -> the U-Net's per-sample timestep values and the scheduler's
// coefficient-table row, then advance the device-side step counter.  First node of the captured denoise-step graph, so a
// replay needs no host-side argument updates at all.
__global__ void step_advance_kernel(const int32_t* __restrict__ schedule, int32_t* __restrict__ state,
                                    float* __restrict__ t_f, int32_t batch, int32_t* __restrict__ row) {
  const int k = state[0], n_steps = state[1];
  const int t = schedule[k < n_steps ? k : n_steps - 1];
  for (int i = threadIdx.x; i < batch; i += blockDim.x) t_f[i] = (float)t;
  __syncthreads();
  if (threadIdx.x == 0) {
    *row = t;
    state[0] = k + 1;
  }
}

// NCHW fp32 -> NHWC image; one thread per pixel.
__global__ void __launch_bounds__(256) latent_to_image_kernel(const float* __restrict__ lat, uint8_t* __restrict__ u8,
                                                              float* __restrict__ f32, int c, int64_t hw,
                                                              int64_t total_px) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_px; p += stride) {
    const int64_t n = p / hw, s = p - n * hw;
    for (int ch = 0; ch < c; ++ch) {
      float v = lat[(n * c + ch) * hw + s];
      // (image / 2 + 0.5).clamp(0, 1)
      v = __fadd_rn(__fdiv_rn(v, 2.0f), 0.5f);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      if (f32) f32[p * c + ch] = v;
      // numpy_to_pil: (x * 255).round().astype(uint8); numpy rounds half to even
      if (u8) u8[p * c + ch] = (uint8_t)__float2int_rn(__fmul_rn(v, 255.0f));
    }
  }
}

static inline int grid_for(int64_t work_items, int threads, int max_blocks = 148 * 8) {
  int64_t b = ceil_div64(work_items, threads);
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_ddpm_step(const float* eps, const float* sample, const float* noise, float* prev, int64_t numel,
                  const float* coef_table, const int32_t* row_dev, int32_t row, void* stream) {
  DSG_CHECK_ARG(numel >= 0, "dsg_ddpm_step: negative size");
  if (numel == 0) return DSG_OK;
  DSG_CHECK_ARG(eps && sample && prev && coef_table, "dsg_ddpm_step: null pointer");
  DSG_CHECK_ARG(((uintptr_t)eps | (uintptr_t)sample | (uintptr_t)prev | (uintptr_t)noise) % 16 == 0,
                "dsg_ddpm_step: pointers must be 16-byte aligned");
  sched_step_kernel<false><<<grid_for(numel / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
      eps, sample, noise, prev, numel, coef_table, row_dev, row);
  DSG_CUDA_LAUNCH_CHECK("dsg_ddpm_step");
  return DSG_OK;
}

int dsg_ddim_step(const float* eps, const float* sample, const float* noise, float* prev, int64_t numel,
                  const float* coef_table, const int32_t* row_dev, int32_t row, void* stream) {
  DSG_CHECK_ARG(numel >= 0, "dsg_ddim_step: negative size");
  if (numel == 0) return DSG_OK;
  DSG_CHECK_ARG(eps && sample && prev && coef_table, "dsg_ddim_step: null pointer");
  DSG_CHECK_ARG(((uintptr_t)eps | (uintptr_t)sample | (uintptr_t)prev | (uintptr_t)noise) % 16 == 0,
                "dsg_ddim_step: pointers must be 16-byte aligned");
  sched_step_kernel<true><<<grid_for(numel / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
      eps, sample, noise, prev, numel, coef_table, row_dev, row);
  DSG_CUDA_LAUNCH_CHECK("dsg_ddim_step");
  return DSG_OK;
}

int dsg_add_noise(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                  const float* sqrt_1mac, int32_t table_len, float* out, int32_t batch, int64_t per_sample,
                  void* stream) {
  DSG_CHECK_ARG(batch >= 0 && batch <= 65535 && per_sample >= 0 && table_len > 0,
                "dsg_add_noise: bad batch/per_sample/table_len");
  if (batch == 0 || per_sample == 0) return DSG_OK;
  DSG_CHECK_ARG(x0 && noise && t && sqrt_ac && sqrt_1mac && out, "dsg_add_noise: null pointer");
  DSG_CHECK_ARG(((uintptr_t)x0 | (uintptr_t)noise | (uintptr_t)out) % 16 == 0,
                "dsg_add_noise: pointers must be 16-byte aligned");
  dim3 grid(grid_for(per_sample / 4 + 1, 256, 148 * 2), batch);
  add_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x0, noise, t, sqrt_ac, sqrt_1mac, out, per_sample,
                                                          table_len);
  DSG_CUDA_LAUNCH_CHECK("dsg_add_noise");
  return DSG_OK;
}

int dsg_step_advance(const int32_t* schedule, int32_t* state, float* t_f, int32_t batch, int32_t* row, void* stream) {
  DSG_CHECK_ARG(schedule && state && t_f && row && batch > 0, "dsg_step_advance: bad arguments");
  step_advance_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(schedule, state, t_f, batch, row);
  DSG_CUDA_LAUNCH_CHECK("dsg_step_advance");
  return DSG_OK;
}

int dsg_latent_to_image(const float* latent, uint8_t* out_u8, float* out_f32, int32_t n, int32_t c, int32_t h,
                        int32_t w, void* stream) {
  DSG_CHECK_ARG(n >= 0 && c > 0 && h > 0 && w > 0, "dsg_latent_to_image: bad shape");
  if (n == 0) return DSG_OK;
  DSG_CHECK_ARG(latent && (out_u8 || out_f32), "dsg_latent_to_image: null pointer");
  const int64_t hw = (int64_t)h * w;
  latent_to_image_kernel<<<grid_for(hw * n, 256), 256, 0, (cudaStream_t)stream>>>(latent, out_u8, out_f32, c, hw,
                                                                                 hw * n);
  DSG_CUDA_LAUNCH_CHECK("dsg_latent_to_image");
  return DSG_OK;
}
}
