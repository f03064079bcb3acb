// conv_small_bwd.cu — backward of the two degenerate convolutions (conv_in 3->C0, conv_out C0->3) and the
// gradient-scale entry point of the backward pass (training path).
//
//   dsg_grad_scale      amax of the incoming fp32 output gradient -> a power-of-two factor s with amax * s in [1, 2)
//                       (device scalars {s, 1/s}): every fp16 activation gradient of the backward pass is carried
//                       times s, every parameter-gradient finaliser multiplies by 1/s, so fp16 range is used whatever
//                       loss scale (GradScaler or none) the caller applies.  Power of two => exact.
//   conv_out data grad  = a conv_in-shaped operation (3 -> C0 channels, NCHW fp32 in, NHWC fp16 out): dsg_conv_in runs
//                       it on the flipped / transposed weights that dsg_conv_out_dgrad_weight prepares (times s).
//   dsg_small_wgrad     weight gradient of either conv: sum over pixels of wide[q][wc] * narrow[c][q +- tap].
// Replaces the cuDNN dgrad/wgrad autograd runs for UNet2DModel.conv_in / conv_out (diffusers 0.20.0 models/unet_2d.py)
// from `accelerator.backward(loss)` (DriveSceneGen/pipeline/training_pipeline.py:86).
#include "common.cuh"
#include "reduce.cuh"

namespace dsg {

constexpr int SB_THREADS = 256;

__global__ void __launch_bounds__(SB_THREADS) amax_partial_kernel(const float* __restrict__ x, int64_t numel,
                                                                  float* __restrict__ partial) {
  float m = 0.f;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(x[i]);
    bad |= !(v <= 3.0e38f);  // inf or nan
    m = fmaxf(m, v);
  }
  if (bad) m = INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[SB_THREADS / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < SB_THREADS / 32; ++i) m = fmaxf(m, sm[i]);
    partial[blockIdx.x] = m;
  }
}
__global__ void __launch_bounds__(256) grad_scale_finalize_kernel(const float* __restrict__ partial, int parts,
                                                                  float* __restrict__ scale) {
  float m = 0.f;
  for (int i = threadIdx.x; i < parts; i += 256) m = fmaxf(m, partial[i]);   // max is order-independent
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
  float s = 1.0f;
  if (m > 0.f && m <= 3.0e38f) {
    int e;
    frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(1 - e) in [1, 2)
    e = 1 - e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    s = ldexpf(1.0f, e);
  }
  scale[0] = s;
  scale[1] = 1.0f / s;
}

// w [nc][wc][3][3] (conv_out weight) -> wt [wc][nc][3][3] = s * w[c][ci][2-ky][2-kx]
__global__ void __launch_bounds__(256) conv_out_dgrad_weight_kernel(const float* __restrict__ w, int nc, int wc,
                                                                    const float* __restrict__ scale,
                                                                    float* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc * wc * 9) return;
  const int t = i % 9, c = (i / 9) % nc, ci = i / (9 * nc);
  const float s = scale ? scale[0] : 1.0f;
  wt[i] = s * w[((int64_t)c * wc + ci) * 9 + (8 - t)];
}

// grid-stride over image rows; thread = (PAIR of wide channels, x slice); 2 * 9 * NC accumulators per thread.
// A thread takes FOUR adjacent pixels per step: their 3 x 3 windows of the narrow tensor overlap, so one (channel, row) of
// the window costs 6 shared-memory loads for 24 FMA pairs (a pixel at a time it was 3 loads per 3).  The loop always runs
// the conv_in geometry (narrow at q + tap - 1); the conv_out form (narrow at q - tap + 1) is the same sum with the tap
// index mirrored, applied when the accumulators are written out.  The reduction scratch reuses the row buffer.
template <int NC>
__global__ void __launch_bounds__(SB_THREADS) small_wgrad_kernel(const __half* __restrict__ wide,
                                                                 const float* __restrict__ narrow, int n, int h, int w,
                                                                 int wc, int sgn, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [NC][3][ws] rows y-1..y+1 of the narrow tensor, zero padded; reduction scratch after
  const int ws = w + 8;          // column x + 1 holds pixel x; columns 0 and w + 1 .. w + 7 are zero
  float* s_rows = sm;
  float* s_red = sm;             // [slices][NC * 9][wc], used after the loop
  const int wc2 = wc >> 1;
  const int c_t = threadIdx.x % wc2, slice = threadIdx.x / wc2, nslices = SB_THREADS / wc2;
  float acc[NC][9][2];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) { acc[c][t][0] = 0.f; acc[c][t][1] = 0.f; }
  float csum[NC];    // per-thread share of the plain sums of the narrow channels (the conv_out bias gradient)
#pragma unroll
  for (int c = 0; c < NC; ++c) csum[c] = 0.f;
  const int64_t rows = (int64_t)n * h;
  const int64_t plane = (int64_t)h * w;
  constexpr int U = 4;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int nn = (int)(r / h), y = (int)(r % h);
    __syncthreads();
    for (int i = threadIdx.x; i < NC * 3 * ws; i += SB_THREADS) {
      const int xx = i % ws - 1, ry = (i / ws) % 3, c = i / (3 * ws);
      const int yy = y + ry - 1;
      float v = 0.f;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = narrow[((int64_t)nn * NC + c) * plane + (int64_t)yy * w + xx];
      s_rows[i] = v;
      if (ry == 1) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) csum[cc] += (c == cc) ? v : 0.f;
      }
    }
    __syncthreads();
    const __half2* wp = reinterpret_cast<const __half2*>(wide + (((int64_t)nn * h + y) * w) * wc) + c_t;
    // the next step's four pixels are in flight while this step's 216 FMAs run (the loop was latency-bound: one round
    // trip to HBM per 4 pixels with nothing to overlap it inside the warp)
    __half2 nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int xx = slice * U + u;
      nxt[u] = xx < w ? wp[(int64_t)xx * wc2] : __float2half2_rn(0.f);
    }
    for (int x0 = slice * U; x0 < w; x0 += nslices * U) {
      float2 a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = __half22float2(nxt[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int xx = x0 + nslices * U + u;
        nxt[u] = xx < w ? wp[(int64_t)xx * wc2] : __float2half2_rn(0.f);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          // pixels x0 .. x0 + 3 with taps kx = 0..2 read smem columns x0 + u + kx = x0 .. x0 + 5
          const float* rowp = s_rows + (c * 3 + ky) * ws + x0;
          float nv[U + 2];
#pragma unroll
          for (int j = 0; j < U + 2; ++j) nv[j] = rowp[j];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int u = 0; u < U; ++u) {
              acc[c][ky * 3 + kx][0] = fmaf(a[u].x, nv[u + kx], acc[c][ky * 3 + kx][0]);
              acc[c][ky * 3 + kx][1] = fmaf(a[u].y, nv[u + kx], acc[c][ky * 3 + kx][1]);
            }
        }
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int to = sgn > 0 ? t : 8 - t;
      s_red[(slice * NC * 9 + c * 9 + to) * wc + 2 * c_t] = acc[c][t][0];
      s_red[(slice * NC * 9 + c * 9 + to) * wc + 2 * c_t + 1] = acc[c][t][1];
    }
  __syncthreads();
  float* o = partial + (int64_t)blockIdx.x * (NC * 9 * wc + NC);
  for (int i = threadIdx.x; i < NC * 9 * wc; i += SB_THREADS) {
    float t = 0.f;
    for (int s = 0; s < nslices; ++s) t += s_red[s * NC * 9 * wc + i];
    o[i] = t;
  }
  // plain sums: warp shuffle, then the 8 warps in a fixed order
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float v = warp_sum(csum[c]);
    if ((threadIdx.x & 31) == 0) s_red[c * 8 + (threadIdx.x >> 5)] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < NC) {
    float t = 0.f;
    for (int i = 0; i < SB_THREADS / 32; ++i) t += s_red[threadIdx.x * 8 + i];
    o[NC * 9 * wc + threadIdx.x] = t;
  }
}

// reduced [nc*9*wc + nc] (already scaled) -> dw (layout: wide_major ? [wc][nc][9] : [nc][wc][9]) and narrow-sum [nc]
__global__ void __launch_bounds__(256) small_wgrad_remap_kernel(const float* __restrict__ red, int nc, int wc,
                                                                int wide_major, float* __restrict__ dw,
                                                                float* __restrict__ nsum) {
  const int stride = nc * 9 * wc + nc;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  const float t = red[i];
  if (i >= nc * 9 * wc) {
    if (nsum) nsum[i - nc * 9 * wc] = t;
    return;
  }
  const int c_t = i % wc, tap = (i / wc) % 9, c = i / (9 * wc);
  const int64_t o = wide_major ? ((int64_t)c_t * nc + c) * 9 + tap : ((int64_t)c * wc + c_t) * 9 + tap;
  dw[o] = t;
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_grad_scale(const float* dout, int64_t numel, float* partial, int32_t parts, float* scale, void* stream) {
  DSG_CHECK_ARG(dout && partial && scale && numel >= 0 && parts >= 1 && parts <= 4096, "dsg_grad_scale: bad args");
  amax_partial_kernel<<<parts, SB_THREADS, 0, (cudaStream_t)stream>>>(dout, numel, partial);
  DSG_CUDA_LAUNCH_CHECK("dsg_grad_scale/amax");
  grad_scale_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial, parts, scale);
  DSG_CUDA_LAUNCH_CHECK("dsg_grad_scale/finalize");
  return DSG_OK;
}

int dsg_conv_out_dgrad_weight(const float* w, int32_t cout, int32_t cin, const float* scale, float* wt, void* stream) {
  DSG_CHECK_ARG(w && wt && cout >= 1 && cout <= 4 && cin > 0, "dsg_conv_out_dgrad_weight: bad args");
  conv_out_dgrad_weight_kernel<<<ceil_div(cout * cin * 9, 256), 256, 0, (cudaStream_t)stream>>>(w, cout, cin, scale,
                                                                                               wt);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_out_dgrad_weight");
  return DSG_OK;
}

int dsg_small_wgrad(const void* wide_h16, const float* narrow_nchw, int32_t n, int32_t h, int32_t w, int32_t wc,
                    int32_t nc, int32_t conv_out_form, float* partial, int32_t parts, const float* inv_scale, float* dw,
                    float* narrow_sum, void* stream) {
  DSG_CHECK_ARG(wide_h16 && narrow_nchw && partial && dw, "dsg_small_wgrad: null pointer");
  DSG_CHECK_ARG(nc >= 1 && nc <= 4 && wc >= 16 && wc <= 2 * SB_THREADS && (2 * SB_THREADS) % wc == 0,
                "dsg_small_wgrad: need 1 <= nc <= 4 and wc an even divisor of 512");
  DSG_CHECK_ARG((uintptr_t)wide_h16 % 4 == 0, "dsg_small_wgrad: unaligned pointer");
  DSG_CHECK_ARG(n >= 0 && h > 0 && w > 0 && parts >= 1, "dsg_small_wgrad: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int nslices = SB_THREADS / (wc / 2);
  const size_t sm_rows = (size_t)nc * 3 * (w + 8), sm_red = (size_t)nslices * nc * 9 * wc;
  const size_t sm = (sm_rows > sm_red ? sm_rows : sm_red) * sizeof(float);
  DSG_CHECK_ARG(sm <= 200 * 1024, "dsg_small_wgrad: row too wide for shared memory");
  const int sgn = conv_out_form ? -1 : 1;
#define DSG_SW_LAUNCH(NCV)                                                                                          \
  do {                                                                                                              \
    if (sm > 48 * 1024)                                                                                             \
      cudaFuncSetAttribute(small_wgrad_kernel<NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);          \
    small_wgrad_kernel<NCV><<<parts, SB_THREADS, sm, st>>>((const __half*)wide_h16, narrow_nchw, n, h, w, wc, sgn,  \
                                                           partial);                                                \
  } while (0)
  switch (nc) {
    case 1: DSG_SW_LAUNCH(1); break;
    case 2: DSG_SW_LAUNCH(2); break;
    case 3: DSG_SW_LAUNCH(3); break;
    default: DSG_SW_LAUNCH(4); break;
  }
#undef DSG_SW_LAUNCH
  DSG_CUDA_LAUNCH_CHECK("dsg_small_wgrad");
  // fixed-order sum over the CTAs' partials into the scratch row that follows them, then the layout change
  const int stride = nc * 9 * wc + nc;
  float* red = partial + (int64_t)parts * stride;
  launch_k(reduce_rows_kernel<1>, dim3(ceil_div(stride, 32)), dim3(256), 0, st, (const float*)partial, 1, parts, stride,
           (int64_t)0, (int64_t)stride, (float*)nullptr, 0, 0, inv_scale, red, (float*)nullptr, (float*)nullptr);
  DSG_CUDA_LAUNCH_CHECK("dsg_small_wgrad/reduce");
  small_wgrad_remap_kernel<<<ceil_div(stride, 256), 256, 0, st>>>(red, nc, wc, conv_out_form ? 0 : 1, dw, narrow_sum);
  DSG_CUDA_LAUNCH_CHECK("dsg_small_wgrad/finalize");
  return DSG_OK;
}
}
