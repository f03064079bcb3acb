// conv_small.cu — the two degenerate convolutions of the U-Net (K = 27 and N = 3): CUDA-core, HBM-bound.
// Replaces UNet2DModel.conv_in (3->C0, NCHW fp32 in, NHWC fp16 out) and conv_out (C0->3, NHWC fp16 in,
// NCHW fp32 out) of diffusers 0.20.0 models/unet_2d.py.
#include "common.cuh"

namespace dsg {

constexpr int CS_THREADS = 256;

// thread = (4 adjacent pixels of one image row, 8 output channels); persistent blocks stage the weights once.
// Weights in smem as [k = ci*9 + ky*3 + kx][cout]: one pair of LDS.128 feeds 32 FMAs.
constexpr int CI_PX = 4;
__global__ void __launch_bounds__(CS_THREADS, 2) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, __half* __restrict__ out,
                                                             int n, int cin, int h, int wd, int cout) {
  extern __shared__ float sw[];  // [cin*9][cout] then bias[cout]
  const int K = cin * 9;
  pdl_sync();
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) {
    const int k = i / cout, co = i - k * cout;  // w is [cout][cin][3][3] = [cout][K]
    sw[i] = w[co * K + k];
  }
  float* sb = sw + K * cout;
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sb[i] = b[i];
  __syncthreads();
  const int gpp = cout >> 3;                 // channel groups per pixel
  const int qpb = CS_THREADS / gpp;          // pixel quads per block iteration
  const int cg = threadIdx.x % gpp, ql = threadIdx.x / gpp;
  if (ql >= qpb) return;
  const int wq = (wd + CI_PX - 1) / CI_PX;   // quads per image row
  const int64_t total_q = (int64_t)n * h * wq;
  const int64_t hw = (int64_t)h * wd;
  float bias8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias8[j] = sb[cg * 8 + j];
  for (int64_t qi = (int64_t)blockIdx.x * qpb + ql; qi < total_q; qi += (int64_t)gridDim.x * qpb) {
    const int xq = (int)(qi % wq);
    const int64_t r = qi / wq;
    const int y = (int)(r % h), nn = (int)(r / h);
    const int x0 = xq * CI_PX;
    float acc[CI_PX][8];
#pragma unroll
    for (int px = 0; px < CI_PX; ++px)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[px][j] = bias8[j];
    for (int ci = 0; ci < cin; ++ci) {
      const float* xp = x + ((int64_t)nn * cin + ci) * hw;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= h) continue;
        float v[CI_PX + 2];
#pragma unroll
        for (int u = 0; u < CI_PX + 2; ++u) {
          const int xc = x0 + u - 1;
          v[u] = (xc >= 0 && xc < wd) ? __ldg(xp + (int64_t)yy * wd + xc) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wr = sw + (ci * 9 + ky * 3 + kx) * cout + cg * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
#pragma unroll
          for (int px = 0; px < CI_PX; ++px) {
            const float a = v[px + kx];
            acc[px][0] = fmaf(a, w0.x, acc[px][0]); acc[px][1] = fmaf(a, w0.y, acc[px][1]);
            acc[px][2] = fmaf(a, w0.z, acc[px][2]); acc[px][3] = fmaf(a, w0.w, acc[px][3]);
            acc[px][4] = fmaf(a, w1.x, acc[px][4]); acc[px][5] = fmaf(a, w1.y, acc[px][5]);
            acc[px][6] = fmaf(a, w1.z, acc[px][6]); acc[px][7] = fmaf(a, w1.w, acc[px][7]);
          }
        }
      }
    }
    __half* op = out + (((int64_t)nn * h + y) * wd + x0) * cout + cg * 8;
#pragma unroll
    for (int px = 0; px < CI_PX; ++px)
      if (x0 + px < wd) stg_v4(op + (int64_t)px * cout, pack8(acc[px]));
  }
}

// warp = 8 adjacent pixels x 4 lanes; each lane owns cin/4 input channels; lanes reduced by shuffle.
// Weights in smem as [tap][co][cin].
template <int MAXCO>
__global__ void __launch_bounds__(CS_THREADS) conv_out_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, float* __restrict__ out,
                                                              int n, int cin, int h, int wd, int cout) {
  extern __shared__ float sw[];  // [9][cout][cin]
  for (int i = threadIdx.x; i < cout * cin * 9; i += blockDim.x) {
    // w is [cout][cin][3][3]
    const int co = i / (cin * 9), r = i - co * cin * 9, ci = r / 9, tap = r - ci * 9;
    sw[(tap * cout + co) * cin + ci] = w[i];
  }
  __syncthreads();
  const int lane4 = threadIdx.x & 3;
  const int64_t hw = (int64_t)h * wd, total = hw * n;
  const int64_t p = (int64_t)blockIdx.x * (CS_THREADS / 4) + (threadIdx.x >> 2);
  const bool valid = p < total;
  const int64_t pc = valid ? p : 0;
  const int nn = (int)(pc / hw);
  const int rem = (int)(pc - (int64_t)nn * hw);
  const int y = rem / wd, xx = rem - y * wd;
  const int cpl = cin >> 2;  // channels per lane (multiple of 8)
  const int c_lo = lane4 * cpl;
  float acc[MAXCO];
#pragma unroll
  for (int j = 0; j < MAXCO; ++j) acc[j] = 0.f;
  if (valid) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xc = xx + kx - 1;
        if (xc < 0 || xc >= wd) continue;
        const __half* xp = x + (((int64_t)nn * h + yy) * wd + xc) * cin + c_lo;
        const float* wt = sw + (ky * 3 + kx) * cout * cin + c_lo;
        for (int c8 = 0; c8 < cpl; c8 += 8) {
          float f[8];
          unpack8(*reinterpret_cast<const uint4*>(xp + c8), f);
#pragma unroll
          for (int j = 0; j < MAXCO; ++j) {
            if (j < cout) {
              const float4 w0 = *reinterpret_cast<const float4*>(wt + j * cin + c8);
              const float4 w1 = *reinterpret_cast<const float4*>(wt + j * cin + c8 + 4);
              acc[j] = fmaf(f[0], w0.x, acc[j]); acc[j] = fmaf(f[1], w0.y, acc[j]);
              acc[j] = fmaf(f[2], w0.z, acc[j]); acc[j] = fmaf(f[3], w0.w, acc[j]);
              acc[j] = fmaf(f[4], w1.x, acc[j]); acc[j] = fmaf(f[5], w1.y, acc[j]);
              acc[j] = fmaf(f[6], w1.z, acc[j]); acc[j] = fmaf(f[7], w1.w, acc[j]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXCO; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
  }
  if (valid && lane4 == 0) {
#pragma unroll
    for (int j = 0; j < MAXCO; ++j)
      if (j < cout) out[((int64_t)nn * cout + j) * hw + rem] = acc[j] + b[j];
  }
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_conv_in(const float* x, const float* w, const float* b, void* out_h16, int32_t n, int32_t cin, int32_t h,
                int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x && w && b && out_h16, "dsg_conv_in: null pointer");
  DSG_CHECK_ARG(cin >= 1 && cin <= 4 && cout % 8 == 0 && cout >= 8 && cout <= 512, "dsg_conv_in: cin<=4, cout%%8==0");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_in: bad shape");
  DSG_CHECK_ARG((uintptr_t)out_h16 % 16 == 0, "dsg_conv_in: out must be 16-byte aligned");
  if (n == 0) return DSG_OK;
  const int gpp = cout / 8, qpb = CS_THREADS / gpp;
  const int64_t total_q = (int64_t)n * h * ((wd + CI_PX - 1) / CI_PX);
  const size_t sm = (size_t)(cin * 9 * cout + cout) * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  int64_t blocks = ceil_div64(total_q, qpb);
  if (blocks > 148 * 2) blocks = 148 * 2;  // persistent: the weight staging is paid once per block
  launch_k(conv_in_kernel, dim3((unsigned)blocks), dim3(CS_THREADS), sm, (cudaStream_t)stream, x, w, b,
           (__half*)out_h16, n, cin, h, wd, cout);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_in");
  return DSG_OK;
}

int dsg_conv_out(const void* x_h16, const float* w, const float* b, float* out, int32_t n, int32_t cin, int32_t h,
                 int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x_h16 && w && b && out, "dsg_conv_out: null pointer");
  DSG_CHECK_ARG(cout >= 1 && cout <= 4 && cin % 32 == 0 && cin >= 32 && cin <= 512,
                "dsg_conv_out: cout<=4, cin%%32==0");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_out: bad shape");
  DSG_CHECK_ARG((uintptr_t)x_h16 % 16 == 0, "dsg_conv_out: x must be 16-byte aligned");
  if (n == 0) return DSG_OK;
  const int64_t total = (int64_t)n * h * wd;
  const size_t sm = (size_t)9 * cout * cin * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(conv_out_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  conv_out_kernel<4><<<(unsigned)ceil_div64(total, CS_THREADS / 4), CS_THREADS, sm, (cudaStream_t)stream>>>(
      (const __half*)x_h16, w, b, out, n, cin, h, wd, cout);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_out");
  return DSG_OK;
}
}
