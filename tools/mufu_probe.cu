// mufu_probe.cu — issue rate of the MUFU (SFU) pipe on B200 for the forms the kernels use:
//   ex2.approx.ftz.f32, ex2.approx.ftz.f16x2, tanh.approx.f32, tanh.approx.f16x2
// Each thread runs 8 independent dependency chains; result = ops per clock per SM (elements, i.e. f16x2 counts 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_probe tools/mufu_probe.cu && tools/mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t* out, int iters, long long* clocks) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0x3c003800u + threadIdx.x + j;   // f32 / f16x2 bit patterns near 1
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[j]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[j]));
      if (MODE == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(v[j]));
      if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(v[j]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

int main() {
  uint32_t* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 4096;
  const char* names[4] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "tanh.approx.f32", "tanh.approx.f16x2"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) probe<0><<<148, 1024>>>(out, iters, clk);
      if (mode == 1) probe<1><<<148, 1024>>>(out, iters, clk);
      if (mode == 2) probe<2><<<148, 1024>>>(out, iters, clk);
      if (mode == 3) probe<3><<<148, 1024>>>(out, iters, clk);
      cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double instr = 1024.0 * iters * 8;           // thread-level instructions per SM
    const int per = (mode == 1 || mode == 3) ? 2 : 1;
    printf("%-22s %8.2f instr/clk/SM  %8.2f elements/clk/SM\n", names[mode], instr / c, per * instr / c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
