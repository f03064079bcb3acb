// umma_probe.cu — stand-alone sm_100a micro-probe (measurement tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu -lcuda
// 1. tcgen05.mma issue/execute rate (cycles per instruction) for M=128, K=16, N in {16..256}, A from smem (SS) and
//    A from TMEM (TS), one CTA per SM on every SM — tells which conv tile shapes are shared-memory-read bound.
// 2. Whether a 128B-swizzled K-major A descriptor may start at a row that is NOT a multiple of 8 (start address
//    + r * 128 B), with the descriptor base_offset field either 0 or (addr >> 7) & 7 — this would let one TMA box
//    serve all nine taps of a 3x3 conv.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../drivescenegen_b200/csrc/common.cuh"

namespace dsg {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace dsg
using namespace dsg;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------- 1. rate probe
// mode 0: SS (A, B from smem).  mode 1: TS (A from TMEM columns 256.., B from smem).
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, int mode, int nbuf, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  // zero the operand area (A: nbuf x 16 KB, B: nbuf x 32 KB)
  for (int i = threadIdx.x; i < nbuf * (16384 + 32768) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_holder);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_holder, 0);
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + nbuf * 16384);
    long long t0 = clock64(), t1 = 0;
    if (elect_one_sync()) {
      for (int it = 0; it < iters; ++it) {
        const int buf = it % nbuf;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t db = umma_desc_sw128(b0 + buf * 32768) + (uint64_t)(2 * k);
          if (mode == 0) {
            const uint64_t da = umma_desc_sw128(a0 + buf * 16384) + (uint64_t)(2 * k);
            umma_f16(tmem_base, da, db, idesc, 1u);
          } else {
            umma_f16_ts(tmem_base, tmem_base + 256 + 8 * k, db, idesc, 1u);
          }
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// ---------------------------------------------------------------- 2. unaligned-start probe
// A region: 24 rows of 64 fp16, value(row, col) = row + col / 64, stored in the TMA SWIZZLE_128B pattern
// (16-byte chunk c of row r at chunk c ^ (r & 7)).  B = 16 x 64 with B[n][k] = (k == n) -> D[m][n] = A[m + r][n].
// Only rows 0..7 of D are checked (M = 128 reads past the initialised rows; zeros there).
__global__ void __launch_bounds__(128, 1) shift_kernel(int r, int use_base_offset, float* out /* [128][16] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  __half* A = reinterpret_cast<__half*>(smem);            // 160 rows x 128 B = 20 KB
  __half* B = reinterpret_cast<__half*>(smem + 20480);    // 16 rows x 128 B
  for (int i = threadIdx.x; i < (20480 + 2048) / 2; i += blockDim.x) A[i] = __float2half(0.f);
  __syncthreads();
  for (int i = threadIdx.x; i < 160 * 64; i += blockDim.x) {
    const int row = i / 64, col = i % 64;
    const int chunk = (col / 8) ^ (row & 7);
    A[row * 64 + chunk * 8 + (col % 8)] = __float2half((float)row + (float)col / 64.f);
  }
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int row = i / 64, col = i % 64;
    const int chunk = (col / 8) ^ (row & 7);
    B[row * 64 + chunk * 8 + (col % 8)] = __float2half(col == row ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<32>(&tmem_holder);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_holder, 0);
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t addr = smem_u32(smem) + (uint32_t)r * 128u;
      uint64_t da = umma_desc_sw128(addr);
      if (use_base_offset) da |= (uint64_t)((addr >> 7) & 7) << 49;
      const uint64_t db = umma_desc_sw128(smem_u32(smem + 20480));
      umma_f16(tmem_base, da, db, umma_idesc_f16(16), 0u);  // k = 0..15 only: D[m][n] = A[m + r][n], n < 16
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[16];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(tmem_base); }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  printf("SMs %d, clock %d kHz\n", sms, khz);
  long long* d_cyc;
  cudaMalloc(&d_cyc, sizeof(long long) * sms);
  const int smem = 2 * (16384 + 32768) + 2048;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int grid : {1, sms}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int nbuf : {1, 2}) {
        for (int n : {16, 32, 64, 96, 128, 192, 256}) {
          rate_kernel<<<grid, 128, smem>>>(n, iters, mode, nbuf, d_cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("rate_kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> h(grid);
          cudaMemcpy(h.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (auto c : h) mx = c > mx ? c : mx;
          const double per = (double)mx / (iters * 4);
          printf("rate grid=%3d %s nbuf=%d N=%3d : %7.1f cycles/MMA  (math floor %5.1f, smem bytes/MMA %5d -> %6.1f B/clk)\n",
                 grid, mode ? "TS" : "SS", nbuf, n, per, n / 2.0, (mode ? 0 : 4096) + n * 32,
                 ((mode ? 0 : 4096) + n * 32) / per);
        }
      }
    }
  }
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * sizeof(float));
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int ubo = 0; ubo < 2; ++ubo) {
    for (int r = 0; r < 10; ++r) {
      cudaMemset(d_out, 0, 128 * 16 * sizeof(float));
      shift_kernel<<<1, 128, 32768>>>(r, ubo, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift_kernel r=%d failed: %s\n", r, cudaGetErrorString(e)); return 1; }
      std::vector<float> h(128 * 16);
      cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
          const float want = (float)(m + r) + (float)n / 64.f;
          if (fabsf(h[m * 16 + n] - want) > 0.02f) ++bad;
        }
      printf("shift r=%d base_offset=%d : %s (%d mismatches)  D[0][0..3] = %.3f %.3f %.3f %.3f  D[9][1] = %.3f\n", r, ubo,
             bad ? "WRONG" : "ok", bad, h[0], h[1], h[2], h[3], h[9 * 16 + 1]);
    }
  }
  return 0;
}
