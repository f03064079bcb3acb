// pack_weights.cu — every conv weight of the model re-packed to its fp16 GEMM layouts in ONE launch, staged through
// shared memory (training path: runs after every optimizer step; SURVEY.md §8 a17 — the per-layer fp32 -> fp16 cast that
// autocast performs for DriveSceneGen/scripts/train.py:24 `mixed_precision='fp16'`).
//
// The packed layouts put the input (forward) or output (data-gradient) channel innermost, the fp32 OIHW source has the
// 3 x 3 taps innermost: an element-per-thread gather (igemm.cu::pack_weight_kernel, kept as the single-job form and as
// the cross-check) reads one float per 36-byte stride (forward layouts) or per cin * 36 bytes (data-gradient layouts)
// and spends two 64-bit divisions per element — 0.55 ms for the 115 M packed elements of the reference U-Net.  Here a
// block loads a contiguous piece of the source into shared memory with coalesced reads and writes whole runs of the
// packed rows from it:
//   forward layouts (modes 0, 1, 2, 3)   one block per output channel: the channel's cin x KK source row (<= 36 KB);
//   data-gradient layouts (10 .. 13)     one block per (128 output channels x 8 input channels) tile — 288-byte source
//                                        runs, 256-byte destination runs (1x1: 64 input channels per tile).
// Same arithmetic, same order of additions as pack_value (bit-identical results; tests/test_gpu_igemm.py).
#include "common.cuh"

namespace dsg {

constexpr int PK_THREADS = 256;
constexpr int PK_CO_TILE = 128;
constexpr int PK_SMEM_MAX = 40 * 1024;   // per block: five blocks per SM; a 1024-channel 3 x 3 row + a 1024-channel shortcut row

__host__ __device__ __forceinline__ int pk_tci(int mode) { return mode == 13 ? 64 : 8; }
__host__ __device__ __forceinline__ int pk_kk(int mode) { return (mode == 3 || mode == 13) ? 1 : 9; }

// forward upsample conv (mode 2): sub-pixel phase (a, b), 2 x 2 tap (ti, tj) = the sum of the 3 x 3 taps that land on it
__device__ __forceinline__ float pk_up_sum(const float* w9, int a, int b, int ti, int tj) {
  float v = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int oy = (a + ky - 1) >= 0 ? (a + ky - 1) / 2 : -1;
    if (oy != ti + a - 1) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ox = (b + kx - 1) >= 0 ? (b + kx - 1) / 2 : -1;
      if (ox != tj + b - 1) continue;
      v += w9[ky * 3 + kx];
    }
  }
  return v;
}

__global__ void __launch_bounds__(PK_THREADS) pack_weights_tiled_kernel(const dsg_pack_job* __restrict__ jobs, int njobs) {
  extern __shared__ float pk_s[];
  int lo = 0, hi = njobs - 1;
  const int64_t blk = blockIdx.x;
  while (lo < hi) {   // last job whose first block is <= blk
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].chunk_begin <= blk) lo = mid; else hi = mid - 1;
  }
  const dsg_pack_job j = jobs[lo];
  const int b = (int)(blk - j.chunk_begin);
  const int tid = threadIdx.x;
  const int mode = j.mode, cin = j.cin, cout = j.cout;
  __half* out = (__half*)j.out;
  const int kt = (int)j.k_total;
  if (mode < 10) {
    // ---------------------------------------------------------------- forward layouts: block = output channel
    const int co = b, KK = pk_kk(mode), nsrc = cin * KK;
    const float* src = j.w + (int64_t)co * nsrc;
    for (int i = tid; i < nsrc; i += PK_THREADS) pk_s[i] = src[i];
    float* s_sc = pk_s + nsrc;
    for (int i = tid; i < j.csc; i += PK_THREADS) s_sc[i] = j.w_sc[(int64_t)co * j.csc + i];
    __syncthreads();
    if (mode == 0 || mode == 1) {
      __half* o = out + (int64_t)co * kt;
      for (int k = tid; k < kt; k += PK_THREADS) {
        float v;
        if (k < 9 * cin) {
          const int tap = k / cin, ci = k - tap * cin;
          v = pk_s[ci * 9 + tap];
        } else {
          v = s_sc[k - 9 * cin];
        }
        o[k] = __float2half_rn(v);
      }
    } else if (mode == 3) {
      __half* o = out + (int64_t)co * kt;
      for (int k = tid; k < kt; k += PK_THREADS) o[k] = __float2half_rn(pk_s[k]);
    } else {  // mode 2: rows (phase * cout + co), k = (ti * 2 + tj) * cin + ci
      for (int e = tid; e < 4 * kt; e += PK_THREADS) {
        const int phase = e / kt, k = e - phase * kt;
        const int tap = k / cin, ci = k - tap * cin;
        out[((int64_t)phase * cout + co) * kt + k] =
            __float2half_rn(pk_up_sum(pk_s + ci * 9, phase >> 1, phase & 1, tap >> 1, tap & 1));
      }
    }
    return;
  }
  // ------------------------------------------------------------------ data-gradient layouts: block = (ci tile, co tile)
  const int TCI = pk_tci(mode), KK = pk_kk(mode), run = TCI * KK, pitch = run + 1;
  const int co_tiles = (cout + PK_CO_TILE - 1) / PK_CO_TILE;
  const int ci0 = (b / co_tiles) * TCI, co0 = (b % co_tiles) * PK_CO_TILE;
  const int nci = min(TCI, cin - ci0), nco = min(PK_CO_TILE, cout - co0);
  const int nrun = nci * KK;
  for (int i = tid; i < nco * run; i += PK_THREADS) {
    const int col = i / run, e = i - col * run;
    pk_s[col * pitch + e] = e < nrun ? j.w[((int64_t)(co0 + col) * cin + ci0) * KK + e] : 0.f;
  }
  __syncthreads();
  if (mode == 10) {          // row ci, k = tap' * cout + co, spatially flipped taps (tap 8 - tap')
    for (int i = tid; i < nci * 9 * PK_CO_TILE; i += PK_THREADS) {
      const int col = i % PK_CO_TILE, r = i / PK_CO_TILE, tap = r % 9, cil = r / 9;
      if (col < nco)
        out[(int64_t)(ci0 + cil) * kt + (int64_t)tap * cout + co0 + col] =
            __float2half_rn(pk_s[col * pitch + cil * 9 + (8 - tap)]);
    }
  } else if (mode == 13) {   // transpose
    for (int i = tid; i < nci * PK_CO_TILE; i += PK_THREADS) {
      const int col = i % PK_CO_TILE, cil = i / PK_CO_TILE;
      if (col < nco) out[(int64_t)(ci0 + cil) * kt + co0 + col] = __float2half_rn(pk_s[col * pitch + cil]);
    }
  } else if (mode == 11) {   // dgrad of the stride-2 conv: row (pa * 2 + pb) * cin + ci, k = (2 x 2 tap) * cout + co
    for (int i = tid; i < nci * 16 * PK_CO_TILE; i += PK_THREADS) {
      const int col = i % PK_CO_TILE, r = i / PK_CO_TILE, tap = r % 4, phase = (r / 4) % 4, cil = r / 16;
      if (col >= nco) continue;
      const int pa = phase >> 1, pb = phase & 1;
      const int ky = 3 - 2 * (tap >> 1) - pa, kx = 3 - 2 * (tap & 1) - pb;
      float v = 0.f;
      if (ky >= 0 && ky < 3 && kx >= 0 && kx < 3) v = pk_s[col * pitch + cil * 9 + ky * 3 + kx];
      out[((int64_t)phase * cin + ci0 + cil) * kt + (int64_t)tap * cout + co0 + col] = __float2half_rn(v);
    }
  } else {                   // mode 12: dgrad of the upsample conv: row ci, k = ((a * 2 + b) * 4 + ti * 2 + tj) * cout + co
    for (int i = tid; i < nci * 16 * PK_CO_TILE; i += PK_THREADS) {
      const int col = i % PK_CO_TILE, r = i / PK_CO_TILE, e = r % 16, cil = r / 16;
      if (col >= nco) continue;
      out[(int64_t)(ci0 + cil) * kt + (int64_t)e * cout + co0 + col] =
          __float2half_rn(pk_up_sum(pk_s + col * pitch + cil * 9, e >> 3, (e >> 2) & 1, (e >> 1) & 1, e & 1));
    }
  }
}

static size_t pk_smem_bytes(int mode, int cin, int csc) {
  if (mode < 10) return ((size_t)cin * pk_kk(mode) + csc) * sizeof(float);
  return (size_t)PK_CO_TILE * (pk_tci(mode) * pk_kk(mode) + 1) * sizeof(float);
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int64_t dsg_pack_job_blocks(int32_t mode, int32_t cout, int32_t cin, int32_t csc) {
  if (!((mode >= 0 && mode <= 3) || (mode >= 10 && mode <= 13)) || cout <= 0 || cin <= 0 || csc < 0) return -1;
  if (pk_smem_bytes(mode, cin, csc) > PK_SMEM_MAX) return -1;   // use dsg_pack_conv_weight for such a layer
  if (mode < 10) return cout;
  return (int64_t)ceil_div(cin, pk_tci(mode)) * ceil_div(cout, PK_CO_TILE);
}

int dsg_pack_conv_weights_batched(const dsg_pack_job* jobs_dev, int32_t njobs, int64_t total_blocks, void* stream) {
  DSG_CHECK_ARG(jobs_dev && njobs >= 1 && total_blocks >= 1 && total_blocks < (int64_t)1 << 31,
                "dsg_pack_conv_weights_batched: bad args");
  static SmemAttrCache cache;
  cudaError_t e = ensure_dyn_smem(cache, pack_weights_tiled_kernel, PK_SMEM_MAX);
  if (e != cudaSuccess) {
    set_error("dsg_pack_conv_weights_batched: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return DSG_ERR_CUDA;
  }
  pack_weights_tiled_kernel<<<(unsigned)total_blocks, PK_THREADS, PK_SMEM_MAX, (cudaStream_t)stream>>>(jobs_dev, njobs);
  DSG_CUDA_LAUNCH_CHECK("dsg_pack_conv_weights_batched");
  return DSG_OK;
}
}
