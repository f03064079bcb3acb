// raster.cu — the byte-image kernels either side of the denoising path.
//
//   image_to_sample_kernel   uint8 HWC raster -> normalised fp32 NCHW sample: the arithmetic of
//                            Image_Dataset.__getitem__ (DriveSceneGen/utils/datasets/dataset.py:20-23,44-47:
//                            ToTensor = x / 255, Normalize([0.5], [0.5]) = (x - 0.5) / 0.5), done after a uint8 H2D copy
//   gray_hist_kernel +       get_gray_image (DriveSceneGen/vectorization/utils/image_utils.py:13-42): 256-bin histogram
//   gray_mask_kernel         of each colour channel, its first maximum, and the mask that is 0 where both the dx and the
//                            dy channel lie within `thresh` of their histogram peak, else 255
//   agent_threshold_kernel   the first three statements of extract_agents on the speed channel
//                            (DriveSceneGen/vectorization/direct/extract_vehicles.py:136-148): (x * 255) truncated to
//                            uint8, the grey conversion of three identical channels (= identity), threshold > t -> 255
//
// All three are HBM-bound byte kernels with integer / exactly-rounded arithmetic: results are bit-identical to the
// reference's numpy / torchvision / OpenCV path (tests/test_gpu_raster.py, tests/golden/raster_*.npz).
#include "common.cuh"

namespace dsg {

// ------------------------------------------------------------------------------------------------ image -> sample
// One thread per 4 pixels: c_img * 4 bytes in, one float4 per output plane.
template <int C_IMG>
__global__ void __launch_bounds__(256) image_to_sample_kernel(const uint8_t* __restrict__ img, float* __restrict__ out,
                                                              int c_out, int64_t hw, int64_t quads_per_image) {
  const int n = blockIdx.y;
  const uint8_t* src = img + (int64_t)n * hw * C_IMG;
  float* dst = out + (int64_t)n * c_out * hw;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int UQ = 4;   // quads per thread per iteration: UQ * C_IMG independent loads in flight
  for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q0 < quads_per_image; q0 += stride * UQ) {
    uint32_t words[UQ][C_IMG];
#pragma unroll
    for (int u = 0; u < UQ; ++u) {
      const int64_t q = q0 + u * stride;
      if (q < quads_per_image) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + q * 4 * C_IMG);
#pragma unroll
        for (int i = 0; i < C_IMG; ++i) words[u][i] = __ldg(wp + i);
      }
    }
#pragma unroll
    for (int u = 0; u < UQ; ++u) {
      const int64_t q = q0 + u * stride;
      if (q >= quads_per_image) break;
#pragma unroll
      for (int ch = 0; ch < C_IMG; ++ch) {
        if (ch >= c_out) break;
        float4 v;
        float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int b = p * C_IMG + ch;  // byte index inside the 4-pixel group
          const float x = (float)((words[u][b >> 2] >> ((b & 3) * 8)) & 0xffu);
          // ToTensor: x.to(float32).div(255); Normalize: sub(mean).div(std) -- three separately rounded fp32 operations
          vp[p] = __fdiv_rn(__fsub_rn(__fdiv_rn(x, 255.0f), 0.5f), 0.5f);
        }
        *reinterpret_cast<float4*>(dst + (int64_t)ch * hw + q * 4) = v;
      }
    }
  }
}

// scalar form for pixel counts that are not a multiple of 4 / unaligned bases
__global__ void __launch_bounds__(256) image_to_sample_scalar_kernel(const uint8_t* __restrict__ img,
                                                                     float* __restrict__ out, int c_img, int c_out,
                                                                     int64_t hw) {
  const int n = blockIdx.y;
  const uint8_t* src = img + (int64_t)n * hw * c_img;
  float* dst = out + (int64_t)n * c_out * hw;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += stride)
    for (int ch = 0; ch < c_out; ++ch)
      dst[(int64_t)ch * hw + p] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[p * c_img + ch], 255.0f), 0.5f), 0.5f);
}

// ------------------------------------------------------------------------------------------------ resize -> sample
// Resize((H, W), antialias=False) of Image_Dataset (dataset.py:20-23) between ToTensor and Normalize: torchvision's
// tensor resize is ATen's CPU upsample_bilinear2d (align_corners = False), restated bit for bit
// (oracle/raster.py::resize_bilinear, pinned by tests/golden/resize_golden.npz which torchvision produced):
//   src = max(fma(scale, i + 0.5, -0.5), 0), scale = float(in) / float(out); i0 = min(int(src), in - 1);
//   i1 = i0 + (i0 < in - 1); l1 = clamp(src - i0, 0, 1); l0 = 1 - l1; equal sizes: plain copy (weights 1, 0)
//   mode 0 (ATen's generic N-d kernel: the multi-threaded host path for outputs with H + W > 128, i.e. the reference's
//           512^2 -> 256^2):  out = fma(top, ly0, bot * ly1), top = fma(a, lx0, b * lx1), bot = fma(c, lx0, d * lx1)
//   mode 1 (ATen's channels-last kernel: single-threaded hosts with C == 3, and outputs with H + W <= 128):
//           w00 = ly0 * lx0 ...; out = fma(d, w11, fma(c, w10, fma(a, w00, b * w01)))
struct ResizeTap { int i0, i1; float l0, l1; };
__device__ __forceinline__ ResizeTap resize_tap(int o, int in_size, int out_size, float scale) {
  ResizeTap t;
  if (in_size == out_size) {
    t.i0 = t.i1 = o; t.l0 = 1.0f; t.l1 = 0.0f;
    return t;
  }
  const float src = fmaxf(__fmaf_rn(scale, __fadd_rn((float)o, 0.5f), -0.5f), 0.0f);
  t.i0 = min((int)src, in_size - 1);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = fminf(fmaxf(__fsub_rn(src, (float)t.i0), 0.0f), 1.0f);
  t.l0 = __fsub_rn(1.0f, t.l1);
  return t;
}
template <typename T> __device__ __forceinline__ float to_unit(T v);
template <> __device__ __forceinline__ float to_unit<uint8_t>(uint8_t v) { return __fdiv_rn((float)v, 255.0f); }  // ToTensor
template <> __device__ __forceinline__ float to_unit<float>(float v) { return v; }   // the .pkl branch holds floats

// One thread per output pixel (x fastest): 4 taps x c_img values in (rows y0 / y1 of the NHWC image), c_out coalesced
// plane stores out.  Grid (x blocks, out_h, n).
template <typename T>
__global__ void __launch_bounds__(256) resize_to_sample_kernel(const T* __restrict__ img, float* __restrict__ out,
                                                               int h, int w, int c_img, int c_out, int out_h, int out_w,
                                                               float scale_h, float scale_w, int mode) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y, n = blockIdx.z;
  if (ox >= out_w) return;
  const ResizeTap ty = resize_tap(oy, h, out_h, scale_h);
  const ResizeTap tx = resize_tap(ox, w, out_w, scale_w);
  const T* base = img + (int64_t)n * h * w * c_img;
  const T* r0 = base + (int64_t)ty.i0 * w * c_img;
  const T* r1 = base + (int64_t)ty.i1 * w * c_img;
  float* dst = out + ((int64_t)n * c_out * out_h + oy) * out_w + ox;
  const float w00 = __fmul_rn(ty.l0, tx.l0), w01 = __fmul_rn(ty.l0, tx.l1);
  const float w10 = __fmul_rn(ty.l1, tx.l0), w11 = __fmul_rn(ty.l1, tx.l1);
  for (int ch = 0; ch < c_out; ++ch) {
    const float a = to_unit<T>(r0[tx.i0 * c_img + ch]), b = to_unit<T>(r0[tx.i1 * c_img + ch]);
    const float c = to_unit<T>(r1[tx.i0 * c_img + ch]), d = to_unit<T>(r1[tx.i1 * c_img + ch]);
    float v;
    if (mode == 0) {
      const float top = __fmaf_rn(a, tx.l0, __fmul_rn(b, tx.l1));
      const float bot = __fmaf_rn(c, tx.l0, __fmul_rn(d, tx.l1));
      v = __fmaf_rn(top, ty.l0, __fmul_rn(bot, ty.l1));
    } else {
      v = __fmaf_rn(d, w11, __fmaf_rn(c, w10, __fmaf_rn(a, w00, __fmul_rn(b, w01))));
    }
    dst[(int64_t)ch * out_h * out_w] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);   // Normalize([0.5], [0.5])
  }
}

// ------------------------------------------------------------------------------------------------ histogram
// np.histogram(v / 255.0, bins=256, range=(0, 1)) for one byte value: the uniform-bin path computes
// floor((x - 0) / (1 - 0) * 256) in float64, moves 256 to 255 and then corrects against the edges i / 256 (exact).
__device__ __forceinline__ int np_hist_bin(int v) {
  const double x = (double)v / 255.0;
  int idx = (int)(x * 256.0);
  if (idx == 256) idx = 255;
  if (x < (double)idx / 256.0) --idx;
  if (idx != 255 && x >= (double)(idx + 1) / 256.0) ++idx;
  return idx;
}

constexpr int GH_THREADS = 256;
constexpr int GH_WARPS = GH_THREADS / 32;

// add `key` (channel * 256 + value, or -1 for nothing) to this warp's private histogram with one shared-memory atomic per
// distinct key in the warp: a BEV raster is mostly one background value, which would otherwise serialise 32 ways
__device__ __forceinline__ void warp_hist_add(uint32_t* wh, int key) {
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(wh + key, (uint32_t)__popc(peers));
}

// hist [n][3][256] must be zero on entry.  Grid (blocks per image, n).
// Vector form (C = 3 or 4, pixel count a multiple of 16): one thread takes 16 consecutive pixels (C 16-byte loads), so it
// holds 16 values of every channel.  Values equal to the thread's first value of that channel (the raster's background,
// for most threads) are counted in registers and added once, warp-aggregated; the others take one shared-memory atomic
// each on the warp's private histogram.
template <int C>
__global__ void __launch_bounds__(GH_THREADS) gray_hist_kernel(const uint8_t* __restrict__ img,
                                                               uint32_t* __restrict__ hist, int c_rt, int64_t hw,
                                                               int vec_ok) {
  __shared__ uint32_t sh[GH_WARPS][3 * 256];
  __shared__ int bin_of[256];
  const int n = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < GH_WARPS * 3 * 256; i += GH_THREADS) (&sh[0][0])[i] = 0;
  bin_of[tid] = np_hist_bin(tid);
  __syncthreads();
  uint32_t* wh = sh[warp];
  const int c = C ? C : c_rt;
  const uint8_t* src = img + (int64_t)n * hw * c;
  const int64_t per_pass = (int64_t)gridDim.x * GH_THREADS;
  if (C != 0 && vec_ok) {
    constexpr int CC = C ? C : 1;
    const int64_t groups = hw / 16;  // 16-pixel groups; trip count is warp-uniform (match_any inside)
    for (int64_t base = (int64_t)blockIdx.x * GH_THREADS; base < groups; base += per_pass) {
      const int64_t g = base + tid;
      const bool live = g < groups;
      uint32_t words[4 * CC];
      if (live) {
        const uint4* vp = reinterpret_cast<const uint4*>(src + g * 16 * CC);
#pragma unroll
        for (int i = 0; i < CC; ++i) {
          const uint4 q = __ldg(vp + i);
          words[4 * i] = q.x, words[4 * i + 1] = q.y, words[4 * i + 2] = q.z, words[4 * i + 3] = q.w;
        }
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        int first = -1, same = 0;
        if (live) {
          first = (words[ch >> 2] >> ((ch & 3) * 8)) & 0xff;
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            const int b = p * CC + ch;
            const int val = (words[b >> 2] >> ((b & 3) * 8)) & 0xff;
            if (val == first) ++same;
            else atomicAdd(wh + ch * 256 + val, 1u);
          }
        }
        const int key = live ? ch * 256 + first : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        // the first lane of each group of equal keys adds the group's total
        const int total = __reduce_add_sync(peers, same);
        if (live && (int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(wh + key, (uint32_t)total);
      }
    }
  } else {
    const int64_t bytes = hw * c;
    for (int64_t base = (int64_t)blockIdx.x * GH_THREADS; base < bytes; base += per_pass) {
      const int64_t b = base + tid;
      const bool live = b < bytes;
      const int ch = live ? (int)(b % c) : 0;
      warp_hist_add(wh, (live && ch < 3) ? ch * 256 + (int)src[b] : -1);
    }
  }
  __syncthreads();
  // fold the warps, map byte value -> numpy bin, one global atomic per non-empty (channel, value)
  for (int i = tid; i < 3 * 256; i += GH_THREADS) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < GH_WARPS; ++w) s += sh[w][i];
    if (s) atomicAdd(hist + ((int64_t)n * 3 + (i >> 8)) * 256 + bin_of[i & 255], s);
  }
}

// ------------------------------------------------------------------------------------------------ mask
// Grid (blocks per image, n).  Prologue: np.argmax of each channel's histogram (first maximum) and, per byte value, whether
// |v / 255 - peak / 256| <= thresh in float64 (combine_dx_dy); then 4 pixels per thread.
__global__ void __launch_bounds__(256) gray_mask_kernel(const uint8_t* __restrict__ img,
                                                        const uint32_t* __restrict__ hist, int32_t* __restrict__ peaks,
                                                        uint8_t* __restrict__ mask, uint8_t* __restrict__ gray3, int c,
                                                        int64_t hw, double thresh, int vec_ok) {
  __shared__ unsigned long long best[3][8];
  __shared__ uint8_t near_peak[2][256];
  const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    // larger count wins; on ties the smaller index wins
    unsigned long long key = ((unsigned long long)hist[((int64_t)n * 3 + ch) * 256 + tid] << 8) | (unsigned)(255 - tid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    if (lane == 0) best[ch][warp] = key;
  }
  __syncthreads();
  int peak[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    unsigned long long key = best[ch][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) key = best[ch][w] > key ? best[ch][w] : key;
    peak[ch] = 255 - (int)(key & 0xff);
  }
  if (blockIdx.x == 0 && tid < 3) peaks[n * 3 + tid] = peak[tid];
  {
    const double x = (double)tid / 255.0;
    near_peak[0][tid] = fabs(x - (double)peak[0] / 256.0) <= thresh;
    near_peak[1][tid] = fabs(x - (double)peak[1] / 256.0) <= thresh;
  }
  __syncthreads();
  const uint8_t* src = img + (int64_t)n * hw * c;
  uint8_t* m = mask + (int64_t)n * hw;
  uint8_t* g3 = gray3 ? gray3 + (int64_t)n * hw * 3 : nullptr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t quads = hw / 4;
    constexpr int UQ = 4;
    for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x + tid; q0 < quads; q0 += stride * UQ) {
      uint32_t words[UQ][4];
#pragma unroll
      for (int u = 0; u < UQ; ++u) {
        const int64_t q = q0 + u * stride;
        if (q < quads) {
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + q * 4 * c);
          words[u][0] = __ldg(wp); words[u][1] = __ldg(wp + 1); words[u][2] = __ldg(wp + 2);
          words[u][3] = c == 4 ? __ldg(wp + 3) : 0u;
        }
      }
#pragma unroll
      for (int u = 0; u < UQ; ++u) {
        const int64_t q = q0 + u * stride;
        if (q >= quads) break;
        uint32_t out = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int b0 = p * c, b1 = p * c + 1;
          const int r = (words[u][b0 >> 2] >> ((b0 & 3) * 8)) & 0xff;
          const int g = (words[u][b1 >> 2] >> ((b1 & 3) * 8)) & 0xff;
          const uint32_t v = (near_peak[0][r] & near_peak[1][g]) ? 0u : 255u;
          out |= v << (p * 8);
        }
        *reinterpret_cast<uint32_t*>(m + q * 4) = out;
        if (g3) {
          // 4 pixels x 3 identical channels = 12 bytes: p0 p0 p0 p1 | p1 p1 p2 p2 | p2 p3 p3 p3
          const uint32_t p0 = out & 0xff, p1 = (out >> 8) & 0xff, p2 = (out >> 16) & 0xff, p3 = out >> 24;
          uint32_t* gp = reinterpret_cast<uint32_t*>(g3 + q * 12);
          gp[0] = p0 | (p0 << 8) | (p0 << 16) | (p1 << 24);
          gp[1] = p1 | (p1 << 8) | (p2 << 16) | (p2 << 24);
          gp[2] = p2 | (p3 << 8) | (p3 << 16) | (p3 << 24);
        }
      }
    }
  } else {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + tid; p < hw; p += stride) {
      const uint8_t v = (near_peak[0][src[p * c]] & near_peak[1][src[p * c + 1]]) ? 0 : 255;
      m[p] = v;
      if (g3) g3[p * 3] = g3[p * 3 + 1] = g3[p * 3 + 2] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ agent threshold
// (x * 255) in fp32 -> astype(uint8) truncates toward zero (values are in [0, 1], so no wrap); cv2.cvtColor(BGR2GRAY) of
// three equal channels v is (v * (1868 + 9617 + 4899) + 8192) >> 14 = v; cv2.threshold(gray, t, 255, THRESH_BINARY).
__global__ void __launch_bounds__(256) agent_threshold_kernel(const float* __restrict__ plane, int64_t plane_stride,
                                                              uint8_t* __restrict__ out, int64_t hw, int thresh,
                                                              int vec_ok) {
  const int n = blockIdx.y;
  const float* src = plane + (int64_t)n * plane_stride;
  uint8_t* dst = out + (int64_t)n * hw;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto one = [thresh](float x) -> uint32_t {
    float t = __fmul_rn(x, 255.0f);
    t = fminf(fmaxf(t, 0.0f), 255.0f);
    const int v = (int)t;  // truncation
    const int gray = (v * 16384 + 8192) >> 14;
    return gray > thresh ? 255u : 0u;
  };
  if (vec_ok) {
    const int64_t quads = hw / 4;
    constexpr int UQ = 4;
    for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q0 < quads; q0 += stride * UQ) {
      float4 v[UQ];
#pragma unroll
      for (int u = 0; u < UQ; ++u)
        if (q0 + u * stride < quads) v[u] = __ldg(reinterpret_cast<const float4*>(src) + q0 + u * stride);
#pragma unroll
      for (int u = 0; u < UQ; ++u)
        if (q0 + u * stride < quads)
          *reinterpret_cast<uint32_t*>(dst + (q0 + u * stride) * 4) =
              one(v[u].x) | (one(v[u].y) << 8) | (one(v[u].z) << 16) | (one(v[u].w) << 24);
    }
  } else {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += stride) dst[p] = (uint8_t)one(src[p]);
  }
}

static inline unsigned blocks_for(int64_t items, int threads, int n_images) {
  // a few waves of 148 SMs over the whole batch
  int64_t want = ceil_div64(items, threads);
  int64_t cap = ceil_div64(148 * 8, n_images > 0 ? n_images : 1);
  if (want > cap) want = cap;
  return (unsigned)(want < 1 ? 1 : want);
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_image_to_sample(const uint8_t* img, float* out, int32_t n, int32_t h, int32_t w, int32_t c_img, int32_t c_out,
                        void* stream) {
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && h >= 0 && w >= 0, "dsg_image_to_sample: bad shape");
  DSG_CHECK_ARG(c_img >= 1 && c_img <= 4 && c_out >= 1 && c_out <= c_img, "dsg_image_to_sample: 1 <= c_out <= c_img <= 4");
  const int64_t hw = (int64_t)h * w;
  if (n == 0 || hw == 0) return DSG_OK;
  DSG_CHECK_ARG(img && out, "dsg_image_to_sample: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = hw % 4 == 0 && (uintptr_t)img % 4 == 0 && (uintptr_t)out % 16 == 0;
  if (vec) {
    dim3 grid(blocks_for(hw / 4, 256, n), n);
    switch (c_img) {
      case 1: image_to_sample_kernel<1><<<grid, 256, 0, st>>>(img, out, c_out, hw, hw / 4); break;
      case 2: image_to_sample_kernel<2><<<grid, 256, 0, st>>>(img, out, c_out, hw, hw / 4); break;
      case 3: image_to_sample_kernel<3><<<grid, 256, 0, st>>>(img, out, c_out, hw, hw / 4); break;
      default: image_to_sample_kernel<4><<<grid, 256, 0, st>>>(img, out, c_out, hw, hw / 4); break;
    }
  } else {
    image_to_sample_scalar_kernel<<<dim3(blocks_for(hw, 256, n), n), 256, 0, st>>>(img, out, c_img, c_out, hw);
  }
  DSG_CUDA_LAUNCH_CHECK("dsg_image_to_sample");
  return DSG_OK;
}

int dsg_resize_to_sample(const void* img, int32_t img_is_f32, float* out, int32_t n, int32_t h, int32_t w, int32_t c_img,
                         int32_t c_out, int32_t out_h, int32_t out_w, int32_t mode, void* stream) {
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && h > 0 && w > 0 && out_h > 0 && out_h <= 65535 && out_w > 0,
                "dsg_resize_to_sample: bad shape");
  DSG_CHECK_ARG(c_img >= 1 && c_out >= 1 && c_out <= c_img, "dsg_resize_to_sample: 1 <= c_out <= c_img");
  DSG_CHECK_ARG(mode == 0 || mode == 1, "dsg_resize_to_sample: mode is 0 (generic) or 1 (channels-last formula)");
  if (n == 0) return DSG_OK;
  DSG_CHECK_ARG(img && out, "dsg_resize_to_sample: null pointer");
  const float scale_h = (float)h / (float)out_h, scale_w = (float)w / (float)out_w;
  dim3 grid((unsigned)ceil_div(out_w, 256), (unsigned)out_h, (unsigned)n);
  const int threads = out_w >= 256 ? 256 : ((out_w + 31) / 32) * 32;
  if (img_is_f32)
    resize_to_sample_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>((const float*)img, out, h, w, c_img, c_out,
                                                                                out_h, out_w, scale_h, scale_w, mode);
  else
    resize_to_sample_kernel<uint8_t><<<grid, threads, 0, (cudaStream_t)stream>>>((const uint8_t*)img, out, h, w, c_img,
                                                                                  c_out, out_h, out_w, scale_h, scale_w,
                                                                                  mode);
  DSG_CUDA_LAUNCH_CHECK("dsg_resize_to_sample");
  return DSG_OK;
}

int dsg_gray_mask(const uint8_t* img, uint32_t* hist, int32_t* peaks, uint8_t* mask, uint8_t* gray3, int32_t n,
                  int32_t h, int32_t w, int32_t c, double thresh, void* stream) {
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && h >= 0 && w >= 0, "dsg_gray_mask: bad shape");
  DSG_CHECK_ARG(c == 3 || c == 4, "dsg_gray_mask: images must have 3 or 4 channels");
  if (n == 0) return DSG_OK;
  const int64_t hw = (int64_t)h * w;
  DSG_CHECK_ARG(hist && peaks && (hw == 0 || (img && mask)), "dsg_gray_mask: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)n * 3 * 256 * sizeof(uint32_t), st);
  if (e != cudaSuccess) {
    set_error("dsg_gray_mask: memset failed: %s", cudaGetErrorString(e));
    return DSG_ERR_CUDA;
  }
  if (hw > 0) {
    const int vec_h = hw % 16 == 0 && (uintptr_t)img % 16 == 0;
    if (vec_h) {
      dim3 grid(blocks_for(hw / 16, GH_THREADS, n), n);
      if (c == 3) gray_hist_kernel<3><<<grid, GH_THREADS, 0, st>>>(img, hist, c, hw, 1);
      else gray_hist_kernel<4><<<grid, GH_THREADS, 0, st>>>(img, hist, c, hw, 1);
    } else {
      gray_hist_kernel<0><<<dim3(blocks_for(hw * c, GH_THREADS, n), n), GH_THREADS, 0, st>>>(img, hist, c, hw, 0);
    }
    DSG_CUDA_LAUNCH_CHECK("dsg_gray_mask/hist");
  }
  // with no pixels every histogram is zero and np.argmax gives bin 0; the mask kernel still writes the peaks
  const int vec_m = hw % 4 == 0 && (uintptr_t)img % 4 == 0 && (uintptr_t)mask % 4 == 0 && (uintptr_t)gray3 % 4 == 0;
  gray_mask_kernel<<<dim3(blocks_for(hw / 4 + 1, 256, n), n), 256, 0, st>>>(img, hist, peaks, mask, gray3, c, hw, thresh,
                                                                           vec_m);
  DSG_CUDA_LAUNCH_CHECK("dsg_gray_mask/mask");
  return DSG_OK;
}

int dsg_agent_threshold(const float* plane, int64_t plane_stride, uint8_t* out, int32_t n, int64_t hw, int32_t thresh,
                        void* stream) {
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && hw >= 0 && plane_stride >= 0, "dsg_agent_threshold: bad shape");
  if (n == 0 || hw == 0) return DSG_OK;
  DSG_CHECK_ARG(plane && out, "dsg_agent_threshold: null pointer");
  const int vec = hw % 4 == 0 && plane_stride % 4 == 0 && (uintptr_t)plane % 16 == 0 && (uintptr_t)out % 4 == 0;
  agent_threshold_kernel<<<dim3(blocks_for(hw / 4 + 1, 256, n), n), 256, 0, (cudaStream_t)stream>>>(
      plane, plane_stride, out, hw, thresh, vec);
  DSG_CUDA_LAUNCH_CHECK("dsg_agent_threshold");
  return DSG_OK;
}
}
