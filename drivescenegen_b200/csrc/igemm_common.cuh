// igemm_common.cuh — pieces shared by the two implicit-GEMM kernels (igemm.cu: tap-streaming; igemm_halo.cu: halo-reuse).
#pragma once
#include "common.cuh"

namespace dsg {

constexpr int IG_BLOCK_M = 128;
constexpr int IG_BLOCK_K = 64;
constexpr int IG_MAX_SRC = 4;

struct IgSrc {
  const __half* ptr;
  int64_t sN, sH, sW;  // element strides
  int C, H, W;         // logical extents (out-of-range reads are zero)
};

// ------------------------------------------------------------------ GroupNorm statistics in the conv epilogue
// Each epilogue warp owns 32 accumulator rows.  Per 32-column slice it reduces its rows to per-column sum / sum of
// squares (warp butterfly) and keeps them in its private quarter of sstat[4][BLOCK_N][2]; after the tile the four
// quarters are combined in a fixed order and added to the tensor's per-channel int64 totals (groupnorm.cu).
__device__ __forceinline__ void epi_stats_slice(const float* f, float* sstat_warp_slice /* [32][2] */, int lane,
                                                bool first) {
  float t[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = f[j];
  const float s1 = warp_colsum32(t, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = f[j] * f[j];
  const float s2 = warp_colsum32(t, lane);
  float2* slot = reinterpret_cast<float2*>(sstat_warp_slice) + lane;
  if (first) {
    *slot = make_float2(s1, s2);
  } else {
    float2 o = *slot;
    *slot = make_float2(o.x + s1, o.y + s2);
  }
}
template <int BLOCK_N>
__device__ __forceinline__ void epi_stats_flush(const float* sstat /* [4][BLOCK_N][2] */, int te /* 0..127 */,
                                                long long* stats_nc /* &stats[(n * cout + n0) * 2] */) {
  for (int idx = te; idx < 2 * BLOCK_N; idx += 128) {
    const float v = ((sstat[idx] + sstat[2 * BLOCK_N + idx]) + sstat[4 * BLOCK_N + idx]) + sstat[6 * BLOCK_N + idx];
    const long long fx = (idx & 1) ? gn_fix_sq(v) : gn_fix_sum(v);
    atomicAdd(reinterpret_cast<unsigned long long*>(stats_nc + idx), (unsigned long long)fx);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

inline int make_map_a(CUtensorMap* m, const IgSrc& s, int N, int TW, int box_h) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DSG_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)s.sW * 2, (cuuint64_t)s.sH * 2, (cuuint64_t)s.sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)IG_BLOCK_K, (cuuint32_t)TW, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)s.ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d N=%d TW=%d TH=%d)", (int)r, s.C, s.W, s.H, N, TW,
              box_h);
    return DSG_ERR_CUDA;
  }
  return DSG_OK;
}

inline int make_map_b(CUtensorMap* m, const __half* w, int64_t k_total, int64_t rows, int block_n) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DSG_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)IG_BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(B) failed: %d (K=%lld rows=%lld bn=%d)", (int)r, (long long)k_total,
              (long long)rows, block_n);
    return DSG_ERR_CUDA;
  }
  return DSG_OK;
}

inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

inline IgSrc dense_src(const void* ptr, int C, int H, int W) {
  IgSrc s;
  s.ptr = (const __half*)ptr; s.C = C; s.H = H; s.W = W;
  s.sW = C; s.sH = (int64_t)W * C; s.sN = (int64_t)H * W * C;
  return s;
}


// igemm_halo.cu: returns DSG_OK, an error, or DSG_HALO_SKIP when the shape is outside what the halo kernel covers
constexpr int DSG_HALO_SKIP = 1;
int launch_halo_conv(const dsg_conv_args* a, int block_n, cudaStream_t st);

}  // namespace dsg
