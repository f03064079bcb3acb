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

// ------------------------------------------------------------------ shared epilogue (TMEM -> fp16 NHWC + GN stats)
// GroupNorm statistics ride along for free-ish: every epilogue warp owns 32 accumulator rows; per 32-column slice it
// adds adjacent channel PAIRS in registers (summed over the MT stacked accumulators first), reduces the 16 pair sums
// across its 32 rows with a halving butterfly (16 shuffles per statistic) and parks them in its quarter of
// sstat[4][BLOCK_N / 2][2].  After the tile the four quarters are combined in a fixed order and added to the tensor's
// int64 totals (groupnorm.cu) in the EVEN channel's slot — consumers only ever sum whole groups, and group sizes
// are even whenever this path is used (the engine falls back to dsg_gn_stats otherwise).

// a[0..15], b[0..15]: per-lane pair sums of two statistics.  On return a[0], b[0] hold the warp totals of pair
// (lane >> 1) (identical in lanes 2k and 2k + 1).  The two chains are independent and interleave.
__device__ __forceinline__ void warp_pairsum16x2(float* a, float* b, int lane) {
#pragma unroll
  for (int off = 16; off >= 2; off >>= 1) {
    const int n = off >> 1;
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float ka = hi ? a[i + n] : a[i], sa = hi ? a[i] : a[i + n];
      const float kb = hi ? b[i + n] : b[i], sb = hi ? b[i] : b[i + n];
      a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
      b[i] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
    }
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
  b[0] += __shfl_xor_sync(0xffffffffu, b[0], 1);
}

// One CTA tile of the epilogue for one warp (32 rows): MT stacked accumulators of BLOCK_N fp32 columns each at
// taddr (+ m * BLOCK_N), slice by slice (c outer, m inner).  The tcgen05.ld of the next slice is in flight while
// the current one is converted; the residual (if any) is register-prefetched two slices ahead.
// Stores: with one pixel per lane a direct STG.128 touches 32 different lines with 16 bytes each — every 32-byte sector
// is written in two halves by two instructions and the LSU processes 32 wavefronts per instruction.  With a staging
// buffer (stage_warp: 2 KB per epilogue warp, nullptr = direct stores) the warp parks its 32 pixels x 64 bytes in shared
// memory (16-byte chunks XOR-swizzled: conflict-free both ways) and reads them back TRANSPOSED: lane L then stores chunk
// (L & 3) of pixel (8 i + L / 4), i = 0..3, so one instruction writes eight whole 64-byte segments (full sectors), for
// any pixel stride / sub-pixel phase.  The pixels' offsets come from their owner lanes by shuffle, once per tile.
constexpr int EPI_STAGE_BYTES_PER_WARP = 2048;
// DSG_EPI_XPOSE=0 keeps the direct stores (A/B switch)
inline bool epi_xpose_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DSG_EPI_XPOSE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int BLOCK_N, int MT>
__device__ __forceinline__ void epi_tile(uint32_t taddr, const float* sb, const bool (&valid)[MT],
                                         const int64_t (&off)[MT], __half* out, const __half* res,
                                         float* sstat_warp /* [BLOCK_N / 2][2] or nullptr */, int lane,
                                         uint8_t* stage_warp = nullptr) {
  constexpr int NCH = BLOCK_N / 32, NIT = NCH * MT;
  int64_t offT[MT][4];
  uint32_t vmask[MT];
  if (stage_warp) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      vmask[m] = __ballot_sync(0xffffffffu, valid[m]);
#pragma unroll
      for (int i = 0; i < 4; ++i) offT[m][i] = __shfl_sync(0xffffffffu, off[m], i * 8 + (lane >> 2));
    }
  }
  uint32_t v[2][32];
  // residual: register-prefetched TWO slices ahead (one ahead left a full L2 / DRAM round trip on every slice)
  uint4 rn[2][4];
#pragma unroll
  for (int pre = 0; pre < 2; ++pre) {
    if (pre < NIT) {
      const int pc = pre / MT, pm = pre % MT;
      if (res && valid[pm]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rn[pre][j] = ldg_nc_v4(res + off[pm] + pc * 32 + j * 8);
      }
    }
  }
  tmem_ld_32x32(taddr, v[0]);
  float a1[16], a2[16];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int c = it / MT, m = it % MT;
    tmem_ld_wait();
    if (it + 1 < NIT) {
      const int nc = (it + 1) / MT, nm = (it + 1) % MT;
      tmem_ld_32x32(taddr + (uint32_t)(nm * BLOCK_N + nc * 32), v[(it + 1) & 1]);
    }
    uint4 rc[4];
    if (res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rc[j] = rn[it & 1][j];
      if (it + 2 < NIT) {
        const int nc = (it + 2) / MT, nm = (it + 2) % MT;
        if (valid[nm]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rn[it & 1][j] = ldg_nc_v4(res + off[nm] + nc * 32 + j * 8);
        }
      }
    }
    float f[32];
    if (valid[m]) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sb + c * 32 + j);
        f[j] = __uint_as_float(v[it & 1][j]) + b4.x;
        f[j + 1] = __uint_as_float(v[it & 1][j + 1]) + b4.y;
        f[j + 2] = __uint_as_float(v[it & 1][j + 2]) + b4.z;
        f[j + 3] = __uint_as_float(v[it & 1][j + 3]) + b4.w;
      }
      if (res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float r[8];
          unpack8(rc[j], r);
#pragma unroll
          for (int u = 0; u < 8; ++u) f[j * 8 + u] += r[u];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = 0.f;  // rows outside the image count as nothing
    }
    // the statistics arithmetic comes BEFORE the stores: placed after them, its first instructions overwrote registers the
    // STG.128s had not read yet and sat on that hazard (every lane stores to its own 128-byte line, the store path is slow)
    if (sstat_warp) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x0 = f[2 * j], x1 = f[2 * j + 1];
        const float s = x0 + x1, qq = fmaf(x0, x0, x1 * x1);
        if (m == 0) { a1[j] = s; a2[j] = qq; } else { a1[j] += s; a2[j] += qq; }
      }
      if (m == MT - 1) {
        warp_pairsum16x2(a1, a2, lane);
        if (!(lane & 1)) reinterpret_cast<float2*>(sstat_warp)[c * 16 + (lane >> 1)] = make_float2(a1[0], a2[0]);
      }
    }
    if (stage_warp) {
      uint8_t* mine = stage_warp + lane * 64;
      const int sw = (lane >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(mine + ((j ^ sw) << 4)) = pack8(f + j * 8);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int px = i * 8 + (lane >> 2);
        const uint4 pk = *reinterpret_cast<const uint4*>(stage_warp + px * 64 + (((lane & 3) ^ ((px >> 1) & 3)) << 4));
        if ((vmask[m] >> px) & 1u) stg_v4(out + offT[m][i] + c * 32 + (lane & 3) * 8, pk);
      }
      __syncwarp();   // the next slice overwrites the staging buffer
    } else if (valid[m]) {
      __half* op = out + off[m] + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) stg_v4(op + j, pack8(f + j));
    }
  }
}

// The same tile through shared memory + TMA stores (BLOCK_N = 64): with one pixel per lane a direct STG.128 touches 32
// different 128-byte lines with 16 bytes each — every 32-byte sector is written twice and the LSU store path, not the
// tensor pipe, paces the cout = 64 layers.  Here each warp parks its 32 pixels x 64 channels of every accumulator (4 KB
// each, 16-byte chunks XOR-swizzled with the pixel index so the STS.128s are conflict-free = the TMA unit's SWIZZLE_128B
// pattern) in its own staging buffers and one elected lane issues ONE tensor store per accumulator (box = 64 channels x
// TW pixels x 32 / TW rows) once its second 32-channel slice is staged; rows / columns outside the image are clipped by
// the TMA unit.  `pending`: stores of this warp's previous tile may still be reading the staging buffers.
template <int BLOCK_N, int MT>
__device__ __forceinline__ void epi_tile_tma(uint32_t taddr, const float* sb, const bool (&valid)[MT],
                                             const int64_t (&off)[MT], const __half* res, float* sstat_warp, int lane,
                                             uint8_t* stage_warp /* [MT][4 KB] */, const CUtensorMap* omap, int c0,
                                             int w0, const int (&hrow)[MT], int n, bool& pending) {
  static_assert(BLOCK_N == 64, "one 128-byte swizzle span per pixel row");
  constexpr int NCH = BLOCK_N / 32, NIT = NCH * MT;
  uint32_t v[2][32];
  uint4 rn[2][4];
#pragma unroll
  for (int pre = 0; pre < 2; ++pre) {
    if (pre < NIT) {
      const int pc = pre / MT, pm = pre % MT;
      if (res && valid[pm]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rn[pre][j] = ldg_nc_v4(res + off[pm] + pc * 32 + j * 8);
      }
    }
  }
  tmem_ld_32x32(taddr, v[0]);
  if (pending) {  // the previous tile's stores have had a whole main loop to drain the staging buffers
    if (lane == 0) bulk_wait_group_read0();
    __syncwarp();
    pending = false;
  }
  float a1[16], a2[16];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int c = it / MT, m = it % MT;
    tmem_ld_wait();
    if (it + 1 < NIT) {
      const int nc = (it + 1) / MT, nm = (it + 1) % MT;
      tmem_ld_32x32(taddr + (uint32_t)(nm * BLOCK_N + nc * 32), v[(it + 1) & 1]);
    }
    uint4 rc[4];
    if (res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rc[j] = rn[it & 1][j];
      if (it + 2 < NIT) {
        const int nc = (it + 2) / MT, nm = (it + 2) % MT;
        if (valid[nm]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rn[it & 1][j] = ldg_nc_v4(res + off[nm] + nc * 32 + j * 8);
        }
      }
    }
    float f[32];
    if (valid[m]) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sb + c * 32 + j);
        f[j] = __uint_as_float(v[it & 1][j]) + b4.x;
        f[j + 1] = __uint_as_float(v[it & 1][j + 1]) + b4.y;
        f[j + 2] = __uint_as_float(v[it & 1][j + 2]) + b4.z;
        f[j + 3] = __uint_as_float(v[it & 1][j + 3]) + b4.w;
      }
      if (res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float r[8];
          unpack8(rc[j], r);
#pragma unroll
          for (int u = 0; u < 8; ++u) f[j * 8 + u] += r[u];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = 0.f;  // rows outside the image count as nothing (and are clipped by the store)
    }
    if (sstat_warp) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x0 = f[2 * j], x1 = f[2 * j + 1];
        const float s = x0 + x1, qq = fmaf(x0, x0, x1 * x1);
        if (m == 0) { a1[j] = s; a2[j] = qq; } else { a1[j] += s; a2[j] += qq; }
      }
      if (m == MT - 1) {
        warp_pairsum16x2(a1, a2, lane);
        if (!(lane & 1)) reinterpret_cast<float2*>(sstat_warp)[c * 16 + (lane >> 1)] = make_float2(a1[0], a2[0]);
      }
    }
    uint8_t* buf = stage_warp + m * 4096 + lane * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(buf + (((c * 4 + j) ^ (lane & 7)) << 4)) = pack8(f + j * 8);
    if (c == NCH - 1) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(omap, stage_warp + m * 4096, c0, w0, hrow[m], n);
        bulk_commit_group();
      }
      pending = true;
    }
  }
}

// combine the four warps' pair totals of one tile (fixed order), convert to the exact fixed-point form and keep them in
// registers (entry idx is always handled by the same thread); the int64 atomics on the tensor's totals are issued only
// when the CTA moves on to another (sample, channel block) or runs out of tiles.  Integer adds: the totals are bit-identical
// to flushing every tile, whatever tiles a CTA happens to process.
template <int BLOCK_N>
struct EpiStatsAcc {
  static constexpr int NV = (BLOCK_N + 127) / 128;
  long long v[NV];
  long long* dst;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = 0;
    dst = nullptr;
  }
  __device__ __forceinline__ void emit(int te) {
    if (dst != nullptr) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int idx = te + k * 128;
        if (idx < BLOCK_N && te < 128 && v[k] != 0)
          atomicAdd(reinterpret_cast<unsigned long long*>(dst + (idx >> 1) * 4 + (idx & 1)), (unsigned long long)v[k]);
        v[k] = 0;
      }
    }
  }
  // NW = epilogue warps that parked pair totals for this tile (4, or 8 when two warps share a TMEM lane quadrant)
  template <int NW = 4>
  __device__ __forceinline__ void add_tile(const float* sstat /* [NW][BLOCK_N / 2][2] */, int te, long long* stats_nc) {
    if (stats_nc != dst) {
      emit(te);
      dst = stats_nc;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int idx = te + k * 128;
      if (idx < BLOCK_N && te < 128) {
        float t = ((sstat[idx] + sstat[BLOCK_N + idx]) + sstat[2 * BLOCK_N + idx]) + sstat[3 * BLOCK_N + idx];
        if constexpr (NW == 8)
          t += ((sstat[4 * BLOCK_N + idx] + sstat[5 * BLOCK_N + idx]) + sstat[6 * BLOCK_N + idx]) + sstat[7 * BLOCK_N + idx];
        v[k] += (idx & 1) ? gn_fix_sq(t) : gn_fix_sum(t);
      }
    }
  }
};

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


inline IgSrc dense_src(const void* ptr, int C, int H, int W) {
  IgSrc s;
  s.ptr = (const __half*)ptr; s.C = C; s.H = H; s.W = W;
  s.sW = C; s.sH = (int64_t)W * C; s.sN = (int64_t)H * W * C;
  return s;
}


// igemm_halo.cu: returns DSG_OK, an error, or DSG_HALO_SKIP when the shape is outside what the halo kernel covers
constexpr int DSG_HALO_SKIP = 1;
int launch_halo_conv(const dsg_conv_args* a, int block_n, int cta_pair, cudaStream_t st);

}  // namespace dsg
