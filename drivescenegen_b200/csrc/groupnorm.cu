// groupnorm.cu — GroupNorm(+SiLU) over NHWC fp16 activations, concat-free (two channel-concatenated sources).
// Replaces torch.nn.GroupNorm(32, C, eps=1e-5) + SiLU inside upstream ResnetBlock2D / Attention / conv_norm_out
// (diffusers 0.20.0 models/resnet.py, models/attention_processor.py) and the
// torch.cat([h, skip], dim=1) of UpBlock2D.forward.  HBM-bound: stats = 1 read, apply = 1 read + 1 write.
#include "common.cuh"

namespace dsg {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_ILP = 4;

__device__ __forceinline__ const __half* src_of(const __half* x1, int c1, const __half* x2, int c2, int64_t pix,
                                                int ch) {
  return ch < c1 ? x1 + pix * c1 + ch : x2 + pix * c2 + (ch - c1);
}

// partial[n][chunk][g] = { sum(x - K_g), sum((x - K_g)^2) } with K_g = x[n][pixel 0][first channel of g]
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const __half* __restrict__ x1, int c1,
                                                              const __half* __restrict__ x2, int c2,
                                                              float* __restrict__ partial, int64_t hw, int groups,
                                                              int64_t px_per_chunk) {
  const int C = c1 + c2, V = C >> 3, cpg = C / groups;
  const int ppi = GN_THREADS / V;  // pixels per iteration
  const int n = blockIdx.y, chunk = blockIdx.x;
  // per-thread per-channel partials, reduced in a fixed order (deterministic, no atomics)
  __shared__ float s_part[2][GN_THREADS * 8];
  const int64_t base_px = (int64_t)n * hw;
  if ((int)threadIdx.x < ppi * V) {
    const int v = threadIdx.x % V, prow = threadIdx.x / V;
    const int ch0 = v << 3;
    float K[8], s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (ch0 + j) / cpg;
      K[j] = __half2float(*src_of(x1, c1, x2, c2, base_px, g * cpg));
      s[j] = 0.f; q[j] = 0.f;
    }
    const bool from1 = ch0 < c1;  // c1 is a multiple of 8, so a vector never straddles the two sources
    const __half* src = from1 ? x1 : x2;
    const int cs = from1 ? c1 : c2, co = from1 ? ch0 : ch0 - c1;
    const int64_t p_begin = (int64_t)chunk * px_per_chunk;
    int64_t p_end = p_begin + px_per_chunk;
    if (p_end > hw) p_end = hw;
    // GN_ILP independent 16-byte loads in flight per thread: the loop is latency-bound otherwise
    for (int64_t p = p_begin + prow; p < p_end; p += (int64_t)ppi * GN_ILP) {
      uint4 raw[GN_ILP];
#pragma unroll
      for (int u = 0; u < GN_ILP; ++u) {
        const int64_t pp = p + (int64_t)u * ppi;
        if (pp < p_end) raw[u] = ldg_nc_v4(src + (base_px + pp) * cs + co);
      }
#pragma unroll
      for (int u = 0; u < GN_ILP; ++u) {
        if (p + (int64_t)u * ppi < p_end) {
          float f[8];
          unpack8(raw[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - K[j];
            s[j] += d;
            q[j] = fmaf(d, d, q[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_part[0][prow * C + ch0 + j] = s[j];
      s_part[1][prow * C + ch0 + j] = q[j];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    float ts = 0.f, tq = 0.f;
    for (int r = 0; r < ppi; ++r)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        ts += s_part[0][r * C + c];
        tq += s_part[1][r * C + c];
      }
    float* o = partial + (((int64_t)n * gridDim.x + chunk) * groups + g) * 2;
    o[0] = ts;
    o[1] = tq;
  }
}

__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const __half* __restrict__ x1, int c1,
                                                              const __half* __restrict__ x2, int c2,
                                                              const float* __restrict__ partial, int chunks,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int act,
                                                              __half* __restrict__ y, int64_t hw, int groups,
                                                              int64_t px_per_cta) {
  const int C = c1 + c2, V = C >> 3, cpg = C / groups;
  const int ppi = GN_THREADS / V;
  const int n = blockIdx.y;
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  const int64_t base_px = (int64_t)n * hw;
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    double S1 = 0.0, S2 = 0.0;
    const float* p = partial + ((int64_t)n * chunks * groups + g) * 2;
    for (int c = 0; c < chunks; ++c) {
      S1 += (double)p[(int64_t)c * groups * 2];
      S2 += (double)p[(int64_t)c * groups * 2 + 1];
    }
    const double cnt = (double)hw * cpg;
    const double K = (double)__half2float(*src_of(x1, c1, x2, c2, base_px, g * cpg));
    const double m1 = S1 / cnt;
    double var = S2 / cnt - m1 * m1;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)(K + m1);
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if ((int)threadIdx.x >= ppi * V) return;
  const int v = threadIdx.x % V, prow = threadIdx.x / V;
  const int ch0 = v << 3;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j, g = ch / cpg;
    a[j] = gamma[ch] * s_rstd[g];
    b[j] = beta[ch] - s_mean[g] * a[j];
  }
  const bool from1 = ch0 < c1;
  const __half* src = from1 ? x1 : x2;
  const int cs = from1 ? c1 : c2, co = from1 ? ch0 : ch0 - c1;
  const int64_t p_begin = (int64_t)blockIdx.x * px_per_cta;
  int64_t p_end = p_begin + px_per_cta;
  if (p_end > hw) p_end = hw;
  for (int64_t p = p_begin + prow; p < p_end; p += (int64_t)ppi * GN_ILP) {
    uint4 raw[GN_ILP];
#pragma unroll
    for (int u = 0; u < GN_ILP; ++u) {
      const int64_t pp = p + (int64_t)u * ppi;
      if (pp < p_end) raw[u] = ldg_nc_v4(src + (base_px + pp) * cs + co);
    }
#pragma unroll
    for (int u = 0; u < GN_ILP; ++u) {
      const int64_t pp = p + (int64_t)u * ppi;
      if (pp < p_end) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(f[j], a[j], b[j]);
          f[j] = act ? silu_f(t) : t;
        }
        stg_v4(y + (base_px + pp) * C + ch0, pack8(f));
      }
    }
  }
}

static int gn_check(const void* x1, int c1, const void* x2, int c2, int n, int64_t hw, int groups, const char* who) {
  DSG_CHECK_ARG(x1 && c1 > 0 && c1 % 8 == 0, "%s: x1 null or c1 not a multiple of 8", who);
  DSG_CHECK_ARG((x2 == nullptr) == (c2 == 0) && c2 % 8 == 0, "%s: x2/c2 mismatch", who);
  const int C = c1 + c2;
  DSG_CHECK_ARG(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "%s: bad groups %d for C=%d", who, groups, C);
  DSG_CHECK_ARG(C / 8 <= GN_THREADS, "%s: C=%d too large (max %d)", who, C, GN_THREADS * 8);
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && hw > 0, "%s: bad n/hw", who);
  DSG_CHECK_ARG(((uintptr_t)x1 | (uintptr_t)x2) % 16 == 0, "%s: pointers must be 16-byte aligned", who);
  return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int32_t dsg_gn_chunks(int64_t hw) {
  int64_t c = ceil_div64(hw, 64);
  if (c > 64) c = 64;
  if (c < 1) c = 1;
  return (int32_t)c;
}

int dsg_gn_stats(const void* x1, int32_t c1, const void* x2, int32_t c2, float* partial, int32_t n, int64_t hw,
                 int32_t groups, void* stream) {
  int rc = gn_check(x1, c1, x2, c2, n, hw, groups, "dsg_gn_stats");
  if (rc) return rc;
  DSG_CHECK_ARG(partial, "dsg_gn_stats: partial is null");
  if (n == 0) return DSG_OK;
  const int chunks = dsg_gn_chunks(hw);
  const int64_t ppc = ceil_div64(hw, chunks);
  gn_stats_kernel<<<dim3(chunks, n), GN_THREADS, 0, (cudaStream_t)stream>>>((const __half*)x1, c1, (const __half*)x2,
                                                                            c2, partial, hw, groups, ppc);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_stats");
  return DSG_OK;
}

int dsg_gn_apply(const void* x1, int32_t c1, const void* x2, int32_t c2, const float* partial, const float* gamma,
                 const float* beta, float eps, int32_t act, void* y, int32_t n, int64_t hw, int32_t groups,
                 void* stream) {
  int rc = gn_check(x1, c1, x2, c2, n, hw, groups, "dsg_gn_apply");
  if (rc) return rc;
  DSG_CHECK_ARG(partial && gamma && beta && y && (uintptr_t)y % 16 == 0, "dsg_gn_apply: null/unaligned pointer");
  if (n == 0) return DSG_OK;
  const int V = (c1 + c2) / 8, ppi = GN_THREADS / V;
  int64_t px_per_cta = (int64_t)ppi * 16;
  int64_t ctas = ceil_div64(hw, px_per_cta);
  if (ctas > 65535) { px_per_cta = ceil_div64(hw, 65535); ctas = ceil_div64(hw, px_per_cta); }
  gn_apply_kernel<<<dim3((unsigned)ctas, n), GN_THREADS, 0, (cudaStream_t)stream>>>(
      (const __half*)x1, c1, (const __half*)x2, c2, partial, dsg_gn_chunks(hw), gamma, beta, eps, act, (__half*)y, hw,
      groups, px_per_cta);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_apply");
  return DSG_OK;
}
}
