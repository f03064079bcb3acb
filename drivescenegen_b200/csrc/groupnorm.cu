// groupnorm.cu — GroupNorm(+SiLU) over NHWC fp16 activations, concat-free (two channel-concatenated sources).
// Replaces torch.nn.GroupNorm(32, C, eps=1e-5) + SiLU inside upstream ResnetBlock2D / Attention / conv_norm_out
// (diffusers 0.20.0 models/resnet.py, models/attention_processor.py) and the
// torch.cat([h, skip], dim=1) of UpBlock2D.forward.
//
// Statistics are kept PER CHANNEL as exact fixed-point integers: stats[n][c] = { sum(x) * 2^24, sum(x^2) * 2^20 }
// (int64).  Producers add tile partials with 64-bit integer atomics — associative, so the totals are bit-reproducible
// whatever the order — either from the conv epilogues (igemm*.cu: the statistics pass then costs no HBM traffic at
// all) or from gn_stats_kernel below (one read of the tensor; used for conv_in's output and stand-alone calls).
// Per-channel totals let ONE set of statistics serve every consumer: a skip tensor is normalised once on the way down
// and once, concatenated with another tensor and therefore with different group boundaries, on the way up.
// (The tcgen05 epilogues store the totals of each channel PAIR in the even channel's slot — half the shuffles — so
// consumers may only sum whole groups; group sizes and source widths are even whenever that path is used.)
// gn_apply_kernel sums the integer totals of each group exactly, forms mean / variance in double precision and
// streams y = act(x * a_c + b_c): one read + one write of the tensor.
#include "common.cuh"

namespace dsg {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_MAX_C = 2048;  // 256 threads x 8 channels
#ifndef GN_ILP_DEF
#define GN_ILP_DEF 4
#endif
constexpr int GN_ILP = GN_ILP_DEF;
#ifndef GN_CTAS_DEF
#define GN_CTAS_DEF 3
#endif
constexpr int GN_CTAS = GN_CTAS_DEF;   // resident gn_apply CTAs per SM (register-limited)

// ------------------------------------------------------------------ per-channel statistics of one tensor
// grid (chunks, n); thread = (pixel row, 8-channel vector).  Block partials in fp32 (<= ~1024 pixels), totals in int64.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const __half* __restrict__ x, int C,
                                                              long long* __restrict__ stats, int64_t hw,
                                                              int64_t px_per_chunk) {
  const int V = C >> 3;
  const int ppi = GN_THREADS / V;  // pixel rows per iteration (V <= 256)
  const int n = blockIdx.y, chunk = blockIdx.x;
  pdl_sync();
  __shared__ float s_part[2][GN_THREADS * 8];
  const int64_t base_px = (int64_t)n * hw;
  if ((int)threadIdx.x < ppi * V) {
    const int v = threadIdx.x % V, prow = threadIdx.x / V;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const int64_t p_begin = (int64_t)chunk * px_per_chunk;
    int64_t p_end = p_begin + px_per_chunk;
    if (p_end > hw) p_end = hw;
    // GN_ILP independent 16-byte loads in flight per thread: the loop is latency-bound otherwise
    for (int64_t p = p_begin + prow; p < p_end; p += (int64_t)ppi * GN_ILP) {
      uint4 raw[GN_ILP];
#pragma unroll
      for (int u = 0; u < GN_ILP; ++u) {
        const int64_t pp = p + (int64_t)u * ppi;
        if (pp < p_end) raw[u] = ldg_nc_v4(x + (base_px + pp) * C + (v << 3));
      }
#pragma unroll
      for (int u = 0; u < GN_ILP; ++u) {
        if (p + (int64_t)u * ppi < p_end) {
          float f[8];
          unpack8(raw[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += f[j];
            q[j] = fmaf(f[j], f[j], q[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_part[0][prow * C + (v << 3) + j] = s[j];
      s_part[1][prow * C + (v << 3) + j] = q[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    float ts = 0.f, tq = 0.f;
    for (int r = 0; r < ppi; ++r) {  // fixed order: the block partial is deterministic
      ts += s_part[0][r * C + c];
      tq += s_part[1][r * C + c];
    }
    long long* o = stats + ((int64_t)n * C + c) * 2;
    atomicAdd(reinterpret_cast<unsigned long long*>(o), (unsigned long long)gn_fix_sum(ts));
    atomicAdd(reinterpret_cast<unsigned long long*>(o + 1), (unsigned long long)gn_fix_sq(tq));
  }
}

// ------------------------------------------------------------------ normalise + affine + activation
__global__ void __launch_bounds__(GN_THREADS, GN_CTAS) gn_apply_kernel(const __half* __restrict__ x1, int c1,
                                                              const long long* __restrict__ st1,
                                                              const __half* __restrict__ x2, int c2,
                                                              const long long* __restrict__ st2,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int act,
                                                              __half* __restrict__ y, int64_t hw, int groups,
                                                              int64_t px_per_cta, double inv_cnt_s, double inv_cnt_q) {
  const int C = c1 + c2, V = C >> 3, cpg = C / groups;
  const int ppi = GN_THREADS / V;
  const int n = blockIdx.y;
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  const bool active = (int)threadIdx.x < ppi * V;
  const int v = active ? threadIdx.x % V : 0, prow = active ? threadIdx.x / V : 0;
  const int ch0 = v << 3;
  const bool from1 = ch0 < c1;  // c1 is a multiple of 8, so a vector never straddles the two sources
  const __half* src = from1 ? x1 : x2;
  const int cs = from1 ? c1 : c2, co = from1 ? ch0 : ch0 - c1;
  const int64_t base_px = (int64_t)n * hw;
  const int64_t p_begin = (int64_t)blockIdx.x * px_per_cta;
  int64_t p_end = p_begin + px_per_cta;
  if (p_end > hw) p_end = hw;
  pdl_sync();
  // the first batch of loads does not depend on the statistics: put it in flight before the prologue
  uint4 raw[GN_ILP];
  int64_t p = p_begin + prow;
  if (active) {
#pragma unroll
    for (int u = 0; u < GN_ILP; ++u) {
      const int64_t pp = p + (int64_t)u * ppi;
      if (pp < p_end) raw[u] = ldg_nc_v4(src + (base_px + pp) * cs + co);
    }
  }
  // the affine parameters do not depend on the statistics either
  float gm8[8], bt8[8];
  if (active) {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + ch0), g1 = *reinterpret_cast<const float4*>(gamma + ch0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + ch0), b1 = *reinterpret_cast<const float4*>(beta + ch0 + 4);
    gm8[0] = g0.x; gm8[1] = g0.y; gm8[2] = g0.z; gm8[3] = g0.w; gm8[4] = g1.x; gm8[5] = g1.y; gm8[6] = g1.z; gm8[7] = g1.w;
    bt8[0] = b0.x; bt8[1] = b0.y; bt8[2] = b0.z; bt8[3] = b0.w; bt8[4] = b1.x; bt8[5] = b1.y; bt8[6] = b1.z; bt8[7] = b1.w;
  }
  // exact integer totals per group (only whole groups are ever summed: conv epilogues store channel PAIRS in the even
  // slot).  Every thread fetches whole channels in ONE round of loads and adds them with shared-memory integer
  // atomics — order-independent — instead of 32 threads walking their group's channels one dependent load at a time.
  __shared__ unsigned long long s_t[GN_MAX_GROUPS][2];
  if ((int)threadIdx.x < groups) { s_t[threadIdx.x][0] = 0ull; s_t[threadIdx.x][1] = 0ull; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const longlong2 tv = *reinterpret_cast<const longlong2*>(
        c < c1 ? st1 + ((int64_t)n * c1 + c) * 2 : st2 + ((int64_t)n * c2 + (c - c1)) * 2);
    if (tv.x != 0 || tv.y != 0) {
      atomicAdd(&s_t[c / cpg][0], (unsigned long long)tv.x);
      atomicAdd(&s_t[c / cpg][1], (unsigned long long)tv.y);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < groups) {
    // mean and variance from the exact integer totals in double (the subtraction cancels); the reciprocal square root
    // itself in fp32 — FP64 sqrt / divide are long sequences on this part and sit on every CTA's critical path
    const int g = threadIdx.x;
    const double mg = (double)(long long)s_t[g][0] * inv_cnt_s;
    double vg = (double)(long long)s_t[g][1] * inv_cnt_q - mg * mg;
    if (vg < 0.0) vg = 0.0;
    s_mean[g] = (float)mg;
    s_rstd[g] = rsqrtf((float)vg + eps);
  }
  __syncthreads();
  if (!active) return;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (ch0 + j) / cpg;
    a[j] = gm8[j] * s_rstd[g];
    b[j] = bt8[j] - s_mean[g] * a[j];
  }
  while (p < p_end) {
    const int64_t pn = p + (int64_t)ppi * GN_ILP;
    uint4 nxt[GN_ILP];
#pragma unroll
    for (int u = 0; u < GN_ILP; ++u) {  // next batch in flight while this one is converted
      const int64_t pp = pn + (int64_t)u * ppi;
      if (pp < p_end) nxt[u] = ldg_nc_v4(src + (base_px + pp) * cs + co);
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
#pragma unroll
    for (int u = 0; u < GN_ILP; ++u) raw[u] = nxt[u];
    p = pn;
  }
}

// ------------------------------------------------------------------ coefficients for the fused (in-conv) form
// grid = n; coef[n][c] = (a / 2, b / 2), GroupNorm(x) = a x + b.  Same statistics arithmetic as gn_apply_kernel.
__global__ void __launch_bounds__(GN_THREADS) gn_coef_kernel(int c1, const long long* __restrict__ st1, int c2,
                                                             const long long* __restrict__ st2,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             float2* __restrict__ coef, int groups, double inv_cnt_s,
                                                             double inv_cnt_q) {
  const int C = c1 + c2, cpg = C / groups;
  const int n = blockIdx.x;
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  __shared__ unsigned long long s_t[GN_MAX_GROUPS][2];
  pdl_sync();
  if ((int)threadIdx.x < groups) { s_t[threadIdx.x][0] = 0ull; s_t[threadIdx.x][1] = 0ull; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const longlong2 tv = *reinterpret_cast<const longlong2*>(
        c < c1 ? st1 + ((int64_t)n * c1 + c) * 2 : st2 + ((int64_t)n * c2 + (c - c1)) * 2);
    if (tv.x != 0 || tv.y != 0) {
      atomicAdd(&s_t[c / cpg][0], (unsigned long long)tv.x);
      atomicAdd(&s_t[c / cpg][1], (unsigned long long)tv.y);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    const double mg = (double)(long long)s_t[g][0] * inv_cnt_s;
    double vg = (double)(long long)s_t[g][1] * inv_cnt_q - mg * mg;
    if (vg < 0.0) vg = 0.0;
    s_mean[g] = (float)mg;
    s_rstd[g] = rsqrtf((float)vg + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cpg;
    const float a = gamma[c] * s_rstd[g];
    const float b = beta[c] - s_mean[g] * a;
    coef[(int64_t)n * C + c] = make_float2(0.5f * a, 0.5f * b);
  }
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_gn_coef(int32_t c1, const void* stats1, int32_t c2, const void* stats2, const float* gamma, const float* beta,
                float eps, float* coef, int32_t n, int64_t hw, int32_t groups, void* stream) {
  DSG_CHECK_ARG(stats1 && c1 > 0 && (stats2 == nullptr) == (c2 == 0) && c2 >= 0, "dsg_gn_coef: stats / channel mismatch");
  const int C = c1 + c2;
  DSG_CHECK_ARG(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "dsg_gn_coef: bad groups %d for C=%d", groups, C);
  DSG_CHECK_ARG(gamma && beta && coef && n >= 0 && hw > 0, "dsg_gn_coef: bad args");
  DSG_CHECK_ARG(((uintptr_t)stats1 | (uintptr_t)stats2 | (uintptr_t)coef) % 16 == 0, "dsg_gn_coef: unaligned pointer");
  if (n == 0) return DSG_OK;
  launch_k(gn_coef_kernel, dim3(n), dim3(GN_THREADS), 0, (cudaStream_t)stream, c1, (const long long*)stats1, c2,
           (const long long*)stats2, gamma, beta, eps, (float2*)coef, groups,
           1.0 / 16777216.0 / ((double)hw * (double)(C / groups)), 1.0 / 1048576.0 / ((double)hw * (double)(C / groups)));
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_coef");
  return DSG_OK;
}

int dsg_gn_stats(const void* x, int32_t c, void* stats, int32_t n, int64_t hw, void* stream) {
  DSG_CHECK_ARG(x && stats && c > 0 && c % 8 == 0 && c <= GN_MAX_C, "dsg_gn_stats: null pointer or bad c=%d", c);
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && hw > 0, "dsg_gn_stats: bad n/hw");
  DSG_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)stats % 8 == 0, "dsg_gn_stats: unaligned pointer");
  if (n == 0) return DSG_OK;
  // chunks of <= ~1024 pixels keep the fp32 block partials accurate; at least ~4 blocks per SM for bandwidth
  int64_t chunks = ceil_div64(hw, 1024);
  const int64_t want = ceil_div64(148 * 4, n);
  if (chunks < want) chunks = want;
  const int64_t max_chunks = ceil_div64(hw, 32);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 65535) chunks = 65535;
  const int64_t ppc = ceil_div64(hw, chunks);
  chunks = ceil_div64(hw, ppc);
  launch_k(gn_stats_kernel, dim3((unsigned)chunks, n), GN_THREADS, 0, (cudaStream_t)stream, (const __half*)x, c,
           (long long*)stats, hw, ppc);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_stats");
  return DSG_OK;
}

int dsg_gn_apply(const void* x1, int32_t c1, const void* stats1, const void* x2, int32_t c2, const void* stats2,
                 const float* gamma, const float* beta, float eps, int32_t act, void* y, int32_t n, int64_t hw,
                 int32_t groups, void* stream) {
  DSG_CHECK_ARG(x1 && stats1 && c1 > 0 && c1 % 8 == 0, "dsg_gn_apply: x1/stats1 null or c1 not a multiple of 8");
  DSG_CHECK_ARG((x2 == nullptr) == (c2 == 0) && (x2 == nullptr) == (stats2 == nullptr) && c2 % 8 == 0,
                "dsg_gn_apply: x2/stats2/c2 mismatch");
  const int C = c1 + c2;
  DSG_CHECK_ARG(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "dsg_gn_apply: bad groups %d for C=%d",
                groups, C);
  DSG_CHECK_ARG(C <= GN_MAX_C, "dsg_gn_apply: C=%d too large (max %d)", C, GN_MAX_C);
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && hw > 0, "dsg_gn_apply: bad n/hw");
  DSG_CHECK_ARG(gamma && beta && y, "dsg_gn_apply: null pointer");
  DSG_CHECK_ARG(((uintptr_t)gamma | (uintptr_t)beta) % 16 == 0, "dsg_gn_apply: gamma / beta must be 16-byte aligned");
  DSG_CHECK_ARG(((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)y) % 16 == 0 &&
                    ((uintptr_t)stats1 | (uintptr_t)stats2) % 16 == 0,
                "dsg_gn_apply: unaligned pointer");
  if (n == 0) return DSG_OK;
  const int V = C / 8, ppi = GN_THREADS / V;
  // about one full wave of CTAs (3 per SM at 74 registers) for big tensors — the per-CTA prologue is amortised and there is no
  // tail wave — but never less than one batch of loads per thread
  const int64_t batch_px = (int64_t)ppi * GN_ILP;
  int64_t per_sample = (num_sms() * GN_CTAS) / n;
  if (per_sample < 1) per_sample = 1;
  int64_t px_per_cta = ceil_div64(ceil_div64(hw, per_sample), batch_px) * batch_px;
  int64_t ctas = ceil_div64(hw, px_per_cta);
  if (ctas > 65535) { px_per_cta = ceil_div64(hw, 65535); ctas = ceil_div64(hw, px_per_cta); }
  launch_k(gn_apply_kernel, dim3((unsigned)ctas, n), GN_THREADS, 0, (cudaStream_t)stream, (const __half*)x1, c1,
           (const long long*)stats1, (const __half*)x2, c2, (const long long*)stats2, gamma, beta, eps, act, (__half*)y,
           hw, groups, px_per_cta, 1.0 / 16777216.0 / ((double)hw * (double)(C / groups)),
           1.0 / 1048576.0 / ((double)hw * (double)(C / groups)));
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_apply");
  return DSG_OK;
}
}
