// groupnorm_bwd.cu — backward of GroupNorm(+SiLU) over NHWC fp16 activations (training path).
// Replaces the autograd backward of torch.nn.GroupNorm(32, C, eps) + SiLU inside upstream ResnetBlock2D / Attention /
// conv_norm_out (diffusers 0.20.0 models/resnet.py; reached from DriveSceneGen/pipeline/training_pipeline.py:86
// `accelerator.backward(loss)`), including the split of the gradient over the two tensors that torch.cat joined.
//
// With xh = (x - mean_g) * rstd_g, y = gamma * xh + beta, a = act(y) and the incoming gradient da:
//   g        = da * act'(y)
//   A_c      = sum_px g            (= d beta_c per sample)       B_c = sum_px g * xh   (= d gamma_c per sample)
//   m1_g     = sum_{c in g} gamma_c A_c / count                  m2_g = sum_{c in g} gamma_c B_c / count
//   dx       = rstd_g * (gamma_c * g - m1_g - xh * m2_g)
// Two streaming passes: gn_bwd_stats_kernel (reads da, x) writes per-(sample, pixel-chunk, channel) partials of A and B
// in a fixed layout; gn_bwd_apply_kernel (reads da, x again) sums the chunk partials in a fixed order — deterministic,
// no atomics — and writes dx.  Both evaluate g with the SAME fp16x2 instruction sequence (one MUFU op per PAIR of
// elements, silu_grad_h2): the fp32 form made both passes issue-bound at ~0.55 of the HBM rate, and storing g after the
// first pass instead (measured) costs a sixth pass over the tensor for nothing — the kernels are memory-bound now.
// Also measured and dropped (profiles/README.md, r2m): both passes as ONE persistent launch interleaved by sample group so
// that pass 2 would find x / dy in the L2 — with the 18 work items per sample and pass that the partial-sum layout allows,
// half of the resident CTAs sit waiting for their group's pass 1 in every wave (2.3x slower); enough items per sample to
// fill a wave with ONE sample's pass would make the per-item prologue (statistics -> mean / rstd) dominate, optionally (+ an addend tensor: the ResnetBlock shortcut's gradient) (+ the previous
// content of the destination: a tensor with two consumers), split over the two concatenated sources.  It can also emit
// per-CTA column sums of what it wrote (the time-embedding / conv1-bias gradient of a ResnetBlock).
// Forward statistics come from the same int64 per-channel totals the forward used (groupnorm.cu).
#include "common.cuh"
#include "reduce.cuh"

namespace dsg {

constexpr int GB_THREADS = 256;
constexpr int GB_MAX_GROUPS = 64;
constexpr int GB_MAX_C = 2048;
constexpr int GB_MAX_CHUNKS = 64;

struct GnBwdArgs {
  const __half* dy;   // [n][hw][C] incoming gradient (w.r.t. the activated, normalised tensor)
  const __half* x1; int c1; const long long* st1;
  const __half* x2; int c2; const long long* st2;
  const float* gamma; const float* beta;
  float eps; int act;
  float* partial;     // [n][chunks + 1][C][2]: chunk partials (sum g, sum g x), then the per-sample (A_c, B_c)
  int chunks;
  float* coef;        // [n][C][4] (behind the partials): ga, ybh, pc, qc of dx = ga g + pc x + qc, y / 2 = x ga / 2 + ybh
  const __half* addend;  // optional [n][hw][C]
  __half* dx1; int acc1;
  __half* dx2; int acc2;
  float* colsum;      // optional [n][ctas][C]: column sums of the GroupNorm term
  float* osum1;       // optional [n][ctas][c1]: column sums of the FINAL value written to dx1
  float* osum2;       // optional [n][ctas][c2]: same for dx2
  int64_t hw; int groups;
  int64_t px_per_block;
  double inv_cnt_s, inv_cnt_q;   // 2^-24 / count and 2^-20 / count (count = hw * channels per group)
};

// mean / rstd per group from the exact integer totals (same arithmetic as gn_apply_kernel)
__device__ __forceinline__ void gn_moments(const GnBwdArgs& a, int n, float* s_mean, float* s_rstd,
                                           unsigned long long (*s_t)[2]) {
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  if ((int)threadIdx.x < a.groups) { s_t[threadIdx.x][0] = 0ull; s_t[threadIdx.x][1] = 0ull; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GB_THREADS) {
    const longlong2 tv = *reinterpret_cast<const longlong2*>(
        c < a.c1 ? a.st1 + ((int64_t)n * a.c1 + c) * 2 : a.st2 + ((int64_t)n * a.c2 + (c - a.c1)) * 2);
    if (tv.x != 0 || tv.y != 0) {
      atomicAdd(&s_t[c / cpg][0], (unsigned long long)tv.x);
      atomicAdd(&s_t[c / cpg][1], (unsigned long long)tv.y);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < a.groups) {
    const int g = threadIdx.x;
    // identical arithmetic to gn_apply_kernel (the backward must see the forward's mean / rstd bit for bit)
    const double mg = (double)(long long)s_t[g][0] * a.inv_cnt_s;
    double vg = (double)(long long)s_t[g][1] * a.inv_cnt_q - mg * mg;
    if (vg < 0.0) vg = 0.0;
    s_mean[g] = (float)mg;
    s_rstd[g] = rsqrtf((float)vg + a.eps);
  }
  __syncthreads();
}

// d/dy [y * sigmoid(y)]
__device__ __forceinline__ float silu_grad_f(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  const float sg = fmaf(0.5f, t, 0.5f);
  return sg * fmaf(y, 1.0f - sg, 1.0f);
}

// The same for two elements in fp16x2, given h = y / 2 (already rounded to fp16): ONE MUFU op per pair and four fp16x2
// instructions.  t = tanh(h); sigmoid = (1 + t) / 2; d/dy silu = sigmoid * (1 + y (1 - sigmoid)) = sigmoid * (1 + h (1 - t)).
// tanh.approx.f16x2 has the same ~2^-11 absolute error near |t| = 1 as tanh.approx.f32, which is what bounds the
// accuracy of either form.
__device__ __forceinline__ __half2 silu_grad_h2(__half2 h) {
  uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(hi));
  const __half2 t = *reinterpret_cast<__half2*>(&ti);
  const __half2 one = __float2half2_rn(1.0f), half_ = __float2half2_rn(0.5f);
  const __half2 sg = __hfma2(half_, t, half_);
  const __half2 w = __hfma2(h, __hsub2(one, t), one);
  return __hmul2(sg, w);
}

// Pass 1.  Per (sample, pixel chunk, channel): sum g and sum g * x (raw x: the centring is applied when the chunks are
// combined, sum g * xh = rstd * (sum g x - mean * sum g), which keeps the per-thread state small).
#ifndef GB_ILP_DEF
#define GB_ILP_DEF 6   /* 4 -> 6: 1-2.5 % (same-box sweep, one-wave grids) */
#endif
constexpr int GB_ILP = GB_ILP_DEF;
template <bool ACT>
__device__ __forceinline__ void gn_bwd_stats_body(const GnBwdArgs& a, const int n, const int chunk,
                                                  const int64_t px_per_block) {
  const int C = a.c1 + a.c2, V = C >> 3, cpg = C / a.groups;
  const int ppi = GB_THREADS / V;
  __shared__ float s_mean[GB_MAX_GROUPS], s_rstd[GB_MAX_GROUPS];
  __shared__ unsigned long long s_t[GB_MAX_GROUPS][2];
  __shared__ float s_part[2][GB_THREADS * 8];
  const bool active = (int)threadIdx.x < ppi * V;
  const int v = active ? threadIdx.x % V : 0, prow = active ? threadIdx.x / V : 0;
  const int ch0 = v << 3;
  const bool from1 = ch0 < a.c1;
  const __half* src = from1 ? a.x1 : a.x2;
  const int cs = from1 ? a.c1 : a.c2, co = from1 ? ch0 : ch0 - a.c1;
  const int64_t base_px = (int64_t)n * a.hw;
  const int64_t p_begin = (int64_t)chunk * px_per_block;
  int64_t p_end = p_begin + px_per_block;
  if (p_end > a.hw) p_end = a.hw;
  int64_t p = p_begin + prow;
  if (ACT) gn_moments(a, n, s_mean, s_rstd, s_t);
  if (active) {
    float gah[8], ybh[8], sA[8], sB[8];   // y / 2 = x * gah + ybh
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ACT) {
        const int ch = ch0 + j, g = ch / cpg;
        const float ga = a.gamma[ch] * s_rstd[g];
        gah[j] = 0.5f * ga;
        ybh[j] = 0.5f * (a.beta[ch] - s_mean[g] * ga);
      }
      sA[j] = 0.f; sB[j] = 0.f;
    }
    // GB_ILP pixels per iteration: all loads of an iteration are issued before the first use; two resident CTAs per SM
    // keep ~64 KB in flight
    for (; p < p_end; p += (int64_t)ppi * GB_ILP) {
      uint4 rx[GB_ILP], rd[GB_ILP];
#pragma unroll
      for (int u = 0; u < GB_ILP; ++u) {
        const int64_t pp = p + (int64_t)u * ppi;
        if (pp < p_end) {
          rx[u] = ldg_nc_v4(src + (base_px + pp) * cs + co);
          rd[u] = ldg_nc_v4(a.dy + (base_px + pp) * C + ch0);
        }
      }
#pragma unroll
      for (int u = 0; u < GB_ILP; ++u) {
        const int64_t pp = p + (int64_t)u * ppi;
        if (pp < p_end) {
          const __half2* hx = reinterpret_cast<const __half2*>(&rx[u]);
          __half2* hd = reinterpret_cast<__half2*>(&rd[u]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fx = __half22float2(hx[k]);
            __half2 g2 = hd[k];
            if (ACT) {
              const __half2 h2 = __floats2half2_rn(fmaf(fx.x, gah[2 * k], ybh[2 * k]),
                                                   fmaf(fx.y, gah[2 * k + 1], ybh[2 * k + 1]));
              g2 = __hmul2(g2, silu_grad_h2(h2));
            }
            const float2 g = __half22float2(g2);   // exactly the g the second pass recomputes (same instructions)
            sA[2 * k] += g.x;
            sA[2 * k + 1] += g.y;
            sB[2 * k] = fmaf(g.x, fx.x, sB[2 * k]);
            sB[2 * k + 1] = fmaf(g.y, fx.y, sB[2 * k + 1]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_part[0][prow * C + ch0 + j] = sA[j];
      s_part[1][prow * C + ch0 + j] = sB[j];
    }
  }
  __syncthreads();
  float* o = a.partial + ((int64_t)n * (a.chunks + 1) + chunk) * C * 2;
  for (int c = threadIdx.x; c < C; c += GB_THREADS) {
    float tA = 0.f, tB = 0.f;
    for (int r = 0; r < ppi; ++r) {  // fixed order
      tA += s_part[0][r * C + c];
      tB += s_part[1][r * C + c];
    }
    reinterpret_cast<float2*>(o)[c] = make_float2(tA, tB);
  }
}
template <bool ACT>
__global__ void __launch_bounds__(GB_THREADS, 2) gn_bwd_stats_kernel(const GnBwdArgs a) {
  pdl_sync();
  gn_bwd_stats_body<ACT>(a, blockIdx.y, blockIdx.x, a.px_per_block);
}

// Between the passes: the chunk partials of a sample are summed ONCE (fixed order), per slice of whole groups, into the
// per-channel coefficients of dx.  Every CTA of the apply pass used to redo this (its own group moments from the
// integer totals + chunks x C x 2 floats of partials): for the 32 x 32 / 64 x 64 layers (512-1024 channels, few pixels)
// that prologue moved as many bytes as the pass itself.  grid (slices, n); a slice = as many whole groups as fit 256
// channels.  Same sums in the same order as before: results are bit-identical to the in-CTA form.
__global__ void __launch_bounds__(GB_THREADS) gn_bwd_coef_kernel(const GnBwdArgs a, const int groups_per_slice) {
  const int C = a.c1 + a.c2, cpg = C / a.groups, n = blockIdx.y;
  const int g0 = blockIdx.x * groups_per_slice;
  const int ng = min(groups_per_slice, a.groups - g0);
  const int c0 = g0 * cpg, nc = ng * cpg;
  __shared__ float s_mean[GB_MAX_GROUPS], s_rstd[GB_MAX_GROUPS], s_m1[GB_MAX_GROUPS], s_m2[GB_MAX_GROUPS];
  __shared__ float s_gA[GB_THREADS], s_gB[GB_THREADS];
  pdl_sync();
  const int tid = threadIdx.x;
  // group moments from the exact integer totals: identical arithmetic to gn_apply_kernel / gn_moments (one thread per
  // group walks its channels; integer sums are order-independent)
  if (tid < ng) {
    const int g = g0 + tid;
    long long ts = 0, tq = 0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const longlong2 tv = *reinterpret_cast<const longlong2*>(
          c < a.c1 ? a.st1 + ((int64_t)n * a.c1 + c) * 2 : a.st2 + ((int64_t)n * a.c2 + (c - a.c1)) * 2);
      ts += tv.x; tq += tv.y;
    }
    const double mg = (double)ts * a.inv_cnt_s;
    double vg = (double)tq * a.inv_cnt_q - mg * mg;
    if (vg < 0.0) vg = 0.0;
    s_mean[tid] = (float)mg;
    s_rstd[tid] = rsqrtf((float)vg + a.eps);
  }
  float tA = 0.f, tBx = 0.f, gm = 0.f, bt = 0.f;
  if (tid < nc) {
    const int c = c0 + tid;
    const float* pp = a.partial + (int64_t)n * (a.chunks + 1) * C * 2;
    for (int k = 0; k < a.chunks; ++k) {  // fixed order
      const float2 t = __ldcg(reinterpret_cast<const float2*>(pp + (int64_t)k * C * 2) + c);
      tA += t.x; tBx += t.y;
    }
    gm = a.gamma[c]; bt = a.beta[c];
  }
  __syncthreads();
  float tB = 0.f;
  if (tid < nc) {
    const int c = c0 + tid, gl = tid / cpg;
    tB = s_rstd[gl] * (tBx - s_mean[gl] * tA);  // sum g * xh
    float* red = a.partial + ((int64_t)n * (a.chunks + 1) + a.chunks) * C * 2;   // per-sample (A_c, B_c) for the params
    reinterpret_cast<float2*>(red)[c] = make_float2(tA, tB);
    s_gA[tid] = gm * tA;
    s_gB[tid] = gm * tB;
  }
  __syncthreads();
  if (tid < ng) {
    float m1 = 0.f, m2 = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { m1 += s_gA[c]; m2 += s_gB[c]; }
    const float inv_cnt = (float)(1.0 / ((double)a.hw * (double)cpg));
    s_m1[tid] = m1 * inv_cnt;
    s_m2[tid] = m2 * inv_cnt;
  }
  __syncthreads();
  if (tid < nc) {
    const int c = c0 + tid, gl = tid / cpg;
    const float mu = s_mean[gl], rs = s_rstd[gl];
    const float ga = gm * rs;
    reinterpret_cast<float4*>(a.coef)[(int64_t)n * C + c] =
        make_float4(ga, 0.5f * (bt - mu * ga), -rs * rs * s_m2[gl], rs * (mu * rs * s_m2[gl] - s_m1[gl]));
  }
}

// Pass 2.  dx = ga * g + pc * x + qc with per-channel pc = -rstd^2 m2, qc = rstd (mean rstd m2 - m1).
#ifndef GB_APPLY_CTAS
#define GB_APPLY_CTAS 2
#endif
#ifndef GB_APPLY_ILP
#define GB_APPLY_ILP 3
#endif
// ADD / ACC / CSUM / OSUM: whether the call has an addend, accumulates into a destination, wants the column sums of the
// GroupNorm term / of the stored values.  Compile-time: the kernel is instruction-issue-bound (ncu: 61 % of all issue
// slots at 23 % occupancy, profiles/r2o_ncu_full_gn_bwd_streamed_vs_register.txt), and the unpack + add sequences of
// the absent operands were a quarter of the inner loop.
template <int ILP, bool ADD, bool ACC, bool CSUM, bool OSUM>
__device__ __forceinline__ void gn_bwd_apply_body(const GnBwdArgs& a, const int n, const int part, const int nparts,
                                                  const int64_t px_per_block) {
  const int C = a.c1 + a.c2, V = C >> 3;
  const int ppi = GB_THREADS / V;
  __shared__ float s_part[2][GB_THREADS * 8];  // column sums
  const bool active = (int)threadIdx.x < ppi * V;
  const int v = active ? threadIdx.x % V : 0, prow = active ? threadIdx.x / V : 0;
  const int ch0 = v << 3;
  const bool from1 = ch0 < a.c1;
  const __half* src = from1 ? a.x1 : a.x2;
  __half* dst = from1 ? a.dx1 : a.dx2;
  const int accum = from1 ? a.acc1 : a.acc2;
  const int cs = from1 ? a.c1 : a.c2, co = from1 ? ch0 : ch0 - a.c1;
  const int64_t base_px = (int64_t)n * a.hw;
  const int64_t p_begin = (int64_t)part * px_per_block;
  int64_t p_end = p_begin + px_per_block;
  if (p_end > a.hw) p_end = a.hw;
  int64_t p = p_begin + prow;
  const bool want_osum = OSUM && (from1 ? a.osum1 : a.osum2) != nullptr;
  float cs_acc[8], os_acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs_acc[j] = 0.f; os_acc[j] = 0.f; }
  if (active) {
    float ga[8], gah[8], ybh[8], pc[8], qc[8];   // y / 2 = x * gah + ybh
    {
      const float4* cb = reinterpret_cast<const float4*>(a.coef) + (int64_t)n * C + ch0;   // gn_bwd_coef_kernel
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(cb + j);
        ga[j] = t.x; gah[j] = 0.5f * t.x; ybh[j] = t.y; pc[j] = t.z; qc[j] = t.w;
      }
    }
    for (; p < p_end; p += (int64_t)ppi * ILP) {
      uint4 rx[ILP], rd[ILP], ra[ADD ? ILP : 1], ro[ACC ? ILP : 1];
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        if (ADD) ra[u] = make_uint4(0, 0, 0, 0);
        if (ACC) ro[u] = make_uint4(0, 0, 0, 0);
        const int64_t pp = p + (int64_t)u * ppi;
        if (pp < p_end) {
          rx[u] = ldg_nc_v4(src + (base_px + pp) * cs + co);
          rd[u] = ldg_nc_v4(a.dy + (base_px + pp) * C + ch0);
          if (ADD) ra[u] = ldg_nc_v4(a.addend + (base_px + pp) * C + ch0);
          if (ACC && accum) ro[u] = *reinterpret_cast<const uint4*>(dst + (base_px + pp) * cs + co);
        }
      }
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        const int64_t pp = p + (int64_t)u * ppi;
        if (pp < p_end) {
          float fx[8], fd[8], fa[8], fo[8], r[8];
          unpack8(rx[u], fx);
          if (ADD) unpack8(ra[u], fa);
          if (ACC) unpack8(ro[u], fo);
          if (a.act) {   // g = dy * SiLU'(y): fp16x2, one MUFU per pair — bit-identical to what pass 1 summed
            __half2* hd = reinterpret_cast<__half2*>(&rd[u]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __half2 h2 = __floats2half2_rn(fmaf(fx[2 * k], gah[2 * k], ybh[2 * k]),
                                                   fmaf(fx[2 * k + 1], gah[2 * k + 1], ybh[2 * k + 1]));
              hd[k] = __hmul2(hd[k], silu_grad_h2(h2));
            }
          }
          unpack8(rd[u], fd);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = fmaf(ga[j], fd[j], fmaf(pc[j], fx[j], qc[j]));
            if (CSUM) cs_acc[j] += d;
            r[j] = d;
            if (ADD) r[j] += fa[j];       // same order of additions as before: (d + addend) + previous content
            if (ACC) r[j] += fo[j];
          }
          const uint4 packed = pack8(r);
          stg_v4(dst + (base_px + pp) * cs + co, packed);
          if (want_osum) {   // sums of what was actually stored (fp16-rounded), like a reader of the tensor would see
            float rr[8];
            unpack8(packed, rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) os_acc[j] += rr[j];
          }
        }
      }
    }
  }
  if (CSUM) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_part[0][prow * C + ch0 + j] = cs_acc[j];
    }
    __syncthreads();
    float* o = a.colsum + ((int64_t)n * nparts + part) * C;
    for (int c = threadIdx.x; c < C; c += GB_THREADS) {
      float t = 0.f;
      for (int r = 0; r < ppi; ++r) t += s_part[0][r * C + c];
      o[c] = t;
    }
  }
  if (OSUM) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_part[1][prow * C + ch0 + j] = os_acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GB_THREADS) {
      float* base = c < a.c1 ? a.osum1 : a.osum2;
      if (!base) continue;
      const int cw = c < a.c1 ? a.c1 : a.c2, cc = c < a.c1 ? c : c - a.c1;
      float t = 0.f;
      for (int r = 0; r < ppi; ++r) t += s_part[1][r * C + c];
      base[((int64_t)n * nparts + part) * cw + cc] = t;
    }
  }
}

template <bool ADD, bool ACC, bool CSUM, bool OSUM>
__global__ void __launch_bounds__(GB_THREADS, GB_APPLY_CTAS) gn_bwd_apply_kernel(const GnBwdArgs a) {
  pdl_sync();
  // four pixels in flight per thread where the registers allow it (1-3 %); with an addend AND a previous content that is
  // 4 x 4 uint4 of loads and spills at two CTAs per SM: three there
  gn_bwd_apply_body<(ADD && ACC) ? GB_APPLY_ILP : GB_APPLY_ILP + 1, ADD, ACC, CSUM, OSUM>(a, blockIdx.y, blockIdx.x,
                                                                                          gridDim.x, a.px_per_block);
}
typedef void (*GnBwdApplyFn)(const GnBwdArgs);
template <int I>
static GnBwdApplyFn gn_bwd_apply_variant() {
  return gn_bwd_apply_kernel<(I & 1) != 0, (I & 2) != 0, (I & 4) != 0, (I & 8) != 0>;
}
static GnBwdApplyFn gn_bwd_apply_pick(bool add, bool acc, bool csum, bool osum) {
  static const GnBwdApplyFn table[16] = {
      gn_bwd_apply_variant<0>(),  gn_bwd_apply_variant<1>(),  gn_bwd_apply_variant<2>(),  gn_bwd_apply_variant<3>(),
      gn_bwd_apply_variant<4>(),  gn_bwd_apply_variant<5>(),  gn_bwd_apply_variant<6>(),  gn_bwd_apply_variant<7>(),
      gn_bwd_apply_variant<8>(),  gn_bwd_apply_variant<9>(),  gn_bwd_apply_variant<10>(), gn_bwd_apply_variant<11>(),
      gn_bwd_apply_variant<12>(), gn_bwd_apply_variant<13>(), gn_bwd_apply_variant<14>(), gn_bwd_apply_variant<15>()};
  return table[(add ? 1 : 0) | (acc ? 2 : 0) | (csum ? 4 : 0) | (osum ? 8 : 0)];
}

// column sums of an fp16 [rows][C] tensor: per-block partials [blocks][C] (fp32), fixed order inside a block
__global__ void __launch_bounds__(GB_THREADS) colsum_h16_kernel(const __half* __restrict__ x, int64_t rows, int C,
                                                                float* __restrict__ partial, int64_t rows_per_block) {
  const int V = C >> 3, ppi = GB_THREADS / V;
  __shared__ float s_part[GB_THREADS * 8];
  const bool active = (int)threadIdx.x < ppi * V;
  const int v = active ? threadIdx.x % V : 0, prow = active ? threadIdx.x / V : 0;
  pdl_sync();
  if (active) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    for (int64_t r = r0 + prow; r < r1; r += 2 * ppi) {
      const uint4 a0 = ldg_nc_v4(x + r * C + (v << 3));
      uint4 a1 = make_uint4(0, 0, 0, 0);
      if (r + ppi < r1) a1 = ldg_nc_v4(x + (r + ppi) * C + (v << 3));
      float f0[8], f1[8];
      unpack8(a0, f0); unpack8(a1, f1);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f0[j] + f1[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[prow * C + (v << 3) + j] = s[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GB_THREADS) {
    float t = 0.f;
    for (int r = 0; r < ppi; ++r) t += s_part[r * C + c];
    partial[(int64_t)blockIdx.x * C + c] = t;
  }
}

// Every finaliser of one backward pass (d gamma / d beta of each GroupNorm, bias / time-embedding column sums) in ONE
// launch from a device-side job table: ~100 launches of 10 us each otherwise.  Same body, same fixed order.
__global__ void __launch_bounds__(256) reduce_rows_batched_kernel(const dsg_reduce_job* __restrict__ jobs, int njobs) {
  __shared__ float smem_f[RR_SMEM_FLOATS];
  __shared__ int s_job;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;   // last job whose first block is <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_job = lo;
  }
  __syncthreads();
  const dsg_reduce_job j = jobs[s_job];
  const int bx = (int)blockIdx.x - j.block_begin;
  if (j.comps == 2)
    reduce_rows_body<2>(j.src, j.n, j.parts, j.c, j.sample_stride, j.part_stride, j.per_n, j.per_n_stride, j.per_n_off,
                        j.inv_scale, j.out0, j.out0b, j.out1, bx, smem_f);
  else
    reduce_rows_body<1>(j.src, j.n, j.parts, j.c, j.sample_stride, j.part_stride, j.per_n, j.per_n_stride, j.per_n_off,
                        j.inv_scale, j.out0, j.out0b, j.out1, bx, smem_f);
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_gn_bwd(const void* dy, const void* x1, int32_t c1, const void* stats1, const void* x2, int32_t c2,
               const void* stats2, const float* gamma, const float* beta, float eps, int32_t act, float* partial,
               int32_t chunks, const void* addend, void* dx1, int32_t acc1, void* dx2, int32_t acc2, float* colsum,
               float* osum1, float* osum2, int32_t parts, int32_t n, int64_t hw, int32_t groups, void* stream) {
  DSG_CHECK_ARG(dy && x1 && stats1 && dx1 && c1 > 0 && c1 % 8 == 0, "dsg_gn_bwd: dy/x1/stats1/dx1 null or bad c1");
  DSG_CHECK_ARG((x2 == nullptr) == (c2 == 0) && (x2 == nullptr) == (stats2 == nullptr) &&
                    (x2 == nullptr) == (dx2 == nullptr) && c2 % 8 == 0,
                "dsg_gn_bwd: x2/stats2/dx2/c2 mismatch");
  const int C = c1 + c2;
  DSG_CHECK_ARG(groups > 0 && groups <= GB_MAX_GROUPS && C % groups == 0 && C <= GB_MAX_C && C / groups <= GB_THREADS,
                "dsg_gn_bwd: bad groups %d for C=%d", groups, C);
  DSG_CHECK_ARG(gamma && beta && partial && chunks >= 1 && chunks <= GB_MAX_CHUNKS, "dsg_gn_bwd: bad partial/chunks");
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && hw > 0, "dsg_gn_bwd: bad n/hw");
  DSG_CHECK_ARG(parts >= 0 && parts <= 65535 && (parts > 0 || (!colsum && !osum1 && !osum2)),
                "dsg_gn_bwd: column sums need an explicit parts > 0");
  DSG_CHECK_ARG(osum2 == nullptr || x2 != nullptr, "dsg_gn_bwd: osum2 without x2");
  DSG_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)dx1 | (uintptr_t)dx2 |
                  (uintptr_t)addend | (uintptr_t)stats1 | (uintptr_t)stats2 | (uintptr_t)partial) % 16) == 0,
                "dsg_gn_bwd: unaligned pointer");
  if (n == 0) return DSG_OK;
  GnBwdArgs a;
  a.dy = (const __half*)dy;
  a.x1 = (const __half*)x1; a.c1 = c1; a.st1 = (const long long*)stats1;
  a.x2 = (const __half*)x2; a.c2 = c2; a.st2 = (const long long*)stats2;
  a.gamma = gamma; a.beta = beta; a.eps = eps; a.act = act;
  a.partial = partial; a.chunks = chunks;
  a.coef = partial + (int64_t)n * (chunks + 1) * C * 2;
  a.addend = (const __half*)addend;
  a.dx1 = (__half*)dx1; a.acc1 = acc1; a.dx2 = (__half*)dx2; a.acc2 = acc2;
  a.colsum = colsum; a.osum1 = osum1; a.osum2 = osum2; a.hw = hw; a.groups = groups;
  a.inv_cnt_s = 1.0 / 16777216.0 / ((double)hw * (double)(C / groups));
  a.inv_cnt_q = 1.0 / 1048576.0 / ((double)hw * (double)(C / groups));
  cudaStream_t st = (cudaStream_t)stream;
  int64_t ctas = parts;
  if (ctas == 0) {
    ctas = (num_sms() * GB_APPLY_CTAS) / n;   // one wave
    if (ctas < 1) ctas = 1;
    const int64_t max_ctas = ceil_div64(hw, 32);
    if (ctas > max_ctas) ctas = max_ctas;
  }
  a.px_per_block = ceil_div64(hw, chunks);
  if (act)
    launch_k(gn_bwd_stats_kernel<true>, dim3((unsigned)chunks, n), dim3(GB_THREADS), 0, st, a);
  else
    launch_k(gn_bwd_stats_kernel<false>, dim3((unsigned)chunks, n), dim3(GB_THREADS), 0, st, a);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd/stats");
  {
    const int cpg = C / groups;
    int gps = GB_THREADS / cpg;   // whole groups per 256-channel slice
    if (gps < 1) gps = 1;
    if (gps > groups) gps = groups;
    launch_k(gn_bwd_coef_kernel, dim3((unsigned)ceil_div(groups, gps), n), dim3(GB_THREADS), 0, st, a, gps);
    DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd/coef");
  }
  a.px_per_block = ceil_div64(hw, ctas);
  launch_k(gn_bwd_apply_pick(addend != nullptr, acc1 != 0 || (acc2 != 0 && c2 > 0), colsum != nullptr,
                             osum1 != nullptr || osum2 != nullptr),
           dim3((unsigned)ctas, n), dim3(GB_THREADS), 0, st, a);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd/apply");
  return DSG_OK;
}

int dsg_gn_bwd_params(const float* partial, int32_t n, int32_t chunks, int32_t c, const float* inv_scale,
                      float* dgamma, float* dbeta, void* stream) {
  DSG_CHECK_ARG(partial && dgamma && dbeta && n >= 0 && chunks >= 1 && c > 0, "dsg_gn_bwd_params: bad args");
  // per-sample slot `chunks` of every sample holds (sum g, sum g * xh) per channel: d beta, d gamma
  launch_k(reduce_rows_kernel<2>, dim3(ceil_div(c, 32)), dim3(256), 0, (cudaStream_t)stream,
           partial + (int64_t)chunks * c * 2, n, 1, c, (int64_t)(chunks + 1) * c * 2, (int64_t)0, (float*)nullptr, 0, 0,
           inv_scale, dbeta, (float*)nullptr, dgamma);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd_params");
  return DSG_OK;
}

int dsg_reduce_rows_batched(const dsg_reduce_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream) {
  DSG_CHECK_ARG(jobs_dev && njobs >= 1 && total_blocks >= 1, "dsg_reduce_rows_batched: bad args");
  reduce_rows_batched_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, njobs);
  DSG_CUDA_LAUNCH_CHECK("dsg_reduce_rows_batched");
  return DSG_OK;
}

int dsg_colsum_h16(const void* x, int64_t rows, int32_t c, float* partial, int32_t parts, void* stream) {
  DSG_CHECK_ARG(x && partial && rows >= 0 && c > 0 && c % 8 == 0 && c <= GB_MAX_C && parts >= 1,
                "dsg_colsum_h16: bad args");
  DSG_CHECK_ARG((uintptr_t)x % 16 == 0, "dsg_colsum_h16: unaligned pointer");
  const int64_t rpb = ceil_div64(rows > 0 ? rows : 1, parts);
  launch_k(colsum_h16_kernel, dim3(parts), dim3(GB_THREADS), 0, (cudaStream_t)stream, (const __half*)x, rows, c, partial,
           rpb);
  DSG_CUDA_LAUNCH_CHECK("dsg_colsum_h16");
  return DSG_OK;
}

int dsg_colsum_finalize(const float* partial, int32_t n, int32_t parts, int32_t c, float* per_n, int32_t per_n_stride,
                        int32_t per_n_off, const float* inv_scale, float* total, float* total2, void* stream) {
  DSG_CHECK_ARG(partial && n >= 0 && parts >= 1 && c > 0, "dsg_colsum_finalize: bad args");
  launch_k(reduce_rows_kernel<1>, dim3(ceil_div(c, 32)), dim3(256), 0, (cudaStream_t)stream, partial, n, parts, c,
           (int64_t)parts * c, (int64_t)c, per_n, per_n_stride, per_n_off, inv_scale, total, total2, (float*)nullptr);
  DSG_CUDA_LAUNCH_CHECK("dsg_colsum_finalize");
  return DSG_OK;
}
}
