// groupnorm_bwd_stream.cu — GroupNorm(+SiLU) backward, TMA-streamed form of groupnorm_bwd.cu (same contract and
// formulas; training path; replaces the autograd backward of torch.nn.GroupNorm(32, C) + SiLU of
// diffusers 0.20.0 models/resnet.py reached from DriveSceneGen/pipeline/training_pipeline.py:86
// `accelerator.backward(loss)`).
//
// groupnorm_bwd.cu loads through registers: a thread issues its 16-byte loads, waits, computes, and every CTA pays a
// prologue (group moments from the integer totals, the sum of up to 64 chunk partials per channel) during which it has
// nothing in flight — 4.0 TB/s of the 6.5 TB/s the HBM delivers.  Here a producer warp streams the tensors through a ring
// of shared-memory stages with 2-D TMA boxes (Cs channels x R pixels) and mbarriers, so loads are in flight regardless of
// what the eight consumer warps are doing (coefficient set-up, block reductions), and the per-sample bookkeeping is a
// kernel of its own:
//
//   work is cut into DOMAINS = (sample, slice of Cs <= 128 channels made of whole groups) — independent GroupNorm
//   problems — and every domain into P pixel ranges ("parts"); an ITEM = (domain, part), ~100-200 KB of x + dy.
//   Persistent CTAs (two per SM) take contiguous runs of items.
//
//   1. gn_bwd_sums_kernel    per item: sum g and sum g x per channel -> rows[domain][part][Cs][2]         (reads x, dy)
//   2. gn_bwd_coef_kernel    per domain: rows added in a fixed order (deterministic) -> (sum g, sum g xh) per channel
//                            (the d beta / d gamma rows), group means m1 / m2, and the per-channel coefficients of dx:
//                            coef[s][c] = (ga, ybh, pc, qc)                                                (a few KB)
//   3. gn_bwd_dx_kernel      per item: dx = ga g + pc x + qc (+ addend, + previous content), optional per-part column
//                            sums                                                           (reads x, dy (, ...), writes dx)
//
// A slice that straddles the boundary of the two concatenated sources is two boxes, each clipped by the TMA unit's
// out-of-bounds rule (zero fill), landing in two tiles of the stage; a thread picks the tile its channels live in.
// Measured and dropped before this form (profiles/r2o_gn_bwd_single_launch_*.txt): both passes in ONE persistent launch
// with the second pass lagging 1-3 L2-sized units behind the first (per-domain counters and flags in global memory).
// Every hop of the chain partial rows -> fence -> counter -> row sums -> coefficients -> fence -> flag -> fetch is a
// trip through a memory system that the kernel itself keeps saturated (~1.4 us each and more), the CTA that arrives
// last does the row sums and is therefore last again at the next unit, and the period per 17 MB unit came out at
// ~29 us whatever the lag: 2.7x slower than two launches.
#include "igemm_common.cuh"

namespace dsg {

constexpr int GS_CONS = 256;               // consumer threads (8 warps)
constexpr int GS_THREADS = GS_CONS + 32;   // + one producer warp
constexpr int GS_MAX_ST = 8;
constexpr int GS_MAX_CS = 128;             // channels per slice
constexpr int GS_MAX_GROUPS = 64;
constexpr int GS_DYN_BUDGET = 92 * 1024;   // ring per CTA (two CTAs per SM)

struct GsMaps { CUtensorMap dy, x1, x2, add, o1, o2; };

struct GsArgs {
  int c1, c2;
  const long long* st1; const long long* st2;
  const float* gamma; const float* beta;
  float eps; int act;
  float* red;            // per-sample (sum g, sum g xh) rows: red + s * red_stride floats, [C][2]
  int64_t red_stride;
  int has_add;
  __half* dx1; int acc1;
  __half* dx2; int acc2;
  float* colsum; float* osum1; float* osum2;   // optional [n][P][C] / [n][P][c1] / [n][P][c2]
  int64_t hw; int groups; int n;
  double inv_cnt_s, inv_cnt_q;
  // plan
  int Cs, J, D, P;       // channels per slice, slices per sample, domains = n * J, parts per domain
  int64_t q;             // pixels per part
  int R, RB;             // rows per stage = RB * ppi
  int nst, stage_bytes, tile_bytes;
  int off_xb, off_dy, off_add, off_oa, off_ob;   // byte offsets of the tiles inside a stage
  // workspace
  float* coef;           // [n][C][4]  ga, ybh, pc, qc   (y / 2 = x * ga / 2 + ybh)
  float* rows;           // [D][P][Cs][2]
};

// mbarrier wait with a 30 s bound (a protocol error must not hang the box)
__device__ __forceinline__ void gs_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 30000000000ull) {
        printf("dsg: gn_bwd_stream mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// d/dy [y sigmoid(y)] for two elements from h = y / 2 — the instruction sequence of groupnorm_bwd.cu::silu_grad_h2:
// ONE MUFU op per pair; both passes evaluate g with it, so the sums and dx see the same g bit for bit
__device__ __forceinline__ __half2 silu_grad_h2s(__half2 h) {
  uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(hi));
  const __half2 t = *reinterpret_cast<__half2*>(&ti);
  const __half2 one = __float2half2_rn(1.0f), half_ = __float2half2_rn(0.5f);
  const __half2 sg = __hfma2(half_, t, half_);
  const __half2 w = __hfma2(h, __hsub2(one, t), one);
  return __hmul2(sg, w);
}

struct GsShared {
  uint64_t full[GS_MAX_ST], empty[GS_MAX_ST];
  float part[2][GS_CONS * 8];     // block reductions
  float mean[GS_MAX_GROUPS], rstd[GS_MAX_GROUPS], m1[GS_MAX_GROUPS], m2[GS_MAX_GROUPS];
  unsigned long long tot[GS_MAX_GROUPS][2];
};

struct GsItem {
  int dom, part, s, c_lo;
  int64_t p0, p1;
  bool in2, straddle;
};
__device__ __forceinline__ void gs_item(const GsArgs& f, int64_t i, GsItem& it) {
  it.dom = (int)(i / f.P);
  it.part = (int)(i % f.P);
  it.s = it.dom / f.J;
  it.c_lo = (it.dom % f.J) * f.Cs;
  it.in2 = it.c_lo >= f.c1;
  it.straddle = it.c_lo < f.c1 && it.c_lo + f.Cs > f.c1;
  it.p0 = (int64_t)it.part * f.q;
  it.p1 = it.p0 + f.q;
  if (it.p1 > f.hw) it.p1 = f.hw;
}

// mean / rstd of the slice's groups from the exact integer totals (the arithmetic of gn_apply_kernel, bit for bit: the
// backward must see the forward's statistics); sh.mean / sh.rstd are indexed by the group's position inside the slice
template <bool ALL>
__device__ __forceinline__ void gs_sync() {
  if (ALL) __syncthreads(); else cons_sync();
}
template <bool ALL>
__device__ void gs_moments(const GsArgs& f, GsShared& sh, const int s, const int c_lo) {
  const int C = f.c1 + f.c2, cpg = C / f.groups, tid = threadIdx.x, ng = f.Cs / cpg;
  if (tid < ng) { sh.tot[tid][0] = 0ull; sh.tot[tid][1] = 0ull; }
  gs_sync<ALL>();
  if (tid < f.Cs) {
    const int c = c_lo + tid;
    const longlong2 tv = *reinterpret_cast<const longlong2*>(
        c < f.c1 ? f.st1 + ((int64_t)s * f.c1 + c) * 2 : f.st2 + ((int64_t)s * f.c2 + (c - f.c1)) * 2);
    if (tv.x != 0 || tv.y != 0) {
      atomicAdd(&sh.tot[tid / cpg][0], (unsigned long long)tv.x);
      atomicAdd(&sh.tot[tid / cpg][1], (unsigned long long)tv.y);
    }
  }
  gs_sync<ALL>();
  if (tid < ng) {
    const double mg = (double)(long long)sh.tot[tid][0] * f.inv_cnt_s;
    double vg = (double)(long long)sh.tot[tid][1] * f.inv_cnt_q - mg * mg;
    if (vg < 0.0) vg = 0.0;
    sh.mean[tid] = (float)mg;
    sh.rstd[tid] = rsqrtf((float)vg + f.eps);
  }
  gs_sync<ALL>();
}

// ------------------------------------------------------------------ the producer warp (both streaming kernels)
template <bool DX>
__device__ __forceinline__ void gs_produce(const GsMaps& maps, const GsArgs& f, GsShared& sh, uint8_t* ring, int64_t i0,
                                           int64_t i1, int lane) {
  if (lane == 0) {
    tma_prefetch_desc(&maps.dy); tma_prefetch_desc(&maps.x1);
    if (f.c2) tma_prefetch_desc(&maps.x2);
  }
  int st = 0;
  uint32_t ph = 0;
  for (int64_t i = i0; i < i1; ++i) {
    GsItem it;
    gs_item(f, i, it);
    const int64_t base = (int64_t)it.s * f.hw;
    const bool old_a = DX && (it.in2 ? f.acc2 : f.acc1), old_b = DX && it.straddle && f.acc2;
    const uint32_t bytes = (uint32_t)f.tile_bytes * (2u + (it.straddle ? 1u : 0u) + ((DX && f.has_add) ? 1u : 0u) +
                                                     (old_a ? 1u : 0u) + (old_b ? 1u : 0u));
    for (int64_t r0 = it.p0; r0 < it.p1; r0 += f.R) {
      if (lane == 0) {
        uint8_t* sb = ring + (size_t)st * f.stage_bytes;
        const int row = (int)(base + r0);
        gs_wait(&sh.empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&sh.full[st], bytes);
        if (it.in2) tma_load_2d(sb, &maps.x2, &sh.full[st], it.c_lo - f.c1, row);
        else tma_load_2d(sb, &maps.x1, &sh.full[st], it.c_lo, row);
        if (it.straddle) tma_load_2d(sb + f.off_xb, &maps.x2, &sh.full[st], it.c_lo - f.c1, row);
        tma_load_2d(sb + f.off_dy, &maps.dy, &sh.full[st], it.c_lo, row);
        if (DX) {
          if (f.has_add) tma_load_2d(sb + f.off_add, &maps.add, &sh.full[st], it.c_lo, row);
          if (old_a) {
            if (it.in2) tma_load_2d(sb + f.off_oa, &maps.o2, &sh.full[st], it.c_lo - f.c1, row);
            else tma_load_2d(sb + f.off_oa, &maps.o1, &sh.full[st], it.c_lo, row);
          }
          if (old_b) tma_load_2d(sb + f.off_ob, &maps.o2, &sh.full[st], it.c_lo - f.c1, row);
        }
      }
      if (++st == f.nst) { st = 0; ph ^= 1; }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void gs_setup(const GsArgs& f, GsShared& sh) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < f.nst; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], GS_CONS / 32); }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_sync();
}

// ------------------------------------------------------------------ 1. partial sums
__global__ void __launch_bounds__(GS_THREADS, 2) gn_bwd_sums_kernel(const __grid_constant__ GsMaps maps, const GsArgs f) {
  extern __shared__ __align__(128) uint8_t gs_dyn[];
  __shared__ GsShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = f.c1 + f.c2, Cs = f.Cs, V = Cs >> 3, ppi = GS_CONS / V, cpg = C / f.groups;
  const int64_t items = (int64_t)f.D * f.P;
  const int64_t i0 = items * blockIdx.x / gridDim.x, i1 = items * (blockIdx.x + 1) / gridDim.x;
  gs_setup(f, sh);
  if (warp == GS_CONS / 32) {
    gs_produce<false>(maps, f, sh, gs_dyn, i0, i1, lane);
    return;
  }
  const bool active = tid < ppi * V;
  const int v = active ? tid % V : 0, prow = active ? tid / V : 0;
  const int col = (v << 3) * 2;               // byte offset of the thread's vector inside a tile row
  int st = 0, cur_dom = -1;
  uint32_t ph = 0;
  float gah[8], ybh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { gah[j] = 0.f; ybh[j] = 0.f; }
  for (int64_t i = i0; i < i1; ++i) {
    GsItem it;
    gs_item(f, i, it);
    if (f.act && it.dom != cur_dom) {   // the SiLU argument's affine map of this thread's channels: y / 2 = x * gah + ybh
      gs_moments<false>(f, sh, it.s, it.c_lo);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int lc = (v << 3) + j, g = lc / cpg;
        const float ga = f.gamma[it.c_lo + lc] * sh.rstd[g];
        gah[j] = 0.5f * ga;
        ybh[j] = 0.5f * __fmaf_rn(-sh.mean[g], ga, f.beta[it.c_lo + lc]);
      }
      cur_dom = it.dom;
    }
    const int xoff = (it.straddle && it.c_lo + (v << 3) >= f.c1) ? f.off_xb : 0;
    float sA[8], sB[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sA[j] = 0.f; sB[j] = 0.f; }
    for (int64_t r0 = it.p0; r0 < it.p1; r0 += f.R) {
      const int nr = (int)((it.p1 - r0 < f.R) ? (it.p1 - r0) : f.R);
      const uint8_t* sb = gs_dyn + (size_t)st * f.stage_bytes;
      gs_wait(&sh.full[st], ph);
      if (active) {
        for (int k = 0; k < f.RB; ++k) {
          const int row = k * ppi + prow;
          if (row < nr) {
            const uint4 rx = *reinterpret_cast<const uint4*>(sb + xoff + (size_t)row * Cs * 2 + col);
            uint4 rd = *reinterpret_cast<const uint4*>(sb + f.off_dy + (size_t)row * Cs * 2 + col);
            const __half2* hx = reinterpret_cast<const __half2*>(&rx);
            __half2* hd = reinterpret_cast<__half2*>(&rd);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const float2 fx = __half22float2(hx[kk]);
              __half2 g2 = hd[kk];
              if (f.act) {
                const __half2 h2 = __floats2half2_rn(fmaf(fx.x, gah[2 * kk], ybh[2 * kk]),
                                                     fmaf(fx.y, gah[2 * kk + 1], ybh[2 * kk + 1]));
                g2 = __hmul2(g2, silu_grad_h2s(h2));
              }
              const float2 g = __half22float2(g2);
              sA[2 * kk] += g.x;
              sA[2 * kk + 1] += g.y;
              sB[2 * kk] = fmaf(g.x, fx.x, sB[2 * kk]);
              sB[2 * kk + 1] = fmaf(g.y, fx.y, sB[2 * kk + 1]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[st]);
      if (++st == f.nst) { st = 0; ph ^= 1; }
    }
    cons_sync();   // the previous item's readers of sh.part are done
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sh.part[0][prow * Cs + (v << 3) + j] = sA[j];
        sh.part[1][prow * Cs + (v << 3) + j] = sB[j];
      }
    }
    cons_sync();
    if (tid < Cs) {
      float tA = 0.f, tB = 0.f;
      for (int r = 0; r < ppi; ++r) {  // fixed order
        tA += sh.part[0][r * Cs + tid];
        tB += sh.part[1][r * Cs + tid];
      }
      reinterpret_cast<float2*>(f.rows + ((int64_t)it.dom * f.P + it.part) * Cs * 2)[tid] = make_float2(tA, tB);
    }
  }
}

// ------------------------------------------------------------------ 2. per-domain coefficients (grid = domains)
__global__ void __launch_bounds__(GS_CONS) gn_bwd_coef_kernel(const GsArgs f) {
  __shared__ GsShared sh;
  const int tid = threadIdx.x;
  const int C = f.c1 + f.c2, Cs = f.Cs, cpg = C / f.groups, ng = Cs / cpg;
  const int dom = blockIdx.x, s = dom / f.J, c_lo = (dom % f.J) * Cs;
  pdl_sync();
  float gm = 0.f, bt = 0.f;
  if (tid < Cs) { gm = f.gamma[c_lo + tid]; bt = f.beta[c_lo + tid]; }
  // rows of the domain, added in a fixed order: thread (kgi, col) takes rows kgi, kgi + kg, ...; the kg partial sums of a
  // column are then added in index order
  float4* fin = reinterpret_cast<float4*>(&sh.part[0][0]);           // [kg][Q] <= 4 KB
  float2* chs = reinterpret_cast<float2*>(&sh.part[0][0]) + 512;     // [Cs] (sum g, sum g x), then (gamma A, gamma B)
  const int Q = Cs >> 1;                                             // float4 per row (<= 64)
  const float4* rows = reinterpret_cast<const float4*>(f.rows + (int64_t)dom * f.P * Cs * 2);
  const int kg = GS_CONS / Q;
  const int col = tid % Q, kgi = tid / Q;
  if (kgi < kg) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = kgi; k < f.P; k += 8 * kg) {   // eight independent loads in flight, added in index order
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (k + u * kg < f.P) ? __ldcg(rows + (int64_t)(k + u * kg) * Q + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    fin[kgi * Q + col] = acc;
  }
  gs_moments<true>(f, sh, s, c_lo);   // (its barriers also publish fin)
  if (tid < Q) {
    float4 acc = fin[tid];
    for (int j = 1; j < kg; ++j) {
      const float4 v = fin[j * Q + tid];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    chs[2 * tid] = make_float2(acc.x, acc.y);
    chs[2 * tid + 1] = make_float2(acc.z, acc.w);
  }
  __syncthreads();
  if (tid < Cs) {
    const float2 t = chs[tid];
    const int g = tid / cpg;
    const float tB = sh.rstd[g] * (t.y - sh.mean[g] * t.x);   // sum g * xh
    reinterpret_cast<float2*>(f.red + (int64_t)s * f.red_stride)[c_lo + tid] = make_float2(t.x, tB);
    chs[tid] = make_float2(gm * t.x, gm * tB);
  }
  __syncthreads();
  if (tid < ng) {
    float m1 = 0.f, m2 = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { m1 += chs[c].x; m2 += chs[c].y; }
    const float inv_cnt = (float)(1.0 / ((double)f.hw * (double)cpg));
    sh.m1[tid] = m1 * inv_cnt;
    sh.m2[tid] = m2 * inv_cnt;
  }
  __syncthreads();
  if (tid < Cs) {
    const int g = tid / cpg;
    const float mu = sh.mean[g], rs = sh.rstd[g];
    const float ga = gm * rs;
    // ybh: the same expression as gn_bwd_sums_kernel's (both passes see the same SiLU argument)
    reinterpret_cast<float4*>(f.coef)[(int64_t)s * C + c_lo + tid] =
        make_float4(ga, 0.5f * __fmaf_rn(-mu, ga, bt), -rs * rs * sh.m2[g], rs * (mu * rs * sh.m2[g] - sh.m1[g]));
  }
}

// ------------------------------------------------------------------ 3. dx
__global__ void __launch_bounds__(GS_THREADS, 2) gn_bwd_dx_kernel(const __grid_constant__ GsMaps maps, const GsArgs f) {
  extern __shared__ __align__(128) uint8_t gs_dyn[];
  __shared__ GsShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = f.c1 + f.c2, Cs = f.Cs, V = Cs >> 3, ppi = GS_CONS / V;
  const int64_t items = (int64_t)f.D * f.P;
  const int64_t i0 = items * blockIdx.x / gridDim.x, i1 = items * (blockIdx.x + 1) / gridDim.x;
  gs_setup(f, sh);
  if (warp == GS_CONS / 32) {
    gs_produce<true>(maps, f, sh, gs_dyn, i0, i1, lane);
    return;
  }
  const bool active = tid < ppi * V;
  const int v = active ? tid % V : 0, prow = active ? tid / V : 0;
  const int col = (v << 3) * 2;
  int st = 0, cur_dom = -1;
  uint32_t ph = 0;
  float gah[8], ga[8], ybh[8], pc[8], qc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { gah[j] = 0.f; ga[j] = 0.f; ybh[j] = 0.f; pc[j] = 0.f; qc[j] = 0.f; }
  for (int64_t i = i0; i < i1; ++i) {
    GsItem it;
    gs_item(f, i, it);
    const int ch = it.c_lo + (v << 3);
    if (it.dom != cur_dom) {
      const float4* cb = reinterpret_cast<const float4*>(f.coef) + (int64_t)it.s * C + ch;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(cb + j);
        ga[j] = t.x; gah[j] = 0.5f * t.x; ybh[j] = t.y; pc[j] = t.z; qc[j] = t.w;
      }
      cur_dom = it.dom;
    }
    const bool from1 = ch < f.c1;
    const int xoff = (it.straddle && !from1) ? f.off_xb : 0;
    const int ooff = (it.straddle && !from1) ? f.off_ob : f.off_oa;
    const int accum = from1 ? f.acc1 : f.acc2;
    __half* dst = from1 ? f.dx1 + ch : f.dx2 + (ch - f.c1);
    const int cs = from1 ? f.c1 : f.c2;
    const bool want_osum = (from1 ? f.osum1 : f.osum2) != nullptr;
    const int64_t base = (int64_t)it.s * f.hw;
    float cs_acc[8], os_acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { cs_acc[j] = 0.f; os_acc[j] = 0.f; }
    for (int64_t r0 = it.p0; r0 < it.p1; r0 += f.R) {
      const int nr = (int)((it.p1 - r0 < f.R) ? (it.p1 - r0) : f.R);
      const uint8_t* sb = gs_dyn + (size_t)st * f.stage_bytes;
      gs_wait(&sh.full[st], ph);
      if (active) {
        for (int k = 0; k < f.RB; ++k) {
          const int row = k * ppi + prow;
          if (row < nr) {
            const uint4 rx = *reinterpret_cast<const uint4*>(sb + xoff + (size_t)row * Cs * 2 + col);
            uint4 rd = *reinterpret_cast<const uint4*>(sb + f.off_dy + (size_t)row * Cs * 2 + col);
            uint4 ra = make_uint4(0, 0, 0, 0), ro = make_uint4(0, 0, 0, 0);
            if (f.has_add) ra = *reinterpret_cast<const uint4*>(sb + f.off_add + (size_t)row * Cs * 2 + col);
            if (accum) ro = *reinterpret_cast<const uint4*>(sb + ooff + (size_t)row * Cs * 2 + col);
            float fx[8], fd[8], fa[8], fo[8], r[8];
            unpack8(rx, fx); unpack8(ra, fa); unpack8(ro, fo);
            if (f.act) {   // g = dy * SiLU'(y): bit-identical to what the first pass summed
              __half2* hd = reinterpret_cast<__half2*>(&rd);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const __half2 h2 = __floats2half2_rn(fmaf(fx[2 * kk], gah[2 * kk], ybh[2 * kk]),
                                                     fmaf(fx[2 * kk + 1], gah[2 * kk + 1], ybh[2 * kk + 1]));
                hd[kk] = __hmul2(hd[kk], silu_grad_h2s(h2));
              }
            }
            unpack8(rd, fd);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float d = fmaf(ga[j], fd[j], fmaf(pc[j], fx[j], qc[j]));
              cs_acc[j] += d;
              r[j] = (d + fa[j]) + fo[j];
            }
            const uint4 packed = pack8(r);
            stg_v4(dst + (base + r0 + row) * cs, packed);
            if (want_osum) {   // sums of what was actually stored (fp16-rounded), like a reader of the tensor would see
              float rr[8];
              unpack8(packed, rr);
#pragma unroll
              for (int j = 0; j < 8; ++j) os_acc[j] += rr[j];
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[st]);
      if (++st == f.nst) { st = 0; ph ^= 1; }
    }
    if (f.colsum) {
      cons_sync();
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh.part[0][prow * Cs + (v << 3) + j] = cs_acc[j];
      }
      cons_sync();
      if (tid < Cs) {
        float t = 0.f;
        for (int r = 0; r < ppi; ++r) t += sh.part[0][r * Cs + tid];
        f.colsum[((int64_t)it.s * f.P + it.part) * C + it.c_lo + tid] = t;
      }
    }
    if (f.osum1 || f.osum2) {
      cons_sync();
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh.part[1][prow * Cs + (v << 3) + j] = os_acc[j];
      }
      cons_sync();
      if (tid < Cs) {
        const int c = it.c_lo + tid;
        float* ob = c < f.c1 ? f.osum1 : f.osum2;
        if (ob) {
          const int cw = c < f.c1 ? f.c1 : f.c2, cc = c < f.c1 ? c : c - f.c1;
          float t = 0.f;
          for (int r = 0; r < ppi; ++r) t += sh.part[1][r * Cs + tid];
          ob[((int64_t)it.s * f.P + it.part) * cw + cc] = t;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ host side: the plan
struct GsPlan {
  bool ok;
  int Cs, J, D, P, R, RB, nst, stage_bytes, tile_bytes, ctas;
  int64_t q;
  int off_xb, off_dy, off_add, off_oa, off_ob;
  size_t dyn_smem;
  int64_t ws_bytes, off_rows;
};

static int gs_max_ctas() {
  static std::atomic<int> cached[kMaxDevices];
  const int d = current_device();
  int v = cached[d].load(std::memory_order_relaxed);
  if (v != 0) return v;
  cudaFuncSetAttribute(gn_bwd_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_DYN_BUDGET);
  cudaFuncSetAttribute(gn_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_DYN_BUDGET);
  (void)cudaGetLastError();
  v = 2 * num_sms();
  cached[d].store(v, std::memory_order_relaxed);
  return v;
}

static int gs_gcd(int a, int b) { return b ? gs_gcd(b, a % b) : a; }

static GsPlan gs_plan(int n, int64_t hw, int c1, int c2, int groups, bool has_add, bool has_acc1, bool has_acc2) {
  GsPlan p = {};
  const int C = c1 + c2;
  static const int enabled = getenv("DSG_GN_BWD_STREAM") ? atoi(getenv("DSG_GN_BWD_STREAM")) : 1;
  static const double item_kb = getenv("DSG_GN_BWD_ITEM_KB") ? atof(getenv("DSG_GN_BWD_ITEM_KB")) : 192.0;
  if (!enabled) return p;
  if (c1 <= 0 || c1 % 8 || c2 < 0 || c2 % 8 || groups <= 0 || groups > GS_MAX_GROUPS || C % groups || n <= 0 || hw <= 0)
    return p;
  if ((double)n * (double)hw > 2.0e9) return p;   // TMA row coordinates are int32
  const int G = gs_max_ctas();
  // slice width: whole groups and whole 16-byte vectors, <= 128 channels, the widest that divides C
  const int cpg = C / groups;
  const int step = cpg / gs_gcd(cpg, 8) * 8;   // lcm(cpg, 8)
  int Cs = 0;
  for (int w = step; w <= GS_MAX_CS && w <= C; w += step)
    if (C % w == 0) Cs = w;
  if (!Cs) return p;
  p.Cs = Cs; p.J = C / Cs; p.D = n * p.J;
  if (p.D > 65535 * 16) return p;
  const int V = Cs / 8, ppi = GS_CONS / V;
  bool straddle = false;
  for (int j = 0; j < p.J; ++j) straddle |= (j * Cs < c1 && (j + 1) * Cs > c1);
  const int ntiles_max = 2 + (straddle ? 1 : 0) + (has_add ? 1 : 0) + ((has_acc1 || has_acc2) ? 1 : 0) +
                         ((straddle && has_acc2) ? 1 : 0);
  p.RB = ntiles_max > 2 ? 1 : 2;
  p.R = ppi * p.RB;
  p.tile_bytes = p.R * Cs * 2;
  const int tstride = (p.tile_bytes + 127) & ~127;
  int off = tstride;
  p.off_xb = off; if (straddle) off += tstride;
  p.off_dy = off; off += tstride;
  p.off_add = off; if (has_add) off += tstride;
  p.off_oa = off; if (has_acc1 || has_acc2) off += tstride;
  p.off_ob = off; if (straddle && has_acc2) off += tstride;
  p.stage_bytes = off;
  p.nst = GS_DYN_BUDGET / p.stage_bytes;
  if (p.nst > GS_MAX_ST) p.nst = GS_MAX_ST;
  if (p.nst < 2) return p;
  p.dyn_smem = (size_t)p.nst * p.stage_bytes;
  // parts: items of ~192 KB of x + dy, but at least ~4 items per CTA of a full grid
  const double dom_bytes = (double)hw * Cs * 4.0;
  double item = item_kb * 1024.0;
  const double fair = dom_bytes * p.D / (4.0 * G);
  if (item > fair) item = fair;
  int64_t q = (int64_t)(item / (Cs * 4.0));
  if (q < p.R) q = p.R;
  q = (q + p.R - 1) / p.R * p.R;               // whole stages: no box reads pixels of the next part
  if (q > hw) q = (hw + p.R - 1) / p.R * p.R;
  p.q = q;
  p.P = (int)((hw + q - 1) / q);
  if (p.P > 65535) return p;
  const int64_t items = (int64_t)p.D * p.P;
  p.ctas = (int)(items < G ? items : G);
  int64_t o = (int64_t)n * C * 4 * 4;          // coef
  o = (o + 255) & ~(int64_t)255;
  p.off_rows = o; o += (int64_t)p.D * p.P * Cs * 2 * 4;
  p.ws_bytes = o;
  p.ok = true;
  return p;
}

// [rows][C] fp16 tensor, box = Cs channels x R rows, dense in shared memory
static int gs_make_map(CUtensorMap* m, const void* ptr, int C, int64_t rows, int Cs, int R) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DSG_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)Cs, (cuuint32_t)R};   // may exceed C: the columns beyond the tensor are zero-filled
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(gn_bwd_stream) failed: %d (C=%d rows=%lld Cs=%d R=%d)", (int)r, C, (long long)rows, Cs, R);
    return DSG_ERR_CUDA;
  }
  return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_gn_bwd_stream_plan(int32_t n, int64_t hw, int32_t c1, int32_t c2, int32_t groups, int32_t has_addend,
                           int32_t acc1, int32_t acc2, int32_t* parts, int64_t* workspace_bytes) {
  DSG_CHECK_ARG(parts && workspace_bytes, "dsg_gn_bwd_stream_plan: null output pointer");
  const GsPlan p = gs_plan(n, hw, c1, c2, groups, has_addend != 0, acc1 != 0, acc2 != 0 && c2 > 0);
  *parts = p.ok ? p.P : 0;
  *workspace_bytes = p.ok ? p.ws_bytes : 0;
  return DSG_OK;
}

int dsg_gn_bwd_stream(const void* dy, const void* x1, int32_t c1, const void* stats1, const void* x2, int32_t c2,
                      const void* stats2, const float* gamma, const float* beta, float eps, int32_t act, float* red,
                      int64_t red_stride, const void* addend, void* dx1, int32_t acc1, void* dx2, int32_t acc2,
                      float* colsum, float* osum1, float* osum2, int32_t parts, int32_t n, int64_t hw, int32_t groups,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  DSG_CHECK_ARG(dy && x1 && stats1 && dx1 && c1 > 0 && c1 % 8 == 0, "dsg_gn_bwd_stream: dy/x1/stats1/dx1 null or bad c1");
  DSG_CHECK_ARG((x2 == nullptr) == (c2 == 0) && (x2 == nullptr) == (stats2 == nullptr) &&
                    (x2 == nullptr) == (dx2 == nullptr) && c2 % 8 == 0 && c2 >= 0,
                "dsg_gn_bwd_stream: x2/stats2/dx2/c2 mismatch");
  DSG_CHECK_ARG(gamma && beta && red && workspace, "dsg_gn_bwd_stream: null gamma/beta/red/workspace");
  DSG_CHECK_ARG(n >= 0 && hw > 0 && groups > 0, "dsg_gn_bwd_stream: bad n/hw/groups");
  DSG_CHECK_ARG(osum2 == nullptr || x2 != nullptr, "dsg_gn_bwd_stream: osum2 without x2");
  DSG_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)dx1 | (uintptr_t)dx2 | (uintptr_t)addend |
                  (uintptr_t)stats1 | (uintptr_t)stats2 | (uintptr_t)workspace) % 16) == 0 &&
                    (uintptr_t)red % 8 == 0 && red_stride % 2 == 0,
                "dsg_gn_bwd_stream: unaligned pointer");
  if (n == 0) return DSG_OK;
  if (c2 == 0) acc2 = 0;
  const GsPlan p = gs_plan(n, hw, c1, c2, groups, addend != nullptr, acc1 != 0, acc2 != 0);
  DSG_CHECK_ARG(p.ok, "dsg_gn_bwd_stream: shape not supported by the streamed form (ask dsg_gn_bwd_stream_plan first)");
  DSG_CHECK_ARG(workspace_bytes >= p.ws_bytes, "dsg_gn_bwd_stream: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)p.ws_bytes);
  DSG_CHECK_ARG((!colsum && !osum1 && !osum2) || parts == p.P,
                "dsg_gn_bwd_stream: column sums need parts == %d (dsg_gn_bwd_stream_plan)", p.P);
  const int C = c1 + c2;
  const int64_t rows = (int64_t)n * hw;
  GsMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = gs_make_map(&maps.dy, dy, C, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if ((rc = gs_make_map(&maps.x1, x1, c1, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (c2 && (rc = gs_make_map(&maps.x2, x2, c2, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (addend && (rc = gs_make_map(&maps.add, addend, C, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (acc1 && (rc = gs_make_map(&maps.o1, dx1, c1, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (acc2 && (rc = gs_make_map(&maps.o2, dx2, c2, rows, p.Cs, p.R)) != DSG_OK) return rc;
  GsArgs f;
  f.c1 = c1; f.st1 = (const long long*)stats1;
  f.c2 = c2; f.st2 = (const long long*)stats2;
  f.gamma = gamma; f.beta = beta; f.eps = eps; f.act = act;
  f.red = red; f.red_stride = red_stride;
  f.has_add = addend != nullptr;
  f.dx1 = (__half*)dx1; f.acc1 = acc1; f.dx2 = (__half*)dx2; f.acc2 = acc2;
  f.colsum = colsum; f.osum1 = osum1; f.osum2 = osum2;
  f.hw = hw; f.groups = groups; f.n = n;
  f.inv_cnt_s = 1.0 / 16777216.0 / ((double)hw * (double)(C / groups));
  f.inv_cnt_q = 1.0 / 1048576.0 / ((double)hw * (double)(C / groups));
  f.Cs = p.Cs; f.J = p.J; f.D = p.D; f.P = p.P; f.q = p.q; f.R = p.R; f.RB = p.RB;
  f.nst = p.nst; f.stage_bytes = p.stage_bytes; f.tile_bytes = p.tile_bytes;
  f.off_xb = p.off_xb; f.off_dy = p.off_dy; f.off_add = p.off_add; f.off_oa = p.off_oa; f.off_ob = p.off_ob;
  uint8_t* ws = (uint8_t*)workspace;
  f.coef = (float*)ws;
  f.rows = (float*)(ws + p.off_rows);
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(gn_bwd_sums_kernel, dim3((unsigned)p.ctas), dim3(GS_THREADS), p.dyn_smem, st, maps, f);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd_stream/sums");
  launch_k(gn_bwd_coef_kernel, dim3((unsigned)p.D), dim3(GS_CONS), 0, st, f);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd_stream/coef");
  launch_k(gn_bwd_dx_kernel, dim3((unsigned)p.ctas), dim3(GS_THREADS), p.dyn_smem, st, maps, f);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd_stream/dx");
  return DSG_OK;
}
}
