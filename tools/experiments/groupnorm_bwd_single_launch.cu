// groupnorm_bwd_fused.cu — GroupNorm(+SiLU) backward as ONE persistent launch whose second pass reads x / dy from the L2.
// Same contract and formulas as groupnorm_bwd.cu (training path; replaces the autograd backward of
// torch.nn.GroupNorm(32, C) + SiLU of diffusers 0.20.0 models/resnet.py reached from
// DriveSceneGen/pipeline/training_pipeline.py:86 `accelerator.backward(loss)`).
//
// The two-kernel form moves the tensor five times (x, dy for the sums; x, dy again and dx for the result) although the
// arithmetic needs three: the sums of a (sample, group) must be complete before its dx can be formed, and a B = 32
// activation tensor (270 MB at 256 x 256 x 64) does not survive in the 126 MB L2 between two launches.  Here the work is
// cut into DOMAINS = (sample, slice of Cs channels made of whole groups) — independent GroupNorm problems of a few MB —
// and UNITS of S consecutive domains (~16 MB of x + dy); every CTA of the grid owns one pixel range ("part") of one
// domain slot of every unit:
//
//   CTA schedule (lag L = 2):   P1(0) P1(1) P1(2) P2(0) P1(3) P2(1) ... P2(U-1)
//
//   P1 = partial sums of g and g x over the CTA's part -> one row per (domain, part) in the workspace; the LAST CTA to
//        finish a domain (one atomic counter per domain) adds the rows in a fixed order (deterministic), forms the group
//        means, writes the per-channel coefficients of that domain's dx and raises the domain's flag;
//   P2 = dx = ga g + pc x + qc (+ addend, + previous content) over the same part, L units later: the chain row -> fence ->
//        counter -> row sums -> coefficients -> fence -> flag -> coefficient fetch is ~8 dependent trips through a
//        loaded memory system (~1.4 us each, measured), so L units of streaming must cover it; x and dy of the L + 1
//        units in flight (~50 MB) stay in the L2.
//
// Data path: a producer warp streams the parts through a ring of shared-memory stages with 2-D TMA boxes (Cs channels
// x R pixels; a slice that straddles the boundary of the two concatenated sources is two boxes, each clipped by the TMA
// unit's out-of-bounds rule) and mbarriers; eight consumer warps work out of shared memory.  Loads are therefore in
// flight regardless of what the consumers are doing (block reductions, the finaliser), which the load-in-registers form
// of groupnorm_bwd.cu could not do.  The producer is also the one that waits for a domain's flag and brings the domain's
// coefficient rows into shared memory, so consumers never spin.
//
// Requires every CTA of the grid to be co-resident (flags are spun on): the grid is sized from the occupancy query, and
// the spin is bounded (trap) so that a protocol error cannot hang the device.  Sync words live at the head of the
// caller's workspace, must be zero on first use, and are left zero by the last CTA to exit.
#include "igemm_common.cuh"

namespace dsg {

constexpr int GF_CONS = 256;               // consumer threads (8 warps)
constexpr int GF_THREADS = GF_CONS + 32;   // + one producer warp
constexpr int GF_MAX_ST = 8;
constexpr int GF_MAX_CS = 128;             // channels per slice
constexpr int GF_MAX_GROUPS = 64;
constexpr int GF_DYN_BUDGET = 92 * 1024;   // ring + coefficient buffer per CTA (two CTAs per SM)

struct GfMaps { CUtensorMap dy, x1, x2, add, o1, o2; };

struct GnFusedArgs {
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
  int Cs, J, D;          // channels per slice, slices per sample, domains = n * J
  int S, P, U, L;        // domains per unit, parts per domain, units, lag
  int64_t q;             // pixels per part
  int R, RB;             // rows per stage = RB * ppi
  int nst, stage_bytes, tile_bytes;
  int off_xb, off_dy, off_add, off_oa, off_ob;   // byte offsets of the tiles inside a stage
  // workspace
  unsigned* sync;        // [0] exited CTAs, [1 + d] P1 arrivals, [1 + D + s] forward coefficients ready, [1 + D + n + d] dx coefficients ready
  float* fwdrow;         // [n][2 C + 2 groups]: (gah, ybh) per channel (y / 2 = x * gah + ybh), then (mean, rstd) per group
  float* coef2;          // [n][C][4]  ga, ybh, pc, qc
  float* rows;           // [D][P][Cs][2]
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin on a flag another CTA of this grid raises (30 s: a stalled co-tenant of the GPU is legitimate, a
// protocol error must still not hang the box)
__device__ __forceinline__ void spin_flag(const unsigned* p) {
  if (ld_acquire_u32(p) != 0) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_u32(p) == 0) {
    __nanosleep(64);
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 30000000000ull) {
      printf("dsg: gn_bwd_fused flag wait timed out (block %d)\n", (int)blockIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// mbarrier wait with the same 30 s bound (common.cuh's mbar_wait traps after 2 s; a consumer may legitimately sit behind a
// producer that waits for a co-tenant of the GPU)
__device__ __forceinline__ void gf_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 30000000000ull) {
        printf("dsg: gn_bwd_fused mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// d/dy [y sigmoid(y)] for two elements from h = y / 2 — the instruction sequence of groupnorm_bwd.cu::silu_grad_h2
__device__ __forceinline__ __half2 silu_grad_h2f(__half2 h) {
  uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(hi));
  const __half2 t = *reinterpret_cast<__half2*>(&ti);
  const __half2 one = __float2half2_rn(1.0f), half_ = __float2half2_rn(0.5f);
  const __half2 sg = __hfma2(half_, t, half_);
  const __half2 w = __hfma2(h, __hsub2(one, t), one);
  return __hmul2(sg, w);
}

struct GfShared {
  uint64_t full[GF_MAX_ST], empty[GF_MAX_ST];
  uint64_t coef_full, coef_empty;
  float part[2][GF_CONS * 8];     // block reductions; the finaliser's scratch
  float mean[GF_MAX_GROUPS], rstd[GF_MAX_GROUPS], m1[GF_MAX_GROUPS], m2[GF_MAX_GROUPS];
  unsigned long long tot[GF_MAX_GROUPS][2];
  int last;
};

// what one (unit, slot) means: sample, slice, which sources the slice touches
struct GfItem {
  int dom, s, c_lo;
  bool in2, straddle;
};
__device__ __forceinline__ bool gf_item(const GnFusedArgs& f, int u, int slot, GfItem& it) {
  it.dom = u * f.S + slot;
  if (u < 0 || u >= f.U || it.dom >= f.D) return false;
  it.s = it.dom / f.J;
  it.c_lo = (it.dom % f.J) * f.Cs;
  it.in2 = it.c_lo >= f.c1;
  it.straddle = it.c_lo < f.c1 && it.c_lo + f.Cs > f.c1;
  return true;
}

// forward coefficients of one sample from the exact integer totals (the arithmetic of gn_apply_kernel, bit for bit)
__device__ void gf_forward_coef(const GnFusedArgs& f, GfShared& sh, const int s) {
  const int C = f.c1 + f.c2, cpg = C / f.groups, tid = threadIdx.x;
  float* row = f.fwdrow + (int64_t)s * (2 * C + 2 * f.groups);
  if (tid < f.groups) { sh.tot[tid][0] = 0ull; sh.tot[tid][1] = 0ull; }
  cons_sync();
  for (int c = tid; c < C; c += GF_CONS) {
    const longlong2 tv = *reinterpret_cast<const longlong2*>(
        c < f.c1 ? f.st1 + ((int64_t)s * f.c1 + c) * 2 : f.st2 + ((int64_t)s * f.c2 + (c - f.c1)) * 2);
    if (tv.x != 0 || tv.y != 0) {
      atomicAdd(&sh.tot[c / cpg][0], (unsigned long long)tv.x);
      atomicAdd(&sh.tot[c / cpg][1], (unsigned long long)tv.y);
    }
  }
  cons_sync();
  if (tid < f.groups) {
    const double mg = (double)(long long)sh.tot[tid][0] * f.inv_cnt_s;
    double vg = (double)(long long)sh.tot[tid][1] * f.inv_cnt_q - mg * mg;
    if (vg < 0.0) vg = 0.0;
    const float mu = (float)mg, rs = rsqrtf((float)vg + f.eps);
    sh.mean[tid] = mu; sh.rstd[tid] = rs;
    reinterpret_cast<float2*>(row + 2 * C)[tid] = make_float2(mu, rs);
  }
  cons_sync();
  for (int c = tid; c < C; c += GF_CONS) {
    const int g = c / cpg;
    const float ga = f.gamma[c] * sh.rstd[g];
    reinterpret_cast<float2*>(row)[c] = make_float2(0.5f * ga, 0.5f * (f.beta[c] - sh.mean[g] * ga));
  }
  __threadfence();
  cons_sync();
  if (tid == 0) st_release_u32(f.sync + 1 + f.D + s, 1u);
}

// Last CTA of a domain: rows -> per-channel sums (fixed order) -> group means -> dx coefficients; raises the flag.
// sh.mean / sh.rstd hold the slice's groups (copied from the coefficient row at the start of this CTA's own P1 item of the
// same domain).
__device__ void gf_finalize(const GnFusedArgs& f, GfShared& sh, const GfItem& it) {
  const int C = f.c1 + f.c2, cpg = C / f.groups, tid = threadIdx.x, Cs = f.Cs, ng = Cs / cpg;
  // per-channel constants first: their loads fly under the row loads
  float gm = 0.f, yb = 0.f;
  if (tid < Cs) {
    gm = f.gamma[it.c_lo + tid];
    yb = __ldcg(f.fwdrow + (int64_t)it.s * (2 * C + 2 * f.groups) + 2 * (it.c_lo + tid) + 1);
  }
  __threadfence();
  float4* fin = reinterpret_cast<float4*>(&sh.part[0][0]);           // [kg][Q] <= 4 KB
  float2* chs = reinterpret_cast<float2*>(&sh.part[0][0]) + 512;     // [Cs] (sum g, sum g x) then (gamma A, gamma B)
  const int Q = Cs >> 1;                                             // float4 per row (<= 64)
  const float4* rows = reinterpret_cast<const float4*>(f.rows + (int64_t)it.dom * f.P * Cs * 2);
  const int kg = GF_CONS / Q;
  const int col = tid % Q, kgi = tid / Q;
  if (kgi < kg) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = kgi;
    for (; k + 15 * kg < f.P; k += 16 * kg) {   // sixteen independent loads in flight, added in index order
      float4 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldcg(rows + (int64_t)(k + u * kg) * Q + col);
#pragma unroll
      for (int u = 0; u < 16; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    {
      float4 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u)
        v[u] = (k + u * kg < f.P) ? __ldcg(rows + (int64_t)(k + u * kg) * Q + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 16; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    fin[kgi * Q + col] = acc;
  }
  cons_sync();
  if (tid < Q) {
    float4 acc = fin[tid];
    for (int j = 1; j < kg; ++j) {   // fixed order
      const float4 v = fin[j * Q + tid];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    chs[2 * tid] = make_float2(acc.x, acc.y);
    chs[2 * tid + 1] = make_float2(acc.z, acc.w);
  }
  cons_sync();
  if (tid < Cs) {
    const float2 t = chs[tid];
    const int g = tid / cpg;
    const float tB = sh.rstd[g] * (t.y - sh.mean[g] * t.x);   // sum g * xh
    reinterpret_cast<float2*>(f.red + (int64_t)it.s * f.red_stride)[it.c_lo + tid] = make_float2(t.x, tB);
    chs[tid] = make_float2(gm * t.x, gm * tB);
  }
  cons_sync();
  if (tid < ng) {
    float m1 = 0.f, m2 = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { m1 += chs[c].x; m2 += chs[c].y; }
    const float inv_cnt = (float)(1.0 / ((double)f.hw * (double)cpg));
    sh.m1[tid] = m1 * inv_cnt;
    sh.m2[tid] = m2 * inv_cnt;
  }
  cons_sync();
  if (tid < Cs) {
    const int g = tid / cpg;
    const float mu = sh.mean[g], rs = sh.rstd[g];
    reinterpret_cast<float4*>(f.coef2)[(int64_t)it.s * C + it.c_lo + tid] =
        make_float4(gm * rs, yb, -rs * rs * sh.m2[g], rs * (mu * rs * sh.m2[g] - sh.m1[g]));
    __threadfence();
  }
  cons_sync();
  if (tid == 0) st_release_u32(f.sync + 1 + f.D + f.n + it.dom, 1u);
}

__global__ void __launch_bounds__(GF_THREADS, 2) gn_bwd_fused_kernel(const __grid_constant__ GfMaps maps,
                                                                      const GnFusedArgs f) {
  extern __shared__ __align__(128) uint8_t gf_dyn[];
  __shared__ GfShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = f.c1 + f.c2, Cs = f.Cs, V = Cs >> 3, ppi = GF_CONS / V, cpg = C / f.groups, ng = Cs / cpg;
  uint8_t* ring = gf_dyn;
  uint8_t* coefbuf = gf_dyn + (size_t)f.nst * f.stage_bytes;
  const int slot = blockIdx.x / f.P, part = blockIdx.x % f.P;
  const int64_t p0 = (int64_t)part * f.q;
  int64_t p1 = p0 + f.q;
  if (p1 > f.hw) p1 = f.hw;

  if (tid == 0) {
    for (int i = 0; i < f.nst; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], GF_CONS / 32); }
    mbar_init(&sh.coef_full, 1);
    mbar_init(&sh.coef_empty, GF_CONS / 32);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_sync();

  if (warp == GF_CONS / 32) {
    // ------------------------------------------------------------ producer: flags, coefficient rows, the data ring
    if (lane == 0) {
      tma_prefetch_desc(&maps.dy); tma_prefetch_desc(&maps.x1);
      if (f.c2) tma_prefetch_desc(&maps.x2);
    }
    int st = 0;
    uint32_t ph = 0, cph = 0;
    for (int w = 0; w < f.U + f.L; ++w) {
      for (int pass = 0; pass < 2; ++pass) {
        GfItem it;
        if (!gf_item(f, pass == 0 ? w : w - f.L, slot, it)) continue;
        const bool p2 = pass == 1;
        if (lane == 0) {
          spin_flag(f.sync + 1 + f.D + (p2 ? f.n + it.dom : it.s));
          asm volatile("fence.proxy.async.global;" ::: "memory");   // rows written through the generic proxy, read by the copy engine
          gf_wait(&sh.coef_empty, cph ^ 1);
          if (p2) {
            mbar_arrive_expect_tx(&sh.coef_full, (uint32_t)Cs * 16u);
            bulk_g2s(coefbuf, f.coef2 + ((int64_t)it.s * C + it.c_lo) * 4, (uint32_t)Cs * 16u, &sh.coef_full);
          } else {
            const float* row = f.fwdrow + (int64_t)it.s * (2 * C + 2 * f.groups);
            mbar_arrive_expect_tx(&sh.coef_full, (uint32_t)Cs * 8u + (uint32_t)ng * 8u);
            bulk_g2s(coefbuf, row + 2 * it.c_lo, (uint32_t)Cs * 8u, &sh.coef_full);
            bulk_g2s(coefbuf + Cs * 8, row + 2 * C + 2 * (it.c_lo / cpg), (uint32_t)ng * 8u, &sh.coef_full);
          }
        }
        cph ^= 1;
        __syncwarp();
        const int64_t base = (int64_t)it.s * f.hw;
        const bool old_a = p2 && (it.in2 ? f.acc2 : f.acc1), old_b = p2 && it.straddle && f.acc2;
        const uint32_t bytes = (uint32_t)f.tile_bytes * (2u + (it.straddle ? 1u : 0u) + ((p2 && f.has_add) ? 1u : 0u) +
                                                         (old_a ? 1u : 0u) + (old_b ? 1u : 0u));
        for (int64_t r0 = p0; r0 < p1; r0 += f.R) {
          if (lane == 0) {
            uint8_t* sb = ring + (size_t)st * f.stage_bytes;
            const int row = (int)(base + r0);
            gf_wait(&sh.empty[st], ph ^ 1);
            mbar_arrive_expect_tx(&sh.full[st], bytes);
            if (it.in2) tma_load_2d(sb, &maps.x2, &sh.full[st], it.c_lo - f.c1, row);
            else tma_load_2d(sb, &maps.x1, &sh.full[st], it.c_lo, row);
            if (it.straddle) tma_load_2d(sb + f.off_xb, &maps.x2, &sh.full[st], it.c_lo - f.c1, row);
            tma_load_2d(sb + f.off_dy, &maps.dy, &sh.full[st], it.c_lo, row);
            if (p2) {
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
    return;
  }

  // -------------------------------------------------------------- consumers
  // forward coefficients (mean / rstd and the SiLU argument's affine map) of the samples this CTA is responsible for
  for (int s = blockIdx.x; s < f.n; s += gridDim.x) gf_forward_coef(f, sh, s);

  const bool active = tid < ppi * V;
  const int v = active ? tid % V : 0, prow = active ? tid / V : 0;
  int st = 0;
  uint32_t ph = 0, cph = 0;

  for (int w = 0; w < f.U + f.L; ++w) {
    for (int pass = 0; pass < 2; ++pass) {
      GfItem it;
      if (!gf_item(f, pass == 0 ? w : w - f.L, slot, it)) continue;
      const int64_t base = (int64_t)it.s * f.hw;
      const int ch = it.c_lo + (v << 3);          // first of this thread's 8 channels
      const bool from1 = ch < f.c1;
      const int xoff = (it.straddle && !from1) ? f.off_xb : 0;
      const int col = (v << 3) * 2;               // byte offset of the thread's vector inside a tile row
      if (pass == 0) {
        // ---------------------------------------------------------- P1: sums of g and g x over this part
        float gah[8], ybh[8], sA[8], sB[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sA[j] = 0.f; sB[j] = 0.f; }
        gf_wait(&sh.coef_full, cph);
        {
          const float4* cb = reinterpret_cast<const float4*>(coefbuf) + (v << 2);   // (gah, ybh) pairs
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 t = cb[j];
            gah[2 * j] = t.x; ybh[2 * j] = t.y; gah[2 * j + 1] = t.z; ybh[2 * j + 1] = t.w;
          }
          if (tid < ng) {   // this slice's (mean, rstd): what a finaliser in this CTA will need
            const float2 t = reinterpret_cast<const float2*>(coefbuf + Cs * 8)[tid];
            sh.mean[tid] = t.x; sh.rstd[tid] = t.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.coef_empty);
        cph ^= 1;
        for (int64_t r0 = p0; r0 < p1; r0 += f.R) {
          const int nr = (int)((p1 - r0 < f.R) ? (p1 - r0) : f.R);
          const uint8_t* sb = ring + (size_t)st * f.stage_bytes;
          gf_wait(&sh.full[st], ph);
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
                    g2 = __hmul2(g2, silu_grad_h2f(h2));
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
        cons_sync();   // the previous user of sh.part (a finaliser or column sums) is done
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
          reinterpret_cast<float2*>(f.rows + ((int64_t)it.dom * f.P + part) * Cs * 2)[tid] = make_float2(tA, tB);
          __threadfence();
        }
        cons_sync();
        if (tid == 0) sh.last = (atomicAdd(f.sync + 1 + it.dom, 1u) == (unsigned)(f.P - 1));
        cons_sync();
        if (sh.last) gf_finalize(f, sh, it);
      } else {
        // ---------------------------------------------------------- P2: dx over the same part, out of the L2
        const int ooff = (it.straddle && !from1) ? f.off_ob : f.off_oa;
        const int accum = from1 ? f.acc1 : f.acc2;
        __half* dst = from1 ? f.dx1 + ch : f.dx2 + (ch - f.c1);
        const int cs = from1 ? f.c1 : f.c2;
        const bool want_osum = (from1 ? f.osum1 : f.osum2) != nullptr;
        float ga[8], ybh[8], pc[8], qc[8], cs_acc[8], os_acc[8];
        gf_wait(&sh.coef_full, cph);
        {
          const float4* cb = reinterpret_cast<const float4*>(coefbuf) + (v << 3);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = cb[j];
            ga[j] = t.x; ybh[j] = t.y; pc[j] = t.z; qc[j] = t.w;
            cs_acc[j] = 0.f; os_acc[j] = 0.f;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.coef_empty);
        cph ^= 1;
        for (int64_t r0 = p0; r0 < p1; r0 += f.R) {
          const int nr = (int)((p1 - r0 < f.R) ? (p1 - r0) : f.R);
          const uint8_t* sb = ring + (size_t)st * f.stage_bytes;
          gf_wait(&sh.full[st], ph);
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
                if (f.act) {
                  __half2* hd = reinterpret_cast<__half2*>(&rd);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    const __half2 h2 = __floats2half2_rn(fmaf(fx[2 * kk], 0.5f * ga[2 * kk], ybh[2 * kk]),
                                                         fmaf(fx[2 * kk + 1], 0.5f * ga[2 * kk + 1], ybh[2 * kk + 1]));
                    hd[kk] = __hmul2(hd[kk], silu_grad_h2f(h2));
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
                if (want_osum) {
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
            f.colsum[((int64_t)it.s * f.P + part) * C + it.c_lo + tid] = t;
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
              ob[((int64_t)it.s * f.P + part) * cw + cc] = t;
            }
          }
        }
      }
    }
  }
  // the last CTA to leave clears the sync words for the next launch on this workspace
  cons_sync();
  if (tid == 0) {
    __threadfence();
    sh.last = (atomicAdd(f.sync, 1u) == gridDim.x - 1);
  }
  cons_sync();
  if (sh.last) {
    __threadfence();
    for (int i = tid; i < 1 + 2 * f.D + f.n; i += GF_CONS) f.sync[i] = 0u;
  }
}

// ------------------------------------------------------------------ host side: the plan
struct GfPlan {
  bool ok;
  int Cs, J, D, S, P, U, L, R, RB, nst, stage_bytes, tile_bytes, ctas;
  int64_t q;
  int off_xb, off_dy, off_add, off_oa, off_ob;
  size_t dyn_smem;
  int64_t sync_bytes, ws_bytes, off_fwd, off_coef2, off_rows;
};

static int gf_max_ctas() {
  static std::atomic<int> cached[kMaxDevices];
  const int d = current_device();
  int v = cached[d].load(std::memory_order_relaxed);
  if (v != 0) return v;
  cudaFuncSetAttribute(gn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_DYN_BUDGET);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_bwd_fused_kernel, GF_THREADS, GF_DYN_BUDGET) != cudaSuccess)
    per_sm = 0;
  (void)cudaGetLastError();
  if (per_sm > 2) per_sm = 2;
  v = per_sm * num_sms();
  const char* e = getenv("DSG_GN_BWD_CTAS");   // e.g. to leave SMs to a concurrent collective
  if (e && atoi(e) > 0 && atoi(e) < v) v = atoi(e);
  if (v <= 0) v = -1;
  cached[d].store(v, std::memory_order_relaxed);
  return v;
}

static int gf_gcd(int a, int b) { return b ? gf_gcd(b, a % b) : a; }

static GfPlan gf_plan(int n, int64_t hw, int c1, int c2, int groups, bool has_add, bool has_acc1, bool has_acc2) {
  GfPlan p = {};
  const int C = c1 + c2;
  static const int enabled = getenv("DSG_GN_BWD_FUSED") ? atoi(getenv("DSG_GN_BWD_FUSED")) : 1;
  static const double target = (getenv("DSG_GN_BWD_UNIT_MB") ? atof(getenv("DSG_GN_BWD_UNIT_MB")) : 17.0) * 1048576.0;
  static const int lag = getenv("DSG_GN_BWD_LAG") ? atoi(getenv("DSG_GN_BWD_LAG")) : 2;
  if (!enabled) return p;
  if (c1 <= 0 || c1 % 8 || c2 < 0 || c2 % 8 || groups <= 0 || groups > GF_MAX_GROUPS || C % groups || n <= 0 || hw <= 0)
    return p;
  if ((double)n * (double)hw > 2.0e9) return p;   // TMA row coordinates are int32
  const int G = gf_max_ctas();
  if (G < 2) return p;
  // slice width: whole groups, whole 16-byte vectors, an even number of groups (16-byte rows of (mean, rstd)), <= 128
  // channels; the widest one whose domain fits the unit target, at least 32 channels (64-byte TMA rows)
  const int cpg = C / groups;
  const int step = cpg / gf_gcd(cpg, 8) * 8;   // lcm(cpg, 8)
  int Cs = 0, smallest = 0;
  for (int w = step; w <= GF_MAX_CS && w <= C; w += step) {
    if (C % w || (w / cpg) % 2 || w < 32) continue;
    if (!smallest) smallest = w;
    if ((double)hw * w * 4.0 <= target) Cs = w;
  }
  if (!Cs) Cs = smallest;
  if (!Cs) return p;
  p.Cs = Cs; p.J = C / Cs; p.D = n * p.J;
  const int V = Cs / 8, ppi = GF_CONS / V;
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
  const int coef_bytes = Cs * 16 + 256;
  p.nst = (GF_DYN_BUDGET - coef_bytes) / p.stage_bytes;
  if (p.nst > GF_MAX_ST) p.nst = GF_MAX_ST;
  if (p.nst < 2) return p;
  p.dyn_smem = (size_t)p.nst * p.stage_bytes + coef_bytes;
  const double dom_bytes = (double)hw * Cs * 4.0;
  int S = (int)(target / dom_bytes);
  if (S < 1) S = 1;
  if (S > p.D) S = p.D;
  if (S > G) S = G;
  for (int d = S; d >= 1 && d * 4 > S * 3; --d)   // a divisor of the domain count when one is close: no ragged last unit
    if (p.D % d == 0) { S = d; break; }
  p.S = S;
  p.U = (p.D + S - 1) / S;
  p.L = lag < 1 ? 1 : lag;
  if (p.L > p.U) p.L = p.U;
  if (p.U > 65535) return p;
  const int Pmax = G / S;
  int64_t q = (hw + Pmax - 1) / Pmax;
  q = (q + p.R - 1) / p.R * p.R;               // whole stages: no box reads pixels of the next part
  p.q = q;
  p.P = (int)((hw + q - 1) / q);
  p.ctas = p.S * p.P;
  p.sync_bytes = ((int64_t)(1 + 2 * p.D + n) * 4 + 255) & ~(int64_t)255;
  int64_t o = p.sync_bytes;
  p.off_fwd = o; o += (int64_t)n * (2 * C + 2 * groups) * 4;
  o = (o + 255) & ~(int64_t)255;
  p.off_coef2 = o; o += (int64_t)n * C * 4 * 4;
  p.off_rows = o; o += (int64_t)p.D * p.P * Cs * 2 * 4;
  p.ws_bytes = o;
  p.ok = true;
  return p;
}

// [rows][C] fp16 tensor, box = Cs channels x R rows, dense in shared memory
static int gf_make_map(CUtensorMap* m, const void* ptr, int C, int64_t rows, int Cs, int R) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DSG_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)Cs, (cuuint32_t)R};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(gn_bwd_fused) failed: %d (C=%d rows=%lld Cs=%d R=%d)", (int)r, C, (long long)rows, Cs, R);
    return DSG_ERR_CUDA;
  }
  return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int dsg_gn_bwd_fused_plan(int32_t n, int64_t hw, int32_t c1, int32_t c2, int32_t groups, int32_t has_addend,
                          int32_t acc1, int32_t acc2, int32_t* parts, int64_t* workspace_bytes, int64_t* sync_bytes) {
  DSG_CHECK_ARG(parts && workspace_bytes && sync_bytes, "dsg_gn_bwd_fused_plan: null output pointer");
  const GfPlan p = gf_plan(n, hw, c1, c2, groups, has_addend != 0, acc1 != 0, acc2 != 0 && c2 > 0);
  *parts = p.ok ? p.P : 0;
  *workspace_bytes = p.ok ? p.ws_bytes : 0;
  *sync_bytes = p.ok ? p.sync_bytes : 0;
  return DSG_OK;
}

int dsg_gn_bwd_fused(const void* dy, const void* x1, int32_t c1, const void* stats1, const void* x2, int32_t c2,
                     const void* stats2, const float* gamma, const float* beta, float eps, int32_t act, float* red,
                     int64_t red_stride, const void* addend, void* dx1, int32_t acc1, void* dx2, int32_t acc2,
                     float* colsum, float* osum1, float* osum2, int32_t parts, int32_t n, int64_t hw, int32_t groups,
                     void* workspace, int64_t workspace_bytes, void* stream) {
  DSG_CHECK_ARG(dy && x1 && stats1 && dx1 && c1 > 0 && c1 % 8 == 0, "dsg_gn_bwd_fused: dy/x1/stats1/dx1 null or bad c1");
  DSG_CHECK_ARG((x2 == nullptr) == (c2 == 0) && (x2 == nullptr) == (stats2 == nullptr) &&
                    (x2 == nullptr) == (dx2 == nullptr) && c2 % 8 == 0 && c2 >= 0,
                "dsg_gn_bwd_fused: x2/stats2/dx2/c2 mismatch");
  DSG_CHECK_ARG(gamma && beta && red && workspace, "dsg_gn_bwd_fused: null gamma/beta/red/workspace");
  DSG_CHECK_ARG(n >= 0 && hw > 0 && groups > 0, "dsg_gn_bwd_fused: bad n/hw/groups");
  DSG_CHECK_ARG(osum2 == nullptr || x2 != nullptr, "dsg_gn_bwd_fused: osum2 without x2");
  DSG_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)dx1 | (uintptr_t)dx2 | (uintptr_t)addend |
                  (uintptr_t)stats1 | (uintptr_t)stats2 | (uintptr_t)workspace) % 16) == 0 &&
                    (uintptr_t)red % 8 == 0 && red_stride % 2 == 0,
                "dsg_gn_bwd_fused: unaligned pointer");
  if (n == 0) return DSG_OK;
  if (c2 == 0) acc2 = 0;
  const GfPlan p = gf_plan(n, hw, c1, c2, groups, addend != nullptr, acc1 != 0, acc2 != 0);
  DSG_CHECK_ARG(p.ok, "dsg_gn_bwd_fused: shape not supported by the fused form (ask dsg_gn_bwd_fused_plan first)");
  DSG_CHECK_ARG(workspace_bytes >= p.ws_bytes, "dsg_gn_bwd_fused: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)p.ws_bytes);
  DSG_CHECK_ARG((!colsum && !osum1 && !osum2) || parts == p.P,
                "dsg_gn_bwd_fused: column sums need parts == %d (dsg_gn_bwd_fused_plan)", p.P);
  const int C = c1 + c2;
  const int64_t rows = (int64_t)n * hw;
  GfMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = gf_make_map(&maps.dy, dy, C, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if ((rc = gf_make_map(&maps.x1, x1, c1, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (c2 && (rc = gf_make_map(&maps.x2, x2, c2, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (addend && (rc = gf_make_map(&maps.add, addend, C, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (acc1 && (rc = gf_make_map(&maps.o1, dx1, c1, rows, p.Cs, p.R)) != DSG_OK) return rc;
  if (acc2 && (rc = gf_make_map(&maps.o2, dx2, c2, rows, p.Cs, p.R)) != DSG_OK) return rc;
  GnFusedArgs f;
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
  f.Cs = p.Cs; f.J = p.J; f.D = p.D; f.S = p.S; f.P = p.P; f.U = p.U; f.L = p.L; f.q = p.q; f.R = p.R; f.RB = p.RB;
  f.nst = p.nst; f.stage_bytes = p.stage_bytes; f.tile_bytes = p.tile_bytes;
  f.off_xb = p.off_xb; f.off_dy = p.off_dy; f.off_add = p.off_add; f.off_oa = p.off_oa; f.off_ob = p.off_ob;
  uint8_t* ws = (uint8_t*)workspace;
  f.sync = (unsigned*)ws;
  f.fwdrow = (float*)(ws + p.off_fwd);
  f.coef2 = (float*)(ws + p.off_coef2);
  f.rows = (float*)(ws + p.off_rows);
  launch_k(gn_bwd_fused_kernel, dim3((unsigned)p.ctas), dim3(GF_THREADS), p.dyn_smem, (cudaStream_t)stream, maps, f);
  DSG_CUDA_LAUNCH_CHECK("dsg_gn_bwd_fused");
  return DSG_OK;
}
}
