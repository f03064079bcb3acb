// wgrad.cu — weight gradients of the U-Net convolutions / linears on tcgen05 tensor cores (training path).
//
// Replaces the cuDNN wgrad kernels autograd reaches from `accelerator.backward(loss)`
// (DriveSceneGen/pipeline/training_pipeline.py:86) for ResnetBlock2D.conv1/conv2/conv_shortcut, Downsample2D.conv,
// Upsample2D.conv and the attention Linear layers of diffusers 0.20.0 (SURVEY.md §2.2 "autograd backward", §8 a17):
//   dW[co][tap][ci] = sum over (n, y, x) of dY[n][y][x][co] * X[n][y + dy(tap)][x + dx(tap)][ci]
// This is a GEMM whose K dimension is the PIXEL axis, so both operands are "MN-major" in NHWC memory: a TMA box of
// 128 pixels x 64 channels (128B-swizzled rows of 64 fp16) is directly a K = 128 slab of an MN-major UMMA operand
// (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: SBO = 1024 B between 8-pixel groups, LBO = the
// distance between 64-channel atoms).  Two tricks make the 3x3 case cheap:
//   * halo reuse — one X box of (TH + 2) x TW pixels serves the three dy taps of a column shift dx: tap dy is the same
//     box at byte offset (dy + 1) * TW * 128 (a multiple of the 1024 B swizzle atom), and
//   * taps as N atoms — those three shifted views are the three 64-wide "atoms" of ONE N = 192 B operand
//     (LBO = TW * 128: the atoms overlap in shared memory), so a single tcgen05.mma covers three taps.
// A = dY (M = 128 output channels, two 64-channel boxes; a 64-channel layer instead stacks dY shifted by one pixel in x
// as its second M atom — a shift of dY by s against the unshifted X is the tap dx = -s — so two of its three column
// taps share one full-height MMA and the level-0 layers do 2/3 of the tensor work they did with don't-care rows),
// D[co][(tap, ci)] accumulates in TMEM over a contiguous range of pixel tiles; the pixel axis is split over CTAs
// (one work item x one split per CTA, one wave) and each CTA stores its fp32 partial tile to a workspace
// [split][co][tap][ci].  wgrad_reduce_kernel sums the splits in a fixed order (deterministic — no atomics), applies
// the loss-scale inverse, folds the 16 sub-pixel taps of the upsample conv back onto its 3x3 weight, and writes the
// fp32 OIHW gradient the optimizer reads.
// Warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-5 = epilogue (TMEM -> workspace).
#include "igemm_common.cuh"

namespace dsg {

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_BOX = 4;
constexpr int WG_MAX_UNIT = 4;
constexpr int WG_MAX_GRP = 3;
constexpr int WG_MAX_MAPS = 4;
constexpr int WG_A_BOX = 16384;          // 128 pixels x 64 channels fp16
constexpr int WG_RING_BYTES = 216 * 1024;

struct WgBox {
  int map;     // X tensor map
  int chunk;   // 64-channel chunk relative to the item's first
  int dx, dy0; // box origin relative to the pixel tile origin
  int bytes;
  int smem_off;  // within the stage's B region
};
struct WgUnit {
  int smem_off;  // B operand start within the stage's B region (box offset + row offset)
  int natoms;    // N = natoms * 64
  int lbo;       // bytes between atoms
  int tmem_col;
  int tap0, tap_step;  // workspace tap of atom a = tap0 + a * tap_step
  int ci0, ci_step;    // workspace channel (relative to the item's first) of atom a = ci0 + a * ci_step
};
struct WgGroup {
  WgBox box[WG_MAX_BOX];
  WgUnit unit[WG_MAX_UNIT];
  int nbox, nunit;
  // A operand (dY): a_nbox 64-channel boxes = M atoms.  Normally atoms are consecutive channel blocks of one output-channel
  // block (a_chan = 0, 1; a_dx = 0).  A 64-channel layer has only one: instead of 64 don't-care accumulator rows, its
  // second atom is the SAME channels shifted by one pixel in x — shifting dY by s against an unshifted X is the tap
  // dx = -s, so one MMA covers two column taps (tap_hi = the workspace tap column of accumulator rows 64..127, else -1).
  int a_nbox;
  int a_chan[2], a_dx[2];
  int tap_hi;
};
struct WgPlan {
  WgGroup grp[WG_MAX_GRP];
  int ngroups;
  int N, H, W;            // pixel grid of dY (per phase)
  int TW, tw_shift, TH;
  int tiles_w, tiles_h;
  int phases;             // 1, or 4 for the upsample conv (dY parity views, X offsets + (pa, pb))
  int co, co_blocks, n_abox;
  int ci, ci_per_item, ci_groups;
  int ktaps;              // taps in the workspace layout (9, 16 or 1)
  int64_t tiles_total;
  int splits;
  int64_t tiles_per_split;
  float* ws;              // [splits][co][ktaps][ci]
  int stage_bytes, nstages, a_bytes;
};
struct alignas(64) WgMaps {
  CUtensorMap a[4];            // dY (one per phase)
  CUtensorMap x[WG_MAX_MAPS];  // X sources
};

// MN-major, 128-byte-swizzled operand descriptor
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset: between 64-element MN atoms
  d |= (uint64_t)64 << 32;                           // stride byte offset: 1024 B between 8-row K groups
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
// kind::f16, fp16 A/B both MN-major, fp32 D, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_f16_mn(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgPlan p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_RING_BYTES);
  uint64_t* full = bars;         // [nstages]
  uint64_t* empty = bars + 8;    // [nstages]
  uint64_t* tfull = bars + 16;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  // work item -> (phase, output-channel block, input-channel group, tap group)
  int item = blockIdx.y;
  const int g = item % p.ngroups; item /= p.ngroups;
  const int cig = item % p.ci_groups; item /= p.ci_groups;
  const int cob = item % p.co_blocks;
  const int phase = item / p.co_blocks;
  const int pa = phase >> 1, pb = phase & 1;
  const WgGroup& G = p.grp[g];
  const int split = blockIdx.x;
  const int64_t t_begin = (int64_t)split * p.tiles_per_split;
  int64_t t_end = t_begin + p.tiles_per_split;
  if (t_end > p.tiles_total) t_end = p.tiles_total;
  const int chunk0 = cig * (p.ci_per_item / 64);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[phase]);
    for (int i = 0; i < WG_MAX_MAPS; ++i) tma_prefetch_desc(&maps.x[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  pdl_sync();  // setup above overlaps the previous kernel's tail

  int stage_bytes_tx = G.a_nbox * WG_A_BOX;
  for (int b = 0; b < G.nbox; ++b) stage_bytes_tx += G.box[b].bytes;

  if (warp == 0) {
    // ===================================================== TMA producer
    int stage = 0; uint32_t ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      int64_t r = t;
      const int tw = (int)(r % p.tiles_w); r /= p.tiles_w;
      const int th = (int)(r % p.tiles_h);
      const int n = (int)(r / p.tiles_h);
      const int h0 = th * p.TH, w0 = tw * p.TW;
      mbar_wait(&empty[stage], ph ^ 1);
      uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes_tx);
        for (int i = 0; i < G.a_nbox; ++i)
          tma_load_4d(sa + i * WG_A_BOX, &maps.a[phase], &full[stage], (cob * 2 + G.a_chan[i]) * 64, w0 + G.a_dx[i], h0,
                      n);
        for (int b = 0; b < G.nbox; ++b) {
          const WgBox& B = G.box[b];
          tma_load_4d(sa + p.a_bytes + B.smem_off, &maps.x[B.map], &full[stage], (chunk0 + B.chunk) * 64,
                      w0 + B.dx + pb, h0 + B.dy0 + pa, n);
        }
      }
      __syncwarp();
      if (++stage == p.nstages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    int stage = 0; uint32_t ph = 0;
    uint32_t accum = 0;
    // a 64-channel layer has one A box: the second M atom then reads the bytes that follow it in shared memory (the
    // stage's X boxes) — finite fp16 values whose products land in accumulator rows 64..127, which nobody reads
    const uint32_t a_lbo = WG_A_BOX;
    for (int64_t t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[stage], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
      const uint64_t da = umma_desc_mn_sw128(sa, a_lbo);
      if (elect_one_sync()) {
        for (int u = 0; u < G.nunit; ++u) {
          const WgUnit& U = G.unit[u];
          const uint64_t db = umma_desc_mn_sw128(sa + (uint32_t)(p.a_bytes + U.smem_off), (uint32_t)U.lbo);
          const uint32_t idesc = umma_idesc_f16_mn(U.natoms * 64);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 pixels per MMA: 2048 B further into both operands
            umma_f16(tmem_base + (uint32_t)U.tmem_col, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc,
                     k == 0 ? accum : 1u);
        }
        umma_commit(&empty[stage]);
        if (t + 1 == t_end) umma_commit(tfull);
      }
      __syncwarp();
      accum = 1;
      if (++stage == p.nstages) { stage = 0; ph ^= 1; }
    }
  } else {
    // ===================================================== epilogue: TMEM -> fp32 workspace partial
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool hi = G.tap_hi >= 0 && row >= 64;          // rows 64..127 = the second column tap of a 64-channel layer
    const int co = G.tap_hi >= 0 ? (row & 63) : cob * 128 + row;
    if (t_begin < t_end) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* wrow = p.ws + ((int64_t)split * p.co + co) * p.ktaps * p.ci;
    for (int u = 0; u < G.nunit; ++u) {
      const WgUnit& U = G.unit[u];
      for (int a = 0; a < U.natoms; ++a) {
        const int tap = (hi ? G.tap_hi : U.tap0) + a * U.tap_step + phase * 4;
        const int ci = chunk0 * 64 + U.ci0 + a * U.ci_step;
        float* dst = wrow + (int64_t)tap * p.ci + ci;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          if (t_begin < t_end) {
            tmem_ld_32x32(taddr + (uint32_t)(U.tmem_col + a * 64 + h * 32), v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;  // an empty split contributes zeros
          }
          if (row < 128 && co < p.co) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + h * 32 + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ split reduction + layout change
// fold table of the upsample conv: 3x3 tap (ky, kx) collects sub-pixel taps (a, b, i(a, ky), j(b, kx))
__device__ __forceinline__ int up_tap_i(int a, int ky) { return a == 0 ? (ky == 0 ? 0 : 1) : (ky == 2 ? 1 : 0); }

// ws [splits][co][ktaps][ci] -> out fp32 [co_count][ci_total][KK] (OIHW), rows co_begin.., columns ci_off..ci_off+ci.
// One block per (output channel, ci_b input channels): phase 1 walks the KK x ci_b slab coalesced along ci, several
// elements per thread with the loads of all splits in flight, sums the splits in a fixed order into a shared tile;
// phase 2 writes the tile transposed, i.e. ci_b * KK consecutive floats of the OIHW gradient.
constexpr int WR_THREADS = 256, WR_CI_MAX = 256, WR_PITCH = WR_CI_MAX + 1;
// EPT = elements per thread (KK * ci_b / 256 rounded up), SPU = splits in flight per element: EPT * SPU independent loads
template <int EPT, int SPU>
__global__ void __launch_bounds__(WR_THREADS) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int co,
                                                                  int ktaps, int ci, int ci_b, int fold_up,
                                                                  int co_begin, int co_count, int ci_total, int ci_off,
                                                                  const float* __restrict__ inv_scale,
                                                                  float* __restrict__ out, int accumulate) {
  const int KK = fold_up ? 9 : ktaps;
  __shared__ float tile[9 * WR_PITCH];
  const int o = blockIdx.y, c0 = blockIdx.x * ci_b;
  const int nc = min(ci_b, ci - c0);
  pdl_sync();
  const float s = inv_scale ? *inv_scale : 1.0f;
  const int64_t split_stride = (int64_t)co * ktaps * ci;
  const float* base0 = ws + ((int64_t)(co_begin + o) * ktaps) * ci + c0;
  if (fold_up) {
    for (int e = threadIdx.x; e < KK * nc; e += WR_THREADS) {
      const int kk = e / nc, c = e - kk * nc;
      const float* base = base0 + c;
      float acc = 0.f;
      const int ky = kk / 3, kx = kk % 3;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int kt = (a * 2 + b) * 4 + up_tap_i(a, ky) * 2 + up_tap_i(b, kx);
          for (int sp = 0; sp < splits; ++sp) acc += base[(int64_t)sp * split_stride + (int64_t)kt * ci];
        }
      tile[kk * WR_PITCH + c] = acc * s;
    }
  } else {
    // the loads of SPU splits x EPT elements are issued before the first add (an element-at-a-time loop left one round
    // trip to DRAM per element, or per split, on the critical path); the sums keep the split order
    const int total = KK * nc;
    int offs[EPT], tidx[EPT];
    float acc[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int e = threadIdx.x + i * WR_THREADS;
      const int ee = e < total ? e : 0;
      const int kk = ee / nc, c = ee - kk * nc;
      offs[i] = kk * ci + c;
      tidx[i] = kk * WR_PITCH + c;
      acc[i] = 0.f;
    }
    const float* bp = base0;
    int sp = 0;
    for (; sp + SPU <= splits; sp += SPU, bp += (int64_t)SPU * split_stride) {
      float v[SPU][EPT];
#pragma unroll
      for (int u = 0; u < SPU; ++u)
#pragma unroll
        for (int i = 0; i < EPT; ++i)
          v[u][i] = threadIdx.x + i * WR_THREADS < total ? __ldg(bp + (int64_t)u * split_stride + offs[i]) : 0.f;
#pragma unroll
      for (int u = 0; u < SPU; ++u)
#pragma unroll
        for (int i = 0; i < EPT; ++i) acc[i] += v[u][i];
    }
    for (; sp < splits; ++sp, bp += split_stride) {
      float v[EPT];
#pragma unroll
      for (int i = 0; i < EPT; ++i) v[i] = threadIdx.x + i * WR_THREADS < total ? __ldg(bp + offs[i]) : 0.f;
#pragma unroll
      for (int i = 0; i < EPT; ++i) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (threadIdx.x + i * WR_THREADS < total) tile[tidx[i]] = acc[i] * s;
  }
  __syncthreads();
  float* obase = out + ((int64_t)o * ci_total + ci_off + c0) * KK;
  for (int e = threadIdx.x; e < KK * nc; e += WR_THREADS) {
    const int c = e / KK, kk = e - c * KK;
    const float v = tile[kk * WR_PITCH + c];
    obase[e] = accumulate ? obase[e] + v : v;
  }
}

// ------------------------------------------------------------------ plain CUDA-core cross-check / small-shape path
// One thread per (co, kk, ci) of the OIHW gradient, looping over every pixel.  mode = the forward conv's mode.
__global__ void __launch_bounds__(128) wgrad_naive_kernel(int mode, const __half* __restrict__ dy,
                                                          const __half* __restrict__ x, int n, int h, int w, int cin,
                                                          int cout, int ci_total, int ci_off,
                                                          const float* __restrict__ inv_scale, float* __restrict__ out,
                                                          int accumulate) {
  // h, w: spatial size of x.  dy: mode 0/3 -> h x w, mode 1 -> h/2 x w/2, mode 2 -> 2h x 2w
  const int KK = mode == 3 ? 1 : 9;
  const int64_t total = (int64_t)cout * KK * cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % cin);
  const int kk = (int)((i / cin) % KK);
  const int co = (int)(i / ((int64_t)cin * KK));
  const int ky = mode == 3 ? 1 : kk / 3, kx = mode == 3 ? 1 : kk % 3;
  const int oh = mode == 1 ? h / 2 : (mode == 2 ? 2 * h : h), ow = mode == 1 ? w / 2 : (mode == 2 ? 2 * w : w);
  float acc = 0.f;
  for (int b = 0; b < n; ++b)
    for (int y = 0; y < oh; ++y)
      for (int xx = 0; xx < ow; ++xx) {
        int sy, sx;
        if (mode == 1) { sy = 2 * y + ky - 1; sx = 2 * xx + kx - 1; }
        else if (mode == 2) {
          const int uy = y + ky - 1, ux = xx + kx - 1;
          if (uy < 0 || uy >= oh || ux < 0 || ux >= ow) continue;
          sy = uy >> 1; sx = ux >> 1;
        } else { sy = y + ky - 1; sx = xx + kx - 1; }
        if (sy < 0 || sy >= h || sx < 0 || sx >= w) continue;
        acc = fmaf(__half2float(dy[(((int64_t)b * oh + y) * ow + xx) * cout + co]),
                   __half2float(x[(((int64_t)b * h + sy) * w + sx) * cin + ci]), acc);
      }
  const float s = inv_scale ? *inv_scale : 1.0f;
  float* op = out + ((int64_t)co * ci_total + ci_off + ci) * KK + kk;
  *op = accumulate ? *op + acc * s : acc * s;
}

// ------------------------------------------------------------------ host side
static int make_map_px(CUtensorMap* m, const IgSrc& s, int N, int TW, int box_h) { return make_map_a(m, s, N, TW, box_h); }

}  // namespace dsg

using namespace dsg;

extern "C" {

// geometry shared by the launcher and the workspace query: pixel-tile shape, work items, pixel-axis splits
static bool wg_geometry(int mode, int n, int h, int w, int cin, int cout, int* tw_o, int* sh_o, int* nchunk_o,
                        int* items_o, int* splits_o, int64_t* tiles_per_split_o) {
  const int gh = mode == 1 ? h / 2 : h, gw = mode == 1 ? w / 2 : w;
  int tw = 0, sh = 0;
  if (gw >= 16) { tw = 16; sh = 4; } else if (gw >= 8) { tw = 8; sh = 3; }
  if (!tw) return false;
  const int th = 128 / tw;
  if (gh < th + 2) return false;
  int nchunk, ngroups, phases = 1;
  if (mode == 0) { nchunk = (cin % 128 == 0) ? 2 : 1; ngroups = cout == 64 ? 2 : 3; }
  else if (mode == 3) { nchunk = (cin % 256 == 0) ? 4 : ((cin % 128 == 0) ? 2 : 1); ngroups = 1; }
  else if (mode == 1) { nchunk = 1; ngroups = 3; }
  else { nchunk = (cin % 128 == 0) ? 2 : 1; ngroups = 2; phases = 4; }
  const int items = phases * ceil_div(cout, 128) * (cin / (nchunk * 64)) * ngroups;
  const int64_t tiles_total = (int64_t)n * ceil_div(gh, th) * ceil_div(gw, tw);
  int64_t splits = num_sms() / items;
  if (splits < 1) splits = 1;
  if (splits > tiles_total) splits = tiles_total;
  const int64_t per = ceil_div64(tiles_total, splits);
  splits = ceil_div64(tiles_total, per);
  *tw_o = tw; *sh_o = sh; *nchunk_o = nchunk; *items_o = items; *splits_o = (int)splits; *tiles_per_split_o = per;
  return true;
}

int64_t dsg_wgrad_workspace_bytes(int32_t mode, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout) {
  int tw, sh, nchunk, items, splits;
  int64_t per;
  if (mode < 0 || mode > 3 || n <= 0 || !wg_geometry(mode, n, h, w, cin, cout, &tw, &sh, &nchunk, &items, &splits, &per))
    return 0;  // the CUDA-core path needs no workspace
  const int ktaps = mode == 3 ? 1 : (mode == 2 ? 16 : 9);
  return (int64_t)splits * cout * ktaps * cin * 4;
}

/* see include/dsg_b200.h */
int dsg_conv_wgrad(const dsg_wgrad_args* a, void* stream) {
  DSG_CHECK_ARG(a != nullptr, "dsg_conv_wgrad: args is null");
  DSG_CHECK_ARG(a->mode >= 0 && a->mode <= 3, "dsg_conv_wgrad: bad mode %d", a->mode);
  DSG_CHECK_ARG(a->n >= 0 && a->h > 0 && a->w > 0 && a->cin > 0 && a->cout > 0 && a->cin % 64 == 0 &&
                    a->cout % 64 == 0,
                "dsg_conv_wgrad: bad shape (channels must be multiples of 64)");
  DSG_CHECK_ARG(a->x && a->dy && a->grad, "dsg_conv_wgrad: null x/dy/grad");
  DSG_CHECK_ARG(a->ci_total >= a->ci_off + a->cin && a->ci_off >= 0, "dsg_conv_wgrad: bad ci_total/ci_off");
  DSG_CHECK_ARG((((uintptr_t)a->x | (uintptr_t)a->dy | (uintptr_t)a->workspace) % 16) == 0,
                "dsg_conv_wgrad: unaligned pointer");
  DSG_CHECK_ARG(a->mode != 1 || (a->h % 2 == 0 && a->w % 2 == 0), "dsg_conv_wgrad: stride-2 conv needs even H, W");
  cudaStream_t st = (cudaStream_t)stream;
  const int KK = a->mode == 3 ? 1 : 9;
  const int64_t out_elems = (int64_t)a->cout * KK * a->cin;
  if (a->n == 0) return DSG_OK;

  // pixel grid of dY (per phase for the upsample conv, whose dY is read through four parity views)
  const int gh = a->mode == 1 ? a->h / 2 : a->h, gw = a->mode == 1 ? a->w / 2 : a->w;
  int tw = 0, sh = 0, nchunk = 1, items = 0, splits = 1;
  int64_t tiles_per_split = 0;
  const bool tc_ok = wg_geometry(a->mode, a->n, a->h, a->w, a->cin, a->cout, &tw, &sh, &nchunk, &items, &splits,
                                 &tiles_per_split) && a->impl != 1;
  const int th = tw ? 128 / tw : 0;
  if (!tc_ok) {
    DSG_CHECK_ARG(a->impl != 2, "dsg_conv_wgrad: shape outside the tcgen05 kernel (needs W >= 8 and H >= tile + 2)");
    const int64_t blocks = ceil_div64(out_elems, 128);
    wgrad_naive_kernel<<<(unsigned)blocks, 128, 0, st>>>(a->mode, (const __half*)a->dy, (const __half*)a->x, a->n,
                                                        a->h, a->w, a->cin, a->cout, a->ci_total, a->ci_off,
                                                        a->inv_scale, a->grad, a->accumulate);
    DSG_CUDA_LAUNCH_CHECK("dsg_conv_wgrad/naive");
    return DSG_OK;
  }
  DSG_CHECK_ARG(a->workspace != nullptr, "dsg_conv_wgrad: workspace is null");

  WgPlan p;
  memset(&p, 0, sizeof(p));
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.N = a->n; p.H = gh; p.W = gw; p.TW = tw; p.tw_shift = sh; p.TH = th;
  p.tiles_w = ceil_div(gw, tw); p.tiles_h = ceil_div(gh, th);
  p.tiles_total = (int64_t)a->n * p.tiles_h * p.tiles_w;
  p.co = a->cout; p.co_blocks = ceil_div(a->cout, 128); p.n_abox = a->cout >= 128 ? 2 : 1;
  p.a_bytes = p.n_abox * WG_A_BOX;
  p.ci = a->cin;
  p.phases = 1;
  const int row_bytes = tw * 128;
  int rc;
  if (a->mode == 0) {
    p.ktaps = 9;
    p.ci_per_item = nchunk * 64;
    const int box_bytes = (th + 2) * row_bytes;
    if (a->cout == 64) {
      // two work-item kinds instead of three: {dx = -1, 0} as the two M atoms of one MMA (dY shifted by +1 / 0 against
      // the unshifted X box), and dx = +1 alone (dY shifted by -1; its second atom is don't-care as before)
      p.ngroups = 2;
      for (int g = 0; g < 2; ++g) {
        WgGroup& G = p.grp[g];
        G.nbox = nchunk; G.nunit = nchunk;
        for (int c = 0; c < nchunk; ++c) {
          G.box[c] = WgBox{0, c, 0, -1, box_bytes, c * box_bytes};
          G.unit[c] = WgUnit{c * box_bytes, 3, row_bytes, c * 256, g == 0 ? 0 : 2, 3, c * 64, 0};
        }
        G.a_nbox = g == 0 ? 2 : 1;
        G.a_chan[0] = G.a_chan[1] = 0;
        G.a_dx[0] = g == 0 ? 1 : -1; G.a_dx[1] = 0;
        G.tap_hi = g == 0 ? 1 : -1;
      }
    } else {
      p.ngroups = 3;
      for (int dx = -1; dx <= 1; ++dx) {
        WgGroup& G = p.grp[dx + 1];
        G.nbox = nchunk; G.nunit = nchunk;
        for (int c = 0; c < nchunk; ++c) {
          G.box[c] = WgBox{0, c, dx, -1, box_bytes, c * box_bytes};
          G.unit[c] = WgUnit{c * box_bytes, 3, row_bytes, c * 256, dx + 1, 3, c * 64, 0};
        }
      }
    }
    rc = make_map_px(&maps.a[0], dense_src(a->dy, a->cout, gh, gw), a->n, tw, th);
    if (rc) return rc;
    rc = make_map_px(&maps.x[0], dense_src(a->x, a->cin, a->h, a->w), a->n, tw, th + 2);
    if (rc) return rc;
  } else if (a->mode == 3) {
    p.ktaps = 1;
    p.ci_per_item = nchunk * 64;
    p.ngroups = 1;
    WgGroup& G = p.grp[0];
    G.nbox = nchunk; G.nunit = 1;
    for (int c = 0; c < nchunk; ++c) G.box[c] = WgBox{0, c, 0, 0, WG_A_BOX, c * WG_A_BOX};
    G.unit[0] = WgUnit{0, nchunk, WG_A_BOX, 0, 0, 0, 0, 64};
    rc = make_map_px(&maps.a[0], dense_src(a->dy, a->cout, gh, gw), a->n, tw, th);
    if (rc) return rc;
    rc = make_map_px(&maps.x[0], dense_src(a->x, a->cin, a->h, a->w), a->n, tw, th);
    if (rc) return rc;
  } else if (a->mode == 1) {
    // X parity views (2i + ph, 2j + pw); tap ky reads view parity {1, 0, 1}[ky] at row shift {-1, 0, 0}[ky]
    p.ktaps = 9;
    p.ci_per_item = 64;
    p.ngroups = 3;
    const int par[3] = {1, 0, 1}, shf[3] = {-1, 0, 0};
    for (int ky = 0; ky < 3; ++ky) {
      WgGroup& G = p.grp[ky];
      G.nbox = 3; G.nunit = 3;
      for (int kx = 0; kx < 3; ++kx) {
        G.box[kx] = WgBox{par[ky] * 2 + par[kx], 0, shf[kx], shf[ky], WG_A_BOX, kx * WG_A_BOX};
        G.unit[kx] = WgUnit{kx * WG_A_BOX, 1, WG_A_BOX, kx * 64, ky * 3 + kx, 0, 0, 0};
      }
    }
    rc = make_map_px(&maps.a[0], dense_src(a->dy, a->cout, gh, gw), a->n, tw, th);
    if (rc) return rc;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        IgSrc s;
        s.ptr = (const __half*)a->x + ((int64_t)ph * a->w + pw) * a->cin;
        s.C = a->cin; s.H = gh; s.W = gw;
        s.sW = 2 * (int64_t)a->cin; s.sH = 2 * (int64_t)a->w * a->cin; s.sN = (int64_t)a->h * a->w * a->cin;
        rc = make_map_px(&maps.x[ph * 2 + pw], s, a->n, tw, th);
        if (rc) return rc;
      }
  } else {  // mode 2: upsample conv — four dY parity views x 2x2 sub-pixel taps on the low-resolution X
    p.ktaps = 16;
    p.phases = 4;
    p.ci_per_item = nchunk * 64;
    p.ngroups = 2;
    const int box_bytes = (th + 1) * row_bytes;
    for (int j = 0; j < 2; ++j) {
      WgGroup& G = p.grp[j];
      G.nbox = nchunk; G.nunit = nchunk;
      for (int c = 0; c < nchunk; ++c) {
        G.box[c] = WgBox{0, c, j - 1, -1, box_bytes, c * box_bytes};
        G.unit[c] = WgUnit{c * box_bytes, 2, row_bytes, c * 128, j, 2, c * 64, 0};
      }
    }
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb) {
        IgSrc s;
        s.ptr = (const __half*)a->dy + ((int64_t)pa * (2 * a->w) + pb) * a->cout;
        s.C = a->cout; s.H = gh; s.W = gw;
        s.sW = 2 * (int64_t)a->cout; s.sH = 2 * (int64_t)(2 * a->w) * a->cout;
        s.sN = (int64_t)(2 * a->h) * (2 * a->w) * a->cout;
        rc = make_map_px(&maps.a[pa * 2 + pb], s, a->n, tw, th);
        if (rc) return rc;
      }
    rc = make_map_px(&maps.x[0], dense_src(a->x, a->cin, a->h, a->w), a->n, tw, th + 1);
    if (rc) return rc;
  }
  for (int i = 1; i < 4; ++i)
    if (p.phases == 1) maps.a[i] = maps.a[0];
  for (int i = 1; i < WG_MAX_MAPS; ++i)
    if (a->mode != 1) maps.x[i] = maps.x[0];
  for (int g = 0; g < p.ngroups; ++g) {
    WgGroup& G = p.grp[g];
    if (G.a_nbox == 0) {   // the usual A operand: consecutive 64-channel blocks of the output-channel block, unshifted
      G.a_nbox = p.n_abox;
      G.a_chan[0] = 0; G.a_chan[1] = 1;
      G.a_dx[0] = G.a_dx[1] = 0;
      G.tap_hi = -1;
    }
  }
  p.ci_groups = a->cin / p.ci_per_item;
  int b_bytes = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    int bb = 0;
    for (int b = 0; b < p.grp[g].nbox; ++b) bb += p.grp[g].box[b].bytes;
    if (bb > b_bytes) b_bytes = bb;
  }
  if (a->mode == 0 && a->cout == 64) p.a_bytes = 2 * WG_A_BOX;   // the pair kind loads two (shifted) A boxes
  p.stage_bytes = p.a_bytes + b_bytes;
  // a 64-channel layer's second (don't-care) M atom reads 16 KB past its single A box: keep that inside the stage
  if (p.n_abox == 1 && p.stage_bytes < 2 * WG_A_BOX) p.stage_bytes = 2 * WG_A_BOX;
  p.stage_bytes = (p.stage_bytes + 1023) & ~1023;
  p.nstages = WG_RING_BYTES / p.stage_bytes;
  if (p.nstages > 8) p.nstages = 8;
  DSG_CHECK_ARG(p.nstages >= 2, "dsg_conv_wgrad: stage does not fit twice in shared memory");
  DSG_CHECK_ARG(items == p.phases * p.co_blocks * p.ci_groups * p.ngroups, "dsg_conv_wgrad: internal item count");
  p.tiles_per_split = tiles_per_split;
  p.splits = splits;
  const int64_t ws_bytes = (int64_t)splits * a->cout * p.ktaps * a->cin * 4;
  DSG_CHECK_ARG(a->workspace_bytes >= ws_bytes, "dsg_conv_wgrad: workspace too small (%lld < %lld bytes)",
                (long long)a->workspace_bytes, (long long)ws_bytes);
  p.ws = (float*)a->workspace;

  const int smem_bytes = WG_RING_BYTES + 256 + 1024;
  static SmemAttrCache attr;
  {
    cudaError_t e = ensure_dyn_smem(attr, wgrad_kernel, (size_t)smem_bytes);
    if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  }
  launch_k(wgrad_kernel, dim3((unsigned)splits, (unsigned)items), dim3(WG_THREADS), smem_bytes, st, maps, p);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_wgrad/tcgen05");
  // ci_b: as wide as possible (more loads in flight per thread) while the grid still fills the machine
  int ci_b = WR_CI_MAX;
  while (ci_b > 64 && (int64_t)ceil_div(a->cin, ci_b) * a->cout < 2 * 148) ci_b >>= 1;
  {
    const int kk_out = a->mode == 2 ? 9 : p.ktaps;
    const int ept = ceil_div(kk_out * (ci_b < a->cin ? ci_b : a->cin), WR_THREADS);
    const dim3 rgrid((unsigned)ceil_div(a->cin, ci_b), (unsigned)a->cout);
#define DSG_WR_LAUNCH(E, U)                                                                                          \
    launch_k(wgrad_reduce_kernel<E, U>, rgrid, dim3(WR_THREADS), 0, st, (const float*)p.ws, splits, a->cout, p.ktaps,  \
             a->cin, ci_b, a->mode == 2 ? 1 : 0, 0, a->cout, a->ci_total, a->ci_off, a->inv_scale, a->grad,            \
             a->accumulate)
    if (ept <= 1) DSG_WR_LAUNCH(1, 16);
    else if (ept <= 3) DSG_WR_LAUNCH(3, 6);
    else if (ept <= 5) DSG_WR_LAUNCH(5, 4);
    else DSG_WR_LAUNCH(9, 2);
#undef DSG_WR_LAUNCH
  }
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_wgrad/reduce");
  return DSG_OK;
}
}
