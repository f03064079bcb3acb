// igemm.cu — implicit-GEMM convolution on tcgen05 tensor cores: the U-Net's dense contractions.
//
// Replaces every cuDNN conv the reference reaches through diffusers 0.20.0 ResnetBlock2D.conv1/conv2/conv_shortcut,
// Downsample2D.conv, Upsample2D (nearest-2x + conv) and the attention Linear layers:
//   GEMM  D[pixel][cout] = sum_k A[pixel][k] * W[cout][k],  k = (tap, cin-chunk)  (+ appended 1x1 shortcut channels)
// A tiles are fetched by TMA straight from the NHWC fp16 activation: one 4-D box (64 ch x TW x TH x 1) per
// (tap, 64-channel chunk) at shifted coordinates; out-of-bounds elements are zero-filled by the TMA unit, which IS
// the conv zero padding.  Stride-2 convs read four parity views of the input (pure stride tricks, same box loads);
// nearest-2x-upsample+conv runs as four 2x2 sub-pixel convs on pre-summed weights.  W tiles are 2-D boxes of the
// packed [rows][K] fp16 weight.  Both land in 128B-swizzled K-major shared memory and feed tcgen05.mma
// (M=128, N=BLOCK_N, K=16, fp16 in / fp32 accumulate in TMEM).  Persistent, warp-specialised:
//   warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> regs -> +bias/temb/residual -> fp16).
// TMEM accumulators are double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
#include "igemm_common.cuh"

namespace dsg {

constexpr int IG_THREADS = 192;
constexpr int IG_MAX_TAPS = 16;
constexpr int IG_A_BYTES = IG_BLOCK_M * IG_BLOCK_K * 2;  // 16 KB

struct IgTap {
  int src, dy, dx, nchunks;
};
struct IgPlan {
  IgSrc src[IG_MAX_SRC];
  IgTap taps[IG_MAX_TAPS];
  int nsrc, ntaps, num_kb;
  int N, OH, OW;      // GEMM pixel grid per phase
  int TW, TH, tw_shift;  // tile = TH x TW pixels (TW a power of two, TH*TW = 128)
  int tiles_w, tiles_h;
  int phases;         // 1, or 4 for the sub-pixel upsample conv
  int cout, n_blocks;
  int64_t k_total;
  int omul;           // output pixel = (h*omul + a, w*omul + b)
  int64_t oN, oH, oW; // output element strides
  const __half* w;
  __half* out;
  const __half* res;
  const float* bias;
  const float* temb;
  int temb_stride, temb_off;
  long long* stats;   // optional per-channel GroupNorm totals of the output: int64 [N][cout][2] (groupnorm.cu)
  int a_bytes;        // bytes one A box delivers (TH clipped to the image height)
  int xpose;           // transposed epilogue stores through shared memory (epi_tile)
  int64_t total_tiles;
};
struct alignas(64) IgMaps {
  CUtensorMap a[IG_MAX_SRC];
  CUtensorMap b;
};

template <int BLOCK_N>
struct IgCfg {
  static constexpr int B_BYTES = BLOCK_N * IG_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = IG_A_BYTES + B_BYTES;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;  // 256 -> 4, 128 -> 6, 64 -> 8
  static constexpr int TMEM_COLS = 2 * BLOCK_N;              // double-buffered fp32 accumulator
  static constexpr int EPI_STAGE = 4 * EPI_STAGE_BYTES_PER_WARP;   // transposed epilogue stores (igemm_common.cuh)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE + 2 * BLOCK_N * 4 /*bias*/ +
                                    4 * BLOCK_N * 4 /*GN stats*/ + 256 /*barriers*/ + 1024 /*align*/;
};

struct TileCoord {
  int n, h0, w0, nb, pa, pb;
};
__device__ __forceinline__ TileCoord decode_tile(const IgPlan& p, int64_t t) {
  TileCoord c;
  c.nb = (int)(t % p.n_blocks); t /= p.n_blocks;
  const int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
  const int th = (int)(t % p.tiles_h); t /= p.tiles_h;
  c.n = (int)(t % p.N);
  const int phase = (int)(t / p.N);
  c.pa = phase >> 1; c.pb = phase & 1;
  c.h0 = th * p.TH; c.w0 = tw * p.TW;
  return c;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ IgMaps maps, const __grid_constant__ IgPlan p) {
  using Cfg = IgCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;                       // [4 warps][2 KB]
  float* sbias = reinterpret_cast<float*>(epi_stage + Cfg::EPI_STAGE);         // [2][BLOCK_N]
  float* sstat = sbias + 2 * BLOCK_N;                                          // [4][BLOCK_N / 2][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sstat + 4 * BLOCK_N);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nsrc; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  pdl_sync();  // setup above overlaps the previous kernel's tail

  // Roles run warp-uniform; only the TMA / tcgen05 instructions are under elect_one_sync() (see common.cuh).
  if (warp == 0) {
    // ===================================================== TMA producer
    {
      int stage = 0; uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int brow = (tc.pa * 2 + tc.pb) * p.cout + tc.nb * BLOCK_N;
        int kb = 0;
        for (int e = 0; e < p.ntaps; ++e) {
          const IgTap tp = p.taps[e];
          const int hy = tc.h0 + tp.dy + tc.pa, wx = tc.w0 + tp.dx + tc.pb;
          for (int ch = 0; ch < tp.nchunks; ++ch, ++kb) {
            mbar_wait(&empty_bar[stage], ph ^ 1);
            uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(p.a_bytes + Cfg::B_BYTES));
              tma_load_4d(sa, &maps.a[tp.src], &full_bar[stage], ch * IG_BLOCK_K, wx, hy, tc.n);
              tma_load_2d(sa + IG_A_BYTES, &maps.b, &full_bar[stage], kb * IG_BLOCK_K, brow);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    {
      const uint32_t idesc = umma_idesc_f16(BLOCK_N);
      int stage = 0; uint32_t ph = 0;
      int acc = 0; uint32_t acc_ph = 0;
      for (int64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + IG_A_BYTES);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < IG_BLOCK_K / 16; ++k)
              umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (kb + 1 == p.num_kb) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
      }
    }
  } else {
    // ===================================================== epilogue warps (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int te = threadIdx.x - 64;
    int acc = 0; uint32_t acc_ph = 0;
    EpiStatsAcc<BLOCK_N> stats_acc;
    stats_acc.init();
    for (int64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      const int n0 = tc.nb * BLOCK_N;
      float* sb = sbias + acc * BLOCK_N;
      for (int j = te; j < BLOCK_N; j += 128) {
        float v = p.bias ? p.bias[n0 + j] : 0.f;
        if (p.temb) v += p.temb[(int64_t)tc.n * p.temb_stride + p.temb_off + n0 + j];
        sb[j] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const int lh = row >> p.tw_shift, lw = row & (p.TW - 1);
      const int h = tc.h0 + lh, w = tc.w0 + lw;
      const bool valid[1] = {(h < p.OH) && (w < p.OW)};
      const int64_t off[1] = {(int64_t)tc.n * p.oN + (int64_t)(h * p.omul + tc.pa) * p.oH +
                              (int64_t)(w * p.omul + tc.pb) * p.oW + n0};
      const uint32_t taddr = tmem_base + (uint32_t)(acc * BLOCK_N) + ((uint32_t)(q * 32) << 16);
      epi_tile<BLOCK_N, 1>(taddr, sb, valid, off, p.out, p.res, p.stats ? sstat + q * BLOCK_N : nullptr, lane,
                           p.xpose ? epi_stage + q * EPI_STAGE_BYTES_PER_WARP : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (p.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        stats_acc.add_tile(sstat, te, p.stats + ((int64_t)tc.n * p.cout + n0) * 2);
      }
      acc ^= 1; if (acc == 0) acc_ph ^= 1;
    }
    stats_acc.emit(te);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ plain CUDA-core cross-check (impl = 1)
// Same plan, same packed weights, fp32 accumulation; one thread per (pixel, cout).  Slow; tests/debugging only.
__global__ void __launch_bounds__(256) igemm_naive_kernel(const __grid_constant__ IgPlan p) {
  const int64_t total = (int64_t)p.phases * p.N * p.OH * p.OW * p.cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int co = (int)(t % p.cout); t /= p.cout;
    const int w = (int)(t % p.OW); t /= p.OW;
    const int h = (int)(t % p.OH); t /= p.OH;
    const int n = (int)(t % p.N);
    const int phase = (int)(t / p.N);
    const int pa = phase >> 1, pb = phase & 1;
    const __half* wr = p.w + ((int64_t)phase * p.cout + co) * p.k_total;
    float acc = 0.f;
    int64_t k = 0;
    for (int e = 0; e < p.ntaps; ++e) {
      const IgTap tp = p.taps[e];
      const IgSrc& s = p.src[tp.src];
      const int hy = h + tp.dy + pa, wx = w + tp.dx + pb;
      const int nc = tp.nchunks * IG_BLOCK_K;
      if (hy >= 0 && hy < s.H && wx >= 0 && wx < s.W) {
        const __half* ap = s.ptr + (int64_t)n * s.sN + (int64_t)hy * s.sH + (int64_t)wx * s.sW;
        for (int c = 0; c < nc; ++c) acc = fmaf(__half2float(ap[c]), __half2float(wr[k + c]), acc);
      }
      k += nc;
    }
    if (p.bias) acc += p.bias[co];
    if (p.temb) acc += p.temb[(int64_t)n * p.temb_stride + p.temb_off + co];
    const int64_t off = (int64_t)n * p.oN + (int64_t)(h * p.omul + pa) * p.oH + (int64_t)(w * p.omul + pb) * p.oW + co;
    if (p.res) acc += __half2float(p.res[off]);
    p.out[off] = __float2half_rn(acc);
    if (p.stats) {
      long long* o = p.stats + ((int64_t)n * p.cout + co) * 2;
      atomicAdd(reinterpret_cast<unsigned long long*>(o), (unsigned long long)gn_fix_sum(acc));
      atomicAdd(reinterpret_cast<unsigned long long*>(o + 1), (unsigned long long)gn_fix_sq(acc * acc));
    }
  }
}

// ------------------------------------------------------------------ weight packing (fp32 OIHW -> fp16 GEMM rows)
// mode 0/1: row co, k = (ky*3+kx)*cin + ci, then k = 9*cin + c for the 1x1 shortcut
// mode 2  : row (a*2+b)*cout + co, k = (i*2+j)*cin + ci, weight = sum of the 3x3 taps that land on source
//           offset (i + a - 1, j + b - 1) after nearest-2x upsampling
// mode 3  : row co, k = ci (w is [cout][cin])
// modes 10-13: the data-gradient ("dgrad") forms of modes 0-3 — the same fp32 OIHW weight, packed so that the
//           gradient w.r.t. the conv INPUT is itself a dsg_conv over the output gradient (cout/cin below are the
//           FORWARD conv's): 10 -> run as mode 0, 11 -> run as mode 2, 12 -> run as mode 4, 13 -> run as mode 3
// one packed element: row r, column k of the [rows][k_total] fp16 GEMM weight (modes above)
__device__ __forceinline__ float pack_value(int mode, const float* __restrict__ w, int cout, int cin,
                                            const float* __restrict__ wsc, int csc, int64_t r, int64_t k) {
    float v = 0.f;
    if (mode == 0 || mode == 1) {
      if (k < (int64_t)9 * cin) {
        const int tap = (int)(k / cin), ci = (int)(k - (int64_t)tap * cin);
        v = w[((r * cin + ci) * 3 + tap / 3) * 3 + tap % 3];
      } else {
        v = wsc[r * csc + (k - (int64_t)9 * cin)];
      }
    } else if (mode == 2) {
      const int phase = (int)(r / cout), co = (int)(r - (int64_t)phase * cout);
      const int a = phase >> 1, b = phase & 1;
      const int tap = (int)(k / cin), ci = (int)(k - (int64_t)tap * cin);
      const int ti = tap >> 1, tj = tap & 1;
      for (int ky = 0; ky < 3; ++ky) {
        // source-row offset of upsampled row (2y + a + ky - 1) relative to y is floor((a + ky - 1) / 2)
        const int oy = (a + ky - 1) >= 0 ? (a + ky - 1) / 2 : -1;
        if (oy != ti + a - 1) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ox = (b + kx - 1) >= 0 ? (b + kx - 1) / 2 : -1;
          if (ox != tj + b - 1) continue;
          v += w[(((int64_t)co * cin + ci) * 3 + ky) * 3 + kx];
        }
      }
    } else if (mode == 3) {
      v = w[r * cin + k];
    } else if (mode == 10) {  // dgrad of mode 0: row ci, k = (ky'*3+kx')*cout + co, spatially flipped taps
      const int tap = (int)(k / cout), co = (int)(k - (int64_t)tap * cout);
      v = w[(((int64_t)co * cin + r) * 3 + (2 - tap / 3)) * 3 + (2 - tap % 3)];
    } else if (mode == 11) {  // dgrad of mode 1 (stride 2), run as a mode-2 conv over dY: row (pa*2+pb)*cin + ci
      const int phase = (int)(r / cin), ci = (int)(r - (int64_t)phase * cin);
      const int pa = phase >> 1, pb = phase & 1;
      const int tap = (int)(k / cout), co = (int)(k - (int64_t)tap * cout);
      const int ky = 3 - 2 * (tap >> 1) - pa, kx = 3 - 2 * (tap & 1) - pb;
      if (ky >= 0 && ky < 3 && kx >= 0 && kx < 3) v = w[(((int64_t)co * cin + ci) * 3 + ky) * 3 + kx];
    } else if (mode == 12) {  // dgrad of mode 2 (upsample conv), run as a mode-4 conv: k = ((a*2+b)*4 + i*2+j)*cout + co
      const int e = (int)(k / cout), co = (int)(k - (int64_t)e * cout);
      const int a = e >> 3, b = (e >> 2) & 1, ti = (e >> 1) & 1, tj = e & 1;
      for (int ky = 0; ky < 3; ++ky) {
        const int oy = (a + ky - 1) >= 0 ? (a + ky - 1) / 2 : -1;
        if (oy != ti + a - 1) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ox = (b + kx - 1) >= 0 ? (b + kx - 1) / 2 : -1;
          if (ox != tj + b - 1) continue;
          v += w[(((int64_t)co * cin + r) * 3 + ky) * 3 + kx];
        }
      }
    } else {  // mode 13: dgrad of mode 3 (transpose)
      v = w[k * cin + r];
    }
    return v;
}

__global__ void __launch_bounds__(256) pack_weight_kernel(int mode, const float* __restrict__ w, int cout, int cin,
                                                          const float* __restrict__ wsc, int csc,
                                                          __half* __restrict__ out, int64_t k_total, int64_t rows) {
  const int64_t total = rows * k_total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k_total, k = i - r * k_total;
    out[i] = __float2half_rn(pack_value(mode, w, cout, cin, wsc, csc, r, k));
  }
}

static int build_plan(const dsg_conv_args* a, IgPlan& p) {
  memset(&p, 0, sizeof(p));
  DSG_CHECK_ARG(a->mode >= 0 && a->mode <= 4, "dsg_conv: bad mode %d", a->mode);
  DSG_CHECK_ARG(a->n >= 0 && a->h > 0 && a->w > 0, "dsg_conv: bad shape");
  DSG_CHECK_ARG(a->n == 0 || (a->x && a->wpacked && (a->out || a->out_nchw_f32)), "dsg_conv: null x/wpacked/out");
  DSG_CHECK_ARG(a->cin > 0 && a->cin % 64 == 0 && a->cout > 0 && (a->cout % 64 == 0 || a->out_nchw_f32),
                "dsg_conv: cin (%d) and cout (%d) must be multiples of 64", a->cin, a->cout);
  DSG_CHECK_ARG(a->csc1 % 64 == 0 && a->csc2 % 64 == 0 && a->csc1 >= 0 && a->csc2 >= 0,
                "dsg_conv: shortcut channels must be multiples of 64");
  DSG_CHECK_ARG((a->sc1 != nullptr) == (a->csc1 > 0) && (a->sc2 != nullptr) == (a->csc2 > 0),
                "dsg_conv: shortcut pointer/channel mismatch");
  DSG_CHECK_ARG(a->csc2 == 0 || a->csc1 > 0, "dsg_conv: sc2 requires sc1");
  DSG_CHECK_ARG((a->csc1 == 0) || a->mode == 0, "dsg_conv: shortcut inputs only with mode 0");
  DSG_CHECK_ARG((((uintptr_t)a->x | (uintptr_t)a->sc1 | (uintptr_t)a->sc2 | (uintptr_t)a->wpacked |
                  (uintptr_t)a->out | (uintptr_t)a->residual) % 16) == 0,
                "dsg_conv: tensor pointers must be 16-byte aligned");
  const int cin_chunks = a->cin / 64;
  p.N = a->n; p.cout = a->cout; p.phases = 1; p.omul = 1;
  int oh = a->h, ow = a->w;
  if (a->mode == 0 || a->mode == 3) {
    p.src[0] = dense_src(a->x, a->cin, a->h, a->w);
    p.nsrc = 1;
    if (a->mode == 0) {
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) p.taps[p.ntaps++] = IgTap{0, ky - 1, kx - 1, cin_chunks};
      if (a->csc1) {
        p.src[p.nsrc] = dense_src(a->sc1, a->csc1, a->h, a->w);
        p.taps[p.ntaps++] = IgTap{p.nsrc++, 0, 0, a->csc1 / 64};
      }
      if (a->csc2) {
        p.src[p.nsrc] = dense_src(a->sc2, a->csc2, a->h, a->w);
        p.taps[p.ntaps++] = IgTap{p.nsrc++, 0, 0, a->csc2 / 64};
      }
    } else {
      p.taps[p.ntaps++] = IgTap{0, 0, 0, cin_chunks};
    }
  } else if (a->mode == 1) {
    DSG_CHECK_ARG(a->h % 2 == 0 && a->w % 2 == 0, "dsg_conv: stride-2 conv needs even H, W");
    oh = a->h / 2; ow = a->w / 2;
    // four parity views: view (ph, pw) holds input pixels (2i + ph, 2j + pw)
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        IgSrc s;
        s.ptr = (const __half*)a->x + ((int64_t)ph * a->w + pw) * a->cin;
        s.C = a->cin; s.H = oh; s.W = ow;
        s.sW = 2 * (int64_t)a->cin; s.sH = 2 * (int64_t)a->w * a->cin; s.sN = (int64_t)a->h * a->w * a->cin;
        p.src[ph * 2 + pw] = s;
      }
    p.nsrc = 4;
    // input row 2*oh + ky - 1: ky=0 -> (oh-1, parity 1); ky=1 -> (oh, parity 0); ky=2 -> (oh, parity 1)
    const int par[3] = {1, 0, 1}, sh[3] = {-1, 0, 0};
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx)
        p.taps[p.ntaps++] = IgTap{par[ky] * 2 + par[kx], sh[ky], sh[kx], cin_chunks};
  } else if (a->mode == 4) {
    // adjoint of mode 2 (nearest-2x upsample + 3x3): a 4x4 stride-2 gather over the four parity views of the
    // high-resolution gradient, weights = the sub-pixel phase weights (pack mode 12)
    DSG_CHECK_ARG(a->h % 2 == 0 && a->w % 2 == 0, "dsg_conv: mode 4 needs even H, W");
    oh = a->h / 2; ow = a->w / 2;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        IgSrc s;
        s.ptr = (const __half*)a->x + ((int64_t)ph * a->w + pw) * a->cin;
        s.C = a->cin; s.H = oh; s.W = ow;
        s.sW = 2 * (int64_t)a->cin; s.sH = 2 * (int64_t)a->w * a->cin; s.sN = (int64_t)a->h * a->w * a->cin;
        p.src[ph * 2 + pw] = s;
      }
    p.nsrc = 4;
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb)
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) p.taps[p.ntaps++] = IgTap{pa * 2 + pb, -(i + pa - 1), -(j + pb - 1), cin_chunks};
  } else {  // mode 2
    p.src[0] = dense_src(a->x, a->cin, a->h, a->w);
    p.nsrc = 1; p.phases = 4; p.omul = 2;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) p.taps[p.ntaps++] = IgTap{0, i - 1, j - 1, cin_chunks};
  }
  p.num_kb = 0;
  for (int e = 0; e < p.ntaps; ++e) p.num_kb += p.taps[e].nchunks;
  p.k_total = (int64_t)p.num_kb * 64;
  p.OH = oh; p.OW = ow;
  int tw = 1, sh = 0;
  while (tw * 2 <= ow && tw * 2 <= 128) { tw *= 2; ++sh; }
  p.TW = tw; p.tw_shift = sh; p.TH = 128 / tw;
  p.tiles_w = ceil_div(ow, p.TW); p.tiles_h = ceil_div(oh, p.TH);
  const int box_h = p.TH < oh ? p.TH : oh;
  p.a_bytes = 128 * p.TW * box_h;
  const int out_h = oh * p.omul, out_w = ow * p.omul;
  p.oW = a->cout; p.oH = (int64_t)out_w * a->cout; p.oN = (int64_t)out_h * out_w * a->cout;
  p.w = (const __half*)a->wpacked; p.out = (__half*)a->out; p.res = (const __half*)a->residual;
  p.bias = a->bias; p.temb = a->temb; p.temb_stride = a->temb_stride; p.temb_off = a->temb_off;
  p.stats = (long long*)a->out_stats;
  return DSG_OK;
}

template <int BLOCK_N>
static int launch_igemm(const IgPlan& plan_in, cudaStream_t st) {
  using Cfg = IgCfg<BLOCK_N>;
  IgPlan p = plan_in;
  p.n_blocks = p.cout / BLOCK_N;
  p.total_tiles = (int64_t)p.phases * p.N * p.tiles_h * p.tiles_w * p.n_blocks;
  p.xpose = epi_xpose_enabled() ? 1 : 0;
  IgMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int box_h = p.a_bytes / (128 * p.TW);
  for (int i = 0; i < p.nsrc; ++i) {
    int rc = make_map_a(&maps.a[i], p.src[i], p.N, p.TW, box_h);
    if (rc) return rc;
  }
  int rc = make_map_b(&maps.b, p.w, p.k_total, (int64_t)p.phases * p.cout, BLOCK_N);
  if (rc) return rc;
  static SmemAttrCache attr;
  {
    cudaError_t e = ensure_dyn_smem(attr, igemm_kernel<BLOCK_N>, (size_t)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("igemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  }
  int64_t grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  launch_k(igemm_kernel<BLOCK_N>, dim3((unsigned)grid), dim3(IG_THREADS), Cfg::SMEM_BYTES, st, maps, p);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv/igemm");
  return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

extern "C" {

int64_t dsg_packed_k(int32_t mode, int32_t cin, int32_t csc) {
  switch (mode) {
    case 0: case 1: return (int64_t)9 * cin + csc;
    case 2: return (int64_t)4 * cin;
    case 3: return cin;
    default: return -1;
  }
}
int64_t dsg_packed_rows(int32_t mode, int32_t cout) { return mode == 2 ? (int64_t)4 * cout : cout; }
/* dgrad packings (pack modes 10-13): K / rows from the FORWARD conv's cin, cout */
int64_t dsg_packed_k_dgrad(int32_t fwd_mode, int32_t cout) {
  switch (fwd_mode) {
    case 0: return (int64_t)9 * cout;
    case 1: return (int64_t)4 * cout;
    case 2: return (int64_t)16 * cout;
    case 3: return cout;
    default: return -1;
  }
}
int64_t dsg_packed_rows_dgrad(int32_t fwd_mode, int32_t cin) { return fwd_mode == 1 ? (int64_t)4 * cin : cin; }

int dsg_pack_conv_weight(int32_t mode, const float* w_oihw, int32_t cout, int32_t cin, const float* w_sc,
                         int32_t csc, void* wpacked, void* stream) {
  DSG_CHECK_ARG(((mode >= 0 && mode <= 3) || (mode >= 10 && mode <= 13)) && w_oihw && wpacked && cout > 0 && cin > 0,
                "dsg_pack_conv_weight: bad args");
  DSG_CHECK_ARG((csc > 0) == (w_sc != nullptr) && (csc == 0 || mode == 0), "dsg_pack_conv_weight: shortcut mismatch");
  const int64_t k_total = mode >= 10 ? dsg_packed_k_dgrad(mode - 10, cout) : dsg_packed_k(mode, cin, csc);
  const int64_t rows = mode >= 10 ? dsg_packed_rows_dgrad(mode - 10, cin) : dsg_packed_rows(mode, cout);
  int64_t blocks = ceil_div64(rows * k_total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mode, w_oihw, cout, cin, w_sc, csc,
                                                                        (__half*)wpacked, k_total, rows);
  DSG_CUDA_LAUNCH_CHECK("dsg_pack_conv_weight");
  return DSG_OK;
}

int dsg_conv_gn_fusable(int32_t mode, int32_t h, int32_t w, int32_t cin, int32_t cin1, int32_t cout) {
  if (mode != 0 || cin % 64 || cin1 % 64 || cin1 <= 0 || cin1 > cin) return 0;
  if (!(cout == 16 || cout % 64 == 0)) return 0;
  const int tw = w >= 16 ? 16 : (w >= 8 ? 8 : 0);
  if (!tw) return 0;
  const int mt = (cout % 256 == 0) ? 1 : 2;   // MT of the kernel variant dsg_conv picks
  return h >= mt * (128 / tw) + 2 ? 1 : 0;
}

int dsg_conv(const dsg_conv_args* a, void* stream) {
  DSG_CHECK_ARG(a != nullptr, "dsg_conv: args is null");
  IgPlan p;
  int rc = build_plan(a, p);
  if (rc) return rc;
  if (a->n == 0) return DSG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->impl == 1) {
    DSG_CHECK_ARG(a->out && !a->out_nchw_f32, "dsg_conv: the cross-check kernel writes the h16 NHWC output only");
    DSG_CHECK_ARG(!a->gn_coef && !a->x2, "dsg_conv: the cross-check kernel has no fused GroupNorm input");
    p.n_blocks = 1;
    const int64_t total = (int64_t)p.phases * p.N * p.OH * p.OW * p.cout;
    int64_t blocks = ceil_div64(total, 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    igemm_naive_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
    DSG_CUDA_LAUNCH_CHECK("dsg_conv/naive");
    return DSG_OK;
  }
  DSG_CHECK_ARG(a->impl == 0 || (a->impl >= 2 && a->impl <= 4), "dsg_conv: bad impl %d", a->impl);
  if (a->gn_coef) {
    DSG_CHECK_ARG(a->mode == 0 && a->impl != 2, "dsg_conv: fused GroupNorm input needs mode 0 on the halo-reuse kernels");
    DSG_CHECK_ARG((a->x2 == nullptr) ? (a->cin1 == a->cin || a->cin1 == 0)
                                     : (a->cin1 > 0 && a->cin1 < a->cin && a->cin1 % 64 == 0),
                  "dsg_conv: bad cin1 %d for cin %d", a->cin1, a->cin);
    DSG_CHECK_ARG(((uintptr_t)a->gn_coef | (uintptr_t)a->x2) % 16 == 0, "dsg_conv: gn_coef / x2 must be 16-byte aligned");
  } else {
    DSG_CHECK_ARG(a->x2 == nullptr, "dsg_conv: x2 is only valid together with gn_coef");
  }
  if (a->out_nchw_f32) {
    DSG_CHECK_ARG(a->mode == 0 && a->cout == 16 && a->cout_real >= 1 && a->cout_real <= 16 && !a->residual &&
                      !a->csc1 && a->impl != 2 && (uintptr_t)a->out_nchw_f32 % 4 == 0,
                  "dsg_conv: conv_out form needs mode 0, cout 16, 1 <= cout_real <= 16, no residual/shortcut");
    rc = launch_halo_conv(a, 16, 0, st);
    if (rc == DSG_HALO_SKIP) { set_error("dsg_conv: conv_out form needs W >= 8 and H >= 18"); return DSG_ERR_UNSUPPORTED; }
    return rc;
  }
  int bn = a->block_n;
  if (bn == 0) bn = (a->cout % 256 == 0) ? 256 : (a->cout % 128 == 0 ? 128 : 64);
  // 1x1 convs over few channels are all epilogue: N = 64 tiles get the eight-warp TMA-store form of the halo kernel
  // (the activation tile is re-read from L2 once per 64 output channels, which is cheap at cin <= 128)
  if (a->block_n == 0 && a->mode == 3 && a->cin <= 128 && a->cout % 64 == 0) bn = 64;
  DSG_CHECK_ARG((bn == 64 || bn == 128 || bn == 256) && a->cout % bn == 0, "dsg_conv: bad block_n %d for cout %d", bn,
                a->cout);
  if (a->impl != 2) {
    // default order: CTA-pair halo kernel (fastest everywhere it applies, bit-identical to the single-CTA form),
    // then the single-CTA halo kernel, then the tap-streaming kernel
    if (a->impl == 0 || a->impl == 4) {
      rc = launch_halo_conv(a, bn, 1, st);
      if (rc != DSG_HALO_SKIP) return rc;
    }
    if (a->impl != 4) {
      rc = launch_halo_conv(a, bn, 0, st);
      if (rc != DSG_HALO_SKIP) return rc;
    }
    if (a->impl >= 3 || a->gn_coef) {
      set_error("dsg_conv: shape/mode not covered by the halo-reuse kernel");
      return DSG_ERR_UNSUPPORTED;
    }
  }
  switch (bn) {
    case 64: return launch_igemm<64>(p, st);
    case 128: return launch_igemm<128>(p, st);
    default: return launch_igemm<256>(p, st);
  }
}
}
