// igemm_halo.cu — halo-reuse implicit-GEMM convolution on tcgen05 tensor cores (3x3 stride-1 and the sub-pixel
// form of nearest-2x-upsample + 3x3).  Same contraction and packed weights as igemm.cu (the tap-streaming kernel),
// but the activation tile is fetched ONCE per column shift instead of once per tap:
//
//   CTA tile = TH x TW output pixels (TW = 16 or 8, TH = MT * 128 / TW), i.e. MT stacked M = 128 accumulators.
//   For each 64-channel chunk and each dx in {-1, 0, +1} one TMA box of (TH + 2) rows x TW pixels x 64 channels
//   lands in a 128B-swizzled smem slot.  Box rows are TW * 128 B = a multiple of the 1024 B swizzle atom, so the
//   A operand of tap (dy, dx) for accumulator m is simply the SAME slot at byte offset (dy + 1 + m * 128 / TW) *
//   TW * 128 — a descriptor start-address bump, no data movement.  Out-of-bounds rows/columns are zero-filled by
//   the TMA unit, which is the conv zero padding.  Weights stream through a second, finer ring (one BLOCK_N x 64
//   tile per tap).  L2 -> smem traffic per 64-channel chunk drops from 9 x 16 KB per 128 pixels to
//   3 x (TH + 2) / TH x 16 KB, and every weight tile is reused by MT accumulators.
//
// Warp roles as in igemm.cu: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-5 = epilogue
// (TMEM -> registers -> +bias/temb/residual -> fp16 NHWC).  TMEM holds 2 x MT accumulators (double-buffered).
// Replaces the cuDNN convs behind ResnetBlock2D.conv1/conv2(+conv_shortcut) and Upsample2D of diffusers 0.20.0
// (SURVEY.md §2.2, §8 a4/a6).
#include "igemm_common.cuh"

namespace dsg {

constexpr int HL_THREADS = 192;
constexpr int HL_THREADS_TMA = 320;    // FUSE = 2: eight epilogue warps (two per TMEM lane quadrant, one accumulator each)
constexpr int HL_THREADS_FUSED = 256;  // + two GroupNorm/SiLU transform warps (256 threads keep 255 registers each)
constexpr int HL_XF_THREADS = 64;
constexpr int HL_MAX_GROUPS = 8;
constexpr int HL_MAX_MAPS = 4;

struct HlGroup {
  int map;         // which activation tensor map (source + box height)
  int dx, dy0;     // box origin relative to the tile origin (phase offset added at run time)
  int nchunks;     // 64-channel chunks of this source
  int chunk_off;   // first 64-channel chunk of this source inside the (concatenated) conv input
  int transform;   // fused GroupNorm + SiLU: the box holds RAW values and is normalised in shared memory
  int ntaps;       // taps served by one box (<= 3)
  int row_off[3];  // first box row of tap t
  int kb_base[3];  // weight k-block (64 columns each) of (tap t, chunk 0)
  int bytes;       // bytes one box delivers
};

struct HlPlan {
  HlGroup grp[HL_MAX_GROUPS];
  int ngroups;
  int N, OH, OW;
  int TW, tw_shift, TH;  // CTA tile = TH x TW pixels
  int tiles_w, tiles_h;
  int phases, omul;
  int cout, n_blocks;
  int64_t oN, oH, oW;  // output element strides
  __half* out;
  float* out_f32;      // BLOCK_N = 16 only: NCHW fp32 output of the first cout_real channels (conv_out)
  int cout_real;
  const __half* res;
  const float* bias;
  const float* temb;
  int temb_stride, temb_off;
  long long* stats;    // optional per-channel GroupNorm totals of the output: int64 [N][cout][2] (groupnorm.cu)
  const float2* coef;  // fused GroupNorm + SiLU of the input: [N][cin_main] (a / 2, b / 2); act = h + h tanh(h), h = a x / 2 + b / 2
  int cin_main;        // channels of the (concatenated) 3x3 input
  int IH, IW;          // extents of the 3x3 input (masking of the zero padding under the fused transform)
  int xpose;           // transposed epilogue stores through shared memory (epi_tile)
  int tma_out;         // the fp16 NHWC output goes through the staging buffers + TMA stores (maps.o)
  int64_t total_tiles;
};

struct alignas(64) HlMaps {
  CUtensorMap a[HL_MAX_MAPS];
  CUtensorMap b;
  CUtensorMap o;       // output tensor (tma_out only): box = 64 channels x TW pixels x 32 / TW rows, SWIZZLE_128B
};

// CG = CTAs per MMA: 1, or 2 for tcgen05 cta_group::2 — a cluster of two CTAs shares every MMA (M = 256): each CTA
// feeds its own 128 pixel rows of A but only HALF of the weight tile, which halves the per-SM shared-memory reads of
// B.  The UMMA operand fetch tops out near 76 B/clk/SM (tools/umma_probe.cu), which caps a lone CTA at 76 % / 60 %
// of the tensor peak for N = 256 / 128; the CTA pair lifts that to ~100 % / 80 %.
template <int BLOCK_N, int MT, int CG, bool STG = false>
struct HlCfg {
  static constexpr int A_SLOT = MT * 16384 + 4096;  // (MT*8 + 2) rows x 16 px x 128 B (the TW = 8 box is smaller)
  static constexpr int B_ROWS = BLOCK_N / CG;       // weight rows this CTA loads per tap
  static constexpr int B_SLOT = B_ROWS * 128;
  static constexpr int NA = BLOCK_N == 128 ? 3 : 4;
  static constexpr int NB = CG == 2 ? (BLOCK_N == 256 ? 8 : (STG ? 11 : 12))
                                    : (BLOCK_N == 16 ? 12 : (BLOCK_N == 64 ? 9 : (BLOCK_N == 128 ? 7 : 4)));
  static constexpr int TMEM_COLS = 2 * MT * BLOCK_N;
  static constexpr int RING_BYTES = NA * A_SLOT + NB * B_SLOT;
  // epilogue staging for the TMA-store form (cout = 64 layers, CTA pairs): 4 KB per epilogue warp and accumulator
  static constexpr bool STAGE = STG;
  static_assert(!STG || (BLOCK_N == 64 && CG == 2), "staging buffers are sized for the BLOCK_N = 64 CTA-pair form");
  static constexpr int STAGE_BYTES = STAGE ? 4 * MT * 4096 : 0;
  // transposed epilogue stores (epi_tile, igemm_common.cuh) wherever 8 KB are left: not the N = 256 forms (their rings
  // fill shared memory; the epilogue hides under a long main loop there) and not the single-CTA N = 128 form
  static constexpr bool XPOSE = !STG && BLOCK_N >= 32 && BLOCK_N != 256 && !(BLOCK_N == 128 && CG == 1);
  static constexpr int XPOSE_BYTES = XPOSE ? 4 * EPI_STAGE_BYTES_PER_WARP : 0;
  static constexpr int SMEM_BYTES =
      RING_BYTES + STAGE_BYTES + XPOSE_BYTES + 2 * BLOCK_N * 4 /*bias*/ + (STG ? 8 : 4) * BLOCK_N * 4 /*GN stats*/ + 512 /*barriers*/ + 1024 /*align*/;
  static_assert(3 * NA + 2 * NB + 4 <= 60, "barrier area overflow");
  static_assert(TMEM_COLS <= 512, "TMEM overflow");
  static_assert(SMEM_BYTES <= 227 * 1024, "smem overflow");
};

struct HlTile {
  int n, h0, w0, nb, pa, pb;
};
// cg CTAs of a cluster share tile t: CTA `rank` takes rows [rank * TH, (rank + 1) * TH) of the cg * TH tall tile
__device__ __forceinline__ HlTile hl_decode(const HlPlan& p, int64_t t, int cg = 1, int rank = 0) {
  HlTile c;
  c.nb = (int)(t % p.n_blocks); t /= p.n_blocks;
  const int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
  const int th = (int)(t % p.tiles_h); t /= p.tiles_h;
  c.n = (int)(t % p.N);
  const int phase = (int)(t / p.N);
  c.pa = phase >> 1; c.pb = phase & 1;
  c.h0 = (th * cg + rank) * p.TH; c.w0 = tw * p.TW;
  return c;
}

// ---- cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a location in this CTA's shared memory) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_holder) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// FUSE = 2: no input transform; the epilogue stages the fp16 tile in shared memory and writes it with TMA tensor stores
// (epi_tile_tma, igemm_common.cuh) — BLOCK_N = 64 CTA pairs only.
// FUSE = 1: the GroupNorm + SiLU that precedes the conv (ResnetBlock2D norm1/norm2 + nonlinearity, conv_norm_out +
// conv_act) is applied to the activation boxes IN SHARED MEMORY by four extra warps between the TMA landing and the
// MMA reading (per-(sample, channel) coefficients from dsg_gn_coef, packed half2 math, padding left at zero): the
// normalised tensor never exists in HBM — one read + one write of every activation less per GroupNorm.
template <int BLOCK_N, int MT, int CG, int FUSE>
__global__ void __launch_bounds__(FUSE == 1 ? HL_THREADS_FUSED : (FUSE == 2 ? HL_THREADS_TMA : HL_THREADS), 1)
igemm_halo_kernel(const __grid_constant__ HlMaps maps, const __grid_constant__ HlPlan p) {
  using Cfg = HlCfg<BLOCK_N, MT, CG, FUSE == 2>;
  constexpr int NA = Cfg::NA, NB = Cfg::NB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + NA * Cfg::A_SLOT;
  uint8_t* stage = smem + Cfg::RING_BYTES;                          // [4 warps][MT][4 KB], 1024-byte aligned
  float* sbias = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::STAGE_BYTES + Cfg::XPOSE_BYTES);  // [2][BLOCK_N]
  constexpr int EW = FUSE == 2 ? 8 : 4;                             // epilogue warps
  constexpr int EPI_THREADS = EW * 32;
  float* sstat = sbias + 2 * BLOCK_N;                               // [EW][BLOCK_N / 2][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sstat + EW * BLOCK_N);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NA;
  uint64_t* b_full = a_empty + NA;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tfull = b_empty + NB;   // [2]
  uint64_t* tempty = tfull + 2;     // [2]
  uint64_t* a_land = tempty + 2;    // [NA] (FUSE only): raw box landed in THIS CTA's shared memory
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(a_land + NA);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int half_rows = 128 >> p.tw_shift;  // image rows per M = 128 accumulator
  const int row_bytes = p.TW * 128;         // one box row in smem

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < HL_MAX_MAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    if (p.tma_out) tma_prefetch_desc(&maps.o);
  }
  // CTA pair (CG = 2): rank 0 is the leader — it owns the "full" barriers (both producers report to them), issues
  // every MMA and collects both epilogues' "accumulator drained" arrivals; "empty"/"accumulator full" barriers
  // exist in both CTAs and are signalled together by multicast commits.
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int64_t tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], CG); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], CG); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EW * CG); }
    if constexpr (FUSE == 1)
      for (int i = 0; i < NA; ++i) mbar_init(&a_land[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_holder);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  pdl_sync();  // everything above overlaps the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    // ===================================================== TMA producer (warp-uniform loop, elected lane issues)
    {
      int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
      for (int64_t t = tile0; t < p.total_tiles; t += tile_step) {
        const HlTile tc = hl_decode(p, t, CG, (int)rank);
        // a CTA pair splits the weight tile: this CTA loads rows [rank * B_ROWS, (rank + 1) * B_ROWS) of it
        const int brow = (tc.pa * 2 + tc.pb) * p.cout + tc.nb * BLOCK_N + (int)rank * Cfg::B_ROWS;
        for (int g = 0; g < p.ngroups; ++g) {
          const HlGroup& G = p.grp[g];
          const int wx = tc.w0 + G.dx + tc.pb, hy = tc.h0 + G.dy0 + tc.pa;
          for (int c = 0; c < G.nchunks; ++c) {
            mbar_wait(&a_empty[as], aph ^ 1);
            if (FUSE == 1 && G.transform) {
              // raw box -> this CTA's own "landed" barrier; the transform warps publish it to the MMA issuer
              if (elect_one_sync()) {
                mbar_arrive_expect_tx(&a_land[as], (uint32_t)G.bytes);
                tma_load_4d(a_ring + as * Cfg::A_SLOT, &maps.a[G.map], &a_land[as], c * IG_BLOCK_K, wx, hy, tc.n);
              }
            } else if (elect_one_sync()) {
              if constexpr (CG == 2) {
                const uint32_t fb = mapa_u32(&a_full[as], 0);
                mbar_arrive_expect_tx_cluster(fb, (uint32_t)G.bytes);
                tma_load_4d_cg2(a_ring + as * Cfg::A_SLOT, &maps.a[G.map], fb, c * IG_BLOCK_K, wx, hy, tc.n);
              } else {
                mbar_arrive_expect_tx(&a_full[as], (uint32_t)G.bytes);
                tma_load_4d(a_ring + as * Cfg::A_SLOT, &maps.a[G.map], &a_full[as], c * IG_BLOCK_K, wx, hy, tc.n);
              }
            }
            __syncwarp();
            if (++as == NA) { as = 0; aph ^= 1; }
            for (int tp = 0; tp < G.ntaps; ++tp) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (elect_one_sync()) {
                const int kcol = (G.kb_base[tp] + G.chunk_off + c) * IG_BLOCK_K;
                if constexpr (CG == 2) {
                  const uint32_t fb = mapa_u32(&b_full[bs], 0);
                  mbar_arrive_expect_tx_cluster(fb, (uint32_t)Cfg::B_SLOT);
                  tma_load_2d_cg2(b_ring + bs * Cfg::B_SLOT, &maps.b, fb, kcol, brow);
                } else {
                  mbar_arrive_expect_tx(&b_full[bs], (uint32_t)Cfg::B_SLOT);
                  tma_load_2d(b_ring + bs * Cfg::B_SLOT, &maps.b, &b_full[bs], kcol, brow);
                }
              }
              __syncwarp();
              if (++bs == NB) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform loop, elected lane issues)
    if (rank == 0) {
      // instruction descriptor: M = 128 per CTA, i.e. 256 for the pair
      const uint32_t idesc = (umma_idesc_f16(BLOCK_N) & ~(0x1Fu << 24)) | ((uint32_t)((128 * CG) >> 4) << 24);
      int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
      int acc = 0; uint32_t acc_ph = 0;
      for (int64_t t = tile0; t < p.total_tiles; t += tile_step) {
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * BLOCK_N);
        uint32_t accum = 0;  // the first k-step of the first tap initialises every accumulator
        for (int g = 0; g < p.ngroups; ++g) {
          const HlGroup& G = p.grp[g];
          for (int c = 0; c < G.nchunks; ++c) {
            mbar_wait(&a_full[as], aph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_ring + as * Cfg::A_SLOT);
            for (int tp = 0; tp < G.ntaps; ++tp) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint64_t db = umma_desc_sw128(smem_u32(b_ring + bs * Cfg::B_SLOT));
              const bool last_tap = (tp + 1 == G.ntaps);
              const bool last_of_tile = last_tap && (c + 1 == G.nchunks) && (g + 1 == p.ngroups);
              if (elect_one_sync()) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                  const uint64_t da = umma_desc_sw128(a_addr + (uint32_t)((G.row_off[tp] + m * half_rows) * row_bytes));
#pragma unroll
                  for (int k = 0; k < IG_BLOCK_K / 16; ++k) {
                    if constexpr (CG == 2)
                      umma_f16_cg2(d_tmem + (uint32_t)(m * BLOCK_N), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                                   idesc, k == 0 ? accum : 1u);
                    else
                      umma_f16(d_tmem + (uint32_t)(m * BLOCK_N), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                               k == 0 ? accum : 1u);
                  }
                }
                if constexpr (CG == 2) {
                  umma_commit_cg2(&b_empty[bs]);                   // both CTAs' weight slots
                  if (last_tap) umma_commit_cg2(&a_empty[as]);     // both CTAs' activation boxes
                  if (last_of_tile) umma_commit_cg2(&tfull[acc]);  // both CTAs' epilogues
                } else {
                  umma_commit(&b_empty[bs]);                 // frees the weight slot when these MMAs retire
                  if (last_tap) umma_commit(&a_empty[as]);   // frees the activation box
                  if (last_of_tile) umma_commit(&tfull[acc]);  // accumulators complete -> epilogue
                }
              }
              __syncwarp();
              accum = 1;
              if (++bs == NB) { bs = 0; bph ^= 1; }
            }
            if (++as == NA) { as = 0; aph ^= 1; }
          }
        }
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
      }
    }
  } else if (FUSE == 1 && warp >= 6) {
    // ===================================================== GroupNorm + SiLU transform warps (FUSE only)
    // A box is [pixel][64 channels] in 128-byte rows, 16-byte chunks XOR-swizzled with (pixel & 7).  Thread tt walks the
    // chunks tt, tt + 64, ...: its pixel index advances by 8 each step, so it always meets the SAME logical channel
    // group lc — its eight (a, b) pairs stay in registers for the whole box.
    const int tt = threadIdx.x - 192;
    const int lc = (tt & 7) ^ ((tt >> 3) & 7);
    int as = 0;
    uint32_t land_ph = 0;  // per-slot phase bits: a_land[s] only advances when slot s carries a TRANSFORMED box
    for (int64_t t = tile0; t < p.total_tiles; t += tile_step) {
      const HlTile tc = hl_decode(p, t, CG, (int)rank);
      for (int g = 0; g < p.ngroups; ++g) {
        const HlGroup& G = p.grp[g];
        if (!G.transform) {
          as = (as + G.nchunks) % NA;
          continue;
        }
        const int wx = tc.w0 + G.dx + tc.pb, hy = tc.h0 + G.dy0 + tc.pa;
        const int nchk = G.bytes >> 4;  // 16-byte chunks in the box
        for (int c = 0; c < G.nchunks; ++c) {
          const float4* cf = reinterpret_cast<const float4*>(p.coef + (int64_t)tc.n * p.cin_main +
                                                             (G.chunk_off + c) * 64 + lc * 8);
          __half2 a2[4], b2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = __ldg(cf + j);  // (a0, b0, a1, b1)
            a2[j] = __floats2half2_rn(v.x, v.z);
            b2[j] = __floats2half2_rn(v.y, v.w);
          }
          mbar_wait(&a_land[as], (land_ph >> as) & 1u);
          land_ph ^= 1u << as;
          uint8_t* box = a_ring + as * Cfg::A_SLOT;
#pragma unroll 4
          for (int q = tt; q < nchk; q += HL_XF_THREADS) {
            const int r = q >> 3;
            const int y = hy + (r >> p.tw_shift), x = wx + (r & (p.TW - 1));
            if (y >= 0 && y < p.IH && x >= 0 && x < p.IW) {  // padding stays zero
              uint4 v = *reinterpret_cast<uint4*>(box + (size_t)q * 16);
              __half2* hv = reinterpret_cast<__half2*>(&v);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __half2 h = __hfma2(hv[j], a2[j], b2[j]);
                uint32_t hu = *reinterpret_cast<const uint32_t*>(&h), tu;
                asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(hu));
                hv[j] = __hfma2(h, *reinterpret_cast<const __half2*>(&tu), h);
              }
              *reinterpret_cast<uint4*>(box + (size_t)q * 16) = v;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA (async proxy)
          asm volatile("bar.sync 2, 64;" ::: "memory");
          if (tt == 0) {
            if constexpr (CG == 2) {
              // plain arrive: each CTA's tensor core reads its OWN shared memory for A, and this CTA's writes were made
              // visible to its async proxy by the fence above (a cluster-scope release costs a full memory barrier)
              mbar_arrive_cluster(mapa_u32(&a_full[as], 0));
            } else {
              mbar_arrive(&a_full[as]);
            }
          }
          if (++as == NA) as = 0;
        }
      }
    }
  } else {
    // ===================================================== epilogue warps (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int te = threadIdx.x - 64;
    const int lh0 = row >> p.tw_shift, lw = row & (p.TW - 1);
    int acc = 0; uint32_t acc_ph = 0;
    EpiStatsAcc<BLOCK_N> stats_acc;
    stats_acc.init();
    bool store_pending = false;
    for (int64_t t = tile0; t < p.total_tiles; t += tile_step) {
      const HlTile tc = hl_decode(p, t, CG, (int)rank);
      const int n0 = tc.nb * BLOCK_N;
      float* sb = sbias + acc * BLOCK_N;
      for (int j = te; j < BLOCK_N; j += EPI_THREADS) {
        float v = p.bias ? p.bias[n0 + j] : 0.f;
        if (p.temb) v += p.temb[(int64_t)tc.n * p.temb_stride + p.temb_off + n0 + j];
        sb[j] = v;
      }
      int64_t off[MT];
      bool valid[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int h = tc.h0 + m * half_rows + lh0, w = tc.w0 + lw;
        valid[m] = (h < p.OH) && (w < p.OW);
        off[m] = (int64_t)tc.n * p.oN + (int64_t)(h * p.omul + tc.pa) * p.oH + (int64_t)(w * p.omul + tc.pb) * p.oW + n0;
        if (p.res && valid[m]) {
#pragma unroll
          for (int j = 0; j < BLOCK_N; j += 64) prefetch_l2(p.res + off[m] + j);
        }
      }
      if constexpr (BLOCK_N == 16) {
        // conv_out form: the first cout_real accumulator columns go to an NCHW fp32 tensor
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        mbar_wait(&tfull[acc], acc_ph);
        tc_fence_after();
        const uint32_t taddr16 = tmem_base + (uint32_t)(acc * MT * BLOCK_N) + ((uint32_t)(q * 32) << 16);
        const int64_t plane = (int64_t)p.OH * p.OW;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          uint32_t v[16];
          tmem_ld_32x16(taddr16 + (uint32_t)(m * BLOCK_N), v);
          tmem_ld_wait();
          if (valid[m]) {
            const int h = tc.h0 + m * half_rows + lh0, w = tc.w0 + lw;
            float* op = p.out_f32 + (int64_t)tc.n * p.cout_real * plane + (int64_t)h * p.OW + w;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < p.cout_real) op[j * plane] = __uint_as_float(v[j]) + sb[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));  // the leader issues the MMAs
        else mbar_arrive(&tempty[acc]);
      }
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
        continue;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * MT * BLOCK_N) + ((uint32_t)(q * 32) << 16);
      if constexpr (FUSE == 2) {
        // warps 2-5 take accumulator 0, warps 6-9 accumulator 1 (same TMEM lane quadrants): half the epilogue latency
        static_assert(MT == 2, "one epilogue warp group per stacked accumulator");
        const int mh = (warp - 2) >> 2;
        const bool valid1[1] = {valid[mh]};
        const int64_t off1[1] = {off[mh]};
        const int hrow1[1] = {tc.h0 + mh * half_rows + ((q * 32) >> p.tw_shift)};
        epi_tile_tma<BLOCK_N, 1>(taddr + (uint32_t)(mh * BLOCK_N), sb, valid1, off1, p.res,
                                 p.stats ? sstat + (mh * 4 + q) * BLOCK_N : nullptr, lane,
                                 stage + (mh * 4 + q) * 4096, &maps.o, n0, tc.w0, hrow1, tc.n, store_pending);
      } else if constexpr (BLOCK_N >= 32) {
        epi_tile<BLOCK_N, MT>(taddr, sb, valid, off, p.out, p.res, p.stats ? sstat + q * BLOCK_N : nullptr, lane,
                              (Cfg::XPOSE && p.xpose) ? stage + q * EPI_STAGE_BYTES_PER_WARP : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));  // the leader issues the MMAs
        else mbar_arrive(&tempty[acc]);
      }
      if (p.stats) {
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        stats_acc.template add_tile<EW>(sstat, te, p.stats + ((int64_t)tc.n * p.cout + n0) * 2);
      }
      acc ^= 1; if (acc == 0) acc_ph ^= 1;
    }
    stats_acc.emit(te);
    if (store_pending && lane == 0) bulk_wait_group0();   // the last tensor stores have left shared memory and landed
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA may leave while the other still signals / reads it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
// DSG_TMA_OUT=0 keeps the direct-store epilogue (A/B switch)
static bool tma_out_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DSG_TMA_OUT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// DSG_HALO_1X1=0 sends 1x1 convs back to the tap-streaming kernel (A/B switch)
static bool halo_1x1_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DSG_HALO_1X1");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int BLOCK_N, int MT, int CG, int FUSE>
static int launch_halo(const dsg_conv_args* a, cudaStream_t st) {
  using Cfg = HlCfg<BLOCK_N, MT, CG, FUSE == 2>;
  HlPlan p;
  memset(&p, 0, sizeof(p));
  const int oh = a->h, ow = a->w;  // GEMM pixel grid (per phase for mode 2)
  int tw, sh;
  if (ow >= 16) { tw = 16; sh = 4; } else if (ow >= 8) { tw = 8; sh = 3; } else return DSG_HALO_SKIP;
  const int th = MT * (128 / tw);
  const int halo = a->mode == 0 ? 2 : (a->mode == 3 ? 0 : 1);
  if (oh < th + halo) return DSG_HALO_SKIP;  // keep every TMA box inside the tensor extents
  if (FUSE == 1 && a->mode != 0) return DSG_HALO_SKIP;                  // fused GroupNorm: plain 3x3 only
  if (FUSE == 2 && a->mode != 0 && a->mode != 3) return DSG_HALO_SKIP;  // TMA-store form: unit-stride outputs only
  p.N = a->n; p.OH = oh; p.OW = ow; p.TW = tw; p.tw_shift = sh; p.TH = th;
  p.tiles_w = ceil_div(ow, tw); p.tiles_h = ceil_div(oh, th * CG);  // a CTA pair stacks its two tiles vertically
  p.cout = a->cout; p.n_blocks = a->cout / BLOCK_N;
  const int cin_chunks = a->cin / 64;
  // the 3x3 input is one tensor, or (fused GroupNorm form) the channel concatenation of two raw tensors
  const int cin1 = (FUSE == 1 && a->x2) ? a->cin1 : a->cin;
  const int nsrc = cin1 < a->cin ? 2 : 1;
  const int src_chunks[2] = {cin1 / 64, (a->cin - cin1) / 64};
  const int src_map[2] = {0, 3};
  HlMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = make_map_a(&maps.a[0], dense_src(a->x, cin1, a->h, a->w), a->n, tw, th + halo);
  if (rc) return rc;
  if (nsrc == 2) {
    rc = make_map_a(&maps.a[3], dense_src(a->x2, a->cin - cin1, a->h, a->w), a->n, tw, th + halo);
    if (rc) return rc;
  }
  int64_t k_total;
  if (a->mode == 0) {
    p.phases = 1; p.omul = 1;
    for (int dx = -1; dx <= 1; ++dx) {
      for (int s = 0; s < nsrc; ++s) {
        HlGroup& G = p.grp[p.ngroups++];
        G.map = src_map[s]; G.dx = dx; G.dy0 = -1; G.nchunks = src_chunks[s]; G.ntaps = 3;
        G.chunk_off = s == 0 ? 0 : src_chunks[0];
        G.transform = FUSE == 1;
        for (int dy = -1; dy <= 1; ++dy) {
          G.row_off[dy + 1] = dy + 1;
          G.kb_base[dy + 1] = ((dy + 1) * 3 + (dx + 1)) * cin_chunks;
        }
        G.bytes = (th + 2) * tw * 128;
      }
    }
    int kb = 9 * cin_chunks;
    const void* scp[2] = {a->sc1, a->sc2};
    const int scc[2] = {a->csc1, a->csc2};
    for (int s = 0; s < 2; ++s) {
      if (!scc[s]) continue;
      IgSrc ss = dense_src(scp[s], scc[s], a->h, a->w);
      rc = make_map_a(&maps.a[1 + s], ss, a->n, tw, th);
      if (rc) return rc;
      HlGroup& G = p.grp[p.ngroups++];
      G.map = 1 + s; G.dx = 0; G.dy0 = 0; G.nchunks = scc[s] / 64; G.ntaps = 1;
      G.chunk_off = 0; G.transform = 0;
      G.row_off[0] = 0; G.kb_base[0] = kb;
      G.bytes = th * tw * 128;
      kb += scc[s] / 64;
    }
    k_total = (int64_t)kb * 64;
  } else if (a->mode == 3) {
    // 1x1 conv / linear layer: one tap, no halo — the "shortcut panel" group alone.  With K this short the kernel is
    // all epilogue, which is what the CTA-pair forms (and the eight-warp TMA-store form for N = 64) are good at.
    p.phases = 1; p.omul = 1;
    HlGroup& G = p.grp[p.ngroups++];
    G.map = 0; G.dx = 0; G.dy0 = 0; G.nchunks = cin_chunks; G.ntaps = 1;
    G.chunk_off = 0; G.transform = 0;
    G.row_off[0] = 0; G.kb_base[0] = 0;
    G.bytes = th * tw * 128;
    k_total = a->cin;
  } else {  // mode 2: four sub-pixel phases, each a 2x2 conv at the input resolution
    p.phases = 4; p.omul = 2;
    for (int j = 0; j < 2; ++j) {
      HlGroup& G = p.grp[p.ngroups++];
      G.map = 0; G.dx = j - 1; G.dy0 = -1; G.nchunks = cin_chunks; G.ntaps = 2;
      G.chunk_off = 0; G.transform = 0;
      for (int i = 0; i < 2; ++i) {
        G.row_off[i] = i;
        G.kb_base[i] = (i * 2 + j) * cin_chunks;
      }
      G.bytes = (th + 1) * tw * 128;
    }
    k_total = (int64_t)4 * a->cin;
  }
  rc = make_map_b(&maps.b, (const __half*)a->wpacked, k_total, (int64_t)p.phases * a->cout, Cfg::B_ROWS);
  if (rc) return rc;
  const int out_h = oh * p.omul, out_w = ow * p.omul;
  p.oW = a->cout; p.oH = (int64_t)out_w * a->cout; p.oN = (int64_t)out_h * out_w * a->cout;
  p.out = (__half*)a->out; p.res = (const __half*)a->residual;
  p.out_f32 = (float*)a->out_nchw_f32; p.cout_real = a->cout_real;
  p.bias = a->bias; p.temb = a->temb; p.temb_stride = a->temb_stride; p.temb_off = a->temb_off;
  p.stats = (long long*)a->out_stats;
  p.coef = (const float2*)a->gn_coef; p.cin_main = a->cin; p.IH = a->h; p.IW = a->w;
  p.total_tiles = (int64_t)p.phases * p.N * p.tiles_h * p.tiles_w * p.n_blocks;
  p.xpose = epi_xpose_enabled() ? 1 : 0;
  if constexpr (FUSE == 2) {
    static_assert(Cfg::STAGE, "no staging buffers in this configuration");
    rc = make_map_a(&maps.o, dense_src(a->out, a->cout, oh, ow), a->n, tw, 32 / tw);
    if (rc) return rc;
    p.tma_out = 1;
  }
  static SmemAttrCache attr;
  {
    cudaError_t e = ensure_dyn_smem(attr, igemm_halo_kernel<BLOCK_N, MT, CG, FUSE>, (size_t)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("igemm_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  }
  const int threads = FUSE == 1 ? HL_THREADS_FUSED : (FUSE == 2 ? HL_THREADS_TMA : HL_THREADS);
  const int64_t slots = num_sms() / CG;  // CTAs, or CTA pairs
  const int64_t grid = (p.total_tiles < slots ? p.total_tiles : slots) * CG;
  if constexpr (CG == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, igemm_halo_kernel<BLOCK_N, MT, CG, FUSE>, maps, p);
    if (e != cudaSuccess) { set_error("igemm_halo (CTA pair): launch: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  } else {
    launch_k(igemm_halo_kernel<BLOCK_N, MT, CG, FUSE>, dim3((unsigned)grid), dim3(threads), Cfg::SMEM_BYTES, st, maps, p);
  }
  DSG_CUDA_LAUNCH_CHECK("dsg_conv/igemm_halo");
  return DSG_OK;
}

int launch_halo_conv(const dsg_conv_args* a, int block_n, int cta_pair, cudaStream_t st) {
  if (a->mode != 0 && a->mode != 2 && a->mode != 3) return DSG_HALO_SKIP;
  if (a->mode == 3 && (!cta_pair || a->csc1 || a->csc2 || !halo_1x1_enabled())) return DSG_HALO_SKIP;
  if (a->gn_coef) {  // fused GroupNorm + SiLU on the input (mode 0 only)
    if (cta_pair) {
      switch (block_n) {
        case 64: return launch_halo<64, 2, 2, 1>(a, st);
        case 128: return launch_halo<128, 2, 2, 1>(a, st);
        case 256: return launch_halo<256, 1, 2, 1>(a, st);
        default: return DSG_HALO_SKIP;
      }
    }
    switch (block_n) {
      case 16: return launch_halo<16, 2, 1, 1>(a, st);
      case 64: return launch_halo<64, 2, 1, 1>(a, st);
      case 128: return launch_halo<128, 2, 1, 1>(a, st);
      case 256: return launch_halo<256, 1, 1, 1>(a, st);
      default: return DSG_HALO_SKIP;
    }
  }
  if (cta_pair) {
    switch (block_n) {
      case 64:
        // FUSE = 2: fp16 NHWC output staged in shared memory and written by TMA tensor stores (plain 3x3 convs)
        // — the layers whose main loop is short enough (3x3 over 64 channels, + at most a 64-channel shortcut panel) for
        // the epilogue to pace the kernel; with more K the deeper weight ring of the plain form wins (measured)
        if (a->out && !a->out_nchw_f32 && tma_out_enabled() &&
            ((a->mode == 0 && a->cin == 64 && a->csc1 + a->csc2 <= 64) || a->mode == 3))
          return launch_halo<64, 2, 2, 2>(a, st);
        return launch_halo<64, 2, 2, 0>(a, st);
      case 128: return launch_halo<128, 2, 2, 0>(a, st);
      case 256: return launch_halo<256, 1, 2, 0>(a, st);
      default: return DSG_HALO_SKIP;
    }
  }
  switch (block_n) {
    case 16: return launch_halo<16, 2, 1, 0>(a, st);
    case 64: return launch_halo<64, 2, 1, 0>(a, st);
    case 128: return launch_halo<128, 2, 1, 0>(a, st);
    case 256: return launch_halo<256, 1, 1, 0>(a, st);
    default: return DSG_HALO_SKIP;
  }
}

}  // namespace dsg
