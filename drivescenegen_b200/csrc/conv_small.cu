// conv_small.cu — the two degenerate convolutions of the U-Net (K = 27 and N = 3): CUDA-core, HBM-bound.
// Replaces UNet2DModel.conv_in (3->C0, NCHW fp32 in, NHWC fp16 out) and conv_out (C0->3, NHWC fp16 in,
// NCHW fp32 out) of diffusers 0.20.0 models/unet_2d.py.
#include "common.cuh"

namespace dsg {

constexpr int CS_THREADS = 256;

// thread = (4 adjacent pixels of one image row, 8 output channels); persistent blocks stage the weights once.
// Weights in smem as [k = ci*9 + ky*3 + kx][cout]: one pair of LDS.128 feeds 32 FMAs.
constexpr int CI_PX = 4;
__global__ void __launch_bounds__(CS_THREADS, 2) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, __half* __restrict__ out,
                                                             int n, int cin, int h, int wd, int cout,
                                                             const float* __restrict__ xscale) {
  extern __shared__ float sw[];  // [cin*9][cout] then bias[cout]
  const int K = cin * 9;
  pdl_sync();
  const float xs = xscale ? *xscale : 1.0f;   // a power of two folded into the weights: exact
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) {
    const int k = i / cout, co = i - k * cout;  // w is [cout][cin][3][3] = [cout][K]
    sw[i] = w[co * K + k] * xs;
  }
  float* sb = sw + K * cout;
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sb[i] = b[i];
  __syncthreads();
  const int gpp = cout >> 3;                 // channel groups per pixel
  const int qpb = CS_THREADS / gpp;          // pixel quads per block iteration
  const int cg = threadIdx.x % gpp, ql = threadIdx.x / gpp;
  if (ql >= qpb) return;
  const int wq = (wd + CI_PX - 1) / CI_PX;   // quads per image row
  const int64_t total_q = (int64_t)n * h * wq;
  const int64_t hw = (int64_t)h * wd;
  float bias8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias8[j] = sb[cg * 8 + j];
  for (int64_t qi = (int64_t)blockIdx.x * qpb + ql; qi < total_q; qi += (int64_t)gridDim.x * qpb) {
    const int xq = (int)(qi % wq);
    const int64_t r = qi / wq;
    const int y = (int)(r % h), nn = (int)(r / h);
    const int x0 = xq * CI_PX;
    float acc[CI_PX][8];
#pragma unroll
    for (int px = 0; px < CI_PX; ++px)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[px][j] = bias8[j];
    for (int ci = 0; ci < cin; ++ci) {
      const float* xp = x + ((int64_t)nn * cin + ci) * hw;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= h) continue;
        float v[CI_PX + 2];
#pragma unroll
        for (int u = 0; u < CI_PX + 2; ++u) {
          const int xc = x0 + u - 1;
          v[u] = (xc >= 0 && xc < wd) ? __ldg(xp + (int64_t)yy * wd + xc) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wr = sw + (ci * 9 + ky * 3 + kx) * cout + cg * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
#pragma unroll
          for (int px = 0; px < CI_PX; ++px) {
            const float a = v[px + kx];
            acc[px][0] = fmaf(a, w0.x, acc[px][0]); acc[px][1] = fmaf(a, w0.y, acc[px][1]);
            acc[px][2] = fmaf(a, w0.z, acc[px][2]); acc[px][3] = fmaf(a, w0.w, acc[px][3]);
            acc[px][4] = fmaf(a, w1.x, acc[px][4]); acc[px][5] = fmaf(a, w1.y, acc[px][5]);
            acc[px][6] = fmaf(a, w1.z, acc[px][6]); acc[px][7] = fmaf(a, w1.w, acc[px][7]);
          }
        }
      }
    }
    __half* op = out + (((int64_t)nn * h + y) * wd + x0) * cout + cg * 8;
#pragma unroll
    for (int px = 0; px < CI_PX; ++px)
      if (x0 + px < wd) stg_v4(op + (int64_t)px * cout, pack8(acc[px]));
  }
}

// ------------------------------------------------------------------ conv_in on the tensor cores, statistics fused
// K = cin * 9 <= 32 and N = 64: far too thin for a tcgen05 tile (the K axis would be padded 20x), but on CUDA cores the
// layer is FP32-FMA bound at ~3x its HBM time.  This form uses warp-level mma.sync m16n8k16 (fp16 operands, fp32
// accumulate — the numerics of every other conv of the U-Net): a warp owns one image row of the block's 8-row strip and
// walks it 16 pixels at a time.  A fragments are gathered from a (8 + 2) x (W + 2) fp16 window staged in shared memory
// (one 16-bit load per element; the k -> (ci, ky, kx) offsets are per-thread constants), all B fragments (the whole
// 32 x 64 weight matrix) live in 32 registers per thread, the fp16 NHWC result goes straight to global memory, and the
// per-channel sum / sum of squares of the next GroupNorm ride along in registers (one int64 atomic per (sample,
// channel, statistic) per strip) — the separate statistics pass over conv_in's output is gone.
constexpr int CM_ROWS = 8;    // rows per strip = warps per block
constexpr int CM_NT = 8;      // n-tiles of 8 output channels (cout = 64)
constexpr int CM_OPITCH = 144; // bytes per pixel row of the per-warp output staging tile

__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16_16(uint32_t addr) {   // the same element 8 pixels to the right
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+16];" : "=h"(v) : "r"(addr));
  return v;
}

// FULL: the row length is a multiple of 16 (no partial tile, no bounds checks in the pixel loop)
template <bool FULL>
__global__ void __launch_bounds__(CM_ROWS * 32, 2) conv_in_mma_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ b,
                                                                      __half* __restrict__ out,
                                                                      long long* __restrict__ stats, int n, int cin,
                                                                      int h, int wd,
                                                                      const float* __restrict__ xscale) {
  extern __shared__ __align__(16) unsigned char cm_smem[];
  const int P = (wd + 2 + 15) & ~15;          // window pitch (halves): column 0 = image column -1; the tail is zero
  const int plane = (CM_ROWS + 2) * P;
  // [cin + 2][CM_ROWS + 2][P]: plane `cin` is all ones, plane `cin + 1` all zeros.  K slots cin * 9 and cin * 9 + 1
  // gather ones and multiply the bias (split into an fp16 head and an fp16 remainder: exact to ~2^-22), the remaining
  // padded slots gather zeros — no bias registers, no predicates in the pixel loop.
  __half* s_in = reinterpret_cast<__half*>(cm_smem);
  float* s_stat = reinterpret_cast<float*>(cm_smem + (size_t)(cin + 2) * plane * 2);  // [8][64][2]
  // per-warp staging of one 16-pixel x 64-channel output tile (pitch 144 B: conflict-free both ways): the accumulator
  // fragments hold 4-byte pieces of 8 different pixels, the tile itself is 2 KB of contiguous NHWC memory
  unsigned char* s_out = reinterpret_cast<unsigned char*>(s_stat + CM_ROWS * 64 * 2) + (threadIdx.x >> 5) * (16 * CM_OPITCH);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = cin * 9;
  for (int i = threadIdx.x; i < plane; i += CM_ROWS * 32) {
    s_in[cin * plane + i] = __float2half_rn(1.f);
    s_in[(cin + 1) * plane + i] = __float2half_rn(0.f);
  }
  pdl_sync();
  // optional power-of-two input scale (conv_out's data gradient: the raw loss gradient is far below fp16 range)
  const float xs = xscale ? *xscale : 1.0f;
  // B fragments: b0 = W[k = 16s + 2t, 2t + 1][n = 8j + g], b1 = W[k + 8, k + 9][n]; W[k][n] = w[n][k] (OIHW, k < K)
  uint32_t bf[2][CM_NT][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int j = 0; j < CM_NT; ++j)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k0 = ks * 16 + hh * 8 + 2 * t, nn = j * 8 + g;
        float wv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = k0 + u;
          const float bh = __half2float(__float2half_rn(b[nn]));
          wv[u] = k < K ? w[nn * K + k] : (k == K ? bh : (k == K + 1 ? b[nn] - bh : 0.f));
        }
        const __half2 hv = __floats2half2_rn(wv[0], wv[1]);
        bf[ks][j][hh] = *reinterpret_cast<const uint32_t*>(&hv);
      }
  // A gather: element i of the thread's fragment list is k = (i >> 2) * 16 + ((i >> 1) & 1) * 8 + 2t + (i & 1);
  // byte offset of that tap inside the window, relative to (row of the warp, pixel column)
  uint32_t aoff[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = (i >> 2) * 16 + ((i >> 1) & 1) * 8 + 2 * t + (i & 1);
    int ci = k < K + 2 ? cin : cin + 1, ky = 0, kx = 0;   // bias slots -> ones plane, the rest -> zeros plane
    if (k < K) { ci = k / 9; ky = (k - ci * 9) / 3; kx = k - ci * 9 - ky * 3; }
    aoff[i] = (uint32_t)(ci * plane + ky * P + kx) * 2u;
  }

  const int strips = (h + CM_ROWS - 1) / CM_ROWS;
  const int64_t items = (int64_t)n * strips;
  const int64_t hw = (int64_t)h * wd;
  const uint32_t row_u32 = smem_u32(s_in) + (uint32_t)(warp * P + g) * 2u;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int nn = (int)(item / strips), y0 = (int)(item % strips) * CM_ROWS;
    __syncthreads();   // the previous strip's window and statistics scratch are no longer in use
    // window rows are spread over the warps; all loads of a row chunk (8 x 32 columns) are issued before the first
    // shared-memory store — a load-store-load-store loop would put one DRAM round trip per 32 columns on the critical path
    for (int rr = warp; rr < cin * (CM_ROWS + 2); rr += CM_ROWS) {
      const int ci = rr / (CM_ROWS + 2), r = rr - ci * (CM_ROWS + 2);
      const int yy = y0 - 1 + r;
      const bool row_ok = yy >= 0 && yy < h;
      const float* xp = x + ((int64_t)nn * cin + ci) * hw + (int64_t)yy * wd;
      __half* sp = s_in + ci * plane + r * P;
      for (int c0 = 0; c0 < P; c0 += 8 * 32) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int xx = c0 + u * 32 + lane - 1;
          v[u] = row_ok && xx >= 0 && xx < wd ? __ldg(xp + xx) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + u * 32 + lane;
          if (c < P) sp[c] = __float2half_rn(v[u] * xs);
        }
      }
    }
    __syncthreads();
    float ssum[CM_NT][2], ssq[CM_NT][2];
#pragma unroll
    for (int j = 0; j < CM_NT; ++j) { ssum[j][0] = ssum[j][1] = 0.f; ssq[j][0] = ssq[j][1] = 0.f; }
    const int y = y0 + warp;
    if (y < h) {
      unsigned char* otile = reinterpret_cast<unsigned char*>(out + (((int64_t)nn * h + y) * wd) * 64) + lane * 16;
      unsigned char* so_w = s_out + g * CM_OPITCH + t * 4;                    // fragment piece (pixel g, n-tile 0)
      const unsigned char* so_r = s_out + (lane >> 3) * CM_OPITCH + (lane & 7) * 16;  // 16 B of pixel lane / 8 (+ 4i)
      uint32_t pa = row_u32;
      for (int x0 = 0; x0 < wd; x0 += 16, pa += 32, otile += 16 * 128) {
        // fragment registers: 0 = (pixel g, k lo pair), 1 = (pixel g + 8, k lo), 2 = (g, k hi), 3 = (g + 8, k hi).
        // The window pitch is padded to 16 columns, so a partial last tile reads zeros / finite values it never stores.
        uint32_t af[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int i0 = ks * 4 + hh * 2;
            af[ks][hh * 2] = lds_u16(pa + aoff[i0]) | (lds_u16(pa + aoff[i0 + 1]) << 16);
            af[ks][hh * 2 + 1] = lds_u16_16(pa + aoff[i0]) | (lds_u16_16(pa + aoff[i0 + 1]) << 16);
          }
        const bool va = FULL || x0 + g < wd, vb = FULL || x0 + g + 8 < wd;
#pragma unroll
        for (int j = 0; j < CM_NT; ++j) {
          float d[4] = {0.f, 0.f, 0.f, 0.f};
          mma_m16n8k16(d, af[0], bf[0][j][0], bf[0][j][1]);
          mma_m16n8k16(d, af[1], bf[1][j][0], bf[1][j][1]);
          *reinterpret_cast<__half2*>(so_w + j * 16) = __floats2half2_rn(d[0], d[1]);
          *reinterpret_cast<__half2*>(so_w + 8 * CM_OPITCH + j * 16) = __floats2half2_rn(d[2], d[3]);
          if (FULL) {
            ssum[j][0] += d[0] + d[2]; ssum[j][1] += d[1] + d[3];
            ssq[j][0] = fmaf(d[0], d[0], fmaf(d[2], d[2], ssq[j][0]));
            ssq[j][1] = fmaf(d[1], d[1], fmaf(d[3], d[3], ssq[j][1]));
          } else {
            if (va) {
              ssum[j][0] += d[0]; ssum[j][1] += d[1];
              ssq[j][0] = fmaf(d[0], d[0], ssq[j][0]); ssq[j][1] = fmaf(d[1], d[1], ssq[j][1]);
            }
            if (vb) {
              ssum[j][0] += d[2]; ssum[j][1] += d[3];
              ssq[j][0] = fmaf(d[2], d[2], ssq[j][0]); ssq[j][1] = fmaf(d[3], d[3], ssq[j][1]);
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // 4 x (32 lanes x 16 B) = the tile's 2 KB, in address order
          if (FULL || x0 + i * 4 + (lane >> 3) < wd)
            stg_v4(otile + i * 512, *reinterpret_cast<const uint4*>(so_r + i * 4 * CM_OPITCH));
        }
        __syncwarp();
      }
    }
    // statistics of the strip: over the 8 pixel lanes of the warp (fixed butterfly), then over the 8 warps (fixed order)
#pragma unroll
    for (int j = 0; j < CM_NT; ++j)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float a = ssum[j][u], q = ssq[j][u];
#pragma unroll
        for (int off = 4; off <= 16; off <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, off);
          q += __shfl_xor_sync(0xffffffffu, q, off);
        }
        if (g == 0) {
          s_stat[(warp * 64 + j * 8 + 2 * t + u) * 2] = a;
          s_stat[(warp * 64 + j * 8 + 2 * t + u) * 2 + 1] = q;
        }
      }
    __syncthreads();
    if (stats != nullptr && threadIdx.x < 128) {
      const int c = threadIdx.x >> 1, k = threadIdx.x & 1;
      float tot = 0.f;
#pragma unroll
      for (int wv = 0; wv < CM_ROWS; ++wv) tot += s_stat[(wv * 64 + c) * 2 + k];
      const long long fx = k ? gn_fix_sq(tot) : gn_fix_sum(tot);
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + ((int64_t)nn * 64 + c) * 2 + k), (unsigned long long)fx);
    }
  }
}

// ------------------------------------------------------------------ conv_out on the tensor cores, GroupNorm+SiLU fused
// conv_norm_out + SiLU + conv_out (64 -> out_channels <= 8) in one pass over the raw 64-channel tensor: a block stages a
// (8 + 2) x (32 + 2)-pixel window with cp.async (double-buffered: the next window is in flight while this one is worked on), applies y = silu(a_c x + b_c) to it in place (the coefficients of
// dsg_gn_coef and the arithmetic of gn_apply_kernel: the staged fp16 values are bit-identical to the unfused path), then
// every warp runs one image row as 2 interleaved m16n8k16 mma.sync tiles: A fragments by ldmatrix.x4 from the window
// (pixel pitch 144 B: conflict-free), all 36 B fragments (K = 9 taps x 64 channels, N = 8) in 72 registers per thread,
// NCHW fp32 stores.  The activated tensor is never written to or re-read from HBM (inference only: the training program
// keeps it for conv_out's weight gradient).
constexpr int CO_ROWS = 8, CO_TW = 32, CO_WW = CO_TW + 2, CO_PITCH = 144;
constexpr int CO_WIN_BYTES = (CO_ROWS + 2) * CO_WW * CO_PITCH;   // one window; the kernel keeps two

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__global__ void __launch_bounds__(CO_ROWS * 32, 2) conv_out_mma_kernel(const __half* __restrict__ x,
                                                                       const float2* __restrict__ coef,
                                                                       const float* __restrict__ w,
                                                                       const float* __restrict__ b,
                                                                       float* __restrict__ out, int n, int h, int wd,
                                                                       int cout) {
  extern __shared__ __align__(16) unsigned char co_smem[];   // 2 x [CO_ROWS + 2][CO_WW] pixels x CO_PITCH bytes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  pdl_sync();
  // B fragments: k-step s = tap * 4 + c16; b0 = W[k = 2t, 2t + 1][n = g], b1 = W[k + 8, k + 9][g]; W[k][n] = w[n][c][tap]
  uint32_t bf[36][2];
#pragma unroll
  for (int s = 0; s < 36; ++s)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int tap = s >> 2, c = (s & 3) * 16 + hh * 8 + 2 * t;
      const float w0 = g < cout ? w[(g * 64 + c) * 9 + tap] : 0.f;
      const float w1 = g < cout ? w[(g * 64 + c + 1) * 9 + tap] : 0.f;
      const __half2 hv = __floats2half2_rn(w0, w1);
      bf[s][hh] = *reinterpret_cast<const uint32_t*>(&hv);
    }
  const float bias0 = 2 * t < cout ? b[2 * t] : 0.f, bias1 = 2 * t + 1 < cout ? b[2 * t + 1] : 0.f;
  // ldmatrix lane address: matrix m = lane / 8 -> pixel rows (lane % 8) + 8 * (m & 1), channel offset 8 * (m >> 1)
  const uint32_t lm_off = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * CO_PITCH + (lane >> 4) * 16);
  const uint32_t win_u32 = smem_u32(co_smem);
  // staging: the thread owns one 16-byte channel chunk of window pixels pslot, pslot + 32, ... (an even deal: a
  // column-wise ownership left the two threads of the 2-pixel halo with twice the work and everyone else at the barrier)
  const int chunk = threadIdx.x & 7, pslot = threadIdx.x >> 3;
  constexpr int NPX = (CO_ROWS + 2) * CO_WW, NPIECE = (NPX + 31) / 32;

  const int strips = (h + CO_ROWS - 1) / CO_ROWS, ctiles = (wd + CO_TW - 1) / CO_TW;
  const int64_t items = (int64_t)n * strips * ctiles;
  const int64_t hw = (int64_t)h * wd;

  // the raw window of `item` into buffer `buf` with cp.async (zeros outside the image), one commit group per call
  auto issue = [&](int64_t item, int buf) {
    if (item < items) {
      const int ct = (int)(item % ctiles);
      const int64_t r2 = item / ctiles;
      const int y0 = (int)(r2 % strips) * CO_ROWS, nn = (int)(r2 / strips), x0 = ct * CO_TW;
      const __half* gsrc = x + (((int64_t)nn * h + (y0 - 1)) * wd + (x0 - 1)) * 64 + chunk * 8;
      const uint32_t sdst = win_u32 + (uint32_t)(buf * CO_WIN_BYTES) + (uint32_t)(chunk * 16);
#pragma unroll
      for (int k = 0; k < NPIECE; ++k) {
        const int p = pslot + 32 * k;        // window pixel: pieces are dealt round-robin, every thread gets 10 or 11
        if (p >= NPX) break;
        const int r = p / CO_WW, c = p - r * CO_WW;
        const int yy = y0 - 1 + r, xx = x0 - 1 + c;
        const uint32_t dst = sdst + (uint32_t)(p * CO_PITCH);
        if (yy >= 0 && yy < h && xx >= 0 && xx < wd) {
          const __half* src = gsrc + ((int64_t)r * wd + c) * 64;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        } else {
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  issue(blockIdx.x, 0);
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x, buf ^= 1) {
    // the window of the NEXT item is fetched while this one is transformed and multiplied: its buffer was last read in
    // the previous iteration, which ended with a block barrier
    issue(item + gridDim.x, buf ^ 1);
    const int ct = (int)(item % ctiles);
    const int64_t r2 = item / ctiles;
    const int y0 = (int)(r2 % strips) * CO_ROWS, nn = (int)(r2 / strips), x0 = ct * CO_TW;
    // per-sample GroupNorm coefficients of this thread's 8 channels: (a / 2, b / 2)
    float ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 cf = coef ? coef[(int64_t)nn * 64 + chunk * 8 + j] : make_float2(0.f, 0.f);
      ca[j] = cf.x; cb[j] = cf.y;
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the group just committed has landed
    // GroupNorm + SiLU in place on the pieces this thread fetched (no block barrier needed in between).  Pixels outside
    // the image stay zero: the conv pads the ACTIVATED tensor.
    const uint32_t wbase = win_u32 + (uint32_t)(buf * CO_WIN_BYTES);
    if (coef) {
#pragma unroll
      for (int k = 0; k < NPIECE; ++k) {
        const int p = pslot + 32 * k;
        if (p >= NPX) break;
        const int r = p / CO_WW, c = p - r * CO_WW;
        const int yy = y0 - 1 + r, xx = x0 - 1 + c;
        if (yy < 0 || yy >= h || xx < 0 || xx >= wd) continue;
        const uint32_t a = wbase + (uint32_t)(p * CO_PITCH + chunk * 16);
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // gn_apply_kernel's arithmetic: fp32, rounded once to fp16
          const float hh = fmaf(f[kk], ca[kk], cb[kk]);
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hh));
          f[kk] = fmaf(hh, th, hh);
        }
        v = pack8(f);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
    __syncthreads();
    const int y = y0 + warp;
    if (y < h) {
      constexpr int NT = CO_TW / 16;
      float acc[NT][4];
#pragma unroll
      for (int i = 0; i < NT; ++i) { acc[i][0] = bias0; acc[i][1] = bias1; acc[i][2] = bias0; acc[i][3] = bias1; }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t rowb = wbase + (uint32_t)(((warp + ky) * CO_WW + kx) * CO_PITCH) + lm_off;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
#pragma unroll
          for (int i = 0; i < NT; ++i) {
            uint32_t af[4];
            ldmatrix_x4(af, rowb + (uint32_t)(i * 16 * CO_PITCH + c16 * 32));
            mma_m16n8k16(acc[i], af, bf[tap * 4 + c16][0], bf[tap * 4 + c16][1]);
          }
        }
      }
      // NCHW fp32: lane (g, t) holds channels 2t, 2t + 1 of pixels x0 + 16 i + g and + 8
      float* orow = out + ((int64_t)nn * cout * h + y) * wd;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        const int xa = x0 + i * 16 + g, xb = xa + 8;
        if (2 * t < cout) {
          if (xa < wd) orow[(int64_t)(2 * t) * hw + xa] = acc[i][0];
          if (xb < wd) orow[(int64_t)(2 * t) * hw + xb] = acc[i][2];
        }
        if (2 * t + 1 < cout) {
          if (xa < wd) orow[(int64_t)(2 * t + 1) * hw + xa] = acc[i][1];
          if (xb < wd) orow[(int64_t)(2 * t + 1) * hw + xb] = acc[i][3];
        }
      }
    }
    __syncthreads();   // every warp is done with this window before the next iteration refills it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// warp = 8 adjacent pixels x 4 lanes; each lane owns cin/4 input channels; lanes reduced by shuffle.
// Weights in smem as [tap][co][cin].
template <int MAXCO>
__global__ void __launch_bounds__(CS_THREADS) conv_out_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, float* __restrict__ out,
                                                              int n, int cin, int h, int wd, int cout) {
  extern __shared__ float sw[];  // [9][cout][cin]
  for (int i = threadIdx.x; i < cout * cin * 9; i += blockDim.x) {
    // w is [cout][cin][3][3]
    const int co = i / (cin * 9), r = i - co * cin * 9, ci = r / 9, tap = r - ci * 9;
    sw[(tap * cout + co) * cin + ci] = w[i];
  }
  __syncthreads();
  const int lane4 = threadIdx.x & 3;
  const int64_t hw = (int64_t)h * wd, total = hw * n;
  const int64_t p = (int64_t)blockIdx.x * (CS_THREADS / 4) + (threadIdx.x >> 2);
  const bool valid = p < total;
  const int64_t pc = valid ? p : 0;
  const int nn = (int)(pc / hw);
  const int rem = (int)(pc - (int64_t)nn * hw);
  const int y = rem / wd, xx = rem - y * wd;
  const int cpl = cin >> 2;  // channels per lane (multiple of 8)
  const int c_lo = lane4 * cpl;
  float acc[MAXCO];
#pragma unroll
  for (int j = 0; j < MAXCO; ++j) acc[j] = 0.f;
  if (valid) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xc = xx + kx - 1;
        if (xc < 0 || xc >= wd) continue;
        const __half* xp = x + (((int64_t)nn * h + yy) * wd + xc) * cin + c_lo;
        const float* wt = sw + (ky * 3 + kx) * cout * cin + c_lo;
        for (int c8 = 0; c8 < cpl; c8 += 8) {
          float f[8];
          unpack8(*reinterpret_cast<const uint4*>(xp + c8), f);
#pragma unroll
          for (int j = 0; j < MAXCO; ++j) {
            if (j < cout) {
              const float4 w0 = *reinterpret_cast<const float4*>(wt + j * cin + c8);
              const float4 w1 = *reinterpret_cast<const float4*>(wt + j * cin + c8 + 4);
              acc[j] = fmaf(f[0], w0.x, acc[j]); acc[j] = fmaf(f[1], w0.y, acc[j]);
              acc[j] = fmaf(f[2], w0.z, acc[j]); acc[j] = fmaf(f[3], w0.w, acc[j]);
              acc[j] = fmaf(f[4], w1.x, acc[j]); acc[j] = fmaf(f[5], w1.y, acc[j]);
              acc[j] = fmaf(f[6], w1.z, acc[j]); acc[j] = fmaf(f[7], w1.w, acc[j]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXCO; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
  }
  if (valid && lane4 == 0) {
#pragma unroll
    for (int j = 0; j < MAXCO; ++j)
      if (j < cout) out[((int64_t)nn * cout + j) * hw + rem] = acc[j] + b[j];
  }
}

}  // namespace dsg

using namespace dsg;

extern "C" {

static int launch_conv_in_mma(const float* x, const float* xscale, const float* w, const float* b, void* out_h16,
                              void* stats, int n, int cin, int h, int wd, int cout, cudaStream_t st);

static int conv_in_f32(const float* x, const float* xscale, const float* w, const float* b, void* out_h16, int32_t n,
                       int32_t cin, int32_t h, int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x && w && b && out_h16, "dsg_conv_in: null pointer");
  DSG_CHECK_ARG(cin >= 1 && cin <= 4 && cout % 8 == 0 && cout >= 8 && cout <= 512, "dsg_conv_in: cin<=4, cout%%8==0");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_in: bad shape");
  DSG_CHECK_ARG((uintptr_t)out_h16 % 16 == 0, "dsg_conv_in: out must be 16-byte aligned");
  if (n == 0) return DSG_OK;
  const int gpp = cout / 8, qpb = CS_THREADS / gpp;
  const int64_t total_q = (int64_t)n * h * ((wd + CI_PX - 1) / CI_PX);
  const size_t sm = (size_t)(cin * 9 * cout + cout) * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  int64_t blocks = ceil_div64(total_q, qpb);
  if (blocks > 148 * 2) blocks = 148 * 2;  // persistent: the weight staging is paid once per block
  launch_k(conv_in_kernel, dim3((unsigned)blocks), dim3(CS_THREADS), sm, (cudaStream_t)stream, x, w, b,
           (__half*)out_h16, n, cin, h, wd, cout, xscale);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_in");
  return DSG_OK;
}

int dsg_conv_in(const float* x, const float* w, const float* b, void* out_h16, int32_t n, int32_t cin, int32_t h,
                int32_t wd, int32_t cout, void* stream) {
  return conv_in_f32(x, nullptr, w, b, out_h16, n, cin, h, wd, cout, stream);
}

int dsg_conv_in_scaled(const float* x, const float* xscale, const float* w, const float* b, void* out_h16, int32_t n,
                       int32_t cin, int32_t h, int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x && xscale && w && b && out_h16, "dsg_conv_in_scaled: null pointer");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_in_scaled: bad shape");
  if (n == 0) return DSG_OK;
  if (launch_conv_in_mma(x, xscale, w, b, out_h16, nullptr, n, cin, h, wd, cout, (cudaStream_t)stream) == 0) {
    DSG_CUDA_LAUNCH_CHECK("dsg_conv_in_scaled");
    return DSG_OK;
  }
  return conv_in_f32(x, xscale, w, b, out_h16, n, cin, h, wd, cout, stream);
}

// tensor-core form when the shape allows it; returns 1 when it does not
static int launch_conv_in_mma(const float* x, const float* xscale, const float* w, const float* b, void* out_h16,
                              void* stats, int n, int cin, int h, int wd, int cout, cudaStream_t st) {
  const size_t sm = (size_t)(cin + 2) * (CM_ROWS + 2) * ((wd + 2 + 15) & ~15) * 2 + CM_ROWS * 64 * 2 * sizeof(float) +
                    CM_ROWS * 16 * CM_OPITCH;
  if (!(cout == 64 && cin >= 1 && cin * 9 + 2 <= 32 && sm <= 100 * 1024 && (uintptr_t)out_h16 % 16 == 0)) return 1;
  const int64_t items = (int64_t)n * ((h + CM_ROWS - 1) / CM_ROWS);
  const int64_t blocks = items < 148 * 2 ? items : 148 * 2;  // persistent: weights are loaded once per block
  if (wd % 16 == 0) {
    if (sm > 48 * 1024)
      cudaFuncSetAttribute(conv_in_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    launch_k(conv_in_mma_kernel<true>, dim3((unsigned)blocks), dim3(CM_ROWS * 32), sm, st, x, w, b, (__half*)out_h16,
             (long long*)stats, n, cin, h, wd, xscale);
  } else {
    if (sm > 48 * 1024)
      cudaFuncSetAttribute(conv_in_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    launch_k(conv_in_mma_kernel<false>, dim3((unsigned)blocks), dim3(CM_ROWS * 32), sm, st, x, w, b, (__half*)out_h16,
             (long long*)stats, n, cin, h, wd, xscale);
  }
  return 0;
}

int dsg_conv_in_stats(const float* x, const float* w, const float* b, void* out_h16, void* stats, int32_t n,
                      int32_t cin, int32_t h, int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x && w && b && out_h16 && stats, "dsg_conv_in_stats: null pointer");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_in_stats: bad shape");
  if (n == 0) return DSG_OK;
  if (launch_conv_in_mma(x, nullptr, w, b, out_h16, stats, n, cin, h, wd, cout, (cudaStream_t)stream) == 0) {
    DSG_CUDA_LAUNCH_CHECK("dsg_conv_in_stats");
    return DSG_OK;
  }
  // other widths: the CUDA-core kernel, then one read of its output for the statistics
  int rc = dsg_conv_in(x, w, b, out_h16, n, cin, h, wd, cout, stream);
  if (rc != DSG_OK) return rc;
  return dsg_gn_stats(out_h16, cout, stats, n, (int64_t)h * wd, stream);
}

int dsg_conv_out_fused(const void* x_h16, const float* gn_coef, const float* w, const float* b, float* out, int32_t n,
                       int32_t cin, int32_t h, int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x_h16 && w && b && out, "dsg_conv_out_fused: null pointer");
  DSG_CHECK_ARG(cin == 64 && cout >= 1 && cout <= 8, "dsg_conv_out_fused: needs cin == 64 and cout <= 8");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_out_fused: bad shape");
  DSG_CHECK_ARG((uintptr_t)x_h16 % 16 == 0 && (uintptr_t)gn_coef % 8 == 0, "dsg_conv_out_fused: unaligned pointer");
  if (n == 0) return DSG_OK;
  const size_t sm = (size_t)2 * CO_WIN_BYTES;
  static SmemAttrCache attr;
  {
    cudaError_t e = ensure_dyn_smem(attr, conv_out_mma_kernel, (size_t)sm);
    if (e != cudaSuccess) { set_error("dsg_conv_out_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  }
  const int64_t items = (int64_t)n * ((h + CO_ROWS - 1) / CO_ROWS) * ((wd + CO_TW - 1) / CO_TW);
  const int64_t blocks = items < 148 * 2 ? items : 148 * 2;
  launch_k(conv_out_mma_kernel, dim3((unsigned)blocks), dim3(CO_ROWS * 32), sm, (cudaStream_t)stream,
           (const __half*)x_h16, (const float2*)gn_coef, w, b, out, n, h, wd, cout);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_out_fused");
  return DSG_OK;
}

int dsg_conv_out(const void* x_h16, const float* w, const float* b, float* out, int32_t n, int32_t cin, int32_t h,
                 int32_t wd, int32_t cout, void* stream) {
  DSG_CHECK_ARG(x_h16 && w && b && out, "dsg_conv_out: null pointer");
  DSG_CHECK_ARG(cout >= 1 && cout <= 4 && cin % 32 == 0 && cin >= 32 && cin <= 512,
                "dsg_conv_out: cout<=4, cin%%32==0");
  DSG_CHECK_ARG(n >= 0 && h > 0 && wd > 0, "dsg_conv_out: bad shape");
  DSG_CHECK_ARG((uintptr_t)x_h16 % 16 == 0, "dsg_conv_out: x must be 16-byte aligned");
  if (n == 0) return DSG_OK;
  const int64_t total = (int64_t)n * h * wd;
  const size_t sm = (size_t)9 * cout * cin * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(conv_out_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  conv_out_kernel<4><<<(unsigned)ceil_div64(total, CS_THREADS / 4), CS_THREADS, sm, (cudaStream_t)stream>>>(
      (const __half*)x_h16, w, b, out, n, cin, h, wd, cout);
  DSG_CUDA_LAUNCH_CHECK("dsg_conv_out");
  return DSG_OK;
}
}
