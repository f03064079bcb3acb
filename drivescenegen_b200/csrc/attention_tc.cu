// attention_tc.cu — tcgen05 self-attention core for head_dim = 8 (the U-Net mid block: 64 heads x 8, 1024 tokens at
// 256^2, 4096 at 512^2).  Replaces F.scaled_dot_product_attention(q, k, v) inside upstream AttnProcessor2_0
// (diffusers 0.20.0 models/attention_processor.py).
//
// One persistent CTA per SM walks over (sample, head) pairs.  K and V of the head (tokens x 16 B each) sit in shared
// memory in their NATURAL layout, which is already a canonical no-swizzle UMMA operand:
//   K  : [key][8 halfs]  = K-major  core matrices (8 keys x 16 B)            -> B operand of S = Q K^T
//   V  : [key][8 halfs]  = MN-major core matrices (8 keys x 8 channels)      -> B operand of O = P V
// head_dim 8 is half of the f16 MMA K = 16: the second K-chunk of Q and K points (leading-byte-offset) at a block of
// zeros.  V is widened to N = 16 by pointing the second channel group (stride-byte-offset) at a constant block whose
// first column is 1.0, so O[:, 8] = sum_k P — the softmax denominator comes out of the tensor core for free.
// Three warpgroups work on their own 128-query tiles.  Each does two passes over the keys in 64-key blocks, with TWO S
// buffers in TMEM: the MMA of block kb + 2 is issued as soon as block kb has been consumed, so the tensor pipe (and the
// commit -> mbarrier round trip) runs under the softmax arithmetic of block kb + 1 instead of stalling the warpgroup.
//   pass 1  S = Q K^T (tcgen05.mma SS, fp32 in TMEM) -> tcgen05.ld -> exact row max
//   pass 2  S again -> p = exp2((s - max) * scale) -> fp16 P written back over S in TMEM (tcgen05.st)
//           -> O += P V (tcgen05.mma with A = P from TMEM) ... -> O / O[:, 8] -> fp16.
// No online rescaling, nothing but Q/K/V/O touches HBM; the kernel is bound by MUFU.EX2 (tokens^2 * heads exps).
#include "common.cuh"

namespace dsg {

constexpr int ATC_WGS = 3;      // warpgroups per CTA: while one takes row maxima / waits / loads, two others keep MUFU busy
constexpr int ATC_THREADS = ATC_WGS * 128;
constexpr int ATC_KB = 64;   // keys per block (two S buffers of 64 fp32 columns per warpgroup)

__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t start, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((start >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

// dbg (optional, tests): float[128*128 + 128*16]: S of (first pair, tile 0, key block 0) and the raw O of that tile.
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n_pairs, int tokens, int heads,
                    float scale_log2e, float* __restrict__ dbg, float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* Ks = smem;                                 // tokens * 16
  uint8_t* Vs = Ks + (size_t)tokens * 16;             // tokens * 16
  uint8_t* Qs = Vs + (size_t)tokens * 16;             // ATC_WGS warpgroups * 2048
  uint8_t* Vc = Qs + ATC_WGS * 2048;                            // 256: 16 keys x {1, 0 x 7}
  uint8_t* Zs = Vc + 256;                             // 2048 of zeros (second K-chunk of Q and K)
  uint64_t* bars = reinterpret_cast<uint64_t*>(Zs + 2048);  // [warpgroup][S buffer]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * ATC_WGS);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wg = warp >> 2, q = warp & 3;
  const int row = (tid & 127);  // query row within the tile == TMEM lane
  const int C = heads * 8, rs = 3 * C;

  for (int i = tid; i < 2048 / 16; i += ATC_THREADS) reinterpret_cast<uint4*>(Zs)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 16) reinterpret_cast<uint4*>(Vc)[tid] = make_uint4(0x00003C00u, 0, 0, 0);  // half 1.0 in element 0
  if (tid == 0) {
    for (int i = 0; i < 2 * ATC_WGS; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_holder);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  const uint32_t t_s = tmem_base + (uint32_t)(wg * 128);        // 2 x [S (fp32, 64 cols) / P (fp16 pairs, 32 cols)]
  const uint32_t t_o = tmem_base + (uint32_t)(ATC_WGS * 128 + wg * 32);  // O (16 cols)
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const uint32_t idesc_qk = (1u << 4) | ((uint32_t)(ATC_KB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_pv = (1u << 4) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t ks_u = smem_u32(Ks), vs_u = smem_u32(Vs), zs_u = smem_u32(Zs), vc_u = smem_u32(Vc);
  const uint32_t qs_u = smem_u32(Qs) + (uint32_t)(wg * 2048);
  uint64_t* bar = &bars[wg * 2];
  uint32_t ph = 0;   // bit b = phase of bar[b]
  const int nkb = tokens / ATC_KB, ntiles = tokens / 128;
  pdl_sync();

  for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    const int b = pair / heads, h = pair - b * heads;
    const __half* base = qkv + (int64_t)b * tokens * rs + h * 8;
    __syncthreads();  // every MMA of the previous pair has been waited for: K / V may be overwritten
    for (int i0 = tid; i0 < tokens; i0 += ATC_THREADS * 4) {
      uint4 kk[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * ATC_THREADS;
        if (i < tokens) {
          kk[u] = ldg_nc_v4(base + (int64_t)i * rs + C);
          vv[u] = ldg_nc_v4(base + (int64_t)i * rs + 2 * C);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * ATC_THREADS;
        if (i < tokens) {
          reinterpret_cast<uint4*>(Ks)[i] = kk[u];
          reinterpret_cast<uint4*>(Vs)[i] = vv[u];
        }
      }
    }
    fence_async_smem();
    __syncthreads();

    for (int qt = wg; qt < ntiles; qt += ATC_WGS) {
      // ---- Q tile of this warpgroup (its previous MMAs have all completed)
      reinterpret_cast<uint4*>(Qs + wg * 2048)[row] = ldg_nc_v4(base + (int64_t)(qt * 128 + row) * rs);
      fence_async_smem();
      tc_fence_before();
      wg_sync(wg);
      const uint64_t dq = umma_desc_plain(qs_u, zs_u - qs_u, 128);
      // S block `blk` of this tile into S buffer `buf`, completion on bar[buf] (issued by one thread of warp 0)
      auto issue_s = [&](int blk, int buf) {
        const uint32_t ka = ks_u + (uint32_t)(blk * ATC_KB * 16);
        umma_f16(t_s + (uint32_t)(buf * ATC_KB), dq, umma_desc_plain(ka, zs_u - ka, 128), idesc_qk, 0u);
        umma_commit(&bar[buf]);
      };
      if (q == 0) {
        if (elect_one_sync()) {
          tc_fence_after();
          issue_s(0, 0);
          issue_s(1, 1);
        }
        __syncwarp();
      }
      // ---- pass 1: exact row maximum
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four chains: one is a serial latency chain
      for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        mbar_wait(&bar[buf], (ph >> buf) & 1u); ph ^= 1u << buf;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < ATC_KB / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_s + lane_off + (uint32_t)(buf * ATC_KB + c * 32), v);
          tmem_ld_wait();
          if (dbg && pair == 0 && qt == 0 && kb < 128 / ATC_KB) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dbg[row * 128 + kb * ATC_KB + c * 32 + j] = __uint_as_float(v[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            m4[(j >> 1) & 3] = fmaxf(m4[(j >> 1) & 3], fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
        }
        tc_fence_before();
        wg_sync(wg);
        if (q == 0) {
          // block kb + 2 of this pass, or (after the last two) blocks 0 and 1 again for pass 2
          const int nk = kb + 2 < nkb ? kb + 2 : kb + 2 - nkb;
          if (elect_one_sync()) {
            tc_fence_after();
            issue_s(nk, buf);
          }
          __syncwarp();
        }
      }
      // ---- pass 2: P = exp2((S - max) * scale) -> TMEM, O += P V
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float msc = m * scale_log2e;
      for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        mbar_wait(&bar[buf], (ph >> buf) & 1u); ph ^= 1u << buf;
        tc_fence_after();
        const uint32_t t_sb = t_s + (uint32_t)(buf * ATC_KB);
#pragma unroll
        for (int c = 0; c < ATC_KB / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_sb + lane_off + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ex2_fast(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -msc));
            const float p1 = ex2_fast(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -msc));
            const __half2 hh = __floats2half2_rn(p0, p1);
            pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          tmem_st_32x16(t_sb + lane_off + (uint32_t)(c * 16), pk);  // over columns already read
        }
        tmem_st_wait();
        tc_fence_before();
        wg_sync(wg);
        if (q == 0) {
          const uint32_t va = vs_u + (uint32_t)(kb * ATC_KB * 16);
          if (elect_one_sync()) {
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < ATC_KB / 16; ++j) {
              const uint32_t vj = va + (uint32_t)(j * 256);
              umma_f16_ts(t_o, t_sb + (uint32_t)(j * 8), umma_desc_plain(vj, 128, vc_u - vj), idesc_pv,
                          (uint32_t)((kb | j) != 0));
            }
            // the tensor pipe runs in order: S of block kb + 2 may overwrite P of block kb once its P V has been issued
            if (kb + 2 < nkb) issue_s(kb + 2, buf);
            else umma_commit(&bar[buf]);
          }
          __syncwarp();
        }
      }
      // ---- the last two commits (blocks nkb - 2 and nkb - 1) cover every P V of the tile
      mbar_wait(&bar[0], ph & 1u); ph ^= 1u;
      mbar_wait(&bar[1], (ph >> 1) & 1u); ph ^= 2u;
      // ---- O / rowsum -> fp16
      tc_fence_after();
      uint32_t o[16];
      tmem_ld_32x16(t_o + lane_off, o);
      tmem_ld_wait();
      if (dbg && pair == 0 && qt == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dbg[128 * 128 + row * 16 + j] = __uint_as_float(o[j]);
      }
      // training: log2-sum-exp of the scaled scores of this row (the backward recomputes p = 2^(s c - lse))
      if (lse_out) lse_out[(int64_t)pair * tokens + qt * 128 + row] = msc + log2f(__uint_as_float(o[8]));
      const float inv = 1.0f / __uint_as_float(o[8]);
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(o[j]) * inv;
      const int b2 = pair / heads;
      stg_v4(out + ((int64_t)b2 * tokens + qt * 128 + row) * C + (pair - b2 * heads) * 8, pack8(f));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int atc_sms() { return num_sms(); }

// returns DSG_OK, an error, or 1 when the shape is outside the kernel (head_dim != 8, tokens % 128, tokens > 4096)
int launch_attention_tc(const __half* qkv, __half* out, int n, int tokens, int heads, int head_dim, float* dbg,
                        float* lse_out, cudaStream_t st) {
  if (head_dim != 8 || tokens % 128 != 0 || tokens > 4096 || tokens < 128) return 1;
  size_t sm = (size_t)tokens * 32 + ATC_WGS * 2048 + 256 + 2048 + 64 + 128;
  if (sm < 120 * 1024) sm = 120 * 1024;  // one CTA per SM: each allocates all 512 TMEM columns
  static SmemAttrCache attr;
  {
    cudaError_t e = ensure_dyn_smem(attr, attention_tc_kernel, sm);
    if (e != cudaSuccess) { set_error("attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  }
  const int pairs = n * heads;
  const int grid = pairs < atc_sms() ? pairs : atc_sms();
  launch_k(attention_tc_kernel, dim3(grid), dim3(ATC_THREADS), sm, st, qkv, out, pairs, tokens, heads,
           1.4426950408889634f / sqrtf(8.0f), dbg, lse_out);
  DSG_CUDA_LAUNCH_CHECK("dsg_attention/tcgen05");
  return DSG_OK;
}

}  // namespace dsg
