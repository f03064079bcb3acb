// attention_bwd_tc.cu — tcgen05 backward of the head_dim-8 self-attention core (training path).
// Same building blocks as the forward kernel (attention_tc.cu): Q, K, V, dO of one (sample, head) sit in shared memory
// in their natural [token][8 halfs] layout, which is at once a K-major operand (scores) and an MN-major operand
// (the "times V"-style products); the 8-deep contraction is padded to the MMA K = 16 with a block of zeros, and the
// probability / score-gradient tiles go back into TMEM as fp16 A operands.  One persistent CTA per SM walks over
// (sample, head) pairs; two warpgroups ping-pong over 128-row tiles.  Two passes per pair:
//   dQ pass    rows = queries:  S = Q K^T, dP = dO V^T (SS MMAs, 64 keys per block) -> p = 2^(s c - lse_row),
//              dS = p (dP - delta_row) / sqrt(d) -> fp16 in TMEM -> dQ += dS K   (TS MMA, K as MN-major operand)
//   dK/dV pass rows = keys:     S^T = K Q^T, dP^T = V dO^T (64 queries per block) -> p = 2^(s c - lse_col),
//              P^T, dS^T -> fp16 in TMEM -> dV += P^T dO, dK += dS^T Q        (TS MMAs)
// lse comes from the forward kernel (log2 units), delta_i = <dO_i, O_i> is computed while dO is staged.
// exp work: 2 x tokens^2 x heads per sample (the probabilities are recomputed in both orientations); the CUDA-core
// version it replaces (attention_bwd.cu) spent 64 FMAs + 3 exps per (query, key) pair.
// dS is carried times 2^10 in fp16 (its magnitude is ~1e-3 of dO's) and the factor is removed from dQ / dK.
#include "common.cuh"

namespace dsg {

constexpr int ABT_THREADS = 256;
constexpr int ABT_KB = 64;  // "other side" rows per block

// helpers shared with attention_tc.cu (kept local: both files are self-contained translation units)
__device__ __forceinline__ uint64_t abt_desc_plain(uint32_t start, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((start >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void abt_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void abt_tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void abt_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void abt_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float abt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void abt_wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

constexpr float ABT_DS_SCALE = 1024.0f;

__global__ void __launch_bounds__(ABT_THREADS, 1)
attention_bwd_tc_kernel(const __half* __restrict__ qkv, const __half* __restrict__ o, const __half* __restrict__ dout,
                        const float* __restrict__ lse, __half* __restrict__ dqkv, int n_pairs, int tokens, int heads,
                        float scale, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* Qs = smem;                                  // tokens * 16 each
  uint8_t* Ks = Qs + (size_t)tokens * 16;
  uint8_t* Vs = Ks + (size_t)tokens * 16;
  uint8_t* Ds = Vs + (size_t)tokens * 16;              // dO
  uint8_t* Zs = Ds + (size_t)tokens * 16;              // 2048 of zeros
  float* s_lse = reinterpret_cast<float*>(Zs + 2048);  // [tokens]
  float* s_del = s_lse + tokens;                       // [tokens]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_del + tokens);  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wg = warp >> 2, q = warp & 3;
  const int row = tid & 127;
  const int C = heads * 8, rs = 3 * C;

  for (int i = tid; i < 2048 / 16; i += ABT_THREADS) reinterpret_cast<uint4*>(Zs)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(tmem_holder);
  abt_fence_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
  const uint32_t t_s = tmem_base + (uint32_t)(wg * 256);    // S (fp32, 64 cols) -> P (fp16, 32 cols)
  const uint32_t t_dp = t_s + 64u;                           // dP (fp32, 64 cols) -> dS (fp16, 32 cols)
  const uint32_t t_a0 = t_s + 128u;                          // accumulator 0 (16 cols): dQ, or dV
  const uint32_t t_a1 = t_s + 160u;                          // accumulator 1 (16 cols): dK
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const uint32_t idesc_sc = (1u << 4) | ((uint32_t)(ABT_KB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_ts = (1u << 4) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t qs_u = smem_u32(Qs), ks_u = smem_u32(Ks), vs_u = smem_u32(Vs), ds_u = smem_u32(Ds),
                 zs_u = smem_u32(Zs);
  uint64_t* bar = &bars[wg];
  uint32_t ph = 0;
  const int nblk = tokens / ABT_KB, ntiles = tokens / 128;
  pdl_sync();

  for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    const int b = pair / heads, h = pair - b * heads;
    const __half* base = qkv + (int64_t)b * tokens * rs + h * 8;
    const __half* obase = o + (int64_t)b * tokens * C + h * 8;
    const __half* dbase = dout + (int64_t)b * tokens * C + h * 8;
    const float* lbase = lse + ((int64_t)b * heads + h) * tokens;
    __half* gbase = dqkv + (int64_t)b * tokens * rs + h * 8;
    __syncthreads();  // every MMA of the previous pair has been waited for: the staged operands may be overwritten
    for (int i = tid; i < tokens; i += ABT_THREADS) {
      const uint4 qq = ldg_nc_v4(base + (int64_t)i * rs);
      const uint4 kk = ldg_nc_v4(base + (int64_t)i * rs + C);
      const uint4 vv = ldg_nc_v4(base + (int64_t)i * rs + 2 * C);
      const uint4 dd = ldg_nc_v4(dbase + (int64_t)i * C);
      const uint4 oo = ldg_nc_v4(obase + (int64_t)i * C);
      reinterpret_cast<uint4*>(Qs)[i] = qq;
      reinterpret_cast<uint4*>(Ks)[i] = kk;
      reinterpret_cast<uint4*>(Vs)[i] = vv;
      reinterpret_cast<uint4*>(Ds)[i] = dd;
      float fo[8], fd[8];
      unpack8(oo, fo); unpack8(dd, fd);
      float dl = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) dl = fmaf(fo[j], fd[j], dl);
      s_del[i] = dl;
      s_lse[i] = lbase[i];
    }
    abt_fence_async();
    __syncthreads();

    // two passes: pass 0 = dQ (rows are queries), pass 1 = dK / dV (rows are keys)
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t a_s = pass == 0 ? qs_u : ks_u;   // A of the score product   (row side)
      const uint32_t a_d = pass == 0 ? ds_u : vs_u;   // A of the dP product
      const uint32_t b_s = pass == 0 ? ks_u : qs_u;   // B of the score product   (column side)
      const uint32_t b_d = pass == 0 ? vs_u : ds_u;   // B of the dP product
      const uint32_t m0 = pass == 0 ? ks_u : ds_u;    // MN-major B of accumulator 0: dQ += dS K   |  dV += P^T dO
      const uint32_t m1 = qs_u;                       // MN-major B of accumulator 1:                 dK += dS^T Q
      for (int tile = wg; tile < ntiles; tile += 2) {
        const int r_glob = tile * 128 + row;
        const uint32_t ta_s = a_s + (uint32_t)(tile * 2048), ta_d = a_d + (uint32_t)(tile * 2048);
        const uint64_t das = abt_desc_plain(ta_s, zs_u - ta_s, 128), dad = abt_desc_plain(ta_d, zs_u - ta_d, 128);
        const float lse_r = s_lse[r_glob], del_r = s_del[r_glob];
        tc_fence_before();
        abt_wg_sync(wg);   // the warpgroup's previous tile has been read out of TMEM
        if (q == 0) {
          if (elect_one_sync()) {
            tc_fence_after();
            umma_f16(t_s, das, abt_desc_plain(b_s, zs_u - b_s, 128), idesc_sc, 0u);
            umma_f16(t_dp, dad, abt_desc_plain(b_d, zs_u - b_d, 128), idesc_sc, 0u);
            umma_commit(bar);
          }
          __syncwarp();
        }
        for (int kb = 0; kb < nblk; ++kb) {
          mbar_wait(bar, ph); ph ^= 1;
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(t_s + lane_off + (uint32_t)(c * 32), sv);
            tmem_ld_32x32(t_dp + lane_off + (uint32_t)(c * 32), dv);
            tmem_ld_wait();
            uint32_t pk[16], dk[16];
            const int col0 = kb * ABT_KB + c * 32;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float l0 = lse_r, l1 = lse_r, e0 = del_r, e1 = del_r;
              if (pass == 1) {
                l0 = s_lse[col0 + 2 * j]; l1 = s_lse[col0 + 2 * j + 1];
                e0 = s_del[col0 + 2 * j]; e1 = s_del[col0 + 2 * j + 1];
              }
              const float p0 = abt_ex2(fmaf(__uint_as_float(sv[2 * j]), scale_log2e, -l0));
              const float p1 = abt_ex2(fmaf(__uint_as_float(sv[2 * j + 1]), scale_log2e, -l1));
              const float g0 = p0 * (__uint_as_float(dv[2 * j]) - e0) * (scale * ABT_DS_SCALE);
              const float g1 = p1 * (__uint_as_float(dv[2 * j + 1]) - e1) * (scale * ABT_DS_SCALE);
              const __half2 hp = __floats2half2_rn(p0, p1), hg = __floats2half2_rn(g0, g1);
              pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
              dk[j] = *reinterpret_cast<const uint32_t*>(&hg);
            }
            if (pass == 1) abt_tmem_st16(t_s + lane_off + (uint32_t)(c * 16), pk);
            abt_tmem_st16(t_dp + lane_off + (uint32_t)(c * 16), dk);
          }
          abt_st_wait();
          tc_fence_before();
          abt_wg_sync(wg);
          if (q == 0) {
            const uint32_t blk = (uint32_t)(kb * ABT_KB * 16);
            const bool more = kb + 1 < nblk;
            if (elect_one_sync()) {
              tc_fence_after();
#pragma unroll
              for (int j = 0; j < ABT_KB / 16; ++j) {
                const uint32_t o0 = m0 + blk + (uint32_t)(j * 256), o1 = m1 + blk + (uint32_t)(j * 256);
                if (pass == 0) {
                  abt_umma_ts(t_a0, t_dp + (uint32_t)(j * 8), abt_desc_plain(o0, 128, zs_u - o0), idesc_ts,
                              (uint32_t)((kb | j) != 0));
                } else {
                  abt_umma_ts(t_a0, t_s + (uint32_t)(j * 8), abt_desc_plain(o0, 128, zs_u - o0), idesc_ts,
                              (uint32_t)((kb | j) != 0));
                  abt_umma_ts(t_a1, t_dp + (uint32_t)(j * 8), abt_desc_plain(o1, 128, zs_u - o1), idesc_ts,
                              (uint32_t)((kb | j) != 0));
                }
              }
              if (more) {
                const uint32_t nb_s = b_s + blk + (uint32_t)(ABT_KB * 16), nb_d = b_d + blk + (uint32_t)(ABT_KB * 16);
                umma_f16(t_s, das, abt_desc_plain(nb_s, zs_u - nb_s, 128), idesc_sc, 0u);
                umma_f16(t_dp, dad, abt_desc_plain(nb_d, zs_u - nb_d, 128), idesc_sc, 0u);
              }
              umma_commit(bar);
            }
            __syncwarp();
          }
        }
        // ---- accumulators -> fp16 gradients
        mbar_wait(bar, ph); ph ^= 1;
        tc_fence_after();
        uint32_t a0[16];
        tmem_ld_32x16(t_a0 + lane_off, a0);
        tmem_ld_wait();
        float f[8];
        if (pass == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(a0[j]) * (1.0f / ABT_DS_SCALE);
          stg_v4(gbase + (int64_t)r_glob * rs, pack8(f));                       // dQ
        } else {
          uint32_t a1[16];
          tmem_ld_32x16(t_a1 + lane_off, a1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(a0[j]);
          stg_v4(gbase + (int64_t)r_glob * rs + 2 * C, pack8(f));               // dV
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(a1[j]) * (1.0f / ABT_DS_SCALE);
          stg_v4(gbase + (int64_t)r_glob * rs + C, pack8(f));                   // dK
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// returns DSG_OK, an error, or 1 when the shape is outside the kernel
int launch_attention_bwd_tc(const __half* qkv, const __half* o, const __half* dout, const float* lse, __half* dqkv, int n,
                            int tokens, int heads, int head_dim, cudaStream_t st) {
  if (head_dim != 8 || tokens % 128 != 0 || tokens > 2048 || tokens < 128) return 1;
  size_t sm = (size_t)tokens * 64 + 2048 + (size_t)tokens * 8 + 64 + 128;
  if (sm < 120 * 1024) sm = 120 * 1024;  // one CTA per SM: each allocates all 512 TMEM columns
  static SmemAttrCache attr;
  {
    cudaError_t e = ensure_dyn_smem(attr, attention_bwd_tc_kernel, sm);
    if (e != cudaSuccess) { set_error("attention_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DSG_ERR_CUDA; }
  }
  const int sms = num_sms();
  const int pairs = n * heads;
  const int grid = pairs < sms ? pairs : sms;
  const float scale = 1.0f / sqrtf(8.0f);
  launch_k(attention_bwd_tc_kernel, dim3(grid), dim3(ABT_THREADS), sm, st, qkv, o, dout, lse, dqkv, pairs, tokens, heads,
           scale, scale * 1.4426950408889634f);
  DSG_CUDA_LAUNCH_CHECK("dsg_attention_bwd/tcgen05");
  return DSG_OK;
}

}  // namespace dsg
