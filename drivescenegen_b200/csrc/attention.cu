// attention.cu — multi-head self-attention core for the U-Net mid block (64 heads x head_dim 8, 1024 tokens at 256^2).
// Replaces F.scaled_dot_product_attention(q, k, v) inside upstream AttnProcessor2_0 (diffusers 0.20.0
// models/attention_processor.py).  v1: flash-style, one query per thread, fp32 math on CUDA
// cores with K/V staged in shared memory; the score matrix never reaches HBM.  exp-throughput bound at d = 8.
#include "common.cuh"

namespace dsg {

constexpr int AT_THREADS = 128;
// keys staged per smem tile (fp32 K and V): 2 * KTILE * D * 4 bytes <= 128 KB
template <int D> struct AtTile { static constexpr int K = D <= 32 ? 512 : 256; };

template <int D>
__global__ void __launch_bounds__(AT_THREADS) attention_kernel(const __half* __restrict__ qkv,
                                                               __half* __restrict__ out, int tokens, int heads,
                                                               float scale_log2e) {
  constexpr int AT_KTILE = AtTile<D>::K;
  extern __shared__ float smf[];
  float* sk = smf;                 // [AT_KTILE][D]
  float* sv = smf + AT_KTILE * D;  // [AT_KTILE][D]
  const int n = blockIdx.z, hd = blockIdx.y;
  const int c = heads * D, row = 3 * c;
  const int qi = blockIdx.x * AT_THREADS + threadIdx.x;
  const bool valid = qi < tokens;
  const __half* base = qkv + (int64_t)n * tokens * row + hd * D;
  float q[D], o[D];
#pragma unroll
  for (int j = 0; j < D; ++j) o[j] = 0.f;
  {
    const __half* qp = base + (int64_t)(valid ? qi : 0) * row;
#pragma unroll
    for (int j = 0; j < D; j += 8) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(qp + j), f);
#pragma unroll
      for (int u = 0; u < 8; ++u) q[j + u] = f[u] * scale_log2e;
    }
  }
  float m = -INFINITY, l = 0.f;
  constexpr int VPR = D / 8;  // uint4 per row slice
  for (int k0 = 0; k0 < tokens; k0 += AT_KTILE) {
    const int nk = min(AT_KTILE, tokens - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * VPR; i += AT_THREADS) {
      const int kr = i / VPR, kv = i - kr * VPR;
      const __half* kp = base + (int64_t)(k0 + kr) * row + c + kv * 8;
      float f[8];
      unpack8(ldg_nc_v4(kp), f);
      *reinterpret_cast<float4*>(sk + kr * D + kv * 8) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(sk + kr * D + kv * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
      unpack8(ldg_nc_v4(kp + c), f);
      *reinterpret_cast<float4*>(sv + kr * D + kv * 8) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(sv + kr * D + kv * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    for (int kb = 0; kb < nk; kb += 8) {
      float s[8];
      float mx = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (kb + u < nk) {
          const float* kr = sk + (kb + u) * D;
          float a = 0.f;
#pragma unroll
          for (int j = 0; j < D; j += 4) {
            const float4 kk = *reinterpret_cast<const float4*>(kr + j);
            a = fmaf(q[j], kk.x, a); a = fmaf(q[j + 1], kk.y, a);
            a = fmaf(q[j + 2], kk.z, a); a = fmaf(q[j + 3], kk.w, a);
          }
          s[u] = a;
          mx = fmaxf(mx, a);
        } else {
          s[u] = -INFINITY;
        }
      }
      if (mx > m) {
        const float corr = exp2f(m - mx);  // m = -inf on the first block -> 0
        l *= corr;
#pragma unroll
        for (int j = 0; j < D; ++j) o[j] *= corr;
        m = mx;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float p = exp2f(s[u] - m);  // masked keys: exp2(-inf) = 0
        l += p;
        const float* vr = sv + (kb + u < nk ? kb + u : 0) * D;
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(vr + j);
          o[j] = fmaf(p, vv.x, o[j]); o[j + 1] = fmaf(p, vv.y, o[j + 1]);
          o[j + 2] = fmaf(p, vv.z, o[j + 2]); o[j + 3] = fmaf(p, vv.w, o[j + 3]);
        }
      }
    }
  }
  if (valid) {
    const float inv = 1.0f / l;
    __half* op = out + ((int64_t)n * tokens + qi) * c + hd * D;
#pragma unroll
    for (int j = 0; j < D; j += 8) {
      float f[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = o[j + u] * inv;
      stg_v4(op + j, pack8(f));
    }
  }
}

template <int D>
static int launch_attention(const __half* qkv, __half* out, int n, int tokens, int heads, cudaStream_t st) {
  const size_t sm = (size_t)2 * AtTile<D>::K * D * sizeof(float);
  if (sm > 48 * 1024)
    cudaFuncSetAttribute(attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)D);
  attention_kernel<D><<<dim3(ceil_div(tokens, AT_THREADS), heads, n), AT_THREADS, sm, st>>>(qkv, out, tokens, heads,
                                                                                          scale_log2e);
  DSG_CUDA_LAUNCH_CHECK("dsg_attention");
  return DSG_OK;
}

}  // namespace dsg

using namespace dsg;

namespace dsg {
int launch_attention_tc(const __half* qkv, __half* out, int n, int tokens, int heads, int head_dim, float* dbg,
                        float* lse_out, cudaStream_t st);  // attention_tc.cu
}

extern "C" int dsg_attention_train_tc_ok(int32_t tokens, int32_t head_dim) {
  return (head_dim == 8 && tokens % 128 == 0 && tokens >= 128 && tokens <= 2048) ? 1 : 0;
}

extern "C" int dsg_attention_train(const void* qkv, void* out, float* lse, int32_t n, int32_t tokens, int32_t heads,
                                   int32_t head_dim, void* stream) {
  DSG_CHECK_ARG(qkv && out && lse, "dsg_attention_train: null pointer");
  DSG_CHECK_ARG(dsg_attention_train_tc_ok(tokens, head_dim), "dsg_attention_train: needs head_dim 8 and tokens a "
                "multiple of 128 in [128, 2048] (use dsg_attention + dsg_attention_bwd with lse = NULL otherwise)");
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && heads > 0 && heads <= 65535, "dsg_attention_train: bad shape");
  DSG_CHECK_ARG(((uintptr_t)qkv | (uintptr_t)out) % 16 == 0, "dsg_attention_train: pointers must be 16-byte aligned");
  if (n == 0) return DSG_OK;
  const int rc = launch_attention_tc((const __half*)qkv, (__half*)out, n, tokens, heads, head_dim, nullptr, lse,
                                     (cudaStream_t)stream);
  if (rc > 0) { dsg::set_error("dsg_attention_train: shape outside the tcgen05 kernel"); return DSG_ERR_UNSUPPORTED; }
  return rc;
}

extern "C" int dsg_attention_ex(const void* qkv, void* out, int32_t n, int32_t tokens, int32_t heads,
                                int32_t head_dim, int32_t impl, float* dbg, void* stream) {
  DSG_CHECK_ARG(qkv && out, "dsg_attention: null pointer");
  DSG_CHECK_ARG(n >= 0 && n <= 65535 && tokens > 0 && heads > 0 && heads <= 65535, "dsg_attention: bad shape");
  DSG_CHECK_ARG(((uintptr_t)qkv | (uintptr_t)out) % 16 == 0, "dsg_attention: pointers must be 16-byte aligned");
  DSG_CHECK_ARG(impl >= 0 && impl <= 2, "dsg_attention: bad impl %d", impl);
  if (n == 0) return DSG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl != 1) {
    const int rc = launch_attention_tc((const __half*)qkv, (__half*)out, n, tokens, heads, head_dim, dbg, nullptr, st);
    if (rc <= 0) return rc;
    if (impl == 2) {
      dsg::set_error("dsg_attention: the tcgen05 kernel needs head_dim 8 and 128 <= tokens <= 4096, tokens %% 128 == 0");
      return DSG_ERR_UNSUPPORTED;
    }
  }
  switch (head_dim) {
    case 8: return launch_attention<8>((const __half*)qkv, (__half*)out, n, tokens, heads, st);
    case 16: return launch_attention<16>((const __half*)qkv, (__half*)out, n, tokens, heads, st);
    case 32: return launch_attention<32>((const __half*)qkv, (__half*)out, n, tokens, heads, st);
    case 64: return launch_attention<64>((const __half*)qkv, (__half*)out, n, tokens, heads, st);
    default:
      dsg::set_error("dsg_attention: head_dim %d unsupported (8/16/32/64)", head_dim);
      return DSG_ERR_UNSUPPORTED;
  }
}

extern "C" int dsg_attention(const void* qkv, void* out, int32_t n, int32_t tokens, int32_t heads, int32_t head_dim,
                             void* stream) {
  return dsg_attention_ex(qkv, out, n, tokens, heads, head_dim, 0, nullptr, stream);
}
