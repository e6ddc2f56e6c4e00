// Statistics pooling over time (StatisticsPooling.forward, pytorch/libs/nnet/pooling.py:58-67).
//
//   mean[b,c] = (1/T) sum_t x[b,t,c];  std[b,c] = sqrt(max((1/T) sum_t (x-mean)^2, eps))
//
// HBM-bound: x (B,T,C) fp32 is read exactly once.  A CTA owns one utterance x one 128-channel
// slab; each of its 8 warps streams every 8th frame (32 lanes x float4 = 512 contiguous bytes per
// frame) in register chunks of kRows frames.  Within a chunk the variance is the true two-pass
// sum (x - chunk_mean)^2 on registers; chunks and warps are merged with Chan's parallel update,
// which is algebraically the reference's two-pass result without a second trip to memory.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace xvb {

constexpr int kPoolWarps = 8;
constexpr int kPoolRows = 8;  // frames held in registers per warp per chunk (8 x float4 = 32 regs)

struct Moments {  // running count / mean / M2 for 4 channels
  float n;
  float mean[4];
  float m2[4];
};

__device__ __forceinline__ void chan_merge(Moments& a, float nb, const float (&mb)[4], const float (&m2b)[4]) {
  if (nb == 0.f) return;
  const float tot = a.n + nb;
  const float wb = nb / tot;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float d = mb[k] - a.mean[k];
    a.mean[k] = fmaf(d, wb, a.mean[k]);
    a.m2[k] += m2b[k] + d * d * a.n * wb;
  }
  a.n = tot;
}

__global__ void __launch_bounds__(kPoolWarps * 32)
stats_pool_kernel(const float* __restrict__ x, long long ldx, int T, int C, float eps, float* __restrict__ out,
                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ldo) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4;
  const int warp = threadIdx.x >> 5;
  const bool active = c < C;  // C % 4 == 0, so a float4 is all-in or all-out
  const float* xb = x + (long long)b * T * ldx + c;

  Moments acc{};
  for (int t0 = warp; t0 < T; t0 += kPoolWarps * kPoolRows) {
    float4 v[kPoolRows];
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < kPoolRows; ++r) {
      const int t = t0 + r * kPoolWarps;
      if (t < T) {
        if (active) v[r] = __ldcs(reinterpret_cast<const float4*>(xb + (long long)t * ldx));  // streaming: read once
        ++cnt;
      }
    }
    if (!active) continue;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < kPoolRows; ++r)
      if (r < cnt) { s[0] += v[r].x; s[1] += v[r].y; s[2] += v[r].z; s[3] += v[r].w; }
    const float inv = 1.f / (float)cnt;
    float m[4] = {s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv};
    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < kPoolRows; ++r)
      if (r < cnt) {
        float d;
        d = v[r].x - m[0]; q[0] = fmaf(d, d, q[0]);
        d = v[r].y - m[1]; q[1] = fmaf(d, d, q[1]);
        d = v[r].z - m[2]; q[2] = fmaf(d, d, q[2]);
        d = v[r].w - m[3]; q[3] = fmaf(d, d, q[3]);
      }
    chan_merge(acc, (float)cnt, m, q);
  }

  __shared__ float sh_n[kPoolWarps][32];
  __shared__ float sh_mean[kPoolWarps][32][4];
  __shared__ float sh_m2[kPoolWarps][32][4];
  const int lane = threadIdx.x & 31;
  sh_n[warp][lane] = acc.n;
#pragma unroll
  for (int k = 0; k < 4; ++k) { sh_mean[warp][lane][k] = acc.mean[k]; sh_m2[warp][lane][k] = acc.m2[k]; }
  __syncthreads();
  if (warp == 0 && active) {
    Moments tot{};
    for (int w = 0; w < kPoolWarps; ++w) {
      float mb[4], qb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { mb[k] = sh_mean[w][lane][k]; qb[k] = sh_m2[w][lane][k]; }
      chan_merge(tot, sh_n[w][lane], mb, qb);
    }
    float sd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sd[k] = sqrtf(fmaxf(tot.m2[k] / (float)T, eps));  // biased var, clamp(min=eps)
    float* ob = out + (long long)b * 2 * C;
    *reinterpret_cast<float4*>(ob + c) = make_float4(tot.mean[0], tot.mean[1], tot.mean[2], tot.mean[3]);
    *reinterpret_cast<float4*>(ob + C + c) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    if (out_hi) {
      __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) { split_bf16(tot.mean[k], h[k], l[k]); split_bf16(sd[k], h[4 + k], l[4 + k]); }
      __nv_bfloat16* oh = out_hi + (long long)b * ldo;
      __nv_bfloat16* ol = out_lo + (long long)b * ldo;
      *reinterpret_cast<uint2*>(oh + c) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
      *reinterpret_cast<uint2*>(ol + c) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      *reinterpret_cast<uint2*>(oh + C + c) = make_uint2(pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
      *reinterpret_cast<uint2*>(ol + C + c) = make_uint2(pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    }
  }
}

}  // namespace xvb

using namespace xvb;

extern "C" int xvb_stats_pool(const float* x, int64_t ldx, int B, int T, int C, float eps, float* out, uint16_t* out_hi,
                              uint16_t* out_lo, int64_t ldo, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && out, "xvb_stats_pool: null pointer");
  XVB_CHECK_ARG(B > 0 && T > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldx >= C, "xvb_stats_pool: need C%%4==0, ldx%%4==0 (C=%d ldx=%lld)", C, (long long)ldx);
  XVB_CHECK_ARG((out_hi != nullptr) == (out_lo != nullptr), "xvb_stats_pool: out_hi/out_lo must both be set or both NULL");
  if (out_hi) XVB_CHECK_ARG(ldo >= 2 * C && ldo % 4 == 0, "xvb_stats_pool: ldo=%lld too small / unaligned", (long long)ldo);
  XVB_CHECK_ARG(((uintptr_t)x | (uintptr_t)out) % 16 == 0 && ((uintptr_t)out_hi | (uintptr_t)out_lo) % 8 == 0, "xvb_stats_pool: unaligned pointer");
  XVB_CHECK_ARG(B <= 65535, "xvb_stats_pool: B=%d exceeds grid.y", B);
  dim3 grid((C + 127) / 128, B);
  stats_pool_kernel<<<grid, kPoolWarps * 32, 0, (cudaStream_t)stream>>>(
      x, ldx, T, C, eps, out, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), ldo);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}
