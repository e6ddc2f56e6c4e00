// Feature-side front-end on the GPU: what sits immediately before the
// extractor in the reference's pipelines, for a ragged batch of utterances stored back to back as
// one (sum_T, F) fp32 matrix with an offsets array (U+1).
//   * energy VAD           : TorchAsvExtractor::ComputeVadEnergy, runtime/extractor/torch_asv_extractor.cc:14-62
//   * per-utterance CMN    : `input_feats - input_feats.mean(0)`, torch_asv_extractor.cc:99-101
//   * sliding-window CMN   : Kaldi apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=W as
//                            called by pytorch/pipeline/extract_xvectors_for_pytorch.sh:105-111
//                            (Kaldi is not vendored: semantics restated, parity unpinned)
//   * voiced-frame select  : index_select(0, nonzero(vad)) (:103-107) / Kaldi select-voiced-frames
// None of this is on the throughput-critical path (bytes: 4*F per frame); the kernels are simple
// one-CTA-per-utterance loops.
#include <cuda_runtime.h>

#include "common.cuh"

namespace xvb {

__global__ void vad_energy_kernel(const float* __restrict__ x, const int32_t* __restrict__ off, int F, float threshold,
                                  float mean_scale, int context, float proportion, uint8_t* __restrict__ voiced,
                                  int32_t* __restrict__ counts) {
  const int u = blockIdx.x;
  const int beg = off[u], T = off[u + 1] - off[u];
  __shared__ float red[32];
  __shared__ float thr_s;
  __shared__ int cnt_s;
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) s += x[(long long)(beg + t) * F];   // column 0 = log-energy
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  if (threadIdx.x == 0) cnt_s = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    thr_s = threshold + (mean_scale != 0.f ? mean_scale * tot / (float)T : 0.f);
  }
  __syncthreads();
  const float thr = thr_s;
  int local = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int num = 0, den = 0;
    for (int t2 = t - context; t2 <= t + context; ++t2)
      if (t2 >= 0 && t2 < T) { ++den; if (x[(long long)(beg + t2) * F] > thr) ++num; }
    const uint8_t v = (float)num >= (float)den * proportion ? 1 : 0;
    voiced[beg + t] = v;
    local += v;
  }
  atomicAdd(&cnt_s, local);
  __syncthreads();
  if (threadIdx.x == 0) counts[u] = cnt_s;
}

// y = x - mean over the window of each frame.  window <= 0: whole utterance (per-utterance CMN).
// Otherwise Kaldi's centred sliding window: [t - W/2, t - W/2 + W) shifted to stay inside [0, T).
__global__ void cmn_kernel(const float* __restrict__ x, const int32_t* __restrict__ off, int F, int window,
                           float* __restrict__ y) {
  const int u = blockIdx.x;
  const int beg = off[u], T = off[u + 1] - off[u];
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float* xc = x + (long long)beg * F + f;
    float* yc = y + (long long)beg * F + f;
    if (window <= 0 || window >= T) {
      double s = 0.0;
      for (int t = 0; t < T; ++t) s += (double)xc[(long long)t * F];
      const float m = (float)(s / (double)T);
      for (int t = 0; t < T; ++t) yc[(long long)t * F] = xc[(long long)t * F] - m;
    } else {
      double s = 0.0;
      int wb = 0, we = 0;  // current window [wb, we)
      for (int t = 0; t < T; ++t) {
        int b = t - window / 2, e = b + window;
        if (b < 0) { e -= b; b = 0; }
        if (e > T) { b -= (e - T); e = T; if (b < 0) b = 0; }
        while (we < e) { s += (double)xc[(long long)we * F]; ++we; }
        while (wb < b) { s -= (double)xc[(long long)wb * F]; ++wb; }
        yc[(long long)t * F] = xc[(long long)t * F] - (float)(s / (double)(e - b));
      }
    }
  }
}

// ordered compaction of the voiced frames of each utterance: one warp per utterance
__global__ void select_frames_kernel(const float* __restrict__ x, const int32_t* __restrict__ off,
                                     const uint8_t* __restrict__ voiced, const int32_t* __restrict__ out_off, int F,
                                     float* __restrict__ y) {
  const int u = blockIdx.x;
  const int beg = off[u], T = off[u + 1] - off[u];
  const int lane = threadIdx.x;
  int written = out_off[u];
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool v = t < T && voiced[beg + t] != 0;
    const unsigned mask = __ballot_sync(0xffffffffu, v);
    const int pos = written + __popc(mask & ((1u << lane) - 1));
    if (v) {
      const float* src = x + (long long)(beg + t) * F;
      float* dst = y + (long long)pos * F;
      for (int f = 0; f < F; ++f) dst[f] = src[f];
    }
    written += __popc(mask);
  }
}

}  // namespace xvb

using namespace xvb;

extern "C" int xvb_vad_energy(const float* x, const int32_t* offsets, int num_utts, int F, float energy_threshold,
                              float energy_mean_scale, int frames_context, float proportion_threshold, uint8_t* voiced,
                              int32_t* voiced_counts, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && offsets && voiced && voiced_counts && num_utts > 0 && F > 0, "xvb_vad_energy: bad arguments");
  XVB_CHECK_ARG(frames_context >= 0 && proportion_threshold > 0.f && proportion_threshold < 1.f && energy_mean_scale >= 0.f,
                "xvb_vad_energy: options out of range (torch_asv_extractor.cc:34-41)");
  vad_energy_kernel<<<num_utts, 256, 0, (cudaStream_t)stream>>>(x, offsets, F, energy_threshold, energy_mean_scale,
                                                                frames_context, proportion_threshold, voiced, voiced_counts);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_cmn(const float* x, const int32_t* offsets, int num_utts, int F, int window, float* y, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && offsets && y && num_utts > 0 && F > 0, "xvb_cmn: bad arguments");
  cmn_kernel<<<num_utts, F < 128 ? ((F + 31) / 32) * 32 : 128, 0, (cudaStream_t)stream>>>(x, offsets, F, window, y);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_select_frames(const float* x, const int32_t* offsets, const uint8_t* voiced, const int32_t* out_offsets,
                                 int num_utts, int F, float* y, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && offsets && voiced && out_offsets && y && num_utts > 0 && F > 0, "xvb_select_frames: bad arguments");
  select_frames_kernel<<<num_utts, 32, 0, (cudaStream_t)stream>>>(x, offsets, voiced, out_offsets, F, y);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}
