// Back-end scoring kernels: embedding pre-processing (score/process.sh `submean`/`norm`/`getmean`),
// cosine (score/score.sh:82-97) and two-covariance PLDA (score/pyplda/gaussian-plda-scoring.py:23-50).
// The score matrices are dense contractions and ride on the tcgen05 layer kernel of tdnn_gemm.cu
// (enroll rows = "frames" with T=1, test rows = output channels); the per-row pieces are
// bandwidth-bound warp kernels.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.cuh"

namespace xvb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row: y = (x - mean) / ||x - mean||
__global__ void center_length_norm_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                          float* __restrict__ y, long long rows, int D) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = xr[c] - (mean ? mean[c] : 0.f);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / sqrtf(ss);
  for (int c = lane; c < D; c += 32) y[row * D + c] = (xr[c] - (mean ? mean[c] : 0.f)) * inv;
}

// partial[g, c] = sum over rows r = g, g+G, ... of x[r, c]   (thread = column, coalesced)
__global__ void column_partial_kernel(const float* __restrict__ x, long long rows, int D, double* __restrict__ partial) {
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    double s = 0.0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) s += (double)x[r * D + c];
    partial[(long long)blockIdx.x * D + c] = s;
  }
}
__global__ void column_final_kernel(const double* __restrict__ partial, int G, long long rows, int D,
                                    float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  double s = 0.0;
  for (int g = 0; g < G; ++g) s += partial[(long long)g * D + c];
  mean[c] = (float)(s / (double)rows);
}

// one CTA per speaker: mean of its member rows (CSR lists), thread = column
__global__ void speaker_mean_kernel(const float* __restrict__ x, int D, const int32_t* __restrict__ offsets,
                                    const int32_t* __restrict__ members, float* __restrict__ out) {
  const int s = blockIdx.x;
  const int beg = offsets[s], end = offsets[s + 1];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    double acc = 0.0;
    for (int i = beg; i < end; ++i) acc += (double)x[(long long)members[i] * D + c];
    out[(long long)s * D + c] = end > beg ? (float)(acc / (double)(end - beg)) : 0.f;
  }
}

// one warp per trial: <e[te], t[tt]> (+ row[te] + col[tt] for PLDA)
__global__ void cosine_trials_kernel(const float* __restrict__ e, const float* __restrict__ t, int D,
                                     const int32_t* __restrict__ te, const int32_t* __restrict__ tt,
                                     const float* __restrict__ row, const float* __restrict__ col,
                                     long long n, float* __restrict__ scores) {
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* a = e + (long long)te[i] * D;
  const float* b = t + (long long)tt[i] * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s = fmaf(a[c], b[c], s);
  s = warp_sum(s);
  if (lane == 0) scores[i] = s + (row ? row[te[i]] : 0.f) + (col ? col[tt[i]] : 0.f);
}

// term[i] = <y_i, x_i> + <x_i, c>   (y = x G computed by the GEMM)
__global__ void plda_rowterm_kernel(const float* __restrict__ x, const float* __restrict__ y, long long ldy,
                                    const float* __restrict__ c, long long rows, int D, float* __restrict__ term) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float xv = x[row * D + k];
    s = fmaf(xv, y[row * ldy + k] + c[k], s);
  }
  s = warp_sum(s);
  if (lane == 0) term[row] = s;
}

// ---------------------------------------------------------------- score normalisation (S-norm / AS-norm)
// One CTA per row of a cohort score matrix: bitonic sort (descending) in shared memory, then mean and
// unbiased standard deviation of the top n entries -- the groupby().head(top_n) / .mean() / .std()
// (ddof = 1) of score/ScoreNormalization.py:151-166 (AS-norm) and :93-98 (S-norm, n = all).
__global__ void topn_mean_std_kernel(const float* __restrict__ S, long long ld, int ncoh, int P, int top_n, int ddof,
                                     float* __restrict__ mean, float* __restrict__ stdv) {
  extern __shared__ float sv[];
  const float* row = S + (long long)blockIdx.x * ld;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sv[i] = i < ncoh ? row[i] : -INFINITY;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sv[i], b = sv[ixj];
          const bool desc = (i & k) == 0;           // descending overall
          if (desc ? (a < b) : (a > b)) { sv[i] = b; sv[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  const int n = top_n > 0 && top_n < ncoh ? top_n : ncoh;
  __shared__ double red[2][32];
  double s1 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s1 += (double)sv[i];
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s1;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[0][w];
  const double mu = tot / (double)n;
  double s2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = (double)sv[i] - mu; s2 += d * d; }
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double q = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) q += red[1][w];
    mean[blockIdx.x] = (float)mu;
    stdv[blockIdx.x] = (float)sqrt(q / (double)(n - ddof));   // ddof = 1 like pandas (n == 1 -> NaN as there); 0 like np.std
  }
}

// Cohort indices of the top_n largest scores of every row, in descending score order (ties: lower index
// first): bitonic sort of (score, index) pairs in shared memory.  For AS-norm with cross selection.
__global__ void topn_index_kernel(const float* __restrict__ S, long long ld, int ncoh, int P, int top_n,
                                  int32_t* __restrict__ idx_out) {
  extern __shared__ float sv[];
  int* si = reinterpret_cast<int*>(sv + P);
  const float* row = S + (long long)blockIdx.x * ld;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { sv[i] = i < ncoh ? row[i] : -INFINITY; si[i] = i; }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sv[i], b = sv[ixj];
          const int ia = si[i], ib = si[ixj];
          const bool a_first = a > b || (a == b && ia < ib);   // the order we want: a before b
          const bool desc = (i & k) == 0;
          if (desc ? !a_first : a_first) { sv[i] = b; sv[ixj] = a; si[i] = ib; si[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < top_n; i += blockDim.x) idx_out[(long long)blockIdx.x * top_n + i] = si[i];
}

// One warp per trial: enroll statistics over the TEST side's top-n cohort, test statistics over the ENROLL
// side's (ScoreNormalization.py:146-160), mean and std(ddof = 1) in double like the table path.
__global__ void snorm_cross_trials_kernel(const float* __restrict__ s, const int32_t* __restrict__ te,
                                          const int32_t* __restrict__ tt, const float* __restrict__ ec, long long lde,
                                          const float* __restrict__ tc, long long ldt, const int32_t* __restrict__ top_e,
                                          const int32_t* __restrict__ top_t, int top_n, long long n,
                                          float* __restrict__ out) {
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int e = te[w], t = tt[w];
  auto stats = [&](const float* row, const int32_t* idx, double& mu, double& sd) {
    double s1 = 0.0;
    for (int i = lane; i < top_n; i += 32) s1 += (double)row[idx[i]];
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    mu = s1 / (double)top_n;
    double s2 = 0.0;
    for (int i = lane; i < top_n; i += 32) { const double d = (double)row[idx[i]] - mu; s2 += d * d; }
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    sd = sqrt(s2 / (double)(top_n - 1));
  };
  double me, se, mt, st;
  stats(ec + (long long)e * lde, top_t + (long long)t * top_n, me, se);
  stats(tc + (long long)t * ldt, top_e + (long long)e * top_n, mt, st);
  if (lane == 0) {
    const double v = (double)s[w];
    out[w] = (float)(0.5 * ((v - me) / se + (v - mt) / st));
  }
}

__global__ void snorm_trials_kernel(const float* __restrict__ s, const int32_t* __restrict__ te,
                                    const int32_t* __restrict__ tt, const float* __restrict__ me,
                                    const float* __restrict__ se, const float* __restrict__ mt,
                                    const float* __restrict__ st, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = s[i];
    const int e = te[i], t = tt[i];
    out[i] = 0.5f * ((v - me[e]) / se[e] + (v - mt[t]) / st[t]);   // ScoreNormalization.py:101-104 / :172-173
  }
}


// PLDA training helpers (score/pyplda/plda_base.py:50-66, :262-287).  Both write TRANSPOSED (D, N) matrices
// so that the following Gram product X^T X is an A.B^T GEMM over K-contiguous rows.
// out[d][i] = sw[spk[i]] * (x[i][d] - mean[spk[i]][d])
__global__ void center_rows_T_kernel(const float* __restrict__ x, const int32_t* __restrict__ spk,
                                     const float* __restrict__ means, const float* __restrict__ sw, long long N, int D,
                                     float* __restrict__ out, long long ldo) {
  __shared__ float tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long i = i0 + r;
    const int d = d0 + threadIdx.x;
    float v = 0.f;
    if (i < N && d < D) {
      const int k = spk[i];
      v = (x[i * D + d] - means[(long long)k * D + d]) * (sw ? sw[k] : 1.f);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r;
    const long long i = i0 + threadIdx.x;
    if (d < D && i < N) out[(long long)d * ldo + i] = tile[threadIdx.x][r];
  }
}

// One EM step's per-class vectors in the basis where within_var = I and between_var = diag(psi):
//   what = n psi / (1 + n psi) * u ; what_T[d][k] = sqrt(w_k) what ; resid_T[d][k] = sqrt(w_k n_k) (u - what)
__global__ void plda_em_rows_T_kernel(const float* __restrict__ u, const float* __restrict__ n, const float* __restrict__ w,
                                      const float* __restrict__ psi, int S, int D, float* __restrict__ what_T,
                                      float* __restrict__ resid_T, long long ldo) {
  __shared__ float ta[32][33], tb[32][33];
  const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, d = d0 + threadIdx.x;
    float a = 0.f, b = 0.f;
    if (k < S && d < D) {
      const float nk = n[k], wk = w ? w[k] : 1.f, p = psi[d], uu = u[(long long)k * D + d];
      const float wh = nk * p / (1.f + nk * p) * uu;
      a = sqrtf(wk) * wh;
      b = sqrtf(wk * nk) * (uu - wh);
    }
    ta[r][threadIdx.x] = a;
    tb[r][threadIdx.x] = b;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, k = k0 + threadIdx.x;
    if (d < D && k < S) {
      what_T[(long long)d * ldo + k] = ta[threadIdx.x][r];
      resid_T[(long long)d * ldo + k] = tb[threadIdx.x][r];
    }
  }
}

// Kaldi-style PLDA scoring in the diagonalised space (score/pyplda/plda_base.py: transform_ivector :93-107,
// get_normalization_factor :151-158, log_likelihood_ratio :109-136).  One warp per row.
// u <- u * sqrt(D / sum_d u_d^2 / (psi_d + 1/n))   (simple: sqrt(D) / ||u||)
__global__ void plda_normalize_rows_kernel(float* __restrict__ u, const float* __restrict__ psi, const float* __restrict__ n,
                                           long long rows, int D, int simple) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float* x = u + r * D;
  const float inv_n = 1.f / (n ? n[r] : 1.f);
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += x[d] * x[d] * (simple ? 1.f : 1.f / (psi[d] + inv_n));
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float f = sqrtf((float)D / s);
  for (int d = lane; d < D; d += 32) x[d] *= f;
}

// LLR(i,j) = [t_j^2 | t_j] . [-1/(2 v_i) | m_i / v_i] + row_i + col_j with m = n psi/(n psi + 1) u, v = 1 + psi/(n psi + 1):
// side 0 (enroll): a = [-1/(2v) | m/v], term = -1/2 (sum log v + sum m^2/v)
// side 1 (test)  : a = [t^2 | t],      term = +1/2 (sum log(psi+1) + sum t^2/(psi+1))
__global__ void plda_llr_operands_kernel(const float* __restrict__ u, const float* __restrict__ psi, const float* __restrict__ n,
                                         long long rows, int D, int side, float* __restrict__ a, float* __restrict__ term) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = u + r * D;
  float* o = a + r * 2 * D;
  const float nn = (side == 0 && n) ? n[r] : 1.f;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float p = psi[d], xv = x[d];
    if (side == 0) {
      const float m = nn * p / (nn * p + 1.f) * xv, v = 1.f + p / (nn * p + 1.f);
      o[d] = -0.5f / v;
      o[D + d] = m / v;
      acc += logf(v) + m * m / v;
    } else {
      o[d] = xv * xv;
      o[D + d] = xv;
      acc += logf(p + 1.f) + xv * xv / (p + 1.f);
    }
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) term[r] = side == 0 ? -0.5f * acc : 0.5f * acc;
}

// out (Ne, Nt) = A (Ne, D) . Bm (Nt, D)^T  [+ row_bias[i] + col_bias[j]] through the tcgen05 layer.
static int matmul_nt(const float* A, int64_t Ne, const float* Bm, int64_t Nt, int D, const float* row_bias,
                     const float* col_bias, float* out, int64_t ldo, uint16_t* out_hi, uint16_t* out_lo,
                     int64_t ldplane, cudaStream_t s, const TrialHist* th = nullptr) {
  const int64_t ldp = round_up(D, 16);
  const bool same = A == Bm && Ne == Nt;   // all pairs of one set: split once
  TempBuf ta(s), tb(s);
  int rc = ta.alloc((size_t)Ne * ldp * 2 * 2);
  if (rc) return rc;
  uint16_t* a_hi = (uint16_t*)ta.p;
  uint16_t* a_lo = a_hi + Ne * ldp;
  rc = xvb_split_f32(A, Ne, D, D, a_hi, a_lo, ldp, s);
  if (rc) return rc;
  uint16_t* b_hi = a_hi;
  uint16_t* b_lo = a_lo;
  if (!same) {
    rc = tb.alloc((size_t)Nt * ldp * 2 * 2);
    if (rc) return rc;
    b_hi = (uint16_t*)tb.p;
    b_lo = b_hi + Nt * ldp;
    rc = xvb_split_f32(Bm, Nt, D, D, b_hi, b_lo, ldp, s);
    if (rc) return rc;
  }
  const int ctx0 = 0;
  xvb_tdnn_args_t g{};
  g.x_hi = a_hi; g.x_lo = a_lo; g.ldx = ldp; g.w_hi = b_hi; g.w_lo = b_lo;
  g.bias = col_bias; g.row_bias = row_bias; g.context_host = &ctx0; g.ntaps = 1;
  g.y_hi = out_hi; g.y_lo = out_lo; g.ldy = ldplane; g.y_f32 = out; g.ldyf = ldo;
  g.B = (int)Ne; g.T = 1; g.Cin = D; g.Cout = (int)Nt;
  return tdnn_affine_impl(g, s, th);
}

}  // namespace xvb

using namespace xvb;

extern "C" int xvb_center_length_norm(const float* x, const float* mean, float* y, int64_t rows, int D, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && y && rows > 0 && D > 0, "xvb_center_length_norm: bad arguments");
  const long long threads = rows * 32;
  center_length_norm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, mean, y, rows, D);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_column_mean(const float* x, int64_t rows, int D, float* mean, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && mean && rows > 0 && D > 0, "xvb_column_mean: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int G = sm_count() * 4;
  if (G > rows) G = (int)rows;
  TempBuf t(s);
  rc = t.alloc((size_t)G * D * sizeof(double));
  if (rc) return rc;
  column_partial_kernel<<<G, D < 512 ? ((D + 31) / 32) * 32 : 512, 0, s>>>(x, rows, D, (double*)t.p);
  XVB_LAUNCH_CHECK();
  column_final_kernel<<<(D + 127) / 128, 128, 0, s>>>((const double*)t.p, G, rows, D, mean);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_speaker_mean(const float* x, int D, const int32_t* offsets, const int32_t* members, int num_spk,
                                float* out, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && offsets && members && out && D > 0 && num_spk > 0, "xvb_speaker_mean: bad arguments");
  speaker_mean_kernel<<<num_spk, D < 256 ? ((D + 31) / 32) * 32 : 256, 0, (cudaStream_t)stream>>>(x, D, offsets, members, out);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_topn_mean_std(const float* S, int64_t lds, int64_t rows, int ncoh, int top_n, float* mean, float* stdv,
                                 void* stream) {
  return xvb_topn_mean_std_ddof(S, lds, rows, ncoh, top_n, 1, mean, stdv, stream);
}

extern "C" int xvb_topn_mean_std_ddof(const float* S, int64_t lds, int64_t rows, int ncoh, int top_n, int ddof, float* mean,
                                      float* stdv, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(S && mean && stdv && rows > 0 && ncoh > 0 && lds >= ncoh && (ddof == 0 || ddof == 1),
                "xvb_topn_mean_std: bad arguments (ddof is 0 or 1)");
  XVB_CHECK_ARG(ncoh <= 32768, "xvb_topn_mean_std: cohort of %d exceeds the 32768 entries one CTA sorts on chip", ncoh);
  int P = 1;
  while (P < ncoh) P <<= 1;
  const size_t smem = (size_t)P * sizeof(float);
  XVB_ENSURE_DYN_SMEM((topn_mean_std_kernel), 32768 * 4);
  topn_mean_std_kernel<<<(unsigned)rows, 512, smem, (cudaStream_t)stream>>>(S, lds, ncoh, P, top_n, ddof, mean, stdv);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_snorm_trials(const float* scores, const int32_t* trial_e, const int32_t* trial_t, int64_t num_trials,
                                const float* mean_e, const float* std_e, const float* mean_t, const float* std_t,
                                float* out, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(scores && trial_e && trial_t && mean_e && std_e && mean_t && std_t && out, "xvb_snorm_trials: null pointer");
  if (num_trials == 0) return XVB_OK;
  long long g = (num_trials + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  snorm_trials_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(scores, trial_e, trial_t, mean_e, std_e, mean_t, std_t,
                                                                    num_trials, out);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_bilinear_trials(const float* enroll, const float* test, int D, const int32_t* trial_e,
                                   const int32_t* trial_t, int64_t num_trials, const float* row_term,
                                   const float* col_term, float* scores, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(enroll && test && trial_e && trial_t && scores && D > 0, "xvb_bilinear_trials: bad arguments");
  if (num_trials == 0) return XVB_OK;
  const long long threads = num_trials * 32;
  cosine_trials_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      enroll, test, D, trial_e, trial_t, row_term, col_term, num_trials, scores);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_cosine_trials(const float* enroll, const float* test, int D, const int32_t* trial_e,
                                 const int32_t* trial_t, int64_t num_trials, float* scores, void* stream) {
  return xvb_bilinear_trials(enroll, test, D, trial_e, trial_t, num_trials, nullptr, nullptr, scores, stream);
}

// y (rows, Dout) = x (rows, D) . M^T, M (Dout, D) -- the small projections of the back-end
// (E' = E.(Lambda+Lambda^T) for PLDA; LDA/whitening transforms of score/process.sh:205-265).
extern "C" int xvb_project(const float* x, int64_t rows, int D, const float* M, int Dout, float* y, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && M && y && rows > 0 && D > 0 && Dout > 0 && Dout % 4 == 0, "xvb_project: bad arguments (Dout%%4==0)");
  return matmul_nt(x, rows, M, Dout, D, nullptr, nullptr, y, Dout, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int xvb_cosine_matrix(const float* enroll, int64_t Ne, const float* test, int64_t Nt, int D, float* S,
                                 int64_t lds, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(enroll && test && S && Ne > 0 && Nt > 0 && D > 0, "xvb_cosine_matrix: bad arguments");
  XVB_CHECK_ARG(Nt % 4 == 0 && lds % 4 == 0 && lds >= Nt, "xvb_cosine_matrix: Nt and lds must be multiples of 4");
  XVB_CHECK_ARG(Ne < (1ll << 31) && Nt < (1ll << 31), "xvb_cosine_matrix: too many rows for one call");
  return matmul_nt(enroll, Ne, test, Nt, D, nullptr, nullptr, S, lds, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int xvb_plda_terms(const float* x, int64_t rows, int D, const float* gamma, const float* c, float* term,
                              void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && gamma && c && term && rows > 0 && D > 0 && D % 4 == 0, "xvb_plda_terms: bad arguments (D%%4==0)");
  cudaStream_t s = (cudaStream_t)stream;
  TempBuf y(s);
  rc = y.alloc((size_t)rows * D * sizeof(float));
  if (rc) return rc;
  // y = x . Gamma^T; Gamma is symmetric (sum of inverses of symmetric matrices), so this is x Gamma.
  rc = matmul_nt(x, rows, gamma, D, D, nullptr, nullptr, (float*)y.p, D, nullptr, nullptr, 0, s);
  if (rc) return rc;
  const long long threads = rows * 32;
  plda_rowterm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(x, (const float*)y.p, D, c, rows, D, term);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_plda_matrix(const float* enroll, int64_t Ne, const float* test, int64_t Nt, int D, const float* L2,
                               const float* row, const float* col, float* S, int64_t lds, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(enroll && test && L2 && S && Ne > 0 && Nt > 0 && D > 0 && D % 4 == 0, "xvb_plda_matrix: bad arguments (D%%4==0)");
  XVB_CHECK_ARG(Nt % 4 == 0 && lds % 4 == 0 && lds >= Nt, "xvb_plda_matrix: Nt and lds must be multiples of 4");
  cudaStream_t s = (cudaStream_t)stream;
  TempBuf el(s);
  rc = el.alloc((size_t)Ne * D * sizeof(float));
  if (rc) return rc;
  // E' = E . L2   (L2 = Lambda + Lambda^T is symmetric, so E L2 = E L2^T)
  rc = matmul_nt(enroll, Ne, L2, D, D, nullptr, nullptr, (float*)el.p, D, nullptr, nullptr, 0, s);
  if (rc) return rc;
  return matmul_nt((const float*)el.p, Ne, test, Nt, D, row, col, S, lds, nullptr, nullptr, 0, s);
}

// Scores -> histogram, fused: S = enroll . test^T + row_term + col_term is consumed tile by tile in
// the GEMM epilogue (shared-memory counters, one u64 flush per CTA), so 10^12 trials cost no HBM
// traffic beyond the embeddings themselves.  SURVEY Appendix A, C4/C5 "fused consumer".
extern "C" int xvb_trial_histogram(const float* enroll, int64_t Ne, const int32_t* enroll_spk, const float* test,
                                   int64_t Nt, const int32_t* test_spk, int D, const float* row_term,
                                   const float* col_term, int symmetric, int unit_first, int unit_stride, float lo,
                                   float hi, int nbins, unsigned long long* hist, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(enroll && test && enroll_spk && test_spk && hist && Ne > 0 && Nt > 0 && D > 0, "xvb_trial_histogram: bad arguments");
  XVB_CHECK_ARG(Ne < (1ll << 31) && Nt < (1ll << 31), "xvb_trial_histogram: too many rows for one call");
  XVB_CHECK_ARG(nbins >= 4 && nbins <= 2048, "xvb_trial_histogram: nbins=%d outside [4, 2048] (2 x nbins u32 counters live in 16 KB of shared memory)", nbins);
  XVB_CHECK_ARG(hi > lo, "xvb_trial_histogram: empty score window [%g, %g)", (double)lo, (double)hi);
  XVB_CHECK_ARG(unit_first >= 0 && unit_stride >= 1, "xvb_trial_histogram: bad row-unit shard %d/%d", unit_first, unit_stride);
  XVB_CHECK_ARG(!symmetric || Ne == Nt, "xvb_trial_histogram: symmetric mode needs one set on both sides (Ne == Nt)");
  TrialHist th{};
  th.hist = hist; th.row_label = enroll_spk; th.col_label = test_spk;
  th.lo = lo; th.inv_w = (float)(nbins - 2) / (hi - lo); th.nbins = nbins; th.symmetric = symmetric ? 1 : 0;
  th.unit_first = unit_first; th.unit_stride = unit_stride;
  return matmul_nt(enroll, Ne, test, Nt, D, row_term, col_term, nullptr, 0, nullptr, nullptr, 0, (cudaStream_t)stream, &th);
}

// Generic A.B^T (+ row/column terms) on the tcgen05 layer: out (M, N) = a (M, K) . b (N, K)^T + row[i] + col[j].
extern "C" int xvb_matmul_nt(const float* a, int64_t M, const float* b, int64_t N, int K, const float* row_bias,
                             const float* col_bias, float* out, int64_t ldo, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(a && b && out && M > 0 && N > 0 && K > 0, "xvb_matmul_nt: bad arguments");
  XVB_CHECK_ARG(N % 4 == 0 && ldo % 4 == 0 && ldo >= N, "xvb_matmul_nt: N and ldo must be multiples of 4");
  XVB_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31), "xvb_matmul_nt: too many rows for one call");
  return matmul_nt(a, M, b, N, K, row_bias, col_bias, out, ldo, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int xvb_center_rows_transposed(const float* x, const int32_t* spk, const float* means, const float* sqrt_weight,
                                          int64_t N, int D, float* out, int64_t ldo, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x && spk && means && out && N > 0 && D > 0 && ldo >= N, "xvb_center_rows_transposed: bad arguments");
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)((D + 31) / 32)), block(32, 8);
  center_rows_T_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, spk, means, sqrt_weight, N, D, out, ldo);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_plda_em_rows(const float* u, const float* n, const float* weight, const float* psi, int S, int D,
                                float* what_T, float* resid_T, int64_t ldo, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(u && n && psi && what_T && resid_T && S > 0 && D > 0 && ldo >= S, "xvb_plda_em_rows: bad arguments");
  dim3 grid((unsigned)((S + 31) / 32), (unsigned)((D + 31) / 32)), block(32, 8);
  plda_em_rows_T_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(u, n, weight, psi, S, D, what_T, resid_T, ldo);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_topn_indices(const float* S, int64_t lds, int64_t rows, int ncoh, int top_n, int32_t* idx, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(S && idx && rows > 0 && ncoh > 0 && lds >= ncoh && top_n >= 1 && top_n <= ncoh, "xvb_topn_indices: bad arguments");
  XVB_CHECK_ARG(ncoh <= 16384, "xvb_topn_indices: cohort of %d exceeds the 16384 (score, index) pairs one CTA sorts on chip", ncoh);
  int P = 1;
  while (P < ncoh) P <<= 1;
  XVB_ENSURE_DYN_SMEM((topn_index_kernel), 16384 * 8);
  topn_index_kernel<<<(unsigned)rows, 512, (size_t)P * 8, (cudaStream_t)stream>>>(S, lds, ncoh, P, top_n, idx);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_snorm_cross_trials(const float* scores, const int32_t* trial_e, const int32_t* trial_t, int64_t num_trials,
                                      const float* enroll_cohort, int64_t lde, const float* test_cohort, int64_t ldt,
                                      const int32_t* top_enroll, const int32_t* top_test, int top_n, float* out,
                                      void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(scores && trial_e && trial_t && enroll_cohort && test_cohort && top_enroll && top_test && out && top_n >= 2,
                "xvb_snorm_cross_trials: bad arguments (top_n >= 2)");
  if (num_trials == 0) return XVB_OK;
  const long long threads = num_trials * 32;
  snorm_cross_trials_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      scores, trial_e, trial_t, enroll_cohort, lde, test_cohort, ldt, top_enroll, top_test, top_n, num_trials, out);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_plda_normalize_rows(float* u, const float* psi, const float* num_examples, int64_t rows, int D,
                                       int simple_length_norm, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(u && psi && rows > 0 && D > 0, "xvb_plda_normalize_rows: bad arguments");
  const long long threads = rows * 32;
  plda_normalize_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, psi, num_examples, rows, D,
                                                                                              simple_length_norm);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

extern "C" int xvb_plda_llr_operands(const float* u, const float* psi, const float* num_examples, int64_t rows, int D, int side,
                                     float* operand, float* term, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(u && psi && operand && term && rows > 0 && D > 0 && (side == 0 || side == 1), "xvb_plda_llr_operands: bad arguments");
  const long long threads = rows * 32;
  plda_llr_operands_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, psi, num_examples, rows, D, side,
                                                                                            operand, term);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}
