// TDNN affine layer as a tcgen05 GEMM (sm_100a).
//
//   y[b,t,n] = epi( bias[n] + sum_{tap} sum_{c} W[n, tap, c] * x[b, t + ctx[tap], c] )
//
// Restates TdnnAffine.forward + ReLU + eval-BatchNorm of the reference
// (pytorch/libs/nnet/components.py:107-149, :410-431) as ONE kernel:
//   * M = B*T frames, N = Cout, K = ntaps*Cin -- only the taps in `context` (the reference's
//     conv1d also multiplies the masked taps, components.py:133-138);
//   * the context splice is never materialised: the A tile of tap `c` is fetched by TMA from the
//     (C, T, B) frame matrix at time coordinate t0 + c; TMA's out-of-bounds zero fill *is*
//     F.pad(..., value=0) (components.py:117) and can never cross an utterance boundary;
//   * fp32-grade accuracy at bf16 tensor rate: operands are bf16 "split planes" (hi, lo) and each
//     K step issues hi*hi + lo*hi + hi*lo into the same fp32 TMEM accumulator;
//   * warp-specialised persistent CTAs: warp0 = TMA producer, warp1 = tcgen05.mma issuer,
//     warps2-5 = epilogue (tcgen05.ld -> +bias -> ReLU -> BN -> split -> global), double-buffered
//     accumulators in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// An M tile is 128 rows = Bb utterances x Tb consecutive frames (Tb*Bb = 128, chosen on the host
// to minimise padding: T=200 -> Tb=8, Bb=16 has none).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace xvb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB per plane per stage
constexpr int kNumThreads = 192;

struct TdnnGemmParams {
  int B, T, Cin, Cout;
  int Tb, Bb, num_t_blk, num_b_blk, num_n_blk, num_tiles;
  int ntaps, cin_p16, num_cblk;
  int ctx[XVB_MAX_TAPS];
  int flags;
  const float* bias;
  const float* scale;
  const float* shift;
  const float* row_bias;  // per-frame additive term (PLDA row term), may be NULL
  __nv_bfloat16* y_hi;
  __nv_bfloat16* y_lo;
  long long ldy;
  float* y_f32;
  long long ldyf;
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // two accumulator stages
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "need at least a double-buffered operand pipeline");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
tdnn_gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                        const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                        const __grid_constant__ TdnnGemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kBBytes = Cfg::kBBytes;
  constexpr int kStageBytes = Cfg::kStageBytes;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_w_hi);
    tma_prefetch_desc(&map_w_lo);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, Cfg::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_kblk = p.ntaps * p.num_cblk;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_blk, n_blk = tile % p.num_n_blk;
        const int b0 = (m_blk / p.num_t_blk) * p.Bb, t0 = (m_blk % p.num_t_blk) * p.Tb;
        const int n0 = n_blk * BLOCK_N;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int tt = t0 + p.ctx[tap];
          for (int cb = 0; cb < p.num_cblk; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* s = smem + stage * kStageBytes;
            mbar_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_3d(s, &map_a_hi, &full_bar[stage], cb * kBlockK, tt, b0);
            tma_load_3d(s + kABytes, &map_a_lo, &full_bar[stage], cb * kBlockK, tt, b0);
            const int kw = tap * p.cin_p16 + cb * kBlockK;
            tma_load_2d(s + 2 * kABytes, &map_w_hi, &full_bar[stage], kw, n0);
            tma_load_2d(s + 2 * kABytes + kBBytes, &map_w_lo, &full_bar[stage], kw, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < num_kblk; ++kb) {
          const int cb = kb % p.num_cblk;
          int nsteps = (p.Cin - cb * kBlockK + 15) >> 4;
          nsteps = nsteps > 4 ? 4 : nsteps;
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da_hi = make_kmajor_desc<128>(sa);
          const uint64_t da_lo = make_kmajor_desc<128>(sa + kABytes);
          const uint64_t db_hi = make_kmajor_desc<128>(sa + 2 * kABytes);
          const uint64_t db_lo = make_kmajor_desc<128>(sa + 2 * kABytes + kBBytes);
          for (int s = 0; s < nsteps; ++s) {
            const uint64_t koff = (uint64_t)(s * 32 >> 4);  // 16 bf16 = 32 bytes along K inside the swizzle row
            umma_bf16<1>(tmem_d, da_lo + koff, db_hi + koff, idesc, accumulate);
            accumulate = 1;
            umma_bf16<1>(tmem_d, da_hi + koff, db_lo + koff, idesc, 1);
            umma_bf16<1>(tmem_d, da_hi + koff, db_hi + koff, idesc, 1);
          }
          umma_commit<1>(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit<1>(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const bool relu = (p.flags & XVB_RELU) != 0;
    const bool bn = (p.flags & XVB_BN) != 0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int m_blk = tile / p.num_n_blk, n_blk = tile % p.num_n_blk;
      const int b = (m_blk / p.num_t_blk) * p.Bb + row / p.Tb;
      const int t = (m_blk % p.num_t_blk) * p.Tb + row % p.Tb;
      const bool valid = (b < p.B) && (t < p.T);
      const long long frame = (long long)b * p.T + t;
      const int n0 = n_blk * BLOCK_N;
      const float rbias = (p.row_bias && valid) ? __ldg(p.row_bias + frame) : 0.f;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        const int n = n0 + ch * 32;
        if (n >= p.Cout) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = n + j;
          float x = __uint_as_float(v[j]) + rbias;
          if (c < p.Cout) {
            if (p.bias) x += __ldg(p.bias + c);
            if (relu) x = fmaxf(x, 0.f);
            if (bn) x = fmaf(x, __ldg(p.scale + c), __ldg(p.shift + c));
          }
          f[j] = x;
        }
        if (valid) {
          if (p.y_f32) {
            float* dst = p.y_f32 + frame * p.ldyf + n;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (n + g * 4 < p.Cout)
                *reinterpret_cast<float4*>(dst + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
          }
          if (p.y_hi) {
            __nv_bfloat16* dh = p.y_hi + frame * p.ldy + n;
            __nv_bfloat16* dl = p.y_lo + frame * p.ldy + n;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (n + g * 8 < p.Cout) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  __nv_bfloat16 h0, l0, h1, l1;
                  split_bf16(f[g * 8 + 2 * k], h0, l0);
                  split_bf16(f[g * 8 + 2 * k + 1], h1, l1);
                  h[k] = pack_bf16x2(h0, h1);
                  l[k] = pack_bf16x2(l0, l1);
                }
                *reinterpret_cast<uint4*>(dh + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(dl + g * 8) = make_uint4(l[0], l[1], l[2], l[3]);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// (C, T, B) bf16 frame matrix with row pitch ld; box = 64 channels x Tb frames x Bb utterances.
static int make_frame_map(CUtensorMap* m, const void* base, int C, int T, int B, long long ld, int Tb, int Bb) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return XVB_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)T};
  cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)Tb, (cuuint32_t)Bb};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(frame map C=%d T=%d B=%d ld=%lld) failed: %d", C, T, B, ld, (int)r); return XVB_ECUDA; }
  return XVB_OK;
}

// (K, Cout) bf16 packed weight, K contiguous; box = 64 x block_n.
static int make_weight_map(CUtensorMap* m, const void* base, long long K, int Cout, int block_n) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return XVB_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weight map K=%lld Cout=%d) failed: %d", K, Cout, (int)r); return XVB_ECUDA; }
  return XVB_OK;
}

// Pick the (Tb, Bb) factorisation of the 128-row M tile with the fewest padded rows.
static void choose_m_tile(int B, int T, int* Tb_out, int* Bb_out) {
  long long best = -1;
  int bt = 128;
  for (int Tb = 128; Tb >= 1; Tb >>= 1) {
    const int Bb = 128 / Tb;
    const long long rows = (long long)((T + Tb - 1) / Tb) * Tb * ((B + Bb - 1) / Bb) * Bb;
    if (best < 0 || rows < best) { best = rows; bt = Tb; }
  }
  *Tb_out = bt;
  *Bb_out = 128 / bt;
}

template <int BLOCK_N>
static int launch_gemm(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const void* w_hi, const void* w_lo,
                       TdnnGemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap mw_hi, mw_lo;
  const long long K = (long long)p.ntaps * p.cin_p16;
  int rc = make_weight_map(&mw_hi, w_hi, K, p.Cout, BLOCK_N);
  if (rc) return rc;
  rc = make_weight_map(&mw_lo, w_lo, K, p.Cout, BLOCK_N);
  if (rc) return rc;
  p.num_n_blk = (p.Cout + BLOCK_N - 1) / BLOCK_N;
  p.num_tiles = p.num_t_blk * p.num_b_blk * p.num_n_blk;
  static bool attr_set = false;
  if (!attr_set) {
    XVB_CUDA(cudaFuncSetAttribute(tdnn_gemm_bf16x3_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    attr_set = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  tdnn_gemm_bf16x3_kernel<BLOCK_N><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, p);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}

}  // namespace xvb

using namespace xvb;

int xvb::tdnn_affine_impl(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, const uint16_t* w_hi,
                          const uint16_t* w_lo, const float* bias, const float* bn_scale, const float* bn_shift,
                          const float* row_bias, int flags, const int* context_host, int ntaps, uint16_t* y_hi,
                          uint16_t* y_lo, int64_t ldy, float* y_f32, int64_t ldyf, int B, int T, int Cin, int Cout,
                          void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x_hi && x_lo && w_hi && w_lo, "xvb_tdnn_affine: null operand pointer");
  XVB_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0, "xvb_tdnn_affine: bad shape B=%d T=%d Cin=%d Cout=%d", B, T, Cin, Cout);
  XVB_CHECK_ARG(ntaps >= 1 && ntaps <= XVB_MAX_TAPS && context_host, "xvb_tdnn_affine: ntaps=%d out of range", ntaps);
  XVB_CHECK_ARG(ldx % 8 == 0 && ldx >= Cin, "xvb_tdnn_affine: ldx=%lld must be a multiple of 8 and >= Cin", (long long)ldx);
  XVB_CHECK_ARG((y_hi != nullptr) == (y_lo != nullptr), "xvb_tdnn_affine: y_hi/y_lo must both be set or both NULL");
  XVB_CHECK_ARG(y_hi || y_f32, "xvb_tdnn_affine: no output requested");
  if (y_hi) XVB_CHECK_ARG(ldy % 8 == 0 && Cout % 8 == 0 && ldy >= Cout, "xvb_tdnn_affine: plane output needs ldy%%8==0, Cout%%8==0");
  if (y_f32) XVB_CHECK_ARG(ldyf % 4 == 0 && Cout % 4 == 0 && ldyf >= Cout, "xvb_tdnn_affine: fp32 output needs ldyf%%4==0, Cout%%4==0");
  XVB_CHECK_ARG(!(flags & XVB_BN) || (bn_scale && bn_shift), "xvb_tdnn_affine: XVB_BN without scale/shift");
  XVB_CHECK_ARG(((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)y_hi |
                 (uintptr_t)y_lo | (uintptr_t)y_f32) % 16 == 0, "xvb_tdnn_affine: pointers must be 16-byte aligned");
  for (int i = 1; i < ntaps; ++i)
    XVB_CHECK_ARG(context_host[i] > context_host[i - 1], "xvb_tdnn_affine: context must be strictly increasing (components.py:34-36)");

  TdnnGemmParams p{};
  p.B = B; p.T = T; p.Cin = Cin; p.Cout = Cout;
  choose_m_tile(B, T, &p.Tb, &p.Bb);
  p.num_t_blk = (T + p.Tb - 1) / p.Tb;
  p.num_b_blk = (B + p.Bb - 1) / p.Bb;
  p.ntaps = ntaps;
  p.cin_p16 = (int)round_up(Cin, 16);
  p.num_cblk = (Cin + kBlockK - 1) / kBlockK;
  for (int i = 0; i < ntaps; ++i) p.ctx[i] = context_host[i];
  p.flags = flags;
  p.bias = bias; p.scale = bn_scale; p.shift = bn_shift; p.row_bias = row_bias;
  p.y_hi = reinterpret_cast<__nv_bfloat16*>(y_hi);
  p.y_lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
  p.ldy = ldy; p.y_f32 = y_f32; p.ldyf = ldyf;

  CUtensorMap ma_hi, ma_lo;
  rc = make_frame_map(&ma_hi, x_hi, Cin, T, B, ldx, p.Tb, p.Bb);
  if (rc) return rc;
  rc = make_frame_map(&ma_lo, x_lo, Cin, T, B, ldx, p.Tb, p.Bb);
  if (rc) return rc;

  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // Wide N tiles when there are enough M tiles to fill the machine, narrow ones for the
  // segment-level layers (M = B rows) so that more SMs get a tile.
  const long long m_tiles = (long long)p.num_t_blk * p.num_b_blk;
  const int sms = sm_count();
  if (Cout >= 256 && m_tiles * ((Cout + 255) / 256) >= sms) return launch_gemm<256>(ma_hi, ma_lo, w_hi, w_lo, p, s);
  if (Cout >= 128 && m_tiles * ((Cout + 127) / 128) >= sms) return launch_gemm<128>(ma_hi, ma_lo, w_hi, w_lo, p, s);
  if (Cout >= 64 && m_tiles * ((Cout + 63) / 64) >= sms / 2) return launch_gemm<64>(ma_hi, ma_lo, w_hi, w_lo, p, s);
  return launch_gemm<32>(ma_hi, ma_lo, w_hi, w_lo, p, s);
}

extern "C" int xvb_tdnn_affine(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, const uint16_t* w_hi,
                               const uint16_t* w_lo, const float* bias, const float* bn_scale, const float* bn_shift,
                               int flags, const int* context_host, int ntaps, uint16_t* y_hi, uint16_t* y_lo,
                               int64_t ldy, float* y_f32, int64_t ldyf, int B, int T, int Cin, int Cout,
                               void* stream) {
  return tdnn_affine_impl(x_hi, x_lo, ldx, w_hi, w_lo, bias, bn_scale, bn_shift, nullptr, flags, context_host, ntaps,
                          y_hi, y_lo, ldy, y_f32, ldyf, B, T, Cin, Cout, stream);
}
