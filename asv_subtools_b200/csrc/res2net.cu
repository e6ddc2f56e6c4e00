// Res2Net block of ECAPA-TDNN as ONE persistent kernel (Res2NetBlock.forward,
// pytorch/model/ecapa_tdnn_xvector.py:61-75).
//
//   y_0 = x_0 ;  y_{i+1} = BN(ReLU(W_i * splice(x_{i+1} + [i>=1] y_i, [-d,0,d]) + b_i)),  i = 0..6
//
// The seven 128->128 dilated TDNN layers form a serial chain, but the chain only couples frames of the
// SAME utterance.  Launching them as seven grid-wide GEMMs (150 pair tiles on 74 SM pairs, a nearly
// empty third wave each, plus launch + drain) ran at ~25 % efficiency.  Here a CTA OWNS utterances: it
// walks its utterance through all seven steps, tile by tile, with the same TMA -> tcgen05 -> TMEM
// pipeline as tdnn_gemm.cu (bf16x3 split operands, fp32 accumulators, double-buffered in TMEM).
// Step i reads chunk i+1 of the input tensor and -- as a second A source accumulated into the same
// TMEM accumulator (W(a+b) = Wa + Wb) -- chunk i of the OUTPUT tensor, which this very CTA wrote in
// step i-1: the only synchronisation is CTA-local (the epilogue drains its TMA stores, then releases
// the producer through an mbarrier), no grid-wide barrier and no extra launches.  Chunk 0 is copied
// through by the epilogue warps on the way.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "ptx.cuh"

namespace xvb {

constexpr int kRW = 128;                           // Res2Net width: channels per chunk = tile N
constexpr int kRStages = 3;
constexpr int kRABytes = 128 * 64 * 2;             // one plane of a 128-row x 64-channel A tile
constexpr int kRBBytes = kRW * 64 * 2;             // one plane of the 128 x 64 weight tile
constexpr int kRStageBytes = 2 * kRABytes + 2 * kRBBytes;   // 64 KB
constexpr int kRSlabBytes = 16384;
constexpr int kRThreads = 64 + 8 * 32;
constexpr int kRSmemBytes = kRStages * kRStageBytes + kRSlabBytes + 2 * 3 * kRW * 4 + 1024 + 256;

struct Res2Params {
  int B, T, C;            // C = scale * 128 channels of the block input / output
  int num_steps;          // scale - 1 (= 7)
  int dilation;
  int num_m;              // ceil(T / 128)
  const float* bias;      // [num_steps][128]
  const float* scale;
  const float* shift;
  const __nv_bfloat16* x_hi;   // block input planes (B,T,ldx): chunk 0 is passed through
  const __nv_bfloat16* x_lo;
  __nv_bfloat16* y_hi;         // block output planes (B,T,ldy)
  __nv_bfloat16* y_lo;
  long long ldx, ldy;
  int debug;              // timing experiments only (XVB_RES2_DEBUG, wrong results): bit0 one A source, bit1 every store issued twice
};

__global__ void __launch_bounds__(kRThreads, 1)
res2net_chain_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                     const __grid_constant__ CUtensorMap map_yin_hi, const __grid_constant__ CUtensorMap map_yin_lo,
                     const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                     const __grid_constant__ CUtensorMap map_yout_hi, const __grid_constant__ CUtensorMap map_yout_lo,
                     const __grid_constant__ Res2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab_base = smem + kRStages * kRStageBytes;
  float* param_base = reinterpret_cast<float*>(slab_base + kRSlabBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slab_base + kRSlabBytes + 2 * 3 * kRW * 4);
  uint64_t* empty_bar = full_bar + kRStages;
  uint64_t* tmem_full_bar = empty_bar + kRStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* step_bar = tmem_empty_bar + 2;        // completes once per step: that step's outputs are in global memory
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(step_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_x_hi); tma_prefetch_desc(&map_x_lo);
    tma_prefetch_desc(&map_yin_hi); tma_prefetch_desc(&map_yin_lo);
    tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
    for (int i = 0; i < kRStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 256); }
    mbar_init(step_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 2 * kRW);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int d = p.dilation;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, steps_done = 0;   // number of completed step_bar phases this thread has consumed
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int st = 0; st < p.num_steps; ++st) {
          for (int m = 0; m < p.num_m; ++m) {
            const int t0 = m * 128;
            for (int src = 0; src < ((st == 0 || (p.debug & 1)) ? 1 : 2); ++src) {
              if (src == 1 && m == 0) {
                // outputs of step st-1 (all tiles of this utterance) must have landed before we read them back
                mbar_wait(step_bar, steps_done & 1);
                ++steps_done;
                fence_proxy_async();
              }
              const CUtensorMap* mh = src == 0 ? &map_x_hi : &map_yin_hi;
              const CUtensorMap* ml = src == 0 ? &map_x_lo : &map_yin_lo;
              const int cbase = src == 0 ? (st + 1) * kRW : st * kRW;   // chunk st+1 of x, chunk st of y
              for (int tap = 0; tap < 3; ++tap) {
                const int tt = t0 + (tap - 1) * d;
                for (int cb = 0; cb < 2; ++cb) {
                  mbar_wait(&empty_bar[stage], phase ^ 1);
                  uint8_t* s = smem + stage * kRStageBytes;
                  mbar_expect_tx(&full_bar[stage], kRStageBytes);
                  tma_load_3d(s, mh, &full_bar[stage], cbase + cb * 64, tt, b);
                  tma_load_3d(s + kRABytes, ml, &full_bar[stage], cbase + cb * 64, tt, b);
                  const int kw = tap * kRW + cb * 64;
                  tma_load_2d(s + 2 * kRABytes, &map_w_hi, &full_bar[stage], kw, st * kRW);
                  tma_load_2d(s + 2 * kRABytes + kRBBytes, &map_w_lo, &full_bar[stage], kw, st * kRW);
                  if (++stage == kRStages) { stage = 0; phase ^= 1; }
                }
              }
            }
          }
        }
        // the last step's stores are drained by the epilogue before it starts the next utterance's
        // chunk-0 copy; consume that phase too so the parity bookkeeping stays in step
        mbar_wait(step_bar, steps_done & 1);
        ++steps_done;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kRW);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int st = 0; st < p.num_steps; ++st) {
          const int nkb = ((st == 0 || (p.debug & 1)) ? 1 : 2) * 6;
          for (int m = 0; m < p.num_m; ++m, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc * kRW;
            uint32_t accumulate = 0;
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
              const uint32_t sa = smem_u32(smem + stage * kRStageBytes);
              const uint64_t da_hi = make_kmajor_desc<128>(sa), da_lo = make_kmajor_desc<128>(sa + kRABytes);
              const uint64_t db_hi = make_kmajor_desc<128>(sa + 2 * kRABytes);
              const uint64_t db_lo = make_kmajor_desc<128>(sa + 2 * kRABytes + kRBBytes);
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint64_t koff = (uint64_t)(s * 32 >> 4);
                umma_bf16<1>(tmem_d, da_lo + koff, db_hi + koff, idesc, accumulate);
                accumulate = 1;
                umma_bf16<1>(tmem_d, da_hi + koff, db_lo + koff, idesc, 1);
                umma_bf16<1>(tmem_d, da_hi + koff, db_hi + koff, idesc, 1);
              }
              umma_commit<1>(&empty_bar[stage]);
              if (++stage == kRStages) { stage = 0; phase ^= 1; }
            }
            umma_commit<1>(&tmem_full_bar[acc]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    const int ew = warp - 2, half = ew >> 2, q = warp & 3;
    const int row = q * 32 + lane, etid = threadIdx.x - 64;
    const bool leader = etid == 0;
    const uint32_t slab = smem_u32(slab_base);
    uint32_t it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
      // chunk 0 passes through (ecapa_tdnn_xvector.py:63-64): 16-byte vectors, 8 per row and plane
      for (int i = etid; i < p.T * 16 * 2; i += 256) {
        const int plane = i / (p.T * 16), r = (i % (p.T * 16)) >> 4, v = i & 15;
        const __nv_bfloat16* src = (plane ? p.x_lo : p.x_hi) + ((long long)b * p.T + r) * p.ldx + v * 8;
        __nv_bfloat16* dst = (plane ? p.y_lo : p.y_hi) + ((long long)b * p.T + r) * p.ldy + v * 8;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
      }
      for (int st = 0; st < p.num_steps; ++st) {
        for (int m = 0; m < p.num_m; ++m, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          const int t0 = m * 128;
          const uint32_t prm = smem_u32(param_base) + acc * (3 * kRW * 4);
          if (etid < kRW) {
            st_shared_f32(prm + etid * 4, __ldg(p.bias + st * kRW + etid));
            st_shared_f32(prm + (kRW + etid) * 4, __ldg(p.scale + st * kRW + etid));
            st_shared_f32(prm + (2 * kRW + etid) * 4, __ldg(p.shift + st * kRW + etid));
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          mbar_wait(&tmem_full_bar[acc], acc_phase);
          tcgen05_fence_after();
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kRW + half * 16;
          auto process = [&](uint32_t (&v)[16], int ch) {
            const int pc = ch * 32 + half * 16;
            float f[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 bb = ld_shared_f4(prm + (pc + 4 * g) * 4);
              const float4 ss = ld_shared_f4(prm + (kRW + pc + 4 * g) * 4);
              const float4 tt = ld_shared_f4(prm + (2 * kRW + pc + 4 * g) * 4);
              f[4 * g + 0] = fmaf(fmaxf(__uint_as_float(v[4 * g + 0]) + bb.x, 0.f), ss.x, tt.x);
              f[4 * g + 1] = fmaf(fmaxf(__uint_as_float(v[4 * g + 1]) + bb.y, 0.f), ss.y, tt.y);
              f[4 * g + 2] = fmaf(fmaxf(__uint_as_float(v[4 * g + 2]) + bb.z, 0.f), ss.z, tt.z);
              f[4 * g + 3] = fmaf(fmaxf(__uint_as_float(v[4 * g + 3]) + bb.w, 0.f), ss.w, tt.w);
            }
            if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t sh = slab + row * 64, sl = slab + 8192 + row * 64;
            const int sw = (row >> 1) & 3;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(f[g * 8 + 2 * k], h0, l0);
                split_bf16(f[g * 8 + 2 * k + 1], h1, l1);
                h[k] = pack_bf16x2(h0, h1);
                l[k] = pack_bf16x2(l0, l1);
              }
              const int c = half * 2 + g;
              st_shared_v4(sh + ((c ^ sw) << 4), h[0], h[1], h[2], h[3]);
              st_shared_v4(sl + ((c ^ sw) << 4), l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (leader) {
              const int n = (st + 1) * kRW + ch * 32;       // output chunk st+1
              tma_store_3d(&map_yout_hi, slab_base, n, t0, b);
              tma_store_3d(&map_yout_lo, slab_base + 8192, n, t0, b);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (p.debug & 2) {   // timing experiment: the same hand-over a second time
              if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              asm volatile("bar.sync 1, 256;" ::: "memory");
              fence_proxy_async();
              asm volatile("bar.sync 2, 256;" ::: "memory");
              if (leader) {
                const int n = (st + 1) * kRW + ch * 32;
                tma_store_3d(&map_yout_hi, slab_base, n, t0, b);
                tma_store_3d(&map_yout_lo, slab_base + 8192, n, t0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
          };
          uint32_t va[16], vb[16];
          tmem_ld_32x16(trow, va);
          tmem_ld_wait();
          tmem_ld_32x16(trow + 32, vb);
          process(va, 0);
          tmem_ld_wait();
          tmem_ld_32x16(trow + 64, va);
          process(vb, 1);
          tmem_ld_wait();
          tmem_ld_32x16(trow + 96, vb);
          process(va, 2);
          tmem_ld_wait();
          process(vb, 3);
          tcgen05_fence_before();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
        // step finished: drain the stores, then let the producer read this step's outputs back
        if (leader) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          fence_proxy_async();
          if (!(p.debug & 1) || st == p.num_steps - 1) mbar_arrive(step_bar);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, 2 * kRW);
}

}  // namespace xvb

using namespace xvb;

extern "C" int xvb_res2net_block(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, const uint16_t* w_hi,
                                 const uint16_t* w_lo, const float* bias, const float* bn_scale, const float* bn_shift,
                                 int dilation, int scale, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, int B, int T,
                                 void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(x_hi && x_lo && w_hi && w_lo && bias && bn_scale && bn_shift && y_hi && y_lo, "xvb_res2net_block: null pointer");
  XVB_CHECK_ARG(B > 0 && T > 0 && scale >= 2 && scale <= 16 && dilation >= 1, "xvb_res2net_block: bad shape");
  const int C = scale * kRW;
  XVB_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "xvb_res2net_block: pitches must be >= %d and multiples of 8", C);
  XVB_CHECK_ARG(((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)y_hi | (uintptr_t)y_lo | (uintptr_t)w_hi | (uintptr_t)w_lo) % 16 == 0,
                "xvb_res2net_block: pointers must be 16-byte aligned");
  XVB_CHECK_ARG(x_hi != y_hi && x_lo != y_lo, "xvb_res2net_block: input and output must be distinct tensors");
  Res2Params p{};
  p.B = B; p.T = T; p.C = C; p.num_steps = scale - 1; p.dilation = dilation; p.num_m = (T + 127) / 128;
  p.bias = bias; p.scale = bn_scale; p.shift = bn_shift;
  p.x_hi = reinterpret_cast<const __nv_bfloat16*>(x_hi); p.x_lo = reinterpret_cast<const __nv_bfloat16*>(x_lo);
  p.y_hi = reinterpret_cast<__nv_bfloat16*>(y_hi); p.y_lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
  p.ldx = ldx; p.ldy = ldy;
  p.debug = getenv("XVB_RES2_DEBUG") ? atoi(getenv("XVB_RES2_DEBUG")) : 0;
  CUtensorMap mx_hi, mx_lo, myi_hi, myi_lo, mw_hi, mw_lo, myo_hi, myo_lo;
  const unsigned long long dx[3] = {(unsigned long long)C, (unsigned long long)T, (unsigned long long)B};
  const unsigned long long sx[2] = {(unsigned long long)ldx * 2, (unsigned long long)ldx * 2 * T};
  const unsigned long long sy[2] = {(unsigned long long)ldy * 2, (unsigned long long)ldy * 2 * T};
  const unsigned box_a[3] = {64u, 128u, 1u}, box_o[3] = {32u, 128u, 1u};
  if ((rc = make_tensor_map(&mx_hi, x_hi, 2, 3, dx, sx, box_a, 128))) return rc;
  if ((rc = make_tensor_map(&mx_lo, x_lo, 2, 3, dx, sx, box_a, 128))) return rc;
  if ((rc = make_tensor_map(&myi_hi, y_hi, 2, 3, dx, sy, box_a, 128))) return rc;
  if ((rc = make_tensor_map(&myi_lo, y_lo, 2, 3, dx, sy, box_a, 128))) return rc;
  if ((rc = make_tensor_map(&myo_hi, y_hi, 2, 3, dx, sy, box_o, 64))) return rc;
  if ((rc = make_tensor_map(&myo_lo, y_lo, 2, 3, dx, sy, box_o, 64))) return rc;
  const unsigned long long dw[2] = {(unsigned long long)3 * kRW, (unsigned long long)(scale - 1) * kRW};
  const unsigned long long sw[1] = {(unsigned long long)3 * kRW * 2};
  const unsigned box_w[2] = {64u, (unsigned)kRW};
  if ((rc = make_tensor_map(&mw_hi, w_hi, 2, 2, dw, sw, box_w, 128))) return rc;
  if ((rc = make_tensor_map(&mw_lo, w_lo, 2, 2, dw, sw, box_w, 128))) return rc;
  XVB_ENSURE_DYN_SMEM((res2net_chain_kernel), kRSmemBytes);
  const int grid = B < sm_count() ? B : sm_count();
  res2net_chain_kernel<<<grid, kRThreads, kRSmemBytes, (cudaStream_t)stream>>>(mx_hi, mx_lo, myi_hi, myi_lo, mw_hi, mw_lo,
                                                                              myo_hi, myo_lo, p);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}
