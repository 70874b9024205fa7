// FP32X: the two square ConditionalLinear layers of a reverse step (latent_model.py:177-183) at FP32 grade on the tensor
// cores -- the precision the reference computes in (latent_model.py:169-184 is plain FP32 PyTorch) -- for any
// feature_dim.  Mainloop: ladine_split.cuh (FP16 hi + lo operands, main + correction accumulators, chunked promotion
// into FP32 registers).  Epilogues as in the 16-bit kernels, but in natural units with the exact-semantics softplus:
//   layer 2   h2 = softplus(A2_t * acc + C2_t)                    -> [hi | lo] FP16 halves of the next GEMM's operand
//   layer 3   h3 = softplus(A3_t * acc + C3_t);  part = h3 . W4^T  -> FP32 partial per 128-column slot (fixed-order sum
//                                                                    in the tail kernel: deterministic)
// Tiles are 128 x 128 (one per CTA at a time, persistent over the static schedule of plan_tiles' slim geometry).
#include "ladine_split.cuh"
#include "ladine_tensor.cuh"

namespace ladine {
namespace {

using namespace split;

template <int LAYER, int CP>
__global__ void __launch_bounds__(kThreads, 1) trunk_split_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sEpi = reinterpret_cast<float*>(smem + kStages * kStageBytes);   // scale[128] | shift[128] | W4[CP][128]
  constexpr int kEpiFloats = SBN * (2 + (LAYER == 3 ? CP : 0));
  Barriers* bars = reinterpret_cast<Barriers*>(sEpi + kEpiFloats);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    init_barriers(bars);
    tma_prefetch_desc(&p.tmA);
  }
  if (warp == 0) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int32_t* my_sched = p.sched + (size_t)blockIdx.x * p.sched_stride;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      Pipe ps;
      for (int it = 0;; ++it) {
        const int32_t code = __ldg(my_sched + it);
        if (code < 0) break;
        const TileCode tc(code);
        produce(smem, bars, ps, &p.tmA, &p.tmB[tc.member], tc.member * p.rows_pad + tc.mb * SBM, tc.nb * SBN, p.Fp, 0, p.KB);
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      Pipe ps;
      uint32_t chunk = 0;
      for (int it = 0;; ++it) {
        if (__ldg(my_sched + it) < 0) break;
        issue(smem, bars, ps, chunk, tmem_base, 0, p.KB);
      }
    }
  } else if (warp < 4) {
    const int et = threadIdx.x;   // 0..127
    const int quad = warp;
    float* sScale = sEpi;
    float* sShift = sEpi + SBN;
    float* sW4 = sEpi + 2 * SBN;
    uint32_t chunk = 0;
    for (int it = 0;; ++it) {
      const int32_t code = __ldg(my_sched + it);
      if (code < 0) break;
      const TileCode tc(code);
      const int member = tc.member, nb = tc.nb;
      {
        // this tile's per-column parameters (natural units in FP32X: the member's tables are not pre-multiplied by log2 e)
        const float s0 = __ldg(p.scale[member] + nb * SBN + et);
        const float h0 = __ldg(p.shift[member] + nb * SBN + et);
        float w4v[LAYER == 3 ? CP : 1];
        if (LAYER == 3) {
#pragma unroll
          for (int c = 0; c < CP; ++c) w4v[c] = __ldg(p.W4[member] + (size_t)c * p.Fp + nb * SBN + et);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the previous tile's readers are done
        sScale[et] = s0;
        sShift[et] = h0;
        if (LAYER == 3) {
#pragma unroll
          for (int c = 0; c < CP; ++c) sW4[c * SBN + et] = w4v[c];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      float acc[SBN];
      collect(bars, chunk, tmem_base, quad, lane, 0, p.KB, acc);

      const int row_m = tc.mb * SBM + quad * 32 + lane;   // row within the member
      const bool valid = row_m < p.rows;
      const size_t grow = (size_t)member * p.rows_pad + row_m;
      if (LAYER == 2) {
        __half* dst = reinterpret_cast<__half*>(p.h_out) + grow * (size_t)(2 * p.Fp) + nb * SBN;
#pragma unroll
        for (int q = 0; q < SBN / 16; ++q) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = 16 * q + 2 * i;
            const float v0 = softplus_precise(fmaf(sScale[j], acc[j], sShift[j]));
            const float v1 = softplus_precise(fmaf(sScale[j + 1], acc[j + 1], sShift[j + 1]));
            hi[i] = Pack16<__half>::pack(v0, v1);
            lo[i] = split_lo(hi[i], v0, v1);
          }
          if (valid) {
            st_global_256(dst + 16 * q, hi);
            st_global_256(dst + p.Fp + 16 * q, lo);
          }
        }
      } else {
        float eacc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) eacc[c] = 0.f;
#pragma unroll
        for (int j = 0; j < SBN; ++j) {
          const float v = softplus_precise(fmaf(sScale[j], acc[j], sShift[j]));
#pragma unroll
          for (int c = 0; c < CP; ++c) eacc[c] = fmaf(v, sW4[c * SBN + j], eacc[c]);
        }
        if (valid) {
          // 128-column slot `nb` of the row's 2 * NB slots (the tail kernel sums them in ascending order)
          float* dst = p.part + (grow * (size_t)(2 * p.NB) + nb) * CP;
#pragma unroll
          for (int c = 0; c < CP; ++c) dst[c] = eacc[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int LAYER, int CP>
cudaError_t launch_t(const GemmParams& p, int grid, cudaStream_t st) {
  const size_t smem = smem_bytes(sizeof(float) * SBN * (2 + (LAYER == 3 ? CP : 0)));
  auto kern = trunk_split_kernel<LAYER, CP>;
  static std::atomic<uint64_t> configured{0};
  cudaError_t e = configure_once(configured, [&] {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_split_gemm(int layer, const GemmParams& p, int grid, int Cp, cudaStream_t st) {
  if (layer == 2) return launch_t<2, 2>(p, grid, st);   // the layer-2 epilogue does not depend on the class count
  switch (Cp) {
    case 2: return launch_t<3, 2>(p, grid, st);
    case 4: return launch_t<3, 4>(p, grid, st);
    case 8: return launch_t<3, 8>(p, grid, st);
    case 16: return launch_t<3, 16>(p, grid, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace ladine
