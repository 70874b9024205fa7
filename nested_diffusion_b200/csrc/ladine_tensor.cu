// Tensor-core path of the LaDiNE sampler for the shipped shape (feature_dim = 4096): the two square
// ConditionalLinear layers of every reverse step run on tcgen05 (UTCHMMA) with FP32 accumulators in
// TMEM, operands staged by TMA, and everything around them fused into the epilogues.
//
// Per reverse step t (diffusion_utils.py:54-92 / :96-111 around latent_model.py:172-184), for ALL
// members x draws x images at once:
//   gemm<2>   h2 = softplus(A2_t * (h1 W2^T) + C2_t)                       -> 16-bit [rows, F]
//   gemm<3>   h3 = softplus(A3_t * (h2 W3^T) + C3_t); part = h3 . W4^T     -> FP32 [rows, F/256, Cp]
//   tailhead  eps = sum(part) + b4; y_{t-1} = posterior(y_t, eps, z);     (FP32, reference op order)
//             h1 = softplus(A1_{t-1} * (W1y y_{t-1} + u) + C1_{t-1}) * xf  -> 16-bit [rows, F]
// where A_l/C_l fold gamma_l[t], the Linear bias and the eval-mode BatchNorm.
//
// GEMM kernel: persistent, warp-specialised, one CTA per SM.
//   warps 0..3  epilogue       (tcgen05.ld 32x32b, scale/shift + softplus in the log2 domain, store; layer 3: fused
//                               lin4 partials and, once a row group is complete, the tail + head of the step)
//   warps 4..7  helpers        (layer 3: tail + head of the reverse step for row groups whose column tiles are all in)
//   warp 8      TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, mbarrier ring)
//   warp 9      MMA issuer     (tcgen05.mma kind::f16, cta_group::1 M=128 / cta_group::2 M=256, N=256 K=16, one thread)
//   TMEM: 2 accumulator stages x 256 columns so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tile geometries (all bit-identical, chosen per call by choose_ctas): 128 x 256 single-CTA tiles, 256 x 256 CTA-pair
// tiles (+ 2 x 64-row half tiles), slim 128 x 128 single-CTA tiles for calls too small to fill the SMs.  Tile order
// (plan_tiles): N-tile-major while a member's activations fit in L2, row-major beyond.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <vector>

#include "ladine_internal.cuh"
#include "ladine_tc.cuh"
#include "ladine_tensor.cuh"

namespace ladine {
namespace {

constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * BN;  // 512
constexpr int kGemmThreads = 320;
constexpr int kEpiThreads = 128;
// Warp roles.  The single-thread TMA producer and MMA issuer get the HIGHEST warp ids: each shares an SM sub-partition
// scheduler with one epilogue warp, and the arbiter favours the higher warp id, so a busy epilogue (fused tail + head)
// can never delay the instructions that keep the tensor pipe fed.
constexpr int kEpiWarps = kEpiThreads / 32;   // warps 0..3: epilogue (TMEM lane quadrant = warp id)
constexpr int kHelperWarp0 = 4;               // warps 4..7: fused tail + head jobs (layer 3), idle otherwise
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
// a member's activations above this size no longer survive in the 126 MB L2 between N tiles (with W and the
// layer's output also passing through): switch the tile order to row-major
constexpr size_t kRowMajorActBytes = (size_t)32 << 20;

// ------------------------------------------------------------------------------------------
// GEMM + fused epilogue
// ------------------------------------------------------------------------------------------
// pipeline geometry per CTA: CTAS = 1 -> A 128x64 + B 256x64 per stage; CTAS = 2 (cta_group::2, the
// CTA pair computes a 256x256 tile) -> A 128x64 + the CTA's half of B 128x64 per stage, so deeper ring
// BNT = tile width: 256, or 128 ("slim" single-CTA tiles for calls too small to fill the SMs with 256-wide tiles:
// twice as many tiles, each 64 KB instead of 96 KB of shared-memory traffic per K block)
template <int CTAS, int BNT = BN>
struct GemmCfg {
  static_assert(BNT == 256 || (BNT == 128 && CTAS == 1), "slim tiles are single-CTA only");
  static constexpr int kStages = (CTAS == 1 && BNT == 256) ? 4 : 6;
  static constexpr int kBRows = BNT / CTAS;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
};
constexpr int kMaxStages = 6;

struct __align__(8) GemmBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[kAccStages];
  uint64_t acc_empty[kAccStages];
  uint32_t tmem_base;
  int published;   // layer 3, fused: tiles of this CTA whose lin4 partials are out (epilogue -> helper warps)
};

// Finish reverse step th.t for the CTA's rows [row_base, row_base + nrows) of member k and (unless it is the last
// step) write their h1 of the next step for the 256 columns of N tile nb.  Called by the 128 epilogue threads.
//   tail : thread e < nrows owns row row_base + e: eps = b4 + sum of the 2 * NB column-slot partials, posterior
//          update (reference op order), y -> sYn (and, from the nb == 0 CTA, to y_next / trajectory / outputs);
//   head : thread e owns columns 2e, 2e+1 of the tile and walks the rows; per-column constants live in registers
//          and the per-image ones (u, xf) are refreshed when the image changes, so the shared-memory/L1 port -- the
//          resource the GEMM mainloop is bound by -- sees one broadcast LDS and one coalesced 128-byte store per row.
// Same arithmetic, in the same order, as tailhead_kernel: fused and unfused chains are bitwise identical.
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

template <typename T16, int CP>
__device__ __forceinline__ void fused_tail_head(const TailHeadParams& th, float* __restrict__ sYn_generic, int k, int nb, int NB,
                                                int row_base, int nrows, int rows_valid, size_t grow_base, int e,
                                                bool publish, bool do_head) {
  const int C = th.C;
  const uint32_t sYn = smem_u32(sYn_generic);   // explicit shared-window accesses (a generic LD costs ~3x an LDS)
  if (e < nrows && row_base + e < rows_valid) {
    const int row_m = row_base + e;
    const size_t grow = grow_base + e;
    const int n = row_m / th.D, d = row_m - n * th.D;
    // The row's partials are 2 * NB * CP contiguous floats written by other SMs: independent 128-bit loads through
    // L2 first (a dependent load-add chain would cost one L2 round trip per slot), then slot-ascending adds.
    float eps[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) eps[c] = c < C ? __ldg(th.b4[k] + c) : 0.f;
    {
      const float4* pp4 = reinterpret_cast<const float4*>(th.part + (grow * NB) * 2 * CP);
      const int n4 = 2 * NB * CP / 4;
      for (int i0 = 0; i0 < n4; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (i0 + j) < n4 ? __ldcg(pp4 + i0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if ((i0 + j) < n4) {
            const float e4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c = (4 * j + q) % CP;   // == (4 * (i0 + j) + q) % CP: i0 is a multiple of 8, CP divides 32
#pragma unroll
              for (int cc = 0; cc < CP; ++cc)
                if (cc == c) eps[cc] += e4[q];
            }
          }
        }
      }
    }
    float yn[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) yn[c] = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      if (c < C) {
        const float mu = __ldg(th.ytmean + ((size_t)k * th.N + n) * C + c);
        const float y = th.y_prev[grow * CP + c];
        float v;
        if (th.t > 0) {
          const float z = th.noise
                              ? __ldg(th.noise + ((((size_t)k * th.D + d) * th.n_slots + th.slot) * th.N + n) * C + c)
                              : philox_normal(th.seed, th.ids.chain(k, d, n), (uint32_t)th.slot, c);
          v = posterior_step_op(th.coef, y, mu, eps[c], z);
        } else {
          v = y0_reparam_op(th.coef, y, mu, eps[c]);
        }
        yn[c] = v;
        if (publish) {
          th.y_next[grow * CP + c] = v;
          if (th.traj_out && th.traj_entry >= 0)
            th.traj_out[((((size_t)k * th.D + d) * th.n_traj + th.traj_entry) * th.N + n) * C + c] = v;
          if (th.write_out) th.y_out[(((size_t)k * th.D + d) * th.N + n) * C + c] = v;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CP; ++c) sts_f32(sYn + (uint32_t)(e * CP + c) * 4u, yn[c]);
    if (publish && th.write_out && th.prob_out) {
      float mx = -INFINITY;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, -(yn[c] - 1.0f) * (yn[c] - 1.0f) / th.temperature);
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(-(yn[c] - 1.0f) * (yn[c] - 1.0f) / th.temperature - mx);
      const size_t o = (((size_t)k * th.D + d) * th.N + n) * C;
      for (int c = 0; c < C; ++c)
        th.prob_out[o + c] = expf(-(yn[c] - 1.0f) * (yn[c] - 1.0f) / th.temperature - mx) / sum;
    }
  }
  if (!do_head) return;
  asm volatile("bar.sync 2, 128;" ::: "memory");   // sYn complete (barrier 2 = the 128 helper threads)

  const int gcol = nb * BN + 2 * e;   // this thread's two features
  float a1[2], c1[2], pc[2][CP];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    a1[j] = __ldg(th.A1[k] + gcol + j);
    c1[j] = __ldg(th.C1[k] + gcol + j);
#pragma unroll
    for (int c = 0; c < CP; ++c) pc[j][c] = a1[j] * __ldg(th.W1y[k] + (size_t)(gcol + j) * CP + c);
  }
  const int last = min(nrows, rows_valid - row_base);   // rows of this CTA that exist
  int n = row_base / th.D, d = row_base - n * th.D;
  float q[2] = {0.f, 0.f}, xl[2] = {0.f, 0.f};
  bool fresh = true;
  uint32_t* hcol = reinterpret_cast<uint32_t*>(reinterpret_cast<T16*>(th.h1) + grow_base * th.Fp + gcol);
  const size_t hstride = (size_t)th.Fp / 2;   // uint32 (two 16-bit values) per row
#pragma unroll 4
  for (int r = 0; r < last; ++r) {
    if (fresh) {
      const float* urow = th.u + ((size_t)k * th.N + n) * th.Fp + gcol;
      const float* xfrow = th.xf + ((size_t)k * th.N + n) * th.Fin + gcol;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        q[j] = fmaf(a1[j], __ldg(urow + j), c1[j]);
        xl[j] = ((gcol + j) < th.Fin ? __ldg(xfrow + j) : 0.f) * kLn2;
      }
      fresh = false;
    }
    float hv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v2 = q[j];
#pragma unroll
      for (int c = 0; c < CP; ++c) v2 = fmaf(pc[j][c], lds_f32(sYn + (uint32_t)(r * CP + c) * 4u), v2);
      const float t = v2 > kSoftplusThreshold * kLog2e ? v2 : lg2_approx(1.0f + ex2_approx(v2));
      hv[j] = t * xl[j];
    }
    hcol[(size_t)r * hstride] = Pack16<T16>::pack(hv[0], hv[1]);   // a warp writes 128 contiguous bytes of the row
    if (++d == th.D) { d = 0; ++n; fresh = true; }
  }
}

template <int LAYER, typename T16, int CP, int CTAS, int BNT>
__global__ void __launch_bounds__(kGemmThreads, 1) trunk_gemm_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<CTAS, BNT>;
  constexpr bool kCanFuse = LAYER == 3 && CP <= 8 && BNT == BN;   // fused tail + head: 256-wide tiles only
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                    // kStages x 16 KiB
  uint8_t* sB = smem + kStages * Cfg::kABytes;           // kStages x 32 (16) KiB
  float* sEpi = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);  // scale[256] shift[256] (W4[CP][256])
  // layer 3: scale[256] | shift[256] | W4[CP][256] | y_next of this CTA's rows [128][CP] (fused tail + head)
  constexpr int kEpiFloats = LAYER == 3 ? BNT * (2 + CP) + BM * CP : 2 * BNT;
  GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(sEpi + kEpiFloats);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;   // position in the CTA pair; 0 issues the MMAs
  const int unit = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;  // tile-scheduling unit (CTA or pair)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      // one arrival: the (leader's) producer arrive.expect_tx; in a pair the peer's TMA bytes complete on the
      // leader's barrier too (its loads for round r+1 cannot start before phase r completed: they wait on
      // the multicast `empty` commit of the MMAs that consumed round r)
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(smem_u32(&bars->acc_full[s]), 1);
      mbar_init(smem_u32(&bars->acc_empty[s]), CTAS * (kEpiThreads / 32));
    }
    bars->published = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&p.tmA);
  }
  if (warp == 0) {
    if (CTAS == 2) tmem_alloc_pair(smem_u32(&bars->tmem_base), kTmemCols);
    else tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int32_t* my_sched = p.sched + (size_t)unit * p.sched_stride;

  if (warp == kProducerWarp) {
    // ===================== TMA producer (one per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const int32_t code = __ldg(my_sched + it);
        if (code < 0) break;
        const TileCode tc(code);
        // a half tile gives each CTA of the pair 64 rows; the 128-row box is still loaded (upper half unused)
        const int arow = tc.member * p.rows_pad + tc.mb * (BM * CTAS) + (int)rank * (tc.half ? BM / 2 : BM);
        const int brow = tc.nb * BNT + (int)rank * Cfg::kBRows;
        const CUtensorMap* tb = &p.tmB[tc.member];
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1u, 0);
          const uint32_t fb = smem_u32(&bars->full[stage]);
          if (CTAS == 1) {
            mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
            tma_load_2d(smem_u32(sA + stage * Cfg::kABytes), &p.tmA, fb, kb * BK, arow);
            tma_load_2d(smem_u32(sB + stage * Cfg::kBBytes), tb, fb, kb * BK, brow);
          } else {
            const uint32_t lfb = mapa_rank(fb, 0);  // both CTAs' bytes complete on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(fb, 2 * Cfg::kStageBytes);
            tma_load_2d_pair(smem_u32(sA + stage * Cfg::kABytes), &p.tmA, lfb, kb * BK, arow);
            tma_load_2d_pair(smem_u32(sB + stage * Cfg::kBBytes), tb, lfb, kb * BK, brow);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (single thread; leader CTA of a pair) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const int32_t code = __ldg(my_sched + it);
        if (code < 0) break;
        const uint32_t idesc = (code & 1) ? p.idesc_half : p.idesc;
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(smem_u32(&bars->acc_empty[as]), aphase ^ 1u, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BNT);
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(smem_u32(&bars->full[stage]), phase, 2);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * Cfg::kABytes));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * Cfg::kBBytes));
#pragma unroll
          for (int k4 = 0; k4 < BK / UK; ++k4) {
            // +32 bytes per K=16 slice inside the 128-byte swizzle row: +2 in the >>4 address field
            if (CTAS == 2)
              umma_f16_pair(d_tmem, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), idesc, (uint32_t)((kb | k4) != 0));
            else
              umma_f16(d_tmem, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), idesc, (uint32_t)((kb | k4) != 0));
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (CTAS == 2) umma_commit_pair(smem_u32(&bars->empty[stage]));
          else umma_commit(smem_u32(&bars->empty[stage]));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CTAS == 2) umma_commit_pair(smem_u32(&bars->acc_full[as]));
        else umma_commit(smem_u32(&bars->acc_full[as]));
      }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (warps 0..3), this CTA's rows x 256 columns =====================
    const int et = threadIdx.x;              // 0..127
    const int quad = warp;                   // TMEM lane quadrant this warp may access
    float* sScale = sEpi;
    float* sShift = sEpi + BNT;
    float* sW4 = sEpi + 2 * BNT;             // [CP][BNT]
    for (int it = 0;; ++it) {
      const int32_t code = __ldg(my_sched + it);
      if (code < 0) break;
      const TileCode tc(code);
      const int member = tc.member, nb = tc.nb;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      // stage this tile's per-column parameters (independent of the accumulator: issued before the wait)
      {
        const float* gs = p.scale[member] + nb * BNT;
        const float* gh = p.shift[member] + nb * BNT;
        constexpr int kHi = BNT == 256 ? 128 : 0;   // second column of this thread (none for 128-wide tiles)
        const float s0 = __ldg(gs + et), s1 = __ldg(gs + et + kHi);
        const float h0 = __ldg(gh + et), h1 = __ldg(gh + et + kHi);
        float w4v[LAYER == 3 ? 2 * CP : 1];
        if (LAYER == 3) {
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            w4v[2 * c] = __ldg(p.W4[member] + (size_t)c * p.Fp + nb * BNT + et);
            w4v[2 * c + 1] = __ldg(p.W4[member] + (size_t)c * p.Fp + nb * BNT + et + kHi);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers are done
        sScale[et] = s0; sScale[et + kHi] = s1;
        sShift[et] = h0; sShift[et + kHi] = h1;
        if (LAYER == 3) {
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            sW4[c * BNT + et] = w4v[2 * c];
            sW4[c * BNT + et + kHi] = w4v[2 * c + 1];
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(smem_u32(&bars->acc_full[as]), aphase, 3);
      tc_fence_after();

      // Accumulator layout (TMEM lane = datapath):
      //   full tile : lane l <-> row l of this CTA's 128 rows, TMEM columns 0..255 <-> output columns 0..255
      //   half tile : this CTA owns 64 rows; lanes 0..63 hold columns 0..127, lanes 64..127 columns 128..255
      //               of rows (lane & 63), both in TMEM columns 0..127 (the UMMA 2-SM M=128 "2x2" layout)
      const bool half = CTAS == 2 && tc.half;
      const int row_m = half ? tc.mb * (BM * CTAS) + (int)rank * (BM / 2) + (quad & 1) * 32 + lane
                             : tc.mb * (BM * CTAS) + (int)rank * BM + quad * 32 + lane;   // row within the member
      const bool valid = row_m < p.rows;
      const size_t grow = (size_t)member * p.rows_pad + row_m;
      const uint32_t taddr = tmem_base + (uint32_t)(as * BNT) + ((uint32_t)(quad * 32) << 16);
      // lin4 partials are kept per 128-column slot so that every tile geometry sums the same values in the
      // same order (tail kernel: fixed-order sum over 2 * NB slots) -> results do not depend on the geometry
      const int nslot = half ? 1 : BNT / 128;
      const int slot0 = half ? (quad >> 1) : 0;
#pragma unroll 1
      for (int sl = 0; sl < nslot; ++sl) {
        const int slot = slot0 + sl;
        float eacc[LAYER == 3 ? CP : 1];
#pragma unroll
        for (int c = 0; c < (LAYER == 3 ? CP : 1); ++c) eacc[c] = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          const int ocol = slot * 128 + ch * 32;                      // output column within the tile
          const uint32_t tcol = (uint32_t)((half ? 0 : sl * 128) + ch * 32);  // TMEM column within the stage
          uint32_t v[32];
          tmem_ld32(taddr + tcol, v);
          tmem_ld_wait();
          float hcol[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 sc = *reinterpret_cast<const float4*>(sScale + ocol + j4 * 4);
            const float4 sh = *reinterpret_cast<const float4*>(sShift + ocol + j4 * 4);
            hcol[4 * j4 + 0] = softplus_log2dom(fmaf(sc.x, __uint_as_float(v[4 * j4 + 0]), sh.x));
            hcol[4 * j4 + 1] = softplus_log2dom(fmaf(sc.y, __uint_as_float(v[4 * j4 + 1]), sh.y));
            hcol[4 * j4 + 2] = softplus_log2dom(fmaf(sc.z, __uint_as_float(v[4 * j4 + 2]), sh.z));
            hcol[4 * j4 + 3] = softplus_log2dom(fmaf(sc.w, __uint_as_float(v[4 * j4 + 3]), sh.w));
          }
          if (LAYER == 2) {
            if (valid) {
              // 32 consecutive 16-bit outputs of this row = 64 B = two full 32-byte sectors: 256-bit stores
              T16* dst = reinterpret_cast<T16*>(p.h_out) + grow * p.Fp + nb * BNT + ocol;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = Pack16<T16>::pack(hcol[16 * q + 2 * i], hcol[16 * q + 2 * i + 1]);
                st_global_256(dst + 16 * q, o);
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) {
              float e = eacc[c];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 w = *reinterpret_cast<const float4*>(sW4 + c * BNT + ocol + j4 * 4);
                e = fmaf(hcol[4 * j4 + 0], w.x, e);
                e = fmaf(hcol[4 * j4 + 1], w.y, e);
                e = fmaf(hcol[4 * j4 + 2], w.z, e);
                e = fmaf(hcol[4 * j4 + 3], w.w, e);
              }
              eacc[c] = e;
            }
          }
        }
        if (LAYER == 3 && valid) {
          // 128-column slot index within the row: 2 * NB slots, tile nb covers BNT / 128 of them
          float* dst = p.part + (grow * (size_t)(2 * p.NB) + nb * (BNT / 128) + slot) * CP;
#pragma unroll
          for (int c = 0; c < CP; ++c) dst[c] = eacc[c];
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer (in the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        const uint32_t eb = smem_u32(&bars->acc_empty[as]);
        if (CTAS == 2) mbar_arrive_cluster(mapa_rank(eb, 0));
        else mbar_arrive(eb);
      }
      if (kCanFuse) {
        if (p.fuse) {
          // publish: my partials are visible GPU-wide (fence) for all 128 rows (barrier) before the row group is
          // signalled; then tell this CTA's helper warps that job `it` exists
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            int* ctr = p.group_arrivals + ((size_t)member * p.mblk_total + tc.mb) * CTAS + rank;
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
            asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&bars->published)), "r"(it + 1) : "memory");
          }
        }
      }
    }
  } else if (warp < kHelperWarp0 + 4) {
    // ===================== helpers (warps 4..7): fused tail + head, decoupled from the accumulator pipeline ==========
    // Job j = scheduled tile j of this CTA.  It can run once (a) this CTA's epilogue has published tile j and (b) all
    // NB column tiles of the row group have arrived.  Arrivals never wait for a helper, so the blocking waits below
    // cannot dead-lock, whatever the order in which CTAs become resident.
    if (kCanFuse) {
      if (p.fuse) {
        const int ht = threadIdx.x - kHelperWarp0 * 32;   // 0..127
        float* sYn = sEpi + BNT * (2 + CP);
        for (int job = 0;; ++job) {
          const int32_t code = __ldg(my_sched + job);
          if (code < 0) break;
          if (ht == 0) {
            const TileCode tc(code);
            const int* ctr = p.group_arrivals + ((size_t)tc.member * p.mblk_total + tc.mb) * CTAS + rank;
            const long long t0 = clock64();
            int pub = 0, seen = 0;
            // relaxed polls with back-off (an acquire load at GPU scope invalidates the SM's L1 on every poll and the
            // spinning thread would steal issue slots from the epilogue); one acquire fence once the data is there
            for (;;) {
              asm volatile("ld.relaxed.cta.shared.s32 %0, [%1];" : "=r"(pub) : "r"(smem_u32(&bars->published)) : "memory");
              if (pub > job) {
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
                if (seen >= p.arrivals_target) break;
              }
              if (clock64() - t0 > 8000000000LL) {
                printf("ladine: helper wait timeout block=%d job=%d published=%d seen=%d target=%d\n", (int)blockIdx.x, job,
                       pub, seen, p.arrivals_target);
                __trap();
              }
              __nanosleep(500);
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
          const TileCode tc(code);
          const bool half = CTAS == 2 && tc.half;
          const int row_base = tc.mb * (BM * CTAS) + (int)rank * (half ? BM / 2 : BM);
          fused_tail_head<T16, CP>(p.th, sYn, tc.member, tc.nb, p.NB, row_base, half ? BM / 2 : BM, p.rows,
                                   (size_t)tc.member * p.rows_pad + row_base, ht, /*publish=*/tc.nb == 0, p.do_head != 0);
          asm volatile("bar.sync 2, 128;" ::: "memory");   // sYn is rewritten by the next job
        }
      }
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer's smem/TMEM stay live until both are done
  if (warp == 0) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// tail (posterior update of step t) + head (lin1 of step t-1)
// ------------------------------------------------------------------------------------------
constexpr int kTailThreads = 128;
constexpr int kTailMaxRows = 32;  // draws handled per CTA

// VEC = features per thread: 4 by default (40 registers -> 12 CTAs = 48 warps per SM; measured 376-410 us per launch at
// config 3 against 456-476 us for VEC = 8 at 64 registers / 8 CTAs), 8 only when forced with the "tail_vec" option.
template <int MODE, typename T16, int CP, int VEC>
__global__ void __launch_bounds__(kTailThreads, (CP <= 4 ? (VEC == 4 ? 12 : 8) : 1)) tailhead_kernel(const __grid_constant__ TailHeadParams p) {
  constexpr int kTailCols = kTailThreads * VEC;   // features per CTA
  __shared__ float sY[kTailMaxRows * CP];
  // grid.x = N * colsplit: CTA (n, cs) produces features [cs * kTailCols, (cs+1) * kTailCols) of image n's draws.
  // Every column split recomputes the (tiny) tail; split 0 alone publishes y / outputs.
  const int n = blockIdx.x / p.colsplit, cs = blockIdx.x - n * p.colsplit, k = blockIdx.y;
  const bool publish = cs == 0;
  const int d0 = blockIdx.z * p.dchunk;
  const int nd = min(p.dchunk, p.D - d0);
  const int tid = threadIdx.x;
  const int C = p.C;

  // ---------------- tail: finish step t for rows (k, n, d0..d0+nd) ----------------
  if (CP != C)
    for (int i = tid; i < nd * CP; i += kTailThreads) sY[i] = 0.f;   // padded classes read as 0 by the head
  if (CP != C) __syncthreads();
  for (int i = tid; i < nd * C; i += kTailThreads) {
    const int dl = i / C, c = i - dl * C;
    const int d = d0 + dl;
    const size_t grow = (size_t)k * p.rows_pad + (size_t)n * p.D + d;
    const float mu = __ldg(p.ytmean + ((size_t)k * p.N + n) * C + c);
    float yn;
    if (MODE == kInit) {
      if (p.y_init) {
        yn = __ldg(p.y_init + (((size_t)k * p.D + d) * p.N + n) * C + c);
      } else {
        const float z = p.noise ? __ldg(p.noise + ((((size_t)k * p.D + d) * p.n_slots + 0) * p.N + n) * C + c)
                                : philox_normal(p.seed, p.ids.chain(k, d, n), 0u, c);
        yn = __fadd_rn(z, mu);
      }
    } else {
      float eps = __ldg(p.b4[k] + c);
      const float* pp = p.part + (grow * p.NB) * 2 * CP + c;
      for (int sl = 0; sl < 2 * p.NB; ++sl) eps += pp[sl * CP];  // fixed order over 128-column slots: deterministic
      const float y = p.y_prev[grow * CP + c];
      if (p.t > 0) {
        const float z = p.noise ? __ldg(p.noise + ((((size_t)k * p.D + d) * p.n_slots + p.slot) * p.N + n) * C + c)
                                : philox_normal(p.seed, p.ids.chain(k, d, n), (uint32_t)p.slot, c);
        yn = posterior_step_op(p.coef, y, mu, eps, z);
      } else {
        yn = y0_reparam_op(p.coef, y, mu, eps);
      }
    }
    sY[dl * CP + c] = yn;
    if (publish) {
      p.y_next[grow * CP + c] = yn;
      if (p.traj_out && p.traj_entry >= 0)
        p.traj_out[((((size_t)k * p.D + d) * p.n_traj + p.traj_entry) * p.N + n) * C + c] = yn;
      if (p.write_out) p.y_out[(((size_t)k * p.D + d) * p.N + n) * C + c] = yn;
    }
  }
  if (MODE == kFinal && !(p.write_out && p.prob_out)) return;
  __syncthreads();
  if (publish && p.write_out && p.prob_out) {
    // softmax(-(y-1)^2 / temperature) -- classification_train_separately.py:392-398
    for (int dl = tid; dl < nd; dl += kTailThreads) {
      float mx = -INFINITY;
      for (int c = 0; c < C; ++c) {
        const float y = sY[dl * CP + c];
        mx = fmaxf(mx, -(y - 1.0f) * (y - 1.0f) / p.temperature);
      }
      float sum = 0.f;
      for (int c = 0; c < C; ++c) {
        const float y = sY[dl * CP + c];
        sum += expf(-(y - 1.0f) * (y - 1.0f) / p.temperature - mx);
      }
      const size_t o = (((size_t)k * p.D + d0 + dl) * p.N + n) * C;
      for (int c = 0; c < C; ++c) {
        const float y = sY[dl * CP + c];
        p.prob_out[o + c] = expf(-(y - 1.0f) * (y - 1.0f) / p.temperature - mx) / sum;
      }
    }
  }
  if (MODE == kFinal) return;

  // ---------------- head: h1 of the next step for the same rows ----------------
  // thread owns VEC consecutive features; their per-column constants stay in registers across draws
  const float* xfrow = p.xf + ((size_t)k * p.N + n) * p.Fin;
  const float* urow = p.u + ((size_t)k * p.N + n) * p.Fp;
  for (int f0 = cs * kTailCols + tid * VEC; f0 < p.Fp; f0 += p.colsplit * kTailCols) {
    // per-column constants of this step, folded once per CTA:  v2 = sum_c pc[c] * y[c] + q  (log2 domain),
    // h1 = softplus(v2 / log2e) * xf = t * xl  with  t = lg2(1 + 2^v2) (or v2 itself when large), xl = xf * ln2
    float q[VEC], xl[VEC], pc[VEC][CP];
    {
      float a1[VEC], c1[VEC], uu[VEC];
#pragma unroll
      for (int v = 0; v < VEC / 4; ++v) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.A1[k] + f0) + v);
        a1[4 * v + 0] = t0.x; a1[4 * v + 1] = t0.y; a1[4 * v + 2] = t0.z; a1[4 * v + 3] = t0.w;
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.C1[k] + f0) + v);
        c1[4 * v + 0] = q0.x; c1[4 * v + 1] = q0.y; c1[4 * v + 2] = q0.z; c1[4 * v + 3] = q0.w;
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(urow + f0) + v);
        uu[4 * v + 0] = u0.x; uu[4 * v + 1] = u0.y; uu[4 * v + 2] = u0.z; uu[4 * v + 3] = u0.w;
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        // 16-bit path: tables are pre-multiplied by log2(e), so softplus = lg2(1 + 2^v2) * ln2 and ln2 rides on xf;
        // FP32X: natural units and the exact-semantics softplus
        xl[j] = ((f0 + j) < p.Fin ? __ldg(xfrow + f0 + j) : 0.f) * (p.split ? 1.0f : kLn2);
        q[j] = fmaf(a1[j], uu[j], c1[j]);
#pragma unroll
        for (int c = 0; c < CP; ++c) pc[j][c] = a1[j] * __ldg(p.W1y[k] + (size_t)(f0 + j) * CP + c);  // padded classes: 0
      }
    }
    T16* hrow = reinterpret_cast<T16*>(p.h1) + ((size_t)k * p.rows_pad + (size_t)n * p.D + d0) * p.ld_h1 + f0;
    for (int dl = 0; dl < nd; ++dl, hrow += p.ld_h1) {
      float yv[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) yv[c] = sY[dl * CP + c];   // padded classes hold 0
      float hv[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float v2 = q[j];
#pragma unroll
        for (int c = 0; c < CP; ++c) v2 = fmaf(pc[j][c], yv[c], v2);
        const float t = p.split ? softplus_precise(v2)
                                : (v2 > kSoftplusThreshold * kLog2e ? v2 : lg2_approx(1.0f + ex2_approx(v2)));
        hv[j] = t * xl[j];
      }
      if (VEC == 8) {
        uint4 o;
        o.x = Pack16<T16>::pack(hv[0], hv[1]);
        o.y = Pack16<T16>::pack(hv[2], hv[3]);
        o.z = Pack16<T16>::pack(hv[VEC - 4], hv[VEC - 3]);
        o.w = Pack16<T16>::pack(hv[VEC - 2], hv[VEC - 1]);
        *reinterpret_cast<uint4*>(hrow) = o;
        if (p.split) {
          uint4 l;
          l.x = split_lo(o.x, hv[0], hv[1]);
          l.y = split_lo(o.y, hv[2], hv[3]);
          l.z = split_lo(o.z, hv[VEC - 4], hv[VEC - 3]);
          l.w = split_lo(o.w, hv[VEC - 2], hv[VEC - 1]);
          *reinterpret_cast<uint4*>(hrow + p.Fp) = l;
        }
      } else {
        uint2 o;
        o.x = Pack16<T16>::pack(hv[0], hv[1]);
        o.y = Pack16<T16>::pack(hv[2], hv[3]);
        *reinterpret_cast<uint2*>(hrow) = o;
        if (p.split) {
          uint2 l;
          l.x = split_lo(o.x, hv[0], hv[1]);
          l.y = split_lo(o.y, hv[2], hv[3]);
          *reinterpret_cast<uint2*>(hrow + p.Fp) = l;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
}  // namespace

// ---- tensor-map helpers (external linkage: shared with ladine_split.cu / ladine_encoder.cu) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t resolve_encode(ladine_handle* h, std::string* err) {
  if (h->encode_tiled) return cudaSuccess;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    *err = "cuTensorMapEncodeTiled is not available from the driver";
    return e != cudaSuccess ? e : cudaErrorNotSupported;
  }
  h->encode_tiled = fn;
  return cudaSuccess;
}

// 2-D K-major tensor map: inner dim = cols (contiguous), outer = rows; box = 64 x box_rows; 128B swizzle
bool make_tmap(ladine_handle* h, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
               bool bf16, std::string* err) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
      out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims,
      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, box_rows);
    *err = buf;
    return false;
  }
  return true;
}

namespace {

// instruction descriptor for kind::f16 (PTX ISA "Instruction descriptor"): D format F32 (1) at [4,6);
// A/B format (0 = F16, 1 = BF16) at [7,10)/[10,13); A,B K-major (0) at 15/16; N>>3 at [17,23); M>>4 at [24,29)
uint32_t make_idesc(bool bf16, int ctas, int n = BN) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((BM * ctas) >> 4) << 24);
}

template <int LAYER, typename T16, int CP, int CTAS, int BNT = BN>
cudaError_t launch_gemm_t(const GemmParams& p, int grid, cudaStream_t st) {
  const size_t smem = tensor_gemm_smem_bytes(CP);
  auto kern = trunk_gemm_kernel<LAYER, T16, CP, CTAS, BNT>;
  // once per (kernel instantiation, device): the call costs several microseconds of host time, which a chain of
  // thousands of small launches cannot afford (a small call is bound by the host's launch rate, not by the GPU)
  static std::atomic<uint64_t> configured{0};
  cudaError_t e = configure_once(configured, [&] {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (CTAS == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

// ctas: 1 = single-CTA 128 x 256 tiles, 2 = CTA pairs, kGeomSlim = single-CTA 128 x 128 tiles
constexpr int kGeomSlim = 3;

template <int LAYER, typename T16>
cudaError_t launch_gemm_c(const GemmParams& p, int grid, int Cp, int ctas, cudaStream_t st) {
  if (LAYER == 2) {  // the layer-2 epilogue does not depend on the class count
    if (ctas == kGeomSlim) return launch_gemm_t<2, T16, 2, 1, 128>(p, grid, st);
    return ctas == 2 ? launch_gemm_t<2, T16, 2, 2>(p, grid, st) : launch_gemm_t<2, T16, 2, 1>(p, grid, st);
  }
#define LADINE_GEMM_CASE(CPV)                                                                          \
  case CPV:                                                                                            \
    if (ctas == kGeomSlim) return launch_gemm_t<3, T16, CPV, 1, 128>(p, grid, st);                \
    return ctas == 2 ? launch_gemm_t<3, T16, CPV, 2>(p, grid, st) : launch_gemm_t<3, T16, CPV, 1>(p, grid, st);
  switch (Cp) {
    LADINE_GEMM_CASE(2)
    LADINE_GEMM_CASE(4)
    LADINE_GEMM_CASE(8)
    LADINE_GEMM_CASE(16)
  }
#undef LADINE_GEMM_CASE
  return cudaErrorInvalidValue;
}

template <int LAYER>
cudaError_t launch_gemm(const GemmParams& p, int grid, bool bf16, int Cp, int ctas, cudaStream_t st) {
  return bf16 ? launch_gemm_c<LAYER, __nv_bfloat16>(p, grid, Cp, ctas, st)
              : launch_gemm_c<LAYER, __half>(p, grid, Cp, ctas, st);
}

template <int MODE, typename T16, int CP, int VEC>
cudaError_t launch_tail_t(const TailHeadParams& p, dim3 grid, cudaStream_t st) {
  auto kern = tailhead_kernel<MODE, T16, CP, VEC>;
  // same shared-memory carveout as the GEMM kernels, so the SMs are not reconfigured at every kernel boundary
  static std::atomic<uint64_t> configured{0};
  cudaError_t e = configure_once(configured, [&] {
    return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  });
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTailThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int MODE>
cudaError_t launch_tail(const TailHeadParams& p, dim3 grid, bool bf16, int Cp, int vec, cudaStream_t st) {
#define LADINE_TAIL_CASE(CPV)                                                                                     \
  case CPV:                                                                                                       \
    if (vec == 4)                                                                                                 \
      return bf16 ? launch_tail_t<MODE, __nv_bfloat16, CPV, 4>(p, grid, st)                                  \
                  : launch_tail_t<MODE, __half, CPV, 4>(p, grid, st);                                        \
    return bf16 ? launch_tail_t<MODE, __nv_bfloat16, CPV, 8>(p, grid, st)                                    \
                : launch_tail_t<MODE, __half, CPV, 8>(p, grid, st);
  switch (Cp) {
    LADINE_TAIL_CASE(2)
    LADINE_TAIL_CASE(4)
    LADINE_TAIL_CASE(8)
    LADINE_TAIL_CASE(16)
  }
#undef LADINE_TAIL_CASE
  return cudaErrorInvalidValue;
}

struct ProfSpan {
  ladine_handle* h;
  cudaStream_t st;
  cudaEvent_t a = nullptr;
  int kind;
  ProfSpan(ladine_handle* h_, cudaStream_t st_, int kind_) : h(h_), st(st_), kind(kind_) {
    if (!h->profiling) return;
    a = take();
    cudaEventRecord(a, st);
  }
  ~ProfSpan() {
    if (!a) return;
    cudaEvent_t b = take();
    cudaEventRecord(b, st);
    h->spans.push_back({a, b, kind});
  }
  cudaEvent_t take() {
    cudaEvent_t e;
    if (!h->pool.empty()) {
      e = h->pool.back();
      h->pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
};

}  // namespace

size_t tensor_gemm_smem_bytes(int Cp) {
  static_assert(GemmCfg<1>::kStages * GemmCfg<1>::kStageBytes == GemmCfg<2>::kStages * GemmCfg<2>::kStageBytes &&
                    GemmCfg<1>::kStages * GemmCfg<1>::kStageBytes == GemmCfg<1, 128>::kStages * GemmCfg<1, 128>::kStageBytes,
                "all pipeline geometries use the same ring size");
  return 1024 /*alignment slack*/ + (size_t)GemmCfg<1>::kStages * GemmCfg<1>::kStageBytes +
         sizeof(float) * (BN * (2 + Cp) + BM * Cp) + sizeof(GemmBarriers);
}

// 1 = cta_group::1 tiles of 128 rows, 2 = CTA pairs (cta_group::2) on 256-row tiles.  Pairs halve the
// B-operand shared-memory/L2 traffic per SM but pad each member's rows to a multiple of 256.

// ------------------------------------------------------------------------------------------
// static tile schedule
// ------------------------------------------------------------------------------------------
struct TilePlan {
  int ctas = 1;
  int tile_rows = BM;      // rows per full tile (128 * ctas)
  int n_full = 0;          // full row tiles per member
  int has_half = 0;        // pair mode: the member's last <= 128 rows run as a half tile (2 x 64 rows, M=128 MMAs)
  int rows_pad = 0;        // row stride between members
  int units = 0;           // CTAs (ctas == 1) or CTA pairs (ctas == 2) that get work
  int stride = 0;          // table row length (entries per unit incl. the -1 terminator)
  std::vector<int32_t> table;
};

// Tiles in canonical order (member, N tile, row tile: the row tiles sharing a W tile are adjacent, so they
// run concurrently and the W tile is fetched from HBM once).  Costs: full = 20, half = 17 -- measured: the
// mainloop is bound by the 128 B/clk shared-memory port (TMA writes + UMMA operand reads), and a half tile
// moves 72 KB per K block against 80 KB for a full pair tile, so it is only ~15 % cheaper although it does
// half the math.  Per-unit quotas come from longest-processing-time assignment; the canonical sequence is
// then dealt to the least-loaded unit that still has quota for that tile kind, which keeps neighbours in time.
// row_major: (member, row tile, N tile) -- the NB column tiles of a row group are adjacent, i.e. run at the same
// time on NB different units (needed by the fused tail + head, which waits for the whole group); otherwise
// (member, N tile, row tile).
TilePlan plan_tiles(int K, int rows, int NB, int ctas, int max_units, bool row_major = false) {
  TilePlan tp;
  tp.ctas = ctas;
  tp.tile_rows = BM * ctas;
  tp.n_full = rows / tp.tile_rows;
  const int rem = rows - tp.n_full * tp.tile_rows;
  tp.has_half = (ctas == 2 && rem > 0 && rem <= BM) ? 1 : 0;
  if (rem > 0 && !tp.has_half) tp.n_full += 1;
  tp.rows_pad = (tp.n_full + tp.has_half) * tp.tile_rows;
  const int n_items = K * NB * (tp.n_full + tp.has_half);
  tp.units = n_items < max_units ? n_items : max_units;
  const int U = tp.units;
  // LPT quotas
  std::vector<int> load(U, 0), qfull(U, 0), qhalf(U, 0);
  auto least = [&](const std::vector<int>& need) {
    int best = -1;
    for (int u = 0; u < U; ++u)
      if (need.empty() || need[u] > 0)
        if (best < 0 || load[u] < load[best]) best = u;
    return best;
  };
  const std::vector<int> none;
  constexpr int kCostFull = 20, kCostHalf = 17;
  for (int i = 0; i < K * NB * tp.n_full; ++i) { const int u = least(none); load[u] += kCostFull; qfull[u] += 1; }
  for (int i = 0; i < K * NB * tp.has_half; ++i) { const int u = least(none); load[u] += kCostHalf; qhalf[u] += 1; }
  int longest = 0;
  for (int u = 0; u < U; ++u) longest = (qfull[u] + qhalf[u]) > longest ? (qfull[u] + qhalf[u]) : longest;
  tp.stride = longest + 1;
  tp.table.assign((size_t)U * tp.stride, -1);
  std::vector<int> fill(U, 0);
  std::fill(load.begin(), load.end(), 0);
  const int MB = tp.n_full + tp.has_half;
  for (int k = 0; k < K; ++k)
    for (int o = 0; o < NB * MB; ++o) {
      const int nb = row_major ? o % NB : o / MB;
      const int mb = row_major ? o / NB : o % MB;
      const int half = (tp.has_half && mb == tp.n_full) ? 1 : 0;
      const int u = least(half ? qhalf : qfull);
      (half ? qhalf : qfull)[u] -= 1;
      load[u] += half ? kCostHalf : kCostFull;
      tp.table[(size_t)u * tp.stride + fill[u]++] = TileCode::pack(k, nb, mb, half);
    }
  return tp;
}

size_t sched_bytes_bound(int K, int rows, int NB) {
  // generous bound for either geometry: every unit's row is at most ceil(items / units) + 2 entries long
  const size_t items = (size_t)K * NB * ((rows + BM - 1) / BM + 1);
  return (items + 148 * 4) * sizeof(int32_t) * 2;
}

int choose_ctas(const ladine_handle* h, int K, int rows, int Fp) {
  // slim 128 x 128 tiles: forced ("ctas" = 3) or, on auto, when even twice as many tiles still fit the SMs in one
  // round (measured at one member x 64 rows: 31.2 -> 27.5 us per layer; with 80 wide tiles -- two rounds of slim
  // ones -- 45.6 -> 54.9 us, hence the one-round condition)
  const long long wide_tiles = (long long)K * ((rows + BM - 1) / BM) * (Fp / BN);
  if (h->ctas == kGeomSlim || (h->ctas == 0 && 2 * wide_tiles <= (long long)h->sm_count)) return kGeomSlim;
  if (h->ctas == 1 || h->ctas == 2) return h->ctas;
  // Makespan of the static schedule in single-CTA 128-row tile times, for both geometries on THIS many SMs (the tile
  // counts quantise differently: one member x 1400 rows is 176 single tiles = 2 rounds on 148 SMs, but 80 pair tiles +
  // 16 half tiles on 74 SM pairs = 2 pair rounds, each 1 / pair_gain as long -- while K = 5 x 1400 rows is 5.95 rounds
  // of single tiles against 6.7 pair rounds).  A full pair tile (256 rows on 2 SMs) takes 1 / pair_gain, a half tile
  // 0.85 / pair_gain (costs 20 : 17, as in plan_tiles, whose longest-processing-time deal is replayed here on counts).
  // Pairs must win by 3 % to be chosen.
  const long long per_col = (long long)K * (Fp / BN);
  const int units2 = h->sm_count / 2 > 0 ? h->sm_count / 2 : 1;
  const int full2 = rows / 256, rem2 = rows - full2 * 256;
  const long long n_full = per_col * (full2 + (rem2 > 128 ? 1 : 0)), n_half = per_col * (rem2 > 0 && rem2 <= 128 ? 1 : 0);
  // full tiles dealt evenly (the first n_full % units2 units carry one more), half tiles to the least-loaded unit
  std::vector<long long> load(units2);
  for (int u = 0; u < units2; ++u) load[u] = (n_full / units2 + (u < n_full % units2 ? 1 : 0)) * 20;
  for (long long i = 0; i < n_half; ++i) *std::min_element(load.begin(), load.end()) += 17;
  const long long max_load = *std::max_element(load.begin(), load.end());
  const double span2 = (double)max_load / 20.0 / h->pair_gain;
  const double span1 = (double)((wide_tiles + h->sm_count - 1) / h->sm_count);
  return span2 * 1.03 < span1 ? 2 : 1;
}

// One lane = one group of members advancing through the reverse steps on its own stream.  Lanes are
// independent chains, so running them on separate streams lets the hardware hand SMs from one lane's
// GEMM CTAs to the next lane's as they retire (no inter-kernel drain/launch gap) and hides the
// tail/head kernel of one lane under the GEMMs of the other.
struct TensorChain {
  ladine_handle* h;
  const ladine_member* const* members;
  ladine_sample_args a;
  const StepCoef* h_coef;
  cudaStream_t st;
  GemmParams g2{}, g3{};
  TailHeadParams tp{};
  dim3 tgrid;
  int grid = 0, K = 0, Fp = 0, Cp = 0, slot_base = 0, ctas = 1, ycur = 0, g3_launches = 0, tail_vec = 8;
  bool fuse = false, split = false;
  float* ybuf[2] = {nullptr, nullptr};
  bool bf16 = false;

  void set_head_rows(int t_next) {
    for (int k = 0; k < K; ++k) {
      tp.A1[k] = members[k]->A[0] + (size_t)t_next * Fp;
      tp.C1[k] = members[k]->Cc[0] + (size_t)t_next * Fp;
    }
  }

  cudaError_t init(ladine_handle* h_, const ladine_member* const* members_, const ladine_sample_args& a_,
                   const ChainIds& ids, const StepCoef* h_coef_, const TensorWorkspace& ws, int n_slots, int n_traj,
                   cudaStream_t st_, bool single_lane, int64_t* launches, std::string* err) {
    h = h_;
    members = members_;
    a = a_;
    h_coef = h_coef_;
    st = st_;
    const ladine_member* m0 = members[0];
    bf16 = m0->precision == LADINE_PREC_BF16;
    split = m0->split != 0;
    Fp = m0->Fp;
    Cp = m0->Cp;
    K = a.K;
    const int rows = a.N * a.D;
    ctas = split ? kGeomSlim : choose_ctas(h, K, rows, Fp);   // FP32X: 128 x 128 tiles (ladine_split.cu)
    const bool slim = ctas == kGeomSlim;
    const int cpu = slim ? 1 : ctas;            // CTAs per scheduling unit (2 for pairs)
    const int bnt = slim ? 128 : BN;            // tile width
    // the fused tail + head spins on other CTAs of the same launch: every CTA must be resident, so it is only
    // used when this chain is the sole lane, and its extra per-column parameters fit in smem up to 8 classes
    fuse = h->fuse && single_lane && Cp <= 8 && !slim && !split;
    const int NBt = Fp / bnt;                   // column tiles per row tile
    if ((rows + BM - 1) / BM > 4096 || NBt > 1024) {
      *err = "too many rows per member for one launch group (max 524288 chains): tile the images (NestedEnsemble does)";
      return cudaErrorInvalidValue;
    }
    // Tile order.  N-tile-major keeps one 2 MiB W tile hot while all row tiles of a member stream past it: right while
    // a member's activations (rows x Fp 16-bit) stay in L2, but beyond that every N tile re-reads them from HBM
    // (F/256 = 16 passes at the shipped width).  Row-major runs the F/256 column tiles of a few row tiles together:
    // the activations are read once and the member's 32 MiB W cycles through L2.  "order" option: 0 auto, 1 / 2 force.
    const size_t act_bytes = (size_t)((rows + BM - 1) / BM) * BM * Fp * 2 * (split ? 2 : 1);
    const bool row_major = h->order == 2 || (h->order == 0 && act_bytes > kRowMajorActBytes);
    const TilePlan plan = plan_tiles(K, rows, NBt, cpu, h->sm_count / cpu, row_major);
    const TilePlan plan3 = (fuse && !row_major) ? plan_tiles(K, rows, NBt, cpu, h->sm_count / cpu, /*row_major=*/true) : plan;
    const int rows_pad = plan.rows_pad;
    const size_t m_total = (size_t)K * rows_pad;
    const size_t sched_half = sched_bytes_bound(K, rows, NBt) / (2 * sizeof(int32_t));  // ints per table
    int32_t* sched3 = ws.sched + sched_half;

    cudaError_t e = resolve_encode(h, err);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(ws.sched, plan.table.data(), plan.table.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(sched3, plan3.table.data(), plan3.table.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    const int mblk_total = plan.n_full + plan.has_half;
    if (e == cudaSuccess && fuse)
      e = cudaMemsetAsync(ws.arrivals, 0, sizeof(int) * (size_t)K * mblk_total * cpu, st);
    if (e != cudaSuccess) return e;
    const uint64_t ld = (uint64_t)Fp * (split ? 2 : 1);   // FP32X: [hi | lo] halves side by side
    if (!make_tmap(h, &g2.tmA, ws.h1, m_total, ld, BM, bf16, err)) return cudaErrorInvalidValue;
    if (!make_tmap(h, &g3.tmA, ws.h2, m_total, ld, BM, bf16, err)) return cudaErrorInvalidValue;
    for (int k = 0; k < K; ++k) {
      if (!make_tmap(h, &g2.tmB[k], members[k]->W2h, Fp, ld, bnt / cpu, bf16, err)) return cudaErrorInvalidValue;
      if (!make_tmap(h, &g3.tmB[k], members[k]->W3h, Fp, ld, bnt / cpu, bf16, err)) return cudaErrorInvalidValue;
      g3.W4[k] = members[k]->W4;
    }
    for (GemmParams* g : {&g2, &g3}) {
      g->Fp = Fp;
      g->NB = Fp / BN;   // 256-column groups: the lin4 partial buffer has 2 * NB slots per row whatever the tile width
      g->KB = Fp / BK;
      g->rows = rows;
      g->rows_pad = rows_pad;
      g->sched = ws.sched;
      g->sched_stride = plan.stride;
      g->idesc = make_idesc(bf16, cpu, bnt);
      g->idesc_half = make_idesc(bf16, 1);  // M = 128 across the pair
      g->fuse = 0;
    }
    g3.sched = sched3;
    g3.sched_stride = plan3.stride;
    g3.fuse = fuse ? 1 : 0;
    g3.mblk_total = mblk_total;
    g3.group_arrivals = ws.arrivals;
    g2.h_out = ws.h2;
    g3.part = ws.part;
    grid = plan.units * cpu;

    for (int k = 0; k < K; ++k) {
      tp.W1y[k] = members[k]->W1y;
      tp.b4[k] = members[k]->b4;
    }
    tp.part = ws.part;
    ybuf[0] = ws.ybuf;
    ybuf[1] = ws.ybuf + m_total * Cp;
    ycur = 0;
    tp.xf = a.xf;
    tp.u = ws.u;
    tp.ytmean = a.ytmean;
    tp.y_init = a.y_init;
    tp.noise = a.noise;
    tp.h1 = ws.h1;
    tp.y_out = a.y_out;
    tp.traj_out = a.traj_out;
    tp.prob_out = a.prob_out;
    tp.temperature = a.prob_out ? a.temperature : 1.0f;
    tp.seed = a.seed;
    tp.ids = ids;
    tp.N = a.N;
    tp.D = a.D;
    tp.C = m0->C;
    tp.Fin = m0->F;
    tp.Fp = Fp;
    tp.ld_h1 = (int)ld;
    tp.split = split ? 1 : 0;
    tp.NB = Fp / BN;
    tp.rows_pad = rows_pad;
    tp.n_slots = n_slots;
    tp.n_traj = n_traj;
    tp.dchunk = a.D < kTailMaxRows ? a.D : kTailMaxRows;
    {
      // features per thread: 4 unless forced ("tail_vec").  Measured: 412-417 us against 472-476 us per launch with 8 at
      // config 3 (102 400 rows; fewer registers -> more resident warps for an issue/latency-bound loop), 42.4 vs 43.9 us
      // at config 2, 8.5 vs 10.5 us at config 1 -- 8 never won, so the round-1 wave-quantisation model that chose it for
      // large grids is gone.
      tail_vec = h->tail_vec == 8 ? 8 : 4;
      tp.colsplit = (Fp + kTailThreads * tail_vec - 1) / (kTailThreads * tail_vec);
    }
    tgrid = dim3(a.N * tp.colsplit, K, (a.D + tp.dchunk - 1) / tp.dchunk);
    slot_base = a.y_init ? 0 : 1;

    // y_T (or the caller's y) and h1 for the first step
    set_head_rows(a.t_first);
    tp.slot = 0;
    tp.traj_entry = a.y_init ? -1 : 0;
    tp.write_out = 0;
    tp.y_prev = ybuf[ycur];
    tp.y_next = ybuf[ycur ^ 1];
    ycur ^= 1;
    e = launch_tail<kInit>(tp, tgrid, bf16, Cp, tail_vec, st);
    if (e == cudaSuccess) ++*launches;
    return e;
  }

  // enqueue reverse step t: GEMM layer 2, GEMM layer 3, tail (+ head of step t-1)
  cudaError_t step(int t, int64_t* launches) {
    for (int k = 0; k < K; ++k) {
      g2.scale[k] = members[k]->A[1] + (size_t)t * Fp;
      g2.shift[k] = members[k]->Cc[1] + (size_t)t * Fp;
      g3.scale[k] = members[k]->A[2] + (size_t)t * Fp;
      g3.shift[k] = members[k]->Cc[2] + (size_t)t * Fp;
    }
    cudaError_t e;
    {
      ProfSpan ps(h, st, 0);
      e = split ? launch_split_gemm(2, g2, grid, Cp, st) : launch_gemm<2>(g2, grid, bf16, Cp, ctas, st);
    }
    if (e != cudaSuccess) return e;
    tp.coef = h_coef[t];
    tp.t = t;
    tp.slot = slot_base + (a.t_first - t);
    tp.traj_entry = slot_base + (a.t_first - t);
    const bool last = (t == a.t_last);
    tp.write_out = last ? 1 : 0;
    tp.y_prev = ybuf[ycur];
    tp.y_next = ybuf[ycur ^ 1];
    ycur ^= 1;
    if (!last) set_head_rows(t - 1);
    if (fuse) {
      g3.th = tp;
      g3.do_head = last ? 0 : 1;
      g3.arrivals_target = (Fp / BN) * (++g3_launches);
      {
        ProfSpan ps(h, st, 1);
        e = launch_gemm<3>(g3, grid, bf16, Cp, ctas, st);
      }
      if (e == cudaSuccess) *launches += 2;
      return e;
    }
    {
      ProfSpan ps(h, st, 1);
      e = split ? launch_split_gemm(3, g3, grid, Cp, st) : launch_gemm<3>(g3, grid, bf16, Cp, ctas, st);
    }
    if (e != cudaSuccess) return e;
    {
      ProfSpan ps(h, st, 2);
      e = last ? launch_tail<kFinal>(tp, tgrid, bf16, Cp, tail_vec, st)
               : launch_tail<kMid>(tp, tgrid, bf16, Cp, tail_vec, st);
    }
    if (e == cudaSuccess) *launches += 3;
    return e;
  }
};

TensorChain* tensor_chain_create(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args& a,
                                 const ChainIds& ids, const StepCoef* h_coef, const TensorWorkspace& ws, int n_slots,
                                 int n_traj, cudaStream_t st, bool single_lane, int64_t* launches, std::string* err,
                                 cudaError_t* status) {
  TensorChain* c = new TensorChain();
  *status = c->init(h, members, a, ids, h_coef, ws, n_slots, n_traj, st, single_lane, launches, err);
  if (*status != cudaSuccess) {
    delete c;
    return nullptr;
  }
  return c;
}
cudaError_t tensor_chain_step(TensorChain* c, int t, int64_t* launches) { return c->step(t, launches); }
void tensor_chain_destroy(TensorChain* c) { delete c; }

int64_t debug_plan(int K, int rows, int Fp, int geometry, int row_major, int units, int32_t* table_out, int64_t cap,
                   int32_t info_out[4]) {
  const int cpu = geometry == 2 ? 2 : 1;
  const int bnt = geometry == kGeomSlim ? 128 : BN;
  const TilePlan plan = plan_tiles(K, rows, Fp / bnt, cpu, units, row_major != 0);
  info_out[0] = plan.units;
  info_out[1] = plan.stride;
  info_out[2] = plan.rows_pad;
  info_out[3] = plan.n_full + plan.has_half;
  const int64_t n = (int64_t)plan.table.size();
  if (n <= cap && table_out) std::copy(plan.table.begin(), plan.table.end(), table_out);
  return n;
}

int debug_geometry(int K, int rows, int Fp, int sm_count) {
  ladine_handle tmp;   // default options: auto geometry, default pair gain
  tmp.sm_count = sm_count;
  return choose_ctas(&tmp, K, rows, Fp);
}

cudaError_t launch_debug_layer(ladine_handle* h, const ladine_member* m, int layer, int t, const void* h_in, int rows,
                               void* h_out, float* part, int32_t* sched_buf, cudaStream_t st, std::string* err) {
  const bool bf16 = m->precision == LADINE_PREC_BF16;
  cudaError_t e = resolve_encode(h, err);
  if (e != cudaSuccess) return e;
  const int Fp = m->Fp;
  // debug entry: "ctas" = 2 exercises the CTA-pair geometry (incl. half tiles), 3 the slim 128-wide tiles
  const bool slim = h->ctas == kGeomSlim || m->split;   // FP32X always runs 128 x 128 tiles
  const int ctas = (h->ctas == 2 && !m->split) ? 2 : 1;
  const int bnt = slim ? 128 : BN;
  const TilePlan plan = plan_tiles(1, rows, Fp / bnt, ctas, h->sm_count / ctas);
  e = cudaMemcpyAsync(sched_buf, plan.table.data(), plan.table.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  GemmParams g{};
  const uint64_t ld = (uint64_t)Fp * (m->split ? 2 : 1);   // FP32X: h_in / h_out rows are [hi | lo]
  if (!make_tmap(h, &g.tmA, h_in, (uint64_t)plan.rows_pad, ld, BM, bf16, err)) return cudaErrorInvalidValue;
  if (!make_tmap(h, &g.tmB[0], layer == 2 ? m->W2h : m->W3h, Fp, ld, bnt / ctas, bf16, err)) return cudaErrorInvalidValue;
  g.scale[0] = m->A[layer - 1] + (size_t)t * Fp;
  g.shift[0] = m->Cc[layer - 1] + (size_t)t * Fp;
  g.W4[0] = m->W4;
  g.h_out = h_out;
  g.part = part;
  g.Fp = Fp;
  g.NB = Fp / BN;
  g.KB = Fp / BK;
  g.rows = rows;
  g.rows_pad = plan.rows_pad;
  g.sched = sched_buf;
  g.sched_stride = plan.stride;
  g.idesc = make_idesc(bf16, ctas, bnt);
  g.idesc_half = make_idesc(bf16, 1);
  const int grid = plan.units * ctas;
  const int geom = slim ? kGeomSlim : ctas;
  if (m->split) return launch_split_gemm(layer, g, grid, m->Cp, st);
  return layer == 2 ? launch_gemm<2>(g, grid, bf16, m->Cp, geom, st)
                    : launch_gemm<3>(g, grid, bf16, m->Cp, geom, st);
}

}  // namespace ladine
