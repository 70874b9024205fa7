// Whole-chain persistent kernel for SMALL calls at the shipped width (BASELINE.json north_star: "a persistent kernel ...
// runs all T reverse steps ... in a single launch"): up to 4 members x up to 128 chains each -- the call shape of the
// reference's own loop (one member, 70 images, one draw per p_sample_loop call: classification_train_separately.py:
// 770-777) and of config 1 -- where the three-launches-per-step path is bound by what ONE SM can pull through TMA:
// a 128 x 256 tile of a [rows, 4096] x [4096, 4096] layer needs 3 MB of operands, ~22 us at ~100 GB/s per SM, while
// 132 of the 148 SMs idle (profiles/README.md "small calls").
//
// Here every square layer is SPLIT-K over the whole chip: CTA (member k, N tile j of 256 columns, K slice s) multiplies
// the [128, F/S] slice of the activations with the [256, F/S] slice of W (tcgen05, FP32 accumulator in TMEM, operands by
// TMA) and stores an FP32 partial tile; after a grid-wide barrier every CTA reduces its 256/S-column share of the layer
// output over the S slices in fixed order (deterministic), applies scale/shift/softplus and writes the 16-bit operand of
// the next layer (layer 2) or its share of the lin4 dot products (layer 3).  The tail of the step (eps, posterior update
// in the reference's operation order, Philox or injected noise) is recomputed identically by every CTA from the lin4
// partials, so the chain state y_t lives in shared memory for all T steps; each CTA then produces its column share of
// the next step's lin1 output.  One cooperative launch per call; 6 grid barriers per reverse step:
//     G2 (split-K GEMM) | R2 (reduce + softplus) | G3 | R3 (reduce + softplus + lin4 share) | T (tail, row owners) | H (head) |
// Weights do not fit in shared memory at F = 4096 (64 MB for two layers against 33 MB of SMEM on the whole chip): they
// stream from L2, where both layers of up to 4 members stay resident.
//
// Arithmetic: same 16-bit operand rounding as the tile kernels, but the K sum is split into S partial sums, so results
// agree with the three-launch path to FP32 rounding noise (~1e-6), not bitwise.  The path is therefore OPT-IN
// (ladine_set_option("persist", 1)); the drop-in p_sample_loop enables it, the sharded ensemble API does not (its
// results are bitwise independent of the partition).
#include "ladine_tc.cuh"
#include "ladine_tensor.cuh"

namespace ladine {
namespace {

constexpr int kPersistMaxK = 4;
constexpr int kPStages = 4;
// 16 classes need 36 KB of per-CTA tables: one ring stage less
__host__ __device__ constexpr int persist_stages(int cp) { return cp > 8 ? 3 : kPStages; }
constexpr int kPABytes = BM * BK * 2;            // 16 KiB
constexpr int kPBBytes = BN * BK * 2;            // 32 KiB
constexpr int kPStageBytes = kPABytes + kPBBytes;
constexpr int kPThreads = 320;                   // warps 0..7 workers (0..3 also drain TMEM), 8 TMA producer, 9 MMA issuer
constexpr int kPWorkers = 256;
constexpr int kPProducerWarp = 8;
constexpr int kPMmaWarp = 9;
constexpr int kPTmemCols = 256;
constexpr int kPMaxShare = 128;                  // columns per CTA in the reduce phases: 256 / S, S >= 2

struct PersistParams {
  CUtensorMap tmH1, tmH2;                 // activations [K * 128, Fp] 16-bit, box 64 x 128
  CUtensorMap tmW2[kPersistMaxK];         // member weights [Fp, Fp] 16-bit, box 64 x 256
  CUtensorMap tmW3[kPersistMaxK];
  const float* A[kPersistMaxK][3];        // per-step scale rows [T, Fp] (x log2 e)
  const float* Cc[kPersistMaxK][3];
  const float* W1y[kPersistMaxK];         // [Fp, Cp]
  const float* W4[kPersistMaxK];          // [Cp, Fp]
  const float* b4[kPersistMaxK];
  void* h1;                               // [K * 128, Fp] 16-bit
  void* h2;
  float* part;                            // [S, K * 128, Fp] FP32 split-K partial tiles (layer 2, then layer 3)
  float* epart;                           // [K * 128, Q, Cp] lin4 partial dot products, Q = NT * S column shares
  float* ybuf;                            // [K * 128, Cp] chain state published by the row owners once per step
  unsigned int* barrier;                  // grid barrier counter (zeroed before the launch)
  const StepCoef* coef;                   // device [T]
  const float* xf;                        // [K, N, Fin]
  const float* u;                         // [K, N, Fp]
  const float* ytmean;                    // [K, N, C]
  const float* y_init;                    // [K, D, N, C] or null
  const float* noise;                     // [K, D, n_slots, N, C] or null
  float* y_out;                           // [K, D, N, C]
  float* traj_out;                        // [K, D, n_traj, N, C] or null
  float* prob_out;                        // [K, D, N, C] or null
  float temperature;
  uint64_t seed;
  ChainIds ids;
  uint32_t idesc;
  int K, N, D, C, Fin, Fp, NT, S, KBS, Q, cs, rows, t_first, t_last, n_slots, n_traj;
  int debug;                              // 1: block 0 prints its per-phase clock totals at the end (ladine_set_option "persist_debug")
};

struct __align__(8) PersistBarriers {
  uint64_t full[kPStages];
  uint64_t empty[kPStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// Grid-wide barrier: every CTA of the (cooperative, co-resident) grid arrives once per call.  Release: thread 0's
// red.release.gpu after the block barrier; acquire: thread 0's ld.acquire.gpu spin, then the block barrier (the pattern of
// CUTLASS' Semaphore).  Bounded spin: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int& epoch) {
  ++epoch;
  __syncthreads();   // every thread's writes of the phase happen-before thread 0's release (cumulativity)
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const unsigned int target = epoch * gridDim.x;
    const long long t0 = clock64();
    unsigned int seen;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      if (seen >= target) break;
      if (clock64() - t0 > 8000000000LL) {
        printf("ladine: grid barrier timeout block=%d epoch=%u seen=%u target=%u\n", (int)blockIdx.x, epoch, seen, target);
        __trap();
      }
    }
  }
  __syncthreads();   // the other threads' reads are ordered after thread 0's acquire; data written by other CTAs is read
                     // with ld.global.cg / TMA (L2), never through this SM's L1
}

template <typename T16, int CP>
__global__ void __launch_bounds__(kPThreads, 1) persistent_chain_kernel(const __grid_constant__ PersistParams p) {
  constexpr int NS = persist_stages(CP);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // NS x 16 KiB
  uint8_t* sB = smem + NS * kPABytes;                   // NS x 32 KiB
  float* sY = reinterpret_cast<float*>(smem + NS * kPStageBytes);   // [128][CP] chain state of this CTA's member
  float* sMu = sY + BM * CP;                            // [128][CP] prior mean of each row (step-invariant)
  float* sW1y = sMu + BM * CP;                          // [128][CP] lin1 weights of this CTA's column share
  float* sW4 = sW1y + kPMaxShare * CP;                  // [CP][128] lin4 weights of the share
  float* sTab = sW4 + kPMaxShare * CP;                  // [6][128]: A2_t C2_t A3_t C3_t A1_{t-1} C1_{t-1} of the share
  StepCoef* sCoef = reinterpret_cast<StepCoef*>(sTab + 6 * kPMaxShare);
  PersistBarriers* bars = reinterpret_cast<PersistBarriers*>(sCoef + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  // CTA -> (member k, N tile j, K slice s); q = column share of the member's layer output owned in the reduce phases
  const int per_member = p.NT * p.S;
  const int k = blockIdx.x / per_member;
  const int q = blockIdx.x - k * per_member;
  const int j = q / p.S, s = q - j * p.S;
  const int col0 = q * p.cs;                            // first of this CTA's cs = 256 / S output columns
  const int R = p.rows, C = p.C;
  const size_t mrow0 = (size_t)k * BM;                  // first row of the member in h1 / h2 / part

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&p.tmH1);
    tma_prefetch_desc(&p.tmH2);
    tma_prefetch_desc(&p.tmW2[k]);
    tma_prefetch_desc(&p.tmW3[k]);
  }
  if (warp == 0) tmem_alloc(smem_u32(&bars->tmem_base), kPTmemCols);
  for (int i = tid; i < BM * CP; i += kPThreads) {
    sY[i] = 0.f;
    const int row = i / CP, c = i - row * CP;
    sMu[i] = (row < R && c < C) ? __ldg(p.ytmean + ((size_t)k * p.N + row / p.D) * C + c) : 0.f;
  }
  for (int i = tid; i < p.cs * CP; i += kPThreads) {   // step-invariant weights of this CTA's column share
    sW1y[i] = __ldg(p.W1y[k] + (size_t)col0 * CP + i);               // [col][c], contiguous in the member's [Fp, Cp] table
    const int c = i / p.cs, cl = i - c * p.cs;
    sW4[c * kPMaxShare + cl] = __ldg(p.W4[k] + (size_t)c * p.Fp + col0 + cl);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // per-phase clock totals of block 0 (debug option): G2, R2, G3, R3, H, the five barriers that follow them, and three
  // sub-phases of H (partial sums / posterior update / head)
  long long acc_clk[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = clock64(), hmark = 0;
  auto mark = [&](int slot) {
    if (p.debug && blockIdx.x == 0) {
      __syncthreads();
      const long long now = clock64();
      acc_clk[slot] += now - tmark;
      tmark = now;
    }
  };
  auto hsub = [&](int slot) {   // workers only (named barrier 2)
    if (p.debug && blockIdx.x == 0 && tid < kPWorkers) {
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const long long now = clock64();
      if (slot >= 0) acc_clk[slot] += now - hmark;
      hmark = now;
    }
  };

  unsigned int epoch = 0;
  // smem ring cursors: `bcur` = where the next W tile goes (it also arms the stage's barrier), `acur` = where the next
  // activation tile goes.  They differ because the W tiles of the next GEMM phase are issued BEFORE the grid barrier
  // (weights are immutable), the activation tiles after it.  The MMA thread walks the ring with `mcur`.
  int bcur = 0, acur = 0, mcur = 0;
  uint32_t bphase = 0, mphase = 0;
  int pre = 0;              // stages of the coming GEMM phase whose W tile is already in flight
  uint32_t acc_phase = 0;   // parity of acc_full, toggles once per GEMM phase

  auto issue_b = [&](int layer, int kb) {
    mbar_wait(smem_u32(&bars->empty[bcur]), bphase ^ 1u, 0);
    const uint32_t fb = smem_u32(&bars->full[bcur]);
    mbar_arrive_expect_tx(fb, kPStageBytes);   // A + B bytes: the stage completes when both have landed
    tma_load_2d(smem_u32(sB + bcur * kPBBytes), layer == 2 ? &p.tmW2[k] : &p.tmW3[k], fb, (s * p.KBS + kb) * BK, j * BN);
    if (++bcur == NS) { bcur = 0; bphase ^= 1u; }
  };

  // ---------------------------------------------------------------------------------------------------------------
  // phase G: split-K GEMM of layer L (2 or 3) -> part[s][member rows][256 columns of tile j].  `next_layer` (0: none):
  // the GEMM phase that follows, whose first W tiles are prefetched.  `t_tab` >= 0: warps 4..7 (idle here) stage the
  // per-step table rows of the column share for the phases that follow (layer 2 only).
  // ---------------------------------------------------------------------------------------------------------------
  auto gemm_phase = [&](int layer, int next_layer, int t_tab, int t_head) {
    if (warp == kPProducerWarp) {
      if (lane == 0) {
        // h1 / h2 were written with generic stores by other CTAs before the grid barrier: order them before the TMA
        // (async proxy) reads of this thread
        asm volatile("fence.proxy.async.global;" ::: "memory");
        const CUtensorMap* ta = layer == 2 ? &p.tmH1 : &p.tmH2;
        for (int kb = 0; kb < p.KBS; ++kb) {
          if (kb >= pre) issue_b(layer, kb);
          tma_load_2d(smem_u32(sA + acur * kPABytes), ta, smem_u32(&bars->full[acur]), (s * p.KBS + kb) * BK, (int)mrow0);
          if (++acur == NS) acur = 0;
        }
        pre = 0;
        if (next_layer) {
          pre = p.KBS < NS ? p.KBS : NS;
          for (int kb = 0; kb < pre; ++kb) issue_b(next_layer, kb);
        }
      }
    } else if (warp == kPMmaWarp) {
      if (lane == 0) {
        tc_fence_after();
        for (int kb = 0; kb < p.KBS; ++kb) {
          mbar_wait(smem_u32(&bars->full[mcur]), mphase, 2);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + mcur * kPABytes));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + mcur * kPBBytes));
#pragma unroll
          for (int k4 = 0; k4 < BK / UK; ++k4)
            umma_f16(tmem_base, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), p.idesc, (uint32_t)((kb | k4) != 0));
          umma_commit(smem_u32(&bars->empty[mcur]));
          if (++mcur == NS) { mcur = 0; mphase ^= 1u; }
        }
        umma_commit(smem_u32(&bars->acc_full));
      }
    } else if (warp < 4) {
      // drain the accumulator: thread = TMEM lane = row; FP32 partial tile, full 32-byte sectors
      mbar_wait(smem_u32(&bars->acc_full), acc_phase, 3);
      tc_fence_after();
      const int r = warp * 32 + lane;
      float* dst = p.part + (((size_t)s * p.K * BM) + mrow0 + r) * p.Fp + j * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(ch * 32), v);
        tmem_ld_wait();
        if (r < R) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = v[8 * g8 + i];
            st_global_256(dst + ch * 32 + 8 * g8, o);
          }
        }
      }
      tc_fence_before();
    } else if (t_tab >= 0) {
      // warps 4..7: table rows of this step for the share (consumed after the next grid barriers, never through L1 of
      // another step: they are written here once per step)
      const int wt = tid - 128;   // 0..127
      for (int i = wt; i < 6 * p.cs; i += 128) {
        const int which = i / p.cs, cl = i - which * p.cs;
        float v = 0.f;
        if (which < 4) {
          const float* base = (which & 1) ? p.Cc[k][1 + (which >> 1)] : p.A[k][1 + (which >> 1)];
          v = __ldg(base + (size_t)t_tab * p.Fp + col0 + cl);
        } else if (t_head >= 0) {
          const float* base = (which & 1) ? p.Cc[k][0] : p.A[k][0];
          v = __ldg(base + (size_t)t_head * p.Fp + col0 + cl);
        }
        sTab[which * kPMaxShare + cl] = v;
      }
      if (wt == 0) *sCoef = p.coef[t_tab];
    }
    acc_phase ^= 1u;
  };

  const int upr = p.cs >> 2;                 // 4-column units per row in this CTA's share (8 for S = 8 ... 32 for S = 2)
  const int units = R * upr;

  // ---------------------------------------------------------------------------------------------------------------
  // phase R2: h2[:, share] = softplus(A2_t * sum_s part + C2_t)  -> 16-bit operand of layer 3
  // ---------------------------------------------------------------------------------------------------------------
  // loads of the S partials of one unit (issued together; summed later in ascending slice order)
  struct Unit4 { float4 v[8]; };
  auto load_unit = [&](int row, int col, Unit4& u4) {
    const float* src = p.part + (mrow0 + row) * p.Fp + col;
    const size_t sstride = (size_t)p.K * BM * p.Fp;
#pragma unroll
    for (int ss = 0; ss < 8; ++ss)
      u4.v[ss] = ss < p.S ? __ldcg(reinterpret_cast<const float4*>(src + ss * sstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sum_unit = [&](const Unit4& u4) {
    float4 z = u4.v[0];
#pragma unroll
    for (int ss = 1; ss < 8; ++ss) {
      if (ss < p.S) { z.x += u4.v[ss].x; z.y += u4.v[ss].y; z.z += u4.v[ss].z; z.w += u4.v[ss].w; }
    }
    return z;
  };

  auto reduce2_phase = [&]() {
    if (tid >= kPWorkers) return;
    // two units per iteration: 2 x S independent 128-bit loads in flight before the first add (L2-latency-bound phase)
    for (int u0 = tid; u0 < units; u0 += 2 * kPWorkers) {
      Unit4 ld[2];
      int rowv[2], clv[2];
      bool livev[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int uidx = u0 + h * kPWorkers;
        livev[h] = uidx < units;
        rowv[h] = livev[h] ? uidx / upr : 0;
        clv[h] = 4 * (livev[h] ? uidx - rowv[h] * upr : 0);
        load_unit(rowv[h], col0 + clv[h], ld[h]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!livev[h]) continue;
        const float4 z = sum_unit(ld[h]);
        const float4 sc = *reinterpret_cast<const float4*>(sTab + 0 * kPMaxShare + clv[h]);
        const float4 sh = *reinterpret_cast<const float4*>(sTab + 1 * kPMaxShare + clv[h]);
        uint2 o;
        o.x = Pack16<T16>::pack(softplus_log2dom(fmaf(sc.x, z.x, sh.x)), softplus_log2dom(fmaf(sc.y, z.y, sh.y)));
        o.y = Pack16<T16>::pack(softplus_log2dom(fmaf(sc.z, z.z, sh.z)), softplus_log2dom(fmaf(sc.w, z.w, sh.w)));
        *reinterpret_cast<uint2*>(reinterpret_cast<T16*>(p.h2) + (mrow0 + rowv[h]) * p.Fp + col0 + clv[h]) = o;
      }
    }
  };

  // ---------------------------------------------------------------------------------------------------------------
  // phase R3: h3[:, share] = softplus(A3_t * sum_s part + C3_t) (kept in FP32); epart[row][q] = h3[row, share] . W4^T
  // ---------------------------------------------------------------------------------------------------------------
  auto reduce3_phase = [&]() {
    if (tid >= kPWorkers) return;
    const int stride2 = 2 * kPWorkers;
    const int padded = (units + stride2 - 1) / stride2 * stride2;   // all lanes of a row group stay in the shuffles
    for (int u0 = tid; u0 < padded; u0 += stride2) {
      Unit4 ld[2];
      int rowv[2], clv[2];
      bool livev[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int uidx = u0 + h * kPWorkers;
        livev[h] = uidx < units;
        rowv[h] = livev[h] ? uidx / upr : 0;
        clv[h] = 4 * (livev[h] ? uidx - rowv[h] * upr : 0);
        if (livev[h]) load_unit(rowv[h], col0 + clv[h], ld[h]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float e[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) e[c] = 0.f;
        if (livev[h]) {
          const float4 z = sum_unit(ld[h]);
          const float4 sc = *reinterpret_cast<const float4*>(sTab + 2 * kPMaxShare + clv[h]);
          const float4 sh = *reinterpret_cast<const float4*>(sTab + 3 * kPMaxShare + clv[h]);
          const float h0 = softplus_log2dom(fmaf(sc.x, z.x, sh.x)), h1v = softplus_log2dom(fmaf(sc.y, z.y, sh.y));
          const float h2v = softplus_log2dom(fmaf(sc.z, z.z, sh.z)), h3v = softplus_log2dom(fmaf(sc.w, z.w, sh.w));
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(sW4 + c * kPMaxShare + clv[h]);
            e[c] = fmaf(h3v, w.w, fmaf(h2v, w.z, fmaf(h1v, w.y, h0 * w.x)));
          }
        }
        // the upr threads of one row are consecutive lanes (upr is a power of two <= 32): fixed-shape tree
        for (int o = upr >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int c = 0; c < CP; ++c) e[c] += __shfl_xor_sync(0xffffffffu, e[c], o);
        }
        const int uidx = u0 + h * kPWorkers;
        if (livev[h] && (uidx % upr) == 0) {
          float* dst = p.epart + ((mrow0 + rowv[h]) * p.Q + q) * CP;
#pragma unroll
          for (int c = 0; c < CP; ++c) dst[c] = e[c];
        }
      }
    }
  };

  // ---------------------------------------------------------------------------------------------------------------
  // phase T (tail): finish step t (t < 0: initialise y_T) for the rows this CTA OWNS (row = q, q + Q, ...): sum the row's
  // Q lin4 partials, apply the posterior update in the reference's operation order, publish y_{t-1} to `ybuf` (and the
  // trajectory / outputs).  Letting every CTA recompute every row instead was measured 3-6x slower: 128 CTAs reading the
  // same partials at the same time serialise on the same L2 lines (7.8k cycles for 64 rows, 18k for 128).
  // A row is handled by `tpr` consecutive lanes: they split the Q partials, and lane c of the group updates class c.
  // ---------------------------------------------------------------------------------------------------------------
  int own_rows = (R - q + p.Q - 1) / p.Q;          // rows q, q + Q, ... < R
  if (own_rows < 0) own_rows = 0;
  int own_cap = (R + p.Q - 1) / p.Q;               // the same for every CTA of the member: fixes the lane grouping
  int rp2 = 8;
  while (rp2 < own_cap) rp2 <<= 1;
  const int tpr = kPWorkers / rp2 < 32 ? kPWorkers / rp2 : 32;   // 4 .. 32 lanes per row (>= C: checked on the host)

  auto tail_phase = [&](int t, int slot, int traj_entry, bool write_out, bool tab_ready) {
    hsub(-1);
    if (tid >= kPWorkers) return;
    const int grp = tid / tpr, sub = tid - grp * tpr;
    const bool live = grp < own_rows;
    const int row = live ? q + grp * p.Q : 0;
    float eps[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) eps[c] = 0.f;
    if (t >= 0 && live) {
      // this lane's run of the row's Q partials is contiguous: [q][CP] floats.  Batches of 8 independent 128-bit loads
      // (one L2 round trip per batch), added in ascending q order.
      const int qn = (p.Q + tpr - 1) / tpr;
      const int q0 = sub * qn, q1 = min(p.Q, q0 + qn);
      const float* src = p.epart + ((mrow0 + row) * p.Q + q0) * CP;
      const bool vec_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
      const float4* src4 = reinterpret_cast<const float4*>(src);
      const int n4 = (q1 > q0 && vec_ok) ? (q1 - q0) * CP / 4 : 0;   // CP >= 4: CP / 4 vectors per q; CP == 2: one per 2 q
      for (int i0 = 0; i0 < n4; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (i0 + i) < n4 ? __ldcg(src4 + i0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e4[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            // element 4 * (i0 + i) + qq of the run belongs to class (that index) % CP; i0 is a multiple of 8
            const int c = (4 * i + qq) % CP;
#pragma unroll
            for (int cc = 0; cc < CP; ++cc)
              if (cc == c) eps[cc] += e4[qq];
          }
        }
      }
      // what the vectors did not cover: an odd tail with 2 classes, or the whole (short, unaligned) run
      for (int qq = q0 + n4 * 4 / CP; qq < q1; ++qq) {
#pragma unroll
        for (int c = 0; c < CP; ++c) eps[c] += __ldcg(p.epart + ((mrow0 + row) * p.Q + qq) * CP + c);
      }
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) {   // butterfly: every lane of the group ends with the row's sums
#pragma unroll
      for (int c = 0; c < CP; ++c) eps[c] += __shfl_xor_sync(0xffffffffu, eps[c], o);
    }
    hsub(10);
    // lane c of the group finishes class c
    const int n = row / p.D, d = row - n * p.D;
    float v = 0.f;
    if (live && sub < C) {
      const int c = sub;
      float e = 0.f;
#pragma unroll
      for (int cc = 0; cc < CP; ++cc)
        if (cc == c) e = eps[cc];
      const float mu = sMu[row * CP + c];
      if (t < 0) {
        if (p.y_init) {
          v = __ldg(p.y_init + (((size_t)k * p.D + d) * p.N + n) * C + c);
        } else {
          const float z = p.noise ? __ldg(p.noise + ((((size_t)k * p.D + d) * p.n_slots + 0) * p.N + n) * C + c)
                                  : philox_normal(p.seed, p.ids.chain(k, d, n), 0u, c);
          v = __fadd_rn(z, mu);
        }
      } else {
        const StepCoef cf = tab_ready ? *sCoef : p.coef[t];
        e += __ldg(p.b4[k] + c);
        const float y = sY[row * CP + c];
        if (t > 0) {
          const float z = p.noise ? __ldg(p.noise + ((((size_t)k * p.D + d) * p.n_slots + slot) * p.N + n) * C + c)
                                  : philox_normal(p.seed, p.ids.chain(k, d, n), (uint32_t)slot, c);
          v = posterior_step_op(cf, y, mu, e, z);
        } else {
          v = y0_reparam_op(cf, y, mu, e);
        }
      }
      p.ybuf[(mrow0 + row) * CP + c] = v;
      if (p.traj_out && traj_entry >= 0)
        p.traj_out[((((size_t)k * p.D + d) * p.n_traj + traj_entry) * p.N + n) * C + c] = v;
      if (write_out) p.y_out[(((size_t)k * p.D + d) * p.N + n) * C + c] = v;
    }
    if (write_out && p.prob_out) {
      // softmax(-(y-1)^2 / temperature) over the classes of the row -- classification_train_separately.py:392-398;
      // the classes sit in lanes 0..C-1 of the group
      const float lg = (live && sub < C) ? -(v - 1.0f) * (v - 1.0f) / p.temperature : -INFINITY;
      float mx = lg;
      for (int o = tpr >> 1; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float ex = (live && sub < C) ? expf(lg - mx) : 0.f;
      float sum = ex;
      for (int o = tpr >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (live && sub < C) p.prob_out[(((size_t)k * p.D + d) * p.N + n) * C + sub] = ex / sum;
    }
    hsub(11);
  };

  // ---------------------------------------------------------------------------------------------------------------
  // phase H (head): fetch y of every row of the member (published by the owners before the grid barrier) and produce
  // lin1 of step t_next for this CTA's column share.
  // ---------------------------------------------------------------------------------------------------------------
  auto head_phase = [&](int t_next, bool tab_ready) {
    hsub(-1);
    if (tid >= kPWorkers) return;
    for (int i = tid; i < R * CP; i += kPWorkers) sY[i] = __ldcg(p.ybuf + mrow0 * CP + i);
    asm volatile("bar.sync 1, 256;" ::: "memory");   // y of every row is in sY (the 256 worker threads)
    if (t_next < 0) return;
    const float* a1g = p.A[k][0] + (size_t)t_next * p.Fp;
    const float* c1g = p.Cc[k][0] + (size_t)t_next * p.Fp;
    // two units per iteration: both units' global loads (u, xf) are in flight before either is used
    for (int u0 = tid; u0 < units; u0 += 2 * kPWorkers) {
      float4 uu[2], xx[2];
      int rowv[2], clv[2];
      bool livev[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int uidx = u0 + h * kPWorkers;
        livev[h] = uidx < units;
        rowv[h] = livev[h] ? uidx / upr : 0;
        clv[h] = 4 * (livev[h] ? uidx - rowv[h] * upr : 0);
        const int n = rowv[h] / p.D, col = col0 + clv[h];
        uu[h] = __ldg(reinterpret_cast<const float4*>(p.u + ((size_t)k * p.N + n) * p.Fp + col));
        const float* xr = p.xf + ((size_t)k * p.N + n) * p.Fin + col;
        if (col + 3 < p.Fin && (p.Fin & 3) == 0) {
          xx[h] = __ldg(reinterpret_cast<const float4*>(xr));
        } else {
          xx[h].x = col + 0 < p.Fin ? __ldg(xr + 0) : 0.f;
          xx[h].y = col + 1 < p.Fin ? __ldg(xr + 1) : 0.f;
          xx[h].z = col + 2 < p.Fin ? __ldg(xr + 2) : 0.f;
          xx[h].w = col + 3 < p.Fin ? __ldg(xr + 3) : 0.f;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!livev[h]) continue;
        const int row = rowv[h], cl = clv[h], col = col0 + cl;
        float4 a, cc;
        if (tab_ready) {
          a = *reinterpret_cast<const float4*>(sTab + 4 * kPMaxShare + cl);
          cc = *reinterpret_cast<const float4*>(sTab + 5 * kPMaxShare + cl);
        } else {
          a = __ldg(reinterpret_cast<const float4*>(a1g + col));
          cc = __ldg(reinterpret_cast<const float4*>(c1g + col));
        }
        const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {cc.x, cc.y, cc.z, cc.w};
        const float uv[4] = {uu[h].x, uu[h].y, uu[h].z, uu[h].w}, xv[4] = {xx[h].x, xx[h].y, xx[h].z, xx[h].w};
        float hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // same arithmetic as tailhead_kernel: v2 = sum_c (A1 * W1y[c]) * y[c] + (A1 * u + C1) in the log2 domain
          float v2 = fmaf(av[i], uv[i], cv[i]);
#pragma unroll
          for (int c = 0; c < CP; ++c) v2 = fmaf(av[i] * sW1y[(cl + i) * CP + c], sY[row * CP + c], v2);
          const float tt = v2 > kSoftplusThreshold * kLog2e ? v2 : lg2_approx(1.0f + ex2_approx(v2));
          hv[i] = tt * (xv[i] * kLn2);
        }
        uint2 o;
        o.x = Pack16<T16>::pack(hv[0], hv[1]);
        o.y = Pack16<T16>::pack(hv[2], hv[3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<T16*>(p.h1) + (mrow0 + row) * p.Fp + col) = o;
      }
    }
    hsub(12);
  };

  // ---------------------------------------------------------------------------------------------------------------
  // the chain
  // ---------------------------------------------------------------------------------------------------------------
  const int slot_base = p.y_init ? 0 : 1;
  if (warp == kPProducerWarp && lane == 0) {   // W tiles of the first GEMM phase: in flight while y_T / h1 are produced
    pre = p.KBS < NS ? p.KBS : NS;
    for (int kb = 0; kb < pre; ++kb) issue_b(2, kb);
  }
  tail_phase(-1, 0, p.y_init ? -1 : 0, false, false);
  grid_barrier(p.barrier, epoch);
  head_phase(p.t_first, false);
  grid_barrier(p.barrier, epoch);
  mark(9);
  for (int t = p.t_first; t >= p.t_last; --t) {
    const bool last = t == p.t_last;
    gemm_phase(2, 3, t, last ? -1 : t - 1);
    mark(0);
    grid_barrier(p.barrier, epoch);
    mark(5);
    reduce2_phase();
    mark(1);
    grid_barrier(p.barrier, epoch);
    mark(6);
    gemm_phase(3, last ? 0 : 2, -1, -1);
    mark(2);
    grid_barrier(p.barrier, epoch);
    mark(7);
    reduce3_phase();
    mark(3);
    grid_barrier(p.barrier, epoch);
    mark(8);
    const int sl = slot_base + (p.t_first - t);
    tail_phase(t, sl, sl, last, true);
    if (last) break;
    grid_barrier(p.barrier, epoch);
    head_phase(t - 1, true);
    mark(4);
    grid_barrier(p.barrier, epoch);
    mark(9);
  }
  if (p.debug && blockIdx.x == 0 && tid == 0) {
    const double n = (double)(p.t_first - p.t_last + 1);
    printf("ladine persist (block 0, clocks per reverse step): G2 %.0f R2 %.0f G3 %.0f R3 %.0f H %.0f | barriers after: "
           "G2 %.0f R2 %.0f G3 %.0f R3 %.0f T|H %.0f | tail sums %.0f + update %.0f, head %.0f\n", acc_clk[0] / n, acc_clk[1] / n,
           acc_clk[2] / n, acc_clk[3] / n, acc_clk[4] / n, acc_clk[5] / n, acc_clk[6] / n, acc_clk[7] / n, acc_clk[8] / n,
           acc_clk[9] / n, acc_clk[10] / n, acc_clk[11] / n, acc_clk[12] / n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kPTmemCols);
  }
}

size_t persist_smem(int Cp) {
  return 1024 + (size_t)persist_stages(Cp) * kPStageBytes + sizeof(float) * (2 * BM * Cp + 2 * kPMaxShare * Cp + 6 * kPMaxShare) +
         sizeof(StepCoef) + sizeof(PersistBarriers);
}

template <typename T16, int CP>
cudaError_t launch_persist_t(const PersistParams& p, int grid, cudaStream_t st) {
  auto kern = persistent_chain_kernel<T16, CP>;
  const size_t smem = persist_smem(CP);
  static std::atomic<uint64_t> configured{0};
  cudaError_t e = configure_once(configured, [&] {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPThreads, smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (per_sm < 1 || grid > per_sm * sms) return cudaErrorCooperativeLaunchTooLarge;   // every CTA must be resident
  void* args[] = {const_cast<PersistParams*>(&p)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(kPThreads), args, smem, st);
}

}  // namespace

// K slices per layer for a call of K members x `rows` chains each, or 0 when the persistent kernel does not apply
int persist_splits(const ladine_handle* h, int K, int rows, int Fp, int C) {
  if (K < 1 || K > kPersistMaxK || rows < 1 || rows > BM || Fp % BN != 0) return 0;
  const int NT = Fp / BN, KB = Fp / BK;
  int S = 8;
  while (S >= 2 && (K * NT * S > h->sm_count || KB % S != 0)) S >>= 1;
  // measured (F = 4096, 70 chains per member): S = 8 (one member) 31 vs 56 us per reverse step on the tile kernels,
  // S = 4 (two members) 50 vs 62, S = 2 (four members) 82 vs 83 -- below 4 slices the split does not pay
  if (S < 4) return 0;
  // the tail phase gives each owned row a group of lanes, one lane per class (see tail_phase)
  const int Q = NT * S, own_cap = (rows + Q - 1) / Q;
  int rp2 = 8;
  while (rp2 < own_cap) rp2 <<= 1;
  const int tpr = kPWorkers / rp2 < 32 ? kPWorkers / rp2 : 32;
  return tpr >= C ? S : 0;
}

size_t persist_workspace_bytes(int K, int Fp, int Cp, int S) {
  const size_t mt = (size_t)K * BM;
  const size_t Q = (size_t)(Fp / BN) * S;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  return up(mt * Fp * 2) * 2 + up((size_t)S * mt * Fp * 4) + up(mt * Q * Cp * 4) + up(mt * Cp * 4) + up(1024);
}

cudaError_t launch_persistent_chain(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args& a,
                                    const ChainIds& ids, const StepCoef* d_coef, const float* d_u, uint8_t* ws, int S,
                                    int n_slots, int n_traj, cudaStream_t st, int64_t* launches, std::string* err) {
  const ladine_member* m0 = members[0];
  const bool bf16 = m0->precision == LADINE_PREC_BF16;
  const int K = a.K, Fp = m0->Fp, Cp = m0->Cp;
  cudaError_t e = resolve_encode(h, err);
  if (e != cudaSuccess) return e;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  const size_t mt = (size_t)K * BM;
  PersistParams p{};
  uint8_t* cur = ws;
  p.h1 = cur; cur += up(mt * Fp * 2);
  p.h2 = cur; cur += up(mt * Fp * 2);
  p.part = reinterpret_cast<float*>(cur); cur += up((size_t)S * mt * Fp * 4);
  p.NT = Fp / BN;
  p.S = S;
  p.Q = p.NT * S;
  p.epart = reinterpret_cast<float*>(cur); cur += up(mt * p.Q * Cp * 4);
  p.ybuf = reinterpret_cast<float*>(cur); cur += up(mt * Cp * 4);
  p.barrier = reinterpret_cast<unsigned int*>(cur);
  // rows of a member beyond `rows` are never written: zero both operand buffers once so the padded MMA rows are finite
  e = cudaMemsetAsync(p.h1, 0, up(mt * Fp * 2) * 2, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(p.barrier, 0, 1024, st);
  if (e != cudaSuccess) return e;
  if (!make_tmap(h, &p.tmH1, p.h1, mt, Fp, BM, bf16, err)) return cudaErrorInvalidValue;
  if (!make_tmap(h, &p.tmH2, p.h2, mt, Fp, BM, bf16, err)) return cudaErrorInvalidValue;
  for (int k = 0; k < K; ++k) {
    if (!make_tmap(h, &p.tmW2[k], members[k]->W2h, Fp, Fp, BN, bf16, err)) return cudaErrorInvalidValue;
    if (!make_tmap(h, &p.tmW3[k], members[k]->W3h, Fp, Fp, BN, bf16, err)) return cudaErrorInvalidValue;
    for (int l = 0; l < 3; ++l) {
      p.A[k][l] = members[k]->A[l];
      p.Cc[k][l] = members[k]->Cc[l];
    }
    p.W1y[k] = members[k]->W1y;
    p.W4[k] = members[k]->W4;
    p.b4[k] = members[k]->b4;
  }
  p.coef = d_coef;
  p.xf = a.xf;
  p.u = d_u;
  p.ytmean = a.ytmean;
  p.y_init = a.y_init;
  p.noise = a.noise;
  p.y_out = a.y_out;
  p.traj_out = a.traj_out;
  p.prob_out = a.prob_out;
  p.temperature = a.prob_out ? a.temperature : 1.0f;
  p.seed = a.seed;
  p.ids = ids;
  const uint32_t fmt = bf16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  p.K = K;
  p.N = a.N;
  p.D = a.D;
  p.C = m0->C;
  p.Fin = m0->F;
  p.Fp = Fp;
  p.KBS = (Fp / BK) / S;
  p.cs = BN / S;
  p.rows = a.N * a.D;
  p.t_first = a.t_first;
  p.t_last = a.t_last;
  p.n_slots = n_slots;
  p.n_traj = n_traj;
  p.debug = h->persist_debug ? 1 : 0;
  const int grid = K * p.NT * S;
#define LADINE_PERSIST_CASE(CPV)                                                     \
  case CPV:                                                                          \
    e = bf16 ? launch_persist_t<__nv_bfloat16, CPV>(p, grid, st) : launch_persist_t<__half, CPV>(p, grid, st); \
    break;
  switch (Cp) {
    LADINE_PERSIST_CASE(2)
    LADINE_PERSIST_CASE(4)
    LADINE_PERSIST_CASE(8)
    LADINE_PERSIST_CASE(16)
    default: e = cudaErrorInvalidValue;
  }
#undef LADINE_PERSIST_CASE
  if (e == cudaSuccess) *launches += 1;
  return e;
}

}  // namespace ladine
