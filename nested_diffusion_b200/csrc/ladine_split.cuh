// FP32-grade GEMM mainloop on the tcgen05 tensor cores: FP16 hi + lo split operands, error-corrected accumulation.
// Shared by the FP32X reverse-step kernels (ladine_split.cu) and the encoder prologue (ladine_encoder.cu).
//
//   A = A_hi + A_lo,  W = W_hi + W_lo  (FP16 parts of operands pre-scaled by powers of two; buffers hold [hi | lo] halves
//   side by side, the lo half starting at column `lo_off`).  Per 64-wide K block ONE pipeline stage carries the four
//   128 x 64 tiles A_hi, A_lo, W_hi, W_lo (64 KiB, 4 TMA loads) and twelve tcgen05.mma (M = N = 128, K = 16) are issued:
//       corr += A_lo . W_hi        corr += A_hi . W_lo        main += A_hi . W_hi          (A_lo . W_lo ~ 2^-22: dropped)
//   into two FP32 accumulators in TMEM.  The tensor core TRUNCATES when it adds into an FP32 accumulator -- about half
//   an ulp of the running sum per instruction, always toward zero, so over a long K loop the error grows linearly and
//   keeps the sign of the running sum (measured: 3.9e-6 of the output at K = 4096, 6e-5 at K = 20000, against ~3e-7 for
//   FP32 FMA).  Two measures bring it back to FP32 grade:
//     * the correction terms (2/3 of the instructions) have their own accumulator, so their truncation is at the scale
//       of the 2^-11-times smaller correction sum (Ootomo & Yokota, error-corrected tensor-core GEMM);
//     * TMEM accumulation is limited to chunks of 8 K blocks (32 main instructions); the epilogue warps promote each
//       chunk (main + corr) into FP32 REGISTER accumulators with round-to-nearest adds while the MMAs of the next chunk
//       run into the other TMEM stage.  Chunk sums have independent signs, so what is left adds up like rounding noise.
//   Shared-memory port per K block and 128 x 128 tile: 64 KiB written by TMA + 96 KiB read by the MMAs = 1280 cycles for
//   768 cycles of tensor pipe: the mainloop is port-bound like the 16-bit kernels, at ~3.5x their cost per output.
#pragma once
#include "ladine_tc.cuh"

namespace ladine {
namespace split {

constexpr int SBM = 128;                       // rows per tile (UMMA M)
constexpr int SBN = 128;                       // columns per tile (UMMA N)
constexpr int SBK = 64;                        // K per stage (one 128-byte swizzle row of FP16)
constexpr int SUK = 16;                        // K per tcgen05.mma
constexpr int kStages = 3;
constexpr int kTile = SBM * SBK * 2;           // 16 KiB
constexpr int kStageBytes = 4 * kTile;         // A_hi | A_lo | W_hi | W_lo
constexpr int kChunk = 8;                      // K blocks accumulated in TMEM before promotion to registers
constexpr int kThreads = 192;                  // warps 0..3 epilogue (TMEM lane quadrant = warp id), 4 producer, 5 MMA
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kTmemCols = 512;                 // 2 stages x (main 128 + corr 128)

struct __align__(8) Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

// kind::f16, FP16 operands, FP32 accumulate, K-major A and B, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(SBN >> 3) << 17) | ((uint32_t)(SBM >> 4) << 24);

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
};

// by thread 0 before the block-wide barrier; warp 0 then allocates the TMEM columns
__device__ __forceinline__ void init_barriers(Barriers* bars) {
  for (int s = 0; s < kStages; ++s) {
    mbar_init(smem_u32(&bars->full[s]), 1);
    mbar_init(smem_u32(&bars->empty[s]), 1);
  }
  for (int s = 0; s < 2; ++s) {
    mbar_init(smem_u32(&bars->acc_full[s]), 1);
    mbar_init(smem_u32(&bars->acc_empty[s]), 4);   // the four epilogue warps
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// TMA producer (one thread): K blocks [k0, k1) of the tile whose A rows start at `arow` and W rows at `brow`
__device__ __forceinline__ void produce(uint8_t* smem, Barriers* bars, Pipe& ps, const CUtensorMap* tmA,
                                        const CUtensorMap* tmB, int arow, int brow, int lo_off, int k0, int k1) {
  for (int kb = k0; kb < k1; ++kb) {
    mbar_wait(smem_u32(&bars->empty[ps.stage]), ps.phase ^ 1u, 0);
    const uint32_t fb = smem_u32(&bars->full[ps.stage]);
    const uint32_t sb = smem_u32(smem + ps.stage * kStageBytes);
    mbar_arrive_expect_tx(fb, kStageBytes);
    tma_load_2d(sb, tmA, fb, kb * SBK, arow);                        // A_hi
    tma_load_2d(sb + kTile, tmA, fb, lo_off + kb * SBK, arow);       // A_lo
    tma_load_2d(sb + 2 * kTile, tmB, fb, kb * SBK, brow);            // W_hi
    tma_load_2d(sb + 3 * kTile, tmB, fb, lo_off + kb * SBK, brow);   // W_lo
    if (++ps.stage == kStages) { ps.stage = 0; ps.phase ^= 1u; }
  }
}

// MMA issuer (one thread): the same K range, chunk by chunk; `chunk` counts chunks over the whole kernel
__device__ __forceinline__ void issue(uint8_t* smem, Barriers* bars, Pipe& ps, uint32_t& chunk, uint32_t tmem_base,
                                      int k0, int k1) {
  for (int kc = k0; kc < k1; kc += kChunk, ++chunk) {
    const int as = (int)(chunk & 1u);
    const uint32_t aphase = (chunk >> 1) & 1u;
    mbar_wait(smem_u32(&bars->acc_empty[as]), aphase ^ 1u, 1);
    tc_fence_after();
    const uint32_t d_main = tmem_base + (uint32_t)(as * 2 * SBN);
    const uint32_t d_corr = d_main + (uint32_t)SBN;
    const int kend = min(kc + kChunk, k1);
    for (int kb = kc; kb < kend; ++kb) {
      mbar_wait(smem_u32(&bars->full[ps.stage]), ps.phase, 2);
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + ps.stage * kStageBytes);
      const uint64_t a_hi = umma_desc_sw128(sb), a_lo = umma_desc_sw128(sb + kTile);
      const uint64_t b_hi = umma_desc_sw128(sb + 2 * kTile), b_lo = umma_desc_sw128(sb + 3 * kTile);
      const uint32_t later = (uint32_t)(kb != kc);   // 0 on the chunk's first K block: overwrite the accumulators
#pragma unroll
      for (int k4 = 0; k4 < SBK / SUK; ++k4)   // +32 bytes per K=16 slice inside the swizzle row: +2 in the >>4 address field
        umma_f16(d_corr, a_lo + (uint64_t)(2 * k4), b_hi + (uint64_t)(2 * k4), kIdesc, later | (uint32_t)(k4 != 0));
#pragma unroll
      for (int k4 = 0; k4 < SBK / SUK; ++k4)
        umma_f16(d_corr, a_hi + (uint64_t)(2 * k4), b_lo + (uint64_t)(2 * k4), kIdesc, 1u);
#pragma unroll
      for (int k4 = 0; k4 < SBK / SUK; ++k4)
        umma_f16(d_main, a_hi + (uint64_t)(2 * k4), b_hi + (uint64_t)(2 * k4), kIdesc, later | (uint32_t)(k4 != 0));
      umma_commit(smem_u32(&bars->empty[ps.stage]));   // frees the smem slot when these MMAs retire
      if (++ps.stage == kStages) { ps.stage = 0; ps.phase ^= 1u; }
    }
    umma_commit(smem_u32(&bars->acc_full[as]));
  }
}

// Epilogue warps: promote every chunk of the K range into this thread's 128 FP32 accumulators (thread = TMEM lane =
// tile row quad * 32 + lane).  Round-to-nearest adds, fixed order: deterministic.
__device__ __forceinline__ void collect(Barriers* bars, uint32_t& chunk, uint32_t tmem_base, int quad, int lane, int k0,
                                        int k1, float (&acc)[SBN]) {
#pragma unroll
  for (int j = 0; j < SBN; ++j) acc[j] = 0.f;
  for (int kc = k0; kc < k1; kc += kChunk, ++chunk) {
    const int as = (int)(chunk & 1u);
    const uint32_t aphase = (chunk >> 1) & 1u;
    mbar_wait(smem_u32(&bars->acc_full[as]), aphase, 3);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t)(as * 2 * SBN) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
    for (int ch = 0; ch < SBN / 32; ++ch) {
      uint32_t v[32], w[32];
      tmem_ld32(taddr + (uint32_t)(ch * 32), v);
      tmem_ld32(taddr + (uint32_t)(SBN + ch * 32), w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        acc[ch * 32 + j] = __fadd_rn(acc[ch * 32 + j], __fadd_rn(__uint_as_float(v[j]), __uint_as_float(w[j])));
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[as]));
  }
}

inline size_t smem_bytes(size_t extra) { return 1024 + (size_t)kStages * kStageBytes + extra + sizeof(Barriers); }

}  // namespace split
}  // namespace ladine
