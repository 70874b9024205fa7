// Parameter blocks and tile geometry shared by the tensor-core kernels of the reverse step
// (ladine_tensor.cu: 16-bit operands; ladine_split.cu: FP32X split operands).
#pragma once
#include <atomic>
#include <string>

#include "ladine_internal.cuh"

namespace ladine {

constexpr int BM = 128;   // rows per tile  (UMMA M)
constexpr int BN = 256;   // cols per tile  (UMMA N)
constexpr int BK = 64;    // K per stage    (64 x 2 B = one 128-byte swizzle row)
constexpr int UK = 16;    // K per tcgen05.mma (kind::f16)

enum TailMode { kInit = 0, kMid = 1, kFinal = 2 };

struct TailHeadParams {
  const float* A1[LADINE_MAX_GROUP];   // row of the NEXT step (t-1), x log2e
  const float* C1[LADINE_MAX_GROUP];
  const float* W1y[LADINE_MAX_GROUP];  // [Fp, Cp]
  const float* b4[LADINE_MAX_GROUP];
  const float* part;   // [M_total, NB, Cp]
  const float* y_prev; // [M_total, Cp]  chain state before this step (ping-pong: column-split CTAs all read it)
  float* y_next;       // [M_total, Cp]  chain state after this step (written by column split 0)
  const float* xf;     // [K, N, Fin]
  const float* u;      // [K, N, Fp]
  const float* ytmean; // [K, N, C]
  const float* y_init; // [K, D, N, C] or null (kInit only)
  const float* noise;  // [K, D, S, N, C] or null
  void* h1;            // [M_total, Fp] 16-bit
  float* y_out;        // kFinal / last step
  float* traj_out;
  float* prob_out;
  float temperature;
  StepCoef coef;       // coefficients of the step being finished (unused for kInit)
  uint64_t seed;
  ChainIds ids;
  int t;               // table index of the step being finished (kInit: unused)
  int slot;            // noise slot / trajectory entry consumed-written by this launch
  int traj_entry;
  int N, D, C, Fin, Fp, NB, rows_pad, n_slots, n_traj, dchunk, write_out, colsplit;
  int ld_h1;           // row stride of h1 in elements: Fp, or 2 * Fp in split (FP32X) mode
  int split;           // FP32X: h1 is written as FP16 hi (columns [0, Fp)) + lo (columns [Fp, 2 * Fp)) parts
};

struct GemmParams {
  CUtensorMap tmA;                     // activations in  [M_total, Fp] 16-bit, box 64 x 128
  CUtensorMap tmB[LADINE_MAX_GROUP];   // member weights  [Fp, Fp]      16-bit, box 64 x (256 / CTAS)
  const float* scale[LADINE_MAX_GROUP];  // A_l[t] row of each member (already offset to row t), x log2e
  const float* shift[LADINE_MAX_GROUP];  // C_l[t] row, x log2e
  const float* W4[LADINE_MAX_GROUP];     // [Cp, Fp]  (layer 3 only)
  void* h_out;                         // layer 2: [M_total, Fp] 16-bit
  float* part;                         // layer 3: [M_total, NB, Cp]
  int Fp, NB, KB;                      // padded feature dim, N tiles (Fp/256), K blocks (Fp/64)
  int rows;                            // valid rows per member
  int rows_pad;                        // row stride between members (multiple of 128 * CTAS)
  // static tile schedule: unit u (a CTA, or a CTA pair) runs sched[u * sched_stride + 0, 1, ...] until a -1.
  // entry = member << 23 | nb << 13 | mb << 1 | half   (half: CTA-pair tile of 2 x 64 rows, M=128 MMAs)
  const int32_t* sched;
  int sched_stride;
  uint32_t idesc;                      // full tiles: M = 128 * CTAS
  uint32_t idesc_half;                 // pair half tiles: M = 128 (64 rows per CTA)
  // ---- layer 3 with the tail + head fused in (fuse != 0) ----
  // After a CTA has written its lin4 partials it signals the row group's arrival counter; when all NB column
  // tiles of the group are in, every one of those CTAs finishes the reverse step for its rows (eps, posterior
  // update -- recomputed identically by each) and produces h1 of the next step for ITS 256 columns.
  int fuse;
  int do_head;                         // 0 on the last step of the chain (no next h1)
  int mblk_total;                      // row tiles per member (full + half)
  int* group_arrivals;                 // [K * mblk_total * CTAS], monotonically increasing over the chain
  int arrivals_target;                 // NB * (layer-3 launches of this chain so far, this one included)
  TailHeadParams th;
};

struct TileCode {
  int member, nb, mb, half;
  __host__ __device__ static int32_t pack(int member, int nb, int mb, int half) {
    return (int32_t)((member << 23) | (nb << 13) | (mb << 1) | half);
  }
  __device__ explicit TileCode(int32_t c) : member(c >> 23), nb((c >> 13) & 1023), mb((c >> 1) & 4095), half(c & 1) {}
};


// ---- host helpers defined in ladine_tensor.cu ----
cudaError_t resolve_encode(ladine_handle* h, std::string* err);
// 2-D K-major tensor map: inner dim = cols (contiguous), outer = rows; box = 64 x box_rows; 128B swizzle
bool make_tmap(ladine_handle* h, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
               bool bf16, std::string* err);

// Run `fn` (a cudaFuncSetAttribute call) the first time a kernel instantiation is launched on the current device; `done`
// is that instantiation's bitmask over device ordinals.
template <typename Fn>
cudaError_t configure_once(std::atomic<uint64_t>& done, Fn fn) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = dev < 64 ? (uint64_t)1 << dev : 0;
  if (bit && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  e = fn();
  if (e == cudaSuccess && bit) done.fetch_or(bit, std::memory_order_release);
  return e;
}

// ---- FP32X kernels (ladine_split.cu) ----
// layer 2 / layer 3 of one reverse step on split operands; `p` as for the 16-bit kernels, with tmA over
// [M_total, 2 * Fp] and tmB over [Fp, 2 * Fp] (hi | lo halves), box 64 x 128, and a slim-geometry tile schedule
cudaError_t launch_split_gemm(int layer, const GemmParams& p, int grid, int Cp, cudaStream_t st);

}  // namespace ladine
