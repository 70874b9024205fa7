// Host-side structures shared by the translation units of libladine.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "ladine_common.cuh"

struct ladine_member {
  int F = 0;        // feature_dim as given
  int Fp = 0;       // padded feature dim (multiple of 32 for FP32-resident, 256 for tensor path)
  int C = 0;        // num_classes
  int Cp = 0;       // padded class count {2,4,8,16}
  int T = 0;        // usable table rows
  int guidance = 0;
  int precision = 0;  // resolved ladine_precision (never AUTO)
  int device = 0;
  // FP32 per-step scale/shift rows, [T, Fp] each.  Tensor path: pre-multiplied by log2(e).
  float* A[3] = {nullptr, nullptr, nullptr};
  float* Cc[3] = {nullptr, nullptr, nullptr};
  float* W1y = nullptr;  // [Fp, Cp]  lin1 weight columns that multiply y_t
  float* W1g = nullptr;  // [Fp, Cp]  lin1 weight columns that multiply y_0_hat (zeros without guidance)
  float* W4 = nullptr;   // [Cp, Fp]
  float* b4 = nullptr;   // [Cp]
  // FP32-resident path: k-major (transposed) FP32 weights  Wt[k * Fp + n] = W[n][k]
  float* W2t = nullptr;
  float* W3t = nullptr;
  // tensor path: 16-bit [Fp, Fp] row-major ([out][in], K contiguous).  FP32X: [Fp, 2 * Fp] FP16, columns [0, Fp) hold
  // the hi part and [Fp, 2 * Fp) the lo part of W * wscale[l] (wscale = a power of two that lifts max|W| to ~2^14 so
  // that the lo parts stay normal FP16 numbers; its inverse is folded into the A_l scale rows)
  void* W2h = nullptr;
  void* W3h = nullptr;
  float* wscale = nullptr;   // FP32X: device [4] = scale2, scale3, 1/scale2, 1/scale3
  int split = 0;             // 1 for FP32X
  uint64_t bytes = 0;
};

struct ladine_handle {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  std::string err;
  int64_t last_launches = 0;
  // grow-only workspace
  void* ws = nullptr;
  uint64_t ws_bytes = 0;
  // grow-only workspace of the encoder prologue (ladine_encode)
  void* enc_ws = nullptr;
  uint64_t enc_ws_bytes = 0;
  int64_t last_encoder_launches = 0;
  // driver entry point for tensor-map encoding (resolved lazily through the runtime)
  void* encode_tiled = nullptr;
  // optional per-kernel event timing (tensor path)
  // lanes: independent member groups run concurrently on internal streams (tensor path)
  static constexpr int kMaxLanes = 4;
  int lanes = 1;
  // fold the tail + head of each step into the layer-3 GEMM kernel (helper warps; single-lane chains, <= 8 classes).
  // Bitwise identical to the separate tail/head kernel; measured 1-3 % SLOWER at config 2 (the helpers compete with
  // the SMEM-port-bound mainloop and the last row groups drain after the last MMA), so it is opt-in.
  bool fuse = false;
  // small calls (<= 4 members x <= 128 chains) on the whole-chain persistent kernel (ladine_persist.cu): opt-in because its
  // split-K sums differ from the tile kernels' by FP32 rounding noise (the tile path is bitwise partition-invariant)
  bool persist = false;
  bool persist_debug = false;   // block 0 of the persistent kernel prints its per-phase clock totals
  int ctas = 0;             // 0 = choose per call, 1 = cta_group::1, 2 = cta_group::2 CTA pairs
  int order = 0;             // GEMM tile order: 0 = auto, 1 = N-tile-major, 2 = row-major
  int tail_vec = 0;          // tail/head features per thread: 0 = default (4), else 4 or 8 (A/B timing)
  double pair_gain = 1.08;   // measured throughput ratio pair/single per useful tile (see choose_ctas)
  cudaStream_t lane_stream[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};  // [0] unused: caller's stream
  // end of the last ladine_sample / ladine_encode on this handle: the next call's stream waits for it before it touches
  // the shared workspace, so calls issued on DIFFERENT streams are ordered instead of overwriting each other
  cudaEvent_t ev_done = nullptr;
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_join[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  bool profiling = false;
  struct Span { cudaEvent_t a, b; int kind; };
  std::vector<Span> spans;       // recorded since the last ladine_get_profile
  std::vector<cudaEvent_t> pool; // reusable events
};

namespace ladine {

inline int cpad_of(int C) { return C <= 2 ? 2 : C <= 4 ? 4 : C <= 8 ? 8 : 16; }

// ---- FP32 SMEM-resident path (ladine_resident.cu) ----
cudaError_t launch_resident(const ladine_handle* h, const ladine_member* const* members, const ladine_sample_args& a,
                            const ChainIds& ids, const StepCoef* d_coef, float* d_u, int n_slots, int n_traj,
                            cudaStream_t st, int64_t* launches);
size_t resident_smem_bytes(int Fp, int Cp);

// ---- tensor-core path (ladine_tensor.cu) ----
struct TensorWorkspace {
  void* h1;      // [K * rows_pad, Fp] 16-bit
  void* h2;      // [K * rows_pad, Fp] 16-bit
  float* part;   // [K * rows_pad, Fp / 256, 2, Cp]  (one lin4 partial per 128-column slot)
  float* ybuf;   // 2 x [K * rows_pad, Cp] (ping-pong chain state)
  float* u;      // [K, N, Fp]
  int32_t* sched;  // static tile schedules of this lane's GEMM launches (layer 2 | layer 3)
  int* arrivals;   // row-group arrival counters of the fused tail + head
};
struct TensorChain;  // one lane: a group of members advancing through the reverse steps on one stream
TensorChain* tensor_chain_create(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args& a,
                                 const ChainIds& ids, const StepCoef* h_coef, const TensorWorkspace& ws, int n_slots,
                                 int n_traj, cudaStream_t st, bool single_lane, int64_t* launches, std::string* err,
                                 cudaError_t* status);
cudaError_t tensor_chain_step(TensorChain* c, int t, int64_t* launches);
void tensor_chain_destroy(TensorChain* c);
cudaError_t launch_debug_layer(ladine_handle* h, const ladine_member* m, int layer, int t, const void* h_in, int rows,
                               void* h_out, float* part, int32_t* sched_buf, cudaStream_t st, std::string* err);
size_t sched_bytes_bound(int K, int rows, int NB);
int64_t debug_plan(int K, int rows, int Fp, int geometry, int row_major, int units, int32_t* table_out, int64_t cap,
                   int32_t info_out[4]);
int debug_geometry(int K, int rows, int Fp, int sm_count);
size_t tensor_gemm_smem_bytes(int Cp);

// ---- whole-chain persistent kernel for small calls (ladine_persist.cu) ----
int persist_splits(const ladine_handle* h, int K, int rows, int Fp, int C);   // K slices per layer, 0 = not applicable
size_t persist_workspace_bytes(int K, int Fp, int Cp, int S);
cudaError_t launch_persistent_chain(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args& a,
                                    const ChainIds& ids, const StepCoef* d_coef, const float* d_u, uint8_t* ws, int S,
                                    int n_slots, int n_traj, cudaStream_t st, int64_t* launches, std::string* err);

// ---- cross-stream ordering of the calls that share a handle's workspaces (ladine_api.cu) ----
cudaError_t order_after_previous_call(ladine_handle* h, cudaStream_t st);
void mark_call_done(ladine_handle* h, cudaStream_t st);

// ---- packed images: a packed member / encoder serialised for an on-disk cache ----
// 128-byte header + the packed device buffers back to back in declaration order.  `layout` changes whenever a packing
// kernel or a buffer layout changes, so a stale file is refused instead of mis-read.
constexpr uint32_t kImageLayout = 1;
struct ImageHeader {
  char magic[8];            // "LADINEM" / "LADINEE"
  uint32_t abi;             // LADINE_ABI_VERSION of the writer
  uint32_t layout;          // kImageLayout
  int32_t dims[20];         // the scalar fields of the packed object (see image_export / image_import)
  uint64_t payload_bytes;   // bytes after the header
  uint64_t checksum;        // image_checksum(payload)
  uint64_t reserved;
};
static_assert(sizeof(ImageHeader) == 120, "ImageHeader layout");
constexpr size_t kImageHeaderBytes = 128;   // the header padded so that the payload stays 16-byte aligned
struct ImageSection {
  void** dev;      // address of the owning pointer inside the packed object
  size_t bytes;
};
uint64_t image_checksum(const void* p, size_t n);
// D2H of every section behind a header into host_dst (synchronises `st`); returns the bytes written or 0 (cap too small)
uint64_t image_export(const char* magic, const int32_t* dims, int n_dims, const std::vector<ImageSection>& secs,
                      void* host_dst, uint64_t cap, cudaStream_t st, cudaError_t* status);
// header checks of an image; on success *hdr_out points into host_src
const char* image_check(const void* host_src, uint64_t bytes, const char* magic, const ImageHeader** hdr_out);
// cudaMalloc + H2D of every section (synchronises `st`); on failure everything allocated here is released
cudaError_t image_import(const void* host_src, const std::vector<ImageSection>& secs, cudaStream_t st);

// ---- shared small kernels (ladine_api.cu) ----
cudaError_t launch_guidance_u(const ladine_member* const* members, int K, int N, const float* y0hat, float* u,
                              cudaStream_t st);

}  // namespace ladine
