// Step-invariant encoder prologue  xf = norm(encoder_x(x))  of the 'linear' ConditionalModel encoder
// (latent_model.py:126-135 -- Linear(Dx,H) BN Softplus Linear(H,H) BN Softplus Linear(H,F) -- and :155, :170-171 norm),
// SURVEY.md §8f-3.  The reference evaluates it in FP32 inside EVERY reverse step; here it runs once per call, FP32-grade,
// on the tensor cores:
//
//   every GEMM operand is split into FP16 hi + lo parts after scaling by a power of two (weights at pack time,
//   activations per call), and per 64-wide K block three groups of tcgen05.mma are issued:
//       corr += A_lo . W_hi      corr += A_hi . W_lo      main += A_hi . W_hi          (FP32 accumulators in TMEM)
//   The tensor core truncates each time it adds into an FP32 accumulator (~half an ulp of the running sum per
//   instruction, a bias that grows linearly with K -- 150528 for the image-sized first layer), so accumulation in TMEM
//   is limited to chunks of 8 K blocks; the epilogue warps promote every chunk (main + corr) into FP32 REGISTER
//   accumulators with round-to-nearest adds while the next chunk runs in the other TMEM stage.
//
// The first layer streams 2.35 GB of weights per member (HBM-bound for the batch sizes of the test path), so its K range
// is split over the SMs: work item = (row tile, 128-column tile, K split); each item writes an FP32 partial tile and a small
// finish kernel sums the splits in fixed order (deterministic), adds the bias, applies the eval-mode BatchNorm and
// softplus.  Layers 2 and 3 (4096 x 4096) reuse the same two kernels.
#include <algorithm>
#include <atomic>
#include <cstring>

#include "ladine_split.cuh"
#include "ladine_tensor.cuh"

struct ladine_encoder {
  int Dx = 0, H = 0, F = 0;          // data_dim, hidden_dim, feature_dim
  int Kp[3] = {0, 0, 0};             // padded K (multiple of 64) of the three layers
  int Np[3] = {0, 0, 0};             // padded output width (multiple of 128)
  int Nout[3] = {0, 0, 0};
  int device = 0;
  __half* Ws[3] = {nullptr, nullptr, nullptr};   // [Np, 2 * Kp]: hi | lo of W * wscale
  float* wscale = nullptr;                       // device [3][2] = scale, 1/scale  (+ 3 abs-max scratch words)
  float* vec[3][5] = {};                         // per layer: bias, bn weight, bn bias, bn mean, bn var  (owned copies)
  float eps = 1e-5f;
  uint64_t bytes = 0;
};

namespace ladine {
namespace {

using namespace split;
constexpr int EBM = SBM, EBN = SBN, EBK = SBK;

struct EncGemmParams {
  CUtensorMap tmA;    // activations [rows, 2 * Kp] FP16 (hi | lo), box 64 x 128, rows beyond `rows` read as zero
  CUtensorMap tmB;    // weights     [Np,   2 * Kp] FP16 (hi | lo), box 64 x 128
  float* partial;     // [S, rows, Np]
  int rows, Np, Kp, KB, MB, NB, S;
};

// work item = (row tile, 128-column tile, K split)
struct EncItem {
  int mb, nb, s, k0, k1;
  __device__ EncItem(const EncGemmParams& p, int item) {
    mb = item % p.MB;           // row tiles of one weight tile are adjacent: they run concurrently and share it in L2
    const int t = item / p.MB;
    s = t % p.S;
    nb = t / p.S;
    k0 = (int)((long long)s * p.KB / p.S);
    k1 = (int)((long long)(s + 1) * p.KB / p.S);
  }
};

__global__ void __launch_bounds__(kThreads, 1) enc_gemm_kernel(const __grid_constant__ EncGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    init_barriers(bars);
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 0) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int n_items = p.MB * p.NB * p.S;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      Pipe ps;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const EncItem it(p, item);
        produce(smem, bars, ps, &p.tmA, &p.tmB, it.mb * EBM, it.nb * EBN, p.Kp, it.k0, it.k1);
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      Pipe ps;
      uint32_t chunk = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const EncItem it(p, item);
        issue(smem, bars, ps, chunk, tmem_base, it.k0, it.k1);
      }
    }
  } else if (warp < 4) {
    const int quad = warp;
    uint32_t chunk = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const EncItem it(p, item);
      float acc[EBN];
      collect(bars, chunk, tmem_base, quad, lane, it.k0, it.k1, acc);
      const int row = it.mb * EBM + quad * 32 + lane;
      if (row < p.rows) {
        float* dst = p.partial + ((size_t)it.s * p.rows + row) * p.Np + it.nb * EBN;
#pragma unroll
        for (int q = 0; q < EBN / 8; ++q) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(acc[8 * q + i]);
          st_global_256(dst + 8 * q, o);
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

// out[m, n] = act(BN(sum_s partial[s][m, n] / (scale_a * scale_w) + bias[n]))   (fixed summation order: deterministic)
struct EncFinishParams {
  const float* partial;   // [S, rows, Np]
  const float* scale_a;   // device scalars: [0] = scale, [1] = 1 / scale (activation operand of this layer)
  const float* scale_w;   // same for the weights
  const float *bias, *bn_w, *bn_b, *bn_mean, *bn_var;
  float* out;             // [rows, ld_out]
  float eps;
  int S, rows, Np, Nout, ld_out, softplus;
};
__global__ void enc_finish_kernel(const __grid_constant__ EncFinishParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.rows * p.Nout) return;
  const int n = (int)(i % p.Nout);
  const size_t m = i / p.Nout;
  float z = 0.f;
  for (int s = 0; s < p.S; ++s) z = __fadd_rn(z, p.partial[((size_t)s * p.rows + m) * p.Np + n]);
  z = z * (p.scale_a[1] * p.scale_w[1]);          // powers of two: exact
  z = __fadd_rn(z, p.bias[n]);
  // eval-mode BatchNorm1d as ATen's CUDA kernel writes it: (x - mean) * invstd * weight + bias
  const float invstd = rsqrtf(p.bn_var[n] + p.eps);
  z = (z - p.bn_mean[n]) * invstd * p.bn_w[n] + p.bn_b[n];
  if (p.softplus) z = softplus_precise(z);
  p.out[m * p.ld_out + n] = z;
}

// max |v| over a buffer (non-negative floats order like their bit patterns)
__global__ void enc_absmax_kernel(const float* __restrict__ v, size_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(v[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// scale = 2^e with max * 2^e in [2^13, 2^14): hi keeps 11 significant bits, lo (<= 2^-11 of hi) stays a normal FP16
// number for every element that matters, nothing overflows.  out[0] = scale, out[1] = 1 / scale.
__global__ void enc_scale_kernel(const unsigned int* __restrict__ absmax_bits, float* __restrict__ out) {
  const float m = __uint_as_float(*absmax_bits);
  int e = 0;
  if (m > 0.f && isfinite(m)) {
    int me;
    frexpf(m, &me);
    e = 14 - me;
  }
  e = max(-60, min(60, e));
  out[0] = ldexpf(1.0f, e);
  out[1] = ldexpf(1.0f, -e);
}
// dst[r, k] = fp16(v * s), dst[r, Kp + k] = fp16(v * s - hi); rows >= rows_src and columns >= K are zero
__global__ void enc_split_kernel(const float* __restrict__ src, int rows_src, int K, size_t ld_src,
                                 const float* __restrict__ scale, int rows_dst, int Kp, __half* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows_dst * Kp) return;
  const int k = (int)(i % Kp);
  const size_t r = i / Kp;
  const float v = (r < (size_t)rows_src && k < K) ? src[r * ld_src + k] * scale[0] : 0.f;
  const __half hi = Pack16<__half>::one(v);
  const __half lo = Pack16<__half>::one(v - __half2float(hi));
  dst[r * 2 * Kp + k] = hi;
  dst[r * 2 * Kp + Kp + k] = lo;
}

inline uint64_t up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

size_t enc_gemm_smem() { return smem_bytes(0); }

// K splits of one layer: fill the SMs when there are few tiles (small batches stream the weights once, HBM-bound)
int enc_splits(int sm_count, int MB, int NB, int KB) {
  int S = sm_count / (MB * NB);
  const int smax = KB / (2 * kChunk) > 0 ? KB / (2 * kChunk) : 1;   // at least two chunks per split
  if (S > smax) S = smax;
  return S < 1 ? 1 : S;
}

}  // namespace
}  // namespace ladine

using namespace ladine;

namespace {
int efail(ladine_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
struct EncDeviceGuard {
  int prev = -1;
  explicit EncDeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~EncDeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};
void free_encoder_buffers(ladine_encoder* e) {
  for (int l = 0; l < 3; ++l) {
    cudaFree(e->Ws[l]);
    for (int v = 0; v < 5; ++v) cudaFree(e->vec[l][v]);
  }
  cudaFree(e->wscale);
}
}  // namespace

extern "C" {

int ladine_pack_encoder(ladine_handle* h, const ladine_encoder_desc* d, void* stream, ladine_encoder** out) {
  if (!h) return LADINE_ERR_INVALID;
  if (!d || !out) return efail(h, LADINE_ERR_INVALID, "null encoder descriptor or output");
  *out = nullptr;
  if (d->struct_size != sizeof(ladine_encoder_desc)) return efail(h, LADINE_ERR_INVALID, "ladine_encoder_desc size mismatch");
  if (d->data_dim < 1 || d->hidden_dim < 1 || d->feature_dim < 1)
    return efail(h, LADINE_ERR_INVALID, "data_dim, hidden_dim, feature_dim must be >= 1");
  for (int l = 0; l < 3; ++l) {
    if (!d->lin_w[l] || !d->lin_b[l] || !d->bn_w[l] || !d->bn_b[l] || !d->bn_mean[l] || !d->bn_var[l])
      return efail(h, LADINE_ERR_INVALID, "null parameter pointer in ladine_encoder_desc");
  }
  EncDeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ladine_encoder* e = new (std::nothrow) ladine_encoder();
  if (!e) return efail(h, LADINE_ERR_NOMEM, "host allocation failed");
  e->Dx = d->data_dim;
  e->H = d->hidden_dim;
  e->F = d->feature_dim;
  e->eps = d->bn_eps;
  e->device = h->device;
  const int K[3] = {e->Dx, e->H, e->H};
  const int N[3] = {e->H, e->H, e->F};
  cudaError_t ce = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
  ok(cudaMalloc(&e->wscale, 16 * sizeof(float)));   // [l][2] scales at 0..5, abs-max scratch words at 8..10
  for (int l = 0; l < 3; ++l) {
    e->Kp[l] = (int)up(K[l], EBK);
    e->Np[l] = (int)up(N[l], EBN);
    e->Nout[l] = N[l];
    const size_t wb = (size_t)e->Np[l] * 2 * e->Kp[l] * sizeof(__half);
    ok(cudaMalloc(&e->Ws[l], wb));
    e->bytes += wb;
    for (int v = 0; v < 5; ++v) ok(cudaMalloc(&e->vec[l][v], (size_t)N[l] * sizeof(float)));
    e->bytes += 5ull * N[l] * sizeof(float);
  }
  if (ce != cudaSuccess) {
    cudaGetLastError();
    free_encoder_buffers(e);
    delete e;
    return efail(h, LADINE_ERR_NOMEM, "device allocation failed while packing an encoder");
  }
  unsigned int* amax = reinterpret_cast<unsigned int*>(e->wscale + 8);
  cudaMemsetAsync(amax, 0, 4 * sizeof(unsigned int), st);
  for (int l = 0; l < 3; ++l) {
    const float* srcv[5] = {d->lin_b[l], d->bn_w[l], d->bn_b[l], d->bn_mean[l], d->bn_var[l]};
    for (int v = 0; v < 5; ++v)
      cudaMemcpyAsync(e->vec[l][v], srcv[v], (size_t)N[l] * sizeof(float), cudaMemcpyDeviceToDevice, st);
    enc_absmax_kernel<<<592, 256, 0, st>>>(d->lin_w[l], (size_t)N[l] * K[l], amax + l);
    enc_scale_kernel<<<1, 1, 0, st>>>(amax + l, e->wscale + 2 * l);
    const size_t tot = (size_t)e->Np[l] * e->Kp[l];
    enc_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d->lin_w[l], N[l], K[l], (size_t)K[l], e->wscale + 2 * l,
                                                                    e->Np[l], e->Kp[l], e->Ws[l]);
  }
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    free_encoder_buffers(e);
    delete e;
    char buf[200];
    snprintf(buf, sizeof buf, "encoder packing kernels: %s", cudaGetErrorString(ce));
    return efail(h, LADINE_ERR_CUDA, buf);
  }
  *out = e;
  return LADINE_OK;
}

int ladine_free_encoder(ladine_handle* h, ladine_encoder* e) {
  if (!e) return LADINE_ERR_INVALID;
  EncDeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  free_encoder_buffers(e);
  delete e;
  (void)h;
  return LADINE_OK;
}

uint64_t ladine_encoder_bytes(const ladine_encoder* e) { return e ? e->bytes : 0; }
int ladine_encoder_dims(const ladine_encoder* e, int32_t dims_out[4]) {
  if (!e || !dims_out) return LADINE_ERR_INVALID;
  const int32_t d[4] = {e->Dx, e->H, e->F, e->device};
  std::memcpy(dims_out, d, sizeof d);
  return LADINE_OK;
}

// ---- packed-encoder images ----
static std::vector<ImageSection> encoder_sections(ladine_encoder* e) {
  std::vector<ImageSection> s;
  s.push_back({reinterpret_cast<void**>(&e->wscale), 16 * sizeof(float)});
  for (int l = 0; l < 3; ++l) {
    s.push_back({reinterpret_cast<void**>(&e->Ws[l]), (size_t)e->Np[l] * 2 * e->Kp[l] * sizeof(__half)});
    for (int v = 0; v < 5; ++v) s.push_back({reinterpret_cast<void**>(&e->vec[l][v]), (size_t)e->Nout[l] * sizeof(float)});
  }
  return s;
}

uint64_t ladine_encoder_image_bytes(const ladine_encoder* e) {
  if (!e) return 0;
  uint64_t total = kImageHeaderBytes;
  for (const auto& s : encoder_sections(const_cast<ladine_encoder*>(e))) total += (s.bytes + 15) / 16 * 16;
  return total;
}

int ladine_encoder_export(ladine_handle* h, const ladine_encoder* e, void* host_dst, uint64_t capacity, void* stream) {
  if (!h) return LADINE_ERR_INVALID;
  if (!e || !host_dst) return efail(h, LADINE_ERR_INVALID, "null encoder or destination");
  if (capacity < ladine_encoder_image_bytes(e))
    return efail(h, LADINE_ERR_INVALID, "destination smaller than ladine_encoder_image_bytes");
  EncDeviceGuard guard(e->device);
  int32_t eps_bits;
  std::memcpy(&eps_bits, &e->eps, 4);
  const int32_t dims[4] = {e->Dx, e->H, e->F, eps_bits};
  cudaError_t ce = cudaSuccess;
  const uint64_t n = image_export("LADINEE", dims, 4, encoder_sections(const_cast<ladine_encoder*>(e)), host_dst, capacity,
                                  static_cast<cudaStream_t>(stream), &ce);
  if (ce != cudaSuccess || n == 0) return efail(h, LADINE_ERR_CUDA, std::string("encoder export: ") + cudaGetErrorString(ce));
  return LADINE_OK;
}

int ladine_encoder_import(ladine_handle* h, const void* host_src, uint64_t bytes, void* stream, ladine_encoder** out) {
  if (!h) return LADINE_ERR_INVALID;
  if (!out) return efail(h, LADINE_ERR_INVALID, "null output");
  *out = nullptr;
  const ImageHeader* hd = nullptr;
  if (const char* why = image_check(host_src, bytes, "LADINEE", &hd)) return efail(h, LADINE_ERR_INVALID, why);
  ladine_encoder* e = new (std::nothrow) ladine_encoder();
  if (!e) return efail(h, LADINE_ERR_NOMEM, "host allocation failed");
  e->Dx = hd->dims[0]; e->H = hd->dims[1]; e->F = hd->dims[2];
  std::memcpy(&e->eps, &hd->dims[3], 4);
  e->device = h->device;
  if (e->Dx < 1 || e->H < 1 || e->F < 1) {
    delete e;
    return efail(h, LADINE_ERR_INVALID, "packed encoder image: inconsistent dimensions");
  }
  const int K[3] = {e->Dx, e->H, e->H};
  const int N[3] = {e->H, e->H, e->F};
  for (int l = 0; l < 3; ++l) {
    e->Kp[l] = (int)up(K[l], EBK);
    e->Np[l] = (int)up(N[l], EBN);
    e->Nout[l] = N[l];
  }
  auto secs = encoder_sections(e);
  uint64_t payload = 0;
  for (const auto& s : secs) payload += (s.bytes + 15) / 16 * 16;
  if (payload != hd->payload_bytes) {
    delete e;
    return efail(h, LADINE_ERR_INVALID, "packed encoder image: inconsistent dimensions");
  }
  EncDeviceGuard guard(h->device);
  cudaError_t ce = image_import(host_src, secs, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) {
    delete e;
    return efail(h, ce == cudaErrorMemoryAllocation ? LADINE_ERR_NOMEM : LADINE_ERR_CUDA,
                 std::string("encoder import: ") + cudaGetErrorString(ce));
  }
  for (int l = 0; l < 3; ++l) e->bytes += (size_t)e->Np[l] * 2 * e->Kp[l] * sizeof(__half) + 5ull * N[l] * sizeof(float);
  *out = e;
  return LADINE_OK;
}

int ladine_encode(ladine_handle* h, const ladine_encoder* const* encoders, int32_t K, const float* x, int32_t N,
                  float* xf_out, void* stream) {
  if (!h) return LADINE_ERR_INVALID;
  if (!encoders || K < 1 || !x || N < 1 || !xf_out) return efail(h, LADINE_ERR_INVALID, "bad ladine_encode arguments");
  const ladine_encoder* e0 = encoders[0];
  for (int k = 0; k < K; ++k) {
    const ladine_encoder* e = encoders[k];
    if (!e) return efail(h, LADINE_ERR_INVALID, "null encoder");
    if (e->device != h->device) return efail(h, LADINE_ERR_INVALID, "encoder packed on another device");
    if (e->Dx != e0->Dx || e->H != e0->H || e->F != e0->F)
      return efail(h, LADINE_ERR_INVALID, "encoders of one call must share data_dim, hidden_dim, feature_dim");
  }
  EncDeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (order_after_previous_call(h, st) != cudaSuccess) return efail(h, LADINE_ERR_CUDA, "stream ordering event");
  std::string err;
  cudaError_t ce = resolve_encode(h, &err);
  if (ce != cudaSuccess) return efail(h, LADINE_ERR_CUDA, err);

  const int rows = N;
  const int MB = (rows + EBM - 1) / EBM;
  int S[3], KB[3], NB[3];
  size_t part_floats = 0;
  for (int l = 0; l < 3; ++l) {
    KB[l] = e0->Kp[l] / EBK;
    NB[l] = e0->Np[l] / EBN;
    S[l] = enc_splits(h->sm_count, MB, NB[l], KB[l]);
    part_floats = std::max(part_floats, (size_t)S[l] * rows * e0->Np[l]);
  }
  // workspace: x split | partial | hidden FP32 | hidden split | activation scales (+ abs-max words)
  uint64_t off = 0;
  const uint64_t o_xs = off; off = up(off + (uint64_t)rows * 2 * e0->Kp[0] * 2, 1024);
  const uint64_t o_part = off; off = up(off + part_floats * 4, 1024);
  const uint64_t o_hid = off; off = up(off + (uint64_t)rows * e0->Np[1] * 4, 1024);
  const uint64_t o_hs = off; off = up(off + (uint64_t)rows * 2 * e0->Kp[1] * 2, 1024);
  const uint64_t o_sc = off; off = up(off + 64 * sizeof(float), 1024);
  if (off > h->enc_ws_bytes) {
    if (h->enc_ws) {
      // earlier stream-ordered work may still read the old buffer: wait for the handle's last call
      if (h->ev_done) cudaEventSynchronize(h->ev_done);
      else cudaDeviceSynchronize();
      cudaFree(h->enc_ws);
      h->enc_ws = nullptr;
      h->enc_ws_bytes = 0;
    }
    if (cudaMalloc(&h->enc_ws, off) != cudaSuccess) {
      cudaGetLastError();
      return efail(h, LADINE_ERR_NOMEM, "encoder workspace allocation failed");
    }
    h->enc_ws_bytes = off;
  }
  uint8_t* ws = static_cast<uint8_t*>(h->enc_ws);
  __half* xs = reinterpret_cast<__half*>(ws + o_xs);
  float* partial = reinterpret_cast<float*>(ws + o_part);
  float* hid = reinterpret_cast<float*>(ws + o_hid);
  __half* hs = reinterpret_cast<__half*>(ws + o_hs);
  float* sc = reinterpret_cast<float*>(ws + o_sc);            // [0..1] scale of x, [2..3] scale of the hidden operand
  unsigned int* amax = reinterpret_cast<unsigned int*>(sc + 8);

  static std::atomic<uint64_t> configured{0};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = dev < 64 ? (uint64_t)1 << dev : 0;
    if (!(configured.load() & bit)) {
      ce = cudaFuncSetAttribute(enc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_gemm_smem());
      if (ce != cudaSuccess) return efail(h, LADINE_ERR_CUDA, "encoder kernel shared-memory configuration failed");
      configured.fetch_or(bit);
    }
  }
  int64_t launches = 0;
  auto split_operand = [&](const float* src, int K_in, size_t ld_src, int Kp, float* scale, unsigned int* am, __half* dst) {
    cudaMemsetAsync(am, 0, sizeof(unsigned int), st);
    // abs-max over the valid region only when the rows are dense (ld == K); otherwise row by row via the 2-D grid
    enc_absmax_kernel<<<592, 256, 0, st>>>(src, (size_t)rows * ld_src, am);
    enc_scale_kernel<<<1, 1, 0, st>>>(am, scale);
    const size_t tot = (size_t)rows * Kp;
    enc_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, rows, K_in, ld_src, scale, rows, Kp, dst);
    launches += 3;
  };
  auto gemm = [&](const __half* A, const ladine_encoder* e, int l) -> bool {
    EncGemmParams g{};
    if (!make_tmap(h, &g.tmA, A, (uint64_t)rows, (uint64_t)2 * e->Kp[l], EBM, false, &err)) return false;
    if (!make_tmap(h, &g.tmB, e->Ws[l], (uint64_t)e->Np[l], (uint64_t)2 * e->Kp[l], EBN, false, &err)) return false;
    g.partial = partial;
    g.rows = rows;
    g.Np = e->Np[l];
    g.Kp = e->Kp[l];
    g.KB = KB[l];
    g.MB = MB;
    g.NB = NB[l];
    g.S = S[l];
    const int items = MB * NB[l] * S[l];
    const int grid = items < h->sm_count ? items : h->sm_count;
    enc_gemm_kernel<<<grid, kThreads, enc_gemm_smem(), st>>>(g);
    launches += 1;
    return true;
  };
  auto finish = [&](const ladine_encoder* e, int l, const float* scale_a, float* out, int ld_out, int softplus) {
    EncFinishParams f{};
    f.partial = partial;
    f.scale_a = scale_a;
    f.scale_w = e->wscale + 2 * l;
    f.bias = e->vec[l][0];
    f.bn_w = e->vec[l][1];
    f.bn_b = e->vec[l][2];
    f.bn_mean = e->vec[l][3];
    f.bn_var = e->vec[l][4];
    f.out = out;
    f.eps = e->eps;
    f.S = S[l];
    f.rows = rows;
    f.Np = e->Np[l];
    f.Nout = e->Nout[l];
    f.ld_out = ld_out;
    f.softplus = softplus;
    const size_t tot = (size_t)rows * e->Nout[l];
    enc_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(f);
    launches += 1;
  };

  split_operand(x, e0->Dx, (size_t)e0->Dx, e0->Kp[0], sc, amax, xs);   // the images are split once for all members
  for (int k = 0; k < K; ++k) {
    const ladine_encoder* e = encoders[k];
    if (!gemm(xs, e, 0)) return efail(h, LADINE_ERR_CUDA, err);
    finish(e, 0, sc, hid, e->H, 1);
    split_operand(hid, e->H, (size_t)e->H, e->Kp[1], sc + 2, amax + 1, hs);
    if (!gemm(hs, e, 1)) return efail(h, LADINE_ERR_CUDA, err);
    finish(e, 1, sc + 2, hid, e->H, 1);
    split_operand(hid, e->H, (size_t)e->H, e->Kp[2], sc + 2, amax + 1, hs);
    if (!gemm(hs, e, 2)) return efail(h, LADINE_ERR_CUDA, err);
    finish(e, 2, sc + 2, xf_out + (size_t)k * rows * e->F, e->F, 0);
  }
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    char buf[200];
    snprintf(buf, sizeof buf, "encoder launch: %s", cudaGetErrorString(ce));
    return efail(h, LADINE_ERR_CUDA, buf);
  }
  h->last_encoder_launches = launches;
  mark_call_done(h, st);
  return LADINE_OK;
}

int64_t ladine_last_encoder_launches(const ladine_handle* h) { return h ? h->last_encoder_launches : 0; }

}  // extern "C"
