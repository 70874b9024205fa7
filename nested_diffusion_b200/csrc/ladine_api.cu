// C ABI of libladine (include/ladine.h): handle/member lifetime, member packing (fold + re-layout),
// and the dispatcher of the hot path.  No torch types, no exceptions across the boundary.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "ladine_internal.cuh"

using namespace ladine;

namespace {

int fail(ladine_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
int fail_cuda(ladine_handle* h, cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return fail(h, LADINE_ERR_CUDA, buf);
}

// ---------------------------------------------------------------------------------------------
// packing kernels
// ---------------------------------------------------------------------------------------------
// A_l[t,f] = E_l[t,f] * s_l[f];  C_l[t,f] = E_l[t,f] * (s_l[f] * b_l[f]) + (beta_l[f] - s_l[f] * mean_l[f])
// with s = bn.weight / sqrt(bn.running_var + eps)   (SURVEY.md §8a).  `gain` = log2(e) on the tensor path.
__global__ void fold_tables_kernel(const float* __restrict__ E, const float* __restrict__ lin_b,
                                   const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                   const float* __restrict__ bn_mean, const float* __restrict__ bn_var, float eps,
                                   float gain, const float* __restrict__ a_gain_dev, int T, int F, int Fp,
                                   float* __restrict__ A, float* __restrict__ Cc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)T * Fp) return;
  // FP32X: the GEMM runs on W * 2^s, so the scale row carries 2^-s (a power of two: exact)
  const float a_gain = a_gain_dev ? *a_gain_dev : 1.0f;
  const int f = (int)(i % Fp);
  const int t = (int)(i / Fp);
  float a = 0.f, c = 0.f;
  if (f < F) {
    const float s = bn_w[f] / sqrtf(bn_var[f] + eps);
    const float e = E[(size_t)t * F + f];
    a = e * s;
    c = e * (s * lin_b[f]) + (bn_b[f] - s * bn_mean[f]);
  }
  A[i] = a * gain * a_gain;
  Cc[i] = c * gain;
}

__global__ void pack_small_kernel(const float* __restrict__ lin1_w, const float* __restrict__ lin4_w,
                                  const float* __restrict__ lin4_b, int F, int Fp, int C, int Cp, int guidance,
                                  float* __restrict__ W1y, float* __restrict__ W1g, float* __restrict__ W4,
                                  float* __restrict__ b4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Fp * Cp) return;
  const int in1 = guidance ? 2 * C : C;
  {  // W1y / W1g : [Fp, Cp]
    const int f = i / Cp, c = i % Cp;
    float wy = 0.f, wg = 0.f;
    if (f < F && c < C) {
      wy = lin1_w[(size_t)f * in1 + c];
      if (guidance) wg = lin1_w[(size_t)f * in1 + C + c];
    }
    W1y[i] = wy;
    W1g[i] = wg;
  }
  {  // W4 : [Cp, Fp]
    const int c = i / Fp, f = i % Fp;
    W4[i] = (f < F && c < C) ? lin4_w[(size_t)c * F + f] : 0.f;
  }
  if (i < Cp) b4[i] = i < C ? lin4_b[i] : 0.f;
}

// Wt[k * Fp + n] = W[n * F + k]  (FP32, zero padded) for the SMEM-resident kernel
__global__ void transpose_pad_kernel(const float* __restrict__ W, int F, int Fp, float* __restrict__ Wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Fp * Fp) return;
  const int k = i / Fp, n = i % Fp;
  Wt[i] = (k < F && n < F) ? W[(size_t)n * F + k] : 0.f;
}

// W16[n * Fp + k] = round16(W[n * F + k]) (zero padded) for the tensor path ([out][in], K contiguous)
template <typename T16>
__global__ void convert_pad_kernel(const float* __restrict__ W, int F, int Fp, T16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Fp * Fp) return;
  const int n = (int)(i / Fp), k = (int)(i % Fp);
  out[i] = Pack16<T16>::one((n < F && k < F) ? W[(size_t)n * F + k] : 0.f);
}

// ---- FP32X: W = hi + lo in FP16 after scaling by a power of two ----
// max |W| over the [F, F] matrix (non-negative floats order like their bit patterns)
__global__ void absmax_kernel(const float* __restrict__ W, size_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(W[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// scale = 2^e with max|W| * 2^e in [2^13, 2^14): the hi part keeps 11 significant bits, the lo part (<= 2^-11 of hi)
// stays far above the FP16 subnormal range for every weight that matters, and nothing overflows (65504 ~ 2^16)
__global__ void split_scale_kernel(const unsigned int* __restrict__ absmax_bits, float* __restrict__ wscale) {
  const int l = threadIdx.x;
  if (l >= 2) return;
  const float m = __uint_as_float(absmax_bits[l]);
  int e = 0;
  if (m > 0.f && isfinite(m)) {
    int me;
    frexpf(m, &me);        // m = f * 2^me, f in [0.5, 1)
    e = 14 - me;           // m * 2^e in [2^13, 2^14)
  }
  e = max(-60, min(60, e));
  wscale[l] = ldexpf(1.0f, e);
  wscale[2 + l] = ldexpf(1.0f, -e);
}
// out[n, k] = fp16(w * s), out[n, Fp + k] = fp16(w * s - hi)      ([Fp, 2 * Fp], zero padded)
__global__ void convert_split_kernel(const float* __restrict__ W, int F, int Fp, const float* __restrict__ scale,
                                     __half* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Fp * Fp) return;
  const int n = (int)(i / Fp), k = (int)(i % Fp);
  const float w = (n < F && k < F) ? W[(size_t)n * F + k] * (*scale) : 0.f;
  const __half hi = Pack16<__half>::one(w);
  const __half lo = Pack16<__half>::one(w - __half2float(hi));
  out[(size_t)n * 2 * Fp + k] = hi;
  out[(size_t)n * 2 * Fp + Fp + k] = lo;
}

// u[k, n, f] = sum_c W1g[f, c] * y_0_hat[k, n, c]   (the step-invariant half of lin1)
struct GuidanceParams {
  const float* W1g[LADINE_MAX_GROUP];
  const float* y0hat;
  float* u;
  int N, C, Cp, Fp;
};
__global__ void guidance_u_kernel(const __grid_constant__ GuidanceParams p) {
  const int n = blockIdx.x, k = blockIdx.y;
  float yh[LADINE_MAX_CLASSES];
  for (int c = 0; c < p.C; ++c) yh[c] = p.y0hat[((size_t)k * p.N + n) * p.C + c];
  for (int f = threadIdx.x; f < p.Fp; f += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < p.C; ++c) acc = fmaf(p.W1g[k][(size_t)f * p.Cp + c], yh[c], acc);
    p.u[((size_t)k * p.N + n) * p.Fp + f] = acc;
  }
}

struct NoiseFillParams {
  float* noise;
  uint64_t seed;
  ChainIds ids;
  int K, D, S, N, C;
};
__global__ void fill_noise_kernel(const __grid_constant__ NoiseFillParams p) {
  const size_t total = (size_t)p.K * p.D * p.S * p.N * p.C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int c = (int)(r % p.C); r /= p.C;
    const int n = (int)(r % p.N); r /= p.N;
    const int s = (int)(r % p.S); r /= p.S;
    const int d = (int)(r % p.D); r /= p.D;
    const int k = (int)r;
    p.noise[i] = philox_normal(p.seed, p.ids.chain(k, d, n), (uint32_t)s, c);
  }
}

template <typename T>
cudaError_t dmalloc(T** p, size_t n, uint64_t* bytes) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
  if (e == cudaSuccess) *bytes += n * sizeof(T);
  return e;
}

void free_member_buffers(ladine_member* m) {
  for (int l = 0; l < 3; ++l) {
    cudaFree(m->A[l]);
    cudaFree(m->Cc[l]);
  }
  cudaFree(m->W1y);
  cudaFree(m->W1g);
  cudaFree(m->W4);
  cudaFree(m->b4);
  cudaFree(m->W2t);
  cudaFree(m->W3t);
  cudaFree(m->W2h);
  cudaFree(m->W3h);
  cudaFree(m->wscale);
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int ensure_workspace(ladine_handle* h, uint64_t bytes) {
  if (bytes <= h->ws_bytes) return LADINE_OK;
  if (h->ws) {
    // earlier stream-ordered work may still read the old buffer: wait for the handle's last call (calls on one handle
    // are chained through ev_done, so the latest event covers all of them) -- not for the whole device
    if (h->ev_done) cudaEventSynchronize(h->ev_done);
    else cudaDeviceSynchronize();
    cudaFree(h->ws);
    h->ws = nullptr;
    h->ws_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&h->ws, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    char buf[128];
    snprintf(buf, sizeof buf, "workspace allocation of %llu bytes failed", (unsigned long long)bytes);
    return fail(h, LADINE_ERR_NOMEM, buf);
  }
  h->ws_bytes = bytes;
  return LADINE_OK;
}

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct SlotInfo {
  int n_steps, n_slots, n_traj;
};
SlotInfo slot_info(const ladine_sample_args& a) {
  SlotInfo s;
  s.n_steps = a.t_first - a.t_last + 1;
  const int noisy = s.n_steps - (a.t_last == 0 ? 1 : 0);  // table index 0 draws no noise
  s.n_slots = (a.y_init ? 0 : 1) + noisy;
  s.n_traj = (a.y_init ? 0 : 1) + s.n_steps;
  return s;
}

int fill_ids(ladine_handle* h, const ladine_sample_args& a, int k0, int kn, ChainIds* ids) {
  for (int k = 0; k < kn; ++k) ids->member_gid[k] = a.member_ids ? a.member_ids[k0 + k] : k0 + k;
  ids->image_offset = a.image_offset;
  ids->images_total = a.images_total > 0 ? a.images_total : a.N;
  ids->draw_offset = a.draw_offset;
  ids->draws_total = a.draws_total > 0 ? a.draws_total : a.D;
  if (ids->image_offset < 0 || ids->draw_offset < 0 || ids->image_offset + a.N > ids->images_total ||
      ids->draw_offset + a.D > ids->draws_total)
    return fail(h, LADINE_ERR_INVALID, "image/draw offsets exceed images_total/draws_total");
  return LADINE_OK;
}

}  // namespace

namespace ladine {
cudaError_t order_after_previous_call(ladine_handle* h, cudaStream_t st) {
  if (!h->ev_done) return cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
  return cudaStreamWaitEvent(st, h->ev_done, 0);
}
void mark_call_done(ladine_handle* h, cudaStream_t st) {
  if (h->ev_done) cudaEventRecord(h->ev_done, st);
}
// ---- packed images (on-disk cache of packed members / encoders) ----
uint64_t image_checksum(const void* p, size_t n) {
  // four interleaved multiply-xor lanes over 64-bit words (memory-bound on the host), tail bytes folded in at the end
  const uint64_t kMul = 0x9E3779B97F4A7C15ull;
  uint64_t lane[4] = {0x243F6A8885A308D3ull, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
  const unsigned char* b = static_cast<const unsigned char*>(p);
  const size_t words = n / 8;
  size_t i = 0;
  for (; i + 4 <= words; i += 4) {
    uint64_t w[4];
    std::memcpy(w, b + 8 * i, 32);
    for (int l = 0; l < 4; ++l) lane[l] = (lane[l] ^ w[l]) * kMul;
  }
  uint64_t acc = lane[0];
  for (int l = 1; l < 4; ++l) acc = (acc ^ (lane[l] >> 29) ^ lane[l]) * kMul;
  for (size_t j = 8 * i; j < n; ++j) acc = (acc ^ b[j]) * kMul;
  return acc ^ (uint64_t)n;
}

uint64_t image_export(const char* magic, const int32_t* dims, int n_dims, const std::vector<ImageSection>& secs,
                      void* host_dst, uint64_t cap, cudaStream_t st, cudaError_t* status) {
  *status = cudaSuccess;
  uint64_t payload = 0;
  for (const auto& s : secs) payload += (s.bytes + 15) / 16 * 16;
  const uint64_t total = kImageHeaderBytes + payload;
  if (!host_dst || cap < total) return host_dst ? 0 : total;
  unsigned char* dst = static_cast<unsigned char*>(host_dst);
  std::memset(dst, 0, kImageHeaderBytes);
  uint64_t off = kImageHeaderBytes;
  for (const auto& s : secs) {
    const uint64_t padded = (s.bytes + 15) / 16 * 16;
    if (padded != s.bytes) std::memset(dst + off + s.bytes, 0, padded - s.bytes);
    cudaError_t e = cudaMemcpyAsync(dst + off, *s.dev, s.bytes, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { *status = e; return 0; }
    off += padded;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { *status = e; return 0; }
  ImageHeader hd{};
  std::memcpy(hd.magic, magic, std::min<size_t>(7, std::strlen(magic)));
  hd.abi = LADINE_ABI_VERSION;
  hd.layout = kImageLayout;
  for (int i = 0; i < n_dims && i < 20; ++i) hd.dims[i] = dims[i];
  hd.payload_bytes = payload;
  hd.checksum = image_checksum(dst + kImageHeaderBytes, payload);
  std::memcpy(dst, &hd, sizeof hd);
  return total;
}

const char* image_check(const void* host_src, uint64_t bytes, const char* magic, const ImageHeader** hdr_out) {
  if (!host_src || bytes < kImageHeaderBytes) return "image shorter than its header";
  const ImageHeader* hd = static_cast<const ImageHeader*>(host_src);
  if ((reinterpret_cast<uintptr_t>(host_src) & 7) != 0) return "image buffer must be 8-byte aligned";
  if (std::strncmp(hd->magic, magic, 8) != 0) return "not a packed image of this kind (magic mismatch)";
  if (hd->abi != LADINE_ABI_VERSION || hd->layout != kImageLayout)
    return "packed image was written by a different library version / packed layout: re-pack from the checkpoint";
  if (hd->payload_bytes != bytes - kImageHeaderBytes) return "packed image is truncated or has trailing bytes";
  if (image_checksum(static_cast<const unsigned char*>(host_src) + kImageHeaderBytes, hd->payload_bytes) != hd->checksum)
    return "packed image checksum mismatch (corrupt file)";
  *hdr_out = hd;
  return nullptr;
}

cudaError_t image_import(const void* host_src, const std::vector<ImageSection>& secs, cudaStream_t st) {
  const unsigned char* src = static_cast<const unsigned char*>(host_src);
  uint64_t off = kImageHeaderBytes;
  cudaError_t e = cudaSuccess;
  size_t done = 0;
  for (; done < secs.size() && e == cudaSuccess; ++done) {
    const auto& s = secs[done];
    *s.dev = nullptr;
    e = cudaMalloc(s.dev, s.bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(*s.dev, src + off, s.bytes, cudaMemcpyHostToDevice, st);
    off += (s.bytes + 15) / 16 * 16;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the host buffer is the caller's again on return
  if (e != cudaSuccess) {
    cudaGetLastError();
    for (size_t i = 0; i < done; ++i) {
      cudaFree(*secs[i].dev);
      *secs[i].dev = nullptr;
    }
  }
  return e;
}

cudaError_t launch_guidance_u(const ladine_member* const* members, int K, int N, const float* y0hat, float* u,
                              cudaStream_t st) {
  GuidanceParams p{};
  for (int k = 0; k < K; ++k) p.W1g[k] = members[k]->W1g;
  p.y0hat = y0hat;
  p.u = u;
  p.N = N;
  p.C = members[0]->C;
  p.Cp = members[0]->Cp;
  p.Fp = members[0]->Fp;
  guidance_u_kernel<<<dim3(N, K), 256, 0, st>>>(p);
  return cudaGetLastError();
}
}  // namespace ladine

// =============================================================================================
extern "C" {

int ladine_version(void) { return LADINE_ABI_VERSION; }

int ladine_create(int device, ladine_handle** out) {
  if (!out) return LADINE_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return LADINE_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LADINE_ERR_CUDA;
  if (prop.major != 10) return LADINE_ERR_UNSUPPORTED;  // sm_100a kernels only; no fallback
  ladine_handle* h = new (std::nothrow) ladine_handle();
  if (!h) return LADINE_ERR_NOMEM;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = h;
  return LADINE_OK;
}

int ladine_destroy(ladine_handle* h) {
  if (!h) return LADINE_ERR_INVALID;
  {
    DeviceGuard g(h->device);
    if (h->ws || h->enc_ws) {
      cudaDeviceSynchronize();
      cudaFree(h->ws);
      cudaFree(h->enc_ws);
    }
    for (auto& s : h->spans) {
      cudaEventDestroy(s.a);
      cudaEventDestroy(s.b);
    }
    for (auto e : h->pool) cudaEventDestroy(e);
    for (int l = 1; l < ladine_handle::kMaxLanes; ++l) {
      if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
      if (h->ev_join[l]) cudaEventDestroy(h->ev_join[l]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
  }
  delete h;
  return LADINE_OK;
}

const char* ladine_last_error(const ladine_handle* h) { return h ? h->err.c_str() : "null handle"; }
int64_t ladine_last_launches(const ladine_handle* h) { return h ? h->last_launches : 0; }
uint64_t ladine_workspace_bytes(const ladine_handle* h) { return h ? h->ws_bytes : 0; }
uint64_t ladine_member_bytes(const ladine_member* m) { return m ? m->bytes : 0; }
int ladine_member_precision(const ladine_member* m) { return m ? m->precision : LADINE_ERR_INVALID; }
int ladine_member_fpad(const ladine_member* m) { return m ? m->Fp : LADINE_ERR_INVALID; }
int ladine_member_cpad(const ladine_member* m) { return m ? m->Cp : LADINE_ERR_INVALID; }
int ladine_member_dims(const ladine_member* m, int32_t dims_out[6]) {
  if (!m || !dims_out) return LADINE_ERR_INVALID;
  const int32_t d[6] = {m->F, m->C, m->T, m->guidance, m->precision, m->device};
  std::memcpy(dims_out, d, sizeof d);
  return LADINE_OK;
}

int ladine_pack_member(ladine_handle* h, const ladine_member_desc* d, void* stream, ladine_member** out) {
  if (!h) return LADINE_ERR_INVALID;
  if (!d || !out) return fail(h, LADINE_ERR_INVALID, "null descriptor or output");
  *out = nullptr;
  if (d->struct_size != sizeof(ladine_member_desc)) return fail(h, LADINE_ERR_INVALID, "ladine_member_desc size mismatch");
  if (d->feature_dim < 1 || d->num_classes < 1 || d->n_steps < 1 || d->emb_rows < d->n_steps)
    return fail(h, LADINE_ERR_INVALID, "feature_dim, num_classes, n_steps must be >= 1 and emb_rows >= n_steps");
  if (d->num_classes > LADINE_MAX_CLASSES)
    return fail(h, LADINE_ERR_UNSUPPORTED, "num_classes > 16 is not supported by the fused kernels");
  const void* ptrs[] = {d->lin1_w, d->lin1_b, d->lin2_w, d->lin2_b, d->lin3_w, d->lin3_b, d->lin4_w, d->lin4_b,
                        d->emb[0], d->emb[1], d->emb[2], d->bn_w[0], d->bn_w[1], d->bn_w[2], d->bn_b[0], d->bn_b[1],
                        d->bn_b[2], d->bn_mean[0], d->bn_mean[1], d->bn_mean[2], d->bn_var[0], d->bn_var[1], d->bn_var[2]};
  for (const void* p : ptrs)
    if (!p) return fail(h, LADINE_ERR_INVALID, "null parameter pointer in ladine_member_desc");

  int prec = d->precision;
  if (prec == LADINE_PREC_AUTO) prec = d->feature_dim <= 128 ? LADINE_PREC_FP32 : LADINE_PREC_FP16;
  if (prec == LADINE_PREC_FP32 && d->feature_dim > 128)
    return fail(h, LADINE_ERR_UNSUPPORTED,
                "FP32 SMEM-resident path needs feature_dim <= 128 (two FP32 square layers must fit in shared memory); "
                "use LADINE_PREC_FP32X (split-operand tensor cores, FP32-grade) for wider members");
  if (prec != LADINE_PREC_FP32 && prec != LADINE_PREC_FP16 && prec != LADINE_PREC_BF16 && prec != LADINE_PREC_FP32X)
    return fail(h, LADINE_ERR_INVALID, "unknown precision");

  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ladine_member* m = new (std::nothrow) ladine_member();
  if (!m) return fail(h, LADINE_ERR_NOMEM, "host allocation failed");
  m->F = d->feature_dim;
  m->C = d->num_classes;
  m->Cp = cpad_of(m->C);
  m->T = d->n_steps;
  m->guidance = d->guidance ? 1 : 0;
  m->precision = prec;
  m->device = h->device;
  const bool tensor = prec != LADINE_PREC_FP32;
  const bool split = prec == LADINE_PREC_FP32X;
  m->split = split ? 1 : 0;
  m->Fp = tensor ? (int)align_up(m->F, 256) : (int)align_up(m->F, 32);
  const int F = m->F, Fp = m->Fp, T = m->T, Cp = m->Cp;

  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
  for (int l = 0; l < 3 && e == cudaSuccess; ++l) {
    ok(dmalloc(&m->A[l], (size_t)T * Fp, &m->bytes));
    ok(dmalloc(&m->Cc[l], (size_t)T * Fp, &m->bytes));
  }
  ok(dmalloc(&m->W1y, (size_t)Fp * Cp, &m->bytes));
  ok(dmalloc(&m->W1g, (size_t)Fp * Cp, &m->bytes));
  ok(dmalloc(&m->W4, (size_t)Cp * Fp, &m->bytes));
  ok(dmalloc(&m->b4, (size_t)Cp, &m->bytes));
  if (tensor) {
    ok(dmalloc(reinterpret_cast<uint16_t**>(&m->W2h), (size_t)Fp * Fp * (split ? 2 : 1), &m->bytes));
    ok(dmalloc(reinterpret_cast<uint16_t**>(&m->W3h), (size_t)Fp * Fp * (split ? 2 : 1), &m->bytes));
    if (split) ok(dmalloc(&m->wscale, 8, &m->bytes));   // [0..3] scales, [4..5] abs-max scratch (as uint bits)
  } else {
    ok(dmalloc(&m->W2t, (size_t)Fp * Fp, &m->bytes));
    ok(dmalloc(&m->W3t, (size_t)Fp * Fp, &m->bytes));
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    free_member_buffers(m);
    delete m;
    return fail(h, LADINE_ERR_NOMEM, "device allocation failed while packing a member");
  }

  const float gain = (tensor && !split) ? kLog2e : 1.0f;   // FP32X keeps natural units (exact-semantics softplus)
  const float* lin_b[3] = {d->lin1_b, d->lin2_b, d->lin3_b};
  const size_t tf = (size_t)T * Fp;
  if (split) {
    unsigned int* amax = reinterpret_cast<unsigned int*>(m->wscale + 4);
    cudaMemsetAsync(amax, 0, 2 * sizeof(unsigned int), st);
    absmax_kernel<<<296, 256, 0, st>>>(d->lin2_w, (size_t)F * F, amax);
    absmax_kernel<<<296, 256, 0, st>>>(d->lin3_w, (size_t)F * F, amax + 1);
    split_scale_kernel<<<1, 32, 0, st>>>(amax, m->wscale);
  }
  for (int l = 0; l < 3; ++l) {
    const float* a_gain = (split && l > 0) ? m->wscale + 2 + (l - 1) : nullptr;
    fold_tables_kernel<<<(unsigned)((tf + 255) / 256), 256, 0, st>>>(d->emb[l], lin_b[l], d->bn_w[l], d->bn_b[l],
                                                                     d->bn_mean[l], d->bn_var[l], d->bn_eps, gain, a_gain,
                                                                     T, F, Fp, m->A[l], m->Cc[l]);
  }
  pack_small_kernel<<<(Fp * Cp + 255) / 256, 256, 0, st>>>(d->lin1_w, d->lin4_w, d->lin4_b, F, Fp, m->C, Cp, m->guidance,
                                                           m->W1y, m->W1g, m->W4, m->b4);
  const size_t ff = (size_t)Fp * Fp;
  const unsigned gff = (unsigned)((ff + 255) / 256);
  if (!tensor) {
    transpose_pad_kernel<<<gff, 256, 0, st>>>(d->lin2_w, F, Fp, m->W2t);
    transpose_pad_kernel<<<gff, 256, 0, st>>>(d->lin3_w, F, Fp, m->W3t);
  } else if (split) {
    convert_split_kernel<<<gff, 256, 0, st>>>(d->lin2_w, F, Fp, m->wscale + 0, static_cast<__half*>(m->W2h));
    convert_split_kernel<<<gff, 256, 0, st>>>(d->lin3_w, F, Fp, m->wscale + 1, static_cast<__half*>(m->W3h));
  } else if (prec == LADINE_PREC_FP16) {
    convert_pad_kernel<__half><<<gff, 256, 0, st>>>(d->lin2_w, F, Fp, static_cast<__half*>(m->W2h));
    convert_pad_kernel<__half><<<gff, 256, 0, st>>>(d->lin3_w, F, Fp, static_cast<__half*>(m->W3h));
  } else {
    convert_pad_kernel<__nv_bfloat16><<<gff, 256, 0, st>>>(d->lin2_w, F, Fp, static_cast<__nv_bfloat16*>(m->W2h));
    convert_pad_kernel<__nv_bfloat16><<<gff, 256, 0, st>>>(d->lin3_w, F, Fp, static_cast<__nv_bfloat16*>(m->W3h));
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    free_member_buffers(m);
    delete m;
    return fail_cuda(h, e, "member packing kernels");
  }
  *out = m;
  return LADINE_OK;
}

int ladine_free_member(ladine_handle* h, ladine_member* m) {
  if (!m) return LADINE_ERR_INVALID;
  DeviceGuard guard(m->device);
  cudaDeviceSynchronize();
  free_member_buffers(m);
  delete m;
  (void)h;
  return LADINE_OK;
}

// ---- packed-member images (checkpoint ingestion with a cache that skips re-packing) ----
static std::vector<ImageSection> member_sections(ladine_member* m) {
  const size_t tf = (size_t)m->T * m->Fp * sizeof(float), fc = (size_t)m->Fp * m->Cp * sizeof(float);
  const size_t ff = (size_t)m->Fp * m->Fp;
  std::vector<ImageSection> s;
  for (int l = 0; l < 3; ++l) {
    s.push_back({reinterpret_cast<void**>(&m->A[l]), tf});
    s.push_back({reinterpret_cast<void**>(&m->Cc[l]), tf});
  }
  s.push_back({reinterpret_cast<void**>(&m->W1y), fc});
  s.push_back({reinterpret_cast<void**>(&m->W1g), fc});
  s.push_back({reinterpret_cast<void**>(&m->W4), fc});
  s.push_back({reinterpret_cast<void**>(&m->b4), (size_t)m->Cp * sizeof(float)});
  if (m->precision != LADINE_PREC_FP32) {
    s.push_back({&m->W2h, ff * 2 * (m->split ? 2 : 1)});
    s.push_back({&m->W3h, ff * 2 * (m->split ? 2 : 1)});
    if (m->split) s.push_back({reinterpret_cast<void**>(&m->wscale), 8 * sizeof(float)});
  } else {
    s.push_back({reinterpret_cast<void**>(&m->W2t), ff * sizeof(float)});
    s.push_back({reinterpret_cast<void**>(&m->W3t), ff * sizeof(float)});
  }
  return s;
}

// Host-only inspection of a packed image (needs no device or handle): what kind it is and whether it is intact.
int ladine_image_info(const void* host_src, uint64_t bytes, int32_t* kind_out, int32_t dims_out[20], const char** why_out) {
  static const char* const kOk = "";
  if (why_out) *why_out = kOk;
  if (kind_out) *kind_out = 0;
  const ImageHeader* hd = nullptr;
  const char* why = image_check(host_src, bytes, "LADINEM", &hd);
  int kind = 1;
  if (why && host_src && bytes >= kImageHeaderBytes && std::strncmp(static_cast<const char*>(host_src), "LADINEE", 8) == 0) {
    why = image_check(host_src, bytes, "LADINEE", &hd);
    kind = 2;
  }
  if (why) {
    if (why_out) *why_out = why;
    return LADINE_ERR_INVALID;
  }
  if (kind_out) *kind_out = kind;
  if (dims_out) std::memcpy(dims_out, hd->dims, sizeof hd->dims);
  return LADINE_OK;
}
uint64_t ladine_image_checksum(const void* host_src, uint64_t bytes) { return host_src ? image_checksum(host_src, bytes) : 0; }

uint64_t ladine_member_image_bytes(const ladine_member* m) {
  if (!m) return 0;
  uint64_t total = kImageHeaderBytes;
  for (const auto& s : member_sections(const_cast<ladine_member*>(m))) total += (s.bytes + 15) / 16 * 16;
  return total;
}

int ladine_member_export(ladine_handle* h, const ladine_member* m, void* host_dst, uint64_t capacity, void* stream) {
  if (!h) return LADINE_ERR_INVALID;
  if (!m || !host_dst) return fail(h, LADINE_ERR_INVALID, "null member or destination");
  if (capacity < ladine_member_image_bytes(m)) return fail(h, LADINE_ERR_INVALID, "destination smaller than ladine_member_image_bytes");
  DeviceGuard guard(m->device);
  const int32_t dims[8] = {m->F, m->Fp, m->C, m->Cp, m->T, m->guidance, m->precision, m->split};
  cudaError_t e = cudaSuccess;
  const uint64_t n = image_export("LADINEM", dims, 8, member_sections(const_cast<ladine_member*>(m)), host_dst, capacity,
                                  static_cast<cudaStream_t>(stream), &e);
  if (e != cudaSuccess || n == 0) return fail_cuda(h, e, "member export");
  return LADINE_OK;
}

int ladine_member_import(ladine_handle* h, const void* host_src, uint64_t bytes, void* stream, ladine_member** out) {
  if (!h) return LADINE_ERR_INVALID;
  if (!out) return fail(h, LADINE_ERR_INVALID, "null output");
  *out = nullptr;
  const ImageHeader* hd = nullptr;
  if (const char* why = image_check(host_src, bytes, "LADINEM", &hd)) return fail(h, LADINE_ERR_INVALID, why);
  ladine_member* m = new (std::nothrow) ladine_member();
  if (!m) return fail(h, LADINE_ERR_NOMEM, "host allocation failed");
  m->F = hd->dims[0]; m->Fp = hd->dims[1]; m->C = hd->dims[2]; m->Cp = hd->dims[3]; m->T = hd->dims[4];
  m->guidance = hd->dims[5]; m->precision = hd->dims[6]; m->split = hd->dims[7];
  m->device = h->device;
  const bool tensor = m->precision != LADINE_PREC_FP32;
  const bool sane = m->F >= 1 && m->C >= 1 && m->C <= LADINE_MAX_CLASSES && m->T >= 1 && m->Cp == cpad_of(m->C) &&
                    (m->precision == LADINE_PREC_FP32 || m->precision == LADINE_PREC_FP16 ||
                     m->precision == LADINE_PREC_BF16 || m->precision == LADINE_PREC_FP32X) &&
                    m->Fp == (int)align_up(m->F, tensor ? 256 : 32) && m->split == (m->precision == LADINE_PREC_FP32X ? 1 : 0) &&
                    (m->precision != LADINE_PREC_FP32 || m->F <= 128);
  auto secs = member_sections(m);
  uint64_t payload = 0;
  for (const auto& s : secs) payload += (s.bytes + 15) / 16 * 16;
  if (!sane || payload != hd->payload_bytes) {
    delete m;
    return fail(h, LADINE_ERR_INVALID, "packed member image: inconsistent dimensions");
  }
  DeviceGuard guard(h->device);
  cudaError_t e = image_import(host_src, secs, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    delete m;
    return e == cudaErrorMemoryAllocation ? fail(h, LADINE_ERR_NOMEM, "device allocation failed while importing a member")
                                          : fail_cuda(h, e, "member import");
  }
  for (const auto& s : secs) m->bytes += s.bytes;
  *out = m;
  return LADINE_OK;
}

static int validate_sample(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args* a) {
  if (!members || !a) return fail(h, LADINE_ERR_INVALID, "null members or args");
  if (a->struct_size != sizeof(ladine_sample_args)) return fail(h, LADINE_ERR_INVALID, "ladine_sample_args size mismatch");
  if (a->K < 1 || a->N < 1 || a->D < 1 || a->T < 1) return fail(h, LADINE_ERR_INVALID, "K, N, D, T must be >= 1");
  if (a->t_first < a->t_last || a->t_last < 0 || a->t_first >= a->T)
    return fail(h, LADINE_ERR_INVALID, "need 0 <= t_last <= t_first < T");
  if (!a->xf || !a->y0hat || !a->ytmean || !a->coef || !a->y_out)
    return fail(h, LADINE_ERR_INVALID, "xf, y0hat, ytmean, coef and y_out are required");
  if (a->prob_out && !(a->temperature > 0.f)) return fail(h, LADINE_ERR_INVALID, "temperature must be > 0");
  const ladine_member* m0 = members[0];
  for (int k = 0; k < a->K; ++k) {
    const ladine_member* m = members[k];
    if (!m) return fail(h, LADINE_ERR_INVALID, "null member");
    if (m->device != h->device) return fail(h, LADINE_ERR_INVALID, "member packed on another device");
    if (m->F != m0->F || m->C != m0->C || m->precision != m0->precision || m->guidance != m0->guidance)
      return fail(h, LADINE_ERR_INVALID, "members of one call must share feature_dim, num_classes, guidance and precision");
    if (m->T < a->T) return fail(h, LADINE_ERR_INVALID, "member packed with fewer table rows than args.T");
  }
  return LADINE_OK;
}

// member groups: sizes differ by at most one, each <= LADINE_MAX_GROUP, at least `lanes` of them when K allows
static std::vector<std::pair<int, int>> plan_groups(int K, int lanes) {
  int n = (K + LADINE_MAX_GROUP - 1) / LADINE_MAX_GROUP;
  const int want = lanes < K ? lanes : K;
  if (n < want) n = want;
  std::vector<std::pair<int, int>> g;
  int k0 = 0;
  for (int i = 0; i < n; ++i) {
    const int kn = K / n + (i < K % n ? 1 : 0);
    g.push_back({k0, kn});
    k0 += kn;
  }
  return g;
}

static int ensure_lane_resources(ladine_handle* h, int lanes) {
  cudaError_t e = cudaSuccess;
  if (!h->ev_fork) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  for (int l = 1; l < lanes && e == cudaSuccess; ++l) {
    if (!h->lane_stream[l]) e = cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking);
    if (e == cudaSuccess && !h->ev_join[l]) e = cudaEventCreateWithFlags(&h->ev_join[l], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) return fail_cuda(h, e, "lane stream/event creation");
  return LADINE_OK;
}

int ladine_sample(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args* a) {
  if (!h) return LADINE_ERR_INVALID;
  int rc = validate_sample(h, members, a);
  if (rc != LADINE_OK) return rc;
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const ladine_member* m0 = members[0];
  const int F = m0->F, Fp = m0->Fp, C = m0->C, Cp = m0->Cp;
  const SlotInfo si = slot_info(*a);
  const bool tensor = m0->precision != LADINE_PREC_FP32;
  const uint64_t act_cols = (uint64_t)Fp * (m0->split ? 2 : 1);   // FP32X: hi | lo halves per activation row
  h->last_launches = 0;

  const StepCoef* h_coef = reinterpret_cast<const StepCoef*>(a->coef);
  static_assert(sizeof(StepCoef) == 8 * sizeof(float), "coef rows are 8 floats");
  // the handle's workspace is shared by successive calls: order this call after the previous one even when the
  // caller switched streams (a no-op when both are on the same stream)
  if (ladine::order_after_previous_call(h, st) != cudaSuccess) return fail(h, LADINE_ERR_CUDA, "stream ordering event");

  // Small calls, opt-in: the whole chain in ONE cooperative launch (split-K over all SMs, grid barriers between phases)
  const int pS = (tensor && !m0->split && h->persist) ? persist_splits(h, a->K, a->N * a->D, Fp, C) : 0;
  if (pS > 0) {
    uint64_t off = 0;
    const uint64_t o_coef = off; off = align_up(off + (uint64_t)a->T * sizeof(StepCoef), 1024);
    const uint64_t o_u = off; off = align_up(off + (uint64_t)a->K * a->N * Fp * 4, 1024);
    const uint64_t o_ws = off; off += persist_workspace_bytes(a->K, Fp, Cp, pS);
    rc = ensure_workspace(h, off);
    if (rc != LADINE_OK) return rc;
    uint8_t* ws = static_cast<uint8_t*>(h->ws);
    cudaError_t e = cudaMemcpyAsync(ws + o_coef, a->coef, (size_t)a->T * sizeof(StepCoef), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail_cuda(h, e, "coefficient upload");
    ChainIds ids{};
    rc = fill_ids(h, *a, 0, a->K, &ids);
    if (rc != LADINE_OK) return rc;
    float* d_u = reinterpret_cast<float*>(ws + o_u);
    e = launch_guidance_u(members, a->K, a->N, a->y0hat, d_u, st);
    if (e != cudaSuccess) return fail_cuda(h, e, "guidance projection kernel");
    h->last_launches += 1;
    std::string err;
    e = launch_persistent_chain(h, members, *a, ids, reinterpret_cast<const StepCoef*>(ws + o_coef), d_u, ws + o_ws, pS,
                                si.n_slots, si.n_traj, st, &h->last_launches, &err);
    if (e == cudaSuccess) {
      ladine::mark_call_done(h, st);
      return LADINE_OK;
    }
    if (e != cudaErrorCooperativeLaunchTooLarge)
      return err.empty() ? fail_cuda(h, e, "persistent chain launch") : fail(h, LADINE_ERR_CUDA, err);
    // not every CTA of the cooperative grid can be resident right now (the occupancy check runs before anything of the
    // chain is enqueued): use the tile kernels for this call
    cudaGetLastError();
    h->last_launches = 0;
  }

  // Tensor path: member groups advance concurrently on up to `lanes` streams (lane 0 = caller's stream).
  // The resident path is a single launch per group and needs no lanes.
  int lanes = tensor ? h->lanes : 1;
  if (lanes < 1) lanes = 1;
  if (lanes > ladine_handle::kMaxLanes) lanes = ladine_handle::kMaxLanes;
  const std::vector<std::pair<int, int>> groups = plan_groups(a->K, lanes);
  if ((int)groups.size() < lanes) lanes = (int)groups.size();
  int gmax = 0;
  for (auto& g : groups) gmax = g.second > gmax ? g.second : gmax;

  // workspace: a coefficient table + one slice per lane, each sized for the largest group
  const int rows = a->N * a->D;
  const int rows_pad = (rows + 255) / 256 * 256;  // covers both tile geometries (128- and 256-row tiles)
  const uint64_t m_total = (uint64_t)gmax * rows_pad;
  uint64_t off = 0;
  const uint64_t o_coef = off; off = align_up(off + (uint64_t)a->T * sizeof(StepCoef), 1024);
  const uint64_t lane_base = off;
  uint64_t lo = 0;
  const uint64_t o_u = lo; lo = align_up(lo + (uint64_t)gmax * a->N * Fp * 4, 1024);
  uint64_t o_h1 = 0, o_h2 = 0, o_part = 0, o_y = 0, o_sched = 0, o_arr = 0;
  if (tensor) {
    o_h1 = lo; lo = align_up(lo + m_total * act_cols * 2, 1024);
    o_h2 = lo; lo = align_up(lo + m_total * act_cols * 2, 1024);
    o_part = lo; lo = align_up(lo + m_total * (Fp / 256) * 2 * Cp * 4, 1024);
    o_sched = lo; lo = align_up(lo + sched_bytes_bound(gmax, rows, Fp / 128), 1024);
    o_arr = lo; lo = align_up(lo + (uint64_t)gmax * ((rows + 127) / 128 + 1) * 2 * sizeof(int), 1024);
    o_y = lo; lo = align_up(lo + 2 * m_total * Cp * 4, 1024);
  }
  const uint64_t lane_bytes = lo;
  rc = ensure_workspace(h, lane_base + lane_bytes * lanes);
  if (rc != LADINE_OK) return rc;
  uint8_t* ws = static_cast<uint8_t*>(h->ws);
  cudaError_t e;

  if (!tensor) {
    const size_t smem = resident_smem_bytes(Fp, Cp);
    if ((int)smem > h->max_smem_optin) return fail(h, LADINE_ERR_UNSUPPORTED, "shared memory budget exceeded");
    e = cudaMemcpyAsync(ws + o_coef, a->coef, (size_t)a->T * sizeof(StepCoef), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail_cuda(h, e, "coefficient upload");
  }
  if (lanes > 1) {
    rc = ensure_lane_resources(h, lanes);
    if (rc != LADINE_OK) return rc;
    e = cudaEventRecord(h->ev_fork, st);  // lanes start after everything already queued on the caller's stream
    for (int l = 1; l < lanes && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(h->lane_stream[l], h->ev_fork, 0);
    if (e != cudaSuccess) return fail_cuda(h, e, "lane fork");
  }

  auto group_args = [&](int k0, int kn) {
    ladine_sample_args g = *a;
    g.K = kn;
    g.xf = a->xf + (size_t)k0 * a->N * F;
    g.y0hat = a->y0hat + (size_t)k0 * a->N * C;
    g.ytmean = a->ytmean + (size_t)k0 * a->N * C;
    const size_t kdnc = (size_t)a->D * a->N * C;
    if (a->y_init) g.y_init = a->y_init + k0 * kdnc;
    if (a->noise) g.noise = a->noise + k0 * kdnc * si.n_slots;
    g.y_out = a->y_out + k0 * kdnc;
    if (a->traj_out) g.traj_out = a->traj_out + k0 * kdnc * si.n_traj;
    if (a->prob_out) g.prob_out = a->prob_out + k0 * kdnc;
    return g;
  };

  int status = LADINE_OK;
  // rounds of up to `lanes` groups; inside a round the lanes' steps are enqueued interleaved so every
  // stream always has work queued (the launch queue is finite: enqueueing lane by lane would serialise them)
  for (size_t g0 = 0; g0 < groups.size() && status == LADINE_OK; g0 += lanes) {
    const int nl = (int)((groups.size() - g0) < (size_t)lanes ? (groups.size() - g0) : lanes);
    TensorChain* chains[ladine_handle::kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    for (int l = 0; l < nl && status == LADINE_OK; ++l) {
      const int k0 = groups[g0 + l].first, kn = groups[g0 + l].second;
      cudaStream_t ls = l == 0 ? st : h->lane_stream[l];
      uint8_t* lw = ws + lane_base + lane_bytes * l;
      const ladine_sample_args g = group_args(k0, kn);
      ChainIds ids{};
      status = fill_ids(h, *a, k0, kn, &ids);
      if (status != LADINE_OK) break;
      float* d_u = reinterpret_cast<float*>(lw + o_u);
      e = launch_guidance_u(members + k0, kn, a->N, g.y0hat, d_u, ls);
      if (e != cudaSuccess) { status = fail_cuda(h, e, "guidance projection kernel"); break; }
      h->last_launches += 1;
      if (!tensor) {
        e = launch_resident(h, members + k0, g, ids, reinterpret_cast<const StepCoef*>(ws + o_coef), d_u, si.n_slots,
                            si.n_traj, ls, &h->last_launches);
        if (e != cudaSuccess) status = fail_cuda(h, e, "resident sampler launch");
      } else {
        TensorWorkspace tw;
        tw.h1 = lw + o_h1;
        tw.h2 = lw + o_h2;
        tw.part = reinterpret_cast<float*>(lw + o_part);
        tw.ybuf = reinterpret_cast<float*>(lw + o_y);
        tw.u = d_u;
        tw.sched = reinterpret_cast<int32_t*>(lw + o_sched);
        tw.arrivals = reinterpret_cast<int*>(lw + o_arr);
        std::string err;
        chains[l] = tensor_chain_create(h, members + k0, g, ids, h_coef, tw, si.n_slots, si.n_traj, ls,
                                        /*single_lane=*/nl == 1, &h->last_launches, &err, &e);
        if (!chains[l]) status = err.empty() ? fail_cuda(h, e, "tensor-core sampler launch") : fail(h, LADINE_ERR_CUDA, err);
      }
    }
    if (tensor) {
      for (int t = a->t_first; t >= a->t_last && status == LADINE_OK; --t) {
        for (int l = 0; l < nl; ++l) {
          e = tensor_chain_step(chains[l], t, &h->last_launches);
          if (e != cudaSuccess) { status = fail_cuda(h, e, "tensor-core sampler step launch"); break; }
        }
      }
      for (int l = 0; l < nl; ++l)
        if (chains[l]) tensor_chain_destroy(chains[l]);
    }
  }
  // join: the caller's stream continues only after every lane has finished (also on the error path)
  for (int l = 1; l < lanes; ++l) {
    cudaError_t je = cudaEventRecord(h->ev_join[l], h->lane_stream[l]);
    if (je == cudaSuccess) je = cudaStreamWaitEvent(st, h->ev_join[l], 0);
    if (je != cudaSuccess && status == LADINE_OK) status = fail_cuda(h, je, "lane join");
  }
  ladine::mark_call_done(h, st);
  return status;
}

int ladine_set_option(ladine_handle* h, const char* key, int64_t value) {
  if (!h || !key) return LADINE_ERR_INVALID;
  if (strcmp(key, "lanes") == 0) {
    if (value < 1 || value > ladine_handle::kMaxLanes) return fail(h, LADINE_ERR_INVALID, "lanes must be in [1, 4]");
    h->lanes = (int)value;
    return LADINE_OK;
  }
  if (strcmp(key, "fuse") == 0) {
    h->fuse = value != 0;
    return LADINE_OK;
  }
  if (strcmp(key, "persist") == 0) {
    h->persist = value != 0;
    return LADINE_OK;
  }
  if (strcmp(key, "persist_debug") == 0) {
    h->persist_debug = value != 0;
    return LADINE_OK;
  }
  if (strcmp(key, "ctas") == 0) {
    if (value < 0 || value > 3) return fail(h, LADINE_ERR_INVALID, "ctas must be 0 (auto), 1, 2 or 3 (slim tiles)");
    h->ctas = (int)value;
    return LADINE_OK;
  }
  if (strcmp(key, "order") == 0) {
    if (value < 0 || value > 2) return fail(h, LADINE_ERR_INVALID, "order must be 0 (auto), 1 (N-tile-major) or 2 (row-major)");
    h->order = (int)value;
    return LADINE_OK;
  }
  if (strcmp(key, "tail_vec") == 0) {
    if (value != 0 && value != 4 && value != 8) return fail(h, LADINE_ERR_INVALID, "tail_vec must be 0 (auto), 4 or 8");
    h->tail_vec = (int)value;
    return LADINE_OK;
  }
  if (strcmp(key, "pair_gain_permille") == 0) {
    if (value < 500 || value > 2000) return fail(h, LADINE_ERR_INVALID, "pair_gain_permille must be in [500, 2000]");
    h->pair_gain = value / 1000.0;
    return LADINE_OK;
  }
  return fail(h, LADINE_ERR_INVALID, std::string("unknown option ") + key);
}

int ladine_fill_noise(ladine_handle* h, const ladine_sample_args* a, int32_t num_classes, float* noise) {
  if (!h) return LADINE_ERR_INVALID;
  if (!a || !noise || num_classes < 1) return fail(h, LADINE_ERR_INVALID, "null args or output");
  if (a->K < 1 || a->K > LADINE_MAX_GROUP * 64 || a->N < 1 || a->D < 1 || a->t_first < a->t_last || a->t_last < 0)
    return fail(h, LADINE_ERR_INVALID, "bad K/N/D/t range");
  DeviceGuard guard(h->device);
  const SlotInfo si = slot_info(*a);
  const size_t kdsnc = (size_t)a->D * si.n_slots * a->N * num_classes;
  for (int k0 = 0; k0 < a->K; k0 += LADINE_MAX_GROUP) {
    const int kn = (a->K - k0) < LADINE_MAX_GROUP ? (a->K - k0) : LADINE_MAX_GROUP;
    NoiseFillParams p{};
    int rc = fill_ids(h, *a, k0, kn, &p.ids);
    if (rc != LADINE_OK) return rc;
    p.noise = noise + k0 * kdsnc;
    p.seed = a->seed;
    p.K = kn;
    p.D = a->D;
    p.S = si.n_slots;
    p.N = a->N;
    p.C = num_classes;
    const size_t total = kn * kdsnc;
    if (total == 0) continue;
    const unsigned grid = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    fill_noise_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(a->stream)>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(h, e, "noise fill kernel");
  return LADINE_OK;
}

int ladine_set_profiling(ladine_handle* h, int enabled) {
  if (!h) return LADINE_ERR_INVALID;
  h->profiling = enabled != 0;
  return LADINE_OK;
}

int ladine_get_profile(ladine_handle* h, float ms_out[4], int64_t count_out[3]) {
  if (!h || !ms_out || !count_out) return LADINE_ERR_INVALID;
  DeviceGuard guard(h->device);
  for (int i = 0; i < 4; ++i) ms_out[i] = 0.f;
  for (int i = 0; i < 3; ++i) count_out[i] = 0;
  std::vector<std::pair<float, float>> gemm;  // [start, end) offsets of GEMM spans from the first recorded event
  cudaEvent_t base = h->spans.empty() ? nullptr : h->spans.front().a;
  for (auto& s : h->spans) {
    cudaError_t e = cudaEventSynchronize(s.b);
    float ms = 0.f, t0 = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, s.a, s.b);
    if (e == cudaSuccess && s.kind < 2) e = cudaEventElapsedTime(&t0, base, s.a);
    if (e != cudaSuccess) return fail_cuda(h, e, "profile event readback");
    ms_out[s.kind] += ms;
    count_out[s.kind] += 1;
    if (s.kind < 2) gemm.push_back({t0, t0 + ms});
  }
  // union of the GEMM busy intervals (lanes overlap; a lane's span also counts time queued behind another lane)
  std::sort(gemm.begin(), gemm.end());
  float cur_a = 0.f, cur_b = -1.f;
  for (auto& iv : gemm) {
    if (cur_b < cur_a || iv.first > cur_b) {
      if (cur_b >= cur_a) ms_out[3] += cur_b - cur_a;
      cur_a = iv.first;
      cur_b = iv.second;
    } else if (iv.second > cur_b) {
      cur_b = iv.second;
    }
  }
  if (cur_b >= cur_a) ms_out[3] += cur_b - cur_a;
  for (auto& s : h->spans) {
    h->pool.push_back(s.a);
    h->pool.push_back(s.b);
  }
  h->spans.clear();
  return LADINE_OK;
}

int32_t ladine_debug_geometry(int32_t K, int32_t rows, int32_t feature_dim_padded, int32_t sm_count) {
  if (K < 1 || rows < 1 || feature_dim_padded < 256 || feature_dim_padded % 256 != 0 || sm_count < 2) return LADINE_ERR_INVALID;
  return ladine::debug_geometry(K, rows, feature_dim_padded, sm_count);
}

int64_t ladine_debug_plan(int32_t K, int32_t rows, int32_t feature_dim_padded, int32_t geometry, int32_t row_major,
                          int32_t units, int32_t* table_out, int64_t cap, int32_t info_out[4]) {
  if (K < 1 || K > LADINE_MAX_GROUP || rows < 1 || feature_dim_padded < 256 || feature_dim_padded % 256 != 0 ||
      geometry < 1 || geometry > 3 || units < 1 || !info_out || (rows + 127) / 128 > 4096)
    return LADINE_ERR_INVALID;
  return ladine::debug_plan(K, rows, feature_dim_padded, geometry, row_major, units, table_out, cap, info_out);
}

int ladine_debug_layer(ladine_handle* h, const ladine_member* member, int layer, int t, const void* h_in, int rows,
                       void* h_out, float* part, void* stream) {
  if (!h) return LADINE_ERR_INVALID;
  if (!member || !h_in || rows < 1 || (layer != 2 && layer != 3) || t < 0 || t >= member->T)
    return fail(h, LADINE_ERR_INVALID, "bad debug-layer arguments");
  if (member->precision == LADINE_PREC_FP32) return fail(h, LADINE_ERR_UNSUPPORTED, "debug layer is for the tensor path");
  if ((layer == 2 && !h_out) || (layer == 3 && !part)) return fail(h, LADINE_ERR_INVALID, "missing output buffer");
  DeviceGuard guard(h->device);
  int rc = ensure_workspace(h, sched_bytes_bound(1, rows, member->Fp / 128));
  if (rc != LADINE_OK) return rc;
  std::string err;
  cudaError_t e = launch_debug_layer(h, member, layer, t, h_in, rows, h_out, part, static_cast<int32_t*>(h->ws),
                                     static_cast<cudaStream_t>(stream), &err);
  if (e != cudaSuccess) {
    if (!err.empty()) return fail(h, LADINE_ERR_CUDA, err);
    return fail_cuda(h, e, "debug layer launch");
  }
  if (!h->ev_done) cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
  ladine::mark_call_done(h, static_cast<cudaStream_t>(stream));
  return LADINE_OK;
}

}  // extern "C"
