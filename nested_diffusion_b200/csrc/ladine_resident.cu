// FP32 SMEM-resident sampler: one CTA owns (member, 32-row tile) and runs every reverse step on-chip.
//
// This is the literal BASELINE.json north-star design and is valid while the member's two square
// layers fit in shared memory in FP32 (feature_dim <= 128: 2 * 128 * 128 * 4 B = 128 KiB).
// Replaces, for those shapes, the whole diffusion_utils.p_sample_loop chain (diffusion_utils.py:133-163)
// and the trunk of latent_model.ConditionalModel.forward (latent_model.py:172-184) in ONE launch:
//   - W2^T / W3^T staged once into SMEM (k-major so a warp reads one 16-byte broadcast per 4 columns),
//   - xf, u = W1g . y_0_hat and the activations kept in SMEM as [feature][row] (conflict-free),
//   - y_t and the posterior update in registers/SMEM, noise injected or Philox4x32-10,
//   - per-step A_l[t]/C_l[t] rows read straight from L2 (6 * F floats per step).
// Thread (warp w, lane l) owns row l of the tile and the NCOL = F/8 columns [w*NCOL, (w+1)*NCOL).
#include "ladine_internal.cuh"

namespace ladine {
namespace {

constexpr int kRows = 32;      // rows per CTA (= warp width: lane <-> row)
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

struct ResidentParams {
  // per member (group of up to LADINE_MAX_GROUP)
  const float* A[LADINE_MAX_GROUP][3];
  const float* Cc[LADINE_MAX_GROUP][3];
  const float* W1y[LADINE_MAX_GROUP];
  const float* W4[LADINE_MAX_GROUP];
  const float* b4[LADINE_MAX_GROUP];
  const float* W2t[LADINE_MAX_GROUP];
  const float* W3t[LADINE_MAX_GROUP];
  const float* xf;      // [K, N, F]   (unpadded F stride = Fin)
  const float* u;       // [K, N, Fp]
  const float* ytmean;  // [K, N, C]
  const float* y_init;  // [K, D, N, C] or null
  const float* noise;   // [K, D, S, N, C] or null
  const StepCoef* coef; // [T]
  float* y_out;
  float* traj_out;
  float* prob_out;
  float temperature;
  uint64_t seed;
  ChainIds ids;
  int N, D, C, Cp, Fin, Fp, rows, t_first, t_last, n_slots, n_traj;
};

template <int NCOL>
__device__ __forceinline__ void dense_layer(const float* __restrict__ hin, const float* __restrict__ Wt, int Fp,
                                            int col0, int lane, float (&acc)[NCOL]) {
#pragma unroll
  for (int j = 0; j < NCOL; ++j) acc[j] = 0.f;
  const float* wrow = Wt + col0;
#pragma unroll 4
  for (int k = 0; k < Fp; ++k) {
    const float hv = hin[k * kRows + lane];
    const float4* wp = reinterpret_cast<const float4*>(wrow + (size_t)k * Fp);
#pragma unroll
    for (int j4 = 0; j4 < NCOL / 4; ++j4) {
      const float4 wv = wp[j4];
      acc[4 * j4 + 0] = fmaf(hv, wv.x, acc[4 * j4 + 0]);
      acc[4 * j4 + 1] = fmaf(hv, wv.y, acc[4 * j4 + 1]);
      acc[4 * j4 + 2] = fmaf(hv, wv.z, acc[4 * j4 + 2]);
      acc[4 * j4 + 3] = fmaf(hv, wv.w, acc[4 * j4 + 3]);
    }
  }
}

template <int NCOL>
__global__ void __launch_bounds__(kThreads, 1) resident_chain_kernel(const __grid_constant__ ResidentParams p) {
  extern __shared__ __align__(16) float smem[];
  const int Fp = p.Fp, C = p.C, Cp = p.Cp;
  float* sW2 = smem;
  float* sW3 = sW2 + Fp * Fp;
  float* sX = sW3 + Fp * Fp;          // xf   [Fp][32]
  float* sU = sX + Fp * kRows;        // u    [Fp][32]
  float* sHa = sU + Fp * kRows;       // h1   [Fp][32]
  float* sHb = sHa + Fp * kRows;      // h2   [Fp][32]
  float* sRed = sHb + Fp * kRows;     // lin4 partials [8][32][Cp]
  float* sY = sRed + kWarps * kRows * Cp;   // y_t  [32][Cp]
  float* sMu = sY + kRows * Cp;             // y_T_mean [32][Cp]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.y;
  const int row0 = blockIdx.x * kRows;
  const int col0 = warp * NCOL;

  // ---- one-time staging -------------------------------------------------------------------
  {
    const float4* g2 = reinterpret_cast<const float4*>(p.W2t[k]);
    const float4* g3 = reinterpret_cast<const float4*>(p.W3t[k]);
    float4* s2 = reinterpret_cast<float4*>(sW2);
    float4* s3 = reinterpret_cast<float4*>(sW3);
    for (int i = tid; i < Fp * Fp / 4; i += kThreads) {
      s2[i] = __ldg(g2 + i);
      s3[i] = __ldg(g3 + i);
    }
    for (int i = tid; i < Fp * kRows; i += kThreads) {
      const int f = i % Fp, r = i / Fp;  // coalesced over f
      const int rg = row0 + r;
      float xv = 0.f, uv = 0.f;
      if (rg < p.rows) {
        const int n = rg / p.D;
        if (f < p.Fin) xv = __ldg(p.xf + ((size_t)k * p.N + n) * p.Fin + f);
        uv = __ldg(p.u + ((size_t)k * p.N + n) * Fp + f);
      }
      sX[f * kRows + r] = xv;
      sU[f * kRows + r] = uv;
    }
  }
  // (row, class) items owned by this thread in the update phase: i = tid + q * 256 < 32 * C
  const int rg_lane = row0 + lane;
  for (int i = tid; i < kRows * C; i += kThreads) {
    const int r = i / C, c = i % C;
    const int rg = row0 + r;
    float mu = 0.f, y = 0.f;
    if (rg < p.rows) {
      const int n = rg / p.D, d = rg % p.D;
      mu = __ldg(p.ytmean + ((size_t)k * p.N + n) * C + c);
      if (p.y_init) {
        y = __ldg(p.y_init + (((size_t)k * p.D + d) * p.N + n) * C + c);
      } else {
        // y_T = z + y_T_mean (diffusion_utils.py:139-140), noise slot 0
        const float z = p.noise ? __ldg(p.noise + ((((size_t)k * p.D + d) * p.n_slots + 0) * p.N + n) * C + c)
                                : philox_normal(p.seed, p.ids.chain(k, d, n), 0u, c);
        y = __fadd_rn(z, mu);
        if (p.traj_out) p.traj_out[((((size_t)k * p.D + d) * p.n_traj + 0) * p.N + n) * C + c] = y;
      }
    }
    sMu[r * Cp + c] = mu;
    sY[r * Cp + c] = y;
  }
  __syncthreads();

  const int slot_base = p.y_init ? 0 : 1;
  float acc[NCOL];

  for (int t = p.t_first; t >= p.t_last; --t) {
    const size_t trow = (size_t)t * Fp;
    // ---- lin1 (+gamma, BN folded) -> softplus -> gate by xf  (latent_model.py:174-177) ----
    {
      float yv[LADINE_MAX_CLASSES];
#pragma unroll
      for (int c = 0; c < LADINE_MAX_CLASSES; ++c) yv[c] = c < C ? sY[lane * Cp + c] : 0.f;
      const float* A1 = p.A[k][0] + trow;
      const float* C1 = p.Cc[k][0] + trow;
#pragma unroll 4
      for (int j = 0; j < NCOL; ++j) {
        const int n = col0 + j;
        float lin = sU[n * kRows + lane];
        const float* w1 = p.W1y[k] + (size_t)n * Cp;
#pragma unroll
        for (int c = 0; c < LADINE_MAX_CLASSES; ++c)
          if (c < C) lin = fmaf(__ldg(w1 + c), yv[c], lin);
        const float v = fmaf(__ldg(A1 + n), lin, __ldg(C1 + n));
        sHa[n * kRows + lane] = softplus_precise(v) * sX[n * kRows + lane];
      }
    }
    __syncthreads();
    // ---- lin2 -> softplus (latent_model.py:178-180) ----
    dense_layer<NCOL>(sHa, sW2, Fp, col0, lane, acc);
    {
      const float* A2 = p.A[k][1] + trow;
      const float* C2 = p.Cc[k][1] + trow;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int n = col0 + j;
        sHb[n * kRows + lane] = softplus_precise(fmaf(__ldg(A2 + n), acc[j], __ldg(C2 + n)));
      }
    }
    __syncthreads();
    // ---- lin3 -> softplus -> lin4 partial (latent_model.py:181-184) ----
    dense_layer<NCOL>(sHb, sW3, Fp, col0, lane, acc);
    {
      const float* A3 = p.A[k][2] + trow;
      const float* C3 = p.Cc[k][2] + trow;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int n = col0 + j;
        acc[j] = softplus_precise(fmaf(__ldg(A3 + n), acc[j], __ldg(C3 + n)));
      }
      for (int c = 0; c < C; ++c) {
        const float* w4 = p.W4[k] + (size_t)c * Fp + col0;
        float pc = 0.f;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) pc = fmaf(acc[j], __ldg(w4 + j), pc);
        sRed[(warp * kRows + lane) * Cp + c] = pc;
      }
    }
    __syncthreads();
    // ---- eps, y_0 reparameterisation, posterior mean, + sqrt(beta_hat) z ----
    {
      const StepCoef kc = p.coef[t];
      const uint32_t slot = (uint32_t)(slot_base + (p.t_first - t));
      for (int i = tid; i < kRows * C; i += kThreads) {
        const int r = i / C, c = i % C;
        const int rg = row0 + r;
        if (rg >= p.rows) continue;
        const int n = rg / p.D, d = rg % p.D;
        float eps = __ldg(p.b4[k] + c);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) eps += sRed[(w * kRows + r) * Cp + c];
        const float y = sY[r * Cp + c], mu = sMu[r * Cp + c];
        float yn;
        if (t > 0) {
          const float z = p.noise ? __ldg(p.noise + ((((size_t)k * p.D + d) * p.n_slots + slot) * p.N + n) * C + c)
                                  : philox_normal(p.seed, p.ids.chain(k, d, n), slot, c);
          yn = posterior_step_op(kc, y, mu, eps, z);
        } else {
          yn = y0_reparam_op(kc, y, mu, eps);  // p_sample_t_1to0, diffusion_utils.py:96-111
        }
        sY[r * Cp + c] = yn;
        if (p.traj_out)
          p.traj_out[((((size_t)k * p.D + d) * p.n_traj + (slot_base + (p.t_first - t))) * p.N + n) * C + c] = yn;
      }
    }
    __syncthreads();
  }

  // ---- write y (and optionally class probabilities) in [K, D, N, C] ----
  if (warp == 0 && rg_lane < p.rows) {
    const int n = rg_lane / p.D, d = rg_lane % p.D;
    const size_t o = (((size_t)k * p.D + d) * p.N + n) * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float y = sY[lane * Cp + c];
      p.y_out[o + c] = y;
      const float lg = -(y - 1.0f) * (y - 1.0f) / p.temperature;
      mx = fmaxf(mx, lg);
    }
    if (p.prob_out) {
      float sum = 0.f;
      for (int c = 0; c < C; ++c) {
        const float y = sY[lane * Cp + c];
        sum += expf(-(y - 1.0f) * (y - 1.0f) / p.temperature - mx);
      }
      for (int c = 0; c < C; ++c) {
        const float y = sY[lane * Cp + c];
        p.prob_out[o + c] = expf(-(y - 1.0f) * (y - 1.0f) / p.temperature - mx) / sum;
      }
    }
  }
}

}  // namespace

size_t resident_smem_bytes(int Fp, int Cp) {
  return sizeof(float) * ((size_t)2 * Fp * Fp + (size_t)4 * Fp * kRows + (size_t)kWarps * kRows * Cp + 2 * kRows * Cp);
}

cudaError_t launch_resident(const ladine_handle* h, const ladine_member* const* members, const ladine_sample_args& a,
                            const ChainIds& ids, const StepCoef* d_coef, float* d_u, int n_slots, int n_traj,
                            cudaStream_t st, int64_t* launches) {
  const ladine_member* m0 = members[0];
  ResidentParams p{};
  for (int k = 0; k < a.K; ++k) {
    const ladine_member* m = members[k];
    for (int l = 0; l < 3; ++l) {
      p.A[k][l] = m->A[l];
      p.Cc[k][l] = m->Cc[l];
    }
    p.W1y[k] = m->W1y;
    p.W4[k] = m->W4;
    p.b4[k] = m->b4;
    p.W2t[k] = m->W2t;
    p.W3t[k] = m->W3t;
  }
  p.xf = a.xf;
  p.u = d_u;
  p.ytmean = a.ytmean;
  p.y_init = a.y_init;
  p.noise = a.noise;
  p.coef = d_coef;
  p.y_out = a.y_out;
  p.traj_out = a.traj_out;
  p.prob_out = a.prob_out;
  p.temperature = a.prob_out ? a.temperature : 1.0f;
  p.seed = a.seed;
  p.ids = ids;
  p.N = a.N;
  p.D = a.D;
  p.C = m0->C;
  p.Cp = m0->Cp;
  p.Fin = m0->F;
  p.Fp = m0->Fp;
  p.rows = a.N * a.D;
  p.t_first = a.t_first;
  p.t_last = a.t_last;
  p.n_slots = n_slots;
  p.n_traj = n_traj;

  const size_t smem = resident_smem_bytes(m0->Fp, m0->Cp);
  dim3 grid((p.rows + kRows - 1) / kRows, a.K);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
  };
  cudaError_t e;
  switch (m0->Fp / 8) {
    case 4: e = go(resident_chain_kernel<4>); break;
    case 8: e = go(resident_chain_kernel<8>); break;
    case 12: e = go(resident_chain_kernel<12>); break;
    case 16: e = go(resident_chain_kernel<16>); break;
    default: return cudaErrorInvalidValue;
  }
  (void)h;
  if (e == cudaSuccess) *launches += 1;
  return e;
}

}  // namespace ladine
