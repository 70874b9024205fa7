/*
 * ladine.h -- C ABI of the B200-native LaDiNE nested-ensemble reverse-diffusion sampler.
 *
 * The reference (xingbpshen/nested-diffusion) is pure Python and has no FFI for this path; the
 * boundary it offers is the Python call
 *     diffusion_utils.p_sample_loop(model, x, y_0_hat, y_T_mean, n_steps, alphas,
 *                                   one_minus_alphas_bar_sqrt, only_last_sample=...)   (diffusion_utils.py:133-163)
 * driven K x 20 times per test batch by classification_train_separately.py:764-784.  This header is
 * what a ctypes / cffi / cgo / JNI stub would bind to replace that call chain; the Python mirror in
 * nested_diffusion_b200/ binds it with ctypes.
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer is a raw CUDA device pointer unless marked HOST;
 *   - all tensors are FP32, dense, row-major, in the layouts of the reference's tensors;
 *   - every call returns 0 on success or a negative ladine_status; nothing throws across the ABI;
 *     ladine_last_error() returns a human-readable message for the last failure on that handle;
 *   - work is enqueued on the caller's CUDA stream (cudaStream_t passed as void*); pointers are
 *     borrowed until that stream-ordered work completes.  ladine_sample / ladine_fill_noise never synchronise
 *     with the host in steady state; the exceptions are stated where they occur: the first call that needs a
 *     LARGER workspace than any earlier one (the old buffer is released after waiting for the handle's previous call),
 *     ladine_free_member / ladine_destroy (device synchronisation before the buffers are freed),
 *     ladine_get_profile (waits for the recorded events) and the packed-image export / import calls (they
 *     synchronise `stream` so that the HOST buffer is the caller's again on return);
 *   - a handle is bound to one device and is not re-entrant (the caller serialises calls per
 *     handle); distinct handles are independent.
 */
#ifndef LADINE_H_
#define LADINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LADINE_ABI_VERSION 3   /* 2: LADINE_PREC_FP32X, ladine_pack_encoder / ladine_encode, option "persist";
                                  3: packed images (ladine_*_export / ladine_*_import) */
#define LADINE_MAX_CLASSES 16   /* num_classes supported by the fused kernels            */
#define LADINE_MAX_GROUP   8    /* members fused into one launch group (larger K loops)  */

typedef struct ladine_handle ladine_handle;
typedef struct ladine_member ladine_member;

typedef enum {
  LADINE_OK = 0,
  LADINE_ERR_INVALID = -1,     /* bad argument (shape, null pointer, range)              */
  LADINE_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                    */
  LADINE_ERR_UNSUPPORTED = -3, /* valid in the reference but not accelerated here        */
  LADINE_ERR_NOMEM = -4        /* workspace allocation failed                            */
} ladine_status;

typedef enum {
  LADINE_PREC_AUTO = 0, /* FP32 SMEM-resident kernel when feature_dim <= 128, else FP16 tensor cores */
  LADINE_PREC_FP32 = 1, /* FP32 FFMA, weights resident in shared memory (feature_dim <= 128)         */
  LADINE_PREC_FP16 = 2, /* tcgen05 kind::f16, FP16 operands, FP32 accumulate in TMEM                 */
  LADINE_PREC_BF16 = 3, /* tcgen05 kind::f16, BF16 operands, FP32 accumulate in TMEM                 */
  LADINE_PREC_FP32X = 4 /* FP32-grade on tensor cores for any feature_dim: every GEMM operand is split into
                           FP16 hi + lo parts (W pre-scaled by a power of two) and each K slice issues three
                           tcgen05.mma (hi.hi + hi.lo + lo.hi) into the same FP32 TMEM accumulator: ~3x the
                           FP16 cost, error ~1e-6 relative (the reference is FP32: latent_model.py:169-184)  */
} ladine_precision;

/*
 * One ensemble member = the trunk of one latent_model.ConditionalModel (latent_model.py:155-184):
 * lin1..lin3 (ConditionalLinear: nn.Linear + nn.Embedding gamma table, latent_model.py:93-105),
 * unetnorm1..3 (BatchNorm1d, eval mode) and lin4.  Pointers are the module's own parameter
 * tensors (state_dict keys in the comments); the library folds and re-lays them out into buffers
 * it owns, so they may be freed after ladine_pack_member returns and the stream is synchronised.
 * The image encoder (encoder_x + norm) is step-invariant and is evaluated by the caller: see xf.
 */
typedef struct {
  uint32_t struct_size;   /* sizeof(ladine_member_desc)                                  */
  int32_t feature_dim;    /* F: config.model.feature_dim                                 */
  int32_t num_classes;    /* C: config.data.num_classes (<= LADINE_MAX_CLASSES)          */
  int32_t n_steps;        /* T: rows of the gamma tables the sampler may index (0..T-1)  */
  int32_t emb_rows;       /* rows actually present in lin*.embed.weight (>= T; T+1 in the reference) */
  int32_t guidance;       /* 1: lin1 takes cat(y, y_0_hat) [2C]; 0: y only [C]            */
  int32_t precision;      /* ladine_precision                                            */
  float bn_eps;           /* BatchNorm1d eps (1e-5)                                      */
  const float* lin1_w;    /* lin1.lin.weight  [F, 2C or C]                               */
  const float* lin1_b;    /* lin1.lin.bias    [F]                                        */
  const float* lin2_w;    /* lin2.lin.weight  [F, F]                                     */
  const float* lin2_b;    /* lin2.lin.bias    [F]                                        */
  const float* lin3_w;    /* lin3.lin.weight  [F, F]                                     */
  const float* lin3_b;    /* lin3.lin.bias    [F]                                        */
  const float* lin4_w;    /* lin4.weight      [C, F]                                     */
  const float* lin4_b;    /* lin4.bias        [C]                                        */
  const float* emb[3];    /* lin{1,2,3}.embed.weight [emb_rows, F]                       */
  const float* bn_w[3];   /* unetnorm{1,2,3}.weight        [F]                           */
  const float* bn_b[3];   /* unetnorm{1,2,3}.bias          [F]                           */
  const float* bn_mean[3];/* unetnorm{1,2,3}.running_mean  [F]                           */
  const float* bn_var[3]; /* unetnorm{1,2,3}.running_var   [F]                           */
} ladine_member_desc;

/*
 * One batched sampling call: K members x D draws x N images, reverse steps t_first .. t_last.
 * Replaces the loop at classification_train_separately.py:767-777 (K*D calls of p_sample_loop);
 * with K = D = 1 it is exactly one p_sample_loop (t_first = T-1, t_last = 0, y_init = NULL),
 * one p_sample (t_first = t_last = t, y_init = y) or p_sample_t_1to0 (t_first = t_last = 0).
 */
typedef struct {
  uint32_t struct_size;     /* sizeof(ladine_sample_args)                                */
  int32_t K;                /* members in this call                                      */
  int32_t N;                /* images                                                    */
  int32_t D;                /* posterior draws per (member, image)                       */
  int32_t T;                /* n_steps of the schedule (rows of coef)                    */
  int32_t t_first;          /* first table index evaluated (T-1 for a full chain)        */
  int32_t t_last;           /* last table index evaluated (0 for a full chain)           */
  const float* xf;          /* [K, N, F]  norm(encoder_x(x)) per member (latent_model.py:170-171) */
  const float* y0hat;       /* [K, N, C]  guidance prediction fed to eps_theta           */
  const float* ytmean;      /* [K, N, C]  prior mean y_T_mean                            */
  const float* y_init;      /* [K, D, N, C] or NULL: NULL draws y_T = y_T_mean + z (diffusion_utils.py:139-140) */
  const float* coef;        /* HOST [T, 8]: inv_q, 1-q, s, gamma0, gamma1, gamma2, sqrt(beta_hat), 0 */
  const float* noise;       /* [K, D, S, N, C] or NULL.  S = (y_init?0:1) + #steps with t>0; slot order = draw order of the reference */
  uint64_t seed;            /* Philox4x32-10 key when noise == NULL                      */
  /* chain identity for the counter-based RNG, so results do not depend on how rows are sharded */
  const int32_t* member_ids;/* HOST [K] global member index, or NULL for 0..K-1          */
  int32_t image_offset;     /* global index of image 0 of this call                      */
  int32_t images_total;     /* global image count (>= image_offset + N); 0 means N       */
  int32_t draw_offset;      /* global index of draw 0 of this call                       */
  int32_t draws_total;      /* global draw count; 0 means D                              */
  float* y_out;             /* [K, D, N, C]  y after step t_last                         */
  float* traj_out;          /* [K, D, n_traj, N, C] or NULL: every y in order (y_T first when drawn here); n_traj = (y_init?0:1) + #steps */
  float* prob_out;          /* [K, D, N, C] or NULL: softmax(-(y-1)^2 / temperature) (classification_train_separately.py:392-398) */
  float temperature;        /* used only when prob_out != NULL                           */
  void* stream;             /* cudaStream_t                                              */
} ladine_sample_args;

int ladine_version(void);

/* Bind a handle to CUDA device `device`.  Fails with LADINE_ERR_UNSUPPORTED unless the device
 * is compute capability 10.x (the kernels are sm_100a-only; there is no fallback path). */
int ladine_create(int device, ladine_handle** out);
int ladine_destroy(ladine_handle* h);
const char* ladine_last_error(const ladine_handle* h);

/* Fold + re-lay-out one member (stream-ordered on `stream`).  The result is owned by the library. */
int ladine_pack_member(ladine_handle* h, const ladine_member_desc* desc, void* stream, ladine_member** out);
int ladine_free_member(ladine_handle* h, ladine_member* m);
/* bytes of device memory held by a packed member; its resolved ladine_precision */
uint64_t ladine_member_bytes(const ladine_member* m);
int ladine_member_precision(const ladine_member* m);

/* The hot path.  members: HOST array of K packed members with identical F, C, T, precision.
 * Sizing note: the throughput per chain is flat from ~6 000 to ~80 000 chains (N x D) per member per call and falls beyond
 * (-8 % at 164 000, -19 % at 256 000: the CTAs of one very long GEMM launch drift apart in their static tile lists and the
 * tiles that share operands stop meeting in L2).  Callers with more chains should split the images over several calls --
 * image_offset / images_total keep the Philox streams, hence the samples, identical; the Python NestedEnsemble does this
 * at 32 768 chains per member. */
int ladine_sample(ladine_handle* h, const ladine_member* const* members, const ladine_sample_args* args);

/* Write the N(0,1) draws ladine_sample would generate itself for (seed, chain ids) into
 * noise [K, D, S, N, C], so a Philox run can be replayed through the injected-noise path. */
int ladine_fill_noise(ladine_handle* h, const ladine_sample_args* args, int32_t num_classes, float* noise);

/* Introspection for benchmarks/tests: kernels launched by the last ladine_sample on this handle,
 * and current workspace bytes. */
int64_t ladine_last_launches(const ladine_handle* h);
uint64_t ladine_workspace_bytes(const ladine_handle* h);

/* Tuning knobs (tensor-core path).
 *   "lanes" (1..4, default 1): groups of members advanced concurrently on internal streams (forked from /
 *       joined to the caller's stream) so one group's tail/head kernel hides under another group's GEMMs;
 *   "ctas" (0 auto | 1 | 2 | 3): GEMM tile geometry -- single CTAs (cta_group::1, 128x256 tiles), CTA pairs
 *       (cta_group::2, 256x256 tiles + 2x64-row half tiles) or slim single-CTA 128x128 tiles; auto picks pairs
 *       when the static schedule of pair tiles on sm_count / 2 SM pairs finishes >= 3 % earlier than that of single
 *       tiles on sm_count SMs, and slim tiles when the call is so small that twice as many tiles still fit the
 *       SMs in one round; all geometries give bit-identical results;
 *   "pair_gain_permille": measured per-tile speed ratio pair/single used by the auto choice (default 1080);
 *   "persist" (0 default | 1): calls of <= 4 members x <= 128 chains run as ONE cooperative launch for the whole chain
 *       (split-K over all SMs, grid barriers between the phases of a step, chain state in shared memory) instead of
 *       three launches per reverse step: ~3x faster for the reference's own call shape (one member, 70 images, one
 *       draw); split-K sums differ from the tile kernels' by FP32 rounding noise, hence opt-in; when the cooperative grid
 *       cannot be fully resident the call falls back to the tile kernels.  "persist_debug" (0 | 1): block 0 of that
 *       kernel prints its per-phase clock totals (device printf) at the end of the chain;
 *   "fuse" (0 default | 1): run the tail + head of each reverse step inside the layer-3 GEMM kernel (helper warps
 *       gated by per-row-group arrival counters) instead of a separate kernel; bitwise identical results;
 *   "order" (0 auto | 1 | 2): GEMM tile order -- N-tile-major (a W tile stays hot while a member's rows stream past
 *       it) or row-major (the activations are read once; chosen automatically when a member's activations
 *       exceed 32 MiB and would otherwise be re-read from HBM for every N tile); results are identical;
 *   "tail_vec" (0 default | 4 | 8): features per thread of the tail/head kernel; 4 (the default: more resident warps) was
 *       faster in every measurement, 8 is kept for A/B timing (results are identical either way). */
int ladine_set_option(ladine_handle* h, const char* key, int64_t value);

/* Optional per-kernel timing of the tensor-core path.  When enabled, ladine_sample brackets every
 * GEMM / tail-head launch with CUDA events on the caller's stream; ladine_get_profile synchronises
 * those events and returns, per kernel family {0: gemm layer 2, 1: gemm layer 3, 2: tail/head},
 * the summed device time in milliseconds and the launch count since the last call, then resets.
 * ms_out[3] is the union of the GEMM spans (wall time during which at least one GEMM launch was in
 * flight; equals ms_out[0] + ms_out[1] with one lane). */
int ladine_set_profiling(ladine_handle* h, int enabled);
int ladine_get_profile(ladine_handle* h, float ms_out[4], int64_t count_out[3]);

/* Debug/test entry: one trunk GEMM layer (2 or 3) of `member` at table index t on `rows` rows.
 * h_in  : [rows_pad, Fpad] 16-bit operands in the member's operand type (rows_pad = rows rounded up to 256)
 * h_out : layer 2 -> [rows_pad, Fpad] 16-bit activations; layer 3 -> unused (may be NULL)
 * part  : layer 3 -> [rows_pad, Fpad/256, 2, Cpad] FP32 partial lin4 sums (one per 128 columns); layer 2 -> unused.
 * Uses (and may grow) the handle's workspace: do not overlap with an in-flight ladine_sample on the handle. */
int ladine_debug_layer(ladine_handle* h, const ladine_member* member, int layer, int t, const void* h_in,
                       int rows, void* h_out, float* part, void* stream);
/* Debug/test entry, host only (needs no device or handle): the static tile schedule a GEMM launch would use.
 *   geometry : 1 = 128x256 single-CTA tiles, 2 = CTA pairs (256x256 + 2x64-row half tiles), 3 = slim 128x128 tiles
 *   row_major: 0 = (member, N tile, row tile) order, 1 = (member, row tile, N tile)
 *   units    : scheduling units available (SMs, or SM pairs for geometry 2)
 * Writes up to `cap` int32 entries -- units_used rows of `stride` entries, each row the tiles of one unit in order,
 * terminated/padded with -1; entry = member << 23 | n_tile << 13 | row_tile << 1 | half -- and returns the number
 * of entries the table holds (> cap: nothing written), or a negative ladine_status.
 * info_out[4] = {units_used, stride, rows_pad, row tiles per member}. */
int64_t ladine_debug_plan(int32_t K, int32_t rows, int32_t feature_dim_padded, int32_t geometry, int32_t row_major,
                          int32_t units, int32_t* table_out, int64_t cap, int32_t info_out[4]);
/* Debug/test entry, host only: the GEMM tile geometry (1, 2 or 3, as above) the auto choice takes for a launch group of K
 * members x `rows` chains each on a device with `sm_count` SMs -- the makespan of the static schedule is compared for
 * single-CTA tiles and CTA pairs (their tile counts quantise differently on a given SM count). */
int32_t ladine_debug_geometry(int32_t K, int32_t rows, int32_t feature_dim_padded, int32_t sm_count);
/*
 * Step-invariant encoder prologue  xf = norm(encoder_x(x))  of the 'linear' ConditionalModel encoder
 * (latent_model.py:126-135: Linear(data_dim, hidden) BN Softplus Linear(hidden, hidden) BN Softplus Linear(hidden, feature);
 * :155 norm = BatchNorm1d(feature); applied at :170-171 inside every denoiser call although it does not depend on t).
 * FP32-grade on the tensor cores (split FP16 operands, error-corrected accumulation).  Pointers of the descriptor are
 * the module's parameter tensors (state_dict keys encoder_x.{0,3,6}.{weight,bias}, encoder_x.{1,4}.* and norm.*);
 * the library keeps its own re-laid-out copies (2 x FP16 per weight: as many bytes as the FP32 original).
 */
typedef struct ladine_encoder ladine_encoder;
typedef struct {
  uint32_t struct_size;            /* sizeof(ladine_encoder_desc) */
  int32_t data_dim, hidden_dim, feature_dim;
  float bn_eps;                    /* nn.BatchNorm1d eps (1e-5) */
  const float* lin_w[3];           /* encoder_x.0 / .3 / .6 .weight : [hidden, data_dim], [hidden, hidden], [feature, hidden] */
  const float* lin_b[3];           /* ... .bias */
  const float* bn_w[3];            /* encoder_x.1 / encoder_x.4 / norm .weight */
  const float* bn_b[3];            /* ... .bias */
  const float* bn_mean[3];         /* ... .running_mean */
  const float* bn_var[3];          /* ... .running_var */
} ladine_encoder_desc;
int ladine_pack_encoder(ladine_handle* h, const ladine_encoder_desc* desc, void* stream, ladine_encoder** out);
int ladine_free_encoder(ladine_handle* h, ladine_encoder* enc);   /* synchronises the device */
uint64_t ladine_encoder_bytes(const ladine_encoder* enc);
/* xf_out[k] = norm(encoder_x_k(x)) for K encoders on the same images: x [N, data_dim] FP32, xf_out [K, N, feature_dim]
 * FP32, both device pointers; enqueued on `stream`.  The images are split into FP16 hi + lo operands once for all K. */
int ladine_encode(ladine_handle* h, const ladine_encoder* const* encoders, int32_t K, const float* x, int32_t N,
                  float* xf_out, void* stream);
int64_t ladine_last_encoder_launches(const ladine_handle* h);

/*
 * Packed images: a packed member / encoder serialised into HOST memory and restored from it, so that a runner can keep
 * the packed form of a checkpoint on disk (keyed by a hash of the checkpoint's content) and skip building the module,
 * loading state['noise_estimator'] into it (classification_train_separately.py:684-697) and re-packing on every run.
 * An image = 128-byte header (magic, ABI version, packed-layout version, dimensions, payload size, checksum) + the packed
 * device buffers.  Import refuses an image of another ABI / layout version, of the wrong kind, truncated or corrupt
 * (LADINE_ERR_INVALID, reason in ladine_last_error) -- the caller then packs from the checkpoint again.
 * host_dst / host_src: HOST pointers (8-byte aligned; pinned memory makes the copies fast).  Both calls synchronise
 * `stream` before returning, so the host buffer is the caller's again on return.  The image is device-independent:
 * it may be imported on any handle.
 */
uint64_t ladine_member_image_bytes(const ladine_member* m);
int ladine_member_export(ladine_handle* h, const ladine_member* m, void* host_dst, uint64_t capacity, void* stream);
int ladine_member_import(ladine_handle* h, const void* host_src, uint64_t bytes, void* stream, ladine_member** out);
uint64_t ladine_encoder_image_bytes(const ladine_encoder* enc);
int ladine_encoder_export(ladine_handle* h, const ladine_encoder* enc, void* host_dst, uint64_t capacity, void* stream);
int ladine_encoder_import(ladine_handle* h, const void* host_src, uint64_t bytes, void* stream, ladine_encoder** out);
/* Host-only inspection (no device, no handle): kind_out = 1 (member) / 2 (encoder), dims_out = the header's dimension
 * words (member: F, Fp, C, Cp, T, guidance, precision, split; encoder: data_dim, hidden_dim, feature_dim, eps bits).
 * Returns LADINE_OK for an intact image of this library's ABI / layout version, else LADINE_ERR_INVALID with the reason in
 * *why_out (a static string).  Image layout: bytes 0..7 magic "LADINEM\0" / "LADINEE\0", u32 ABI version, u32 packed-layout
 * version, 20 x i32 dimensions, u64 payload bytes, u64 checksum of the payload (ladine_image_checksum), header padded to
 * 128 bytes, then the packed buffers, each padded to 16 bytes. */
int ladine_image_info(const void* host_src, uint64_t bytes, int32_t* kind_out, int32_t dims_out[20], const char** why_out);
uint64_t ladine_image_checksum(const void* host_src, uint64_t bytes);
/* dimensions of a packed object (what a wrapper needs after an import):
 * member  -> dims_out[6] = {feature_dim, num_classes, n_steps, guidance, precision, device}
 * encoder -> dims_out[4] = {data_dim, hidden_dim, feature_dim, device} */
int ladine_member_dims(const ladine_member* m, int32_t dims_out[6]);
int ladine_encoder_dims(const ladine_encoder* enc, int32_t dims_out[4]);

/* padded feature dim and padded class count used by the packed layout */
int ladine_member_fpad(const ladine_member* m);
int ladine_member_cpad(const ladine_member* m);

#ifdef __cplusplus
}
#endif
#endif /* LADINE_H_ */
