/*
 * stoke_b200.h -- C ABI of libstoke_b200.so (hand-written sm_100a kernels for the post-backward gradient path of
 * fidelity/stoke).  Plain C: opaque context, raw device pointers, sizes, and a cudaStream_t passed as void*.  No torch
 * types.  Every function returns 0 on success or a negative stk_status; the message is available through
 * stk_last_error().  No function synchronises the device unless its comment says so.  A context may be used from any
 * host thread (internally locked); all launches go to the stream the caller passes.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *
 *   stk_ctx_create / stk_comm_*     process-group bring-up            stoke/distributed.py:491-538 (DistributedDDP)
 *   stk_mem_*                        DDP bucket storage (torch Reducer buckets, NCCL buffers)  stoke/extensions.py:207-215
 *   stk_grad_accumulate              local accumulation under no_sync  stoke/distributed.py:648-669, stoke/stoke.py:978-984
 *   stk_grad_reduce                  DDP bucket copy-in + all-reduce + copy-out (stoke/extensions.py:207-215), fused with
 *                                    GradScaler.unscale_ (stoke/fp16.py:180-183,222-225) and the norm / inf reductions of
 *                                    clip_grad_norm_ (stoke/fp16.py:233); reduce-scatter flavour = fairscale SDDP/OSS
 *                                    reduce-to-owner (stoke/extensions.py:277-285)
 *   stk_optim_step                   clip_grad_norm_/clip_grad_value_ scaling (stoke/fp16.py:184,233), scaler.step gate +
 *                                    optimizer.step() (stoke/fp16.py:298,805), OSS shard step + parameter broadcast
 *                                    (stoke/extensions.py:136-141 -> fairscale OSS.step)
 *   stk_step_epilogue                scaler.update() (stoke/fp16.py:806) + zero_grad bookkeeping (stoke/utils.py:83-106)
 *   stk_loss_sync                    detach_and_sync_loss: item + barrier + all_reduce + item  stoke/distributed.py:619-646
 *   stk_barrier                      torch.distributed.barrier()      stoke/distributed.py:673
 *   stk_bcast                        DDP init param sync / per-forward buffer broadcast (broadcast_buffers=True,
 *                                    stoke/configs.py:182)
 *   stk_randperm / stk_argsort_u32 / stk_sampler_*    BucketedDistributedSampler  stoke/data.py:156-266, 380-498
 */
#ifndef STOKE_B200_H
#define STOKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STK_MAX_WORLD 8
#define STK_IPC_HANDLE_BYTES 64

typedef struct stk_ctx stk_ctx;

typedef enum {
  STK_OK = 0,
  STK_ERR_INVALID = -1,   /* bad argument */
  STK_ERR_CUDA = -2,      /* a CUDA runtime call failed (no device, OOM, launch failure, ...) */
  STK_ERR_STATE = -3,     /* call order violated (e.g. reduce before comm connect) */
  STK_ERR_PEER = -4,      /* a peer did not arrive within the spin bound (dead or out-of-order rank) */
  STK_ERR_UNSUPPORTED = -5
} stk_status;

typedef enum { STK_F32 = 0, STK_BF16 = 1, STK_F16 = 2 } stk_dtype;

typedef enum {
  STK_REDUCE_ALL = 0,     /* all-reduce: owner reduces its shard and pushes the result to every rank */
  STK_REDUCE_SCATTER = 1  /* reduce-scatter: owner keeps its shard (ZeRO-1 / OSS) */
} stk_reduce_mode;

typedef enum { STK_NORM_NONE = 0, STK_NORM_L2 = 1, STK_NORM_INF = 2, STK_NORM_P = 3 } stk_norm_kind;
typedef enum { STK_CLIP_NONE = 0, STK_CLIP_NORM = 1, STK_CLIP_VALUE = 2 } stk_clip_kind;
typedef enum { STK_OPT_ADAM = 0, STK_OPT_ADAMW = 1, STK_OPT_SGD = 2 } stk_optim_kind;

/* flags of stk_grad_reduce */
#define STK_RF_FINAL 1u        /* last bucket of this optimizer step: finish norm / found_inf (cross-rank exchange) */
#define STK_RF_ZERO_INPUT 2u   /* zero the local gradient bucket (and accumulator) after it has been consumed */
#define STK_RF_UNSCALE 4u      /* multiply by 1/loss_scale (device scaler state) and test for inf/nan */

typedef struct {
  int sm_major, sm_minor, sm_count;
  int rank, world, device;
  int peer_access;        /* 1 if every peer's memory is mapped */
  int multicast;          /* 1 if NVLS multicast objects can be created on this device (not used yet) */
  size_t hbm_bytes;
} stk_caps_t;

/* device-resident loss-scaler / step state (one per context); mirrors torch.amp.GradScaler's state_dict */
typedef struct {
  float scale;            /* current loss scale */
  float growth_factor, backoff_factor;
  int32_t growth_interval;
  int32_t growth_tracker;
  int32_t enabled;        /* 0: scale is fixed at 1 and never updated */
  int32_t found_inf;      /* result of the last finished reduce (all ranks agree) */
  float grad_norm;        /* total gradient norm of the last finished reduce (before clipping) */
  int64_t opt_steps;      /* optimizer steps actually applied (not skipped) */
  int64_t skipped_steps;
} stk_scaler_state_t;

typedef struct {
  int kind;               /* stk_optim_kind */
  double lr, beta1, beta2, eps, weight_decay;   /* Adam / AdamW */
  double momentum, dampening;                   /* SGD */
  int nesterov;
  int maximize;
  int clip_kind;          /* stk_clip_kind */
  double clip_max_norm;   /* STK_CLIP_NORM: max_norm (norm kind / p were given to stk_grad_reduce) */
  double clip_value;      /* STK_CLIP_VALUE */
} stk_optim_hyper_t;

/* ---- library / context -------------------------------------------------------------------------------------------- */
int stk_version(void);
const char* stk_last_error(stk_ctx* ctx); /* ctx may be NULL: last error of the calling thread */

int stk_ctx_create(int rank, int world, int device, unsigned flags, stk_ctx** out);
int stk_ctx_destroy(stk_ctx* ctx);
int stk_caps(stk_ctx* ctx, stk_caps_t* out);

/* ---- peer-visible device memory (cudaMalloc + CUDA IPC; the 64-byte handles are exchanged by the caller) ---------- */
int stk_mem_alloc_shared(stk_ctx* ctx, size_t bytes, void** local_ptr, unsigned char handle_out[STK_IPC_HANDLE_BYTES]);
/* handles: world * 64 bytes in rank order; peer_ptrs_out[rank] == local_ptr */
int stk_mem_open_peers(stk_ctx* ctx, void* local_ptr, const unsigned char* handles, void** peer_ptrs_out);
int stk_mem_free_shared(stk_ctx* ctx, void* local_ptr);

/* signal pads (flags + scalar slots) used by every cross-rank kernel: two-phase like the buffers above */
int stk_comm_local(stk_ctx* ctx, unsigned char handle_out[STK_IPC_HANDLE_BYTES]);
int stk_comm_connect(stk_ctx* ctx, const unsigned char* handles);
/* copies the device error word to the host (synchronises `stream`); returns STK_ERR_PEER if a spin bound was hit */
int stk_comm_check(stk_ctx* ctx, void* stream);

/* tuning knobs.  STK_OPT_K1_ALGO: how the cross-rank K1 brings the peers' 16-bit gradients into the SM -- 0 = 16-byte
 * register loads, 1 = bulk-async copies (TMA engine) staged through shared memory (default; +10..20 % bus bandwidth).  Same
 * results either way; fp32 gradients and launches with a local accumulator always take flavour 0.
 * Also settable at context creation through the environment variable STK_K1_ALGO=ldg|bulk. */
#define STK_OPT_K1_ALGO 1
int stk_option_set(stk_ctx* ctx, int key, int value);

/* launch timing for bench.py's roofline: when enabled, K1 (kind 0), K2 (kind 1) and the accumulate kernel (kind 2) are
 * bracketed by CUDA events on the launch stream; stk_profile_read synchronises those events, returns the summed
 * duration and the launch count since the last read, and clears the list. */
int stk_profile_enable(stk_ctx* ctx, int on);
int stk_profile_read(stk_ctx* ctx, int kind, double* ms_total, int* launches);
/* K1 only: time between its start and end barriers (the NVLink data phase) taken with the device timer by block 0 --
 * excludes the wait for the slowest rank to arrive, which host-side events include; ms_zero_tail (may be NULL) receives
 * the time block 0 then spent zeroing its part of the local bucket (HBM work).  Synchronises `stream`. */
int stk_profile_read_k1_device(stk_ctx* ctx, double* ms_total, int* launches, double* ms_zero_tail, void* stream);

/* ---- scaler / step state ------------------------------------------------------------------------------------------- */
int stk_scaler_set(stk_ctx* ctx, const stk_scaler_state_t* st, void* stream);
int stk_scaler_get(stk_ctx* ctx, stk_scaler_state_t* st, void* stream); /* synchronises `stream` */
void* stk_scaler_scale_ptr(stk_ctx* ctx); /* device float*: the live loss scale (for scaler.scale(loss)) */

/* ---- K1: gradient path --------------------------------------------------------------------------------------------- */
/* acc[i] (+)= float(grad[i]); optionally zero grad.  Local, HBM-bound.  first != 0 overwrites acc. */
int stk_grad_accumulate(stk_ctx* ctx, void* grad, int grad_dtype, float* acc, size_t n, int first, int zero_grad,
                        void* stream);

/* main[i] = (sum_r (grad_r[i] [+ acc_r[i]])) * mul * (1/scale)   for i in the shard this rank owns (all i if world==1),
 * written to every rank (STK_REDUCE_ALL) or kept locally (STK_REDUCE_SCATTER); accumulates the norm partial and the
 * inf/nan flag; with STK_RF_FINAL finishes them across ranks into the scaler state.
 *   grad_ptrs / acc_ptrs / out_ptrs: `world` device pointers in rank order (peer mappings); acc_ptrs may be NULL.
 *   n: elements in the bucket (multiple of 8);  mul: e.g. 1/world;  norm_kind / norm_p: which norm to accumulate. */
int stk_grad_reduce(stk_ctx* ctx, int mode, void* const* grad_ptrs, int grad_dtype, float* const* acc_ptrs,
                    void* const* out_ptrs, int out_dtype, size_t n, double mul, int norm_kind, double norm_p,
                    unsigned flags, void* stream);

/* element range [begin, end) of the shard `rank` owns in a bucket of n elements (same partition the kernels use) */
int stk_shard_range(size_t n, int world, int rank, size_t* begin, size_t* end);

/* ---- K2: fused optimizer step (+ K3 parameter all-gather when sharded) ---------------------------------------------- */
/* Updates master/exp_avg/exp_avg_sq over [0, n_local) from grad (fp32, already reduced & unscaled); applies the clip
 * coefficient from the scaler state; skips everything if found_inf.  If lp_ptrs != NULL the updated parameters are also
 * written as lp_dtype to lp_ptrs[r] + lp_offset for r in [0, lp_world) (lp_world == 1: local low-precision copy;
 * lp_world == world: sharded step pushing its shard to every rank = parameter all-gather). */
int stk_optim_step(stk_ctx* ctx, const stk_optim_hyper_t* hyper, float* master, float* exp_avg, float* exp_avg_sq,
                   const float* grad, size_t n_local, void* const* lp_ptrs, int lp_world, int lp_dtype, size_t lp_offset,
                   void* stream);
/* scaler.update(), opt_steps/skipped_steps bookkeeping, reset of the per-step accumulators */
int stk_step_epilogue(stk_ctx* ctx, void* stream);

/* ---- small collectives on the signal pads -------------------------------------------------------------------------- */
/* mean over ranks of *loss_dev (float32/bf16/f16 scalar) -> *out_host (pinned, mapped) ; synchronises `stream` */
int stk_loss_sync(stk_ctx* ctx, const void* loss_dev, int dtype, double* out_host, void* stream);
int stk_barrier(stk_ctx* ctx, void* stream);
/* every rank copies bytes from ptrs[root] to ptrs[rank] (peer pull) with start/end barriers */
int stk_bcast(stk_ctx* ctx, void* const* ptrs, size_t bytes, int root, void* stream);

/* ---- sampler (stoke/data.py) ---------------------------------------------------------------------------------------- */
/* torch.randperm(n, generator=g) for each n in lens[0..k) drawn from ONE mt19937 seeded with `seed` (host, bit-exact
 * with torch's CPU generator); out receives the k permutations back to back as int32. */
int stk_randperm(uint64_t seed, const int64_t* lens, int k, int32_t* out_host);

typedef struct {
  int64_t n, buckets, batch_size, world, rank;
  int32_t drop_last, allow_bucket_overlap, shuffle;
  /* derived (filled by stk_sampler_plan) */
  int64_t slice_size, per_bucket, slices_per_bucket, rounded_per_bucket, rounded_per_replica;
  int64_t bucket_base, bucket_rem;   /* np.array_split: first bucket_rem buckets have bucket_base + 1 elements */
  int64_t n_bucket_batches, n_overlap_batches, n_batches;
  int32_t needs_padding;
} stk_sampler_plan_t;

/* Fills the derived fields; returns STK_ERR_INVALID with the reference's message for its three ValueError guards. */
int stk_sampler_plan(stk_sampler_plan_t* plan);
/* Positions (into the permuted bucket) of the padded last slice for a bucket of length `bucket_len`:
 * out[j], j in [0, slice_size), is the position that lands at offset j of the last slice (stoke/data.py:450-498). */
int stk_sampler_last_slice(const stk_sampler_plan_t* plan, int64_t bucket_len, int32_t* out_host);

/* stable argsort of u32 keys on the device (LSD radix, 8-bit digits): idx_out[i] = index of the i-th smallest key.
 * tmp must hold stk_argsort_tmp_bytes(n) bytes. */
size_t stk_argsort_tmp_bytes(size_t n);
int stk_argsort_u32(stk_ctx* ctx, const uint32_t* keys, size_t n, int64_t* idx_out, void* tmp, void* stream);

/* out[j] for j in [0, rounded_per_replica): this replica's epoch indices.
 *   sorted_idx: int64[n] (device); bucket_perm: int32[n] per-bucket permutations back to back (NULL if !shuffle);
 *   batch_perm: int32[n_batches] (NULL if !shuffle); last_slice: int32[2][slice_size] tables for bucket lengths
 *   bucket_base + 1 and bucket_base (NULL if !needs_padding). */
int stk_sampler_indices(stk_ctx* ctx, const stk_sampler_plan_t* plan, const int64_t* sorted_idx,
                        const int32_t* bucket_perm, const int32_t* batch_perm, const int32_t* last_slice,
                        int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STOKE_B200_H */
