// k1_reduce.cu -- K1: the gradient path between backward and the optimizer step.
//
//   k_grad_accumulate : acc (+)= float(grad)                      (local micro-steps under no_sync; HBM-bound)
//   k_grad_reduce     : main = (sum over ranks of grad [+ acc]) * mul * inv_scale   for the shard this rank owns,
//                       pushed to every rank (all-reduce) or kept (reduce-scatter); fused inf/nan test and
//                       norm partial (sum g^2 | max |g| | sum |g|^p); last block finishes them across ranks.
//
// Replaces, in one pass: DDP bucket copy-in (/W), NCCL all-reduce, bucket copy-out, GradScaler.unscale_, and the two
// reduction passes of clip_grad_norm_ (see include/stoke_b200.h for reference file:line).
//
// Cross-rank protocol (world > 1): block b of every rank (1) start barrier: "my gradients are complete" (the kernel is
// stream-ordered after backward), (2) reads its vectors of the owned shard from all W peers over NVLink (16-byte loads,
// >= 8 in flight per thread), reduces in rank order in fp32 registers, and stores the result straight into every
// rank's main-grad buffer (posted 16-byte peer stores), (3) end barrier: all my reads of peer gradients and all my
// writes to peer buffers are done.  No staging buffer, no second pass: bus bytes per GPU and direction =
// (W-1)/W * n * (b_in + b_out).  All blocks are co-resident (cooperative launch) because they spin on peers.
#include "ctx.cuh"

namespace stk {

struct ReduceParams {
  PtrTable grad;      // W peer pointers to the gradient bucket
  PtrTable acc;       // W peer pointers to the fp32 local accumulators (p[0] == nullptr: none)
  PtrTable out;       // W peer pointers to the main-grad bucket
  PeerPads pads;
  stk_scaler_state_t* scaler;
  StepAccum* accum;
  float* blk_partial;   // [grid] per-block norm partials
  float* grp_partial;   // [grid / 64 + 1] per-group partials
  uint32_t* grp_count;  // [grid / 64 + 1] group tickets
  unsigned long long* prof_ns;  // optional {sum of barrier-to-barrier ns, launches} (block 0), nullptr when off
  size_t vec_begin, vec_end;  // owned shard in units of 8 elements
  size_t vec_total, vec_per_shard;  // whole bucket / shard stride (for zeroing the local bucket after the end barrier)
  float mul;
  float norm_p;
  int rank, world, n_dst;
  int norm_kind;
  uint32_t flags;
  uint32_t epoch, aux_epoch;
};

template <int DT>
struct InVec;  // 8 input elements -> 8 floats
template <>
struct InVec<STK_F32> {
  static constexpr int kBytes = 32;
  __device__ static void load(const void* base, size_t v, float (&f)[8]) {
    f8 a = ld_stream_f8(reinterpret_cast<const float*>(base) + v * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = a.v[i];
  }
  __device__ static void zero(void* base, size_t v) {
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    st_stream_f8(reinterpret_cast<float*>(base) + v * 8, z);
  }
};
template <>
struct InVec<STK_BF16> {
  static constexpr int kBytes = 16;
  __device__ static void load(const void* base, size_t v, float (&f)[8]) {
    uint4 u = ld_stream16(reinterpret_cast<const uint4*>(base) + v);
    f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
    f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
  }
  __device__ static void zero(void* base, size_t v) {
    st_stream16(reinterpret_cast<uint4*>(base) + v, make_uint4(0, 0, 0, 0));
  }
};
template <>
struct InVec<STK_F16> {
  static constexpr int kBytes = 16;
  __device__ static void load(const void* base, size_t v, float (&f)[8]) {
    uint4 u = ld_stream16(reinterpret_cast<const uint4*>(base) + v);
    f[0] = f16lo(u.x); f[1] = f16hi(u.x); f[2] = f16lo(u.y); f[3] = f16hi(u.y);
    f[4] = f16lo(u.z); f[5] = f16hi(u.z); f[6] = f16lo(u.w); f[7] = f16hi(u.w);
  }
  __device__ static void zero(void* base, size_t v) {
    st_stream16(reinterpret_cast<uint4*>(base) + v, make_uint4(0, 0, 0, 0));
  }
};

template <int DT>
__device__ __forceinline__ void store_out(void* base, size_t v, const float (&f)[8]) {
  if constexpr (DT == STK_F32) {
    st_stream_f8(reinterpret_cast<float*>(base) + v * 8, f);  // one 32-byte store: full sectors over NVLink
  } else {
    st_stream16(reinterpret_cast<uint4*>(base) + v,
                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int IN_DT>
__global__ void __launch_bounds__(256) k_grad_accumulate(void* __restrict__ grad, float* __restrict__ acc, size_t nvec,
                                                         int first, int zero_grad) {
  // one vector per thread, one block per chunk: measured 10-15% faster than a persistent grid-stride loop on B200
  // (tools/membench.cu; profiles/membench_r01.md)
  const size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v < nvec) {
    float g[8];
    InVec<IN_DT>::load(grad, v, g);
    if (!first) {
      float a[8];
      InVec<STK_F32>::load(acc, v, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += a[i];
    }
    store_out<STK_F32>(acc, v, g);
    if (zero_grad) InVec<IN_DT>::zero(grad, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int W>
struct Unroll {
  static constexpr int value = W == 1 ? 2 : (W == 2 ? 2 : 1);
};

// W_T: compile-time world size (1, 2, 4, 8) or 0 = runtime p.world
template <int IN_DT, int OUT_DT, int W_T>
__global__ void __launch_bounds__(W_T == 1 ? 256 : 512, W_T == 1 ? 6 : 1) k_grad_reduce(const ReduceParams p) {
  __shared__ float s_red[32];
  __shared__ unsigned s_bad[32];
  const int W = W_T ? W_T : p.world;
  constexpr int U = Unroll<W_T>::value;
  constexpr int WMAX = W_T ? W_T : kMaxWorld;
  const bool has_acc = p.acc.p[0] != nullptr;
  const bool zero_in = (p.flags & STK_RF_ZERO_INPUT) && W == 1;

  if (W > 1) block_barrier_all_ranks(p.pads, p.rank, W, 0, p.epoch);
  // device-side timing of the data phase (after the start barrier = after the slowest rank has arrived, up to the end
  // barrier): the NVLink time of this launch without the cross-rank launch skew that host-side events include
  unsigned long long t_begin = 0;
  if (p.prof_ns && blockIdx.x == 0 && threadIdx.x == 0) t_begin = globaltimer_ns();

  float inv_scale = 1.f;
  if (p.flags & STK_RF_UNSCALE) inv_scale = (float)(1.0 / (double)p.scaler->scale);
  const float mul = p.mul;

  float part = 0.f;   // norm partial of this thread
  bool bad = false;   // saw inf/nan

  // W == 1 (local, HBM-bound): one-shot launch, each block owns U * blockDim contiguous vectors (no grid-stride loop:
  // 10-15% faster on B200, tools/membench.cu).  W > 1 (NVLink-bound, blocks spin on peers): persistent grid-stride.
  const size_t stride = (W_T == 1) ? size_t(blockDim.x) : size_t(gridDim.x) * blockDim.x;
  const size_t first = (W_T == 1) ? p.vec_begin + size_t(blockIdx.x) * blockDim.x * U + threadIdx.x
                                  : p.vec_begin + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t step = (W_T == 1) ? ~size_t(0) / 2 : stride * U;
  for (size_t v0 = first; v0 < p.vec_end; v0 += step) {
    float g[U][WMAX][8];
    float a[U][8];
    // issue every load before the first use (latency over NVLink is ~2 us)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = v0 + u * stride;
      if (v < p.vec_end) {
#pragma unroll
        for (int r = 0; r < WMAX; ++r)
          if (r < W) InVec<IN_DT>::load(p.grad.p[r], v, g[u][r]);
      }
    }
    if (has_acc) {
      // accumulators are summed per rank before the cross-rank sum: (g_r + acc_r) is what rank r's param.grad held
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t v = v0 + u * stride;
        if (v < p.vec_end) {
#pragma unroll
          for (int r = 0; r < WMAX; ++r)
            if (r < W) {
              InVec<STK_F32>::load(p.acc.p[r], v, a[u]);
#pragma unroll
              for (int i = 0; i < 8; ++i) g[u][r][i] += a[u][i];
            }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = v0 + u * stride;
      if (v < p.vec_end) {
        float s[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = g[u][0][i];
#pragma unroll
          for (int r = 1; r < WMAX; ++r)
            if (r < W) x += g[u][r][i];
          x *= mul;
          bad |= !finitef(x);
          x *= inv_scale;
          s[i] = x;
          if (p.norm_kind == STK_NORM_L2) part = fmaf(x, x, part);
          else if (p.norm_kind == STK_NORM_INF) part = fmaxf(part, fabsf(x));
          else if (p.norm_kind == STK_NORM_P) part += __powf(fabsf(x), p.norm_p);
        }
        if (OUT_DT == STK_BF16 && p.norm_kind != STK_NORM_NONE) { /* norm is of the fp32 value, before rounding */ }
#pragma unroll
        for (int d = 0; d < WMAX; ++d)
          if (d < p.n_dst) {
            // rotate destinations so the W owners do not all hit the same peer at the same time
            int dst = p.n_dst == 1 ? p.rank : (p.rank + d) % W;
            store_out<OUT_DT>(p.out.p[dst], v, s);
          }
        if (zero_in) InVec<IN_DT>::zero(p.grad.p[0], v);
      }
    }
  }

  // ---- per-block partials (fixed tree: run-to-run deterministic); ONE block-wide barrier, then only warp 0 continues ----
  {
    const bool mx = p.norm_kind == STK_NORM_INF;
    const float wsum = mx ? warp_reduce<true>(part) : warp_reduce<false>(part);
    const unsigned wbad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
      s_red[threadIdx.x >> 5] = wsum;
      s_bad[threadIdx.x >> 5] = wbad;
    }
  }
  __syncthreads();

  // block partial (warp 0): needed before the end barrier at W > 1, where it rides to the peers with the barrier flag
  float blk = 0.f;
  unsigned any_bad = 0;
  if (threadIdx.x < 32) {
    const unsigned nwarp = (blockDim.x + 31) >> 5;
    blk = threadIdx.x < nwarp ? s_red[threadIdx.x] : 0.f;
    blk = (p.norm_kind == STK_NORM_INF) ? warp_reduce<true>(blk) : warp_reduce<false>(blk);
    any_bad = __any_sync(0xffffffffu, threadIdx.x < nwarp && s_bad[threadIdx.x] != 0);
  }

  if (W > 1) {
    // Publish this block's (norm partial, inf flag) to every rank BEFORE signalling the end barrier: the same thread then
    // does fence.sys + st.release of the flag, so whoever sees the flag sees the partial.  Every rank later sums all
    // W x grid partials in the same (rank, block) order -> bit-identical totals everywhere, with no third cross-GPU
    // round trip after the data phase.
    if (threadIdx.x < (unsigned)W) {
      RankScalars* slot = &p.pads.p[threadIdx.x]->blk_scal[p.rank][blockIdx.x];
      st_relaxed_sys_f32(&slot->norm_partial, blk);
      st_relaxed_sys_u32(&slot->found_inf, any_bad);
    }
    block_barrier_all_ranks(p.pads, p.rank, W, 1, p.epoch);
  }
  if (W > 1 && (p.flags & STK_RF_ZERO_INPUT)) {
    // Zero the LOCAL gradient bucket inside the kernel (no separate memset between backward and step).  Block b may only
    // clear what the peers' blocks b have finished reading -- exactly this block's own index pattern, replicated in every
    // shard: shard q of my bucket is read by rank q's block b at the same offsets, and that block has passed the end barrier.
    for (int q = 0; q < W; ++q) {
      const size_t qb = p.vec_per_shard * q;
      size_t qe = qb + p.vec_per_shard;
      if (qe > p.vec_total) qe = p.vec_total;
      for (size_t v = qb + size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < qe; v += size_t(gridDim.x) * blockDim.x)
        InVec<IN_DT>::zero(p.grad.p[p.rank], v);
    }
  }
  if (p.prof_ns && blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&p.prof_ns[0], globaltimer_ns() - t_begin);
    atomicAdd(&p.prof_ns[1], 1ull);
  }

  // Only warp 0 stays for the bookkeeping: the other warps retire now, so a block never sits idle on the ticket's
  // fence + atomic round trip.
  if (threadIdx.x >= 32) return;
  const unsigned lane = threadIdx.x;
  const bool mx = p.norm_kind == STK_NORM_INF;

  if (W > 1) {
    // ---- cross-rank flavour: plain ticket over the (<= SM count) co-resident blocks ----
    unsigned last = 0;
    if (lane == 0) {
      __threadfence();
      last = (atomicAdd(&p.accum->blocks_done, 1u) == gridDim.x - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // every local block has passed its end barrier, so every peer block's partial has landed in this rank's pad
    float total = 0.f;
    unsigned inf = 0;
    for (int r = 0; r < W; ++r) {
      float x = 0.f;
      for (unsigned b = lane; b < gridDim.x; b += 32) {
        const RankScalars* slot = &p.pads.p[p.rank]->blk_scal[r][b];
        const float y = ld_relaxed_sys_f32(&slot->norm_partial);
        x = mx ? fmaxf(x, y) : x + y;
        inf |= ld_relaxed_sys_u32(&slot->found_inf);
      }
      x = mx ? warp_reduce<true>(x) : warp_reduce<false>(x);
      total = mx ? fmaxf(total, x) : total + x;
    }
    inf = __any_sync(0xffffffffu, inf != 0);
    if (lane == 0) {
      float run = p.accum->norm_partial;   // running over the buckets of this optimizer step (global values)
      run = mx ? fmaxf(run, total) : run + total;
      unsigned run_inf = p.accum->found_inf | (inf ? 1u : 0u);
      p.accum->blocks_done = 0;
      if (p.flags & STK_RF_FINAL) {
        float norm = run;
        if (p.norm_kind == STK_NORM_L2) norm = sqrtf(run);
        else if (p.norm_kind == STK_NORM_P) norm = powf(run, 1.f / p.norm_p);
        p.scaler->grad_norm = norm;
        // the inf gate belongs to the loss scaler (GradScaler.step); without one the reference steps regardless
        p.scaler->found_inf = (run_inf && (p.flags & STK_RF_UNSCALE)) ? 1 : 0;
        run = 0.f;
        run_inf = 0;
      }
      p.accum->norm_partial = run;
      p.accum->found_inf = run_inf;
    }
    return;
  }

  // ---- local flavour (W == 1) ----
  // Two-level ticket (groups of 64 blocks): the last block of a group folds the group's partials, the last group folds
  // the group partials -- fixed order at both levels (deterministic), and the serial tail stays short even with the tens
  // of thousands of one-shot blocks of a W == 1 launch.
  const unsigned grp = blockIdx.x >> 6, ngroups = (gridDim.x + 63) >> 6;
  const unsigned gsize = min(64u, gridDim.x - (grp << 6));
  unsigned last = 0;
  if (lane == 0) {
    if (p.norm_kind != STK_NORM_NONE) p.blk_partial[blockIdx.x] = blk;
    if (any_bad) atomicOr(&p.accum->found_inf, 1u);
    __threadfence();
    last = (atomicAdd(&p.grp_count[grp], 1u) == gsize - 1);
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  {
    float x = 0.f;
    if (p.norm_kind != STK_NORM_NONE) {
      const float a = lane < gsize ? __ldcg(&p.blk_partial[(grp << 6) + lane]) : 0.f;
      const float b = lane + 32 < gsize ? __ldcg(&p.blk_partial[(grp << 6) + 32 + lane]) : 0.f;
      x = mx ? warp_reduce<true>(fmaxf(a, b)) : warp_reduce<false>(a + b);
    }
    last = 0;
    if (lane == 0) {
      p.grp_partial[grp] = x;
      p.grp_count[grp] = 0;
      __threadfence();
      last = (atomicAdd(&p.accum->blocks_done, 1u) == ngroups - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
  }

  // ---- last group of this rank: fold the bucket into the step accumulators; on FINAL exchange across ranks ----
  __threadfence();
  float tot = 0.f;
  if (p.norm_kind != STK_NORM_NONE) {
    float x = 0.f;
    for (unsigned i = lane; i < ngroups; i += 32) {  // fixed lane/iteration order -> deterministic
      const float y = __ldcg(&p.grp_partial[i]);
      x = mx ? fmaxf(x, y) : x + y;
    }
    tot = mx ? warp_reduce<true>(x) : warp_reduce<false>(x);
  }
  if (lane == 0) {
    float run = p.accum->norm_partial;
    run = (p.norm_kind == STK_NORM_INF) ? fmaxf(run, tot) : run + tot;
    p.accum->norm_partial = run;
    p.accum->blocks_done = 0;
    __threadfence();
  }
  if (!(p.flags & STK_RF_FINAL)) return;
  if (lane == 0) {
    const float total = p.accum->norm_partial;
    const uint32_t inf = atomicOr(&p.accum->found_inf, 0u);
    float norm = total;
    if (p.norm_kind == STK_NORM_L2) norm = sqrtf(total);
    else if (p.norm_kind == STK_NORM_P) norm = powf(total, 1.f / p.norm_p);
    p.scaler->grad_norm = norm;
    // the inf gate belongs to the loss scaler (GradScaler.step); without one the reference steps regardless
    p.scaler->found_inf = (inf && (p.flags & STK_RF_UNSCALE)) ? 1 : 0;
    p.accum->norm_partial = 0.f;
    p.accum->found_inf = 0u;
  }
}

template <int IN_DT, int OUT_DT>
static cudaError_t launch_reduce(stk_ctx* c, const ReduceParams& p, int grid, bool coop, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(p.world == 1 ? 256 : 512);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = coop ? 1 : 0;
  ProfScope prof(c, 0, s);
  switch (p.world) {
    case 1: return cudaLaunchKernelEx(&cfg, k_grad_reduce<IN_DT, OUT_DT, 1>, p);
    case 2: return cudaLaunchKernelEx(&cfg, k_grad_reduce<IN_DT, OUT_DT, 2>, p);
    case 4: return cudaLaunchKernelEx(&cfg, k_grad_reduce<IN_DT, OUT_DT, 4>, p);
    case 8: return cudaLaunchKernelEx(&cfg, k_grad_reduce<IN_DT, OUT_DT, 8>, p);
    default: return cudaLaunchKernelEx(&cfg, k_grad_reduce<IN_DT, OUT_DT, 0>, p);
  }
}

}  // namespace stk

using namespace stk;

extern "C" {

int stk_grad_accumulate(stk_ctx* c, void* grad, int grad_dtype, float* acc, size_t n, int first, int zero_grad,
                        void* stream) {
  STK_REQUIRE(c, c && grad && acc, "stk_grad_accumulate: NULL argument");
  STK_REQUIRE(c, n % 8 == 0, "stk_grad_accumulate: n must be a multiple of 8");
  if (n == 0) return STK_OK;
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t nvec = n / 8;
  const unsigned grid = (unsigned)((nvec + 255) / 256);
  ProfScope prof(c, 2, s);
  switch (grad_dtype) {
    case STK_F32: k_grad_accumulate<STK_F32><<<grid, 256, 0, s>>>(grad, acc, nvec, first, zero_grad); break;
    case STK_BF16: k_grad_accumulate<STK_BF16><<<grid, 256, 0, s>>>(grad, acc, nvec, first, zero_grad); break;
    case STK_F16: k_grad_accumulate<STK_F16><<<grid, 256, 0, s>>>(grad, acc, nvec, first, zero_grad); break;
    default: return stk_fail(c, STK_ERR_INVALID, "stk_grad_accumulate: bad dtype");
  }
  STK_CUDA(c, cudaGetLastError());
  return STK_OK;
}

int stk_grad_reduce(stk_ctx* c, int mode, void* const* grad_ptrs, int grad_dtype, float* const* acc_ptrs,
                    void* const* out_ptrs, int out_dtype, size_t n, double mul, int norm_kind, double norm_p,
                    unsigned flags, void* stream) {
  STK_REQUIRE(c, c && grad_ptrs && out_ptrs, "stk_grad_reduce: NULL argument");
  STK_REQUIRE(c, n % 8 == 0, "stk_grad_reduce: n must be a multiple of 8");
  STK_REQUIRE(c, mode == STK_REDUCE_ALL || mode == STK_REDUCE_SCATTER, "stk_grad_reduce: bad mode");
  STK_REQUIRE(c, out_dtype == STK_F32 || out_dtype == STK_BF16, "stk_grad_reduce: out dtype must be f32 or bf16");
  STK_REQUIRE(c, norm_kind >= STK_NORM_NONE && norm_kind <= STK_NORM_P, "stk_grad_reduce: bad norm kind");
  if (c->world > 1 && !c->comm_ready) return stk_fail(c, STK_ERR_STATE, "stk_grad_reduce before stk_comm_connect");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int W = c->world;

  ReduceParams p{};
  for (int r = 0; r < W; ++r) {
    STK_REQUIRE(c, grad_ptrs[r] && out_ptrs[r], "stk_grad_reduce: NULL peer pointer");
    p.grad.p[r] = grad_ptrs[r];
    p.out.p[r] = out_ptrs[r];
    p.acc.p[r] = acc_ptrs ? acc_ptrs[r] : nullptr;
  }
  p.pads = c->pads;
  p.scaler = c->scaler_dev;
  p.accum = c->accum_dev;
  p.blk_partial = c->blk_partial_dev;
  p.grp_partial = c->grp_partial_dev;
  p.grp_count = c->grp_count_dev;
  p.prof_ns = c->profiling ? c->prof_ns_dev : nullptr;
  size_t b = 0, e = 0;
  stk_shard_range(n, W, c->rank, &b, &e);
  p.vec_begin = b / 8;
  p.vec_end = (e + 7) / 8;
  p.vec_total = (n + 7) / 8;
  p.vec_per_shard = ((n + 7) / 8 + W - 1) / W;
  p.mul = (float)mul;
  p.norm_p = (float)norm_p;
  p.rank = c->rank;
  p.world = W;
  p.n_dst = (mode == STK_REDUCE_ALL) ? W : 1;
  p.norm_kind = norm_kind;
  p.flags = flags;
  p.epoch = ++c->blk_epoch;
  p.aux_epoch = (flags & STK_RF_FINAL) ? ++c->aux_epoch[0] : c->aux_epoch[0];

  // grid: identical on every rank (depends on n and W only).  Cross-rank kernels spin on peers, so every block must be
  // resident: one 512-thread block per SM, cooperative launch.
  const size_t nvec_shard = ((n + 7) / 8 + W - 1) / W;
  const int U = W == 1 ? 2 : (W == 2 ? 2 : 1);
  const size_t threads = W == 1 ? 256 : 512;
  size_t want = (nvec_shard + threads * U - 1) / (threads * U);
  int grid;
  if (W > 1) grid = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count));
  else grid = (int)std::max<size_t>(1, want);  // one-shot: every block does one chunk
  {
    int rc = stk_grow_partials(c, (size_t)grid, s);
    if (rc != STK_OK) return rc;
    p.blk_partial = c->blk_partial_dev;
    p.grp_partial = c->grp_partial_dev;
    p.grp_count = c->grp_count_dev;
  }
  const bool coop = W > 1;
  if (coop && grid > kMaxReduceBlocks) grid = kMaxReduceBlocks;

  cudaError_t err;
#define STK_DISPATCH(IN, OUT) err = launch_reduce<IN, OUT>(c, p, grid, coop, s)
  if (grad_dtype == STK_BF16 && out_dtype == STK_F32) STK_DISPATCH(STK_BF16, STK_F32);
  else if (grad_dtype == STK_BF16 && out_dtype == STK_BF16) STK_DISPATCH(STK_BF16, STK_BF16);
  else if (grad_dtype == STK_F32 && out_dtype == STK_F32) STK_DISPATCH(STK_F32, STK_F32);
  else if (grad_dtype == STK_F32 && out_dtype == STK_BF16) STK_DISPATCH(STK_F32, STK_BF16);
  else if (grad_dtype == STK_F16 && out_dtype == STK_F32) STK_DISPATCH(STK_F16, STK_F32);
  else return stk_fail(c, STK_ERR_INVALID, "stk_grad_reduce: unsupported dtype combination");
#undef STK_DISPATCH
  if (err != cudaSuccess) return stk_fail(c, STK_ERR_CUDA, std::string("k_grad_reduce launch: ") + cudaGetErrorString(err));

  return STK_OK;
}

}  // extern "C"
