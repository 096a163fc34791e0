// sampler.cu -- K4: BucketedDistributedSampler (stoke/data.py:111-516) as host planning + device index kernels.
//
//   stk_randperm          host: torch's CPU randperm (mt19937 + Fisher-Yates, `random() % (n - i)`), bit-exact; the swap
//                         chain is inherently serial, so the permutations are drawn on the host and uploaded.
//   stk_sampler_plan      host: the integer sizes and the three ValueError guards (data.py:219-260).
//   stk_sampler_last_slice host: the padding rule of _handle_padding (data.py:450-498) as a position table.
//   stk_argsort_u32       device: stable LSD radix argsort (8-bit digits; histogram -> scan -> ranked scatter), the
//                         user-side `np.argsort(lengths, kind="stable")` that produces sorted_idx.
//   stk_sampler_indices   device: one thread per output index: batch shuffle -> bucket/slice -> replica stride ->
//                         (padding table | residual batches) -> bucket permutation -> sorted_idx gather.
#include <algorithm>
#include <vector>

#include "ctx.cuh"

// ---- host: mt19937 exactly as at::mt19937 (seed truncated to 32 bits) ---------------------------------------------------
namespace {
struct MT19937 {
  uint32_t mt[624];
  int idx;
  explicit MT19937(uint64_t seed) {
    mt[0] = (uint32_t)(seed & 0xffffffffu);
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  void refill() {
    for (int k = 0; k < 624; ++k) {
      uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    idx = 0;
  }
  uint32_t next() {
    if (idx >= 624) refill();
    uint32_t y = mt[idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};
}  // namespace

extern "C" {

int stk_randperm(uint64_t seed, const int64_t* lens, int k, int32_t* out) {
  if (!lens || !out || k < 0) return stk_fail(nullptr, STK_ERR_INVALID, "stk_randperm: bad argument");
  MT19937 gen(seed);
  size_t off = 0;
  for (int b = 0; b < k; ++b) {
    const int64_t n = lens[b];
    if (n < 0 || n >= (int64_t)(0xffffffffu / 20)) return stk_fail(nullptr, STK_ERR_INVALID, "stk_randperm: length out of range");
    int32_t* r = out + off;
    for (int64_t i = 0; i < n; ++i) r[i] = (int32_t)i;
    for (int64_t i = 0; i < n - 1; ++i) {
      const int64_t z = (int64_t)(gen.next() % (uint64_t)(n - i));
      std::swap(r[i], r[z + i]);
    }
    off += (size_t)n;
  }
  return STK_OK;
}

int stk_sampler_plan(stk_sampler_plan_t* p) {
  if (!p) return stk_fail(nullptr, STK_ERR_INVALID, "stk_sampler_plan: NULL plan");
  if (p->n <= 0 || p->buckets <= 0 || p->batch_size <= 0 || p->world <= 0 || p->rank < 0 || p->rank >= p->world)
    return stk_fail(nullptr, STK_ERR_INVALID, "stk_sampler_plan: bad sizes");
  auto size = [&](int64_t n, int64_t d) { return p->drop_last ? n / d : (n + d - 1) / d; };
  p->slice_size = p->batch_size * p->world;
  p->per_bucket = size(p->n, p->buckets);
  p->slices_per_bucket = size(p->per_bucket, p->slice_size);
  if (p->per_bucket < p->slice_size)
    return stk_fail(nullptr, STK_ERR_INVALID,
                    "Stoke -- Resulting number of slices (batch * replicas) per bucket (" + std::to_string(p->per_bucket) +
                        ") is less than the batch size (" + std::to_string(p->batch_size) + ")");
  if (p->slices_per_bucket < 2)
    return stk_fail(nullptr, STK_ERR_INVALID,
                    "Stoke -- Number of slices per bucket " + std::to_string(p->slices_per_bucket) +
                        " is less than 2 which is not recommended");
  if (p->per_bucket < 100)
    return stk_fail(nullptr, STK_ERR_INVALID,
                    "Stoke -- Number of samples per bucket " + std::to_string(p->per_bucket) +
                        " is less than 100 which is not recommended as this might lead to dropping of excessive data");
  p->rounded_per_bucket = p->slice_size * p->slices_per_bucket;
  p->rounded_per_replica = p->slices_per_bucket * p->batch_size * p->buckets;
  p->bucket_base = p->n / p->buckets;
  p->bucket_rem = p->n % p->buckets;
  p->n_bucket_batches = p->buckets * p->slices_per_bucket;
  p->n_overlap_batches = 0;
  if (p->allow_bucket_overlap) {
    const int64_t resid = p->n - p->rounded_per_bucket * p->buckets;
    int64_t q = resid / p->slice_size;
    if ((resid % p->slice_size != 0) && (resid < 0)) --q;  // python floor division (data.py:256-260)
    p->rounded_per_replica += q * p->batch_size;
    // data.py:419-434: residual batches exist only with drop_last and only if len(residual) > slice_size
    if (p->drop_last && resid > p->slice_size) p->n_overlap_batches = resid / p->slice_size;
  }
  p->n_batches = p->n_bucket_batches + p->n_overlap_batches;
  // smallest bucket is bucket_base long (bucket_base + 1 for the first bucket_rem)
  p->needs_padding = (p->rounded_per_bucket > p->bucket_base) ? 1 : 0;
  return STK_OK;
}

int stk_sampler_last_slice(const stk_sampler_plan_t* p, int64_t len, int32_t* out) {
  if (!p || !out) return stk_fail(nullptr, STK_ERR_INVALID, "stk_sampler_last_slice: NULL argument");
  const int64_t S = p->slice_size, W = p->world, bs = p->batch_size, ns = p->slices_per_bucket;
  const int64_t head = (ns - 1) * S;
  const int64_t n_short = len - head;  // elements already in the last slice
  if (n_short < 0 || n_short > S) return stk_fail(nullptr, STK_ERR_INVALID, "stk_sampler_last_slice: bucket length out of range");
  for (int64_t j = 0; j < n_short; ++j) out[j] = (int32_t)(head + j);
  if (n_short == S) return STK_OK;
  // np.array_split(short, W): the first (n_short % W) parts hold one more element
  std::vector<int64_t> deficit(W);
  for (int64_t r = 0; r < W; ++r) deficit[r] = bs - (n_short / W + (r < n_short % W ? 1 : 0));
  // replica r pads from bucket[r : W * deficit_r : W]
  std::vector<std::vector<int32_t>> pads(W);
  for (int64_t r = 0; r < W; ++r) {
    const int64_t stop = std::min<int64_t>(W * deficit[r], len);
    for (int64_t q = r; q < stop; q += W) pads[r].push_back((int32_t)q);
  }
  int64_t first = 0;
  bool uniform = true;
  for (int64_t r = 1; r < W; ++r) uniform &= (deficit[r] == deficit[0]);
  if (!uniform) first = std::max_element(deficit.begin(), deficit.end()) - deficit.begin();  // first largest deficit
  size_t longest = 0;
  for (auto& v : pads) longest = std::max(longest, v.size());
  int64_t j = n_short;
  for (size_t i = 0; i < longest; ++i)
    for (int64_t rr = 0; rr < W; ++rr) {
      const auto& v = pads[(first + rr) % W];
      if (i < v.size()) {
        if (j >= S) return stk_fail(nullptr, STK_ERR_INVALID, "stk_sampler_last_slice: padded slice overflows");
        out[j++] = v[i];
      }
    }
  if (j != S) return stk_fail(nullptr, STK_ERR_INVALID, "stk_sampler_last_slice: padded slice is short (reference would assert)");
  return STK_OK;
}

}  // extern "C"

// ---- device: stable LSD radix argsort -------------------------------------------------------------------------------------
namespace stk {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                       // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per block
constexpr int kSortWarps = kSortThreads / 32;

__global__ void k_iota(int32_t* idx, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int32_t)i;
}

// hist[digit * nblocks + block] = number of keys of this block's tile with that digit
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint32_t* __restrict__ keys, size_t n, int shift,
                                                             uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = size_t(blockIdx.x) * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    size_t i = base + size_t(k) * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_cnt[(keys[i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  hist[size_t(threadIdx.x) * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan over hist laid out digit-major (single block; total = 256 * nblocks entries)
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t* __restrict__ hist, size_t total) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t base = 0; base < total; base += 1024) {
    size_t i = base + threadIdx.x;
    uint32_t x = i < total ? hist[i] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += y;
      }
      s_warp[lane] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    if (i < total) hist[i] = carry + s_warp[warp] + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
    __syncthreads();
  }
}

// Stable scatter: warp w of a block owns 256 consecutive keys of the tile, lane-contiguous in 8 rounds of 32.
template <bool LAST>
__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const uint32_t* __restrict__ keys_in,
                                                                const int32_t* __restrict__ idx_in, size_t n, int shift,
                                                                const uint32_t* __restrict__ hist, int nblocks,
                                                                uint32_t* __restrict__ keys_out,
                                                                int32_t* __restrict__ idx_out, int64_t* __restrict__ idx_out64) {
  __shared__ uint32_t s_cnt[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) s_cnt[warp][d] = 0;
  __syncwarp();
  const size_t wbase = size_t(blockIdx.x) * kSortTile + size_t(warp) * (32 * kSortItems);
  uint32_t key[kSortItems];
  int32_t val[kSortItems];
  uint32_t rank_in_warp[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const size_t i = wbase + size_t(k) * 32 + lane;
    const bool ok = i < n;
    key[k] = ok ? keys_in[i] : 0xffffffffu;
    val[k] = ok ? idx_in[i] : -1;
    const uint32_t d = (key[k] >> shift) & 0xffu;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    unsigned peers = __match_any_sync(0xffffffffu, ok ? d : (0x100u + lane));
    peers &= act;
    const uint32_t before = ok ? s_cnt[warp][d] : 0u;
    rank_in_warp[k] = before + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (ok && (peers & ((1u << lane) - 1u)) == 0) s_cnt[warp][d] = before + __popc(peers);  // leader updates the count
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan of each digit's count across the warps of the block, plus the global base of (digit, block)
  {
    const int d = threadIdx.x;  // 256 threads == 256 digits
    uint32_t run = hist[size_t(d) * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = s_cnt[w][d];
      s_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const size_t i = wbase + size_t(k) * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[k] >> shift) & 0xffu;
      const uint32_t pos = s_cnt[warp][d] + rank_in_warp[k];
      if (LAST) idx_out64[pos] = (int64_t)val[k];
      else {
        keys_out[pos] = key[k];
        idx_out[pos] = val[k];
      }
    }
  }
}

// ---- device: epoch indices of one replica -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sampler_indices(const stk_sampler_plan_t p, const int64_t* __restrict__ sorted_idx,
                                                         const int32_t* __restrict__ bucket_perm,
                                                         const int32_t* __restrict__ batch_perm,
                                                         const int32_t* __restrict__ last_slice, int64_t* __restrict__ out) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= p.rounded_per_replica) return;
  const int64_t bs = p.batch_size, S = p.slice_size, W = p.world, ns = p.slices_per_bucket;
  const int64_t bdst = j / bs, t = j % bs;
  const int64_t b = batch_perm ? (int64_t)batch_perm[bdst] : bdst;
  const int64_t in_slice = p.rank + t * W;  // slice[rank : S : W][t]
  int64_t bucket, pos;
  if (b < p.n_bucket_batches) {
    bucket = b / ns;
    const int64_t k = b % ns;
    const int64_t len = p.bucket_base + (bucket < p.bucket_rem ? 1 : 0);
    pos = k * S + in_slice;
    if (k == ns - 1 && ns * S > len) pos = last_slice[(len == p.bucket_base ? S : 0) + in_slice];
  } else {
    // residual batches: tails bucket[rounded:] chained over buckets (data.py:421-433)
    const int64_t q = (b - p.n_bucket_batches) * S + in_slice;
    const int64_t r0 = p.bucket_base - p.rounded_per_bucket, r1 = r0 + 1;
    const int64_t first = p.bucket_rem * r1;
    int64_t off;
    if (q < first) {
      bucket = q / r1;
      off = q % r1;
    } else {
      bucket = p.bucket_rem + (q - first) / r0;
      off = (q - first) % r0;
    }
    pos = p.rounded_per_bucket + off;
  }
  const int64_t start = bucket * p.bucket_base + (bucket < p.bucket_rem ? bucket : p.bucket_rem);
  const int64_t src = bucket_perm ? (int64_t)bucket_perm[start + pos] : pos;
  out[j] = sorted_idx[start + src];
}

}  // namespace stk

using namespace stk;

extern "C" {

size_t stk_argsort_tmp_bytes(size_t n) {
  const size_t nblocks = (n + kSortTile - 1) / kSortTile;
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  return al(n * 4) * 2 /*keys a/b*/ + al(n * 4) * 2 /*idx a/b*/ + al(256 * nblocks * 4) + 256;
}

int stk_argsort_u32(stk_ctx* c, const uint32_t* keys, size_t n, int64_t* idx_out, void* tmp, void* stream) {
  STK_REQUIRE(c, c && keys && idx_out && tmp, "stk_argsort_u32: NULL argument");
  STK_REQUIRE(c, n > 0 && n < (size_t(1) << 31), "stk_argsort_u32: n out of range");
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int nblocks = (int)((n + kSortTile - 1) / kSortTile);
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  char* base = static_cast<char*>(tmp);
  uint32_t* kbuf[2] = {reinterpret_cast<uint32_t*>(base), reinterpret_cast<uint32_t*>(base + al(n * 4))};
  int32_t* ibuf[2] = {reinterpret_cast<int32_t*>(base + 2 * al(n * 4)), reinterpret_cast<int32_t*>(base + 3 * al(n * 4))};
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + 4 * al(n * 4));
  k_iota<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ibuf[0], n);
  const uint32_t* kin = keys;
  int cur = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    k_radix_hist<<<nblocks, kSortThreads, 0, s>>>(kin, n, shift, hist, nblocks);
    k_radix_scan<<<1, 1024, 0, s>>>(hist, size_t(256) * nblocks);
    if (pass == 3)
      k_radix_scatter<true><<<nblocks, kSortThreads, 0, s>>>(kin, ibuf[cur], n, shift, hist, nblocks, nullptr, nullptr, idx_out);
    else
      k_radix_scatter<false><<<nblocks, kSortThreads, 0, s>>>(kin, ibuf[cur], n, shift, hist, nblocks, kbuf[pass & 1],
                                                              ibuf[cur ^ 1], nullptr);
    kin = kbuf[pass & 1];
    cur ^= 1;
  }
  STK_CUDA(c, cudaGetLastError());
  return STK_OK;
}

int stk_sampler_indices(stk_ctx* c, const stk_sampler_plan_t* plan, const int64_t* sorted_idx, const int32_t* bucket_perm,
                        const int32_t* batch_perm, const int32_t* last_slice, int64_t* out, void* stream) {
  STK_REQUIRE(c, c && plan && sorted_idx && out, "stk_sampler_indices: NULL argument");
  STK_REQUIRE(c, plan->rounded_per_replica > 0 && plan->slice_size > 0, "stk_sampler_indices: plan was not filled by stk_sampler_plan");
  STK_REQUIRE(c, !plan->needs_padding || last_slice, "stk_sampler_indices: padding tables required");
  STK_REQUIRE(c, !plan->shuffle || (bucket_perm && batch_perm), "stk_sampler_indices: permutations required when shuffle is set");
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((plan->rounded_per_replica + 255) / 256);
  k_sampler_indices<<<grid, 256, 0, s>>>(*plan, sorted_idx, bucket_perm, batch_perm, last_slice, out);
  STK_CUDA(c, cudaGetLastError());
  return STK_OK;
}

}  // extern "C"
