// ctx.cuh -- host-side context shared by the translation units of libstoke_b200.so
#pragma once
#include <map>
#include <vector>
#include <mutex>
#include <string>

#include "common.cuh"

struct StepAccum {          // per-optimizer-step accumulators on the device (local shard)
  float norm_partial;       // running sum / max over the buckets reduced so far
  uint32_t found_inf;
  uint32_t blocks_done;     // last-block detection counter (self-resetting)
  uint32_t pad_;
};

struct stk_ctx {
  int rank = 0, world = 1, device = 0;
  int sm_count = 148;
  std::mutex mu;
  std::string err;
  // peer-visible allocations made through stk_mem_alloc_shared: local ptr -> peer mappings
  struct Shared {
    size_t bytes;
    void* peers[STK_MAX_WORLD];
    bool opened;
  };
  std::map<void*, Shared> shared;
  // signal pads
  stk::SignalPad* pad_local = nullptr;
  stk::PeerPads pads{};
  bool comm_ready = false;
  uint32_t blk_epoch = 0;       // block-barrier epoch (identical sequence on every rank)
  uint32_t aux_epoch[4] = {0, 0, 0, 0};
  int k1_algo = 1;              // cross-rank K1 flavour: 0 = register-staged loads (k1_reduce.cu), 1 = bulk-async (k1_bulk.cu, default)
  // device state
  stk_scaler_state_t* scaler_dev = nullptr;
  StepAccum* accum_dev = nullptr;
  float* blk_partial_dev = nullptr;   // [blk_partial_cap] per-block norm partials of K1
  float* grp_partial_dev = nullptr;   // [blk_partial_cap / 64 + 1] per-group partials (two-level ticket)
  uint32_t* grp_count_dev = nullptr;  // [blk_partial_cap / 64 + 1] group tickets (self-resetting)
  size_t blk_partial_cap = 0;
  // optional launch timing (stk_profile_*): CUDA-event pairs recorded around the kernel launch, on the launch stream
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[3];  // 0: K1 reduce, 1: K2 optimizer step, 2: accumulate
  unsigned long long* prof_ns_dev = nullptr;                   // {ns, launches} written by K1's block 0 (device timer)
  std::map<const void*, int> occupancy;                       // kernel -> resident blocks per SM (cached query)
  // pinned, mapped host scratch
  double* host_scratch = nullptr;     // [16]
  double* host_scratch_dev = nullptr; // device alias of host_scratch
};

extern thread_local std::string g_tls_err;
int stk_fail(stk_ctx* ctx, int code, const std::string& msg);

#define STK_CUDA(ctx, call)                                                                              \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return stk_fail(ctx, STK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
  } while (0)

#define STK_REQUIRE(ctx, cond, msg)                                   \
  do {                                                                \
    if (!(cond)) return stk_fail(ctx, STK_ERR_INVALID, (msg));        \
  } while (0)

int stk_grow_partials(stk_ctx* c, size_t blocks, cudaStream_t s);

struct ProfScope {  // records an event pair around a launch when profiling is on
  stk_ctx* c;
  int kind;
  cudaStream_t s;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(stk_ctx* ctx, int k, cudaStream_t st) : c(ctx), kind(k), s(st) {
    if (c->profiling && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, s);
  }
  ~ProfScope() {
    if (e0 && e1) {
      cudaEventRecord(e1, s);
      c->prof[kind].emplace_back(e0, e1);
    }
  }
};

template <typename K>
static inline int blocks_per_sm(stk_ctx* c, K kernel, int threads) {
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = c->occupancy.find(key);
  if (it != c->occupancy.end()) return it->second;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  c->occupancy[key] = per_sm;
  return per_sm;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
