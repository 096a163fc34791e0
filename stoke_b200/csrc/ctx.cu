// ctx.cu -- context, peer-visible memory (CUDA IPC), signal pads, scaler state.  Host code + two tiny kernels.
#include <cstdlib>
#include <cstring>

#include "ctx.cuh"

thread_local std::string g_tls_err;

int stk_fail(stk_ctx* ctx, int code, const std::string& msg) {
  g_tls_err = msg;
  if (ctx) ctx->err = msg;
  return code;
}

int stk_grow_partials(stk_ctx* c, size_t blocks, cudaStream_t s) {
  if (blocks <= c->blk_partial_cap) return STK_OK;
  const size_t groups = blocks / 64 + 1;
  float *bp = nullptr, *gp = nullptr;
  uint32_t* gc = nullptr;
  STK_CUDA(c, cudaMalloc(&bp, sizeof(float) * blocks));
  STK_CUDA(c, cudaMalloc(&gp, sizeof(float) * groups));
  STK_CUDA(c, cudaMalloc(&gc, sizeof(uint32_t) * groups));
  STK_CUDA(c, cudaMemset(gc, 0, sizeof(uint32_t) * groups));
  if (c->blk_partial_dev) {
    STK_CUDA(c, cudaStreamSynchronize(s));  // a previous launch may still use the old arrays
    cudaFree(c->blk_partial_dev);
    cudaFree(c->grp_partial_dev);
    cudaFree(c->grp_count_dev);
  }
  c->blk_partial_dev = bp;
  c->grp_partial_dev = gp;
  c->grp_count_dev = gc;
  c->blk_partial_cap = blocks;
  return STK_OK;
}

extern "C" {

int stk_version(void) { return 100; }

const char* stk_last_error(stk_ctx* ctx) {
  if (ctx && !ctx->err.empty()) return ctx->err.c_str();
  return g_tls_err.c_str();
}

int stk_ctx_create(int rank, int world, int device, unsigned flags, stk_ctx** out) {
  (void)flags;
  STK_REQUIRE(nullptr, out != nullptr, "stk_ctx_create: out is NULL");
  STK_REQUIRE(nullptr, world >= 1 && world <= STK_MAX_WORLD, "stk_ctx_create: world must be in [1, 8]");
  STK_REQUIRE(nullptr, rank >= 0 && rank < world, "stk_ctx_create: rank out of range");
  int ndev = 0;
  STK_CUDA(nullptr, cudaGetDeviceCount(&ndev));
  STK_REQUIRE(nullptr, device >= 0 && device < ndev, "stk_ctx_create: no such CUDA device");
  DeviceGuard g(device);
  cudaDeviceProp prop;
  STK_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return stk_fail(nullptr, STK_ERR_UNSUPPORTED, "stoke_b200 needs an sm_100 (Blackwell) device");
  stk_ctx* c = new stk_ctx();
  c->rank = rank;
  c->world = world;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (const char* algo = std::getenv("STK_K1_ALGO")) c->k1_algo = (std::strcmp(algo, "bulk") == 0) ? 1 : 0;
  cudaError_t e;
  if ((e = cudaMalloc(&c->scaler_dev, sizeof(stk_scaler_state_t))) != cudaSuccess ||
      (e = cudaMalloc(&c->accum_dev, sizeof(StepAccum))) != cudaSuccess ||
      (e = cudaMalloc(&c->prof_ns_dev, 4 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(c->prof_ns_dev, 0, 4 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaHostAlloc(&c->host_scratch, sizeof(double) * 16, cudaHostAllocMapped)) != cudaSuccess ||
      (e = cudaHostGetDevicePointer(&c->host_scratch_dev, c->host_scratch, 0)) != cudaSuccess) {
    delete c;
    return stk_fail(nullptr, STK_ERR_CUDA, std::string("stk_ctx_create: ") + cudaGetErrorString(e));
  }
  stk_scaler_state_t st{};
  st.scale = 1.f;
  st.growth_factor = 2.f;
  st.backoff_factor = 0.5f;
  st.growth_interval = 2000;
  cudaMemcpy(c->scaler_dev, &st, sizeof(st), cudaMemcpyHostToDevice);
  cudaMemset(c->accum_dev, 0, sizeof(StepAccum));
  if (stk_grow_partials(c, stk::kMaxBlocks, nullptr) != STK_OK) {
    delete c;
    return STK_ERR_CUDA;
  }
  for (int i = 0; i < STK_MAX_WORLD; ++i) c->pads.p[i] = nullptr;
  *out = c;
  return STK_OK;
}

int stk_ctx_destroy(stk_ctx* c) {
  if (!c) return STK_OK;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->shared) {
    if (kv.second.opened)
      for (int r = 0; r < c->world; ++r)
        if (r != c->rank && kv.second.peers[r]) cudaIpcCloseMemHandle(kv.second.peers[r]);
    cudaFree(kv.first);
  }
  if (c->comm_ready)
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->pads.p[r]) cudaIpcCloseMemHandle(c->pads.p[r]);
  if (c->pad_local) cudaFree(c->pad_local);
  cudaFree(c->scaler_dev);
  cudaFree(c->accum_dev);
  cudaFree(c->blk_partial_dev);
  cudaFree(c->grp_partial_dev);
  cudaFree(c->grp_count_dev);
  cudaFree(c->prof_ns_dev);
  cudaFreeHost(c->host_scratch);
  delete c;
  return STK_OK;
}

int stk_caps(stk_ctx* c, stk_caps_t* out) {
  STK_REQUIRE(c, c && out, "stk_caps: NULL argument");
  DeviceGuard g(c->device);
  cudaDeviceProp prop;
  STK_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
  out->sm_major = prop.major;
  out->sm_minor = prop.minor;
  out->sm_count = prop.multiProcessorCount;
  out->rank = c->rank;
  out->world = c->world;
  out->device = c->device;
  out->peer_access = (c->world == 1) || c->comm_ready;
  out->multicast = 0;
  out->hbm_bytes = prop.totalGlobalMem;
  return STK_OK;
}

// ---- peer-visible memory ---------------------------------------------------------------------------------------------
int stk_mem_alloc_shared(stk_ctx* c, size_t bytes, void** local_ptr, unsigned char handle_out[STK_IPC_HANDLE_BYTES]) {
  STK_REQUIRE(c, c && local_ptr && handle_out, "stk_mem_alloc_shared: NULL argument");
  STK_REQUIRE(c, bytes > 0, "stk_mem_alloc_shared: zero bytes");
  static_assert(sizeof(cudaIpcMemHandle_t) == STK_IPC_HANDLE_BYTES, "IPC handle size");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  void* p = nullptr;
  size_t rounded = (bytes + 255) & ~size_t(255);
  STK_CUDA(c, cudaMalloc(&p, rounded));
  STK_CUDA(c, cudaMemset(p, 0, rounded));
  std::memset(handle_out, 0, STK_IPC_HANDLE_BYTES);
  if (c->world > 1) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      return stk_fail(c, STK_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, sizeof(h));
  }
  stk_ctx::Shared s{};
  s.bytes = rounded;
  s.opened = false;
  for (int r = 0; r < STK_MAX_WORLD; ++r) s.peers[r] = nullptr;
  s.peers[c->rank] = p;
  c->shared[p] = s;
  *local_ptr = p;
  return STK_OK;
}

int stk_mem_open_peers(stk_ctx* c, void* local_ptr, const unsigned char* handles, void** peer_ptrs_out) {
  STK_REQUIRE(c, c && local_ptr && peer_ptrs_out, "stk_mem_open_peers: NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->shared.find(local_ptr);
  STK_REQUIRE(c, it != c->shared.end(), "stk_mem_open_peers: pointer was not allocated by stk_mem_alloc_shared");
  DeviceGuard g(c->device);
  if (!it->second.opened && c->world > 1) {
    STK_REQUIRE(c, handles != nullptr, "stk_mem_open_peers: handles is NULL");
    for (int r = 0; r < c->world; ++r) {
      if (r == c->rank) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles + size_t(r) * STK_IPC_HANDLE_BYTES, sizeof(h));
      void* p = nullptr;
      STK_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      it->second.peers[r] = p;
    }
  }
  it->second.opened = true;
  for (int r = 0; r < c->world; ++r) peer_ptrs_out[r] = it->second.peers[r];
  return STK_OK;
}

int stk_mem_free_shared(stk_ctx* c, void* local_ptr) {
  STK_REQUIRE(c, c != nullptr, "stk_mem_free_shared: NULL ctx");
  if (!local_ptr) return STK_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->shared.find(local_ptr);
  STK_REQUIRE(c, it != c->shared.end(), "stk_mem_free_shared: unknown pointer");
  DeviceGuard g(c->device);
  if (it->second.opened)
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && it->second.peers[r]) cudaIpcCloseMemHandle(it->second.peers[r]);
  cudaFree(local_ptr);
  c->shared.erase(it);
  return STK_OK;
}

// ---- signal pads -----------------------------------------------------------------------------------------------------
int stk_comm_local(stk_ctx* c, unsigned char handle_out[STK_IPC_HANDLE_BYTES]) {
  STK_REQUIRE(c, c && handle_out, "stk_comm_local: NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  if (!c->pad_local) {
    STK_CUDA(c, cudaMalloc(&c->pad_local, sizeof(stk::SignalPad)));
    STK_CUDA(c, cudaMemset(c->pad_local, 0, sizeof(stk::SignalPad)));
    STK_CUDA(c, cudaDeviceSynchronize());
  }
  std::memset(handle_out, 0, STK_IPC_HANDLE_BYTES);
  if (c->world > 1) {
    cudaIpcMemHandle_t h;
    STK_CUDA(c, cudaIpcGetMemHandle(&h, c->pad_local));
    std::memcpy(handle_out, &h, sizeof(h));
  }
  return STK_OK;
}

int stk_comm_connect(stk_ctx* c, const unsigned char* handles) {
  STK_REQUIRE(c, c != nullptr, "stk_comm_connect: NULL ctx");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->pad_local) return stk_fail(c, STK_ERR_STATE, "stk_comm_connect before stk_comm_local");
  if (c->comm_ready) return STK_OK;
  DeviceGuard g(c->device);
  c->pads.p[c->rank] = c->pad_local;
  if (c->world > 1) {
    STK_REQUIRE(c, handles != nullptr, "stk_comm_connect: handles is NULL");
    for (int r = 0; r < c->world; ++r) {
      if (r == c->rank) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles + size_t(r) * STK_IPC_HANDLE_BYTES, sizeof(h));
      void* p = nullptr;
      STK_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      c->pads.p[r] = reinterpret_cast<stk::SignalPad*>(p);
    }
  }
  c->comm_ready = true;
  return STK_OK;
}

int stk_comm_check(stk_ctx* c, void* stream) {
  STK_REQUIRE(c, c != nullptr, "stk_comm_check: NULL ctx");
  if (!c->pad_local) return STK_OK;
  DeviceGuard g(c->device);
  uint32_t err = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  STK_CUDA(c, cudaMemcpyAsync(&err, &c->pad_local->error, sizeof(err), cudaMemcpyDeviceToHost, s));
  STK_CUDA(c, cudaStreamSynchronize(s));
  if (err) return stk_fail(c, STK_ERR_PEER, "a peer rank did not arrive within the spin bound (dead or out-of-order rank)");
  return STK_OK;
}

// ---- scaler state ----------------------------------------------------------------------------------------------------
int stk_scaler_set(stk_ctx* c, const stk_scaler_state_t* st, void* stream) {
  STK_REQUIRE(c, c && st, "stk_scaler_set: NULL argument");
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // staged through the pinned scratch so the async copy reads stable memory
  std::lock_guard<std::mutex> lk(c->mu);
  STK_CUDA(c, cudaStreamSynchronize(s));
  std::memcpy(c->host_scratch + 4, st, sizeof(*st));
  STK_CUDA(c, cudaMemcpyAsync(c->scaler_dev, c->host_scratch + 4, sizeof(*st), cudaMemcpyHostToDevice, s));
  STK_CUDA(c, cudaStreamSynchronize(s));
  return STK_OK;
}

int stk_scaler_get(stk_ctx* c, stk_scaler_state_t* st, void* stream) {
  STK_REQUIRE(c, c && st, "stk_scaler_get: NULL argument");
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  STK_CUDA(c, cudaMemcpyAsync(st, c->scaler_dev, sizeof(*st), cudaMemcpyDeviceToHost, s));
  STK_CUDA(c, cudaStreamSynchronize(s));
  return STK_OK;
}

void* stk_scaler_scale_ptr(stk_ctx* c) { return c ? static_cast<void*>(&c->scaler_dev->scale) : nullptr; }

int stk_option_set(stk_ctx* c, int key, int value) {
  STK_REQUIRE(c, c != nullptr, "stk_option_set: NULL ctx");
  std::lock_guard<std::mutex> lk(c->mu);
  switch (key) {
    case STK_OPT_K1_ALGO:
      STK_REQUIRE(c, value == 0 || value == 1, "stk_option_set: K1 algo must be 0 (ldg) or 1 (bulk)");
      c->k1_algo = value;
      return STK_OK;
    default:
      return stk_fail(c, STK_ERR_INVALID, "stk_option_set: unknown key");
  }
}

int stk_profile_enable(stk_ctx* c, int on) {
  STK_REQUIRE(c, c != nullptr, "stk_profile_enable: NULL ctx");
  std::lock_guard<std::mutex> lk(c->mu);
  c->profiling = on != 0;
  return STK_OK;
}

int stk_profile_read(stk_ctx* c, int kind, double* ms_total, int* launches) {
  STK_REQUIRE(c, c && ms_total && launches && kind >= 0 && kind < 3, "stk_profile_read: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  double tot = 0.0;
  int n = 0;
  for (auto& pr : c->prof[kind]) {
    float ms = 0.f;
    if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
      tot += ms;
      ++n;
    }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  c->prof[kind].clear();
  *ms_total = tot;
  *launches = n;
  return STK_OK;
}

int stk_profile_read_k1_device(stk_ctx* c, double* ms_total, int* launches, double* ms_zero_tail, void* stream) {
  STK_REQUIRE(c, c && ms_total && launches, "stk_profile_read_k1_device: NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long v[4] = {0, 0, 0, 0};
  STK_CUDA(c, cudaMemcpyAsync(v, c->prof_ns_dev, sizeof(v), cudaMemcpyDeviceToHost, s));
  STK_CUDA(c, cudaMemsetAsync(c->prof_ns_dev, 0, sizeof(v), s));
  STK_CUDA(c, cudaStreamSynchronize(s));
  *ms_total = (double)v[0] * 1e-6;
  *launches = (int)v[1];
  if (ms_zero_tail) *ms_zero_tail = (double)v[2] * 1e-6;
  return STK_OK;
}

int stk_shard_range(size_t n, int world, int rank, size_t* begin, size_t* end) {
  if (!begin || !end || world < 1 || rank < 0 || rank >= world) return stk_fail(nullptr, STK_ERR_INVALID, "stk_shard_range: bad argument");
  // shards are multiples of 8 elements (16 B of bf16 / 32 B of fp32) so every vector access stays aligned
  size_t vecs = (n + 7) / 8;
  size_t per = (vecs + world - 1) / world;
  size_t b = per * rank * 8, e = per * (rank + 1) * 8;
  if (b > n) b = n;
  if (e > n) e = n;
  *begin = b;
  *end = e;
  return STK_OK;
}

}  // extern "C"
