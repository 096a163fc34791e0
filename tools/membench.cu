// membench.cu -- standalone HBM streaming microbenchmark used to choose launch shapes for K1/K2 (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench tools/membench.cu && ./membench
// Pattern "adam": 4 fp32 input streams, 3 fp32 + 1 bf16 output streams (30 B/elem), like k_optim_step.
// Pattern "k1":   1 bf16 input stream, 1 fp32 + 1 bf16(zero) output streams (8 B/elem), like k_grad_reduce at W=1.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  f8 r;
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& a) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]), "f"(a.v[3]), "f"(a.v[4]), "f"(a.v[5]), "f"(a.v[6]), "f"(a.v[7]) : "memory");
}
__device__ __forceinline__ uint4 ld16(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- adam-like: vec = 8 floats per thread per item --------------------------------------------------------------
template <int U, bool PERSIST>
__global__ void adam_like(const float* __restrict__ g, float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                          uint4* __restrict__ lp, size_t nvec) {
  const size_t stride = PERSIST ? size_t(gridDim.x) * blockDim.x : 0;
  size_t i0 = (size_t(blockIdx.x) * blockDim.x) * (PERSIST ? 1 : U) + threadIdx.x;
  do {
    f8 a[U], b[U], c[U], d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      size_t i = i0 + size_t(u) * (PERSIST ? stride : blockDim.x);
      if (i < nvec) { a[u] = ld8(g + i * 8); b[u] = ld8(w + i * 8); c[u] = ld8(m + i * 8); d[u] = ld8(v + i * 8); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      size_t i = i0 + size_t(u) * (PERSIST ? stride : blockDim.x);
      if (i < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float gg = a[u].v[k], mm = c[u].v[k], vv = d[u].v[k];
          mm = fmaf(0.1f, gg - mm, mm); vv = fmaf(0.001f * gg, gg, vv * 0.999f);
          b[u].v[k] -= 1e-3f * (mm / (sqrtf(vv) + 1e-8f)); c[u].v[k] = mm; d[u].v[k] = vv;
        }
        st8(w + i * 8, b[u]); st8(m + i * 8, c[u]); st8(v + i * 8, d[u]);
        uint4 q;
        q.x = __float_as_uint(b[u].v[0]) >> 16 | (__float_as_uint(b[u].v[1]) & 0xffff0000u);
        q.y = __float_as_uint(b[u].v[2]) >> 16 | (__float_as_uint(b[u].v[3]) & 0xffff0000u);
        q.z = __float_as_uint(b[u].v[4]) >> 16 | (__float_as_uint(b[u].v[5]) & 0xffff0000u);
        q.w = __float_as_uint(b[u].v[6]) >> 16 | (__float_as_uint(b[u].v[7]) & 0xffff0000u);
        st16(lp + i, q);
      }
    }
    i0 += stride * U;
  } while (PERSIST && i0 < nvec);
}

// ---- k1-like: read bf16 vec (16 B), write fp32 (32 B) + zero bf16 (16 B) ---------------------------------------
template <int U, bool PERSIST>
__global__ void k1_like(uint4* __restrict__ g, float* __restrict__ out, size_t nvec, float* __restrict__ partial) {
  const size_t stride = PERSIST ? size_t(gridDim.x) * blockDim.x : 0;
  size_t i0 = (size_t(blockIdx.x) * blockDim.x) * (PERSIST ? 1 : U) + threadIdx.x;
  float part = 0.f;
  do {
    uint4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      size_t i = i0 + size_t(u) * (PERSIST ? stride : blockDim.x);
      if (i < nvec) a[u] = ld16(g + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      size_t i = i0 + size_t(u) * (PERSIST ? stride : blockDim.x);
      if (i < nvec) {
        f8 r;
        r.v[0] = __uint_as_float(a[u].x << 16); r.v[1] = __uint_as_float(a[u].x & 0xffff0000u);
        r.v[2] = __uint_as_float(a[u].y << 16); r.v[3] = __uint_as_float(a[u].y & 0xffff0000u);
        r.v[4] = __uint_as_float(a[u].z << 16); r.v[5] = __uint_as_float(a[u].z & 0xffff0000u);
        r.v[6] = __uint_as_float(a[u].w << 16); r.v[7] = __uint_as_float(a[u].w & 0xffff0000u);
#pragma unroll
        for (int k = 0; k < 8; ++k) part = fmaf(r.v[k], r.v[k], part);
        st8(out + i * 8, r);
        st16(g + i, make_uint4(0, 0, 0, 0));
      }
    }
    i0 += stride * U;
  } while (PERSIST && i0 < nvec);
  if (part == 123.456f) partial[0] = part;
}

template <int U, bool PERSIST>
__global__ void copy_like(const float* __restrict__ a, float* __restrict__ b, size_t nvec) {
  const size_t stride = PERSIST ? size_t(gridDim.x) * blockDim.x : 0;
  size_t i0 = (size_t(blockIdx.x) * blockDim.x) * (PERSIST ? 1 : U) + threadIdx.x;
  do {
    f8 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t i = i0 + size_t(u) * (PERSIST ? stride : blockDim.x); if (i < nvec) x[u] = ld8(a + i * 8); }
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t i = i0 + size_t(u) * (PERSIST ? stride : blockDim.x); if (i < nvec) st8(b + i * 8, x[u]); }
    i0 += stride * U;
  } while (PERSIST && i0 < nvec);
}

static float time_it(void (*launch)(void*), void* ctx, void* flush, size_t flush_bytes, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int i = 0; i < iters + 3; ++i) {
    if (flush) cudaMemsetAsync(flush, i, flush_bytes);
    cudaEventRecord(e0); launch(ctx); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1); if (i >= 3) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2];
}
#include <algorithm>

struct Bufs { float *g, *w, *m, *v, *out; uint4 *lp, *g16; size_t nvec; int sms; float* partial; };
static Bufs B;
static int G_threads, G_grid_mul;

template <int U, bool P> static void run_adam(void*) {
  int threads = G_threads;
  size_t blocks = P ? size_t(B.sms) * G_grid_mul : (B.nvec + size_t(threads) * U - 1) / (size_t(threads) * U);
  adam_like<U, P><<<(unsigned)blocks, threads>>>(B.g, B.w, B.m, B.v, B.lp, B.nvec);
}
template <int U, bool P> static void run_k1(void*) {
  int threads = G_threads;
  size_t blocks = P ? size_t(B.sms) * G_grid_mul : (B.nvec + size_t(threads) * U - 1) / (size_t(threads) * U);
  k1_like<U, P><<<(unsigned)blocks, threads>>>(B.g16, B.out, B.nvec, B.partial);
}
template <int U, bool P> static void run_copy(void*) {
  int threads = G_threads;
  size_t blocks = P ? size_t(B.sms) * G_grid_mul : (B.nvec + size_t(threads) * U - 1) / (size_t(threads) * U);
  copy_like<U, P><<<(unsigned)blocks, threads>>>(B.g, B.w, B.nvec);
}

int main(int argc, char** argv) {
  size_t n = argc > 1 ? atoll(argv[1]) : 25557040;
  n = (n + 7) / 8 * 8;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  B.sms = prop.multiProcessorCount; B.nvec = n / 8;
  cudaMalloc(&B.g, n * 4); cudaMalloc(&B.w, n * 4); cudaMalloc(&B.m, n * 4); cudaMalloc(&B.v, n * 4); cudaMalloc(&B.out, n * 4);
  cudaMalloc(&B.lp, n * 2); cudaMalloc(&B.g16, n * 2); cudaMalloc(&B.partial, 4);
  cudaMemset(B.g, 0, n * 4); cudaMemset(B.w, 0, n * 4); cudaMemset(B.m, 0, n * 4); cudaMemset(B.v, 0x3f, n * 4); cudaMemset(B.g16, 0, n * 2);
  void* flush; size_t fb = 512u << 20; cudaMalloc(&flush, fb);
  printf("n=%zu sms=%d\n", n, B.sms);
  auto report = [&](const char* name, float ms, double bytes) {
    printf("%-40s threads=%4d gridmul=%2d  %8.1f us  %7.0f GB/s\n", name, G_threads, G_grid_mul, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
  };
  for (int threads : {128, 256, 512}) {
    G_threads = threads;
    G_grid_mul = 0;
    report("copy  oneshot U=1", time_it(run_copy<1, false>, 0, flush, fb, 15), n * 8.0);
    report("copy  oneshot U=4", time_it(run_copy<4, false>, 0, flush, fb, 15), n * 8.0);
    report("adam  oneshot U=1", time_it(run_adam<1, false>, 0, flush, fb, 15), n * 30.0);
    report("adam  oneshot U=2", time_it(run_adam<2, false>, 0, flush, fb, 15), n * 30.0);
    report("k1    oneshot U=1", time_it(run_k1<1, false>, 0, flush, fb, 15), n * 8.0);
    report("k1    oneshot U=2", time_it(run_k1<2, false>, 0, flush, fb, 15), n * 8.0);
    report("k1    oneshot U=4", time_it(run_k1<4, false>, 0, flush, fb, 15), n * 8.0);
    for (int mul : {1, 2, 4, 8}) {
      if (threads * mul > 2048) continue;
      G_grid_mul = mul;
      report("copy  persist U=4", time_it(run_copy<4, true>, 0, flush, fb, 15), n * 8.0);
      report("adam  persist U=1", time_it(run_adam<1, true>, 0, flush, fb, 15), n * 30.0);
      report("adam  persist U=2", time_it(run_adam<2, true>, 0, flush, fb, 15), n * 30.0);
      report("k1    persist U=2", time_it(run_k1<2, true>, 0, flush, fb, 15), n * 8.0);
      report("k1    persist U=4", time_it(run_k1<4, true>, 0, flush, fb, 15), n * 8.0);
    }
  }
  return 0;
}
