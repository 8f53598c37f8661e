// Host-side SIMT simulator for kernel-logic debugging (DEVELOPMENT / TEST TOOL ONLY).
//
// Compiling the .cu sources with  g++ -DGNF_EMU  runs every kernel on CPU threads: one
// std::thread per CUDA thread of a block, blocks executed one after another, __syncthreads
// = std::barrier, warp shuffles through a per-warp exchange buffer.  It exists so that the
// indexing / tiling logic of the hand-written kernels can be checked against the oracle in
// the GPU-less build container before GPU minutes are spent (tests/test_emu_*.py).  It is
// never built by __graft_entry__.build() into the product library, never loaded by the
// product package, and is not a fallback: the product path needs libgnf_sm100.so.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace emu {
struct BlockCtx {
  unsigned char* smem;
  std::barrier<>* block_bar;
  std::vector<std::unique_ptr<std::barrier<>>>* warp_bars;
  uint64_t* warp_xchg;  // [nwarps][32]
};
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_ctx = nullptr;
inline thread_local unsigned t_linear = 0;

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (nblocks == 0 || nthreads == 0) return;
  unsigned char* smem = (unsigned char*)aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
  std::barrier<> block_bar(nthreads);
  std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
  for (unsigned w = 0; w < nwarps; ++w) {
    unsigned cnt = (w == nwarps - 1) ? nthreads - 32 * w : 32;
    warp_bars.emplace_back(new std::barrier<>(cnt));
  }
  std::vector<uint64_t> xchg(nwarps * 32);
  BlockCtx ctx{smem, &block_bar, &warp_bars, xchg.data()};
  auto worker = [&](unsigned lin) {
    t_ctx = &ctx;
    t_linear = lin;
    t_blockDim = block;
    t_gridDim = grid;
    t_threadIdx = dim3(lin % block.x, (lin / block.x) % block.y, lin / (block.x * block.y));
    for (size_t b = 0; b < nblocks; ++b) {
      t_blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((size_t)grid.x * grid.y)));
      body();
      block_bar.arrive_and_wait();
    }
  };
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (unsigned i = 0; i < nthreads; ++i) th.emplace_back(worker, i);
  for (auto& t : th) t.join();
  free(smem);
}
inline void warp_sync() { (*t_ctx->warp_bars)[t_linear / 32]->arrive_and_wait(); }
template <class T>
inline T shfl_generic(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "");
  uint64_t* buf = t_ctx->warp_xchg + (t_linear / 32) * 32;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  buf[t_linear % 32] = raw;
  warp_sync();
  uint64_t got = buf[src_lane & 31];
  warp_sync();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

inline void __syncthreads() { emu::t_ctx->block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::shfl_generic(v, (int)(emu::t_linear % 32) ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned dlt, int = 32) {
  int lane = emu::t_linear % 32;
  int src = lane + (int)dlt;
  T got = emu::shfl_generic(v, src > 31 ? lane : src);
  return got;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::shfl_generic(v, src); }

inline float atomicAdd(float* addr, float val) {
  auto* a = reinterpret_cast<std::atomic<uint32_t>*>(addr);
  uint32_t old = a->load(std::memory_order_relaxed);
  for (;;) {
    float f;
    memcpy(&f, &old, 4);
    f += val;
    uint32_t nw;
    memcpy(&nw, &f, 4);
    if (a->compare_exchange_weak(old, nw)) { float r; memcpy(&r, &old, 4); return r; }
  }
}
inline unsigned atomicAdd(unsigned* addr, unsigned val) { return reinterpret_cast<std::atomic<uint32_t>*>(addr)->fetch_add(val); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }

#define GNF_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::t_ctx->smem)
#define GNF_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=] { kernel(__VA_ARGS__); })
