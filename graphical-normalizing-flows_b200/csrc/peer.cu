// Gradient all-reduce over NVLink peer memory: ONE kernel per step instead of a NCCL call (the data-parallel path's
// only exchange step is the average of the flat gradient bucket, 0.03 - 17 MB; the reference's nn.DataParallel gathers gradients
// on GPU 0, ImageExperiments.py:168).  Every rank owns a cudaMalloc'ed bucket + flag block, exported with cudaIpcGetMemHandle and
// opened by every other rank (one process per GPU), so that a kernel can load and store its peers' buckets directly:
//   barrier A   every rank's gradients are packed into its bucket
//   reduce      rank r sums slice r of ALL buckets (peer loads, fixed rank order: every replica gets bit-identical averages),
//               scales by 1/world and stores the result into slice r of ALL buckets (peer stores)
//   barrier B   all pushes have landed: every bucket holds the averaged gradient
// No element is touched by two ranks, so there is no third phase and no hazard with the next step's packing copy.  Barriers are
// per-rank epoch flags in peer memory (st.release.sys / ld.acquire.sys); the epoch lives on the device, so the launch replays
// inside a CUDA graph.  The grid (one CTA per SM at most) must be co-resident: the kernel is the last one of the backward.
#include "common.cuh"

#ifndef GNF_EMU
namespace gnf {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 512;
// flag block layout (unsigned words): [0, 16) barrier A arrivals, [16, 32) barrier B arrivals, 32 block counter, 33 epoch
constexpr int kPeerFlagA = 0, kPeerFlagB = 16, kPeerCounter = 32, kPeerEpoch = 33, kPeerFlagWords = 64;

struct PeerArgs {
  float* buf[kPeerMaxWorld];
  unsigned* flag[kPeerMaxWorld];
  int rank, world;
  long long n4;                 // float4 elements of the bucket
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer (or own) bucket data: never from a stale L1 line of an earlier step
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// bounded spin (a peer that never arrives traps the kernel instead of hanging the GPU forever: ~10 s)
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch) {
  for (long long it = 0; it < (1ll << 31); ++it) {
    if ((int)(ld_acquire_sys(p) - epoch) >= 0) return;
    __nanosleep(64);
  }
  __trap();
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs a) {
  __shared__ unsigned s_epoch;
  unsigned* mine = a.flag[a.rank];
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(mine + kPeerEpoch) + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  // ---- barrier A: this rank's packing copy is complete (stream order); tell everyone, wait for everyone
  if (blockIdx.x == 0 && threadIdx.x < a.world) st_release_sys(a.flag[threadIdx.x] + kPeerFlagA + a.rank, epoch);
  if (threadIdx.x < a.world) wait_flag(mine + kPeerFlagA + threadIdx.x, epoch);
  __syncthreads();
  // ---- reduce slice `rank` of every bucket, push the average into every bucket
  const long long per = (a.n4 + a.world - 1) / a.world;
  const long long lo = per * a.rank, hi = (lo + per < a.n4) ? lo + per : a.n4;
  const float inv = 1.f / (float)a.world;
  for (long long i = lo + (long long)blockIdx.x * kPeerThreads + threadIdx.x; i < hi; i += (long long)gridDim.x * kPeerThreads) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < a.world; ++r) {
      const float4 v = ld_peer(reinterpret_cast<const float4*>(a.buf[r]) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
#pragma unroll 4
    for (int r = 0; r < a.world; ++r) *(reinterpret_cast<float4*>(a.buf[r]) + i) = s;
  }
  // ---- barrier B: the last block of this rank to finish announces it (after a system-scope fence over every block's pushes)
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(mine + kPeerCounter, 1u);
    if (prev == gridDim.x - 1) {
      mine[kPeerCounter] = 0u;
      mine[kPeerEpoch] = epoch;
      __threadfence_system();
      for (int r = 0; r < a.world; ++r) st_release_sys(a.flag[r] + kPeerFlagB + a.rank, epoch);
    }
  }
  if (threadIdx.x < a.world) wait_flag(mine + kPeerFlagB + threadIdx.x, epoch);
  __syncthreads();
}

}  // namespace gnf
using namespace gnf;
#endif

extern "C" {

int gnf_peer_alloc(size_t bytes, void** out) {
#ifdef GNF_EMU
  (void)bytes; (void)out;
  return gnf::fail(GNF_ERR_UNSUPPORTED, "peer memory has no host-simulator flavour");
#else
  if (!out || bytes == 0) return fail(GNF_ERR_INVALID, "gnf_peer_alloc: bad arguments");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return fail(GNF_ERR_PEER, "gnf_peer_alloc: cudaMalloc(%zu) failed", bytes); }
  cudaMemset(p, 0, bytes);
  *out = p;
  return 0;
#endif
}

int gnf_peer_free(void* p) {
#ifndef GNF_EMU
  if (p) cudaFree(p);
#else
  (void)p;
#endif
  return 0;
}

int gnf_peer_export(const void* p, unsigned char* handle64) {
#ifdef GNF_EMU
  (void)p; (void)handle64;
  return gnf::fail(GNF_ERR_UNSUPPORTED, "peer memory has no host-simulator flavour");
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (!p || !handle64 || cudaIpcGetMemHandle(&h, const_cast<void*>(p)) != cudaSuccess) {
    cudaGetLastError();
    return fail(GNF_ERR_PEER, "gnf_peer_export: cudaIpcGetMemHandle failed");
  }
  memcpy(handle64, &h, 64);
  return 0;
#endif
}

int gnf_peer_import(const unsigned char* handle64, void** out) {
#ifdef GNF_EMU
  (void)handle64; (void)out;
  return gnf::fail(GNF_ERR_UNSUPPORTED, "peer memory has no host-simulator flavour");
#else
  cudaIpcMemHandle_t h;
  if (!handle64 || !out) return fail(GNF_ERR_INVALID, "gnf_peer_import: bad arguments");
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(GNF_ERR_PEER, "gnf_peer_import: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
  *out = p;
  return 0;
#endif
}

int gnf_peer_close(void* p) {
#ifndef GNF_EMU
  if (p) cudaIpcCloseMemHandle(p);
#else
  (void)p;
#endif
  return 0;
}

size_t gnf_peer_flag_bytes(void) {
#ifdef GNF_EMU
  return 0;
#else
  return kPeerFlagWords * sizeof(unsigned);
#endif
}

int gnf_peer_allreduce_avg(float* const* bufs, unsigned* const* flags, int rank, int world, long long numel, gnf_stream_t stream) {
#ifdef GNF_EMU
  (void)bufs; (void)flags; (void)rank; (void)world; (void)numel; (void)stream;
  return gnf::fail(GNF_ERR_UNSUPPORTED, "peer memory has no host-simulator flavour");
#else
  if (!bufs || !flags || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || numel < 0 || (numel % 4) != 0)
    return fail(GNF_ERR_INVALID, "gnf_peer_allreduce_avg: bad arguments (world <= %d, numel %% 4 == 0)", kPeerMaxWorld);
  if (numel == 0 || world == 1) return 0;
  PeerArgs a;
  for (int r = 0; r < kPeerMaxWorld; ++r) { a.buf[r] = r < world ? bufs[r] : nullptr; a.flag[r] = r < world ? flags[r] : nullptr; }
  for (int r = 0; r < world; ++r)
    if (!a.buf[r] || !a.flag[r] || (reinterpret_cast<uintptr_t>(a.buf[r]) & 15) != 0) return fail(GNF_ERR_INVALID, "gnf_peer_allreduce_avg: peer pointer %d is NULL / unaligned", r);
  a.rank = rank; a.world = world; a.n4 = numel / 4;
  const long long per = (a.n4 + world - 1) / world;
  long long blocks = (per + kPeerThreads - 1) / kPeerThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > kNumSMs) blocks = kNumSMs;                   // co-resident grid: the barrier between the blocks of a rank spins
  GNF_LAUNCH(peer_allreduce_kernel, (int)blocks, kPeerThreads, 0, (cudaStream_t)stream, a);
  return check_launch("gnf_peer_allreduce_avg");
#endif
}

}  // extern "C"
