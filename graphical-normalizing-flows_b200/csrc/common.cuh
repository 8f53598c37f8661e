// Shared declarations for libgnf_sm100: platform switch (CUDA / host SIMT simulator), error
// reporting behind the C-ABI, cp.async wrappers, Philox4x32-10.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef GNF_EMU
#include "cpu_emu.h"
#define GNF_NOINLINE
#else
#define GNF_NOINLINE __noinline__
#include <cuda_runtime.h>
#define GNF_SMEM(T, name)                                              \
  extern __shared__ __align__(1024) unsigned char _gnf_smem_raw[];     \
  T* name = reinterpret_cast<T*>(_gnf_smem_raw)
#define GNF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include "../../include/gnf.h"

namespace gnf {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
// Check the launch status of the kernels enqueued by an entry point.
int check_launch(const char* what);

#ifndef GNF_EMU
// Fork / join inside one entry point: independent groups of small kernels are enqueued on library-owned side streams between an
// event fork from, and an event join back into, the caller's stream -- inside a stream capture they become parallel branches of the
// graph.  Streams and events are created per device on first use (the warm-up call before a capture); when that is impossible the
// calls degrade to the caller's stream.
struct Branches {
  static constexpr int kSide = 2;
  cudaStream_t side[kSide];
  cudaEvent_t fork_ev[kSide], join_ev[kSide];
  bool ok;
  cudaStream_t begin(cudaStream_t main, int k) const {        // side stream k now follows everything enqueued on `main` so far
    if (!ok) return main;
    cudaEventRecord(fork_ev[k], main);
    cudaStreamWaitEvent(side[k], fork_ev[k], 0);
    return side[k];
  }
  void end(cudaStream_t main, int k) const {                  // `main` now follows everything enqueued on side stream k
    if (!ok) return;
    cudaEventRecord(join_ev[k], side[k]);
    cudaStreamWaitEvent(main, join_ev[k], 0);
  }
};
const Branches& branches();
#endif

#ifndef GNF_EMU
// Launch with the programmatic-stream-serialization attribute: the kernel may begin (its prologue, up to pdl_prologue_done()) while the
// previous kernel of the stream drains.  Only for kernels that call pdl_prologue_done() before their first global access.  Used for the
// persistent tcgen05 kernels (prologue = TMEM allocation + barrier init): cfg4 +0.9 %.  NOT for the small kernels between them: their
// early-resident CTAs, spinning in griddepcontrol.wait, take SM slots from the kernels of the step's parallel branches (-0.5 %).
bool pdl_enabled();
template <class... Params, class... Args>
static inline void launch_pdl(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  if (!pdl_enabled()) { kernel<<<grid, block, smem, s>>>(Params(args)...); return; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, Params(args)...);
}
#define GNF_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
// Called by every thread of such a kernel before its first global access (after the prologue, if it has one): lets the NEXT kernel of
// the stream start on SMs this grid has left, and holds this grid until the PREVIOUS grid has completed and flushed.  A no-op for
// launches without the attribute.
__device__ __forceinline__ void pdl_prologue_done() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#else
#define GNF_LAUNCH_PDL GNF_LAUNCH
inline void pdl_prologue_done() {}
#endif

// Zero-fill of several small buffers in ONE launch (inside a captured step every cudaMemsetAsync is a graph node of its own, ~1.5 us
// each on the critical path; the fused UMNN backward alone had ten).
struct ZeroList {
  static constexpr int kMax = 24;       // 2 GNF_MAX_LAYERS gradient tensors + a few work buffers
  float* p[kMax];
  long long n[kMax];      // floats
  int count = 0;
  void add(float* ptr, size_t floats) { if (ptr && floats && count < kMax) { p[count] = ptr; n[count] = (long long)floats; ++count; } }
};
void zero_many(const ZeroList& z, cudaStream_t s);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

#ifdef GNF_EMU
static constexpr int kNumSMs = 4;
#else
static constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
#endif

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte copies, global -> shared
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef GNF_EMU
  memcpy(smem_dst, gmem_src, 16);
#else
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef GNF_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef GNF_EMU
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011); one call -> 4 x 32 random bits.
// ---------------------------------------------------------------------------------------------
struct Philox {
  static constexpr unsigned kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __host__ __device__ static inline uint4 gen(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    unsigned c0 = (unsigned)ctr_lo, c1 = (unsigned)(ctr_lo >> 32), c2 = (unsigned)ctr_hi, c3 = (unsigned)(ctr_hi >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      unsigned hi0 = (unsigned)(((uint64_t)kM0 * c0) >> 32), lo0 = kM0 * c0;
      unsigned hi1 = (unsigned)(((uint64_t)kM1 * c2) >> 32), lo1 = kM1 * c2;
      unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += kW0; k1 += kW1;
    }
    uint4 out;
    out.x = c0; out.y = c1; out.z = c2; out.w = c3;
    return out;
  }
  // 23-bit uniform strictly inside (0,1): (n + 0.5) * 2^-23 with n < 2^23 is exactly representable in fp32
  // (n + 0.5 with n up to 2^24 - 1 is not: it would round up to 1.0 and make the Gumbel -log(-log(u)) infinite).
  __host__ __device__ static inline float u01(unsigned bits) { return ((float)(bits >> 9) + 0.5f) * (1.0f / 8388608.0f); }
};

}  // namespace gnf
