// K2: NOTEARS-style acyclicity term  t = tr((I + alpha A∘A)^p) - d  and its gradient
// (DAGConditioner.get_power_trace, models/Conditionners/DAGConditioner.py:176-194).
// The multiplication chain follows torch.matrix_power (p<=3 special-cased, else binary
// decomposition) so rounding matches the reference to summation order.
//   d <= 64 : one CTA, every matrix of the chain lives in shared memory (one launch).
//   d  > 64 : host-driven chain on the fp32 tile-GEMM engine, matrices in the caller's workspace.
#include "gemm.cuh"

namespace gnf {

constexpr int kSmallD = 64;
constexpr int kPTThreads = 512;

// Shared-memory matrices are stored with a padded leading dimension kLD (zero padded to 64 x 64) so that a thread can
// own a 2-row x 4-column register block: per k it reads two A scalars (broadcast within the warp) and one float4 of B.
constexpr int kLD = kSmallD + 4;
constexpr int kMatFloats = kSmallD * kLD;

__device__ __forceinline__ void sm_matmul(float* __restrict__ C, const float* __restrict__ A, const float* __restrict__ Bm, int d) {
  const int rb = threadIdx.x >> 4, cb = threadIdx.x & 15;       // 32 row pairs x 16 column quads
  const int r0 = 2 * rb, c0 = 4 * cb;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  if (r0 < d && c0 < d) {
    const float* a0 = A + r0 * kLD;
    const float* a1 = a0 + kLD;
#pragma unroll 4
    for (int k = 0; k < d; ++k) {
      const float x0 = a0[k], x1 = a1[k];
      const float4 b = *reinterpret_cast<const float4*>(Bm + k * kLD + c0);
      acc[0][0] = fmaf(x0, b.x, acc[0][0]); acc[0][1] = fmaf(x0, b.y, acc[0][1]);
      acc[0][2] = fmaf(x0, b.z, acc[0][2]); acc[0][3] = fmaf(x0, b.w, acc[0][3]);
      acc[1][0] = fmaf(x1, b.x, acc[1][0]); acc[1][1] = fmaf(x1, b.y, acc[1][1]);
      acc[1][2] = fmaf(x1, b.z, acc[1][2]); acc[1][3] = fmaf(x1, b.w, acc[1][3]);
    }
  }
  // rows / columns >= d of every operand are zero, so the padded part of C stays zero
  *reinterpret_cast<float4*>(C + r0 * kLD + c0) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
  *reinterpret_cast<float4*>(C + (r0 + 1) * kLD + c0) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
  __syncthreads();
}
__device__ __forceinline__ void sm_copy(float* __restrict__ C, const float* __restrict__ A) {
  for (int e = threadIdx.x; e < kMatFloats; e += blockDim.x) C[e] = A[e];
  __syncthreads();
}

// Returns a pointer (inside smem) to Bm^p.  bufs: Bm, Z0, Z1, R0, R1 each kMatFloats floats.
__device__ const float* sm_matrix_power(float* smem, int d, int p) {
  float* Bm = smem;
  float* Z[2] = {smem + kMatFloats, smem + 2 * kMatFloats};
  float* R[2] = {smem + 3 * kMatFloats, smem + 4 * kMatFloats};
  if (p == 0) {
    for (int e = threadIdx.x; e < kMatFloats; e += blockDim.x) {
      const int r = e / kLD, c = e % kLD;
      R[0][e] = (r == c && r < d) ? 1.f : 0.f;
    }
    __syncthreads();
    return R[0];
  }
  if (p == 1) return Bm;
  if (p == 2) { sm_matmul(R[0], Bm, Bm, d); return R[0]; }
  if (p == 3) { sm_matmul(Z[0], Bm, Bm, d); sm_matmul(R[0], Z[0], Bm, d); return R[0]; }
  const float* z = nullptr;
  const float* res = nullptr;
  int zi = 0, ri = 0;
  while (p > 0) {
    const int bit = p & 1;
    p >>= 1;
    if (z == nullptr) {
      z = Bm;
    } else {
      sm_matmul(Z[zi], z, z, d);
      z = Z[zi];
      zi ^= 1;
    }
    if (bit) {
      if (res == nullptr) {
        sm_copy(R[ri], z);
      } else {
        sm_matmul(R[ri], res, z, d);
      }
      res = R[ri];
      ri ^= 1;
    }
  }
  return res;
}

__device__ __forceinline__ void sm_build_B(float* Bm, const float* __restrict__ A, int d, float alpha) {
  for (int e = threadIdx.x; e < kMatFloats; e += blockDim.x) {
    const int r = e / kLD, c = e % kLD;
    float v = 0.f;
    if (r < d && c < d) {
      const float a = A[r * d + c];
      v = ((r == c) ? 1.f : 0.f) + alpha * (a * a);
    }
    Bm[e] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPTThreads) power_trace_small_fwd(const float* __restrict__ A, int d, float alpha, int p, float* __restrict__ t_out) {
  GNF_SMEM(float, smem);
  sm_build_B(smem, A, d, alpha);
  const float* M = sm_matrix_power(smem, d, p);
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += M[i * kLD + i];
    *t_out = s - (float)d;
  }
}

__global__ void __launch_bounds__(kPTThreads) power_trace_small_bwd(const float* __restrict__ A, int d, float alpha, int p, const float* __restrict__ gt,
                                                                    float* __restrict__ dA) {
  GNF_SMEM(float, smem);
  if (p <= 0) {
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) dA[e] = 0.f;
    return;
  }
  sm_build_B(smem, A, d, alpha);
  const float* G = sm_matrix_power(smem, d, p - 1);
  const float scale = (*gt) * (float)p * 2.f * alpha;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    dA[e] = scale * A[e] * G[j * kLD + i];
  }
}

// Training flavour of the small-d forward: one launch leaves everything the backward needs.  G = B^(p-1) by the same binary
// chain, t = tr(G B) - d = sum_ik G[i][k] B[k][i] - d (a dot product instead of the last matrix product), and G goes to global
// memory so that the backward is one elementwise pass (trace_bwd_kernel) instead of a second single-CTA chain of p-1 products
// (21 us at d = 63, p = 13, on the critical path of every training step).  The last product's rounding order differs from
// torch.matrix_power's (B^12 B instead of B^5 B^8): ~1e-6 relative on t; the no-grad path keeps torch's order.
__global__ void __launch_bounds__(kPTThreads) power_trace_small_fwd_save(const float* __restrict__ A, int d, float alpha, int p, float* __restrict__ t_out,
                                                                         float* __restrict__ Gout) {
  GNF_SMEM(float, smem);
  if (p <= 0) {                                  // tr(I) - d = 0, no dependence on A
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) Gout[e] = 0.f;
    if (threadIdx.x == 0) *t_out = 0.f;
    return;
  }
  sm_build_B(smem, A, d, alpha);
  const float* G = sm_matrix_power(smem, d, p - 1);
  const float* Bm = smem;
  float s = 0.f;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, k = e % d;
    const float gv = G[i * kLD + k];
    Gout[e] = gv;
    s = fmaf(gv, Bm[k * kLD + i], s);
  }
  __syncthreads();                               // every read of the chain's buffers is done: reuse one as the reduction scratch
  float* red = smem + kMatFloats;
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)blockDim.x; ++i) tot += red[i];
    *t_out = tot - (float)d;
  }
}

// ------------------------------- large-d helpers -------------------------------
__global__ void build_B_kernel(const float* __restrict__ A, float* __restrict__ Bm, int d, float alpha) {
  const size_t n = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float a = A[e];
    Bm[e] = ((e / d == e % d) ? 1.f : 0.f) + alpha * (a * a);
  }
}
__global__ void eye_kernel(float* __restrict__ M, int d) {
  const size_t n = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) M[e] = (e / d == e % d) ? 1.f : 0.f;
}
__global__ void trace_kernel(const float* __restrict__ M, int d, float* __restrict__ t_out) {
  GNF_SMEM(float, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) s += M[(size_t)i * d + i];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)blockDim.x; ++i) tot += red[i];
    *t_out = tot - (float)d;
  }
}
__global__ void trace_bwd_kernel(const float* __restrict__ A, const float* __restrict__ G, int d, float alpha, int p, const float* __restrict__ gt, float* __restrict__ dA) {
  const float scale = (*gt) * (float)p * 2.f * alpha;
  const size_t n = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / d, j = e % d;
    dA[e] = scale * A[e] * G[j * d + i];
  }
}

static void big_matmul(float* C, const float* A, const float* Bm, int d, cudaStream_t s) {
  launch_gemm<TileBig>(LoadRowMajorA{A, d}, LoadRowMajorB{Bm, d}, EpiStore{C, d}, d, d, d, 1, s);
}
static inline int pt_blocks(int d) {
  size_t b = ((size_t)d * d + 255) / 256;
  return b > (size_t)4 * kNumSMs ? 4 * kNumSMs : (int)b;
}
// Host-driven torch.matrix_power chain; returns pointer to the result inside `work`.
static const float* big_matrix_power(float* work, int d, int p, cudaStream_t s) {
  const size_t n = (size_t)d * d;
  float* Bm = work;
  float* Z[2] = {work + n, work + 2 * n};
  float* R[2] = {work + 3 * n, work + 4 * n};
  if (p == 0) { GNF_LAUNCH(eye_kernel, pt_blocks(d), 256, 0, s, R[0], d); return R[0]; }
  if (p == 1) return Bm;
  if (p == 2) { big_matmul(R[0], Bm, Bm, d, s); return R[0]; }
  if (p == 3) { big_matmul(Z[0], Bm, Bm, d, s); big_matmul(R[0], Z[0], Bm, d, s); return R[0]; }
  const float* z = nullptr;
  const float* res = nullptr;
  int zi = 0, ri = 0;
  while (p > 0) {
    const int bit = p & 1;
    p >>= 1;
    if (!z) z = Bm;
    else { big_matmul(Z[zi], z, z, d, s); z = Z[zi]; zi ^= 1; }
    if (bit) {
      if (!res) cudaMemcpyAsync(R[ri], z, n * sizeof(float), cudaMemcpyDeviceToDevice, s);
      else big_matmul(R[ri], res, z, d, s);
      res = R[ri];
      ri ^= 1;
    }
  }
  return res;
}

}  // namespace gnf

using namespace gnf;

extern "C" {

size_t gnf_power_trace_workspace_bytes(int d) {
  if (d <= kSmallD) return 16;
  return (size_t)5 * d * d * sizeof(float);
}

int gnf_power_trace_fwd(const float* A, int d, float alpha, int p, float* t_out, void* work, size_t work_bytes,
                        gnf_stream_t stream) {
  if (!A || !t_out || d <= 0 || p < 0) return fail(GNF_ERR_INVALID, "gnf_power_trace_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (d <= kSmallD) {
    const size_t smem = (size_t)5 * kMatFloats * sizeof(float);
#ifndef GNF_EMU
    cudaFuncSetAttribute(power_trace_small_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(5 * kMatFloats * sizeof(float)));
#endif
    GNF_LAUNCH(power_trace_small_fwd, 1, kPTThreads, smem, s, A, d, alpha, p, t_out);
    return check_launch("gnf_power_trace_fwd");
  }
  if (!work || work_bytes < gnf_power_trace_workspace_bytes(d)) return fail(GNF_ERR_WORKSPACE, "gnf_power_trace_fwd: workspace too small");
  float* w = (float*)work;
  GNF_LAUNCH(build_B_kernel, pt_blocks(d), 256, 0, s, A, w, d, alpha);
  const float* M = big_matrix_power(w, d, p, s);
  GNF_LAUNCH(trace_kernel, 1, 256, 256 * sizeof(float), s, M, d, t_out);
  return check_launch("gnf_power_trace_fwd");
}

int gnf_power_trace_fwd_save(const float* A, int d, float alpha, int p, float* t_out, float* G_out, gnf_stream_t stream) {
  if (!A || !t_out || !G_out || d <= 0 || p < 0) return fail(GNF_ERR_INVALID, "gnf_power_trace_fwd_save: bad arguments");
  if (d > kSmallD) return fail(GNF_ERR_UNSUPPORTED, "gnf_power_trace_fwd_save: d <= %d only (larger d: gnf_power_trace_fwd / _bwd)", kSmallD);
  const size_t smem = (size_t)5 * kMatFloats * sizeof(float);
#ifndef GNF_EMU
  cudaFuncSetAttribute(power_trace_small_fwd_save, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  GNF_LAUNCH(power_trace_small_fwd_save, 1, kPTThreads, smem, (cudaStream_t)stream, A, d, alpha, p, t_out, G_out);
  return check_launch("gnf_power_trace_fwd_save");
}

int gnf_power_trace_bwd_saved(const float* A, const float* G, int d, float alpha, int p, const float* gt, float* dA, gnf_stream_t stream) {
  if (!A || !G || !gt || !dA || d <= 0 || p < 0) return fail(GNF_ERR_INVALID, "gnf_power_trace_bwd_saved: bad arguments");
  GNF_LAUNCH(trace_bwd_kernel, pt_blocks(d), 256, 0, (cudaStream_t)stream, A, G, d, alpha, p, gt, dA);
  return check_launch("gnf_power_trace_bwd_saved");
}

int gnf_power_trace_bwd(const float* A, int d, float alpha, int p, const float* gt, float* dA, void* work,
                        size_t work_bytes, gnf_stream_t stream) {
  if (!A || !gt || !dA || d <= 0 || p < 0) return fail(GNF_ERR_INVALID, "gnf_power_trace_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (d <= kSmallD) {
    const size_t smem = (size_t)5 * kMatFloats * sizeof(float);
#ifndef GNF_EMU
    cudaFuncSetAttribute(power_trace_small_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(5 * kMatFloats * sizeof(float)));
#endif
    GNF_LAUNCH(power_trace_small_bwd, 1, kPTThreads, smem, s, A, d, alpha, p, gt, dA);
    return check_launch("gnf_power_trace_bwd");
  }
  if (p == 0) {
    cudaMemsetAsync(dA, 0, (size_t)d * d * sizeof(float), s);
    return check_launch("gnf_power_trace_bwd");
  }
  if (!work || work_bytes < gnf_power_trace_workspace_bytes(d)) return fail(GNF_ERR_WORKSPACE, "gnf_power_trace_bwd: workspace too small");
  float* w = (float*)work;
  GNF_LAUNCH(build_B_kernel, pt_blocks(d), 256, 0, s, A, w, d, alpha);
  const float* G = big_matrix_power(w, d, p - 1, s);
  GNF_LAUNCH(trace_bwd_kernel, pt_blocks(d), 256, 0, s, A, G, d, alpha, p, gt, dA);
  return check_launch("gnf_power_trace_bwd");
}

}  // extern "C"
