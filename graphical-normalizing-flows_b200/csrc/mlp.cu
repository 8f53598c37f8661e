// Conditioner MLP engine: Linear(+ReLU) forward / dgrad / wgrad on the fp32 FFMA tile GEMM, the
// DAGConditioner's first layer with the masked embedding generated inside the operand loader, and
// the small pack / reduce helpers around them.  See include/gnf.h for the reference lines each
// entry point replaces.
#include "gemm.cuh"

namespace gnf {

// ------------------------------------------------------------------------------------------------
// DAG gate: e[b,i,j] and its partial derivatives  (DAGConditioner.py:94-153)
// ------------------------------------------------------------------------------------------------
struct GateCtx {
  const float* x;   // [B,d]
  const float* P;   // [d,d] importance table
  const float* n1;  // replayed noise (nullable)
  const float* n2;
  uint64_t seed, offset;
  const uint64_t* offset_dev;  // optional device-side addend to `offset`
  float T;
  int mode, d;
};

__device__ __forceinline__ void gate_noise(const GateCtx& g, size_t idx, float& a, float& b) {
  if (g.n1) {
    a = __ldg(g.n1 + idx);
    b = g.n2 ? __ldg(g.n2 + idx) : 0.f;
    return;
  }
  const uint4 r = Philox::gen(g.seed, (uint64_t)idx, g.offset + (g.offset_dev ? *g.offset_dev : 0ull));
  if (g.mode == GNF_GATE_GUMBEL) {
    a = Philox::u01(r.x);
    b = Philox::u01(r.y);
  } else {  // standard normal by Box-Muller
    const float u = Philox::u01(r.x), v = Philox::u01(r.y);
    a = sqrtf(-2.f * logf(u)) * cosf(6.283185307179586f * v);
    b = 0.f;
  }
}

// The heavy part of the stochastic gate is NOT inlined: the tile GEMM evaluates the gate once per element of its fully
// unrolled register tiles (32-64 call sites per kernel); inlined, Philox + six transcendentals per site made the layer-1
// kernels 0.2-0.4 MB of code and ncu showed 23-53 % of their warp samples stalled on instruction fetch (stall_no_inst).
// Everything goes in and out by value (registers): an out-of-line gate_e taking the context by reference and returning
// through pointers cost 20 % at d = 784, where the gate math is the whole kernel.
//   returns (G, dG/dp) of the relaxed Bernoulli gate, DAGConditioner.stochastic_gate (DAGConditioner.py:94-103)
__device__ __forceinline__ float2 gumbel_gate_math_inl(float p, float na, float nb, float T) {
  const float eps = 1e-6f;
  const float g1 = -logf(-logf(na)), g2 = -logf(-logf(nb));
  const float z1 = expf((logf(p + eps) + g1) / T);
  const float z2 = expf((logf(1.f - p + eps) + g2) / T);
  const float G = z1 / (z1 + z2);
  return make_float2(G, (G * (1.f - G) / T) * (1.f / (p + eps) + 1.f / (1.f - p + eps)));
}
// same with the uniforms drawn in place (training path: no replayed noise)
__device__ __forceinline__ float2 gumbel_gate_philox_inl(float p, float T, uint64_t seed, uint64_t idx, uint64_t offset) {
  const uint4 r = Philox::gen(seed, idx, offset);
  return gumbel_gate_math_inl(p, Philox::u01(r.x), Philox::u01(r.y), T);
}
__device__ GNF_NOINLINE float2 gumbel_gate_math(float p, float na, float nb, float T) { return gumbel_gate_math_inl(p, na, nb, T); }
__device__ GNF_NOINLINE float2 gumbel_gate_philox(float p, float T, uint64_t seed, uint64_t idx, uint64_t offset) {
  return gumbel_gate_philox_inl(p, T, seed, idx, offset);
}

// standard normal by Box-Muller (DAGConditioner.noiser_gate, DAGConditioner.py:114-116), out of line for the same reason
__device__ GNF_NOINLINE float normal_philox(uint64_t seed, uint64_t idx, uint64_t offset) {
  const uint4 r = Philox::gen(seed, idx, offset);
  const float u = Philox::u01(r.x), v = Philox::u01(r.y);
  return sqrtf(-2.f * logf(u)) * cosf(6.283185307179586f * v);
}

// kInl: inline the Gumbel math after all -- for wide flows (d >= 256: d^2 gate evaluations per sample) the gate math IS the
// kernel and the call overhead costs 10-15 % (cfg5, d = 784), while the short-K kernels of narrow flows are fetch-bound.
template <bool kGrad, bool kInl>
__device__ __forceinline__ float gate_e(const GateCtx& g, int b, int i, int j, float* de_dx, float* de_dP) {
  const float p = __ldg(g.P + (size_t)i * g.d + j);
  const float xv = __ldg(g.x + (size_t)b * g.d + j);
  if (g.mode == GNF_GATE_TABLE) {
    if (kGrad) { *de_dx = p; *de_dP = xv; }
    return xv * p;
  }
  const size_t idx = ((size_t)b * g.d + i) * g.d + j;
  if (g.mode == GNF_GATE_GUMBEL) {
    float2 G;
    const uint64_t off = g.offset + (g.offset_dev ? *g.offset_dev : 0ull);
    if (g.n1) G = kInl ? gumbel_gate_math_inl(p, __ldg(g.n1 + idx), __ldg(g.n2 + idx), g.T) : gumbel_gate_math(p, __ldg(g.n1 + idx), __ldg(g.n2 + idx), g.T);
    else G = kInl ? gumbel_gate_philox_inl(p, g.T, g.seed, (uint64_t)idx, off) : gumbel_gate_philox(p, g.T, g.seed, (uint64_t)idx, off);
    if (kGrad) {
      *de_dx = G.x;
      *de_dP = xv * G.y;
    }
    return xv * G.x;
  }
  // noiser gate: e = P*(x + n*sqrt((1-P)^2))
  const float na = g.n1 ? __ldg(g.n1 + idx) : normal_philox(g.seed, (uint64_t)idx, g.offset + (g.offset_dev ? *g.offset_dev : 0ull));
  const float a = fabsf(1.f - p);
  if (kGrad) {
    const float sgn = (1.f - p) > 0.f ? 1.f : ((1.f - p) < 0.f ? -1.f : 0.f);
    *de_dx = p;
    *de_dP = xv + na * a - p * na * sgn;
  }
  return p * (xv + na * a);
}

constexpr int kGateInlineMinD = 256;

template <bool kInl>
struct LoadDagA {  // A(m,k) = e[b,i,j], m = b*d+i, k = j
  static constexpr bool kContigK = true;
  GateCtx g;
  __device__ __forceinline__ float operator()(int m, int k) const { return gate_e<false, kInl>(g, m / g.d, m % g.d, k, nullptr, nullptr); }
};
template <bool kInl>
struct LoadDagB {  // B(k,n) = e[m=k, j=n]
  static constexpr bool kContigK = false;
  GateCtx g;
  __device__ __forceinline__ float operator()(int k, int n) const { return gate_e<false, kInl>(g, k / g.d, k % g.d, n, nullptr, nullptr); }
};
template <bool kInl>
struct EpiDagDgrad {  // ebar[m,j] -> dx[b,j] += ebar*de/dx ; dP[i,j] += ebar*de/dP
  GateCtx g;
  float* dx;
  float* dP;
  __device__ __forceinline__ void operator()(int m, int n, const float* v, int nv) const {
    const int b = m / g.d, i = m % g.d;
    for (int jj = 0; jj < nv; ++jj) {
      float ddx, ddp;
      gate_e<true, kInl>(g, b, i, n + jj, &ddx, &ddp);
      atomicAdd(dx + (size_t)b * g.d + n + jj, v[jj] * ddx);
      atomicAdd(dP + (size_t)i * g.d + n + jj, v[jj] * ddp);
    }
  }
};

static int make_gate(GateCtx* out, const float* x, const float* P, const gnf_gate_t* gate, int d) {
  if (!gate) return fail(GNF_ERR_INVALID, "gate descriptor is NULL");
  if (gate->mode < GNF_GATE_TABLE || gate->mode > GNF_GATE_NOISER) return fail(GNF_ERR_UNSUPPORTED, "unknown gate mode %d", gate->mode);
  if (gate->mode == GNF_GATE_GUMBEL && !(gate->temperature > 0.f)) return fail(GNF_ERR_INVALID, "gumble_T must be > 0");
  if (gate->mode == GNF_GATE_GUMBEL && ((gate->noise1 == nullptr) != (gate->noise2 == nullptr)))
    return fail(GNF_ERR_INVALID, "Gumbel replay needs both noise tensors");
  out->x = x; out->P = P; out->n1 = gate->noise1; out->n2 = gate->noise2;
  out->seed = gate->seed; out->offset = gate->offset; out->offset_dev = gate->offset_dev;
  out->T = gate->temperature; out->mode = gate->mode; out->d = d;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Small kernels
// ------------------------------------------------------------------------------------------------
__global__ void zero2d_kernel(float* p, int ld, int rows, int cols) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[(i / cols) * ld + (i % cols)] = 0.f;
}
static void zero2d(float* p, int ld, int rows, int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return;
  if (ld == cols) { cudaMemsetAsync(p, 0, (size_t)rows * cols * sizeof(float), s); return; }
  const size_t n = (size_t)rows * cols;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  GNF_LAUNCH(zero2d_kernel, blocks, 256, 0, s, p, ld, rows, cols);
}

// out[c] += sum_{q in chunk} Y[q*rowstride + c]  for the [Q, C] view (C = period*ldy)
__global__ void colsum_kernel(const float* __restrict__ Y, float* __restrict__ out, int Q, int C, int ldy, int N, size_t rowstride, int q_per) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int q0 = blockIdx.y * q_per;
  const int q1 = (q0 + q_per < Q) ? q0 + q_per : Q;
  if (c < C && (c % ldy) < N) {
    float s = 0.f;
    for (int q = q0; q < q1; ++q) s += __ldg(Y + (size_t)q * rowstride + c);
    atomicAdd(out + (size_t)(c / ldy) * N + (c % ldy), s);
  }
}

__global__ void relu_mask_kernel(float* dY, int lddy, const float* __restrict__ act, int ldact, int M, int N) {
  const size_t n = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / N, c = i % N;
    if (!(act[m * ldact + c] > 0.f)) dY[m * lddy + c] = 0.f;
  }
}

__global__ void pack_rows_kernel(const float* __restrict__ W, const float* __restrict__ mask, const int32_t* __restrict__ perm, float* __restrict__ out, int R, int K) {
  const size_t n = (size_t)R * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K, k = i % K;
    const size_t src = (size_t)(perm ? perm[r] : (int)r) * K + k;
    out[i] = mask ? W[src] * mask[src] : W[src];
  }
}
__global__ void unpack_rows_kernel(const float* __restrict__ dWp, const float* __restrict__ mask, const int32_t* __restrict__ perm, float* __restrict__ dW, int R, int K) {
  const size_t n = (size_t)R * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K, k = i % K;
    const size_t dst = (size_t)(perm ? perm[r] : (int)r) * K + k;
    dW[dst] = mask ? dWp[i] * mask[dst] : dWp[i];
  }
}
__global__ void unpack_vec_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm, float* __restrict__ dst, int R) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) dst[perm ? perm[i] : i] = src[i];
}

__global__ void dag_importance_kernel(const float* __restrict__ A, int n, int imp, float h, float* __restrict__ P, float* __restrict__ dPdA) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float a = A[i];
    float p, dp;
    if (imp == GNF_IMP_RAW) {
      p = a; dp = 1.f;
    } else if (imp == GNF_IMP_HARD_SQ) {
      const float a2 = a * a;
      const float keep = (a2 > h) ? 1.f : 0.f;
      p = a2 * keep; dp = 2.f * a * keep;
    } else {
      const float s = 1.f / (1.f + expf(-2.f * (a * a)));   // sigmoid(2 A^2)
      p = 2.f * (s - .5f);
      dp = 8.f * a * s * (1.f - s);
      if (imp == GNF_IMP_HARD_SOFT) {
        const float keep = (p > h) ? 1.f : 0.f;
        p *= keep; dp *= keep;
      }
    }
    P[i] = p;
    if (dPdA) dPdA[i] = dp;
  }
}

__global__ void dag_bias_table_kernel(const float* __restrict__ W1, int ldw, const float* __restrict__ b1, float* __restrict__ T, int d, int N, int hot) {
  const int rows = hot ? d : 1;
  const size_t n = (size_t)rows * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / N), c = (int)(idx % N);
    T[idx] = (hot ? W1[(size_t)c * ldw + d + i] : 0.f) + (b1 ? b1[c] : 0.f);
  }
}
__global__ void dag_bias_table_bwd_kernel(const float* __restrict__ dT, float* __restrict__ dW1, int ldw, float* __restrict__ db1, int d, int N, int hot) {
  const int rows = hot ? d : 1;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < rows; ++i) {
      const float v = dT[(size_t)i * N + c];
      s += v;
      if (hot) dW1[(size_t)c * ldw + d + i] = v;
    }
    if (db1) db1[c] = s;
  }
}
__global__ void dag_finish_dA_kernel(const float* __restrict__ dP, const float* __restrict__ dPdA, float* __restrict__ dA, int n, int accumulate) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = dP[i] * dPdA[i];
    dA[i] = accumulate ? dA[i] + v : v;
  }
}
__global__ void dag_dump_noise_kernel(GateCtx g, float* n1, float* n2, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a, b;
    gate_noise(g, i, a, b);
    n1[i] = a;
    if (n2) n2[i] = b;
  }
}

static inline int ew_blocks(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > (size_t)8 * kNumSMs) b = (size_t)8 * kNumSMs;
  return b < 1 ? 1 : (int)b;
}
}  // namespace gnf

using namespace gnf;

extern "C" {

int gnf_linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* bias, int bias_period, float* Y,
                   int ldy, int M, int N, int K, int relu, gnf_stream_t stream) {
  if (!X || !W || !Y || M < 0 || N <= 0 || K <= 0 || ldx < K || ldw < K || ldy < N) return fail(GNF_ERR_INVALID, "gnf_linear_fwd: bad arguments");
  if (M == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  LoadRowMajorA al{X, ldx};
  LoadWeightT bl{W, ldw};
  EpiBiasAct epi{Y, ldy, bias, N, bias_period < 1 ? 1 : bias_period, relu};
  launch_gemm_auto(al, bl, epi, M, N, K, false, s);
  return check_launch("gnf_linear_fwd");
}

int gnf_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* act, int ldact, float* dX,
                     int lddx, int M, int N, int K, gnf_stream_t stream) {
  if (!dY || !W || !dX || M < 0 || N <= 0 || K <= 0 || lddy < N || ldw < K || lddx < K) return fail(GNF_ERR_INVALID, "gnf_linear_dgrad: bad arguments");
  if (M == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  LoadRowMajorA al{dY, lddy};        // A(m, n) reduction over n
  LoadRowMajorB bl{W, ldw};          // B(n, k) = W[n, k]
  EpiMaskStore epi{dX, lddx, act, ldact};
  launch_gemm_auto(al, bl, epi, M, K, N, false, s);
  return check_launch("gnf_linear_dgrad");
}

int gnf_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, int M, int N, int K,
                     gnf_stream_t stream) {
  if (!dY || !X || !dW || M < 0 || N <= 0 || K <= 0 || lddy < N || ldx < K || lddw < K) return fail(GNF_ERR_INVALID, "gnf_linear_wgrad: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  zero2d(dW, lddw, N, K, s);
  if (M == 0) return check_launch("gnf_linear_wgrad");
  LoadColMajorA al{dY, lddy};        // A(n, m) = dY[m, n]
  LoadRowMajorB bl{X, ldx};          // B(m, k) = X[m, k]
  EpiAtomicAdd epi{dW, lddw};
  launch_gemm_auto(al, bl, epi, N, K, M, true, s);
  return check_launch("gnf_linear_wgrad");
}

int gnf_colsum(const float* Y, int ldy, float* out, int M, int N, int period, gnf_stream_t stream) {
  if (!Y || !out || M < 0 || N <= 0 || period < 1 || ldy < N || (M % period) != 0) return fail(GNF_ERR_INVALID, "gnf_colsum: bad arguments (M must be a multiple of period)");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, (size_t)period * N * sizeof(float), s);
  if (M == 0) return check_launch("gnf_colsum");
  const int Q = M / period, C = period * ldy;
  const int cblocks = ceil_div(C, 256);
  int qsplit = ceil_div(4 * kNumSMs, cblocks);
  if (qsplit > ceil_div(Q, 8)) qsplit = ceil_div(Q, 8);
  if (qsplit < 1) qsplit = 1;
  const int q_per = ceil_div(Q, qsplit);
  GNF_LAUNCH(colsum_kernel, dim3(cblocks, ceil_div(Q, q_per)), 256, 0, s, Y, out, Q, C, ldy, N, (size_t)period * ldy, q_per);
  return check_launch("gnf_colsum");
}

int gnf_relu_mask(float* dY, int lddy, const float* act, int ldact, int M, int N, gnf_stream_t stream) {
  if (!dY || !act || M < 0 || N <= 0) return fail(GNF_ERR_INVALID, "gnf_relu_mask: bad arguments");
  if (M == 0) return 0;
  GNF_LAUNCH(relu_mask_kernel, ew_blocks((size_t)M * N), 256, 0, (cudaStream_t)stream, dY, lddy, act, ldact, M, N);
  return check_launch("gnf_relu_mask");
}

int gnf_pack_rows(const float* W, const float* mask, const int32_t* perm, float* out, int R, int K, gnf_stream_t stream) {
  if (!W || !out || R <= 0 || K <= 0) return fail(GNF_ERR_INVALID, "gnf_pack_rows: bad arguments");
  GNF_LAUNCH(pack_rows_kernel, ew_blocks((size_t)R * K), 256, 0, (cudaStream_t)stream, W, mask, perm, out, R, K);
  return check_launch("gnf_pack_rows");
}
int gnf_unpack_rows(const float* dWp, const float* mask, const int32_t* perm, float* dW, int R, int N, int K, gnf_stream_t stream) {
  if (!dWp || !dW || R <= 0 || K <= 0 || N < R) return fail(GNF_ERR_INVALID, "gnf_unpack_rows: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), s);
  GNF_LAUNCH(unpack_rows_kernel, ew_blocks((size_t)R * K), 256, 0, s, dWp, mask, perm, dW, R, K);
  return check_launch("gnf_unpack_rows");
}
int gnf_unpack_vec(const float* src, const int32_t* perm, float* dst, int R, int N, gnf_stream_t stream) {
  if (!src || !dst || R <= 0 || N < R) return fail(GNF_ERR_INVALID, "gnf_unpack_vec: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(dst, 0, (size_t)N * sizeof(float), s);
  GNF_LAUNCH(unpack_vec_kernel, ew_blocks((size_t)R), 256, 0, s, src, perm, dst, R);
  return check_launch("gnf_unpack_vec");
}

int gnf_dag_importance(const float* A, int d, int imp, float h_thresh, float* P, float* dPdA, gnf_stream_t stream) {
  if (!A || !P || d <= 0 || imp < GNF_IMP_RAW || imp > GNF_IMP_HARD_SQ) return fail(GNF_ERR_INVALID, "gnf_dag_importance: bad arguments");
  GNF_LAUNCH(dag_importance_kernel, ew_blocks((size_t)d * d), 256, 0, (cudaStream_t)stream, A, d * d, imp, h_thresh, P, dPdA);
  return check_launch("gnf_dag_importance");
}
int gnf_dag_bias_table(const float* W1, int ldw, const float* b1, float* T, int d, int N, int hot, gnf_stream_t stream) {
  if (!W1 || !T || d <= 0 || N <= 0 || ldw < (hot ? 2 * d : d)) return fail(GNF_ERR_INVALID, "gnf_dag_bias_table: bad arguments");
  GNF_LAUNCH(dag_bias_table_kernel, ew_blocks((size_t)(hot ? d : 1) * N), 256, 0, (cudaStream_t)stream, W1, ldw, b1, T, d, N, hot);
  return check_launch("gnf_dag_bias_table");
}
int gnf_dag_bias_table_bwd(const float* dT, float* dW1, int ldw, float* db1, int d, int N, int hot, gnf_stream_t stream) {
  if (!dT || (hot && !dW1) || d <= 0 || N <= 0) return fail(GNF_ERR_INVALID, "gnf_dag_bias_table_bwd: bad arguments");
  GNF_LAUNCH(dag_bias_table_bwd_kernel, ew_blocks((size_t)N), 256, 0, (cudaStream_t)stream, dT, dW1, ldw, db1, d, N, hot);
  return check_launch("gnf_dag_bias_table_bwd");
}

int gnf_dag_l1_fwd(const float* x, const float* P, const gnf_gate_t* gate, const float* W1, int ldw, const float* T,
                   int bias_period, float* Y, int ldy, int B, int d, int N, int relu, gnf_stream_t stream) {
  if (!x || !P || !W1 || !Y || B < 0 || d <= 0 || N <= 0 || ldw < d || ldy < N) return fail(GNF_ERR_INVALID, "gnf_dag_l1_fwd: bad arguments");
  GateCtx g;
  if (int e = make_gate(&g, x, P, gate, d)) return e;
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  LoadWeightT bl{W1, ldw};
  EpiBiasAct epi{Y, ldy, T, N, bias_period < 1 ? 1 : bias_period, relu};
  const int M = B * d;
  if (d >= kGateInlineMinD) launch_gemm_auto(LoadDagA<true>{g}, bl, epi, M, N, d, false, s);
  else launch_gemm_auto(LoadDagA<false>{g}, bl, epi, M, N, d, false, s);
  return check_launch("gnf_dag_l1_fwd");
}

int gnf_dag_l1_wgrad(const float* dY, int lddy, const float* x, const float* P, const gnf_gate_t* gate, float* dW1,
                     int ldw, int B, int d, int N, gnf_stream_t stream) {
  if (!dY || !x || !P || !dW1 || B < 0 || d <= 0 || N <= 0 || ldw < d || lddy < N) return fail(GNF_ERR_INVALID, "gnf_dag_l1_wgrad: bad arguments");
  GateCtx g;
  if (int e = make_gate(&g, x, P, gate, d)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  zero2d(dW1, ldw, N, d, s);
  if (B == 0) return check_launch("gnf_dag_l1_wgrad");
  const int M = B * d;
  LoadColMajorA al{dY, lddy};   // A(n, m) = dY[m, n]
  EpiAtomicAdd epi{dW1, ldw};
  if (d >= kGateInlineMinD) launch_gemm_auto(al, LoadDagB<true>{g}, epi, N, d, M, true, s);    // B(m, j) = e[m, j]
  else launch_gemm_auto(al, LoadDagB<false>{g}, epi, N, d, M, true, s);
  return check_launch("gnf_dag_l1_wgrad");
}

int gnf_dag_l1_dgrad(const float* dY, int lddy, const float* W1, int ldw, const float* x, const float* P,
                     const gnf_gate_t* gate, float* dx, float* dP, int B, int d, int N, gnf_stream_t stream) {
  if (!dY || !W1 || !x || !P || !dx || !dP || B < 0 || d <= 0 || N <= 0 || ldw < d || lddy < N) return fail(GNF_ERR_INVALID, "gnf_dag_l1_dgrad: bad arguments");
  GateCtx g;
  if (int e = make_gate(&g, x, P, gate, d)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(dx, 0, (size_t)B * d * sizeof(float), s);
  cudaMemsetAsync(dP, 0, (size_t)d * d * sizeof(float), s);
  if (B == 0) return check_launch("gnf_dag_l1_dgrad");
  const int M = B * d;
  LoadRowMajorA al{dY, lddy};   // A(m, n)
  LoadRowMajorB bl{W1, ldw};    // B(n, j) = W1[n, j]
  // N = d <= 64 gives one column of tiles (50 CTAs at cfg4): split the reduction over the layer width.  The epilogue is
  // linear in the accumulator and already reduces with atomics, so partial sums need no second pass.
  if (d >= kGateInlineMinD) launch_gemm_auto(al, bl, EpiDagDgrad<true>{g, dx, dP}, M, d, N, true, s);
  else launch_gemm_auto(al, bl, EpiDagDgrad<false>{g, dx, dP}, M, d, N, true, s);
  return check_launch("gnf_dag_l1_dgrad");
}

int gnf_dag_finish_dA(const float* dP, const float* dPdA, float* dA, int d, int accumulate, gnf_stream_t stream) {
  if (!dP || !dPdA || !dA || d <= 0) return fail(GNF_ERR_INVALID, "gnf_dag_finish_dA: bad arguments");
  GNF_LAUNCH(dag_finish_dA_kernel, ew_blocks((size_t)d * d), 256, 0, (cudaStream_t)stream, dP, dPdA, dA, d * d, accumulate);
  return check_launch("gnf_dag_finish_dA");
}

int gnf_dag_dump_noise(const gnf_gate_t* gate, float* n1, float* n2, int B, int d, gnf_stream_t stream) {
  if (!gate || !n1 || B <= 0 || d <= 0) return fail(GNF_ERR_INVALID, "gnf_dag_dump_noise: bad arguments");
  GateCtx g;
  gnf_gate_t tmp = *gate;
  tmp.noise1 = tmp.noise2 = nullptr;
  if (int e = make_gate(&g, nullptr, nullptr, &tmp, d)) return e;
  const size_t n = (size_t)B * d * d;
  GNF_LAUNCH(dag_dump_noise_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, g, n1, n2, n);
  return check_launch("gnf_dag_dump_noise");
}

}  // extern "C"
