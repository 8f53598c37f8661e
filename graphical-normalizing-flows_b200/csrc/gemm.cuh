// fp32 FFMA tile-GEMM engine with functor operand loaders and functor epilogues.
//
//   C(m,n) = sum_k A(m,k) * B(k,n)
//
// A(m,k) and B(k,n) are produced by loader functors (plain strided memory, transposed memory, or
// the DAG masked-embedding generator), so an operand such as the [B*d, d] masked input never has to
// exist in memory.  The epilogue functor receives 4 consecutive n per call (bias/ReLU store,
// ReLU-masked store, split-K atomic accumulate, fused dx/dP reduction, ...).
//
// This is the strict-fp32 path (bit-level comparable with an fp32 reference up to summation order).
#pragma once
#include "common.cuh"

namespace gnf {

template <int BM_, int BN_, int BK_, int TM_, int TN_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int THREADS = (BM / TM) * (BN / TN);
  static constexpr int PAD = 4;
  static constexpr int LDA = BM + PAD, LDB = BN + PAD;
  static constexpr int SMEM_FLOATS = 2 * BK * (LDA + LDB);
  static constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float);
  static_assert(TM % 4 == 0 && TN % 4 == 0, "thread tile must be a multiple of 4");
  static_assert((BM * BK) % THREADS == 0 && (BN * BK) % THREADS == 0, "loader mapping");
};

using TileBig = TileCfg<128, 128, 16, 8, 8>;   // 256 threads, 8x8 per thread
using TileMid = TileCfg<128, 64, 16, 8, 4>;    // 256 threads, 8x4 per thread (32 < N <= 64)
using TileSkinny = TileCfg<128, 32, 16, 4, 4>; // 256 threads, 4x4 per thread (N <= 32)
using TileTiny = TileCfg<32, 32, 64, 4, 4>;    // 64 threads, 64-deep k tiles (the loop is one global round trip per k tile): N <= 32 and too few 128-row tiles to fill the GPU (the 630 -> 30 output layer)
using TileShort = TileCfg<32, 128, 16, 4, 8>;  // 128 threads: M <= 32 with a wide output (the output layer's wgrad: 30 x 630)

template <class Cfg, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS) gemm_kernel(ALoad al, BLoad bl, Epi epi, int M, int N, int K, int k_per_split) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN;
  constexpr int THREADS = Cfg::THREADS, LDA = Cfg::LDA, LDB = Cfg::LDB;
  constexpr int A_PER = BM * BK / THREADS, B_PER = BN * BK / THREADS;
  constexpr int RCH = TM / 4, CCH = TN / 4;
  GNF_SMEM(float, smem);
  float* As = smem;                 // [2][BK][LDA]
  float* Bs = smem + 2 * BK * LDA;  // [2][BK][LDB]

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const int ntiles = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  const int tx = t % (BN / TN), ty = t / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[A_PER], rb[B_PER];
  auto load_regs = [&](int tile) {
    const int k0 = kbeg + tile * BK;
#pragma unroll
    for (int e = 0; e < A_PER; ++e) {
      const int idx = t + e * THREADS;
      int mm, kk;
      if (ALoad::kContigK) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
      const int k = k0 + kk;
      ra[e] = (m0 + mm < M && k < kend) ? al(m0 + mm, k) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < B_PER; ++e) {
      const int idx = t + e * THREADS;
      int nn, kk;
      if (BLoad::kContigK) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
      const int k = k0 + kk;
      rb[e] = (n0 + nn < N && k < kend) ? bl(k, n0 + nn) : 0.f;
    }
  };
  auto store_smem = [&](int buf) {
    float* as = As + buf * BK * LDA;
    float* bs = Bs + buf * BK * LDB;
#pragma unroll
    for (int e = 0; e < A_PER; ++e) {
      const int idx = t + e * THREADS;
      int mm, kk;
      if (ALoad::kContigK) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
      as[kk * LDA + mm] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < B_PER; ++e) {
      const int idx = t + e * THREADS;
      int nn, kk;
      if (BLoad::kContigK) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
      bs[kk * LDB + nn] = rb[e];
    }
  };

  if (ntiles > 0) {
    load_regs(0);
    store_smem(0);
  }
  __syncthreads();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_regs(tile + 1);
    const float* as = As + buf * BK * LDA;
    const float* bs = Bs + buf * BK * LDB;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int rc = 0; rc < RCH; ++rc) {
        const float4 v = *reinterpret_cast<const float4*>(&as[kk * LDA + rc * (BM / RCH) + ty * 4]);
        a[rc * 4 + 0] = v.x; a[rc * 4 + 1] = v.y; a[rc * 4 + 2] = v.z; a[rc * 4 + 3] = v.w;
      }
#pragma unroll
      for (int cc = 0; cc < CCH; ++cc) {
        const float4 v = *reinterpret_cast<const float4*>(&bs[kk * LDB + cc * (BN / CCH) + tx * 4]);
        b[cc * 4 + 0] = v.x; b[cc * 4 + 1] = v.y; b[cc * 4 + 2] = v.z; b[cc * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (tile + 1 < ntiles) store_smem(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int rc = 0; rc < RCH; ++rc)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + rc * (BM / RCH) + ty * 4 + i;
      if (m >= M) continue;
#pragma unroll
      for (int cc = 0; cc < CCH; ++cc) {
        const int n = n0 + cc * (BN / CCH) + tx * 4;
        if (n >= N) continue;
        float v[4] = {acc[rc * 4 + i][cc * 4 + 0], acc[rc * 4 + i][cc * 4 + 1], acc[rc * 4 + i][cc * 4 + 2],
                      acc[rc * 4 + i][cc * 4 + 3]};
        epi(m, n, v, (N - n < 4) ? N - n : 4);
      }
    }
}

// ------------------------------------------------------------------------------------------------
// Operand loaders
// ------------------------------------------------------------------------------------------------
struct LoadRowMajorA {  // A(m,k) = p[m*ld + k]
  static constexpr bool kContigK = true;
  const float* p; int ld;
  __device__ __forceinline__ float operator()(int m, int k) const { return __ldg(p + (size_t)m * ld + k); }
};
struct LoadColMajorA {  // A(m,k) = p[k*ld + m]
  static constexpr bool kContigK = false;
  const float* p; int ld;
  __device__ __forceinline__ float operator()(int m, int k) const { return __ldg(p + (size_t)k * ld + m); }
};
struct LoadWeightT {    // B(k,n) = W[n*ld + k]   (nn.Linear weight used as x @ W^T)
  static constexpr bool kContigK = true;
  const float* p; int ld;
  __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(p + (size_t)n * ld + k); }
};
struct LoadRowMajorB {  // B(k,n) = p[k*ld + n]
  static constexpr bool kContigK = false;
  const float* p; int ld;
  __device__ __forceinline__ float operator()(int k, int n) const { return __ldg(p + (size_t)k * ld + n); }
};

// ------------------------------------------------------------------------------------------------
// Epilogues
// ------------------------------------------------------------------------------------------------
struct EpiBiasAct {  // Y = act(acc + bias[(m % period), n])
  float* Y; int ldy; const float* bias; int bias_ld; int period; int relu;
  __device__ __forceinline__ void operator()(int m, int n, const float* v, int nv) const {
    float* y = Y + (size_t)m * ldy + n;
    const float* bp = bias ? bias + (size_t)(period > 1 ? (m % period) : 0) * bias_ld + n : nullptr;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r = v[j];
      if (bp && j < nv) r += __ldg(bp + j);
      if (relu) r = fmaxf(r, 0.f);
      o[j] = r;
    }
    if (nv == 4 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
      *reinterpret_cast<float4*>(y) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int j = 0; j < nv; ++j) y[j] = o[j];
    }
  }
};
struct EpiMaskStore {  // dX = acc * (act > 0)
  float* dX; int ld; const float* act; int ldact;
  __device__ __forceinline__ void operator()(int m, int n, const float* v, int nv) const {
    for (int j = 0; j < nv; ++j) {
      float r = v[j];
      if (act && !(__ldg(act + (size_t)m * ldact + n + j) > 0.f)) r = 0.f;
      dX[(size_t)m * ld + n + j] = r;
    }
  }
};
struct EpiAtomicAdd {  // C += acc   (split-K)
  float* C; int ld;
  __device__ __forceinline__ void operator()(int m, int n, const float* v, int nv) const {
    for (int j = 0; j < nv; ++j) atomicAdd(C + (size_t)m * ld + n + j, v[j]);
  }
};
struct EpiStoreSplit {  // P[split][m, n] = acc   (split-K partial tiles, summed in fixed order by a second kernel: deterministic)
  float* P; int ld; size_t split_stride;
  __device__ __forceinline__ void operator()(int m, int n, const float* v, int nv) const {
    float* dst = P + (size_t)blockIdx.z * split_stride + (size_t)m * ld + n;
    for (int j = 0; j < nv; ++j) dst[j] = v[j];
  }
};
struct EpiStore {  // C = acc
  float* C; int ld;
  __device__ __forceinline__ void operator()(int m, int n, const float* v, int nv) const {
    for (int j = 0; j < nv; ++j) C[(size_t)m * ld + n + j] = v[j];
  }
};

template <class Cfg, class ALoad, class BLoad, class Epi>
static inline void launch_gemm(const ALoad& al, const BLoad& bl, const Epi& epi, int M, int N, int K, int splits, cudaStream_t stream);

// Pick the tile by output width: 128x128 (N > 64), 128x64 (32 < N <= 64), 128x32 (N <= 32).
template <class ALoad, class BLoad, class Epi>
static inline void launch_gemm_auto(const ALoad& al, const BLoad& bl, const Epi& epi, int M, int N, int K, bool split_k, cudaStream_t stream) {
  auto splits = [&](int bm, int bn) {
    if (!split_k) return 1;
    const int tiles = ceil_div(M, bm) * ceil_div(N, bn);
    int want = ceil_div(2 * kNumSMs, tiles);
    const int maxs = ceil_div(K, 4 * 16);
    if (want > maxs) want = maxs;
    return want < 1 ? 1 : want;
  };
  if (N <= 32 && !split_k && ceil_div(M, 128) < kNumSMs) launch_gemm<TileTiny>(al, bl, epi, M, N, K, 1, stream);
  else if (N <= 32) launch_gemm<TileSkinny>(al, bl, epi, M, N, K, splits(128, 32), stream);
  else if (M <= 32 && N > 64) launch_gemm<TileShort>(al, bl, epi, M, N, K, splits(32, 128), stream);
  else if (N <= 64) launch_gemm<TileMid>(al, bl, epi, M, N, K, splits(128, 64), stream);
  else launch_gemm<TileBig>(al, bl, epi, M, N, K, splits(128, 128), stream);
}

template <class Cfg, class ALoad, class BLoad, class Epi>
static inline void launch_gemm(const ALoad& al, const BLoad& bl, const Epi& epi, int M, int N, int K, int splits,
                               cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  if (splits < 1) splits = 1;
  int ktiles = ceil_div(K > 0 ? K : 1, Cfg::BK);
  int k_per_split = ceil_div(ktiles, splits) * Cfg::BK;
  splits = ceil_div(K > 0 ? K : 1, k_per_split);
  dim3 grid(ceil_div(M, Cfg::BM), ceil_div(N, Cfg::BN), splits);
#ifndef GNF_EMU
  if (Cfg::SMEM_BYTES > 48 * 1024)         // per launch: the attribute is per device (a process-wide flag breaks a second GPU)
    cudaFuncSetAttribute(gemm_kernel<Cfg, ALoad, BLoad, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)Cfg::SMEM_BYTES);
#endif
  GNF_LAUNCH((gemm_kernel<Cfg, ALoad, BLoad, Epi>), grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, al, bl, epi, M, N,
             K, k_per_split);
}

}  // namespace gnf
