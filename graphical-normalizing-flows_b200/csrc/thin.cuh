// Skinny layers of the hot path whose REDUCTION is a handful of terms: the input cotangent of a conditioner output layer with few
// outputs (cfg5: 1024 -> 2, dX[78400, 1024] = dY[78400, 2] W) is a pure streaming pass -- write 321 MB, read the ReLU mask -- that
// the 128 x 128 register-tiled GEMM runs at 294 us; with the whole weight matrix in shared memory and the mask loads issued ahead
// of the contraction it takes 122 us (profiles/r02y_thin_bench.txt).  Measured and NOT kept: the same idea for reductions of ~30
// terms and for thin outputs (cfg4's 630 -> 30 layer): 12 - 38 us against 8 - 30 us of the tile GEMM.
#pragma once
#include "common.cuh"

namespace gnf {

constexpr int kThinT = 8;                  // the thin (reduction) dimension at most
constexpr int kThinThreads = 256;
constexpr int kThinMaxWide = 4096;         // wide dimension: T x wide floats of shared memory

// Weight image loader: dst[a * lda_dst + b] = Wp[a * s_a + b * s_b] for a < A, b < Bn (zero beyond, up to Ap x Bp), with coalesced
// global reads whichever of the two source strides is 1: blocks of 32 x 32 go through a per-warp 32 x 33 transposition tile
// when the source is contiguous along a.  All threads of the CTA call it; the caller synchronises afterwards.
__device__ __forceinline__ void thin_load_weights(float* __restrict__ dst, int lda_dst, const float* __restrict__ Wp, long long s_a, long long s_b,
                                                  int A, int Bn, int Ap, int Bp, float* __restrict__ tile) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // every load is unconditional, from a clamped address, and eight (or 32) of them are in flight per thread: the image is loaded
  // by every CTA before it can start, a dependent-load chain here costs more than the layer itself
  if (s_a == 1 && s_b != 1) {
    const int ba = (Ap + 31) / 32, bb = (Bp + 31) / 32;
    for (int blk = warp; blk < ba * bb; blk += kThinThreads / 32) {
      const int a0 = (blk % ba) * 32, b0 = (blk / ba) * 32;
      const int a = a0 + lane, ac = a < A ? a : A - 1;
#pragma unroll
      for (int r = 0; r < 32; ++r) {                           // row b0 + r of the source, 32 consecutive a
        const int b = b0 + r, bc = b < Bn ? b : Bn - 1;
        const float v = __ldg(Wp + ac + bc * s_b);
        tile[r * 33 + lane] = (a < A && b < Bn) ? v : 0.f;
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; ++r) {                           // dst row a0 + r, 32 consecutive b
        const int aa = a0 + r, b = b0 + lane;
        if (aa < Ap && b < Bp) dst[(size_t)aa * lda_dst + b] = tile[lane * 33 + r];
      }
      __syncwarp();
    }
  } else {
    const int total = Ap * Bp;
    for (int base = 0; base < total; base += kThinThreads * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * kThinThreads + tid;
        const int ia = (idx < total ? idx : total - 1) / Bp, ib = (idx < total ? idx : total - 1) - ia * Bp;
        const float x = __ldg(Wp + (ia < A ? ia : A - 1) * s_a + (ib < Bn ? ib : Bn - 1) * s_b);
        v[u] = (ia < A && ib < Bn) ? x : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * kThinThreads + tid;
        if (idx < total) { const int ia = idx / Bp; dst[(size_t)ia * lda_dst + (idx - ia * Bp)] = v[u]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Thin REDUCTION: out[m, w] = epi( sum_t in[m, t] * Wt(t, w) ), t < T <= 32, w < Wd.  Lane = output column (coalesced stores),
// a warp computes kThinRows rows at a time so that every weight read from shared memory feeds kThinRows FMAs; the rows' T inputs
// sit in a small per-warp tile and are read as broadcasts.  epi: + bias[w], ReLU, or the dgrad mask (act[m, w] > 0).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kThinRows = 4;
constexpr int kThinSlots = 8;              // output columns per lane and pass: 256 columns per pass over the wide dimension

__global__ void __launch_bounds__(kThinThreads) thin_red_kernel(const float* __restrict__ in, int ldin, const float* __restrict__ Wp, long long s_t,
                                                                long long s_w, const float* __restrict__ bias, int relu, const float* __restrict__ act,
                                                                int ldact, float* __restrict__ out, int ldout, int M, int T, int Wd) {
  GNF_SMEM(float, smem);
  const int WdP = (Wd + 31) / 32 * 32;
  float* Ws = smem;                                          // [T][WdP]: Ws[t*WdP + w], zero for w >= Wd
  float* xin = Ws + (size_t)T * WdP;                         // [8 warps][32 x 33]: transposition tile of the weight loader, then the rows' inputs
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  thin_load_weights(Ws, WdP, Wp, s_t, s_w, T, Wd, T, WdP, xin + warp * (32 * 33));
  __syncthreads();
  float* xr = xin + warp * (32 * 33);
  const int nwarps = gridDim.x * (kThinThreads / 32);
  for (int m0 = (blockIdx.x * (kThinThreads / 32) + warp) * kThinRows; m0 < M; m0 += nwarps * kThinRows) {
#pragma unroll
    for (int r = 0; r < kThinRows; ++r) {
      const int mr = m0 + r;
      const float v = __ldg(in + (size_t)(mr < M ? mr : M - 1) * ldin + (lane < T ? lane : T - 1));
      xr[r * 32 + lane] = (mr < M && lane < T) ? v : 0.f;
    }
    __syncwarp();
    for (int wb = 0; wb < WdP; wb += 32 * kThinSlots) {
      // the dgrad mask of the pass first: its loads fly under the contraction
      float keep[kThinRows][kThinSlots];
      if (act) {
#pragma unroll
        for (int r = 0; r < kThinRows; ++r)
#pragma unroll
          for (int j = 0; j < kThinSlots; ++j) {
            const int mr = m0 + r, w = wb + 32 * j + lane;
            keep[r][j] = __ldg(act + (size_t)(mr < M ? mr : M - 1) * ldact + (w < Wd ? w : Wd - 1));
          }
      }
      float acc[kThinRows][kThinSlots];
#pragma unroll
      for (int r = 0; r < kThinRows; ++r)
#pragma unroll
        for (int j = 0; j < kThinSlots; ++j) acc[r][j] = 0.f;
      for (int t = 0; t < T; ++t) {
        float xv[kThinRows];
#pragma unroll
        for (int r = 0; r < kThinRows; ++r) xv[r] = xr[r * 32 + t];
        const float* wrow = Ws + (size_t)t * WdP + wb + lane;
#pragma unroll
        for (int j = 0; j < kThinSlots; ++j) {
          const float wv = (wb + 32 * j < WdP) ? wrow[32 * j] : 0.f;
#pragma unroll
          for (int r = 0; r < kThinRows; ++r) acc[r][j] = fmaf(xv[r], wv, acc[r][j]);
        }
      }
#pragma unroll
      for (int j = 0; j < kThinSlots; ++j) {
        const int w = wb + 32 * j + lane;
        if (w < Wd) {
          const float bv = bias ? __ldg(bias + w) : 0.f;
#pragma unroll
          for (int r = 0; r < kThinRows; ++r) {
            const int mr = m0 + r;
            if (mr < M) {
              float v = acc[r][j] + bv;
              if (relu) v = fmaxf(v, 0.f);
              if (act && !(keep[r][j] > 0.f)) v = 0.f;
              out[(size_t)mr * ldout + w] = v;
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

static inline size_t thin_red_smem(int T, int Wd) { return ((size_t)T * ((Wd + 31) / 32 * 32) + (size_t)(kThinThreads / 32) * 32 * 33) * sizeof(float); }

// CTAs per SM that the shared-memory footprint allows (227 KB per SM, 1 KB of bookkeeping per CTA), at most `cap`
static inline int thin_per_sm(size_t smem, int cap) {
  const int fit = (int)((227 * 1024) / (smem + 1024));
  return fit < 1 ? 1 : (fit > cap ? cap : fit);
}

}  // namespace gnf
