// Multi-tensor Adam step: every parameter tensor of the flow in ONE launch (torch's fused Adam walks the tensors in 64 K-element
// chunks -- ~20 CTAs for the 0.95 M parameters of the BSDS300-shape flow, 39 us per step; this kernel spreads the same 26 MB of
// traffic over the whole chip).  Arithmetic = torch.optim.Adam (the optimizer of the reference's drivers, UCIExperiments.py:100,
// ToyExperiments.py:59: L2 weight decay folded into the gradient, no amsgrad), evaluated per element in fp32 with the bias
// corrections 1 - beta^t computed in double from a DEVICE step counter, so that the step replays inside a CUDA graph.
#include "common.cuh"

namespace gnf {

constexpr int kAdamMaxTensors = 48;       // per launch (kernel parameter space); more tensors = more launches
constexpr int kAdamChunk = 4096;          // elements per block
constexpr int kAdamThreads = 256;

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long n[kAdamMaxTensors];
  int first_chunk[kAdamMaxTensors + 1];   // block index of the tensor's first chunk
  int count;
};

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(AdamTable tb, const long long* __restrict__ step, float lr, float beta1,
                                                                  float beta2, float eps, float wd) {
  GNF_SMEM(float, s_corr);                 // [2] bias corrections, computed once per block in double
  if (threadIdx.x == 0) {
    const double t = (double)(*step + 1);
    s_corr[0] = (float)(1.0 - pow((double)beta1, t));
    s_corr[1] = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  __syncthreads();
  int ti = 0;
  while (ti + 1 < tb.count && (int)blockIdx.x >= tb.first_chunk[ti + 1]) ++ti;
  const long long base = (long long)((int)blockIdx.x - tb.first_chunk[ti]) * kAdamChunk;
  const long long n = tb.n[ti];
  float* __restrict__ p = tb.p[ti];
  const float* __restrict__ g = tb.g[ti];
  float* __restrict__ m = tb.m[ti];
  float* __restrict__ v = tb.v[ti];
  const float step_size = lr / s_corr[0], inv_sqrt_bc2 = 1.f / s_corr[1];
  auto update = [&](float pv, float gv0, float& mv, float& vv) {
    const float gv = fmaf(wd, pv, gv0);
    mv = mv + (1.f - beta1) * (gv - mv);                                 // torch: exp_avg.lerp_(grad, 1 - beta1)
    vv = beta2 * vv + (1.f - beta2) * gv * gv;                           // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    return pv - step_size * (mv / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  };
  const long long end = (n < base + kAdamChunk) ? n : base + kAdamChunk;
#ifndef GNF_EMU
  // 16-byte path: the whole chunk of a thread (4 x float4 of each of p, g, m, v) is in flight before the first update -- with scalar
  // loads the 16 trips per thread were 16 serial round trips (17 us for 26 MB)
  if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) {
    constexpr int kPer = kAdamChunk / (4 * kAdamThreads);                // float4 per thread
    float4 pv[kPer], gv[kPer], mv[kPer], vv[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const long long i = base + 4 * (threadIdx.x + (long long)u * kAdamThreads);
      if (i + 4 <= end) {
        pv[u] = *reinterpret_cast<const float4*>(p + i); gv[u] = *reinterpret_cast<const float4*>(g + i);
        mv[u] = *reinterpret_cast<const float4*>(m + i); vv[u] = *reinterpret_cast<const float4*>(v + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const long long i = base + 4 * (threadIdx.x + (long long)u * kAdamThreads);
      if (i + 4 <= end) {
        pv[u].x = update(pv[u].x, gv[u].x, mv[u].x, vv[u].x); pv[u].y = update(pv[u].y, gv[u].y, mv[u].y, vv[u].y);
        pv[u].z = update(pv[u].z, gv[u].z, mv[u].z, vv[u].z); pv[u].w = update(pv[u].w, gv[u].w, mv[u].w, vv[u].w);
        *reinterpret_cast<float4*>(m + i) = mv[u]; *reinterpret_cast<float4*>(v + i) = vv[u]; *reinterpret_cast<float4*>(p + i) = pv[u];
      } else {
        for (long long e = i; e < end; ++e) {                            // ragged tail of the tensor (at most 3 elements, one thread)
          float me = m[e], ve = v[e];
          p[e] = update(p[e], g[e], me, ve);
          m[e] = me; v[e] = ve;
        }
      }
    }
    return;
  }
#endif
  for (long long i = base + threadIdx.x; i < end; i += kAdamThreads) {
    float mv = m[i], vv = v[i];
    p[i] = update(p[i], g[i], mv, vv);
    m[i] = mv;
    v[i] = vv;
  }
}

}  // namespace gnf

using namespace gnf;

extern "C" {

int gnf_adam_step(const gnf_adam_tensor_t* tensors, int n_tensors, const int64_t* step_dev, float lr, float beta1, float beta2, float eps,
                  float weight_decay, gnf_stream_t stream) {
  if (!tensors || n_tensors < 0 || !step_dev) return fail(GNF_ERR_INVALID, "gnf_adam_step: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  for (int t0 = 0; t0 < n_tensors; t0 += kAdamMaxTensors) {
    AdamTable tb;
    tb.count = n_tensors - t0 < kAdamMaxTensors ? n_tensors - t0 : kAdamMaxTensors;
    int blocks = 0;
    for (int i = 0; i < tb.count; ++i) {
      const gnf_adam_tensor_t& a = tensors[t0 + i];
      if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq || a.numel < 0) return fail(GNF_ERR_INVALID, "gnf_adam_step: tensor %d has a NULL pointer", t0 + i);
      tb.p[i] = a.param; tb.g[i] = a.grad; tb.m[i] = a.exp_avg; tb.v[i] = a.exp_avg_sq; tb.n[i] = a.numel;
      tb.first_chunk[i] = blocks;
      blocks += (int)((a.numel + kAdamChunk - 1) / kAdamChunk);
    }
    tb.first_chunk[tb.count] = blocks;
    for (int i = tb.count; i < kAdamMaxTensors; ++i) { tb.p[i] = nullptr; tb.g[i] = nullptr; tb.m[i] = nullptr; tb.v[i] = nullptr; tb.n[i] = 0; tb.first_chunk[i + 1] = blocks; }
    if (blocks == 0) continue;
    GNF_LAUNCH(adam_step_kernel, blocks, kAdamThreads, 2 * sizeof(float), s, tb, reinterpret_cast<const long long*>(step_dev), lr, beta1, beta2, eps, weight_decay);
  }
  return check_launch("gnf_adam_step");
}

}  // extern "C"
