// K3 on the 5th-generation tensor cores: fused Clenshaw-Curtis UMNN integral, forward, "fast" mode
// (single-pass TF32 operands, fp32 accumulation in TMEM; per-sample log-likelihood tolerance 2e-3).
//
// Design (one CTA per SM, persistent over 128-node-row tiles):
//   * every layer's weight matrix is packed once per call into the UMMA canonical K-major shared-memory
//     image and pulled into shared memory by bulk async copies (TMA unit, mbarrier completion); it stays
//     resident for the CTA's lifetime -- B operand of every tcgen05.mma;
//   * the activation chain never touches shared or global memory: thread t owns tile row t = TMEM lane t;
//     the layer input lives in TMEM columns [0, NP) (A operand, TS form), the accumulator in columns
//     [NP, 2NP); the epilogue (tcgen05.ld -> bias + ReLU in registers -> tcgen05.st) rewrites the A region
//     for the next layer;
//   * the last Linear (N = 1), ELU + 1.05, the CC weighting and the per-row segment reduction are done in
//     registers / a 128-float shared buffer, emitting z, jac, logdet (+ zrev) directly.
#include "tc_common.cuh"

#ifndef GNF_EMU
namespace gnf {

constexpr int kTcRows = 128;

struct TcPlan {
  int L, NP, alloc_cols;
  int kp[GNF_MAX_LAYERS];          // K of layer l rounded up to 8 (number of A columns consumed)
  size_t off[GNF_MAX_LAYERS];      // float offset of layer l's image
  size_t off_bias, off_wlast, total_floats;
};

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static int make_tc_plan(const gnf_mlp_t* net, TcPlan* pl) {
  if (!net || net->n_layers < 2 || net->n_layers > GNF_MAX_LAYERS) return fail(GNF_ERR_UNSUPPORTED, "umnn tc: integrand needs 2..%d linear layers", GNF_MAX_LAYERS);
  if (net->dims[net->n_layers] != 1) return fail(GNF_ERR_UNSUPPORTED, "umnn tc: integrand output size must be 1");
  const int L = net->n_layers - 1;
  int maxh = 0;
  for (int l = 1; l <= L; ++l) maxh = net->dims[l] > maxh ? net->dims[l] : maxh;
  const int NP = round_up(maxh, 16);
  if (NP > 256 || round_up(net->dims[0], 8) > NP) return fail(GNF_ERR_UNSUPPORTED, "umnn tc: widths out of range (hidden max %d, input %d)", maxh, net->dims[0]);
  size_t off = 0;
  for (int l = 0; l < L; ++l) {
    pl->kp[l] = round_up(net->dims[l], 8);
    pl->off[l] = off;
    off += (size_t)pl->kp[l] * NP;
  }
  pl->off_bias = off; off += (size_t)L * NP;
  pl->off_wlast = off; off += NP;
  pl->total_floats = off;
  pl->L = L; pl->NP = NP;
  int a = 32;
  while (a < 2 * NP) a <<= 1;
  pl->alloc_cols = a;
  const size_t smem = off * sizeof(float) + 1024;
  if (smem > 227 * 1024) return fail(GNF_ERR_UNSUPPORTED, "umnn tc: resident weights need %zu B of shared memory (> 227 KB); use the strict kernel", smem);
  return 0;
}

struct TcPackArgs {
  const float* W[GNF_MAX_LAYERS];
  const float* b[GNF_MAX_LAYERS];
  size_t off[GNF_MAX_LAYERS], off_bias, off_wlast;
  int dims[GNF_MAX_LAYERS + 1], kp[GNF_MAX_LAYERS];
  int L, NP;
};

// Canonical K-major / no-swizzle image of layer l:  img[(k/4)][n][k%4] = W[n][k]  (zero padded)
__global__ void umnn_tc_pack_kernel(TcPackArgs a, float* __restrict__ ws, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i >= a.off_wlast) {
      const int k = (int)(i - a.off_wlast);
      if (k < a.dims[a.L]) v = a.W[a.L][k];
    } else if (i >= a.off_bias) {
      const int e = (int)(i - a.off_bias);
      const int l = e / a.NP, n = e % a.NP;
      if (n < a.dims[l + 1]) v = a.b[l][n];
    } else {
      int l = 0;
      while (l + 1 < a.L && i >= a.off[l + 1]) ++l;
      const size_t e = i - a.off[l];
      const int kc = (int)(e / ((size_t)a.NP * 4)), n = (int)((e / 4) % a.NP), kk = (int)(e % 4);
      const int k = kc * 4 + kk;
      if (n < a.dims[l + 1] && k < a.dims[l]) v = a.W[l][(size_t)n * a.dims[l] + k];
    }
    ws[i] = v;
  }
}

struct TcFwdParams {
  const float* x; const float* h; const float* ccw; const float* ccn; const float* image; const float* blast;
  float* z; float* zrev; float* jac; float* logdet;
  int R, d, E, S, L, alloc_cols;
  long long Q;
  int kp[GNF_MAX_LAYERS];
  unsigned off[GNF_MAX_LAYERS];     // float offsets inside the image
  unsigned off_bias, off_wlast, total_floats;
  long long* trace;                 // debug: per-phase SM clock stamps of CTA 0 (NULL in production)
};

static long long* g_tc_trace = nullptr;

template <int NP>
__global__ void __launch_bounds__(kTcRows) umnn_fwd_tc_kernel(TcFwdParams p) {
  using namespace tc;
  GNF_SMEM(float, smem);
  float* img = smem;                                                        // weight images + bias + wlast
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ((p.total_floats + 3) / 4) * 4);   // [0]: weights landed, [1]: MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* red = reinterpret_cast<float*>(bars + 4);                          // [128]
  const int t = threadIdx.x, warp = t >> 5;
  const int nodes = p.S + 1;

  if (warp == 0) tmem_alloc(tmem_slot, p.alloc_cols);
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const uint32_t tA = tmem_base + lane_sel;             // activations: columns [0, NP)
  const uint32_t tD = tmem_base + lane_sel + NP;        // accumulator: columns [NP, 2NP)

  if (t == 0) {
    const uint32_t bytes = p.total_floats * 4u;
    mbar_expect_tx(&bars[0], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t n = (bytes - o < 32768u) ? bytes - o : 32768u;
      bulk_g2s(reinterpret_cast<char*>(img) + o, reinterpret_cast<const char*>(p.image) + o, n, &bars[0]);
    }
  }
  mbar_wait(&bars[0], 0);

  const float* bias = img + p.off_bias;
  const float* wlast = img + p.off_wlast;
  const float blast = __ldg(p.blast);
  constexpr uint32_t idesc = make_idesc_tf32(kTcRows, NP);
  const uint32_t img_addr = smem_u32(img);
  uint32_t phase = 0;
  const long long ntiles = (p.Q + kTcRows - 1) / kTcRows;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long q = tile * kTcRows + t;
    const bool valid = q < p.Q;
    int r = 0, kn = 0;
    float xv = 0.f;
    if (valid) { r = (int)(q / nodes); kn = (int)(q % nodes); xv = __ldg(p.x + r); }
    // ---- layer-0 input row [X, h_r, 0...] -> TMEM A region
    for (int c = 0; c < p.kp[0]; c += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = c + j;
        float val = 0.f;
        if (valid) {
          if (k == 0) val = (xv * (__ldg(p.ccn + kn) + 1.f)) / 2.f;
          else if (k <= p.E) val = __ldg(p.h + (size_t)r * p.E + (k - 1));
        }
        v[j] = val;
      }
      tmem_st8(tA + c, v);
    }
    tmem_wait_st();
    fence_before_sync();
    __syncthreads();

    float y = 0.f;
    for (int l = 0; l < p.L; ++l) {
      if (t == 0) {
        fence_after_sync();
        const uint32_t wbase = img_addr + p.off[l] * 4u;
        const int nk = p.kp[l] / 8;
        for (int ks = 0; ks < nk; ++ks) {
          // B: [k/4][n][k%4] image -> lbo (between 16-byte K chunks) = NP*16, sbo (between 8-row groups) = 128
          const uint64_t bdesc = make_smem_desc(wbase + (uint32_t)ks * 2u * NP * 16u, NP * 16u, 128u);
          mma_tf32_ts(tmem_base + NP, tmem_base + ks * 8, bdesc, idesc, ks > 0 ? 1u : 0u);
        }
        mma_commit(&bars[1]);
      }
      mbar_wait(&bars[1], phase);
      phase ^= 1u;
      fence_after_sync();
      const float* bl = bias + l * NP;
      if (l + 1 < p.L) {
#pragma unroll 1
        for (int c = 0; c < NP; c += 16) {
          float v[16];
          tmem_ld16(tD + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + bl[c + j], 0.f);
          tmem_st16(tA + c, v);
        }
        tmem_wait_st();
        fence_before_sync();
        __syncthreads();
      } else {
#pragma unroll 1
        for (int c = 0; c < NP; c += 16) {
          float v[16];
          tmem_ld16(tD + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) y = fmaf(fmaxf(v[j] + bl[c + j], 0.f), wlast[c + j], y);
        }
        y += blast;
      }
    }
    // ---- integrand value, CC-weighted segment sums (same epilogue as the strict kernel)
    const float f = (y > 0.f ? y : expm1f(y)) + 1.05f;
    float wv = 0.f;
    if (valid) {
      wv = __ldg(p.ccw + kn) * f;
      if (kn == 0) {
        p.jac[r] = f;
        if (p.logdet) atomicAdd(p.logdet + r / p.d, logf(f));
      }
    }
    red[t] = wv;
    fence_before_sync();
    __syncthreads();
    if (valid && (kn == 0 || t == 0)) {
      float s = 0.f;
      int rem = nodes - kn;
      if (rem > kTcRows - t) rem = kTcRows - t;
      for (int i = 0; i < rem; ++i) s += red[t + i];
      float c = s * xv / 2.f;
      if (kn == 0) c += __ldg(p.h + (size_t)r * p.E);
      atomicAdd(p.z + r, c);
      if (p.zrev) { const int b = r / p.d, i = r % p.d; atomicAdd(p.zrev + (size_t)b * p.d + (p.d - 1 - i), c); }
    }
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.alloc_cols);
}

// ------------------------------------------------------------------------------------------------
// v2: two tiles in flight per CTA, three rotating TMEM regions.
//
// Warpgroup g (128 threads) owns the CTA's tiles of parity g; a ninth warp is the MMA issuer.  The tensor pipe
// executes the MMAs of the two tiles alternately (global MMA sequence m = 0,1,2,...; m % 2 = warpgroup); while tile
// X's layer runs on the tensor cores, tile Y's epilogue (TMEM -> registers -> bias/ReLU -> TMEM, IN PLACE) runs on
// the CUDA cores.  Hand-shake per tile: epi_done[g] (epilogue warps -> issuer) and mma_done[g] (tcgen05.commit).
// With the in-place epilogue a tile needs two regions only while its MMA runs (A and D) and one otherwise, so
// three NP-column regions suffice:   A(m) = m % 3,  D(m) = (m + 2) % 3   (D(m) is the region freed by MMA m-1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <int NP>
__global__ void __launch_bounds__(2 * kTcRows + 32) umnn_fwd_tc2_kernel(TcFwdParams p) {
  using namespace tc;
  GNF_SMEM(float, smem);
  float* img = smem;
  // [0]: weights landed, [1],[2]: MMA done (tile parity 0/1), [3],[4]: epilogue done / input staged (parity 0/1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ((p.total_floats + 3) / 4) * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* red_all = reinterpret_cast<float*>(bars + 6);                                   // [2][128]
  const int tid = threadIdx.x, g = (tid >> 7) & 1, t = tid & 127, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == 8;
  float* red = red_all + g * kTcRows;
  const int nodes = p.S + 1;
  const int L = p.L;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 4);                  // one arrive per epilogue warp
    mbar_init(&bars[4], 4);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;

  if (tid == 0) {
    const uint32_t bytes = p.total_floats * 4u;
    mbar_expect_tx(&bars[0], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t n = (bytes - o < 32768u) ? bytes - o : 32768u;
      bulk_g2s(reinterpret_cast<char*>(img) + o, reinterpret_cast<const char*>(p.image) + o, n, &bars[0]);
    }
  }
  mbar_wait(&bars[0], 0);

  const float* bias = img + p.off_bias;
  const float* wlast = img + p.off_wlast;
  const float blast = __ldg(p.blast);
  constexpr uint32_t idesc = make_idesc_tf32(kTcRows, NP);
  const uint32_t img_addr = smem_u32(img);
  const long long ntiles = (p.Q + kTcRows - 1) / kTcRows;
  // CTA-local tiles c = 0,1,..: global tile blockIdx.x + c*gridDim.x; both warpgroups run the same number of
  // iterations (a warpgroup without a tile processes an all-invalid one) so the MMA hand-shake never stalls.
  const long long cta_tiles = (ntiles > (long long)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_local = (cta_tiles + 1) / 2;
  const int K0P = p.kp[0];

  // per-thread description of the current tile row, and prefetched input of the next tile
  bool valid = false; int r = 0, kn = 0; float xv = 0.f;
  float pre[32];
  auto load_tile_row = [&](long long c_local) {
    const long long c = 2 * c_local + g;
    const long long q = (c < cta_tiles) ? ((long long)blockIdx.x + c * gridDim.x) * kTcRows + t : p.Q;
    valid = q < p.Q;
    r = 0; kn = 0; xv = 0.f;
    if (valid) { r = (int)(q / nodes); kn = (int)(q % nodes); xv = ldg_pinned(p.x + r); }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float val = 0.f;
      if (valid && k < K0P) {
        if (k == 0) val = (xv * (ldg_pinned(p.ccn + kn) + 1.f)) / 2.f;
        else if (k <= p.E) val = ldg_pinned(p.h + (size_t)r * p.E + (k - 1));
      }
      pre[k] = val;
    }
  };
  auto store_input = [&](uint32_t region_addr) {
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      if (c < K0P) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = pre[c + j];
        tmem_st8(region_addr + c, v);
      }
    }
  };

  const long long n_iter = n_local * L;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && (is_issuer ? lane == 0 : t == 0);
  auto stamp = [&](long long i, int gg, int slot) {
    if (tracing && i < 24) p.trace[(i * 2 + gg) * 8 + slot] = clock64();
  };

  if (is_issuer) {
    // ===================== MMA issuer warp =====================
    {   // the whole warp runs the loop converged; one elected lane issues (tc_common.cuh: mma_*_w)
      for (long long i = 0; i < n_iter; ++i) {
        const int layer = (int)(i % L);
        const uint32_t wbase = img_addr + p.off[layer] * 4u;
        const int nk = p.kp[layer] / 8;
        for (int gg = 0; gg < 2; ++gg) {
          const long long m = 2 * i + gg;
          const uint32_t rA = (uint32_t)(m % 3) * NP, rD = (uint32_t)((m + 2) % 3) * NP;
          // the tile's layer input is staged in TMEM (its previous epilogue, or the tile prologue)
          mbar_wait(&bars[3 + gg], (uint32_t)(i & 1));
          stamp(i, gg, 0);
          // D(m) is the A region of MMA m-1: that MMA must have finished reading it
          if (m > 0) {
            const long long io = (gg == 1) ? i : i - 1;
            mbar_wait(&bars[1 + (1 - gg)], (uint32_t)(io & 1));
          }
          stamp(i, gg, 1);
          fence_after_sync();
          uint64_t bdesc = make_smem_desc(wbase, NP * 16u, 128u);
          constexpr uint64_t kDescStep = (2u * NP * 16u) >> 4;   // one k-step = two 16-byte K chunks of the image
          uint32_t ta = tmem_base + rA;
          const uint32_t td = tmem_base + rD;
          mma_tf32_ts_w(td, ta, bdesc, idesc, 0u);
#pragma unroll 4
          for (int ks = 1; ks < nk; ++ks) {
            bdesc += kDescStep;
            ta += 8;
            mma_tf32_ts_w(td, ta, bdesc, idesc, 1u);
          }
          mma_commit_w(&bars[1 + gg]);
          stamp(i, gg, 2);
        }
      }
    }
  } else {
  // ===================== epilogue warpgroups =====================
  // prologue: first tile's input into region A(m = g) = g
  load_tile_row(0);
  store_input(tmem_base + lane_sel + (uint32_t)(g % 3) * NP);
  tmem_wait_st();
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive_cta(&bars[3 + g]);

  for (long long i = 0; i < n_iter; ++i) {
    const int layer = (int)(i % L);
    const long long m = 2 * i + g;
    const uint32_t rD = (uint32_t)((m + 2) % 3) * NP;
    // what this thread needs after the last layer: its own row's result context, then the next tile's input
    bool cur_valid = valid; int cur_r = r, cur_kn = kn; float cur_xv = xv;
    if (layer == L - 1) load_tile_row(i / L + 1);      // prefetch while the tensor cores work
    mbar_wait(&bars[1 + g], (uint32_t)(i & 1));
    stamp(i, g, 3);
    fence_after_sync();
    const uint32_t R = tmem_base + lane_sel + rD;
    const float* bl = bias + layer * NP;
    // Software-pipelined epilogue over 32-column chunks (a 16-column tail when NP % 32 != 0): chunk c+1's TMEM load is
    // in flight while chunk c is processed.  TMEM reads are the scarce resource (~64 B/clk/SM), so the loads are kept
    // back to back; tcgen05.wait::ld waits for ALL outstanding loads, hence exactly one chunk of look-ahead.
    constexpr int NC = (NP + 31) / 32;
    uint32_t buf[2][32];
    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
    if (NP >= 32) tmem_ld32p(R, buf[0]); else tmem_ld16p(R, buf[0]);
    tmem_wait_ld();
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
      const int c = ci * 32;
      const int w = (NP - c >= 32) ? 32 : 16;               // width of this chunk
      uint32_t* cur = buf[ci & 1];
      uint32_t* nxt = buf[(ci + 1) & 1];
      if (ci + 1 < NC) {
        if (NP - (c + 32) >= 32) tmem_ld32p(R + c + 32, nxt); else tmem_ld16p(R + c + 32, nxt);
      }
      if (layer < L - 1) {
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          if (j4 < w) {
            const float4 b4 = *reinterpret_cast<const float4*>(bl + c + j4);
            cur[j4 + 0] = __float_as_uint(fmaxf(__uint_as_float(cur[j4 + 0]) + b4.x, 0.f));
            cur[j4 + 1] = __float_as_uint(fmaxf(__uint_as_float(cur[j4 + 1]) + b4.y, 0.f));
            cur[j4 + 2] = __float_as_uint(fmaxf(__uint_as_float(cur[j4 + 2]) + b4.z, 0.f));
            cur[j4 + 3] = __float_as_uint(fmaxf(__uint_as_float(cur[j4 + 3]) + b4.w, 0.f));
          }
        }
        if (w == 32) tmem_st32p(R + c, cur); else tmem_st16p(R + c, cur);
      } else {
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          if (j4 < w) {
            const float4 b4 = *reinterpret_cast<const float4*>(bl + c + j4);
            const float4 w4 = *reinterpret_cast<const float4*>(wlast + c + j4);
            y0 = fmaf(fmaxf(__uint_as_float(cur[j4 + 0]) + b4.x, 0.f), w4.x, y0);
            y1 = fmaf(fmaxf(__uint_as_float(cur[j4 + 1]) + b4.y, 0.f), w4.y, y1);
            y2 = fmaf(fmaxf(__uint_as_float(cur[j4 + 2]) + b4.z, 0.f), w4.z, y2);
            y3 = fmaf(fmaxf(__uint_as_float(cur[j4 + 3]) + b4.w, 0.f), w4.w, y3);
          }
        }
      }
      if (ci + 1 < NC) tmem_wait_ld();
    }
    if (layer == L - 1) {
      float y = (y0 + y1) + (y2 + y3);
      // next tile's input row goes in place into this region (its A for MMA m+2); hand the region to the issuer
      // BEFORE the reduction / atomics so that they stay off the tensor pipe's critical path
      store_input(R);
      tmem_wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&bars[3 + g]);
      stamp(i, g, 4);
      y += blast;
      const float f = (y > 0.f ? y : expm1f(y)) + 1.05f;
      float wv = 0.f;
      if (cur_valid) {
        wv = __ldg(p.ccw + cur_kn) * f;
        if (cur_kn == 0) {
          p.jac[cur_r] = f;
          if (p.logdet) atomicAdd(p.logdet + cur_r / p.d, logf(f));
        }
      }
      named_bar_sync(1 + g, kTcRows);           // previous tile's readers of red[] are done
      red[t] = wv;
      named_bar_sync(1 + g, kTcRows);
      if (cur_valid && (cur_kn == 0 || t == 0)) {
        float sacc = 0.f;
        int rem = nodes - cur_kn;
        if (rem > kTcRows - t) rem = kTcRows - t;
        for (int k = 0; k < rem; ++k) sacc += red[t + k];
        float c = sacc * cur_xv / 2.f;
        if (cur_kn == 0) c += __ldg(p.h + (size_t)cur_r * p.E);
        atomicAdd(p.z + cur_r, c);
        if (p.zrev) { const int b = cur_r / p.d, ii = cur_r % p.d; atomicAdd(p.zrev + (size_t)b * p.d + (p.d - 1 - ii), c); }
      }
      stamp(i, g, 5);
    } else {
      tmem_wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&bars[3 + g]);
      stamp(i, g, 4);
    }
  }
  }  // epilogue warpgroups
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Self-test kernel: C[128, N] = A[128, K] * W[N, K]^T with one CTA, TS (A via TMEM) or SS (A via smem) form.
// Exercises exactly the descriptor / TMEM-layout conventions the fused kernel relies on.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int N, int K, int NPc, int KP, int alloc_cols, int ss_mode) {
  using namespace tc;
  GNF_SMEM(float, smem);
  float* wimg = smem;                                   // [KP/4][NPc][4]
  float* aimg = wimg + (size_t)KP * NPc;                // [KP/4][128][4]  (SS mode)
  uint64_t* bars = reinterpret_cast<uint64_t*>(aimg + (size_t)KP * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tmem_alloc(tmem_slot, alloc_cols);
  if (t == 0) { mbar_init(&bars[0], 1); fence_mbar_init(); }
  for (int i = t; i < KP * NPc; i += 128) {
    const int kc = i / (NPc * 4), n = (i / 4) % NPc, kk = i % 4, k = kc * 4 + kk;
    wimg[i] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
  }
  for (int i = t; i < KP * 128; i += 128) {
    const int kc = i / (128 * 4), row = (i / 4) % 128, kk = i % 4, k = kc * 4 + kk;
    aimg[i] = (k < K) ? A[(size_t)row * K + k] : 0.f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const uint32_t colA = 0, colD = 256;                  // A columns [0, KP), D columns [256, 256+NPc)
  if (!ss_mode) {
    for (int c = 0; c < KP; c += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < K) ? A[(size_t)t * K + c + j] : 0.f;
      tmem_st8(tmem_base + lane_sel + colA + c, v);
    }
    tmem_wait_st();
  }
  fence_before_sync();
  __syncthreads();
  if (t == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(128, NPc);
    for (int ks = 0; ks < KP / 8; ++ks) {
      const uint64_t bdesc = make_smem_desc(smem_u32(wimg) + (uint32_t)ks * 2u * NPc * 16u, NPc * 16u, 128u);
      if (ss_mode) {
        const uint64_t adesc = make_smem_desc(smem_u32(aimg) + (uint32_t)ks * 2u * 128u * 16u, 128u * 16u, 128u);
        mma_tf32_ss(tmem_base + colD, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
      } else {
        mma_tf32_ts(tmem_base + colD, tmem_base + colA + ks * 8, bdesc, idesc, ks > 0 ? 1u : 0u);
      }
    }
    mma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  fence_after_sync();
  for (int c = 0; c < NPc; c += 16) {
    float v[16];
    tmem_ld16(tmem_base + lane_sel + colD + c, v);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c + j < N) C[(size_t)t * N + c + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, alloc_cols);
}

// ------------------------------------------------------------------------------------------------
// Micro-probe (measurement tool, not on the product path): TMEM read bandwidth, tcgen05.mma issue rate, and whether
// the two overlap.  mode bit0: 128 threads stream tcgen05.ld.x32 over 256 columns `iters` times; bit1: one thread
// issues `iters` x 19 TF32 MMAs (M128 x N160 x K8, A from TMEM, B from shared memory).  out[0..1] = elapsed clocks of
// the load warps / the MMA thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160) tc_probe_kernel(int mode, int iters, long long* out) {
  using namespace tc;
  GNF_SMEM(float, smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 160 * 152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (t == 0) { mbar_init(&bars[0], 1); fence_mbar_init(); }
  for (int i = t; i < 160 * 152; i += 160) smem[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    if (mode & 1) {
      uint32_t acc = 0;
      const long long c0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 256; c += 32) {
          uint32_t r[32];
          tmem_ld32p(tmem_base + lane_sel + 256 + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
      }
      const long long c1 = clock64();
      if (t == 0) out[0] = c1 - c0;
      if (acc == 0x12345678u) out[3] = acc;
    }
  } else if (t == 128) {
    if (mode & 2) {
      const uint32_t idesc = make_idesc_tf32(128, 160);
      const uint32_t wbase = smem_u32(smem);
      const long long c0 = clock64();
      for (int it = 0; it < iters; ++it) {
        uint64_t bdesc = make_smem_desc(wbase, 160 * 16u, 128u);
        for (int ks = 0; ks < 19; ++ks) {
          mma_tf32_ts(tmem_base + 0, tmem_base + 160 + ks * 8, bdesc, idesc, ks > 0 ? 1u : 0u);
          bdesc += (2u * 160 * 16u) >> 4;
        }
      }
      mma_commit(&bars[0]);
      mbar_wait(&bars[0], 0);
      const long long c1 = clock64();
      out[1] = c1 - c0;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int NP>
static int launch_tc_fwd(const TcFwdParams& p, size_t smem, cudaStream_t s) {
  const long long ntiles = (p.Q + kTcRows - 1) / kTcRows;
  if constexpr (3 * NP <= 512) {
    // two tiles in flight need 3 regions of NP columns and the next tile's input prefetched in 32 registers
    if (p.kp[0] <= 32 && smem + kTcRows * sizeof(float) + 64 <= 227 * 1024) {
      const size_t smem2 = smem + kTcRows * sizeof(float) + 64;
      cudaFuncSetAttribute(umnn_fwd_tc2_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      long long grid = kNumSMs;
      if (grid > (ntiles + 1) / 2) grid = (ntiles + 1) / 2;
      GNF_LAUNCH(umnn_fwd_tc2_kernel<NP>, (unsigned)grid, 2 * kTcRows + 32, smem2, s, p);
      return 0;
    }
  }
  cudaFuncSetAttribute(umnn_fwd_tc_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long grid = kNumSMs;
  if (grid > ntiles) grid = ntiles;
  GNF_LAUNCH(umnn_fwd_tc_kernel<NP>, (unsigned)grid, kTcRows, smem, s, p);
  return 0;
}

}  // namespace gnf
using namespace gnf;
#endif  // !GNF_EMU

extern "C" {

size_t gnf_umnn_tc_workspace_bytes(const gnf_mlp_t* net) {
#ifdef GNF_EMU
  (void)net;
  gnf::set_error("tensor-core kernels have no host-simulator flavour");
  return 0;
#else
  TcPlan pl;
  if (make_tc_plan(net, &pl)) return 0;
  return pl.total_floats * sizeof(float);
#endif
}

int gnf_umnn_fwd_tc(const float* x, const float* h, const gnf_mlp_t* net, int S, const float* ccw, const float* ccn, float* z,
                    float* zrev, float* jac, float* logdet, int R, int d, void* work, size_t work_bytes, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!x || !h || !net || !ccw || !ccn || !z || !jac || R < 0 || d <= 0 || S < 1 || (R % d) != 0) return fail(GNF_ERR_INVALID, "gnf_umnn_fwd_tc: bad arguments");
  TcPlan pl;
  if (int e = make_tc_plan(net, &pl)) return e;
  if (!work || work_bytes < pl.total_floats * sizeof(float)) return fail(GNF_ERR_WORKSPACE, "gnf_umnn_fwd_tc: workspace too small");
  if ((reinterpret_cast<uintptr_t>(work) & 15) != 0) return fail(GNF_ERR_INVALID, "gnf_umnn_fwd_tc: workspace must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (R == 0) return 0;
  TcPackArgs a;
  for (int l = 0; l < GNF_MAX_LAYERS; ++l) { a.W[l] = nullptr; a.b[l] = nullptr; a.off[l] = 0; a.kp[l] = 0; }
  for (int l = 0; l <= pl.L; ++l) { a.W[l] = net->W[l]; a.b[l] = net->b[l]; }
  for (int l = 0; l < pl.L; ++l) { a.off[l] = pl.off[l]; a.kp[l] = pl.kp[l]; }
  a.off_bias = pl.off_bias; a.off_wlast = pl.off_wlast;
  for (int l = 0; l <= net->n_layers; ++l) a.dims[l] = net->dims[l];
  a.L = pl.L; a.NP = pl.NP;
  int blocks = (int)((pl.total_floats + 255) / 256);
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  GNF_LAUNCH(umnn_tc_pack_kernel, blocks, 256, 0, s, a, (float*)work, pl.total_floats);
  cudaMemsetAsync(z, 0, (size_t)R * sizeof(float), s);
  if (zrev) cudaMemsetAsync(zrev, 0, (size_t)R * sizeof(float), s);
  if (logdet) cudaMemsetAsync(logdet, 0, (size_t)(R / d) * sizeof(float), s);
  TcFwdParams p;
  p.x = x; p.h = h; p.ccw = ccw; p.ccn = ccn; p.image = (const float*)work; p.blast = net->b[pl.L];
  p.z = z; p.zrev = zrev; p.jac = jac; p.logdet = logdet;
  p.R = R; p.d = d; p.E = net->dims[0] - 1; p.S = S; p.L = pl.L; p.alloc_cols = pl.alloc_cols;
  p.Q = (long long)R * (S + 1);
  for (int l = 0; l < GNF_MAX_LAYERS; ++l) { p.kp[l] = l < pl.L ? pl.kp[l] : 0; p.off[l] = l < pl.L ? (unsigned)pl.off[l] : 0; }
  p.off_bias = (unsigned)pl.off_bias; p.off_wlast = (unsigned)pl.off_wlast; p.total_floats = (unsigned)pl.total_floats;
  p.trace = g_tc_trace;
  const size_t smem = ((pl.total_floats + 3) / 4 * 4) * sizeof(float) + 4 * sizeof(uint64_t) + kTcRows * sizeof(float);
  int e = 0;
  switch (pl.NP) {
    case 16: e = launch_tc_fwd<16>(p, smem, s); break;
    case 32: e = launch_tc_fwd<32>(p, smem, s); break;
    case 48: e = launch_tc_fwd<48>(p, smem, s); break;
    case 64: e = launch_tc_fwd<64>(p, smem, s); break;
    case 80: e = launch_tc_fwd<80>(p, smem, s); break;
    case 96: e = launch_tc_fwd<96>(p, smem, s); break;
    case 112: e = launch_tc_fwd<112>(p, smem, s); break;
    case 128: e = launch_tc_fwd<128>(p, smem, s); break;
    case 144: e = launch_tc_fwd<144>(p, smem, s); break;
    case 160: e = launch_tc_fwd<160>(p, smem, s); break;
    case 176: e = launch_tc_fwd<176>(p, smem, s); break;
    case 192: e = launch_tc_fwd<192>(p, smem, s); break;
    case 208: e = launch_tc_fwd<208>(p, smem, s); break;
    case 224: e = launch_tc_fwd<224>(p, smem, s); break;
    case 240: e = launch_tc_fwd<240>(p, smem, s); break;
    default: e = launch_tc_fwd<256>(p, smem, s); break;
  }
  if (e) return e;
  return check_launch("gnf_umnn_fwd_tc");
#endif
}

#ifdef GNF_DEVTOOLS
/* Debug: subsequent gnf_umnn_fwd_tc calls record SM-clock stamps of CTA 0's phases into `buf`
 * (48 x 8 x int64, device); NULL disables.  Not thread safe; measurement tool only. */
int gnf_tc_set_trace(long long* buf) {
#ifndef GNF_EMU
  gnf::g_tc_trace = buf;
#else
  (void)buf;
#endif
  return 0;
}
#endif

#ifdef GNF_DEVTOOLS
/* Measurement tool: see tc_probe_kernel.  out: 4 x int64 (device). */
int gnf_tc_probe(int mode, int iters, long long* out, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  const size_t smem = 160 * 152 * sizeof(float) + 64;
  cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GNF_LAUNCH(tc_probe_kernel, 1, 160, smem, (cudaStream_t)stream, mode, iters, out);
  return check_launch("gnf_tc_probe");
#endif
}
#endif

/* C[128,N] = A[128,K] W[N,K]^T on one CTA through tcgen05 (mode 0: A staged in TMEM, mode 1: A in shared memory). */
int gnf_tc_selftest(const float* A, const float* W, float* C, int N, int K, int mode, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!A || !W || !C || N < 1 || N > 256 || K < 1 || K > 256) return fail(GNF_ERR_INVALID, "gnf_tc_selftest: bad arguments");
  const int NPc = round_up(N, 16), KP = round_up(K, 8);
  const size_t smem = ((size_t)KP * NPc + (size_t)KP * 128) * sizeof(float) + 64;
  if (smem > 227 * 1024) return fail(GNF_ERR_UNSUPPORTED, "gnf_tc_selftest: too large");
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GNF_LAUNCH(tc_selftest_kernel, 1, 128, smem, (cudaStream_t)stream, A, W, C, N, K, NPc, KP, 512, mode);
  return check_launch("gnf_tc_selftest");
#endif
}

}  // extern "C"
