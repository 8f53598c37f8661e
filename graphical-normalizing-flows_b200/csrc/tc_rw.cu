// Resident-weight layer GEMM on tcgen05 / TMEM for the layer-wise UMNN engine (umnn_lw.cu):
//     C[M, NP] = epi( A[M, KP] * B[NP, KP]^T ),     NP <= 160, KP <= 160,   M = B*d*(S+2) node-rows (10^5 .. 10^7)
// in 3xTF32 (fp32-equivalent) or single-pass TF32.  (IntegrandNet hidden layers, MonotonicNormalizer.py:12-38, forward
// and dgrad over all quadrature node-rows.)
//
// The generic engine (tc_gemm.cu) re-streams the whole weight matrix per 128-row tile, splits both operands into hi/lo
// in shared memory, and pays a 700-instruction transposing epilogue per 16 columns: at K = N = 150 it runs at 5x the
// HBM time of the layer.  Here the shapes are fixed by the integrand width, so:
//   * B (the layer's weights) is split into TF32 hi / lo ONCE per call by a pack kernel, in the UMMA canonical K-major
//     image; each persistent CTA pulls both images (2 x NP x KP x 4 B <= 195 KB) into shared memory with bulk async
//     copies and keeps them for its lifetime;
//   * A is never staged as an MMA operand in shared memory: loader thread t owns tile row t = TMEM lane t, splits its
//     row chunk in registers and writes A_hi / A_lo into TMEM (tcgen05.st); the MMAs are TS form (A from TMEM, B from
//     shared memory);
//   * accumulation order: all correction products (a_lo*b_hi, a_hi*b_lo) first, while the accumulator is ~2^-11 of its
//     final size, then the a_hi*b_hi chain.  The tensor core accumulates with truncation, so what matters is the number
//     of additions made at full accumulator magnitude: KP/8 = 19 here instead of 57 (measured: gradient error vs float64
//     2e-5, the same as the FFMA engine; a single 57-step chain gives 3.5e-4);
//   * global memory is only ever touched with full 128-byte lines per 8 lanes.  Row-owner threads reading / writing
//     their own rows directly was measured at 18-40 k clocks per tile (scripts/rw_trace.py: 16-byte row-owner stores
//     leave half-written sectors that L2 must fill from DRAM; 32-byte ones still make 32 requests per instruction), so
//     every 32-row x 32-column block goes through a 4 KB XOR-swizzled shared-memory block per warp that turns the
//     coalesced (4 rows x 128 B per instruction) global layout into the row-owner layout and back, conflict-free;
//   * epilogue thread t owns row t: tcgen05.ld 32 columns -> bias + ReLU (or ReLU-mask bits) -> staged store, the ReLU
//     bit mask of the output row as one word per 32 columns.
// Roles: warps 0-3 epilogue, warps 4-11 loaders (two per TMEM lane quarter, alternating 32-column chunks), warp 12 the
// MMA issuer (one thread; it shared a loader warp at first, and that warp's serial load -> issue -> load chain was the
// whole tile period: profiles/r01z_rw_gemm_trace.txt).  A and D are single-buffered in TMEM (A_hi 160 + A_lo 160 + D 160
// columns), so both are held as briefly as possible: every loader thread has exactly ONE load group outstanding -- its
// next chunk -- issued as soon as the current one is staged, converts it to the row-owner layout and splits it BEFORE
// waiting for A to be free; an epilogue thread drains three accumulator blocks into registers at once (tcgen05.ld.x32)
// and hands D back as soon as the last block is read, so the next tile's MMAs overlap most of the epilogue and the
// next tile's loads overlap the MMAs.
#include "tc_common.cuh"
#include "tc_rw.h"

#ifndef GNF_EMU
namespace gnf {

constexpr int kRwRows = 128;
constexpr int kRwEpiWarps = 4, kRwLoadWarps = 8;
constexpr int kRwEpiStageFloats = 32 * 32;                     // epilogue warp: one 32-row x 32-column transposition block
constexpr int kRwLoadStageFloats = 32 * 16;                    // loader warp: one 32-row x 16-column block (two passes per chunk)
constexpr int kRwStageFloatsTotal = kRwEpiWarps * kRwEpiStageFloats + kRwLoadWarps * kRwLoadStageFloats;
constexpr int kRwThreads = 16 * 32;   // 4 epilogue + 8 loader + 1 MMA issuer warp (+3 idle: registers come in 4-warp granules): 128 registers per thread
constexpr int kRwColAhi = 0, kRwColAlo = 160, kRwColD = 320;   // TMEM columns
constexpr int kRwMaxK = 160, kRwMaxN = 160;
constexpr int kRwMaxChunks = kRwMaxK / 32;
constexpr size_t kRwSmemBudget = 227 * 1024;

__device__ __forceinline__ uint32_t rw_rn_tf32(uint32_t u) { return (u + 0x1000u) & 0xffffe000u; }

// image[(k/4)][n][k%4] (hi), same (lo), bias[NP]
__global__ void rw_pack_kernel(const float* __restrict__ W, long long ldw, int N, int K, int transpose, const float* __restrict__ bias,
                               float* __restrict__ image, int NP, int KP) {
  const int total = NP * KP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + NP; i += gridDim.x * blockDim.x) {
    if (i >= total) {
      const int n = i - total;
      image[2 * (size_t)total + n] = (bias && n < N) ? bias[n] : 0.f;
      continue;
    }
    const int kc = i / (NP * 4), n = (i / 4) % NP, k = kc * 4 + (i % 4);
    float v = 0.f;
    if (n < N && k < K) v = transpose ? W[(long long)k * ldw + n] : W[(long long)n * ldw + k];
    const float hi = __uint_as_float(rw_rn_tf32(__float_as_uint(v)));
    const float lo = __uint_as_float(rw_rn_tf32(__float_as_uint(v - hi)));
    image[i] = hi;
    image[(size_t)total + i] = lo;
  }
}

__device__ __forceinline__ float4 rw_ldg16(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void rw_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// Measurement: role r of CTA 0 appends SM-clock stamps to row r of the trace buffer (kRwTraceLen stamps per role).
constexpr int kRwTraceLen = 256;
struct RwTrace {
  long long* buf; int n;
  __device__ __forceinline__ void stamp() { if (buf && n < kRwTraceLen) buf[n++] = clock64(); }
};

// NB = NP / 32 output column blocks, NCH = ceil(KP / 32) input column chunks (compile time: whole rows live in registers).
// ablation bits (skip stores / loads / MMAs) exist in the development build only
#ifdef GNF_DEVTOOLS
#define RW_DEBUG (p.debug)
#else
#define RW_DEBUG 0
#endif
template <int NB, int NCH>
__global__ void __launch_bounds__(kRwThreads, 1) rw_gemm_kernel(RwGemmParams p) {
  using namespace tc;
  GNF_SMEM(float, smem);
  constexpr int NP = NB * 32;
  const int KP = p.KP;
  const uint32_t img_floats = 2u * (uint32_t)NP * (uint32_t)KP + (uint32_t)NP;
  const float* bias_s = smem + 2 * NP * KP;
  float* stage_all = smem + ((img_floats + 31) / 32) * 32;   // per-warp transposition blocks, XOR-swizzled 16-byte slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + kRwStageFloatsTotal);
  uint64_t* w_full = bars;                   // weight images landed
  uint64_t* a_full = bars + 1;               // [kRwMaxChunks] loaders -> MMA (one arrive per quarter warp)
  uint64_t* a_empty = bars + 1 + kRwMaxChunks;   // MMA -> loaders (tcgen05.commit)
  uint64_t* d_full = a_empty + 1;            // MMA -> epilogue (tcgen05.commit)
  uint64_t* d_empty = d_full + 1;            // epilogue -> MMA (one arrive per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    mbar_init(w_full, 1);
    for (int c = 0; c < kRwMaxChunks; ++c) mbar_init(&a_full[c], 4);          // one arrive per lane quarter
    mbar_init(a_empty, 1);
    mbar_init(d_full, 1);
    mbar_init(d_empty, kRwEpiWarps);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  if (tid == 0) {
    const uint32_t bytes = img_floats * 4u;
    mbar_expect_tx(w_full, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t n = (bytes - o < 32768u) ? bytes - o : 32768u;
      bulk_g2s(reinterpret_cast<char*>(smem) + o, reinterpret_cast<const char*>(p.image) + o, n, w_full);
    }
  }

  const long long ntiles = ((long long)p.M + kRwRows - 1) / kRwRows;
  const long long n_local = (ntiles > (long long)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nksteps = KP / 8;
  const bool split = p.passes == 3;
  const int sub = lane >> 3, piece = lane & 7;               // coalesced view of a 32 x 32 block: lane <-> (row 4i + sub, 16-byte piece)

  if (warp >= kRwEpiWarps + kRwLoadWarps) {
    // ===================== MMA issuer (one thread) =====================
    if (warp == kRwEpiWarps + kRwLoadWarps) {   // the whole warp runs the loop converged; one elected lane issues
      const uint32_t idesc = make_idesc_tf32(kRwRows, NP);
      const uint32_t whi = smem_u32(smem), wlo = whi + (uint32_t)NP * (uint32_t)KP * 4u;
      const uint64_t dhi0 = make_smem_desc(whi, (uint32_t)NP * 16u, 128u), dlo0 = make_smem_desc(wlo, (uint32_t)NP * 16u, 128u);
      const uint64_t dstep = (uint64_t)((2u * (uint32_t)NP * 16u) >> 4);      // one k-step (8 k) = two 16-byte K chunks of the image
      const uint32_t tD = tmem_base + kRwColD, tAhi = tmem_base + kRwColAhi, tAlo = tmem_base + kRwColAlo;
      RwTrace trm = {(p.trace && blockIdx.x == 0 && lane == 0) ? p.trace : nullptr, 0};
      mbar_wait(w_full, 0);
      for (long long tl = 0; tl < n_local; ++tl) {
        trm.stamp();                                         // per tile: start, D free, chunk c issued (x NCH), all issued
        if (tl > 0) mbar_wait(d_empty, (uint32_t)((tl - 1) & 1));
        trm.stamp();
        // correction products of chunk c as soon as all four lane quarters have handed it over; after the last chunk the
        // a_hi * b_hi chain and the two commits (A free for the loaders, D full for the epilogue)
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          mbar_wait(&a_full[c], (uint32_t)(tl & 1));
          fence_after_sync();
          const int k0 = c * 4, k1 = (RW_DEBUG & 4) ? k0 + (c == 0 ? 1 : 0) : ((k0 + 4 < nksteps) ? k0 + 4 : nksteps);
          if (split) {
            for (int kk = k0; kk < k1; ++kk) mma_tf32_ts_w(tD, tAlo + kk * 8, dhi0 + dstep * kk, idesc, (c == 0 && kk == k0) ? 0u : 1u);
            for (int kk = k0; kk < k1; ++kk) mma_tf32_ts_w(tD, tAhi + kk * 8, dlo0 + dstep * kk, idesc, 1u);
          } else {
            for (int kk = k0; kk < k1; ++kk) mma_tf32_ts_w(tD, tAhi + kk * 8, dhi0 + dstep * kk, idesc, (c == 0 && kk == k0) ? 0u : 1u);
          }
          trm.stamp();
        }
        if (split && !(RW_DEBUG & 4))
          for (int kk = 0; kk < nksteps; ++kk) mma_tf32_ts_w(tD, tAhi + kk * 8, dhi0 + dstep * kk, idesc, 1u);
        mma_commit_w(a_empty);
        mma_commit_w(d_full);
        trm.stamp();
      }
    }
  } else if (warp >= kRwEpiWarps) {
    // ===================== loaders =====================
    const int lw = warp - kRwEpiWarps, quarter = lw & 3, half = lw >> 2;     // chunk parity this warp serves
    float* stage = stage_all + kRwEpiWarps * kRwEpiStageFloats + lw * kRwLoadStageFloats;
    const uint32_t tA = tmem_base + lane_sel;
    RwTrace tr = {(p.trace && blockIdx.x == 0 && lane == 0 && quarter == 0) ? p.trace + (1 + half) * kRwTraceLen : nullptr, 0};
    const int n_my = (NCH - half + 1) / 2;                   // chunks per tile served by this warp (0 for odd warps when NCH == 1)
    // the thread's share (coalesced view) of one 32-column chunk: 8 x 16 bytes, rows 4i + sub.  Exactly one load group
    // is outstanding per thread (its next chunk); memory parallelism comes from the eight loader warps.
    float4 buf[8];
    // Unconditional loads from clamped coordinates: a conditional around an inline-asm load is a branch per load (measured in
    // the wgrad loader: 100 clocks per load to issue).  Rows past M and prefetches past the end read valid memory whose
    // values are never used: rows >= M are never stored by the epilogue.
    auto issue = [&](long long tl, int c) {
      if (RW_DEBUG & 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
      }
      const long long tlc = tl < n_local ? tl : n_local - 1;
      const int cc = c < NCH ? c : NCH - 1;
      const long long row0 = ((long long)blockIdx.x + tlc * gridDim.x) * kRwRows + quarter * 32 + sub;
      const float* src = p.A + cc * 32 + 4 * piece;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = (row0 + 4 * i < p.M) ? row0 + 4 * i : (long long)p.M - 1;
        buf[i] = rw_ldg16(src + r * p.lda);
      }
    };
    issue(0, half);
    for (long long tl = 0; tl < n_local; ++tl)
    for (int j = 0; j < n_my; ++j) {                         // (no 64-bit divisions in this loop: they cost a call each)
      const int c = half + 2 * j;
      tr.stamp();                                            // per chunk: start, split in registers + A free, handed over
      // coalesced view -> row-owner view through the 32 x 16 staging block (16-byte slots XOR-swizzled by (row >> 1) & 3),
      // split into TF32 hi / lo in registers -- all of it BEFORE waiting for A, so that the hand-over after the previous
      // tile's last MMA is four TMEM stores
      uint32_t hi[2][16], lo[2][16];
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        if ((piece >> 2) == hb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int R = 4 * i + sub;
            *reinterpret_cast<float4*>(stage + R * 16 + 4 * ((piece & 3) ^ ((R >> 1) & 3))) = buf[i];
          }
        }
        if (hb == 1) {                                       // registers are free: this thread's next chunk goes in flight
          if (j + 1 < n_my) issue(tl, c + 2); else issue(tl + 1, half);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 q = *reinterpret_cast<const float4*>(stage + lane * 16 + 4 * (i ^ ((lane >> 1) & 3)));
          const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t h = rw_rn_tf32(__float_as_uint(v[e]));
            hi[hb][4 * i + e] = h;
            lo[hb][4 * i + e] = rw_rn_tf32(__float_as_uint(v[e] - __uint_as_float(h)));
          }
        }
        __syncwarp();                                        // every lane is done reading the staging block
      }
      if (j == 0 && tl > 0) {                                // the previous tile's MMAs must be done with A
        mbar_wait(a_empty, (uint32_t)((tl - 1) & 1));
        fence_after_sync();
      }
      tr.stamp();
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        tmem_st16p(tA + kRwColAhi + c * 32 + 16 * hb, hi[hb]);
        if (split) tmem_st16p(tA + kRwColAlo + c * 32 + 16 * hb, lo[hb]);
      }
      tmem_wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) rw_mbar_arrive(&a_full[c]);
      tr.stamp();
    }
  } else {
    // ===================== epilogue =====================
    mbar_wait(w_full, 0);                                    // bias lives behind the images
    const uint32_t tD = tmem_base + lane_sel + kRwColD;
    float* stage = stage_all + warp * kRwEpiStageFloats;
    const bool bit_mask = p.epi == RW_EPI_MASK && p.mask_bits != nullptr;
    const bool act_mask = p.epi == RW_EPI_MASK && !bit_mask && p.act != nullptr;
    RwTrace tr = {(p.trace && blockIdx.x == 0 && tid == 0) ? p.trace + 3 * kRwTraceLen : nullptr, 0};
    constexpr int W = NB < 3 ? NB : 3;                       // accumulator blocks held in registers at once (128 registers per thread)
    for (long long tl = 0; tl < n_local; ++tl) {
      tr.stamp();                                            // per tile: start, D full, first window drained, each block stored
      const long long row = ((long long)blockIdx.x + tl * gridDim.x) * kRwRows + warp * 32 + lane;
      const bool valid = row < p.M;
      const long long row0 = row - lane + sub;
      uint32_t mw[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) mw[i] = (bit_mask && valid) ? __ldg(p.mask_bits + row * p.mask_ld + i) : 0u;
      mbar_wait(d_full, (uint32_t)(tl & 1));
      fence_after_sync();
      tr.stamp();
      // drain up to three 32-column blocks of the accumulator row into registers at once; D goes back to the MMA issuer as
      // soon as the last block has been read (for NB = 5: after the second block is staged), not after the last store
      uint32_t acc[W][32];
#pragma unroll
      for (int ci = 0; ci < W; ++ci) tmem_ld32p(tD + ci * 32, acc[ci]);
      tmem_wait_ld();
      if (NB <= W) {
        fence_before_sync();
        __syncwarp();
        if (lane == 0) rw_mbar_arrive(d_empty);
      }
      tr.stamp();
#pragma unroll
      for (int ci = 0; ci < NB; ++ci) {
        uint32_t* cur = acc[ci % W];
        const int c = ci * 32;
        if (p.epi == RW_EPI_BIAS_ACT) {
          uint32_t obits = 0u;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + j4);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = __uint_as_float(cur[j4 + e]) + bb[e];
              if (p.relu) v = fmaxf(v, 0.f);
              obits |= (v > 0.f ? 1u : 0u) << (j4 + e);
              cur[j4 + e] = __float_as_uint(v);
            }
          }
          if (p.bits_out && valid) p.bits_out[row * p.bits_ld + ci] = obits;
        } else if (bit_mask) {
          const uint32_t w = mw[ci];
#pragma unroll
          for (int j = 0; j < 32; ++j) cur[j] = ((w >> j) & 1u) ? cur[j] : 0u;
        } else if (act_mask) {
          if (valid) {
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.act + row * p.ldact + c + j4));
              cur[j4 + 0] = a4.x > 0.f ? cur[j4 + 0] : 0u;
              cur[j4 + 1] = a4.y > 0.f ? cur[j4 + 1] : 0u;
              cur[j4 + 2] = a4.z > 0.f ? cur[j4 + 2] : 0u;
              cur[j4 + 3] = a4.w > 0.f ? cur[j4 + 3] : 0u;
            }
          }
        }
        // row-owner -> coalesced through the swizzled staging block: each store instruction writes 4 rows x 128 bytes
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<uint4*>(stage + lane * 32 + 4 * (j4 ^ (lane & 7))) = make_uint4(cur[4 * j4], cur[4 * j4 + 1], cur[4 * j4 + 2], cur[4 * j4 + 3]);
        __syncwarp();
        if (ci + W < NB) {                                   // the block's registers are free: next accumulator block
          tmem_ld32p(tD + (ci + W) * 32, acc[ci % W]);
          tmem_wait_ld();
          if (ci + W == NB - 1) {                            // that was the last one: D goes back before this block's global stores
            fence_before_sync();
            __syncwarp();
            if (lane == 0) rw_mbar_arrive(d_empty);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int R = 4 * i + sub;
          const uint4 q = *reinterpret_cast<const uint4*>(stage + R * 32 + 4 * (piece ^ (R & 7)));
          if (row0 + 4 * i < p.M && !(RW_DEBUG & 1)) *reinterpret_cast<uint4*>(p.C + (row0 + 4 * i) * p.ldc + c + 4 * piece) = q;
        }
        __syncwarp();
        tr.stamp();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static long long* g_rw_trace = nullptr;
static int g_rw_debug = 0;    // measurement: bit0 skip global stores, bit1 skip global loads, bit2 skip the MMAs

static size_t rw_smem_bytes(int NP, int KP) {
  return ((rw_image_floats(NP, KP) + 31) / 32 * 32 + (size_t)kRwStageFloatsTotal) * sizeof(float) +
         (4 + kRwMaxChunks) * sizeof(uint64_t) + 16;
}

bool rw_supported(int N, int K) {
  const int NP = (N + 31) / 32 * 32, KP = (K + 7) / 8 * 8;
  if (N < 1 || K < 1 || NP > kRwMaxN || KP > kRwMaxK) return false;
  return rw_smem_bytes(NP, KP) <= kRwSmemBudget;
}

void rw_pack_image(const float* W, long long ldw, int N, int K, int transpose, const float* bias, float* image, int NP, int KP,
                   cudaStream_t s) {
  const int total = NP * KP + NP;
  int blocks = (total + 255) / 256;
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  GNF_LAUNCH(rw_pack_kernel, blocks, 256, 0, s, W, ldw, N, K, transpose, bias, image, NP, KP);
}

int launch_rw_gemm(const RwGemmParams& p, cudaStream_t s) {
  if (p.M <= 0) return 0;
  if (p.passes != 1 && p.passes != 3) return fail(GNF_ERR_INVALID, "resident-weight GEMM: passes must be 1 or 3");
  if (p.NP % 32 || p.KP % 8 || p.NP < 32 || p.KP < 8 || p.NP > kRwMaxN || p.KP > kRwMaxK)
    return fail(GNF_ERR_UNSUPPORTED, "resident-weight GEMM: padded widths %d x %d out of range", p.NP, p.KP);
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (p.lda % 4) || p.lda < (p.KP + 31) / 32 * 32 || (reinterpret_cast<uintptr_t>(p.C) & 15) ||
      (p.ldc % 4) || p.ldc < p.NP || (reinterpret_cast<uintptr_t>(p.image) & 15) || (p.act && ((reinterpret_cast<uintptr_t>(p.act) & 15) || (p.ldact % 4))))
    return fail(GNF_ERR_INVALID, "resident-weight GEMM: operands must be 16-byte aligned with leading dimensions padded to 32 floats");
  const size_t smem = rw_smem_bytes(p.NP, p.KP);
  if (smem > kRwSmemBudget) return fail(GNF_ERR_UNSUPPORTED, "resident-weight GEMM: weight images need %zu B of shared memory", smem);
  const long long ntiles = ((long long)p.M + kRwRows - 1) / kRwRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  RwGemmParams q = p;
  q.trace = g_rw_trace;
  q.debug = g_rw_debug;
  const int NB = p.NP / 32, NCH = (p.KP + 31) / 32;
#define RW_CASE(nb, nch)                                                                                                  \
  if (NB == nb && NCH == nch) {                                                                                           \
    cudaFuncSetAttribute(rw_gemm_kernel<nb, nch>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRwSmemBudget);       \
    GNF_LAUNCH((rw_gemm_kernel<nb, nch>), grid, kRwThreads, smem, s, q);                                                  \
    return 0;                                                                                                             \
  }
#define RW_ROW(nb) RW_CASE(nb, 1) RW_CASE(nb, 2) RW_CASE(nb, 3) RW_CASE(nb, 4) RW_CASE(nb, 5)
  RW_ROW(1) RW_ROW(2) RW_ROW(3) RW_ROW(4) RW_ROW(5)
#undef RW_ROW
#undef RW_CASE
  return fail(GNF_ERR_UNSUPPORTED, "resident-weight GEMM: no kernel for %d x %d", p.NP, p.KP);
}

}  // namespace gnf
using namespace gnf;
#endif

extern "C" {

size_t gnf_linear_rw_workspace_bytes(int N, int K) {
#ifdef GNF_EMU
  (void)N; (void)K;
  gnf::set_error("tensor-core kernels have no host-simulator flavour");
  return 0;
#else
  if (!rw_supported(N, K)) { set_error("resident-weight GEMM: %d x %d weights do not fit shared memory as hi/lo TF32 images", N, K); return 0; }
  return rw_image_floats((N + 31) / 32 * 32, (K + 7) / 8 * 8) * sizeof(float);
#endif
}

#ifdef GNF_DEVTOOLS
int gnf_linear_rw_set_debug(int bits) {
#ifndef GNF_EMU
  gnf::g_rw_debug = bits;
#else
  (void)bits;
#endif
  return 0;
}
#endif

#ifdef GNF_DEVTOOLS
int gnf_linear_rw_set_trace(long long* buf) {
#ifndef GNF_EMU
  gnf::g_rw_trace = buf;
#else
  (void)buf;
#endif
  return 0;
}
#endif

int gnf_linear_fwd_rw(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, uint32_t* bits_out,
                      int M, int N, int K, int relu, int passes, void* work, size_t work_bytes, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!X || !W || !Y || M < 0 || N <= 0 || K <= 0 || ldw < K) return fail(GNF_ERR_INVALID, "gnf_linear_fwd_rw: bad arguments");
  if (!rw_supported(N, K)) return fail(GNF_ERR_UNSUPPORTED, "gnf_linear_fwd_rw: %d x %d weights do not fit", N, K);
  const int NP = (N + 31) / 32 * 32, KP = (K + 7) / 8 * 8;
  if (!work || work_bytes < rw_image_floats(NP, KP) * sizeof(float)) return fail(GNF_ERR_WORKSPACE, "gnf_linear_fwd_rw: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  rw_pack_image(W, ldw, N, K, 0, bias, (float*)work, NP, KP, s);
  RwGemmParams p = {};
  p.A = X; p.lda = ldx; p.C = Y; p.ldc = ldy; p.image = (const float*)work;
  p.M = M; p.NP = NP; p.KP = KP; p.passes = passes; p.epi = RW_EPI_BIAS_ACT; p.relu = relu;
  p.bits_out = bits_out; p.bits_ld = NP / 32;
  if (int e = launch_rw_gemm(p, s)) return e;
  return check_launch("gnf_linear_fwd_rw");
#endif
}

int gnf_linear_dgrad_rw(const float* dY, int lddy, const float* W, int ldw, const float* act, int ldact, const uint32_t* mask_bits,
                        float* dX, int lddx, int M, int N, int K, int passes, void* work, size_t work_bytes, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!dY || !W || !dX || M < 0 || N <= 0 || K <= 0 || ldw < K) return fail(GNF_ERR_INVALID, "gnf_linear_dgrad_rw: bad arguments");
  if (!rw_supported(K, N)) return fail(GNF_ERR_UNSUPPORTED, "gnf_linear_dgrad_rw: %d x %d weights do not fit", N, K);
  const int NP = (K + 31) / 32 * 32, KP = (N + 7) / 8 * 8;           // output width = in-features, reduction = out-features
  if (!work || work_bytes < rw_image_floats(NP, KP) * sizeof(float)) return fail(GNF_ERR_WORKSPACE, "gnf_linear_dgrad_rw: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  rw_pack_image(W, ldw, K, N, 1, nullptr, (float*)work, NP, KP, s);
  RwGemmParams p = {};
  p.A = dY; p.lda = lddy; p.C = dX; p.ldc = lddx; p.image = (const float*)work;
  p.M = M; p.NP = NP; p.KP = KP; p.passes = passes; p.epi = RW_EPI_MASK;
  p.mask_bits = mask_bits; p.mask_ld = NP / 32; p.act = act; p.ldact = ldact;
  if (int e = launch_rw_gemm(p, s)) return e;
  return check_launch("gnf_linear_dgrad_rw");
#endif
}

}  // extern "C"
