// Conditioner MLP engine on the 5th-generation tensor cores (tcgen05 / TMEM), TF32 operands with fp32
// accumulation, in two precisions:
//   passes = 1 : single-pass TF32            (fast mode: log-likelihood tolerance 2e-3)
//   passes = 3 : 3xTF32 split  a = a_hi + a_lo,  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo   (fp32-equivalent: strict mode)
//
// One persistent, warp-specialised kernel serves Linear forward, dgrad and wgrad:
//   warps 0-3  epilogue   : thread t <-> accumulator row (TMEM lane) t; TMEM -> registers -> fused epilogue -> global
//   warp  4    MMA issuer : one thread issues tcgen05.mma, commits to mbarriers
//   warps 5-12 producers  : global -> registers -> (hi/lo split) -> 128B-swizzled UMMA shared-memory tiles
// Operands are staged by the producer warps rather than by TMA tensor maps because every operand needs a
// transformation on the way (hi/lo split, zero padding of ragged K / N tails), and because one code path can then
// read either orientation (k-contiguous or row-contiguous global memory) coalesced and still hand the tensor core
// a canonical K-major or MN-major tile.  A k-chunk is 32 fp32 = one 128-byte swizzle atom per tile row.
// Accumulators are double buffered in TMEM (2 x <=256 columns) so a tile's epilogue overlaps the next tile's MMAs.
#include "tc_common.cuh"

#ifndef GNF_EMU
namespace gnf {

constexpr int kGemmBM = 128;
constexpr int kGemmKC = 32;              // k-chunk: 32 fp32 = 128 bytes = one swizzle atom
constexpr int kEpiThreads = 128, kProdThreads = 256, kGemmThreads = kEpiThreads + 32 + kProdThreads;
constexpr int kMaxStages = 6;

enum { TCG_EPI_BIAS_ACT = 0, TCG_EPI_MASK = 1, TCG_EPI_ATOMIC = 2 };
enum { TCG_SRC_K = 0, TCG_SRC_MN = 1 };  // global memory contiguous along the reduction index / along the row index

struct TcGemmParams {
  // C[m, n] = sum_k A(m, k) * B(n, k);  A: Mrows x Kred,  B: Ncols x Kred
  const float* A; long long lda; int a_src;    // TCG_SRC_K: A(m,k) = A[m*lda + k];  TCG_SRC_MN: A(m,k) = A[k*lda + m]
  const float* B; long long ldb; int b_src;    // same convention with n in place of m
  int M, N, K;
  int BN, stages, passes, splits, k_per_split;
  int epi;
  float* C; long long ldc;
  const float* bias; int bias_ld, bias_period, relu;   // TCG_EPI_BIAS_ACT
  const float* act; long long ldact;                    // TCG_EPI_MASK
};

// K-major fp32 tiles use SWIZZLE_128B (16-byte chunks XOR (row & 7), 8-row atoms: SBO = 1024).  MN-major tiles of
// a 32-bit type must use SWIZZLE_128B_BASE32B (32-byte granules XOR (row & 3), 4-row atoms: SBO = 512); LBO is the
// stride between 32-element MN slabs.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(((mn_major ? 4096u : 0u) >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(((mn_major ? 512u : 1024u) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)(mn_major ? 1 : 2) << 61;           // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Fill one operand tile (rows x 32 k) for one k-chunk.  Tile image: 128-byte rows, 16-byte chunks XOR-swizzled by
// (row & 7).  K-major: tile row = operand row.  MN-major: slabs of 32 operand rows; tile row = k index inside the chunk.
// cp.async (LDGSTS) straight into the swizzled tile: no register staging, so a producer thread can have every
// vector of several stages in flight.  src_bytes < VEC*4 zero-fills the remainder (ragged K / row tails).
template <int VEC>
__device__ __forceinline__ void cp_async_vec(char* smem_dst, const float* gmem_src, int src_bytes) {
  const uint32_t d = tc::smem_u32(smem_dst);
  if (VEC == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
  else if (VEC == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
// 3xTF32 split of a landed raw fp32 tile, in place:  hi = rn_tf32(x) (overwrites the raw value), lo = rn_tf32(x - hi).
// The tensor core truncates the fp32 bits it reads (measured: scripts/tf32_round_probe.py); feeding it operands that
// already sit on the TF32 grid makes that truncation a no-op, and round-to-nearest keeps the representation error of
// hi + lo at ~2^-22 |x| and unbiased (truncating both would leave a one-sided 2^-20 error that survives cancellation).
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
template <int VEC>
__device__ __forceinline__ void split_vec(char* img_hi, char* img_lo, int off) {
  float v[VEC];
  if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(img_hi + off); v[0] = q.x; v[1 % VEC] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w; }
  else if (VEC == 2) { const float2 q = *reinterpret_cast<const float2*>(img_hi + off); v[0] = q.x; v[1 % VEC] = q.y; }
  else v[0] = *reinterpret_cast<const float*>(img_hi + off);
  float hi[VEC], lo[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { hi[j] = rn_tf32(v[j]); lo[j] = rn_tf32(v[j] - hi[j]); }
  if (VEC == 4) {
    *reinterpret_cast<float4*>(img_hi + off) = make_float4(hi[0], hi[1 % VEC], hi[2 % VEC], hi[3 % VEC]);
    *reinterpret_cast<float4*>(img_lo + off) = make_float4(lo[0], lo[1 % VEC], lo[2 % VEC], lo[3 % VEC]);
  } else if (VEC == 2) {
    *reinterpret_cast<float2*>(img_hi + off) = make_float2(hi[0], hi[1 % VEC]);
    *reinterpret_cast<float2*>(img_lo + off) = make_float2(lo[0], lo[1 % VEC]);
  } else {
    *reinterpret_cast<float*>(img_hi + off) = hi[0];
    *reinterpret_cast<float*>(img_lo + off) = lo[0];
  }
}

// K-major operand (global memory contiguous along k): tile row = operand row, 128 bytes (32 k) per row, 16-byte chunks
// XOR-swizzled by (row & 7).  Thread -> (row within a pass, vector within the row); passes advance by a multiple of
// 8 rows, so the swizzle term and the in-row offset are per-thread constants.
// kSplitPass = false: issue the cp.async copies of the raw tile.  true: produce the lo tile from the landed raw tile.
template <int VEC, bool kSplitPass>
__device__ __forceinline__ void fill_k(char* img_hi, char* img_lo, const float* __restrict__ g, long long ld, int r0, int r_end,
                                       int k0, int k_end, int rows, int ptid) {
  constexpr int VPR = 32 / VEC, STEP = kProdThreads / VPR;
  const int cv = ptid % VPR, rb = ptid / VPR, e = cv * VEC;
  const int soff0 = rb * 128 + ((((e >> 2) ^ (rb & 7)) << 4) | ((e & 3) << 2));
  const int npass = rows / STEP;
  if (kSplitPass) {
#pragma unroll 4
    for (int ps = 0; ps < npass; ++ps) split_vec<VEC>(img_hi, img_lo, soff0 + ps * STEP * 128);
    return;
  }
  int nk = k_end - (k0 + e);                             // valid elements of this thread's vector along k
  nk = nk < 0 ? 0 : (nk > VEC ? VEC : nk);
  const float* p0 = g + (long long)(r0 + rb) * ld + k0 + e;
#pragma unroll 4
  for (int ps = 0; ps < npass; ++ps) {
    const bool ok = (r0 + rb + ps * STEP) < r_end && nk > 0;
    cp_async_vec<VEC>(img_hi + soff0 + ps * STEP * 128, ok ? p0 + (long long)ps * STEP * ld : g, ok ? nk * 4 : 0);
  }
}

// MN-major operand (global memory contiguous along the operand-row index): slabs of 32 operand rows; tile row = k index
// inside the chunk (32 per chunk), 32-byte granules XOR-swizzled by (k & 3) (SWIZZLE_128B_BASE32B).  Producer warp w
// owns k rows w, w+8, w+16, w+24; lanes sweep the operand rows.
template <int VEC, bool kSplitPass>
__device__ __forceinline__ void fill_mn(char* img_hi, char* img_lo, const float* __restrict__ g, long long ld, int r0, int r_end,
                                        int k0, int k_end, int rows, int ptid) {
  const int w = ptid >> 5, lane = ptid & 31;
  const int per = rows / VEC;                            // vectors per k row
  for (int rv = lane; rv < per; rv += 32) {
    const int r = rv * VEC;
    const int slab = r >> 5, e = r & 31;
    int nr = r_end - (r0 + r);
    nr = nr < 0 ? 0 : (nr > VEC ? VEC : nr);
    const float* p0 = g + (long long)(k0 + w) * ld + r0 + r;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = w + 8 * u;
      const int off = slab * 4096 + kk * 128 + ((((e >> 2) ^ ((kk & 3) << 1)) << 4) | ((e & 3) << 2));
      if (kSplitPass) {
        split_vec<VEC>(img_hi, img_lo, off);
      } else {
        const bool ok = (k0 + kk) < k_end && nr > 0;
        cp_async_vec<VEC>(img_hi + off, ok ? p0 + (long long)(8 * u) * ld : g, ok ? nr * 4 : 0);
      }
    }
  }
}

template <bool kSplitPass>
__device__ __forceinline__ void fill_dispatch(int vec, char* hi, char* lo, const float* g, long long ld, int src, int r0, int r_end,
                                              int k0, int k_end, int rows, int ptid) {
  if (src == TCG_SRC_K) {
    if (vec == 4) fill_k<4, kSplitPass>(hi, lo, g, ld, r0, r_end, k0, k_end, rows, ptid);
    else if (vec == 2) fill_k<2, kSplitPass>(hi, lo, g, ld, r0, r_end, k0, k_end, rows, ptid);
    else fill_k<1, kSplitPass>(hi, lo, g, ld, r0, r_end, k0, k_end, rows, ptid);
  } else {
    if (vec == 4) fill_mn<4, kSplitPass>(hi, lo, g, ld, r0, r_end, k0, k_end, rows, ptid);
    else if (vec == 2) fill_mn<2, kSplitPass>(hi, lo, g, ld, r0, r_end, k0, k_end, rows, ptid);
    else fill_mn<1, kSplitPass>(hi, lo, g, ld, r0, r_end, k0, k_end, rows, ptid);
  }
}

__global__ void __launch_bounds__(kGemmThreads) tc_gemm_kernel(TcGemmParams p, int vecA, int vecB) {
  using namespace tc;
  GNF_SMEM(char, smem);
  const int BN = p.BN;
  const bool split = p.passes == 3;
  const uint32_t a_bytes = kGemmBM * 128u, b_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (split ? 2u : 1u);
  // stage layout: [A_hi][B_hi]([A_lo][B_lo])
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;                       // [stages]  producers -> MMA   (count 8: one arrive per producer warp)
  uint64_t* empty = bars + kMaxStages;         // [stages]  MMA -> producers   (tcgen05.commit)
  uint64_t* tfull = bars + 2 * kMaxStages;     // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;                // [2]       epilogue -> MMA    (count 4: one arrive per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], kProdThreads / 32); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], kEpiThreads / 32); }
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = (p.M + kGemmBM - 1) / kGemmBM, tiles_n = (p.N + BN - 1) / BN;
  const long long total = (long long)tiles_m * tiles_n * p.splits;

  if (warp >= 5) {
    // ===================== producers =====================
    // Software pipeline with a lag of kLag stages: the copies of chunk `it` are issued (cp.async, one commit group per
    // chunk) before chunk it-kLag is finished (wait for its group, produce the lo tiles, proxy fence, signal the MMA).
    const int kLag = p.stages >= 3 ? 2 : 1;              // the ring must hold the in-flight chunks plus one being consumed
    const int ptid = tid - (kEpiThreads + 32);
    long long it = 0;                                    // running k-chunk counter (stage ring position)
    auto finish = [&](long long j) {                     // chunk j's copies have landed (caller waited on its group)
      const int s = (int)(j % p.stages);
      char* st = smem + (size_t)s * stage_bytes;
      if (split) {
        fill_dispatch<true>(vecA, st, st + a_bytes + b_bytes, nullptr, 0, p.a_src, 0, 0, 0, 0, kGemmBM, ptid);
        fill_dispatch<true>(vecB, st + a_bytes, st + 2 * a_bytes + b_bytes, nullptr, 0, p.b_src, 0, 0, 0, 0, BN, ptid);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
      const int tm = (int)(w % tiles_m), tn = (int)((w / tiles_m) % tiles_n), sp = (int)(w / ((long long)tiles_m * tiles_n));
      const int kbeg = sp * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
      for (int k0 = kbeg; k0 < kend; k0 += kGemmKC, ++it) {
        const int s = (int)(it % p.stages);
        const uint32_t par = (uint32_t)((it / p.stages) & 1);
        mbar_wait(&empty[s], par ^ 1u);
        char* st = smem + (size_t)s * stage_bytes;
        fill_dispatch<false>(vecA, st, nullptr, p.A, p.lda, p.a_src, tm * kGemmBM, p.M, k0, kend, kGemmBM, ptid);
        fill_dispatch<false>(vecB, st + a_bytes, nullptr, p.B, p.ldb, p.b_src, tn * BN, p.N, k0, kend, BN, ptid);
        cp_async_commit();
        if (it >= kLag) {
          if (kLag == 2) cp_async_wait<2>(); else cp_async_wait<1>();
          finish(it - kLag);
        }
      }
    }
    // drain
    cp_async_wait<0>();
    for (long long j = (it > kLag ? it - kLag : 0); j < it; ++j) finish(j);
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(kGemmBM, BN) | (p.a_src == TCG_SRC_MN ? (1u << 15) : 0u) | (p.b_src == TCG_SRC_MN ? (1u << 16) : 0u);
      // per k-step (8 k) descriptor advance: K-major: 32 bytes inside the swizzle atom; MN-major: 8 tile rows = 1024 bytes
      const uint32_t a_step = (p.a_src == TCG_SRC_K) ? 2u : 64u, b_step = (p.b_src == TCG_SRC_K) ? 2u : 64u;
      const bool a_mn = p.a_src == TCG_SRC_MN, b_mn = p.b_src == TCG_SRC_MN;
      long long it = 0;
      int tcount = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x, ++tcount) {
        const int sp = (int)(w / ((long long)tiles_m * tiles_n));
        const int kbeg = sp * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
        const int acc = tcount & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((tcount >> 1) & 1) ^ 1));
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        bool first = true;
        for (int k0 = kbeg; k0 < kend; k0 += kGemmKC, ++it) {
          const int s = (int)(it % p.stages);
          mbar_wait(&full[s], (uint32_t)((it / p.stages) & 1));
          fence_after_sync();
          const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da_hi = make_sw128_desc(st, a_mn), db_hi = make_sw128_desc(st + a_bytes, b_mn);
          const uint64_t da_lo = make_sw128_desc(st + a_bytes + b_bytes, a_mn), db_lo = make_sw128_desc(st + 2 * a_bytes + b_bytes, b_mn);
#pragma unroll
          for (int ks = 0; ks < kGemmKC / 8; ++ks) {
            const uint64_t oa = (uint64_t)(a_step * ks), ob = (uint64_t)(b_step * ks);
            mma_tf32_ss(d_tmem, da_hi + oa, db_hi + ob, idesc, first ? 0u : 1u);
            first = false;
            if (split) {
              mma_tf32_ss(d_tmem, da_lo + oa, db_hi + ob, idesc, 1u);
              mma_tf32_ss(d_tmem, da_hi + oa, db_lo + ob, idesc, 1u);
            }
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[acc]);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int t = tid;                                   // accumulator row / TMEM lane
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    int tcount = 0;
    for (long long w = blockIdx.x; w < total; w += gridDim.x, ++tcount) {
      const int tm = (int)(w % tiles_m), tn = (int)((w / tiles_m) % tiles_n);
      const int acc = tcount & 1;
      mbar_wait(&tfull[acc], (uint32_t)((tcount >> 1) & 1));
      fence_after_sync();
      const int m = tm * kGemmBM + t, n0 = tn * BN;
      const uint32_t src = tmem_base + lane_sel + (uint32_t)acc * 256u;
      uint32_t cur[16], nxt[16];
      tmem_ld16_nowait(src, cur);
      tmem_wait_ld();
      for (int c = 0; c < BN; c += 16) {
        if (c + 16 < BN) tmem_ld16_nowait(src + c + 16, nxt);
        if (m < p.M) {
          const int n = n0 + c;
          if (p.epi == TCG_EPI_BIAS_ACT) {
            float* y = p.C + (long long)m * p.ldc + n;
            const float* bp = p.bias ? p.bias + (long long)(p.bias_period > 1 ? (m % p.bias_period) : 0) * p.bias_ld + n : nullptr;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float r = __uint_as_float(cur[j]);
              if (bp && n + j < p.N) r += __ldg(bp + j);
              o[j] = p.relu ? fmaxf(r, 0.f) : r;
            }
            if (n + 16 <= p.N && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            } else if (n + 16 <= p.N && (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) *reinterpret_cast<float2*>(y + j) = make_float2(o[j], o[j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (n + j < p.N) y[j] = o[j];
            }
          } else if (p.epi == TCG_EPI_MASK) {
            float* y = p.C + (long long)m * p.ldc + n;
            const float* ap = p.act ? p.act + (long long)m * p.ldact + n : nullptr;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n + j < p.N) {
                float r = __uint_as_float(cur[j]);
                if (ap && !(__ldg(ap + j) > 0.f)) r = 0.f;
                y[j] = r;
              }
            }
          } else {
            float* y = p.C + (long long)m * p.ldc + n;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < p.N) atomicAdd(y + j, __uint_as_float(cur[j]));
          }
        }
        if (c + 16 < BN) {
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static inline int vec_width(const float* p, long long ld) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((a & 15) == 0 && (ld % 4) == 0) return 4;
  if ((a & 7) == 0 && (ld % 2) == 0) return 2;
  return 1;
}

static int pick_bn(int N) {
  int best = 64;
  long long best_cost = -1;
  const int cands[7] = {64, 96, 128, 160, 192, 224, 256};
  for (int i = 0; i < 7; ++i) {
    const int bn = cands[i];
    const long long padded = (long long)((N + bn - 1) / bn) * bn;
    // padded width is wasted tensor work; small tiles pay more A re-reads and more per-tile overhead
    const long long cost = padded * 8 + (long long)((N + bn - 1) / bn) * 160;
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

static int launch_tc_gemm(TcGemmParams p, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0) return 0;
  if (p.passes != 1 && p.passes != 3) return fail(GNF_ERR_INVALID, "tensor-core GEMM: passes must be 1 or 3");
  p.BN = pick_bn(p.N);
  // every candidate is a multiple of 32: K-major fills advance 8/16/32 rows per pass, MN-major tiles are 32-row slabs
  const uint32_t stage_bytes = (uint32_t)(kGemmBM + p.BN) * 128u * (p.passes == 3 ? 2u : 1u);
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(GNF_ERR_UNSUPPORTED, "tensor-core GEMM: tile does not fit shared memory");
  p.stages = stages;
  const int tiles = ((p.M + kGemmBM - 1) / kGemmBM) * ((p.N + p.BN - 1) / p.BN);
  int splits = 1;
  if (p.epi == TCG_EPI_ATOMIC) {
    splits = (kNumSMs + tiles - 1) / tiles;
    const int max_splits = (p.K + 8 * kGemmKC - 1) / (8 * kGemmKC);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  const int kchunks = (p.K + kGemmKC - 1) / kGemmKC;
  p.k_per_split = ((kchunks + splits - 1) / splits) * kGemmKC;
  p.splits = (p.K + p.k_per_split - 1) / p.k_per_split;
  if (p.splits < 1) p.splits = 1;
  const long long total = (long long)tiles * p.splits;
  const size_t smem = (size_t)stages * stage_bytes + (2 * kMaxStages + 4) * sizeof(uint64_t) + 16;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
  const int grid = (int)(total < kNumSMs ? total : kNumSMs);
  GNF_LAUNCH(tc_gemm_kernel, grid, kGemmThreads, smem, s, p, vec_width(p.A, p.lda), vec_width(p.B, p.ldb));
  return 0;
}

}  // namespace gnf
using namespace gnf;
#endif

extern "C" {

int gnf_linear_fwd_tc(const float* X, int ldx, const float* W, int ldw, const float* bias, int bias_period, float* Y, int ldy,
                      int M, int N, int K, int relu, int passes, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!X || !W || !Y || M < 0 || N <= 0 || K <= 0 || ldx < K || ldw < K || ldy < N) return fail(GNF_ERR_INVALID, "gnf_linear_fwd_tc: bad arguments");
  TcGemmParams p = {};
  p.A = X; p.lda = ldx; p.a_src = TCG_SRC_K;
  p.B = W; p.ldb = ldw; p.b_src = TCG_SRC_K;
  p.M = M; p.N = N; p.K = K; p.passes = passes;
  p.epi = TCG_EPI_BIAS_ACT; p.C = Y; p.ldc = ldy; p.bias = bias; p.bias_ld = N; p.bias_period = bias_period < 1 ? 1 : bias_period; p.relu = relu;
  if (int e = launch_tc_gemm(p, (cudaStream_t)stream)) return e;
  return check_launch("gnf_linear_fwd_tc");
#endif
}

int gnf_linear_dgrad_tc(const float* dY, int lddy, const float* W, int ldw, const float* act, int ldact, float* dX, int lddx,
                        int M, int N, int K, int passes, gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!dY || !W || !dX || M < 0 || N <= 0 || K <= 0 || lddy < N || ldw < K || lddx < K) return fail(GNF_ERR_INVALID, "gnf_linear_dgrad_tc: bad arguments");
  TcGemmParams p = {};
  p.A = dY; p.lda = lddy; p.a_src = TCG_SRC_K;         // A(m, n): reduction over n, contiguous
  p.B = W; p.ldb = ldw; p.b_src = TCG_SRC_MN;          // B(k_out, n) = W[n*ldw + k_out]: contiguous along the output index
  p.M = M; p.N = K; p.K = N; p.passes = passes;
  p.epi = TCG_EPI_MASK; p.C = dX; p.ldc = lddx; p.act = act; p.ldact = ldact;
  if (int e = launch_tc_gemm(p, (cudaStream_t)stream)) return e;
  return check_launch("gnf_linear_dgrad_tc");
#endif
}

int gnf_linear_wgrad_tc(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, int M, int N, int K, int passes,
                        gnf_stream_t stream) {
#ifdef GNF_EMU
  return gnf::fail(GNF_ERR_UNSUPPORTED, "tensor-core kernels have no host-simulator flavour");
#else
  if (!dY || !X || !dW || M < 0 || N <= 0 || K <= 0 || lddy < N || ldx < K || lddw < K) return fail(GNF_ERR_INVALID, "gnf_linear_wgrad_tc: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (lddw == K) cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), s);
  else for (int n = 0; n < N; ++n) cudaMemsetAsync(dW + (size_t)n * lddw, 0, (size_t)K * sizeof(float), s);
  if (M == 0) return check_launch("gnf_linear_wgrad_tc");
  TcGemmParams p = {};
  p.A = dY; p.lda = lddy; p.a_src = TCG_SRC_MN;        // A(n, m) = dY[m*lddy + n]: contiguous along the output row index
  p.B = X; p.ldb = ldx; p.b_src = TCG_SRC_MN;          // B(k, m) = X[m*ldx + k]
  p.M = N; p.N = K; p.K = M; p.passes = passes;
  p.epi = TCG_EPI_ATOMIC; p.C = dW; p.ldc = lddw;
  if (int e = launch_tc_gemm(p, s)) return e;
  return check_launch("gnf_linear_wgrad_tc");
#endif
}

}  // extern "C"
