// Stage 1 shared by the 3D and 4D paths: scalar field -> low / near bitplanes (+ min / max, + dilated row flags).
// Rows are runs of the contiguous (last) axis; a row index decomposes as ((a*rb + b)*rc + c).
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

struct MinMaxKeys {
  unsigned long long min_key, max_key;     // order-preserving encodings, combined with atomicMin / atomicMax
};

// exact n / d for 32-bit n, d via one 64-bit multiply-high (M = ceil(2^64 / d))
struct FastDiv {
  unsigned long long M;
  unsigned d;
  __host__ void init(unsigned dd) {
    d = dd;
    M = dd <= 1 ? 0ull : (~0ull / dd) + 1ull;
  }
  __device__ __forceinline__ unsigned div(unsigned n) const { return d <= 1 ? n : (unsigned)__umul64hi((unsigned long long)n, M); }
};

__device__ __forceinline__ unsigned long long order_key(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// ------------------------------------------------------------------------------------------------
// Stage 1: field -> low bitplane (the only pass over the whole field; HBM-bound), plus the "any sample
// near the isovalue" flag and optional min/max.  Near words are written with the low words; a word
// that has a near sample also raises the (dilated) row flags that switch later stages to their exact paths.
//   k_bitplane_generic : any shape (rows padded to whole words), one warp converts 32 samples per step.
//   k_bitplane_vec     : rows a multiple of 32 samples: the volume is one linear stream; 128-bit loads.
//   k_bitplane_tma     : same stream staged through shared memory by TMA bulk copies (cp.async.bulk +
//                        mbarrier full/empty ring): one producer lane + 8 consumer warps per CTA; consumers
//                        read the staged samples bank-conflict-free and ballot straight into bit words.
// ------------------------------------------------------------------------------------------------
struct RowGeom {
  uint8_t* rowflag;
  uint32_t* wordflag;                      // optional (3D): per row, one bit per group of 2^wshift words, see flag_rows
  FastDiv divW, divRc, divRb;
  int rb, rc;                              // row = (a*rb + b)*rc + c   (3D: a = 0, b = i, c = j)
  int W, wshift;
};

// word `word` holds a near sample: flag every row whose 3x3(x3) row neighbourhood contains it.  The word flags say
// where in those rows: bit (w >> wshift) for the words w-1, w, w+1 (an edge or voxel of word w-1 reaches sample 0 of
// word w; the exact paths of word w+1 look back at sample 31 of word w), so that later stages take their exact path
// for 3 words of a flagged row, not for all of it.
__device__ __noinline__ void flag_rows(const RowGeom& rg, unsigned word) {
  const unsigned row = rg.divW.div(word);
  const unsigned ab = rg.divRc.div(row), c = row - ab * (unsigned)rg.rc;
  const unsigned a = rg.divRb.div(ab), b = ab - a * (unsigned)rg.rb;
  uint32_t m = 0;
  if (rg.wordflag) {
    const int w = (int)(word - row * (unsigned)rg.W);
    for (int q = max(w - 1, 0); q <= min(w + 1, rg.W - 1); ++q) m |= 1u << (q >> rg.wshift);
  }
  for (int da = 0; da < 3; ++da)
    for (int db = 0; db < 3; ++db)
      for (int dc = 0; dc < 3; ++dc)
        if ((int)a - da >= 0 && (int)b - db >= 0 && (int)c - dc >= 0) {
          const size_t r = ((size_t)(a - da) * rg.rb + (b - db)) * rg.rc + (c - dc);
          rg.rowflag[r] = 1;
          if (rg.wordflag) atomicOr(&rg.wordflag[r], m);
        }
}

template <typename T>
__device__ __forceinline__ void minmax_commit(T mn, T mx, bool anynear, MinMaxKeys* ctr) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  (void)anynear;
  if (lane_id() == 0) {
    if (mn <= mx) {
      atomicMin(&ctr->min_key, order_key((double)mn));
      atomicMax(&ctr->max_key, order_key((double)mx));
    }
  }
}

template <typename T, int UNROLL, bool MINMAX>
__global__ void __launch_bounds__(256) k_bitplane_generic(const T* __restrict__ f, unsigned nrows, int n2, int W,
                                                          T thr, T near_lo, T near_hi, uint32_t* __restrict__ bits,
                                                          uint32_t* __restrict__ nbits, RowGeom rg, MinMaxKeys* ctr) {
  ctr_pdl_enter();
  const unsigned lane = lane_id();
  const unsigned nwords = nrows * (unsigned)W;
  const unsigned warps = gridDim.x * (blockDim.x >> 5);
  unsigned g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  T mn = INFINITY, mx = -INFINITY;
  bool anynear = false;
  for (; g < nwords; g += warps * UNROLL) {
    T val[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      unsigned gu = g + (unsigned)u * warps;
      ok[u] = false;
      val[u] = (T)0;
      if (gu < nwords) {
        unsigned row = gu / (unsigned)W;
        int k = (int)(gu - row * (unsigned)W) * 32 + (int)lane;
        if (k < n2) {
          ok[u] = true;
          val[u] = __ldg(f + (size_t)row * n2 + k);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      unsigned gu = g + (unsigned)u * warps;
      if (gu < nwords) {                       // warp-uniform
        unsigned wl = __ballot_sync(0xffffffffu, ok[u] && (val[u] < thr));
        unsigned wn = __ballot_sync(0xffffffffu, ok[u] && (val[u] >= near_lo) && (val[u] <= near_hi));
        if (lane == 0) {
          bits[gu] = wl;
          nbits[gu] = wn;
          if (wn) flag_rows(rg, gu);
        }
        if (ok[u]) {
          if (MINMAX) {
            mn = fmin(mn, val[u]);             // fmin/fmax ignore NaN
            mx = fmax(mx, val[u]);
          }
        }
      }
    }
  }
  minmax_commit(mn, mx, anynear, ctr);
}

// 16 bytes = VEC samples per lane; one warp-wide 128-bit load covers 32*VEC samples = VEC bit words.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int VEC = 4;
  typedef float4 type;
  __device__ static __forceinline__ void get(const float4& v, float out[4]) { out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
};
template <>
struct Vec16<double> {
  static constexpr int VEC = 2;
  typedef double2 type;
  __device__ static __forceinline__ void get(const double2& v, double out[2]) { out[0] = v.x; out[1] = v.y; }
};

template <typename T, int UNROLL, bool MINMAX>
__global__ void __launch_bounds__(256) k_bitplane_vec(const T* __restrict__ f, size_t nsamp, T thr, T near_lo, T near_hi,
                                                      uint32_t* __restrict__ bits, uint32_t* __restrict__ nbits, RowGeom rg,
                                                      MinMaxKeys* ctr) {
  typedef typename Vec16<T>::type V;
  constexpr int VEC = Vec16<T>::VEC;
  constexpr int GROUP = 32 / VEC;               // lanes per output word
  ctr_pdl_enter();
  const unsigned lane = lane_id();
  const size_t nchunks = (nsamp + 32 * VEC - 1) / (32 * VEC);
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  size_t c = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  T mn = INFINITY, mx = -INFINITY;
  bool anynear = false;
  const V* fv = reinterpret_cast<const V*>(f);
  for (; c < nchunks; c += warps * UNROLL) {
    V v[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      size_t cu = c + (size_t)u * warps;
      size_t s0 = (cu * 32 + lane) * VEC;
      ok[u] = cu < nchunks && s0 < nsamp;
      if (ok[u]) v[u] = __ldcs(fv + cu * 32 + lane);       // streaming load (evict-first)
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      size_t cu = c + (size_t)u * warps;
      if (cu >= nchunks) continue;                          // warp-uniform
      T x[VEC];
      Vec16<T>::get(v[u], x);
      unsigned lo = 0, nr = 0;
      if (ok[u]) {
        T m4 = x[0], M4 = x[0];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          lo |= (x[q] < thr ? 1u : 0u) << q;
          m4 = fmin(m4, x[q]);
          M4 = fmax(M4, x[q]);
        }
        if (MINMAX) {
          mn = fmin(mn, m4);
          mx = fmax(mx, M4);
        }
        if (!(M4 < near_lo || m4 > near_hi)) {              // rarely taken: some sample may be in the hull
#pragma unroll
          for (int q = 0; q < VEC; ++q) nr |= ((x[q] >= near_lo) && (x[q] <= near_hi) ? 1u : 0u) << q;
        }
      }
      unsigned wl = lo << (VEC * (lane & (GROUP - 1)));
#pragma unroll
      for (int o = 1; o < GROUP; o <<= 1) wl |= __shfl_xor_sync(0xffffffffu, wl, o);
      unsigned wn = 0;
      if (__any_sync(0xffffffffu, nr != 0)) {               // rare
        wn = nr << (VEC * (lane & (GROUP - 1)));
#pragma unroll
        for (int o = 1; o < GROUP; o <<= 1) wn |= __shfl_xor_sync(0xffffffffu, wn, o);
      }
      if ((lane & (GROUP - 1)) == 0 && ok[u]) {
        const size_t word = cu * VEC + lane / GROUP;
        bits[word] = wl;
        nbits[word] = wn;
        if (wn) flag_rows(rg, (unsigned)word);
      }
    }
  }
  minmax_commit(mn, mx, anynear, ctr);
}

// ---- TMA (bulk async copy) variant ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// The field is read exactly once: its lines enter L2 marked evict-first, so that they do not push out the bitplanes this
// kernel writes (32 MiB that stage 2 reads right away) nor the dirty lines of the previous extraction's mesh.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

constexpr int TMA_STAGES_DEFAULT = 3;              // x 16 KiB per CTA; 3 CTAs per SM (measured best of 1..6 x 2..6)
constexpr int TMA_CHUNK = 16384;                 // bytes per stage
constexpr int TMA_CONSUMER_WARPS = 8;

template <typename T, bool MINMAX>
__global__ void __launch_bounds__((TMA_CONSUMER_WARPS + 1) * 32) k_bitplane_tma(const T* __restrict__ f, size_t nbytes,
                                                                                  T thr, T near_lo, T near_hi,
                                                                                  uint32_t* __restrict__ bits,
                                                                                  uint32_t* __restrict__ nbits, RowGeom rg,
                                                                                  MinMaxKeys* ctr, int TMA_STAGES, int l2_hint) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)TMA_STAGES * TMA_CHUNK);
  uint64_t* empty = full + TMA_STAGES;
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TMA_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nchunks = (nbytes + TMA_CHUNK - 1) / TMA_CHUNK;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(f);
  if (warp == TMA_CONSUMER_WARPS) {
    if (lane == 0) {
      unsigned it = 0;
      const uint64_t pol = l2_evict_first_policy();
      for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
        const int s = it % TMA_STAGES;
        const unsigned ph = (it / TMA_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);                           // first round passes immediately
        size_t off = c * (size_t)TMA_CHUNK;
        unsigned bytes = (unsigned)((nbytes - off) < (size_t)TMA_CHUNK ? (nbytes - off) : (size_t)TMA_CHUNK);
        mbar_expect_tx(&full[s], bytes);
        if (l2_hint) tma_bulk_g2s_hint(smem + (size_t)s * TMA_CHUNK, src + off, bytes, &full[s], pol);
        else tma_bulk_g2s(smem + (size_t)s * TMA_CHUNK, src + off, bytes, &full[s]);
      }
    }
    return;
  }
  constexpr int NW = TMA_CHUNK / (32 * (int)sizeof(T));          // bit words per stage
  constexpr int WPW = NW / TMA_CONSUMER_WARPS;                   // words per warp per stage (<= 32)
  static_assert(WPW >= 1 && WPW <= 32, "stage geometry");
  uint32_t* swords = reinterpret_cast<uint32_t*>(empty + TMA_STAGES) + warp * 32;   // per-warp word staging
  // conservative centre / radius form of the hull test: |x - vc| <= vr  (superset of [near_lo, near_hi])
  const T vc = (T)0.5 * near_lo + (T)0.5 * near_hi;
  const T vr = (near_hi - near_lo) * (T)0.55 + fabs(vc) * (T)1e-6 + (T)1e-30;
  T mn = INFINITY, mx = -INFINITY;
  const bool anynear = false;
  unsigned it = 0;
  // The producer above starts fetching at once: the field is complete before the kernel in front of this one passed its
  // own wait (common.cuh, rule 2).  The consumers wait here, in front of their first write: the flags they set are
  // zeroed by that kernel, and the bit planes may still be read by the previous run's last kernel.
  ctr_pdl_enter();
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int s = it % TMA_STAGES;
    const unsigned ph = (it / TMA_STAGES) & 1u;
    const size_t off = c * (size_t)TMA_CHUNK;
    const unsigned bytes = (unsigned)((nbytes - off) < (size_t)TMA_CHUNK ? (nbytes - off) : (size_t)TMA_CHUNK);
    const unsigned words_here = bytes / (32u * (unsigned)sizeof(T));
    const T* sv = reinterpret_cast<const T*>(smem + (size_t)s * TMA_CHUNK) + (size_t)warp * WPW * 32;
    mbar_wait(&full[s], ph);
    T val[WPW];
#pragma unroll
    for (int q = 0; q < WPW; ++q) val[q] = sv[q * 32 + lane];    // conflict-free: lane <-> bank
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);                       // stage is in registers: release it early
    const unsigned wbase = warp * WPW;
    const size_t word0 = off / (32 * sizeof(T)) + wbase;
    if (words_here < (unsigned)NW) {
      // partial last chunk: samples past the end are stale shared memory; neutralise them
#pragma unroll
      for (int q = 0; q < WPW; ++q)
        if (wbase + q >= words_here) val[q] = INFINITY;
    }
    T dist = INFINITY;
#pragma unroll
    for (int q = 0; q < WPW; ++q) {
      const unsigned wl = __ballot_sync(0xffffffffu, val[q] < thr);
      if (lane == 0) swords[q] = wl;
      dist = fmin(dist, fabs(val[q] - vc));                     // NaN never wins: not near
      if (MINMAX) {
        mn = fmin(mn, val[q]);
        mx = fmax(mx, val[q]);
      }
    }
    unsigned mine_n = 0;
    if (__any_sync(0xffffffffu, dist <= vr)) {                   // rare: materialise the near words of this warp
#pragma unroll
      for (int q = 0; q < WPW; ++q) {
        const unsigned wn = __ballot_sync(0xffffffffu, (val[q] >= near_lo) && (val[q] <= near_hi));
        if ((int)lane == q) {
          mine_n = wn;
          if (wn) flag_rows(rg, (unsigned)(word0 + q));
        }
      }
    }
    __syncwarp();
    if ((int)lane < WPW && wbase + lane < words_here) {
      bits[word0 + lane] = swords[lane];
      nbits[word0 + lane] = mine_n;
    }
    __syncwarp();
  }
  minmax_commit(mn, mx, anynear, ctr);
}


double key_to_double(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
  memcpy(&d, &u, 8);
  return d;
}

template <typename T>
void thresholds(double v, T& thr, T& nlo, T& nhi);

template <>
void thresholds<double>(double v, double& thr, double& nlo, double& nhi) {
  thr = v;
  double r = (1e-8 + 1e-5 * fabs(v)) * 1.0001;
  nlo = v - r;
  nhi = v + r;
}
template <>
void thresholds<float>(double v, float& thr, float& nlo, float& nhi) {
  thr = (float)v;                               // f < v  <=>  f < thr, thr = smallest float >= v
  if ((double)thr < v) thr = nextafterf(thr, INFINITY);
  double r = (1e-8 + 1e-5 * fabs(v)) * 1.0001;
  nlo = (float)(v - r);
  if ((double)nlo > v - r) nlo = nextafterf(nlo, -INFINITY);
  nhi = (float)(v + r);
  if ((double)nhi < v + r) nhi = nextafterf(nhi, INFINITY);
}

enum BitplaneKind { BP_AUTO = 0, BP_GENERIC = 1, BP_VEC = 2, BP_TMA = 3 };

BitplaneKind bitplane_choice() {
  const char* e = getenv("CTR_BITPLANE");
  if (!e) return BP_AUTO;
  if (!strcmp(e, "generic")) return BP_GENERIC;
  if (!strcmp(e, "vec")) return BP_VEC;
  if (!strcmp(e, "tma")) return BP_TMA;
  return BP_AUTO;
}

template <typename T, bool MINMAX>
int launch_bitplane(ctr_ctx* ctx, const T* dfield, unsigned nrows, int n2, int W, int rb, int rc, double iso,
                    uint32_t* bits, uint32_t* nbits, uint8_t* rowflag, MinMaxKeys* dctr, bool clear_rowflag = true,
                    uint32_t* wordflag = nullptr) {
  cudaStream_t st = ctx->stream;
  T thr, nlo, nhi;
  thresholds<T>(iso, thr, nlo, nhi);
  const size_t nsamp = (size_t)nrows * n2;
  const size_t nwords = (size_t)nrows * W;
  const bool linear_ok = (n2 % 32 == 0) && (((uintptr_t)dfield) % 16 == 0);
  BitplaneKind kind = bitplane_choice();
  if (kind == BP_AUTO) kind = linear_ok ? BP_TMA : BP_GENERIC;
  if (!linear_ok) kind = BP_GENERIC;
  RowGeom rg;
  rg.rowflag = rowflag;
  rg.wordflag = wordflag;
  rg.W = W;
  rg.wshift = 0;
  while ((W >> rg.wshift) > 32) ++rg.wshift;
  rg.divW.init((unsigned)W);
  rg.divRc.init((unsigned)rc);
  rg.divRb.init((unsigned)rb);
  rg.rb = rb;
  rg.rc = rc;
  if (clear_rowflag) CTR_CUDA(ctx, cudaMemsetAsync(rg.rowflag, 0, (size_t)nrows, st));
  if (kind == BP_TMA) {
    static const int TMA_STAGES = getenv("CTR_BP_STAGES") ? atoi(getenv("CTR_BP_STAGES")) : TMA_STAGES_DEFAULT;
    static const int ctas_per_sm = getenv("CTR_BP_CTAS") ? atoi(getenv("CTR_BP_CTAS")) : 3;
    static const int l2_hint = getenv("CTR_BP_L2HINT") ? atoi(getenv("CTR_BP_L2HINT")) : 1;   // 0.428 -> 0.418 ms per 512^3 step
    const int smem = TMA_STAGES * TMA_CHUNK + 2 * TMA_STAGES * 8 + TMA_CONSUMER_WARPS * 32 * 4 + 64;
    // this header is compiled into two translation units (anonymous namespaces): each has its own kernel instances
    const int ti = CTR_BP_ATTR_BASE + (sizeof(T) == 4 ? 0 : 1) + (MINMAX ? 2 : 0);
    if (!(ctx->attr_mask & (1u << ti))) {
      CTR_CUDA(ctx, cudaFuncSetAttribute(k_bitplane_tma<T, MINMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TMA_CHUNK + 1024));
      ctx->attr_mask |= 1u << ti;
    }
    const size_t nbytes = nsamp * sizeof(T);
    const size_t nchunks = (nbytes + TMA_CHUNK - 1) / TMA_CHUNK;
    int blocks = (int)std::min<size_t>(nchunks, (size_t)ctx->sm_count * ctas_per_sm);
    ctr_launch_dep(k_bitplane_tma<T, MINMAX>, blocks, (TMA_CONSUMER_WARPS + 1) * 32, smem, st, dfield, nbytes, thr, nlo, nhi, bits, nbits, rg,
                   dctr, (int)TMA_STAGES, (int)l2_hint);
  } else if (kind == BP_VEC) {
    const size_t nchunks = (nsamp + 32 * Vec16<T>::VEC - 1) / (32 * Vec16<T>::VEC);
    size_t need = (nchunks + 8 * 4 - 1) / (8 * 4);
    int blocks = (int)std::min<size_t>(std::max<size_t>(need, 1), (size_t)ctx->sm_count * 8);
    k_bitplane_vec<T, 4, MINMAX><<<blocks, 256, 0, st>>>(dfield, nsamp, thr, nlo, nhi, bits, nbits, rg, dctr);
  } else {
    size_t need = (nwords + 8 * 4 - 1) / (8 * 4);
    int blocks = (int)std::min<size_t>(std::max<size_t>(need, 1), (size_t)ctx->sm_count * 8);
    k_bitplane_generic<T, 4, MINMAX><<<blocks, 256, 0, st>>>(dfield, nrows, n2, W, thr, nlo, nhi, bits, nbits, rg, dctr);
  }
  ctx->launches++;
  CTR_CUDA(ctx, cudaGetLastError());
  return 0;
}


}  // namespace
