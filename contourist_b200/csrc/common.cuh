// Shared plumbing of the engine: context, grow-only device buffers, error handling, warp helpers,
// decoupled-lookback tile status.  Host language above this is Python (ctypes), see engine.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/contourist_b200.h"

#define CTR_NSTAGE 8

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct ctr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t stream2 = nullptr;      // side stream: stage 4 runs beside stage 3
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  int64_t launches = 0;
  bool timing = false;
  cudaEvent_t ev[CTR_NSTAGE + 1] = {};
  bool ev_set[CTR_NSTAGE + 1] = {};
  float stage_ms[CTR_NSTAGE] = {};
  int sm_count = 148;

  // staging + shared scratch
  DevBuf field;              // device copy of a host field
  DevBuf bits, nbits;        // bitplanes: low (f < v) and near (conservative allclose hull)
  DevBuf vbase, tbase;       // per-word exclusive offsets (uint32)
  DevBuf list_v, list_t;     // compacted active word lists (uint32 word index)
  DevBuf tile_state;         // decoupled-lookback status words
  DevBuf counters;           // small block of device counters (see each path)
  DevBuf wmask;              // 3D: per-word (active-owner, emitting-voxel) masks
  DevBuf wdir;               // 3D: per-word direction prefix (dirpack) of the vertex numbering
  DevBuf vox_tab;            // 3D: corner bits -> triangle list of a voxel (256 x 12 words)
  void* counters_host = nullptr;  // pinned mirror

  // outputs (3D)
  DevBuf verts, normals, tris, keys, lowmin, cells, codes;
  // last-run descriptors
  int last_kind = 0;         // 0 none, 3 = mt3d, 2 = mt2d, 4 = mp4d
  uint32_t last_flags = 0;
  int64_t last_counts[8] = {};
  unsigned char last3_params[160] = {};   // parameters of the last completed ctr_mt3d_run
  // 3D: an extraction enqueued by ctr_mt3d_enqueue and not yet finished
  bool pending3 = false;
  alignas(8) unsigned char pending3_params[160] = {};
  // 3D: capacities of the output pools / work lists kept from earlier runs, last list lengths, launch coverage
  size_t spec_v = 0, spec_t = 0, spec_own = 0, spec_cell = 0, last_own = 0, last_cell = 0, cover_own = 0, cover_cell = 0;
  size_t spec_w = 0, last_w = 0, cover_w = 0;   // interesting-word list of the split stage 2
  DevBuf wlist;

  // 2D / 4D extra outputs are declared in their own translation units via these generic slots
  DevBuf aux[32];
};

static inline int ctr_fail(ctr_ctx* ctx, int code, const char* what, const char* detail = "") {
  if (ctx) {
    ctx->err = std::string(what) + (detail && detail[0] ? ": " : "") + (detail ? detail : "");
  }
  return code;
}

#define CTR_CUDA(ctx, call)                                                                   \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      char b__[256];                                                                          \
      snprintf(b__, sizeof b__, "%s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__);  \
      return ctr_fail(ctx, e__ == cudaErrorMemoryAllocation ? CTR_ERR_OOM : CTR_ERR_CUDA,     \
                      "CUDA error", b__);                                                     \
    }                                                                                         \
  } while (0)

// grow-only allocation with 25% head-room so repeated runs of similar size never reallocate
static inline int ctr_ensure(ctr_ctx* ctx, DevBuf& b, size_t bytes, bool exact = false) {
  if (bytes <= b.cap && b.p) return 0;
  if (b.p) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = exact ? bytes : bytes + bytes / 4;
  if (want < 256) want = 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess && want != bytes) {
    want = bytes < 256 ? 256 : bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    char m[128];
    snprintf(m, sizeof m, "cudaMalloc of %zu bytes failed", want);
    return ctr_fail(ctx, CTR_ERR_OOM, m);
  }
  b.cap = want;
  return 0;
}

static inline void ctr_stage_mark(ctr_ctx* ctx, int idx) {
  if (!ctx->timing) return;
  if (!ctx->ev[idx]) cudaEventCreate(&ctx->ev[idx]);
  cudaEventRecord(ctx->ev[idx], ctx->stream);
  ctx->ev_set[idx] = true;
}

// ------------------------------------------------------------------------------------------ device
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ unsigned long long warp_incl_scan_u64(unsigned long long v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (unsigned)o) v += n;
  }
  return v;
}

__device__ __forceinline__ unsigned warp_incl_scan_u32(unsigned v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (unsigned)o) v += n;
  }
  return v;
}

// Decoupled look-back (Merrill & Garland) over tiles handed out by ticket.  One 64-bit status word per
// tile per scan: bits 63..62 = state (0 invalid, 1 aggregate, 2 inclusive prefix), low 62 bits = value
// (two packed 31-bit sums).  A single word carries state and value, so no fence is needed between them.
#define CTR_LB_AGG (1ull << 62)
#define CTR_LB_INC (2ull << 62)
#define CTR_LB_VAL ((1ull << 62) - 1)

__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by one full warp.  Returns the exclusive prefix of `tile` (sum of aggregates of tiles < tile)
// and publishes the inclusive prefix.  status has one word per tile, zero-initialised.
// The two halves can be called apart: publish the aggregate as soon as it is known, do whatever does not need the
// prefix, then resolve it -- by then the predecessors have usually published theirs and nobody polls.
__device__ __forceinline__ void lb_publish(unsigned long long* status, int tile, unsigned long long aggregate) {
  if (lane_id() == 0) lb_store(&status[tile], (tile == 0 ? CTR_LB_INC : CTR_LB_AGG) | aggregate);
}

__device__ __forceinline__ unsigned long long lb_resolve(unsigned long long* status, int tile, unsigned long long aggregate) {
  const unsigned lane = lane_id();
  if (tile == 0) return 0ull;
  unsigned long long excl = 0;
  int idx = tile - 1;
  while (true) {
    int my = idx - (int)lane;
    unsigned long long s;
    if (my >= 0) {
      do {
        s = lb_load(&status[my]);
      } while ((s >> 62) == 0ull);
    } else {
      s = CTR_LB_INC;  // virtual tile before the first: inclusive prefix 0
    }
    unsigned inc = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
    unsigned long long v = s & CTR_LB_VAL;
    if (inc) {
      int first = __ffs(inc) - 1;
      if ((int)lane > first) v = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (inc) break;
    idx -= 32;
  }
  if (lane == 0) lb_store(&status[tile], CTR_LB_INC | (excl + aggregate));
  return excl;
}

__device__ __forceinline__ unsigned long long lb_lookback(unsigned long long* status, int tile,
                                                          unsigned long long aggregate) {
  lb_publish(status, tile, aggregate);
  return lb_resolve(status, tile, aggregate);
}

// Block-wide look-back over two status arrays at once.  Every thread of the block polls one predecessor, so a window
// of THREADS tiles is resolved per iteration -- with a few hundred tiles in flight the warp version above needs ~10
// serial L2 round trips while the rest of the block idles at a barrier; this one needs one or two.
// Called by all threads; returns the exclusive prefixes of `tile` and publishes the inclusive ones.
template <int THREADS>
__device__ __forceinline__ void lb_lookback2_block(unsigned long long* st_a, unsigned long long* st_b, int tile,
                                                   unsigned long long agg_a, unsigned long long agg_b,
                                                   unsigned long long& excl_a, unsigned long long& excl_b) {
  constexpr int NW = THREADS / 32;
  __shared__ int s_first[2][NW];
  __shared__ unsigned long long s_sum[2][NW];
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  excl_a = excl_b = 0ull;
  if (tile == 0) {                                   // block-uniform
    if (tid == 0) {
      lb_store(&st_a[0], CTR_LB_INC | agg_a);
      lb_store(&st_b[0], CTR_LB_INC | agg_b);
    }
    return;
  }
  if (tid == 0) {
    lb_store(&st_a[tile], CTR_LB_AGG | agg_a);
    lb_store(&st_b[tile], CTR_LB_AGG | agg_b);
  }
  bool done_a = false, done_b = false;
  for (int idx = tile - 1;; idx -= THREADS) {
    const int my = idx - tid;
    unsigned long long sa = CTR_LB_INC, sb = CTR_LB_INC;   // virtual tiles before the first: inclusive prefix 0
    if (my >= 0) {
      if (!done_a) do { sa = lb_load(&st_a[my]); } while ((sa >> 62) == 0ull);
      if (!done_b) do { sb = lb_load(&st_b[my]); } while ((sb >> 62) == 0ull);
    }
    const unsigned ia = __ballot_sync(0xffffffffu, (sa >> 62) == 2ull), ib = __ballot_sync(0xffffffffu, (sb >> 62) == 2ull);
    if (lane == 0) {
      s_first[0][warp] = ia ? warp * 32 + (__ffs(ia) - 1) : THREADS;
      s_first[1][warp] = ib ? warp * 32 + (__ffs(ib) - 1) : THREADS;
    }
    __syncthreads();
    int fa = THREADS, fb = THREADS;                  // nearest inclusive predecessor in this window (thread index)
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      fa = min(fa, s_first[0][q]);
      fb = min(fb, s_first[1][q]);
    }
    unsigned long long va = (!done_a && tid <= fa) ? (sa & CTR_LB_VAL) : 0ull;
    unsigned long long vb = (!done_b && tid <= fb) ? (sb & CTR_LB_VAL) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      va += __shfl_xor_sync(0xffffffffu, va, o);
      vb += __shfl_xor_sync(0xffffffffu, vb, o);
    }
    if (lane == 0) {
      s_sum[0][warp] = va;
      s_sum[1][warp] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      excl_a += s_sum[0][q];
      excl_b += s_sum[1][q];
    }
    done_a = done_a || fa < THREADS;
    done_b = done_b || fb < THREADS;
    __syncthreads();                                 // s_first / s_sum are rewritten by the next iteration
    if (done_a && done_b) break;
  }
  if (tid == 0) {
    lb_store(&st_a[tile], CTR_LB_INC | (excl_a + agg_a));
    lb_store(&st_b[tile], CTR_LB_INC | (excl_b + agg_b));
  }
}

// Single-array wide look-back: warp 0 polls THREADS predecessors per iteration (THREADS/32 loads in flight per lane),
// the other warps park at the barrier instead of spinning.  Called by all threads.
template <int THREADS>
__device__ __forceinline__ unsigned long long lb_lookback_block(unsigned long long* st, int tile, unsigned long long agg) {
  constexpr int PER = THREADS / 32;
  __shared__ unsigned long long s_excl;
  const int tid = (int)threadIdx.x, lane = tid & 31;
  if (tile == 0) {                                   // block-uniform
    if (tid == 0) lb_store(&st[0], CTR_LB_INC | agg);
    return 0ull;
  }
  if (tid < 32) {
    if (lane == 0) lb_store(&st[tile], CTR_LB_AGG | agg);
    unsigned long long excl = 0ull;
    for (int idx = tile - 1;; idx -= THREADS) {
      // lane l looks at predecessors idx - (l*PER + q), q = 0..PER-1 (nearest first)
      unsigned long long v[PER];
      bool all_valid;
      do {
        all_valid = true;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int my = idx - (lane * PER + q);
          v[q] = my >= 0 ? lb_load(&st[my]) : CTR_LB_INC;
          all_valid = all_valid && (v[q] >> 62) != 0ull;
        }
        if (!all_valid) __nanosleep(64);
      } while (!all_valid);
      // sum up to and including the nearest inclusive entry
      unsigned long long sum = 0ull;
      bool found = false;
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        if (!found) sum += v[q] & CTR_LB_VAL;
        found = found || (v[q] >> 62) == 2ull;
      }
      const unsigned fm = __ballot_sync(0xffffffffu, found);
      if (fm) {
        const int first = __ffs(fm) - 1;
        if (lane > first) sum = 0ull;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      excl += sum;
      if (fm) break;
    }
    if (lane == 0) {
      lb_store(&st[tile], CTR_LB_INC | (excl + agg));
      s_excl = excl;
    }
  }
  __syncthreads();
  const unsigned long long e = s_excl;
  __syncthreads();
  return e;
}

#endif  // __CUDACC__
