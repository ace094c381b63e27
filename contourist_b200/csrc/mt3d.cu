// 3D marching tetrahedra on sm_100a -- bitplane-first design.
//
//   k_bitplane   : stream the scalar field ONCE (the only pass that touches all of it), write two bitplanes
//                  (1 bit/sample each): low = f < v (tetrahedral.py:572) and near = conservative hull of the
//                  two np.allclose tests (tetrahedral.py:391,576); min/max (grid_field.py:79-80).
//   k_count_scan : from the bitplanes (L2 resident), per 32-voxel word: triangle count, distinct-edge
//                  (vertex) count, crossing count; fused single-pass decoupled-lookback exclusive scan
//                  -> per-word output offsets + compacted active-word lists.
//   k_emit_verts : per active owner word: interpolate edge crossings (tetrahedral.py:471-487), gradient
//                  normals, world transform (grid_field.py:89-93); vertex id = rank of the edge key
//                  (the reference's dict dedup, tetrahedral.py:184-188, as a perfect hash).
//   k_emit_tris  : per active voxel word: 6 Kuhn tets per voxel (tetrahedral.py:32-39,554-595) ->
//                  triangles of vertex ids, wound so the normal points to the high side.
//   k_codes      : optional parity output: (voxel, 30-bit case code) recomputed from the raw samples.
//
// Exactness: voxels / tets / edges whose outcome could depend on an np.allclose test are detected from the
// near bitplane and re-evaluated from the samples in fp64 (cell_emit_exact / edge_used_exact); everything
// else is decided by bit logic.  There is no CPU fallback anywhere.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "bitplane.cuh"
#include "tables.h"

namespace {

__constant__ uint8_t c_tri_n[6][16];
__constant__ uint8_t c_tri_e[6][16][6];
__constant__ uint8_t c_edge_s[19];
__constant__ uint8_t c_edge_d[19];
__constant__ uint8_t c_tetmask[8][8];
__constant__ uint8_t c_tet[6][4];

struct Counters {                    // device counter block (mirrored to pinned host memory)
  // stage 1 (bitplane)
  unsigned long long min_key;        // order-preserving encoding of fmin
  unsigned long long max_key;        // order-preserving encoding of fmax
  // stage 2 (count + scan)
  unsigned long long n_cells;        // emitting voxels
  unsigned long long n_cross;        // strict crossings
  unsigned long long total_vt;       // packed totals over the scanned range: T << 31 | V
  unsigned long long v_emit;         // vertex count at the start of plane i_hi (vertices this call emits)
  unsigned int ticket;
  unsigned int n_codes;              // slots handed out by k_codes
  unsigned int claim_v, claim_t;     // next unclaimed segment batch of the persistent stage 3 / 4 kernels
};



template <typename T>
struct Grid {
  const T* f;
  const uint32_t* bits;
  const uint32_t* nbits;
  int n0, n1, n2, W;
  int i_lo, i_hi, i_hiv;             // emit planes [i_lo, i_hi); scanned owner planes [i_lo, i_hiv)
  long long plane_offset;
  double v;                          // isovalue
  double tolv;                       // 1e-8 + 1e-5*|v|
  // allclose handling is local: rowflag[i*n1+j] != 0 iff some sample of rows (i..i+2, j..j+2) lies inside the
  // conservative allclose hull; kernels read it (`near`) before touching a word of row (i, j).
  const uint8_t* rowflag;
  FastDiv divW, divN1;
  __device__ __forceinline__ void word_coords(unsigned gw, int& i, int& j, int& w) const {
    unsigned row = divW.div(gw);
    w = (int)(gw - row * (unsigned)W);
    unsigned ii = divN1.div(row);
    i = (int)ii;
    j = (int)(row - ii * (unsigned)n1);
  }
};


// ------------------------------------------------------------------------------------------------
// Exact (fp64) evaluation from the samples -- only for voxels/edges flagged by the near bitplane.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool near_a(double f, double v) {   // np.allclose(value, f): tetrahedral.py:391
  return fabs(v - f) <= __dadd_rn(1e-8, __dmul_rn(1e-5, fabs(f)));
}

template <typename T>
__device__ __forceinline__ double sample(const Grid<T>& g, int i, int j, int k) {
  return (double)g.f[((long long)i * g.n1 + j) * g.n2 + k];
}

// 6-bit mask of the tets of voxel (i,j,k) that emit triangles; also returns the 30-bit case code.
template <typename T>
__device__ __noinline__ unsigned cell_emit_exact(const Grid<T>& g, int i, int j, int k, unsigned* code_out) {
  if (code_out) *code_out = 0;
  if (i < 0 || j < 0 || k < 0 || i >= g.n0 - 1 || j >= g.n1 - 1 || k >= g.n2 - 1) return 0;
  double fv[8];
  bool all_a = true;
  unsigned low = 0, nb = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    fv[c] = sample(g, i + ((c >> 2) & 1), j + ((c >> 1) & 1), k + (c & 1));
    all_a = all_a && near_a(fv[c], g.v);
    low |= (fv[c] < g.v ? 1u : 0u) << c;
    nb |= (fabs(fv[c] - g.v) <= g.tolv ? 1u : 0u) << c;       // np.allclose(values, value): tetrahedral.py:576
  }
  unsigned emit = 0, code = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    unsigned m = 0, alln = 1;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int c = c_tet[t][b];
      m |= ((low >> c) & 1u) << b;
      alln &= (nb >> c) & 1u;
    }
    code |= (m | (alln << 4)) << (5 * t);
    if (m != 0 && m != 15 && !alln) emit |= 1u << t;
  }
  if (code_out) *code_out = code;
  // border_voxel (tetrahedral.py:391-394): the min<=v<=max half is implied by any crossing tet
  return all_a ? 0u : emit;
}

// Is the crossing edge p -> p+d used by any emitted triangle?  (OR over the voxels / tets that contain it.)
template <typename T>
__device__ __noinline__ bool edge_used_exact(const Grid<T>& g, int i, int j, int k, int d) {
  for (int s = 0; s < 8; ++s) {
    if (s & d) continue;
    unsigned tm = c_tetmask[d][s];
    if (!tm) continue;
    unsigned e = cell_emit_exact(g, i - ((s >> 2) & 1), j - ((s >> 1) & 1), k - (s & 1), nullptr);
    if (e & tm) return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
// Bit-sliced neighbourhood of one word: rows (i+a, j+b), a,b in {0,1}; P = bits at k, S = bits at k+1.
// ------------------------------------------------------------------------------------------------
struct Planes {
  uint32_t P[4];    // index a*2+b
  uint32_t S[4];
  uint32_t kpt;     // bits with k < n2
  uint32_t kp1;     // bits with k+1 < n2
  bool has_i1, has_j1;
};

__device__ __forceinline__ uint32_t low_mask(int n) {   // n lowest bits set, n in [0,32]
  return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

template <typename T>
__device__ __forceinline__ void load_planes(const Grid<T>& g, const uint32_t* __restrict__ plane, int i, int j, int w,
                                            Planes& pl) {
  pl.has_i1 = (i + 1 < g.n0);
  pl.has_j1 = (j + 1 < g.n1);
  int rem = g.n2 - w * 32;                   // samples from this word's first bit to the end of the row
  pl.kpt = low_mask(rem < 0 ? 0 : rem);
  pl.kp1 = low_mask(rem - 1 < 0 ? 0 : rem - 1);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      bool ok = (a == 0 || pl.has_i1) && (b == 0 || pl.has_j1);
      uint32_t p = 0, nx = 0;
      if (ok) {
        const unsigned base = ((unsigned)(i + a) * (unsigned)g.n1 + (unsigned)(j + b)) * (unsigned)g.W + (unsigned)w;
        p = plane[base];
        if (w + 1 < g.W) nx = plane[base + 1];
      }
      pl.P[a * 2 + b] = p;
      pl.S[a * 2 + b] = (p >> 1) | (nx << 31);
    }
}

// crossing words for the 7 edge directions of the owner row (index d-1), masked to edges inside the grid
__device__ __forceinline__ void cross_words(const Planes& pl, uint32_t x[7]) {
  uint32_t A = pl.P[0];
  uint32_t vj = pl.has_j1 ? 0xffffffffu : 0u, vi = pl.has_i1 ? 0xffffffffu : 0u;
  x[0] = (A ^ pl.S[0]) & pl.kp1;                 // d=1 (0,0,1)
  x[1] = (A ^ pl.P[1]) & pl.kpt & vj;            // d=2 (0,1,0)
  x[2] = (A ^ pl.S[1]) & pl.kp1 & vj;            // d=3 (0,1,1)
  x[3] = (A ^ pl.P[2]) & pl.kpt & vi;            // d=4 (1,0,0)
  x[4] = (A ^ pl.S[2]) & pl.kp1 & vi;            // d=5 (1,0,1)
  x[5] = (A ^ pl.P[3]) & pl.kpt & vi & vj;       // d=6 (1,1,0)
  x[6] = (A ^ pl.S[3]) & pl.kp1 & vi & vj;       // d=7 (1,1,1)
}

// the "other endpoint" plane for direction d (1..7) out of a Planes
__device__ __forceinline__ uint32_t dir_plane(const Planes& pl, int d) {
  int ab = d >> 1;
  return (d & 1) ? pl.S[ab] : pl.P[ab];
}

// corner planes of the voxel row: corner c = a*4+b*2+dk
__device__ __forceinline__ uint32_t corner_plane(const Planes& pl, int c) {
  return (c & 1) ? pl.S[c >> 1] : pl.P[c >> 1];
}

// Per-tet words: odd = 1-3 split (1 triangle), two = 2-2 split (2 triangles); cand = crossing & all four near.
__device__ __forceinline__ void tet_words(const Planes& pl, const Planes* npl, uint32_t cellmask, uint32_t odd[6],
                                          uint32_t two[6], uint32_t& cand) {
  const uint32_t A = pl.P[0], H = pl.S[3];
  const int xs[6] = {1, 3, 2, 6, 4, 5};          // tets [A,H,x,y]: (B,D),(D,C),(C,G),(G,E),(E,F),(F,B)
  const int ys[6] = {3, 2, 6, 4, 5, 1};
  cand = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    uint32_t X = corner_plane(pl, xs[t]), Y = corner_plane(pl, ys[t]);
    uint32_t par = A ^ H ^ X ^ Y;
    uint32_t dif = (A ^ H) | (A ^ X) | (A ^ Y);
    odd[t] = par & cellmask;
    two[t] = ~par & dif & cellmask;
    if (npl) {
      uint32_t nn = npl->P[0] & npl->S[3] & corner_plane(*npl, xs[t]) & corner_plane(*npl, ys[t]);
      cand |= nn & dif & cellmask;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// used-edge words of an owner word (index d-1), allclose-ambiguous edges resolved exactly (serial, rare)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void resolve_used_exact(const Grid<T>& g, int i, int j, int w, uint32_t used[7]) {
  Planes npl;
  load_planes(g, g.nbits, i, j, w, npl);
  for (int d = 1; d <= 7; ++d) {
    uint32_t c = used[d - 1] & npl.P[0] & dir_plane(npl, d);
    while (c) {
      int b = __ffs(c) - 1;
      c &= c - 1;
      if (!edge_used_exact(g, i, j, w * 32 + b, d)) used[d - 1] &= ~(1u << b);
    }
  }
}

template <typename T>
__device__ __forceinline__ void owner_used(const Grid<T>& g, bool near, const Planes& pl, int i, int j, int w,
                                           uint32_t used[7]) {
  cross_words(pl, used);
  if (near) resolve_used_exact(g, i, j, w, used);
}

__device__ __forceinline__ unsigned tet_mask_of(unsigned corner8, int t) {
  // tets [A,H,x,y] with (x,y) = (B,D),(D,C),(C,G),(G,E),(E,F),(F,B)  (tetrahedral.py:32-39)
  const int xs[6] = {1, 3, 2, 6, 4, 5};
  const int ys[6] = {3, 2, 6, 4, 5, 1};
  return (corner8 & 1u) | (((corner8 >> 7) & 1u) << 1) | (((corner8 >> xs[t]) & 1u) << 2) | (((corner8 >> ys[t]) & 1u) << 3);
}

// Exact (allclose-aware) corrections of one word, the rare part of stages 2 and 4:
//   *ncross : strict-crossing count (grid_field.py:81): a crossing whose HIGH endpoint equals the isovalue is not strict;
//   odd/two : bits of tets that do not emit under the exact rules (tetrahedral.py:391,576) are cleared.
template <typename T>
__device__ __noinline__ void word_exact_fix(const Grid<T>& g, const Planes& pl, int i, int j, int w, unsigned* ncross,
                                            uint32_t odd[6], uint32_t two[6]) {
  Planes npl;
  load_planes(g, g.nbits, i, j, w, npl);
  if (ncross) {
    uint32_t xs[7];
    cross_words(pl, xs);
    for (int d = 1; d <= 7; ++d) {
      uint32_t A = pl.P[0], O = dir_plane(pl, d);
      uint32_t c = xs[d - 1] & pl.kp1 & ((~A & npl.P[0]) | (~O & dir_plane(npl, d)));
      while (c) {
        int b = __ffs(c) - 1;
        c &= c - 1;
        bool a_high = !((A >> b) & 1u);
        int k = w * 32 + b;
        double fh = a_high ? sample(g, i, j, k) : sample(g, i + ((d >> 2) & 1), j + ((d >> 1) & 1), k + (d & 1));
        if (fh == g.v) --*ncross;
      }
    }
  }
  uint32_t o2[6], t2[6], cand;
  tet_words(pl, &npl, pl.kp1, o2, t2, cand);
  while (cand) {
    int b = __ffs(cand) - 1;
    cand &= cand - 1;
    const unsigned e = cell_emit_exact(g, i, j, w * 32 + b, nullptr);
    const uint32_t keep = ~(1u << b);
    for (int q = 0; q < 6; ++q)
      if (!((e >> q) & 1u)) {
        odd[q] &= keep;
        two[q] &= keep;
      }
  }
}

// carry-save adders over bit-sliced counters (one counter per bit position of the word)
__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy) {
  const uint32_t ab = a ^ b;
  s = ab ^ c;
  cy = (a & b) | (c & ab);
}
// per-voxel number of triangles (0..12) = sum(odd) + 2*sum(two) as four bit planes
__device__ __forceinline__ void slice_tris(const uint32_t odd[6], const uint32_t two[6], uint32_t s[4]) {
  uint32_t a0, a1, a2, b0, b1, b2, p, q, cp, cq, c;
  full_add(odd[0], odd[1], odd[2], p, cp);
  full_add(odd[3], odd[4], odd[5], q, cq);
  a0 = p ^ q;
  c = p & q;
  full_add(cp, cq, c, a1, a2);
  full_add(two[0], two[1], two[2], p, cp);
  full_add(two[3], two[4], two[5], q, cq);
  b0 = p ^ q;
  c = p & q;
  full_add(cp, cq, c, b1, b2);
  // count = A + 2B
  s[0] = a0;
  s[1] = a1 ^ b0;
  const uint32_t k1 = a1 & b0;
  uint32_t k2;
  full_add(a2, b1, k1, s[2], k2);
  s[3] = b2 ^ k2;          // count <= 12: no carry out of bit 3
}

// n-th (0-based) set bit of x; x has more than n bits set.  (__fns costs ~50 instructions; this is ~25.)
__device__ __forceinline__ int nth_set_bit(uint32_t x, unsigned n) {
  int b = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const uint32_t below = (1u << (b + step)) - 1u;      // b + step <= 31
    if ((unsigned)__popc(x & below) <= n) b += step;
  }
  return b;
}

// Vertex order inside a word: direction-major, then bit (k).  dirpack byte q-1 (q = 1..6) = number of used edges of
// the word in the directions before direction index q (index = d-1); id = vbase[word] + dirbase(q) + rank of the bit.
// Kept as two 32-bit halves (bytes 0..3 | bytes 4..5) so that every extraction is a 32-bit shift.
__device__ __forceinline__ uint2 dir_pack(const uint32_t x[7]) {
  unsigned c = 0;
  uint2 dp = make_uint2(0u, 0u);
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    c += __popc(x[q]);
    if (q < 4) dp.x |= c << (8 * q);
    else dp.y |= c << (8 * (q - 4));
  }
  return dp;
}
__device__ __forceinline__ unsigned dir_base(uint2 dp, int q) {   // q = d-1 in 0..6 (compile-time constant when unrolled)
  return q == 0 ? 0u : q <= 4 ? (dp.x >> (8 * (q - 1))) & 255u : (dp.y >> (8 * (q - 5))) & 255u;
}

// ------------------------------------------------------------------------------------------------
// Stage 2: per-word counts from the bitplanes + exclusive scan, in ONE launch of a few hundred CTAs.
// A CTA takes a contiguous chunk of words (ticket order) and
//   1  walks it in sub-tiles of 1024 words: every thread tests 4 words (does anything cross here?), the CTA
//      compacts the interesting ones in shared memory and deals them out evenly: vertex / triangle counts
//      (popcounts of bit-sliced words) -> shared memory, per-direction prefix (dirpack) -> wdir[word];
//   2  publishes the chunk aggregate and sums the aggregates of ALL earlier chunks (one status word each, read
//      in parallel by the CTA's threads: no serial look-back chain; earlier tickets have started, so no deadlock);
//   3  scans its counts warp-contiguously (lane = word: coalesced) -> wpre[word] = (vbase, tbase).
// Nothing is compacted for the later stages: they expand words -> vertices / voxels themselves, warp by warp.
// ------------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 256;
constexpr int CS_ITEMS = 4;
constexpr int CS_SUB = CS_THREADS * CS_ITEMS;
constexpr unsigned CS_MAX_CHUNK = 14 * 1024;          // words per chunk: 4 bytes of shared memory each

#define CTR_ST_VALID (1ull << 63)

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* s_warp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                   // s_warp may still be read from a previous use
  if (lane_id() == 0) s_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long t = 0;
#pragma unroll
  for (int q = 0; q < CS_THREADS / 32; ++q) t += s_warp[q];
  return t;
}

// quick test of one word: does any of its 7 x 32 owned edges cross?  32-bit word indices, no masks beyond the row ends
template <typename T>
__device__ __forceinline__ bool word_interesting(const Grid<T>& g, unsigned gw, unsigned plane_words) {
  const unsigned row = g.divW.div(gw);
  const unsigned w = gw - row * (unsigned)g.W;
  const unsigned i = g.divN1.div(row);
  const unsigned j = row - i * (unsigned)g.n1;
  const bool hi = (int)i + 1 < g.n0, hj = (int)j + 1 < g.n1, hw = (int)w + 1 < g.W;
  const uint32_t* p = g.bits + gw;
  const uint32_t A = p[0];
  const int rem = g.n2 - (int)w * 32;
  const uint32_t kpt = low_mask(rem), kp1 = low_mask(rem - 1);
  uint32_t any_t = 0, any_1 = 0;                     // crossings towards k (kpt mask) and towards k+1 (kp1 mask)
  {
    const uint32_t n0 = hw ? p[1] : 0u;
    any_1 |= A ^ __funnelshift_r(A, n0, 1);
  }
  if (hj) {
    const uint32_t b = p[g.W], nb = hw ? p[g.W + 1] : 0u;
    any_t |= A ^ b;
    any_1 |= A ^ __funnelshift_r(b, nb, 1);
  }
  if (hi) {
    const uint32_t b = p[plane_words], nb = hw ? p[plane_words + 1] : 0u;
    any_t |= A ^ b;
    any_1 |= A ^ __funnelshift_r(b, nb, 1);
    if (hj) {
      const uint32_t c = p[plane_words + g.W], nc = hw ? p[plane_words + g.W + 1] : 0u;
      any_t |= A ^ c;
      any_1 |= A ^ __funnelshift_r(c, nc, 1);
    }
  }
  return ((any_t & kpt) | (any_1 & kp1)) != 0;
}

template <typename T>
__global__ void __launch_bounds__(CS_THREADS, 3) k_count_scan(const __grid_constant__ Grid<T> g, unsigned word0,
                                                              unsigned nwords_scan, unsigned chunk, uint2* __restrict__ wpre,
                                                              uint2* __restrict__ wdir, unsigned long long* status,
                                                              Counters* ctr, int nchunks) {
  extern __shared__ uint32_t s_cnt[];                // per word of the chunk: t << 16 | v
  __shared__ unsigned short s_list[CS_SUB];
  __shared__ unsigned long long s_warp[CS_THREADS / 32];
  __shared__ unsigned s_tile, s_nint;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket, 1u);
  __syncthreads();
  const unsigned c = s_tile;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned plane_words = (unsigned)g.n1 * (unsigned)g.W;
  const unsigned emit_end = (unsigned)g.i_hi * plane_words;      // words below this are emitted
  const unsigned c0 = c * chunk;
  const unsigned cn = min(chunk, nwords_scan - c0);

  // ---- 1: counts
  unsigned ncross = 0, ncells = 0;
  for (unsigned sub = 0; sub < cn; sub += CS_SUB) {
    if (threadIdx.x == 0) s_nint = 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CS_ITEMS; ++q) {
      const unsigned wl = (unsigned)q * CS_THREADS + threadIdx.x;
      bool interesting = false;
      if (sub + wl < cn) {
        interesting = word_interesting(g, word0 + c0 + sub + wl, plane_words);
        s_cnt[sub + wl] = 0;
      }
      const unsigned m = __ballot_sync(0xffffffffu, interesting);
      unsigned base = 0;
      if (lane == 0 && m) base = atomicAdd(&s_nint, (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (interesting) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)wl;
    }
    __syncthreads();
    const unsigned nint = s_nint;
    for (unsigned idx = threadIdx.x; idx < nint; idx += CS_THREADS) {
      const unsigned wl = s_list[idx];
      const unsigned gw = word0 + c0 + sub + wl;
      int i, j, w;
      g.word_coords(gw, i, j, w);
      const bool near = g.rowflag[(unsigned)i * (unsigned)g.n1 + (unsigned)j] != 0;
      Planes pl;
      load_planes(g, g.bits, i, j, w, pl);
      uint32_t x[7];
      owner_used(g, near, pl, i, j, w, x);
      unsigned v = 0;
#pragma unroll
      for (int d = 0; d < 7; ++d) v += __popc(x[d]);
      if (v) wdir[gw] = dir_pack(x);
      unsigned t = 0;
      if (pl.has_i1 && pl.has_j1 && i < g.i_hi) {
        // strict crossings for owners inside the voxel range (grid_field.py:64-84)
        unsigned nc = 0;
        if (!near) {
#pragma unroll
          for (int d = 0; d < 7; ++d) nc += __popc(x[d] & pl.kp1);
        } else {
          uint32_t xs[7];
          cross_words(pl, xs);
#pragma unroll
          for (int d = 0; d < 7; ++d) nc += __popc(xs[d] & pl.kp1);
        }
        uint32_t odd[6], two[6], cand;
        tet_words(pl, nullptr, pl.kp1, odd, two, cand);
        if (near) word_exact_fix(g, pl, i, j, w, &nc, odd, two);
        uint32_t em = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          t += __popc(odd[q]) + 2 * __popc(two[q]);
          em |= odd[q] | two[q];
        }
        ncross += nc;
        ncells += __popc(em);
      }
      s_cnt[sub + wl] = (t << 16) | v;
    }
  }
  {
    unsigned long long cc = ((unsigned long long)ncross << 32) | ncells;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, o);
    if (lane == 0 && cc) {
      if (cc & 0xffffffffull) atomicAdd(&ctr->n_cells, cc & 0xffffffffull);
      if (cc >> 32) atomicAdd(&ctr->n_cross, cc >> 32);
    }
  }
  __syncthreads();

  // ---- 2: chunk aggregate (T << 31 | V), published; exclusive prefix = sum over all earlier chunks
  // warp q scans the contiguous words [q*per, (q+1)*per) of the chunk in step 3; its total is needed first
  const unsigned per = ((cn + CS_THREADS - 1) / CS_THREADS) * 32u;          // multiple of 32, 8*per >= cn
  const unsigned wbeg = min(warp * per, cn), wend = min(wbeg + per, cn);
  unsigned long long wsum = 0;
  for (unsigned q = wbeg + lane; q < wend; q += 32) {
    const uint32_t e = s_cnt[q];
    wsum += ((unsigned long long)(e >> 16) << 31) | (e & 0xffffu);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if (lane == 0) s_warp[warp] = wsum;
  __syncthreads();
  unsigned long long agg = 0, woff = 0;
#pragma unroll
  for (int q = 0; q < CS_THREADS / 32; ++q) {
    if (q < (int)warp) woff += s_warp[q];
    agg += s_warp[q];
  }
  if (threadIdx.x == 0) lb_store(&status[c], CTR_ST_VALID | agg);
  unsigned long long pre = 0;
  for (unsigned p = threadIdx.x; p < c; p += CS_THREADS) {
    unsigned long long s;
    do {
      s = lb_load(&status[p]);
    } while (!(s & CTR_ST_VALID));
    pre += s & ~CTR_ST_VALID;
  }
  const unsigned long long excl = block_sum_u64(pre, s_warp);

  // ---- 3: warp-contiguous scan, coalesced (vbase, tbase) writes
  unsigned long long run = excl + woff;
  uint2* out = wpre + word0 + c0;
  const unsigned emit_rel = emit_end - (word0 + c0);             // wraps to a huge value when emit_end is before this chunk
  for (unsigned q0 = wbeg; q0 < wend; q0 += 32) {
    const unsigned q = q0 + lane;
    unsigned long long val = 0;
    if (q < wend) {
      const uint32_t e = s_cnt[q];
      val = ((unsigned long long)(e >> 16) << 31) | (e & 0xffffu);
    }
    const unsigned long long inc = warp_incl_scan_u64(val);
    if (q < wend) {
      const unsigned long long ex = run + inc - val;
      const uint32_t vb = (uint32_t)ex & 0x7fffffffu;
      out[q] = make_uint2(vb, (uint32_t)(ex >> 31));
      if (q == emit_rel && g.i_hiv > g.i_hi) ctr->v_emit = vb;
    }
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (c == (unsigned)nchunks - 1 && threadIdx.x == 0) {
    const unsigned long long tot = excl + agg;
    ctr->total_vt = tot;
    wpre[word0 + nwords_scan] = make_uint2((uint32_t)(tot & 0x7fffffffull), (uint32_t)(tot >> 31));   // sentinel
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 3: vertices.  Persistent warps; a warp takes 64 consecutive words at a time (2048 owner points, two words per
// lane).  Each lane recomputes the used-edge words of its words into shared memory, then the warp deals the
// segment's vertices out one per lane per round: a lane finds its word by binary search over the warp's prefix
// sums (shuffles), the direction from the word's dirpack and the owner point as the n-th set bit of that direction's
// word.  Consecutive lanes write consecutive vertex ids.  The id of an edge is a perfect hash of its key:
// vbase[word] + dirbase + rank (the reference's dict dedup, tetrahedral.py:184-188, without a table).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
// geometry-mode arithmetic: IEEE in fp64 mode (bit-exact vs the oracle); fast approximations in fp32 mode (1e-4 rel)
__device__ __forceinline__ double quot(double a, double b) { return a / b; }
__device__ __forceinline__ float quot(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double inv_sqrt(double a) { return 1.0 / sqrt(a); }
__device__ __forceinline__ float inv_sqrt(float a) { return rsqrtf(a); }

// central differences (one-sided on the boundary), in grid units
template <typename T, typename G>
__device__ __forceinline__ void grad_at(const Grid<T>& g, const T* __restrict__ c, int i, int j, int k, long long s0,
                                        long long s1, G out[3]) {
  {
    const bool lo = i > 0, hi = i < g.n0 - 1;
    G a = (G)c[hi ? s0 : 0], b = (G)c[lo ? -s0 : 0];
    out[0] = (a - b) * ((lo && hi) ? (G)0.5 : (G)1);
  }
  {
    const bool lo = j > 0, hi = j < g.n1 - 1;
    G a = (G)c[hi ? s1 : 0], b = (G)c[lo ? -s1 : 0];
    out[1] = (a - b) * ((lo && hi) ? (G)0.5 : (G)1);
  }
  {
    const bool lo = k > 0, hi = k < g.n2 - 1;
    G a = (G)c[hi ? 1 : 0], b = (G)c[lo ? -1 : 0];
    out[2] = (a - b) * ((lo && hi) ? (G)0.5 : (G)1);
  }
}

template <typename G>
struct Xform {
  G origin[3], delta[3], inv_delta[3];
  G den_tol;                 // largest G <= 1e-8 (np.allclose(fhigh - flow, 0), tetrahedral.py:483)
};

// lane o of the warp = last lane whose exclusive prefix <= slot (prefixes are non-decreasing; lanes with a
// zero count are skipped because the next lane has the same prefix)
__device__ __forceinline__ int warp_find_owner(unsigned excl, unsigned slot) {
  int o = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int probe = o + step;
    const unsigned e = __shfl_sync(0xffffffffu, excl, probe & 31);
    if (probe < 32 && e <= slot) o = probe;
  }
  return o;
}

constexpr int EV_THREADS = 256;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int EV_BLOCKS_PER_SM = 4;
constexpr int SEG_WORDS = 64;                        // words per warp step in stages 3 and 4 (two per lane)
constexpr unsigned SEG_BATCH = 4;                    // segments claimed per atomic

template <typename T, typename G>
__global__ void __launch_bounds__(EV_THREADS, EV_BLOCKS_PER_SM)
    k_emit_verts(const __grid_constant__ Grid<T> g, unsigned word0, unsigned nwords_emit, const uint2* __restrict__ wpre,
                 const uint2* __restrict__ wdir, const __grid_constant__ Xform<G> xf, G* __restrict__ verts,
                 G* __restrict__ normals, unsigned long long* __restrict__ keys, uint8_t* __restrict__ lowmin, unsigned cap_v,
                 unsigned* claim) {
  __shared__ uint32_t s_x[EV_WARPS][SEG_WORDS][10];  // 7 used-edge words, the word's own low bits, dirpack
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t (*sx)[10] = s_x[warp];
  const G v = (G)g.v;
  const long long s1 = g.n2, s0 = (long long)g.n1 * g.n2;
  const unsigned nseg = (nwords_emit + SEG_WORDS - 1) / SEG_WORDS;
  // segments are claimed in batches of SEG_BATCH by an atomic counter (work per segment varies by orders of magnitude)
  for (unsigned seg = 0, seg_end = 0;; ++seg) {
    if (seg == seg_end) {
      if (lane == 0) seg = atomicAdd(claim, SEG_BATCH);
      seg = __shfl_sync(0xffffffffu, seg, 0);
      seg_end = seg + SEG_BATCH;
    }
    if (seg >= nseg) break;
    const unsigned seg0 = seg * SEG_WORDS;
    unsigned cnt[2] = {0, 0}, vb0 = 0;
    {
      const unsigned rel = seg0 + lane * 2;
      if (rel < nwords_emit) {
        const unsigned gw = word0 + rel;
        vb0 = wpre[gw].x;
        const unsigned a1 = wpre[gw + 1].x;
        cnt[0] = a1 - vb0;
        if (rel + 1 < nwords_emit) cnt[1] = wpre[gw + 2].x - a1;
      }
    }
    const unsigned both = cnt[0] + cnt[1];
    const unsigned incl = warp_incl_scan_u32(both);
    const unsigned excl = incl - both;
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    const unsigned vfirst = __shfl_sync(0xffffffffu, vb0, 0);       // ids of a segment's vertices are consecutive
    __syncwarp();                                                   // previous segment's readers are done with sx
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!cnt[q]) continue;
      const unsigned gw = word0 + seg0 + lane * 2 + q;
      int i, j, w;
      g.word_coords(gw, i, j, w);
      const bool near = g.rowflag[(unsigned)i * (unsigned)g.n1 + (unsigned)j] != 0;
      Planes pl;
      load_planes(g, g.bits, i, j, w, pl);
      uint32_t x[7];
      owner_used(g, near, pl, i, j, w, x);
      const uint2 dp = wdir[gw];
      uint32_t* dst = sx[lane * 2 + q];
#pragma unroll
      for (int d = 0; d < 7; ++d) dst[d] = x[d];
      dst[7] = pl.P[0];
      dst[8] = dp.x;
      dst[9] = dp.y;
    }
    __syncwarp();
    for (unsigned base = 0; base < total; base += 32) {
      const unsigned vtx = base + lane;
      const int o = warp_find_owner(excl, vtx);
      const unsigned o_excl = __shfl_sync(0xffffffffu, excl, o);
      const unsigned o_c0 = __shfl_sync(0xffffffffu, cnt[0], o);
      const unsigned id = vfirst + vtx;
      if (vtx >= total || id >= cap_v) continue;
      unsigned r = vtx - o_excl;
      const unsigned second = r >= o_c0 ? 1u : 0u;
      r -= second ? o_c0 : 0u;
      const unsigned widx = (unsigned)o * 2u + second;
      const unsigned gw = word0 + seg0 + widx;
      const uint32_t* sw = sx[widx];
      const uint2 dp = make_uint2(sw[8], sw[9]);
      // direction index q = number of prefixes <= r; owner point = (r - dirbase)-th set bit of that direction's word
      int q = 0;
      unsigned qb = 0;
#pragma unroll
      for (int t = 1; t <= 6; ++t) {
        const unsigned cb = dir_base(dp, t);
        if (cb <= r) {
          q = t;
          qb = cb;
        }
      }
      const int b = nth_set_bit(sw[q], r - qb);
      const int d = q + 1;
      int oi, oj, ow;
      g.word_coords(gw, oi, oj, ow);
      const int ok = ow * 32 + b;
      const bool p_low = ((sw[7] >> b) & 1u) != 0;
      const int di = (d >> 2) & 1, dj = (d >> 1) & 1, dk = d & 1;
      const T* cp = g.f + ((long long)oi * g.n1 + oj) * g.n2 + ok;
      const T* cq = cp + (di ? s0 : 0) + (dj ? s1 : 0) + dk;
      // all samples this vertex needs are requested before the first one is used (one memory latency per round)
      const bool inner = normals && oi > 0 && oi + di < g.n0 - 1 && oj > 0 && oj + dj < g.n1 - 1 && ok > 0 && ok + dk < g.n2 - 1;
      T sp[6], sq[6];
      if (inner) {
        sp[0] = cp[s0]; sp[1] = cp[-s0]; sp[2] = cp[s1]; sp[3] = cp[-s1]; sp[4] = cp[1]; sp[5] = cp[-1];
        sq[0] = cq[s0]; sq[1] = cq[-s0]; sq[2] = cq[s1]; sq[3] = cq[-s1]; sq[4] = cq[1]; sq[5] = cq[-1];
      }
      const G fp = (G)*cp, fq = (G)*cq;
      // tetrahedral.py:476-487: key oriented (low, high) by value; ratio = (z-flow)/(fhigh-flow), 0.5 if ~0
      const G flow = p_low ? fp : fq, fhigh = p_low ? fq : fp;
      const G den = fhigh - flow;
      const G ratio = (fabs(den) <= xf.den_tol) ? (G)0.5 : quot(v - flow, den);
      // x = low + ratio*(high - low); (high-low) is +-1 or 0 per axis so the product is exact
      const G step = p_low ? ratio : -ratio;
      const int pl[3] = {p_low ? oi : oi + di, p_low ? oj : oj + dj, p_low ? ok : ok + dk};
      const int dd[3] = {di, dj, dk};
      G* vo = verts + (size_t)id * 3;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const G c0 = (ax == 0) ? (G)((long long)pl[0] + g.plane_offset) : (G)pl[ax];
        const G xx = dd[ax] ? add_rn(c0, step) : c0;
        vo[ax] = add_rn(mul_rn(xx, xf.delta[ax]), xf.origin[ax]);   // grid_field.py:93
      }
      if (normals) {
        G gp[3], gq[3];
        if (inner) {
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            gp[ax] = ((G)sp[2 * ax] - (G)sp[2 * ax + 1]) * (G)0.5;
            gq[ax] = ((G)sq[2 * ax] - (G)sq[2 * ax + 1]) * (G)0.5;
          }
        } else {
          grad_at<T, G>(g, cp, oi, oj, ok, s0, s1, gp);
          grad_at<T, G>(g, cq, oi + di, oj + dj, ok + dk, s0, s1, gq);
        }
        G nn[3], len2 = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const G gl = p_low ? gp[ax] : gq[ax], gh = p_low ? gq[ax] : gp[ax];
          nn[ax] = add_rn(gl, mul_rn(ratio, gh - gl)) * xf.inv_delta[ax];
          len2 = add_rn(len2, mul_rn(nn[ax], nn[ax]));
        }
        const G inv = len2 > (G)0 ? inv_sqrt(len2) : (G)0;
        G* no = normals + (size_t)id * 3;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) no[ax] = nn[ax] * inv;
      }
      if (keys) {
        const long long lin = (((long long)oi + g.plane_offset) * g.n1 + oj) * g.n2 + ok;
        keys[id] = ((unsigned long long)lin << 3) | (unsigned)d;
        lowmin[id] = p_low ? 1 : 0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 4: triangles.  Same persistent warp-per-64-words expansion, one lane per emitting voxel.  The id of each of
// the voxel's 19 edges is vbase[owner word] + dirbase + rank of the owner bit in that direction's crossing word, all
// from the 2x2 rows of bit words the voxel touches.  A 256-entry table turns the 8 corner bits into the voxel's
// triangle list (6 Kuhn tets, tetrahedral.py:32-39,554-595; wound towards the high side); the triangles of a
// round are staged in shared memory and written out coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int ET_THREADS = 128;
constexpr int ET_WARPS = ET_THREADS / 32;
constexpr int ET_BLOCKS_PER_SM = 6;
__constant__ uint32_t c_tri_packed[6 * 16];    // per tet: n | slot0<<2 | slot1<<7 | ... (6 slots x 5 bits)

// edge slot e (tables.h CTR_EDGE3_*): owner corner s = a*4 + c*2 + dk and direction d, as compile-time constants
__device__ __forceinline__ constexpr int edge_s(int e) {
  return e < 7 ? 0 : e < 10 ? 1 : e < 13 ? 2 : e < 14 ? 3 : e < 17 ? 4 : e < 18 ? 5 : 6;
}
__device__ __forceinline__ constexpr int edge_d(int e) {
  return e < 7 ? e + 1 : e == 7 ? 2 : e == 8 ? 4 : e == 9 ? 6 : e == 10 ? 1 : e == 11 ? 4 : e == 12 ? 5 : e == 13 ? 4
         : e == 14 ? 1 : e == 15 ? 2 : e == 16 ? 3 : e == 17 ? 2 : 1;
}

// emitting-voxel word of (i, j, w) and the bit-sliced per-voxel triangle counts
template <typename T>
__device__ __forceinline__ uint32_t emit_word(const Grid<T>& g, bool near, int i, int j, int w, uint32_t* s4) {
  Planes pl;
  load_planes(g, g.bits, i, j, w, pl);
  if (!(pl.has_i1 && pl.has_j1 && i < g.i_hi)) {
    if (s4) s4[0] = s4[1] = s4[2] = s4[3] = 0;
    return 0;
  }
  uint32_t odd[6], two[6], cand;
  tet_words(pl, nullptr, pl.kp1, odd, two, cand);
  if (near) word_exact_fix(g, pl, i, j, w, nullptr, odd, two);
  if (s4) slice_tris(odd, two, s4);
  return odd[0] | odd[1] | odd[2] | odd[3] | odd[4] | odd[5] | two[0] | two[1] | two[2] | two[3] | two[4] | two[5];
}

// used-edge words of the voxel's four owner rows when some sample nearby is allclose to the isovalue (rare)
template <typename T>
__device__ __noinline__ void rows_used_exact(const Grid<T>& g, int i, int j, int w, uint32_t X[4][7]) {
  for (int ab = 0; ab < 4; ++ab) {
    Planes pl;
    load_planes(g, g.bits, i + (ab >> 1), j + (ab & 1), w, pl);
    owner_used(g, true, pl, i + (ab >> 1), j + (ab & 1), w, X[ab]);
  }
}

// exact emitting-tet mask of a voxel in a near row (rare)
template <typename T>
__device__ __noinline__ unsigned cell_emit_near(const Grid<T>& g, int i, int j, int w, int b, unsigned c8, bool* exact) {
  unsigned emit = 0;
  for (int t = 0; t < 6; ++t) {
    const unsigned tm = tet_mask_of(c8, t);
    if (tm != 0 && tm != 15) emit |= 1u << t;
  }
  Planes npl;
  load_planes(g, g.nbits, i, j, w, npl);
  unsigned n8 = 0;
  for (int c = 0; c < 8; ++c) n8 |= ((corner_plane(npl, c) >> b) & 1u) << c;
  bool cand = false;
  for (int t = 0; t < 6; ++t) cand = cand || (((emit >> t) & 1u) && tet_mask_of(n8, t) == 15);
  *exact = cand;
  if (cand) emit = cell_emit_exact(g, i, j, w * 32 + b, nullptr);
  return emit;
}

template <typename T>
__global__ void __launch_bounds__(ET_THREADS, ET_BLOCKS_PER_SM)
    k_emit_tris(const __grid_constant__ Grid<T> g, unsigned word0, unsigned nwords_emit, const uint2* __restrict__ wpre,
                const uint2* __restrict__ wdir, const uint32_t* __restrict__ vox_tab, int* __restrict__ tris, unsigned cap_t,
                unsigned* claim) {
  __shared__ unsigned s_ids[ET_WARPS][19][32];
  __shared__ int s_stage[ET_WARPS][32 * 12 * 3];
  __shared__ uint32_t s_ew[ET_WARPS][SEG_WORDS][6];   // emitting-voxel word, 4 bit planes of per-voxel triangle counts, tbase
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t (*ew)[6] = s_ew[warp];
  unsigned (*ids)[32] = s_ids[warp];
  int* stage = s_stage[warp];
  const unsigned plane_words = (unsigned)g.n1 * (unsigned)g.W;
  const unsigned uW = (unsigned)g.W;
  const size_t gcap = (size_t)cap_t * 3;
  const unsigned nseg = (nwords_emit + SEG_WORDS - 1) / SEG_WORDS;
  for (unsigned seg = 0, seg_end = 0;; ++seg) {
    if (seg == seg_end) {
      if (lane == 0) seg = atomicAdd(claim, SEG_BATCH);
      seg = __shfl_sync(0xffffffffu, seg, 0);
      seg_end = seg + SEG_BATCH;
    }
    if (seg >= nseg) break;
    const unsigned seg0 = seg * SEG_WORDS;
    unsigned tcnt[2] = {0, 0}, tb[2] = {0, 0};
    {
      const unsigned rel = seg0 + lane * 2;
      if (rel < nwords_emit) {
        const unsigned gw = word0 + rel;
        tb[0] = wpre[gw].y;
        tb[1] = wpre[gw + 1].y;
        tcnt[0] = tb[1] - tb[0];
        if (rel + 1 < nwords_emit) tcnt[1] = wpre[gw + 2].y - tb[1];
      }
    }
    if (!__any_sync(0xffffffffu, (tcnt[0] | tcnt[1]) != 0)) continue;
    __syncwarp();                                                   // previous segment's readers are done with ew
    unsigned ccnt[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!tcnt[q]) continue;
      const unsigned gw = word0 + seg0 + lane * 2 + q;
      int i, j, w;
      g.word_coords(gw, i, j, w);
      const bool near = g.rowflag[(unsigned)i * (unsigned)g.n1 + (unsigned)j] != 0;
      uint32_t s4[4];
      const uint32_t em = emit_word(g, near, i, j, w, s4);
      uint32_t* dst = ew[lane * 2 + q];
      dst[0] = em;
      dst[1] = s4[0];
      dst[2] = s4[1];
      dst[3] = s4[2];
      dst[4] = s4[3];
      dst[5] = tb[q];
      ccnt[q] = __popc(em);
    }
    __syncwarp();
    const unsigned both = ccnt[0] + ccnt[1];
    const unsigned incl = warp_incl_scan_u32(both);
    const unsigned excl = incl - both;
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    for (unsigned base = 0; base < total; base += 32) {
      const unsigned slot = base + lane;
      const bool act = slot < total;
      const int o = warp_find_owner(excl, slot);
      const unsigned o_excl = __shfl_sync(0xffffffffu, excl, o);
      const unsigned o_c0 = __shfl_sync(0xffffffffu, ccnt[0], o);
      unsigned toff = 0, nt = 0, c8 = 0, emit = 0;
      bool table_path = true;
      if (act) {
        unsigned r = slot - o_excl;
        const unsigned second = r >= o_c0 ? 1u : 0u;
        r -= second ? o_c0 : 0u;
        const unsigned widx = (unsigned)o * 2u + second;
        const unsigned gw = word0 + seg0 + widx;
        const uint32_t* e5 = ew[widx];
        const int b = nth_set_bit(e5[0], r);
        const uint32_t below = (1u << b) - 1u;
        toff = e5[5] + __popc(e5[1] & below) + 2u * __popc(e5[2] & below) + 4u * __popc(e5[3] & below) +
               8u * __popc(e5[4] & below);
        nt = ((e5[1] >> b) & 1u) | (((e5[2] >> b) & 1u) << 1) | (((e5[3] >> b) & 1u) << 2) | (((e5[4] >> b) & 1u) << 3);
        const unsigned row = g.divW.div(gw);
        const unsigned w = gw - row * uW;
        const bool near = g.rowflag[row] != 0;
        // the voxel's 2x2 rows: word w (P) and the word after it; S = the same rows shifted by one sample in k
        const bool next_ok = (w + 1 < uW);
        const unsigned wi[4] = {gw, gw + uW, gw + plane_words, gw + plane_words + uW};
        uint32_t P[4], S[4];
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
          P[ab] = g.bits[wi[ab]];
          const uint32_t nx = next_ok ? g.bits[wi[ab] + 1] : 0u;
          S[ab] = __funnelshift_r(P[ab], nx, 1);
          c8 |= (__funnelshift_r(P[ab], nx, b) & 3u) << (ab * 2);    // bits b, b+1 of the row = corners (ab, dk)
        }
        unsigned vb[4], vb1[4];
        uint2 dp[4], dp1[4];
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
          vb[ab] = wpre[wi[ab]].x;
          dp[ab] = wdir[wi[ab]];
        }
        // used-edge words per owner row (index ab) and direction (index d-1); only the 14 that voxel edges use
        uint32_t X[4][7];
        if (near) {
          const unsigned i = g.divN1.div(row);
          rows_used_exact(g, (int)i, (int)(row - i * (unsigned)g.n1), (int)w, X);
        } else {
          X[0][0] = P[0] ^ S[0]; X[0][1] = P[0] ^ P[1]; X[0][2] = P[0] ^ S[1]; X[0][3] = P[0] ^ P[2];
          X[0][4] = P[0] ^ S[2]; X[0][5] = P[0] ^ P[3]; X[0][6] = P[0] ^ S[3];
          X[1][0] = P[1] ^ S[1]; X[1][3] = P[1] ^ P[3]; X[1][4] = P[1] ^ S[3];
          X[2][0] = P[2] ^ S[2]; X[2][1] = P[2] ^ P[3]; X[2][2] = P[2] ^ S[3];
          X[3][0] = P[3] ^ S[3];
        }
        // owner points at k+1: bit b+1 of the same word, or (b == 31) bit 0 of the next word with rank 0
        uint32_t below1 = below | (1u << b);
        if (b == 31) {
          below1 = 0u;
#pragma unroll
          for (int ab = 0; ab < 3; ++ab) {
            vb1[ab] = wpre[wi[ab] + 1].x;
            dp1[ab] = wdir[wi[ab] + 1];
          }
        } else {
#pragma unroll
          for (int ab = 0; ab < 3; ++ab) {
            vb1[ab] = vb[ab];
            dp1[ab] = dp[ab];
          }
        }
#pragma unroll
        for (int e = 0; e < 19; ++e) {
          const int s = edge_s(e), d = edge_d(e);
          const int ab = s >> 1;
          unsigned id;
          if (s & 1) id = vb1[ab] + dir_base(dp1[ab], d - 1) + __popc(X[ab][d - 1] & below1);
          else id = vb[ab] + dir_base(dp[ab], d - 1) + __popc(X[ab][d - 1] & below);
          ids[e][lane] = id;
        }
        if (near) {
          const unsigned i = g.divN1.div(row);
          bool exact;
          emit = cell_emit_near(g, (int)i, (int)(row - i * (unsigned)g.n1), (int)w, b, c8, &exact);
          table_path = !exact;
        }
      }
      // stage this round's triangles (contiguous in the output: voxels are in word/bit order) and write them coalesced
      const unsigned nact = min(32u, total - base);
      const unsigned first = __shfl_sync(0xffffffffu, toff, 0);
      const unsigned end = __shfl_sync(0xffffffffu, toff + nt, (int)nact - 1);
      if (act) {
        int* dst = stage + (toff - first) * 3u;
        if (table_path) {
          const uint32_t* tab = vox_tab + c8 * 12;
          for (unsigned t = 0; t < nt; ++t) {
            const uint32_t e = __ldg(tab + t);          // three byte offsets into this lane's column of ids
            const char* col = (const char*)&ids[0][lane];
            dst[0] = *(const int*)(col + (e & 0xfffu));
            dst[1] = *(const int*)(col + ((e >> 12) & 0xfffu));
            dst[2] = *(const int*)(col + (e >> 24 << 4));
            dst += 3;
          }
        } else {
          for (int t = 0; t < 6; ++t) {
            if (!((emit >> t) & 1u)) continue;
            const uint32_t e = c_tri_packed[t * 16 + tet_mask_of(c8, t)];
            dst[0] = (int)ids[(e >> 2) & 31u][lane];
            dst[1] = (int)ids[(e >> 7) & 31u][lane];
            dst[2] = (int)ids[(e >> 12) & 31u][lane];
            if ((e & 3u) == 2u) {
              dst[3] = (int)ids[(e >> 17) & 31u][lane];
              dst[4] = (int)ids[(e >> 22) & 31u][lane];
              dst[5] = (int)ids[(e >> 27) & 31u][lane];
            }
            dst += (e & 3u) * 3;
          }
        }
      }
      __syncwarp();
      {
        const size_t gbase = (size_t)first * 3;
        unsigned nint = (end - first) * 3u;
        if (gbase + nint > gcap) nint = gbase < gcap ? (unsigned)(gcap - gbase) : 0u;
        int* out = tris + gbase;
        for (unsigned q = lane; q < nint; q += 32) out[q] = stage[q];
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Parity output: (voxel, case code) of emitting voxels, recomputed from the samples in fp64
// (independent of the bit logic above).  Unordered (slots are handed out per warp by an atomic counter).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_codes(const __grid_constant__ Grid<T> g, unsigned word0, unsigned nwords_emit,
                                               const uint2* __restrict__ wpre, Counters* ctr, unsigned cap,
                                               long long* __restrict__ cells, uint32_t* __restrict__ codes) {
  const unsigned lane = lane_id();
  const unsigned seg0 = (blockIdx.x * 8u + (threadIdx.x >> 5)) * 32u;
  if (seg0 >= nwords_emit) return;
  const bool have = seg0 + lane < nwords_emit;
  const unsigned gw = word0 + seg0 + (have ? lane : 0u);
  uint32_t em = 0;
  if (have && wpre[gw + 1].y != wpre[gw].y) {
    int i, j, w;
    g.word_coords(gw, i, j, w);
    em = emit_word(g, g.rowflag[(size_t)i * g.n1 + j] != 0, i, j, w, nullptr);
  }
  const unsigned ccnt = __popc(em);
  const unsigned incl = warp_incl_scan_u32(ccnt);
  const unsigned excl = incl - ccnt;
  const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  unsigned out0 = 0;
  if (lane == 0) out0 = atomicAdd(&ctr->n_codes, total);
  out0 = __shfl_sync(0xffffffffu, out0, 0);
  for (unsigned base = 0; base < total; base += 32) {
    const unsigned slot = base + lane;
    const int o = warp_find_owner(excl, slot);
    const unsigned o_excl = __shfl_sync(0xffffffffu, excl, o);
    const unsigned o_gw = __shfl_sync(0xffffffffu, gw, o);
    const uint32_t o_em = __shfl_sync(0xffffffffu, em, o);
    if (slot >= total || out0 + slot >= cap) continue;
    const int b = nth_set_bit(o_em, slot - o_excl);
    int i, j, w;
    g.word_coords(o_gw, i, j, w);
    const int k = w * 32 + b;
    unsigned code = 0;
    const unsigned e = cell_emit_exact(g, i, j, k, &code);
    cells[out0 + slot] = e ? (((long long)i + g.plane_offset) * (g.n1 - 1) + j) * (long long)(g.n2 - 1) + k : -1;
    codes[out0 + slot] = code;
  }
}

// one launch instead of four memsets / copies in front of every run
__global__ void k_reset3(Counters* ctr, uint4* rowflag16, size_t n16, unsigned long long* tile_state, size_t ntile) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)gridDim.x * blockDim.x;
  if (t == 0) {
    Counters c;
    memset(&c, 0, sizeof c);
    c.min_key = ~0ull;
    *ctr = c;
  }
  for (size_t q = t; q < n16; q += n) rowflag16[q] = make_uint4(0u, 0u, 0u, 0u);
  for (size_t q = t; q < ntile; q += n) tile_state[q] = 0ull;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool g_tables_loaded[64] = {};

int load_tables(ctr_ctx* ctx) {
  if (ctx->device < 64 && g_tables_loaded[ctx->device]) return 0;
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_tri_n, CTR_TRI3_N_H, sizeof(CTR_TRI3_N_H)));
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_tri_e, CTR_TRI3_E_H, sizeof(CTR_TRI3_E_H)));
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_edge_s, CTR_EDGE3_S_H, sizeof(CTR_EDGE3_S_H)));
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_edge_d, CTR_EDGE3_D_H, sizeof(CTR_EDGE3_D_H)));
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_tetmask, CTR_TETMASK3_H, sizeof(CTR_TETMASK3_H)));
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_tet, CTR_TET3_H, sizeof(CTR_TET3_H)));
  uint32_t packed[96];
  for (int t = 0; t < 6; ++t)
    for (int m = 0; m < 16; ++m) {
      uint32_t e = CTR_TRI3_N_H[t][m];
      for (int q = 0; q < 6; ++q) e |= (uint32_t)(CTR_TRI3_E_H[t][m][q] & 31u) << (2 + 5 * q);
      packed[t * 16 + m] = e;
    }
  CTR_CUDA(ctx, cudaMemcpyToSymbol(c_tri_packed, packed, sizeof(packed)));
  if (ctx->device < 64) g_tables_loaded[ctx->device] = true;
  return 0;
}

// CTR_DEBUG_SYNC=1: synchronise after every launch so that a device fault is reported with the kernel's name
#define CTR_DBG(ctx, name)                                                                       \
  do {                                                                                           \
    static const bool dbg__ = getenv("CTR_DEBUG_SYNC") != nullptr;                               \
    if (dbg__) {                                                                                 \
      cudaError_t e__ = cudaStreamSynchronize((ctx)->stream);                                    \
      if (e__ != cudaSuccess) return ctr_fail(ctx, CTR_ERR_CUDA, name, cudaGetErrorString(e__)); \
    }                                                                                            \
  } while (0)

template <typename T>
int run_typed(ctr_ctx* ctx, const ctr_mt3d_params* p, ctr_mt3d_counts* out) {
  const int n0 = (int)p->n0, n1 = (int)p->n1, n2 = (int)p->n2;
  const int W = (n2 + 31) / 32;
  const long long nrows = (long long)n0 * n1;
  const long long nwords = nrows * W;
  const size_t nsamp = (size_t)nrows * n2;
  cudaStream_t st = ctx->stream;
  if (nwords >= (1ll << 31) - 64) return ctr_fail(ctx, CTR_ERR_UNSUPPORTED, "volume too large for 32-bit word indices");
  int rc;
  if ((rc = load_tables(ctx))) return rc;
  ctr_stage_mark(ctx, 0);
  const T* dfield;
  if (p->flags & CTR_FIELD_ON_DEVICE) {
    dfield = (const T*)p->field;
  } else {
    if ((rc = ctr_ensure(ctx, ctx->field, nsamp * sizeof(T)))) return rc;
    CTR_CUDA(ctx, cudaMemcpyAsync(ctx->field.p, p->field, nsamp * sizeof(T), cudaMemcpyHostToDevice, st));
    dfield = (const T*)ctx->field.p;
  }
  ctr_stage_mark(ctx, 1);

  Grid<T> g;
  g.f = dfield;
  g.n0 = n0; g.n1 = n1; g.n2 = n2; g.W = W;
  g.i_lo = (int)p->i_lo;
  g.i_hi = (int)p->i_hi;
  g.i_hiv = std::min(g.i_hi + 1, n0);
  g.plane_offset = p->plane_offset;
  g.v = p->isovalue;
  g.tolv = 1e-8 + 1e-5 * fabs(p->isovalue);
  g.divW.init((unsigned)W);
  g.divN1.init((unsigned)n1);
  const long long plane_words = (long long)n1 * W;
  const unsigned word0 = (unsigned)((long long)g.i_lo * plane_words);
  const unsigned nscan = (unsigned)((long long)(g.i_hiv - g.i_lo) * plane_words);
  const unsigned nemit = (unsigned)((long long)(g.i_hi - g.i_lo) * plane_words);
  unsigned chunk = (nscan + 3u * (unsigned)ctx->sm_count - 1u) / (3u * (unsigned)ctx->sm_count);
  chunk = std::min<unsigned>(std::max<unsigned>((chunk + CS_SUB - 1) / CS_SUB * CS_SUB, CS_SUB), CS_MAX_CHUNK);
  const int ntiles = (int)((nscan + chunk - 1) / chunk);

  if ((rc = ctr_ensure(ctx, ctx->bits, (size_t)(nwords + 4) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->nbits, (size_t)(nwords + 4) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->vbase, (size_t)(nwords + 4) * 8))) return rc;     // wpre: (vbase, tbase) per word
  if ((rc = ctr_ensure(ctx, ctx->tbase, (size_t)(nwords + 4) * 8))) return rc;     // wdir: dirpack per word
  if (!ctx->vox_tab.p) {
    // corner bits -> triangle list of the whole voxel (tet order, then triangle order; one triangle per word)
    static uint32_t tab[256 * 12];
    const int xs[6] = {1, 3, 2, 6, 4, 5}, ys[6] = {3, 2, 6, 4, 5, 1};
    for (unsigned c8 = 0; c8 < 256; ++c8) {
      uint32_t* e = tab + c8 * 12;
      for (int q = 0; q < 12; ++q) e[q] = 0;
      int k = 0;
      for (int t = 0; t < 6; ++t) {
        const unsigned m = (c8 & 1u) | (((c8 >> 7) & 1u) << 1) | (((c8 >> xs[t]) & 1u) << 2) | (((c8 >> ys[t]) & 1u) << 3);
        for (int tri = 0; tri < CTR_TRI3_N_H[t][m]; ++tri, ++k) {
          // byte offsets of the three edge slots in a lane's column of s_ids[19][32]: slot * 128
          const uint32_t a = CTR_TRI3_E_H[t][m][tri * 3 + 0], b = CTR_TRI3_E_H[t][m][tri * 3 + 1], c = CTR_TRI3_E_H[t][m][tri * 3 + 2];
          e[k] = (a * 128u) | ((b * 128u) << 12) | ((c * 8u) << 24);
        }
      }
    }
    if ((rc = ctr_ensure(ctx, ctx->vox_tab, sizeof tab, true))) return rc;
    CTR_CUDA(ctx, cudaMemcpyAsync(ctx->vox_tab.p, tab, sizeof tab, cudaMemcpyHostToDevice, st));
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
  }
  if ((rc = ctr_ensure(ctx, ctx->counters, sizeof(Counters)))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->aux[4], (size_t)nrows + 32))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->tile_state, (size_t)ntiles * 8 + 16))) return rc;
  if (!ctx->counters_host) CTR_CUDA(ctx, cudaMallocHost(&ctx->counters_host, 256));
  g.bits = (const uint32_t*)ctx->bits.p;
  g.nbits = (const uint32_t*)ctx->nbits.p;
  g.rowflag = (const uint8_t*)ctx->aux[4].p;
  Counters* dctr = (Counters*)ctx->counters.p;

  k_reset3<<<64, 256, 0, st>>>(dctr, (uint4*)ctx->aux[4].p, (size_t)(nrows + 15) / 16, (unsigned long long*)ctx->tile_state.p,
                               (size_t)ntiles);
  ctx->launches++;
  CTR_DBG(ctx, "k_reset3");
  if (p->flags & CTR_WANT_MINMAX)
    rc = launch_bitplane<T, true>(ctx, dfield, (unsigned)nrows, n2, W, n0, n1, p->isovalue, (uint32_t*)ctx->bits.p,
                                  (uint32_t*)ctx->nbits.p, (uint8_t*)ctx->aux[4].p, (MinMaxKeys*)dctr, false);
  else
    rc = launch_bitplane<T, false>(ctx, dfield, (unsigned)nrows, n2, W, n0, n1, p->isovalue, (uint32_t*)ctx->bits.p,
                                   (uint32_t*)ctx->nbits.p, (uint8_t*)ctx->aux[4].p, (MinMaxKeys*)dctr, false);
  if (rc) return rc;
  CTR_DBG(ctx, "k_bitplane");
  ctr_stage_mark(ctx, 2);

  if (ntiles > 0) {
    static bool attr_set[2] = {false, false};
    if (!attr_set[sizeof(T) == 4 ? 0 : 1]) {
      CTR_CUDA(ctx, cudaFuncSetAttribute(k_count_scan<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_MAX_CHUNK * 4));
      attr_set[sizeof(T) == 4 ? 0 : 1] = true;
    }
    k_count_scan<T><<<ntiles, CS_THREADS, (size_t)chunk * 4, st>>>(g, word0, nscan, chunk, (uint2*)ctx->vbase.p,
                                                                   (uint2*)ctx->tbase.p,
                                                                   (unsigned long long*)ctx->tile_state.p, dctr, ntiles);
    ctx->launches++;
    CTR_DBG(ctx, "k_count_scan");
  }
  CTR_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  ctr_stage_mark(ctx, 3);

  const bool geom = !(p->flags & CTR_NO_GEOMETRY);
  const bool f64 = (p->flags & CTR_GEOM_F64) != 0;
  const size_t gsz = f64 ? 8 : 4;
  const bool want_n = (p->flags & CTR_WANT_NORMALS) != 0, want_k = (p->flags & CTR_WANT_KEYS) != 0;
  Xform<double> xd;
  Xform<float> xs;
  for (int a = 0; a < 3; ++a) {
    xd.origin[a] = p->origin[a];
    xd.delta[a] = p->delta[a];
    xd.inv_delta[a] = 1.0 / p->delta[a];
    xs.origin[a] = (float)p->origin[a];
    xs.delta[a] = (float)p->delta[a];
    xs.inv_delta[a] = (float)(1.0 / p->delta[a]);
  }
  xd.den_tol = 1e-8;
  xs.den_tol = 1e-8f;
  if ((double)xs.den_tol > 1e-8) xs.den_tol = nextafterf(xs.den_tol, 0.f);
  // Output pools are grow-only.  When a previous run left capacities behind, stages 3-4 are enqueued right away
  // against those capacities (kernels never write past them) and the counts are checked after the single
  // synchronisation at the end; otherwise (first run, or the guess was too small) the counts are read first.
  auto ensure_outputs = [&](size_t cv, size_t ct) -> int {
    int r;
    if ((r = ctr_ensure(ctx, ctx->verts, cv * 3 * gsz, true))) return r;
    if (want_n && (r = ctr_ensure(ctx, ctx->normals, cv * 3 * gsz, true))) return r;
    if (want_k && (r = ctr_ensure(ctx, ctx->keys, cv * 8, true))) return r;
    if (want_k && (r = ctr_ensure(ctx, ctx->lowmin, cv, true))) return r;
    if ((r = ctr_ensure(ctx, ctx->tris, ct * 12, true))) return r;
    return 0;
  };
  auto launch_emit = [&](size_t cv, size_t ct, bool again) -> int {
    if (again) CTR_CUDA(ctx, cudaMemsetAsync(&dctr->claim_v, 0, 2 * sizeof(unsigned), st));
    const unsigned capv = (unsigned)std::min<size_t>(cv, 0x7fffffffu), capt = (unsigned)std::min<size_t>(ct, 0x7fffffffu);
    unsigned long long* dkeys = want_k ? (unsigned long long*)ctx->keys.p : nullptr;
    uint8_t* dlow = want_k ? (uint8_t*)ctx->lowmin.p : nullptr;
    if (nemit) {
      const unsigned segs = (nemit + SEG_WORDS - 1) / SEG_WORDS;
      const unsigned vb = std::min<unsigned>((segs + EV_WARPS - 1) / EV_WARPS, (unsigned)ctx->sm_count * EV_BLOCKS_PER_SM);
      const uint2* wpre = (const uint2*)ctx->vbase.p;
      const uint2* wdir = (const uint2*)ctx->tbase.p;
      if (f64)
        k_emit_verts<T, double><<<vb, EV_THREADS, 0, st>>>(g, word0, nemit, wpre, wdir, xd, (double*)ctx->verts.p,
                                                            want_n ? (double*)ctx->normals.p : nullptr, dkeys, dlow, capv, &dctr->claim_v);
      else
        k_emit_verts<T, float><<<vb, EV_THREADS, 0, st>>>(g, word0, nemit, wpre, wdir, xs, (float*)ctx->verts.p,
                                                           want_n ? (float*)ctx->normals.p : nullptr, dkeys, dlow, capv, &dctr->claim_v);
      ctx->launches++;
      CTR_DBG(ctx, "k_emit_verts");
      ctr_stage_mark(ctx, 4);
      const unsigned tb = std::min<unsigned>((segs + ET_WARPS - 1) / ET_WARPS, (unsigned)ctx->sm_count * ET_BLOCKS_PER_SM);
      k_emit_tris<T><<<tb, ET_THREADS, 0, st>>>(g, word0, nemit, wpre, wdir, (const uint32_t*)ctx->vox_tab.p,
                                                (int*)ctx->tris.p, capt, &dctr->claim_t);
      ctx->launches++;
      CTR_DBG(ctx, "k_emit_tris");
    } else {
      ctr_stage_mark(ctx, 4);
    }
    ctr_stage_mark(ctx, 5);
    CTR_CUDA(ctx, cudaGetLastError());
    return 0;
  };
  Counters h;
  unsigned long long totV = 0, totT = 0, nV = 0;
  auto read_counts = [&]() -> int {
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(&h, ctx->counters_host, sizeof h);
    totV = h.total_vt & 0x7fffffffull;
    totT = h.total_vt >> 31;
    nV = (g.i_hiv > g.i_hi) ? h.v_emit : totV;
    return 0;
  };
  bool emitted = false;
  const bool speculate = geom && ctx->spec_v > 0 && ctx->spec_t > 0 && !(p->flags & CTR_WANT_CODES);
  if (speculate) {
    if ((rc = ensure_outputs(ctx->spec_v, ctx->spec_t))) return rc;
    if ((rc = launch_emit(ctx->spec_v, ctx->spec_t, false))) return rc;
    if ((rc = read_counts())) return rc;
    emitted = nV <= ctx->spec_v && totT <= ctx->spec_t;
  } else {
    if ((rc = read_counts())) return rc;
  }
  if (totV >= 0x7ffffff0ull || totT >= 0x7ffffff0ull)
    return ctr_fail(ctx, CTR_ERR_OVERFLOW, "more than 2^31 vertices or triangles in one call; shard the volume");
  if (geom && !emitted) {
    ctx->spec_v = std::max<size_t>(ctx->spec_v, (size_t)nV + (size_t)nV / 4 + 1024);
    ctx->spec_t = std::max<size_t>(ctx->spec_t, (size_t)totT + (size_t)totT / 4 + 1024);
    if ((rc = ensure_outputs(ctx->spec_v, ctx->spec_t))) return rc;
    if ((rc = launch_emit(ctx->spec_v, ctx->spec_t, speculate))) return rc;
  } else if (!geom) {
    ctr_stage_mark(ctx, 4);
    ctr_stage_mark(ctx, 5);
  }

  out->n_verts = (int64_t)nV;
  out->n_tris = (int64_t)totT;
  out->n_active_cells = (int64_t)h.n_cells;
  out->n_crossings = (int64_t)h.n_cross;
  out->n_codes = 0;
  const bool mm = (p->flags & CTR_WANT_MINMAX) && h.min_key != ~0ull;
  out->fmin = mm ? key_to_double(h.min_key) : NAN;
  out->fmax = mm ? key_to_double(h.max_key) : NAN;

  if (p->flags & CTR_WANT_CODES) {
    const size_t nCell = (size_t)h.n_cells;
    if ((rc = ctr_ensure(ctx, ctx->cells, nCell * 8 + 8))) return rc;
    if ((rc = ctr_ensure(ctx, ctx->codes, nCell * 4 + 4))) return rc;
    if (nCell && nemit) {
      const unsigned segs = (nemit + 31) / 32;
      k_codes<T><<<(segs + 7) / 8, 256, 0, st>>>(g, word0, nemit, (const uint2*)ctx->vbase.p, dctr, (unsigned)nCell,
                                                 (long long*)ctx->cells.p, (uint32_t*)ctx->codes.p);
      ctx->launches++;
    }
    out->n_codes = (int64_t)nCell;
  }
  ctr_stage_mark(ctx, 6);
  CTR_CUDA(ctx, cudaGetLastError());
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  if (ctx->timing) {
    for (int s = 0; s < 6; ++s) {
      ctx->stage_ms[s] = 0.f;
      if (ctx->ev_set[s] && ctx->ev_set[s + 1]) cudaEventElapsedTime(&ctx->stage_ms[s], ctx->ev[s], ctx->ev[s + 1]);
    }
  }
  ctx->last_kind = 3;
  ctx->last_flags = p->flags;
  ctx->last_counts[0] = (int64_t)nV;
  ctx->last_counts[1] = (int64_t)totT;
  ctx->last_counts[2] = out->n_codes;
  return 0;
}

}  // namespace

extern "C" int ctr_mt3d_run(ctr_ctx* ctx, const ctr_mt3d_params* p, ctr_mt3d_counts* out) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (!p || !out || !p->field) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "null argument");
  if (p->n0 < 2 || p->n1 < 2 || p->n2 < 2) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "grid must have at least 2 samples per axis");
  if (p->n0 > 0x7ffffff0ll || p->n1 > 0x7ffffff0ll || p->n2 > 0x7ffffff0ll) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "axis too long");
  if (p->i_lo < 0 || p->i_hi > p->n0 || p->i_lo >= p->i_hi) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "bad slab range [i_lo, i_hi)");
  if (!(p->isovalue == p->isovalue)) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "isovalue is NaN");
  for (int a = 0; a < 3; ++a)
    if (p->delta[a] == 0.0) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "delta must be non-zero");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  memset(out, 0, sizeof *out);
  ctx->last_kind = 0;
  if (p->dtype == CTR_F32) return run_typed<float>(ctx, p, out);
  if (p->dtype == CTR_F64) return run_typed<double>(ctx, p, out);
  return ctr_fail(ctx, CTR_ERR_BAD_ARG, "dtype must be CTR_F32 or CTR_F64");
}

extern "C" int ctr_mt3d_fetch(ctr_ctx* ctx, void* verts, void* normals, int32_t* tris, uint64_t* keys,
                              uint8_t* lowmin, int64_t* cells, uint32_t* codes) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (ctx->last_kind != 3) return ctr_fail(ctx, CTR_ERR_STATE, "no completed ctr_mt3d_run to fetch from");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t fl = ctx->last_flags;
  const size_t gsz = (fl & CTR_GEOM_F64) ? 8 : 4;
  const size_t nV = (size_t)ctx->last_counts[0], nT = (size_t)ctx->last_counts[1], nC = (size_t)ctx->last_counts[2];
  cudaStream_t st = ctx->stream;
  const bool geom = !(fl & CTR_NO_GEOMETRY);
  if (verts) {
    if (!geom) return ctr_fail(ctx, CTR_ERR_STATE, "run had CTR_NO_GEOMETRY");
    if (nV) CTR_CUDA(ctx, cudaMemcpyAsync(verts, ctx->verts.p, nV * 3 * gsz, cudaMemcpyDeviceToHost, st));
  }
  if (normals) {
    if (!geom || !(fl & CTR_WANT_NORMALS)) return ctr_fail(ctx, CTR_ERR_STATE, "run did not compute normals");
    if (nV) CTR_CUDA(ctx, cudaMemcpyAsync(normals, ctx->normals.p, nV * 3 * gsz, cudaMemcpyDeviceToHost, st));
  }
  if (tris) {
    if (!geom) return ctr_fail(ctx, CTR_ERR_STATE, "run had CTR_NO_GEOMETRY");
    if (nT) CTR_CUDA(ctx, cudaMemcpyAsync(tris, ctx->tris.p, nT * 12, cudaMemcpyDeviceToHost, st));
  }
  if (keys || lowmin) {
    if (!geom || !(fl & CTR_WANT_KEYS)) return ctr_fail(ctx, CTR_ERR_STATE, "run did not compute keys");
    if (keys && nV) CTR_CUDA(ctx, cudaMemcpyAsync(keys, ctx->keys.p, nV * 8, cudaMemcpyDeviceToHost, st));
    if (lowmin && nV) CTR_CUDA(ctx, cudaMemcpyAsync(lowmin, ctx->lowmin.p, nV, cudaMemcpyDeviceToHost, st));
  }
  if (cells || codes) {
    if (!(fl & CTR_WANT_CODES)) return ctr_fail(ctx, CTR_ERR_STATE, "run did not compute case codes");
    if (cells && nC) CTR_CUDA(ctx, cudaMemcpyAsync(cells, ctx->cells.p, nC * 8, cudaMemcpyDeviceToHost, st));
    if (codes && nC) CTR_CUDA(ctx, cudaMemcpyAsync(codes, ctx->codes.p, nC * 4, cudaMemcpyDeviceToHost, st));
  }
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}

extern "C" int ctr_mt3d_device_ptrs(ctr_ctx* ctx, void** verts, void** normals, int32_t** tris) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (ctx->last_kind != 3) return ctr_fail(ctx, CTR_ERR_STATE, "no completed ctr_mt3d_run");
  if (verts) *verts = ctx->verts.p;
  if (normals) *normals = (ctx->last_flags & CTR_WANT_NORMALS) ? ctx->normals.p : nullptr;
  if (tris) *tris = (int32_t*)ctx->tris.p;
  return 0;
}
