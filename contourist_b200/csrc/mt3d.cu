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
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tables.h"

namespace {

__constant__ uint8_t c_tri_n[6][16];
__constant__ uint8_t c_tri_e[6][16][6];
__constant__ uint8_t c_edge_s[19];
__constant__ uint8_t c_edge_d[19];
__constant__ uint8_t c_tetmask[8][8];
__constant__ uint8_t c_tet[6][4];

struct Counters {                    // device counter block (mirrored to pinned host memory)
  unsigned long long min_key;        // order-preserving encoding of fmin
  unsigned long long max_key;        // order-preserving encoding of fmax
  unsigned long long n_cells;        // emitting voxels
  unsigned long long n_cross;        // strict crossings
  unsigned long long total_vt;       // packed totals over the scanned range: T << 31 | V
  unsigned long long total_act;      // packed active-word counts: actT << 31 | actV
  unsigned long long v_emit;         // vertex count at the start of plane i_hi (vertices this call emits)
  unsigned int any_near;
  unsigned int ticket;
  unsigned int n_codes;
  unsigned int pad;
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
  int any_near;                      // filled in-kernel from *near_flag
  const unsigned* near_flag;
};

__device__ __forceinline__ unsigned long long order_key(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// ------------------------------------------------------------------------------------------------
// Stage 1: field -> bitplanes.  One warp converts 32 consecutive samples of a row per step.
// ------------------------------------------------------------------------------------------------
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) k_bitplane(const T* __restrict__ f, long long nrows, int n2, int W,
                                                  T thr, T near_lo, T near_hi, uint32_t* __restrict__ bits,
                                                  uint32_t* __restrict__ nbits, Counters* ctr) {
  const unsigned lane = lane_id();
  const long long nwords = nrows * W;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  T mn = INFINITY, mx = -INFINITY;
  unsigned anynear = 0;
  for (; g < nwords; g += warps * UNROLL) {
    T val[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long gu = g + (long long)u * warps;
      ok[u] = false;
      val[u] = (T)0;
      if (gu < nwords) {
        long long row = gu / W;
        int w = (int)(gu - row * W);
        int k = w * 32 + (int)lane;
        if (k < n2) {
          ok[u] = true;
          val[u] = __ldg(f + row * (long long)n2 + k);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long gu = g + (long long)u * warps;
      if (gu < nwords) {                       // warp-uniform
        bool lo = ok[u] && (val[u] < thr);
        bool nr = ok[u] && (val[u] >= near_lo) && (val[u] <= near_hi);
        unsigned wl = __ballot_sync(0xffffffffu, lo);
        unsigned wn = __ballot_sync(0xffffffffu, nr);
        if (lane == 0) {
          bits[gu] = wl;
          nbits[gu] = wn;
        }
        anynear |= wn;
        if (ok[u]) {
          mn = fmin(mn, val[u]);               // fmin/fmax ignore NaN
          mx = fmax(mx, val[u]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    atomicMin(&ctr->min_key, order_key((double)mn));
    atomicMax(&ctr->max_key, order_key((double)mx));
    if (anynear) atomicOr(&ctr->any_near, 1u);
  }
}

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
__device__ unsigned cell_emit_exact(const Grid<T>& g, int i, int j, int k, unsigned* code_out) {
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
__device__ bool edge_used_exact(const Grid<T>& g, int i, int j, int k, int d) {
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
        long long base = ((long long)(i + a) * g.n1 + (j + b)) * g.W + w;
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
// Stage 2: counts per word + fused decoupled-lookback scan.
// ------------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 256;
constexpr int CS_ITEMS = 4;
constexpr int CS_TILE = CS_THREADS * CS_ITEMS;

template <typename T>
__global__ void __launch_bounds__(CS_THREADS) k_count_scan(Grid<T> gin, long long word0, long long nwords_scan,
                                                           uint32_t* __restrict__ vbase, uint32_t* __restrict__ tbase,
                                                           uint32_t* __restrict__ list_v, uint32_t* __restrict__ list_t,
                                                           unsigned long long* status_vt, unsigned long long* status_act,
                                                           Counters* ctr, int ntiles) {
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_warp_vt[CS_THREADS / 32], s_warp_act[CS_THREADS / 32];
  __shared__ unsigned long long s_excl_vt, s_excl_act;
  Grid<T> g = gin;
  g.any_near = (int)*gin.near_flag;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket, 1u);
  __syncthreads();
  const int tile = (int)s_tile;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

  unsigned cv[CS_ITEMS], ct[CS_ITEMS];
  bool emitv[CS_ITEMS];
  unsigned ncells = 0, ncross = 0;
  const long long plane_words = (long long)g.n1 * g.W;
  const long long first = (long long)tile * CS_TILE + (long long)threadIdx.x * CS_ITEMS;
#pragma unroll
  for (int it = 0; it < CS_ITEMS; ++it) {
    cv[it] = 0;
    ct[it] = 0;
    emitv[it] = false;
    long long rel = first + it;
    if (rel >= nwords_scan) continue;
    long long gw = word0 + rel;
    long long row = gw / g.W;
    int w = (int)(gw - row * g.W);
    int i = (int)(row / g.n1), j = (int)(row - (long long)i * g.n1);
    Planes pl, npl;
    load_planes(g, g.bits, i, j, w, pl);
    uint32_t x[7];
    cross_words(pl, x);
    if (g.any_near) load_planes(g, g.nbits, i, j, w, npl);
    // ---- vertices owned by this word (edge p -> p+d is used iff it crosses, modulo allclose skips)
    unsigned v = 0;
#pragma unroll
    for (int d = 1; d <= 7; ++d) {
      uint32_t xd = x[d - 1];
      if (g.any_near) {
        uint32_t c = xd & npl.P[0] & dir_plane(npl, d);
        xd &= ~c;
        while (c) {
          int b = __ffs(c) - 1;
          c &= c - 1;
          if (edge_used_exact(g, i, j, w * 32 + b, d)) ++v;
        }
      }
      v += __popc(xd);
    }
    cv[it] = v;
    emitv[it] = v && gw < (long long)g.i_hi * plane_words;
    // ---- strict crossings for owners inside the voxel range (grid_field.py:64-84)
    if (pl.has_i1 && pl.has_j1) {
#pragma unroll
      for (int d = 1; d <= 7; ++d) {
        uint32_t xd = x[d - 1] & pl.kp1;
        ncross += __popc(xd);
        if (g.any_near) {
          // not strict when the HIGH endpoint equals the isovalue exactly (then (f0-v)*(f1-v) == 0)
          uint32_t A = pl.P[0], O = dir_plane(pl, d);
          uint32_t c = xd & ((~A & npl.P[0]) | (~O & dir_plane(npl, d)));
          while (c) {
            int b = __ffs(c) - 1;
            c &= c - 1;
            bool a_high = !((A >> b) & 1u);
            int k = w * 32 + b;
            double fh = a_high ? sample(g, i, j, k) : sample(g, i + ((d >> 2) & 1), j + ((d >> 1) & 1), k + (d & 1));
            if (fh == g.v) --ncross;
          }
        }
      }
    }
    // ---- triangles of the voxels in this word
    if (pl.has_i1 && pl.has_j1 && i < g.i_hi) {
      uint32_t odd[6], two[6], cand;
      tet_words(pl, g.any_near ? &npl : nullptr, pl.kp1, odd, two, cand);
      uint32_t emitting = 0;
      unsigned t = 0;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        t += __popc(odd[q] & ~cand) + 2 * __popc(two[q] & ~cand);
        emitting |= (odd[q] | two[q]) & ~cand;
      }
      while (cand) {
        int b = __ffs(cand) - 1;
        cand &= cand - 1;
        unsigned e = cell_emit_exact(g, i, j, w * 32 + b, nullptr);
        if (e) emitting |= 1u << b;
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if ((e >> q) & 1u) t += ((odd[q] >> b) & 1u) ? 1u : 2u;
      }
      ct[it] = t;
      ncells += __popc(emitting);
    }
  }
  // ---- block scan of (V, T) and (activeV, activeT), packed 31 bits each
  unsigned long long loc_vt = 0, loc_act = 0;
#pragma unroll
  for (int it = 0; it < CS_ITEMS; ++it) {
    loc_vt += ((unsigned long long)ct[it] << 31) | cv[it];
    loc_act += ((unsigned long long)(ct[it] ? 1u : 0u) << 31) | (emitv[it] ? 1u : 0u);
  }
  unsigned long long inc_vt = warp_incl_scan_u64(loc_vt), inc_act = warp_incl_scan_u64(loc_act);
  if (lane == 31) {
    s_warp_vt[warp] = inc_vt;
    s_warp_act[warp] = inc_act;
  }
  // block-level counters
  unsigned long long cc = ((unsigned long long)ncross << 32) | ncells;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, o);
  if (lane == 0 && cc) {
    if (cc & 0xffffffffull) atomicAdd(&ctr->n_cells, cc & 0xffffffffull);
    if (cc >> 32) atomicAdd(&ctr->n_cross, cc >> 32);
  }
  __syncthreads();
  unsigned long long woff_vt = 0, woff_act = 0, blk_vt = 0, blk_act = 0;
#pragma unroll
  for (int q = 0; q < CS_THREADS / 32; ++q) {
    if (q < (int)warp) {
      woff_vt += s_warp_vt[q];
      woff_act += s_warp_act[q];
    }
    blk_vt += s_warp_vt[q];
    blk_act += s_warp_act[q];
  }
  if (warp == 0) {
    unsigned long long e = lb_lookback(status_vt, tile, blk_vt);
    if (lane == 0) s_excl_vt = e;
  } else if (warp == 1) {
    unsigned long long e = lb_lookback(status_act, tile, blk_act);
    if (lane == 0) s_excl_act = e;
  }
  __syncthreads();
  unsigned long long run_vt = s_excl_vt + woff_vt + inc_vt - loc_vt;
  unsigned long long run_act = s_excl_act + woff_act + inc_act - loc_act;
#pragma unroll
  for (int it = 0; it < CS_ITEMS; ++it) {
    long long rel = first + it;
    if (rel >= nwords_scan) break;
    long long gw = word0 + rel;
    unsigned vb = (unsigned)(run_vt & 0x7fffffffull), tb = (unsigned)(run_vt >> 31);
    vbase[gw] = vb;
    tbase[gw] = tb;
    if (g.i_hiv > g.i_hi && gw == (long long)g.i_hi * plane_words) ctr->v_emit = vb;
    if (emitv[it]) list_v[(unsigned)(run_act & 0x7fffffffull)] = (uint32_t)gw;
    if (ct[it]) list_t[(unsigned)(run_act >> 31)] = (uint32_t)gw;
    run_vt += ((unsigned long long)ct[it] << 31) | cv[it];
    run_act += ((unsigned long long)(ct[it] ? 1u : 0u) << 31) | (emitv[it] ? 1u : 0u);
  }
  if (tile == ntiles - 1 && threadIdx.x == 0) {
    ctr->total_vt = s_excl_vt + blk_vt;
    ctr->total_act = s_excl_act + blk_act;
  }
}

// ------------------------------------------------------------------------------------------------
// used-edge words of an owner word, computed cooperatively by one warp (exact resolution via ballot)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void owner_used_warp(const Grid<T>& g, int i, int j, int w, uint32_t used[7], uint32_t* Aword) {
  if (i >= g.n0 || j >= g.n1 || w >= g.W) {
#pragma unroll
    for (int d = 0; d < 7; ++d) used[d] = 0;
    if (Aword) *Aword = 0;
    return;
  }
  Planes pl;
  load_planes(g, g.bits, i, j, w, pl);
  cross_words(pl, used);
  if (Aword) *Aword = pl.P[0];
  if (g.any_near) {
    Planes npl;
    load_planes(g, g.nbits, i, j, w, npl);
    const unsigned lane = lane_id();
#pragma unroll
    for (int d = 1; d <= 7; ++d) {
      uint32_t c = used[d - 1] & npl.P[0] & dir_plane(npl, d);
      if (c) {                                   // warp-uniform
        bool r = ((c >> lane) & 1u) && edge_used_exact(g, i, j, w * 32 + (int)lane, d);
        uint32_t fix = __ballot_sync(0xffffffffu, r);
        used[d - 1] = (used[d - 1] & ~c) | fix;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 3: vertices.  One warp per active owner word; lane = owner point.
// ------------------------------------------------------------------------------------------------
template <typename T, typename G>
__device__ __forceinline__ void grad_at(const Grid<T>& g, int i, int j, int k, G out[3]) {
  const int n[3] = {g.n0, g.n1, g.n2};
  const int p[3] = {i, j, k};
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    int lo[3] = {i, j, k}, hi[3] = {i, j, k};
    lo[ax] = p[ax] > 0 ? p[ax] - 1 : 0;
    hi[ax] = p[ax] < n[ax] - 1 ? p[ax] + 1 : n[ax] - 1;
    G s = (p[ax] == 0 || p[ax] == n[ax] - 1) ? (G)1 : (G)0.5;
    G a = (G)g.f[((long long)hi[0] * g.n1 + hi[1]) * g.n2 + hi[2]];
    G b = (G)g.f[((long long)lo[0] * g.n1 + lo[1]) * g.n2 + lo[2]];
    out[ax] = (a - b) * s;
  }
}

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

struct Xform {
  double origin[3], delta[3];
};

template <typename T, typename G>
__global__ void __launch_bounds__(128) k_emit_verts(Grid<T> gin, const uint32_t* __restrict__ list_v, unsigned n_active,
                                                    const uint32_t* __restrict__ vbase, Xform xf, G* __restrict__ verts,
                                                    G* __restrict__ normals, unsigned long long* __restrict__ keys,
                                                    uint8_t* __restrict__ lowmin) {
  Grid<T> g = gin;
  g.any_near = (int)*gin.near_flag;
  const unsigned lane = lane_id();
  unsigned a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_active) return;
  const long long gw = list_v[a];
  long long row = gw / g.W;
  const int w = (int)(gw - row * g.W);
  const int i = (int)(row / g.n1), j = (int)(row - (long long)i * g.n1);
  uint32_t used[7], Aw;
  owner_used_warp(g, i, j, w, used, &Aw);
  const uint32_t below = (1u << lane) - 1u;
  unsigned rank = 0, m7 = 0;
#pragma unroll
  for (int d = 0; d < 7; ++d) {
    rank += __popc(used[d] & below);
    m7 |= ((used[d] >> lane) & 1u) << d;
  }
  if (!m7) return;
  unsigned id = vbase[gw] + rank;
  const int k = w * 32 + (int)lane;
  const bool p_low = (Aw >> lane) & 1u;
  const G fp = (G)g.f[((long long)i * g.n1 + j) * g.n2 + k];
  const G v = (G)g.v;
  G gp[3];
  if (normals) grad_at<T, G>(g, i, j, k, gp);
  const long long lin = (((long long)i + g.plane_offset) * g.n1 + j) * g.n2 + k;
#pragma unroll 1
  for (int d = 1; d <= 7; ++d) {
    if (!((m7 >> (d - 1)) & 1u)) continue;
    const int di = (d >> 2) & 1, dj = (d >> 1) & 1, dk = d & 1;
    const G fq = (G)g.f[((long long)(i + di) * g.n1 + (j + dj)) * g.n2 + (k + dk)];
    // tetrahedral.py:476-487: key oriented (low, high) by value; ratio = (z-flow)/(fhigh-flow), 0.5 if ~0
    const G flow = p_low ? fp : fq, fhigh = p_low ? fq : fp;
    const G den = fhigh - flow;
    G ratio = (fabs((double)den) <= 1e-8) ? (G)0.5 : (v - flow) / den;
    // x = low + ratio*(high - low); (high-low) is +-1 or 0 per axis so the product is exact
    const G sgn = p_low ? (G)1 : (G)-1;
    const int pl[3] = {p_low ? i : i + di, p_low ? j : j + dj, p_low ? k : k + dk};
    const int dd[3] = {di, dj, dk};
    G pos[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      G base = (G)pl[ax];
      if (ax == 0) base = (G)((long long)pl[0] + g.plane_offset);
      G x = dd[ax] ? add_rn(base, mul_rn(ratio, sgn)) : base;
      pos[ax] = add_rn(mul_rn(x, (G)xf.delta[ax]), (G)xf.origin[ax]);   // grid_field.py:93
    }
    verts[(size_t)id * 3 + 0] = pos[0];
    verts[(size_t)id * 3 + 1] = pos[1];
    verts[(size_t)id * 3 + 2] = pos[2];
    if (normals) {
      G gq[3];
      grad_at<T, G>(g, i + di, j + dj, k + dk, gq);
      G nn[3], len2 = 0;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        G gl = p_low ? gp[ax] : gq[ax], gh = p_low ? gq[ax] : gp[ax];
        nn[ax] = (gl + ratio * (gh - gl)) / (G)xf.delta[ax];
        len2 += nn[ax] * nn[ax];
      }
      G len = sqrt(len2);
      normals[(size_t)id * 3 + 0] = len > (G)0 ? nn[0] / len : (G)0;
      normals[(size_t)id * 3 + 1] = len > (G)0 ? nn[1] / len : (G)0;
      normals[(size_t)id * 3 + 2] = len > (G)0 ? nn[2] / len : (G)0;
    }
    if (keys) {
      keys[id] = ((unsigned long long)lin << 3) | (unsigned)d;
      lowmin[id] = p_low ? 1 : 0;
    }
    ++id;
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 4: triangles.  One warp per active voxel word; lane = voxel.
// ------------------------------------------------------------------------------------------------
constexpr int ET_WARPS = 4;

template <typename T>
__global__ void __launch_bounds__(ET_WARPS * 32) k_emit_tris(Grid<T> gin, const uint32_t* __restrict__ list_t,
                                                             unsigned n_active, const uint32_t* __restrict__ vbase,
                                                             const uint32_t* __restrict__ tbase, int* __restrict__ tris) {
  __shared__ unsigned s_ids[ET_WARPS][19][32];
  Grid<T> g = gin;
  g.any_near = (int)*gin.near_flag;
  const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
  unsigned a = blockIdx.x * ET_WARPS + wib;
  if (a >= n_active) return;
  const long long gw = list_t[a];
  long long row = gw / g.W;
  const int w = (int)(gw - row * g.W);
  const int i = (int)(row / g.n1), j = (int)(row - (long long)i * g.n1);

  // owner (idbase, mask7) at bit `lane` and `lane+1` for the 4 owner rows (a,b)
  unsigned idb[8], msk[8];   // index = corner s = a*4 + b*2 + dk
#pragma unroll
  for (int ab = 0; ab < 4; ++ab) {
    const int ii = i + (ab >> 1), jj = j + (ab & 1);
    uint32_t u[7], un[7];
    owner_used_warp(g, ii, jj, w, u, nullptr);
    const bool lane31_next = (w + 1 < g.W);
    if (lane31_next) owner_used_warp(g, ii, jj, w + 1, un, nullptr);   // warp-uniform branch
    const long long wi = ((long long)ii * g.n1 + jj) * g.W + w;
    const bool row_ok = ii < g.n0 && jj < g.n1;
    const unsigned base0 = row_ok ? vbase[wi] : 0u;
    const uint32_t below = (1u << lane) - 1u;
    unsigned rank = 0, m0 = 0, m1 = 0;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      rank += __popc(u[d] & below);
      m0 |= ((u[d] >> lane) & 1u) << d;
      if (lane < 31) m1 |= ((u[d] >> (lane + 1)) & 1u) << d;
    }
    unsigned id0 = base0 + rank, id1 = id0 + __popc(m0);
    if (lane == 31) {
      m1 = 0;
      id1 = 0;
      if (lane31_next && row_ok) {
        id1 = vbase[wi + 1];
#pragma unroll
        for (int d = 0; d < 7; ++d) m1 |= (un[d] & 1u) << d;
      }
    }
    idb[ab * 2 + 0] = id0;
    msk[ab * 2 + 0] = m0;
    idb[ab * 2 + 1] = id1;
    msk[ab * 2 + 1] = m1;
  }
#pragma unroll
  for (int e = 0; e < 19; ++e) {
    const int s = c_edge_s[e], d = c_edge_d[e];               // constant-bank loads (uniform)
    s_ids[wib][e][lane] = idb[s] + __popc(msk[s] & ((1u << (d - 1)) - 1u));
  }
  // voxel classification
  Planes pl;
  load_planes(g, g.bits, i, j, w, pl);
  const bool cell_ok = pl.has_i1 && pl.has_j1 && ((pl.kp1 >> lane) & 1u);
  unsigned cb = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) cb |= ((corner_plane(pl, c) >> lane) & 1u) << c;
  unsigned tm[6];
  unsigned emit = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    unsigned m = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) m |= ((cb >> c_tet[t][b]) & 1u) << b;
    tm[t] = m;
    if (cell_ok && m != 0 && m != 15) emit |= 1u << t;
  }
  if (g.any_near && emit) {
    Planes npl;
    load_planes(g, g.nbits, i, j, w, npl);
    unsigned nb = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) nb |= ((corner_plane(npl, c) >> lane) & 1u) << c;
    bool cand = false;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      bool alln = true;
#pragma unroll
      for (int b = 0; b < 4; ++b) alln = alln && ((nb >> c_tet[t][b]) & 1u);
      cand = cand || (alln && ((emit >> t) & 1u));
    }
    if (cand) emit = cell_emit_exact(g, i, j, w * 32 + (int)lane, nullptr);
  }
  unsigned nt = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t)
    if ((emit >> t) & 1u) nt += c_tri_n[t][tm[t]];
  unsigned incl = warp_incl_scan_u32(nt);
  size_t o = (size_t)tbase[gw] + incl - nt;
  __syncwarp();
#pragma unroll 1
  for (int t = 0; t < 6; ++t) {
    if (!((emit >> t) & 1u)) continue;
    const unsigned m = tm[t];
    const int n = c_tri_n[t][m];
    for (int q = 0; q < n; ++q) {
      int* dst = tris + o * 3;
      dst[0] = (int)s_ids[wib][c_tri_e[t][m][q * 3 + 0]][lane];
      dst[1] = (int)s_ids[wib][c_tri_e[t][m][q * 3 + 1]][lane];
      dst[2] = (int)s_ids[wib][c_tri_e[t][m][q * 3 + 2]][lane];
      ++o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Parity output: (voxel, case code) of emitting voxels, recomputed from the samples (independent of the
// bit logic above).  Unordered (slot by atomic).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_codes(Grid<T> g, const uint32_t* __restrict__ list_t, unsigned n_active,
                                               long long* __restrict__ cells, uint32_t* __restrict__ codes,
                                               Counters* ctr) {
  const unsigned lane = lane_id();
  unsigned a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_active) return;
  const long long gw = list_t[a];
  long long row = gw / g.W;
  const int w = (int)(gw - row * g.W);
  const int i = (int)(row / g.n1), j = (int)(row - (long long)i * g.n1);
  const int k = w * 32 + (int)lane;
  unsigned code = 0;
  unsigned e = cell_emit_exact(g, i, j, k, &code);
  if (e) {
    unsigned slot = atomicAdd(&ctr->n_codes, 1u);
    cells[slot] = (((long long)i + g.plane_offset) * (g.n1 - 1) + j) * (long long)(g.n2 - 1) + k;
    codes[slot] = code;
  }
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
  if (ctx->device < 64) g_tables_loaded[ctx->device] = true;
  return 0;
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

template <typename T>
int run_typed(ctr_ctx* ctx, const ctr_mt3d_params* p, ctr_mt3d_counts* out) {
  const int n0 = (int)p->n0, n1 = (int)p->n1, n2 = (int)p->n2;
  const int W = (n2 + 31) / 32;
  const long long nrows = (long long)n0 * n1;
  const long long nwords = nrows * W;
  const size_t nsamp = (size_t)nrows * n2;
  cudaStream_t st = ctx->stream;
  if (nwords >= (1ll << 31)) return ctr_fail(ctx, CTR_ERR_UNSUPPORTED, "volume too large for 32-bit word indices");
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
  if ((rc = ctr_ensure(ctx, ctx->bits, (size_t)(nwords + 1) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->nbits, (size_t)(nwords + 1) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->vbase, (size_t)(nwords + 1) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->tbase, (size_t)(nwords + 1) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->list_v, (size_t)(nwords + 1) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->list_t, (size_t)(nwords + 1) * 4))) return rc;
  if ((rc = ctr_ensure(ctx, ctx->counters, sizeof(Counters)))) return rc;
  if (!ctx->counters_host) CTR_CUDA(ctx, cudaMallocHost(&ctx->counters_host, 256));

  Counters init;
  memset(&init, 0, sizeof init);
  init.min_key = ~0ull;
  init.max_key = 0ull;
  memcpy(ctx->counters_host, &init, sizeof init);
  CTR_CUDA(ctx, cudaMemcpyAsync(ctx->counters.p, ctx->counters_host, sizeof init, cudaMemcpyHostToDevice, st));
  Counters* dctr = (Counters*)ctx->counters.p;

  T thr, nlo, nhi;
  thresholds<T>(p->isovalue, thr, nlo, nhi);
  {
    const int warps_per_block = 8;
    long long need = (nwords + warps_per_block * 4 - 1) / (warps_per_block * 4);
    int blocks = (int)std::min<long long>(need, (long long)ctx->sm_count * 8);
    if (blocks < 1) blocks = 1;
    k_bitplane<T, 4><<<blocks, 256, 0, st>>>(dfield, nrows, n2, W, thr, nlo, nhi, (uint32_t*)ctx->bits.p,
                                             (uint32_t*)ctx->nbits.p, dctr);
    ctx->launches++;
  }
  ctr_stage_mark(ctx, 2);
  Counters h;

  Grid<T> g;
  g.f = dfield;
  g.bits = (const uint32_t*)ctx->bits.p;
  g.nbits = (const uint32_t*)ctx->nbits.p;
  g.n0 = n0; g.n1 = n1; g.n2 = n2; g.W = W;
  g.i_lo = (int)p->i_lo;
  g.i_hi = (int)p->i_hi;
  g.i_hiv = std::min(g.i_hi + 1, n0);
  g.plane_offset = p->plane_offset;
  g.v = p->isovalue;
  g.tolv = 1e-8 + 1e-5 * fabs(p->isovalue);
  g.any_near = 0;
  g.near_flag = &dctr->any_near;

  const long long plane_words = (long long)n1 * W;
  const long long word0 = (long long)g.i_lo * plane_words;
  const long long nscan = (long long)(g.i_hiv - g.i_lo) * plane_words;
  const int ntiles = (int)((nscan + CS_TILE - 1) / CS_TILE);
  if ((rc = ctr_ensure(ctx, ctx->tile_state, (size_t)ntiles * 16 + 16))) return rc;
  CTR_CUDA(ctx, cudaMemsetAsync(ctx->tile_state.p, 0, (size_t)ntiles * 16, st));
  unsigned long long* st_vt = (unsigned long long*)ctx->tile_state.p;
  unsigned long long* st_act = st_vt + ntiles;
  if (ntiles > 0) {
    k_count_scan<T><<<ntiles, CS_THREADS, 0, st>>>(g, word0, nscan, (uint32_t*)ctx->vbase.p, (uint32_t*)ctx->tbase.p,
                                                   (uint32_t*)ctx->list_v.p, (uint32_t*)ctx->list_t.p, st_vt, st_act,
                                                   dctr, ntiles);
    ctx->launches++;
  }
  ctr_stage_mark(ctx, 3);
  CTR_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(&h, ctx->counters_host, sizeof h);
  const unsigned long long totV = h.total_vt & 0x7fffffffull, totT = h.total_vt >> 31;
  const unsigned actV = (unsigned)(h.total_act & 0x7fffffffull), actT = (unsigned)(h.total_act >> 31);
  const unsigned long long nV = (g.i_hiv > g.i_hi) ? h.v_emit : totV;
  if (totV >= 0x7fffffffull || totT >= 0x7fffffffull)
    return ctr_fail(ctx, CTR_ERR_OVERFLOW, "more than 2^31 vertices or triangles in one call; shard the volume");

  out->n_verts = (int64_t)nV;
  out->n_tris = (int64_t)totT;
  out->n_active_cells = (int64_t)h.n_cells;
  out->n_crossings = (int64_t)h.n_cross;
  out->n_codes = 0;
  out->fmin = key_to_double(h.min_key);
  out->fmax = key_to_double(h.max_key);

  const bool f64 = (p->flags & CTR_GEOM_F64) != 0;
  const size_t gsz = f64 ? 8 : 4;
  if (!(p->flags & CTR_NO_GEOMETRY)) {
    if ((rc = ctr_ensure(ctx, ctx->verts, (size_t)nV * 3 * gsz))) return rc;
    if (p->flags & CTR_WANT_NORMALS)
      if ((rc = ctr_ensure(ctx, ctx->normals, (size_t)nV * 3 * gsz))) return rc;
    if (p->flags & CTR_WANT_KEYS) {
      if ((rc = ctr_ensure(ctx, ctx->keys, (size_t)nV * 8))) return rc;
      if ((rc = ctr_ensure(ctx, ctx->lowmin, (size_t)nV))) return rc;
    }
    if ((rc = ctr_ensure(ctx, ctx->tris, (size_t)totT * 12))) return rc;
    Xform xf;
    for (int a = 0; a < 3; ++a) {
      xf.origin[a] = p->origin[a];
      xf.delta[a] = p->delta[a];
    }
    unsigned long long* dkeys = (p->flags & CTR_WANT_KEYS) ? (unsigned long long*)ctx->keys.p : nullptr;
    uint8_t* dlow = (p->flags & CTR_WANT_KEYS) ? (uint8_t*)ctx->lowmin.p : nullptr;
    if (actV) {
      int blocks = (int)((actV + 3) / 4);
      if (f64) {
        k_emit_verts<T, double><<<blocks, 128, 0, st>>>(g, (const uint32_t*)ctx->list_v.p, actV,
                                                        (const uint32_t*)ctx->vbase.p, xf, (double*)ctx->verts.p,
                                                        (p->flags & CTR_WANT_NORMALS) ? (double*)ctx->normals.p : nullptr,
                                                        dkeys, dlow);
      } else {
        k_emit_verts<T, float><<<blocks, 128, 0, st>>>(g, (const uint32_t*)ctx->list_v.p, actV,
                                                       (const uint32_t*)ctx->vbase.p, xf, (float*)ctx->verts.p,
                                                       (p->flags & CTR_WANT_NORMALS) ? (float*)ctx->normals.p : nullptr,
                                                       dkeys, dlow);
      }
      ctx->launches++;
    }
    ctr_stage_mark(ctx, 4);
    if (actT) {
      int blocks = (int)((actT + ET_WARPS - 1) / ET_WARPS);
      k_emit_tris<T><<<blocks, ET_WARPS * 32, 0, st>>>(g, (const uint32_t*)ctx->list_t.p, actT,
                                                       (const uint32_t*)ctx->vbase.p, (const uint32_t*)ctx->tbase.p,
                                                       (int*)ctx->tris.p);
      ctx->launches++;
    }
    ctr_stage_mark(ctx, 5);
  } else {
    ctr_stage_mark(ctx, 4);
    ctr_stage_mark(ctx, 5);
  }
  if (p->flags & CTR_WANT_CODES) {
    if ((rc = ctr_ensure(ctx, ctx->cells, (size_t)h.n_cells * 8 + 8))) return rc;
    if ((rc = ctr_ensure(ctx, ctx->codes, (size_t)h.n_cells * 4 + 4))) return rc;
    if (actT) {
      k_codes<T><<<(actT + 3) / 4, 128, 0, st>>>(g, (const uint32_t*)ctx->list_t.p, actT, (long long*)ctx->cells.p,
                                                 (uint32_t*)ctx->codes.p, dctr);
      ctx->launches++;
    }
    out->n_codes = (int64_t)h.n_cells;
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
