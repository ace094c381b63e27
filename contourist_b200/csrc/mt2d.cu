// 2D marching triangles, all levels fused in one pass over the field (sm_100a).
//
//   k2d_count : dense pass.  A CTA owns a tile of 32 x 1024 squares: it reads the 33 x 1025 samples ONCE
//               (the field is never re-read densely, whatever the number of levels), classifies each sample
//               against the sorted level list (lt = #levels < f, eq = f equals a level), counts the contour
//               segments of every square for all levels from the four corner classes, and joins a
//               decoupled-lookback scan -> compacted, ordered list of active squares with segment offsets.
//   k2d_emit  : sparse pass.  One thread per active square: re-reads its 4 samples, evaluates the
//               reference's inclusive crossing predicate per level in fp64 and writes the segments
//               (level tag, two edge keys, two interpolated end points).
//   k_minmax  : field min / max for Linear2DContour.get_values (multiple_2d_contour.py:100-108).
//
// Reference lines (under /root/reference/contourist/): triangulated.py:10-14 (triangulated grid),
// :347-362 (key (low, high) exists iff f(low) <= z <= f(high); ratio 0.5 when |den| <= 1e-8),
// :66-77 + :295-305 (two keys are joined iff they are adjacent_pairs of each other = share their low or
// their high end inside one grid triangle), multiple_2d_contour.py:50-75 (level classification of edges).
#include <math.h>
#include <cmath>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int MAXL = 64;
constexpr int T2_COLS = 1024;                       // squares per tile row: 4 consecutive columns per thread
constexpr int T2_THREADS = 256;
constexpr int T2_ROWS = 32;

template <typename T>
struct Levels2D {
  T up[MAXL];      // smallest T strictly greater than level l:  level_l < f  <=>  f >= up[l]
  T eqv[MAXL];     // level l as a T when exactly representable, else NaN
  double z[MAXL];
  int n;
  T g0, ginv;      // first guess of the class of f: (f - g0) * ginv + 1 (exact for evenly spaced levels; fixed up otherwise)
};

struct Counters2D {
  unsigned long long total;     // packed: segments << 31 | active squares (written by k2d_tile_scan)
  unsigned long long min_key, max_key;
  unsigned int slots, pad;      // list slots handed out to the tiles (= active squares)
};

__device__ __forceinline__ unsigned long long order_key2(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// lt = number of levels strictly below f, eq = f equals level lt.  The thresholds live in shared memory, padded:
// upp[0] = -inf, upp[1..n] = up[0..n-1], upp[n+1] = NaN (never reached), eqp[n] = NaN, so class k is right iff
// f >= upp[k] and not f >= upp[k+1]
// and no index needs a bound check.  A guess (from the level spacing, or the neighbour's class: fields are smooth) is
// verified with two comparisons and only walked when wrong, so the result is exact for any increasing level list.
template <typename T>
__device__ __forceinline__ int class_fix(const T* __restrict__ upp, T f, int k) {
  if (!((f >= upp[k]) && !(f >= upp[k + 1]))) {
    if (f != f) return 0;                            // NaN: below every level, equal to none
    while (f >= upp[k + 1]) ++k;
    while (!(f >= upp[k])) --k;
  }
  return k;
}
template <typename T>
__device__ __forceinline__ int class_guess(int n, T g0, T ginv, T f) {
  const T q = fmin((f - g0) * ginv, (T)n);           // NaN -> n
  return f < g0 ? 0 : min(n, (int)q + 1);
}
template <typename T>
__device__ __forceinline__ unsigned class_byte(const T* __restrict__ eqp, T f, int k) {
  return (unsigned)k | ((f == eqp[k]) ? 128u : 0u);
}

__device__ __forceinline__ int tri_count(unsigned a, unsigned b, unsigned c) {
  // classes: lt in bits 0..6, eq in bit 7; le = lt + eq.  Segments of one triangle over all levels.
  const int lta = a & 127, ltb = b & 127, ltc = c & 127;
  const int lea = lta + (a >> 7), leb = ltb + (b >> 7), lec = ltc + (c >> 7);
  int n = 0;
  n += max(0, min(leb, lec) - lta) + max(0, lea - max(ltb, ltc));
  n += max(0, min(lea, lec) - ltb) + max(0, leb - max(lta, ltc));
  n += max(0, min(lea, leb) - ltc) + max(0, lec - max(lta, ltb));
  return n;
}

struct Shared2D {
  uint32_t cls[T2_ROWS + 1][T2_THREADS + 1];         // 4 class bytes per word (columns 4t..4t+3); word 256 = halo column
  uint32_t act[T2_ROWS * (T2_COLS / 32)];            // bitmap of the squares that may emit, row-major
  unsigned long long warp_sum[T2_THREADS / 32];
  unsigned long long excl;                           // first list slot of the tile
};

__device__ __forceinline__ unsigned cls_at(const Shared2D& sh, int r, int c) {   // class byte of sample (r, c) of the tile
  return (sh.cls[r][c >> 2] >> (8 * (c & 3))) & 255u;
}

template <typename T>
__global__ void __launch_bounds__(T2_THREADS) k2d_count(const T* __restrict__ f, int n0, int n1, int i_lo, int i_hi,
                                                        Levels2D<T> lv, int tiles_j, int ntiles, int vec_ok,
                                                        uint32_t* __restrict__ sq_lin, uint32_t* __restrict__ sq_base,
                                                        uint32_t* __restrict__ sq_cls, unsigned cap,
                                                        unsigned long long* __restrict__ tile_cnt, Counters2D* ctr) {
  __shared__ Shared2D sh;
  __shared__ T s_upp[MAXL + 2], s_eqp[MAXL + 1], s_lim[MAXL + 1];   // s_lim[k]: samples of class k below it carry no equality flag
  const int nl = lv.n;
  if ((int)threadIdx.x <= nl + 1) s_upp[threadIdx.x] = threadIdx.x == 0 ? (T)-INFINITY : ((int)threadIdx.x <= nl ? lv.up[threadIdx.x - 1] : (T)NAN);
  if ((int)threadIdx.x <= nl) {
    const T e = (int)threadIdx.x < nl ? lv.eqv[threadIdx.x] : (T)NAN;
    s_eqp[threadIdx.x] = e;
    s_lim[threadIdx.x] = (int)threadIdx.x < nl ? (e == e ? e : lv.up[threadIdx.x]) : (T)INFINITY;
  }
  __syncthreads();
  const int tile = (int)blockIdx.x;
  const int ti = tile / tiles_j, tj = tile - ti * tiles_j;
  const int i0 = i_lo + ti * T2_ROWS, j0 = tj * T2_COLS;
  const int t = threadIdx.x;
  const unsigned lane = lane_id(), warp = t >> 5;
  const int jc = j0 + 4 * t;                         // first of this thread's 4 columns

  // ---- classify 33 x 1025 samples: 4 consecutive columns per thread (one 128-bit load per row), 3 rows in flight
  constexpr int RB = 3;
  static_assert((T2_ROWS + 1) % RB == 0, "row batches");
  int kprev = 0;
  for (int r0 = 0; r0 <= T2_ROWS; r0 += RB) {
    T v[RB][4];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int i = i0 + r0 + q;
      const T* p = f + (size_t)i * n1 + jc;
      if (i < n0 && vec_ok && jc + 3 < n1) {
        if (sizeof(T) == 4) {
          const float4 x = *reinterpret_cast<const float4*>(p);
          v[q][0] = (T)x.x; v[q][1] = (T)x.y; v[q][2] = (T)x.z; v[q][3] = (T)x.w;
        } else {
          const double2 x = *reinterpret_cast<const double2*>(p), y = *reinterpret_cast<const double2*>(p + 2);
          v[q][0] = (T)x.x; v[q][1] = (T)x.y; v[q][2] = (T)y.x; v[q][3] = (T)y.y;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[q][u] = (i < n0 && jc + u < n1) ? p[u] : (T)NAN;
      }
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      uint32_t word;
      int k = (r0 + q == 0) ? class_fix(s_upp, v[q][0], class_guess(nl, lv.g0, lv.ginv, v[q][0])) : kprev;
      // fast path: all four samples strictly inside class k (fields are smooth: almost always) -> 2 comparisons for 4 samples
      const T lo4 = fmin(fmin(v[q][0], v[q][1]), fmin(v[q][2], v[q][3]));
      const T hi4 = fmax(fmax(v[q][0], v[q][1]), fmax(v[q][2], v[q][3]));
      const bool nan4 = (v[q][0] != v[q][0]) | (v[q][1] != v[q][1]) | (v[q][2] != v[q][2]) | (v[q][3] != v[q][3]);
      if (lo4 >= s_upp[k] && hi4 < s_lim[k] && !nan4) {
        word = (uint32_t)k * 0x01010101u;
      } else {
        // second path, branch-free per sample: every sample within one class of k (smooth fields: nearly always when
        // the first test fails) -- the four thresholds and three level values around k are loaded once for the group
        const T t_lo = s_upp[max(k - 1, 0)], t_a = s_upp[k], t_b = s_upp[k + 1], t_c = s_upp[min(k + 2, nl + 1)];
        const T e_m = s_eqp[max(k - 1, 0)], e_0 = s_eqp[k], e_p = s_eqp[min(k + 1, nl)];
        bool all_ok = true;
        word = 0;
        int k0 = k;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const T f = v[q][u];
          const bool up = f >= t_b, dn = !(f >= t_a);
          const int ku = k + (up ? 1 : 0) - (dn ? 1 : 0);
          all_ok = all_ok && (up ? !(f >= t_c) : (dn ? (k > 0 && f >= t_lo) : true));
          const T ev = up ? e_p : (dn ? e_m : e_0);
          word |= ((uint32_t)ku | (f == ev ? 128u : 0u)) << (8 * u);
          if (u == 0) k0 = ku;
        }
        if (all_ok) {
          kprev = k0;
        } else {
          word = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            k = class_fix(s_upp, v[q][u], k);
            word |= class_byte(s_eqp, v[q][u], k) << (8 * u);
            if (u == 0) kprev = k;
          }
        }
      }
      sh.cls[r0 + q][t] = word;
    }
  }
  if (t <= T2_ROWS) {                                  // halo column j0 + 1024: one row per thread
    uint32_t ch = 0;
    if (i0 + t < n0 && j0 + T2_COLS < n1) {
      const T x = f[(size_t)(i0 + t) * n1 + j0 + T2_COLS];
      ch = class_byte(s_eqp, x, class_fix(s_upp, x, class_guess(nl, lv.g0, lv.ginv, x)));
    }
    sh.cls[t][T2_THREADS] = ch;
  }
  __syncthreads();

  // ---- squares that may emit: 4 per thread and row, decided on packed class bytes (A = B = C = D without an
  // equality flag <=> no level touches the square); 8 lanes make one bitmap word
  for (int r = 0; r < T2_ROWS; ++r) {
    const int i = i0 + r;
    const uint32_t W0 = sh.cls[r][t], W1 = sh.cls[r][t + 1], X0 = sh.cls[r + 1][t], X1 = sh.cls[r + 1][t + 1];
    const uint32_t B = __funnelshift_r(W0, W1, 8), D = __funnelshift_r(X0, X1, 8);
    const uint32_t m = (W0 ^ B) | (W0 ^ X0) | (W0 ^ D) | (W0 & 0x80808080u);
    uint32_t nib = 0;
    if (m && i + 1 < n0 && i < i_hi) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (((m >> (8 * u)) & 255u) && jc + u + 1 < n1) nib |= 1u << u;
    }
    uint32_t wbits = nib << (4 * (lane & 7u));
    wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
    wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
    wbits |= __shfl_xor_sync(0xffffffffu, wbits, 4);
    if ((lane & 7u) == 0) sh.act[r * (T2_COLS / 32) + (t >> 3)] = wbits;
  }
  __syncthreads();

  // ---- count: thread t owns bitmap words 4t..4t+3 (row-major order of the tile's squares)
  unsigned long long loc = 0;
#pragma unroll 1
  for (int u = 0; u < 4; ++u) {
    const int q = 4 * t + u;
    uint32_t bits = sh.act[q];
    const int r = q / (T2_COLS / 32), c0 = (q % (T2_COLS / 32)) * 32;
    while (bits) {
      const int c = c0 + __ffs(bits) - 1;
      bits &= bits - 1;
      const unsigned A = cls_at(sh, r, c), Bc = cls_at(sh, r, c + 1), C = cls_at(sh, r + 1, c), Dc = cls_at(sh, r + 1, c + 1);
      const unsigned n = (unsigned)(tri_count(A, C, Dc) + tri_count(A, Bc, Dc));
      if (n) loc += ((unsigned long long)n << 31) | 1u;
    }
  }
  const unsigned long long inc = warp_incl_scan_u64(loc);
  if (lane == 31) sh.warp_sum[warp] = inc;
  __syncthreads();
  unsigned long long woff = 0, blk = 0;
#pragma unroll
  for (int q = 0; q < T2_THREADS / 32; ++q) {
    if (q < (int)warp) woff += sh.warp_sum[q];
    blk += sh.warp_sum[q];
  }
  // The tile's list entries go to slots from one atomic counter (tiles in any order, squares of a tile in order) and
  // carry segment offsets RELATIVE to the tile; k2d_tile_scan turns the per-tile segment counts into tile offsets, so
  // the segment arrays come out in tile order without any tile waiting for its predecessors (a look-back scan here
  // cost a fifth of the kernel: blocks finish out of order and sat polling).
  if (t == 0) {
    sh.excl = (blk & 0x7fffffffull) ? atomicAdd(&ctr->slots, (unsigned)(blk & 0x7fffffffull)) : 0u;
    tile_cnt[tile] = blk;
  }
  __syncthreads();
  unsigned long long run = woff + inc - loc;
  const unsigned slot0 = (unsigned)sh.excl;
  // ---- compact list of active squares with their tile-relative segment offsets
#pragma unroll 1
  for (int u = 0; u < 4; ++u) {
    const int q = 4 * t + u;
    uint32_t bits = sh.act[q];
    const int r = q / (T2_COLS / 32), c0 = (q % (T2_COLS / 32)) * 32;
    while (bits) {
      const int c = c0 + __ffs(bits) - 1;
      bits &= bits - 1;
      const unsigned A = cls_at(sh, r, c), Bc = cls_at(sh, r, c + 1), C = cls_at(sh, r + 1, c), Dc = cls_at(sh, r + 1, c + 1);
      const unsigned n = (unsigned)(tri_count(A, C, Dc) + tri_count(A, Bc, Dc));
      if (!n) continue;
      const unsigned slot = slot0 + (unsigned)(run & 0x7fffffffull);
      if (slot < cap) {
        sq_lin[slot] = (uint32_t)((size_t)(i0 + r) * n1 + (j0 + c));
        sq_base[slot] = (uint32_t)(run >> 31);
        sq_cls[slot] = A | (Bc << 8) | (C << 16) | (Dc << 24);
      }
      run += ((unsigned long long)n << 31) | 1u;
    }
  }
}

// exclusive scan of the per-tile segment counts (one block; a few thousand tiles) -> tile_off, grand totals
__global__ void __launch_bounds__(1024) k2d_tile_scan(const unsigned long long* __restrict__ tile_cnt, int ntiles,
                                                      uint32_t* __restrict__ tile_off, Counters2D* ctr) {
  // 8 consecutive tiles per thread, so the 8192 tiles of a 16384^2 field are one round of independent loads, one warp
  // scan and two barriers (one tile per thread took eight such rounds one after the other: 20 us of pure latency)
  constexpr int PER = 8;
  __shared__ unsigned long long s_warp[32];
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned long long carry = 0;
  for (int base = 0; base < ntiles; base += 1024 * PER) {
    const int q = base + (int)threadIdx.x * PER;
    unsigned long long c[PER], mine = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      c[u] = q + u < ntiles ? tile_cnt[q + u] : 0ull;
      mine += c[u];
    }
    const unsigned long long inc = warp_incl_scan_u64(mine);
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
      if (w < (int)warp) woff += s_warp[w];
      tot += s_warp[w];
    }
    unsigned long long run = carry + woff + inc - mine;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      if (q + u < ntiles) tile_off[q + u] = (uint32_t)(run >> 31);
      run += c[u];
    }
    carry += tot;
  }
  if (threadIdx.x == 0) ctr->total = carry;
}

__device__ __forceinline__ double mul_rn2(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn2(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn2(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn2(float a, float b) { return __fadd_rn(a, b); }

struct Xform2 {
  double origin[2], delta[2];
};

// write one end point: key (low -> high) between integer points, interpolated position
template <typename G>
__device__ __forceinline__ void put_end(int li, int lj, int hi_, int hj, double flow, double fhigh, double z, int n1,
                                        long long row_offset, const Xform2& xf, unsigned long long* key, G* pos) {
  const int pi = min(li, hi_), pj = min(lj, hj);
  const int d = (max(li, hi_) - pi) * 2 + (max(lj, hj) - pj);
  const unsigned lowmin = (li == pi && lj == pj) ? 1u : 0u;
  const unsigned long long lin = (unsigned long long)((long long)pi + row_offset) * (unsigned long long)n1 + (unsigned)pj;
  *key = ((lin * 4ull + (unsigned)d) << 1) | lowmin;
  const G fl = (G)flow, fh = (G)fhigh;
  const G den = fh - fl;
  const G ratio = (fabs((double)den) <= 1e-8) ? (G)0.5 : ((G)z - fl) / den;
  const G x0 = add_rn2((G)((long long)li + row_offset), mul_rn2(ratio, (G)(hi_ - li)));
  const G x1 = add_rn2((G)lj, mul_rn2(ratio, (G)(hj - lj)));
  pos[0] = add_rn2(mul_rn2(x0, (G)xf.delta[0]), (G)xf.origin[0]);
  pos[1] = add_rn2(mul_rn2(x1, (G)xf.delta[1]), (G)xf.origin[1]);
}

template <typename T, typename G>
__global__ void __launch_bounds__(128) k2d_emit(const T* __restrict__ f, int n1, long long row_offset, Levels2D<T> lv,
                                                const uint32_t* __restrict__ sq_lin, const uint32_t* __restrict__ sq_base,
                                                const uint32_t* __restrict__ sq_cls, unsigned n_sq,
                                                const uint32_t* __restrict__ tile_off, int i_lo, int tiles_j, Xform2 xf,
                                                uint8_t* __restrict__ seg_level, unsigned long long* __restrict__ seg_keys,
                                                G* __restrict__ seg_pos) {
  const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_sq) return;
  const unsigned lin = sq_lin[a];
  const int i = (int)(lin / (unsigned)n1), j = (int)(lin - (unsigned)i * (unsigned)n1);
  const uint32_t cls = sq_cls[a];
  size_t o = (size_t)tile_off[((i - i_lo) / T2_ROWS) * tiles_j + j / T2_COLS] + sq_base[a];
  const double fA = (double)f[(size_t)i * n1 + j], fB = (double)f[(size_t)i * n1 + j + 1];
  const double fC = (double)f[(size_t)(i + 1) * n1 + j], fD = (double)f[(size_t)(i + 1) * n1 + j + 1];
  int lmin = 127, lmax = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = (cls >> (8 * q)) & 255;
    lmin = min(lmin, c & 127);
    lmax = max(lmax, (c & 127) + (c >> 7));
  }
  // triangles T0 = {A, C, D} = {(i,j),(i+1,j),(i+1,j+1)}, T1 = {A, B, D} = {(i,j),(i,j+1),(i+1,j+1)}
#pragma unroll 1
  for (int tri = 0; tri < 2; ++tri) {
    const int pi[3] = {i, tri == 0 ? i + 1 : i, i + 1};
    const int pj[3] = {j, tri == 0 ? j : j + 1, j + 1};
    const double fv[3] = {fA, tri == 0 ? fC : fB, fD};
    const int cl[3] = {(int)(cls & 255u), (int)((cls >> (tri == 0 ? 16 : 8)) & 255u), (int)(cls >> 24)};
    const int lt[3] = {cl[0] & 127, cl[1] & 127, cl[2] & 127};
    const int le[3] = {lt[0] + (cl[0] >> 7), lt[1] + (cl[1] >> 7), lt[2] + (cl[2] >> 7)};
#pragma unroll 1
    for (int l = lmin; l < lmax; ++l) {
      const double z = lv.z[l];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int b = (v + 1) % 3, c = (v + 2) % 3;
        // the inclusive predicates f(low) <= z <= f(high), decided on the sample classes so that the
        // emission can never disagree with the counts of k2d_count (level l >= lt(x) <=> f(x) <= z_l, l < le(x) <=> z_l <= f(x))
        // low-isolated: keys (v -> b), (v -> c)
        if (l >= lt[v] && l < le[b] && l < le[c]) {
          seg_level[o] = (uint8_t)l;
          put_end<G>(pi[v], pj[v], pi[b], pj[b], fv[v], fv[b], z, n1, row_offset, xf, &seg_keys[o * 2], &seg_pos[o * 4]);
          put_end<G>(pi[v], pj[v], pi[c], pj[c], fv[v], fv[c], z, n1, row_offset, xf, &seg_keys[o * 2 + 1], &seg_pos[o * 4 + 2]);
          ++o;
        }
        // high-isolated: keys (b -> v), (c -> v)
        if (l >= lt[b] && l >= lt[c] && l < le[v]) {
          seg_level[o] = (uint8_t)l;
          put_end<G>(pi[b], pj[b], pi[v], pj[v], fv[b], fv[v], z, n1, row_offset, xf, &seg_keys[o * 2], &seg_pos[o * 4]);
          put_end<G>(pi[c], pj[c], pi[v], pj[v], fv[c], fv[v], z, n1, row_offset, xf, &seg_keys[o * 2 + 1], &seg_pos[o * 4 + 2]);
          ++o;
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_minmax(const T* __restrict__ f, size_t n, Counters2D* ctr) {
  T mn = INFINITY, mx = -INFINITY;
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) {
    T v = __ldg(f + s);
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane_id() == 0 && mn <= mx) {
    atomicMin(&ctr->min_key, order_key2((double)mn));
    atomicMax(&ctr->max_key, order_key2((double)mx));
  }
}

double key_to_double2(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
  memcpy(&d, &u, 8);
  return d;
}

template <typename T>
void make_levels(const double* z, int n, Levels2D<T>& lv);
template <>
void make_levels<double>(const double* z, int n, Levels2D<double>& lv) {
  lv.n = n;
  for (int l = 0; l < n; ++l) {
    lv.z[l] = z[l];
    lv.up[l] = nextafter(z[l], INFINITY);
    lv.eqv[l] = z[l];
  }
}
template <>
void make_levels<float>(const double* z, int n, Levels2D<float>& lv) {
  lv.n = n;
  for (int l = 0; l < n; ++l) {
    lv.z[l] = z[l];
    float c = (float)z[l];
    if ((double)c > z[l]) {
      lv.up[l] = c;                      // c is already the smallest float above the level
      lv.eqv[l] = NAN;
    } else if ((double)c == z[l]) {
      lv.up[l] = nextafterf(c, INFINITY);
      lv.eqv[l] = c;
    } else {
      lv.up[l] = nextafterf(c, INFINITY);
      lv.eqv[l] = NAN;
    }
  }
}

template <typename T>
void make_guess(Levels2D<T>& lv) {
  lv.g0 = lv.up[0];
  const double span = (double)lv.up[lv.n - 1] - (double)lv.up[0];
  lv.ginv = (lv.n > 1 && span > 0 && std::isfinite(span)) ? (T)((lv.n - 1) / span) : (T)0;
}

template <typename T>
int run2d(ctr_ctx* ctx, const ctr_mt2d_params* p, ctr_mt2d_counts* out) {
  const int n0 = (int)p->n0, n1 = (int)p->n1;
  const size_t nsamp = (size_t)n0 * n1;
  cudaStream_t st = ctx->stream;
  int rc;
  ctr_stage_mark(ctx, 0);
  const T* df;
  if (p->flags & CTR_FIELD_ON_DEVICE) {
    df = (const T*)p->field;
  } else {
    if ((rc = ctr_ensure(ctx, ctx->field, nsamp * sizeof(T)))) return rc;
    CTR_CUDA(ctx, cudaMemcpyAsync(ctx->field.p, p->field, nsamp * sizeof(T), cudaMemcpyHostToDevice, st));
    df = (const T*)ctx->field.p;
  }
  ctr_stage_mark(ctx, 1);
  Levels2D<T> lv = {};
  make_levels<T>(p->levels, p->nlevels, lv);
  make_guess<T>(lv);
  const int i_lo = (int)p->i_lo, i_hi = (int)std::min<int64_t>(p->i_hi, n0 - 1);
  const int rows = std::max(i_hi - i_lo, 0);
  const int tiles_i = (rows + T2_ROWS - 1) / T2_ROWS, tiles_j = (n1 - 1 + T2_COLS - 1) / T2_COLS;
  const int ntiles = tiles_i * tiles_j;
  const int vec_ok = ((n1 & 3) == 0 && (((uintptr_t)df) & 15) == 0) ? 1 : 0;      // rows start 16-byte aligned
  if ((rc = ctr_ensure(ctx, ctx->counters, 256))) return rc;
  if (!ctx->counters_host) CTR_CUDA(ctx, cudaMallocHost(&ctx->counters_host, 1024));
  if ((rc = ctr_ensure(ctx, ctx->tile_state, (size_t)ntiles * 12 + 32))) return rc;      // per-tile counts (u64), then offsets (u32)
  Counters2D* dctr = (Counters2D*)ctx->counters.p;
  unsigned long long* tile_cnt = (unsigned long long*)ctx->tile_state.p;
  uint32_t* tile_off = (uint32_t*)(tile_cnt + ntiles + 1);
  DevBuf& b_lin = ctx->aux[5];
  DevBuf& b_base = ctx->aux[6];
  DevBuf& b_cls = ctx->aux[7];
  size_t want = std::max<size_t>(nsamp / 16, 1 << 14);
  Counters2D h;
  unsigned long long nsq = 0, nseg = 0;
  for (int attempt = 0; attempt < 3; ++attempt) {
    if ((rc = ctr_ensure(ctx, b_lin, want * 4))) return rc;
    if ((rc = ctr_ensure(ctx, b_base, want * 4))) return rc;
    if ((rc = ctr_ensure(ctx, b_cls, want * 4))) return rc;
    const unsigned cap = (unsigned)std::min<size_t>(std::min(b_lin.cap, std::min(b_base.cap, b_cls.cap)) / 4, 0x7fffffffu);
    Counters2D init;
    memset(&init, 0, sizeof init);
    init.min_key = ~0ull;
    memcpy(ctx->counters_host, &init, sizeof init);
    CTR_CUDA(ctx, cudaMemcpyAsync(dctr, ctx->counters_host, sizeof init, cudaMemcpyHostToDevice, st));
    if (ntiles > 0) {
      k2d_count<T><<<ntiles, T2_THREADS, 0, st>>>(df, n0, n1, i_lo, i_hi, lv, tiles_j, ntiles, vec_ok, (uint32_t*)b_lin.p,
                                               (uint32_t*)b_base.p, (uint32_t*)b_cls.p, cap, tile_cnt, dctr);
      k2d_tile_scan<<<1, 1024, 0, st>>>(tile_cnt, ntiles, tile_off, dctr);
      ctx->launches += 2;
    }
    if (p->flags & CTR_WANT_MINMAX) {
      k_minmax<T><<<ctx->sm_count * 8, 256, 0, st>>>(df, nsamp, dctr);
      ctx->launches++;
    }
    CTR_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, dctr, sizeof(Counters2D), cudaMemcpyDeviceToHost, st));
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(&h, ctx->counters_host, sizeof h);
    nsq = h.total & 0x7fffffffull;
    nseg = h.total >> 31;
    if (nsq <= cap) break;
    if (attempt == 2) return ctr_fail(ctx, CTR_ERR_STATE, "list capacity did not converge");
    want = nsq;
  }
  ctr_stage_mark(ctx, 2);
  if (nseg >= 0x7ffffff0ull) return ctr_fail(ctx, CTR_ERR_OVERFLOW, "more than 2^31 segments in one call; shard the field");
  out->n_segments = (int64_t)nseg;
  out->n_active_squares = (int64_t)nsq;
  const bool mm = (p->flags & CTR_WANT_MINMAX) && h.min_key != ~0ull;
  out->fmin = mm ? key_to_double2(h.min_key) : NAN;
  out->fmax = mm ? key_to_double2(h.max_key) : NAN;
  const bool f64 = (p->flags & CTR_GEOM_F64) != 0;
  DevBuf& b_lvl = ctx->aux[8];
  DevBuf& b_keys = ctx->aux[9];
  DevBuf& b_pos = ctx->aux[10];
  if (!(p->flags & CTR_NO_GEOMETRY)) {
    if ((rc = ctr_ensure(ctx, b_lvl, (size_t)nseg + 16))) return rc;
    if ((rc = ctr_ensure(ctx, b_keys, (size_t)nseg * 16 + 16))) return rc;
    if ((rc = ctr_ensure(ctx, b_pos, (size_t)nseg * 4 * (f64 ? 8 : 4) + 16))) return rc;
    Xform2 xf;
    for (int a = 0; a < 2; ++a) {
      xf.origin[a] = p->origin[a];
      xf.delta[a] = p->delta[a];
    }
    if (nsq) {
      const int blocks = (int)((nsq + 127) / 128);
      if (f64)
        k2d_emit<T, double><<<blocks, 128, 0, st>>>(df, n1, p->row_offset, lv, (const uint32_t*)b_lin.p,
                                                    (const uint32_t*)b_base.p, (const uint32_t*)b_cls.p, (unsigned)nsq, tile_off, i_lo,
                                                    tiles_j, xf, (uint8_t*)b_lvl.p, (unsigned long long*)b_keys.p, (double*)b_pos.p);
      else
        k2d_emit<T, float><<<blocks, 128, 0, st>>>(df, n1, p->row_offset, lv, (const uint32_t*)b_lin.p,
                                                   (const uint32_t*)b_base.p, (const uint32_t*)b_cls.p, (unsigned)nsq, tile_off, i_lo,
                                                   tiles_j, xf, (uint8_t*)b_lvl.p, (unsigned long long*)b_keys.p, (float*)b_pos.p);
      ctx->launches++;
    }
  }
  ctr_stage_mark(ctx, 3);
  CTR_CUDA(ctx, cudaGetLastError());
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  if (ctx->timing) {
    for (int s = 0; s < 3; ++s) {
      ctx->stage_ms[s] = 0.f;
      if (ctx->ev_set[s] && ctx->ev_set[s + 1]) cudaEventElapsedTime(&ctx->stage_ms[s], ctx->ev[s], ctx->ev[s + 1]);
    }
  }
  ctx->last_kind = 2;
  ctx->poly_counts[0] = ctx->poly_counts[1] = 0;
  ctx->last_flags = p->flags;
  ctx->last_counts[0] = (int64_t)nseg;
  return 0;
}

}  // namespace

extern "C" int ctr_mt2d_run(ctr_ctx* ctx, const ctr_mt2d_params* p, ctr_mt2d_counts* out) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (!p || !out || !p->field || !p->levels) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "null argument");
  if (p->n0 < 2 || p->n1 < 2) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "grid must have at least 2 samples per axis");
  if ((unsigned long long)p->n0 * (unsigned long long)p->n1 >= (1ull << 32))
    return ctr_fail(ctx, CTR_ERR_UNSUPPORTED, "2D field too large for 32-bit sample indices; shard it");
  // the totals travel packed as segments << 31 | squares: the squares of one call must stay below 2^31
  if (p->i_hi > p->i_lo && (unsigned long long)(p->i_hi - p->i_lo) * (unsigned long long)(p->n1 - 1) >= (1ull << 31) - 16)
    return ctr_fail(ctx, CTR_ERR_UNSUPPORTED, "more than 2^31 squares in one call; pass a narrower row band [i_lo, i_hi)");
  if (p->nlevels < 1 || p->nlevels > MAXL) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "1..64 levels supported");
  for (int l = 0; l < p->nlevels; ++l) {
    if (!(p->levels[l] == p->levels[l])) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "NaN level");
    if (l && !(p->levels[l] > p->levels[l - 1])) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "levels must be strictly increasing");
  }
  if (p->i_lo < 0 || p->i_hi > p->n0 || p->i_lo >= p->i_hi) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "bad row range [i_lo, i_hi)");
  for (int a = 0; a < 2; ++a)
    if (p->delta[a] == 0.0) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "delta must be non-zero");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  memset(out, 0, sizeof *out);
  ctx->last_kind = 0;
  if (p->dtype == CTR_F32) return run2d<float>(ctx, p, out);
  if (p->dtype == CTR_F64) return run2d<double>(ctx, p, out);
  return ctr_fail(ctx, CTR_ERR_BAD_ARG, "dtype must be CTR_F32 or CTR_F64");
}

extern "C" int ctr_mt2d_fetch(ctr_ctx* ctx, uint8_t* seg_level, uint64_t* seg_keys, void* seg_pos) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (ctx->last_kind != 2) return ctr_fail(ctx, CTR_ERR_STATE, "no completed ctr_mt2d_run to fetch from");
  if (ctx->last_flags & CTR_NO_GEOMETRY) return ctr_fail(ctx, CTR_ERR_STATE, "run had CTR_NO_GEOMETRY");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->last_counts[0];
  const size_t gsz = (ctx->last_flags & CTR_GEOM_F64) ? 8 : 4;
  cudaStream_t st = ctx->stream;
  if (n) {
    if (seg_level) CTR_CUDA(ctx, cudaMemcpyAsync(seg_level, ctx->aux[8].p, n, cudaMemcpyDeviceToHost, st));
    if (seg_keys) CTR_CUDA(ctx, cudaMemcpyAsync(seg_keys, ctx->aux[9].p, n * 16, cudaMemcpyDeviceToHost, st));
    if (seg_pos) CTR_CUDA(ctx, cudaMemcpyAsync(seg_pos, ctx->aux[10].p, n * 4 * gsz, cudaMemcpyDeviceToHost, st));
  }
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}
