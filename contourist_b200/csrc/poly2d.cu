// Polylines of the 2D contours on the device: the reference chains its contour pairs a vertex at
// a time (triangulated.py:221-305 find_adjacencies / get_contour_sequences); here the segment soup of the last
// ctr_mt2d_run is ranked by pointer jumping over darts.
//
//   vertex  = (level, edge key); a key of one level lies on at most two segments unless a sample sits exactly on the level;
//   dart    = a segment with a direction: dart 2s+q leaves end q of segment s.  The successor of a dart arriving at a
//             vertex is the dart that leaves it along the vertex's OTHER segment, or -- at the end of an open contour --
//             the dart back along the same segment.  An open contour of m segments is then ONE cycle of 2m darts (there
//             and back), a closed contour of m segments TWO cycles of m darts (one per direction);
//   k_p_insert / k_p_succ : (level, key) -> open-addressing table of 16-byte slots holding the (up to two) darts that
//             leave the vertex; successor and priority of every dart;
//   k_p_min  : pointer jumping, log2(longest contour) rounds: every dart learns the FIRST dart of its cycle = the dart
//             that leaves the open contour's smaller end / the closed contour's smallest vertex;
//   k_p_cut / k_p_rank : every cycle cut in front of its first dart, list ranking by pointer jumping: position of every dart;
//   k_p_mark / scans / k_p_points : of a closed contour's two directions the one that starts towards the smaller
//             neighbour is kept, of an open contour's there-and-back tour the first half; points written in position
//             order, one contour after the other;
//   k_p_dedupe* : consecutive np.allclose points dropped (triangulated.py:269), a contour that returns to its start is closed.
// The numpy statement of the same rules is contourist_b200/triangulated.py rank_darts / chain_segments (host, kept for
// levels where a key has more than two segments: there the reference's result is a matter of visiting order, and the
// facade's fixed-order walk decides).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "uf_hash.cuh"

namespace {

constexpr unsigned NIL = 0xffffffffu;

struct PolyCounters {
  unsigned long long junction_levels;   // bit l: some key of level l has more than two segments
  unsigned changed;
  unsigned live;
  unsigned n_poly;
  unsigned n_points;
  unsigned n_points_final;
  unsigned pad;
};

struct __align__(16) VSlot {
  unsigned long long key;
  int a, b;                             // the darts that leave this vertex (-1: none)
};

__device__ __forceinline__ unsigned long long vkey(const uint8_t* __restrict__ level, const unsigned long long* __restrict__ keys,
                                                   unsigned dart) {
  // (level, key): keys are < 2^35 ((lin << 3 | d << 1 | lowmin), lin < 2^32), levels < 64
  return ((unsigned long long)level[dart >> 1] << 40) | keys[dart];
}

__device__ __forceinline__ VSlot* vslot(VSlot* tab, size_t mask, unsigned long long key, bool insert) {
  size_t s = (size_t)ufh::mix64(key) & mask;
  while (true) {
    unsigned long long k = tab[s].key;
    if (k == key) return tab + s;
    if (k == ufh::EMPTY) {
      if (!insert) return nullptr;
      k = atomicCAS(&tab[s].key, ufh::EMPTY, key);
      if (k == ufh::EMPTY || k == key) return tab + s;
    }
    s = (s + 1) & mask;
  }
}

__global__ void k_p_init(VSlot* tab, size_t nslots, PolyCounters* ctr) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)gridDim.x * blockDim.x;
  for (size_t q = t; q < nslots; q += n) reinterpret_cast<int4*>(tab)[q] = make_int4(-1, -1, -1, -1);
  if (t == 0) memset(ctr, 0, sizeof *ctr);
}

// dart d = 2s+q leaves end q of segment s: registered at the vertex of that end
__global__ void k_p_insert(const uint8_t* __restrict__ level, const unsigned long long* __restrict__ keys, unsigned nd, VSlot* tab,
                           size_t mask, PolyCounters* ctr) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  if (keys[d] == keys[d ^ 1u]) return;                           // a segment from a key to itself: no dart
  VSlot* s = vslot(tab, mask, vkey(level, keys, d), true);
  if (atomicCAS(&s->a, -1, (int)d) == -1) return;
  if (atomicCAS(&s->b, -1, (int)d) == -1) return;
  atomicOr(&ctr->junction_levels, 1ull << level[d >> 1]);        // a third segment at this key
}

// successor, priority: (0 for darts that leave an end of an open contour, else 1) << 63 | (level, key) of the tail.
// Within one cycle the smallest priority is unique: ends are visited once, and a closed contour's cycle leaves every
// vertex once.
__global__ void k_p_succ(const uint8_t* __restrict__ level, const unsigned long long* __restrict__ keys, unsigned nd,
                         const VSlot* __restrict__ tab, size_t mask, unsigned* __restrict__ succ,
                         unsigned long long* __restrict__ best_p, unsigned* __restrict__ best_d, unsigned* __restrict__ jump) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  if (keys[d] == keys[d ^ 1u]) {                                 // dead dart: a cycle of its own, never kept
    succ[d] = d;
    jump[d] = d;
    best_p[d] = ~0ull;
    best_d[d] = d;
    return;
  }
  const VSlot* hs = vslot(const_cast<VSlot*>(tab), mask, vkey(level, keys, d ^ 1u), false);     // vertex the dart arrives at
  const int rev = (int)(d ^ 1u);
  int other = hs->a == rev ? hs->b : hs->a;
  const unsigned sc = other >= 0 ? (unsigned)other : (unsigned)rev;                              // an end: turn round
  const unsigned long long tk = vkey(level, keys, d);
  const VSlot* ts = vslot(const_cast<VSlot*>(tab), mask, tk, false);
  const bool leaves_end = ts->b < 0;
  succ[d] = sc;
  jump[d] = sc;
  best_p[d] = (leaves_end ? 0ull : (1ull << 63)) | tk;
  best_d[d] = d;
}

// one round of pointer jumping: minimum over twice as many darts ahead
__global__ void k_p_min(unsigned nd, const unsigned long long* __restrict__ bp, const unsigned* __restrict__ bd,
                        const unsigned* __restrict__ jmp, unsigned long long* __restrict__ bp2, unsigned* __restrict__ bd2,
                        unsigned* __restrict__ jmp2, PolyCounters* ctr) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const unsigned j = jmp[d];
  unsigned long long p = bp[d];
  unsigned b = bd[d];
  const unsigned long long pj = bp[j];
  if (pj < p) {
    p = pj;
    b = bd[j];
    ctr->changed = 1u;
  }
  bp2[d] = p;
  bd2[d] = b;
  jmp2[d] = jmp[j];
}

__global__ void k_p_cut(unsigned nd, const unsigned* __restrict__ succ, const unsigned* __restrict__ start, unsigned* __restrict__ nxt,
                        unsigned* __restrict__ dist) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const bool last = succ[d] == start[d];
  nxt[d] = last ? NIL : succ[d];
  dist[d] = last ? 0u : 1u;
}

__global__ void k_p_rank(unsigned nd, const unsigned* __restrict__ nxt, const unsigned* __restrict__ dist, unsigned* __restrict__ nxt2,
                         unsigned* __restrict__ dist2, PolyCounters* ctr) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const unsigned n = nxt[d];
  if (n == NIL) {
    nxt2[d] = NIL;
    dist2[d] = dist[d];
    return;
  }
  dist2[d] = dist[d] + dist[n];
  const unsigned nn = nxt[n];
  nxt2[d] = nn;
  if (nn != NIL) ctr->live = 1u;
}

// per dart: is it the first dart of a kept contour, and how many points does that contour have
__global__ void k_p_mark(unsigned nd, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ start,
                         const unsigned* __restrict__ dist, unsigned* __restrict__ npts) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  unsigned n = 0;
  if (start[d] == d && keys[d] != keys[d ^ 1u]) {
    const unsigned len = dist[d] + 1u;
    const unsigned twin = start[d ^ 1u];
    if (twin == d) n = len / 2u + 1u;                                                 // open: m segments, m + 1 points
    else if (keys[d ^ 1u] < keys[twin ^ 1u]) n = len;                                   // closed, this direction kept: m points
  }
  npts[d] = n;
}

// exclusive scan of 32-bit counts, three kernels (block sums, one block over them, add back)
constexpr int PS_THREADS = 256, PS_PER = 8, PS_TILE = PS_THREADS * PS_PER;
__global__ void __launch_bounds__(PS_THREADS) k_ps_local(const unsigned* __restrict__ in, unsigned n, unsigned* __restrict__ out,
                                                         unsigned* __restrict__ block_sum, unsigned* __restrict__ nonzero_out) {
  __shared__ unsigned s_warp[PS_THREADS / 32], s_nz[PS_THREADS / 32];
  const unsigned q0 = blockIdx.x * PS_TILE + threadIdx.x * PS_PER;
  unsigned v[PS_PER], sum = 0, nz = 0;
#pragma unroll
  for (int u = 0; u < PS_PER; ++u) {
    v[u] = q0 + u < n ? in[q0 + u] : 0u;
    sum += v[u];
    nz += v[u] ? 1u : 0u;
  }
  const unsigned inc = warp_incl_scan_u32(sum), incz = warp_incl_scan_u32(nz);
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  if (lane == 31) {
    s_warp[warp] = inc;
    s_nz[warp] = incz;
  }
  __syncthreads();
  unsigned woff = 0, wz = 0, tot = 0, totz = 0;
#pragma unroll
  for (int w = 0; w < PS_THREADS / 32; ++w) {
    if (w < (int)warp) {
      woff += s_warp[w];
      wz += s_nz[w];
    }
    tot += s_warp[w];
    totz += s_nz[w];
  }
  unsigned run = woff + inc - sum, runz = wz + incz - nz;
#pragma unroll
  for (int u = 0; u < PS_PER; ++u) {
    if (q0 + u < n) {
      out[q0 + u] = run;
      if (nonzero_out) nonzero_out[q0 + u] = runz;
    }
    run += v[u];
    runz += v[u] ? 1u : 0u;
  }
  if (threadIdx.x == 0) {
    block_sum[2 * blockIdx.x] = tot;
    block_sum[2 * blockIdx.x + 1] = totz;
  }
}
__global__ void __launch_bounds__(1024) k_ps_blocks(unsigned* __restrict__ block_sum, unsigned nblocks, unsigned* __restrict__ total) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0ull;
  __syncthreads();
  for (unsigned base = 0; base < nblocks; base += 1024) {
    const unsigned q = base + threadIdx.x;
    const unsigned long long mine = q < nblocks ? ((unsigned long long)block_sum[2 * q + 1] << 32 | block_sum[2 * q]) : 0ull;
    const unsigned long long inc = warp_incl_scan_u64(mine);
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned long long woff = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < (int)(threadIdx.x >> 5)) woff += s_warp[w];
      tot += s_warp[w];
    }
    const unsigned long long ex = s_carry + woff + inc - mine;
    if (q < nblocks) {
      block_sum[2 * q] = (unsigned)(ex & 0xffffffffull);
      block_sum[2 * q + 1] = (unsigned)(ex >> 32);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    total[0] = (unsigned)(s_carry & 0xffffffffull);
    total[1] = (unsigned)(s_carry >> 32);
  }
}
__global__ void k_ps_add(unsigned* __restrict__ out, unsigned* __restrict__ nonzero_out, unsigned n, const unsigned* __restrict__ block_sum) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  out[q] += block_sum[2 * (q / PS_TILE)];
  if (nonzero_out) nonzero_out[q] += block_sum[2 * (q / PS_TILE) + 1];
}

// contour descriptors + points: the first dart writes (level, closed, start key, first point), every kept dart the
// point it arrives at
template <typename G>
__global__ void k_p_points(unsigned nd, const uint8_t* __restrict__ level, const unsigned long long* __restrict__ keys,
                           const G* __restrict__ pos, const unsigned* __restrict__ start, const unsigned* __restrict__ dist,
                           const unsigned* __restrict__ npts, const unsigned* __restrict__ off, const unsigned* __restrict__ pidx,
                           G* __restrict__ pts, unsigned* __restrict__ poly_of_point, int* __restrict__ p_level,
                           uint8_t* __restrict__ p_closed, unsigned long long* __restrict__ p_key, unsigned* __restrict__ p_off) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const unsigned sd = start[d];
  const unsigned n = npts[sd];
  if (!n) return;
  const unsigned len = dist[sd] + 1u;
  const unsigned p = len - 1u - dist[d];                      // position of this dart behind the first one
  const bool open = start[sd ^ 1u] == sd;
  const unsigned o = off[sd], pl = pidx[sd];
  if (d == sd) {
    p_level[pl] = (int)level[d >> 1];
    p_closed[pl] = open ? 0 : 1;
    p_key[pl] = keys[d];
    p_off[pl] = o;
    pts[(size_t)o * 2] = pos[(size_t)d * 2];                   // end q of segment s is row 2s+q of the position array
    pts[(size_t)o * 2 + 1] = pos[(size_t)d * 2 + 1];
    poly_of_point[o] = pl;
  }
  if (p + 1u < n) {                                            // open: p < m; closed: p < m - 1
    const unsigned h = d ^ 1u;
    pts[(size_t)(o + p + 1u) * 2] = pos[(size_t)h * 2];
    pts[(size_t)(o + p + 1u) * 2 + 1] = pos[(size_t)h * 2 + 1];
    poly_of_point[o + p + 1u] = pl;
  }
}

template <typename G>
__device__ __forceinline__ bool close2(const G* a, const G* b) {        // np.allclose(a, b)
  return fabs((double)a[0] - (double)b[0]) <= 1e-8 + 1e-5 * fabs((double)b[0]) &&
         fabs((double)a[1] - (double)b[1]) <= 1e-8 + 1e-5 * fabs((double)b[1]);
}

// keep[q] = 0 for a point np.allclose to the one before it in the same contour (triangulated.py:269)
template <typename G>
__global__ void k_p_dedupe_flag(unsigned np_, const G* __restrict__ pts, const unsigned* __restrict__ poly_of_point,
                                const unsigned* __restrict__ p_off, unsigned* __restrict__ keep) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np_) return;
  const bool first = p_off[poly_of_point[q]] == q;
  keep[q] = (first || !close2(pts + (size_t)q * 2, pts + (size_t)(q - 1) * 2)) ? 1u : 0u;
}

template <typename G>
__global__ void k_p_dedupe_move(unsigned np_, const G* __restrict__ pts, const unsigned* __restrict__ keep, const unsigned* __restrict__ idx,
                                G* __restrict__ out) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np_ || !keep[q]) return;
  out[(size_t)idx[q] * 2] = pts[(size_t)q * 2];
  out[(size_t)idx[q] * 2 + 1] = pts[(size_t)q * 2 + 1];
}

// new offsets of the contours; a contour whose (filtered) ends are np.allclose is closed
template <typename G>
__global__ void k_p_dedupe_polys(unsigned npoly, unsigned np_, unsigned np_final, const unsigned* __restrict__ idx,
                                 const unsigned* __restrict__ old_off, const unsigned* __restrict__ sorted_old_off_next,
                                 const G* __restrict__ out, unsigned* __restrict__ new_off, unsigned* __restrict__ new_len,
                                 uint8_t* __restrict__ p_closed) {
  const unsigned pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npoly) return;
  const unsigned o = old_off[pl], e = sorted_old_off_next[pl];                  // old range [o, e)
  const unsigned no = idx[o], ne = e < np_ ? idx[e] : np_final;
  new_off[pl] = no;
  new_len[pl] = ne - no;
  if (ne - no > 1u && close2(out + (size_t)no * 2, out + (size_t)(ne - 1u) * 2)) p_closed[pl] |= 2;   // bit 0 stays structural
}

__global__ void k_p_old_ends(unsigned npoly, const unsigned* __restrict__ p_off, const unsigned* __restrict__ npts_of_poly,
                             unsigned* __restrict__ end_of) {
  const unsigned pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npoly) return;
  end_of[pl] = p_off[pl] + npts_of_poly[pl];
}

__global__ void k_p_len(unsigned nd, const unsigned* __restrict__ npts, const unsigned* __restrict__ pidx, unsigned* __restrict__ len_of_poly) {
  const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd || !npts[d]) return;
  len_of_poly[pidx[d]] = npts[d];
}

int scan_u32(ctr_ctx* ctx, const unsigned* in, unsigned n, unsigned* out, unsigned* nonzero_out, unsigned* block_sum, unsigned* total) {
  cudaStream_t st = ctx->stream;
  const unsigned nb = (n + PS_TILE - 1) / PS_TILE;
  k_ps_local<<<nb, PS_THREADS, 0, st>>>(in, n, out, block_sum, nonzero_out);
  k_ps_blocks<<<1, 1024, 0, st>>>(block_sum, nb, total);
  k_ps_add<<<(n + 255) / 256, 256, 0, st>>>(out, nonzero_out, n, block_sum);
  ctx->launches += 3;
  return 0;
}

template <typename G>
int polylines_typed(ctr_ctx* ctx, ctr_poly_counts* out) {
  cudaStream_t st = ctx->stream;
  const size_t S = (size_t)ctx->last_counts[0];
  memset(out, 0, sizeof *out);
  ctx->poly_counts[0] = ctx->poly_counts[1] = 0;
  if (S == 0) return 0;
  if (S >= (1ull << 30)) return ctr_fail(ctx, CTR_ERR_OVERFLOW, "more than 2^30 segments: chain row bands separately");
  const unsigned nd = (unsigned)(2 * S);
  const uint8_t* level = (const uint8_t*)ctx->aux[8].p;
  const unsigned long long* keys = (const unsigned long long*)ctx->aux[9].p;
  const G* pos = (const G*)ctx->aux[10].p;
  size_t nslots = 1024;
  while (nslots < (size_t)nd * 2) nslots <<= 1;
  const unsigned nb_scan = (nd + PS_TILE - 1) / PS_TILE;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_tab = carve(nslots * sizeof(VSlot)), o_succ = carve((size_t)nd * 4), o_bp = carve((size_t)nd * 8),
               o_bp2 = carve((size_t)nd * 8), o_bd = carve((size_t)nd * 4), o_bd2 = carve((size_t)nd * 4), o_j = carve((size_t)nd * 4),
               o_j2 = carve((size_t)nd * 4), o_d = carve((size_t)nd * 4), o_d2 = carve((size_t)nd * 4), o_np = carve((size_t)nd * 4),
               o_off = carve((size_t)nd * 4), o_pi = carve((size_t)nd * 4), o_bs = carve((size_t)nb_scan * 8 + 64), o_ctr = carve(256),
               o_tot = carve(64);
  DevBuf& b_s = ctx->aux[36];
  int rc;
  if ((rc = ctr_ensure(ctx, b_s, off))) return rc;
  char* base = (char*)b_s.p;
  VSlot* tab = (VSlot*)(base + o_tab);
  unsigned* succ = (unsigned*)(base + o_succ);
  unsigned long long *bp = (unsigned long long*)(base + o_bp), *bp2 = (unsigned long long*)(base + o_bp2);
  unsigned *bd = (unsigned*)(base + o_bd), *bd2 = (unsigned*)(base + o_bd2), *jmp = (unsigned*)(base + o_j), *jmp2 = (unsigned*)(base + o_j2);
  unsigned *dist = (unsigned*)(base + o_d), *dist2 = (unsigned*)(base + o_d2);
  unsigned *npts = (unsigned*)(base + o_np), *poff = (unsigned*)(base + o_off), *pidx = (unsigned*)(base + o_pi);
  unsigned* bsum = (unsigned*)(base + o_bs);
  PolyCounters* dctr = (PolyCounters*)(base + o_ctr);
  unsigned* dtot = (unsigned*)(base + o_tot);
  const unsigned blocks = (nd + 255) / 256;
  PolyCounters h;
  k_p_init<<<ctx->sm_count * 8, 256, 0, st>>>(tab, nslots, dctr);
  k_p_insert<<<blocks, 256, 0, st>>>(level, keys, nd, tab, nslots - 1, dctr);
  k_p_succ<<<blocks, 256, 0, st>>>(level, keys, nd, tab, nslots - 1, succ, bp, bd, jmp);
  ctx->launches += 3;
  // first dart of every cycle
  for (int round = 0; round < 40; round += 2) {
    CTR_CUDA(ctx, cudaMemsetAsync(&dctr->changed, 0, 4, st));
    k_p_min<<<blocks, 256, 0, st>>>(nd, bp, bd, jmp, bp2, bd2, jmp2, dctr);
    k_p_min<<<blocks, 256, 0, st>>>(nd, bp2, bd2, jmp2, bp, bd, jmp, dctr);
    ctx->launches += 2;
    CTR_CUDA(ctx, cudaMemcpyAsync(&h, dctr, sizeof h, cudaMemcpyDeviceToHost, st));
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
    if (h.junction_levels) {
      out->junction_levels = h.junction_levels;                 // the host's fixed-order walk decides such levels
      return 0;
    }
    if (!h.changed) break;
  }
  unsigned* start = bd;                                         // first dart of each dart's cycle
  // positions: cut in front of the first dart, rank
  unsigned *nxt = jmp, *nxt2 = jmp2;
  k_p_cut<<<blocks, 256, 0, st>>>(nd, succ, start, nxt, dist);
  ctx->launches++;
  for (int round = 0; round < 40; round += 2) {
    CTR_CUDA(ctx, cudaMemsetAsync(&dctr->live, 0, 4, st));
    k_p_rank<<<blocks, 256, 0, st>>>(nd, nxt, dist, nxt2, dist2, dctr);
    k_p_rank<<<blocks, 256, 0, st>>>(nd, nxt2, dist2, nxt, dist, dctr);
    ctx->launches += 2;
    CTR_CUDA(ctx, cudaMemcpyAsync(&h, dctr, sizeof h, cudaMemcpyDeviceToHost, st));
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
    if (!h.live) break;
  }
  // kept contours, their sizes and places
  k_p_mark<<<blocks, 256, 0, st>>>(nd, keys, start, dist, npts);
  ctx->launches++;
  scan_u32(ctx, npts, nd, poff, pidx, bsum, dtot);
  unsigned tot[2];
  CTR_CUDA(ctx, cudaMemcpyAsync(tot, dtot, 8, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  const unsigned np_ = tot[0], npoly = tot[1];
  if (!npoly) return 0;
  DevBuf &b_pts = ctx->aux[37], &b_pts2 = ctx->aux[38], &b_desc = ctx->aux[39], &b_tmp = ctx->aux[40];
  if ((rc = ctr_ensure(ctx, b_pts, (size_t)np_ * 2 * sizeof(G) + 16))) return rc;
  if ((rc = ctr_ensure(ctx, b_pts2, (size_t)np_ * 2 * sizeof(G) + 16))) return rc;
  // descriptors: level i32, start key u64, offset u32, length u32, closed u8
  const size_t d_key = 0, d_lvl = (size_t)npoly * 8, d_off = d_lvl + (size_t)npoly * 4, d_len = d_off + (size_t)npoly * 4,
               d_cl = d_len + (size_t)npoly * 4;
  if ((rc = ctr_ensure(ctx, b_desc, d_cl + npoly + 64))) return rc;
  // scratch per point / per contour
  const unsigned nb2 = (np_ + PS_TILE - 1) / PS_TILE;
  size_t t_off = 0;
  auto carve2 = [&](size_t bytes) {
    const size_t o = t_off;
    t_off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t t_pop = carve2((size_t)np_ * 4), t_keep = carve2((size_t)np_ * 4), t_idx = carve2((size_t)np_ * 4),
               t_bs = carve2((size_t)nb2 * 8 + 64), t_ooff = carve2((size_t)npoly * 4), t_oend = carve2((size_t)npoly * 4),
               t_olen = carve2((size_t)npoly * 4);
  if ((rc = ctr_ensure(ctx, b_tmp, t_off))) return rc;
  char* tb = (char*)b_tmp.p;
  unsigned *pop = (unsigned*)(tb + t_pop), *keep = (unsigned*)(tb + t_keep), *idx = (unsigned*)(tb + t_idx), *bs2 = (unsigned*)(tb + t_bs),
           *ooff = (unsigned*)(tb + t_ooff), *oend = (unsigned*)(tb + t_oend), *olen = (unsigned*)(tb + t_olen);
  char* db = (char*)b_desc.p;
  unsigned long long* p_key = (unsigned long long*)(db + d_key);
  int* p_level = (int*)(db + d_lvl);
  unsigned* p_off = (unsigned*)(db + d_off);
  unsigned* p_len = (unsigned*)(db + d_len);
  uint8_t* p_closed = (uint8_t*)(db + d_cl);
  k_p_points<G><<<blocks, 256, 0, st>>>(nd, level, keys, pos, start, dist, npts, poff, pidx, (G*)b_pts.p, pop, p_level, p_closed, p_key, ooff);
  k_p_len<<<blocks, 256, 0, st>>>(nd, npts, pidx, olen);
  const unsigned pb = (npoly + 255) / 256, qb = (np_ + 255) / 256;
  k_p_old_ends<<<pb, 256, 0, st>>>(npoly, ooff, olen, oend);
  k_p_dedupe_flag<G><<<qb, 256, 0, st>>>(np_, (const G*)b_pts.p, pop, ooff, keep);
  ctx->launches += 4;
  scan_u32(ctx, keep, np_, idx, nullptr, bs2, dtot);
  CTR_CUDA(ctx, cudaMemcpyAsync(tot, dtot, 8, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  const unsigned np_final = tot[0];
  k_p_dedupe_move<G><<<qb, 256, 0, st>>>(np_, (const G*)b_pts.p, keep, idx, (G*)b_pts2.p);
  k_p_dedupe_polys<G><<<pb, 256, 0, st>>>(npoly, np_, np_final, idx, ooff, oend, (const G*)b_pts2.p, p_off, p_len, p_closed);
  ctx->launches += 2;
  CTR_CUDA(ctx, cudaGetLastError());
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->poly_counts[0] = npoly;
  ctx->poly_counts[1] = np_final;
  out->n_polylines = npoly;
  out->n_points = np_final;
  return 0;
}

}  // namespace

extern "C" int ctr_mt2d_polylines(ctr_ctx* ctx, ctr_poly_counts* out) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (!out) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "null argument");
  if (ctx->last_kind != 2 || (ctx->last_flags & CTR_NO_GEOMETRY))
    return ctr_fail(ctx, CTR_ERR_STATE, "no completed ctr_mt2d_run with geometry to chain");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->last_flags & CTR_GEOM_F64) return polylines_typed<double>(ctx, out);
  return polylines_typed<float>(ctx, out);
}

extern "C" int ctr_mt2d_polylines_fetch(ctr_ctx* ctx, int32_t* level, uint8_t* closed, uint64_t* start_key, uint32_t* offset,
                                        uint32_t* length, void* points) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (ctx->last_kind != 2) return ctr_fail(ctx, CTR_ERR_STATE, "no completed ctr_mt2d_run");
  const size_t npoly = (size_t)ctx->poly_counts[0], np_ = (size_t)ctx->poly_counts[1];
  if (!npoly) return 0;
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t gsz = (ctx->last_flags & CTR_GEOM_F64) ? 8 : 4;
  const char* db = (const char*)ctx->aux[39].p;
  const size_t d_lvl = npoly * 8, d_off = d_lvl + npoly * 4, d_len = d_off + npoly * 4, d_cl = d_len + npoly * 4;
  if (start_key) CTR_CUDA(ctx, cudaMemcpyAsync(start_key, db, npoly * 8, cudaMemcpyDeviceToHost, st));
  if (level) CTR_CUDA(ctx, cudaMemcpyAsync(level, db + d_lvl, npoly * 4, cudaMemcpyDeviceToHost, st));
  if (offset) CTR_CUDA(ctx, cudaMemcpyAsync(offset, db + d_off, npoly * 4, cudaMemcpyDeviceToHost, st));
  if (length) CTR_CUDA(ctx, cudaMemcpyAsync(length, db + d_len, npoly * 4, cudaMemcpyDeviceToHost, st));
  if (closed) CTR_CUDA(ctx, cudaMemcpyAsync(closed, db + d_cl, npoly, cudaMemcpyDeviceToHost, st));
  if (points && np_) CTR_CUDA(ctx, cudaMemcpyAsync(points, ctx->aux[38].p, np_ * 2 * gsz, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}
