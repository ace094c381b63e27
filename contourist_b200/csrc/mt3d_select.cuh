// Seeded extraction -- included by mt3d.cu inside its anonymous namespace.
//
// The reference tracks the surface from seed segments: find_initial_voxels (tetrahedral.py:396-441, restated on the
// host side of the binding: it reads a handful of samples) and a flood fill over the 26-neighbourhood of "border"
// voxels (expand_voxels / in_range, tetrahedral.py:443-469).  Its result is therefore the full scan restricted to the
// 26-connected components of border voxels that contain a seed voxel.  That is what these kernels compute on the mesh
// of the last full-volume ctr_mt3d_run:
//   nodes      = the emitting voxels of the run's work list  +  "connectors": border voxels that emit nothing (a corner
//                with f == value exactly on an otherwise high voxel, or every crossing tet skipped by np.allclose);
//                the flood fill walks through those, so they have to be in the graph;
//   k_sel_insert / k_sel_union : voxel id -> node in an open-addressing table; every node looks up its 13 "forward"
//                neighbours and unites with the ones that exist (lock-free union-find, uf_hash.cuh);
//   k_sel_seeds : the components of the seed voxels get flagged;
//   k_sel_mark  : every emitting voxel copies its component's flag to its triangles (contiguous in the output);
//   k_sel_used, k_flag_scan x 2, k_sel_rows / k_sel_tris : the kept triangles and the vertices they use are compacted
//                in place order (single-pass look-back scans), ids rewritten.
constexpr int SC_THREADS = 256;
constexpr int SC_PER = 8;
constexpr int SC_TILE = SC_THREADS * SC_PER;

struct SelCounters {
  unsigned n_extra;                  // connector voxels found
  unsigned ticket[2];                // tile tickets of the two scans
  unsigned pad;
  unsigned long long total[2];       // kept triangles, used vertices
};

template <typename T>
__device__ __forceinline__ unsigned long long vox_key(const Grid<T>& g, int i, int j, int k) {
  return ((unsigned long long)i * (unsigned long long)g.n1 + (unsigned long long)j) * (unsigned long long)g.n2 + (unsigned long long)k;
}

template <typename T>
__global__ void k_sel_cells(Grid<T> g, const unsigned long long* __restrict__ cell_id, unsigned n_cell,
                            unsigned long long* __restrict__ node_vox) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cell) return;
  const unsigned long long id = cell_id[c];
  int i, j, w;
  g.word_coords((unsigned)(id >> 28), i, j, w);
  node_vox[c] = vox_key(g, i, j, w * 32 + (int)((id >> 14) & 31u));
}

// node_vox == nullptr: count only
template <typename T>
__global__ void k_sel_connectors(Grid<T> g, unsigned nwords, unsigned long long* __restrict__ node_vox, unsigned base,
                                 unsigned cap, unsigned* n_extra) {
  const unsigned gw = blockIdx.x * blockDim.x + threadIdx.x;
  if (gw >= nwords) return;
  int i, j, w;
  g.word_coords(gw, i, j, w);
  if (i >= g.n0 - 1 || j >= g.n1 - 1) return;
  if (!g.rowflag[(size_t)i * g.n1 + j]) return;      // no sample of the voxel row anywhere near the isovalue
  Planes npl;
  load_planes(g, g.nbits, i, j, w, npl);
  uint32_t cand = (npl.P[0] | npl.P[1] | npl.P[2] | npl.P[3] | npl.S[0] | npl.S[1] | npl.S[2] | npl.S[3]) & npl.kp1;
  while (cand) {
    const int b = __ffs(cand) - 1;
    cand &= cand - 1;
    const int k = w * 32 + b;
    if (cell_emit_exact(g, i, j, k, nullptr)) continue;          // an emitting voxel: already a node
    double mn = INFINITY, mx = -INFINITY;
    bool all_a = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double f = sample(g, i + ((c >> 2) & 1), j + ((c >> 1) & 1), k + (c & 1));
      mn = fmin(mn, f);
      mx = fmax(mx, f);
      all_a = all_a && near_a(f, g.v);
    }
    if (mn <= g.v && g.v <= mx && !all_a) {                       // border_voxel, tetrahedral.py:383-394
      const unsigned slot = atomicAdd(n_extra, 1u);
      if (node_vox && base + slot < cap) node_vox[base + slot] = vox_key(g, i, j, k);
    }
  }
}

__global__ void k_sel_init(ufh::Slot* tab, size_t nslots, int* parent, uint8_t* flag, unsigned n_nodes) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)gridDim.x * blockDim.x;
  for (size_t q = t; q < nslots; q += n) reinterpret_cast<int4*>(tab)[q] = make_int4(-1, -1, -1, 0);
  for (size_t q = t; q < n_nodes; q += n) {
    parent[q] = (int)q;
    flag[q] = 0;
  }
}

__global__ void k_sel_insert(const unsigned long long* __restrict__ node_vox, unsigned n_nodes, ufh::Slot* tab, size_t mask) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_nodes) return;
  ufh::hash_slot(tab, mask, node_vox[c], true)->tri = (int)c;    // voxel ids are unique: one writer per slot
}

__global__ void k_sel_union(const unsigned long long* __restrict__ node_vox, unsigned n_nodes, ufh::Slot* tab, size_t mask,
                            int* parent, long long n1, long long n2) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_nodes) return;
  const long long key = (long long)node_vox[c];
  // the 13 offsets that follow (0,0,0) lexicographically: each adjacent pair is visited once.  A step that leaves the
  // voxel range lands on sample index n1-1 or n2-1 of some row, which is never a voxel, so it is simply not found.
  for (int o = 14; o < 27; ++o) {
    const int di = o / 9 - 1, dj = (o / 3) % 3 - 1, dk = o % 3 - 1;
    const long long nk = key + ((long long)di * n1 + dj) * n2 + dk;
    if (nk < 0) continue;
    const ufh::Slot* s = ufh::hash_slot(tab, mask, (unsigned long long)nk, false);
    if (s->key == (unsigned long long)nk) ufh::uf_union(parent, (int)c, s->tri);
  }
}

__global__ void k_sel_seeds(const int* __restrict__ seeds, unsigned n_seeds, int n0, int n1, int n2, ufh::Slot* tab,
                            size_t mask, int* parent, uint8_t* flag) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_seeds) return;
  const int i = seeds[q * 3], j = seeds[q * 3 + 1], k = seeds[q * 3 + 2];
  if (i < 0 || j < 0 || k < 0 || i >= n0 - 1 || j >= n1 - 1 || k >= n2 - 1) return;   // a "leak" voxel of the reference
  const unsigned long long key = ((unsigned long long)i * (unsigned long long)n1 + (unsigned long long)j) * (unsigned long long)n2 + (unsigned long long)k;
  const ufh::Slot* s = ufh::hash_slot(tab, mask, key, false);
  if (s->key == key) flag[ufh::uf_find(parent, s->tri)] = 1;
}

__global__ void k_sel_mark(const unsigned long long* __restrict__ cell_id, unsigned n_cell, const uint4* __restrict__ wrec, int* parent, const uint8_t* __restrict__ flag,
                           uint8_t* __restrict__ keep_t, unsigned n_tris, unsigned* n_sel_cells) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cell) return;
  int r = (int)c;
  for (int q = parent[r]; q != r; q = parent[r]) r = q;          // unions are over: read-only walk
  if (!flag[r]) return;
  atomicAdd(n_sel_cells, 1u);
  const unsigned long long id = cell_id[c];
  const unsigned c8 = (unsigned)(id & 255u), emit = (unsigned)((id >> 8) & 63u);
  unsigned nt = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t)
    if ((emit >> t) & 1u) nt += (__popc(tet_mask_of(c8, t)) == 2) ? 2u : 1u;
  const unsigned t0 = wrec[(unsigned)(id >> 28)].y + ((unsigned)(id >> 19) & 511u);
  for (unsigned q = 0; q < nt; ++q)
    if (t0 + q < n_tris) keep_t[t0 + q] = 1;
}

__global__ void k_sel_used(const int* __restrict__ tris, unsigned n_tris, const uint8_t* __restrict__ keep_t, unsigned id_base,
                           uint8_t* __restrict__ used_v, unsigned n_verts) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tris || !keep_t[t]) return;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const unsigned v = (unsigned)tris[(size_t)t * 3 + q] - id_base;
    if (v < n_verts) used_v[v] = 1;
  }
}

// idx[q] = number of set flags before q; *total = number of set flags.  Tiles take tickets, prefixes by look-back.
__global__ void __launch_bounds__(SC_THREADS) k_flag_scan(const uint8_t* __restrict__ flags, unsigned n, uint32_t* __restrict__ idx,
                                                          unsigned long long* status, unsigned* ticket,
                                                          unsigned long long* total, int ntiles) {
  __shared__ unsigned s_tile;
  __shared__ unsigned s_warp[SC_THREADS / 32];
  __shared__ unsigned long long s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = (int)s_tile;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned q0 = (unsigned)tile * SC_TILE + threadIdx.x * SC_PER;
  unsigned f[SC_PER], cnt = 0;
#pragma unroll
  for (int u = 0; u < SC_PER; ++u) {
    f[u] = (q0 + u < n && flags[q0 + u]) ? 1u : 0u;
    cnt += f[u];
  }
  const unsigned inc = warp_incl_scan_u32(cnt);
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned woff = 0, blk = 0;
#pragma unroll
  for (int w = 0; w < SC_THREADS / 32; ++w) {
    if (w < (int)warp) woff += s_warp[w];
    blk += s_warp[w];
  }
  if (warp == 0) {
    const unsigned long long e = lb_lookback(status, tile, (unsigned long long)blk);
    if (lane == 0) s_excl = e;
  }
  __syncthreads();
  unsigned run = (unsigned)s_excl + woff + inc - cnt;
#pragma unroll
  for (int u = 0; u < SC_PER; ++u) {
    if (q0 + u < n) idx[q0 + u] = run;
    run += f[u];
  }
  if (tile == ntiles - 1 && threadIdx.x == 0) *total = s_excl + blk;
}

// rows of `wpr` 32-bit words (or single bytes when wpr == 0): dst[idx[r]] = src[r] for flagged rows
__global__ void k_sel_rows(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint8_t* __restrict__ flags,
                           const uint32_t* __restrict__ idx, unsigned n, int wpr) {
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || !flags[r]) return;
  const size_t d = idx[r];
  if (wpr == 0) {
    reinterpret_cast<uint8_t*>(dst)[d] = reinterpret_cast<const uint8_t*>(src)[r];
    return;
  }
  for (int q = 0; q < wpr; ++q) dst[d * wpr + q] = src[(size_t)r * wpr + q];
}

__global__ void k_sel_tris(const int* __restrict__ src, int* __restrict__ dst, const uint8_t* __restrict__ keep_t,
                           const uint32_t* __restrict__ idx_t, const uint32_t* __restrict__ idx_v, unsigned n_tris,
                           unsigned id_base) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tris || !keep_t[t]) return;
  const size_t d = idx_t[t];
#pragma unroll
  for (int q = 0; q < 3; ++q) dst[d * 3 + q] = (int)(idx_v[(unsigned)src[(size_t)t * 3 + q] - id_base] + id_base);
}

template <typename T>
int select_typed(ctr_ctx* ctx, const ctr_mt3d_params* p, const int32_t* seeds, int64_t n_seeds, int64_t* out_verts,
                 int64_t* out_tris, int64_t* out_voxels) {
  const int n0 = (int)p->n0, n1 = (int)p->n1, n2 = (int)p->n2;
  const int W = (n2 + 31) / 32;
  const long long nwords = (long long)n0 * n1 * W;
  cudaStream_t st = ctx->stream;
  Grid<T> g;
  g.f = (p->flags & CTR_FIELD_ON_DEVICE) ? (const T*)p->field : (const T*)ctx->field.p;
  g.n0 = n0; g.n1 = n1; g.n2 = n2; g.W = W;
  g.i_lo = 0; g.i_hi = n0; g.i_hiv = n0;
  g.plane_offset = p->plane_offset;
  g.id_base = (unsigned)p->vert_id_base;
  g.v = p->isovalue;
  g.tolv = 1e-8 + 1e-5 * fabs(p->isovalue);
  g.any_near = 0;
  g.divW.init((unsigned)W);
  g.divN1.init((unsigned)n1);
  g.bits = (const uint32_t*)ctx->bits.p;
  g.nbits = (const uint32_t*)ctx->nbits.p;
  g.rowflag = (const uint8_t*)ctx->aux[4].p;
  g.wordflag = (const uint32_t*)ctx->aux[32].p;
  g.exactflag = (uint32_t*)ctx->aux[35].p;
  g.wshift = 0;
  while ((W >> g.wshift) > 32) ++g.wshift;
  const unsigned nV = (unsigned)ctx->last_counts[0], nT = (unsigned)ctx->last_counts[1];
  const unsigned n_cell = (unsigned)ctx->last_cell;
  const uint32_t fl = p->flags;
  const size_t gsz = (fl & CTR_GEOM_F64) ? 8 : 4;
  const bool want_n = (fl & CTR_WANT_NORMALS) != 0, want_k = (fl & CTR_WANT_KEYS) != 0;
  int rc;
  DevBuf& b_ctr = ctx->aux[31];
  if ((rc = ctr_ensure(ctx, b_ctr, sizeof(SelCounters) + 64))) return rc;
  SelCounters* dctr = (SelCounters*)b_ctr.p;
  SelCounters h;
  CTR_CUDA(ctx, cudaMemsetAsync(dctr, 0, sizeof(SelCounters), st));
  // connectors, counted first (rare: they need a sample exactly on the isovalue)
  const unsigned wb = (unsigned)((nwords + 255) / 256);
  k_sel_connectors<T><<<wb, 256, 0, st>>>(g, (unsigned)nwords, nullptr, 0u, 0u, &dctr->n_extra);
  ctx->launches++;
  CTR_CUDA(ctx, cudaMemcpyAsync(&h, dctr, sizeof h, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  const unsigned n_extra = h.n_extra;
  const size_t n_nodes = (size_t)n_cell + n_extra;
  if (n_nodes >= 0x7fffffffull) return ctr_fail(ctx, CTR_ERR_OVERFLOW, "too many border voxels for one selection");
  size_t nslots = 1024;
  while (nslots < n_nodes * 2) nslots <<= 1;
  const int tiles_t = (int)(((size_t)nT + SC_TILE - 1) / SC_TILE), tiles_v = (int)(((size_t)nV + SC_TILE - 1) / SC_TILE);
  // one scratch allocation, carved (256-byte aligned pieces)
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_vox = carve(n_nodes * 8 + 8), o_tab = carve(nslots * sizeof(ufh::Slot)), o_par = carve(n_nodes * 4 + 4),
               o_flag = carve(n_nodes + 1), o_keep = carve((size_t)nT + 8), o_used = carve((size_t)nV + 8),
               o_it = carve((size_t)nT * 4 + 4), o_iv = carve((size_t)nV * 4 + 4), o_st = carve((size_t)tiles_t * 8 + 8),
               o_sv = carve((size_t)tiles_v * 8 + 8), o_out = carve(std::max<size_t>((size_t)nV * 3 * gsz, (size_t)nT * 12) + 16);
  DevBuf& b_s = ctx->aux[30];
  if ((rc = ctr_ensure(ctx, b_s, off))) return rc;
  char* base = (char*)b_s.p;
  unsigned long long* node_vox = (unsigned long long*)(base + o_vox);
  ufh::Slot* tab = (ufh::Slot*)(base + o_tab);
  int* parent = (int*)(base + o_par);
  uint8_t* flag = (uint8_t*)(base + o_flag);
  uint8_t* keep_t = (uint8_t*)(base + o_keep);
  uint8_t* used_v = (uint8_t*)(base + o_used);
  uint32_t* idx_t = (uint32_t*)(base + o_it);
  uint32_t* idx_v = (uint32_t*)(base + o_iv);
  unsigned long long* st_t = (unsigned long long*)(base + o_st);
  unsigned long long* st_v = (unsigned long long*)(base + o_sv);
  void* tmp = base + o_out;
  DevBuf& b_seeds = ctx->aux[29];
  if ((rc = ctr_ensure(ctx, b_seeds, (size_t)std::max<int64_t>(n_seeds, 1) * 12))) return rc;
  if (n_seeds) CTR_CUDA(ctx, cudaMemcpyAsync(b_seeds.p, seeds, (size_t)n_seeds * 12, cudaMemcpyHostToDevice, st));
  CTR_CUDA(ctx, cudaMemsetAsync(dctr, 0, sizeof(SelCounters), st));
  CTR_CUDA(ctx, cudaMemsetAsync(keep_t, 0, (size_t)nT + 8, st));
  CTR_CUDA(ctx, cudaMemsetAsync(used_v, 0, (size_t)nV + 8, st));
  CTR_CUDA(ctx, cudaMemsetAsync(st_t, 0, (size_t)tiles_t * 8 + 8, st));
  CTR_CUDA(ctx, cudaMemsetAsync(st_v, 0, (size_t)tiles_v * 8 + 8, st));
  const unsigned nb = (unsigned)((n_nodes + 255) / 256), cb = (n_cell + 255) / 256, tb = (nT + 255) / 256, vb = (nV + 255) / 256;
  if (n_cell) k_sel_cells<T><<<cb, 256, 0, st>>>(g, (const unsigned long long*)ctx->aux[2].p, n_cell, node_vox);
  if (n_extra) k_sel_connectors<T><<<wb, 256, 0, st>>>(g, (unsigned)nwords, node_vox, n_cell, (unsigned)n_nodes, &dctr->n_extra);
  k_sel_init<<<ctx->sm_count * 8, 256, 0, st>>>(tab, nslots, parent, flag, (unsigned)n_nodes);
  if (n_nodes) {
    k_sel_insert<<<nb, 256, 0, st>>>(node_vox, (unsigned)n_nodes, tab, nslots - 1);
    k_sel_union<<<nb, 256, 0, st>>>(node_vox, (unsigned)n_nodes, tab, nslots - 1, parent, n1, n2);
  }
  if (n_seeds)
    k_sel_seeds<<<(unsigned)((n_seeds + 255) / 256), 256, 0, st>>>((const int*)b_seeds.p, (unsigned)n_seeds, n0, n1, n2, tab,
                                                                    nslots - 1, parent, flag);
  if (n_cell)
    k_sel_mark<<<cb, 256, 0, st>>>((const unsigned long long*)ctx->aux[2].p, n_cell,
                                   (const uint4*)ctx->wdir.p, parent, flag, keep_t, nT, &dctr->pad);
  if (nT) {
    k_sel_used<<<tb, 256, 0, st>>>((const int*)ctx->tris.p, nT, keep_t, g.id_base, used_v, nV);
    k_flag_scan<<<tiles_t, SC_THREADS, 0, st>>>(keep_t, nT, idx_t, st_t, &dctr->ticket[0], &dctr->total[0], tiles_t);
  }
  if (nV) k_flag_scan<<<tiles_v, SC_THREADS, 0, st>>>(used_v, nV, idx_v, st_v, &dctr->ticket[1], &dctr->total[1], tiles_v);
  ctx->launches += 9;
  CTR_CUDA(ctx, cudaGetLastError());
  CTR_CUDA(ctx, cudaMemcpyAsync(&h, dctr, sizeof h, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  const size_t newT = (size_t)h.total[0], newV = (size_t)h.total[1];
  // compaction through the scratch piece, array by array
  auto rows = [&](DevBuf& b, int wpr, size_t row_bytes) -> int {
    if (!nV || !b.p) return 0;
    k_sel_rows<<<vb, 256, 0, st>>>((const uint32_t*)b.p, (uint32_t*)tmp, used_v, idx_v, nV, wpr);
    ctx->launches++;
    if (newV) CTR_CUDA(ctx, cudaMemcpyAsync(b.p, tmp, newV * row_bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  if ((rc = rows(ctx->verts, (int)(3 * gsz / 4), 3 * gsz))) return rc;
  if (want_n && (rc = rows(ctx->normals, (int)(3 * gsz / 4), 3 * gsz))) return rc;
  if (want_k && (rc = rows(ctx->keys, 2, 8))) return rc;
  if (want_k && (rc = rows(ctx->lowmin, 0, 1))) return rc;
  if (nT) {
    k_sel_tris<<<tb, 256, 0, st>>>((const int*)ctx->tris.p, (int*)tmp, keep_t, idx_t, idx_v, nT, g.id_base);
    ctx->launches++;
    if (newT) CTR_CUDA(ctx, cudaMemcpyAsync(ctx->tris.p, tmp, newT * 12, cudaMemcpyDeviceToDevice, st));
  }
  CTR_CUDA(ctx, cudaGetLastError());
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->last_counts[0] = (int64_t)newV;
  ctx->last_counts[1] = (int64_t)newT;
  ctx->last3_edited |= 1u;
  if (out_verts) *out_verts = (int64_t)newV;
  if (out_tris) *out_tris = (int64_t)newT;
  if (out_voxels) *out_voxels = (int64_t)h.pad;
  return 0;
}
