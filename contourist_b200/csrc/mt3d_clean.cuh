// The reference's mesh post-processing on the device mesh of the last ctr_mt3d_run --
// included by mt3d.cu inside its anonymous namespace, after mt3d_select.cuh (whose scan / compaction kernels it uses).
//
//   tetrahedral.py:190-215    quantize_interpolations : interpolations that truncate to the same cell of the
//                             `divisions` grid become one vertex; simplices that lose a vertex, or equal another, go;
//   tetrahedral.py:353-375    remove_tiny_simplices   : a simplex whose extent is < epsilon of the grid on every axis
//                             is dropped and its vertices moved to one point;
//   surface_geometry.py:14-50 clean_triangles         : zero-area triangles are dropped, coincident vertices of those
//                             triangles merged, vertices renumbered over what the kept triangles use.
// In the reference all three depend on CPython dict / set iteration order.  Here the order is fixed: wherever the
// reference keeps "whichever came first / last", the vertex with the SMALLEST id survives (ids are the engine's
// deterministic numbering), and the sequential overwrite passes become "decide every simplex from the positions
// before the pass, then merge connected clusters" (lock-free union-find whose roots are minima).  oracle/post3d.py
// states the same rules in numpy and is pinned to final meshes of the unmodified reference.
//
//   k_c_qinsert / k_c_qmap : quantum cell (3 x 21 bits) -> open-addressing table, value = smallest vertex id in the cell;
//   k_c_qtris / k_c_qdedupe: triangles through that map; the distinct ones go into a second table keyed by a 64-bit
//                            hash of the sorted triple, value = smallest triangle index; a triangle that finds another
//                            index there compares the triples (equal: duplicate, dropped; different: a hash collision,
//                            reported to the host, which retries with another seed);
//   k_c_tiny / k_c_tinypos : extent test, union of the three vertices, positions of non-roots overwritten by their root's;
//   k_c_flat / k_c_final   : cross-product test, union of np.allclose vertex pairs of flat triangles, kept triangles
//                            mapped through the merge and checked for lost vertices; used-vertex flags;
//   k_flag_scan x 2, k_sel_rows, k_c_tris_out : order-preserving compaction of vertices, normals, keys and triangles;
//   k_c_transform          : grid -> world (grid_field.py:89-93) at the very end, after the optional orientation pass,
//                            which the reference also runs in grid coordinates (tetrahedral.py:611-617).

struct CleanCounters {
  unsigned n_quant, n_tiny, n_flat, collision;
};

__device__ __forceinline__ double c_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double c_sub(double a, double b) { return __dsub_rn(a, b); }

template <typename G>
__device__ __forceinline__ unsigned long long quantum_key(const G* __restrict__ verts, unsigned v, const double ex[3]) {
  unsigned long long key = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const long long q = (long long)c_mul((double)verts[(size_t)v * 3 + a], ex[a]);     // astype(int): towards zero
    key |= ((unsigned long long)q & 0x1fffffull) << (21 * a);
  }
  return key;
}

struct Expander {
  double ex[3];
  double inv[3];
};

__global__ void k_c_init(ufh::Slot* tab_v, size_t nslots_v, ufh::Slot* tab_t, size_t nslots_t, int* parent_a, int* parent_b,
                         uint8_t* used, unsigned nv) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)gridDim.x * blockDim.x;
  for (size_t q = t; q < nslots_v; q += n) reinterpret_cast<int4*>(tab_v)[q] = make_int4(-1, -1, 0x7fffffff, 0);
  for (size_t q = t; q < nslots_t; q += n) reinterpret_cast<int4*>(tab_t)[q] = make_int4(-1, -1, 0x7fffffff, 0);
  for (size_t q = t; q < nv; q += n) {
    parent_a[q] = (int)q;
    parent_b[q] = (int)q;
    used[q] = 0;
  }
}

template <typename G>
__global__ void k_c_qinsert(const G* __restrict__ verts, unsigned nv, Expander e, ufh::Slot* tab, size_t mask) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  atomicMin(&ufh::hash_slot(tab, mask, quantum_key(verts, v, e.ex), true)->tri, (int)v);
}

template <typename G>
__global__ void k_c_qmap(const G* __restrict__ verts, unsigned nv, Expander e, ufh::Slot* tab, size_t mask, int* __restrict__ rep) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  rep[v] = ufh::hash_slot(tab, mask, quantum_key(verts, v, e.ex), false)->tri;
}

__device__ __forceinline__ void sort3(int& a, int& b, int& c) {
  int t;
  if (a > b) { t = a; a = b; b = t; }
  if (b > c) { t = b; b = c; c = t; }
  if (a > b) { t = a; a = b; b = t; }
}

__device__ __forceinline__ unsigned long long triple_hash(int a, int b, int c, unsigned long long seed) {
  sort3(a, b, c);
  unsigned long long h = ufh::mix64(((unsigned long long)(unsigned)a << 32 | (unsigned)b) + seed);
  h = ufh::mix64(h ^ ((unsigned long long)(unsigned)c * 0x9e3779b97f4a7c15ull));
  return h == ufh::EMPTY ? 0ull : h;
}

// quantize: triangles through the representative map; keep[t] = 1 for triangles that still have three vertices
__global__ void k_c_qtris(const int* __restrict__ tris, unsigned nt, const int* __restrict__ rep, int* __restrict__ tris2,
                          uint8_t* __restrict__ keep, ufh::Slot* tab, size_t mask, unsigned long long seed) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const int a = rep[tris[(size_t)t * 3]], b = rep[tris[(size_t)t * 3 + 1]], c = rep[tris[(size_t)t * 3 + 2]];
  tris2[(size_t)t * 3] = a;
  tris2[(size_t)t * 3 + 1] = b;
  tris2[(size_t)t * 3 + 2] = c;
  const bool distinct = a != b && a != c && b != c;                                    // tetrahedral.py:208
  keep[t] = distinct ? 1 : 0;
  if (distinct) atomicMin(&ufh::hash_slot(tab, mask, triple_hash(a, b, c, seed), true)->tri, (int)t);
}

// simplex_sets is a set of frozensets (tetrahedral.py:209): of equal triples the first in triangle order stays
__global__ void k_c_qdedupe(const int* __restrict__ tris2, unsigned nt, uint8_t* __restrict__ keep, ufh::Slot* tab, size_t mask,
                            unsigned long long seed, CleanCounters* ctr) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  if (!keep[t]) {
    atomicAdd(&ctr->n_quant, 1u);
    return;
  }
  int a = tris2[(size_t)t * 3], b = tris2[(size_t)t * 3 + 1], c = tris2[(size_t)t * 3 + 2];
  const int m = ufh::hash_slot(tab, mask, triple_hash(a, b, c, seed), false)->tri;
  if (m == (int)t) return;
  int ma = tris2[(size_t)m * 3], mb = tris2[(size_t)m * 3 + 1], mc = tris2[(size_t)m * 3 + 2];
  sort3(a, b, c);
  sort3(ma, mb, mc);
  if (a == ma && b == mb && c == mc) {
    keep[t] = 0;
    atomicAdd(&ctr->n_quant, 1u);
  } else {
    atomicAdd(&ctr->collision, 1u);
  }
}

template <typename G>
__global__ void k_c_tiny(const G* __restrict__ verts, const int* __restrict__ tris2, unsigned nt, uint8_t* __restrict__ keep,
                         Expander e, double epsilon, int* parent, CleanCounters* ctr) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || !keep[t]) return;
  const int v[3] = {tris2[(size_t)t * 3], tris2[(size_t)t * 3 + 1], tris2[(size_t)t * 3 + 2]};
  double worst = -INFINITY;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const double p0 = (double)verts[(size_t)v[0] * 3 + ax], p1 = (double)verts[(size_t)v[1] * 3 + ax],
                 p2 = (double)verts[(size_t)v[2] * 3 + ax];
    const double d = c_mul(c_sub(fmax(p0, fmax(p1, p2)), fmin(p0, fmin(p1, p2))), e.inv[ax]);   // tetrahedral.py:363-365
    worst = fmax(worst, d);
  }
  if (worst < epsilon) {                                                                // :366
    keep[t] = 0;
    atomicAdd(&ctr->n_tiny, 1u);
    ufh::uf_union(parent, v[0], v[1]);
    ufh::uf_union(parent, v[0], v[2]);
  }
}

// non-roots take their root's position; roots are never written, so this runs in place
template <typename G>
__global__ void k_c_tinypos(G* __restrict__ verts, unsigned nv, int* parent) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  int r = (int)v;
  for (int q = parent[r]; q != r; q = parent[r]) r = q;
  if (r == (int)v) return;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) verts[(size_t)v * 3 + ax] = verts[(size_t)r * 3 + ax];
}

__device__ __forceinline__ bool close_to(const double p[3], const double q[3]) {       // np.allclose(p, q)
  bool ok = true;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) ok = ok && fabs(c_sub(p[ax], q[ax])) <= __dadd_rn(1e-8, c_mul(1e-5, fabs(q[ax])));
  return ok;
}

template <typename G>
__global__ void k_c_flat(const G* __restrict__ verts, const int* __restrict__ tris2, unsigned nt, uint8_t* __restrict__ keep,
                         int* parent, CleanCounters* ctr) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || !keep[t]) return;
  const int v[3] = {tris2[(size_t)t * 3], tris2[(size_t)t * 3 + 1], tris2[(size_t)t * 3 + 2]};
  double P[3][3];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) P[q][ax] = (double)verts[(size_t)v[q] * 3 + ax];
  double u[3], w[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    u[ax] = c_sub(P[0][ax], P[2][ax]);                                                  // A - C
    w[ax] = c_sub(P[1][ax], P[2][ax]);                                                  // B - C
  }
  const double cx = c_sub(c_mul(u[1], w[2]), c_mul(u[2], w[1]));
  const double cy = c_sub(c_mul(u[2], w[0]), c_mul(u[0], w[2]));
  const double cz = c_sub(c_mul(u[0], w[1]), c_mul(u[1], w[0]));
  if (fabs(cx) <= 1e-8 && fabs(cy) <= 1e-8 && fabs(cz) <= 1e-8) {                       // np.allclose(cross, 0)
    keep[t] = 0;
    atomicAdd(&ctr->n_flat, 1u);
    if (close_to(P[0], P[1])) ufh::uf_union(parent, v[0], v[1]);                        // surface_geometry.py:39-44
    if (close_to(P[0], P[2])) ufh::uf_union(parent, v[0], v[2]);
    if (close_to(P[1], P[2])) ufh::uf_union(parent, v[1], v[2]);
  }
}

__global__ void k_c_final(int* __restrict__ tris2, unsigned nt, uint8_t* __restrict__ keep, int* parent, uint8_t* __restrict__ used,
                          CleanCounters* ctr) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || !keep[t]) return;
  int v[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    int r = tris2[(size_t)t * 3 + q];
    for (int p = parent[r]; p != r; p = parent[r]) r = p;
    v[q] = r;
  }
  if (v[0] == v[1] || v[0] == v[2] || v[1] == v[2]) {
    keep[t] = 0;
    atomicAdd(&ctr->n_flat, 1u);
    return;
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    tris2[(size_t)t * 3 + q] = v[q];
    used[v[q]] = 1;
  }
}

__global__ void k_c_tris_out(const int* __restrict__ tris2, int* __restrict__ dst, const uint8_t* __restrict__ keep,
                             const uint32_t* __restrict__ idx_t, const uint32_t* __restrict__ idx_v, unsigned nt) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || !keep[t]) return;
  const size_t d = idx_t[t];
#pragma unroll
  for (int q = 0; q < 3; ++q) dst[d * 3 + q] = (int)idx_v[tris2[(size_t)t * 3 + q]];
}

template <typename G>
__global__ void k_c_transform(G* __restrict__ verts, G* __restrict__ normals, unsigned nv, Xform xf) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax)
    verts[(size_t)v * 3 + ax] = add_rn(mul_rn(verts[(size_t)v * 3 + ax], (G)xf.delta[ax]), (G)xf.origin[ax]);   // grid_field.py:93
  if (normals) {
    G n[3], len2 = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      n[ax] = normals[(size_t)v * 3 + ax] * (G)xf.inv_delta[ax];
      len2 += n[ax] * n[ax];
    }
    const G inv = len2 > (G)0 ? inv_sqrt(len2) : (G)0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) normals[(size_t)v * 3 + ax] = n[ax] * inv;
  }
}

template <typename G>
int clean_typed(ctr_ctx* ctx, const ctr_clean_params* cp, ctr_clean_counts* out) {
  cudaStream_t st = ctx->stream;
  const unsigned nV = (unsigned)ctx->last_counts[0], nT = (unsigned)ctx->last_counts[1];
  const uint32_t fl = ctx->last_flags;
  const size_t gsz = sizeof(G);
  const bool want_n = (fl & CTR_WANT_NORMALS) != 0, want_k = (fl & CTR_WANT_KEYS) != 0;
  Expander e;
  for (int a = 0; a < 3; ++a) {
    e.ex[a] = (double)(long long)(((double)cp->divisions * 1.0) / (double)cp->corner[a]);   // tetrahedral.py:192
    e.inv[a] = 1.0 / (double)cp->corner[a];                                                  // :360
  }
  Xform xf;
  bool identity = true;
  for (int a = 0; a < 3; ++a) {
    xf.origin[a] = cp->origin[a];
    xf.delta[a] = cp->delta[a];
    xf.inv_delta[a] = 1.0 / cp->delta[a];
    identity = identity && cp->origin[a] == 0.0 && cp->delta[a] == 1.0;
  }
  int rc;
  size_t nslots_v = 1024, nslots_t = 1024;
  while (nslots_v < (size_t)nV * 2) nslots_v <<= 1;
  while (nslots_t < (size_t)nT * 2) nslots_t <<= 1;
  const int tiles_t = (int)(((size_t)nT + SC_TILE - 1) / SC_TILE), tiles_v = (int)(((size_t)nV + SC_TILE - 1) / SC_TILE);
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_tv = carve(nslots_v * sizeof(ufh::Slot)), o_tt = carve(nslots_t * sizeof(ufh::Slot)),
               o_rep = carve((size_t)nV * 4 + 4), o_pa = carve((size_t)nV * 4 + 4), o_pb = carve((size_t)nV * 4 + 4),
               o_used = carve((size_t)nV + 8), o_keep = carve((size_t)nT + 8), o_t2 = carve((size_t)nT * 12 + 16),
               o_it = carve((size_t)nT * 4 + 4), o_iv = carve((size_t)nV * 4 + 4), o_st = carve((size_t)tiles_t * 8 + 8),
               o_sv = carve((size_t)tiles_v * 8 + 8), o_sel = carve(sizeof(SelCounters) + 64), o_cc = carve(sizeof(CleanCounters) + 64),
               o_out = carve(std::max<size_t>((size_t)nV * 3 * gsz, (size_t)nT * 12) + 16);
  DevBuf& b_s = ctx->aux[28];
  if ((rc = ctr_ensure(ctx, b_s, off))) return rc;
  char* base = (char*)b_s.p;
  ufh::Slot* tab_v = (ufh::Slot*)(base + o_tv);
  ufh::Slot* tab_t = (ufh::Slot*)(base + o_tt);
  int* rep = (int*)(base + o_rep);
  int* parent_a = (int*)(base + o_pa);
  int* parent_b = (int*)(base + o_pb);
  uint8_t* used_v = (uint8_t*)(base + o_used);
  uint8_t* keep_t = (uint8_t*)(base + o_keep);
  int* tris2 = (int*)(base + o_t2);
  uint32_t* idx_t = (uint32_t*)(base + o_it);
  uint32_t* idx_v = (uint32_t*)(base + o_iv);
  unsigned long long* st_t = (unsigned long long*)(base + o_st);
  unsigned long long* st_v = (unsigned long long*)(base + o_sv);
  SelCounters* dsel = (SelCounters*)(base + o_sel);
  CleanCounters* dcc = (CleanCounters*)(base + o_cc);
  void* tmp = base + o_out;
  G* verts = (G*)ctx->verts.p;
  const int* tris = (const int*)ctx->tris.p;
  const unsigned vb = (nV + 255) / 256, tb = (nT + 255) / 256;
  CleanCounters hc;
  memset(&hc, 0, sizeof hc);
  SelCounters hs;
  memset(&hs, 0, sizeof hs);
  if (nV && nT) {
    for (int attempt = 0;; ++attempt) {
      const unsigned long long seed = 0x51ed270b1a2c3d4full * (unsigned long long)(attempt + 1);
      CTR_CUDA(ctx, cudaMemsetAsync(dcc, 0, sizeof(CleanCounters), st));
      k_c_init<<<ctx->sm_count * 8, 256, 0, st>>>(tab_v, nslots_v, tab_t, nslots_t, parent_a, parent_b, used_v, nV);
      k_c_qinsert<G><<<vb, 256, 0, st>>>(verts, nV, e, tab_v, nslots_v - 1);
      k_c_qmap<G><<<vb, 256, 0, st>>>(verts, nV, e, tab_v, nslots_v - 1, rep);
      k_c_qtris<<<tb, 256, 0, st>>>(tris, nT, rep, tris2, keep_t, tab_t, nslots_t - 1, seed);
      k_c_qdedupe<<<tb, 256, 0, st>>>(tris2, nT, keep_t, tab_t, nslots_t - 1, seed, dcc);
      ctx->launches += 5;
      CTR_CUDA(ctx, cudaMemcpyAsync(&hc, dcc, sizeof hc, cudaMemcpyDeviceToHost, st));
      CTR_CUDA(ctx, cudaStreamSynchronize(st));
      if (!hc.collision) break;
      if (attempt == 3) return ctr_fail(ctx, CTR_ERR_STATE, "triangle de-duplication: hash collisions under four seeds");
    }
    k_c_tiny<G><<<tb, 256, 0, st>>>(verts, tris2, nT, keep_t, e, cp->epsilon, parent_a, dcc);
    k_c_tinypos<G><<<vb, 256, 0, st>>>(verts, nV, parent_a);
    if (!(cp->flags & CTR_CLEAN_NO_TRIANGLES)) k_c_flat<G><<<tb, 256, 0, st>>>(verts, tris2, nT, keep_t, parent_b, dcc);
    k_c_final<<<tb, 256, 0, st>>>(tris2, nT, keep_t, parent_b, used_v, dcc);
    CTR_CUDA(ctx, cudaMemsetAsync(dsel, 0, sizeof(SelCounters), st));
    CTR_CUDA(ctx, cudaMemsetAsync(st_t, 0, (size_t)tiles_t * 8 + 8, st));
    CTR_CUDA(ctx, cudaMemsetAsync(st_v, 0, (size_t)tiles_v * 8 + 8, st));
    k_flag_scan<<<tiles_t, SC_THREADS, 0, st>>>(keep_t, nT, idx_t, st_t, &dsel->ticket[0], &dsel->total[0], tiles_t);
    k_flag_scan<<<tiles_v, SC_THREADS, 0, st>>>(used_v, nV, idx_v, st_v, &dsel->ticket[1], &dsel->total[1], tiles_v);
    ctx->launches += 6;
    CTR_CUDA(ctx, cudaGetLastError());
    CTR_CUDA(ctx, cudaMemcpyAsync(&hc, dcc, sizeof hc, cudaMemcpyDeviceToHost, st));
    CTR_CUDA(ctx, cudaMemcpyAsync(&hs, dsel, sizeof hs, cudaMemcpyDeviceToHost, st));
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
  }
  const size_t newT = (size_t)hs.total[0], newV = (size_t)hs.total[1];
  auto rows = [&](DevBuf& b, int wpr, size_t row_bytes) -> int {
    if (!nV || !b.p) return 0;
    k_sel_rows<<<vb, 256, 0, st>>>((const uint32_t*)b.p, (uint32_t*)tmp, used_v, idx_v, nV, wpr);
    ctx->launches++;
    if (newV) CTR_CUDA(ctx, cudaMemcpyAsync(b.p, tmp, newV * row_bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  if (nV && nT) {
    if ((rc = rows(ctx->verts, (int)(3 * gsz / 4), 3 * gsz))) return rc;
    if (want_n && (rc = rows(ctx->normals, (int)(3 * gsz / 4), 3 * gsz))) return rc;
    if (want_k && (rc = rows(ctx->keys, 2, 8))) return rc;
    if (want_k && (rc = rows(ctx->lowmin, 0, 1))) return rc;
    k_c_tris_out<<<tb, 256, 0, st>>>(tris2, (int*)tmp, keep_t, idx_t, idx_v, nT);
    ctx->launches++;
    if (newT) CTR_CUDA(ctx, cudaMemcpyAsync(ctx->tris.p, tmp, newT * 12, cudaMemcpyDeviceToDevice, st));
    CTR_CUDA(ctx, cudaGetLastError());
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
  }
  ctx->last_counts[0] = (int64_t)newV;
  ctx->last_counts[1] = (int64_t)newT;
  ctx->last3_edited |= 2u;
  out->n_verts = (int64_t)newV;
  out->n_tris = (int64_t)newT;
  out->n_quantized = hc.n_quant;
  out->n_tiny = hc.n_tiny;
  out->n_flat = hc.n_flat;
  out->n_components = 0;
  out->n_flipped = 0;
  if ((cp->flags & CTR_CLEAN_ORIENT) && newT) {
    // surface_geometry.py:52-140, in grid coordinates like the reference (tetrahedral.py:611-617)
    if ((rc = ctr_mt3d_orient_reference(ctx, &out->n_components, &out->n_flipped))) return rc;
  }
  if (!identity && newV) {
    k_c_transform<G><<<(unsigned)((newV + 255) / 256), 256, 0, st>>>(verts, want_n ? (G*)ctx->normals.p : nullptr, (unsigned)newV, xf);
    ctx->launches++;
    CTR_CUDA(ctx, cudaGetLastError());
    CTR_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return 0;
}
