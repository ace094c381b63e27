// Multi-GPU entry points of the C ABI (north_star): one process per GPU, every rank extracts its z-slab
// on its own; NCCL carries only
//   ctr_allgather_offsets : all-gather of the ranks' {n_verts, n_tris} straight from the device counters the last run
//                           published (no host round trip) -> exclusive prefix = global vertex / triangle offsets;
//   ctr_gather_mesh       : the optional gather of the mesh to one rank: variable-size point-to-point transfers
//                           (ncclSend / ncclRecv in one group) of positions, normals and triangles, triangle ids made
//                           global on the receiver (k_g_add_offset).
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process -- torch's -- or on the loader path), so
// that the library itself has no link-time dependency on it and still loads where there is no NCCL at all.
#include <dlfcn.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt64 = 4 };          // ncclDataType_t values used here (nccl.h: ncclInt8 = 0, ncclInt64 = 4)

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};

Nccl g_nccl;

bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);       // the copy already in the process (torch's), if any
    if (h) break;
  }
  if (!h)
    for (const char* n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
  if (!h) {
    g_nccl.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "");
    return false;
  }
#define CTR_SYM(field, name)                                          \
  *(void**)(&g_nccl.field) = dlsym(h, name);                          \
  if (!g_nccl.field) {                                                \
    g_nccl.err = std::string("libnccl lacks ") + name;                \
    return false;                                                     \
  }
  CTR_SYM(GetUniqueId, "ncclGetUniqueId")
  CTR_SYM(CommInitRank, "ncclCommInitRank")
  CTR_SYM(CommDestroy, "ncclCommDestroy")
  CTR_SYM(AllGather, "ncclAllGather")
  CTR_SYM(Send, "ncclSend")
  CTR_SYM(Recv, "ncclRecv")
  CTR_SYM(GroupStart, "ncclGroupStart")
  CTR_SYM(GroupEnd, "ncclGroupEnd")
  CTR_SYM(GetErrorString, "ncclGetErrorString")
#undef CTR_SYM
  g_nccl.lib = h;
  return true;
}

#define CTR_NCCL(ctx, call)                                                                        \
  do {                                                                                             \
    int r__ = (call);                                                                              \
    if (r__ != ncclSuccess) return ctr_fail(ctx, CTR_ERR_CUDA, "NCCL error", g_nccl.GetErrorString(r__)); \
  } while (0)

// triangles of rank r arrive with that rank's local ids (ids >= its vertex count name the next rank's first vertices):
// global id = local id + vertex offset of r, for both
__global__ void k_g_add_offset(int* __restrict__ tris, size_t n, int off) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) tris[t] += off;
}

}  // namespace

struct ctr_comm_state {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  long long* d_mine = nullptr;       // {n_verts, n_tris} of the last run, device
  long long* d_all = nullptr;        // [nranks][2], device
  long long* h_all = nullptr;        // pinned mirror
  std::vector<long long> counts;     // last all-gather, host
  bool have_counts = false;
  DevBuf g_verts, g_normals, g_tris; // gathered mesh (root)
  long long g_nv = 0, g_nt = 0;
  size_t g_gsz = 4;
  bool g_normals_valid = false;
};

extern "C" int ctr_comm_unique_id(void* id128) {
  if (!id128) return CTR_ERR_BAD_ARG;
  if (!nccl_load()) return CTR_ERR_UNSUPPORTED;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return CTR_ERR_CUDA;
  memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int ctr_comm_init(ctr_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "bad communicator arguments");
  if (!nccl_load()) return ctr_fail(ctx, CTR_ERR_UNSUPPORTED, g_nccl.err.c_str());
  if (ctx->comm) return ctr_fail(ctx, CTR_ERR_STATE, "the context already has a communicator");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  ctr_comm_state* cs = new ctr_comm_state();
  cs->rank = rank;
  cs->nranks = nranks;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  int r = g_nccl.CommInitRank(&cs->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    delete cs;
    return ctr_fail(ctx, CTR_ERR_CUDA, "ncclCommInitRank", g_nccl.GetErrorString(r));
  }
  cudaError_t e = cudaMalloc(&cs->d_mine, 16);
  if (e == cudaSuccess) e = cudaMalloc(&cs->d_all, (size_t)nranks * 16);
  if (e == cudaSuccess) e = cudaMallocHost(&cs->h_all, (size_t)nranks * 16);
  if (e != cudaSuccess) {
    g_nccl.CommDestroy(cs->comm);
    delete cs;
    return ctr_fail(ctx, CTR_ERR_OOM, "communicator buffers", cudaGetErrorString(e));
  }
  cs->counts.assign((size_t)nranks * 2, 0);
  ctx->comm = cs;
  // from now on every 3D run of this context publishes its counts where the all-gather reads them
  ctx->publish3 = cs->d_mine;
  return 0;
}

extern "C" int ctr_comm_destroy(ctr_ctx* ctx) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  ctr_comm_state* cs = ctx->comm;
  if (!cs) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->publish3 == cs->d_mine) ctx->publish3 = nullptr;
  if (cs->comm) g_nccl.CommDestroy(cs->comm);
  if (cs->d_mine) cudaFree(cs->d_mine);
  if (cs->d_all) cudaFree(cs->d_all);
  if (cs->h_all) cudaFreeHost(cs->h_all);
  for (DevBuf* b : {&cs->g_verts, &cs->g_normals, &cs->g_tris})
    if (b->p) cudaFree(b->p);
  delete cs;
  ctx->comm = nullptr;
  return 0;
}

extern "C" int ctr_allgather_offsets(ctr_ctx* ctx, int64_t* counts_out, int64_t* my_offsets, int64_t* totals) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  ctr_comm_state* cs = ctx->comm;
  if (!cs) return ctr_fail(ctx, CTR_ERR_STATE, "no communicator: call ctr_comm_init first");
  if (ctx->last_kind != 3 && !ctx->pending3) return ctr_fail(ctx, CTR_ERR_STATE, "no 3D run whose counts could be gathered");
  if (ctx->publish3 != cs->d_mine) return ctr_fail(ctx, CTR_ERR_STATE, "ctr_mt3d_publish_counts redirected the counts of this context");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // the run stored {n_verts, n_tris} at d_mine behind its last kernel on this stream: the collective reads them there
  CTR_NCCL(ctx, g_nccl.AllGather(cs->d_mine, cs->d_all, 2, ncclInt64, cs->comm, st));
  CTR_CUDA(ctx, cudaMemcpyAsync(cs->h_all, cs->d_all, (size_t)cs->nranks * 16, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  long long ov = 0, ot = 0, tv = 0, tt = 0;
  for (int r = 0; r < cs->nranks; ++r) {
    cs->counts[2 * r] = cs->h_all[2 * r];
    cs->counts[2 * r + 1] = cs->h_all[2 * r + 1];
    if (r < cs->rank) {
      ov += cs->h_all[2 * r];
      ot += cs->h_all[2 * r + 1];
    }
    tv += cs->h_all[2 * r];
    tt += cs->h_all[2 * r + 1];
  }
  cs->have_counts = true;
  if (counts_out) memcpy(counts_out, cs->h_all, (size_t)cs->nranks * 16);
  if (my_offsets) {
    my_offsets[0] = ov;
    my_offsets[1] = ot;
  }
  if (totals) {
    totals[0] = tv;
    totals[1] = tt;
  }
  return 0;
}

extern "C" int ctr_gather_mesh(ctr_ctx* ctx, int root, int64_t* total_verts, int64_t* total_tris) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  ctr_comm_state* cs = ctx->comm;
  if (!cs) return ctr_fail(ctx, CTR_ERR_STATE, "no communicator: call ctr_comm_init first");
  if (!cs->have_counts) return ctr_fail(ctx, CTR_ERR_STATE, "call ctr_allgather_offsets after the run first");
  if (root < 0 || root >= cs->nranks) return ctr_fail(ctx, CTR_ERR_BAD_ARG, "bad root");
  if (ctx->last_kind != 3 || (ctx->last_flags & CTR_NO_GEOMETRY)) return ctr_fail(ctx, CTR_ERR_STATE, "no completed ctr_mt3d_run with geometry");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t gsz = (ctx->last_flags & CTR_GEOM_F64) ? 8 : 4;
  const bool want_n = (ctx->last_flags & CTR_WANT_NORMALS) != 0;
  long long tv = 0, tt = 0;
  for (int r = 0; r < cs->nranks; ++r) {
    tv += cs->counts[2 * r];
    tt += cs->counts[2 * r + 1];
  }
  if (tv >= 0x7fffffffll) return ctr_fail(ctx, CTR_ERR_OVERFLOW, "gathered mesh needs more than 2^31 vertex ids");
  const long long my_v = cs->counts[2 * cs->rank], my_t = cs->counts[2 * cs->rank + 1];
  if (my_v != ctx->last_counts[0] || my_t != ctx->last_counts[1])
    return ctr_fail(ctx, CTR_ERR_STATE, "the gathered counts are not those of this context's last run");
  int rc;
  if (cs->rank == root) {
    if ((rc = ctr_ensure(ctx, cs->g_verts, (size_t)tv * 3 * gsz + 16))) return rc;
    if (want_n && (rc = ctr_ensure(ctx, cs->g_normals, (size_t)tv * 3 * gsz + 16))) return rc;
    if ((rc = ctr_ensure(ctx, cs->g_tris, (size_t)tt * 12 + 16))) return rc;
  }
  CTR_NCCL(ctx, g_nccl.GroupStart());
  if (cs->rank == root) {
    long long ov = 0, ot = 0;
    for (int r = 0; r < cs->nranks; ++r) {
      const long long nv = cs->counts[2 * r], nt = cs->counts[2 * r + 1];
      char* dv = (char*)cs->g_verts.p + (size_t)ov * 3 * gsz;
      char* dn = want_n ? (char*)cs->g_normals.p + (size_t)ov * 3 * gsz : nullptr;
      char* dt = (char*)cs->g_tris.p + (size_t)ot * 12;
      if (r == root) {
        if (nv) CTR_CUDA(ctx, cudaMemcpyAsync(dv, ctx->verts.p, (size_t)nv * 3 * gsz, cudaMemcpyDeviceToDevice, st));
        if (nv && want_n) CTR_CUDA(ctx, cudaMemcpyAsync(dn, ctx->normals.p, (size_t)nv * 3 * gsz, cudaMemcpyDeviceToDevice, st));
        if (nt) CTR_CUDA(ctx, cudaMemcpyAsync(dt, ctx->tris.p, (size_t)nt * 12, cudaMemcpyDeviceToDevice, st));
      } else {
        if (nv) CTR_NCCL(ctx, g_nccl.Recv(dv, (size_t)nv * 3 * gsz, ncclInt8, r, cs->comm, st));
        if (nv && want_n) CTR_NCCL(ctx, g_nccl.Recv(dn, (size_t)nv * 3 * gsz, ncclInt8, r, cs->comm, st));
        if (nt) CTR_NCCL(ctx, g_nccl.Recv(dt, (size_t)nt * 12, ncclInt8, r, cs->comm, st));
      }
      ov += nv;
      ot += nt;
    }
  } else {
    if (my_v) CTR_NCCL(ctx, g_nccl.Send(ctx->verts.p, (size_t)my_v * 3 * gsz, ncclInt8, root, cs->comm, st));
    if (my_v && want_n) CTR_NCCL(ctx, g_nccl.Send(ctx->normals.p, (size_t)my_v * 3 * gsz, ncclInt8, root, cs->comm, st));
    if (my_t) CTR_NCCL(ctx, g_nccl.Send(ctx->tris.p, (size_t)my_t * 12, ncclInt8, root, cs->comm, st));
  }
  CTR_NCCL(ctx, g_nccl.GroupEnd());
  if (cs->rank == root) {
    long long ov = 0, ot = 0;
    for (int r = 0; r < cs->nranks; ++r) {
      const long long nv = cs->counts[2 * r], nt = cs->counts[2 * r + 1];
      if (nt && ov) {
        const size_t n = (size_t)nt * 3;
        k_g_add_offset<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((int*)cs->g_tris.p + (size_t)ot * 3, n, (int)ov);
        ctx->launches++;
      }
      ov += nv;
      ot += nt;
    }
    CTR_CUDA(ctx, cudaGetLastError());
    cs->g_nv = tv;
    cs->g_nt = tt;
    cs->g_gsz = gsz;
    cs->g_normals_valid = want_n;
  }
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  if (total_verts) *total_verts = tv;
  if (total_tris) *total_tris = tt;
  return 0;
}

extern "C" int ctr_gathered_fetch(ctr_ctx* ctx, void* verts, void* normals, int32_t* tris) {
  if (!ctx) return CTR_ERR_BAD_ARG;
  ctr_comm_state* cs = ctx->comm;
  if (!cs || !cs->g_tris.p) return ctr_fail(ctx, CTR_ERR_STATE, "no gathered mesh on this rank");
  CTR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (verts && cs->g_nv) CTR_CUDA(ctx, cudaMemcpyAsync(verts, cs->g_verts.p, (size_t)cs->g_nv * 3 * cs->g_gsz, cudaMemcpyDeviceToHost, st));
  if (normals) {
    if (!cs->g_normals_valid) return ctr_fail(ctx, CTR_ERR_STATE, "the gathered runs had no normals");
    if (cs->g_nv) CTR_CUDA(ctx, cudaMemcpyAsync(normals, cs->g_normals.p, (size_t)cs->g_nv * 3 * cs->g_gsz, cudaMemcpyDeviceToHost, st));
  }
  if (tris && cs->g_nt) CTR_CUDA(ctx, cudaMemcpyAsync(tris, cs->g_tris.p, (size_t)cs->g_nt * 12, cudaMemcpyDeviceToHost, st));
  CTR_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}
