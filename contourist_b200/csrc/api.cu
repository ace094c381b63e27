// Context management of the C ABI (include/contourist_b200.h).
#include "common.cuh"

static std::string g_create_err;

extern "C" int ctr_create(int device, ctr_ctx** out) {
  if (!out) return CTR_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return CTR_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_err = "device index out of range";
    return CTR_ERR_BAD_ARG;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return CTR_ERR_CUDA;
  }
  ctr_ctx* c = new ctr_ctx();
  c->device = device;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete c;
    return CTR_ERR_CUDA;
  }
  c->own_stream = true;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = c;
  return 0;
}

static void free_buf(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

extern "C" void ctr_destroy(ctr_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  ctr_comm_destroy(c);
  DevBuf* all[] = {&c->field, &c->bits, &c->nbits, &c->vbase, &c->tbase, &c->list_v, &c->list_t, &c->tile_state,
                   &c->counters, &c->wmask, &c->wdir, &c->wlist, &c->vox_tab, &c->verts, &c->normals, &c->tris, &c->keys, &c->lowmin, &c->cells, &c->codes};
  for (DevBuf* b : all) free_buf(*b);
  for (DevBuf& b : c->aux) free_buf(b);
  if (c->counters_host) cudaFreeHost(c->counters_host);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (c->ev_enqueued) cudaEventDestroy(c->ev_enqueued);
  if (c->ev_tail) cudaEventDestroy(c->ev_tail);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char* ctr_last_error(const ctr_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

extern "C" int ctr_set_stream(ctr_ctx* c, void* s) {
  if (!c) return CTR_ERR_BAD_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return 0;
}

extern "C" int ctr_set_timing(ctr_ctx* c, int enabled) {
  if (!c) return CTR_ERR_BAD_ARG;
  c->timing = enabled != 0;
  return 0;
}

extern "C" int ctr_stage_times(ctr_ctx* c, float* ms, int n) {
  if (!c || !ms) return CTR_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) ms[i] = i < CTR_NSTAGE ? c->stage_ms[i] : 0.f;
  return 0;
}

extern "C" int64_t ctr_kernel_launches(const ctr_ctx* c) { return c ? c->launches : 0; }

extern "C" int ctr_host_alloc(ctr_ctx* c, uint64_t bytes, void** out) {
  if (!c || !out) return CTR_ERR_BAD_ARG;
  *out = nullptr;
  CTR_CUDA(c, cudaSetDevice(c->device));
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ctr_fail(c, CTR_ERR_OOM, "cudaHostAlloc failed", cudaGetErrorString(e));
  }
  return 0;
}

extern "C" int ctr_host_free(ctr_ctx* c, void* p) {
  if (!c) return CTR_ERR_BAD_ARG;
  if (p) CTR_CUDA(c, cudaFreeHost(p));
  return 0;
}

extern "C" int ctr_stage_upload(ctr_ctx* c, const void* host, uint64_t dst_offset, uint64_t bytes, uint64_t total_bytes,
                                void** device_base) {
  if (!c) return CTR_ERR_BAD_ARG;
  if (!device_base || (bytes && !host)) return ctr_fail(c, CTR_ERR_BAD_ARG, "null argument");
  if (dst_offset > total_bytes || bytes > total_bytes - dst_offset) return ctr_fail(c, CTR_ERR_BAD_ARG, "piece outside the staging buffer");
  CTR_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = ctr_ensure(c, c->field, total_bytes ? total_bytes : 1))) return rc;
  *device_base = c->field.p;
  if (bytes)
    CTR_CUDA(c, cudaMemcpyAsync((char*)c->field.p + dst_offset, host, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

extern "C" int ctr_wait_for(ctr_ctx* c, ctr_ctx* other) {
  if (!c || !other) return CTR_ERR_BAD_ARG;
  if (c->device != other->device) return ctr_fail(c, CTR_ERR_BAD_ARG, "contexts on different devices");
  if (c == other || c->stream == other->stream) return 0;          // one stream: already ordered
  CTR_CUDA(c, cudaSetDevice(c->device));
  if (!other->ev_tail) CTR_CUDA(c, cudaEventCreateWithFlags(&other->ev_tail, cudaEventDisableTiming));
  CTR_CUDA(c, cudaEventRecord(other->ev_tail, other->stream));
  CTR_CUDA(c, cudaStreamWaitEvent(c->stream, other->ev_tail, 0));
  return 0;
}
