// nf_ctx.cu -- context object of libnaviflow_b200 (stream, reduction scratch, error text).
#include <stdlib.h>

#include "nf_common.cuh"

#define NF_VERSION 100

extern "C" int nf_version(void) { return NF_VERSION; }

static thread_local std::string g_create_error;

extern "C" int nf_ctx_create(nf_ctx** out, int device, void* cuda_stream) {
  if (!out) return NF_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return NF_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return NF_ERR_ARG;
  }
  if (cudaSetDevice(device) != cudaSuccess) return NF_ERR_CUDA;
  nf_ctx* c = new nf_ctx();
  c->device = device;
  // NULL = the legacy default stream (what PyTorch uses unless told otherwise), so that library launches are
  // ordered with the caller's own work on that stream
  c->stream = (cudaStream_t)cuda_stream;
  c->owns_stream = false;
  bool ok = cudaMalloc(&c->partials, sizeof(double) * NF_MAX_PARTIALS * NF_MAX_RED) == cudaSuccess &&
            cudaMalloc(&c->ticket, sizeof(unsigned int) * 4) == cudaSuccess &&
            cudaMalloc(&c->scalars, sizeof(double) * NF_NUM_SCALARS) == cudaSuccess &&
            cudaMallocHost(&c->scalars_host, sizeof(double) * NF_NUM_SCALARS) == cudaSuccess;
  if (ok) {
    ok = cudaMemsetAsync(c->ticket, 0, sizeof(unsigned int) * 4, c->stream) == cudaSuccess &&
         cudaMemsetAsync(c->scalars, 0, sizeof(double) * NF_NUM_SCALARS, c->stream) == cudaSuccess &&
         cudaStreamSynchronize(c->stream) == cudaSuccess;
  }
  if (!ok) {
    g_create_error = std::string("context allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    nf_ctx_destroy(c);
    return NF_ERR_ALLOC;
  }
  *out = c;
  return NF_OK;
}

extern "C" int nf_ctx_destroy(nf_ctx* c) {
  if (!c) return NF_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->partials) cudaFree(c->partials);
  if (c->ticket) cudaFree(c->ticket);
  if (c->scalars) cudaFree(c->scalars);
  if (c->scalars_host) cudaFreeHost(c->scalars_host);
  if (c->owns_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return NF_OK;
}

extern "C" const char* nf_last_error(nf_ctx* c) {
  if (!c) return g_create_error.c_str();
  return c->err.c_str();
}

extern "C" int nf_sync(nf_ctx* c) {
  if (!c) return NF_ERR_ARG;
  NF_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return NF_OK;
}

extern "C" int64_t nf_launch_count(nf_ctx* c) { return c ? c->launches : 0; }

extern "C" void* nf_ctx_stream(nf_ctx* c) { return c ? (void*)c->stream : nullptr; }

int nf_read_scalars(nf_ctx* ctx, int first, int count, double* out_host) {
  NF_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->scalars_host + first, ctx->scalars + first, count * sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->stream));
  NF_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < count; ++k) out_host[k] = ctx->scalars_host[first + k];
  return NF_OK;
}

// NF_PDL=0: launch every kernel with the full stream dependency (see nf_common.cuh)
bool nfi_pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("NF_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
