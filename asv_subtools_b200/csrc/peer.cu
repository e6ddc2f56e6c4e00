// Embedding table replicated over the GPUs of one node WITHOUT a collective on the critical path.
//
// BASELINE configs[3] gathers every rank's (n, D) embeddings into one (world x n, D) table on every GPU before
// all-pairs scoring (the reference writes ark files per job and cats the scp lists,
// extract_xvectors_for_pytorch.sh:147-151).  The NCCL form is one all-gather AFTER the shard is extracted.  Here
// the table of every rank is a cudaMalloc allocation exported through CUDA IPC and mapped by its peers over
// NVLink, and each batch's embeddings are stored into all world copies the moment the batch's last layer has
// produced them (xvb_scatter_rows, one small kernel per batch on the batch's lane stream): the transfer rides
// under the next batches' GEMMs and the step ends with a barrier instead of a 2 GB exchange.
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"

namespace xvb {

struct PeerPtrs {
  float* p[XVB_MAX_PEERS];
};

// rows x D floats (D % 4 == 0) from src (pitch D) into every destination at row offset row0 (pitch ld)
__global__ void scatter_rows_kernel(const float* __restrict__ src, long long rows, int D, PeerPtrs dst, int ndst, long long row0,
                                    long long ld) {
  const int vec = D >> 2;
  const long long total = rows * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec;
    const int c = (int)(i - r * vec);
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * D) + c);
#pragma unroll 1
    for (int k = 0; k < ndst; ++k) *(reinterpret_cast<float4*>(dst.p[k] + (row0 + r) * ld) + c) = v;
  }
}

}  // namespace xvb

using namespace xvb;

extern "C" int xvb_ipc_alloc(void** ptr, size_t bytes) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(ptr && bytes > 0, "xvb_ipc_alloc: bad arguments");
  XVB_CUDA(cudaMalloc(ptr, bytes));
  return XVB_OK;
}

extern "C" int xvb_ipc_free(void* ptr) {
  if (ptr) XVB_CUDA(cudaFree(ptr));
  return XVB_OK;
}

extern "C" int xvb_ipc_export(void* ptr, void* handle64) {
  XVB_CHECK_ARG(ptr && handle64, "xvb_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == XVB_IPC_HANDLE_BYTES, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  XVB_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof h);
  return XVB_OK;
}

extern "C" int xvb_ipc_open(const void* handle64, void** ptr) {
  XVB_CHECK_ARG(ptr && handle64, "xvb_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  XVB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return XVB_OK;
}

extern "C" int xvb_ipc_close(void* ptr) {
  if (ptr) XVB_CUDA(cudaIpcCloseMemHandle(ptr));
  return XVB_OK;
}

extern "C" int xvb_scatter_rows(const float* src, int64_t rows, int D, float* const* tables, int ntables, int64_t row0,
                                int64_t ld, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  XVB_CHECK_ARG(src && tables && rows > 0 && D > 0 && D % 4 == 0 && ld % 4 == 0 && ld >= D && row0 >= 0,
                "xvb_scatter_rows: bad arguments (D and ld multiples of 4)");
  XVB_CHECK_ARG(ntables >= 1 && ntables <= XVB_MAX_PEERS, "xvb_scatter_rows: %d tables (1..%d)", ntables, XVB_MAX_PEERS);
  PeerPtrs d{};
  for (int k = 0; k < ntables; ++k) {
    XVB_CHECK_ARG(tables[k] && (uintptr_t)tables[k] % 16 == 0, "xvb_scatter_rows: table %d is null or unaligned", k);
    d.p[k] = tables[k];
  }
  const long long total = rows * (D >> 2);
  long long g = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 4;      // a small kernel on purpose: it runs next to the GEMMs
  if (g > cap) g = cap;
  scatter_rows_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(src, rows, D, d, ntables, row0, ld);
  XVB_LAUNCH_CHECK();
  return XVB_OK;
}
