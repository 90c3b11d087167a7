// Peer-memory plumbing of the synchronised BatchNorm (include/dcgc.h: dcgc_p2p_*): device allocations that the other
// processes of the node map through CUDA IPC and write with plain stores over NVLink.
#include <string.h>

#include "common.h"

extern "C" int64_t dcgc_bn_sync_mailbox_bytes(int32_t world, int32_t cap) {
  if (world < 1 || world > DCGC_SYNC_MAX_RANKS || cap < 1) return -1;
  return (int64_t)world * 2 * (2 * (int64_t)cap + 2) * 8;
}

extern "C" int dcgc_p2p_alloc(int64_t bytes, void** ptr_out, void* handle_out) {
  DCGC_CHECK_ARG(bytes > 0 && ptr_out && handle_out, "dcgc_p2p_alloc: bad arguments");
  void* p = nullptr;
  DCGC_CUDA_CALL(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    dcgc_set_error("dcgc_p2p_alloc: %s", cudaGetErrorString(e));
    return DCGC_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return DCGC_OK;
}

extern "C" int dcgc_p2p_open(const void* handle, void** ptr_out) {
  DCGC_CHECK_ARG(handle && ptr_out, "dcgc_p2p_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  DCGC_CUDA_CALL(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return DCGC_OK;
}

extern "C" int dcgc_p2p_close(void* ptr) {
  if (ptr) DCGC_CUDA_CALL(cudaIpcCloseMemHandle(ptr));
  return DCGC_OK;
}

extern "C" int dcgc_p2p_free(void* ptr) {
  if (ptr) DCGC_CUDA_CALL(cudaFree(ptr));
  return DCGC_OK;
}
