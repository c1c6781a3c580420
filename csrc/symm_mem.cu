// Symmetric memory: every rank allocates the same buffers and maps every peer's copy, so
// kernels address peer memory directly over NVLink (the transport that replaces the
// reference's gRPC parameter-server traffic).
//
// Backend 1 (this file, always available): cudaMalloc + CUDA IPC handles.  The 64-byte handles
// are exchanged by the Python side over the existing process group; opening a handle enables
// peer access lazily.
#include <string.h>

#include "host_utils.h"

extern "C" {

int dm_set_device(int device) {
  DM_CUDA_OK(cudaSetDevice(device));
  DM_CUDA_OK(cudaFree(0));
  return 0;
}

int dm_symm_alloc(unsigned long long bytes, void** out) {
  void* p = nullptr;
  DM_CUDA_OK(cudaMalloc(&p, bytes));
  DM_CUDA_OK(cudaMemset(p, 0, bytes));
  DM_CUDA_OK(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int dm_symm_free(void* p) {
  DM_CUDA_OK(cudaFree(p));
  return 0;
}

int dm_symm_ipc_handle(void* p, void* handle64) {
  cudaIpcMemHandle_t h;
  DM_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

int dm_symm_ipc_open(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  DM_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return 0;
}

int dm_symm_ipc_close(void* p) {
  DM_CUDA_OK(cudaIpcCloseMemHandle(p));
  return 0;
}

int dm_can_access_peer(int dev, int peer) {
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, dev, peer) != cudaSuccess) return -1;
  return ok;
}

int dm_memset_async(void* p, int value, unsigned long long bytes, void* stream) {
  DM_CUDA_OK(cudaMemsetAsync(p, value, bytes, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
