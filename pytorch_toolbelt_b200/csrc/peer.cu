// Peer memory for the multi-GPU halo exchange (include/tb200.h, "peer memory"): CUDA-IPC exportable allocations and the
// two stream-ordered flag primitives the ragged-strip merge is choreographed with. One process per GPU; the data itself is
// written by the merge kernels (peer stores over NVLink), these kernels only order it.
#include <string.h>

#include "common.cuh"

namespace tb {

__global__ void peer_signal_kernel(unsigned long long* flag, unsigned long long value) {
  // everything launched before this kernel on the stream has completed; make it visible system-wide, then publish
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

__global__ void peer_wait_kernel(const unsigned long long* flag, unsigned long long value, long long timeout_ns,
                                 unsigned long long* status) {
  unsigned long long seen = 0;
  long long start = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
    if (seen >= value) return;
    long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - start > timeout_ns) break;
    __nanosleep(200);
  }
  if (status) *status = seen + 1ull;  // timed out: record what was seen (+1 so that 0 still means success)
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_peer_alloc(size_t bytes, void** dev_ptr_out) {
  TB_REQUIRE(dev_ptr_out && bytes > 0, TB_ERR_ARG, "tb_peer_alloc: bad argument");
  void* p = nullptr;
  TB_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaMemset");
  }
  *dev_ptr_out = p;
  return TB_OK;
}

int tb_peer_free(void* dev_ptr) {
  if (dev_ptr) TB_CUDA(cudaFree(dev_ptr));
  return TB_OK;
}

int tb_peer_export(const void* dev_ptr, void* handle_out_64) {
  TB_REQUIRE(dev_ptr && handle_out_64, TB_ERR_ARG, "tb_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  TB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle_out_64, &h, sizeof(h));
  return TB_OK;
}

int tb_peer_open(const void* handle_64, void** dev_ptr_out) {
  TB_REQUIRE(handle_64 && dev_ptr_out, TB_ERR_ARG, "tb_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, sizeof(h));
  void* p = nullptr;
  TB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr_out = p;
  return TB_OK;
}

int tb_peer_close(void* dev_ptr) {
  if (dev_ptr) TB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return TB_OK;
}

int tb_peer_read(const uint64_t* word, uint64_t* value_out_host) {
  TB_REQUIRE(word && value_out_host, TB_ERR_ARG, "tb_peer_read: null pointer");
  TB_CUDA(cudaMemcpy(value_out_host, word, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return TB_OK;
}

int tb_peer_signal(uint64_t* flag, uint64_t value, void* stream) {
  TB_REQUIRE(flag, TB_ERR_ARG, "tb_peer_signal: null flag");
  peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(flag), (unsigned long long)value);
  TB_LAUNCH_CHECK("peer_signal_kernel");
  return TB_OK;
}

int tb_peer_wait(const uint64_t* flag, uint64_t value, int timeout_ms, uint64_t* status, void* stream) {
  TB_REQUIRE(flag && timeout_ms > 0, TB_ERR_ARG, "tb_peer_wait: bad argument");
  peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(flag), (unsigned long long)value,
                                                      (long long)timeout_ms * 1000000ll,
                                                      reinterpret_cast<unsigned long long*>(status));
  TB_LAUNCH_CHECK("peer_wait_kernel");
  return TB_OK;
}

}  // extern "C"
