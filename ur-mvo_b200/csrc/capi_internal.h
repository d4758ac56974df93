// csrc/capi_internal.h — pieces of the C-ABI layer shared by its translation units (capi.cu,
// fm_capi.cu): the context object behind urmvo_ctx and the error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/urmvo_b200.h"

namespace urmvo {
// records the message returned by urmvo_last_error() (thread-local) and returns `code`
int set_error(int code, const std::string& msg);
// Host threads the library may use for flattening / subset drawing: URMVO_B200_HOST_THREADS if set
// (one process per GPU on a shared host should divide the cores), else hardware_concurrency().
int host_threads();
}  // namespace urmvo

#define CU_TRY(expr)                                                                                  \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return urmvo::set_error(URMVO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

struct urmvo_ctx {
  int device = 0;
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  // optional NCCL communicator for the point-sharded BA (urmvo_comm_init)
  void* comm = nullptr;
  int rank = 0, world = 1;
  // grow-only device workspace lent to the one-shot BA calls (no cudaMalloc / cudaFree per call)
  unsigned char* ws_dev = nullptr;
  size_t ws_bytes = 0;
  bool ws_in_use = false;
  // reusable pinned staging buffer for the one-shot entry points
  void* pinned = nullptr;
  size_t pinned_size = 0;
  int ensure_pinned(size_t n) {
    if (n <= pinned_size) return 0;
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_size = 0;
    size_t want = std::max(n, (size_t)1 << 20);
    if (cudaMallocHost(&pinned, want) != cudaSuccess) return -1;
    pinned_size = want;
    return 0;
  }
};
