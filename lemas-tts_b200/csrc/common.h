// Host-side helpers shared by the translation units of liblemas_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/lemas_b200.h"

namespace lemas {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int sm_count();
void count_launches(int n);
int64_t launches_so_far();

#define LEMAS_CUDA_OK(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::lemas::fail(LEMAS_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) +    \
                                               " at " __FILE__ ":" + std::to_string(__LINE__));      \
  } while (0)

// after a kernel launch: surface launch errors and account the launch (lemas_launch_count)
#define LEMAS_LAUNCHED(n)                    \
  do {                                       \
    LEMAS_CUDA_OK(cudaGetLastError());       \
    ::lemas::count_launches(n);              \
  } while (0)

#define LEMAS_REQUIRE(cond, msg)                                                   \
  do {                                                                             \
    if (!(cond)) return ::lemas::fail(LEMAS_ERR_INVALID, std::string(msg));        \
  } while (0)

#define LEMAS_TRY(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != LEMAS_OK) return _rc; \
  } while (0)

// fp16 row-major tensor map with a 64-element (128 B) inner box and the 128-byte swizzle.
// dims/strides innermost first; strides in bytes for dims 1.. (multiples of 16).
int make_tensor_map_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box);

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember it per (kernel instantiation, device), so a
// process that drives several GPUs configures each of them.  `done` is a function-local static bitmask of the caller.
template <class K>
inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, unsigned long long& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && ((done >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) done |= 1ull << dev;
  return e;
}

// LEMAS_PDL=0 disables programmatic dependent launch (A/B measurements)
bool pdl_enabled();

// Launch with the programmatic-stream-serialization attribute: the kernel's prologue (barrier init, TMEM
// allocation, descriptor prefetch, instruction fetch) overlaps the tail of the previous kernel in the stream; the
// kernel calls pdl_wait() before its first global access.  Works eagerly and under stream capture.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace lemas
