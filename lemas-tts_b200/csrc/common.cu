#include "common.h"
#include <cstring>

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace lemas {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static std::atomic<int64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launches_so_far() { return g_launches.load(); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LEMAS_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // The driver API is resolved at run time so the library links without libcuda (none in the build container).
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(LEMAS_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string m = "CUDA error: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") rank " +
                    std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) m += " " + std::to_string(dims[i]);
    m += " strides";
    for (int i = 0; i + 1 < rank; ++i) m += " " + std::to_string(strides_bytes[i]);
    return fail(LEMAS_ERR_CUDA, m);
  }
  return LEMAS_OK;
}

}  // namespace lemas

extern "C" {

const char* lemas_last_error(void) { return lemas::g_last_error.c_str(); }
int lemas_version(void) { return 100; }

int64_t lemas_launch_count(void) { return lemas::g_launches.load(); }

int lemas_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(lemas_gemm_desc);
    case 1: return (int)sizeof(lemas_dit_config);
    case 2: return (int)sizeof(lemas_dit_layer);
    case 3: return (int)sizeof(lemas_dit_weights);
    case 4: return (int)sizeof(lemas_sample_args);
    case 5: return (int)sizeof(lemas_vocos_layer);
    case 6: return (int)sizeof(lemas_vocos_weights);
    case 7: return (int)sizeof(lemas_text_block);
    case 8: return (int)sizeof(lemas_text_weights);
    case 9: return (int)sizeof(lemas_prosody_tdnn);
    case 10: return (int)sizeof(lemas_prosody_block);
    case 11: return (int)sizeof(lemas_prosody_weights);
    case 12: return (int)sizeof(lemas_bigvgan_block);
    case 13: return (int)sizeof(lemas_bigvgan_stage);
    case 14: return (int)sizeof(lemas_bigvgan_weights);
  }
  return -1;
}

int lemas_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
  // zero-filled device memory of the CURRENT device plus the 64-byte CUDA IPC handle another process opens it with
  if (!ptr || !handle64 || bytes <= 0) return lemas::fail(LEMAS_ERR_INVALID, "lemas_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    return lemas::fail(LEMAS_ERR_CUDA, std::string("CUDA error: lemas_peer_alloc: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return LEMAS_OK;
}

int lemas_peer_open(const void* handle64, void** ptr) {
  // maps the peer process's allocation for kernels of the CURRENT device (peer access over NVLink is enabled lazily)
  if (!ptr || !handle64) return lemas::fail(LEMAS_ERR_INVALID, "lemas_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess)
    return lemas::fail(LEMAS_ERR_CUDA, std::string("CUDA error: lemas_peer_open: ") + cudaGetErrorString(e));
  *ptr = p;
  return LEMAS_OK;
}

int lemas_peer_close(void* ptr) {
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) { (void)cudaGetLastError(); return LEMAS_ERR_CUDA; }
  return LEMAS_OK;
}

int lemas_peer_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) { (void)cudaGetLastError(); return LEMAS_ERR_CUDA; }
  return LEMAS_OK;
}

int lemas_device_supported(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return 0;
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? 1 : 0;
}
}
