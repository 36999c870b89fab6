#include "host_common.h"

#include <stdarg.h>
#include <string.h>

#include "../../include/qfx.h"

namespace qfx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime API, so the library links (and loads) without a driver.
static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (encode_tiled_fn)p;
  }
  return fn;
}

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  encode_tiled_fn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return -2;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (base %p rank %d dims %llu,%llu stride0 %llu box %u,%u)", (int)r,
              base, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    return -3;
  }
  return 0;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace qfx

extern "C" const char* qfx_last_error(void) { return qfx::g_err; }
extern "C" int qfx_version(void) { return 1; }

// ---------------------------------------------------------------------------------------------------- peer memory
// Sharded frozen weights: each rank exports its shard through CUDA IPC; a block's weights are assembled with copy-engine pulls.
static_assert(sizeof(cudaIpcMemHandle_t) == QFX_PEER_HANDLE_BYTES, "handle size");

#define QFX_RT(call)                                                            \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      qfx::set_error("%s: %s", #call, cudaGetErrorString(e_));                  \
      return -1;                                                                \
    }                                                                           \
  } while (0)

extern "C" int qfx_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle) {
  if (bytes <= 0 || !ptr || !handle) {
    qfx::set_error("qfx_peer_alloc: bad arguments");
    return -1;
  }
  QFX_RT(cudaMalloc(ptr, (size_t)bytes));  // a dedicated allocation: the handle maps exactly this range (offset 0)
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {
    qfx::set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    cudaFree(*ptr);
    *ptr = nullptr;
    return -1;
  }
  memcpy(handle, &h, sizeof(h));
  return 0;
}
extern "C" int qfx_peer_free(void* ptr) {
  QFX_RT(cudaFree(ptr));
  return 0;
}
extern "C" int qfx_peer_open(const unsigned char* handle, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  QFX_RT(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int qfx_peer_close(void* ptr) {
  QFX_RT(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int qfx_peer_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  QFX_RT(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
