// Host-side runtime of the C-ABI library: last-error string, TMA descriptor cache, device queries.
// The library never allocates or frees device memory and holds no references to caller buffers; the descriptor
// cache is keyed by (pointer, shape, pitch, box) and is the only mutable global state (mutex-guarded).
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>

namespace obt {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return OBT_ERR_CUDA;
  }
  return OBT_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency,
// so the library loads on a machine without a GPU driver).
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  return fn;
}

struct MapKey {
  uint64_t v[10];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 10; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

static int encode(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                  const uint32_t* box) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  int dev = 0;
  cudaGetDevice(&dev);
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = static_cast<uint64_t>(rank) | (static_cast<uint64_t>(dev) << 8);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[5 + i] = strides_elems[i];
    key.v[8] |= static_cast<uint64_t>(box[i]) << (16 * i);
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return OBT_OK;
    }
  }
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return OBT_ERR_CUDA;
  }
  if ((reinterpret_cast<uint64_t>(base) & 15) != 0) {
    set_last_error("TMA base pointer %p is not 16-byte aligned", base);
    return OBT_ERR_INVALID;
  }
  cuuint64_t gdim[3] = {1, 1, 1};
  cuuint64_t gstride[2] = {0, 0};  // byte strides of dims 1..rank-1
  cuuint32_t gbox[3] = {1, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    if (i > 0) {
      gstride[i - 1] = strides_elems[i] * 2;
      if ((gstride[i - 1] & 15) != 0) {
        set_last_error("TMA stride %llu bytes (dim %d) is not a multiple of 16", (unsigned long long)gstride[i - 1], i);
        return OBT_ERR_INVALID;
      }
    }
  }
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstride, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r,
                   rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2], gbox[0],
                   gbox[1], gbox[2]);
    return OBT_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps.emplace(key, m);
  }
  *out = m;
  return OBT_OK;
}

int get_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_elems,
                      uint32_t box0, uint32_t box1) {
  uint64_t dims[3] = {dim0, dim1, 1};
  uint64_t strides[3] = {1, stride1_elems, 0};
  uint32_t box[3] = {box0, box1, 1};
  return encode(out, base, 2, dims, strides, box);
}

int get_tensor_map_3d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, uint32_t box2) {
  uint64_t dims[3] = {dim0, dim1, dim2};
  uint64_t strides[3] = {1, stride1_elems, stride2_elems};
  uint32_t box[3] = {box0, box1, box2};
  return encode(out, base, 3, dims, strides, box);
}

}  // namespace obt

extern "C" const char* obt_last_error(void) { return obt::g_err; }

extern "C" int obt_version(void) { return 100; }

extern "C" void obt_clear_descriptor_cache(void) {
  std::lock_guard<std::mutex> lk(obt::g_map_mu);
  obt::g_maps.clear();
}
