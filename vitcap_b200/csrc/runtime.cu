// Host-side runtime shared by all entry points: error string, launch checks, SM count, and the TMA
// descriptor factory (cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so the library
// has no link-time dependency on libcuda and loads on a GPU-less build box).
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace vc {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return VC_ERR_LAUNCH;
  }
  return VC_OK;
}

static int g_pdl = -1;
int pdl_mode() {
  if (g_pdl < 0) {
    const char* e = getenv("VITCAP_PDL");
    g_pdl = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? (e[0] - '0') : 1;
  }
  return g_pdl;
}
void set_pdl_mode(int mode) { g_pdl = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }

// The library holds sm_100a code only (tcgen05 / TMEM / TMA): on any other device a launch would fail late with "no kernel
// image is available". Checked once per process by the first compute entry point (capi.cu VC_COUNT).
int check_device() {
  static int state = 0;                       // 0 = not checked, 1 = ok, -1 = unsupported
  if (state == 0) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
      cudaGetLastError();
      set_last_error("vitcap_b200: no usable CUDA device");
      return VC_ERR_UNSUPPORTED;
    }
    const char* fake = getenv("VITCAP_FAKE_CC");          // test hook: pretend the device reports this compute capability
    if (fake != nullptr && fake[0] != 0) { major = atoi(fake) / 10; minor = atoi(fake) % 10; }
    state = (major == 10) ? 1 : -1;
    if (state < 0) set_last_error("vitcap_b200 is built for sm_100a (B200) only; this device reports compute capability %d.%d", major, minor);
  }
  if (state < 0) {
    if (last_error()[0] == 0) set_last_error("vitcap_b200 is built for sm_100a (B200) only");
    return VC_ERR_UNSUPPORTED;
  }
  return VC_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  uint64_t v[8];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 8; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

static int encode(CUtensorMap* out, const TmapKey& key, int rank, const void* base, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box,
#ifdef VC_STORE_F16
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
#else
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
#endif
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return VC_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return VC_ERR_DRIVER; }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu box %u,%u pitch %llu", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], (unsigned long long)strides_bytes[0]);
    return VC_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> g(g_tmap_mu);
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps.emplace(key, *out);
  return VC_OK;
}

int get_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  TmapKey key = {{(uint64_t)(uintptr_t)base, rows, cols, ld, box_rows, box_cols, 2, 0}};
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(out, key, 2, base, dims, strides, box);
}

int get_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    uint32_t box_cols) {
  TmapKey key = {{(uint64_t)(uintptr_t)base, rows, cols, ld, box_rows, box_cols, 2, 4}};
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(out, key, 2, base, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

int get_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint64_t ld1, uint64_t ld2,
                     uint32_t box_rows, uint32_t box_cols) {
  TmapKey key = {{(uint64_t)(uintptr_t)base, d2, d1, d0, ld1, ld2, ((uint64_t)box_rows << 32) | box_cols, 3}};
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1 * 2, ld2 * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  return encode(out, key, 3, base, dims, strides, box);
}

}  // namespace vc
