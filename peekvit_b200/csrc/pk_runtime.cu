// peekvit_b200 — per-process runtime: error text, device context, TMA descriptor cache.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

#include <cstdarg>
#include <mutex>
#include <unordered_map>

namespace pk {

static thread_local char tl_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PK_OK;
  set_last_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return PK_ERR_CUDA;
}

struct Context {
  int device = -1;
  int sms = 0;
  unsigned int* flag = nullptr;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
};
// One context per device of the process (a model on cuda:1 next to one on cuda:0): every entry point works on the
// context of the CALLER'S CURRENT device, which the host wrappers set to the device of the tensors they pass.
constexpr int kMaxDevices = 64;
static Context g_ctxs[kMaxDevices];
static std::mutex g_mu;

static Context& cur_ctx() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return g_ctxs[d];
}
#define g_ctx (cur_ctx())

int num_sms() { return g_ctx.sms > 0 ? g_ctx.sms : 148; }
unsigned int* device_flag_ptr() { return g_ctx.flag; }

static int init_context(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (device < 0 || device >= kMaxDevices) {
    set_last_error("pk_init: device %d out of range", device);
    return PK_ERR_INVALID;
  }
  Context& c = g_ctxs[device];
  if (c.device == device && c.flag) return PK_OK;
  int prev = -1;
  PK_CHECK_CUDA(cudaGetDevice(&prev));          // the caller's current device is left as it was
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
  PK_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("peekvit_b200 needs an sm_100a device (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    return PK_ERR_UNSUPPORTED;
  }
  c.sms = prop.multiProcessorCount;
  PK_CHECK_CUDA(cudaMalloc(&c.flag, sizeof(unsigned int)));
  PK_CHECK_CUDA(cudaMemset(c.flag, 0, sizeof(unsigned int)));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return PK_ERR_UNSUPPORTED;
  }
  c.encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  c.device = device;
  return PK_OK;
}

struct TmapKey {
  const void* base;
  uint64_t rows, cols, ld;
  uint32_t box_rows, box_cols;
  int elem_bytes, swizzle;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols &&
           elem_bytes == o.elem_bytes && swizzle == o.swizzle;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h = h * 1000003u ^ k.rows;
    h = h * 1000003u ^ k.cols;
    h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ (static_cast<size_t>(k.box_rows) << 16 | k.box_cols);
    h = h * 1000003u ^ static_cast<size_t>(k.elem_bytes * 1024 + k.swizzle);
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  if (!g_ctx.encode) {
    set_last_error("pk_init() has not been called");
    return PK_ERR_INVALID;
  }
  TmapKey key{base, rows, cols, ld_elems, box_rows, box_cols, elem_bytes, swizzle_bytes};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) {
    *out = it->second;
    return PK_OK;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = g_ctx.encode(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): base=%p elem=%d rows=%llu cols=%llu ld=%llu box=%ux%u swizzle=%d", static_cast<int>(r),
                   base, elem_bytes, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols,
                   swizzle_bytes);
    return PK_ERR_CUDA;
  }
  if (g_tmaps.size() > 8192) g_tmaps.clear();
  g_tmaps.emplace(key, *out);
  return PK_OK;
}

// 3-D bf16 view (column, row in sample, sample) of a packed [batch*rows, ld] buffer; box = 64 columns x
// box_rows rows x 1 sample, 128-byte swizzle.  Rows past `rows` of a sample are out of bounds (zero-filled
// on load, clipped on store) instead of aliasing the next sample.
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps3;
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld_elems,
                      uint32_t box_rows) {
  if (!g_ctx.encode) {
    set_last_error("pk_init() has not been called");
    return PK_ERR_INVALID;
  }
  TmapKey key{base, rows, cols, ld_elems, box_rows, static_cast<uint32_t>(batch), 2, 128};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tmaps3.find(key);
  if (it != g_tmaps3.end()) {
    *out = it->second;
    return PK_OK;
  }
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {ld_elems * 2ull, rows * ld_elems * 2ull};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = g_ctx.encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(3d) failed (%d): base=%p cols=%llu rows=%llu batch=%llu ld=%llu box_rows=%u", static_cast<int>(r),
                   base, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch, (unsigned long long)ld_elems, box_rows);
    return PK_ERR_CUDA;
  }
  if (g_tmaps3.size() > 4096) g_tmaps3.clear();
  g_tmaps3.emplace(key, *out);
  return PK_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, base, 2, rows, cols, ld_elems, box_rows, box_cols, 128);
}

}  // namespace pk

extern "C" int pk_abi_version(void) { return PK_ABI_VERSION; }

extern "C" int pk_init(int device) { return pk::init_context(device); }

extern "C" const char* pk_last_error(void) { return pk::tl_error; }

extern "C" int pk_num_sms(void) { return pk::num_sms(); }

extern "C" int pk_device_flag_async(unsigned int* host_dst, void* stream) {
  using namespace pk;
  if (!g_ctx.flag || !host_dst) {
    set_last_error("pk_device_flag_async: pk_init() has not been called or null destination");
    return PK_ERR_INVALID;
  }
  return check_cuda(cudaMemcpyAsync(host_dst, g_ctx.flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)),
                    "pk_device_flag_async");
}

extern "C" int pk_device_flag(int reset) {
  using namespace pk;
  if (!g_ctx.flag) {
    set_last_error("pk_init() has not been called");
    return PK_ERR_INVALID;
  }
  unsigned int v = 0;
  PK_CHECK_CUDA(cudaDeviceSynchronize());
  PK_CHECK_CUDA(cudaMemcpy(&v, g_ctx.flag, sizeof(v), cudaMemcpyDeviceToHost));
  if (reset && v != 0) PK_CHECK_CUDA(cudaMemset(g_ctx.flag, 0, sizeof(unsigned int)));
  return static_cast<int>(v);
}
