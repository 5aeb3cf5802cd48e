// Library-wide pieces of the C-ABI: version, error text, launch counter, DLPack descriptors.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "apg_common.h"

namespace apg {

std::string& last_error() {
  static thread_local std::string err;
  return err;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

static std::atomic<int64_t> g_launches{0};
void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace apg

extern "C" {

int apg_version(void) { return APG_VERSION; }

const char* apg_last_error(void) { return apg::last_error().c_str(); }

int64_t apg_launch_count(void) { return apg::g_launches.load(std::memory_order_relaxed); }

void apg_struct_sizes(int32_t out[5]) {
  out[0] = (int32_t)sizeof(apg_image_cfg);
  out[1] = (int32_t)sizeof(apg_image_buffers);
  out[2] = (int32_t)sizeof(apg_lidar_cfg);
  out[3] = (int32_t)sizeof(apg_lidar_buffers);
  out[4] = (int32_t)sizeof(apg_pcg64_state);
}

// ---- DLPack v0.8 descriptor (layout restated from the public dlpack.h; no header dependency) ----------
typedef struct {
  int32_t device_type;  // kDLCUDA = 2
  int32_t device_id;
} ApgDLDevice;
typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} ApgDLDataType;
typedef struct {
  void* data;
  ApgDLDevice device;
  int32_t ndim;
  ApgDLDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
} ApgDLTensor;
typedef struct ApgDLManagedTensor {
  ApgDLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct ApgDLManagedTensor*);
} ApgDLManagedTensor;

static void apg_dlpack_deleter(ApgDLManagedTensor* self) {
  if (!self) return;
  free(self->dl_tensor.shape);
  free(self);
}

void* apg_dlpack_wrap(void* d_ptr, int32_t device, int32_t dtype_code, int32_t bits, int32_t ndim,
                      const int64_t* shape) {
  if (ndim < 0 || ndim > 8 || (ndim > 0 && !shape)) {
    apg::fail(APG_ERR_INVALID, "apg_dlpack_wrap: bad ndim/shape");
    return nullptr;
  }
  ApgDLManagedTensor* m = static_cast<ApgDLManagedTensor*>(calloc(1, sizeof(ApgDLManagedTensor)));
  if (!m) return nullptr;
  m->dl_tensor.data = d_ptr;
  m->dl_tensor.device.device_type = 2;
  m->dl_tensor.device.device_id = device;
  m->dl_tensor.ndim = ndim;
  m->dl_tensor.dtype.code = static_cast<uint8_t>(dtype_code);
  m->dl_tensor.dtype.bits = static_cast<uint8_t>(bits);
  m->dl_tensor.dtype.lanes = 1;
  m->dl_tensor.shape = static_cast<int64_t*>(malloc(sizeof(int64_t) * (ndim > 0 ? ndim : 1)));
  for (int i = 0; i < ndim; ++i) m->dl_tensor.shape[i] = shape[i];
  m->dl_tensor.strides = nullptr;  // compact row-major
  m->dl_tensor.byte_offset = 0;
  m->manager_ctx = nullptr;
  m->deleter = apg_dlpack_deleter;
  return m;
}

}  // extern "C"
