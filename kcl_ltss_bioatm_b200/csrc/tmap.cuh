// Host-side TMA tensor-map construction.  The driver symbol is resolved at run time through the
// CUDA runtime (no link-time dependency on libcuda, so the library also loads on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace plume {

void set_error(const std::string& msg);  // api.cu

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
    set_error(std::string("cuTensorMapEncodeTiled not available: ") + cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(sym);
  return fn;
}

// A strided NHWC bf16 view: `ptr` addresses channel 0 of pixel (0,0,0); strides are in elements.
struct ActView {
  const void* ptr;
  int C, W, H, N;
  long long pix_stride, row_stride, img_stride;
  long long plane = 0;  // bf16x3 mode: element offset from `ptr` (the hi plane) to the lo plane of the same view
};
// The lo plane of a bf16x3 view as a view of its own.
inline ActView lo_plane(const ActView& v) {
  ActView l = v;
  l.ptr = static_cast<const char*>(v.ptr) + v.plane * 2;
  return l;
}

inline ActView dense_view(const void* ptr, int N, int H, int W, int C_total, int c_off, int C) {
  ActView v;
  v.ptr = static_cast<const char*>(ptr) + static_cast<long long>(c_off) * 2;
  v.C = C; v.W = W; v.H = H; v.N = N;
  v.pix_stride = C_total;
  v.row_stride = static_cast<long long>(W) * C_total;
  v.img_stride = static_cast<long long>(H) * W * C_total;
  return v;
}

// bf16 4-D map (C, W, H, N), 128-byte swizzle, zero fill outside the tensor.
inline int make_act_map(CUtensorMap* m, const ActView& v, int box_c, int box_w, int box_h,
                        int box_n) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  const cuuint64_t es = 2;  // bytes per bf16 element
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.pix_stride * es, (cuuint64_t)v.row_stride * es,
                           (cuuint64_t)v.img_stride * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) || (strides[0] & 15) || (strides[1] & 15) ||
      (strides[2] & 15)) {
    set_error("activation view is not 16-byte aligned (channel offsets/strides must be multiples of 8)");
    return -1;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled(4d) failed: %d dims=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u)",
             (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)dims[2], (unsigned long long)dims[3], box[0], box[1], box[2],
             box[3]);
    set_error(buf);
    return -1;
  }
  return 0;
}

// bf16 row-major matrix [rows][cols] (cols contiguous), 128-byte swizzle; box = (box_cols, box_rows).
inline int make_mat_map(CUtensorMap* m, const void* ptr, long long rows, long long cols,
                        int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15)) {
    set_error("weight matrix is not 16-byte aligned");
    return -1;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(2d) failed: %d rows=%lld cols=%lld box=(%d,%d)",
             (int)r, rows, cols, box_cols, box_rows);
    set_error(buf);
    return -1;
  }
  return 0;
}

}  // namespace plume
