// Host helper: build CUtensorMap descriptors without linking libcuda (driver entry point lookup).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdexcept>
#include <string>

namespace nrl {

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || ptr == nullptr)
      throw std::runtime_error("cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 2D row-major tensor [rows, cols] of `elem_bytes`-wide elements, box [box_rows, box_cols], 128B swizzle.
// box_cols * elem_bytes must be 128 (one swizzle row).
inline CUtensorMap make_tma_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                               uint32_t box_rows, uint32_t box_cols, CUtensorMapDataType dtype, int elem_bytes) {
  // The driver call needs the primary context current on THIS thread.  Autograd worker threads only get it once a
  // runtime-API call has run there, and the caching allocator can satisfy torch::empty without one.
  thread_local bool ctx_bound = (cudaFree(nullptr) == cudaSuccess);
  (void)ctx_bound;
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (box_cols * elem_bytes != 128) throw std::runtime_error("TMA box inner extent must be 128 bytes");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15))
    throw std::runtime_error("TMA tensors need 16-byte aligned base and row stride");
  CUresult r = tensor_map_encoder()(&m, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string(int(r)));
  return m;
}

// 3D view [batch][rows][cols] of 2-byte elements (cols contiguous), box [1][box_rows][64], 128B swizzle: the batched GEMM's
// operands (e.g. per-head slices of a packed [T, H, 64] projection) without materialising per-batch copies.
inline CUtensorMap make_tma_3d(const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                               uint64_t batch_stride_bytes, uint32_t box_rows, uint32_t box_cols, CUtensorMapDataType dtype, int elem_bytes) {
  thread_local bool ctx_bound = (cudaFree(nullptr) == cudaSuccess);
  (void)ctx_bound;
  CUtensorMap m;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {row_stride_bytes, batch_stride_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (box_cols * elem_bytes != 128) throw std::runtime_error("TMA box inner extent must be 128 bytes");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15) || (batch_stride_bytes & 15))
    throw std::runtime_error("TMA tensors need 16-byte aligned base and strides");
  CUresult r = tensor_map_encoder()(&m, dtype, 3, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(3D) failed: " + std::to_string(int(r)));
  return m;
}

// Same, without swizzle: the box lands densely in shared memory (box_cols * elem_bytes per row, a multiple of 16).
inline CUtensorMap make_tma_2d_plain(const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                                     uint32_t box_rows, uint32_t box_cols, CUtensorMapDataType dtype, int elem_bytes) {
  thread_local bool ctx_bound = (cudaFree(nullptr) == cudaSuccess);
  (void)ctx_bound;
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if ((box_cols * elem_bytes) % 16 != 0 || box_cols > 256 || box_rows > 256)
    throw std::runtime_error("TMA box: inner extent must be a multiple of 16 bytes, extents <= 256");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15))
    throw std::runtime_error("TMA tensors need 16-byte aligned base and row stride");
  CUresult r = tensor_map_encoder()(&m, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string(int(r)));
  return m;
}

}  // namespace nrl
