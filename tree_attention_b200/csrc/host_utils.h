// Host helpers shared by the kernel launchers: CUDA error checks and TMA tensor-map creation
// (cuTensorMapEncodeTiled is fetched through the runtime's driver entry point, so nothing links
// against libcuda directly).
// Reference: none (host helpers of the native layer; /root/reference/model.py has no native code).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace ta {

#define TA_CUDA_CHECK(expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(_e) + " at " +       \
                               __FILE__ + ":" + std::to_string(__LINE__) + " in " #expr);           \
    }                                                                                               \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 4-D tensor map over a (B, H, S, D)-indexed tensor given as element strides; the innermost
// dimension D must be contiguous.  Box = (box_d, box_s, 1, 1).  elem_bytes in {1, 2, 4}.
inline CUtensorMap make_tmap_bhsd(const void* base, int elem_bytes, int64_t B, int64_t H, int64_t S, int64_t D,
                                  int64_t stride_b, int64_t stride_h, int64_t stride_s, int box_d, int box_s,
                                  CUtensorMapSwizzle swizzle) {
  CUtensorMap m;
  CUtensorMapDataType dt = elem_bytes == 2   ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                           : elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(stride_s * elem_bytes), (cuuint64_t)(stride_h * elem_bytes),
                           (cuuint64_t)(stride_b * elem_bytes)};
  // A size-1 dimension may carry any stride from PyTorch; TMA needs multiples of 16 B.
  for (int i = 0; i < 3; ++i)
    if (dims[i + 1] == 1 || strides[i] == 0) strides[i] = (cuuint64_t)(D * elem_bytes) * (i == 0 ? 1 : dims[1]);
  cuuint32_t box[4] = {(cuuint32_t)box_d, (cuuint32_t)box_s, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) throw std::runtime_error("TMA base must be 16B aligned");
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0) throw std::runtime_error("TMA strides must be multiples of 16 bytes");
  CUresult r = get_encode_tiled()(&m, dt, 4, const_cast<void*>(base), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return m;
}

inline int num_sms(int device = -1) {
  static int cached[64] = {0};
  if (device < 0) TA_CUDA_CHECK(cudaGetDevice(&device));
  if (!cached[device]) TA_CUDA_CHECK(cudaDeviceGetAttribute(&cached[device], cudaDevAttrMultiProcessorCount, device));
  return cached[device];
}

}  // namespace ta
