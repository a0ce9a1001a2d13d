// TMA (cp.async.bulk.tensor) helpers for sm_100a: host-side tensor-map encoding without linking libcuda
// (cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint) and the device-side PTX wrappers.
//
// "Split" activation format (the TMA-native layout of the fused ResBlock units):
//   an activation tensor [B, C, L] is stored pre-activated and pre-split as fp16 hi / lo in the blocked order the
//   UMMA A operand uses:   halfs[B][2 (hi, lo)][C/8][L][8]
//   i.e. plane p = (b*2 + half)*(C/8) + kc holds rows t = 0..L-1 of 16 bytes (8 channels) each.  4*B*C*L bytes, the
//   same as the fp32 tensor.  A consumer's A tile (rows g0 .. g0+rows-1 of every plane) is fetched by TMA through a 2-D
//   tensor map over 8-byte elements {2*L, planes} with a {2*RB, 1} box; rows outside [0, L) are zero-filled by the
//   hardware = the convolution's zero padding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

namespace fv {

typedef CUresult (*tma_encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tma_encode_fn_t tma_encode_fn() {
  static tma_encode_fn_t fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tma_encode_fn_t>(p);
  });
  return fn;
}

constexpr int TMA_SPLIT_RB = 128;   // rows per full box of the split format (2 KB)

// Tensor map over a split activation buffer: `planes` planes of L rows x 16 B; box = `box_rows` rows of one plane.
// Returns false when the driver entry point is missing or the shape is not encodable (caller falls back).
inline bool tma_encode_split(CUtensorMap* tm, const void* base, long long L, long long planes, int box_rows) {
  tma_encode_fn_t fn = tma_encode_fn();
  if (!fn || box_rows <= 0 || box_rows > 128 || L <= 0 || planes <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)(2 * L), (cuuint64_t)planes};
  const cuuint64_t gstride[1] = {(cuuint64_t)(16 * L)};     // bytes between planes (multiple of 16)
  const cuuint32_t box[2] = {(cuuint32_t)(2 * box_rows), 1u};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tensor map over a plain fp32 activation [rows_total = B*C][L]: box = `box_cols` consecutive samples of `box_rows`
// channels.  L*4 must be a multiple of 16.
inline bool tma_encode_f32_rows(CUtensorMap* tm, const void* base, long long L, long long rows_total, int box_cols,
                                int box_rows) {
  tma_encode_fn_t fn = tma_encode_fn();
  if (!fn || box_cols <= 0 || box_cols > 256 || box_rows <= 0 || box_rows > 256 || (L & 3) || (box_cols & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)L, (cuuint64_t)rows_total};
  const cuuint64_t gstride[1] = {(cuuint64_t)(4 * L)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// global (tensor map, coordinates {c0 = fastest, c1}) -> shared; completes `box bytes` on the mbarrier.  Elements of the
// box that fall outside the tensor are written as zeros (and still counted in the byte total).
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
#endif

}  // namespace fv
