// TMA window loads for the tiled CUDA-core kernels (flow warp, box filters): one thread issues a single
// cp.async.bulk.tensor for the whole (window x 64-byte channel chunk) box of a channels-last bf16 tensor instead of
// every thread computing addresses and bounds for a handful of 16-byte cp.async (ncu on the warp gather: ~45
// instructions per 16 bytes, 20 % of the kernel's instruction stream).  Out-of-image pixels are zero-filled by the
// unit; SWIZZLE_64B is exactly swz_slot()'s pattern (16-byte slot ^ ((pixel >> 1) & 3)) when the window base is
// 512-byte aligned.  Device side: one mbarrier per CTA, phases alternate per load.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>

namespace tmaw {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// called by ONE thread: arm the barrier with the bytes of all the copies of this phase ...
__device__ __forceinline__ void arm(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
// ... and start the copy of box (c0, x0, y0, b)
__device__ __forceinline__ void copy(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int x0, int y0, int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(x0), "r"(y0), "r"(b) : "memory");
}
__device__ __forceinline__ void load(void* dst, const CUtensorMap* map, uint64_t* bar, uint32_t bytes, int c0, int x0, int y0,
                                     int b) {
  arm(bar, bytes);
  copy(dst, map, bar, c0, x0, y0, b);
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "TMAW_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TMAW_DONE;\n"
      "bra TMAW_WAIT;\n"
      "TMAW_DONE:\n}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeFn encoder() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  });
  return fn;
}
// map over a dense channels-last bf16 tensor [N, H, W, C]; box = {32 channels, ww, wh, 1}.  false if it cannot be built
// (the callers then keep their cp.async path).
inline bool make_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int ww, int wh, bool swizzle) {
  EncodeFn enc = encoder();
  if (!enc || C % 32 != 0 || ((uintptr_t)base & 15) != 0 || ww > 256 || wh > 256) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)ww, (cuuint32_t)wh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tmaw
