// Standalone hardware probe (NOT part of liblcgan_b200.so): can ONE haloed activation tile in shared
// memory serve all nine taps of a 3x3 stride-1 conv through UMMA descriptor start offsets?
//
//   lattice tile 8 wide x 16 tall (128 rows of M), stored WITH its halo as one (8+2) x (16+2)-pixel TMA box
//   (kc channels per pixel, SWIZZLE_64B for kc = 32 / SWIZZLE_128B for kc = 64); tap (dy, dx) reads
//       start = base + (dy*10 + dx) * row_bytes,   SBO = 10 * row_bytes   (one stored row between 8-row groups)
//   B is an identity matrix, so D[r][n] = A[r][n]: the accumulator IS the A tile the tensor core saw.
//   The host compares it with x[m0 + r/8 + dy - 1][n0 + r%8 + dx - 1][:] for every tap and every value of the
//   descriptor's base_offset field (bits 49-51).  See DESIGN.md section 9 item 2.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC \
//             scratch/probe/probe_halo.cu lcgan_b200/csrc/api.cu -lcuda -o scratch/probe/libprobe_halo.so
// run:    python scratch/probe/probe_halo.py
#include "../../lcgan_b200/csrc/conv_tc.cu"

namespace {

constexpr int kPW = 10, kPH = 18;            // stored tile: (8 + 2) x (16 + 2) pixels

__global__ void __launch_bounds__(128, 1)
probe_halo_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw, int kc, int n0,
                  int m0, int b, int base_off, int nrow0, float* __restrict__ out /* [9][128][32] */) {
  extern __shared__ uint8_t raw[];
  uint8_t* a = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const uint32_t row_bytes = (uint32_t)kc * 2;
  const uint32_t a_bytes = kPW * kPH * row_bytes;
  uint8_t* bt = a + ((a_bytes + 1023) & ~1023u);
  const uint32_t b_bytes = 32u * row_bytes;                      // rows [nrow0, nrow0 + 32) of the identity, kc columns
  uint64_t* full = (uint64_t*)(bt + ((b_bytes + 1023) & ~1023u));
  uint64_t* done = full + 1;
  uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    mbar_expect_tx(full, a_bytes + b_bytes, leader);
    tma_load_4d(a, &tmx, full, 0, n0 - 1, m0 - 1, b, leader);
    tma_load_2d(bt, &tmw, full, 0, nrow0, leader);
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(32, false, false);        // N = 32: D[r][n] = A[r][nrow0 + n]
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3, dx = t % 3;
      uint64_t ad = make_desc(smem_u32(a) + (uint32_t)(dy * kPW + dx) * row_bytes, 16, kPW * row_bytes, kc);
      ad |= (uint64_t)(base_off & 7) << 49;
      const uint64_t bd = make_desc(smem_u32(bt), 16, 8 * row_bytes, kc);
      for (int k = 0; k < kc / 16; ++k)
        umma_f16(tmem_base + (uint32_t)(t * 32), ad + 2 * k, bd + 2 * k, idesc, k != 0, leader);
    }
    umma_commit(done, leader);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < 32; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 32 + c), v);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) out[((size_t)t * 128 + r) * 32 + c + i] = __uint_as_float(v[i]);
    }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

}  // namespace

// x [N,H,W,C] bf16 channels-last (C = 32 or 64), eye [C][C] bf16 identity, out [9][128][32] f32 =
// channels [nrow0, nrow0 + 32) of the tile each tap saw
extern "C" int probe_halo(const void* x, const void* eye, float* out, int N, int H, int W, int C, int n0, int m0, int b,
                          int base_off, int nrow0, void* stream) {
  LCGAN_CHECK(C == 32 || C == 64, "probe_halo: C must be 32 or 64");
  CUtensorMap tmx, tmw;
  if (int e = make_act_map(&tmx, x, N, H, W, C, kPW, kPH, 1, 1, C)) return e;
  LCGAN_CHECK(nrow0 >= 0 && nrow0 + 32 <= C, "probe_halo: bad nrow0");
  if (int e = make_w_map(&tmw, eye, C, C, 32, C)) return e;
  const int smem = 1024 + ((kPW * kPH * C * 2 + 1023) & ~1023) + ((32 * C * 2 + 1023) & ~1023) + 64;
  LCGAN_CUDA(cudaFuncSetAttribute(probe_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_halo_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmx, tmw, C, n0, m0, b, base_off, nrow0, out);
  LCGAN_LAUNCH_CHECK();
  return 0;
}
