// Shared helpers for the lcgan_b200 kernels (sm_100a).
#pragma once
#include <mutex>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/lcgan_b200.h"

typedef __nv_bfloat16 bf16;

void lcgan_set_error(const char* fmt, ...);

#define LCGAN_CHECK(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      lcgan_set_error(__VA_ARGS__);            \
      return 1;                                \
    }                                          \
  } while (0)

#define LCGAN_CUDA(expr)                                                                \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      lcgan_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

#define LCGAN_LAUNCH_CHECK()                                                            \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      lcgan_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 3;                                                                         \
    }                                                                                   \
  } while (0)

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void pack(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void unpack(float* f) const {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
};


// 16-byte read-only global load that the compiler may not sink towards its first use: a batch of
// these issued back to back keeps that many requests in flight per thread (nvcc otherwise schedules
// each load of an unrolled streaming loop right before its consumer to save registers).
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <typename T> __device__ __forceinline__ void unpack_raw16(const uint4& u, float* f);
template <> __device__ __forceinline__ void unpack_raw16<bf16>(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <> __device__ __forceinline__ void unpack_raw16<float>(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// ---- shared-memory tile staging (cp.async, zero fill outside the image) ------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 16 : 0;                          // 0 source bytes: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_pending() {   // <= N groups still in flight
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
// window pixel (r, q) <- image pixel (wy0 + r, wx0 + q), the 64-byte channel chunk starting at
// element c0; PITCH bytes per window pixel; NT threads cooperate
// SWZ (PITCH must be 64): the 16-byte slot v of window pixel pq is stored at slot v ^ ((pq >> 1) & 3),
// which makes 16-byte reads of one slot from 8 consecutive pixels bank-conflict free without padding
// (read back with swz_slot).
__device__ __forceinline__ int swz_slot(int pq, int v) { return v ^ ((pq >> 1) & 3); }
template <typename T, int WW, int WH, int PITCH, int NT, bool SWZ = false>
__device__ __forceinline__ void load_window(unsigned char* sm, const T* __restrict__ img, int H, int W, int C,
                                            int c0, int wy0, int wx0) {
  constexpr int NV = WW * WH * 4;
  for (int i = threadIdx.x; i < NV; i += NT) {
    const int v = i & 3, pq = i >> 2;
    const int q = pq % WW, r = pq / WW;
    const int y = wy0 + r, x = wx0 + q;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(img);
    if (ok) src = reinterpret_cast<const unsigned char*>(img + ((int64_t)y * W + x) * C + c0) + v * 16;
    cp_async16(sm + pq * PITCH + (SWZ ? swz_slot(pq, v) : v) * 16, src, ok);
  }
}

// One-time initialisation PER DEVICE (cudaFuncSetAttribute applies to the current device only): run(f) calls f() the first
// time it is reached with each current device and returns f's result for that device afterwards.
struct DeviceOnce {
  std::mutex m;
  bool done[64] = {};
  int result[64] = {};
  template <class F> int run(F f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return f();
    std::lock_guard<std::mutex> g(m);
    if (!done[dev]) { result[dev] = f(); done[dev] = true; }
    return result[dev];
  }
};

// thin.cu: special-shape kernels tried before the generic tiled ones (-1 = not applicable)
int lcgan_thin_forward(const lcgan_tapconv& d, const void* x, const void* w, void* y, const float* rowscale,
                       const float* bias, const void* residual, cudaStream_t s);
int lcgan_thin_wgrad(const lcgan_tapconv& d, const void* x, const void* g, float* dw, float scale, cudaStream_t s);

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- deterministic mode (lcgan_set_deterministic) ------------------------------------------------
// Reductions that normally finish with fp32 atomics (split-K weight gradients, per-(b,c) sums) take
// TURNS instead: the contributors of one output region own a semaphore and add their partial sums with
// plain loads/stores in a fixed order (contributor k waits until k-1 has published), so two runs are
// bit-identical.  Contributors are ordered by their linear block index, and blocks are dispatched in
// that order, so a waiting block only ever waits on blocks that are already resident or finished (the
// serial split-K argument).  The last contributor resets the semaphore to 0: the pool needs no memset,
// but kernels using it must not run concurrently on two streams.
bool lcgan_det_enabled();
constexpr int kDetSems = 32768;

__device__ __forceinline__ void det_wait_turn(int* sem, int turn) {
  int v;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(sem) : "memory");
    if (v != turn) __nanosleep(32);
  } while (v != turn);
}
__device__ __forceinline__ void det_pass_turn(int* sem, int turn, int nturns) {
  const int next = turn + 1 == nturns ? 0 : turn + 1;
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(sem), "r"(next) : "memory");
}
// dst += v inside a turn (L2-coherent read: other blocks wrote the previous value)
__device__ __forceinline__ void det_add(float* dst, float v) { __stcg(dst, __ldcg(dst) + v); }
__device__ __forceinline__ void det_add4(float* dst, float4 v) {
  float4 o = __ldcg(reinterpret_cast<const float4*>(dst));
  o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
  __stcg(reinterpret_cast<float4*>(dst), o);
}
// whole-block turn: call with all threads of the block
__device__ __forceinline__ void det_block_begin(int* sem, int turn) {
  if (threadIdx.x == 0) det_wait_turn(sem, turn);
  __syncthreads();
}
__device__ __forceinline__ void det_block_end(int* sem, int turn, int nturns) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) det_pass_turn(sem, turn, nturns);
}
