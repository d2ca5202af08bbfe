// Flow warp of SynthesisBlock (reference custom_layers.py:127-134,151,161-165):
//   grid = linspace(-1,1) coordinates + tanh(flow) * max_flow_scale
//   out  = grid_sample(x, grid, mode='bicubic', padding_mode='zeros', align_corners=False)
// tanh, coordinate generation and the 16-tap bicubic gather are one kernel; the backward produces
// the feature gradient (vector atomics into an fp32 accumulator) and the flow gradient (warp-shuffle
// reduction over channels) in one pass.  A group of G lanes (G = min(32, C/V)) owns one pixel and
// strides over its 16-byte channel vectors, so every gather tap is a contiguous run.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
constexpr float kA = -0.75f;

__device__ __forceinline__ float cub_near(float u) { return ((kA + 2.f) * u - (kA + 3.f)) * u * u + 1.f; }
__device__ __forceinline__ float cub_far(float u) { return ((kA * u - 5.f * kA) * u + 8.f * kA) * u - 4.f * kA; }
__device__ __forceinline__ float dcub_near(float u) { return (3.f * (kA + 2.f) * u - 2.f * (kA + 3.f)) * u; }
__device__ __forceinline__ float dcub_far(float u) { return (3.f * kA * u - 10.f * kA) * u + 8.f * kA; }

__device__ __forceinline__ void cubic_w(float t, float* w) {
  w[0] = cub_far(t + 1.f); w[1] = cub_near(t); w[2] = cub_near(1.f - t); w[3] = cub_far(2.f - t);
}
__device__ __forceinline__ void cubic_dw(float t, float* w) {
  w[0] = dcub_far(t + 1.f); w[1] = dcub_near(t); w[2] = -dcub_near(1.f - t); w[3] = -dcub_far(2.f - t);
}

struct PixCoord {
  int x0, y0;       // floor of the source index
  float tx, ty;     // fractional parts
  float th0, th1;   // tanh(flow)
};

__device__ __forceinline__ PixCoord source_index(const float* __restrict__ flow, int64_t pix, int h, int w,
                                                 int H, int W, float scale) {
  PixCoord pc;
  const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
  pc.th0 = tanhf(f.x);
  pc.th1 = tanhf(f.y);
  const float gx = (2.f * (float)w / (float)(W - 1) - 1.f) + pc.th0 * scale;
  const float gy = (2.f * (float)h / (float)(H - 1) - 1.f) + pc.th1 * scale;
  const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f;
  const float iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  pc.x0 = (int)fx; pc.y0 = (int)fy;
  pc.tx = ix - fx; pc.ty = iy - fy;
  return pc;
}

template <typename T, int V> __device__ __forceinline__ void ldv(const T* p, float* f) {
  if constexpr (V == 1) f[0] = ldf(p); else { Vec16<T> v; v.load(p); v.unpack(f); }
}
template <typename T, int V> __device__ __forceinline__ void stv(T* p, const float* f) {
  if constexpr (V == 1) stf(p, f[0]); else { Vec16<T> v; v.pack(f); v.store(p); }
}

// Tap table of one output pixel: 16 (offset, weight) pairs with out-of-range taps clamped to a
// valid address and given weight 0, so the gather below is branch-free and all 16 vector loads of a
// channel vector are in flight together.
struct Taps {
  int off[16];     // pixel offset (y*W + x) inside the image
  float w[16];
};

__device__ __forceinline__ void make_taps(const PixCoord& pc, int H, int W, const float* wx, const float* wy, Taps& t) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = pc.y0 - 1 + j;
    const bool oky = yy >= 0 && yy < H;
    const int yc = min(max(yy, 0), H - 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = pc.x0 - 1 + i;
      const bool ok = oky && xx >= 0 && xx < W;
      const int xc = min(max(xx, 0), W - 1);
      t.off[j * 4 + i] = yc * W + xc;
      t.w[j * 4 + i] = ok ? wy[j] * wx[i] : 0.f;
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
warp_fwd_kernel(const T* __restrict__ x, const float* __restrict__ flow, T* __restrict__ out, int N, int H,
                int W, int C, float scale, int G) {
  const int cv = C / V;
  const int64_t npix = (int64_t)N * H * W;
  const int gl = threadIdx.x % G;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t pix = blockIdx.x * (int64_t)(kThreads / G) + threadIdx.x / G; pix < npix; pix += groups_per_grid) {
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((int64_t)W * H));
    const PixCoord pc = source_index(flow, pix, h, w, H, W, scale);
    float wx[4], wy[4];
    cubic_w(pc.tx, wx);
    cubic_w(pc.ty, wy);
    Taps tp;
    make_taps(pc, H, W, wx, wy, tp);
    const T* xb = x + (int64_t)b * H * W * C;
    for (int v = gl; v < cv; v += G) {
      float f[16][V];
#pragma unroll
      for (int k = 0; k < 16; ++k) ldv<T, V>(xb + (int64_t)tp.off[k] * C + v * V, f[k]);
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(f[k][i], tp.w[k], acc[i]);
      stv<T, V>(out + pix * C + v * V, acc);
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
warp_bwd_kernel(const T* __restrict__ x, const float* __restrict__ flow, const T* __restrict__ dout,
                float* __restrict__ dx, float* __restrict__ dflow, int N, int H, int W, int C, float scale,
                int G) {
  const int cv = C / V;
  const int64_t npix = (int64_t)N * H * W;
  const int gl = threadIdx.x % G;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / G);
  // all lanes of a warp run the same trip count (npix rounded up) so the shuffles stay converged
  const int64_t npix_pad = (npix + (kThreads / G) - 1) / (kThreads / G) * (kThreads / G);
  for (int64_t pix = blockIdx.x * (int64_t)(kThreads / G) + threadIdx.x / G; pix < npix_pad; pix += groups_per_grid) {
    const bool live = pix < npix;
    float gix = 0.f, giy = 0.f;
    PixCoord pc{};
    if (live) {
      const int w = (int)(pix % W);
      const int h = (int)((pix / W) % H);
      const int b = (int)(pix / ((int64_t)W * H));
      pc = source_index(flow, pix, h, w, H, W, scale);
      float wx[4], wy[4], dwx[4], dwy[4];
      cubic_w(pc.tx, wx); cubic_w(pc.ty, wy);
      cubic_dw(pc.tx, dwx); cubic_dw(pc.ty, dwy);
      Taps tp, tgx, tgy;                       // value weights and the two derivative weight sets
      make_taps(pc, H, W, wx, wy, tp);
      make_taps(pc, H, W, dwx, wy, tgx);
      make_taps(pc, H, W, wx, dwy, tgy);
      const int64_t boff = (int64_t)b * H * W * C;
      for (int v = gl; v < cv; v += G) {
        float g[V];
        ldv<T, V>(dout + pix * C + v * V, g);
        float f[16][V];
#pragma unroll
        for (int k = 0; k < 16; ++k) ldv<T, V>(x + boff + (int64_t)tp.off[k] * C + v * V, f[k]);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < V; ++i) dot = fmaf(f[k][i], g[i], dot);
          gix = fmaf(dot, tgx.w[k], gix);
          giy = fmaf(dot, tgy.w[k], giy);
          const float wgt = tp.w[k];
          if (wgt != 0.f) {
            float* dp = dx + boff + (int64_t)tp.off[k] * C + v * V;
            if constexpr (V % 4 == 0) {
#pragma unroll
              for (int i = 0; i < V; i += 4)
                atomicAdd(reinterpret_cast<float4*>(dp + i),
                          make_float4(g[i] * wgt, g[i + 1] * wgt, g[i + 2] * wgt, g[i + 3] * wgt));
            } else {
#pragma unroll
              for (int i = 0; i < V; ++i) atomicAdd(dp + i, g[i] * wgt);
            }
          }
        }
      }
    }
    // reduce over the G lanes of the group
    for (int o = G >> 1; o > 0; o >>= 1) {
      gix += __shfl_xor_sync(0xffffffffu, gix, o);
      giy += __shfl_xor_sync(0xffffffffu, giy, o);
    }
    if (live && gl == 0) {
      // d ix / d gx = W/2 ; d gx / d flow = scale * (1 - tanh^2)
      const float d0 = gix * (0.5f * (float)W) * scale * (1.f - pc.th0 * pc.th0);
      const float d1 = giy * (0.5f * (float)H) * scale * (1.f - pc.th1 * pc.th1);
      *reinterpret_cast<float2*>(dflow + pix * 2) = make_float2(d0, d1);
    }
  }
}

__device__ __forceinline__ void ld4(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float* f) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}

// Backward, run variant: a thread owns ONE channel vector of a run of kRun consecutive output pixels
// of a row and keeps the 4x4 bicubic footprint of the feature gradient in registers.  Consecutive
// pixels of a smooth flow field move the footprint by ~1 source column, so only the column that
// leaves the window is flushed (one vector atomic per row) - about 4 atomics per source pixel instead
// of 16.  The L2 atomic units are the bottleneck of this kernel (measured: 442 G fp32 adds/s, which
// is their peak), so atomics saved is time saved.  Any footprint jump (row change, backward step,
// step > 3 columns) flushes the whole window and re-bases it: always correct, only slower.
constexpr int kRun = 32;

template <typename T>
__global__ void __launch_bounds__(128)
warp_bwd_run_kernel(const T* __restrict__ x, const float* __restrict__ flow, const T* __restrict__ dout,
                    float* __restrict__ dx, float* __restrict__ dflow, int N, int H, int W, int C, float scale,
                    int G, int run) {
  constexpr int V = 4;                        // 4 channels per lane: one RED.128 per footprint cell
  const int cv = C / V;                       // == G (one vector per lane, G lanes per run)
  (void)cv;
  const int runs_per_row = W / run;
  const int64_t nruns = (int64_t)N * H * runs_per_row;
  const int gl = threadIdx.x % G;
  const int groups_per_block = 128 / G;
  const int64_t nruns_pad = (nruns + groups_per_block - 1) / groups_per_block * groups_per_block;
  for (int64_t rid = blockIdx.x * (int64_t)groups_per_block + threadIdx.x / G; rid < nruns_pad;
       rid += (int64_t)gridDim.x * groups_per_block) {
    const bool live = rid < nruns;
    const int rx = (int)(rid % runs_per_row);
    const int h = (int)((rid / runs_per_row) % H);
    const int b = (int)(rid / ((int64_t)runs_per_row * H));
    const int64_t boff = (int64_t)b * H * W * C + gl * V;
    float acc[4][4][V];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) acc[j][i][k] = 0.f;
    int wx0 = 0, wy0 = 0;
    bool valid = false;
    auto flush_col = [&](int col, int sx) {           // col is a compile-time constant at every call site
      if (sx < 0 || sx >= W) return;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sy = wy0 + j;
        if (sy < 0 || sy >= H) continue;
        float* dp = dx + boff + ((int64_t)sy * W + sx) * C;
        atomicAdd(reinterpret_cast<float4*>(dp),
                  make_float4(acc[j][col][0], acc[j][col][1], acc[j][col][2], acc[j][col][3]));
      }
    };
    for (int xi = 0; xi < run; ++xi) {
      const int w = rx * run + xi;
      const int64_t pix = ((int64_t)b * H + h) * W + w;
      float gix = 0.f, giy = 0.f;
      PixCoord pc{};
      if (live) {
        pc = source_index(flow, pix, h, w, H, W, scale);
        float wx[4], wy[4], dwx[4], dwy[4];
        cubic_w(pc.tx, wx); cubic_w(pc.ty, wy);
        cubic_dw(pc.tx, dwx); cubic_dw(pc.ty, dwy);
        const int ox = pc.x0 - 1, oy = pc.y0 - 1;
        if (!valid || oy != wy0 || ox < wx0 || ox > wx0 + 3) {
          if (valid) { flush_col(0, wx0); flush_col(1, wx0 + 1); flush_col(2, wx0 + 2); flush_col(3, wx0 + 3); }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int k = 0; k < V; ++k) acc[j][i][k] = 0.f;
          wx0 = ox; wy0 = oy; valid = true;
        } else {
          while (wx0 < ox) {                  // slide one source column to the right
            flush_col(0, wx0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int k = 0; k < V; ++k) {
                acc[j][0][k] = acc[j][1][k]; acc[j][1][k] = acc[j][2][k];
                acc[j][2][k] = acc[j][3][k]; acc[j][3][k] = 0.f;
              }
            ++wx0;
          }
        }
        float g[V];
        ld4(dout + pix * C + gl * V, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int yy = oy + j;
          const bool oky = yy >= 0 && yy < H;
          const int yc = min(max(yy, 0), H - 1);
          float f[4][V];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int xc = min(max(ox + i, 0), W - 1);
            ld4(x + boff + ((int64_t)yc * W + xc) * C, f[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = oky && (ox + i) >= 0 && (ox + i) < W;
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < V; ++k) dot = fmaf(f[i][k], g[k], dot);
            if (ok) {
              gix = fmaf(dot, wy[j] * dwx[i], gix);
              giy = fmaf(dot, dwy[j] * wx[i], giy);
            }
            const float wgt = wy[j] * wx[i];
#pragma unroll
            for (int k = 0; k < V; ++k) acc[j][i][k] = fmaf(g[k], wgt, acc[j][i][k]);
          }
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) {
        gix += __shfl_xor_sync(0xffffffffu, gix, o);
        giy += __shfl_xor_sync(0xffffffffu, giy, o);
      }
      if (live && gl == 0) {
        const float d0 = gix * (0.5f * (float)W) * scale * (1.f - pc.th0 * pc.th0);
        const float d1 = giy * (0.5f * (float)H) * scale * (1.f - pc.th1 * pc.th1);
        *reinterpret_cast<float2*>(dflow + pix * 2) = make_float2(d0, d1);
      }
    }
    if (live && valid) { flush_col(0, wx0); flush_col(1, wx0 + 1); flush_col(2, wx0 + 2); flush_col(3, wx0 + 3); }
  }
}

inline int group_size(int cv) {
  if (cv & (cv - 1)) return 1;   // not a power of two: one lane per pixel
  return cv < 32 ? cv : 32;
}

inline int grid_for_groups(int64_t npix, int G) {
  const int per_block = kThreads / G;
  int64_t b = (npix + per_block - 1) / per_block;
  const int64_t cap = 148LL * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int lcgan_warp_fwd(const void* x, const float* flow, void* out, int dt, int N, int H, int W, int C,
                              float flow_scale, void* stream) {
  LCGAN_CHECK(x && flow && out && N > 0 && H > 1 && W > 1 && C > 0, "warp_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t npix = (int64_t)N * H * W;
#define CALL(T, V)                                                                         \
  do {                                                                                     \
    const int G = group_size(C / V);                                                       \
    warp_fwd_kernel<T, V><<<grid_for_groups(npix, G), kThreads, 0, s>>>(                   \
        (const T*)x, flow, (T*)out, N, H, W, C, flow_scale, G);                            \
  } while (0)
  if (dt == LCGAN_F32) { if (C % 4 == 0) CALL(float, 4); else CALL(float, 1); }
  else if (dt == LCGAN_BF16) { if (C % 8 == 0) CALL(bf16, 8); else CALL(bf16, 1); }
  else { lcgan_set_error("warp_fwd: bad dtype %d", dt); return 1; }
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_warp_bwd(const void* x, const float* flow, const void* dout, float* dx_acc, float* dflow,
                              int dt, int N, int H, int W, int C, float flow_scale, void* stream) {
  LCGAN_CHECK(x && flow && dout && dx_acc && dflow && N > 0 && H > 1 && W > 1 && C > 0, "warp_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t npix = (int64_t)N * H * W;
#define CALL(T, V)                                                                         \
  do {                                                                                     \
    const int G = group_size(C / V);                                                       \
    const int run = W < kRun ? W : kRun;                                                   \
    const int G4 = C / 4;                                                                  \
    if (C % 4 == 0 && (G4 & (G4 - 1)) == 0 && G4 <= 32 && W % run == 0 && V > 1 &&          \
        getenv("LCGAN_WARP_BWD_V1") == nullptr) {                                          \
      const int64_t nruns = (int64_t)N * H * (W / run);                                    \
      const int gpb = 128 / G4;                                                            \
      int64_t blocks = (nruns + gpb - 1) / gpb;                                            \
      if (blocks > 148LL * 64) blocks = 148LL * 64;                                        \
      warp_bwd_run_kernel<T><<<(int)blocks, 128, 0, s>>>(                                  \
          (const T*)x, flow, (const T*)dout, dx_acc, dflow, N, H, W, C, flow_scale, G4, run); \
    } else {                                                                               \
      warp_bwd_kernel<T, V><<<grid_for_groups(npix, G), kThreads, 0, s>>>(                 \
          (const T*)x, flow, (const T*)dout, dx_acc, dflow, N, H, W, C, flow_scale, G);    \
    }                                                                                      \
  } while (0)
  if (dt == LCGAN_F32) { if (C % 4 == 0) CALL(float, 4); else CALL(float, 1); }
  else if (dt == LCGAN_BF16) { if (C % 8 == 0) CALL(bf16, 8); else CALL(bf16, 1); }
  else { lcgan_set_error("warp_bwd: bad dtype %d", dt); return 1; }
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}
