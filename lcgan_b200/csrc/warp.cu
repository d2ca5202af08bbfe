// Flow warp of SynthesisBlock (reference custom_layers.py:127-134,151,161-165):
//   grid = linspace(-1,1) coordinates + tanh(flow) * max_flow_scale
//   out  = grid_sample(x, grid, mode='bicubic', padding_mode='zeros', align_corners=False)
// tanh, coordinate generation and the 16-tap bicubic gather are one kernel; the backward produces
// the feature gradient (vector atomics into an fp32 accumulator) and the flow gradient (warp-shuffle
// reduction over channels) in one pass.  A group of G lanes (G = min(32, C/V)) owns one pixel and
// strides over its 16-byte channel vectors, so every gather tap is a contiguous run.
#include "common.cuh"
#include "tma_window.cuh"
#include <stdlib.h>
#include <limits.h>
#include <mutex>

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
  float ix, iy;     // unrounded source index
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
  pc.ix = ix; pc.iy = iy;
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


// ------------------------------------------------------------------------------------------
// Tiled kernels.  A CTA owns a 32x16 pixel tile of one image (one thread per pixel, a warp per row)
// and stages a window of the other tensor in shared memory, one 64-byte channel chunk at a time
// (32 bf16 / 16 fp32 channels; the four 16-byte slots of a pixel are XOR-swizzled so that 16-byte
// loads at a one-pixel lane stride are bank-conflict free).  The learned flows are smooth: a tile's
// samples land in a patch of about the tile's size, displaced by up to tens of pixels once training
// is under way (measured at 1024^2: < 3 px at init, mean 3 px / max 18 px after 56 iterations).  So
// the window is PLACED per tile and every feature vector is read from HBM/L2 about once.
//   * forward / dflow: window = bounding box of the tile's footprints (block reduction); a pixel
//     whose footprint leaves the 44x28 window falls back to global loads on its own.
//   * dx: GATHER formulation - source pixel s collects w(s,p) g[p] from the output pixels p whose
//     footprint contains s, w = k(|s.x - ix_p|) k(|s.y - iy_p|) with k the cubic convolution kernel.
//     No atomics, no fp32 accumulator tensor, dx is written once in the activation dtype.  A pre-pass
//     over the flow records, per SOURCE tile, the bounding box of the output pixels that reach it;
//     the dx kernel walks that box in 48x32 windows (one window when the flow is smooth).  A box
//     needing more than kMaxWindows windows raises a device flag: the tiled result is then
//     discarded and the atomic scatter kernels above (which otherwise exit at once) redo the
//     tensor.  Every decision is taken on the device, so the sequence is CUDA-graph capturable.
// ------------------------------------------------------------------------------------------
constexpr int kTW = 32, kTH = 16, kTileThreads = kTW * kTH;
constexpr int kPixB = 64;                                // XOR-swizzled 16-byte slots (swz_slot), no padding
constexpr int kFW = 44, kFH = 28;                        // footprint window (fwd, dflow): tile + 12
constexpr int kBW = 48, kBH = 32;                        // contributor window (dx): tile + 16
constexpr int kMaxWindows = 6;                           // dx: windows per source tile before giving up
constexpr int kFwdSmem = kFW * kFH * kPixB + 16;                     //  78 864 B (2 CTAs / SM)
constexpr int kFwdSmem8 = kFW * (8 + 12) * kPixB + 16;                //  56 336 B (4 CTAs / SM, 8-row tiles)
constexpr int kDxSmem = kBW * kBH * (kPixB + 8) + 32;                // 110 624 B (2 CTAs / SM)

// the scatter (atomic) kernels run only when the tiled dx kernel raised the flag
__device__ __forceinline__ bool tiled_ok(const int* __restrict__ flag) { return *flag == 0; }

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
                int G, const int* __restrict__ skip_if_tiled) {
  if (skip_if_tiled && tiled_ok(skip_if_tiled)) return;      // the tiled gather kernels did the work
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
          if (wgt != 0.f && dx) {                // dx == nullptr: flow gradient only
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

// dx by GATHER over a fixed window, for images too small for the tiled kernels (the 8x8 / 16x16 blocks): source
// pixel s collects w(s,p) g[p] from every output pixel p with |p - s| <= R per axis, R = the largest displacement
// the bounded flow (|tanh| * scale) plus the bicubic footprint allow.  No atomics: deterministic mode uses it in
// place of the scatter kernel.  A group of G lanes owns one source pixel and strides over its channel vectors.
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
warp_dx_window_kernel(const float* __restrict__ flow, const T* __restrict__ dout, T* __restrict__ dx, int N, int H,
                      int W, int C, float scale, int G, int Rx, int Ry) {
  const int cv = C / V;
  const int64_t npix = (int64_t)N * H * W;
  const int gl = threadIdx.x % G;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t pix = blockIdx.x * (int64_t)(kThreads / G) + threadIdx.x / G; pix < npix; pix += groups_per_grid) {
    const int sx = (int)(pix % W);
    const int sy = (int)((pix / W) % H);
    const int b = (int)(pix / ((int64_t)W * H));
    const int64_t boff = (int64_t)b * H * W;
    for (int v0 = gl; v0 < cv; v0 += G) {
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
      for (int h = max(sy - Ry, 0); h <= min(sy + Ry, H - 1); ++h)
        for (int w = max(sx - Rx, 0); w <= min(sx + Rx, W - 1); ++w) {
          const int64_t p = boff + (int64_t)h * W + w;
          const PixCoord pc = source_index(flow, p, h, w, H, W, scale);
          const int i = sx - pc.x0 + 1, j = sy - pc.y0 + 1;       // tap indices of s in p's footprint
          if (i < 0 || i > 3 || j < 0 || j > 3) continue;
          float wx[4], wy[4];
          cubic_w(pc.tx, wx);
          cubic_w(pc.ty, wy);
          const float wgt = wy[j] * wx[i];
          float g[V];
          ldv<T, V>(dout + p * C + v0 * V, g);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = fmaf(g[k], wgt, acc[k]);
        }
      stv<T, V>(dx + pix * C + v0 * V, acc);
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
                    int G, int run, const int* __restrict__ skip_if_tiled) {
  if (skip_if_tiled && tiled_ok(skip_if_tiled)) return;      // the tiled gather kernels did the work
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


// Mixed-precision FMA of sm_100 (PTX fma.rn.f32.bf16 -> SASS FHFMA.BF16 with .H0/.H1 operand selects): both factors are
// bf16 halves of packed registers, the product is exact and the accumulate is fp32 - so a packed bf16 pair needs no
// unpack instructions (ncu on the unpack + FFMA form: 54 % ALU-pipe, 31 % FMA-pipe, issue-bound).
//   acc0 += lo(x) * lo(w),  acc1 += hi(x) * lo(w)
__device__ __forceinline__ void fhfma_pair(uint32_t x, uint32_t w, float& acc0, float& acc1) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, wl, %0;\n\tfma.rn.f32.bf16 %1, xh, wl, %1;\n\t}"
      : "+f"(acc0), "+f"(acc1) : "r"(x), "r"(w));
}
//   acc += lo(x) * lo(g) + hi(x) * hi(g)
__device__ __forceinline__ void fhfma_dot(uint32_t x, uint32_t g, float& acc) {
  asm("{\n\t.reg .b16 xl, xh, gl, gh;\n\t"
      "mov.b32 {xl, xh}, %1;\n\tmov.b32 {gl, gh}, %2;\n\t"
      "fma.rn.f32.bf16 %0, xl, gl, %0;\n\tfma.rn.f32.bf16 %0, xh, gh, %0;\n\t}"
      : "+f"(acc) : "r"(x), "r"(g));
}
// A tap weight as the bf16 operand of fhfma_pair.  Rounding the 16 bicubic weights to bf16 (2^-9 relative each)
// perturbs the result by less than the bf16 rounding of the stored output does (sum w^2 < 1); fp32 mode is untouched.
__device__ __forceinline__ uint32_t bf16_weight(float w) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(w)); }

template <typename T>
__device__ __forceinline__ void fma_vec(const uint4& u, float w, uint32_t wb, float* acc) {
  if constexpr (sizeof(T) == 2) {
    fhfma_pair(u.x, wb, acc[0], acc[1]); fhfma_pair(u.y, wb, acc[2], acc[3]);
    fhfma_pair(u.z, wb, acc[4], acc[5]); fhfma_pair(u.w, wb, acc[6], acc[7]);
  } else {
    acc[0] = fmaf(__uint_as_float(u.x), w, acc[0]); acc[1] = fmaf(__uint_as_float(u.y), w, acc[1]);
    acc[2] = fmaf(__uint_as_float(u.z), w, acc[2]); acc[3] = fmaf(__uint_as_float(u.w), w, acc[3]);
  }
}
template <typename T>
__device__ __forceinline__ void dot_vec(const uint4& u, const uint4& g, float& dot) {
  if constexpr (sizeof(T) == 2) {
    fhfma_dot(u.x, g.x, dot); fhfma_dot(u.y, g.y, dot); fhfma_dot(u.z, g.z, dot); fhfma_dot(u.w, g.w, dot);
  } else {
    dot = fmaf(__uint_as_float(u.x), __uint_as_float(g.x), dot); dot = fmaf(__uint_as_float(u.y), __uint_as_float(g.y), dot);
    dot = fmaf(__uint_as_float(u.z), __uint_as_float(g.z), dot); dot = fmaf(__uint_as_float(u.w), __uint_as_float(g.w), dot);
  }
}

// acc[0..CC) += w * (64-byte pixel chunk at p)
template <typename T>
__device__ __forceinline__ void fma_chunk(const unsigned char* p, float w, float* acc) {
  constexpr int E = 16 / sizeof(T);
  const uint32_t wb = sizeof(T) == 2 ? bf16_weight(w) : 0u;
#pragma unroll
  for (int v = 0; v < 4; ++v) fma_vec<T>(*reinterpret_cast<const uint4*>(p + v * 16), w, wb, acc + v * E);
}
// <chunk at p, chunk g> with g kept packed (4 x uint4)
template <typename T>
__device__ __forceinline__ float dot_chunk(const unsigned char* p, const uint4* g) {
  float dot = 0.f;
#pragma unroll
  for (int v = 0; v < 4; ++v) dot_vec<T>(*reinterpret_cast<const uint4*>(p + v * 16), g[v], dot);
  return dot;
}
// the same on a swizzled window pixel (win = window base, pq = pixel index in the window)
template <typename T>
__device__ __forceinline__ void fma_chunk_win(const unsigned char* win, int pq, float w, float* acc) {
  constexpr int E = 16 / sizeof(T);
  const unsigned char* p = win + pq * kPixB;
  const int x = (pq >> 1) & 3;
  const uint32_t wb = sizeof(T) == 2 ? bf16_weight(w) : 0u;
#pragma unroll
  for (int v = 0; v < 4; ++v) fma_vec<T>(*reinterpret_cast<const uint4*>(p + ((v ^ x) << 4)), w, wb, acc + v * E);
}
template <typename T>
__device__ __forceinline__ float dot_chunk_win(const unsigned char* win, int pq, const uint4* g) {
  const unsigned char* p = win + pq * kPixB;
  const int x = (pq >> 1) & 3;
  float dot = 0.f;
#pragma unroll
  for (int v = 0; v < 4; ++v) dot_vec<T>(*reinterpret_cast<const uint4*>(p + ((v ^ x) << 4)), g[v], dot);
  return dot;
}
template <typename T>
__device__ __forceinline__ void store_chunk(T* dst, const float* acc) {
  constexpr int E = 16 / sizeof(T);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    Vec16<T> o;
    o.pack(acc + v * E);
    o.store(dst + v * E);
  }
}

// BWD = false: out[p] = sum_taps w x[tap]          (the forward warp)
// BWD = true : dflow[p] from <x[tap], g[p]> and the derivative weights
// Pre-pass: per SOURCE tile, the bounding box of the output pixels whose 4x4 footprint touches it.
// bbox[tile] = {max(-p.x), max(p.x), max(-p.y), max(p.y)}, pre-set very negative (memset 0x80).  A warp
// (32 pixels of a row) aggregates by source tile before the atomics.
// (called by all 32 lanes; live = this lane holds an output pixel (w, h) of image b with footprint origin x0, y0)
__device__ __forceinline__ void bbox_update(int* __restrict__ bbox, bool live, int b, int h, int w, int x0, int y0, int H,
                                            int W, int tiles_x, int tiles_y) {
  int fx0 = 1, fx1 = 0, fy0 = 1, fy1 = 0;                              // empty footprint
  if (live) {
    fx0 = max(x0 - 1, 0); fx1 = min(x0 + 2, W - 1);
    fy0 = max(y0 - 1, 0); fy1 = min(y0 + 2, H - 1);
  }
  const bool any = fx0 <= fx1 && fy0 <= fy1;
  const int tx0 = fx0 / kTW, tx1 = fx1 / kTW, ty0 = fy0 / kTH, ty1 = fy1 / kTH;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int tx = (k & 1) ? tx1 : tx0, ty = (k & 2) ? ty1 : ty0;
    const bool dup = ((k & 1) && tx1 == tx0) || ((k & 2) && ty1 == ty0);
    const int id = (any && !dup) ? (b * tiles_y + ty) * tiles_x + tx : -1;
    const unsigned grp = __match_any_sync(0xffffffffu, id);
    const int a0 = __reduce_max_sync(grp, -w), a1 = __reduce_max_sync(grp, w);
    const int a2 = __reduce_max_sync(grp, -h), a3 = __reduce_max_sync(grp, h);
    if (id >= 0 && (int)(threadIdx.x & 31) == __ffs(grp) - 1) {
      int* bb = bbox + (int64_t)id * 4;
      atomicMax(bb + 0, a0); atomicMax(bb + 1, a1); atomicMax(bb + 2, a2); atomicMax(bb + 3, a3);
    }
  }
}

__global__ void __launch_bounds__(256)
warp_bbox_kernel(const float* __restrict__ flow, int N, int H, int W, float scale, int* __restrict__ bbox) {
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const int64_t npix = (int64_t)N * H * W;
  const int64_t npix_pad = (npix + 31) / 32 * 32;
  for (int64_t pix = blockIdx.x * 256LL + threadIdx.x; pix < npix_pad; pix += (int64_t)gridDim.x * 256) {
    const bool live = pix < npix;
    int w = 0, h = 0, b = 0, x0 = 0, y0 = 0;
    if (live) {
      w = (int)(pix % W); h = (int)((pix / W) % H); b = (int)(pix / ((int64_t)W * H));
      const PixCoord pc = source_index(flow, pix, h, w, H, W, scale);
      x0 = pc.x0; y0 = pc.y0;
    }
    bbox_update(bbox, live, b, h, w, x0, y0, H, W, tiles_x, tiles_y);
  }
}

// TMA = true (bf16): the window arrives by one bulk tensor copy (tma_window.cuh) instead of per-thread cp.async
// TH = tile height (a warp per row): 16 -> 2 CTAs of 512 threads per SM, 8 -> 4 CTAs of 256 (more independent
// flow-load -> window-load -> gather latency chains in flight per SM)
template <typename T, bool BWD, bool TMA = false, int TH = kTH>
__global__ void __launch_bounds__(kTW * TH, 1024 / (kTW * TH))
warp_tile_gather_kernel(const __grid_constant__ CUtensorMap tmx, const T* __restrict__ x, const float* __restrict__ flow,
                        const T* __restrict__ g, T* __restrict__ out, float* __restrict__ dflow, int H, int W, int C,
                        float scale, const float* __restrict__ cs = nullptr, int* __restrict__ bbox = nullptr) {
  // cs [N,C] (forward only, optional): out = warp(x) * cs[b,c] - the style of the modulated conv that consumes
  // the warped features (the to-RGB block), folded into this pass
  constexpr int CC = 64 / sizeof(T);
  constexpr int FH = TH + 12;                              // window height
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + TH - 1) / TH;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int px = tx * kTW + (threadIdx.x & 31), py = ty * TH + (threadIdx.x >> 5);
  const bool live = px < W && py < H;
  const T* img = x + (int64_t)b * H * W * C;
  const int64_t pix = ((int64_t)b * H + py) * W + px;
  int* org = reinterpret_cast<int*>(smem + kFW * FH * kPixB);   // window origin (min x0, min y0 of the tile)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kFW * FH * kPixB + 8);
  uint32_t phase = 0;
  if (threadIdx.x < 2) org[threadIdx.x] = INT_MAX;
  if (TMA && threadIdx.x == 0) tmaw::bar_init(bar);
  __syncthreads();
  PixCoord pc{};
  float wx[4], wy[4], dwx[4], dwy[4];
  int mx = INT_MAX, my = INT_MAX;
  if (live) {
    pc = source_index(flow, pix, py, px, H, W, scale);
    cubic_w(pc.tx, wx);
    cubic_w(pc.ty, wy);
    if constexpr (BWD) { cubic_dw(pc.tx, dwx); cubic_dw(pc.ty, dwy); }
    // footprints entirely outside the image contribute nothing and must not drag the window away
    if (pc.x0 + 2 >= 0 && pc.x0 - 1 < W && pc.y0 + 2 >= 0 && pc.y0 - 1 < H) { mx = pc.x0; my = pc.y0; }
  }
  if constexpr (BWD) {
    // the dx kernel's per-source-tile contributor boxes, from the coordinates this pass computes anyway
    if (bbox) bbox_update(bbox, live, b, py, px, pc.x0, pc.y0, H, W, (W + kTW - 1) / kTW, (H + kTH - 1) / kTH);
  }
  mx = __reduce_min_sync(0xffffffffu, mx);
  my = __reduce_min_sync(0xffffffffu, my);
  if ((threadIdx.x & 31) == 0) { atomicMin(org + 0, mx); atomicMin(org + 1, my); }
  __syncthreads();
  const int wx0 = min(max(org[0] - 1, -4), W), wy0 = min(max(org[1] - 1, -4), H);
  const bool inwin = live && pc.x0 - 1 >= wx0 && pc.x0 + 2 < wx0 + kFW && pc.y0 - 1 >= wy0 && pc.y0 + 2 < wy0 + FH;
  float gix = 0.f, giy = 0.f;
  for (int c0 = 0; c0 < C; c0 += CC) {
    if (c0) __syncthreads();                             // everyone is done with the previous chunk
    if constexpr (TMA) {
      if (threadIdx.x == 0) tmaw::load(smem, &tmx, bar, kFW * FH * kPixB, c0, wx0, wy0, b);
    } else {
      load_window<T, kFW, FH, kPixB, kTW * TH, true>(smem, img, H, W, C, c0, wy0, wx0);
    }
    uint4 gv[4];
    if (BWD && live) {
#pragma unroll
      for (int v = 0; v < 4; ++v) gv[v] = *reinterpret_cast<const uint4*>(g + pix * C + c0 + v * (16 / sizeof(T)));
    }
    if constexpr (TMA) {
      tmaw::wait(bar, phase);
      phase ^= 1;
    } else {
      cp_async_wait_all();
      __syncthreads();
    }
    if (!live) continue;
    float acc[CC];
#pragma unroll
    for (int i = 0; i < CC; ++i) acc[i] = 0.f;
    if (inwin) {
      const int pq0 = (pc.y0 - 1 - wy0) * kFW + (pc.x0 - 1 - wx0);
      // one footprint row at a time (4 taps = 16 vector loads in flight); the row weights rotate
      // through scalars so the row loop needs no dynamic register indexing
      float ry0 = wy[0], ry1 = wy[1], ry2 = wy[2], ry3 = wy[3];
      float dy0 = 0.f, dy1 = 0.f, dy2 = 0.f, dy3 = 0.f;
      if constexpr (BWD) { dy0 = dwy[0]; dy1 = dwy[1]; dy2 = dwy[2]; dy3 = dwy[3]; }
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pq = pq0 + j * kFW + i;
          if constexpr (BWD) {
            const float dot = dot_chunk_win<T>(smem, pq, gv);
            gix = fmaf(dot, ry0 * dwx[i], gix);
            giy = fmaf(dot, dy0 * wx[i], giy);
          } else {
            fma_chunk_win<T>(smem, pq, ry0 * wx[i], acc);
          }
        }
        ry0 = ry1; ry1 = ry2; ry2 = ry3;
        dy0 = dy1; dy1 = dy2; dy2 = dy3;
      }
    } else {
      // footprint leaves the window: same arithmetic from global memory (out-of-image taps weigh 0)
      float ry0 = wy[0], ry1 = wy[1], ry2 = wy[2], ry3 = wy[3];
      float dy0 = 0.f, dy1 = 0.f, dy2 = 0.f, dy3 = 0.f;
      if constexpr (BWD) { dy0 = dwy[0]; dy1 = dwy[1]; dy2 = dwy[2]; dy3 = dwy[3]; }
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int yy = pc.y0 - 1 + j;
        const bool oky = yy >= 0 && yy < H;
        const T* rowp = img + (int64_t)min(max(yy, 0), H - 1) * W * C + c0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int xx = pc.x0 - 1 + i;
          const bool ok = oky && xx >= 0 && xx < W;
          const unsigned char* p = reinterpret_cast<const unsigned char*>(rowp + (int64_t)min(max(xx, 0), W - 1) * C);
          if constexpr (BWD) {
            const float dot = ok ? dot_chunk<T>(p, gv) : 0.f;
            gix = fmaf(dot, ry0 * dwx[i], gix);
            giy = fmaf(dot, dy0 * wx[i], giy);
          } else {
            fma_chunk<T>(p, ok ? ry0 * wx[i] : 0.f, acc);
          }
        }
        ry0 = ry1; ry1 = ry2; ry2 = ry3;
        dy0 = dy1; dy1 = dy2; dy2 = dy3;
      }
    }
    if constexpr (!BWD) {
      if (cs) {
        const float4* cp = reinterpret_cast<const float4*>(cs + (int64_t)b * C + c0);
#pragma unroll
        for (int i = 0; i < CC / 4; ++i) {
          const float4 c4 = __ldg(cp + i);
          acc[4 * i] *= c4.x; acc[4 * i + 1] *= c4.y; acc[4 * i + 2] *= c4.z; acc[4 * i + 3] *= c4.w;
        }
      }
      store_chunk<T>(out + pix * C + c0, acc);
    }
  }
  if (BWD && live) {
    const float d0 = gix * (0.5f * (float)W) * scale * (1.f - pc.th0 * pc.th0);
    const float d1 = giy * (0.5f * (float)H) * scale * (1.f - pc.th1 * pc.th1);
    *reinterpret_cast<float2*>(dflow + pix * 2) = make_float2(d0, d1);
  }
}

// cubic convolution kernel of a distance d >= 0 (0 for d >= 2): the tap weight of source pixel s for a
// sample at ix is k(|s - ix|) - identical to cubic_w(frac)[s - floor(ix) + 1]
__device__ __forceinline__ float cubic_k(float d) {
  const bool nr = d <= 1.f;
  const float c3 = nr ? (kA + 2.f) : kA;
  const float c2 = nr ? -(kA + 3.f) : -5.f * kA;
  const float c1 = nr ? 0.f : 8.f * kA;
  const float c0 = nr ? 1.f : -4.f * kA;
  return fmaf(fmaf(fmaf(c3, d, c2), d, c1), d, c0);
}

template <typename T, bool TMA = false>
__global__ void __launch_bounds__(kTileThreads, 2)
warp_tile_dx_kernel(const __grid_constant__ CUtensorMap tmg, const float* __restrict__ flow, const T* __restrict__ g,
                    T* __restrict__ dx, const int* __restrict__ bbox, int* __restrict__ flag, int H, int W, int C,
                    float scale) {
  constexpr int CC = 64 / sizeof(T);
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* gwin = smem;
  float2* coord = reinterpret_cast<float2*>(smem + kBW * kBH * kPixB);
  int* lb = reinterpret_cast<int*>(smem + kBW * kBH * (kPixB + 8));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBW * kBH * (kPixB + 8) + 16);
  uint32_t phase = 0;
  if (TMA && threadIdx.x == 0) tmaw::bar_init(bar);      // made visible by the first __syncthreads of the window loop
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  int t = blockIdx.x;
  const int* bb = bbox + (int64_t)t * 4;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  // a warp owns an 8 x 4 patch, not 32 pixels of a row: its lanes walk the candidate offsets in lockstep and run the
  // union of their hit sets, which grows with the spread of the displacement over the warp (ncu, 8 %-stretch flow:
  // 33 hit iterations per warp for 16-20 per lane with row-shaped warps).  A quarter-warp is still 8 consecutive
  // pixels of a row, so the swizzled 16-byte window reads stay conflict-free and the stores 512-byte runs.
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (wid & 3) * 8 + (lane & 7), ly = (wid >> 2) * 4 + (lane >> 3);
  const int px = tx * kTW + lx, py = ty * kTH + ly;
  const bool live = px < W && py < H;
  const int64_t pix = ((int64_t)b * H + py) * W + px;
  const T* gimg = g + (int64_t)b * H * W * C;
  // bounding box of the contributing output pixels
  const int bx0 = -bb[0], bx1 = bb[1], by0 = -bb[2], by1 = bb[3];
  const bool empty = bb[1] < 0 || bb[3] < 0;                 // still at the memset value: nothing reaches the tile
  const int nwx = empty ? 0 : (bx1 - bx0 + kBW) / kBW, nwy = empty ? 0 : (by1 - by0 + kBH) / kBH;
  if (nwx * nwy > kMaxWindows) {                             // rough flow: the scatter kernels redo the tensor
    if (threadIdx.x == 0) atomicExch(flag, 1);
    return;
  }
  const float sxf = (float)px, syf = (float)py;
  for (int c0 = 0; c0 < C; c0 += CC) {
    float acc[CC];
#pragma unroll
    for (int i = 0; i < CC; ++i) acc[i] = 0.f;
    for (int wi = 0; wi < nwx * nwy; ++wi) {
      const int wx0 = bx0 + (wi % nwx) * kBW, wy0 = by0 + (wi / nwx) * kBH;
      __syncthreads();                                       // previous window fully consumed
      if constexpr (TMA) {
        if (threadIdx.x == 0) tmaw::load(gwin, &tmg, bar, kBW * kBH * kPixB, c0, wx0, wy0, b);
      } else {
        load_window<T, kBW, kBH, kPixB, kTileThreads, true>(gwin, gimg, H, W, C, c0, wy0, wx0);
      }
      if (c0 == 0 || nwx * nwy > 1) {
        // sample positions of the window's output pixels + their integer displacement range
        if (threadIdx.x < 4) lb[threadIdx.x] = INT_MIN;
        __syncthreads();
        int m0 = INT_MIN, m1 = INT_MIN, m2 = INT_MIN, m3 = INT_MIN;
        for (int i = threadIdx.x; i < kBW * kBH; i += kTileThreads) {
          const int q = i % kBW, r = i / kBW;
          const int yy = wy0 + r, xx = wx0 + q;
          float2 c = make_float2(-1e8f, -1e8f);
          if (yy >= 0 && yy < H && xx >= 0 && xx < W && xx <= bx1 && yy <= by1) {
            const PixCoord pc = source_index(flow, ((int64_t)b * H + yy) * W + xx, yy, xx, H, W, scale);
            // only samples that can reach this tile take part in the candidate range
            if (pc.ix > (float)(tx * kTW - 2) && pc.ix < (float)(tx * kTW + kTW + 1) &&
                pc.iy > (float)(ty * kTH - 2) && pc.iy < (float)(ty * kTH + kTH + 1)) {
              c = make_float2(pc.ix, pc.iy);
              const int ex = pc.x0 - xx, ey = pc.y0 - yy;
              m0 = max(m0, -ex); m1 = max(m1, ex); m2 = max(m2, -ey); m3 = max(m3, ey);
            }
          }
          coord[i] = c;
        }
        m0 = __reduce_max_sync(0xffffffffu, m0); m1 = __reduce_max_sync(0xffffffffu, m1);
        m2 = __reduce_max_sync(0xffffffffu, m2); m3 = __reduce_max_sync(0xffffffffu, m3);
        if (lane == 0) { atomicMax(lb + 0, m0); atomicMax(lb + 1, m1); atomicMax(lb + 2, m2); atomicMax(lb + 3, m3); }
      }
      if constexpr (TMA) {
        tmaw::wait(bar, phase);
        phase ^= 1;
      } else {
        cp_async_wait_all();
      }
      __syncthreads();
      if (!live || lb[1] == INT_MIN) continue;               // no sample of this window reaches the tile
      // contributors of s: p in [s - e_max - 2, s - e_min + 1], clipped to the window
      const int qx0 = max(px - lb[1] - 2 - wx0, 0), qx1 = min(px + lb[0] + 1 - wx0, kBW - 1);
      const int qy0 = max(py - lb[3] - 2 - wy0, 0), qy1 = min(py + lb[2] + 1 - wy0, kBH - 1);
      for (int qy = qy0; qy <= qy1; ++qy) {
        for (int qx = qx0; qx <= qx1; ++qx) {
          const int qi = qy * kBW + qx;
          const float2 c = coord[qi];
          const float ddx = fabsf(sxf - c.x), ddy = fabsf(syf - c.y);
          if (ddx < 2.f && ddy < 2.f) fma_chunk_win<T>(gwin, qi, cubic_k(ddx) * cubic_k(ddy), acc);
        }
      }
    }
    if (live) store_chunk<T>(dx + pix * C + c0, acc);
  }
}

__global__ void __launch_bounds__(256)
zero_unless_tiled_kernel(float4* __restrict__ p, int64_t n4, const int* __restrict__ flag) {
  if (flag && tiled_ok(flag)) return;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(256)
cast_unless_tiled_kernel(const float4* __restrict__ in, bf16* __restrict__ out, int64_t n4, const int* __restrict__ flag) {
  if (flag && tiled_ok(flag)) return;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = in[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), c = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&c));
  }
}

inline bool tile_eligible(int dt, int H, int W, int C) {
  const int cc = dt == LCGAN_BF16 ? 32 : 16;
  return C % cc == 0 && W >= kTW && H >= kTH && getenv("LCGAN_WARP_NO_TILE") == nullptr;
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

template <typename K>
static int opt_in_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : 1;
}

static int tile_kernels_ready() {
  static DeviceOnce once;
  return once.run([] {
    int err = 0;
    err |= opt_in_smem(warp_tile_gather_kernel<bf16, false>, kFwdSmem);
    err |= opt_in_smem(warp_tile_gather_kernel<bf16, true>, kFwdSmem);
    err |= opt_in_smem(warp_tile_gather_kernel<bf16, false, true>, kFwdSmem);
    err |= opt_in_smem(warp_tile_gather_kernel<bf16, true, true>, kFwdSmem);
    err |= opt_in_smem(warp_tile_gather_kernel<bf16, false, true, 8>, kFwdSmem8);
    err |= opt_in_smem(warp_tile_gather_kernel<bf16, true, true, 8>, kFwdSmem8);
    err |= opt_in_smem(warp_tile_dx_kernel<bf16, true>, kDxSmem);
    err |= opt_in_smem(warp_tile_gather_kernel<float, false>, kFwdSmem);
    err |= opt_in_smem(warp_tile_gather_kernel<float, true>, kFwdSmem);
    err |= opt_in_smem(warp_tile_dx_kernel<bf16>, kDxSmem);
    err |= opt_in_smem(warp_tile_dx_kernel<float>, kDxSmem);
    return err;
  });
}

static bool tile8() { static const bool on = getenv("LCGAN_WARP_TH16") == nullptr; return on; }
static bool use_tma() { static const bool on = getenv("LCGAN_WARP_NO_TMA") == nullptr; return on; }

static void launch_tile_fwd(const void* x, const float* flow, void* out, const float* cs, int dt, int N, int H, int W,
                            int C, float flow_scale, int tiles, cudaStream_t s) {
  CUtensorMap tm{};
  if (dt == LCGAN_BF16 && use_tma() && tile8() && tmaw::make_map(&tm, x, N, H, W, C, kFW, 20, true))
    warp_tile_gather_kernel<bf16, false, true, 8><<<N * ((H + 7) / 8) * ((W + kTW - 1) / kTW), kTW * 8, kFwdSmem8, s>>>(
        tm, (const bf16*)x, flow, nullptr, (bf16*)out, nullptr, H, W, C, flow_scale, cs);
  else if (dt == LCGAN_BF16 && use_tma() && tmaw::make_map(&tm, x, N, H, W, C, kFW, kFH, true))
    warp_tile_gather_kernel<bf16, false, true><<<tiles, kTileThreads, kFwdSmem, s>>>(
        tm, (const bf16*)x, flow, nullptr, (bf16*)out, nullptr, H, W, C, flow_scale, cs);
  else if (dt == LCGAN_BF16)
    warp_tile_gather_kernel<bf16, false><<<tiles, kTileThreads, kFwdSmem, s>>>(
        tm, (const bf16*)x, flow, nullptr, (bf16*)out, nullptr, H, W, C, flow_scale, cs);
  else
    warp_tile_gather_kernel<float, false><<<tiles, kTileThreads, kFwdSmem, s>>>(
        tm, (const float*)x, flow, nullptr, (float*)out, nullptr, H, W, C, flow_scale, cs);
}

extern "C" int lcgan_warp_fwd(const void* x, const float* flow, void* out, int dt, int N, int H, int W, int C,
                              float flow_scale, void* stream) {
  LCGAN_CHECK(x && flow && out && N > 0 && H > 1 && W > 1 && C > 0, "warp_fwd: bad arguments");
  LCGAN_CHECK(dt == LCGAN_F32 || dt == LCGAN_BF16, "warp_fwd: bad dtype %d", dt);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t npix = (int64_t)N * H * W;
  if (tile_eligible(dt, H, W, C)) {
    LCGAN_CHECK(tile_kernels_ready() == 0, "warp_fwd: cannot opt in to %d bytes of shared memory", kDxSmem);
    const int tiles = N * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW);
    launch_tile_fwd(x, flow, out, nullptr, dt, N, H, W, C, flow_scale, tiles, s);
    LCGAN_LAUNCH_CHECK();
    return 0;
  }
#define CALL(T, V)                                                                         \
  do {                                                                                     \
    const int G = group_size(C / V);                                                       \
    warp_fwd_kernel<T, V><<<grid_for_groups(npix, G), kThreads, 0, s>>>(                   \
        (const T*)x, flow, (T*)out, N, H, W, C, flow_scale, G);                            \
  } while (0)
  if (dt == LCGAN_F32) { if (C % 4 == 0) CALL(float, 4); else CALL(float, 1); }
  else { if (C % 8 == 0) CALL(bf16, 8); else CALL(bf16, 1); }
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// Forward warp with a per-(image, channel) output scale (tiled shapes only: W >= 32, H >= 16, C % 32 (bf16) / 16).
extern "C" int lcgan_warp_fwd_cs(const void* x, const float* flow, void* out, const float* cs, int dt, int N, int H,
                                 int W, int C, float flow_scale, void* stream) {
  LCGAN_CHECK(x && flow && out && cs && N > 0 && H > 1 && W > 1 && C > 0, "warp_fwd_cs: bad arguments");
  LCGAN_CHECK((dt == LCGAN_F32 || dt == LCGAN_BF16) && tile_eligible(dt, H, W, C) && C % 4 == 0,
              "warp_fwd_cs: shape not eligible for the tiled kernel (use lcgan_warp_fwd + lcgan_modulate)");
  LCGAN_CHECK(tile_kernels_ready() == 0, "warp_fwd_cs: cannot opt in to %d bytes of shared memory", kDxSmem);
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles = N * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW);
  launch_tile_fwd(x, flow, out, cs, dt, N, H, W, C, flow_scale, tiles, s);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// the scatter kernels: dx_acc (f32, zeroed) += ..., dflow written; skip = device bounds (or null)
static void launch_atomic_bwd(const void* x, const float* flow, const void* dout, float* dx_acc, float* dflow, int dt,
                              int N, int H, int W, int C, float flow_scale, const int* skip, cudaStream_t s) {
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
          (const T*)x, flow, (const T*)dout, dx_acc, dflow, N, H, W, C, flow_scale, G4, run, skip); \
    } else {                                                                               \
      warp_bwd_kernel<T, V><<<grid_for_groups(npix, G), kThreads, 0, s>>>(                 \
          (const T*)x, flow, (const T*)dout, dx_acc, dflow, N, H, W, C, flow_scale, G, skip); \
    }                                                                                      \
  } while (0)
  if (dt == LCGAN_F32) { if (C % 4 == 0) CALL(float, 4); else CALL(float, 1); }
  else { if (C % 8 == 0) CALL(bf16, 8); else CALL(bf16, 1); }
#undef CALL
}

extern "C" int lcgan_warp_bwd(const void* x, const float* flow, const void* dout, float* dx_acc, float* dflow,
                              int dt, int N, int H, int W, int C, float flow_scale, void* stream) {
  LCGAN_CHECK(x && flow && dout && dx_acc && dflow && N > 0 && H > 1 && W > 1 && C > 0, "warp_bwd: bad arguments");
  LCGAN_CHECK(dt == LCGAN_F32 || dt == LCGAN_BF16, "warp_bwd: bad dtype %d", dt);
  launch_atomic_bwd(x, flow, dout, dx_acc, dflow, dt, N, H, W, C, flow_scale, nullptr, (cudaStream_t)stream);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_warp_bwd_tiled(const void* x, const float* flow, const void* dout, void* dx, float* dflow,
                                    float* ws_acc, int* ws_bounds, int dt, int N, int H, int W, int C,
                                    float flow_scale, void* stream) {
  LCGAN_CHECK(x && flow && dout && dx && dflow && ws_bounds && N > 0 && H > 1 && W > 1 && C > 0,
              "warp_bwd_tiled: bad arguments");
  LCGAN_CHECK(dt == LCGAN_F32 || dt == LCGAN_BF16, "warp_bwd_tiled: bad dtype %d", dt);
  LCGAN_CHECK(dt == LCGAN_F32 || ws_acc, "warp_bwd_tiled: bf16 needs the f32 scratch accumulator");
  LCGAN_CHECK(((int64_t)N * H * W * C) % 4 == 0, "warp_bwd_tiled: element count must be a multiple of 4");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n4 = (int64_t)N * H * W * C / 4;
  float* acc = dt == LCGAN_F32 ? (float*)dx : ws_acc;     // fp32: the scatter kernels accumulate in place
  const int egrid = (int)(n4 / 256 + 1 < 148LL * 16 ? n4 / 256 + 1 : 148LL * 16);
  const int* skip = nullptr;
  if (tile_eligible(dt, H, W, C)) {
    LCGAN_CHECK(tile_kernels_ready() == 0, "warp_bwd_tiled: cannot opt in to %d bytes of shared memory", kDxSmem);
    const int tiles = N * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW);
    int* bbox = ws_bounds;                                // [tiles][4]
    int* flag = ws_bounds + (size_t)tiles * 4;            // raised by the dx kernel when a tile's box is too large
    LCGAN_CUDA(cudaMemsetAsync(bbox, 0x80, (size_t)tiles * 4 * sizeof(int), s));
    LCGAN_CUDA(cudaMemsetAsync(flag, 0, 4 * sizeof(int), s));
    const int64_t npix = (int64_t)N * H * W;
    const int bgrid = (int)(npix / 256 + 1 < 148LL * 8 ? npix / 256 + 1 : 148LL * 8);
    CUtensorMap tmx{}, tmg{};
    const bool tma = dt == LCGAN_BF16 && use_tma() && tmaw::make_map(&tmx, x, N, H, W, C, kFW, tile8() ? 20 : kFH, true) &&
                     tmaw::make_map(&tmg, dout, N, H, W, C, kBW, kBH, true);
    // the contributor boxes come from the flow-gradient pass on the TMA path, from a pre-pass otherwise
    if (!tma) warp_bbox_kernel<<<bgrid, 256, 0, s>>>(flow, N, H, W, flow_scale, bbox);
    if (tma) {
      if (tile8())
        warp_tile_gather_kernel<bf16, true, true, 8><<<N * ((H + 7) / 8) * ((W + kTW - 1) / kTW), kTW * 8, kFwdSmem8, s>>>(
            tmx, (const bf16*)x, flow, (const bf16*)dout, nullptr, dflow, H, W, C, flow_scale, nullptr, bbox);
      else
        warp_tile_gather_kernel<bf16, true, true><<<tiles, kTileThreads, kFwdSmem, s>>>(
            tmx, (const bf16*)x, flow, (const bf16*)dout, nullptr, dflow, H, W, C, flow_scale, nullptr, bbox);
      warp_tile_dx_kernel<bf16, true><<<tiles, kTileThreads, kDxSmem, s>>>(tmg, flow, (const bf16*)dout, (bf16*)dx, bbox,
                                                                          flag, H, W, C, flow_scale);
    } else if (dt == LCGAN_BF16) {
      warp_tile_gather_kernel<bf16, true><<<tiles, kTileThreads, kFwdSmem, s>>>(
          tmx, (const bf16*)x, flow, (const bf16*)dout, nullptr, dflow, H, W, C, flow_scale);
      warp_tile_dx_kernel<bf16><<<tiles, kTileThreads, kDxSmem, s>>>(tmg, flow, (const bf16*)dout, (bf16*)dx, bbox, flag,
                                                                    H, W, C, flow_scale);
    } else {
      warp_tile_gather_kernel<float, true><<<tiles, kTileThreads, kFwdSmem, s>>>(
          tmx, (const float*)x, flow, (const float*)dout, nullptr, dflow, H, W, C, flow_scale);
      warp_tile_dx_kernel<float><<<tiles, kTileThreads, kDxSmem, s>>>(tmg, flow, (const float*)dout, (float*)dx, bbox,
                                                                     flag, H, W, C, flow_scale);
    }
    LCGAN_LAUNCH_CHECK();
    skip = flag;                                          // the rest runs only if a tile gave up
  }
  if (!skip && lcgan_det_enabled()) {
    // deterministic mode, image too small for the tiles: flow gradient from the per-pixel kernel (no dx),
    // dx from the window gather (no atomics)
    const int Rx = (int)(2.5f + fabsf(flow_scale) * 0.5f * (float)W) + 1, Ry = (int)(2.5f + fabsf(flow_scale) * 0.5f * (float)H) + 1;
    const int64_t npix = (int64_t)N * H * W;
#define CALLD(T, V)                                                                                             \
  do {                                                                                                          \
    const int G = group_size(C / V);                                                                            \
    warp_bwd_kernel<T, V><<<grid_for_groups(npix, G), kThreads, 0, s>>>(                                        \
        (const T*)x, flow, (const T*)dout, nullptr, dflow, N, H, W, C, flow_scale, G, nullptr);                 \
    warp_dx_window_kernel<T, V><<<grid_for_groups(npix, G), kThreads, 0, s>>>(                                  \
        flow, (const T*)dout, (T*)dx, N, H, W, C, flow_scale, G, Rx, Ry);                                       \
  } while (0)
    if (dt == LCGAN_F32) { if (C % 4 == 0) CALLD(float, 4); else CALLD(float, 1); }
    else { if (C % 8 == 0) CALLD(bf16, 8); else CALLD(bf16, 1); }
#undef CALLD
    LCGAN_LAUNCH_CHECK();
    return 0;
  }
  zero_unless_tiled_kernel<<<egrid, 256, 0, s>>>(reinterpret_cast<float4*>(acc), n4, skip);
  launch_atomic_bwd(x, flow, dout, acc, dflow, dt, N, H, W, C, flow_scale, skip, s);
  if (dt == LCGAN_BF16)
    cast_unless_tiled_kernel<<<egrid, 256, 0, s>>>(reinterpret_cast<const float4*>(acc), (bf16*)dx, n4, skip);
  LCGAN_LAUNCH_CHECK();
  return 0;
}
