// Memory-bound kernels on dense channels-last tensors: box filter (+fused activation / mask),
// 2x2 pool, nearest x2, skip-path upsample, epilogue backward, style modulation, casts.
// One thread handles one 16-byte channel vector of one pixel (V=1 fallback for odd C), so a warp
// reads/writes contiguous 512-byte runs along the channel-innermost axis.
#include "common.cuh"
#include "tma_window.cuh"
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
__device__ int g_sems[kDetSems];     // deterministic-mode turn semaphores (common.cuh)
constexpr int kStreamStages = 4;     // cp.async depth of the streaming reductions (act_bwd, modulate_bwd)

template <typename T, int V>
__device__ __forceinline__ void ldv(const T* p, float* f) {
  if constexpr (V == 1) {
    f[0] = ldf(p);
  } else {
    Vec16<T> v; v.load(p); v.unpack(f);
  }
}
template <typename T, int V>
__device__ __forceinline__ void stv(T* p, const float* f) {
  if constexpr (V == 1) {
    stf(p, f[0]);
  } else {
    Vec16<T> v; v.pack(f); v.store(p);
  }
}

inline int grid_for(int64_t n) {
  int64_t b = (n + kThreads - 1) / kThreads;
  static const int per_sm = getenv("LCGAN_GRID_CAP") ? atoi(getenv("LCGAN_GRID_CAP")) : 64;
  const int64_t cap = 148LL * per_sm;   // grid-stride beyond 64 CTAs per SM
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
box3_kernel(const T* __restrict__ a, const T* __restrict__ mask, T* __restrict__ out, int N, int H,
            int W, int C, float pre_slope, float pre_gain, float post_slope, float post_gain) {
  const int cv = C / V;
  const int64_t total = (int64_t)N * H * W * cv;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(idx % cv) * V;
    int64_t p = idx / cv;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        const int64_t off = (((int64_t)b * H + yy) * W + xx) * C + c;
        float f[V];
        ldv<T, V>(a + off, f);
        if (mask) {
          float m[V];
          ldv<T, V>(mask + off, m);
#pragma unroll
          for (int i = 0; i < V; ++i) f[i] *= (m[i] > 0.f ? pre_gain : pre_gain * pre_slope);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += f[i];
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float v = acc[i] * (1.f / 9.f);
      acc[i] = (v > 0.f ? v : v * post_slope) * post_gain;
    }
    stv<T, V>(out + idx * V, acc);
  }
}

// Strip variant: one thread owns a column of R output rows for one channel vector and slides a
// window of horizontal 3-sums down it, so each input vector is loaded 3(R+2)/R times instead of 9.
template <typename T, int V, int R, bool MASK>
__global__ void __launch_bounds__(kThreads)
box3_strip_kernel(const T* __restrict__ a, const T* __restrict__ mask, T* __restrict__ out, int N, int H,
                  int W, int C, float pre_slope, float pre_gain, float post_slope, float post_gain) {
  const int cv = C / V, HS = H / R;
  const int64_t total = (int64_t)N * HS * W * cv;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(idx % cv) * V;
    int64_t p = idx / cv;
    const int x = (int)(p % W); p /= W;
    const int y0 = (int)(p % HS) * R;
    const int b = (int)(p / HS);
    const int64_t img = (int64_t)b * H * W * C;
    // all (R+2)*3 vector loads are issued before any is consumed (memory-level parallelism), then the
    // horizontal 3-sums h[r] of rows y0-1 .. y0+R are combined vertically
    float h[R + 2][V];
#pragma unroll
    for (int r = 0; r < R + 2; ++r) {
      const int yy = y0 - 1 + r;
      const bool oky = yy >= 0 && yy < H;
      const int yc = min(max(yy, 0), H - 1);
      float f[3][V];
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        const bool ok = oky && xx >= 0 && xx < W;
        const int xc = min(max(xx, 0), W - 1);
        const int64_t off = img + ((int64_t)yc * W + xc) * C + c;
        ldv<T, V>(a + off, f[dx + 1]);
        if constexpr (MASK) {
          float m[V];
          ldv<T, V>(mask + off, m);
#pragma unroll
          for (int i = 0; i < V; ++i) f[dx + 1][i] *= (m[i] > 0.f ? pre_gain : pre_gain * pre_slope);
        }
        if (!ok) {
#pragma unroll
          for (int i = 0; i < V; ++i) f[dx + 1][i] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) h[r][i] = f[0][i] + f[1][i] + f[2][i];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float v = (h[r][i] + h[r + 1][i] + h[r + 2][i]) * (1.f / 9.f);
        o[i] = (v > 0.f ? v : v * post_slope) * post_gain;
      }
      stv<T, V>(out + img + ((int64_t)(y0 + r) * W + x) * C + c, o);
    }
  }
}


// Tiled variant (no mask): a CTA stages a 34x18 pixel window of one 64-byte channel chunk in shared
// memory with cp.async (zero fill = the box filter's zero padding) and each thread slides a window
// of horizontal 3-sums down 8 rows of one 16-byte channel vector.  The loads need no registers, so
// enough bytes are in flight regardless of instruction scheduling (the strip kernel's 30 dependent
// vector loads per thread left it latency-bound at 2.2 TB/s).
constexpr int kBoxTW = 32, kBoxTH = 16, kBoxWW = kBoxTW + 2;   // tile height: 16 (8 with the mask window, 48 KiB static limit)
// TMA = true (bf16): the window(s) arrive by one bulk tensor copy each (tma_window.cuh; out-of-image pixels zero-filled by
// the unit) instead of ~10 bounds-checked 16-byte cp.async per thread - about 40 % of this kernel's instructions.
template <typename T, bool MASK, int TH, bool TMA = false>
__global__ void __launch_bounds__(kThreads)
box3_tile_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmm,
                 const T* __restrict__ a, const T* __restrict__ mask, T* __restrict__ out, int H, int W, int C,
                 float pre_slope, float pre_gain, float post_slope, float post_gain,
                 const float* __restrict__ cs = nullptr, float* __restrict__ red = nullptr,
                 const T* __restrict__ post_mask = nullptr) {
  // post_mask (MASK == false only): out = box3(a) * (post_mask > 0 ? post_gain : post_gain * post_slope) - the box
  // filter's backward followed by the backward of the leaky-relu of the conv BEFORE it (y = post_mask), one pass;
  // red[b,c] += sum_p out then is that conv's bias gradient
  // cs [N,C] (optional): per-(image, channel) scale - the style modulation of the NEXT layer folded into this
  // pass: forward out = post(box3(a)) * cs; backward (MASK) pre(a) = a * cs * lrelu'(mask * cs), and
  // red[b,c] += sum over the tile's own pixels of a * mask (the style gradient, divided by cs on the host)
  constexpr int E = 16 / sizeof(T);
  constexpr int kBoxWH = TH + 2, SR = TH / 2;               // window height, rows per thread
  constexpr int kWin = kBoxWW * kBoxWH * 64;
  __shared__ __align__(128) unsigned char sm[MASK ? 2 * kWin : kWin];
  __shared__ __align__(8) uint64_t bar;
  const int tiles_x = (W + kBoxTW - 1) / kBoxTW, tiles_y = (H + TH - 1) / TH;
  // the channel chunks of a tile are neighbouring CTAs (chunk fastest): they run together and read
  // whole DRAM pages; chunk-major order swept the tensor once per chunk, 64 bytes of every pixel
  const int nchunk = C / (64 / (int)sizeof(T));
  int t = blockIdx.x / nchunk;
  const int c0 = (blockIdx.x % nchunk) * (64 / (int)sizeof(T));
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  if constexpr (TMA) {
    static_assert(kWin % 128 == 0, "the mask window must start 128-byte aligned");
    if (threadIdx.x == 0) {
      tmaw::bar_init(&bar);
      tmaw::arm(&bar, MASK ? 2 * kWin : kWin);
      tmaw::copy(sm, &tma, &bar, c0, tx * kBoxTW - 1, ty * TH - 1, b);
      if constexpr (MASK) tmaw::copy(sm + kWin, &tmm, &bar, c0, tx * kBoxTW - 1, ty * TH - 1, b);
    }
    __syncthreads();                                          // the barrier is initialised
    tmaw::wait(&bar, 0);
  } else {
    load_window<T, kBoxWW, kBoxWH, 64, kThreads>(sm, a + (int64_t)b * H * W * C, H, W, C, c0, ty * TH - 1, tx * kBoxTW - 1);
    if constexpr (MASK)   // the activation mask of the backward pass: a is scaled by the leaky-relu slope of mask
      load_window<T, kBoxWW, kBoxWH, 64, kThreads>(sm + kWin, mask + (int64_t)b * H * W * C, H, W, C, c0,
                                                   ty * TH - 1, tx * kBoxTW - 1);
    cp_async_wait_all();
    __syncthreads();
  }
  const int v = threadIdx.x & 3, xq = (threadIdx.x >> 2) & 31, strip = threadIdx.x >> 7;
  const int ox = tx * kBoxTW + xq;
  float csv[E], racc[E];
#pragma unroll
  for (int i = 0; i < E; ++i) { csv[i] = cs ? cs[(int64_t)b * C + c0 + v * E + i] : 1.f; racc[i] = 0.f; }
  const bool col_live = ox < W;
  if (!col_live && !red) return;                            // (with a reduction every thread reaches the barrier below)
  const float neg = pre_gain * pre_slope;
  float h0[E], h1[E], h2[E];
#pragma unroll
  for (int rr = 0; rr < SR + 2; ++rr) {
    const unsigned char* p = sm + ((strip * SR + rr) * kBoxWW + xq) * 64 + v * 16;
    float f[3][E];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Vec16<T> u;
      u.v = *reinterpret_cast<const decltype(u.v)*>(p + 64 * k);
      u.unpack(f[k]);
      if constexpr (MASK) {
        Vec16<T> m;
        float mf[E];
        m.v = *reinterpret_cast<const decltype(m.v)*>(p + kWin + 64 * k);
        m.unpack(mf);
        if (red && k == 1 && rr >= 1 && rr <= SR && col_live) {      // the tile's own pixels: centre column, own rows
#pragma unroll
          for (int i = 0; i < E; ++i) racc[i] = fmaf(f[k][i], mf[i], racc[i]);
        }
#pragma unroll
        for (int i = 0; i < E; ++i) f[k][i] *= csv[i] * (mf[i] * csv[i] > 0.f ? pre_gain : neg);
      }
    }
#pragma unroll
    for (int i = 0; i < E; ++i) { h0[i] = h1[i]; h1[i] = h2[i]; h2[i] = f[0][i] + f[1][i] + f[2][i]; }
    if (rr >= 2) {
      const int oy = ty * TH + strip * SR + rr - 2;
      if (oy < H && col_live) {
        float o[E];
        const int64_t oidx = (((int64_t)b * H + oy) * W + ox) * C + c0 + v * E;
        if (!MASK && post_mask) {
          Vec16<T> m;
          float mf[E];
          m.load(post_mask + oidx);
          m.unpack(mf);
#pragma unroll
          for (int i = 0; i < E; ++i) {
            o[i] = (h0[i] + h1[i] + h2[i]) * (1.f / 9.f) * (mf[i] > 0.f ? post_gain : post_gain * post_slope);
            racc[i] += o[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < E; ++i) {
            const float sv = (h0[i] + h1[i] + h2[i]) * (1.f / 9.f);
            o[i] = (sv > 0.f ? sv : sv * post_slope) * post_gain;
            if constexpr (!MASK) o[i] *= csv[i];
          }
        }
        Vec16<T> w;
        w.pack(o);
        w.store(out + (((int64_t)b * H + oy) * W + ox) * C + c0 + v * E);
      }
    }
  }
  {
    if (red) {
      // rows of the window beyond the image contribute zeros (zero-filled window), columns beyond it were skipped;
      // sum the 64 threads (32 columns x 2 strips) that share a channel vector, one atomic per channel and CTA
      __syncthreads();                                        // the windows are no longer read: reuse them
      float* rs = reinterpret_cast<float*>(sm);
#pragma unroll
      for (int i = 0; i < E; ++i) rs[threadIdx.x * E + i] = racc[i];
      __syncthreads();
      if (threadIdx.x < 4 * E) {
        const int vv = threadIdx.x / E, i = threadIdx.x % E;
        float t = 0.f;
        for (int j = 0; j < kThreads / 4; ++j) t += rs[(j * 4 + vv) * E + i];
        atomicAdd(red + (int64_t)b * C + c0 + vv * E + i, t);
      }
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
pool2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, float scale) {
  const int cv = C / V, OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)N * OH * OW * cv;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(idx % cv) * V;
    int64_t p = idx / cv;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const int b = (int)(p / OH);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[V];
        ldv<T, V>(x + (((int64_t)b * H + 2 * oy + dy) * W + 2 * ox + dx) * C + c, f);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += f[i];
      }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= scale;
    stv<T, V>(y + idx * V, acc);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
up2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, float scale) {
  const int cv = C / V, OH = H * 2, OW = W * 2;
  const int64_t total = (int64_t)N * OH * OW * cv;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(idx % cv) * V;
    int64_t p = idx / cv;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const int b = (int)(p / OH);
    float f[V];
    ldv<T, V>(x + (((int64_t)b * H + oy / 2) * W + ox / 2) * C + c, f);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] *= scale;
    stv<T, V>(y + idx * V, f);
  }
}

// out = a + scale * nearest_up2(s) ; s [N,H,W,C], a / out [N,2H,2W,C]: the gradient of a tensor that feeds both a layer
// (gradient a) and a 2x2 average pool (gradient s) in one pass - autograd's own route is up2 (write N) + add (read 2N,
// write N).  IDX = uint32_t when the vector count fits (64-bit div / mod dominate these index-decoding kernels).
template <typename T, int V, typename IDX>
__global__ void __launch_bounds__(kThreads)
up2_add_kernel(const T* __restrict__ a, const T* __restrict__ s, T* __restrict__ out, int N, int H, int W, int C,
               float scale) {
  const IDX cv = C / V, OH = H * 2, OW = W * 2;
  const IDX total = (IDX)N * OH * OW * cv;
  for (IDX idx = blockIdx.x * (IDX)kThreads + threadIdx.x; idx < total; idx += (IDX)gridDim.x * kThreads) {
    const IDX c = (idx % cv) * V;
    IDX p = idx / cv;
    const IDX ox = p % OW; p /= OW;
    const IDX oy = p % OH;
    const IDX b = p / OH;
    float f[V], g[V];
    ldv<T, V>(a + (int64_t)idx * V, f);
    ldv<T, V>(s + (((int64_t)b * H + oy / 2) * W + ox / 2) * C + c, g);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = fmaf(g[i], scale, f[i]);
    stv<T, V>(out + (int64_t)idx * V, f);
  }
}

// out[y, x] = (1/9) sum_{i,j = -1..2} w_i w_j g[2y+i, 2x+j], w = (1,2,2,1), zeros outside: sum-pool2(box3(g)), the
// gradient of box3(nearest_up2(.)), in one pass (box3 + pool2 separately: read N, write N, read N, write N/4)
template <typename T, int V, typename IDX>
__global__ void __launch_bounds__(kThreads)
box3_pool2_kernel(const T* __restrict__ g, T* __restrict__ out, int N, int H, int W, int C) {
  const IDX cv = C / V, OH = H / 2, OW = W / 2;
  const IDX total = (IDX)N * OH * OW * cv;
  for (IDX idx = blockIdx.x * (IDX)kThreads + threadIdx.x; idx < total; idx += (IDX)gridDim.x * kThreads) {
    const IDX c = (idx % cv) * V;
    IDX p = idx / cv;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const IDX b = p / OH;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = -1; i <= 2; ++i) {
      const int y = 2 * oy + i;
      const float wy = (i == 0 || i == 1) ? 2.f : 1.f;
#pragma unroll
      for (int j = -1; j <= 2; ++j) {
        const int x = 2 * ox + j;
        const float wgt = wy * ((j == 0 || j == 1) ? 2.f : 1.f);
        if (y >= 0 && y < H && x >= 0 && x < W) {
          float f[V];
          ldv<T, V>(g + (((int64_t)b * H + y) * W + x) * C + c, f);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = fmaf(f[k], wgt, acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] *= (1.f / 9.f);
    stv<T, V>(out + (int64_t)idx * V, acc);
  }
}

// out = box3(nearest_up2(s)) + t ; s [N,H,W,C] -> out [N,2H,2W,C]
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
up2box_add_kernel(const T* __restrict__ s, const T* __restrict__ t, T* __restrict__ out, int N, int H,
                  int W, int C) {
  const int cv = C / V, OH = H * 2, OW = W * 2;
  const int64_t total = (int64_t)N * OH * OW * cv;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(idx % cv) * V;
    int64_t p = idx / cv;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const int b = (int)(p / OH);
    // the 3x3 window on the x2 grid covers at most 2 low-res rows/cols with weights (2,1) or (1,2);
    // branch-free: out-of-range rows/cols are clamped and weigh 0, so the five loads issue together
    const int y0 = (oy - 1) >> 1, y1 = (oy + 1) >> 1;   // oy-1 may be -1 -> y0 = -1 (arith shift)
    const int x0 = (ox - 1) >> 1, x1 = (ox + 1) >> 1;
    const float wy[2] = {y0 >= 0 ? ((oy & 1) ? 2.f : 1.f) : 0.f, y1 < H ? ((oy & 1) ? 1.f : 2.f) : 0.f};
    const float wx[2] = {x0 >= 0 ? ((ox & 1) ? 2.f : 1.f) : 0.f, x1 < W ? ((ox & 1) ? 1.f : 2.f) : 0.f};
    const int ys[2] = {max(y0, 0), min(y1, H - 1)};
    const int xs[2] = {max(x0, 0), min(x1, W - 1)};
    float acc[V], f[4][V];
    ldv<T, V>(t + idx * V, acc);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        ldv<T, V>(s + (((int64_t)b * H + ys[j]) * W + xs[i]) * C + c, f[j * 2 + i]);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float wgt = wy[j] * wx[i] * (1.f / 9.f);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += f[j * 2 + i][k] * wgt;
      }
    stv<T, V>(out + idx * V, acc);
  }
}

// The same with one thread per LOW-RES pixel vector and its 2 x 2 outputs: out(2y+py, 2x+px) =
// (1/9) sum_{a,b} wy[py][a] wx[px][b] s[y+a-1, x+b-1] with w[0] = (1,2,0), w[1] = (0,2,1) - horizontal sums
// L = s[x-1] + 2 s[x], R = 2 s[x] + s[x+1] per low-res row, then two vertical combinations each.  One index decode
// (32-bit) and 9 + 4 loads per four outputs instead of a 64-bit decode and 5 loads per output.
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
up2box_add_quad_kernel(const T* __restrict__ s, const T* __restrict__ t, T* __restrict__ out, int N, int H, int W, int C) {
  const uint32_t cv = C / V;
  const uint32_t total = (uint32_t)N * H * W * cv;
  const int OW = 2 * W;
  for (uint32_t idx = blockIdx.x * kThreads + threadIdx.x; idx < total; idx += gridDim.x * kThreads) {
    const uint32_t c = (idx % cv) * V;
    uint32_t p = idx / cv;
    const int x = (int)(p % (uint32_t)W); p /= (uint32_t)W;
    const int y = (int)(p % (uint32_t)H);
    const uint32_t b = p / (uint32_t)H;
    const T* sb = s + (size_t)b * H * W * C + c;
    float L[3][V], R[3][V];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int yy = y + a - 1;
      float m[V], l[V], r[V];
      const bool oky = yy >= 0 && yy < H;
      const T* row = sb + (size_t)(oky ? yy : y) * W * C;
      ldv<T, V>(row + (size_t)x * C, m);
      ldv<T, V>(row + (size_t)(x > 0 ? x - 1 : x) * C, l);
      ldv<T, V>(row + (size_t)(x + 1 < W ? x + 1 : x) * C, r);
      const float wl = (oky && x > 0) ? 1.f : 0.f, wr = (oky && x + 1 < W) ? 1.f : 0.f, wm = oky ? 2.f : 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        L[a][i] = fmaf(l[i], wl, m[i] * wm);
        R[a][i] = fmaf(r[i], wr, m[i] * wm);
      }
    }
    const size_t o00 = (((size_t)b * 2 * H + 2 * y) * OW + 2 * x) * C + c;
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const size_t o = o00 + ((size_t)py * OW + px) * C;
        float f[V];
        ldv<T, V>(t + o, f);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float v = px == 0 ? (py == 0 ? L[0][i] + 2.f * L[1][i] : 2.f * L[1][i] + L[2][i])
                                  : (py == 0 ? R[0][i] + 2.f * R[1][i] : 2.f * R[1][i] + R[2][i]);
          f[i] = fmaf(v, 1.f / 9.f, f[i]);
        }
        stv<T, V>(out + o, f);
      }
  }
}

// Epilogue backward with per-(b,c) reductions.  Block = (pixel chunk, b); threads along channels
// first so loads coalesce; each thread owns one channel vector, keeps V partial sums over its
// pixels and finishes with one atomic per (b,c).
template <typename T, int V, int STAGES>
__global__ void __launch_bounds__(kThreads)
act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ gout,
               const float* __restrict__ d, float* __restrict__ r0, float* __restrict__ r1, int P, int C,
               float slope, float gain, int pix_per_block, int* sems) {
  const int cv = C / V;
  const int b = blockIdx.y;
  (void)pix_per_block;                       // only sizes the grid: blocks interleave over the pixels
  const float inv_pos = 1.f / gain, inv_neg = 1.f / (gain * slope);
  for (int cg = 0; cg < cv; cg += kThreads) {
    const int ncv = min(kThreads, cv - cg);
    const int lanes = kThreads / ncv;          // pixels processed per block iteration
    const bool active = (int)threadIdx.x < lanes * ncv;
    const int c = (cg + (int)threadIdx.x % ncv) * V;
    float s0[V], s1[V], dd[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s0[i] = 0.f; s1[i] = 0.f; dd[i] = (d && active) ? d[(int64_t)b * C + c + i] : 1.f; }
    const int64_t base = (int64_t)b * P * C + c;
    int p = active ? blockIdx.x * lanes + threadIdx.x / ncv : P;   // inactive threads: empty pixel range
    const int pstep = gridDim.x * lanes;
    auto body = [&](float* g, const float* yy, int64_t off) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const bool pos = yy[i] > 0.f;
        const float dz = g[i] * (pos ? gain : gain * slope);
        s0[i] += dz;
        s1[i] += dz * yy[i] * (pos ? inv_pos : inv_neg);
        g[i] = dz * dd[i];
      }
      stv<T, V>(gout + off, g);
    };
    if constexpr (V > 1) {
      // STAGES pixels in flight per thread through cp.async (the loads hold no registers and
      // cannot be sunk to their consumers by the compiler, which left a plain unrolled loop at
      // 4.0 TB/s); every thread reads back only the slots it filled itself: no block barrier.
      __shared__ uint4 stage[STAGES][2][kThreads];
      int pl = p;
#pragma unroll
      for (int st = 0; st < STAGES; ++st, pl += pstep) {
        if (pl < P) {
          cp_async16(&stage[st][0][threadIdx.x], dy + base + (int64_t)pl * C, true);
          cp_async16(&stage[st][1][threadIdx.x], y + base + (int64_t)pl * C, true);
        }
        cp_async_commit();
      }
      for (int it = 0; p < P; p += pstep, pl += pstep, ++it) {
        cp_async_wait_pending<STAGES - 1>();
        const int st = it % STAGES;
        float g[V], yy[V];
        unpack_raw16<T>(stage[st][0][threadIdx.x], g);
        unpack_raw16<T>(stage[st][1][threadIdx.x], yy);
        if (pl < P) {
          cp_async16(&stage[st][0][threadIdx.x], dy + base + (int64_t)pl * C, true);
          cp_async16(&stage[st][1][threadIdx.x], y + base + (int64_t)pl * C, true);
        }
        cp_async_commit();
        body(g, yy, base + (int64_t)p * C);
      }
      cp_async_wait_pending<0>();
    } else {
      for (; p < P; p += pstep) {
        const int64_t off = base + (int64_t)p * C;
        float g[V], yy[V];
        ldv<T, V>(dy + off, g);
        ldv<T, V>(y + off, yy);
        body(g, yy, off);
      }
    }
    if (r0 || r1) {
      // fixed-order sum over the pixel lanes of the block through shared memory, then ONE atomic per
      // (b,c) and block.  (Per-thread atomics put gridDim.x * lanes ~ 10^4 adds on each of a few hundred
      // addresses - at batch 4 the L2 atomic units serialised them into 0.5 ms per launch.)
      // Deterministic mode: the blocks of an image add in blockIdx.x order instead (common.cuh).
      __shared__ float red[kThreads * V];
      const bool first = active && (int)threadIdx.x < ncv;
      auto lane_sum = [&](float* sv) {
        __syncthreads();
        if (active) {
#pragma unroll
          for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = sv[i];
        }
        __syncthreads();
        if (first)
          for (int l = 1; l < lanes; ++l)
#pragma unroll
            for (int i = 0; i < V; ++i) sv[i] += red[(l * ncv + threadIdx.x) * V + i];
      };
      if (r0) lane_sum(s0);
      if (r1) lane_sum(s1);
      if (sems) det_block_begin(sems + b, blockIdx.x);
      if (first) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if (sems) {
            if (r0) det_add(r0 + (int64_t)b * C + c + i, s0[i]);
            if (r1) det_add(r1 + (int64_t)b * C + c + i, s1[i]);
          } else {
            if (r0) atomicAdd(r0 + (int64_t)b * C + c + i, s0[i]);
            if (r1) atomicAdd(r1 + (int64_t)b * C + c + i, s1[i]);
          }
        }
      }
      if (sems) det_block_end(sems + b, blockIdx.x, gridDim.x);
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
modulate_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ xs, int N, int P, int C) {
  const int cv = C / V;
  const int64_t total = (int64_t)N * P * cv;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(idx % cv) * V;
    const int b = (int)(idx / ((int64_t)P * cv));
    float f[V];
    ldv<T, V>(x + idx * V, f);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] *= s[(int64_t)b * C + c + i];
    stv<T, V>(xs + idx * V, f);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
modulate_bwd_kernel(const T* __restrict__ x, const T* __restrict__ t, const float* __restrict__ s,
                    T* __restrict__ dx, float* __restrict__ ds, int P, int C, int pix_per_block, int* sems) {
  const int cv = C / V;
  const int b = blockIdx.y;
  (void)pix_per_block;                       // only sizes the grid: blocks interleave over the pixels
  for (int cg = 0; cg < cv; cg += kThreads) {
    const int ncv = min(kThreads, cv - cg);
    const int lanes = kThreads / ncv;
    const bool active = (int)threadIdx.x < lanes * ncv;
    const int c = (cg + (int)threadIdx.x % ncv) * V;
    float acc[V], ss[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { acc[i] = 0.f; ss[i] = active ? s[(int64_t)b * C + c + i] : 0.f; }
    const int64_t base = (int64_t)b * P * C + c;
    int p = active ? blockIdx.x * lanes + threadIdx.x / ncv : P;
    const int pstep = gridDim.x * lanes;
    if constexpr (V > 1) {
      __shared__ uint4 stage[kStreamStages][2][kThreads];   // cp.async pipeline, see act_bwd_kernel
      int pl = p;
#pragma unroll
      for (int st = 0; st < kStreamStages; ++st, pl += pstep) {
        if (pl < P) {
          cp_async16(&stage[st][0][threadIdx.x], x + base + (int64_t)pl * C, true);
          cp_async16(&stage[st][1][threadIdx.x], t + base + (int64_t)pl * C, true);
        }
        cp_async_commit();
      }
      for (int it = 0; p < P; p += pstep, pl += pstep, ++it) {
        cp_async_wait_pending<kStreamStages - 1>();
        const int st = it % kStreamStages;
        float xv[V], tv[V];
        unpack_raw16<T>(stage[st][0][threadIdx.x], xv);
        unpack_raw16<T>(stage[st][1][threadIdx.x], tv);
        if (pl < P) {
          cp_async16(&stage[st][0][threadIdx.x], x + base + (int64_t)pl * C, true);
          cp_async16(&stage[st][1][threadIdx.x], t + base + (int64_t)pl * C, true);
        }
        cp_async_commit();
#pragma unroll
        for (int i = 0; i < V; ++i) { acc[i] += xv[i] * tv[i]; tv[i] *= ss[i]; }
        stv<T, V>(dx + base + (int64_t)p * C, tv);
      }
      cp_async_wait_pending<0>();
    } else {
      for (; p < P; p += pstep) {
        const int64_t off = base + (int64_t)p * C;
        float xv[V], tv[V];
        ldv<T, V>(x + off, xv);
        ldv<T, V>(t + off, tv);
#pragma unroll
        for (int i = 0; i < V; ++i) { acc[i] += xv[i] * tv[i]; tv[i] *= ss[i]; }
        stv<T, V>(dx + off, tv);
      }
    }
    {
      // block-level sum over the pixel lanes, then one atomic per (b,c) and block (see act_bwd_kernel)
      __shared__ float red[kThreads * V];
      const bool first = active && (int)threadIdx.x < ncv;
      __syncthreads();
      if (active) {
#pragma unroll
        for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = acc[i];
      }
      __syncthreads();
      if (first)
        for (int l = 1; l < lanes; ++l)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += red[(l * ncv + threadIdx.x) * V + i];
      if (sems) det_block_begin(sems + b, blockIdx.x);
      if (first) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if (sems) det_add(ds + (int64_t)b * C + c + i, acc[i]);
          else atomicAdd(ds + (int64_t)b * C + c + i, acc[i]);
        }
      }
      if (sems) det_block_end(sems + b, blockIdx.x, gridDim.x);
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    stf(out + i, ldf(in + i));
}

template <typename T> constexpr int vec_of() { return Vec16<T>::N; }

// dispatch on dtype and on whether C admits 16-byte vectors
#define DISPATCH_TV(dt, C, CALL)                                          \
  do {                                                                    \
    if ((dt) == LCGAN_F32) {                                              \
      if ((C) % 4 == 0) { CALL(float, 4); } else { CALL(float, 1); }      \
    } else if ((dt) == LCGAN_BF16) {                                      \
      if ((C) % 8 == 0) { CALL(bf16, 8); } else { CALL(bf16, 1); }        \
    } else {                                                              \
      lcgan_set_error("bad dtype code %d", (int)(dt));                    \
      return 1;                                                           \
    }                                                                     \
  } while (0)

}  // namespace

// The flow fields (2 fp32 channels, custom_layers.py:150-151: box_filter(flow)): one thread per pixel, float2 taps,
// 32-bit indices.  The generic scalar kernel spent ~125 us per launch on them (64-bit div / mod per element, one
// channel per thread): 5 ms per iteration at 1024^2.
__global__ void __launch_bounds__(kThreads)
box3_c2_kernel(const float2* __restrict__ a, float2* __restrict__ out, int N, int H, int W) {
  const uint32_t total = (uint32_t)N * H * W;
  for (uint32_t idx = blockIdx.x * kThreads + threadIdx.x; idx < total; idx += gridDim.x * kThreads) {
    const int x = (int)(idx % (uint32_t)W);
    const uint32_t t = idx / (uint32_t)W;
    const int y = (int)(t % (uint32_t)H);
    const float2* img = a + (size_t)(t / (uint32_t)H) * H * W;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float2 v = img[(size_t)yy * W + xx];
          acc.x += v.x; acc.y += v.y;
        }
      }
    }
    out[idx] = make_float2(acc.x * (1.f / 9.f), acc.y * (1.f / 9.f));
  }
}

// the tiled kernel for every flavour (mask / style / reduction / post-mask); bf16 windows arrive by TMA
static void launch_box_tile(const void* a, const void* mask, void* out, int dt, int N, int H, int W, int C, float pre_slope,
                            float pre_gain, float post_slope, float post_gain, const float* cs, float* red,
                            const void* post_mask, cudaStream_t s) {
  static const bool no_tma = getenv("LCGAN_BOX_NO_TMA") != nullptr;
  const int cc = dt == LCGAN_BF16 ? 32 : 16;
  const int th = mask ? 8 : kBoxTH;
  const dim3 grid(N * ((H + th - 1) / th) * ((W + kBoxTW - 1) / kBoxTW) * (C / cc));
  CUtensorMap ta{}, tm{};
  const bool tma = dt == LCGAN_BF16 && !no_tma && tmaw::make_map(&ta, a, N, H, W, C, kBoxWW, th + 2, false) &&
                   (!mask || tmaw::make_map(&tm, mask, N, H, W, C, kBoxWW, th + 2, false));
#define BT(T, M, THH, TMA)                                                                                        \
  box3_tile_kernel<T, M, THH, TMA><<<grid, kThreads, 0, s>>>(ta, tm, (const T*)a, (const T*)mask, (T*)out, H, W, C, \
                                                             pre_slope, pre_gain, post_slope, post_gain, cs, red,  \
                                                             (const T*)post_mask)
  if (dt == LCGAN_BF16) {
    if (tma) { if (mask) BT(bf16, true, 8, true); else BT(bf16, false, 16, true); }
    else { if (mask) BT(bf16, true, 8, false); else BT(bf16, false, 16, false); }
  } else {
    if (mask) BT(float, true, 8, false); else BT(float, false, 16, false);
  }
#undef BT
}

extern "C" int lcgan_box3(const void* a, const void* mask, void* out, int dt, int N, int H, int W, int C,
                          float pre_slope, float pre_gain, float post_slope, float post_gain, void* stream) {
  LCGAN_CHECK(a && out && N > 0 && H > 0 && W > 0 && C > 0, "box3: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int cc = dt == LCGAN_BF16 ? 32 : 16;
  if (dt == LCGAN_F32 && C == 2 && !mask && pre_slope == 1.f && pre_gain == 1.f && post_slope == 1.f && post_gain == 1.f &&
      (int64_t)N * H * W < (1LL << 31) - (1LL << 24) && (uintptr_t)a % 8 == 0 && (uintptr_t)out % 8 == 0) {
    box3_c2_kernel<<<grid_for((int64_t)N * H * W), kThreads, 0, s>>>((const float2*)a, (float2*)out, N, H, W);
  } else
  if ((dt == LCGAN_BF16 || dt == LCGAN_F32) && C % cc == 0 && W >= kBoxTW && H >= kBoxTH &&
      getenv("LCGAN_BOX_NO_TILE") == nullptr) {
    launch_box_tile(a, mask, out, dt, N, H, W, C, pre_slope, pre_gain, post_slope, post_gain, nullptr, nullptr, nullptr, s);
  } else if (H % 8 == 0 && (int64_t)N * (H / 8) * W * (C / 8) >= 148LL * 64) {
#define CALL(T, V)                                                                                   \
  do {                                                                                               \
    const int g_ = grid_for((int64_t)N * (H / 8) * W * (C / V));                                     \
    if (mask)                                                                                        \
      box3_strip_kernel<T, V, 8, true><<<g_, kThreads, 0, s>>>(                                      \
          (const T*)a, (const T*)mask, (T*)out, N, H, W, C, pre_slope, pre_gain, post_slope, post_gain); \
    else                                                                                             \
      box3_strip_kernel<T, V, 8, false><<<g_, kThreads, 0, s>>>(                                     \
          (const T*)a, (const T*)mask, (T*)out, N, H, W, C, pre_slope, pre_gain, post_slope, post_gain); \
  } while (0)
    DISPATCH_TV(dt, C, CALL);
#undef CALL
  } else {
#define CALL(T, V)                                                                              \
  box3_kernel<T, V><<<grid_for((int64_t)N * H * W * (C / V)), kThreads, 0, s>>>(                \
      (const T*)a, (const T*)mask, (T*)out, N, H, W, C, pre_slope, pre_gain, post_slope, post_gain)
    DISPATCH_TV(dt, C, CALL);
#undef CALL
  }
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// out = box3(a) * lrelu'(y) * gain and r0[b,c] += sum_p out: the backward of "conv -> lrelu*gain -> box filter" from the
// gradient of the box filter's output to the gradient of the conv's pre-activation, in one pass.  Tile shapes only.
extern "C" int lcgan_box3_postmask(const void* a, const void* y, void* out, float* r0, int dt, int N, int H, int W,
                                   int C, float slope, float gain, void* stream) {
  LCGAN_CHECK(a && y && out && N > 0 && H > 0 && W > 0 && C > 0, "box3_postmask: bad arguments");
  const int cc = dt == LCGAN_BF16 ? 32 : 16;
  LCGAN_CHECK((dt == LCGAN_BF16 || dt == LCGAN_F32) && C % cc == 0 && W >= kBoxTW && H >= kBoxTH,
              "box3_postmask: needs C %% %d == 0, W >= %d, H >= %d (use lcgan_box3 + lcgan_act_bwd otherwise)", cc, kBoxTW, kBoxTH);
  cudaStream_t s = (cudaStream_t)stream;
  launch_box_tile(a, nullptr, out, dt, N, H, W, C, 1.f, 1.f, slope, gain, nullptr, r0, y, s);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// Box filter with the next layer's style folded in (see box3_tile_kernel): tile shapes only.
extern "C" int lcgan_box3_cs(const void* a, const void* mask, void* out, const float* cs, float* red, int dt, int N,
                             int H, int W, int C, float pre_slope, float pre_gain, float post_slope, float post_gain,
                             void* stream) {
  LCGAN_CHECK(a && out && cs && N > 0 && H > 0 && W > 0 && C > 0, "box3_cs: bad arguments");
  LCGAN_CHECK(!red || mask, "box3_cs: the reduction needs the mask tensor");
  const int cc = dt == LCGAN_BF16 ? 32 : 16;
  LCGAN_CHECK((dt == LCGAN_BF16 || dt == LCGAN_F32) && C % cc == 0 && W >= kBoxTW && H >= kBoxTH,
              "box3_cs: needs C %% %d == 0, W >= %d, H >= %d (use lcgan_box3 + lcgan_modulate otherwise)", cc, kBoxTW, kBoxTH);
  cudaStream_t s = (cudaStream_t)stream;
  launch_box_tile(a, mask, out, dt, N, H, W, C, pre_slope, pre_gain, post_slope, post_gain, cs, red, nullptr, s);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_pool2(const void* x, void* y, int dt, int N, int H, int W, int C, float scale, void* stream) {
  LCGAN_CHECK(x && y && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "pool2: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
#define CALL(T, V)                                                                              \
  pool2_kernel<T, V><<<grid_for((int64_t)N * (H / 2) * (W / 2) * (C / V)), kThreads, 0, s>>>(   \
      (const T*)x, (T*)y, N, H, W, C, scale)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_up2(const void* x, void* y, int dt, int N, int H, int W, int C, float scale, void* stream) {
  LCGAN_CHECK(x && y && N > 0 && H > 0 && W > 0 && C > 0, "up2: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
#define CALL(T, V)                                                                              \
  up2_kernel<T, V><<<grid_for((int64_t)N * H * W * 4 * (C / V)), kThreads, 0, s>>>(             \
      (const T*)x, (T*)y, N, H, W, C, scale)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_up2_add(const void* a, const void* sm, void* out, int dt, int N, int H, int W, int C, float scale,
                             void* stream) {
  LCGAN_CHECK(a && sm && out && N > 0 && H > 0 && W > 0 && C > 0, "up2_add: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
#define CALL(T, V)                                                                                       \
  do {                                                                                                   \
    const int64_t total_ = (int64_t)N * H * W * 4 * (C / V);                                             \
    if (total_ < (1LL << 31) - (1LL << 24))                                                              \
      up2_add_kernel<T, V, uint32_t><<<grid_for(total_), kThreads, 0, s>>>((const T*)a, (const T*)sm, (T*)out, N, H, W, C, scale); \
    else                                                                                                 \
      up2_add_kernel<T, V, int64_t><<<grid_for(total_), kThreads, 0, s>>>((const T*)a, (const T*)sm, (T*)out, N, H, W, C, scale);  \
  } while (0)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_box3_pool2(const void* g, void* out, int dt, int N, int H, int W, int C, void* stream) {
  LCGAN_CHECK(g && out && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "box3_pool2: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
#define CALL(T, V)                                                                                       \
  do {                                                                                                   \
    const int64_t total_ = (int64_t)N * (H / 2) * (W / 2) * (C / V);                                     \
    if (total_ < (1LL << 31) - (1LL << 24))                                                              \
      box3_pool2_kernel<T, V, uint32_t><<<grid_for(total_), kThreads, 0, s>>>((const T*)g, (T*)out, N, H, W, C); \
    else                                                                                                 \
      box3_pool2_kernel<T, V, int64_t><<<grid_for(total_), kThreads, 0, s>>>((const T*)g, (T*)out, N, H, W, C);  \
  } while (0)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_up2box_add(const void* sk, const void* t, void* out, int dt, int N, int H, int W, int C,
                                void* stream) {
  LCGAN_CHECK(sk && t && out && N > 0 && H > 0 && W > 0 && C > 0, "up2box_add: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  static const bool no_quad = getenv("LCGAN_UP2BOX_NO_QUAD") != nullptr;
#define CALL(T, V)                                                                              \
  do {                                                                                          \
    if (V > 1 && !no_quad && (int64_t)N * H * W * (C / V) < (1LL << 31) - (1LL << 24))          \
      up2box_add_quad_kernel<T, V><<<grid_for((int64_t)N * H * W * (C / V)), kThreads, 0, s>>>( \
          (const T*)sk, (const T*)t, (T*)out, N, H, W, C);                                      \
    else                                                                                        \
      up2box_add_kernel<T, V><<<grid_for((int64_t)N * H * W * 4 * (C / V)), kThreads, 0, s>>>(  \
          (const T*)sk, (const T*)t, (T*)out, N, H, W, C);                                      \
  } while (0)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

static inline int pix_per_block_for(int N, int P, int blocks_per_sm = 8) {
  // aim for ~8 CTAs per SM overall, at least 64 pixels per block
  int64_t want_blocks = 148LL * blocks_per_sm;
  int64_t per_b = (want_blocks + N - 1) / N;
  int ppb = (int)((P + per_b - 1) / per_b);
  if (ppb < 64) ppb = 64;
  return ppb;
}

extern "C" int lcgan_act_bwd(const void* dy, const void* y, void* gout, const float* d, float* r0, float* r1,
                             int dt, int N, int P, int C, float slope, float gain, void* stream) {
  LCGAN_CHECK(dy && y && gout && N > 0 && P > 0 && C > 0, "act_bwd: bad arguments");
  LCGAN_CHECK(slope > 0.f && gain > 0.f, "act_bwd: slope and gain must be positive");
  LCGAN_CHECK(N <= 65535, "act_bwd: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  const char* e_bps = getenv("LCGAN_ACT_BPS");          // tuning knobs (experiments only)
  const int ppb = pix_per_block_for(N, P, e_bps ? atoi(e_bps) : 4);   // one resident wave (measured best: 3-6)
  dim3 grid(ceil_div(P, ppb), N);
  int* sems = nullptr;
  if (lcgan_det_enabled() && (r0 || r1)) {
    LCGAN_CHECK(N <= kDetSems, "act_bwd: batch too large for deterministic mode");
    LCGAN_CUDA(cudaGetSymbolAddress((void**)&sems, g_sems));
  }
#define CALL(T, V)                                                                              \
  act_bwd_kernel<T, V, kStreamStages><<<grid, kThreads, 0, s>>>((const T*)dy, (const T*)y, (T*)gout, d, r0, \
                                                                r1, P, C, slope, gain, ppb, sems)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_modulate(const void* x, const float* sc, void* xs, int dt, int N, int P, int C, void* stream) {
  LCGAN_CHECK(x && sc && xs && N > 0 && P > 0 && C > 0, "modulate: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
#define CALL(T, V)                                                                              \
  modulate_kernel<T, V><<<grid_for((int64_t)N * P * (C / V)), kThreads, 0, s>>>((const T*)x, sc, (T*)xs, N, P, C)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_modulate_bwd(const void* x, const void* t, const float* sc, void* dx, float* ds, int dt,
                                  int N, int P, int C, void* stream) {
  LCGAN_CHECK(x && t && sc && dx && ds && N > 0 && P > 0 && C > 0, "modulate_bwd: bad arguments");
  LCGAN_CHECK(N <= 65535, "modulate_bwd: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  const char* e_bps = getenv("LCGAN_ACT_BPS");          // tuning knob (experiments only)
  const int ppb = pix_per_block_for(N, P, e_bps ? atoi(e_bps) : 3);   // measured best: 1-3 (6.0 vs 5.2 TB/s at 8)
  dim3 grid(ceil_div(P, ppb), N);
  int* sems = nullptr;
  if (lcgan_det_enabled()) {
    LCGAN_CHECK(N <= kDetSems, "modulate_bwd: batch too large for deterministic mode");
    LCGAN_CUDA(cudaGetSymbolAddress((void**)&sems, g_sems));
  }
#define CALL(T, V)                                                                              \
  modulate_bwd_kernel<T, V><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)t, sc, (T*)dx, ds, P, C, ppb, sems)
  DISPATCH_TV(dt, C, CALL);
#undef CALL
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_cast(const void* in, void* out, int dt_in, int dt_out, int64_t n, void* stream) {
  LCGAN_CHECK(in && out && n >= 0, "cast: bad arguments");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for(n);
  if (dt_in == LCGAN_F32 && dt_out == LCGAN_BF16) cast_kernel<float, bf16><<<g, kThreads, 0, s>>>((const float*)in, (bf16*)out, n);
  else if (dt_in == LCGAN_BF16 && dt_out == LCGAN_F32) cast_kernel<bf16, float><<<g, kThreads, 0, s>>>((const bf16*)in, (float*)out, n);
  else if (dt_in == LCGAN_F32 && dt_out == LCGAN_F32) cast_kernel<float, float><<<g, kThreads, 0, s>>>((const float*)in, (float*)out, n);
  else if (dt_in == LCGAN_BF16 && dt_out == LCGAN_BF16) cast_kernel<bf16, bf16><<<g, kThreads, 0, s>>>((const bf16*)in, (bf16*)out, n);
  else { lcgan_set_error("cast: bad dtype codes %d -> %d", dt_in, dt_out); return 1; }
  LCGAN_LAUNCH_CHECK();
  return 0;
}
