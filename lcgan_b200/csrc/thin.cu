// Memory-bound contractions that are not GEMM-shaped (SURVEY "first/last layers are not GEMM-shaped"):
//   thin-out  : Cout <= 4  (flow layers C->2, to-RGB C->3, data-gradient of from-RGB)     reads X once
//   thin-in   : Cin  <= 4  (from-RGB 3->C, data-gradients of the flow / to-RGB layers)    writes Y once
//   thin-wgrad: weight gradient when one side has <= 4 channels                           reads both once
//   skinny    : linear layers with batch <= 32 rows (mapping network, style affines, demod
//               coefficients, heads' weight gradients)                                   reads W once
// They are selected inside lcgan_tapconv_simt / lcgan_tapconv_wgrad_simt, so the C ABI is unchanged.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxThin = 4;
__device__ int g_sems[kDetSems];     // deterministic-mode turn semaphores (common.cuh)
static int* det_sems() {             // nullptr unless deterministic mode is on
  int* p = nullptr;
  if (lcgan_det_enabled()) cudaGetSymbolAddress((void**)&p, g_sems);
  return p;
}
constexpr int kSmemFloats = 11 * 1024;   // 44 KiB of weights in static shared memory

template <typename T, int V> __device__ __forceinline__ void ldv(const T* p, float* f) {
  if constexpr (V == 1) f[0] = ldf(p); else { Vec16<T> v; v.load(p); v.unpack(f); }
}
template <typename T, int V> __device__ __forceinline__ void stv(T* p, const float* f) {
  if constexpr (V == 1) stf(p, f[0]); else { Vec16<T> v; v.pack(f); v.store(p); }
}

__device__ __forceinline__ float ldw(const void* w, int w_dtype, int64_t i) {
  return w_dtype == LCGAN_F32 ? reinterpret_cast<const float*>(w)[i]
                              : __bfloat162float(reinterpret_cast<const bf16*>(w)[i]);
}

__device__ __forceinline__ float epilogue(const lcgan_tapconv& d, float acc, int b, int o, const float* rowscale,
                                          const float* bias) {
  float v = acc * d.acc_scale;
  if (rowscale) v *= rowscale[(int64_t)b * d.Cout + o];
  if (bias) v += bias[o] * d.bias_scale;
  return (v > 0.f ? v : v * d.slope) * d.gain;
}

// ------------------------------------------------------------------------------------------
// thin-out: a group of G lanes (G = min(32, Cin/V), a power of two) owns one lattice point; lane g
// reads channel vectors g, g+G, ... so every load instruction covers contiguous memory, then the
// Cout (<= 4) partial sums are combined with log2(G) shuffles.  Weights sit in shared memory and are
// read as 16-byte vectors.  (A thread-per-point variant was measured 2x slower on the 64..512-channel
// flow layers: its strided 16-byte loads thrash L1.)
// ------------------------------------------------------------------------------------------
template <typename TX, typename TY, int V>
__global__ void __launch_bounds__(kThreads)
thin_out_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const void* __restrict__ w, TY* __restrict__ y,
                const float* __restrict__ rowscale, const float* __restrict__ bias, const TY* __restrict__ residual,
                int G) {
  extern __shared__ __align__(16) float ws[];               // [o][t][c], sized by the launch (<= kSmemFloats)
  const int tc = d.ntaps * d.Cin;
  for (int i = threadIdx.x; i < d.Cout * tc; i += kThreads) {
    const int o = i / tc, r = i - o * tc, t = r / d.Cin, c = r - t * d.Cin;
    ws[i] = ldw(w, d.w_dtype, (int64_t)o * d.w_ld + (int64_t)d.wtap[t] * d.Cin + c);
  }
  __syncthreads();
  const int cv = d.Cin / V;
  // 32-bit lattice index (the host falls back to the generic kernel beyond 2^31 points): the 64-bit
  // div/mod chain of the decode was most of this kernel's instructions
  const uint32_t rows = (uint32_t)d.N * d.MH * d.MW;
  const uint32_t gl = threadIdx.x % G, gpb = kThreads / G;
  const uint32_t rows_pad = (rows + gpb - 1) / gpb * gpb;
  const uint32_t MW = d.MW, MH = d.MH;
  for (uint32_t r = blockIdx.x * gpb + threadIdx.x / G; r < rows_pad; r += gridDim.x * gpb) {
    const bool live = r < rows;
    float acc[kMaxThin] = {0.f, 0.f, 0.f, 0.f};
    int b = 0, m = 0, n = 0;
    if (live) {
      n = (int)(r % MW);
      const uint32_t q = r / MW;
      m = (int)(q % MH);
      b = (int)(q / MH);
      for (int t = 0; t < d.ntaps; ++t) {
        const int iy = m * d.is + d.dy[t], ix = n * d.is + d.dx[t];
        if (iy < 0 || iy >= d.IH || ix < 0 || ix >= d.IW) continue;
        const TX* xp = x + b * d.xs_n + iy * d.xs_h + ix * d.xs_w;
        const float* wt = ws + t * d.Cin;
        for (int v = (int)gl; v < cv; v += G) {
          float f[V];
          if constexpr (V == 1) f[0] = ldf(xp + (int64_t)v * d.xs_c); else ldv<TX, V>(xp + v * V, f);
#pragma unroll
          for (int o = 0; o < kMaxThin; ++o) {
            if (o < d.Cout) {
              const float* wp = wt + o * tc + v * V;
              if constexpr (V % 4 == 0) {
#pragma unroll
                for (int i = 0; i < V; i += 4) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wp + i);
                  acc[o] = fmaf(f[i], w4.x, acc[o]); acc[o] = fmaf(f[i + 1], w4.y, acc[o]);
                  acc[o] = fmaf(f[i + 2], w4.z, acc[o]); acc[o] = fmaf(f[i + 3], w4.w, acc[o]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < V; ++i) acc[o] = fmaf(f[i], wp[i], acc[o]);
              }
            }
          }
        }
      }
    }
    for (int s = G >> 1; s > 0; s >>= 1) {
#pragma unroll
      for (int o = 0; o < kMaxThin; ++o)
        if (o < d.Cout) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    }
    if (live && gl == 0) {
      const int64_t base = b * d.ys_n + (int64_t)(m * d.os + d.py) * d.ys_h + (int64_t)(n * d.os + d.px) * d.ys_w;
#pragma unroll
      for (int o = 0; o < kMaxThin; ++o) {
        if (o < d.Cout) {
          float v = epilogue(d, acc[o], b, o, rowscale, bias);
          if (residual) v += ldf(residual + base + o * d.ys_c);
          stf(y + base + o * d.ys_c, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// thin-in: one thread = one output channel vector of one lattice point
// ------------------------------------------------------------------------------------------
template <typename TX, typename TY, int V, int VPT>
__global__ void __launch_bounds__(kThreads)
thin_in_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const void* __restrict__ w, TY* __restrict__ y,
               const float* __restrict__ rowscale, const float* __restrict__ bias, const TY* __restrict__ residual) {
  // one thread = VPT output channel vectors of one lattice point, INTERLEAVED over the og threads of
  // the point (thread g owns vectors g, g+og, ...): a warp stores contiguous runs, and the weights
  // are laid out [t][c][quad][o/4] (V%4==0) so that neighbouring threads read neighbouring 16-byte
  // quads - the former [t][c][o] layout with 4 consecutive vectors per thread put the threads of a
  // quarter-warp 128 bytes apart, an 8-way bank conflict on every weight load (measured 0.25 TB/s).
  extern __shared__ __align__(16) float ws[];               // sized by the launch (<= kSmemFloats)
  const int tc = d.ntaps * d.Cin;
  constexpr int Q = V % 4 == 0 ? V / 4 : 1;                 // 16-byte quads per output vector
  const int nvec = d.Cout / V;                              // output vectors per lattice point
  for (int i = threadIdx.x; i < d.Cout * tc; i += kThreads) {
    const int o = i % d.Cout, r = i / d.Cout, t = r / d.Cin, c = r - t * d.Cin;
    const float wv = ldw(w, d.w_dtype, (int64_t)o * d.w_ld + (int64_t)d.wtap[t] * d.Cin + c);
    if constexpr (V % 4 == 0) {
      const int vec = o / V, q = (o % V) / 4, e = o % 4;
      ws[(r * Q + q) * (nvec * 4) + vec * 4 + e] = wv;
    } else {
      ws[i] = wv;
    }
  }
  __syncthreads();
  const uint32_t og = d.Cout / (V * VPT);
  const uint32_t total = (uint32_t)d.N * d.MH * d.MW * og;   // < 2^31 (host-checked)
  const uint32_t MW = d.MW, MH = d.MH;
  for (uint32_t idx = blockIdx.x * kThreads + threadIdx.x; idx < total; idx += gridDim.x * kThreads) {
    const int g = (int)(idx % og);
    uint32_t r = idx / og;
    const int n = (int)(r % MW); r /= MW;
    const int m = (int)(r % MH);
    const int b = (int)(r / MH);
    float acc[VPT][V];
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
    for (int t = 0; t < d.ntaps; ++t) {
      const int iy = m * d.is + d.dy[t], ix = n * d.is + d.dx[t];
      if (iy < 0 || iy >= d.IH || ix < 0 || ix >= d.IW) continue;
      const TX* xp = x + b * d.xs_n + iy * d.xs_h + ix * d.xs_w;
      for (int c = 0; c < d.Cin; ++c) {
        const float xv = ldf(xp + c * d.xs_c);
        const float* wp = ws + (t * d.Cin + c) * d.Cout;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const int vec = j * og + g;
          if constexpr (V % 4 == 0) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              const float4 w4 = *reinterpret_cast<const float4*>(wp + q * (nvec * 4) + vec * 4);
              acc[j][4 * q] = fmaf(xv, w4.x, acc[j][4 * q]); acc[j][4 * q + 1] = fmaf(xv, w4.y, acc[j][4 * q + 1]);
              acc[j][4 * q + 2] = fmaf(xv, w4.z, acc[j][4 * q + 2]); acc[j][4 * q + 3] = fmaf(xv, w4.w, acc[j][4 * q + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < V; ++i) acc[j][i] = fmaf(xv, wp[vec * V + i], acc[j][i]);
          }
        }
      }
    }
    const int64_t base = b * d.ys_n + (int64_t)(m * d.os + d.py) * d.ys_h + (int64_t)(n * d.os + d.px) * d.ys_w;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int oj = (j * og + g) * V;
#pragma unroll
      for (int i = 0; i < V; ++i) acc[j][i] = epilogue(d, acc[j][i], b, oj + i, rowscale, bias);
      if constexpr (V == 1) {
        if (residual) acc[j][0] += ldf(residual + base + oj * d.ys_c);
        stf(y + base + oj * d.ys_c, acc[j][0]);
      } else {
        if (residual) {
          float rr[V];
          ldv<TY, V>(residual + base + oj, rr);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[j][i] += rr[i];
        }
        stv<TY, V>(y + base + oj, acc[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// thin-up2: ALL FOUR output phases of conv_transpose2d(k3, s2, p1, op1) with Cout <= 4 (the flow
// layers, custom_layers.py:78 with C -> 2) in one pass over X: a group of G lanes owns input pixel
// (m, n), reads the channel vectors of x[m..m+1][n..n+1] once and produces the 2x2 output block
//   out(2m  , 2n  ) = x[m,n] w11
//   out(2m  , 2n+1) = x[m,n] w12 + x[m,n+1] w10
//   out(2m+1, 2n  ) = x[m,n] w21 + x[m+1,n] w01
//   out(2m+1, 2n+1) = x[m,n] w22 + x[m,n+1] w20 + x[m+1,n] w02 + x[m+1,n+1] w00
// (the four phase launches of the generic path each re-read X and ran at ~1 TB/s).
// ------------------------------------------------------------------------------------------
template <typename TX, typename TY, int V, int CO>
__global__ void __launch_bounds__(kThreads)
thin_up2_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const void* __restrict__ w, TY* __restrict__ y,
                const float* __restrict__ rowscale, const float* __restrict__ bias, int G) {
  // weights [tap 9][o][quad][vector][4]: lane g reads quad q of its vector at a 16-byte lane stride
  // (conflict-free; [t][o][c] put the lanes 32 bytes apart, a 2-way conflict on every weight load)
  extern __shared__ __align__(16) float ws[];               // sized by the launch (<= kSmemFloats)
  constexpr int Q = V % 4 == 0 ? V / 4 : 1, E = V % 4 == 0 ? 4 : 1;
  const int Cin = d.Cin, oc = CO * Cin, cv = Cin / V;
  for (int i = threadIdx.x; i < 9 * oc; i += kThreads) {
    const int t = i / oc, r = i - t * oc, o = r / Cin, c = r - o * Cin;
    const int vec = c / V, q = (c % V) / E, e = c % E;
    ws[((t * CO + o) * Q + q) * (cv * E) + vec * E + e] = ldw(w, d.w_dtype, (int64_t)o * d.w_ld + (int64_t)t * Cin + c);
  }
  __syncthreads();
  const uint32_t rows = (uint32_t)d.N * d.IH * d.IW;       // < 2^31 (checked on the host)
  const uint32_t gl = threadIdx.x % G, gpb = kThreads / G;
  const uint32_t rows_pad = (rows + gpb - 1) / gpb * gpb;
  const uint32_t IW = d.IW, IH = d.IH;
  for (uint32_t r = blockIdx.x * gpb + threadIdx.x / G; r < rows_pad; r += gridDim.x * gpb) {
    const bool live = r < rows;
    float acc[4][CO];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int o = 0; o < CO; ++o) acc[ph][o] = 0.f;
    uint32_t b = 0, m = 0, n = 0;
    if (live) {
      n = r % IW;
      const uint32_t q = r / IW;
      m = q % IH;
      b = q / IH;
      const bool okx = n + 1 < IW, oky = m + 1 < IH;
      const TX* p00 = x + b * d.xs_n + (int64_t)m * d.xs_h + (int64_t)n * d.xs_w;
      const TX* p01 = okx ? p00 + d.xs_w : p00;
      const TX* p10 = oky ? p00 + d.xs_h : p00;
      const TX* p11 = p10 + (okx ? d.xs_w : 0);
      const float m01 = okx ? 1.f : 0.f, m10 = oky ? 1.f : 0.f, m11 = m01 * m10;
      for (int v = gl; v < cv; v += G) {
        float f00[V], f01[V], f10[V], f11[V];
        ldv<TX, V>(p00 + v * V, f00);
        ldv<TX, V>(p01 + v * V, f01);
        ldv<TX, V>(p10 + v * V, f10);
        ldv<TX, V>(p11 + v * V, f11);
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          auto dot = [&](const float* f, int t) {
            const float* wp = ws + ((t * CO + o) * Q) * (cv * E) + v * E;
            float sacc = 0.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              if constexpr (E == 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(wp + q * (cv * E));
                sacc = fmaf(f[4 * q], w4.x, sacc); sacc = fmaf(f[4 * q + 1], w4.y, sacc);
                sacc = fmaf(f[4 * q + 2], w4.z, sacc); sacc = fmaf(f[4 * q + 3], w4.w, sacc);
              } else {
                sacc = fmaf(f[q], wp[q * (cv * E)], sacc);
              }
            }
            return sacc;
          };
          acc[0][o] += dot(f00, 4);
          acc[1][o] += dot(f00, 5) + m01 * dot(f01, 3);
          acc[2][o] += dot(f00, 7) + m10 * dot(f10, 1);
          acc[3][o] += dot(f00, 8) + m01 * dot(f01, 6) + m10 * dot(f10, 2) + m11 * dot(f11, 0);
        }
      }
    }
    for (int sft = G >> 1; sft > 0; sft >>= 1) {
#pragma unroll
      for (int ph = 0; ph < 4; ++ph)
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[ph][o] += __shfl_xor_sync(0xffffffffu, acc[ph][o], sft);
    }
    if (live) {
#pragma unroll
      for (int ph = 0; ph < 4; ++ph)
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[ph][o] = epilogue(d, acc[ph][o], b, o, rowscale, bias);
      if (CO == 2 && sizeof(TY) == 4 && d.ys_c == 1 && d.ys_w == 2 && (d.ys_h & 3) == 0 && (d.ys_n & 3) == 0) {
        // dense [.., 2H, 2W, 2] f32: output row 2m+j holds the two phases (j,0),(j,1) of this pixel as
        // one 16-byte store; lane j of the group writes row j
        if (gl < 2 || G == 1) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if ((int)gl == j || G == 1) {
              float* dst = reinterpret_cast<float*>(y) + b * d.ys_n + (int64_t)(2 * m + j) * d.ys_h + (int64_t)(2 * n) * 2;
              *reinterpret_cast<float4*>(dst) = make_float4(acc[2 * j][0], acc[2 * j][CO - 1], acc[2 * j + 1][0], acc[2 * j + 1][CO - 1]);
            }
          }
        }
      } else {
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
          if ((int)gl == (G >= 4 ? ph : 0)) {
            const int64_t base = b * d.ys_n + (int64_t)(2 * m + (ph >> 1)) * d.ys_h + (int64_t)(2 * n + (ph & 1)) * d.ys_w;
#pragma unroll
            for (int o = 0; o < CO; ++o) stf(y + base + o * d.ys_c, acc[ph][o]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// thin-wgrad: thread = (wide channel vector, tap, pixel lane); accumulates acc[thin][V] over its
// lattice points, then one atomic per element.  kThinIsG: the thin side is G (Cout <= 4) and the
// wide side is X at the tap-shifted position; otherwise thin = X (Cin <= 4), wide = G.
// ------------------------------------------------------------------------------------------
template <typename TW, typename TT, int V, bool kThinIsG>
__global__ void __launch_bounds__(kThreads)
thin_wgrad_kernel(const lcgan_tapconv d, const TW* __restrict__ wide, const TT* __restrict__ thin,
                  float* __restrict__ dw, float scale, int cv, int lanes, int64_t rows_per_block, int* sems) {
  __shared__ float red[kThreads * kMaxThin * V / 2];   // half the threads park their partials at a time
  const int per_lane = cv * d.ntaps;
  const int lane = threadIdx.x / per_lane;
  const bool active = lane < lanes;
  const int rem = threadIdx.x - lane * per_lane;
  const int t = active ? rem / cv : 0, v = active ? rem - t * cv : 0;
  const int nthin = kThinIsG ? d.Cout : d.Cin;
  const int64_t rows = (int64_t)d.N * d.MH * d.MW;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  float acc[kMaxThin][V];
#pragma unroll
  for (int o = 0; o < kMaxThin; ++o)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[o][i] = 0.f;
  if (active) {
    // incremental (b, m, n) decode: one division pair up front, then carries
    int64_t r = r_begin + lane;
    int n = (int)(r % d.MW);
    int64_t q = r / d.MW;
    int m = (int)(q % d.MH);
    int b = (int)(q / d.MH);
    for (; r < r_end; r += lanes) {
      const int iy = m * d.is + d.dy[t], ix = n * d.is + d.dx[t];
      if (iy >= 0 && iy < d.IH && ix >= 0 && ix < d.IW) {
        const int oy = m * d.os + d.py, ox = n * d.os + d.px;
        float f[V], th[kMaxThin];
        if constexpr (kThinIsG) {
          ldv<TW, V>(wide + b * d.xs_n + iy * d.xs_h + ix * d.xs_w + v * V, f);
          const TT* tp = thin + b * d.ys_n + oy * d.ys_h + ox * d.ys_w;
#pragma unroll
          for (int o = 0; o < kMaxThin; ++o) th[o] = o < nthin ? ldf(tp + o * d.ys_c) : 0.f;
        } else {
          ldv<TW, V>(wide + b * d.ys_n + oy * d.ys_h + ox * d.ys_w + v * V, f);
          const TT* tp = thin + b * d.xs_n + iy * d.xs_h + ix * d.xs_w;
#pragma unroll
          for (int o = 0; o < kMaxThin; ++o) th[o] = o < nthin ? ldf(tp + o * d.xs_c) : 0.f;
        }
#pragma unroll
        for (int o = 0; o < kMaxThin; ++o)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[o][i] = fmaf(th[o], f[i], acc[o][i]);
      }
      n += lanes;
      while (n >= d.MW) { n -= d.MW; if (++m == d.MH) { m = 0; ++b; } }
    }
  }
  // tree-reduce the pixel lanes through shared memory, then ONE atomic per weight element per block
  int span = 1;
  while (span < lanes) span <<= 1;
  for (int h = span >> 1; h >= 1; h >>= 1) {
    const bool writer = active && lane >= h && lane < 2 * h;
    const bool reader = active && lane < h && lane + h < lanes;
    __syncthreads();
    if (writer) {
      float* p = red + ((lane - h) * per_lane + rem) * (kMaxThin * V);
#pragma unroll
      for (int o = 0; o < kMaxThin; ++o)
#pragma unroll
        for (int i = 0; i < V; ++i) p[o * V + i] = acc[o][i];
    }
    __syncthreads();
    if (reader) {
      const float* p = red + (lane * per_lane + rem) * (kMaxThin * V);
#pragma unroll
      for (int o = 0; o < kMaxThin; ++o)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[o][i] += p[o * V + i];
    }
  }
  if (sems) det_block_begin(sems, blockIdx.x);       // deterministic mode: blocks add in index order
  if (active && lane == 0) {
    const int64_t wcol0 = (int64_t)d.wtap[t] * d.Cin;
    for (int o = 0; o < nthin; ++o)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float* dst = kThinIsG ? dw + (int64_t)o * d.w_ld + wcol0 + v * V + i
                              : dw + (int64_t)(v * V + i) * d.w_ld + wcol0 + o;
        if (sems) det_add(dst, acc[o][i] * scale);
        else atomicAdd(dst, acc[o][i] * scale);
      }
  }
  if (sems) det_block_end(sems, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------------
// thin-up2 weight gradient: all 9 taps of conv_transpose2d(k3, s2, p1, op1) with Cout <= 4 in one pass
// over X.  dW[o][t][c] = sum_{b,m,n} g[b, 2m-1+ki, 2n-1+kj, o] x[b,m,n,c], t = ki*3+kj.  Thread =
// (4 input channels, pixel lane): one 4-channel load of x and the 3x3 neighbourhood of g (shared by
// the threads of the pixel through L1) feed 9*CO*4 accumulators; pixel lanes are tree-reduced through
// shared memory and each block finishes with one atomic per weight element.  (The per-phase generic
// kernel made one thread per tap re-read X: 9 passes, 0.4 TB/s.)
// ------------------------------------------------------------------------------------------
template <typename TX, typename TG, int CO>
__global__ void __launch_bounds__(kThreads)
thin_up2_wgrad_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const TG* __restrict__ g,
                      float* __restrict__ dw, float scale, int cv4, int lanes, uint32_t rows_per_block, int* sems) {
  constexpr int NA = 9 * CO * 4;
  __shared__ float red[(kThreads / 2) * NA];
  const int v = threadIdx.x % cv4, lane = threadIdx.x / cv4;
  const uint32_t rows = (uint32_t)d.N * d.IH * d.IW;
  const uint32_t r_begin = blockIdx.x * rows_per_block;
  const uint32_t r_end = min(rows, r_begin + rows_per_block);
  const uint32_t IW = d.IW, IH = d.IH;
  const int OH = 2 * d.IH, OW = 2 * d.IW;
  float acc[9][CO][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][o][i] = 0.f;
  for (uint32_t r = r_begin + lane; r < r_end; r += lanes) {
    const int n = (int)(r % IW);
    const uint32_t q = r / IW;
    const int m = (int)(q % IH), b = (int)(q / IH);
    float xv[4];
    const TX* xp = x + b * d.xs_n + (int64_t)m * d.xs_h + (int64_t)n * d.xs_w + v * 4;
    if constexpr (sizeof(TX) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(xp);
      xv[0] = __uint_as_float(u.x << 16); xv[1] = __uint_as_float(u.x & 0xffff0000u);
      xv[2] = __uint_as_float(u.y << 16); xv[3] = __uint_as_float(u.y & 0xffff0000u);
    } else {
      const float4 u = *reinterpret_cast<const float4*>(xp);
      xv[0] = u.x; xv[1] = u.y; xv[2] = u.z; xv[3] = u.w;
    }
    const TG* gb = g + b * d.ys_n;
#pragma unroll
    for (int ki = 0; ki < 3; ++ki) {
      const int oy = 2 * m - 1 + ki;
      const bool oky = oy >= 0 && oy < OH;
#pragma unroll
      for (int kj = 0; kj < 3; ++kj) {
        const int ox = 2 * n - 1 + kj;
        const bool ok = oky && ox >= 0 && ox < OW;
        const TG* gp = gb + (int64_t)(ok ? oy : 0) * d.ys_h + (int64_t)(ok ? ox : 0) * d.ys_w;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const float gv = ok ? ldf(gp + o * d.ys_c) : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[ki * 3 + kj][o][i] = fmaf(gv, xv[i], acc[ki * 3 + kj][o][i]);
        }
      }
    }
  }
  // tree-reduce the pixel lanes (threads with equal v) through shared memory
  int span = 1;
  while (span < lanes) span <<= 1;
  for (int h = span >> 1; h >= 1; h >>= 1) {
    const bool writer = lane >= h && lane < 2 * h;
    const bool reader = lane < h && lane + h < lanes;
    __syncthreads();
    if (writer) {
      float* p = red + ((lane - h) * cv4 + v) * NA;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int o = 0; o < CO; ++o)
#pragma unroll
          for (int i = 0; i < 4; ++i) p[(t * CO + o) * 4 + i] = acc[t][o][i];
    }
    __syncthreads();
    if (reader) {
      const float* p = red + (lane * cv4 + v) * NA;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int o = 0; o < CO; ++o)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[t][o][i] += p[(t * CO + o) * 4 + i];
    }
  }
  if (sems) det_block_begin(sems, blockIdx.x);       // deterministic mode: blocks add in index order
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < CO; ++o)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* dst = dw + (int64_t)o * d.w_ld + (int64_t)t * d.Cin + v * 4 + i;
          if (sems) det_add(dst, acc[t][o][i] * scale);
          else atomicAdd(dst, acc[t][o][i] * scale);
        }
  }
  if (sems) det_block_end(sems, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------------
// skinny linear: out[m][n] = sum_k x[m][k] w[n][k], m <= 32.  One warp per output feature, lanes
// stride over K (coalesced w row), 32 accumulators per lane, shuffle reduction.
// ------------------------------------------------------------------------------------------
constexpr int kSkinnyKC = 256;    // K chunk staged in shared memory: 32 rows x 256 floats = 32 KiB

template <typename TX, typename TWt, typename TY>
__global__ void __launch_bounds__(kThreads)
skinny_linear_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const TWt* __restrict__ w, TY* __restrict__ y,
                     const float* __restrict__ rowscale, const float* __restrict__ bias, const TY* __restrict__ residual) {
  __shared__ float xs[32][kSkinnyKC];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n = blockIdx.x * (kThreads / 32) + warp;
  const bool nlive = n < d.Cout;
  const int M = d.N, K = d.Cin;
  float acc[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) acc[m] = 0.f;
  const TWt* wr = w + (int64_t)(nlive ? n : 0) * d.w_ld;
  static_assert(kThreads == kSkinnyKC, "the staging below gives thread t column t of the chunk");
  for (int k0 = 0; k0 < K; k0 += kSkinnyKC) {
    const int kc = min(kSkinnyKC, K - k0);
    // These layers are latency-, not bandwidth-bound (1 MB of weights, 64 KB of x; 24 us per launch in the graph-replay
    // profile): the staging loop kept ~4 of its 32 loads per thread in flight.  All 32 x loads and the 8 weight loads
    // are now issued before anything waits on them.
    float wv[kSkinnyKC / 32];
#pragma unroll
    for (int j = 0; j < kSkinnyKC / 32; ++j) {
      const int k = lane + 32 * j;
      wv[j] = (nlive && k < kc) ? ldf(wr + k0 + k) : 0.f;
    }
    float stage[32];
    {
      const int k = threadIdx.x;
#pragma unroll
      for (int m = 0; m < 32; ++m) stage[m] = (m < M && k < kc) ? ldf(x + m * d.xs_n + (int64_t)(k0 + k) * d.xs_c) : 0.f;
    }
    __syncthreads();                                          // the previous chunk is consumed
#pragma unroll
    for (int m = 0; m < 32; ++m) xs[m][threadIdx.x] = stage[m];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSkinnyKC / 32; ++j) {
#pragma unroll
      for (int m = 0; m < 32; ++m) acc[m] = fmaf(xs[m][lane + 32 * j], wv[j], acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < 32; ++m) acc[m] = warp_sum(acc[m]);
  float mine = 0.f;
#pragma unroll
  for (int m = 0; m < 32; ++m) if (lane == m) mine = acc[m];
  if (nlive && lane < M) {
    float v = epilogue(d, mine, lane, n, rowscale, bias);
    const int64_t off = lane * d.ys_n + n * d.ys_c;
    if (residual) v += ldf(residual + off);
    stf(y + off, v);
  }
}

// dW[n][k] += scale * sum_m g[m][n] x[m][k]; thread = one (n, 4 consecutive k)
template <typename TX, typename TG>
__global__ void __launch_bounds__(kThreads)
skinny_wgrad_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const TG* __restrict__ g, float* __restrict__ dw,
                    float scale) {
  const int K4 = (d.Cin + 3) / 4;
  const int64_t total = (int64_t)d.Cout * K4;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
    const int k0 = (int)(idx % K4) * 4;
    const int n = (int)(idx / K4);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m = 0; m < d.N; ++m) {
      const float gv = ldf(g + m * d.ys_n + n * d.ys_c);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (k0 + i < d.Cin) acc[i] = fmaf(gv, ldf(x + m * d.xs_n + (k0 + i) * d.xs_c), acc[i]);
    }
    float* o = dw + (int64_t)n * d.w_ld + k0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k0 + i < d.Cin) o[i] += acc[i] * scale;
  }
}

inline int pow2_group(int cv) { return (cv & (cv - 1)) ? 1 : (cv < 32 ? cv : 32); }

inline bool is_point(const lcgan_tapconv& d) {
  return d.IH == 1 && d.IW == 1 && d.OH == 1 && d.OW == 1 && d.ntaps == 1 && d.MH == 1 && d.MW == 1;
}

inline bool dense_inner(int64_t sc, int64_t sw, int64_t sh, int64_t sn, int C, int vec) {
  return sc == 1 && C % vec == 0 && sw % vec == 0 && sh % vec == 0 && sn % vec == 0;
}

inline int grid_cap(int64_t blocks, int per_sm) {
  const int64_t cap = 148LL * per_sm;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}


// ------------------------------------------------------------------------------------------
// Pointwise (1x1) thin layers at 32 wide channels - to-RGB 32 -> 3 and from-RGB 3 -> 32 of the 1024x1024
// model (cnn.py:19,87) and the data gradient of the former: ONE THREAD PER PIXEL.  A pixel is 64 bytes of
// bf16, so a warp covers 2 KB of contiguous activations and 128 contiguous bytes of every fp32 image plane;
// weights (pre-multiplied by acc_scale * gain) are broadcast reads from shared memory.  The generic thin
// kernels above split a pixel over 4 lanes (shuffle reduction, one lane in four storing 4-byte pieces):
// 1.49 ms for 32 -> 3 @1024^2 batch 32 against a 0.39 ms HBM floor.
// ------------------------------------------------------------------------------------------
template <typename TY, int CO>
__global__ void __launch_bounds__(kThreads)
pw_out32_kernel(const lcgan_tapconv d, const bf16* __restrict__ x, const void* __restrict__ w, TY* __restrict__ y,
                const float* __restrict__ rowscale, const float* __restrict__ bias, const TY* __restrict__ residual) {
  __shared__ float ws[CO][32];
  __shared__ float bs[CO];
  const float asg = d.acc_scale * d.gain;
  for (int i = threadIdx.x; i < CO * 32; i += kThreads) {
    const int o = i / 32, c = i % 32;
    ws[o][c] = o < d.Cout ? ldw(w, d.w_dtype, (int64_t)o * d.w_ld + (int64_t)d.wtap[0] * 32 + c) * asg : 0.f;
  }
  if (threadIdx.x < CO) bs[threadIdx.x] = (bias && (int)threadIdx.x < d.Cout) ? bias[threadIdx.x] * d.bias_scale * d.gain : 0.f;
  __syncthreads();
  const uint32_t hw = (uint32_t)d.MH * d.MW, rows = (uint32_t)d.N * hw, MW = d.MW;
  for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < rows; r += gridDim.x * kThreads) {
    const uint32_t b = r / hw, p = r - b * hw, m = p / MW, n = p - m * MW;
    const uint4* xp = reinterpret_cast<const uint4*>(x + (int64_t)r * 32);     // dense channels-last: pixel r
    const float4* cp = d.colscale ? reinterpret_cast<const float4*>(d.colscale + (int64_t)b * 32) : nullptr;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 u = xp[v];
      float f[8];
      unpack_raw16<bf16>(u, f);
      if (cp) {
        // the modulated activation as the tensor-core path would have read it: x * s rounded to bf16
        const float4 c0 = __ldg(cp + 2 * v), c1 = __ldg(cp + 2 * v + 1);
        const float cs8[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __bfloat162float(__float2bfloat16_rn(f[i] * cs8[i]));
      }
#pragma unroll
      for (int o = 0; o < CO; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[o] = fmaf(f[i], ws[o][v * 8 + i], acc[o]);
    }
    const int64_t base = (int64_t)b * d.ys_n + (int64_t)m * d.ys_h + (int64_t)n * d.ys_w;
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      if (o < d.Cout) {
        float t = acc[o];
        if (rowscale) t *= rowscale[(int64_t)b * d.Cout + o];
        t += bs[o];
        t = fmaxf(t, t * d.slope);
        if (residual) t += ldf(residual + base + o * d.ys_c);
        stf(y + base + o * d.ys_c, t);
      }
    }
  }
}

template <typename TX, int CI>
__global__ void __launch_bounds__(kThreads)
pw_in32_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const void* __restrict__ w, bf16* __restrict__ y,
               const float* __restrict__ rowscale, const float* __restrict__ bias, const bf16* __restrict__ residual) {
  __shared__ float ws[CI][32];
  __shared__ float bs[32];
  const float asg = d.acc_scale * d.gain;
  for (int i = threadIdx.x; i < CI * 32; i += kThreads) {
    const int c = i / 32, o = i % 32;
    ws[c][o] = c < d.Cin ? ldw(w, d.w_dtype, (int64_t)o * d.w_ld + (int64_t)d.wtap[0] * d.Cin + c) * asg : 0.f;
  }
  if (threadIdx.x < 32) bs[threadIdx.x] = bias ? bias[threadIdx.x] * d.bias_scale * d.gain : 0.f;
  __syncthreads();
  // A thread computes the 64 bytes of one pixel, but a warp stores through a 2 KB transpose tile so that every store
  // instruction writes 512 contiguous bytes: per-thread 16-byte stores at a 64-byte lane stride ran this kernel at
  // 2.7 TB/s where the load-side mirror image (pw_out32) reaches 3.9 and fully coalesced passes 5+.
  __shared__ uint4 tr[kThreads / 32][32 * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t hw = (uint32_t)d.MH * d.MW, rows = (uint32_t)d.N * hw, MW = d.MW;
  for (uint32_t rb = blockIdx.x * kThreads + warp * 32; rb < rows; rb += gridDim.x * kThreads) {
    const uint32_t r = min(rb + lane, rows - 1);                                 // (lanes past the end recompute the last pixel)
    const uint32_t b = r / hw, p = r - b * hw, m = p / MW, n = p - m * MW;
    const TX* xp = x + (int64_t)b * d.xs_n + (int64_t)m * d.xs_h + (int64_t)n * d.xs_w;
    float xv[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) xv[c] = c < d.Cin ? ldf(xp + c * d.xs_c) : 0.f;
    const uint4* rp = residual ? reinterpret_cast<const uint4*>(residual + (int64_t)r * 32) : nullptr;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < CI; ++c) t = fmaf(xv[c], ws[c][v * 8 + i], t);
        if (rowscale) t *= rowscale[(int64_t)b * 32 + v * 8 + i];
        t += bs[v * 8 + i];
        f[i] = fmaxf(t, t * d.slope);
      }
      if (rp) {
        float g[8];
        unpack_raw16<bf16>(rp[v], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += g[i];
      }
      Vec16<bf16> o;
      o.pack(f);
      tr[warp][lane * 4 + (v ^ ((lane >> 1) & 3))] = o.v;                       // XOR-swizzled slots: conflict-free both ways
    }
    __syncwarp();
    uint4* yw = reinterpret_cast<uint4*>(y + (int64_t)rb * 32);                 // dense channels-last: pixels rb .. rb+31
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane, px = idx >> 2, c = idx & 3;
      if (rb + px < rows) yw[idx] = tr[warp][px * 4 + (c ^ ((px >> 1) & 3))];
    }
    __syncwarp();
  }
}

// per-image weight gradient of the pointwise 32 -> CO layer: dwp[b][o][c] += sum_p g[b,p,o] x[b,p,c]
template <typename TG, int CO>
__global__ void __launch_bounds__(kThreads)
pw_wgrad32_kernel(const lcgan_tapconv d, const bf16* __restrict__ x, const TG* __restrict__ g, float* __restrict__ dwp) {
  const int b = blockIdx.y;
  const uint32_t hw = (uint32_t)d.MH * d.MW, MW = d.MW;
  float acc[CO][32];
#pragma unroll
  for (int o = 0; o < CO; ++o)
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[o][c] = 0.f;
  for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < hw; p += gridDim.x * kThreads) {
    const uint32_t m = p / MW, n = p - m * MW;
    const uint4* xp = reinterpret_cast<const uint4*>(x + ((int64_t)b * hw + p) * 32);
    const TG* gp = g + (int64_t)b * d.ys_n + (int64_t)m * d.ys_h + (int64_t)n * d.ys_w;
    float gv[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) gv[o] = o < d.Cout ? ldf(gp + o * d.ys_c) : 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float f[8];
      unpack_raw16<bf16>(xp[v], f);
#pragma unroll
      for (int o = 0; o < CO; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[o][v * 8 + i] = fmaf(gv[o], f[i], acc[o][v * 8 + i]);
    }
  }
  // warp reduction, then the 8 warps of the block through shared memory, one atomic per element and block
  __shared__ float red[kThreads / 32][CO * 32];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int o = 0; o < CO; ++o)
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float t = warp_sum(acc[o][c]);
      if (lane == c) red[warp][o * 32 + c] = t;
    }
  __syncthreads();
  for (int e = threadIdx.x; e < d.Cout * 32; e += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kThreads / 32; ++w8) t += red[w8][e];
    atomicAdd(dwp + (int64_t)b * d.Cout * 32 + e, t);
  }
}

// 1x1, unit strides, no offset
static bool is_pointwise(const lcgan_tapconv& d) {
  return d.ntaps == 1 && d.dy[0] == 0 && d.dx[0] == 0 && d.is == 1 && d.os == 1 && d.py == 0 && d.px == 0 &&
         d.slope > 0.f && d.slope <= 1.f && d.gain > 0.f && (int64_t)d.N * d.MH * d.MW < (1LL << 31) - (1 << 20);
}

}  // namespace

// returns -1 when the descriptor is not one of the special shapes (caller falls back to the tiled kernel)
int lcgan_thin_forward(const lcgan_tapconv& d, const void* x, const void* w, void* y, const float* rowscale,
                       const float* bias, const void* residual, cudaStream_t s) {
  const bool xf = d.x_dtype == LCGAN_F32, yf = d.y_dtype == LCGAN_F32;
  const int64_t rows = (int64_t)d.N * d.MH * d.MW;
  // ---- skinny linear -------------------------------------------------------------------
  if (is_point(d) && d.N <= 32) {
    const int grid = (d.Cout + kThreads / 32 - 1) / (kThreads / 32);
#define SK(TXT, TWT, TYT)                                                                                   \
    skinny_linear_kernel<TXT, TWT, TYT><<<grid, kThreads, 0, s>>>(d, (const TXT*)x, (const TWT*)w, (TYT*)y, \
                                                                  rowscale, bias, (const TYT*)residual)
    const bool wf = d.w_dtype == LCGAN_F32;
    if (xf && wf && yf) SK(float, float, float);
    else if (xf && wf && !yf) SK(float, float, bf16);
    else if (!xf && !wf && yf) SK(bf16, bf16, float);
    else if (!xf && !wf && !yf) SK(bf16, bf16, bf16);
    else return -1;
#undef SK
    LCGAN_LAUNCH_CHECK();
    return 0;
  }
  // ---- pointwise layers at 32 wide channels: one thread per pixel -------------------------------
  if (is_pointwise(d) && getenv("LCGAN_NO_PW") == nullptr) {
    const int grid = grid_cap((rows + kThreads - 1) / kThreads, 16);
    if (d.Cin == 32 && d.Cout <= kMaxThin && d.x_dtype == LCGAN_BF16 && (uintptr_t)x % 16 == 0 &&
        dense_inner(d.xs_c, d.xs_w, d.xs_h, d.xs_n, 32, 8) && d.xs_w == 32 && (d.MH == 1 || d.xs_h == (int64_t)d.MW * 32) &&
        (d.N == 1 || d.xs_n == (int64_t)d.MH * d.MW * 32)) {
      if (yf) pw_out32_kernel<float, kMaxThin><<<grid, kThreads, 0, s>>>(d, (const bf16*)x, w, (float*)y, rowscale, bias, (const float*)residual);
      else pw_out32_kernel<bf16, kMaxThin><<<grid, kThreads, 0, s>>>(d, (const bf16*)x, w, (bf16*)y, rowscale, bias, (const bf16*)residual);
      LCGAN_LAUNCH_CHECK();
      return 0;
    }
    if (d.Cout == 32 && d.Cin <= kMaxThin && d.y_dtype == LCGAN_BF16 && (uintptr_t)y % 16 == 0 &&
        (!residual || (uintptr_t)residual % 16 == 0) && d.ys_c == 1 && d.ys_w == 32 &&
        (d.MH == 1 || d.ys_h == (int64_t)d.MW * 32) && (d.N == 1 || d.ys_n == (int64_t)d.MH * d.MW * 32)) {
      if (xf) pw_in32_kernel<float, kMaxThin><<<grid, kThreads, 0, s>>>(d, (const float*)x, w, (bf16*)y, rowscale, bias, (const bf16*)residual);
      else pw_in32_kernel<bf16, kMaxThin><<<grid, kThreads, 0, s>>>(d, (const bf16*)x, w, (bf16*)y, rowscale, bias, (const bf16*)residual);
      LCGAN_LAUNCH_CHECK();
      return 0;
    }
  }
  if (d.colscale) {
    lcgan_set_error("tapconv: colscale is only implemented by the pointwise 32-channel thin kernel");
    return 1;
  }
  if (rows * (d.Cin <= kMaxThin ? d.Cout : 1) >= (1LL << 31) - (1 << 20)) return -1;   // 32-bit indices below
  // ---- thin-out --------------------------------------------------------------------------
  const size_t wbytes = (size_t)d.Cout * d.ntaps * d.Cin * sizeof(float);   // weight staging (dynamic smem)
  if (d.Cout <= kMaxThin && d.Cout * d.ntaps * d.Cin <= kSmemFloats) {
    const int vec = xf ? 4 : 8;
    const bool v_ok = dense_inner(d.xs_c, d.xs_w, d.xs_h, d.xs_n, d.Cin, vec) && ((uintptr_t)x % 16 == 0);
#define TO(TXT, TYT, VV)                                                                                    \
    do {                                                                                                    \
      const int G = pow2_group(d.Cin / VV);                                                                 \
      const int grid = grid_cap((rows + kThreads / G - 1) / (kThreads / G), 16);                            \
      thin_out_kernel<TXT, TYT, VV><<<grid, kThreads, wbytes, s>>>(d, (const TXT*)x, w, (TYT*)y, rowscale, bias, \
                                                              (const TYT*)residual, G);                     \
    } while (0)
    if (xf && yf) { if (v_ok) TO(float, float, 4); else TO(float, float, 1); }
    else if (xf && !yf) { if (v_ok) TO(float, bf16, 4); else TO(float, bf16, 1); }
    else if (!xf && yf) { if (v_ok) TO(bf16, float, 8); else TO(bf16, float, 1); }
    else { if (v_ok) TO(bf16, bf16, 8); else TO(bf16, bf16, 1); }
#undef TO
    LCGAN_LAUNCH_CHECK();
    return 0;
  }
  // ---- thin-in ---------------------------------------------------------------------------
  if (d.Cin <= kMaxThin && d.Cout * d.ntaps * d.Cin <= kSmemFloats) {
    const int vec = yf ? 4 : 8;
    const bool v_ok = dense_inner(d.ys_c, d.ys_w, d.ys_h, d.ys_n, d.Cout, vec) && ((uintptr_t)y % 16 == 0) &&
                      (!residual || (uintptr_t)residual % 16 == 0);
#define TI(TXT, TYT, VV)                                                                                    \
    do {                                                                                                    \
      if (VV > 1 && d.Cout % (VV * 4) == 0) {                                                               \
        const int grid = grid_cap((rows * (d.Cout / (VV * 4)) + kThreads - 1) / kThreads, 32);              \
        thin_in_kernel<TXT, TYT, VV, 4><<<grid, kThreads, wbytes, s>>>(d, (const TXT*)x, w, (TYT*)y, rowscale,   \
                                                                  bias, (const TYT*)residual);              \
      } else {                                                                                              \
        const int grid = grid_cap((rows * (d.Cout / VV) + kThreads - 1) / kThreads, 32);                    \
        thin_in_kernel<TXT, TYT, VV, 1><<<grid, kThreads, wbytes, s>>>(d, (const TXT*)x, w, (TYT*)y, rowscale,   \
                                                                  bias, (const TYT*)residual);              \
      }                                                                                                     \
    } while (0)
    if (xf && yf) { if (v_ok) TI(float, float, 4); else TI(float, float, 1); }
    else if (xf && !yf) { if (v_ok) TI(float, bf16, 8); else TI(float, bf16, 1); }
    else if (!xf && yf) { if (v_ok) TI(bf16, float, 4); else TI(bf16, float, 1); }
    else { if (v_ok) TI(bf16, bf16, 8); else TI(bf16, bf16, 1); }
#undef TI
    LCGAN_LAUNCH_CHECK();
    return 0;
  }
  return -1;
}

int lcgan_thin_wgrad(const lcgan_tapconv& d, const void* x, const void* g, float* dw, float scale, cudaStream_t s) {
  const bool xf = d.x_dtype == LCGAN_F32, gf = d.y_dtype == LCGAN_F32;
  const int64_t rows = (int64_t)d.N * d.MH * d.MW;
  if (is_point(d) && d.N <= 32) {
    const int64_t total = (int64_t)d.Cout * ((d.Cin + 3) / 4);
    const int grid = grid_cap((total + kThreads - 1) / kThreads, 32);
#define SW(TXT, TGT) skinny_wgrad_kernel<TXT, TGT><<<grid, kThreads, 0, s>>>(d, (const TXT*)x, (const TGT*)g, dw, scale)
    if (xf && gf) SW(float, float); else if (xf) SW(float, bf16); else if (gf) SW(bf16, float); else SW(bf16, bf16);
#undef SW
    LCGAN_LAUNCH_CHECK();
    return 0;
  }
  const bool thin_g = d.Cout <= kMaxThin, thin_x = d.Cin <= kMaxThin;
  if (!thin_g && !thin_x) return -1;
  // the wide side must be channel-innermost with 16-byte vectors
  const bool wide_is_x = thin_g;
  const int Cw = wide_is_x ? d.Cin : d.Cout;
  const bool wf = wide_is_x ? xf : gf;
  const int vec = wf ? 4 : 8;
  const bool ok = wide_is_x ? (dense_inner(d.xs_c, d.xs_w, d.xs_h, d.xs_n, d.Cin, vec) && (uintptr_t)x % 16 == 0)
                            : (dense_inner(d.ys_c, d.ys_w, d.ys_h, d.ys_n, d.Cout, vec) && (uintptr_t)g % 16 == 0);
  if (!ok) return -1;
  const int cv = Cw / vec;
  const int per_lane = cv * d.ntaps;
  if (per_lane > kThreads) return -1;
  int lanes = 1;
  while (lanes * 2 * per_lane <= kThreads) lanes *= 2;     // power of two (tree reduction in smem)
  int64_t blocks = 148LL * 4;
  int64_t rpb = (rows + blocks - 1) / blocks;
  if (rpb < 8LL * lanes) rpb = 8LL * lanes;
  blocks = (rows + rpb - 1) / rpb;
#define TWG(TWIDE, TTHIN, VV, THIN_G)                                                                        \
  thin_wgrad_kernel<TWIDE, TTHIN, VV, THIN_G><<<(int)blocks, kThreads, 0, s>>>(                              \
      d, (const TWIDE*)(THIN_G ? x : g), (const TTHIN*)(THIN_G ? g : x), dw, scale, cv, lanes, rpb, det_sems())
  if (thin_g) {
    if (xf && gf) TWG(float, float, 4, true); else if (xf) TWG(float, bf16, 4, true);
    else if (gf) TWG(bf16, float, 8, true); else TWG(bf16, bf16, 8, true);
  } else {
    if (gf && xf) TWG(float, float, 4, false); else if (gf) TWG(float, bf16, 4, false);
    else if (xf) TWG(bf16, float, 8, false); else TWG(bf16, bf16, 8, false);
  }
#undef TWG
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// Fused x2 transposed conv with <= 4 output channels (all phases).  d describes phase (0,0) of the
// plan: N, IH, IW, Cin, Cout, strides, dtypes, w_ld and the epilogue constants are used.
extern "C" int lcgan_tapconv_up2_thin_eligible(const lcgan_tapconv* d) {
  if (!d || d->Cout > kMaxThin || d->os != 2 || d->is != 1 || d->noise) return 0;
  if (9 * d->Cout * d->Cin > kSmemFloats) return 0;
  if (d->OH != 2 * d->IH || d->OW != 2 * d->IW) return 0;
  if ((int64_t)d->N * d->IH * d->IW >= (1LL << 31) - 65536) return 0;   // 32-bit pixel index in the kernel
  const int vec = d->x_dtype == LCGAN_F32 ? 4 : 8;
  if (d->x_dtype != LCGAN_F32 && d->x_dtype != LCGAN_BF16) return 0;
  return dense_inner(d->xs_c, d->xs_w, d->xs_h, d->xs_n, d->Cin, vec) ? 1 : 0;
}

extern "C" int lcgan_tapconv_up2_thin(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                                      const float* rowscale, const float* bias, void* stream) {
  LCGAN_CHECK(lcgan_tapconv_up2_thin_eligible(d), "tapconv_up2_thin: descriptor not eligible");
  LCGAN_CHECK(x && w2 && y && (uintptr_t)x % 16 == 0, "tapconv_up2_thin: bad pointers");
  cudaStream_t s = (cudaStream_t)stream;
  const bool xf = d->x_dtype == LCGAN_F32, yf = d->y_dtype == LCGAN_F32;
  const int64_t rows = (int64_t)d->N * d->IH * d->IW;
  const size_t wbytes = (size_t)9 * d->Cout * d->Cin * sizeof(float);
#define TU(TXT, TYT, VV)                                                                                     \
  do {                                                                                                       \
    const int G = pow2_group(d->Cin / VV);                                                                   \
    const int grid = grid_cap((rows + kThreads / G - 1) / (kThreads / G), 16);                               \
    switch (d->Cout) {                                                                                       \
      case 1: thin_up2_kernel<TXT, TYT, VV, 1><<<grid, kThreads, wbytes, s>>>(*d, (const TXT*)x, w2, (TYT*)y, rowscale, bias, G); break; \
      case 2: thin_up2_kernel<TXT, TYT, VV, 2><<<grid, kThreads, wbytes, s>>>(*d, (const TXT*)x, w2, (TYT*)y, rowscale, bias, G); break; \
      case 3: thin_up2_kernel<TXT, TYT, VV, 3><<<grid, kThreads, wbytes, s>>>(*d, (const TXT*)x, w2, (TYT*)y, rowscale, bias, G); break; \
      default: thin_up2_kernel<TXT, TYT, VV, 4><<<grid, kThreads, wbytes, s>>>(*d, (const TXT*)x, w2, (TYT*)y, rowscale, bias, G); break; \
    }                                                                                                        \
  } while (0)
  if (xf && yf) TU(float, float, 4); else if (xf) TU(float, bf16, 4); else if (yf) TU(bf16, float, 8); else TU(bf16, bf16, 8);
#undef TU
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_tapconv_up2_thin_wgrad(const lcgan_tapconv* d, const void* x, const void* g, float* dw2,
                                            float scale, void* stream) {
  // d as for lcgan_tapconv_up2_thin (y_* describe G [N, 2H, 2W, Cout]); dw2 [Cout][9*Cin] f32, accumulated
  LCGAN_CHECK(lcgan_tapconv_up2_thin_eligible(d), "tapconv_up2_thin_wgrad: descriptor not eligible");
  LCGAN_CHECK(x && g && dw2 && (uintptr_t)x % 16 == 0, "tapconv_up2_thin_wgrad: bad pointers");
  const int cv4 = d->Cin / 4;
  LCGAN_CHECK(d->Cin % 4 == 0 && cv4 <= kThreads && kThreads % cv4 == 0 && d->Cout <= 2,
              "tapconv_up2_thin_wgrad: needs Cin = 4 * (a divisor of %d) and Cout <= 2", kThreads);
  cudaStream_t s = (cudaStream_t)stream;
  const bool xf = d->x_dtype == LCGAN_F32, gf = d->y_dtype == LCGAN_F32;
  const int64_t rows = (int64_t)d->N * d->IH * d->IW;
  const int lanes = kThreads / cv4;
  int64_t blocks = 148LL * 4;
  int64_t rpb = (rows + blocks - 1) / blocks;
  if (rpb < 8LL * lanes) rpb = 8LL * lanes;
  blocks = (rows + rpb - 1) / rpb;
#define UW(TXT, TGT)                                                                                         \
  do {                                                                                                       \
    if (d->Cout == 1)                                                                                        \
      thin_up2_wgrad_kernel<TXT, TGT, 1><<<(int)blocks, kThreads, 0, s>>>(*d, (const TXT*)x, (const TGT*)g, dw2, scale, cv4, lanes, (uint32_t)rpb, det_sems()); \
    else                                                                                                     \
      thin_up2_wgrad_kernel<TXT, TGT, 2><<<(int)blocks, kThreads, 0, s>>>(*d, (const TXT*)x, (const TGT*)g, dw2, scale, cv4, lanes, (uint32_t)rpb, det_sems()); \
  } while (0)
  if (xf && gf) UW(float, float); else if (xf) UW(float, bf16); else if (gf) UW(bf16, float); else UW(bf16, bf16);
#undef UW
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_pw_wgrad32(const lcgan_tapconv* d, const void* x, const void* g, float* dwp, void* stream) {
  LCGAN_CHECK(d && x && g && dwp, "pw_wgrad32: null argument");
  LCGAN_CHECK(is_pointwise(*d) && d->Cin == 32 && d->Cout >= 1 && d->Cout <= kMaxThin && d->x_dtype == LCGAN_BF16 &&
              (uintptr_t)x % 16 == 0 && d->xs_c == 1 && d->xs_w == 32 && (d->MH == 1 || d->xs_h == (int64_t)d->MW * 32) &&
              (d->N == 1 || d->xs_n == (int64_t)d->MH * d->MW * 32) && d->N <= 65535,
              "pw_wgrad32: needs a 1x1 layer, 32 dense channels-last bf16 input channels, Cout <= %d", kMaxThin);
  const int64_t hw = (int64_t)d->MH * d->MW;
  int gx = (int)((hw + kThreads * 16 - 1) / (kThreads * 16));       // >= 16 pixels per thread
  const int cap = (148 * 4 + d->N - 1) / d->N;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (d->y_dtype == LCGAN_F32) pw_wgrad32_kernel<float, kMaxThin><<<dim3(gx, d->N), kThreads, 0, s>>>(*d, (const bf16*)x, (const float*)g, dwp);
  else pw_wgrad32_kernel<bf16, kMaxThin><<<dim3(gx, d->N), kThreads, 0, s>>>(*d, (const bf16*)x, (const bf16*)g, dwp);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Flow-layer weight gradient on the tensor cores.  dW[o][t][c] = sum g[b, 2m-1+ki, 2n-1+kj, o] x[b,m,n,c] has only
// 18 = 9 taps x 2 channels columns on the gradient side, so the gradient is first gathered into a bf16
// channels-last tensor G18[b,m,n, t*2+o] (32 channels, the last 14 zero; out-of-range taps zero) - one thread per
// input-lattice pixel, 9 float2 reads, one 64-byte write - and the contraction over pixels is then an ordinary
// pointwise weight gradient (lcgan_tapconv_wgrad_tc: X[px x Cin]^T G18[px x 32]).  The CUDA-core all-tap kernel
// it replaces was issue-bound at 9 % of HBM (2.25 ms for 64 -> 2 @512^2 -> 1024^2, batch 32).
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(kThreads)
flow_grad_im2col_kernel(const float* __restrict__ g, bf16* __restrict__ out, int N, int H, int W) {
  const int64_t npix = (int64_t)N * H * W;
  const int OH = 2 * H, OW = 2 * W;
  for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < npix; p += (int64_t)gridDim.x * kThreads) {
    const int n = (int)(p % W);
    const int m = (int)((p / W) % H);
    const int b = (int)(p / ((int64_t)W * H));
    const float2* gb = reinterpret_cast<const float2*>(g) + (int64_t)b * OH * OW;
    float f[32];
#pragma unroll
    for (int i = 18; i < 32; ++i) f[i] = 0.f;
#pragma unroll
    for (int ki = 0; ki < 3; ++ki) {
      const int oy = 2 * m - 1 + ki;
#pragma unroll
      for (int kj = 0; kj < 3; ++kj) {
        const int ox = 2 * n - 1 + kj;
        const bool ok = oy >= 0 && oy < OH && ox >= 0 && ox < OW;
        const float2 v = ok ? gb[(int64_t)oy * OW + ox] : make_float2(0.f, 0.f);
        f[(ki * 3 + kj) * 2] = v.x;
        f[(ki * 3 + kj) * 2 + 1] = v.y;
      }
    }
    uint4* op = reinterpret_cast<uint4*>(out + p * 32);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      Vec16<bf16> o;
      o.pack(f + 8 * v);
      op[v] = o.v;
    }
  }
}
}  // namespace

extern "C" int lcgan_flow_grad_im2col(const float* g, void* out, int N, int H, int W, void* stream) {
  LCGAN_CHECK(g && out && N > 0 && H > 0 && W > 0 && (uintptr_t)g % 8 == 0 && (uintptr_t)out % 16 == 0,
              "flow_grad_im2col: bad arguments");
  const int64_t npix = (int64_t)N * H * W;
  flow_grad_im2col_kernel<<<grid_cap((npix + kThreads - 1) / kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(
      g, (bf16*)out, N, H, W);
  LCGAN_LAUNCH_CHECK();
  return 0;
}
