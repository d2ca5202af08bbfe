// Generic tap-convolution on CUDA cores: any strides, any dtype mix, fp32 accumulate.
// This is the universal path (fp32-accurate mode, K=3 / N=2,3 edge layers, odd channel counts);
// the tensor-core path in conv_tc.cu takes the GEMM-shaped bf16 layers.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int BK = 16;

template <typename TX, typename TW, typename TY, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(kThreads)
tapconv_simt_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const TW* __restrict__ w,
                    TY* __restrict__ y, const float* __restrict__ rowscale,
                    const float* __restrict__ bias, const TY* __restrict__ residual) {
  constexpr int TXN = BN / TN;   // threads along output channels
  constexpr int TYN = BM / TM;   // threads along lattice rows
  static_assert(TXN * TYN == kThreads, "tile/thread mismatch");
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ int rb[BM], rm[BM], rn[BM];

  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int o0 = blockIdx.y * BN;
  const int64_t rows = (int64_t)d.N * d.MH * d.MW;
  for (int r = tid; r < BM; r += kThreads) {
    int64_t g = row0 + r;
    if (g < rows) {
      rn[r] = (int)(g % d.MW);
      int64_t q = g / d.MW;
      rm[r] = (int)(q % d.MH);
      rb[r] = (int)(q / d.MH);
    } else {
      rb[r] = -1; rm[r] = 0; rn[r] = 0;
    }
  }
  __syncthreads();

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int ty = tid / TXN, tx = tid % TXN;

  for (int t = 0; t < d.ntaps; ++t) {
    const int ddy = d.dy[t], ddx = d.dx[t];
    const int64_t wcol0 = (int64_t)d.wtap[t] * d.Cin;
    for (int c0 = 0; c0 < d.Cin; c0 += BK) {
      for (int i = tid; i < BM * BK; i += kThreads) {
        const int k = i % BK, r = i / BK;
        const int b = rb[r], c = c0 + k;
        float v = 0.f;
        if (b >= 0 && c < d.Cin) {
          const int iy = rm[r] * d.is + ddy, ix = rn[r] * d.is + ddx;
          if (iy >= 0 && iy < d.IH && ix >= 0 && ix < d.IW)
            v = ldf(x + b * d.xs_n + iy * d.xs_h + ix * d.xs_w + c * d.xs_c);
        }
        As[k][r] = v;
      }
      for (int i = tid; i < BN * BK; i += kThreads) {
        const int k = i % BK, o = i / BK;
        const int c = c0 + k;
        float v = 0.f;
        if (o0 + o < d.Cout && c < d.Cin) v = ldf(w + (int64_t)(o0 + o) * d.w_ld + wcol0 + c);
        Bs[k][o] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty + i * TYN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + j * TXN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty + i * TYN;
    const int b = rb[r];
    if (b < 0) continue;
    const int oy = rm[r] * d.os + d.py, ox = rn[r] * d.os + d.px;
    const int64_t base = b * d.ys_n + oy * d.ys_h + ox * d.ys_w;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int o = o0 + tx + j * TXN;
      if (o >= d.Cout) continue;
      float v = acc[i][j] * d.acc_scale;
      if (rowscale) v *= rowscale[(int64_t)b * d.Cout + o];
      if (bias) v += bias[o] * d.bias_scale;
      if (d.noise) v += d.noise[(int64_t)oy * d.OW + ox] * d.noise_scale;
      v = (v > 0.f ? v : v * d.slope) * d.gain;
      if (residual) v += ldf(residual + base + o * d.ys_c);
      stf(y + base + o * d.ys_c, v);
    }
  }
}

__device__ int g_sems[kDetSems];   // deterministic-mode turn semaphores (self-resetting)

template <typename TX, typename TG>
__global__ void __launch_bounds__(kThreads)
tapconv_wgrad_simt_kernel(const lcgan_tapconv d, const TX* __restrict__ x, const TG* __restrict__ g,
                          float* __restrict__ dw, float scale, int64_t rows_per_split, int ctiles, int* sems) {
  constexpr int BO = 64, BC = 64;
  __shared__ float Gs[BK][BO + 1];
  __shared__ float Xs[BK][BC + 1];
  __shared__ int kb[BK], km[BK], kn[BK];
  const int tid = threadIdx.x;
  const int o0 = (blockIdx.x / ctiles) * BO, c0 = (blockIdx.x % ctiles) * BC;
  const int t = blockIdx.y;
  const int ddy = d.dy[t], ddx = d.dx[t];
  const int64_t rows = (int64_t)d.N * d.MH * d.MW;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(rows, r_begin + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid / 16, tx = tid % 16;

  for (int64_t k0 = r_begin; k0 < r_end; k0 += BK) {
    if (tid < BK) {
      int64_t r = k0 + tid;
      if (r < r_end) {
        kn[tid] = (int)(r % d.MW);
        int64_t q = r / d.MW;
        km[tid] = (int)(q % d.MH);
        kb[tid] = (int)(q / d.MH);
      } else {
        kb[tid] = -1; km[tid] = 0; kn[tid] = 0;
      }
    }
    __syncthreads();
    for (int i = tid; i < BO * BK; i += kThreads) {
      const int o = i % BO, k = i / BO;
      const int b = kb[k];
      float v = 0.f;
      if (b >= 0 && o0 + o < d.Cout) {
        const int oy = km[k] * d.os + d.py, ox = kn[k] * d.os + d.px;
        v = ldf(g + b * d.ys_n + oy * d.ys_h + ox * d.ys_w + (o0 + o) * d.ys_c);
      }
      Gs[k][o] = v;
    }
    for (int i = tid; i < BC * BK; i += kThreads) {
      const int c = i % BC, k = i / BC;
      const int b = kb[k];
      float v = 0.f;
      if (b >= 0 && c0 + c < d.Cin) {
        const int iy = km[k] * d.is + ddy, ix = kn[k] * d.is + ddx;
        if (iy >= 0 && iy < d.IH && ix >= 0 && ix < d.IW)
          v = ldf(x + b * d.xs_n + iy * d.xs_h + ix * d.xs_w + (c0 + c) * d.xs_c);
      }
      Xs[k][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Gs[k][ty + i * 16];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[k][tx + j * 16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int64_t wcol0 = (int64_t)d.wtap[t] * d.Cin;
  // deterministic mode: the K-splits of this (tile, tap) add in split order (common.cuh)
  int* sem = sems ? sems + blockIdx.y * gridDim.x + blockIdx.x : nullptr;
  if (sem) det_block_begin(sem, blockIdx.z);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + ty + i * 16;
    if (o >= d.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx + j * 16;
      if (c >= d.Cin) continue;
      float* dst = dw + (int64_t)o * d.w_ld + wcol0 + c;
      if (sem) det_add(dst, acc[i][j] * scale);
      else atomicAdd(dst, acc[i][j] * scale);
    }
  }
  if (sem) det_block_end(sem, blockIdx.z, gridDim.z);
}

template <typename TX, typename TW, typename TY>
int launch_fwd(const lcgan_tapconv& d, const void* x, const void* w, void* y, const float* rowscale,
               const float* bias, const void* residual, cudaStream_t s) {
  const int64_t rows = (int64_t)d.N * d.MH * d.MW;
  if (d.Cout <= 8) {
    dim3 grid(ceil_div(rows, 128), ceil_div(d.Cout, 8));
    tapconv_simt_kernel<TX, TW, TY, 128, 8, 4, 1><<<grid, kThreads, 0, s>>>(
        d, (const TX*)x, (const TW*)w, (TY*)y, rowscale, bias, (const TY*)residual);
  } else {
    dim3 grid(ceil_div(rows, 64), ceil_div(d.Cout, 64));
    tapconv_simt_kernel<TX, TW, TY, 64, 64, 4, 4><<<grid, kThreads, 0, s>>>(
        d, (const TX*)x, (const TW*)w, (TY*)y, rowscale, bias, (const TY*)residual);
  }
  LCGAN_LAUNCH_CHECK();
  return 0;
}

template <typename TX, typename TW>
int dispatch_y(const lcgan_tapconv& d, const void* x, const void* w, void* y, const float* rs,
               const float* bias, const void* res, cudaStream_t s) {
  if (d.y_dtype == LCGAN_F32) return launch_fwd<TX, TW, float>(d, x, w, y, rs, bias, res, s);
  return launch_fwd<TX, TW, bf16>(d, x, w, y, rs, bias, res, s);
}

template <typename TX>
int dispatch_w(const lcgan_tapconv& d, const void* x, const void* w, void* y, const float* rs,
               const float* bias, const void* res, cudaStream_t s) {
  if (d.w_dtype == LCGAN_F32) return dispatch_y<TX, float>(d, x, w, y, rs, bias, res, s);
  return dispatch_y<TX, bf16>(d, x, w, y, rs, bias, res, s);
}

int check_desc(const lcgan_tapconv* d) {
  LCGAN_CHECK(d != nullptr, "tapconv: null descriptor");
  LCGAN_CHECK(d->ntaps >= 1 && d->ntaps <= LCGAN_MAX_TAPS, "tapconv: ntaps=%d out of range", d->ntaps);
  LCGAN_CHECK(d->N > 0 && d->MH > 0 && d->MW > 0 && d->Cin > 0 && d->Cout > 0, "tapconv: empty dims");
  LCGAN_CHECK(d->os >= 1 && d->is >= 1, "tapconv: bad lattice strides");
  LCGAN_CHECK((d->MH - 1) * d->os + d->py < d->OH && (d->MW - 1) * d->os + d->px < d->OW,
              "tapconv: lattice exceeds output extent");
  LCGAN_CHECK(d->x_dtype >= 0 && d->x_dtype <= 1 && d->y_dtype >= 0 && d->y_dtype <= 1 &&
              d->w_dtype >= 0 && d->w_dtype <= 1, "tapconv: bad dtype code");
  return 0;
}

}  // namespace

extern "C" int lcgan_tapconv_simt(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                                  const float* rowscale, const float* bias, const void* residual,
                                  void* stream) {
  if (int e = check_desc(d)) return e;
  LCGAN_CHECK(x && w2 && y, "tapconv_simt: null tensor pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (!d->noise) {                               // the special-shape kernels have no noise term
    const int e = lcgan_thin_forward(*d, x, w2, y, rowscale, bias, residual, s);
    if (e >= 0) return e;
  }
  LCGAN_CHECK(d->colscale == nullptr, "tapconv_simt: colscale is only implemented by the pointwise thin kernel");
  if (d->x_dtype == LCGAN_F32) return dispatch_w<float>(*d, x, w2, y, rowscale, bias, residual, s);
  return dispatch_w<bf16>(*d, x, w2, y, rowscale, bias, residual, s);
}

extern "C" int lcgan_tapconv_wgrad_simt(const lcgan_tapconv* d, const void* x, const void* g,
                                        float* dw2, float scale, void* stream) {
  if (int e = check_desc(d)) return e;
  LCGAN_CHECK(x && g && dw2, "tapconv_wgrad_simt: null tensor pointer");
  cudaStream_t s = (cudaStream_t)stream;
  {
    const int e = lcgan_thin_wgrad(*d, x, g, dw2, scale, s);
    if (e >= 0) return e;
  }
  const int64_t rows = (int64_t)d->N * d->MH * d->MW;
  const int otiles = ceil_div(d->Cout, 64), ctiles = ceil_div(d->Cin, 64);
  const int base_blocks = otiles * ctiles * d->ntaps;
  int splits = (4 * 148 + base_blocks - 1) / base_blocks;
  const int64_t max_splits = (rows + 255) / 256;   // at least 256 rows per split
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rps = (rows + splits - 1) / splits;
  rps = (rps + BK - 1) / BK * BK;
  splits = (int)((rows + rps - 1) / rps);
  dim3 grid(otiles * ctiles, d->ntaps, splits);
  int* sems = nullptr;
  if (lcgan_det_enabled()) {
    LCGAN_CHECK(otiles * ctiles * d->ntaps <= kDetSems, "tapconv_wgrad_simt: too many tiles for deterministic mode");
    LCGAN_CUDA(cudaGetSymbolAddress((void**)&sems, g_sems));
  }
#define WG(TXT, TGT)                                                                     \
  tapconv_wgrad_simt_kernel<TXT, TGT><<<grid, kThreads, 0, s>>>(*d, (const TXT*)x, (const TGT*)g, \
                                                                dw2, scale, rps, ctiles, sems)
  if (d->x_dtype == LCGAN_F32 && d->y_dtype == LCGAN_F32) WG(float, float);
  else if (d->x_dtype == LCGAN_F32) WG(float, bf16);
  else if (d->y_dtype == LCGAN_F32) WG(bf16, float);
  else WG(bf16, bf16);
#undef WG
  LCGAN_LAUNCH_CHECK();
  return 0;
}
