// Per-sample loss reductions (reference loss.py:9-24, cnn.py:40-41) and the multi-tensor EMA
// (ema.py:26-32).  One warp per sample with shuffle reductions for the [b,256] embedding losses;
// block + atomic reduction for the R1 sum of squares over [b, 3*R*R].
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
__device__ int g_sems[kDetSems];     // deterministic-mode turn semaphores (common.cuh)

__global__ void __launch_bounds__(kThreads)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, int B, int D) {
  const int row = blockIdx.x * (kThreads / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= B) return;
  const float* xr = x + (int64_t)row * D;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) ss = fmaf(xr[i], xr[i], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int i = lane; i < D; i += 32) y[(int64_t)row * D + i] = xr[i] * inv;
  if (lane == 0) inv_norm[row] = inv;
}

__global__ void __launch_bounds__(kThreads)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv_norm, const float* __restrict__ dy,
                  float* __restrict__ dx, int B, int D) {
  const int row = blockIdx.x * (kThreads / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= B) return;
  const float* yr = y + (int64_t)row * D;
  const float* gr = dy + (int64_t)row * D;
  float dot = 0.f;
  for (int i = lane; i < D; i += 32) dot = fmaf(yr[i], gr[i], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  for (int i = lane; i < D; i += 32) dx[(int64_t)row * D + i] = (gr[i] - yr[i] * dot) * inv;
}

__global__ void __launch_bounds__(kThreads)
contrastive_fwd_kernel(const float* __restrict__ a, const float* __restrict__ p, const float* __restrict__ n,
                       float* __restrict__ l, float* __restrict__ sig, int B, int D, float inv_tau) {
  const int row = blockIdx.x * (kThreads / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= B) return;
  const int64_t o = (int64_t)row * D;
  float ap = 0.f, an = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float av = a[o + i];
    ap = fmaf(av, p[o + i], ap);
    an = fmaf(av, n[o + i], an);
  }
  ap = warp_sum(ap); an = warp_sum(an);
  if (lane == 0) {
    const float u = (an - ap) * inv_tau;
    l[row] = fmaxf(u, 0.f) + log1pf(expf(-fabsf(u)));       // softplus(u) = -log(e^p/(e^p+e^n))
    sig[row] = 1.f / (1.f + expf(-u));
  }
}

__global__ void __launch_bounds__(kThreads)
contrastive_bwd_kernel(const float* __restrict__ a, const float* __restrict__ p, const float* __restrict__ n,
                       const float* __restrict__ sig, const float* __restrict__ dl, float* __restrict__ da,
                       float* __restrict__ dp, float* __restrict__ dn, int B, int D, float inv_tau) {
  const int64_t total = (int64_t)B * D;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int row = (int)(i / D);
    const float k = dl[row] * sig[row] * inv_tau;
    const float av = a[i];
    da[i] = k * (n[i] - p[i]);
    dp[i] = -k * av;
    dn[i] = k * av;
  }
}

__global__ void __launch_bounds__(kThreads)
sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t L, int* sems) {
  const int b = blockIdx.y;
  const float* xr = x + (int64_t)b * L;
  float acc = 0.f;
  const int64_t L4 = ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) ? L / 4 : 0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < L4; i += (int64_t)gridDim.x * kThreads) {
    const float4 v = reinterpret_cast<const float4*>(xr)[i];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  for (int64_t i = L4 * 4 + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < L; i += (int64_t)gridDim.x * kThreads)
    acc = fmaf(xr[i], xr[i], acc);
  acc = warp_sum(acc);
  __shared__ float part[kThreads / 32];
  if (threadIdx.x % 32 == 0) part[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      if (sems) {                                   // deterministic mode: blocks of a row add in order
        det_wait_turn(sems + b, blockIdx.x);
        det_add(out + b, v);
        __threadfence();
        det_pass_turn(sems + b, blockIdx.x, gridDim.x);
      } else {
        atomicAdd(out + b, v);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
rowscale_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, int64_t L) {
  const int b = blockIdx.y;
  const float k = s[b];
  const float* xr = x + (int64_t)b * L;
  float* yr = y + (int64_t)b * L;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < L; i += (int64_t)gridDim.x * kThreads)
    yr[i] = xr[i] * k;
}

__global__ void __launch_bounds__(kThreads)
ema_lerp_kernel(float* const* __restrict__ dst, const float* const* __restrict__ src,
                const int64_t* __restrict__ numel, float decay, const float* __restrict__ decay_dev) {
  if (decay_dev) decay = *decay_dev;
  float* d = dst[blockIdx.y];
  const float* s = src[blockIdx.y];
  const int64_t n = numel[blockIdx.y];
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float sv = s[i];
    d[i] = fmaf(decay, d[i] - sv, sv);   // src.lerp(dst, decay)
  }
}

}  // namespace

extern "C" int lcgan_l2norm_fwd(const float* x, float* y, float* inv_norm, int B, int D, void* stream) {
  LCGAN_CHECK(x && y && inv_norm && B > 0 && D > 0, "l2norm_fwd: bad arguments");
  l2norm_fwd_kernel<<<ceil_div(B, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(x, y, inv_norm, B, D);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int B, int D,
                                void* stream) {
  LCGAN_CHECK(y && inv_norm && dy && dx && B > 0 && D > 0, "l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<ceil_div(B, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(y, inv_norm, dy, dx, B, D);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_contrastive_fwd(const float* a, const float* p, const float* n, float* l, float* sig, int B,
                                     int D, float tau, void* stream) {
  LCGAN_CHECK(a && p && n && l && sig && B > 0 && D > 0 && tau > 0.f, "contrastive_fwd: bad arguments");
  contrastive_fwd_kernel<<<ceil_div(B, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(a, p, n, l, sig, B, D,
                                                                                           1.f / tau);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_contrastive_bwd(const float* a, const float* p, const float* n, const float* sig,
                                     const float* dl, float* da, float* dp, float* dn, int B, int D, float tau,
                                     void* stream) {
  LCGAN_CHECK(a && p && n && sig && dl && da && dp && dn && B > 0 && D > 0 && tau > 0.f,
              "contrastive_bwd: bad arguments");
  contrastive_bwd_kernel<<<ceil_div((int64_t)B * D, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      a, p, n, sig, dl, da, dp, dn, B, D, 1.f / tau);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_sumsq(const float* x, float* out, int B, int64_t L, void* stream) {
  LCGAN_CHECK(x && out && B > 0 && B <= 65535 && L > 0, "sumsq: bad arguments");
  int bx = ceil_div(L, (int64_t)kThreads * 16);
  const int cap = ceil_div(148 * 8, B);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  int* sems = nullptr;
  if (lcgan_det_enabled()) {
    LCGAN_CHECK(B <= kDetSems, "sumsq: batch too large for deterministic mode");
    LCGAN_CUDA(cudaGetSymbolAddress((void**)&sems, g_sems));
  }
  sumsq_kernel<<<dim3(bx, B), kThreads, 0, (cudaStream_t)stream>>>(x, out, L, sems);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_rowscale(const float* x, const float* s, float* y, int B, int64_t L, void* stream) {
  LCGAN_CHECK(x && s && y && B > 0 && B <= 65535 && L > 0, "rowscale: bad arguments");
  int bx = ceil_div(L, (int64_t)kThreads * 8);
  const int cap = ceil_div(148 * 16, B);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  rowscale_kernel<<<dim3(bx, B), kThreads, 0, (cudaStream_t)stream>>>(x, s, y, L);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_ema_lerp(float* const* dst, const float* const* src, const int64_t* numel, int n, float decay,
                              const float* decay_dev, void* stream) {
  LCGAN_CHECK(dst && src && numel && n > 0 && n <= 65535, "ema_lerp: bad arguments");
  ema_lerp_kernel<<<dim3(16, n), kThreads, 0, (cudaStream_t)stream>>>(dst, src, numel, decay, decay_dev);
  LCGAN_LAUNCH_CHECK();
  return 0;
}
