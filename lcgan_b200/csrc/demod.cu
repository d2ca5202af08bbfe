// Weight-/[b,C]-sized pieces of the modulated convolution (reference custom_layers.py:62-68) and of the
// fused epilogue's backward, each as ONE launch instead of a chain of elementwise torch kernels:
//   demod_fwd    d[b,o]  = rsqrt(sum_c s[b,c]^2 Wsq[o,c] + eps)          (Wsq = sum_k (w c)^2, optim.cu)
//   demod_bwd_s  ds[b,c] = 2 s[b,c] sum_o dq[b,o] Wsq[o,c],   dq = -0.5 dd d^3
//   demod_bwd_w  dw[o,c,k] = 2 c^2 q(w[o,c,k]) sum_b dq[b,o] s[b,c]^2    (q = rounding to the conv's dtype)
//   epilogue_grads  db[o] = bias_scale sum_b r0[b,o];  dd[b,o] = (r1[b,o] - bias[o] bias_scale r0[b,o]) / d[b,o]
// These are latency-, not bandwidth-bound (<= 512x512 tables): they exist to keep the per-iteration launch
// count down (the batch-independent floor of the data-parallel step).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
demod_fwd_kernel(const float* __restrict__ s, const float* __restrict__ wsq, float* __restrict__ d, int B, int O,
                 int I, float eps) {
  const int warp = blockIdx.x * (kThreads / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp >= B * O) return;
  const int b = warp / O, o = warp - b * O;
  const float* sr = s + (int64_t)b * I;
  const float* wr = wsq + (int64_t)o * I;
  float acc = 0.f;
  for (int c = lane; c < I; c += 32) acc = fmaf(sr[c] * sr[c], wr[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) d[warp] = rsqrtf(acc + eps);
}

__global__ void __launch_bounds__(kThreads)
demod_bwd_s_kernel(const float* __restrict__ dd, const float* __restrict__ d, const float* __restrict__ s,
                   const float* __restrict__ wsq, float* __restrict__ ds, int B, int O, int I) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * kThreads + threadIdx.x;
  __shared__ float dq[512];
  float acc = 0.f;
  for (int o0 = 0; o0 < O; o0 += 512) {
    __syncthreads();
    for (int o = threadIdx.x; o < 512 && o0 + o < O; o += kThreads) {
      const float dv = d[(int64_t)b * O + o0 + o];
      dq[o] = -0.5f * dd[(int64_t)b * O + o0 + o] * dv * dv * dv;
    }
    __syncthreads();
    if (c < I) {
      const int n = min(512, O - o0);
      for (int o = 0; o < n; ++o) acc = fmaf(dq[o], wsq[(int64_t)(o0 + o) * I + c], acc);
    }
  }
  if (c < I) ds[(int64_t)b * I + c] = 2.f * s[(int64_t)b * I + c] * acc;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
demod_bwd_w_kernel(const float* __restrict__ dd, const float* __restrict__ d, const float* __restrict__ s,
                   const float* __restrict__ w, float* __restrict__ dw, int B, int O, int I, int K, float c2) {
  const int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (idx >= (int64_t)O * I) return;
  const int o = (int)(idx / I), c = (int)(idx - (int64_t)o * I);
  float t = 0.f;
  for (int b = 0; b < B; ++b) {
    const float dv = d[(int64_t)b * O + o], sv = s[(int64_t)b * I + c];
    t = fmaf(-0.5f * dd[(int64_t)b * O + o] * dv * dv * dv, sv * sv, t);
  }
  t *= 2.f * c2;
  for (int k = 0; k < K; ++k) {
    T q;
    stf(&q, w[idx * K + k]);
    dw[idx * K + k] = ldf(&q) * t;
  }
}

__global__ void __launch_bounds__(kThreads)
epilogue_grads_kernel(const float* __restrict__ r0, const float* __restrict__ r1, const float* __restrict__ bias,
                      const float* __restrict__ d, float bias_scale, float* __restrict__ db, float* __restrict__ dd,
                      int B, int O) {
  const int o = blockIdx.x * kThreads + threadIdx.x;
  if (o >= O) return;
  const float be = bias ? bias[o] * bias_scale : 0.f;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float v0 = r0[(int64_t)b * O + o];
    acc += v0;
    if (dd) dd[(int64_t)b * O + o] = (r1[(int64_t)b * O + o] - be * v0) / d[(int64_t)b * O + o];
  }
  if (db) db[o] = acc * bias_scale;
}

}  // namespace

extern "C" int lcgan_demod_fwd(const float* s, const float* wsq, float* d, int B, int O, int I, float eps, void* stream) {
  LCGAN_CHECK(s && wsq && d && B > 0 && O > 0 && I > 0, "demod_fwd: bad arguments");
  demod_fwd_kernel<<<ceil_div((int64_t)B * O, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(s, wsq, d, B, O, I, eps);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_demod_bwd(const float* dd, const float* d, const float* s, const float* wsq, const float* w,
                               float* ds, float* dw, int B, int O, int I, int K, float wscale, int dt, void* stream) {
  LCGAN_CHECK(dd && d && s && wsq && B > 0 && B <= 65535 && O > 0 && I > 0 && K > 0, "demod_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (ds) demod_bwd_s_kernel<<<dim3(ceil_div(I, kThreads), B), kThreads, 0, st>>>(dd, d, s, wsq, ds, B, O, I);
  if (dw) {
    LCGAN_CHECK(w != nullptr, "demod_bwd: dw needs w");
    const int g = ceil_div((int64_t)O * I, kThreads);
    if (dt == LCGAN_BF16) demod_bwd_w_kernel<bf16><<<g, kThreads, 0, st>>>(dd, d, s, w, dw, B, O, I, K, wscale * wscale);
    else demod_bwd_w_kernel<float><<<g, kThreads, 0, st>>>(dd, d, s, w, dw, B, O, I, K, wscale * wscale);
  }
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_epilogue_grads(const float* r0, const float* r1, const float* bias, const float* d,
                                    float bias_scale, float* db, float* dd, int B, int O, void* stream) {
  LCGAN_CHECK(r0 && B > 0 && O > 0 && (db || dd), "epilogue_grads: bad arguments");
  LCGAN_CHECK(!dd || (r1 && d), "epilogue_grads: dd needs r1 and d");
  epilogue_grads_kernel<<<ceil_div(O, kThreads), kThreads, 0, (cudaStream_t)stream>>>(r0, r1, bias, d, bias_scale, db, dd,
                                                                                    B, O);
  LCGAN_LAUNCH_CHECK();
  return 0;
}
