// Multi-tensor optimizer step (reference worker.py:98-110: torch.optim.Adam(betas=(0, 0.99), eps=1e-8) over
// 165 generator / 60 discriminator tensors) and multi-tensor weight packing.
//
// Up to LCGAN_MT_MAX tensors per launch; their pointers travel BY VALUE in the kernel parameters, so
// there is no device-side pointer table to build, and a CUDA graph capture records them like any other
// argument.  blockIdx.y = tensor, blockIdx.x strides over its elements (16-byte vectors when aligned).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBlocksPerTensor = 64;

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), per element:
//   m = lerp(m, g, 1-b1);  v = v*b2 + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// with bc1 = 1-b1^t, bc2 = 1-b2^t, t = this tensor's own step count (a tensor whose gradient was None
// in an iteration is skipped by torch and keeps its count - the caller simply leaves it out).
__global__ void __launch_bounds__(kThreads)
adam_kernel(const lcgan_adam_chunk ch, float lr, float beta1, float beta2, float eps) {
  const int k = blockIdx.y;
  float* __restrict__ p = ch.p[k];
  const float* __restrict__ g = ch.g[k];
  float* __restrict__ m = ch.m[k];
  float* __restrict__ v = ch.v[k];
  const int64_t n = ch.numel[k];
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*ch.step[k]) + 1.0;
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    sh[0] = (float)((double)lr / bc1);
    sh[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = sh[0], bc2_sqrt = sh[1];
  const float om_b1 = 1.f - beta1, om_b2 = 1.f - beta2;
  auto upd = [&](float& pv, float gv, float& mv, float& vv) {
    mv = beta1 == 0.f ? gv : (om_b1 < 0.5f ? mv + om_b1 * (gv - mv) : gv - (gv - mv) * (1.f - om_b1));
    vv = fmaf(om_b2 * gv, gv, vv * beta2);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = fmaf(-step_size, mv / denom, pv);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(v) |
                     (m ? reinterpret_cast<uintptr_t>(m) : 0)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 mv = m ? reinterpret_cast<float4*>(m)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upd(pv.x, gv.x, mv.x, vv.x); upd(pv.y, gv.y, mv.y, vv.y); upd(pv.z, gv.z, mv.z, vv.z); upd(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (m) reinterpret_cast<float4*>(m)[i] = mv;
  }
  for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    float pv = p[i], vv = v[i], mv = m ? m[i] : 0.f;
    upd(pv, g[i], mv, vv);
    p[i] = pv; v[i] = vv;
    if (m) m[i] = mv;
  }
}

__global__ void adam_count_kernel(const lcgan_adam_chunk ch) {
  const int k = threadIdx.x;
  if (k < ch.count) *ch.step[k] += 1.f;
}

// Weight packs of the tap convolution (ops.pack_weight): for a parameter w[O][I][K] (K = kh*kw, fp32)
//   mode 0: dst[o][k*I + c] = w[o][c][k]      (forward / weight-gradient layout)
//   mode 1: dst[c][k*O + o] = w[o][c][k]      (data-gradient layout)
//   mode 2: dst[o][c] = sum_k (q(w[o][c][k]) * scale)^2   (f32; q = rounding to the pack dtype)  - the
//           demodulation table Wsq of custom_layers.py:65-67
//   mode 3: the [4*O][4*I] phase-major / tap-major weight of the fused x2 transposed conv (K = 9)
// Modes 0 / 1 go through a shared-memory tile of 32 output channels x 32 input channels x K: the source rows
// (32*K contiguous floats) are read coalesced, and the destination is written in runs of 32 contiguous elements
// (c-fastest for mode 0, o-fastest for mode 1).  (A destination-ordered gather read the 260 MB of discriminator
// weights at an 8x sector amplification: 2.4 ms per refresh, 6 % of the batch-4 iteration.)
constexpr int kPT = 32;                       // tile edge (channels)
constexpr int kPackMaxK = 9;

template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_kernel(const lcgan_pack_chunk ch) {
  const int j = blockIdx.y;
  const float* __restrict__ w = ch.src[j];
  const int O = ch.O[j], I = ch.I[j], K = ch.K[j], mode = ch.mode[j];
  if (mode == 3) {
    // fused x2 transposed conv (lcgan_tapconv_tc_blocked): dst[(py*2+px)*O + o][(dy*2+dx)*I + c] =
    // w[o][c][ki][kj], ki = py+1-2dy, kj = px+1-2dx, zero where a kernel index falls outside 0..2 (K = 9)
    T* dst = reinterpret_cast<T*>(ch.dst[j]);
    const int64_t n = 16LL * O * I;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
      const int c = (int)(i % I);
      int64_t r = i / I;
      const int tap = (int)(r % 4); r /= 4;
      const int o = (int)(r % O), ph = (int)(r / O);
      const int ki = (ph >> 1) + 1 - 2 * (tap >> 1), kj = (ph & 1) + 1 - 2 * (tap & 1);
      const bool ok = ki >= 0 && ki <= 2 && kj >= 0 && kj <= 2;
      stf(dst + i, ok ? w[((int64_t)o * I + c) * 9 + ki * 3 + kj] : 0.f);
    }
    return;
  }
  if (mode == 2) {
    const float sc = ch.scale[j];
    float* dst = reinterpret_cast<float*>(ch.dst[j]);
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < (int64_t)O * I; i += (int64_t)gridDim.x * kThreads) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        T q;
        stf(&q, w[i * K + k]);
        const float f = ldf(&q) * sc;
        acc = fmaf(f, f, acc);
      }
      dst[i] = acc;
    }
    return;
  }
  T* dst = reinterpret_cast<T*>(ch.dst[j]);
  __shared__ float tile[kPT][kPT * kPackMaxK + 1];
  const int tiles_c = (I + kPT - 1) / kPT, tiles_o = (O + kPT - 1) / kPT;
  const int row = kPT * K;                    // floats of one source row inside the tile
  for (int t = blockIdx.x; t < tiles_o * tiles_c; t += gridDim.x) {
    const int o0 = (t / tiles_c) * kPT, c0 = (t % tiles_c) * kPT;
    const int no = min(kPT, O - o0), nc = min(kPT, I - c0);
    __syncthreads();
    for (int e = threadIdx.x; e < no * nc * K; e += kThreads) {
      const int r = e / (nc * K), q = e - r * (nc * K);
      tile[r][q] = w[((int64_t)(o0 + r) * I + c0) * K + q];
    }
    __syncthreads();
    if (mode == 0) {
      for (int e = threadIdx.x; e < no * K * nc; e += kThreads) {      // (r, k, cc), cc fastest
        const int cc = e % nc, rk = e / nc, k = rk % K, r = rk / K;
        stf(dst + (int64_t)(o0 + r) * K * I + (int64_t)k * I + c0 + cc, tile[r][cc * K + k]);
      }
    } else {
      for (int e = threadIdx.x; e < nc * K * no; e += kThreads) {      // (cc, k, r), r fastest
        const int r = e % no, ck = e / no, k = ck % K, cc = ck / K;
        stf(dst + (int64_t)(c0 + cc) * K * O + (int64_t)k * O + o0 + r, tile[r][cc * K + k]);
      }
    }
  }
  (void)row;
}

}  // namespace

extern "C" int lcgan_adam_step(const lcgan_adam_chunk* ch, float lr, float beta1, float beta2, float eps, void* stream) {
  LCGAN_CHECK(ch && ch->count > 0 && ch->count <= LCGAN_MT_MAX, "adam_step: bad chunk");
  LCGAN_CHECK(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
              "adam_step: bad hyper-parameters");
  for (int k = 0; k < ch->count; ++k)
    LCGAN_CHECK(ch->p[k] && ch->g[k] && ch->v[k] && ch->step[k] && ch->numel[k] >= 0 && (beta1 == 0.f || ch->m[k]),
                "adam_step: null pointer in chunk entry %d", k);
  cudaStream_t s = (cudaStream_t)stream;
  adam_kernel<<<dim3(kBlocksPerTensor, ch->count), kThreads, 0, s>>>(*ch, lr, beta1, beta2, eps);
  adam_count_kernel<<<1, LCGAN_MT_MAX, 0, s>>>(*ch);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_pack_weights(const lcgan_pack_chunk* ch, int dt, void* stream) {
  LCGAN_CHECK(ch && ch->count > 0 && ch->count <= LCGAN_MT_MAX, "pack_weights: bad chunk");
  for (int k = 0; k < ch->count; ++k)
    LCGAN_CHECK(ch->src[k] && ch->dst[k] && ch->O[k] > 0 && ch->I[k] > 0 && ch->K[k] > 0 && ch->K[k] <= kPackMaxK &&
                ch->mode[k] >= 0 && ch->mode[k] <= 3 && (ch->mode[k] != 3 || ch->K[k] == 9),
                "pack_weights: bad chunk entry %d", k);
  cudaStream_t s = (cudaStream_t)stream;
  // 37 KB of static shared memory per block.  The two 8192x2048 head weights dominate (16 384 tiles each), and a
  // tile is latency-bound (4 KB in, barrier, 2 KB out): 4 blocks per SM and tensor keep enough tiles in flight;
  // the blocks of small tensors exit at once
  const int bx = 592;
  if (dt == LCGAN_BF16) pack_kernel<bf16><<<dim3(bx, ch->count), kThreads, 0, s>>>(*ch);
  else if (dt == LCGAN_F32) pack_kernel<float><<<dim3(bx, ch->count), kThreads, 0, s>>>(*ch);
  else { lcgan_set_error("pack_weights: bad dtype code %d", dt); return 1; }
  LCGAN_LAUNCH_CHECK();
  return 0;
}
