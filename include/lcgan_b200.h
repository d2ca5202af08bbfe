/* lcgan_b200 C ABI  --  the drop-in boundary for the LC-GAN training hot path on B200 (sm_100a).
 *
 * The reference (rakutentech/lcgan) has no FFI of its own: its hot path is PyTorch library calls
 * made from custom_layers.py / cnn.py / loss.py.  Each entry point below replaces one of those
 * library call sites (cited per function as reference file:line).  The host side
 * (lcgan_b200/ops.py) wraps them in torch.autograd.Functions; INTEGRATION.md shows the ctypes
 * binding.  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises
 *     the device, never allocates or frees device memory (workspaces are caller-provided);
 *   - returns 0 on success, non-zero on error; lcgan_last_error() returns a message for the
 *     calling thread;
 *   - dtype codes: 0 = float32, 1 = bfloat16.
 */
#ifndef LCGAN_B200_H
#define LCGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCGAN_F32 0
#define LCGAN_BF16 1
#define LCGAN_MAX_TAPS 9

/* A "tap convolution": the one contraction every conv / transposed conv / linear layer and their
 * data- and weight-gradients reduce to.
 *
 *   for a lattice point (b, m, n), m < MH, n < MW:
 *     out pixel (oy, ox) = (m*os + py, n*os + px)
 *     acc[o] = sum_t sum_c  X[b, m*is + dy[t], n*is + dx[t], c] * W2[o][wtap[t]*Cin + c]
 *     Y[b, oy, ox, o] = lrelu(acc[o]*acc_scale*rowscale[b,o] + bias[o]*bias_scale + noise[oy,ox]*noise_scale, slope) * gain
 *                       + R[b,oy,ox,o]
 *   out-of-range input pixels read as zero.
 *
 * X and Y are addressed with explicit element strides, so NCHW fp32 images and channels-last bf16
 * activations go through the same descriptor.
 */
typedef struct lcgan_tapconv {
  int32_t N, IH, IW, Cin;          /* input tensor  */
  int32_t OH, OW, Cout;            /* output tensor */
  int64_t xs_n, xs_h, xs_w, xs_c;  /* input element strides  */
  int64_t ys_n, ys_h, ys_w, ys_c;  /* output (and residual) element strides */
  int32_t x_dtype, y_dtype, w_dtype;
  int32_t MH, MW;                  /* lattice extent */
  int32_t os, py, px;              /* lattice -> output pixel */
  int32_t is;                      /* lattice -> input pixel stride */
  int32_t ntaps;
  int32_t dy[LCGAN_MAX_TAPS], dx[LCGAN_MAX_TAPS], wtap[LCGAN_MAX_TAPS];
  int64_t w_ld;                    /* row stride (elements) of W2[Cout][w_ld] */
  float acc_scale;                 /* equalized-lr constant c (weights are packed unscaled) */
  float bias_scale, slope, gain;   /* slope = 1 -> no activation */
  const float* noise;              /* optional f32 plane [OH][OW] (device) added before the activation:
                                      + noise[oy][ox] * noise_scale - the noise injection of
                                      custom_layers.py:108-110; NULL = none */
  float noise_scale;
  const float* colscale;           /* optional f32 [N][Cin] (device): X[b,..,c] is multiplied by colscale[b,c] as it is
                                      read - the style modulation of custom_layers.py:62-64 in the prologue.  Only the
                                      pointwise thin kernels implement it (lcgan_tapconv_simt with Cin 32, 1x1,
                                      Cout <= 4); every other path rejects a non-NULL value */
} lcgan_tapconv;

const char* lcgan_last_error(void);
int lcgan_version(void);
/* 1 if the tcgen05 path can take this descriptor (channels-last bf16, Cin%64==0, ...) */
int lcgan_tapconv_tc_eligible(const lcgan_tapconv* d);

/* lcgan_tapconv_tc with BLOCKED output channels: channel o is stored at element offset
 * (o / cblk) * ys_blk + o % cblk of its lattice point's output pixel and takes rowscale[b, o % cperiod],
 * bias[o % cperiod].  Used to run conv_transpose2d(k3, s2, p1, op1) (custom_layers.py:78) as one launch
 * over the input lattice (4 taps, Cout' = 4 Cout, cblk = 2 Cout, ys_blk = one output row). */
int lcgan_tapconv_tc_blocked(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                             const float* rowscale, const float* bias, int cblk, int64_t ys_blk,
                             int cperiod, void* stream);

/* conv_transpose2d(k3, s2, p1, op1) with Cout <= 4 (the flow layers, custom_layers.py:78 with C -> 2):
 * all four output phases in one pass over X.  d = the descriptor of phase (0,0) of the x2 plan
 * (N, IH, IW, Cin, Cout, strides, dtypes, w_ld and the epilogue constants are read). */
int lcgan_tapconv_up2_thin_eligible(const lcgan_tapconv* d);
int lcgan_tapconv_up2_thin(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                           const float* rowscale, const float* bias, void* stream);

/* weight gradient of the same layer, all 9 taps in one pass over X (Cout <= 2, Cin % 4 == 0, Cin/4 a
 * divisor of 256): dw2 [Cout][9*Cin] f32 += scale * sum g x; d->y_* describe G [N, 2H, 2W, Cout]. */
int lcgan_tapconv_up2_thin_wgrad(const lcgan_tapconv* d, const void* x, const void* g, float* dw2,
                                 float scale, void* stream);

/* Gather for the flow layers' weight gradient on the tensor cores: g [N,2H,2W,2] f32 channels-last (the gradient of a
 * C -> 2 conv_transpose2d(k3,s2,p1,op1), custom_layers.py:78,150) -> out [N,H,W,32] bf16 channels-last with
 * out[b,m,n, (ki*3+kj)*2 + o] = g[b, 2m-1+ki, 2n-1+kj, o] (0 outside the image; channels 18..31 zero).  The weight
 * gradient is then the pointwise lcgan_tapconv_wgrad_tc of X against this tensor. */
int lcgan_flow_grad_im2col(const float* g, void* out, int N, int H, int W, void* stream);

/* Forward-type tap conv on CUDA cores (any strides/dtypes; fp32 accumulate).
 * Replaces F.conv2d / F.conv_transpose2d / F.linear call sites (custom_layers.py:25,41,43,78,83)
 * and their autograd data-gradients.  rowscale [N,Cout] f32, bias [Cout] f32, residual like Y;
 * each may be NULL. */
int lcgan_tapconv_simt(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                       const float* rowscale, const float* bias, const void* residual, void* stream);

/* Same contraction on the 5th-gen tensor cores: tcgen05.mma, TMEM accumulators, TMA-fed operands
 * (bf16 channels-last X, bf16 W2, Cin % 64 == 0).  Y may be bf16 or f32 channels-last. */
int lcgan_tapconv_tc(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                     const float* rowscale, const float* bias, const void* residual, void* stream);

/* Per-image weight gradient of a pointwise (1x1) layer with 32 input channels and Cout <= 4 (the to-RGB conv,
 * cnn.py:87): dwp[b][o][c] += sum_p G[b,p,o] * X[b,p,c], f32 [N][Cout][32], caller-zeroed.  With the unmodulated X
 * this one pass yields both gradients of the modulated layer: dW = sum_b s[b,c] dwp[b], ds[b,c] = sum_o w[o,c] dwp[b]. */
int lcgan_pw_wgrad32(const lcgan_tapconv* d, const void* x, const void* g, float* dwp, void* stream);

/* Weight gradient of the same contraction (autograd of custom_layers.py:41,43,78,83,25):
 *   dW2[o][wtap[t]*Cin + c] += scale * sum_{b,m,n} G[b, m*os+py, n*os+px, o] * X[b, m*is+dy[t], n*is+dx[t], c]
 * G is addressed with the descriptor's Y strides/dtype.  dW2 is f32 [Cout][w_ld], accumulated into
 * (caller zeroes it). */
int lcgan_tapconv_wgrad_simt(const lcgan_tapconv* d, const void* x, const void* g, float* dw2,
                             float scale, void* stream);
int lcgan_tapconv_wgrad_tc(const lcgan_tapconv* d, const void* x, const void* g, float* dw2,
                           float scale, void* stream);

/* ---- memory-bound kernels; tensors are dense channels-last [N,H,W,C] of dtype `dt` ---------- */

/* out = post(box3(pre(a)))  (F.avg_pool2d(3,1,1), custom_layers.py:137,197, fused with the
 * neighbouring leaky-relu*gain of :155,205 or with its backward mask).
 * pre(a) = a * (mask > 0 ? pre_gain : pre_gain*pre_slope) when mask != NULL, else a.
 * post(v) = (v > 0 ? v : v*post_slope) * post_gain. */
int lcgan_box3(const void* a, const void* mask, void* out, int dt, int N, int H, int W, int C,
               float pre_slope, float pre_gain, float post_slope, float post_gain, void* stream);

/* out = box3(a) * (y > 0 ? gain : gain*slope);  r0[b,c] += sum_p out (f32 [N,C], caller-zeroed, may be NULL):
 * the backward of F.avg_pool2d(3,1,1) (custom_layers.py:197) fused with the backward of the leaky-relu*gain of the conv
 * in front of it (:205), whose stored output is y.  Tile shapes only (C % 32 (bf16) / 16 (f32), W >= 32, H >= 16). */
int lcgan_box3_postmask(const void* a, const void* y, void* out, float* r0, int dt, int N, int H, int W, int C,
                        float slope, float gain, void* stream);

/* lcgan_box3 with the NEXT layer's style modulation (custom_layers.py:62-64) folded in; cs [N,C] f32:
 *   mask == NULL:  out = post(box3(a)) * cs[b,c]
 *   mask != NULL:  out = box3(a * cs[b,c] * (mask*cs > 0 ? pre_gain : pre_gain*pre_slope)),
 *                  red[b,c] += sum_p a*mask   (f32, caller-zeroed, may be NULL)
 * Shapes the tiled kernel takes only: C % 32 == 0 (bf16) / 16 (f32), W >= 32, H >= 16. */
int lcgan_box3_cs(const void* a, const void* mask, void* out, const float* cs, float* red, int dt, int N,
                  int H, int W, int C, float pre_slope, float pre_gain, float post_slope, float post_gain,
                  void* stream);

/* y[b,i,j,c] = scale * sum_{2x2} x[b,2i+a,2j+b,c]   (F.avg_pool2d(2,2), custom_layers.py:202; scale=.25) */
int lcgan_pool2(const void* x, void* y, int dt, int N, int H, int W, int C, float scale, void* stream);
/* y[b,2i+a,2j+b,c] = scale * x[b,i,j,c]   (F.interpolate nearest x2, custom_layers.py:146; adjoint of pool2) */
int lcgan_up2(const void* x, void* y, int dt, int N, int H, int W, int C, float scale, void* stream);
/* out = a + scale * nearest_up2(s): s [N,H,W,C], a and out [N,2H,2W,C].  The gradient of a tensor that feeds both a layer
 * (gradient a) and F.avg_pool2d(x, 2) (gradient s, scale .25) - DiscriminatorBlock's input, custom_layers.py:206-216 - in
 * one pass instead of up2 + add. */
int lcgan_up2_add(const void* a, const void* s, void* out, int dt, int N, int H, int W, int C, float scale, void* stream);

/* out [N,H/2,W/2,C] = sum-pool2(box3(g)), g [N,H,W,C]: the gradient of the block skip `box_filter(upsample(skip))`
 * (custom_layers.py:146-147,159) with respect to skip, in one pass. */
int lcgan_box3_pool2(const void* g, void* out, int dt, int N, int H, int W, int C, void* stream);

/* out = box3(nearest_up2(s)) + t   (custom_layers.py:146-147,159); s [N,H,W,C], t/out [N,2H,2W,C] */
int lcgan_up2box_add(const void* s, const void* t, void* out, int dt, int N, int H, int W, int C, void* stream);

/* Backward of the fused epilogue  y = lrelu(acc*d + bias)*gain:
 *   dz = dy * gain * (y > 0 ? 1 : slope);  gout = dz * d[b,c] (d may be NULL)
 *   r0[b,c] += sum_p dz;   r1[b,c] += sum_p dz * z,  z = y/gain (y>0) or y/(gain*slope)
 * r0/r1 (f32, caller-zeroed) may be NULL.  dy/y/gout are [N,P,C] channels-last. */
int lcgan_act_bwd(const void* dy, const void* y, void* gout, const float* d, float* r0, float* r1,
                  int dt, int N, int P, int C, float slope, float gain, void* stream);

/* xs[b,p,c] = x[b,p,c] * s[b,c]   (style modulation, custom_layers.py:62-64, shared-weight form) */
int lcgan_modulate(const void* x, const float* s, void* xs, int dt, int N, int P, int C, void* stream);
/* dx = t * s ; ds[b,c] += sum_p x*t */
int lcgan_modulate_bwd(const void* x, const void* t, const float* s, void* dx, float* ds,
                       int dt, int N, int P, int C, void* stream);

/* Flow warp (custom_layers.py:127-134,151,161-165): grid = linspace coords + tanh(flow)*scale,
 * bicubic (A=-0.75), zeros padding, align_corners=False.  x/out [N,H,W,C] dt; flow [N,H,W,2] f32
 * (pre-tanh). */
int lcgan_warp_fwd(const void* x, const float* flow, void* out, int dt, int N, int H, int W, int C,
                   float flow_scale, void* stream);
/* the same with out = warp(x) * cs[b,c] (cs [N,C] f32): the style modulation of the to-RGB conv that consumes the
 * warped features (custom_layers.py:62-64), folded into the warp pass.  Tiled shapes only (W >= 32, H >= 16,
 * C % 32 == 0 for bf16 / 16 for f32). */
int lcgan_warp_fwd_cs(const void* x, const float* flow, void* out, const float* cs, int dt, int N, int H, int W,
                      int C, float flow_scale, void* stream);
/* dx_acc [N,H,W,C] f32 (caller-zeroed, atomically accumulated); dflow [N,H,W,2] f32 (written). */
int lcgan_warp_bwd(const void* x, const float* flow, const void* dout, float* dx_acc, float* dflow,
                   int dt, int N, int H, int W, int C, float flow_scale, void* stream);

/* Backward of the flow warp with dx written once in the activation dtype: dx [N,H,W,C] dt and
 * dflow [N,H,W,2] f32.  Smooth flows (the output pixels that reach a 32x16 source tile fit a few
 * 48x32 windows - decided per tile on the device from the field itself) run as shared-memory tiled
 * gathers without atomics; otherwise the scatter kernels of lcgan_warp_bwd accumulate into ws_acc
 * and the result is cast.  ws_acc: f32 scratch of N*H*W*C elements (bf16 only, fp32 accumulates in
 * dx); ws_bounds: 4 * N * ceil(H/16) * ceil(W/32) + 4 ints of scratch.  No host synchronisation:
 * CUDA-graph capturable. */
int lcgan_warp_bwd_tiled(const void* x, const float* flow, const void* dout, void* dx, float* dflow,
                         float* ws_acc, int* ws_bounds, int dt, int N, int H, int W, int C,
                         float flow_scale, void* stream);

/* dtype/layout conversion out = (T_out) in, both dense with the same element order */
int lcgan_cast(const void* in, void* out, int dt_in, int dt_out, int64_t n, void* stream);

/* ---- loss reductions ---------------------------------------------------------------------- */
/* y = x / max(||x||_2, 1e-12) per row (F.normalize, cnn.py:40-41); x,y [B,D] f32; inv_norm [B] out */
int lcgan_l2norm_fwd(const float* x, float* y, float* inv_norm, int B, int D, void* stream);
int lcgan_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int B, int D, void* stream);
/* per-sample InfoNCE with one negative (loss.py:9-15): l[b] = softplus((a.n - a.p)/tau);
 * sig[b] = sigmoid((a.n - a.p)/tau) kept for backward. */
int lcgan_contrastive_fwd(const float* a, const float* p, const float* n, float* l, float* sig,
                          int B, int D, float tau, void* stream);
int lcgan_contrastive_bwd(const float* a, const float* p, const float* n, const float* sig,
                          const float* dl, float* da, float* dp, float* dn, int B, int D, float tau, void* stream);
/* out[b] = sum_i x[b,i]^2   (R1, loss.py:21-23); x [B,L] f32 */
int lcgan_sumsq(const float* x, float* out, int B, int64_t L, void* stream);
/* y[b,i] = x[b,i] * s[b]  (backward of sumsq) */
int lcgan_rowscale(const float* x, const float* s, float* y, int B, int64_t L, void* stream);

/* multi-tensor EMA (ema.py:26-32): dst = src + decay*(dst - src) over n contiguous f32 spans.
 * decay_dev (device scalar, may be NULL) overrides `decay`: a captured CUDA graph can then follow the
 * ema.py:19-23 start_iter schedule without re-capture. */
int lcgan_ema_lerp(float* const* dst, const float* const* src, const int64_t* numel, int n,
                   float decay, const float* decay_dev, void* stream); /* dst/src/numel: DEVICE arrays of length n */

/* ---- multi-tensor optimizer step and weight packing (optim.cu) ------------------------------- */
#define LCGAN_MT_MAX 48
/* Up to LCGAN_MT_MAX f32 tensors; the struct itself is a HOST object passed by value to the kernel. */
typedef struct lcgan_adam_chunk {
  float* p[LCGAN_MT_MAX];          /* parameters (updated in place) */
  const float* g[LCGAN_MT_MAX];    /* gradients */
  float* m[LCGAN_MT_MAX];          /* exp_avg (may be NULL when beta1 == 0: m = g) */
  float* v[LCGAN_MT_MAX];          /* exp_avg_sq */
  float* step[LCGAN_MT_MAX];       /* per-tensor completed-step counters (device f32 scalars; incremented) */
  int64_t numel[LCGAN_MT_MAX];
  int32_t count;
} lcgan_adam_chunk;
/* torch.optim.Adam(betas, eps, weight_decay=0, amsgrad=False) step (worker.py:98-110) for every tensor
 * of the chunk in one launch; bias corrections use each tensor's own step counter, like torch. */
int lcgan_adam_step(const lcgan_adam_chunk* chunk, float lr, float beta1, float beta2, float eps, void* stream);

typedef struct lcgan_pack_chunk {
  const float* src[LCGAN_MT_MAX];  /* w [O][I][K] f32 (K = kh*kw) */
  void* dst[LCGAN_MT_MAX];
  int32_t O[LCGAN_MT_MAX], I[LCGAN_MT_MAX], K[LCGAN_MT_MAX];
  int32_t mode[LCGAN_MT_MAX];      /* 0: [O][K*I] (forward), 1: [I][K*O] (data gradient), 2: Wsq [O][I] f32,
                                      3: [4*O][4*I] fused x2 transposed-conv weight (lcgan_tapconv_tc_blocked) */
  float scale[LCGAN_MT_MAX];       /* mode 2: sum_k (round(w)*scale)^2 */
  int32_t count;
} lcgan_pack_chunk;
/* Tap-conv weight packs (W2 layouts of lcgan_tapconv) and demodulation tables (custom_layers.py:65-67)
 * for up to LCGAN_MT_MAX weights in one launch; dt = dtype code of the packs. */
int lcgan_pack_weights(const lcgan_pack_chunk* chunk, int dt, void* stream);

/* ---- weight-/[b,C]-sized pieces of the modulated convolution (demod.cu; custom_layers.py:62-68) --- */
/* d[b,o] = rsqrt(sum_c s[b,c]^2 wsq[o,c] + eps); s [B,I], wsq [O,I] (lcgan_pack_weights mode 2), d [B,O], all f32 */
int lcgan_demod_fwd(const float* s, const float* wsq, float* d, int B, int O, int I, float eps, void* stream);
/* backward of the above through q = sum_c s^2 wsq: with dq = -0.5 dd d^3,
 *   ds[b,c] = 2 s[b,c] sum_o dq[b,o] wsq[o,c]                         (ds may be NULL)
 *   dw[o,c,k] = 2 wscale^2 round_dt(w[o,c,k]) sum_b dq[b,o] s[b,c]^2    (dw [O,I,K] f32, may be NULL) */
int lcgan_demod_bwd(const float* dd, const float* d, const float* s, const float* wsq, const float* w,
                    float* ds, float* dw, int B, int O, int I, int K, float wscale, int dt, void* stream);
/* Parameter-side gradients of the fused epilogue from the per-(b,o) sums of lcgan_act_bwd:
 *   db[o] = bias_scale sum_b r0[b,o];   dd[b,o] = (r1[b,o] - bias[o] bias_scale r0[b,o]) / d[b,o]
 * db or dd may be NULL (dd needs r1 and d; bias may be NULL). */
int lcgan_epilogue_grads(const float* r0, const float* r1, const float* bias, const float* d, float bias_scale,
                         float* db, float* dd, int B, int O, void* stream);

/* Deterministic mode: reductions that finish with fp32 atomics (split-K weight gradients, per-(b,c)
 * sums) take ordered turns instead, so repeated runs are bit-identical.  Returns the previous setting.
 * Kernels launched in this mode must not run concurrently on two streams. */
int lcgan_set_deterministic(int on);

/* Debug: raw tcgen05 / TMA issue rates (scratch/rates.py).  mode 0 / 1: `iters` back-to-back M=128 x n x K=16 MMAs with
 * K-major / MN-major shared-memory operands; mode 2: `iters` TMA loads of a 16 x ht box of act [N,H,W,C] into a 4-slot
 * ring.  out_cycles[block] = elapsed SM clocks.  Not used by the product path. */
int lcgan_debug_tc_rate(int mode, int n, int iters, const void* act, int N, int H, int W, int C, int ht,
                        long long* out_cycles, int blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif
