// Tap-convolution on the 5th-generation tensor cores (sm_100a):
//   tcgen05.mma (cta_group::1, kind::f16, bf16 x bf16 -> fp32) issued by one elected thread,
//   accumulators in TMEM, operands staged in shared memory by TMA (cp.async.bulk.tensor) through an
//   mbarrier ring, epilogue (demod row scale, bias, leaky-relu, gain, residual) applied on the
//   tcgen05.ld'ed accumulator and stored straight to the channels-last output.
//
// Implicit GEMM without im2col: the activation operand of a tap is ONE 4-D TMA box
//   {64 channels, wt, ht, nt} of the channels-last tensor at the tap's shifted origin - out-of-range
//   coordinates are zero-filled by TMA, which is exactly the convolution's zero padding, and
//   elementStrides = 2 walks the stride-2 lattices.  The 128 rows of the box land in shared memory
//   as a K-major SWIZZLE_128B tile, directly consumable by the UMMA descriptor.
//
// Forward-type kernel:  D[128 lattice points x BN out-channels] += A[128 x 64] * W2[BN x 64]^T
//                       over taps x (Cin/64) k-blocks.                  (A, B K-major)
// Weight-gradient kernel: D[128 out-ch x BN in-ch] += G[128 px x 128 ch]^T * X[128 px x BN ch]
//                       over the lattice points of a K-split.           (A, B MN-major, same TMA tiles)
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace {

constexpr int kStages = 3;                     // weight-gradient kernel: 3 x 64 KiB
constexpr int kFwdStages = 4;                  // forward kernel: 4 x 32 KiB
constexpr int kAccStages = 4;                  // TMEM accumulator ring (4 x 128 columns = all 512)
constexpr int kTileM = 128;                    // lattice points (fwd) / pixels per k-block (wgrad)
constexpr int kBlockK = 64;                    // max channels per k-block (one 128-byte swizzle row); 32 -> 64-byte rows
constexpr int kMaxBN = 128;
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KiB
constexpr int kBBytes = kMaxBN * kBlockK * 2;  // 16 KiB
constexpr int kThreads = 192;                  // warp0: TMA, warp1: MMA, warps 2-5: epilogue
constexpr int kMaxSmem = 232448;                // opt-in dynamic shared memory per block on sm_100 (227 KiB)
constexpr int kFwdThreads = 608;               // warps: 0 TMA, 1+6 MMA issuers, 2-5 / 7-10 / 11-14 / 15-18 epilogue groups

struct TcParams {
  int N, Cin, Cout;
  int wt, ht, nt, tiles_w, tiles_h;
  int is, os, py, px;
  int ntaps, kpt, kc;                           // kc = channels per k-block (64 or 32), kpt = Cin / kc
  int dy[LCGAN_MAX_TAPS], dx[LCGAN_MAX_TAPS], wtap[LCGAN_MAX_TAPS];
  int BN, n_tiles, total_tiles;
  int lw, lh;                                   // log2(tiles_w), log2(tiles_h): tile decode by shifts
  int rowshare;                                 // 3x3 stride-1: 1/2 = one tall A tile per dx serves the 3 dy taps;
                                                // 3 = ONE haloed (wt+2) x (ht+2) tile serves all nine taps
  int grp_wtap[3][3];                           // [dx+1][dy+1] -> weight tap index
  int stages, a_bytes, b_bytes, wres_bytes;     // rowshare == 2: all 9 taps' weights stay resident in smem
  long long ys_n, ys_h, ys_w;
  int y_f32;
  float acc_scale, bias_scale, slope, gain;
  // blocked output channels (lcgan_tapconv_tc_blocked): channel o lands at (o / cblk) * ys_blk + o % cblk
  // and takes rowscale / bias of channel o % cperiod; cblk == 0: plain channel-innermost output
  int cblk, cperiod;
  long long ys_blk;
  const float* noise;                           // [OH][OW] f32 plane added before the activation (or null)
  float noise_scale;
  int OW;
  int tiles_w_px;                               // rowshare == 5: output width (columns past it are not stored)
  int dbg;                                    // LCGAN_TC_DEBUG timing experiments (1: no MMAs, 2: no TMA loads, 4: no stores)
};

struct WgParams {
  int N, Cin, Cout;
  int wt, ht, nt, tiles_w, tiles_h, tiles_total;
  int is, os, py, px;
  int ntaps;
  int dy[LCGAN_MAX_TAPS], dx[LCGAN_MAX_TAPS], wtap[LCGAN_MAX_TAPS];
  int BN, ctiles, kcg, kcx;                     // channels per TMA box of G and of X (64 or 32)
  int tiles_per_split;
  long long w_ld;
  float scale;
  int* sems;                                    // deterministic mode: ordered K-split accumulation (common.cuh)
};

__device__ int g_sems[kDetSems];

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Single-thread instructions take a `leader` flag and are predicated INSIDE the asm: the surrounding
// loop is executed by all 32 lanes with warp-uniform values, so descriptor/coordinate arithmetic
// stays in the uniform datapath instead of per-operand R2UR moves under a divergent `if (lane == 0)`
// (measured: ~150 cycles of issue overhead per MMA with the divergent form vs a 45-64 cycle MMA).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes, uint32_t leader) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n"
               ::"r"(smem_u32(bar)), "r"(bytes), "r"(leader) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            uint32_t leader) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %7, 0;\n"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(leader) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %5, 0;\n"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(leader) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                         uint32_t leader) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t leader) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n"
               ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1, SWIZZLE_128B).
// kc = channels per swizzle row: 64 -> SWIZZLE_128B (layout 2), 32 -> SWIZZLE_64B (layout 4).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int kc) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)(kc == 64 ? 2 : 4) << 61;     // LayoutType::SWIZZLE_128B / SWIZZLE_64B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M=128.
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                          // c_format = F32
  d |= 1u << 7;                          // a_format = BF16
  d |= 1u << 10;                         // b_format = BF16
  d |= (a_mn_major ? 1u : 0u) << 15;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(kTileM >> 4) << 24;
  return d;
}

struct Smem {
  uint8_t* base;          // 1024-byte aligned; stage i: A at base + i*stage_bytes, B right after A
  uint32_t stage_bytes, a_bytes;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* done;         // [kAccStages] accumulator-full barriers (wgrad uses done[0] only)
  uint64_t* acc_empty;    // [kAccStages]
  uint64_t* wfull;        // resident-weight region filled
  uint8_t* wres;          // resident weights (small-channel mode), 1024-byte aligned
  uint32_t* tmem_slot;
  __device__ __forceinline__ uint8_t* a(int st) const { return base + (size_t)st * stage_bytes; }
  __device__ __forceinline__ uint8_t* b(int st) const { return base + (size_t)st * stage_bytes + a_bytes; }
};

__device__ __forceinline__ Smem carve(uint8_t* raw, int stages, int a_bytes, int b_bytes, int wres_bytes = 0) {
  Smem s;
  s.wres = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  s.base = s.wres + wres_bytes;
  s.stage_bytes = a_bytes + b_bytes;
  s.a_bytes = a_bytes;
  uint8_t* tail = s.base + (size_t)stages * s.stage_bytes;
  s.full = (uint64_t*)tail;
  s.empty = s.full + stages;
  s.done = s.empty + stages;
  s.acc_empty = s.done + kAccStages;
  s.wfull = s.acc_empty + kAccStages;
  s.tmem_slot = (uint32_t*)(s.wfull + 1);
  return s;
}

// ---------------------------------------------------------------------------------------------
// forward-type kernel
// ---------------------------------------------------------------------------------------------
// LEAN: the epilogue of the dominant HBM-bound shapes (one 32-channel tile, bf16 output, unit output stride, no
// residual / noise / blocked layout, resident-weight modes) with everything per-tile hoisted: ncu had the generic
// epilogue at ~300 instructions per thread and tile for 16 outputs (8 warps x 300 = 2 400 of the 2 575 warp
// instructions a tile costs; 644 issue cycles per scheduler against the 690-cycle HBM budget of a tile).
template <bool LEAN>
__global__ void __launch_bounds__(kFwdThreads, 1)
tapconv_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw, const TcParams p,
                  void* __restrict__ y, const float* __restrict__ rowscale, const float* __restrict__ bias,
                  const void* __restrict__ residual) {
  // Persistent: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...  Three decoupled roles:
  //   warp 0   TMA producer, runs ahead through the 4-stage smem ring across tile boundaries
  //   warp 1   MMA issuer, accumulates tile i into TMEM stage i%2
  //   warps2-5 epilogue of tile i overlaps the MMAs of tile i+1 (TMEM double buffer)
  extern __shared__ uint8_t smem_raw[];
  const Smem s = carve(smem_raw, p.stages, p.a_bytes, p.b_bytes, p.wres_bytes);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmw);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(s.wfull, 1);
    for (int i = 0; i < kAccStages; ++i) { mbar_init(&s.done[i], p.rowshare >= 2 ? 1 : 2); mbar_init(&s.acc_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAccStages * kMaxBN>(s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;
  const int nkb = (p.rowshare ? 3 : p.ntaps) * p.kpt;
  const uint32_t wtile = (uint32_t)p.BN * p.kc * 2;           // bytes of one tap's weight tile

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();
      const uint32_t tx_bytes = p.rowshare ? (uint32_t)p.a_bytes + 3 * wtile : (kTileM + p.BN) * p.kc * 2;
      int g = 0;                                            // k-block counter across tiles
      if (p.rowshare == 4) {
        // x2 transposed conv, haloed: ONE (wt+1) x (ht+1) box of the input lattice per tile; the four taps
        // (dy, dx in {0,1}) are descriptor offsets into it, the nine live (tap, phase) weight blocks stay resident
        // Resident weights: per tap the blocks of the phases it reaches, in ACCUMULATOR-COLUMN order.  The accumulator
        // holds the phases in the order 0, 1, 3, 2 so that each tap's phases are contiguous columns: tap (0,0) -> all
        // four (one N = 4*Cout MMA), tap (0,1) -> phases 1, 3, tap (1,0) -> phases 3, 2, tap (1,1) -> phase 3.
        const uint32_t wblk = (uint32_t)p.cperiod * p.kc * 2;
        mbar_expect_tx(s.wfull, 9 * wblk, leader);
        {
          const int order[9][2] = {{0, 0}, {0, 1}, {0, 3}, {0, 2}, {1, 1}, {1, 3}, {2, 3}, {2, 2}, {3, 3}};   // (tap, phase)
          for (int idx = 0; idx < 9; ++idx)
            tma_load_2d(s.wres + idx * wblk, &tmw, s.wfull, order[idx][0] * p.Cin, order[idx][1] * p.cperiod, leader);
        }
        const uint32_t box_bytes = (uint32_t)(p.wt + 1) * (p.ht + 1) * p.kc * 2;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++g) {
          const int tw = tile & (p.tiles_w - 1), th = (tile >> p.lw) & (p.tiles_h - 1), tb = tile >> (p.lw + p.lh);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          mbar_expect_tx(&s.full[st], box_bytes, leader);
          tma_load_4d(s.a(st), &tmx, &s.full[st], 0, tw * p.wt, th * p.ht, tb, leader);
        }
      } else if (p.rowshare == 5) {
        // dx-on-N mode (see the MMA issuer): a 16-pixel-wide, (8+2)-row box per tile, 14 output columns of it live;
        // resident weights ordered [dy][dx] so that the three dx taps of one dy are ONE B operand of 3*Cout rows
        mbar_expect_tx(s.wfull, 9 * wtile, leader);
        for (int dyi = 0; dyi < 3; ++dyi)
          for (int j = 0; j < 3; ++j)
            tma_load_2d(s.wres + (dyi * 3 + j) * wtile, &tmw, s.wfull, p.grp_wtap[j][dyi] * p.Cin, 0, leader);
        const uint32_t box_bytes = 16u * 10u * p.kc * 2;
        const int per_img = p.tiles_w * p.tiles_h;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++g) {
          const int tb = tile / per_img, tr = tile - tb * per_img;
          const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          if (p.dbg & 2) { if (leader) mbar_arrive(&s.full[st]); continue; }
          mbar_expect_tx(&s.full[st], box_bytes, leader);
          tma_load_4d(s.a(st), &tmx, &s.full[st], 0, tw * 14 - 1, th * 8 - 1, tb, leader);
        }
      } else if (p.rowshare == 6) {
        // haloed STRIDE-2 mode (3x3, 32 input channels): the input is seen as [N, H, W/2, 2*32] - a PAIR of pixels is one
        // 128-byte swizzle row - and ONE unit-stride box of (8+1) pairs x (2*16+1) rows per 8 x 16 lattice tile replaces
        // nine element-stride-2 boxes (which fetch 64-byte pieces at 128-byte stride: 0.42 of HBM).  Tap (dy, dx) of
        // lattice point (m, n) is input pixel (2m+dy-1, 2n+dx-1) = box row 2m+dy, pair n + ((dx+1)>>1), half (dx != 1):
        // a descriptor start offset, with consecutive 8-pair groups two box rows apart (SBO).
        mbar_expect_tx(s.wfull, 9 * wtile, leader);
        for (int j = 0; j < 3; ++j)
          for (int dyi = 0; dyi < 3; ++dyi)
            tma_load_2d(s.wres + (j * 3 + dyi) * wtile, &tmw, s.wfull, p.grp_wtap[j][dyi] * p.Cin, 0, leader);
        const uint32_t box_bytes = 9u * 33u * 128u;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++g) {
          const int tw = tile & (p.tiles_w - 1), th = (tile >> p.lw) & (p.tiles_h - 1), tb = tile >> (p.lw + p.lh);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          mbar_expect_tx(&s.full[st], box_bytes, leader);
          tma_load_4d(s.a(st), &tmx, &s.full[st], 0, tw * 8 - 1, th * 32 - 1, tb, leader);
        }
      } else if (p.rowshare == 3) {
        // haloed small-channel mode: resident weights, and ONE TMA box per tile - the (wt+2) x (ht+2) pixel
        // neighbourhood of the 8 x 16 lattice tile.  Tap (dy, dx) is read by the MMA straight out of that
        // box through its descriptor start offset (rows (dy*(wt+2) + dx) further on, SBO = one stored row of
        // wt+2 pixels): the swizzle XOR is a function of the absolute shared-memory address for both the
        // TMA write and the MMA read (probed on the B200: profiles/r02_probe_halo_single_copy_tile.txt).
        mbar_expect_tx(s.wfull, 9 * p.kpt * wtile, leader);
        for (int j = 0; j < 3; ++j)
          for (int dyi = 0; dyi < 3; ++dyi)
            for (int cb = 0; cb < p.kpt; ++cb)
              tma_load_2d(s.wres + ((j * 3 + dyi) * p.kpt + cb) * wtile, &tmw, s.wfull,
                          p.grp_wtap[j][dyi] * p.Cin + cb * p.kc, 0, leader);
        const uint32_t box_bytes = (uint32_t)(p.wt + 2) * (p.ht + 2) * p.kc * 2;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++g) {
          const int tw = tile & (p.tiles_w - 1), th = (tile >> p.lw) & (p.tiles_h - 1), tb = tile >> (p.lw + p.lh);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          if (p.dbg & 2) { if (leader) mbar_arrive(&s.full[st]); continue; }
          mbar_expect_tx(&s.full[st], box_bytes, leader);
          tma_load_4d(s.a(st), &tmx, &s.full[st], 0, tw * p.wt - 1, th * p.ht - 1, tb, leader);
        }
      } else if (p.rowshare == 2) {
        // small-channel mode: the 9 taps' weights are loaded once and stay in smem; one stage = the
        // three tall (dx = -1,0,+1) A tiles of a whole output tile -> 3*kpt TMA issues per tile
        const uint32_t a_tile = (uint32_t)(p.ht + 2) * 16 * p.kc * 2;
        mbar_expect_tx(s.wfull, 9 * p.kpt * wtile, leader);
        for (int j = 0; j < 3; ++j)
          for (int dyi = 0; dyi < 3; ++dyi)
            for (int cb = 0; cb < p.kpt; ++cb)
              tma_load_2d(s.wres + ((j * 3 + dyi) * p.kpt + cb) * wtile, &tmw, s.wfull,
                          p.grp_wtap[j][dyi] * p.Cin + cb * p.kc, 0, leader);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++g) {
          const int tw = tile & (p.tiles_w - 1), th = (tile >> p.lw) & (p.tiles_h - 1), tb = tile >> (p.lw + p.lh);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          mbar_expect_tx(&s.full[st], 3 * p.kpt * a_tile, leader);
          for (int j = 0; j < 3; ++j)
            for (int cb = 0; cb < p.kpt; ++cb)
              tma_load_4d(s.a(st) + (j * p.kpt + cb) * a_tile, &tmx, &s.full[st], cb * p.kc, tw * p.wt + j - 1,
                          th * p.ht - 1, tb * p.nt, leader);
        }
      } else
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt_i = tile % p.n_tiles;
        const int t = tile / p.n_tiles;
        const int tw = t & (p.tiles_w - 1), th = (t >> p.lw) & (p.tiles_h - 1), tb = t >> (p.lw + p.lh);
        const int n0 = tw * p.wt, m0 = th * p.ht, b0 = tb * p.nt, o0 = nt_i * p.BN;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.empty[st], ph ^ 1);
          const int tap = kb / p.kpt, cb = kb - tap * p.kpt;
          mbar_expect_tx(&s.full[st], tx_bytes, leader);
          if (p.rowshare) {
            // `tap` is the dx group: a (ht+2)-row tile starting one row above serves dy = -1, 0, +1
            tma_load_4d(s.a(st), &tmx, &s.full[st], cb * p.kc, n0 + tap - 1, m0 - 1, b0, leader);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi)
              tma_load_2d(s.b(st) + dyi * wtile, &tmw, &s.full[st], p.grp_wtap[tap][dyi] * p.Cin + cb * p.kc, o0, leader);
          } else {
            tma_load_4d(s.a(st), &tmx, &s.full[st], cb * p.kc, n0 * p.is + p.dx[tap], m0 * p.is + p.dy[tap], b0, leader);
            tma_load_2d(s.b(st), &tmw, &s.full[st], p.wtap[tap] * p.Cin + cb * p.kc, o0, leader);
          }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    {
      // Resident mode: two issuing warps alternate tiles - one thread needs ~120 cycles of descriptor
      // moves per MMA while an N<=64 MMA executes in 45.  Safe only because there every tile owns
      // exactly one smem stage and the stage / accumulator counts are even, so each mbarrier is waited
      // on by ONE warp, in order (a warp that skipped uses of a barrier would alias its phase parity;
      // that is why the k-block-ring mode keeps a single issuer).
      const int mine = warp == 1 ? 0 : 1;
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc(p.BN, false, false);
      int g = 0, li = 0;
      if (p.rowshare == 4) {
        mbar_wait(s.wfull, 0);
        const uint32_t row16 = (uint32_t)(p.kc * 2) >> 4, pitch = (uint32_t)(p.wt + 1) * row16;
        const uint32_t w_tl = ((uint32_t)p.cperiod * p.kc * 2) >> 4;
        const uint32_t idesc4 = make_idesc(p.cperiod, false, false);          // N = Cout per (tap, phase) block
        const uint32_t idesc_n2 = make_idesc(2 * p.cperiod, false, false), idesc_n4 = make_idesc(4 * p.cperiod, false, false);
        const uint64_t bd0 = make_desc(smem_u32(s.wres), 16, 16 * p.kc, p.kc);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li, ++g) {
          if ((li & 1) != mine) continue;
          const int as = li % kAccStages, aph = (li / kAccStages) & 1;
          mbar_wait(&s.acc_empty[as], aph ^ 1);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(as * kMaxBN);
          const uint64_t ad0 = make_desc(smem_u32(s.a(st)), 16, (uint32_t)(p.wt + 1) * p.kc * 2, p.kc);
          // one MMA per (tap, K step) over all the phases the tap reaches (N = 4, 2, 2, 1 x Cout): 4 instead of 9
          // MMAs per K step - an N <= 128 MMA costs ~60 cycles of operand streaming whatever N is
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint64_t ad = ad0 + (uint32_t)((t >> 1) * pitch + (t & 1) * row16);
            const int first_blk = t == 0 ? 0 : (t == 1 ? 4 : (t == 2 ? 6 : 8));
            const int first_col = t == 0 ? 0 : (t == 1 ? 1 : 2);          // accumulator position of the tap's first phase
            const uint32_t idn = t == 0 ? idesc_n4 : (t == 3 ? idesc4 : idesc_n2);
            const uint32_t tcol = tacc + (uint32_t)(first_col * p.cperiod);
            const uint64_t bd = bd0 + (uint32_t)(first_blk * w_tl);
            // tap (0,0) reaches every phase first: it initialises the four accumulators
            umma_f16(tcol, ad, bd, idn, t != 0, leader);
            umma_f16(tcol, ad + 2, bd + 2, idn, 1, leader);
            if (p.kc == 64) {
              umma_f16(tcol, ad + 4, bd + 4, idn, 1, leader);
              umma_f16(tcol, ad + 6, bd + 6, idn, 1, leader);
            }
          }
          umma_commit(&s.empty[st], leader);
          umma_commit(&s.done[as], leader);
        }
      } else if (p.rowshare == 5) {
        // dx-on-N: D'[(m, n'), (dx, o)] = sum_dy sum_c X[m + dy - 1, n', c] W[o, dy, dx, c] - one MMA per (dy, K step)
        // with N = 3 * Cout; the epilogue adds the three dx column groups of neighbouring pixels.  Timing the haloed
        // mode with its memory traffic switched off (LCGAN_TC_DEBUG) showed its 18 N = 32 MMAs per tile alone take the
        // kernel's whole duration: ~64 cycles each, the cost of streaming the 4 KB A operand out of shared memory,
        // not of the 16-cycle tensor work.  Here an A operand is read once per dy instead of once per tap.
        mbar_wait(s.wfull, 0);
        const uint32_t idesc5 = make_idesc(3 * p.BN, false, false);
        const uint32_t a_dy = (uint32_t)(16 * p.kc * 2) >> 4;             // one stored row of 16 pixels, 16-byte units
        const uint32_t w_dy = (3 * wtile) >> 4;
        const uint64_t bd0 = make_desc(smem_u32(s.wres), 16, 16 * p.kc, p.kc);
        const int nacc = 3 * p.BN <= kMaxBN ? kAccStages : 2;
        const uint32_t acc_stride = (uint32_t)(kAccStages * kMaxBN / nacc);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li, ++g) {
          if ((li & 1) != mine) continue;
          const int as = li % nacc, aph = (li / nacc) & 1;
          mbar_wait(&s.acc_empty[as], aph ^ 1);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)as * acc_stride;
          const uint64_t ad0 = make_desc(smem_u32(s.a(st)), 16, 16 * p.kc, p.kc);
          if (!(p.dbg & 1))
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi) {
            const uint64_t ad = ad0 + (uint32_t)(dyi * a_dy);
            const uint64_t bd = bd0 + (uint32_t)(dyi * w_dy);
            umma_f16(tacc, ad, bd, idesc5, dyi != 0, leader);
            umma_f16(tacc, ad + 2, bd + 2, idesc5, 1, leader);
            if (p.kc == 64) {
              umma_f16(tacc, ad + 4, bd + 4, idesc5, 1, leader);
              umma_f16(tacc, ad + 6, bd + 6, idesc5, 1, leader);
            }
          }
          umma_commit(&s.empty[st], leader);
          umma_commit(&s.done[as], leader);
        }
      } else if (p.rowshare == 6) {
        mbar_wait(s.wfull, 0);
        const uint32_t pitch16 = 9 * 8;                                   // one box row of 9 pixel pairs, 16-byte units
        const uint32_t w_tl = wtile >> 4;
        const uint64_t bd0 = make_desc(smem_u32(s.wres), 16, 16 * 32, 32);      // weights: 64-byte rows (K = 32)
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li, ++g) {
          if ((li & 1) != mine) continue;
          const int as = li % kAccStages, aph = (li / kAccStages) & 1;
          mbar_wait(&s.acc_empty[as], aph ^ 1);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(as * kMaxBN);
          // A: 128-byte rows (pixel pairs, SWIZZLE_128B); the wanted pixel is one half of the row = two K steps of it
          const uint64_t ad0 = make_desc(smem_u32(s.a(st)), 16, 2 * 9 * 128, 64);
          uint32_t first = 0;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const uint64_t ad = ad0 + (uint32_t)(dyi * pitch16 + ((j + 1) >> 1) * 8 + (j != 1 ? 4 : 0));
              const uint64_t bd = bd0 + (uint32_t)((j * 3 + dyi) * w_tl);
              umma_f16(tacc, ad, bd, idesc, first, leader);
              umma_f16(tacc, ad + 2, bd + 2, idesc, 1, leader);
              first = 1;
            }
          }
          umma_commit(&s.empty[st], leader);
          umma_commit(&s.done[as], leader);
        }
      } else if (p.rowshare == 3) {
        mbar_wait(s.wfull, 0);
        const uint32_t row16 = (uint32_t)(p.kc * 2) >> 4;               // one stored pixel, in 16-byte units
        const uint32_t pitch = (uint32_t)(p.wt + 2) * row16;              // one stored row of wt+2 pixels
        const uint32_t w_tl = wtile >> 4;
        const uint64_t bd0 = make_desc(smem_u32(s.wres), 16, 16 * p.kc, p.kc);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li, ++g) {
          if ((li & 1) != mine) continue;
          const int as = li % kAccStages, aph = (li / kAccStages) & 1;
          mbar_wait(&s.acc_empty[as], aph ^ 1);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(as * kMaxBN);
          // 8-row core groups = the 8 pixels of one lattice row; consecutive groups one stored row apart
          const uint64_t ad0 = make_desc(smem_u32(s.a(st)), 16, (uint32_t)(p.wt + 2) * p.kc * 2, p.kc);
          uint32_t first = 0;
          if (!(p.dbg & 1))
#pragma unroll
          for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const uint64_t ad = ad0 + (uint32_t)(dyi * pitch + j * row16);
              const uint64_t bd = bd0 + (uint32_t)((j * 3 + dyi) * w_tl);
              umma_f16(tacc, ad, bd, idesc, first, leader);
              umma_f16(tacc, ad + 2, bd + 2, idesc, 1, leader);
              if (p.kc == 64) {
                umma_f16(tacc, ad + 4, bd + 4, idesc, 1, leader);
                umma_f16(tacc, ad + 6, bd + 6, idesc, 1, leader);
              }
              first = 1;
            }
          }
          umma_commit(&s.empty[st], leader);
          umma_commit(&s.done[as], leader);
        }
      } else if (p.rowshare == 2) {
        const uint32_t a_tile = (uint32_t)(p.ht + 2) * 16 * p.kc * 2;
        mbar_wait(s.wfull, 0);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li, ++g) {
          if ((li & 1) != mine) continue;
          const int as = li % kAccStages, aph = (li / kAccStages) & 1;
          mbar_wait(&s.acc_empty[as], aph ^ 1);
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(as * kMaxBN);
          // descriptors differ only in the 14-bit start-address field (16-byte units): one make_desc
          // per tile, then integer offsets - the issue loop must stay far below the 16..64 cycles an
          // N = 32..128 MMA takes
          const uint64_t ad0 = make_desc(smem_u32(s.a(st)), 16, 16 * p.kc, p.kc);
          const uint64_t bd0 = make_desc(smem_u32(s.wres), 16, 16 * p.kc, p.kc);
          const uint32_t a_tl = a_tile >> 4, a_dy = (uint32_t)(16 * p.kc * 2) >> 4, w_tl = wtile >> 4;
          uint32_t first = 0;
#pragma unroll
          for (int j = 0; j < 3; ++j)
            for (int cb = 0; cb < p.kpt; ++cb) {
              const uint64_t aj = ad0 + (uint32_t)((j * p.kpt + cb) * a_tl);
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi) {
                const uint64_t ad = aj + (uint32_t)(dyi * a_dy);
                const uint64_t bd = bd0 + (uint32_t)(((j * 3 + dyi) * p.kpt + cb) * w_tl);
                if (p.kc == 32) {
                  umma_f16(tacc, ad, bd, idesc, first, leader);
                  umma_f16(tacc, ad + 2, bd + 2, idesc, 1, leader);
                } else {
                  umma_f16(tacc, ad, bd, idesc, first, leader);
                  umma_f16(tacc, ad + 2, bd + 2, idesc, 1, leader);
                  umma_f16(tacc, ad + 4, bd + 4, idesc, 1, leader);
                  umma_f16(tacc, ad + 6, bd + 6, idesc, 1, leader);
                }
                first = 1;
              }
            }
          umma_commit(&s.empty[st], leader);
          umma_commit(&s.done[as], leader);
        }
      } else
      // k-block-ring mode, two issuers: warp 1 takes the even k-blocks (even smem stages), warp 6 the
      // odd ones, each into its OWN accumulator (columns (2*pair + issuer)*128); the epilogue adds the
      // two partial tiles.  Every mbarrier is still used by one warp in order (stage count is even),
      // both wait the same acc_empty phase, and done[pair] collects one commit from each.
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
        const int ap = li & 1, aph = (li >> 1) & 1;
        mbar_wait(&s.acc_empty[ap], aph ^ 1);               // epilogue has drained this accumulator pair
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)((2 * ap + mine) * kMaxBN);
        uint32_t first = 0;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          if ((g & 1) != mine) continue;
          const int st = g % p.stages, ph = (g / p.stages) & 1;
          mbar_wait(&s.full[st], ph);
          tc_fence_after();
          const int ndy = p.rowshare ? 3 : 1;
          for (int dyi = 0; dyi < ndy; ++dyi) {
            // row-shared: the dy tap's 128 rows start dyi*16 pixels (= whole swizzle atoms) into the tile
            const uint64_t ad = make_desc(smem_u32(s.a(st)) + dyi * 16 * p.kc * 2, 16, 16 * p.kc, p.kc);   // SBO = 8 rows
            const uint64_t bd = make_desc(smem_u32(s.b(st)) + dyi * wtile, 16, 16 * p.kc, p.kc);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)   // +32 bytes along K inside the swizzle row
              if (k * 16 < p.kc) { umma_f16(tacc, ad + 2 * k, bd + 2 * k, idesc, first, leader); first = 1; }
          }
          umma_commit(&s.empty[st], leader);
        }
        umma_commit(&s.done[ap], leader);
      }
    }
  } else if ((warp >= 2 && warp <= 5) || warp >= 7) {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31  (= tile rows).  FOUR groups of four warps
    // (2-5, 7-10, 11-14, 15-18): groups alternate tiles (li & 1) and split a tile's 16-column chunks
    // (even / odd chunks).  ncu showed the small-channel layers epilogue-bound - a single group busy
    // 100 % of the time at one dependent instruction per ~4 cycles, the MMA issuers stalled on
    // acc_empty, tensor pipe 16 % active.  Tile li uses accumulator (pair) li % 4 (li % 2), so the two
    // groups with parity li & 1 are the only waiters of its barriers (8 arrivals free an accumulator).
    const int grp = warp <= 5 ? 0 : (warp - 7) / 4 + 1;
    const int eg = grp & 1, half = grp >> 1;
    const int q = warp % 4;
    const int r = q * 32 + lane;
    const int ni = r % p.wt, mi = (r / p.wt) % p.ht, bi = r / (p.wt * p.ht);
    // Folded epilogue constants: for gain > 0, lrelu(v) * gain = lrelu(v * gain), and for 0 < slope <= 1,
    // lrelu(t) = max(t, slope * t): three instructions per element (fma, mul, max).  ncu had this kernel
    // at 28 thread-instructions per output element with the issue slots 58 % busy - tile decode by
    // division and a four-step epilogue - which is what bounded the small-channel layers.
    const float asg = p.acc_scale * p.gain, bsg = p.bias_scale * p.gain;
    // one 16-column chunk per thread and tile (BN <= 32, a single channel tile): its bias and the current
    // image's row scales stay in registers
    // (also the blocked x2 form with 16 / 32 real channels: a thread's chunks, 32 columns apart, all map to the
    // same 16 channels)
    const bool cache = p.n_tiles == 1 && ((p.cblk == 0 && p.BN <= 32) || (p.cblk >= 16 && (p.cperiod == 16 || p.cperiod == 32)));
    const int co = p.cblk ? (half * 16) % p.cperiod : half * 16;       // first channel of this thread's chunks
    const int crow_c = p.cblk ? p.cperiod : p.Cout;
    float csc[16], cbi[16];
    int cached_b = -1;
    if (cache) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int o = co + i;
        cbi[i] = (bias && o < crow_c) ? bias[o] * bsg : 0.f;
        csc[i] = asg;
      }
    }
    int li = 0;
    if constexpr (LEAN) if (p.rowshare == 5) {
      // accumulator row r = m * 16 + n' (8 rows of 16 stored pixels): lane n' = 1..14 owns output column n0 + n' - 1 and
      // adds dx = 0 from its left neighbour's row, dx = 1 from its own, dx = 2 from its right neighbour's
      const int m = r >> 4, np = r & 15;
      const int nacc = 3 * p.BN <= kMaxBN ? kAccStages : 2;
      const uint32_t acc_stride = (uint32_t)(kAccStages * kMaxBN / nacc);
      const float slope = p.slope;
      const int per_img = p.tiles_w * p.tiles_h;
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
        if ((li & 1) != eg) continue;
        const int b = tile / per_img, tr = tile - b * per_img;
        const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
        const int n = tw * 14 + np - 1, mm = th * 8 + m;
        const bool live = np >= 1 && np <= 14 && n < p.tiles_w_px;
        bf16* const yp = reinterpret_cast<bf16*>(y) + (long long)b * p.ys_n + (long long)mm * p.ys_h + (long long)n * p.ys_w;
        if (cache && rowscale && b != cached_b) {
          const float4* rs = reinterpret_cast<const float4*>(rowscale + (long long)b * 32 + half * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 sc = rs[i];
            csc[4 * i] = sc.x * asg; csc[4 * i + 1] = sc.y * asg; csc[4 * i + 2] = sc.z * asg; csc[4 * i + 3] = sc.w * asg;
          }
          cached_b = b;
        }
        const int as = li % nacc, aph = (li / nacc) & 1;
        mbar_wait(&s.done[as], aph);
        tc_fence_after();
        const uint32_t tcol = tlane + (uint32_t)as * acc_stride;
        for (int c = half * 16; c < p.BN; c += 32) {
          uint32_t v0[16], v1[16], v2[16];
          tmem_ld16(tcol + (uint32_t)c, v0);
          tmem_ld16(tcol + (uint32_t)(p.BN + c), v1);
          tmem_ld16(tcol + (uint32_t)(2 * p.BN + c), v2);
          tmem_ld_wait();
          if (c + 32 >= p.BN) {
            // last chunk is in registers: hand the accumulator back before the arithmetic and the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acc_empty[as]);
          }
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float lft = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[i]), 1);
            const float rgt = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[i]), 1);
            f[i] = lft + __uint_as_float(v1[i]) + rgt;
          }
          if (live) {
            if (cache) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], csc[i], cbi[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float4 sc = make_float4(asg, asg, asg, asg), bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rowscale) {
                  const float4 t4 = *reinterpret_cast<const float4*>(rowscale + (long long)b * p.Cout + c + 4 * i);
                  sc = make_float4(t4.x * asg, t4.y * asg, t4.z * asg, t4.w * asg);
                }
                if (bias) {
                  const float4 t4 = *reinterpret_cast<const float4*>(bias + c + 4 * i);
                  bb = make_float4(t4.x * bsg, t4.y * bsg, t4.z * bsg, t4.w * bsg);
                }
                f[4 * i] = fmaf(f[4 * i], sc.x, bb.x); f[4 * i + 1] = fmaf(f[4 * i + 1], sc.y, bb.y);
                f[4 * i + 2] = fmaf(f[4 * i + 2], sc.z, bb.z); f[4 * i + 3] = fmaf(f[4 * i + 3], sc.w, bb.w);
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], f[i] * slope);
            Vec16<bf16> o0v, o1v;
            o0v.pack(f); o1v.pack(f + 8);
            if (!(p.dbg & 4)) { o0v.store(yp + c); o1v.store(yp + c + 8); }
          }
        }
      }
      li = -1;                                                // (done: skip the other forms below)
    }
    if (li < 0) {
    } else if constexpr (LEAN) {
      // per-thread constants: element offset of (row of the tile, 16-channel half), tile steps
      bf16* const ybase = reinterpret_cast<bf16*>(y) + (long long)mi * p.ys_h + (long long)ni * p.ys_w + half * 16;
      const long long wstep = (long long)p.wt * p.ys_w, hstep = (long long)p.ht * p.ys_h;
      const float slope = p.slope;
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
        if ((li & 1) != eg) continue;
        const int tw = tile & (p.tiles_w - 1), th = (tile >> p.lw) & (p.tiles_h - 1), tb = tile >> (p.lw + p.lh);
        const int b = tb * p.nt + bi;
        const bool live = b < p.N;
        if (rowscale && live && b != cached_b) {
          const float4* rs = reinterpret_cast<const float4*>(rowscale + (long long)b * 32 + half * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 sc = rs[i];
            csc[4 * i] = sc.x * asg; csc[4 * i + 1] = sc.y * asg; csc[4 * i + 2] = sc.z * asg; csc[4 * i + 3] = sc.w * asg;
          }
          cached_b = b;
        }
        const int as = li % kAccStages, aph = (li / kAccStages) & 1;
        mbar_wait(&s.done[as], aph);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(tcol + (uint32_t)(as * kMaxBN), v);
        tmem_ld_wait();
        // the accumulator is in registers: hand the stage back before the arithmetic and the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.acc_empty[as]);
        if (live) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float t = fmaf(__uint_as_float(v[i]), csc[i], cbi[i]);
            f[i] = fmaxf(t, t * slope);
          }
          bf16* yp = ybase + (long long)b * p.ys_n + th * hstep + tw * wstep;
          Vec16<bf16> o0v, o1v;
          o0v.pack(f); o1v.pack(f + 8);
          if (!(p.dbg & 4)) { o0v.store(yp); o1v.store(yp + 8); }
        }
      }
    } else
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
      if ((li & 1) != eg) continue;
      int nt_i = 0, t = tile;
      if (p.n_tiles > 1) { nt_i = tile % p.n_tiles; t = tile / p.n_tiles; }
      const int tw = t & (p.tiles_w - 1), th = (t >> p.lw) & (p.tiles_h - 1), tb = t >> (p.lw + p.lh);
      const int n0 = tw * p.wt, m0 = th * p.ht, b0 = tb * p.nt, o0 = nt_i * p.BN;
      const int b = b0 + bi;
      const bool live = b < p.N;
      const long long pix = (long long)b * p.ys_n + (long long)((m0 + mi) * p.os + p.py) * p.ys_h +
                            (long long)((n0 + ni) * p.os + p.px) * p.ys_w;
      const float nzg = (p.noise && live) ? p.noise[(long long)((m0 + mi) * p.os + p.py) * p.OW + (n0 + ni) * p.os + p.px] *
                                            p.noise_scale * p.gain : 0.f;
      if (cache && rowscale && live && b != cached_b) {
        const float4* rs = reinterpret_cast<const float4*>(rowscale + (long long)b * crow_c + co);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (co + 4 * i < crow_c) {
            const float4 sc = rs[i];
            csc[4 * i] = sc.x * asg; csc[4 * i + 1] = sc.y * asg; csc[4 * i + 2] = sc.z * asg; csc[4 * i + 3] = sc.w * asg;
          }
        }
        cached_b = b;
      }
      // resident modes: 4 single accumulators; ring mode: 2 pairs of partial accumulators
      const bool ring = p.rowshare < 2;
      const int as = ring ? (li & 1) : li % kAccStages, aph = ring ? (li >> 1) & 1 : (li / kAccStages) & 1;
      mbar_wait(&s.done[as], aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ring ? 2 * as : as) * kMaxBN);
      // which issuers contributed k-blocks to this tile (both unless the tile has a single k-block)
      const int g0 = li * nkb;
      const bool has0 = !ring || nkb > 1 || (g0 & 1) == 0, has1 = ring && (nkb > 1 || (g0 & 1) == 1);
      // 16 accumulator columns -> scale, bias, activation, residual, store
      auto finish16 = [&](const uint32_t* v, int oc) {   // oc: first accumulator column of the chunk
        if (live && p.cblk == 4) {
          // narrow blocked form (2-channel x2 transposed conv): columns 0..7 = (py, px, ch); each py half is one
          // float4 = the two horizontally adjacent output pixels x 2 channels
          if (oc == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float f[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int o = i & 1;
                float t = __uint_as_float(v[4 * h + i]) * asg;
                if (rowscale) t *= rowscale[(long long)b * 2 + o];
                if (bias) t = fmaf(bias[o], bsg, t);
                f[i] = fmaxf(t, t * p.slope);
              }
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + pix + (long long)h * p.ys_blk) =
                  make_float4(f[0], f[1], f[2], f[3]);
            }
          }
        } else if (live && oc < p.Cout) {
          // o: channel whose rowscale / bias apply; yo: element offset of the chunk inside the pixel
          const int o = p.cblk ? oc % p.cperiod : oc;
          // (haloed x2 form: accumulator columns hold the phases in the order 0, 1, 3, 2)
          if (p.rowshare == 4 && oc >= 2 * p.cperiod) oc ^= p.cperiod;
          const long long yo = p.cblk ? (long long)(oc / p.cblk) * p.ys_blk + oc % p.cblk : oc;
          const int crow = p.cblk ? p.cperiod : p.Cout;
          float f[16];
          if (cache) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaf(__uint_as_float(v[i]), csc[i], cbi[i] + nzg);
          } else {
            if (rowscale) {
              const float4* rs = reinterpret_cast<const float4*>(rowscale + (long long)b * crow + o);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 sc = rs[i];
                f[4 * i] = __uint_as_float(v[4 * i]) * (sc.x * asg); f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) * (sc.y * asg);
                f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) * (sc.z * asg); f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) * (sc.w * asg);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * asg;
            }
            if (bias) {
              const float4* bp = reinterpret_cast<const float4*>(bias + o);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 bb = bp[i];
                f[4 * i] = fmaf(bb.x, bsg, f[4 * i]); f[4 * i + 1] = fmaf(bb.y, bsg, f[4 * i + 1]);
                f[4 * i + 2] = fmaf(bb.z, bsg, f[4 * i + 2]); f[4 * i + 3] = fmaf(bb.w, bsg, f[4 * i + 3]);
              }
            }
            if (p.noise) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += nzg;
            }
          }
          if (p.slope != 1.f) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], f[i] * p.slope);
          }
          if (p.y_f32) {
            float* yp = reinterpret_cast<float*>(y) + pix + yo;
            if (residual) {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(residual) + pix + yo);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 rr = rp[i];
                f[4 * i] += rr.x; f[4 * i + 1] += rr.y; f[4 * i + 2] += rr.z; f[4 * i + 3] += rr.w;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(yp)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            bf16* yp = reinterpret_cast<bf16*>(y) + pix + yo;
            if (residual) {
              const bf16* rp = reinterpret_cast<const bf16*>(residual) + pix + yo;
              Vec16<bf16> r0, r1;
              r0.load(rp); r1.load(rp + 8);
              float g[16];
              r0.unpack(g); r1.unpack(g + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += g[i];
            }
            Vec16<bf16> o0v, o1v;
            o0v.pack(f); o1v.pack(f + 8);
            o0v.store(yp); o1v.store(yp + 8);
          }
        }
      };
      for (int c = half * 16; c < p.BN; c += 32) {
        uint32_t va[16];
        if (has0) {
          tmem_ld16(trow + c, va);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) va[i] = 0u;
        }
        if (has1) {
          uint32_t wa[16];
          tmem_ld16(trow + kMaxBN + c, wa);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) va[i] = __float_as_uint(__uint_as_float(va[i]) + __uint_as_float(wa[i]));
        }
        tmem_ld_wait();
        finish16(va, o0 + c);
      }
      // this warp is done reading the TMEM stage: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kAccStages * kMaxBN>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel: one CTA = (128 out-channels x BN in-channels) of one tap over a K-split
// ---------------------------------------------------------------------------------------------
constexpr int kWgABytes = kTileM * 128 * 2;   // G tile: 128 px x 128 ch = two 64-channel boxes
constexpr int kWgBBytes = kTileM * kMaxBN * 2;

__global__ void __launch_bounds__(kThreads)
tapconv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmg, const __grid_constant__ CUtensorMap tmx,
                        const WgParams p, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  const Smem s = carve(smem_raw, kStages, kWgABytes, kWgBBytes);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int o0 = (blockIdx.x / p.ctiles) * 128, c0 = (blockIdx.x % p.ctiles) * p.BN;
  const int tap = blockIdx.y;
  const int t_begin = blockIdx.z * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int nkb = t_end - t_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmg);
    tma_prefetch_desc(&tmx);
    for (int i = 0; i < kStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.done[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kMaxBN>(s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;
  const int nbx = p.BN / p.kcx, nbg = 128 / p.kcg;   // TMA boxes of the X and G operands
  const uint32_t gbox = kTileM * p.kcg * 2, xbox = kTileM * p.kcx * 2;   // bytes per box

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();
      const uint32_t tx_bytes = kWgABytes + nbx * xbox;
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % kStages, ph = (kb / kStages) & 1;
        mbar_wait(&s.empty[st], ph ^ 1);
        int t = t_begin + kb;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tb = t / p.tiles_h;
        const int n0 = tw * p.wt, m0 = th * p.ht, b0 = tb * p.nt;
        mbar_expect_tx(&s.full[st], tx_bytes, leader);
        for (int j = 0; j < nbg; ++j)      // boxes past Cout are out of range: TMA zero-fills them
          tma_load_4d(s.a(st) + j * gbox, &tmg, &s.full[st], o0 + p.kcg * j, n0 * p.os + p.px, m0 * p.os + p.py, b0, leader);
        for (int j = 0; j < nbx; ++j)
          tma_load_4d(s.b(st) + j * xbox, &tmx, &s.full[st], c0 + p.kcx * j, n0 * p.is + p.dx[tap],
                      m0 * p.is + p.dy[tap], b0, leader);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc(p.BN, true, true);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % kStages, ph = (kb / kStages) & 1;
        mbar_wait(&s.full[st], ph);
        tc_fence_after();
        // MN-major SWIZZLE_128B: 64-channel x 8-pixel atoms of 1024 B; LBO = next 64 channels (one box),
        // SBO = next 8 pixels; one MMA consumes 16 pixels = 2048 B along K.
        const uint64_t ad = make_desc(smem_u32(s.a(st)), gbox, 16 * p.kcg, p.kcg);
        const uint64_t bd = make_desc(smem_u32(s.b(st)), xbox, 16 * p.kcx, p.kcx);
#pragma unroll
        for (int k = 0; k < kTileM / 16; ++k)   // 16 pixels = 16 rows of kc*2 bytes along K
          umma_f16(tmem_base, ad + (uint64_t)(k * 2 * p.kcg), bd + (uint64_t)(k * 2 * p.kcx), idesc, (kb | k) != 0, leader);
        umma_commit(&s.empty[st], leader);
      }
      umma_commit(&s.done[0], leader);
    }
  } else {
    const int q = warp % 4;
    const int o = o0 + q * 32 + lane;
    mbar_wait(&s.done[0], 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float* row = dw + (long long)o * p.w_ld + (long long)p.wtap[tap] * p.Cin + c0;
    // deterministic mode: each epilogue warp owns 32 rows of the tile; the K-splits (blockIdx.z) add in order
    int* sem = p.sems ? p.sems + ((blockIdx.y * gridDim.x + blockIdx.x) * 4 + q) : nullptr;
    if (sem) {
      if (lane == 0) det_wait_turn(sem, blockIdx.z);
      __syncwarp();
    }
    for (int c = 0; c < p.BN; c += 16) {
      uint32_t v[16];
      tmem_ld16(trow + c, v);
      tmem_ld_wait();
      if (o < p.Cout && nkb > 0) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          if (c0 + c + i < p.Cin) {
            const float4 add = make_float4(__uint_as_float(v[i]) * p.scale, __uint_as_float(v[i + 1]) * p.scale,
                                           __uint_as_float(v[i + 2]) * p.scale, __uint_as_float(v[i + 3]) * p.scale);
            if (sem) det_add4(row + c + i, add);
            else atomicAdd(reinterpret_cast<float4*>(row + c + i), add);
          }
        }
      }
    }
    if (sem) {
      __threadfence();
      __syncwarp();
      if (lane == 0) det_pass_turn(sem, blockIdx.z, gridDim.z);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kMaxBN>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel for small channel counts ("taps along N"): one CTA accumulates ALL taps
//   D[128 out-ch x (ntaps*Cin)] += G[128 px x 128 ch]^T * [X_tap0 | X_tap1 | ...][128 px x ntaps*Cin]
// The tap tiles sit side by side in shared memory, one TMA box each, so a single MN-major B
// descriptor (LBO = box bytes) spans them: each 16-pixel K step is 1-2 wide MMAs (N <= 256) instead
// of ntaps narrow ones, and the G tile is read once per step instead of once per tap.
// Needs ntaps*Cin <= 288 (TMEM columns, 2 smem stages): the 32-channel 3x3 layers and the 64->x
// transposed-conv phases of the 512/1024 models.
// ---------------------------------------------------------------------------------------------
constexpr int kTnStages = 2;
constexpr int kTnMaxN = 288;
constexpr int kTnXBytes = kTileM * kTnMaxN * 2;   // 72 KiB

__global__ void __launch_bounds__(kThreads, 1)
tapconv_wgrad_tn_kernel(const __grid_constant__ CUtensorMap tmg, const __grid_constant__ CUtensorMap tmx,
                        const WgParams p, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  const Smem s = carve(smem_raw, kTnStages, kWgABytes, kTnXBytes);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int o0 = blockIdx.x * 128;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int nkb = t_end - t_begin;
  const int ntot = p.ntaps * p.Cin;                 // accumulator columns
  const int bpt = p.Cin / p.kcx;                    // X boxes per tap

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmg);
    tma_prefetch_desc(&tmx);
    for (int i = 0; i < kTnStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.done[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;
  const int nbg = 128 / p.kcg;
  const uint32_t gbox = kTileM * p.kcg * 2, xbox = kTileM * p.kcx * 2;

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();
      const uint32_t tx_bytes = kWgABytes + p.ntaps * bpt * xbox;
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % kTnStages, ph = (kb / kTnStages) & 1;
        mbar_wait(&s.empty[st], ph ^ 1);
        int t = t_begin + kb;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tb = t / p.tiles_h;
        const int n0 = tw * p.wt, m0 = th * p.ht, b0 = tb * p.nt;
        mbar_expect_tx(&s.full[st], tx_bytes, leader);
        for (int j = 0; j < nbg; ++j)
          tma_load_4d(s.a(st) + j * gbox, &tmg, &s.full[st], o0 + p.kcg * j, n0 * p.os + p.px, m0 * p.os + p.py, b0, leader);
        for (int tap = 0; tap < p.ntaps; ++tap)
          for (int j = 0; j < bpt; ++j)
            tma_load_4d(s.b(st) + (tap * bpt + j) * xbox, &tmx, &s.full[st], p.kcx * j, n0 * p.is + p.dx[tap],
                        m0 * p.is + p.dy[tap], b0, leader);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_one();
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % kTnStages, ph = (kb / kTnStages) & 1;
        mbar_wait(&s.full[st], ph);
        tc_fence_after();
        const uint64_t ad = make_desc(smem_u32(s.a(st)), gbox, 16 * p.kcg, p.kcg);
        for (int n0c = 0; n0c < ntot; n0c += 256) {          // N chunks of whole boxes, <= 256 columns
          const int nn = min(256, ntot - n0c);
          const uint32_t idesc = make_idesc(nn, true, true);
          const uint64_t bd = make_desc(smem_u32(s.b(st)) + (n0c / p.kcx) * xbox, xbox, 16 * p.kcx, p.kcx);
#pragma unroll
          for (int k = 0; k < kTileM / 16; ++k)
            umma_f16(tmem_base + (uint32_t)n0c, ad + (uint64_t)(k * 2 * p.kcg), bd + (uint64_t)(k * 2 * p.kcx), idesc,
                     (kb | k) != 0, leader);
        }
        umma_commit(&s.empty[st], leader);
      }
      umma_commit(&s.done[0], leader);
    }
  } else {
    const int q = warp % 4;
    const int o = o0 + q * 32 + lane;
    mbar_wait(&s.done[0], 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    int* sem = p.sems ? p.sems + (blockIdx.x * 4 + q) : nullptr;      // K-splits (blockIdx.y) add in order
    if (sem) {
      if (lane == 0) det_wait_turn(sem, blockIdx.y);
      __syncwarp();
    }
    for (int c = 0; c < ntot; c += 16) {
      uint32_t v[16];
      tmem_ld16(trow + c, v);
      tmem_ld_wait();
      if (o < p.Cout && nkb > 0) {
        const int tap = c / p.Cin, ci = c - tap * p.Cin;      // 16 | Cin, so a 16-column group stays in one tap
        float* row = dw + (long long)o * p.w_ld + (long long)p.wtap[tap] * p.Cin + ci;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 add = make_float4(__uint_as_float(v[i]) * p.scale, __uint_as_float(v[i + 1]) * p.scale,
                                         __uint_as_float(v[i + 2]) * p.scale, __uint_as_float(v[i + 3]) * p.scale);
          if (sem) det_add4(row + i, add);
          else atomicAdd(reinterpret_cast<float4*>(row + i), add);
        }
      }
    }
    if (sem) {
      __threadfence();
      __syncwarp();
      if (lane == 0) det_pass_turn(sem, blockIdx.y, gridDim.y);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------
// weight-gradient kernel for 3x3 stride-1 layers with 32 / 64 channels, TRANSPOSED and HALOED:
//   D^T[(dx, c) x (dy, o)] += X_shift(dy, dx)[px x c]^T * G[px x o]      over the pixels of a K-split
// The layer input is loaded ONCE per 8 x 16 lattice tile as the (8+2) x (16+2)-pixel box of the forward
// kernel's haloed mode; a tap is a descriptor start offset into it.  Two more descriptor facts make the
// shape efficient:
//   * the dx = 0,1,2 views differ by one stored pixel, so with LBO = one pixel row they ARE the consecutive
//     M blocks of a single MN-major A operand: one MMA covers three taps (M = 3 x 32 of 128 rows for 32
//     channels; for 64 channels taps dx = 0,1 fill M = 128 and dx = 2 takes a second MMA);
//   * putting (dx, c) on M and o on N avoids padding a 32-channel gradient to M = 128 (the layout that made
//     the previous small-channel kernel tensor-bound at 4x the useful work).
// Rows of A past the last tap read whatever follows in shared memory: each D row depends on its own A row
// only, and those rows are never stored.  ncu on the previous kernel: 27.9 GB L2->SM per launch for 4.3 GB of
// operands (every tap re-loaded its own shifted tile); here each pixel of X and G enters shared memory once.
// ---------------------------------------------------------------------------------------------
constexpr int kWhStagesMax = 6;

struct WhParams {
  int N, Cin, Cout;
  int tiles_w, tiles_h, lw, lh, tiles_total, tiles_per_split;
  int kcx, kcg;                                 // channels per swizzle row of X / G (= Cin, Cout: 32 or 64)
  int stages, x_bytes, g_bytes;                 // per-stage bytes (x_bytes padded to the swizzle period)
  int grp_wtap[3][3];                           // [dx+1][dy+1] -> weight tap index
  long long w_ld;
  float scale;
  int* sems;
  int dyn;                                      // dy on N: one MMA per K step covers the three dy taps too (see the kernel)
};

__global__ void __launch_bounds__(kThreads, 1)
tapconv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmg, const __grid_constant__ CUtensorMap tmx,
                          const WhParams p, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  const Smem s = carve(smem_raw, p.stages, p.x_bytes, p.g_bytes);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int t_begin = blockIdx.x * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int nkb = t_end - t_begin;
  constexpr int WT = 8, HT = 16;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmg);
    tma_prefetch_desc(&tmx);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.done[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;
  const uint32_t rowx = (uint32_t)p.kcx * 2, rowg = (uint32_t)p.kcg * 2;       // bytes of one stored pixel
  const int nh = p.Cin == 64 ? 2 : 1;                                           // MMAs per (dy, K-step)

  if (warp == 0) {
    const uint32_t leader = elect_one();
    // dyn: the row halo moves from X to G (X box (WT+2) x HT, G box WT x (HT+2))
    const uint32_t tx_bytes = p.dyn ? (uint32_t)(WT + 2) * HT * rowx + (uint32_t)WT * (HT + 2) * rowg
                                    : (uint32_t)(WT + 2) * (HT + 2) * rowx + (uint32_t)WT * HT * rowg;
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % p.stages, ph = (kb / p.stages) & 1;
      mbar_wait(&s.empty[st], ph ^ 1);
      const int t = t_begin + kb;
      const int tw = t & (p.tiles_w - 1), th = (t >> p.lw) & (p.tiles_h - 1), tb = t >> (p.lw + p.lh);
      mbar_expect_tx(&s.full[st], tx_bytes, leader);
      tma_load_4d(s.a(st), &tmx, &s.full[st], 0, tw * WT - 1, th * HT - (p.dyn ? 0 : 1), tb, leader);
      tma_load_4d(s.b(st), &tmg, &s.full[st], 0, tw * WT, th * HT - (p.dyn ? 1 : 0), tb, leader);
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(p.Cout, true, true);
    const uint32_t row16 = rowx >> 4, pitch16 = (uint32_t)(WT + 2) * row16;
    if (p.dyn) {
      // dW[o][dy,dx][c] = sum_q X[q + (0, dx-1), c] G[q - (dy-1, 0), o]: with q = the X row, the dy = 2, 1, 0 views of G
      // start one stored lattice row (8 pixels) apart, so - like the dx views of X on M - they are the consecutive
      // N blocks of ONE MN-major B operand (LBO = 8 pixels): D^T[(dx, c) x (2-dy, o)], 8 (16 for 64 channels) MMAs
      // of N = 3 * Cout per tile instead of 24 (48) of N = Cout.  The N = Cout form ran at ~53-61 cycles per MMA
      // whatever N was - streaming the 4 KB A operand out of shared memory - and bounded the kernel (1.31 ms for
      // 32 -> 32 @1024^2 against 0.66 ms of HBM time).
      const uint32_t idesc3 = make_idesc(3 * p.Cout, true, true);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % p.stages, ph = (kb / p.stages) & 1;
        mbar_wait(&s.full[st], ph);
        tc_fence_after();
        const uint64_t ad0 = make_desc(smem_u32(s.a(st)), rowx, (uint32_t)(WT + 2) * rowx, p.kcx);
        const uint64_t bd0 = make_desc(smem_u32(s.b(st)), 8 * rowg, 8 * rowg, p.kcg);
#pragma unroll
        for (int k = 0; k < HT / 2; ++k) {
          const uint64_t ad = ad0 + (uint32_t)(2 * k * pitch16);
          const uint64_t bd = bd0 + (uint32_t)(k * 16 * (rowg >> 4));
          umma_f16(tmem_base, ad, bd, idesc3, (kb | k) != 0, leader);
          if (nh == 2) umma_f16(tmem_base + (uint32_t)(3 * p.Cout), ad + 2 * row16, bd, idesc3, (kb | k) != 0, leader);
        }
        umma_commit(&s.empty[st], leader);
      }
    } else
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % p.stages, ph = (kb / p.stages) & 1;
      mbar_wait(&s.full[st], ph);
      tc_fence_after();
      // A = X: MN-major, M blocks (taps dx, dx+1, ..) one stored pixel apart, 8-pixel K groups one stored row apart
      const uint64_t ad0 = make_desc(smem_u32(s.a(st)), rowx, (uint32_t)(WT + 2) * rowx, p.kcx);
      // B = G: MN-major, one channel box, 8-pixel K groups contiguous
      const uint64_t bd0 = make_desc(smem_u32(s.b(st)), (uint32_t)WT * HT * rowg, 8 * rowg, p.kcg);
#pragma unroll
      for (int k = 0; k < HT / 2; ++k) {                   // one MMA K-step = 16 pixels = two lattice rows
        const uint64_t bd = bd0 + (uint32_t)(k * 16 * (rowg >> 4));
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
          const uint64_t ad = ad0 + (uint32_t)((dyi + 2 * k) * pitch16);
          umma_f16(tmem_base + (uint32_t)(dyi * nh * p.Cout), ad, bd, idesc, (kb | k) != 0, leader);
          if (nh == 2) umma_f16(tmem_base + (uint32_t)((dyi * 2 + 1) * p.Cout), ad + 2 * row16, bd, idesc, (kb | k) != 0, leader);
        }
      }
      umma_commit(&s.empty[st], leader);
    }
    umma_commit(&s.done[0], leader);
  } else {
    const int q = warp % 4;
    const int r = q * 32 + lane;                           // accumulator row = (tap dx, input channel c)
    mbar_wait(&s.done[0], 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    int* sem = p.sems ? p.sems + q : nullptr;              // deterministic mode: K-splits add in order
    if (sem) {
      if (lane == 0) det_wait_turn(sem, blockIdx.x);
      __syncwarp();
    }
    for (int dyi = 0; dyi < 3; ++dyi)
      for (int h = 0; h < nh; ++h) {
        const int dxi = nh == 2 ? 2 * h + r / 64 : r / 32;
        const int c = nh == 2 ? r % 64 : r % 32;
        const bool live = dxi < 3 && nkb > 0;
        // dyn: accumulator h holds the N blocks j = 2 - dy side by side
        float* col = dw + (long long)p.grp_wtap[live ? dxi : 0][dyi] * p.Cin + c;
        const uint32_t acol = p.dyn ? (uint32_t)((h * 3 + (2 - dyi)) * p.Cout) : (uint32_t)((dyi * nh + h) * p.Cout);
        for (int o0 = 0; o0 < p.Cout; o0 += 16) {
          uint32_t v[16];
          tmem_ld16(trow + acol + (uint32_t)o0, v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float* dst = col + (long long)(o0 + i) * p.w_ld;
              if (sem) det_add(dst, __uint_as_float(v[i]) * p.scale);
              else atomicAdd(dst, __uint_as_float(v[i]) * p.scale);
            }
          }
        }
      }
    if (sem) {
      __threadfence();
      __syncwarp();
      if (lane == 0) det_pass_turn(sem, blockIdx.x, gridDim.x);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 4-D map over a dense channels-last bf16 tensor [N, H, W, C]: box {64, wt, ht, nt} walked with
// element stride `es` along W and H.
int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int wt, int ht, int nt, int es, int kc,
                 int extra_rows = 0, int extra_cols = 0) {
  EncodeTiledFn enc = get_encode();
  LCGAN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(wt * es + extra_cols), (cuuint32_t)(ht * es + extra_rows), (cuuint32_t)nt};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LCGAN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (N=%d H=%d W=%d C=%d box=%d,%d,%d es=%d)",
              (int)r, N, H, W, C, wt, ht, nt, es);
  return 0;
}

int make_w_map(CUtensorMap* m, const void* base, int rows, long long ld, int box_rows, int kc) {
  EncodeTiledFn enc = get_encode();
  LCGAN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LCGAN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d (rows=%d ld=%lld)", (int)r, rows, ld);
  return 0;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// lattice tile {wt, ht, nt} with wt*ht*nt == 128
bool lattice_tile(int MW, int MH, int* wt, int* ht, int* nt) {
  if (!is_pow2(MW) || !is_pow2(MH)) return false;
  *wt = MW < 16 ? MW : 16;
  const int rem = kTileM / *wt;
  *ht = MH < rem ? MH : rem;
  *nt = rem / *ht;
  return true;
}

// dense channels-last; strides of size-1 dimensions are irrelevant (a [b,K] matrix viewed as [b,K,1,1])
bool dense_cl(int64_t sn, int64_t sh, int64_t sw, int64_t sc, int N, int H, int W, int C) {
  return sc == 1 && (W == 1 || sw == C) && (H == 1 || sh == (int64_t)W * C) && (N == 1 || sn == (int64_t)H * W * C);
}

int fwd_smem_bytes() { return 3 * (20 * 1024 + 3 * kBBytes) + 1024 + 256; }   // row-shared mode is the larger one

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
int wg_smem_bytes() { return kStages * (kWgABytes + kWgBBytes) + 1024 + 256; }
int tn_smem_bytes() { return kTnStages * (kWgABytes + kTnXBytes) + 1024 + 256; }

}  // namespace

static int tc_eligible(const lcgan_tapconv* d, bool narrow_out);
extern "C" int lcgan_tapconv_tc_eligible(const lcgan_tapconv* d) { return tc_eligible(d, false); }

// narrow_out: 8 output channels (the blocked form of a 2-channel x2 transposed conv), stored as float4 pieces
static int tc_eligible(const lcgan_tapconv* d, bool narrow_out) {
  if (!d) return 0;
  if (d->x_dtype != LCGAN_BF16) return 0;
  if (d->Cin % 32 != 0 || (narrow_out ? d->Cout != 8 : d->Cout % 16 != 0)) return 0;
  if (!dense_cl(d->xs_n, d->xs_h, d->xs_w, d->xs_c, d->N, d->IH, d->IW, d->Cin)) return 0;
  // output: channel-innermost, 16-byte aligned pixel rows (dense channels-last or any such strides)
  const int al = narrow_out ? 4 : 8;          // elements per 16 bytes of the narrowest store (f32 float4 / bf16 x 8)
  if (d->ys_c != 1 || (d->OW > 1 && d->ys_w % al != 0) || (d->OH > 1 && d->ys_h % al != 0) ||
      (d->N > 1 && d->ys_n % al != 0)) return 0;
  int wt, ht, nt;
  if (!lattice_tile(d->MW, d->MH, &wt, &ht, &nt)) return 0;
  if (d->is < 1 || d->is > 2 || d->os < 1 || d->os > 2) return 0;
  if (d->ntaps < 1 || d->ntaps > LCGAN_MAX_TAPS) return 0;
  return 1;
}

static int tapconv_tc_launch(const lcgan_tapconv* d, const void* x, const void* w2, void* y, const float* rowscale,
                             const float* bias, const void* residual, void* stream, int cblk, long long ys_blk,
                             int cperiod);

extern "C" int lcgan_tapconv_tc(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                                const float* rowscale, const float* bias, const void* residual, void* stream) {
  return tapconv_tc_launch(d, x, w2, y, rowscale, bias, residual, stream, 0, 0, 0);
}

// Blocked output channels: channel o of the contraction is stored at element offset
// (o / cblk) * ys_blk + o % cblk of its lattice point's output pixel and uses rowscale / bias of channel
// o % cperiod (rowscale is [N, cperiod]).  This runs conv_transpose2d(k3, s2, p1, op1) as ONE launch over
// the input lattice: 4 taps (the 2x2 input neighbourhood), Cout' = 4 Cout with the unused (phase, tap)
// weight blocks zero, cblk = 2 Cout (the two horizontally adjacent output pixels), ys_blk = one output row.
extern "C" int lcgan_tapconv_tc_blocked(const lcgan_tapconv* d, const void* x, const void* w2, void* y,
                                        const float* rowscale, const float* bias, int cblk, int64_t ys_blk,
                                        int cperiod, void* stream) {
  const bool narrow = d && cblk == 4 && cperiod == 2 && d->Cout == 8 && d->y_dtype == LCGAN_F32 && ys_blk % 4 == 0;
  LCGAN_CHECK(narrow || (d && cblk > 0 && cblk % 16 == 0 && cperiod > 0 && cperiod % 16 == 0 && d->Cout % cblk == 0 &&
                         cblk % cperiod == 0 && ys_blk % 8 == 0),
              "tapconv_tc_blocked: need cblk %% 16 == 0, cperiod %% 16 == 0, cperiod | cblk | Cout, ys_blk %% 8 == 0 "
              "(or the narrow form: Cout 8, cblk 4, cperiod 2, f32 output)");
  return tapconv_tc_launch(d, x, w2, y, rowscale, bias, nullptr, stream, cblk, ys_blk, cperiod);
}

static int tapconv_tc_launch(const lcgan_tapconv* d, const void* x, const void* w2, void* y, const float* rowscale,
                             const float* bias, const void* residual, void* stream, int cblk, long long ys_blk,
                             int cperiod) {
  LCGAN_CHECK(tc_eligible(d, cblk == 4), "tapconv_tc: descriptor not eligible for the tensor-core path");
  LCGAN_CHECK(d->w_dtype == LCGAN_BF16, "tapconv_tc: weights must be bf16");
  LCGAN_CHECK(d->colscale == nullptr, "tapconv_tc: colscale is only implemented by the pointwise thin kernel");
  LCGAN_CHECK(x && w2 && y, "tapconv_tc: null tensor pointer");
  LCGAN_CHECK(((uintptr_t)x % 16 == 0) && ((uintptr_t)w2 % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
              (d->w_ld % 8 == 0), "tapconv_tc: operands must be 16-byte aligned");
  TcParams p{};
  p.N = d->N; p.Cin = d->Cin; p.Cout = d->Cout;
  lattice_tile(d->MW, d->MH, &p.wt, &p.ht, &p.nt);
  p.tiles_w = d->MW / p.wt; p.tiles_h = d->MH / p.ht;
  p.is = d->is; p.os = d->os; p.py = d->py; p.px = d->px;
  p.kc = d->Cin % 64 == 0 ? 64 : 32;
  p.ntaps = d->ntaps; p.kpt = d->Cin / p.kc;
  for (int t = 0; t < d->ntaps; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.wtap[t] = d->wtap[t]; }
  p.BN = d->Cout >= kMaxBN ? kMaxBN : (d->Cout < 16 ? 16 : d->Cout);   // Cout % 16 == 0 (narrow form: 8 of 16 columns,
                                                                        // the weight box rows past Cout are zero-filled by TMA)
  p.ys_n = d->ys_n; p.ys_h = d->ys_h; p.ys_w = d->ys_w;
  p.y_f32 = d->y_dtype == LCGAN_F32;
  p.acc_scale = d->acc_scale; p.bias_scale = d->bias_scale; p.slope = d->slope; p.gain = d->gain;
  p.cblk = cblk; p.cperiod = cperiod; p.ys_blk = ys_blk;
  p.noise = d->noise; p.noise_scale = d->noise_scale; p.OW = d->OW;
  LCGAN_CHECK(!(d->noise && cblk), "tapconv_tc_blocked: no noise term in the blocked-output form");

  LCGAN_CHECK(d->gain > 0.f && d->slope > 0.f && d->slope <= 1.f, "tapconv_tc: needs gain > 0 and 0 < slope <= 1");
  // full 3x3 stride-1 tap set?
  bool full3x3 = false;
  if (d->ntaps == 9 && d->is == 1) {
    int seen = 0;
    for (int t = 0; t < 9; ++t) {
      const int dy = d->dy[t], dx = d->dx[t];
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1) { seen = -1; break; }
      p.grp_wtap[dx + 1][dy + 1] = d->wtap[t];
      seen |= 1 << ((dy + 1) * 3 + dx + 1);
    }
    full3x3 = seen == 0x1FF;
  }
  // full 3x3 stride-2 tap set (input offsets -1, 0, +1 around 2m, 2n)?
  bool s2_3x3 = false;
  if (d->ntaps == 9 && d->is == 2 && d->os == 1) {
    int seen = 0;
    for (int t = 0; t < 9; ++t) {
      const int dy = d->dy[t], dx = d->dx[t];
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1) { seen = -1; break; }
      p.grp_wtap[dx + 1][dy + 1] = d->wtap[t];
      seen |= 1 << ((dy + 1) * 3 + dx + 1);
    }
    s2_3x3 = seen == 0x1FF;
  }
  p.rowshare = 0;
  p.wres_bytes = 0;
  int smem_bytes = fwd_smem_bytes();
  const int budget = 3 * (20 * 1024 + 3 * kBBytes);        // bytes available for the operand ring
  bool up2_taps = cblk > 4 && d->ntaps == 4 && d->is == 1 && d->os == 1;
  for (int t = 0; up2_taps && t < 4; ++t) up2_taps = d->dy[t] == (t >> 1) && d->dx[t] == (t & 1) && d->wtap[t] == t;
  if (up2_taps && (d->Cin == 32 || d->Cin == 64) && d->Cout <= kMaxBN && d->Cout == 4 * cperiod && cblk == 2 * cperiod &&
      (cperiod & (cperiod - 1)) == 0 &&
      d->MW % 8 == 0 && d->MH % 16 == 0 && getenv("LCGAN_NO_UP2_HALO") == nullptr) {
    // x2 transposed conv in the blocked form, haloed: 8 x 16 input-lattice tiles, one (8+1) x (16+1) box per tile,
    // the 9 live (tap, phase) weight blocks resident, 4 phase accumulators of Cout columns each
    p.rowshare = 4;
    p.wt = 8; p.ht = 16; p.nt = 1;
    p.tiles_w = d->MW / p.wt; p.tiles_h = d->MH / p.ht;
    p.wres_bytes = (9 * cperiod * d->Cin * 2 + 1023) & ~1023;
    p.a_bytes = ((p.wt + 1) * (p.ht + 1) * p.kc * 2 + 1023) & ~1023;
    p.b_bytes = 0;
    p.stages = (kMaxSmem - 2048 - p.wres_bytes) / p.a_bytes;
    if (p.stages > 12) p.stages = 12;
    p.stages &= ~1;
    smem_bytes = 1024 + p.wres_bytes + p.stages * p.a_bytes + 256;
  } else if (s2_3x3 && d->Cin == 32 && d->Cout <= kMaxBN && d->MW % 8 == 0 && d->MH % 16 == 0 && d->IW == 2 * d->MW &&
             d->IH == 2 * d->MH && cblk == 0 && getenv("LCGAN_NO_S2_HALO") == nullptr) {
    // haloed stride-2 mode (see the producer): pixel pairs as 128-byte rows, one (8+1) x (32+1) box per tile
    p.rowshare = 6;
    p.wt = 8; p.ht = 16; p.nt = 1;
    p.tiles_w = d->MW / p.wt; p.tiles_h = d->MH / p.ht;
    p.wres_bytes = 9 * p.BN * d->Cin * 2;
    p.a_bytes = (9 * 33 * 128 + 1023) & ~1023;
    p.b_bytes = 0;
    p.stages = (kMaxSmem - 2048 - p.wres_bytes) / p.a_bytes;
    if (p.stages > 12) p.stages = 12;
    p.stages &= ~1;
    smem_bytes = 1024 + p.wres_bytes + p.stages * p.a_bytes + 256;
  } else if (full3x3 && (d->Cin == 32 || d->Cin == 64) && (d->Cout == 32 || d->Cout == 64) && d->MW >= 16 &&
             d->MH % 8 == 0 && d->os == 1 && d->px == 0 && d->py == 0 && cblk == 0 && !p.y_f32 && !residual && !d->noise &&
             getenv("LCGAN_DXN") != nullptr) {
    // dx-on-N mode (OPT-IN, LCGAN_DXN=1): 16 stored x 8 rows per tile (14 live output columns), one (16) x (8+2) box per
    // tile, weights resident as three [dx][Cout] x Cin operands (one per dy).  Measured on the B200 (32 -> 32 @1024^2,
    // batch 32): its MMA chain costs 0.11 ms where the nine-tap haloed mode's costs 0.67 ms, but the epilogue's two
    // shuffles per output and three TMEM loads per chunk make the epilogue skeleton alone 0.99 ms (0.43 ms there):
    // 1.46 ms against 1.06 ms end to end.  Kept, tested, for the TMA-store epilogue that would make it pay.
    p.rowshare = 5;
    p.wt = 16; p.ht = 8; p.nt = 1;
    p.tiles_w = (d->MW + 13) / 14; p.tiles_h = d->MH / 8;
    p.tiles_w_px = d->MW;
    p.wres_bytes = 9 * p.BN * d->Cin * 2;
    p.a_bytes = 16 * 10 * p.kc * 2;
    p.b_bytes = 0;
    p.stages = (kMaxSmem - 2048 - p.wres_bytes) / p.a_bytes;
    if (p.stages > 12) p.stages = 12;
    p.stages &= ~1;
    smem_bytes = 1024 + p.wres_bytes + p.stages * p.a_bytes + 256;
  } else if (full3x3 && (d->Cin == 32 || d->Cin == 64) && d->Cout <= kMaxBN && d->MW % 8 == 0 && d->MH % 16 == 0 &&
      9 * p.BN * d->Cin * 2 <= 80 * 1024 && getenv("LCGAN_NO_HALO") == nullptr) {
    // haloed small-channel mode: 8-wide x 16-tall lattice tiles, ONE (8+2) x (16+2) pixel box per tile serves all
    // nine taps, weights resident; stage stride padded to the swizzle period
    p.rowshare = 3;
    p.wt = 8; p.ht = 16; p.nt = 1;
    p.tiles_w = d->MW / p.wt; p.tiles_h = d->MH / p.ht;
    p.wres_bytes = 9 * p.BN * d->Cin * 2;
    p.a_bytes = ((p.wt + 2) * (p.ht + 2) * p.kc * 2 + 1023) & ~1023;
    p.b_bytes = 0;
    p.stages = (kMaxSmem - 2048 - p.wres_bytes) / p.a_bytes;
    if (p.stages > 12) p.stages = 12;
    p.stages &= ~1;                                         // even: see the two-issuer note in the kernel
    smem_bytes = 1024 + p.wres_bytes + p.stages * p.a_bytes + 256;
  } else if (full3x3 && p.wt == 16 && p.ht == 8 && p.nt == 1 && getenv("LCGAN_NO_ROWSHARE") == nullptr) {
    // row-shared mode: 16-wide, 8-tall, single-image tile; one tall A tile per dx serves the three dy taps
    p.rowshare = 1;
    if (d->Cout <= kMaxBN && 9 * p.BN * d->Cin * 2 <= 80 * 1024 && getenv("LCGAN_NO_RESIDENT") == nullptr) {
      // small-channel layers: resident weights, one stage per output tile
      p.rowshare = 2;
      p.wres_bytes = 9 * p.BN * d->Cin * 2;
      p.a_bytes = 3 * p.kpt * (p.ht + 2) * 16 * p.kc * 2;
      p.b_bytes = 0;
      p.stages = (budget - p.wres_bytes) / p.a_bytes;
      if (p.stages > 8) p.stages = 8;
      p.stages &= ~1;                                         // even: see the two-issuer note in the kernel
      if (p.stages < 2) { p.rowshare = 1; p.wres_bytes = 0; }
    }
    if (p.rowshare == 1 && p.kc == 64 && getenv("LCGAN_ROWSHARE_WIDE") == nullptr) p.rowshare = 0;   // measured slower on 64-ch k-blocks
  }
  if (p.rowshare >= 2) {
  } else if (p.rowshare) {
    p.a_bytes = (p.ht + 2) * 16 * p.kc * 2;                 // 20 KiB (kc=64) / 10 KiB (kc=32)
    p.b_bytes = 3 * p.BN * p.kc * 2;
    p.stages = 2;                                           // even: one issuer per stage parity
    if ((p.a_bytes + p.b_bytes) * 4 <= budget) p.stages = 4;
    if ((p.a_bytes + p.b_bytes) * 6 <= budget) p.stages = 6;
  } else {
    p.a_bytes = kABytes; p.b_bytes = kBBytes; p.stages = kFwdStages;
  }
  for (p.lw = 0; (1 << p.lw) < p.tiles_w; ++p.lw) {}
  for (p.lh = 0; (1 << p.lh) < p.tiles_h; ++p.lh) {}
  CUtensorMap tmx, tmw;
  if (p.rowshare == 6) {
    // [N, IH, IW/2, 64]: the same bytes with pixel pairs as the innermost 64 "channels"; box 64 x (8+1) x (32+1)
    if (int e = make_act_map(&tmx, x, d->N, d->IH, d->IW / 2, 64, 8, 32, 1, 1, 64, 1, 1)) return e;
  } else
  if (int e = make_act_map(&tmx, x, d->N, d->IH, d->IW, d->Cin, p.wt, p.ht, p.nt, d->is, p.kc,
                           p.rowshare == 4 ? 1 : (p.rowshare ? 2 : 0), p.rowshare == 4 ? 1 : (p.rowshare == 3 ? 2 : 0))) return e;
  const int tiles_img = (d->N + p.nt - 1) / p.nt;          // (the small-channel modes reset nt to 1)
  if (int e = make_w_map(&tmw, w2, d->Cout, d->w_ld, p.rowshare == 4 ? cperiod : p.BN, p.kc)) return e;

  static DeviceOnce once;
  const cudaError_t attr_err = (cudaError_t)once.run([] {
    cudaError_t attr_err = cudaSuccess;
    attr_err = cudaFuncSetAttribute(tapconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(tapconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    return (int)attr_err;
  });
  LCGAN_CHECK(attr_err == cudaSuccess, "tapconv_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(attr_err));
  p.n_tiles = (d->Cout + p.BN - 1) / p.BN;
  p.total_tiles = p.tiles_w * p.tiles_h * tiles_img * p.n_tiles;
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();   // persistent: one CTA per SM
  static const int dbg = getenv("LCGAN_TC_DEBUG") ? atoi(getenv("LCGAN_TC_DEBUG")) : 0;
  p.dbg = dbg;
  static const bool no_lean = getenv("LCGAN_TC_NO_LEAN") != nullptr;
  const bool lean = p.rowshare == 5 || (!no_lean && p.n_tiles == 1 && p.BN == 32 && d->Cout == 32 && p.cblk == 0 && p.rowshare >= 2 &&
                    !p.y_f32 && !residual && !p.noise && p.os == 1 && p.px == 0 && p.py == 0);
  if (lean)
    tapconv_tc_kernel<true><<<grid, kFwdThreads, smem_bytes, (cudaStream_t)stream>>>(tmx, tmw, p, y, rowscale, bias, residual);
  else
    tapconv_tc_kernel<false><<<grid, kFwdThreads, smem_bytes, (cudaStream_t)stream>>>(tmx, tmw, p, y, rowscale, bias, residual);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcgan_tapconv_wgrad_tc(const lcgan_tapconv* d, const void* x, const void* g, float* dw2, float scale,
                                      void* stream) {
  // here the descriptor's X side is the layer input and its Y side is the output gradient G
  LCGAN_CHECK(lcgan_tapconv_tc_eligible(d), "tapconv_wgrad_tc: descriptor not eligible");
  LCGAN_CHECK(d->y_dtype == LCGAN_BF16 && d->Cout % 32 == 0 &&
              dense_cl(d->ys_n, d->ys_h, d->ys_w, d->ys_c, d->N, d->OH, d->OW, d->Cout),
              "tapconv_wgrad_tc: G must be dense channels-last bf16 with Cout %% 32 == 0");
  LCGAN_CHECK(x && g && dw2 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)dw2 % 16 == 0) &&
              d->w_ld % 4 == 0, "tapconv_wgrad_tc: bad pointers/alignment");
  WgParams p{};
  p.N = d->N; p.Cin = d->Cin; p.Cout = d->Cout;
  lattice_tile(d->MW, d->MH, &p.wt, &p.ht, &p.nt);
  p.tiles_w = d->MW / p.wt; p.tiles_h = d->MH / p.ht;
  const int tiles_b = (d->N + p.nt - 1) / p.nt;
  p.tiles_total = p.tiles_w * p.tiles_h * tiles_b;
  p.is = d->is; p.os = d->os; p.py = d->py; p.px = d->px;
  p.ntaps = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) { p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.wtap[t] = d->wtap[t]; }
  p.kcx = d->Cin % 64 == 0 ? 64 : 32;
  p.kcg = d->Cout % 64 == 0 ? 64 : 32;
  p.BN = d->Cin >= kMaxBN ? kMaxBN : d->Cin;   // multiple of kcx, <= 128 (a partial last tile is zero-filled)
  p.ctiles = (d->Cin + p.BN - 1) / p.BN;
  const int otiles = (d->Cout + 127) / 128;
  const int base = otiles * p.ctiles * d->ntaps;
  int splits = (2 * 148 + base - 1) / base;
  if (splits > p.tiles_total) splits = p.tiles_total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles_total + splits - 1) / splits;
  splits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  p.w_ld = d->w_ld; p.scale = scale;
  p.sems = nullptr;
  if (lcgan_det_enabled()) {
    LCGAN_CHECK(base * 4 <= kDetSems, "tapconv_wgrad_tc: too many tiles for deterministic mode");
    LCGAN_CUDA(cudaGetSymbolAddress((void**)&p.sems, g_sems));
  }

  // 3x3 stride-1, 32 / 64 channels on both sides: transposed, haloed kernel
  if (d->ntaps == 9 && d->is == 1 && d->os == 1 && (d->Cin == 32 || d->Cin == 64) && (d->Cout == 32 || d->Cout == 64) &&
      d->MW % 8 == 0 && d->MH % 16 == 0 && getenv("LCGAN_NO_WGRAD_HALO") == nullptr) {
    WhParams h{};
    int seen = 0;
    for (int t = 0; t < 9; ++t) {
      const int dy = d->dy[t], dx = d->dx[t];
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1) { seen = -1; break; }
      h.grp_wtap[dx + 1][dy + 1] = d->wtap[t];
      seen |= 1 << ((dy + 1) * 3 + dx + 1);
    }
    if (seen == 0x1FF) {
      h.N = d->N; h.Cin = d->Cin; h.Cout = d->Cout;
      h.tiles_w = d->MW / 8; h.tiles_h = d->MH / 16;
      for (h.lw = 0; (1 << h.lw) < h.tiles_w; ++h.lw) {}
      for (h.lh = 0; (1 << h.lh) < h.tiles_h; ++h.lh) {}
      h.tiles_total = h.tiles_w * h.tiles_h * d->N;
      int hs = sm_count();
      if (hs > h.tiles_total) hs = h.tiles_total;
      h.tiles_per_split = (h.tiles_total + hs - 1) / hs;
      hs = (h.tiles_total + h.tiles_per_split - 1) / h.tiles_per_split;
      h.kcx = d->Cin; h.kcg = d->Cout;
      h.dyn = getenv("LCGAN_WG_NO_DYN") == nullptr;
      h.x_bytes = ((h.dyn ? 10 * 16 : 10 * 18) * h.kcx * 2 + 1023) & ~1023;
      h.g_bytes = ((h.dyn ? 8 * 18 : 8 * 16) * h.kcg * 2 + 1023) & ~1023;
      h.stages = (kMaxSmem - 2048) / (h.x_bytes + h.g_bytes);
      if (h.stages > kWhStagesMax) h.stages = kWhStagesMax;
      h.w_ld = d->w_ld; h.scale = scale; h.sems = nullptr;
      if (lcgan_det_enabled()) LCGAN_CUDA(cudaGetSymbolAddress((void**)&h.sems, g_sems));
      const int smem = 1024 + h.stages * (h.x_bytes + h.g_bytes) + 256;
      CUtensorMap tmgh, tmxh;
      if (int e = make_act_map(&tmgh, g, d->N, d->OH, d->OW, d->Cout, 8, 16, 1, 1, h.kcg, h.dyn ? 2 : 0, 0)) return e;
      if (int e = make_act_map(&tmxh, x, d->N, d->IH, d->IW, d->Cin, 8, 16, 1, 1, h.kcx, h.dyn ? 0 : 2, 2)) return e;
      static DeviceOnce once3;
      const cudaError_t attr_err3 = (cudaError_t)once3.run([] {
        cudaError_t attr_err3 = cudaSuccess;
        attr_err3 = cudaFuncSetAttribute(tapconv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        return (int)attr_err3;
      });
      LCGAN_CHECK(attr_err3 == cudaSuccess, "tapconv_wgrad_halo: cannot raise dynamic shared memory: %s",
                  cudaGetErrorString(attr_err3));
      tapconv_wgrad_halo_kernel<<<hs, kThreads, smem, (cudaStream_t)stream>>>(tmgh, tmxh, h, dw2);
      LCGAN_LAUNCH_CHECK();
      return 0;
    }
  }

  if (d->ntaps * d->Cin <= kTnMaxN && d->Cin % 16 == 0 && getenv("LCGAN_NO_WGRAD_TN") == nullptr) {
    // small-channel layers: all taps along N, persistent-ish K split (one CTA per SM)
    int tsplits = sm_count() / otiles;
    if (tsplits > p.tiles_total) tsplits = p.tiles_total;
    if (tsplits < 1) tsplits = 1;
    p.tiles_per_split = (p.tiles_total + tsplits - 1) / tsplits;
    tsplits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
    CUtensorMap tmg2, tmx2;
    if (int e = make_act_map(&tmg2, g, d->N, d->OH, d->OW, d->Cout, p.wt, p.ht, p.nt, d->os, p.kcg)) return e;
    if (int e = make_act_map(&tmx2, x, d->N, d->IH, d->IW, d->Cin, p.wt, p.ht, p.nt, d->is, p.kcx)) return e;
    static DeviceOnce once2;
    const cudaError_t attr_err2 = (cudaError_t)once2.run([] {
      cudaError_t attr_err2 = cudaSuccess;
      attr_err2 = cudaFuncSetAttribute(tapconv_wgrad_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tn_smem_bytes());
      return (int)attr_err2;
    });
    LCGAN_CHECK(attr_err2 == cudaSuccess, "tapconv_wgrad_tn: cannot raise dynamic shared memory: %s",
                cudaGetErrorString(attr_err2));
    dim3 grid2(otiles, tsplits);
    tapconv_wgrad_tn_kernel<<<grid2, kThreads, tn_smem_bytes(), (cudaStream_t)stream>>>(tmg2, tmx2, p, dw2);
    LCGAN_LAUNCH_CHECK();
    return 0;
  }

  CUtensorMap tmg, tmx;
  if (int e = make_act_map(&tmg, g, d->N, d->OH, d->OW, d->Cout, p.wt, p.ht, p.nt, d->os, p.kcg)) return e;
  if (int e = make_act_map(&tmx, x, d->N, d->IH, d->IW, d->Cin, p.wt, p.ht, p.nt, d->is, p.kcx)) return e;
  static DeviceOnce once;
  const cudaError_t attr_err = (cudaError_t)once.run([] {
    cudaError_t attr_err = cudaSuccess;
    attr_err = cudaFuncSetAttribute(tapconv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes());
    return (int)attr_err;
  });
  LCGAN_CHECK(attr_err == cudaSuccess, "tapconv_wgrad_tc: cannot raise dynamic shared memory: %s",
              cudaGetErrorString(attr_err));
  dim3 grid(otiles * p.ctiles, d->ntaps, splits);
  tapconv_wgrad_tc_kernel<<<grid, kThreads, wg_smem_bytes(), (cudaStream_t)stream>>>(tmg, tmx, p, dw2);
  LCGAN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Micro-benchmarks of the raw tcgen05 / TMA issue rates (used to choose tile shapes).
// mode 0: `iters` back-to-back MMAs M=128 x N x K=16 (K-major SW128 operands in smem garbage)
// mode 1: same with MN-major operands
// mode 2: `iters` TMA loads of the activation box described by the tensor map into a 4-slot ring
// out[blockIdx.x] = elapsed SM clocks.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1)
tc_rate_kernel(const __grid_constant__ CUtensorMap tm, int mode, int n, int iters, int box_bytes, int kc,
               long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[5];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0) {
    const uint32_t leader = 1;
    if (mode < 2) {
      const bool mn = mode == 1;
      const uint32_t idesc = make_idesc(n, mn, mn);
      const uint64_t ad = make_desc(smem_u32(base), mn ? 16384 : 16, mn ? 1024 : 16 * kc, kc);
      const uint64_t bd = make_desc(smem_u32(base) + 32768, mn ? 16384 : 16, mn ? 1024 : 16 * kc, kc);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) umma_f16(tmem_base, ad, bd, idesc, 1, leader);
      umma_commit(&bar[4], leader);
      mbar_wait(&bar[4], 0);
      t1 = clock64();
    } else {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int st = i & 3;
        if (i >= 4) mbar_wait(&bar[st], ((i >> 2) - 1) & 1);
        mbar_expect_tx(&bar[st], box_bytes, leader);
        tma_load_4d(base + st * 32768, &tm, &bar[st], 0, (i * 16) & 1023, ((i >> 6) * 8) & 1023, blockIdx.x & 31, leader);
      }
      for (int st = 0; st < 4 && st < iters; ++st) {
        const int last = ((iters - 1 - st) >> 2);     // index of the last use of this slot
        if (iters > st) mbar_wait(&bar[st], last & 1);
      }
      t1 = clock64();
    }
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}
}  // namespace

extern "C" int lcgan_debug_tc_rate(int mode, int n, int iters, const void* act, int N, int H, int W, int C, int ht,
                                   long long* out_cycles, int blocks, void* stream) {
  CUtensorMap tm;
  const int kc = C % 64 == 0 ? 64 : 32;
  if (int e = make_act_map(&tm, act, N, H, W, C, 16, ht, 1, 1, kc)) return e;
  cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 2048);
  tc_rate_kernel<<<blocks, 128, 4 * 32768 + 2048, (cudaStream_t)stream>>>(tm, mode, n, iters, 16 * ht * kc * 2, kc,
                                                                       out_cycles);
  LCGAN_LAUNCH_CHECK();
  return 0;
}
