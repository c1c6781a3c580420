// conv2 (5x5 SAME, 32 -> 64 channels, 14x14 feature maps) as tcgen05 implicit GEMMs.
// reference op K5 (tf.nn.conv2d #2, src/mnist.py:119-122) = 82 % of the model FLOPs, fused with
// K2/K3/K4 (bias, ReLU, 2x2 max-pool, src/mnist.py:123-127) in the forward epilogue.
//
// No im2col buffer exists anywhere: activations stay NHWC bf16 and every filter tap is one 4-D TMA
// box (channels, 16 pixels of a row, 8 rows, 1 image) fetched at a shifted coordinate; the SAME
// padding halo and the 14 -> 16 row/column round-up are TMA out-of-bounds zero fill.  An M tile is
// therefore 128 pixels (8 x 16, 112 or 84 of them real) of one image; two tiles cover an image.
//
//   fwd    D[pixel, co]  += X[pixel + tap, ci] (K-major, 64B swizzle)  * W[tap][ci][co] (MN-major, 128B)
//   dgrad  D[pixel, ci]  += dY[pixel - tap, co] (K-major, 128B swizzle) * W[tap][ci][co] (K-major,  128B)
//   wgrad  D[(tap,ci), co] += X[pixel + tap, ci] (MN-major, 64B)        * dY[pixel, co]   (MN-major, 128B),
//          K = pixels, split over CTAs, fp32 atomics into the HWIO gradient tensor.
// The HWIO bf16 weight image (800 x 64) is loaded once per CTA and stays resident in shared memory
// (100 KB) while the persistent CTA walks its tiles; accumulators are double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "host_utils.h"

namespace dm {

constexpr int CV_THREADS = 192;
constexpr int W_BYTES = 800 * 64 * 2;   // 102400

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
constexpr int FW_STAGES = 4;
constexpr int FW_A_BYTES = 20 * 12 * 32 * 2;   // ONE patch per tile: 20 rows x 12 pixels x 32 ch = 15 KB, all 25 taps inside
struct FwSmem {
  static constexpr int A_OFF = W_BYTES;
  static constexpr int BAR_OFF = A_OFF + FW_STAGES * FW_A_BYTES;
  static constexpr int SCR_OFF = BAR_OFF + 512;                 // epilogue transpose scratch: 4 warps x 32 rows x 128 B
  static constexpr int BIAS_OFF = SCR_OFF + 4 * 4096;           // 64 floats
  static constexpr int TOTAL = BIAS_OFF + 256 + 1024;
};

// EPI = 1: the 2x2 max-pool is an in-thread max.  Each epilogue warp transposes its 32 pixels x 32 channels through a
//   swizzled 4 KB shared-memory scratch so that a lane ends up with ALL FOUR window positions of the 8 channels it stores
//   (8 STS.128 + 8 LDS.128 per chunk) -- ~200 instructions per 32-channel chunk.
// EPI = 0: first version, 2 warp shuffles + selects per value (~560 instructions per chunk; with one warp per scheduler
//   the epilogue, not the MMAs, set the kernel's pace: 1.8 us vs 1.25 us per tile).
template <int EPI>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv2_fwd_kernel(const __grid_constant__ CUtensorMap tmX,   // a1 NHWC [B,14,14,32], box (32,12,20,1), 64B swizzle
                 const __grid_constant__ CUtensorMap tmW,   // W [800][64], box (64,200), 128B swizzle
                 const float* __restrict__ bias,            // [64]
                 __nv_bfloat16* __restrict__ out,           // pooled [B,7,7,64]
                 uint8_t* __restrict__ code,                // [B,7,7,64]
                 int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + FwSmem::BAR_OFF);
  uint64_t* a_empty = a_full + FW_STAGES;
  uint64_t* acc_full = a_empty + FW_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w_full + 1);
  float* s_bias = reinterpret_cast<float*>(smem + FwSmem::BIAS_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < FW_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_holder);
  if (warp == 2) { s_bias[lane] = __ldg(bias + lane); s_bias[32 + lane] = __ldg(bias + 32 + lane); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();
  // (no early launch_dependents: resident-but-blocked CTAs of the next kernel steal SM resources from this one)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, W_BYTES);
      for (int i = 0; i < 4; ++i) tma_load_2d(smem + i * 25600, &tmW, w_full, 0, 200 * i);
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        // tile = 16 rows x 8 columns of one image (columns w0 .. w0+7); its receptive field is the 20 x 12 patch
        // starting at (-2, w0-2).  One TMA box; halo, rows 14-15 and columns 14-15 are OOB zero fill.
        const int img = t >> 1, w0 = (t & 1) * 8;
        const int s = it % FW_STAGES;
        mbar_wait(&a_empty[s], ((it / FW_STAGES) & 1) ^ 1);
        mbar_expect_tx(&a_full[s], FW_A_BYTES);
        tma_load_4d(smem + FwSmem::A_OFF + s * FW_A_BYTES, &tmX, &a_full[s], 0, w0 - 2, -2, img);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, /*A MN*/ false, /*B MN*/ true);
    mbar_wait(w_full, 0);
    int it = 0, tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      mbar_wait(&acc_empty[buf], ((tl >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      {
        const int s = it % FW_STAGES;
        mbar_wait(&a_full[s], (it / FW_STAGES) & 1);
        tc_fence_after_sync();
        if (elect_one()) {
          // output pixel (h, w) reads patch pixel (h + kh, w + kw): the window of tap (kh,kw) starts (kh*12 + kw)
          // rows into the patch; 8-row groups (fixed h) are one patch row = 12 pixels = 768 B apart.  A descriptor
          // may start at any row of the swizzled buffer (the hardware XORs absolute address bits:
          // profiles/umma_row_shift_probe_r1.txt).  Descriptors = one base + compile-time offsets (>> 4).
          const uint64_t da0 = make_smem_desc(smem_u32(smem + FwSmem::A_OFF + s * FW_A_BYTES), 16, 768, SWZ_64B);
          const uint64_t db0 = make_smem_desc(smem_u32(smem), 8192, 1024, SWZ_128B);
          const uint32_t tm = tmem_base + buf * 64;
#pragma unroll
          for (int tap = 0; tap < 25; ++tap) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t da = da0 + (uint64_t)((((tap / 5) * 12 + (tap % 5)) * 64 + 32 * k) >> 4);
              const uint64_t db = db0 + (uint64_t)((tap * 4096 + 2048 * k) >> 4);     // 16 ci rows per step
              umma_bf16(tm, da, db, idesc, (tap | k) != 0);
            }
          }
          umma_commit(&a_empty[s]);
          umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        ++it;
      }
    }
  } else {
    const int q = warp & 3;
    int tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      const int img = t >> 1, w0 = (t & 1) * 8;
      mbar_wait(&acc_full[buf], (tl >> 1) & 1);
      tc_fence_after_sync();
      // accumulator row = h*8 + w: this warp holds rows h = 4q .. 4q+3; lane = (h & 3) * 8 + w
      const int w = w0 + (lane & 7), h = 4 * q + (lane >> 3);
      const int ph = h >> 1, pw = w >> 1;
      const bool valid = (h < 14) && (w < 14);
      const int qp = (((lane >> 3) & 1) << 1) | (lane & 1);   // position inside the 2x2 pooling window
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64 + c * 32, v);
        tmem_ld_wait();
        if (EPI == 1) {
          uint8_t* scr = smem + FwSmem::SCR_OFF + q * 4096;
          __syncwarp();                                   // previous chunk's reads are done
#pragma unroll
          for (int k = 0; k < 8; ++k)                     // row = lane (pixel), 16-byte chunk k at k ^ (lane & 7)
            *reinterpret_cast<uint4*>(scr + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          __syncwarp();
          const int l0 = lane & ~9;                       // lane of window position 0 (dy = 0, dx = 0)
          float val[4][8];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int rp = l0 + (p & 1) + 8 * (p >> 1);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float4 t = *reinterpret_cast<const float4*>(scr + rp * 128 + (((2 * qp + e) ^ (rp & 7)) << 4));
              val[p][4 * e] = t.x; val[p][4 * e + 1] = t.y; val[p][4 * e + 2] = t.z; val[p][4 * e + 3] = t.w;
            }
          }
          if (valid) {
            uint32_t pk[4];
            uint32_t cd[2] = {0u, 0u};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              float m = val[0][jj];
              uint32_t idx = 0;
#pragma unroll
              for (int p = 1; p < 4; ++p)
                if (val[p][jj] > m) { m = val[p][jj]; idx = p; }
              m += s_bias[c * 32 + qp * 8 + jj];
              const bool active = m > 0.f;
              cd[jj >> 2] |= (idx | (active ? 4u : 0u)) << ((jj & 3) * 8);
              const float o = active ? m : 0.f;
              if (jj & 1) pk[jj >> 1] = pack_bf16x2(__uint_as_float(pk[jj >> 1]), o);
              else pk[jj >> 1] = __float_as_uint(o);
            }
            const size_t off = (((size_t)img * 7 + ph) * 7 + pw) * 64 + c * 32 + qp * 8;
            *reinterpret_cast<uint4*>(out + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint2*>(code + off) = make_uint2(cd[0], cd[1]);
          }
          continue;
        }
        uint32_t bits0 = 0, bits1 = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float a = __uint_as_float(v[j]);
          const float b = __shfl_xor_sync(0xffffffffu, a, 1);
          const float odd = (lane & 1) ? a : b, even = (lane & 1) ? b : a;
          bits0 |= (odd > even ? 1u : 0u) << j;           // argmax column inside my row pair
          const float m1 = fmaxf(a, b);
          const float o = __shfl_xor_sync(0xffffffffu, m1, 8);
          const float bot = (lane & 8) ? m1 : o, top = (lane & 8) ? o : m1;
          bits1 |= (bot > top ? 1u : 0u) << j;            // argmax row
          v[j] = __float_as_uint(fmaxf(m1, o));
        }
        const uint32_t other = __shfl_xor_sync(0xffffffffu, bits0, 8);
        const uint32_t bits_top = (lane & 8) ? other : bits0, bits_bot = (lane & 8) ? bits0 : other;
        if (valid) {
          // every lane of the 2x2 window now holds the pooled values of all 32 channels; each writes 8
          uint32_t pk[4];
          uint32_t cd[2] = {0u, 0u};
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            // select channel (qp*8 + jj) with compile-time register indices
            float m = 0.f;
            uint32_t i0t = 0, i0b = 0, i1 = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s)
              if (qp == s) {
                m = __uint_as_float(v[s * 8 + jj]);
                i0t = (bits_top >> (s * 8 + jj)) & 1u;
                i0b = (bits_bot >> (s * 8 + jj)) & 1u;
                i1 = (bits1 >> (s * 8 + jj)) & 1u;
              }
            m += bias[c * 32 + qp * 8 + jj];
            const bool active = m > 0.f;
            const uint32_t idx = i1 ? (2u | i0b) : i0t;
            cd[jj >> 2] |= (idx | (active ? 4u : 0u)) << ((jj & 3) * 8);
            const float o = active ? m : 0.f;
            if (jj & 1) pk[jj >> 1] = pack_bf16x2(__uint_as_float(pk[jj >> 1]), o);
            else pk[jj >> 1] = __float_as_uint(o);
          }
          const size_t off = (((size_t)img * 7 + ph) * 7 + pw) * 64 + c * 32 + qp * 8;
          *reinterpret_cast<uint4*>(out + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint2*>(code + off) = make_uint2(cd[0], cd[1]);
        }
      }
      tc_fence_before_sync();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------
// dgrad
// ------------------------------------------------------------------------------------------------------
constexpr int DG_STAGES = 3;
constexpr int DG_A_BYTES = 20 * 12 * 64 * 2;   // ONE patch per tile: 20 rows x 12 pixels x 64 ch = 30 KB
struct DgSmem {
  static constexpr int A_OFF = W_BYTES;
  static constexpr int BAR_OFF = A_OFF + DG_STAGES * DG_A_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

// Register cap: the register file is per scheduler (4 x 16 K).  At 230 registers/thread the two schedulers that hold two
// of this CTA's six warps had 1.5 K registers left, so no other kernel's CTA (which needs a warp on every scheduler) could
// share the SM: fc2_wgrad and the early aggregation kernel only started as these CTAs exited (profiles/bench_r1_call30_1gpu.txt,
// profiles/coresidency_probe_r1.txt).  128 registers leave 8 K per scheduler.
__global__ void __maxnreg__(128)
conv2_dgrad_kernel(const __grid_constant__ CUtensorMap tmDY,  // dY NHWC [B,14,14,64], box (64,12,20,1), 128B swizzle
                   const __grid_constant__ CUtensorMap tmW,
                   __nv_bfloat16* __restrict__ dx,            // [B,14,14,32]
                   int num_tiles, unsigned long long* __restrict__ dbg) {   // dbg: optional timeline of CTA 0
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + DgSmem::BAR_OFF);
  uint64_t* a_empty = a_full + DG_STAGES;
  uint64_t* acc_full = a_empty + DG_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < DG_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  const bool dbg_on = dbg != nullptr && blockIdx.x == 0;
  if (dbg_on && threadIdx.x == 0) dbg[0] = globaltimer_ns();
  if (warp == 1) tmem_alloc<64>(tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();
  // (no early launch_dependents: resident-but-blocked CTAs of the next kernel steal SM resources from this one)
  if (dbg_on && threadIdx.x == 0) dbg[1] = globaltimer_ns();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, W_BYTES);
      for (int i = 0; i < 4; ++i) tma_load_2d(smem + i * 25600, &tmW, w_full, 0, 200 * i);
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int img = t >> 1, w0 = (t & 1) * 8;
        const int s = it % DG_STAGES;
        mbar_wait(&a_empty[s], ((it / DG_STAGES) & 1) ^ 1);
        if (dbg_on && it < 20) dbg[8 + it] = globaltimer_ns();          // producer: slot free, issuing load `it`
        mbar_expect_tx(&a_full[s], DG_A_BYTES);
        // dX[y,x] += dY[y + 2 - kh, x + 2 - kw] * W[kh,kw]: same 20 x 12 patch at (-2, w0-2); tap (kh,kw)
        // starts (4-kh) rows and (4-kw) pixels into it
        tma_load_4d(smem + DgSmem::A_OFF + s * DG_A_BYTES, &tmDY, &a_full[s], 0, w0 - 2, -2, img);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
    mbar_wait(w_full, 0);
    if (dbg_on && lane == 0) dbg[2] = globaltimer_ns();                       // weights resident
    int it = 0, tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      mbar_wait(&acc_empty[buf], ((tl >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      {
        const int s = it % DG_STAGES;
        mbar_wait(&a_full[s], (it / DG_STAGES) & 1);
        tc_fence_after_sync();
        if (dbg_on && lane == 0 && it < 20) dbg[32 + it] = globaltimer_ns();   // MMA: data of load `it` landed
        if (elect_one()) {
          // tap (kh,kw) starts (4-kh) patch rows and (4-kw) pixels into the patch; 128 B per pixel, groups 12*128 B apart
          const uint64_t da0 = make_smem_desc(smem_u32(smem + DgSmem::A_OFF + s * DG_A_BYTES), 16, 1536, SWZ_128B);
          const uint64_t db0 = make_smem_desc(smem_u32(smem), 16, 1024, SWZ_128B);   // W[tap]: 32 ci rows x 64 co, K-major
          const uint32_t tm = tmem_base + buf * 32;
#pragma unroll
          for (int tap = 0; tap < 25; ++tap) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = da0 + (uint64_t)((((4 - tap / 5) * 12 + (4 - tap % 5)) * 128 + 32 * k) >> 4);
              const uint64_t db = db0 + (uint64_t)((tap * 4096 + 32 * k) >> 4);
              umma_bf16(tm, da, db, idesc, (tap | k) != 0);
            }
          }
          umma_commit(&a_empty[s]);
          umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        ++it;
      }
    }
  } else {
    const int q = warp & 3;
    int tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      const int img = t >> 1, w0 = (t & 1) * 8;
      mbar_wait(&acc_full[buf], (tl >> 1) & 1);
      tc_fence_after_sync();
      if (dbg_on && q == 0 && lane == 0 && tl < 8) dbg[56 + tl] = globaltimer_ns();   // epilogue: accumulator of tile tl complete
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 32, v);
      tmem_ld_wait();
      tc_fence_before_sync();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      const int h = 4 * q + (lane >> 3), w = w0 + (lane & 7);
      if (h < 14 && w < 14) {
        __nv_bfloat16* o = dx + (((size_t)img * 14 + h) * 14 + w) * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(o + j) =
              make_uint4(pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                         pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                         pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])),
                         pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (dbg_on && threadIdx.x == 0) dbg[3] = globaltimer_ns();
  if (warp == 1) tmem_dealloc<64>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------
// wgrad: CTA = (group of 4 taps, slice of the pixel tiles).  Per pixel tile (16 rows x 8 columns) a stage
// holds the same 20 x 12 input patch as the forward pass (15 KB, A operand for every tap of the group) and
// the 16 x 8 tile of dY (16 KB, B operand).  M = (tap, ci): the four taps of a group are M-chunks whose
// windows start a uniform distance apart inside the patch:
//     group g < 5 : taps (kh = 0..3, kw = g)   -> chunk stride = one patch row = 768 B
//     group 5     : taps (kh = 4, kw = 0..3)   -> chunk stride = one pixel     =  64 B
//     group 6     : tap  (kh = 4, kw = 4) + three padding chunks (computed, never stored)
// ------------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 4;   // 4 x 32 KB: leaves room for two conv1_wgrad CTAs (37 KB each) on the same SM
constexpr int WG_A_BYTES = 20 * 12 * 32 * 2;   // 15 KB patch
constexpr int WG_B_BYTES = 128 * 64 * 2;       // 16 KB dY tile
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES + 1024;   // keep the dY tile 1024-aligned (128B swizzle)
constexpr int WG_GROUPS = 7;
struct WgSmem {
  static constexpr int BAR_OFF = WG_STAGES * WG_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024 + 4096;     // + slack: group 6's padding chunks read past the patch
};

// GPC = tap groups per CTA.  Every group of a CTA reuses the stage's patch + dY tile, so the shared-memory fill traffic
// (the bound of the first version: 7 groups x 21 pixel slices re-read X and dY seven times = 111 MB through L2 per step at
// batch 256) drops to ceil(7 / GPC) passes, at the price of ceil(7 / GPC) x fewer pixel slices, i.e. more atomics per output.
template <int GPC>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv2_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,    // a1, box (32,12,20,1), 64B swizzle
                   const __grid_constant__ CUtensorMap tmDY,   // dY, box (64,8,16,1), 128B swizzle
                   float* __restrict__ g_w,                    // [25][32][64] fp32, accumulated atomically
                   int num_tiles, int splits) {
  constexpr int SETS = (WG_GROUPS + GPC - 1) / GPC;
  constexpr uint32_t TM_COLS = GPC * 64 <= 64 ? 64 : (GPC * 64 <= 128 ? 128 : (GPC * 64 <= 256 ? 256 : 512));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WgSmem::BAR_OFF);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* acc_full = empty + WG_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int set = blockIdx.x % SETS, split = blockIdx.x / SETS;
  const int g_first = set * GPC;
  const int ng = (g_first + GPC <= WG_GROUPS) ? GPC : (WG_GROUPS - g_first);   // groups this CTA really owns
  const int t_begin = (int)(((long long)num_tiles * split) / splits);
  const int t_end = (int)(((long long)num_tiles * (split + 1)) / splits);
  const int nt = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TM_COLS>(tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();
  // (no early launch_dependents: resident-but-blocked CTAs of the next kernel steal SM resources from this one)

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int t = t_begin + i, img = t >> 1, w0 = (t & 1) * 8;
        const int s = i % WG_STAGES;
        mbar_wait(&empty[s], ((i / WG_STAGES) & 1) ^ 1);
        uint8_t* sA = smem + s * WG_STAGE_BYTES;
        mbar_expect_tx(&full[s], WG_A_BYTES + WG_B_BYTES);
        tma_load_4d(sA, &tmX, &full[s], 0, w0 - 2, -2, img);
        tma_load_4d(sA + WG_A_BYTES + 1024, &tmDY, &full[s], 0, w0, 0, img);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
    for (int i = 0; i < nt; ++i) {
      const int s = i % WG_STAGES;
      mbar_wait(&full[s], (i / WG_STAGES) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t b_addr = smem_u32(smem + s * WG_STAGE_BYTES) + WG_A_BYTES + 1024;
        // B (MN-major, 128B swizzle): [pixel][64 co], 8-pixel groups 1024 B apart; K step = 2048 B.
        const uint64_t db0 = make_smem_desc(b_addr, 8192, 1024, SWZ_128B);
#pragma unroll
        for (int gi = 0; gi < GPC; ++gi) {
          if (gi < ng) {
            const int group = g_first + gi;
            // first tap of the group (as a pixel offset into the patch) and the byte distance between its four M-chunks
            const int base_px = group < 5 ? group : (group == 5 ? 4 * 12 : 4 * 12 + 4);
            const uint32_t lbo = group < 5 ? 768u : 64u;
            // A (MN-major, 64B swizzle): k = pixel; 8-pixel groups (one tile row) are one patch row = 768 B apart,
            //   M-chunks (taps) `lbo` apart; a K step = 16 pixels = two tile rows = 1536 B.
            const uint64_t da0 = make_smem_desc(smem_u32(smem + s * WG_STAGE_BYTES) + base_px * 64, lbo, 768, SWZ_64B);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem_base + gi * 64, da0 + (uint64_t)((1536 * k) >> 4), db0 + (uint64_t)((2048 * k) >> 4), idesc, (i | k) != 0);
          }
        }
        umma_commit(&empty[s]);
        if (i == nt - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (nt > 0) {
    const int q = warp & 3;               // rows 32q .. 32q+31 = M-chunk q of the group, row = ci
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int gi = 0; gi < ng; ++gi) {
      const int group = g_first + gi;
      const int tap = group < 5 ? q * 5 + group : (group == 5 ? 20 + q : (q == 0 ? 24 : 25));
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + gi * 64 + c * 32, v);
        tmem_ld_wait();
        if (tap < 25) {
          float* o = g_w + ((size_t)tap * 32 + lane) * 64 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_f32x4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TM_COLS>(tmem_base);
}

}  // namespace dm

extern "C" {

// a1: [B,14,14,32] bf16;  w_bf16: HWIO [800][64] bf16;  out/code: [B,7,7,64]
int dm_conv2_fwd(const void* a1, const void* w_bf16, const void* bias, void* out, void* code, int B, void* stream) {
  using namespace dm;
  CUtensorMap tmX, tmW;
  if (make_tmap_nhwc_bf16(&tmX, a1, 32, 14, 14, B, 32, 12, 20, 64)) return 101;
  if (make_tmap_2d_bf16(&tmW, w_bf16, 64, 800, 64, 64, 200, 128)) return 102;
  static const int epi = env_int("DMNIST_CONV2_EPI", 1);
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwSmem::TOTAL));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_fwd_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwSmem::TOTAL));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_fwd_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const int tiles = 2 * B;
  const int grid = tiles < 148 ? tiles : 148;
  auto kern = epi == 0 ? conv2_fwd_kernel<0> : conv2_fwd_kernel<1>;
  return (int)launch_kernel(kern, dim3(grid), dim3(CV_THREADS), FwSmem::TOTAL,
                            reinterpret_cast<cudaStream_t>(stream), tmX, tmW, reinterpret_cast<const float*>(bias),
                            reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<uint8_t*>(code), tiles);
}

// dy: [B,14,14,64] bf16 -> dx: [B,14,14,32] bf16
int dm_conv2_dgrad_dbg(const void* dy, const void* w_bf16, void* dx, int B, void* dbg, void* stream);
int dm_conv2_dgrad(const void* dy, const void* w_bf16, void* dx, int B, void* stream) {
  return dm_conv2_dgrad_dbg(dy, w_bf16, dx, B, nullptr, stream);
}
int dm_conv2_dgrad_dbg(const void* dy, const void* w_bf16, void* dx, int B, void* dbg, void* stream) {
  using namespace dm;
  CUtensorMap tmDY, tmW;
  if (make_tmap_nhwc_bf16(&tmDY, dy, 64, 14, 14, B, 64, 12, 20, 128)) return 101;
  if (make_tmap_2d_bf16(&tmW, w_bf16, 64, 800, 64, 64, 200, 128)) return 102;
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DgSmem::TOTAL));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_dgrad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const int tiles = 2 * B;
  const int grid = tiles < g_max_ctas ? tiles : g_max_ctas;
  return (int)launch_kernel(conv2_dgrad_kernel, dim3(grid), dim3(CV_THREADS), DgSmem::TOTAL,
                            reinterpret_cast<cudaStream_t>(stream), tmDY, tmW, reinterpret_cast<__nv_bfloat16*>(dx), tiles,
                            reinterpret_cast<unsigned long long*>(dbg));
}

// g_w ([25][32][64] fp32) must be zeroed by the caller; accumulated with atomics.
int dm_conv2_wgrad(const void* a1, const void* dy, void* g_w, int B, void* stream) {
  using namespace dm;
  CUtensorMap tmX, tmDY;
  if (make_tmap_nhwc_bf16(&tmX, a1, 32, 14, 14, B, 32, 12, 20, 64)) return 101;
  if (make_tmap_nhwc_bf16(&tmDY, dy, 64, 14, 14, B, 64, 8, 16, 128)) return 102;
  const int tiles = 2 * B;
  static const int gpc = env_int("DMNIST_WGRAD_GPC", 1);      // tap groups per CTA: 1 (first version), 2, 4
  const int sets = gpc == 1 ? 7 : (gpc == 2 ? 4 : 2);
  int splits = g_max_ctas / sets;
  if (splits > tiles) splits = tiles;
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem::TOTAL));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem::TOTAL));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem::TOTAL));
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* gw = reinterpret_cast<float*>(g_w);
  if (gpc == 1) return (int)launch_kernel(conv2_wgrad_kernel<1>, dim3(sets * splits), dim3(CV_THREADS), WgSmem::TOTAL, st, tmX, tmDY, gw, tiles, splits);
  if (gpc == 2) return (int)launch_kernel(conv2_wgrad_kernel<2>, dim3(sets * splits), dim3(CV_THREADS), WgSmem::TOTAL, st, tmX, tmDY, gw, tiles, splits);
  return (int)launch_kernel(conv2_wgrad_kernel<4>, dim3(sets * splits), dim3(CV_THREADS), WgSmem::TOTAL, st, tmX, tmDY, gw, tiles, splits);
}

}  // extern "C"
