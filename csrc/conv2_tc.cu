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
constexpr int FW_STAGES = 6;
constexpr int FW_A_BYTES = 12 * 16 * 32 * 2;   // one kw-patch: 12 rows x 16 pixels x 32 ch = 12 KB (5 kh taps inside)
struct FwSmem {
  static constexpr int A_OFF = W_BYTES;
  static constexpr int BAR_OFF = A_OFF + FW_STAGES * FW_A_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

__global__ void __launch_bounds__(CV_THREADS, 1)
conv2_fwd_kernel(const __grid_constant__ CUtensorMap tmX,   // a1 NHWC [B,14,14,32], box (32,16,12,1), 64B swizzle
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
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, W_BYTES);
      for (int i = 0; i < 4; ++i) tma_load_2d(smem + i * 25600, &tmW, w_full, 0, 200 * i);
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int img = t >> 1, h0 = (t & 1) * 8;
        for (int kw = 0; kw < 5; ++kw, ++it) {
          // one patch per horizontal tap offset: rows h0-2 .. h0+9; the five vertical taps are row offsets
          // inside it (16 pixels = 1024 B, a multiple of the 512 B swizzle period)
          const int s = it % FW_STAGES;
          mbar_wait(&a_empty[s], ((it / FW_STAGES) & 1) ^ 1);
          mbar_expect_tx(&a_full[s], FW_A_BYTES);
          tma_load_4d(smem + FwSmem::A_OFF + s * FW_A_BYTES, &tmX, &a_full[s], 0, kw - 2, h0 - 2, img);
        }
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
      for (int kw = 0; kw < 5; ++kw, ++it) {
        const int s = it % FW_STAGES;
        mbar_wait(&a_full[s], (it / FW_STAGES) & 1);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + FwSmem::A_OFF + s * FW_A_BYTES);
#pragma unroll
          for (int kh = 0; kh < 5; ++kh) {
            const uint32_t b_addr = smem_u32(smem + (kh * 5 + kw) * 4096);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t da = make_smem_desc(a_addr + kh * 1024 + 32 * k, 16, 512, SWZ_64B);  // 64 B rows, 8-row groups 512 B
              const uint64_t db = make_smem_desc(b_addr + 2048 * k, 8192, 1024, SWZ_128B);      // 16 ci rows per step
              umma_bf16(tmem_base + buf * 64, da, db, idesc, (kw | kh | k) != 0);
            }
          }
          umma_commit(&a_empty[s]);
          if (kw == 4) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    int tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      const int img = t >> 1, h0 = (t & 1) * 8;
      mbar_wait(&acc_full[buf], (tl >> 1) & 1);
      tc_fence_after_sync();
      const int w = lane & 15;
      const int ph = (h0 >> 1) + q, pw = w >> 1;
      const bool valid = (h0 + 2 * q < 14) && (w < 14);
      const int qp = ((lane >> 4) << 1) | (lane & 1);   // position inside the 2x2 pooling window
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64 + c * 32, v);
        tmem_ld_wait();
        uint32_t bits0 = 0, bits1 = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float a = __uint_as_float(v[j]);
          const float b = __shfl_xor_sync(0xffffffffu, a, 1);
          const float odd = (lane & 1) ? a : b, even = (lane & 1) ? b : a;
          bits0 |= (odd > even ? 1u : 0u) << j;           // argmax column inside my row pair
          const float m1 = fmaxf(a, b);
          const float o = __shfl_xor_sync(0xffffffffu, m1, 16);
          const float bot = (lane & 16) ? m1 : o, top = (lane & 16) ? o : m1;
          bits1 |= (bot > top ? 1u : 0u) << j;            // argmax row
          v[j] = __float_as_uint(fmaxf(m1, o));
        }
        const uint32_t other = __shfl_xor_sync(0xffffffffu, bits0, 16);
        const uint32_t bits_top = (lane & 16) ? other : bits0, bits_bot = (lane & 16) ? bits0 : other;
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
constexpr int DG_STAGES = 4;
constexpr int DG_A_BYTES = 12 * 16 * 64 * 2;   // one kw-patch: 12 rows x 16 pixels x 64 ch = 24 KB
struct DgSmem {
  static constexpr int A_OFF = W_BYTES;
  static constexpr int BAR_OFF = A_OFF + DG_STAGES * DG_A_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

__global__ void __launch_bounds__(CV_THREADS, 1)
conv2_dgrad_kernel(const __grid_constant__ CUtensorMap tmDY,  // dY NHWC [B,14,14,64], box (64,16,12,1), 128B swizzle
                   const __grid_constant__ CUtensorMap tmW,
                   __nv_bfloat16* __restrict__ dx,            // [B,14,14,32]
                   int num_tiles) {
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
  if (warp == 1) tmem_alloc<64>(tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, W_BYTES);
      for (int i = 0; i < 4; ++i) tma_load_2d(smem + i * 25600, &tmW, w_full, 0, 200 * i);
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int img = t >> 1, h0 = (t & 1) * 8;
        for (int kw = 0; kw < 5; ++kw, ++it) {
          const int s = it % DG_STAGES;
          mbar_wait(&a_empty[s], ((it / DG_STAGES) & 1) ^ 1);
          mbar_expect_tx(&a_full[s], DG_A_BYTES);
          // dX[y,x] += dY[y - (kh-2), x - (kw-2)] * W[kh,kw]: patch rows h0-2 .. h0+9 at column offset 2-kw;
          // tap kh starts at patch row 4-kh
          tma_load_4d(smem + DgSmem::A_OFF + s * DG_A_BYTES, &tmDY, &a_full[s], 0, 2 - kw, h0 - 2, img);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
    mbar_wait(w_full, 0);
    int it = 0, tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      mbar_wait(&acc_empty[buf], ((tl >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      for (int kw = 0; kw < 5; ++kw, ++it) {
        const int s = it % DG_STAGES;
        mbar_wait(&a_full[s], (it / DG_STAGES) & 1);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + DgSmem::A_OFF + s * DG_A_BYTES);
#pragma unroll
          for (int kh = 0; kh < 5; ++kh) {
            const uint32_t b_addr = smem_u32(smem + (kh * 5 + kw) * 4096);   // W[tap]: 32 ci rows x 64 co (128 B), K-major
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc(a_addr + (4 - kh) * 2048 + 32 * k, 16, 1024, SWZ_128B);
              const uint64_t db = make_smem_desc(b_addr + 32 * k, 16, 1024, SWZ_128B);
              umma_bf16(tmem_base + buf * 32, da, db, idesc, (kw | kh | k) != 0);
            }
          }
          umma_commit(&a_empty[s]);
          if (kw == 4) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    int tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      const int img = t >> 1, h0 = (t & 1) * 8;
      mbar_wait(&acc_full[buf], (tl >> 1) & 1);
      tc_fence_after_sync();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 32, v);
      tmem_ld_wait();
      tc_fence_before_sync();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      const int h = h0 + 2 * q + (lane >> 4), w = lane & 15;
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
  if (warp == 1) tmem_dealloc<64>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------
// wgrad: CTA = (group of 4 taps, slice of the pixel tiles)
// ------------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = 4 * 128 * 32 * 2;   // 4 taps x 8 KB
constexpr int WG_B_BYTES = 128 * 64 * 2;       // 16 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_GROUPS = 7;                   // ceil(25 / 4)
struct WgSmem {
  static constexpr int BAR_OFF = WG_STAGES * WG_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

__global__ void __launch_bounds__(CV_THREADS, 1)
conv2_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,    // a1, box (32,16,8,1), 64B swizzle
                   const __grid_constant__ CUtensorMap tmDY,   // dY, box (64,16,8,1), 128B swizzle
                   float* __restrict__ g_w,                    // [25][32][64] fp32, accumulated atomically
                   int num_tiles, int splits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WgSmem::BAR_OFF);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* acc_full = empty + WG_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x % WG_GROUPS, split = blockIdx.x / WG_GROUPS;
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
  if (warp == 1) tmem_alloc<64>(tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int t = t_begin + i, img = t >> 1, h0 = (t & 1) * 8;
        const int s = i % WG_STAGES;
        mbar_wait(&empty[s], ((i / WG_STAGES) & 1) ^ 1);
        uint8_t* sA = smem + s * WG_STAGE_BYTES;
        mbar_expect_tx(&full[s], WG_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tap = min(group * 4 + j, 24);   // the last group is padded with copies of tap 24
          tma_load_4d(sA + j * 8192, &tmX, &full[s], 0, tap % 5 - 2, h0 + tap / 5 - 2, img);
        }
        tma_load_4d(sA + WG_A_BYTES, &tmDY, &full[s], 0, 0, h0, img);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
    for (int i = 0; i < nt; ++i) {
      const int s = i % WG_STAGES;
      mbar_wait(&full[s], (i / WG_STAGES) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * WG_STAGE_BYTES);
        const uint32_t b_addr = a_addr + WG_A_BYTES;
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 16 pixels per step
          // A: [pixel][32 ci] per tap, 64 B rows; 8-pixel groups 512 B apart; the 4 taps (M chunks) 8192 B apart
          const uint64_t da = make_smem_desc(a_addr + 1024 * k, 8192, 512, SWZ_64B);
          // B: [pixel][64 co], 128 B rows; 8-pixel groups 1024 B apart
          const uint64_t db = make_smem_desc(b_addr + 2048 * k, 8192, 1024, SWZ_128B);
          umma_bf16(tmem_base, da, db, idesc, (i | k) != 0);
        }
        umma_commit(&empty[s]);
        if (i == nt - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (nt > 0) {
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const int tap = group * 4 + q;        // rows 32q .. 32q+31 belong to tap q of the group, row = ci
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (tap < 25) {
        float* o = g_w + ((size_t)tap * 32 + lane) * 64 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(o + j, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem_base);
}

}  // namespace dm

extern "C" {

// a1: [B,14,14,32] bf16;  w_bf16: HWIO [800][64] bf16;  out/code: [B,7,7,64]
int dm_conv2_fwd(const void* a1, const void* w_bf16, const void* bias, void* out, void* code, int B, void* stream) {
  using namespace dm;
  CUtensorMap tmX, tmW;
  if (make_tmap_nhwc_bf16(&tmX, a1, 32, 14, 14, B, 32, 16, 12, 64)) return 101;
  if (make_tmap_2d_bf16(&tmW, w_bf16, 64, 800, 64, 64, 200, 128)) return 102;
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwSmem::TOTAL));
    configured = true;
  }
  const int tiles = 2 * B;
  const int grid = tiles < 148 ? tiles : 148;
  conv2_fwd_kernel<<<grid, CV_THREADS, FwSmem::TOTAL, reinterpret_cast<cudaStream_t>(stream)>>>(
      tmX, tmW, reinterpret_cast<const float*>(bias), reinterpret_cast<__nv_bfloat16*>(out),
      reinterpret_cast<uint8_t*>(code), tiles);
  return (int)cudaGetLastError();
}

// dy: [B,14,14,64] bf16 -> dx: [B,14,14,32] bf16
int dm_conv2_dgrad(const void* dy, const void* w_bf16, void* dx, int B, void* stream) {
  using namespace dm;
  CUtensorMap tmDY, tmW;
  if (make_tmap_nhwc_bf16(&tmDY, dy, 64, 14, 14, B, 64, 16, 12, 128)) return 101;
  if (make_tmap_2d_bf16(&tmW, w_bf16, 64, 800, 64, 64, 200, 128)) return 102;
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DgSmem::TOTAL));
    configured = true;
  }
  const int tiles = 2 * B;
  const int grid = tiles < 148 ? tiles : 148;
  conv2_dgrad_kernel<<<grid, CV_THREADS, DgSmem::TOTAL, reinterpret_cast<cudaStream_t>(stream)>>>(
      tmDY, tmW, reinterpret_cast<__nv_bfloat16*>(dx), tiles);
  return (int)cudaGetLastError();
}

// g_w ([25][32][64] fp32) must be zeroed by the caller; accumulated with atomics.
int dm_conv2_wgrad(const void* a1, const void* dy, void* g_w, int B, void* stream) {
  using namespace dm;
  CUtensorMap tmX, tmDY;
  if (make_tmap_nhwc_bf16(&tmX, a1, 32, 14, 14, B, 32, 16, 8, 64)) return 101;
  if (make_tmap_nhwc_bf16(&tmDY, dy, 64, 14, 14, B, 64, 16, 8, 128)) return 102;
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(conv2_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem::TOTAL));
    configured = true;
  }
  const int tiles = 2 * B;
  int splits = 148 / WG_GROUPS;   // 21 -> 147 CTAs
  if (splits > tiles) splits = tiles;
  conv2_wgrad_kernel<<<WG_GROUPS * splits, CV_THREADS, WgSmem::TOTAL, reinterpret_cast<cudaStream_t>(stream)>>>(
      tmX, tmDY, reinterpret_cast<float*>(g_w), tiles, splits);
  return (int)cudaGetLastError();
}

}  // extern "C"
