// tcgen05 GEMM for sm_100a:  D[M,N] (+)= A[M,K] * B[K,N],  bf16 in, fp32 accumulate in TMEM.
//
// Covers every dense GEMM of the models (reference ops K6/K8: tf.matmul, src/mnist.py:136,145,
// and their gradients) without a single transpose kernel: each operand may be K-major or
// MN-major in memory, so the TF layouts ([in,out] weights, [batch,features] activations) are
// consumed as they are:
//     fc fwd    D[b,out]  = X[b,in]   (A K-major)  * W[in,out]  (B MN-major)
//     fc dgrad  D[b,in]   = dY[b,out] (A K-major)  * W[in,out]  (B K-major:  rows=in, K=out contiguous)
//     fc wgrad  D[in,out] = X[b,in]   (A MN-major) * dY[b,out]  (B MN-major),  K = batch
//
// Structure (one 128 x BN output tile per CTA, optional split-K over gridDim.z):
//     warp 0      TMA producer      cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx
//     warp 1      MMA issuer        one elected lane issues tcgen05.mma; tcgen05.commit frees the stage
//     warps 2-5   epilogue          tcgen05.ld (32 lanes x 32 columns) -> registers -> global
#include "common.cuh"
#include "host_utils.h"

namespace dm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;

enum GemmEpilogue : int { EPI_STORE_F32 = 0, EPI_ATOMIC_F32 = 1, EPI_STORE_BF16 = 2, EPI_BIAS_RELU_BF16 = 3,
                          // fc1 dgrad fused with the backward of maxpool2 + ReLU2 (reference K4/K3 gradients): a 64-column
                          // tile is one pooled position x 64 channels; the epilogue scatters each value to the argmax
                          // pixel of its 2x2 window (dense bf16 [B,14,14,64], zeros elsewhere) and reduces the conv2
                          // bias gradient -- the separate unpool kernel and the [B,3136] intermediate disappear.
                          EPI_UNPOOL2_BF16 = 4 };

struct GemmParams {
  int M, N;        // logical output extent (predication)
  int num_kb;      // ceil(K / 64)
  int ldo;         // output leading dimension (elements)
  void* out;
  long long split_stride;   // EPI_STORE_F32 with split-K: split z writes its partial tile at out + z*split_stride
  const float* bias;        // EPI_BIAS_RELU_BF16: per-column bias (reference K2/K3 fused into the producer)
  const uint8_t* code;      // EPI_UNPOOL2_BF16: pooling codes [M][N] (bits 0-1 argmax position, bit 2 ReLU active)
  float* g_bias;            // EPI_UNPOOL2_BF16: conv2 bias gradient [64], accumulated atomically (pre-zeroed)
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // (a 3-deep ring for the 128-wide tiles would let fc1_wgrad share SMs with fc1_dgrad; measured: both then take 12 us
  //  instead of 9.4 + 7.8 back to back and the critical path loses 3 us -- profiles/bench_r1_call35_1gpu.txt -- so 4 it is)
  static constexpr int STAGES = GEMM_STAGES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + alignment slack
};

// The unpool epilogue is instruction-heavy: it gets eight epilogue warps (two per TMEM lane quarter, one 32-column
// chunk each) instead of four, so two warps share every scheduler and hide each other's latencies.
template <int EPI>
constexpr int gemm_threads() { return EPI == EPI_UNPOOL2_BF16 ? GEMM_THREADS + 128 : GEMM_THREADS; }

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms repeat every 1024 B: align the ring so descriptors need no base offset.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full_bar = empty_bar + S::STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GEMM_BM;
  const int n0 = blockIdx.y * BN;
  // split-K: this CTA owns k-blocks [kb_begin, kb_end)
  const int kb_begin = (int)(((long long)p.num_kb * blockIdx.z) / gridDim.z);
  const int kb_end = (int)(((long long)p.num_kb * (blockIdx.z + 1)) / gridDim.z);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();      // everything above is private set-up; operands / outputs belong to earlier kernels until here
  // (no early launch_dependents: resident-but-blocked CTAs of the next kernel steal SM resources from this one)

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S::STAGES;
        const uint32_t ph = (i / S::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sA = smem + s * S::STAGE_BYTES;
        uint8_t* sB = sA + S::A_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        const int k0 = (kb_begin + i) * GEMM_BK;
        if (!A_MN) {
          tma_load_2d(sA, &tmA, &full_bar[s], k0, m0);                 // box {64 k, 128 m}
        } else {
          tma_load_2d(sA, &tmA, &full_bar[s], m0, k0);                 // box {64 m, 64 k}
          tma_load_2d(sA + 8192, &tmA, &full_bar[s], m0 + 64, k0);
        }
        if (!B_MN) {
          tma_load_2d(sB, &tmB, &full_bar[s], k0, n0);                 // box {64 k, BN n}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S::STAGES;
      const uint32_t ph = (i / S::STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        // K-major:  128 B rows (64 bf16 of K), 8-row groups 1024 B apart; K advances 32 B per UMMA_K.
        // MN-major: 128 B k-rows (64 bf16 of M/N), 8-k groups 1024 B apart, 64-wide M/N chunks 8192 B
        //           apart; K advances 16 k-rows = 2048 B per UMMA_K.
        const uint64_t da0 = A_MN ? make_smem_desc(a_addr, 8192, 1024, SWZ_128B)
                                  : make_smem_desc(a_addr, 16, 1024, SWZ_128B);
        const uint64_t db0 = B_MN ? make_smem_desc(b_addr, 8192, 1024, SWZ_128B)
                                  : make_smem_desc(b_addr, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          const uint64_t da = desc_advance(da0, (A_MN ? 2048u : 32u) * k);
          const uint64_t db = desc_advance(db0, (B_MN ? 2048u : 32u) * k);
          umma_bf16(tmem_base, da, db, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[s]);                  // stage reusable once these MMAs retire
        if (i == nkb - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue -----------------------------------
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const int c_begin = EPI == EPI_UNPOOL2_BF16 ? (warp - 2) >> 2 : 0;
    const int c_end = EPI == EPI_UNPOOL2_BF16 ? c_begin + 1 : BN / 32;
    uint32_t cw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (EPI == EPI_UNPOOL2_BF16 && row < p.M) {
      // pooling codes of this thread's 32 channels: an input of the forward pass, fetched while the MMAs run
      const uint4* cp = reinterpret_cast<const uint4*>(p.code + (size_t)row * p.N + n0 + c_begin * 32);
      const uint4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
      cw[0] = c0.x; cw[1] = c0.y; cw[2] = c0.z; cw[3] = c0.w; cw[4] = c1.x; cw[5] = c1.y; cw[6] = c1.z; cw[7] = c1.w;
    }
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0;
      }
      const int col0 = n0 + c * 32;
      if (row < p.M) {
        if (EPI == EPI_STORE_F32) {
          float* o = reinterpret_cast<float*>(p.out) + (size_t)blockIdx.z * p.split_stride + (size_t)row * p.ldo + col0;
          if (col0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                               __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) o[j] = __uint_as_float(v[j]);
          }
        } else if (EPI == EPI_ATOMIC_F32) {
          float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
          if (col0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_f32x4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(o + j, __uint_as_float(v[j]));
          }
        } else if (EPI == EPI_UNPOOL2_BF16) {
          // handled below (needs the whole warp, also for rows beyond M)
        } else {
          if (EPI == EPI_BIAS_RELU_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float b = (col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
              v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]) + b, 0.f));
            }
          }
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0;
          if (col0 + 32 <= p.N && (p.ldo & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              w.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              w.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
              w.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
              *reinterpret_cast<uint4*>(o + j) = w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) o[j] = __float2bfloat16(__uint_as_float(v[j]));
          }
        }
      }
      if (EPI == EPI_UNPOOL2_BF16) {
        // thread = batch row, 32 consecutive channels (ch0 ..) of pooled position `pos`
        const int pos = col0 >> 6, ch0 = col0 & 63;
        const int ph = pos / 7, pw = pos - ph * 7;
        float gs[32];                                   // masked gradient (bias-gradient contribution)
        if (row < p.M) {
          // code byte = argmax position (bits 0-1) | ReLU active (bit 2).  Byte-parallel: for window position q4,
          // x = (codes & 7) ^ (4 | q4) is zero exactly in the matching bytes; x + 0x7f sets bit 7 of the others
          // (no carries: x < 8), so ~(x + 0x7f..) & 0x80.. flags the matches and PRMT's sign-replicate mode widens
          // each flag to a 16-bit lane mask for the packed bf16 pairs.
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
#pragma unroll
          for (int j = 0; j < 32; ++j) gs[j] = ((cw[j >> 2] >> ((j & 3) * 8)) & 4u) ? __uint_as_float(v[j]) : 0.f;
          __nv_bfloat16* dy = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {              // the four pixels of the 2x2 window: 64 B each
            const int y = 2 * ph + (q4 >> 1), x = 2 * pw + (q4 & 1);
            uint4* dst = reinterpret_cast<uint4*>(dy + (((size_t)row * 14 + y) * 14 + x) * 64 + ch0);
            uint32_t o[16];
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) {            // code word w8 <-> channels 4*w8 .. 4*w8+3 <-> pk[2*w8], pk[2*w8+1]
              const uint32_t xz = (cw[w8] & 0x07070707u) ^ (0x04040404u | (0x01010101u * (uint32_t)q4));
              const uint32_t hit = ~(xz + 0x7f7f7f7fu) & 0x80808080u;
              o[2 * w8] = pk[2 * w8] & prmt(hit, 0u, 0x9988u);          // bytes {0,0,1,1} sign-replicated
              o[2 * w8 + 1] = pk[2 * w8 + 1] & prmt(hit, 0u, 0xbbaau);  // bytes {2,2,3,3}
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) gs[j] = 0.f;
        }
        // column sums over the warp's 32 rows: butterfly transpose-reduce (31 shuffles), lane l ends with channel ch0 + l
#pragma unroll
        for (int st = 0; st < 5; ++st) {
          const int off = 16 >> st;                     // compile-time after unrolling: gs[] stays in registers
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < off) {
              const float send = upper ? gs[i] : gs[i + off];
              const float keep = upper ? gs[i + off] : gs[i];
              gs[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
        }
        atomicAdd(p.g_bias + ch0 + lane, gs[0]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int splits,
                       cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI>;
  constexpr int smem = GemmSmem<BN>::TOTAL;
  static bool configured = false;
  if (!configured) {
    DM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  dim3 grid((p.M + GEMM_BM - 1) / GEMM_BM, (p.N + BN - 1) / BN, splits);
  return (int)launch_kernel(kern, grid, dim3(gemm_threads<EPI>()), smem, stream, tmA, tmB, p);
}

template <int BN, int EPI>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                          int splits, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, EPI>(tmA, tmB, p, splits, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, EPI>(tmA, tmB, p, splits, stream);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false, EPI>(tmA, tmB, p, splits, stream);
  return launch_gemm<BN, true, true, EPI>(tmA, tmB, p, splits, stream);
}

}  // namespace dm

// D[M,N] = A * B with bf16 operands.
//   a_mn == 0: A is [M rows][K cols] (lda elements between rows);   a_mn == 1: A is [K rows][M cols].
//   b_mn == 0: B is [N rows][K cols] (ldb);                          b_mn == 1: B is [K rows][N cols].
//   epi: 0 store fp32, 1 atomicAdd fp32 (split-K; caller zeroes), 2 store bf16, 3 bias+ReLU -> bf16.   bn: 64, 128 or 256.
//   split-K without atomics: epi 0 with splits > 1 and split_stride > 0 -> partial sums at out + z*split_stride
//   (the consumer adds the `splits` partials while loading).
extern "C" int dm_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, int lda, int ldb, int ldo,
                            int a_mn, int b_mn, int epi, int splits, int bn, long long split_stride, const void* bias,
                            void* stream_) {
  using namespace dm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if ((bn != 64 && bn != 128 && bn != 256) || epi < 0 || epi > 3 || splits < 1) return -1;
  if (epi == EPI_BIAS_RELU_BF16 && bias == nullptr) return -4;
  if (splits > 1 && !(epi == EPI_ATOMIC_F32 || (epi == EPI_STORE_F32 && split_stride > 0))) return -2;
  if ((lda & 7) || (ldb & 7)) return -3;   // TMA: global strides are multiples of 16 bytes
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 128, 128);
  else       rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64, 128);
  if (rc) return 100 + rc;
  if (!b_mn) rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)bn, 128);
  else       rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64, 128);
  if (rc) return 200 + rc;
  GemmParams p{M, N, (K + GEMM_BK - 1) / GEMM_BK, ldo, out, splits > 1 ? split_stride : 0,
               reinterpret_cast<const float*>(bias), nullptr, nullptr};
  if (splits > p.num_kb && epi == EPI_ATOMIC_F32) splits = p.num_kb > 0 ? p.num_kb : 1;
#define DM_DISPATCH(BN_, EPI_) return dispatch_major<BN_, EPI_>(a_mn != 0, b_mn != 0, tmA, tmB, p, splits, stream)
  if (bn == 64) {
    if (epi == 0) DM_DISPATCH(64, 0);
    if (epi == 1) DM_DISPATCH(64, 1);
    if (epi == 2) DM_DISPATCH(64, 2);
    DM_DISPATCH(64, 3);
  } else if (bn == 128) {
    if (epi == 0) DM_DISPATCH(128, 0);
    if (epi == 1) DM_DISPATCH(128, 1);
    if (epi == 2) DM_DISPATCH(128, 2);
    DM_DISPATCH(128, 3);
  } else {
    // 128 x 256 tiles: twice the FLOPs per operand byte fetched from L2 (the large MLP GEMMs are L2->SM bandwidth bound with
    // 128 x 128 tiles: 2048 CTAs x 2 MB of operands each), M128 N256 K16 instructions run at the tensor-pipe floor
    if (epi == 0) DM_DISPATCH(256, 0);
    if (epi == 2) DM_DISPATCH(256, 2);
    if (epi == 3) DM_DISPATCH(256, 3);
    return -1;
  }
#undef DM_DISPATCH
}

// fc1 dgrad fused with the maxpool2/ReLU2 backward and the conv2 bias gradient:
//   dxfc[b, j] = sum_k dh[b, k] * W1[j, k]   (A = dh [B,512] K-major, B = W1 [3136,512] rows = j, K contiguous)
//   dy2[b, 2ph+dy, 2pw+dx, c] = dxfc[b, (ph*7+pw)*64 + c] where code says (dy,dx) is the argmax and ReLU was active
//   g_bias2[c] += sum of the scattered values.          dy2: [B,14,14,64] bf16, fully written (zeros included).
extern "C" int dm_fc1_dgrad_unpool(const void* dh, const void* w1_bf16, const void* code2, void* dy2, void* g_bias2, int B,
                                   void* stream_) {
  using namespace dm;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, dh, 512, (uint64_t)B, 512, 64, 128, 128);
  if (rc) return 100 + rc;
  rc = make_tmap_2d_bf16(&tmB, w1_bf16, 512, 3136, 512, 64, 64, 128);
  if (rc) return 200 + rc;
  GemmParams p{B, 3136, 512 / GEMM_BK, 0, dy2, 0, nullptr, reinterpret_cast<const uint8_t*>(code2),
               reinterpret_cast<float*>(g_bias2)};
  return launch_gemm<64, false, false, EPI_UNPOOL2_BF16>(tmA, tmB, p, 1, reinterpret_cast<cudaStream_t>(stream_));
}
