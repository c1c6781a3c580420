// Hardware probe: can a tcgen05 shared-memory descriptor start at an arbitrary ROW offset inside a
// TMA-written swizzled buffer (needed to slide a 5x5 filter window over one resident activation patch),
// and which `base_offset` convention does that require?
//
// A buffer of 144 rows x 64 bf16 (128 B rows, SWIZZLE_128B) or x 32 bf16 (64 B rows, SWIZZLE_64B) is loaded by
// TMA; for every row shift r the kernel multiplies rows [r, r+128) by a 64x64 (or 32-wide K) matrix B and
// writes D.  mode 0: base_offset = 0;  mode 1: base_offset = (start_address >> 7) & 7.
// The host compares with the exact product, so the result table says which encoding the hardware expects.
#include "../common.cuh"
#include "../host_utils.h"

namespace dm {

template <int ROW_BYTES>   // 128 or 64
__global__ void __launch_bounds__(128, 1)
umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        float* __restrict__ out, int shift, int mode) {
  constexpr int K = ROW_BYTES / 2;              // bf16 elements per row
  constexpr uint32_t SWZ = ROW_BYTES == 128 ? SWZ_128B : SWZ_64B;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                           // 144 rows
  uint8_t* sB = smem + 144 * 128;               // 64 rows (N) x K, K-major
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 144 * 128 + 64 * 128);
  uint64_t* done = bar + 1;
  uint32_t* holder = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(holder);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *holder;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 144 * ROW_BYTES + 64 * ROW_BYTES);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t a_addr = smem_u32(sA) + shift * ROW_BYTES;
    const uint32_t b_addr = smem_u32(sB);
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    for (int k = 0; k < K / 16; ++k) {
      uint64_t da = make_smem_desc(a_addr + 32 * k, 16, 8 * ROW_BYTES, SWZ);
      if (mode == 1) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
      const uint64_t db = make_smem_desc(b_addr + 32 * k, 16, 8 * ROW_BYTES, SWZ);
      umma_bf16(tmem, da, db, idesc, k != 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after_sync();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

}  // namespace dm

// A: [144][K] bf16, B: [64][K] bf16, out: [128][64] fp32 = A[shift:shift+128] @ B^T.  row_bytes in {64, 128}.
extern "C" int dm_umma_shift_probe(const void* A, const void* B, void* out, int row_bytes, int shift, int mode,
                                   void* stream) {
  using namespace dm;
  const int K = row_bytes / 2;
  CUtensorMap tmA, tmB;
  if (make_tmap_2d_bf16(&tmA, A, K, 144, K, K, 144, row_bytes)) return 101;
  if (make_tmap_2d_bf16(&tmB, B, K, 64, K, K, 64, row_bytes)) return 102;
  const int smem = 144 * 128 + 64 * 128 + 64 + 1024;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (row_bytes == 128) {
    DM_CUDA_OK(cudaFuncSetAttribute(umma_shift_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_shift_probe_kernel<128><<<1, 128, smem, st>>>(tmA, tmB, reinterpret_cast<float*>(out), shift, mode);
  } else {
    DM_CUDA_OK(cudaFuncSetAttribute(umma_shift_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_shift_probe_kernel<64><<<1, 128, smem, st>>>(tmA, tmB, reinterpret_cast<float*>(out), shift, mode);
  }
  return (int)cudaGetLastError();
}
