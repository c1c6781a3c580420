// On-device litmus test of the memory-ordering pattern every aggregation kernel relies on (the reference has no
// race checking at all; here the cross-GPU protocol is exercised directly, thousands of rounds inside ONE kernel launch):
//
//   writer GPU                                     reader GPU
//   all threads: plain stores  data[i] = round     thread 0: spin  ld.acquire.sys(flag) >= round
//   bar.sync                                        bar.sync
//   thread 0: fence.sys ; st.release.sys(flag)      all threads: read the writer's data[i]
//             (flag lives on the reader)                 (a) peer load  ld.global.L1::no_allocate
//   thread 0: spin ld.acquire.sys(ack) >= round          (b) multimem.ld_reduce over the multicast mapping (own copy is 0)
//                                                    any value != round is a violation (message passing broken)
//                                                    bar.sync ; thread 0: st.release.sys(ack on the writer)
//
// This is exactly "gradients written by earlier threads -> bar.sync -> one thread fences and releases a flag on the peer ->
// the peer acquires the flag -> its other threads read the data over NVLink / through the NVSwitch reduction".
#include "fused_sync.cuh"

namespace dm {

struct LitmusArgs {
  float* data_local;            // my symmetric data buffer (n floats)
  const float* data_peer;       // the writer's buffer as mapped into the reader (P2P)
  const float* data_mc;         // multicast mapping (or null)
  volatile uint32_t* flag_local;   // words in MY control area: [0] = flag (written by the writer), [32] = ack (written by the reader)
  volatile uint32_t* flag_peer;    // the same words on the peer
  unsigned long long* result;   // [0] violations (peer loads), [1] violations (multimem), [2] rounds completed
  int n, rounds, role;          // role 0 = writer, 1 = reader
  unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(256, 1) litmus_mp_kernel(LitmusArgs a) {
  __shared__ int s_abort;
  unsigned long long bad_p2p = 0, bad_mc = 0;
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  for (int round = 1; round <= a.rounds; ++round) {
    if (a.role == 0) {
      for (int i = threadIdx.x; i < a.n; i += blockDim.x) a.data_local[i] = (float)round;     // weak stores
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(a.flag_peer, (uint32_t)round);
        if (!spin_until([&] { return ld_relaxed_sys(a.flag_local + 32) >= (uint32_t)round; }, a.timeout_ns, a.flag_local + 32)) s_abort = 1;
      }
      __syncthreads();
    } else {
      if (threadIdx.x == 0)
        if (!spin_until([&] { return ld_relaxed_sys(a.flag_local) >= (uint32_t)round; }, a.timeout_ns, a.flag_local)) s_abort = 1;
      __syncthreads();
      if (!s_abort) {
        for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
          float v;
          asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(a.data_peer + i) : "memory");
          if (v != (float)round) ++bad_p2p;
        }
        if (a.data_mc != nullptr) {
          for (int i = threadIdx.x * 4; i + 3 < a.n; i += blockDim.x * 4) {
            const float4 v = multimem_ld_reduce_f4(a.data_mc + i);          // sum over {writer: round, reader: 0}
            if (v.x != (float)round || v.y != (float)round || v.z != (float)round || v.w != (float)round) ++bad_mc;
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys(a.flag_peer + 32, (uint32_t)round);
    }
    if (s_abort) break;
    if (threadIdx.x == 0) a.result[2] = (unsigned long long)round;
  }
  if (bad_p2p) atomicAdd(&a.result[0], bad_p2p);
  if (bad_mc) atomicAdd(&a.result[1], bad_mc);
  if (threadIdx.x == 0 && s_abort) a.result[3] = 1ull;
}

}  // namespace dm

extern "C" int dm_litmus_mp(int role, void* data_local, const void* data_peer, const void* data_mc, void* flags_local,
                            void* flags_peer, void* result4, int n, int rounds, double timeout_ms, void* stream_) {
  using namespace dm;
  LitmusArgs a;
  a.data_local = reinterpret_cast<float*>(data_local);
  a.data_peer = reinterpret_cast<const float*>(data_peer);
  a.data_mc = reinterpret_cast<const float*>(data_mc);
  a.flag_local = reinterpret_cast<volatile uint32_t*>(flags_local);
  a.flag_peer = reinterpret_cast<volatile uint32_t*>(flags_peer);
  a.result = reinterpret_cast<unsigned long long*>(result4);
  a.n = n; a.rounds = rounds; a.role = role;
  a.timeout_ns = (unsigned long long)(timeout_ms * 1e6);
  litmus_mp_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(a);
  return (int)cudaGetLastError();
}
