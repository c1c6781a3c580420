// Shared definitions of the gradient-aggregation kernels (csrc/fused_sync.cu: single-kernel K-of-N / interval path;
// csrc/fused_bucket.cu: bucketed, overlapped K == N path): the symmetric control block, the peer pointer tables and the
// system-scope memory / multimem helpers.
#pragma once

#include <stddef.h>
#include <string.h>

#include "common.cuh"
#include "host_utils.h"

namespace dm {

constexpr int SYNC_MAX_RANKS = 8;
constexpr int SYNC_RING = 64;          // commit-word ring (steps)
constexpr int SYNC_THREADS = 512;
constexpr int TIMING_RING = 1024;      // per-rank arrival-timestamp ring (steps)

// Control block, one per rank, in symmetric memory.  Words written remotely are spread over
// separate 128-byte lines.
struct SyncCtrl {
  // ---- written by peers -----------------------------------------------------------------
  volatile uint32_t arrive[SYNC_MAX_RANKS * 32];     // [p*32]: peer p arrived for epoch (value = epoch+1)
  volatile uint32_t done[SYNC_MAX_RANKS * 32];       // [p*32]: peer p's pushes for step s landed (value = s+1)
  volatile unsigned long long commit_local[SYNC_RING];  // ((step+1) << 32) | mask, broadcast by the committer
  // ---- authoritative on the chief only ------------------------------------------------------
  unsigned int bitmap[SYNC_RING];                    // arrival bitmap of step s at [s % RING]
  unsigned long long commit[SYNC_RING];              // ((step+1) << 32) | mask
  volatile uint32_t global_step;                     // number of committed steps
  volatile uint32_t last_in_mask[SYNC_MAX_RANKS];    // (last step in whose mask rank q was) + 1
  uint32_t pad0[32];
  // ---- local ----------------------------------------------------------------------------------
  uint32_t epoch;                // global step of my weights (device-resident so CUDA graphs replay)
  uint32_t cta_counter;          // grid-wide completion counter
  volatile uint32_t decided_tag; // epoch+1 once `decided_mask` is valid for this launch
  volatile uint32_t decided_mask;
  volatile uint32_t decided_late;   // 1: my gradient is not part of the mean
  volatile uint32_t decided_target; // epoch to move to after this launch
  uint32_t error;                // watchdog: 1 = arrival timeout, 2 = done timeout
  uint32_t accepted_steps, dropped_steps;
  uint32_t last_mask, last_count, last_late;
  uint32_t pad1[32];
  unsigned long long t_arrive[TIMING_RING];   // %globaltimer at arrival, per local step (cdf telemetry)
  unsigned long long t_start[TIMING_RING];    // %globaltimer stamped by the step's first kernel
  unsigned long long t_phase[8];              // last launch: kernel start, decided, reduced, pushed, landed, shadowed
  // ---- bucketed (overlapped) aggregation: the early bucket has its own handshake words ---------------------------
  volatile uint32_t arrive_e[SYNC_MAX_RANKS * 32];   // [p*32]: peer p's early-bucket gradients are complete (epoch+1)
  volatile uint32_t done_e[SYNC_MAX_RANKS * 32];     // [p*32]: peer p's early-bucket pushes have landed (epoch+1)
  uint32_t cta_counter_e;
  uint32_t cta_counter2;         // late bucket kernel: CTAs that have finished (the last one closes the step)
  uint32_t pad2[30];
  unsigned long long t_phase_e[8];            // early kernel: start, arrived, reduced (CTA 0), all pushed, all landed, applied (CTA 0)
};

struct SyncPeers {
  SyncCtrl* ctrl[SYNC_MAX_RANKS];
  float* params[SYNC_MAX_RANKS];
  const float* grads[SYNC_MAX_RANKS];
};

struct SyncArgs {
  int rank, nranks, k;
  int numel4;                 // arena length in float4
  float lr0, decay_rate;      // staircase exponential decay evaluated on device (reference K12)
  int decay_steps;
  float drop_keep;            // gradient drop-connect keep probability, <= 0: off (reference K14)
  uint32_t drop_seed;
  unsigned long long timeout_ns;
  __nv_bfloat16* shadow;      // local bf16 copy of the parameter arena (tensor-core operand source)
  // NVLS (NVSwitch multicast) views of the two arenas, null when the fabric has none.  A load from mc_grads returns the
  // SUM over all ranks computed inside the switch; a store to mc_params lands in every rank's arena.
  const float* mc_grads;
  float* mc_params;
};

// ---- system-scope memory helpers ----------------------------------------------------------------
DMNIST_DEVICE uint32_t ld_acquire_sys(const volatile uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DMNIST_DEVICE unsigned long long ld_acquire_sys64(const volatile unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DMNIST_DEVICE void st_release_sys(volatile uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
DMNIST_DEVICE void st_release_sys64(volatile unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DMNIST_DEVICE float4 ld_peer_f4(const float* p) {   // peer data: read once, keep out of L1
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
DMNIST_DEVICE void st_peer_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// In-switch reduction: one 16-byte load returns the element-wise fp32 sum of the same address on every rank.
DMNIST_DEVICE float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
// Multicast store: the switch replicates the 16 bytes into every rank's copy.
DMNIST_DEVICE void multimem_st_f4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

DMNIST_DEVICE float device_lr(const SyncArgs& a, uint32_t step) {
  const float p = (float)(step / (uint32_t)max(a.decay_steps, 1));
  return a.lr0 * __powf(a.decay_rate, p);
}

// Poll until pred() or the watchdog fires.
template <class Pred>
DMNIST_DEVICE bool spin_until(Pred pred, unsigned long long timeout_ns) {
  if (pred()) return true;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (!pred()) {
    if (++spins > 64) {          // busy-poll first (the common wait is a few microseconds), then back off
      __nanosleep(64);
      if ((spins & 255) == 0 && globaltimer_ns() - t0 > timeout_ns) return false;
    }
  }
  return true;
}

}  // namespace dm
