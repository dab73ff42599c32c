// cta_shim.h -- TEST INFRASTRUCTURE: enough of the CUDA vocabulary to compile a multi-warp,
// single-CTA kernel body for the host and run it as one OS thread per CUDA thread.
//
// Same model as simt_shim.h (lanes run freely and meet only where the kernel synchronises),
// extended to a whole CTA: __syncthreads is a barrier over all threads, warp collectives
// (__shfl*_sync, __ballot_sync, __any_sync) are per-warp barriers around an exchange buffer,
// cp.async / TMA bulk copies are performed at ISSUE time (the earliest moment the hardware may
// read the source -- the most stale view the kernel has to tolerate), and an mbarrier is a
// pending-byte count plus a phase counter.  Everything else is the kernel's own source.
#pragma once
#define FMB_SIMT_HOST 1
#include <cuda_runtime.h>  // vector types + the annotation macros replaced below

#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

using std::max;
using std::min;

namespace simt {

constexpr int kMaxWarps = 32;
struct Dim {
  unsigned x;
};
extern thread_local Dim tid;
extern Dim bdim;
extern pthread_barrier_t cta_barrier;
extern pthread_barrier_t warp_barrier[kMaxWarps];
extern pthread_barrier_t named_barrier[4];
extern uint64_t xchg[kMaxWarps][32];

inline int warp() { return (int)(tid.x >> 5); }
inline int lane() { return (int)(tid.x & 31); }
inline void wsync() { pthread_barrier_wait(&warp_barrier[warp()]); }

template <class T>
inline T exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "exchange width");
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  __atomic_store_n(&xchg[warp()][lane()], bits, __ATOMIC_RELAXED);
  wsync();
  const uint64_t got = __atomic_load_n(&xchg[warp()][src & 31], __ATOMIC_RELAXED);
  wsync();  // nobody overwrites its slot before all have read
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace simt

#define threadIdx (simt::tid)
#define blockDim (simt::bdim)

inline void __syncthreads() { pthread_barrier_wait(&simt::cta_barrier); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::wsync(); }

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  return simt::exchange(v, src);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int d) {
  const int l = simt::lane();
  return simt::exchange(v, l >= d ? l - d : l);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
  return simt::exchange(v, simt::lane() ^ m);
}
inline unsigned __ballot_sync(unsigned, bool pred) {
  __atomic_store_n(&simt::xchg[simt::warp()][simt::lane()], pred ? 1ull : 0ull, __ATOMIC_RELAXED);
  simt::wsync();
  unsigned m = 0;
  for (int l = 0; l < 32; l++)
    if (__atomic_load_n(&simt::xchg[simt::warp()][l], __ATOMIC_RELAXED)) m |= 1u << l;
  simt::wsync();
  return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) {
  }
  return old;
}

// ---- the device primitives of fm_device.cuh the ordered kernel uses -------------------------
namespace fmb {

// mbarrier word: high 32 bits = completed phases, low 32 bits = pending transaction bytes
inline void mbar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_RELEASE); }
inline void fence_mbar_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  __atomic_fetch_add(bar, (uint64_t)bytes, __ATOMIC_ACQ_REL);
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  const uint64_t now = __atomic_sub_fetch(bar, (uint64_t)bytes, __ATOMIC_ACQ_REL);
  if ((now & 0xffffffffull) == 0) __atomic_fetch_add(bar, 1ull << 32, __ATOMIC_ACQ_REL);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (;;) {
    const uint64_t v = __atomic_load_n(bar, __ATOMIC_ACQUIRE);
    if ((v & 0xffffffffull) == 0 && (((v >> 32) & 1u) != parity)) return;
    sched_yield();
  }
}
inline uint64_t policy_evict_first() { return 0; }
inline void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t) {
  memcpy(dst, src, bytes);
  mbar_complete_tx(bar, bytes);
}
inline void cp_async_16(void* dst, const void* src) { memcpy(dst, src, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_1() {}
inline void cp_async_wait_0() {}
// bar.sync id, n: the host driver initialises named_barrier[id] for n threads before the launch
inline void named_bar_sync(int id, int) { pthread_barrier_wait(&simt::named_barrier[id]); }

}  // namespace fmb
