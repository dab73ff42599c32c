// fm_peer.cu -- parameter averaging over NVLink peer memory (the per-epoch exchange of
// the row-sharded multi-GPU path), without NCCL.
//
// The packed fp32 state of C2 is 0.6 MB: an NCCL all-reduce of that size is pure
// latency (~50 us measured next to an ~80 us epoch).  Every rank instead maps the
// peers' state buffers (CUDA IPC across processes, plain peer access inside one
// process) and ONE kernel per rank does a one-shot all-reduce:
//
//   1. signal: write this epoch's sequence number into slot [self] of every peer's
//      flag block (st.release.sys over NVLink),
//   2. wait until the own flag block shows the sequence number for all ranks
//      (ld.acquire.sys) -- every peer's epoch kernel has then finished (stream order
//      on the peer) and its state buffer `cur` is final,
//   3. read all G `cur` buffers (peer loads travel NVLink) in rank order -- the same
//      order on every rank, so all replicas end bit-identical -- and write the mean
//      into the LOCAL `next` buffer.
//
// The context then swaps cur/next.  Double buffering removes the second barrier: a
// slow peer may still read our old `cur` while we already train into `next`; the old
// buffer is only overwritten by the NEXT averaging kernel, after that kernel's
// barrier has proven that every peer finished this one.
#include <algorithm>

#include "fmb200_internal.h"

namespace fmb {

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PeerArgs {
  unsigned int* flags[FMB200_MAX_PEERS];  // flag block of every rank (mapped)
  const float4* cur[FMB200_MAX_PEERS];    // current state buffer of every rank (mapped)
  float4* next_local;
  int world, rank;
  unsigned int seq;
  uint64_t n_vec;  // float4 elements
  float inv_world;
};

__global__ void __launch_bounds__(256) fm_peer_mean_kernel(const PeerArgs a) {
  if (blockIdx.x == 0 && threadIdx.x < a.world) st_release_sys(a.flags[threadIdx.x] + a.rank, a.seq);
  if (threadIdx.x < a.world) {
    const unsigned int* mine = a.flags[a.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(mine) - a.seq) < 0) {
    }
  }
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < a.n_vec;
       i += (uint64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < a.world; q++) {
      const float4 v = __ldcv(a.cur[q] + i);  // never from a stale L1 line
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    s.x *= a.inv_world;
    s.y *= a.inv_world;
    s.z *= a.inv_world;
    s.w *= a.inv_world;
    a.next_local[i] = s;
  }
}

// Cross-GPU barrier on the stream (no data): same signal / wait through the peers' flag
// blocks, on its own flag row and sequence so it never interferes with the averaging.
__global__ void fm_peer_barrier_kernel(const PeerArgs a) {
  if (threadIdx.x < a.world) {
    st_release_sys(a.flags[threadIdx.x] + FMB200_MAX_PEERS + a.rank, a.seq);
    const unsigned int* mine = a.flags[a.rank] + FMB200_MAX_PEERS + threadIdx.x;
    while ((int)(ld_acquire_sys(mine) - a.seq) < 0) {
    }
  }
}

cudaError_t launch_peer_barrier(fmb200_ctx* c) {
  PeerArgs a;
  for (int q = 0; q < c->peer_world; q++) a.flags[q] = reinterpret_cast<unsigned int*>(c->peer_base[q]);
  a.world = c->peer_world;
  a.rank = c->peer_rank;
  a.seq = ++c->peer_bar_seq;
  fm_peer_barrier_kernel<<<1, 32, 0, c->stream>>>(a);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_peer_mean(fmb200_ctx* c) {
  PeerArgs a;
  const int cur = c->peer_cur;
  for (int q = 0; q < c->peer_world; q++) {
    a.flags[q] = reinterpret_cast<unsigned int*>(c->peer_base[q]);
    a.cur[q] = reinterpret_cast<const float4*>(c->peer_base[q] + c->comm_hdr + (size_t)cur * c->comm_buf_bytes);
  }
  a.next_local = reinterpret_cast<float4*>(c->comm_base + c->comm_hdr + (size_t)(cur ^ 1) * c->comm_buf_bytes);
  a.world = c->peer_world;
  a.rank = c->peer_rank;
  a.seq = ++c->peer_seq;
  a.n_vec = (c->p32.n_floats + 3) / 4;
  a.inv_world = 1.f / (float)c->peer_world;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((a.n_vec + 255) / 256, (uint64_t)c->sm_count * 2));
  fm_peer_mean_kernel<<<grid, 256, 0, c->stream>>>(a);
  c->launches++;
  c->peer_cur = cur ^ 1;
  c->p32.base = reinterpret_cast<float*>(c->comm_base + c->comm_hdr + (size_t)c->peer_cur * c->comm_buf_bytes);
  return cudaGetLastError();
}

}  // namespace fmb
