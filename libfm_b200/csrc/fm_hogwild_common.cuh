// fm_hogwild_common.cuh -- pieces shared by the two HOGWILD epoch kernels
// (fm_hogwild.cu: sub-warp row groups, any k / row length; fm_rowlane.cu: one lane
// per row for short rows with k <= 8): launch arguments, the TMA tile producer and
// the mean-field step scale.  See fm_hogwild.cu for the design notes.
#pragma once
#include "fm_device.cuh"
#include "fmb200_internal.h"

namespace fmb {

constexpr int HW_NSTAGE = 3;
constexpr int HW_MAX_THREADS = 256;
constexpr int HW_HDR_BYTES = 256;  // mbarriers [0,64) + per-warp bias partials [64,192): [2 slots][8 warps] float2

struct HogwildArgs {
  const uint64_t* row_ptr;
  const uint32_t* col;
  const float* val;
  const float* target;
  uint64_t n_rows;
  uint32_t n_tiles;
  int tile_rows;       // TR (multiple of 32)
  uint32_t tile_cap;   // max staged entries per tile (multiple of 4)
  uint32_t stage_bytes;
  float* w0;
  float* w;
  float* v;
  int gp;  // float4 chunks per V row (kp / 4)
  int ws;  // stride of w in floats
  int use_w0, use_w, task;
  float lr, reg0, regw, regv, min_target, max_target;
  const float* feat_cnt;  // occurrences of each feature in this data set (DAMP)
  float conc_scale;       // rows processed concurrently / n_rows: count -> concurrency
  float w0_conc;          // rows in flight w.r.t. the bias (tile granularity)
  float hot_thr;          // COMBINE: occurrence count from which a feature is parked in the CTA's hot table
  unsigned int* sched;    // [0] next unclaimed tile, [1] CTAs that ran dry (both 0 between launches)
  int global_entries;     // rows too long for the staging ring: ids / values are read from global
                          // memory, only row offsets and targets are staged (tile_cap == 0)
};

__device__ __forceinline__ unsigned char* stage_base(unsigned char* smem, const HogwildArgs& a,
                                                     int stage) {
  return smem + HW_HDR_BYTES + (size_t)stage * a.stage_bytes;
}

// TMA producer: stage one tile whose entry range [nb, ne) is already known.
__device__ __forceinline__ void issue_tile(const HogwildArgs& a, unsigned char* smem,
                                           uint64_t* bars, uint32_t tile, int stage,
                                           uint64_t policy, uint64_t nb, uint64_t ne) {
  const int TR = a.tile_rows;
  const uint64_t r0 = (uint64_t)tile * TR;
  const uint64_t ab = nb & ~3ull;
  const uint64_t ae = (ne + 3ull) & ~3ull;
  const uint32_t ebytes = a.global_entries ? 0u : (uint32_t)(ae - ab) * 4u;
  const uint32_t rp_bytes = (uint32_t)(TR + 2) * 8u;
  const uint32_t y_bytes = (uint32_t)TR * 4u;
  unsigned char* sb = stage_base(smem, a, stage);
  uint64_t* bar = bars + stage;
  mbar_arrive_expect_tx(bar, rp_bytes + y_bytes + 2u * ebytes);
  bulk_g2s_hint(sb, a.row_ptr + r0, rp_bytes, bar, policy);
  bulk_g2s_hint(sb + rp_bytes, a.target + r0, y_bytes, bar, policy);
  if (ebytes) {
    unsigned char* cb = sb + rp_bytes + y_bytes;
    bulk_g2s_hint(cb, a.col + ab, ebytes, bar, policy);
    bulk_g2s_hint(cb + (size_t)a.tile_cap * 4u, a.val + ab, ebytes, bar, policy);
  }
}

// gamma(c, u) = (1 - (1-u)^c) / (c*u): scale of each of c concurrent steps whose
// sequential execution would contract the residual by (1-u) per step
__device__ __forceinline__ float gamma_scale(float c, float u) {
  if (c <= 1.f || u <= 0.f) return 1.f;
  const float q = c * u;
  if (q < 1e-3f) return 1.f;
  const float a = fmaxf(1.f - u, 0.f);
  const float ac = a > 0.f ? __expf(c * __logf(a)) : 0.f;
  return fminf(1.f, (1.f - ac) / q);
}


// The bias sector is reduced into by every CTA once per tile; loads of it queue behind
// those reductions at its single L2 slice.  So ONE lane per CTA fetches it per tile
// (issued at the top of the tile, consumed after the gathers are in flight) and hands
// it to the other warps through shared memory + named barrier 1.
struct BiasFetch {
  float* slot;   // [2] floats in the smem header, indexed by tile parity
  float pending; // lane 0 of warp 0: the in-flight value
  __device__ __forceinline__ void issue(const HogwildArgs& a, bool use_w0, int tid) {
    pending = 0.f;
    if (use_w0 && tid == 0) pending = ld_cg_f(a.w0);  // warp 0 fetches the bias
  }
  // returns the tile's bias in every thread of the CTA
  __device__ __forceinline__ float get(bool use_w0, int tid, int it, int nthreads) {
    if (!use_w0) return 0.f;
    float* s = slot + (it & 1);
    if (tid < 32) {
      const float v = __shfl_sync(0xffffffffu, pending, 0);
      if (tid == 0) *s = v;
      __threadfence_block();
      named_bar_arrive(1, nthreads);
      return v;
    }
    named_bar_sync(1, nthreads);
    return *s;
  }
};

// Dynamic tile scheduler: CTAs claim row tiles from a global counter in file order, so
// the tail of the epoch is balanced (a static round-robin leaves 1/9 of the CTAs a whole
// tile short on C2) and the rows in flight stay one contiguous window.  Used by thread 0
// only.  The last CTA to run dry resets the two words for the next launch.
constexpr uint32_t HW_NO_TILE = 0xffffffffu;
struct TileSched {
  unsigned int* w;
  uint32_t n_tiles;
  bool dry;
  __device__ __forceinline__ uint32_t claim() {
    if (dry) return HW_NO_TILE;
    const uint32_t t = atomicAdd(w, 1u);
    if (t >= n_tiles) {
      dry = true;
      return HW_NO_TILE;
    }
    return t;
  }
  // split claim: fire() only issues the atomic (its latency must not stall the producer
  // thread, which also processes rows); resolve() interprets the value a tile later
  __device__ __forceinline__ uint32_t fire() { return dry ? HW_NO_TILE : atomicAdd(w, 1u); }
  __device__ __forceinline__ uint32_t resolve(uint32_t raw) {
    if (raw >= n_tiles) {
      dry = true;
      return HW_NO_TILE;
    }
    return raw;
  }
  // `last_raw`: the claim still in flight.  Its value must have RETURNED (the counter
  // increment performed) before this CTA reports itself dry, or the reset below could be
  // overtaken by it and the next launch would start at tile 1.
  __device__ __forceinline__ void finish(unsigned int n_ctas, uint32_t last_raw) {
    unsigned int inc = (last_raw == 0x7fffffffu) ? 2u : 1u;  // data dependence on the return value
    __threadfence();
    if (atomicAdd(w + 1, inc) == n_ctas - 1) {
      w[0] = 0u;
      w[1] = 0u;
    }
  }
};

using HogwildKernelFn = void (*)(const HogwildArgs);

// fm_rowlane.cu: kernel for (float4 chunks per row gp in {1,2}, rows of at most Z entries)
HogwildKernelFn pick_rowlane_kernel(int gp, int max_row_nnz, bool damp, bool combine);
// warp-specialised variant: blockDim = rows_per_tile + 32, smem header 512 B
HogwildKernelFn pick_rowlane_ws_kernel(int gp, int max_row_nnz, bool damp, bool combine);

}  // namespace fmb
