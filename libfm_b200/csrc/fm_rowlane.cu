// fm_rowlane.cu -- HOGWILD epoch kernel for SHORT ROWS and k <= 8: one lane per row.
//
// Same algorithm, staging and concurrency control as fm_sgd_hogwild_kernel
// (fm_hogwild.cu) -- this is the instruction-lean mapping for the one-hot
// "MovieLens" shape (BASELINE config C2: k=8, 2 nnz/row).  The sub-warp row-group
// kernel spends ~30 warp instructions per example there (segmented shuffles,
// per-lane address arithmetic replicated over 4 lanes per row); the measured
// bound of the shape is the SM's L1TEX/LSU rate for scattered 32-byte sectors
// (profiles/r01_red_microbench.txt), so everything else has to get out of the way.
//
// Mapping
//  * lane t of a CTA owns row t of the staged tile (rows_per_tile == blockDim.x);
//    all per-row math (fm_model::predict, reference fm_model.h:105-127; fm_SGD,
//    fm_sgd.h:33-51) happens in that lane's registers: no shuffles for sums.
//  * a factor row of k=8 floats is one 32-byte sector = two float4.  Letting every
//    lane fetch its own two halves would cost two sector requests per row and
//    instruction; instead lane PAIRS (2j, 2j+1) co-operate: instruction A fetches
//    the row of lane 2j (even lane low half, odd lane high half), instruction B the
//    row of lane 2j+1, and one 4-float __shfl_xor swaps the halves into place.
//    The write-back mirrors it (swap, then two red.global.add.v4.f32 whose lane
//    pairs cover one full sector each).  Sector requests per (row, entry): 1 gather
//    + 1 reduction for V, 1 + 1 for w -- the minimum for this layout.
//  * GP == 1 (k <= 4): a factor row is a single float4; every lane fetches its own.
#include <algorithm>

#include "fm_hogwild_common.cuh"

namespace fmb {

template <int GP>
struct FactorRow {
  float v[4 * GP];
};

// CTA-level write combining for the hottest features of skewed data (COMBINE kernels).
// Same-address L2 reductions serialise: a feature that appears in 11% of the rows costs
// the whole chip ~4 ns per reduction.  After the in-warp merge, steps for features with
// at least `hot_thr` occurrences are parked in a small direct-mapped shared-memory table
// (tag = feature id) and leave the CTA as ONE reduction per feature and tile.  A slot that
// is taken by another feature simply sends the step to L2 as before.
constexpr int HOT_SLOTS = 128;
struct HotEntry {
  uint32_t tag;   // feature id, HOT_EMPTY when free
  float dw;
  float dv[8];
  uint32_t pad[6];  // 64 bytes: one entry per pair of banks rows, no false sharing of tags
};
constexpr uint32_t HOT_EMPTY = 0xffffffffu;

// returns true when the step was parked in the table
template <int K>
__device__ __forceinline__ bool hot_park(HotEntry* tab, uint32_t id, const float (&d)[K], float dw) {
  HotEntry* e = tab + (id & (HOT_SLOTS - 1));
  const uint32_t prev = atomicCAS(&e->tag, HOT_EMPTY, id);
  if (prev != HOT_EMPTY && prev != id) return false;
#pragma unroll
  for (int f = 0; f < K; ++f) atomicAdd(&e->dv[f], d[f]);
  atomicAdd(&e->dw, dw);
  return true;
}

// flush one table: thread t handles slot t (call with t < HOT_SLOTS after a barrier)
template <int GP>
__device__ __forceinline__ void hot_flush(const HogwildArgs& a, HotEntry* tab, int t) {
  HotEntry* e = tab + t;
  const uint32_t id = e->tag;
  if (id == HOT_EMPTY) return;
  red_add_f4(a.v + (size_t)id * GP * 4, e->dv[0], e->dv[1], e->dv[2], e->dv[3]);
  if (GP == 2) red_add_f4(a.v + (size_t)id * GP * 4 + 4, e->dv[4], e->dv[5], e->dv[6], e->dv[7]);
  if (a.use_w) red_add_f(a.w + (size_t)id * a.ws, e->dw);
  e->tag = HOT_EMPTY;
  e->dw = 0.f;
#pragma unroll
  for (int f = 0; f < 8; ++f) e->dv[f] = 0.f;
}

// One tile, one lane per row: gather, score, multiplier, write-back.  `get_w0` is called
// once the gathers are in flight and returns the tile's bias.  Returns this lane's loss
// multiplier and joint curvature (zero for lanes without a row) for the bias step.
template <int GP, int Z, bool DAMP, bool COMBINE, typename W0F>
__device__ __forceinline__ void rowlane_tile(const HogwildArgs& a, const uint64_t* rp,
                                             const float* ys, const uint32_t* ids,
                                             const float* xs, int rows_here, int tid, W0F get_w0,
                                             float& mult_out, float& hjoint_out,
                                             HotEntry* hot_tab = nullptr) {
  constexpr int K = 4 * GP;
  const int lane = tid & 31;
  const int odd = lane & 1;
  const float4* V4 = reinterpret_cast<const float4*>(a.v);
  const bool use_w = a.use_w != 0;
  const bool use_w0 = a.use_w0 != 0;
  const float lr = a.lr;
  const uint64_t ab = rp[0] & ~3ull;
  // ---- this lane's row ----
  const bool valid = tid < rows_here;
  int beg = 0, cnt = 0;
  float y = 0.f;
  if (valid) {
    beg = (int)(rp[tid] - ab);
    cnt = (int)(rp[tid + 1] - ab) - beg;
    y = ys[tid];
  }
  uint32_t id[Z];
  float x[Z], wv[Z];
  FactorRow<GP> vr[Z];
  // ---- gather: all entries in flight at once ----
#pragma unroll
  for (int e = 0; e < Z; ++e) {
    const bool on = e < cnt;
    id[e] = on ? ids[beg + e] : 0u;
    x[e] = on ? xs[beg + e] : 0.f;
  }
#pragma unroll
  for (int e = 0; e < Z; ++e) {
    if (GP == 2) {
      // a missing entry (ragged rows) fetches nothing: an unconditional gather of feature 0's sector would add
      // L2 loads on a line the rows that really contain feature 0 are reducing into
      const uint32_t gid = (e < cnt) ? id[e] : 0xffffffffu;
      const uint32_t pid = __shfl_xor_sync(0xffffffffu, gid, 1);
      const uint32_t idA = odd ? pid : gid;  // row of the even lane
      const uint32_t idB = odd ? gid : pid;  // row of the odd lane
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 la = idA != 0xffffffffu ? ld_cg_f4(V4 + (size_t)idA * 2 + odd) : zero4;
      const float4 lb = idB != 0xffffffffu ? ld_cg_f4(V4 + (size_t)idB * 2 + odd) : zero4;
      const float4 send = odd ? la : lb;
      float4 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
      recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
      recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
      const float4 lo = odd ? recv : la;
      const float4 hi = odd ? lb : recv;
      vr[e].v[0] = lo.x; vr[e].v[1] = lo.y; vr[e].v[2] = lo.z; vr[e].v[3] = lo.w;
      vr[e].v[4] = hi.x; vr[e].v[5] = hi.y; vr[e].v[6] = hi.z; vr[e].v[7] = hi.w;
    } else {
      const float4 l = (e < cnt) ? ld_cg_f4(V4 + (size_t)id[e]) : make_float4(0.f, 0.f, 0.f, 0.f);
      vr[e].v[0] = l.x; vr[e].v[1] = l.y; vr[e].v[2] = l.z; vr[e].v[3] = l.w;
    }
    wv[e] = (use_w && e < cnt) ? ld_cg_f(a.w + (size_t)id[e] * a.ws) : 0.f;
  }

  // ---- fm_model::predict in registers (fm_model.h:105-127) ----
  float sum[K];
#pragma unroll
  for (int f = 0; f < K; ++f) sum[f] = 0.f;
  float sq = 0.f, lin = 0.f, xx = 0.f;
#pragma unroll
  for (int e = 0; e < Z; ++e) {
#pragma unroll
    for (int f = 0; f < K; ++f) {
      const float d = vr[e].v[f] * x[e];
      sum[f] += d;
      sq += d * d;
    }
    lin += wv[e] * x[e];
    xx += x[e] * x[e];
  }
  float s2 = 0.f;
#pragma unroll
  for (int f = 0; f < K; ++f) s2 += sum[f] * sum[f];
  const float w0 = get_w0();
  const float p = w0 + lin + 0.5f * (s2 - sq);

  // ---- loss multiplier (fm_learn_sgd_element.h:58-65) ----
  float mult, curv;
  if (a.task == FMB200_TASK_REGRESSION) {
    const float pc = fmaxf(a.min_target, fminf(a.max_target, p));
    mult = pc - y;
    const float den = p - y;
    curv = (pc == p) ? 1.f : (fabsf(den) > 1e-12f ? fminf(fmaxf(mult / den, 0.f), 1.f) : 0.f);
  } else {
    const float sg = 1.f / (1.f + __expf(-y * p));
    mult = -y * (1.f - sg);
    curv = sg * (1.f - sg);
  }
  if (!valid) {
    mult = 0.f;
    curv = 0.f;
  }
  // joint curvature of the row's whole parameter set (see fm_hogwild.cu)
  const float hrow = (use_w ? xx : 0.f) + fmaxf((xx - 2.f) * s2 + sq, 0.f);
  const float hjoint = DAMP ? curv * ((use_w0 ? 1.f : 0.f) + hrow) : curv;

  // ---- fm_SGD write-back (fm_sgd.h:38-50) ----
  const float nlr_mult = -lr * mult;
  const float nlr_regv = -lr * a.regv;
  const float nlr_regw = -lr * a.regw;
#pragma unroll
  for (int e = 0; e < Z; ++e) {
    const bool on = e < cnt;
    float sv = 1.f, sw = 1.f;
    if (DAMP) {
      const float conc = on ? __ldg(a.feat_cnt + id[e]) * a.conc_scale : 0.f;
      if (conc > 1.f) {
        sv = gamma_scale(conc, lr * (hjoint + a.regv));
        sw = gamma_scale(conc, lr * (hjoint + a.regw));
      }
    }
    const float x2 = x[e] * x[e];
    float d[K];
#pragma unroll
    for (int f = 0; f < K; ++f)
      d[f] = sv * (nlr_mult * (sum[f] * x[e] - vr[e].v[f] * x2) + nlr_regv * vr[e].v[f]);
    float dw = sw * (nlr_mult * x[e] + nlr_regw * wv[e]);
    bool on_c = on;  // this lane still owns a reduction for entry e
    if (COMBINE) {
      // Skewed data: several rows of a warp hit the same feature.  Sum their steps
      // inside the warp (log-step segmented reduction over the lanes that share the
      // id, after "Voting and Shuffling to Optimize Atomic Operations") and let the
      // lowest lane issue ONE reduction: hot rows serialise at L2, so every merged
      // reduction is time saved for the whole chip.
      const uint32_t key = on ? id[e] : (0x80000000u | (uint32_t)lane);  // inactive: unique
      unsigned peers = __match_any_sync(0xffffffffu, key);
      if (__any_sync(0xffffffffu, __popc(peers) > 1)) {
        const int first = __ffs(peers) - 1;
        int rel = __popc(peers << (31 - lane) << 1);  // peers below this lane
        peers &= (0xfffffffeu << lane);               // peers above this lane
        while (__any_sync(0xffffffffu, peers != 0u)) {
          const int next = __ffs(peers);
          const int src = next ? next - 1 : lane;
#pragma unroll
          for (int f = 0; f < K; ++f) {
            const float t = __shfl_sync(0xffffffffu, d[f], src);
            if (next) d[f] += t;
          }
          const float tw = __shfl_sync(0xffffffffu, dw, src);
          if (next) dw += tw;
          const unsigned done = __ballot_sync(0xffffffffu, rel & 1);
          peers &= ~done;
          rel >>= 1;
        }
        on_c = on && (lane == first);
      }
      // hot features leave the CTA once per tile
      if (hot_tab != nullptr && on_c && __ldg(a.feat_cnt + id[e]) >= a.hot_thr) {
        if (hot_park<K>(hot_tab, id[e], d, dw)) on_c = false;
      }
    }
    if (GP == 2) {
      // swap halves inside the lane pair so that each reduction covers a full sector
      const uint32_t pid = __shfl_xor_sync(0xffffffffu, id[e], 1);
      const bool pon = __shfl_xor_sync(0xffffffffu, (int)on_c, 1) != 0;
      const uint32_t idA = odd ? pid : id[e];
      const uint32_t idB = odd ? id[e] : pid;
      const bool onA = odd ? pon : on_c;
      const bool onB = odd ? on_c : pon;
      float4 send, keep;
      if (odd) {
        send = make_float4(d[0], d[1], d[2], d[3]);  // my low half goes to the even lane
        keep = make_float4(d[4], d[5], d[6], d[7]);
      } else {
        send = make_float4(d[4], d[5], d[6], d[7]);  // my high half goes to the odd lane
        keep = make_float4(d[0], d[1], d[2], d[3]);
      }
      float4 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
      recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
      recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
      // row A (even lane's): even writes its low half, odd writes the received high half
      const float4 va = odd ? recv : keep;
      // row B (odd lane's): even writes the received low half, odd writes its high half
      const float4 vb = odd ? keep : recv;
      if (onA) red_add_f4(a.v + ((size_t)idA * 2 + odd) * 4, va.x, va.y, va.z, va.w);
      if (onB) red_add_f4(a.v + ((size_t)idB * 2 + odd) * 4, vb.x, vb.y, vb.z, vb.w);
    } else {
      if (on_c) red_add_f4(a.v + (size_t)id[e] * 4, d[0], d[1], d[2], d[3]);
    }
    if (on_c && use_w) red_add_f(a.w + (size_t)id[e] * a.ws, dw);
  }

  mult_out = mult;
  hjoint_out = valid ? hjoint : 0.f;
}

template <int GP, int Z, bool DAMP, bool COMBINE>
__global__ void __launch_bounds__(HW_MAX_THREADS, 3) fm_sgd_rowlane_kernel(const HogwildArgs a) {
  constexpr int K = 4 * GP;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  float* s_acc = reinterpret_cast<float*>(smem + 64);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int odd = lane & 1;
  const int TR = a.tile_rows;  // == blockDim.x

  uint64_t policy = 0;
  if (tid == 0) {
    for (int i = 0; i < HW_NSTAGE; i++) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  if (tid == (int)blockDim.x - 32) policy = policy_evict_first();
  // COMBINE: two hot-feature tables (ping-pong by tile parity) behind the staging ring
  HotEntry* hot = reinterpret_cast<HotEntry*>(smem + HW_HDR_BYTES + (size_t)HW_NSTAGE * a.stage_bytes);
  if (COMBINE) {
    for (int i = tid; i < 2 * HOT_SLOTS; i += blockDim.x) {
      hot[i].tag = HOT_EMPTY;
      hot[i].dw = 0.f;
#pragma unroll
      for (int f = 0; f < 8; ++f) hot[i].dv[f] = 0.f;
    }
  }
  __syncthreads();
  uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem + 208);  // [HW_NSTAGE] tile staged per stage
  TileSched sched{a.sched, a.n_tiles, false};
  // producer duties (tile claims, TMA issue, bias reduction) sit on lane 0 of the LAST
  // warp; warp 0 fetches and publishes the bias -- nobody waits on the producer before
  // the end-of-tile barrier
  const int ptid = (int)blockDim.x - 32;
  uint32_t claim_raw = HW_NO_TILE;  // producer: a claim in flight (resolved one tile later)
  if (tid == ptid) {
    for (int i = 0; i < HW_NSTAGE; i++) {
      const uint32_t t = sched.claim();
      s_tile[i] = t;
      if (t != HW_NO_TILE) {
        const uint64_t r0 = (uint64_t)t * TR, r1 = min(r0 + (uint64_t)TR, a.n_rows);
        issue_tile(a, smem, bars, t, i, policy, __ldg(a.row_ptr + r0), __ldg(a.row_ptr + r1));
      }
    }
    claim_raw = sched.fire();
  }
  __syncthreads();

  const float4* V4 = reinterpret_cast<const float4*>(a.v);
  const bool use_w = a.use_w != 0;
  const bool use_w0 = a.use_w0 != 0;
  const float lr = a.lr;

  int it = 0;
  for (;; ++it) {
    const int stage = it % HW_NSTAGE;
    const uint32_t parity = (uint32_t)(it / HW_NSTAGE) & 1u;
    const uint64_t tile = s_tile[stage];
    if (tile == HW_NO_TILE) break;  // this CTA's claims ran dry
    uint32_t nt = HW_NO_TILE;
    uint64_t nt_nb = 0, nt_ne = 0;
    if (tid == ptid) {
      nt = sched.resolve(claim_raw);  // fired a tile ago: long since returned
      claim_raw = sched.fire();       // not looked at before the next tile
      if (nt != HW_NO_TILE) {
        const uint64_t r0 = (uint64_t)nt * TR, r1 = min(r0 + (uint64_t)TR, a.n_rows);
        nt_nb = __ldg(a.row_ptr + r0);
        nt_ne = __ldg(a.row_ptr + r1);
      }
    }
    BiasFetch bias;
    bias.slot = reinterpret_cast<float*>(smem + 192);
    bias.issue(a, use_w0, tid);
    mbar_wait(bars + stage, parity);

    unsigned char* sb = stage_base(smem, a, stage);
    const uint64_t* rp = reinterpret_cast<const uint64_t*>(sb);
    const float* ys = reinterpret_cast<const float*>(sb + (size_t)(TR + 2) * 8);
    const uint32_t* ids = reinterpret_cast<const uint32_t*>(sb + (size_t)(TR + 2) * 8 + (size_t)TR * 4);
    const float* xs = reinterpret_cast<const float*>(ids + a.tile_cap);
    const uint64_t row0 = tile * (uint64_t)TR;
    const int rows_here = (int)min((uint64_t)TR, a.n_rows - row0);
    const uint64_t ab = rp[0] & ~3ull;

    float mult, hj, w0 = 0.f;
    rowlane_tile<GP, Z, DAMP, COMBINE>(
        a, rp, ys, ids, xs, rows_here, tid,
        [&]() { return w0 = bias.get(use_w0, tid, it, (int)blockDim.x); }, mult, hj,
        COMBINE ? hot + (it & 1) * HOT_SLOTS : nullptr);
    // ---- bias: one damped reduction into the global w0 per tile ----
    float2* s_part = reinterpret_cast<float2*>(s_acc) + (it & 1) * 8;  // [2 slots][8 warps]
    if (use_w0) {
      const float msum = warp_sum(mult);
      const float hsum = warp_sum(hj);
      if (lane == 0) s_part[tid >> 5] = make_float2(msum, hsum);
    }
    __syncthreads();
    // this tile's hot table is complete: one reduction per parked feature, while the next
    // tile already accumulates into the other table
    if (COMBINE) {
      for (int i = tid; i < HOT_SLOTS; i += blockDim.x) hot_flush<GP>(a, hot + (it & 1) * HOT_SLOTS, i);
    }
    if (tid == ptid) {
      s_tile[stage] = nt;
      if (nt != HW_NO_TILE) issue_tile(a, smem, bars, nt, stage, policy, nt_nb, nt_ne);
      if (use_w0) {
        float M = 0.f, H = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) {
          M += s_part[i].x;
          H += s_part[i].y;
        }
        const float T = (float)rows_here;
        M += T * a.reg0 * w0;
        const float gsc = gamma_scale(fmaxf(a.w0_conc, 1.f), lr * (H / T + a.reg0));
        red_add_f(a.w0, -lr * gsc * M);
      }
    }
  }
  if (tid == ptid) sched.finish(gridDim.x, claim_raw);
}

// ---------------------------------------------------------------------------
// Warp-specialised variant: the CTA has one extra PRODUCER warp (tile claims, TMA issue,
// bias fetch and bias reduction) and the consumer warps never meet in a block barrier:
//   full[s]  : producer -> consumers, mbarrier (1 arrival + TMA transaction bytes);
//              the tile id and the bias for the tile ride along in shared memory
//   empty[s] : consumers -> producer, mbarrier (one arrival per consumer warp, issued
//              after the warp has stored its partial bias sums for the tile)
// A warp that finishes its 32 rows early starts the next staged tile at once; the
// per-tile __syncthreads of the kernel above (top stall in its ncu capture) is gone.
// The bias is read when a stage is FILLED, i.e. HW_NSTAGE tiles ahead of its use; the
// launcher widens the bias' concurrency window accordingly.
constexpr int WS_HDR = 512;  // full[3] @0, empty[3] @64, tile[3] @128, w0[3] @160, partials @192: [3][8] float2

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int GP, int Z, bool DAMP, bool COMBINE>
__global__ void __launch_bounds__(HW_MAX_THREADS + 32, 3) fm_sgd_rowlane_ws_kernel(const HogwildArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + 64);
  uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem + 128);
  float* s_w0 = reinterpret_cast<float*>(smem + 160);
  float2* s_part = reinterpret_cast<float2*>(smem + 192);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int TR = a.tile_rows;       // == number of consumer threads
  const int n_cwarps = TR >> 5;
  const bool use_w0 = a.use_w0 != 0;
  auto stage_ptr = [&](int stage) { return smem + WS_HDR + (size_t)stage * a.stage_bytes; };

  if (tid == 0) {
    for (int i = 0; i < HW_NSTAGE; i++) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, (uint32_t)n_cwarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= TR) {
    // ===================== producer warp (one lane) =====================
    if (lane != 0) return;
    TileSched sched{a.sched, a.n_tiles, false};
    const uint64_t policy = policy_evict_first();
    auto fill = [&](int stage, uint32_t t) {
      s_tile[stage] = t;
      if (t == HW_NO_TILE) {
        mbar_arrive(full + stage);  // wakes the consumers; they see NO_TILE and leave
        return;
      }
      s_w0[stage] = use_w0 ? ld_cg_f(a.w0) : 0.f;
      const uint64_t r0 = (uint64_t)t * TR, r1 = min(r0 + (uint64_t)TR, a.n_rows);
      const uint64_t nb = __ldg(a.row_ptr + r0), ne = __ldg(a.row_ptr + r1);
      const uint64_t ab = nb & ~3ull, ae = (ne + 3ull) & ~3ull;
      const uint32_t ebytes = (uint32_t)(ae - ab) * 4u;
      const uint32_t rp_bytes = (uint32_t)(TR + 2) * 8u, y_bytes = (uint32_t)TR * 4u;
      unsigned char* sb = stage_ptr(stage);
      mbar_arrive_expect_tx(full + stage, rp_bytes + y_bytes + 2u * ebytes);  // release: tile id + bias visible
      bulk_g2s_hint(sb, a.row_ptr + r0, rp_bytes, full + stage, policy);
      bulk_g2s_hint(sb + rp_bytes, a.target + r0, y_bytes, full + stage, policy);
      if (ebytes) {
        unsigned char* cb = sb + rp_bytes + y_bytes;
        bulk_g2s_hint(cb, a.col + ab, ebytes, full + stage, policy);
        bulk_g2s_hint(cb + (size_t)a.tile_cap * 4u, a.val + ab, ebytes, full + stage, policy);
      }
    };
    for (int i = 0; i < HW_NSTAGE; i++) fill(i, sched.claim());
    for (int it = 0;; ++it) {
      const int stage = it % HW_NSTAGE;
      const uint32_t t = s_tile[stage];
      if (t == HW_NO_TILE) break;  // the consumers left at this stage without arriving
      mbar_wait(empty + stage, (uint32_t)(it / HW_NSTAGE) & 1u);
      if (use_w0) {
        float M = 0.f, H = 0.f;
        for (int i = 0; i < n_cwarps; i++) {
          const float2 p = s_part[stage * 8 + i];
          M += p.x;
          H += p.y;
        }
        const uint64_t row0 = (uint64_t)t * TR;
        const float T = (float)min((uint64_t)TR, a.n_rows - row0);
        M += T * a.reg0 * s_w0[stage];
        const float gsc = gamma_scale(fmaxf(a.w0_conc, 1.f), a.lr * (H / T + a.reg0));
        red_add_f(a.w0, -a.lr * gsc * M);
      }
      fill(stage, sched.claim());
    }
    sched.finish(gridDim.x, 0u);
    return;
  }

  // ========================= consumer warps =========================
  for (int it = 0;; ++it) {
    const int stage = it % HW_NSTAGE;
    mbar_wait(full + stage, (uint32_t)(it / HW_NSTAGE) & 1u);
    const uint32_t tile = s_tile[stage];
    if (tile == HW_NO_TILE) break;
    unsigned char* sb = stage_ptr(stage);
    const uint64_t* rp = reinterpret_cast<const uint64_t*>(sb);
    const float* ys = reinterpret_cast<const float*>(sb + (size_t)(TR + 2) * 8);
    const uint32_t* ids = reinterpret_cast<const uint32_t*>(sb + (size_t)(TR + 2) * 8 + (size_t)TR * 4);
    const float* xs = reinterpret_cast<const float*>(ids + a.tile_cap);
    const uint64_t row0 = (uint64_t)tile * TR;
    const int rows_here = (int)min((uint64_t)TR, a.n_rows - row0);
    const float w0 = s_w0[stage];
    float mult, hj;
    rowlane_tile<GP, Z, DAMP, COMBINE>(a, rp, ys, ids, xs, rows_here, tid, [&]() { return w0; }, mult, hj);
    if (use_w0) {
      const float msum = warp_sum(mult);
      const float hsum = warp_sum(hj);
      if (lane == 0) s_part[stage * 8 + (tid >> 5)] = make_float2(msum, hsum);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);  // release: partials + "done reading the stage"
  }
}

template <int GP, int Z>
static HogwildKernelFn pick_d(bool damp, bool combine) {
  if (combine)
    return damp ? fm_sgd_rowlane_kernel<GP, Z, true, true> : fm_sgd_rowlane_kernel<GP, Z, false, true>;
  return damp ? fm_sgd_rowlane_kernel<GP, Z, true, false> : fm_sgd_rowlane_kernel<GP, Z, false, false>;
}

template <int GP>
static HogwildKernelFn pick_z(int z, bool damp, bool combine) {
  if (z <= 1) return pick_d<GP, 1>(damp, combine);
  if (z <= 2) return pick_d<GP, 2>(damp, combine);
  if (z <= 4) return pick_d<GP, 4>(damp, combine);
  return nullptr;
}

template <int GP, int Z>
static HogwildKernelFn pick_d_ws(bool damp, bool combine) {
  if (combine)
    return damp ? fm_sgd_rowlane_ws_kernel<GP, Z, true, true> : fm_sgd_rowlane_ws_kernel<GP, Z, false, true>;
  return damp ? fm_sgd_rowlane_ws_kernel<GP, Z, true, false> : fm_sgd_rowlane_ws_kernel<GP, Z, false, false>;
}

template <int GP>
static HogwildKernelFn pick_z_ws(int z, bool damp, bool combine) {
  if (z <= 1) return pick_d_ws<GP, 1>(damp, combine);
  if (z <= 2) return pick_d_ws<GP, 2>(damp, combine);
  if (z <= 4) return pick_d_ws<GP, 4>(damp, combine);
  return nullptr;
}

HogwildKernelFn pick_rowlane_ws_kernel(int gp, int max_row_nnz, bool damp, bool combine) {
  if (gp == 1) return pick_z_ws<1>(max_row_nnz, damp, combine);
  if (gp == 2) return pick_z_ws<2>(max_row_nnz, damp, combine);
  return nullptr;
}

HogwildKernelFn pick_rowlane_kernel(int gp, int max_row_nnz, bool damp, bool combine) {
  if (gp == 1) return pick_z<1>(max_row_nnz, damp, combine);
  if (gp == 2) return pick_z<2>(max_row_nnz, damp, combine);
  return nullptr;
}

}  // namespace fmb
