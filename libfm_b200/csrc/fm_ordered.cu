// fm_ordered.cu -- launcher of the sequentially consistent epoch (fm_ordered.cuh) and the
// index work it needs per uploaded data set: for every entry the distance to the previous
// entry naming the same feature, for every row the distance to the nearest earlier row it
// depends on.  Both are pure functions of col[] / row_ptr[] (bit-exact ordering work) and
// are built once per upload, on the device, the first time an ORDERED epoch needs them.
#include <algorithm>
#include <cstdio>

#include <cub/device/device_radix_sort.cuh>

#include "fm_ordered.cuh"
#include "fmb200_internal.h"

namespace fmb {

namespace {

__global__ void ord_iota_kernel(uint32_t* p, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = (uint32_t)i;
}

// row containing entry e: the last r with row_ptr[r] <= e (empty rows skipped by construction)
__device__ __forceinline__ uint64_t row_of(const uint64_t* __restrict__ rp, uint64_t n_rows, uint64_t e) {
  uint64_t lo = 0, hi = n_rows;  // invariant: rp[lo] <= e < rp[hi]
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (rp[mid] <= e) lo = mid;
    else hi = mid;
  }
  return lo;
}

// `ids` / `ent`: the entries stably sorted by feature id, i.e. each feature's occurrences in
// file order.  Neighbours with equal id are consecutive occurrences of one feature.
__global__ void ord_link_kernel(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ ent,
                                uint64_t nnz, const uint64_t* __restrict__ rp, uint64_t n_rows,
                                uint32_t* __restrict__ link, uint32_t* __restrict__ rowdep) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nnz;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t e = ent[i];
    uint32_t L = ORD_NONE;
    if (i > 0 && ids[i - 1] == ids[i]) {
      const uint32_t pe = ent[i - 1];
      L = e - pe;
      const uint64_t r = row_of(rp, n_rows, e), pr = row_of(rp, n_rows, pe);
      atomicMin(rowdep + r, (uint32_t)(r - pr));
    }
    link[e] = L;
  }
}

// shape[0]: bit 0 = every value is exactly 1, bit 1 = every row has exactly z entries (the one-hot two-field
// shape of ratings data); the epoch kernel reads the word and takes its one-hot path when both hold
__global__ void ord_shape_kernel(const float* __restrict__ val, uint64_t nnz, const uint64_t* __restrict__ rp,
                                 uint64_t n_rows, uint32_t z, uint32_t* shape) {
  uint32_t clear = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * blockDim.x)
    if (val[i] != 1.0f) clear |= 1u;
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (uint64_t)gridDim.x * blockDim.x)
    if (rp[r + 1] - rp[r] != z) clear |= 2u;
  if (clear) atomicAnd(shape, ~clear);
}

// blockDim <= ORD_SMAX * GL (one group of GL lanes per example of a run): small k leaves the
// register file to few threads (k <= 8: 128 threads)
template <int GL, int KF, int TASK, int ZF = 0>
__global__ void __launch_bounds__((ORD_SMAX * GL < 256 ? 256 : (ORD_SMAX * GL < ORD_MAX_THREADS ? ORD_SMAX * GL : ORD_MAX_THREADS)), 1)
    fm_sgd_ordered_kernel(const OrderedArgs a) {
  extern __shared__ __align__(128) unsigned char ord_smem[];
  ordered_epoch_body<GL, KF, TASK, ZF>(a, ord_smem);
}

// warp-specialised form: ORD_SMAX * GL compute threads, ORD_PARKED threads that leave after the set-up and
// ORD_HELPERS helper threads (write-back, fetch).  ONE helper warp: every further helper warp slowed the epoch
// by ~5% although the helpers idle most of the time -- their bursts of shared-memory / LSU traffic delay the
// compute warps' loads (r02 calls M, N, the same build on one box, C2-shaped 200 000 rows: 1 / 2 / 3 / 4 helper
// warps = 6.30 / 6.64-6.74 / 6.92 / 7.33 ms; which schedulers the helpers sit on matters little).  The parked
// threads keep the helper warp's index a multiple of 4 apart from compute warp 3.
constexpr int ORD_PARKED = 96;
constexpr int ORD_HELPERS = 32;
template <int GL, int KF, int TASK, int ZF = 0>
__global__ void __launch_bounds__(ORD_SMAX * GL + ORD_PARKED + ORD_HELPERS, 1)
    fm_sgd_ordered_ws_kernel(const OrderedArgs a) {
  extern __shared__ __align__(128) unsigned char ord_smem[];
  ordered_epoch_body_ws<GL, KF, TASK, ZF>(a, ord_smem, ORD_SMAX * GL, ORD_PARKED);
}

using OrdFn = void (*)(const OrderedArgs);

// (GL lanes per example, KF consecutive factors per lane): KF <= 8; k <= 8 runs one lane per example
inline void ordered_shape(int k, int* GL, int* KF) {
  if (k <= 8) {
    *GL = 1;
    *KF = k <= 1 ? 1 : (k <= 2 ? 2 : (k <= 4 ? 4 : 8));
    return;
  }
  *KF = 8;
  int g = 2;
  while (g * 8 < k) g <<= 1;
  *GL = g;
}

// register-resident fast path: k in {2,4,8} exactly, rows of at most 2 / 4 entries
template <int TASK>
OrdFn pick_fast_kernel(int k, uint32_t max_row_nnz) {
  if (max_row_nnz < 1 || max_row_nnz > 4) return nullptr;
  const bool z2 = max_row_nnz <= 2;
  if (k == 2) return z2 ? fm_sgd_ordered_kernel<1, 2, TASK, 2> : fm_sgd_ordered_kernel<1, 2, TASK, 4>;
  if (k == 4) return z2 ? fm_sgd_ordered_kernel<1, 4, TASK, 2> : fm_sgd_ordered_kernel<1, 4, TASK, 4>;
  if (k == 8) return z2 ? fm_sgd_ordered_kernel<1, 8, TASK, 2> : fm_sgd_ordered_kernel<1, 8, TASK, 4>;
  return nullptr;
}

// warp-specialised kernels: k <= 32 (GL <= 4: 128 GL compute threads + 128 helpers fit one CTA)
template <int TASK>
OrdFn pick_ws_kernel(int k, uint32_t max_row_nnz, int* ncompute) {
  *ncompute = ORD_SMAX;
  if (max_row_nnz >= 1 && max_row_nnz <= 4 && (k == 2 || k == 4 || k == 8)) {
    const bool z2 = max_row_nnz <= 2;
    if (k == 2) return z2 ? fm_sgd_ordered_ws_kernel<1, 2, TASK, 2> : fm_sgd_ordered_ws_kernel<1, 2, TASK, 4>;
    if (k == 4) return z2 ? fm_sgd_ordered_ws_kernel<1, 4, TASK, 2> : fm_sgd_ordered_ws_kernel<1, 4, TASK, 4>;
    return z2 ? fm_sgd_ordered_ws_kernel<1, 8, TASK, 2> : fm_sgd_ordered_ws_kernel<1, 8, TASK, 4>;
  }
  if (k <= 1) return fm_sgd_ordered_ws_kernel<1, 1, TASK>;
  if (k <= 2) return fm_sgd_ordered_ws_kernel<1, 2, TASK>;
  if (k <= 4) return fm_sgd_ordered_ws_kernel<1, 4, TASK>;
  if (k <= 8) return fm_sgd_ordered_ws_kernel<1, 8, TASK>;
  if (k <= 16) {
    *ncompute = 2 * ORD_SMAX;
    return fm_sgd_ordered_ws_kernel<2, 8, TASK>;
  }
  if (k <= 32) {
    *ncompute = 4 * ORD_SMAX;
    return fm_sgd_ordered_ws_kernel<4, 8, TASK>;
  }
  return nullptr;
}

template <int TASK>
OrdFn pick_kernel(int k) {
  if (k <= 1) return fm_sgd_ordered_kernel<1, 1, TASK>;
  if (k <= 2) return fm_sgd_ordered_kernel<1, 2, TASK>;
  if (k <= 4) return fm_sgd_ordered_kernel<1, 4, TASK>;
  if (k <= 8) return fm_sgd_ordered_kernel<1, 8, TASK>;
  if (k <= 16) return fm_sgd_ordered_kernel<2, 8, TASK>;
  if (k <= 32) return fm_sgd_ordered_kernel<4, 8, TASK>;
  if (k <= 64) return fm_sgd_ordered_kernel<8, 8, TASK>;
  if (k <= 128) return fm_sgd_ordered_kernel<16, 8, TASK>;
  return fm_sgd_ordered_kernel<32, 8, TASK>;
}

int grid_for(const fmb200_ctx* c, uint64_t work) {
  const uint64_t blocks = (work + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 8;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

// Build link[] / rowdep[] of a data set on c->stream (no host sync).
cudaError_t build_ordered_links(fmb200_ctx* c, DataSlot& d) {
  if (d.links_ready) return cudaSuccess;
  if (d.nnz >= 0xffffffffull) return cudaErrorInvalidValue;
  cudaError_t e;
  const uint64_t cap_e = d.cap_nnz + 16, cap_r = d.cap_rows + 520;
  if (!d.link) {
    if ((e = cudaMalloc(&d.link, cap_e * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.rowdep, (cap_r + 4) * sizeof(uint32_t))) != cudaSuccess) return e;  // + the shape word
  }
  d.ord_shape = d.rowdep + cap_r;
  {
    if ((e = cudaMemsetAsync(d.ord_shape, 0x03, sizeof(uint32_t), c->stream)) != cudaSuccess) return e;
    ord_shape_kernel<<<grid_for(c, d.nnz + d.n_rows), 256, 0, c->stream>>>(d.val, d.nnz, d.row_ptr, d.n_rows,
                                                                           d.max_row_nnz, d.ord_shape);
    c->launches++;
  }
  if ((e = cudaMemsetAsync(d.link, 0xff, cap_e * sizeof(uint32_t), c->stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(d.rowdep, 0xff, cap_r * sizeof(uint32_t), c->stream)) != cudaSuccess) return e;
  if (d.nnz > 0) {
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (uint64_t)c->n) bits++;
    size_t tmp_bytes = 0;
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d.col, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (uint64_t)d.nnz, 0, bits, c->stream);
    if (e != cudaSuccess) return e;
    // scratch of the index build: [ids | ent_in | ent | sort temp].  Kept with the slot for data sets up to
    // 64 M entries (a re-upload then rebuilds its index without a single allocation or host sync -- the
    // end-to-end path uploads a fresh data set every step); larger ones release it right away.
    const size_t words = ((size_t)d.nnz + 63) & ~(size_t)63;
    const size_t need = 3 * words * sizeof(uint32_t) + ((tmp_bytes + 255) & ~(size_t)255) + 256;
    if (d.ord_scratch_bytes < need) {
      if (d.ord_scratch) cudaFree(d.ord_scratch);
      d.ord_scratch = nullptr;
      d.ord_scratch_bytes = 0;
      if ((e = cudaMalloc(&d.ord_scratch, need)) != cudaSuccess) return e;
      d.ord_scratch_bytes = need;
    }
    uint32_t* ids = static_cast<uint32_t*>(d.ord_scratch);
    uint32_t* ent_in = ids + words;
    uint32_t* ent = ent_in + words;
    void* tmp = ent + words;
    ord_iota_kernel<<<grid_for(c, d.nnz), 256, 0, c->stream>>>(ent_in, d.nnz);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d.col, ids, ent_in, ent, (uint64_t)d.nnz, 0, bits, c->stream);
    if (e != cudaSuccess) return e;
    ord_link_kernel<<<grid_for(c, d.nnz), 256, 0, c->stream>>>(ids, ent, d.nnz, d.row_ptr, d.n_rows, d.link, d.rowdep);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    c->launches += 3;  // iota, link + the library's sort passes counted as one
    if (d.nnz > (64ull << 20)) {
      if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
      cudaFree(d.ord_scratch);
      d.ord_scratch = nullptr;
      d.ord_scratch_bytes = 0;
    }
  }
  d.links_ready = true;
  return cudaSuccess;
}

// Pick the tile geometry; false if not even one row fits the ring (caller falls back to the
// row-at-a-time kernel, which is also sequentially consistent).
bool ordered_geometry(const fmb200_ctx* c, const DataSlot& d, int* TR_out, uint32_t* TE_out, int* kw_out,
                      int* rs_out, size_t* smem_out) {
  const int k = c->k;
  const int kw = (k & 1) ? k + 1 : k;
  const int rs = kw + 2;
  const size_t limit = (size_t)c->max_smem_optin;
  for (int TR = 256; TR >= 1; TR >>= 1) {
    uint64_t span;
    if (TR >= 32) {
      int idx = 0;
      while ((32 << idx) < TR) idx++;
      span = d.tile_span[idx];
    } else {
      span = (uint64_t)TR * d.max_row_nnz + 6;
      if (span > d.tile_span[0]) span = d.tile_span[0];
    }
    const uint64_t TE64 = ((span + 3) & ~3ull) + 4;
    if (TE64 > (1u << 20)) continue;
    const uint32_t TE = (uint32_t)TE64;
    const size_t need = ord_smem_bytes(TR, TE, rs);
    if (need <= limit) {
      *TR_out = TR;
      *TE_out = TE;
      *kw_out = kw;
      *rs_out = rs;
      *smem_out = need;
      return true;
    }
  }
  return false;
}

cudaError_t launch_sgd_ordered(fmb200_ctx* c, DataSlot& d, bool* handled) {
  *handled = false;
  if (d.n_rows == 0) {
    *handled = true;
    return cudaSuccess;
  }
  if (d.nnz >= 0xffffffffull) return cudaSuccess;  // 32-bit entry distances: not eligible
  int TR = 0, kw = 0, rs = 0;
  uint32_t TE = 0;
  size_t smem = 0;
  if (!ordered_geometry(c, d, &TR, &TE, &kw, &rs, &smem)) return cudaSuccess;
  cudaError_t e = build_ordered_links(c, d);
  if (e != cudaSuccess) return e;

  OrderedArgs a{};
  a.row_ptr = d.row_ptr;
  a.col = d.col;
  a.val = d.val;
  a.target = d.target;
  a.link = d.link;
  a.rowdep = d.rowdep;
  a.shape = d.ord_shape;
  a.n_rows = d.n_rows;
  a.n_tiles = (uint32_t)((d.n_rows + TR - 1) / TR);
  a.tile_rows = TR;
  a.tile_cap = TE;
  a.w0 = c->p64.w0();
  a.w = c->p64.w();
  a.v = c->p64.v();
  a.k = c->k;
  a.kw = kw;
  a.rs = rs;
  a.use_w0 = c->k0;
  a.use_w = c->k1;
  a.lr = c->hp.lr;
  a.reg0 = c->hp.reg0;
  a.regw = c->hp.regw;
  a.regv = c->hp.regv;
  a.min_target = c->hp.min_target;
  a.max_target = c->hp.max_target;
  a.csr_bytes = ord_csr_bytes(TR, TE);
  a.rec_bytes = TE * (uint32_t)rs * 8u;
  a.debug = (c->tune_variant >= 100) ? c->tune_variant - 100 : 0;  // timing experiments (wrong results)
  a.prof = nullptr;
  const bool want_prof = (a.debug & 32) != 0;  // variant 132 (+ skip bits): print the phase timers
  a.debug &= 31;
  if (want_prof) {
    if ((e = cudaMalloc(&a.prof, 16 * sizeof(unsigned long long))) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(a.prof, 0, 16 * sizeof(unsigned long long), c->stream)) != cudaSuccess) return e;
  }

  int GL = 1, KF = 1;
  ordered_shape(c->k, &GL, &KF);
  // one group of GL lanes per example of a run: min(ORD_SMAX, 1024 / GL) examples
  int threads = std::min(ORD_SMAX * GL, ORD_MAX_THREADS);
  const int bound = std::max(threads, 256);  // the kernel's launch bound
  if (c->tune_threads)  // fewer threads = shorter runs; more = helper warps for the fetch issue / write-back
    threads = std::min(bound, std::max(32, (c->tune_threads / (32 > GL ? 32 : GL)) * (32 > GL ? 32 : GL)));
  OrdFn fn = c->hp.task == FMB200_TASK_REGRESSION ? pick_kernel<0>(c->k) : pick_kernel<1>(c->k);
  if (c->tune_variant != 1) {  // variant 1 forces the general path
    OrdFn fast = c->hp.task == FMB200_TASK_REGRESSION ? pick_fast_kernel<0>(c->k, d.max_row_nnz)
                                                      : pick_fast_kernel<1>(c->k, d.max_row_nnz);
    if (fast != nullptr) fn = fast;
  }
  // the warp-specialised form is the default where it exists (variant 1 / 2 and explicit thread counts keep
  // the single-role kernels: comparisons, tests)
  if (c->tune_variant != 1 && c->tune_variant != 2 && !c->tune_threads) {
    int ncompute = 0;
    OrdFn ws = c->hp.task == FMB200_TASK_REGRESSION ? pick_ws_kernel<0>(c->k, d.max_row_nnz, &ncompute)
                                                    : pick_ws_kernel<1>(c->k, d.max_row_nnz, &ncompute);
    if (ws != nullptr) {
      fn = ws;
      threads = ncompute + ORD_PARKED + ORD_HELPERS;
    }
  }
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fn<<<1, threads, smem, c->stream>>>(a);
  c->launches++;
  c->last_cfg = EpochConfig{GL, std::min(ORD_SMAX, (threads >= ORD_SMAX * GL ? ORD_SMAX * GL : threads) / GL), TR, 1, threads, (int)smem, 0};
  *handled = true;
  if (want_prof) {
    unsigned long long h[16];
    if ((e = cudaMemcpyAsync(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
    cudaFree(a.prof);
    static const char* name[16] = {"A scores", "A barrier", "chain", "chain barrier", "check+sgd", "closing barrier",
                                   "tile prologue", "wait for helpers", "H write-back", "H csr", "H fetch issue",
                                   "H fetch landing", "H wait for compute", "-", "-", "between"};
    fprintf(stderr, "[ordered phases, cycles per tile (%u tiles)]", a.n_tiles);
    for (int i = 0; i < 16; i++)
      if (h[i]) fprintf(stderr, " %s=%.0f", name[i], (double)h[i] / a.n_tiles);
    fprintf(stderr, "\n");
  }
  return cudaGetLastError();
}

}  // namespace fmb
