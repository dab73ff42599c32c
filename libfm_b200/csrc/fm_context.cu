// fm_context.cu -- the C ABI declared in include/fmb200.h: context lifetime,
// host<->HBM layout conversion (bit-exact index/ordering work), epoch /
// evaluate / predict entry points.  CUDA only: no CPU fallback exists.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "fmb200_internal.h"

using namespace fmb;

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CK(expr)                                                                      \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                          \
  } while (0)

#define NEED_CTX(c) \
  if ((c) == nullptr) return fail("null context")

// No C++ exception may cross the C ABI: entry points that allocate host staging run
// their body through this.
template <class F>
int guarded(F&& body) {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    return fail("out of host memory");
  } catch (...) {
    return fail("unexpected C++ exception");
  }
}

int bind(fmb200_ctx* c) {
  CK(cudaSetDevice(c->device));
  return 0;
}

void free_slot(DataSlot& s) {
  if (s.row_ptr) cudaFree(s.row_ptr);
  if (s.col) cudaFree(s.col);
  if (s.val) cudaFree(s.val);
  if (s.target) cudaFree(s.target);
  if (s.feat_cnt) cudaFree(s.feat_cnt);
  if (s.link) cudaFree(s.link);
  if (s.rowdep) cudaFree(s.rowdep);
  if (s.ord_scratch) cudaFree(s.ord_scratch);
  if (s.d_flag) cudaFree(s.d_flag);
  if (s.h_flag) cudaFreeHost(s.h_flag);
  if (s.ready) cudaEventDestroy(s.ready);
  s = DataSlot();
}

// slack (in elements) behind every CSR array so that whole-tile TMA bulk copies
// of the last tile stay inside the allocation
constexpr uint64_t kRowSlack = 512 + 8;
constexpr uint64_t kEntrySlack = 16;

// Make `slot` ready to receive a data set of (n_rows, nnz): drain an earlier asynchronous
// upload, (re)allocate, reset the bookkeeping.  Enqueues only the memsets of fresh buffers.
int upload_begin(fmb200_ctx* c, int slot, uint64_t n_rows, uint64_t nnz, cudaStream_t st) {
  DataSlot& s = c->slots[slot];
  if (s.pending) {  // an earlier asynchronous upload into this slot: drain it first
    CK(cudaEventSynchronize(s.ready));
    s.pending = false;
  }
  // re-uploads into a slot reuse its buffers when they are large enough
  if (!(s.row_ptr && s.cap_rows >= n_rows && s.cap_nnz >= nnz)) {
    free_slot(s);
    CK(cudaMalloc(&s.row_ptr, (n_rows + 1 + kRowSlack) * sizeof(uint64_t)));
    CK(cudaMalloc(&s.target, (n_rows + kRowSlack) * sizeof(float)));
    CK(cudaMalloc(&s.col, (nnz + kEntrySlack) * sizeof(uint32_t)));
    CK(cudaMalloc(&s.val, (nnz + kEntrySlack) * sizeof(float)));
    CK(cudaMalloc(&s.feat_cnt, sizeof(float) * (size_t)(c->n ? c->n : 1)));
    CK(cudaMalloc(&s.d_flag, 16 * sizeof(unsigned int)));
    CK(cudaHostAlloc((void**)&s.h_flag, 16 * sizeof(unsigned int), cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
    // the slack is only ever read by whole-tile bulk copies and never used
    CK(cudaMemsetAsync(s.row_ptr, 0, (n_rows + 1 + kRowSlack) * sizeof(uint64_t), st));
    CK(cudaMemsetAsync(s.target, 0, (n_rows + kRowSlack) * sizeof(float), st));
    CK(cudaMemsetAsync(s.col, 0, (nnz + kEntrySlack) * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(s.val, 0, (nnz + kEntrySlack) * sizeof(float), st));
    s.cap_rows = n_rows;
    s.cap_nnz = nnz;
  }
  s.present = false;
  s.links_ready = false;
  s.upload_gen = ++c->upload_counter;
  s.n_rows = n_rows;
  s.nnz = nnz;
  return 0;
}

// Inspection runs on the device: offsets monotone and consistent, longest row, tile
// spans, largest column id (the reference asserts id < num_attribute per access,
// fm_model.h:112), and the per-feature occurrence counts used by the HOGWILD damping.
// Leaves the results in the slot's pinned flag mirror; upload_finish() collects them.
int upload_inspect(fmb200_ctx* c, int slot, cudaStream_t st) {
  DataSlot& s = c->slots[slot];
  cudaStream_t saved = c->stream;
  c->stream = st;  // the launch helpers enqueue on c->stream
  cudaError_t e1 = cudaMemsetAsync(s.d_flag, 0, 16 * sizeof(unsigned int), st);
  cudaError_t e2 = launch_csr_inspect(c, s.row_ptr, s.n_rows, s.nnz, s.d_flag);
  cudaError_t e3 = launch_feature_counts(c, s.col, s.nnz, s.feat_cnt, s.d_flag + 8, s.d_flag + 9);
  c->stream = saved;
  CK(e1);
  CK(e2);
  CK(e3);
  CK(cudaMemcpyAsync(s.h_flag, s.d_flag, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(s.ready, st));
  s.pending = true;
  return 0;
}

// Enqueue the copies + the device-side inspection of one SoA data set on `st`.
int upload_enqueue(fmb200_ctx* c, int slot, uint64_t n_rows, uint64_t nnz, const uint64_t* row_ptr,
                   const uint32_t* col, const float* val, const float* target, cudaStream_t st) {
  if (upload_begin(c, slot, n_rows, nnz, st)) return 1;
  DataSlot& s = c->slots[slot];
  CK(cudaMemcpyAsync(s.row_ptr, row_ptr, (n_rows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.target, target, n_rows * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.col, col, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.val, val, nnz * sizeof(float), cudaMemcpyHostToDevice, st));
  return upload_inspect(c, slot, st);
}

// One-hot rows of fixed width z: ids [n_rows*z] and targets cross PCIe; row offsets and the
// all-ones values are written by a kernel (fm_upload.cu).
int upload_onehot_enqueue(fmb200_ctx* c, int slot, uint64_t n_rows, uint32_t z, const uint32_t* ids,
                          const float* target, cudaStream_t st) {
  const uint64_t nnz = n_rows * z;
  if (upload_begin(c, slot, n_rows, nnz, st)) return 1;
  DataSlot& s = c->slots[slot];
  CK(cudaMemcpyAsync(s.target, target, n_rows * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.col, ids, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  cudaStream_t saved = c->stream;
  c->stream = st;
  cudaError_t e = launch_onehot_fill(c, n_rows, z, s.row_ptr, s.val);
  c->stream = saved;
  CK(e);
  return upload_inspect(c, slot, st);
}

// RAII for the temporaries of the AoS upload
struct DevTmp {
  void* p = nullptr;
  ~DevTmp() {
    if (p) cudaFree(p);
  }
};

// The one host sync of an upload: wait for the slot's event and read the verdict.
int upload_finish(fmb200_ctx* c, int slot) {
  DataSlot& s = c->slots[slot];
  if (!s.pending) return 0;
  CK(cudaEventSynchronize(s.ready));
  s.pending = false;
  const unsigned int* h = s.h_flag;
  if (h[0] & 1u) return fail("row_ptr[0] must be 0");
  if (h[0] & 2u) return fail("row_ptr is not monotone");
  if (h[0] & 4u) return fail("row_ptr[n_rows] != nnz");
  if (h[0] & 8u) return fail("a row is longer than 2^32-1 entries");
  if (s.nnz > 0 && h[8] >= c->n)
    return fail("feature id %u out of range (num_attribute=%u)", h[8], c->n);
  s.max_row_nnz = h[1];
  for (int i = 0; i < 5; i++) s.tile_span[i] = h[2 + i];
  s.max_feat_cnt = h[9];
  s.present = true;
  return 0;
}

int upload_common(fmb200_ctx* c, int slot, uint64_t n_rows, uint64_t nnz, const uint64_t* row_ptr,
                  const uint32_t* col, const float* val, const float* target) {
  if (upload_enqueue(c, slot, n_rows, nnz, row_ptr, col, val, target, c->stream)) return 1;
  return upload_finish(c, slot);
}

int need_slot(fmb200_ctx* c, int slot) {
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (c->slots[slot].pending && upload_finish(c, slot)) return 1;
  if (!c->slots[slot].present) return fail("slot %d holds no data", slot);
  return 0;
}

int ensure_partials(fmb200_ctx* c, int n_blocks) {
  if (c->n_partials < n_blocks) {
    if (c->d_partials) cudaFree(c->d_partials);
    c->d_partials = nullptr;
    CK(cudaMalloc(&c->d_partials, sizeof(double) * 3 * n_blocks));
    c->n_partials = n_blocks;
  }
  return 0;
}

int metric_blocks(fmb200_ctx* c, const DataSlot& s) {
  uint64_t want = (s.n_rows + 255) / 256;
  uint64_t cap = (uint64_t)c->sm_count * 8;
  uint64_t b = want < cap ? want : cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

// Everything of fmb200_create that can fail after the context object exists; the caller
// destroys the partially built context on a non-zero return.
static int create_resources(fmb200_ctx* c, int device, const cudaDeviceProp& prop, uint32_t n_attr,
                            int num_factor, int use_w0, int use_w) {
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->n = n_attr;
  c->k = num_factor;
  c->kp = (num_factor + 3) & ~3;
  c->k0 = use_w0 != 0;
  c->k1 = use_w != 0;
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  c->p32.ws = (n_attr <= 131072u) ? 8 : 1;
  const uint64_t n4 = ((uint64_t)n_attr * c->p32.ws + 3) & ~3ull;
  c->p32.off_w = 4;
  c->p32.off_v = 4 + n4;
  c->p32.n_floats = 4 + n4 + (uint64_t)n_attr * c->kp;
  c->p64.off_v = Params64::off_w + (((uint64_t)n_attr + 1) & ~1ull);
  c->p64.n_doubles = c->p64.off_v + (uint64_t)n_attr * num_factor + 2;
  c->comm_buf_bytes = (c->p32.n_floats * sizeof(float) + 255) & ~(size_t)255;
  c->comm_cnt_floats = ((size_t)n_attr + 63) & ~(size_t)63;
  const size_t comm_total = c->comm_hdr + 3 * c->comm_buf_bytes + (3 * c->comm_cnt_floats + 2 * (size_t)fmb::FMB_PEER_PART) * sizeof(float);
  CK(cudaMalloc(&c->comm_base, comm_total));
  CK(cudaMemsetAsync(c->comm_base, 0, comm_total, c->stream));
  c->p32.base = reinterpret_cast<float*>(c->comm_base + c->comm_hdr);
  c->peer_base[0] = c->comm_base;
  CK(cudaMalloc(&c->p64.base, c->p64.n_doubles * sizeof(double)));
  CK(cudaMemsetAsync(c->p32.base, 0, c->p32.n_floats * sizeof(float), c->stream));
  CK(cudaMemsetAsync(c->p64.base, 0, c->p64.n_doubles * sizeof(double), c->stream));
  CK(cudaMalloc(&c->d_sched, 2 * sizeof(unsigned int)));
  CK(cudaMemsetAsync(c->d_sched, 0, 2 * sizeof(unsigned int), c->stream));
  CK(cudaMalloc(&c->d_flag, 16 * sizeof(unsigned int)));
  CK(cudaHostAlloc((void**)&c->h_flag, 16 * sizeof(unsigned int), cudaHostAllocDefault));
  {
    const size_t need = std::max(c->p32.n_floats * sizeof(float), c->p64.n_doubles * sizeof(double));
    if (need <= (64u << 20)) {
      CK(cudaHostAlloc(&c->h_stage, need, cudaHostAllocDefault));
      c->h_stage_bytes = need;
    }
  }
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" {

const char* fmb200_last_error(void) { return g_err; }

int fmb200_create(fmb200_ctx** out, int device, uint32_t n_attr, int num_factor, int use_w0,
                  int use_w) {
  if (out == nullptr) return fail("null out pointer");
  *out = nullptr;
  if (num_factor < 0) return fail("num_factor must be >= 0");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("no CUDA device available (%s): libfmb200 has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail("device %d out of range (count %d)", device, count);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                prop.minor);
  fmb200_ctx* c = new (std::nothrow) fmb200_ctx();
  if (!c) return fail("out of host memory");
  if (create_resources(c, device, prop, n_attr, num_factor, use_w0, use_w)) {
    fmb200_destroy(c);  // releases whatever was allocated; g_err keeps the cause
    return 1;
  }
  *out = c;
  return 0;
}

void fmb200_destroy(fmb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 0; i < FMB200_MAX_SLOTS; i++) free_slot(c->slots[i]);
  for (int q = 0; q < FMB200_MAX_PEERS; q++)
    if (c->peer_ipc[q] && c->peer_base[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
  if (c->comm_base) cudaFree(c->comm_base);
  if (c->p64.base) cudaFree(c->p64.base);
  if (c->sgda_grad_w) cudaFree(c->sgda_grad_w);
  if (c->sgda_grad_v) cudaFree(c->sgda_grad_v);
  if (c->sgda_reg_w) cudaFree(c->sgda_reg_w);
  if (c->sgda_reg_v) cudaFree(c->sgda_reg_v);
  if (c->sgda_group) cudaFree(c->sgda_group);
  if (c->d_partials) cudaFree(c->d_partials);
  if (c->d_pred) cudaFree(c->d_pred);
  if (c->d_sched) cudaFree(c->d_sched);
  if (c->d_flag) cudaFree(c->d_flag);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int fmb200_set_hparams(fmb200_ctx* c, int task, double learn_rate, double reg0, double regw,
                       double regv, double min_target, double max_target) {
  NEED_CTX(c);
  if (task != FMB200_TASK_REGRESSION && task != FMB200_TASK_CLASSIFICATION)
    return fail("unknown task");  // fm_learn.h:99-101 throws "unknown task"
  c->hp.task = task;
  c->hp.lr = learn_rate;
  c->hp.reg0 = reg0;
  c->hp.regw = regw;
  c->hp.regv = regv;
  c->hp.min_target = min_target;
  c->hp.max_target = max_target;
  return 0;
}

int fmb200_set_mode(fmb200_ctx* c, int mode) {
  NEED_CTX(c);
  if (mode != FMB200_MODE_INORDER && mode != FMB200_MODE_HOGWILD && mode != FMB200_MODE_ORDERED)
    return fail("unknown mode %d", mode);
  if (mode == c->mode) return 0;
  if (bind(c)) return 1;
  // carry the live state into the representation of the new mode (INORDER and ORDERED share
  // the fp64 state)
  const bool was64 = c->mode != FMB200_MODE_HOGWILD, is64 = mode != FMB200_MODE_HOGWILD;
  if (is64 && c->k > 256) return fail("num_factor > 256 is not supported in the fp64 modes");
  if (was64 && !is64) {
    CK(launch_p64_to_p32(c));
  } else if (!was64 && is64) {
    CK(launch_p32_to_p64(c));
  }
  CK(cudaStreamSynchronize(c->stream));
  c->mode = mode;
  c->peer_base_valid = false;
  return 0;
}

int fmb200_upload_data(fmb200_ctx* c, int slot, uint64_t n_rows, uint64_t nnz,
                       const uint64_t* row_ptr, const uint32_t* col, const float* val,
                       const float* target) {
  NEED_CTX(c);
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (!row_ptr || (n_rows && !target) || (nnz && (!col || !val))) return fail("null data pointer");
  if (n_rows > 0xffffffffull) return fail("row count exceeds the reference's uint range");
  if (bind(c)) return 1;
  return upload_common(c, slot, n_rows, nnz, row_ptr, col, val, target);
}

int fmb200_upload_data_async(fmb200_ctx* c, int slot, uint64_t n_rows, uint64_t nnz,
                             const uint64_t* row_ptr, const uint32_t* col, const float* val,
                             const float* target) {
  NEED_CTX(c);
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (!row_ptr || (n_rows && !target) || (nnz && (!col || !val))) return fail("null data pointer");
  if (n_rows > 0xffffffffull) return fail("row count exceeds the reference's uint range");
  if (bind(c)) return 1;
  if (c->copy_stream == nullptr) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  return upload_enqueue(c, slot, n_rows, nnz, row_ptr, col, val, target, c->copy_stream);
}

// reference layout, util/fmatrix.h:34-42
struct AosEntry {
  uint32_t id;
  float value;
};
struct AosRow {
  const AosEntry* data;
  uint32_t size;
};
static_assert(sizeof(AosRow) == 16 && sizeof(AosEntry) == 8, "LP64 layout of sparse_row/sparse_entry");

// rows scattered over the heap: gather them on the host (slow path)
static int upload_aos_host_gather(fmb200_ctx* c, int slot, uint64_t n_rows, const AosRow* r,
                                  const float* target) {
  return guarded([&]() -> int {
    std::vector<uint64_t> rp(n_rows + 1);
    rp[0] = 0;
    for (uint64_t i = 0; i < n_rows; i++) rp[i + 1] = rp[i] + r[i].size;
    const uint64_t nnz = rp[n_rows];
    std::vector<uint32_t> col(nnz ? nnz : 1);
    std::vector<float> val(nnz ? nnz : 1);
    for (uint64_t i = 0; i < n_rows; i++) {
      const AosEntry* e = r[i].data;
      uint64_t o = rp[i];
      for (uint32_t j = 0; j < r[i].size; j++) {
        col[o + j] = e[j].id;
        val[o + j] = e[j].value;
      }
    }
    return upload_common(c, slot, n_rows, nnz, rp.data(), col.data(), val.data(), target);
  });
}

int fmb200_upload_data_aos(fmb200_ctx* c, int slot, uint64_t n_rows, const void* rows,
                           const float* target) {
  NEED_CTX(c);
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (n_rows && (!rows || !target)) return fail("null data pointer");
  if (n_rows > 0xffffffffull) return fail("row count exceeds the reference's uint range");
  if (bind(c)) return 1;
  const AosRow* r = static_cast<const AosRow*>(rows);
  // The reference keeps all entries in ONE block (Data.h:238,260).  The row array crosses PCIe
  // as it is; the device scans the sizes into row offsets and checks that every row pointer is
  // where a contiguous block puts it.  Only then is the block itself read (8 B per entry, one
  // copy) and split into ids / values on the device.
  uint64_t first = 0;
  while (first < n_rows && r[first].size == 0) first++;
  if (first == n_rows) {  // no entries at all
    std::vector<uint64_t> rp;
    try {
      rp.assign(n_rows + 1, 0);
    } catch (const std::bad_alloc&) {
      return fail("out of host memory");
    }
    uint32_t dc = 0;
    float dv = 0.f;
    return upload_common(c, slot, n_rows, 0, rp.data(), &dc, &dv, target);
  }
  const unsigned long long base = (unsigned long long)(uintptr_t)r[first].data;
  cudaStream_t st = c->stream;
  DevTmp d_rows, d_rp, d_scr, d_ent;
  unsigned int* flag = c->d_flag;
  CK(cudaMalloc(&d_rows.p, n_rows * sizeof(AosRow)));
  CK(cudaMalloc(&d_rp.p, (n_rows + 1) * sizeof(uint64_t)));
  CK(cudaMalloc(&d_scr.p, (aos_scan_tiles(n_rows) + 1) * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(flag, 0, 16 * sizeof(unsigned int), st));
  CK(cudaMemcpyAsync(d_rows.p, r, n_rows * sizeof(AosRow), cudaMemcpyHostToDevice, st));
  CK(launch_aos_to_csr(c, d_rows.p, nullptr, n_rows, 0, base, static_cast<unsigned long long*>(d_scr.p),
                       static_cast<uint64_t*>(d_rp.p), nullptr, nullptr, flag));
  uint64_t nnz = 0;
  CK(cudaMemcpyAsync(c->h_flag, flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&nnz, static_cast<uint64_t*>(d_rp.p) + n_rows, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (c->h_flag[0] & 1u) return upload_aos_host_gather(c, slot, n_rows, r, target);
  if (upload_begin(c, slot, n_rows, nnz, st)) return 1;
  DataSlot& s = c->slots[slot];
  CK(cudaMalloc(&d_ent.p, (nnz ? nnz : 1) * sizeof(AosEntry)));
  CK(cudaMemcpyAsync(d_ent.p, r[first].data, nnz * sizeof(AosEntry), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.row_ptr, d_rp.p, (n_rows + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(s.target, target, n_rows * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(launch_aos_split(c, d_ent.p, nnz, s.col, s.val));
  if (upload_inspect(c, slot, st)) return 1;
  return upload_finish(c, slot);  // syncs: the temporaries may be released
}

int fmb200_upload_onehot(fmb200_ctx* c, int slot, uint64_t n_rows, uint32_t nnz_per_row,
                         const uint32_t* ids, const float* target) {
  NEED_CTX(c);
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (n_rows && (!target || (nnz_per_row && !ids))) return fail("null data pointer");
  if (n_rows > 0xffffffffull) return fail("row count exceeds the reference's uint range");
  if (bind(c)) return 1;
  if (upload_onehot_enqueue(c, slot, n_rows, nnz_per_row, ids, target, c->stream)) return 1;
  return upload_finish(c, slot);
}

int fmb200_upload_onehot_async(fmb200_ctx* c, int slot, uint64_t n_rows, uint32_t nnz_per_row,
                               const uint32_t* ids, const float* target) {
  NEED_CTX(c);
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (n_rows && (!target || (nnz_per_row && !ids))) return fail("null data pointer");
  if (n_rows > 0xffffffffull) return fail("row count exceeds the reference's uint range");
  if (bind(c)) return 1;
  if (c->copy_stream == nullptr) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  return upload_onehot_enqueue(c, slot, n_rows, nnz_per_row, ids, target, c->copy_stream);
}

int fmb200_host_alloc(void** out, uint64_t bytes) {
  if (!out) return fail("null out pointer");
  *out = nullptr;
  CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}

int fmb200_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return 0;
}

int fmb200_free_data(fmb200_ctx* c, int slot) {
  NEED_CTX(c);
  if (slot < 0 || slot >= FMB200_MAX_SLOTS) return fail("slot %d out of range", slot);
  if (bind(c)) return 1;
  CK(cudaStreamSynchronize(c->stream));
  if (c->slots[slot].pending) CK(cudaEventSynchronize(c->slots[slot].ready));
  free_slot(c->slots[slot]);
  return 0;
}

int fmb200_set_params(fmb200_ctx* c, double w0, const double* w, const double* v) {
  NEED_CTX(c);
  if ((c->n && !w) || ((uint64_t)c->n * c->k && !v)) return fail("null parameter pointer");
  if (bind(c)) return 1;
  return guarded([&]() -> int {
  const uint32_t n = c->n;
    const int k = c->k, kp = c->kp;
    // fp64 image: [w0 | w | V attribute-major]
    std::vector<double> h64(c->p64.n_doubles);
    h64[0] = w0;
    for (uint32_t i = 0; i < n; i++) h64[Params64::off_w + i] = w[i];
    double* hv = h64.data() + c->p64.off_v;
    for (int f = 0; f < k; f++)
      for (uint32_t i = 0; i < n; i++) hv[(size_t)i * k + f] = v[(size_t)f * n + i];
    // fp32 packed image
    std::vector<float> h32(c->p32.n_floats, 0.f);
    h32[0] = (float)w0;
    for (uint32_t i = 0; i < n; i++) h32[c->p32.off_w + (size_t)i * c->p32.ws] = (float)w[i];
    float* hv32 = h32.data() + c->p32.off_v;
    for (int f = 0; f < k; f++)
      for (uint32_t i = 0; i < n; i++) hv32[(size_t)i * kp + f] = (float)v[(size_t)f * n + i];
    CK(cudaMemcpyAsync(c->p64.base, h64.data(), h64.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->p32.base, h32.data(), h32.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->peer_base_valid = false;
    c->hogwild_fresh = true;
    return 0;
  });
}

int fmb200_get_params(fmb200_ctx* c, double* w0, double* w, double* v) {
  NEED_CTX(c);
  if (!w0 || (c->n && !w) || ((uint64_t)c->n * c->k && !v)) return fail("null parameter pointer");
  if (bind(c)) return 1;
  return guarded([&]() -> int {
  const uint32_t n = c->n;
    const int k = c->k, kp = c->kp;
    if (c->mode != FMB200_MODE_HOGWILD) {
      std::vector<double> pageable;
      double* h = static_cast<double*>(c->h_stage);  // pinned staging for small models
      if (h == nullptr) {
        pageable.resize(c->p64.n_doubles);
        h = pageable.data();
      }
      CK(cudaMemcpyAsync(h, c->p64.base, c->p64.n_doubles * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      *w0 = h[0];
      for (uint32_t i = 0; i < n; i++) w[i] = h[Params64::off_w + i];
      const double* hv = h + c->p64.off_v;
      for (int f = 0; f < k; f++)
        for (uint32_t i = 0; i < n; i++) v[(size_t)f * n + i] = hv[(size_t)i * k + f];
    } else {
      std::vector<float> pageable;
      float* h = static_cast<float*>(c->h_stage);
      if (h == nullptr) {
        pageable.resize(c->p32.n_floats);
        h = pageable.data();
      }
      CK(cudaMemcpyAsync(h, c->p32.base, c->p32.n_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      *w0 = h[0];
      for (uint32_t i = 0; i < n; i++) w[i] = h[c->p32.off_w + (size_t)i * c->p32.ws];
      const float* hv = h + c->p32.off_v;
      for (uint32_t i = 0; i < n; i++)
        for (int f = 0; f < k; f++) v[(size_t)f * n + i] = hv[(size_t)i * kp + f];
    }
    return 0;
  });
}

int fmb200_sgd_epoch_async(fmb200_ctx* c, int slot) {
  NEED_CTX(c);
  if (need_slot(c, slot)) return 1;
  if (bind(c)) return 1;
  DataSlot& d = c->slots[slot];
  if (c->mode == FMB200_MODE_ORDERED) {
    if (c->k > 256) return fail("num_factor > 256 is not supported in the fp64 modes");
    bool handled = false;
    CK(launch_sgd_ordered(c, d, &handled));
    // shapes the ring cannot hold run row-at-a-time: the same order, just slower
    if (!handled) CK(launch_sgd_inorder(c, d));
  } else if (c->mode == FMB200_MODE_INORDER) {
    if (c->k > 256) return fail("num_factor > 256 is not supported in the fp64 modes");
    CK(launch_sgd_inorder(c, d));
  } else {
    if (c->kp > 128) return fail("num_factor > 128 is not supported in HOGWILD mode");
    CK(peer_before_epoch(c, d));  // multi-GPU: theta0 + shard counts for the mean-field combine
    CK(launch_sgd_hogwild(c, d));
  }
  return 0;
}

int fmb200_sync(fmb200_ctx* c) {
  NEED_CTX(c);
  if (bind(c)) return 1;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_sgd_epoch(fmb200_ctx* c, int slot, double* device_seconds) {
  NEED_CTX(c);
  if (bind(c)) return 1;
  CK(cudaEventRecord(c->ev0, c->stream));
  if (fmb200_sgd_epoch_async(c, slot)) return 1;
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (device_seconds) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *device_seconds = (double)ms * 1e-3;
  }
  return 0;
}

int fmb200_evaluate(fmb200_ctx* c, int slot, double* sum_sq_err, double* sum_abs_err,
                    uint64_t* n_correct) {
  NEED_CTX(c);
  if (need_slot(c, slot)) return 1;
  if (bind(c)) return 1;
  const DataSlot& d = c->slots[slot];
  double sq = 0, ab = 0, ok = 0;
  if (d.n_rows > 0) {
    const int nb = metric_blocks(c, d);
    if (ensure_partials(c, nb)) return 1;
    if (c->mode != FMB200_MODE_HOGWILD && c->hp.task == FMB200_TASK_REGRESSION) {
      // fp64 modes: the reference sums err*err and |err| left to right over the rows
      // (fm_learn.h:136-146); a tree reduction may differ in the last ulps and flip a printed
      // digit.  The kernel writes the per-row error, the host adds them in row order.
      if (c->pred_cap < d.n_rows) {
        if (c->d_pred) cudaFree(c->d_pred);
        c->d_pred = nullptr;
        c->pred_cap = 0;
        CK(cudaMalloc(&c->d_pred, d.n_rows * sizeof(double)));
        c->pred_cap = d.n_rows;
      }
      CK(launch_predict64(c, d, 2, c->d_pred, nullptr, nb));
      std::vector<double> err;
      try {
        err.resize(d.n_rows);
      } catch (const std::bad_alloc&) {
        return fail("out of host memory");
      }
      CK(cudaMemcpyAsync(err.data(), c->d_pred, d.n_rows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      for (uint64_t i = 0; i < d.n_rows; i++) {
        sq += err[i] * err[i];
        ab += std::abs(err[i]);
      }
    } else {
      if (c->mode != FMB200_MODE_HOGWILD) {
        CK(launch_predict64(c, d, 0, nullptr, c->d_partials, nb));
      } else {
        CK(launch_predict32(c, d, 0, nullptr, c->d_partials, nb));
      }
      std::vector<double> h;
      try {
        h.resize(3 * (size_t)nb);
      } catch (const std::bad_alloc&) {
        return fail("out of host memory");
      }
      CK(cudaMemcpyAsync(h.data(), c->d_partials, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      for (int b = 0; b < nb; b++) {  // fixed order: deterministic result
        sq += h[3 * b + 0];
        ab += h[3 * b + 1];
        ok += h[3 * b + 2];
      }
    }
  }
  if (sum_sq_err) *sum_sq_err = sq;
  if (sum_abs_err) *sum_abs_err = ab;
  if (n_correct) *n_correct = (uint64_t)llround(ok);
  return 0;
}

int fmb200_predict(fmb200_ctx* c, int slot, int transform, double* out) {
  NEED_CTX(c);
  if (need_slot(c, slot)) return 1;
  if (bind(c)) return 1;
  const DataSlot& d = c->slots[slot];
  if (d.n_rows == 0) return 0;
  if (!out) return fail("null output pointer");
  if (c->pred_cap < d.n_rows) {
    if (c->d_pred) cudaFree(c->d_pred);
    c->d_pred = nullptr;
    c->pred_cap = 0;
    CK(cudaMalloc(&c->d_pred, d.n_rows * sizeof(double)));
    c->pred_cap = d.n_rows;
  }
  const int nb = metric_blocks(c, d);
  if (c->mode != FMB200_MODE_HOGWILD) {
    CK(launch_predict64(c, d, transform, c->d_pred, nullptr, nb));
  } else {
    CK(launch_predict32(c, d, transform, c->d_pred, nullptr, nb));
  }
  CK(cudaMemcpyAsync(out, c->d_pred, d.n_rows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_sgda_begin(fmb200_ctx* c, uint32_t n_groups, const uint32_t* attr_group) {
  NEED_CTX(c);
  if (bind(c)) return 1;
  if (c->mode == FMB200_MODE_HOGWILD) return fail("SGDA runs on the fp64 state: set INORDER or ORDERED mode first");
  if (n_groups == 0 || n_groups > 1024) return fail("n_groups must be in [1,1024]");
  if (n_groups > 1 && !attr_group) return fail("attr_group is required for more than one group");
  if (attr_group)
    for (uint32_t i = 0; i < c->n; i++)
      if (attr_group[i] >= n_groups) return fail("attr_group[%u] = %u >= n_groups", i, attr_group[i]);
  const size_t n1 = c->n ? c->n : 1, nk = (size_t)c->n * c->k ? (size_t)c->n * c->k : 1;
  const size_t gk = (size_t)n_groups * (c->k ? c->k : 1);
  if (!c->sgda_grad_w) {
    CK(cudaMalloc(&c->sgda_grad_w, n1 * sizeof(double)));
    CK(cudaMalloc(&c->sgda_grad_v, nk * sizeof(double)));
    CK(cudaMalloc(&c->sgda_group, n1 * sizeof(uint32_t)));
  }
  if (c->sgda_groups != n_groups) {
    if (c->sgda_reg_w) cudaFree(c->sgda_reg_w);
    if (c->sgda_reg_v) cudaFree(c->sgda_reg_v);
    c->sgda_reg_w = c->sgda_reg_v = nullptr;
    CK(cudaMalloc(&c->sgda_reg_w, n_groups * sizeof(double)));
    CK(cudaMalloc(&c->sgda_reg_v, gk * sizeof(double)));
    c->sgda_groups = n_groups;
  }
  // init(): grad_w = grad_v = 0 (:73-74); learn(): w = 0, reg_w = reg_v = 0 (:283-291)
  CK(cudaMemsetAsync(c->sgda_grad_w, 0, n1 * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->sgda_grad_v, 0, nk * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->sgda_reg_w, 0, n_groups * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->sgda_reg_v, 0, gk * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->p64.w(), 0, n1 * sizeof(double), c->stream));
  if (attr_group) CK(cudaMemcpyAsync(c->sgda_group, attr_group, c->n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  else CK(cudaMemsetAsync(c->sgda_group, 0, n1 * sizeof(uint32_t), c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_sgda_epoch(fmb200_ctx* c, int train_slot, int val_slot, int lambda_steps, double* device_seconds) {
  NEED_CTX(c);
  if (need_slot(c, train_slot) || need_slot(c, val_slot)) return 1;
  if (bind(c)) return 1;
  if (c->mode == FMB200_MODE_HOGWILD) return fail("SGDA runs on the fp64 state: set INORDER or ORDERED mode first");
  if (c->sgda_groups == 0) return fail("call fmb200_sgda_begin first");
  if (c->k > 256) return fail("num_factor > 256 is not supported in the fp64 modes");
  CK(cudaEventRecord(c->ev0, c->stream));
  CK(launch_sgda_epoch(c, c->slots[train_slot], c->slots[val_slot], lambda_steps));
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (device_seconds) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *device_seconds = (double)ms * 1e-3;
  }
  return 0;
}

int fmb200_sgda_get_reg(fmb200_ctx* c, double* reg_w, double* reg_v) {
  NEED_CTX(c);
  if (bind(c)) return 1;
  if (c->sgda_groups == 0) return fail("call fmb200_sgda_begin first");
  if (reg_w) CK(cudaMemcpyAsync(reg_w, c->sgda_reg_w, c->sgda_groups * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (reg_v && c->k)
    CK(cudaMemcpyAsync(reg_v, c->sgda_reg_v, (size_t)c->sgda_groups * c->k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_mcmc_eterms(fmb200_ctx* c, int slot, double* e_out) {
  NEED_CTX(c);
  if (need_slot(c, slot)) return 1;
  if (bind(c)) return 1;
  if (c->mode == FMB200_MODE_HOGWILD)
    return fail("e-terms are computed from the fp64 state: set INORDER or ORDERED mode first");
  const DataSlot& d = c->slots[slot];
  if (d.n_rows == 0) return 0;
  if (!e_out) return fail("null output pointer");
  if (c->pred_cap < d.n_rows) {
    if (c->d_pred) cudaFree(c->d_pred);
    c->d_pred = nullptr;
    c->pred_cap = 0;
    CK(cudaMalloc(&c->d_pred, d.n_rows * sizeof(double)));
    c->pred_cap = d.n_rows;
  }
  CK(launch_mcmc_eterms(c, d, c->d_pred));
  CK(cudaMemcpyAsync(e_out, c->d_pred, d.n_rows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_params_device(fmb200_ctx* c, void** device_ptr, uint64_t* n_floats) {
  NEED_CTX(c);
  if (c->mode != FMB200_MODE_HOGWILD) return fail("packed fp32 state is live only in HOGWILD mode");
  if (device_ptr) *device_ptr = c->p32.base;
  if (n_floats) *n_floats = c->p32.n_floats;
  return 0;
}

int fmb200_params_layout(fmb200_ctx* c, uint64_t* off_w, int* ws, uint64_t* off_v, int* kp) {
  NEED_CTX(c);
  if (off_w) *off_w = c->p32.off_w;
  if (ws) *ws = c->p32.ws;
  if (off_v) *off_v = c->p32.off_v;
  if (kp) *kp = c->kp;
  return 0;
}

int fmb200_scale_params(fmb200_ctx* c, double factor) {
  NEED_CTX(c);
  if (c->mode != FMB200_MODE_HOGWILD) return fail("scale_params applies to the HOGWILD state");
  if (bind(c)) return 1;
  CK(launch_scale_p32(c, (float)factor));
  c->peer_base_valid = false;
  return 0;
}

int fmb200_peer_export(fmb200_ctx* c, void* handle) {
  NEED_CTX(c);
  if (!handle) return fail("null handle pointer");
  if (bind(c)) return 1;
  static_assert(sizeof(cudaIpcMemHandle_t) == FMB200_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->comm_base));
  memcpy(handle, &h, sizeof(h));
  return 0;
}

static int peer_precheck(fmb200_ctx* c, int world, int rank) {
  if (world < 1 || world > FMB200_MAX_PEERS) return fail("world %d outside [1,%d]", world, FMB200_MAX_PEERS);
  if (rank < 0 || rank >= world) return fail("rank %d outside world %d", rank, world);
  if (c->peer_seq != 0 || c->peer_world != 1) return fail("peers are already attached");
  if (c->mode != FMB200_MODE_HOGWILD) return fail("peer averaging applies to the HOGWILD state");
  return 0;
}

int fmb200_peer_attach_ipc(fmb200_ctx* c, int world, int rank, const void* handles) {
  NEED_CTX(c);
  if (!handles) return fail("null handles");
  if (peer_precheck(c, world, rank)) return 1;
  if (bind(c)) return 1;
  const unsigned char* hb = static_cast<const unsigned char*>(handles);
  for (int q = 0; q < world; q++) {
    if (q == rank) {
      c->peer_base[q] = c->comm_base;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)q * FMB200_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[q] = static_cast<unsigned char*>(p);
    c->peer_ipc[q] = true;
  }
  CK(peer_preload_kernels());
  c->peer_world = world;
  c->peer_rank = rank;
  return 0;
}

int fmb200_peer_attach_local(fmb200_ctx* c, int world, int rank, fmb200_ctx* const* all) {
  NEED_CTX(c);
  if (!all) return fail("null context list");
  if (peer_precheck(c, world, rank)) return 1;
  if (bind(c)) return 1;
  for (int q = 0; q < world; q++) {
    if (!all[q] || all[q]->p32.n_floats != c->p32.n_floats) return fail("peer %d has a different model shape", q);
    if (q != rank && all[q]->device != c->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, c->device, all[q]->device));
      if (!can) return fail("device %d cannot access device %d", c->device, all[q]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(all[q]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail("cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
    }
    c->peer_base[q] = all[q]->comm_base;
  }
  CK(peer_preload_kernels());
  c->peer_world = world;
  c->peer_rank = rank;
  return 0;
}

int fmb200_allreduce_mean(fmb200_ctx* c) {
  NEED_CTX(c);
  if (c->mode != FMB200_MODE_HOGWILD) return fail("peer averaging applies to the HOGWILD state");
  if (c->peer_world <= 1) return 0;
  if (bind(c)) return 1;
  CK(launch_peer_mean(c));
  return 0;
}

int fmb200_allreduce_meanfield(fmb200_ctx* c) {
  NEED_CTX(c);
  if (c->mode != FMB200_MODE_HOGWILD) return fail("the peer exchange applies to the HOGWILD state");
  if (c->peer_world <= 1) return 0;
  if (!c->peer_base_valid) return fail("no epoch has run since the state was last set: nothing to combine");
  if (bind(c)) return 1;
  CK(launch_peer_meanfield(c));
  return 0;
}

int fmb200_peer_barrier(fmb200_ctx* c) {
  NEED_CTX(c);
  if (c->peer_world <= 1) return 0;
  if (bind(c)) return 1;
  CK(launch_peer_barrier(c));
  return 0;
}

int fmb200_stream(fmb200_ctx* c, void** cuda_stream) {
  NEED_CTX(c);
  if (cuda_stream) *cuda_stream = (void*)c->stream;
  return 0;
}

int fmb200_kernel_launches(fmb200_ctx* c, uint64_t* count) {
  NEED_CTX(c);
  if (count) *count = c->launches;
  return 0;
}

int fmb200_last_epoch_config(fmb200_ctx* c, int* lanes_per_row, int* slots, int* rows_per_tile,
                             int* grid, int* block, int* smem_bytes, int* damp) {
  NEED_CTX(c);
  if (lanes_per_row) *lanes_per_row = c->last_cfg.lanes_per_row;
  if (slots) *slots = c->last_cfg.slots;
  if (rows_per_tile) *rows_per_tile = c->last_cfg.rows_per_tile;
  if (grid) *grid = c->last_cfg.grid;
  if (block) *block = c->last_cfg.block;
  if (smem_bytes) *smem_bytes = c->last_cfg.smem;
  if (damp) *damp = c->last_cfg.damp;
  return 0;
}

int fmb200_download_data(fmb200_ctx* c, int slot, uint64_t* n_rows, uint64_t* nnz, uint64_t* row_ptr,
                         uint32_t* col, float* val, float* target) {
  NEED_CTX(c);
  if (need_slot(c, slot)) return 1;
  if (bind(c)) return 1;
  const DataSlot& d = c->slots[slot];
  if (n_rows) *n_rows = d.n_rows;
  if (nnz) *nnz = d.nnz;
  if (row_ptr) CK(cudaMemcpyAsync(row_ptr, d.row_ptr, (d.n_rows + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  if (col && d.nnz) CK(cudaMemcpyAsync(col, d.col, d.nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (val && d.nnz) CK(cudaMemcpyAsync(val, d.val, d.nnz * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (target && d.n_rows) CK(cudaMemcpyAsync(target, d.target, d.n_rows * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_ordered_index(fmb200_ctx* c, int slot, uint32_t* link, uint32_t* rowdep) {
  NEED_CTX(c);
  if (need_slot(c, slot)) return 1;
  if (bind(c)) return 1;
  DataSlot& d = c->slots[slot];
  if (d.nnz >= 0xffffffffull) return fail("the ORDERED index needs nnz < 2^32-1");
  CK(build_ordered_links(c, d));
  if (link && d.nnz)
    CK(cudaMemcpyAsync(link, d.link, d.nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (rowdep && d.n_rows)
    CK(cudaMemcpyAsync(rowdep, d.rowdep, d.n_rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int fmb200_set_tuning(fmb200_ctx* c, int ctas_per_sm, int rows_per_tile, int threads, int damp,
                      int variant) {
  NEED_CTX(c);
  if (threads && (threads % 32 != 0 || threads < 32 || threads > 1024))
    return fail("threads must be a multiple of 32 in [32,1024]");
  if (rows_per_tile && (rows_per_tile < 32 || rows_per_tile > 512))
    return fail("rows_per_tile must be in [32,512]");
  c->tune_ctas_per_sm = ctas_per_sm;
  c->tune_rows_per_tile = rows_per_tile;
  c->tune_threads = threads;
  c->tune_damp = damp;
  c->tune_variant = variant;
  return 0;
}

}  // extern "C"
