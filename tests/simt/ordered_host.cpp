// ordered_host.cpp -- TEST INFRASTRUCTURE.  Compiles libfm_b200/csrc/fm_ordered.cuh (the ORDERED
// epoch kernel's own source) for the host through cta_shim.h, runs it as one OS thread per CUDA
// thread on seeded synthetic data and compares the resulting w0 / w / V with the sequential oracle
// (oracle/fm_oracle.c: fmo_sgd_epoch).  The kernel re-associates three sums, so the comparison is
// to 1e-10 relative, not bitwise.  Exit status 0 = every case within tolerance.
// Also the SPEC of the link / rowdep index the device builds in fm_ordered.cu (build_links below).
#include "cta_shim.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../libfm_b200/csrc/fm_ordered.cuh"

namespace simt {
thread_local Dim tid;
Dim bdim;
pthread_barrier_t cta_barrier;
pthread_barrier_t warp_barrier[kMaxWarps];
pthread_barrier_t named_barrier[4];
uint64_t xchg[kMaxWarps][32];
}  // namespace simt

extern "C" void fmo_sgd_epoch(uint32_t n, int k, int k0, int k1, double* w0, double* w, double* v, double lr,
                              double reg0, double regw, double regv, int task, double min_target,
                              double max_target, uint64_t n_rows, const uint64_t* row_ptr,
                              const uint32_t* col, const float* val, const float* target);

namespace {

struct Rng {  // xorshift64*
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 1) {}
  uint64_t next() {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return s * 0x2545F4914F6CDD1Dull;
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(uni() * n); }
  double gauss() {
    double a = 0;
    for (int i = 0; i < 6; i++) a += uni();
    return (a - 3.0) * 1.41;
  }
};

struct Case {
  const char* name;
  uint64_t n_rows;
  uint32_t n_feat;
  int k, k0, k1, task;
  double regs[3];
  int max_len;     // 0: exactly 2 one-hot entries (user, item); else 0..max_len entries with real values
  int dup_every;   // every so many rows repeat the first id inside the row
  int TR;          // forced tile rows
  int warps;
  double lr;
  int helper_warps = 0;  // > 0: the warp-specialised driver
};

struct Csr {
  std::vector<uint64_t> row_ptr;
  std::vector<uint32_t> col;
  std::vector<float> val, target;
};

Csr make_data(const Case& c, uint64_t seed) {
  Rng r(seed);
  Csr d;
  d.row_ptr.push_back(0);
  for (uint64_t i = 0; i < c.n_rows; i++) {
    if (c.max_len == 0) {
      const uint32_t half = c.n_feat / 2;
      d.col.push_back(r.below(half));
      d.val.push_back(1.f);
      d.col.push_back(half + r.below(c.n_feat - half));
      d.val.push_back(1.f);
    } else {
      const int size = (int)r.below(c.max_len + 1);
      for (int j = 0; j < size; j++) {
        d.col.push_back(r.below(c.n_feat));
        d.val.push_back((float)(r.gauss() * 0.7));
      }
      if (c.dup_every && size >= 2 && i % c.dup_every == 0) {
        const size_t b = d.row_ptr.back();
        d.col[b + size - 1] = d.col[b];
        if (size >= 3 && i % (2 * c.dup_every) == 0) d.col[b + 1] = d.col[b];  // a triple
      }
    }
    d.row_ptr.push_back(d.col.size());
    if (c.task == 0) d.target.push_back((float)(1 + r.below(5)));
    else d.target.push_back(r.uni() < 0.5 ? -1.f : 1.f);
  }
  return d;
}

// the index the ORDERED epoch needs (what fm_ordered.cu builds on the device)
void build_links(const Csr& d, uint32_t n_feat, std::vector<uint32_t>& link, std::vector<uint32_t>& rowdep) {
  const uint64_t n_rows = d.row_ptr.size() - 1;
  link.assign(d.col.size(), fmb::ORD_NONE);
  rowdep.assign(n_rows, fmb::ORD_NONE);
  std::vector<int64_t> last(n_feat, -1), lastrow(n_feat, -1);
  for (uint64_t r = 0; r < n_rows; r++)
    for (uint64_t e = d.row_ptr[r]; e < d.row_ptr[r + 1]; e++) {
      const uint32_t f = d.col[e];
      if (last[f] >= 0) {
        link[e] = (uint32_t)(e - (uint64_t)last[f]);
        rowdep[r] = std::min(rowdep[r], (uint32_t)(r - (uint64_t)lastrow[f]));
      }
      last[f] = (int64_t)e;
      lastrow[f] = (int64_t)r;
    }
}

// helper_warps > 0 runs the warp-specialised driver: `nthreads` compute threads + 32 * helper_warps helpers
template <int GL, int KF, int ZF = 0>
void run_threads(const fmb::OrderedArgs& a, unsigned char* smem, int task, int nthreads, int helper_warps = 0) {
  const int ncompute = nthreads;
  nthreads += 32 * helper_warps;
  simt::bdim.x = (unsigned)nthreads;
  pthread_barrier_init(&simt::cta_barrier, nullptr, nthreads);
  pthread_barrier_init(&simt::named_barrier[1], nullptr, ncompute);
  if (helper_warps) pthread_barrier_init(&simt::named_barrier[2], nullptr, 32 * helper_warps);
  for (int w = 0; w < nthreads / 32; w++) pthread_barrier_init(&simt::warp_barrier[w], nullptr, 32);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t]() {
      simt::tid.x = (unsigned)t;
      if (helper_warps) {
        if (task == 0) fmb::ordered_epoch_body_ws<GL, KF, 0, ZF>(a, smem, ncompute, 0);
        else fmb::ordered_epoch_body_ws<GL, KF, 1, ZF>(a, smem, ncompute, 0);
      } else {
        if (task == 0) fmb::ordered_epoch_body<GL, KF, 0, ZF>(a, smem);
        else fmb::ordered_epoch_body<GL, KF, 1, ZF>(a, smem);
      }
    });
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&simt::cta_barrier);
  pthread_barrier_destroy(&simt::named_barrier[1]);
  if (helper_warps) pthread_barrier_destroy(&simt::named_barrier[2]);
  for (int w = 0; w < nthreads / 32; w++) pthread_barrier_destroy(&simt::warp_barrier[w]);
}

void dispatch(int k, const fmb::OrderedArgs& a, unsigned char* smem, int task, int nthreads, int max_nnz, int hw) {
  // the register-resident fast path where the launcher would pick it (fm_ordered.cu::pick_fast_kernel)
  if (k == 8 && max_nnz >= 1 && max_nnz <= 2) return run_threads<1, 8, 2>(a, smem, task, nthreads, hw);
  if (k == 8 && max_nnz >= 1 && max_nnz <= 4) return run_threads<1, 8, 4>(a, smem, task, nthreads, hw);
  if (k == 4 && max_nnz >= 1 && max_nnz <= 4) return run_threads<1, 4, 4>(a, smem, task, nthreads, hw);
  if (k <= 1) run_threads<1, 1>(a, smem, task, nthreads, hw);
  else if (k <= 2) run_threads<1, 2>(a, smem, task, nthreads, hw);
  else if (k <= 4) run_threads<1, 4>(a, smem, task, nthreads, hw);
  else if (k <= 8) run_threads<1, 8>(a, smem, task, nthreads, hw);
  else if (k <= 16) run_threads<2, 8>(a, smem, task, nthreads, hw);
  else if (k <= 32) run_threads<4, 8>(a, smem, task, nthreads, hw);
  else if (k <= 64) run_threads<8, 8>(a, smem, task, nthreads, hw);
  else if (k <= 128) run_threads<16, 8>(a, smem, task, nthreads, hw);
  else run_threads<32, 8>(a, smem, task, nthreads, hw);
}

bool run_case(const Case& c) {
  Csr d = make_data(c, 1234 + c.n_rows + c.k);
  const uint64_t N = c.n_rows, nnz = d.col.size();
  const uint32_t n = c.n_feat;
  const int k = c.k;
  std::vector<uint32_t> link, rowdep;
  build_links(d, n, link, rowdep);

  // padded copies (the device arrays carry slack for whole-tile bulk copies)
  const size_t RS = 600, ES = 32;
  std::vector<uint64_t> rp(N + 1 + RS, nnz);
  std::copy(d.row_ptr.begin(), d.row_ptr.end(), rp.begin());
  std::vector<float> tg(N + RS, 0.f);
  std::copy(d.target.begin(), d.target.end(), tg.begin());
  std::vector<uint32_t> rd(N + RS, fmb::ORD_NONE);
  std::copy(rowdep.begin(), rowdep.end(), rd.begin());
  std::vector<uint32_t> col(nnz + ES, 0), lk(nnz + ES, fmb::ORD_NONE);
  std::vector<float> val(nnz + ES, 0.f);
  std::copy(d.col.begin(), d.col.end(), col.begin());
  std::copy(d.val.begin(), d.val.end(), val.begin());
  std::copy(link.begin(), link.end(), lk.begin());

  // state: kernel layout [w0, pad | w (even) | V[n][k] | pad], oracle layout factor-major
  Rng r(99);
  const uint64_t off_w = 2, off_v = off_w + ((n + 1ull) & ~1ull);
  std::vector<double> st_raw(off_v + (uint64_t)n * k + 4 + 2, 0.0);
  double* st = st_raw.data();
  if (((uintptr_t)st & 15) != 0) st++;  // 16-byte alignment of the base
  std::vector<double> ow(n), ov((size_t)k * n);
  double ow0 = 0.05;
  st[0] = ow0;
  for (uint32_t i = 0; i < n; i++) {
    ow[i] = r.gauss() * 0.1;
    st[off_w + i] = ow[i];
  }
  for (int f = 0; f < k; f++)
    for (uint32_t i = 0; i < n; i++) {
      const double x = r.gauss() * 0.1;
      ov[(size_t)f * n + i] = x;
      st[off_v + (size_t)i * k + f] = x;
    }

  // geometry
  const int TR = c.TR;
  uint64_t span = 0;
  for (uint64_t r0 = 0; r0 < N; r0 += TR) {
    const uint64_t r1 = std::min<uint64_t>(r0 + TR, N);
    const uint64_t ab = d.row_ptr[r0] & ~3ull, ae = (d.row_ptr[r1] + 3) & ~3ull;
    span = std::max(span, ae - ab);
  }
  const uint32_t TE = (uint32_t)span + 4;
  const int kw = (k & 1) ? k + 1 : k, rs = kw + 2;
  const size_t smem_bytes = fmb::ord_smem_bytes(TR, TE, rs);
  std::vector<unsigned char> smem_raw(smem_bytes + 256, 0xcd);  // poison: unfetched records must never be read
  unsigned char* smem = smem_raw.data();
  smem += (128 - ((uintptr_t)smem & 127)) & 127;

  fmb::OrderedArgs a{};  // (zero: a.prof, the phase timers, must be null here)
  a.row_ptr = rp.data();
  a.col = col.data();
  a.val = val.data();
  a.target = tg.data();
  a.link = lk.data();
  a.rowdep = rd.data();
  uint32_t shape = 3u;  // what fm_ordered.cu::ord_shape_kernel computes
  {
    uint64_t mx = 0;
    for (uint64_t i = 0; i < N; i++) mx = std::max<uint64_t>(mx, d.row_ptr[i + 1] - d.row_ptr[i]);
    for (uint64_t i = 0; i < N; i++)
      if (d.row_ptr[i + 1] - d.row_ptr[i] != mx) shape &= ~2u;
    for (float x : d.val)
      if (x != 1.f) shape &= ~1u;
  }
  a.shape = &shape;
  a.n_rows = N;
  a.n_tiles = (uint32_t)((N + TR - 1) / TR);
  a.tile_rows = TR;
  a.tile_cap = TE;
  a.w0 = st;
  a.w = st + off_w;
  a.v = st + off_v;
  a.k = k;
  a.kw = kw;
  a.rs = rs;
  a.use_w0 = c.k0;
  a.use_w = c.k1;
  a.lr = c.lr;
  a.reg0 = c.regs[0];
  a.regw = c.regs[1];
  a.regv = c.regs[2];
  a.min_target = c.task == 0 ? 1.0 : -1.0;
  a.max_target = c.task == 0 ? 5.0 : 1.0;
  a.csr_bytes = fmb::ord_csr_bytes(TR, TE);
  a.rec_bytes = TE * (uint32_t)rs * 8u;
  a.debug = 0;

  int max_nnz = 0;
  for (uint64_t i = 0; i < N; i++) max_nnz = std::max<int>(max_nnz, (int)(d.row_ptr[i + 1] - d.row_ptr[i]));
  const int epochs = 2;
  for (int ep = 0; ep < epochs; ep++) {
    if (N > 0) dispatch(k, a, smem, c.task, c.warps * 32, max_nnz, c.helper_warps);
    fmo_sgd_epoch(n, k, c.k0, c.k1, &ow0, ow.data(), ov.data(), c.lr, c.regs[0], c.regs[1], c.regs[2], c.task,
                  a.min_target, a.max_target, N, d.row_ptr.data(), d.col.data(), d.val.data(), d.target.data());
  }
  double worst = 0;
  auto cmp = [&](double got, double want) {
    const double err = fabs(got - want) / (1e-3 + fabs(want));
    if (!(err <= worst)) worst = err;  // NaN-safe
  };
  if (c.k0) cmp(st[0], ow0);
  for (uint32_t i = 0; i < n; i++) cmp(st[off_w + i], ow[i]);
  for (int f = 0; f < k; f++)
    for (uint32_t i = 0; i < n; i++) cmp(st[off_v + (size_t)i * k + f], ov[(size_t)f * n + i]);
  const bool ok = worst <= 1e-10;
  printf("%-28s rows=%-6llu k=%-3d TR=%-3d warps=%d+%d tiles=%u  worst rel err %.3g  %s\n", c.name,
         (unsigned long long)N, k, TR, c.warps, c.helper_warps, a.n_tiles, worst, ok ? "ok" : "FAIL");
  return ok;
}

}  // namespace

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "--quick");
  std::vector<Case> cases = {
      {"c2_like", 1200, 1000, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 64, 4, 0.02},
      {"c2_like_small_tiles", 700, 300, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 8, 2, 0.02},
      {"hot_features", 700, 40, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 32, 2, 0.02},
      {"ragged_real_values", 700, 400, 8, 1, 1, 0, {0, 0, 0}, 4, 0, 32, 2, 0.02},
      {"dups_in_row", 700, 300, 8, 1, 1, 0, {0.01, 0.02, 0.03}, 4, 5, 16, 2, 0.02},
      {"classification", 700, 500, 8, 1, 1, 1, {0, 0, 0}, 0, 0, 32, 2, 0.05},
      {"no_bias_no_linear", 700, 500, 8, 0, 0, 0, {0, 0, 0}, 0, 0, 32, 2, 0.02},
      {"k3_odd_reg", 700, 401, 3, 1, 1, 0, {0.01, 0.02, 0.03}, 4, 7, 32, 2, 0.02},
      {"k16_longer_rows", 500, 600, 16, 1, 1, 0, {0, 0, 0.01}, 12, 9, 8, 2, 0.01},
      {"k40_two_per_lane", 300, 300, 40, 1, 1, 1, {0, 0, 0}, 6, 0, 4, 2, 0.02},
      {"k1", 600, 200, 1, 1, 1, 0, {0, 0, 0}, 3, 4, 32, 1, 0.02},
      {"k0_linear_only", 600, 200, 0, 1, 1, 0, {0, 0, 0}, 3, 0, 32, 1, 0.02},
      {"single_row_tiles", 200, 100, 8, 1, 1, 0, {0, 0, 0}, 5, 3, 1, 2, 0.02},
      {"tiny", 5, 6, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 32, 2, 0.02},
      // the bias chain under stress: big steps push scores across the clamps (contradicted guesses, re-walks)
      {"oh_clamp_heavy", 700, 120, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 32, 4, 0.3},
      {"oh_regularised", 700, 300, 8, 1, 1, 0, {0.02, 0.01, 0.03}, 0, 0, 32, 4, 0.05},
      {"k8_max2_real_clamps", 700, 200, 8, 1, 1, 0, {0, 0.01, 0.01}, 2, 0, 16, 2, 0.25},
      {"k4_short_rows", 700, 200, 4, 1, 1, 0, {0, 0, 0.02}, 4, 6, 32, 2, 0.1},
      // the warp-specialised driver (helper warps write tile T-1 back and fetch tile T+1 while tile T runs)
      {"ws_oh_clamp_heavy", 700, 120, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 32, 4, 0.3, 1},
      {"ws_c2_like", 1200, 1000, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 64, 4, 0.02, 2},
      {"ws_small_tiles_hot", 700, 60, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 8, 2, 0.02, 1},
      {"ws_ragged_dups_k3", 700, 300, 3, 1, 1, 0, {0.01, 0.02, 0.03}, 4, 5, 16, 2, 0.02, 2},
      {"ws_k16_classification", 500, 600, 16, 1, 1, 1, {0, 0, 0.01}, 12, 9, 8, 2, 0.02, 1},
      {"ws_single_row_tiles", 200, 100, 8, 1, 1, 0, {0, 0, 0}, 5, 3, 1, 2, 0.02, 1},
      {"ws_tiny", 5, 6, 8, 1, 1, 0, {0, 0, 0}, 0, 0, 32, 2, 0.02, 1},
  };
  if (quick) cases.resize(3);
  bool all = true;
  for (const Case& c : cases) all = run_case(c) && all;
  printf(all ? "ALL OK\n" : "FAILURES\n");
  return all ? 0 : 1;
}
