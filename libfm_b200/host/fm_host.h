// fm_host.h -- host-side model image, R-style log and the SGD learner of the
// drop-in command line.  The learner keeps the reference's vtable surface
// (init / learn / evaluate / predict; reference src/libfm/src/fm_learn.h:31-60)
// but every pass over the data is ONE call into libfmb200 (include/fmb200.h).
#pragma once
#include <sys/resource.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "fmb200.h"
#include "sparse_data.h"

#ifdef FMB200_WITH_NCCL
#include <cuda_runtime_api.h>
#include <nccl.h>
#endif

namespace host {

// ---- RNG: libc rand() exactly as the reference consumes it -----------------
// (reference src/util/random.h:148-174; seeded by srand() at libfm.cpp:115-116)
inline double ran_uniform() { return rand() / ((double)RAND_MAX + 1); }
inline double ran_gaussian() {  // Leva's ratio-of-uniforms method
  double u, v, x, y, q;
  for (;;) {
    do {
      u = ran_uniform();
    } while (u == 0.0);
    v = 1.7156 * (ran_uniform() - 0.5);
    x = u - 0.449871;
    y = std::fabs(v) + 0.386595;
    q = x * x + y * (0.19600 * y - 0.25472 * x);
    if (q < 0.27597) break;
    if (!((q > 0.27846) || ((v * v) > (-4.0 * u * u * std::log(u))))) break;
  }
  return v / u;
}

// ---- model image ------------------------------------------------------------
// v is factor-major [num_factor][num_attribute] like the reference's
// DMatrixDouble (util/matrix.h:152-175) so that the init draw order and the
// model file layout coincide with it.
struct HostModel {
  uint32_t num_attribute = 0;
  int num_factor = 0;
  bool k0 = true, k1 = true;
  double reg0 = 0, regw = 0, regv = 0;
  double init_mean = 0, init_stdev = 0.01;
  double w0 = 0;
  std::vector<double> w, v;

  double& V(int f, uint32_t i) { return v[(size_t)f * num_attribute + i]; }

  // fm_model::init, fm_core/fm_model.h:91-99 + matrix.h:398-404
  void init() {
    w0 = 0;
    w.assign(num_attribute, 0.0);
    v.resize((size_t)num_factor * num_attribute);
    const bool constant = (init_stdev == 0.0) || std::isnan(init_stdev);
    for (auto& x : v) x = constant ? init_mean : init_mean + init_stdev * ran_gaussian();
  }

  void debug() const {  // fm_model.h:80-89
    std::cout << "num_attributes=" << num_attribute << std::endl;
    std::cout << "use w0=" << k0 << std::endl;
    std::cout << "use w1=" << k1 << std::endl;
    std::cout << "dim v =" << num_factor << std::endl;
    std::cout << "reg_w0=" << reg0 << std::endl;
    std::cout << "reg_w=" << regw << std::endl;
    std::cout << "reg_v=" << regv << std::endl;
    std::cout << "init ~ N(" << init_mean << "," << init_stdev << ")" << std::endl;
  }

  // text checkpoint, byte-compatible with fm_model::saveModel (fm_model.h:132-154)
  void save(const std::string& path) {
    std::ofstream out(path.c_str());
    if (k0) out << "#global bias W0" << std::endl << w0 << std::endl;
    if (k1) {
      out << "#unary interactions Wj" << std::endl;
      for (uint32_t i = 0; i < num_attribute; i++) out << w[i] << std::endl;
    }
    out << "#pairwise interactions Vj,f" << std::endl;
    for (uint32_t i = 0; i < num_attribute; i++) {
      for (int f = 0; f < num_factor; f++) {
        out << V(f, i);
        if (f != num_factor - 1) out << ' ';
      }
      out << std::endl;
    }
  }

  // fm_model::loadModel (fm_model.h:160-190): 1 = ok, 0 = malformed / missing.
  // Deviation: the reference's splitString yields no token for a line without a
  // blank, so num_factor == 1 files always read as malformed there; here they load.
  int load(const std::string& path) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) return 0;
    std::string line;
    if (k0) {
      if (!std::getline(in, line)) return 0;
      if (!std::getline(in, line)) return 0;
      w0 = atof(line.c_str());
    }
    if (k1) {
      if (!std::getline(in, line)) return 0;
      for (uint32_t i = 0; i < num_attribute; i++) {
        if (!std::getline(in, line)) return 0;
        w[i] = atof(line.c_str());
      }
    }
    if (!std::getline(in, line)) return 0;
    for (uint32_t i = 0; i < num_attribute; i++) {
      if (!std::getline(in, line)) return 0;
      std::vector<std::string> tok;
      size_t a = 0;
      for (;;) {
        size_t b = line.find(' ', a);
        tok.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
      }
      if ((int)tok.size() != num_factor) return 0;
      for (int f = 0; f < num_factor; f++) V(f, i) = atof(tok[f].c_str());
    }
    return 1;
  }
};

// ---- R-style measurement log (reference src/util/rlog.h:56-103) -------------
class RLog {
 public:
  explicit RLog(std::ostream* out) : out_(out) {}
  void add_field(const std::string& name, double dflt) {
    for (const auto& h : header_)
      if (h == name) throw "the field " + name + " already exists";
    header_.push_back(name);
    default_[name] = dflt;
  }
  void init() {
    write_row(true);
    reset();
  }
  void log(const std::string& field, double d) { value_[field] = d; }
  void new_line() {
    write_row(false);
    reset();
  }

 private:
  void reset() {
    value_.clear();
    for (const auto& h : header_) value_[h] = default_[h];
  }
  void write_row(bool names) {
    if (!out_) return;
    for (size_t i = 0; i < header_.size(); i++) {
      if (names) *out_ << header_[i];
      else *out_ << value_[header_[i]];
      *out_ << (i + 1 < header_.size() ? "\t" : "\n");
    }
    out_->flush();
  }
  std::ostream* out_;
  std::vector<std::string> header_;
  std::map<std::string, double> default_, value_;
};

inline double user_seconds() {  // util.h:71-81
  struct rusage ru;
  getrusage(RUSAGE_SELF, &ru);
  return (double)ru.ru_utime.tv_sec + (double)ru.ru_utime.tv_usec / 1e6;
}

// ---- the learner --------------------------------------------------------------
// fm_learn_sgd_element over N GPUs: rows shard contiguously, one replica of
// w0|w|V per GPU, one NCCL all-reduce + 1/N scale per epoch.
class GpuSgdLearner {
 public:
  HostModel* fm = nullptr;
  double min_target = 0, max_target = 0;
  int task = FMB200_TASK_REGRESSION;
  int num_iter = 100;
  double learn_rate = 0;
  double learn_rates[3] = {0, 0, 0};
  RLog* log = nullptr;
  int mode = FMB200_MODE_HOGWILD;
  int num_gpus = 1;
  int first_device = 0;

  ~GpuSgdLearner() {
#ifdef FMB200_WITH_NCCL
    for (auto c : comms_) ncclCommDestroy(c);
#endif
    for (auto c : ctx_) fmb200_destroy(c);
  }

  static void ck(int rc) {
    if (rc != 0) throw std::string(fmb200_last_error());
  }

  // fm_learn::init + fm_learn_sgd_element::init (fm_learn.h:73-91, fm_learn_sgd_element.h:40-46)
  void init() {
    if (log) {
      const double nan = std::numeric_limits<double>::quiet_NaN();
      if (task == FMB200_TASK_REGRESSION) {
        log->add_field("rmse", nan);
        log->add_field("mae", nan);
      } else {
        log->add_field("accuracy", nan);
      }
      log->add_field("time_pred", nan);
      log->add_field("time_learn", nan);
      log->add_field("time_learn2", nan);
      log->add_field("time_learn4", nan);
      log->add_field("rmse_train", nan);
    }
    if (num_gpus > 1 && mode != FMB200_MODE_HOGWILD)
      throw std::string("-gpus > 1 requires -mode hogwild (the ordered / in-order epoch is one dependency chain)");
    ctx_.resize(num_gpus, nullptr);
    for (int g = 0; g < num_gpus; g++) {
      ck(fmb200_create(&ctx_[g], first_device + g, fm->num_attribute, fm->num_factor, fm->k0, fm->k1));
      ck(fmb200_set_mode(ctx_[g], mode));
    }
    // per-epoch exchange: NVLink peer-memory averaging when the devices can map each
    // other, NCCL otherwise (or when FMB200_CLI_NCCL is set)
    if (num_gpus > 1 && getenv("FMB200_CLI_NCCL") == nullptr) {
      use_peer_ = true;
      for (int g = 0; g < num_gpus && use_peer_; g++)
        if (fmb200_peer_attach_local(ctx_[g], num_gpus, g, ctx_.data()) != 0) use_peer_ = false;
      if (!use_peer_) std::cerr << "note: peer access unavailable (" << fmb200_last_error() << "), using NCCL" << std::endl;
    }
    if (use_peer_) return;
#ifdef FMB200_WITH_NCCL
    if (num_gpus > 1) {
      std::vector<int> devs(num_gpus);
      for (int g = 0; g < num_gpus; g++) devs[g] = first_device + g;
      comms_.resize(num_gpus);
      if (ncclCommInitAll(comms_.data(), num_gpus, devs.data()) != ncclSuccess)
        throw std::string("ncclCommInitAll failed");
    }
#else
    if (num_gpus > 1) throw std::string("this build has no NCCL: -gpus must be 1");
#endif
  }

  void push_state() {
    for (auto c : ctx_) {
      ck(fmb200_set_hparams(c, task, learn_rate, fm->reg0, fm->regw, fm->regv, min_target, max_target));
      ck(fmb200_set_params(c, fm->w0, fm->w.data(), fm->v.data()));
    }
  }
  void pull_state() { ck(fmb200_get_params(ctx_[0], &fm->w0, fm->w.data(), fm->v.data())); }

  // slot 0 = the GPU's train shard, slot 1 = test (GPU 0 only)
  void attach(const SparseData& train, const SparseData& test) {
    const uint64_t n = train.num_cases();
    for (int g = 0; g < num_gpus; g++) {
      const uint64_t lo = n * g / num_gpus, hi = n * (g + 1) / num_gpus;
      std::vector<uint64_t> rp(hi - lo + 1);
      const uint64_t base = train.row_ptr[lo];
      for (uint64_t r = lo; r <= hi; r++) rp[r - lo] = train.row_ptr[r] - base;
      ck(fmb200_upload_data(ctx_[g], 0, hi - lo, rp.back(), rp.data(), train.col.data() + base,
                            train.val.data() + base, train.target.data() + lo));
    }
    ck(fmb200_upload_data(ctx_[0], 1, test.num_cases(), test.num_values(), test.row_ptr.data(),
                          test.col.data(), test.val.data(), test.target.data()));
    n_train_ = n;
    n_test_ = test.num_cases();
  }

  // the body of fm_learn_sgd_element::learn's epoch (fm_learn_sgd_element.h:56-67)
  double epoch() {
    const auto t0 = std::chrono::steady_clock::now();
    for (auto c : ctx_) ck(fmb200_sgd_epoch_async(c, 0));
    if (use_peer_)
      for (auto c : ctx_) ck(fmb200_allreduce_mean(c));
#ifdef FMB200_WITH_NCCL
    if (num_gpus > 1 && !use_peer_) {
      ncclGroupStart();
      for (int g = 0; g < num_gpus; g++) {
        void *buf = nullptr, *st = nullptr;
        uint64_t cnt = 0;
        ck(fmb200_params_device(ctx_[g], &buf, &cnt));
        ck(fmb200_stream(ctx_[g], &st));
        ncclAllReduce(buf, buf, cnt, ncclFloat, ncclSum, comms_[g], (cudaStream_t)st);
      }
      ncclGroupEnd();
      for (auto c : ctx_) ck(fmb200_scale_params(c, 1.0 / num_gpus));
    }
#endif
    for (auto c : ctx_) ck(fmb200_sync(c));
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }

  // fm_learn::evaluate (fm_learn.h:93-153); which = 0 train (all shards), 1 test
  double evaluate(int which) {
    // wall clock: the pass runs on the GPU, user-CPU time (the reference's clock) would read ~0
    const auto t0 = std::chrono::steady_clock::now();
    double sq = 0, ab = 0;
    uint64_t ok = 0;
    const int ng = which == 0 ? num_gpus : 1;
    for (int g = 0; g < ng; g++) {
      double a = 0, b = 0;
      uint64_t c = 0;
      ck(fmb200_evaluate(ctx_[g], which, &a, &b, &c));
      sq += a;
      ab += b;
      ok += c;
    }
    const double n = (double)(which == 0 ? n_train_ : n_test_);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (task == FMB200_TASK_REGRESSION) {
      const double rmse = std::sqrt(sq / n);
      if (log) {
        log->log("rmse", rmse);
        log->log("mae", ab / n);
        log->log("time_pred", dt);
      }
      return rmse;
    }
    const double acc = (double)ok / n;
    if (log) {
      log->log("accuracy", acc);
      log->log("time_pred", dt);
    }
    return acc;
  }

  // fm_learn_sgd::learn + fm_learn_sgd_element::learn (fm_learn_sgd.h:55-65, ..._element.h:48-78)
  void learn() {
    std::cout << "learnrate=" << learn_rate << std::endl;
    std::cout << "learnrates=" << learn_rates[0] << "," << learn_rates[1] << "," << learn_rates[2] << std::endl;
    std::cout << "#iterations=" << num_iter << std::endl;
    std::cout.flush();
    std::cout << "SGD: DON'T FORGET TO SHUFFLE THE ROWS IN TRAINING DATA TO GET THE BEST RESULTS." << std::endl;
    for (int i = 0; i < num_iter; i++) {
      const double dt = epoch();
      const double tr = evaluate(0);
      const double te = evaluate(1);
      std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << tr << "\tTest=" << te << std::endl;
      if (log) {
        log->log("rmse_train", tr);
        log->log("time_learn", dt);
        log->new_line();
      }
    }
  }

  // fm_learn_sgd::predict (fm_learn_sgd.h:76-90) on the test slot
  void predict_test(std::vector<double>& out) {
    out.resize(n_test_);
    ck(fmb200_predict(ctx_[0], 1, 1, out.data()));
  }

  void debug() const {  // fm_learn_sgd.h:71-74 + fm_learn.h:107-111
    std::cout << "num_iter=" << num_iter << std::endl;
    std::cout << "task=" << task << std::endl;
    std::cout << "min_target=" << min_target << std::endl;
    std::cout << "max_target=" << max_target << std::endl;
  }

 private:
  std::vector<fmb200_ctx*> ctx_;
#ifdef FMB200_WITH_NCCL
  std::vector<ncclComm_t> comms_;
#endif
  uint64_t n_train_ = 0, n_test_ = 0;
  bool use_peer_ = false;
};

}  // namespace host
