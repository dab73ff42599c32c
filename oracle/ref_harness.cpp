// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI shim around the UNMODIFIED reference (srendle/libfm) headers, which
// are compiled IN PLACE from /root/reference/src (never copied into this repo).
// It exists so the parity tests and the CPU-baseline leg of bench.py can drive
// the reference's own classes (fm_model, fm_learn_sgd_element, Data) on
// in-memory CSR inputs and read back full-precision state.  Output:
// oracle/_ref/libfm_ref.so (git-ignored; travels to the GPU box with gpurun).
//
// Nothing under libfm_b200/ may link or dlopen this file's output.
//
// Reference entry points exercised (all in /root/reference/src):
//   fm_model::init / predict                 fm_core/fm_model.h:91-127
//   fm_SGD                                   fm_core/fm_sgd.h:33-51
//   fm_learn_sgd_element::learn              libfm/src/fm_learn_sgd_element.h:48-78
//   fm_learn::evaluate                       libfm/src/fm_learn.h:93-153
//   fm_learn_sgd::predict                    libfm/src/fm_learn_sgd.h:76-90
//   Data::load                               libfm/src/Data.h:113-290
//   fm_model::saveModel / loadModel          fm_core/fm_model.h:132-190
//   fm_learn_mcmc::predict_data_and_write_to_eterms  libfm/src/fm_learn_mcmc.h:148-378
//   Data::create_data_t                      libfm/src/Data.h:292-337
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <iterator>
#include <algorithm>
#include <iomanip>
#include <stdint.h>

// include order matters: the reference headers are not self-contained
#include "util/util.h"
#include "fm_core/fm_model.h"
#include "libfm/src/Data.h"
#include "libfm/src/fm_learn.h"
#include "libfm/src/fm_learn_sgd.h"
#include "libfm/src/fm_learn_sgd_element.h"
#include "libfm/src/fm_learn_sgd_element_adapt_reg.h"
#include "libfm/src/fm_learn_mcmc_simultaneous.h"

namespace {

// silence the reference's chatter on std::cout while a call is in flight
struct CoutMute {
  std::streambuf* saved;
  std::ostringstream sink;
  CoutMute() { saved = std::cout.rdbuf(sink.rdbuf()); }
  ~CoutMute() { std::cout.rdbuf(saved); }
};

struct RefData {
  Data* d;
  sparse_entry<DATA_FLOAT>* entries;  // owned when built from CSR
  RefData() : d(NULL), entries(NULL) {}
};

struct RefFm {
  fm_model fm;
  DataMetaInfo* meta;
};

char g_err[512];

template <typename F> int guarded(F f) {
  try {
    f();
    return 0;
  } catch (std::string& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.c_str());
  } catch (char const*& e) {
    snprintf(g_err, sizeof(g_err), "%s", e);
  } catch (...) {
    snprintf(g_err, sizeof(g_err), "unknown exception");
  }
  return 1;
}

void setup_learner(fm_learn_sgd_element& l, RefFm* m, int task, double lr, int num_iter,
                   double min_target, double max_target, RLog* log) {
  l.fm = &m->fm;
  l.meta = m->meta;
  l.task = task;
  l.min_target = min_target;
  l.max_target = max_target;
  l.num_iter = num_iter;
  l.learn_rate = lr;
  l.log = log;
  l.init();
  l.learn_rates.init(lr);
  if (log != NULL) log->init();
}

// Data::create_data_t is protected: reached through a derived view of the same object
struct DataProbe : public Data {
  DataProbe() : Data(0, true, false) {}
  void make_t() { create_data_t(); }
};

// the e-term pass is a protected member: a derived probe exposes it (nothing is modified)
struct McmcProbe : public fm_learn_mcmc_simultaneous {
  void eterms(DVector<Data*>& d, DVector<e_q_term*>& c) { predict_data_and_write_to_eterms(d, c); }
};

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err; }

// fm_learn_sgd_element_adapt_reg::learn (SGDA) with the given attribute groups; returns the learned
// regularisation values (reg_w [n_groups], reg_v [n_groups][k]); the model stays in the fm handle.
int ref_sgda_learn(void* fm_h, void* train_h, void* val_h, void* test_h, const uint32_t* group, uint32_t n_groups,
                   int task, double lr, int num_iter, double min_target, double max_target, double* reg_w,
                   double* reg_v) {
  RefFm* m = (RefFm*)fm_h;
  RefData* tr = (RefData*)train_h;
  RefData* va = (RefData*)val_h;
  RefData* te = (RefData*)test_h;
  return guarded([&]() {
    CoutMute mute;
    DataMetaInfo meta(m->fm.num_attribute);
    for (uint i = 0; i < m->fm.num_attribute; i++) meta.attr_group(i) = group[i];
    meta.num_attr_groups = n_groups;
    meta.num_relations = 0;
    fm_learn_sgd_element_adapt_reg l;
    l.fm = &m->fm;
    l.meta = &meta;
    l.task = task;
    l.min_target = min_target;
    l.max_target = max_target;
    l.num_iter = num_iter;
    l.learn_rate = lr;
    l.log = NULL;
    l.validation = va->d;
    l.init();
    l.learn_rates.init(lr);
    l.learn(*tr->d, *te->d);
    for (uint g = 0; g < n_groups; g++) {
      reg_w[g] = l.reg_w(g);
      for (int f = 0; f < m->fm.num_factor; f++) reg_v[(size_t)g * m->fm.num_factor + f] = l.reg_v(g, f);
    }
  });
}

// fm_learn_mcmc::predict_data_and_write_to_eterms on one data set (no relations): e_out[c] is the
// e-term = the model's score of case c, accumulated feature-major through the transposed copy
// exactly as the MCMC/ALS learner does once per iteration (fm_learn_mcmc_simultaneous.h:69,122).
int ref_mcmc_eterms(void* fm_h, void* data_h, double* e_out) {
  RefFm* m = (RefFm*)fm_h;
  RefData* r = (RefData*)data_h;
  return guarded([&]() {
    CoutMute mute;
    if (r->d->data_t == NULL) static_cast<DataProbe*>(r->d)->make_t();
    McmcProbe l;
    l.fm = &m->fm;
    l.meta = m->meta;
    l.task = 0;
    l.log = NULL;
    DVector<Data*> main_data(1);
    DVector<e_q_term*> main_cache(1);
    e_q_term* cache = new e_q_term[r->d->num_cases > 0 ? r->d->num_cases : 1];
    main_data(0) = r->d;
    main_cache(0) = cache;
    l.eterms(main_data, main_cache);
    for (uint c = 0; c < r->d->num_cases; c++) e_out[c] = cache[c].e;
    delete[] cache;
  });
}

// srand(seed) then fm_model::init(): the exact draw order of libfm.cpp:115-116,245-257
void* ref_fm_create(uint32_t n_attr, int k, int k0, int k1, double init_mean, double init_stdev,
                    long seed) {
  RefFm* m = new RefFm();
  srand(seed);
  m->fm.num_attribute = n_attr;
  m->fm.init_mean = init_mean;
  m->fm.init_stdev = init_stdev;
  m->fm.k0 = k0 != 0;
  m->fm.k1 = k1 != 0;
  m->fm.num_factor = k;
  m->fm.init();
  m->meta = new DataMetaInfo(n_attr);
  m->meta->num_relations = 0;
  return m;
}

void ref_fm_destroy(void* h) {
  RefFm* m = (RefFm*)h;
  delete m->meta;
  delete m;
}

void ref_fm_set_reg(void* h, double reg0, double regw, double regv) {
  RefFm* m = (RefFm*)h;
  m->fm.reg0 = reg0;
  m->fm.regw = regw;
  m->fm.regv = regv;
}

// v is factor-major [k][n] exactly as DMatrixDouble stores it (matrix.h:152-175)
void ref_fm_get_params(void* h, double* w0, double* w, double* v) {
  RefFm* m = (RefFm*)h;
  *w0 = m->fm.w0;
  memcpy(w, m->fm.w.value, sizeof(double) * m->fm.num_attribute);
  memcpy(v, m->fm.v.value[0], sizeof(double) * (size_t)m->fm.num_attribute * m->fm.num_factor);
}

void ref_fm_set_params(void* h, double w0, const double* w, const double* v) {
  RefFm* m = (RefFm*)h;
  m->fm.w0 = w0;
  memcpy(m->fm.w.value, w, sizeof(double) * m->fm.num_attribute);
  memcpy(m->fm.v.value[0], v, sizeof(double) * (size_t)m->fm.num_attribute * m->fm.num_factor);
}

int ref_fm_save_model(void* h, const char* path) {
  RefFm* m = (RefFm*)h;
  return guarded([&]() { m->fm.saveModel(path); });
}

int ref_fm_load_model(void* h, const char* path) {
  RefFm* m = (RefFm*)h;
  return m->fm.loadModel(path);  // 1 = ok, 0 = malformed (fm_model.h:160-190)
}

// Build a Data object holding a LargeSparseMatrixMemory over one contiguous
// sparse_entry[] block -- the same in-memory shape Data::load produces for text
// input (Data.h:180,238,260).
void* ref_data_from_csr(uint64_t n_rows, const uint64_t* row_ptr, const uint32_t* col,
                        const float* val, const float* target, int num_feature) {
  RefData* r = new RefData();
  r->d = new Data(0, true, false);
  LargeSparseMatrixMemory<DATA_FLOAT>* mat = new LargeSparseMatrixMemory<DATA_FLOAT>();
  r->d->data = mat;
  uint64_t nnz = row_ptr[n_rows];
  r->entries = new sparse_entry<DATA_FLOAT>[nnz > 0 ? nnz : 1];
  for (uint64_t j = 0; j < nnz; j++) {
    r->entries[j].id = col[j];
    r->entries[j].value = val[j];
  }
  mat->data.setSize(n_rows);
  for (uint64_t i = 0; i < n_rows; i++) {
    mat->data.value[i].data = r->entries + row_ptr[i];
    mat->data.value[i].size = (uint)(row_ptr[i + 1] - row_ptr[i]);
  }
  mat->num_cols = num_feature;
  mat->num_values = nnz;
  r->d->target.setSize(n_rows);
  r->d->min_target = +std::numeric_limits<DATA_FLOAT>::max();
  r->d->max_target = -std::numeric_limits<DATA_FLOAT>::max();
  for (uint64_t i = 0; i < n_rows; i++) {
    r->d->target.value[i] = target[i];
    r->d->min_target = std::min(target[i], r->d->min_target);
    r->d->max_target = std::max(target[i], r->d->max_target);
  }
  r->d->num_feature = num_feature;
  r->d->num_cases = n_rows;
  return r;
}

// Data::load on a libfm text file or a convert-produced binary pair
void* ref_data_load(const char* filename) {
  RefData* r = new RefData();
  r->d = new Data(0, true, false);
  CoutMute mute;
  if (guarded([&]() { r->d->load(filename); })) {
    delete r;
    return NULL;
  }
  return r;
}

void ref_data_info(void* h, uint64_t* n_rows, uint64_t* nnz, int* num_feature, float* min_target,
                   float* max_target) {
  RefData* r = (RefData*)h;
  *n_rows = r->d->data->getNumRows();
  *nnz = r->d->data->getNumValues();
  *num_feature = r->d->num_feature;
  *min_target = r->d->min_target;
  *max_target = r->d->max_target;
}

// dump what the loader produced, walking the reference's own row cursor
void ref_data_to_csr(void* h, uint64_t* row_ptr, uint32_t* col, float* val, float* target) {
  RefData* r = (RefData*)h;
  LargeSparseMatrix<DATA_FLOAT>* x = r->d->data;
  uint64_t pos = 0;
  row_ptr[0] = 0;
  for (x->begin(); !x->end(); x->next()) {
    sparse_row<DATA_FLOAT>& row = x->getRow();
    for (uint j = 0; j < row.size; j++) {
      col[pos] = row.data[j].id;
      val[pos] = row.data[j].value;
      pos++;
    }
    row_ptr[x->getRowIndex() + 1] = pos;
    target[x->getRowIndex()] = r->d->target(x->getRowIndex());
  }
}

// classification target remap, libfm.cpp:302-303
void ref_data_binarize_targets(void* h) {
  RefData* r = (RefData*)h;
  for (uint i = 0; i < r->d->target.dim; i++) {
    r->d->target(i) = (r->d->target(i) <= 0.0) ? -1.0 : 1.0;
  }
}

void ref_data_destroy(void* h) {
  RefData* r = (RefData*)h;
  delete[] r->entries;  // Data itself never frees (reference leaks by design)
  delete r;
}

// fm_model::predict for one row (fm_model.h:105-127); also returns sum/sum_sqr
double ref_predict_row(void* hm, uint32_t size, const uint32_t* col, const float* val, double* sum,
                       double* sum_sqr) {
  RefFm* m = (RefFm*)hm;
  std::vector<sparse_entry<FM_FLOAT> > e(size > 0 ? size : 1);
  for (uint32_t j = 0; j < size; j++) {
    e[j].id = col[j];
    e[j].value = val[j];
  }
  sparse_row<FM_FLOAT> row;
  row.data = &e[0];
  row.size = size;
  DVector<double> s(m->fm.num_factor), ss(m->fm.num_factor);
  double p = m->fm.predict(row, s, ss);
  for (int f = 0; f < m->fm.num_factor; f++) {
    if (sum) sum[f] = s(f);
    if (sum_sqr) sum_sqr[f] = ss(f);
  }
  return p;
}

// one fm_SGD step on one row with an explicit multiplier (fm_sgd.h:33-51)
void ref_sgd_row(void* hm, double lr, uint32_t size, const uint32_t* col, const float* val,
                 double mult, const double* sum) {
  RefFm* m = (RefFm*)hm;
  std::vector<sparse_entry<FM_FLOAT> > e(size > 0 ? size : 1);
  for (uint32_t j = 0; j < size; j++) {
    e[j].id = col[j];
    e[j].value = val[j];
  }
  sparse_row<FM_FLOAT> row;
  row.data = &e[0];
  row.size = size;
  DVector<double> s(m->fm.num_factor);
  for (int f = 0; f < m->fm.num_factor; f++) s(f) = sum[f];
  fm_SGD(&m->fm, lr, row, mult, s);
}

// fm_learn_sgd_element::learn for num_iter epochs.  Per epoch the reference
// itself logs rmse_train and time_learn to its RLog; we hand it a string stream
// and parse the columns back.  out_* may be NULL; each has num_iter slots.
int ref_sgd_learn(void* hm, void* htrain, void* htest, int task, double lr, int num_iter,
                  double min_target, double max_target, double* out_train_metric,
                  double* out_test_metric, double* out_time_learn) {
  RefFm* m = (RefFm*)hm;
  RefData* tr = (RefData*)htrain;
  RefData* te = (RefData*)htest;
  std::ostringstream rl;
  return guarded([&]() {
    CoutMute mute;
    RLog log(&rl);
    fm_learn_sgd_element l;
    setup_learner(l, m, task, lr, num_iter, min_target, max_target, &log);
    l.learn(*tr->d, *te->d);
    // parse TSV: header line then num_iter lines
    std::istringstream in(rl.str());
    std::string line;
    std::getline(in, line);
    std::vector<std::string> hdr = tokenize(line, "\t");
    int c_train = -1, c_test = -1, c_time = -1;
    for (size_t i = 0; i < hdr.size(); i++) {
      if (hdr[i] == "rmse_train") c_train = i;
      if (hdr[i] == "time_learn") c_time = i;
      // after learn() the generic metric column holds the TEST value
      // (last evaluate() wins, fm_learn_sgd_element.h:69-70)
      if (hdr[i] == "rmse" || hdr[i] == "accuracy") c_test = i;
    }
    for (int it = 0; it < num_iter; it++) {
      if (!std::getline(in, line)) throw std::string("rlog short");
      std::vector<std::string> f = tokenize(line, "\t");
      if (out_train_metric) out_train_metric[it] = atof(f[c_train].c_str());
      if (out_test_metric) out_test_metric[it] = atof(f[c_test].c_str());
      if (out_time_learn) out_time_learn[it] = atof(f[c_time].c_str());
    }
  });
}

// fm_learn::evaluate: RMSE (task 0) or accuracy (task 1), full double precision
double ref_evaluate(void* hm, void* hdata, int task, double min_target, double max_target) {
  RefFm* m = (RefFm*)hm;
  RefData* d = (RefData*)hdata;
  CoutMute mute;
  fm_learn_sgd_element l;
  setup_learner(l, m, task, 0.0, 0, min_target, max_target, NULL);
  return l.evaluate(*d->d);
}

// fm_learn_sgd::predict: clamped score (task 0) or sigmoid (task 1) per row
int ref_predict(void* hm, void* hdata, int task, double min_target, double max_target,
                double* out) {
  RefFm* m = (RefFm*)hm;
  RefData* d = (RefData*)hdata;
  return guarded([&]() {
    CoutMute mute;
    fm_learn_sgd_element l;
    setup_learner(l, m, task, 0.0, 0, min_target, max_target, NULL);
    DVector<double> pred;
    pred.setSize(d->d->num_cases);
    l.predict(*d->d, pred);
    memcpy(out, pred.value, sizeof(double) * d->d->num_cases);
  });
}

// the reference's RNG primitives, for pinning the oracle's restatement
void ref_srand(long seed) { srand(seed); }
double ref_ran_gaussian() { return ran_gaussian(); }
double ref_ran_uniform() { return ran_uniform(); }

}  // extern "C"
