/* oracle/wavefront_emul.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU emulation of the SCHEDULE of fm_sgd_inorder_wavefront_kernel
 * (libfm_b200/csrc/fm_inorder.cu): 32 "lanes" take the next 32 examples, the hashed
 * conflict test picks the prefix P, the lanes gather and form their addends from the
 * state as it is at the start of the step, the bias chain runs over the P examples,
 * and the lanes then apply fm_SGD from their cached values.  tests/test_oracle.py
 * asserts that this reordering is bit-identical to the sequential loop
 * (fmo_sgd_epoch, i.e. the reference's fm_learn_sgd_element::learn) -- the argument
 * that lets the device kernel claim sequential equivalence.
 *
 * Same arithmetic contract as fm_oracle.c: -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WF_Z 4
#define WF_K 8
#define WF_HASH 2048

static uint32_t wf_hash(uint32_t id) { return (uint32_t)(id * 2654435761u) >> (32 - 11); }

/* returns the number of steps taken (n_rows / steps = mean prefix), or 0 when the
 * shape is not eligible (k > WF_K or a row longer than WF_Z). */
uint64_t fmo_sgd_epoch_wavefront(uint32_t n, int k, int k0, int k1, double* w0p, double* w, double* v,
                                 double lr, double reg0, double regw, double regv, int task,
                                 double min_target, double max_target, uint64_t n_rows,
                                 const uint64_t* row_ptr, const uint32_t* col, const float* val,
                                 const float* target) {
  if (k > WF_K) return 0;
  for (uint64_t r = 0; r < n_rows; r++)
    if (row_ptr[r + 1] - row_ptr[r] > WF_Z) return 0;
  static unsigned int hash[WF_HASH];
  memset(hash, 0, sizeof(hash));
  unsigned int seq = 0;
  uint64_t steps = 0;
  double w0 = k0 ? *w0p : 0.0;
  uint64_t base = 0;
  while (base < n_rows) {
    int size[32], dup[32], valid[32];
    uint32_t id[32][WF_Z];
    double x[32][WF_Z];
    /* (1) */
    if (++seq == (1u << 27)) {
      memset(hash, 0, sizeof(hash));
      seq = 1;
    }
    for (int lane = 0; lane < 32; lane++) {
      uint64_t r = base + lane;
      valid[lane] = r < n_rows;
      size[lane] = valid[lane] ? (int)(row_ptr[r + 1] - row_ptr[r]) : 0;
      dup[lane] = 0;
      for (int j = 0; j < size[lane]; j++) {
        id[lane][j] = col[row_ptr[r] + j];
        x[lane][j] = (double)val[row_ptr[r] + j];
        for (int j2 = 0; j2 < j; j2++)
          if (id[lane][j] == id[lane][j2]) dup[lane] = 1;
      }
    }
    for (int lane = 0; lane < 32; lane++) /* atomicMax of all lanes */
      for (int j = 0; j < size[lane]; j++) {
        unsigned int tag = (seq << 5) | (unsigned int)(31 - lane);
        unsigned int* h = &hash[wf_hash(id[lane][j])];
        if (tag > *h) *h = tag;
      }
    int P = 32;
    for (int lane = 0; lane < 32; lane++) {
      int conflict = 0;
      for (int j = 0; j < size[lane]; j++) {
        unsigned int h = hash[wf_hash(id[lane][j])];
        if (31 - (int)(h & 31u) < lane) conflict = 1;
      }
      if (conflict || !valid[lane]) {
        P = lane;
        break;
      }
    }
    /* (2) gather + addends, all from the state at the start of the step */
    double wv[32][WF_Z], vv[32][WF_Z][WF_K], sum[32][WF_K], add[32][WF_Z + WF_K];
    for (int lane = 0; lane < P; lane++) {
      for (int j = 0; j < size[lane]; j++) {
        wv[lane][j] = k1 ? w[id[lane][j]] : 0.0;
        for (int f = 0; f < k; f++) vv[lane][j][f] = v[(size_t)f * n + id[lane][j]];
      }
      for (int j = 0; j < WF_Z; j++) add[lane][j] = (j < size[lane] && k1) ? wv[lane][j] * x[lane][j] : -0.0;
      for (int f = k; f < WF_K; f++) add[lane][WF_Z + f] = -0.0;
      for (int f = 0; f < k; f++) {
        double sf = 0, ss = 0;
        for (int j = 0; j < size[lane]; j++) {
          double d = vv[lane][j][f] * x[lane][j];
          sf += d;
          ss += d * d;
        }
        sum[lane][f] = sf;
        add[lane][WF_Z + f] = 0.5 * (sf * sf - ss);
      }
    }
    /* (3) chain */
    double mult_of[32];
    for (int t = 0; t < P; t++) {
      double y = (double)target[base + t];
      /* unused slots hold -0.0, the exact identity of IEEE addition; w0 stays +0.0 without bias */
      double pr = 0.0 + w0;
      for (int a = 0; a < WF_Z + WF_K; a++) pr += add[t][a];
      double mult = 0;
      if (task == 0) { /* the kernel's select form of -(y - fmax(min, fmin(max, pr))) */
        double m_lo = -(y - min_target), m_hi = -(y - max_target), m_mid = -(y - pr);
        int hi = !(pr <= max_target);
        int lo = hi ? (max_target < min_target) : (pr < min_target);
        mult = lo ? m_lo : (hi ? m_hi : m_mid);
      } else {
        mult = -y * (1.0 - 1.0 / (1.0 + exp(-y * pr)));
      }
      if (k0) w0 -= lr * (mult + reg0 * w0);
      mult_of[t] = mult;
    }
    /* (4) scatter from the cached values (memory re-read only for rows with a repeated id) */
    for (int lane = 0; lane < P; lane++) {
      double m = mult_of[lane];
      if (k1)
        for (int j = 0; j < size[lane]; j++) {
          double* wi = &w[id[lane][j]];
          double c = dup[lane] ? *wi : wv[lane][j];
          c -= lr * (m * x[lane][j] + regw * c);
          *wi = c;
        }
      for (int f = 0; f < k; f++)
        for (int j = 0; j < size[lane]; j++) {
          double* vp = &v[(size_t)f * n + id[lane][j]];
          double c = dup[lane] ? *vp : vv[lane][j][f];
          double grad = sum[lane][f] * x[lane][j] - c * x[lane][j] * x[lane][j];
          c -= lr * (m * grad + regv * c);
          *vp = c;
        }
    }
    base += (uint64_t)P;
    steps++;
  }
  if (k0) *w0p = w0;
  return steps;
}
