"""ctypes front-ends of the two CPU checkers.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libfm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfm_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "libFM")
REF_CONVERT = os.path.join(HERE, "_ref", "convert")
REF_CLI_B200 = os.path.join(HERE, "_ref", "libFM_b200")  # reference main() + integration/fm_learn_sgd_b200.h


def build() -> None:
    """Compile the C restatement, and the reference shim when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "port", "ref"], check=True,
                   stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _csr(data):
    return (C.c_uint64(data.num_cases), _p(data.row_ptr, C.c_uint64), _p(data.col, C.c_uint32),
            _p(data.val, C.c_float), _p(data.target, C.c_float))


class Port:
    """fm_oracle.c -- state is held here as numpy float64, v factor-major [k][n]."""

    def __init__(self, n, k, k0=True, k1=True):
        if not os.path.exists(PORT_SO):
            build()
        self.lib = C.CDLL(PORT_SO)
        self.lib.fmo_ran_gaussian.restype = C.c_double
        self.lib.fmo_ran_uniform.restype = C.c_double
        self.lib.fmo_predict_row.restype = C.c_double
        self.n, self.k, self.k0, self.k1 = int(n), int(k), int(bool(k0)), int(bool(k1))
        self.w0 = C.c_double(0.0)
        self.w = np.zeros(self.n, dtype=np.float64)
        self.v = np.zeros((self.k, self.n), dtype=np.float64)
        self.reg0 = self.regw = self.regv = 0.0

    def init(self, seed, mean=0.0, stdev=0.1):
        self.lib.fmo_srand(C.c_long(seed))
        self.lib.fmo_init(C.c_uint32(self.n), self.k, C.c_double(mean), C.c_double(stdev),
                          C.byref(self.w0), _p(self.w, C.c_double), _p(self.v, C.c_double))

    def set_params(self, w0, w, v):
        self.w0 = C.c_double(float(w0))
        self.w = np.array(w, dtype=np.float64, copy=True)
        self.v = np.array(v, dtype=np.float64, copy=True).reshape(self.k, self.n)

    def sgd_epoch(self, data, task, lr, min_target, max_target):
        self.lib.fmo_sgd_epoch(C.c_uint32(self.n), self.k, self.k0, self.k1, C.byref(self.w0),
                               _p(self.w, C.c_double), _p(self.v, C.c_double), C.c_double(lr),
                               C.c_double(self.reg0), C.c_double(self.regw), C.c_double(self.regv),
                               task, C.c_double(min_target), C.c_double(max_target), *_csr(data))

    def sgd_epoch_wavefront(self, data, task, lr, min_target, max_target):
        """wavefront_emul.c: the schedule of fm_sgd_inorder_wavefront_kernel; returns its step count
        (0 = shape not eligible, nothing done)."""
        self.lib.fmo_sgd_epoch_wavefront.restype = C.c_uint64
        return self.lib.fmo_sgd_epoch_wavefront(
            C.c_uint32(self.n), self.k, self.k0, self.k1, C.byref(self.w0), _p(self.w, C.c_double),
            _p(self.v, C.c_double), C.c_double(lr), C.c_double(self.reg0), C.c_double(self.regw),
            C.c_double(self.regv), task, C.c_double(min_target), C.c_double(max_target), *_csr(data))

    def evaluate(self, data, task, min_target, max_target):
        sq, ab, ok = C.c_double(), C.c_double(), C.c_uint64()
        self.lib.fmo_evaluate(C.c_uint32(self.n), self.k, self.k0, self.k1, self.w0,
                              _p(self.w, C.c_double), _p(self.v, C.c_double), task,
                              C.c_double(min_target), C.c_double(max_target), *_csr(data),
                              C.byref(sq), C.byref(ab), C.byref(ok))
        return sq.value, ab.value, ok.value

    def metric(self, data, task, min_target, max_target):
        sq, ab, ok = self.evaluate(data, task, min_target, max_target)
        return float(np.sqrt(sq / data.num_cases)) if task == 0 else ok / data.num_cases

    def predict(self, data, task, min_target, max_target, transform=True):
        out = np.empty(data.num_cases, dtype=np.float64)
        self.lib.fmo_predict(C.c_uint32(self.n), self.k, self.k0, self.k1, self.w0,
                             _p(self.w, C.c_double), _p(self.v, C.c_double), task,
                             C.c_double(min_target), C.c_double(max_target), int(transform),
                             C.c_uint64(data.num_cases), _p(data.row_ptr, C.c_uint64),
                             _p(data.col, C.c_uint32), _p(data.val, C.c_float), _p(out, C.c_double))
        return out

    def sgda_begin(self, group=None):
        """state of fm_learn_sgd_element_adapt_reg (:60-90 init, :281-292 learn prologue)"""
        self.group = np.zeros(self.n, dtype=np.uint32) if group is None else np.ascontiguousarray(group, dtype=np.uint32)
        self.n_groups = int(self.group.max()) + 1 if self.n else 1
        self.grad_w = np.zeros(self.n)
        self.grad_v = np.zeros((self.k, self.n))
        self.reg_w = np.zeros(self.n_groups)
        self.reg_v = np.zeros((self.n_groups, max(self.k, 1)))[:, :self.k].copy()
        self.w[:] = 0.0  # :283

    def sgda_epoch(self, train, val, task, lr, min_target, max_target, lambda_steps):
        self.lib.fmo_sgda_epoch(C.c_uint32(self.n), self.k, self.k0, self.k1, C.byref(self.w0),
                                _p(self.w, C.c_double), _p(self.v, C.c_double), _p(self.grad_w, C.c_double),
                                _p(self.grad_v, C.c_double), _p(self.reg_w, C.c_double), _p(self.reg_v, C.c_double),
                                _p(self.group, C.c_uint32), C.c_uint32(self.n_groups), C.c_double(lr), task,
                                C.c_double(min_target), C.c_double(max_target), int(lambda_steps),
                                *_csr(train), *_csr(val))

    def mcmc_eterms(self, data):
        """fm_learn_mcmc::predict_data_and_write_to_eterms (fm_learn_mcmc.h:148-378), one data set."""
        out = np.empty(data.num_cases, dtype=np.float64)
        self.lib.fmo_mcmc_eterms(C.c_uint32(self.n), self.k, self.k0, self.k1, self.w0,
                                 _p(self.w, C.c_double), _p(self.v, C.c_double),
                                 C.c_uint64(data.num_cases), _p(data.row_ptr, C.c_uint64),
                                 _p(data.col, C.c_uint32), _p(data.val, C.c_float), _p(out, C.c_double))
        return out

    def predict_row(self, col, val):
        col = np.ascontiguousarray(col, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float32)
        s = np.zeros(max(self.k, 1))
        ss = np.zeros(max(self.k, 1))
        p = self.lib.fmo_predict_row(C.c_uint32(self.n), self.k, self.k0, self.k1, self.w0,
                                     _p(self.w, C.c_double), _p(self.v, C.c_double),
                                     C.c_uint32(col.size), _p(col, C.c_uint32), _p(val, C.c_float),
                                     _p(s, C.c_double), _p(ss, C.c_double))
        return p, s[:self.k], ss[:self.k]


class Ref:
    """The reference's own fm_model / fm_learn_sgd_element / Data via ref_harness.cpp."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                build()
            if not os.path.exists(REF_SO):
                raise RuntimeError("oracle/_ref/libfm_ref.so unavailable (no /root/reference here "
                                   "and no prebuilt copy)")
            L = C.CDLL(REF_SO)
            L.ref_fm_create.restype = C.c_void_p
            L.ref_fm_create.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_long]
            L.ref_data_from_csr.restype = C.c_void_p
            L.ref_data_load.restype = C.c_void_p
            L.ref_data_load.argtypes = [C.c_char_p]
            L.ref_predict_row.restype = C.c_double
            L.ref_evaluate.restype = C.c_double
            L.ref_last_error.restype = C.c_char_p
            L.ref_ran_gaussian.restype = C.c_double
            L.ref_ran_uniform.restype = C.c_double
            cls._lib = L
        return cls._lib

    def __init__(self, n, k, k0=True, k1=True, init_mean=0.0, init_stdev=0.1, seed=42):
        L = self.lib()
        self.n, self.k = int(n), int(k)
        self.h = C.c_void_p(L.ref_fm_create(self.n, self.k, int(k0), int(k1), init_mean, init_stdev, seed))
        self._data = {}

    def set_reg(self, reg0, regw, regv):
        self.lib().ref_fm_set_reg(self.h, C.c_double(reg0), C.c_double(regw), C.c_double(regv))

    def get_params(self):
        w0 = C.c_double()
        w = np.empty(self.n)
        v = np.empty((self.k, self.n))
        self.lib().ref_fm_get_params(self.h, C.byref(w0), _p(w, C.c_double), _p(v, C.c_double))
        return w0.value, w, v

    def set_params(self, w0, w, v):
        w = np.ascontiguousarray(w, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.lib().ref_fm_set_params(self.h, C.c_double(w0), _p(w, C.c_double), _p(v, C.c_double))

    def data(self, d):
        key = id(d)
        if key not in self._data:
            h = self.lib().ref_data_from_csr(*_csr(d), C.c_int(d.num_feature))
            self._data[key] = (C.c_void_p(h), d)
        return self._data[key][0]

    def learn(self, train, test, task, lr, num_iter, min_target, max_target):
        """fm_learn_sgd_element::learn; returns per-epoch (train_metric, test_metric, time_learn)."""
        tr = np.zeros(num_iter)
        te = np.zeros(num_iter)
        tm = np.zeros(num_iter)
        rc = self.lib().ref_sgd_learn(self.h, self.data(train), self.data(test), task, C.c_double(lr),
                                      num_iter, C.c_double(min_target), C.c_double(max_target),
                                      _p(tr, C.c_double), _p(te, C.c_double), _p(tm, C.c_double))
        if rc != 0:
            raise RuntimeError(self.lib().ref_last_error().decode())
        return tr, te, tm

    def evaluate(self, d, task, min_target, max_target):
        return self.lib().ref_evaluate(self.h, self.data(d), task, C.c_double(min_target),
                                       C.c_double(max_target))

    def predict(self, d, task, min_target, max_target):
        out = np.empty(d.num_cases)
        rc = self.lib().ref_predict(self.h, self.data(d), task, C.c_double(min_target),
                                    C.c_double(max_target), _p(out, C.c_double))
        if rc != 0:
            raise RuntimeError(self.lib().ref_last_error().decode())
        return out

    def sgda_learn(self, train, val, test, group, task, lr, num_iter, min_target, max_target):
        """fm_learn_sgd_element_adapt_reg::learn; returns (reg_w, reg_v); parameters via get_params()"""
        group = np.ascontiguousarray(group, dtype=np.uint32)
        ng = int(group.max()) + 1
        reg_w = np.zeros(ng)
        reg_v = np.zeros((ng, self.k))
        rc = self.lib().ref_sgda_learn(self.h, self.data(train), self.data(val), self.data(test),
                                       _p(group, C.c_uint32), C.c_uint32(ng), task, C.c_double(lr), num_iter,
                                       C.c_double(min_target), C.c_double(max_target), _p(reg_w, C.c_double),
                                       _p(reg_v, C.c_double))
        if rc != 0:
            raise RuntimeError(self.lib().ref_last_error().decode())
        return reg_w, reg_v

    def mcmc_eterms(self, d):
        """the reference's own e-term pass (through its transposed copy of the data)"""
        out = np.empty(d.num_cases)
        rc = self.lib().ref_mcmc_eterms(self.h, self.data(d), _p(out, C.c_double))
        if rc != 0:
            raise RuntimeError(self.lib().ref_last_error().decode())
        return out

    def predict_row(self, col, val):
        col = np.ascontiguousarray(col, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float32)
        s = np.zeros(max(self.k, 1))
        ss = np.zeros(max(self.k, 1))
        p = self.lib().ref_predict_row(self.h, C.c_uint32(col.size), _p(col, C.c_uint32),
                                       _p(val, C.c_float), _p(s, C.c_double), _p(ss, C.c_double))
        return p, s[:self.k], ss[:self.k]

    def save_model(self, path):
        return self.lib().ref_fm_save_model(self.h, path.encode())

    def load_model(self, path):
        return self.lib().ref_fm_load_model(self.h, path.encode())

    @classmethod
    def load_data(cls, filename):
        """Data::load on a file; returns (row_ptr, col, val, target, num_feature, min_t, max_t)."""
        L = cls.lib()
        h = L.ref_data_load(filename.encode())
        if not h:
            raise RuntimeError(L.ref_last_error().decode())
        h = C.c_void_p(h)
        n_rows, nnz, nf = C.c_uint64(), C.c_uint64(), C.c_int()
        mn, mx = C.c_float(), C.c_float()
        L.ref_data_info(h, C.byref(n_rows), C.byref(nnz), C.byref(nf), C.byref(mn), C.byref(mx))
        row_ptr = np.zeros(n_rows.value + 1, dtype=np.uint64)
        col = np.zeros(max(nnz.value, 1), dtype=np.uint32)
        val = np.zeros(max(nnz.value, 1), dtype=np.float32)
        tgt = np.zeros(n_rows.value, dtype=np.float32)
        L.ref_data_to_csr(h, _p(row_ptr, C.c_uint64), _p(col, C.c_uint32), _p(val, C.c_float),
                          _p(tgt, C.c_float))
        return row_ptr, col[:nnz.value], val[:nnz.value], tgt, nf.value, mn.value, mx.value
