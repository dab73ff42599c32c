"""Python mirror of the reference's interface for the SGD path, over the C ABI.

Names follow the reference so that the parity tests read like its own code:

  Data                ~ class Data               (reference src/libfm/src/Data.h:47-74)
  FmModel             ~ class fm_model           (src/fm_core/fm_model.h:36-66)
  FmLearnSgdElement   ~ class fm_learn_sgd_element (src/libfm/src/fm_learn_sgd_element.h,
                        fm_learn_sgd.h, fm_learn.h)

All compute happens in libfmb200.so on the GPU; this module only marshals
numpy buffers.  The compiled drop-in command line lives in host/ (C++).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi

TASK_REGRESSION = 0  # fm_learn.h:47
TASK_CLASSIFICATION = 1  # fm_learn.h:48
MODE_INORDER = 0
MODE_HOGWILD = 1
MODE_ORDERED = 2  # sequentially consistent, fp64, parallel over conflict-free runs (fm_ordered.cuh)


class FmError(RuntimeError):
    """The reference throws std::string / const char* (caught at libfm.cpp:436-440)."""


def _p(arr, typ):
    return arr.ctypes.data_as(C.POINTER(typ))


class Data:
    """Row-major sparse design matrix + targets (CSR, SoA)."""

    def __init__(self, row_ptr, col, val, target, num_feature=None):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.target = np.ascontiguousarray(target, dtype=np.float32)
        self.num_cases = int(self.row_ptr.shape[0] - 1)
        if num_feature is None:
            # Data.h:227-229: one more than the largest id seen
            num_feature = int(self.col.max()) + 1 if self.col.size else 0
        self.num_feature = int(num_feature)
        if self.target.size:
            self.min_target = float(self.target.min())  # Data.h:207-208
            self.max_target = float(self.target.max())
        else:
            self.min_target = float(np.finfo(np.float32).max)
            self.max_target = -float(np.finfo(np.float32).max)

    @property
    def num_values(self) -> int:
        return int(self.row_ptr[-1])

    def rows(self, lo, hi) -> "Data":
        """Row shard [lo, hi) -- the multi-GPU partition unit."""
        a, b = int(self.row_ptr[lo]), int(self.row_ptr[hi])
        return Data(self.row_ptr[lo:hi + 1] - self.row_ptr[lo], self.col[a:b], self.val[a:b],
                    self.target[lo:hi], self.num_feature)

    def binarize_targets(self) -> None:
        """libfm.cpp:302-303: classification targets become -1 / +1."""
        self.target = np.where(self.target <= 0.0, -1.0, 1.0).astype(np.float32)

    @staticmethod
    def load(filename: str) -> "Data":
        """libfm text format, as Data::load parses it (Data.h:180-290):
        `target id:value id:value ...`, blank lines and `#` lines skipped."""
        row_ptr = [0]
        col, val, target = [], [], []
        try:
            f = open(filename, "r")
        except OSError:
            raise FmError("unable to open " + filename)
        with f:
            for line in f:
                s = line.strip(" \t\r\n")
                if not s or s[0] == "#":
                    continue
                tok = s.split()
                try:
                    target.append(np.float32(tok[0]))
                    for t in tok[1:]:
                        if t[0] == "#":
                            break
                        i, v = t.split(":")
                        col.append(int(i))
                        val.append(np.float32(v))
                except (ValueError, IndexError):
                    raise FmError('cannot parse line "' + line.rstrip("\n") + '"')
                row_ptr.append(len(col))
        return Data(np.array(row_ptr, dtype=np.uint64), np.array(col, dtype=np.uint32),
                    np.array(val, dtype=np.float32), np.array(target, dtype=np.float32))


def pinned_copy(arr: np.ndarray) -> np.ndarray:
    """Copy `arr` into page-locked memory from fmb200_host_alloc (kept alive by the array)."""
    lib = _capi.load()
    p = C.c_void_p()
    if lib.fmb200_host_alloc(C.byref(p), arr.nbytes) != 0:
        raise FmError(lib.fmb200_last_error().decode())
    buf = (C.c_char * max(arr.nbytes, 1)).from_address(p.value)
    out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size).reshape(arr.shape)
    out[...] = arr
    return out  # never freed explicitly: process-lifetime staging buffers


class _LibcRand:
    """glibc srand()/rand(): the reference's only entropy source (random.h:172-174)."""

    def __init__(self):
        self.libc = C.CDLL(None)
        self.libc.rand.restype = C.c_int
        self.libc.srand.argtypes = [C.c_uint]

    def srand(self, seed: int) -> None:
        self.libc.srand(C.c_uint(seed & 0xFFFFFFFF))

    def uniform(self) -> float:
        return self.libc.rand() / (2147483647.0 + 1.0)

    def gaussian(self) -> float:
        # Leva's ratio-of-uniforms method, random.h:148-162
        while True:
            u = self.uniform()
            while u == 0.0:
                u = self.uniform()
            v = 1.7156 * (self.uniform() - 0.5)
            x = u - 0.449871
            y = abs(v) + 0.386595
            q = x * x + y * (0.19600 * y - 0.25472 * x)
            if q < 0.27597:
                break
            if not ((q > 0.27846) or ((v * v) > (-4.0 * u * u * math.log(u)))):
                break
        return v / u


class FmModel:
    """Host image of the FM parameters; v is factor-major [num_factor][num_attribute]."""

    def __init__(self, num_attribute: int, num_factor: int, k0: bool = True, k1: bool = True):
        self.num_attribute = int(num_attribute)
        self.num_factor = int(num_factor)
        self.k0, self.k1 = bool(k0), bool(k1)
        self.reg0 = self.regw = self.regv = 0.0
        self.init_mean = 0.0
        self.init_stdev = 0.01  # fm_model.h:72
        self.w0 = 0.0
        self.w = np.zeros(self.num_attribute, dtype=np.float64)
        self.v = np.zeros((self.num_factor, self.num_attribute), dtype=np.float64)

    def init(self, seed: int | None = None) -> None:
        """fm_model::init (fm_model.h:91-99): w0 = 0, w = 0, v ~ N(mean, stdev) drawn
        factor-outer / attribute-inner from libc rand() (matrix.h:398-404)."""
        rng = _LibcRand()
        if seed is not None:
            rng.srand(seed)  # libfm.cpp:115-116
        self.w0 = 0.0
        self.w[:] = 0.0
        if self.init_stdev == 0.0 or math.isnan(self.init_stdev):
            self.v[:] = self.init_mean
            return
        flat = self.v.reshape(-1)
        for i in range(flat.shape[0]):
            flat[i] = self.init_mean + self.init_stdev * rng.gaussian()

    def init_numpy(self, seed: int) -> None:
        """Fast non-reference init for large synthetic benchmarks."""
        r = np.random.default_rng(seed)
        self.w0 = 0.0
        self.w[:] = 0.0
        self.v[:] = self.init_mean + self.init_stdev * r.standard_normal(self.v.shape)

    def saveModel(self, path: str) -> None:
        """fm_model::saveModel text layout (fm_model.h:132-154), %g-style numbers."""
        def g(x):
            return "%g" % x
        with open(path, "w") as f:
            if self.k0:
                f.write("#global bias W0\n" + g(self.w0) + "\n")
            if self.k1:
                f.write("#unary interactions Wj\n")
                for i in range(self.num_attribute):
                    f.write(g(self.w[i]) + "\n")
            f.write("#pairwise interactions Vj,f\n")
            for i in range(self.num_attribute):
                f.write(" ".join(g(self.v[q, i]) for q in range(self.num_factor)) + "\n")


class FmLearnSgdElement:
    """fm_learn_sgd_element on a B200: one libfmb200 context (one GPU)."""

    def __init__(self, fm: FmModel, device: int = 0, mode: int = MODE_HOGWILD):
        self.lib = _capi.load()
        self.fm = fm
        self.task = TASK_REGRESSION
        self.learn_rate = 0.0
        self.num_iter = 100  # libfm.cpp:274
        self.min_target = 0.0
        self.max_target = 0.0
        self.mode = mode
        self._ctx = C.c_void_p()
        self._check(self.lib.fmb200_create(C.byref(self._ctx), device, fm.num_attribute,
                                           fm.num_factor, int(fm.k0), int(fm.k1)))
        self._check(self.lib.fmb200_set_mode(self._ctx, mode))
        self._slots = {}
        self.push_params()

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise FmError(self.lib.fmb200_last_error().decode())

    def close(self) -> None:
        if self._ctx:
            self.lib.fmb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mode(self, mode: int) -> None:
        self._check(self.lib.fmb200_set_mode(self._ctx, mode))
        self.mode = mode

    def set_tuning(self, ctas_per_sm=0, rows_per_tile=0, threads=0, damp=0, variant=0) -> None:
        self._check(self.lib.fmb200_set_tuning(self._ctx, ctas_per_sm, rows_per_tile, threads, damp,
                                               variant))

    def push_hparams(self) -> None:
        self._check(self.lib.fmb200_set_hparams(self._ctx, self.task, self.learn_rate, self.fm.reg0,
                                                self.fm.regw, self.fm.regv, self.min_target,
                                                self.max_target))

    def push_params(self) -> None:
        fm = self.fm
        w = np.ascontiguousarray(fm.w, dtype=np.float64)
        v = np.ascontiguousarray(fm.v, dtype=np.float64)
        self._check(self.lib.fmb200_set_params(self._ctx, float(fm.w0), _p(w, C.c_double),
                                               _p(v, C.c_double)))

    def pull_params(self) -> None:
        fm = self.fm
        w0 = C.c_double()
        w = np.empty(fm.num_attribute, dtype=np.float64)
        v = np.empty((fm.num_factor, fm.num_attribute), dtype=np.float64)
        self._check(self.lib.fmb200_get_params(self._ctx, C.byref(w0), _p(w, C.c_double),
                                               _p(v, C.c_double)))
        fm.w0, fm.w, fm.v = w0.value, w, v

    def upload(self, data: Data, slot: int) -> None:
        self._check(self.lib.fmb200_upload_data(
            self._ctx, slot, data.num_cases, data.num_values, _p(data.row_ptr, C.c_uint64),
            _p(data.col, C.c_uint32), _p(data.val, C.c_float), _p(data.target, C.c_float)))
        # the Data object is kept alive with its slot: id() values are reused after garbage
        # collection, and a recycled id must never map to a stale upload
        for key in [k for k, (s, _) in self._slots.items() if s == slot]:
            del self._slots[key]
        self._slots[id(data)] = (slot, data)

    def upload_onehot(self, data: Data, slot: int) -> None:
        """fmb200_upload_onehot: ids + targets only (rows of a fixed width, every value 1)."""
        z = data.num_values // max(data.num_cases, 1)
        if data.num_values != z * data.num_cases or not np.all(data.val == 1.0) or \
                not np.array_equal(data.row_ptr, np.arange(data.num_cases + 1, dtype=np.uint64) * np.uint64(z)):
            raise FmError("upload_onehot needs fixed-width rows with every value 1")
        self._check(self.lib.fmb200_upload_onehot(self._ctx, slot, data.num_cases, z,
                                                  _p(data.col, C.c_uint32), _p(data.target, C.c_float)))
        for key in [k for k, (s, _) in self._slots.items() if s == slot]:
            del self._slots[key]
        self._slots[id(data)] = (slot, data)

    def upload_aos(self, data: Data, slot: int, contiguous: bool = True) -> None:
        """fmb200_upload_data_aos from a host image of the reference's containers
        (sparse_row[] pointing into sparse_entry[]; util/fmatrix.h:34-42)."""
        ent = np.empty(max(data.num_values, 1), dtype=[("id", np.uint32), ("value", np.float32)])
        ent["id"][:data.num_values] = data.col
        ent["value"][:data.num_values] = data.val
        rows = np.zeros(max(data.num_cases, 1), dtype=[("data", np.uint64), ("size", np.uint32), ("pad", np.uint32)])
        sizes = np.diff(data.row_ptr.astype(np.int64)).astype(np.uint32)
        keep = [ent]
        if contiguous:
            rows["data"][:data.num_cases] = ent.ctypes.data + 8 * data.row_ptr[:-1]
        else:  # every row in its own allocation, as a loader without the block would do
            for r in range(data.num_cases):
                a, b = int(data.row_ptr[r]), int(data.row_ptr[r + 1])
                part = ent[a:b].copy()
                keep.append(part)
                rows["data"][r] = part.ctypes.data
        rows["size"][:data.num_cases] = sizes
        self._check(self.lib.fmb200_upload_data_aos(self._ctx, slot, data.num_cases,
                                                    rows.ctypes.data_as(C.c_void_p), _p(data.target, C.c_float)))
        del keep
        for key in [k for k, (s, _) in self._slots.items() if s == slot]:
            del self._slots[key]
        self._slots[id(data)] = (slot, data)

    def release(self, data: Data) -> None:
        """Free the device copy of `data` and its slot."""
        ent = self._slots.pop(id(data), None)
        if ent is not None:
            self._check(self.lib.fmb200_free_data(self._ctx, ent[0]))

    def _slot_of(self, data: Data) -> int:
        if id(data) not in self._slots:
            used = {s for s, _ in self._slots.values()}
            free = [s for s in range(8) if s not in used]
            if not free:
                raise FmError("all 8 data slots are in use: release() one first")
            self.upload(data, free[0])
        return self._slots[id(data)][0]

    # -- the reference's learner surface -----------------------------------
    def sgd_epoch(self, train: Data) -> float:
        """One pass of the row loop, fm_learn_sgd_element.h:56-67.  Returns device seconds."""
        sec = C.c_double()
        self._check(self.lib.fmb200_sgd_epoch(self._ctx, self._slot_of(train), C.byref(sec)))
        return sec.value

    def evaluate(self, data: Data) -> float:
        """fm_learn::evaluate (fm_learn.h:93-153): RMSE or accuracy."""
        sq, ab, ok = C.c_double(), C.c_double(), C.c_uint64()
        self._check(self.lib.fmb200_evaluate(self._ctx, self._slot_of(data), C.byref(sq),
                                             C.byref(ab), C.byref(ok)))
        self.last_mae = ab.value / max(1, data.num_cases)
        if self.task == TASK_REGRESSION:
            return math.sqrt(sq.value / data.num_cases)
        return ok.value / data.num_cases

    def predict(self, data: Data, transform: bool = True) -> np.ndarray:
        """fm_learn_sgd::predict (fm_learn_sgd.h:76-90)."""
        out = np.empty(data.num_cases, dtype=np.float64)
        self._check(self.lib.fmb200_predict(self._ctx, self._slot_of(data), int(transform),
                                            _p(out, C.c_double)))
        return out

    # -- SGDA: fm_learn_sgd_element_adapt_reg ----------------------------------
    def sgda_begin(self, attr_group=None) -> None:
        if attr_group is None:
            self._sgda_groups = 1
            self._check(self.lib.fmb200_sgda_begin(self._ctx, 1, None))
        else:
            g = np.ascontiguousarray(attr_group, dtype=np.uint32)
            self._sgda_groups = int(g.max()) + 1
            self._check(self.lib.fmb200_sgda_begin(self._ctx, self._sgda_groups, _p(g, C.c_uint32)))

    def sgda_epoch(self, train: Data, validation: Data, lambda_steps: bool) -> float:
        sec = C.c_double()
        self._check(self.lib.fmb200_sgda_epoch(self._ctx, self._slot_of(train), self._slot_of(validation),
                                               int(lambda_steps), C.byref(sec)))
        return sec.value

    def sgda_reg(self):
        reg_w = np.zeros(self._sgda_groups)
        reg_v = np.zeros((self._sgda_groups, self.fm.num_factor))
        self._check(self.lib.fmb200_sgda_get_reg(self._ctx, _p(reg_w, C.c_double), _p(reg_v, C.c_double)))
        return reg_w, reg_v

    def mcmc_eterms(self, data: Data) -> np.ndarray:
        """fm_learn_mcmc::predict_data_and_write_to_eterms (fm_learn_mcmc.h:148-378) for one data set."""
        out = np.empty(data.num_cases, dtype=np.float64)
        self._check(self.lib.fmb200_mcmc_eterms(self._ctx, self._slot_of(data), _p(out, C.c_double)))
        return out

    def learn(self, train: Data, test: Data, log=None):
        """fm_learn_sgd_element::learn (fm_learn_sgd_element.h:48-78)."""
        self.push_hparams()
        hist = []
        for i in range(self.num_iter):
            t = self.sgd_epoch(train)
            tr, te = self.evaluate(train), self.evaluate(test)
            hist.append((tr, te, t))
            if log is not None:
                log("#Iter=%3d\tTrain=%g\tTest=%g" % (i, tr, te))
        self.pull_params()
        return hist

    # -- extras for bench / multi-GPU ---------------------------------------
    def params_device(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.fmb200_params_device(self._ctx, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def params_layout(self) -> dict:
        off_w, off_v, ws, kp = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int()
        self._check(self.lib.fmb200_params_layout(self._ctx, C.byref(off_w), C.byref(ws), C.byref(off_v), C.byref(kp)))
        return {"off_w": off_w.value, "ws": ws.value, "off_v": off_v.value, "kp": kp.value,
                "n": self.fm.num_attribute}

    def stream(self) -> int:
        s = C.c_void_p()
        self._check(self.lib.fmb200_stream(self._ctx, C.byref(s)))
        return s.value or 0

    def kernel_launches(self) -> int:
        n = C.c_uint64()
        self._check(self.lib.fmb200_kernel_launches(self._ctx, C.byref(n)))
        return n.value

    def download(self, slot: int) -> Data:
        """The device CSR of a slot, copied back (tests of the upload paths)."""
        nr, nz = C.c_uint64(), C.c_uint64()
        self._check(self.lib.fmb200_download_data(self._ctx, slot, C.byref(nr), C.byref(nz), None, None, None, None))
        rp = np.zeros(nr.value + 1, dtype=np.uint64)
        col = np.zeros(max(nz.value, 1), dtype=np.uint32)
        val = np.zeros(max(nz.value, 1), dtype=np.float32)
        tg = np.zeros(max(nr.value, 1), dtype=np.float32)
        self._check(self.lib.fmb200_download_data(self._ctx, slot, None, None, _p(rp, C.c_uint64),
                                                  _p(col, C.c_uint32), _p(val, C.c_float), _p(tg, C.c_float)))
        return Data(rp, col[:nz.value], val[:nz.value], tg[:nr.value], 0)

    def ordered_index(self, data: Data):
        """(link, rowdep) of fm_ordered.cu for `data` (tests)."""
        link = np.empty(max(data.num_values, 1), dtype=np.uint32)
        rowdep = np.empty(max(data.num_cases, 1), dtype=np.uint32)
        self._check(self.lib.fmb200_ordered_index(self._ctx, self._slot_of(data), _p(link, C.c_uint32),
                                                  _p(rowdep, C.c_uint32)))
        return link[:data.num_values], rowdep[:data.num_cases]

    def epoch_config(self) -> dict:
        v = [C.c_int() for _ in range(7)]
        self._check(self.lib.fmb200_last_epoch_config(self._ctx, *[C.byref(x) for x in v]))
        keys = ["lanes_per_row", "slots", "rows_per_tile", "grid", "block", "smem_bytes", "damp"]
        return dict(zip(keys, [x.value for x in v]))
