"""GPU: the upload paths that convert layouts on the device (libfm_b200/csrc/fm_upload.cu) -- bit-exact
index / byte work, checked by copying the device CSR back (fmb200_download_data):

 * fmb200_upload_data_aos: the reference's sparse_row[] -> sparse_entry[] containers
   (util/fmatrix.h:34-42; one contiguous block as Data::load allocates it, Data.h:238,260):
   offsets by a device scan, AoS -> SoA split on the device; scattered rows fall back to a host gather;
 * fmb200_upload_onehot: ids + targets only, offsets and values materialised on the device.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import make_learner
from libfm_b200 import MODE_INORDER, Data, FmError, synth

pytestmark = pytest.mark.gpu


def _learner(n, k=4):
    cfg = dict(n=n, k=k, k0=1, k1=1, task=0, lr=0.01, regs=np.zeros(3), min_target=1.0, max_target=5.0)
    r = np.random.default_rng(0)
    return make_learner(cfg, (0.1, r.standard_normal(n) * 0.1, r.standard_normal((k, n)) * 0.1), mode=MODE_INORDER)


def _same(a: Data, b: Data):
    assert np.array_equal(a.row_ptr, b.row_ptr)
    assert np.array_equal(a.col, b.col)
    assert np.array_equal(a.val.view(np.uint32), b.val.view(np.uint32))  # bit pattern, NaN-safe
    assert np.array_equal(a.target.view(np.uint32), b.target.view(np.uint32))


@pytest.mark.parametrize("case", ["ragged", "c2_size", "leading_empty", "all_empty", "one_row", "no_rows"])
def test_aos_upload_device_conversion_is_bit_exact(case, built_lib):
    if case == "ragged":
        d = synth.ragged(5_000, 300, 9, seed=21)
    elif case == "c2_size":  # > 256 scan tiles: the carry loop of the second pass
        d = synth.movielens_1m_shaped(seed=7)
    elif case == "leading_empty":
        d = synth.ragged(3_000, 100, 5, seed=22, empty_frac=0.6)
        d = Data(np.concatenate([np.zeros(40, dtype=np.uint64), d.row_ptr]), d.col, d.val,
                 np.concatenate([np.ones(40, dtype=np.float32), d.target]), 100)
    elif case == "all_empty":
        d = Data(np.zeros(101, dtype=np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32),
                 np.ones(100, np.float32), 10)
    elif case == "one_row":
        d = Data(np.array([0, 3], dtype=np.uint64), np.array([1, 2, 3], np.uint32),
                 np.array([0.5, -1.5, 2.0], np.float32), np.array([4.0], np.float32), 10)
    else:
        d = Data(np.zeros(1, dtype=np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32),
                 np.zeros(0, np.float32), 10)
    l = _learner(max(d.num_feature, 10))
    l.upload_aos(d, 1, contiguous=True)
    _same(l.download(1), d)
    l.close()


def test_aos_upload_scattered_rows_fall_back_to_host_gather(built_lib):
    d = synth.ragged(400, 50, 6, seed=23)
    l = _learner(50)
    l.upload_aos(d, 1, contiguous=False)
    _same(l.download(1), d)
    l.close()


def test_aos_upload_rejects_bad_ids(built_lib):
    d = synth.ragged(400, 50, 6, seed=24)
    l = _learner(20)  # ids up to 49 >= num_attribute 20
    with pytest.raises(FmError, match="out of range"):
        l.upload_aos(d, 1)
    l.close()


@pytest.mark.parametrize("z", [1, 2, 3])
def test_onehot_upload_equals_soa_upload(z, built_lib):
    n_rows, per = 20_000, 300
    r = np.random.default_rng(z)
    cols = (r.integers(0, per, size=(n_rows, z)) + np.arange(z) * per).astype(np.uint32)
    d = Data(np.arange(n_rows + 1, dtype=np.uint64) * np.uint64(z), cols.reshape(-1), np.ones(n_rows * z, np.float32),
             r.integers(1, 6, size=n_rows).astype(np.float32), per * z)
    l = _learner(per * z)
    l.upload_onehot(d, 2)
    _same(l.download(2), d)
    # and it trains exactly like the SoA upload
    l.sgd_epoch(d)          # slot 2 (bound to d by upload_onehot)
    l.pull_params()
    w0a, va = l.fm.w0, l.fm.v.copy()
    m = _learner(per * z)
    m.upload(d, 0)
    m.sgd_epoch(d)
    m.pull_params()
    assert w0a == m.fm.w0 and np.array_equal(va, m.fm.v)
    l.close()
    m.close()


def test_onehot_async_ping_pong(built_lib):
    from libfm_b200.model import pinned_copy
    d = synth.two_field(30_000, 600, 400, seed=3)
    l = _learner(1000)
    ids, tg = pinned_copy(d.col), pinned_copy(d.target)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    for slot in (2, 3, 2):
        assert l.lib.fmb200_upload_onehot_async(l._ctx, slot, d.num_cases, 2, P(ids, C.c_uint32),
                                                P(tg, C.c_float)) == 0
        assert l.lib.fmb200_sgd_epoch(l._ctx, slot, None) == 0, l.lib.fmb200_last_error()
    _same(l.download(2), d)
    _same(l.download(3), d)
    l.close()
