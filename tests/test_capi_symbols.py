"""CPU: the C-ABI library loads and exports every symbol include/fmb200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "fmb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmb200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from libfm_b200 import _capi
    assert _declared() == sorted(_capi.SYMBOLS)


def test_library_exports_every_symbol(built_lib):
    for name in _declared():
        assert hasattr(built_lib, name), name


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        return
    ctx = C.c_void_p()
    rc = built_lib.fmb200_create(C.byref(ctx), 0, 10, 4, 1, 1)
    assert rc != 0 and not ctx.value
    assert b"no CPU path" in built_lib.fmb200_last_error()
