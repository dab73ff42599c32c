"""CPU: the ORDERED epoch kernel's OWN SOURCE (libfm_b200/csrc/fm_ordered.cuh) compiled for the host
through tests/simt/cta_shim.h and run as one OS thread per CUDA thread -- __syncthreads / warp
collectives as pthread barriers, cp.async and TMA bulk copies performed at issue time (the most stale
view the hardware may give), mbarriers as byte counters -- against the sequential oracle
(oracle/fm_oracle.c) on 25 shapes (18 through the single-role driver, 7 through the warp-specialised one; four of them take big
steps so that scores cross the clamps and the bias chain is re-walked): tiles of 1..64 rows, 1..4 warps, k in {0,1,3,8,16,40}, repeated
features inside a row, hot features (runs of one row), classification, no bias.  The kernel
re-associates sums, so the bar is 1e-10 relative on every parameter after two epochs.

What this cannot show is hardware behaviour (memory ordering of cp.async against st.global, device
exp()); the device run is tests/test_ordered_gpu.py."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SIMT = os.path.join(ROOT, "tests", "simt")
OUT = os.path.join(SIMT, "_build")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC),
                                reason="needs g++ and the CUDA headers")


def test_ordered_kernel_source_on_host_threads_matches_oracle():
    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(OUT, "fm_oracle.o")
    subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", "fm_oracle.c"),
                    "-o", obj], check=True)
    exe = os.path.join(OUT, "ordered_host")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-I" + CUDA_INC,
                        os.path.join(SIMT, "ordered_host.cpp"), obj, "-o", exe, "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if " ok" in l]
    assert len(lines) == 25 and "ALL OK" in r.stdout, r.stdout
