"""Host logic of the pageable-input path (point-cloud-utils_b200/csrc/staging.h): the pinned ring's slot protocol --
copy threads, two generation counters per slot running on across jobs, coalesced transfers, lazy slot release --
compiled against a stand-in CUDA runtime (tests/stager_stub/cuda_runtime.h) and stressed on the CPU.  The GPU side
of the same path is covered by tests/test_gpu_full_size.py::test_pageable_inputs_through_the_pinned_ring."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_ring_protocol_delivers_every_byte(tmp_path):
    exe = str(tmp_path / "stager_stress")
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(HERE, "stager_stub"),
           "-I", os.path.join(ROOT, "point-cloud-utils_b200", "csrc"), os.path.join(HERE, "stager_stub", "stress.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    out = subprocess.run([exe, "120"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith("ok")
