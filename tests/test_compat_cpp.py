"""include/gsdr_compat.hpp: a caller written like the reference's TXRX worker threads compiles with
a plain C++11 compiler (no nvcc, no CUDA headers) against libgsdr.so; on a GPU it also runs."""
import os
import subprocess

import pytest

from common import ROOT, has_gpu


def build(tmp_path):
    exe = str(tmp_path / "link_thread_style")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "link_thread_style.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "gpu_sdr_b200"), "-lgsdr", "-Wl,-rpath," + os.path.join(ROOT, "gpu_sdr_b200")])
    return exe


def test_reference_style_caller_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = build(tmp_path)
    if not has_gpu():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "ERROR" in r.stderr  # print_error + exit(-1), never a silent fallback


@pytest.mark.gpu
def test_reference_style_caller_runs(tmp_path, gpu_required):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
