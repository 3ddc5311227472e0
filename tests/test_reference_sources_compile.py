"""'Compiles unchanged', literally: the reference's own cpp/USRP_server_link_threads.cpp -- the file that constructs the
buffer wrappers (:121,136,191,212), calls process() (:666), get() (:584) and close() (:475-521) -- is compiled from where
it lies under /root/reference, unmodified, with include/gsdr_compat.hpp standing in for the four headers whose classes
live behind the C-ABI (tests/cpp/reference_shim/gsdr_for_reference_server.hpp).  UHD / HDF5 / Boost are name-only
stand-ins (oracle/ref_stubs, tests/cpp/thirdparty_stubs): the translation unit is compiled, never linked or run.
The object file must bind to the gsdr C-ABI and to none of the reference's kernels or cuBLAS / cuFFT."""
import os
import shutil
import subprocess

import pytest

from common import ROOT

REF = "/root/reference"
SRC = os.path.join(REF, "cpp", "USRP_server_link_threads.cpp")


@pytest.mark.skipif(not os.path.exists(SRC), reason="reference tree not present (GPU box): compiled in the build container only")
@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_reference_link_threads_compiles_against_the_compat_header(tmp_path):
    obj = tmp_path / "link_threads.o"
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-c", "-w",
           "-include", os.path.join(ROOT, "tests", "cpp", "thirdparty_stubs", "gsdr_more_stubs.hpp"),
           "-include", os.path.join(ROOT, "tests", "cpp", "reference_shim", "gsdr_for_reference_server.hpp"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp", "thirdparty_stubs"),
           "-I", os.path.join(ROOT, "oracle", "ref_stubs"), "-I", os.path.join(REF, "headers"), "-I", cuda_inc,
           SRC, "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-C", str(obj)], capture_output=True, text=True, check=True).stdout
    undefined = {line.split(" U ", 1)[1].strip() for line in syms.splitlines() if " U " in line}
    for need in ("gsdr_rx_create", "gsdr_rx_process", "gsdr_rx_destroy", "gsdr_tx_create", "gsdr_tx_get", "gsdr_tx_destroy",
                 "gsdr_pool_create", "gsdr_pool_get", "gsdr_pool_trash", "gsdr_pool_close"):
        assert need in undefined, f"{need} is not referenced by the reference's link threads"
    # the worker loops themselves are in the object, built from the reference's source
    assert "TXRX::rx_single_link(" in syms and "TXRX::tx_single_link(" in syms
    # nothing of the replaced implementation is pulled in
    for banned in ("cufft", "cublas", "polyphase_filter", "tone_select", "chirp_demodulator", "FIR::"):
        assert not any(banned in u for u in undefined), banned
