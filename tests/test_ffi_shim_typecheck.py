"""
CPU: csrc/ffi/xla_ffi.cc cannot be BUILT here (the XLA FFI headers come with jaxlib, absent from this image),
but it can be type-checked: tests/ffi_mock/xla/ffi/api/ffi.h is a stand-in whose handler macro static_asserts
that every implementation takes exactly the context / attributes / operands / results its binding declares
(the round-1 defect: `solve` was called with three operands while its binding declared two).  Also checks that
jax_ffi.py registers exactly the handler symbols the shim defines.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "jax_fdm_b200", "csrc", "ffi", "xla_ffi.cc")


def _cuda_include():
    for c in ("/usr/local/cuda/include", os.path.join(os.environ.get("CUDA_HOME", ""), "include")):
        if os.path.exists(os.path.join(c, "cuda_runtime_api.h")):
            return c
    return None


@pytest.mark.skipif(shutil.which("g++") is None or _cuda_include() is None, reason="needs g++ and the CUDA headers")
def test_shim_type_checks_against_the_mock_header():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "ffi_mock"),
           "-I", os.path.join(ROOT, "include"), "-I", _cuda_include(), SHIM]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_python_side_registers_the_symbols_the_shim_defines():
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\(fdm_b200_([a-z_0-9]+),", open(SHIM).read()))
    src = open(os.path.join(ROOT, "jax_fdm_b200", "jax_ffi.py")).read()
    m = re.search(r"_TARGETS = \[(.*?)\]", src, re.S)
    registered = set(re.findall(r'"([a-z_0-9]+)"', m.group(1)))
    assert registered == defined, registered ^ defined
    # every wrapper calls a registered target
    called = set(re.findall(r'self\._call\("([a-z_0-9]+)"', src)) | set(re.findall(r'ffi_call\("fdm_b200_([a-z_0-9]+)"', src))
    assert called <= defined, called - defined
