"""
CPU: the C ABI as a C compiler sees it.  include/fdm_b200.h is compiled as plain C99 (gcc -Wall -Werror -pedantic: no
C++, no CUDA, no torch types in it) into a program that prints sizeof / offsetof of every option struct that crosses the
boundary; the ctypes mirrors in jax_fdm_b200/_lib.py must agree field by field, and so must the enum values the Python
side hard-codes.  A field added on one side only shows up here, not as a silent misread on the GPU.
"""
import ctypes as C
import os
import shutil
import subprocess

import pytest

from jax_fdm_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = {"fdm_solve_opts": L.SolveOpts, "fdm_fp_opts": L.FpOpts, "fdm_fa_opts": L.FaOpts, "fdm_goal_program": L.GoalProgram}
ENUMS = {"FDM_FA_Q": L.FA_Q, "FDM_FA_X_FREE": L.FA_X_FREE, "FDM_FA_XYZ": L.FA_XYZ, "FDM_FA_LOSS": L.FA_LOSS,
         "FDM_FA_Q_BAR": L.FA_Q_BAR, "FDM_FA_LOADS_BAR": L.FA_LOADS_BAR, "FDM_FA_XYZ_FIXED_BAR": L.FA_XYZ_FIXED_BAR,
         "FDM_FA_INFO_FORWARD": L.FA_INFO_FORWARD, "FDM_FA_INFO_ADJOINT": L.FA_INFO_ADJOINT, "FDM_FA_LENGTHS": L.FA_LENGTHS,
         "FDM_FA_FORCES": L.FA_FORCES, "FDM_FA_RESIDUALS": L.FA_RESIDUALS, "FDM_FA_VECTORS": L.FA_VECTORS,
         "FDM_FA_XYZ_FIXED": L.FA_XYZ_FIXED, "FDM_FA_LOADS": L.FA_LOADS, "FDM_INFO_STRIDE": L.INFO_STRIDE}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_header_is_plain_c_and_ctypes_mirrors_match(tmp_path):
    lines = ["#include <stddef.h>", "#include <stdio.h>", '#include "fdm_b200.h"', "int main(void) {"]
    for cname, cls in STRUCTS.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname}*)0)->{fname}));')
    for ename in ENUMS:
        lines.append(f'  printf("{ename} %d\\n", (int){ename});')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                    str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = {}
    for line in out.splitlines():
        key, *vals = line.split()
        seen[key] = [int(v) for v in vals]
    for cname, cls in STRUCTS.items():
        assert seen[cname] == [C.sizeof(cls)], (cname, seen[cname], C.sizeof(cls))
        for fname, ftype in cls._fields_:
            assert seen[f"{cname}.{fname}"] == [getattr(cls, fname).offset, C.sizeof(ftype)], (cname, fname)
    for ename, value in ENUMS.items():
        assert seen[ename] == [value], ename


def test_every_struct_field_of_the_header_is_mirrored():
    """The other direction: a field present in the header but missing from the ctypes mirror (same size by luck of
    padding) is caught by comparing the field NAMES parsed from the header."""
    import re
    text = open(os.path.join(ROOT, "include", "fdm_b200.h")).read()
    for cname, cls in STRUCTS.items():
        m = re.search(r"typedef struct " + cname + r" \{(.*?)\} " + cname + ";", text, re.S)
        assert m, cname
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        names = [re.split(r"[\s\*]+", part.strip())[-1]                      # `int32_t a, b;` declares a and b
                 for decl in body.split(";") if decl.strip() for part in decl.split(",")]
        assert names == [f for f, _ in cls._fields_], (cname, names)
