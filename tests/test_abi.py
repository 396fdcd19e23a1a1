"""The C-ABI library loads and exports every symbol include/moc_b200.h declares; without a
CUDA device every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    """names of the functions declared in include/moc_b200.h"""
    text = open(os.path.join(ROOT, "include", "moc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    text = re.sub(r"enum\s*\{.*?\}\s*;", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(names) - {"defined"})


def test_header_and_binding_agree(built):
    assert declared_functions() == sorted(api.EXPORTED)


def test_library_exports_every_declared_symbol(built):
    L = C.CDLL(api.LIB_PATH, mode=C.RTLD_LOCAL)
    for name in declared_functions():
        assert hasattr(L, name), f"libmoc_b200.so lacks {name}"


def test_struct_sizes_are_the_reference_layout(built):
    # sizeof(Input)=152, Params=64 (LP64); checked against the reference itself in
    # tests/test_oracle_vs_ref.py::test_struct_layout_matches_reference
    assert C.sizeof(api.Input) == 152
    assert C.sizeof(api.Params) == 64
    assert C.sizeof(api.CommGrid) == 48


def test_library_is_sm100a_only(built):
    """one architecture, no PTX-JIT fallback path for other GPUs"""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(api.lib().moc_device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback(built):
    host = m.HostProblem(m.derive(m.input_from_values(CASES["tiny"])), seed=1)
    with pytest.raises(m.MocError, match="no usable CUDA device"):
        m.DeviceProblem(host)
    host.close()


def test_bad_arguments_are_reported(built):
    L = api.lib()
    g = api.CommGrid()
    assert L.moc_make_grid(2, 2, 2, 8, C.byref(g)) == -1          # rank out of range
    assert b"bad grid" in L.moc_last_error()
    inp = m.default_input()
    assert L.moc_read_input_file(C.byref(inp), b"/nonexistent/file.in") == -1
    assert b"cannot open" in L.moc_last_error()
    argv = (C.c_char_p * 2)(b"prog", b"-x")
    assert L.moc_read_CLI(2, argv, C.byref(inp)) == -1
    assert b"usage" in L.moc_last_error()
