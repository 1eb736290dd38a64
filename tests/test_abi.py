"""The C-ABI library loads, exports every symbol include/csolve_b200.h declares, and the device
entry points fail loudly (never fall back to a CPU path) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

import csolve_b200 as cb
import util
from csolve_b200 import instances as I

HEADER = os.path.join(util.ROOT, "include", "csolve_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csolve_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = cb.library()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", cb.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (csolve_\w+)", out))
    assert set(names) <= exported


def test_abi_version():
    assert cb.library().csolve_abi_version() == 5


def test_header_compiles_as_c():
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", cb.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_reference_oracle():
    """nothing under csolve_b200/, include/ or integration/ may use oracle/"""
    for base in ("csolve_b200", "include", "integration"):
        for dp, _, files in os.walk(os.path.join(util.ROOT, base)):
            for f in files:
                if f.endswith((".py", ".c", ".cpp", ".cu", ".cuh", ".h", ".hpp")):
                    text = open(os.path.join(dp, f), errors="replace").read()
                    assert "csolve_oracle" not in text and "liboracle" not in text and "orc_" not in text, os.path.join(dp, f)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_device_is_a_loud_error():
    m = cb.Model(I.queens(4))
    with pytest.raises(cb.CsolveError) as e:
        cb.GpuProblem(m)
    assert e.value.code == -7 and "no CPU fallback" in e.value.message


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_comm_and_group_need_a_device_too():
    """the multi-GPU entry points (csolve_gpu_comm_*, csolve_gpu_group_*) have no CPU fallback either"""
    with pytest.raises(cb.CsolveError) as e:
        cb.Comm(0, 0, 2)
    assert e.value.code == -7
    with pytest.raises(cb.CsolveError) as e:
        cb.GpuGroup(2)
    assert e.value.code == -7
    with pytest.raises(cb.CsolveError) as e:
        cb.device_count()
    assert e.value.code == -7
    with pytest.raises(cb.CsolveError) as e:
        cb.Comm(0, 3, 2)                      # rank outside the world: rejected before any device call
    assert e.value.code == -1


def test_invalid_models_are_rejected_before_touching_the_device():
    """compile_model() validation through the harness (same code csolve_gpu_load runs first)"""
    hc = util.harness_lib()
    m = cb.Model(I.queens(4))
    d = m.flat
    bad = cb.FlatModel.from_buffer_copy(d)
    bad.obj_var = 2            # ALL with an objective variable
    assert hc.hc_load(bad, 1) == -1
    bad = cb.FlatModel.from_buffer_copy(d)
    bad.n_clauses = d.n_clauses - 1
    assert hc.hc_load(bad, 1) == -1
    assert hc.hc_load(d, 1) == 0


def test_ctypes_mirrors_have_the_header_layout(tmp_path):
    """every field of the option / result structs sits where the C header puts it (a C program prints offsetof)"""
    structs = {"csolve_solve_options": cb.host._SolveOptions, "csolve_gpu_result": cb.host._GpuResult,
               "csolve_flat_model": cb.FlatModel}
    lines = []
    for cname, ct in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include "csolve_b200.h"\nint main(void) {\n%s\nreturn 0; }\n' % "\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)


def test_backjump_kernel_is_in_the_library():
    """csolve_solve_options.backjump runs the depth-first phase on an instance of k_search compiled in its own unit
    (csrc/kernels_bj.cu); the instances of kernels.cu are the ones every other test and the bench run"""
    out = subprocess.run(["cuobjdump", "-res-usage", cb.library_path()], capture_output=True, text=True).stdout
    names = re.findall(r"Function (\w+):", out)
    bj = [n for n in names if "csolve_dev_bj" in n]
    assert len(bj) == 1 and "k_searchILb0ELb1ELb0ELb0E" in bj[0], bj
    assert any("csolve_dev8k_searchILb0ELb1ELb0ELb0E" in n for n in names)
