"""CPU checks of the drop-in boundary: libdge.so builds, loads, and exports every symbol include/dge.h
declares; without a GPU it refuses loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dge.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dge_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol(dge_lib):
    syms = header_symbols()
    assert len(syms) >= 25
    L = dge_lib.lib()
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, "declared in include/dge.h but not exported: %s" % missing


def test_header_is_plain_c(tmp_path):
    """The ABI must be bindable from C (JNI glue / Panama): the header compiles as C11 on its own."""
    c = tmp_path / "t.c"
    c.write_text('#include "dge.h"\nint main(void){dge_sgns_params p; (void)p; return DGE_OK;}\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                           "-o", str(tmp_path / "t.o")])


def test_sgns_params_struct_layout_matches_ctypes(dge_lib):
    p = dge_lib.sgns_params()
    assert (p.dim, p.window, p.negative, p.min_count, p.epochs) == (20, 8, 5, 2, 1)
    assert (p.neg_table_size, p.exp_table_size, p.concurrency) == (100000, 1000, 0)
    assert abs(p.lr - 0.025) < 1e-9 and abs(p.min_lr - 1e-4) < 1e-9 and p.seed == 1


def test_flag_constants_match_the_header(dge_lib):
    """The F_* constants of the ctypes binding are the DGE_SGNS_F_* values of include/dge.h."""
    src = open(os.path.join(ROOT, "include", "dge.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    header = {}
    for name, val in re.findall(r"DGE_SGNS_F_([A-Z0-9_]+)\s*=\s*([^,}]+)", src):
        header[name] = eval(val.strip(), {"__builtins__": {}})
    assert len(header) >= 20
    checked = 0
    for name, val in header.items():
        if hasattr(dge_lib, "F_" + name):
            assert getattr(dge_lib, "F_" + name) == val, name
            checked += 1
    assert checked >= 12


def test_no_silent_cpu_fallback(dge_lib):
    """Without a CUDA device dge_create must fail with DGE_E_NO_DEVICE and say so."""
    h = C.c_void_p()
    rc = dge_lib.lib().dge_create(0, C.byref(h))
    if rc == 0:  # a GPU box: nothing to check here
        dge_lib.lib().dge_destroy(h)
        pytest.skip("GPU present")
    assert rc == -2
    msg = dge_lib.lib().dge_last_error(None).decode()
    assert "no CPU fallback" in msg or "sm_100a" in msg
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Context(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under embedding_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "embedding_b200")):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|libdge_oracle|#include\s+\"[^\"]*oracle", txt):
                    bad.append(f)
    assert not bad, bad
