"""CPU: the C-ABI shared library loads and exports every symbol include/urmvo_b200.h declares.
No compute call is made (there is no GPU here); creating a context must fail loudly."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "urmvo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(urmvo_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("urmvo_local_ba", "urmvo_local_ba_batch", "urmvo_pose_only_batch", "urmvo_two_view", "urmvo_create",
                 "urmvo_destroy", "urmvo_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import urmvo_b200 as U
    lib = U.load_library()
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(U.EXPORTED_SYMBOLS) == _declared()
    assert lib.urmvo_version() >= 100


def test_signatures_are_plain_c():
    """No torch / C++ types in the boundary: the header must compile as C."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "urmvo_b200.h"\nint main(void){return urmvo_version()==0;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src])


def test_no_cpu_fallback():
    import torch
    import urmvo_b200 as U
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is exercised on the CPU container")
    with pytest.raises(U.UrmvoError, match="no CPU fallback"):
        U.Context(0)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under ur-mvo_b200/ or include/ may reference it."""
    bad = []
    for base in ("ur-mvo_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep) or "lib" in dp.split(os.sep) or "__pycache__" in dp:
                continue
            for f in fn:
                if f.endswith((".cu", ".cuh", ".h", ".cc", ".cpp", ".py", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"pyoracle|liburmvo_oracle|oracle\.h|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
