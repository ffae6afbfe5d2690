"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200sched.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200s_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for s in ("b200s_init", "b200s_snapshot_begin", "b200s_snapshot_commit", "b200s_pods_upload", "b200s_eval",
              "b200s_eval_combined", "b200s_score_batch", "b200s_fetch_scores", "b200s_comm_init"):
        assert s in syms


def test_library_exports_every_declared_symbol(engine_mod):
    lib = engine_mod.load_library()
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/b200sched.h but not exported: {missing}"
    assert sorted(engine_mod.EXPORTS) == header_symbols()
    assert lib.b200s_version() == 100


def test_no_cpu_fallback_without_gpu(engine_mod):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine_mod.B200SError) as ei:
        engine_mod.Engine(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or
    dlopen anything from oracle/."""
    pkg = os.path.join(ROOT, "scheduler-plugins_b200")
    bad = []
    for dp, _, fns in os.walk(pkg):
        if os.sep + "build" in dp or dp.endswith("lib"):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                for line in txt.splitlines():
                    code = line.split("//")[0].split("#")[0] if not fn.endswith(".py") else line.split("#")[0]
                    if re.search(r"liboracle|pyoracle|from oracle|import oracle|oracle/", code):
                        bad.append((fn, line.strip()))
    assert not bad, bad


def test_header_is_plain_c_and_the_example_links():
    """include/b200sched.h must be usable from C (what cgo compiles): examples/cycle.c passes gcc -std=c11 -pedantic
    -Werror, links against the built library, and -- on a box without a GPU -- fails loudly with exit code 2."""
    import subprocess
    import tempfile

    import torch

    src = os.path.join(ROOT, "examples", "cycle.c")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + inc, src], check=True)
    libdir = os.path.join(ROOT, "scheduler-plugins_b200", "lib")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "cycle")
        subprocess.run(["gcc", "-std=c11", "-I" + inc, src, "-L" + libdir, "-lb200sched", "-Wl,-rpath," + libdir, "-o", exe],
                       check=True)
        if not torch.cuda.is_available():
            r = subprocess.run([exe], capture_output=True, text=True)
            assert r.returncode == 2 and "no CPU fallback" in r.stderr
