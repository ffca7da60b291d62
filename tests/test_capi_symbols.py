"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/mira_b200.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mira_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mira_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_python_binds():
    from mira_b200 import _native as N
    assert _declared_symbols() == sorted(N.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from mira_b200 import _native as N
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(N.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mira_b200 import CommitmentKey, CudaError, combine_partials
    with pytest.raises(CudaError):
        CommitmentKey(0, bytes(64 * 4))
    with pytest.raises(CudaError):
        combine_partials(0, bytes(128))


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load, link or call oracle/."""
    pats = [r'#include\s*"[^"]*oracle', r"libmira_oracle", r"oracle_lib", r"\boracle_[a-z0-9_]+\s*\(", r"import\s+oracle",
            r"-lmira_oracle"]
    pkg = os.path.join(ROOT, "mira_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for p in pats:
                    assert not re.search(p, txt), (os.path.join(dirpath, f), p)
