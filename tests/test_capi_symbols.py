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


def test_rust_bindings_are_current_and_complete():
    """integration/crates/mira-b200-sys/src/lib.rs (the `extern "C"` block of the Rust -sys crate) is generated from the
    header: the committed file must be what the generator emits now, and must declare exactly the exported symbols."""
    import sys
    from mira_b200 import _native as N
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_bindings as g
    with open(g.OUT) as f:
        committed = f.read()
    assert committed == g.generate(), "run `python tools/gen_rust_bindings.py`"
    declared = set(re.findall(r"pub fn (mira_\w+)\(", committed))
    assert declared == set(N.SYMBOLS)
    # a few signatures spelled out, so that a parser regression cannot hide behind the self-comparison above
    assert "pub fn mira_msm_commit(ctx: *mut mira_msm_ctx, scalars: *const c_void, n: usize, out_affine: *mut c_void) -> c_int;" in committed
    assert "pub fn mira_msm_ctx_create(curve: c_int, bases: *const c_void, n_bases: usize, bases_on_device: c_int, device: c_int, out: *mut *mut mira_msm_ctx) -> c_int;" in committed
    assert "outs_dev: *const *mut c_void" in committed and "pub selectors: *const *const c_void," in committed
    assert "pub fn mira_msm_ctx_destroy(ctx: *mut mira_msm_ctx);" in committed
