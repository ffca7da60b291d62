"""The C++ host-side mirror (include/mira_commitment.hpp) compiled with g++ against libmira_b200.so: on a CPU box it
must fail loudly (CudaError, no fallback); on the GPU box it must reproduce the oracle's commitments and the
reference's TooLongInput behaviour."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "commitment_mirror_test.cpp")


def _build(tmp_path):
    import oracle_lib
    oracle_lib.lib()                                          # makes sure oracle/libmira_oracle.so exists
    from mira_b200 import _native as N
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    exe = str(tmp_path / "commitment_mirror_test")
    libdir, oradir = os.path.join(ROOT, "mira_b200"), os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", oradir, SRC, "-o", exe,
                    "-L", libdir, "-lmira_b200", "-L", oradir, "-lmira_oracle",
                    f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{oradir}"], check=True)
    return exe


def test_cpp_mirror_fails_loudly_without_gpu(tmp_path):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = _build(tmp_path)
    r = subprocess.run([exe, "64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("CudaError") == 2


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle_on_gpu(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = _build(tmp_path)
    r = subprocess.run([exe, "3000", "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("C++ mirror ok") == 2 and "Can't commit too long input: input len: 3001, but limit is 3000" in r.stdout


@pytest.mark.gpu
def test_cpp_witness_mirror_matches_oracle_on_gpu(tmp_path):
    """include/mira_witness.hpp (GraphEvaluator / PlonkEvalDomain / fold / fft) driven from C++ against the oracle."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle_lib
    oracle_lib.lib()
    exe = str(tmp_path / "witness_mirror_test")
    libdir, oradir = os.path.join(ROOT, "mira_b200"), os.path.join(ROOT, "oracle")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", oradir,
                    os.path.join(ROOT, "tests", "cpp", "witness_mirror_test.cpp"), "-o", exe, "-L", libdir, "-lmira_b200", "-L", oradir,
                    "-lmira_oracle", "-Xlinker", f"-rpath={libdir}", "-Xlinker", f"-rpath={oradir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C++ witness mirror ok" in r.stdout and "EvalError" in r.stdout
