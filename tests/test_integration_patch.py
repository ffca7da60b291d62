"""The Rust side of the drop-in is shipped as FILES: integration/mira-b200.patch (edits to the
reference: Cargo feature `b200`, `CommitmentKey::commit`, `GraphEvaluator::to_bytecode`, `commit_cross_terms`, the new
src/b200.rs shim) and integration/crates/mira-b200-sys (FFI declarations generated from include/mira_b200.h).  There is
no rustc in this image, so what can be checked is: the patch applies cleanly to the reference tree, touches the lines it
claims to, calls only functions the header declares, and keeps the reference's tracing span names."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "mira-b200.patch")
REFERENCE = "/root/reference"


def _patch_text():
    with open(PATCH) as f:
        return f.read()


def test_patch_applies_to_the_reference(tmp_path):
    if not os.path.isdir(REFERENCE):
        pytest.skip("the reference tree is not on this box")
    work = tmp_path / "mira"
    shutil.copytree(REFERENCE, work, ignore=shutil.ignore_patterns(".git", "target"))
    subprocess.run(["git", "init", "-q", "."], cwd=work, check=True)
    subprocess.run(["git", "apply", "--check", "--verbose", PATCH], cwd=work, check=True)
    subprocess.run(["git", "apply", PATCH], cwd=work, check=True)
    # the -sys crate goes where the patched Cargo.toml looks for it
    shutil.copytree(os.path.join(ROOT, "integration", "crates", "mira-b200-sys"), work / "crates" / "mira-b200-sys")
    cargo = (work / "Cargo.toml").read_text()
    assert 'b200 = ["dep:mira-b200-sys"]' in cargo and 'path = "crates/mira-b200-sys"' in cargo
    assert (work / "crates" / "mira-b200-sys" / "src" / "lib.rs").exists() and (work / "src" / "b200.rs").exists()
    commitment = (work / "src" / "commitment.rs").read_text()
    i = commitment.index("pub fn commit(&self, v: &[C::Scalar])")
    body = commitment[i:i + 900]
    # the length check stays first, the GPU path sits before the CPU multiexp, which stays as the other curves' path
    assert body.index("self.ck.len() >= v.len()") < body.index("crate::b200::commit::<C>") < body.index("best_multiexp(v")
    assert "TooLongInput" in body
    ge = (work / "src" / "polynomial" / "graph_evaluator.rs").read_text()
    assert "pub(crate) fn to_bytecode(&self) -> Bytecode<F>" in ge
    vanilla = (work / "src" / "nifs" / "vanilla" / "mod.rs").read_text()
    for span in ('info_span!("evaluation")', 'info_span!("commit")', "#[instrument(skip_all)]\n    pub fn commit_cross_terms"):
        assert span in vanilla                      # .scripts/analyze_profiling.py keys on these names
    assert "crate::b200::commit_cross_terms::<C>(ck, &data, &evaluators)?" in vanilla


def test_shim_calls_only_declared_functions_with_the_declared_arity():
    """Every `sys::mira_*` call of src/b200.rs exists in the generated bindings and passes as many arguments as the
    header declares (a cheap stand-in for the type check rustc would do)."""
    from mira_b200 import _native as N
    patch = _patch_text()
    shim = "\n".join(l[1:] for l in patch.split("diff --git a/src/b200.rs")[1].split("diff --git")[0].splitlines() if l.startswith("+"))
    with open(os.path.join(ROOT, "integration", "crates", "mira-b200-sys", "src", "lib.rs")) as f:
        bindings = f.read()
    arity = {m.group(1): (0 if not m.group(2).strip() else m.group(2).count(":"))
             for m in re.finditer(r"pub fn (mira_\w+)\(([^)]*)\)", bindings)}
    calls = list(re.finditer(r"sys::(mira_\w+)\(", shim))
    assert len(calls) >= 12
    for m in calls:
        name = m.group(1)
        assert name in N.SYMBOLS and name in arity, name
        # count top-level commas of the call's argument list
        depth, args, j = 1, 1, m.end()
        empty = True
        while depth:
            ch = shim[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                args += 1
            elif not ch.isspace():
                empty = False
            j += 1
        inner = shim[m.end():j - 1].strip()
        n_args = 0 if not inner else args - (1 if inner.endswith(",") else 0)
        assert n_args == arity[name], (name, n_args, arity[name])
    for const in set(re.findall(r"sys::(MIRA_\w+)", shim)):
        assert f"pub const {const}:" in bindings, const


def test_bytecode_encoding_in_the_patch_matches_the_header():
    """`to_bytecode` (the Rust serialiser) and include/mira_b200.h document the same opcode / operand-kind numbers as
    the Python model the parity tests build programs with."""
    patch = _patch_text()
    for line in ("Calculation::Add(a, b) => (0,", "Calculation::Sub(a, b) => (1,", "Calculation::Mul(a, b) => (2,",
                 "Calculation::Square(a) => (3,", "Calculation::Double(a) => (4,", "Calculation::Negate(a) => (5,",
                 "Calculation::Store(a) => (7,", "ValueSource::Constant(index) => (0u32,", "ValueSource::Intermediate(index) => (1,",
                 "ValueSource::Fixed { index, rotation } => (2,", "ValueSource::Poly { index, rotation } => (3,",
                 "ValueSource::Challenge { index } => (4,"):
        assert line in patch, line
    assert "(6, [start, factor].into_iter().chain(parts.iter()).collect())" in patch      # Horner: start, factor, parts..
    header = open(os.path.join(ROOT, "include", "mira_b200.h")).read()
    assert "0 Add(a,b) 1 Sub(a,b) 2 Mul(a,b) 3 Square(a)" in header and "0 Constant 1 Intermediate 2 Fixed 3 Poly 4 Challenge" in header
    import graph_evaluator_model as G
    ge = G.GraphEvaluator.new(G.Polynomial(0) * G.Polynomial(1) + G.Challenge(0), 97)
    code, ops, i = ge.encode()["code"], [], 0
    while i < len(code):                                                     # walk the records: header, target, operands
        ops.append(code[i] & 0xFF)
        i += 2 + 2 * (code[i] >> 8)
    assert i == len(code) and 2 in ops and 0 in ops and ops[-1] == 7         # a Mul, an Add, and the final Store
