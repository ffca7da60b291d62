#!/usr/bin/env python3
"""Regenerates the fixtures in tests/golden/.  Run in the build container (needs /root/reference for the KAT):

    python tests/golden/make_golden.py

* fft_kat_fr.json   — the known-answer vector of the reference's own test, copied out of
                      /root/reference/src/fft.rs:239-258 (input 0..7, k = 3, BN254 Fr).  REFERENCE-PINNED.
* lagrange_kat_fr.json — /root/reference/src/polynomial/lagrange.rs:113-126 (values of L_i(2), n = 4).  REFERENCE-PINNED.
* commit_vectors.json — commit() outputs of the CPU oracle (oracle/mira_oracle.c) on seeded synthetic inputs.
                      ORACLE-GENERATED (the Rust reference cannot be built here: "parity unpinned" at commit()).
                      They freeze today's oracle so a later change to it, or to the generators, is caught; the GPU
                      path is compared with the same file on the box (where /root/reference does not exist).
* eval_vectors.json — SHA-256 of oracle_eval_rows outputs for the MainGate-shaped cross-term programs.  ORACLE-GENERATED.
"""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import graph_evaluator_model as G  # noqa: E402
import oracle_lib as O  # noqa: E402
import pyref as R  # noqa: E402
from witness_util import Domain, pack_program  # noqa: E402

REF = "/root/reference"


def fft_kat():
    src = open(os.path.join(REF, "src/fft.rs")).read()
    body = src[src.index("fn fft_simple_input_test"):]
    body = body[:body.index(".map(|s| Fr::from_str_vartime")]
    out = re.findall(r'"(\d+)"', body)
    assert len(out) == 8
    return {"source": "src/fft.rs:239-258", "field": "bn256::Fr", "log_n": 3, "input": list(range(8)), "output": out}


def lagrange_kat():
    src = open(os.path.join(REF, "src/polynomial/lagrange.rs")).read()
    i = src.index("mod tests") if "mod tests" in src else 0
    body = src[i:]
    vals = re.findall(r'"(0x[0-9a-fA-F]+|\d{20,})"', body)
    return {"source": "src/polynomial/lagrange.rs:113-126", "raw_strings": vals}


def commit_vectors():
    out = []
    for curve in (R.BN254, R.GRUMPKIN):
        for n, dist in ((1, 0), (7, 0), (64, 1), (1000, 0), (4099, 1)):
            bases = O.gen_bases(curve, 0x4D495241, n)
            sc = O.gen_scalars(curve, 0x4D495242 + n, n, dist)
            out.append({"curve": curve, "n": n, "dist": dist, "seed_bases": 0x4D495241, "seed_scalars": 0x4D495242 + n,
                        "commit_hex": O.commit(curve, bases, sc).hex(),
                        "bases_sha256": hashlib.sha256(bases).hexdigest(), "scalars_sha256": hashlib.sha256(sc).hexdigest()})
    return out


def eval_vectors():
    out = []
    for T, ng in ((5, 1), (5, 2)):
        progs, meta = G.cross_term_programs(T, ng, R.R_)
        d = Domain(R.R_, 64, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=4242 + ng, sparse=True)
        for k, p in enumerate(progs):
            res = O.eval_rows(R.FR, pack_program(p), d.as_bytes())
            out.append({"T": T, "n_gates": ng, "term": k + 1, "rows": 64, "domain_seed": 4242 + ng, "nodes": len(p.calculations),
                        "sha256": hashlib.sha256(res).hexdigest()})
    return out


if __name__ == "__main__":
    for name, fn in (("fft_kat_fr.json", fft_kat), ("lagrange_kat_fr.json", lagrange_kat),
                     ("commit_vectors.json", commit_vectors), ("eval_vectors.json", eval_vectors)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
