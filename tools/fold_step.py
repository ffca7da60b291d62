"""Replay of the hot-path work of ONE SnarkStar IVC fold step (BASELINE.json's second metric, "IVC fold-step ms").

The Rust IVC driver cannot be built in this image (no cargo/rustc; SURVEY.md §7 "hard parts"), so this replays the
exact shapes `IVC::fold_step` sends through the hot path (`cargo re-groth16-dev`, k = 19) with
synthetic data through the same entry points a Rust shim would call:

  per circuit (primary: BN254 / Fr, 2 x MainGate<5> = 14 advice columns, 6 cross terms;
               secondary: Grumpkin / Fq, 1 x MainGate<5> = 7 advice columns, 5 cross terms):
    generate_plonk_trace   ck.commit(W2)                      1 MSM of num_advice * 2^k          (src/plonk/mod.rs:674-707)
    commit_cross_terms     T_j = evaluate(expr_j) over 2^k rows, ck.commit(T_j) for each j        (src/nifs/vanilla/mod.rs:80-140)
    fold                   W = W1 + r*W2 ; E = E + sum_j r^j T_j                                  (src/plonk/mod.rs:1097-1134)

`run_gpu`: everything in HBM; the incoming witness W2 is the only host->device traffic (the e2e leg copies it from
pinned memory inside the timed region; commitments come back as 64 bytes each).
`run_cpu`: the same sequence through the CPU oracle with all host cores (bench.py's cpu_baseline / --impl reference).
Both return the list of commitments so the caller can check they agree bit for bit.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import graph_evaluator_model as G  # noqa: E402  (program construction only: what the Rust side hands over)
import pyref as R  # noqa: E402
from witness_util import pack_program  # noqa: E402

CIRCUITS = [  # name, curve, scalar field id, modulus, MainGate instances
    ("secondary", R.GRUMPKIN, R.FQ, R.P, 1),
    ("primary", R.BN254, R.FR, R.R_, 2),
]
SEED = 0x4D495241


def shapes(log_rows: int):
    out = []
    for name, curve, field, m, n_gates in CIRCUITS:
        progs, meta = G.cross_term_programs(5, n_gates, m)
        out.append({"name": name, "curve": curve, "field": field, "m": m, "progs": [pack_program(p) for p in progs],
                    "muls_per_row": sum(p.counts()["mul"] for p in progs), "meta": meta, "rows": 1 << log_rows,
                    "n_w": meta["num_advice"] << log_rows})
    return out


def points_per_step(sh) -> int:
    return sum(s["n_w"] + len(s["progs"]) * s["rows"] for s in sh)


# ------------------------------------------------------------------------------------------ GPU arm
class GpuFoldStep:
    def __init__(self, log_rows: int, device: int = 0, batch: bool = True):
        import torch
        self.batch = batch
        import gpu_util
        from mira_b200 import CommitmentKey
        from mira_b200 import witness as W
        self.torch, self.W = torch, W
        self.sh = shapes(log_rows)
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.state = []
        for ci, s in enumerate(self.sh):
            rows, n_w, curve = s["rows"], s["n_w"], s["curve"]
            gen = lambda seed, n, dist: gpu_util.gen_scalars_dev(curve, SEED + 1000 * ci + seed, n, dist, device=device)
            bases = gpu_util.gen_bases_dev(curve, SEED + ci, n_w, device=device)
            ck = CommitmentKey(curve, bases, device=device, on_device=True)
            ck.prepare(n_w)
            ck.prepare(rows)
            st = {"ck": ck, "bases": bases,
                  "fixed": [gen(10 + i, rows, 1) for i in range(s["meta"]["num_fixed"])],
                  "W1": gen(1, n_w, 0),            # accumulator (relaxed witness): dense
                  "E": gen(2, rows, 0),
                  "W2": gen(3, n_w, 1),            # incoming witness: sparse, as real advice columns are
                  "W_out": torch.empty(n_w * 32, dtype=torch.uint8, device=f"cuda:{device}"),
                  "E_out": torch.empty(rows * 32, dtype=torch.uint8, device=f"cuda:{device}"),
                  "T": [torch.empty(rows * 32, dtype=torch.uint8, device=f"cuda:{device}") for _ in s["progs"]],
                  "ch": gpu_util.to_bytes(gen(4, s["meta"]["num_challenges"], 0)),
                  "r": gpu_util.to_bytes(gen(5, 1, 0)),
                  "progs": [W.GraphEvaluator(s["field"], p["code"], p["constants"], p["rotations"], p["num_intermediates"])
                            for p in s["progs"]]}
            st["W2_host"] = torch.empty(n_w * 32, dtype=torch.uint8, pin_memory=True)
            st["W2_host"].copy_(st["W2"])
            self.state.append(st)
        torch.cuda.synchronize()

    def step(self, from_host: bool):
        torch, W = self.torch, self.W
        commits = []
        sh = self.stream.cuda_stream
        with torch.cuda.stream(self.stream):
            for s, st in zip(self.sh, self.state):
                ck = st["ck"]
                if from_host:
                    # the witness arrives in (pinned) host memory: the commit pipelines its H2D behind the accumulation,
                    # and evaluation / fold then read the commit's own device copy — W2 crosses PCIe once
                    commits.append(ck.commit(st["W2_host"]))
                    w2 = ck.scalars_device()
                else:
                    w2 = st["W2"]
                    commits.append(ck.commit_device(w2.data_ptr(), s["n_w"], sh))
                dom = W.PlonkEvalDomain(s["meta"]["num_advice"], 0, st["ch"], [], st["fixed"], [st["W1"]], [w2])
                if self.batch:      # all cross terms in one launch, shared sub-products computed once
                    W.evaluate_rows_multi(st["progs"], dom, outs=st["T"], stream=sh)
                else:
                    for prog, t in zip(st["progs"], st["T"]):
                        prog.evaluate_rows(dom, out=t, stream=sh)
                if self.batch:      # all cross-term commitments of the fold in one call
                    commits += ck.commit_batch_device([t.data_ptr() for t in st["T"]], s["rows"], sh)
                else:
                    commits += [ck.commit_device(t.data_ptr(), s["rows"], sh) for t in st["T"]]
                W.fold_w(s["field"], st["W1"], w2, st["r"], out=st["W_out"], stream=sh)
                W.fold_e(s["field"], st["E"], st["T"], st["r"], out=st["E_out"], stream=sh)
        return commits

    def launches_per_step(self) -> int:
        n = 0
        for s, st in zip(self.sh, self.state):
            # every commit reports its own launch count; + one kernel per evaluation and two folds
            n += len(s["progs"]) + 2
        return n

    def host_inputs(self):
        """Everything the CPU arm needs to redo the same step (used by the parity check at small sizes)."""
        import gpu_util
        out = []
        for s, st in zip(self.sh, self.state):
            out.append({k: ([gpu_util.to_bytes(x) for x in st[k]] if isinstance(st[k], list) else gpu_util.to_bytes(st[k]))
                        for k in ("fixed", "W1", "E", "W2", "bases")} | {"ch": st["ch"], "r": st["r"]})
        return out


# ------------------------------------------------------------------------------------------ GPU arm, row-sharded
class ShardedGpuFoldStep:
    """The same fold step with the 2^k rows cut into `world` contiguous ranges, one per rank (cross-term
    evaluation and the fold shard by row range with no exchange; the commitments shard by point range).

    Rank g holds rows [lo, hi) of every column.  Its key shard is laid out column by column,
    `ck[c * 2^k + lo .. c * 2^k + hi)` for c = 0 .. num_advice - 1, so that the local witness (the same rows of each
    advice column, concatenated) commits against the whole shard and a local cross-term vector against its first
    `hi - lo` points — exactly the key prefix the unsharded `ck.commit(T_j)` uses.  MainGate queries `Rotation::cur()`
    only (asserted), so a row needs no neighbour and the evaluator runs on the local columns as a smaller circuit.
    Every commitment becomes a 128-byte XYZZ partial in one device buffer; `gather` is the single collective
    (all_gather of 13 x 128 B per rank) and `combine` folds the ranks' partials on the device."""

    def __init__(self, log_rows: int, rank: int, world: int, device: int = 0):
        import torch
        import gpu_util
        from mira_b200 import CommitmentKey
        from mira_b200 import witness as W
        from mira_b200.sharding import row_shard_key_ranges, shard_range
        self.torch, self.W = torch, W
        self.rank, self.world, self.device = rank, world, device
        self.sh = shapes(log_rows)
        self.stream = torch.cuda.Stream(device=device)
        self.state = []
        self.n_commits = sum(1 + len(s["progs"]) for s in self.sh)
        self.partials = torch.zeros(self.n_commits * 128, dtype=torch.uint8, device=f"cuda:{device}")
        for ci, s in enumerate(self.sh):
            rows, curve, cols = s["rows"], s["curve"], s["meta"]["num_advice"]
            lo, hi = shard_range(rows, world, rank)
            loc = hi - lo
            assert all(r == 0 for p in s["progs"] for r in p["rotations"]), "row sharding without a halo needs Rotation::cur() only"
            gen = lambda seed, n, dist, first=0: gpu_util.gen_scalars_dev(curve, SEED + 1000 * ci + seed, n, dist, first=first, device=device)
            col_major = lambda seed, dist: torch.cat([gen(seed, loc, dist, first=c * rows + lo) for c in range(cols)])
            bases = torch.cat([gpu_util.gen_bases_dev(curve, SEED + ci, k_hi - k_lo, first=k_lo, device=device)
                               for k_lo, k_hi in row_shard_key_ranges(rows, cols, world, rank)])
            ck = CommitmentKey(curve, bases, device=device, on_device=True)
            ck.prepare(cols * loc)
            ck.prepare(loc)
            del bases
            st = {"ck": ck, "loc": loc, "n_w": cols * loc,
                  "fixed": [gen(10 + i, loc, 1, first=lo) for i in range(s["meta"]["num_fixed"])],
                  "W1": col_major(1, 0), "E": gen(2, loc, 0, first=lo), "W2": col_major(3, 1),
                  "W_out": torch.empty(cols * loc * 32, dtype=torch.uint8, device=f"cuda:{device}"),
                  "E_out": torch.empty(loc * 32, dtype=torch.uint8, device=f"cuda:{device}"),
                  "T": [torch.empty(loc * 32, dtype=torch.uint8, device=f"cuda:{device}") for _ in s["progs"]],
                  "ch": gpu_util.to_bytes(gen(4, s["meta"]["num_challenges"], 0)),
                  "r": gpu_util.to_bytes(gen(5, 1, 0)),
                  "progs": [W.GraphEvaluator(s["field"], p["code"], p["constants"], p["rotations"], p["num_intermediates"])
                            for p in s["progs"]]}
            self.state.append(st)
        torch.cuda.synchronize()

    def step(self):
        """Queues the rank's share of the step; returns the device buffer of its n_commits XYZZ partials."""
        torch, W = self.torch, self.W
        sh = self.stream.cuda_stream
        base = self.partials.data_ptr()
        j = 0
        with torch.cuda.stream(self.stream):
            for s, st in zip(self.sh, self.state):
                ck = st["ck"]
                ck.partial_batch_device([st["W2"].data_ptr()], st["n_w"], base + 128 * j, sh)
                j += 1
                dom = W.PlonkEvalDomain(s["meta"]["num_advice"], 0, st["ch"], [], st["fixed"], [st["W1"]], [st["W2"]])
                W.evaluate_rows_multi(st["progs"], dom, outs=st["T"], stream=sh)
                ck.partial_batch_device([t.data_ptr() for t in st["T"]], st["loc"], base + 128 * j, sh)
                j += len(st["T"])
                W.fold_w(s["field"], st["W1"], st["W2"], st["r"], out=st["W_out"], stream=sh)
                W.fold_e(s["field"], st["E"], st["T"], st["r"], out=st["E_out"], stream=sh)
        return self.partials

    def curves(self):
        """Curve of each commitment slot, in the order `step` writes them."""
        return [s["curve"] for s in self.sh for _ in range(1 + len(s["progs"]))]

    def combine(self, gathered, world: int):
        """`gathered`: device tensor [world][n_commits][128] (what all_gather of the ranks' buffers gives).  Returns the
        commitments in the unsharded step's order.  Both circuits' slots are combined by their own curve."""
        from mira_b200 import combine_partials_device
        out, j = [], 0
        with self.torch.cuda.stream(self.stream):
            for s in self.sh:
                cnt = 1 + len(s["progs"])
                out += combine_partials_device(s["curve"], gathered.data_ptr() + 128 * j, world, cnt, self.n_commits * 128,
                                               self.device, self.stream.cuda_stream)
                j += cnt
        return out


# ------------------------------------------------------------------------------------------ CPU arm
class CpuFoldStep:
    """The same step through oracle/ (test infrastructure: only bench.py's CPU legs and tests use this)."""

    def __init__(self, log_rows: int, inputs=None, threads: int = 0):
        import numpy as np
        import oracle_lib as O
        self.O, self.np = O, np
        self.L = O.lib()
        self.L.oracle_eval_rows_mt.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                               C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
        self.L.oracle_eval_rows_mt.restype = C.c_int
        self.L.oracle_fold_w_mt.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        self.L.oracle_fold_e_mt.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        self.threads = threads
        self.sh = shapes(log_rows)
        self.state = []
        for ci, s in enumerate(self.sh):
            rows, n_w, curve = s["rows"], s["n_w"], s["curve"]
            if inputs is None:
                gen = lambda seed, n, dist: O.gen_scalars(curve, SEED + 1000 * ci + seed, n, dist)
                d = {"fixed": [gen(10 + i, rows, 1) for i in range(s["meta"]["num_fixed"])], "W1": gen(1, n_w, 0), "E": gen(2, rows, 0),
                     "W2": gen(3, n_w, 1), "ch": gen(4, s["meta"]["num_challenges"], 0), "r": gen(5, 1, 0)}
            else:
                d = inputs[ci]
            arr = lambda b: np.frombuffer(bytearray(b), dtype=np.uint8)
            st = {"bases": arr(d["bases"]) if "bases" in d else arr(O.gen_bases(curve, SEED + ci, n_w)), "fixed": [arr(x) for x in d["fixed"]], "W1": arr(d["W1"]),
                  "E": arr(d["E"]), "W2": arr(d["W2"]), "ch": arr(d["ch"]), "r": arr(d["r"]),
                  "T": [np.zeros(rows * 32, dtype=np.uint8) for _ in s["progs"]],
                  "W_out": np.zeros(n_w * 32, dtype=np.uint8), "E_out": np.zeros(rows * 32, dtype=np.uint8)}
            self.state.append(st)

    def _commit(self, curve, bases, scalars, n) -> bytes:
        out = C.create_string_buffer(64)
        rc = self.L.oracle_commit(curve, bases.ctypes.data, len(bases) // 64, scalars.ctypes.data, n, self.threads, out)
        assert rc == 0
        return out.raw

    def step(self):
        O, L = self.O, self.L
        commits = []
        for s, st in zip(self.sh, self.state):
            commits.append(self._commit(s["curve"], st["bases"], st["W2"], s["n_w"]))
            fx = (C.c_void_p * len(st["fixed"]))(*[f.ctypes.data for f in st["fixed"]])
            w1 = (C.c_void_p * 1)(st["W1"].ctypes.data)
            w2 = (C.c_void_p * 1)(st["W2"].ctypes.data)
            l1 = (C.c_uint64 * 1)(s["n_w"])
            dom = O.EvalDomainStruct(s["rows"], 0, len(st["fixed"]), s["meta"]["num_advice"], 0, len(st["ch"]) // 32, 1, 1, 0,
                                     None, C.cast(fx, C.c_void_p), C.cast(w1, C.c_void_p), C.cast(l1, C.c_void_p),
                                     C.cast(w2, C.c_void_p), C.cast(l1, C.c_void_p), st["ch"].ctypes.data)
            for p, t in zip(s["progs"], st["T"]):
                code = (C.c_uint32 * len(p["code"]))(*p["code"])
                rots = (C.c_int32 * max(len(p["rotations"]), 1))(*p["rotations"])
                rc = L.oracle_eval_rows_mt(s["field"], code, len(p["code"]), p["constants"], len(p["constants"]) // 32, rots,
                                           len(p["rotations"]), p["num_intermediates"], C.byref(dom), self.threads, t.ctypes.data)
                assert rc == 0
                commits.append(self._commit(s["curve"], st["bases"], t, s["rows"]))
            L.oracle_fold_w_mt(s["field"], st["W1"].ctypes.data, st["W2"].ctypes.data, s["n_w"], st["r"].ctypes.data, self.threads,
                               st["W_out"].ctypes.data)
            ts = (C.c_void_p * len(st["T"]))(*[t.ctypes.data for t in st["T"]])
            L.oracle_fold_e_mt(s["field"], st["E"].ctypes.data, ts, len(st["T"]), s["rows"], st["r"].ctypes.data, self.threads,
                               st["E_out"].ctypes.data)
        return commits


if __name__ == "__main__":   # small self-check on a GPU box: GPU step == CPU step, bit for bit
    import hashlib
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    g = GpuFoldStep(lg)
    got = g.step(from_host=True)
    g.torch.cuda.synchronize()
    cpu = CpuFoldStep(lg, inputs=g.host_inputs())
    t0 = time.time()
    want = cpu.step()
    print(f"cpu step {time.time() - t0:.2f}s; commitments equal: {got == want} ({len(got)} commitments)")
    import gpu_util
    for st_g, st_c in zip(g.state, cpu.state):
        assert gpu_util.to_bytes(st_g["W_out"]) == st_c["W_out"].tobytes() and gpu_util.to_bytes(st_g["E_out"]) == st_c["E_out"].tobytes()
    assert got == want
    print("fold outputs equal; sha256(commitments) =", hashlib.sha256(b"".join(got)).hexdigest())
