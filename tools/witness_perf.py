"""Phase timing of the witness-side kernels on one GPU (development aid; bench.py --workload ... is the contract).

    python tools/witness_perf.py [log_rows=19]

Workload = the shapes of one SnarkStar fold step: cross-term programs of the primary
(2 x MainGate<5>, 6 terms) and secondary (1 x MainGate<5>, 5 terms) circuits over 2^k rows, W fold over
14*2^k elements, E fold with 6 terms, FFT round trips."""
import json
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch

import graph_evaluator_model as G
import gpu_util
import pyref as R
from mira_b200 import witness as W
from witness_util import pack_program

FR = R.FR
log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 19
rows = 1 << log_rows
HBM = 6456.5


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def column(seed, n, dist=1):
    return gpu_util.gen_scalars_dev(R.BN254, seed, n, dist)


for name, n_gates in (("secondary", 1), ("primary", 2)):
    progs, meta = G.cross_term_programs(5, n_gates, R.R_)
    fixed = [column(100 + i, rows) for i in range(meta["num_fixed"])]
    w1 = [column(200, meta["num_advice"] * rows)]
    w2 = [column(201, meta["num_advice"] * rows)]
    ch = gpu_util.to_bytes(column(300, meta["num_challenges"], 0))
    dom = W.PlonkEvalDomain(meta["num_advice"], 0, ch, [], fixed, w1, w2)
    out = torch.empty(rows * 32, dtype=torch.uint8, device="cuda")
    tot_ms, tot_mul = 0.0, 0
    for k, p in enumerate(progs):
        pk = pack_program(p)
        prog = W.GraphEvaluator(FR, pk["code"], pk["constants"], pk["rotations"], pk["num_intermediates"])
        ms = timed(lambda: prog.evaluate_rows(dom, out=out), reps=3, warm=1)
        st = prog.stats()
        tot_ms += ms
        tot_mul += st["muls"]
        print(json.dumps({"circuit": name, "term": k + 1, "rows": rows, "ms": round(ms, 3), **st,
                          "Gmul_s": round(st["muls"] * rows / ms / 1e6, 1),
                          "wideMAC_T_s": round(st["muls"] * rows * 136 / ms / 1e9, 3)}), flush=True)
    print(json.dumps({"circuit": name, "all_terms_ms": round(tot_ms, 3), "muls_per_row": tot_mul,
                      "Mrows_s": round(rows / tot_ms / 1e3, 2)}), flush=True)
    del fixed, w1, w2

n = 14 * rows
a, b = column(1, n), column(2, n)
r = gpu_util.to_bytes(column(3, 1, 0))
o = torch.empty_like(a)
ms = timed(lambda: W.fold_w(FR, a, b, r, out=o))
print(json.dumps({"kernel": "fold_w", "n": n, "ms": round(ms, 3), "GB_s": round(n * 96 / ms / 1e6, 1), "frac_hbm": round(n * 96 / ms / 1e6 / HBM, 3)}))
e = column(4, rows)
ts = [column(10 + k, rows) for k in range(6)]
oe = torch.empty_like(e)
ms = timed(lambda: W.fold_e(FR, e, ts, r, out=oe))
print(json.dumps({"kernel": "fold_e(6 terms)", "n": rows, "ms": round(ms, 3), "GB_s": round(rows * 8 * 32 / ms / 1e6, 1)}))
del a, b, o
for k in (10, 16, 20, 24):
    x = column(50 + k, 1 << k, 0)
    ms = timed(lambda: W.fft(FR, x, k), reps=3, warm=1)
    nbytes = (1 << k) * 64 * (1 + max(k - 10, 0)) + (1 << k) * 64   # tile pass + one pass per global stage + bitrev
    print(json.dumps({"kernel": "fft", "log_n": k, "ms": round(ms, 3), "alg_GB_s": round(nbytes / ms / 1e6, 1),
                      "Melem_s": round((1 << k) / ms / 1e3, 1)}))

# lookup argument (a6): multiplicities and the two inverse vectors, 2^22 elements (uniform values: all distinct)
k = 22
l_col, t_col = column(70, 1 << k, 0), column(71, 1 << k, 0)
ms = timed(lambda: W.evaluate_m(FR, l_col, t_col), reps=3, warm=1)
print(json.dumps({"kernel": "evaluate_m", "log_n": k, "ms": round(ms, 3), "Melem_s": round((1 << k) / ms / 1e3, 1)}))
m_col = W.evaluate_m(FR, l_col, t_col)
ms = timed(lambda: W.evaluate_h_g(FR, l_col, t_col, r, m_col), reps=3, warm=1)
print(json.dumps({"kernel": "evaluate_h_g", "log_n": k, "ms": round(ms, 3), "Melem_s": round((1 << k) / ms / 1e3, 1),
                  "alg_GB_s": round((1 << k) * 5 * 32 / ms / 1e6, 1)}))
