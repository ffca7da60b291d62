"""Development aid: merged cross-term evaluation of the primary IVC circuit over 2^19 rows (for ncu)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import graph_evaluator_model as G, gpu_util, pyref as R
from mira_b200 import witness as W
from witness_util import pack_program
rows = 1 << 19
for ng in (1, 2):
    progs, meta = G.cross_term_programs(5, ng, R.R_)
    col = lambda seed, n, dist=1: gpu_util.gen_scalars_dev(R.BN254, seed, n, dist)
    fixed = [col(100 + i, rows) for i in range(meta["num_fixed"])]
    dom = W.PlonkEvalDomain(meta["num_advice"], 0, gpu_util.to_bytes(col(300, meta["num_challenges"], 0)), [], fixed,
                            [col(200, meta["num_advice"] * rows, 0)], [col(201, meta["num_advice"] * rows)])
    gp = []
    for p in progs:
        pk = pack_program(p)
        gp.append(W.GraphEvaluator(R.FR, pk["code"], pk["constants"], pk["rotations"], pk["num_intermediates"]))
    outs = [torch.empty(rows * 32, dtype=torch.uint8, device="cuda") for _ in gp]
    for _ in range(3):
        W.evaluate_rows_multi(gp, dom, outs=outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        W.evaluate_rows_multi(gp, dom, outs=outs)
    e1.record(); torch.cuda.synchronize()
    print(ng, "merged ms", round(e0.elapsed_time(e1) / 5, 3), gp[0].stats())
