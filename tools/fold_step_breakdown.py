"""Development aid: wall-clock breakdown of the fold-step replay's calls (every commit is host-blocking)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import torch
import fold_step as F
g = F.GpuFoldStep(19)
W = g.W
for _ in range(2):
    g.step(False)
torch.cuda.synchronize()
tot = {}
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.time(); r = fn(); torch.cuda.synchronize()
    tot[name] = tot.get(name, 0) + (time.time() - t0) * 1e3
    return r
for rep in range(3):
    sh = g.stream.cuda_stream
    with torch.cuda.stream(g.stream):
        for s, st in zip(g.sh, g.state):
            ck = st["ck"]; nm = s["name"]
            timed(nm + " commit W2", lambda: ck.commit_device(st["W2"].data_ptr(), s["n_w"], sh))
            print(nm, "W2 window", ck.stats()["window_bits"], "launches", ck.stats()["kernel_launches"]) if rep == 0 else None
            dom = W.PlonkEvalDomain(s["meta"]["num_advice"], 0, st["ch"], [], st["fixed"], [st["W1"]], [st["W2"]])
            timed(nm + " eval", lambda: W.evaluate_rows_multi(st["progs"], dom, outs=st["T"], stream=sh))
            timed(nm + " commit T (batch)", lambda: ck.commit_batch_device([t.data_ptr() for t in st["T"]], s["rows"], sh))
            print(nm, "T window", ck.stats()["window_bits"], "launches", ck.stats()["kernel_launches"]) if rep == 0 else None
            timed(nm + " fold", lambda: (W.fold_w(s["field"], st["W1"], st["W2"], st["r"], out=st["W_out"], stream=sh),
                                         W.fold_e(s["field"], st["E"], st["T"], st["r"], out=st["E_out"], stream=sh)))
for k, v in tot.items():
    print(f"{k:32s} {v / 3:7.3f} ms")
print("sum", round(sum(tot.values()) / 3, 3))
