"""One rank's share of a row-sharded fold step, on ONE GPU (rank 0 of `world`): where do its milliseconds go?
    python tools/fold_rank_profile.py [--world 8] [--log-rows 19] [--reps 20]
Prints the step time (CUDA events, queued back to back) and a per-call breakdown (each call followed by a stream
synchronisation, so launch gaps and host round trips show)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--log-rows", type=int, default=19)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import fold_step as F
    from mira_b200 import witness as W
    g = F.ShardedGpuFoldStep(args.log_rows, 0, args.world)
    for _ in range(3):
        g.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(g.stream)
    for _ in range(args.reps):
        g.step()
    e1.record(g.stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.reps * 1e3
    out = {"world": args.world, "log_rows": args.log_rows, "step_ms_events": round(e0.elapsed_time(e1) / args.reps, 3), "step_ms_wall": round(wall, 3)}
    # per-call breakdown
    sh = g.stream.cuda_stream
    parts = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        g.stream.synchronize()
        parts[name] = parts.get(name, 0.0) + (time.perf_counter() - t) * 1e3
    base = g.partials.data_ptr()
    for _ in range(args.reps):
        j = 0
        with torch.cuda.stream(g.stream):
            for s, st in zip(g.sh, g.state):
                ck = st["ck"]
                nm = s["name"]
                timed(f"{nm}: commit W2 ({st['n_w']} sparse)", lambda: ck.partial_batch_device([st["W2"].data_ptr()], st["n_w"], base + 128 * j, sh))
                j += 1
                dom = W.PlonkEvalDomain(s["meta"]["num_advice"], 0, st["ch"], [], st["fixed"], [st["W1"]], [st["W2"]])
                timed(f"{nm}: evaluate {len(st['T'])} terms", lambda: W.evaluate_rows_multi(st["progs"], dom, outs=st["T"], stream=sh))
                timed(f"{nm}: commit {len(st['T'])} x {st['loc']} (batch)",
                      lambda: ck.partial_batch_device([t.data_ptr() for t in st["T"]], st["loc"], base + 128 * j, sh))
                stt = ck.stats()
                parts[f"{nm}: batch window/launches"] = f"c={stt['window_bits']} launches={stt['kernel_launches']}"
                j += len(st["T"])
                timed(f"{nm}: folds", lambda: (W.fold_w(s["field"], st["W1"], st["W2"], st["r"], out=st["W_out"], stream=sh),
                                               W.fold_e(s["field"], st["E"], st["T"], st["r"], out=st["E_out"], stream=sh)))
    out["per_call_ms"] = {k: (round(v / args.reps, 3) if isinstance(v, float) else v) for k, v in parts.items()}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
