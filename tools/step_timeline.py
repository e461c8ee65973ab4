#!/usr/bin/env python
"""Device timeline of ONE replay of the captured training step (CUPTI through torch.profiler): every kernel with its stream, start
and duration, sorted by start, plus the idle gaps of the union of all streams.  Used to find the real critical path of the
launch-bound config-2 step.

    python tools/step_timeline.py [--workload config2] > gpurun_out/timeline.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    import bench
    from igcn_b200 import train as T
    from igcn_b200.data import Batch, SubjectSet
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda", 0)
    w = bench.WORKLOADS[a.workload]
    model, sub, _ = bench.build_problem(w, 0, dev)
    model = model.to(dev).train()
    opt = T.FlatAdam(model.parameters(), lr=1e-3)
    batch = Batch.collate(SubjectSet(sub), np.arange(w["B"]), dev)
    g = T.GraphedTrainStep(model, opt, batch, bench.LAMBDA, None, True)
    for _ in range(20):
        g()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            g()
        torch.cuda.synchronize()
    # the chrome trace carries the stream of every kernel (FunctionEvent does not)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        prof.export_chrome_trace(f.name)
        trace = json.load(open(f.name))
    evs = []
    for e in trace["traceEvents"]:
        if e.get("cat") == "kernel":
            evs.append((float(e["ts"]), float(e["dur"]), e["name"], e.get("args", {}).get("stream")))
    evs.sort()
    # three identical replays back to back: the middle third of the kernels, starting at the first kernel of a step
    evs = [e for e in evs if "Memcpy" not in e[2] and "Memset" not in e[2]]
    per = len(evs) // 3
    first = next(i for i in range(per, 2 * per + 1) if "dropout_masks" in evs[i][2] or "counter_inc" in evs[i][2])
    last = evs[first:first + per]
    t0 = last[0][0]
    rows = [dict(start_us=round(s - t0, 2), dur_us=round(d, 2), stream=st, name=n.replace("igcn::", "").replace("void ", "").split("(")[0][:70])
            for s, d, n, st in last]
    end = max(r["start_us"] + r["dur_us"] for r in rows)
    # idle time of the union of all kernels
    iv = sorted((r["start_us"], r["start_us"] + r["dur_us"]) for r in rows)
    idle, cur = [], iv[0][1]
    for s, e in iv[1:]:
        if s > cur + 0.3:
            idle.append((round(cur, 2), round(s - cur, 2)))
        cur = max(cur, e)
    print(json.dumps(dict(workload=a.workload, step_us=round(end, 2), kernels=len(rows), idle_total_us=round(sum(g for _, g in idle), 2),
                          idle_gaps=idle, timeline=rows)))


if __name__ == "__main__":
    main()
