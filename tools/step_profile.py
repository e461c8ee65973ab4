#!/usr/bin/env python
"""Kernel-level breakdown of ONE training step with torch.profiler (CUPTI), for choosing what to fuse next.
    python tools/step_profile.py [--workload config2] [--graph]
"""
import argparse, collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from igcn_b200 import train as T
    from igcn_b200.data import Batch, SubjectSet
    w = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    model, sub, _ = bench.build_problem(w, 0, dev)
    model = model.to(dev).train()
    opt = T.FlatAdam(model.parameters(), lr=1e-3)
    batch = Batch.collate(SubjectSet(sub), np.arange(w["B"]), dev)
    step = lambda: T.train_step(model, batch, opt, bench.LAMBDA)
    for _ in range(3):
        step()
    if a.graph:
        gs = T.GraphedTrainStep(model, opt, batch, bench.LAMBDA)
        step = gs
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = re.sub(r"<.*", "", e.name)[:60]
            agg[k][0] += 1
            agg[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
            tot += e.device_time if hasattr(e, "device_time") else e.cuda_time
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=50))
    ew = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and ("elementwise" in e.name or "reduce_kernel" in e.name):
            m = re.search(r"(\w+Functor\w*|\w+_kernel_cuda\w*|\w+Ops?\b|\w+_cuda\w*)", e.name.split("elementwise_kernel")[-1] if "elementwise" in e.name else e.name)
            k = (m.group(1) if m else e.name[-60:])
            ew[k][0] += 1
            ew[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    print("-- elementwise / reduce kernels by functor")
    for k, (c, t) in sorted(ew.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%5d %9.1f us  %s" % (c, t, k))
    n = sum(c for c, _ in agg.values())
    print("kernels in one step: %d, summed device time %.1f us" % (n, tot))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%5d %9.1f us %5.1f%%  %s" % (c, t, 100 * t / tot, k))


if __name__ == "__main__":
    main()
