#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top=25]
Prints, per CUDA source line: stall samples, warp instructions executed, dominant stall reasons, shared wavefronts.
"""
import csv
import io
import subprocess
import sys


def _i(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0


def main():
    rep = sys.argv[1]
    kfilter = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fn, path, hdr = None, None, None
    agg = {}
    cur = None
    seen_fn = set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1]
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or (kfilter and (fn is None or kfilter not in fn)):
            continue
        if r[0] != "":
            cur = (fn, path.split("/")[-1], int(r[0]), r[1].strip())
            agg.setdefault(cur, dict(samples=0, inst=0, wave=0, wave_ideal=0, stalls={}))
            continue
        if cur is None:
            continue
        d = dict(zip(hdr, r))
        a = agg[cur]
        a["samples"] += _i(d.get("# Samples"))
        a["inst"] += _i(d.get("Instructions Executed"))
        a["wave"] += _i(d.get("L1 Wavefronts Shared"))
        a["wave_ideal"] += _i(d.get("L1 Wavefronts Shared Ideal"))
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                a["stalls"][k] = a["stalls"].get(k, 0) + _i(v)
    fns = sorted({k[0] for k in agg})
    for f in fns:
        items = [(k, v) for k, v in agg.items() if k[0] == f]
        tot_s = sum(v["samples"] for _, v in items) or 1
        tot_i = sum(v["inst"] for _, v in items) or 1
        print("== %s : %d samples, %d warp-instructions" % (f, tot_s, tot_i))
        for k, v in sorted(items, key=lambda kv: -kv[1]["samples"])[:top]:
            st = ", ".join("%s %d%%" % (n[6:], 100 * c // max(v["samples"], 1)) for n, c in sorted(v["stalls"].items(), key=lambda x: -x[1])[:3])
            print("%5.1f%% smp %5.1f%% inst  wave %8d/%8d  %s:%d  %s   [%s]" % (100.0 * v["samples"] / tot_s, 100.0 * v["inst"] / tot_i,
                  v["wave"], v["wave_ideal"], k[1], k[2], k[3][:70], st))


if __name__ == "__main__":
    main()
