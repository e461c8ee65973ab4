#!/usr/bin/env python
"""Summarises an ncu launch list (CSV written by
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv <cmd>)
into per-kernel calls / time / share, and writes the per-launch DRAM traffic of the igcn kernels that bench.py reports as
`roofline.traffic`.

    python tools/ncu_launches.py launches.csv [--traffic profiles/ncu_traffic_config2.json] [--summary out.txt]
"""
import argparse
import collections
import csv
import json
import re

# kernel function -> C-ABI operation (the tag prefix bench.py uses)
OPS = [("sgcn_fwd_mma", "sgcn_encoder_fwd"), ("sgcn_bwd_mma", "sgcn_encoder_bwd"), ("attn_mma_fwd", "cross_attn_fwd"),
       ("attn_bwd2", "cross_attn_bwd"), ("attn_tables", "cross_attn_tables"), ("attn_chain", "cross_attn_chain"),
       ("go_small_fwd", "go_layer_fwd"), ("go_small_bwd", "go_layer_bwd"),
       ("sgcn_fwd_h16", "sgcn_encoder_fwd"), ("sgcn_encoder_fwd", "sgcn_encoder_fwd"), ("sgcn_bwd_h16", "sgcn_encoder_bwd"),
       ("sgcn_encoder_bwd", "sgcn_encoder_bwd"), ("attn_rows_fwd", "cross_attn_fwd"), ("cross_attn_fwd", "cross_attn_fwd"),
       ("attn_rows_bwd", "cross_attn_bwd"), ("cross_attn_bwd", "cross_attn_bwd"), ("go_layer_fwd", "go_layer_fwd"),
       ("go_layer_bwd", "go_layer_bwd"), ("go_spmm_fwd", "go_spmm_fwd"), ("go_spmm_bwd", "go_spmm_bwd"), ("tc_gemm_kernel", "tc_gemm"),
       ("tc_split", "tc_split"), ("bn_act_fwd", "bn_act_fwd"), ("bn_act_bwd", "bn_act_bwd"), ("collate_kernel", "collate_csr")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic")
    ap.add_argument("--summary")
    a = ap.parse_args()
    lines = [l for l in open(a.csv, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = collections.defaultdict(lambda: dict(calls=set(), ns=0.0, rd=0.0, wr=0.0))
    unit = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name)[:70]
        m, v = r["Metric Name"], float(r["Metric Value"].replace(",", "") or 0)
        unit[m] = r["Metric Unit"]
        d = per[name]
        d["calls"].add(r["ID"])
        if m == "gpu__time_duration.sum":
            d["ns"] += v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1)
        elif m == "dram__bytes_read.sum":
            d["rd"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
        elif m == "dram__bytes_write.sum":
            d["wr"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    tot = sum(d["ns"] for d in per.values())
    out = ["%d launches, %.1f us summed kernel time (ncu: cold caches, serialised)" % (sum(len(d["calls"]) for d in per.values()), tot / 1e3),
           "%6s %10s %6s %12s %12s  kernel" % ("calls", "us", "share", "rd B/launch", "wr B/launch")]
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(d["calls"])
        out.append("%6d %10.1f %5.1f%% %12.0f %12.0f  %s" % (n, d["ns"] / 1e3, 100 * d["ns"] / tot, d["rd"] / n, d["wr"] / n, k))
    text = "\n".join(out)
    print(text)
    if a.summary:
        open(a.summary, "w").write(text + "\n")
    if a.traffic:
        tr = {}
        for k, d in per.items():
            for pat, op in OPS:
                if pat in k:
                    n = len(d["calls"])
                    t = tr.setdefault(op, [0.0, 0])
                    t[0] += d["rd"] + d["wr"]
                    t[1] += n
                    break
        json.dump({op: int(b / max(n, 1)) for op, (b, n) in sorted(tr.items())}, open(a.traffic, "w"), indent=1)


if __name__ == "__main__":
    main()
