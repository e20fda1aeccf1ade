"""Per-kernel-class summary of an ncu launch list of ONE training step.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file launches.csv python tests/prof_step.py 32 1
    python multi-feature-vit_b200/tools/ncu_classes.py launches.csv profiles/r01_kernel_traffic.json

GEMM launches are classed by position (before / after the fusion forward kernel) and epilogue: forward = every GEMM
before the fusion forward (`fusion_fwd_kernel` / `fus2_ln0_kernel`); in the backward the fp32 reduce-add instance is a weight gradient, the rest are dgrads.
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def load(path):
    rows = OrderedDict()
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(rows.values())


def classify(launches):
    seen_fusion = False
    for l in launches:
        n = l["name"]
        if "fusion_fwd_kernel" in n or "fus2_ln0_kernel" in n:  # single-kernel / batched fusion forward
            seen_fusion = True
        if "gemm_bf16_kernel" in n:
            m = re.search(r"gemm_bf16_kernel<\s*(?:\(int\))?(\d+),\s*(?:\(int\))?(\d+),\s*(?:\(int\))?(\d+)(?:,\s*(?:\(int\))?\d+)?>", n)
            epi = int(m.group(3)) if m else -1
            l["kernel"] = "gemm<%s,%s,epi%s>" % (m.group(1), m.group(2), m.group(3)) if m else "gemm"
            l["cls"] = "gemm_fwd" if not seen_fusion else ("gemm_wgrad" if epi == 4 else "gemm_dgrad")
        elif "attn_fwd" in n:
            l["cls"], l["kernel"] = "attn_fwd", n.split("(")[0].split("::")[-1]
        elif "attn_bwd" in n or "attn_delta" in n or "attn_dq_convert" in n:
            l["cls"], l["kernel"] = "attn_bwd", n.split("(")[0].split("::")[-1]
        elif "ln_fwd" in n:
            l["cls"], l["kernel"] = "ln_fwd", "ln_fwd"
        elif "ln_bwd" in n:
            l["cls"], l["kernel"] = "ln_bwd", "ln_bwd"
        elif "fusion_" in n or "fus2_" in n:
            l["cls"], l["kernel"] = "fusion", n.split("(")[0].split("::")[-1].split("<")[0]
        elif "sgd_kernel" in n or "adam_kernel" in n:
            l["cls"], l["kernel"] = "sgd", "sgd"
        else:
            l["cls"], l["kernel"] = "other", n.split("(")[0].split("::")[-1].split("<")[0]
    return launches


def main():
    src, out = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    launches = classify(load(src))
    total = sum(l["gpu__time_duration.sum"] for l in launches) / 1e3
    cls = defaultdict(lambda: {"n": 0, "us": 0.0, "bytes": 0.0})
    ker = defaultdict(lambda: {"n": 0, "us": 0.0})
    for l in launches:
        c = cls[l["cls"]]
        c["n"] += 1
        c["us"] += l["gpu__time_duration.sum"] / 1e3
        c["bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        k = ker[l["cls"] + " " + l["kernel"]]
        k["n"] += 1
        k["us"] += l["gpu__time_duration.sum"] / 1e3
    print("%d launches, %.1f us serialised" % (len(launches), total))
    print("| class | launches | us / step | share | DRAM MB / launch |\n|---|---|---|---|---|")
    for name, c in sorted(cls.items(), key=lambda kv: -kv[1]["us"]):
        print("| %s | %d | %.0f | %.1f %% | %.1f |" % (name, c["n"], c["us"], 100 * c["us"] / total, c["bytes"] / c["n"] / 1e6))
    print()
    for name, k in sorted(ker.items(), key=lambda kv: -kv[1]["us"]):
        print("%-46s x%-3d %8.1f us  avg %6.1f us" % (name, k["n"], k["us"], k["us"] / k["n"]))
    if out:
        doc = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control "
                         "none --profile-from-start off, python tests/prof_step.py 32 1 (one MF-ViT CA step, 32 pairs); "
                         "serialised, cold-cache launches",
               "total_us": round(total, 1),
               "classes": {n: {"launches_per_step": c["n"], "us_per_step": round(c["us"], 1),
                               "share": round(c["us"] / total, 4), "dram_bytes_per_launch": int(c["bytes"] / c["n"])}
                           for n, c in sorted(cls.items(), key=lambda kv: -kv[1]["us"])}}
        with open(out, "w") as f:
            json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
