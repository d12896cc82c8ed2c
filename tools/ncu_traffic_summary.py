"""Per-family DRAM traffic from an ncu launch list that carries dram__bytes_read.sum / dram__bytes_write.sum:

    KGAN_NCU_RANGE=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
    python tools/ncu_traffic_summary.py gpurun_out/traffic.csv 1024 [shape, default ntu120] [iterations in the range: KGAN_NCU_STEPS, default 5] > profiles/rN_traffic_b1024.json

KGAN_NCU_RANGE=1 makes bench.py bracket the eager roofline pass (one n_critic cycle: 5 iterations) with cudaProfilerStart/Stop,
so the capture holds exactly the launches that `roofline.algorithmic_bytes_per_launch` averages over.  Families are the ones
ops.py reports (`_run(family, ...)`); bench.py puts `dram_bytes_per_launch` of the dominant family into `roofline.traffic`."""
import csv
import json
import re
import sys
from collections import defaultdict

FAMILY = [
    (r"tapconv_fwd_build_k<(\(bool\))?(1|true)>", "gcn_fused_tf32"),
    (r"tapconv_fwd_tma_k|tapconv_fwd_umma|tapconv_fwd_build_k", "tapconv_fwd_tf32"),
    (r"tapconv_wgrad_tma_k|tapconv_wgrad_umma|tapconv_wgrad_thin", "tapconv_wgrad_tf32"),
    (r"tapconv_fwd_thin|tapconv_fwd_simt", "tapconv_fwd"),
    (r"tapconv_wgrad_simt", "tapconv_wgrad"),
    (r"tapconv_pack", "tapconv_pack"),
    (r"adjmix_bwd_a", "adjmix_bwd_a"),
    (r"adjmix_", "adjmix"),
    (r"plane_spmm", "plane_spmm"),
    (r"chan_reduce", "reduce"),
    (r"bn_", "batchnorm"),
    (r"adam_k", "adam"),
    (r"kgan::", "pointwise"),
    (r"at::|cub::", "torch"),
]


def family(kernel):
    for pat, fam in FAMILY:
        if re.search(pat, kernel):
            return fam
    return "other"


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v


def main():
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per_id = defaultdict(dict)
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        d = per_id[r["ID"]]
        d["kernel"] = r["Kernel Name"]
        m = r["Metric Name"]
        if m == "dram__bytes_read.sum":
            d["rd"] = to_bytes(v, r["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            d["wr"] = to_bytes(v, r["Metric Unit"])
        elif m == "gpu__time_duration.sum":
            d["us"] = to_us(v, r["Metric Unit"])
    fam = defaultdict(lambda: {"launches": 0, "rd": 0.0, "wr": 0.0, "us": 0.0})
    for d in per_id.values():
        a = fam[family(d["kernel"])]
        a["launches"] += 1
        a["rd"] += d.get("rd", 0.0)
        a["wr"] += d.get("wr", 0.0)
        a["us"] += d.get("us", 0.0)
    total_us = sum(a["us"] for a in fam.values())
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over the eager "
                     "roofline pass of bench.py (%s iterations)" % (sys.argv[4] if len(sys.argv) > 4 else "5") + "; per-launch times are cold-cache and serialised",
           "per_gpu_batch": int(sys.argv[2]) if len(sys.argv) > 2 else None, "shape": sys.argv[3] if len(sys.argv) > 3 else "ntu120", "launches": len(per_id), "total_us": total_us, "families": {}}
    for k, a in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        n = a["launches"]
        out["families"][k] = {"launches": n, "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "dram_read_bytes_per_launch": a["rd"] / n,
                              "dram_write_bytes_per_launch": a["wr"] / n, "us_per_launch": a["us"] / n, "share_of_time": a["us"] / total_us}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
