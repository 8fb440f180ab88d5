#!/usr/bin/env python
"""profiles/r2_traffic.json from the two per-layer ncu --set full captures (tools/ncu_summary.py CSVs).

The captures hold, in launch order, the five kernels of one encoder layer: QKV GEMM, attention, out_proj GEMM, fc1 GEMM,
fc2 GEMM.  bench.py reads this file for `roofline.traffic` (DRAM bytes per launch of the same kernel role).

    python tools/make_traffic_json.py profiles/r2e_vision_layer_summary.csv profiles/r2e_text_layer_summary.csv
"""
import csv
import json
import os
import sys

ROLES = ["gemm[ln1+qkv]", "attention", "gemm[out_proj+resid]", "gemm[ln2+fc1+gelu]", "gemm[fc2+resid]"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = {"source": [os.path.basename(a) for a in sys.argv[1:3]],
           "note": "per launch, ncu --set full --clock-control none (cold-cache, serialised): dram__bytes_read.sum + "
                   "dram__bytes_write.sum, sm__pipe_tensor_cycles_active, gpu__time_duration"}
    for tower, path in zip(("vision", "text"), sys.argv[1:3]):
        rows = list(csv.DictReader(open(path)))
        assert len(rows) == len(ROLES), (path, len(rows))
        for role, r in zip(ROLES, rows):
            out[f"{tower}/{role}"] = {"traffic_mb": round(float(r["dram_read_MB"]) + float(r["dram_write_MB"]), 1),
                                      "dram_read_mb": float(r["dram_read_MB"]), "dram_write_mb": float(r["dram_write_MB"]),
                                      "tensor_active_pct": float(r["tensor_pct"]), "us_under_ncu": float(r["us"]),
                                      "dram_pct_of_peak": float(r["dram_pct"]), "kernel": r["kernel"][:60]}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:600])


if __name__ == "__main__":
    main()
