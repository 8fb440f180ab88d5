#!/usr/bin/env python
"""Summarise an ncu --set full report (.ncu-rep) into one CSV row per captured launch: duration, DRAM traffic,
tensor-pipe activity, issue activity, registers, top stall reasons.  Runs here (no GPU): `ncu -i` only reads the file.

    python tools/ncu_summary.py gpurun_out/r2b_vision_layer.ncu-rep [out.csv]
"""
import csv
import io
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    out = []
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for r in data:
        d = {"kernel": r[name_i][:90]}
        for k, short in KEYS.items():
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                u = units[hdr.index(k)]
                try:
                    f = float(v)
                    if short == "us":
                        f = f / 1000.0 if u in ("ns", "nsecond") else (f * 1000.0 if u in ("ms", "msecond") else f)
                    if short.endswith("_MB"):
                        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * f
                    d[short] = round(f, 3)
                except ValueError:
                    d[short] = v
        st = []
        for i in stall_cols:
            try:
                st.append((float(r[i]), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        d["top_stalls"] = " ".join(f"{n}:{v:.2f}" for v, n in st[:4])
        out.append(d)
    cols = ["kernel"] + [v for v in KEYS.values()] + ["top_stalls"]
    w = csv.DictWriter(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout, fieldnames=cols, extrasaction="ignore")
    w.writeheader()
    for d in out:
        w.writerow(d)


if __name__ == "__main__":
    main()
