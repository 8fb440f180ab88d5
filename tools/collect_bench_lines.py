#!/usr/bin/env python
"""Collect the bench.py JSON lines of the round's gpurun calls (gpurun_out/r2*.json[l]) into profiles/r2_bench_lines.json,
trimmed to the fields the docs quote (the full lines of the final run are kept whole)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LABELS = {
    "r2a_bench": "start of round 2 (round-1 kernels, new bench contract)", "r2a_bench_ref": "reference arm, start of round",
    "r2a_bench_cfg3": "cfg3, start of round", "r2a_pdl1": "PLIP_PDL=1, 200 steps", "r2a_pdl0": "PLIP_PDL=0, 200 steps",
    "r2b_bench_v2": "attention v2", "r2b_bench_v1": "PLIP_ATT_V1=1 (round-1 attention), same box",
    "r2b_bench_v2c3": "attention v2, 3 CTAs / SM (128 registers)", "r2b_bench_fc2bn192": "fc2 with 192-wide N tiles",
    "r2d_bench_bf16": "packed-fp32 epilogues, bf16 operands", "r2d_bench_fp16": "packed-fp32 epilogues, fp16 operands",
    "r2d_bench_bf16_b": "packed-fp32 epilogues, bf16 operands (repeat)",
    "r2e_bench": "per-warp bias slices + graphs + tensor-core similarity (full extras)", "r2e_bench_mb512": "PLIP_BENCH_MB=512",
    "r2e_bench_nograph": "PLIP_GRAPH_MAX=0", "r2e_bench_cfg5_125k": "cfg5 with a 125k-tile gallery on one GPU (= one rank of N=8)",
    "r2n_bench_a1": "two-pair-cluster A-B-A: default", "r2n_bench_quad": "two-pair-cluster A-B-A: PLIP_GEMM_QUAD=1",
    "r2n_bench_a2": "two-pair-cluster A-B-A: default again", "r2o_bench": "with extra.last_layer_pruning_opt_in",
    "r2q_bench": "software-pipelined 16-bit epilogue", "r2r_bench_a1": "pipelined fp32 epilogues (A)",
    "r2r_bench_b": "PLIP_GEMM_F32_SERIAL=1 (B)", "r2r_bench_a2": "pipelined fp32 epilogues (A again)",
    "r2s_bench": "END OF ROUND driver-style run of HEAD (slower box: 1447 MHz under the cap)",
    "r2s_bench_bn256": "PLIP_GEMM_BN=256 forced on every GEMM (out_proj A/B)",
    "r2f_bench": "FINAL driver-style run", "r2f_bench_reference": "FINAL reference arm", "r2f_bench_cfg3": "FINAL cfg3 line",
}
KEEP_FULL = {"r2f_bench", "r2f_bench_reference", "r2f_bench_cfg3", "r2e_bench", "r2s_bench"}


def trim(b):
    out = {k: b.get(k) for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "dtype", "clocks", "gpu_launches")
           if k in b}
    out["config"] = b.get("config", {}).get("name")
    if "e2e" in b:
        out["e2e"] = {k: v for k, v in b["e2e"].items() if k != "path"}
    ex = b.get("extra") or {}
    if "kernels_in_step" in ex:
        out["kernels_in_step_us"] = {k["kernel"]: round(k["us"], 1) for k in ex["kernels_in_step"] if k["share_of_step"] > 0.02}
    for k in ("vision_tower_1024_bf16", "text_tower_1024x77", "small_batch_latency_forward", "similarity_block", "fused_topk50_merge"):
        if k in ex:
            out[k] = ex[k]
    if b.get("cpu_baseline"):
        out["cpu_baseline"] = {k: b["cpu_baseline"][k] for k in ("value", "cores", "kind")}
    return out


def main():
    res = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2*.json")) + glob.glob(os.path.join(ROOT, "gpurun_out", "r2*.jsonl"))):
        key = os.path.basename(f).rsplit(".", 1)[0]
        lines = []
        for l in open(f):
            if l.strip().startswith("{"):
                try:
                    lines.append(json.loads(l))
                except json.JSONDecodeError:   # a pretty-printed (multi-line) JSON file, not a bench line
                    break
        if not lines or "value" not in lines[0]:
            continue
        for i, b in enumerate(lines):
            k = key if len(lines) == 1 else f"{key}#{b.get('config', {}).get('name', i)}"
            res[k] = {"label": LABELS.get(key, ""), "line": b if key in KEEP_FULL else trim(b)}
    json.dump(res, open(os.path.join(ROOT, "profiles", "r2_bench_lines.json"), "w"), indent=1)
    for k, v in res.items():
        b = v["line"]
        print(f"{k:40s} {b.get('value', 0):>12.0f} {b.get('unit', ''):9s} {b.get('ms_per_step', 0):9.2f} ms  {v['label']}")


if __name__ == "__main__":
    main()
