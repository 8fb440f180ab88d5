"""Where does a layer GEMM's time go?  For the shapes of the text / vision towers, times the shipped epilogue against
the diagnostic variants of the same kernel: PLIP_GEMM_DBG=1 (no global stores), 2 (LayerNorm-fold math only: no GELU, no
packing, no staging, no stores) and EPI_NULL (accumulators read from TMEM and dropped).  One subprocess per (shape, variant): the switch is
read once per process.  500 back-to-back launches per number (long enough for the power cap to settle).
Usage: python tools/epilogue_probe.py [out.json]"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {  # name: (M, N, K, epi)
    "text_qkv": (78848, 1536, 512, 5),
    "text_fc1": (78848, 2048, 512, 6),
    "vision_qkv": (51200, 2304, 768, 5),
    "vision_fc1": (51200, 3072, 768, 6),
}
VARIANTS = [("full", 0, None), ("no_stores", 1, None), ("fold_math_only", 2, None), ("epi_null", 0, 7)]


def run(name, epi_override):
    import torch
    from plip_b200._lib import lib, check
    M, N, K, epi = SHAPES[name]
    if epi_override is not None:
        epi = epi_override
    L = lib(strict=False)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    colsum = torch.randn(N, generator=g).to(dev)
    stats = torch.zeros(M, 8, 2, device=dev)   # kStatSlots = 8
    stats[:, 0, 0] = 0.1 * K
    stats[:, 0, 1] = 1.0 * K
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if epi != 7 else torch.float32)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        check(L.plip_dbg_gemm(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), out.data_ptr(), N, None, epi, 2, 256,
                              colsum.data_ptr(), stats.data_ptr(), 1, None, None, st), "gemm")
    for _ in range(20):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 500
    print(json.dumps({"us": us, "tflops": 2.0 * M * N * K / us / 1e6}))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] in SHAPES:
        run(sys.argv[1], None if sys.argv[2] == "-" else int(sys.argv[2]))
        sys.exit(0)
    res = {}
    only = os.environ.get("PROBE_VARIANTS")   # e.g. "full,epi_null"
    for name in SHAPES:
        for vname, dbg, epi in VARIANTS:
            if only and vname not in only.split(","):
                continue
            env = dict(os.environ, PLIP_GEMM_DBG=str(dbg))
            try:
                r = subprocess.run([sys.executable, __file__, name, "-" if epi is None else str(epi)], env=env,
                                   capture_output=True, text=True, timeout=180)
                line = (r.stdout.strip().splitlines() or ["{}"])[-1]
                res[f"{name}/{vname}"] = json.loads(line) if r.returncode == 0 else {"rc": r.returncode, "err": r.stderr[-300:]}
            except subprocess.TimeoutExpired:
                res[f"{name}/{vname}"] = {"timeout": True}
            print(name, vname, res[f"{name}/{vname}"], flush=True)
    if len(sys.argv) > 1 and sys.argv[1] not in SHAPES:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
