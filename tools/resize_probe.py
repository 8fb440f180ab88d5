"""Timing of the device resize + crop kernel (plip_resize_crop_u8) against PIL on one host thread.

Usage: python tools/resize_probe.py [out.json].  Roofline: HBM — algorithmic bytes = packed source bytes + 150,528
tile bytes per image; peak from MEASURED_PEAKS.json when present."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import PIL.Image
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from plip_b200 import Engine, preprocess as P  # noqa: E402
from plip_b200.synthetic import make_state_dict  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    peaks = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    hbm = None
    if peaks.exists():
        pk = json.loads(peaks.read_text())
        hbm = pk.get("hbm_gbs")
    eng = Engine(make_state_dict(seed=0), device="cuda:0", max_micro_batch=64)
    rng = np.random.default_rng(0)
    rows = []
    for (h, w, n) in [(256, 256, 1024), (512, 512, 512), (1000, 1000, 256), (2000, 1500, 64), (96, 96, 1024)]:
        base = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(8)]
        arrs = [base[i % 8] for i in range(n)]
        buf, d = P.pack_rgb(arrs)
        src = buf.cuda()
        tiles = eng.resize_crop(src, d)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            eng.resize_crop(src, d, out=tiles)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        t0 = time.perf_counter()
        for a in base:
            np.asarray(P.resize_center_crop(PIL.Image.fromarray(a)))
        pil_ms = (time.perf_counter() - t0) / len(base) * 1e3
        nbytes = sum(a.nbytes for a in arrs) + n * 224 * 224 * 3
        row = {"h": h, "w": w, "n": n, "gpu_ms": round(ms, 4), "gpu_us_per_image": round(ms * 1e3 / n, 3),
               "pil_ms_per_image_1thread": round(pil_ms, 3), "algorithmic_GB": round(nbytes / 1e9, 4),
               "achieved_GBps": round(nbytes / 1e9 / (ms / 1e3), 1),
               "hbm_frac": round(nbytes / 1e9 / (ms / 1e3) / hbm, 4) if hbm else None,
               "bit_exact_vs_pil": bool(np.array_equal(tiles[:8].cpu().numpy(),
                                                       np.stack([np.asarray(P.resize_center_crop(PIL.Image.fromarray(a))) for a in base])))}
        print(json.dumps(row), flush=True)
        rows.append(row)
    if out_path:
        Path(out_path).write_text(json.dumps({"hbm_peak_GBps": hbm, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
