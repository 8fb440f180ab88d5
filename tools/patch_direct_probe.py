"""Vision tower at batch 1024 (bf16 NCHW pixels) with the in-step kernel profile: run once with PLIP_PATCH_DIRECT unset and
once with PLIP_PATCH_DIRECT=1 to compare im2col + patch GEMM against the direct (4-D tensor map) patch GEMM."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plip_b200 import synthetic  # noqa: E402
from plip_b200.engine import Engine  # noqa: E402

eng = Engine(synthetic.make_state_dict(), max_micro_batch=1024)
px = [synthetic.pixel_values(1024, seed=s).to(torch.bfloat16).cuda() for s in (1, 2)]
for i in range(3):
    eng.encode_images(px[i & 1])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20):
    eng.encode_images(px[i & 1])
b.record()
torch.cuda.synchronize()
res = {"direct": os.environ.get("PLIP_PATCH_DIRECT", "0"), "vision_tower_ms": a.elapsed_time(b) / 20}
eng.profile(True)
for i in range(3):
    eng.encode_images(px[i & 1])
torch.cuda.synchronize()
rows = eng.profile_read()
eng.profile(False)
res["kernels_us"] = {r["name"]: round(1e3 * r["total_ms"] / max(r["launches"], 1), 1) for r in rows
                     if "patch" in r["name"] or "im2col" in r["name"]}
print(json.dumps(res))
