"""Do the two towers overlap usefully when run on two streams with separate workspaces?"""
import torch
from oracle import weights, synth
from plip_b200.engine import Engine
sd = weights.make_state_dict(0)
ev = Engine(sd, max_micro_batch=1024)
et = Engine(sd, max_micro_batch=1024)
px = torch.randn(1024, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
ids = synth.token_ids(1024, full_length=True)[0].cuda()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def seq():
    a = ev.encode_images(px, normalize=True); b = ev.encode_text(ids, normalize=True); return ev.similarity(a, b, normalize_image=False, normalize_text=False)
def par():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): a = ev.encode_images(px, normalize=True)
    with torch.cuda.stream(s2): b = et.encode_text(ids, normalize=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
    return ev.similarity(a, b, normalize_image=False, normalize_text=False)
for name, fn in (("sequential", seq), ("two_streams", par), ("sequential", seq), ("two_streams", par)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "ms/step", e0.elapsed_time(e1) / 10)
a = seq(); b = par(); print("max diff", (a - b).abs().max().item())
