export PYTHONPATH=.
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 200 python - <<'PY'
import torch, time
from oracle import weights, synth
from plip_b200.engine import Engine
eng = Engine(weights.make_state_dict(0), max_micro_batch=1024)
ids, mask = synth.token_ids(4096, seed=3, min_len=8)
ids[:, 16:] = 49407; ids[:, 15] = 49407          # every caption ends within 16 tokens (typical prompt length)
idp = ids.pin_memory()
for name, fn in (("full 77", lambda: eng.encode_text(ids.cuda())), ("prefix 16", lambda: eng.encode_text(ids.cuda(), prefix_len=16)), ("host path (auto prefix)", lambda: eng.encode_text_host(idp))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): out = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(name, f"{dt*1e3:.2f} ms per 4096 captions -> {4096/dt:,.0f} captions/s")
PY
