"""N = 2 GPUs under NCCL: ``ShardedCLIP`` with the real CUDA engine on every rank must reproduce the single-GPU
result (BASELINE configs[3] / configs[4] flows, and the bench step) — VERDICT r1 missing #1.

Shard sizes are even and start on even image indices, so every image keeps its position parity inside a packed
attention tile (2 images per 128-row tile): the embeddings are then BIT-identical to the single-GPU run, and so is
everything derived from them.  Skipped on boxes with fewer than 2 GPUs (run with ``gpurun --gpus 2``)."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_IMG, N_CLS, N_Q, TOPK = 300, 64, 40, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    from oracle import synth
    tiles = torch.from_numpy(synth.tiles_u8(N_IMG, seed=41))
    cls_ids = synth.token_ids(N_CLS, seed=42)[0]
    q_ids = synth.token_ids(N_Q, seed=43)[0]
    return tiles, cls_ids, q_ids


def _worker(rank, ws, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    try:
        from oracle import weights
        from plip_b200 import distributed as D
        from plip_b200.engine import Engine
        torch.set_grad_enabled(False)
        eng = Engine(weights.make_state_dict(0, "rich"), device=dev, max_micro_batch=64)
        sh = D.ShardedCLIP.from_engine(eng)
        tiles, cls_ids, q_ids = _inputs()
        sl = sh.local_slice(N_IMG)
        pred, logits, all_img = sh.zero_shot(tiles[sl].to(dev), cls_ids.to(dev), N_IMG)
        qs = sh.local_slice(N_Q)
        chunks = (tiles[sl][i:i + 64].to(dev) for i in range(0, sl.stop - sl.start, 64))     # streamed gallery
        block, gal, q_all = sh.retrieval(chunks, q_ids[qs].to(dev), N_Q)
        top_i, top_v = sh.retrieval_topk(gal, q_all, TOPK, N_IMG)
        lpi = sh.clip_forward(tiles[sl].to(dev), q_ids[qs].to(dev))
        torch.cuda.synchronize()
        torch.save({k: v.cpu() for k, v in dict(pred=pred, logits=logits, all_img=all_img, block=block, q_all=q_all,
                                                 top_i=top_i, top_v=top_v, lpi=lpi).items()},
                   os.path.join(outdir, f"r{rank}.pt"))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_clip_two_gpus_equals_single_gpu(state_dict):
    ws = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(ws, _free_port(), d), nprocs=ws, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(ws)]
    from plip_b200.engine import Engine
    eng = Engine(state_dict, max_micro_batch=64)
    tiles, cls_ids, q_ids = _inputs()
    img = eng.encode_images(tiles.cuda(), normalize=True)
    cls = eng.encode_text(cls_ids.cuda(), normalize=True)
    q = eng.encode_text(q_ids.cuda(), normalize=True)
    s = eng.logit_scale_exp
    ref_logits = eng.similarity(img, cls, normalize_image=False, normalize_text=False).cpu()
    ref_block = eng.similarity(img, q, normalize_image=False, normalize_text=False).cpu()
    ref_ti, ref_tv = eng.similarity_topk(q, img, TOPK, scale=1.0, normalize_query=False, normalize_space=False)
    assert torch.equal(torch.cat([r["logits"] for r in res]), ref_logits)               # bit for bit
    assert torch.equal(torch.cat([r["pred"] for r in res]), ref_logits.argmax(-1))
    assert torch.equal(torch.cat([r["block"] for r in res]), ref_block)
    assert torch.equal(torch.cat([r["lpi"] for r in res]), ref_block)
    for r in res:
        assert torch.equal(r["all_img"], img.cpu()) and torch.equal(r["q_all"], q.cpu())
        assert torch.equal(r["top_i"], ref_ti.cpu().to(torch.int64)) and torch.equal(r["top_v"], ref_tv.cpu())
    assert abs(s - 14.285) < 0.1
    eng.close()
