"""N>1 host path on CPU: world_size-2 gloo processes exercise the sharding / all-gather logic with a
deterministic stand-in encoder (the CUDA engine is rank-local and needs no collective)."""
import os
import socket
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plip_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _encoders():
    g = torch.Generator().manual_seed(7)
    wi = torch.randn(12, 512, generator=g)
    wt = torch.randn(5, 512, generator=g)

    def norm(x):
        return x / x.norm(dim=-1, keepdim=True)

    enc_i = lambda rows: norm(rows.float() @ wi)          # noqa: E731
    enc_t = lambda rows: norm(rows.float() @ wt)          # noqa: E731
    sim = lambda a, b, s: s * a @ b.t()                   # noqa: E731
    return enc_i, enc_t, sim


def _topk(q, space, k):
    """Stand-in for Engine.similarity_topk: descending scores, ties by lower index."""
    sc = q @ space.t()
    val, idx = torch.sort(sc, dim=1, descending=True, stable=True)
    return idx[:, :k].to(torch.int32), val[:, :k]


def _data():
    g = torch.Generator().manual_seed(11)
    return torch.randn(23, 12, generator=g), torch.randn(9, 5, generator=g), torch.randn(4, 5, generator=g)


def _worker(rank, ws, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        imgs, queries, classes = _data()
        enc_i, enc_t, sim = _encoders()
        sh = D.ShardedCLIP(enc_i, enc_t, sim, 14.3, topk=_topk)
        assert (sh.rank, sh.world_size) == (rank, ws)
        # uneven all-gather restores order
        sl = sh.local_slice(imgs.shape[0])
        gathered = D.all_gather_rows(imgs[sl].contiguous())
        assert torch.equal(gathered, imgs)
        gathered2 = D.all_gather_rows(imgs[sl].contiguous(), D.shard_counts(imgs.shape[0], ws))
        assert torch.equal(gathered2, imgs)
        # async variant (equal blocks): result identical, usable after work.wait()
        eq = imgs[:20].view(2, 10, 12)[rank].contiguous()
        got, work = D.all_gather_rows_async(eq)
        work.wait()
        assert torch.equal(got, imgs[:20])
        # cfg4 flow
        pred, logits, all_img = sh.zero_shot(imgs[sl], classes, imgs.shape[0])
        # cfg5 flow
        qs = sh.local_slice(queries.shape[0])
        block, gal, q_all = sh.retrieval(imgs[sl], queries[qs], queries.shape[0])
        # gallery streamed in chunks == gallery as one tensor
        block_c, _, _ = sh.retrieval(iter([imgs[sl][:5], imgs[sl][5:]]), queries[qs], queries.shape[0])
        assert torch.equal(block_c, block)
        # retrieval head over the sharded gallery (retrieval.py:13-16): global top-k identical on every rank
        top_i, top_v = sh.retrieval_topk(gal, q_all, 6, imgs.shape[0])
        # bench step: local images x captions of all ranks (equal caption blocks -> async gather path)
        cap_local = queries[:8].view(2, 4, 5)[rank].contiguous()
        lpi = sh.clip_forward(imgs[sl], cap_local)
        torch.save({"pred": pred, "logits": logits, "all_img": all_img, "block": block, "q_all": q_all, "lo": sl.start,
                    "top_i": top_i, "top_v": top_v, "lpi": lpi}, os.path.join(outdir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    ws = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(ws, _free_port(), d), nprocs=ws, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(ws)]
    imgs, queries, classes = _data()
    enc_i, enc_t, sim = _encoders()
    ref_img, ref_q, ref_c = enc_i(imgs), enc_t(queries), enc_t(classes)
    ref_logits = sim(ref_img, ref_c, 14.3)
    ref_block = sim(ref_img, ref_q, 14.3)
    assert torch.allclose(torch.cat([r["logits"] for r in res]), ref_logits, atol=1e-5)
    assert torch.equal(torch.cat([r["pred"] for r in res]), ref_logits.argmax(-1))
    for r in res:
        assert torch.allclose(r["all_img"], ref_img, atol=1e-6)      # gathered table identical on every rank
        assert torch.allclose(r["q_all"], ref_q, atol=1e-6)
    assert torch.allclose(torch.cat([r["block"] for r in res]), ref_block, atol=1e-5)
    assert [r["lo"] for r in res] == [0, 12]
    full = ref_q @ ref_img.t()                                           # [n_queries, n_gallery]
    ref_v, ref_i = torch.sort(full, dim=1, descending=True, stable=True)
    for r in res:
        assert torch.equal(r["top_i"], ref_i[:, :6]) and torch.allclose(r["top_v"], ref_v[:, :6], atol=1e-6)
    assert torch.allclose(torch.cat([r["lpi"] for r in res]), sim(ref_img, enc_t(queries[:8]), 14.3), atol=1e-5)
