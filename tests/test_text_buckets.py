"""Length-bucket plan of ``plip_encode_text_host`` (only the positions up to a caption's first EOS
matter, TF:modeling_clip.py:571-584).  The planner is host code: exercised here without a GPU."""
import ctypes as C

import numpy as np
import pytest

from plip_b200._lib import lib

CAP = 8


def plan(lens, seq_len=77, cap=CAP):
    L = lib()
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n = len(lens)
    perm = np.empty(n, np.int32)
    start = np.empty(cap + 1, np.int32)
    prefix = np.empty(cap, np.int32)
    nb = L.plip_dbg_text_bucket_plan(lens.ctypes.data, n, seq_len, perm.ctypes.data, start.ctypes.data,
                                     prefix.ctypes.data, cap)
    assert 1 <= nb <= cap
    return nb, perm, start[:nb + 1].copy(), prefix[:nb].copy()


def cost(lens, start, prefix):
    return sum(max((start[k + 1] - start[k]) * prefix[k], 8192) + 2048 for k in range(len(prefix)))


def check_valid(lens, nb, perm, start, prefix):
    n = len(lens)
    assert sorted(perm.tolist()) == list(range(n))                      # a permutation
    sl = np.asarray(lens)[perm]
    assert np.all(np.diff(sl) >= 0)                                       # sorted by length ...
    same = np.flatnonzero(np.diff(sl) == 0)
    assert np.all(perm[same] < perm[same + 1])                            # ... stably
    assert start[0] == 0 and start[-1] == n and np.all(np.diff(start) > 0)
    for k in range(nb):
        seg = sl[start[k]:start[k + 1]]
        assert seg.max() == prefix[k]                                     # prefix = longest caption of the bucket
    assert np.all(np.diff(prefix) > 0)


def test_small_or_uniform_batches_stay_in_one_bucket():
    for lens in ([12] * 8, list(range(5, 17)), [77] * 4096, [9] * 100000, [5, 77]):
        nb, perm, start, prefix = plan(lens)
        assert nb == 1 and prefix[0] == max(lens) and perm.tolist() == sorted(range(len(lens)), key=lambda i: (lens[i], i))


def test_mixed_lengths_are_bucketed_and_cheaper():
    rng = np.random.default_rng(0)
    lens = rng.integers(8, 78, 4096)
    nb, perm, start, prefix = plan(lens)
    check_valid(lens, nb, perm, start, prefix)
    assert nb >= 3
    single = cost(lens, np.array([0, len(lens)]), np.array([lens.max()]))
    assert cost(lens, start, prefix) < 0.75 * single
    # typical prompts plus a few long captions: the long tail gets its own bucket
    lens2 = np.concatenate([rng.integers(10, 16, 8000), rng.integers(60, 78, 300)])
    nb2, perm2, start2, prefix2 = plan(lens2)
    check_valid(lens2, nb2, perm2, start2, prefix2)
    assert nb2 >= 2 and prefix2[0] <= 15 and cost(lens2, start2, prefix2) < 0.4 * cost(lens2, np.array([0, len(lens2)]), np.array([77]))


def test_plan_is_optimal_for_its_cost_model():
    """Brute force over all contiguous partitions of the distinct lengths (few distinct values)."""
    import itertools
    rng = np.random.default_rng(3)
    for trial in range(20):
        vals = np.sort(rng.choice(np.arange(1, 78), size=rng.integers(1, 7), replace=False))
        counts = rng.integers(1, 6000, size=len(vals))
        lens = np.repeat(vals, counts)
        rng.shuffle(lens)
        nb, perm, start, prefix = plan(lens)
        check_valid(lens, nb, perm, start, prefix)
        best = None
        m = len(vals)
        for r in range(m):
            for cuts in itertools.combinations(range(1, m), r):
                b = [0, *cuts, m]
                c = sum(max(int(counts[b[i]:b[i + 1]].sum()) * int(vals[b[i + 1] - 1]), 8192) + 2048 for i in range(len(b) - 1))
                best = c if best is None else min(best, c)
        assert cost(lens, start, prefix) == best, (vals, counts)


def test_clamping_shorter_seq_len_and_bad_arguments():
    nb, perm, start, prefix = plan([3, 40, 99, 0], seq_len=40)              # lengths clamp to [1, seq_len]
    assert nb == 1 and prefix[0] == 40
    L = lib()
    buf = (C.c_int32 * 16)()
    assert L.plip_dbg_text_bucket_plan(None, 4, 77, None, C.addressof(buf), C.addressof(buf), 8) == -2
    assert L.plip_dbg_text_bucket_plan(C.addressof(buf), 4, 78, None, C.addressof(buf), C.addressof(buf), 8) == -2


@pytest.mark.gpu
def test_bucketed_host_path_matches_single_pass(engine):
    """A large mixed-length batch through the host path (several buckets, results un-permuted) against the
    device path that processes every caption at full length."""
    import torch
    from plip_b200 import synthetic as synth
    ids, mask = synth.token_ids(3000, seed=77, min_len=6)
    lens = mask.sum(1).numpy()
    assert plan(lens)[0] >= 2
    host = engine.encode_text_host(ids, mask)
    dev = engine.encode_text(ids.cuda(), mask.cuda()).cpu()
    cos = torch.nn.functional.cosine_similarity(host, dev, dim=1)
    assert (1 - cos).max().item() < 1e-5
    host32 = engine.encode_text_host(ids.to(torch.int32), None, normalize=True)     # int32 ids, no mask, normalised
    assert torch.allclose(host32.norm(dim=1), torch.ones(3000), atol=1e-5)
    assert (1 - torch.nn.functional.cosine_similarity(host32, dev, dim=1)).max().item() < 1e-5
