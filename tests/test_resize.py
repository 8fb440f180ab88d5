"""Image resize + crop: oracle pinned on PIL, host planning, and the CUDA kernel against PIL.

The reference resizes with PIL on the host (``plip.py:35`` through CLIPProcessor; ``reproducibility/embedders/
transform.py:45-52`` through torchvision on PIL images).  Integer work: every comparison here is bit-exact."""
import ctypes as C
import hashlib
import os

import numpy as np
import PIL.Image
import pytest

from oracle import resize_oracle as R
from plip_b200 import preprocess as P

# (h, w) source sizes: downscale, upscale, identity, one-axis identity, extreme aspect, odd sizes
SIZES = [(256, 256), (300, 500), (1000, 777), (96, 96), (224, 224), (225, 224), (500, 224), (333, 1999),
         (1536, 2048), (100, 150), (231, 229)]


def _img(rng, h, w, kind="noise"):
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = ((xx * 7 + yy * 3) % 256).astype(np.uint8)           # sharp diagonal stripes: exercises the clamp
    return np.stack([base, 255 - base, ((xx // 8 + yy // 8) % 2 * 255).astype(np.uint8)], axis=-1)


def _golden():
    import importlib.util
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize_golden.npz"))
    spec = importlib.util.spec_from_file_location("make_resize_golden",
                                                  os.path.join(os.path.dirname(__file__), "golden", "make_resize_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = [(int(h), int(w), int(seed), str(k), str(c))
             for (h, w, seed), k, c in zip(g["cases"], g["kinds"], g["crops"])]
    return mod.make_image, cases, [str(x) for x in g["sha256"]], g["patches"]


def _sha(tile):
    return hashlib.sha256(np.ascontiguousarray(tile).tobytes()).hexdigest()


def test_oracle_and_host_path_match_committed_golden_tiles():
    """Frozen outputs of the reference's own pipelines (torchvision transform / transformers PIL processor,
    tests/golden/make_resize_golden.py): the oracle and the PIL host route must reproduce them byte for byte."""
    make_image, cases, shas, patches = _golden()
    assert len(cases) >= 8
    for (h, w, seed, kind, crop), sha, patch in zip(cases, shas, patches):
        a = make_image(h, w, seed, kind)
        nw, nh, left, top = P.resize_plan(w, h, crop=crop)
        tile = R.resize_crop_u8(a, nw, nh, left, top)
        assert np.array_equal(tile[:24, :24], patch) and _sha(tile) == sha, (h, w, crop)
        host = P.to_uint8_tiles([a], crop=crop)[0]
        assert _sha(host) == sha, (h, w, crop)


def _pil_tile(a, crop):
    return np.asarray(P.resize_center_crop(PIL.Image.fromarray(a), P.SIZE, crop))


@pytest.mark.parametrize("h,w", SIZES)
def test_oracle_matches_pil_resize(h, w):
    rng = np.random.default_rng(h * 10007 + w)
    for kind in ("noise", "stripes"):
        a = _img(rng, h, w, kind)
        nw, nh, left, top = P.resize_plan(w, h)
        ref = np.asarray(PIL.Image.fromarray(a).resize((nw, nh), resample=PIL.Image.BICUBIC))
        got = R.resize_bicubic_u8(a, nw, nh)
        assert np.array_equal(ref, got)
        assert np.array_equal(R.resize_crop_u8(a, nw, nh, left, top), _pil_tile(a, "floor"))


def test_oracle_matches_pil_on_random_sizes():
    """Seeded sweep over arbitrary (also non-aspect-preserving) size pairs, up- and down-scaling."""
    rng = np.random.default_rng(2024)
    for _ in range(25):
        h, w = (int(v) for v in rng.integers(1, 400, 2))
        nh, nw = (int(v) for v in rng.integers(1, 400, 2))
        a = _img(rng, h, w, "noise" if rng.random() < 0.7 else "stripes")
        ref = np.asarray(PIL.Image.fromarray(a).resize((nw, nh), resample=PIL.Image.BICUBIC))
        assert np.array_equal(R.resize_bicubic_u8(a, nw, nh), ref), (h, w, nh, nw)


def test_resize_plan_conventions():
    # floor = CLIPImageProcessor.center_crop, round = torchvision CenterCrop (banker's rounding of x.5)
    assert P.resize_plan(224, 224) == (224, 224, 0, 0)
    assert P.resize_plan(500, 300) == (373, 224, 74, 0)            # excess 149: floor -> 74
    assert P.resize_plan(500, 300, crop="round") == (373, 224, 74, 0)   # 74.5 -> 74 (half to even)
    assert P.resize_plan(300, 502, crop="floor") == (224, 374, 0, 75)
    assert P.resize_plan(300, 506, crop="floor")[3] == 76 and P.resize_plan(300, 506, crop="round")[3] == 76
    assert P.resize_plan(229, 224, crop="floor")[2] == 2 and P.resize_plan(229, 224, crop="round")[2] == 2
    assert P.resize_plan(227, 224, crop="floor")[2] == 1 and P.resize_plan(227, 224, crop="round")[2] == 2
    with pytest.raises(ValueError):
        P.resize_plan(10, 10, crop="centre")


def test_host_tiles_match_reference_processors():
    """``to_uint8_tiles`` equals torchvision's PIL transform (crop='round') and transformers' PIL image
    processor (crop='floor') on non-square, non-224 inputs."""
    rng = np.random.default_rng(5)
    imgs = [PIL.Image.fromarray(_img(rng, h, w)) for h, w in [(300, 401), (517, 233), (224, 224), (100, 150), (227, 224)]]
    tv = pytest.importorskip("torchvision.transforms")
    t = tv.Compose([tv.Resize(224, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop(224)])
    assert np.array_equal(P.to_uint8_tiles(imgs, crop="round"), np.stack([np.asarray(t(im)) for im in imgs]))
    try:
        from transformers import CLIPImageProcessorPil as Proc     # transformers >= 5: the PIL backend
    except ImportError:
        pytest.skip("no PIL-backed CLIP image processor in this transformers")
    o = Proc()(images=imgs, return_tensors="np", do_normalize=False, do_rescale=False)["pixel_values"]
    assert np.array_equal(P.to_uint8_tiles(imgs, crop="floor"), np.transpose(o, (0, 2, 3, 1)))


def test_pack_rgb_descriptors():
    rng = np.random.default_rng(2)
    arrs = [_img(rng, 30, 50), _img(rng, 224, 224), _img(rng, 301, 17)]
    buf, d = P.pack_rgb(arrs, crop="floor")
    assert d.dtype == P.RESIZE_DESC_DTYPE and d.dtype.itemsize == 32 and buf.dtype.is_floating_point is False
    off = 0
    for a, row in zip(arrs, d):
        h, w = a.shape[:2]
        assert int(row["offset"]) == off and off % 16 == 0
        assert (int(row["width"]), int(row["height"])) == (w, h)
        assert (int(row["new_width"]), int(row["new_height"]), int(row["left"]), int(row["top"])) == P.resize_plan(w, h)
        assert np.array_equal(buf.numpy()[off:off + h * w * 3].reshape(h, w, 3), a)
        off += (h * w * 3 + 15) // 16 * 16
    assert buf.numel() == off
    with pytest.raises(ValueError):
        P.pack_rgb([np.zeros((4, 4), np.uint8)])


@pytest.mark.parametrize("in_size,out_size", [(256, 224), (96, 224), (224, 224), (1999, 1344), (3000, 224), (225, 224),
                                              (150, 336), (7000, 224)])
def test_library_filter_rows_match_oracle(in_size, out_size):
    """The filter-bank code the kernel runs (host instantiation, no GPU) against the oracle's coefficients."""
    from plip_b200._lib import lib
    L = lib()
    xmins, counts, kk = R.coefficients(in_size, out_size)
    ks = kk.shape[1]
    buf = (C.c_int32 * ks)()
    xmin, cnt = C.c_int(), C.c_int()
    step = max(1, out_size // 97)
    for xx in list(range(0, out_size, step)) + [out_size - 1]:
        rc = L.plip_dbg_resize_filter(in_size, out_size, xx, C.addressof(buf), ks, C.byref(xmin), C.byref(cnt))
        assert rc == ks
        assert (xmin.value, cnt.value) == (int(xmins[xx]), int(counts[xx]))
        assert list(buf[:cnt.value]) == kk[xx, :cnt.value].tolist()
    assert L.plip_dbg_resize_filter(in_size, out_size, 0, C.addressof(buf), ks - 1, C.byref(xmin), C.byref(cnt)) == -ks
    assert L.plip_dbg_resize_filter(in_size, out_size, out_size, C.addressof(buf), ks, C.byref(xmin), C.byref(cnt)) == -2


def test_device_resizable_predicate_matches_library_validation():
    """``preprocess.device_resizable`` (used to route oversized images to PIL) against the library's own check.
    On a CPU-only box the call stops at the first CUDA call, after descriptor validation; never run this where a
    GPU is present (the pointers are fake)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("validation-only probe: needs a box without a GPU")
    from plip_b200._lib import ResizeDesc, last_error, lib
    L = lib()
    sizes = [(256, 256), (1000, 1000), (4000, 3000), (3000, 5000), (6000, 6000), (6500, 6500), (7000, 7000), (8000, 8000),
             (12000, 9000), (8, 60000), (60000, 8), (20000, 300), (300, 20000), (65536, 224), (224, 65536), (5000, 7000)]
    seen = set()
    for w, h in sizes:
        nw, nh, left, top = P.resize_plan(w, h)
        d = ResizeDesc(0, w, h, nw, nh, left, top)
        assert L.plip_resize_crop_u8(4096, 1 << 40, C.byref(d), 1, 4096, None) != 0
        lib_ok = "shrinks too much" not in last_error() and "invalid size" not in last_error() \
            and "smaller than" not in last_error()
        assert lib_ok == P.device_resizable(w, h), (w, h, last_error())
        seen.add(lib_ok)
    assert seen == {True, False}


def test_plip_encode_images_routes_oversized_images_to_pil():
    """Flush routing of ``PLIP.encode_images`` with a stub engine: all-224 batches go straight to the host path,
    resizable batches to the device kernel, batches holding an image the kernel cannot take to PIL."""
    import torch
    from plip_b200.plip import PLIP

    class StubEngine:
        max_micro_batch, device = 4, "cpu"

        def __init__(self):
            self.calls = []

        def encode_images_host(self, tiles, normalize=False):
            self.calls.append(("host", tuple(tiles.shape)))
            return torch.zeros(tiles.shape[0], 512)

        def resize_crop(self, src, descs):
            self.calls.append(("device_resize", len(descs)))
            return torch.zeros(len(descs), 224, 224, 3, dtype=torch.uint8)

        def encode_images(self, tiles, normalize=False):
            self.calls.append(("device", tuple(tiles.shape)))
            return torch.zeros(tiles.shape[0], 512)

    class StubModel:
        engine = StubEngine()

    plip = PLIP.__new__(PLIP)
    plip.model, plip.num_workers, plip.device_resize = StubModel(), 0, True
    rng = np.random.default_rng(0)
    import plip_b200.plip as mod
    real_pack = mod.pack_rgb
    mod.pack_rgb = lambda arrs, crop="floor", pinned=False: real_pack(arrs, crop=crop, pinned=False)   # no CUDA here
    try:
        tiles224 = [_img(rng, 224, 224) for _ in range(4)]
        mixed = [_img(rng, 300, 260), _img(rng, 224, 224), _img(rng, 240, 500), _img(rng, 224, 224)]
        out = plip.encode_images(tiles224 + mixed, batch_size=2)
        assert out.shape == (8, 512)
        assert StubModel.engine.calls == [("host", (4, 224, 224, 3)), ("device_resize", 4), ("device", (4, 224, 224, 3))]
        StubModel.engine.calls.clear()
        assert not P.device_resizable(8, 9000)
        plip.encode_images([np.zeros((9000, 8, 3), np.uint8), tiles224[0]], batch_size=2)
        assert StubModel.engine.calls == [("host", (2, 224, 224, 3))]               # PIL route, then the host path
        StubModel.engine.calls.clear()
        plip.device_resize = False
        plip.encode_images(mixed, batch_size=4)
        assert StubModel.engine.calls == [("host", (4, 224, 224, 3))]
    finally:
        mod.pack_rgb = real_pack


# ---- GPU ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("crop", ["floor", "round"])
def test_device_resize_matches_pil(engine, crop):
    import torch
    rng = np.random.default_rng(11)
    shapes = SIZES + [(2240, 3000), (4000, 3000), (64, 64), (225, 1200), (5000, 700)]
    arrs = [_img(rng, h, w, "noise" if i % 3 else "stripes") for i, (h, w) in enumerate(shapes)]
    buf, d = P.pack_rgb(arrs, crop=crop)
    tiles = engine.resize_crop(buf.cuda(), d)
    torch.cuda.synchronize()
    got = tiles.cpu().numpy()
    assert got.shape == (len(arrs), 224, 224, 3) and got.dtype == np.uint8
    for i, a in enumerate(arrs):
        ref = _pil_tile(a, crop)
        assert np.array_equal(got[i], ref), f"image {i} {a.shape}: {(got[i] != ref).sum()} bytes differ"
    # the oracle agrees too (one mid-size case; the rest is covered on the CPU tier)
    nw, nh, left, top = P.resize_plan(500, 300, crop=crop)
    assert np.array_equal(got[1], R.resize_crop_u8(arrs[1], nw, nh, left, top))


@pytest.mark.gpu
def test_device_resize_matches_committed_golden_tiles(engine):
    make_image, cases, shas, patches = _golden()
    for crop in ("floor", "round"):
        sel = [i for i, c in enumerate(cases) if c[4] == crop]
        arrs = [make_image(*cases[i][:4]) for i in sel]
        buf, d = P.pack_rgb(arrs, crop=crop)
        got = engine.resize_crop(buf.cuda(), d).cpu().numpy()
        for j, i in enumerate(sel):
            assert np.array_equal(got[j][:24, :24], patches[i]) and _sha(got[j]) == shas[i], cases[i]


@pytest.mark.gpu
def test_device_resize_many_images_and_arbitrary_window(engine):
    """More images than one launch carries (512), and a crop window that is not centred."""
    import torch
    rng = np.random.default_rng(12)
    arrs = [_img(rng, int(rng.integers(200, 330)), int(rng.integers(200, 330))) for _ in range(530)]
    buf, d = P.pack_rgb(arrs)
    d = d.copy()
    d["new_width"][:3], d["new_height"][:3], d["left"][:3], d["top"][:3] = 300, 260, (0, 76, 31), (36, 0, 17)
    got = engine.resize_crop(buf.cuda(), d).cpu().numpy()
    for i, a in enumerate(arrs):
        nw, nh, left, top = (int(d[k][i]) for k in ("new_width", "new_height", "left", "top"))
        ref = np.asarray(PIL.Image.fromarray(a).resize((nw, nh), resample=PIL.Image.BICUBIC))[top:top + 224, left:left + 224]
        assert np.array_equal(got[i], ref), f"image {i}"


@pytest.mark.gpu
def test_device_resize_source_ending_mid_word(engine):
    """The kernel fetches source bytes as aligned 32-bit words; a buffer whose size is not a multiple of 4 must
    still be read correctly (and never past its end) for the last pixels of the last image."""
    import torch
    rng = np.random.default_rng(14)
    for h, w in [(225, 225), (227, 301), (451, 233)]:
        a = _img(rng, h, w)
        assert (h * w * 3) % 4 != 0
        src = torch.from_numpy(a.reshape(-1).copy()).cuda()            # exactly h*w*3 bytes, no padding
        d = np.zeros(1, dtype=P.RESIZE_DESC_DTYPE)
        d[0] = (0, w, h) + P.resize_plan(w, h)
        # crop windows that touch the right / bottom edge of the resized image
        nw, nh = int(d["new_width"][0]), int(d["new_height"][0])
        d["left"][0], d["top"][0] = nw - 224, nh - 224
        got = engine.resize_crop(src, d).cpu().numpy()[0]
        ref = np.asarray(PIL.Image.fromarray(a).resize((nw, nh), resample=PIL.Image.BICUBIC))[nh - 224:, nw - 224:]
        assert np.array_equal(got, ref)


@pytest.mark.gpu
def test_device_resize_rejects_bad_descriptors(engine):
    import torch
    src = torch.zeros(300 * 300 * 3, dtype=torch.uint8, device="cuda")
    d = np.zeros(1, dtype=P.RESIZE_DESC_DTYPE)
    d[0] = (0, 300, 300, 224, 224, 0, 0)
    engine.resize_crop(src, d)                                     # fine
    for bad, msg in [((16, 300, 300, 224, 224, 0, 0), "exceeds"), ((0, 300, 300, 200, 224, 0, 0), "smaller"),
                     ((0, 300, 300, 224, 224, 1, 0), "crop origin"), ((0, 0, 300, 224, 224, 0, 0), "invalid size")]:
        d[0] = bad
        with pytest.raises(RuntimeError, match=msg):
            engine.resize_crop(src, d)
    big = torch.zeros(60000 * 8 * 3, dtype=torch.uint8, device="cuda")
    d[0] = (0, 8, 60000, 224, 224, 0, 0)                           # 268x vertical shrink: filter bank > shared memory
    with pytest.raises(RuntimeError, match="shrinks too much"):
        engine.resize_crop(big, d)


@pytest.mark.gpu
def test_plip_encode_images_device_resize_equals_pil_route(state_dict):
    """``PLIP.encode_images`` on non-224 images: device resize and host PIL resize feed identical tiles, so the
    embeddings are bit-identical; mixed batches with 224x224 tiles keep their order."""
    from plip_b200 import PLIP
    rng = np.random.default_rng(13)
    shapes = [(224, 224), (300, 400), (512, 512), (224, 224), (180, 260), (700, 300), (224, 224)]
    imgs = [PIL.Image.fromarray(_img(rng, h, w)) for h, w in shapes]
    plip = PLIP.from_state_dict(state_dict, max_micro_batch=64)
    a = plip.encode_images(imgs, batch_size=3)
    plip.device_resize = False
    b = plip.encode_images(imgs, batch_size=3)
    assert a.shape == (7, 512) and a.dtype == np.float32

    def one_minus_cos(x, y):
        return float((1.0 - (x * y).sum(1) / np.linalg.norm(x, axis=1) / np.linalg.norm(y, axis=1)).max())

    assert one_minus_cos(a, b) <= 1e-6          # identical tiles -> same embeddings (bitwise unless batching differs)
    only224 = plip.encode_images([imgs[0], imgs[3], imgs[6]], batch_size=2)
    assert one_minus_cos(only224, a[[0, 3, 6]]) <= 1e-5
