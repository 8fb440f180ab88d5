"""Python handle on the CUDA engine: device memory and streams come from torch, compute from the
C ABI (``libplip_b200.so``).  There is no fallback path: a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Tuple, Union

import numpy as np
import torch

from ._lib import check, lib
from .weights import operand_format, pack_state_dict

PIX_F32_NCHW, PIX_BF16_NCHW, PIX_U8_NHWC = 0, 1, 2
IDS_I32, IDS_I64 = 0, 1
EMBED_DIM = 512
IMAGE_SIZE = 224
MAX_TEXT_LEN = 77


def _pixel_format(t: Union[torch.Tensor, np.ndarray]) -> int:
    """Validate an image batch and return its C-ABI pixel format (error text mirrors TF:204-207)."""
    shape, dtype = tuple(t.shape), t.dtype
    if dtype in (torch.uint8, np.dtype("uint8")):
        if len(shape) != 4 or shape[3] != 3:
            raise ValueError(f"uint8 images must be [n,224,224,3] (NHWC), got {shape}")
        if shape[1] != IMAGE_SIZE or shape[2] != IMAGE_SIZE:
            raise ValueError(f"Input image size ({shape[1]}*{shape[2]}) doesn't match model (224*224).")
        return PIX_U8_NHWC
    if len(shape) != 4 or shape[1] != 3:
        raise ValueError(f"pixel_values must be [n,3,224,224], got {shape}")
    if shape[2] != IMAGE_SIZE or shape[3] != IMAGE_SIZE:
        raise ValueError(f"Input image size ({shape[2]}*{shape[3]}) doesn't match model (224*224).")
    if dtype in (torch.float32, np.dtype("float32")):
        return PIX_F32_NCHW
    if dtype == torch.bfloat16:
        return PIX_BF16_NCHW
    raise TypeError(f"unsupported pixel dtype {dtype} (float32, bfloat16 or uint8)")


def _ids_dtype(dtype) -> int:
    if dtype in (torch.int64, np.dtype("int64")):
        return IDS_I64
    if dtype in (torch.int32, np.dtype("int32")):
        return IDS_I32
    raise TypeError(f"input_ids must be int32 or int64, got {dtype}")


def _check_ids(input_ids, attention_mask) -> Tuple[int, int]:
    if len(input_ids.shape) != 2:
        raise ValueError(f"input_ids must be [n, seq_len], got {tuple(input_ids.shape)}")
    n, s = int(input_ids.shape[0]), int(input_ids.shape[1])
    if s > MAX_TEXT_LEN:
        raise ValueError(
            "Sequence length must be less than max_position_embeddings (got `sequence length`: "
            f"{s} and max_position_embeddings: {MAX_TEXT_LEN}")  # TF:243-247
    if attention_mask is not None and tuple(attention_mask.shape) != (n, s):
        raise ValueError(f"attention_mask shape {tuple(attention_mask.shape)} != input_ids shape {(n, s)}")
    return n, s


@torch.no_grad()
def similarity_topk(query: torch.Tensor, space: torch.Tensor, k: int, scale: float = 1.0, normalize_query: bool = True,
                    normalize_space: bool = False, device: Union[int, str, torch.device, None] = None):
    """``plip_similarity_topk`` needs no engine handle (no weights involved): fused scores + top-k over ``space`` rows on
    ``device`` (default: the current CUDA device).  Returns ``(idx int32 [n,k], val f32 [n,k])``, best first."""
    if not torch.cuda.is_available():
        raise RuntimeError("plip_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = lib()
    q = query.to(torch.float32).to(dev, non_blocking=True).contiguous()
    s = space.to(torch.float32).to(dev, non_blocking=True).contiguous()
    n = int(q.shape[0])
    idx = torch.empty(n, k, device=dev, dtype=torch.int32)
    val = torch.empty(n, k, device=dev, dtype=torch.float32)
    if n == 0:
        return idx, val
    with torch.cuda.device(dev):
        check(L.plip_similarity_topk(q.data_ptr(), n, s.data_ptr(), int(s.shape[0]), C.c_float(scale), int(normalize_query),
                                     int(normalize_space), int(k), idx.data_ptr(), val.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), "plip_similarity_topk")
    return idx, val


class Engine:
    """One engine per CUDA device: packed bf16/fp32 weights + workspace for ``max_micro_batch``."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], device: Union[int, str, torch.device, None] = None,
                 max_micro_batch: int = 1024, operand_dtype="bf16"):
        """``operand_dtype``: 16-bit format of the GEMM / attention operands — ``"bf16"`` (default, BASELINE's dtype) or
        ``"fp16"`` (same speed, 3 more significand bits: 6-8x smaller end-to-end logits error, 65504 range; what the
        reference's OpenAI-clip flavour runs on a GPU).  Accumulation / residual stream / softmax are fp32 in both."""
        if not torch.cuda.is_available():
            raise RuntimeError("plip_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._L = lib()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"plip_b200 runs on CUDA devices only, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        fmt, _ = operand_format(operand_dtype)
        self.operand_dtype = "fp16" if fmt == 1 else "bf16"
        blob, scale = pack_state_dict(state_dict, self.operand_dtype)
        self.logit_scale_exp = float(scale)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.init()
            check(self._L.plip_create_ex(blob.data_ptr(), blob.numel(), C.c_float(scale), self.device.index,
                                         int(max_micro_batch), fmt, C.byref(h)), "plip_create_ex")
        self._h = h
        self.max_micro_batch = int(max_micro_batch)

    def set_text_pooling(self, no_eos_argmax: bool) -> None:
        """Captions without an eos token: pool position 0 (HF, ``eos_token_id == 49407``; default) or the first argmax
        of the ids (legacy HF configs with ``eos_token_id == 2``, OpenAI clip).  See ``plip_set_text_pooling``."""
        check(self._L.plip_set_text_pooling(self._h, int(bool(no_eos_argmax))), "plip_set_text_pooling")

    def set_last_layer_pruning(self, on: bool) -> None:
        """Embedding calls run the last encoder layer's out_proj / LayerNorm 2 / MLP on the pooled rows only (CLS, first
        eos) — same embeddings, less work; hidden-state requests are never pruned.  See ``plip_set_last_layer_pruning``."""
        check(self._L.plip_set_last_layer_pruning(self._h, int(bool(on))), "plip_set_last_layer_pruning")

    @property
    def last_layer_pruning(self) -> bool:
        return bool(self._L.plip_last_layer_pruning(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.plip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ---------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t.contiguous()

    def upload_async(self, t: torch.Tensor):
        """Start copying a host tensor to the device on the engine's copy stream; returns ``(device tensor, event)``.
        The consumer stream must ``wait_event(event)`` before the first use (pinned sources overlap with compute)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            d = t.contiguous().to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        d.record_stream(cur)
        return d, ev

    # ---- device-tensor API -------------------------------------------------------------------
    @torch.no_grad()
    def encode_images(self, pixels: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """``get_image_features``: ``[n,3,224,224]`` f32/bf16 or ``[n,224,224,3]`` u8 -> ``[n,512]`` f32 (device)."""
        fmt = _pixel_format(pixels)
        n = int(pixels.shape[0])
        if n == 0:
            return torch.empty(0, EMBED_DIM, device=self.device)
        pixels = self._dev(pixels)
        out = torch.empty(n, EMBED_DIM, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self._L.plip_encode_images(self._h, pixels.data_ptr(), fmt, n, out.data_ptr(), int(normalize),
                                             self._stream()), "plip_encode_images")
        return out

    @torch.no_grad()
    def encode_text(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                    normalize: bool = False, prefix_len: Optional[int] = None) -> torch.Tensor:
        """``get_text_features``: ids ``[n,<=77]`` int32/int64 (+ optional mask) -> ``[n,512]`` f32 (device).

        ``prefix_len``: process only the first ``prefix_len`` positions of every row.  Exact whenever every
        caption's first eos lies inside the prefix (causal attention); the host path (``encode_text_host``)
        finds the longest caption itself, a device caller can pass it to avoid a sync."""
        n, s = _check_ids(input_ids, attention_mask)
        p_len = s if prefix_len is None else int(prefix_len)
        if not 1 <= p_len <= s:
            raise ValueError(f"prefix_len {p_len} out of [1, {s}]")
        if n == 0:
            return torch.empty(0, EMBED_DIM, device=self.device)
        idt = _ids_dtype(input_ids.dtype)
        ids = self._dev(input_ids)
        mask = None
        if attention_mask is not None:
            mask = self._dev(attention_mask.to(input_ids.dtype))
        out = torch.empty(n, EMBED_DIM, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self._L.plip_encode_text_prefix(self._h, ids.data_ptr(), idt,
                                                  mask.data_ptr() if mask is not None else None, n, s, p_len,
                                                  out.data_ptr(), int(normalize), self._stream()), "plip_encode_text")
        return out

    @torch.no_grad()
    def similarity(self, image_embeds: torch.Tensor, text_embeds: torch.Tensor, scale: Optional[float] = None,
                   normalize_image: bool = True, normalize_text: bool = True) -> torch.Tensor:
        """``logits_per_image[n,m] = scale * norm(image) @ norm(text).T`` (TF:923-930), fp32."""
        a = self._dev(image_embeds.to(torch.float32))
        b = self._dev(text_embeds.to(torch.float32))
        if a.shape[-1] != EMBED_DIM or b.shape[-1] != EMBED_DIM:
            raise ValueError("embeddings must have 512 columns")
        n, m = int(a.shape[0]), int(b.shape[0])
        ld = (m + 127) // 128 * 128          # the tensor-core path writes whole 128-column tiles
        out = torch.empty(n, ld, device=self.device, dtype=torch.float32)
        if n == 0 or m == 0:
            return out[:, :m]
        s = self.logit_scale_exp if scale is None else float(scale)
        with torch.cuda.device(self.device):
            check(self._L.plip_similarity(a.data_ptr(), n, b.data_ptr(), m, C.c_float(s), int(normalize_image),
                                          int(normalize_text), out.data_ptr(), ld, self._stream()), "plip_similarity")
        return out[:, :m]

    @torch.no_grad()
    def similarity_topk(self, query: torch.Tensor, space: torch.Tensor, k: int, scale: float = 1.0,
                        normalize_query: bool = True, normalize_space: bool = False):
        """Fused scores + top-k over ``space`` rows: returns ``(idx int32 [n,k], val f32 [n,k])``, descending."""
        return similarity_topk(query, space, k, scale, normalize_query, normalize_space, device=self.device)

    @torch.no_grad()
    def l2_normalize_(self, x: torch.Tensor) -> torch.Tensor:
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        if x.shape[0]:
            with torch.cuda.device(self.device):
                check(self._L.plip_l2_normalize(x.data_ptr(), int(x.shape[0]), int(x.shape[1]), self._stream()),
                      "plip_l2_normalize")
        return x

    @torch.no_grad()
    def resize_crop(self, src: torch.Tensor, descs: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Packed RGB uint8 images on the device + host descriptors (``preprocess.pack_rgb``) -> uint8 tiles
        ``[n,224,224,3]``; Pillow-exact bicubic resize and crop (``plip_resize_crop_u8``)."""
        from .preprocess import RESIZE_DESC_DTYPE
        assert src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous() and src.dim() == 1
        descs = np.ascontiguousarray(descs, dtype=RESIZE_DESC_DTYPE)
        n = int(descs.shape[0])
        if out is None:
            out = torch.empty((n, 224, 224, 3), device=src.device, dtype=torch.uint8)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == n * 224 * 224 * 3
        if n:
            with torch.cuda.device(src.device):
                check(self._L.plip_resize_crop_u8(src.data_ptr(), int(src.numel()), descs.ctypes.data, n,
                                                  out.data_ptr(), self._stream()), "plip_resize_crop_u8")
        return out

    # ---- host-buffer API (copies inside the call) ------------------------------------------------
    def encode_images_host(self, pixels: Union[np.ndarray, torch.Tensor], normalize: bool = False,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host array in, host ``[n,512]`` f32 tensor out; H2D/D2H pipelined inside the C call."""
        fmt = _pixel_format(pixels)
        t = torch.from_numpy(np.ascontiguousarray(pixels)) if isinstance(pixels, np.ndarray) else pixels.contiguous()
        assert not t.is_cuda, "encode_images_host takes host memory; use encode_images for device tensors"
        n = int(t.shape[0])
        if out is None:
            out = torch.empty(n, EMBED_DIM, dtype=torch.float32)
        if n:
            check(self._L.plip_encode_images_host(self._h, t.data_ptr(), fmt, n, out.data_ptr(), int(normalize)),
                  "plip_encode_images_host")
        return out

    def encode_text_host(self, input_ids: Union[np.ndarray, torch.Tensor], attention_mask=None,
                         normalize: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        ids = torch.from_numpy(np.ascontiguousarray(input_ids)) if isinstance(input_ids, np.ndarray) else input_ids.contiguous()
        n, s = _check_ids(ids, attention_mask)
        idt = _ids_dtype(ids.dtype)
        mask = None
        if attention_mask is not None:
            mask = (torch.from_numpy(np.ascontiguousarray(attention_mask)) if isinstance(attention_mask, np.ndarray)
                    else attention_mask).to(ids.dtype).contiguous()
        if out is None:
            out = torch.empty(n, EMBED_DIM, dtype=torch.float32)
        if n:
            check(self._L.plip_encode_text_host(self._h, ids.data_ptr(), idt, mask.data_ptr() if mask is not None else None,
                                                n, s, out.data_ptr(), int(normalize)), "plip_encode_text_host")
        return out

    # ---- in-step kernel timing (bench.py) ------------------------------------------------------
    def profile(self, on: bool) -> None:
        """Bracket every kernel launch of the following tower calls with CUDA events (``plip_profile_enable``)."""
        check(self._L.plip_profile_enable(self._h, int(bool(on))), "plip_profile_enable")

    def profile_read(self):
        """Rows ``{name, launches, total_ms, flops, bytes}`` aggregated per (tower, kernel role) since ``profile(True)``."""
        from ._lib import KernelTime
        buf = (KernelTime * 64)()
        cnt = C.c_int(0)
        check(self._L.plip_profile_read(self._h, buf, 64, C.byref(cnt)), "plip_profile_read")
        return [{"name": buf[i].name.decode(), "launches": int(buf[i].launches), "total_ms": float(buf[i].total_ms),
                 "flops": float(buf[i].flops), "bytes": float(buf[i].bytes)} for i in range(min(cnt.value, 64))]

    # ---- test hook ---------------------------------------------------------------------------
    @torch.no_grad()
    def hidden_states(self, tower: str, inputs: torch.Tensor, num_layers: int,
                      attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Residual stream after ``num_layers`` encoder layers (fp32), for layer-wise parity tests."""
        x = self._dev(inputs)
        n = int(x.shape[0])
        if tower == "vision":
            fmt, t, shape = _pixel_format(x), 0, (n, 50, 768)
        else:
            fmt, t, shape = _ids_dtype(x.dtype), 1, (n, 77, 512)
            assert x.shape[1] == 77
        mask = self._dev(attention_mask.to(x.dtype)) if attention_mask is not None else None
        out = torch.empty(shape, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self._L.plip_dbg_hidden_states(self._h, t, x.data_ptr(), fmt,
                                                 mask.data_ptr() if mask is not None else None, n, int(num_layers),
                                                 out.data_ptr(), self._stream()), "plip_dbg_hidden_states")
        return out
