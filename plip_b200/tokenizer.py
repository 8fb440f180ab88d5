r"""CLIP byte-level BPE tokenizer (captions -> ``[n,77]`` token ids) without the ``clip`` / ``tokenizers`` packages.

The reference tokenises on the host in two flavours:

* ``plip.py:57-58`` — ``CLIPProcessor(text=…, max_length=77, padding="max_length", truncation=True)``: ids
  ``<|startoftext|> … <|endoftext|>``, truncated to 77 keeping the final ``<|endoftext|>``, padded with
  ``<|endoftext|>`` (49407), plus an ``attention_mask``  -> :meth:`ClipTokenizer.__call__`;
* ``reproducibility/embedders/plip.py:65`` — ``clip.tokenize(captions, truncate=True)`` (OpenAI ``clip`` package,
  not installed here): same ids, padded with **0**, int32 ``[n,77]``, ``RuntimeError`` when a caption is too long
  and ``truncate`` is false  -> :meth:`ClipTokenizer.tokenize`.

Both read the same published algorithm (OpenAI ``simple_tokenizer.py`` / ``transformers`` ``CLIPTokenizer``):
NFC + whitespace collapse + lower-casing, a regex pre-split, bytes mapped to printable unicode, greedy
lowest-rank-first pair merging with an ``</w>`` end-of-word marker.  They differ in details that the two front ends
keep apart: ``clip.tokenize`` un-escapes HTML entities, uses Python's ``\s`` (which includes U+001C..U+001F) and
``str.lower()`` (word-final sigma); ``transformers`` uses Unicode ``White_Space`` and lower-cases character by
character.  The vocabulary / merge table are assets of the
checkpoint (``vocab.json`` + ``merges.txt``, or OpenAI's ``bpe_simple_vocab_16e6.txt.gz``) — none is on this box, so
``tests/test_tokenizer.py`` pins the implementation against ``transformers.CLIPTokenizer`` on a synthetic merge table.
Host-side plumbing only: the device path starts at ``input_ids``.
"""
from __future__ import annotations

import gzip
import html
import json
import os
import unicodedata
from functools import lru_cache
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import regex as re

BOS_TOKEN = "<|startoftext|>"
EOS_TOKEN = "<|endoftext|>"
CONTEXT_LENGTH = 77

# OpenAI's simple_tokenizer runs on Python's `regex` (\s = str.isspace(): includes the separators U+001C..U+001F);
# transformers' CLIPTokenizer runs on the Rust regex crate (\s = Unicode White_Space, which excludes them).
_PATTERN = re.compile(
    r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)
_PATTERN_HF = re.compile(
    r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\p{White_Space}\p{L}\p{N}]+""")
_WS_HF = re.compile(r"\p{White_Space}+")


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """Reversible byte -> printable unicode character table of GPT-2 / CLIP byte-level BPE."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, (chr(c) for c in cs)))


def base_vocab() -> List[str]:
    """The 512 single-byte symbols (plain and with ``</w>``) every CLIP vocabulary starts with."""
    v = list(bytes_to_unicode().values())
    return v + [s + "</w>" for s in v]


# ---- ftfy.fix_text, as OpenAI clip's basic_clean calls it (``embedders/plip.py:65`` -> clip.tokenize) ------------------
# ``ftfy`` is a third-party package (un-pinned by the reference, absent from this image).  When it is importable it is
# used; otherwise the deterministic character-level fixes of its DEFAULT configuration (ftfy 6.x ``TextFixerConfig``:
# fix_latin_ligatures, fix_character_width, uncurl_quotes, fix_line_breaks, remove_terminal_escapes,
# remove_control_chars, NFC) are restated here.  Its mojibake repair (fix_encoding, a heuristic search over
# mis-decodings) is NOT restated: text that was decoded with the wrong codec tokenises differently without ftfy.
_LIGATURES = {"\uFB00": "ff", "\uFB01": "fi", "\uFB02": "fl", "\uFB03": "ffi", "\uFB04": "ffl", "\uFB05": "\u017Ft", "\uFB06": "st",
              "\u0132": "IJ", "\u0133": "ij", "\u0149": "\u02BCn", "\u01F1": "DZ", "\u01F2": "Dz", "\u01F3": "dz",
              "\u01C4": "D\u017D", "\u01C5": "D\u017E", "\u01C6": "d\u017E", "\u01C7": "LJ", "\u01C8": "Lj", "\u01C9": "lj",
              "\u01CA": "NJ", "\u01CB": "Nj", "\u01CC": "nj"}
_FIX_TABLE = {ord(k): v for k, v in _LIGATURES.items()}
_FIX_TABLE.update({c: "'" for c in [0x02BC] + list(range(0x2018, 0x201C))})        # uncurl_quotes: single
_FIX_TABLE.update({c: '"' for c in range(0x201C, 0x2020)})                         # uncurl_quotes: double
_FIX_TABLE.update({c: chr(c - 0xFEE0) for c in range(0xFF01, 0xFF5F)})             # fix_character_width: fullwidth ASCII
_FIX_TABLE[0x3000] = " "                                                           # ideographic space
_FIX_TABLE.update({0x2028: "\n", 0x2029: "\n", 0x0085: "\n"})                      # fix_line_breaks
_FIX_TABLE.update({c: None for c in list(range(0x00, 0x09)) + [0x0B] + list(range(0x0E, 0x20)) + [0x7F, 0xFEFF]
                   + list(range(0x206A, 0x2070)) + list(range(0xFFF9, 0xFFFD))})   # remove_control_chars
_ANSI_RE = re.compile(r"\033\[((?:\d|;)*)([a-zA-Z])")                              # remove_terminal_escapes


def fix_text(text: str) -> str:
    """``ftfy.fix_text`` if the package is present, else its deterministic default fixes (see above)."""
    try:
        import ftfy                                  # noqa: PLC0415 - optional dependency of the reference
        return ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(text)
    text = _ANSI_RE.sub("", text.replace("\r\n", "\n").replace("\r", "\n"))
    return unicodedata.normalize("NFC", text.translate(_FIX_TABLE))


def _clean(text: str, openai: bool) -> str:
    if openai:   # clip.simple_tokenizer: basic_clean (ftfy.fix_text + 2x html.unescape) + whitespace_clean + lower
        text = html.unescape(html.unescape(fix_text(text))).strip()
        return re.sub(r"\s+", " ", text).strip().lower()
    # transformers: normalizers.Sequence([NFC(), Replace(Regex(r"\s+"), " "), Lowercase()]).  `tokenizers` lowercases
    # character by character, i.e. without str.lower()'s context rule for a word-final capital sigma.
    return "".join(c.lower() for c in _WS_HF.sub(" ", unicodedata.normalize("NFC", text)))


class ClipTokenizer:
    """``vocab``: token string -> id; ``merges``: ranked list of symbol pairs."""

    def __init__(self, vocab: Dict[str, int], merges: Sequence[Tuple[str, str]]):
        self.encoder = dict(vocab)
        self.decoder = {i: t for t, i in self.encoder.items()}
        self.bpe_ranks = {tuple(m): i for i, m in enumerate(merges)}
        if BOS_TOKEN not in self.encoder or EOS_TOKEN not in self.encoder:
            raise ValueError(f"vocabulary lacks {BOS_TOKEN} / {EOS_TOKEN}")
        self.bos_token_id = self.encoder[BOS_TOKEN]
        self.eos_token_id = self.encoder[EOS_TOKEN]
        self.unk_token_id = self.eos_token_id                 # transformers: unk_token = "<|endoftext|>"
        self._byte = bytes_to_unicode()
        self._cache: Dict[str, Tuple[str, ...]] = {BOS_TOKEN: (BOS_TOKEN,), EOS_TOKEN: (EOS_TOKEN,)}

    # ---- constructors ----------------------------------------------------------------------------
    @classmethod
    def from_files(cls, vocab_json: str, merges_txt: str) -> "ClipTokenizer":
        """HuggingFace checkpoint layout (``vocab.json`` + ``merges.txt``, first line a ``#version`` header)."""
        with open(vocab_json, encoding="utf-8") as f:
            vocab = json.load(f)
        with open(merges_txt, encoding="utf-8") as f:
            lines = f.read().split("\n")
        if lines and lines[0].startswith("#"):
            lines = lines[1:]
        merges = [tuple(l.split()) for l in lines if l.strip()]
        return cls(vocab, merges)

    @classmethod
    def from_openai_bpe(cls, path: str, vocab_size: int = 49408) -> "ClipTokenizer":
        """OpenAI ``bpe_simple_vocab_16e6.txt.gz``: the vocabulary is *derived* from the merge list
        (512 byte symbols, one token per merge, then the two specials)."""
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(l.split()) for l in lines[1:vocab_size - 512 - 2 + 1]]
        return cls.from_merges(merges)

    @classmethod
    def from_merges(cls, merges: Sequence[Tuple[str, str]]) -> "ClipTokenizer":
        tokens = base_vocab() + ["".join(m) for m in merges] + [BOS_TOKEN, EOS_TOKEN]
        return cls({t: i for i, t in enumerate(tokens)}, merges)

    @classmethod
    def from_pretrained(cls, directory: str) -> "ClipTokenizer":
        """A checkpoint directory (or a single ``.txt.gz`` / ``.txt`` merge file)."""
        if os.path.isfile(directory):
            return cls.from_openai_bpe(directory)
        vj, mt = os.path.join(directory, "vocab.json"), os.path.join(directory, "merges.txt")
        if os.path.exists(vj) and os.path.exists(mt):
            return cls.from_files(vj, mt)
        for name in ("bpe_simple_vocab_16e6.txt.gz", "bpe_simple_vocab_16e6.txt"):
            if os.path.exists(os.path.join(directory, name)):
                return cls.from_openai_bpe(os.path.join(directory, name))
        raise FileNotFoundError(f"no vocab.json + merges.txt or bpe_simple_vocab_16e6.txt.gz under {directory}")

    # ---- BPE ---------------------------------------------------------------------------------------
    def _bpe(self, token: str) -> Tuple[str, ...]:
        hit = self._cache.get(token)
        if hit is not None:
            return hit
        word: Tuple[str, ...] = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = set(zip(word[:-1], word[1:]))
            first, second = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if (first, second) not in self.bpe_ranks:
                break
            out: List[str] = []
            i = 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    out.append(first + second)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
        self._cache[token] = word
        return word

    def encode(self, text: str, openai: bool = False) -> List[int]:
        """Caption -> token ids, without the start / end tokens.  ``openai`` selects ``clip.tokenize``'s text
        cleaning and whitespace definition instead of ``transformers.CLIPTokenizer``'s."""
        ids: List[int] = []
        for tok in (_PATTERN if openai else _PATTERN_HF).findall(_clean(text, openai)):
            if tok in (BOS_TOKEN, EOS_TOKEN):
                ids.append(self.encoder[tok])
                continue
            sym = "".join(self._byte[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder.get(piece, self.unk_token_id) for piece in self._bpe(sym))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        inv = {c: b for b, c in self._byte.items()}
        text = "".join(self.decoder[int(i)] for i in ids if int(i) not in (self.bos_token_id, self.eos_token_id))
        return bytearray(inv[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")

    # ---- batch front ends ---------------------------------------------------------------------------
    def tokenize(self, texts: Union[str, Sequence[str]], context_length: int = CONTEXT_LENGTH,
                 truncate: bool = False) -> np.ndarray:
        """``clip.tokenize``: int32 ``[n, context_length]``, zero padded; too-long captions raise unless
        ``truncate`` (then the last kept token becomes ``<|endoftext|>``)."""
        if isinstance(texts, str):
            texts = [texts]
        out = np.zeros((len(texts), context_length), dtype=np.int32)
        for i, t in enumerate(texts):
            ids = [self.bos_token_id] + self.encode(t, openai=True) + [self.eos_token_id]
            if len(ids) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {t} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = self.eos_token_id
            out[i, :len(ids)] = ids
        return out

    def __call__(self, text: Union[str, Sequence[str]], max_length: int = CONTEXT_LENGTH, padding: str = "max_length",
                 truncation: bool = True, return_tensors: Optional[str] = None, **_) -> Dict[str, object]:
        """``CLIPProcessor(text=…)`` / ``CLIPTokenizer(...)``: ``input_ids`` padded with ``<|endoftext|>`` and
        ``attention_mask`` (int64).  ``padding``: ``"max_length"`` or ``"longest"`` / ``True``."""
        if isinstance(text, str):
            text = [text]
        rows = []
        for t in text:
            ids = [self.bos_token_id] + self.encode(t) + [self.eos_token_id]
            if truncation and len(ids) > max_length:
                ids = ids[:max_length - 1] + [self.eos_token_id]
            rows.append(ids)
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        ids = np.full((len(rows), width), self.eos_token_id, dtype=np.int64)
        mask = np.zeros((len(rows), width), dtype=np.int64)
        for i, r in enumerate(rows):
            if len(r) > width:
                raise ValueError(f"caption {i} has {len(r)} tokens (> {width}) and truncation is off")
            ids[i, :len(r)] = r
            mask[i, :len(r)] = 1
        if return_tensors == "pt":
            import torch
            return {"input_ids": torch.from_numpy(ids), "attention_mask": torch.from_numpy(mask)}
        return {"input_ids": ids, "attention_mask": mask}


def find_tokenizer(*candidates: Optional[str]) -> Optional[ClipTokenizer]:
    """First loadable tokenizer among the given directories / files and ``$PLIP_B200_TOKENIZER``."""
    for c in list(candidates) + [os.environ.get("PLIP_B200_TOKENIZER")]:
        if not c:
            continue
        try:
            return ClipTokenizer.from_pretrained(c)
        except (FileNotFoundError, NotADirectoryError, ValueError):
            continue
    return None
