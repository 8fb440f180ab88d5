"""Built-in CLIP BPE tokenizer against ``transformers.CLIPTokenizer`` on a synthetic merge table
(no real vocabulary is on this box), plus the two padding conventions of the reference's call sites
(``plip.py:57-58`` HF processor; ``reproducibility/embedders/plip.py:65`` ``clip.tokenize(..., truncate=True)``)."""
import collections
import gzip
import json

import numpy as np
import pytest

from plip_b200.tokenizer import ClipTokenizer, _PATTERN, base_vocab, bytes_to_unicode, find_tokenizer

CORPUS = """an h&e image patch of colorectal adenocarcinoma epithelium tumor stroma lymphocytes mucosa debris adipose
tissue normal colon mucosa smooth muscle cancer-associated stroma benign malignant glands nuclei pleomorphism mitotic
figures 40x magnification squamous cell carcinoma of the lung, poorly differentiated; necrosis present. don't can't
it's we're they've i'm you'll he'd the quick brown fox jumps over the lazy dog 0123456789 2023 ki-67 her2+ 3.5mm
café naïve résumé"""

CAPTIONS = ["An H&E image patch of colorectal adenocarcinoma epithelium", "  Tumor   stroma\tlymphocytes!! ",
            "don't you'll HE'D", "café naïve 3.5mm Ki-67 HER2+ 2023", "", "x" * 300,
            "squamous cell carcinoma of the lung, poorly differentiated; necrosis present. " * 5,
            "日本語 テスト ünïcödé", "a<|endoftext|>b", "Ki-67 &amp; HER2 &lt;3", "normal\ncolon\r\nmucosa",
            "ΟΔΟΣ ΣΟΦΟΣ, İstanbul STRASSE ẞ", "a\x1fb\x1cc\x85d\u2003e", "cafe\u0301 nai\u0308ve"]


def _train_merges(n=400):
    """A tiny BPE learner over CORPUS (most frequent pair first, ties by symbol order)."""
    bu = bytes_to_unicode()
    cnt = collections.Counter()
    for t in _PATTERN.findall(CORPUS):
        s = [bu[b] for b in t.encode()]
        s[-1] += "</w>"
        cnt[tuple(s)] += 1
    merges = []
    for _ in range(n):
        pc = collections.Counter()
        for w, c in cnt.items():
            for p in zip(w[:-1], w[1:]):
                pc[p] += c
        if not pc:
            break
        (a, b), _c = max(pc.items(), key=lambda kv: (kv[1], kv[0]))
        merges.append((a, b))
        new = collections.Counter()
        for w, c in cnt.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        cnt = new
    return merges


@pytest.fixture(scope="module")
def merges():
    return _train_merges()


@pytest.fixture(scope="module")
def tok(merges):
    return ClipTokenizer.from_merges(merges)


def test_vocab_layout(tok, merges):
    # OpenAI layout: 256 byte symbols, 256 with </w>, one token per merge, then <|startoftext|>, <|endoftext|>
    assert len(base_vocab()) == 512 and len(set(base_vocab())) == 512
    assert len(tok.encoder) == 512 + len(merges) + 2
    assert tok.bos_token_id == len(tok.encoder) - 2 and tok.eos_token_id == len(tok.encoder) - 1
    assert tok.encoder["".join(merges[0])] == 512


def test_matches_transformers_clip_tokenizer(tok, merges):
    tr = pytest.importorskip("transformers")
    hf = tr.CLIPTokenizer(vocab=dict(tok.encoder), merges=list(merges))
    assert (hf.bos_token_id, hf.eos_token_id) == (tok.bos_token_id, tok.eos_token_id)
    ours = tok(CAPTIONS, max_length=77, padding="max_length", truncation=True)
    ref = hf(CAPTIONS, max_length=77, padding="max_length", truncation=True)
    assert ours["input_ids"].tolist() == ref["input_ids"]
    assert ours["attention_mask"].tolist() == ref["attention_mask"]
    assert ours["input_ids"].dtype == np.int64 and ours["input_ids"].shape == (len(CAPTIONS), 77)
    for t in CAPTIONS[:4]:
        assert tok.encode(t) == hf.encode(t, add_special_tokens=False)
    longest = tok(CAPTIONS[:4], padding="longest")
    ref2 = hf(CAPTIONS[:4], padding="longest")
    assert longest["input_ids"].tolist() == ref2["input_ids"]
    pt = tok(CAPTIONS[:2], return_tensors="pt")
    assert pt["input_ids"].shape == (2, 77) and str(pt["input_ids"].dtype) == "torch.int64"


def test_openai_tokenize_convention(tok):
    """``clip.tokenize``: int32, zero padding, eot forced on truncation, RuntimeError when too long."""
    out = tok.tokenize(CAPTIONS, truncate=True)
    hf_style = tok(CAPTIONS)["input_ids"]
    mask = tok(CAPTIONS)["attention_mask"]
    assert out.dtype == np.int32 and out.shape == (len(CAPTIONS), 77)
    # clip.tokenize also un-escapes HTML entities, lower-cases with str.lower() (word-final sigma) and treats
    # U+001C..U+001F as white space: compare the two conventions on the captions free of those
    plain = [i for i, c in enumerate(CAPTIONS) if "&" not in c and "Σ" not in c and "\x1f" not in c]
    assert len(plain) >= 10
    for row in out:                                                   # zero padding after the closing eos
        e = int(np.flatnonzero(row == tok.eos_token_id).max())
        assert not row[e + 1:].any()
    assert np.array_equal(out[plain], (hf_style * mask)[plain])       # same ids, zeros where HF pads eos
    amp = CAPTIONS.index("Ki-67 &amp; HER2 &lt;3")
    assert np.array_equal(out[amp], tok.tokenize(["Ki-67 & HER2 <3"])[0])
    assert (out[:, 0] == tok.bos_token_id).all()
    assert out[5, 76] == tok.eos_token_id                                            # 300-character caption: truncated
    with pytest.raises(RuntimeError, match="too long for context length 77"):
        tok.tokenize([CAPTIONS[5]])
    assert tok.tokenize("tumor").shape == (1, 77)
    # first-eos pooling position (what the text tower uses) agrees between the two conventions
    assert np.array_equal((out == tok.eos_token_id).argmax(1)[plain], (hf_style == tok.eos_token_id).argmax(1)[plain])


def test_openai_flavour_applies_ftfy_default_fixes(tok):
    """clip.tokenize cleans captions with ftfy.fix_text (embedders/plip.py:65): curly quotes, ligatures, fullwidth
    characters and control characters must tokenise like their plain spellings (Twitter-sourced OpenPath
    captions are full of them).  Vectors: ftfy 6.x default configuration (uncurl_quotes, fix_latin_ligatures,
    fix_character_width, remove_control_chars, fix_line_breaks)."""
    from plip_b200.tokenizer import fix_text
    pairs = [("It\u2019s a \u201cfine\u201d tumor", "It's a \"fine\" tumor"),
             ("\ufb01brosis and in\ufb02ammation", "fibrosis and inflammation"),
             ("\uff28\uff06\uff25 stain\u3000image", "H&E stain image"),
             ("mitotic\x07 figure\ufeff", "mitotic figure"),
             ("line one\u2028line two", "line one\nline two")]
    for raw, plain in pairs:
        assert fix_text(raw) == plain, (raw, fix_text(raw))
        assert np.array_equal(tok.tokenize([raw])[0], tok.tokenize([plain])[0])
    assert fix_text("Ki-67 &amp; HER2") == "Ki-67 & HER2"                   # unescape_html


def test_decode_round_trip(tok):
    for t in ["tumor stroma", "café naïve"]:
        assert tok.decode(tok.encode(t)).strip() == t
    # every pre-token ends in </w>, so punctuation / digits come back space-separated (as with clip's decoder)
    assert tok.decode(tok.encode("ki-67 her2+")).strip() == "ki - 6 7 her 2 +"


def test_asset_loading(tmp_path, tok, merges):
    # HF layout
    d = tmp_path / "hf"
    d.mkdir()
    (d / "vocab.json").write_text(json.dumps(tok.encoder), encoding="utf-8")
    (d / "merges.txt").write_text("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n", encoding="utf-8")
    a = ClipTokenizer.from_pretrained(str(d))
    # OpenAI layout: header line + merges (the vocabulary is derived)
    g = tmp_path / "bpe_simple_vocab_16e6.txt.gz"
    with gzip.open(g, "wb") as f:
        f.write(('"bpe_simple_vocab_16e6.txt#version: 0.2\n' + "\n".join(" ".join(m) for m in merges) + "\n").encode())
    b = ClipTokenizer.from_openai_bpe(str(g), vocab_size=512 + len(merges) + 2)
    c = find_tokenizer(None, str(tmp_path / "missing"), str(tmp_path))     # directory holding the .gz
    assert c is not None and len(c.encoder) >= 512 + 2
    for t in CAPTIONS:
        assert a.encode(t) == tok.encode(t) == b.encode(t)
    assert find_tokenizer(None, str(tmp_path / "missing")) is None
    with pytest.raises(ValueError):
        ClipTokenizer({"a": 0}, [])


def test_embedder_uses_builtin_tokenizer(tmp_path, tok, merges, monkeypatch):
    """``CLIPEmbedder`` without the ``clip`` package: merge table found through ``$PLIP_B200_TOKENIZER``."""
    import torch
    from plip_b200 import embedders as E
    g = tmp_path / "bpe_simple_vocab_16e6.txt.gz"
    with gzip.open(g, "wb") as f:
        f.write(("header\n" + "\n".join(" ".join(m) for m in merges) + "\n").encode())
    monkeypatch.setenv("PLIP_B200_TOKENIZER", str(g))
    fn = E._default_tokenize(None)
    assert fn is not None
    ids = fn(["tumor stroma", "normal colon mucosa"])
    assert isinstance(ids, torch.Tensor) and ids.shape == (2, 77) and ids.dtype == torch.int32
    monkeypatch.delenv("PLIP_B200_TOKENIZER")
    try:
        import clip  # noqa: F401
    except Exception:  # noqa: BLE001
        assert E._default_tokenize(None) is None


@pytest.mark.gpu
def test_plip_encode_text_strings(state_dict, tok):
    """``PLIP.encode_text(List[str])`` end to end: built-in tokenizer -> ids -> text tower == encode_token_ids."""
    import torch
    from plip_b200 import PLIP
    from plip_b200.synthetic import EOS
    # synthetic weights have the real 49408-row embedding table: map the toy ids into it (eos must be 49407)
    class Shifted(ClipTokenizer):
        def __call__(self, text, **kw):
            enc = tok(text, **{k: v for k, v in kw.items() if k != "return_tensors"})
            ids = enc["input_ids"].copy()
            ids[ids == tok.eos_token_id] = EOS
            ids[ids == tok.bos_token_id] = EOS - 1
            return {"input_ids": torch.from_numpy(ids), "attention_mask": torch.from_numpy(enc["attention_mask"])}
    shifted = Shifted(tok.encoder, list(tok.bpe_ranks))
    plip = PLIP.from_state_dict(state_dict, max_micro_batch=64, tokenizer=shifted)
    caps = ["An H&E image patch of tumor", "An H&E image patch of normal colon mucosa", "stroma"]
    a = plip.encode_text(caps, batch_size=2)
    enc = shifted(caps)
    b = plip.encode_token_ids(enc["input_ids"], enc["attention_mask"])
    assert a.shape == (3, 512) and a.dtype == np.float32
    cos = (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
    assert (1 - cos).max() <= 1e-5
    plip.tokenizer = None
    with pytest.raises(RuntimeError, match="no tokenizer available"):
        plip.encode_text(caps, batch_size=2)


def test_random_strings_match_transformers(tok, merges):
    """Property test: arbitrary unicode captions (letters, digits, punctuation, whitespace, emoji, combining marks)
    tokenise exactly like ``transformers.CLIPTokenizer`` on the same merge table."""
    tr = pytest.importorskip("transformers")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import strategies as st
    hf = tr.CLIPTokenizer(vocab=dict(tok.encoder), merges=list(merges))
    alphabet = st.one_of(st.sampled_from(list("tumor stroma adenocarcinoma h&e 0123456789 .,;:!?'\"-+()/\t\n  ")),
                         # assigned code points only: what an unassigned one is depends on the Unicode version of
                         # each regex engine (Python `regex` vs the Rust crate behind transformers)
                         # and no arbitrary combining marks (canonical reordering of marks added in recent Unicode
                         # versions differs between normalisers); the common ones are listed explicitly
                         st.sampled_from(list("\u0301\u0308\u0327\u0303\x1c\x1f\x85\xa0\u2003\u3000\u200b\ufeffΣİẞ")),
                         st.characters(blacklist_categories=("Cs", "Cn", "Co", "Mn", "Mc", "Me"), max_codepoint=0x1F9FF))

    @hyp.settings(max_examples=300, deadline=None, derandomize=True)
    @hyp.given(st.text(alphabet=alphabet, max_size=120))
    def check(text):
        ours = tok([text])["input_ids"][0].tolist()
        ref = hf([text], max_length=77, padding="max_length", truncation=True)["input_ids"][0]
        assert ours == ref, repr(text)

    check()
