"""Token spans as trimmed byte ranges of the packed UTF-8 buffer against the reference's token
texts: the committed outputs of the reference's own tokenize() (tests/golden/reference_outputs.json) and the
oracle's restatement of `text[s:e].strip()` (default_tokenizer.py:151-158).  Needs a B200 (-m gpu)."""
import json
from pathlib import Path

import numpy as np
import pytest

import corpus
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from latok_b200.engine import Engine
    with Engine(0) as e:
        yield e


def _packed(engine, texts):
    from latok_b200.core.default_tokenizer import tokenize_packed
    from latok_b200.engine import pack_strings
    buf, off = pack_strings(texts)
    return tokenize_packed(buf, off, engine=engine)


def test_golden_token_texts(engine):
    recs = json.loads((Path(__file__).parent / "golden" / "reference_outputs.json").read_text())["records"]
    recs = [r for r in recs if r["text"]]
    pt = _packed(engine, [r["text"] for r in recs])
    for i, r in enumerate(recs):
        assert pt.tokens(i) == r["tokens"], r["text"]


@pytest.mark.parametrize("profile,seed", [("mixed", 11), ("ascii", 12), ("mixed", 13)])
def test_fuzz_token_texts(engine, profile, seed):
    texts = [t for t in corpus.fuzz_strings(seed, 1500, 200, profile) if t] + corpus.FIXTURES
    texts = [t for t in texts if t]
    pt = _packed(engine, texts)
    for i, t in enumerate(texts):
        assert pt.tokens(i) == oracle.tokens(t), t


def test_long_multibyte_strings_cross_blocks(engine):
    # strings far longer than the 512-byte counting blocks and the 128 KB scan groups, with every UTF-8 length
    rng = np.random.default_rng(5)
    alphabet = ["a", "B", " ", "é", "日", "😀", ",", "　", "@", "x", "ß", " ", "1"]
    texts = ["".join(rng.choice(alphabet, size=n)) for n in (5000, 70000, 300000, 3)]
    pt = _packed(engine, texts)
    for i, t in enumerate(texts):
        assert pt.tokens(i) == oracle.tokens(t)


def test_arrow_view_and_edge_cases(engine):
    texts = ["This is a #test!", "", " ", "a", "日本語 のテキスト、です。", "x" * 1000 + " y"]
    pt = _packed(engine, texts)
    want = [oracle.tokens(t) if t else [] for t in texts]
    assert [pt.tokens(i) for i in range(len(texts))] == want
    assert pt.to_arrow().to_pylist() == want
    # every range lies inside its string and ranges are ordered and disjoint
    from latok_b200.engine import pack_strings
    _, off = pack_strings(texts)
    for i in range(len(texts)):
        sp = pt.byte_spans[pt.tok_offsets[i]:pt.tok_offsets[i + 1]]
        assert np.all(sp[:, 0] < sp[:, 1]) and np.all(sp[1:, 0] >= sp[:-1, 1])
        if len(sp):
            assert sp[0, 0] >= off[i] and sp[-1, 1] <= off[i + 1]


def test_needs_spans(engine):
    from latok_b200.engine import SPLITS, pack_strings
    buf, off = pack_strings(["abc def"])
    engine.submit(buf, off, SPLITS)
    with pytest.raises(RuntimeError):
        engine.token_bytes()


def test_timing_cli_outfile_matches_reference_format(tmp_path):
    """tools/time_tokenizer.py (counterpart of scripts/timing/time_tokenizer.py:65-123): csv.gz in, one line of
    tab-separated tokens per row out."""
    import csv
    import gzip
    import subprocess
    import sys
    texts = [t for t in corpus.FIXTURES + corpus.fuzz_strings(21, 300, 120) if "\x00" not in t]
    texts = [t for t in texts if not any(0xD800 <= ord(c) <= 0xDFFF for c in t)] + ["", "   "]
    src = tmp_path / "in.csv.gz"
    with gzip.open(src, "wt", encoding="utf-8", newline="") as f:
        w = csv.writer(f)
        for i, t in enumerate(texts):
            w.writerow([i, json.dumps(t)])
    root = Path(__file__).resolve().parent.parent
    want = b"".join("\t".join(oracle.tokens(t.strip()) if t.strip() else []).encode("utf-8") + b"\n" for t in texts)
    for reader in ("native", "python"):        # latok_reader.cpp into pinned buffers / the csv + json modules
        out = tmp_path / f"out_{reader}.tsv"
        r = subprocess.run([sys.executable, str(root / "tools" / "time_tokenizer.py"), str(src), "--outfile", str(out),
                            "--batch", "64", "--batch-bytes", "65536", "--reader", reader], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert out.read_bytes() == want, reader
        stats = json.loads(r.stdout.strip().splitlines()[-1])
        assert stats["lines"] == len(texts) and stats["reader"] == reader


def test_device_resident_text_and_device_output(engine):
    """Text submitted from device memory, byte ranges written to a caller's device array (on_device = 1)."""
    import ctypes as C
    import torch
    from latok_b200 import _lib
    from latok_b200.engine import SPANS, pack_strings
    texts = [t for t in corpus.fuzz_strings(31, 800, 150, "mixed") if t]
    buf, off = pack_strings(texts)
    d_buf = torch.from_numpy(buf.copy()).cuda()
    d_off = torch.from_numpy(off.copy()).cuda()
    engine.submit_device(d_buf.data_ptr(), d_off.data_ptr(), len(texts), len(buf), SPANS)
    r = engine.fetch()
    d_out = torch.empty((r.n_tokens, 2), dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().latok_b200_fetch_token_bytes(engine._h, r.n_tokens, d_out.data_ptr(), 1))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    assert np.array_equal(got, engine.token_bytes())
    mv = memoryview(buf)
    for i in (0, 1, len(texts) // 2, len(texts) - 1):
        toks = [bytes(mv[b:e]).decode("utf-8", "surrogatepass") for b, e in got[r.tok_offsets[i]:r.tok_offsets[i + 1]]]
        assert toks == oracle.tokens(texts[i])
    # too small a caller array is rejected (the call takes its capacity), and so is a misaligned device pointer
    with pytest.raises(ValueError):
        _lib.check(_lib.load().latok_b200_fetch_token_bytes(engine._h, r.n_tokens - 1, d_out.data_ptr(), 1))
    with pytest.raises(ValueError):
        _lib.check(_lib.load().latok_b200_fetch_token_bytes(engine._h, r.n_tokens, d_out.data_ptr() + 8, 1))
