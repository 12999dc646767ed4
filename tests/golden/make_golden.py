#!/usr/bin/env python3
"""Generate the committed golden fixtures.  Runs ONLY in the build container (needs
/root/reference): everything written here is output of the reference itself.

  notebook_cell.json    the reference's one stored golden vector, copied out of the executed
                        notebook cell notebooks/scratch/LaTokenizer.ipynb:1263-1430 (input :1444)
  reference_outputs.json  per fixture string: feature matrix (25-bit word per character),
                        split mask, list(tokenize(text)), and for strings <= 127 chars the
                        featurize spans + feature vectors -- produced by the reference's own
                        latok.c (compiled unmodified, `make -C oracle ref`) driven by the
                        reference's own default_tokenizer.py
  fuzz_digest.json      sha256 digests of the same outputs over seeded fuzz corpora
  codepoint_classes.bin 12-bit base-feature word for every code point 0..0x10FFFF as read off
                        the reference's _gen_parse_matrix (zlib-compressed uint16)

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import re
import sys
import zlib
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import ref_driver  # noqa: E402
import corpus  # noqa: E402

FUZZ_SETS = [  # (seed, count, max_len, profile)
    (101, 4000, 60, "mixed"),
    (102, 4000, 120, "ascii"),
    (103, 4000, 100, "marks"),
    (104, 300, 700, "mixed"),
]


def pack_rows(m: np.ndarray):
    w = (m.astype(np.int64) << np.arange(25)).sum(axis=1)
    return [int(x) for x in w]


def notebook_cell():
    nb = json.load(open("/root/reference/notebooks/scratch/LaTokenizer.ipynb"))
    cell = nb["cells"][0]
    text = re.search(r'text = "(.*)"', "".join(cell["source"])).group(1)
    plain = "".join(cell["outputs"][0]["data"]["text/plain"])
    # the DataFrame repr is printed in column chunks; stitch them back by row index
    rows = {}
    columns = []
    for chunk in plain.split("\n\n"):
        lines = [ln for ln in chunk.split("\n") if ln.strip()]
        if not lines:
            continue
        header = lines[0].replace("\\", "").split()
        first = not columns
        for ln in lines[1:]:
            ln = ln.replace("\\", "")
            idx = int(ln.split()[0])
            if first:
                # "idx  <char>  values..." ; the char may be a space, so slice after the index
                vals = ln.split()[1:]
                need = len(header) - 1
                vals = vals[-need:]
                rows.setdefault(idx, []).extend(int(v) for v in vals)
            else:
                rows.setdefault(idx, []).extend(int(v) for v in ln.split()[1:])
        columns.extend(header[1:] if first else header)
    assert columns[0] == "Splits" and len(columns) == 26, columns
    n = len(text)
    data = np.array([rows[i] for i in range(n)], dtype=np.int64)
    assert data.shape == (n, 26)
    return {"source": "notebooks/scratch/LaTokenizer.ipynb:1263-1430", "text": text,
            "columns": columns, "splits": data[:, 0].tolist(), "matrix": data[:, 1:].tolist()}


def reference_record(rp, text):
    m = rp._gen_parse_matrix(text)
    splits = rp.gen_split_mask(m)
    rec = {"text": text, "matrix_words": pack_rows(m), "splits": [int(x) for x in splits],
           "tokens": list(rp.tokenize(text))}
    if len(text) <= 127:
        toks = list(rp.featurize(text))
        rec["feat_spans"] = [[int(t.start_idx), int(t.end_idx)] for t in toks]
        rec["feat_texts"] = [t.text for t in toks]
        rec["feats"] = [[int(v) for v in t.features] for t in toks]
    return rec


def digest(rp, strings):
    h_m, h_s, h_t = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    n_tok = 0
    for t in strings:
        if not t:
            continue
        m = rp._gen_parse_matrix(t)
        s = rp.gen_split_mask(m)
        toks = list(rp.tokenize(t))
        h_m.update(np.ascontiguousarray(m).tobytes())
        h_s.update(np.ascontiguousarray(s).tobytes())
        h_t.update(("\x00".join(toks) + "\x01").encode("utf-8", "surrogatepass"))
        n_tok += len(toks)
    return {"matrix": h_m.hexdigest(), "splits": h_s.hexdigest(), "tokens": h_t.hexdigest(), "n_tokens": n_tok}


def main():
    rp = ref_driver.ref_python()
    assert rp is not None, "needs /root/reference and a built oracle/_ref"
    ext = ref_driver.ext()

    cell = notebook_cell()
    # the compiled reference must reproduce its own stored cell before we trust it as a generator
    m = ext._gen_parse_matrix(cell["text"])
    assert m.tolist() == cell["matrix"], "compiled reference disagrees with the notebook cell (matrix)"
    assert rp.gen_split_mask(m).tolist() == cell["splits"], "compiled reference disagrees with the notebook cell (splits)"
    json.dump(cell, open(HERE / "notebook_cell.json", "w"), indent=0)

    recs = [reference_record(rp, t) for t in corpus.FIXTURES]
    json.dump({"generator": "tests/golden/make_golden.py", "records": recs},
              open(HERE / "reference_outputs.json", "w"), ensure_ascii=True, indent=0)

    dig = []
    for seed, count, max_len, profile in FUZZ_SETS:
        strings = corpus.fuzz_strings(seed, count, max_len, profile)
        d = digest(rp, strings)
        d.update(seed=seed, count=count, max_len=max_len, profile=profile)
        dig.append(d)
    json.dump({"generator": "tests/golden/make_golden.py", "sets": dig}, open(HERE / "fuzz_digest.json", "w"), indent=1)

    # every code point through the reference's _gen_parse_matrix (surrogates are legal in str)
    words = np.zeros(0x110000, dtype=np.uint16)
    step = 4096
    for base in range(0, 0x110000, step):
        s = "".join(chr(c) for c in range(base, base + step))
        mm = ext._gen_parse_matrix(s)
        words[base:base + step] = (mm[:, :12].astype(np.uint16) << np.arange(12, dtype=np.uint16)).sum(axis=1)
    (HERE / "codepoint_classes.bin").write_bytes(zlib.compress(words.tobytes(), 9))
    print("notebook cell ok; fixtures:", len(recs), "fuzz sets:", len(dig),
          "classes distinct:", len(set(words.tolist())))


if __name__ == "__main__":
    main()
