#!/usr/bin/env python3
"""Derive the LaTok character-class ranges from the reference's generated header.

Run in the build container only (it needs /root/reference).  It reads the
reference's type-record / index tables as *data* (latok.h:66-570 records,
:1814 SHIFT, :1815-2422 index1, :2424-4173 index2), projects the 20 flag bits
of every code point onto the 12 base features the hot path consumes
(latok.c:87-98) and writes a run-length list

    <first-cp-hex> <last-cp-hex> <12-bit-feature-hex>

to latok_b200/data/ucd11_latok_classes.txt.  That file is the single source of
truth for both the CUDA class table (tools/gen_tables.py) and the CPU oracle;
nothing else in this repo reads the reference header.

Feature bit order = feature column order (offsets.py:24-35):
  0 ALPHA 1 ALPHA_NUM 2 NUM 3 LOWER 4 UPPER 5 SPACE 6 SYMBOL 7 TWITTER
  8 '@'   9 ':'      10 '/' 11 '.'
"""
import re
import sys
from pathlib import Path

REF_H = Path("/root/reference/latok/core/src/latok/latok.h")
OUT = Path(__file__).resolve().parent.parent / "latok_b200" / "data" / "ucd11_latok_classes.txt"

# flag masks, latok.h:3-22
F_ALPHA, F_LOWER, F_SPACE, F_UPPER = 0x01, 0x08, 0x20, 0x80
F_PRINTABLE, F_NUMERIC = 0x400, 0x800
F_SPECIALS, F_AT, F_COLON, F_SLASH, F_PERIOD = 0x8000, 0x10000, 0x20000, 0x40000, 0x80000


def base_features(flags: int) -> int:
    """12 base feature bits of one code point, latok.c:87-98."""
    alpha = 1 if flags & F_ALPHA else 0
    num = 1 if flags & F_NUMERIC else 0
    alnum = alpha | num
    space = 1 if flags & F_SPACE else 0
    symbol = 1 if (flags & F_PRINTABLE) and not alnum and not space else 0
    bits = [
        alpha, alnum, num,
        1 if flags & F_LOWER else 0,
        1 if flags & F_UPPER else 0,
        space, symbol,
        1 if flags & F_SPECIALS else 0,
        1 if flags & F_AT else 0,
        1 if flags & F_COLON else 0,
        1 if flags & F_SLASH else 0,
        1 if flags & F_PERIOD else 0,
    ]
    return sum(b << i for i, b in enumerate(bits))


def parse_header(path: Path):
    text = path.read_text()
    rec_block = re.search(r"_TtUnicode_TypeRecords\[\]\s*=\s*\{(.*?)\n\};", text, re.S).group(1)
    flags = [int(m.group(1)) for m in re.finditer(r"\{[^{}]*?,\s*(-?\d+)\s*\}", rec_block)]
    shift = int(re.search(r"#define\s+SHIFT\s+(\d+)", text).group(1))
    idx1 = [int(x) for x in re.findall(r"\d+", re.search(r"index1\[\]\s*=\s*\{(.*?)\};", text, re.S).group(1))]
    idx2 = [int(x) for x in re.findall(r"\d+", re.search(r"index2\[\]\s*=\s*\{(.*?)\};", text, re.S).group(1))]
    return flags, shift, idx1, idx2


def main():
    flags, shift, idx1, idx2 = parse_header(REF_H)
    mask = (1 << shift) - 1
    feats = []
    for cp in range(0x110000):
        rec = idx2[(idx1[cp >> shift] << shift) + (cp & mask)]
        feats.append(base_features(flags[rec]))
    runs = []
    start = 0
    for cp in range(1, 0x110000 + 1):
        if cp == 0x110000 or feats[cp] != feats[start]:
            runs.append((start, cp - 1, feats[start]))
            start = cp
    distinct = sorted({f for _, _, f in runs})
    with OUT.open("w") as fh:
        fh.write("# LaTok base-feature classes per code point, UCD 11.0.0 + LaTok's 5 special flags.\n")
        fh.write("# Derived data (tools/derive_classes.py); columns: first last features(hex, bit i = feature column i)\n")
        fh.write(f"# runs={len(runs)} distinct_feature_words={len(distinct)}\n")
        for a, b, f in runs:
            fh.write(f"{a:X} {b:X} {f:03X}\n")
    print(f"records={len(flags)} shift={shift} index1={len(idx1)} index2={len(idx2)}")
    print(f"runs={len(runs)} distinct={len(distinct)}: {[hex(d) for d in distinct]}")
    print(f"wrote {OUT}")


if __name__ == "__main__":
    sys.exit(main())
