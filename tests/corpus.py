"""Shared test corpora: the fixture strings of BASELINE config #1 (SURVEY.md section 4 items 1-4 and
the quirk corpus of section 8c) and a seeded mixed-Unicode fuzz generator."""
from __future__ import annotations

import numpy as np

# strings the reference itself names (default_tokenizer.py:195-198, numpy_tokenizer.py:286,
# time_tokenizer.py:113) followed by the quirk corpus.
FIXTURES = [
    "This is a #test! Testing, Testing, 1 2 3",
    "can\u2019t wait to get my glasses back \U0001F913",
    "IKR!! IM LIKE \"WHERE'S MY DADDY AT? \U0001F440) https://t.co/jM3qLZijMc",
    "$#@^:a./",
    "This is a test line, just to get things warmed up...",
    " ",
    "a",
    "ab",
    "abc",
    "a b",
    "  ",
    "a  b",
    " a",
    "a ",
    "!",
    "!!",
    "! ",
    " !",
    "a@b",
    "#foo",
    " #a b,c",
    "#a,b c,d",
    "fooBar, a@b",
    "a@b.c,d@e.f fooBar, x-y z",
    "a@b,c@d,e@f one,two three,four five,six",
    "a@b,c@d,e@f  one,two   three,four five,six",
    "camelCaseWord HTTPServer XMLHttpRequest",
    "see http://foo.com/bar?x=1#frag and more",
    "email me at joe.blow@example.com or .@joe ok",
    ".@joe",
    "x .@joe, y",
    "@",
    "@a",
    "a@",
    "$AAPL ^GSPC #tag @user",
    "$ AAPL # tag",
    "http://",
    "a://b",
    "1://b ://",
    "\u65e5\u672c\u8a9e\u306e\u30c6\u30ad\u30b9\u30c8\u3001\u3067\u3059\u3002",
    "\u00dcn\u00efc\u00f6d\u00e9 \u01c4 \u01c5 \u01c6 \u00bd \u00b2 \u0663",
    "tab\tnew\nline\xa0nbsp\u3000ideo",
    "zero\u200bwidth\u00adsoft\ufeffbom",
    "\u24b6\u24d0 circled \u2160\u2170 roman",
    "lone \ud800 surrogate \udfff end",
    "\U0010ffff max \U000e0100 vs \U0003134f",
    "emoji\U0001F600\U0001F64Fend \U0001F1FA\U0001F1F8",
    "\u2028line\u2029para\u0085nel\u1680ogham\u205fmmsp",
    "a\u0301e\u0301 combining",
    "x" * 127,
    "y" * 128,
    "ab " * 85,            # 255 chars
    "ab " * 85 + "c",      # 256
    "ab " * 85 + "cd",     # 257
    "#" * 64,
    "a@b " * 70,
    "A" * 250,
    ("word " * 40 + "http://t.co/abc ") * 3,
]

EMPTY = ""  # Q1: reference raises IndexError; batch API returns zero characters / zero tokens

_ASCII_WORD = "abcdefghijklmnopqrstuvwxyz"
_POOLS = {
    "lower": _ASCII_WORD,
    "upper": _ASCII_WORD.upper(),
    "digit": "0123456789",
    "punct": ".,!?'\"():;-_/\\@#$^&*+=<>[]{}|~`%",
    "space": " \t\n\r\x0b\x0c\x1c\x1d\x1e\x1f\x85\xa0\u1680\u2000\u2005\u200a\u2028\u2029\u202f\u205f\u3000",
    "latin": "".join(chr(c) for c in range(0xC0, 0x250)),
    "cjk": "".join(chr(c) for c in list(range(0x4E00, 0x4E80)) + list(range(0x3040, 0x30FF)) + list(range(0xAC00, 0xAC40))),
    "emoji": "".join(chr(c) for c in range(0x1F600, 0x1F650)),
    "mbpunct": "\u3001\u3002\u300c\u300d\u2026\u2014\u2019\u201c\u201d\u00b7\u00a1\u00bf",
    "numeric": "\u00bd\u00b2\u0663\u2160\u2170\u3007\u0969\u2460",
    "odd": "\u200b\u00ad\ufeff\x00\x01\x7f\u0080\u009f\ud800\udfff\U0010ffff\U000e0100\U000e01ef\U0003134f\u01c5\u24b6\u24d0\u0345",
}
_SPECIALS = ["@", "#", "$", "^", ":", "/", ".", "://", ".@", " .@", " #", " @", "@@", "//"]


def fuzz_string(rng: np.random.Generator, max_len: int = 60, profile: str = "mixed") -> str:
    n = int(rng.integers(0, max_len + 1))
    if profile == "ascii":
        pools, weights = ["lower", "upper", "digit", "punct", "space"], [0.5, 0.1, 0.08, 0.12, 0.2]
    elif profile == "marks":
        pools, weights = ["lower", "upper", "digit", "punct", "space"], [0.45, 0.05, 0.05, 0.1, 0.1]
    else:
        pools = list(_POOLS)
        weights = [0.25, 0.06, 0.05, 0.1, 0.14, 0.08, 0.08, 0.05, 0.05, 0.04, 0.10]
    weights = np.array(weights) / np.sum(weights)
    out = []
    while len(out) < n:
        if profile != "ascii" and rng.random() < (0.25 if profile == "marks" else 0.08):
            out.extend(_SPECIALS[int(rng.integers(len(_SPECIALS)))])
            continue
        pool = _POOLS[pools[int(rng.choice(len(pools), p=weights))]]
        run = int(rng.integers(1, 5))
        for _ in range(run):
            out.append(pool[int(rng.integers(len(pool)))])
    return "".join(out[:n])


def fuzz_strings(seed: int, count: int, max_len: int = 60, profile: str = "mixed"):
    rng = np.random.default_rng(seed)
    return [fuzz_string(rng, max_len, profile) for _ in range(count)]
