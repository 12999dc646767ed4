"""Seeded synthetic corpora of the BASELINE.json configs, generated directly
as packed (uint8 UTF-8 buffer, int64 offsets) with vectorised NumPy -- no per-string Python loop.

    tweets(n, seed=20240601)         config #2: ~140-char ASCII-heavy tweets
    long_docs(n, doc_bytes, seed)    config #3: long documents, newline every ~80 chars, some with
                                     >= 32 KB space-free runs and chunks holding several marks
    mixed_unicode(n, seed)           config #4: accented Latin / CJK / emoji / multi-byte punctuation
    scaling_batch(n_chars, seed)     config #5: config-#2 text up to a character budget

Host-side test/bench data only; nothing here is on the tokenization path.
"""
from __future__ import annotations

import numpy as np

_LOWER = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)
_UPPER = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)
_ALNUM = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
_DIGIT = np.frombuffer(b"0123456789", dtype=np.uint8)
_PUNCT = np.frombuffer(b".,!?'\"():;-", dtype=np.uint8)


class _Pool:
    """A list of byte strings stored as one flat buffer + offsets."""

    def __init__(self):
        self.chunks, self.lens = [], []

    def add(self, items):
        base = len(self.lens)
        for b in items:
            self.chunks.append(b)
            self.lens.append(len(b))
        return np.arange(base, len(self.lens))

    def freeze(self):
        self.data = np.frombuffer(b"".join(self.chunks), dtype=np.uint8)
        self.len = np.asarray(self.lens, dtype=np.int64)
        self.off = np.zeros(len(self.lens), dtype=np.int64)
        np.cumsum(self.len[:-1], out=self.off[1:])
        return self


def _rand_words(rng, n, alphabet, lo, hi):
    lens = rng.integers(lo, hi + 1, size=n)
    flat = alphabet[rng.integers(0, len(alphabet), size=int(lens.sum()))].tobytes()
    out, p = [], 0
    for ln in lens:
        out.append(flat[p:p + ln])
        p += ln
    return out


def _ragged_gather(pool: _Pool, ids: np.ndarray, sep: np.ndarray):
    """Concatenate pool items `ids`, each preceded by sep[i] space characters.  Returns (bytes, end offset
    of every item in the output)."""
    lens = pool.len[ids] + sep
    ends = np.cumsum(lens)
    total = int(ends[-1]) if len(ends) else 0
    out = np.full(total, 0x20, dtype=np.uint8)
    starts = ends - pool.len[ids]                       # where the item's own bytes begin
    idx = np.repeat(pool.off[ids] - starts, pool.len[ids]) + _arange_runs(starts, pool.len[ids])
    pos = np.repeat(starts, pool.len[ids]) + _arange_within(pool.len[ids])
    out[pos] = pool.data[idx]
    return out, ends


def _arange_within(lens):
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    first = np.cumsum(lens) - lens
    return np.arange(total, dtype=np.int64) - np.repeat(first, lens)


def _arange_runs(starts, lens):
    return np.repeat(starts, lens) + _arange_within(lens)


def _tweet_pool(rng):
    pool = _Pool()
    words = _rand_words(rng, 20000, _LOWER, 1, 12)
    ids = {"word": pool.add(words)}
    caps = [w.capitalize() for w in words[:3000]]
    camel = [a + b.capitalize() for a, b in zip(words[100:2100], words[2100:4100])]
    allcaps = [w.upper() for w in words[50:2050]]
    ids["case"] = pool.add(caps + camel + allcaps)
    ids["num"] = pool.add([str(int(x)).encode() for x in rng.integers(0, 100000, size=5000)])
    ids["emoji"] = pool.add([chr(c).encode("utf-8") for c in range(0x1F600, 0x1F650)])
    users = _rand_words(rng, 8000, _ALNUM, 3, 12)
    ids["mention"] = pool.add([b"@" + u for u in users[:4000]])
    ids["tag"] = pool.add([b"#" + w for w in words[4000:8000] if len(w) > 1])
    ids["tick"] = pool.add([b"$" + w for w in _rand_words(rng, 500, _UPPER, 2, 5)])
    ids["url"] = pool.add([b"https://t.co/" + u for u in _rand_words(rng, 20000, _ALNUM, 10, 10)])
    ids["email"] = pool.add([a + b"@" + b + b".com" for a, b in zip(words[8000:10000], words[10000:12000])])
    ids["dotat"] = pool.add([b".@" + u for u in users[4000:6000]])
    ids["punct"] = pool.add([bytes([c]) for c in _PUNCT])
    ids["nl"] = pool.add([b"\n"])
    return pool.freeze(), ids


def _zipf_choice(rng, ids, size, a=1.3):
    p = 1.0 / np.arange(1, len(ids) + 1) ** a
    cdf = np.cumsum(p / p.sum())
    return ids[np.minimum(np.searchsorted(cdf, rng.random(size)), len(ids) - 1)]


def _token_stream(rng, pool, ids, n_tok, emoji_frac=0.02):
    """Token ids + number of spaces before each token (0 for attached punctuation)."""
    kind = rng.random(n_tok)
    tok = _zipf_choice(rng, ids["word"], n_tok)
    m = kind < 0.06
    tok[m] = ids["case"][rng.integers(0, len(ids["case"]), size=int(m.sum()))]
    m = (kind >= 0.06) & (kind < 0.09)
    tok[m] = ids["num"][rng.integers(0, len(ids["num"]), size=int(m.sum()))]
    m = (kind >= 0.09) & (kind < 0.09 + emoji_frac)
    tok[m] = ids["emoji"][rng.integers(0, len(ids["emoji"]), size=int(m.sum()))]
    sep = np.ones(n_tok, dtype=np.int64)
    sep[rng.random(n_tok) < 0.03] = 2
    # punctuation attached to ~15 % of the words: an extra zero-separator token right after them
    attach = np.nonzero(rng.random(n_tok) < 0.15)[0]
    punct = ids["punct"][rng.integers(0, len(ids["punct"]), size=len(attach))]
    tok = np.insert(tok, attach + 1, punct)
    sep = np.insert(sep, attach + 1, 0)
    return tok, sep


def _cut(tok_ends_chars, targets):
    """Token index boundaries so that string i holds about targets[i] characters."""
    return np.searchsorted(tok_ends_chars, np.cumsum(targets), side="left") + 1


def _char_len(pool, ids_map):
    # characters per pool item (emoji are 4 bytes / 1 char; everything else here is ASCII)
    clen = pool.len.copy()
    clen[ids_map["emoji"]] = 1
    return clen


def _assemble(rng, pool, ids, n_strings, targets, specials, newline_every=0, emoji_frac=0.02, chunk=200000):
    """Build strings chunk by chunk.  `specials`: list of (pool-id array, per-string probability)."""
    bufs, lens = [], []
    clen = _char_len(pool, ids)
    for c0 in range(0, n_strings, chunk):
        tg = targets[c0:c0 + chunk]
        n = len(tg)
        n_tok = int(tg.sum() / 5.2) + 64
        tok, sep = _token_stream(rng, pool, ids, n_tok, emoji_frac)
        ends_c = np.cumsum(clen[tok] + sep)
        while ends_c[-1] < tg.sum() + 16:
            t2, s2 = _token_stream(rng, pool, ids, n_tok // 4 + 64, emoji_frac)
            tok, sep = np.concatenate([tok, t2]), np.concatenate([sep, s2])
            ends_c = np.cumsum(clen[tok] + sep)
        bnd = np.minimum(_cut(ends_c, tg), len(tok))
        first = np.concatenate([[0], bnd[:-1]])
        ntok = np.maximum(bnd - first, 1)
        # per-string specials: overwrite one random token of the string
        for pid, prob in specials:
            hit = np.nonzero(rng.random(n) < prob)[0]
            pos = first[hit] + (rng.random(len(hit)) * ntok[hit]).astype(np.int64)
            pos = np.minimum(pos, len(tok) - 1)
            tok[pos] = pid[rng.integers(0, len(pid), size=len(hit))]
            sep[pos] = np.maximum(sep[pos], 1)
        if newline_every:
            nl = np.nonzero(rng.random(len(tok)) < 1.0 / max(newline_every / 6.2, 1.0))[0]
            tok = np.insert(tok, nl, ids["nl"][0])
            sep = np.insert(sep, nl, 0)
            bnd = bnd + np.searchsorted(nl, bnd, side="left")
            first = np.concatenate([[0], bnd[:-1]])
        sep[first[first < len(sep)]] = 0            # no leading space
        used = int(bnd[-1])
        data, ends_b = _ragged_gather(pool, tok[:used], sep[:used])
        str_end = ends_b[np.minimum(bnd, used) - 1]
        str_len = np.diff(np.concatenate([[0], str_end]))
        bufs.append(data[:int(str_end[-1])])
        lens.append(str_len)
    lens = np.concatenate(lens)
    offsets = np.zeros(n_strings + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return np.concatenate(bufs), offsets


def _tweet_specials(ids):
    return [(ids["mention"], 0.4), (ids["tag"], 0.3), (ids["tick"], 0.03), (ids["url"], 0.25),
            (ids["email"], 0.02), (ids["dotat"], 0.01)]


def tweets(n_strings: int = 1_000_000, seed: int = 20240601):
    """Config #2: tweet-sized ASCII-heavy strings, length ~ clip(N(140, 35), 20, 280) characters."""
    rng = np.random.default_rng(seed)
    pool, ids = _tweet_pool(rng)
    targets = np.clip(rng.normal(140, 35, size=n_strings), 20, 280).astype(np.int64)
    return _assemble(rng, pool, ids, n_strings, targets, _tweet_specials(ids))


def scaling_batch(n_chars: int = 1_000_000_000, seed: int = 20240605):
    """Config #5: config-#2 text with about n_chars characters in total."""
    return tweets(max(int(n_chars / 140.5), 1), seed)


def long_docs(n_docs: int = 100_000, doc_bytes: int = 65536, seed: int = 20240602):
    """Config #3: documents of ~doc_bytes each with a newline every ~80 characters.  About 1 % of the
    documents get a space-free run of >= doc_bytes/2 (>= 32 KB at the default size) and about 2 % get
    chunks holding three or more marks (the backlog of latok.c:225-238)."""
    rng = np.random.default_rng(seed)
    pool, ids = _tweet_pool(rng)
    targets = np.full(n_docs, doc_bytes, dtype=np.int64)
    per_doc = max(doc_bytes // 140, 1)
    specials = [(p, min(1.0, pr)) for p, pr in _tweet_specials(ids)]
    buf, offsets = _assemble(rng, pool, ids, n_docs, targets, specials, newline_every=80,
                             chunk=max(1, (1 << 28) // doc_bytes))
    buf = buf.copy()
    # sprinkle more marks: tweets carry ~1 special per string, documents should carry ~1 per 140 chars
    n_extra = int(len(buf) / 140 * 0.9)
    pos = rng.integers(1, max(len(buf) - 24, 2), size=n_extra)
    for pid, frac in ((ids["mention"], 0.45), (ids["tag"], 0.3), (ids["email"], 0.05)):
        sel = pos[rng.random(len(pos)) < frac]
        item = pid[rng.integers(0, len(pid), size=len(sel))]
        ln = np.minimum(pool.len[item], 20)
        # overwrite in place: " " + item (keeps document sizes and UTF-8 validity only where the target is ASCII)
        ok = np.ones(len(sel), dtype=bool)
        for k in range(21):
            ok &= buf[np.minimum(sel + k, len(buf) - 1)] < 0x80
        sel, item, ln = sel[ok], item[ok], ln[ok]
        buf[sel] = 0x20
        idx = np.repeat(pool.off[item], ln) + _arange_within(ln)
        buf[np.repeat(sel + 1, ln) + _arange_within(ln)] = pool.data[idx]
    starts = offsets[:-1]
    size = np.diff(offsets)
    # space-free runs: replace whitespace by ',' over the second half of ~1 % of the documents
    for d in np.nonzero(rng.random(n_docs) < 0.01)[0]:
        a = int(starts[d] + size[d] // 4)
        b = int(starts[d] + size[d] // 4 + max(size[d] // 2, 1))
        seg = buf[a:b]
        seg[(seg == 0x20) | (seg == 0x0A)] = 0x2C
    # backlog: turn the separators of a short stretch into ',' so several marks share one chunk
    for d in np.nonzero(rng.random(n_docs) < 0.02)[0]:
        a = int(starts[d] + rng.integers(0, max(size[d] - 400, 1)))
        seg = buf[a:a + 300]
        seg[(seg == 0x20) | (seg == 0x0A)] = 0x2C
    return buf, offsets


def _mixed_pool(rng):
    pool = _Pool()
    ids = {}

    def cps_words(n, ranges, lo, hi):
        cps = np.concatenate([np.arange(a, b) for a, b in ranges])
        lens = rng.integers(lo, hi + 1, size=n)
        flat = cps[rng.integers(0, len(cps), size=int(lens.sum()))]
        out, p = [], 0
        for ln in lens:
            out.append("".join(map(chr, flat[p:p + ln])).encode("utf-8"))
            p += ln
        return out
    ids["latin"] = pool.add(cps_words(6000, [(0xC0, 0xD7), (0xD8, 0xF7), (0xF8, 0x250), (0x61, 0x7B)], 2, 10))
    ids["cjk"] = pool.add(cps_words(6000, [(0x4E00, 0x9FEF), (0x3041, 0x3097), (0x30A1, 0x30FB), (0xAC00, 0xD7A4)], 1, 6))
    ids["emoji"] = pool.add([chr(c).encode("utf-8") for c in list(range(0x1F600, 0x1F650)) + list(range(0x1F300, 0x1F340))
                             + [0x2764, 0x2728, 0x263A]])
    ids["mbpunct"] = pool.add([s.encode("utf-8") for s in "、 。 「 」 … — ’ “ ” 　    ".split(" ") if s]
                              + ["　".encode(), " ".encode(), " ".encode()])
    ascii_words = _rand_words(rng, 4000, _LOWER, 1, 10)
    users = _rand_words(rng, 1000, _ALNUM, 3, 10)
    ids["ascii"] = pool.add(ascii_words + [w.capitalize() for w in ascii_words[:500]]
                            + [b"@" + u for u in users[:400]] + [b"#" + w for w in ascii_words[500:900]]
                            + [b"https://t.co/" + u for u in _rand_words(rng, 500, _ALNUM, 10, 10)]
                            + [a + b"@" + b + b".org" for a, b in zip(ascii_words[1000:1200], ascii_words[1200:1400])]
                            + [bytes([c]) for c in _PUNCT] + [str(int(x)).encode() for x in rng.integers(0, 9999, size=300)])
    return pool.freeze(), ids


def mixed_unicode(n_strings: int = 1_000_000, seed: int = 20240603, mean_chars: int = 160):
    """Config #4: ~160-character strings mixing accented Latin (35 %), CJK (30 %), emoji (10 %),
    multi-byte punctuation / spaces (5 %) and ASCII with marks (20 %)."""
    rng = np.random.default_rng(seed)
    pool, ids = _mixed_pool(rng)
    # characters per pool item
    clen = np.array([len(bytes(pool.data[o:o + l]).decode("utf-8")) for o, l in zip(pool.off, pool.len)], dtype=np.int64)
    targets = np.clip(rng.normal(mean_chars, 40, size=n_strings), 10, 400).astype(np.int64)
    bufs, lens = [], []
    groups = [("latin", 0.35), ("cjk", 0.30), ("emoji", 0.10), ("mbpunct", 0.05), ("ascii", 0.20)]
    cdf = np.cumsum([g[1] for g in groups])
    chunk = 200000
    for c0 in range(0, n_strings, chunk):
        tg = targets[c0:c0 + chunk]
        n_tok = int(tg.sum() / 4.0) + 64
        while True:
            kind = np.searchsorted(cdf, rng.random(n_tok))
            tok = np.empty(n_tok, dtype=np.int64)
            for g, (name, _) in enumerate(groups):
                m = kind == g
                tok[m] = ids[name][rng.integers(0, len(ids[name]), size=int(m.sum()))]
            # CJK text and emoji mostly run together without spaces
            sep = np.ones(n_tok, dtype=np.int64)
            sep[((kind == 1) | (kind == 2) | (kind == 3)) & (rng.random(n_tok) < 0.7)] = 0
            ends_c = np.cumsum(clen[tok] + sep)
            if ends_c[-1] >= tg.sum() + 16:
                break
            n_tok = int(n_tok * 1.3)
        bnd = np.minimum(_cut(ends_c, tg), len(tok))
        first = np.concatenate([[0], bnd[:-1]])
        sep[first[first < len(sep)]] = 0
        used = int(bnd[-1])
        data, ends_b = _ragged_gather(pool, tok[:used], sep[:used])
        str_end = ends_b[np.minimum(bnd, used) - 1]
        bufs.append(data[:int(str_end[-1])])
        lens.append(np.diff(np.concatenate([[0], str_end])))
    lens = np.concatenate(lens)
    offsets = np.zeros(n_strings + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return np.concatenate(bufs), offsets


def to_strings(buf: np.ndarray, offsets: np.ndarray, lo: int = 0, hi: int | None = None):
    """Decode strings [lo, hi) of a packed corpus back to Python str (for oracle checks on samples)."""
    hi = len(offsets) - 1 if hi is None else hi
    raw = buf.tobytes() if isinstance(buf, np.ndarray) else bytes(buf)
    return [raw[offsets[i]:offsets[i + 1]].decode("utf-8", "surrogatepass") for i in range(lo, hi)]
