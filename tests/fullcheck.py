"""Exact parity of a WHOLE batch against the CPU oracle: the batch is cut into runs of whole strings, a pool of
processes runs the oracle's C batch path (oracle.tokenize_batch_utf8) on each run and compares it with the matching
slices of the GPU result -- every string, every output array.  Test infrastructure (uses oracle/)."""
from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np

_G = {}


def _work(ab):
    from oracle import oracle
    a, b = ab
    buf, off, r, feats = _G["buf"], _G["off"], _G["r"], _G["feats"]
    sub = np.ascontiguousarray(buf[off[a]:off[b]])
    so = (off[a:b + 1] - off[a]).astype(np.int64)
    o = oracle.tokenize_batch_utf8(sub, so, feats=feats)
    c0, c1 = int(r.char_offsets[a]), int(r.char_offsets[b])
    t0, t1 = int(r.tok_offsets[a]), int(r.tok_offsets[b])
    bad = []
    if c1 - c0 != o["n_chars"] or not np.array_equal(r.char_offsets[a:b + 1] - c0, o["char_offsets"]):
        bad.append("char_offsets")
    elif not np.array_equal(r.splits[c0:c1], o["splits"]):
        bad.append("splits")
    if t1 - t0 != o["n_tokens"] or not np.array_equal(r.tok_offsets[a:b + 1] - t0, o["tok_offsets"]):
        bad.append("tok_offsets")
    else:
        if not np.array_equal(r.spans[t0:t1], o["spans"]):
            bad.append("spans")
        if feats and not np.array_equal(r.tok_feats[t0:t1], o["tok_feats"]):
            bad.append("tok_feats")
    return (a, b, bad) if bad else None


def compare_all(buf, off, r, feats=False, target_bytes=4 << 20, procs=None):
    """Returns the number of strings compared; raises AssertionError naming the first differing run of strings."""
    S = len(off) - 1
    cuts = [0]
    while cuts[-1] < S:
        j = int(np.searchsorted(off, off[cuts[-1]] + target_bytes, side="left"))
        cuts.append(min(max(j, cuts[-1] + 1), S))
    tasks = list(zip(cuts[:-1], cuts[1:]))
    _G.update(buf=buf, off=off, r=r, feats=feats)
    procs = procs or min(os.cpu_count() or 1, 32)
    try:
        with mp.get_context("fork").Pool(procs) as pool:        # (children only touch NumPy and the oracle library)
            res = [x for x in pool.imap_unordered(_work, tasks, chunksize=1) if x]
    finally:
        _G.clear()
    assert not res, f"{len(res)} of {len(tasks)} runs of strings differ from the oracle; first: strings {min(res)[:2]}: {min(res)[2]}"
    return S
