#!/usr/bin/env python3
"""Quick device-side timing of the tokenize kernel on a synthetic workload (development aid).
   python tools/prof_run.py [tweets|mixed|docs] [n_strings] [reps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(1, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import synth
from latok_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "tweets"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
what = int(sys.argv[4]) if len(sys.argv) > 4 else 3
buf, off = {"tweets": synth.tweets, "mixed": synth.mixed_unicode}.get(wl, lambda k: synth.long_docs(k, 65536))(n)
best = 1e9
with Engine(0) as e:
    for i in range(reps):
        e.submit(buf, off, what)
        c, t = e.sizes()
        r = e.fetch() if i == 0 else None
        import ctypes as C
        from latok_b200 import _lib
        ms = C.c_float(0); w = C.c_int64(0)
        _lib.check(_lib.load().latok_b200_last_stats(e._h, C.byref(ms), C.byref(w)))
        if i == 0 and what & 2:
            e.token_bytes(); e.token_bytes()        # (the first call pays for loading the kernels)
            print(f"token byte ranges: {e.token_bytes_ms():.3f} ms ({(len(buf) + 24 * t + 40 * len(off)) / e.token_bytes_ms() / 1e6:.1f} GB/s algorithmic)")
        alg = len(buf) + c + 8 * t + 16 * (len(off))
        best = min(best, ms.value)
        print(f"{wl} S={len(off)-1} B={len(buf)} C={c} T={t} kernel={ms.value:.3f} ms  in={len(buf)/ms.value/1e6:.1f} GB/s  alg={alg/ms.value/1e6:.1f} GB/s ({alg/ms.value/1e6/6545.9*100:.1f}% of 6545.9) walks={w.value}")
print(f"{wl} best of {reps}: {best:.4f} ms")
