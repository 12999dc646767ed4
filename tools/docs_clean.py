import sys; sys.path.insert(0,'.'); sys.path.insert(1,'tests')
import numpy as np, ctypes as C
from latok_b200 import _lib
import synth
from latok_b200.engine import Engine
buf, off = synth.long_docs(3000, 65536)
keep = [i for i in range(3000) if (buf[off[i]:off[i+1]] == 0x2C).mean() < 0.02]
print("kept", len(keep), "of 3000 documents")
parts = [buf[off[i]:off[i+1]] for i in keep]
nb = np.concatenate(parts); no = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
with Engine(0) as e:
    for i in range(3):
        e.submit(nb, no, 3); c, t = e.sizes()
        ms = C.c_float(0); w = C.c_int64(0)
        _lib.check(_lib.load().latok_b200_last_stats(e._h, C.byref(ms), C.byref(w)))
        alg = len(nb) + c + 8 * t + 16 * len(no)
        print(f"clean docs B={len(nb)} kernel={ms.value:.3f} ms in={len(nb)/ms.value/1e6:.1f} GB/s alg={alg/ms.value/1e6:.1f} GB/s walks={w.value}")
