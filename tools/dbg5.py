"""Developer aid: run the split-mask + span path (v5 kernel) against the oracle on a ladder of cases and print the
first difference with context.  GPU box only:  python tools/dbg5.py [level]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import corpus
from oracle import oracle
from latok_b200.engine import Engine


def check(e, texts, label, what=3):
    t0 = time.time()
    try:
        r = e.run(texts, what)
    except Exception as ex:
        print(f"[{label}] EXCEPTION {type(ex).__name__}: {ex}")
        return False
    o = oracle.tokenize_batch(texts)
    ok = True
    if r.n_chars != o["n_chars"]:
        print(f"[{label}] n_chars {r.n_chars} vs {o['n_chars']}"); ok = False
    if ok and not np.array_equal(r.char_offsets, o["char_offsets"]):
        bad = np.nonzero(r.char_offsets != o["char_offsets"])[0]
        print(f"[{label}] char_offsets differ at {bad[:8]}: {r.char_offsets[bad[:8]]} vs {o['char_offsets'][bad[:8]]}"); ok = False
    if ok:
        bad = np.nonzero(r.splits != o["splits"])[0]
        if bad.size:
            b = int(bad[0])
            print(f"[{label}] splits differ at {bad.size} chars, first {bad[:8]}; around: got {r.splits[max(0,b-8):b+8].tolist()} want {o['splits'][max(0,b-8):b+8].tolist()}")
            ok = False
    if r.n_tokens != o["n_tokens"]:
        print(f"[{label}] n_tokens {r.n_tokens} vs {o['n_tokens']}"); ok = False
    elif not np.array_equal(r.tok_offsets, o["tok_offsets"]):
        bad = np.nonzero(r.tok_offsets != o["tok_offsets"])[0]
        print(f"[{label}] tok_offsets differ at {bad[:8]}: {r.tok_offsets[bad[:8]]} vs {o['tok_offsets'][bad[:8]]}"); ok = False
    else:
        bad = np.nonzero((r.spans != o["spans"]).any(axis=1))[0]
        if bad.size:
            print(f"[{label}] spans differ at {bad.size} tokens, first {bad[:6]}: got {r.spans[bad[:4]].tolist()} want {o['spans'][bad[:4]].tolist()}"); ok = False
    print(f"[{label}] {'ok' if ok else 'FAIL'}  S={len(texts)} C={r.n_chars} T={r.n_tokens} kernel={r.kernel_ms:.3f} ms walks={r.lookahead_walks} ({time.time()-t0:.1f}s)")
    return ok


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    e = Engine(0)
    ok = True
    ok &= check(e, ["This is a #test! Testing, Testing, 1 2 3"], "one")
    ok &= check(e, corpus.FIXTURES[:8], "fix8")
    ok &= check(e, corpus.FIXTURES, "fixtures")
    for t in corpus.FIXTURES:
        if not check(e, [t], repr(t[:24])):
            ok = False
    ok &= check(e, [], "none"); ok &= check(e, [""], "empty"); ok &= check(e, ["", "a", "", "", "b c", ""], "ragged")
    if level >= 2:
        ok &= check(e, ["word " * 700], "3500 ascii")
        ok &= check(e, ["word " * 2000], "10k ascii")
        ok &= check(e, ["camelCaseWord HTTPServer " * 900], "22k camel")
        ok &= check(e, ["a@b.c,d@e.f fooBar, x-y z " * 3000], "marks 78k")
        ok &= check(e, ["日本語のテキスト、です。 café naïve " * 2000], "cjk")
        ok &= check(e, ["x" * 40000], "x40000")
        ok &= check(e, ["a@b" + "x" * 20000 + " tail"], "mark+x20000")
        ok &= check(e, ["a@b,c@d,e@f one,two three,four five,six " * 800], "backlog")
        for seed, count, ml, prof in [(1, 6000, 60, "mixed"), (2, 6000, 140, "ascii"), (3, 6000, 100, "marks"), (4, 400, 3000, "mixed")]:
            ok &= check(e, corpus.fuzz_strings(seed, count, ml, prof), f"fuzz {prof} {seed}")
    if level >= 3:
        import synth
        buf, offs = synth.tweets(200000, seed=7) if hasattr(synth, "tweets") else (None, None)
        if buf is not None:
            texts = [bytes(buf[offs[i]:offs[i + 1]]).decode("utf-8") for i in range(20000)]
            ok &= check(e, texts, "tweets20k")
    print("ALL OK" if ok else "SOME FAILED")
    e.close()


if __name__ == "__main__":
    main()
