#!/usr/bin/env python3
"""Builds the pair (CUDA library, CPU oracle) over the class ranges of the UCD that ships with the running Python
(15.0 on CPython 3.12) into latok_b200/_variants/ (git-ignored, travels to the GPU box), for
tests/test_gpu_extended.py::test_ucd15_library (which runs tools/ucd_check.py against the pair).

    python tools/build_ucd_variant.py

The default build (UCD 11, bit-exact with the reference) is untouched: the generated tables are put back afterwards.
"""
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
VAR = ROOT / "latok_b200" / "_variants"
LIB, ORACLE = VAR / "liblatok_ucd.so", VAR / "liblatok_oracle_ucd.so"


def main():
    VAR.mkdir(exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        ranges = tmp / "ucd.txt"
        subprocess.run([sys.executable, str(ROOT / "tools" / "regen_classes.py"), "--out", str(ranges)], check=True,
                       stdout=subprocess.DEVNULL)
        env = dict(os.environ, LATOK_CLASSES=str(ranges), LATOK_LOW_LIMIT="0x32400")
        try:
            subprocess.run([sys.executable, "-m", "latok_b200.build"], check=True, cwd=ROOT, stdout=subprocess.DEVNULL,
                           env=dict(env, LATOK_B200_LIB_OUT=str(LIB)))
            gen = tmp / "gen"
            subprocess.run([sys.executable, str(ROOT / "tools" / "gen_tables.py"), str(gen)], check=True, env=env,
                           stdout=subprocess.DEVNULL)
            orc = tmp / "orc"
            (orc / "_gen").mkdir(parents=True)
            shutil.copy(ROOT / "oracle" / "latok_oracle.c", orc / "latok_oracle.c")
            shutil.copy(gen / "oracle_runs.h", orc / "_gen" / "oracle_runs.h")
            subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", str(orc / "latok_oracle.c"), "-o", str(ORACLE)], check=True)
        finally:
            # back to the UCD-11 tables the default library and oracle are built from
            subprocess.run([sys.executable, str(ROOT / "tools" / "gen_tables.py")], check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
    print(LIB)
    print(ORACLE)


if __name__ == "__main__":
    main()
