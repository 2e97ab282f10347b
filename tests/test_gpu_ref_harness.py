"""The drop-in proof: the UNMODIFIED reference harness (framework/decodeUtil.c evaluate(),
framework/huffdata.c loaders and byte comparison, framework/timing.c), compiled where it lies
under /root/reference by `make -C oracle refharness` (build container only; the binary travels
in oracle/_ref/), with b200Approach and b200ApproachMulti registered the way
framework/mainrun.c:480-501 registers approaches.  evaluate() exits non-zero on the first
differing byte (framework/decodeUtil.c:47-52)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
FILES = os.path.join(ROOT, "oracle", "_ref", "files")


def test_unmodified_reference_harness_accepts_b200approach():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    p = subprocess.run([BIN, FILES], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l.split() for l in p.stdout.splitlines() if l.strip()]
    # "%17s %8s     %.9f ms": 5 corpora x 2 approaches, every first decode byte-checked by the reference
    assert [(l[0], l[1]) for l in lines] == [(a, n) for a in ("b200", "b200multi")
                                             for n in ("paper1", "hello", "news", "kjv", "book2")], p.stdout
    assert "problem with" not in p.stderr and "different" not in p.stdout
