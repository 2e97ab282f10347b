"""Prefix/scaling sweep (the reference's graphtest, framework/mainrun.c:361-410, test
names graph1..4 / quickgraph1..3): decode prefixes of growing length that end on a
codeword boundary with b200Approach (whole approach call, host buffers) and with
the CPU paths, min of 1 + repeats wall-clock runs, first run byte-checked.  Shows
where the GPU path overtakes the reference's CPU decoders.

    python tools/graph_sweep.py [corpus=kjv] [points=24] [repeats=5]      # on a B200
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O                  # noqa: E402
import huffmandecoderongpus_b200 as hb  # noqa: E402


def best(fn, want, reps):
    b = None
    for r in range(1 + reps):
        t0 = time.perf_counter()
        got = fn()
        dt = time.perf_counter() - t0
        if r == 0:
            assert np.array_equal(got, want)
        b = dt if b is None else min(b, dt)
    return b


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "kjv"
    points = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    st = O.load_huff(O.corpus_path(name))
    use_ref = O.ref() is not None
    print(f"# {name}: {st.bits} bits, {st.usize} symbols; CPU paths = "
          f"{'unmodified reference' if use_ref else 'oracle port'}, 1 thread; times in ms")
    print(f"{'bits':>10} {'symbols':>9} | {'b200 wall':>10} {'b200 dev':>9} | {'simpleDecode':>12} {'jumptable8':>10} | b200 vs best CPU")
    targets = sorted(set(int(st.bits * (i / points) ** 2) for i in range(1, points + 1)) | {256, 2048, 16384})
    for target in targets:
        bits, usize = O.prefix_sizes(st, max(target, 8))
        if usize == 0:
            continue
        want = O.simple_decode(st, bits=bits)
        tb = best(lambda: hb.b200_approach(st.tree, st.data, bits, usize), want, reps)
        dev = hb.lib().b200ApproachLastDeviceMs()
        if use_ref:
            ts = best(lambda: O.ref_decode(st, "simpleDecode", bits=bits, usize=usize), want, min(reps, 2))
            tj = best(lambda: O.ref_decode(st, "jumptableApproach", 8, bits=bits, usize=usize), want, min(reps, 2))
        else:
            ts = best(lambda: O.simple_decode(st, bits=bits), want, min(reps, 2))
            tj = best(lambda: O.jumptable_decode(st, 8, bits=bits), want, min(reps, 2))
        print(f"{bits:>10} {usize:>9} | {tb * 1e3:10.4f} {dev:9.4f} | {ts * 1e3:12.4f} {tj * 1e3:10.4f} | {min(ts, tj) / tb:8.2f}x")


if __name__ == "__main__":
    main()
