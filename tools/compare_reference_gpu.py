"""Side-by-side timing on the shipped corpora: the reference's own CUDA approaches
(fastgpu.cu / fastgpuOpt1.cu, UNMODIFIED, recompiled for sm_100a into
oracle/_ref/libref_cuda.so by `make -C oracle refcuda`) against b200Approach, with
the reference's protocol (framework/decodeUtil.c:30-70: zeroed output, first run
byte-checked, min wall time over 1 + REPEATS runs; the whole call is timed,
including the reference's per-call cudaMalloc / copies / cudaFree).

    python tools/compare_reference_gpu.py [repeats]      # on a B200
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O            # noqa: E402
import huffmandecoderongpus_b200 as hb  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    so = os.path.join(O.REF_DIR, "libref_cuda.so")
    if not os.path.exists(so):
        sys.exit("oracle/_ref/libref_cuda.so missing: run `make -C oracle refcuda` in the build container")
    ref = C.CDLL(so)
    fns = {}
    for name in ("fastgpuApproach", "fastgpuApproachOpt1"):
        f = getattr(ref, name)
        f.restype = None
        f.argtypes = [C.POINTER(O.RefCompressed), C.POINTER(O.RefUnCompressed), C.c_void_p]
        fns[name] = f
    ours = hb.lib().b200Approach

    def run(fn, cd_t, ucd_t, st, want):
        best = None
        for r in range(1 + reps):
            out = np.zeros(st.usize + 16, dtype=np.uint8)
            cd = cd_t(st.bits, st.nodes, st.usize, st.tree.ctypes.data, st.data.ctypes.data)
            ucd = ucd_t(st.usize, out.ctypes.data)
            t0 = time.perf_counter()
            fn(C.byref(cd), C.byref(ucd), None)
            dt = time.perf_counter() - t0
            if r == 0 and not np.array_equal(out[: st.usize], want):
                return None
            best = dt if best is None else min(best, dt)
        return best

    print(f"{'dataset':>9} {'bytes out':>10} | {'fastgpu ms':>11} {'fastgpuOpt1 ms':>15} {'b200 ms':>9} | speed-up vs best reference GPU path")
    for name in ("hello", "paper1", "news", "book2", "world192", "bible", "kjv", "ecoli"):
        p = O.corpus_path(name)
        if p is None:
            continue
        st = O.load_huff(p)
        want = O.simple_decode(st)
        t = {}
        for k, f in fns.items():
            if k == "fastgpuApproachOpt1" and name == "hello":
                t[k] = None     # its 200x400-bit launch shape needs more than one block's worth of input
                continue
            t[k] = run(f, O.RefCompressed, O.RefUnCompressed, st, want)
        tb = run(ours, hb.RefCompressedData, hb.RefUnCompressedData, st, want)
        fm = lambda x: "   wrong/na" if x is None else f"{x * 1e3:11.3f}"
        refbest = min(v for v in t.values() if v is not None) if any(v is not None for v in t.values()) else None
        sp = f"{refbest / tb:8.1f}x" if refbest and tb else "n/a"
        print(f"{name:>9} {st.usize:>10} | {fm(t['fastgpuApproach'])} {fm(t['fastgpuApproachOpt1']):>15} {fm(tb):>9} | {sp}")


if __name__ == "__main__":
    main()
