"""Decodes for compute-sanitizer (memcheck / racecheck): every kernel path on streams that
exercise it -- shipped corpora (all end in a partial tile), a code with 20-bit codewords
(world192: the multi-level table and the long-codeword fallbacks), a fixed-length code (ecoli:
closed-form chains), a synthetic stream long enough for the transducer sync kernel and the
flat emit kernel, a two-shard decode.  Host buffers only (no torch).  Every result is
checked against the reference's digest / the generator."""
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import huffmandecoderongpus_b200 as hb  # noqa: E402
import oracle_lib as O  # noqa: E402

names = sys.argv[1:] or ["paper1", "kjv", "ecoli", "world192"]
for name in names:
    f = hb.HuffFile.load(O.corpus_path(name))
    for sync in ("auto", "fsm", "probe"):
        for emit in ("words", "words32", "bytes", "flat"):
            c = hb.Context(0)
            c.set_sync_path(sync)
            c.set_emit_path(emit)
            out = np.zeros(f.usize + 16, dtype=np.uint8)
            res = hb.decode_host(c, f.tree, f.data, f.bits, out[: f.usize])
            ok = res["n_symbols"] == f.usize and O.sha256(out[: f.usize]) == O.CORPORA[name][2]
            print(f"{name:9s} sync={sync:5s} emit={emit:5s} launches={res['launches']} {'ok' if ok else 'MISMATCH'}", flush=True)
            assert ok
            c.close()
# a synthetic stream of 2^22 symbols: full tiles through the transducer kernel, chunked host path
m = hb.Model(0)
fs, syms = m.huff_file_cpu(0x48554646, 1 << 22)
for emit in ("words", "words32", "flat"):
    c = hb.Context(0)
    c.set_sync_path("fsm")
    c.set_emit_path(emit)
    c.set_host_chunk(1 << 18)
    out = np.zeros(fs.usize, dtype=np.uint8)
    res = hb.decode_host(c, fs.tree, fs.data, fs.bits, out)
    assert res["n_symbols"] == fs.usize and np.array_equal(out, syms)
    print(f"english4m sync=fsm emit={emit} chunked host path launches={res['launches']} ok", flush=True)
    c.close()
# one process, every visible device (one shard each)
mm = hb.Multi(0)
out = np.zeros(fs.usize, dtype=np.uint8)
res = mm.decode_host(fs.tree, fs.data, fs.bits, out)
assert res["n_symbols"] == fs.usize and np.array_equal(out, syms)
print(f"english4m hb_multi_decode_host devices={res['n_devices']} ok", flush=True)
mm.close()
ctx = hb.Context(0)
f = hb.HuffFile.load(O.corpus_path("hello"))
out = np.zeros(f.usize, dtype=np.uint8)
hb.decode_onethread(ctx, f.tree, f.data, f.bits, out)
assert O.sha256(out) == O.CORPORA["hello"][2]
print("hello onethread ok", flush=True)
ctx.close()
