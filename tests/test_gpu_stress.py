"""Randomised parity on the GPU: random prefix codes (random complete trees, codewords of up
to 32 bits), random symbol streams of random length drawn from random distributions, every
kernel path, words per thread and output alignment, against the CPU oracle.  Seeded; the
number of cases grows with HB_STRESS_CASES (default sized for a few seconds)."""
import os

import numpy as np
import pytest

import oracle_lib as O
import huffmandecoderongpus_b200 as hb

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("seed", range(int(os.environ.get("HB_STRESS_SEEDS", "6"))))
def test_random_codes_and_streams(seed):
    rng = np.random.default_rng(1000 + seed)
    dev = torch.device("cuda:0")
    ncases = int(os.environ.get("HB_STRESS_CASES", "10"))
    for case in range(ncases):
        nleaves = int(rng.choice([2, 3, 5, 17, 64, 200, 256]))
        maxlen = int(rng.choice([4, 9, 13, 20, 32]))
        lengths = O.random_lengths(rng, nleaves, maxlen)
        tree, codes = O.tree_from_lengths(lengths)
        # symbol distribution: uniform, skewed towards short codes, or towards long ones
        w = np.array([2.0 ** (-l) for l in lengths])
        mode = int(rng.integers(3))
        p = np.ones(len(lengths)) if mode == 0 else (w if mode == 1 else 1.0 / w)
        p = p / p.sum()
        n = int(rng.choice([1, 7, 300, 5000, 70000, 400000, 1500000]))
        syms = rng.choice(len(lengths), size=n, p=p)
        data, bits = O.encode_with_codes(codes, syms)
        want = (syms & 255).astype(np.uint8)
        f = hb.HuffFile(tree, data, bits, n)
        wpt = int(rng.choice([4, 8, 16]))
        ctx = hb.Context(0, stream=torch.cuda.current_stream().cuda_stream, words_per_thread=wpt)
        ctx.set_sync_path(str(rng.choice(["auto", "fsm", "probe"])))
        ctx.set_sync_copies(int(rng.choice([-1, 0, 1, 2])))
        ctx.set_emit_lane_subsequences(int(rng.choice([1, 2])))
        ctx.set_emit_path(str(rng.choice(["auto", "bytes", "flat", "words", "words32", "words32", "words32w", "words32w", "words64w"])))
        ctx.set_emit_table(int(rng.choice([0, 9, 10, 11, 12, 13, 14, 15])), int(rng.choice([-1, 3, 2, 0])))
        cb = hb.Codebook(ctx, tree)
        nb = (bits + 7) // 8
        comp = torch.zeros((nb + 15) // 16 * 16 + 32, dtype=torch.uint8, device=dev)
        comp[:nb] = torch.from_numpy(np.ascontiguousarray(data[:nb])).to(dev)
        off = int(rng.integers(16))
        raw = torch.full((n + 64 + off,), 0xA5, dtype=torch.uint8, device=dev)
        res = hb.decode_device(ctx, cb, comp.data_ptr(), comp.numel(), bits, raw[off:].data_ptr(), n)
        got = raw.cpu().numpy()
        tag = (seed, case, nleaves, maxlen, n, wpt, off)
        assert res["n_symbols"] == n, tag
        assert np.array_equal(got[off: off + n], want), tag
        assert (got[:off] == 0xA5).all() and (got[off + n:] == 0xA5).all(), tag   # nothing outside the slice
        cb.close()
        ctx.close()
