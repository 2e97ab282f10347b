"""Throughput on codes that synchronise badly:
 * "even", "even16", "mult4": every codeword length a multiple of 2 / 2 / 4 -- chains of entry
   offsets off that residue class never merge with the true one; since round 2 they are not
   followed at all (hb_stream_args.hstep), which removed the 18x cliff of round 1
 * "mult3": multiples of 3 -- no power-of-two factor, the full path
 * "7or8": 64 codes of 7 bits + 128 codes of 8 bits -- almost fixed length, still slow
 * "english": the bench model, for scale
Builds the stream on the CPU (small), decodes on the GPU, checks bytes."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O                  # noqa: E402
import huffmandecoderongpus_b200 as hb  # noqa: E402


tree_from_lengths, encode = O.tree_from_lengths, O.encode_with_codes


def run(name, tree, data, bits, syms, ctx, dev):
    cb = hb.Codebook(ctx, tree)
    nb = (bits + 7) // 8
    comp = torch.zeros((nb + 15) // 16 * 16 + 32, dtype=torch.uint8, device=dev)
    comp[:nb] = torch.from_numpy(data[:nb]).to(dev)
    out = torch.zeros(syms.size + 64, dtype=torch.uint8, device=dev)
    best = None
    for _ in range(4):
        res = hb.decode_device(ctx, cb, comp.data_ptr(), comp.numel(), bits, out.data_ptr(), syms.size)
        best = res["ms_total"] if best is None else min(best, res["ms_total"])
    ok = res["n_symbols"] == syms.size and np.array_equal(out[: syms.size].cpu().numpy(), syms)
    print(f"{name:>14}: maxlen {cb.maxlen} minlen {cb.minlen}, {syms.size} symbols, {nb} B in: "
          f"{best:.3f} ms = {syms.size / best / 1e6:.1f} GB/s out, sync {res['ms_sync']:.3f} emit {res['ms_emit']:.3f}  {'OK' if ok else 'MISMATCH'}")


def main():
    dev = torch.device("cuda:0")
    ctx = hb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_phase_timing("always")
    rng = np.random.default_rng(1)
    for log2n in (24, 26):
        n = 1 << log2n
        for name, lengths in (("even", [2, 2, 2, 4, 4, 4, 4]), ("even16", [2] * 3 + [4] * 3 + [6] * 3 + [8] * 3 + [10] * 3 + [12] * 3 + [14] * 3 + [16] * 4),
                              ("mult4", [4] * 15 + [8] * 16), ("mult3", [3] * 7 + [6] * 8), ("7or8", [7] * 64 + [8] * 128)):
            p = np.array([2.0 ** -l for l in lengths])
            syms = rng.choice(len(lengths), size=n, p=p / p.sum()).astype(np.uint8)
            tree, codes = tree_from_lengths(lengths)
            data, bits = encode(codes, syms)
            st = O.Stream(tree, data, bits, n)
            assert np.array_equal(O.simple_decode(st, bits=min(bits, 1 << 20))[:1000], syms[:1000])
            run(f"{name}/2^{log2n}", tree, data, bits, syms, ctx, dev)
        m = hb.Model(hb.MODEL_ENGLISH)
        f, syms = m.huff_file_cpu(7, n)
        run(f"english/2^{log2n}", f.tree, f.data, f.bits, syms, ctx, dev)


if __name__ == "__main__":
    main()
