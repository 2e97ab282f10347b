"""Where the transducer sync kernel (plus one probe launch for the partial tile) overtakes a
single probe-kernel launch: device time of mid-size English-like streams per sync path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import huffmandecoderongpus_b200 as hb  # noqa: E402

SEED = 0x48554646
dev = torch.device("cuda:0")
m = hb.Model(hb.MODEL_ENGLISH)
for log2n in (20, 22, 23, 24, 25, 26, 27):
    n = 1 << log2n
    row = []
    for path in ("probe", "fsm", "auto"):
        ctx = hb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
        ctx.set_sync_path(path)
        ctx.set_phase_timing("never")
        cb = hb.Codebook(ctx, m.tree)
        bits = hb.gen_count_bits_device(ctx, m, SEED, 0, n)
        nb = (bits + 7) // 8
        comp = torch.zeros((nb + 15) // 16 * 16 + 32, dtype=torch.uint8, device=dev)
        hb.gen_encode_device(ctx, m, SEED, 0, n, comp.data_ptr(), comp.numel())
        out = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
        best = min(hb.decode_device(ctx, cb, comp.data_ptr(), comp.numel(), bits, out.data_ptr(), n)["ms_total"]
                   for _ in range(20))
        row.append(best)
        cb.close()
        ctx.close()
    tiles = (bits + 65535) // 65536
    print(f"2^{log2n} symbols, {nb / 1e6:7.2f} MB in, {tiles:6d} tiles: probe {row[0]:.4f}  fsm {row[1]:.4f}  auto {row[2]:.4f} ms")
