"""Per-phase device time of small decodes (the shipped corpora are launch-latency bound)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, huffmandecoderongpus_b200 as hb, oracle_lib as O
dev=torch.device("cuda:0")
wpt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx=hb.Context(0, stream=torch.cuda.current_stream().cuda_stream, words_per_thread=wpt)
if len(sys.argv) > 1: ctx.set_phase_timing(sys.argv[1])   # always | never | auto
for name in ("hello","paper1","news","world192","kjv","ecoli"):
    f=hb.HuffFile.load(O.corpus_path(name))
    cb=hb.Codebook(ctx,f.tree)
    n=(f.nbytes+15)//16*16+16
    comp=torch.zeros(n,dtype=torch.uint8,device=dev); comp[:f.nbytes]=torch.from_numpy(f.data[:f.nbytes]).to(dev)
    out=torch.zeros(f.usize+64,dtype=torch.uint8,device=dev)
    best=None
    for _ in range(30):
        r=hb.decode_device(ctx,cb,comp.data_ptr(),comp.numel(),f.bits,out.data_ptr(),f.usize)
        if best is None or r["ms_total"]<best["ms_total"]: best=r
    print(name, {k:round(best[k],4) for k in ("ms_sync","ms_scan","ms_emit","ms_total")}, best["launches"], best["tiles"])
