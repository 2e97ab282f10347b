"""Copies-only probe of the host <-> device paths the end-to-end numbers run over.

  python tools/pcie_probe.py                                  one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
                                                              N ranks, one GPU each, all at once

No decode kernel runs: only the pinned-memory copies of bench.py's end-to-end legs (per rank
588.5 MB up, 1 GiB down: one english1g shard).  With N ranks the copies of all ranks run
simultaneously between barriers, which is what the N-GPU end-to-end leg does; the aggregate
tells whether the host (memory, root complexes) or the decode limits that leg.
Prints one JSON line on rank 0."""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

NIN, NOUT = 588_538_415, 1 << 30
d_out = torch.empty(NOUT, dtype=torch.uint8, device=dev)
h_out = torch.empty(NOUT, dtype=torch.uint8, pin_memory=True)
h_in = torch.empty(NIN, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(NIN, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=4):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def up():
    d_in.copy_(h_in, non_blocking=True)


def down():
    h_out.copy_(d_out, non_blocking=True)


def up_then_down():          # the N > 1 end-to-end leg: upload, (exchange), download
    d_in.copy_(h_in, non_blocking=True)
    h_out.copy_(d_out, non_blocking=True)


def both():                  # both directions at once
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


K = 18
ci, co = (NIN + K - 1) // K, (NOUT + K - 1) // K


def pipeline():              # hb_decode_host's pattern: chunk k's download behind chunk k's upload
    evs = []
    with torch.cuda.stream(s1):
        for k in range(K):
            d_in[k * ci:(k + 1) * ci].copy_(h_in[k * ci:(k + 1) * ci], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s1)
            evs.append(e)
    with torch.cuda.stream(s2):
        for k in range(K):
            s2.wait_event(evs[k])
            h_out[k * co:(k + 1) * co].copy_(d_out[k * co:(k + 1) * co], non_blocking=True)


res = {}
for name, fn, nbytes in (("h2d", up, NIN), ("d2h", down, NOUT), ("h2d_then_d2h", up_then_down, NIN + NOUT),
                         ("h2d_and_d2h_concurrent", both, NIN + NOUT), ("chunked_pipeline", pipeline, NIN + NOUT)):
    dt = timed(fn)
    res[name] = {"ms": dt * 1e3, "per_rank_GBps": nbytes / dt / 1e9, "aggregate_GBps": world * nbytes / dt / 1e9,
                 "decoded_equiv_GBps_aggregate": world * NOUT / dt / 1e9 if "d2h" in name or "pipeline" in name else None}
if rank == 0:
    print(json.dumps({"probe": "pinned host<->device copies, no kernels", "ranks": world,
                      "bytes_up_per_rank": NIN, "bytes_down_per_rank": NOUT, "results": res}), flush=True)
if world > 1:
    dist.destroy_process_group()
