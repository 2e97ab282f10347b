#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 speculative parallel Huffman decoder.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): decoded GB/s, device-timed, on the synthetic English-like
stream of config 4 (2^30 symbols per GPU, order-0 histogram of bible.txt, encoded
in the reference's .huff bit format by the bundled generator).  A step is one
whole decode of the stream(s): N=1 decodes the 2^30-symbol stream; N>1 decodes
one stream of N * 2^30 symbols partitioned by compressed byte range, one shard
per GPU, with one NCCL all-gather of the 32-entry boundary maps (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SEED = 0x48554646  # "HUFF"
WORKLOADS = {
    # name: (model kind, log2 symbols per GPU, description)
    "english1g": (0, 30, "synthetic English-like (order-0 histogram of bible.txt), 2^30 symbols per GPU"),
    "fib4g": (1, 32, "synthetic Fibonacci-skewed 256-symbol alphabet (max code length > 20), 2^32 symbols per GPU"),
    "english64m": (0, 26, "synthetic English-like, 2^26 symbols per GPU (quick check)"),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (the interface behind
    nvidia-smi's clocks.sm / clocks_event_reasons.*) every ~10 ms by a thread;
    only samples taken inside the timed region are kept."""

    def __init__(self, index):
        self.index = index
        self.samples = []   # (t, sm_mhz, reasons_bitmask)
        self.stop_flag = False
        self.ok = False
        self.sm_max = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; map through CUDA_VISIBLE_DEVICES when it is numeric
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        if not self.ok:
            return
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.th.join(timeout=2)
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        sm = sorted(s[1] for s in inside)
        mask = 0
        for s in inside:
            mask |= s[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                "samples": len(inside), "reasons": sorted(v for k, v in names.items() if mask & k)}


# ---------------------------------------------------------------------------------
# CPU legs (the oracle / the compiled reference are used ONLY here, as the baseline)
# ---------------------------------------------------------------------------------

def cpu_sample(kind, log2n):
    """First 2^log2n symbols of the workload, built by the CPU generator."""
    import numpy as np
    import huffmandecoderongpus_b200 as hb
    import oracle_lib as O
    m = hb.Model(kind)
    f, syms = m.huff_file_cpu(SEED, 1 << log2n)
    return O.Stream(f.tree, f.data, f.bits, f.usize), syms


def cpu_time_paths(st, syms, reps):
    """Reference protocol (framework/decodeUtil.c:30-70): first run byte-checked,
    min wall time over the runs.  Returns {path: seconds}, kind."""
    import numpy as np
    import oracle_lib as O
    use_ref = O.ref() is not None and st.bits < 2 ** 31
    out = {}

    def timed(fn):
        best = None
        for r in range(reps):
            t0 = time.perf_counter()
            got = fn()
            dt = time.perf_counter() - t0
            if r == 0:
                assert np.array_equal(got, syms), "CPU baseline output differs from the generated symbols"
            best = dt if best is None else min(best, dt)
        return best

    if use_ref:
        out["simpleDecode"] = timed(lambda: O.ref_decode(st, "simpleDecode"))
        out["jumptableApproach_jb8"] = timed(lambda: O.ref_decode(st, "jumptableApproach", 8))
        kind = "reference"
    else:
        out["simpleDecode"] = timed(lambda: O.simple_decode(st))
        out["jumptableApproach_jb8"] = timed(lambda: O.jumptable_decode(st, 8))
        kind = "port"
    return out, kind


def run_reference_arm(args, kind, desc):
    """--impl reference: the reference's own CPU decode path on this box's host
    cores (single-threaded: the reference has no threads), each step a bounded
    sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    log2s = min(args.cpu_sample_log2, 26)
    st, syms = cpu_sample(kind, log2s)
    import oracle_lib as O
    use_ref = O.ref() is not None
    fn = (lambda: O.ref_decode(st, "jumptableApproach", 8)) if use_ref else (lambda: O.jumptable_decode(st, 8))
    import numpy as np
    for _ in range(max(args.warmup, 1)):
        got = fn()
    assert np.array_equal(got, syms)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    gbs = args.steps * syms.size / dt / 1e9
    sample = (f"first 2^{log2s} symbols of the workload ({st.nbytes} compressed bytes), "
              f"{'unmodified reference' if use_ref else 'oracle port of'} jumptableApproach(jumpbits=8), 1 thread")
    line = {
        "impl": "reference", "metric": "decoded_GB_per_s", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "sample": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1,
                         "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="english1g", choices=list(WORKLOADS))
    ap.add_argument("--wpt", type=int, default=0, help="words per thread (0 = library default)")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--ep-wf", type=int, default=0, help="EP-table index bits of the flat emit kernel (0 = auto)")
    ap.add_argument("--ep-copies-log2", type=int, default=-1, help="log2 of the EP-table copies (-1 = auto)")
    ap.add_argument("--emit-path", default="auto", choices=["auto", "bytes", "words", "flat"],
                    help="staging stores: bytes (hb_emit_kernel) or whole words (hb_emitw_kernel, default)")
    ap.add_argument("--sync-path", default="auto", choices=["auto", "probe", "fsm"],
                    help="auto: transducer sync kernel on full tiles; probe: probe sync kernel only")
    ap.add_argument("--cpu-sample-log2", type=int, default=27)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--host-chunk-mib", type=int, default=0,
                    help="chunk size of the pipelined host path in MiB (0 = library default, 32)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    kind, log2n, desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, kind, desc)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import huffmandecoderongpus_b200 as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_per = 1 << log2n
    n_total = n_per * world
    ctx = hb.Context(local, stream=torch.cuda.current_stream().cuda_stream,
                     words_per_thread=args.wpt, ctas_per_sm=args.ctas_per_sm)
    ctx.set_sync_path(args.sync_path)
    if args.host_chunk_mib:
        ctx.set_host_chunk(args.host_chunk_mib << 20)
    ctx.set_emit_path(args.emit_path)
    ctx.set_emit_table(args.ep_wf, args.ep_copies_log2)
    model = hb.Model(kind)
    cb = hb.Codebook(ctx, model.tree)

    # ---- build the stream on the device (setup, untimed) -------------------------
    bits_total = hb.gen_count_bits_device(ctx, model, SEED, 0, n_total)
    nbytes_total = (bits_total + 7) // 8
    per = (nbytes_total // world) // 16 * 16
    a = rank * per
    b = nbytes_total if rank == world - 1 else (rank + 1) * per
    last = rank == world - 1
    whole = torch.zeros((nbytes_total + 15) // 16 * 16 + 64, dtype=torch.uint8, device=dev)
    assert hb.gen_encode_device(ctx, model, SEED, 0, n_total, whole.data_ptr(), whole.numel()) == bits_total
    halo_end = min(nbytes_total, b + 16)
    comp = torch.zeros((halo_end - a + 15) // 16 * 16 + 16, dtype=torch.uint8, device=dev)
    comp[: halo_end - a] = whole[a:halo_end]
    del whole
    torch.cuda.empty_cache()
    bits_own = bits_total - 8 * a if last else 8 * (b - a)
    bits_avail = bits_own if last else min(bits_total - 8 * a, 8 * (halo_end - a))
    comp_bytes_own = (bits_own + 7) // 8

    cap = int(n_per * 1.02) + (1 << 16)
    out = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
    my_map = torch.zeros(32, dtype=torch.int64, device=dev)
    all_maps = torch.zeros(32 * world, dtype=torch.int64, device=dev)
    eb = torch.zeros(4, dtype=torch.int64, device=dev)

    def step(want_result=False):
        hb.shard_map(ctx, cb, comp.data_ptr(), comp.numel(), bits_own, bits_avail, my_map.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(all_maps, my_map)
            hb.shard_compose(ctx, all_maps.data_ptr(), world, rank, eb.data_ptr())
            ebp = eb.data_ptr()
        else:
            ebp = None
        return hb.shard_emit(ctx, cb, comp.data_ptr(), comp.numel(), bits_own, bits_avail, ebp,
                             out.data_ptr(), cap, want_result=want_result)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of this very configuration (untimed) -------------------------
    res = step(want_result=True)
    n_mine = res["n_symbols"]
    launches_per_step = res["launches"] + (1 if world > 1 else 0)   # + hb_compose_kernel
    bad = hb.gen_verify_device(ctx, model, SEED, res["out_base"], n_mine, out.data_ptr())
    tot = torch.tensor([n_mine, bad], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    assert int(tot[0]) == n_total and int(tot[1]) == 0, f"decode mismatch: {tot.tolist()} vs {n_total}"

    # ---- timed region ---------------------------------------------------------------
    for _ in range(args.warmup - 1):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timing_begin(args.steps)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    w1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    phases = ctx.timing_collect()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0])
    ms_step = ms_max / args.steps
    value = n_total / (ms_step * 1e-3) / 1e9
    in_gbs = nbytes_total / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (rank 0's CUDA events over the timed steps) --
    peak, peak_src = measured_peak()
    b_alg = comp_bytes_own + n_mine                 # compressed read once + decoded written once
    k_ms = {k: phases[k] / max(phases["steps"], 1) for k in ("sync", "scan", "emit", "total")}
    dom = max(("sync", "emit"), key=lambda k: k_ms[k])
    sync_name = "hb_sync_kernel" if args.sync_path == "probe" else "hb_fsm_sync_kernel"
    emit_name = {"bytes": "hb_emit_kernel", "words": "hb_emitw_kernel"}.get(args.emit_path, "hb_emitf_kernel")
    dom_name = {"sync": sync_name, "emit": emit_name}[dom]
    achieved = b_alg / (k_ms[dom] * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this
    # same command (profiles/r01_traffic.json); only quoted for the configuration it was taken on
    traffic = None
    try:
        if args.workload == "english1g" and (args.wpt or 8) == 8:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)["kernels"][dom_name]["dram_bytes"]
    except Exception:
        traffic = None
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": b_alg,
        "kernel_ms": {sync_name: k_ms["sync"], "hb_scan_*+hb_fix": k_ms["scan"], emit_name: k_ms["emit"]},
        "decode_achieved": b_alg / (k_ms["total"] * 1e-3) / 1e9,
        "decode_frac": b_alg / (k_ms["total"] * 1e-3) / 1e9 / peak,
        "kernel_share_of_step": k_ms[dom] / k_ms["total"],
    }

    # ---- end to end through host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_comp = torch.empty(comp.numel(), dtype=torch.uint8, pin_memory=True)
        h_comp.copy_(comp)
        h_out = torch.empty(cap + 64, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        steps_e = max(args.e2e_steps, 1)

        def e2e_step():
            if world == 1:
                # the user-facing call: host buffers in, host buffer out
                hb.decode_host(ctx, model.tree, h_comp.numpy(), bits_total, h_out.numpy()[:cap])
            else:
                comp.copy_(h_comp, non_blocking=True)
                step()
                h_out[:n_mine].copy_(out[:n_mine], non_blocking=True)
                torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        assert np.array_equal(h_out[:4096].numpy(), out[:4096].cpu().numpy())
        e2e = {"value": n_total * steps_e / dt / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(comp_bytes_own), "d2h_bytes_per_step": int(n_mine),
               "steps": steps_e, "ms_per_step": dt / steps_e * 1e3,
               "path": "hb_decode_host (pinned host buffers)" if world == 1
                       else "pinned H2D + hb_shard_map/all_gather/hb_shard_emit + pinned D2H per rank"}
        del h_comp, h_out

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        log2s = min(args.cpu_sample_log2, log2n)
        st, syms = cpu_sample(kind, log2s)
        secs, ckind = cpu_time_paths(st, syms, reps=3)
        best = min(secs, key=secs.get)
        cpu = {"value": syms.size / secs[best] / 1e9, "unit": "GB/s", "cores": 1, "kind": ckind,
               "path": best,
               "sample": f"first 2^{log2s} symbols of the workload ({st.nbytes} compressed bytes), "
                         f"min of 3 runs, first run byte-checked, 1 thread of {os.cpu_count()} host CPUs",
               "all_paths_GBps": {k: syms.size / v / 1e9 for k, v in secs.items()}}

    if rank == 0:
        line = {
            "metric": "decoded_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "seed": SEED,
                       "symbols_total": n_total, "compressed_bytes_total": int(nbytes_total),
                       "bits_total": int(bits_total), "max_code_length": model.maxlen,
                       "words_per_thread": args.wpt or 8, "sync_path": args.sync_path, "emit_path": args.emit_path,
                       "parallelism": f"byte-range shards x{world}, 1 NCCL all-gather of 32-entry maps" if world > 1 else "single GPU",
                       "l2": "inputs and outputs larger than L2 (no flush needed)",
                       "compressed_input_GB_per_s": in_gbs},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    cb.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
