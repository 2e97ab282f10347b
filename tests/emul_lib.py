"""ctypes driver of tests/emul/emul.cpp (CPU emulation of the CUDA tile algorithm).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import huffmandecoderongpus_b200 as hb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "emul.cpp")
SO = os.path.join(ROOT, "tests", "emul", "libemul.so")
CSRC = os.path.join(ROOT, "huffmandecoderongpus_b200", "csrc")


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "tiles", "rounds_total", "rounds_max", "probes_walk", "probes_rewalk", "probes_hyp",
        "probes_fix", "probes_emit", "warp_iters_walk", "warp_iters_rewalk", "warp_iters_hyp",
        "warp_iters_emit", "hyp_unmerged", "tiles_entry_nonzero", "long_probes")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC, os.path.join(CSRC, "hb_core.cuh"), os.path.join(CSRC, "hb_format.h")]
        if (not os.path.exists(SO)) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                                   SRC, "-o", SO])
        L = C.CDLL(SO)
        L.emul_run.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                               C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                               C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                               C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                               C.c_void_p, C.POINTER(Stats), C.c_uint32,
                               C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def gmod_of(lut, origin_byte=0):
    """(gmod, gorg) as make_args in csrc/hb_api.cu computes them; origin_byte = None: the shard's
    position in the stream is unknown -> only the power-of-two part of the gcd is usable"""
    g = int(lut["len_gcd"]) or 1
    if origin_byte is None:
        h = 1
        while h < 32 and g % (2 * h) == 0:
            h *= 2
        return h, 0
    return g, (origin_byte % g) * 8 % g


def words_of(data: np.ndarray, nbytes: int) -> np.ndarray:
    """compressed bytes -> little-endian u32 words (zero padded)"""
    n = (nbytes + 3) // 4
    buf = np.zeros(n * 4, dtype=np.uint8)
    buf[:nbytes] = data[:nbytes]
    return buf.view("<u4").copy()


def run(lut, words, bits_own, bits_avail, wpt=4, T=256, entry=0, base=0, emit=True,
        out_capacity=None, out_offset=0, emit_win=0, sync_mode=2, emit_mode=1, ep_wf=10, origin_byte=0):
    """Returns (out bytes, shard_map[32], result[4], stats dict, rc).
    sync_mode: 0 = probe sync kernel only, 1 = transducer kernel on full tiles (the
    product's default dispatch), 2 = both, failing (rc -101) unless they agree.
    emit_mode: 0 = byte-store emit walk, 1 = word-store walk (WPT >= 2; narrower test
    shapes fall back to bytes), 2 = flat walk (EP-table of ep_wf index bits) on every tile
    but the last one (wpt >= 4), word-store walk on the last, 3 = word-store walk with the E32-table
    (index width ep_wf, 0 = the E64 default), 4 = the same per warp tile of 32 subsequences."""
    cap = int(out_capacity if out_capacity is not None else bits_own + 64)
    raw = np.zeros(cap + 64 + out_offset, dtype=np.uint8)
    out = raw[out_offset:]
    smap = np.zeros(32, dtype=np.uint64)
    res = np.zeros(4, dtype=np.uint64)
    st = Stats()
    ent = np.ascontiguousarray(lut["entries"], dtype=np.uint32)
    words = np.ascontiguousarray(words, dtype=np.uint32)
    stab = np.ascontiguousarray(lut["stab"], dtype=np.uint32)
    etab = np.ascontiguousarray(lut["etab"], dtype=np.uint32)
    e64 = np.ascontiguousarray(lut["e64"], dtype=np.uint32)
    fsm = np.ascontiguousarray(lut["fsm"], dtype=np.uint16)
    fdepth = np.ascontiguousarray(lut["fsm_depth"], dtype=np.uint8)
    fpstep = np.ascontiguousarray(lut["fsm_pstep"], dtype=np.uint16)
    rc = lib().emul_run(ent.ctypes.data, lut["w1"], lut["maxlen"], lut["minlen"],
                        stab.ctypes.data, etab.ctypes.data, lut["wf"], words.ctypes.data,
                        words.size, bits_own, bits_avail, wpt, T, int(emit), entry, base,
                        out.ctypes.data, cap, smap.ctypes.data, res.ctypes.data, C.byref(st), emit_win,
                        sync_mode, lut["fsm_states"], fsm.ctypes.data, fdepth.ctypes.data,
                        fpstep.ctypes.data, emit_mode, e64.ctypes.data, lut["wf64"], ep_wf, *gmod_of(lut, origin_byte))
    return out, smap, res, st.as_dict(), rc


def decode(stream, wpt=4, T=256, lut=None, **kw):
    """Whole-stream emulated decode of an oracle_lib.Stream / hb.HuffFile."""
    lut = lut or hb.build_lut(stream.tree)
    w = words_of(stream.data, (stream.bits + 7) // 8)
    out, smap, res, st, rc = run(lut, w, stream.bits, stream.bits, wpt, T, **kw)
    return out[: int(res[0])], st, rc
