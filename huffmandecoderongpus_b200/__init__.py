"""huffmandecoderongpus_b200 -- ctypes binding of libhuffb200.so (include/huffb200.h).

The product is the C-ABI shared library (hand-written sm_100a CUDA + C host
code); this module only exposes it to Python tests and bench.py.  There is no
Python or CPU decode path: if the library is missing, or no CUDA device is
present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhuffb200.so")

HB_OK = 0
ERRORS = {
    -1: "HB_ERR_CUDA", -2: "HB_ERR_TREE", -3: "HB_ERR_CODELEN", -4: "HB_ERR_ARG",
    -5: "HB_ERR_NOMEM", -6: "HB_ERR_OUTPUT_FULL", -7: "HB_ERR_IO", -8: "HB_ERR_FORMAT",
    -9: "HB_ERR_STATE",
}

MODEL_ENGLISH, MODEL_FIBONACCI, MODEL_DNA, MODEL_UNIFORM8 = 0, 1, 2, 3

NODE_DTYPE = np.dtype([("sym", np.uint8), ("izero", np.int32), ("ione", np.int32)], align=True)


class HuffError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {ERRORS.get(code, code)} {detail}".strip())


class Node(C.Structure):
    _fields_ = [("sym", C.c_uint8), ("izero", C.c_int32), ("ione", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("n_symbols", C.c_uint64), ("out_base", C.c_uint64),
                ("exit_offset", C.c_uint32), ("entry_offset", C.c_uint32),
                ("launches", C.c_uint32), ("tiles", C.c_uint32),
                ("ms_total", C.c_float), ("ms_sync", C.c_float),
                ("ms_scan", C.c_float), ("ms_emit", C.c_float)]


class MultiResult(C.Structure):
    _fields_ = [("n_symbols", C.c_uint64), ("n_devices", C.c_int32), ("launches", C.c_uint32),
                ("ms_device_max", C.c_float), ("ms_wall", C.c_float),
                ("shard_symbols", C.c_uint64 * 8), ("shard_ms", C.c_float * 8)]


class HuffFileC(C.Structure):
    _fields_ = [("nodes", C.c_int32), ("wide", C.c_int32), ("bits", C.c_uint64),
                ("usize", C.c_uint64), ("tree", C.POINTER(Node)),
                ("data", C.POINTER(C.c_uint8))]


class ModelC(C.Structure):
    _fields_ = [("nodes", C.c_int32), ("tree", Node * 511), ("cum", C.c_uint32 * 256),
                ("symtab", C.c_uint8 * 256), ("code", C.c_uint32 * 256), ("codelen", C.c_uint8 * 256),
                ("maxlen", C.c_uint32), ("minlen", C.c_uint32), ("nsyms", C.c_uint32)]


class LutC(C.Structure):
    _fields_ = [("entries", C.POINTER(C.c_uint32)), ("n_entries", C.c_uint32),
                ("w1", C.c_uint32), ("maxlen", C.c_uint32), ("minlen", C.c_uint32),
                ("n_leaves", C.c_uint32), ("wf", C.c_uint32),
                ("stab", C.POINTER(C.c_uint32)), ("etab", C.POINTER(C.c_uint32)),
                ("code", C.c_uint32 * 256), ("codelen", C.c_uint8 * 256),
                ("fsm_states", C.c_uint32), ("fsm", C.POINTER(C.c_uint16)),
                ("fsm_bstep", C.POINTER(C.c_uint16)), ("fsm_depth", C.c_uint8 * 256),
                ("fsm_pstep", C.c_uint16 * 256), ("e64", C.POINTER(C.c_uint32)),
                ("fsm_node", C.c_int32 * 256), ("node_state", C.POINTER(C.c_int32)),
                ("wf64", C.c_uint32), ("implied_avg_len", C.c_double), ("len_gcd", C.c_uint32)]


class RefCompressedData(C.Structure):
    """struct CompressedData, reference framework/huffdata.h:26-32"""
    _fields_ = [("bits", C.c_int), ("nodes", C.c_int), ("uncompressedsize", C.c_int),
                ("tree", C.c_void_p), ("data", C.c_void_p)]


class RefUnCompressedData(C.Structure):
    """struct UnCompressedData, reference framework/huffdata.h:34-37"""
    _fields_ = [("uncompressedsize", C.c_int), ("data", C.c_void_p)]


_lib = None


def lib():
    """Load libhuffb200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C huffmandecoderongpus_b200` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no Python/CPU fallback for the decode path.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.hb_strerror.restype = C.c_char_p
    L.hb_strerror.argtypes = [i32]
    L.hb_version.restype = C.c_char_p
    L.hb_last_error.restype = C.c_char_p
    L.hb_last_error.argtypes = [vp]
    L.hb_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.hb_ctx_destroy.argtypes = [vp]
    L.hb_ctx_destroy.restype = None
    L.hb_ctx_configure.argtypes = [vp, i32, i32]
    L.hb_ctx_set_sync_path.argtypes = [vp, i32]
    L.hb_ctx_set_phase_timing.argtypes = [vp, i32]
    L.hb_codebook_download_table.argtypes = [vp, i32, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.hb_ctx_set_sync_copies.argtypes = [vp, i32]
    L.hb_ctx_set_emit_lane_subsequences.argtypes = [vp, i32]
    L.hb_ctx_last_emit_kernel.argtypes = [vp]
    L.hb_ctx_last_emit_kernel.restype = C.c_char_p
    L.hb_ctx_set_emit_path.argtypes = [vp, i32]
    L.hb_ctx_set_emit_table.argtypes = [vp, i32, i32]
    L.hb_ctx_set_shard_origin.argtypes = [vp, u64, i32]
    L.hb_ctx_sync.argtypes = [vp]
    L.hb_ctx_set_host_chunk.argtypes = [vp, u64]
    L.hb_ctx_timing_begin.argtypes = [vp, i32]
    L.hb_ctx_timing_collect.argtypes = [vp, C.POINTER(C.c_double * 4), C.POINTER(i32)]
    L.hb_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(u64)]
    L.hb_codebook_create.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.hb_codebook_destroy.argtypes = [vp]
    L.hb_codebook_destroy.restype = None
    L.hb_codebook_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.hb_decode_device.argtypes = [vp, vp, vp, u64, u64, vp, u64, C.POINTER(Result)]
    L.hb_multi_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.hb_multi_destroy.argtypes = [vp]
    L.hb_multi_destroy.restype = None
    L.hb_multi_devices.argtypes = [vp]
    L.hb_multi_last_error.argtypes = [vp]
    L.hb_multi_last_error.restype = C.c_char_p
    L.hb_multi_load.argtypes = [vp, vp, i32, vp, u64]
    L.hb_multi_generate.argtypes = [vp, i32, u64, u64, C.POINTER(u64)]
    L.hb_multi_decode.argtypes = [vp, C.POINTER(MultiResult)]
    L.hb_multi_download.argtypes = [vp, vp, u64]
    L.hb_multi_verify.argtypes = [vp, i32, u64, C.POINTER(u64)]
    L.hb_multi_decode_host.argtypes = [vp, vp, i32, vp, u64, vp, u64, C.POINTER(MultiResult)]
    L.hb_host_pin.argtypes = [vp, u64]
    L.hb_host_unpin.argtypes = [vp]
    L.hb_shard_result.argtypes = [vp, C.POINTER(Result)]
    L.hb_shard_map_host.argtypes = [vp, vp, vp, u64, u64, u64, vp]
    L.hb_shard_emit_host.argtypes = [vp, vp, vp, vp, u64, C.POINTER(Result)]
    L.hb_decode_onethread.argtypes = [vp, vp, i32, vp, u64, vp, u64, C.POINTER(Result)]
    L.hb_shard_map.argtypes = [vp, vp, vp, u64, u64, u64, vp]
    L.hb_shard_compose.argtypes = [vp, vp, i32, i32, vp]
    L.hb_peer_export.argtypes = [vp, vp]
    L.hb_peer_connect.argtypes = [vp, i32, i32, vp]
    L.hb_peer_connect_local.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.hb_shard_exchange.argtypes = [vp, u64, vp]
    L.hb_peer_close.argtypes = [vp]
    L.hb_peer_disconnect.argtypes = [vp]
    L.hb_shard_emit.argtypes = [vp, vp, vp, u64, u64, u64, vp, vp, u64, C.POINTER(Result)]
    L.hb_decode_host.argtypes = [vp, vp, i32, vp, u64, vp, u64, C.POINTER(Result)]
    L.hb_huff_load.argtypes = [C.c_char_p, C.POINTER(HuffFileC)]
    L.hb_huff_save.argtypes = [C.c_char_p, C.POINTER(HuffFileC), i32]
    L.hb_huff_free.argtypes = [C.POINTER(HuffFileC)]
    L.hb_huff_free.restype = None
    L.hb_model_build.argtypes = [i32, C.POINTER(ModelC)]
    L.hb_gen_symbols_cpu.argtypes = [C.POINTER(ModelC), u64, u64, u64, vp]
    L.hb_gen_symbols_cpu.restype = None
    L.hb_encode_bits_cpu.argtypes = [C.POINTER(ModelC), vp, u64]
    L.hb_encode_bits_cpu.restype = u64
    L.hb_encode_cpu.argtypes = [C.POINTER(ModelC), vp, u64, vp]
    L.hb_encode_cpu.restype = None
    L.hb_gen_encode_device.argtypes = [vp, C.POINTER(ModelC), u64, u64, u64, vp, u64, C.POINTER(u64)]
    L.hb_gen_verify_device.argtypes = [vp, C.POINTER(ModelC), u64, u64, u64, vp, C.POINTER(u64)]
    L.hb_gen_count_bits_device.argtypes = [vp, C.POINTER(ModelC), u64, u64, u64, C.POINTER(u64)]
    L.hb_lut_build.argtypes = [vp, i32, i32, i32, C.POINTER(LutC)]
    L.hb_lut_sizeof.restype = C.c_size_t
    assert L.hb_lut_sizeof() == C.sizeof(LutC), "LutC no longer mirrors struct hb_lut (csrc/hb_lut.h)"
    L.hb_lut_free.argtypes = [C.POINTER(LutC)]
    L.hb_lut_free.restype = None
    for name in ("b200Approach",):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [C.POINTER(RefCompressedData), C.POINTER(RefUnCompressedData), vp]
    L.b200ApproachLastDeviceMs.restype = C.c_double
    L.b200ApproachLastSymbols.restype = C.c_ulonglong
    L.b200ApproachShutdown.restype = None
    _lib = L
    return L


def _check(rc, where, ctx=None):
    if rc != HB_OK:
        detail = ""
        if ctx is not None and rc == -1:
            detail = lib().hb_last_error(ctx).decode(errors="replace")
        raise HuffError(rc, where, detail)


# ---- host-side pieces (no GPU needed) ----------------------------------------

class HuffFile:
    """A .huff stream loaded by the C host loader (hb_huff_load)."""

    def __init__(self, tree, data, bits, usize, wide=False):
        self.tree = np.ascontiguousarray(tree, dtype=NODE_DTYPE)
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.bits = int(bits)
        self.usize = int(usize)
        self.wide = bool(wide)
        self.nodes = int(self.tree.shape[0])

    @property
    def nbytes(self):
        return (self.bits + 7) // 8

    @classmethod
    def load(cls, path):
        f = HuffFileC()
        _check(lib().hb_huff_load(os.fsencode(path), C.byref(f)), f"hb_huff_load({path})")
        try:
            tree = np.ctypeslib.as_array(C.cast(f.tree, C.POINTER(C.c_uint8)),
                                         shape=(f.nodes * 12,)).copy().view(NODE_DTYPE)
            nbytes = (f.bits + 7) // 8
            data = np.ctypeslib.as_array(f.data, shape=(nbytes + 16,)).copy()
            return cls(tree, data, f.bits, f.usize, f.wide)
        finally:
            lib().hb_huff_free(C.byref(f))

    def save(self, path, wide=None):
        wide = self.wide if wide is None else wide
        f = HuffFileC(self.nodes, int(wide), self.bits, self.usize,
                      C.cast(self.tree.ctypes.data, C.POINTER(Node)),
                      C.cast(self.data.ctypes.data, C.POINTER(C.c_uint8)))
        _check(lib().hb_huff_save(os.fsencode(path), C.byref(f), int(wide)), "hb_huff_save")


def build_lut(tree, w1_max=0, w2_max=0):
    """Host table construction (hb_lut_build); returns a dict of numpy data."""
    tree = np.ascontiguousarray(tree, dtype=NODE_DTYPE)
    lut = LutC()
    _check(lib().hb_lut_build(tree.ctypes.data, int(tree.shape[0]), w1_max, w2_max, C.byref(lut)),
           "hb_lut_build")
    try:
        ns = int(lut.fsm_states)
        return {
            "fsm_states": ns,
            "fsm": (np.ctypeslib.as_array(lut.fsm, shape=(ns * 256,)).copy() if ns
                    else np.zeros(1, np.uint16)),
            "fsm_bstep": (np.ctypeslib.as_array(lut.fsm_bstep, shape=(ns * 2,)).copy() if ns
                          else np.zeros(2, np.uint16)),
            "fsm_depth": np.array(lut.fsm_depth, dtype=np.uint8),
            "fsm_pstep": np.array(lut.fsm_pstep, dtype=np.uint16),
            "entries": np.ctypeslib.as_array(lut.entries, shape=(lut.n_entries,)).copy(),
            "w1": lut.w1, "maxlen": lut.maxlen, "minlen": lut.minlen, "n_leaves": lut.n_leaves,
            "wf": lut.wf,
            "stab": np.ctypeslib.as_array(lut.stab, shape=(1 << lut.wf,)).copy(),
            "etab": np.ctypeslib.as_array(lut.etab, shape=(1 << lut.wf,)).copy(),
            "wf64": lut.wf64, "implied_avg_len": lut.implied_avg_len, "len_gcd": lut.len_gcd,
            "e64": np.ctypeslib.as_array(lut.e64, shape=(2 << lut.wf64,)).copy(),
            "code": np.array(lut.code, dtype=np.uint32), "codelen": np.array(lut.codelen, dtype=np.uint8),
        }
    finally:
        lib().hb_lut_free(C.byref(lut))


class Model:
    """Synthetic-stream model (hb_model_build): tree + sampler + code table."""

    def __init__(self, kind):
        self.kind = kind
        self.c = ModelC()
        _check(lib().hb_model_build(kind, C.byref(self.c)), "hb_model_build")
        self.nodes = self.c.nodes
        self.maxlen, self.minlen, self.nsyms = self.c.maxlen, self.c.minlen, self.c.nsyms
        raw = np.frombuffer(bytes(self.c.tree), dtype=np.uint8)[: self.nodes * 12]
        self.tree = raw.view(NODE_DTYPE).copy()

    def symbols_cpu(self, seed, first, n):
        out = np.empty(n, dtype=np.uint8)
        lib().hb_gen_symbols_cpu(C.byref(self.c), seed, first, n, out.ctypes.data)
        return out

    def encode_cpu(self, syms):
        syms = np.ascontiguousarray(syms, dtype=np.uint8)
        bits = lib().hb_encode_bits_cpu(C.byref(self.c), syms.ctypes.data, syms.size)
        out = np.zeros((bits + 7) // 8 + 32, dtype=np.uint8)
        lib().hb_encode_cpu(C.byref(self.c), syms.ctypes.data, syms.size, out.ctypes.data)
        return out, int(bits)

    def huff_file_cpu(self, seed, n):
        syms = self.symbols_cpu(seed, 0, n)
        data, bits = self.encode_cpu(syms)
        return HuffFile(self.tree, data, bits, n, wide=bits >= 2 ** 31), syms


# ---- device objects -------------------------------------------------------------

class Context:
    def __init__(self, device=0, stream=None, words_per_thread=0, ctas_per_sm=0):
        self.h = C.c_void_p()
        # stream: a cudaStream_t handle as an int; 0 (torch's default stream) is passed as
        # cudaStreamLegacy (0x1), None lets the context own a non-blocking stream
        if stream is not None and int(stream) == 0:
            stream = 1
        _check(lib().hb_ctx_create(device, C.c_void_p(stream) if stream is not None else None, C.byref(self.h)),
               "hb_ctx_create (a CUDA device is required; there is no CPU fallback)")
        self.device = device
        if words_per_thread or ctas_per_sm:
            self.configure(words_per_thread, ctas_per_sm)

    def configure(self, words_per_thread=0, ctas_per_sm=0):
        _check(lib().hb_ctx_configure(self.h, words_per_thread, ctas_per_sm), "hb_ctx_configure")

    def set_phase_timing(self, mode):
        """"auto", "always" or "never": whether decodes record the per-phase events."""
        _check(lib().hb_ctx_set_phase_timing(self.h, {"auto": 0, "always": 1, "never": 2}[mode]),
               "hb_ctx_set_phase_timing")

    def set_emit_lane_subsequences(self, n=1):
        """hb_emit32w_kernel: consecutive subsequences per lane (1 or 2)"""
        _check(lib().hb_ctx_set_emit_lane_subsequences(self.h, n), "hb_ctx_set_emit_lane_subsequences")

    def last_emit_kernel(self):
        """name of the emit kernel the last decode used for the bulk of its tiles"""
        return lib().hb_ctx_last_emit_kernel(self.h).decode()

    def set_sync_copies(self, log2_copies=-1):
        """copies of the transducer table in the sync kernel (log2: 0, 1, 2; -1 = automatic)"""
        _check(lib().hb_ctx_set_sync_copies(self.h, log2_copies), "hb_ctx_set_sync_copies")

    def set_emit_path(self, path):
        """emit kernel: "bytes" (byte stores), "words" (whole 32-bit words, one loop per stream
        word), "flat" (hb_emitf_kernel on every tile but the last; experimental, slower) or
        "auto" (= words)."""
        _check(lib().hb_ctx_set_emit_path(self.h, {"auto": 0, "bytes": 1, "words": 2, "flat": 3, "words32": 4, "words32w": 5, "words64w": 6}[path]),
               "hb_ctx_set_emit_path")

    def set_emit_table(self, index_bits=0, log2_copies=-1):
        """EP-table geometry of the flat emit kernel (0 / -1 = automatic)."""
        _check(lib().hb_ctx_set_emit_table(self.h, index_bits, log2_copies), "hb_ctx_set_emit_table")

    def set_shard_origin(self, first_byte=None):
        """index, in the whole stream, of the first byte of the shards decoded next (None = unknown)"""
        _check(lib().hb_ctx_set_shard_origin(self.h, 0 if first_byte is None else first_byte,
                                             0 if first_byte is None else 1), "hb_ctx_set_shard_origin")

    def set_sync_path(self, path):
        """"auto" (transducer sync kernel for streams of two waves of tiles or more), "probe"
        (probe kernel only) or "fsm" (transducer kernel whenever the code has one)."""
        _check(lib().hb_ctx_set_sync_path(self.h, {"auto": 0, "probe": 1, "fsm": 2}[path]),
               "hb_ctx_set_sync_path")

    def set_host_chunk(self, nbytes):
        _check(lib().hb_ctx_set_host_chunk(self.h, nbytes), "hb_ctx_set_host_chunk")

    def sync(self):
        _check(lib().hb_ctx_sync(self.h), "hb_ctx_sync", self.h)

    def timing_begin(self, max_steps):
        _check(lib().hb_ctx_timing_begin(self.h, max_steps), "hb_ctx_timing_begin", self.h)

    def timing_collect(self):
        ms = (C.c_double * 4)()
        n = C.c_int()
        _check(lib().hb_ctx_timing_collect(self.h, C.byref(ms), C.byref(n)), "hb_ctx_timing_collect", self.h)
        return {"sync": ms[0], "scan": ms[1], "emit": ms[2], "total": ms[3], "steps": n.value}

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        _check(lib().hb_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)),
               "hb_device_info")
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def close(self):
        if self.h:
            # a codebook must not outlive its context (hb_codebook_destroy uses the context's
            # device and stream): close the ones still open first
            for cb in list(getattr(self, "_codebooks", ())):
                cb.close()
            lib().hb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Codebook:
    def __init__(self, ctx: Context, tree):
        self.ctx = ctx
        tree = np.ascontiguousarray(tree, dtype=NODE_DTYPE)
        self.h = C.c_void_p()
        _check(lib().hb_codebook_create(ctx.h, tree.ctypes.data, int(tree.shape[0]), C.byref(self.h)),
               "hb_codebook_create", ctx.h)
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().hb_codebook_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.maxlen, self.minlen, self.w1, self.n_entries = a.value, b.value, c.value, d.value
        if not hasattr(ctx, "_codebooks"):
            ctx._codebooks = []
        ctx._codebooks.append(self)

    def table(self, which):
        """Device-resident code table copied back to the host: "lut", "stab", "etab", "e64"
        (u32 arrays) or "fsm" (u16)."""
        idx = {"lut": 0, "stab": 1, "etab": 2, "e64": 3, "fsm": 5}[which]
        buf = np.zeros(1 << 20, dtype=np.uint8)
        n = C.c_uint64()
        _check(lib().hb_codebook_download_table(self.h, idx, buf.ctypes.data, buf.size, C.byref(n)),
               "hb_codebook_download_table", self.ctx.h)
        return buf[: n.value].view(np.uint16 if which == "fsm" else np.uint32).copy()

    def close(self):
        if self.h:
            if self.ctx.h:   # after Context.close() the library has already released it
                lib().hb_codebook_destroy(self.h)
            self.h = C.c_void_p()
            if self in getattr(self.ctx, "_codebooks", ()):
                self.ctx._codebooks.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _res_dict(r: Result):
    return {k: getattr(r, k) for k, _ in Result._fields_}


def decode_device(ctx: Context, cb: Codebook, d_comp: int, comp_bytes: int, bits: int,
                  d_out: int, out_capacity: int, want_result=True):
    """hb_decode_device on raw device pointers (e.g. torch tensor .data_ptr()).  want_result=False:
    the kernels are queued and nothing is read back (no host synchronisation)."""
    r = Result()
    _check(lib().hb_decode_device(ctx.h, cb.h, d_comp, comp_bytes, bits, d_out, out_capacity,
                                  C.byref(r) if want_result else None), "hb_decode_device", ctx.h)
    return _res_dict(r) if want_result else None


def shard_map(ctx, cb, d_comp, comp_bytes, bits_own, bits_avail, d_map):
    _check(lib().hb_shard_map(ctx.h, cb.h, d_comp, comp_bytes, bits_own, bits_avail, d_map),
           "hb_shard_map", ctx.h)


def shard_compose(ctx, d_all_maps, n_ranks, rank, d_entry_base):
    _check(lib().hb_shard_compose(ctx.h, d_all_maps, n_ranks, rank, d_entry_base),
           "hb_shard_compose", ctx.h)


PEER_HANDLE_BYTES = 64


def peer_export(ctx) -> bytes:
    """hb_peer_export: the 64-byte IPC handle of this rank's exchange table (pass it to every rank)."""
    buf = C.create_string_buffer(PEER_HANDLE_BYTES)
    _check(lib().hb_peer_export(ctx.h, buf), "hb_peer_export", ctx.h)
    return buf.raw


def peer_connect(ctx, rank: int, handles) -> None:
    """hb_peer_connect: handles = the exported handles of all ranks, in rank order.  The caller must
    put a barrier between this call and the first shard_exchange."""
    blob = b"".join(handles)
    assert len(blob) == PEER_HANDLE_BYTES * len(handles)
    _check(lib().hb_peer_connect(ctx.h, rank, len(handles), blob), "hb_peer_connect", ctx.h)


def peer_connect_local(ctxs) -> None:
    """hb_peer_connect_local for contexts of this process: ctxs[r] is rank r."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    for r, c in enumerate(ctxs):
        _check(lib().hb_peer_connect_local(c.h, r, len(ctxs), arr), "hb_peer_connect_local", c.h)


def peer_disconnect(ctx) -> None:
    """hb_peer_disconnect: unmap the other ranks' tables (call on every rank, then a barrier, before closing)."""
    _check(lib().hb_peer_disconnect(ctx.h), "hb_peer_disconnect", ctx.h)


def shard_exchange(ctx, seq: int, d_entry_base: int) -> None:
    """hb_shard_exchange: between shard_map and shard_emit; replaces all-gather + shard_compose."""
    _check(lib().hb_shard_exchange(ctx.h, seq, d_entry_base), "hb_shard_exchange", ctx.h)


def shard_emit(ctx, cb, d_comp, comp_bytes, bits_own, bits_avail, d_entry_base, d_out,
               out_capacity, want_result=True):
    r = Result()
    _check(lib().hb_shard_emit(ctx.h, cb.h, d_comp, comp_bytes, bits_own, bits_avail,
                               d_entry_base, d_out, out_capacity,
                               C.byref(r) if want_result else None), "hb_shard_emit", ctx.h)
    return _res_dict(r) if want_result else None


def shard_map_host(ctx, cb, h_comp: np.ndarray, comp_bytes, bits_own, bits_avail, d_map):
    _check(lib().hb_shard_map_host(ctx.h, cb.h, h_comp.ctypes.data, comp_bytes, bits_own, bits_avail, d_map),
           "hb_shard_map_host", ctx.h)


def shard_emit_host(ctx, cb, d_entry_base, h_out: np.ndarray):
    r = Result()
    _check(lib().hb_shard_emit_host(ctx.h, cb.h, d_entry_base, h_out.ctypes.data, h_out.size, C.byref(r)),
           "hb_shard_emit_host", ctx.h)
    return _res_dict(r)


def decode_host(ctx: Context, tree, data, bits: int, out: np.ndarray):
    """hb_decode_host: host buffers in, host buffer out (upload, decode, download)."""
    tree = np.ascontiguousarray(tree, dtype=NODE_DTYPE)
    r = Result()
    _check(lib().hb_decode_host(ctx.h, tree.ctypes.data, int(tree.shape[0]), data.ctypes.data, bits,
                                out.ctypes.data, out.size, C.byref(r)), "hb_decode_host", ctx.h)
    return _res_dict(r)


def b200_approach(tree, data, bits: int, usize: int):
    """Call the drop-in approach exactly as the reference's evaluate() would:
    reference-layout structs, caller-allocated zeroed output of usize+3 bytes
    (framework/decodeUtil.c:37-43).  Returns the output buffer [0:usize]."""
    tree = np.ascontiguousarray(tree, dtype=NODE_DTYPE)
    out = np.zeros(usize + 3, dtype=np.uint8)
    cd = RefCompressedData(bits, int(tree.shape[0]), usize, tree.ctypes.data, data.ctypes.data)
    ucd = RefUnCompressedData(usize, out.ctypes.data)
    lib().b200Approach(C.byref(cd), C.byref(ucd), None)
    return out[:usize]


def gen_encode_device(ctx: Context, model: Model, seed: int, first: int, n: int, d_comp: int,
                      capacity: int) -> int:
    bits = C.c_uint64()
    _check(lib().hb_gen_encode_device(ctx.h, C.byref(model.c), seed, first, n, d_comp, capacity,
                                      C.byref(bits)), "hb_gen_encode_device", ctx.h)
    return bits.value


def gen_count_bits_device(ctx: Context, model: Model, seed: int, first: int, n: int) -> int:
    bits = C.c_uint64()
    _check(lib().hb_gen_count_bits_device(ctx.h, C.byref(model.c), seed, first, n, C.byref(bits)),
           "hb_gen_count_bits_device", ctx.h)
    return bits.value


def gen_verify_device(ctx: Context, model: Model, seed: int, first: int, n: int, d_out: int) -> int:
    bad = C.c_uint64()
    _check(lib().hb_gen_verify_device(ctx.h, C.byref(model.c), seed, first, n, d_out, C.byref(bad)),
           "hb_gen_verify_device", ctx.h)
    return bad.value


# ---- one process, N devices -----------------------------------------------------------

def _multi_dict(r: MultiResult):
    n = r.n_devices
    return {"n_symbols": r.n_symbols, "n_devices": n, "launches": r.launches,
            "ms_device_max": r.ms_device_max, "ms_wall": r.ms_wall,
            "shard_symbols": list(r.shard_symbols[:n]), "shard_ms": list(r.shard_ms[:n])}


class Multi:
    """hb_multi_*: one stream over several GPUs of this process (byte-range shards, maps
    exchanged by peer copies)."""

    def __init__(self, n_devices=0, devices=None):
        self.h = C.c_void_p()
        arr = None
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            n_devices = len(devices)
        _check(lib().hb_multi_create(arr, n_devices, C.byref(self.h)), "hb_multi_create")
        self.n = lib().hb_multi_devices(self.h)

    def _ck(self, rc, where):
        if rc != 0:
            raise HuffError(rc, where, lib().hb_multi_last_error(self.h).decode())

    def decode_host(self, tree, data, bits, out):
        tree = np.ascontiguousarray(tree)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        r = MultiResult()
        self._ck(lib().hb_multi_decode_host(self.h, tree.ctypes.data, len(tree), data.ctypes.data, bits,
                                            out.ctypes.data, out.size, C.byref(r)), "hb_multi_decode_host")
        return _multi_dict(r)

    def load(self, tree, data, bits):
        tree = np.ascontiguousarray(tree)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        self._ck(lib().hb_multi_load(self.h, tree.ctypes.data, len(tree), data.ctypes.data, bits), "hb_multi_load")

    def generate(self, kind, seed, n_symbols):
        bits = C.c_uint64()
        self._ck(lib().hb_multi_generate(self.h, kind, seed, n_symbols, C.byref(bits)), "hb_multi_generate")
        return bits.value

    def decode(self):
        r = MultiResult()
        self._ck(lib().hb_multi_decode(self.h, C.byref(r)), "hb_multi_decode")
        return _multi_dict(r)

    def download(self, out):
        self._ck(lib().hb_multi_download(self.h, out.ctypes.data, out.size), "hb_multi_download")

    def verify(self, kind, seed):
        bad = C.c_uint64()
        self._ck(lib().hb_multi_verify(self.h, kind, seed, C.byref(bad)), "hb_multi_verify")
        return bad.value

    def close(self):
        if self.h:
            lib().hb_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_onethread(ctx: Context, tree, data, bits: int, out: np.ndarray):
    """the reference's onethread approach: the whole stream on one device thread (debug aid)"""
    tree = np.ascontiguousarray(tree)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    r = Result()
    _check(lib().hb_decode_onethread(ctx.h, tree.ctypes.data, len(tree), data.ctypes.data, bits,
                                     out.ctypes.data, out.size, C.byref(r)), "hb_decode_onethread", ctx.h)
    return _res_dict(r)
