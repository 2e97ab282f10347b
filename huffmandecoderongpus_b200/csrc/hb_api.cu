/*
 * hb_api.cu -- implementation of the C ABI declared in include/huffb200.h:
 * context, codebook upload, kernel orchestration, timing.  No CPU decode path
 * exists here: every decode entry point launches the sm_100a kernels or fails.
 */
#include "huffb200.h"
#include "hb_lut.h"
#include "hb_kernels.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#define HB_NEV 5

struct hb_codebook {
    hb_ctx *ctx;
    hb_lut lut;        /* host copy (entries kept for the encoder / tests) */
    uint32_t *d_lut;
    uint8_t *d_fsm;    /* [fsm u16 x states*256][depth u8 x 256][pstep u16 x 256], or NULL */
    double implied_avg_len;   /* sum over leaves of 2^-len * len */
    /* E32-tables of hb_emit32_kernel, one per index width, built on the device on first use (by a kernel,
     * stream-ordered) and kept: the emit CTAs then only copy theirs from L2 */
    uint32_t *d_e32[HB_E32_WF_MAX + 1] = {};
    uint32_t *d_e64[HB_E32_WF_MAX + 1] = {};   /* E64-tables of other widths than lut.wf64 (hb_emitw_kernel), same scheme */
};

struct hb_buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct hb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaDeviceProp prop;
    int wpt = 8;
    int ctas_per_sm = 0;
    int sync_path = HB_SYNC_AUTO;
    int emit_path = HB_EMIT_AUTO;
    int ep_wf = 0, ep_rshift = -1;   /* EP-/E32-table geometry of the flat / 32-bit emit kernels (0 / -1 = automatic) */
    int fsm_copies = -1;             /* transducer table copies in the sync kernel: log2; -1 = automatic = one (measured: 4 copies
                                      * in one 1024-thread CTA 0.494 ms against 0.438 ms with one copy per CTA and 48 warps per SM) */
    int warp_spl = 1;                /* subsequences per lane of the warp-autonomous emit kernel: 2 halves the per-warp-tile
                                      * work per stream bit but was measured slower (english1g emit 0.628 against 0.558 ms: the
                                      * second subsequence's words are loaded in the middle of the chain, and the doubled
                                      * staging slices leave room for a 14-bit table only) */
    bool auto_warp_emit = true;      /* HB_EMIT_AUTO picks the warp-autonomous E32 kernel (english1g emit 0.576 ms against 0.622) */
    uint32_t e64_wide = 11;          /* E64 index width for short codes in large streams (set from measurements) */
    uint32_t smem_base = 0x400;      /* shared-window address at which a kernel's dynamic shared memory begins (measured) */
    int phase_timing = HB_PHASES_AUTO;
    bool fuse_small = false;      /* set by hb_decode_device: single shard, nobody reads the map between the phases */
    bool map_fused = false;       /* the last hb_shard_map left up/top to hb_scan_small_kernel */
    bool map_notop = false;       /* ... skipped hb_scan_top_kernel: hb_scan_downfix_kernel follows the one chain itself */
    uint32_t last_launches = 0;   /* kernels launched by the last map + emit pair */
    const char *last_emit = "";   /* the emit kernel the last hb_shard_emit launched for the bulk of the tiles */
    cudaEvent_t ev0[HB_NEV];   /* default event set */
    cudaEvent_t *ev = nullptr; /* set used by the current step */
    cudaEvent_t *tim_ev = nullptr; /* optional ring: tim_cap steps x HB_NEV events */
    int tim_cap = 0, tim_n = 0;
    char err[256];
    /* scratch (grow-only) */
    hb_buf subs, tmaps, wmaps, cmaps, cprefix, tile_entry, tile_base, misc;
    /* misc layout (u64 words): [0..31] shard map, [32..35] result, [36] status, [40..42] zero entry_base */
    /* state of the last hb_shard_map */
    bool have_map = false;
    uint32_t map_ntiles = 0, map_ncta = 0;
    int map_wpt = 0;
    /* host-buffer path: device buffers and the last code table are kept across
     * calls (the reference harness calls an approach 26 times with the same tree) */
    hb_buf d_comp, d_out;
    hb_codebook *host_cb = nullptr;
    hb_node_abi *host_tree = nullptr;
    int host_nodes = 0;
    /* pipelined host path: copy streams, per-chunk events, pinned staging */
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_aux = nullptr;            /* the partial tile's sync kernel, beside the transducer kernel */
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    cudaEvent_t *pipe_ev = nullptr;   /* 2 per chunk: upload done, emit done */
    int pipe_cap = 0;
    uint64_t *h_pipe = nullptr;       /* pinned: 32 map words + 4 entry/base words per chunk */
    hb_buf d_eb;
    uint64_t pipe_chunk_bytes = 32ull << 20;
    bool pipe_chunk_default = true;  /* ... not set by the caller: 64 MiB chunks for streams of at least 512 MiB */
    cudaEvent_t pipe_t0 = nullptr, pipe_t1 = nullptr;
    uint64_t *h_res = nullptr; /* pinned, 8 words */
    uint64_t hs_readable = 0, hs_own = 0, hs_avail = 0;   /* shard of the last hb_shard_map_host */
    bool origin_known = false;    /* hb_ctx_set_shard_origin */
    uint64_t origin_byte = 0;
    /* map exchange by peer stores (hb_peer_*, one process per GPU) */
    uint64_t *peer_tab = nullptr;                 /* this rank's exchange table (cudaMalloc: exportable) */
    uint64_t *peer_ptr[HB_MULTI_MAX] = {};        /* tables of the ranks to the right, opened from their handles */
    int peer_rank = -1, peer_n = 0;
    bool peer_ipc = false;                        /* peer_ptr came from cudaIpcOpenMemHandle */
};

static const char *const k_errs[] = {
    "ok", "CUDA runtime error", "malformed Huffman tree", "codeword longer than 32 bits",
    "bad argument", "out of memory", "output buffer too small", "I/O error",
    "not a HUFF/HUF8 file", "call order",
};

extern "C" const char *hb_strerror(int code) {
    int i = -code;
    if (i < 0 || i >= (int)(sizeof(k_errs) / sizeof(k_errs[0]))) return "unknown error";
    return k_errs[i];
}

extern "C" const char *hb_version(void) { return "huffb200 0.1 (sm_100a)"; }

extern "C" const char *hb_last_error(const hb_ctx *ctx) { return ctx ? ctx->err : "no context"; }

static int cuda_fail(hb_ctx *ctx, cudaError_t e, const char *what) {
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
    return HB_ERR_CUDA;
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);     \
    } while (0)

static int ensure(hb_ctx *ctx, hb_buf &b, size_t bytes) {
    if (bytes <= b.cap) return HB_OK;
    if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return HB_ERR_NOMEM;
    }
    b.cap = want;
    return HB_OK;
}

__global__ void hb_smem_base_kernel(uint32_t *out) {
    extern __shared__ __align__(16) uint32_t smem_probe[];
    if (threadIdx.x == 0) *out = (uint32_t)__cvta_generic_to_shared(smem_probe);
}

extern "C" int hb_ctx_create(int device, void *cuda_stream, hb_ctx **out) {
    if (!out) return HB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return HB_ERR_CUDA;   /* no CPU fallback */
    }
    hb_ctx *ctx = new (std::nothrow) hb_ctx();
    if (!ctx) return HB_ERR_NOMEM;
    ctx->err[0] = 0;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess) {
        delete ctx;
        return HB_ERR_CUDA;
    }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return HB_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    int nev = 0;
    for (; nev < HB_NEV; nev++)
        if (cudaEventCreate(&ctx->ev0[nev]) != cudaSuccess) break;
    ctx->ev = ctx->ev0;
    if (nev < HB_NEV || cudaMallocHost((void **)&ctx->h_res, 8 * sizeof(uint64_t)) != cudaSuccess) {
        cudaGetLastError();
        for (int i = 0; i < nev; i++) cudaEventDestroy(ctx->ev0[i]);
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return HB_ERR_CUDA;
    }
    /* where dynamic shared memory begins in the shared window: hb_emit32_kernel places its table at
     * a multiple of the table size (it verifies the address itself and reports HB_ST_LAYOUT) */
    {
        uint32_t *d_base = nullptr;
        if (cudaMalloc((void **)&d_base, sizeof(uint32_t)) == cudaSuccess) {
            hb_smem_base_kernel<<<1, 32, 16, ctx->stream>>>(d_base);
            uint32_t b = 0;
            if (cudaMemcpyAsync(&b, d_base, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                cudaStreamSynchronize(ctx->stream) == cudaSuccess)
                ctx->smem_base = b & 0xffffffu;    /* the bits above are the CTA's rank in its cluster */
            cudaFree(d_base);
        }
        cudaGetLastError();
    }
    *out = ctx;
    return HB_OK;
}

extern "C" void hb_ctx_destroy(hb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    hb_buf *bufs[] = { &ctx->subs, &ctx->tmaps, &ctx->wmaps, &ctx->cmaps, &ctx->cprefix,
                       &ctx->tile_entry, &ctx->tile_base, &ctx->misc, &ctx->d_comp, &ctx->d_out };
    for (hb_buf *b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->host_cb) hb_codebook_destroy(ctx->host_cb);
    free(ctx->host_tree);
    for (int i = 0; i < 2 * ctx->pipe_cap; i++) cudaEventDestroy(ctx->pipe_ev[i]);
    free(ctx->pipe_ev);
    if (ctx->h_pipe) cudaFreeHost(ctx->h_pipe);
    if (ctx->pipe_t0) { cudaEventDestroy(ctx->pipe_t0); cudaEventDestroy(ctx->pipe_t1); }
    if (ctx->s_aux) { cudaStreamDestroy(ctx->s_aux); cudaEventDestroy(ctx->aux_fork); cudaEventDestroy(ctx->aux_join); }
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->d_eb.p) cudaFree(ctx->d_eb.p);
    for (int i = 0; i < HB_NEV; i++) cudaEventDestroy(ctx->ev0[i]);
    for (int i = 0; i < ctx->tim_cap * HB_NEV; i++) cudaEventDestroy(ctx->tim_ev[i]);
    free(ctx->tim_ev);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    hb_peer_close(ctx);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int hb_ctx_configure(hb_ctx *ctx, int words_per_thread, int ctas_per_sm) {
    if (!ctx) return HB_ERR_ARG;
    if (words_per_thread != 0 && words_per_thread != 4 && words_per_thread != 8 &&
        words_per_thread != 16)
        return HB_ERR_ARG;
    if (ctas_per_sm < 0 || ctas_per_sm > 32) return HB_ERR_ARG;
    ctx->wpt = words_per_thread ? words_per_thread : 8;
    ctx->ctas_per_sm = ctas_per_sm;
    return HB_OK;
}

extern "C" int hb_ctx_set_sync_path(hb_ctx *ctx, int path) {
    if (!ctx || path < HB_SYNC_AUTO || path > HB_SYNC_FSM) return HB_ERR_ARG;
    ctx->sync_path = path;
    return HB_OK;
}

extern "C" const char *hb_ctx_last_emit_kernel(const hb_ctx *ctx) { return ctx ? ctx->last_emit : ""; }

extern "C" int hb_ctx_set_emit_lane_subsequences(hb_ctx *ctx, int n) {
    if (!ctx || (n != 1 && n != 2)) return HB_ERR_ARG;
    ctx->warp_spl = n;
    return HB_OK;
}

extern "C" int hb_ctx_set_sync_copies(hb_ctx *ctx, int log2_copies) {
    if (!ctx || log2_copies < -1 || log2_copies > 2) return HB_ERR_ARG;
    ctx->fsm_copies = log2_copies;
    return HB_OK;
}

extern "C" int hb_ctx_set_phase_timing(hb_ctx *ctx, int mode) {
    if (!ctx || mode < HB_PHASES_AUTO || mode > HB_PHASES_NEVER) return HB_ERR_ARG;
    ctx->phase_timing = mode;
    return HB_OK;
}

extern "C" int hb_ctx_set_emit_path(hb_ctx *ctx, int path) {
    if (!ctx || (path != HB_EMIT_AUTO && path != HB_EMIT_BYTES && path != HB_EMIT_WORDS && path != HB_EMIT_FLAT &&
                 path != HB_EMIT_WORDS32 && path != HB_EMIT_WORDS32W && path != HB_EMIT_WORDS64W))
        return HB_ERR_ARG;
    ctx->emit_path = path;
    return HB_OK;
}

extern "C" int hb_ctx_set_emit_table(hb_ctx *ctx, int index_bits, int log2_copies) {
    if (!ctx) return HB_ERR_ARG;
    if (index_bits != 0 && (index_bits < HB_EP_WF_MIN || index_bits > HB_E32_WF_MAX)) return HB_ERR_ARG;
    if (log2_copies < -1 || log2_copies > 4) return HB_ERR_ARG;
    ctx->ep_wf = index_bits;
    ctx->ep_rshift = log2_copies;
    return HB_OK;
}

extern "C" int hb_ctx_set_shard_origin(hb_ctx *ctx, uint64_t first_byte, int known) {
    if (!ctx) return HB_ERR_ARG;
    ctx->origin_known = known != 0;
    ctx->origin_byte = known ? first_byte : 0;
    return HB_OK;
}

extern "C" int hb_ctx_set_host_chunk(hb_ctx *ctx, uint64_t bytes) {
    if (!ctx) return HB_ERR_ARG;
    ctx->pipe_chunk_bytes = bytes ? bytes : (32ull << 20);
    ctx->pipe_chunk_default = bytes == 0;
    return HB_OK;
}

extern "C" int hb_ctx_sync(hb_ctx *ctx) {
    if (!ctx) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

extern "C" int hb_device_info(hb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                              uint64_t *total_mem) {
    if (!ctx) return HB_ERR_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (total_mem) *total_mem = (uint64_t)ctx->prop.totalGlobalMem;
    return HB_OK;
}

/* ---- codebook ------------------------------------------------------------ */

extern "C" int hb_codebook_create(hb_ctx *ctx, const hb_node_abi *tree, int nodes,
                                  hb_codebook **out) {
    if (!ctx || !tree || !out) return HB_ERR_ARG;
    *out = nullptr;
    hb_codebook *cb = new (std::nothrow) hb_codebook();
    if (!cb) return HB_ERR_NOMEM;
    cb->ctx = ctx;
    cb->d_lut = nullptr;
    cb->d_fsm = nullptr;
    /* host: validation, state numbering, single-symbol table, per-symbol codes; the
     * multi-symbol tables and the transducer table are built on the device */
    int rc = hb_lut_build_small(tree, nodes, 0, 0, &cb->lut);
    if (rc != HB_OK) { delete cb; return rc; }
    cb->implied_avg_len = cb->lut.implied_avg_len;
    cudaSetDevice(ctx->device);
    /* device layout: [single-symbol LUT][S-table][E-table][E64-table] */
    const size_t n1 = cb->lut.n_entries, nf = (size_t)1 << cb->lut.wf;
    const size_t ns = cb->lut.fsm_states;
    /* ONE stream-ordered allocation: [tables][transducer + depth + partial steps][build
     * inputs: node array, state of every node, node of every state] */
    const size_t lut_bytes = (sizeof(uint32_t) * (n1 + 4 * nf) + 15) & ~(size_t)15;
    const size_t fsm_bytes = ns ? ns * 512 + 256 + 512 : 0;
    const size_t tree_bytes = sizeof(hb_node_abi) * (size_t)nodes;
    const size_t in_bytes = ((tree_bytes + 15) & ~(size_t)15) + sizeof(int32_t) * ((size_t)nodes + 256);
    uint8_t *d_all = nullptr, *d_in = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_all, lut_bytes + fsm_bytes + in_bytes, ctx->stream);
    if (e == cudaSuccess) {
        cb->d_lut = (uint32_t *)d_all;
        cb->d_fsm = ns ? d_all + lut_bytes : nullptr;
        d_in = d_all + lut_bytes + fsm_bytes;
    }
    hb_build_args ba;
    memset(&ba, 0, sizeof(ba));
    if (e == cudaSuccess) {
        ba.tree = (const hb_node_abi *)d_in;
        int32_t *d_ns = (int32_t *)(d_in + ((tree_bytes + 15) & ~(size_t)15));
        ba.node_state = d_ns;
        ba.state_node = d_ns + nodes;
        ba.nstates = (uint32_t)ns;
        ba.wf = cb->lut.wf;
        ba.wf64 = cb->lut.wf64;
        ba.stab = cb->d_lut + n1;
        ba.etab = cb->d_lut + n1 + nf;
        ba.e64 = cb->d_lut + n1 + 2 * nf;
        ba.fsm = ns ? (uint16_t *)cb->d_fsm : nullptr;
        e = cudaMemcpyAsync(d_in, tree, tree_bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && ns)
            e = cudaMemcpyAsync(d_ns, cb->lut.node_state, sizeof(int32_t) * (size_t)nodes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && ns)
            e = cudaMemcpyAsync(d_ns + nodes, cb->lut.fsm_node, sizeof(int32_t) * 256, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(cb->d_lut, cb->lut.entries, sizeof(uint32_t) * n1, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && ns)
        e = cudaMemcpyAsync(cb->d_fsm + ns * 512, cb->lut.fsm_depth, 256, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && ns)
        e = cudaMemcpyAsync(cb->d_fsm + ns * 512 + 256, cb->lut.fsm_pstep, 512, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const size_t entries = nf + ns * 256;
        hb_build_tables_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, ctx->stream>>>(ba);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   /* the host arrays are reused */
    if (e != cudaSuccess) {
        if (d_all) cudaFreeAsync(d_all, ctx->stream);
        hb_lut_free(&cb->lut);
        delete cb;
        return cuda_fail(ctx, e, "codebook build");
    }
    *out = cb;
    return HB_OK;
}

extern "C" void hb_codebook_destroy(hb_codebook *cb) {
    if (!cb) return;
    cudaSetDevice(cb->ctx->device);
    if (cb->d_lut) cudaFreeAsync(cb->d_lut, cb->ctx->stream);   /* d_fsm lives in the same allocation */
    for (uint32_t *p : cb->d_e32) if (p) cudaFreeAsync(p, cb->ctx->stream);
    for (uint32_t *p : cb->d_e64) if (p) cudaFreeAsync(p, cb->ctx->stream);
    hb_lut_free(&cb->lut);
    delete cb;
}

extern "C" int hb_codebook_download_table(const hb_codebook *cb, int which, void *dst, uint64_t capacity,
                                          uint64_t *bytes) {
    if (!cb || !dst || !bytes) return HB_ERR_ARG;
    hb_ctx *ctx = cb->ctx;
    const size_t n1 = cb->lut.n_entries, nf = (size_t)1 << cb->lut.wf, ns = cb->lut.fsm_states;
    const uint8_t *src = nullptr;
    size_t n = 0;
    switch (which) {
    case HB_TABLE_LUT: src = (const uint8_t *)cb->d_lut; n = 4 * n1; break;
    case HB_TABLE_S:   src = (const uint8_t *)(cb->d_lut + n1); n = 4 * nf; break;
    case HB_TABLE_E:   src = (const uint8_t *)(cb->d_lut + n1 + nf); n = 4 * nf; break;
    case HB_TABLE_E64: src = (const uint8_t *)(cb->d_lut + n1 + 2 * nf); n = (size_t)8 << cb->lut.wf64; break;
    case HB_TABLE_FSM: src = cb->d_fsm; n = ns * 512; break;
    default: return HB_ERR_ARG;
    }
    *bytes = n;
    if (n > capacity) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n) CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

extern "C" int hb_codebook_info(const hb_codebook *cb, uint32_t *maxlen, uint32_t *minlen,
                                uint32_t *w1, uint32_t *n_entries) {
    if (!cb) return HB_ERR_ARG;
    if (maxlen) *maxlen = cb->lut.maxlen;
    if (minlen) *minlen = cb->lut.minlen;
    if (w1) *w1 = cb->lut.w1;
    if (n_entries) *n_entries = cb->lut.n_entries;
    return HB_OK;
}

/* ---- launch helpers ------------------------------------------------------ */

/* dynamic shared memory (the fast table is a static 16 KB array in each kernel) */
template <int WPT>
static size_t sync_smem_bytes(uint32_t) {
    return sizeof(uint32_t) * ((size_t)HB_T * WPT + 4 + (size_t)WPT * HB_T + HB_T + HB_T + 16);
}

static size_t emit_smem_bytes(uint32_t, uint32_t stage_bytes) {
    return sizeof(uint32_t) * 16 + stage_bytes;
}

/* Emit-kernel staging: a window of `win` output bytes per tile pass plus one
 * thread's worth of overhang (at most ceil(S / minlen) symbols) plus alignment.
 * The window is sized for the output a tile produces when symbols occur with
 * the probabilities the code table implies (avg code length = sum 2^-len * len),
 * with 20 % headroom; more compressible tiles simply take several windows. */
static void stage_geometry(const hb_codebook *cb, int wpt, uint32_t *win, uint32_t *stage_bytes) {
    const uint32_t S = 32u * (uint32_t)wpt;
    const uint32_t max_c = (S + cb->lut.minlen - 1) / cb->lut.minlen;
    const uint32_t worst = HB_T * max_c;
    double avg = cb->implied_avg_len > 1.0 ? cb->implied_avg_len : 1.0;
    uint32_t typical = (uint32_t)((double)HB_T * S * 1.2 / avg) + 64;
    uint32_t w = typical < worst ? typical : worst;
    if (w < max_c) w = max_c;   /* a window holds at least one thread's output (emit kernel invariant) */
    w = (w + 15u) & ~15u;
    *win = w;
    *stage_bytes = (w + max_c + 16u + 15u) & ~15u;
}

template <typename K>
static int grid_for(hb_ctx *ctx, K kernel, size_t smem, uint32_t ntiles, int *grid, int threads = HB_T) {
    /* static (16 KB fast table) + dynamic can exceed the 48 KB default: always opt in */
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) {
        snprintf(ctx->err, sizeof(ctx->err), "kernel does not fit an SM (smem %zu)", smem);
        return HB_ERR_CUDA;
    }
    if (ctx->ctas_per_sm > 0 && ctx->ctas_per_sm < occ) occ = ctx->ctas_per_sm;
    uint64_t g = (uint64_t)occ * (uint64_t)ctx->prop.multiProcessorCount;
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    *grid = (int)g;
    return HB_OK;
}

static int make_args(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp, uint64_t comp_bytes,
                     uint64_t bits_own, uint64_t bits_avail, hb_stream_args *a) {
    if (!ctx || !cb || cb->ctx != ctx) return HB_ERR_ARG;
    if (bits_avail < bits_own) return HB_ERR_ARG;
    if (bits_avail && !d_comp) return HB_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_comp) & 15u) != 0) return HB_ERR_ARG;
    if (comp_bytes < bits_avail / 8 + (bits_avail % 8 != 0)) return HB_ERR_ARG;
    const uint64_t tile_bits = (uint64_t)HB_T * 32u * (uint64_t)ctx->wpt;
    uint64_t ntiles = (bits_own + tile_bits - 1) / tile_bits;
    if (ntiles > 0x7fffffffull) return HB_ERR_ARG;
    a->words = (const uint32_t *)d_comp;
    a->nwords = (comp_bytes + 3) / 4;   /* the word holding the last byte must be readable */
    a->bits_own = bits_own;
    a->bits_avail = bits_avail;
    a->ntiles = (uint32_t)ntiles;
    a->lut = cb->d_lut;
    a->w1 = cb->lut.w1;
    a->maxlen = cb->lut.maxlen;
    a->minlen = cb->lut.minlen;
    a->fast = cb->d_lut + cb->lut.n_entries;   /* S-table; the E-table follows it */
    a->wf = cb->lut.wf;
    /* entry offsets that cannot occur are not followed (hb_stream_args.gmod) */
    if (ctx->origin_known) {
        a->gmod = cb->lut.len_gcd ? cb->lut.len_gcd : 1u;
        a->gorg = (uint32_t)((ctx->origin_byte % a->gmod) * 8u % a->gmod);
    } else {
        a->gmod = 1u;
        while (a->gmod < 32u && cb->lut.len_gcd % (2u * a->gmod) == 0u) a->gmod *= 2u;
        a->gorg = 0u;
    }
    return HB_OK;
}

static uint64_t *misc_words(hb_ctx *ctx) { return (uint64_t *)ctx->misc.p; }

/* The three inner events (sync | scan | emit boundaries) cost ~4 us each: a third of the
 * device time of a decode of the shipped corpora.  They are recorded for streams of more
 * than one scan CTA's worth of tiles, while a timing ring is armed (bench), or on request;
 * otherwise hb_result reports ms_total only. */
static bool phase_events(const hb_ctx *ctx, uint32_t ntiles) {
    if (ctx->phase_timing == HB_PHASES_ALWAYS) return true;
    if (ctx->phase_timing == HB_PHASES_NEVER) return false;
    return ntiles > 1024u || (ctx->tim_ev && ctx->ev != ctx->ev0);
}

/* Transducer sync kernel over tiles [0, n_full).  G groups of HB_T threads share one
 * table copy per CTA; G is chosen for the most resident warps per SM. */
template <int WPT, int G, int LC>
static int try_fsm_geometry(hb_ctx *ctx, size_t table_bytes, size_t extra_bytes, int *occ, size_t *smem) {
    *smem = (table_bytes << LC) + extra_bytes + (size_t)G * hb_fsm_group_words<WPT>() * sizeof(uint32_t);
    *occ = 0;
    if (*smem > (size_t)ctx->prop.sharedMemPerBlockOptin) return HB_OK;
    CK(cudaFuncSetAttribute(hb_fsm_sync_kernel<WPT, G, LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, hb_fsm_sync_kernel<WPT, G, LC>, G * HB_T, *smem));
    return HB_OK;
}

template <int WPT, int G, int LC>
static int run_fsm_sync(hb_ctx *ctx, const hb_stream_args &a, const hb_fsm_args &fa, uint32_t n_full,
                        int occ, size_t smem) {
    if (ctx->ctas_per_sm > 0 && ctx->ctas_per_sm < occ) occ = ctx->ctas_per_sm;
    uint64_t grid = (uint64_t)occ * (uint64_t)ctx->prop.multiProcessorCount;
    const uint64_t need = ((uint64_t)n_full + G - 1) / G;
    if (grid > need) grid = need;
    hb_fsm_sync_kernel<WPT, G, LC><<<(int)grid, G * HB_T, smem, ctx->stream>>>(
        a, fa, n_full, (uint16_t *)ctx->subs.p, (uint32_t *)ctx->tmaps.p);
    CK(cudaGetLastError());
    ctx->last_launches++;
    return HB_OK;
}

template <int WPT>
static int launch_fsm_sync(hb_ctx *ctx, const hb_codebook *cb, const hb_stream_args &a, uint32_t n_full) {
    const size_t ns = cb->lut.fsm_states;
    hb_fsm_args fa;
    fa.tab = (const uint16_t *)cb->d_fsm;
    fa.depth = cb->d_fsm + ns * 512;
    fa.pstep = (const uint16_t *)(cb->d_fsm + ns * 512 + 256);
    fa.nstates = (uint32_t)ns;
    const size_t table_bytes = ns * 512, extra = 256 + 512 + ((size_t)4 << a.w1);   /* + level 1 of the LUT */
    int rc;
    /* on request (A/B, tests): two or four copies of the table on disjoint banks, one CTA of four groups
     * per SM, 8-word subsequences.  Fewer bank-conflict replays, but 32 instead of 48 warps per SM, and the
     * walk is a chain of dependent lookups that needs the warps: measured slower, never automatic. */
    if (WPT == 8 && ctx->fsm_copies > 0) {
        int occ = 0;
        size_t smem = 0;
        if (ctx->fsm_copies != 1) {
            if ((rc = try_fsm_geometry<8, 4, 2>(ctx, table_bytes, extra, &occ, &smem))) return rc;
            if (occ > 0) return run_fsm_sync<8, 4, 2>(ctx, a, fa, n_full, occ, smem);
        }
        if (ctx->fsm_copies == 1) {
            if ((rc = try_fsm_geometry<8, 4, 1>(ctx, table_bytes, extra, &occ, &smem))) return rc;
            if (occ > 0) return run_fsm_sync<8, 4, 1>(ctx, a, fa, n_full, occ, smem);
        }
    }
    int occ[3] = {0, 0, 0};
    size_t smem[3] = {0, 0, 0};
    if ((rc = try_fsm_geometry<WPT, 1, 0>(ctx, table_bytes, extra, &occ[0], &smem[0]))) return rc;
    if ((rc = try_fsm_geometry<WPT, 2, 0>(ctx, table_bytes, extra, &occ[1], &smem[1]))) return rc;
    if ((rc = try_fsm_geometry<WPT, 4, 0>(ctx, table_bytes, extra, &occ[2], &smem[2]))) return rc;
    /* most resident warps; on a tie the larger group count (fewer table copies) */
    int best = -1, best_warps = 0;
    for (int i = 0; i < 3; i++) {
        const int warps = occ[i] * (1 << i) * (HB_T / 32);
        if (warps >= best_warps && warps > 0) { best = i; best_warps = warps; }
    }
    switch (best) {
    case 0: return run_fsm_sync<WPT, 1, 0>(ctx, a, fa, n_full, occ[0], smem[0]);
    case 1: return run_fsm_sync<WPT, 2, 0>(ctx, a, fa, n_full, occ[1], smem[1]);
    case 2: return run_fsm_sync<WPT, 4, 0>(ctx, a, fa, n_full, occ[2], smem[2]);
    }
    snprintf(ctx->err, sizeof(ctx->err), "transducer table (%zu B) does not fit an SM", table_bytes + extra);
    return HB_ERR_CUDA;
}

template <int WPT>
static int launch_map(hb_ctx *ctx, const hb_codebook *cb, const hb_stream_args &a, uint64_t *d_map) {
    const uint32_t ncta = (a.ntiles + 1023u) / 1024u;
    int rc;
    if ((rc = ensure(ctx, ctx->subs, sizeof(uint16_t) * (size_t)a.ntiles * HB_T))) return rc;
    if ((rc = ensure(ctx, ctx->tmaps, sizeof(uint32_t) * (size_t)a.ntiles * 32))) return rc;
    if ((rc = ensure(ctx, ctx->wmaps, sizeof(uint64_t) * (size_t)ncta * 32 * 32))) return rc;
    if ((rc = ensure(ctx, ctx->cmaps, sizeof(uint64_t) * (size_t)ncta * 32))) return rc;
    if ((rc = ensure(ctx, ctx->cprefix, sizeof(uint64_t) * (size_t)ncta * 32))) return rc;
    if ((rc = ensure(ctx, ctx->tile_entry, (size_t)a.ntiles))) return rc;
    if ((rc = ensure(ctx, ctx->tile_base, sizeof(uint64_t) * (size_t)a.ntiles))) return rc;

    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    ctx->last_launches = 0;
    uint32_t tile0 = 0;
    /* codes whose lengths share a factor that is not a power of two: the probe kernel, which can start
     * every subsequence's chain at the first offset of the right residue class (hb_first_entry) */
    const bool odd_factor = (a.gmod & (a.gmod - 1u)) != 0u;
    if (ctx->sync_path != HB_SYNC_PROBE && cb->d_fsm && a.minlen != a.maxlen && !odd_factor) {
        /* full tiles: byte-step transducer kernel -- from two waves of tiles on; below
         * that one probe-kernel launch for everything is quicker */
        const uint32_t n_full = (uint32_t)(a.bits_own / ((uint64_t)HB_T * 32u * WPT));
        const uint32_t min_tiles = ctx->sync_path == HB_SYNC_FSM ? 1u
                                 : 2u * HB_SYNC_MIN_CTAS * (uint32_t)ctx->prop.multiProcessorCount;
        if (n_full >= min_tiles) tile0 = n_full;
    }
    const bool split = tile0 > 0 && tile0 < a.ntiles;   /* transducer kernel + one partial tile */
    cudaStream_t probe_stream = ctx->stream;
    if (split) {
        /* the partial tile's single CTA goes FIRST, on a forked stream: it finds a free slot
         * now, whereas behind the persistent transducer kernel it would run alone at the end */
        if (!ctx->s_aux) {
            CK(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
        }
        CK(cudaEventRecord(ctx->aux_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_aux, ctx->aux_fork, 0));
        probe_stream = ctx->s_aux;
    }
    if (tile0 < a.ntiles) {
        size_t smem = sync_smem_bytes<WPT>(a.wf);
        int grid = 1;
        if ((rc = grid_for(ctx, hb_sync_kernel<WPT>, smem, a.ntiles - tile0, &grid))) return rc;
        hb_sync_kernel<WPT><<<grid, HB_T, smem, probe_stream>>>(a, tile0, (uint16_t *)ctx->subs.p,
                                                               (uint32_t *)ctx->tmaps.p);
        CK(cudaGetLastError());
        ctx->last_launches++;
    }
    if (split) CK(cudaEventRecord(ctx->aux_join, ctx->s_aux));
    if (tile0 > 0 && (rc = launch_fsm_sync<WPT>(ctx, cb, a, tile0))) return rc;
    if (split) CK(cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));
    if (phase_events(ctx, a.ntiles)) CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->map_fused = ctx->fuse_small && ncta == 1 && !d_map;
    ctx->map_notop = ctx->fuse_small && ncta > 1 && ncta <= HB_NOTOP_MAX_CTAS && !d_map;
    if (!ctx->map_fused) {
        hb_scan_up_kernel<<<ncta, HB_SCAN_T, 0, ctx->stream>>>((const uint32_t *)ctx->tmaps.p, a.ntiles,
                                                              (uint64_t *)ctx->wmaps.p,
                                                              (uint64_t *)ctx->cmaps.p);
        CK(cudaGetLastError());
        ctx->last_launches++;
        if (!ctx->map_notop) {
            hb_scan_top_kernel<<<1, 32, 0, ctx->stream>>>((const uint64_t *)ctx->cmaps.p, ncta,
                                                          (uint64_t *)ctx->cprefix.p, misc_words(ctx));
            CK(cudaGetLastError());
            ctx->last_launches++;
        }
    }
    if (d_map)
        CK(cudaMemcpyAsync(d_map, misc_words(ctx), 32 * sizeof(uint64_t), cudaMemcpyDeviceToDevice,
                           ctx->stream));
    if (phase_events(ctx, a.ntiles)) CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->have_map = true;
    ctx->map_ntiles = a.ntiles;
    ctx->map_ncta = ncta;
    ctx->map_wpt = WPT;
    return HB_OK;
}

/* Flat emit kernel over tiles [0, ntiles_run): EP-table geometry, group count, window. */
template <int G, int NP>
static int run_emit_flat(hb_ctx *ctx, const hb_stream_args &ae, uint32_t rshift, uint32_t ntiles_run,
                         void *d_out, uint64_t out_capacity, uint32_t win, uint32_t stage, size_t smem) {
    CK(cudaFuncSetAttribute(hb_emitf_kernel<8, G, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t grid = (uint64_t)ctx->prop.multiProcessorCount;
    const uint64_t need = ((uint64_t)ntiles_run + G - 1) / G;
    if (grid > need) grid = need;
    uint64_t *misc = misc_words(ctx);
    hb_emitf_kernel<8, G, NP><<<(int)grid, G * HB_T, smem, ctx->stream>>>(
        ae, rshift, ntiles_run, (const uint16_t *)ctx->subs.p, (const uint64_t *)ctx->tile_base.p,
        (uint8_t *)d_out, out_capacity, win, stage, (uint32_t *)(misc + 36));
    CK(cudaGetLastError());
    return HB_OK;
}

static int launch_emit_flat(hb_ctx *ctx, const hb_codebook *cb, const hb_stream_args &a, uint32_t ntiles_run,
                            void *d_out, uint64_t out_capacity, bool *launched) {
    *launched = false;
    hb_stream_args ae = a;
    uint32_t wf = ctx->ep_wf ? (uint32_t)ctx->ep_wf : 10u;
    if (wf > HB_EP_WF_MAX) wf = HB_EP_WF_MAX;
    if (wf > cb->lut.maxlen && cb->lut.maxlen >= HB_EP_WF_MIN) wf = cb->lut.maxlen;   /* no longer codeword exists */
    uint32_t rshift = ctx->ep_rshift >= 0 ? (uint32_t)ctx->ep_rshift : 4u;
    const size_t limit = (size_t)ctx->prop.sharedMemPerBlockOptin;
    while (rshift > 0 && ((size_t)8 << (wf + rshift)) > limit / 2 + limit / 8) rshift--;
    ae.wf = wf;
    const size_t table = (size_t)8 << (wf + rshift);
    /* staging window: the expected output of a tile plus head room, shrunk (never below one
     * thread's worth) until at least two groups fit beside the table */
    uint32_t win = 0, stage = 0;
    stage_geometry(cb, 8, &win, &stage);
    const uint32_t max_c = (256u + cb->lut.minlen - 1) / cb->lut.minlen;
    int G = 0;
    for (int g = 4; g >= 2 && !G; g--) {
        const size_t per = (limit - table) / (size_t)g;
        const size_t fixed = sizeof(uint32_t) * (size_t)hb_emitf_group_words<8>(0);
        if (per < fixed + max_c + 64u) continue;
        uint32_t w = (uint32_t)((per - fixed - max_c - 32u) & ~(size_t)15);
        if (g > 2 && w < win - win / 8) continue;     /* rather fewer groups than many windows per tile */
        if (w < ((max_c + 15u) & ~15u)) continue;
        if (w < win) { win = w; stage = (w + max_c + 16u + 15u) & ~15u; }
        G = g;
    }
    if (!G) return HB_OK;
    const size_t smem = table + (size_t)G * sizeof(uint32_t) * hb_emitf_group_words<8>(stage);
    const bool np3 = wf <= 10u;
    int rc = HB_OK;
    switch (G * 2 + (np3 ? 1 : 0)) {
    case 9: rc = run_emit_flat<4, 3>(ctx, ae, rshift, ntiles_run, d_out, out_capacity, win, stage, smem); break;
    case 8: rc = run_emit_flat<4, 2>(ctx, ae, rshift, ntiles_run, d_out, out_capacity, win, stage, smem); break;
    case 7: rc = run_emit_flat<3, 3>(ctx, ae, rshift, ntiles_run, d_out, out_capacity, win, stage, smem); break;
    case 6: rc = run_emit_flat<3, 2>(ctx, ae, rshift, ntiles_run, d_out, out_capacity, win, stage, smem); break;
    case 5: rc = run_emit_flat<2, 3>(ctx, ae, rshift, ntiles_run, d_out, out_capacity, win, stage, smem); break;
    case 4: rc = run_emit_flat<2, 2>(ctx, ae, rshift, ntiles_run, d_out, out_capacity, win, stage, smem); break;
    }
    if (rc == HB_OK) *launched = true;
    return rc;
}

__global__ void __launch_bounds__(256)
hb_build_e32_kernel(const uint32_t *__restrict__ lut, uint32_t w1, uint32_t wf, uint32_t *__restrict__ out) {
    const uint32_t x = blockIdx.x * 256u + threadIdx.x;
    if (x >> wf) return;
    const hb_lutref slow{lut, lut, (1u << w1) - 1u};
    out[x] = hb_e32_entry(slow, x, wf);
}

__global__ void __launch_bounds__(256)
hb_build_e64_kernel(const uint32_t *__restrict__ lut, uint32_t w1, uint32_t wf, uint32_t *__restrict__ out) {
    const uint32_t x = blockIdx.x * 256u + threadIdx.x;
    if (x >> wf) return;
    const hb_lutref slow{lut, lut, (1u << w1) - 1u};
    hb_e64_entry(slow, x, wf, out + 2 * x, out + 2 * x + 1);
}

static int e64_table(hb_ctx *ctx, const hb_codebook *cb, uint32_t wf, const uint32_t **out) {
    hb_codebook *m = const_cast<hb_codebook *>(cb);
    if (!m->d_e64[wf]) {
        uint32_t *p = nullptr;
        CK(cudaMallocAsync((void **)&p, sizeof(uint32_t) * 2 << wf, ctx->stream));
        hb_build_e64_kernel<<<((1u << wf) + 255u) / 256u, 256, 0, ctx->stream>>>(cb->d_lut, cb->lut.w1, wf, p);
        CK(cudaGetLastError());
        m->d_e64[wf] = p;
    }
    *out = m->d_e64[wf];
    return HB_OK;
}

/* the codebook's E32-table of index width wf (built on first use, on the context's stream) */
static int e32_table(hb_ctx *ctx, const hb_codebook *cb, uint32_t wf, const uint32_t **out) {
    hb_codebook *m = const_cast<hb_codebook *>(cb);     /* a cache: not part of the codebook's value */
    if (!m->d_e32[wf]) {
        uint32_t *p = nullptr;
        CK(cudaMallocAsync((void **)&p, sizeof(uint32_t) << wf, ctx->stream));
        hb_build_e32_kernel<<<((1u << wf) + 255u) / 256u, 256, 0, ctx->stream>>>(cb->d_lut, cb->lut.w1, wf, p);
        CK(cudaGetLastError());
        m->d_e32[wf] = p;
    }
    *out = m->d_e32[wf];
    return HB_OK;
}

template <int WPT>
static int launch_emit(hb_ctx *ctx, const hb_codebook *cb, const hb_stream_args &a,
                       const uint64_t *d_entry_base, void *d_out, uint64_t out_capacity) {
    const uint32_t ncta = ctx->map_ncta;
    uint64_t *misc = misc_words(ctx);
    int rc;
    if (phase_events(ctx, a.ntiles)) CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    uint32_t scan_launches = 2;
    if (ctx->map_fused) {
        /* small single-shard stream: up, top, down and fix in one launch */
        hb_scan_small_kernel<WPT><<<1, HB_SCAN_T, 0, ctx->stream>>>(
            a, (const uint32_t *)ctx->tmaps.p, d_entry_base, misc, (uint8_t *)ctx->tile_entry.p,
            (uint64_t *)ctx->tile_base.p, misc + 32, (uint16_t *)ctx->subs.p);
        CK(cudaGetLastError());
        scan_launches = 1;
        if (a.minlen == a.maxlen) {
            hb_fix_fixed_kernel<WPT><<<a.ntiles, HB_T, 0, ctx->stream>>>(
                a, (const uint8_t *)ctx->tile_entry.p, (uint16_t *)ctx->subs.p);
            CK(cudaGetLastError());
            scan_launches = 2;
        }
    } else {
        /* down-sweep; the owner of a tile whose true entry offset is not 0 re-chains its head */
        const size_t cm_bytes = ctx->map_notop ? (size_t)ncta * 32 * sizeof(uint64_t) : 0;
        if (cm_bytes)
            CK(cudaFuncSetAttribute(hb_scan_downfix_kernel<WPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    HB_NOTOP_MAX_CTAS * 32 * (int)sizeof(uint64_t)));
        hb_scan_downfix_kernel<WPT><<<ncta, HB_SCAN_T, cm_bytes, ctx->stream>>>(
            a, (const uint32_t *)ctx->tmaps.p, (const uint64_t *)ctx->wmaps.p,
            (const uint64_t *)ctx->cprefix.p, misc, d_entry_base, (uint8_t *)ctx->tile_entry.p,
            (uint64_t *)ctx->tile_base.p, misc + 32, (uint16_t *)ctx->subs.p,
            ctx->map_notop ? (const uint64_t *)ctx->cmaps.p : nullptr);
        CK(cudaGetLastError());
        scan_launches = 1;
        if (a.minlen == a.maxlen) {
            hb_fix_fixed_kernel<WPT><<<a.ntiles, HB_T, 0, ctx->stream>>>(
                a, (const uint8_t *)ctx->tile_entry.p, (uint16_t *)ctx->subs.p);
            CK(cudaGetLastError());
            scan_launches = 2;
        }
    }
    ctx->last_launches += scan_launches + 1;   /* + the emit kernel below */
    if (phase_events(ctx, a.ntiles)) CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    hb_stream_args ae = a;
    uint32_t win = 0, stage = 0;
    stage_geometry(cb, WPT, &win, &stage);
    size_t smem = emit_smem_bytes(a.wf, stage);
    int grid = 1;
    /* on request (A/B, tests): all tiles but the last one through the flat kernel.  Measured
     * slower than hb_emitw_kernel on both bench workloads (profiles/r02a_flat_emit.md), so
     * HB_EMIT_AUTO never picks it. */
    uint32_t tile0 = 0;
    if (WPT == 8 && ctx->emit_path == HB_EMIT_FLAT && a.ntiles >= 2) {
        bool launched = false;
        if ((rc = launch_emit_flat(ctx, cb, a, a.ntiles - 1u, d_out, out_capacity, &launched))) return rc;
        if (launched) { tile0 = a.ntiles - 1u; ctx->last_launches++; ctx->last_emit = "hb_emitf_kernel"; }
    }
    /* staging stores: whole words, three symbols per probe (english1g 0.75 ms vs 0.81 with
     * byte stores); the byte-store kernel on request */
    /* 32-bit table entries, three symbols per probe, 4 (or 8) copies on disjoint banks, the table
     * at a multiple of its size (hb_emit32_kernel): streams of at least four tiles per SM */
    bool done32 = false;
    /* warp-autonomous pipeline over the E64-table (four symbols per probe): on request, and by default for codes
     * whose implied mean length is at most 3.5 bits (three symbols do not fill a probe there) in streams of at
     * least four tiles per SM: fib4g emit 1.681 -> 1.568 ms with 11 bits x 4 copies (12 x 2: 1.605, 10 x 8: 1.594,
     * 11 x 2: 1.648); english1g would lose (0.677 against 0.557 ms) */
    if (WPT >= 2 && tile0 == 0 &&
        (ctx->emit_path == HB_EMIT_WORDS64W ||
         (ctx->emit_path == HB_EMIT_AUTO && ctx->auto_warp_emit && cb->lut.wf64 < HB_WF_MAX &&
          a.ntiles >= 4u * (uint32_t)ctx->prop.multiProcessorCount))) {
        const size_t limit = (size_t)ctx->prop.sharedMemPerBlockOptin;
        const uint32_t Sbits = 32u * (uint32_t)WPT;
        const uint32_t max_c = (Sbits + cb->lut.minlen - 1) / cb->lut.minlen;
        const double avg = cb->implied_avg_len > 1.0 ? cb->implied_avg_len : 1.0;
        uint32_t winw = (uint32_t)(32.0 * Sbits * 1.25 / avg) + 64u;
        if (winw > 32u * max_c) winw = 32u * max_c;
        if (winw < max_c) winw = max_c;
        winw = (winw + 15u) & ~15u;
        const uint32_t stg = (winw + max_c + 32u + 15u) & ~15u;
        uint32_t wfx = ctx->ep_wf >= 9 && ctx->ep_wf <= 14 ? (uint32_t)ctx->ep_wf : ctx->e64_wide;
        if (wfx > cb->lut.maxlen && cb->lut.maxlen >= 9u) wfx = cb->lut.maxlen;
        uint32_t rs = ctx->ep_rshift >= 0 ? (ctx->ep_rshift > 3 ? 3u : (uint32_t)ctx->ep_rshift) : 2u;
        while (rs > 0 && ((size_t)8 << (wfx + rs)) + 32u * (size_t)stg > limit) rs--;
        while (wfx > 9u && ((size_t)8 << (wfx + rs)) + 32u * (size_t)stg > limit) wfx--;
        const size_t tab = (size_t)8 << (wfx + rs), total = tab + 32u * (size_t)stg;
        /* unasked, only with a table of at least 10 bits beside the whole window; else the group kernels */
        if (total <= limit && (ctx->emit_path == HB_EMIT_WORDS64W || wfx >= 10u)) {
            ae.wf = wfx;
            if (wfx == cb->lut.wf64) ae.fast = a.fast + ((size_t)2 << a.wf);
            else if ((rc = e64_table(ctx, cb, wfx, &ae.fast))) return rc;
            const uint64_t units = (uint64_t)a.ntiles * (HB_T / 32u);
            uint64_t g = (units + 31) / 32;
            if (g > (uint64_t)ctx->prop.multiProcessorCount) g = (uint64_t)ctx->prop.multiProcessorCount;
            CK(cudaFuncSetAttribute(hb_emit32w_kernel<WPT, true, 1, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
            hb_emit32w_kernel<WPT, true, 1, true><<<(int)g, 1024, total, ctx->stream>>>(
                ae, rs, 0u, (uint32_t)tab, (const uint16_t *)ctx->subs.p, (const uint64_t *)ctx->tile_base.p,
                (const uint64_t *)(misc + 32), (uint8_t *)d_out, out_capacity, winw, stg, (uint32_t *)(misc + 36));
            done32 = true;
            ctx->last_emit = "hb_emit32w_kernel<E64>";
        }
    }
    const bool want32w = !done32 && (ctx->emit_path == HB_EMIT_WORDS32W ||
                         (ctx->emit_path == HB_EMIT_AUTO && ctx->auto_warp_emit &&
                          a.ntiles - tile0 >= 4u * (uint32_t)ctx->prop.multiProcessorCount));
    if (WPT >= 2 && want32w) {
        /* warp-autonomous variant: one staging slice per warp (32 subsequences' worth of output, 25 % head
         * room, one thread's overhang), the table in front of the slices */
        const size_t limit = (size_t)ctx->prop.sharedMemPerBlockOptin;
        const uint32_t per_sm = (a.ntiles - tile0) / (uint32_t)ctx->prop.multiProcessorCount;
        uint32_t wf_want = per_sm >= 32u ? 15u : (per_sm >= 8u ? 14u : 12u);
        if (ctx->ep_wf >= 9 && ctx->ep_wf <= HB_E32_WF_MAX) wf_want = (uint32_t)ctx->ep_wf;
        if (wf_want > cb->lut.maxlen && cb->lut.maxlen >= 9u) wf_want = cb->lut.maxlen;
        const uint32_t wf_min = ctx->emit_path == HB_EMIT_WORDS32W ? 9u : 12u;
        /* ctx->warp_spl subsequences per lane (A/B knob), else one.  For each: the widest table beside which
         * the whole window still fits (a second window per warp tile costs far more than an index bit: fib4g
         * 2.74 ms with 15 bits and two windows, 1.69 ms with 14 and one); on request (HB_EMIT_WORDS32W) a
         * smaller window with the narrowest table rather than no launch. */
        for (uint32_t spl = (uint32_t)ctx->warp_spl; spl >= 1u && !done32; spl--) {
            const uint32_t Sbits = 32u * (uint32_t)WPT * spl;      /* stream bits per lane */
            const uint32_t max_c = (Sbits + cb->lut.minlen - 1) / cb->lut.minlen;
            const double avg = cb->implied_avg_len > 1.0 ? cb->implied_avg_len : 1.0;
            uint32_t winw = (uint32_t)(32.0 * Sbits * 1.25 / avg) + 64u;
            if (winw > 32u * max_c) winw = 32u * max_c;
            if (winw < max_c) winw = max_c;
            winw = (winw + 15u) & ~15u;
            for (int pass = 0; pass < 2 && !done32; pass++)
            for (uint32_t wfx = wf_want; wfx >= wf_min && !done32; wfx--) {
                const size_t tab = (size_t)4 << wfx;
                size_t room = limit > tab ? limit - tab : 0;
                uint32_t ww = winw;
                if (pass == 1)      /* second pass: shrink the window, never below one thread's output */
                    while (ww > max_c + 16u && 32u * (size_t)((ww + max_c + 32u + 15u) & ~15u) > room) ww -= 16u;
                const uint32_t stg = (ww + max_c + 32u + 15u) & ~15u;
                if (ww < max_c || 32u * (size_t)stg > room) continue;
                if (pass == 1 && (ctx->emit_path != HB_EMIT_WORDS32W || spl > 1u)) break;   /* rather one subsequence per lane / the group kernels */
                ae.wf = wfx;
                if ((rc = e32_table(ctx, cb, wfx, &ae.fast))) return rc;
                const size_t total = tab + 32u * (size_t)stg;
                const uint64_t units = (uint64_t)(a.ntiles - tile0) * (HB_T / (32u * spl));
                uint64_t g = (units + 31) / 32;
                if (g > (uint64_t)ctx->prop.multiProcessorCount) g = (uint64_t)ctx->prop.multiProcessorCount;
#define HB_LAUNCH_EMIT32W(SPL)                                                                                      \
                do {                                                                                                \
                    CK(cudaFuncSetAttribute(hb_emit32w_kernel<WPT, true, SPL>,                                      \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));              \
                    hb_emit32w_kernel<WPT, true, SPL><<<(int)g, 1024, total, ctx->stream>>>(                        \
                        ae, 0u, 0u, (uint32_t)tab, (const uint16_t *)ctx->subs.p, (const uint64_t *)ctx->tile_base.p, \
                        (const uint64_t *)(misc + 32), (uint8_t *)d_out, out_capacity, ww, stg,                     \
                        (uint32_t *)(misc + 36));                                                                   \
                } while (0)
                if (spl == 2u) HB_LAUNCH_EMIT32W(2);
                else HB_LAUNCH_EMIT32W(1);
#undef HB_LAUNCH_EMIT32W
                done32 = true;
                ctx->last_emit = "hb_emit32w_kernel";
            }
        }
    }
    if (!done32 && WPT >= 2 && (ctx->emit_path == HB_EMIT_WORDS32 ||
                     (ctx->emit_path == HB_EMIT_AUTO && cb->lut.wf64 == HB_WF_MAX &&   /* short codes: four symbols per probe pay (fib4g 1.75 vs 1.85 ms) */
                      a.ntiles - tile0 >= 4u * (uint32_t)ctx->prop.multiProcessorCount))) {
        constexpr uint32_t G = 4;
        /* index width: the table is built once per codebook and width (e32_table) and every CTA copies
         * it from L2 (4 .. 128 KB).  English text: 2.3 symbols per probe with 12 bits, 2.65 with 14, 2.8
         * with 15, and a tenth of the long-codeword fallbacks (english1g emit 0.723 / 0.651 / 0.641 ms). */
        const uint32_t per_sm = (a.ntiles - tile0) / (uint32_t)ctx->prop.multiProcessorCount;
        uint32_t wf32 = per_sm >= 32u ? 15u : (per_sm >= 8u ? 14u : 12u);
        if (ctx->ep_wf >= 9 && ctx->ep_wf <= HB_E32_WF_MAX) wf32 = (uint32_t)ctx->ep_wf;
        if (wf32 > cb->lut.maxlen && cb->lut.maxlen >= 9u) wf32 = cb->lut.maxlen;   /* no longer codeword exists */
        uint32_t rs = ctx->ep_rshift >= 0 ? (ctx->ep_rshift > 3 ? 3u : (uint32_t)ctx->ep_rshift) : (wf32 >= 14u ? 0u : (wf32 == 13u ? 1u : 2u));   /* copies: what fits 64 KB (no measurable effect: not bound by replays) */
        const size_t limit = (size_t)ctx->prop.sharedMemPerBlockOptin;
        const size_t grp = sizeof(uint32_t) * (size_t)hb_emitw_group_words(stage);
        for (;; rs--) {
            const size_t tab = (size_t)4 << (wf32 + rs);
            /* aligned to its size when a multiple of it lies inside the window (up to 64 KB); else at offset 0 */
            const size_t abs0 = ((size_t)ctx->smem_base + tab - 1) / tab * tab;
            const bool add = abs0 + tab > (size_t)ctx->smem_base + limit;
            const size_t tab_off = add ? 0 : abs0 - ctx->smem_base;
            size_t n_before = tab_off / grp;
            if (n_before > G) n_before = G;
            const size_t total = tab_off + tab + (G - n_before) * grp;
            if (total <= limit) {
                ae.wf = wf32;
                if ((rc = e32_table(ctx, cb, wf32, &ae.fast))) return rc;
                const uint32_t need = (a.ntiles - tile0 + G - 1u) / G;
#define HB_LAUNCH_EMIT32(ADD)                                                                                  \
                do {                                                                                           \
                    if ((rc = grid_for(ctx, hb_emit32_kernel<WPT, G, ADD>, total, need, &grid, G * HB_T))) return rc;   \
                    hb_emit32_kernel<WPT, G, ADD><<<grid, G * HB_T, total, ctx->stream>>>(                     \
                        ae, tile0, rs, (uint32_t)tab_off, (uint32_t)n_before, (const uint16_t *)ctx->subs.p,  \
                        (const uint64_t *)ctx->tile_base.p, (const uint64_t *)(misc + 32), (uint8_t *)d_out,   \
                        out_capacity, win, stage, (uint32_t *)(misc + 36));                                    \
                } while (0)
                if (add) HB_LAUNCH_EMIT32(true);
                else HB_LAUNCH_EMIT32(false);
#undef HB_LAUNCH_EMIT32
                done32 = true;
                ctx->last_emit = "hb_emit32_kernel";
                break;
            }
            if (rs == 0) break;
        }
    }
    if (done32) {
    } else if (ctx->emit_path != HB_EMIT_BYTES) {
        ae.fast = a.fast + ((size_t)2 << a.wf);   /* E64-table */
        ae.wf = cb->lut.wf64;                     /* ... and its own index width */
        /* another width on request (hb_ctx_set_emit_table), or 13 bits for short codes in large streams:
         * 3.6 instead of 3.2 symbols per probe on the Fibonacci-skewed model */
        uint32_t wfx = ae.wf;
        if (ctx->ep_wf >= 9 && ctx->ep_wf <= 14) wfx = (uint32_t)ctx->ep_wf;
        else if (ctx->emit_path == HB_EMIT_AUTO && cb->lut.wf64 < HB_WF_MAX &&
                 a.ntiles - tile0 >= 32u * (uint32_t)ctx->prop.multiProcessorCount) wfx = ctx->e64_wide;
        if (wfx > cb->lut.maxlen && cb->lut.maxlen >= 9u) wfx = cb->lut.maxlen;
        if (wfx != ae.wf) {
            if ((rc = e64_table(ctx, cb, wfx, &ae.fast))) return rc;
            ae.wf = wfx;
        }
        /* G groups of 256 threads share R copies of the table (hb_tables64) */
        const uint32_t left = a.ntiles - tile0;
        const uint32_t grp = sizeof(uint32_t) * hb_emitw_group_words(stage);
        const size_t optin = (size_t)ctx->prop.sharedMemPerBlockOptin;
        /* groups: 4 (one 1024-thread CTA per SM) for streams of at least four tiles per SM, else 1;
         * copies: as requested, else 4 (measured: english1g emit 0.775 -> 0.747 ms, fib4g 1.85 -> 1.75;
         * 2 copies: no change) -- fewer of either when table + groups would not fit an SM */
        uint32_t G = left >= 4u * (uint32_t)ctx->prop.multiProcessorCount || ctx->ep_rshift > 0 ? 4u : 1u;
        uint32_t rs = ctx->ep_rshift >= 0 ? (ctx->ep_rshift > 3 ? 3u : (uint32_t)ctx->ep_rshift) : (G == 4u ? 2u : 0u);
        while (rs > 0 && ((size_t)8 << (ae.wf + rs)) + (size_t)G * grp > optin) rs--;
        while (G > 1 && ((size_t)8 << (ae.wf + rs)) + (size_t)G * grp > optin) G >>= 1;
        if (((size_t)8 << ae.wf) + (size_t)grp > optin && ae.wf != cb->lut.wf64) {
            /* a wide table beside a large staging window (very short codes): the codebook's own width */
            ae.fast = a.fast + ((size_t)2 << a.wf);
            ae.wf = cb->lut.wf64;
        }
        smem = ((size_t)8 << (ae.wf + rs)) + (size_t)G * grp;   /* the table sits in front of the groups */
        const uint32_t need = (left + G - 1u) / G;
#define HB_LAUNCH_EMITW(GG)                                                                              \
        do {                                                                                             \
            if ((rc = grid_for(ctx, hb_emitw_kernel<WPT, GG>, smem, need, &grid, GG * HB_T))) return rc;   \
            hb_emitw_kernel<WPT, GG><<<grid, GG * HB_T, smem, ctx->stream>>>(                             \
                ae, tile0, rs, (const uint16_t *)ctx->subs.p, (const uint64_t *)ctx->tile_base.p,         \
                (const uint64_t *)(misc + 32), (uint8_t *)d_out, out_capacity, win, stage,               \
                (uint32_t *)(misc + 36));                                                                \
        } while (0)
        if (G == 4) HB_LAUNCH_EMITW(4);
        else if (G == 2) HB_LAUNCH_EMITW(2);
        else HB_LAUNCH_EMITW(1);
#undef HB_LAUNCH_EMITW
        if (!tile0) ctx->last_emit = "hb_emitw_kernel";
    } else {
        ae.fast = a.fast + ((size_t)1 << a.wf);   /* E-table */
        ctx->last_emit = "hb_emit_kernel";
        if ((rc = grid_for(ctx, hb_emit_kernel<WPT>, smem, a.ntiles, &grid))) return rc;
        hb_emit_kernel<WPT><<<grid, HB_T, smem, ctx->stream>>>(
            ae, (const uint16_t *)ctx->subs.p, (const uint64_t *)ctx->tile_base.p,
            (const uint64_t *)(misc + 32), (uint8_t *)d_out, out_capacity, win,
            (uint32_t *)(misc + 36));
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    return HB_OK;
}

static int prepare_misc(hb_ctx *ctx) {
    int rc = ensure(ctx, ctx->misc, 64 * sizeof(uint64_t));
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->misc.p, 0, 64 * sizeof(uint64_t), ctx->stream));
    return HB_OK;
}

extern "C" int hb_shard_map(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp,
                            uint64_t comp_bytes, uint64_t bits_own, uint64_t bits_avail,
                            uint64_t *d_map) {
    hb_stream_args a;
    int rc = make_args(ctx, cb, d_comp, comp_bytes, bits_own, bits_avail, &a);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = prepare_misc(ctx))) return rc;
    ctx->have_map = false;
    /* per-step event set: a slot of the timing ring while one is armed */
    if (ctx->tim_ev && ctx->tim_n < ctx->tim_cap) ctx->ev = ctx->tim_ev + (size_t)ctx->tim_n++ * HB_NEV;
    else ctx->ev = ctx->ev0;
    if (a.ntiles == 0) {
        /* empty shard: identity map (entry e -> exit e, 0 symbols) */
        uint64_t ident[32];
        for (int e = 0; e < 32; e++) ident[e] = (uint64_t)e;
        CK(cudaMemcpyAsync(ctx->misc.p, ident, sizeof(ident), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (d_map)
            CK(cudaMemcpyAsync(d_map, ctx->misc.p, sizeof(ident), cudaMemcpyDeviceToDevice, ctx->stream));
        for (int i = 0; i < HB_NEV; i++) CK(cudaEventRecord(ctx->ev[i], ctx->stream));
        ctx->have_map = true;
        ctx->map_ntiles = 0;
        ctx->map_ncta = 0;
        ctx->map_wpt = ctx->wpt;
        return HB_OK;
    }
    switch (ctx->wpt) {
    case 4: return launch_map<4>(ctx, cb, a, d_map);
    case 8: return launch_map<8>(ctx, cb, a, d_map);
    case 16: return launch_map<16>(ctx, cb, a, d_map);
    }
    return HB_ERR_ARG;
}

extern "C" int hb_shard_compose(hb_ctx *ctx, const uint64_t *d_all_maps, int n_ranks, int rank,
                                uint64_t *d_entry_base) {
    if (!ctx || !d_all_maps || !d_entry_base || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    hb_compose_kernel<<<1, 32, 0, ctx->stream>>>(d_all_maps, n_ranks, rank, d_entry_base);
    CK(cudaGetLastError());
    return HB_OK;
}

/* ---- map exchange without NCCL: one process per GPU, peer stores over NVLink ------------------
 * Every rank owns an exchange table in its own memory: HB_PEER_DEPTH ring slots x HB_MULTI_MAX
 * source ranks x (32 map words + a sequence flag).  hb_exchange_kernel, launched behind the scan of
 * hb_shard_map, (1) stores this rank's map into slot seq % DEPTH of the tables of the ranks to its
 * RIGHT (warp w serves rank + 1 + w; relaxed system-scope stores, a system fence, then the flag with
 * release semantics) and (2) has one thread wait for the flags of the ranks to its LEFT in its own
 * table and compose their maps into (entry offset, output base) -- what NCCL's all-gather plus
 * hb_compose_kernel did in two launches and ~30 us.  Rank 0 waits for nobody and every rank stores
 * before it waits, so the chain cannot deadlock; a rank that never shows up is reported after 5 s
 * (HB_ST_PEER_TIMEOUT -> HB_ERR_STATE) instead of hanging the device.  The ring lets a rank run up
 * to HB_PEER_DEPTH - 1 decodes ahead of its neighbours without overwriting a map not yet read. */
#define HB_PEER_DEPTH 1024u
#define HB_PEER_SLOT 40u
#define HB_ST_PEER_TIMEOUT 4u
struct hb_peer_tabs { uint64_t *p[HB_MULTI_MAX]; };

__global__ void __launch_bounds__(256)
hb_exchange_kernel(const uint64_t *__restrict__ my_map, hb_peer_tabs tabs, const uint64_t *own_tab, int rank,
                   int n_ranks, uint64_t seq, uint64_t *__restrict__ entry_base, uint32_t *__restrict__ status) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const size_t ring = (size_t)(seq % HB_PEER_DEPTH) * HB_MULTI_MAX * HB_PEER_SLOT;
    const int dst = rank + 1 + w;
    if (dst < n_ranks) {
        uint64_t *t = tabs.p[dst] + ring + (size_t)rank * HB_PEER_SLOT;
        const uint64_t v = my_map[l];
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(t + l), "l"(v) : "memory");
        __threadfence_system();
        __syncwarp();
        if (l == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(t + 32), "l"(seq) : "memory");
    }
    if (threadIdx.x == 7 * 32) {     /* at most seven peers: warp 7 never stores */
        uint32_t cur = 0;
        uint64_t base = 0, t0, t1;
        bool ok = true;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int r = 0; r < rank && ok; r++) {
            const uint64_t *sl = own_tab + ring + (size_t)r * HB_PEER_SLOT;
            for (;;) {
                uint64_t f;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(sl + 32) : "memory");
                if (f == seq) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 5000000000ull) { ok = false; break; }
            }
            if (!ok) break;
            uint64_t m;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(m) : "l"(sl + cur) : "memory");
            base += m >> 8;
            cur = (uint32_t)m & 31u;
        }
        if (!ok) { atomicOr(status, HB_ST_PEER_TIMEOUT); cur = 0; base = 0; }
        entry_base[0] = cur;
        entry_base[1] = base;
        entry_base[2] = base + (my_map[cur] >> 8);
    }
}

/* unmap the other ranks' tables (this rank's own table stays: others may still store into it) */
extern "C" int hb_peer_disconnect(hb_ctx *ctx) {
    if (!ctx) return HB_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->peer_n) cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < HB_MULTI_MAX; r++)
        if (ctx->peer_ptr[r]) {
            if (ctx->peer_ipc) cudaIpcCloseMemHandle(ctx->peer_ptr[r]);
            ctx->peer_ptr[r] = nullptr;
        }
    ctx->peer_rank = -1;
    ctx->peer_n = 0;
    cudaGetLastError();
    return HB_OK;
}

extern "C" int hb_peer_close(hb_ctx *ctx) {
    if (!ctx) return HB_ERR_ARG;
    hb_peer_disconnect(ctx);
    if (ctx->peer_tab) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->peer_tab); ctx->peer_tab = nullptr; }
    cudaGetLastError();
    return HB_OK;
}

static int peer_table(hb_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->peer_tab) {
        const size_t bytes = sizeof(uint64_t) * HB_PEER_DEPTH * HB_MULTI_MAX * HB_PEER_SLOT;
        CK(cudaMalloc((void **)&ctx->peer_tab, bytes));
        CK(cudaMemset(ctx->peer_tab, 0, bytes));
        CK(cudaDeviceSynchronize());
    }
    return HB_OK;
}

extern "C" int hb_peer_export(hb_ctx *ctx, void *handle) {
    if (!ctx || !handle) return HB_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == HB_PEER_HANDLE_BYTES, "IPC handle size");
    int rc = peer_table(ctx);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->peer_tab));
    memcpy(handle, &h, sizeof(h));
    return HB_OK;
}

extern "C" int hb_peer_connect(hb_ctx *ctx, int rank, int n_ranks, const void *handles) {
    if (!ctx || !handles || n_ranks < 1 || n_ranks > HB_MULTI_MAX || rank < 0 || rank >= n_ranks) return HB_ERR_ARG;
    if (!ctx->peer_tab) return HB_ERR_STATE;       /* hb_peer_export first */
    CK(cudaSetDevice(ctx->device));
    for (int r = rank + 1; r < n_ranks; r++) {
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)r * HB_PEER_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_ptr[r] = (uint64_t *)p;
    }
    ctx->peer_rank = rank;
    ctx->peer_n = n_ranks;
    ctx->peer_ipc = true;
    return HB_OK;
}

/* the same for contexts of ONE process (several devices, or several contexts on one device: tests) */
extern "C" int hb_peer_connect_local(hb_ctx *ctx, int rank, int n_ranks, hb_ctx *const *ctxs) {
    if (!ctx || !ctxs || n_ranks < 1 || n_ranks > HB_MULTI_MAX || rank < 0 || rank >= n_ranks || ctxs[rank] != ctx)
        return HB_ERR_ARG;
    int rc = peer_table(ctx);
    if (rc) return rc;
    for (int r = rank + 1; r < n_ranks; r++) {
        if (!ctxs[r]) return HB_ERR_ARG;
        if ((rc = peer_table(ctxs[r]))) return rc;
        if (ctxs[r]->device != ctx->device) {
            CK(cudaSetDevice(ctx->device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[r]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        ctx->peer_ptr[r] = ctxs[r]->peer_tab;
    }
    ctx->peer_rank = rank;
    ctx->peer_n = n_ranks;
    ctx->peer_ipc = false;
    return HB_OK;
}

extern "C" int hb_shard_exchange(hb_ctx *ctx, uint64_t seq, uint64_t *d_entry_base) {
    if (!ctx || !d_entry_base || seq == 0) return HB_ERR_ARG;
    if (!ctx->have_map || ctx->peer_rank < 0) return HB_ERR_STATE;
    CK(cudaSetDevice(ctx->device));
    hb_peer_tabs tabs;
    for (int r = 0; r < HB_MULTI_MAX; r++) tabs.p[r] = ctx->peer_ptr[r];
    uint64_t *misc = misc_words(ctx);
    hb_exchange_kernel<<<1, 256, 0, ctx->stream>>>(misc, tabs, ctx->peer_tab, ctx->peer_rank, ctx->peer_n, seq,
                                                   d_entry_base, (uint32_t *)(misc + 36));
    CK(cudaGetLastError());
    return HB_OK;
}

static int finish_result(hb_ctx *ctx, uint32_t ntiles, uint32_t launches, hb_result *res) {
    uint64_t *misc = misc_words(ctx);
    CK(cudaMemcpyAsync(ctx->h_res, misc + 32, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memset(res, 0, sizeof(*res));
    res->n_symbols = ctx->h_res[0];
    res->exit_offset = (uint32_t)ctx->h_res[1];
    res->entry_offset = (uint32_t)ctx->h_res[2];
    res->out_base = ctx->h_res[3];
    res->launches = launches;
    res->tiles = ntiles;
    float ms = 0;
    if (phase_events(ctx, ntiles)) {
        if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) res->ms_sync = ms;
        float s1 = 0, s2 = 0;
        cudaEventElapsedTime(&s1, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&s2, ctx->ev[2], ctx->ev[3]);
        res->ms_scan = s1 + s2;
        if (cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]) == cudaSuccess) res->ms_emit = ms;
    }
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]) == cudaSuccess) res->ms_total = ms;
    cudaGetLastError();
    if ((uint32_t)ctx->h_res[4] & HB_ST_LAYOUT) {
        snprintf(ctx->err, sizeof(ctx->err), "hb_emit32_kernel: table not aligned to its size (dynamic shared memory base moved)");
        return HB_ERR_CUDA;
    }
    if ((uint32_t)ctx->h_res[4] & HB_ST_PEER_TIMEOUT) {
        snprintf(ctx->err, sizeof(ctx->err), "hb_shard_exchange: a rank to the left never delivered its map (5 s)");
        return HB_ERR_STATE;
    }
    if ((uint32_t)ctx->h_res[4] & HB_ST_OUTPUT_FULL) return HB_ERR_OUTPUT_FULL;
    return HB_OK;
}

extern "C" int hb_shard_emit(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp,
                             uint64_t comp_bytes, uint64_t bits_own, uint64_t bits_avail,
                             const uint64_t *d_entry_base, void *d_out, uint64_t out_capacity,
                             hb_result *res) {
    hb_stream_args a;
    int rc = make_args(ctx, cb, d_comp, comp_bytes, bits_own, bits_avail, &a);
    if (rc) return rc;
    if (!ctx->have_map || ctx->map_ntiles != a.ntiles || ctx->map_wpt != ctx->wpt)
        return HB_ERR_STATE;
    if (!d_out && out_capacity) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t launches = 0;
    if (a.ntiles == 0) {
        /* result = entry/base passthrough, zero symbols */
        if (res) {
            uint64_t eb[2] = {0, 0};
            if (d_entry_base) {
                CK(cudaMemcpyAsync(ctx->h_res, d_entry_base, 2 * sizeof(uint64_t),
                                   cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                eb[0] = ctx->h_res[0]; eb[1] = ctx->h_res[1];
            }
            memset(res, 0, sizeof(*res));
            res->entry_offset = res->exit_offset = (uint32_t)eb[0];
            res->out_base = eb[1];
        }
        return HB_OK;
    }
    switch (ctx->wpt) {
    case 4: rc = launch_emit<4>(ctx, cb, a, d_entry_base, d_out, out_capacity); break;
    case 8: rc = launch_emit<8>(ctx, cb, a, d_entry_base, d_out, out_capacity); break;
    case 16: rc = launch_emit<16>(ctx, cb, a, d_entry_base, d_out, out_capacity); break;
    default: rc = HB_ERR_ARG;
    }
    if (rc) return rc;
    launches = ctx->last_launches;
    if (res) return finish_result(ctx, a.ntiles, launches, res);
    return HB_OK;
}

extern "C" int hb_shard_result(hb_ctx *ctx, hb_result *res) {
    if (!ctx || !res) return HB_ERR_ARG;
    if (!ctx->have_map) return HB_ERR_STATE;
    CK(cudaSetDevice(ctx->device));
    if (ctx->map_ntiles == 0) {
        CK(cudaStreamSynchronize(ctx->stream));
        memset(res, 0, sizeof(*res));
        return HB_OK;
    }
    return finish_result(ctx, ctx->map_ntiles, ctx->last_launches, res);
}

extern "C" int hb_host_pin(const void *ptr, uint64_t bytes) {
    if (!ptr || !bytes) return HB_ERR_ARG;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) return HB_OK;
    cudaGetLastError();
    cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return HB_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return HB_ERR_CUDA; }
    return HB_OK;
}

extern "C" int hb_host_unpin(const void *ptr) {
    if (!ptr) return HB_ERR_ARG;
    cudaError_t e = cudaHostUnregister(const_cast<void *>(ptr));
    cudaGetLastError();
    return e == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_decode_device(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp,
                                uint64_t comp_bytes, uint64_t bits, void *d_out,
                                uint64_t out_capacity, hb_result *res) {
    if (!ctx) return HB_ERR_ARG;   /* res == NULL: nothing is read back, the call returns with the kernels queued */
    ctx->fuse_small = true;    /* nobody reads the shard map between the two phases */
    const bool ok0 = ctx->origin_known;
    const uint64_t ob0 = ctx->origin_byte;
    ctx->origin_known = true;  /* a whole stream: its first byte is byte 0 */
    ctx->origin_byte = 0;
    int rc = hb_shard_map(ctx, cb, d_comp, comp_bytes, bits, bits, nullptr);
    ctx->fuse_small = false;
    if (rc == HB_OK)
        rc = hb_shard_emit(ctx, cb, d_comp, comp_bytes, bits, bits, nullptr, d_out, out_capacity, res);
    ctx->origin_known = ok0;
    ctx->origin_byte = ob0;
    return rc;
}

/* Large host-resident streams: the compressed bytes are cut into chunks (byte-range
 * shards on ONE GPU).  All uploads are queued on a copy stream; chunk k is mapped
 * as soon as it has landed, its 32-entry map is read back (256 B, the only host
 * synchronisation per chunk), the host composes the entry offset / output base,
 * the chunk is emitted and its bytes are downloaded on a second copy stream --
 * so upload k+1, decode k and download k-1 overlap on the two PCIe directions. */
static int decode_host_pipelined_run(hb_ctx *ctx, hb_codebook *cb, const uint8_t *data, uint64_t bits,
                                     uint8_t *out, uint64_t out_capacity, hb_result *res);

static int decode_host_pipelined(hb_ctx *ctx, hb_codebook *cb, const uint8_t *data, uint64_t bits,
                                 uint8_t *out, uint64_t out_capacity, hb_result *res) {
    const bool ok0 = ctx->origin_known;
    const uint64_t ob0 = ctx->origin_byte;
    int rc = decode_host_pipelined_run(ctx, cb, data, bits, out, out_capacity, res);
    ctx->origin_known = ok0;
    ctx->origin_byte = ob0;
    if (rc != HB_OK) {
        /* copies may still be in flight to/from the caller's buffers: drain them
         * before the error is reported */
        if (ctx->s_h2d) cudaStreamSynchronize(ctx->s_h2d);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
        cudaGetLastError();
    }
    return rc;
}

static int decode_host_pipelined_run(hb_ctx *ctx, hb_codebook *cb, const uint8_t *data, uint64_t bits,
                                     uint8_t *out, uint64_t out_capacity, hb_result *res) {
    const uint64_t nbytes = (bits + 7) / 8;
    const uint64_t tile_bytes = (uint64_t)HB_T * 4u * (uint64_t)ctx->wpt;
    /* a chunk costs ~50 us of its own (kernel ramps, one host round trip for its map), the first upload and
     * the last download are not overlapped: english1g end to end 27.1 / 25.2 / 23.6 / 23.1 ms with chunks of
     * 8 / 16 / 32 / 64 MiB (profiles/r02i_host_chunk_sweep.txt) */
    uint64_t chunk = ctx->pipe_chunk_bytes;
    if (ctx->pipe_chunk_default && nbytes >= (512ull << 20)) chunk = 64ull << 20;
    uint64_t cbytes = chunk / tile_bytes * tile_bytes;
    if (cbytes < tile_bytes) cbytes = tile_bytes;
    const int K = (int)((nbytes + cbytes - 1) / cbytes);
    int rc;
    if (!ctx->s_h2d) CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    if (!ctx->s_d2h) CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    if (!ctx->pipe_t0) { CK(cudaEventCreate(&ctx->pipe_t0)); CK(cudaEventCreate(&ctx->pipe_t1)); }
    if (K > ctx->pipe_cap) {
        /* grow events and pinned words together; pipe_cap only moves once both exist */
        cudaEvent_t *ne = (cudaEvent_t *)realloc(ctx->pipe_ev, sizeof(cudaEvent_t) * 2 * (size_t)K);
        if (!ne) return HB_ERR_NOMEM;
        ctx->pipe_ev = ne;
        uint64_t *nh = nullptr;
        cudaError_t e = cudaMallocHost((void **)&nh, sizeof(uint64_t) * 36 * (size_t)K);
        int made = 2 * ctx->pipe_cap;
        for (; e == cudaSuccess && made < 2 * K; made++)
            e = cudaEventCreateWithFlags(&ctx->pipe_ev[made], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            for (int i = 2 * ctx->pipe_cap; i < made - 1; i++) cudaEventDestroy(ctx->pipe_ev[i]);
            if (nh) cudaFreeHost(nh);
            return cuda_fail(ctx, e, "pipeline setup");
        }
        if (ctx->h_pipe) cudaFreeHost(ctx->h_pipe);
        ctx->h_pipe = nh;
        ctx->pipe_cap = K;
    }
    if ((rc = ensure(ctx, ctx->d_eb, sizeof(uint64_t) * 4 * (size_t)K))) return rc;
    const uint64_t padded = (nbytes + 15) & ~15ull;
    if ((rc = ensure(ctx, ctx->d_comp, padded + 32))) return rc;
    if ((rc = ensure(ctx, ctx->d_out, out_capacity + 16))) return rc;
    uint8_t *d_comp = (uint8_t *)ctx->d_comp.p, *d_out = (uint8_t *)ctx->d_out.p;

    /* the bytes behind the stream are zero before ANY chunk can read them (a chunk's halo may
     * reach past the end when the last chunk is shorter than 16 bytes), then every upload
     * (chunk + 16-byte halo), one event each */
    CK(cudaMemsetAsync(d_comp + nbytes, 0, padded + 32 - nbytes, ctx->s_h2d));
    for (int k = 0; k < K; k++) {
        const uint64_t a = (uint64_t)k * cbytes;
        const uint64_t b = a + cbytes + 16 < nbytes ? a + cbytes + 16 : nbytes;
        CK(cudaMemcpyAsync(d_comp + a, data + a, b - a, cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaEventRecord(ctx->pipe_ev[2 * k], ctx->s_h2d));
    }
    CK(cudaEventRecord(ctx->pipe_t0, ctx->stream));
    uint32_t cur = 0, launches = 0, tiles = 0;
    uint64_t base = 0;
    for (int k = 0; k < K; k++) {
        const uint64_t a = (uint64_t)k * cbytes;
        const bool last = k == K - 1;
        const uint64_t own = last ? bits - 8 * a : 8 * cbytes;
        /* never more valid bits than the stream has: the halo of the last-but-one chunk ends
         * where the data ends */
        const uint64_t avail = last ? own : (bits - 8 * a < 8 * (cbytes + 16) ? bits - 8 * a : 8 * (cbytes + 16));
        const uint64_t readable = last || a + cbytes + 16 > padded + 32 ? padded + 32 - a : cbytes + 16;
        uint64_t *h = ctx->h_pipe + 36 * (size_t)k;
        CK(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 * k], 0));
        ctx->origin_known = true;    /* chunk k begins at byte a of the stream */
        ctx->origin_byte = a;
        if ((rc = hb_shard_map(ctx, cb, d_comp + a, readable, own, avail, nullptr))) return rc;
        CK(cudaMemcpyAsync(h, ctx->misc.p, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const uint64_t m = h[cur];
        uint64_t n = m >> 8;
        if (last && n && own + (m & 31u) > avail) n--;   /* cut-off last codeword emits nothing */
        if (base + n > out_capacity) return HB_ERR_OUTPUT_FULL;
        h[32] = cur; h[33] = base; h[34] = 0; h[35] = 0;
        uint64_t *d_eb = (uint64_t *)ctx->d_eb.p + 4 * (size_t)k;
        CK(cudaMemcpyAsync(d_eb, h + 32, 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = hb_shard_emit(ctx, cb, d_comp + a, readable, own, avail, d_eb, d_out + base,
                                out_capacity - base, nullptr))) return rc;
        CK(cudaEventRecord(ctx->pipe_ev[2 * k + 1], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->pipe_ev[2 * k + 1], 0));
        if (n) CK(cudaMemcpyAsync(out + base, d_out + base, n, cudaMemcpyDeviceToHost, ctx->s_d2h));
        launches += ctx->last_launches;
        tiles += ctx->map_ntiles;
        base += n;
        cur = (uint32_t)(m & 31u);
    }
    CK(cudaEventRecord(ctx->pipe_t1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->s_d2h));
    memset(res, 0, sizeof(*res));
    res->n_symbols = base;
    res->exit_offset = cur;
    res->launches = launches;
    res->tiles = tiles;
    float ms = 0;
    /* first map to last emit on the compute stream (includes waits for uploads) */
    if (cudaEventElapsedTime(&ms, ctx->pipe_t0, ctx->pipe_t1) == cudaSuccess) res->ms_total = ms;
    /* output-full raised inside a kernel */
    CK(cudaMemcpy(ctx->h_res, misc_words(ctx) + 36, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if ((uint32_t)ctx->h_res[0] & HB_ST_LAYOUT) return HB_ERR_CUDA;
    if ((uint32_t)ctx->h_res[0] & HB_ST_OUTPUT_FULL) return HB_ERR_OUTPUT_FULL;
    return HB_OK;
}

extern "C" int hb_decode_host(hb_ctx *ctx, const hb_node_abi *tree, int nodes,
                              const uint8_t *data, uint64_t bits, uint8_t *out,
                              uint64_t out_capacity, hb_result *res) {
    if (!ctx || !tree || (!data && bits) || (!out && out_capacity)) return HB_ERR_ARG;
    hb_result local;
    if (!res) res = &local;
    CK(cudaSetDevice(ctx->device));
    int rc = HB_OK;
    {   /* reuse the cached code table when the tree is the same, field by field */
        bool same = ctx->host_cb && ctx->host_nodes == nodes;
        for (int i = 0; same && i < nodes; i++)
            same = tree[i].sym == ctx->host_tree[i].sym && tree[i].izero == ctx->host_tree[i].izero &&
                   tree[i].ione == ctx->host_tree[i].ione;
        if (!same) {
            if (ctx->host_cb) { hb_codebook_destroy(ctx->host_cb); ctx->host_cb = nullptr; }
            free(ctx->host_tree);
            ctx->host_tree = nullptr;
            ctx->host_nodes = 0;
            rc = hb_codebook_create(ctx, tree, nodes, &ctx->host_cb);
            if (rc) return rc;
            ctx->host_tree = (hb_node_abi *)malloc(sizeof(hb_node_abi) * (size_t)nodes);
            if (!ctx->host_tree) { hb_codebook_destroy(ctx->host_cb); ctx->host_cb = nullptr; return HB_ERR_NOMEM; }
            for (int i = 0; i < nodes; i++) {
                ctx->host_tree[i].sym = tree[i].sym;
                ctx->host_tree[i].izero = tree[i].izero;
                ctx->host_tree[i].ione = tree[i].ione;
            }
            ctx->host_nodes = nodes;
        }
    }
    hb_codebook *cb = ctx->host_cb;
    const uint64_t nbytes = (bits + 7) / 8;
    const uint64_t padded = (nbytes + 15) & ~15ull;
    if (nbytes >= 2 * ctx->pipe_chunk_bytes)
        return decode_host_pipelined(ctx, cb, data, bits, out, out_capacity, res);
    do {
        if ((rc = ensure(ctx, ctx->d_comp, padded + 16))) break;
        if ((rc = ensure(ctx, ctx->d_out, out_capacity + 16))) break;
        cudaError_t e = cudaSuccess;
        if (nbytes) e = cudaMemcpyAsync(ctx->d_comp.p, data, nbytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemsetAsync((uint8_t *)ctx->d_comp.p + nbytes, 0, padded + 16 - nbytes, ctx->stream);
        if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "upload"); break; }
        rc = hb_decode_device(ctx, cb, ctx->d_comp.p, padded + 16, bits, ctx->d_out.p, out_capacity, res);
        if (rc) break;
        if (res->n_symbols) {
            e = cudaMemcpyAsync(out, ctx->d_out.p, res->n_symbols, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "download"); break; }
        }
    } while (0);
    return rc;
}

/* ---- onethread: the whole stream on ONE device thread (debug aid) ------------------
 * The reference registers a `<<<1,1>>>` serial tree walk as its "onethread" approach
 * (framework/onethread.cu:13-52; the loop of simpleDecode, framework/mainrun.c:38-55).
 * Same thing here over the uploaded node array: a device-side statement of the stream
 * semantics that shares nothing with the table-driven kernels -- useful to tell a table
 * bug from a data bug.  Not a fast path. */
__global__ void hb_onethread_kernel(const hb_node_abi *__restrict__ tree, const uint8_t *__restrict__ data,
                                    uint64_t bits, uint8_t *__restrict__ out, uint64_t out_capacity,
                                    uint64_t *__restrict__ n_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int32_t node = 0;
    uint64_t n = 0;
    bool full = false;
    for (uint64_t pos = 0; pos < bits; pos++) {
        const uint32_t bit = (data[pos >> 3] >> (pos & 7)) & 1u;
        node = bit ? tree[node].ione : tree[node].izero;
        if (tree[node].izero == -1 && tree[node].ione == -1) {
            if (n < out_capacity) out[n] = tree[node].sym; else full = true;
            n++;
            node = 0;
        }
    }
    n_out[0] = n;
    n_out[1] = full ? 1u : 0u;
}

extern "C" int hb_decode_onethread(hb_ctx *ctx, const hb_node_abi *tree, int nodes, const uint8_t *data,
                                   uint64_t bits, uint8_t *out, uint64_t out_capacity, hb_result *res) {
    if (!ctx || !tree || nodes < 1 || (!data && bits) || (!out && out_capacity)) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < nodes; i++) {   /* the walk indexes the array with these: keep them inside it */
        const int32_t z = tree[i].izero, o = tree[i].ione;
        if ((z == -1) != (o == -1) || z < -1 || o < -1 || z >= nodes || o >= nodes) return HB_ERR_TREE;
    }
    if (tree[0].izero == -1) return HB_ERR_TREE;
    const uint64_t nbytes = (bits + 7) / 8;
    int rc;
    if ((rc = ensure(ctx, ctx->d_comp, nbytes + 16))) return rc;
    if ((rc = ensure(ctx, ctx->d_out, out_capacity + 16))) return rc;
    hb_node_abi *d_tree = nullptr;
    uint64_t *d_n = nullptr;
    CK(cudaMallocAsync((void **)&d_tree, sizeof(hb_node_abi) * (size_t)nodes + 16, ctx->stream));
    CK(cudaMallocAsync((void **)&d_n, 2 * sizeof(uint64_t), ctx->stream));
    CK(cudaMemcpyAsync(d_tree, tree, sizeof(hb_node_abi) * (size_t)nodes, cudaMemcpyHostToDevice, ctx->stream));
    if (nbytes) CK(cudaMemcpyAsync(ctx->d_comp.p, data, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev0[0], ctx->stream));
    hb_onethread_kernel<<<1, 1, 0, ctx->stream>>>(d_tree, (const uint8_t *)ctx->d_comp.p, bits,
                                                  (uint8_t *)ctx->d_out.p, out_capacity, d_n);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev0[4], ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_res, d_n, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint64_t n = ctx->h_res[0];
    const bool full = ctx->h_res[1] != 0;
    if (n && !full) {
        CK(cudaMemcpyAsync(out, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaFreeAsync(d_tree, ctx->stream));
    CK(cudaFreeAsync(d_n, ctx->stream));
    if (res) {
        memset(res, 0, sizeof(*res));
        res->n_symbols = full ? out_capacity : n;
        res->launches = 1;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev0[0], ctx->ev0[4]) == cudaSuccess) res->ms_total = ms;
    }
    return full ? HB_ERR_OUTPUT_FULL : HB_OK;
}

/* ---- rank-local halves of a multi-GPU decode with HOST buffers -----------------------
 * One process per GPU (torch.distributed, MPI ...): upload + map, the caller exchanges the
 * 32-entry maps, compose (hb_shard_compose), then emit + download.  The copies are DMA
 * transfers when the host buffers are page-locked (hb_host_pin). */
extern "C" int hb_shard_map_host(hb_ctx *ctx, const hb_codebook *cb, const uint8_t *h_comp,
                                 uint64_t comp_bytes, uint64_t bits_own, uint64_t bits_avail,
                                 uint64_t *d_map) {
    if (!ctx || !cb || (!h_comp && comp_bytes)) return HB_ERR_ARG;
    if (comp_bytes < bits_avail / 8 + (bits_avail % 8 != 0)) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint64_t readable = (comp_bytes + 15) / 16 * 16 + 32;
    int rc;
    if ((rc = ensure(ctx, ctx->d_comp, readable))) return rc;
    uint8_t *d = (uint8_t *)ctx->d_comp.p;
    if (comp_bytes) CK(cudaMemcpyAsync(d, h_comp, comp_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d + comp_bytes, 0, readable - comp_bytes, ctx->stream));
    ctx->hs_readable = readable;
    ctx->hs_own = bits_own;
    ctx->hs_avail = bits_avail;
    return hb_shard_map(ctx, cb, d, readable, bits_own, bits_avail, d_map);
}

extern "C" int hb_shard_emit_host(hb_ctx *ctx, const hb_codebook *cb, const uint64_t *d_entry_base,
                                  uint8_t *h_out, uint64_t out_capacity, hb_result *res) {
    if (!ctx || !cb || (!h_out && out_capacity)) return HB_ERR_ARG;
    if (!ctx->have_map || !ctx->hs_readable) return HB_ERR_STATE;
    hb_result local;
    if (!res) res = &local;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->d_out, out_capacity + 16))) return rc;
    rc = hb_shard_emit(ctx, cb, ctx->d_comp.p, ctx->hs_readable, ctx->hs_own, ctx->hs_avail, d_entry_base,
                       ctx->d_out.p, out_capacity, res);
    ctx->hs_readable = 0;
    if (rc) return rc;
    if (res->n_symbols) {
        CK(cudaMemcpyAsync(h_out, ctx->d_out.p, res->n_symbols, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return HB_OK;
}

/* used by hb_gen.cu (setup-only generator) to run on the context's device/stream */
extern "C" int hb_gen_ctx_stream(hb_ctx *ctx, int *device, void **stream) {
    if (!ctx) return HB_ERR_ARG;
    *device = ctx->device;
    *stream = (void *)ctx->stream;
    return HB_OK;
}

/* ---- per-step phase timing over many steps (bench) -------------------------
 * hb_ctx_timing_begin arms a ring of max_steps event sets; every following
 * hb_shard_map/hb_shard_emit pair records into the next set without any host
 * synchronisation.  hb_ctx_timing_collect synchronises the stream and returns
 * the summed CUDA-event milliseconds {sync, scan(+exchange), emit, total}. */
extern "C" int hb_ctx_timing_begin(hb_ctx *ctx, int max_steps) {
    if (!ctx || max_steps < 0 || max_steps > 4096) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (max_steps > ctx->tim_cap) {
        cudaEvent_t *ne = (cudaEvent_t *)realloc(ctx->tim_ev, sizeof(cudaEvent_t) * (size_t)max_steps * HB_NEV);
        if (!ne) return HB_ERR_NOMEM;
        ctx->tim_ev = ne;
        for (int i = ctx->tim_cap * HB_NEV; i < max_steps * HB_NEV; i++) CK(cudaEventCreate(&ctx->tim_ev[i]));
        ctx->tim_cap = max_steps;
    }
    ctx->tim_n = 0;
    ctx->ev = ctx->ev0;
    return HB_OK;
}

extern "C" int hb_ctx_timing_collect(hb_ctx *ctx, double ms[4], int *steps) {
    if (!ctx || !ms) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ms[0] = ms[1] = ms[2] = ms[3] = 0.0;
    for (int k = 0; k < ctx->tim_n; k++) {
        cudaEvent_t *e = ctx->tim_ev + (size_t)k * HB_NEV;
        float a = 0, b = 0, c = 0, d = 0, t = 0;
        CK(cudaEventElapsedTime(&a, e[0], e[1]));
        CK(cudaEventElapsedTime(&b, e[1], e[2]));
        CK(cudaEventElapsedTime(&c, e[2], e[3]));
        CK(cudaEventElapsedTime(&d, e[3], e[4]));
        CK(cudaEventElapsedTime(&t, e[0], e[4]));
        ms[0] += a; ms[1] += b + c; ms[2] += d; ms[3] += t;
    }
    if (steps) *steps = ctx->tim_n;
    ctx->tim_n = ctx->tim_cap;   /* disarm until the next begin */
    ctx->ev = ctx->ev0;
    return HB_OK;
}

/* ---- plain device memory for C hosts ----------------------------------------- */
extern "C" int hb_dev_alloc(hb_ctx *ctx, uint64_t bytes, void **d_ptr) {
    if (!ctx || !d_ptr) return HB_ERR_ARG;
    *d_ptr = nullptr;
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) { cudaGetLastError(); *d_ptr = nullptr; return HB_ERR_NOMEM; }
    CK(cudaMemsetAsync(*d_ptr, 0, bytes, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

extern "C" int hb_dev_free(hb_ctx *ctx, void *d_ptr) {
    if (!ctx) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (d_ptr) CK(cudaFree(d_ptr));
    return HB_OK;
}

extern "C" int hb_dev_upload(hb_ctx *ctx, void *d_dst, const void *h_src, uint64_t bytes) {
    if (!ctx || (bytes && (!d_dst || !h_src))) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

extern "C" int hb_dev_download(hb_ctx *ctx, void *h_dst, const void *d_src, uint64_t bytes) {
    if (!ctx || (bytes && (!h_dst || !d_src))) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}
