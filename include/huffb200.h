/*
 * huffb200.h -- C ABI of libhuffb200.so, the B200 (sm_100a) speculative
 * parallel Huffman decoder that replaces the reference's GPU decode backend.
 *
 * Plain pointers and sizes only; callable from C without CUDA headers.  The
 * reference conflates allocation, transfer, decode and download in one call
 * (framework/fastgpu.cu:140-332, timed as a whole by framework/decodeUtil.c:41-43);
 * this ABI separates them so a harness can report device time on resident
 * data AND end-to-end time:
 *
 *   reference interface replaced                        entry point here
 *   --------------------------------------------------  -----------------------------
 *   fastgpuApproach(cd, out, NULL)   fastgpu.cu:140      b200Approach (b200approach.h)
 *     cudaMalloc/cudaMemcpy tree+data fastgpu.cu:196-201  hb_codebook_create, hb_decode_host
 *     decodeAllBits..calcresult       fastgpu.cu:214-305  hb_decode_device
 *     (no multi-GPU in the reference)                     hb_shard_map / _compose / _emit
 *   loadHuffFile                     huffdata.c:27-68    hb_huff_load / hb_huff_free
 *   tableHeight / tableMinDepth      huffdata.c:224,272  hb_codebook_info
 *
 * Every function returns HB_OK (0) or a negative HB_ERR_* code; nothing here
 * prints or exits (the reference's print-and-exit behaviour, fastgpu.cu:16-31,
 * is reproduced only at the approach boundary, b200Approach).
 * There is no CPU fallback: without a CUDA device hb_ctx_create fails.
 */
#ifndef HUFFB200_H_
#define HUFFB200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK               0
#define HB_ERR_CUDA        -1  /* CUDA runtime error (hb_last_error has the text) */
#define HB_ERR_TREE        -2  /* malformed tree: bad child index, half-leaf, cycle, leaf root */
#define HB_ERR_CODELEN     -3  /* a codeword is longer than 32 bits */
#define HB_ERR_ARG         -4  /* bad argument (NULL, misaligned device pointer, sizes) */
#define HB_ERR_NOMEM       -5
#define HB_ERR_OUTPUT_FULL -6  /* decoded length exceeds out_capacity (nothing past it is written) */
#define HB_ERR_IO          -7
#define HB_ERR_FORMAT      -8  /* not a HUFF / HUF8 file */
#define HB_ERR_STATE       -9  /* call order (e.g. hb_shard_emit without hb_shard_map) */

/* Tree node: identical layout to the reference's struct HuffNode
 * (framework/huffdata.h:12-16): node 0 is the root, leaf <=> izero == ione == -1. */
typedef struct hb_node_abi {
    uint8_t sym;
    int32_t izero;
    int32_t ione;
} hb_node_abi;

typedef struct hb_ctx hb_ctx;           /* one device + one stream + scratch */
typedef struct hb_codebook hb_codebook; /* lookup tables resident on the device */

typedef struct hb_result {
    uint64_t n_symbols;   /* decoded bytes written by this call */
    uint64_t out_base;    /* index of this shard's first symbol in the whole stream */
    uint32_t exit_offset; /* bit offset, past the end of the owned range, of the next codeword */
    uint32_t entry_offset;/* bit offset of the first owned codeword */
    uint32_t launches;    /* kernels launched by this call */
    uint32_t tiles;
    float ms_total;       /* CUDA-event time of all kernels of this call */
    float ms_sync;        /* phase 1: per-subsequence chains + tile maps */
    float ms_scan;        /* phase 2: map composition across tiles */
    float ms_emit;        /* phase 3: decode + coalesced write */
} hb_result;

const char *hb_strerror(int code);
const char *hb_version(void);

/* ---- context ------------------------------------------------------------ */
/* cuda_stream: a cudaStream_t created by the caller (e.g. torch's current
 * stream) or NULL to let the context own a non-blocking stream. */
int  hb_ctx_create(int device, void *cuda_stream, hb_ctx **ctx);
void hb_ctx_destroy(hb_ctx *ctx);
const char *hb_last_error(const hb_ctx *ctx);
/* words_per_thread: 4, 8 or 16 32-bit words per subsequence (0 = default);
 * ctas_per_sm: persistent CTAs per SM (0 = occupancy-derived). */
int  hb_ctx_configure(hb_ctx *ctx, int words_per_thread, int ctas_per_sm);
/* Which sync kernel resolves the chains of full tiles.  The byte-step transducer kernel
 * needs a code table with a transducer (every tree with at most 256 internal nodes that
 * is not a fixed-length code); the multi-symbol probe kernel handles everything,
 * including the partial tile at the end of a shard.
 *   HB_SYNC_AUTO   transducer kernel for streams of at least two waves of tiles, one
 *                  probe-kernel launch for smaller ones (a second launch and a second
 *                  table load cost more than they save there)
 *   HB_SYNC_PROBE  probe kernel only        HB_SYNC_FSM  transducer kernel whenever possible
 * Identical records either way; the knob exists for A/B measurement and tests. */
#define HB_SYNC_AUTO  0
#define HB_SYNC_PROBE 1
#define HB_SYNC_FSM   2
int  hb_ctx_set_sync_path(hb_ctx *ctx, int path);
/* Copies of the transducer table in the sync kernel's shared memory, each on its own banks (fewer
 * bank-conflict replays): log2 of the count, 0 / 1 / 2, or -1 = automatic (four copies when the
 * table is small enough and the stream large enough).  A/B knob. */
int  hb_ctx_set_sync_copies(hb_ctx *ctx, int log2_copies);
/* How the emit kernel fills its staging buffer:
 *   HB_EMIT_WORDS   whole 32-bit words assembled in registers, up to four symbols per table
 *                   probe (hb_emitw_kernel)
 *   HB_EMIT_BYTES   byte stores, two symbols per probe (hb_emit_kernel)
 *   HB_EMIT_AUTO    = HB_EMIT_WORDS (the fastest on every workload measured)
 * Identical output; the knob exists for A/B measurement and tests. */
#define HB_EMIT_AUTO  0
#define HB_EMIT_BYTES 1
#define HB_EMIT_WORDS 2
#define HB_EMIT_FLAT  3   /* hb_emitf_kernel (one loop per subsequence, lane-private table copies) on
                             every tile but the last; 8 words per thread only.  Experimental: measured
                             slower than HB_EMIT_WORDS, never chosen by HB_EMIT_AUTO */
#define HB_EMIT_WORDS32 4 /* hb_emit32_kernel: word stores, 32-bit table entries with up to three symbols, 4
                             (or 8) copies of the table on disjoint banks */
#define HB_EMIT_WORDS32W 5 /* hb_emit32w_kernel: the same probes, every warp on its own (own staging slice, own bulk
                             store, no block-level barriers) */
#define HB_EMIT_WORDS64W 6 /* hb_emit32w_kernel<E64>: the warp-autonomous pipeline over the E64-table (up to four
                             symbols per probe): codes so short that three symbols do not fill a probe */
int  hb_ctx_set_emit_path(hb_ctx *ctx, int path);
/* HB_EMIT_WORDS32W: consecutive subsequences a lane decodes in one go (1 or 2; default 1: 2 was measured
 * slower).  A/B knob. */
int  hb_ctx_set_emit_lane_subsequences(hb_ctx *ctx, int n);
/* name of the emit kernel the last decode used for the bulk of its tiles ("" before the first) */
const char *hb_ctx_last_emit_kernel(const hb_ctx *ctx);
/* EP-table of the flat emit kernel: index width in bits (8..12, 0 = automatic) and log2 of the
 * number of copies interleaved in shared memory (0..4, -1 = automatic).  A/B knob. */
int  hb_ctx_set_emit_table(hb_ctx *ctx, int index_bits, int log2_copies);
/* hb_result's per-phase times need three extra CUDA events per decode (~4 us each).
 * HB_PHASES_AUTO records them for streams of more than 1024 tiles and while a timing ring
 * is armed (hb_ctx_timing_begin); smaller decodes then report ms_total only. */
#define HB_PHASES_AUTO   0
#define HB_PHASES_ALWAYS 1
#define HB_PHASES_NEVER  2
int  hb_ctx_set_phase_timing(hb_ctx *ctx, int mode);
int  hb_ctx_sync(hb_ctx *ctx);
/* hb_decode_host cuts streams of at least two chunks into chunks of this many
 * compressed bytes (rounded to whole tiles) and overlaps upload, decode and
 * download; 0 restores the default (32 MiB). */
int  hb_ctx_set_host_chunk(hb_ctx *ctx, uint64_t bytes);
/* Where the next shards begin in their stream: first_byte = index, in the WHOLE stream, of the
 * first byte handed to hb_shard_map (known = 0: unknown, the default).  Optional, and only a
 * speed matter: for codes whose codeword lengths all share a factor g that is not a power of
 * two, it lets the kernels skip entry offsets that cannot occur at that position (their chains
 * never merge with the true one and are slow to follow).  hb_decode_device, hb_decode_host and
 * hb_multi_* set it themselves. */
int  hb_ctx_set_shard_origin(hb_ctx *ctx, uint64_t first_byte, int known);
/* Phase timing over many steps without host synchronisation in between:
 * _begin arms a ring of max_steps CUDA-event sets (one per following
 * hb_shard_map + hb_shard_emit pair); _collect synchronises the stream and
 * returns the summed milliseconds ms[0..3] = {sync, scan (+ exchange between
 * the two calls), emit, total} and the number of steps recorded. */
int  hb_ctx_timing_begin(hb_ctx *ctx, int max_steps);
int  hb_ctx_timing_collect(hb_ctx *ctx, double ms[4], int *steps);
int  hb_device_info(hb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                    uint64_t *total_mem);

/* Device memory for C hosts that keep streams resident without CUDA headers
 * (zero-filled, 256-byte aligned).  hb_dev_upload / _download are synchronous. */
int  hb_dev_alloc(hb_ctx *ctx, uint64_t bytes, void **d_ptr);
int  hb_dev_free(hb_ctx *ctx, void *d_ptr);
int  hb_dev_upload(hb_ctx *ctx, void *d_dst, const void *h_src, uint64_t bytes);
int  hb_dev_download(hb_ctx *ctx, void *h_dst, const void *d_src, uint64_t bytes);

/* ---- codebook ------------------------------------------------------------ */
int  hb_codebook_create(hb_ctx *ctx, const hb_node_abi *tree, int nodes,
                        hb_codebook **cb);
void hb_codebook_destroy(hb_codebook *cb);
/* Copy one of the device-resident code tables back to the host (tests: the tables are
 * built by a kernel and compared with csrc/hb_lut.c's host construction). */
#define HB_TABLE_LUT 0   /* single-symbol multi-level table (built on the host) */
#define HB_TABLE_S   1
#define HB_TABLE_E   2
#define HB_TABLE_E64 3
#define HB_TABLE_FSM 5   /* byte-step transducer; 0 bytes when the tree has none */
int  hb_codebook_download_table(const hb_codebook *cb, int which, void *dst, uint64_t capacity,
                                uint64_t *bytes);
int  hb_codebook_info(const hb_codebook *cb, uint32_t *maxlen, uint32_t *minlen,
                      uint32_t *w1, uint32_t *n_entries);

/* ---- whole stream, device-resident input and output ----------------------- */
/* d_comp: device pointer, 16-byte aligned, comp_bytes >= ceil(bits/8) readable
 * bytes (the kernels read whole 32-bit words: comp_bytes rounded up to a multiple of 4 must
 * be readable, which any 16-byte-granular allocation gives).  d_out: device pointer with out_capacity bytes.  Runs on the
 * context's stream, synchronises it, and fills *res.  res == NULL: the kernels are only queued (no
 * synchronisation, nothing read back; hb_ctx_sync / the next call with a result orders after them). */
int hb_decode_device(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp,
                     uint64_t comp_bytes, uint64_t bits, void *d_out,
                     uint64_t out_capacity, hb_result *res);

/* ---- byte-range shards (multi-GPU or streaming) --------------------------- */
/* A shard owns the codewords that START in its first bits_own bits; d_comp
 * holds bits_avail >= bits_own bits (the surplus is the halo into the next
 * shard, >= 32 bits unless the stream ends).  hb_shard_map leaves, in device
 * memory d_map (32 x u64), the shard's transfer map: entry e ->
 * (symbol count << 8) | exit offset.  Asynchronous on the context's stream. */
int hb_shard_map(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp,
                 uint64_t comp_bytes, uint64_t bits_own, uint64_t bits_avail,
                 uint64_t *d_map);
/* d_all_maps: n_ranks x 32 x u64 (the all-gathered maps, rank-major, device).
 * Composes ranks 0..rank-1 from entry offset 0 and writes
 * d_entry_base[0] = this rank's entry offset, [1] = its output base,
 * [2] = total symbols of all ranks.  Asynchronous. */
int hb_shard_compose(hb_ctx *ctx, const uint64_t *d_all_maps, int n_ranks,
                     int rank, uint64_t *d_entry_base);
/* The exchange + compose WITHOUT a collective library, for one process per GPU on one NVLink domain:
 * every rank exports a small exchange table (hb_peer_export: a 64-byte CUDA IPC handle the caller
 * passes around once, e.g. with torch.distributed / MPI all-gather), opens the tables of the ranks
 * to its right (hb_peer_connect: handles = n_ranks x 64 bytes, rank-major; a barrier of the caller
 * must follow before the first exchange), and then, per decode, calls hb_shard_exchange between
 * hb_shard_map and hb_shard_emit: ONE kernel stores this rank's map into its right neighbours'
 * tables (peer stores over NVLink), waits for the maps of its left neighbours and writes
 * d_entry_base[0..2] = entry offset, output base, symbols through this rank.  seq: the same value
 * on every rank for the same decode, > 0 and increasing by one (a rank may run up to 1023 decodes
 * ahead of another).  A rank that never delivers is reported by hb_shard_emit / hb_shard_result as
 * HB_ERR_STATE after 5 s.  Replaces the caller's all-gather + hb_shard_compose. */
#define HB_PEER_HANDLE_BYTES 64
int hb_peer_export(hb_ctx *ctx, void *handle);
int hb_peer_connect(hb_ctx *ctx, int rank, int n_ranks, const void *handles);
/* the same for contexts of one process (ctxs[rank] == ctx; several devices, or several contexts on one) */
int hb_peer_connect_local(hb_ctx *ctx, int rank, int n_ranks, hb_ctx *const *ctxs);
int hb_shard_exchange(hb_ctx *ctx, uint64_t seq, uint64_t *d_entry_base);
/* Teardown: hb_peer_disconnect unmaps the other ranks' tables; hb_peer_close also frees this rank's own
 * (hb_ctx_destroy calls it).  With several processes: disconnect everywhere, a barrier, then close, so
 * that no table is freed while a neighbour still has it mapped. */
int hb_peer_disconnect(hb_ctx *ctx);
int hb_peer_close(hb_ctx *ctx);
/* Second half: fix every tile's entry offset / output base from
 * d_entry_base[0..1] (NULL = entry 0, base 0) and write the shard's symbols to
 * d_out[0 .. n).  Must follow hb_shard_map on the same context with the same
 * stream arguments.  Synchronises and fills *res when res != NULL. */
int hb_shard_emit(hb_ctx *ctx, const hb_codebook *cb, const void *d_comp,
                  uint64_t comp_bytes, uint64_t bits_own, uint64_t bits_avail,
                  const uint64_t *d_entry_base, void *d_out,
                  uint64_t out_capacity, hb_result *res);

/* The same two halves with HOST buffers (one process per GPU): hb_shard_map_host uploads the
 * shard (h_comp: comp_bytes bytes, halo included) and maps it; after the exchange and
 * hb_shard_compose, hb_shard_emit_host emits it and downloads the bytes into h_out. */
int hb_shard_map_host(hb_ctx *ctx, const hb_codebook *cb, const uint8_t *h_comp, uint64_t comp_bytes,
                      uint64_t bits_own, uint64_t bits_avail, uint64_t *d_map);
int hb_shard_emit_host(hb_ctx *ctx, const hb_codebook *cb, const uint64_t *d_entry_base,
                       uint8_t *h_out, uint64_t out_capacity, hb_result *res);

/* Result of the last hb_shard_emit on this context (synchronises its stream): for callers
 * that queued the emit with res == NULL to keep several devices busy at once. */
int hb_shard_result(hb_ctx *ctx, hb_result *res);

/* ---- one process, N devices ---------------------------
 * The stream is cut into byte-range shards, one per device; the 32-entry shard maps travel by
 * peer copies ordered with events (256 B each; no NCCL, no host round trip), every device
 * composes the maps to its left and emits its shard into its own output slice.
 * devices == NULL: devices 0 .. n_devices-1; n_devices <= 0: all visible devices. */
#define HB_MULTI_MAX 8
typedef struct hb_multi hb_multi;
typedef struct hb_multi_result {
    uint64_t n_symbols;                   /* whole stream */
    int32_t  n_devices;                   /* shards actually used (tiny streams use fewer devices) */
    uint32_t launches;                    /* kernels launched on all devices */
    float    ms_device_max;               /* max over devices of the CUDA-event time map .. emit */
    float    ms_wall;                     /* host wall clock of the call */
    uint64_t shard_symbols[HB_MULTI_MAX];
    float    shard_ms[HB_MULTI_MAX];
} hb_multi_result;
int  hb_multi_create(const int *devices, int n_devices, hb_multi **m);
void hb_multi_destroy(hb_multi *m);
int  hb_multi_devices(const hb_multi *m);
const char *hb_multi_last_error(const hb_multi *m);
/* resident: shard a host stream over the devices (or build a synthetic one on them with the
 * bundled generator), then decode as often as wanted; device-timed */
int  hb_multi_load(hb_multi *m, const hb_node_abi *tree, int nodes, const uint8_t *data, uint64_t bits);
int  hb_multi_generate(hb_multi *m, int model_kind, uint64_t seed, uint64_t n_symbols, uint64_t *bits_out);
int  hb_multi_decode(hb_multi *m, hb_multi_result *res);
int  hb_multi_download(hb_multi *m, uint8_t *out, uint64_t out_capacity);
/* every output slice against regenerated symbols (streams made by hb_multi_generate) */
int  hb_multi_verify(hb_multi *m, int model_kind, uint64_t seed, uint64_t *mismatches);
/* host buffers in and out: upload, decode, download (what b200ApproachMulti calls) */
int  hb_multi_decode_host(hb_multi *m, const hb_node_abi *tree, int nodes, const uint8_t *data,
                          uint64_t bits, uint8_t *out, uint64_t out_capacity, hb_multi_result *res);

/* Page-lock / release a caller-owned host range (cudaHostRegister, portable) so that the
 * copies of hb_decode_host / hb_multi_decode_host are true DMA transfers that overlap across
 * devices.  Ranges that are already page-locked are accepted (HB_OK).  Optional. */
int  hb_host_pin(const void *ptr, uint64_t bytes);
int  hb_host_unpin(const void *ptr);

/* ---- host buffers (what the approach call does) --------------------------- */
/* Upload tree + data, decode, download.  data must have >= ceil(bits/8) bytes.
 * Device buffers are cached in the context and grow on demand. */
int hb_decode_host(hb_ctx *ctx, const hb_node_abi *tree, int nodes,
                   const uint8_t *data, uint64_t bits, uint8_t *out,
                   uint64_t out_capacity, hb_result *res);

/* The whole stream on ONE device thread: the reference's "onethread" approach
 * (framework/onethread.cu:13-52), a bit-serial walk of the uploaded node array that shares
 * nothing with the table-driven kernels.  Debug aid (seconds per 100 MB), host buffers. */
int hb_decode_onethread(hb_ctx *ctx, const hb_node_abi *tree, int nodes, const uint8_t *data,
                        uint64_t bits, uint8_t *out, uint64_t out_capacity, hb_result *res);

/* ---- .huff container ------------------------------------------------------ */
/* "HUFF" (reference framework/huffdata.c:27-68: BE i32 nodes, bits, usize) and
 * "HUF8" (this repo: BE u64 bits, usize) files.  data gets >= 16 zero bytes of
 * padding.  Free with hb_huff_free. */
typedef struct hb_huff_file {
    int32_t      nodes;
    int32_t      wide;    /* 1 = HUF8 */
    uint64_t     bits;
    uint64_t     usize;
    hb_node_abi *tree;
    uint8_t     *data;
} hb_huff_file;
int  hb_huff_load(const char *path, hb_huff_file *out);
int  hb_huff_save(const char *path, const hb_huff_file *in, int wide);
void hb_huff_free(hb_huff_file *f);

/* ---- bundled synthetic-stream generator (the reference has no
 * encoder).  Symbol i of a stream is a pure function of (model, seed, i). ---- */
#define HB_MODEL_ENGLISH   0  /* order-0 byte histogram of the reference's files/bible.txt */
#define HB_MODEL_FIBONACCI 1  /* 256 symbols, Fibonacci-skewed, 20 < max code length <= 32 */
#define HB_MODEL_DNA       2  /* 4 equiprobable symbols: all codes 2 bits, never self-synchronises */
#define HB_MODEL_UNIFORM8  3  /* 8 equiprobable symbols: all codes 3 bits (adversarial for sync) */

typedef struct hb_model {
    int32_t     nodes;           /* Huffman tree in the reference's node format */
    hb_node_abi tree[511];
    uint32_t    cum[256];        /* sampling thresholds over the nsyms present symbols, in symbol
                                    order: draw u32 u, pick the largest k < nsyms with cum[k] <= u */
    uint8_t     symtab[256];     /* k -> symbol value */
    uint32_t    code[256];       /* LSB-first code bits per symbol value */
    uint8_t     codelen[256];
    uint32_t    maxlen, minlen, nsyms;
} hb_model;
int hb_model_build(int kind, hb_model *m);
/* CPU generation / encoding (spec of the stream; used for small cases and tests) */
void     hb_gen_symbols_cpu(const hb_model *m, uint64_t seed, uint64_t first, uint64_t n, uint8_t *out);
uint64_t hb_encode_bits_cpu(const hb_model *m, const uint8_t *syms, uint64_t n);           /* total bits */
void     hb_encode_cpu(const hb_model *m, const uint8_t *syms, uint64_t n, uint8_t *out);   /* out zeroed, ceil(bits/8)+8 bytes */
/* GPU generation / encoding, device-resident: d_comp must hold comp_capacity
 * zero-initialisable bytes; *bits_out receives the stream length. */
int hb_gen_encode_device(hb_ctx *ctx, const hb_model *m, uint64_t seed,
                         uint64_t first_symbol, uint64_t n_symbols,
                         void *d_comp, uint64_t comp_capacity, uint64_t *bits_out);
/* Compare d_out[0..n) with regenerated symbols first_symbol.. ; *mismatches out. */
int hb_gen_verify_device(hb_ctx *ctx, const hb_model *m, uint64_t seed,
                         uint64_t first_symbol, uint64_t n_symbols,
                         const void *d_out, uint64_t *mismatches);
/* Exact encoded length, in bits, of symbols [first, first + n) (device-side count). */
int hb_gen_count_bits_device(hb_ctx *ctx, const hb_model *m, uint64_t seed,
                             uint64_t first_symbol, uint64_t n_symbols,
                             uint64_t *bits_out);

#ifdef __cplusplus
}
#endif
#endif /* HUFFB200_H_ */
