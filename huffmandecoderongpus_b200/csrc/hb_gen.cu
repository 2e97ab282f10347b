/*
 * hb_gen.cu -- GPU twin of the bundled generator/encoder in hb_model.c, so that
 * GB-scale synthetic streams (BASELINE.json configs 4 and 5) are produced and
 * verified on the device in a fraction of a second.  Bench/test SETUP only:
 * nothing here is on the timed decode path.  (The chunk-offset prefix sum uses
 * CUB; it is plumbing for stream construction, not a decode kernel.)
 */
#include "huffb200.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdint.h>
#include <stdio.h>

#define HB_GEN_CHUNK 512   /* symbols per thread */
#define HB_GEN_T 128

extern "C" int hb_gen_ctx_stream(hb_ctx *ctx, int *device, void **stream);

struct hb_gen_tables {
    uint32_t nsyms;
    uint32_t cum[256];
    uint8_t symtab[256];
    uint32_t code[256];
    uint8_t codelen[256];
};

__device__ __forceinline__ uint32_t hb_gen_u32(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

/* s_cum[k], s_sym[k]: thresholds and values of the nsyms present symbols */
struct hb_gen_sampler {
    const uint32_t *cum;
    const uint8_t *sym;
    uint32_t nsyms;
    __device__ __forceinline__ uint32_t operator()(uint32_t u) const {
        uint32_t lo = 0, hi = nsyms - 1;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            uint32_t mid = (lo + hi + 1) >> 1;
            if (lo < hi) { if (cum[mid] <= u) lo = mid; else hi = mid - 1; }
        }
        return sym[lo];
    }
};

__global__ void __launch_bounds__(HB_GEN_T)
hb_gen_count_kernel(const hb_gen_tables *__restrict__ tb, uint64_t seed, uint64_t first,
                    uint64_t n, uint64_t nchunks, uint64_t *__restrict__ chunk_bits) {
    __shared__ uint32_t s_cum[256];
    __shared__ uint8_t s_sym[256];
    __shared__ uint8_t s_len[256];
    for (int i = threadIdx.x; i < 256; i += HB_GEN_T) { s_cum[i] = tb->cum[i]; s_sym[i] = tb->symtab[i]; s_len[i] = tb->codelen[i]; }
    __syncthreads();
    const hb_gen_sampler sample{s_cum, s_sym, tb->nsyms};
    uint64_t c = (uint64_t)blockIdx.x * HB_GEN_T + threadIdx.x;
    if (c >= nchunks) return;
    uint64_t i0 = c * HB_GEN_CHUNK, i1 = i0 + HB_GEN_CHUNK < n ? i0 + HB_GEN_CHUNK : n;
    uint32_t bits = 0;
    for (uint64_t i = i0; i < i1; i++) bits += s_len[sample(hb_gen_u32(seed, first + i))];
    chunk_bits[c] = bits;
}

__global__ void __launch_bounds__(HB_GEN_T)
hb_gen_encode_kernel(const hb_gen_tables *__restrict__ tb, uint64_t seed, uint64_t first,
                     uint64_t n, uint64_t nchunks, const uint64_t *__restrict__ chunk_end,
                     uint32_t *__restrict__ words) {
    __shared__ uint32_t s_cum[256];
    __shared__ uint32_t s_code[256];
    __shared__ uint8_t s_sym[256];
    __shared__ uint8_t s_len[256];
    for (int i = threadIdx.x; i < 256; i += HB_GEN_T) {
        s_cum[i] = tb->cum[i]; s_code[i] = tb->code[i]; s_sym[i] = tb->symtab[i]; s_len[i] = tb->codelen[i];
    }
    __syncthreads();
    const hb_gen_sampler sample{s_cum, s_sym, tb->nsyms};
    uint64_t c = (uint64_t)blockIdx.x * HB_GEN_T + threadIdx.x;
    if (c >= nchunks) return;
    uint64_t i0 = c * HB_GEN_CHUNK, i1 = i0 + HB_GEN_CHUNK < n ? i0 + HB_GEN_CHUNK : n;
    uint64_t bit0 = c ? chunk_end[c - 1] : 0ull;
    uint64_t bit1 = chunk_end[c];
    uint64_t widx = bit0 >> 5;
    const uint64_t wlast = bit1 >> 5;        /* word holding the first bit after my range */
    uint64_t acc = 0;                        /* pending bits for word widx, LSB first */
    uint32_t nacc = (uint32_t)(bit0 & 31u);  /* low nacc bits of the first word belong to my left neighbour */
    bool first_word = nacc != 0;
    for (uint64_t i = i0; i < i1; i++) {
        uint32_t s = sample(hb_gen_u32(seed, first + i));
        acc |= (uint64_t)s_code[s] << nacc;
        nacc += s_len[s];
        if (nacc >= 32) {
            uint32_t wv = (uint32_t)acc;
            /* a word is mine alone unless it is the first (shared with the left
             * neighbour) or the last (shared with the right neighbour) */
            if (first_word || widx == wlast) atomicOr(words + widx, wv);
            else words[widx] = wv;
            first_word = false;
            widx++;
            acc >>= 32;
            nacc -= 32;
        }
    }
    if (nacc) atomicOr(words + widx, (uint32_t)acc);
}

__global__ void __launch_bounds__(256)
hb_gen_verify_kernel(const hb_gen_tables *__restrict__ tb, uint64_t seed, uint64_t first,
                     uint64_t n, const uint8_t *__restrict__ out,
                     unsigned long long *__restrict__ mismatches) {
    __shared__ uint32_t s_cum[256];
    __shared__ uint8_t s_sym[256];
    for (int i = threadIdx.x; i < 256; i += 256) { s_cum[i] = tb->cum[i]; s_sym[i] = tb->symtab[i]; }
    __syncthreads();
    const hb_gen_sampler sample{s_cum, s_sym, tb->nsyms};
    unsigned long long bad = 0;
    const uint64_t stride = (uint64_t)gridDim.x * 256 * 16;
    for (uint64_t i0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16; i0 < n; i0 += stride) {
        if (i0 + 16 <= n && ((reinterpret_cast<uintptr_t>(out) + i0) & 15u) == 0) {
            uint4 q = *reinterpret_cast<const uint4 *>(out + i0);
            uint32_t wv[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
            for (int k = 0; k < 16; k++) {
                uint32_t want = sample(hb_gen_u32(seed, first + i0 + k));
                bad += ((wv[k >> 2] >> (8 * (k & 3))) & 0xffu) != want;
            }
        } else {
            for (uint64_t i = i0; i < n && i < i0 + 16; i++)
                bad += out[i] != sample(hb_gen_u32(seed, first + i));
        }
    }
    for (int d = 16; d; d >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, d);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

#define GCK(call)                                                              \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "hb_gen: %s: %s\n", #call, cudaGetErrorString(e_)); \
            rc = HB_ERR_CUDA;                                                  \
            goto done;                                                         \
        }                                                                      \
    } while (0)

static void fill_tables(const hb_model *m, hb_gen_tables *t) {
    t->nsyms = m->nsyms;
    for (int i = 0; i < 256; i++) { t->cum[i] = m->cum[i]; t->symtab[i] = m->symtab[i]; t->code[i] = m->code[i]; t->codelen[i] = m->codelen[i]; }
}

/* count (and optionally encode) symbols [first, first+n) */
static int gen_run(hb_ctx *ctx, const hb_model *m, uint64_t seed, uint64_t first, uint64_t n,
                   void *d_comp, uint64_t comp_capacity, uint64_t *bits_out, bool encode) {
    int device = 0, rc = HB_OK;
    void *sp = nullptr;
    if (!ctx || !m || !bits_out) return HB_ERR_ARG;
    if ((rc = hb_gen_ctx_stream(ctx, &device, &sp))) return rc;
    cudaStream_t st = (cudaStream_t)sp;
    hb_gen_tables ht, *dt = nullptr;
    uint64_t *d_bits = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    uint64_t total = 0;
    const uint64_t nchunks = (n + HB_GEN_CHUNK - 1) / HB_GEN_CHUNK;
    *bits_out = 0;
    if (n == 0) return HB_OK;
    fill_tables(m, &ht);
    GCK(cudaSetDevice(device));
    GCK(cudaMalloc((void **)&dt, sizeof(ht)));
    GCK(cudaMemcpyAsync(dt, &ht, sizeof(ht), cudaMemcpyHostToDevice, st));
    GCK(cudaMalloc((void **)&d_bits, sizeof(uint64_t) * nchunks));
    {
        unsigned grid = (unsigned)((nchunks + HB_GEN_T - 1) / HB_GEN_T);
        hb_gen_count_kernel<<<grid, HB_GEN_T, 0, st>>>(dt, seed, first, n, nchunks, d_bits);
        GCK(cudaGetLastError());
        GCK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d_bits, d_bits, (int64_t)nchunks, st));
        GCK(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
        GCK(cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, d_bits, d_bits, (int64_t)nchunks, st));
        GCK(cudaMemcpyAsync(&total, d_bits + (nchunks - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        GCK(cudaStreamSynchronize(st));
        *bits_out = total;
        if (encode) {
            const uint64_t need = (total + 7) / 8;
            if (!d_comp || comp_capacity < ((need + 3) & ~3ull) ||
                (reinterpret_cast<uintptr_t>(d_comp) & 3u)) { rc = HB_ERR_ARG; goto done; }
            GCK(cudaMemsetAsync(d_comp, 0, comp_capacity, st));
            hb_gen_encode_kernel<<<grid, HB_GEN_T, 0, st>>>(dt, seed, first, n, nchunks, d_bits,
                                                            (uint32_t *)d_comp);
            GCK(cudaGetLastError());
            GCK(cudaStreamSynchronize(st));
        }
    }
done:
    if (d_tmp) cudaFree(d_tmp);
    if (d_bits) cudaFree(d_bits);
    if (dt) cudaFree(dt);
    return rc;
}

extern "C" int hb_gen_encode_device(hb_ctx *ctx, const hb_model *m, uint64_t seed,
                                    uint64_t first_symbol, uint64_t n_symbols, void *d_comp,
                                    uint64_t comp_capacity, uint64_t *bits_out) {
    return gen_run(ctx, m, seed, first_symbol, n_symbols, d_comp, comp_capacity, bits_out, true);
}

extern "C" int hb_gen_count_bits_device(hb_ctx *ctx, const hb_model *m, uint64_t seed,
                                        uint64_t first_symbol, uint64_t n_symbols,
                                        uint64_t *bits_out) {
    return gen_run(ctx, m, seed, first_symbol, n_symbols, nullptr, 0, bits_out, false);
}

extern "C" int hb_gen_verify_device(hb_ctx *ctx, const hb_model *m, uint64_t seed,
                                    uint64_t first_symbol, uint64_t n_symbols, const void *d_out,
                                    uint64_t *mismatches) {
    int device = 0, rc = HB_OK;
    void *sp = nullptr;
    if (!ctx || !m || !mismatches || (!d_out && n_symbols)) return HB_ERR_ARG;
    if ((rc = hb_gen_ctx_stream(ctx, &device, &sp))) return rc;
    cudaStream_t st = (cudaStream_t)sp;
    hb_gen_tables ht, *dt = nullptr;
    unsigned long long *d_bad = nullptr, bad = 0;
    *mismatches = 0;
    if (n_symbols == 0) return HB_OK;
    fill_tables(m, &ht);
    GCK(cudaSetDevice(device));
    GCK(cudaMalloc((void **)&dt, sizeof(ht)));
    GCK(cudaMemcpyAsync(dt, &ht, sizeof(ht), cudaMemcpyHostToDevice, st));
    GCK(cudaMalloc((void **)&d_bad, sizeof(bad)));
    GCK(cudaMemsetAsync(d_bad, 0, sizeof(bad), st));
    {
        uint64_t nthreads = (n_symbols + 15) / 16;
        uint64_t grid = (nthreads + 255) / 256;
        if (grid > 148 * 16) grid = 148 * 16;
        hb_gen_verify_kernel<<<(unsigned)grid, 256, 0, st>>>(dt, seed, first_symbol, n_symbols,
                                                             (const uint8_t *)d_out, d_bad);
        GCK(cudaGetLastError());
    }
    GCK(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    *mismatches = bad;
done:
    if (d_bad) cudaFree(d_bad);
    if (dt) cudaFree(dt);
    return rc;
}
