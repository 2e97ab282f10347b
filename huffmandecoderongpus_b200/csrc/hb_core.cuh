/*
 * hb_core.cuh -- per-thread decode primitives shared by the sm_100a kernels
 * (hb_kernels.cu) and by the CPU emulation harness used in tests/emul/ to
 * check the tile algorithm without a GPU.  Everything here is plain integer
 * code; HB_HD expands to __host__ __device__ under nvcc and to nothing under
 * a host compiler.
 *
 * Stream semantics (reference framework/mainrun.c:38-55, simpleDecode): bit p
 * of the stream is (data[p/8] >> (p%8)) & 1; bits are consumed root->leaf,
 * 0 -> izero, 1 -> ione.  Read as little-endian 32-bit words, bit p is bit
 * (p & 31) of word p >> 5, so a codeword starting at p is the low bits of the
 * 64-bit window {word[p>>5], word[(p>>5)+1]} shifted right by p & 31.
 */
#ifndef HB_CORE_CUH_
#define HB_CORE_CUH_

#include <stdint.h>

#ifdef __CUDACC__
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD static inline
#endif

#include "hb_format.h"

/* LUT entry (u32):
 *   leaf: bit31 = 0, [7:0] bits consumed at this level (1..width), [15:8] symbol
 *   link: bit31 = 1, [7:0] bits consumed at this level (== this table's width),
 *         [12:8] width of the next table (1..31), [30:13] base index of the
 *         next table inside the same entry array.
 * Level 1 (width w1) lives at index 0 and is staged in shared memory; deeper
 * levels are read from global memory (rare: only codes longer than w1).     */

struct hb_lutref {
    const uint32_t *l1;   /* level-1 table (shared memory on device) */
    const uint32_t *all;  /* whole entry array (global memory on device) */
    uint32_t mask1;       /* (1 << w1) - 1 */
};

HB_HD uint32_t hb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
#endif
}

HB_HD uint32_t hb_bit(uint32_t pos) {  /* 1 << (pos & 31) */
    return 1u << (pos & 31u);
}

HB_HD int hb_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

/* Decode one codeword whose first bit is bit (pos & 31) of lo; hi is the next
 * stream word.  Returns the codeword length; *sym receives the symbol.       */
HB_HD uint32_t hb_probe(const hb_lutref &lut, uint32_t lo, uint32_t hi,
                        uint32_t pos, uint32_t *sym) {
    uint32_t win = hb_funnel_r(lo, hi, pos);
    uint32_t ent = lut.l1[win & lut.mask1];
    if (ent & HB_LUT_LINK) {
        /* second and deeper levels: 64-bit window, global-memory tables */
        uint64_t w64 = (((uint64_t)hi << 32) | lo) >> (pos & 31u);
        uint32_t used = 0;
        do {
            used += ent & 0xffu;
            uint32_t nw = (ent >> 8) & 31u;
            uint32_t base = (ent >> 13) & 0x3ffffu;
            uint32_t idx = (uint32_t)(w64 >> used) & ((1u << nw) - 1u);
            ent = lut.all[base + idx];
        } while (ent & HB_LUT_LINK);
        *sym = (ent >> 8) & 0xffu;
        return used + (ent & 0xffu);
    }
    *sym = (ent >> 8) & 0xffu;
    return ent & 0xffu;
}

/* ---- tile map entries -----------------------------------------------------
 * A transfer map describes a run of the stream [a, b): for every candidate
 * entry offset e (a true codeword starts at a + e, e < 32) it gives the exit
 * offset x (the first codeword starting at or after b starts at b + x) and
 * the number of codewords that start in [a + e, b).  Packed as
 * (count << 8) | x; tile maps use u32, composed maps u64.                   */

HB_HD uint32_t hb_map_pack32(uint32_t x, uint32_t count) { return (count << 8) | x; }
HB_HD uint64_t hb_map_pack64(uint32_t x, uint64_t count) { return (count << 8) | x; }

/* per-subsequence record written by the sync kernel: entry offset (5 bits) and
 * symbol count (11 bits: at most 512 codewords start in a 512-bit run)        */
HB_HD uint16_t hb_sub_pack(uint32_t e, uint32_t c) { return (uint16_t)((c << 5) | e); }
HB_HD uint32_t hb_sub_entry(uint16_t s) { return s & 31u; }
HB_HD uint32_t hb_sub_count(uint16_t s) { return (uint32_t)s >> 5; }

/* ==== per-thread chain walks over one subsequence ==========================
 * A subsequence is WPT consecutive 32-bit words (S = 32*WPT bits) owned by one
 * thread; w[WPT] is the first word of the next subsequence (a codeword may
 * straddle the boundary by at most 31 bits).  Positions are bit offsets from
 * the start of the subsequence.  lim (0..S) is the number of leading bit
 * positions at which a codeword may START here (S, except where the stream
 * ends inside the subsequence).  V[j] has bit (p & 31) set iff a codeword of
 * the chain starts at position p = 32*j + (p & 31).                          */

/* Walk the chain that starts at entry offset e.  Fills V, returns the position
 * of the first codeword start at or after lim (>= S when lim == S). */
template <int WPT>
HB_HD uint32_t hb_walk(const hb_lutref &lut, const uint32_t (&w)[WPT + 1],
                       uint32_t lim, uint32_t e, uint32_t (&V)[WPT]) {
    uint32_t pos = e;
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        uint32_t limj = lim < 32u * (j + 1) ? lim : 32u * (j + 1);
        uint32_t v = 0;
        while (pos < limj) {
            uint32_t sym;
            v |= hb_bit(pos);
            pos += hb_probe(lut, w[j], w[j + 1], pos, &sym);
        }
        V[j] = v;
    }
    return pos;
}

/* Re-walk from a new entry offset e until the chain hits a position already in
 * V (from there on both chains are identical) or runs past lim.  V becomes the
 * chain of e.  Returns true when it merged (the end position is unchanged);
 * otherwise *endpos receives the new end position. */
template <int WPT>
HB_HD bool hb_rewalk(const hb_lutref &lut, const uint32_t (&w)[WPT + 1],
                     uint32_t lim, uint32_t e, uint32_t (&V)[WPT],
                     uint32_t *endpos) {
    uint32_t pos = e;
    bool merged = false;
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        if (!merged) {
            uint32_t limj = lim < 32u * (j + 1) ? lim : 32u * (j + 1);
            uint32_t v = 0;
            while (pos < limj) {
                uint32_t b = hb_bit(pos);
                if (V[j] & b) {           /* old chain passes through pos */
                    v |= V[j] & ~(b - 1u); /* keep its starts at and after pos */
                    merged = true;
                    break;
                }
                uint32_t sym;
                v |= b;
                pos += hb_probe(lut, w[j], w[j + 1], pos, &sym);
            }
            V[j] = v;
        }
    }
    if (!merged) *endpos = pos;
    return merged;
}

/* Walk the chain of entry e and hand every symbol to sink(n, sym), n = 0,1,...
 * Returns the number of symbols (codewords that start before lim). */
template <int WPT, class Sink>
HB_HD uint32_t hb_walk_emit(const hb_lutref &lut, const uint32_t (&w)[WPT + 1],
                            uint32_t lim, uint32_t e, Sink &sink) {
    uint32_t pos = e, n = 0;
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        uint32_t limj = lim < 32u * (j + 1) ? lim : 32u * (j + 1);
        while (pos < limj) {
            uint32_t sym;
            pos += hb_probe(lut, w[j], w[j + 1], pos, &sym);
            sink(n, sym);
            n++;
        }
    }
    return n;
}

/* ==== tile-level walks over shared-memory copies ===========================
 * comp: the tile's T*WPT words followed by the first word of the next tile.
 * Vs:   converged chain of the "entry offset 0" hypothesis, Vs[j*T + t].
 * cs:   exclusive prefix of the per-subsequence symbol counts of that chain.
 * tile_lim: number of leading bit positions of the tile at which a codeword
 *       may start (T*S except in the last tile); avail: number of stream bits
 *       from the start of the tile to the end of the data (a codeword that
 *       would end after avail is incomplete and is not counted, as in the
 *       reference's serial decoder which only emits on reaching a leaf).     */

/* Chain of a non-zero tile entry offset e: follow it until it joins the
 * hypothesis-0 chain (then exit and tail count are those of hypothesis 0) or
 * leaves the tile.  Outputs packed (count << 8) | exit. */
template <int WPT, int T>
HB_HD uint32_t hb_hyp_walk(const hb_lutref &lut, const uint32_t *comp,
                           const uint32_t *Vs, const uint32_t *cs,
                           uint32_t C0, uint32_t X0, uint32_t tile_lim,
                           uint32_t avail, uint32_t e) {
    constexpr uint32_t S = 32u * WPT;
    uint32_t q = e, n = 0;
    for (;;) {
        if (q >= tile_lim) return hb_map_pack32((q - tile_lim) & 31u, n);
        uint32_t t = q / S, j = (q >> 5) & (WPT - 1u);
        uint32_t vw = Vs[j * T + t], b = hb_bit(q);
        if (vw & b) {
            uint32_t below = cs[t] + hb_popc(vw & (b - 1u));
            for (uint32_t jj = 0; jj < j; jj++) below += hb_popc(Vs[jj * T + t]);
            return hb_map_pack32(X0 & 31u, n + C0 - below);
        }
        uint32_t sym;
        uint32_t len = hb_probe(lut, comp[q >> 5], comp[(q >> 5) + 1], q, &sym);
        if (q + len > avail) return hb_map_pack32(0, n);
        q += len;
        n++;
    }
}

/* The true entry offset E of a tile differs from the stored hypothesis: redo
 * the (entry, count) records of the leading subsequences until the chain of E
 * enters a subsequence at the offset already on record. */
template <int WPT, int T>
HB_HD void hb_fix_entries(const hb_lutref &lut, const uint32_t *comp,
                          uint16_t *sub, uint32_t tile_lim, uint32_t avail,
                          uint32_t E) {
    constexpr uint32_t S = 32u * WPT;
    uint32_t e = E;
    for (uint32_t t = 0; t < (uint32_t)T; t++) {
        uint32_t s0 = t * S;
        if (s0 >= tile_lim) break;
        if (hb_sub_entry(sub[t]) == e) break;
        uint32_t lim = tile_lim - s0 < S ? tile_lim - s0 : S;
        uint32_t pos = e, n = 0;
        while (pos < lim) {
            uint32_t q = s0 + pos, sym;
            uint32_t len = hb_probe(lut, comp[q >> 5], comp[(q >> 5) + 1], q, &sym);
            if (q + len > avail) { pos = S + 31u; break; }
            pos += len;
            n++;
        }
        sub[t] = hb_sub_pack(e, n);
        e = (pos - S) & 31u;
    }
}

#endif /* HB_CORE_CUH_ */
