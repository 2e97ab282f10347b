/*
 * emul.cpp -- TEST INFRASTRUCTURE ONLY: CPU emulation of the CUDA kernels'
 * tile algorithm (huffmandecoderongpus_b200/csrc/hb_kernels.cuh).
 *
 * The build container has no GPU, so the per-thread device functions in
 * hb_core.cuh (shared verbatim with the kernels) are driven here by plain
 * loops that mirror the kernels' phase structure: one loop over "threads" per
 * __syncthreads-delimited phase.  tests/test_emul.py checks the result against
 * the oracle on every corpus, for several (words-per-thread, threads-per-tile)
 * shapes and shard splits, and reads the work counters below to estimate SIMT
 * cost before spending GPU time.  The product never links this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <array>
#include <vector>
#include <algorithm>

#include "hb_core.cuh"

struct emul_stats {
    uint64_t tiles;
    uint64_t rounds_total;      /* stitch rounds summed over tiles */
    uint64_t rounds_max;
    uint64_t probes_walk;       /* codewords decoded in the first walk */
    uint64_t probes_rewalk;     /* ... in stitch rounds */
    uint64_t probes_hyp;        /* ... by hypothesis lanes */
    uint64_t probes_fix;        /* ... by the emit kernel's entry fix-up */
    uint64_t probes_emit;
    uint64_t warp_iters_walk;   /* sum over warps of max-over-lanes trip count */
    uint64_t warp_iters_rewalk;
    uint64_t warp_iters_hyp;
    uint64_t warp_iters_emit;
    uint64_t hyp_unmerged;      /* hypothesis lanes that left the tile unmerged */
    uint64_t tiles_entry_nonzero;
    uint64_t long_probes;       /* probes that needed a second-level table */
};

static thread_local uint64_t g_probe_count;

/* counting wrappers: the probe itself is hb_probe from hb_core.cuh */
template <int WPT>
static uint32_t walk_counted(const hb_lutref &lut, const uint32_t (&w)[WPT + 1], uint32_t lim,
                             uint32_t e, uint32_t (&V)[WPT], uint64_t *n) {
    uint32_t end = hb_walk<WPT>(lut, w, lim, e, V);
    uint64_t c = 0;
    for (int j = 0; j < WPT; j++) c += hb_popc(V[j]);
    *n = c;
    return end;
}

template <int WPT, int T>
struct Emul {
    static constexpr uint32_t S = 32u * WPT;
    static constexpr uint32_t TS = T * S;

    const uint32_t *words; uint64_t nwords, bits_own, bits_avail; uint32_t ntiles;
    hb_lutref lut; uint32_t maxlen;
    std::vector<uint16_t> subs;
    std::vector<uint32_t> tmaps;
    std::vector<uint64_t> wmaps, cmaps, cprefix;
    uint64_t shard_map[32];
    std::vector<uint8_t> tile_entry;
    std::vector<uint64_t> tile_base;
    emul_stats st;

    void load(uint64_t wbase, uint32_t (&w)[WPT + 1]) {
        for (int j = 0; j <= WPT; j++) w[j] = (wbase + j < nwords) ? words[wbase + j] : 0u;
    }

    /* mirrors hb_sync_kernel, one tile */
    void sync_tile(uint32_t tile) {
        const uint64_t tile_bit0 = (uint64_t)tile * TS;
        std::vector<uint32_t> s_comp(T * WPT + 4, 0), s_V(WPT * T, 0), s_cs(T, 0), s_end(T, 0);
        std::vector<uint32_t> lim(T), e(T, 0), endpos(T), c(T);
        std::vector<std::array<uint32_t, WPT + 1>> W(T);
        std::vector<std::array<uint32_t, WPT>> V(T);
        std::vector<uint32_t> trips(T, 0);
        for (int t = 0; t < T; t++) {
            uint32_t w[WPT + 1];
            load((uint64_t)tile * (T * WPT) + (uint64_t)t * WPT, w);
            for (int j = 0; j <= WPT; j++) W[t][j] = w[j];
            for (int j = 0; j < WPT; j++) s_comp[t * WPT + j] = w[j];
            if (t == T - 1) s_comp[T * WPT] = w[WPT];
            const uint64_t sub0 = tile_bit0 + (uint64_t)t * S;
            lim[t] = sub0 >= bits_own ? 0u : (bits_own - sub0 < S ? (uint32_t)(bits_own - sub0) : S);
            uint32_t v[WPT];
            uint64_t n;
            endpos[t] = walk_counted<WPT>(lut, w, lim[t], 0u, v, &n);
            for (int j = 0; j < WPT; j++) V[t][j] = v[j];
            st.probes_walk += n;
            trips[t] = (uint32_t)n;
            s_end[t] = endpos[t];
        }
        for (int w0 = 0; w0 < T; w0 += 32)
            st.warp_iters_walk += *std::max_element(trips.begin() + w0, trips.begin() + std::min(T, w0 + 32));
        uint64_t rounds = 0;
        for (;;) {
            bool any = false;
            std::vector<uint32_t> new_end(endpos);
            std::fill(trips.begin(), trips.end(), 0);
            for (int t = 0; t < T; t++) {
                if (t > 0 && lim[t] > 0) {
                    uint32_t en = (s_end[t - 1] - S) & 31u;
                    if (en != e[t]) {
                        e[t] = en;
                        uint32_t w[WPT + 1], v[WPT], np = 0;
                        for (int j = 0; j <= WPT; j++) w[j] = W[t][j];
                        for (int j = 0; j < WPT; j++) v[j] = V[t][j];
                        uint32_t before = 0;
                        for (int j = 0; j < WPT; j++) before += hb_popc(v[j]);
                        /* count probes = new starts added before the merge point */
                        uint32_t old[WPT];
                        for (int j = 0; j < WPT; j++) old[j] = v[j];
                        bool merged = hb_rewalk<WPT>(lut, w, lim[t], en, v, &np);
                        uint32_t added = 0;
                        for (int j = 0; j < WPT; j++) added += hb_popc(v[j] & ~old[j]);
                        st.probes_rewalk += added;
                        trips[t] = added;
                        for (int j = 0; j < WPT; j++) V[t][j] = v[j];
                        if (!merged && np != endpos[t]) { new_end[t] = np; any = true; }
                    }
                }
            }
            for (int w0 = 0; w0 < T; w0 += 32)
                st.warp_iters_rewalk += *std::max_element(trips.begin() + w0, trips.begin() + std::min(T, w0 + 32));
            endpos = new_end;
            rounds++;
            if (!any) break;
            for (int t = 0; t < T; t++) s_end[t] = endpos[t];
        }
        st.rounds_total += rounds;
        st.rounds_max = std::max<uint64_t>(st.rounds_max, rounds);
        uint32_t C0 = 0;
        for (int t = 0; t < T; t++) {
            const uint64_t sub0 = tile_bit0 + (uint64_t)t * S;
            uint32_t cc = 0;
            for (int j = 0; j < WPT; j++) { cc += hb_popc(V[t][j]); s_V[j * T + t] = V[t][j]; }
            if (cc && sub0 + endpos[t] > bits_avail) cc--;
            c[t] = cc;
            subs[(uint64_t)tile * T + t] = hb_sub_pack(e[t], cc);
            s_cs[t] = C0;
            C0 += cc;
        }
        /* note: the kernel re-publishes s_end only inside the loop; after the
         * final (no-change) round s_end equals endpos */
        for (int t = 0; t < T; t++) s_end[t] = endpos[t];
        const uint64_t own_left = bits_own - tile_bit0, av_left = bits_avail - tile_bit0;
        const uint32_t tile_lim = own_left < TS ? (uint32_t)own_left : TS;
        const uint32_t avail = av_left < 0xffffffffull ? (uint32_t)av_left : 0xffffffffu;
        const uint32_t tl = (tile_lim - 1u) / S;
        const uint32_t X0 = (tl * S + s_end[tl] - tile_lim) & 31u;
        uint32_t maxtrip = 0;
        for (uint32_t t = 0; t < 32; t++) {
            uint32_t m = hb_map_pack32(X0, C0);
            if (t > 0 && t < maxlen) {
                m = hb_hyp_walk<WPT, T>(lut, s_comp.data(), s_V.data(), s_cs.data(), C0, X0, tile_lim, avail, t);
                /* cost accounting: re-run to count steps */
                uint32_t q = t, steps = 0;
                bool merged = false;
                while (q < tile_lim) {
                    uint32_t tt = q / S, j = (q >> 5) & (WPT - 1u);
                    if (s_V[j * T + tt] & hb_bit(q)) { merged = true; break; }
                    uint32_t sym, len = hb_probe(lut, s_comp[q >> 5], s_comp[(q >> 5) + 1], q, &sym);
                    if (q + len > avail) break;
                    q += len; steps++;
                }
                st.probes_hyp += steps;
                maxtrip = std::max(maxtrip, steps);
                if (!merged) st.hyp_unmerged++;
            }
            tmaps[(uint64_t)tile * 32 + t] = m;
        }
        st.warp_iters_hyp += maxtrip;
    }

    /* mirrors hb_scan_up_kernel / hb_scan_top_kernel */
    void scan_up() {
        const uint32_t ncta = (ntiles + 1023u) / 1024u;
        wmaps.assign((size_t)ncta * 32 * 32, 0); cmaps.assign((size_t)ncta * 32, 0); cprefix.assign((size_t)ncta * 32, 0);
        for (uint32_t cta = 0; cta < ncta; cta++) {
            for (uint32_t wid = 0; wid < 32; wid++) {
                uint64_t gw = (uint64_t)cta * 32 + wid;
                for (uint32_t lane = 0; lane < 32; lane++) {
                    uint32_t cur = lane, cnt = 0;
                    for (int j = 0; j < 32; j++) {
                        uint64_t tile = gw * 32 + j;
                        uint32_t m = tile < ntiles ? tmaps[tile * 32 + cur] : hb_map_pack32(cur, 0);
                        cnt += m >> 8; cur = m & 31u;
                    }
                    wmaps[gw * 32 + lane] = hb_map_pack64(cur, cnt);
                }
            }
            for (uint32_t lane = 0; lane < 32; lane++) {
                uint32_t c2 = lane; uint64_t n2 = 0;
                for (int j = 0; j < 32; j++) {
                    uint64_t m = wmaps[((uint64_t)cta * 32 + j) * 32 + c2];
                    n2 += m >> 8; c2 = (uint32_t)m & 31u;
                }
                cmaps[(uint64_t)cta * 32 + lane] = hb_map_pack64(c2, n2);
            }
        }
        for (uint32_t lane = 0; lane < 32; lane++) {
            uint32_t cur = lane; uint64_t cnt = 0;
            for (uint32_t c = 0; c < ncta; c++) {
                cprefix[(uint64_t)c * 32 + lane] = hb_map_pack64(cur, cnt);
                uint64_t m = cmaps[(uint64_t)c * 32 + cur];
                cnt += m >> 8; cur = (uint32_t)m & 31u;
            }
            shard_map[lane] = hb_map_pack64(cur, cnt);
        }
    }

    /* mirrors hb_scan_down_kernel */
    void scan_down(uint32_t E, uint64_t B, uint64_t *result) {
        const uint32_t ncta = (ntiles + 1023u) / 1024u;
        tile_entry.assign(ntiles, 0); tile_base.assign(ntiles, 0);
        result[0] = shard_map[E] >> 8; result[1] = shard_map[E] & 31u; result[2] = E; result[3] = B;
        for (uint32_t cta = 0; cta < ncta; cta++) {
            uint64_t cp = cprefix[(uint64_t)cta * 32 + E];
            uint32_t cur = (uint32_t)cp & 31u; uint64_t b = cp >> 8;
            uint32_t we[32]; uint64_t wb[32];
            for (int j = 0; j < 32; j++) {
                we[j] = cur; wb[j] = b;
                uint64_t m = wmaps[((uint64_t)cta * 32 + j) * 32 + cur];
                b += m >> 8; cur = (uint32_t)m & 31u;
            }
            for (uint32_t wid = 0; wid < 32; wid++) {
                uint64_t gw = (uint64_t)cta * 32 + wid;
                uint32_t c2 = we[wid]; uint64_t b2 = wb[wid];
                for (int j = 0; j < 32; j++) {
                    uint64_t tile = gw * 32 + j;
                    if (tile < ntiles) { tile_entry[tile] = (uint8_t)c2; tile_base[tile] = b2; }
                    uint32_t m = tile < ntiles ? tmaps[tile * 32 + c2] : hb_map_pack32(c2, 0);
                    b2 += m >> 8; c2 = m & 31u;
                }
            }
        }
    }

    struct Sink { uint8_t *p; void operator()(uint32_t n, uint32_t sym) const { p[n] = (uint8_t)sym; } };

    /* mirrors hb_emit_kernel, one tile; returns false on output overflow */
    bool emit_tile(uint32_t tile, uint8_t *out, uint64_t out_capacity, uint32_t stage_bytes) {
        const uint64_t tile_bit0 = (uint64_t)tile * TS;
        const uint32_t E = tile_entry[tile];
        const uint64_t B = tile_base[tile];
        std::vector<uint32_t> s_comp(T * WPT + 4, 0);
        std::vector<uint16_t> s_sub(T);
        std::vector<uint8_t> s_out(stage_bytes + 64, 0xEE);
        std::vector<std::array<uint32_t, WPT + 1>> W(T);
        for (int t = 0; t < T; t++) {
            uint32_t w[WPT + 1];
            load((uint64_t)tile * (T * WPT) + (uint64_t)t * WPT, w);
            for (int j = 0; j <= WPT; j++) W[t][j] = w[j];
            for (int j = 0; j < WPT; j++) s_comp[t * WPT + j] = w[j];
            if (t == T - 1) s_comp[T * WPT] = w[WPT];
            s_sub[t] = subs[(uint64_t)tile * T + t];
        }
        if (E != 0) {
            st.tiles_entry_nonzero++;
            const uint64_t own_left = bits_own - tile_bit0, av_left = bits_avail - tile_bit0;
            const uint32_t tile_lim = own_left < TS ? (uint32_t)own_left : TS;
            const uint32_t avail = av_left < 0xffffffffull ? (uint32_t)av_left : 0xffffffffu;
            std::vector<uint16_t> before(s_sub);
            hb_fix_entries<WPT, T>(lut, s_comp.data(), s_sub.data(), tile_lim, avail, E);
            for (int t = 0; t < T; t++) if (before[t] != s_sub[t]) st.probes_fix += hb_sub_count(s_sub[t]);
        }
        uint32_t o = 0;
        std::vector<uint32_t> off(T), trips(T, 0);
        for (int t = 0; t < T; t++) { off[t] = o; o += hb_sub_count(s_sub[t]); }
        const uint32_t nk = o;
        const uint32_t al = (uint32_t)((reinterpret_cast<uintptr_t>(out) + B) & 15u);
        for (int t = 0; t < T; t++) {
            const uint64_t sub0 = tile_bit0 + (uint64_t)t * S;
            const uint32_t lim = sub0 >= bits_own ? 0u : (bits_own - sub0 < S ? (uint32_t)(bits_own - sub0) : S);
            const uint32_t e = hb_sub_entry(s_sub[t]), c = hb_sub_count(s_sub[t]);
            if (c) {
                if (al + off[t] + c + 1 > stage_bytes) return false;   /* staging bound violated */
                uint32_t w[WPT + 1];
                for (int j = 0; j <= WPT; j++) w[j] = W[t][j];
                Sink sink{s_out.data() + al + off[t]};
                uint32_t n = hb_walk_emit<WPT>(lut, w, lim, e, sink);
                if (n != c && n != c + 1) return false;               /* record/chain mismatch */
                st.probes_emit += n;
                trips[t] = n;
            }
        }
        for (int w0 = 0; w0 < T; w0 += 32)
            st.warp_iters_emit += *std::max_element(trips.begin() + w0, trips.begin() + std::min(T, w0 + 32));
        if (B + nk > out_capacity) return false;
        /* same vector/partial split as the kernel */
        uint8_t *gbase = out + B - al;
        const uint32_t endb = al + nk, nvec = (endb + 15u) >> 4;
        for (uint32_t v = 0; v < nvec; v++) {
            const uint32_t b0 = v << 4;
            if (b0 >= al && b0 + 16u <= endb) memcpy(gbase + b0, s_out.data() + b0, 16);
            else {
                const uint32_t lo = b0 < al ? al : b0, hi = b0 + 16u < endb ? b0 + 16u : endb;
                for (uint32_t i = lo; i < hi; i++) gbase[i] = s_out[i];
            }
        }
        return true;
    }
};

template <int WPT, int T>
static int run(const uint32_t *lut_entries, uint32_t w1, uint32_t maxlen, uint32_t minlen,
               const uint32_t *words, uint64_t nwords, uint64_t bits_own, uint64_t bits_avail,
               int have_entry, uint32_t entry, uint64_t base, uint8_t *out, uint64_t out_capacity,
               uint64_t *shard_map, uint64_t *result, emul_stats *stats) {
    Emul<WPT, T> E;
    memset(&E.st, 0, sizeof(E.st));
    E.words = words; E.nwords = nwords; E.bits_own = bits_own; E.bits_avail = bits_avail;
    E.lut = hb_lutref{lut_entries, lut_entries, (1u << w1) - 1u};
    E.maxlen = maxlen;
    const uint64_t tile_bits = (uint64_t)E.TS;
    E.ntiles = (uint32_t)((bits_own + tile_bits - 1) / tile_bits);
    E.subs.assign((size_t)E.ntiles * T, 0);
    E.tmaps.assign((size_t)E.ntiles * 32, 0);
    E.st.tiles = E.ntiles;
    for (uint32_t tile = 0; tile < E.ntiles; tile++) E.sync_tile(tile);
    E.scan_up();
    if (E.ntiles == 0) for (int e = 0; e < 32; e++) E.shard_map[e] = (uint64_t)e;
    if (shard_map) memcpy(shard_map, E.shard_map, sizeof(E.shard_map));
    int rc = 0;
    if (have_entry) {
        uint64_t res[4] = { 0, entry, entry, base };
        if (E.ntiles) E.scan_down(entry & 31u, base, res);
        uint32_t S = 32u * WPT;
        uint32_t stage = ((T * ((S + minlen - 1) / minlen) + 32u) + 15u) & ~15u;
        for (uint32_t tile = 0; tile < E.ntiles; tile++)
            if (!E.emit_tile(tile, out, out_capacity, stage)) { rc = -6; break; }
        if (result) memcpy(result, res, sizeof(res));
    }
    if (stats) *stats = E.st;
    return rc;
}

extern "C" int emul_run(const uint32_t *lut_entries, uint32_t w1, uint32_t maxlen, uint32_t minlen,
                        const uint32_t *words, uint64_t nwords, uint64_t bits_own,
                        uint64_t bits_avail, int wpt, int T, int have_entry, uint32_t entry,
                        uint64_t base, uint8_t *out, uint64_t out_capacity, uint64_t *shard_map,
                        uint64_t *result, emul_stats *stats) {
#define CASE(W, TT)                                                                              \
    if (wpt == W && T == TT)                                                                     \
        return run<W, TT>(lut_entries, w1, maxlen, minlen, words, nwords, bits_own, bits_avail,  \
                          have_entry, entry, base, out, out_capacity, shard_map, result, stats)
    CASE(4, 256); CASE(8, 256); CASE(16, 256);
    CASE(1, 4); CASE(2, 8); CASE(4, 32); CASE(1, 64);
#undef CASE
    return -4;
}
