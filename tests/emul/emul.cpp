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

template <int WPT, int T>
struct Emul {
    static constexpr uint32_t S = 32u * WPT;
    static constexpr uint32_t TS = T * S;

    const uint32_t *words; uint64_t nwords, bits_own, bits_avail; uint32_t ntiles;
    hb_tables tbS, tbE; uint32_t maxlen, minlen, gmod = 1, gorg = 0;   /* see hb_stream_args */
    bool possible(uint32_t tile, uint32_t e) const {
        return gmod <= 1u || (gorg + (tile % gmod) * (TS % gmod) + e) % gmod == 0u;
    }
    uint32_t first_entry(uint32_t tile, uint32_t sub_bit0) const {   /* hb_first_entry in hb_kernels.cuh */
        if (gmod <= 1u) return 0u;
        const uint32_t p = (gorg + (tile % gmod) * (TS % gmod) + sub_bit0 % gmod) % gmod;
        return p ? gmod - p : 0u;
    }
    hb_tables64 tbE64; int emit_mode = 0;   /* 0 byte stores (E-table), 1 word stores (E64-table), 3 word stores (E32-table) */
    hb_tables32 tbE32; std::vector<uint32_t> e32tab;
    bool flat = false;                       /* flat walk (EP-table) on all tiles but the last */
    std::vector<uint32_t> eptab; uint32_t ep_wf = 10;   /* plain EP-table */
    uint64_t flat_tiles = 0;
    hb_fsm fsm; bool have_fsm = false; int sync_mode = 0;   /* 0 probe, 1 transducer on full tiles, 2 both + compare */
    uint64_t fsm_tiles = 0, fsm_mismatch = 0;
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
    static uint32_t total(const uint32_t (&rec)[WPT]) {
        uint32_t c = 0;
        for (int j = 0; j < WPT; j++) c += hb_rec_cnt(rec[j]);
        return c;
    }

    /* mirrors hb_sync_kernel, one tile */
    void sync_tile(uint32_t tile) {
        const uint64_t tile_bit0 = (uint64_t)tile * TS;
        std::vector<uint32_t> s_comp(T * WPT + 4, 0), s_rec(WPT * T, 0), s_cs(T, 0), s_land(T, 0);
        std::vector<uint32_t> lim(T), e(T, 0);
        std::vector<std::array<uint32_t, WPT + 1>> W(T);
        std::vector<std::array<uint32_t, WPT>> R(T);
        std::vector<uint32_t> trips(T, 0);
        for (int t = 0; t < T; t++) {
            uint32_t w[WPT + 1], rec[WPT];
            load((uint64_t)tile * (T * WPT) + (uint64_t)t * WPT, w);
            for (int j = 0; j <= WPT; j++) W[t][j] = w[j];
            for (int j = 0; j < WPT; j++) s_comp[t * WPT + j] = w[j];
            if (t == T - 1) s_comp[T * WPT] = w[WPT];
            const uint64_t sub0 = tile_bit0 + (uint64_t)t * S;
            lim[t] = sub0 >= bits_own ? 0u : (bits_own - sub0 < S ? (uint32_t)(bits_own - sub0) : S);
            const bool fixed_len = minlen == maxlen;
            e[t] = fixed_len ? hb_fixed_next(0u, maxlen, (uint32_t)t * S) - (uint32_t)t * S
                             : first_entry(tile, (uint32_t)t * S);
            hb_walk<WPT>(tbS, w, lim[t], e[t], rec);
            /* cross-check: the multi-symbol walk equals the symbol-by-symbol walk */
            if (lim[t] == S) {
                uint32_t acc = e[t];
                for (int j = 0; j < WPT; j++) {
                    uint32_t land, cnt;
                    hb_word_slow(tbS.slow, w[j], w[j + 1], 32u, acc, land, cnt);
                    if (hb_rec_pack(land, cnt) != rec[j]) st.long_probes |= 1ull << 63;  /* flag */
                }
            }
            for (int j = 0; j < WPT; j++) R[t][j] = rec[j];
            st.probes_walk += total(rec);
            trips[t] = total(rec);
            s_land[t] = hb_rec_land(rec[WPT - 1]);
        }
        for (int w0 = 0; w0 < T; w0 += 32)
            st.warp_iters_walk += *std::max_element(trips.begin() + w0, trips.begin() + std::min(T, w0 + 32));
        uint64_t rounds = 0;
        for (;;) {
            bool any = false;
            std::vector<uint32_t> new_land(s_land);
            std::fill(trips.begin(), trips.end(), 0);
            for (int t = 0; t < T; t++) {
                if (t > 0 && lim[t] > 0) {
                    uint32_t en = s_land[t - 1];
                    if (en != e[t]) {
                        e[t] = en;
                        uint32_t w[WPT + 1], rec[WPT], old[WPT];
                        for (int j = 0; j <= WPT; j++) w[j] = W[t][j];
                        for (int j = 0; j < WPT; j++) old[j] = rec[j] = R[t][j];
                        bool changed = hb_rewalk<WPT>(tbS, w, lim[t], en, rec);
                        uint32_t walked = 0;   /* words re-walked, for the cost estimate */
                        for (int j = 0; j < WPT; j++) {
                            walked++;
                            if (hb_rec_land(rec[j]) == hb_rec_land(old[j])) break;
                        }
                        for (int j = 0; j < (int)walked && j < WPT; j++) { st.probes_rewalk += hb_rec_cnt(rec[j]); trips[t] += hb_rec_cnt(rec[j]); }
                        for (int j = 0; j < WPT; j++) R[t][j] = rec[j];
                        new_land[t] = hb_rec_land(rec[WPT - 1]);
                        if (changed) any = true;
                    }
                }
            }
            for (int w0 = 0; w0 < T; w0 += 32)
                st.warp_iters_rewalk += *std::max_element(trips.begin() + w0, trips.begin() + std::min(T, w0 + 32));
            rounds++;
            if (!any) break;
            s_land = new_land;
        }
        st.rounds_total += rounds;
        st.rounds_max = std::max<uint64_t>(st.rounds_max, rounds);
        uint32_t C0 = 0;
        for (int t = 0; t < T; t++) {
            uint32_t cc = 0;
            for (int j = 0; j < WPT; j++) { cc += hb_rec_cnt(R[t][j]); s_rec[j * T + t] = R[t][j]; }
            subs[(uint64_t)tile * T + t] = hb_sub_pack(e[t], cc);
            s_cs[t] = C0;
            C0 += cc;
        }
        const uint64_t own_left = bits_own - tile_bit0;
        const uint32_t tile_lim = own_left < TS ? (uint32_t)own_left : TS;
        const uint32_t wl = (tile_lim - 1u) >> 5;
        const uint32_t X0 = hb_rec_land(s_rec[(wl % WPT) * T + wl / WPT]);
        for (uint32_t t = 0; t < 32; t++) {
            uint32_t m = hb_map_pack32(X0, C0);
            if (t > 0 && t < maxlen && possible(tile, t)) {
                if (minlen == maxlen) {
                    const uint32_t n = hb_fixed_count(t, maxlen, 0u, tile_lim);
                    m = hb_map_pack32((hb_fixed_next(t, maxlen, tile_lim) - tile_lim) & 31u, n);
                } else {
                    m = hb_hyp_walk<WPT, T>(tbS, s_comp.data(), s_rec.data(), s_cs.data(), C0, X0, tile_lim, t);
                }
                if ((m & 31u) != X0) st.hyp_unmerged++;
            }
            tmaps[(uint64_t)tile * 32 + t] = m;
        }
    }

    /* mirrors hb_fsm_sync_kernel, one FULL tile (one group of T threads) */
    void sync_tile_fsm(uint32_t tile) {
        std::vector<uint16_t> s_rec(WPT * T, 0);
        std::vector<uint32_t> s_cs(T, 0), s_exit(T, 0), st_in(T, 0);
        std::vector<std::array<uint32_t, WPT + 1>> W(T);
        std::vector<std::array<uint32_t, WPT>> R(T);
        const uint64_t tbase = (uint64_t)tile * (T * WPT);
        for (int t = 0; t < T; t++) {
            uint32_t w[WPT + 1], rec[WPT];
            load(tbase + (uint64_t)t * WPT, w);
            for (int j = 0; j <= WPT; j++) W[t][j] = w[j];
            hb_fsm_walk<WPT>(fsm, w, 0u, rec);
            {   /* the walks over bank-separated table copies (2 and 4 copies, this thread's copy index
                 * inside the spread bytes) read the same entries */
                uint32_t v1[2 * WPT], v2[2 * WPT], r1[WPT], r2[WPT];
                const uint32_t c1 = ((uint32_t)t & 1u) << 12, c2 = ((uint32_t)t & 3u) << 10;
                for (int j = 0; j < WPT; j++) {
                    hb_fsmc_spread<1>(w[j], c1 * 0x10001u, v1[2 * j], v1[2 * j + 1]);
                    hb_fsmc_spread<2>(w[j], c2 * 0x10001u, v2[2 * j], v2[2 * j + 1]);
                }
                hb_fsmc_walk<WPT, 1>(fsm, v1, 0u, r1);
                hb_fsmc_walk<WPT, 2>(fsm, v2, 0u, r2);
                for (int j = 0; j < WPT; j++) if (r1[j] != rec[j] || r2[j] != rec[j]) fsm_mismatch++;
            }
            for (int j = 0; j < WPT; j++) R[t][j] = rec[j];
            s_exit[t] = hb_frec_state(rec[WPT - 1]);
        }
        uint64_t rounds = 0;
        for (;;) {
            bool any = false;
            std::vector<uint32_t> new_exit(s_exit);
            for (int t = 1; t < T; t++) {
                const uint32_t sn = s_exit[t - 1];
                if (sn != st_in[t]) {
                    st_in[t] = sn;
                    uint32_t w[WPT + 1], rec[WPT];
                    for (int j = 0; j <= WPT; j++) w[j] = W[t][j];
                    for (int j = 0; j < WPT; j++) rec[j] = R[t][j];
                    {
                        uint32_t v2[2 * WPT], r2[WPT];
                        for (int j = 0; j < WPT; j++) {
                            hb_fsmc_spread<2>(w[j], (((uint32_t)t & 3u) << 10) * 0x10001u, v2[2 * j], v2[2 * j + 1]);
                            r2[j] = rec[j];
                        }
                        const bool ch2 = hb_fsmc_rewalk<WPT, 2>(fsm, v2, sn, r2);
                        const bool ch = hb_fsm_rewalk<WPT>(fsm, w, sn, rec);
                        if (ch != ch2) fsm_mismatch++;
                        for (int j = 0; j < WPT; j++) if (r2[j] != rec[j]) fsm_mismatch++;
                        if (ch) any = true;
                    }
                    for (int j = 0; j < WPT; j++) R[t][j] = rec[j];
                    new_exit[t] = hb_frec_state(rec[WPT - 1]);
                }
            }
            rounds++;
            if (!any) break;
            s_exit = new_exit;
        }
        st.rounds_total += rounds;
        st.rounds_max = std::max<uint64_t>(st.rounds_max, rounds);
        uint32_t E0 = 0, X0 = 0, d0 = 0;
        for (int t = 0; t < T; t++) {
            uint32_t ends = 0;
            for (int j = 0; j < WPT; j++) { ends += hb_frec_ends(R[t][j]); s_rec[j * T + t] = (uint16_t)R[t][j]; }
            const uint32_t d_in = fsm.depth[st_in[t]], d_out = fsm.depth[hb_frec_state(R[t][WPT - 1])];
            uint32_t e = 0;
            if (d_in) e = hb_fsm_fwd(tbS.slow, words[tbase + (uint64_t)t * WPT - 1], W[t][0], d_in);
            subs[(uint64_t)tile * T + t] = hb_sub_pack(e, ends - (d_in ? 1u : 0u) + (d_out ? 1u : 0u));
            if (t == T - 1) { X0 = hb_fsm_fwd(tbS.slow, W[t][WPT - 1], W[t][WPT], d_out); d0 = d_out; }
            s_cs[t] = E0;
            E0 += ends;
        }
        auto word = [&](uint32_t i) -> uint32_t { return tbase + i < nwords ? words[tbase + i] : 0u; };
        for (uint32_t t = 0; t < 32; t++) {
            uint32_t m = hb_map_pack32(X0, E0 + (d0 ? 1u : 0u));
            if (t > 0 && t < maxlen && possible(tile, t)) {
                m = hb_fsm_hyp_walk<WPT, T>(fsm, tbS.slow, word, s_rec.data(), s_cs.data(), E0, X0, d0, t);
                if ((m & 31u) != X0) st.hyp_unmerged++;
            }
            tmaps[(uint64_t)tile * 32 + t] = m;
        }
        fsm_tiles++;
    }

    /* dispatch of launch_map in hb_api.cu */
    void sync_all() {
        const bool odd_factor = (gmod & (gmod - 1u)) != 0u;      /* launch_map: probe kernel only */
        const bool use_fsm = sync_mode != 0 && have_fsm && minlen != maxlen && !odd_factor;
        const uint32_t n_full = use_fsm ? (uint32_t)(bits_own / TS) : 0u;
        for (uint32_t tile = 0; tile < ntiles; tile++) {
            if (tile < n_full) {
                sync_tile_fsm(tile);
                if (sync_mode == 2) {   /* the probe kernel must produce the very same records */
                    std::vector<uint16_t> a(subs.begin() + (size_t)tile * T, subs.begin() + (size_t)(tile + 1) * T);
                    std::vector<uint32_t> b(tmaps.begin() + (size_t)tile * 32, tmaps.begin() + (size_t)(tile + 1) * 32);
                    sync_tile(tile);
                    if (!std::equal(a.begin(), a.end(), subs.begin() + (size_t)tile * T)) fsm_mismatch++;
                    if (!std::equal(b.begin(), b.end(), tmaps.begin() + (size_t)tile * 32)) fsm_mismatch++;
                }
            } else sync_tile(tile);
        }
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
        uint64_t total = shard_map[E] >> 8;
        if (total && bits_own + (shard_map[E] & 31u) > bits_avail) total--;
        result[0] = total; result[1] = shard_map[E] & 31u; result[2] = E; result[3] = B;
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

    /* mirrors hb_fix_kernel, one tile */
    void fix_tile(uint32_t tile) {
        const uint32_t E = tile_entry[tile];
        if (E == 0) return;
        st.tiles_entry_nonzero++;
        const uint64_t own_left = bits_own - (uint64_t)tile * TS;
        const uint32_t tile_lim = own_left < TS ? (uint32_t)own_left : TS;
        const uint64_t base = (uint64_t)tile * (T * WPT);
        auto word = [&](uint32_t i) -> uint32_t { return base + i < nwords ? words[base + i] : 0u; };
        std::vector<uint16_t> before(subs.begin() + (size_t)tile * T, subs.begin() + (size_t)(tile + 1) * T);
        if (minlen == maxlen) {   /* mirrors hb_fix_fixed_kernel */
            for (uint32_t t = 0; t < (uint32_t)T; t++) {
                const uint32_t s0 = t * S;
                if (s0 >= tile_lim) break;
                const uint32_t s1 = tile_lim - s0 < S ? tile_lim : s0 + S;
                const uint32_t first = hb_fixed_next(E, maxlen, s0);
                subs[(size_t)tile * T + t] = hb_sub_pack((first - s0) & 31u, hb_fixed_count(E, maxlen, s0, s1));
            }
        } else
        hb_fix_entries<WPT, T>(tbS, word, subs.data() + (size_t)tile * T, tile_lim, E);
        for (int t = 0; t < T; t++)
            if (before[t] != subs[(size_t)tile * T + t]) st.probes_fix += hb_sub_count(subs[(size_t)tile * T + t]);
    }

    /* mirrors hb_emitf_kernel, one full tile that is not the last one */
    struct ColVec {
        const uint32_t *w; uint32_t n, k;
        uint32_t next() { return k < n ? w[k++] : (k++, 0xdeadbeefu); }
    };
    bool emit_tile_flat(uint32_t tile, uint8_t *out, uint64_t out_capacity, uint32_t win, uint32_t stage_bytes) {
        constexpr uint32_t LA = 6;
        const uint64_t B = tile_base[tile];
        uint32_t o_acc = 0;
        std::vector<uint32_t> off(T), cnt(T);
        for (int t = 0; t < T; t++) { off[t] = o_acc; cnt[t] = hb_sub_count(subs[(uint64_t)tile * T + t]); o_acc += cnt[t]; }
        const uint32_t nk = o_acc;
        if (B + nk > out_capacity) return false;
        hb_ptab tb{eptab.data(), 0u, 3u, ((1u << ep_wf) - 1u) << 3, tbE.slow};
        uint32_t lo_b = 0;
        for (uint32_t wb = 0; wb == 0 || wb < nk; wb += win) {
            std::vector<uint8_t> s_out(stage_bytes + 64, 0xEE);
            const uint32_t al = (uint32_t)((reinterpret_cast<uintptr_t>(out) + B + wb) & 15u);
            uint32_t hi_b = nk;
            struct Fin { uint8_t *at; uint32_t w; };
            std::vector<Fin> fins;
            for (int t = 0; t < T; t++) {
                const uint32_t o = off[t], c = cnt[t];
                const bool mine = c && o >= wb && o - wb < win;
                if (!mine) continue;
                const uint32_t e = hb_sub_entry(subs[(uint64_t)tile * T + t]);
                if (al + (o - wb) + c + 3 + 8 > stage_bytes) return false;      /* staging bound violated */
                uint32_t w[WPT + LA];
                for (uint32_t j = 0; j < WPT + LA; j++) {
                    const uint64_t i = (uint64_t)tile * (T * WPT) + (uint64_t)t * WPT + j;
                    w[j] = i < nwords ? words[i] : 0u;
                }
                ColVec col{w, WPT + LA, 0};
                uint8_t *dst = s_out.data() + al + (o - wb);
                const uint32_t mis = (al + (o - wb)) & 3u;
                uint8_t *wend = s_out.data() + ((al + (o - wb) + c + 3u) & ~3u);
                uint32_t fin = ep_wf <= 10 ? hb_emit_flat<3>(tb, col, e, dst, mis, wend)
                                           : hb_emit_flat<2>(tb, col, e, dst, mis, wend);
                if (col.k > WPT + LA) return false;                          /* read past the look-ahead rows */
                fins.push_back(Fin{wend - 4, fin});
                st.probes_emit += c;
                if (o + c - wb >= win && o + c < nk) hi_b = o + c;
            }
            /* worst store order: every slice's first word lands last (zeros below its first
             * byte) ... then the barrier, then the final words once more */
            for (int t = 0; t < T; t++) {
                const uint32_t o = off[t], c = cnt[t];
                if (!(c && o >= wb && o - wb < win)) continue;
                const uint32_t a0 = al + (o - wb);
                for (uint32_t i = a0 & ~3u; i < a0; i++) s_out[i] = 0;
            }
            for (const Fin &f : fins) for (int i = 0; i < 4; i++) f.at[i] = (uint8_t)(f.w >> (8 * i));
            if (lo_b < hi_b) {
                uint8_t *gbase = out + B + wb - al;
                const uint32_t begb = al + (lo_b - wb), endb = al + (hi_b - wb);
                for (uint32_t i = begb; i < endb; i++) gbase[i] = s_out[i];
            }
            lo_b = hi_b > lo_b ? hi_b : lo_b;
        }
        flat_tiles++;
        return true;
    }

    /* mirrors hb_emit32w_kernel, one tile: every group of 32 lanes (a warp) on its own, SPL consecutive
     * subsequences per lane, with its own output base (tile base + the symbols in front of it), windows
     * and copy-out */
    template <int SPL>
    bool emit_tile_warp(uint32_t tile, uint8_t *out, uint64_t out_capacity, uint32_t win, uint32_t stage_bytes,
                        uint64_t total_valid) {
        const uint64_t tile_bit0 = (uint64_t)tile * TS;
        constexpr int L = T / SPL >= 32 ? 32 : T / SPL;      /* lanes per warp tile */
        uint32_t front = 0;
        for (int q = 0; q < T / (L * SPL); q++) {
            const int t0 = q * L * SPL;
            const uint64_t B = tile_base[tile] + front;
            uint32_t o_acc = 0;
            std::vector<uint32_t> off(L), cnt(L), cnt0(L);
            for (int l = 0; l < L; l++) {
                off[l] = o_acc;
                cnt0[l] = hb_sub_count(subs[(uint64_t)tile * T + t0 + l * SPL]);
                cnt[l] = cnt0[l] + (SPL == 2 ? hb_sub_count(subs[(uint64_t)tile * T + t0 + l * SPL + 1]) : 0u);
                o_acc += cnt[l];
            }
            const uint32_t nk = o_acc;
            front += nk;
            uint32_t nvalid = nk;
            if (B >= total_valid) nvalid = 0;
            else if (B + nk > total_valid) nvalid = (uint32_t)(total_valid - B);
            if (B + nvalid > out_capacity) return false;
            uint32_t lo_b = 0;
            for (uint32_t wb = 0; wb == 0 || wb < nk; wb += win) {
                std::vector<uint8_t> s_out(stage_bytes + 64, 0xEE);
                const uint32_t al = (uint32_t)((reinterpret_cast<uintptr_t>(out) + B + wb) & 15u);
                uint32_t hi_b = nk;
                std::vector<hb_tail> tails;
                for (int l = 0; l < L; l++) {
                    const int t = t0 + l * SPL;
                    const uint32_t o = off[l], c = cnt[l], c0 = cnt0[l];
                    const bool mine = c && o >= wb && o - wb < win;
                    if (!mine) continue;
                    const uint64_t sub0 = tile_bit0 + (uint64_t)t * S;
                    const uint32_t lim = sub0 >= bits_own ? 0u
                                       : (bits_own - sub0 < SPL * S ? (uint32_t)(bits_own - sub0) : SPL * S);
                    const uint32_t e = hb_sub_entry(subs[(uint64_t)tile * T + t]);
                    if (al + (o - wb) + c > stage_bytes) return false;      /* staging bound violated */
                    uint32_t w[WPT + 1];
                    load((uint64_t)tile * (T * WPT) + (uint64_t)t * WPT, w);
                    uint8_t *dst = s_out.data() + al + (o - wb);
                    const uint8_t canary = dst[c];
                    const uint32_t mis = (al + (o - wb)) & 3u;
                    uint32_t n = c;
                    if (lim != SPL * S || WPT < 2) {
                        n = hb_emit_clipped32<WPT>(tbE32, w, lim < S ? lim : S, e, c0, dst);
                        if (SPL == 2) {
                            load((uint64_t)tile * (T * WPT) + (uint64_t)(t + 1) * WPT, w);
                            n += hb_emit_clipped32<WPT>(tbE32, w, lim > S ? lim - S : 0u,
                                                        hb_sub_entry(subs[(uint64_t)tile * T + t + 1]), c - c0, dst + c0);
                        }
                    } else if constexpr (WPT >= 2) {
                        hb_w32 st;
                        hb_w32_begin(st, e, dst, mis);
                        for (int h = 0; h < SPL; h++) {
                            /* the chain runs on (its last probe may already have taken codewords that start
                             * in the second subsequence: st.acc is a probe position, not the recorded entry) */
                            if (h) load((uint64_t)tile * (T * WPT) + (uint64_t)(t + 1) * WPT, w);
                            hb_emit_words32_part<WPT>(tbE32, w, st, c, dst, h == SPL - 1);
                        }
                        tails.push_back(hb_w32_tail(st));
                        if ((uint32_t)((tails.back().at + tails.back().k) - dst) != c) return false;
                    }
                    if (n != c) return false;
                    if (dst[c] != canary) return false;
                    st.probes_emit += n;
                    if (o + c - wb >= win && o + c < nk) hi_b = o + c;
                }
                for (const hb_tail &tl : tails) hb_store_tail(tl);
                if (hi_b > nvalid) hi_b = nvalid;
                if (lo_b < hi_b) {
                    uint8_t *gbase = out + B + wb - al;
                    const uint32_t begb = al + (lo_b - wb), endb = al + (hi_b - wb);
                    for (uint32_t i = begb; i < endb; i++) gbase[i] = s_out[i];
                }
                lo_b = hi_b > lo_b ? hi_b : lo_b;
            }
        }
        return true;
    }

    /* mirrors hb_emit_kernel, one tile; returns false on output overflow */
    bool emit_tile(uint32_t tile, uint8_t *out, uint64_t out_capacity, uint32_t win, uint32_t stage_bytes,
                   uint64_t total_valid) {
        const uint64_t tile_bit0 = (uint64_t)tile * TS;
        const uint64_t B = tile_base[tile];
        uint32_t o_acc = 0;
        std::vector<uint32_t> off(T), cnt(T), trips(T, 0);
        for (int t = 0; t < T; t++) { off[t] = o_acc; cnt[t] = hb_sub_count(subs[(uint64_t)tile * T + t]); o_acc += cnt[t]; }
        const uint32_t nk = o_acc;
        uint32_t nvalid = nk;
        if (B >= total_valid) nvalid = 0;
        else if (B + nk > total_valid) nvalid = (uint32_t)(total_valid - B);
        if (B + nvalid > out_capacity) return false;
        uint32_t lo_b = 0;
        for (uint32_t wb = 0; wb == 0 || wb < nk; wb += win) {
            std::vector<uint8_t> s_out(stage_bytes + 64, 0xEE);
            const uint32_t al = (uint32_t)((reinterpret_cast<uintptr_t>(out) + B + wb) & 15u);
            uint32_t hi_b = nk;
            std::vector<hb_tail> tails;   /* hb_emitw_kernel: stored after the barrier that ends the decode */
            for (int t = 0; t < T; t++) {
                const uint32_t o = off[t], c = cnt[t];
                const bool mine = c && o >= wb && o - wb < win;
                if (!mine) continue;
                const uint64_t sub0 = tile_bit0 + (uint64_t)t * S;
                const uint32_t lim = sub0 >= bits_own ? 0u : (bits_own - sub0 < S ? (uint32_t)(bits_own - sub0) : S);
                const uint32_t e = hb_sub_entry(subs[(uint64_t)tile * T + t]);
                if (al + (o - wb) + c > stage_bytes) return false;      /* staging bound violated */
                uint32_t w[WPT + 1];
                load((uint64_t)tile * (T * WPT) + (uint64_t)t * WPT, w);
                uint8_t *dst = s_out.data() + al + (o - wb);
                const uint8_t canary = dst[c];
                /* the staging buffer is 16-byte aligned on the device: only the index counts */
                const uint32_t mis = (al + (o - wb)) & 3u;
                uint32_t n = c;
                if (lim != S && (emit_mode == 0 || WPT < 2)) n = hb_emit_slow<WPT>(tbE.slow, w, lim, e, c, dst);
                else if (lim != S) n = emit_mode == 3 ? hb_emit_clipped32<WPT>(tbE32, w, lim, e, c, dst)
                                                      : hb_emit_clipped<WPT>(tbE64, w, lim, e, c, dst);
                else if (emit_mode == 0 || WPT < 2) n = hb_emit_fast<WPT>(tbE, w, e, c, dst);
                else if constexpr (WPT >= 2) {
                    tails.push_back(emit_mode == 3 ? hb_emit_words32<WPT>(tbE32, w, e, c, dst, mis)
                                                   : hb_emit_words<WPT>(tbE64, w, e, c, dst, mis));
                    /* whole words inside the slice plus the tail make exactly c bytes */
                    if ((uint32_t)((tails.back().at + tails.back().k) - dst) != c) return false;
                }
                if (n != c) return false;                                /* record/chain mismatch */
                if (dst[c] != canary) return false;                      /* wrote past its slice */
                st.probes_emit += n;
                trips[t] = n;
                if (o + c - wb >= win && o + c < nk) hi_b = o + c;
            }
            for (const hb_tail &tl : tails) hb_store_tail(tl);
            if (hi_b > nvalid) hi_b = nvalid;
            if (lo_b < hi_b) {
                uint8_t *gbase = out + B + wb - al;
                const uint32_t begb = al + (lo_b - wb), endb = al + (hi_b - wb);
                const uint32_t v0 = begb >> 4, nvec = (endb + 15u) >> 4;
                for (uint32_t v = v0; v < nvec; v++) {
                    const uint32_t b0 = v << 4;
                    if (b0 >= begb && b0 + 16u <= endb) memcpy(gbase + b0, s_out.data() + b0, 16);
                    else {
                        const uint32_t lo = b0 < begb ? begb : b0, hi = b0 + 16u < endb ? b0 + 16u : endb;
                        for (uint32_t i = lo; i < hi; i++) gbase[i] = s_out[i];
                    }
                }
            }
            lo_b = hi_b > lo_b ? hi_b : lo_b;
        }
        for (int w0 = 0; w0 < T; w0 += 32)
            st.warp_iters_emit += *std::max_element(trips.begin() + w0, trips.begin() + std::min(T, w0 + 32));
        return true;
    }
};

template <int WPT, int T>
static int run(const uint32_t *lut_entries, uint32_t w1, uint32_t maxlen, uint32_t minlen,
               const uint32_t *stab, const uint32_t *etab, uint32_t wf,
               const uint32_t *words, uint64_t nwords, uint64_t bits_own, uint64_t bits_avail,
               int have_entry, uint32_t entry, uint64_t base, uint8_t *out, uint64_t out_capacity,
               uint64_t *shard_map, uint64_t *result, emul_stats *stats, uint32_t emit_win,
               int sync_mode, uint32_t fsm_states, const uint16_t *fsm_tab, const uint8_t *fsm_depth,
               const uint16_t *fsm_pstep, int emit_mode, const uint32_t *e64, uint32_t wf64, uint32_t ep_wf,
               uint32_t gmod, uint32_t gorg) {
    Emul<WPT, T> E;
    E.gmod = gmod ? gmod : 1u;
    E.gorg = gorg;
    E.emit_mode = emit_mode == 2 ? 1 : emit_mode;
    E.flat = emit_mode == 2 && WPT >= 4;
    E.sync_mode = sync_mode;
    E.have_fsm = fsm_states != 0;
    E.fsm = hb_fsm{fsm_tab, 0u, fsm_depth, fsm_pstep};
    memset(&E.st, 0, sizeof(E.st));
    E.words = words; E.nwords = nwords; E.bits_own = bits_own; E.bits_avail = bits_avail;
    hb_lutref slow{lut_entries, lut_entries, (1u << w1) - 1u};
    E.tbS = hb_tables{stab, 0u, ((1u << wf) - 1u) << 2, slow};
    E.tbE = hb_tables{etab, 0u, ((1u << wf) - 1u) << 2, slow};
    E.tbE64 = hb_tables64{e64, 0u, ((1u << wf64) - 1u) << 3, slow, 3u, 0u};
    if (emit_mode >= 3 && emit_mode <= 5) {   /* E32-table exactly as hb_emit32_kernel builds it (index width: ep_wf, else wf64) */
        uint32_t wf32 = ep_wf ? ep_wf : wf64;
        if (wf32 > maxlen && maxlen >= 9u) wf32 = maxlen;
        E.e32tab.resize((size_t)1 << wf32);
        for (uint32_t x = 0; x < (1u << wf32); x++) E.e32tab[x] = hb_e32_entry(slow, x, wf32);
        E.tbE32 = hb_tables32{E.e32tab.data(), ((1u << wf32) - 1u) << 2, 0u, slow, 2u, wf32};
    }
    E.maxlen = maxlen; E.minlen = minlen;
    const uint64_t tile_bits = (uint64_t)E.TS;
    E.ntiles = (uint32_t)((bits_own + tile_bits - 1) / tile_bits);
    E.subs.assign((size_t)E.ntiles * T, 0);
    E.tmaps.assign((size_t)E.ntiles * 32, 0);
    E.st.tiles = E.ntiles;
    E.sync_all();
    E.scan_up();
    if (E.ntiles == 0) for (int e = 0; e < 32; e++) E.shard_map[e] = (uint64_t)e;
    if (shard_map) memcpy(shard_map, E.shard_map, sizeof(E.shard_map));
    int rc = 0;
    if (have_entry) {
        uint64_t res[4] = { 0, entry, entry, base };
        if (E.ntiles) E.scan_down(entry & 31u, base, res);
        for (uint32_t tile = 0; tile < E.ntiles; tile++) E.fix_tile(tile);
        uint32_t S = 32u * WPT;
        uint32_t max_c = (S + minlen - 1) / minlen;
        uint32_t win = emit_win ? emit_win : ((T * max_c + 15u) & ~15u);   /* 0 = worst case, one window */
        if (win < ((max_c + 15u) & ~15u)) win = (max_c + 15u) & ~15u;      /* a window holds at least one thread's output */
        uint32_t stage = (win + max_c + 16u + 15u) & ~15u;
        if (E.flat) {   /* EP-table exactly as the kernel builds it */
            E.ep_wf = ep_wf ? ep_wf : 10u;
            E.eptab.resize((size_t)2 << E.ep_wf);
            for (uint32_t x = 0; x < (1u << E.ep_wf); x++) hb_ep_entry(slow, x, E.ep_wf, &E.eptab[2 * x], &E.eptab[2 * x + 1]);
        }
        for (uint32_t tile = 0; tile < E.ntiles; tile++) {
            const bool flat = E.flat && tile + 1 < E.ntiles;
            if (emit_mode == 4 || emit_mode == 5) {    /* warp tiles: a window sized for 32 lanes' output */
                const uint32_t spl = emit_mode == 5 && T >= 2 ? 2u : 1u;
                const uint32_t L = T / spl >= 32 ? 32u : (uint32_t)T / spl;
                const uint32_t mc = spl * max_c;
                uint32_t ww = emit_win ? emit_win : ((L * mc + 15u) & ~15u);
                if (ww < ((mc + 15u) & ~15u)) ww = (mc + 15u) & ~15u;
                const uint32_t stg = (ww + mc + 32u + 15u) & ~15u;
                const bool ok = spl == 2 ? E.template emit_tile_warp<2>(tile, out, out_capacity, ww, stg, res[0])
                                         : E.template emit_tile_warp<1>(tile, out, out_capacity, ww, stg, res[0]);
                if (!ok) { rc = -6; break; }
                continue;
            }
            if (flat ? !E.emit_tile_flat(tile, out, out_capacity, win, stage + 16)
                     : !E.emit_tile(tile, out, out_capacity, win, stage, res[0])) { rc = -6; break; }
        }
        if (E.flat && E.ntiles > 1 && E.flat_tiles != E.ntiles - 1) rc = -103;
        if (result) memcpy(result, res, sizeof(res));
    }
    if (E.st.long_probes >> 63) rc = -100;   /* fast and slow word walks disagreed */
    if (E.fsm_mismatch) rc = -101;           /* transducer and probe sync kernels disagreed */
    if (sync_mode && E.have_fsm && minlen != maxlen && !(E.gmod & (E.gmod - 1u)) && E.fsm_tiles != bits_own / E.TS) rc = -102;
    if (stats) *stats = E.st;
    return rc;
}

extern "C" int emul_run(const uint32_t *lut_entries, uint32_t w1, uint32_t maxlen, uint32_t minlen,
                        const uint32_t *stab, const uint32_t *etab, uint32_t wf,
                        const uint32_t *words, uint64_t nwords, uint64_t bits_own,
                        uint64_t bits_avail, int wpt, int T, int have_entry, uint32_t entry,
                        uint64_t base, uint8_t *out, uint64_t out_capacity, uint64_t *shard_map,
                        uint64_t *result, emul_stats *stats, uint32_t emit_win,
                        int sync_mode, uint32_t fsm_states, const uint16_t *fsm_tab,
                        const uint8_t *fsm_depth, const uint16_t *fsm_pstep, int emit_mode,
                        const uint32_t *e64, uint32_t wf64, uint32_t ep_wf, uint32_t gmod, uint32_t gorg) {
#define CASE(W, TT)                                                                              \
    if (wpt == W && T == TT)                                                                     \
        return run<W, TT>(lut_entries, w1, maxlen, minlen, stab, etab, wf, words, nwords,        \
                          bits_own, bits_avail, have_entry, entry, base, out, out_capacity,      \
                          shard_map, result, stats, emit_win, sync_mode, fsm_states, fsm_tab,    \
                          fsm_depth, fsm_pstep, emit_mode, e64, wf64, ep_wf, gmod, gorg)
    CASE(4, 256); CASE(8, 256); CASE(16, 256);
    CASE(1, 4); CASE(2, 8); CASE(4, 32); CASE(1, 64);
#undef CASE
    return -4;
}
