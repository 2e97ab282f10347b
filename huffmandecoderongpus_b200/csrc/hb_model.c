/*
 * hb_model.c -- bundled synthetic-stream generator, CPU side (plain C).
 *
 * The reference ships no encoder; its .huff files were made by
 * an external tool.  This file defines the synthetic streams of BASELINE.json
 * configs 4 and 5: a symbol model (weights -> Huffman tree in the reference's
 * node format + sampling thresholds), a counter-based generator (symbol i is a
 * pure function of (model, seed, i), so any range can be regenerated on any
 * device for verification) and a bit packer that follows the format's bit
 * order (LSB-first within a byte, framework/mainrun.c:45).
 * hb_gen.cu is the GPU twin and must produce bit-identical streams.
 */
#include "huffb200.h"
#include "hb_format.h"
#include "hb_english_hist.h"

#include <stdlib.h>
#include <string.h>

/* ---- weights -------------------------------------------------------------- */

static void model_weights(int kind, uint64_t w[256]) {
    memset(w, 0, sizeof(uint64_t) * 256);
    switch (kind) {
    case HB_MODEL_ENGLISH:
        for (int s = 0; s < 256; s++) w[s] = hb_english_hist[s];
        break;
    case HB_MODEL_FIBONACCI: {
        /* w[s] = F(40 - 1 - s) for s < 40, 1 beyond: the 217 weight-1 symbols
         * form a deep subtree half way down the Fibonacci spine, giving a
         * maximum code length in (20, 32] (asserted in hb_model_build) */
        uint64_t f[64];
        f[0] = 1; f[1] = 1;
        for (int i = 2; i < 64; i++) f[i] = f[i - 1] + f[i - 2];
        for (int s = 0; s < 256; s++) w[s] = f[s < 39 ? 39 - s : 0];
        break;
    }
    case HB_MODEL_DNA:
        w['a'] = w['c'] = w['g'] = w['t'] = 1;
        break;
    case HB_MODEL_UNIFORM8:
        for (int s = 0; s < 8; s++) w['A' + s] = 1;
        break;
    default:
        break;
    }
}

/* ---- Huffman tree ----------------------------------------------------------
 * Deterministic: repeatedly join the two lightest roots; ties go to the one
 * created first (leaves in symbol order, then internal nodes in creation
 * order).  The lighter root becomes the 0-child. */

typedef struct bnode {
    uint64_t w;
    int left, right; /* -1 for leaves */
    int sym;
} bnode;

static int emit_tree(const bnode *bn, int root, hb_node_abi *tree) {
    /* breadth-first numbering, root = node 0 (reference convention) */
    int queue[511], idx_of[511], head = 0, tail = 0, n = 0;
    queue[tail++] = root;
    idx_of[root] = n++;
    while (head < tail) {
        int v = queue[head++];
        if (bn[v].left >= 0) {
            idx_of[bn[v].left] = n++;  queue[tail++] = bn[v].left;
            idx_of[bn[v].right] = n++; queue[tail++] = bn[v].right;
        }
    }
    for (int i = 0; i < tail; i++) {
        int v = queue[i];
        hb_node_abi *o = &tree[idx_of[v]];
        if (bn[v].left < 0) {
            o->sym = (uint8_t)bn[v].sym; o->izero = -1; o->ione = -1;
        } else {
            o->sym = 0; o->izero = idx_of[bn[v].left]; o->ione = idx_of[bn[v].right];
        }
    }
    return n;
}

static void walk_codes(const hb_node_abi *tree, int v, int depth, uint32_t bits, hb_model *m) {
    if (tree[v].izero == -1) {
        m->code[tree[v].sym] = bits;
        m->codelen[tree[v].sym] = (uint8_t)depth;
        if ((uint32_t)depth > m->maxlen) m->maxlen = (uint32_t)depth;
        if ((uint32_t)depth < m->minlen) m->minlen = (uint32_t)depth;
        return;
    }
    walk_codes(tree, tree[v].izero, depth + 1, bits, m);
    walk_codes(tree, tree[v].ione, depth + 1, bits | (1u << depth), m);
}

int hb_model_build(int kind, hb_model *m) {
    if (!m || kind < 0 || kind > HB_MODEL_UNIFORM8) return HB_ERR_ARG;
    memset(m, 0, sizeof(*m));
    uint64_t w[256];
    model_weights(kind, w);
    bnode bn[511];
    int alive[256], na = 0, nb = 0;
    uint64_t total = 0;
    for (int s = 0; s < 256; s++) {
        if (!w[s]) continue;
        bn[nb].w = w[s]; bn[nb].left = bn[nb].right = -1; bn[nb].sym = s;
        alive[na++] = nb++;
        total += w[s];
    }
    if (na < 2 || total >= (1ull << 32)) return HB_ERR_ARG;
    m->nsyms = (uint32_t)na;
    while (na > 1) {
        int a = -1, b = -1; /* positions in alive[] of the two lightest */
        for (int i = 0; i < na; i++) {
            int v = alive[i];
            if (a < 0 || bn[v].w < bn[alive[a]].w || (bn[v].w == bn[alive[a]].w && v < alive[a])) {
                b = a; a = i;
            } else if (b < 0 || bn[v].w < bn[alive[b]].w ||
                       (bn[v].w == bn[alive[b]].w && v < alive[b])) {
                b = i;
            }
        }
        bn[nb].w = bn[alive[a]].w + bn[alive[b]].w;
        bn[nb].left = alive[a]; bn[nb].right = alive[b]; bn[nb].sym = 0;
        int hi = a > b ? a : b, lo = a > b ? b : a;
        alive[hi] = alive[--na];
        alive[lo] = nb++;
    }
    m->nodes = emit_tree(bn, alive[0], m->tree);
    m->minlen = 0xffffffffu;
    walk_codes(m->tree, 0, 0, 0, m);
    if (m->maxlen > HB_MAX_CODELEN) return HB_ERR_CODELEN;
    if (kind == HB_MODEL_FIBONACCI && !(m->maxlen > 20 && m->maxlen <= 32)) return HB_ERR_CODELEN;
    /* sampling thresholds over the present symbols, in symbol order:
     * cum[k] = floor(2^32 * (weight of the present symbols before the k-th) / total) */
    uint64_t pre = 0;
    int k = 0;
    for (int s = 0; s < 256; s++) {
        if (!w[s]) continue;
        m->cum[k] = (uint32_t)((pre << 32) / total);
        m->symtab[k] = (uint8_t)s;
        pre += w[s];
        k++;
    }
    for (; k < 256; k++) { m->cum[k] = 0xffffffffu; m->symtab[k] = m->symtab[m->nsyms - 1]; }
    return HB_OK;
}

/* ---- generator ------------------------------------------------------------ */

static inline uint32_t gen_u32(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull; /* splitmix64 */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

static inline uint32_t sample(const uint32_t *cum, uint32_t nsyms, uint32_t u) {
    /* largest k in [0,nsyms) with cum[k] <= u (cum is non-decreasing, cum[0] = 0) */
    uint32_t lo = 0, hi = nsyms - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (cum[mid] <= u) lo = mid; else hi = mid - 1;
    }
    return lo;
}

void hb_gen_symbols_cpu(const hb_model *m, uint64_t seed, uint64_t first, uint64_t n, uint8_t *out) {
    for (uint64_t i = 0; i < n; i++) out[i] = m->symtab[sample(m->cum, m->nsyms, gen_u32(seed, first + i))];
}

uint64_t hb_encode_bits_cpu(const hb_model *m, const uint8_t *syms, uint64_t n) {
    uint64_t bits = 0;
    for (uint64_t i = 0; i < n; i++) bits += m->codelen[syms[i]];
    return bits;
}

void hb_encode_cpu(const hb_model *m, const uint8_t *syms, uint64_t n, uint8_t *out) {
    uint64_t acc = 0; /* pending bits, LSB first */
    int nacc = 0;
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; i++) {
        acc |= (uint64_t)m->code[syms[i]] << nacc;
        nacc += m->codelen[syms[i]];
        while (nacc >= 8) {
            out[o++] = (uint8_t)acc;
            acc >>= 8;
            nacc -= 8;
        }
    }
    if (nacc) out[o++] = (uint8_t)acc;
}
