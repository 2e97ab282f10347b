/*
 * hb_lut.c -- Huffman tree -> multi-level lookup table (host, plain C).
 *
 * The reference decodes by walking the node array bit by bit on the device
 * (framework/fastgpu.cu:46-66) and only has single-level 2^height tables on
 * the CPU (framework/mainrun.c:142-195, capped at height 23 by mask2).  Here
 * the first level is W1 <= 11 bits wide so that it fits shared memory; longer
 * codes chain through sub-tables of at most W2 bits (world192 has 20-bit
 * codes, the adversarial config up to 32).
 */
#include "hb_lut.h"
#include "hb_format.h"

#include <stdlib.h>
#include <string.h>

#define HB_W1_DEFAULT 11
#define HB_W2_DEFAULT 10
#define HB_MAX_ENTRIES (1u << 18)

typedef struct builder {
    const hb_node *tree;
    int nodes;
    int w1, w2;
    uint32_t *ent;
    uint32_t n, cap;
    int32_t *sub_of_node; /* memoised sub-table base per internal node, -1 = none */
    uint8_t *height;      /* subtree height per node */
    int err;
} builder;

static int is_leaf(const hb_node *n) { return n->izero == -1 && n->ione == -1; }
static int build_fast_tables(const hb_node *tree, hb_lut *out);
static int build_fsm(const hb_node *tree, int nodes, hb_lut *out, int with_table);

/* iterative validation: every reachable node is a full internal node or a leaf,
 * indices in range, no node reached twice (=> a tree, no cycles), depth <= 32 */
static int validate(const hb_node *tree, int nodes, uint8_t *height,
                    uint32_t *maxlen, uint32_t *minlen, uint32_t *nleaves) {
    if (nodes < 3 || !tree) return HB_ERR_TREE;
    if (is_leaf(&tree[0])) return HB_ERR_TREE;
    uint8_t *seen = (uint8_t *)calloc((size_t)nodes, 1);
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nodes + 16);
    uint8_t *depth = (uint8_t *)calloc((size_t)nodes, 1);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    if (!seen || !stack || !depth || !order) {
        free(seen); free(stack); free(depth); free(order);
        return HB_ERR_NOMEM;
    }
    int rc = HB_OK, sp = 0, no = 0;
    uint32_t mx = 0, mn = 0xffffffffu, nl = 0;
    stack[sp++] = 0;
    seen[0] = 1;
    while (sp > 0 && rc == HB_OK) {
        int32_t v = stack[--sp];
        order[no++] = v;
        const hb_node *nd = &tree[v];
        if (is_leaf(nd)) {
            uint32_t d = depth[v];
            if (d > mx) mx = d;
            if (d < mn) mn = d;
            nl++;
            continue;
        }
        int32_t ch[2] = { nd->izero, nd->ione };
        for (int k = 0; k < 2; k++) {
            int32_t c = ch[k];
            if (c <= 0 || c >= nodes || seen[c]) { rc = HB_ERR_TREE; break; }
            if (depth[v] + 1 > HB_MAX_CODELEN) { rc = HB_ERR_CODELEN; break; }
            seen[c] = 1;
            depth[c] = (uint8_t)(depth[v] + 1);
            stack[sp++] = c;
        }
    }
    if (rc == HB_OK) {
        /* heights bottom-up: children appear after parents in `order` */
        for (int i = no - 1; i >= 0; i--) {
            int32_t v = order[i];
            if (is_leaf(&tree[v])) height[v] = 0;
            else {
                uint8_t a = height[tree[v].izero], b = height[tree[v].ione];
                height[v] = (uint8_t)(1 + (a > b ? a : b));
            }
        }
        *maxlen = mx; *minlen = mn; *nleaves = nl;
    }
    free(seen); free(stack); free(depth); free(order);
    return rc;
}

static uint32_t alloc_table(builder *b, int width) {
    uint32_t need = 1u << width;
    if (b->n + need > HB_MAX_ENTRIES) { b->err = HB_ERR_NOMEM; return 0; }
    if (b->n + need > b->cap) {
        uint32_t nc = b->cap ? b->cap : 4096;
        while (nc < b->n + need) nc *= 2;
        uint32_t *ne = (uint32_t *)realloc(b->ent, sizeof(uint32_t) * nc);
        if (!ne) { b->err = HB_ERR_NOMEM; return 0; }
        b->ent = ne;
        b->cap = nc;
    }
    uint32_t base = b->n;
    b->n += need;
    return base;
}

static uint32_t build_table(builder *b, int32_t root, int width);

/* fill all slots of table `base` (index width `width`) whose low `depth` bits
 * equal `prefix` with what the walk from `node` finds */
static void fill(builder *b, uint32_t base, int width, int32_t node, int depth,
                 uint32_t prefix) {
    if (b->err) return;
    const hb_node *nd = &b->tree[node];
    if (is_leaf(nd)) {
        uint32_t e = (uint32_t)depth | ((uint32_t)nd->sym << 8);
        for (uint32_t i = prefix; i < (1u << width); i += (1u << depth))
            b->ent[base + i] = e;
        return;
    }
    if (depth == width) {
        int nw = b->height[node] < b->w2 ? b->height[node] : b->w2;
        uint32_t sub = (b->sub_of_node[node] >= 0) ? (uint32_t)b->sub_of_node[node]
                                                  : build_table(b, node, nw);
        if (b->err) return;
        b->ent[base + prefix] = HB_LUT_LINK | (uint32_t)width | ((uint32_t)nw << 8) |
                                (sub << 13);
        return;
    }
    fill(b, base, width, nd->izero, depth + 1, prefix);
    fill(b, base, width, nd->ione, depth + 1, prefix | (1u << depth));
}

static uint32_t build_table(builder *b, int32_t root, int width) {
    uint32_t base = alloc_table(b, width);
    if (b->err) return 0;
    b->sub_of_node[root] = (int32_t)base;
    fill(b, base, width, root, 0, 0);
    return base;
}

static void collect_codes(const hb_node *tree, int32_t node, int depth,
                          uint32_t bits, hb_lut *out, uint8_t *have) {
    const hb_node *nd = &tree[node];
    if (is_leaf(nd)) {
        if (!have[nd->sym]) {
            have[nd->sym] = 1;
            out->code[nd->sym] = bits;
            out->codelen[nd->sym] = (uint8_t)depth;
        }
        return;
    }
    collect_codes(tree, nd->izero, depth + 1, bits, out, have);
    collect_codes(tree, nd->ione, depth + 1, bits | (depth < 32 ? (1u << depth) : 0u), out, have);
}

static int lut_build(const hb_node *tree, int nodes, int w1_max, int w2_max, hb_lut *out, int small);

int hb_lut_build(const hb_node *tree, int nodes, int w1_max, int w2_max, hb_lut *out) {
    return lut_build(tree, nodes, w1_max, w2_max, out, 0);
}

int hb_lut_build_small(const hb_node *tree, int nodes, int w1_max, int w2_max, hb_lut *out) {
    return lut_build(tree, nodes, w1_max, w2_max, out, 1);
}

static int lut_build(const hb_node *tree, int nodes, int w1_max, int w2_max, hb_lut *out, int small) {
    if (!out) return HB_ERR_ARG;
    memset(out, 0, sizeof(*out));
    if (nodes > (1 << 20)) return HB_ERR_TREE;
    builder b;
    memset(&b, 0, sizeof(b));
    b.tree = tree;
    b.nodes = nodes;
    b.height = (uint8_t *)calloc((size_t)(nodes > 0 ? nodes : 1), 1);
    if (!b.height) return HB_ERR_NOMEM;
    int rc = validate(tree, nodes, b.height, &out->maxlen, &out->minlen, &out->n_leaves);
    if (rc != HB_OK) { free(b.height); return rc; }
    b.w1 = (w1_max > 0 ? w1_max : HB_W1_DEFAULT);
    b.w2 = (w2_max > 0 ? w2_max : HB_W2_DEFAULT);
    if (b.w1 > 13) b.w1 = 13;
    if (b.w2 > 12) b.w2 = 12;
    if ((uint32_t)b.w1 > out->maxlen) b.w1 = (int)out->maxlen;
    b.sub_of_node = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    if (!b.sub_of_node) { free(b.height); return HB_ERR_NOMEM; }
    for (int i = 0; i < nodes; i++) b.sub_of_node[i] = -1;
    build_table(&b, 0, b.w1);
    free(b.sub_of_node);
    free(b.height);
    if (b.err) { free(b.ent); return b.err; }
    out->entries = b.ent;
    out->n_entries = b.n;
    out->w1 = (uint32_t)b.w1;
    uint8_t have[256];
    memset(have, 0, sizeof(have));
    collect_codes(tree, 0, 0, 0, out, have);
    out->wf = HB_WF_MAX;
    {   /* mean codeword length under the code's own implied distribution (iterative DFS) */
        double acc = 0.0;
        uint32_t len_gcd = 0;
        int32_t st_node[2 * HB_MAX_CODELEN + 4];
        int st_depth[2 * HB_MAX_CODELEN + 4];
        int sp = 1;
        st_node[0] = 0; st_depth[0] = 0;
        while (sp > 0) {
            const int32_t v = st_node[--sp];
            const int d = st_depth[sp];
            if (is_leaf(&tree[v])) {
                acc += (double)d / (double)(1ull << d);
                uint32_t a = len_gcd, b = (uint32_t)d;     /* gcd of all codeword lengths */
                while (b) { const uint32_t r = a % b; a = b; b = r; }
                len_gcd = a;
                continue;
            }
            st_node[sp] = tree[v].izero; st_depth[sp++] = d + 1;
            st_node[sp] = tree[v].ione;  st_depth[sp++] = d + 1;
        }
        out->implied_avg_len = acc;
        out->len_gcd = len_gcd ? len_gcd : 1u;
        /* measured: fib4g (mean 3.0 bits) emits 7 % faster through an 11-bit table (half the
         * shared memory: 4 CTAs per SM instead of 3), english1g (4.26) 11 % slower */
        out->wf64 = acc <= 3.5 ? HB_WF_MAX - 1 : HB_WF_MAX;
    }
    int frc = small ? HB_OK : build_fast_tables(tree, out);
    if (frc != HB_OK) { hb_lut_free(out); return frc; }
    frc = build_fsm(tree, nodes, out, !small);
    if (frc != HB_OK) { hb_lut_free(out); return frc; }
    return HB_OK;
}

/* ---- byte-step transducer -----------------------------------------------------
 * The reference's CPU jump table (framework/jumptableapproach.c:40-99) keys its rows
 * by the partial-codeword prefix and emits symbols; this one keys them by the
 * internal tree node, consumes exactly 8 bits per step and only counts the codewords
 * that end inside them, which is all the sync kernel needs. */
static int build_fsm(const hb_node *tree, int nodes, hb_lut *out, int with_table) {
    out->fsm_states = 0;
    out->node_state = NULL;
    for (int i = 0; i < 256; i++) out->fsm_node[i] = -1;
    out->fsm = NULL;
    out->fsm_bstep = NULL;
    memset(out->fsm_depth, 0, sizeof(out->fsm_depth));
    /* number the internal nodes breadth first from the root */
    int32_t *state_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    int32_t *node_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    if (!state_of || !node_of) { free(state_of); free(node_of); return HB_ERR_NOMEM; }
    for (int i = 0; i < nodes; i++) state_of[i] = -1;
    uint32_t ns = 0;
    uint8_t *depth = (uint8_t *)calloc((size_t)nodes, 1);
    if (!depth) { free(state_of); free(node_of); return HB_ERR_NOMEM; }
    state_of[0] = 0; node_of[ns++] = 0;
    for (uint32_t q = 0; q < ns; q++) {
        const hb_node *nd = &tree[node_of[q]];
        int32_t ch[2] = { nd->izero, nd->ione };
        for (int k = 0; k < 2; k++)
            if (!is_leaf(&tree[ch[k]])) {
                state_of[ch[k]] = (int32_t)ns;
                depth[ns] = (uint8_t)(depth[q] + 1);   /* indexed by state */
                node_of[ns++] = ch[k];
            }
    }
    if (ns > HB_FSM_MAX_STATES) { free(state_of); free(node_of); free(depth); return HB_OK; }
    if (with_table) out->fsm = (uint16_t *)malloc(sizeof(uint16_t) * 256 * (size_t)ns);
    out->fsm_bstep = (uint16_t *)malloc(sizeof(uint16_t) * 2 * (size_t)ns);
    if ((with_table && !out->fsm) || !out->fsm_bstep) { free(state_of); free(node_of); free(depth); return HB_ERR_NOMEM; }
    for (uint32_t s = 0; s < ns; s++) {
        out->fsm_depth[s] = depth[s];
        out->fsm_node[s] = node_of[s];
        for (uint32_t bit = 0; bit < 2; bit++) {
            int32_t c = bit ? tree[node_of[s]].ione : tree[node_of[s]].izero;
            out->fsm_bstep[2 * s + bit] = is_leaf(&tree[c]) ? (uint16_t)0x100u : (uint16_t)state_of[c];
        }
        for (uint32_t b = 0; with_table && b < 256; b++) {
            int32_t node = node_of[s];
            uint32_t ends = 0;
            for (int i = 0; i < 8; i++) {
                node = ((b >> i) & 1u) ? tree[node].ione : tree[node].izero;
                if (is_leaf(&tree[node])) { ends++; node = 0; }
            }
            out->fsm[s * 256 + b] = (uint16_t)(((uint32_t)state_of[node] << 8) | ends);
        }
    }
    memset(out->fsm_pstep, 0, sizeof(out->fsm_pstep));
    for (uint32_t r = 1; r < 8; r++)
        for (uint32_t x = 0; x < (1u << r); x++) {
            int32_t node = 0;
            uint32_t ends = 0;
            for (uint32_t i = 0; i < r; i++) {
                node = ((x >> i) & 1u) ? tree[node].ione : tree[node].izero;
                if (is_leaf(&tree[node])) { ends++; node = 0; }
            }
            out->fsm_pstep[(1u << r) + x] = (uint16_t)(((uint32_t)state_of[node] << 8) | ends);
        }
    out->fsm_states = ns;
    out->node_state = state_of;
    free(node_of); free(depth);
    return HB_OK;
}

/* ---- multi-symbol tables -----------------------------------------------------
 * The reference has a CPU precedent for several symbols per table entry
 * (decodeBigtableMultiSym, framework/mainrun.c:209-247,300-352: up to 6 symbols
 * per 2^height-entry cell); here the index is only wf <= 12 bits wide so that a
 * table fits shared memory, and the sync kernel's variant carries the start
 * offsets instead of the symbols. */
static int build_fast_tables(const hb_node *tree, hb_lut *out) {
    uint32_t wf = HB_WF_MAX;   /* always full width: short codes then yield several symbols per probe */
    uint32_t n = 1u << wf;
    out->wf = wf;
    out->stab = (uint32_t *)malloc(sizeof(uint32_t) * n);
    out->etab = (uint32_t *)malloc(sizeof(uint32_t) * n);
    const uint32_t wf64 = out->wf64;
    out->e64 = (uint32_t *)malloc(sizeof(uint32_t) * 2 * ((size_t)1 << wf64));
    if (!out->stab || !out->etab || !out->e64) return HB_ERR_NOMEM;
    for (uint32_t x = 0; x < n; x++) {
        uint32_t sm = 0, nsym = 0, used = 0;       /* unlimited symbols (S-table) */
        uint32_t e_syms = 0, e_nsym = 0, e_used = 0; /* at most HB_E_MAXSYM (E-table) */
        uint32_t x_syms = 0, x_nsym = 0, x_used = 0; /* at most HB_E64_MAXSYM (E64-table) */
        uint32_t pos = 0;
        for (;;) {
            int32_t node = 0;
            uint32_t p = pos;
            while (!is_leaf(&tree[node]) && p < wf) {
                node = ((x >> p) & 1u) ? tree[node].ione : tree[node].izero;
                p++;
            }
            if (!is_leaf(&tree[node])) break;      /* next codeword does not fit */
            sm |= 1u << pos;
            nsym++;
            used = p;
            if (e_nsym < HB_E_MAXSYM) {
                e_syms |= (uint32_t)tree[node].sym << (8 * e_nsym);
                e_nsym++;
                e_used = p;
            }
            if (x_nsym < HB_E64_MAXSYM && p <= wf64) {
                x_syms |= (uint32_t)tree[node].sym << (8 * x_nsym);
                x_nsym++;
                x_used = p;
            }
            pos = p;
            if (pos >= wf) break;
        }
        if (nsym == 0) {
            out->stab[x] = HB_FAST_MARK << 16;
            out->etab[x] = HB_FAST_MARK << 16;
        } else {
            out->stab[x] = sm | (used << 16) | (nsym << 24);
            out->etab[x] = e_syms | (e_used << 16) | (e_nsym << 24);
        }
        if (x < (1u << wf64)) {   /* only the low wf64 index bits count for this table */
            if (x_nsym == 0) {
                out->e64[2 * x] = 0;
                out->e64[2 * x + 1] = 0x3210u | (HB_E64_MARK << 26);
            } else {
                out->e64[2 * x] = x_syms;
                out->e64[2 * x + 1] = (0x3210u + 0x1111u * x_nsym) | ((8u * x_nsym) << 16) | (x_used << 26);
            }
        }
    }
    return HB_OK;
}

size_t hb_lut_sizeof(void) { return sizeof(hb_lut); }   /* the Python mirror of the struct checks itself against this */

void hb_lut_free(hb_lut *lut) {
    if (!lut) return;
    free(lut->entries);
    free(lut->stab);
    free(lut->etab);
    free(lut->e64);
    lut->e64 = NULL;
    free(lut->fsm);
    free(lut->fsm_bstep);
    free(lut->node_state);
    lut->node_state = NULL;
    lut->fsm = lut->fsm_bstep = NULL;
    lut->fsm_states = 0;
    lut->entries = NULL;
    lut->stab = lut->etab = NULL;
    lut->n_entries = 0;
}
