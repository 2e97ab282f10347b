/*
 * huff_oracle.c -- TEST INFRASTRUCTURE ONLY (see huff_oracle.h).
 *
 * Plain-C restatement of the reference's CPU decode path with 64-bit stream
 * positions.  Not linked into, or called from, the product library.
 */
#include "huff_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORA_PAD 16 /* zero bytes kept after the payload (reference keeps 3) */

/* ---- .huff container ----------------------------------------------------
 * reference framework/huffdata.c:21-25 (readBint: big-endian int32) and
 * :27-68 (loadHuffFile): "HUFF", BE i32 nodes, bits, uncompressedsize, then
 * nodes x {u8 sym, BE i32 izero, BE i32 ione}, then ceil(bits/8) data bytes.
 * "HUF8" (this repo): same, but bits and uncompressedsize are BE u64. */

static int rd_be32(FILE *f, int32_t *v) {
    unsigned char b[4];
    if (fread(b, 1, 4, f) != 4) return -1;
    *v = (int32_t)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) |
                   ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
    return 0;
}

static int rd_be64(FILE *f, uint64_t *v) {
    unsigned char b[8];
    if (fread(b, 1, 8, f) != 8) return -1;
    uint64_t x = 0;
    for (int i = 0; i < 8; i++) x = (x << 8) | b[i];
    *v = x;
    return 0;
}

ora_stream *ora_load_huff(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    ora_stream *s = (ora_stream *)calloc(1, sizeof(*s));
    char magic[4];
    int32_t v32;
    if (!s || fread(magic, 1, 4, f) != 4) goto bad;
    if (memcmp(magic, "HUFF", 4) == 0) s->wide = 0;
    else if (memcmp(magic, "HUF8", 4) == 0) s->wide = 1;
    else goto bad;
    if (rd_be32(f, &s->nodes) || s->nodes <= 0) goto bad;
    if (s->wide) {
        if (rd_be64(f, &s->bits) || rd_be64(f, &s->usize)) goto bad;
    } else {
        if (rd_be32(f, &v32) || v32 < 0) goto bad;
        s->bits = (uint64_t)v32;
        if (rd_be32(f, &v32) || v32 < 0) goto bad;
        s->usize = (uint64_t)v32;
    }
    s->tree = (ora_node *)calloc((size_t)s->nodes, sizeof(ora_node));
    if (!s->tree) goto bad;
    for (int32_t i = 0; i < s->nodes; i++) {
        int c = fgetc(f);
        if (c == EOF) goto bad;
        s->tree[i].sym = (uint8_t)c;
        if (rd_be32(f, &s->tree[i].izero) || rd_be32(f, &s->tree[i].ione))
            goto bad;
    }
    {
        uint64_t nbytes = (s->bits + 7) / 8;
        s->data = (uint8_t *)calloc((size_t)nbytes + ORA_PAD, 1);
        if (!s->data) goto bad;
        if (nbytes && fread(s->data, 1, (size_t)nbytes, f) != nbytes) goto bad;
    }
    fclose(f);
    return s;
bad:
    if (f) fclose(f);
    ora_free_stream(s);
    return NULL;
}

void ora_free_stream(ora_stream *s) {
    if (!s) return;
    free(s->tree);
    free(s->data);
    free(s);
}

/* ---- tree metrics: reference framework/huffdata.c:224-238,272-278 -------- */

int ora_tree_height(const ora_node *tree, int r) {
    if (tree[r].izero == -1) return 0;
    int a = ora_tree_height(tree, tree[r].izero);
    int b = ora_tree_height(tree, tree[r].ione);
    return 1 + (a > b ? a : b);
}

int ora_tree_mindepth(const ora_node *tree, int r) {
    if (tree[r].izero == -1) return 0;
    int a = ora_tree_mindepth(tree, tree[r].izero);
    int b = ora_tree_mindepth(tree, tree[r].ione);
    return 1 + (a < b ? a : b);
}

int ora_tree_size(const ora_node *tree, int r) {
    if (tree[r].izero == -1) return 1;
    return 1 + ora_tree_size(tree, tree[r].izero) +
           ora_tree_size(tree, tree[r].ione);
}

static inline int ora_bit(const uint8_t *data, uint64_t p) {
    return (data[p >> 3] >> (p & 7)) & 1;
}

static inline int ora_is_leaf(const ora_node *n) {
    return n->izero == -1 && n->ione == -1;
}

/* ---- serial oracle: reference framework/mainrun.c:38-55 ------------------ */

uint64_t ora_simple_decode(const ora_node *tree, const uint8_t *data,
                           uint64_t bits, uint8_t *out, uint64_t outcap) {
    uint64_t n = 0;
    int32_t node = 0;
    for (uint64_t p = 0; p < bits; p++) {
        node = ora_bit(data, p) ? tree[node].ione : tree[node].izero;
        if (ora_is_leaf(&tree[node])) {
            if (n < outcap) out[n] = tree[node].sym;
            n++;
            node = 0;
        }
    }
    return n;
}

/* ---- prefix truncation: reference framework/mainrun.c:361-385 ------------ */

void ora_prefix_sizes(const ora_node *tree, const uint8_t *data,
                      uint64_t targetbits, uint64_t *bits_out,
                      uint64_t *usize_out) {
    uint64_t n = 0, last = 0;
    int32_t node = 0;
    for (uint64_t p = 0; p < targetbits; p++) {
        node = ora_bit(data, p) ? tree[node].ione : tree[node].izero;
        if (ora_is_leaf(&tree[node])) {
            n++;
            node = 0;
            last = p;
        }
    }
    /* the reference sets bits = lastokaypos + 1 even when no codeword
     * completed (lastokaypos stays 0), mainrun.c:383 */
    *bits_out = last + 1;
    *usize_out = n;
}

/* ---- jump-table FSM: reference framework/jumptableapproach.c ------------- *
 * A state is "the tree node reached by the bits of the unfinished codeword".
 * The reference memoises states by (prebits, prebitsnum) (:45-52); a prefix
 * identifies exactly one tree node, so memoising by node index is the same
 * state set.  Row layout per state: 2^jumpbits cells {next, nsym, syms[]}.
 * Chunk bits are consumed LSB-first (:79 "reverse the bits").
 * Deviation (documented): the reference refuses jumpbits/mindepth > 7
 * (:146-147, 7-symbol cells); cells here hold up to 16 symbols so that trees
 * with a 1-bit code can still be timed.  On every input the reference accepts
 * the output is identical (tests/test_oracle.py).                            */

#define ORA_JT_MAXSYM 16

typedef struct ora_jcell {
    int32_t next;
    uint8_t nsym;
    uint8_t syms[ORA_JT_MAXSYM];
} ora_jcell;

typedef struct ora_jt {
    int jumpbits, width, nstates, cap;
    const ora_node *tree;
    int32_t *state_of_node; /* node -> state id or -1 */
    int32_t *node_of_state;
    int32_t *depth_of_state; /* bits already consumed inside the codeword */
    ora_jcell *cells;
} ora_jt;

static int32_t jt_state(ora_jt *jt, int32_t node, int32_t depth) {
    if (jt->state_of_node[node] >= 0) return jt->state_of_node[node];
    int32_t id = jt->nstates++;
    jt->state_of_node[node] = id;
    jt->node_of_state[id] = node;
    jt->depth_of_state[id] = depth;
    for (int c = 0; c < jt->width; c++) {
        int32_t cur = node, d = depth;
        ora_jcell *cell = &jt->cells[(size_t)id * jt->width + c];
        int ns = 0;
        for (int j = 0; j < jt->jumpbits; j++) {
            cur = ((c >> j) & 1) ? jt->tree[cur].ione : jt->tree[cur].izero;
            d++;
            if (jt->tree[cur].ione == -1) {
                cell->syms[ns++] = jt->tree[cur].sym;
                cur = 0;
                d = 0;
            }
        }
        cell->nsym = (uint8_t)ns;
        int32_t nx = jt_state(jt, cur, d);
        /* cells may have moved?  no: storage is allocated up front */
        jt->cells[(size_t)id * jt->width + c].next = nx;
    }
    return id;
}

uint64_t ora_jumptable_decode(const ora_node *tree, int nodes,
                              const uint8_t *data, uint64_t bits, int jumpbits,
                              uint8_t *out, uint64_t outcap) {
    if (jumpbits < 1 || jumpbits > 15) return (uint64_t)-1;
    int mind = ora_tree_mindepth(tree, 0);
    if (mind < 1 || jumpbits / mind > ORA_JT_MAXSYM) return (uint64_t)-1;

    ora_jt jt;
    jt.jumpbits = jumpbits;
    jt.width = 1 << jumpbits;
    jt.nstates = 0;
    jt.cap = nodes; /* at most one state per tree node */
    jt.tree = tree;
    jt.state_of_node = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    jt.node_of_state = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    jt.depth_of_state = (int32_t *)malloc(sizeof(int32_t) * (size_t)nodes);
    jt.cells = (ora_jcell *)malloc(sizeof(ora_jcell) * (size_t)nodes * jt.width);
    if (!jt.state_of_node || !jt.node_of_state || !jt.depth_of_state || !jt.cells) {
        free(jt.state_of_node); free(jt.node_of_state);
        free(jt.depth_of_state); free(jt.cells);
        return (uint64_t)-1;
    }
    for (int i = 0; i < nodes; i++) jt.state_of_node[i] = -1;
    jt_state(&jt, 0, 0);

    uint64_t n = 0, pos = 0;
    int32_t st = 0;
    const unsigned mask = (unsigned)jt.width - 1u;

    if (jumpbits == 8) {
        /* :174-188 one whole byte per step */
        uint64_t nbytes = bits >> 3;
        for (uint64_t b = 0; b < nbytes; b++) {
            const ora_jcell *c = &jt.cells[(size_t)st * jt.width + data[b]];
            for (int j = 0; j < c->nsym; j++) {
                if (n < outcap) out[n] = c->syms[j];
                n++;
            }
            st = c->next;
        }
        pos = nbytes << 3;
    } else {
        /* :207-240 unaligned jumpbits-wide window per step; the reference
         * stops while pos < bits - jumpbits */
        while (pos + (uint64_t)jumpbits < bits) {
            uint64_t by = pos >> 3;
            unsigned w = (unsigned)data[by] | ((unsigned)data[by + 1] << 8) |
                         ((unsigned)data[by + 2] << 16);
            unsigned idx = (w >> (pos & 7)) & mask;
            const ora_jcell *c = &jt.cells[(size_t)st * jt.width + idx];
            for (int j = 0; j < c->nsym; j++) {
                if (n < outcap) out[n] = c->syms[j];
                n++;
            }
            st = c->next;
            pos += (uint64_t)jumpbits;
        }
    }
    /* bit-serial tail (:190-205,242-257).  The reference rewinds by the
     * state's prefix length and restarts at the root; continuing from the
     * state's node is the same walk. */
    {
        int32_t node = jt.node_of_state[st];
        for (; pos < bits; pos++) {
            node = ora_bit(data, pos) ? tree[node].ione : tree[node].izero;
            if (ora_is_leaf(&tree[node])) {
                if (n < outcap) out[n] = tree[node].sym;
                n++;
                node = 0;
            }
        }
    }
    free(jt.state_of_node);
    free(jt.node_of_state);
    free(jt.depth_of_state);
    free(jt.cells);
    return n;
}

/* ---- per-offset phase statement: reference framework/pes.c:30-46 --------- */

void ora_decode_all_bits(const ora_node *tree, const uint8_t *data,
                         uint64_t bits, uint8_t *sym_out, int32_t *len_out) {
    for (uint64_t b = 0; b < bits; b++) {
        uint64_t p = b;
        int32_t node = 0;
        while (tree[node].izero != -1 && p < bits) {
            node = ora_bit(data, p) ? tree[node].ione : tree[node].izero;
            p++;
        }
        sym_out[b] = tree[node].sym;
        len_out[b] = (int32_t)(p - b);
    }
}

/* ---- the reference parallel algorithm run serially: pes.c:48-104,106-209 - *
 * level[k][b] = total length of 2^k consecutive codewords starting at b, or
 * -1 if that runs past the end; index[b] = output position of the codeword
 * starting at b (true starts only), filled top-down from index[0] = 0.       */

uint64_t ora_pes_decode(const ora_node *tree, const uint8_t *data,
                        uint64_t bits, uint8_t *out, uint64_t outcap) {
    if (bits == 0) return 0;
    uint8_t *sym = (uint8_t *)malloc((size_t)bits);
    int64_t *index = (int64_t *)malloc(sizeof(int64_t) * (size_t)bits);
    int32_t **level = (int32_t **)calloc(64, sizeof(int32_t *));
    level[0] = (int32_t *)malloc(sizeof(int32_t) * (size_t)bits);
    ora_decode_all_bits(tree, data, bits, sym, level[0]);
    for (uint64_t b = 0; b < bits; b++) index[b] = -1;

    int nlev = 0;
    for (;;) { /* pes.c:48-71, loop at :158-167 */
        int32_t *cur = level[nlev];
        int32_t *nxt = (int32_t *)malloc(sizeof(int32_t) * (size_t)bits);
        level[nlev + 1] = nxt;
        for (uint64_t b = 0; b < bits; b++) {
            int32_t s = cur[b];
            if (s == -1 || b + (uint64_t)s > bits) { nxt[b] = -1; continue; }
            /* the reference reads one-past-the-end here when b + s == bits
             * (benign race); that element means "no further
             * codeword", i.e. -1 */
            int32_t w = (b + (uint64_t)s == bits) ? -1 : cur[b + (uint64_t)s];
            if (w == -1 || b + (uint64_t)s + (uint64_t)w > bits) nxt[b] = -1;
            else nxt[b] = s + w;
        }
        int32_t probe = cur[0];
        nlev++;
        if (probe == -1) break;
    }
    index[0] = 0; /* pes.c:182 */
    for (int k = nlev; k > 0; k--) { /* pes.c:73-85, loop at :188-192 */
        const int32_t *lv = level[k - 1];
        int64_t add = (int64_t)1 << (k - 1);
        for (uint64_t b = 0; b < bits; b++) {
            int32_t off = lv[b];
            if (off != -1 && index[b] != -1 && b + (uint64_t)off < bits)
                index[b + (uint64_t)off] = index[b] + add;
        }
    }
    int64_t maxidx = 0;
    for (uint64_t b = 0; b < bits; b++) { /* pes.c:87-96 */
        if (index[b] != -1) {
            if ((uint64_t)index[b] < outcap) out[index[b]] = sym[b];
            if (index[b] > maxidx) maxidx = index[b];
        }
    }
    for (int k = 0; k <= nlev; k++) free(level[k]);
    free(level);
    free(index);
    free(sym);
    return (uint64_t)maxidx + 1; /* pes.c:98-104,203-204 */
}
