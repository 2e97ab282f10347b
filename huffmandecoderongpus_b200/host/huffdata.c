/*
 * huffdata.c -- host data layer of the C harness: .huff / plaintext loading,
 * byte comparison and tree metrics with the contracts of the reference's
 * framework/huffdata.c (loadHuffFile :27-68, loadTextFile :152-165,
 * compareUnCompressedData :183-203, loadTestData :205-215, tableHeight /
 * tableMinDepth / treeSize :224-278), built on the library's container reader
 * (hb_huff_load).  Unlike the reference, a missing file is reported (NULL),
 * not dereferenced.
 */
#include "huffdata.h"
#include "huffb200.h"
#include "sha256.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct CompressedDataL *loadHuffFileL(const char *filename) {
    hb_huff_file f;
    int rc = hb_huff_load(filename, &f);
    if (rc != HB_OK) {
        fprintf(stderr, "loadHuffFile(%s): %s\n", filename, hb_strerror(rc));
        return NULL;
    }
    struct CompressedDataL *cd = (struct CompressedDataL *)malloc(sizeof(*cd));
    if (!cd) { hb_huff_free(&f); return NULL; }
    cd->bits = f.bits;
    cd->nodes = f.nodes;
    cd->uncompressedsize = f.usize;
    cd->tree = (struct HuffNode *)f.tree;   /* same layout, checked in b200approach.c */
    cd->data = f.data;
    return cd;
}

struct CompressedData *loadHuffFile(const char *filename) {
    struct CompressedDataL *l = loadHuffFileL(filename);
    if (!l) return NULL;
    if (l->bits > 0x7fffffffull || l->uncompressedsize > 0x7fffffffull) {
        fprintf(stderr, "loadHuffFile(%s): stream needs the 64-bit loader\n", filename);
        freeCompressedDataL(l);
        return NULL;
    }
    struct CompressedData *cd = (struct CompressedData *)malloc(sizeof(*cd));
    if (!cd) { freeCompressedDataL(l); return NULL; }
    cd->bits = (int)l->bits;
    cd->nodes = l->nodes;
    cd->uncompressedsize = (int)l->uncompressedsize;
    cd->tree = l->tree;
    cd->data = l->data;
    free(l);
    return cd;
}

void freeCompressedData(struct CompressedData *cd) {
    if (!cd) return;
    free(cd->tree);
    free(cd->data);
    free(cd);
}

void freeCompressedDataL(struct CompressedDataL *cd) {
    if (!cd) return;
    free(cd->tree);
    free(cd->data);
    free(cd);
}

struct UnCompressedData *newUnCompressedData(int size) {
    struct UnCompressedData *u = (struct UnCompressedData *)malloc(sizeof(*u));
    if (!u) return NULL;
    u->uncompressedsize = size;
    u->data = (unsigned char *)malloc((size_t)size + 3);   /* +3 as in the reference */
    if (!u->data) { free(u); return NULL; }
    return u;
}

struct UnCompressedData *loadTextFile(const char *filename) {
    FILE *f = fopen(filename, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    struct UnCompressedData *u = (size >= 0 && size < 0x7fffffffL) ? newUnCompressedData((int)size) : NULL;
    if (u && size && fread(u->data, 1, (size_t)size, f) != (size_t)size) {
        freeUnCompressedData(u);
        u = NULL;
    }
    fclose(f);
    return u;
}

void freeUnCompressedData(struct UnCompressedData *u) {
    if (!u) return;
    free(u->data);
    free(u);
}

void clearUnCompressedData(struct UnCompressedData *u) {
    memset(u->data, 0, (size_t)u->uncompressedsize);
}

int compareUnCompressedData(struct UnCompressedData *a, struct UnCompressedData *b) {
    if (a->uncompressedsize != b->uncompressedsize) {
        printf("different size! : %d %d\n", a->uncompressedsize, b->uncompressedsize);
        return -1;
    }
    int diffs = 0;
    for (int i = 0; i < a->uncompressedsize; i++) {
        if (a->data[i] != b->data[i]) {
            if (diffs < 10)
                printf("different at: %d  val1: %d  val2: %d\n", i, a->data[i], b->data[i]);
            diffs++;
        }
    }
    if (diffs) printf("differences %d / %d\n", diffs, a->uncompressedsize);
    return diffs ? -1 : 0;
}

struct TestData *loadTestData(const char *filename, const char *name) {
    char huffname[1024];
    snprintf(huffname, sizeof(huffname), "%s.huff", filename);   /* naming rule of huffdata.c:210-213 */
    struct TestData *td = (struct TestData *)calloc(1, sizeof(*td));
    if (!td) return NULL;
    td->cd = loadHuffFile(huffname);
    td->ucd = loadTextFile(filename);   /* may be NULL: plaintext not shipped */
    td->name = strdup(name);
    if (!td->cd) { freeTestData(td); return NULL; }
    return td;
}

void freeTestData(struct TestData *td) {
    if (!td) return;
    freeCompressedData(td->cd);
    freeUnCompressedData(td->ucd);
    free(td->name);
    free(td);
}

void infoCompressedData(struct CompressedData *cd) {
    printf("nodes %d, bits %d, uncompressedsize %d\n", cd->nodes, cd->bits, cd->uncompressedsize);
}

void infoTestData(struct TestData *td) {
    printf("%s ", td->name);
    infoCompressedData(td->cd);
}

/* tree metrics, iterative (explicit stack) */
static int tree_metric(struct HuffNode *tree, int r, int want_min, int want_size) {
    int stack_n[64 * 2 + 8], stack_d[64 * 2 + 8], sp = 0;
    int best = want_min ? 0x7fffffff : 0, size = 0;
    stack_n[sp] = r; stack_d[sp++] = 0;
    while (sp > 0) {
        int v = stack_n[--sp], d = stack_d[sp];
        size++;
        if (tree[v].izero == -1) {
            if (want_min ? d < best : d > best) best = d;
            continue;
        }
        if (sp + 2 > (int)(sizeof(stack_n) / sizeof(stack_n[0]))) return -1;   /* deeper than 64 */
        stack_n[sp] = tree[v].izero; stack_d[sp++] = d + 1;
        stack_n[sp] = tree[v].ione;  stack_d[sp++] = d + 1;
    }
    return want_size ? size : best;
}

int tableHeight(struct HuffNode *tree, int r) { return tree_metric(tree, r, 0, 0); }
int tableMinDepth(struct HuffNode *tree, int r) { return tree_metric(tree, r, 1, 0); }
int treeSize(struct HuffNode *tree, int r) { return tree_metric(tree, r, 0, 1); }

int digestUnCompressedData(struct UnCompressedData *u, char hex[65]) {
    sha256_hex(u->data, (size_t)u->uncompressedsize, hex);
    return 0;
}
