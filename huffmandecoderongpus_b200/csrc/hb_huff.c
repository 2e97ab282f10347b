/*
 * hb_huff.c -- .huff container reader / writer (host, plain C).
 *
 * Format, as read by the reference's loadHuffFile (framework/huffdata.c:27-68):
 *   "HUFF" | BE i32 nodes | BE i32 bits | BE i32 uncompressedsize |
 *   nodes x { u8 sym, BE i32 izero, BE i32 ione } | ceil(bits/8) data bytes
 * The reference's header fields are 32-bit signed (framework/huffdata.h:26-32),
 * which caps a stream at 2^31-1 bits; "HUF8" is the same layout with BE u64
 * bits and uncompressedsize for the GB-scale synthetic configurations.
 */
#include "huffb200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HB_DATA_PAD 16

static int get32(FILE *f, int32_t *v) {
    unsigned char b[4];
    if (fread(b, 1, 4, f) != 4) return -1;
    *v = (int32_t)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]);
    return 0;
}

static int get64(FILE *f, uint64_t *v) {
    int32_t hi, lo;
    if (get32(f, &hi) || get32(f, &lo)) return -1;
    *v = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    return 0;
}

static void put32(unsigned char *b, uint32_t v) {
    b[0] = (unsigned char)(v >> 24); b[1] = (unsigned char)(v >> 16);
    b[2] = (unsigned char)(v >> 8);  b[3] = (unsigned char)v;
}

int hb_huff_load(const char *path, hb_huff_file *out) {
    if (!path || !out) return HB_ERR_ARG;
    memset(out, 0, sizeof(*out));
    FILE *f = fopen(path, "rb");
    if (!f) return HB_ERR_IO;
    int rc = HB_ERR_FORMAT;
    char magic[4];
    int32_t v;
    if (fread(magic, 1, 4, f) != 4) goto done;
    if (!memcmp(magic, "HUFF", 4)) out->wide = 0;
    else if (!memcmp(magic, "HUF8", 4)) out->wide = 1;
    else goto done;
    if (get32(f, &out->nodes) || out->nodes <= 0 || out->nodes > (1 << 20)) goto done;
    if (out->wide) {
        if (get64(f, &out->bits) || get64(f, &out->usize)) goto done;
    } else {
        if (get32(f, &v) || v < 0) goto done;
        out->bits = (uint64_t)v;
        if (get32(f, &v) || v < 0) goto done;
        out->usize = (uint64_t)v;
    }
    out->tree = (hb_node_abi *)calloc((size_t)out->nodes, sizeof(hb_node_abi));
    if (!out->tree) { rc = HB_ERR_NOMEM; goto done; }
    {
        /* 9 bytes per node on disk */
        size_t raw = (size_t)out->nodes * 9;
        unsigned char *buf = (unsigned char *)malloc(raw);
        if (!buf) { rc = HB_ERR_NOMEM; goto done; }
        if (fread(buf, 1, raw, f) != raw) { free(buf); goto done; }
        for (int32_t i = 0; i < out->nodes; i++) {
            const unsigned char *p = buf + (size_t)i * 9;
            out->tree[i].sym = p[0];
            out->tree[i].izero = (int32_t)(((uint32_t)p[1] << 24) | ((uint32_t)p[2] << 16) |
                                           ((uint32_t)p[3] << 8) | p[4]);
            out->tree[i].ione = (int32_t)(((uint32_t)p[5] << 24) | ((uint32_t)p[6] << 16) |
                                          ((uint32_t)p[7] << 8) | p[8]);
        }
        free(buf);
    }
    {
        /* the header is untrusted: the data bytes it promises must exist in the file before
         * anything is allocated for them (and bits + 7 must not wrap) */
        const uint64_t nbytes = out->bits / 8 + (out->bits % 8 != 0);
        long here = ftell(f), end = -1;
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0) end = ftell(f);
        if (here < 0 || end < here || fseek(f, here, SEEK_SET) != 0) { rc = HB_ERR_IO; goto done; }
        if (nbytes > (uint64_t)(end - here)) goto done;
        out->data = (uint8_t *)calloc((size_t)nbytes + HB_DATA_PAD, 1);
        if (!out->data) { rc = HB_ERR_NOMEM; goto done; }
        if (nbytes && fread(out->data, 1, (size_t)nbytes, f) != nbytes) goto done;
    }
    rc = HB_OK;
done:
    fclose(f);
    if (rc != HB_OK) hb_huff_free(out);
    return rc;
}

int hb_huff_save(const char *path, const hb_huff_file *in, int wide) {
    if (!path || !in || !in->tree || in->nodes <= 0) return HB_ERR_ARG;
    if (!wide && (in->bits > 0x7fffffffull || in->usize > 0x7fffffffull)) return HB_ERR_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return HB_ERR_IO;
    unsigned char hdr[24];
    size_t hl;
    memcpy(hdr, wide ? "HUF8" : "HUFF", 4);
    put32(hdr + 4, (uint32_t)in->nodes);
    if (wide) {
        put32(hdr + 8, (uint32_t)(in->bits >> 32));  put32(hdr + 12, (uint32_t)in->bits);
        put32(hdr + 16, (uint32_t)(in->usize >> 32)); put32(hdr + 20, (uint32_t)in->usize);
        hl = 24;
    } else {
        put32(hdr + 8, (uint32_t)in->bits);
        put32(hdr + 12, (uint32_t)in->usize);
        hl = 16;
    }
    int ok = fwrite(hdr, 1, hl, f) == hl;
    for (int32_t i = 0; ok && i < in->nodes; i++) {
        unsigned char nb[9];
        nb[0] = in->tree[i].sym;
        put32(nb + 1, (uint32_t)in->tree[i].izero);
        put32(nb + 5, (uint32_t)in->tree[i].ione);
        ok = fwrite(nb, 1, 9, f) == 9;
    }
    uint64_t nbytes = (in->bits + 7) / 8;
    if (ok && nbytes) ok = fwrite(in->data, 1, (size_t)nbytes, f) == nbytes;
    if (fclose(f) != 0) ok = 0;
    return ok ? HB_OK : HB_ERR_IO;
}

void hb_huff_free(hb_huff_file *f) {
    if (!f) return;
    free(f->tree);
    free(f->data);
    f->tree = NULL;
    f->data = NULL;
}
