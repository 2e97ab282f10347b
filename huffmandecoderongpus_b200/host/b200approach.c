/*
 * b200approach.c -- approach-table entry point over the C ABI (include/huffb200.h).
 *
 * Replaces the host orchestration of the reference's fastgpuApproach
 * (framework/fastgpu.cu:140-332).  Like the reference's OpenCL approach, which
 * keeps platform/context/queue in statics across calls
 * (framework/openclapproach.c:231-234,286-328), the device context is created
 * on first use and reused by the 26 back-to-back calls of evaluate()
 * (framework/decodeUtil.c:41-58).
 */
#include "b200approach.h"
#include "huffdata.h"
#include "huffb200.h"

#include <stdio.h>
#include <stdlib.h>

static hb_ctx *g_ctx;
static hb_multi *g_multi;
static int g_multi_n;
static hb_result g_last;
static int g_atexit;

/* evaluate() calls an approach 26 times with the same two buffers (framework/decodeUtil.c:
 * 41-58): page-lock them on first sight, so that the copies are DMA transfers instead of
 * staged pageable copies (and overlap across devices in the multi-GPU approach).  A small
 * cache; a range that overlaps a new one is released first.
 * OFF unless the host program asks for it (b200ApproachPinBuffers(1), or B200_PIN=1 in the
 * environment): a range that stays page-locked after its owner freed it makes any LATER CUDA
 * copy of the application from memory overlapping it fail, so the program that turns this
 * on must call b200ApproachReleaseBuffers() before it frees the buffers (host/decodeUtil.c
 * does, at the end of evaluate()). */
#define NPINS 4
static struct { const unsigned char *p; size_t n; unsigned age; } g_pins[NPINS];
static unsigned g_pin_clock;
static int g_pin_enabled = -1;

void b200ApproachPinBuffers(int on) { g_pin_enabled = on ? 1 : 0; }

static void pin_range(const void *ptr, size_t n) {
    if (g_pin_enabled < 0) { const char *s = getenv("B200_PIN"); g_pin_enabled = s && s[0] == '1'; }
    if (!g_pin_enabled || !ptr || n < 4096) return;
    const unsigned char *p = (const unsigned char *)ptr;
    int slot = -1;
    for (int i = 0; i < NPINS; i++) {
        if (!g_pins[i].p) { if (slot < 0) slot = i; continue; }
        if (g_pins[i].p == p && g_pins[i].n >= n) { g_pins[i].age = ++g_pin_clock; return; }
        if (p < g_pins[i].p + g_pins[i].n && g_pins[i].p < p + n) {   /* stale overlap: buffer was freed and reused */
            hb_host_unpin(g_pins[i].p);
            g_pins[i].p = NULL;
            if (slot < 0) slot = i;
        }
    }
    if (slot < 0) {
        slot = 0;
        for (int i = 1; i < NPINS; i++) if (g_pins[i].age < g_pins[slot].age) slot = i;
        hb_host_unpin(g_pins[slot].p);
        g_pins[slot].p = NULL;
    }
    if (hb_host_pin(p, n) == HB_OK) { g_pins[slot].p = p; g_pins[slot].n = n; g_pins[slot].age = ++g_pin_clock; }
}

static void unpin_all(void) {
    for (int i = 0; i < NPINS; i++)
        if (g_pins[i].p) { hb_host_unpin(g_pins[i].p); g_pins[i].p = NULL; }
}

void b200ApproachReleaseBuffers(void) { unpin_all(); }

static void fatal(const char *what, int rc) {
    printf("b200Approach failed in %s: %s (%s)\n", what, hb_strerror(rc),
           g_ctx ? hb_last_error(g_ctx) : "");
    exit(-1);
}

void b200ApproachShutdown(void) {
    unpin_all();
    if (g_ctx) {
        hb_ctx_destroy(g_ctx);
        g_ctx = NULL;
    }
    if (g_multi) {
        hb_multi_destroy(g_multi);
        g_multi = NULL;
    }
}

static void ensure_ctx(void) {
    if (g_ctx) return;
    int dev = 0;
    const char *s = getenv("B200_DEVICE");
    if (s) dev = atoi(s);
    int rc = hb_ctx_create(dev, NULL, &g_ctx);
    if (rc != HB_OK) fatal("hb_ctx_create (no CUDA device? there is no CPU fallback)", rc);
    s = getenv("B200_WPT");
    if (s) {
        rc = hb_ctx_configure(g_ctx, atoi(s), 0);
        if (rc != HB_OK) fatal("hb_ctx_configure", rc);
    }
    if (!g_atexit) {
        atexit(b200ApproachShutdown);
        g_atexit = 1;
    }
}

void b200ApproachL(struct CompressedDataL *cd, struct UnCompressedDataL *uncompressed,
                   void *paramdata) {
    (void)paramdata;
    ensure_ctx();
    /* struct HuffNode and hb_node_abi have the same layout (checked below) */
    _Static_assert(sizeof(struct HuffNode) == sizeof(hb_node_abi), "node layout");
    pin_range(cd->data, (size_t)((cd->bits + 7) / 8));
    pin_range(uncompressed->data, (size_t)uncompressed->uncompressedsize);
    int rc = hb_decode_host(g_ctx, (const hb_node_abi *)cd->tree, cd->nodes, cd->data, cd->bits,
                            uncompressed->data, uncompressed->uncompressedsize, &g_last);
    if (rc != HB_OK) fatal("hb_decode_host", rc);
}

/* One stream over several GPUs of this process.  paramdata: int * = device count (NULL:
 * B200_DEVICES, else every visible device) -- the reference passes approach parameters the
 * same way (jumpbits, framework/mainrun.c:442,500-501). */
void b200ApproachMultiL(struct CompressedDataL *cd, struct UnCompressedDataL *uncompressed,
                        void *paramdata) {
    int want = paramdata ? *(int *)paramdata : 0;
    if (!want && getenv("B200_DEVICES")) want = atoi(getenv("B200_DEVICES"));
    if (g_multi && g_multi_n != want) { hb_multi_destroy(g_multi); g_multi = NULL; }
    if (!g_multi) {
        int rc = hb_multi_create(NULL, want, &g_multi);
        if (rc != HB_OK) fatal("hb_multi_create (no CUDA device? there is no CPU fallback)", rc);
        g_multi_n = want;
        if (!g_atexit) { atexit(b200ApproachShutdown); g_atexit = 1; }
    }
    pin_range(cd->data, (size_t)((cd->bits + 7) / 8));
    pin_range(uncompressed->data, (size_t)uncompressed->uncompressedsize);
    hb_multi_result r;
    int rc = hb_multi_decode_host(g_multi, (const hb_node_abi *)cd->tree, cd->nodes, cd->data, cd->bits,
                                  uncompressed->data, uncompressed->uncompressedsize, &r);
    if (rc != HB_OK) {
        printf("b200ApproachMulti failed: %s (%s)\n", hb_strerror(rc), hb_multi_last_error(g_multi));
        exit(-1);
    }
    g_last.n_symbols = r.n_symbols;
    g_last.ms_total = r.ms_device_max;
    g_last.launches = r.launches;
}

void b200ApproachMulti(struct CompressedData *cd, struct UnCompressedData *uncompressed,
                       void *paramdata) {
    struct CompressedDataL cdl;
    struct UnCompressedDataL ul;
    cdl.bits = (uint64_t)(cd->bits < 0 ? 0 : cd->bits);
    cdl.nodes = cd->nodes;
    cdl.uncompressedsize = (uint64_t)(cd->uncompressedsize < 0 ? 0 : cd->uncompressedsize);
    cdl.tree = cd->tree;
    cdl.data = cd->data;
    ul.uncompressedsize = (uint64_t)(uncompressed->uncompressedsize < 0 ? 0 : uncompressed->uncompressedsize);
    ul.data = uncompressed->data;
    b200ApproachMultiL(&cdl, &ul, paramdata);
}

void b200Approach(struct CompressedData *cd, struct UnCompressedData *uncompressed,
                  void *paramdata) {
    struct CompressedDataL cdl;
    struct UnCompressedDataL ul;
    cdl.bits = (uint64_t)(cd->bits < 0 ? 0 : cd->bits);
    cdl.nodes = cd->nodes;
    cdl.uncompressedsize = (uint64_t)(cd->uncompressedsize < 0 ? 0 : cd->uncompressedsize);
    cdl.tree = cd->tree;
    cdl.data = cd->data;
    ul.uncompressedsize = (uint64_t)(uncompressed->uncompressedsize < 0 ? 0 : uncompressed->uncompressedsize);
    ul.data = uncompressed->data;
    b200ApproachL(&cdl, &ul, paramdata);
}

/* the reference's onethreadApproach (framework/onethread.cu:33-52): one device thread */
void onethreadApproach(struct CompressedData *cd, struct UnCompressedData *uncompressed, void *paramdata) {
    (void)paramdata;
    ensure_ctx();
    int rc = hb_decode_onethread(g_ctx, (const hb_node_abi *)cd->tree, cd->nodes, cd->data,
                                 (uint64_t)(cd->bits < 0 ? 0 : cd->bits), uncompressed->data,
                                 (uint64_t)(uncompressed->uncompressedsize < 0 ? 0 : uncompressed->uncompressedsize),
                                 &g_last);
    if (rc != HB_OK) fatal("hb_decode_onethread", rc);
}

double b200ApproachLastDeviceMs(void) { return g_last.ms_total; }
unsigned long long b200ApproachLastSymbols(void) { return g_last.n_symbols; }
