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
static hb_result g_last;
static int g_atexit;

static void fatal(const char *what, int rc) {
    printf("b200Approach failed in %s: %s (%s)\n", what, hb_strerror(rc),
           g_ctx ? hb_last_error(g_ctx) : "");
    exit(-1);
}

void b200ApproachShutdown(void) {
    if (g_ctx) {
        hb_ctx_destroy(g_ctx);
        g_ctx = NULL;
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
    int rc = hb_decode_host(g_ctx, (const hb_node_abi *)cd->tree, cd->nodes, cd->data, cd->bits,
                            uncompressed->data, uncompressed->uncompressedsize, &g_last);
    if (rc != HB_OK) fatal("hb_decode_host", rc);
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

double b200ApproachLastDeviceMs(void) { return g_last.ms_total; }
unsigned long long b200ApproachLastSymbols(void) { return g_last.n_symbols; }
