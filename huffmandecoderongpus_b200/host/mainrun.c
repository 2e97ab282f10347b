/*
 * mainrun.c -- C harness with the shape of the reference's framework/mainrun.c:
 * an approach table (newDecoder, :480-501), datasets loaded by name (:505-509),
 * one positional test name, "name dataset time" lines (evalandshow, :412-420).
 *
 * Differences the survey asked for: the test names kjv,
 * ecoli, world192 and bible exist; corpora without a shipped plaintext are
 * checked by SHA-256; the data directory is an argument instead of a
 * hard-coded ../../files; device time and GB/s are appended to each line.
 *
 *   HuffFramework <test> [files-dir]
 *   tests: hello paper1 news book2 kjv ecoli world192 bible bigtable all
 *          quickgraph1..3 graph1..4 (the reference's names, framework/mainrun.c:590-616;
 *          quickgraph / graph = the GPU ones, quickgraph2 / graph2) kjvprof opt bts
 *          multi (every corpus through b200ApproachMulti) onethread
 *          synth1g synthfib synth16g (B200_DEVICES=N: one stream over N GPUs)
 *
 * The GPU approach (b200Approach) is always registered.  The CPU baselines
 * (the oracle's restatement of simpleDecode / jumptableApproach) are only
 * linked with -DWITH_ORACLE_BASELINES (make HuffFrameworkBaselines): the
 * product binary has no CPU decode path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200approach.h"
#include "decodeUtil.h"
#include "huffb200.h"
#include "huffdata.h"

#ifdef WITH_ORACLE_BASELINES
#include <dlfcn.h>
#include "huff_oracle.h"
static void simpleDecode(struct CompressedData *cd, struct UnCompressedData *u, void *p) {
    (void)p;
    ora_simple_decode((const ora_node *)cd->tree, cd->data, (uint64_t)cd->bits, u->data,
                      (uint64_t)u->uncompressedsize);
}
static void jumptableApproach(struct CompressedData *cd, struct UnCompressedData *u, void *p) {
    ora_jumptable_decode((const ora_node *)cd->tree, cd->nodes, cd->data, (uint64_t)cd->bits,
                         *(int *)p, u->data, (uint64_t)u->uncompressedsize);
}
#endif

struct dataset {
    const char *name, *file, *sha256;   /* sha256 of the reference's own decode */
    struct TestData *td;
};

static struct dataset g_sets[] = {
    { "hello", "hello", "a591a6d40bf420404a011733cfb7b190d62c65bf0bcda32b57b277d9ad9f146e", NULL },
    { "paper1", "paper1", "8d9c42d9fa58b5bce1a8b5fae3cc27c9eb7cc7a032bc12a633d44e816497e143", NULL },
    { "news", "news", "7f0482f9774681429eb7021050c17966f6acf19450e170de6611e1ed953d42e8", NULL },
    { "book2", "book2", "c8538730cf2ce6a243acf3eb299c43d619b5c695d892f4884df796c13081fdf8", NULL },
    { "kjv", "kjv.txt", "e4e21579f6360b35e66dc97b67cd732a3f759623e41e4e077bec039eeb79fd0a", NULL },
    { "ecoli", "E.coli", "9125dfd87315961ef4286f3856098069e050cc3a2abe65735fe43e69d1996f40", NULL },
    { "world192", "world192.txt", "1aebdc97d29904b25791da9aa32be90b69d7da6dc0ac9b95512ed27ed40d2112", NULL },
    { "bible", "bible.txt", "4e0a7e8dff7d9c82dbded57305c0ca3cdd3c4ca014db27121782fe9710f4723f", NULL },
};
#define NSETS ((int)(sizeof(g_sets) / sizeof(g_sets[0])))

static const char *g_dir = "../../oracle/_ref/files";

static struct dataset *get_set(const char *name) {
    for (int i = 0; i < NSETS; i++) {
        if (strcmp(g_sets[i].name, name)) continue;
        if (!g_sets[i].td) {
            char path[1024];
            snprintf(path, sizeof(path), "%s/%s", g_dir, g_sets[i].file);
            g_sets[i].td = loadTestData(path, g_sets[i].name);
            if (!g_sets[i].td) {
                fprintf(stderr, "error exit: cannot load %s.huff\n", path);
                exit(1);
            }
        }
        return &g_sets[i];
    }
    return NULL;
}

static void evalandshow(struct decoder *d, struct dataset *s, int withcheck) {
    struct evalresult r = evaluate_ex(d, s->td, withcheck, s->sha256);
    const double out_b = (double)s->td->cd->uncompressedsize;
    const double in_b = (double)((s->td->cd->bits + 7) / 8);
    if (d->paramdata != NULL)   /* the reference prints seconds in this form, mainrun.c:413-415 */
        printf("%17s %8s  %2d %.9f", d->name, s->name, *(int *)d->paramdata, r.min_seconds);
    else
        printf("%17s %8s     %.9f ms", d->name, s->name, r.min_seconds * 1000.0);
    printf("   | %.4f GB/s out, %.4f GB/s in (wall)", out_b / r.min_seconds / 1e9, in_b / r.min_seconds / 1e9);
    if (r.min_device_ms >= 0)
        printf(", device %.4f ms = %.3f GB/s out", r.min_device_ms, out_b / (r.min_device_ms * 1e-3) / 1e9);
    printf(" [%s]\n", r.checked == 1 ? "bytes ok" : r.checked == 2 ? "sha256 ok" : "unchecked");
    fflush(stdout);
}

/* reference graphtest + setTargetSizes (framework/mainrun.c:361-410): decode
 * prefixes of growing length that end on a codeword boundary */
static void graphtest(struct decoder *d, struct dataset *s, int incs) {
    struct CompressedData cut = *s->td->cd;
    struct UnCompressedData plain;
    struct TestData red;
    red.name = s->td->name;
    red.cd = &cut;
    red.ucd = s->td->ucd ? &plain : NULL;
    /* one serial pass records, for every target, the last whole-codeword boundary */
    const struct CompressedData *cd = s->td->cd;
    int node = 0, nsym = 0, lastok = 0, next = incs;
    for (int pos = 0; pos < cd->bits; pos++) {
        if (pos == next) {
            cut.bits = lastok + 1;
            cut.uncompressedsize = nsym;
            if (s->td->ucd) { plain.data = s->td->ucd->data; plain.uncompressedsize = nsym; }
            struct evalresult r = evaluate_ex(d, &red, s->td->ucd != NULL, NULL);
            printf("%8d  %.9f  %.6f\n", next, r.min_seconds, r.min_device_ms);
            fflush(stdout);
            next += incs;
        }
        int bit = (cd->data[pos >> 3] >> (pos & 7)) & 1;
        node = bit ? cd->tree[node].ione : cd->tree[node].izero;
        if (cd->tree[node].izero == -1 && cd->tree[node].ione == -1) { nsym++; node = 0; lastok = pos; }
    }
}

/* BASELINE.json configs 4/5 over N GPUs of this process (B200_DEVICES=N): ONE stream, built
 * with the bundled generator, cut into byte-range shards, decoded resident (hb_multi_*),
 * every output slice verified on its device */
static void synth_multi(int kind, unsigned log2n, const char *label, int ndev) {
    hb_multi *m = NULL;
    int rc = hb_multi_create(NULL, ndev, &m);
    if (rc) { printf("hb_multi_create(%d): %s\n", ndev, hb_strerror(rc)); exit(-1); }
    const uint64_t n = 1ull << log2n, seed = 0x48554646ull;
    uint64_t bits = 0, bad = 0;
    if ((rc = hb_multi_generate(m, kind, seed, n, &bits))) {
        printf("hb_multi_generate: %s (%s)\n", hb_strerror(rc), hb_multi_last_error(m));
        exit(-1);
    }
    hb_multi_result res, best;
    memset(&best, 0, sizeof(best));
    for (int i = 0; i < 5; i++) {
        if ((rc = hb_multi_decode(m, &res))) {
            printf("hb_multi_decode: %s (%s)\n", hb_strerror(rc), hb_multi_last_error(m));
            exit(-1);
        }
        if (i == 0 || res.ms_device_max < best.ms_device_max) best = res;
    }
    if ((rc = hb_multi_verify(m, kind, seed, &bad))) { printf("hb_multi_verify: %s\n", hb_strerror(rc)); exit(-1); }
    const double ms = best.ms_device_max;
    printf("%17s %8s  %2d %.9f ms   | %.3f GB/s out, %.3f GB/s in (device, max over %d GPUs; wall %.3f ms), "
           "bits %llu, symbols %llu, mismatches %llu\n", "b200multi", label, best.n_devices, ms,
           (double)n / (ms * 1e-3) / 1e9, (double)((bits + 7) / 8) / (ms * 1e-3) / 1e9, best.n_devices,
           best.ms_wall, (unsigned long long)bits, (unsigned long long)best.n_symbols, (unsigned long long)bad);
    for (int i = 0; i < best.n_devices; i++)
        printf("%17s shard %d: %llu symbols, %.4f ms\n", "", i, (unsigned long long)best.shard_symbols[i], best.shard_ms[i]);
    if (bad || best.n_symbols != n) { fprintf(stderr, "problem with : b200multi\n"); exit(1); }
    hb_multi_destroy(m);
}

/* BASELINE.json configs 4/5 from the C side: build the stream on the device with
 * the bundled generator, decode it resident, verify every byte on the device */
static void synth(int kind, unsigned log2n, const char *label) {
    if (getenv("B200_DEVICES") && atoi(getenv("B200_DEVICES")) > 0) {
        synth_multi(kind, log2n, label, atoi(getenv("B200_DEVICES")));
        return;
    }
    hb_ctx *ctx = NULL;
    hb_model *m = (hb_model *)malloc(sizeof(*m));
    int rc = hb_ctx_create(0, NULL, &ctx);
    if (rc || !m) { printf("hb_ctx_create: %s\n", hb_strerror(rc)); exit(-1); }
    if ((rc = hb_model_build(kind, m))) { printf("hb_model_build: %s\n", hb_strerror(rc)); exit(-1); }
    const uint64_t n = 1ull << log2n, seed = 0x48554646ull;
    uint64_t bits = 0, bad = 0;
    void *d_comp = NULL, *d_out = NULL;
    hb_gen_count_bits_device(ctx, m, seed, 0, n, &bits);
    uint64_t cap = ((bits + 7) / 8 + 15) / 16 * 16 + 64;
    if (hb_dev_alloc(ctx, cap, &d_comp) || hb_dev_alloc(ctx, n + 64, &d_out)) { printf("device alloc failed\n"); exit(-1); }
    if ((rc = hb_gen_encode_device(ctx, m, seed, 0, n, d_comp, cap, &bits))) { printf("encode: %s\n", hb_strerror(rc)); exit(-1); }
    hb_codebook *cb = NULL;
    if ((rc = hb_codebook_create(ctx, m->tree, m->nodes, &cb))) { printf("codebook: %s\n", hb_strerror(rc)); exit(-1); }
    hb_result res;
    double best = -1;
    for (int i = 0; i < 5; i++) {
        if ((rc = hb_decode_device(ctx, cb, d_comp, cap, bits, d_out, n, &res))) {
            printf("decode: %s (%s)\n", hb_strerror(rc), hb_last_error(ctx));
            exit(-1);
        }
        if (best < 0 || res.ms_total < best) best = res.ms_total;
    }
    hb_gen_verify_device(ctx, m, seed, 0, n, d_out, &bad);
    printf("%17s %8s     %.9f ms   | %.3f GB/s out, %.3f GB/s in (device), bits %llu, max code length %u, "
           "symbols %llu, mismatches %llu\n", "b200", label, best, (double)n / (best * 1e-3) / 1e9,
           (double)((bits + 7) / 8) / (best * 1e-3) / 1e9, (unsigned long long)bits, m->maxlen,
           (unsigned long long)res.n_symbols, (unsigned long long)bad);
    if (bad || res.n_symbols != n) { fprintf(stderr, "problem with : b200\n"); exit(1); }
    hb_codebook_destroy(cb);
    hb_dev_free(ctx, d_comp);
    hb_dev_free(ctx, d_out);
    hb_ctx_destroy(ctx);
    free(m);
}

int main(int argc, char *argv[]) {
    const char *testname = argc > 1 ? argv[1] : "hello";
    if (argc > 2) g_dir = argv[2];
    else if (getenv("HUFF_FILES")) g_dir = getenv("HUFF_FILES");
    fprintf(stderr, "running test: %s\n", testname);

    /* this harness owns the buffers it hands to the approach and releases them in evaluate_ex:
     * let the approach page-lock them (B200_PIN=0 to measure without) */
    b200ApproachPinBuffers(!(getenv("B200_PIN") && getenv("B200_PIN")[0] == '0'));
    struct decoder *b200 = newDecoder(b200Approach, NULL, "b200");
    /* the multi-GPU approach takes its device count the way jumptable takes jumpbits
     * (framework/mainrun.c:442,500-501); 0 = every visible device */
    static int ndev = 0;
    if (getenv("B200_DEVICES")) ndev = atoi(getenv("B200_DEVICES"));
    struct decoder *b200multi = newDecoder(b200ApproachMulti, &ndev, "b200multi");
    struct decoder *onethread = newDecoder(onethreadApproach, NULL, "onethread");
#ifdef WITH_ORACLE_BASELINES
    static int jumpbits = 8;
    struct decoder *simpledec = newDecoder(simpleDecode, NULL, "simpleDecode");
    struct decoder *jumptable = newDecoder(jumptableApproach, &jumpbits, "jumptableApproach");
    /* the reference's remaining CPU approaches (framework/mainrun.c:
     * 496-501), taken UNMODIFIED from oracle/_ref/libref.so (make -C oracle ref) when that
     * library is present: same struct layout, same approach signature */
    struct decoder *refdec[5] = { NULL, NULL, NULL, NULL, NULL }, *refpes = NULL;
    int nref = 0;
    {
        const char *lib = getenv("HB_REF_LIB") ? getenv("HB_REF_LIB") : "../../oracle/_ref/libref.so";
        void *h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
        static const char *const names[5] = { "decodeBigtablev1", "decodeBigtableMultiSym",
                                              "decodeBigtableSimple", "linApproach", "jumptableApproach" };
        static const char *const labels[5] = { "ref:dbtV1", "ref:dbtMultiSym", "ref:dbtSimple",
                                               "ref:linApproach", "ref:jumptable" };
        for (int i = 0; h && i < 5; i++) {
            void *f = dlsym(h, names[i]);
            if (f) refdec[nref++] = newDecoder((void (*)(struct CompressedData *, struct UnCompressedData *, void *))f,
                                               i >= 3 ? &jumpbits : NULL, labels[i]);
        }
        if (h && dlsym(h, "pesApproach"))
            refpes = newDecoder((void (*)(struct CompressedData *, struct UnCompressedData *, void *))dlsym(h, "pesApproach"),
                                NULL, "ref:pes");
        if (!h) fprintf(stderr, "(no %s: reference CPU approaches not listed)\n", lib);
    }
#endif
    const char *suite_bigtable[] = { "paper1", "hello", "news", "kjv", "book2" };   /* order of mainrun.c:558-562 */
    const char *suite_all[] = { "hello", "paper1", "news", "book2", "world192", "bible", "kjv", "ecoli" };
    const char **suite = NULL;
    int ns = 0;
    const char *one[1];

    if (!strcmp(testname, "bigtable")) { suite = suite_bigtable; ns = 5; }
    else if (!strcmp(testname, "all")) { suite = suite_all; ns = 8; }
    else if (get_set(testname)) { one[0] = testname; suite = one; ns = 1; }
    /* the reference's sweeps (framework/mainrun.c:590-616): 2 = the GPU approach, 1 = the serial
     * decoder, 3 = decodeBigtableMultiSym, 4 = pes; the CPU ones need HuffFrameworkBaselines */
    else if (!strcmp(testname, "quickgraph") || !strcmp(testname, "quickgraph2")) { graphtest(b200, get_set("paper1"), 10000); }
    else if (!strcmp(testname, "graph") || !strcmp(testname, "graph2")) { graphtest(b200, get_set("kjv"), 500000); }
    else if (!strcmp(testname, "kjvprof")) { evalandshow(b200, get_set("kjv"), 1); }   /* one approach, one corpus: run it under a profiler */
    else if (!strcmp(testname, "opt")) { evalandshow(b200, get_set("kjv"), 1); evalandshow(b200multi, get_set("kjv"), 1); }
    else if (!strcmp(testname, "onethread")) { evalandshow(onethread, get_set("hello"), 1); evalandshow(onethread, get_set("paper1"), 1); }
    else if (!strcmp(testname, "multi")) {
        for (int i = 0; i < 8; i++) evalandshow(b200multi, get_set(suite_all[i]), 1);
    }
#ifdef WITH_ORACLE_BASELINES
    else if (!strcmp(testname, "quickgraph1")) { graphtest(simpledec, get_set("paper1"), 10000); }
    else if (!strcmp(testname, "graph1")) { graphtest(simpledec, get_set("kjv"), 500000); }
    else if (!strcmp(testname, "quickgraph3") && nref > 1) { graphtest(refdec[1], get_set("paper1"), 10000); }
    else if (!strcmp(testname, "graph3") && nref > 1) { graphtest(refdec[1], get_set("kjv"), 500000); }
    else if (!strcmp(testname, "graph4") && refpes) { graphtest(refpes, get_set("kjv"), 500000); }
    else if (!strcmp(testname, "bts") && nref > 2) {
        for (int i = 0; i < 5; i++) evalandshow(refdec[2], get_set(suite_bigtable[i]), 1);
    }
#else
    else if (!strcmp(testname, "quickgraph1") || !strcmp(testname, "quickgraph3") || !strcmp(testname, "graph1") ||
             !strcmp(testname, "graph3") || !strcmp(testname, "graph4") || !strcmp(testname, "bts")) {
        fprintf(stderr, "error exit: %s sweeps a CPU approach; this binary has no CPU decode path "
                        "(build and run HuffFrameworkBaselines)\n", testname);
        return 1;
    }
#endif
    else if (!strcmp(testname, "synth1g")) { synth(HB_MODEL_ENGLISH, 30, "synth1g"); }
    else if (!strcmp(testname, "synthfib")) { synth(HB_MODEL_FIBONACCI, 32, "synthfib"); }
    else if (!strcmp(testname, "synth16g")) { synth(HB_MODEL_FIBONACCI, 34, "synth16g"); }   /* BASELINE config 5 at its stated size */
    else { fprintf(stderr, "error exit: unknown test %s\n", testname); return 1; }

    if (suite) {
        for (int i = 0; i < ns; i++) infoTestData(get_set(suite[i])->td);
        for (int i = 0; i < ns; i++) evalandshow(b200, get_set(suite[i]), 1);
#ifdef WITH_ORACLE_BASELINES
        for (int i = 0; i < ns; i++) evalandshow(simpledec, get_set(suite[i]), 1);
        for (int i = 0; i < ns; i++) evalandshow(jumptable, get_set(suite[i]), 1);
        for (int k = 0; k < nref; k++)
            for (int i = 0; i < ns; i++) evalandshow(refdec[k], get_set(suite[i]), 1);
#endif
    }
    for (int i = 0; i < NSETS; i++) freeTestData(g_sets[i].td);
    freeDecoder(b200);
    freeDecoder(b200multi);
    freeDecoder(onethread);
#ifdef WITH_ORACLE_BASELINES
    if (refpes) freeDecoder(refpes);
    freeDecoder(simpledec);
    freeDecoder(jumptable);
    for (int k = 0; k < nref; k++) freeDecoder(refdec[k]);
#endif
    return 0;
}
