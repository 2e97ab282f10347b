#include "decodeUtil.h"
#include "b200approach.h"

#include <err.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct decoder *newDecoder(decoder_fn f, void *paramdata, const char *name) {
    struct decoder *d = (struct decoder *)malloc(sizeof(*d));
    if (!d) return NULL;
    d->decoder_function = f;
    d->paramdata = paramdata;
    d->name = name;
    return d;
}

void freeDecoder(struct decoder *d) { free(d); }

static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC_RAW, &ts);   /* same clock as framework/time.h:20 */
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int repeats(void) {
    const char *s = getenv("B200_REPEATS");
    int r = s ? atoi(s) : REPEATS;
    return r < 0 ? 0 : r;
}

struct evalresult evaluate_ex(struct decoder *d, struct TestData *td, int withcheck,
                              const char *want_sha256) {
    struct evalresult r = { 0.0, -1.0, 0 };
    const int is_b200 = d->decoder_function == (decoder_fn)b200Approach ||
                        d->decoder_function == (decoder_fn)b200ApproachMulti;
    struct UnCompressedData *out = newUnCompressedData(td->cd->uncompressedsize);
    if (!out) err(1, "out of memory");
    const int n = 1 + repeats();
    for (int i = 0; i < n; i++) {
        clearUnCompressedData(out);   /* stale data must not fake a pass */
        double t0 = now_seconds();
        d->decoder_function(td->cd, out, d->paramdata);
        double dt = now_seconds() - t0;
        if (i == 0 || dt < r.min_seconds) r.min_seconds = dt;
        if (is_b200) {
            double ms = b200ApproachLastDeviceMs();
            if (r.min_device_ms < 0 || ms < r.min_device_ms) r.min_device_ms = ms;
        }
        if (i == 0 && withcheck) {
            int bad = 0;
            out->uncompressedsize = td->cd->uncompressedsize;
            if (td->ucd) {
                bad = compareUnCompressedData(out, td->ucd) != 0;
                r.checked = 1;
            } else if (want_sha256) {
                char hex[65];
                digestUnCompressedData(out, hex);
                bad = strcmp(hex, want_sha256) != 0;
                if (bad) printf("sha256 %s, expected %s\n", hex, want_sha256);
                r.checked = 2;
            }
            if (bad) {
                fprintf(stderr, "problem with : %s\n", d->name);
                errx(1, "decode problem");
            }
        }
    }
    /* the approach may have page-locked `out` (b200ApproachPinBuffers): release it before the
     * memory goes back to the allocator */
    b200ApproachReleaseBuffers();
    freeUnCompressedData(out);
    return r;
}

double evaluate(struct decoder *d, struct TestData *td, int withcheck) {
    return evaluate_ex(d, td, withcheck, NULL).min_seconds;
}
