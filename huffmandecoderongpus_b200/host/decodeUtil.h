/*
 * decodeUtil.h -- approach ("decoder") objects and the evaluation protocol of
 * the reference harness (framework/decodeUtil.h:14-31): an approach is
 *   void f(struct CompressedData *cd, struct UnCompressedData *out, void *param)
 * registered with newDecoder() and timed by evaluate().
 */
#ifndef B200_DECODEUTIL_H_
#define B200_DECODEUTIL_H_

#include "huffdata.h"

typedef void (*decoder_fn)(struct CompressedData *cd, struct UnCompressedData *uncompressed,
                           void *paramdata);

struct decoder {
    decoder_fn decoder_function;
    void *paramdata;
    const char *name;
};

struct decoder *newDecoder(decoder_fn f, void *paramdata, const char *name);
void freeDecoder(struct decoder *d);

#define REPEATS 25   /* reference framework/decodeUtil.h:26; override with env B200_REPEATS */

struct evalresult {
    double min_seconds;     /* min wall time over 1 + repeats runs (the reference's number) */
    double min_device_ms;   /* min CUDA-event time, when the approach reports one (< 0 otherwise) */
    int checked;            /* 1: bytes compared with the plaintext, 2: SHA-256 compared, 0: unchecked */
};

/* reference framework/decodeUtil.c:30-70: zero the output, run once and check
 * against the plaintext (or its SHA-256 when the plaintext is not shipped),
 * then `repeats` more zeroed runs; abort with err(1) on a mismatch. */
double evaluate(struct decoder *d, struct TestData *td, int withcheck);
struct evalresult evaluate_ex(struct decoder *d, struct TestData *td, int withcheck,
                              const char *want_sha256);

#endif
