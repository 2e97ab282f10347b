/*
 * ref_harness_main.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A main() that drives the UNMODIFIED reference harness -- framework/decodeUtil.c
 * (newDecoder, evaluate: 1 checked + 25 timed runs), framework/huffdata.c (loadTestData,
 * compareUnCompressedData) and framework/timing.c, compiled where they lie under
 * /root/reference by `make -C oracle refharness` -- with b200Approach registered the way
 * framework/mainrun.c:480-501 registers its approaches, on the five corpora of the bigtable
 * suite (framework/mainrun.c:558-562).  It proves the drop-in claim: reference structs,
 * reference loader, reference byte comparison, our approach function linked from
 * libhuffb200.so.  The reference's own main (framework/mainrun.c) is not used because it
 * hard-codes ../../files and has no b200 entry; nothing of it is copied here.
 *
 *   ref_harness <files-dir>        (files-dir holds NAME and NAME.huff)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "decodeUtil.h"   /* the reference's own headers (-I /root/reference/framework) */
#include "huffdata.h"

void b200Approach(struct CompressedData *cd, struct UnCompressedData *uncompressed, void *paramdata);
void b200ApproachMulti(struct CompressedData *cd, struct UnCompressedData *uncompressed, void *paramdata);

int main(int argc, char *argv[]) {
    const char *dir = argc > 1 ? argv[1] : "files";
    static const char *const names[5] = { "paper1", "hello", "news", "kjv.txt", "book2" };
    static const char *const labels[5] = { "paper1", "hello", "news", "kjv", "book2" };
    struct decoder *b200 = newDecoder(b200Approach, NULL, "b200");
    struct decoder *multi = newDecoder(b200ApproachMulti, NULL, "b200multi");
    for (int k = 0; k < 2; k++) {
        struct decoder *d = k ? multi : b200;
        for (int i = 0; i < 5; i++) {
            char path[1024];
            snprintf(path, sizeof(path), "%s/%s", dir, names[i]);
            struct TestData *td = loadTestData(path, (char *)labels[i]);
            /* evaluate() compares the first decode with the plaintext and exits with
             * err(1, "decode problem") on any difference (framework/decodeUtil.c:47-52) */
            double s = evaluate(d, td, 1);
            printf("%17s %8s     %.9f ms\n", d->name, td->name, s * 1000.0);   /* framework/mainrun.c:417-418 */
            fflush(stdout);
            freeTestData(td);
        }
    }
    freeDecoder(b200);
    freeDecoder(multi);
    return 0;
}
