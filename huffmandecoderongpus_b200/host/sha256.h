/* sha256.h -- minimal SHA-256 (FIPS 180-4) for the harness's digest checks of
 * corpora whose plaintext is not shipped (kjv.txt, E.coli). */
#ifndef B200_SHA256_H_
#define B200_SHA256_H_
#include <stddef.h>
#include <stdint.h>
void sha256_hex(const unsigned char *data, size_t len, char out_hex[65]);
#endif
