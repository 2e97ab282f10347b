/*
 * huff_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, 64-bit positions) of the reference's serial Huffman
 * decode path.  It is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, link or call anything in oracle/.  Nothing in the product
 * (huffmandecoderongpus_b200/, include/) depends on it.
 *
 * Parity pin: every function here is checked by tests/test_oracle.py against
 *   (1) the hello golden vector (reference framework/mainrun.c:659-663),
 *   (2) the SHA-256 digests of the reference's own simpleDecode output on all
 *       eight shipped .huff corpora, and
 *   (3) the UNMODIFIED reference objects compiled into oracle/_ref/libref.so
 *       (oracle/Makefile), when that library is present.
 *
 * Each function cites the reference file:line it restates.
 */
#ifndef HUFF_ORACLE_H_
#define HUFF_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* In-memory node, same field meaning as reference framework/huffdata.h:12-16:
 * node 0 is the root; a leaf has izero == ione == -1; bit 0 selects izero. */
typedef struct ora_node {
    uint8_t sym;
    int32_t izero;
    int32_t ione;
} ora_node;

typedef struct ora_stream {
    int32_t   nodes;
    uint64_t  bits;   /* exact number of valid bits, LSB-first inside each byte */
    uint64_t  usize;  /* decoded length in bytes */
    ora_node *tree;
    uint8_t  *data;   /* ceil(bits/8) bytes followed by >= 16 zero bytes */
    int       wide;   /* 0: "HUFF" 32-bit header, 1: "HUF8" 64-bit header */
} ora_stream;

/* reference framework/huffdata.c:27-68 (loadHuffFile); also accepts the
 * 64-bit "HUF8" container this repo adds for > 2^31-bit streams.
 * Returns NULL on any format error. */
ora_stream *ora_load_huff(const char *path);
void ora_free_stream(ora_stream *s);

/* reference framework/huffdata.c:224-238,272-278 */
int ora_tree_height(const ora_node *tree, int r);
int ora_tree_mindepth(const ora_node *tree, int r);
int ora_tree_size(const ora_node *tree, int r);

/* reference framework/mainrun.c:38-55 (simpleDecode): bit-serial tree walk.
 * Writes at most outcap bytes; returns the number of symbols the walk emits. */
uint64_t ora_simple_decode(const ora_node *tree, const uint8_t *data,
                           uint64_t bits, uint8_t *out, uint64_t outcap);

/* reference framework/jumptableapproach.c:40-99 (makejumptables) and :109-265
 * (jumptableApproach): finite-state decoder consuming jumpbits bits per step.
 * Table construction is inside the call, as in the reference.  Returns the
 * number of symbols written, or (uint64_t)-1 if jumpbits is unsupported. */
uint64_t ora_jumptable_decode(const ora_node *tree, int nodes,
                              const uint8_t *data, uint64_t bits, int jumpbits,
                              uint8_t *out, uint64_t outcap);

/* reference framework/mainrun.c:361-385 (setTargetSizes): longest prefix of
 * whole codewords that fits in targetbits.  Outputs its bit length and symbol
 * count. */
void ora_prefix_sizes(const ora_node *tree, const uint8_t *data,
                      uint64_t targetbits, uint64_t *bits_out,
                      uint64_t *usize_out);

/* reference framework/pes.c:30-104, restated with 64-bit indices and a
 * shrinking number of doubling levels kept in memory (two at a time are
 * enough for the bottom-up pass; the top-down pass recomputes).  Only meant
 * for small inputs (O(bits * log bits) work, 13 bytes of scratch per bit per
 * level).  Returns the decoded length. */
uint64_t ora_pes_decode(const ora_node *tree, const uint8_t *data,
                        uint64_t bits, uint8_t *out, uint64_t outcap);

/* Per-bit-offset phase statement, reference framework/pes.c:30-46
 * (decodeAllBits): for every bit offset b < bits the codeword length found by
 * walking from the root (clamped at the end of the stream) and the symbol of
 * the node reached. */
void ora_decode_all_bits(const ora_node *tree, const uint8_t *data,
                         uint64_t bits, uint8_t *sym_out, int32_t *len_out);

#ifdef __cplusplus
}
#endif
#endif
