/*
 * hb_lut.h -- host-side code-table construction: Huffman tree (the reference's
 * node array, framework/huffdata.h:12-16) -> multi-level lookup table consumed
 * by the kernels (entry format in hb_core.cuh).  Plain C.
 */
#ifndef HB_LUT_H_
#define HB_LUT_H_

#include <stddef.h>
#include <stdint.h>
#include "huffb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* same layout as the reference's struct HuffNode (12 bytes, 3 bytes padding) */
typedef hb_node_abi hb_node;

typedef struct hb_lut {
    uint32_t *entries;   /* level 1 first (1 << w1 entries), then sub-tables */
    uint32_t  n_entries;
    uint32_t  w1;        /* level-1 index width in bits */
    uint32_t  maxlen;    /* longest codeword (reference tableHeight)   */
    uint32_t  minlen;    /* shortest codeword (reference tableMinDepth) */
    uint32_t  n_leaves;
    /* multi-symbol tables (hb_format.h), 1 << wf entries each */
    uint32_t  wf;
    uint32_t *stab;
    uint32_t *etab;
    /* canonical per-symbol code (first leaf found for each symbol), used by the
     * bundled encoder: code bits LSB-first in stream order */
    uint32_t  code[256];
    uint8_t   codelen[256];
    /* byte-step transducer (hb_format.h); fsm_states == 0 when the tree has more
     * than HB_FSM_MAX_STATES internal nodes (the probe-based sync path is used) */
    uint32_t  fsm_states;
    uint16_t *fsm;         /* fsm_states * 256 entries */
    uint16_t *fsm_bstep;   /* fsm_states * 2 entries */
    uint8_t   fsm_depth[256];
    uint16_t  fsm_pstep[256];
    uint32_t *e64;         /* E64-table: 2 << wf64 words (lo, hi interleaved) */
    /* transducer state numbering (breadth first, root = 0): node index of every state
     * and state of every node (-1 for leaves); node_state has `nodes` entries */
    int32_t   fsm_node[256];
    int32_t  *node_state;
    uint32_t  wf64;              /* index width of the E64-table (hb_format.h) */
    double    implied_avg_len;   /* sum over leaves of 2^-len * len */
    uint32_t  len_gcd;           /* gcd of all codeword lengths (1 for almost every real code) */
} hb_lut;

/* Validate the tree and build the table.  w1_max/w2_max cap the widths of the
 * first and of every deeper level (0 = defaults 11 / 10).
 * Returns 0 or a negative HB_ERR_* code (include/huffb200.h). */
int hb_lut_build(const hb_node *tree, int nodes, int w1_max, int w2_max, hb_lut *out);
/* Same, without the multi-symbol tables and the transducer table (stab, etab, e64,
 * fsm stay NULL; everything else is filled): the device builds those itself
 * (hb_build_tables_kernel), the host only validates and numbers the states. */
int hb_lut_build_small(const hb_node *tree, int nodes, int w1_max, int w2_max, hb_lut *out);
void hb_lut_free(hb_lut *lut);
size_t hb_lut_sizeof(void);

#ifdef __cplusplus
}
#endif
#endif
