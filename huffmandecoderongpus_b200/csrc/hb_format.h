/* hb_format.h -- constants shared by the C table builder and the CUDA code */
#ifndef HB_FORMAT_H_
#define HB_FORMAT_H_
#define HB_MAX_CODELEN 32          /* longest supported codeword (bits) */
#define HB_LUT_LINK 0x80000000u    /* single-symbol LUT entry flag: continue in a sub-table */

/* Multi-symbol ("fast") tables, indexed by the next wf stream bits (LSB first).
 * From offset 0 of the index, consecutive codewords are decoded while they lie
 * entirely inside the wf bits.  Both tables pack, in their high half,
 *     [23:16] B    = bits consumed by those codewords
 *     [31:24] nsym = how many there are
 * so that `acc += entry >> 16` advances a packed (symbols << 8 | bit position)
 * accumulator in one add.  Low half:
 *   S-table (sync kernel): [15:0] bitmask of the codeword START offsets
 *   E-table (emit kernel): [7:0] first symbol, [15:8] second symbol (at most
 *                          HB_E_MAXSYM symbols per entry)
 * When not even the first codeword fits (code longer than wf bits) the entry is
 * the marker: nsym = 0, B = HB_FAST_MARK, low half 0; adding it pushes the
 * position byte past any legal value, which ends the probe loop.
 *   E64-table (emit kernel, word-granular stores), indexed by the next wf64 <= wf bits
 *   (wf64 = 11 when the code's implied mean codeword length is at most 3.5 bits: four
 *   codewords then usually fit and the table takes half the shared memory -- else 12):
 *   two u32 per index,
 *       lo: first .. fourth symbol, one byte each (at most HB_E64_MAXSYM)
 *       hi: [15:0]  PRMT selector that shifts nsym new bytes into a 4-byte window whose
 *                   newest byte is on top: 0x3210 + 0x1111 * nsym
 *           [21:16] 8 * nsym, [31:26] B (bits consumed)
 *                   marker: nsym = 0 (selector 0x3210), B = HB_E64_MARK, lo = 0 */
#define HB_FAST_MARK 0xE0u
#define HB_E_MAXSYM 2
#define HB_E64_MAXSYM 4
#define HB_E64_MARK 48u            /* > 31 + HB_WF_MAX: a position no real probe can reach */
#define HB_WF_MAX 12               /* widest fast-table index (16 KB per table) */

/* EP-table (flat emit kernel, hb_emitf_kernel): indexed by the next wfp stream bits; built by
 * every CTA straight into shared memory from the single-symbol table, R = 1 << rshift copies
 * interleaved entry by entry (copy r of entry x at byte (x << (3 + rshift)) + 8 r), so that
 * the lanes of a warp read from disjoint banks (R = 16: lane l uses copy l & 15 and an LDS.64
 * of a warp completes in its minimum of two wavefronts whatever the indices are).
 *     lo: first .. fourth symbol, one byte each (at most HB_E64_MAXSYM)
 *     hi: [15:0]  PRMT selector 0x3210 + 0x1111 * nsym
 *         [20:16] bits consumed (<= wfp), [26:21] 8 * nsym
 *         bit 31  marker: not even the first codeword fits wfp bits (nsym = 0, bits = 0) */
#define HB_EP_MARK 0x80000000u
#define HB_EP_WF_MIN 8
#define HB_EP_WF_MAX 12
#define HB_E32_WF_MAX 15          /* widest E32-table index (hb_emit32_kernel): 128 KB for one copy */

/* Byte-step transducer of the sync kernel's fast path (the GPU counterpart of the
 * reference's jump table, framework/jumptableapproach.c:40-99, with jumpbits = 8 and
 * no symbol output).  States are the internal nodes of the tree, root = state 0, at
 * most HB_FSM_MAX_STATES of them (every tree over a byte alphabet qualifies).
 *   fsm[s * 256 + b] (u16), b = the next 8 stream bits, bit 0 first:
 *       [15:8] state after the 8 bits, [3:0] number of codewords that END inside them
 *   fsm_depth[s]  = bits of the unfinished codeword already consumed in state s
 *   fsm_bstep[2 * s + bit] (u16): [7:0] next state, bit 8 = this bit ended a codeword
 *   fsm_pstep[(1 << r) + x] (u16), r = 1..7: the root's entry for a step of only r
 *       bits x (same packing) -- aligns a chain that starts inside a byte          */
#define HB_FSM_MAX_STATES 256
#endif
