/* hb_format.h -- constants shared by the C table builder and the CUDA code */
#ifndef HB_FORMAT_H_
#define HB_FORMAT_H_
#define HB_MAX_CODELEN 32          /* longest supported codeword (bits) */
#define HB_LUT_LINK 0x80000000u    /* LUT entry flag: continue in a sub-table */
#endif
