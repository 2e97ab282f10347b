/*
 * huffdata.h -- host data model of the C harness.
 *
 * struct HuffNode / CompressedData / UnCompressedData / TestData keep the
 * field order, types and meaning of the reference's framework/huffdata.h:12-37,
 * so an approach function compiled here can be registered in the reference's
 * own approach table and vice versa.  The reference's header
 * fields are 32-bit; streams beyond 2^31-1 bits use the *L ("long") twins.
 */
#ifndef B200_HUFFDATA_H_
#define B200_HUFFDATA_H_

#include <stdint.h>

struct HuffNode {
    unsigned char sym;
    int izero;
    int ione;
};

struct CompressedData {
    int bits;
    int nodes;
    int uncompressedsize;
    struct HuffNode *tree;
    unsigned char *data;
};

struct UnCompressedData {
    int uncompressedsize;
    unsigned char *data;
};

struct TestData {
    struct CompressedData *cd;
    struct UnCompressedData *ucd;
    char *name;
};

/* 64-bit twins for the GB-scale synthetic configurations */
struct CompressedDataL {
    uint64_t bits;
    int nodes;
    uint64_t uncompressedsize;
    struct HuffNode *tree;
    unsigned char *data;
};

struct UnCompressedDataL {
    uint64_t uncompressedsize;
    unsigned char *data;
};

/* loaders: same contracts as the reference (framework/huffdata.c:27-68,152-222)
 * but they return NULL instead of crashing on a missing file */
struct CompressedData *loadHuffFile(const char *filename);
struct CompressedDataL *loadHuffFileL(const char *filename);   /* HUFF or HUF8 */
void freeCompressedData(struct CompressedData *cd);
void freeCompressedDataL(struct CompressedDataL *cd);
struct UnCompressedData *loadTextFile(const char *filename);
struct UnCompressedData *newUnCompressedData(int size);
void freeUnCompressedData(struct UnCompressedData *ucd);
void clearUnCompressedData(struct UnCompressedData *ucd);
int compareUnCompressedData(struct UnCompressedData *a, struct UnCompressedData *b);
struct TestData *loadTestData(const char *filename, const char *name);

void freeTestData(struct TestData *td);
void infoCompressedData(struct CompressedData *cd);
void infoTestData(struct TestData *td);
int tableHeight(struct HuffNode *tree, int r);
int tableMinDepth(struct HuffNode *tree, int r);
int treeSize(struct HuffNode *tree, int r);
int digestUnCompressedData(struct UnCompressedData *u, char hex[65]);

#endif
