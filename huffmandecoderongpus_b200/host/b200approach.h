/*
 * b200approach.h -- the drop-in approach entry point.
 *
 * Same shape as every approach of the reference (framework/decodeUtil.h:15-16,
 * e.g. framework/fastgpu.h:13-15): register it with
 *     newDecoder(b200Approach, NULL, "b200")
 * in framework/mainrun.c's approach table and link libhuffb200.so.
 */
#ifndef B200APPROACH_H_
#define B200APPROACH_H_

#ifdef __cplusplus
extern "C" {
#endif

struct CompressedData;
struct UnCompressedData;
struct CompressedDataL;
struct UnCompressedDataL;

/* Decode cd into uncompressed->data (host memory, caller-allocated
 * uncompressedsize+3 bytes, as framework/decodeUtil.c:37 does).  paramdata is
 * ignored (NULL in the reference for GPU approaches).  Errors: message on
 * stdout and exit(-1), like framework/fastgpu.cu:16-31. */
void b200Approach(struct CompressedData *cd, struct UnCompressedData *uncompressed,
                  void *paramdata);

/* 64-bit variant for streams beyond 2^31-1 bits */
void b200ApproachL(struct CompressedDataL *cd, struct UnCompressedDataL *uncompressed,
                   void *paramdata);

/* The same stream over several GPUs of this process (byte-range shards, maps exchanged by
 * peer copies; include/huffb200.h hb_multi_*).  paramdata: int * = number of devices, NULL =
 * the B200_DEVICES environment variable, else every visible device.  Registered next to
 * b200Approach: newDecoder(b200ApproachMulti, &ndev, "b200multi"). */
void b200ApproachMulti(struct CompressedData *cd, struct UnCompressedData *uncompressed,
                       void *paramdata);
void b200ApproachMultiL(struct CompressedDataL *cd, struct UnCompressedDataL *uncompressed,
                        void *paramdata);

/* the reference's debug approach of that name (framework/onethread.cu:33-52, registered at
 * framework/mainrun.c:480): the whole stream on one device thread */
void onethreadApproach(struct CompressedData *cd, struct UnCompressedData *uncompressed, void *paramdata);

/* Optional: keep the caller's two buffers page-locked across the 26 back-to-back calls of
 * evaluate() (DMA copies instead of staged pageable ones; kjv 0.6 -> 0.27 ms per call).  A
 * program that turns this on must call b200ApproachReleaseBuffers() before it frees the
 * buffers it passed in.  Also enabled by B200_PIN=1 in the environment. */
void b200ApproachPinBuffers(int on);
void b200ApproachReleaseBuffers(void);

/* device milliseconds (CUDA events, kernels only) of the last call, and the
 * number of symbols it produced */
double b200ApproachLastDeviceMs(void);
unsigned long long b200ApproachLastSymbols(void);
/* release the cached context (optional; also run at process exit) */
void b200ApproachShutdown(void);

#ifdef __cplusplus
}
#endif
#endif
