/* otezip/zlib.h — minimal zlib-compatible declarations for consumers that include <zlib.h>-style code against this
 * library (reference: src/include/otezip/zlib.h:1-108; same names, constants and z_stream layout).  The six functions
 * are exported by libotezip_b200.so (otezip_b200/csrc/host/zcompat.c): one-shot streams decoded / encoded by the GPU
 * kernels, as the CLI's -d / -g modes use them (src/main.c:590-832). */
#ifndef OTEZIP_ZLIB_H
#define OTEZIP_ZLIB_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "zstream.h"   /* Bytef, uInt, uLong, voidpf, z_stream (field order is ABI) */

typedef voidpf (*alloc_func)(voidpf opaque, uInt items, uInt size);
typedef void (*free_func)(voidpf opaque, voidpf address);
typedef z_stream *z_streamp;

#define Z_NULL 0

/* return codes */
#define Z_OK 0
#define Z_STREAM_END 1
#define Z_NEED_DICT 2
#define Z_ERRNO (-1)
#define Z_STREAM_ERROR (-2)
#define Z_DATA_ERROR (-3)
#define Z_MEM_ERROR (-4)
#define Z_BUF_ERROR (-5)
#define Z_VERSION_ERROR (-6)

/* flush values */
#define Z_NO_FLUSH 0
#define Z_PARTIAL_FLUSH 1
#define Z_SYNC_FLUSH 2
#define Z_FULL_FLUSH 3
#define Z_FINISH 4
#define Z_BLOCK 5
#define Z_TREES 6

/* levels, strategies, data types */
#define Z_NO_COMPRESSION 0
#define Z_BEST_SPEED 1
#define Z_BEST_COMPRESSION 9
#define Z_DEFAULT_COMPRESSION (-1)
#define Z_FILTERED 1
#define Z_HUFFMAN_ONLY 2
#define Z_RLE 3
#define Z_FIXED 4
#define Z_DEFAULT_STRATEGY 0
#define Z_BINARY 0
#define Z_TEXT 1
#define Z_ASCII Z_TEXT
#define Z_UNKNOWN 2

#define MAX_WBITS 15
#define Z_DEFLATED 8

#ifdef __cplusplus
extern "C" {
#endif
int inflateInit2_(z_streamp strm, int windowBits, const char *version, int stream_size);
int inflate(z_streamp strm, int flush);
int inflateEnd(z_streamp strm);
int deflateInit2_(z_streamp strm, int level, int method, int windowBits, int memLevel, int strategy, const char *version,
	int stream_size);
int deflate(z_streamp strm, int flush);
int deflateEnd(z_streamp strm);
#ifdef __cplusplus
}
#endif

#define ZLIB_VERSION "1.2.11"
#define inflateInit(strm) inflateInit2_((strm), MAX_WBITS, ZLIB_VERSION, (int)sizeof(z_stream))
#define inflateInit2(strm, windowBits) inflateInit2_((strm), (windowBits), ZLIB_VERSION, (int)sizeof(z_stream))
#define deflateInit(strm, level) deflateInit2_((strm), (level), Z_DEFLATED, MAX_WBITS, 8, Z_DEFAULT_STRATEGY, ZLIB_VERSION, (int)sizeof(z_stream))
#define deflateInit2(strm, level, method, windowBits, memLevel, strategy) \
	deflateInit2_((strm), (level), (method), (windowBits), (memLevel), (strategy), ZLIB_VERSION, (int)sizeof(z_stream))

#endif /* OTEZIP_ZLIB_H */
