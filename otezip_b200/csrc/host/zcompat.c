/* zcompat.c — the zlib-named entry points the otezip CLI links against for its single-stream gzip
 * modes (-d / -g; /root/reference/src/main.c:31-36, :590-832).  Those modes are not part of the
 * batched archive hot path ("next"): the symbols exist so that the unchanged CLI
 * relinks against libotezip_b200.so, and they fail loudly instead of decoding on the CPU.
 */
#include <stdio.h>

#include "otezip/zstream.h"

#define Z_STREAM_ERROR (-2)

static int not_on_gpu_path(const char *what) {
	fprintf (stderr, "otezip-b200: %s: single-stream gzip modes are not on the GPU path of this build\n", what);
	return Z_STREAM_ERROR;
}

int inflateInit2(z_stream *strm, int windowBits) {
	(void)strm;
	(void)windowBits;
	return not_on_gpu_path ("inflateInit2");
}
int inflateInit2_(z_stream *strm, int windowBits, const char *version, int stream_size) {
	(void)version;
	(void)stream_size;
	return inflateInit2 (strm, windowBits);
}
int inflate(z_stream *strm, int flush) {
	(void)strm;
	(void)flush;
	return not_on_gpu_path ("inflate");
}
int inflateEnd(z_stream *strm) {
	(void)strm;
	return Z_STREAM_ERROR;
}
int deflateInit2(z_stream *strm, int level, int method, int windowBits, int memLevel, int strategy) {
	(void)strm;
	(void)level;
	(void)method;
	(void)windowBits;
	(void)memLevel;
	(void)strategy;
	return not_on_gpu_path ("deflateInit2");
}
int deflateInit2_(z_stream *strm, int level, int method, int windowBits, int memLevel, int strategy, const char *version, int stream_size) {
	(void)version;
	(void)stream_size;
	return deflateInit2 (strm, level, method, windowBits, memLevel, strategy);
}
int deflate(z_stream *strm, int flush) {
	(void)strm;
	(void)flush;
	return not_on_gpu_path ("deflate");
}
int deflateEnd(z_stream *strm) {
	(void)strm;
	return Z_STREAM_ERROR;
}
