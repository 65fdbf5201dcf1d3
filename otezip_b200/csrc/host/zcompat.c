/* zcompat.c — the zlib-named entry points the otezip CLI links against for its single-stream gzip modes
 * (-d / -g; /root/reference/src/main.c:31-36, :590-832), served by the same GPU kernels as the archive path
 * (SURVEY.md §8f rank 1).  Reference side: inflateInit2 / inflate / inflateEnd in
 * src/lib/deflate-dec.inc.c:452-843 (wrapper detection :463-477, gzip / zlib header skipping :361-443, no
 * trailer check) and deflateInit2 / deflate / deflateEnd in src/lib/deflate-enc.inc.c:199-541.
 *
 * These are ONE-SHOT implementations: the first inflate() call takes everything at next_in as the complete
 * stream, decodes it as a batch of one on the GPU and then hands the bytes out across calls as avail_out
 * allows (Z_BUF_ERROR while output space is missing, Z_STREAM_END once everything is delivered — the
 * protocol main.c:617-660 drives).  deflate() requires Z_FINISH with the whole input.  No CPU codec: without
 * a device they return Z_STREAM_ERROR.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "otezip/zstream.h"
#include "otz_gpu.h"

#define Z_OK 0
#define Z_STREAM_END 1
#define Z_STREAM_ERROR (-2)
#define Z_DATA_ERROR (-3)
#define Z_MEM_ERROR (-4)
#define Z_BUF_ERROR (-5)
#define Z_FINISH 4

otz_ctx *otezip_b200_ctx(void);

enum { WRAP_RAW, WRAP_ZLIB, WRAP_GZIP, WRAP_AUTO };

struct zc_state {
	int inflating;
	int wrap;
	uint8_t *out;       /* fully decoded stream (inflate) */
	uint64_t out_len, served;
	int decoded;
};

static int wrap_of(int windowBits) { /* dec:463-477 */
	if (windowBits < 0) {
		return WRAP_RAW;
	}
	if (windowBits >= 40) {
		return WRAP_AUTO;
	}
	if (windowBits >= 24) {
		return WRAP_GZIP;
	}
	return WRAP_ZLIB;
}

static int new_state(z_stream *strm, int inflating, int windowBits) {
	if (!strm) {
		return Z_STREAM_ERROR;
	}
	struct zc_state *st = (struct zc_state *)calloc (1, sizeof (*st));
	if (!st) {
		return Z_MEM_ERROR;
	}
	st->inflating = inflating;
	st->wrap = wrap_of (windowBits);
	strm->state = st;
	strm->total_in = 0;
	strm->total_out = 0;
	return Z_OK;
}

int inflateInit2(z_stream *strm, int windowBits) {
	return new_state (strm, 1, windowBits);
}
int inflateInit2_(z_stream *strm, int windowBits, const char *version, int stream_size) {
	(void)version;
	(void)stream_size;
	return inflateInit2 (strm, windowBits);
}

/* dec:361-416 */
static long skip_gzip(const uint8_t *b, size_t n) {
	if (n < 10 || b[0] != 0x1f || b[1] != 0x8b || b[2] != 8) {
		return -1;
	}
	uint8_t fl = b[3];
	size_t pos = 10;
	if (fl & 0x04) {
		if (pos + 2 > n) {
			return -1;
		}
		pos += 2 + (size_t)(b[pos] | (b[pos + 1] << 8));
		if (pos > n) {
			return -1;
		}
	}
	for (int k = 0; k < 2; k++) {
		if (fl & (k == 0 ? 0x08 : 0x10)) {
			while (pos < n && b[pos]) {
				pos++;
			}
			if (pos >= n) {
				return -1;
			}
			pos++;
		}
	}
	if (fl & 0x02) {
		pos += 2;
		if (pos > n) {
			return -1;
		}
	}
	return (long)pos;
}

/* dec:419-443 */
static long skip_zlib(const uint8_t *b, size_t n) {
	if (n < 2 || (b[0] & 0x0f) != 8 || ((b[0] << 8) | b[1]) % 31 != 0) {
		return -1;
	}
	size_t pos = 2;
	if (b[1] & 0x20) {
		pos += 4;
		if (pos > n) {
			return -1;
		}
	}
	return (long)pos;
}

/* decode the raw stream [p, p+n) as a batch of one: a 30-byte local header in front makes it an "archive" */
static int gpu_inflate_all(struct zc_state *st, const uint8_t *p, size_t n) {
	otz_ctx *ctx = otezip_b200_ctx ();
	if (!ctx) {
		return Z_STREAM_ERROR;
	}
	void *img = NULL;
	if (otz_host_alloc (n + 30 + 64, &img) != OTZ_SUCCESS) {
		return Z_MEM_ERROR;
	}
	uint8_t *im = (uint8_t *)img;
	memset (im, 0, 30);
	im[0] = 0x50;
	im[1] = 0x4b;
	im[2] = 0x03;
	im[3] = 0x04;
	memcpy (im + 30, p, n);
	otz_extract_opts o = { 1, 0, 0, 0 }; /* no zip-bomb rule on a bare stream */
	int rc = Z_DATA_ERROR;
	uint64_t cap = n * 4 < 65536 ? 65536 : (uint64_t)n * 4;
	for (;;) {
		if (cap > 0x7fffffffULL) {
			cap = 0x7fffffffULL;
		}
		void *out = NULL;
		if (otz_host_alloc (cap + 64, &out) != OTZ_SUCCESS) {
			rc = Z_MEM_ERROR;
			break;
		}
		otz_entry e;
		memset (&e, 0, sizeof (e));
		e.comp_size = (uint32_t)n;
		e.uncomp_size = (uint32_t)cap;
		e.method = OTZ_M_DEFLATE;
		uint32_t crc = 0, produced = 0;
		int32_t status = 0;
		if (otz_extract_host_ex (ctx, im, n + 30, &e, 1, &o, (uint8_t *)out, cap, &crc, &status, &produced) != OTZ_SUCCESS) {
			fprintf (stderr, "otezip-b200: %s\n", otz_last_error ());
			otz_host_free (out);
			rc = Z_STREAM_ERROR;
			break;
		}
		if (OTZ_ST_CODE (status) == OTZ_ST_OK) {
			st->out = (uint8_t *)malloc (produced ? produced : 1);
			if (!st->out) {
				otz_host_free (out);
				rc = Z_MEM_ERROR;
				break;
			}
			memcpy (st->out, out, produced);
			st->out_len = produced;
			otz_host_free (out);
			rc = Z_OK;
			break;
		}
		otz_host_free (out);
		if (OTZ_ST_CODE (status) == OTZ_ST_OVERFLOW && cap < 0x7fffffffULL) {
			cap *= 4; /* the guess was too small: decode again into a larger arena */
			continue;
		}
		rc = OTZ_ST_CODE (status) == OTZ_ST_TRUNCATED ? Z_BUF_ERROR : Z_DATA_ERROR;
		break;
	}
	otz_host_free (img);
	return rc;
}

int inflate(z_stream *strm, int flush) {
	(void)flush;
	if (!strm || !strm->state) {
		return Z_STREAM_ERROR;
	}
	struct zc_state *st = (struct zc_state *)strm->state;
	if (!st->inflating) {
		return Z_STREAM_ERROR;
	}
	if (!st->decoded) {
		const uint8_t *p = strm->next_in;
		size_t n = strm->avail_in;
		long skip = 0;
		if (st->wrap == WRAP_GZIP || (st->wrap == WRAP_AUTO && n >= 2 && p[0] == 0x1f && p[1] == 0x8b)) {
			skip = skip_gzip (p, n); /* dec:557-573 */
		} else if (st->wrap != WRAP_RAW) {
			skip = skip_zlib (p, n);
		}
		if (skip < 0) {
			return Z_DATA_ERROR;
		}
		int rc = gpu_inflate_all (st, p + skip, n - (size_t)skip);
		if (rc != Z_OK) {
			return rc;
		}
		st->decoded = 1;
		strm->next_in += n; /* the reference reads no trailer either; everything counts as consumed */
		strm->total_in += n;
		strm->avail_in = 0;
	}
	uint64_t left = st->out_len - st->served;
	uint64_t k = left < strm->avail_out ? left : strm->avail_out;
	memcpy (strm->next_out, st->out + st->served, k);
	st->served += k;
	strm->next_out += k;
	strm->avail_out -= (uInt)k;
	strm->total_out += k;
	return st->served == st->out_len ? Z_STREAM_END : Z_BUF_ERROR;
}

int inflateEnd(z_stream *strm) {
	if (!strm || !strm->state) {
		return Z_STREAM_ERROR;
	}
	struct zc_state *st = (struct zc_state *)strm->state;
	free (st->out);
	free (st);
	strm->state = NULL;
	return Z_OK;
}

int deflateInit2(z_stream *strm, int level, int method, int windowBits, int memLevel, int strategy) {
	(void)level;
	(void)memLevel;
	(void)strategy;
	if (method != 8) {
		return Z_STREAM_ERROR;
	}
	int rc = new_state (strm, 0, windowBits);
	if (rc == Z_OK && ((struct zc_state *)strm->state)->wrap == WRAP_ZLIB) {
		/* a zlib wrapper needs an Adler-32 trailer, which is not on this path */
		free (strm->state);
		strm->state = NULL;
		return Z_STREAM_ERROR;
	}
	return rc;
}
int deflateInit2_(z_stream *strm, int level, int method, int windowBits, int memLevel, int strategy, const char *version, int stream_size) {
	(void)version;
	(void)stream_size;
	return deflateInit2 (strm, level, method, windowBits, memLevel, strategy);
}

static void le32(uint8_t *p, uint32_t v) {
	p[0] = (uint8_t)v;
	p[1] = (uint8_t)(v >> 8);
	p[2] = (uint8_t)(v >> 16);
	p[3] = (uint8_t)(v >> 24);
}

int deflate(z_stream *strm, int flush) {
	if (!strm || !strm->state || ((struct zc_state *)strm->state)->inflating) {
		return Z_STREAM_ERROR;
	}
	if (flush != Z_FINISH) {
		return Z_STREAM_ERROR; /* one-shot only (main.c:697 passes Z_FINISH) */
	}
	struct zc_state *st = (struct zc_state *)strm->state;
	otz_ctx *ctx = otezip_b200_ctx ();
	if (!ctx) {
		return Z_STREAM_ERROR;
	}
	const uint32_t n = strm->avail_in;
	void *in = NULL, *out = NULL;
	if (otz_host_alloc ((uint64_t)n + 64, &in) != OTZ_SUCCESS || otz_host_alloc ((uint64_t)n + 64, &out) != OTZ_SUCCESS) {
		if (in) {
			otz_host_free (in);
		}
		return Z_MEM_ERROR;
	}
	memcpy (in, strm->next_in, n);
	uint64_t in_ofs = 0, out_ofs = 0, total = 0;
	uint32_t in_len = n, out_size = 0, crc = 0;
	uint16_t method = OTZ_M_DEFLATE, method_out = 0;
	int rc = Z_STREAM_ERROR;
	if (otz_deflate_host (ctx, (const uint8_t *)in, n, &in_ofs, &in_len, &method, 1, (uint8_t *)out, n, &out_ofs, &out_size, &crc, &method_out,
		&total) == OTZ_SUCCESS) {
		const int gz = st->wrap == WRAP_GZIP;
		/* an incompressible source comes back as STORE: frame it as stored DEFLATE blocks (container work only) */
		uint64_t body = method_out == OTZ_M_DEFLATE ? out_size : (uint64_t)n + 5ull * ((n + 65534u) / 65535u) + (n ? 0 : 5);
		uint64_t need = body + (gz ? 18 : 0);
		if (need > strm->avail_out) {
			rc = Z_BUF_ERROR;
		} else {
			uint8_t *w = strm->next_out;
			if (gz) {
				static const uint8_t hdr[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3 };
				memcpy (w, hdr, 10);
				w += 10;
			}
			if (method_out == OTZ_M_DEFLATE) {
				memcpy (w, (uint8_t *)out + out_ofs, out_size);
				w += out_size;
			} else {
				const uint8_t *src = (const uint8_t *)in;
				uint32_t left = n;
				do {
					uint32_t k = left > 65535u ? 65535u : left;
					*w++ = (uint8_t)(left == k);
					*w++ = (uint8_t)k;
					*w++ = (uint8_t)(k >> 8);
					*w++ = (uint8_t)~k;
					*w++ = (uint8_t)(~k >> 8);
					memcpy (w, src, k);
					w += k;
					src += k;
					left -= k;
				} while (left);
			}
			if (gz) {
				le32 (w, crc);
				le32 (w + 4, n);
				w += 8;
			}
			strm->total_out += (uLong)(w - strm->next_out);
			strm->avail_out -= (uInt)(w - strm->next_out);
			strm->next_out = w;
			strm->next_in += n;
			strm->total_in += n;
			strm->avail_in = 0;
			rc = Z_STREAM_END;
		}
	} else {
		fprintf (stderr, "otezip-b200: %s\n", otz_last_error ());
	}
	otz_host_free (in);
	otz_host_free (out);
	return rc;
}

int deflateEnd(z_stream *strm) {
	if (!strm || !strm->state) {
		return Z_STREAM_ERROR;
	}
	free (strm->state);
	strm->state = NULL;
	return Z_OK;
}
