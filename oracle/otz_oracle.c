/* otz_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see otz_oracle.h).
 *
 * Plain-C restatement of the reference hot path.  Citations are
 * /root/reference-relative: dec = src/lib/deflate-dec.inc.c,
 * zstd = src/lib/zstd.inc.c, crc = src/lib/crc32.inc.c, otezip.c = src/lib/otezip.c.
 */
#include "otz_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ CRC-32 */

static uint32_t crc_tab[256];
static int crc_tab_ready;

/* crc:3-36 is the byte-wise table of the reflected polynomial 0xEDB88320. */
static void crc_tab_init(void) {
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t c = i;
		for (int k = 0; k < 8; k++) {
			c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
		}
		crc_tab[i] = c;
	}
	crc_tab_ready = 1;
}

/* crc:40-47 */
uint32_t otzo_crc32(uint32_t crc, const void *buf, size_t len) {
	if (!crc_tab_ready) {
		crc_tab_init ();
	}
	const uint8_t *p = (const uint8_t *)buf;
	crc = ~crc;
	for (size_t i = 0; i < len; i++) {
		crc = crc_tab[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
	}
	return ~crc;
}

/* ----------------------------------------------------------------- inflate */

/* Canonical Huffman decoder tables: count[len], symbols sorted by (len, sym).
 * For a complete, non-over-subscribed code this decodes exactly what the
 * reference's (length, code) linear scan decodes (dec:86-119 assigns the RFC
 * 1951 canonical codes; dec:671-691 matches them MSB-first). */
typedef struct {
	uint16_t count[16];
	uint16_t symbol[288];
} huff_t;

typedef struct {
	const uint8_t *in;
	uint32_t in_len;
	uint64_t bitpos; /* bits consumed so far; the reference has loaded ceil(bitpos/8) bytes (dec:44-61) */
} bits_t;

/* dec:44-61 get_bit: -1 when no input byte is left */
static int getbit(bits_t *b) {
	if ((b->bitpos >> 3) >= b->in_len) {
		return -1;
	}
	int bit = (b->in[b->bitpos >> 3] >> (b->bitpos & 7)) & 1;
	b->bitpos++;
	return bit;
}

/* dec:64-83 get_bits: LSB-first, -1 on exhaustion */
static int getbits(bits_t *b, int n) {
	int v = 0;
	for (int i = 0; i < n; i++) {
		int bit = getbit (b);
		if (bit < 0) {
			return -1;
		}
		v |= bit << i;
	}
	return v;
}

/* Build canonical tables.  Returns 0 complete, >0 incomplete (left-over code
 * space), <0 over-subscribed.  dec:86-119 performs no such check (SURVEY
 * Appendix B.5): the oracle is strict, as zlib is, and so is the CUDA path. */
static int huff_build(huff_t *h, const uint8_t *lengths, int n) {
	uint16_t offs[16];
	memset (h->count, 0, sizeof (h->count));
	for (int i = 0; i < n; i++) {
		h->count[lengths[i]]++;
	}
	int left = 1;
	for (int len = 1; len <= 15; len++) {
		left <<= 1;
		left -= h->count[len];
		if (left < 0) {
			return left;
		}
	}
	offs[1] = 0;
	for (int len = 1; len < 15; len++) {
		offs[len + 1] = offs[len] + h->count[len];
	}
	for (int i = 0; i < n; i++) {
		if (lengths[i]) {
			h->symbol[offs[lengths[i]]++] = (uint16_t)i;
		}
	}
	return left;
}

/* dec:671-691 / dec:743-764: accumulate MSB-first, return the symbol whose
 * canonical (length, code) matches; -1 no input, -2 no code within 15 bits. */
static int huff_decode(bits_t *b, const huff_t *h) {
	int code = 0, first = 0, index = 0;
	for (int len = 1; len <= 15; len++) {
		int bit = getbit (b);
		if (bit < 0) {
			return -1;
		}
		code |= bit;
		int count = h->count[len];
		if (code - count < first) {
			return h->symbol[index + (code - first)];
		}
		index += count;
		first += count;
		first <<= 1;
		code <<= 1;
	}
	return -2;
}

static const uint16_t len_base[29] = { /* dec:720-722 */
	3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t len_extra[29] = { /* dec:723-725 */
	0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t dst_base[30] = { /* dec:766-768 */
	1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
	6145, 8193, 12289, 16385, 24577 };
static const uint8_t dst_extra[30] = { /* dec:769-771 */
	0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

/* dec:122-266 read_dynamic_huffman.  Strict where the reference is lax:
 * HLIT>286 / HDIST>30 (dec:165-168 overflow), over-subscribed or incomplete
 * sets, missing end-of-block code → data error. */
static int read_dynamic(bits_t *b, huff_t *lit, huff_t *dist) {
	static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 }; /* dec:146-148 */
	int hlit = getbits (b, 5), hdist = getbits (b, 5), hclen = getbits (b, 4);
	if (hlit < 0 || hdist < 0 || hclen < 0) {
		return OTZO_DATA_ERROR;
	}
	hlit += 257;
	hdist += 1;
	hclen += 4;
	if (hlit > 286 || hdist > 30) {
		return OTZO_DATA_ERROR;
	}
	uint8_t cl[19] = { 0 };
	for (int i = 0; i < hclen; i++) {
		int v = getbits (b, 3);
		if (v < 0) {
			return OTZO_DATA_ERROR;
		}
		cl[order[i]] = (uint8_t)v;
	}
	huff_t clh;
	if (huff_build (&clh, cl, 19) != 0) {
		return OTZO_DATA_ERROR;
	}
	uint8_t lens[286 + 30] = { 0 };
	int idx = 0;
	while (idx < hlit + hdist) {
		int sym = huff_decode (b, &clh);
		if (sym < 0) {
			return OTZO_DATA_ERROR;
		}
		if (sym < 16) {
			lens[idx++] = (uint8_t)sym;
			continue;
		}
		int val = 0, rep;
		if (sym == 16) { /* dec:209-219 */
			if (idx == 0) {
				return OTZO_DATA_ERROR;
			}
			val = lens[idx - 1];
			rep = getbits (b, 2);
			if (rep < 0) {
				return OTZO_DATA_ERROR;
			}
			rep += 3;
		} else if (sym == 17) { /* dec:221-228 */
			rep = getbits (b, 3);
			if (rep < 0) {
				return OTZO_DATA_ERROR;
			}
			rep += 3;
		} else { /* dec:230-237 */
			rep = getbits (b, 7);
			if (rep < 0) {
				return OTZO_DATA_ERROR;
			}
			rep += 11;
		}
		if (idx + rep > hlit + hdist) { /* dec:244 */
			return OTZO_DATA_ERROR;
		}
		while (rep--) {
			lens[idx++] = (uint8_t)val;
		}
	}
	if (lens[256] == 0) {
		return OTZO_DATA_ERROR; /* no end-of-block code */
	}
	/* incomplete sets are accepted only in zlib's one case: a single 1-bit code */
	int r = huff_build (lit, lens, hlit);
	if (r < 0 || (r > 0 && !(hlit - lit->count[0] == 1 && lit->count[1] == 1))) {
		return OTZO_DATA_ERROR;
	}
	r = huff_build (dist, lens + hlit, hdist);
	if (r < 0 || (r > 0 && hdist - dist->count[0] != 0 && !(hdist - dist->count[0] == 1 && dist->count[1] == 1))) {
		return OTZO_DATA_ERROR;
	}
	return OTZO_OK;
}

/* dec:322-349 */
static void fixed_tables(huff_t *lit, huff_t *dist) {
	uint8_t l[288];
	int i = 0;
	for (; i < 144; i++) { l[i] = 8; }
	for (; i < 256; i++) { l[i] = 9; }
	for (; i < 280; i++) { l[i] = 7; }
	for (; i < 288; i++) { l[i] = 8; }
	huff_build (lit, l, 288);
	for (i = 0; i < 30; i++) { l[i] = 5; }
	huff_build (dist, l, 30);
}

/* The reference's main loop (dec:610-817) advances one "step" per iteration:
 * state 0 = 3 header bits, state 1 = block set-up (stored copy / fixed tables /
 * dynamic header), state 2 = one literal, end-of-block or length+distance pair.
 * After every step that does not finish the stream it returns Z_BUF_ERROR if
 * avail_in == 0 (dec:811-816) — and avail_in drops to 0 as soon as the last
 * input byte has been LOADED (dec:51-54), i.e. once ceil(bitpos/8) == in_len.
 * ref_check() is that test; the first time it fires *ref_ret is latched and the
 * oracle keeps decoding to deliver the RFC 1951 verdict as well. */
#define REF_CHECK() \
	do { \
		if (!ref_done && ((b.bitpos + 7) >> 3) >= in_len) { \
			*ref_ret = OTZO_BUF_ERROR; \
			ref_done = 1; \
		} \
	} while (0)
#define FAIL(code) \
	do { \
		if (!ref_done) { \
			*ref_ret = (code); \
		} \
		*rfc_ret = (code) == OTZO_OK ? OTZO_BUF_ERROR : (code); \
		*total_out = op; \
		return; \
	} while (0)

void otzo_inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap,
	uint32_t *total_out, int *ref_ret, int *rfc_ret) {
	bits_t b = { in, in_len, 0 };
	huff_t lit, dist;
	uint32_t op = 0;
	int ref_done = 0;
	*ref_ret = OTZO_BUF_ERROR;
	*rfc_ret = OTZO_BUF_ERROR;
	*total_out = 0;
	/* dec:610: the loop body never runs with avail_in == 0 */
	if (in_len == 0) {
		*ref_ret = out_cap == 0 ? OTZO_OK : OTZO_BUF_ERROR; /* dec:825-830 */
		return;
	}
	for (;;) {
		/* state 0, dec:613-627 */
		int final = getbit (&b);
		int btype = getbits (&b, 2);
		if (final < 0 || btype < 0) {
			FAIL (OTZO_DATA_ERROR);
		}
		REF_CHECK ();
		/* state 1, dec:629-660 */
		if (btype == 0) {
			/* dec:269-319: drop the partial byte, LEN/NLEN, straight copy */
			uint64_t pos = (b.bitpos + 7) >> 3;
			if (in_len - pos < 4) {
				FAIL (OTZO_DATA_ERROR);
			}
			uint32_t len = in[pos] | (in[pos + 1] << 8), nlen = in[pos + 2] | (in[pos + 3] << 8);
			pos += 4;
			if (len != ((~nlen) & 0xFFFFu) || in_len - pos < len) {
				FAIL (OTZO_DATA_ERROR);
			}
			if (out_cap - op < len) {
				FAIL (OTZO_BUF_ERROR);
			}
			memcpy (out + op, in + pos, len);
			op += len;
			b.bitpos = (pos + len) << 3;
			if (final) {
				break; /* dec:806-808 */
			}
			REF_CHECK ();
			continue;
		} else if (btype == 1) {
			fixed_tables (&lit, &dist);
		} else if (btype == 2) {
			if (read_dynamic (&b, &lit, &dist) != OTZO_OK) {
				FAIL (OTZO_DATA_ERROR);
			}
		} else {
			FAIL (OTZO_DATA_ERROR); /* dec:657-658 */
		}
		REF_CHECK ();
		/* state 2, dec:662-799: one symbol per step */
		for (;;) {
			int sym = huff_decode (&b, &lit);
			if (sym < 0) {
				FAIL (OTZO_DATA_ERROR);
			}
			if (sym < 256) {
				if (op >= out_cap) {
					FAIL (OTZO_OK); /* dec:700-703: pending literal, Z_OK, never STREAM_END */
				}
				out[op++] = (uint8_t)sym;
			} else if (sym == 256) {
				break;
			} else if (sym <= 285) {
				int li = sym - 257;
				int length = len_base[li];
				if (len_extra[li]) {
					int x = getbits (&b, len_extra[li]);
					if (x < 0) {
						FAIL (OTZO_DATA_ERROR);
					}
					length += x;
				}
				int ds = huff_decode (&b, &dist);
				if (ds < 0 || ds > 29) {
					FAIL (OTZO_DATA_ERROR);
				}
				uint32_t distance = dst_base[ds];
				if (dst_extra[ds]) {
					int x = getbits (&b, dst_extra[ds]);
					if (x < 0) {
						FAIL (OTZO_DATA_ERROR);
					}
					distance += (uint32_t)x;
				}
				/* dec:785 only rejects distance > 32768 and would read its
				 * uninitialised window (dec:493); a reach before the start of
				 * the output is invalid per RFC 1951 → strict error. */
				if (distance > op) {
					FAIL (OTZO_DATA_ERROR);
				}
				if ((uint32_t)length > out_cap - op) {
					FAIL (OTZO_OK); /* dec:535-541, 791-793: pending copy, Z_OK */
				}
				for (int i = 0; i < length; i++, op++) { /* dec:521-533 */
					out[op] = out[op - distance];
				}
			} else {
				FAIL (OTZO_DATA_ERROR); /* dec:794-797 */
			}
			REF_CHECK ();
		}
		if (final) {
			break; /* dec:714-716 */
		}
		REF_CHECK ();
	}
	if (!ref_done) {
		*ref_ret = OTZO_STREAM_END;
	}
	*rfc_ret = OTZO_STREAM_END;
	*total_out = op;
}

/* ------------------------------------------------ method 93 (reference container) */

/* zstd:479-705 driven one-shot as otezip.c:542-555.  See SURVEY.md Appendix C. */
int otzo_zstdref_decode(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, uint32_t *total_out) {
	uint32_t ip = 0, op = 0;
	*total_out = 0;
	if (in_len < 5) {
		return OTZO_BUF_ERROR; /* zstd:490-492 */
	}
	uint32_t magic = in[0] | (in[1] << 8) | (in[2] << 16) | ((uint32_t)in[3] << 24);
	if (magic != 0xFD2FB528u) {
		return OTZO_DATA_ERROR; /* zstd:495-498 */
	}
	ip = 5; /* zstd:501-507: one descriptor byte, value unused */
	while (in_len - ip > 0) { /* zstd:511 */
		if (in_len - ip < 3) {
			return OTZO_BUF_ERROR; /* zstd:698-700 */
		}
		uint8_t h = in[ip];
		int last = h & 1, type = (h >> 1) & 3; /* zstd:554-556 */
		uint32_t bsz = in[ip + 1] | (in[ip + 2] << 8); /* zstd:559-560 */
		ip += 3;
		if (type != 0 && type != 2) {
			return OTZO_DATA_ERROR; /* zstd:689-692 */
		}
		if (in_len - ip < bsz) {
			return OTZO_BUF_ERROR; /* zstd:570-576, 635-641 */
		}
		if (type == 2 && bsz == 0) {
			return OTZO_DATA_ERROR; /* zstd:189-191 → :647-649 */
		}
		if (out_cap - op < bsz) {
			/* zstd:608-632 / :676-683: spills what fits, then returns Z_OK at :546-548 */
			*total_out = out_cap;
			return OTZO_OK;
		}
		memcpy (out + op, in + ip, bsz);
		op += bsz;
		ip += bsz;
		*total_out = op;
		if (last) {
			return OTZO_STREAM_END; /* zstd:695-697 */
		}
	}
	return OTZO_OK; /* zstd:704 */
}

/* ------------------------------------------------------- container layer */

static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t rd32(const uint8_t *p) {
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

#define MAX_PAYLOAD (2ULL * 1024 * 1024 * 1024) /* otezip.c:102 */

int otzo_load_central(const uint8_t *img, uint64_t img_len, otzo_entry **entries, uint32_t *n_entries) {
	*entries = NULL;
	*n_entries = 0;
	/* otezip.c:199-272 */
	if (img_len < 22) {
		return -2;
	}
	uint64_t search = img_len < 0x10000 + 22 ? img_len : 0x10000 + 22;
	const uint8_t *buf = img + (img_len - search);
	uint32_t cd_size = 0, cd_ofs = 0;
	uint16_t n = 0;
	int found = 0;
	for (uint64_t i = search - 22 + 1; i-- > 0;) {
		if (rd32 (buf + i) != 0x06054b50u) {
			continue;
		}
		uint16_t ents = rd16 (buf + i + 10);
		uint32_t sz = rd32 (buf + i + 12), ofs = rd32 (buf + i + 16);
		if (ofs > (uint32_t)img_len || (uint64_t)ofs + sz > img_len) {
			continue; /* otezip.c:235-239 */
		}
		if (ents > 0 && sz >= 4) { /* otezip.c:243-261 */
			if ((uint64_t)ofs + 4 > img_len || rd32 (img + ofs) != 0x02014b50u) {
				continue;
			}
		}
		n = ents;
		cd_size = sz;
		cd_ofs = ofs;
		found = 1;
		break;
	}
	if (!found) {
		return -2;
	}
	/* otezip.c:275-396 */
	if ((uint64_t)cd_ofs + cd_size > img_len) {
		return -2;
	}
	if (n == 0) {
		return cd_size != 0 ? -2 : 0; /* otezip.c:305-312 */
	}
	if ((uint64_t)n * 46 > cd_size) {
		return -2; /* otezip.c:332-335 */
	}
	otzo_entry *e = (otzo_entry *)calloc (n, sizeof (*e));
	if (!e) {
		return -1;
	}
	const uint8_t *cd = img + cd_ofs;
	uint64_t off = 0;
	for (uint32_t i = 0; i < n; i++) {
		if (off + 46 > cd_size || rd32 (cd + off) != 0x02014b50u) {
			free (e);
			return -2; /* otezip.c:349-352 */
		}
		const uint8_t *h = cd + off;
		uint64_t fl = rd16 (h + 28), xl = rd16 (h + 30), cl = rd16 (h + 32);
		uint64_t esz = 46 + fl + xl + cl;
		if (esz > cd_size - off) {
			free (e);
			return -2; /* otezip.c:365-368 */
		}
		e[i].method = rd16 (h + 10);
		e[i].file_time = rd16 (h + 12);
		e[i].file_date = rd16 (h + 14);
		e[i].crc32 = rd32 (h + 16);
		e[i].comp_size = rd32 (h + 20);
		e[i].uncomp_size = rd32 (h + 24);
		e[i].external_attr = rd32 (h + 38);
		e[i].local_hdr_ofs = rd32 (h + 42);
		e[i].name_ofs = (uint32_t)(cd_ofs + off + 46);
		e[i].name_len = (uint16_t)fl;
		if (e[i].comp_size > MAX_PAYLOAD || e[i].uncomp_size > MAX_PAYLOAD) {
			free (e);
			return -2; /* otezip.c:380-383 */
		}
		off += esz;
	}
	*entries = e;
	*n_entries = n;
	return 0;
}

int otzo_extract_entry(const uint8_t *img, uint64_t img_len, const otzo_entry *e, const otzo_opts *o,
	uint8_t *out, uint32_t *crc_out, int *crc_mismatch) {
	*crc_out = 0;
	*crc_mismatch = 0;
	/* otezip.c:403-423 */
	if ((uint64_t)e->local_hdr_ofs > img_len || img_len - e->local_hdr_ofs < 30) {
		return -1;
	}
	const uint8_t *lfh = img + e->local_hdr_ofs;
	if (rd32 (lfh) != 0x04034b50u) {
		return -1;
	}
	/* otezip.c:429-446 */
	uint64_t data_ofs = (uint64_t)e->local_hdr_ofs + 30 + rd16 (lfh + 26) + rd16 (lfh + 28);
	if (data_ofs > img_len) {
		return -1;
	}
	if (e->comp_size > MAX_PAYLOAD || e->uncomp_size > MAX_PAYLOAD) {
		return -1;
	}
	if (data_ofs + e->comp_size > img_len) {
		return -1;
	}
	/* otezip.c:454-462 */
	if (!o->ignore_zipbomb && e->comp_size > 0) {
		uint64_t allowed = (uint64_t)e->comp_size * o->max_ratio + o->max_slack;
		if ((uint64_t)e->uncomp_size > allowed) {
			return -1;
		}
	}
	const uint8_t *c = img + data_ofs;
	if (e->method == 0) { /* otezip.c:481-487 */
		if (e->comp_size != e->uncomp_size) {
			return -1;
		}
		memcpy (out, c, e->uncomp_size);
	} else if (e->method == 8) { /* otezip.c:490-532 */
		uint32_t tot;
		int ref_ret, rfc_ret;
		memset (out, 0, e->uncomp_size); /* otezip.c:500 */
		otzo_inflate_raw (c, e->comp_size, out, e->uncomp_size, &tot, &ref_ret, &rfc_ret);
		if (ref_ret != OTZO_STREAM_END) {
			return -1; /* otezip.c:525-529; total_out is not compared (Appendix B.7) */
		}
	} else if (e->method == 93) { /* otezip.c:535-561 */
		uint32_t tot;
		memset (out, 0, e->uncomp_size);
		int r = otzo_zstdref_decode (c, e->comp_size, out, e->uncomp_size, &tot);
		if (r != OTZO_STREAM_END || tot != e->uncomp_size) {
			return -1;
		}
	} else {
		return -1; /* methods outside the hot path: oracle scope ends here */
	}
	/* otezip.c:667-679 */
	*crc_out = otzo_crc32 (0, out, e->uncomp_size);
	if (*crc_out != e->crc32) {
		*crc_mismatch = 1;
		if (o->verify_crc) {
			return -1;
		}
	}
	return 0;
}

void otzo_extract_range(const uint8_t *img, uint64_t img_len, const otzo_entry *ents, uint32_t first, uint32_t last,
	const otzo_opts *o, uint8_t *out, const uint64_t *out_ofs, uint32_t *crc, int32_t *status) {
	for (uint32_t i = first; i < last; i++) {
		int mm;
		status[i] = otzo_extract_entry (img, img_len, &ents[i], o, out + out_ofs[i], &crc[i], &mm);
		if (mm && status[i] == 0) {
			status[i] = 0x100; /* accepted with the reference's CRC warning (otezip.c:676) */
		}
	}
}
