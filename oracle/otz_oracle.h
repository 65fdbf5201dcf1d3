/* otz_oracle.h — CPU oracle for the otezip archive-codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker.  The product
 * (otezip_b200/csrc) never links, calls or falls back to this code.
 *
 * Every function is a plain-C restatement of the reference algorithm
 * (trufae/otezip, paths relative to /root/reference) and cites the lines it
 * follows.  Parity is PINNED: tests/test_oracle_vs_ref.py checks each function
 * against the compiled, unmodified reference (oracle/_ref/libotezip_ref.so,
 * built by oracle/Makefile from the sources where they lie) and against the
 * known-answer vectors the reference's own tests hold.
 */
#ifndef OTZ_ORACLE_H
#define OTZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* zlib-style return codes, src/lib/deflate.inc.c:33-42 */
#define OTZO_OK 0
#define OTZO_STREAM_END 1
#define OTZO_DATA_ERROR (-3)
#define OTZO_BUF_ERROR (-5)

/* src/lib/crc32.inc.c:40-47 — IEEE reflected CRC-32, chainable seed. */
uint32_t otzo_crc32(uint32_t crc, const void *buf, size_t len);

/* src/lib/deflate-dec.inc.c:547-831 driven as otezip.c:503-529 drives it:
 * raw stream, one call, Z_FINISH, output buffer exactly out_cap bytes.
 *   *ref_ret  = what the reference's inflate() returns (OTZO_STREAM_END on
 *               success; anything else makes zip_fopen_index return NULL);
 *   *rfc_ret  = what a correct RFC 1951 decoder returns for the same stream
 *               (differs from *ref_ret only by the reference's EOB rule,
 *               SURVEY.md F1 / dec:811-816);
 *   *total_out= bytes produced by the full (RFC) decode.
 * Streams that are invalid per RFC 1951 give OTZO_DATA_ERROR in both (the
 * reference has undefined behaviour on some of them, dec:785/:766-774). */
void otzo_inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap,
	uint32_t *total_out, int *ref_ret, int *rfc_ret);

/* src/lib/zstd.inc.c:479-705 — the reference's method-93 raw-block container
 * (NOT RFC 8878).  Returns the reference's return code; *total_out as the
 * reference's strm.total_out. */
int otzo_zstdref_decode(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap,
	uint32_t *total_out);

/* One central-directory entry, the fields of struct otezip_entry
 * (src/include/otezip/zip.h:78-88) that the read path consumes. */
typedef struct {
	uint32_t local_hdr_ofs, comp_size, uncomp_size, crc32;
	uint16_t method, file_time, file_date;
	uint32_t external_attr;
	uint32_t name_ofs; /* offset of the name inside the archive image */
	uint16_t name_len;
} otzo_entry;

/* src/lib/otezip.c:199-272 + :275-396 on an in-memory archive image.
 * Returns 0, -1 (OTEZIP_ERR_READ) or -2 (OTEZIP_ERR_INCONS).  *entries is
 * malloc'd (caller frees), NULL when the archive is empty. */
int otzo_load_central(const uint8_t *img, uint64_t img_len, otzo_entry **entries, uint32_t *n_entries);

typedef struct {
	int verify_crc;          /* otezip_verify_crc,          otezip.c:157 */
	int ignore_zipbomb;      /* otezip_ignore_zipbomb,      otezip.c:166 */
	uint64_t max_ratio;      /* otezip_max_expansion_ratio, otezip.c:164 */
	uint64_t max_slack;      /* otezip_max_expansion_slack, otezip.c:165 */
} otzo_opts;

/* src/lib/otezip.c:399-684 — extract one entry from the image into `out`
 * (uncomp_size bytes, caller-owned).  Returns 0 when the reference would hand
 * the buffer to the caller, -1 when zip_fopen_index would return NULL.
 * *crc_out = CRC-32 of the produced buffer (computed whenever decode succeeded),
 * *crc_mismatch = 1 when it differs from the directory value. */
int otzo_extract_entry(const uint8_t *img, uint64_t img_len, const otzo_entry *e, const otzo_opts *o,
	uint8_t *out, uint32_t *crc_out, int *crc_mismatch);

/* Batch driver used for CPU-side timing and bulk parity: entries [first,last)
 * extracted into out + out_ofs[i]; status[i] = 0 / -1 as above. */
void otzo_extract_range(const uint8_t *img, uint64_t img_len, const otzo_entry *ents, uint32_t first, uint32_t last,
	const otzo_opts *o, uint8_t *out, const uint64_t *out_ofs, uint32_t *crc, int32_t *status);

#ifdef __cplusplus
}
#endif
#endif
