// k_crc32.cuh — CRC-32 (IEEE, reflected) over device-resident byte ranges.
//
// Replaces otezip_crc32 (/root/reference/src/lib/crc32.inc.c:40-47), which walks
// one byte per iteration through a 256-entry table.  Here one warp owns one
// chunk; the 32 lanes read it as coalesced 16-byte vectors (lane j takes vector
// j of every 512-byte row) and each lane runs a slice-by-16 update whose tables
// already include the 496 zero bytes that separate a lane's consecutive
// vectors.  CRC is linear over GF(2), so the chunk remainder is the XOR of the
// lane remainders, each multiplied by x^(8*distance to the chunk end), and the
// entry CRC is the XOR of the chunk remainders multiplied the same way.
#pragma once
#include "otz_common.cuh"

// Fold lags (tools/crc_fold_search.py): x^(4096*11+25) + x^(4096*9-31) + x^(4096*5+27) + x^(4096*3-2) + 1 is,
// after reversal, a multiple of the CRC-32 polynomial, i.e. u[t] = s[t] ^ u[t-L0] ^ u[t-L1] ^ u[t-L2] ^ u[t-L3]
// with L = 4096*m + r eliminates stream bits without changing M(x) mod P.
#define FOLD_NLAG 4
#define FOLD_M0 3
#define FOLD_R0 (-2)
#define FOLD_M1 5
#define FOLD_R1 27
#define FOLD_M2 9
#define FOLD_R2 (-31)
#define FOLD_M3 11
#define FOLD_R3 25
#define FOLD_MMAX 11
#define FOLD_H (FOLD_MMAX + 2)      // rows of history kept in registers (ring, statically indexed)
#define FOLD_K (FOLD_MMAX + 1)      // zero rows appended so that every data row is eliminated
#define OTZ_CRC_CHUNK (FOLD_H * 512u * 40u)   // 266,240 bytes: 40 blocks of FOLD_H rows
#define OTZ_CRC_FOLD_MIN 8192u      // shorter ranges use the table path

// Load the 16 KiB of skip tables into shared memory (whole CTA).
__device__ __forceinline__ void crc_tables_to_smem(uint32_t *s_skip, const OtzCrcTables *__restrict__ t) {
	const uint4 *src = reinterpret_cast<const uint4 *>(&t->skip[0][0]);
	uint4 *dst = reinterpret_cast<uint4 *>(s_skip);
	for (int i = threadIdx.x; i < 16 * 256 / 4; i += blockDim.x) {
		dst[i] = src[i];
	}
}

// One slice-by-16 step: fold the lane state into the first word, then 16 lookups.
__device__ __forceinline__ uint32_t crc_step16(uint32_t s, uint4 v, const uint32_t *__restrict__ sk) {
	uint32_t a = v.x ^ s;
	uint32_t r = sk[0 * 256 + (a & 0xFF)] ^ sk[1 * 256 + ((a >> 8) & 0xFF)] ^ sk[2 * 256 + ((a >> 16) & 0xFF)] ^
		sk[3 * 256 + (a >> 24)];
	r ^= sk[4 * 256 + (v.y & 0xFF)] ^ sk[5 * 256 + ((v.y >> 8) & 0xFF)] ^ sk[6 * 256 + ((v.y >> 16) & 0xFF)] ^
		sk[7 * 256 + (v.y >> 24)];
	r ^= sk[8 * 256 + (v.z & 0xFF)] ^ sk[9 * 256 + ((v.z >> 8) & 0xFF)] ^ sk[10 * 256 + ((v.z >> 16) & 0xFF)] ^
		sk[11 * 256 + (v.z >> 24)];
	r ^= sk[12 * 256 + (v.w & 0xFF)] ^ sk[13 * 256 + ((v.w >> 8) & 0xFF)] ^ sk[14 * 256 + ((v.w >> 16) & 0xFF)] ^
		sk[15 * 256 + (v.w >> 24)];
	return r;
}

// Zero the bytes of a 16-byte vector outside [lo, hi) (byte indices within the vector).
__device__ __forceinline__ uint4 mask_vec(uint4 v, int lo, int hi) {
	uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
	for (int k = 0; k < 4; k++) {
		int a = lo - 4 * k, b = hi - 4 * k;  // keep bytes [a, b) of word k
		a = a < 0 ? 0 : (a > 4 ? 4 : a);
		b = b < 0 ? 0 : (b > 4 ? 4 : b);
		uint32_t m = (b > a) ? ((b - a == 4) ? 0xFFFFFFFFu : (((1u << (8 * (b - a))) - 1u) << (8 * a))) : 0u;
		w[k] &= m;
	}
	return make_uint4(w[0], w[1], w[2], w[3]);
}

// Pure remainder R(M) = M(x) * x^32 mod P of the n bytes at p (any alignment), computed by
// one full warp.  Valid in every lane.  s_skip: the 16 skip tables in shared memory.
__device__ __forceinline__ uint32_t crc_raw_warp(const uint8_t *__restrict__ p, uint64_t n, const uint32_t *__restrict__ s_skip,
	const OtzCrcTables *__restrict__ tabs) {
	const int lane = threadIdx.x & 31;
	if (n == 0) {
		return 0;
	}
	const uint64_t pa = reinterpret_cast<uint64_t>(p);
	const uint64_t A = pa & ~15ull;
	const uint32_t headpad = (uint32_t)(pa - A);
	const uint64_t L = headpad + n;
	const uint64_t V = (L + 15) >> 4;  // vectors in the aligned span
	const uint32_t tail_hi = (uint32_t)(L - ((V - 1) << 4));  // valid bytes of the last vector (1..16)
	uint32_t s = 0;
	uint64_t cnt = 0;
	const uint4 *base = reinterpret_cast<const uint4 *>(A);
	// main loop, 4 rows per iteration so four 16-byte loads are in flight per lane
	uint64_t v = lane;
	for (; v + 97 < V; v += 128) {
		uint4 q0 = ld_stream16(base + v), q1 = ld_stream16(base + v + 32), q2 = ld_stream16(base + v + 64),
		      q3 = ld_stream16(base + v + 96);
		if (v == 0) {
			q0 = mask_vec(q0, headpad, 16);
		}
		s = crc_step16(s, q0, s_skip);
		s = crc_step16(s, q1, s_skip);
		s = crc_step16(s, q2, s_skip);
		s = crc_step16(s, q3, s_skip);
		cnt += 4;
	}
	for (; v < V; v += 32) {
		uint4 q = ld_stream16(base + v);
		int lo = (v == 0) ? (int)headpad : 0;
		int hi = (v == V - 1) ? (int)tail_hi : 16;
		if (lo != 0 || hi != 16) {
			q = mask_vec(q, lo, hi);
		}
		s = crc_step16(s, q, s_skip);
		cnt++;
	}
	// lane state stands at A + 16*lane + 512*cnt; move it to the true end p+n (always backwards)
	int64_t d = (int64_t)L - (int64_t)(16 * lane) - (int64_t)(cnt << 9);
	uint32_t r = cnt ? crc_mulmod(s, tabs->xp8[d + OTZ_XP8_BIAS]) : 0u;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		r ^= __shfl_xor_sync(0xFFFFFFFFu, r, o);
	}
	return r;
}

// ---------------------------------------------------------------------------------------------
// Fold path.  The chunk is viewed as rows of 512 bytes (lane j owns the 16-byte vector j of every
// row).  Each new row is XORed with four earlier rows shifted by a few bits (funnel shifts inside the
// lane's four words, one shuffle for the word that crosses a lane boundary), which removes the old
// rows from the polynomial without changing its remainder.  FOLD_K zero rows appended at the end
// collect what is left; only those 6 KiB go through the table-driven update.  ~2 ALU ops per byte,
// no shared-memory traffic in the main loop.
template <int M, int RS, int J>
__device__ __forceinline__ void fold_term(uint4 &v, const uint4 (&hist)[FOLD_H], int lane, bool pad_row) {
	// row J of the current block pulls from the row M rows earlier; in a zero (padding) row J only
	// data rows may be sources: M > J, and the cross-row spill of a right shift needs M - 1 > J.
	if (pad_row && M <= J) {
		if (RS > 0 && M == J) {
			// the aligned source row is a padding row, but the bits that cross the row boundary still
			// come from the last data row (lane 31, word 3) and land in lane 0, word 0
			const uint4 older = hist[(J - M - 1 + 2 * FOLD_H) % FOLD_H];
			const uint32_t prev = __shfl_sync(0xFFFFFFFFu, older.w, 31);
			if (lane == 0) {
				v.x ^= prev >> (32 - (RS > 0 ? RS : 1));
			}
		}
		return;
	}
	const uint4 src = hist[(J - M + 2 * FOLD_H) % FOLD_H];
	if (RS == 0) {
		v.x ^= src.x; v.y ^= src.y; v.z ^= src.z; v.w ^= src.w;
	} else if (RS > 0) {
		const uint4 older = hist[(J - M - 1 + 2 * FOLD_H) % FOLD_H];
		const uint32_t give = lane == 31 ? older.w : src.w;
		const uint32_t prev = __shfl_sync(0xFFFFFFFFu, give, (lane + 31) & 31);
		v.x ^= __funnelshift_l(prev, src.x, RS);
		v.y ^= __funnelshift_l(src.x, src.y, RS);
		v.z ^= __funnelshift_l(src.y, src.z, RS);
		v.w ^= __funnelshift_l(src.z, src.w, RS);
	} else {
		const uint4 newer = hist[(J - M + 1 + 2 * FOLD_H) % FOLD_H];
		const bool newer_is_pad = pad_row && (M - 1 <= J);
		const uint32_t give = lane == 0 ? (newer_is_pad ? 0u : newer.x) : src.x;
		const uint32_t next = __shfl_sync(0xFFFFFFFFu, give, (lane + 1) & 31);
		v.x ^= __funnelshift_r(src.x, src.y, -RS);
		v.y ^= __funnelshift_r(src.y, src.z, -RS);
		v.z ^= __funnelshift_r(src.z, src.w, -RS);
		v.w ^= __funnelshift_r(src.w, next, -RS);
	}
}

template <int J>
__device__ __forceinline__ void fold_row(uint4 v, uint4 (&hist)[FOLD_H], int lane, bool pad_row) {
	fold_term<FOLD_M0, FOLD_R0, J>(v, hist, lane, pad_row);
	fold_term<FOLD_M1, FOLD_R1, J>(v, hist, lane, pad_row);
	fold_term<FOLD_M2, FOLD_R2, J>(v, hist, lane, pad_row);
	fold_term<FOLD_M3, FOLD_R3, J>(v, hist, lane, pad_row);
	hist[J] = v;
}

template <int J, bool EDGE>
__device__ __forceinline__ void fold_rows_from(const uint4 (&data)[FOLD_H], uint4 (&hist)[FOLD_H], int lane, bool pad_row) {
	if constexpr (J < FOLD_H) {
		if (!(pad_row && J >= FOLD_K)) {
			fold_row<J>(data[J], hist, lane, pad_row);
		}
		fold_rows_from<J + 1, EDGE>(data, hist, lane, pad_row);
	}
}

// Pure remainder of the n bytes at p (n >= OTZ_CRC_FOLD_MIN), whole warp, valid in every lane.
__device__ __noinline__ uint32_t crc_raw_warp_fold(const uint8_t *__restrict__ p, uint64_t n, const uint32_t *__restrict__ s_skip,
	const OtzCrcTables *__restrict__ tabs) {
	const int lane = threadIdx.x & 31;
	const uint64_t pa = reinterpret_cast<uint64_t>(p);
	const uint64_t A = pa & ~15ull;
	const uint32_t headpad = (uint32_t)(pa - A);
	const uint64_t L = headpad + n;
	const int64_t V = (int64_t)((L + 15) >> 4);                 // vectors in the aligned span
	const uint32_t tail_hi = (uint32_t)(L - ((uint64_t)(V - 1) << 4));
	const int64_t rows = (V + 31) >> 5;
	const int64_t blocks = (rows + FOLD_H - 1) / FOLD_H;
	const int64_t lead = blocks * FOLD_H - rows;               // virtual zero rows in front (leading zeros are free)
	const uint4 *base = reinterpret_cast<const uint4 *>(A);
	uint4 hist[FOLD_H];
#pragma unroll
	for (int j = 0; j < FOLD_H; j++) {
		hist[j] = make_uint4(0, 0, 0, 0);
	}
	uint4 data[FOLD_H];
	for (int64_t b = 0; b < blocks; b++) {
		const int64_t v0 = (b * FOLD_H - lead) * 32 + lane;      // vector index of row 0 of this block for this lane
		if (b > 0 && b + 1 < blocks) {
#pragma unroll
			for (int j = 0; j < FOLD_H; j++) {
				data[j] = ld_stream16(base + v0 + 32 * j);
			}
		} else {
#pragma unroll
			for (int j = 0; j < FOLD_H; j++) {
				const int64_t v = v0 + 32 * j;
				uint4 q = make_uint4(0, 0, 0, 0);
				if (v >= 0 && v < V) {
					q = ld_stream16(base + v);
					const int lo = v == 0 ? (int)headpad : 0, hi = v == V - 1 ? (int)tail_hi : 16;
					if (lo != 0 || hi != 16) {
						q = mask_vec(q, lo, hi);
					}
				}
				data[j] = q;
			}
		}
		fold_rows_from<0, false>(data, hist, lane, false);
	}
	// FOLD_K zero rows: afterwards hist[0..FOLD_K) is everything that is left of the chunk
#pragma unroll
	for (int j = 0; j < FOLD_H; j++) {
		data[j] = make_uint4(0, 0, 0, 0);
	}
	fold_rows_from<0, false>(data, hist, lane, true);
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < FOLD_K; j++) {
		s = crc_step16(s, hist[j], s_skip);
	}
	// the lane state stands 16*lane bytes past the end of the residue rows; the residue itself is the chunk
	// followed by (row padding + FOLD_K rows) zero bytes
	uint32_t r = crc_mulmod(s, tabs->xp8[OTZ_XP8_BIAS - 16 * lane]);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		r ^= __shfl_xor_sync(0xFFFFFFFFu, r, o);
	}
	const uint32_t trail = (uint32_t)((uint64_t)rows * 512u - L);   // zero bytes up to the end of the last data row
	r = crc_mulmod(r, tabs->xp8[OTZ_XP8_BIAS - (int)trail]);
	return crc_mulmod(r, tabs->x_inv_fold);
}

struct OtzCrcChunk {
	uint32_t entry;
	uint32_t chunk;
};

// grid: persistent, any size; block: multiple of 32.  One warp per chunk.
__global__ void __launch_bounds__(256, 2) k_crc_chunks(const uint8_t *__restrict__ archive, const uint8_t *__restrict__ out,
	const otz_entry *__restrict__ ents, const OtzEntryState *__restrict__ est, const int32_t *__restrict__ status,
	const OtzCrcChunk *__restrict__ chunks, uint32_t n_chunks, uint32_t *__restrict__ acc, const OtzCrcTables *__restrict__ tabs,
	int verify_only, int only_ref_rejected) {
	__shared__ __align__(16) uint32_t s_skip[16 * 256];
	crc_tables_to_smem(s_skip, tabs);
	__syncthreads();
	const uint32_t warps_per_cta = blockDim.x >> 5;
	const uint32_t total_warps = gridDim.x * warps_per_cta;
	const int lane = threadIdx.x & 31;
	for (uint32_t c = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); c < n_chunks; c += total_warps) {
		const OtzCrcChunk ck = chunks[c];
		const int32_t est_ = status[ck.entry];
		if (OTZ_ST_CODE(est_) != OTZ_ST_OK || (only_ref_rejected && !(est_ & OTZ_STF_REF_EOB))) {
			continue;   // failed; or (method 93) already CRC'd inside k_zstdref
		}
		const otz_entry e = ents[ck.entry];
		const uint64_t off = (uint64_t)ck.chunk * OTZ_CRC_CHUNK;
		const uint64_t len = min((uint64_t)OTZ_CRC_CHUNK, (uint64_t)e.uncomp_size - off);
		const uint8_t *src = (verify_only && e.method == OTZ_M_STORE) ? archive + est[ck.entry].data_ofs + off
		                                                               : out + e.out_ofs + off;
		uint32_t raw = len >= OTZ_CRC_FOLD_MIN ? crc_raw_warp_fold(src, len, s_skip, tabs) : crc_raw_warp(src, len, s_skip, tabs);
		if (lane == 0) {
			uint64_t after = (uint64_t)e.uncomp_size - off - len;
			uint32_t shifted = after ? crc_mulmod(raw, crc_xpow8(after, tabs->x2n)) : raw;
			atomicXor(&acc[ck.entry], shifted);
		}
	}
}

// crc32(M) = R(M) ^ crc32(0^N), crc32(0^N) = ~(x^(8N) * 0xFFFFFFFF); compare with the directory value
// (otezip.c:667-679).  One thread per entry.
__global__ void k_crc_finalize(const otz_entry *__restrict__ ents, uint32_t n, const uint32_t *__restrict__ acc,
	uint32_t *__restrict__ crc_out, int32_t *__restrict__ status, const OtzCrcTables *__restrict__ tabs) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	int32_t st = status[i];
	if (OTZ_ST_CODE(st) != OTZ_ST_OK || (ents[i].flags & OTZ_EF_CHUNK)) {
		crc_out[i] = 0;   // failed, or a chunk row (its parent row carries the CRC)
		return;
	}
	uint32_t N = ents[i].uncomp_size;
	uint32_t crc = acc[i] ^ ~crc_mulmod(crc_xpow8(N, tabs->x2n), 0xFFFFFFFFu);
	crc_out[i] = crc;
	if (crc != ents[i].crc32) {
		status[i] = st | OTZ_STF_CRC_MISMATCH;
	}
}
