// k_zstd_tok.cuh — phase A of the two-phase Zstandard (RFC 8878) path: tokens for k_inflate_lz<W, true>.
//
// Not part of the reference (its method 93 is a raw-block container, /root/reference/src/lib/zstd.inc.c:479-705, and it
// rejects real frames: SURVEY.md F3); parity is pinned against libzstd 1.5.5 (tests/test_gpu_zstd.py).
//
// The token stream of an entry = the literal bytes of all its blocks back to back (dense, in stream order) + 8-byte
// sequence records {literal run | (length - 3) << 9, offset}, long runs / matches split into <= 511 / <= 258 byte
// pieces, repeat offsets resolved.  The format hands out two kinds of parallelism, and the two kernels below take one
// each (round 1 had ONE thread do both for an entry, five such threads per warp, each in a different phase of its
// entry: 4 of 32 lanes busy, every warp instruction paid five times):
//
//   k_zstd_lit   FOUR lanes per entry (the four Huffman streams of a literals section are independent by construction),
//                eight entries per warp, one 4 KiB decoding table per entry in shared memory.  The four lanes walk the
//                frame / block structure together (the same instructions: no cost in lock-step), build the table
//                together and then decode one stream each; raw / RLE literal sections and raw blocks are copied by the
//                four lanes in turns.
//   k_zstd_seq   ONE lane per entry for the serial part: the three interleaved FSE states of a sequences section.  Only
//                the FSE decoding tables are needed here, as 16-bit entries 2.5 KiB per lane (round 1: 9.7 KB per lane
//                for everything = 20 entries per SM; here up to 80), and the code -> base / extra-bits tables are shared
//                by the CTA.  The chain state -> table entry -> bit counts -> next state is a few shared-memory round
//                trips per sequence (with the tables in global memory it was L2 round trips: 3x slower).
//
// Both kernels are written as warp-level ROUNDS so that the lanes stay in lock-step where the instructions are: every
// lane (group) first walks its entry to the next block that has work for this kernel — fetching a new entry from the
// work counter whenever its own is finished —, then all of them build their tables, then all of them run their decoding
// loop.  The walk (zs_walk) is the one piece of code both kernels share, so they agree on the structure and on every
// structural error.  The two kernels run at the same time (the sequence kernel leaves a fifth of an SM's shared memory and
// most of its issue slots free); k_zstd_join releases an entry to the LZ executor when both accepted it, and otherwise gives it
// the status of the error that comes first in the stream.
#pragma once
#include "k_zstd.cuh"

// worst-case token scratch of an entry of n output bytes: literals + 8 bytes per record (a match piece is >= 3 bytes,
// an escape exactly 511 literals)
__host__ __device__ __forceinline__ uint64_t zs_scratch_bytes(uint64_t n) { return ((n + 8 * (n / 3 + n / 511 + 8) + 64 + 15) / 16) * 16; }

struct ZsTok {
	uint8_t *L;        // literal bytes, ascending
	uint2 *seq_end;    // records, descending from here
	uint32_t nl;       // literals so far (both kernels count them)
	uint32_t nseq;     // records written
	uint32_t run;      // literals since the last match
	__device__ __forceinline__ void rec(uint32_t w0, uint32_t off) {
		seq_end[-1 - (int32_t)nseq] = make_uint2(w0, off);
		nseq++;
	}
	__device__ __forceinline__ void match(uint32_t len, uint32_t off) {   // len >= 3
		while (run >= 511u) {
			rec(511u, 0u);
			run -= 511u;
		}
		uint32_t r = run;
		run = 0;
		while (len) {
			const uint32_t piece = len > 258u ? (len - 258u < 3u ? len - 3u : 258u) : len;
			rec(r | ((piece - 3u) << 9), off);
			r = 0;
			len -= piece;
		}
	}
};

#define ZS_PH_LIT 0
#define ZS_PH_SEQ 1

// where an entry stands between two blocks
#define ZSW_IN_FRAME 1u     // inside a frame (behind its header)
#define ZSW_FRAME_END 2u    // the last block of the frame has been handed out
#define ZSW_CKSUM 4u        // the frame ends with a 4-byte checksum
#define ZSW_FRESH 8u        // a frame header was read since the last compressed block: repeat offsets / tables start over
#define ZSW_HAS_FCS 16u
struct ZsWalk {
	const uint8_t *in;
	uint32_t n, cap;
	uint32_t ip;
	uint32_t op;            // output position (ZS_PH_SEQ; the literal kernel does not know the match lengths)
	uint32_t frame_start;
	uint32_t n_frames;
	uint32_t fl;
	uint32_t pos;           // frame headers + blocks read so far: where an error was found
	uint64_t fcs;
	int32_t err;
	__device__ __forceinline__ void init(const uint8_t *in_, uint32_t n_, uint32_t cap_) {
		in = in_;
		n = n_;
		cap = cap_;
		ip = op = frame_start = n_frames = fl = pos = 0;
		fcs = 0;
		err = 0;
	}
};

// a compressed block, its literals section header parsed (RFC 8878 3.1.1.3.1)
struct ZsBlk {
	const uint8_t *bp;
	uint32_t bsz;
	uint32_t ltype, regen, lcomp, lhdr, nstreams;
	uint32_t lsec;   // bytes of the whole literals section
};

// Advance entry W to its next compressed block (0: B is filled) or to its end (1: W.err, or all frames read).  Frame
// headers, skippable frames, raw and RLE blocks are consumed on the way: their tokens are emitted by the kernel whose
// part they are (PH; `sub` of `nsub` lanes share the copies of the literal kernel).
template <int PH>
__device__ int zs_walk(ZsWalk &W, ZsTok &T, ZsBlk &B, uint32_t sub, uint32_t nsub) {
	const uint8_t *const in = W.in;
	const uint32_t n = W.n, cap = W.cap;
#define ZSW_FAIL(code_)  \
	do {                 \
		W.err = (code_); \
		return 1;        \
	} while (0)
	for (;;) {
		if (W.err) {
			return 1;
		}
		if (!(W.fl & ZSW_IN_FRAME)) {
			if (W.ip >= n) {
				return 1;
			}
			W.pos++;
			if (n - W.ip < 4) {
				ZSW_FAIL(OTZ_ST_TRUNCATED);
			}
			const uint32_t magic = ld_le32(in + W.ip);
			if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {   // skippable frame
				if (n - W.ip < 8 || n - W.ip - 8 < ld_le32(in + W.ip + 4)) {
					ZSW_FAIL(OTZ_ST_TRUNCATED);
				}
				W.ip += 8 + ld_le32(in + W.ip + 4);
				continue;
			}
			if (magic != 0xFD2FB528u) {
				ZSW_FAIL(OTZ_ST_DATA);
			}
			uint32_t ip = W.ip + 4;
			if (ip >= n) {
				ZSW_FAIL(OTZ_ST_TRUNCATED);
			}
			// ---- frame header (RFC 8878 3.1.1.1)
			const uint32_t fhd = in[ip++];
			const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_cksum = (fhd >> 2) & 1, did_flag = fhd & 3;
			if (fhd & 0x08) {
				ZSW_FAIL(OTZ_ST_DATA);
			}
			const uint32_t did_len = did_flag == 3 ? 4 : did_flag;
			const uint32_t fcs_len = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8);
			if (n - ip < (single ? 0 : 1) + did_len + fcs_len) {
				ZSW_FAIL(OTZ_ST_TRUNCATED);
			}
			if (!single) {
				ip++;   // window descriptor: the whole output is addressable here
			}
			uint32_t did = 0;
			for (uint32_t i = 0; i < did_len; i++) {
				did |= (uint32_t)in[ip++] << (8 * i);
			}
			if (did) {
				ZSW_FAIL(OTZ_ST_DATA);   // dictionaries are not available to a ZIP entry
			}
			uint64_t fcs = 0;
			for (uint32_t i = 0; i < fcs_len; i++) {
				fcs |= (uint64_t)in[ip++] << (8 * i);
			}
			if (fcs_len == 2) {
				fcs += 256;
			}
			W.fcs = fcs;
			W.ip = ip;
			W.frame_start = W.op;
			W.fl = ZSW_IN_FRAME | ZSW_FRESH | (has_cksum ? ZSW_CKSUM : 0u) | (fcs_len ? ZSW_HAS_FCS : 0u);
			continue;
		}
		if (W.fl & ZSW_FRAME_END) {
			if (W.fl & ZSW_CKSUM) {
				if (n - W.ip < 4) {
					ZSW_FAIL(OTZ_ST_TRUNCATED);
				}
				W.ip += 4;   // XXH64 low word: not verified here, the ZIP CRC-32 covers the entry
			}
			if (PH == ZS_PH_SEQ && (W.fl & ZSW_HAS_FCS) && W.fcs != (uint64_t)(W.op - W.frame_start)) {
				ZSW_FAIL(OTZ_ST_DATA);
			}
			W.n_frames++;
			W.fl &= ZSW_FRESH;   // (a frame without a compressed block leaves the flag to the next one: harmless)
			continue;
		}
		// ---- block header
		W.pos++;
		if (n - W.ip < 3) {
			ZSW_FAIL(OTZ_ST_TRUNCATED);
		}
		const uint32_t bh = in[W.ip] | (in[W.ip + 1] << 8) | (in[W.ip + 2] << 16);
		W.ip += 3;
		const uint32_t last = bh & 1, btype = (bh >> 1) & 3, bsz = bh >> 3;
		if (btype == 3 || bsz > ZS_BLOCK_MAX) {
			ZSW_FAIL(OTZ_ST_DATA);
		}
		if (last) {
			W.fl |= ZSW_FRAME_END;
		}
		if (btype == 0) {   // raw
			if (n - W.ip < bsz) {
				ZSW_FAIL(OTZ_ST_TRUNCATED);
			}
			if ((PH == ZS_PH_SEQ ? cap - W.op : cap - T.nl) < bsz) {   // (every literal is an output byte)
				ZSW_FAIL(OTZ_ST_OVERFLOW);
			}
			if (PH == ZS_PH_LIT) {
				for (uint32_t i = sub; i < bsz; i += nsub) {
					T.L[T.nl + i] = in[W.ip + i];
				}
			}
			T.nl += bsz;
			T.run += bsz;
			W.ip += bsz;
			W.op += bsz;
		} else if (btype == 1) {   // RLE
			if (n - W.ip < 1) {
				ZSW_FAIL(OTZ_ST_TRUNCATED);
			}
			const uint32_t nlit = bsz <= 3u ? bsz : 1u;   // one literal, the rest a run-length match of offset 1
			if ((PH == ZS_PH_SEQ ? cap - W.op : cap - T.nl) < (PH == ZS_PH_SEQ ? bsz : nlit)) {
				ZSW_FAIL(OTZ_ST_OVERFLOW);
			}
			const uint8_t b = in[W.ip++];
			if (PH == ZS_PH_LIT && sub == 0) {
				for (uint32_t i = 0; i < nlit; i++) {
					T.L[T.nl + i] = b;
				}
			}
			T.nl += nlit;
			T.run += nlit;
			if (PH == ZS_PH_SEQ && bsz > 3u) {
				T.match(bsz - 1u, 1u);
			}
			W.op += bsz;
		} else {   // compressed
			if (n - W.ip < bsz || bsz < 2) {
				ZSW_FAIL(n - W.ip < bsz ? OTZ_ST_TRUNCATED : OTZ_ST_DATA);
			}
			const uint8_t *const bp = in + W.ip;
			W.ip += bsz;
			// ---- literals section header
			const uint32_t b0 = bp[0];
			const uint32_t ltype = b0 & 3, sf = (b0 >> 2) & 3;
			uint32_t regen, lcomp = 0, lhdr, nstreams = 1;
			if (ltype < 2) {
				if ((sf & 1) == 0) {
					regen = b0 >> 3;
					lhdr = 1;
				} else if (sf == 1) {
					regen = (b0 >> 4) | (bp[1] << 4);
					lhdr = 2;
				} else {
					if (bsz < 3) {
						ZSW_FAIL(OTZ_ST_DATA);
					}
					regen = (b0 >> 4) | (bp[1] << 4) | (bp[2] << 12);
					lhdr = 3;
				}
			} else {
				if (bsz < 5) {
					ZSW_FAIL(OTZ_ST_DATA);
				}
				const uint64_t v = (uint64_t)ld_le32(bp) | ((uint64_t)bp[4] << 32);
				if (sf <= 1) {
					regen = (uint32_t)(v >> 4) & 0x3FF;
					lcomp = (uint32_t)(v >> 14) & 0x3FF;
					lhdr = 3;
					nstreams = sf == 0 ? 1 : 4;
				} else if (sf == 2) {
					regen = (uint32_t)(v >> 4) & 0x3FFF;
					lcomp = (uint32_t)(v >> 18) & 0x3FFF;
					lhdr = 4;
					nstreams = 4;
				} else {
					regen = (uint32_t)(v >> 4) & 0x3FFFF;
					lcomp = (uint32_t)(v >> 22) & 0x3FFFF;
					lhdr = 5;
					nstreams = 4;
				}
			}
			if (regen > ZS_BLOCK_MAX) {
				ZSW_FAIL(OTZ_ST_DATA);
			}
			if (regen > cap - T.nl) {
				ZSW_FAIL(OTZ_ST_OVERFLOW);   // every literal is an output byte
			}
			const uint32_t lsec = lhdr + (ltype == 0 ? regen : ltype == 1 ? 1u : lcomp);
			if (lsec > bsz) {
				ZSW_FAIL(OTZ_ST_DATA);
			}
			B.bp = bp;
			B.bsz = bsz;
			B.ltype = ltype;
			B.regen = regen;
			B.lcomp = lcomp;
			B.lhdr = lhdr;
			B.nstreams = nstreams;
			B.lsec = lsec;
			return 0;
		}
	}
#undef ZSW_FAIL
}

// what a kernel found for an entry: 0 = accepted, else (position of the error << 8) | status code.  The literal kernel's
// part of a block comes before the sequence kernel's.
__device__ __forceinline__ uint32_t zs_verdict(const ZsWalk &W, int ph) { return W.err ? ((2u * W.pos + (uint32_t)ph) << 8) | (uint32_t)W.err : 0u; }

// ------------------------------------------------------------------------------------------------ literals
struct __align__(16) ZsLitSmem {
	uint16_t huf[1 << ZS_HUF_LOG_MAX];   // (symbol << 8) | nbits
	uint32_t seq_ll[64];                 // FSE table of the weights (zs_read_huf)
	uint32_t huf_log, have_huf;
	uint32_t k;                          // list slot fetched by the group's first lane
	uint32_t pad;                        // 4368 bytes = 1092 words = 4 (mod 32): equal indices of the 8 groups fall into different banks
};
#define ZS_LIT_WARPS 1   // (35 KB of shared memory per CTA: one fits next to a CTA of k_zstd_seq, which runs at the same time)

// grid: persistent, ZS_LIT_WARPS warps per CTA, 8 entries per warp (4 lanes each)
__global__ void __launch_bounds__(32 * ZS_LIT_WARPS) k_zstd_lit(const uint8_t *__restrict__ archive, const otz_entry *__restrict__ ents,
	const OtzEntryState *__restrict__ est, const int32_t *__restrict__ status, const uint32_t *__restrict__ list, uint32_t n_list,
	uint8_t *__restrict__ scratch, const uint64_t *__restrict__ tok_ofs, uint32_t *__restrict__ litres, uint32_t *__restrict__ work_counter) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const uint32_t lane = threadIdx.x & 31u, sub = lane & 3u;
	const uint32_t gmask = 0xFu << (lane & ~3u);
	ZsLitSmem &S = reinterpret_cast<ZsLitSmem *>(smem_raw)[threadIdx.x >> 2];
	ZsWalk W;
	ZsTok T;
	ZsBlk B;
	uint32_t k = 0;
	bool busy = false, dry = false;
	W.init(nullptr, 0, 0);
	T.L = nullptr;
	T.seq_end = nullptr;
	T.nl = T.nseq = T.run = 0;
	for (;;) {
		// ---- every group: on to its next Huffman-coded literals section (the four lanes of a group move together)
		bool blk = false;
		while (!blk && !dry) {
			if (!busy) {
				if (sub == 0) {
					S.k = atomicAdd(work_counter, 1u);
				}
				__syncwarp(gmask);
				k = S.k;
				__syncwarp(gmask);
				if (k >= n_list) {
					dry = true;
					break;
				}
				const uint32_t ei = list[k];
				if (status[ei] != OTZ_ST_PENDING) {
					if (sub == 0) {
						litres[k] = 0u;   // resolved as a reference container (or failed earlier): k_zstd_seq skips it too
					}
					continue;
				}
				const otz_entry e = ents[ei];
				W.init(archive + est[ei].data_ofs, e.comp_size, e.uncomp_size);
				T.L = scratch + tok_ofs[k];
				T.nl = T.run = 0;
				S.have_huf = 0;
				busy = true;
			}
			if (zs_walk<ZS_PH_LIT>(W, T, B, sub, 4u)) {
				if (sub == 0) {
					litres[k] = zs_verdict(W, ZS_PH_LIT);
				}
				busy = false;
				continue;
			}
			if (W.fl & ZSW_FRESH) {
				W.fl &= ~ZSW_FRESH;
				__syncwarp(gmask);
				S.have_huf = 0;
				__syncwarp(gmask);
			}
			uint8_t *const lit = T.L + T.nl;   // the literals of this block go straight into the token scratch
			if (B.ltype == 0) {
				for (uint32_t i = sub; i < B.regen; i += 4) {
					lit[i] = B.bp[B.lhdr + i];
				}
				T.nl += B.regen;
			} else if (B.ltype == 1) {
				const uint8_t b = B.bp[B.lhdr];
				for (uint32_t i = sub; i < B.regen; i += 4) {
					lit[i] = b;
				}
				T.nl += B.regen;
			} else {
				blk = true;
			}
		}
		if (!__any_sync(0xFFFFFFFFu, blk)) {
			break;   // (no group has a block, so every group ran dry)
		}
		// ---- tables: the four lanes of a group build the same table with the same stores
		const uint8_t *sp = nullptr;
		uint32_t sn = 0, cnt = 0;
		uint8_t *dst = nullptr;
		if (blk) {
			const uint8_t *hp = B.bp + B.lhdr;
			uint32_t hrem = B.lcomp;
			OTZ_CHK(T.nl + B.regen <= W.cap, OTZ_CK_ZS_LIT);
			__syncwarp(gmask);
			if (B.ltype == 2) {
				const uint32_t used = zs_read_huf(S, hp, hrem);
				if (used == 0) {
					W.err = OTZ_ST_DATA;
				}
				hp += used;
				hrem -= used;
			} else if (!S.have_huf) {
				W.err = OTZ_ST_DATA;   // treeless literals without a tree
			}
			__syncwarp(gmask);
			if (!W.err) {
				if (B.nstreams == 1) {
					if (sub == 0) {
						sp = hp;
						sn = hrem;
						cnt = B.regen;
						dst = T.L + T.nl;
					}
				} else if (hrem < 6) {
					W.err = OTZ_ST_DATA;
				} else {
					const uint32_t s1 = ld_le16(hp), s2 = ld_le16(hp + 2), s3 = ld_le16(hp + 4);
					const uint32_t q = (B.regen + 3) / 4;
					if (6ull + s1 + s2 + s3 > hrem || B.regen < 3 * q) {
						W.err = OTZ_ST_DATA;
					} else {
						const uint32_t o0 = 6, o1 = 6 + s1, o2 = o1 + s2, o3 = o2 + s3;
						const uint32_t so = sub == 0 ? o0 : sub == 1 ? o1 : sub == 2 ? o2 : o3;
						const uint32_t se = sub == 0 ? o1 : sub == 1 ? o2 : sub == 2 ? o3 : hrem;
						sp = hp + so;
						sn = se - so;
						cnt = sub == 3 ? B.regen - 3 * q : q;
						dst = T.L + T.nl + sub * q;
					}
				}
			}
		}
		// ---- decode: one stream per lane, five symbols per 64-bit window (zs_huf_stream in lock-step over the warp)
		bool ok = true;
		if (blk && !W.err && (sub == 0 || B.nstreams == 4)) {
			ok = zs_huf_stream(S, sp, sn, dst, cnt);
		}
		if (blk) {
			// (the group's lanes leave the decoder at different times: its verdict needs all four)
			const uint32_t bad = __ballot_sync(gmask, !ok) & gmask;
			if (bad && !W.err) {
				W.err = OTZ_ST_DATA;
			}
			T.nl += B.regen;
		}
	}
}

// ------------------------------------------------------------------------------------------------ sequences
#define ZS_SEQ_TAB_BYTES (2 * ((1 << ZS_LL_LOG_MAX) + (1 << ZS_OF_LOG_MAX) + (1 << ZS_ML_LOG_MAX)))   // 16-bit FSE tables of one lane: 2.5 KiB
#define ZS_SEQ_WARPS 8      // (the kernel is bound by the latency of the per-entry chain: more warps with fewer lanes each hide a little
                            // more of it — 10,000 entries: 3 warps 9.9 ms, 4: 9.3, 6: 9.1, 8: 8.6)
#define ZS_SEQ_LPW_MAX 10   // lanes per warp that fit one SM's shared memory (8 x 10 x 2.5 KiB = 200 KiB)

// grid: persistent, ONE CTA of ZS_SEQ_WARPS warps per SM; the first `lpw` lanes of every warp work, each with its FSE
// tables in its own slot of the dynamic shared memory (ZS_SEQ_WARPS * lpw * ZS_SEQ_TAB_BYTES)
__global__ void __launch_bounds__(32 * ZS_SEQ_WARPS, 1) k_zstd_seq(const uint8_t *__restrict__ archive, const otz_entry *__restrict__ ents,
	const OtzEntryState *__restrict__ est, int32_t *__restrict__ status, const uint32_t *__restrict__ list, uint32_t n_list,
	uint8_t *__restrict__ scratch, const uint64_t *__restrict__ tok_ofs, I2TokRes *__restrict__ tokres, uint32_t *__restrict__ seqres,
	uint32_t *__restrict__ work_counter, uint32_t lpw) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	// code -> base | extra bits << 24 (the lanes index these with different codes: shared memory, not the constant bank)
	__shared__ uint32_t s_ll[36], s_ml[53];
	const uint32_t lane = threadIdx.x & 31u;
	for (uint32_t i = threadIdx.x; i < 36; i += blockDim.x) {
		s_ll[i] = c_zs_ll_base[i] | ((uint32_t)c_zs_ll_bits[i] << 24);
	}
	for (uint32_t i = threadIdx.x; i < 53; i += blockDim.x) {
		s_ml[i] = c_zs_ml_base[i] | ((uint32_t)c_zs_ml_bits[i] << 24);
	}
	__syncthreads();
	uint16_t *const t_ll = reinterpret_cast<uint16_t *>(smem_raw + (size_t)((threadIdx.x >> 5) * lpw + min(lane, lpw - 1u)) * ZS_SEQ_TAB_BYTES);
	uint16_t *const t_of = t_ll + (1 << ZS_LL_LOG_MAX);
	uint16_t *const t_ml = t_of + (1 << ZS_OF_LOG_MAX);
	ZsWalk W;
	ZsTok T;
	ZsBlk B;
	ZsBackW b;
	uint32_t k = 0, ei = 0;
	const uint8_t *rec_floor = nullptr;   // (bounds-check build: the records of the entry must stay above its literals)
	bool busy = false, dry = lane >= lpw;
	uint32_t rep1 = 1, rep2 = 4, rep3 = 8;
	uint32_t ll_log = 0, of_log = 0, ml_log = 0, have_ll = 0, have_of = 0, have_ml = 0;
	W.init(nullptr, 0, 0);
	T.L = nullptr;
	T.seq_end = nullptr;
	T.nl = T.nseq = T.run = 0;
	b.init(nullptr, 0);
	for (;;) {
		// ---- every lane: on to its next block that has sequences
		bool blk = false;
		uint32_t nseq = 0, regen = 0, lit_pos = 0;
		const uint8_t *sp = nullptr;
		uint32_t srem = 0, shdr = 0;
		while (!blk && !dry) {
			if (!busy) {
				k = atomicAdd(work_counter, 1u);
				if (k >= n_list) {
					dry = true;
					break;
				}
				ei = list[k];
				tokres[k].ok = 0u;
				seqres[k] = 0u;
				if (status[ei] != OTZ_ST_PENDING) {
					continue;   // resolved as a reference container (or failed earlier)
				}
				const otz_entry e = ents[ei];
				W.init(archive + est[ei].data_ofs, e.comp_size, e.uncomp_size);
				T.seq_end = reinterpret_cast<uint2 *>(scratch + tok_ofs[k + 1]);
				T.nl = T.nseq = T.run = 0;
				rec_floor = scratch + tok_ofs[k];
				busy = true;
			}
			if (zs_walk<ZS_PH_SEQ>(W, T, B, 0u, 1u)) {
				// the entry is finished: the verdict of both kernels
				if (!W.err) {
					W.err = W.n_frames == 0 ? OTZ_ST_DATA : (W.op == W.cap ? 0 : OTZ_ST_SIZE);
					W.pos++;
				}
				// this kernel's verdict; k_zstd_join combines it with the literal kernel's (the two run at the same time)
				const uint32_t vs = zs_verdict(W, ZS_PH_SEQ);
				seqres[k] = vs;
				if (vs == 0u) {
					I2TokRes r;
					r.nseq = T.nseq;
					r.nlit = T.nl;
					r.status = OTZ_ST_OK | OTZ_STF_REF_EOB;   // a valid stream that the reference rejects
					r.ok = 1u;
					tokres[k] = r;
				}
				busy = false;
				continue;
			}
			if (W.fl & ZSW_FRESH) {
				W.fl &= ~ZSW_FRESH;
				rep1 = 1;
				rep2 = 4;
				rep3 = 8;
				have_ll = have_of = have_ml = 0;
			}
			T.nl += B.regen;
			// ---- sequences section header
			sp = B.bp + B.lsec;
			srem = B.bsz - B.lsec;
			if (srem < 1) {
				W.err = OTZ_ST_DATA;
				continue;
			}
			nseq = sp[0];
			shdr = 1;
			if (nseq >= 128) {
				if (nseq == 255) {
					if (srem < 3) {
						W.err = OTZ_ST_DATA;
						continue;
					}
					nseq = sp[1] + (sp[2] << 8) + 0x7F00;
					shdr = 3;
				} else {
					if (srem < 2) {
						W.err = OTZ_ST_DATA;
						continue;
					}
					nseq = ((nseq - 128) << 8) + sp[1];
					shdr = 2;
				}
			}
			regen = B.regen;
			if (nseq == 0) {
				// literals only
				if (W.cap - W.op < regen) {
					W.err = OTZ_ST_OVERFLOW;
					continue;
				}
				T.run += regen;
				W.op += regen;
				continue;
			}
			if (srem < shdr + 1) {
				W.err = OTZ_ST_DATA;
				continue;
			}
			blk = true;
		}
		if (!__any_sync(0xFFFFFFFFu, blk)) {
			break;   // (no lane has a block, so every lane ran dry)
		}
		// ---- tables and initial states
		uint32_t st_ll = 0, st_of = 0, st_ml = 0;
		if (blk) {
			const uint32_t modes = sp[shdr];
			if (modes & 3) {
				W.err = OTZ_ST_DATA;
			} else {
				uint32_t o = shdr + 1;
				int r = zs_seq_table(modes >> 6, sp + o, srem - o, t_ll, &ll_log, &have_ll, c_zs_ll_default, 36, 6, 35, ZS_LL_LOG_MAX);
				if (r >= 0) {
					o += r;
					r = zs_seq_table((modes >> 4) & 3, sp + o, srem - o, t_of, &of_log, &have_of, c_zs_of_default, 29, 5, 31, ZS_OF_LOG_MAX);
				}
				if (r >= 0) {
					o += r;
					r = zs_seq_table((modes >> 2) & 3, sp + o, srem - o, t_ml, &ml_log, &have_ml, c_zs_ml_default, 53, 6, 52, ZS_ML_LOG_MAX);
				}
				if (r >= 0) {
					o += r;
				}
				if (r < 0 || o > srem || !b.init(sp + o, srem - o)) {
					W.err = OTZ_ST_DATA;
				} else {
					st_ll = b.read(ll_log);
					st_of = b.read(of_log);
					st_ml = b.read(ml_log);
				}
			}
			blk = !W.err;
		}
		// ---- the sequences, all lanes in lock-step
		uint32_t op = W.op;
		const uint32_t cap = W.cap, frame_start = W.frame_start;
		int32_t err = 0;
		uint32_t kk = 0;
		bool on = blk;
		while (__any_sync(0xFFFFFFFFu, on)) {
			if (on) {
				uint32_t lc, oc, mc, u_ll, u_of, u_ml, b_ll, b_of, b_ml;
				OTZ_CHK(st_ll < (1u << ll_log) && st_of < (1u << of_log) && st_ml < (1u << ml_log), OTZ_CK_ZS_TABLE);
				zs_fse16(t_ll[st_ll], ll_log, lc, u_ll, b_ll);
				zs_fse16(t_of[st_of], of_log, oc, u_of, b_of);
				zs_fse16(t_ml[st_ml], ml_log, mc, u_ml, b_ml);
				if (lc > 35 || mc > 52 || oc > 31) {
					err = OTZ_ST_DATA;
					on = false;
				} else {
					// one window for the whole sequence: three extra-bit fields, then (unless it is the last
					// sequence) the three state updates, in the RFC 8878 4.1.1 order
					const bool last_seq = kk + 1 >= nseq;
					const uint32_t xl = s_ll[lc], xm = s_ml[mc];
					const uint32_t n_ml = xm >> 24, n_ll = xl >> 24;
					if (last_seq) {
						u_ll = u_ml = u_of = 0u;
					}
					uint64_t x = b.top64();
					uint32_t ofv;
					if (oc + n_ml + n_ll + u_ll + u_ml + u_of > 64u) {
						ofv = (1u << oc) + zs_take(x, oc);   // (offset codes of 30+ bits: the rest is at most 58 bits)
						b.skip(oc);
						x = b.top64();
						b.skip(n_ml + n_ll + u_ll + u_ml + u_of);
					} else {
						b.skip(oc + n_ml + n_ll + u_ll + u_ml + u_of);
						ofv = (1u << oc) + zs_take(x, oc);
					}
					const uint32_t mlv = (xm & 0xFFFFFFu) + zs_take(x, n_ml);
					const uint32_t llv = (xl & 0xFFFFFFu) + zs_take(x, n_ll);
					st_ll = b_ll + zs_take(x, u_ll);
					st_ml = b_ml + zs_take(x, u_ml);
					st_of = b_of + zs_take(x, u_of);
					// repeat offsets (RFC 8878 3.1.1.5), without branches: idx 0 = a new offset, 1..3 = rep1..3, 4 = rep1 - 1
					const uint32_t idx = ofv > 3u ? 0u : ofv + (llv == 0u ? 1u : 0u);
					const uint32_t offset = idx == 0u ? ofv - 3u : idx == 1u ? rep1 : idx == 2u ? rep2 : idx == 3u ? rep3 : rep1 - 1u;
					const bool bad = b.pos() < 0 || offset == 0u;
					rep3 = (idx == 0u || idx >= 3u) ? rep2 : rep3;
					rep2 = idx == 1u ? rep2 : rep1;
					rep1 = offset;
					if (bad) {
						err = OTZ_ST_DATA;
						on = false;
					} else if (llv > regen - lit_pos || llv + mlv > cap - op || offset > op + llv - frame_start) {   // (llv, mlv < 2^18)
						err = llv > regen - lit_pos ? OTZ_ST_DATA : (llv + mlv > cap - op ? OTZ_ST_OVERFLOW : OTZ_ST_DATA);
						on = false;
					} else {
						op += llv;
						lit_pos += llv;
						T.run += llv;
						T.match(mlv, offset);
						OTZ_CHK(reinterpret_cast<const uint8_t *>(T.seq_end - T.nseq) >= rec_floor + T.nl, OTZ_CK_ZS_REC);
						op += mlv;
						kk++;
						on = kk < nseq;
					}
				}
			}
		}
		if (blk) {
			if (!err && b.pos() != 0) {
				err = OTZ_ST_DATA;   // the bitstream must be consumed exactly
			}
			// ---- literals after the last sequence stay pending in the run
			const uint32_t tail = regen - lit_pos;
			if (!err && cap - op < tail) {
				err = OTZ_ST_OVERFLOW;
			}
			if (err) {
				W.err = err;
			} else {
				T.run += tail;
				W.op = op + tail;
			}
		}
	}
}

// the verdicts of the two tokenizers: an entry is executed when both accepted it, else it gets the status of the error
// that comes first in the stream (0 = accepted, else (position << 8) | status code)
__global__ void k_zstd_join(const uint32_t *__restrict__ list, uint32_t n_list, const uint32_t *__restrict__ litres, const uint32_t *__restrict__ seqres,
	I2TokRes *__restrict__ tokres, int32_t *__restrict__ status) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_list) {
		return;
	}
	const uint32_t vl = litres[k], vs = seqres[k];
	const uint32_t v = vs == 0u ? vl : vl == 0u ? vs : min(vs, vl);
	if (v) {
		tokres[k].ok = 0u;
		status[list[k]] = (int32_t)(v & 0xFFu);
	}
}
