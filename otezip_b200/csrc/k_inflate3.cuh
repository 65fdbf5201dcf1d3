// k_inflate3.cuh — phase A of the two-phase raw-DEFLATE (RFC 1951) decoder: one WARP per stream, the 32 lanes
// decode 32 consecutive pieces of the stream speculatively and synchronise on the symbol boundaries.
//
// Replaces the entropy stage of inflate() as driven by otezip_extract_entry
// (/root/reference/src/lib/otezip.c:503-529; decoder src/lib/deflate-dec.inc.c:547-831, "dec" below).
//
// A Huffman-coded DEFLATE block is a serial bit chain, but a decoder that starts at a wrong bit offset falls into
// step with the true chain after a few dozen symbols (the code is a prefix code; the literal/length and distance
// alphabets alternate only inside a match).  The warp therefore cuts the block into ROUNDS of 32 x S bits
// (S = 256 or 512) and, with ONE table set per block in shared memory:
//
//   pass 1   lane j decodes from bit R + j*S (a guess, except for lane 0) to the end of its piece and records
//            every symbol boundary it visits in a bitmap and the position where it left the piece (its exit);
//   sync     lane j restarts at the exit of lane j-1 and walks until it steps on a boundary of its own pass 1 —
//            from there on the two decodes are identical, so its exit is the one pass 1 found — or leaves the piece
//            at a new exit; repeated until no exit changes (lane 0 is exact from the start, so after iteration i
//            lanes 0..i are exact: at most 31 iterations, 2-3 in practice);
//   decode   every lane decodes its piece from its true start and writes the tokens — literal bytes and one 32-bit
//            record per match {literal run : 9, length - 3 : 8, distance - 1 : 15} — into its own temp slot in
//            global memory (L2 resident), counting literals, matches and output bytes on the way;
//   place    warp prefix sums give every lane its place in the token stream and in the output, and the tokens move
//            there: literals dense and ascending, records descending — the format k_inflate_lz (k_inflate2.cuh)
//            executes.  (Round 2 first decoded every piece twice here, once to count and once to emit.)
//
// Three decodes of every symbol instead of one, but 32 symbols per warp instruction on one stream: a 64 KiB entry
// is ~9 rounds instead of 4,400 serial steps, a 16 MiB entry needs neither a block search nor a lane per block, and
// table memory is per warp, not per stream (32 warps per SM).  Block headers (dec:122-266) are parsed by the warp
// in lock-step from a register bit buffer; the tables are built by the warp (i2_build_table, k_inflate2.cuh).
//
// Like its predecessor the kernel only commits streams that are plainly valid (final block reached, exactly
// uncomp_size bytes, regular stored-block headers, tables within the budget, no distance beyond the output);
// anything else goes to the fallback list and is decoded from scratch by k_inflate, whose status words define the
// behaviour in those cases.  The reference's end-of-input rule (dec:811-816) is evaluated on the true
// chain and reported as OTZ_STF_REF_EOB exactly as k_inflate reports it.
//
// HUGE streams (list slots below n_huge): the output position of every round is known here, so the token stream is
// cut into segments at round boundaries (each >= 64 KiB of output) for the parallel execution over 16-bit symbols
// (k_inflate_lz<.., PAR>, k_seg_window, k_seg_translate): the segment table I2SegCtl is filled directly and
// k_seg_stitch only has to place the segments in the symbol buffer.
#pragma once
#include "k_inflate2.cuh"

#define I3_WARPS 4         // NW = 1: independent warps (streams) per CTA
// WPLMAX = words per piece (lane and round), at most: 16 (512 bits) or, for CTA groups when a batch has few huge streams, 32
#define I3_BAD 4u          // (kinds 0..3 are I2_K_*)
// temp slot of a lane: the tokens of its piece before their place in the stream's scratch is known (a piece spans at most
// 32 * WPLMAX + 20 bits; a literal code has >= 1 bit, a match >= 2)
#define I3_TMP_LIT 576u
#define I3_TMP_REC 272u
#define I3_TMP_BYTES (I3_TMP_LIT + 4u * I3_TMP_REC)
#define I3_SEG_MIN 524288u // least output bytes of a segment of a huge stream (k_seg_window resolves the last 32 KiB of each one serially)

struct I3BuildScratch {
	uint32_t cnt[16];
	uint32_t first15[16];
	uint32_t limit15[16];
	uint32_t offs[16];
	uint32_t run[16];
	uint16_t sorted[320];
};

// shared memory of one GROUP = the NW warps that decode one stream together (NW * 32 pieces per round)
template <int NW, int WPLMAX = 16>
struct __align__(16) I3Smem {
	static constexpr int G = 32 * NW;
	uint16_t lit[I2_LIT_CAP];
	uint16_t dst[I2_DST_CAP];
	uint32_t xo[G], xk[G];   // exit (bit position, kind) of every piece
	uint32_t ws[NW][12];     // per-warp partial results
	uint32_t bc[8];          // broadcast slots
	union {
		struct {
			uint32_t stage[G * (WPLMAX + 3)];   // words of piece t at [t * (words per piece + 3) ...]: odd stride, no bank conflicts between lanes at the same offset
			uint32_t vis[WPLMAX * G];           // symbol boundaries visited in pass 1: word w of piece t at [w * G + t]
		} r;
		struct {
			I3BuildScratch b;
			uint8_t lens[320];
			uint8_t pre[128];
		} h;
	} u;
};

template <int NW>
__device__ __forceinline__ void i3_sync() {
	if (NW == 1) {
		__syncwarp();
	} else {
		__syncthreads();
	}
}
template <int NW>
__device__ __forceinline__ bool i3_any(bool p) {
	if (NW == 1) {
		return __any_sync(0xFFFFFFFFu, p) != 0;
	} else {
		return __syncthreads_or(p) != 0;
	}
}

// One symbol of the chain at bit q of the lane's staged piece (pw[i] = word i of the stream, counted from the aligned
// base): kind (I2_K_LIT / LEN / EOB, I3_BAD), `next` = the bit behind it (behind the distance code for a match).
// LVL >= 1: also the literal byte or match length (`val`); LVL 2: and the distance.  Straight-line: the distance
// look-up is done by every lane and used by the matches, so that a warp step is one pass over the same instructions;
// only second-level tables (rare on a true chain) branch.
template <int LVL>
__device__ __forceinline__ uint32_t i3_step(const uint16_t *__restrict__ lit, const uint16_t *__restrict__ dst, const uint32_t *pw, uint32_t q, uint32_t Pend,
	uint32_t &next, uint32_t &val, uint32_t &dist) {
	const uint32_t *w = pw + (q >> 5);
	const uint32_t lo = w[0], hi = w[1], nx = w[2];
	const uint32_t bits = __funnelshift_r(lo, hi, q);
	uint32_t e = lit[bits & ((1u << I2_LIT_ROOT) - 1u)];
	uint32_t used = 0;
	if (((e >> 4) & 3u) == I2_K_LINK) {
		const uint32_t sb = I2_LIT_LINK_BITS(e);   // 0: no code matches (dec:693-695) — e stays a LINK
		const uint32_t e2 = lit[I2_LIT_LINK_OFS(e) + ((bits >> I2_LIT_ROOT) & ((1u << sb) - 1u))];
		e = sb ? e2 : e;
		used = I2_LIT_ROOT;
	}
	const uint32_t kind = (e >> 4) & 3u, tb = used + (e & 15u);
	const uint32_t q2 = q + tb;   // tb <= 20: the distance code starts in the same word or the next one
	const bool same = ((q2 ^ q) >> 5) == 0u;
	const uint32_t bits2 = __funnelshift_r(same ? lo : hi, same ? hi : nx, q2);
	uint32_t d = dst[bits2 & ((1u << I2_DST_ROOT) - 1u)];
	uint32_t used2 = 0;
	const bool isl = kind == I2_K_LEN;
	if (isl && (d >> 14) != 0u) {
		const uint32_t sb = I2_DST_LINK_BITS(d);   // 0: dec:762-764 — d stays a LINK
		const uint32_t d2 = dst[I2_DST_LINK_OFS(d) + ((bits2 >> I2_DST_ROOT) & ((1u << sb) - 1u))];
		d = sb ? d2 : d;
		used2 = I2_DST_ROOT;
	}
	const uint32_t t3 = used2 + (d & 31u);
	next = isl ? q2 + t3 : q2;
	if (LVL >= 1) {
		const uint32_t xb = (e >> 6) & 7u;
		const uint32_t len = i2_len_base((e >> 9) & 31u, xb) + ((bits >> (tb - xb)) & ((1u << xb) - 1u));
		val = isl ? len : (e >> 6) & 0xFFu;
	}
	if (LVL >= 2) {
		const uint32_t x3 = (d >> 5) & 15u;
		dist = i2_dist_base((d >> 9) & 31u, x3) + ((bits2 >> (t3 - x3)) & ((1u << x3) - 1u));
	}
	const bool bad = kind == I2_K_LINK || (isl && (d >> 14) != 0u) || q >= Pend;   // invalid code / out of input
	return bad ? I3_BAD : kind;
}

// Bit reader of the header parser: every thread of the group runs it on the same stream (uniform, no divergence),
// 64-bit window in registers refilled one aligned word at a time.  P = position of the window's bit 0, in bits from `inw`.
struct I3HdrBits {
	const uint32_t *inw;
	uint32_t nw;      // words of the stream (reads beyond return 0)
	uint32_t wi;      // next word to load
	uint64_t buf;
	uint32_t have;
	uint32_t P;
	__device__ __forceinline__ void init(const uint32_t *inw_, uint32_t nw_, uint32_t P_) {
		inw = inw_;
		nw = nw_;
		P = P_;
		wi = P_ >> 5;
		const uint32_t a = wi < nw ? __ldg(inw + wi) : 0u, b = wi + 1u < nw ? __ldg(inw + wi + 1u) : 0u;
		buf = (((uint64_t)b << 32) | a) >> (P_ & 31u);
		have = 64u - (P_ & 31u);
		wi += 2;
	}
	__device__ __forceinline__ void fill() {   // > 32 valid bits afterwards
		if (have <= 32u) {
			const uint32_t a = wi < nw ? __ldg(inw + wi) : 0u;
			buf |= (uint64_t)a << have;
			have += 32u;
			wi++;
		}
	}
	__device__ __forceinline__ uint32_t peek() const { return (uint32_t)buf; }
	__device__ __forceinline__ void skip(uint32_t n) {
		buf >>= n;
		have -= n;
		P += n;
	}
};

// what the header parser asks for
#define I3_A_NONE 0u       // stands at the next block header
#define I3_A_BUILD 1u      // code lengths are in S.u.h.lens: build the tables, then decode symbols
#define I3_A_COMMIT 2u     // the stream (or chunk) ended here
#define I3_A_FALLBACK 3u   // k_inflate takes the stream
#define I3_A_STORED 4u     // a stored block with payload

// exclusive prefix sums of (a, b) over the group + their totals.  NW > 1: per-warp sums go through S.ws[.][slot, slot + 1]
// (one barrier; the slots are not reused before the next barrier of the caller)
template <int NW, typename SM>
__device__ __forceinline__ void i3_scan2(SM &S, uint32_t lane, uint32_t warp, int slot, uint32_t a, uint32_t b, uint32_t &apre, uint32_t &bpre,
	uint32_t &atot, uint32_t &btot) {
	uint32_t ai = a, bi = b;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, ai, d), y = __shfl_up_sync(0xFFFFFFFFu, bi, d);
		if ((int)lane >= d) {
			ai += x;
			bi += y;
		}
	}
	if (NW == 1) {
		atot = __shfl_sync(0xFFFFFFFFu, ai, 31);
		btot = __shfl_sync(0xFFFFFFFFu, bi, 31);
		apre = ai - a;
		bpre = bi - b;
	} else {
		if (lane == 31) {
			S.ws[warp][slot] = ai;
			S.ws[warp][slot + 1] = bi;
		}
		__syncthreads();
		uint32_t ao = 0, bo = 0, at = 0, bt = 0;
#pragma unroll
		for (int v = 0; v < NW; v++) {
			const uint32_t x = S.ws[v][slot], y = S.ws[v][slot + 1];
			ao += (uint32_t)v < warp ? x : 0u;
			bo += (uint32_t)v < warp ? y : 0u;
			at += x;
			bt += y;
		}
		apre = ai - a + ao;
		bpre = bi - b + bo;
		atot = at;
		btot = bt;
	}
}

// grid: persistent.  NW = 1: I3_WARPS independent warps per CTA, one stream each; NW > 1: the NW warps of a CTA
// decode one (huge) stream together.  Every group pulls list slots [k0, k1) from *work_counter (longest streams first).
template <int NW, int MINB = (NW == 1 ? 8 : 16 / NW), int WPLMAX = 16>
__global__ void __launch_bounds__(32 * (NW == 1 ? I3_WARPS : NW), MINB) k_inflate_spec(const uint8_t *__restrict__ archive,
	const otz_entry *__restrict__ ents, const OtzEntryState *__restrict__ est, const int32_t *__restrict__ status, const uint32_t *__restrict__ list,
	uint32_t k0, uint32_t k1, uint32_t *__restrict__ work_counter, uint8_t *__restrict__ scratch, const uint64_t *__restrict__ tok_ofs,
	I2TokRes *__restrict__ tokres, uint32_t *__restrict__ fb_list, uint32_t *__restrict__ fb_count, uint32_t n_huge, I2SegCtl seg, uint32_t seg_min,
	uint8_t *__restrict__ tmp) {
	static_assert(32 * WPLMAX + 20 <= I3_TMP_LIT && (32 * WPLMAX + 20) / 2 <= I3_TMP_REC, "temp slot too small for this piece length");
	extern __shared__ __align__(16) uint8_t smem_raw[];
	constexpr uint32_t G = 32u * NW;
	constexpr uint32_t I3_TMP_SMEM = 4u * WPLMAX;   // literals of a piece kept in shared memory: the WPLMAX words of the lane's boundary bitmap
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = NW == 1 ? 0u : threadIdx.x >> 5;   // warp within the group
	const uint32_t tid = NW == 1 ? lane : threadIdx.x;        // thread within the group = its piece
	typedef I3Smem<NW, WPLMAX> Smem;
	Smem &S = reinterpret_cast<Smem *>(smem_raw)[NW == 1 ? threadIdx.x >> 5 : 0];
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint8_t *const tml = tmp + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * I3_TMP_BYTES;   // this lane's temp literals
	uint32_t *const tmr = reinterpret_cast<uint32_t *>(tml + I3_TMP_LIT);                          // and records

	for (;;) {
		uint32_t k = 0;
		if (NW == 1) {
			if (lane == 0) {
				k = atomicAdd(work_counter, 1u);
			}
			k = __shfl_sync(0xFFFFFFFFu, k, 0);
		} else {
			__syncthreads();
			if (tid == 0) {
				S.bc[0] = atomicAdd(work_counter, 1u);
			}
			__syncthreads();
			k = S.bc[0];
		}
		k += k0;
		if (k >= k1) {
			break;
		}
		const bool huge = k < n_huge;
		const uint32_t ei = list[k];
		if (OTZ_ST_CODE(status[ei]) != OTZ_ST_OK) {
			if (tid == 0) {
				if (huge) {
					seg.count[k] = 0;
				} else {
					tokres[k].ok = 0u;   // failed in k_resolve: nothing to decode
				}
			}
			continue;
		}
		const otz_entry ent = ents[ei];
		const uint8_t *const in = archive + est[ei].data_ofs;
		const uint32_t comp = ent.comp_size, cap = ent.uncomp_size, rflags = ent.flags;
		const bool chunk_mid = (rflags & OTZ_EF_CHUNK) && !(rflags & OTZ_EF_LAST_CHUNK);
		// bit positions count from the aligned word that holds the first byte
		const uint32_t skipb = (uint32_t)(reinterpret_cast<uint64_t>(in) & 3u);
		const uint32_t *const inw = reinterpret_cast<const uint32_t *>(in - skipb);
		const uint32_t P0 = 8u * skipb, Pend = P0 + 8u * comp;
		const uint32_t nw = (Pend + 31u) >> 5;
		// pieces of 512 bits once a stream is long enough to fill a few rounds of them
		// (longer pieces amortise the synchronisation walk — its length is set by the slowest of the G lanes, not by the piece —
		// and are worth their shared memory only where a stream has hundreds of rounds: the 4-warp groups)
		const uint32_t wpl = (WPLMAX >= 32 && comp >= 768u * G) ? 32u : comp >= 192u * G ? 16u : 8u;   // words per piece
		const uint32_t S_bits = 32u * wpl;
		uint8_t *const litp = scratch + tok_ofs[k];
		uint32_t *const seq_end = reinterpret_cast<uint32_t *>(scratch + tok_ofs[k + 1]);
		// totals of the stream so far; run_carry = literals since the last match (they belong to the next record)
		uint32_t P = P0, nl_tot = 0, nseq_tot = 0, ob_tot = 0, run_carry = 0;
		uint32_t final_blk = 0, ref_eob = 0;
		// HUGE: the open segment
		uint32_t nseg = 0, sg_lit0 = 0, sg_seq0 = 0, sg_out0 = 0, sg_minsrc = 0xFFFFFFFFu;
		const uint32_t sg_target = max(seg_min, cap / (I2_MAXSEG - 8u) + 1u);
		uint32_t act = comp == 0u ? I3_A_FALLBACK : I3_A_NONE;   // dec:610: k_inflate answers TRUNCATED
		if (huge && tid == 0) {
			seg.start[k * I2_MAXSEG] = 0u;
		}

// dec:811-816 as k_inflate evaluates it: after a step that leaves the stream unfinished
#define I3_STEP_CHECK(pos_)                      \
	do {                                         \
		if ((pos_) > Pend) {                     \
			act = I3_A_FALLBACK;                 \
		} else if (Pend - (pos_) < 8u) {         \
			ref_eob = 1u;                        \
		}                                        \
	} while (0)

		// (every condition below is uniform over the group: all threads take the same path to every barrier)
		while (act != I3_A_COMMIT && act != I3_A_FALLBACK) {
			// ---------------------------------------------------------------- block header (dec:613-627), uniform
			uint32_t hlit = 0, hdist = 0, st_src = 0, st_len = 0;
			{
				I3HdrBits hb;
				hb.init(inw, nw, P);
				uint32_t bits = hb.peek();
				final_blk = bits & 1u;
				const uint32_t btype = (bits >> 1) & 3u;
				hb.skip(3);
				I3_STEP_CHECK(hb.P);
				if (act == I3_A_FALLBACK) {
					break;
				}
				if (chunk_mid && final_blk) {
					act = I3_A_FALLBACK;   // a final block inside a chunk that is not the last: not a chunk k_deflate wrote
					break;
				}
				if (btype == 0u) {
					// stored block (dec:269-319).  Anything irregular — a bad length pair, a payload that runs past the
					// input — is left to k_inflate, which knows what the reference answers
					const uint32_t rem = Pend - hb.P;
					const uint32_t bpos = comp - (rem >> 3);   // the partial byte is dropped
					if (comp - bpos < 4u || (i2_ld_le16(in + bpos) ^ i2_ld_le16(in + bpos + 2)) != 0xFFFFu) {
						act = I3_A_FALLBACK;
					} else {
						st_len = i2_ld_le16(in + bpos);
						st_src = bpos + 4u;
						if (comp - st_src < st_len) {
							act = I3_A_FALLBACK;
						} else if (st_len != 0u) {
							act = I3_A_STORED;
						} else {
							P = P0 + 8u * st_src;
							if (final_blk) {
								act = I3_A_COMMIT;
							} else if (st_src >= comp) {
								act = chunk_mid ? I3_A_COMMIT : I3_A_FALLBACK;   // end of this chunk / unfinished stream out of input
							}
						}
					}
				} else if (btype == 3u) {
					act = I3_A_FALLBACK;   // dec:657-658
				} else if (btype == 1u) {
					hb.fill();
					if ((hb.peek() & 127u) == 0u) {
						// empty fixed block (zlib's Z_FINISH tail): end-of-block is the 7-bit code 0000000
						hb.skip(7);
						P = hb.P;
						if (final_blk) {
							act = I3_A_COMMIT;
						} else {
							I3_STEP_CHECK(P);
						}
					} else {
						i3_sync<NW>();
						for (uint32_t i = tid; i < 320u; i += G) {   // dec:322-349
							S.u.h.lens[i] = i < 144u ? 8 : i < 256u ? 9 : i < 280u ? 7 : i < 288u ? 8 : 5;
						}
						hlit = 288;
						hdist = 32;
						P = hb.P;
						act = I3_A_BUILD;
					}
				} else {
					// dynamic block header, dec:122-266
					uint8_t *const lens = S.u.h.lens;
					uint8_t *const pre = S.u.h.pre;
					hb.fill();
					bits = hb.peek();
					hlit = (bits & 31u) + 257u;
					hdist = ((bits >> 5) & 31u) + 1u;
					const uint32_t hclen = ((bits >> 10) & 15u) + 4u;
					hb.skip(14);
					bool bad = hlit > 286u || hdist > 30u;
					uint64_t cl = 0;    // 19 code-length-code lengths, 3 bits each
					uint64_t cnt = 0;   // packed byte counters per length
					for (uint32_t i = 0; i < hclen; i++) {
						hb.fill();
						const uint32_t v = hb.peek() & 7u;
						hb.skip(3);
						cl |= (uint64_t)v << (3u * c_cl_order[i]);
						cnt += 1ull << (8u * v);
					}
					uint64_t nextc = 0;   // packed next canonical code per length
					{
						int left = 1;
						uint32_t code = 0;
						for (uint32_t l = 1; l <= 7; l++) {
							const uint32_t c = (uint32_t)(cnt >> (8u * l)) & 0xFFu;
							left = (left << 1) - (int)c;
							bad |= left < 0;
							code = (code + (l > 1 ? (uint32_t)(cnt >> (8u * (l - 1))) & 0xFFu : 0u)) << 1;
							nextc |= (uint64_t)(code & 0xFFu) << (8u * l);
						}
						bad |= left != 0;   // the code-length code must be complete
					}
					if (!bad) {
						i3_sync<NW>();   // (the round data of the previous block shares this memory)
						for (uint32_t s = 0; s < 19; s++) {
							const uint32_t l = (uint32_t)(cl >> (3u * s)) & 7u;
							if (l) {
								const uint32_t c = (uint32_t)(nextc >> (8u * l)) & 0xFFu;
								nextc += 1ull << (8u * l);
								const uint32_t rev = __brev(c) >> (32u - l);
								// the 2^(7-l) slots of this code, spread over the threads
								for (uint32_t x = rev + (tid << l); x < 128u; x += (G << l)) {
									pre[x] = (uint8_t)(s | (l << 5));
								}
							}
						}
						i3_sync<NW>();
						const uint32_t total = hlit + hdist;
						uint32_t idx = 0, prev = 0;
						while (idx < total) {
							hb.fill();
							bits = hb.peek();
							const uint32_t e = pre[bits & 127u];
							const uint32_t sym = e & 31u, cb = e >> 5;
							if (sym < 16u) {
								hb.skip(cb);
								if (tid == 0) {
									lens[idx] = (uint8_t)sym;
								}
								idx++;
								prev = sym;
								continue;
							}
							uint32_t rep, val = 0;
							if (sym == 16u) {   // dec:209-219
								if (idx == 0) {
									bad = true;
									break;
								}
								val = prev;
								rep = 3u + ((bits >> cb) & 3u);
								hb.skip(cb + 2);
							} else if (sym == 17u) {   // dec:221-228
								rep = 3u + ((bits >> cb) & 7u);
								hb.skip(cb + 3);
							} else {   // dec:230-237
								rep = 11u + ((bits >> cb) & 127u);
								hb.skip(cb + 7);
							}
							if (idx + rep > total) {
								bad = true;   // dec:244
								break;
							}
							for (uint32_t i = tid; i < rep; i += G) {
								lens[idx + i] = (uint8_t)val;
							}
							idx += rep;
							prev = val;
						}
						for (uint32_t i = total + tid; i < 320u; i += G) {
							lens[i] = 0;
						}
						i3_sync<NW>();
						bad = bad || lens[256] == 0;   // no end-of-block code
					}
					P = hb.P;
					act = (bad || P > Pend) ? I3_A_FALLBACK : I3_A_BUILD;
				}
			}
			if (act == I3_A_COMMIT || act == I3_A_FALLBACK) {
				break;
			}
			if (act == I3_A_STORED) {
				// payload of a stored block: the bytes join the literals of the stream (a run of 511 literals or more
				// becomes escape records when the next match is emitted)
				if (ob_tot + st_len > cap) {
					act = I3_A_FALLBACK;   // k_inflate reports the overflow (dec:296-300)
					break;
				}
				OTZ_CHK((uint64_t)st_src + st_len <= comp && litp + nl_tot + st_len <= reinterpret_cast<uint8_t *>(seq_end - nseq_tot), OTZ_CK_SPEC_STORED);
				for (uint32_t i = tid; i < st_len; i += G) {
					litp[nl_tot + i] = in[st_src + i];
				}
				nl_tot += st_len;
				ob_tot += st_len;
				run_carry += st_len;
				const uint32_t npos = st_src + st_len;
				P = P0 + 8u * npos;
				if (final_blk) {
					act = I3_A_COMMIT;
				} else if (npos >= comp) {
					act = chunk_mid ? I3_A_COMMIT : I3_A_FALLBACK;
				} else {
					act = I3_A_NONE;
				}
				continue;
			}
			if (act == I3_A_NONE) {
				continue;   // an empty stored / fixed block was consumed
			}
			// ---------------------------------------------------------------- tables of the block (one warp builds them)
			{
				i3_sync<NW>();
				int r = 0;
				if (warp == 0) {
					uint32_t lensr[10];
#pragma unroll
					for (int j = 0; j < 10; j++) {
						lensr[j] = S.u.h.lens[32 * j + lane];
					}
					__syncwarp();
					r = i2_build_table<false, I2_LIT_ROOT, I2_LIT_CAP>(S.u.h.b, lensr, 0u, hlit, S.lit);
					if (!r) {
						r = i2_build_table<true, I2_DST_ROOT, I2_DST_CAP>(S.u.h.b, lensr, hlit, hdist, S.dst);
					}
					if (NW > 1 && lane == 0) {
						S.bc[1] = (uint32_t)r;
					}
				}
				i3_sync<NW>();
				if (NW > 1) {
					r = (int)S.bc[1];
				}
				if (r) {
					act = I3_A_FALLBACK;
					break;
				}
				I3_STEP_CHECK(P);
				if (act == I3_A_FALLBACK) {
					break;
				}
			}
			// ---------------------------------------------------------------- symbols of the block, round by round
			const uint16_t *const lit = S.lit;
			const uint16_t *const dst = S.dst;
			uint32_t *const vis = S.u.r.vis;
			uint32_t *const myst = S.u.r.stage + tid * (wpl + 3u);
			bool block_done = false;
			while (!block_done) {
				const uint32_t R = P, w0 = R >> 5;
				// stage the round: every thread its piece + the 3 words behind it (zeros behind the stream)
				i3_sync<NW>();
				{
					const uint32_t wb = w0 + tid * wpl;
					for (uint32_t i = 0; i < wpl + 3u; i++) {
						myst[i] = wb + i < nw ? __ldg(inw + wb + i) : 0u;
					}
					// (the words of the next round, about G pieces further on: an L2 hit instead of an HBM round trip at its start)
					if (wb + G * wpl < nw) {
						asm volatile("prefetch.global.L2 [%0];" ::"l"(inw + wb + G * wpl));
					}
					for (uint32_t w = 0; w < wpl; w++) {
						vis[w * G + tid] = 0u;
					}
				}
				// pw[i] = word i of the stream for the words this thread may touch
				const uint32_t *const pw = myst - (w0 + tid * wpl);
				const uint32_t base = R + tid * S_bits, end = base + S_bits;
				// ---- pass 1: from the guessed start to the end of the piece
				uint32_t E, Ek;
				{
					uint32_t q = base, curw = 0, curmask = 0, kd_exit = I3_BAD;
					bool on = base < Pend;
					while (__any_sync(0xFFFFFFFFu, on)) {
						uint32_t nx, v_, d_;
						const uint32_t kd = i3_step<0>(lit, dst, pw, q, Pend, nx, v_, d_);
						if (on) {
							const uint32_t bit = q - base, w = bit >> 5;
							OTZ_CHK(bit < S_bits && (q >> 5) - (w0 + tid * wpl) <= wpl, OTZ_CK_SPEC_PIECE);
							OTZ_CHK(w < wpl, OTZ_CK_SPEC_VIS);
							if (w != curw) {
								vis[curw * G + tid] = curmask;
								curw = w;
								curmask = 0;
							}
							curmask |= 1u << (bit & 31u);
							q = kd == I3_BAD ? q : nx;
							const bool stop = kd == I3_BAD || kd == I2_K_EOB || q >= end;
							kd_exit = stop ? (kd == I3_BAD || kd == I2_K_EOB ? kd : 0u) : kd_exit;
							on = !stop;
						}
					}
					vis[curw * G + tid] = curmask;
					E = q;
					Ek = kd_exit;
				}
				// ---- synchronisation: every piece restarts at the exit of the piece before it until it meets its own pass 1
				uint32_t start = base, outp = E, outk = Ek;   // outk: 0 = left the piece, I2_K_EOB, I3_BAD
				if (NW > 1) {
					S.xo[tid] = outp;
					S.xk[tid] = outk;
					__syncthreads();
				}
				for (;;) {
					uint32_t pout, pk;
					if (NW == 1) {
						pout = __shfl_up_sync(0xFFFFFFFFu, outp, 1);
						pk = __shfl_up_sync(0xFFFFFFFFu, outk, 1);
					} else {
						pout = tid ? S.xo[tid - 1u] : 0u;
						pk = tid ? S.xk[tid - 1u] : I3_BAD;
					}
					const bool need = tid > 0u && pk == 0u && pout != start;
					if (!i3_any<NW>(need)) {   // (NW > 1: also the barrier between reading and rewriting the exits)
						break;
					}
					uint32_t q = need ? pout : base, k2 = 0u;
					bool hit = false, on = need;
					start = need ? pout : start;
					while (__any_sync(0xFFFFFFFFu, on)) {
						uint32_t nx, v_, d_;
						const uint32_t kd = i3_step<0>(lit, dst, pw, q, Pend, nx, v_, d_);
						if (on) {
							const uint32_t bit = q - base;
							OTZ_CHK(bit < S_bits && (q >> 5) - (w0 + tid * wpl) <= wpl, OTZ_CK_SPEC_PIECE);
							if ((vis[(bit >> 5) * G + tid] >> (bit & 31u)) & 1u) {
								hit = true;
								on = false;
							} else {
								q = kd == I3_BAD ? q : nx;
								const bool stop = kd == I3_BAD || kd == I2_K_EOB || q >= end;
								k2 = stop && (kd == I3_BAD || kd == I2_K_EOB) ? kd : k2;
								on = !stop;
							}
						}
					}
					if (need) {
						outp = hit ? E : q;
						outk = hit ? Ek : k2;
					}
					if (NW > 1) {
						if (need) {
							S.xo[tid] = outp;
							S.xk[tid] = outk;
						}
						__syncthreads();
					}
				}
				// ---- the true chain: pieces up to the first one that did not simply leave its range
				uint32_t T, lastl, lastk, Pnext;
				if (NW == 1) {
					const uint32_t pout = __shfl_up_sync(0xFFFFFFFFu, outp, 1);
					T = lane == 0u ? R : pout;
					const uint32_t stopm = __ballot_sync(0xFFFFFFFFu, outk != 0u);
					lastl = stopm ? (uint32_t)__ffs(stopm) - 1u : 31u;
					lastk = __shfl_sync(0xFFFFFFFFu, outk, lastl);
					Pnext = __shfl_sync(0xFFFFFFFFu, outp, lastl);
				} else {
					T = tid == 0u ? R : S.xo[tid - 1u];
					const uint32_t stopm = __ballot_sync(0xFFFFFFFFu, outk != 0u);
					if (lane == 0) {
						S.ws[warp][6] = stopm ? 32u * warp + (uint32_t)__ffs(stopm) - 1u : 0xFFFFFFFFu;
					}
					__syncthreads();
					lastl = G - 1u;
#pragma unroll
					for (int v = NW - 1; v >= 0; v--) {
						const uint32_t x = S.ws[v][6];
						lastl = x != 0xFFFFFFFFu ? x : lastl;
					}
					lastk = S.xk[lastl];
					Pnext = S.xo[lastl];
				}
				if (lastk == I3_BAD) {
					act = I3_A_FALLBACK;   // invalid code or out of input on the true chain: k_inflate names the error
					break;
				}
				const bool valid = tid <= lastl;
				// ---- decode: the tokens of the piece go to the lane's temp slot (their place in the stream is not known yet), counted on the way
				uint32_t nl = 0, nm = 0, ob = 0, lead = 0, since = 0, esc_in = 0;   // esc_in: escape records in front of matches that are not the first of the piece
				uint32_t lastp = 0;
				int32_t need = -0x40000000;   // max over the matches of (distance - output bytes of the piece in front of the match)
				{
					uint32_t q = T;
					bool on = valid;
					while (__any_sync(0xFFFFFFFFu, on)) {
						uint32_t nx, v_, d_;
						const uint32_t kd = i3_step<2>(lit, dst, pw, q, Pend, nx, v_, d_);
						if (on) {
							OTZ_CHK(q >= R && q - base < S_bits && (q >> 5) - (w0 + tid * wpl) <= wpl, OTZ_CK_SPEC_PIECE);
							const bool isl = kd == I2_K_LEN, isb = kd == I2_K_LIT;
							if (isb) {
								OTZ_CHK(nl < I3_TMP_LIT, OTZ_CK_SPEC_LIT);
								// (the first I3_TMP_SMEM literals stay on chip, in the lane's words of the boundary bitmap — not needed any
								// more in this round: one scattered byte store to global memory per literal is what the tokenizer can afford)
								if (nl < I3_TMP_SMEM) {
									reinterpret_cast<uint8_t *>(&vis[(nl >> 2) * G + tid])[nl & 3u] = (uint8_t)v_;
								} else {
									tml[nl] = (uint8_t)v_;
								}
							}
							if (isl) {
								OTZ_CHK(nm < I3_TMP_REC, OTZ_CK_SPEC_SEQ);
								need = max(need, (int32_t)d_ - (int32_t)ob);
								lead = nm == 0u ? since : lead;
								esc_in += nm != 0u ? since / I2_SEQ_ESC : 0u;   // (only a piece of more than 511 bits can hold such a run)
								tmr[nm] = (nm == 0u ? 0u : since & 511u) | ((v_ - 3u) << 9) | ((d_ - 1u) << 17);   // (first record: its run is set below)
							}
							nl += isb;
							nm += isl;
							ob += isb ? 1u : isl ? v_ : 0u;
							since = isl ? 0u : since + isb;
							lastp = (isl || isb) ? nx : lastp;
							q = nx;
							on = (isl || isb) && q < end;   // (I3_BAD cannot happen on the verified chain)
						}
					}
				}
				// place of every piece in the token stream and in the output
				uint32_t nlpre, obpre, tot_nl, tot_ob;
				// literals pending in front of the piece (since the last match of the stream before it)
				const uint32_t mm = __ballot_sync(0xFFFFFFFFu, nm != 0u);
				uint32_t nl_w = nl;   // inclusive sum inside the warp
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, nl_w, d);
					if ((int)lane >= d) {
						nl_w += x;
					}
				}
				const uint32_t Lm = mm ? 31u - (uint32_t)__clz(mm) : 0u;
				// literals pending behind the last match of this warp (or all of its literals when it has none)
				const uint32_t w_tail = __shfl_sync(0xFFFFFFFFu, since, Lm) + (__shfl_sync(0xFFFFFFFFu, nl_w, 31) - __shfl_sync(0xFFFFFFFFu, nl_w, Lm));
				if (NW > 1) {
					const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, nl_w, 31);
					if (lane == 0) {
						S.ws[warp][3] = mm != 0u;
						S.ws[warp][4] = mm ? w_tail : wtot;
					}
				}
				i3_scan2<NW, Smem>(S, lane, warp, 0, nl, ob, nlpre, obpre, tot_nl, tot_ob);   // (NW > 1: barrier — ws[.][3,4] are visible too)
				if (ob_tot + tot_ob > cap) {
					act = I3_A_FALLBACK;   // dec:700-703, dec:791-793: k_inflate reports the overflow
					break;
				}
				uint32_t w_carry = run_carry, new_carry;   // pending literals at the start of this warp / at the end of the round
				if (NW == 1) {
					new_carry = mm ? w_tail : run_carry + tot_nl;
				} else {
					uint32_t run = run_carry;
#pragma unroll
					for (int v = 0; v < NW; v++) {
						w_carry = (uint32_t)v == warp ? run : w_carry;
						run = S.ws[v][3] ? S.ws[v][4] : run + S.ws[v][4];
					}
					new_carry = run;
				}
				const uint32_t below = mm & lt_mask;
				const uint32_t Lp = below ? 31u - (uint32_t)__clz(below) : 0u;
				const uint32_t tl = __shfl_sync(0xFFFFFFFFu, since, Lp), sl = __shfl_sync(0xFFFFFFFFu, nl_w, Lp);
				const uint32_t carry_in = below ? tl + ((nl_w - nl) - sl) : w_carry + (nl_w - nl);
				const uint32_t esc = nm ? (carry_in + lead) / I2_SEQ_ESC : 0u;
				uint32_t sqpre, dummy_pre, tot_sq, dummy_tot;
				i3_scan2<NW, Smem>(S, lane, warp, 8, nm + esc + esc_in, 0u, sqpre, dummy_pre, tot_sq, dummy_tot);
				// ---- the tokens move from the temp slots to their places
				// (a run of 511 literals and more INSIDE one piece — one-bit literal codes — would need escape records between the
				// piece's own records: such a stream goes to k_inflate, through the flags of the round below)
				uint32_t bad = esc_in != 0u, minsrc = 0xFFFFFFFFu;
				{
					uint8_t *lp = litp + nl_tot + nlpre;
					OTZ_CHK(nl == 0u || (lp >= litp && lp + nl <= reinterpret_cast<uint8_t *>(seq_end)), OTZ_CK_SPEC_LIT);
					const uint32_t n_on = min(nl, I3_TMP_SMEM);
					for (uint32_t i = 0; i < n_on; i++) {
						lp[i] = reinterpret_cast<const uint8_t *>(&vis[(i >> 2) * G + tid])[i & 3u];
					}
					for (uint32_t i = I3_TMP_SMEM; i < nl; i++) {
						lp[i] = tml[i];
					}
					if (nm) {
						const uint32_t opos = ob_tot + obpre;
						bad |= (int64_t)need > (int64_t)opos;   // reaches before the start of the output (strict; dec:785 does not check)
						minsrc = opos - (uint32_t)need;
						uint32_t *sp = seq_end - (nseq_tot + sqpre);
						uint32_t run = carry_in + lead;
						while (run >= I2_SEQ_ESC) {
							*--sp = I2_SEQ_ESC;
							run -= I2_SEQ_ESC;
						}
						OTZ_CHK(sp - nm >= reinterpret_cast<uint32_t *>(litp + nl_tot + tot_nl) && sp <= seq_end, OTZ_CK_SPEC_SEQ);
						*--sp = (tmr[0] & ~511u) | run;
						for (uint32_t j = 1; j < nm; j++) {
							*--sp = tmr[j];
						}
					}
				}
				// dec:811-816 on the symbols of this round (positions ascend: the last one of a piece decides)
				bad |= lastp > Pend;
				const bool near_end = valid && lastp != 0u && lastp <= Pend && Pend - lastp < 8u;
#pragma unroll
				for (int d = 16; d > 0; d >>= 1) {
					minsrc = min(minsrc, __shfl_xor_sync(0xFFFFFFFFu, minsrc, d));
				}
				uint32_t flags = (__any_sync(0xFFFFFFFFu, bad != 0u) ? 1u : 0u) | (__any_sync(0xFFFFFFFFu, near_end) ? 2u : 0u);
				if (NW > 1) {
					if (lane == 0) {
						S.ws[warp][5] = flags;
						S.ws[warp][7] = minsrc;
					}
					__syncthreads();
					flags = 0;
#pragma unroll
					for (int v = 0; v < NW; v++) {
						flags |= S.ws[v][5];
						minsrc = min(minsrc, S.ws[v][7]);
					}
				}
				if (flags & 1u) {
					act = I3_A_FALLBACK;
					break;
				}
				ref_eob |= (flags >> 1) & 1u;
				sg_minsrc = min(sg_minsrc, minsrc);
				run_carry = new_carry;
				nl_tot += tot_nl;
				nseq_tot += tot_sq;
				ob_tot += tot_ob;
				P = Pnext;
				block_done = lastk == I2_K_EOB;
				// HUGE: close a segment behind the last match once it is long enough
				if (huge && (ob_tot - run_carry) - sg_out0 >= sg_target && nseg < I2_MAXSEG - 2u) {
					OTZ_CHK(nseg + 1u < I2_MAXSEG, OTZ_CK_SEG_TABLE);
					if (tid == 0) {
						I2SegRes r;
						r.nseq = nseq_tot - sg_seq0;
						r.nlit = (nl_tot - run_carry) - sg_lit0;
						r.produced = (ob_tot - run_carry) - sg_out0;
						r.end_bit = P - P0 + 1u;   // (only has to match the next segment's start)
						r.reach = sg_minsrc < sg_out0 ? sg_out0 - sg_minsrc : 0u;
						r.flags = I2_SEGF_OK;
						r.scr_lo = tok_ofs[k] + sg_lit0;
						r.scr_hi = tok_ofs[k + 1] - 4ull * sg_seq0;
						seg.res[k * I2_MAXSEG + nseg] = r;
						seg.start[k * I2_MAXSEG + nseg + 1u] = r.end_bit;
					}
					nseg++;
					sg_seq0 = nseq_tot;
					sg_lit0 = nl_tot - run_carry;
					sg_out0 = ob_tot - run_carry;
					sg_minsrc = 0xFFFFFFFFu;
				}
			}
			if (act == I3_A_FALLBACK) {
				break;
			}
			// end of block, dec:711-716 (P stands behind the code)
			if (final_blk) {
				act = I3_A_COMMIT;
			} else {
				act = I3_A_NONE;
				I3_STEP_CHECK(P);
			}
		}
#undef I3_STEP_CHECK
		if (act == I3_A_COMMIT && (P > Pend || ob_tot != cap)) {
			act = I3_A_FALLBACK;
		}
		if (tid == 0) {
			if (huge) {
				// k_seg_stitch accepts the chain or hands the stream to k_inflate
				I2SegRes r;
				r.nseq = nseq_tot - sg_seq0;
				r.nlit = nl_tot - sg_lit0;
				r.produced = ob_tot - sg_out0;
				r.end_bit = P - P0 + 1u;
				r.reach = sg_minsrc < sg_out0 ? sg_out0 - sg_minsrc : 0u;
				r.flags = act == I3_A_COMMIT ? (I2_SEGF_OK | I2_SEGF_FINAL | (ref_eob ? I2_SEGF_REF_EOB : 0u)) : 0u;
				r.scr_lo = tok_ofs[k] + sg_lit0;
				r.scr_hi = tok_ofs[k + 1] - 4ull * sg_seq0;
				seg.res[k * I2_MAXSEG + nseg] = r;
				seg.count[k] = nseg + 1u;
			} else if (act == I3_A_COMMIT) {
				I2TokRes r;
				r.nseq = nseq_tot;
				r.nlit = nl_tot;
				r.status = (rflags & OTZ_EF_CHUNK) ? OTZ_ST_OK : (OTZ_ST_OK | (ref_eob ? OTZ_STF_REF_EOB : 0));
				r.ok = 1u;
				tokres[k] = r;
			} else {
				fb_list[atomicAdd(fb_count, 1u)] = ei;
				tokres[k].ok = 0u;
			}
		}
		i3_sync<NW>();
	}
}
