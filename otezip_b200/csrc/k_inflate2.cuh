// k_inflate2.cuh — two-phase batched raw-DEFLATE (RFC 1951) decoder: one LANE per stream for the entropy
// stage, one WARP per stream for the LZ77 stage.
//
// Replaces inflateInit2/inflate/inflateEnd as driven by otezip_extract_entry
// (/root/reference/src/lib/otezip.c:503-529; decoder src/lib/deflate-dec.inc.c:547-831, "dec" below), like
// k_inflate.cuh, but splits the decoder the way the work parallelises:
//
//   phase A  k_inflate_tok   Huffman decoding is a serial bit chain per stream, so a warp that decodes one
//            stream spends 32 lanes on one symbol.  Here every lane of a warp owns a different stream: its
//            own bit reader (64-bit window in registers; the stream is staged through a per-lane shared-memory
//            ring of eight 16-byte vectors filled by cp.async three vectors ahead), its own 16-bit two-level
//            tables in shared memory (9-bit / 7-bit roots, 1796 bytes per lane, bank-skewed) and it emits, per
//            stream, the literal bytes (dense) and one 32-bit sequence record per match
//            {literal run : 9, length-3 : 8, distance-1 : 15}.  All lanes step in lock-step through straight-
//            line code — one literal/length look-up and one distance look-up per step, the step software-
//            pipelined over two iterations (parse step i, emit step i-1), everything unusual behind one warp
//            vote — so one warp instruction advances up to 28 streams.  Dynamic block headers (dec:122-266) are
//            parsed by the lanes that need one, all at the same time; the decoding tables are then built by the
//            whole warp, one lane's block at a time (code lengths in registers, canonical order by match_any
//            ranks, every table slot computed independently from the 15-bit left-aligned code boundaries).
//   phase B  k_inflate_lz    one warp per stream executes the sequences: 32 records per coalesced load, warp
//            prefix sums give every record its literal and output position; two batches are in flight — while
//            batch k executes, batch k+1 has been scanned, its match descriptors compacted into near / far
//            lists and the sources of its far matches (those that left the shared-memory ring) are on their
//            way into a staging buffer by cp.async — so executing a batch touches shared memory only; near
//            matches run ring -> ring in stream order; completed 512-byte segments leave as coalesced 16-byte
//            stores.  The same kernel executes Zstandard sequences (8-byte records, k_zstd.cuh) and the token
//            chains of segmented huge streams (below).
//
// Phase A only commits streams that are plainly valid: final block reached, exactly uncomp_size bytes, regular
// stored-block headers (their payload joins the literals), tables within the fixed budget.  Anything else (errors,
// short streams, exotic code sets) is appended to a fallback list and decoded from scratch by k_inflate, whose
// status words define the behaviour in those cases; the two kernels agree bit for bit on every stream both
// can decode, so the split is invisible to the caller.  The reference's end-of-input rule (dec:811-816,
// SURVEY.md F1) is evaluated in phase A and reported as OTZ_STF_REF_EOB exactly as k_inflate reports it.
#pragma once
#include "otz_common.cuh"
#include "k_inflate.cuh"

#define I2_LIT_ROOT 9
#define I2_DST_ROOT 7
#define I2_LIT_CAP 704   // 512 root slots + 192 second-level slots
#define I2_DST_CAP 192   // 128 root slots + 64 second-level slots
#define I2_LANES 28      // most table slots per warp (4 warps x 28 slots fit the 227 KB of one SM)
#define I2_SLOT_BYTES 1796   // (704 + 192) * 2 + 4: an odd number of 32-bit words, so equal indices of different lanes hit different banks
#define I2_LENS_OFS 0        // header parse scratch inside the lane's slot (dead once the tables are built)
#define I2_PRE_OFS 320

// 16-bit entries.  tb = bits the symbol consumes at this table level INCLUDING its extra bits, so the bit
// position advances by one field of the entry; what the extra bits mean is worked out off the critical path.
//   literal/length table: [3:0] tb, [5:4] kind, then  LIT: [13:6] byte   LEN: [8:6] extra bits, [13:9] symbol-257
//                         LINK: tb = 0, [8:6] index bits of the second-level table (0 = invalid code),
//                               [15:9] (its offset - 512) / 2   (second-level tables are even-sized and packed)
//   distance table:       [4:0] tb, [8:5] extra bits, [13:9] symbol, [15:14] kind (0 = symbol, 3 = LINK/invalid)
//                         LINK: tb = 0, [8:5] index bits (0 = invalid code), [13:9] (offset - 128) / 2
// A LINK entry has tb = 0: the straight-line decoder may add every entry's tb to the bit position blindly.
#define I2_K_LIT 0u
#define I2_K_LEN 1u
#define I2_K_EOB 2u
#define I2_K_LINK 3u
#define I2_LIT_ENTRY(tb, kind, rest) ((uint16_t)((tb) | ((kind) << 4) | ((rest) << 6)))
#define I2_LIT_LINK(bits, off) I2_LIT_ENTRY(0u, I2_K_LINK, (bits) | ((((off) - (1u << I2_LIT_ROOT)) >> 1) << 3))
#define I2_LIT_INVALID I2_LIT_ENTRY(0u, I2_K_LINK, 0u)
#define I2_LIT_LINK_BITS(e) (((e) >> 6) & 7u)
#define I2_LIT_LINK_OFS(e) ((1u << I2_LIT_ROOT) + (((e) >> 9) << 1))
#define I2_DST_ENTRY(tb, xb, sym) ((uint16_t)((tb) | ((xb) << 5) | ((sym) << 9)))
#define I2_DST_LINK(bits, off) ((uint16_t)(((bits) << 5) | ((((off) - (1u << I2_DST_ROOT)) >> 1) << 9) | (3u << 14)))
#define I2_DST_INVALID I2_DST_LINK(0u, (1u << I2_DST_ROOT))
#define I2_DST_LINK_BITS(d) (((d) >> 5) & 15u)
#define I2_DST_LINK_OFS(d) ((1u << I2_DST_ROOT) + ((((d) >> 9) & 31u) << 1))

// sequence record: [8:0] literal run, [16:9] match length - 3, [31:17] distance - 1
#define I2_SEQ_ESC 511u   // literal run of exactly 511 bytes and no match

struct I2TokRes {
	uint32_t nseq;    // sequence records written (descending from the end of the stream's scratch)
	uint32_t nlit;    // literal bytes written (ascending from the start)
	int32_t status;   // status word to report when phase B has produced the bytes
	uint32_t ok;      // 1: phase B executes this stream; 0: it went to the fallback list
};

struct I2WarpScratch {
	uint32_t ring[8 * 32 * 4];   // input staging: 8 vectors of 16 bytes per lane, [vector slot][lane][word]
	uint32_t cnt[16];
	uint32_t first15[16];   // first canonical code of each length, left-aligned to 15 bits
	uint32_t limit15[16];   // one past the last code of each length, left-aligned to 15 bits
	uint32_t offs[16];      // index in sorted[] of the first symbol of each length
	uint32_t run[16];
	uint16_t sorted[320];
	uint16_t len_base[32];    // dec:720-725 by symbol - 257
	uint16_t dist_base[32];   // dec:766-771 by symbol
};

#define I2_SMEM_BYTES(lanes) ((int)sizeof(I2WarpScratch) + (int)(lanes) * I2_SLOT_BYTES)

// worst-case scratch of a stream of n output bytes: literals + 4 bytes per match (>= 3 bytes each) + escapes
__host__ __device__ __forceinline__ uint64_t i2_scratch_bytes(uint64_t n) {
	return ((n + 4 * (n / 3 + n / 511 + 4) + 16 + 15) / 16) * 16;
}

// length / distance bases (dec:720-725, dec:766-771) from the symbol
__device__ __forceinline__ uint32_t i2_len_base(uint32_t v, uint32_t xb) {
	return v < 8u ? 3u + v : v == 28u ? 258u : 3u + ((4u + (v & 3u)) << xb);
}
__device__ __forceinline__ uint32_t i2_dist_base(uint32_t ds, uint32_t xb) { return ds < 4u ? 1u + ds : 1u + ((2u + (ds & 1u)) << xb); }

// ------------------------------------------------------------------------------------------------
// per-lane bit reader: {lo,hi} is a 64-bit window of the stream, pos < 32 after norm(); nx is the word after hi.
// The stream is staged through shared memory as 16-byte vectors, eight slots per lane, requested with cp.async
// three vectors (>= 8 decoding steps) ahead of their first use: in lock-step execution no lane's cache miss stalls
// the other 31, and a refill is branch-free (two selects and one shared-memory load).
struct I2Reader {
	const uint4 *b16;        // 16-byte aligned base of the stream
	uint32_t nvec;           // vectors that overlap the stream
	uint32_t wi;             // word index (from b16) of nx
	uint32_t lo, hi, nx, pos;
	uint32_t wi_end;         // words_left() = wi_end - wi
	uint32_t pad_bits;
	uint32_t fu;             // vectors [0, fu) have been requested
	volatile uint32_t *col;  // this lane's column of the staging ring (written by cp.async)
	uint32_t col_sa;         // its shared-window address

	// 32-bit words of the stream not yet moved into {lo,hi} (as BitReader::words_left in k_inflate.cuh)
	__device__ __forceinline__ int32_t words_left() const { return (int32_t)(wi_end - wi); }
	__device__ __forceinline__ uint32_t word(uint32_t i) const { return col[((i & 28u) << 5) | (i & 3u)]; }
	// Keep vectors up to (wi >> 2) + 3 requested; at most one new vector per call, which is enough for one decoding
	// step (<= 48 bits).  Straight-line: the copy is predicated, not branched around.
	__device__ __forceinline__ void top() {
		const uint32_t pred = fu <= (wi >> 2) + 3u;
		const uint32_t cc = fu < nvec ? fu : nvec;   // never more than one vector past the stream
		const uint32_t sa = col_sa + ((fu & 7u) << 9);
		// one (possibly empty) copy group per call: "at most 6 groups pending" then means that everything requested
		// seven or more steps ago has landed — a vector is requested at least 8 steps before its first word is read
		asm volatile(
			"{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n\t"
			"cp.async.commit_group;\n\tcp.async.wait_group 6;" ::"r"(sa),
			"l"(b16 + cc), "r"(pred));
		fu += pred;
	}
	__device__ __forceinline__ void init(const uint8_t *p, uint64_t nbytes) {
		const uint64_t a = reinterpret_cast<uint64_t>(p);
		const uint32_t skipb = (uint32_t)(a & 3), i0 = (uint32_t)(a & 15) >> 2;
		b16 = reinterpret_cast<const uint4 *>(a & ~15ull);
		const uint32_t nw = (uint32_t)((skipb + nbytes + 3) >> 2);
		nvec = (i0 + nw + 3) >> 2;
		pad_bits = (uint32_t)(((uint64_t)nw << 5) - ((skipb + nbytes) << 3));
		// (a restart in the middle of a stream — behind a stored block, at a chunk — must not race with copies that are
		// still on their way into the same slots)
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		for (fu = 0; fu < 4u; fu++) {
			const uint32_t cc = fu < nvec ? fu : nvec;
			const uint32_t sa = col_sa + ((fu & 7u) << 9);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" ::"r"(sa), "l"(b16 + cc) : "memory");
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		lo = word(i0);
		hi = word(i0 + 1);
		wi = i0 + 2;
		nx = word(wi);
		wi_end = wi + nw - 2u;
		pos = 8 * skipb;
	}
	// pos < 64 on entry; the caller keeps the staging ring ahead with top()
	__device__ __forceinline__ void norm() {
		const bool p = pos >= 32u;
		lo = p ? hi : lo;
		hi = p ? nx : hi;
		pos &= 31u;
		wi += p;
		nx = word(wi);
	}
	// for the (branchy, rare) header code: refill check + norm
	__device__ __forceinline__ void norm_hdr() {
		top();
		norm();
	}
	__device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, pos); }
	__device__ __forceinline__ int64_t remaining_bits() const { return ((int64_t)words_left() << 5) + 64 - (int64_t)pos - (int64_t)pad_bits; }
};

// ------------------------------------------------------------------------------------------------
// Warp-cooperative build of one two-level table from code lengths held in registers.
// lens[j] = code length of combined symbol 32*j + lane (literal/length alphabet first, then distances);
// this call covers combined symbols [b0, b0 + n).  Returns 0 ok, 1 = not usable on the fast path
// (invalid / incomplete set in a shape k_inflate has to judge, or second-level budget exceeded).
template <bool IS_DIST>
__device__ __forceinline__ uint16_t i2_symbol_entry(uint32_t s, uint32_t cb) {
	if (IS_DIST) {
		const uint32_t xb = s < 4u ? 0u : (s - 2u) >> 1;
		return s < 30u ? I2_DST_ENTRY(cb + xb, xb, s) : I2_DST_INVALID;
	}
	if (s < 256u) {
		return I2_LIT_ENTRY(cb, I2_K_LIT, s);
	}
	if (s == 256u) {
		return I2_LIT_ENTRY(cb, I2_K_EOB, 0u);
	}
	const uint32_t v = s - 257u;
	const uint32_t xb = (v < 8u || v == 28u) ? 0u : (v - 4u) >> 2;
	return s < 286u ? I2_LIT_ENTRY(cb + xb, I2_K_LEN, xb | (v << 3)) : I2_LIT_INVALID;
}

// length and position in sorted[] of the code that covers the 15-bit left-aligned value c15 (len 16 = none);
// lim[j] = S.limit15[j], held in registers by the caller
template <typename WS>
__device__ __forceinline__ void i2_lookup15(const WS &S, const uint32_t (&lim)[16], uint32_t c15, uint32_t &len, uint32_t &idx) {
	uint32_t l = 1;
#pragma unroll
	for (int j = 1; j <= 15; j++) {
		l += (c15 >= lim[j]);
	}
	len = l;
	const uint32_t ll = l > 15u ? 15u : l;
	idx = S.offs[ll] + ((c15 - S.first15[ll]) >> (15u - ll));
}

template <bool IS_DIST, int ROOT, int CAP, typename WS>
__device__ __noinline__ int i2_build_table(WS &S, const uint32_t (&lens)[10], uint32_t b0, uint32_t n, uint16_t *tbl) {
	const uint32_t lane = threadIdx.x & 31u;
	constexpr uint16_t INVALID = IS_DIST ? I2_DST_INVALID : I2_LIT_INVALID;
	if (lane < 16) {
		S.cnt[lane] = 0;
	}
	__syncwarp();
	uint32_t mylen[10];
#pragma unroll
	for (int j = 0; j < 10; j++) {
		const uint32_t s = 32u * j + lane - b0;
		mylen[j] = s < n ? lens[j] : 0u;
		if (mylen[j]) {
			atomicAdd(&S.cnt[mylen[j]], 1u);
		}
	}
	__syncwarp();
	int left = 1;
	uint32_t ncodes = 0;
	{
		uint32_t code = 0, off = 0;
		for (uint32_t l = 1; l <= 15; l++) {
			const uint32_t c = S.cnt[l];
			left = (left << 1) - (int)c;
			if (left < 0) {
				return 1;   // over-subscribed
			}
			code = (code + (l > 1 ? S.cnt[l - 1] : 0u)) << 1;
			if (lane == 0) {
				S.offs[l] = off;
				S.run[l] = off;
				S.first15[l] = code << (15u - l);
				S.limit15[l] = (code + c) << (15u - l);
			}
			off += c;
			ncodes += c;
		}
	}
	constexpr uint32_t ROOTSZ = 1u << ROOT;
	if (ncodes == 0) {
		if (!IS_DIST) {
			return 1;
		}
		for (uint32_t k = lane; k < ROOTSZ; k += 32) {
			tbl[k] = INVALID;   // a block of literals only: any distance code is an error
		}
		__syncwarp();
		return 0;
	}
	if (left > 0 && !(ncodes == 1 && S.cnt[1] == 1)) {
		return 1;   // incomplete set (zlib accepts only a single 1-bit code): k_inflate reports it
	}
	__syncwarp();
	// canonical order: sorted[] = symbols by (length, symbol)
#pragma unroll
	for (int j = 0; j < 10; j++) {
		const uint32_t l = mylen[j];
		const uint32_t m = __match_any_sync(0xFFFFFFFFu, l);
		const uint32_t rank = __popc(m & ((1u << lane) - 1u));
		if (l) {
			S.sorted[S.run[l] + rank] = (uint16_t)(32u * j + lane - b0);
		}
		__syncwarp();
		if (l && rank == 0) {
			S.run[l] += __popc(m);
		}
		__syncwarp();
	}
	// root slots: slot k holds the code whose bits, LSB first, are a prefix of k
	uint32_t lim[16];
#pragma unroll
	for (int j = 0; j < 16; j++) {
		lim[j] = S.limit15[j];
	}
	const uint32_t long15 = S.limit15[ROOT];          // first 15-bit value whose code is longer than ROOT bits
	const uint32_t end15 = S.limit15[15];             // one past the last covered value (32768 when complete)
	for (uint32_t k = lane; k < ROOTSZ; k += 32) {
		const uint32_t c15 = (__brev(k) >> (32 - ROOT)) << (15 - ROOT);
		uint16_t e = INVALID;
		if (c15 < long15) {
			uint32_t len, idx;
			i2_lookup15(S, lim, c15, len, idx);
			e = i2_symbol_entry<IS_DIST>(S.sorted[idx], len);
		}
		tbl[k] = e;
	}
	// prefixes of codes longer than ROOT bits: one second-level table each, sized by its longest code
	if (end15 > long15) {
		const uint32_t p0 = long15 >> (15 - ROOT), p1 = (end15 + (1u << (15 - ROOT)) - 1u) >> (15 - ROOT);
		uint32_t next_free = ROOTSZ;
		for (uint32_t pb = p0; pb < p1; pb += 32) {
			const uint32_t p = pb + lane;
			uint32_t sub_bits = 0;
			if (p < p1) {
				uint32_t v15 = ((p + 1u) << (15 - ROOT)) - 1u;
				v15 = v15 < end15 ? v15 : end15 - 1u;
				uint32_t len, idx;
				i2_lookup15(S, lim, v15, len, idx);
				sub_bits = len - ROOT;
			}
			const uint32_t size = p < p1 ? (1u << sub_bits) : 0u;
			uint32_t incl = size;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
				if ((int)lane >= d) {
					incl += t;
				}
			}
			const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
			if (next_free + total > (uint32_t)CAP) {
				return 1;   // second-level budget exceeded: k_inflate (worst-case tables) takes the stream
			}
			const uint32_t sub_off = next_free + incl - size;
			next_free += total;
			if (p < p1) {
				tbl[__brev(p) >> (32 - ROOT)] = IS_DIST ? I2_DST_LINK(sub_bits, sub_off) : I2_LIT_LINK(sub_bits, sub_off);
				for (uint32_t t = 0; t < size; t++) {
					const uint32_t c15 = (p << (15 - ROOT)) | ((__brev(t) >> (32u - sub_bits)) << (15u - ROOT - sub_bits));
					uint16_t e = INVALID;
					if (c15 < end15) {
						uint32_t len, idx;
						i2_lookup15(S, lim, c15, len, idx);
						e = i2_symbol_entry<IS_DIST>(S.sorted[idx], len - ROOT);
					}
					tbl[sub_off + t] = e;
				}
			}
		}
	}
	__syncwarp();
	return 0;
}

// ------------------------------------------------------------------------------------------------
// Segmented decode of HUGE streams.  A stream advances one symbol per step wherever it is decoded, so one 16 MiB entry
// (1.1 M symbols) sets the critical path of a whole batch.  Its entropy stage, however, only needs to know where
// blocks start: k_block_search tests every bit offset of the stream for a plausible dynamic-block header (the
// checks are so selective that false positives are practically absent), each candidate becomes a SEGMENT that a lane
// of k_inflate_tok<true> decodes from its candidate to the next live one, and k_seg_stitch accepts the stream only if
// the segments chain exactly (each one ends where the next starts, the last one with the final block, the sizes add
// up, no match reaches before the stream).  Tokens do not care where they were produced: k_inflate_lz walks the
// chain.  Nothing depends on the search being right — a wrong candidate is skipped (the decoder that passes it marks
// it dead) or fails the chain, and the stream then goes to k_inflate like every other declined stream.
#define I2_MAXSEG 256
#define I2_PREWIN 32768u   // elements of history a segment may refer to (DEFLATE window)
#define I2_SEGF_OK 1u
#define I2_SEGF_FINAL 2u
#define I2_SEGF_DEAD 4u
#define I2_SEGF_REF_EOB 8u

struct I2SegRes {
	uint32_t nseq, nlit;    // tokens of the segment
	uint32_t produced;      // output bytes of the segment
	uint32_t end_bit;       // bit position (in the stream) where the segment stopped
	uint32_t reach;         // how far a match reaches before the segment's first output byte
	uint32_t flags;         // I2_SEGF_*
	uint64_t scr_lo, scr_hi;   // its slice of the stream's token scratch (literals up from lo, records down from hi)
};

struct I2SegCtl {
	uint32_t *count;     // [n_huge] candidates found / segments of the stream
	uint32_t *start;     // [n_huge][I2_MAXSEG] start bit of every segment, ascending, [0] = 0
	uint32_t *bucket;    // [n_huge][I2_MAXSEG] first candidate in every 1/255 of the stream (0xFFFFFFFF = none)
	I2SegRes *res;       // [n_huge][I2_MAXSEG]
	uint32_t *items;     // compact work list: stream << 16 | segment
	uint32_t *n_items;
	uint32_t *live;      // [n_huge][I2_MAXSEG] the chain of segments that make up the stream
	uint32_t *nlive;     // [n_huge] length of the chain (0 = the stream was declined)
	int32_t *seg_status; // [n_huge] status word of an accepted stream
	// parallel execution of the chain (k_inflate_lz<.., PAR> + k_seg_window + k_seg_translate)
	uint32_t *par;       // [n_huge] 1 = the segments of the stream are executed in parallel into the symbol buffer
	uint32_t *par_items; // work list: stream << 16 | position in the chain
	uint32_t *n_par;
	uint64_t *sym_start; // [n_huge][I2_MAXSEG] by chain position: first element of the segment in the symbol buffer
	uint32_t *out_start; // [n_huge][I2_MAXSEG] by chain position: first output byte of the segment in the entry
	unsigned long long *sym_top;   // bump allocator of the symbol buffer (elements)
	uint64_t sym_cap;    // its capacity (0 = no parallel execution)
	uint32_t out_mis;    // address of the output arena & 15
};

// ------------------------------------------------------------------------------------------------
// lane states
#define I2_S_IDLE 0u     // needs a stream
#define I2_S_HDR 1u      // at a block header
#define I2_S_BUILD 2u    // code lengths parsed into the slot; waiting for the cooperative table build
#define I2_S_DEC 3u      // decoding symbols
#define I2_S_DONE 4u     // no more work
#define I2_S_COPY 5u     // at the payload of a stored block; waiting for the cooperative copy

__device__ __forceinline__ uint32_t i2_ld_le16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// what a lane has to do after i2_header()
#define I2_A_NONE 0u       // still at a block header (an empty stored / fixed block was consumed)
#define I2_A_BUILD 1u      // code lengths are in the slot: build the tables, then decode
#define I2_A_COMMIT 2u     // the stream (or chunk) ended here
#define I2_A_FALLBACK 3u   // k_inflate takes the stream
#define I2_A_STORED 4u     // a stored block with payload: st_src / st_len say where it is

// dec:811-816 as k_inflate evaluates it: after a step that leaves the stream unfinished
#define I2_HDR_STEP_CHECK()                               \
	if (br.words_left() <= 1) {                           \
		const int64_t rem_ = br.remaining_bits();         \
		if (rem_ < 0) {                                   \
			return I2_A_FALLBACK;                         \
		} else if (rem_ < 8) {                            \
			ref_eob = 1u;                                 \
		}                                                 \
	}

// One block header (dec:613-627) of the lane's stream: stored blocks (dec:269-319; the payload of a non-empty one is
// copied by the whole warp afterwards), fixed (dec:322-349) and dynamic (dec:122-266) code lengths into the lane's
// slot.  Lanes run this in lock-step.
__device__ __forceinline__ uint32_t i2_header(I2Reader &br, uint8_t *slot, const uint8_t *in, uint32_t comp, uint32_t rflags, uint32_t &final_blk,
	uint32_t &ref_eob, uint32_t &hlit, uint32_t &hdist, uint32_t &st_src, uint32_t &st_len) {
	const bool chunk_mid = (rflags & OTZ_EF_CHUNK) && !(rflags & OTZ_EF_LAST_CHUNK);
	br.norm_hdr();
	uint32_t bits = br.peek();
	final_blk = bits & 1u;
	const uint32_t btype = (bits >> 1) & 3u;
	br.pos += 3;
	I2_HDR_STEP_CHECK();
	if (chunk_mid && final_blk) {
		return I2_A_FALLBACK;   // a final block inside a chunk that is not the last: k_inflate fails the chunk
	}
	if (btype == 0) {
		// stored block (dec:269-319).  Anything irregular — a bad length pair, a payload that runs past the input — is
		// left to k_inflate, which knows what the reference answers
		const int64_t rem = br.remaining_bits();
		const uint64_t bpos = (uint64_t)comp - (uint64_t)(rem >> 3);
		if ((uint64_t)comp - bpos < 4 || (i2_ld_le16(in + bpos) ^ i2_ld_le16(in + bpos + 2)) != 0xFFFFu) {
			return I2_A_FALLBACK;
		} else if (i2_ld_le16(in + bpos) != 0u) {
			st_len = i2_ld_le16(in + bpos);
			st_src = (uint32_t)(bpos + 4);
			if ((uint64_t)comp - (bpos + 4) < st_len) {
				return I2_A_FALLBACK;
			} else {
				return I2_A_STORED;
			}
		} else {
			const uint64_t npos = bpos + 4;
			br.init(in + npos, (uint64_t)comp - npos);
			if (final_blk) {
				return I2_A_COMMIT;
			} else if (npos >= comp) {
				if (chunk_mid) {
					return I2_A_COMMIT;   // end of this chunk
				} else {
					return I2_A_FALLBACK;   // unfinished stream out of input
				}
			}
		}
	} else if (btype == 3) {
		return I2_A_FALLBACK;   // dec:657-658
	} else if (btype == 1) {
		br.norm_hdr();
		if ((br.peek() & 127u) == 0u) {
			// empty fixed block (zlib's Z_FINISH tail): end-of-block is the 7-bit code 0000000
			br.pos += 7;
			if (final_blk) {
				return I2_A_COMMIT;
			} else {
				I2_HDR_STEP_CHECK();
			}
		} else {
			uint8_t *lens = slot + I2_LENS_OFS;   // dec:322-349
			for (int i = 0; i < 320; i++) {
				lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
			}
			hlit = 288;
			hdist = 32;
			return I2_A_BUILD;
		}
	} else {
		// dynamic block header, dec:122-266
		uint8_t *lens = slot + I2_LENS_OFS;
		uint8_t *pre = slot + I2_PRE_OFS;
		br.norm_hdr();
		bits = br.peek();
		hlit = (bits & 31u) + 257u;
		hdist = ((bits >> 5) & 31u) + 1u;
		const uint32_t hclen = ((bits >> 10) & 15u) + 4u;
		br.pos += 14;
		bool bad = hlit > 286u || hdist > 30u;
		uint64_t cl = 0;    // 19 code-length-code lengths, 3 bits each
		uint64_t cnt = 0;   // packed byte counters per length
		for (uint32_t i = 0; i < hclen; i++) {
			br.norm_hdr();
			const uint32_t v = br.peek() & 7u;
			br.pos += 3;
			cl |= (uint64_t)v << (3u * c_cl_order[i]);
			cnt += 1ull << (8u * v);
		}
		uint64_t next = 0;   // packed next canonical code per length
		{
			int left = 1;
			uint32_t code = 0;
			for (uint32_t l = 1; l <= 7; l++) {
				const uint32_t c = (uint32_t)(cnt >> (8u * l)) & 0xFFu;
				left = (left << 1) - (int)c;
				bad |= left < 0;
				code = (code + (l > 1 ? (uint32_t)(cnt >> (8u * (l - 1))) & 0xFFu : 0u)) << 1;
				next |= (uint64_t)(code & 0xFFu) << (8u * l);
			}
			bad |= left != 0;   // the code-length code must be complete
		}
		if (!bad) {
			for (uint32_t s = 0; s < 19; s++) {
				const uint32_t l = (uint32_t)(cl >> (3u * s)) & 7u;
				if (l) {
					const uint32_t c = (uint32_t)(next >> (8u * l)) & 0xFFu;
					next += 1ull << (8u * l);
					const uint32_t rev = __brev(c) >> (32u - l);
					for (uint32_t x = rev; x < 128u; x += (1u << l)) {
						pre[x] = (uint8_t)(s | (l << 5));
					}
				}
			}
			const uint32_t total = hlit + hdist;
			uint32_t idx = 0, prev = 0;
			while (idx < total) {
				br.norm_hdr();
				bits = br.peek();
				const uint32_t e = pre[bits & 127u];
				const uint32_t sym = e & 31u, cb = e >> 5;
				if (sym < 16u) {
					br.pos += cb;
					lens[idx++] = (uint8_t)sym;
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
					br.pos += cb + 2;
				} else if (sym == 17u) {   // dec:221-228
					rep = 3u + ((bits >> cb) & 7u);
					br.pos += cb + 3;
				} else {   // dec:230-237
					rep = 11u + ((bits >> cb) & 127u);
					br.pos += cb + 7;
				}
				if (idx + rep > total) {
					bad = true;   // dec:244
					break;
				}
				for (uint32_t i = 0; i < rep; i++) {
					lens[idx + i] = (uint8_t)val;
				}
				idx += rep;
				prev = val;
			}
			for (uint32_t i = total; i < 320u; i++) {
				lens[i] = 0;
			}
			bad |= !bad && lens[256] == 0;   // no end-of-block code
		}
		if (bad || br.remaining_bits() < 0) {
			return I2_A_FALLBACK;
		} else {
			return I2_A_BUILD;
		}
	}
		return I2_A_NONE;
}
#undef I2_HDR_STEP_CHECK

// phase A.  grid: persistent, one warp per CTA; the first `lanes_active` (<= I2_LANES) lanes of every warp pull list
// indices from *work_counter.  (Few streams are spread over all resident warps rather than packed into few:
// a lock-step step costs the same whatever the number of live lanes.)
// SEG: the work items are the segments of huge streams (`list`, `tok_ofs` index those streams; see I2SegCtl).
template <bool SEG>
__global__ void __launch_bounds__(32) k_inflate_tok(const uint8_t *__restrict__ archive, const otz_entry *__restrict__ ents,
	const OtzEntryState *__restrict__ est, const int32_t *__restrict__ status, const uint32_t *__restrict__ list, uint32_t n_list,
	uint32_t *__restrict__ work_counter, uint8_t *__restrict__ scratch, const uint64_t *__restrict__ tok_ofs, I2TokRes *__restrict__ tokres,
	uint32_t *__restrict__ fb_list, uint32_t *__restrict__ fb_count, uint32_t lanes_active, I2SegCtl seg) {
	if (SEG) {
		n_list = *seg.n_items;
	}
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const uint32_t lane = threadIdx.x;
	// shared memory: the warp scratch, then `lanes_active` table slots.  Lanes without a slot never become active;
	// their (predicated-off) table reads go to slot 0
	I2WarpScratch &WS = *reinterpret_cast<I2WarpScratch *>(smem_raw);
	uint8_t *const slots = smem_raw + sizeof(I2WarpScratch);
	uint8_t *const slot = slots + (lane < lanes_active ? lane : 0u) * I2_SLOT_BYTES;
	const uint16_t *const lit = reinterpret_cast<const uint16_t *>(slot);
	const uint16_t *const dst = lit + I2_LIT_CAP;

	{
		const uint32_t xl = (lane < 8u || lane == 28u) ? 0u : (lane - 4u) >> 2, xd = lane < 4u ? 0u : (lane - 2u) >> 1;
		WS.len_base[lane] = (uint16_t)(lane < 29u ? i2_len_base(lane, xl) : 0u);
		WS.dist_base[lane] = (uint16_t)(lane < 30u ? i2_dist_base(lane, xd) : 0u);
	}
	__syncwarp();
	uint32_t state = lane < lanes_active ? I2_S_IDLE : I2_S_DONE;
	I2Reader br;
	// lanes that never get a stream still run the (predicated-off) refill logic: give them a harmless source
	br.b16 = reinterpret_cast<const uint4 *>(reinterpret_cast<uint64_t>(archive) & ~15ull);
	br.nvec = br.wi = br.lo = br.hi = br.nx = br.pos = br.pad_bits = br.wi_end = 0;
	br.fu = 8;   // nothing to request
	br.col = WS.ring + lane * 4u;
	br.col_sa = (uint32_t)__cvta_generic_to_shared(WS.ring + lane * 4u);
	const uint8_t *in = nullptr;
	uint32_t comp = 0, cap = 0, rflags = 0, k = 0, ei = 0;
	// output bytes so far = nl + mb; the current literal run = nl - nl0
	uint32_t nl = 0, nl0 = 0, mb = 0, nseq = 0;
	uint8_t *litp = nullptr;
	uint32_t *seqp = nullptr;
	uint32_t final_blk = 0, ref_eob = 0, hlit = 0, hdist = 0;
	uint32_t st_src = 0, st_len = 0;   // payload of the stored block the lane stands at (offset in the stream, bytes)
	// step parsed but not yet emitted (software pipeline of the decode loop): entries and bit windows
	bool pv = false;
	uint32_t pe = 0, pd = 0, pb = 0, pb2 = 0;
	// SEG: stream / segment of this lane, the next candidate start, bytes a match reached before the segment
	uint32_t sh = 0, sj = 0, sjn = 0, s_next = 0xFFFFFFFFu, s_own = 0, reach = 0;
	uint8_t *seq_floor = nullptr;   // SEG: the records must stay above the literals (scr_lo .. scr_hi is a guess)

// leave the current stream: FALLBACK = hand it to k_inflate, COMMIT = release it to phase B
// (SEG: the segment is left without the OK flag / with its result; k_seg_stitch decides about the stream)
#define I2_FALLBACK()                                                     \
	do {                                                                  \
		if (!SEG) {                                                       \
			fb_list[atomicAdd(fb_count, 1u)] = ei;                        \
			tokres[k].ok = 0u;                                            \
		}                                                                 \
		state = I2_S_IDLE;                                                \
	} while (0)
#define I2_SEG_COMMIT(final_)                                                               \
	do {                                                                                    \
		const int64_t rem_c = br.remaining_bits();                                          \
		if (rem_c >= 0) {                                                                   \
			I2SegRes *r_ = &seg.res[sh * I2_MAXSEG + sj];                                   \
			r_->nseq = nseq;                                                                \
			r_->nlit = nl;                                                                  \
			r_->produced = nl + mb;                                                         \
			r_->end_bit = (uint32_t)((int64_t)comp * 8 - rem_c);                            \
			r_->reach = reach;                                                              \
			atomicOr(&r_->flags, I2_SEGF_OK | ((final_) ? I2_SEGF_FINAL : 0u) | (ref_eob ? I2_SEGF_REF_EOB : 0u)); \
		}                                                                                   \
		state = I2_S_IDLE;                                                                  \
	} while (0)
#define I2_COMMIT()                                                                         \
	do {                                                                                    \
		if (SEG) {                                                                          \
			I2_SEG_COMMIT(true);                                                            \
		} else if (br.remaining_bits() < 0 || nl + mb != cap) {                             \
			I2_FALLBACK();                                                                  \
		} else {                                                                            \
			I2TokRes r;                                                                     \
			r.nseq = nseq;                                                                  \
			r.nlit = nl;                                                                    \
			r.status = (rflags & OTZ_EF_CHUNK) ? OTZ_ST_OK : (OTZ_ST_OK | (ref_eob ? OTZ_STF_REF_EOB : 0)); \
			r.ok = 1u;                                                                      \
			tokres[k] = r;                                                                  \
			state = I2_S_IDLE;                                                              \
		}                                                                                   \
	} while (0)
// dec:811-816 as k_inflate evaluates it: after a step that leaves the stream unfinished
#define I2_STEP_CHECK()                                   \
	if (br.words_left() <= 1) {                             \
		const int64_t rem_ = br.remaining_bits();         \
		if (rem_ < 0) {                                   \
			I2_FALLBACK();                                \
		} else if (rem_ < 8) {                            \
			ref_eob = 1u;                                 \
		}                                                 \
	}
// one literal byte / one match into the stream's token scratch
#define I2_EMIT_LIT(byte_)                      \
	do {                                        \
		litp[nl] = (uint8_t)(byte_);            \
		nl++;                                   \
	} while (0)
#define I2_EMIT_MATCH(run_, len_, dist_)                                              \
	do {                                                                              \
		*--seqp = (run_) | (((len_) - 3u) << 9) | (((dist_) - 1u) << 17);             \
		nseq++;                                                                       \
		nl0 = nl;                                                                     \
		mb += (len_);                                                                 \
	} while (0)

	for (;;) {
		// ---- (1) hand streams to idle lanes
		{
			const uint32_t idle = __ballot_sync(0xFFFFFFFFu, state == I2_S_IDLE);
			if (idle) {
				uint32_t base = 0;
				if (lane == (uint32_t)(__ffs(idle) - 1)) {
					base = atomicAdd(work_counter, (uint32_t)__popc(idle));
				}
				base = __shfl_sync(0xFFFFFFFFu, base, __ffs(idle) - 1);
				if (state == I2_S_IDLE) {
					k = base + __popc(idle & ((1u << lane) - 1u));
					if (k >= n_list) {
						state = I2_S_DONE;
					} else if (SEG) {
						const uint32_t item = seg.items[k];
						sh = item >> 16;
						sj = item & 0xFFFFu;
						ei = list[sh];
						const otz_entry e = ents[ei];
						in = archive + est[ei].data_ofs;
						comp = e.comp_size;
						cap = e.uncomp_size;
						rflags = e.flags;
						const uint32_t cnt = seg.count[sh];
						s_own = seg.start[sh * I2_MAXSEG + sj];
						sjn = sj + 1;
						s_next = sjn < cnt ? seg.start[sh * I2_MAXSEG + sjn] : 0xFFFFFFFFu;
						I2SegRes *r_ = &seg.res[sh * I2_MAXSEG + sj];
						litp = scratch + r_->scr_lo;
						seq_floor = litp;
						seqp = reinterpret_cast<uint32_t *>(scratch + r_->scr_hi);
						nl = nl0 = mb = nseq = 0;
						ref_eob = 0;
						reach = 0;
						pv = false;
						if (r_->scr_hi - r_->scr_lo >= 128u) {
							br.init(in + (s_own >> 3), (uint64_t)comp - (s_own >> 3));
							br.pos += s_own & 7u;
							state = I2_S_HDR;
						}   // else: two candidates a few bits apart, no room for tokens — the segment simply fails (stays idle)
					} else {
						ei = list[k];
						if (OTZ_ST_CODE(status[ei]) == OTZ_ST_OK) {
							const otz_entry e = ents[ei];
							in = archive + est[ei].data_ofs;
							comp = e.comp_size;
							cap = e.uncomp_size;
							rflags = e.flags;
							litp = scratch + tok_ofs[k];
							seqp = reinterpret_cast<uint32_t *>(scratch + tok_ofs[k + 1]);
							nl = nl0 = mb = nseq = 0;
							ref_eob = 0;
							if (comp == 0) {
								I2_FALLBACK();   // dec:610: k_inflate answers TRUNCATED
							} else {
								br.init(in, comp);
								state = I2_S_HDR;
							}
						} else {
							tokres[k].ok = 0u;   // failed in k_resolve: nothing to decode (stays idle, picks the next one)
						}
					}
				}
			}
			if (__all_sync(0xFFFFFFFFu, state == I2_S_DONE)) {
				break;
			}
		}
		// ---- (2) block headers (dec:613-627), every lane that stands at one, in lock-step
		if (SEG && state == I2_S_HDR) {
			// a block boundary: the segment ends where the next live candidate starts; candidates it has run past
			// were false (or inside a block) and are marked dead
			const uint32_t P = (uint32_t)((int64_t)comp * 8 - br.remaining_bits());
			if (P != s_own) {
				const uint32_t cnt = seg.count[sh];
				while (P > s_next) {
					atomicOr(&seg.res[sh * I2_MAXSEG + sjn].flags, I2_SEGF_DEAD);
					sjn++;
					s_next = sjn < cnt ? seg.start[sh * I2_MAXSEG + sjn] : 0xFFFFFFFFu;
				}
				if (P == s_next) {
					I2_SEG_COMMIT(false);
				}
			}
		}
		if (state == I2_S_HDR) {
			const uint32_t act_ = i2_header(br, slot, in, comp, rflags, final_blk, ref_eob, hlit, hdist, st_src, st_len);
			if (act_ == I2_A_BUILD) {
				state = I2_S_BUILD;
			} else if (act_ == I2_A_STORED) {
				state = I2_S_COPY;
			} else if (act_ == I2_A_COMMIT) {
				I2_COMMIT();
			} else if (act_ == I2_A_FALLBACK) {
				I2_FALLBACK();
			}
		}
		// ---- (2b) tables, one requesting lane at a time, whole warp
		{
			uint32_t need = __ballot_sync(0xFFFFFFFFu, state == I2_S_BUILD);
			while (need) {
				const int x = __ffs(need) - 1;
				need &= need - 1;
				__syncwarp();
				const uint8_t *xl = slots + x * I2_SLOT_BYTES + I2_LENS_OFS;
				uint32_t lens[10];
#pragma unroll
				for (int j = 0; j < 10; j++) {
					lens[j] = xl[32 * j + lane];
				}
				__syncwarp();
				const uint32_t xh = __shfl_sync(0xFFFFFFFFu, hlit, x), xd = __shfl_sync(0xFFFFFFFFu, hdist, x);
				uint16_t *xt = reinterpret_cast<uint16_t *>(slots + x * I2_SLOT_BYTES);
				int r = i2_build_table<false, I2_LIT_ROOT, I2_LIT_CAP>(WS, lens, 0u, xh, xt);
				if (!r) {
					r = i2_build_table<true, I2_DST_ROOT, I2_DST_CAP>(WS, lens, xh, xd, xt + I2_LIT_CAP);
				}
				__syncwarp();
				if ((int)lane == x) {
					if (r) {
						I2_FALLBACK();
					} else {
						state = I2_S_DEC;
						I2_STEP_CHECK();
					}
				}
			}
		}
		// ---- (2c) payloads of stored blocks (dec:269-319), one requesting lane at a time, whole warp: the bytes join the
		// literals of the stream (a run of more than 511 literals becomes escape records when the next match is emitted)
		{
			uint32_t need = __ballot_sync(0xFFFFFFFFu, state == I2_S_COPY);
			while (need) {
				const int x = __ffs(need) - 1;
				need &= need - 1;
				bool room = nl + mb + st_len <= cap;   // (else k_inflate reports the overflow, dec:296-300)
				if (SEG) {
					room = room && (int64_t)((uint8_t *)seqp - (seq_floor + nl)) >= (int64_t)st_len + (int64_t)(4u * ((nl - nl0 + st_len) / I2_SEQ_ESC)) + 64;
				}
				room = __shfl_sync(0xFFFFFFFFu, (int)room, x) != 0;
				const uint32_t xlen = __shfl_sync(0xFFFFFFFFu, st_len, x);
				const uint8_t *sp = reinterpret_cast<const uint8_t *>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)reinterpret_cast<uint64_t>(in + st_src), x));
				uint8_t *dp = reinterpret_cast<uint8_t *>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)reinterpret_cast<uint64_t>(litp + nl), x));
				if (room) {
#pragma unroll 4
					for (uint32_t i = lane; i < xlen; i += 32) {
						dp[i] = sp[i];
					}
				}
				__syncwarp();
				if ((int)lane == x) {
					if (!room) {
						I2_FALLBACK();
					} else {
						nl += st_len;
						const uint64_t npos = (uint64_t)st_src + st_len;
						const bool chunk_mid = (rflags & OTZ_EF_CHUNK) && !(rflags & OTZ_EF_LAST_CHUNK);
						br.init(in + npos, (uint64_t)comp - npos);
						if (final_blk) {
							I2_COMMIT();
						} else if (npos >= comp) {
							if (chunk_mid) {
								I2_COMMIT();   // end of this chunk
							} else {
								I2_FALLBACK();   // unfinished stream out of input
							}
						} else {
							state = I2_S_HDR;
						}
					}
				}
			}
		}
		// ---- (3) symbols (dec:662-799).  One literal/length symbol and — used by the lanes that got a length —
		// one distance symbol per step, as straight-line code: every lane runs the same instructions.  A step is
		// software-pipelined over two iterations: iteration i PARSES step i (bit position, two table lookups — the
		// serial chain of the stream) and EMITS step i-1 (length/distance arithmetic, bounds, token stores) from the
		// entries and bit windows saved in registers, so the two dependency chains interleave in the one warp.
		// Everything unusual (second-level tables, end of block, long literal runs, errors, end of input) is left
		// untouched by the main path and finished, lane by lane, behind ONE warp vote at the end of the iteration.
		while (__all_sync(0xFFFFFFFFu, state == I2_S_DEC || state == I2_S_DONE) && __any_sync(0xFFFFFFFFu, state == I2_S_DEC)) {
			bool leave = false;
#pragma unroll 1
			for (int burst = 0; burst < 64 && !leave; burst++) {
				const bool act = state == I2_S_DEC;
				// ---- emit step i-1
				bool e_special;
				{
					const uint32_t tb = pe & 15u, kind = (pe >> 4) & 3u, xb = (pe >> 6) & 7u, v = (pe >> 9) & 31u;
					const uint32_t length = WS.len_base[v] + ((pb >> (tb - xb)) & ((1u << xb) - 1u));
					const uint32_t tb2 = pd & 31u, dxb = (pd >> 5) & 15u, ds = (pd >> 9) & 31u;
					const uint32_t dist = WS.dist_base[ds] + ((pb2 >> (tb2 - dxb)) & ((1u << dxb) - 1u));
					const uint32_t run = nl - nl0, opos = nl + mb;
					const bool e_lit = pv && kind == I2_K_LIT, e_len = pv && kind == I2_K_LEN;
					// SEG (not the first segment): the output position in the stream is unknown; remember how far back
					// the matches reach instead (k_seg_stitch checks it against the position the chain gives)
					const bool e_ok = e_len && run < I2_SEQ_ESC && (SEG ? (s_own != 0u || dist <= opos) : dist <= opos);
					if (SEG && e_ok && dist > opos) {
						reach = max(reach, dist - opos);
					}
					if (e_lit) {
						I2_EMIT_LIT((pe >> 6) & 0xFFu);
					}
					if (e_ok) {
						I2_EMIT_MATCH(run, length, dist);
					}
					// (the token scratch has room for the one literal or record that may exceed `cap` here)
					pv = e_len && !e_ok;   // still pending only if the special path has to finish it
					e_special = pv || (act && (nl + mb > cap || (SEG && (uint8_t *)seqp - (seq_floor + nl) < 32)));
				}
				// ---- parse step i
				br.top();
				br.norm();
				const uint32_t bits = br.peek();
				const uint32_t e = lit[bits & ((1u << I2_LIT_ROOT) - 1u)];
				const uint32_t kind = (e >> 4) & 3u;
				br.pos += act ? (e & 15u) : 0u;   // (0 for a LINK entry)
				br.norm();
				const uint32_t bits2 = br.peek();
				const uint32_t d = dst[bits2 & ((1u << I2_DST_ROOT) - 1u)];
				const bool p_len = act && kind == I2_K_LEN, p_root = (d >> 14) == 0u;
				br.pos += (p_len && p_root) ? (d & 31u) : 0u;
				const bool p_plain = act && (kind == I2_K_LIT || (p_len && p_root));
				const bool special = e_special || (act && (!p_plain || br.words_left() <= 1));
				if (__any_sync(0xFFFFFFFFu, special)) {
					if (special) {
						bool bad = false;
						// -- finish step i-1 (a match after >= 511 literals, or an error)
						if (pv) {
							const uint32_t tb = pe & 15u, xb = (pe >> 6) & 7u, v = (pe >> 9) & 31u;
							const uint32_t length = WS.len_base[v] + ((pb >> (tb - xb)) & ((1u << xb) - 1u));
							const uint32_t tb2 = pd & 31u, dxb = (pd >> 5) & 15u, ds = (pd >> 9) & 31u;
							const uint32_t dist = WS.dist_base[ds] + ((pb2 >> (tb2 - dxb)) & ((1u << dxb) - 1u));
							uint32_t r3 = nl - nl0;
							if ((dist > nl + mb && !(SEG && s_own != 0u)) ||
								(SEG && (uint8_t *)seqp - (seq_floor + nl) < (int64_t)(4u * (r3 / I2_SEQ_ESC) + 32u))) {
								bad = true;   // reaches before the start of the output (strict; dec:785 does not check) / slice full
							} else {
								if (SEG && dist > nl + mb) {
									reach = max(reach, dist - (nl + mb));
								}
								while (r3 >= I2_SEQ_ESC) {
									*--seqp = I2_SEQ_ESC;
									nseq++;
									r3 -= I2_SEQ_ESC;
								}
								I2_EMIT_MATCH(r3, length, dist);
							}
							pv = false;
						}
						bad = bad || nl + mb > cap;   // dec:700-703, dec:791-793
						bad = bad || (SEG && (uint8_t *)seqp - (seq_floor + nl) < 32);   // the guessed scratch slice is full
						// -- step i, when it is not a plain literal / root-level match: decode and emit it here
						bool eob = false;
						if (!bad && act && !p_plain) {
							bool want_dist = p_len;   // (then the length code is consumed, the distance code is not)
							uint32_t len2 = 0;
							if (p_len) {
								const uint32_t tb = e & 15u, xb = (e >> 6) & 7u;
								len2 = WS.len_base[(e >> 9) & 31u] + ((bits >> (tb - xb)) & ((1u << xb) - 1u));
							} else if (kind == I2_K_EOB) {
								eob = true;
							} else {
								// second-level literal/length table
								const uint32_t sb = I2_LIT_LINK_BITS(e);
								if (sb == 0u) {
									bad = true;   // dec:693-695: no code matches
								} else {
									const uint32_t e2 = lit[I2_LIT_LINK_OFS(e) + ((bits >> I2_LIT_ROOT) & ((1u << sb) - 1u))];
									const uint32_t k2 = (e2 >> 4) & 3u, t2 = e2 & 15u;
									if (k2 == I2_K_LINK) {
										bad = true;
									} else {
										br.pos += I2_LIT_ROOT + t2;
										if (k2 == I2_K_LIT) {
											I2_EMIT_LIT((e2 >> 6) & 0xFFu);
											bad = nl + mb > cap;
										} else if (k2 == I2_K_EOB) {
											eob = true;
										} else {
											const uint32_t x2 = (e2 >> 6) & 7u;
											len2 = WS.len_base[(e2 >> 9) & 31u] + ((bits >> (I2_LIT_ROOT + t2 - x2)) & ((1u << x2) - 1u));
											want_dist = true;
										}
									}
								}
							}
							if (!bad && want_dist) {
								br.norm_hdr();
								const uint32_t b3 = br.peek();
								uint32_t dd = dst[b3 & ((1u << I2_DST_ROOT) - 1u)];
								uint32_t used = 0;
								if ((dd >> 14) != 0u) {
									const uint32_t sb = I2_DST_LINK_BITS(dd);
									if (sb == 0u) {
										bad = true;   // dec:762-764
									} else {
										dd = dst[I2_DST_LINK_OFS(dd) + ((b3 >> I2_DST_ROOT) & ((1u << sb) - 1u))];
										used = I2_DST_ROOT;
										bad = (dd >> 14) != 0u;
									}
								}
								if (!bad) {
									const uint32_t t3 = dd & 31u, x3 = (dd >> 5) & 15u;
									const uint32_t dist3 = WS.dist_base[(dd >> 9) & 31u] + ((b3 >> (used + t3 - x3)) & ((1u << x3) - 1u));
									br.pos += used + t3;
									uint32_t r3 = nl - nl0;
									if ((dist3 > nl + mb && !(SEG && s_own != 0u)) ||
										(SEG && (uint8_t *)seqp - (seq_floor + nl) < (int64_t)(4u * (r3 / I2_SEQ_ESC) + 32u))) {
										bad = true;
									} else {
										if (SEG && dist3 > nl + mb) {
											reach = max(reach, dist3 - (nl + mb));
										}
										while (r3 >= I2_SEQ_ESC) {
											*--seqp = I2_SEQ_ESC;
											nseq++;
											r3 -= I2_SEQ_ESC;
										}
										I2_EMIT_MATCH(r3, len2, dist3);
										bad = nl + mb > cap;
									}
								}
							}
						}
						if (bad) {
							I2_FALLBACK();
						} else if (eob) {
							// end of block, dec:711-716 (pos already stands behind the code)
							if (final_blk) {
								I2_COMMIT();
							} else {
								state = I2_S_HDR;
								I2_STEP_CHECK();
							}
						} else if (act) {
							I2_STEP_CHECK();
						}
					}
					leave = __any_sync(0xFFFFFFFFu, state != I2_S_DEC && state != I2_S_DONE);
				}
				// step i waits for the next iteration if it is plain and its lane is still decoding
				pv = pv || (p_plain && state == I2_S_DEC);
				pe = p_plain ? e : pe;
				pd = p_plain ? d : pd;
				pb = p_plain ? bits : pb;
				pb2 = p_plain ? bits2 : pb2;
			}
		}
	}
#undef I2_FALLBACK
#undef I2_SEG_COMMIT
#undef I2_COMMIT
#undef I2_STEP_CHECK
#undef I2_EMIT_LIT
#undef I2_EMIT_MATCH
}

// ------------------------------------------------------------------------------------------------
// k_block_search: one thread per byte of a huge stream tests its 8 bit offsets for a dynamic-block header:
// BTYPE = 2, HLIT <= 29, HDIST <= 29, a complete code-length code (Kraft sum), then — for the ~0.1 % that get this
// far — the code lengths themselves: exactly HLIT + HDIST of them, an end-of-block code, complete literal/length
// and distance codes.  grid: (ceil(max comp / 256), n_huge).
__device__ __forceinline__ uint32_t i2_bits_at(const uint8_t *p, uint64_t bit, uint32_t n) {   // n <= 24
	// two aligned words and one funnel shift (the image is padded: the word behind the stream is readable)
	const uint64_t a = reinterpret_cast<uint64_t>(p) + (bit >> 3);
	const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~3ull);
	const uint32_t sh = (uint32_t)(a & 3ull) * 8u + (uint32_t)(bit & 7u);   // <= 31
	return __funnelshift_r(__ldg(w), __ldg(w + 1), sh) & ((1u << n) - 1u);
}

__device__ __noinline__ bool i2_plausible_header(const uint8_t *in, uint64_t nbits, uint64_t p, uint32_t hlit, uint32_t hdist, uint32_t hclen) {
	// (in + nbits/8 + 4 is readable: the image is padded)
	uint8_t pre[128];
	uint32_t cl[19];
	uint64_t q = p + 17;
	for (uint32_t i = 0; i < 19; i++) {
		cl[i] = 0;
	}
	for (uint32_t i = 0; i < hclen; i++) {
		cl[c_cl_order[i]] = i2_bits_at(in, q, 3);
		q += 3;
	}
	uint32_t next[8], cnt[8];
	for (int l = 0; l < 8; l++) {
		cnt[l] = 0;
	}
	for (uint32_t i = 0; i < 19; i++) {
		cnt[cl[i]]++;
	}
	uint32_t code = 0;
	cnt[0] = 0;
	for (int l = 1; l < 8; l++) {
		code = (code + cnt[l - 1]) << 1;
		next[l] = code;
	}
	for (uint32_t sym = 0; sym < 19; sym++) {
		const uint32_t l = cl[sym];
		if (l) {
			const uint32_t c = next[l]++;
			const uint32_t rev = __brev(c) >> (32u - l);
			for (uint32_t x = rev; x < 128u; x += (1u << l)) {
				pre[x] = (uint8_t)(sym | (l << 5));
			}
		}
	}
	const uint32_t total = hlit + hdist;
	uint32_t idx = 0, prev = 0;
	uint32_t kl = 0, kd = 0, nd = 0;   // Kraft sums (units of 2^-15), distance codes used
	bool eob = false;
	while (idx < total) {
		if (q + 14 > nbits) {
			return false;
		}
		const uint32_t bits = i2_bits_at(in, q, 14);
		const uint32_t e = pre[bits & 127u];
		const uint32_t sym = e & 31u, cb = e >> 5;
		uint32_t rep = 1, val = sym;
		if (sym < 16u) {
			q += cb;
			prev = sym;
		} else if (sym == 16u) {
			if (idx == 0) {
				return false;
			}
			val = prev;
			rep = 3u + ((bits >> cb) & 3u);
			q += cb + 2;
		} else if (sym == 17u) {
			val = 0;
			rep = 3u + ((bits >> cb) & 7u);
			q += cb + 3;
			prev = 0;
		} else {
			val = 0;
			rep = 11u + ((bits >> cb) & 127u);
			q += cb + 7;
			prev = 0;
		}
		if (idx + rep > total) {
			return false;
		}
		if (val) {
			for (uint32_t i = 0; i < rep; i++) {
				const uint32_t sidx = idx + i;
				if (sidx < hlit) {
					kl += 32768u >> val;
					eob = eob || sidx == 256u;
				} else {
					kd += 32768u >> val;
					nd++;
				}
			}
			if (kl > 32768u || kd > 32768u) {
				return false;   // over-subscribed: what most false candidates are after a few dozen lengths
			}
		}
		idx += rep;
	}
	return eob && kl == 32768u && (kd == 32768u || nd <= 1u);
}

// grid: persistent (any size); task t = 256 consecutive bytes of one stream, task_ofs[h] = first task of stream h.
// Every thread tests the 8 bit offsets of its byte: BTYPE = 2, HLIT <= 29, HDIST <= 29, and the Kraft sum
// of the code-length code — its (HCLEN + 4) 3-bit lengths are looked up four at a time in a 4096-entry table of
// sum(2^(7 - len)), saturated at 255 — must be exactly 128 (a complete code; zlib never emits another one).
__global__ void __launch_bounds__(256) k_block_search(const uint8_t *__restrict__ archive, const otz_entry *__restrict__ ents,
	const OtzEntryState *__restrict__ est, const int32_t *__restrict__ status, const uint32_t *__restrict__ list, uint32_t n_huge,
	const uint32_t *__restrict__ task_ofs, uint2 *__restrict__ surv, uint32_t surv_cap, uint32_t *__restrict__ n_surv) {
	__shared__ uint8_t s_kraft[4096];
	for (uint32_t i = threadIdx.x; i < 4096u; i += blockDim.x) {
		uint32_t sum = 0;
#pragma unroll
		for (int f = 0; f < 4; f++) {
			const uint32_t l = (i >> (3 * f)) & 7u;
			sum += l ? (128u >> l) : 0u;
		}
		s_kraft[i] = (uint8_t)min(sum, 255u);
	}
	__syncthreads();
	const uint32_t n_tasks = task_ofs[n_huge];
	for (uint32_t t = blockIdx.x; t < n_tasks; t += gridDim.x) {
		// stream of this task: the last h with task_ofs[h] <= t
		uint32_t lo_h = 0, hi_h = n_huge;
		while (hi_h - lo_h > 1) {
			const uint32_t mid = (lo_h + hi_h) >> 1;
			if (task_ofs[mid] <= t) {
				lo_h = mid;
			} else {
				hi_h = mid;
			}
		}
		const uint32_t h = lo_h;
		const uint32_t ei = list[h];
		if (OTZ_ST_CODE(status[ei]) != OTZ_ST_OK) {
			continue;
		}
		const uint32_t comp = ents[ei].comp_size;
		const uint32_t byte = (t - task_ofs[h]) * 256u + threadIdx.x;
		if (byte + 12u > comp) {
			continue;   // a block header needs more than that; the tail belongs to the last segment anyway
		}
		// 12 bytes from `byte` on, as aligned words (the image is padded) shifted into place
		const uint8_t *q = archive + est[ei].data_ofs + byte;
		const uint32_t sh = (uint32_t)(reinterpret_cast<uint64_t>(q) & 3u) * 8u;
		const uint32_t *qw = reinterpret_cast<const uint32_t *>(q - (sh >> 3));
		const uint32_t w0 = __ldg(qw), w1 = __ldg(qw + 1), w2 = __ldg(qw + 2), w3 = __ldg(qw + 3);
		const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
		// first the header fields of all 8 offsets (22 % pass), then the Kraft sums of those that passed
		uint32_t cand = 0;
#pragma unroll
		for (uint32_t o = 0; o < 8u; o++) {
			const uint32_t x0 = __funnelshift_r(b0, b1, o);
			const bool head = ((x0 >> 1) & 3u) == 2u && ((x0 >> 3) & 31u) <= 29u && ((x0 >> 8) & 31u) <= 29u;
			cand |= head ? (1u << o) : 0u;
		}
		uint32_t hits = 0;
		while (cand) {
			const uint32_t o = __ffs(cand) - 1u;
			cand &= cand - 1u;
			// bits o .. o + 80 of the 96 loaded ones: header word (17 bits) and the 57 bits behind it
			const uint32_t x0 = __funnelshift_r(b0, b1, o), x1 = __funnelshift_r(b1, b2, o), x2 = b2 >> o;
			const uint32_t hc = ((x0 >> 13) & 15u) + 4u;
			const uint64_t v2 = (((uint64_t)__funnelshift_r(x1, x2, 17) << 32) | __funnelshift_r(x0, x1, 17)) & ((1ull << (3u * hc)) - 1ull);
			const uint32_t lo32 = (uint32_t)v2, hi32 = (uint32_t)(v2 >> 32);
			const uint32_t kr = (uint32_t)s_kraft[lo32 & 4095u] + s_kraft[(lo32 >> 12) & 4095u] + s_kraft[__funnelshift_r(lo32, hi32, 24) & 4095u] +
				s_kraft[(hi32 >> 4) & 4095u] + s_kraft[(hi32 >> 16) & 4095u];
			hits |= kr == 128u ? (1u << o) : 0u;
		}
		if (byte == 0) {
			hits &= ~1u;   // bit 0 is the stream's own start
		}
		// survivors (~0.1 % of the offsets) go to a list: checking their code lengths here, one lane at a time,
		// would cost more than everything else together
		while (hits) {
			const uint32_t o = __ffs(hits) - 1u;
			hits &= hits - 1u;
			const uint32_t at = atomicAdd(n_surv, 1u);
			if (at < surv_cap) {
				surv[at] = make_uint2(h, byte * 8u + o);
			}
		}
	}
}

// second stage of the search: one thread per surviving offset decodes the code lengths of the would-be header
__global__ void __launch_bounds__(256) k_block_verify(const uint8_t *__restrict__ archive, const otz_entry *__restrict__ ents,
	const OtzEntryState *__restrict__ est, const uint32_t *__restrict__ list, const uint2 *__restrict__ surv, uint32_t surv_cap,
	const uint32_t *__restrict__ n_surv, I2SegCtl seg) {
	const uint32_t n = min(*n_surv, surv_cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint2 sv = surv[i];
		const uint32_t h = sv.x, ei = list[h];
		const uint8_t *in = archive + est[ei].data_ofs;
		const uint32_t w = i2_bits_at(in, sv.y, 17);
		if (i2_plausible_header(in, (uint64_t)ents[ei].comp_size * 8u, sv.y, ((w >> 3) & 31u) + 257u, ((w >> 8) & 31u) + 1u, ((w >> 13) & 15u) + 4u)) {
			const uint32_t at = atomicAdd(&seg.count[h], 1u);
			if (at < I2_MAXSEG - 1u) {
				seg.start[h * I2_MAXSEG + 1u + at] = sv.y;
			}
			// for streams with more candidates than that: the first one of every 1/255 of the stream, so that the
			// segments stay evenly spaced
			const uint32_t bucket = (uint32_t)(((uint64_t)sv.y * (I2_MAXSEG - 1u)) / ((uint64_t)ents[ei].comp_size * 8u));
			atomicMin(&seg.bucket[h * I2_MAXSEG + 1u + min(bucket, (uint32_t)I2_MAXSEG - 2u)], sv.y);
		}
	}
}

// One thread per huge stream: candidates (all of them, or the first one of every bucket when there are more than 255) in
// ascending order behind the true start (bit 0), result rows cleared, every
// segment gets a slice of the stream's token scratch in proportion to its share of the compressed bits, and the
// segments are appended to the work list.
__global__ void k_seg_prepare(const otz_entry *__restrict__ ents, const uint32_t *__restrict__ list, uint32_t n_huge,
	const uint64_t *__restrict__ tok_ofs, I2SegCtl seg) {
	const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h >= n_huge) {
		return;
	}
	uint32_t *st = seg.start + h * I2_MAXSEG;
	uint32_t n = seg.count[h];
	st[0] = 0;
	if (n <= I2_MAXSEG - 1u) {
		for (uint32_t i = 2; i <= n; i++) {   // insertion sort of st[1..n]
			const uint32_t v = st[i];
			uint32_t j = i;
			while (j > 1 && st[j - 1] > v) {
				st[j] = st[j - 1];
				j--;
			}
			st[j] = v;
		}
		n += 1;
	} else {
		const uint32_t *bk = seg.bucket + h * I2_MAXSEG;
		n = 1;
		for (uint32_t b = 1; b < I2_MAXSEG; b++) {   // the buckets are in ascending order by construction
			const uint32_t v = bk[b];
			if (v != 0xFFFFFFFFu) {
				st[n++] = v;
			}
		}
	}
	seg.count[h] = n;
	seg.nlive[h] = 0;
	seg.par[h] = 0;
	const uint32_t ei = list[h];
	const double bits_total = (double)ents[ei].comp_size * 8.0;
	const uint64_t base = tok_ofs[h], total = tok_ofs[h + 1] - tok_ofs[h];
	const uint32_t at = atomicAdd(seg.n_items, n);
	uint64_t lo = base;
	for (uint32_t j = 0; j < n; j++) {
		const uint64_t hi = j + 1 < n ? base + ((uint64_t)((double)total * ((double)st[j + 1] / bits_total)) & ~15ull) : base + total;
		I2SegRes r;
		r.nseq = r.nlit = r.produced = r.end_bit = r.reach = r.flags = 0;
		r.scr_lo = lo;
		r.scr_hi = hi;
		seg.res[h * I2_MAXSEG + j] = r;
		seg.items[at + j] = (h << 16) | j;
		lo = hi;
	}
}

// One thread per huge stream: follow the chain of segments.  The stream is accepted (nlive > 0) only if every link
// fits; otherwise it is appended to the fallback list of k_inflate.
//
// An accepted stream with more than one segment gets room in the symbol buffer (if there is any left): per segment
// I2_PREWIN marker elements, then its output as 16-bit symbols, placed so that symbol i of the segment and output byte i
// have the same index modulo 16 (k_seg_translate works on whole vectors of both).
__global__ void k_seg_stitch(const otz_entry *__restrict__ ents, const int32_t *__restrict__ status, const uint32_t *__restrict__ list,
	uint32_t n_huge, I2SegCtl seg, uint32_t *__restrict__ fb_list, uint32_t *__restrict__ fb_count) {
	const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h >= n_huge) {
		return;
	}
	const uint32_t ei = list[h];
	if (OTZ_ST_CODE(status[ei]) != OTZ_ST_OK) {
		return;
	}
	const uint32_t n = seg.count[h], cap = ents[ei].uncomp_size;
	const I2SegRes *res = seg.res + h * I2_MAXSEG;
	const uint32_t *st = seg.start + h * I2_MAXSEG;
	uint32_t *live = seg.live + h * I2_MAXSEG;
	uint32_t j = 0, nlive = 0;
	uint64_t out = 0;
	bool ok = true, done = false, ref_eob = false;
	while (ok && !done) {
		const I2SegRes r = res[j];
		if (!(r.flags & I2_SEGF_OK) || r.reach > out || out + r.produced > cap) {
			ok = false;
			break;
		}
		live[nlive++] = j;
		out += r.produced;
		if (r.flags & I2_SEGF_FINAL) {
			done = true;
			ref_eob = (r.flags & I2_SEGF_REF_EOB) != 0u;
			break;
		}
		uint32_t jn = j + 1;
		while (jn < n && st[jn] < r.end_bit) {
			jn++;
		}
		if (jn >= n || st[jn] != r.end_bit) {
			ok = false;
			break;
		}
		j = jn;
	}
	if (ok && done && out == cap) {
		seg.nlive[h] = nlive;
		seg.seg_status[h] = OTZ_ST_OK | (ref_eob ? OTZ_STF_REF_EOB : 0);
		uint32_t par = 0;
		if (nlive > 1 && seg.sym_cap) {
			const uint64_t need = (uint64_t)cap + (uint64_t)nlive * (I2_PREWIN + 16u);
			const uint64_t base = atomicAdd(seg.sym_top, (unsigned long long)need);
			if (base + need <= seg.sym_cap) {
				par = 1;
				uint64_t at = base;
				uint32_t o = 0;
				const uint32_t omis = (uint32_t)((seg.out_mis + ents[ei].out_ofs) & 15u);
				for (uint32_t c = 0; c < nlive; c++) {
					uint64_t first = at + I2_PREWIN;
					first += ((omis + o) - (uint32_t)first) & 15u;
					seg.sym_start[h * I2_MAXSEG + c] = first;
					seg.out_start[h * I2_MAXSEG + c] = o;
					const uint32_t pr = res[live[c]].produced;
					at += I2_PREWIN + 16u + pr;
					o += pr;
				}
				const uint32_t w = atomicAdd(seg.n_par, nlive);
				for (uint32_t c = 0; c < nlive; c++) {
					seg.par_items[w + c] = (h << 16) | c;
				}
			}
		}
		seg.par[h] = par;
	} else {
		seg.par[h] = 0;
		seg.nlive[h] = 0;
		fb_list[atomicAdd(fb_count, 1u)] = ei;
	}
}

// Parallel execution of a huge stream.  Every segment of the chain is executed by its own warp over 16-bit symbols:
// values below 256 are bytes, 256 + j stands for "byte j of the 32 KiB before this segment", unknown while the
// segments before it are still being executed (k_inflate_lz<.., PAR> prefills those markers and copies them around
// like any other element).  k_seg_window then walks the chain of ONE stream per CTA and resolves only the last
// 32 KiB of every segment — the window of the next one — straight into the output; step c reads what steps < c
// wrote.  k_seg_translate finally resolves everything in front of those tails, all segments at once.
__global__ void __launch_bounds__(1024) k_seg_window(uint8_t *__restrict__ out, const otz_entry *__restrict__ ents, const uint32_t *__restrict__ list,
	uint32_t n_huge, const uint16_t *__restrict__ sym, int32_t *__restrict__ status, uint32_t *__restrict__ produced_out, I2SegCtl seg) {
	const uint32_t h = blockIdx.x;
	if (h >= n_huge || !seg.par[h]) {
		return;
	}
	const uint32_t ei = list[h];
	uint8_t *const o = out + ents[ei].out_ofs;
	const uint32_t nlive = seg.nlive[h];
	for (uint32_t c = 0; c < nlive; c++) {
		const uint32_t pr = seg.res[h * I2_MAXSEG + seg.live[h * I2_MAXSEG + c]].produced;
		const uint32_t os = seg.out_start[h * I2_MAXSEG + c];
		const uint16_t *sp = sym + seg.sym_start[h * I2_MAXSEG + c];
		const uint32_t t0 = pr > I2_PREWIN ? pr - I2_PREWIN : 0u;
		for (uint32_t i = t0 + threadIdx.x; i < pr; i += blockDim.x) {
			const uint32_t v = __ldcs(sp + i);
			// (a marker only exists where a match reached, and k_seg_stitch checked reach <= os)
			o[os + i] = v < 256u ? (uint8_t)v : __ldcg(o + (os - I2_PREWIN + (v - 256u)));
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		status[ei] = seg.seg_status[h];
		produced_out[ei] = ents[ei].uncomp_size;
	}
}

// One CTA per work item (segment): symbols [0, produced - 32 KiB) -> bytes.  The 32 KiB before the segment (final
// since k_seg_window) are staged in shared memory: in text most vectors still hold a marker or two.
#define I2_TR_THREADS 512
__global__ void __launch_bounds__(I2_TR_THREADS) k_seg_translate(uint8_t *__restrict__ out, const otz_entry *__restrict__ ents, const uint32_t *__restrict__ list,
	const uint16_t *__restrict__ sym, uint32_t *__restrict__ work_counter, I2SegCtl seg) {
	__shared__ __align__(16) uint8_t s_win[I2_PREWIN + 16];
	__shared__ uint32_t s_k;
	const uint32_t n = *seg.n_par;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0) {
			s_k = atomicAdd(work_counter, 1u);
		}
		__syncthreads();
		const uint32_t k = s_k;
		if (k >= n) {
			break;
		}
		const uint32_t it = seg.par_items[k], h = it >> 16, c = it & 0xFFFFu;
		const I2SegRes *r_ = &seg.res[h * I2_MAXSEG + seg.live[h * I2_MAXSEG + c]];
		const uint32_t pr = r_->produced;
		if (pr <= I2_PREWIN) {
			continue;
		}
		const uint32_t body = pr - I2_PREWIN, os = seg.out_start[h * I2_MAXSEG + c];
		uint8_t *const o = out + ents[list[h]].out_ofs + os;        // o[i] <- sp[i]
		const uint16_t *const sp = sym + seg.sym_start[h * I2_MAXSEG + c];
		const uint32_t omis = (uint32_t)(reinterpret_cast<uint64_t>(o) & 15u);
		if (r_->reach) {   // (no reach, no markers; and the first segment of a stream has nothing in front of it)
			// (a segment in the first 32 KiB of its entry: nothing is read in front of the entry's output)
			const int64_t avail = (int64_t)os + (int64_t)((reinterpret_cast<uint64_t>(o) - os) & 15u);
			const uint4 *wa = reinterpret_cast<const uint4 *>(o - I2_PREWIN - omis);
			for (uint32_t x = threadIdx.x; x < I2_PREWIN / 16u + 1u; x += blockDim.x) {
				if ((int64_t)(I2_PREWIN + omis) - (int64_t)(16u * x) <= avail) {
					reinterpret_cast<uint4 *>(s_win)[x] = __ldcg(wa + x);
				}
			}
		}
		__syncthreads();
		const uint8_t *const win = s_win + omis;                    // marker j -> win[j]
		const uint32_t head = min(body, (16u - omis) & 15u);
		const uint32_t nvec = (body - head) >> 4, tail0 = head + (nvec << 4);
		for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) {
			const uint32_t v = sp[i];
			o[i] = v < 256u ? (uint8_t)v : win[(v - 256u) & (I2_PREWIN - 1u)];
		}
		for (uint32_t i = tail0 + threadIdx.x; i < body; i += blockDim.x) {
			const uint32_t v = sp[i];
			o[i] = v < 256u ? (uint8_t)v : win[(v - 256u) & (I2_PREWIN - 1u)];
		}
#pragma unroll 2
		for (uint32_t x = threadIdx.x; x < nvec; x += blockDim.x) {
			const uint32_t i = head + (x << 4);
			const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(sp + i)), b = __ldcs(reinterpret_cast<const uint4 *>(sp + i + 8));
			uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
			uint32_t r[4];
			if (((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) & 0xFF00FF00u) == 0u) {
#pragma unroll
				for (int j = 0; j < 4; j++) {
					r[j] = __byte_perm(w[2 * j], w[2 * j + 1], 0x6420);
				}
			} else {
#pragma unroll
				for (int j = 0; j < 4; j++) {
					uint32_t acc = 0;
#pragma unroll
					for (int t = 0; t < 4; t++) {
						const uint32_t v = (w[2 * j + (t >> 1)] >> (16 * (t & 1))) & 0xFFFFu;
						const uint32_t byte = v < 256u ? v : (uint32_t)win[(v - 256u) & (I2_PREWIN - 1u)];
						acc |= byte << (8 * t);
					}
					r[j] = acc;
				}
			}
			__stcs(reinterpret_cast<uint4 *>(o + i), make_uint4(r[0], r[1], r[2], r[3]));
		}
	}
}

// ------------------------------------------------------------------------------------------------
// phase B.  One warp per committed stream; W = bytes of shared-memory ring per warp.  Ring index of output
// byte p is (p + mis) & (W-1) with mis = dst & 15 (as OutRing in k_inflate.cuh), so ring vectors line up with
// 16-byte aligned global vectors.
//
// The sequence records are executed in batches of up to 32 (one per lane), software-pipelined over two batches:
// while batch k runs, batch k+1 has already been scanned (warp prefix sums give every record its literal and
// output positions), its match descriptors sit in shared memory, the first literals of every run are in
// registers and — the point of the exercise — the sources of its FAR matches (those that left the ring; they
// were flushed to HBM before batch k started, see SPAN_MAX) are on their way into a staging buffer as 16-byte
// cp.async copies.  Executing a batch therefore touches shared memory only.
// The executor is written over an element type T: bytes for ordinary streams, 16-bit SYMBOLS for the segments of a
// huge stream that are executed in parallel (k_seg_window below): positions, distances and lengths are in elements,
// VEC = elements per 16-byte vector.
template <typename T>
struct I2Elem {
	static constexpr uint32_t VEC = 16u / sizeof(T);
	static constexpr uint32_t STAGE_VECS = sizeof(T) == 1 ? 64u : 96u;   // staging vectors per batch; far matches beyond them are fetched directly
	static constexpr uint32_t STAGE = STAGE_VECS * VEC;
};

template <bool PAR>
struct I2ElemOf {
	typedef uint8_t type;
};
template <>
struct I2ElemOf<true> {
	typedef uint16_t type;
};

template <int W>
struct I2Ring {
	static constexpr uint32_t MASK = W - 1;
	static constexpr uint32_t SEG = 512;
	// output bytes per batch.  W >= 2 * SPAN_MAX + SEG + 258 guarantees that a far source of batch k+1 has been
	// flushed to HBM before batch k starts (when its copy is issued).
	static constexpr uint32_t SPAN_MAX = W >= 8192 ? 2048u : 1024u;
};

template <int W, typename T = uint8_t>
struct __align__(16) I2LzSmem {
	T ring[W];
	T stage[2][I2Elem<T>::STAGE];   // (directly behind the ring: a match source is ONE index from the ring's base, see I2Batch::mb)
};

// write ring[a, b) (linear positions) to HBM; whole warp, ring contents visible (caller synced)
template <int W, typename T>
__device__ __forceinline__ void i2_flush_range(T *gbase, const T *ring, uint32_t a, uint32_t b, uint32_t lane) {
	constexpr uint32_t MASK = W - 1, VEC = I2Elem<T>::VEC;
	const uint32_t a16 = (a + VEC - 1u) & ~(VEC - 1u), b16 = b & ~(VEC - 1u);
	if (a16 >= b16) {
		for (uint32_t x = a + lane; x < b; x += 32) {
			gbase[x] = ring[x & MASK];
		}
		return;
	}
	for (uint32_t x = a + lane; x < a16; x += 32) {
		gbase[x] = ring[x & MASK];
	}
#pragma unroll 1
	for (uint32_t x = a16 + VEC * lane; x < b16; x += 32u * VEC) {
		*reinterpret_cast<uint4 *>(gbase + x) = *reinterpret_cast<const uint4 *>(ring + (x & MASK));
	}
	for (uint32_t x = b16 + lane; x < b; x += 32) {
		gbase[x] = ring[x & MASK];
	}
}

// overlapping LZ77 copy (distance < length): periodic extension of the last `dd` bytes, whole warp (dec:521-533)
template <int W, typename T>
__device__ __noinline__ void i2_copy_periodic(T *rb, uint32_t dq, uint32_t sq, uint32_t dd, uint32_t len, uint32_t lane) {
	constexpr uint32_t MASK = W - 1;
	uint32_t r = dd > lane ? lane : lane % dd;
	const uint32_t step = dd > 32u ? 32u : 32u % dd;
	for (uint32_t x = lane; x < len; x += 32) {
		rb[(dq + x) & MASK] = rb[(sq + r) & MASK];
		r += step;
		r = r >= dd ? r - dd : r;
	}
}

// what the scan of one batch leaves in registers for its execution
struct I2Batch {
	uint32_t ntake;     // records in the batch (0 = none left)
	uint32_t tot_l;     // literals it consumes
	uint32_t q_end;     // linear output position behind it
	uint32_t lr;        // this lane's record: literal run,
	uint32_t my_lit;    //   its first literal,
	uint32_t my_out;    //   its linear output position
	uint32_t lit4;      //   and its first four literal bytes
	// the match of this lane's record, as the executor wants it (it walks the records in stream order and fetches
	// these two words with one shuffle each):
	uint32_t ma;        //   ring index of the destination | length << 16 | I2_MF_* << 28 | bit 31 = any flag   (length 0: no match)
	uint32_t mb;        //   source, as an element index from the ring's base: a ring index (< W), or an index into the staging
	                    //   buffer behind the ring (>= W: a FAR source, fetched while the batch before ran)
	uint32_t m_src;     //   linear position of the source (far matches that did not fit the staging buffer are fetched directly)
	uint32_t m_dist;    //   distance (overlapping matches are extended periodically)
};
#define I2_MF_LONG 1u       // longer than 64 elements, or a range that wraps around the ring
#define I2_MF_PERIODIC 2u   // distance < length
#define I2_MF_DIRECT 4u     // far source not staged: read from HBM

// scan records [b, b + 32) (this lane holds record b + lane in `rec`): positions by warp prefix sums, the match descriptor
// of every record (kept in its lane), and the copies of the batch's far sources into staging buffer `buf` are started.
// WIDE: 8-byte records {literal run | (length - 3) << 9, distance} (Zstandard: distances beyond 32 KiB); `rec` is the first
// word, `wdist` the second.  Otherwise the distance sits in rec[31:17].
template <int W, bool WIDE, typename T>
__device__ __forceinline__ I2Batch i2_scan_batch(I2LzSmem<W, T> &S, int buf, uint32_t rec, uint32_t wdist, uint32_t b, uint32_t nseq, uint32_t q,
	uint32_t lp, const uint8_t *__restrict__ lits, const T *gbase, uint32_t lane) {
	constexpr uint32_t MASK = I2Ring<W>::MASK, SPAN_MAX = I2Ring<W>::SPAN_MAX, VEC = I2Elem<T>::VEC, STAGE_VECS = I2Elem<T>::STAGE_VECS;
	const uint32_t lt_mask = (1u << lane) - 1u;
	I2Batch B;
	const bool have = b + lane < nseq;
	const uint32_t lr = have ? rec & 511u : 0u;
	const uint32_t ml = (have && lr != I2_SEQ_ESC) ? ((rec >> 9) & 255u) + 3u : 0u;
	const uint32_t dist = WIDE ? wdist : (rec >> 17) + 1u;
	uint32_t lsum = lr, osum = lr + ml;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, lsum, d), c = __shfl_up_sync(0xFFFFFFFFu, osum, d);
		if ((int)lane >= d) {
			lsum += a;
			osum += c;
		}
	}
	// the longest prefix of the batch whose output fits SPAN_MAX (one record is at most 769 bytes)
	const uint32_t fit = __ballot_sync(0xFFFFFFFFu, have && osum <= SPAN_MAX);
	B.ntake = __popc(fit);   // osum is monotonic, so `fit` is a prefix mask
	const uint32_t last = B.ntake ? B.ntake - 1u : 0u;
	B.tot_l = B.ntake ? __shfl_sync(0xFFFFFFFFu, lsum, last) : 0u;
	const uint32_t tot_o = B.ntake ? __shfl_sync(0xFFFFFFFFu, osum, last) : 0u;
	const bool mine = lane < B.ntake;
	B.lr = mine ? lr : 0u;
	B.my_lit = lp + lsum - lr;
	B.my_out = q + osum - lr - ml;
	B.q_end = q + tot_o;
	// the first literals of the run travel in a register
	B.lit4 = 0;
#pragma unroll
	for (int t = 0; t < 4; t++) {
		if ((uint32_t)t < B.lr) {
			B.lit4 |= (uint32_t)lits[B.my_lit + t] << (8 * t);
		}
	}
	// far: the source lies below the ring window [q_end - W, q_end) of this batch
	const uint32_t mq = B.my_out + lr;   // match destination
	const bool is_match = mine && ml != 0u;
	const bool is_far = is_match && dist > (uint32_t)W - (B.q_end - mq);
	const uint32_t far_m = __ballot_sync(0xFFFFFFFFu, is_far);
	const uint32_t src_lin = mq - dist;
	// LONG: more than 64 elements, or the destination / a ring source wraps around the ring
	uint32_t flags = is_match ? (((ml > 64u || (mq & MASK) + ml > (uint32_t)W || (!is_far && (src_lin & MASK) + ml > (uint32_t)W)) ? I2_MF_LONG : 0u) |
		((!is_far && dist < ml) ? I2_MF_PERIODIC : 0u)) : 0u;
	B.mb = src_lin & MASK;
	B.m_src = src_lin;
	B.m_dist = dist;
	if (far_m) {
		// staging vectors per far match (the source is copied as whole 16-byte vectors)
		const uint32_t soff = src_lin & (VEC - 1u);
		const uint32_t nch = is_far ? (soff + ml + VEC - 1u) / VEC : 0u;
		uint32_t incl = nch;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if ((int)lane >= d) {
				incl += a;
			}
		}
		if (is_far) {
			if (incl <= STAGE_VECS) {
				const uint32_t cst = incl - nch;
				B.mb = (uint32_t)W + (uint32_t)buf * I2Elem<T>::STAGE + VEC * cst + soff;
				// every lane fetches the vectors of its own match (cp.async groups are per thread: the executor waits for
				// its own group and then syncs the warp)
				uint32_t sa = (uint32_t)__cvta_generic_to_shared(&S.stage[buf][VEC * cst]);
				const T *g = gbase + (src_lin & ~(VEC - 1u));
#pragma unroll 1
				for (uint32_t v = 0; v < nch; v++) {
					asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
					sa += 16u;
					g += VEC;
				}
			} else {
				flags |= I2_MF_DIRECT;
			}
		}
	}
	B.ma = (mq & MASK) | ((is_match ? ml : 0u) << 16) | (flags << 28) | (flags ? 0x80000000u : 0u);
	asm volatile("cp.async.commit_group;" ::: "memory");
	return B;
}

// SEG: `list` holds huge streams; the tokens of a stream are the chain of segments k_seg_stitch accepted, walked by
// this warp.  PAR: the work items are single segments of such chains (seg.par_items), executed over 16-bit symbols
// into the symbol buffer `out` with the 32 KiB before the segment standing in as MARKERS (k_seg_window resolves them).
template <int W, bool WIDE, bool SEG, bool PAR = false>
__global__ void __launch_bounds__(128) k_inflate_lz(uint8_t *__restrict__ out, const otz_entry *__restrict__ ents, const uint32_t *__restrict__ list,
	uint32_t n_list, uint32_t *__restrict__ work_counter, const uint8_t *__restrict__ scratch, const uint64_t *__restrict__ tok_ofs,
	const I2TokRes *__restrict__ tokres, int32_t *__restrict__ status, uint32_t *__restrict__ produced_out, I2SegCtl seg) {
	typedef typename I2ElemOf<PAR>::type T;
	static_assert(!PAR || (SEG && !WIDE), "PAR executes segments of DEFLATE streams");
	extern __shared__ __align__(16) uint8_t smem_raw[];
	constexpr uint32_t MASK = I2Ring<W>::MASK, SEGB = I2Ring<W>::SEG, SPAN_MAX = I2Ring<W>::SPAN_MAX, VEC = I2Elem<T>::VEC;
	const uint32_t lane = threadIdx.x & 31u;
	I2LzSmem<W, T> &S = reinterpret_cast<I2LzSmem<W, T> *>(smem_raw)[threadIdx.x >> 5];
	T *const rb = S.ring;
	if (PAR) {
		n_list = *seg.n_par;
	}
	for (;;) {
		uint32_t k = 0;
		if (lane == 0) {
			k = atomicAdd(work_counter, 1u);
		}
		k = __shfl_sync(0xFFFFFFFFu, k, 0);
		if (k >= n_list) {
			break;
		}
		I2TokRes tr;
		uint32_t nsegs = 1, par_c = 0;
		if (PAR) {
			const uint32_t it = seg.par_items[k];
			k = it >> 16;        // from here on: the stream
			par_c = it & 0xFFFFu;
			tr.ok = 1;
			tr.status = 0;
			tr.nseq = tr.nlit = 0;
		} else if (SEG) {
			nsegs = seg.par[k] ? 0u : seg.nlive[k];
			tr.ok = nsegs != 0u;
			tr.status = seg.seg_status[k];
			tr.nseq = tr.nlit = 0;
		} else {
			tr = tokres[k];
		}
		if (!tr.ok) {
			continue;
		}
		const uint32_t ei = list[k];
		const otz_entry e = ents[ei];
		T *const dstp = PAR ? reinterpret_cast<T *>(out) + seg.sym_start[k * I2_MAXSEG + par_c] : reinterpret_cast<T *>(out + e.out_ofs);
		const uint32_t mis = (uint32_t)((reinterpret_cast<uint64_t>(dstp) / sizeof(T)) & (VEC - 1u));
		// PAR: linear position I2_PREWIN is the segment's first element; the markers sit below it
		T *const gbase = dstp - mis - (PAR ? I2_PREWIN : 0u);
		uint32_t q = mis + (PAR ? I2_PREWIN : 0u), qf = q;   // linear write position / position up to which HBM holds the data
		if (PAR) {
			// marker for the element m places before the segment: 256 + (I2_PREWIN - m).  HBM gets all the matches can
			// reach, the ring the part of them a near match can address.
			const uint32_t reach = min(seg.res[k * I2_MAXSEG + seg.live[k * I2_MAXSEG + par_c]].reach, (uint32_t)I2_PREWIN);
			for (uint32_t m = 1u + lane; m <= reach; m += 32) {
				const T v = (T)(256u + I2_PREWIN - m);
				gbase[q - m] = v;
				if (m <= (uint32_t)W) {
					rb[(q - m) & MASK] = v;
				}
			}
			__syncwarp();
		}
		for (uint32_t sgi = 0; sgi < nsegs; sgi++) {
			const uint8_t *lits;
			const uint32_t *seq_end;
			uint32_t nseq, nlit;
			if (SEG) {
				const I2SegRes *r_ = &seg.res[k * I2_MAXSEG + seg.live[k * I2_MAXSEG + (PAR ? par_c : sgi)]];
				lits = scratch + r_->scr_lo;
				seq_end = reinterpret_cast<const uint32_t *>(scratch + r_->scr_hi);
				nseq = r_->nseq;
				nlit = r_->nlit;
			} else {
				lits = scratch + tok_ofs[k];
				seq_end = reinterpret_cast<const uint32_t *>(scratch + tok_ofs[k + 1]);
				nseq = tr.nseq;
				nlit = tr.nlit;
			}
			uint32_t lp = 0, b = 0;       // literals / records consumed
			__syncwarp();
			// records b + lane (recA) and b + 32 + lane (recB); WIDE: their second words in offA / offB
			const uint2 *const seq_end2 = reinterpret_cast<const uint2 *>(seq_end);
			uint32_t recA = 0, recB = 0, offA = 0, offB = 0;
			if (WIDE) {
				if (lane < nseq) {
					const uint2 r = seq_end2[-1 - (int32_t)lane];
					recA = r.x;
					offA = r.y;
				}
				if (32u + lane < nseq) {
					const uint2 r = seq_end2[-33 - (int32_t)lane];
					recB = r.x;
					offB = r.y;
				}
			} else {
				recA = lane < nseq ? __ldcs(seq_end - 1 - lane) : 0u;
				recB = 32u + lane < nseq ? __ldcs(seq_end - 33 - lane) : 0u;
			}
			int buf = 0;
			I2Batch cur = i2_scan_batch<W, WIDE, T>(S, 0, recA, offA, 0u, nseq, q, lp, lits, gbase, lane);
			while (cur.ntake) {
				// ---- scan batch k+1 and start its far copies
				const uint32_t b2 = b + cur.ntake;
				{
					const uint32_t j = cur.ntake + lane;
					const uint32_t fromA = __shfl_sync(0xFFFFFFFFu, recA, j & 31u), fromB = __shfl_sync(0xFFFFFFFFu, recB, j & 31u);
					recA = j < 32u ? fromA : fromB;
					if (WIDE) {
						const uint32_t oA = __shfl_sync(0xFFFFFFFFu, offA, j & 31u), oB = __shfl_sync(0xFFFFFFFFu, offB, j & 31u);
						offA = j < 32u ? oA : oB;
						recB = offB = 0;
						if (b2 + 32u + lane < nseq) {
							const uint2 r = seq_end2[-33 - (int32_t)(b2 + lane)];
							recB = r.x;
							offB = r.y;
						}
					} else {
						recB = b2 + 32u + lane < nseq ? __ldcs(seq_end - 33 - (b2 + lane)) : 0u;
					}
				}
				const I2Batch nxt = i2_scan_batch<W, WIDE, T>(S, buf ^ 1, recA, offA, b2, nseq, cur.q_end, lp + cur.tot_l, lits, gbase, lane);
				// ---- execute batch k: its far sources have landed in stage[buf]
				asm volatile("cp.async.wait_group 1;" ::: "memory");
				__syncwarp();
				{
					// literal runs, all records at once
					const uint32_t n4 = min(cur.lr, 4u);
					for (uint32_t t = 0; t < n4; t++) {
						rb[(cur.my_out + t) & MASK] = (T)(uint8_t)(cur.lit4 >> (8u * t));
					}
#pragma unroll 1
					for (uint32_t t = 4; t < cur.lr; t++) {
						rb[(cur.my_out + t) & MASK] = lits[cur.my_lit + t];
					}
				}
				// the matches in stream order: one shuffle pair fetches the record's descriptor; the common case — up to 64 elements,
				// no overlap, neither range wraps around the ring — is two loads and two stores per lane without any index masking
				// (source: ring or staging buffer, one index space); everything else takes the side exit
				const uint32_t lane32 = lane + 32u;
#pragma unroll 1
				for (uint32_t r = 0; r < cur.ntake; r++) {
					const uint32_t a = __shfl_sync(0xFFFFFFFFu, cur.ma, r), bsrc = __shfl_sync(0xFFFFFFFFu, cur.mb, r);
					__syncwarp();   // earlier ring stores are visible to the loads below
					if ((int32_t)a >= 0) {
						const uint32_t len = a >> 16, dq = a & 0xFFFFu;
						T v0 = 0, v1 = 0;
						if (lane < len) {
							v0 = rb[bsrc + lane];
						}
						if (lane32 < len) {
							v1 = rb[bsrc + lane32];
						}
						if (lane < len) {
							rb[dq + lane] = v0;
						}
						if (lane32 < len) {
							rb[dq + lane32] = v1;
						}
						continue;
					}
					const uint32_t len = (a >> 16) & 0x1FFu, dq = a & 0xFFFFu, fl = (a >> 28) & 7u;
					const uint32_t smask = bsrc >= (uint32_t)W ? 0xFFFFFFFFu : MASK;   // staged sources are not wrapped
					if (fl & I2_MF_PERIODIC) {
						i2_copy_periodic<W, T>(rb, dq, bsrc, __shfl_sync(0xFFFFFFFFu, cur.m_dist, r), len, lane);
					} else if (fl & I2_MF_DIRECT) {
						const uint32_t src = __shfl_sync(0xFFFFFFFFu, cur.m_src, r);
#pragma unroll 1
						for (uint32_t x = lane; x < len; x += 32) {
							rb[(dq + x) & MASK] = __ldcg(gbase + src + x);
						}
					} else {
#pragma unroll 1
						for (uint32_t x = lane; x < len; x += 32) {
							rb[(dq + x) & MASK] = rb[(bsrc + x) & smask];
						}
					}
				}
				b = b2;
				lp += cur.tot_l;
				q = cur.q_end;
				__syncwarp();
				const uint32_t qa = q & ~(SEGB - 1u);
				if (qa > qf) {
					i2_flush_range<W, T>(gbase, rb, qf, qa, lane);
					qf = qa;
					__syncwarp();
				}
				cur = nxt;
				buf ^= 1;
			}
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			// literals after the last match
			while (lp < nlit) {
				const uint32_t n = min(nlit - lp, SPAN_MAX);
				for (uint32_t t = lane; t < n; t += 32) {
					rb[(q + t) & MASK] = lits[lp + t];
				}
				lp += n;
				q += n;
				__syncwarp();
				const uint32_t qa = q & ~(SEGB - 1u);
				if (qa > qf) {
					i2_flush_range<W, T>(gbase, rb, qf, qa, lane);
					qf = qa;
					__syncwarp();
				}
			}
		}   // segments
		if (q > qf) {
			i2_flush_range<W, T>(gbase, rb, qf, q, lane);
		}
		if (!PAR && lane == 0) {
			status[ei] = tr.status;
			produced_out[ei] = e.uncomp_size;
		}
		__syncwarp();
	}
}
