// k_zstd_enc.cuh — Zstandard (RFC 8878) block encoder for method-93 entries, the emit phase of k_deflate_chunks.
//
// SURVEY.md §8f rank 3.  The reference's "zstd" writer is a raw-block stub whose frames its own reader rejects
// (/root/reference/src/lib/zstd.inc.c:172-182, :269) and whose output always loses against STORE
// (/root/reference/src/lib/otezip.c:894-899), so `otezip -c … -z zstd` never writes a method-93 entry.  This is the
// replacement: real frames that libzstd, 7-Zip and this library's own k_zstd* decoders read.
//
// The LZ77 phase is the DEFLATE compressor's (k_deflate.cuh phase 1: 4-byte hash, lazy greedy parse, matches of 4..258
// bytes at distances <= 32768 that never cross a chunk start) — every such match is a legal Zstandard sequence.  Per
// chunk of <= 65,280 bytes ONE block:
//   literals section   Raw_Literals_Block (3-byte header), the chunk's literal bytes in order (warp-parallel gather);
//   sequences section  count, Symbol_Compression_Modes = 0 (all three Predefined_Mode), then the FSE bitstream: the
//                      sequences last to first, states initialised from the last one, offset / match-length /
//                      literal-length state bits, then the three extra-bit fields (offset_value = offset + 3: repeat
//                      offsets are never emitted), the final states, a 1 bit.  The three state chains are serial, so
//                      one lane writes the stream (~2,400 sequences per chunk of log text, ~7 % of the chunk's time).
// A block that would not be smaller than its input becomes a Raw_Block; the first chunk of an entry carries the frame
// header (magic, Window_Descriptor 64 KiB, 4-byte Frame_Content_Size, no checksum), the last one the Last_Block bit;
// chunk outputs concatenate with memcpy exactly like the DEFLATE chunks.
//
// The encoding tables come from the predefined distributions of RFC 8878 §3.1.1.3.2.2, built on the host when a context
// is created (zse_build_tables, the construction every FSE encoder uses) and kept in constant memory.
#pragma once
#include "otz_common.cuh"

#define ZSE_SEQ_OFS 32640u   // word offset of the sequence array inside a chunk's token scratch (2 words per sequence, <= 16,320 sequences)

struct ZseTables {
	uint16_t ll_tab[64], ml_tab[64], of_tab[32];   // next-state tables
	uint32_t ll_dnb[36], ml_dnb[53], of_dnb[29];   // deltaNbBits per symbol
	int32_t ll_dfs[36], ml_dfs[53], of_dfs[29];    // deltaFindState per symbol
	uint8_t ll_code[64], ml_code[128];             // literal-length / (match length - 3) -> code for the small values
	uint8_t ll_bits[36], ml_bits[53];              // extra bits per code
};
__constant__ ZseTables c_zse;

// ---------------------------------------------------------------- host: tables from the predefined distributions
static inline int zse_highbit(uint32_t v) {
	int r = -1;
	while (v) {
		r++;
		v >>= 1;
	}
	return r;
}
static inline void zse_build_ctable(const int *norm, int n, int log, uint16_t *tab, uint32_t *dnb, int32_t *dfs) {
	const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
	int cumul[64], sym[64], high = size - 1;
	cumul[0] = 0;
	for (int s = 0; s < n; s++) {
		if (norm[s] == -1) {
			cumul[s + 1] = cumul[s] + 1;
			sym[high--] = s;   // "less than one" probability: one cell at the end of the table
		} else {
			cumul[s + 1] = cumul[s] + norm[s];
		}
	}
	int pos = 0;
	for (int s = 0; s < n; s++) {
		for (int k = 0; k < norm[s]; k++) {
			sym[pos] = s;
			pos = (pos + step) & mask;
			while (pos > high) {
				pos = (pos + step) & mask;
			}
		}
	}
	for (int u = 0; u < size; u++) {
		tab[cumul[sym[u]]++] = (uint16_t)(size + u);
	}
	int total = 0;
	for (int s = 0; s < n; s++) {
		const int c = norm[s];
		if (c == 0) {
			dnb[s] = (uint32_t)(((log + 1) << 16) - (1 << log));
			dfs[s] = 0;
		} else if (c == -1 || c == 1) {
			dnb[s] = (uint32_t)((log << 16) - (1 << log));
			dfs[s] = total - 1;
			total++;
		} else {
			const int max_bits_out = log - zse_highbit((uint32_t)c - 1);
			dnb[s] = (uint32_t)((max_bits_out << 16) - (c << max_bits_out));
			dfs[s] = total - c;
			total += c;
		}
	}
}
static inline void zse_build_tables(ZseTables *t) {
	static const int ll_norm[36] = { 4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1 };
	static const int ml_norm[53] = { 1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
		1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1 };
	static const int of_norm[29] = { 1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1 };
	static const uint8_t ll_bits[36] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
	static const uint8_t ml_bits[53] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
		0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
	memset(t, 0, sizeof(*t));
	zse_build_ctable(ll_norm, 36, 6, t->ll_tab, t->ll_dnb, t->ll_dfs);
	zse_build_ctable(ml_norm, 53, 6, t->ml_tab, t->ml_dnb, t->ml_dfs);
	zse_build_ctable(of_norm, 29, 5, t->of_tab, t->of_dnb, t->of_dfs);
	memcpy(t->ll_bits, ll_bits, 36);
	memcpy(t->ml_bits, ml_bits, 53);
	// code of a value = the last code whose baseline does not exceed it (baselines: 0..15 one by one, then in steps of 2^bits)
	for (int c = 0, base = 0; c < 36 && base < 64; c++) {
		for (int k = 0; k < (1 << ll_bits[c]) && base + k < 64; k++) {
			t->ll_code[base + k] = (uint8_t)c;
		}
		base += 1 << ll_bits[c];
	}
	for (int c = 0, base = 0; c < 53 && base < 128; c++) {
		for (int k = 0; k < (1 << ml_bits[c]) && base + k < 128; k++) {
			t->ml_code[base + k] = (uint8_t)c;
		}
		base += 1 << ml_bits[c];
	}
}

// ---------------------------------------------------------------- device
__device__ __forceinline__ uint32_t zse_ll_code(uint32_t ll) { return ll < 64u ? c_zse.ll_code[ll] : 50u - (uint32_t)__clz(ll); }   // highbit + 19
__device__ __forceinline__ uint32_t zse_ml_code(uint32_t mb) { return mb < 128u ? c_zse.ml_code[mb] : 67u - (uint32_t)__clz(mb); }  // highbit + 36

// forward bit writer of the sequences section (the decoder reads it backwards); one lane
struct ZseBits {
	uint8_t *p;
	const uint8_t *limit;   // writing stops (and the block is given up) once p would reach it
	uint64_t acc;
	uint32_t n;
	__device__ __forceinline__ void add(uint32_t v, uint32_t nb) {
		acc |= (uint64_t)(v & ((nb < 32u ? (1u << nb) : 0u) - 1u)) << n;
		n += nb;
	}
	__device__ __forceinline__ void flush() {   // leaves fewer than 8 bits in the accumulator
		while (n >= 8u && p < limit) {
			*p++ = (uint8_t)acc;
			acc >>= 8;
			n -= 8u;
		}
	}
};

// One Zstandard block (and, for the first chunk of an entry, the frame header) from the chunk's LZ77 tokens.
// tok[0, ntok): 0x80000000 | (distance - 1) << 8 | (length - 3) for a match, the byte for a literal (k_deflate.cuh).
// Returns the bytes written to co.  Whole warp.
__device__ __noinline__ uint32_t zse_emit_block(const uint8_t *__restrict__ src, uint32_t n, uint32_t *tok, uint32_t ntok, uint8_t *co, bool first, bool last,
	uint32_t entry_len, int lane) {
	uint32_t hdr = 0;
	if (first) {
		if (lane == 0) {
			co[0] = 0x28;   // magic 0xFD2FB528, little endian
			co[1] = 0xB5;
			co[2] = 0x2F;
			co[3] = 0xFD;
			co[4] = 0x80;   // Frame_Header_Descriptor: 4-byte Frame_Content_Size, window descriptor present, no checksum, no dictionary
			co[5] = 0x30;   // Window_Descriptor: 64 KiB (the matches reach at most 32 KiB back)
			co[6] = (uint8_t)entry_len;
			co[7] = (uint8_t)(entry_len >> 8);
			co[8] = (uint8_t)(entry_len >> 16);
			co[9] = (uint8_t)(entry_len >> 24);
		}
		hdr = 10;
	}
	uint8_t *const blk = co + hdr;   // Block_Header at blk[0, 3), literals header at blk[3, 6), literal bytes from blk + 6
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t *const seqs = tok + ZSE_SEQ_OFS;
	uint32_t nlit = 0, nseq = 0, prev_l = 0;
	const bool room = ntok <= ZSE_SEQ_OFS;   // (a chunk with more tokens than that is nearly all literals: raw block)
	if (room) {
		for (uint32_t base = 0; base < ntok; base += 32) {
			const bool valid = base + lane < ntok;
			const uint32_t tk = valid ? tok[base + lane] : 0u;
			const bool ism = valid && (tk >> 31), isl = valid && !(tk >> 31);
			const uint32_t lm = __ballot_sync(0xFFFFFFFFu, isl), mm = __ballot_sync(0xFFFFFFFFu, ism);
			const uint32_t L = nlit + __popc(lm & lt_mask);   // literals in front of this token
			if (isl) {
				blk[6u + L] = (uint8_t)tk;
			}
			const uint32_t below = mm & lt_mask;
			const uint32_t lprev = __shfl_sync(0xFFFFFFFFu, L, below ? 31 - __clz(below) : 0);
			if (ism) {
				const uint32_t k = nseq + __popc(below);
				seqs[2u * k] = (L - (below ? lprev : prev_l)) | ((tk & 0xFFu) << 24);   // literal length (<= 65,280) | (match length - 3) << 24
				seqs[2u * k + 1u] = ((tk >> 8) & 0x7FFFu) + 1u;                              // offset
			}
			if (mm) {
				prev_l = __shfl_sync(0xFFFFFFFFu, L, 31 - __clz(mm));
			}
			nlit += __popc(lm);
			nseq += __popc(mm);
		}
	}
	__syncwarp();
	uint32_t csize = 0;   // size of the compressed block's content, 0 = give up
	if (room && nseq != 0u && 6u + nlit + 4u < n) {
		if (lane == 0) {
			ZseBits bw;
			bw.p = blk + 6u + nlit;
			bw.limit = blk + 3u + n - 1u;   // the content must come out smaller than the chunk
			bw.acc = 0;
			bw.n = 0;
			if (nseq < 128u) {
				*bw.p++ = (uint8_t)nseq;
			} else if (nseq < 0x7F00u) {
				*bw.p++ = (uint8_t)((nseq >> 8) + 128u);
				*bw.p++ = (uint8_t)nseq;
			} else {
				*bw.p++ = 255;
				*bw.p++ = (uint8_t)(nseq - 0x7F00u);
				*bw.p++ = (uint8_t)((nseq - 0x7F00u) >> 8);
			}
			*bw.p++ = 0;   // Symbol_Compression_Modes: Predefined_Mode x 3
			uint32_t s_ll, s_ml, s_of;
			{
				const uint32_t w0 = seqs[2u * (nseq - 1u)], ofv = seqs[2u * (nseq - 1u) + 1u] + 3u;
				const uint32_t ll = w0 & 0xFFFFFFu, mb = w0 >> 24;
				const uint32_t llc = zse_ll_code(ll), mlc = zse_ml_code(mb), ofc = 31u - (uint32_t)__clz(ofv);
				uint32_t nb = (c_zse.ml_dnb[mlc] + (1u << 15)) >> 16;
				s_ml = c_zse.ml_tab[(int32_t)(((nb << 16) - c_zse.ml_dnb[mlc]) >> nb) + c_zse.ml_dfs[mlc]];
				nb = (c_zse.of_dnb[ofc] + (1u << 15)) >> 16;
				s_of = c_zse.of_tab[(int32_t)(((nb << 16) - c_zse.of_dnb[ofc]) >> nb) + c_zse.of_dfs[ofc]];
				nb = (c_zse.ll_dnb[llc] + (1u << 15)) >> 16;
				s_ll = c_zse.ll_tab[(int32_t)(((nb << 16) - c_zse.ll_dnb[llc]) >> nb) + c_zse.ll_dfs[llc]];
				bw.add(ll, c_zse.ll_bits[llc]);
				bw.add(mb, c_zse.ml_bits[mlc]);
				bw.add(ofv, ofc);
				bw.flush();
			}
			for (uint32_t k = nseq - 1u; k-- > 0u && bw.p < bw.limit;) {
				const uint32_t w0 = seqs[2u * k], ofv = seqs[2u * k + 1u] + 3u;
				const uint32_t ll = w0 & 0xFFFFFFu, mb = w0 >> 24;
				const uint32_t llc = zse_ll_code(ll), mlc = zse_ml_code(mb), ofc = 31u - (uint32_t)__clz(ofv);
				uint32_t nb = (s_of + c_zse.of_dnb[ofc]) >> 16;
				bw.add(s_of, nb);
				s_of = c_zse.of_tab[(int32_t)(s_of >> nb) + c_zse.of_dfs[ofc]];
				nb = (s_ml + c_zse.ml_dnb[mlc]) >> 16;
				bw.add(s_ml, nb);
				s_ml = c_zse.ml_tab[(int32_t)(s_ml >> nb) + c_zse.ml_dfs[mlc]];
				nb = (s_ll + c_zse.ll_dnb[llc]) >> 16;
				bw.add(s_ll, nb);
				s_ll = c_zse.ll_tab[(int32_t)(s_ll >> nb) + c_zse.ll_dfs[llc]];
				bw.flush();   // (<= 17 bits of states before, <= 38 bits of extras now: fits the 64-bit accumulator)
				bw.add(ll, c_zse.ll_bits[llc]);
				bw.add(mb, c_zse.ml_bits[mlc]);
				bw.add(ofv, ofc);
				bw.flush();
			}
			bw.add(s_ml, 6);
			bw.add(s_of, 5);
			bw.add(s_ll, 6);
			bw.add(1u, 1);   // end mark: the decoder finds the last set bit
			bw.flush();
			if (bw.n && bw.p < bw.limit) {
				*bw.p++ = (uint8_t)bw.acc;
				bw.n = 0;
			}
			if (bw.n == 0u && bw.p < bw.limit) {
				csize = (uint32_t)(bw.p - (blk + 3u));
				const uint32_t bh = (last ? 1u : 0u) | (2u << 1) | (csize << 3);   // Compressed_Block
				blk[0] = (uint8_t)bh;
				blk[1] = (uint8_t)(bh >> 8);
				blk[2] = (uint8_t)(bh >> 16);
				blk[3] = (uint8_t)((3u << 2) | ((nlit & 15u) << 4));   // Raw_Literals_Block, Size_Format 11 (20-bit size)
				blk[4] = (uint8_t)(nlit >> 4);
				blk[5] = (uint8_t)(nlit >> 12);
			}
		}
		csize = __shfl_sync(0xFFFFFFFFu, csize, 0);
	}
	if (csize == 0u) {
		// Raw_Block: header + the chunk's bytes
		if (lane == 0) {
			const uint32_t bh = (last ? 1u : 0u) | (n << 3);
			blk[0] = (uint8_t)bh;
			blk[1] = (uint8_t)(bh >> 8);
			blk[2] = (uint8_t)(bh >> 16);
		}
		tile_copy<32>(blk + 3, src, n, lane);
		csize = n;
	}
	__syncwarp();
	return hdr + 3u + csize;
}
