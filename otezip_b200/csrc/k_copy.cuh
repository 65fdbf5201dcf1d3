// k_copy.cuh — cooperative byte-range copies and the two copy-shaped decoders:
//   STORE             /root/reference/src/lib/otezip.c:481-487
//   method 93         /root/reference/src/lib/zstd.inc.c:479-705 (the reference's raw-block
//                     container — magic 28 B5 2F FD, one descriptor byte, [flags][len16] blocks;
//                     SURVEY.md Appendix C), driven as otezip.c:535-561 drives it.
#pragma once
#include "otz_common.cuh"
#include "k_crc32.cuh"

// Copy n bytes src -> dst with G cooperating lanes (lane in [0,G)); any alignment on either side.
// Stores are 16-byte vectors on the destination grid; loads are aligned 32-bit words funnel-shifted
// into place (or 16-byte vectors when both sides share their alignment).  No overlap allowed.
template <int G>
__device__ __forceinline__ void tile_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint64_t n, int lane) {
	uint64_t h = (16 - (reinterpret_cast<uint64_t>(dst) & 15)) & 15;
	if (h > n) {
		h = n;
	}
	for (uint64_t i = lane; i < h; i += G) {
		dst[i] = src[i];
	}
	dst += h;
	src += h;
	n -= h;
	const uint64_t nv = n >> 4;
	const uint64_t sa = reinterpret_cast<uint64_t>(src);
	if ((sa & 15) == 0) {
		const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
		uint4 *d4 = reinterpret_cast<uint4 *>(dst);
		uint64_t v = lane;
		// software pipeline: the four vectors of trip t+1 are requested before those of trip t are stored,
		// so a lane keeps 64-128 bytes in flight (a warp 2-4 KiB) without more resident warps
		if (v + 3 * G < nv) {
			uint4 a = ld_stream16(s4 + v), b = ld_stream16(s4 + v + G), c = ld_stream16(s4 + v + 2 * G), d = ld_stream16(s4 + v + 3 * G);
			for (; v + 7 * G < nv; v += 4 * G) {
				const uint4 a2 = ld_stream16(s4 + v + 4 * G), b2 = ld_stream16(s4 + v + 5 * G), c2 = ld_stream16(s4 + v + 6 * G),
				            d2 = ld_stream16(s4 + v + 7 * G);
				d4[v] = a;
				d4[v + G] = b;
				d4[v + 2 * G] = c;
				d4[v + 3 * G] = d;
				a = a2;
				b = b2;
				c = c2;
				d = d2;
			}
			d4[v] = a;
			d4[v + G] = b;
			d4[v + 2 * G] = c;
			d4[v + 3 * G] = d;
			v += 4 * G;
		}
		for (; v < nv; v += G) {
			d4[v] = ld_stream16(s4 + v);
		}
	} else {
		const uint32_t sh = (uint32_t)(sa & 3) * 8;
		const uint32_t *w = reinterpret_cast<const uint32_t *>(sa & ~3ull);
		uint4 *d4 = reinterpret_cast<uint4 *>(dst);
		uint64_t v = lane;
		// 4 vectors per lane per trip (20 independent word loads), the next trip's loads in flight before this trip's stores
		if (v + 3 * G < nv) {
			uint32_t a[4][5];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t *p = w + 4 * (v + k * G);
#pragma unroll
				for (int i = 0; i < 4; i++) {
					a[k][i] = __ldg(p + i);
				}
				a[k][4] = sh ? __ldg(p + 4) : 0u;
			}
			for (; v + 7 * G < nv; v += 4 * G) {
				uint32_t n[4][5];
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const uint32_t *p = w + 4 * (v + (k + 4) * G);
#pragma unroll
					for (int i = 0; i < 4; i++) {
						n[k][i] = __ldg(p + i);
					}
					n[k][4] = sh ? __ldg(p + 4) : 0u;
				}
#pragma unroll
				for (int k = 0; k < 4; k++) {
					uint4 o;
					o.x = __funnelshift_r(a[k][0], a[k][1], sh);
					o.y = __funnelshift_r(a[k][1], a[k][2], sh);
					o.z = __funnelshift_r(a[k][2], a[k][3], sh);
					o.w = __funnelshift_r(a[k][3], a[k][4], sh);
					d4[v + k * G] = o;
#pragma unroll
					for (int i = 0; i < 5; i++) {
						a[k][i] = n[k][i];
					}
				}
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				uint4 o;
				o.x = __funnelshift_r(a[k][0], a[k][1], sh);
				o.y = __funnelshift_r(a[k][1], a[k][2], sh);
				o.z = __funnelshift_r(a[k][2], a[k][3], sh);
				o.w = __funnelshift_r(a[k][3], a[k][4], sh);
				d4[v + k * G] = o;
			}
			v += 4 * G;
		}
		for (; v < nv; v += G) {
			const uint32_t *p = w + 4 * v;
			uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3);
			uint32_t w4 = sh ? __ldg(p + 4) : 0u;
			uint4 o;
			o.x = __funnelshift_r(w0, w1, sh);
			o.y = __funnelshift_r(w1, w2, sh);
			o.z = __funnelshift_r(w2, w3, sh);
			o.w = __funnelshift_r(w3, w4, sh);
			d4[v] = o;
		}
	}
	const uint64_t done = nv << 4;
	for (uint64_t i = done + lane; i < n; i += G) {
		dst[i] = src[i];
	}
}

// STORE extract: one warp per chunk of the shared chunk list, copy only (the CRC is taken by k_crc_chunks from the
// output while it is still warm in L2; a copy-only kernel keeps four times as many warps, i.e. loads, in flight).
__global__ void __launch_bounds__(256, 4) k_store_copy(const uint8_t *__restrict__ archive, uint8_t *__restrict__ out,
	const otz_entry *__restrict__ ents, const OtzEntryState *__restrict__ est, const int32_t *__restrict__ status,
	const OtzCrcChunk *__restrict__ chunks, uint32_t n_chunks) {
	const uint32_t warps_per_cta = blockDim.x >> 5;
	const uint32_t total_warps = gridDim.x * warps_per_cta;
	const int lane = threadIdx.x & 31;
	for (uint32_t c = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); c < n_chunks; c += total_warps) {
		const OtzCrcChunk ck = chunks[c];
		const otz_entry e = ents[ck.entry];
		if (e.method != OTZ_M_STORE || OTZ_ST_CODE(status[ck.entry]) != OTZ_ST_OK) {
			continue;
		}
		const uint64_t off = (uint64_t)ck.chunk * OTZ_CRC_CHUNK;
		const uint64_t len = min((uint64_t)OTZ_CRC_CHUNK, (uint64_t)e.uncomp_size - off);
		tile_copy<32>(out + e.out_ofs + off, archive + est[ck.entry].data_ofs + off, len, lane);
	}
}

// Method 93, reference container: one warp per entry walks the block chain and copies payloads (software-pipelined
// 16-byte copies, 2-4 KiB in flight per warp).  The CRC is taken afterwards by k_crc_chunks from the output, most of
// which is still in L2: a copy-only kernel needs a quarter of the registers of the fused one, so four times as many
// warps keep loads in flight.
__global__ void __launch_bounds__(256, 4) k_zstdref(const uint8_t *__restrict__ archive, uint8_t *__restrict__ out,
	const otz_entry *__restrict__ ents, const OtzEntryState *__restrict__ est, int32_t *__restrict__ status,
	const uint32_t *__restrict__ list, uint32_t n_list) {
	const uint32_t warps_per_cta = blockDim.x >> 5;
	const uint32_t total_warps = gridDim.x * warps_per_cta;
	const int lane = threadIdx.x & 31;
	for (uint32_t k = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); k < n_list; k += total_warps) {
		const uint32_t ei = list[k];
		if (OTZ_ST_CODE(status[ei]) != OTZ_ST_OK) {
			continue;
		}
		const otz_entry e = ents[ei];
		const uint8_t *in = archive + est[ei].data_ofs;
		uint8_t *dst = out + e.out_ofs;
		const uint32_t n = e.comp_size, cap = e.uncomp_size;
		int32_t st = OTZ_ST_OK;
		uint32_t ip = 5, op = 0;
		if (n < 5) {
			st = OTZ_ST_TRUNCATED;  // zstd:490-492
		} else if (ld_le32(in) != 0xFD2FB528u) {
			st = OTZ_ST_DATA;  // zstd:495-498
		} else {
			for (;;) {
				if (n - ip < 3) {
					st = OTZ_ST_TRUNCATED;  // zstd:511 falls out with Z_OK / :698-700 Z_BUF_ERROR
					break;
				}
				const uint32_t h = in[ip];
				const uint32_t bsz = in[ip + 1] | (in[ip + 2] << 8);  // zstd:559-560
				const uint32_t type = (h >> 1) & 3;
				ip += 3;
				if (type != 0 && type != 2) {
					st = OTZ_ST_DATA;  // zstd:689-692
					break;
				}
				if (n - ip < bsz) {
					st = OTZ_ST_TRUNCATED;  // zstd:570-576, :635-641
					break;
				}
				if (type == 2 && bsz == 0) {
					st = OTZ_ST_DATA;  // zstd:189-191 -> :647-649
					break;
				}
				if (cap - op < bsz) {
					st = OTZ_ST_OVERFLOW;  // zstd:608-632, :676-683 then :546-548
					break;
				}
				tile_copy<32>(dst + op, in + ip, bsz, lane);
				ip += bsz;
				op += bsz;
				if (h & 1) {
					break;  // zstd:695-697
				}
			}
			if (st == OTZ_ST_OK && op != cap) {
				st = OTZ_ST_SIZE;  // otezip.c:555
			}
		}
		if (lane == 0) {
			if (st != OTZ_ST_OK) {
				// not a consistent reference container: if it carries the Zstandard magic, k_zstd_lit / k_zstd_seq try it as an
				// RFC 8878 frame (the reference itself would reject it either way)
				// (a frame, or a skippable frame in front of one: RFC 8878 3.1.2)
				const uint32_t mg = n >= 4 ? ld_le32(in) : 0u;
				status[ei] = (mg == 0xFD2FB528u || (mg & 0xFFFFFFF0u) == 0x184D2A50u) ? OTZ_ST_PENDING : st;
			}
		}
	}
}
