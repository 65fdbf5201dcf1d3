// spec_proto.cpp — CPU model of the warp-per-stream speculative DEFLATE tokenizer (k_inflate_spec).
//
// Design study, not product code: measures, on real zlib streams, how the "32 lanes start at guessed bit offsets
// and decode until they meet a position the next lane has visited" scheme behaves (lock-step steps per pass,
// iterations of the synchronisation loop) and checks that the tokens it yields equal those of a serial decode.
//
//   g++ -O2 -o /tmp/spec_proto tools/spec_proto.cpp -lz && /tmp/spec_proto <file with raw deflate streams>
//   input file: repeated { u32 comp_len, u32 uncomp_len, comp bytes }
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Tab {
	// full 15-bit tables, entry = sym | len << 16 (0 = invalid)
	std::vector<uint32_t> lit, dst;
};

static const uint16_t LBASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t LXB[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t DBASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t DXB[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

static uint64_t peek(const uint8_t *in, uint64_t nbits, uint64_t p) {
	uint64_t v = 0;
	const uint64_t b = p >> 3;
	for (int i = 0; i < 8; i++) {
		const uint64_t idx = b + i;
		v |= (uint64_t)(idx * 8 < nbits + 64 && idx < (nbits + 7) / 8 ? in[idx] : 0) << (8 * i);
	}
	return v >> (p & 7);
}

static bool build(const uint8_t *lens, int n, std::vector<uint32_t> &t) {
	t.assign(32768, 0);
	int cnt[16] = { 0 };
	for (int i = 0; i < n; i++) cnt[lens[i]]++;
	cnt[0] = 0;
	int code = 0, next[16];
	for (int l = 1; l <= 15; l++) {
		code = (code + cnt[l - 1]) << 1;
		next[l] = code;
	}
	for (int s = 0; s < n; s++) {
		const int l = lens[s];
		if (!l) continue;
		int c = next[l]++;
		int rev = 0;
		for (int i = 0; i < l; i++) rev |= ((c >> i) & 1) << (l - 1 - i);
		for (int x = rev; x < 32768; x += 1 << l) t[x] = (uint32_t)s | (l << 16) | 0x80000000u;
	}
	return true;
}

enum { K_LIT, K_MATCH, K_EOB, K_BAD };
struct Step { int kind; uint32_t val, len, dist; uint64_t next; };

static Step step(const Tab &T, const uint8_t *in, uint64_t nbits, uint64_t p) {
	Step s{ K_BAD, 0, 0, 0, p };
	uint64_t w = peek(in, nbits, p);
	uint32_t e = T.lit[w & 32767];
	if (!e) return s;
	uint32_t sym = e & 0xFFFF, l = (e >> 16) & 15;
	p += l;
	w >>= l;
	if (sym < 256) { s.kind = K_LIT; s.val = sym; s.next = p; return s; }
	if (sym == 256) { s.kind = K_EOB; s.next = p; return s; }
	if (sym > 285) return s;
	const uint32_t v = sym - 257;
	s.len = LBASE[v] + (uint32_t)(w & ((1u << LXB[v]) - 1));
	p += LXB[v];
	w = peek(in, nbits, p);
	e = T.dst[w & 32767];
	if (!e) return s;
	sym = e & 0xFFFF; l = (e >> 16) & 15;
	if (sym > 29) return s;
	p += l; w >>= l;
	s.dist = DBASE[sym] + (uint32_t)(w & ((1u << DXB[sym]) - 1));
	p += DXB[sym];
	s.kind = K_MATCH; s.next = p;
	return s;
}

struct Stats {
	uint64_t rounds = 0, p1_steps = 0, p2_steps = 0, p3_steps = 0, sum_sym = 0, iters = 0, p2_rounds_gt1 = 0, unsynced = 0, lanes = 0;
	uint64_t hist_iter[40] = { 0 };
};

int main(int argc, char **argv) {
	if (argc < 2) return 1;
	const uint32_t S = argc > 2 ? atoi(argv[2]) : 256;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 1;
	Stats st;
	uint64_t total_out = 0, streams = 0, blocks = 0;
	for (;;) {
		uint32_t hdr[2];
		if (fread(hdr, 4, 2, f) != 2) break;
		std::vector<uint8_t> comp(hdr[0] + 16);
		if (fread(comp.data(), 1, hdr[0], f) != hdr[0]) break;
		const uint8_t *in = comp.data();
		const uint64_t nbits = (uint64_t)hdr[0] * 8;
		uint64_t p = 0;
		std::vector<uint8_t> out;
		streams++;
		bool final_blk = false;
		while (!final_blk) {
			uint64_t w = peek(in, nbits, p);
			final_blk = w & 1;
			const uint32_t bt = (w >> 1) & 3;
			p += 3;
			if (bt == 0) {
				p = (p + 7) & ~7ull;
				const uint32_t len = in[p / 8] | (in[p / 8 + 1] << 8);
				p += 32;
				for (uint32_t i = 0; i < len; i++) out.push_back(in[p / 8 + i]);
				p += 8ull * len;
				continue;
			}
			Tab T;
			uint8_t lens[320] = { 0 };
			uint32_t hlit = 288, hdist = 32;
			if (bt == 1) {
				for (int i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
				for (int i = 0; i < 32; i++) lens[288 + i] = 5;
			} else {
				w = peek(in, nbits, p);
				hlit = (w & 31) + 257; hdist = ((w >> 5) & 31) + 1;
				const uint32_t hclen = ((w >> 10) & 15) + 4;
				p += 14;
				static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
				uint8_t cl[19] = { 0 };
				for (uint32_t i = 0; i < hclen; i++) { cl[order[i]] = peek(in, nbits, p) & 7; p += 3; }
				std::vector<uint32_t> ct;
				build(cl, 19, ct);
				uint32_t idx = 0;
				while (idx < hlit + hdist) {
					w = peek(in, nbits, p);
					const uint32_t e = ct[w & 32767];
					const uint32_t sym = e & 0xFFFF, l = (e >> 16) & 15;
					p += l; w >>= l;
					if (sym < 16) lens[idx++] = sym;
					else if (sym == 16) { uint32_t r = 3 + (w & 3); p += 2; while (r--) { lens[idx] = lens[idx - 1]; idx++; } }
					else if (sym == 17) { uint32_t r = 3 + (w & 7); p += 3; idx += r; }
					else { uint32_t r = 11 + (w & 127); p += 7; idx += r; }
				}
			}
			build(lens, hlit, T.lit);
			build(lens + hlit, hdist, T.dst);
			blocks++;
			// ---- rounds of 32 lanes x S bits
			bool eob = false;
			while (!eob) {
				const uint64_t R = p;
				uint64_t E[32], Xs[32];      // spec exit, spec start
				int Ek[32];                  // exit kind: 0 normal, 1 EOB, 2 bad
				std::vector<std::vector<uint8_t>> vis(32, std::vector<uint8_t>(S, 0));
				uint32_t maxsteps = 0;
				for (int j = 0; j < 32; j++) {
					uint64_t q = R + (uint64_t)j * S;
					const uint64_t end = R + (uint64_t)(j + 1) * S;
					Xs[j] = q;
					uint32_t n = 0;
					Ek[j] = 0;
					while (q < end) {
						vis[j][q - (R + (uint64_t)j * S)] = 1;
						if (q >= nbits) { Ek[j] = 2; break; }
						Step s = step(T, in, nbits, q);
						n++;
						if (s.kind == K_BAD) { Ek[j] = 2; break; }
						q = s.next;
						if (s.kind == K_EOB) { Ek[j] = 1; break; }
					}
					E[j] = q;
					maxsteps = std::max(maxsteps, n);
				}
				st.p1_steps += maxsteps;
				// ---- sync loop
				uint64_t start[33], outp[32];
				int outk[32];
				bool need[32];
				for (int j = 0; j < 32; j++) { outp[j] = E[j]; outk[j] = Ek[j]; start[j] = Xs[j]; need[j] = false; }
				uint32_t iters = 0;
				for (;;) {
					bool any = false;
					for (int j = 1; j < 32; j++) {
						need[j] = false;
						if (outk[j - 1] == 0 && outp[j - 1] != start[j]) { need[j] = true; any = true; }
					}
					if (!any) break;
					iters++;
					uint32_t ms = 0;
					// all lanes in parallel from the previous iteration's outputs
					uint64_t nout[32]; int nk[32];
					for (int j = 1; j < 32; j++) {
						nout[j] = outp[j]; nk[j] = outk[j];
						if (!need[j]) continue;
						uint64_t q = outp[j - 1];
						start[j] = q;
						const uint64_t base = R + (uint64_t)j * S, end = base + S;
						uint32_t n = 0;
						int k = 0;
						bool hit = false;
						while (q < end) {
							if (vis[j][q - base]) { hit = true; break; }
							if (q >= nbits) { k = 2; break; }
							Step s = step(T, in, nbits, q);
							n++;
							if (s.kind == K_BAD) { k = 2; break; }
							q = s.next;
							if (s.kind == K_EOB) { k = 1; break; }
						}
						if (hit) { nout[j] = E[j]; nk[j] = Ek[j]; }
						else { nout[j] = q; nk[j] = k; st.unsynced++; }
						ms = std::max(ms, n);
					}
					for (int j = 1; j < 32; j++) { outp[j] = nout[j]; outk[j] = nk[j]; }
					st.p2_steps += ms;
				}
				st.iters += iters;
				st.hist_iter[std::min<uint32_t>(iters, 39)]++;
				// ---- valid lanes: chain from lane 0 while exits are normal
				int nvalid = 32;
				for (int j = 0; j < 32; j++) {
					if (outk[j] != 0) { nvalid = j + 1; break; }
				}
				// ---- emit (serial order here; count pass and emit pass have the same lock-step cost)
				uint32_t ms = 0;
				for (int j = 0; j < nvalid; j++) {
					uint64_t q = j == 0 ? R : outp[j - 1];
					const uint64_t end = R + (uint64_t)(j + 1) * S;
					uint32_t n = 0;
					while (q < end) {
						Step s = step(T, in, nbits, q);
						n++;
						if (s.kind == K_BAD) { fprintf(stderr, "bad symbol in verified chain\n"); return 2; }
						q = s.next;
						if (s.kind == K_EOB) { eob = true; break; }
						if (s.kind == K_LIT) out.push_back((uint8_t)s.val);
						else {
							if (s.dist > out.size()) { fprintf(stderr, "dist too far\n"); return 2; }
							for (uint32_t i = 0; i < s.len; i++) out.push_back(out[out.size() - s.dist]);
						}
					}
					if (q != outp[j] && !(eob)) { fprintf(stderr, "chain mismatch\n"); return 2; }
					st.sum_sym += n;
					ms = std::max(ms, n);
					p = q;
					if (eob) break;
				}
				st.p3_steps += ms;
				st.rounds++;
				st.lanes += nvalid;
			}
		}
		if (out.size() != hdr[1]) { fprintf(stderr, "stream %llu: size %zu != %u\n", (unsigned long long)streams, out.size(), hdr[1]); return 2; }
		total_out += out.size();
	}
	printf("S=%u streams=%llu blocks=%llu out=%llu rounds=%llu\n", S, (unsigned long long)streams, (unsigned long long)blocks, (unsigned long long)total_out, (unsigned long long)st.rounds);
	printf("per round: pass1 lock-step steps %.1f, sync steps %.1f (iters %.2f), emit steps %.1f, symbols %.1f (%.1f per valid lane), valid lanes %.1f, unsynced lane-iters %.3f\n",
		(double)st.p1_steps / st.rounds, (double)st.p2_steps / st.rounds, (double)st.iters / st.rounds, (double)st.p3_steps / st.rounds,
		(double)st.sum_sym / st.rounds, (double)st.sum_sym / st.lanes, (double)st.lanes / st.rounds, (double)st.unsynced / st.rounds);
	printf("bytes out per symbol %.2f; lock-step steps per output KB (p1 + sync + 2 x emit) %.1f\n", (double)total_out / st.sum_sym,
		(double)(st.p1_steps + st.p2_steps + 2 * st.p3_steps) / (total_out / 1024.0));
	printf("iteration histogram:");
	for (int i = 0; i < 40; i++) if (st.hist_iter[i]) printf(" %d:%llu", i, (unsigned long long)st.hist_iter[i]);
	printf("\n");
	return 0;
}
