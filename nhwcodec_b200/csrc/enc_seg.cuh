// enc_seg.cuh -- segment-parallel forms of the two byte-stream stages:
//   peephole passes over the luma scan (encoder/nhw_encoder.c:2136-2252)
//   statistics + code emission of wavlts2packet (encoder/compress_pixel.c:81-107, 279-361)
//
// Both are forward scans whose only state is "where did the current run of zeros (byte 128)
// start" plus a few skip windows, so a stream is cut into fixed-size segments, one thread per
// segment; a thread owns every token that STARTS in its segment (a run of zeros is one token
// group owned by the segment where it starts, however far it reaches).  Run lengths are found
// by walking inside the segment and hopping over all-zero segments through a per-segment
// summary.  The functions are plain (no barriers/atomics inside) so the host harness can run
// the identical per-segment logic sequentially; the kernels add the scans and atomics.
#pragma once
#include "enc_ll_par.cuh"

#define SEG_THREADS 256

// Zero runs are measured on a bitmap of the stream (bit set = byte is not the zero symbol 128)
// instead of byte by byte: the streams are mostly zeros and a run of a thousand zeros is 32 words.
struct NzBits {
	const uint32_t *w;      // bit k of word j describes stream byte p1 + 32*j + k
	int p1, n;              // first stream position covered, number of bits (multiple of 1024)
	const uint32_t *sum;    // second level: bit k of word j = "w[32*j + k] != 0" (n / 1024 words), or NULL
};

NHW_HD int nhw_ctz(uint32_t m)
{
#ifdef __CUDA_ARCH__
	return __ffs((int)m) - 1;
#else
	return __builtin_ctz(m);
#endif
}
NHW_HD int nhw_clz(uint32_t m)
{
#ifdef __CUDA_ARCH__
	return __clz((int)m);
#else
	return __builtin_clz(m);
#endif
}

// nonzero mask of 4 stream bytes packed in a word -> 4 bits
NHW_HD uint32_t nz_mask4(uint32_t w)
{
	const uint32_t x = w ^ 0x80808080u;
	const uint32_t t = ((x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u) >> 7;   // 0x01 per nonzero byte
	return ((t * 0x00204081u) >> 21) & 15u;
}

// first position >= i (i >= p1) holding a non-zero byte, `end` if there is none before it
NHW_HD int nz_next(const NzBits &b, int i, int end)
{
	const int k = i - b.p1, nw = b.n >> 5;
	if (k >= b.n) return end;
	int wi = k >> 5;
	uint32_t m = b.w[wi] & (0xffffffffu << (k & 31));
	if (!m && b.sum) {
		// hop over all-zero words through the summary level
		if (++wi >= nw) return end;
		int si = wi >> 5;
		const int ns = nw >> 5;
		uint32_t ms = b.sum[si] & (0xffffffffu << (wi & 31));
		while (!ms) {
			if (++si >= ns) return end;
			ms = b.sum[si];
		}
		wi = (si << 5) + nhw_ctz(ms);
		m = b.w[wi];
	}
	while (!m) {
		if (++wi >= nw) return end;
		m = b.w[wi];
	}
	const int r = b.p1 + (wi << 5) + nhw_ctz(m);
	return r < end ? r : end;
}

// last position <= i holding a non-zero byte, p1 - 1 if there is none
NHW_HD int nz_prev(const NzBits &b, int i)
{
	const int k = i - b.p1;
	if (k < 0) return b.p1 - 1;
	int wi = k >> 5;
	uint32_t m = b.w[wi] & (0xffffffffu >> (31 - (k & 31)));
	if (!m && b.sum) {
		if (--wi < 0) return b.p1 - 1;
		int si = wi >> 5;
		uint32_t ms = b.sum[si] & (0xffffffffu >> (31 - (wi & 31)));
		while (!ms) {
			if (--si < 0) return b.p1 - 1;
			ms = b.sum[si];
		}
		wi = (si << 5) + 31 - nhw_clz(ms);
		m = b.w[wi];
	}
	while (!m) {
		if (--wi < 0) return b.p1 - 1;
		m = b.w[wi];
	}
	return b.p1 + (wi << 5) + 31 - nhw_clz(m);
}

struct SegStream {
	const uint8_t *s;       // whole scan buffer (im.scan)
	int p1, p2;             // [p1, p2) is the stream
	int S;                  // segment size, (p2-p1)/SEG_THREADS
	NzBits nz;              // non-zero bitmap of [p1, p2)
	const int *bnd;         // optional: segment t = [bnd[t], bnd[t+1]) instead of equal sizes (work-balanced cut)
};
NHW_HD int seg_start(const SegStream &st, int t) { return st.bnd ? st.bnd[t] : st.p1 + t * st.S; }

// number of consecutive 128s starting at i (s[i]==128), never past p2
NHW_HD int seg_run_len(const SegStream &st, int i) { return nz_next(st.nz, i, st.p2) - i; }

// =====================================================================================
// peephole
// =====================================================================================
// pass A candidate: a +-8 byte, three zeros, a +-8 byte (all on the un-edited stream)
NHW_HD bool peep_pair_candidate(const uint8_t *s, int i, int N)
{
	if (i < 0 || i >= N - 4) return false;
	const int x = s[i], y = s[i + 4];
	return (x == 136 || x == 120) && s[i + 1] == 128 && s[i + 2] == 128 && s[i + 3] == 128 && (y == 136 || y == 120);
}

// pass A for the chain headed at i (candidate whose predecessor i-4 is not one): the reference
// pairs greedily left to right, so chain members alternate merged / swallowed.
NHW_HD void peep_merge_chain(uint8_t *s, int i, int N)
{
	for (;;) {
		const int x = s[i], y = s[i + 4];
		const bool next_is_candidate = peep_pair_candidate(s, i + 4, N);
		const bool after_next = next_is_candidate && peep_pair_candidate(s, i + 8, N);
		s[i] = (uint8_t)(132 + (x == 120 ? 2 : 0) + (y == 120 ? 1 : 0));
		s[i + 4] = 201;
		if (!after_next) return;
		i += 8;
	}
}

// passes B and C for position i of the stream `d` (= after pass A and the forced zero ends),
// result byte returned; sel1/sel2 tell which select counter to bump.
// B: isolated / paired +-8 bytes lose their code and keep only a sign bit (153/155, 157/159).
// C: a 153/155 that ends a zero run too long for one run code goes back to a coded byte.
NHW_HD int peep_select_byte(const uint8_t *d, const NzBits &nz, int i, int N, int &sel1, int &sel2)
{
	sel1 = sel2 = 0;
	int v = d[i];
	if (i < 4 || i >= N - 4) {
		// position N-4 can still be rewritten as the right neighbour of a pair starting at N-5
		if (i != N - 4) return v;
	}
	// (1) am I the right half of a pair flagged at i-1 ?
	if (i - 1 >= 4 && i - 1 < N - 4) {
		const int p = i - 1, c = d[p];
		if ((c == 136 || c == 120) && (v == 120 || v == 136)) {
			const bool zl = d[p - 1] == 128;
			const bool b1 = d[p + 2] == 128 && zl && d[p - 2] == 128 && d[p - 3] == 128 && d[p - 4] == 128;
			const bool b2 = zl && d[p + 2] == 128 && d[p + 3] == 128 && d[p + 4] == 128 && d[p + 5] == 128;
			if (b1 || b2) { sel2 = 1; return v == 120 ? 157 : 159; }
		}
	}
	if (i >= N - 4) return v;
	// (2) my own decision
	if (v != 136 && v != 120) return v;
	const bool nxt = d[i + 1] == 120 || d[i + 1] == 136;
	const bool zl = d[i - 1] == 128;
	if (d[i + 2] == 128 && nxt && zl && d[i - 2] == 128 && d[i - 3] == 128 && d[i - 4] == 128) return v;   // pair: my right neighbour changes, not me
	if (zl && nxt && d[i + 2] == 128 && d[i + 3] == 128 && d[i + 4] == 128 && d[i + 5] == 128) return v;
	const bool b3 = zl && d[i - 2] == 128 && d[i - 3] == 128 && d[i - 4] == 128 && d[i + 1] == 128;
	const bool b4 = zl && d[i + 1] == 128 && d[i + 2] == 128 && d[i + 3] == 128 && d[i + 4] == 128;
	if (!(b3 || b4)) return v;
	sel1 = 1;
	int out = v == 136 ? 153 : 155;
	// pass C: zero run [p, i) of length L ends right before me
	const int L = i - 1 - nz_prev(nz, i - 1);
	if (L >= 253) {
		const int p = i - L;
		int m = 0, i_last = 0;
		if (i - 2 >= p + 255) { m = (i - 2 - (p + 255)) / 254 + 1; i_last = p + 255 + 254 * (m - 1); }
		const int count_final = m > 0 ? i - i_last : L - 1;
		if ((m > 0 && i_last >= i - 3) || count_final >= 252) out = out == 153 ? 124 : 123;
	}
	return out;
}

// =====================================================================================
// entropy stage
// =====================================================================================
// first position of segment t that starts a token owned by t
NHW_HD int seg_first_token(const SegStream &st, int t)
{
	const int start = seg_start(st, t);
	int i = start;
	for (int k = 1; k <= 4; k++)
		if (i - k >= st.p1 && st.s[i - k] > 131 && st.s[i - k] < 136) { i = i - k + 5; break; }
	if (i < st.p2 && st.s[i] == 128 && i > st.p1 && st.s[i - 1] == 128) i += seg_run_len(st, i);
	return i;
}

// statistics of segment t (compress_pixel.c:81-107): every byte of the segment that is not a
// zero, plus every zero run that starts in it, chunked by the ">255 -> 254 + rest" rule.
template <typename Add>
NHW_HD void seg_stats(const SegStream &st, int t, Add add /* (is_run, index) */)
{
	const int start = seg_start(st, t), end = seg_start(st, t + 1);
	const int last = st.p2 - 1;   // the reference's loop never starts a token at the last byte
	int i = start;
	if (i < end && st.s[i] == 128 && i > st.p1 && st.s[i - 1] == 128) i += seg_run_len(st, i);
	while (i < end && i < last) {
		if (st.s[i] != 128) { add(false, (int)st.s[i]); i++; continue; }
		int L = seg_run_len(st, i);
		i += L;
		while (L > 255) { add(true, 254); L -= 254; }
		if (L == 1) add(false, 128); else add(true, L);
	}
}

struct SegCount { int bits, n1, n2; };

// Emission of segment t (compress_pixel.c:279-361).  emit(rank_code, len) is called for every
// code in order, bit1/bit2 for the select bits.  rank tables: sym_rank[byte], run_rank[len].
template <typename Emit, typename Bit1, typename Bit2>
NHW_HD int seg_emit(const SegStream &st, int t, const int *sym_rank, const int *run_rank, int select, bool zone,
                    Emit emit, Bit1 bit1, Bit2 bit2)
{
	const int end = seg_start(st, t + 1), last = st.p2 - 1;
	int i = seg_first_token(st, t);
	auto put = [&](int pos) -> int {
		pos &= 0xffff;
		if (pos >= 110 && pos < 174 && zone) { emit((1u << 6) | (uint32_t)(pos - 110), 15); return 0; }
		if (pos >= 174 && zone) pos -= 64;
		if (pos >= NHW_CODE_DEPTH) return 1;
		emit(nhw_code_bits[pos], (int)nhw_code_len[pos]);
		return 0;
	};
	int bad = 0;
	while (i < end && i < last) {
		const int v = st.s[i];
		if (v == 153) { bit1(0); i++; }
		else if (v == 155) { bit1(1); i++; }
		else if (v == 157) { bit2(0); i++; }
		else if (v == 159) { bit2(1); i++; }
		else if (v != 128) { bad |= put(sym_rank[v]); i += (v > 131 && v < 136) ? 5 : 1; }
		else {
			int L = seg_run_len(st, i);
			i += L;
			while (L > 255) { bad |= put(run_rank[254]); L -= 254; }
			if (L == 1) bad |= put(sym_rank[128]);
			else if (L < select) { for (int k = 0; k < L; k++) bad |= put(sym_rank[128]); }
			else bad |= put(run_rank[L]);
		}
	}
	return bad;
}

// OR `len` bits of `code` into a big-endian-within-word bit string at absolute bit `off`
template <typename Or>
NHW_HD void seg_put_bits(Or or_word, int word0, long off, uint32_t code, int len)
{
	const int w = word0 + (int)(off >> 5), used = (int)(off & 31) + len;
	if (used <= 32) or_word(w, code << (32 - used));
	else {
		const int spill = used - 32;
		or_word(w, code >> spill);
		or_word(w + 1, (code & ((1u << spill) - 1u)) << (32 - spill));
	}
}
