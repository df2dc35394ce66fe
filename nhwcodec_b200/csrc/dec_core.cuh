// dec_core.cuh -- decoder stage functions (host/device), q17..q23.
//
//   dec_ll_dpcm          : LL byte decode of parse_file          decoder/nhw_decoder.c:1663-2026
//   dec_prefix_luma/_chroma : retrieve_pixel_Y_comp / _UV_comp   decoder/compress_pixel.c:49-444, 446-641
//   dec_expand_list      : res1/res3/res5 position lists          decoder/nhw_decoder.c:93-491
//   dec_luma_* / dec_chroma_* : inline stages of decode_image     decoder/nhw_decoder.c:71-1474
//   ycc_to_rgb           : write_image_bmp                        decoder/nhw_decoder_cli.c:108-291
//
// The prefix decoder does not use the reference's hand-built tables (decoder/tables.h): the code is
// static (rank r <-> nhw_code_bits[r], nhw_code_len[r], the same pairs the encoder writes), so a
// rank is found by matching the next bits against those pairs; the zone escape (nine bits
// 000000001 + 6 bits) is checked first when the stream uses it.
#pragma once
#include "enc_seg.cuh"

struct DecDesc {   // one per image, filled on the host from the .nhw header
	int32_t status, quality, byte0;
	int32_t size_tree1, size_tree2, size_data1, size_data2, tree_end, exw_Y_end;
	int32_t res1_len, res1_bit_len, res3_len, res3_bit_len, res4_len, res5_len, res5_bit_len;
	int32_t select1, select2, highres_comp_len, end_ch_res;
	int32_t res6_len, res6_bit_len, char_res1_len, qsetting3_len;   // q22/q23
	uint32_t off_tree1, off_tree2, off_exw, off_res1, off_res1_bit, off_res1_word, off_res4, off_res3, off_res3_bit,
	    off_res3_word, off_res5, off_res5_bit, off_res5_word, off_sel1, off_sel2, off_u64, off_v64, off_highres,
	    off_ch_res, off_words;
	uint32_t off_res6, off_res6_bit, off_res6_word, off_char_res1, off_qsetting3;
	uint32_t blob_len;
};

struct DecImg {
	const uint8_t *blob;     // the .nhw bytes
	const DecDesc *d;
	int16_t *proc, *jpeg, *aux;       // luma planes (512x512)
	int16_t *uvcoef;                  // im_nhw3: 131072 interleaved chroma coefficients
	int16_t *cproc, *cjpeg, *caux;    // chroma planes of one component (256x256)
	uint8_t *res_comp;                // 24577 LL bytes
	uint16_t *list[8];                // expanded position lists (res1 -,+ ; res5 -,+ ; res3 x4)
	int32_t *list_len;                // 8 lengths + [8] = stale `count`, [9] = edge-flag count, [10] = exw split, [11],[12] = hq lists
	uint32_t *hq_list[2];             // q22/q23: res6 positions, -32 then +32 (flat indices into the half-synthesised plane)
	uint32_t *mbits;                  // luma marker bitmap (512 rows x 16 words), written by the inverse scan
	uint32_t *cmark;                  // chroma markers of this component: [0] = count, then (code << 24 | cell) entries
	uint16_t *flags;                  // edge-flag positions
	uint16_t *book;                   // rank -> (run<<8 | byte)
	const uint16_t *lut;              // primary table of the static prefix code (dec_build_lut)
	uint8_t *yuv;                     // Y, U, V u8 planes 512x512 each
};

NHW_HD int nhw_extra_value(int word) { return word >= 0 && word < 110 ? (int)nhw_extra_table[word] : 0; }

// ---- bit reader over the packed 32-bit words (stored little-endian, consumed MSB first).
// The stream may start at any byte address: words are fetched with aligned 32-bit loads and
// re-aligned with a funnel shift; a 64-bit buffer holds the next bits MSB first.
struct BitReader {
	const uint32_t *ap;      // aligned pointer at or below the first stream byte
	int sh;                  // 8 * (first byte address & 3)
	long widx;               // next aligned word to fetch
	uint32_t carry;          // ap[widx], already fetched
	unsigned long long buf;  // unread bits, MSB first
	int nb;                  // valid bits in buf
	long pos;                // bits consumed so far
	NHW_HD void init(const uint8_t *b)
	{
		const uintptr_t a = (uintptr_t)b;
		ap = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
		sh = (int)(a & 3) * 8;
		widx = 0;
		carry = ap[0];
		buf = 0;
		nb = 0;
		pos = 0;
		fill();
		fill();
	}
	NHW_HD void fill()       // append the next stream word (nb <= 32 on entry)
	{
		const uint32_t nxt = ap[++widx];
		const uint32_t w = sh ? ((carry >> sh) | (nxt << (32 - sh))) : carry;
		carry = nxt;
		buf |= (unsigned long long)w << (32 - nb);
		nb += 32;
	}
	NHW_HD uint32_t peek(int n) const { return (uint32_t)(buf >> (64 - n)); }   // next n (<= 32) bits
	NHW_HD void skip(int n)
	{
		buf <<= n;
		nb -= n;
		pos += n;
		if (nb <= 32) fill();
	}
};

// Decode tables of the static prefix code, two levels.  Level 1 is indexed by the next 12 bits:
// (length << 10) | rank for every code of at most 12 bits; 0x8000 | k where longer codes (13..20
// bits, ranks NHW_LONG_FIRST..) start, k selecting a 256-entry level-2 table indexed by the 8 bits
// that follow; 0 = no code.  Level-2 entries have the same (length << 10) | rank form.
#define NHW_LUT_BITS 12
#define NHW_LONG_FIRST 98
#define NHW_LUT2_TABLES 8
#define NHW_LUT_WORDS ((1 << NHW_LUT_BITS) + NHW_LUT2_TABLES * 256)
inline void dec_build_lut(uint16_t *lut /* NHW_LUT_WORDS */)
{
	for (int i = 0; i < NHW_LUT_WORDS; i++) lut[i] = 0;
	int ntab = 0;
	for (int r = 0; r < NHW_CODE_DEPTH; r++) {
		const int len = h_nhw_code_len[r];
		const uint32_t code = h_nhw_code_bits[r];
		if (len <= NHW_LUT_BITS) {
			const uint32_t first = code << (NHW_LUT_BITS - len);
			for (uint32_t k = 0; k < (1u << (NHW_LUT_BITS - len)); k++) lut[first + k] = (uint16_t)((len << 10) | r);
		} else {
			const uint32_t p = code >> (len - NHW_LUT_BITS);
			if (!lut[p]) lut[p] = (uint16_t)(0x8000 | ntab++);
			uint16_t *sub = lut + (1 << NHW_LUT_BITS) + (lut[p] & 255) * 256;
			const uint32_t rest = (code << (20 - len)) & 255u;   // the code's bits 13..20, left aligned
			for (uint32_t k = 0; k < (1u << (20 - len)); k++) sub[rest + k] = (uint16_t)((len << 10) | r);
		}
	}
}

// rank of the next code (static prefix code).  Returns -1 if nothing matches.
NHW_HD int dec_next_rank(BitReader &br, bool zone, const uint16_t *lut)
{
	const uint32_t v = br.peek(20);
	if (zone && (v >> 11) == 1) {   // 000000001 + 6 bits
		const int r = (int)((v >> 5) & 63) + 110;
		br.skip(15);
		return r;
	}
	int ent = lut[v >> (20 - NHW_LUT_BITS)];
	if (ent & 0x8000) ent = lut[(1 << NHW_LUT_BITS) + ((ent & 255) << 8) + (v & 255u)];
	if (!ent) return -1;
	const int r = ent & 1023;
	br.skip(ent >> 10);
	return (zone && r >= 110) ? r + 64 : r;
}

// ---- codebook: un-RLE, re-interleave, build rank -> symbol table (compress_pixel.c:92-118, 455-478)
NHW_HDN int dec_build_book(const uint8_t *tree, int size, int marker, int e_override, uint16_t *book, uint8_t *tmp /* 2*1024 */)
{
	uint8_t *flat = tmp, *inter = tmp + 1024;
	int e = 0;
	for (int i = 0; i < size && e < 1000; i++) {
		if (tree[i] == marker) {
			for (int j = 0; j < tree[i + 1] && e < 1000; j++) flat[e++] = (uint8_t)marker;
			i++;
		} else flat[e++] = tree[i];
	}
	if (e_override >= 0) e = e_override;
	for (int i = 0; i < 1024; i++) inter[i] = 0;
	int j = 0;
	for (int i = 0; i < e; i += 2) inter[i] = flat[j++];
	for (int i = 1; i < e; i += 2) inter[i] = flat[j++];
	int n = 0;
	for (int i = 0; i < e; i++) {
		if (marker == 3) {   // luma: 3 = "zero run" marker followed by the run length
			if (inter[i] == 3) { book[n++] = (uint16_t)((inter[i + 1] << 8) | 128); i++; }
			else book[n++] = (uint16_t)(256 | inter[i]);
		} else {             // chroma: even byte = (128, run length) pair, odd byte = symbol | 1
			if (!(inter[i] & 1)) { book[n++] = (uint16_t)((inter[i + 1] << 8) | inter[i]); i++; }
			else book[n++] = (uint16_t)(256 | (inter[i] & 0xfe));
		}
	}
	return n;
}

// ---- rank -> action.  What the decoders do with a symbol depends on the symbol alone (decoder/compress_pixel.c:153-441:
// a chain of up to fifteen comparisons per symbol, 479-641 for chroma), so it is worked out once per codebook entry instead
// of once per decoded symbol: the 16-bit book entries are widened IN PLACE (from the top down, entry i only overwrites
// entries 2i and 2i + 1) to
//   bits 0-1  kind: 0 literal   1 zero run   2 literal that arms the "mem2" rule (luma 136 / 120)   3 pair (luma 132..135)
//   bits 8-15 run length (kind 1);  bit 8: the pair's second value is -11 (kind 3)
//   bits 16-31 the literal / first value of the pair (int16)
// The serial decoders are latency chains: this takes a dependent table look-up and the comparison chain off every symbol.
#define NHW_ACT_ENTRIES 512
NHW_HD uint32_t dec_action_of(int sym, bool luma)
{
	const int word = sym & 0xff, run = sym >> 8;
	if (word == 0x80) return 1u | ((uint32_t)(run & 255) << 8);
	int kind = 0, v, second_neg = 0;
	const int x = word < 110 ? nhw_extra_value(word) : 0;
	if (luma) {
		if (word == 136) { v = 11; kind = 2; }
		else if (word == 120) { v = -11; kind = 2; }
		else if (word >= 132 && word <= 135) { v = word < 134 ? 11 : -11; kind = 3; second_neg = word & 1; }
		else if (word == 127) v = 1008;
		else if (word == 129) v = 1009;
		else if (word == 125) v = 1006;
		else if (word == 126) v = 1007;
		else if (word == 121) v = 1010;
		else if (word == 122) v = 1011;
		else if (word == 124) v = 11;
		else if (word == 123) v = -11;
		else if (x > 0) v = 123 + (x << 3);
		else if (x < 0) v = (x << 3) - 123;
		else v = word > 0x80 ? word - 125 : word - 131;
	} else {
		if (word < 110 && x > 0) v = 123 + (x << 3);
		else if (word < 110 && x < 0) v = (x << 3) - 123;
		else if (word >= 110 && word == 124) v = 5005;
		else if (word >= 110 && word == 126) v = 5006;
		else if (word >= 110 && word == 122) v = 5003;
		else if (word >= 110 && word == 130) v = 5004;
		else v = word > 0x80 ? word - 125 : word - 131;
	}
	return (uint32_t)kind | ((uint32_t)second_neg << 8) | ((uint32_t)(uint16_t)(int16_t)v << 16);
}
NHW_HDN const uint32_t *dec_build_actions(uint16_t *book /* >= 1024 entries */, bool luma)
{
	uint32_t *act = reinterpret_cast<uint32_t *>(book);
	for (int i = NHW_ACT_ENTRIES - 1; i >= 0; i--) {
		const int sym = book[i];
		act[i] = dec_action_of(sym, luma);
	}
	return act;
}

// ---- luma prefix decode + run / select-bit logic (compress_pixel.c:120-444)
NHW_HDN int dec_prefix_luma(const DecImg &im, int16_t *im3 /* 262144, zeroed */, const uint32_t *act /* dec_build_actions */)
{
	const DecDesc *d = im.d;
	const uint8_t *sel1 = im.blob + d->off_sel1, *sel2 = im.blob + d->off_sel2;
	const bool zone = d->byte0 < 4;
	BitReader br;
	br.init(im.blob + d->off_words);
	const long nbits = (long)d->size_data1 * 32;
	const int p1 = 262144;
	int e = 0, mem = 0, mem2 = 0, ac1 = 0, run_over = -257, t = 0, t2 = 0;
	// select bits past the section the header declares read as 0 (a well-formed stream never gets there)
	auto bit1 = [&](int k) { return (k >> 3) < d->select1 ? (sel1[k >> 3] >> (7 - (k & 7))) & 1 : 0; };
	auto bit2 = [&](int k) { return (k >> 3) < d->select2 ? (sel2[k >> 3] >> (7 - (k & 7))) & 1 : 0; };
	// The plane starts zeroed and is written at increasing positions only, so "is cell e-k still
	// zero" is answered from a shift register of the last 32 positions instead of reading it back:
	// bit j of hist = a non-zero value was stored at position e-1-j.
	uint32_t hist = 0;
	auto z = [&](int k) { return k >= 0 ? !((hist >> (e - 1 - k)) & 1u) : true; };   // k in [e-5, e-1]; below 0: zero guard
	// (the last symbol may run a few cells past the plane: those stores are dropped, the plane's guard band stays zero)
	auto put = [&](int v) { if (e < p1) im3[e] = (int16_t)v; e++; hist = (hist << 1) | (v != 0 ? 1u : 0u); };
	auto advance = [&](int n) { e += n; hist = n >= 32 ? 0u : hist << n; };
	while (br.pos < nbits + 64) {
		const int dec = dec_next_rank(br, zone, im.lut);
		if (dec < 0) return NHW_ERR_CODEBOOK_DEV;
		const uint32_t a = act[dec];
		if ((a & 3u) == 1u) {
			const int run = (int)((a >> 8) & 255u);
			mem++;
			if (mem2 == 1) {
				if (e >= 5 && z(e - 2) && z(e - 3) && z(e - 4) && z(e - 5)) { put(bit2(t2++) ? 11 : -11); }
				else if (run >= 4 && z(e - 2)) { put(bit2(t2++) ? 11 : -11); }
				mem2 = 0;
			} else if (mem == 2 && !ac1) {
				if (e >= 4 && z(e - 1) && z(e - 2) && z(e - 3) && z(e - 4) && (e + run - 257) >= run_over) {
					put(bit1(t++) ? -11 : 11);
					mem = 1;
				} else if (run >= 4 && e > 0 && z(e - 1) && !ac1 && (e + run - 257) >= run_over) {
					put(bit1(t++) ? -11 : 11);
					mem = 1;
				}
			} else if (run >= 4 && e > 0 && z(e - 1) && !ac1 && (e + run - 257) >= run_over) {
				put(bit1(t++) ? -11 : 11);
				mem = 1;
			}
			if (run == 254) { ac1 = 1; mem = 0; run_over = e; }
			else ac1 = 0;
			advance(run);
		} else {
			mem = 0; mem2 = 0; ac1 = 0;
			put((int)(int16_t)(a >> 16));
			if ((a & 3u) == 2u) mem2 = 1;
			else if ((a & 3u) == 3u) { advance(3); put((a & 0x100u) ? -11 : 11); }
		}
		if (e >= p1 - 1) return 0;
	}
	return NHW_ERR_CODEBOOK_DEV;   // ran out of bits
}

// ---- chroma prefix decode (compress_pixel.c:479-641): no zone, no select bits
NHW_HDN int dec_prefix_chroma(const DecImg &im, int16_t *im3 /* 131072, zeroed */, const uint32_t *act /* dec_build_actions */)
{
	const DecDesc *d = im.d;
	BitReader br;
	br.init(im.blob + d->off_words + 4 * (size_t)d->size_data1);
	const long nbits = (long)(d->size_data2 - d->size_data1) * 32;
	const int p1 = 131071;
	int e = 0;
	while (br.pos < nbits + 64) {
		const int dec = dec_next_rank(br, false, im.lut);
		if (dec < 0) return NHW_ERR_CODEBOOK_DEV;
		const uint32_t a = act[dec];
		if (a & 1u) e += (int)((a >> 8) & 255u);
		else {
			if (e < 131072) im3[e] = (int16_t)(a >> 16);   // (a run may carry e past the end: nothing is stored there)
			e++;
		}
		if (e >= p1 - 1) return 0;
	}
	return NHW_ERR_CODEBOOK_DEV;
}

// ---- LL bytes (parse_file, nhw_decoder.c:1663-2026).  All arithmetic is modulo 256 like the
// reference's unsigned char stores.
// Returns 0, or NHW_ERR_STREAM_DEV when the code runs past the section lengths the header declares.
NHW_HDN int dec_ll_dpcm(const DecImg &im)
{
	const DecDesc *d = im.d;
	const uint8_t *ch = im.blob + d->off_ch_res, *hr = im.blob + d->off_highres;
	const int ch_len = d->end_ch_res, hr_len = d->highres_comp_len;
	uint8_t *o = im.res_comp;
	const int q = d->quality, mode = d->byte0 & 3;
	int j = 1, i = 1, a = 0;
	// `last` mirrors o[j-1] (without the chroma LSBs added on the way out, see below), so that the
	// recurrence runs in registers instead of through memory
	uint8_t last = ch[0];
	o[0] = last;
	const uint8_t *ub = im.blob + d->off_u64, *vb = im.blob + d->off_v64;
	auto emit = [&](int v) {
		last = (uint8_t)v;
		uint8_t w = last;
		if (j >= 16384 && q > 15) {   // res_U_64 / res_V_64 LSB planes (nhw_decoder.c:1983-2026)
			const int k = (j - 16384) & 4095;
			const uint8_t *pl = j < 20480 ? ub : vb;
			w = (uint8_t)(w + (((pl[k >> 3] >> (7 - (k & 7))) & 1) << 1));
		}
		o[j++] = w;
	};
	auto rel = [&](int delta) { emit(last + delta); };
	auto triple = [&](int c0, int c1) {   // 3 deltas packed in two bytes (marker 64)
		rel((((c0 >> 1) & 31) << 1) - 32);
		rel(((((c0 & 1) << 3) | (c1 >> 5)) << 1) - 16);
		rel(((c1 & 31) << 1) - 32);
	};
	for (; j < 16384; i++) {
		if (i + 1 >= ch_len) return NHW_ERR_STREAM_DEV;   // every code may take a second byte
		const int c = ch[i];
		if (c >= 128) {
			if (q > 15) { if (a >= hr_len) return NHW_ERR_STREAM_DEV; emit(hr[a++]); }
			emit((c - 128) << 1);
		} else if (mode == 0) {
			if (c < 16) {
				const int run = (c >> 3) & 1;
				const uint8_t v = last;
				for (int e = 0; e < run + 2; e++) emit(v);
				const int k = c & 7;
				if (k == 1) rel(2);
				else if (k == 2) { rel(2); rel(-2); }
				else if (k == 3) { rel(2); rel(0); }
				else if (k == 4) { rel(-2); rel(2); }
				else if (k == 5) { rel(-2); rel(0); }
				else if (k == 6) rel(-2);
				else if (k == 7) rel(4);
			} else if (c < 32) {
				rel(c >= 24 ? 4 : 2);
				rel(((c & 7) << 1) - 8);
			} else if (c < 64) {
				const int x = c - 32;
				rel(((x >> 3) << 1) - 6);
				rel(((x & 7) << 1) - 8);
			} else { i++; triple(c - 64, ch[i]); }
		} else if (mode == 1) {
			if (c < 32) {
				const int run = (c >> 2) & 7;
				const uint8_t v = last;
				for (int e = 0; e < run + 2; e++) emit(v);
				const int k = c & 3;
				if (k == 1) rel(2);
				else if (k == 2) rel(-2);
				else if (k == 3) rel(0);
			} else if (c < 64) {
				const int x = c - 32;
				rel(((x >> 3) << 1) - 4);
				rel(((x & 7) << 1) - 8);
			} else { i++; triple(c - 64, ch[i]); }
		} else {
			if (c < 64) {
				const int run = c & 63;
				const uint8_t v = last;
				for (int e = 0; e < run + 2; e++) emit(v);
			} else { i++; triple(c - 64, ch[i]); }
		}
	}
	j = 16384;
	if (i >= ch_len) return NHW_ERR_STREAM_DEV;
	emit(ch[i++]);
	for (; j < 24576; i++) {
		if (i >= ch_len) return NHW_ERR_STREAM_DEV;
		const int c = ch[i];
		if (c >= 192) {
			const int x = c - 192, k = x >> 2;
			const int d0 = k < 2 ? 0 : (k == 2 || k == 4 || k == 5) ? 4 : -4;
			const int d1 = (k == 0 || k == 4 || k == 6) ? 4 : (k == 1 || k == 5 || k == 7) ? -4 : 0;
			rel(d0);
			rel(d1);
			const int m = x & 3;
			rel(m == 0 ? 0 : m == 1 ? 4 : m == 2 ? -4 : 8);
		} else if (c >= 128) emit((c - 128) << 2);
		else if (c >= 64) {
			int run = (c >> 3) & 7;
			const uint8_t v = last;
			if (run == 7) {
				run = (c & 7) + 7;
				for (int e = 0; e < run + 2; e++) emit(v);
			} else {
				for (int e = 0; e < run + 2; e++) emit(v);
				const int k = c & 7;
				if (k == 1) rel(4);
				else if (k == 2) { rel(4); rel(-4); }
				else if (k == 3) { rel(4); rel(-4); rel(0); }
				else if (k == 4) { rel(-4); rel(4); rel(0); }
				else if (k == 5) { rel(-4); rel(4); }
				else if (k == 6) rel(-4);
				else if (k == 7) rel(8);
			}
		} else {
			rel(((c >> 3) << 2) - 16);
			rel(((c & 7) << 2) - 16);
		}
	}
	return 0;
}

// ---- position list expansion (nhw_decoder.c:93-183 and its res5/res3 twins).
// in: packed list (pair-delta bytes >=128, 127 = row step), LSB plane.  out: col + (row<<8).
// Returns the number of entries; cap = 8*bit_len like the reference's calloc.
#define NHW_CAP_HQ_LIST 32768
template <typename T>
NHW_HDN int dec_expand_list(const uint8_t *res, int len, const uint8_t *bits, int bit_len, T *out, int cap_limit = 1 << 30)
{
	const int cap = (bit_len << 3) < cap_limit ? (bit_len << 3) : cap_limit;
	for (int i = 0; i < cap; i++) out[i] = 0;
	int stage = 0, count;
	auto last = [&]() { return stage > 0 ? (int)out[stage - 1] : 0; };   // [-1] reads the zero guard
	int prev = res[0];                       // value of res[i-1] as the reference leaves it behind
	if (res[0] == 127) count = 1;
	else { out[stage++] = (T)(res[0] << 1); count = 0; }
	for (int i = 1; i < len; i++) {
		int cur = res[i];
		if (cur >= 128) {
			const int e = (cur - 128) >> 4, scan = cur & 15;
			if (prev != 127) {
				int j = (last() & 255) + (e << 1);
				if (j >= 254) { count++; cur = 127; }
				else if (stage < cap) out[stage++] = (T)(j + (count << 8));
				j += scan << 1;
				if (j >= 254) { count++; cur = 127; }
				else if (stage < cap) out[stage++] = (T)(j + (count << 8));
			} else { cur = 127; count += 2; }
		} else if (cur == 127) count++;
		else {
			if ((cur << 1) < (last() & 255) && prev != 127) count++;
			if (stage < cap) out[stage++] = (T)((cur << 1) + (count << 8));
		}
		prev = cur;
	}
	for (int i = 0; i < cap; i++) out[i] = (T)(out[i] + ((bits[i >> 3] >> (7 - (i & 7))) & 1));
	return stage;
}
