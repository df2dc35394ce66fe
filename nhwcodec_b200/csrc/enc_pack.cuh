// enc_pack.cuh -- entropy stage and container writer.
//
//   packet_stream_image : wavlts2packet, encoder/compress_pixel.c:53-469 (one call per stream:
//                         part 0 = 262144 luma bytes, part 1 = 131072 interleaved chroma bytes)
//   write_stream_image  : write_compressed_file, encoder/nhw_encoder.c:3100-3220
//
// Serial per image in this first version: histogram -> alphabet -> rank sort -> rank code
// emission, MSB-first into 32-bit words.
#pragma once
#include "enc_c.cuh"
#include "nhw_tables.cuh"

#define NHW_ERR_CODEBOOK_DEV (-4)
#define NHW_ERR_OVERFLOW_DEV (-5)
#define NHW_ERR_STREAM_DEV (-7)
#define NHW_WORDS_LIMIT (131072 - 4)   // ENC_WORDS_CAP minus slack, enc_batch.cuh

struct PackState {
	int rle_buf[256];     // symbol histogram, later symbol -> rank
	int rle_128[256];     // zero-run-length histogram, later run length -> rank
	uint32_t weight[354];
	uint16_t sym[354];    // (run length << 8) | 128  or  (1 << 8) | byte
};

// candidate alphabet in the reference's enumeration order (compress_pixel.c:136-229)
template <typename F>
NHW_HD void for_each_symbol(F f)
{
	for (int i = 0; i < 109; i += 2) f(i);
	f(112);
	for (int i = 120; i < 141; i++) f(i);
	for (int i = 144; i < 256; i += 4) f(i);
}

NHW_HD void put_bits(uint32_t *w, int &a, int &pack, uint32_t code, int len)
{
	pack += len;
	if (pack <= 32) w[a] |= code << (32 - pack);
	else {
		int match = pack - 32;
		w[a] |= code >> match;
		a++;
		w[a] |= (code & ((1u << match) - 1u)) << (32 - match);
		pack = match;
	}
}

// Alphabet, rank order and rank tables from the two histograms (compress_pixel.c:129-273),
// in three steps so that the CUDA path can run the sort across a CTA:
//   pack_enumerate : candidate alphabet in the reference's enumeration order; raises `select`
//                    (minimum coded zero-run length) until at most 354 entries remain
//   (sort)         : stable, by decreasing weight (the reference's bubble sort keeps ties in order)
//   pack_finish    : symbol -> rank / run length -> rank tables, overflow checks
NHW_HDN int pack_enumerate(PackState &st, int &select, int &k)
{
	for (;;) {
		uint32_t w128 = st.rle_buf[128] > 0 ? (uint32_t)st.rle_buf[128] : 0u;
		for (int j = 2; j < 256; j++) if (st.rle_128[j] > 0) w128 += (uint32_t)(j * st.rle_128[j]);
		for (int j = 2; j < select; j++) st.rle_128[j] = 0;
		for (int j = select; j < 256; j++) if (st.rle_128[j] > 0) w128 -= (uint32_t)(j * st.rle_128[j]);
		st.rle_buf[128] = (int)w128;
		k = 0;
		for (int j = select; j < 256; j++)
			if (st.rle_128[j] > 0) { st.sym[k] = (uint16_t)((j << 8) | 128); st.weight[k] = (uint32_t)st.rle_128[j]; k++; }
		for_each_symbol([&](int i) {
			if (st.rle_buf[i] > 0) { st.sym[k] = (uint16_t)((1 << 8) | i); st.weight[k] = (uint32_t)st.rle_buf[i]; k++; }
		});
		if (k <= 354) return 0;
		select++;
		if (select >= 100) return NHW_ERR_CODEBOOK_DEV;
	}
}

NHW_HDN int pack_finish(PackState &st, int part, int select, int k, int &b)
{
	for (int i = 0; i < k; i++) {
		if ((st.sym[i] >> 8) == 1) st.rle_buf[st.sym[i] & 0xff] = i;
		else st.rle_128[st.sym[i] >> 8] = i;
	}
	b = st.sym[0] == ((1 << 8) | 128) ? 1 : 0;
	if (part == 0 && b == 0 && k > 290) return NHW_ERR_CODEBOOK_DEV;
	if (part == 1 && select != 4 && k > 290) return NHW_ERR_CODEBOOK_DEV;
	const bool zone = (part == 0 && select == 4 && b == 1);
	if (!zone && k > 290) return NHW_ERR_CODEBOOK_DEV;   // the reference would index past its code table
	return 0;
}

NHW_HDN int pack_alphabet(PackState &st, int part, int &select, int &k, int &b)
{
	int rc = pack_enumerate(st, select, k);
	if (rc) return rc;
	for (int i = 1; i < k; i++) {   // stable insertion sort, decreasing weight
		uint32_t wv = st.weight[i];
		uint16_t sv = st.sym[i];
		int j = i - 1;
		while (j >= 0 && st.weight[j] < wv) { st.weight[j + 1] = st.weight[j]; st.sym[j + 1] = st.sym[j]; j--; }
		st.weight[j + 1] = wv;
		st.sym[j + 1] = sv;
	}
	return pack_finish(st, part, select, k, b);
}

// Codebook section of the container for one stream (compress_pixel.c:400-461)
// scratch: 2 x 1024 bytes (the raw list, then its de-interleaved copy); NULL = use the image's list scratch
// The reference de-interleaves into ONE 580-byte array (compress_pixel.c:58) that it fills for the luma book and then
// again for the chroma book, and its run-length loop does not stop at the end of the list while it keeps seeing the
// marker byte (:412-414, :446-448).  So when the chroma list ends on its marker (128), the run goes on into whatever
// the LUMA list left behind at those positions -- a luma symbol byte 128 there makes the last run one longer (seen on
// about 1 image in 4000).  `de` therefore persists from part 0 to part 1 (the caller passes the same scratch), is
// cleared once per image, and the run loop reads on past n.  Beyond the luma list the array is never-written stack
// memory in the reference; here it reads as 0, like every other such read.
#define NHW_BOOK_ARRAY 580
NHW_HDN void pack_codebook(const EncImg &im, const PackState &st, int part, int k, uint8_t *scratch = nullptr)
{
	EncHdr *h = im.hdr;
	// ---- codebook: ranked symbol list, de-interleaved (even then odd positions), runs of the
	// marker byte compressed (compress_pixel.c:400-461)
	const int marker = part ? 128 : 3;
	uint8_t *raw = scratch ? scratch : im.tmp3 + 40000, *de = raw + 1000;
	int n = 0;
	for (int i = 0; i < k; i++) {
		if ((st.sym[i] >> 8) == 1) raw[n++] = (uint8_t)(part ? ((st.sym[i] & 0xff) | 1) : (st.sym[i] & 0xff));
		else { raw[n++] = (uint8_t)marker; raw[n++] = (uint8_t)(st.sym[i] >> 8); }
	}
	if (part) h->tree_end = n;
	if (!part) for (int i = 0; i < NHW_BOOK_ARRAY + 4; i++) de[i] = 0;
	int m = 0;
	for (int i = 0; i < n; i += 2) de[m++] = raw[i];
	for (int i = 1; i < n; i += 2) de[m++] = raw[i];
	uint8_t *outb = part ? im.codebook2 : im.codebook1;
	int o = 0, run = 0;
	for (int i = 0; i < n; i++) {
		while (i < NHW_BOOK_ARRAY && de[i] == marker) { run++; i++; }
		if (run > 0) { outb[o++] = (uint8_t)marker; outb[o++] = (uint8_t)run; run = 0; i--; }
		else outb[o++] = de[i];
	}
	if (part) h->size_tree2 = o; else h->size_tree1 = o;
}

// Returns 0 or NHW_ERR_CODEBOOK_DEV.  `a` is the running word index across both parts.
NHW_HDN int packet_stream_image(const EncImg &im, int part, int &a)
{
	PackState &st = *static_cast<PackState *>(im.pack_scratch);
	uint8_t *s = im.scan;
	EncHdr *h = im.hdr;
	uint32_t *words = im.words;
	const int p1 = part ? 262144 : 0, p2 = part ? 393216 : 262144;
	int select = part ? 3 : 4;
	uint8_t saved = 0;
	if (!part) { saved = s[262144]; s[262144] = 3; }
	else s[393215] = s[393214];
	for (int i = 0; i < 256; i++) { st.rle_buf[i] = 0; st.rle_128[i] = 0; }
	// ---- statistics (compress_pixel.c:81-107)
	{
		int e = 1, c = 0;
		for (int i = p1; i < p2 - 1; i++) {
			if (s[i] == 128) {
				while (i < p2 - 1 && s[i + 1] == 128) {
					e++; c = 1;
					if (e > 255) { st.rle_128[254]++; e = 1; c = 0; continue; }
					i++;
				}
			}
			if (c) st.rle_128[e]++;
			else st.rle_buf[s[i]]++;
			e = 1; c = 0;
		}
	}
	int k = 0, b = 0;
	{
		const int rc = pack_alphabet(st, part, select, k, b);
		if (rc) return rc;
	}
	const bool zone = (part == 0 && select == 4 && b == 1);
	// ---- emission (compress_pixel.c:279-361)
	uint8_t *s1 = im.tmp1, *s2 = im.tmp2;
	int c = 0, j = 0, e = 1, pack = 0, tag = 0;
	for (int i = p1; i < p2 - 1; i++) {
		const int pixel = s[i];
		int pos;
		bool direct = false;
		if (pixel == 153) { s1[c++] = 0; continue; }
		if (pixel == 155) { s1[c++] = 1; continue; }
		if (pixel == 157) { s2[j++] = 0; continue; }
		if (pixel == 159) { s2[j++] = 1; continue; }
		if (pixel != 128 && pixel < 136 && pixel > 120) {
			pos = st.rle_buf[pixel] & 0xffff;
			if (pixel > 131) i += 4;
			direct = true;
		} else if (pixel == 128) {
			bool overflow = false;
			while (i < p2 - 1 && s[i + 1] == 128) {
				e++;
				if (e > 255) { e = 254; i--; overflow = true; break; }
				i++;
			}
			if (!overflow && e > 1 && e < select) { i -= e - 1; tag = e; e = 1; }
		}
		for (;;) {
			if (!direct) pos = (e == 1 ? st.rle_buf[pixel] : st.rle_128[e]) & 0xffff;
			direct = false;
			if (pos >= 110 && pos < 174 && zone) put_bits(words, a, pack, (1u << 6) | (uint32_t)(pos - 110), 15);
			else {
				if (pos >= 174 && zone) pos -= 64;
				if (pos >= NHW_CODE_DEPTH) return NHW_ERR_CODEBOOK_DEV;   // byte outside the alphabet
				put_bits(words, a, pack, nhw_code_bits[pos], nhw_code_len[pos]);
			}
			if (a >= NHW_WORDS_LIMIT) return NHW_ERR_OVERFLOW_DEV;
			e = 1;
			if (tag > 0) { tag--; if (tag > 0) { i++; continue; } }
			break;
		}
	}
	// ---- side outputs
	if (part == 0) {
		h->size_data1 = a + 1;
		h->wavelet_type = (select > 4 || b == 0) ? 4 : 0;
		for (int t = 0; t < 8; t++) { s1[c + t] = 0; s2[j + t] = 0; }
		int n1 = 0, n2 = 0;
		for (int i = 0; i < ((c >> 3) + 1) * 8; i += 8) {
			int v = 0;
			for (int t = 0; t < 8; t++) v |= (s1[i + t] & 1) << (7 - t);
			im.sel1[n1++] = (uint8_t)v;
		}
		for (int i = 0; i < ((j >> 3) + 1) * 8; i += 8) {
			int v = 0;
			for (int t = 0; t < 8; t++) v |= (s2[i + t] & 1) << (7 - t);
			im.sel2[n2++] = (uint8_t)v;
		}
		h->select1 = n1;
		h->select2 = n2;
	} else h->size_data2 = a + 1;
	pack_codebook(im, st, part, k);
	if (!part) s[262144] = saved;
	return 0;
}

// ---- container.  Returns the stream length.
NHW_HD void put16(uint8_t *&p, int v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p += 2; }
NHW_HD void put32(uint8_t *&p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); p += 4; }

// The container as an ordered list of byte ranges: `hdr` receives the header fields (<= 64 bytes, returned
// length), section(src, n) is called for every section in file order.  Multi-byte section elements
// (char_res1, high_qsetting3, the code words) are little-endian in memory already.
template <typename Section>
NHW_HD int stream_layout(const EncImg &im, uint8_t *hdr, Section section)
{
	const EncHdr *h = im.hdr;
	const int q = h->quality;
	uint8_t *p = hdr;
	const int exw_end = h->exw_y_len + 2 + h->exw_u_len + 2 + h->exw_v_len;
	*p++ = (uint8_t)(h->res_low + h->wavelet_type);
	*p++ = (uint8_t)q;
	put16(p, h->size_tree1);
	put16(p, h->size_tree2);
	put32(p, (uint32_t)h->size_data1);
	put32(p, (uint32_t)h->size_data2);
	put16(p, h->tree_end);
	put16(p, exw_end);
	if (q > 12) put16(p, h->res1_len);
	if (q >= 19) { put16(p, h->res3_len); put16(p, h->res3_bit_len); }
	if (q > 17) put16(p, h->res4_len);
	if (q > 12) put16(p, h->res1_bit_len);
	if (q >= 21) { put16(p, h->res5_len); put16(p, h->res5_bit_len); }
	if (q > 21) { put32(p, (uint32_t)h->res6_len); put16(p, h->res6_bit_len); put16(p, h->char_res1_len); }
	if (q > 22) put16(p, h->qsetting3_len);
	put16(p, h->select1);
	put16(p, h->select2);
	if (q > 15) put16(p, h->highres_comp_len);
	put16(p, h->end_ch_res);
	const int hdr_len = (int)(p - hdr);
	section(hdr, hdr_len);
	section(im.codebook1, h->size_tree1);
	section(im.codebook2, h->size_tree2);
	section(im.exw, h->exw_y_len);
	section(nullptr, 2);                       // 0,0 separator
	section(im.exw_uv, h->exw_u_len);
	section(nullptr, 2);
	section(im.exw_uv + 16384, h->exw_v_len);
	if (q > 12) { section(im.res1, h->res1_len); section(im.res1_bit, h->res1_bit_len); section(im.res1_word, h->res1_word_len); }
	if (q > 17) section(im.res4, h->res4_len);
	if (q >= 19) { section(im.res3, h->res3_len); section(im.res3_bit, h->res3_bit_len); section(im.res3_word, h->res3_word_len); }
	if (q >= 21) { section(im.res5, h->res5_len); section(im.res5_bit, h->res5_bit_len); section(im.res5_word, h->res5_word_len); }
	if (q > 21) {
		section(im.res6, h->res6_len); section(im.res6_bit, h->res6_bit_len); section(im.res6_word, h->res6_word_len);
		section(reinterpret_cast<const uint8_t *>(im.char_res1), 2 * h->char_res1_len);
	}
	if (q > 22) section(reinterpret_cast<const uint8_t *>(im.qsetting3), 4 * h->qsetting3_len);
	section(im.sel1, h->select1);
	section(im.sel2, h->select2);
	if (q > 15) {
		section(im.res_uv64, 1024);
		section(im.highres_word, h->highres_comp_len);
	}
	section(im.llcode, h->end_ch_res);
	section(reinterpret_cast<const uint8_t *>(im.words), 4 * h->size_data2);
	return hdr_len;
}

NHW_HDN int write_stream_image(const EncImg &im, uint8_t *out)
{
	uint8_t hdr[64];
	uint8_t *p = out;
	stream_layout(im, hdr, [&](const uint8_t *src, int n) {
		for (int i = 0; i < n; i++) p[i] = src ? src[i] : (uint8_t)0;
		p += n;
	});
	return (int)(p - out);
}
