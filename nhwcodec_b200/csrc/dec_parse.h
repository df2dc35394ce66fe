// dec_parse.h -- host-side header walk of one .nhw stream (decoder/nhw_decoder.c:1494-1661,
// SURVEY.md Appendix A): fills the per-image descriptor the decode kernels work from.
// Host/device: the host API walks headers on the host (api.cu), the device-resident API in a kernel (decode.cu:
// kd_parse_headers); the host test harness compiles it with g++.
//
// The stream is untrusted: every length the device code later uses as a loop bound or an index
// range is validated here against the capacity of the decode workspace (decode.cu: DOFF_*), and no
// byte beyond `len` is read.  Return codes: 0 ok, -7 not an .nhw stream / truncated / inconsistent
// (NHW_ERR_STREAM), -3 quality outside 1..23 (NHW_ERR_QUALITY), -5 a section larger than the
// workspace holds (NHW_ERR_OVERFLOW).
#pragma once
#include <stdint.h>
#include <string.h>

#include "dec_core.cuh"

// capacity limits of the decode workspace, in list entries (decode.cu: 8 lists x 65536 u16)
#define NHW_DEC_LIST_ENTRIES 65536
#define NHW_DEC_BOOK_BYTES 1000          // flat codebook bytes dec_build_book expands (its scratch is 1024)

NHW_HDN int nhw_parse_header(const uint8_t *p, size_t len, DecDesc *d)
{
	*d = DecDesc{};
	if (len < 2) return -7;
	size_t pos = 0;
	bool trunc = false;
	auto u8 = [&]() -> int { if (pos + 1 > len) { trunc = true; return 0; } return p[pos++]; };
	auto u16 = [&]() -> int { if (pos + 2 > len) { trunc = true; return 0; } int v = p[pos] | (p[pos + 1] << 8); pos += 2; return v; };
	auto u32 = [&]() -> int {
		if (pos + 4 > len) { trunc = true; return 0; }
		uint32_t v = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16) | ((uint32_t)p[pos + 3] << 24);
		pos += 4;
		return (int)v;
	};
	d->byte0 = u8();
	const int q = d->quality = u8();
	if (d->byte0 > 6) return -7;                       // "Not an .nhw file"
	if (q < 1 || q > 23) return -3;
	d->size_tree1 = u16(); d->size_tree2 = u16();
	d->size_data1 = u32(); d->size_data2 = u32();
	d->tree_end = u16(); d->exw_Y_end = u16();
	if (q > 12) d->res1_len = u16();
	if (q >= 19) { d->res3_len = u16(); d->res3_bit_len = u16(); }
	if (q > 17) d->res4_len = u16();
	if (q > 12) d->res1_bit_len = u16();
	if (q >= 21) { d->res5_len = u16(); d->res5_bit_len = u16(); }
	if (q > 21) { d->res6_len = u32(); d->res6_bit_len = u16(); d->char_res1_len = u16(); }
	if (q > 22) d->qsetting3_len = u16();
	d->select1 = u16(); d->select2 = u16();
	if (q > 15) d->highres_comp_len = u16();
	d->end_ch_res = u16();
	if (trunc) return -7;
	if (d->size_data1 <= 0 || d->size_data2 < d->size_data1 || d->res6_len < 0) return -7;
	auto take = [&](uint32_t &off, size_t n) { off = (uint32_t)pos; pos += n; };
	take(d->off_tree1, d->size_tree1);
	take(d->off_tree2, d->size_tree2);
	take(d->off_exw, d->exw_Y_end);
	if (q > 12) { take(d->off_res1, d->res1_len); take(d->off_res1_bit, d->res1_bit_len); take(d->off_res1_word, d->res1_bit_len); }
	if (q > 17) take(d->off_res4, d->res4_len);
	if (q >= 19) { take(d->off_res3, d->res3_len); take(d->off_res3_bit, d->res3_bit_len); take(d->off_res3_word, 2 * (size_t)d->res3_bit_len); }
	if (q >= 21) { take(d->off_res5, d->res5_len); take(d->off_res5_bit, d->res5_bit_len); take(d->off_res5_word, d->res5_bit_len); }
	if (q > 21) {
		take(d->off_res6, (size_t)(uint32_t)d->res6_len); take(d->off_res6_bit, d->res6_bit_len); take(d->off_res6_word, d->res6_bit_len);
		take(d->off_char_res1, 2 * (size_t)d->char_res1_len);
	}
	if (q > 22) take(d->off_qsetting3, 4 * (size_t)d->qsetting3_len);
	take(d->off_sel1, d->select1);
	take(d->off_sel2, d->select2);
	if (q > 15) { take(d->off_u64, 512); take(d->off_v64, 512); take(d->off_highres, d->highres_comp_len); }
	take(d->off_ch_res, d->end_ch_res);
	take(d->off_words, 4 * (size_t)d->size_data2);
	d->blob_len = (uint32_t)len;
	if (pos > len) return -7;
	// ---- capacities of the decode workspace (the reference callocs 8*bit_len entries per list; ours hold 65536)
	if ((d->res1_bit_len << 3) > NHW_DEC_LIST_ENTRIES || (d->res5_bit_len << 3) > NHW_DEC_LIST_ENTRIES ||
	    (d->res3_bit_len << 3) > NHW_DEC_LIST_ENTRIES)
		return -5;
	if (q > 21 && (d->res6_bit_len << 3) > NHW_CAP_HQ_LIST) return -5;   // more res6 entries than the decode workspace holds
	// the chroma codebook takes its length from the header; the flat form of either book is at most 1000 bytes
	if (d->tree_end > NHW_DEC_BOOK_BYTES) return -7;
	// a position list starts with one entry and needs one LSB byte per 8 entries: empty sections cannot be walked
	if (q > 12 && (d->res1_len < 1 || d->res1_bit_len < 1)) return -7;
	if (q >= 19 && (d->res3_len < 1 || d->res3_bit_len < 1)) return -7;
	if (q >= 21 && (d->res5_len < 1 || d->res5_bit_len < 1)) return -7;
	if (q > 21 && (d->res6_len < 1 || d->res6_bit_len < 1)) return -7;
	if (d->end_ch_res < 2) return -7;                  // one seed byte for luma, one for chroma at least
	return 0;
}
