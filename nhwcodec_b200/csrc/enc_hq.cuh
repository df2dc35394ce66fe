// enc_hq.cuh -- the q22/q23 luma side channel (`res6`, `char_res1`, `high_qsetting3`).
//
// Reference behaviour:
//   snapshot of the first horizontal pass      encoder/wavelet_filterbank.c:107-112 (im_quality_setting)
//   copy of the reconstructed LL1              encoder/nhw_encoder.c:766-777        (im_wavelet_first_order)
//   residual codes folded into that copy       encoder/nhw_encoder.c:1426-1496
//   LH1 rebuilt from its quantised bytes       encoder/image_processing.c:523-556   (im_recons_wavelet_band)
//   half synthesis, comparison, lists          encoder/wavelet_filterbank.c:498-707
//
// What it is: the decoder's first inverse pass of level 1 produces, for the low band, exactly
// upfilter53I(LL1') + upfilter53III(LH1'); the encoder runs that half-synthesis on what the decoder
// will have, compares it with the true first-pass low band it kept from the analysis, and sends
// the positions where they differ by more than 34 (30 at q23) as +-32 corrections, plus a second
// list of +-56 corrections at q23.
#pragma once
#include "enc_pack.cuh"
#include "nhw_tables.cuh"

#define NHW_CAP_CHAR_RES1 256   // the reference mallocs 256 entries (encoder/wavelet_filterbank.c:563)
#define NHW_CAP_QSETTING3 16384
NHW_HD int nhw_extra_value_enc(int a) { return a >= 0 && a < 110 ? (int)nhw_extra_table[a] : 0; }

// ---- residual codes -> corrections of the LL1 copy (E17).  ll1 row r, column j < 254 -> fo[j*256 + r + k]
NHW_HD int hq_e17_delta(int code, int (&d)[3])
{
	d[0] = d[1] = d[2] = 0;
	if (code == 141) d[0] = -5;
	else if (code == 140) d[0] = 5;
	else if (code == 144) d[0] = -3;
	else if (code == 145) d[0] = 3;
	else if (code == 121) { d[0] = -4; d[1] = -3; }
	else if (code == 122) { d[0] = 4; d[1] = 3; }
	else if (code == 123) { d[0] = 2; d[1] = 2; d[2] = 2; }
	else if (code == 124) { d[0] = -2; d[1] = -2; d[2] = -2; }
	else if (code == 126) { d[0] = 9; d[1] = 3; }
	else if (code == 125) { d[0] = -9; d[1] = -3; }
	else if (code == 148) d[0] = -8;
	else if (code == 149) d[0] = 8;
	else return 0;
	return 1;
}

// ---- LH1 from its bytes.  byte_at(c) = quantised byte of LH1 cell c = row*256 + col (the band is
// rows 0..255, columns 256..511 of the coefficient plane).  The reference walks the cells in order,
// a 127/129 byte ("three in a row") writes the cells left and right of it as well and skips the next
// cell of its row; the last write to a cell wins.
NHW_HD bool hq_is_pattern(int a) { return a == 127 || a == 129; }
template <typename ByteAt>
NHW_HD bool hq_visited_pattern(ByteAt byte_at, int c)
{
	if (!hq_is_pattern(byte_at(c))) return false;
	// skipped iff the cell on its left, in the same row, is a visited pattern: runs of pattern bytes alternate
	int run = 0;
	const int col = c & 255;
	while (run < col && hq_is_pattern(byte_at(c - 1 - run))) run++;
	return (run & 1) == 0;
}
template <typename ByteAt>
NHW_HD int hq_band_cell(ByteAt byte_at, int c)
{
	if (c + 1 < 65536 && hq_visited_pattern(byte_at, c + 1)) return byte_at(c + 1) == 127 ? 5 : -5;
	const bool left_vp = c > 0 && hq_visited_pattern(byte_at, c - 1);
	const bool skipped = left_vp && (c & 255) != 0;
	if (!skipped) {
		const int a = byte_at(c);
		if (a == 127) return 6;
		if (a == 129) return -7;
		if (a != 128) {
			if (a & 7) {
				const int e = nhw_extra_value_enc(a);
				return e > 0 ? 123 + (e << 3) : (e << 3) - 123;
			}
			return a > 128 ? a - 125 : a - 131;
		}
	}
	if (left_vp) return byte_at(c - 1) == 127 ? 5 : -5;
	return 0;
}

// ---- half synthesis of row r and comparison with the kept first pass.  tag: 0 none, 1 = +32 (30000),
// 2 = -32 (31000), 3 = +56 (32000), 4 = -56 (32500)
NHW_HD void hq_tag_pair(const EncImg &im, int q, int r, int t)   // outputs 2t, 2t+1 of row r
{
	const int16_t *fo = im.hq_fo + r * 256, *band = im.hq_band + r * 256, *qs = im.hq_qs + r * 512;
	uint8_t *tag = im.hq_tag + r * 512;
	const int thr = q > 22 ? 30 : 34;
	auto l = [&](int k) { return (int)fo[k]; };
	auto h = [&](int k) { return (int)band[k]; };
	int v[2];
	inverse_pair(l, h, t, 256, false, v[0], v[1]);
	for (int k = 0; k < 2; k++) {
		const int d = qs[2 * t + k] - v[k];
		int g = 0;
		if (nhw_iabs(d) > thr) {
			if (q > 22 && nhw_iabs(d) > 56) g = d > 0 ? 3 : 4;
			else g = d > 0 ? 1 : 2;
		}
		tag[2 * t + k] = (uint8_t)g;
	}
}

// positions + sign words of one row (two halves, each closed by the marker 254); returns entries written
// to pos (markers included); nw = sign words written.  pos/wrd may be NULL to count only.
NHW_HD int hq_collect_row(const EncImg &im, int r, uint8_t *pos, uint8_t *wrd, int &nw)
{
	const uint8_t *tag = im.hq_tag + r * 512;
	int n = 0;
	nw = 0;
	for (int j = 0; j < 512; j++) {
		if (j == 254 || j == 510) {
			if (pos) pos[n] = 254;
			n++;
			j++;
		} else if (tag[j] == 1 || tag[j] == 2) {
			if (pos) { pos[n] = (uint8_t)(j & 255); wrd[nw] = (uint8_t)(tag[j] - 1); }
			n++;
			nw++;
		}
	}
	return n;
}
