// dec_stages.cuh -- inline stages of decode_image (decoder/nhw_decoder.c:71-1474) and the colour
// conversion of write_image_bmp (decoder/nhw_decoder_cli.c:108-291), q17..q23.
// *_row / *_strip functions are independent per row/strip; *_image functions are raster-ordered.
#pragma once
#include "dec_core.cuh"

// ---- D3: expand + split the side-channel lists (nhw_decoder.c:93-491).  list_len[8] receives
// the value the reference's `count` variable is left with (it is read, stale, by D4).
#define DEC_LIST_CAP 65536   // entries per expanded list (and of `tmp`); nhw_parse_header rejects streams that need more
NHW_HDN void dec_lists_image(const DecImg &im, uint16_t *tmp /* DEC_LIST_CAP */)
{
	const DecDesc *d = im.d;
	const int q = d->quality;
	int stale = 262144;   // q <= 12 expands no list: `count` still holds the de-scan loop's final value (nhw_decoder.c:71-91)
	for (int k = 0; k < 8; k++) im.list_len[k] = 0;
	for (int pass = 0; pass < 2; pass++) {   // res1 (q>12), then res5 (q>=21): one word bit per entry
		if (pass == 0 ? !(q > 12) : !(q >= 21)) continue;
		const uint8_t *res = im.blob + (pass ? d->off_res5 : d->off_res1);
		const uint8_t *bits = im.blob + (pass ? d->off_res5_bit : d->off_res1_bit);
		const uint8_t *word = im.blob + (pass ? d->off_res5_word : d->off_res1_word);
		const int len = pass ? d->res5_len : d->res1_len, bit_len = pass ? d->res5_bit_len : d->res1_bit_len;
		dec_expand_list(res, len, bits, bit_len, tmp, DEC_LIST_CAP);
		uint16_t *minus = im.list[2 * pass], *plus = im.list[2 * pass + 1];
		int nm = 0, np = 0, c = 0;
		for (int i = 0; i < bit_len - 1 && c + 8 <= DEC_LIST_CAP; i++)   // the header check keeps 8*bit_len within the lists
			for (int b = 7; b >= 0; b--) {
				if ((word[i] >> b) & 1) minus[nm++] = tmp[c++];
				else plus[np++] = tmp[c++];
			}
		im.list_len[2 * pass] = nm;
		im.list_len[2 * pass + 1] = np;
		stale = c;
	}
	if (q >= 19) {   // res3: two word bits per entry, four classes
		dec_expand_list(im.blob + d->off_res3, d->res3_len, im.blob + d->off_res3_bit, d->res3_bit_len, tmp, DEC_LIST_CAP);
		const uint8_t *word = im.blob + d->off_res3_word;
		int n[4] = {0, 0, 0, 0}, c = 0;
		for (int i = 0; i < (d->res3_bit_len << 1) - 2 && c + 4 <= DEC_LIST_CAP; i++)
			for (int b = 6; b >= 0; b -= 2) {
				const int sel = (word[i] >> b) & 3;   // 0 -> nhwres4 (+4,+3), 1 -> nhwres3 (-4,-3), 2 -> +2 x3, 3 -> -2 x3
				im.list[4 + sel][n[sel]++] = tmp[c++];
			}
		for (int k = 0; k < 4; k++) im.list_len[4 + k] = n[k];
		stale = c;
	}
	im.list_len[8] = stale;
}

// ---- q22/q23: expand + split the res6 list (nhw_decoder.c:282-388); positions are flat indices
// (half-row * 256 + column) into the plane after the first half of the level-1 synthesis
NHW_HDN void dec_hq_lists_image(const DecImg &im, uint32_t *tmp /* NHW_CAP_HQ_LIST */)
{
	const DecDesc *d = im.d;
	im.list_len[11] = im.list_len[12] = 0;
	if (d->quality <= 21) return;
	dec_expand_list(im.blob + d->off_res6, d->res6_len, im.blob + d->off_res6_bit, d->res6_bit_len, tmp, NHW_CAP_HQ_LIST);
	const uint8_t *word = im.blob + d->off_res6_word;
	int nm = 0, np = 0, c = 0;
	for (int i = 0; i < d->res6_bit_len - 1 && c + 8 <= NHW_CAP_HQ_LIST; i++)
		for (int b = 7; b >= 0; b--) {
			if ((word[i] >> b) & 1) im.hq_list[0][nm++] = tmp[c++];
			else im.hq_list[1][np++] = tmp[c++];
		}
	im.list_len[11] = nm;
	im.list_len[12] = np;
}

// q22/q23 add-backs inside wavelet_synthesis2 (decoder/wavelet_filterbank.c:296-348), entry k of the
// concatenation [-32 list | +32 list | char_res1 | high_qsetting3]: target cell and amount
NHW_HD bool dec_hq_addback(const DecImg &im, int k, int &pos, int &amount)
{
	const DecDesc *d = im.d;
	const int n0 = im.list_len[11], n1 = im.list_len[12];
	if (k < n0) { pos = (int)im.hq_list[0][k]; amount = -32; return true; }
	k -= n0;
	if (k < n1) { pos = (int)im.hq_list[1][k]; amount = 32; return true; }
	k -= n1;
	if (k < d->char_res1_len) {
		const uint8_t *p = im.blob + d->off_char_res1 + 2 * k;
		const int v = p[0] | (p[1] << 8), m = v & 3;
		pos = ((v - m) << 1) + (m < 2 ? 254 : 255);
		amount = (m & 1) ? -32 : 32;
		return true;
	}
	k -= d->char_res1_len;
	if (d->quality > 22 && k < d->qsetting3_len) {
		const uint8_t *p = im.blob + d->off_qsetting3 + 4 * k;
		const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
		pos = (int)(v >> 1);
		amount = (v & 1) ? -56 : 56;
		return true;
	}
	return false;
}
NHW_HD int dec_hq_addback_count(const DecImg &im)
{
	return im.d->quality > 21 ? im.list_len[11] + im.list_len[12] + im.d->char_res1_len + (im.d->quality > 22 ? im.d->qsetting3_len : 0) : 0;
}

// ---- D4: marker expansion, in place and in raster order (nhw_decoder.c:493-607)
NHW_HDN void dec_y_markers_image(const DecImg &im)
{
	int16_t *J = im.jpeg;
	const int q = im.d->quality;
	for (int r = 0; r < 256; r++)
		for (int j = 0; j < 512; j++) {
			const int s = r * YW + j, v = J[s];
			if (v <= 1000) continue;
			if (v == 1008) { J[s - 1] = 5; J[s + 1] = 5; J[s] = (int16_t)(j < 256 ? 5 : 6); }
			else if (v == 1009) { J[s - 1] = -5; J[s + 1] = -5; J[s] = (int16_t)(j < 256 ? -6 : -7); }
			else if (v == 1010) { J[s] = 5; J[s + 1] = 5; J[s + YW] = 5; J[s + YW + 1] = 5; }
			else if (v == 1011) { J[s] = -5; J[s + 1] = -5; J[s + YW] = -5; J[s + YW + 1] = -5; }
			else if (v == 1006) { J[s] = -6; J[s + 1] = -6; }
			else if (v == 1007) { J[s] = 6; J[s + 1] = 6; }
		}
	int count = im.list_len[8];
	for (int half = 0; half < 2; half++)
		for (int r = 256; r < 512; r++)
			for (int j = half ? 256 : 0; j < (half ? 512 : 256); j++) {
				const int s = r * YW + j, v = J[s];
				if (v > 1000) {
					if (v == 1008) { J[s - 1] = 5; J[s] = 6; J[s + 1] = 5; }
					else if (v == 1009) { J[s - 1] = -5; J[s] = -7; J[s + 1] = -5; }
					else if (v == 1006 || v == 1007) {
						const int16_t w = (int16_t)(v == 1006 ? -7 : 7);
						if ((s & 511) < 256) { J[s] = w; J[s + 1] = w; }
						else { J[s - 256] = w; J[s - 768] = w; J[s] = 0; }
					}
				} else if (half && nhw_iabs(v) > 8 && nhw_iabs(v) < 16 && q < 23) {
					if (j > 256 && j < 511) {
						if (nhw_iabs(J[s - 1]) < 8) count++;
						if (nhw_iabs(J[s + 1]) < 8) count++;
						if (nhw_iabs(J[s - YW]) < 8) count++;
						if (nhw_iabs(J[s + YW]) < 8) count++;
						if (count >= 2) J[s] += v > 0 ? 1 : -1;
						count = 0;
					}
				}
			}
}

// ---- D5-D7: LL2 fill, res4 parity restore, exw overrides (nhw_decoder.c:609-658)
// res4 parity restore + exw overrides; returns where the chroma exw entries start
// (res4_done: the parity restore has been applied already -- the CUDA kernel does it with a warp, kd_y_ll)
NHW_HDN int dec_y_ll_overrides(const DecImg &im, bool res4_done = false)
{
	int16_t *J = im.jpeg;
	const DecDesc *d = im.d;
	if (d->quality > 17 && !res4_done) {
		const uint8_t *r4 = im.blob + d->off_res4;
		int count = 0;
		for (int i = 0; i < d->res4_len && count < 128; i++) {   // (a well-formed list names 128 rows; more is garbage)
			if (r4[i] == 128) { count++; continue; }
			const int col = r4[i] > 128 ? r4[i] - 129 : r4[i] - 1;
			if (col >= 0 && col <= 124) {
				const int e = (count << 9) + col;
				for (int k = 0; k < 4; k++)
					if (!(J[e + k] & 1)) J[e + k]++;
			}
			if (r4[i] > 128) count++;
		}
	}
	const uint8_t *x = im.blob + d->off_exw;
	int i = 0;
	for (; i < d->exw_Y_end; i += 3) {
		if (!x[i] && !x[i + 1]) break;
		int col = x[i + 1], val;
		if (col >= 128) { val = x[i + 2] + 255; col -= 128; }
		else val = -x[i + 2];
		J[(x[i] << 9) + col] = (int16_t)val;
	}
	return i;   // exw1: where the chroma entries start (after the 0,0 separator)
}

NHW_HD int dec_lap8(const int16_t *P, int s, int stride)
{
	return (P[s] << 3) - P[s - 1] - P[s + 1] - P[s - stride] - P[s + stride] - P[s - stride - 1] - P[s + stride - 1] -
	       P[s - stride + 1] - P[s + stride + 1];
}

// ---- D14: conditional 5-tap smoothing at the flagged positions, list order (nhw_decoder.c:848-867)
NHW_HDN void dec_y_smooth_flags_plane(const DecImg &im, int16_t *J /* the half-synthesised plane, transposed */)
{
	for (int i = 0; i < im.list_len[9]; i++) {
		const int s = ((im.flags[i] >> 8) << 10) + (im.flags[i] & 255);
		const int res = dec_lap8(J, s, YW);
		if (nhw_iabs(res) < 116) J[s] = (int16_t)(((J[s] << 2) + J[s - 1] + J[s + 1] + J[s - YW] + J[s + YW] + 4) >> 3);
	}
}

NHW_HD uint8_t dec_clip8(int v) { return (uint8_t)((v >> 8) != 0 ? (v < 0 ? 0 : 255) : v); }

// the exw escape entries of one chroma component; exw_pos = index into the exw list (returned advanced past them)
NHW_HDN int dec_c_ll_overrides(const DecImg &im, int exw_pos)
{
	int16_t *J = im.cjpeg;
	const DecDesc *d = im.d;
	const uint8_t *x = im.blob + d->off_exw;
	int i = exw_pos + 2;
	for (; i < d->exw_Y_end; i += 3) {
		if (!x[i] && !x[i + 1]) break;
		int col = x[i + 1], val;
		if (col >= 128) { val = x[i + 2] + 255; col -= 128; }
		else val = -x[i + 2];
		J[(x[i] << 8) + col] = (int16_t)val;
	}
	return i;
}

// LL fill + exw overrides; exw_pos = index into the exw list (advanced past this component)
NHW_HDN int dec_c_ll_image(const DecImg &im, int is_v, int exw_pos)
{
	int16_t *J = im.cjpeg;
	const DecDesc *d = im.d;
	const uint8_t *src = im.res_comp + (is_v ? 20480 : 16384);
	const int bias = d->quality > 15 ? 0 : 1;
	for (int r = 0; r < 64; r++)
		for (int j = 0; j < 64; j++) J[r * CW + j] = (int16_t)(src[r * 64 + j] + bias);
	return dec_c_ll_overrides(im, exw_pos);
}

// markers 5003..5006 in the level-1 bands push +-6 / +-4,+-4 into the reconstructed LL
// (nhw_decoder.c:991-1069).  P = reconstruction (128x128 region of cproc), J = band plane.
NHW_HDN void dec_c_markers_image(const DecImg &im)
{
	int16_t *P = im.cproc, *J = im.cjpeg;
	for (int r = 0; r < 256; r++)
		for (int j = (r < 128 ? 128 : 0); j < 256; j++) {
			const int s = r * CW + j, v = J[s];
			if (v <= 5000) continue;
			int t = s;
			if (r < 128) t -= 128;
			else t -= 32768 + (j < 128 ? 0 : 128);
			if (v == 5005) { P[t] -= 4; P[t + 1] -= 4; J[s] = 0; }
			else if (v == 5006) { P[t] += 4; P[t + 1] += 4; J[s] = 0; }
			else if (v == 5003) { P[t] -= 6; J[s] = 0; }
			else if (v == 5004) { P[t] += 6; J[s] = 0; }
		}
}

// ---- D17: YCbCr -> RGB (nhw_decoder_cli.c:139-229), IEEE-exact like the encoder's colour stage.
// mode 0: q>=20   1: q18,19 (Y scaled in float first)   2: q17   3: q<=16, float32 fixed-point form
struct DecColor { int mode; float y_inv; };

// the per-quality luma gain of write_image_bmp (nhw_decoder_cli.c:168-169,204,239-254); q0 has none (unsupported)
NHW_HD DecColor dec_color_of(int quality)
{
	DecColor c;
	c.mode = quality >= 20 ? 0 : quality >= 18 ? 1 : quality == 17 ? 2 : 3;
	// (a table, not a switch: see DESIGN.md on sparse switches in device code)
	const float gain[24] = {1.0f, 2.060881f, 1.985939f, 1.916257f, 1.820444f, 1.741126f, 1.665887f, 1.587597f, 1.521263f,
	                        1.392014f, 1.281502f, 1.190611f, 1.177434f, 1.186945f, 1.138331f, 1.048174f, 1.012139f,
	                        1.063830f, 1.075269f, 1.025641f, 1.0f, 1.0f, 1.0f, 1.0f};
	c.y_inv = gain[quality < 0 ? 0 : quality > 23 ? 23 : quality];
	return c;
}

#ifdef __CUDA_ARCH__
#define NHW_DMUL(a, b) __dmul_rn((a), (b))
#define NHW_DADD(a, b) __dadd_rn((a), (b))
#define NHW_DSUB(a, b) __dsub_rn((a), (b))
#define NHW_FMUL(a, b) __fmul_rn((a), (b))
#define NHW_D2I(a) __double2int_rz(a)
#define NHW_FADD(a, b) __fadd_rn((a), (b))
#define NHW_I2F(a) __int2float_rn(a)
#define NHW_F2I(a) __float2int_rz(a)
#else
#define NHW_FADD(a, b) ((a) + (b))
#define NHW_I2F(a) ((float)(a))
#define NHW_F2I(a) ((int)(a))
#define NHW_DMUL(a, b) ((a) * (b))
#define NHW_DADD(a, b) ((a) + (b))
#define NHW_DSUB(a, b) ((a) - (b))
#define NHW_FMUL(a, b) ((a) * (b))
#define NHW_D2I(a) ((int)(a))
#endif

// ---- integer form of the q >= 20 matrix.  R, G, B are trunc(Y + k1 U' + k2 V' + 0.5) of values that lie exactly on
// a 1e-3 (R, B) or 1e-5 (G) grid, and the double expression only leaves the exact value by ~1e-13: the truncation is
// the integer quotient, except where the exact value IS an integer (B at U' = +-125, G on a few (U', V') pairs), where
// the rounding of the partial products decides and the IEEE expression is evaluated.  Negative values clip to 0 either
// way.  Checked against the IEEE form on all 2^24 (Y, U, V) triples (nhw_debug_dec_color_check, tests/test_decode_gpu.py).
NHW_HD bool dec_ycc_to_rgb_q20_int(int y8, int u8, int v8, uint8_t *rgb)
{
	const int U = u8 - 128, V = v8 - 128;
	const int tr = 1000 * y8 + 500 + 1402 * V;
	const int tb = 1000 * y8 + 500 + 1772 * U;
	const int tg = 100000 * y8 + 50000 - 34414 * U - 71414 * V;
	if (U == 125 || U == -125) return false;
	const uint32_t qg = tg > 0 ? (uint32_t)(((unsigned long long)(uint32_t)tg * 2814749768ull) >> 48) : 0u;
	if (tg > 0 && qg * 100000u == (uint32_t)tg) return false;
	const uint32_t qr = tr > 0 ? (uint32_t)(((unsigned long long)(uint32_t)tr * 0x10624dd3ull) >> 38) : 0u;
	const uint32_t qb = tb > 0 ? (uint32_t)(((unsigned long long)(uint32_t)tb * 0x10624dd3ull) >> 38) : 0u;
	rgb[0] = (uint8_t)(qr > 255u ? 255u : qr);
	rgb[1] = (uint8_t)(qg > 255u ? 255u : qg);
	rgb[2] = (uint8_t)(qb > 255u ? 255u : qb);
	return true;
}

NHW_HD void dec_ycc_to_rgb_ieee(int y8, int u8, int v8, const DecColor &c, uint8_t *rgb);
NHW_HD void dec_ycc_to_rgb(int y8, int u8, int v8, const DecColor &c, uint8_t *rgb)
{
	if (c.mode == 0 && dec_ycc_to_rgb_q20_int(y8, u8, v8, rgb)) return;
	dec_ycc_to_rgb_ieee(y8, u8, v8, c, rgb);
}

NHW_HD void dec_ycc_to_rgb_ieee(int y8, int u8, int v8, const DecColor &c, uint8_t *rgb)
{
	if (c.mode == 3) {
		// q <= 16 (nhw_decoder_cli.c:256-276): integer matrix on un-centred U, V, one float32 multiply by the
		// luma gain, +128.5f, truncate, >> 8.  -56992-128, 34784-128, -70688-128 are R/G/B_COMP (decoder/codec.h:96-98).
		const int Y = y8 * 298;
		const int r = NHW_F2I(NHW_FADD(NHW_FMUL(NHW_I2F(Y + 409 * v8 - 57120), c.y_inv), 128.5f)) >> 8;
		const int g = NHW_F2I(NHW_FADD(NHW_FMUL(NHW_I2F(Y - 100 * u8 - 208 * v8 + 34656), c.y_inv), 128.5f)) >> 8;
		const int b = NHW_F2I(NHW_FADD(NHW_FMUL(NHW_I2F(Y + 516 * u8 - 70816), c.y_inv), 128.5f)) >> 8;
		rgb[0] = dec_clip8(r);
		rgb[1] = dec_clip8(g);
		rgb[2] = dec_clip8(b);
		return;
	}
	const double U = (double)(u8 - 128), V = (double)(v8 - 128);
	double Y = (double)y8;
	if (c.mode == 1) Y = (double)NHW_FMUL((float)y8, c.y_inv);
	double r = NHW_DADD(Y, NHW_DMUL(1.402, V));
	double g = NHW_DSUB(NHW_DSUB(Y, NHW_DMUL(0.34414, U)), NHW_DMUL(0.71414, V));
	double b = NHW_DADD(Y, NHW_DMUL(1.772, U));
	if (c.mode == 2) {
		const double k = (double)c.y_inv;
		r = NHW_DMUL(r, k);
		g = NHW_DMUL(g, k);
		b = NHW_DMUL(b, k);
	}
	rgb[0] = dec_clip8(NHW_D2I(NHW_DADD(r, 0.5)));
	rgb[1] = dec_clip8(NHW_D2I(NHW_DADD(g, 0.5)));
	rgb[2] = dec_clip8(NHW_D2I(NHW_DADD(b, 0.5)));
}
