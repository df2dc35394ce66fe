// enc_ll.cuh -- coding of the lowest band: luma LL2 (128x128) and chroma LL (2 x 64x64).
//
//   y_ll2_to_bytes_image : nhw_encoder.c:636-743   LL2 -> tree1 / ch_res / exw_Y / nhw_res4
//   ll_dpcm_luma_image   : Y_highres_compression, compress_pixel.c:471-876
//   ll_dpcm_chroma_image : highres_compression,   compress_pixel.c:878-1022
//
// All of it is strictly sequential with look-ahead (a code is chosen from the next two
// differences and consumes a variable number of samples): one thread per image.
#pragma once
#include "enc_y1.cuh"

// ---- E10 + E11: LL2 plane -> byte list.  Values outside 0..255 go to the exw_Y escape list;
// rows of four odd samples in a row are tagged and announced in nhw_res4 (q>17).
// P = LL2 band at row stride PS (see y_recons_ll2_core)
NHW_HDN void y_ll2_to_bytes_core(const EncImg &im, int16_t *P, int PS, int q)
{
	EncHdr *h = im.hdr;
	int res = 0;
	if (q > 17) {
		for (int r = 0; r < 128; r++) {
			int c = r * PS, stage = 0;
			for (int j = 0; j < 125; j++, c++) {
				if (nhw_odd(P[c]) && nhw_odd(P[c + 1]) && nhw_odd(P[c + 2]) && nhw_odd(P[c + 3]) &&
				    nhw_iabs(P[c] - P[c + 3]) > 1) {
					P[c] += 24000; P[c + 1] += 16000; P[c + 2] += 16000; P[c + 3] += 16000;
					res++; stage++; j += 3; c += 3;
				}
			}
			if (!stage) res++;
		}
		h->res4_len = res;
	}
	int a = 0, e = 0;
	res = 0;
	for (int r = 0; r < 128; r++) {
		int stage = 0;
		int c = r * PS;
		for (int j = 0; j < 128; j++, c++) {
			int scan = P[c];
			if (q > 17 && scan > 10000) {
				if (scan > 20000) { scan -= 24000; im.res4[res++] = (uint8_t)(j + 1); stage++; }
				else scan -= 16000;
			} else if (nhw_odd(scan) && j > 0 && nhw_odd(P[c + 1])) {
				if (j < 126 && nhw_odd(P[c + 2])) {
					if (nhw_iabs(scan - P[c + 2]) > 1 && q > 17) P[c + 1]++;
				} else if (r < 127 && nhw_odd(P[c + PS]) && nhw_odd(P[c + PS + 1]) && !nhw_odd(P[c + PS + 2])) {
					if (P[c + PS] < 10000 && q > 17) P[c + PS]++;
				}
			} else if (nhw_odd(scan) && r >= 1 && r < 125) {
				if (nhw_odd(P[c + PS]) && nhw_odd(P[c + PS + 1])) {
					if (nhw_odd(P[c + 2 * PS]) && !nhw_odd(P[c + 3 * PS])) {
						if (P[c + PS] < 10000 && q > 17) P[c + PS]++;
					}
				}
			}
			if (scan > 255 && (j > 0 || r > 0)) {
				im.exw[e++] = (uint8_t)r;
				im.exw[e++] = (uint8_t)(j + 128);
				int y = scan - 255;
				if (y > 255) y = 255;
				im.exw[e++] = (uint8_t)y;
				im.tree1[a] = im.tree1[a - 1];
				im.ch_res[a] = im.tree1[a - 1];
				a++;
			} else if (scan < 0 && (j > 0 || r > 0)) {
				im.exw[e++] = (uint8_t)r;
				im.exw[e++] = (uint8_t)j;
				if (scan < -255) scan = -255;
				im.exw[e++] = (uint8_t)(-scan);
				im.tree1[a] = im.tree1[a - 1];
				im.ch_res[a] = im.tree1[a - 1];
				a++;
			} else {
				if (scan > 255) scan = 255;
				else if (scan < 0) scan = 0;
				im.ch_res[a] = (uint8_t)scan;
				im.tree1[a++] = (uint8_t)(scan & 254);
			}
			P[c] = 0;
		}
		if (q > 17) {
			if (!stage) im.res4[res++] = 128;
			else im.res4[res - 1] += 128;
		}
	}
	h->exw_y_len = e;
}

NHW_HDN void y_ll2_to_bytes_image(const EncImg &im, int q) { y_ll2_to_bytes_core(im, im.proc, YW, q); }

NHW_HDN int ll_dpcm_luma_core(const EncImg &im, const uint8_t *x, uint8_t *work, int q);
NHW_HDN int ll_dpcm_luma_image(const EncImg &im, int q) { return ll_dpcm_luma_core(im, im.tree1, im.tmp1, q); }

// ---- shared pieces of the three DPCM modes (compress_pixel.c:511-830) ----
struct LlCoder {
	const uint8_t *x;      // samples (tree1), readable a few bytes past the end (zero there)
	uint8_t *out;          // code bytes
	const uint8_t *full;   // un-truncated samples (ch_res of E11) for highres_word
	uint8_t *word;         // highres_word
	uint16_t *mem;         // highres_mem
	int j, nmem, q;
};

// raw escape: 128, then the two samples halved (q>15 also remembers the dropped LSB's sample)
NHW_HD void ll_emit_raw(LlCoder &c, int &i)
{
	c.out[c.j++] = 128;
	c.out[c.j++] = (uint8_t)(128 + (c.x[i] >> 1));
	if (c.q > 15) {
		c.out[c.j++] = (uint8_t)(128 + (c.x[i + 1] >> 1));
		c.word[c.nmem] = c.full[i];
		c.mem[c.nmem++] = (uint16_t)i;
		i++;
	}
}

// three-difference code (marker 64): `scan`,`count`,`e` are already biased
NHW_HD void ll_emit_triple(LlCoder &c, int &i, int scan, int count, int e)
{
	if (scan == 64 || count == 32 || e == 64) { ll_emit_raw(c, i); return; }
	count >>= 1;
	c.out[c.j++] = 64;
	c.out[c.j++] = (uint8_t)(64 + scan + (count >> 3));
	c.out[c.j++] = (uint8_t)(((count & 7) << 5) + (e >> 1));
	i += 2;
}

NHW_HD bool ll_triple_ok(const LlCoder &c, int i) { return nhw_iabs(c.x[i + 2] - c.x[i + 1]) <= 32 && i < 16382; }

// luma LL2: chooses RES_LOW mode 0/1/2 from run statistics, then codes.  Returns the mode.
// x = the 16384 LL2 bytes followed by at least 32 readable zero bytes (tree1, or a shared-memory
// copy); work = >= 24640 bytes of scratch for the marked code (tmp1, or shared memory).
NHW_HDN int ll_dpcm_luma_core(const EncImg &im, const uint8_t *x, uint8_t *work, int q)
{
	EncHdr *h = im.hdr;
	const int N = 16384;
	// run statistics (compress_pixel.c:482-502)
	int e = 0, Y = 0, a = 0;
	for (int i = 1; i < N; i++) {
		while (x[i] == x[i - 1]) {
			e++;
			if (e < 16) { if (e == 8) a++; i++; continue; }
			if (e == 16) Y++;
			break;
		}
		e = 0;
	}
	a += Y;
	int mode = Y > 299 ? 2 : (a > 179 ? 1 : 0);

	LlCoder c;
	c.x = x; c.out = work; c.full = im.ch_res; c.word = im.highres_word; c.mem = im.highres_mem;
	c.j = 1; c.nmem = 0; c.q = q;
	c.out[0] = x[0];
	a = 0;
	for (int i = 1; i < N; i++) {
		int scan = x[i] - x[i - 1];
		int count = x[i + 1] - x[i];
		if (scan == 0 && count == 0) {
			if (mode == 0) {
				if (x[i + a + 2] == x[i + a + 1]) a++;
				i += a + 2;
				int code = a << 3;
				int d = x[i] - x[i - 1], d2 = x[i + 1] - x[i];
				if (d == 2) {
					if (d2 == -2) { code += 2; i++; }
					else if (d2 == 0) { code += 3; i++; }
					else code += 1;
				} else if (d == -2) {
					if (d2 == 2) { code += 4; i++; }
					else if (d2 == 0) { code += 5; i++; }
					else code += 6;
				} else if (d == 4) code += 7;
				else i--;
				c.out[c.j++] = (uint8_t)code;
			} else if (mode == 1) {
				while (x[i + a + 2] == x[i + a + 1]) { a++; if (a >= 7) break; }
				i += a + 2;
				int code = a << 2;
				int d = x[i] - x[i - 1];
				if (d == 2) code += 1;
				else if (d == -2) code += 2;
				else if (d == 0) code += 3;
				else i--;
				c.out[c.j++] = (uint8_t)code;
			} else {
				while (x[i + a + 2] == x[i + a + 1]) { a++; if (a >= 63) break; }
				i += a + 1;
				c.out[c.j++] = (uint8_t)a;
			}
			a = 0;
		} else if (mode == 0 && nhw_iabs(scan) <= 6 && nhw_iabs(count) <= 8) {
			scan += 6; count += 8;
			if (scan == 12 || count == 16) {
				if (ll_triple_ok(c, i)) ll_emit_triple(c, i, scan + 26, count + 8, x[i + 2] - x[i + 1] + 32);
				else ll_emit_raw(c, i);
			} else {
				if (scan < 8) c.out[c.j++] = (uint8_t)(32 + (scan << 2) + (count >> 1));
				else if (scan == 8) c.out[c.j++] = (uint8_t)(16 + (count >> 1));
				else c.out[c.j++] = (uint8_t)(24 + (count >> 1));
				i++;
			}
		} else if (mode == 1 && nhw_iabs(scan) <= 4 && nhw_iabs(count) <= 8) {
			scan += 4; count += 8;
			if (scan == 8 || count == 16) {
				if (ll_triple_ok(c, i)) ll_emit_triple(c, i, scan + 28, count + 8, x[i + 2] - x[i + 1] + 32);
				else ll_emit_raw(c, i);
			} else {
				c.out[c.j++] = (uint8_t)(32 + (scan << 2) + (count >> 1));
				i++;
			}
		} else if (nhw_iabs(scan) <= 32 && nhw_iabs(count) <= 16 && ll_triple_ok(c, i)) {
			ll_emit_triple(c, i, scan + 32, count + 16, x[i + 2] - x[i + 1] + 32);
		} else {
			ll_emit_raw(c, i);
		}
	}
	// strip the 64 / 128 markers (compress_pixel.c:832-866); result is highres_comp -> im.llcode
	const int j = c.j;
	uint8_t *dst = im.llcode;
	dst[0] = c.out[0];
	int o = 1, i = 1;
	for (; i < j - 1; i++) {
		if (c.out[i] == 64) { dst[o++] = c.out[i + 1]; dst[o++] = c.out[i + 2]; i += 2; }
		else if (c.out[i] == 128) {
			if (q > 15) { i++; dst[o++] = c.out[i + 1]; i++; }
			else { i++; dst[o++] = c.out[i]; }
		} else dst[o++] = c.out[i];
	}
	if (i < j) dst[o++] = c.out[j - 1];
	h->highres_comp_len = c.nmem;
	h->highres_mem_len = c.nmem;
	h->res_low = mode;           // setup->RES_LOW; becomes RES_HIGH in highres_compression
	h->y_res_comp = o;
	return mode;
}
