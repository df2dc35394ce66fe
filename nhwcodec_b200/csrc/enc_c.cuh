// enc_c.cuh -- chroma encoder stages (encoder/nhw_encoder.c:2255-2868; U and V run the same
// skeleton on 256x256 planes, stride 256, with two small differences noted below), the chroma
// quantisers offsetUV_recons256 / offsetUV (encoder/image_processing.c:3192-3353, 108-183)
// and the chroma LL coder highres_compression (encoder/compress_pixel.c:878-1022).
// The q <= 16 additions (pre_processing_UV, thresholds, LL smoothing) are in enc_lowq.cuh.
#pragma once
#include "enc_y3.cuh"

#define CW 256   // chroma row stride

// ---- offsetUV_recons256: LL (64x64) part.  comp=1 first call, comp=0 second call.
NHW_HD void c_recons_ll_row(const EncImg &im, int r /* 0..63 */, int comp, int q = 20)
{
	const int16_t *P = im.cproc + r * CW;
	int16_t *J = im.cjpeg + r * CW;
	if (comp && q <= 15) {   // the decoder adds 1 to every chroma LL sample below q16 (nhw_decoder.c:953-962)
		for (int j = 0; j < 64; j++) J[j] = (int16_t)((P[j] & 65532) + 1);
	} else if (comp) {
		for (int j = 0; j < 64; j += 2) {
			if (r == 0) { J[j] = P[j]; J[j + 1] = (int16_t)(P[j + 1] & 65534); }
			else { J[j] = (int16_t)(P[j] & 65534); J[j + 1] = P[j + 1]; }
		}
	} else {
		for (int j = 0; j < 64; j++) J[j] = (P[j] > 0 && P[j] < 256) ? (int16_t)(P[j] & 65534) : P[j];
	}
}

// ---- offsetUV_recons256: detail rows of the 128x128 level-2 region
NHW_HD void c_recons_quant_row(const EncImg &im, int r /* 0..127 */, int m1, int comp)
{
	const int16_t *P = im.cproc + r * CW;
	int16_t *J = im.cjpeg + r * CW;
	for (int j = r < 64 ? 64 : 0; j < 128; j++) {
		int a = P[j];
		if ((a == -7 || a == -8) && !comp) {
			if (j < 127 && (P[j + 1] == -7 || P[j + 1] == -8)) { J[j] = -11; J[j + 1] = -11; j++; continue; }
		}
		if (a < 0) {
			a = -a;
			if (P[j + 1] < 0 && P[j + 1] > -8) { if ((a & 7) < 6) a &= 65528; }
			else if ((a & 7) < 7) a &= 65528;
			a = -a;
		}
		if (a < m1 && a > -m1) { J[j] = 0; continue; }
		a += 128;
		if (a < 0) a = -((-a) & 65528);
		else a &= 65528;
		J[j] = (int16_t)(a > 128 ? a - 125 : a - 131);
	}
}

// ---- chroma LL1 correction against the trial reconstruction (nhw_encoder.c:2316-2335,
// 2629-2647).  U tests the right neighbour with >=0 / <=0, V with >0 / <0.
NHW_HD void c_correct_row(const EncImg &im, int r /* 0..127 */, int is_v)
{
	const int16_t *P = im.cproc + r * CW, *L = im.cll1 + r * 128;
	int16_t *J = im.cjpeg + r * CW;
	for (int j = 0; j < 128; j++) {
		int scan = P[j] - L[j];
		int nx = P[j + 1] - L[j + 1];
		int d = 0;
		if (scan > 10) d = -6;
		else if (scan > 7) d = -3;
		else if (scan > 4) d = -2;
		else if (scan > 3) d = -1;
		else if (scan > 2 && (is_v ? nx > 0 : nx >= 0)) d = -1;
		else if (scan < -10) d = 6;
		else if (scan < -7) d = 3;
		else if (scan < -4) d = 2;
		else if (scan < -3) d = 1;
		else if (scan < -2 && (is_v ? nx < 0 : nx <= 0)) d = 1;
		J[j] = (int16_t)(L[j] + d);
	}
}

// ---- residual tags 12400 / 12600 / 12900 / 13000 dropped into the first free (|c|<8) cell of
// the three level-1 bands (nhw_encoder.c:2372-2424), q>=18.  Sequential along a row.
NHW_HD bool c_drop_tag(int16_t *P, int scan, int tag)
{
	if (nhw_iabs(P[scan + 128]) < 8) { P[scan + 128] = (int16_t)tag; return true; }
	if (nhw_iabs(P[scan + 32768]) < 8) { P[scan + 32768] = (int16_t)tag; return true; }
	if (nhw_iabs(P[scan + 32768 + 128]) < 8) { P[scan + 32768 + 128] = (int16_t)tag; return true; }
	return false;
}

NHW_HD void c_residual_tags_row(const EncImg &im, int q, int r /* 0..127 */)
{
	if (q < 18) return;
	int16_t *P = im.cproc;
	const int16_t *L = im.cll1;
	const int res_uv = q > 17 ? 4 : 5;
	int scan = r * CW, count = r * 128;
	for (int j = 0; j < 128; j++, scan++, count++) {
		int d = P[scan] - L[count];
		if (d > 3 && d < 7) {
			int n = P[scan + 1] - L[count + 1];
			if (n > 2 && n < 7) {
				if (c_drop_tag(P, scan, 12400)) { count++; scan++; j++; continue; }
			}
		} else if (d < -3 && d > -7) {
			int n = P[scan + 1] - L[count + 1];
			if (n < -2 && n > -8) {
				if (c_drop_tag(P, scan, 12600)) { count++; scan++; j++; continue; }
			}
		}
		if (nhw_iabs(d) > res_uv) {
			if (d > 0) c_drop_tag(P, scan, 12900);
			else if (d == -5) { if ((P[scan + 1] - L[count + 1]) < 0) c_drop_tag(P, scan, 13000); }
			else c_drop_tag(P, scan, 13000);
		}
	}
}

// ---- chroma LL (64x64) -> tree1 bytes + exw escapes (nhw_encoder.c:2482-2515, 2781-2813).
// exw entries go to a per-component list (im.exw_uv for U, im.exw_uv+16384 for V) that the
// container writer splices after the luma list with the two-zero separators.
NHW_HDN int c_ll_to_bytes_image(const EncImg &im, int is_v)
{
	int16_t *P = im.cproc;
	uint8_t *t = im.tree1;
	uint8_t *exw = im.exw_uv + (is_v ? 16384 : 0);
	int a = is_v ? 20480 : 16384, e = 0;
	for (int r = 0; r < 64; r++) {
		for (int j = 0; j < 64; j++) {
			int scan = P[r * CW + j];
			if (scan > 255 && (j > 0 || r > 0)) {
				exw[e++] = (uint8_t)r; exw[e++] = (uint8_t)(j + 128);
				int y = scan - 255;
				exw[e++] = (uint8_t)(y > 255 ? 255 : y);
				t[a] = t[a - 1]; a++;
			} else if (scan < 0 && (j > 0 || r > 0)) {
				exw[e++] = (uint8_t)r; exw[e++] = (uint8_t)j;
				if (scan < -255) scan = -255;
				exw[e++] = (uint8_t)(-scan);
				t[a] = t[a - 1]; a++;
			} else {
				if (scan > 255) scan = 255;
				else if (scan < 0) scan = 0;
				t[a++] = (uint8_t)(scan & 254);
			}
			P[r * CW + j] = 0;
		}
	}
	return e;
}

// bit 1 of every chroma LL byte, 8 per byte (res_U_64 / res_V_64, nhw_encoder.c:2517-2537), q>15
NHW_HD void c_ll_bit1_plane(const EncImg &im, int is_v)
{
	const uint8_t *t = im.tree1 + (is_v ? 20480 : 16384);
	uint8_t *o = im.res_uv64 + (is_v ? 512 : 0);
	for (int k = 0; k < 512; k++) {
		int b = 0;
		for (int i = 0; i < 8; i++) b |= ((t[8 * k + i] >> 1) & 1) << (7 - i);
		o[k] = (uint8_t)b;
	}
}

// ---- highres_compression (compress_pixel.c:878-1022): both chroma LL planes, appended to
// the luma LL code.  in: tree1[16384..24575]; io: im.llcode (highres_comp) from y_res_comp on.
// Core: x = the LL bytes indexed as in tree1 (x[16384..24575] valid and already masked with 252, readable up
// to x[24575 + 20]); out[j0..] receives the code; returns the end index.
// One iteration of the coder's loop at position i: it emits exactly one byte and moves on by 1..17 positions, and it
// carries no state from one iteration to the next -- a pure function of the position (the parallel form,
// k_c_ll_code, evaluates it everywhere and then finds the positions the coder really visits).
NHW_HD int c_ll_step(const uint8_t *x, int i, int &byte)
{
	int scan = x[i] - x[i - 1];
	int count = x[i + 1] - x[i];
	if (scan == 0 && count == 0) {
		int a = 0, res = 0;
		while (x[i + a + 2] == x[i + a + 1]) {
			a++;
			if (a < 7) continue;
			res = 1;             // a==7 || res==1
			if (a >= 14) break;
		}
		i += a + 1;
		if (res == 1) byte = 64 + (7 << 3) + a - 7;
		else {
			i++;
			int code = 64 + (a << 3);
			int d = x[i] - x[i - 1], d2 = x[i + 1] - x[i];
			if (d == 4) {
				if (d2 == -4) {
					if (x[i + 2] - x[i + 1] == 0) { code += 3; i += 2; }
					else { code += 2; i++; }
				} else code += 1;
			} else if (d == -4) {
				if (d2 == 4) {
					if (x[i + 2] - x[i + 1] == 0) { code += 4; i += 2; }
					else { code += 5; i++; }
				} else code += 6;
			} else if (d == 8) code += 7;
			else i--;
			byte = code;
		}
	} else if (nhw_iabs(scan) <= 4 && nhw_iabs(count) <= 4) {
		int res = 0;
		if (!scan && count == 4) res = 0;
		else if (!scan && count == -4) res = 1;
		else if (scan == 4 && !count) res = 2;
		else if (scan == -4 && !count) res = 3;
		else if (scan == 4 && count == 4) res = 4;
		else if (scan == 4 && count == -4) res = 5;
		else if (scan == -4 && count == 4) res = 6;
		else if (scan == -4 && count == -4) res = 7;
		int d3 = x[i + 2] - x[i + 1];
		if (d3 == 0) { byte = 192 + (res << 2); i += 2; }
		else if (d3 == 4) { byte = 192 + (res << 2) + 1; i += 2; }
		else if (d3 == -4) { byte = 192 + (res << 2) + 2; i += 2; }
		else if (d3 == 8) { byte = 192 + (res << 2) + 3; i += 2; }
		else { byte = ((scan + 16) << 1) + ((count + 16) >> 2); i++; }
	} else if (nhw_iabs(scan) <= 16 && nhw_iabs(count) <= 16) {
		scan += 16; count += 16;
		if (scan == 32 || count == 32) byte = 128 + (x[i] >> 2);
		else { byte = (scan << 1) + (count >> 2); i++; }
	} else byte = 128 + (x[i] >> 2);
	byte &= 255;
	return i + 1;
}
NHW_HDN int ll_dpcm_chroma_core(const uint8_t *x, uint8_t *out, int j0)
{
	int j = j0;
	out[j++] = x[16384];
	for (int i = 16385; i < 24576;) {
		int byte;
		i = c_ll_step(x, i, byte);
		out[j++] = (uint8_t)byte;
	}
	return j;
}

NHW_HDN void ll_dpcm_chroma_image(const EncImg &im)
{
	for (int i = 16384; i < 24576; i++) im.tree1[i] &= 252;
	im.hdr->end_ch_res = ll_dpcm_chroma_core(im.tree1, im.llcode, im.hdr->y_res_comp);
}
