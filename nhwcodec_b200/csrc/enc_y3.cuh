// enc_y3.cuh -- luma: quantisation of the coefficient plane to the byte alphabet
// (offsetY, encoder/image_processing.c:185-521, q>16 branches), the serpentine scan
// (encoder/nhw_encoder.c:2108-2132) and the three peephole passes over the byte stream
// (encoder/nhw_encoder.c:2136-2252).
#pragma once
#include "enc_y2.cuh"

// The two escape ladders (extra_words1/2, encoder/tree.h:54-55) are not perfectly regular, so
// they stay tables: a __constant__ copy for the kernels, a host copy for the test harness.
#define NHW_EXTRA_WORDS1 {10, 12, 14, 18, 20, 22, 26, 28, 30, 34, 36, 38, 42, 44, 46, 50, 52, 54, 58}
#define NHW_EXTRA_WORDS2 {60, 62, 66, 68, 70, 74, 76, 78, 82, 84, 86, 90, 92, 94, 98, 100, 102, 106, 108}
#ifdef __CUDACC__
__constant__ uint8_t c_extra_words1[19] = NHW_EXTRA_WORDS1;
__constant__ uint8_t c_extra_words2[19] = NHW_EXTRA_WORDS2;
#endif
static const uint8_t h_extra_words1[19] = NHW_EXTRA_WORDS1;
static const uint8_t h_extra_words2[19] = NHW_EXTRA_WORDS2;
#ifdef __CUDA_ARCH__
#define NHW_EXTRA1(k) c_extra_words1[k]
#define NHW_EXTRA2(k) c_extra_words2[k]
#else
#define NHW_EXTRA1(k) h_extra_words1[k]
#define NHW_EXTRA2(k) h_extra_words2[k]
#endif

// ---- offsetY loop 4 (image_processing.c:312-519): coefficient -> byte
NHW_HDN void y_offset_quant_image(const EncImg &im, int m1)
{
	int16_t *P = im.proc;
	for (int i = 0; i < 4 * 65536; i++) {
		const bool inrow = (i & 511) < 511;
		int a = P[i];
		if (a > 10000) {
			int b = a == 10100 ? 128 : a == 12700 ? 127 : a == 12900 ? 129 : a == 10204 ? 125 : a == 10300 ? 126 :
			        a == 12100 ? 121 : a == 12200 ? 122 : -1;
			if (b >= 0) { P[i] = (int16_t)b; continue; }
		}
		if (a > 127) {
			int k = ((a & 0xfff8) - 128) >> 3;
			P[i] = NHW_EXTRA1(k > 18 ? 18 : k);
			continue;
		} else if (a < -127) {
			int k = (((-a) & 0xfff8) - 128) >> 3;
			P[i] = NHW_EXTRA2(k > 18 ? 18 : k);
			continue;
		}
		if (a < -12 && ((-a) & 7) == 6) {
			if (inrow && P[i + 1] == -7) P[i + 1] = -9;
		}
		if (a < 0) {
			if (a == -7 && P[i + 1] == 8 && inrow) { P[i] = -8; a = -8; }
			a = -a;
			if (a > 14 && (a & 7) == 7 && P[i + 1] > 0 && P[i + 1] < 8) a -= 2;
			if ((a & 7) < 7) a &= 504;
			a = -a;
		} else if (a == 8 && P[i + 1] == -7 && inrow) P[i + 1] = -8;
		else if (a > 12 && (a & 7) >= 6) {
			if (inrow && P[i + 1] == 7) P[i + 1] = 9;
		}
		if (a < m1 && a > -m1) { P[i] = 128; continue; }
		P[i] = (int16_t)((a + 128) & 248);
	}
}

// ---- E23: serpentine scan, 4-column strips, two rows per step, second row reversed
NHW_HD void y_scan_strip(const EncImg &im, int strip /* 0..127 */)
{
	const int16_t *P = im.proc + strip * 4;
	uint8_t *s = im.scan + strip * 2048;
	for (int k = 0; k < 256; k++) {
		const int16_t *r0 = P + (2 * k) * YW, *r1 = r0 + YW;
		s[0] = (uint8_t)r0[0]; s[1] = (uint8_t)r0[1]; s[2] = (uint8_t)r0[2]; s[3] = (uint8_t)r0[3];
		s[4] = (uint8_t)r1[3]; s[5] = (uint8_t)r1[2]; s[6] = (uint8_t)r1[1]; s[7] = (uint8_t)r1[0];
		s += 8;
	}
}
