// enc_ll_par.cuh -- parallel form of the luma LL2 coding (encoder/nhw_encoder.c:636-743 and
// Y_highres_compression, encoder/compress_pixel.c:471-876).
//
//  LL2 -> bytes : tagging rows (independent) -> nudge/untag wavefront (same footprint as the
//                 recons pass, skew 3) -> byte stores for every cell -> the rare escape cells in
//                 raster order by one thread.
//  DPCM coder   : one iteration of the reference's coding loop is a pure function of the position
//                 (ll_dpcm_step): it is evaluated for EVERY position in parallel, one thread then
//                 follows the chain of "next position" links through shared memory to find which
//                 positions are really visited and where their bytes go, and all threads emit.
//                 The code is written in its final (marker-stripped) form directly.
#pragma once
#include "enc_par.cuh"

// ---- LL2 -> bytes ---------------------------------------------------------------------------------
// quad tagging of one row + its nhw_res4 entries (column+1 of each tagged quad; last one +128,
// or a lone 128).  Returns the number of entries (1..32).
NHW_HD int ll2_bytes_tag_row(int16_t *P, int PS, int r, uint8_t *r4)
{
	int n = 0, c = r * PS;
	for (int j = 0; j < 125; j++, c++) {
		if (nhw_odd(P[c]) && nhw_odd(P[c + 1]) && nhw_odd(P[c + 2]) && nhw_odd(P[c + 3]) && nhw_iabs(P[c] - P[c + 3]) > 1) {
			P[c] += 24000; P[c + 1] += 16000; P[c + 2] += 16000; P[c + 3] += 16000;
			r4[n++] = (uint8_t)(j + 1);
			j += 3;
			c += 3;
		}
	}
	if (n == 0) { r4[0] = 128; return 1; }
	r4[n - 1] += 128;
	return n;
}

// wavefront cell: value of the LL2 sample after un-tagging / parity nudges; the band cell is zeroed
NHW_HD int ll2_bytes_cell(int16_t *P, int PS, int16_t *V, int q, int r, int j)
{
	const int a = r * PS + j;
	int scan = P[a];
	if (q > 17 && scan > 10000) scan -= scan > 20000 ? 24000 : 16000;
	else ll2_parity_nudge(P, PS, a, r, j, scan, q);
	V[r * 128 + j] = (int16_t)scan;
	P[a] = 0;
	return 1;
}

NHW_HD bool ll2_is_escape(int scan, int a) { return (scan > 255 || scan < 0) && a > 0; }

// ---- DPCM coder ------------------------------------------------------------------------------------
// run statistics of one run of equal neighbours starting at i (x[i]==x[i-1], x[i-1]!=x[i-2] or i==1):
// adds the run's contribution to (a8, y16) exactly like the reference's counting loop
// (compress_pixel.c:482-502), including its walk into the zero tail past the last sample.
NHW_HD void ll_stats_run(const uint8_t *x, int i, int N, int &a8, int &y16)
{
	int k = 0;
	while (i + k < N && x[i + k] == x[i + k - 1]) k++;
	int full = k >> 4, rem = k & 15;
	if (i + k >= N && rem != 0 && x[N - 1] == 0) { full++; rem = 0; }   // the count runs on through the zero tail
	y16 += full;
	a8 += full + (rem >= 8 ? 1 : 0);
}

struct LlStep { int next, nbytes, raw; uint8_t b[2]; };

// one iteration of the coding loop at position i (1 <= i < 16384), marker-stripped output
NHW_HD LlStep ll_dpcm_step(const uint8_t *x, int i, int mode, int q)
{
	LlStep s;
	s.nbytes = 1; s.raw = 0; s.b[0] = s.b[1] = 0;
	const int i0 = i;
	int scan = x[i] - x[i - 1];
	int count = x[i + 1] - x[i];
	auto raw = [&]() {          // 128, x[i]>>1, (x[i+1]>>1): only the last byte survives the strip pass
		if (q > 15) { s.b[0] = (uint8_t)(128 + (x[i + 1] >> 1)); s.raw = 1; i++; }
		else s.b[0] = (uint8_t)(128 + (x[i] >> 1));
	};
	auto triple = [&](int sc, int co, int e) {
		if (sc == 64 || co == 32 || e == 64) { raw(); return; }
		co >>= 1;
		s.b[0] = (uint8_t)(64 + sc + (co >> 3));
		s.b[1] = (uint8_t)(((co & 7) << 5) + (e >> 1));
		s.nbytes = 2;
		i += 2;
	};
	const bool tri_ok = nhw_iabs(x[i + 2] - x[i + 1]) <= 32 && i < 16382;
	const int e32 = x[i + 2] - x[i + 1] + 32;
	if (scan == 0 && count == 0) {
		int a = 0;
		if (mode == 0) {
			if (x[i + 2] == x[i + 1]) a = 1;
			i += a + 2;
			int code = a << 3;
			const int d = x[i] - x[i - 1], d2 = x[i + 1] - x[i];
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
			s.b[0] = (uint8_t)code;
		} else if (mode == 1) {
			while (x[i + a + 2] == x[i + a + 1]) { a++; if (a >= 7) break; }
			i += a + 2;
			int code = a << 2;
			const int d = x[i] - x[i - 1];
			if (d == 2) code += 1;
			else if (d == -2) code += 2;
			else if (d == 0) code += 3;
			else i--;
			s.b[0] = (uint8_t)code;
		} else {
			while (x[i + a + 2] == x[i + a + 1]) { a++; if (a >= 63) break; }
			i += a + 1;
			s.b[0] = (uint8_t)a;
		}
	} else if (mode == 0 && nhw_iabs(scan) <= 6 && nhw_iabs(count) <= 8) {
		scan += 6; count += 8;
		if (scan == 12 || count == 16) {
			if (tri_ok) triple(scan + 26, count + 8, e32);
			else raw();
		} else {
			if (scan < 8) s.b[0] = (uint8_t)(32 + (scan << 2) + (count >> 1));
			else if (scan == 8) s.b[0] = (uint8_t)(16 + (count >> 1));
			else s.b[0] = (uint8_t)(24 + (count >> 1));
			i++;
		}
	} else if (mode == 1 && nhw_iabs(scan) <= 4 && nhw_iabs(count) <= 8) {
		scan += 4; count += 8;
		if (scan == 8 || count == 16) {
			if (tri_ok) triple(scan + 28, count + 8, e32);
			else raw();
		} else {
			s.b[0] = (uint8_t)(32 + (scan << 2) + (count >> 1));
			i++;
		}
	} else if (nhw_iabs(scan) <= 32 && nhw_iabs(count) <= 16 && tri_ok) {
		triple(scan + 32, count + 16, e32);
	} else raw();
	s.next = i + 1;   // the for-loop's i++
	(void)i0;
	return s;
}
