// enc_lowq.cuh -- encoder stages that only exist at the low quality settings (q <= 16):
//
//   y_e7_kill_row        nhw_encoder.c:285-309    q<=11  isolated small coefficients of one level-2 band
//   y_e8_smooth_image    nhw_encoder.c:311-621    q<=12  LL2 smoothing + zeroing of descendants
//   y_e14_lowq_image     nhw_encoder.c:804-967    q<=15  level-1 band thresholds
//   y_offset_quant_lowq_image  image_processing.c:312-519 (q<=16 arms)  coefficient -> byte with the cyclic counters
//   c_pre_uv_cell        image_processing.c:2428-2464  q<=14  chroma pre-filter
//   c_thresholds_cell    nhw_encoder.c:2277-2308  q<=16  chroma level-1 band thresholds
//   c_ll_smooth_image    nhw_encoder.c:2438-2478  q<=11  chroma LL smoothing
//
// These are raster-ordered, in-place passes over flat plane memory (a neighbour at the end of a row is the first
// cell of the next one): *_image functions are one walk per image, *_row / *_cell functions are independent units.
// The batch is the parallel axis for the *_image forms.
#pragma once
#include "enc_c.cuh"

// ---- E7: rows 128..255 x cols 0..255 of the plane (one level-2 detail band pair), in place along the row:
// the left neighbour is read after its own turn, the right one before.  Column 0's left neighbour is the previous
// row's last cell (a level-1 cell this stage never writes), so rows are independent.
NHW_HD void y_e7_kill_row(const EncImg &im, int q, int ratio, int r /* 128..255 */)
{
	const int top = q > 6 ? 10 : 11;
	int16_t *P = im.proc + r * YW;
	for (int j = 0; j < 256; j++) {
		const int v = nhw_iabs(P[j]);
		if (v < ratio || v >= top) continue;
		const bool l = nhw_iabs(P[j - 1]) < ratio, rt = nhw_iabs(P[j + 1]) < ratio;
		if (l && rt) P[j] = 0;
		else if (v == ratio && (l || rt)) P[j] = 0;
	}
}

// ---- E8.  A smoothed LL2 sample also silences its descendants: the 2x2 cells below it in the three level-1
// bands, and (q <= 11) the co-located cells of the three level-2 detail bands.
struct E8Thr { int t1, t2, t3, t4, t5, t6, t7; };
NHW_HD E8Thr e8_thresholds(int q)
{
	E8Thr t = {8, 13, 6, 11, 34, 14, 0};                      // q12
	if (q <= 11 && q >= 8) { t.t6 = 15; t.t7 = 15; }
	else if (q == 7) t = E8Thr{10, 15, 9, 14, 36, 17, 17};
	else if (q <= 6 && q >= 4) t = E8Thr{11, 15, 10, 15, 36, 17, 17};
	else if (q == 3) t = E8Thr{11, 15, 10, 15, 36, 18, 18};
	else if (q == 2) t = E8Thr{11, 15, 10, 15, 36, 19, 20};
	else if (q == 1) t = E8Thr{11, 15, 10, 15, 36, 20, 21};
	return t;
}
NHW_HD void e8_kill(int16_t *P, int at, int below) { if (nhw_iabs(P[at]) < below) P[at] = 0; }
// the twelve level-1 descendants of LL2 cell p (flat plane index): 2x2 blocks at twice the coordinates in the
// band right of the level-2 region (a), below it (b) and diagonal (c)
NHW_HD void e8_silence_children(int16_t *P, int p, int a, int b, int c)
{
	const int base = p << 1;
	e8_kill(P, base + 256, a); e8_kill(P, base + 257, a); e8_kill(P, base + 768, a); e8_kill(P, base + 769, a);
	e8_kill(P, base + 131072, b); e8_kill(P, base + 131073, b); e8_kill(P, base + 131584, b); e8_kill(P, base + 131585, b);
	e8_kill(P, base + 131328, c); e8_kill(P, base + 131329, c); e8_kill(P, base + 131840, c); e8_kill(P, base + 131841, c);
}
NHW_HD void e8_silence_siblings(int16_t *P, int p)
{
	e8_kill(P, p + 128, 11); e8_kill(P, p + 65536, 12); e8_kill(P, p + 65536 + 128, 13);
}

// `cursor` models the reference's `count` variable: it is assigned inside some branches only and read, stale, by
// the q <= 11 tail of the third pass, so it is threaded through all four passes (65536 on entry: the value the
// LL1 copy / correction loop leaves behind).
// B = the LL2 band (128 x 128) at row stride BS -- the plane itself (BS = 512) or a staged copy.  The walk never reads
// the cells it silences, and silencing is idempotent ("zero if below a threshold"; of two requests for the same cell
// the larger threshold wins), so the requests go to a sink: E8Direct applies them to the plane at once (serial form),
// the CUDA kernel records them and applies them with all its lanes after the walk (E8Deferred, encode.cu).
struct E8Direct {
	int16_t *P;
	NHW_HD void children(int p, int a, int b, int c) const { e8_silence_children(P, p, a, b, c); }
	NHW_HD void siblings(int p) const { e8_silence_siblings(P, p); }
};
template <typename Sink>
NHW_HDN void y_e8_smooth_walk(int16_t *B, int BS, const Sink &P_, int q)
{
	const E8Thr t = e8_thresholds(q);
	const bool deep = q <= 11;
	int cursor = 65536;
	// pass 1: five-sample windows along each LL2 row
	for (int r = 0; r < 128; r++)
		for (int j = 0, s = r * YW, b = r * BS; j < 124; j++, s++, b++) {
			const int v0 = B[b], v1 = B[b + 1], v2 = B[b + 2], v3 = B[b + 3], v4 = B[b + 4];
			bool hit = false;
			if (nhw_iabs(v4 - v0) < t.t1 && nhw_iabs(v4 - v3) < t.t1 && nhw_iabs(v1 - v0) < t.t1 && nhw_iabs(v3 - v1) < t.t1 &&
			    nhw_iabs(v3 - v2) < t.t2 - 2) {
				const int up = v3 - v1;             // slope across the middle sample
				if (up > 5 && v2 >= v3) B[b + 2] = (int16_t)v3;
				else if (-up > 5 && v2 <= v3) B[b + 2] = (int16_t)v3;
				else if (-up > 5 && v2 >= v1) B[b + 2] = (int16_t)v1;
				else if (up > 5 && v2 <= v1) B[b + 2] = (int16_t)v1;
				else if (v3 > v2 && v2 > v1) {}
				else if (v1 > v2 && v2 > v3) {}
				else B[b + 2] = (int16_t)((v3 + v1) >> 1);
				hit = true;
			} else if (nhw_iabs(v4 - v0) < t.t2 + 1 && nhw_iabs(v4 - v3) < t.t2 + 1 && nhw_iabs(v1 - v0) < t.t2 + 1) {
				if (nhw_iabs(v3 - v1) < t.t2 + 6 && nhw_iabs(v3 - v2) < t.t2 + 6)
					hit = (v3 >= v2 && v2 >= v1) || (v3 <= v2 && v2 <= v1);
			}
			if (!hit) continue;
			for (int k = 1; k < 4; k++) P_.children(s + k, t.t6, t.t6 + 6, t.t5);
			cursor = 4;
			if (deep)
				for (int k = 1; k < 4; k++) P_.siblings(s + k);
		}
	// passes 2 and 3: the centre of a plus-shaped neighbourhood becomes the rounded mean of its four arms
	for (int pass = 0; pass < 2; pass++)
		for (int r = 0; r < 126; r++)
			for (int j = 0, s = r * YW, b = r * BS; j < 126; j++, s++, b++) {
				const int up = B[b + 1], dn = B[b + 2 * BS + 1], lf = B[b + BS], rt = B[b + BS + 2], ce = B[b + BS + 1];
				bool outer, inner;
				if (pass == 0) {
					outer = nhw_iabs(up - dn) < t.t3 && nhw_iabs(lf - rt) < t.t3;
					inner = outer && nhw_iabs(ce - lf) < t.t4 - 1 && nhw_iabs(up - ce) < t.t4;
				} else {
					outer = nhw_iabs(B[b + 2] - up) < t.t3 && nhw_iabs(up - B[b]) < t.t3 && nhw_iabs(B[b] - lf) < t.t3 &&
					        nhw_iabs(B[b + 2] - rt) < t.t3;
					inner = outer && nhw_iabs(dn - lf) < t.t3 && nhw_iabs(lf - ce) < t.t4;
				}
				if (inner) {
					const int mean = (up + dn + lf + rt + (pass == 0 ? 2 : 1)) >> 2;
					if (nhw_iabs(mean - lf) < 5 || nhw_iabs(mean - rt) < 5) B[b + BS + 1] = (int16_t)mean;
					cursor = s + YW + 1;
					P_.children(cursor, t.t6, t.t6 + 6, 32);
				}
				// pass 2 silences the siblings together with the children; pass 3 does it one test further out, with
				// whatever the cursor holds
				if (deep && (pass == 0 ? inner : outer))
					for (int k = -1; k <= 1; k++) P_.siblings(cursor + k);
			}
	if (!deep) return;
	// pass 4: three flat samples in a row
	for (int r = 0; r < 128; r++)
		for (int j = 0, s = r * YW, b = r * BS; j < 126; j++, s++, b++) {
			const int v0 = B[b], v1 = B[b + 1], v2 = B[b + 2];
			if (nhw_iabs(v2 - v1) < t.t7 && nhw_iabs(v2 - v0) < t.t7 && nhw_iabs(v1 - v0) < t.t7) {
				P_.children(s + 1, t.t6, t.t6 + 6, 34);
				P_.siblings(s + 1);
			}
		}
}
NHW_HDN void y_e8_smooth_band(int16_t *B, int BS, int16_t *P, int q) { y_e8_smooth_walk(B, BS, E8Direct{P}, q); }
NHW_HDN void y_e8_smooth_image(const EncImg &im, int q) { y_e8_smooth_band(im.proc, YW, im.proc, q); }

// ---- E14 below q16.  q14/q15: pointwise.  q <= 13: thresholds chosen from a global count (q <= 12), then three
// in-place walks that look at the parent sample in the level-2 snapshot (im.ll2s, flat index) and at both
// neighbours, zeroing a cell together with one of them.
// q14 / q15: pointwise, one row (256..511) at a time
NHW_HD void y_e14_q14_row(const EncImg &im, int q, int ratio, int r)
{
	const int hi2 = q == 15 ? 19 : 20;
	int16_t *row = im.proc + r * YW;
	for (int j = 0; j < 256; j++) {
		const int v = nhw_iabs(row[j]);
		if (v >= ratio && v < 11) row[j] = 0;
	}
	for (int j = 256; j < 512; j++) {
		const int v = nhw_iabs(row[j]);
		if (v >= ratio && v < hi2) row[j] = (int16_t)(row[j] >= 14 ? 7 : row[j] <= -14 ? -7 : 0);
	}
}
NHW_HDN void y_e14_lowq_image(const EncImg &im, int q, int ratio)
{
	int16_t *P = im.proc;
	const int16_t *S = im.ll2s;
	if (q >= 14) {
		for (int r = 256; r < 512; r++) y_e14_q14_row(im, q, ratio, r);
		return;
	}
	int t1 = 15, t2 = 27, t3 = 10, t4 = 6, t5 = 3;   // q13
	if (q <= 12) {
		t1 = 16; t2 = 28; t3 = 11; t4 = 8; t5 = 5;
		int busy = 0;
		for (int i = 131072; i < 262144; i++) busy += nhw_iabs(P[i]) >= 12;
		if (busy > 12500) { t1 = 19; t2 = 31; t3 = 13; t4 = 9; t5 = 6; }
		else if (busy > 10000) { t1 = 18; t2 = 30; t3 = 12; t4 = 8; t5 = 6; }
		else if (busy >= 7000) { t1 = 17; t2 = 29; t3 = 11; t4 = 8; t5 = 5; }
		if (q == 11) {
			if (busy > 12500) { t1++; t2++; t3++; t4++; t5++; }
			else t1++;
		} else if (q <= 10) {
			if (busy > 12500) { t1 += 3; t2 += 3; t3 += 2; t4 += 3; t5 += 3; }
			else { t1 += 3; t2 += 2; t3 += 2; t4 += 2; t5 += 2; }
		}
	}
	// a small cell goes when its parent is small, or together with the neighbour it nearly cancels
	auto weak = [&](int s, int top, int parent, int parent_top) {
		const int v = nhw_iabs(P[s]);
		if (v < ratio || v >= top) return;
		if (nhw_iabs(parent) < parent_top) P[s] = 0;
		else if (nhw_iabs(P[s] + P[s - 1]) < t5 && nhw_iabs(P[s + 1]) < t5) { P[s] = 0; P[s - 1] = 0; }
		else if (nhw_iabs(P[s] + P[s + 1]) < t5 && nhw_iabs(P[s - 1]) < t5) { P[s] = 0; P[s + 1] = 0; }
	};
	auto lonely = [&](int s) { return nhw_iabs(P[s - 1]) < ratio && nhw_iabs(P[s + 1]) < ratio; };
	for (int r = 0; r < 256; r++)
		for (int j = 256, s = r * YW + 256; j < 512; j++, s++) {
			weak(s, t3 + 2, S[((r * 256 + (j - 256)) >> 1) + 128], t4);
			const int v = nhw_iabs(P[s]);
			if (v >= ratio && v < t3 && lonely(s)) P[s] = 0;
		}
	for (int r = 256; r < 512; r++) {
		for (int j = 0, s = r * YW; j < 256; j++, s++) {
			weak(s, t1 + 2, S[(((r - 256) * 256 + j) >> 1) + 32768], t4);
			const int v = nhw_iabs(P[s]);
			if (v >= ratio && v < t1 && (lonely(s) || v < t1 - 4)) P[s] = 0;
		}
		for (int j = 256, s = r * YW + 256; j < 511; j++, s++) {
			weak(s, t2 + 1, S[(((r - 256) * 256 + (j - 256)) >> 1) + 32768 + 128], t4 + 1);
			const int v = nhw_iabs(P[s]);
			if (v >= ratio && v < t2 && (lonely(s) || v < t2 - 5)) {
				if (q > 10) P[s] = (int16_t)(P[s] >= 16 ? 7 : P[s] <= -16 ? -7 : 0);
				else P[s] = 0;
			}
		}
	}
}

// ---- offsetY_recons256's isolated-coefficient shrink at q <= 16 (image_processing.c:3137-3160): a diagonal neighbour
// only blocks from 16 up.  A blocking diagonal neighbour of exactly +-16 may itself shrink first, so the rule needs
// the rows above final and the rows below untouched; the four direct neighbours cannot change while this cell is a
// candidate (they would need it to be below 8), so the cells of one row are independent.
NHW_HD void y_recons_shrink_lowq_cell(int16_t *J, int r, int j)
{
	const int e = r * YW + j;
	if (nhw_iabs(J[e]) < 8) return;
	if (nhw_iabs(J[e - YW - 1]) >= 16 || nhw_iabs(J[e - YW]) >= 8 || nhw_iabs(J[e - YW + 1]) >= 16 || nhw_iabs(J[e - 1]) >= 8 ||
	    nhw_iabs(J[e + 1]) >= 8 || nhw_iabs(J[e + YW - 1]) >= 16 || nhw_iabs(J[e + YW]) >= 8 || nhw_iabs(J[e + YW + 1]) >= 16)
		return;
	if (r >= 128 || j >= 128) J[e] += J[e] > 0 ? -1 : 1;
}

// ---- offsetY, the coefficient -> byte loop at q <= 16 (image_processing.c:312-519).  Differences from the q > 16
// form (y_offset_quant_image): no pattern codes exist; negative values are cut on the QuantCycle (restarted per
// row); and pairs of neighbouring large values whose magnitudes both end in 6/7 trade two units on a three-state
// cycle that is never restarted (image-serial).
NHW_HDN void y_offset_quant_lowq_image(const EncImg &im, int m1)
{
	int16_t *P = im.proc;
	QuantCycle cyc;
	cyc.reset();
	int trade = 0;          // the never-reset cycle (quant4)
	for (int i = 0; i < 4 * 65536; i++) {
		const int col = i & 511;
		const bool inrow = col < 511;
		if (col == 0) cyc.reset();
		int a = P[i];
		if (a > 10000) {   // (no stage produces these codes at q <= 16; kept for the reference's order of tests)
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
			a = -cyc.cut(a, 504);
		} else if (a == 8 && P[i + 1] == -7 && inrow) P[i + 1] = -8;
		else if (a > 12 && (a & 7) >= 6) {
			if (inrow && P[i + 1] == 7) P[i + 1] = 9;
		}
		if (a >= 14 && P[i + 1] >= 14 && (i >= 2 * 65536 || col >= 256)) {
			const int nx = P[i + 1];
			if (((a & 510) & 7) == 6 && ((nx & 510) & 7) == 6 && ((a & 1) || (nx & 1))) {
				// a neighbour that is negative with a magnitude ending in 6/7 (or in -3..-7) blocks the trade on its side
				auto blocks = [](int v) { return (v < -2 && v > -8) || (v < -7 && ((-v) & 7) >= 6); };
				bool bl = false, br = false;
				if (col > 0 && col < 510) { bl = blocks(P[i - 1]); br = blocks(P[i + 2]); }
				if (trade == 0) {
					const bool same_step = (a & 504) == (nx & 504);
					const bool first = same_step ? a >= nx : a <= nx;
					if (first) { if (!bl) { a += 2; P[i + 1] = (int16_t)(nx - 2); } }
					else if (!br) P[i + 1] = (int16_t)(nx + 2);
				}
				trade = trade == 2 ? 0 : trade + 1;
			}
		}
		if (a < m1 && a > -m1) { P[i] = 128; continue; }
		P[i] = (int16_t)((a + 128) & 248);
	}
}

// Row form of the loop above.  R = the row before the stage (read only), W = where its bytes go (NULL: dry run),
// next0 = the first cell of the next row before the stage (0 after the last row).  `trade` = the state of the
// never-restarted cycle when the row starts; the state it leaves is returned.  spill != 0: the row's last cell traded
// with the NEXT row's first cell (the reference's flat indexing lets it) -- the caller then has to take the image
// through the serial form.  A row is walked once per possible incoming state (dry) to chain the rows, then once for real.
NHW_HD int y_offset_quant_lowq_row(const int16_t *R, int16_t *W, int r, int m1, int next0, int trade, int &spill)
{
	QuantCycle cyc;
	cyc.reset();
	spill = 0;
	int cur = R[0], prevq = 0;
	for (int col = 0; col < 512; col++) {
		const bool inrow = col < 511;
		int nx = inrow ? (int)R[col + 1] : next0;   // the next cell as it stands; look-ahead rules of this cell rewrite it
		int a = cur, outv = -1;
		if (a > 10000) {
			outv = a == 10100 ? 128 : a == 12700 ? 127 : a == 12900 ? 129 : a == 10204 ? 125 : a == 10300 ? 126 : a == 12100 ? 121 : a == 12200 ? 122 : -1;
		}
		if (outv < 0 && a > 127) { const int k = ((a & 0xfff8) - 128) >> 3; outv = NHW_EXTRA1(k > 18 ? 18 : k); }
		else if (outv < 0 && a < -127) { const int k = (((-a) & 0xfff8) - 128) >> 3; outv = NHW_EXTRA2(k > 18 ? 18 : k); }
		if (outv < 0) {
			if (a < -12 && ((-a) & 7) == 6) { if (inrow && nx == -7) nx = -9; }
			if (a < 0) {
				if (a == -7 && nx == 8 && inrow) a = -8;
				a = -a;
				if (a > 14 && (a & 7) == 7 && nx > 0 && nx < 8) a -= 2;
				a = -cyc.cut(a, 504);
			} else if (a == 8 && nx == -7 && inrow) nx = -8;
			else if (a > 12 && (a & 7) >= 6) { if (inrow && nx == 7) nx = 9; }
			if (a >= 14 && nx >= 14 && (r >= 256 || col >= 256)) {
				if (((a & 510) & 7) == 6 && ((nx & 510) & 7) == 6 && ((a & 1) || (nx & 1))) {
					auto blocks = [](int v) { return (v < -2 && v > -8) || (v < -7 && ((-v) & 7) >= 6); };
					bool bl = false, br = false;
					if (col > 0 && col < 510) { bl = blocks(prevq); br = blocks(R[col + 2]); }
					if (trade == 0) {
						const bool same_step = (a & 504) == (nx & 504);
						const bool first = same_step ? a >= nx : a <= nx;
						int d = 0;
						if (first) { if (!bl) { a += 2; d = -2; } }
						else if (!br) d = 2;
						nx += d;
						if (!inrow) spill = d;
					}
					trade = trade == 2 ? 0 : trade + 1;
				}
			}
			outv = (a < m1 && a > -m1) ? 128 : ((a + 128) & 248);
		}
		if (W) W[col] = (int16_t)outv;
		prevq = outv;
		cur = nx;
	}
	return trade;
}

// ---- chroma pre-filter (pre_processing_UV): 8-neighbour Laplacian of the un-filtered plane nudges the sample
// src = the plane before the stage (u8 4:2:0 bytes); returns the new sample.  Border cells are not visited.
NHW_HD int c_pre_uv_cell(const uint8_t *src, int q, int r, int j)
{
	const int v = src[r * 256 + j];
	if (r < 1 || r > 254 || j < 1 || j > 254) return v;
	const uint8_t *p = src + r * 256 + j;
	const int res = 8 * v - p[-1] - p[1] - p[-256] - p[256] - p[-257] - p[255] - p[-255] - p[257];
	if (q < 14) {
		if (nhw_iabs(res) >= 14) return res > 0 ? v - 2 : v + 2;
		if (nhw_iabs(res) > 5) return res > 0 ? v - 1 : v + 1;
		return v;
	}
	return res > 5 ? v - 1 : res < -5 ? v + 1 : v;
}

// ---- chroma level-1 band thresholds at q <= 16, applied between the two analysis levels: cell (r, j) of the
// transposed coefficient plane, outside the 128x128 level-2 region
NHW_HD int c_threshold_cell(int v, int ratio, int r, int j)
{
	if (r < 128 && j < 128) return v;
	const int top = r < 128 ? 24 : (j < 128 ? 32 : 48);
	const int m = nhw_iabs(v);
	return (m >= ratio && m < top) ? 0 : v;
}

// ---- chroma LL smoothing at q <= 11: the two plus-shaped passes of E8 on the 64x64 band, without descendants
NHW_HDN void c_ll_smooth_image(const EncImg &im)
{
	int16_t *P = im.cproc;
	for (int pass = 0; pass < 2; pass++)
		for (int r = 0; r < 62; r++)
			for (int j = 0, s = r * CW; j < 62; j++, s++) {
				const int up = P[s + 1], dn = P[s + 2 * CW + 1], lf = P[s + CW], rt = P[s + CW + 2], ce = P[s + CW + 1];
				bool go;
				if (pass == 0) go = nhw_iabs(up - dn) < 5 && nhw_iabs(lf - rt) < 5 && nhw_iabs(ce - lf) < 7 && nhw_iabs(up - ce) < 8;
				else
					go = nhw_iabs(P[s + 2] - up) < 5 && nhw_iabs(up - P[s]) < 5 && nhw_iabs(P[s] - lf) < 5 && nhw_iabs(P[s + 2] - rt) < 5 &&
					     nhw_iabs(dn - lf) < 5 && nhw_iabs(lf - ce) < 8;
				if (go) P[s + CW + 1] = (int16_t)((up + dn + lf + rt + (pass == 0 ? 2 : 1)) >> 2);
			}
}
