// enc_y1.cuh -- luma encoder stages between the first analysis and the LL2 coder:
// the closed loop that corrects LL1 against what the decoder will reconstruct
// (encoder/nhw_encoder.c:141-283) and the quantise->dequantise model it uses
// (offsetY_recons256, encoder/image_processing.c:2600-3190).
//
// Every function is either *_row (rows are independent: one thread per row) or *_image
// (raster dependencies across rows: one thread per image).  Flat plane indices are kept
// where the reference relies on them (neighbours that wrap to the adjacent row).
// The *_row / *_image forms cover q1..q23; the cell-group forms in enc_cells.cuh are the q17..q23 fast path.
#pragma once
#include "dwt_core.cuh"

#define YW 512   // luma row stride

NHW_HD bool nhw_odd(int v) { return (v & 1) == 1; }
NHW_HD bool in4to7(int v) { return v > 3 && v <= 7; }
NHW_HD bool in_m7to_m4(int v) { return v < -3 && v >= -7; }

// ---- E6a (nhw_encoder.c:144-177): tag LL1 cells from the level-2 detail coefficient below
// them; +16000 / +12000 are consumed by e6c after the trial reconstruction.
NHW_HD void y_e6a_tag_row(const EncImg &im, int r)
{
	const int16_t *P = im.proc;
	for (int j = 0; j < 256; j++) {
		if (r < 128 && j < 128) continue;
		int scan = r * YW + j;
		int st = P[scan];
		int add = 0;
		if (st < -7) {
			int low = (-st) & 7;
			if (low == 7 || low == 0) add = 16000;
		} else if (st < -4) {
			add = 12000;
		} else if (st >= 0) {
			if (st >= 2 && st < 5) {
				if (scan >= YW + 1 && scan < 2 * 65536 - YW - 1) {
					if (P[scan - (YW + 1)] != 0 || P[scan + (YW + 1)] != 0) add = 12000;
				}
			} else if ((st & 7) == 0 || (st & 7) == 1) add = 12000;
			else if (st > 4 && st <= 7) add = 16000;
		}
		if (add) im.ll1[r * 256 + j] = (int16_t)(im.ll1[r * 256 + j] + add);
	}
}

// ---- offsetY_recons256, LL2 part (image_processing.c:2610-2737).  Serial: a row nudges
// cells of the next row before that row is visited.
// P = LL2 band at row stride PS (the plane itself, PS = 512, or a shared-memory copy of its
// first 128 rows x 132 columns); J = im_jpeg (row stride 512); tmp = 16384 cells of scratch.
NHW_HDN void y_recons_ll2_core(int16_t *P, int PS, int16_t *J, int16_t *tmp, const uint16_t *hmem, int hmem_len, int q, int part)
{
	if (q > 17) {
		for (int r = 0; r < 128; r++) {
			int a = r * PS;
			for (int j = 0; j < 125; j++, a++) {
				if (nhw_odd(P[a]) && nhw_odd(P[a + 1]) && nhw_odd(P[a + 2]) && nhw_odd(P[a + 3]) &&
				    nhw_iabs(P[a] - P[a + 3]) > 1) {
					P[a] += 16000;
					P[a + 2] += 16000;
					if (!part) { P[a + 1] += 16000; P[a + 3] += 16000; }
					j += 3;
					a += 3;
				}
			}
		}
	}
	for (int r = 0; r < 128; r++) {
		int a = r * PS, aj = r * YW;
		for (int j = 0; j < 128; j++, a++, aj++) {
			if (P[a] > 10000) {
				if (!part) J[aj] = P[a];
				else {
					P[a] -= 16000;
					J[aj] = P[a];
					J[aj + 1] = (P[a + 1] > 0 && P[a + 1] < 256) ? (int16_t)(P[a + 1] & 65534) : P[a + 1];
					j++;
					a++;
					aj++;
				}
				continue;
			} else if (nhw_odd(P[a]) && j > 0 && nhw_odd(P[a + 1])) {
				if (j < 126 && nhw_odd(P[a + 2])) {
					if (nhw_iabs(P[a] - P[a + 2]) > 1 && q > 17) P[a + 1]++;
				} else if (r < 127 && nhw_odd(P[a + PS]) && nhw_odd(P[a + PS + 1]) && !nhw_odd(P[a + PS + 2])) {
					if (P[a + PS] < 10000 && q > 17) P[a + PS]++;
				}
			} else if (nhw_odd(P[a]) && r >= 1 && r < 125) {
				if (nhw_odd(P[a + PS]) && nhw_odd(P[a + PS + 1])) {
					if (nhw_odd(P[a + 2 * PS]) && !nhw_odd(P[a + 3 * PS])) {
						if (P[a + PS] < 10000 && q > 17) P[a + PS]++;
					}
				}
			}
			if (part) J[aj] = (P[a] > 0 && P[a] < 256) ? (int16_t)(P[a] & 65534) : P[a];
		}
	}
	if (!part) {
		// highres_tmp (image_processing.c:2715-2736)
		int t = 0;
		for (int r = 0; r < 128; r++) {
			int a = r * PS, aj = r * YW;
			for (int j = 0; j < 128; j++, a++, aj++) {
				if (P[a] < 10000) {
					tmp[t++] = P[a];
					J[aj] = (P[a] >= 0 && P[a] < 256) ? (int16_t)(P[a] & 65534) : P[a];
				} else {
					P[a] -= 16000;
					tmp[t++] = P[a];
					J[aj] = P[a];
				}
			}
		}
		if (q > 15) {
			for (int k = 0; k < hmem_len; k++) {
				int m = hmem[k];
				J[((m >> 7) << 9) + (m & 127)] = tmp[m];
			}
		}
	}
}

// ---- offsetY_recons256, 3-in-a-row / vertical-pair substitutions in the level-2 detail
// bands (image_processing.c:2757-2849).  Serial: writes into the next row.
NHW_HD void recons_pattern_cell(int16_t *P, int16_t *J, int &a, int &j)
{
	int v = P[a];
	if (v > 3 && v < 8) {
		if (in4to7(P[a - 1])) {
			if (in4to7(P[a + 1])) {
				P[a - 1] = 15300; P[a] = 0; J[a] = 5; J[a + 1] = 5; j++; a++;
			} else if (in4to7(P[a + YW - 1])) {
				if (in4to7(P[a + YW])) {
					P[a - 1] = 15500; J[a] = 5;
					P[a + YW - 1] = 15500; J[a + YW] = 5;
					P[a + YW] = 0;
					j++; a++;
				}
			}
		}
	} else if (v < -3 && v > -8) {
		if (in_m7to_m4(P[a - 1])) {
			if (in_m7to_m4(P[a + 1])) {
				P[a - 1] = 15400; P[a] = 0; J[a] = -6; J[a + 1] = -5; j++; a++;
			} else if (in_m7to_m4(P[a + YW - 1])) {
				if (in_m7to_m4(P[a + YW])) {
					P[a - 1] = 15600; J[a] = -5;
					P[a + YW - 1] = 15600; J[a + YW] = -5;
					P[a + YW] = 0;
					j++; a++;
				}
			}
		}
	}
}

// ---- offsetY_recons256, second call only: like-signed 5..7 pairs (image_processing.c:2851-2905)
NHW_HD void y_recons_tag57_row(const EncImg &im, int r)
{
	int16_t *P = im.proc;
	int j = r < 128 ? 128 : 0;
	int a = r * YW + j;
	for (; j < 255; j++, a++) {
		int v = P[a], w = P[a + 1];
		if (v >= 5 && v <= 7) {
			if (w >= 5 && w <= 7) { P[a] = 15700; j++; a++; }
		} else if (v <= -5 && v >= -7) {
			if (w <= -5 && w >= -7) { P[a] = 15800; j++; a++; }
		}
	}
}

// The q <= 16 quantisers cut negative values of the form -(8k+7) on a cycle: of the cells of a row that hold -15
// the first of every six is cut to -8 and the other five stay -15; of the cells holding -(8k+7) <= -23 the first of
// every four is cut.  Everything else is cut (image_processing.c:357-410,
// 2938-2990).  The two counters restart at the beginning of every row (of the region being walked).
struct QuantCycle {
	int n15, n23;
	NHW_HD void reset() { n15 = 0; n23 = 0; }
	// a = magnitude of a negative coefficient; returns the magnitude after the q <= 16 rounding rule
	NHW_HD int cut(int a, int mask)
	{
		if (a == 15) {
			const bool first = n15 == 0;
			n15 = n15 == 5 ? 0 : n15 + 1;
			return first ? (a & mask) : a;
		}
		if (a > 22 && (a & 7) == 7) {
			const bool first = n23 == 0;
			n23 = n23 == 3 ? 0 : n23 + 1;
			return first ? (a & mask) : a;
		}
		return a & mask;
	}
};

// ---- offsetY_recons256, dead-zone quantise + dequantise of one detail row into im_jpeg
// (image_processing.c:2909-3133)
NHW_HD void y_recons_quant_row(const EncImg &im, int r, int m1, int part, int q = 20)
{
	int16_t *P = im.proc + r * YW, *J = im.jpeg + r * YW;
	QuantCycle cyc;
	cyc.reset();
	for (int j = r < 128 ? 128 : 0; j < 256; j++) {
		int a = P[j];
		if (a > 15000) {
			if (a == 15300) { J[j] = 5; j += 2; }
			else if (a == 15400) { J[j] = -5; j += 2; }
			else if (a == 15500) { J[j] = 5; j++; }
			else if (a == 15600) { J[j] = -5; j++; }
			else if (a == 15700) { J[j] = 6; J[j + 1] = 6; j++; }
			else if (a == 15800) { J[j] = -6; J[j + 1] = -6; j++; }
			continue;
		}
		if (a < -12 && ((-a) & 7) == 6) {
			if (j < 255 && P[j + 1] == -7) P[j + 1] = -8;
		}
		if (a < 0) {
			if (a == -7 && j < 255 && P[j + 1] == 8) { P[j] = -8; a = -8; }
			a = -a;
			if (q <= 16) a = cyc.cut(a, 65528);
			else if ((a & 7) < 7) a &= 65528;
			a = -a;
		} else if (a == 8 && j < 255 && P[j + 1] == -7) P[j + 1] = -8;
		else if (a > 12 && !part && (a & 7) >= 6) {
			if (j < 255 && P[j + 1] == 7) P[j + 1] = 8;
		}
		if (a < m1 && a > -m1) { J[j] = 0; continue; }
		a += 128;
		if (a < 0) a = -((-a) & 65528);
		else a &= 65528;
		J[j] = (int16_t)(a > 128 ? a - 125 : a - 131);
	}
}

// ---- E6c (nhw_encoder.c:183-216): un-tag LL1 and push the +-1 into the trial reconstruction
NHW_HD void y_e6c_apply_row(const EncImg &im, int r)
{
	int16_t *P = im.proc, *L = im.ll1 + r * 256;
	for (int j = 0; j < 256; j++) {
		int d;
		if (L[j] > 14000) { L[j] -= 16000; d = 1; }
		else if (L[j] > 10000) { L[j] -= 12000; d = -1; }
		else continue;
		if (r < 128 && j >= 128) P[2 * r + ((j - 128) << 10) + YW] += d;
		else if (r >= 128 && j < 128) P[2 * (r - 128) + (j << 10) + 1] += d;
		else if (r >= 128 && j >= 128) P[2 * (r - 128) + ((j - 128) << 10) + YW + 1] += d;
	}
}

// ---- E6d (nhw_encoder.c:218-279): correct LL1 against the trial reconstruction.  Sequential
// along a row (reads the already-corrected left neighbour), rows independent.
NHW_HD int e6d_soften(int a)
{
	if (nhw_iabs(a) <= 4) return a;
	if (a > 0) return a > 11 ? a - 7 : a > 7 ? a - 4 : a > 5 ? a - 2 : a - 1;
	return a < -11 ? a + 7 : a < -7 ? a + 4 : a < -5 ? a + 2 : a + 1;
}

// One step of E6d: the correction of a cell from its own difference `scan`, the right neighbour's (original)
// difference `nxt` and the left neighbour's difference AFTER its correction `prev`.  Only |scan| in 2..4 looks
// at the neighbours at all.
NHW_HD bool e6d_needs_neighbours(int scan) { const int a = nhw_iabs(scan); return a >= 2 && a <= 4; }
NHW_HD int e6d_delta(int scan, int nxt, int prev)
{
	if (scan > 11) return -7;
	if (scan > 7) return -4;
	if (scan > 5) return -2;
	if (scan > 4) return -1;
	if (scan < -11) return 7;
	if (scan < -7) return 4;
	if (scan < -5) return 2;
	if (scan < -4) return 1;
	if (nhw_iabs(scan) <= 1) return 0;
	const int a = e6d_soften(nxt) + prev;
	if (scan >= 4 && a >= 1) return -1;
	if (scan <= -4 && a <= -1) return 1;
	if (scan == 3 && a >= 0) return -1;
	if (scan == -3 && a <= 0) return 1;
	if (nhw_iabs(a) >= 3) {
		if (scan > 0 && a > 0) return -1;
		if (scan < 0 && a < 0) return 1;
		if (a >= 5) return -2;
		if (a <= -5) return 2;
		if (a >= 4) return -1;
		if (a <= -4) return 1;
	}
	return 0;
}

// Correction of cell j from the row's differences sc[-1 .. 256] (sc[j] = P[j] - L[j] before the pass).  The
// left-to-right dependency only runs through unbroken stretches of cells with |difference| in 2..4, so a cell
// finds the start of its stretch and replays it: cells are independent of each other.
NHW_HD int e6d_delta_at(const int16_t *sc, int j)
{
	if (!e6d_needs_neighbours(sc[j])) return e6d_delta(sc[j], 0, 0);
	int start = j;
	while (start > 0 && e6d_needs_neighbours(sc[start - 1])) start--;
	int prev = start > 0 ? sc[start - 1] + e6d_delta(sc[start - 1], 0, 0) : sc[-1];
	int d = 0;
	for (int t = start; t <= j; t++) {
		d = e6d_delta(sc[t], sc[t + 1], prev);
		prev = sc[t] + d;
	}
	return d;
}

// P, J, L point at column 0 of the row; P[-1], P[256], L[-1], L[256] are the flat neighbours the
// reference reads at the row ends.  J may alias L (J[j] is written after the last read of L[j]).
NHW_HD void y_e6d_correct_cells(int16_t *P, int16_t *J, const int16_t *L)
{
	int prev = P[-1] - L[-1];       // left neighbour's difference AFTER its correction
	int cur = P[0] - L[0];
	for (int j = 0; j < 256; j++) {
		const int nxt = P[j + 1] - L[j + 1];
		const int scan = cur;
		const int d = e6d_delta(scan, nxt, prev);
		const int l = L[j];
		J[j] = (int16_t)(l + d);
		P[j] = (int16_t)(P[j] + d);
		prev = scan + d;
		cur = nxt;
	}
}
