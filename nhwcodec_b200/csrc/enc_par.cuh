// enc_par.cuh -- parallel forms of the raster-order luma stages.
//
// The reference walks planes in raster order and edits them in place, so a cell can depend on
// cells edited earlier.  For each stage the dependency footprint is small and fixed; this file
// states it and exposes the stage as
//   * a WAVEFRONT cell: thread = row, step t handles column (t - skew*row); a row may run ahead
//     of the row below by exactly `skew` columns, which is the smallest lag that lets it read
//     the row below un-edited and the row above fully edited;
//   * a ROW function where rows only couple through one boundary cell, handled explicitly;
//   * a COLUMN function (residual coding walks columns).
// The same functions are run in the same schedule by tests/hostemu (sequentially, rows of a
// step in reverse order) to check the analysis against the reference taps.
#pragma once
#include "enc_hq.cuh"

// ---- wavefront geometry of each stage -------------------------------------------------------
struct WfGeom { int r0, rows, c0, cols, skew; };

NHW_HD int wf_shrink_cell(const EncImg &im, int r, int j)
{
	int16_t *J = im.jpeg;
	const int e = r * YW + j;
	if (nhw_iabs(J[e]) < 8) return 1;
	if (nhw_iabs(J[e - YW - 1]) >= 8 || nhw_iabs(J[e - YW]) >= 8 || nhw_iabs(J[e - YW + 1]) >= 8 ||
	    nhw_iabs(J[e - 1]) >= 8 || nhw_iabs(J[e + 1]) >= 8 || nhw_iabs(J[e + YW - 1]) >= 8 ||
	    nhw_iabs(J[e + YW]) >= 8 || nhw_iabs(J[e + YW + 1]) >= 8)
		return 1;
	if (r >= 128 || j >= 128) J[e] += J[e] > 0 ? -1 : 1;
	return 1;
}

NHW_HD int wf_offset_patterns_cell(const EncImg &im, int r, int j)
{
	int16_t *P = im.proc;
	const int a = r * YW + j;
	const int v = P[a];
	if (v > 3 && v < 8) {
		if (in4to7(P[a - 1])) {
			if (in4to7(P[a + 1])) { P[a] = 12700; P[a - 1] = 10100; return 2; }
			if (in4to7(P[a + YW - 1]) && in4to7(P[a + YW])) {
				P[a - 1] = 12100; P[a] = 10100; P[a + YW - 1] = 10100; P[a + YW] = 10100; return 2;
			}
		}
	} else if (v < -3 && v > -8) {
		if (in_m7to_m4(P[a - 1])) {
			if (in_m7to_m4(P[a + 1])) { P[a] = 12900; P[a - 1] = 10100; return 2; }
			if (in_m7to_m4(P[a + YW - 1]) && in_m7to_m4(P[a + YW])) {
				P[a - 1] = 12200; P[a] = 10100; P[a + YW - 1] = 10100; P[a + YW] = 10100; return 2;
			}
		}
	}
	return 1;
}

// ---- row forms ---------------------------------------------------------------------------------
// offsetY like-signed 5..7 pairs (image_processing.c:291-311): sideways only.
NHW_HD void y_offset_pairs57_row(const EncImg &im, int r /* 0..255 */)
{
	int16_t *P = im.proc + r * YW;
	for (int j = 0; j < 255; j++) {
		int v = P[j], w = P[j + 1];
		if (v >= 5 && v <= 7) { if (w >= 5 && w <= 7) { P[j] = 10300; j++; } }
		else if (v <= -5 && v >= -7) { if (w <= -5 && w >= -7) { P[j] = 10204; j++; } }
	}
}

// offsetY loop 1 (image_processing.c:194-237).  The only cross-row read is P[i-1] at column 0
// (the previous row's last cell) tested as <= 0; this pass only ever decrements cells > 15, so
// that test has the same outcome before and after the previous row ran: rows are independent.
NHW_HD void y_offset_mult8_row(const EncImg &im, int r /* 0..511 */)
{
	int16_t *P = im.proc;
	for (int col = r < 256 ? 256 : 0; col < 511; col++) {
		const int i = r * YW + col;
		if (!(P[i] > 7 && P[i + 1] > 7)) continue;
		int a = P[i];
		if ((a & 7) || (P[i + 1] & 7)) continue;
		if (a > 15) {
			if (i > 0) {
				if (P[i - 1] <= 0) P[i]--;
				else if (P[i + 1] > 15) { if (col < 510 && P[i + 2] <= 0) P[i + 1]--; }
			}
		} else if (P[i + 1] > 15) {
			if (col < 510 && P[i + 2] <= 0) P[i + 1]--;
		}
	}
}

// ---- residual coding with concurrent columns ---------------------------------------------------
// snapshot layout inside the scratch plane: rows 0..257 of `proc` at the same flat indices,
// then the 65536 LL1 cells followed by 1024 zeros (reads past the end must see 0).
#define E16_SNAP_P_CELLS (258 * 512)
#define E16_SNAP_L_OFF (260 * 512)
#define E16_SNAP_L_CELLS (65536 + 1024)

NHW_HD void y_recons_ll2_tag_row(int16_t *P, int PS, int r, int part)
{
	int a = r * PS;
	for (int j = 0; j < 125; j++, a++) {
		if (nhw_odd(P[a]) && nhw_odd(P[a + 1]) && nhw_odd(P[a + 2]) && nhw_odd(P[a + 3]) && nhw_iabs(P[a] - P[a + 3]) > 1) {
			P[a] += 16000;
			P[a + 2] += 16000;
			if (!part) { P[a + 1] += 16000; P[a + 3] += 16000; }
			j += 3;
			a += 3;
		}
	}
}

// parity nudges shared by the recons pass and the LL2 -> bytes pass (same code in the reference)
NHW_HD void ll2_parity_nudge(int16_t *P, int PS, int a, int r, int j, int v, int q)
{
	if (nhw_odd(v) && j > 0 && nhw_odd(P[a + 1])) {
		if (j < 126 && nhw_odd(P[a + 2])) {
			if (nhw_iabs(v - P[a + 2]) > 1 && q > 17) P[a + 1]++;
		} else if (r < 127 && nhw_odd(P[a + PS]) && nhw_odd(P[a + PS + 1]) && !nhw_odd(P[a + PS + 2])) {
			if (P[a + PS] < 10000 && q > 17) P[a + PS]++;
		}
	} else if (nhw_odd(v) && r >= 1 && r < 125) {
		if (nhw_odd(P[a + PS]) && nhw_odd(P[a + PS + 1])) {
			if (nhw_odd(P[a + 2 * PS]) && !nhw_odd(P[a + 3 * PS])) {
				if (P[a + PS] < 10000 && q > 17) P[a + PS]++;
			}
		}
	}
}

NHW_HD int y_recons_ll2_cell(int16_t *P, int PS, int16_t *J, int q, int part, int r, int j)
{
	const int a = r * PS + j, aj = r * YW + j;
	if (P[a] > 10000) {
		if (!part) { J[aj] = P[a]; return 1; }
		P[a] -= 16000;
		J[aj] = P[a];
		J[aj + 1] = (P[a + 1] > 0 && P[a + 1] < 256) ? (int16_t)(P[a + 1] & 65534) : P[a + 1];
		return 2;
	}
	ll2_parity_nudge(P, PS, a, r, j, P[a], q);
	if (part) J[aj] = (P[a] > 0 && P[a] < 256) ? (int16_t)(P[a] & 65534) : P[a];
	return 1;
}
