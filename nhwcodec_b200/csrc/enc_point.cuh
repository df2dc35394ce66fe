// enc_point.cuh -- pointwise forms of the two "quantise to the byte alphabet" loops and of the
// scan permutations that follow them:
//   offsetY loop 4   encoder/image_processing.c:312-519 (q>16)   + scan  encoder/nhw_encoder.c:2111-2132
//   offsetUV         encoder/image_processing.c:108-183          + scan  encoder/nhw_encoder.c:2542-2570
//
// Both loops look sequential (they write the cell ahead of the one being quantised) but the
// look-ahead only ever turns a +-7 into +-8/+-9, and a cell in that state never triggers a
// look-ahead write itself; so the value a cell is quantised from is a function of its own and its
// left neighbour's ORIGINAL values, and the output byte of a cell is a function of the original
// values of (left, self, right).  The chroma loop additionally pairs up runs of {-7,-8} greedily
// from the left, which is the parity of the cell's offset inside its run.  The host harness
// (tests/hostemu) runs these against the reference's taps.
#pragma once
#include "enc_c.cuh"

// om1 / o / op1: original values of the left neighbour, the cell, and the cell that follows it in
// flat order (the next row's first cell from the last column, 0 after the last row).
NHW_HD int y_quant_byte(int om1, int o, int op1, bool has_left /* col >= 1 */, bool inrow /* col < 511 */, int m1)
{
	int a = o;
	if (has_left && om1 >= -127 && om1 <= 127) {   // what the left neighbour's step wrote into this cell
		if (o == -7) {
			if (om1 < -12 && ((-om1) & 7) == 6) a = -9;
			else if (om1 == 8) a = -8;
		} else if (o == 7) {
			if (om1 > 12 && (om1 & 7) >= 6) a = 9;
		}
	}
	if (a > 10000) {
		const int b = a == 10100 ? 128 : a == 12700 ? 127 : a == 12900 ? 129 : a == 10204 ? 125 : a == 10300 ? 126 :
		              a == 12100 ? 121 : a == 12200 ? 122 : -1;
		if (b >= 0) return b;
	}
	if (a > 127) {
		const int k = ((a & 0xfff8) - 128) >> 3;
		return NHW_EXTRA1(k > 18 ? 18 : k);
	}
	if (a < -127) {
		const int k = (((-a) & 0xfff8) - 128) >> 3;
		return NHW_EXTRA2(k > 18 ? 18 : k);
	}
	if (a < 0) {
		if (a == -7 && op1 == 8 && inrow) a = -8;
		a = -a;
		if (a > 14 && (a & 7) == 7 && op1 > 0 && op1 < 8) a -= 2;
		if ((a & 7) < 7) a &= 504;
		a = -a;
	}
	if (a < m1 && a > -m1) return 128;
	return (a + 128) & 248;
}

// position of luma cell (row, col) in the scan: 4-column strips, two rows per step, second row reversed
NHW_HD int y_scan_pos(int row, int col) { return (col >> 2) * 2048 + (row >> 1) * 8 + ((row & 1) ? 7 - (col & 3) : (col & 3)); }

NHW_HD bool c_pairable(int v) { return v == -7 || v == -8; }

// a cell whose step turns a following 7 into 8 (image_processing.c:176-178) -- unless it is a 7 that was
// itself turned into 8, which is why runs of 7s alternate
NHW_HD bool c_bumps_next(int v) { return v > 6 && v <= 127 && (v & 7) >= 6; }

// left_run: number of consecutive cells, immediately left of the cell in its row, whose ORIGINAL value is
//   pairable (if the cell is pairable) / equal to 7 (if the cell is 7); otherwise unused.
// run_pre_bumps: the cell left of that run exists in the row and c_bumps_next() holds for it.
NHW_HD int c_quant_byte(int o, int op1, int left_run, bool run_pre_bumps, bool inrow /* col < 255 */, int m2)
{
	if (c_pairable(o)) {
		if (left_run & 1) return 120;                       // second half of a pair
		if (inrow && c_pairable(op1)) return 120;           // first half
	}
	int a = o;
	if (o == 7 && ((left_run & 1) != 0) != run_pre_bumps) a = 8;   // 7s alternate, phase set by what precedes the run
	if (a > 10000) {
		const int b = a == 12400 ? 124 : a == 12600 ? 126 : a == 12900 ? 122 : a == 13000 ? 130 : -1;
		if (b >= 0) return b;
	}
	if (a > 127) {
		const int k = ((a & 0xfff8) - 128) >> 3;
		return NHW_EXTRA1(k > 18 ? 18 : k);
	}
	if (a < -127) {
		const int k = (((-a) & 0xfff8) - 128) >> 3;
		return NHW_EXTRA2(k > 18 ? 18 : k);
	}
	if (a < 0) {
		a = -a;
		if (op1 < 0 && op1 > -8) { if ((a & 7) < 6) a &= 504; }
		else if ((a & 7) < 7) a &= 504;
		a = -a;
	}
	if (a < m2 && a > -m2) return 128;
	return (a + 128) & 248;
}
