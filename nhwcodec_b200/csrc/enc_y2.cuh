// enc_y2.cuh -- luma encoder stages after the LL2 coder (encoder/nhw_encoder.c:749-2252):
// level-1 thresholds, pattern tags, residual side channels (res1/res3/res5), clean-up,
// quantisation to bytes (offsetY, encoder/image_processing.c:185-521), serpentine scan and
// the byte-stream peephole passes.  q17..q23 (the q>=22 side channels res6 / char_res1 / high_qsetting3 live in enc_hq.cuh).
#pragma once
#include "enc_ll.cuh"

// ---- E14 (nhw_encoder.c:783-803), q16..q19: small level-1 coefficients -> +-7
NHW_HD void y_e14_threshold_row(const EncImg &im, int q, int ratio, int r /* 256..511 */)
{
	if (!(q < 20 && q > 15)) return;
	int16_t *P = im.proc + r * YW;
	for (int j = 0; j < 512; j++) {
		int v = nhw_iabs(P[j]);
		if (v >= ratio && (j < 256 ? v < 9 : v <= 14)) P[j] = (int16_t)(P[j] > 0 ? 7 : -7);
	}
}

// ---- E15 (nhw_encoder.c:970-1074), q>16: mark runs of three small coefficients in two of
// the level-1 detail bands.  Band A = rows 1..254, cols 257..510; band B = rows 257..510,
// cols 1..254.  Only sideways neighbours matter (the reference's "row above" sub-branches sit
// behind conditions an earlier branch of the same chain already claims, so they never run):
// rows are independent.
NHW_HD void y_e15_tags_row(const EncImg &im, int r /* 1..510 */)
{
	int j0, j1;
	bool band_a;
	if (r >= 1 && r <= 254) { j0 = 257; j1 = 511; band_a = true; }
	else if (r >= 257 && r <= 510) { j0 = 1; j1 = 255; band_a = false; }
	else return;
	int16_t *P = im.proc + r * YW;
	for (int j = j0; j < j1; j++) {
		int v = P[j];
		if (v > 4 && v < 8) {
			if (in4to7(P[j - 1]) && in4to7(P[j + 1])) { P[j] = 12700; P[j - 1] = 10100; P[j + 1] = 10100; }
		} else if (v < -4 && v > -8) {
			if (in_m7to_m4(P[j - 1]) && in_m7to_m4(P[j + 1])) { P[j] = 12900; P[j - 1] = 10100; P[j + 1] = 10100; }
		} else if (v == 8) {
			if ((P[j - 1] & 65534) == 6 || (P[j + 1] & 65534) == 6) P[j] = 10;
			else if (band_a && P[j + 1] == 8) { P[j] = 9; P[j + 1] = 9; }
		} else if (v == -8) {
			if (((-P[j - 1]) & 65534) == 6 || ((-P[j + 1]) & 65534) == 6) P[j] = -9;
			else if (band_a && P[j + 1] == -8) { P[j] = -9; P[j + 1] = -9; }
		}
	}
}

// ---- E16 (nhw_encoder.c:1084-1325), q>12: residual LL1 - reconstruction, walked down one
// column; codes 12100..14900 go into res256 (im.ll1), corrections into the two cells below
// and into one cell of the vertical-detail band.  Columns are independent given that
// neighbours' cells are only read (scan+1 / stage-1 belong to other columns: see note).
NHW_HD int res_setting_of(int q) { return q >= 20 ? 3 : q >= 18 ? 4 : q >= 15 ? 6 : 8; }

NHW_HD void e16_band_fix_w1(int16_t *P, int stage, int left)
{
	if (P[stage] == 7) { if (left >= 0 && left < 8) P[stage] += 2; }
	else if (P[stage] == 8) { if (left >= -2 && left < 8) P[stage] += 2; }
}
NHW_HD void e16_band_fix_w2(int16_t *P, int stage, int left)
{
	if (P[stage] < -14) { if (!((-P[stage]) & 7) || ((-P[stage]) & 7) == 7) P[stage]++; }
	else if (P[stage] == 7 || (P[stage] & 65534) == 8) { if (left >= -2) P[stage] += 3; }
}
NHW_HD void e16_band_fix_w3(int16_t *P, int16_t *Lc /* the LL1 cell */, int stage, int q, int left)
{
	if (q >= 21) { *Lc = 14500; return; }
	if (P[stage] < -14) { if (!((-P[stage]) & 7) || ((-P[stage]) & 7) == 7) P[stage]++; }
	else if (P[stage] >= 0 && ((P[stage] + 2) & 65532) == 8) { if (left >= -2) P[stage] = 10; }
	else if (P[stage] > 14 && (P[stage] & 7) == 7) P[stage]++;
}
NHW_HD void e16_band_fix_w5(int16_t *P, int16_t *Lc /* the LL1 cell */, int stage, int res, int q, int left)
{
	*Lc = 14000;
	if (res == -4) {
		if (P[stage] == -7 || P[stage] == -8) { if (left < 2 && left > -8) P[stage] = -9; }
	} else if (res < -6) {
		if (res < -7 && q >= 21) *Lc = 14900;
		else if (P[stage] < -14) { if (!((-P[stage]) & 7) || ((-P[stage]) & 7) == 7) P[stage]++; }
		else if (P[stage] == 7 || P[stage] == 8) { if (left >= -1 && left < 8) P[stage] += 3; }
	}
}

// One column j of the stage.  Pn/Ln are where cells of column j+1 (and the P(j,255) cell read
// at row 0) come from: the reference walks columns left to right, so columns 0..254 must see
// column j+1 as it was BEFORE the stage (a snapshot when columns run concurrently), while
// column 255 wraps to cells that earlier columns have already finalised (live planes).
NHW_HDN void y_e16_residual_col(const EncImg &im, int q, int j, const int16_t *Pn, const int16_t *Ln)
{
	int16_t *P = im.proc, *L = im.ll1;
	const int rs = res_setting_of(q);
	{
		int scan = j, count = j;
		for (int row = 0; row < 255; row++, scan += YW, count += 256) {
			const int stage = (j << 9) + row + 256;
			int res = P[scan] - L[count];
			int a = P[scan + YW] - L[count + 256];
			int b = P[scan + 2 * YW] - L[count + 512];   // two rows down (reads past the band on the last rows)
			enum { NONE, W1, W2, W3, W5 } go = NONE;
			if (res == 2 && a == 2 && b >= 2) {
				if (b < 5 || b > 6) { L[count] = 12400; P[scan + YW] -= 2; P[scan + 2 * YW] -= 2; }
			} else if (((res == 2 && a == 3) || (res == 3 && a == 2)) && b > 1 && b < 6) {
				L[count] = 12400; P[scan + YW] -= 2; P[scan + 2 * YW] -= 2;
			} else if (res == 3 && a == 3) {
				if (b > 0 && b < 6) { L[count] = 12400; P[scan + YW] -= 2; P[scan + 2 * YW] -= 2; }
				else if (q >= 19) { L[count] = 12100; P[scan + YW] = L[count + 256]; }
			} else if (a == -4 && (res == 2 || res == 3) && (b == 2 || b == 3)) {
				if (res == 2 && b == 2) P[scan + YW]++;
				else { L[count] = 12400; P[scan + YW] -= 2; P[scan + 2 * YW] -= 2; }
			} else if (res == 1 && a == 3 && b == 2) {
				if (row > 0 && (P[scan - YW] - L[count - 256]) >= 0) { L[count] = 12400; P[scan + YW] -= 2; P[scan + 2 * YW] -= 2; }
			} else if ((res == 3 || res == 4 || res == 5 || res > 6) && (a == 3 || (a & 65534) == 4)) {
				if (res > 6) { L[count] = 12500; P[scan + YW] = L[count + 256]; }
				else if (q >= 19) { L[count] = 12100; P[scan + YW] = L[count + 256]; }
				else if (q == 18) {
					if (res < 5 && a == 5) L[count + 256] = 14100;
					else if (res >= 5) L[count] = 14100;
					else if (res == 3 && a >= 4) L[count + 256] = 14100;
					P[scan + YW] = L[count + 256];
				}
			} else if ((res == 2 || res == 3) && (a == 2 || a == 3)) {
				if (b == 0 || b == 1) {
					int c1 = Pn[scan + 1] - Ln[count + 1];
					if (c1 == 2 || c1 == 3) {
						int c2 = Pn[scan + YW + 1] - Ln[count + 257];
						if (c2 == 2 || c2 == 3) {
							if ((Pn[scan + 2 * YW + 1] - Ln[count + 513]) > 0) { L[count] = 12400; P[scan + YW] -= 2; P[scan + 2 * YW] -= 2; }
						}
					}
				}
			} else if (a == 4 && (res == -2 || res == -3) && (-b == 2 || -b == 3)) {
				if (res == -2 && -b == 2) P[scan + YW]--;
				else { L[count] = 12300; P[scan + YW] += 2; P[scan + 2 * YW] += 2; }
			} else if ((res == -3 || res == -4 || res == -5 || res < -7) && (a == -3 || a == -4 || a == -5)) {
				if (res < -7) { L[count] = 12600; P[scan + YW] = L[count + 256]; }
				else if (q >= 19) { L[count] = 12200; P[scan + YW] = L[count + 256]; }
				else if (q == 18) {
					if (res > -5 && a == -5) L[count + 256] = 14000;
					else if (res <= -5) L[count] = 14000;
					else if (res == -3 && a <= -4) L[count + 256] = 14000;
					P[scan + YW] = L[count + 256];
				}
			} else if (a == -2 || a == -3) {
				if (res == -2 || res == -3) {
					if (-b > 0) { L[count] = 12300; P[scan + YW] += 2; P[scan + 2 * YW] += 2; }
					else if (res == -3 && q >= 21) L[count] = 14500;
					else if (b == 0) {
						int c1 = Pn[scan + 1] - Ln[count + 1];
						if (c1 == -2 || c1 == -3) {
							int c2 = Pn[scan + YW + 1] - Ln[count + 257];
							if (c2 == -2 || c2 == -3) {
								if ((Pn[scan + 2 * YW + 1] - Ln[count + 513]) < 0) { L[count] = 12300; P[scan + YW] += 2; P[scan + 2 * YW] += 2; }
							}
						}
					} else if (res == -2) go = W2;
					else go = W3;
				} else if (res == -1 && a == -3 && b == -2) {
					if (row > 0 && (P[scan - YW] - L[count - 256]) <= 0) { L[count] = 12300; P[scan + YW] += 2; P[scan + 2 * YW] += 2; }
				} else if (res == -1) {
					if (-b == 3) { L[count] = 12300; P[scan + YW] += 2; P[scan + 2 * YW] += 2; }
					else go = W1;
				} else if (res == -4) {
					if (-b > 1 && -b < 4) { L[count] = 12300; P[scan + YW] += 2; P[scan + 2 * YW] += 2; }
					else go = W5;
				}
			} else if (res == 0 || res == -1) go = W1;
			else if (res == -2) go = W2;
			else if (res == -3) go = W3;
			else if (res < -rs) go = W5;

			if (go != NONE) {
				const int left = row == 0 ? Pn[stage - 1] : P[stage - 1];
				if (go == W1) e16_band_fix_w1(P, stage, left);
				else if (go == W2) e16_band_fix_w2(P, stage, left);
				else if (go == W3) e16_band_fix_w3(P, L + count, stage, q, left);
				else if (go == W5) e16_band_fix_w5(P, L + count, stage, res, q, left);
			}
		}
	}
}

// Which arm of the rule chain of y_e16_residual_col a step takes is a function of the three differences (and of the
// quality through rs); the arms' own sub-conditions stay with the arms.  Evaluated once per (res, a, b) into a table
// (2431 entries: differences beyond the thresholds the chain tests are clamped), a step then costs one look-up
// instead of up to fourteen compound tests.
NHW_HD int e16_arm(int res, int a, int b, int rs)
{
	if (res == 2 && a == 2 && b >= 2) return 0;
	if (((res == 2 && a == 3) || (res == 3 && a == 2)) && b > 1 && b < 6) return 1;
	if (res == 3 && a == 3) return 2;
	if (a == -4 && (res == 2 || res == 3) && (b == 2 || b == 3)) return 3;
	if (res == 1 && a == 3 && b == 2) return 4;
	if ((res == 3 || res == 4 || res == 5 || res > 6) && (a == 3 || (a & 65534) == 4)) return 5;
	if ((res == 2 || res == 3) && (a == 2 || a == 3)) return 6;
	if (a == 4 && (res == -2 || res == -3) && (-b == 2 || -b == 3)) return 7;
	if ((res == -3 || res == -4 || res == -5 || res < -7) && (a == -3 || a == -4 || a == -5)) return 8;
	if (a == -2 || a == -3) return 9;
	if (res == 0 || res == -1) return 10;
	if (res == -2) return 11;
	if (res == -3) return 12;
	if (res < -rs) return 13;
	return 14;
}
#define E16_LUT_SIZE (17 * 13 * 11)
NHW_HD int e16_clamp(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
NHW_HD int e16_lut_index(int res, int a, int b)
{
	return ((e16_clamp(res, -9, 7) + 9) * 13 + (e16_clamp(a, -6, 6) + 6)) * 11 + (e16_clamp(b, -4, 6) + 4);
}
NHW_HD int e16_lut_entry(int idx, int rs)
{
	const int b = idx % 11 - 4, a = (idx / 11) % 13 - 6, res = idx / (11 * 13) - 9;
	return e16_arm(res, a, b, rs);
}

// The same walk with the four rows a step touches held in registers: the next rows are loaded ahead of time and a step no
// longer waits for the previous step's stores to come back from memory (the column walk is a 255-step dependency
// chain, one column per thread).
// LOCKSTEP (CUDA only): all columns of an image advance one row per barrier and column 255 runs E16_LAG rows behind.
// Column 255 is the one column that reads live data of another column -- the LL1 codes of column 0 up to three rows
// further down (its "next" cell in flat order) -- and in the reference it runs last; behind a lag of four rows it
// sees exactly those final values, so it no longer needs a serial pass of its own after the other 255 columns.
#define E16_LAG 4
template <bool LOCKSTEP> NHW_HD void e16_lockstep()
{
#ifdef __CUDA_ARCH__
	if (LOCKSTEP) __syncthreads();
#endif
}
template <bool LOCKSTEP>
NHW_HDN void y_e16_residual_col_t(const EncImg &im, int q, int j, const int16_t *Pn, const int16_t *Ln, const uint8_t *lut, int lag)
{
	int16_t *P = im.proc, *L = im.ll1;
	const int rs = res_setting_of(q);
	{
		int scan = j, count = j;
		// cells of rows row-1 .. row+2 of this column; o* = what memory holds for the ones not stored yet
		int pm1 = 0, lm1 = 0, p0 = P[j], l0 = L[j], p1 = P[j + YW], l1 = L[j + 256], p2 = P[j + 2 * YW], l2 = L[j + 512];
		int ol0 = l0, op1 = p1, ol1 = l1, op2 = p2;
		int np = P[j + 3 * YW], nl = L[j + 768];   // next row to enter the window (reads past the band on the last rows)
		for (int it = 0; it < 255 + E16_LAG; it++) {
			const int row = it - lag;
			if (row >= 0 && row < 255) {
			scan = row * YW + j; count = row * 256 + j;
			const int stage = (j << 9) + row + 256;
			int res = p0 - l0;
			int a = p1 - l1;
			int b = p2 - l2;
			enum { NONE, W1, W2, W3, W5 } go = NONE;
			const int arm = lut ? lut[e16_lut_index(res, a, b)] : e16_arm(res, a, b, rs);
			if (arm == 0) {
				if (b < 5 || b > 6) { l0 = 12400; p1 -= 2; p2 -= 2; }
			} else if (arm == 1) {
				l0 = 12400; p1 -= 2; p2 -= 2;
			} else if (arm == 2) {
				if (b > 0 && b < 6) { l0 = 12400; p1 -= 2; p2 -= 2; }
				else if (q >= 19) { l0 = 12100; p1 = l1; }
			} else if (arm == 3) {
				if (res == 2 && b == 2) p1++;
				else { l0 = 12400; p1 -= 2; p2 -= 2; }
			} else if (arm == 4) {
				if (row > 0 && (pm1 - lm1) >= 0) { l0 = 12400; p1 -= 2; p2 -= 2; }
			} else if (arm == 5) {
				if (res > 6) { l0 = 12500; p1 = l1; }
				else if (q >= 19) { l0 = 12100; p1 = l1; }
				else if (q == 18) {
					if (res < 5 && a == 5) l1 = 14100;
					else if (res >= 5) l0 = 14100;
					else if (res == 3 && a >= 4) l1 = 14100;
					p1 = l1;
				}
			} else if (arm == 6) {
				if (b == 0 || b == 1) {
					int c1 = Pn[scan + 1] - Ln[count + 1];
					if (c1 == 2 || c1 == 3) {
						int c2 = Pn[scan + YW + 1] - Ln[count + 257];
						if (c2 == 2 || c2 == 3) {
							if ((Pn[scan + 2 * YW + 1] - Ln[count + 513]) > 0) { l0 = 12400; p1 -= 2; p2 -= 2; }
						}
					}
				}
			} else if (arm == 7) {
				if (res == -2 && -b == 2) p1--;
				else { l0 = 12300; p1 += 2; p2 += 2; }
			} else if (arm == 8) {
				if (res < -7) { l0 = 12600; p1 = l1; }
				else if (q >= 19) { l0 = 12200; p1 = l1; }
				else if (q == 18) {
					if (res > -5 && a == -5) l1 = 14000;
					else if (res <= -5) l0 = 14000;
					else if (res == -3 && a <= -4) l1 = 14000;
					p1 = l1;
				}
			} else if (arm == 9) {
				if (res == -2 || res == -3) {
					if (-b > 0) { l0 = 12300; p1 += 2; p2 += 2; }
					else if (res == -3 && q >= 21) l0 = 14500;
					else if (b == 0) {
						int c1 = Pn[scan + 1] - Ln[count + 1];
						if (c1 == -2 || c1 == -3) {
							int c2 = Pn[scan + YW + 1] - Ln[count + 257];
							if (c2 == -2 || c2 == -3) {
								if ((Pn[scan + 2 * YW + 1] - Ln[count + 513]) < 0) { l0 = 12300; p1 += 2; p2 += 2; }
							}
						}
					} else if (res == -2) go = W2;
					else go = W3;
				} else if (res == -1 && a == -3 && b == -2) {
					if (row > 0 && (pm1 - lm1) <= 0) { l0 = 12300; p1 += 2; p2 += 2; }
				} else if (res == -1) {
					if (-b == 3) { l0 = 12300; p1 += 2; p2 += 2; }
					else go = W1;
				} else if (res == -4) {
					if (-b > 1 && -b < 4) { l0 = 12300; p1 += 2; p2 += 2; }
					else go = W5;
				}
			} else if (arm == 10) go = W1;
			else if (arm == 11) go = W2;
			else if (arm == 12) go = W3;
			else if (arm == 13) go = W5;

			if (go != NONE) {
				const int left = row == 0 ? Pn[stage - 1] : P[stage - 1];
				if (go == W1) e16_band_fix_w1(P, stage, left);
				else if (go == W2) e16_band_fix_w2(P, stage, left);
				else if (go == W3) { int16_t lc = (int16_t)l0; e16_band_fix_w3(P, &lc, stage, q, left); l0 = lc; }
				else if (go == W5) { int16_t lc = (int16_t)l0; e16_band_fix_w5(P, &lc, stage, res, q, left); l0 = lc; }
			}
			// row `row` of LL1 and row+1 of the plane are final now
			if (l0 != ol0) L[count] = (int16_t)l0;
			if (p1 != op1) P[scan + YW] = (int16_t)p1;
			pm1 = p0; lm1 = l0;
			p0 = p1; l0 = l1; ol0 = ol1;
			p1 = p2; op1 = op2; l1 = l2; ol1 = l2;
			p2 = np; op2 = np; l2 = nl;
			if (row < 253) { np = P[scan + 4 * YW]; nl = L[count + 1024]; }   // up to row 256
			}
			e16_lockstep<LOCKSTEP>();
		}
		scan = 255 * YW + j; count = 255 * 256 + j;
		// the window still holds row 255 of LL1 (a q18 rule of the last step may have coded it) and row 256 of the plane
		if (l0 != ol0) L[count] = (int16_t)l0;
		if (p1 != op1) P[scan + YW] = (int16_t)p1;
	}
}

NHW_HDN void y_e16b_classify_col_w(const EncImg &im, int q, int j, int &w1, int &w3, int &w5)
{
	int16_t *P = im.proc, *L = im.ll1;
	const int rs = res_setting_of(q);
	{
		// the band cell of the previous row ("left"), this row's band cell, LL1 code and reconstruction sample are
		// carried in registers and the next row's are loaded a step ahead: no step waits for its predecessor's stores
		int left = P[(j << 9) + 255];
		int npv = P[(j << 9) + 256], nl = L[j], np = P[j];
		for (int row = 0; row < 256; row++) {
			const int scan = row * YW + j, count = row * 256 + j;
			const int stage = (j << 9) + row + 256;
			int pv = npv, lc = nl;
			const int p = np, opv = pv, olc = lc;
			if (row < 255) { npv = P[stage + 1]; nl = L[count + 256]; np = P[scan + YW]; }
			if (lc < 12000) {
				int res = p - lc;
				lc = 0;
				if (res == 0 || res == 1) {
					if (pv == -7 || pv == -8) { if (left < 2 && left > -8) pv = -9; }
				} else if (res == 2) {
					if (pv > 15 && !(pv & 7)) pv--;
					else if (pv == -7 || pv == -8) { if (left <= 1) pv = -9; }
					else if (pv == -6) { if (left <= -1 && left > -8) pv = -9; }
				} else if (res == 3) {
					if (q >= 21) { lc = 144; w5++; }
					else if (pv > 15 && !(pv & 7)) pv--;
					else if (pv <= 0 && (((-pv) + 2) & 65532) == 8) { if (left <= 2) pv = -10; }
				} else if (res > rs) {
					lc = 141; w1++;
					if (res == 4) {
						if (pv == 7 || (pv & 65534) == 8) { if (left >= 0 && left < 8) pv += 2; }
					} else if (res > 6) {
						if (res > 7 && q >= 21) { lc = 148; w5++; w1++; }
						else if (pv > 15 && !(pv & 7)) pv--;
						else if (pv == -6 || pv == -7 || pv == -8) { if (left < 0 && left > -8) pv = -9; }
					}
				}
			} else {
				// every code left by the column pass is a multiple of 100; the byte code is code/100
				const int v = lc;
				const bool w1c = v == 14000 || v == 14100, w3c = v == 12100 || v == 12200 || v == 12300 || v == 12400;
				const bool w5c = v == 14500, w31 = v == 12500 || v == 12600, w51 = v == 14900;
				if (w1c || w3c || w5c || w31 || w51) {
					lc = v / 100;
					w1 += (w1c || w31 || w51) ? 1 : 0;
					w3 += (w3c || w31) ? 1 : 0;
					w5 += (w5c || w51) ? 1 : 0;
				}
			}
			if (pv != opv) P[stage] = (int16_t)pv;
			if (lc != olc) L[count] = (int16_t)lc;
			left = pv;
		}
	}
}

// ---- E18 (nhw_encoder.c:1498-1887): turn the codes left in res256 into the res1/res3/res5
// side channels: column positions per row (254 = end of row), pair-delta packed, with the
// positions' LSBs and the 1- or 2-bit words in separate bit planes.
// which: 1, 3 or 5.  Scratch: tmp1 (positions), tmp2 (copy), tmp3 (words).
// Step 1, one row: collect the column positions (and word bits) of the codes that belong to list
// `which`, rewrite those cells, end the row with the 254 marker.  With pos == NULL only counts.
// Returns the number of codes found (the row contributes that many + 1 positions).
NHW_HD int y_e18_collect_row(const EncImg &im, int which, int row, uint8_t *pos, uint8_t *wrd)
{
	int16_t *L = im.ll1 + row * 256;
	int n = 0;
	for (int j = 0; j < 254; j++) {
		const int v = L[j];
		if (v == 0) continue;
		int nv = -1, w = 0;
		if (which == 1) {
			if (v == 141) { nv = 0; w = 1; } else if (v == 140) { nv = 0; w = 0; }
			else if (v == 126) { nv = 122; w = 0; } else if (v == 125) { nv = 121; w = 1; }
			else if (v == 148) { nv = 144; w = 1; } else if (v == 149) { nv = 145; w = 0; }
		} else if (which == 3) {
			if (v == 121) { nv = 0; w = 1; } else if (v == 122) { nv = 0; w = 0; }
			else if (v == 123) { nv = 0; w = 2; } else if (v == 124) { nv = 0; w = 3; }
		} else {
			if (v == 144) { nv = 0; w = 1; } else if (v == 145) { nv = 0; w = 0; }
		}
		if (nv < 0) continue;
		if (pos) { pos[n] = (uint8_t)j; wrd[n] = (uint8_t)w; L[j] = (int16_t)nv; }
		n++;
	}
	if (pos) { pos[n] = 254; L[254] = 0; L[255] = 0; }
	return n;
}

NHW_HDN void y_e18_finish_list_image(const EncImg &im, int which, int count, int e);

// The three collection passes as one function of the cell's value before the stage: list 1 rewrites some of
// its codes into codes of lists 3 and 5, which then collect them (q >= 19 / q >= 21).  member bit k = the cell is
// an entry of list 1 / 3 / 5 (k = 0 / 1 / 2), w[k] its word value; returns the cell's value after all passes.
NHW_HD int y_e18_classify(int v, int q, int &member, int (&w)[3])
{
	member = 0;
	w[0] = w[1] = w[2] = 0;
	if (v == 141) { member |= 1; w[0] = 1; v = 0; }
	else if (v == 140) { member |= 1; w[0] = 0; v = 0; }
	else if (v == 126) { member |= 1; w[0] = 0; v = 122; }
	else if (v == 125) { member |= 1; w[0] = 1; v = 121; }
	else if (v == 148) { member |= 1; w[0] = 1; v = 144; }
	else if (v == 149) { member |= 1; w[0] = 0; v = 145; }
	if (q >= 19 && v >= 121 && v <= 124) { member |= 2; w[1] = v == 121 ? 1 : v == 122 ? 0 : v == 123 ? 2 : 3; v = 0; }
	if (q >= 21 && (v == 144 || v == 145)) { member |= 4; w[2] = v == 144 ? 1 : 0; v = 0; }
	return v;
}

// Steps 2..6 on the collected lists: tmp1 = `count` positions, tmp3 = `e` word values.
NHW_HDN void y_e18_finish_list_image(const EncImg &im, int which, int count, int e)
{
	EncHdr *h = im.hdr;
	uint8_t *pos = im.tmp1, *cpy = im.tmp2, *wrd = im.tmp3;
	uint8_t *out, *out_bit, *out_word;
	if (which == 1) { out = im.res1; out_bit = im.res1_bit; out_word = im.res1_word; }
	else if (which == 3) { out = im.res3; out_bit = im.res3_bit; out_word = im.res3_word; }
	else if (which == 5) { out = im.res5; out_bit = im.res5_bit; out_word = im.res5_word; }
	else { out = im.res6; out_bit = im.res6_bit; out_word = im.res6_word; }
	for (int i = 0; i < 8; i++) wrd[e + i] = 0;   // the reference reads up to 7 entries past the end
	// drop end-of-row markers the decoder can infer from a decreasing position
	for (int i = 0; i < count; i++) cpy[i] = pos[i];
	int len = 1;
	for (int i = 1; i < count - 1; i++) {
		if (cpy[i] == 254 && cpy[i - 1] != 254 && cpy[i + 1] != 254) {
			if (cpy[i - 1] <= cpy[i + 1]) pos[len++] = cpy[i];
		} else pos[len++] = cpy[i];
	}
	pos[len++] = cpy[count - 1];
	// LSB plane of the real positions (cpy reused as the compacted list)
	int np = 0;
	for (int i = 0; i < len; i++)
		if (pos[i] != 254) cpy[np++] = pos[i];
	for (int i = 0; i < 8; i++) cpy[np + i] = 0;
	const int bit_len = (np >> 3) + 1;
	for (int i = 0, o = 0; i < ((np >> 3) << 3) + 8; i += 8) {
		int b = 0;
		for (int k = 0; k < 8; k++) b |= (cpy[i + k] & 1) << (7 - k);
		out_bit[o++] = (uint8_t)b;
	}
	// halve the positions and merge (small delta, small delta) pairs into one byte >= 128
	int olen = 1;
	out[0] = pos[0] >> 1;
	for (int i = 1; i < len - 1; i++) {
		int cur = pos[i] >> 1, d1 = cur - (pos[i - 1] >> 1);
		if (d1 >= 0 && d1 < 8) {
			int d2 = (pos[i + 1] >> 1) - cur;
			if (d2 >= 0 && d2 < 16) { out[olen++] = (uint8_t)(128 + (d1 << 4) + d2); i++; }
			else out[olen++] = (uint8_t)cur;
		} else out[olen++] = (uint8_t)cur;
	}
	// word plane
	int wbytes = 0;
	for (int i = 0; i < ((e >> 3) << 3) + 8; i += 8) {
		if (which == 3) {
			out_word[wbytes++] = (uint8_t)(((wrd[i] & 3) << 6) | ((wrd[i + 1] & 3) << 4) | ((wrd[i + 2] & 3) << 2) | (wrd[i + 3] & 3));
			out_word[wbytes++] = (uint8_t)(((wrd[i + 4] & 3) << 6) | ((wrd[i + 5] & 3) << 4) | ((wrd[i + 6] & 3) << 2) | (wrd[i + 7] & 3));
		} else {
			int b = 0;
			for (int k = 0; k < 8; k++) b |= (wrd[i + k] & 1) << (7 - k);
			out_word[wbytes++] = (uint8_t)b;
		}
	}
	if (which == 1) { h->res1_len = olen; h->res1_bit_len = bit_len; h->res1_word_len = wbytes; }
	else if (which == 3) { h->res3_len = olen; h->res3_bit_len = bit_len; h->res3_word_len = wbytes; }
	else if (which == 5) { h->res5_len = olen; h->res5_bit_len = bit_len; h->res5_word_len = wbytes; }
	else { h->res6_len = olen; h->res6_bit_len = bit_len; h->res6_word_len = wbytes; }
}

// ---- E19 (nhw_encoder.c:1893-1910): restore the level-2 region from the resIII snapshot,
// zeroing LL2 except tagged cells
NHW_HD void y_e19_restore_row(const EncImg &im, int r /* 0..255 */)
{
	int16_t *P = im.proc + r * YW;
	const int16_t *S = im.ll2s + r * 256;
	for (int j = 0; j < 256; j++) P[j] = (j < 128 && r < 128 && S[j] <= 8000) ? (int16_t)0 : S[j];
}

// ---- E20 (nhw_encoder.c:1914-2098): clean-up of the three level-1 detail bands.
// pass 0: rows 1..254, cols 257..510   pass 1: rows 256..510, cols 1..255
// pass 2: rows 256..510, cols 257..510.  In place, reads the row above after it was edited.
NHW_HD void e20_cell(int16_t *P, int a, int j, int jmax, int lo, int yw, int yw2, int pass)
{
	if (nhw_iabs(P[a]) >= lo) {
		if (nhw_iabs(P[a]) < yw2) {
			int cnt = 0;
			if (nhw_iabs(P[a - 1]) + 2 >= 8) cnt++;
			if (nhw_iabs(P[a + 1]) + 2 >= 8) cnt++;
			if (nhw_iabs(P[a - YW]) + 2 >= 8) cnt++;
			if (nhw_iabs(P[a + YW]) + 2 >= 8) cnt++;
			if (cnt < 3 && P[a] < yw && P[a] > -yw) {
				if (pass == 0) { if (P[a] < -6) P[a] = -7; else if (P[a] > 6) P[a] = 7; }
				else P[a] = (int16_t)(P[a] < 0 ? -7 : 7);
			} else if (pass == 1 && cnt == 0 && nhw_iabs(P[a]) < yw2) P[a] = (int16_t)(P[a] < 0 ? -7 : 7);
		}
	} else P[a] = 0;
	if (nhw_iabs(P[a]) > 6) {
		int e = P[a];
		if (e >= 8 && (e & 7) < 2) {
			if (P[a + 1] > 7 && P[a + 1] < 10000) P[a + 1]--;
		} else if (e == -7 && P[a + 1] == 8) P[a] = -8;
		else if (e == 8 && P[a + 1] == -7) P[a + 1] = -8;
		else if (e < -7 && ((-e) & 7) < 2) {
			if (P[a + 1] < -14 && P[a + 1] < 10000) {
				if (((-P[a + 1]) & 7) == 7) P[a + 1]++;
				else if (((-P[a + 1]) & 7) < 2 && j < jmax && P[a + 2] <= 0) P[a + 1]++;
			}
		}
	}
}

// ---- E20 without the wavefront.  The neighbour test of the rule above (|n| + 2 >= 8) has the same outcome on a
// neighbour before and after that neighbour's own turn, except that an already visited neighbour (left, above)
// with |n| < lo has been zeroed: "visited and |n| >= lo, or not visited and |n| >= 6".  So the count is a function
// of the plane as it was before the pass; what remains sequential is only the one-cell look-ahead nudge along a
// row (P[a+1]--, P[a+1]++, P[a+1] = -8), and a cell can only receive one if its value is > 7, == -7 or < -14.
// The final value of a cell is found by replaying its row from the start of its unbroken stretch of such cells.
// P = the plane BEFORE the pass (read only); regions of the three passes never read each other's cells.
struct E20Pass { int r0, r1, j0, j1, jmax, lo, yw, yw2, pass; };   // rows [r0, r1), columns [j0, j1)
NHW_HD E20Pass e20_pass(int q, int ratio, int pass)
{
	E20Pass g;
	g.pass = pass;
	if (pass == 0) { g.r0 = 1; g.r1 = 255; g.j0 = 257; g.j1 = 511; g.jmax = 510; g.lo = ratio - 2; if (q > 22) { g.yw = 8; g.yw2 = 4; } else { g.yw = 9; g.yw2 = 9; } }
	else if (pass == 1) { g.r0 = 256; g.r1 = 511; g.j0 = 1; g.j1 = 256; g.jmax = 254; g.lo = ratio - 2; if (q > 22) { g.yw = 8; g.yw2 = 4; } else if (q > 17) { g.yw = 8; g.yw2 = 9; } else { g.yw = 9; g.yw2 = 9; } }
	else { g.r0 = 256; g.r1 = 511; g.j0 = 257; g.j1 = 511; g.jmax = 510; g.lo = ratio - 1; g.yw = q > 22 ? 8 : 11; g.yw2 = g.yw; }
	return g;
}
NHW_HD bool e20_can_receive(int v) { return v > 7 || v == -7 || v < -14; }
// one cell's turn: v = its value when its turn comes (original + what the left neighbour did to it); returns its
// final value and, in `give`, what it does to the cell on its right (0 none, -1, +1, or 100 = "becomes -8")
// the values version: left / n1 / n2 = the cells at j-1, j+1, j+2, up / dn = above and below, all as before the stage
NHW_HD int e20_turn_v(const E20Pass &g, int r, int j, int left, int v, int n1, int n2, int up, int dn, int &give)
{
	if (nhw_iabs(v) >= g.lo) {
		if (nhw_iabs(v) < g.yw2) {
			int cnt = 0;
			if (nhw_iabs(left) >= (j > g.j0 ? g.lo : 6)) cnt++;     // visited neighbours survive only from lo up
			if (nhw_iabs(n1) >= 6) cnt++;
			if (nhw_iabs(up) >= (r > g.r0 ? g.lo : 6)) cnt++;
			if (nhw_iabs(dn) >= 6) cnt++;
			if (cnt < 3 && v < g.yw && v > -g.yw) {
				if (g.pass == 0) { if (v < -6) v = -7; else if (v > 6) v = 7; }
				else v = v < 0 ? -7 : 7;
			} else if (g.pass == 1 && cnt == 0 && nhw_iabs(v) < g.yw2) v = v < 0 ? -7 : 7;
		}
	} else v = 0;
	give = 0;
	if (nhw_iabs(v) > 6) {
		const int e = v;
		if (e >= 8 && (e & 7) < 2) {
			if (n1 > 7 && n1 < 10000) give = -1;
		} else if (e == -7 && n1 == 8) v = -8;
		else if (e == 8 && n1 == -7) give = 100;
		else if (e < -7 && ((-e) & 7) < 2) {
			if (n1 < -14 && n1 < 10000) {
				if (((-n1) & 7) == 7) give = 1;
				else if (((-n1) & 7) < 2 && j < g.jmax && n2 <= 0) give = 1;
			}
		}
	}
	return v;
}
// `row` points at column 0 of row r (of the plane, stride S = YW, or of a staged copy of it with its own stride)
NHW_HD int e20_turn(const int16_t *row, int S, const E20Pass &g, int r, int j, int v, int &give)
{
	return e20_turn_v(g, r, j, row[j - 1], v, row[j + 1], row[j + 2], row[j - S], row[j + S], give);
}
// what the cells on the left of column c pass into it (0 none, -1, +1, 100 = "becomes -8")
NHW_HD int e20_give_into(const int16_t *row, int S, const E20Pass &g, int r, int c)
{
	if (c <= g.j0 || !e20_can_receive(row[c])) return 0;
	int start = c - 1;
	while (start > g.j0 && e20_can_receive(row[start])) start--;
	int give = 0;
	for (int x = start; x < c; x++) {
		int in = row[x];
		if (give == 100) in = -8;
		else in += give;
		e20_turn(row, S, g, r, x, in, give);
	}
	return give;
}
// j in [j0, j1]: column j1 is outside the region (it has no turn) but still receives from the last cell of the row
NHW_HD int e20_final_cell(const int16_t *row, int S, const E20Pass &g, int r, int j)
{
	int start = j;
	while (start > g.j0 && e20_can_receive(row[start])) start--;
	// `start` receives nothing (it cannot, or it is the first cell of the row): replay start .. j
	int give = 0, v = 0;
	for (int x = start; x <= j; x++) {
		int in = row[x];
		if (give == 100) in = -8;
		else in += give;
		if (x == g.j1) return in;
		v = e20_turn(row, S, g, r, x, in, give);
	}
	return v;
}
