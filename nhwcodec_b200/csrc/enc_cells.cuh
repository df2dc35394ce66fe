// enc_cells.cuh -- cell-parallel forms of luma stages the reference writes as in-place row loops.  One call handles
// a GROUP of 8 consecutive cells of a row (one 16-byte load / store); every value is computed from the plane as it
// was BEFORE the stage (the executor loads, synchronises, then stores; rows never straddle two CTAs), so groups are
// independent.  For each stage the comment states why the row loop reduces to this.  tests/hostemu runs the same
// functions against the reference's taps.
#pragma once
#include "enc_point.cuh"
#include "cells8.cuh"

// ---- E6a (nhw_encoder.c:144-177; row form y_e6a_tag_row): reads the level-2 detail cell and two diagonal
// neighbours, adds a tag to the LL1 copy.  Pointwise as written.  Returns whether L changed.
NHW_HD bool y_e6a_tag_cells(const int16_t *P /* plane */, int16_t *Lrow /* ll1 row r */, int r, int g)
{
	const int j0 = g * 8;
	if (r < 128 && j0 < 128) return false;
	int v[8], add[8];
	ld8(P + r * YW + j0, v);
	bool any = false;
	for (int k = 0; k < 8; k++) {
		const int scan = r * YW + j0 + k, st = v[k];
		int a = 0;
		if (st < -7) {
			const int low = (-st) & 7;
			if (low == 7 || low == 0) a = 16000;
		} else if (st < -4) a = 12000;
		else if (st >= 0) {
			if (st >= 2 && st < 5) {
				if (scan >= YW + 1 && scan < 2 * 65536 - YW - 1) {
					if (P[scan - (YW + 1)] != 0 || P[scan + (YW + 1)] != 0) a = 12000;
				}
			} else if ((st & 7) == 0 || (st & 7) == 1) a = 12000;
			else if (st > 4 && st <= 7) a = 16000;
		}
		add[k] = a;
		any |= a != 0;
	}
	if (!any) return false;
	int l[8];
	ld8(Lrow + j0, l);
	for (int k = 0; k < 8; k++) l[k] += add[k];
	st8(Lrow + j0, l);
	return true;
}

// ---- E6c (nhw_encoder.c:183-216; row form y_e6c_apply_row): un-tag LL1, push +-1 into the trial reconstruction
// at the transposed position.  Every target is written by exactly one source cell.
NHW_HD void y_e6c_apply_cells(int16_t *P, int16_t *Lrow, int r, int g)
{
	const int j0 = g * 8;
	int l[8];
	ld8(Lrow + j0, l);
	bool any = false;
	for (int k = 0; k < 8; k++) {
		const int j = j0 + k;
		int d;
		if (l[k] > 14000) { l[k] -= 16000; d = 1; }
		else if (l[k] > 10000) { l[k] -= 12000; d = -1; }
		else continue;
		any = true;
		if (r < 128 && j >= 128) P[2 * r + ((j - 128) << 10) + YW] += d;
		else if (r >= 128 && j < 128) P[2 * (r - 128) + (j << 10) + 1] += d;
		else if (r >= 128 && j >= 128) P[2 * (r - 128) + ((j - 128) << 10) + YW + 1] += d;
	}
	if (any) st8(Lrow + j0, l);
}

// ---- offsetY loop 1 (image_processing.c:194-237; row form y_offset_mult8_row).  A turn at cell i (columns
// c0(r) .. 510) fires when P[i] and P[i+1] are both positive multiples of 8; it decrements P[i] (when P[i] > 15
// and P[i-1] <= 0) or else P[i+1] (when P[i+1] > 15 and P[i+2] <= 0, column < 510).  A decremented cell is still
// positive, a cell <= 0 is never touched, and a turn that decrements its right neighbour needs P[i+2] <= 0 -- which
// makes the neighbour's own turn fail whether or not it was decremented; so every condition has the same outcome on
// the values before the stage.  o[0..7] = the final values of the group; returns false when nothing changed.
NHW_HD bool offset_mult8_fires(int a, int b) { return a > 7 && b > 7 && !(a & 7) && !(b & 7); }
NHW_HD bool y_offset_mult8_cells(const int16_t *P, int r, int g, int *o)
{
	const int c = g * 8, c0 = r < 256 ? 256 : 0;
	if (c < c0) return false;
	int w[12];   // columns c-2 .. c+9
	const int16_t *R = P + r * YW;
	w[0] = R[c - 2]; w[1] = R[c - 1];
	ld8(R + c, w + 2);
	w[10] = R[c + 8]; w[11] = R[c + 9];
	bool any = false;
	for (int k = 0; k < 8; k++) {
		const int col = c + k;
		const int *v = w + 2 + k;   // v[0] = this cell
		int dec = 0;
		// its own turn
		if (col < 511 && offset_mult8_fires(v[0], v[1]) && v[0] > 15 && v[-1] <= 0) dec = 1;
		// the turn of the cell on its left
		else if (col - 1 >= c0 && col - 1 < 510 && offset_mult8_fires(v[-1], v[0]) && v[0] > 15 && v[1] <= 0 &&
		         !(v[-1] > 15 && v[-2] <= 0))
			dec = 1;
		o[k] = v[0] - dec;
		any |= dec != 0;
	}
	return any;
}

// ---- offsetY like-signed 5..7 pairs (image_processing.c:291-311; row form y_offset_pairs57_row), rows 0..255,
// columns 0..255.  The loop pairs cells of the same class greedily from the left and tags the first of each pair;
// the tag is behind the cursor, so a cell is tagged iff its right neighbour is of its class and it sits at an even
// offset in its run of that class (runs start at column 0 of a row at the earliest).
NHW_HD int pairs57_class(int v) { return (v >= 5 && v <= 7) ? 1 : (v <= -5 && v >= -7) ? 2 : 0; }
NHW_HD bool y_offset_pairs57_cells(const int16_t *P, int r, int g, int *o)
{
	const int c = g * 8;
	const int16_t *R = P + r * YW;
	int v[9];
	ld8(R + c, v);
	v[8] = c + 8 < 256 ? (int)R[c + 8] : 0;
	int any_cls = 0;
	for (int k = 0; k < 8; k++) any_cls |= pairs57_class(v[k]);
	if (!any_cls) return false;
	// offset parity of the first cell in its run
	int cls = pairs57_class(v[0]), par = 0;
	if (cls) { int s = c; while (s > 0 && pairs57_class(R[s - 1]) == cls) { s--; par ^= 1; } }
	bool any = false;
	for (int k = 0; k < 8; k++) {
		const int cl = pairs57_class(v[k]);
		if (k) { if (cl && cl == cls) par ^= 1; else par = 0; }
		cls = cl;
		o[k] = v[k];
		if (cl && !par && c + k < 255 && pairs57_class(v[k + 1]) == cl) { o[k] = cl == 1 ? 10300 : 10204; any = true; }
	}
	return any;
}

// ---- E6d (nhw_encoder.c:218-279), 8 cells: like e6d_delta_at, but the stretch before the group is replayed once
// and the group is then walked left to right.  sc = the row's differences (sc[-1] .. sc[256]).
NHW_HD void e6d_delta_cells(const int16_t *sc, int j0, const int *own /* sc[j0 .. j0+7] */, int *d)
{
	int prev;
	if (e6d_needs_neighbours(own[0])) {
		int start = j0;
		while (start > 0 && e6d_needs_neighbours(sc[start - 1])) start--;
		prev = start > 0 ? sc[start - 1] + e6d_delta(sc[start - 1], 0, 0) : sc[-1];
		for (int t = start; t < j0; t++) prev = sc[t] + e6d_delta(sc[t], sc[t + 1], prev);
	} else prev = 0;   // not read
	for (int k = 0; k < 8; k++) {
		const int s = own[k];
		d[k] = e6d_delta(s, k < 7 ? own[k + 1] : (int)sc[j0 + 8], prev);
		prev = s + d[k];
	}
}

// ---- offsetY_recons256: like-signed 5..7 pairs (second call only; row form y_recons_tag57_row) followed by the
// dead-zone quantise + dequantise of a detail row into im_jpeg (row form y_recons_quant_row).
// Pair tags: same greedy pairing as above, the value a cell has after that loop is a function of its run parity.
NHW_HD int recons_tag57_at(const int16_t *R /* row */, int j, int jr0)
{
	const int v = R[j], cl = pairs57_class(v);
	if (!cl || j >= 255 || pairs57_class(R[j + 1]) != cl) return v;
	int par = 0;
	for (int s = j; s > jr0 && pairs57_class(R[s - 1]) == cl; s--) par ^= 1;
	return par ? v : (cl == 1 ? 15700 : 15800);
}
// The quantiser walks a row with a cursor that jumps over 1 or 2 cells after a pattern tag, and a visited cell may
// rewrite the NEXT cell (-7 -> -8, 7 -> 8) before that cell's turn.  Two facts bound the history a cell depends on:
//   * a cell is visited whenever the two cells before it hold no tag (whatever skipped them ends before it);
//   * a rewritten cell only ever passes on one more rewrite (a 7 turned 8 can turn the next -7 into -8; a -8 does
//     nothing), and a cell that rewrites a 7 is > 12, i.e. not rewritten itself.
// So behind four tag-free cells k..k+3 the loop can be re-started at k+2 and is exact from k+4 on; it is replayed
// from there (or from the start of the row) on private state up to the group, then run over the group.  Only
// im_jpeg is written (the in-place edits of the band are dead: both callers rebuild the region afterwards).
// o[k] / bit k of the result = value / "written" for the 8 cells of the group.
#define RQ_NONE 0x40000000
NHW_HD int rq_val(const int16_t *R, int j, int jr0, int part) { return part ? (int)R[j] : recons_tag57_at(R, j, jr0); }
// one turn: a = the cell (with a pending rewrite applied), nx = the next cell or RQ_NONE at the end of the row.
// Returns the cursor step; out / out_next = values for im_jpeg (RQ_NONE: not written), pend = rewrite of the next cell
NHW_HD int rq_turn(int a, int nx, int m1, int part, int &out, int &out_next, int &pend)
{
	out = RQ_NONE; out_next = RQ_NONE; pend = RQ_NONE;
	if (a > 15000) {
		if (a == 15300) { out = 5; return 3; }
		if (a == 15400) { out = -5; return 3; }
		if (a == 15500) { out = 5; return 2; }
		if (a == 15600) { out = -5; return 2; }
		if (a == 15700) { out = 6; out_next = 6; return 2; }
		if (a == 15800) { out = -6; out_next = -6; return 2; }
		return 1;
	}
	if (a < -12 && ((-a) & 7) == 6) { if (nx == -7) pend = -8; }
	if (a < 0) {
		if (a == -7 && nx == 8) a = -8;
		a = -a;
		if ((a & 7) < 7) a &= 65528;
		a = -a;
	} else if (a == 8 && nx == -7) pend = -8;
	else if (a > 12 && !part && (a & 7) >= 6) { if (nx == 7) pend = 8; }
	if (a < m1 && a > -m1) out = 0;
	else {
		a += 128;
		if (a < 0) a = -((-a) & 65528);
		else a &= 65528;
		out = a > 128 ? a - 125 : a - 131;
	}
	return 1;
}
NHW_HD int y_recons_quant_cells(const int16_t *R /* band row, before the stage */, int r, int g, int m1, int part, int *o)
{
	const int c = g * 8, jr0 = r < 128 ? 128 : 0;
	if (c < jr0) return 0;
	// the group's cells (and the two after it) as this call sees them
	int v[10], T[9];
	ld8(R + c, v);
	v[8] = R[c + 8]; v[9] = R[c + 9];
	if (part) { for (int k = 0; k < 9; k++) T[k] = v[k]; }
	else {
		int cls_prev = 0, par = 0;
		const int cl0 = pairs57_class(v[0]);
		if (cl0) for (int s = c; s > jr0 && pairs57_class(R[s - 1]) == cl0; s--) par ^= 1;
		for (int k = 0; k < 9; k++) {
			const int cl = pairs57_class(v[k]);
			if (k) par = (cl && cl == cls_prev) ? par ^ 1 : 0;
			cls_prev = cl;
			T[k] = (cl && c + k < 255 && pairs57_class(v[k + 1]) == cl && !par) ? (cl == 1 ? 15700 : 15800) : v[k];
		}
	}
	// replay up to the group -- unless the two cells before it are plainly inert (|v| <= 4: no tag, no pair tag, no
	// rewrite of their right neighbour, and a 7 -> 8 rewrite cannot reach them): then the cursor lands on the group's
	// first cell with nothing pending
	int j = c, clean = 0;
	const bool inert = c - 2 >= jr0 && nhw_iabs(R[c - 1]) <= 4 && nhw_iabs(R[c - 2]) <= 4;
	while (!inert && j > jr0) {
		j--;
		clean = rq_val(R, j, jr0, part) > 15000 ? 0 : clean + 1;
		if (clean == 4) { j += 2; break; }
	}
	int mask = 0, pend = RQ_NONE, out, out_next;
	while (j < c) {
		const int a = pend != RQ_NONE ? pend : rq_val(R, j, jr0, part);
		const int nx = j + 1 < c ? rq_val(R, j + 1, jr0, part) : T[0];
		const int step = rq_turn(a, nx, m1, part, out, out_next, pend);
		if (j + 1 == c && out_next != RQ_NONE) { o[0] = out_next; mask |= 1; }
		j += step;
	}
	for (int k = 0; k < 8; k++) {
		if (j != c + k) continue;
		const int a = pend != RQ_NONE ? pend : T[k];
		const int nx = c + k < 255 ? T[k + 1] : RQ_NONE;
		j += rq_turn(a, nx, m1, part, out, out_next, pend);
		if (out != RQ_NONE) { o[k] = out; mask |= 1 << k; }
		if (k < 7 && out_next != RQ_NONE) { o[k + 1] = out_next; mask |= 2 << k; }
	}
	return mask;
}

// ---- E14 + E15 (nhw_encoder.c:783-803, 970-1074; row forms y_e14_threshold_row, y_e15_tags_row).
// E14 is pointwise.  E15 visits every cell of a band row in order, reads (left, self, right) and may rewrite all
// three; but a cell whose value lies outside +-4..+-8 neither triggers a rule nor is ever rewritten, so the loop is
// replayed on private state from the nearest such cell on the left (or from the start of the band row), one turn
// past the group (that turn may still rewrite the group's last cell).
NHW_HD int e14_value(int v, int q, int ratio, int r, int j)
{
	if (r >= 256 && q < 20 && q > 15) {
		const int a = nhw_iabs(v);
		if (a >= ratio && (j < 256 ? a < 9 : a <= 14)) return v > 0 ? 7 : -7;
	}
	return v;
}
NHW_HD bool e15_active(int v) { const int a = nhw_iabs(v); return a >= 4 && a <= 8; }
NHW_HD void e15_turn(int &prev, int &cur, int &nx, bool band_a)
{
	const int v = cur;
	if (v > 4 && v < 8) {
		if (in4to7(prev) && in4to7(nx)) { cur = 12700; prev = 10100; nx = 10100; }
	} else if (v < -4 && v > -8) {
		if (in_m7to_m4(prev) && in_m7to_m4(nx)) { cur = 12900; prev = 10100; nx = 10100; }
	} else if (v == 8) {
		if ((prev & 65534) == 6 || (nx & 65534) == 6) cur = 10;
		else if (band_a && nx == 8) { cur = 9; nx = 9; }
	} else if (v == -8) {
		if (((-prev) & 65534) == 6 || ((-nx) & 65534) == 6) cur = -9;
		else if (band_a && nx == -8) { cur = -9; nx = -9; }
	}
}
NHW_HD bool y_e14_e15_cells(const int16_t *R /* row r */, int r, int g, int q, int ratio, int *o)
{
	const int c = g * 8;
	int raw[8], w[11];   // w[i] = current value of column c - 1 + i
	ld8(R + c, raw);
	bool changed = false, any_active = false;
	for (int k = 0; k < 8; k++) {
		w[k + 1] = e14_value(raw[k], q, ratio, r, c + k);
		any_active |= e15_active(w[k + 1]);
	}
	int j0 = 0, j1 = 0;
	bool band_a = false;
	if (r >= 1 && r <= 254) { j0 = 257; j1 = 511; band_a = true; }
	else if (r >= 257 && r <= 510) { j0 = 1; j1 = 255; }
	// a cell that gets rewritten is itself in +-4..+-8
	if (any_active && j1 && c + 7 >= j0 - 1 && c <= j1) {
		w[0] = c > 0 ? e14_value(R[c - 1], q, ratio, r, c - 1) : 0;
		w[9] = e14_value(R[c + 8], q, ratio, r, c + 8);
		w[10] = e14_value(R[c + 9], q, ratio, r, c + 9);
		// Rewrites only ever take cells out of the classes the rules test, so a turn that cannot fire on the values
		// before the stage cannot fire at all: the group can only change through the turns at columns c-1 .. c+8,
		// and those need three like-signed 4..7 cells in a row or a +-8 in the middle.
		{
			const int wm = c > 1 ? e14_value(R[c - 2], q, ratio, r, c - 2) : 0;
			uint32_t pos = 0, neg = 0, mid = 0;   // bit i = column c - 2 + i
			for (int i = 0; i < 12; i++) {
				const int v = i == 0 ? wm : w[i - 1];
				if (in4to7(v)) pos |= 1u << i;
				if (in_m7to_m4(v)) neg |= 1u << i;
				if (v == 8 || v == -8) mid |= 1u << i;
			}
			const uint32_t can = ((pos & (pos << 1) & (pos >> 1)) | (neg & (neg << 1) & (neg >> 1)) | mid) & 0x7feu;   // turns c-1 .. c+8
			if (!can) { for (int k = 0; k < 8; k++) { o[k] = w[k + 1]; changed |= o[k] != raw[k]; } return changed; }
		}
		if (c > j0 && e15_active(w[0])) {   // replay the turns between the last inactive cell and the group
			int s = c - 1;
			while (s > j0 && e15_active(e14_value(R[s - 1], q, ratio, r, s - 1))) s--;
			int prev = e14_value(R[s - 1], q, ratio, r, s - 1), cur = e14_value(R[s], q, ratio, r, s);
			for (int j = s; j < c; j++) {
				int nx = j + 1 < c ? e14_value(R[j + 1], q, ratio, r, j + 1) : w[1];
				e15_turn(prev, cur, nx, band_a);
				prev = cur;
				cur = nx;
			}
			w[0] = prev;
			w[1] = cur;
		}
		for (int k = 0; k <= 8; k++) {
			const int j = c + k;
			if (j >= j0 && j < j1) e15_turn(w[k], w[k + 1], w[k + 2], band_a);
		}
	}
	for (int k = 0; k < 8; k++) { o[k] = w[k + 1]; changed |= o[k] != raw[k]; }
	return changed;
}

// ---- chroma: offsetUV_recons256 (row forms c_recons_ll_row + c_recons_quant_row), one group of a 128-column row of
// the level-2 region.  The band is only read.  The one sequential element (second call: greedy pairing of
// neighbouring -7/-8 cells into -11,-11) is the parity of a cell's offset in its run of such cells.
NHW_HD int c_recons_quant_value(int a, int nx, int m1)
{
	if (a < 0) {
		a = -a;
		if (nx < 0 && nx > -8) { if ((a & 7) < 6) a &= 65528; }
		else if ((a & 7) < 7) a &= 65528;
		a = -a;
	}
	if (a < m1 && a > -m1) return 0;
	a += 128;
	if (a < 0) a = -((-a) & 65528);
	else a &= 65528;
	return a > 128 ? a - 125 : a - 131;
}
NHW_HD void c_recons_cells(const int16_t *R /* cproc row r */, int r, int g, int m1, int comp, int *o)
{
	const int c = g * 8;
	int v[9];
	ld8(R + c, v);
	v[8] = R[c + 8];
	if (r < 64 && c < 64) {   // LL
		for (int k = 0; k < 8; k++) {
			if (comp) o[k] = (r == 0) == ((k & 1) == 1) ? (int)(int16_t)(v[k] & 65534) : v[k];
			else o[k] = (v[k] > 0 && v[k] < 256) ? (v[k] & 65534) : v[k];
		}
		return;
	}
	const int jr0 = r < 64 ? 64 : 0;
	int par = 0;
	bool prev_pairable = false;
	if (!comp && c_pairable(v[0])) {
		for (int s = c; s > jr0 && c_pairable(R[s - 1]); s--) par ^= 1;
		prev_pairable = true;   // only its parity is used below
	}
	for (int k = 0; k < 8; k++) {
		const int j = c + k;
		if (!comp && c_pairable(v[k])) {
			if (k) par = prev_pairable ? par ^ 1 : 0;
			prev_pairable = true;
			if (par || (j < 127 && c_pairable(v[k + 1]))) { o[k] = -11; continue; }
		} else prev_pairable = false;
		o[k] = c_recons_quant_value(v[k], v[k + 1], m1);
	}
}

// ---- chroma LL1 correction (row form c_correct_row): pointwise
NHW_HD void c_correct_cells(const int16_t *P /* cproc row */, const int16_t *L /* cll1 row */, int g, int is_v, int *o)
{
	const int c = g * 8;
	int p[9], l[9];
	ld8(P + c, p); p[8] = P[c + 8];
	ld8(L + c, l); l[8] = L[c + 8];
	for (int k = 0; k < 8; k++) {
		const int scan = p[k] - l[k], nx = p[k + 1] - l[k + 1];
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
		o[k] = l[k] + d;
	}
}

// ---- E20 (enc_y2.cuh: e20_final_cell states the rule), 8 cells of a row: what the cells on the left pass into the
// group is replayed once, then the group is walked left to right.  row / up / dn point at column 0 of the row, the
// row above and the row below as they were before the stage; columns outside [j0, j1] keep their value.
NHW_HD void e20_cells8(const int16_t *up, const int16_t *row, const int16_t *dn, const E20Pass &g, int r, int c, int *o)
{
	int m[11], u[8], d[8];   // m[i] = column c - 1 + i
	ld8(row + c, m + 1);
	ld8(up + c, u);
	ld8(dn + c, d);
	m[0] = row[c - 1]; m[9] = row[c + 8]; m[10] = row[c + 9];
	int give = e20_give_into(row, (int)(dn - row), g, r, c);
	for (int k = 0; k < 8; k++) {
		const int j = c + k;
		int in = m[k + 1];
		if (j < g.j0 || j > g.j1) { o[k] = in; give = 0; continue; }
		if (give == 100) in = -8;
		else in += give;
		if (j == g.j1) { o[k] = in; give = 0; }
		else o[k] = e20_turn_v(g, r, j, m[k], in, m[k + 2], m[k + 3], u[k], d[k], give);
	}
}
// the cell just right of the group (only needed where that is column j1 = 256 of pass 1)
NHW_HD int e20_edge_cell(const int16_t *row, int S, const E20Pass &g, int r) { return e20_final_cell(row, S, g, r, g.j1); }

// ---- chroma residual tags (nhw_encoder.c:2372-2424; row form c_residual_tags_row), q >= 18.  A visited cell whose
// residual and the next one are both medium and like-signed drops a pair tag (12400 / 12600) into the first free of
// its three band cells and the cursor skips the next cell; otherwise a large residual drops 12900 / 13000.  The band
// cells of a sample are its own, so whether the pair rule fires is a function of the data before the stage, and
// "fire, then skip one" picks every other cell of a run of such cells: the parity of the offset in the run.
// One coupling: the residual that follows column 127 is read at column 128, which is the first band cell of the
// row's column 0 -- that cell's turn is replayed where it matters.
NHW_HD bool c_tag_free(const int16_t *P, int scan)
{
	return nhw_iabs(P[scan + 128]) < 8 || nhw_iabs(P[scan + 32768]) < 8 || nhw_iabs(P[scan + 32768 + 128]) < 8;
}
NHW_HD bool c_tag_pair_fires(int d, int n) { return (d > 3 && d < 7 && n > 2 && n < 7) || (d < -3 && d > -7 && n < -2 && n > -8); }
// the tag a visited cell drops when it does not fire the pair rule (0: none)
NHW_HD int c_tag_single(int d, int n, int res_uv)
{
	if (nhw_iabs(d) <= res_uv) return 0;
	if (d > 0) return 12900;
	if (d == -5) return n < 0 ? 13000 : 0;
	return 13000;
}
// residual of column j of row r, j <= 128; column 128 = the first band cell of column 0 after that cell's turn
NHW_HD int c_tag_diff(const int16_t *P, const int16_t *L, int r, int j, int res_uv)
{
	const int scan = r * CW + j, count = r * 128 + j;
	if (j < 128) return P[scan] - L[count];
	int v = P[scan];
	if (v < 12000 && nhw_iabs(v) < 8) {   // still free: does column 0 (always visited) drop a tag here?
		const int d0 = P[r * CW] - L[r * 128], n0 = P[r * CW + 1] - L[r * 128 + 1];
		if (c_tag_pair_fires(d0, n0)) v = d0 > 0 ? 12400 : 12600;
		else { const int t = c_tag_single(d0, n0, res_uv); if (t) v = t; }
	}
	return v - L[count];
}
// tags[k] = what cell 8g+k drops (0: nothing); the plane is only read: the caller drops the tags once every group of
// the row has decided (a dropped tag makes a band cell look taken)
NHW_HD void c_residual_tags_cells(const int16_t *P, const int16_t *L, int q, int r, int g, int *tags)
{
	for (int k = 0; k < 8; k++) tags[k] = 0;
	if (q < 18) return;
	const int res_uv = q > 17 ? 4 : 5, c = g * 8;
	int d[10];
	for (int k = 0; k < 9; k++) d[k] = c + k <= 128 ? c_tag_diff(P, L, r, c + k, res_uv) : 0;
	// is the first cell of the group skipped?  Count the cells right before it that fire in a row.
	bool skip = false;
	{
		int k = c - 1, cnt = 0, dn = d[0];
		while (k >= 0) {
			const int dk = P[r * CW + k] - L[r * 128 + k];
			if (!(c_tag_pair_fires(dk, dn) && c_tag_free(P, r * CW + k))) break;
			cnt++;
			dn = dk;
			k--;
		}
		skip = (cnt & 1) != 0;
	}
	for (int k = 0; k < 8; k++) {
		const int j = c + k, scan = r * CW + j;
		if (skip) { skip = false; continue; }
		int tag;
		if (c_tag_pair_fires(d[k], d[k + 1]) && c_tag_free(P, scan)) { tag = d[k] > 0 ? 12400 : 12600; skip = true; }
		else tag = c_tag_single(d[k], d[k + 1], res_uv);
		tags[k] = tag;
	}
}
