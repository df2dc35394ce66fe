// enc_ll2_masks.cuh -- the parity nudges of the LL2 band as bit-mask algebra.
//   offsetY_recons256, LL2 part   encoder/image_processing.c:2610-2737   (cell form: y_recons_ll2_cell, enc_par.cuh)
//   LL2 -> bytes                  encoder/nhw_encoder.c:661-743          (cell form: ll2_bytes_cell, enc_ll_par.cuh)
// Both walk the 128 x 128 band in raster order and run ll2_parity_nudge (enc_par.cuh) at every visited, untagged
// cell: depending on the PARITY of the cell, its right neighbours and the cells below it, it increments the next
// cell of the row (rule A1) or the cell below (rules A2, B).  What makes this tractable:
//   * an increment only ever hits an odd cell (every rule tests its target's parity first), so a cell is
//     incremented at most once and the state of the walk is the odd-mask of the band, which only loses bits;
//   * the value tests (|P[a] - P[a+2]| > 1, P[a+PS] < 10000) are only evaluated on odd = never incremented cells,
//     and the +16000/+24000 tags are even: both are fixed masks, taken after the quad-tagging pass;
//   * rules read the rows below as nobody has changed them yet (a row's own increments of the row below lie
//     behind the columns it still reads), so a row depends on the row above only through the odd bits that row
//     cleared in it;
//   * inside a row, A1 at column j clears the parity of j+1, which switches off every rule at j+1: "fire, then
//     not at the next column" = every other bit of each run of columns where A1 can fire.
// So one thread walks the 128 rows on 128-bit masks (about a hundred instructions per row instead of 3 x 128
// barrier-stepped wavefront steps), and the increments are applied by all threads afterwards.
#pragma once
#include "enc_patterns.cuh"

#define LL2M_ROWS 132
struct Ll2Masks {
	uint64_t odd[LL2M_ROWS][3];   // parity of columns 0..131 of rows 0..127 (the rules look 2 columns past the band); rows >= 128 zero
	uint64_t tag[LL2M_ROWS][2];   // > 10000 after the quad tagging
	uint64_t d2inc[LL2M_ROWS][2]; // in: |P[j] - P[j+2]| > 1, on the tagged values; out: the cells that get +1
};

// masks of one band row; R = the row after quad tagging, 132 readable columns
NHW_HD void ll2_masks_row(const int16_t *R, uint64_t *odd /*3*/, uint64_t *tag /*2*/, uint64_t *d2 /*2*/)
{
	odd[0] = odd[1] = odd[2] = 0; tag[0] = tag[1] = 0; d2[0] = d2[1] = 0;
	for (int j = 0; j < 132; j++) {
		const int v = R[j];
		if (v & 1) odd[j >> 6] |= 1ull << (j & 63);
		if (j < 128) {
			if (v > 10000) tag[j >> 6] |= 1ull << (j & 63);
			if (j < 126 && nhw_iabs(v - R[j + 2]) > 1) d2[j >> 6] |= 1ull << (j & 63);
		}
	}
}

// 128-bit helpers (two little-endian words); `hi` = the bits that shift in from columns 128..
NHW_HD void b128_shr(const uint64_t *a, uint64_t hi, int k, uint64_t *o) { o[0] = (a[0] >> k) | (a[1] << (64 - k)); o[1] = (a[1] >> k) | (hi << (64 - k)); }
NHW_HD void b128_shl1(const uint64_t *a, uint64_t *o) { o[1] = (a[1] << 1) | (a[0] >> 63); o[0] = a[0] << 1; }
NHW_HD void b128_alternate(const uint64_t *e, uint64_t *sel)
{
	const uint64_t EVEN = 0x5555555555555555ull;
	uint64_t sh[2], es[2], x[2];
	b128_shl1(e, sh);
	es[0] = e[0] & ~sh[0] & EVEN; es[1] = e[1] & ~sh[1] & EVEN;
	x[0] = e[0] + es[0];
	x[1] = e[1] + es[1] + (uint64_t)(x[0] < e[0]);
	sel[0] = (e[0] & ~x[0] & EVEN) | (e[0] & x[0] & ~EVEN);
	sel[1] = (e[1] & ~x[1] & EVEN) | (e[1] & x[1] & ~EVEN);
}

// skip_after_tag: the first recons call steps over the cell that follows a tagged one (it takes no turn).
// On return m.d2inc[r] holds the increment mask of row r.
NHW_HD void ll2_nudge_solve(Ll2Masks &m, int q, bool skip_after_tag)
{
	if (q <= 17) {   // every increment sits behind q > 17
		for (int r = 0; r < 128; r++) m.d2inc[r][0] = m.d2inc[r][1] = 0;
		return;
	}
	uint64_t O[2] = {m.odd[0][0], m.odd[0][1]}, carry[2] = {0, 0};
	for (int r = 0; r < 128; r++) {
		const uint64_t *T = m.tag[r], *Tb = m.tag[r + 1];
		uint64_t C[2], tsh[2];
		b128_shl1(T, tsh);
		C[0] = ~T[0] & (skip_after_tag ? ~tsh[0] : ~0ull);
		C[1] = ~T[1] & (skip_after_tag ? ~tsh[1] : ~0ull);
		uint64_t O1[2], O2[2], X[2], F1[2], F1s[2], cur[2], Acond[2], A1br[2], R1[2], R2[2];
		b128_shr(O, m.odd[r][2], 1, O1);
		b128_shr(O, m.odd[r][2], 2, O2);
		const uint64_t lo_ok = ~1ull, hi125 = ~0ull >> 2;   // column > 0; column < 126
		X[0] = C[0] & O[0] & O1[0] & O2[0] & m.d2inc[r][0] & lo_ok;
		X[1] = C[1] & O[1] & O1[1] & O2[1] & m.d2inc[r][1] & hi125;
		b128_alternate(X, F1);
		b128_shl1(F1, F1s);
		cur[0] = O[0] & ~F1s[0]; cur[1] = O[1] & ~F1s[1];
		Acond[0] = C[0] & cur[0] & O1[0] & lo_ok; Acond[1] = C[1] & cur[1] & O1[1];
		A1br[0] = Acond[0] & O2[0]; A1br[1] = Acond[1] & O2[1] & hi125;
		uint64_t incb[2] = {0, 0};
		const uint64_t *Ob = m.odd[r + 1];
		b128_shr(Ob, Ob[2], 1, R1);
		b128_shr(Ob, Ob[2], 2, R2);
		if (r < 127) {
			incb[0] = Acond[0] & ~A1br[0] & Ob[0] & R1[0] & ~R2[0] & ~Tb[0];
			incb[1] = Acond[1] & ~A1br[1] & Ob[1] & R1[1] & ~R2[1] & ~Tb[1];
		}
		if (r >= 1 && r < 125) {
			const uint64_t *Obb = m.odd[r + 2], *Obbb = m.odd[r + 3];
			incb[0] |= C[0] & cur[0] & ~(O1[0] & lo_ok) & Ob[0] & R1[0] & Obb[0] & ~Obbb[0] & ~Tb[0];
			incb[1] |= C[1] & cur[1] & ~O1[1] & Ob[1] & R1[1] & Obb[1] & ~Obbb[1] & ~Tb[1];
		}
		m.d2inc[r][0] = carry[0] | F1s[0]; m.d2inc[r][1] = carry[1] | F1s[1];
		carry[0] = incb[0]; carry[1] = incb[1];
		O[0] = Ob[0] & ~incb[0]; O[1] = Ob[1] & ~incb[1];
	}
}

// ---- what the two passes leave in a cell once the increments are known
// recons: P = the band cell (un-tagged, incremented), J = im_jpeg, tmp = highres_tmp (second call only)
NHW_HD void ll2_recons_apply_cell(int16_t *P, int16_t *J, int16_t *tmp, int inc, int part)
{
	int v = *P + inc;   // rule A1 also increments tagged cells
	if (v > 10000) { v -= 16000; *J = (int16_t)v; }
	else *J = (v > 0 && v < 256) ? (int16_t)(v & 65534) : (int16_t)v;
	*P = (int16_t)v;
	if (!part) *tmp = (int16_t)v;
}
// bytes pass: the sample value that goes on to the byte coder
NHW_HD int ll2_bytes_value(int p, int inc, int q)
{
	p += inc;   // rule A1 also increments tagged cells
	if (q > 17 && p > 10000) return p - (p > 20000 ? 24000 : 16000);
	return p;
}
