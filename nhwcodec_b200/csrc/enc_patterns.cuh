// enc_patterns.cuh -- the "small coefficient pattern" substitutions of the level-2 region as bit-mask algebra:
//   offsetY_recons256 patterns  encoder/image_processing.c:2757-2849   (serial form: recons_pattern_cell, enc_y1.cuh)
//   offsetY patterns            encoder/image_processing.c:239-290     (wavefront form: wf_offset_patterns_cell)
//
// Both loops walk a region in raster order; at cursor (r, j), with the cell and its left neighbour in the same
// class (4..7 or -7..-4), they fire
//   * a TRIPLE when the right neighbour is in the class too, else
//   * a BLOCK when the two cells below (r+1, j-1) and (r+1, j) are,
// rewrite some of those cells with tags / zeros and skip the next column.  What makes this tractable:
//   * every rewrite takes a cell OUT of its class and nothing ever enters one, cells a row rewrites in itself lie
//     behind its cursor, and the cells it reads in the row below are not the ones its earlier blocks rewrote:
//     whether an event CAN fire at (r, j) is a function of the class bits of row r as the row above left them and
//     of the original class bits of row r+1;
//   * "fire, then skip one column" picks every other cell in each run of cells that can fire (counted from the
//     start of the run): a carry trick on the 256-bit row mask;
//   * a row depends on the row above only through the blocks that fired there (they clear two class bits).
// So rows are solved on 256-bit masks (a few hundred instructions per row instead of a 256-column wavefront), and
// the rewrites -- no two events touch the same cell -- are applied in parallel afterwards.
#pragma once
#include "enc_cells.cuh"

struct PatMasks {
	uint64_t pos[257][4], neg[257][4];   // class bits of the rows as they are before the stage (bit j = column j)
	uint64_t ft[256][4], fb[256][4];     // cursors where a triple / a block fired
};

NHW_HD void pat_class_bits8(const int *v, uint32_t &pos8, uint32_t &neg8)
{
	pos8 = 0; neg8 = 0;
	for (int k = 0; k < 8; k++) {
		if (v[k] > 3 && v[k] < 8) pos8 |= 1u << k;
		else if (v[k] < -3 && v[k] > -8) neg8 |= 1u << k;
	}
}

// 256-bit helpers, little-endian words
NHW_HD void b256_shl1(const uint64_t *a, uint64_t *o) { o[3] = (a[3] << 1) | (a[2] >> 63); o[2] = (a[2] << 1) | (a[1] >> 63); o[1] = (a[1] << 1) | (a[0] >> 63); o[0] = a[0] << 1; }
NHW_HD void b256_shr1(const uint64_t *a, uint64_t *o) { o[0] = (a[0] >> 1) | (a[1] << 63); o[1] = (a[1] >> 1) | (a[2] << 63); o[2] = (a[2] >> 1) | (a[3] << 63); o[3] = a[3] >> 1; }
NHW_HD void b256_add(const uint64_t *a, const uint64_t *b, uint64_t *o)
{
	uint64_t carry = 0;
	for (int w = 0; w < 4; w++) {
		const uint64_t s = a[w] + b[w], s2 = s + carry;
		carry = (uint64_t)(s < a[w]) | (uint64_t)(s2 < s);
		o[w] = s2;
	}
}
// every other bit of each run of ones, counted from the run's lowest bit
NHW_HD void b256_alternate(const uint64_t *e, uint64_t *sel)
{
	const uint64_t EVEN = 0x5555555555555555ull;
	uint64_t sh[4], es[4], x[4];
	b256_shl1(e, sh);
	for (int w = 0; w < 4; w++) es[w] = e[w] & ~sh[w] & EVEN;   // runs that start on an even column
	b256_add(e, es, x);                                         // ... are wiped out by the carry
	for (int w = 0; w < 4; w++) sel[w] = (e[w] & ~x[w] & EVEN) | (e[w] & x[w] & ~EVEN);
}

// kind 0: offsetY_recons256 (rows 0..254; cursor columns 129..254 above row 128, 1..254 from row 128 on)
// kind 1: offsetY           (rows 0..255; cursor columns 1..254)
NHW_HD int pat_rows(int kind) { return kind ? 256 : 255; }
// one row, given the class bits the blocks of the row above cleared
NHW_HD void pat_solve_row(PatMasks &m, int kind, int r, const uint64_t *clear)
{
	uint64_t range[4] = {~1ull, ~0ull, ~0ull, ~0ull >> 1};     // columns 1..254
	if (!kind && r < 128) { range[0] = 0; range[1] = 0; range[2] = ~1ull; }   // columns 129..254
	uint64_t t[4] = {0, 0, 0, 0}, bk[4] = {0, 0, 0, 0};
	for (int s = 0; s < 2; s++) {
		uint64_t c[4], l[4], rt[4], bl[4];
		const uint64_t *cur = s ? m.neg[r] : m.pos[r], *b = s ? m.neg[r + 1] : m.pos[r + 1];
		for (int w = 0; w < 4; w++) c[w] = cur[w] & ~clear[w];
		b256_shl1(c, l);
		b256_shr1(c, rt);
		b256_shl1(b, bl);
		for (int w = 0; w < 4; w++) {
			const uint64_t pair = c[w] & l[w];
			t[w] |= pair & rt[w];
			bk[w] |= pair & ~rt[w] & bl[w] & b[w];
		}
	}
	uint64_t e[4], sel[4];
	for (int w = 0; w < 4; w++) { t[w] &= range[w]; bk[w] &= range[w]; e[w] = t[w] | bk[w]; }
	b256_alternate(e, sel);
	for (int w = 0; w < 4; w++) { m.ft[r][w] = sel[w] & t[w]; m.fb[r][w] = sel[w] & bk[w]; }
}
// a block at cursor j takes (r+1, j-1) and (r+1, j) out of their class
NHW_HD bool pat_cleared_by(const PatMasks &m, int r, uint64_t *clear)
{
	uint64_t fbr[4];
	b256_shr1(m.fb[r], fbr);
	uint64_t any = 0;
	for (int w = 0; w < 4; w++) { clear[w] = m.fb[r][w] | fbr[w]; any |= clear[w]; }
	return any != 0;
}
// Rows only depend on the row above through its blocks, which are rare: every row is first solved as if the row
// above had fired none (pat_solve_row with clear = 0, all rows in parallel), then one walk down the rows redoes
// those that follow a row with blocks.
NHW_HD void pat_fixup(PatMasks &m, int kind)
{
	const int rows = pat_rows(kind);
	for (int r = 1; r < rows; r++) {
		uint64_t clear[4];
		if (pat_cleared_by(m, r - 1, clear)) pat_solve_row(m, kind, r, clear);
	}
}

// the rewrites of one fired event
NHW_HD void pat_apply(int16_t *P, int16_t *J, int kind, int r, int j, bool triple, bool positive)
{
	const int a = r * YW + j;
	if (!kind) {
		if (triple) {
			P[a - 1] = (int16_t)(positive ? 15300 : 15400); P[a] = 0;
			J[a] = (int16_t)(positive ? 5 : -6); J[a + 1] = (int16_t)(positive ? 5 : -5);
		} else {
			const int16_t tag = (int16_t)(positive ? 15500 : 15600), jv = (int16_t)(positive ? 5 : -5);
			P[a - 1] = tag; J[a] = jv;
			P[a + YW - 1] = tag; J[a + YW] = jv;
			P[a + YW] = 0;
		}
	} else {
		if (triple) { P[a] = (int16_t)(positive ? 12700 : 12900); P[a - 1] = 10100; }
		else { P[a - 1] = (int16_t)(positive ? 12100 : 12200); P[a] = 10100; P[a + YW - 1] = 10100; P[a + YW] = 10100; }
	}
}
