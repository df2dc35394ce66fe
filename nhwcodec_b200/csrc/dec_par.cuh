// dec_par.cuh -- parallel forms of the decoder's raster-order stages (same approach as
// enc_par.cuh: wavefront cells with the smallest skew the dependency footprint allows, and
// pointwise forms where the raster order turns out not to matter).
//
//   D8   shrink isolated coefficients     decoder/nhw_decoder.c:685-711   wavefront, skew 2
//   D11  edge flags (+16000 in place)     decoder/nhw_decoder.c:789-825   wavefront over pairs, skew 2
//   D16  chroma sharpen                   decoder/nhw_decoder.c:1085-1109 wavefront, skew 2
//   D16  chroma 2x upsample               decoder/nhw_decoder.c:1137-1181 pointwise
// In all three wavefront stages a cell reads its 8 neighbours in place: the row above and the left
// neighbour already edited, the right neighbour and the row below not yet; a row may therefore run
// ahead of the row below by two cells (two pairs for D11) and no more.
#pragma once
#include "dec_stages.cuh"
#include "cells8.cuh"

NHW_HD int dwf_shrink_cell(int16_t *J, int r, int j)
{
	const int s = r * YW + j;
	if (nhw_iabs(J[s]) <= 8) return 1;
	if (nhw_iabs(J[s - YW - 1]) > 8 || nhw_iabs(J[s - YW]) > 8 || nhw_iabs(J[s - YW + 1]) > 8 ||
	    nhw_iabs(J[s - 1]) > 8 || nhw_iabs(J[s + 1]) > 8 || nhw_iabs(J[s + YW - 1]) > 8 ||
	    nhw_iabs(J[s + YW]) > 8 || nhw_iabs(J[s + YW + 1]) > 8)
		return 1;
	if (r >= 128 || j >= 128) J[s] += J[s] > 0 ? -1 : 1;
	return 1;
}

// D8 at q <= 16: same rule with the diagonal threshold at 16.  Valid when rows above r are final and rows below
// untouched; cells of row r may run in any order (see kd_shrink_y_lowq).
NHW_HD void dec_shrink_lowq_cell(int16_t *J, int r, int j)
{
	const int s = r * YW + j;
	if (nhw_iabs(J[s]) <= 8) return;
	if (nhw_iabs(J[s - YW - 1]) > 16 || nhw_iabs(J[s - YW]) > 8 || nhw_iabs(J[s - YW + 1]) > 16 ||
	    nhw_iabs(J[s - 1]) > 8 || nhw_iabs(J[s + 1]) > 8 || nhw_iabs(J[s + YW - 1]) > 16 ||
	    nhw_iabs(J[s + YW]) > 8 || nhw_iabs(J[s + YW + 1]) > 16)
		return;
	if (r >= 128 || j >= 128) J[s] += J[s] > 0 ? -1 : 1;
}

NHW_HD int dwf_edge_cell(int16_t *P, int r, int p)
{
	const int s = r * YW + 1 + 2 * p;
	const int res = dec_lap8(P, s, YW);
	const int cnt = dec_lap8(P, s + 1, YW);
	if (res > 41 && res < 108 && cnt < 16) P[s] += 16000;
	else if (res < -41 && res > -108 && cnt > -16) P[s] += 16000;
	else if (cnt > 41 && cnt < 108 && res < 16) P[s + 1] += 16000;
	else if (cnt < -41 && cnt > -108 && res > -16) P[s + 1] += 16000;
	return 1;
}

NHW_HD int dwf_sharpen_cell(int16_t *P, int thr, int r, int j)
{
	const int s = r * CW + j;
	const int res = dec_lap8(P, s, CW);
	if (nhw_iabs(res) > thr) {
		if (res > 0) P[s] += res > 160 ? 3 : 2;
		else P[s] -= res < -160 ? 3 : 2;
	}
	return 1;
}

// D16: the 2x2 output cells of chroma cell (r, c) from the un-clipped sharpened plane
// (clip, then rows, then columns; the last row / column is repeated)
NHW_HD void dec_c_upsample_cell(const int16_t *P, uint8_t *out /* 512x512 */, int r, int c)
{
	const int c1 = c < 255 ? c + 1 : 255, r1 = r < 255 ? r + 1 : 255;
	const int a00 = dec_clip8(P[r * CW + c]), a01 = dec_clip8(P[r * CW + c1]);
	const int a10 = dec_clip8(P[r1 * CW + c]), a11 = dec_clip8(P[r1 * CW + c1]);
	const int b0 = r < 255 ? (a00 + a10 + 1) >> 1 : a00;      // odd output row
	const int b1 = r < 255 ? (a01 + a11 + 1) >> 1 : a01;
	uint8_t *o = out + (2 * r) * 512 + 2 * c;
	o[0] = (uint8_t)a00;
	o[1] = (uint8_t)(c < 255 ? (a00 + a01 + 1) >> 1 : a00);
	o[512] = (uint8_t)b0;
	o[513] = (uint8_t)(c < 255 ? (b0 + b1 + 1) >> 1 : b0);
}

// ---- D4 in parallel form ---------------------------------------------------------------------
// dec_y_markers_image walks the band plane in three sweeps (rows 0..255; rows 256..511 left half;
// rows 256..511 right half).  Marker codes (> 1000) are sparse and only WRITE constants, so they are
// collected in sweep order and applied one after the other (a marker overwritten before its turn
// is no longer one: re-checked when applied).  The right-half sweep also nudges |v| in 9..15 by one
// when at least two of its 4-neighbours are small, in place; a nudged cell stays >= 9 in magnitude,
// so the only things a cell's test can see change are marker writes.  For the cell at s, in sweep
// order: the left and upper neighbours have had ALL their writers' turns (writers sit at most one
// cell to the right), the right and lower neighbours none -- so the test reads the plane after the
// markers (J) for the former and a snapshot taken before them (S) for the latter.
// W: cells of the right half written by a right-half marker.  A: cells holding an applied 1008/1009.
// ATOMIC: several threads apply (non-overlapping) markers at once, so the bit masks are updated with atomicOr (device only)
template <bool ATOMIC = false>
NHW_HD void dec_marker_apply(int16_t *J, int s, bool lower, uint32_t *W, uint32_t *A)
{
	const int v = J[s];
	if (v <= 1000) return;
	const int j = s & 511;
	auto set_bit = [&](uint32_t *M, int k) {
#ifdef __CUDA_ARCH__
		if (ATOMIC) { atomicOr(&M[k >> 5], 1u << (k & 31)); return; }
#endif
		M[k >> 5] |= 1u << (k & 31);
	};
	auto mark = [&](int t) {   // t in the right half of the lower rows
		if (W && (t & 511) >= 256 && (t >> 9) >= 256 && (t >> 9) < 512) {
			const int k = (((t >> 9) - 256) << 8) + ((t & 511) - 256);
			set_bit(W, k);
		}
	};
	if (!lower) {
		if (v == 1008) { if (s > 0) J[s - 1] = 5; J[s + 1] = 5; J[s] = (int16_t)(j < 256 ? 5 : 6); }
		else if (v == 1009) { if (s > 0) J[s - 1] = -5; J[s + 1] = -5; J[s] = (int16_t)(j < 256 ? -6 : -7); }
		else if (v == 1010) { J[s] = 5; J[s + 1] = 5; J[s + YW] = 5; J[s + YW + 1] = 5; }
		else if (v == 1011) { J[s] = -5; J[s + 1] = -5; J[s + YW] = -5; J[s + YW + 1] = -5; }
		else if (v == 1006) { J[s] = -6; J[s + 1] = -6; }
		else if (v == 1007) { J[s] = 6; J[s + 1] = 6; }
		return;
	}
	if (v == 1008 || v == 1009) {
		const int sg = v == 1008 ? 1 : -1;
		// (at the plane's very last cell the reference writes one element past its buffer: dropped, the guard stays zero)
		J[s - 1] = (int16_t)(5 * sg); J[s] = (int16_t)(v == 1008 ? 6 : -7); if (s + 1 < 512 * 512) J[s + 1] = (int16_t)(5 * sg);
		mark(s - 1); mark(s); if (j < 511) mark(s + 1);
		if (A && j >= 256) {
			const int k = (((s >> 9) - 256) << 8) + (j - 256);
			set_bit(A, k);
		}
	} else if (v == 1006 || v == 1007) {
		const int16_t w = (int16_t)(v == 1006 ? -7 : 7);
		if (j < 256) { J[s] = w; J[s + 1] = w; mark(s + 1); }
		else { J[s - 256] = w; J[s - 768] = w; J[s] = 0; mark(s); }
	}
}

// the nudge rule of the right-half sweep for cell s = (r, j), 256 <= r, 256 < j < 511.
// S = snapshot of the plane before the right-half markers, J = plane after them.
NHW_HD bool dec_dense_qualifies(const int16_t *S, const uint32_t *A, int s)
{
	const int k = (((s >> 9) - 256) << 8) + ((s & 511) - 256);
	if ((A[(k - 1) >> 5] >> ((k - 1) & 31)) & 1u) return false;   // an applied 1008/1009 on the left rewrote this cell first
	const int v = S[s];
	return v <= 1000 && nhw_iabs(v) > 8 && nhw_iabs(v) < 16;
}
NHW_HD int dec_dense_count(const int16_t *J, const int16_t *S, int s)
{
	const int below = (s >> 9) < 511 ? (int)S[s + YW] : 0;   // past the last row: the zero guard
	return (nhw_iabs(J[s - 1]) < 8) + (nhw_iabs(S[s + 1]) < 8) + (nhw_iabs(J[s - YW]) < 8) + (nhw_iabs(below) < 8);
}
