// cells8.cuh -- helpers shared by the cell-group forms of the encoder (enc_cells.cuh) and decoder (dec_par.cuh):
// 8 consecutive int16 cells move as one 16-byte word.
#pragma once
#include "dwt_core.cuh"

NHW_HD void ld8(const int16_t *p, int *v)
{
#ifdef __CUDA_ARCH__
	const uint4 w = *reinterpret_cast<const uint4 *>(p);
	v[0] = (int16_t)(w.x & 0xffff); v[1] = (int16_t)(w.x >> 16);
	v[2] = (int16_t)(w.y & 0xffff); v[3] = (int16_t)(w.y >> 16);
	v[4] = (int16_t)(w.z & 0xffff); v[5] = (int16_t)(w.z >> 16);
	v[6] = (int16_t)(w.w & 0xffff); v[7] = (int16_t)(w.w >> 16);
#else
	for (int k = 0; k < 8; k++) v[k] = p[k];
#endif
}
NHW_HD void st8(int16_t *p, const int *v)
{
#ifdef __CUDA_ARCH__
	uint4 w;
	w.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
	w.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
	w.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
	w.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
	*reinterpret_cast<uint4 *>(p) = w;
#else
	for (int k = 0; k < 8; k++) p[k] = (int16_t)v[k];
#endif
}

// ---- isolated-coefficient shrink of the level-2 region: encoder offsetY_recons256 (image_processing.c:3162-3187,
// wavefront form wf_shrink_cell, |J| >= 8) and its decoder twin (wavefront form dwf_shrink_cell, |J| > 8).
// A cell with |J| >= t moves one step towards zero when none of its 8 neighbours has |J| >= t.  Raster order does
// not matter: a neighbour that could change (|n| >= t) would itself need this cell to be below t to do so, and then
// this cell is no candidate; neighbours below t never change.  So the test reads the same on the plane before the
// stage, and a value read while another group stores its result gives the same answer too.
NHW_HD bool shrink_cells8(const int16_t *J /* plane, row stride 512 */, int r, int g, int t /* 8 encoder, 9 decoder */, int *o)
{
	if (r < 1 || r > 254) return false;
	const int c = g * 8;
	const int16_t *R = J + r * 512;
	int v[10], up[10], dn[10];
	ld8(R + c, v + 1);
	bool cand = false;
	for (int k = 1; k <= 8; k++) cand |= nhw_iabs(v[k]) >= t;
	if (!cand) return false;
	v[0] = R[c - 1]; v[9] = R[c + 8];
	ld8(R - 512 + c, up + 1); up[0] = R[-512 + c - 1]; up[9] = R[-512 + c + 8];
	ld8(R + 512 + c, dn + 1); dn[0] = R[512 + c - 1]; dn[9] = R[512 + c + 8];
	bool any = false;
	for (int k = 1; k <= 8; k++) {
		const int j = c + k - 1;
		o[k - 1] = v[k];
		if (j < 1 || j > 254 || nhw_iabs(v[k]) < t) continue;
		if (nhw_iabs(up[k - 1]) >= t || nhw_iabs(up[k]) >= t || nhw_iabs(up[k + 1]) >= t || nhw_iabs(v[k - 1]) >= t ||
		    nhw_iabs(v[k + 1]) >= t || nhw_iabs(dn[k - 1]) >= t || nhw_iabs(dn[k]) >= t || nhw_iabs(dn[k + 1]) >= t)
			continue;
		if (r >= 128 || j >= 128) { o[k - 1] += v[k] > 0 ? -1 : 1; any = true; }
	}
	return any;
}

