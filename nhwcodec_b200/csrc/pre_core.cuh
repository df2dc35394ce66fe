// pre_core.cuh -- pair rule of the luma pre-sharpening pass (pre_processing loop B,
// encoder/image_processing.c:770-836,1926-1990, q17..q21), shared by front.cu and front_fused.cu.
#pragma once
#include "dwt_core.cuh"

// Loop B (image_processing.c:770-836,1926-1990 for q>16): per horizontal pair (res,count)
// of kernel values, nudge the two pixels.  `a` is the flag the PREVIOUS pair (raster order)
// leaves behind; it depends on that pair's own values only.
NHW_HD int pair_flag(int res, int cnt)
{
	int ar = nhw_iabs(res), ac = nhw_iabs(cnt);
	if (ar > 10 && ar < 32 && ac >= 23) return 0;   // the two `continue` exits
	return (ac >= 16 && ac < 32 && ar >= 23) ? 1 : 0;
}

NHW_HD void pair_nudge(int res, int cnt, int a, int &d0, int &d1)
{
	int e;
	d0 = 0;
	d1 = 0;
	if (res > 201) { d0 -= 2; e = 4; }
	else if (res < -201) { d0 += 2; e = 3; }
	else if (res > 176) { d0 -= 1; e = 2; }
	else if (res < -176) { d0 += 1; e = 1; }
	else e = 0;
	if (cnt > 201) { if (e == 0 || e == 3) d1 -= 2; else if (e != 4) d1 -= 1; }
	else if (cnt < -201) { if (e == 0 || e == 4) d1 += 2; else if (e != 3) d1 += 1; }
	else if (cnt > 176) { if (e != 4) d1 -= 1; }
	else if (cnt < -176) { if (e != 3) d1 += 1; }

	if (res < 32 && res > 10) {
		if (nhw_iabs(cnt) >= 23) {
			if (res < 16) { if (cnt > 0 && cnt < 32 && res > 11) d1 += 1; d0 += 1; }
			else d0 += a ? 1 : 2;
			return;
		}
	} else if (res > -32 && res < -10) {
		if (nhw_iabs(cnt) >= 23) {
			if (res > -16) { if (cnt < 0 && cnt > -32 && res < -11) d1 -= 1; d0 -= 1; }
			else d0 -= a ? 1 : 2;
			return;
		}
	}
	if (cnt < 32 && cnt > 10) {
		if (nhw_iabs(res) >= 23) {
			if (cnt < 16) { if (res > 0 && res < 32 && cnt > 11) d0 += 1; d1 += 1; }
			else d1 += 2;
		}
	} else if (cnt > -32 && cnt < -10) {
		if (nhw_iabs(res) >= 23) {
			if (cnt > -16) { if (res < 0 && res > -32 && cnt < -11) d0 -= 1; d1 -= 1; }
			else d1 -= 2;
		}
	}
}


// ---- table form of the pair rule -------------------------------------------------------
// Every comparison above is against one of {0, +-10, +-11, +-16, +-23, +-32, +-176, +-201}, so a
// value only matters through which of 17 intervals it falls in, and a pair through
// (interval of res, interval of cnt, a).  One 16-bit entry per interval pair:
//   bits 0-2  d0 + 2 when a == 0     bits 3-5  d0 + 2 when a == 1
//   bits 6-8  d1 + 2                 bit  9    the flag this pair leaves behind
// pair_cat_table maps clamp(v, -255, 255) + 256 to the interval.  Checked against the branchy
// form over [-300, 300]^2 x {0,1} by tests/test_host_logic_cpu.py.
#define PAIR_CATS 17
NHW_HD int pair_cat(int v)
{
	const int a = v < 0 ? -v : v;
	int c = (a > 0) + (a > 10) + (a > 11) + (a > 15) + (a > 22) + (a > 31) + (a > 176) + (a > 201);
	return v < 0 ? 8 - c : 8 + c;
}
inline void pair_build_tables(uint8_t cat[512], uint16_t lut[PAIR_CATS * PAIR_CATS])
{
	static const int rep[PAIR_CATS] = {-250, -190, -100, -27, -19, -13, -11, -5, 0, 5, 11, 13, 19, 27, 100, 190, 250};
	for (int i = 0; i < 512; i++) {
		int v = i - 256;
		cat[i] = (uint8_t)pair_cat(v < -255 ? -255 : v);
	}
	for (int r = 0; r < PAIR_CATS; r++)
		for (int c = 0; c < PAIR_CATS; c++) {
			int d00, d10, d01, d11;
			pair_nudge(rep[r], rep[c], 0, d00, d10);
			pair_nudge(rep[r], rep[c], 1, d01, d11);
			lut[r * PAIR_CATS + c] = (uint16_t)((d00 + 2) | ((d01 + 2) << 3) | ((d10 + 2) << 6) | (pair_flag(rep[r], rep[c]) << 9));
		}
}
