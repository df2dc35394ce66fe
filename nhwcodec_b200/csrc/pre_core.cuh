// pre_core.cuh -- pair rule of the luma pre-sharpening pass (pre_processing loop B,
// encoder/image_processing.c:770-836,1926-1990, q17..q21), shared by front.cu and front_fused.cu.
#pragma once
#include "dwt_core.cuh"

// Loop B (image_processing.c:770-836,1926-1990 for q>16): per horizontal pair (res,count)
// of kernel values, nudge the two pixels.  `a` is the flag the PREVIOUS pair (raster order)
// leaves behind; it depends on that pair's own values only.
__device__ __forceinline__ int pair_flag(int res, int cnt)
{
	int ar = nhw_iabs(res), ac = nhw_iabs(cnt);
	if (ar > 10 && ar < 32 && ac >= 23) return 0;   // the two `continue` exits
	return (ac >= 16 && ac < 32 && ar >= 23) ? 1 : 0;
}

__device__ __forceinline__ void pair_nudge(int res, int cnt, int a, int &d0, int &d1)
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

