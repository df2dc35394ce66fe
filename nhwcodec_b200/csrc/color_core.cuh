// color_core.cuh -- per-pixel arithmetic of the reference's RGB -> YCbCr transform
// (downsample_YUV420, encoder/colorspace.c:55-214), shared by the stage kernel (front.cu)
// and the fused front end (front_fused.cu).
#pragma once
#include <stdint.h>

struct ColorParams {
	int mode;     // 0: q>=20   1: q18,q19   2: q17   3: q<=16 (integer)
	double yq;    // mode 1: (double)(float)Y_quant   encoder/colorspace.c:104-105
	int qtz;      // mode 3: encoder/colorspace.c:174-189
};

// IEEE-exact, never contracted into FMA: the reference's x86-64 build has no FMA and the
// truncations below sit on the rounding of every partial sum (hard part 2).
__device__ __forceinline__ void rgb_to_ycc(int c0, int c1, int c2, const ColorParams &p, int &Y, int &U, int &V)
{
	if (p.mode == 3) {
		Y = (((66 * c0 + 129 * c1 + 25 * c2) * p.qtz + 4194304) >> 23) + 16;
		U = (((-38 * c0 - 74 * c1 + 112 * c2) * p.qtz + 4194304) >> 23) + 128;
		V = (((112 * c0 - 94 * c1 - 18 * c2) * p.qtz + 4194304) >> 23) + 128;
	} else {
		double d0 = (double)c0, d1 = (double)c1, d2 = (double)c2;
		double s = __dadd_rn(__dadd_rn(__dmul_rn(0.299, d0), __dmul_rn(0.587, d1)), __dmul_rn(0.114, d2));
		double bu = __dadd_rn(__dsub_rn(__dmul_rn(-0.1687, d0), __dmul_rn(0.3313, d1)), __dmul_rn(0.5, d2));
		double bv = __dsub_rn(__dsub_rn(__dmul_rn(0.5, d0), __dmul_rn(0.4187, d1)), __dmul_rn(0.0813, d2));
		if (p.mode == 1) s = __dmul_rn(s, p.yq);
		else if (p.mode == 2) {
			s = __dmul_rn(s, 0.94);
			bu = __dmul_rn(bu, 0.94);
			bv = __dmul_rn(bv, 0.94);
		}
		Y = __double2int_rz(__dadd_rn(s, 0.5));
		float fu = __double2float_rn(bu), fv = __double2float_rn(bv);
		U = __float2int_rz(__fadd_rn(fu, fu >= 0.0f ? 128.5f : 128.4f));
		V = __float2int_rz(__fadd_rn(fv, fv >= 0.0f ? 128.5f : 128.4f));
	}
	if (U >> 8) U = U < 0 ? 0 : 255;
	if (V >> 8) V = V < 0 ? 0 : 255;
}


// ---- integer form of the q>=20 transform ---------------------------------------------
// The double/float expression above only ever leaves the exact rational value by ~1e-13
// (Y) or by float rounding that cannot cross an integer (U, V: the exact value lies on a
// 1e-4 grid, half a float ulp is < 8e-6), so trunc() of it equals the integer quotient
// except where the exact Y sum is an integer itself (299c0+587c1+114c2+500 divisible by
// 1000, 16782 of the 2^24 triples): there the rounding of the partial sums decides and the
// IEEE expression is evaluated.  Verified exhaustively over all 2^24 triples against the
// IEEE path (tests/test_frontend_gpu.py::test_color_fast_path_exhaustive).
__device__ __forceinline__ void rgb_to_ycc_q20(int c0, int c1, int c2, int &Y, int &U, int &V)
{
	const uint32_t s = 299u * c0 + 587u * c1 + 114u * c2 + 500u;
	const uint32_t q = __umulhi(s, 0x10624dd3u) >> 6;   // s / 1000
	Y = (int)q;
	if (s - 1000u * q == 0u) {
		double d0 = (double)c0, d1 = (double)c1, d2 = (double)c2;
		double t = __dadd_rn(__dadd_rn(__dmul_rn(0.299, d0), __dmul_rn(0.587, d1)), __dmul_rn(0.114, d2));
		Y = __double2int_rz(__dadd_rn(t, 0.5));
	}
	const int eu = -1687 * c0 - 3313 * c1 + 5000 * c2;   // 10000 * (the double expression, exact)
	const int ev = 5000 * c0 - 4187 * c1 - 813 * c2;
	const uint32_t vu = (uint32_t)(eu + (eu >= 0 ? 1285000 : 1284000));   // +128.5 / +128.4f (= 128.4 - 6.1e-6)
	const uint32_t vv = (uint32_t)(ev + (ev >= 0 ? 1285000 : 1284000));
	U = (int)(__umulhi(vu, 0xD1B71759u) >> 13);           // / 10000
	V = (int)(__umulhi(vv, 0xD1B71759u) >> 13);
	U = U > 255 ? 255 : U;
	V = V > 255 ? 255 : V;
}

// four pixels at once: one branch for the (rare) exact-tie case instead of one per pixel;
// uv[k] = U | V << 16
__device__ __forceinline__ void rgb4_to_ycc_q20(const int (&c0)[4], const int (&c1)[4], const int (&c2)[4], int (&Y)[4],
                                                uint32_t (&uv)[4])
{
	uint32_t rem[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t s = 299u * c0[k] + 587u * c1[k] + 114u * c2[k] + 500u;
		const uint32_t q = __umulhi(s, 0x10624dd3u) >> 6;
		Y[k] = (int)q;
		rem[k] = s - 1000u * q;
		const int eu = -1687 * c0[k] - 3313 * c1[k] + 5000 * c2[k];
		const int ev = 5000 * c0[k] - 4187 * c1[k] - 813 * c2[k];
		const uint32_t vu = (uint32_t)(eu + (eu >= 0 ? 1285000 : 1284000));
		const uint32_t vv = (uint32_t)(ev + (ev >= 0 ? 1285000 : 1284000));
		const uint32_t U = min(__umulhi(vu, 0xD1B71759u) >> 13, 255u);
		const uint32_t V = min(__umulhi(vv, 0xD1B71759u) >> 13, 255u);
		uv[k] = U | (V << 16);
	}
	if (rem[0] == 0u || rem[1] == 0u || rem[2] == 0u || rem[3] == 0u) {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (rem[k] == 0u) {
				const double d0 = (double)c0[k], d1 = (double)c1[k], d2 = (double)c2[k];
				const double t = __dadd_rn(__dadd_rn(__dmul_rn(0.299, d0), __dmul_rn(0.587, d1)), __dmul_rn(0.114, d2));
				Y[k] = __double2int_rz(__dadd_rn(t, 0.5));
			}
		}
	}
}

// ---- the same on packed pixels (px = c0 | c1 << 8 | c2 << 16): the three dot products as mixed-sign dp2a pairs
// (16-bit signed coefficients x unsigned pixel bytes), two instructions each instead of three multiply-adds, and no
// per-channel unpacking.  Same integers as rgb4_to_ycc_q20 (exhaustive test: test_color_fast_path_exhaustive).
__device__ __forceinline__ int dp2a_lo_su(int a, uint32_t b, int c)
{
	int d;
	asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, uint32_t b, int c)
{
	int d;
	asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
#define NHW_PACK16(lo, hi) ((int)(((uint32_t)(uint16_t)(int16_t)(lo)) | ((uint32_t)(uint16_t)(int16_t)(hi) << 16)))
__device__ __forceinline__ void rgb4px_to_ycc_q20(const uint32_t (&px)[4], int (&Y)[4], uint32_t (&uv)[4])
{
	uint32_t rem[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t s = (uint32_t)dp2a_hi_su(NHW_PACK16(114, 0), px[k], dp2a_lo_su(NHW_PACK16(299, 587), px[k], 500));
		const uint32_t q = __umulhi(s, 0x10624dd3u) >> 6;
		Y[k] = (int)q;
		rem[k] = s - 1000u * q;
		const int eu = dp2a_hi_su(NHW_PACK16(5000, 0), px[k], dp2a_lo_su(NHW_PACK16(-1687, -3313), px[k], 0));
		const int ev = dp2a_hi_su(NHW_PACK16(-813, 0), px[k], dp2a_lo_su(NHW_PACK16(5000, -4187), px[k], 0));
		const uint32_t vu = (uint32_t)(eu + (eu >= 0 ? 1285000 : 1284000));
		const uint32_t vv = (uint32_t)(ev + (ev >= 0 ? 1285000 : 1284000));
		const uint32_t U = min(__umulhi(vu, 0xD1B71759u) >> 13, 255u);
		const uint32_t V = min(__umulhi(vv, 0xD1B71759u) >> 13, 255u);
		uv[k] = U | (V << 16);
	}
	if (rem[0] == 0u || rem[1] == 0u || rem[2] == 0u || rem[3] == 0u) {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (rem[k] == 0u) {
				const double d0 = (double)(px[k] & 255u), d1 = (double)((px[k] >> 8) & 255u), d2 = (double)((px[k] >> 16) & 255u);
				const double t = __dadd_rn(__dadd_rn(__dmul_rn(0.299, d0), __dmul_rn(0.587, d1)), __dmul_rn(0.114, d2));
				Y[k] = __double2int_rz(__dadd_rn(t, 0.5));
			}
		}
	}
}
