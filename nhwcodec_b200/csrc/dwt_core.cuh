// dwt_core.cuh -- per-output arithmetic of the reference's 1-D integer filters, written as
// "give me output e of this line" functions over an abstract loader so the same code serves
// row walks, column walks out of shared memory, and the host-compiled test harness.
//
//   forward, first pass : downfilter53IV   encoder/filters.c:346-386
//   forward, second pass: downfilter53VI   encoder/filters.c:203-287   (rows of the low band)
//                         downfilter53     encoder/filters.c:55-114    (rows of the high band)
//   inverse             : upfilter53I/III/VI   encoder/filters.c:521-572 == decoder/filters.c:143-194
#pragma once
#include "enc_img.cuh"

// symmetric rounding division by 2^sh (half = 2^(sh-1)): v>=0 ? (v+half)>>sh : -((-v+half)>>sh).
// For v<0 that is ceil((v-half)/2^sh) = (v+half-1)>>sh, hence the branch-free form.
NHW_HD int nhw_sround(int v, int half, int sh) { return (v + half + (v >> 31)) >> sh; }
NHW_HD int nhw_iabs(int v) { return v < 0 ? -v : v; }

// 5-tap low of every forward filter, mirror extension x[-1]=x[1], x[-2]=x[2], x[N]=x[N-2]
template <typename Ld>
NHW_HD int tap_low(Ld ld, int e, int N)
{
	int c = 2 * e;
	int xm2 = ld(c >= 2 ? c - 2 : 2), xm1 = ld(c >= 1 ? c - 1 : 1), x0 = ld(c), xp1 = ld(c + 1);
	int xp2 = ld(c + 2 < N ? c + 2 : N - 2);
	return 6 * x0 + 2 * (xm1 + xp1) - (xm2 + xp2);
}

// first-pass outputs (no normalisation, stored as int16 by the caller)
template <typename Ld>
NHW_HD int first_pass_high(Ld ld, int e, int N)
{
	if (e == N / 2 - 1) return (ld(N - 1) - ld(N - 2)) << 1;
	return 2 * ld(2 * e + 1) - (ld(2 * e) + ld(2 * e + 2));
}

// high-pass residue of the second-pass filters (filters.c:62-84,212-231): the pair parity
// flag `m` makes the odd output of each pair round its predictor up when both sums are odd.
template <typename Ld>
NHW_HD int tap_high_lifted(Ld ld, int e)
{
	int a = ld(2 * e) + ld(2 * e + 2);
	if ((e & 1) && (a & 1) && ((ld(2 * e - 2) + ld(2 * e)) & 1)) a++;
	return ld(2 * e + 1) - (a >> 1);
}

// remainder fed forward by downfilter53VI's low band (filters.c:245-246,266-274)
NHW_HD int vi_remainder(int r)
{
	if (r >= 0) { int q = r & 63; return q < 32 ? (q >> 2) : -((64 - q) >> 2); }
	int q = (-r) & 63;
	return q < 32 ? -(q >> 2) : ((64 - q) >> 2);
}

// one output of the second (column) pass.  `fine` selects downfilter53VI vs downfilter53.
template <typename Ld>
NHW_HD int second_pass_low(Ld ld, int e, int N, bool fine)
{
	int r = tap_low(ld, e, N);
	if (!fine) return nhw_sround(r, 8, 4);
	int acc = r;
	if (e > 0) acc += vi_remainder(tap_low(ld, e - 1, N));
	return nhw_sround((int)(int16_t)acc, 32, 6);
}

template <typename Ld>
NHW_HD int second_pass_high(Ld ld, int e, int N, bool fine)
{
	if (e == N / 2 - 1) {
		int d = ld(N - 1) - ld(N - 2);
		return fine ? (d >> 3) : ((d + 1) >> 1);
	}
	int r = tap_high_lifted(ld, e);
	if (fine) return nhw_sround(r, 4, 3);
	return r > 0 ? ((r + 1) >> 1) : (r >> 1);
}

// ---- inverse: outputs 2t and 2t+1 of upfilter53I(low) followed by upfilter53III / VI(high),
// M = band length, mirror h[-1]=h[0], h[M]=h[M-1], l[M]=l[M-1].  Every store in the
// reference is to a short, so intermediate results wrap to int16 exactly where it does.
template <typename Ll, typename Lh>
NHW_HD void inverse_pair(Ll l, Lh h, int t, int M, bool normalise, int &even, int &odd)
{
	int lt = l(t), lt1 = l(t + 1 < M ? t + 1 : M - 1);
	int ht = h(t), hm = h(t > 0 ? t - 1 : 0), hp = h(t + 1 < M ? t + 1 : M - 1);
	int16_t ev = (int16_t)(lt << 3);
	int16_t od = (int16_t)((t < M - 1) ? ((lt1 + lt) << 2) : (lt << 3));
	ev = (int16_t)(ev - ((ht + hm) << 1));
	od = (int16_t)(od + (6 * ht - hp - hm));
	if (normalise) {
		if (ev > 0) ev = (int16_t)(ev + 32);
		ev = (int16_t)(ev >> 6);
		if (od > 0) od = (int16_t)(od + 32);
		od = (int16_t)(od >> 6);
	}
	even = ev;
	odd = od;
}
