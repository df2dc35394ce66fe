// tests/hostcheck/check_tables.cpp -- TEST TOOLING: host-side checks of table / closed forms used
// by the CUDA kernels against the branchy statements of the same rules.
#include <cstdio>
#include <cstdint>
#include "../../nhwcodec_b200/csrc/pre_core.cuh"

static int sround_ref(int v, int half, int sh) { return v >= 0 ? ((v + half) >> sh) : -((-v + half) >> sh); }

int main()
{
	long bad = 0;
	for (int v = -70000; v <= 70000; v++) {
		if (nhw_sround(v, 32, 6) != sround_ref(v, 32, 6)) bad++;
		if (nhw_sround(v, 8, 4) != sround_ref(v, 8, 4)) bad++;
		if (nhw_sround(v, 4, 3) != sround_ref(v, 4, 3)) bad++;
	}
	printf("sround mismatches %ld\n", bad);
	static uint8_t cat[512];
	static uint16_t lut[PAIR_CATS * PAIR_CATS];
	pair_build_tables(cat, lut);
	long badp = 0;
	for (int res = -2100; res <= 2100; res++)
		for (int cnt = -2100; cnt <= 2100; cnt += (cnt > -320 && cnt < 320) ? 1 : 7)
			for (int a = 0; a < 2; a++) {
				int d0, d1;
				pair_nudge(res, cnt, a, d0, d1);
				const int f = pair_flag(res, cnt);
				const int cr = cat[(res < -255 ? -255 : res > 255 ? 255 : res) + 256];
				const int cc = cat[(cnt < -255 ? -255 : cnt > 255 ? 255 : cnt) + 256];
				const int e = lut[cr * PAIR_CATS + cc];
				const int t0 = ((e >> (3 * a)) & 7) - 2, t1 = ((e >> 6) & 7) - 2, tf = e >> 9;
				if (t0 != d0 || t1 != d1 || tf != f) badp++;
			}
	printf("pair table mismatches %ld\n", badp);
	return (bad || badp) ? 1 : 0;
}
