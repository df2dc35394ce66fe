// tests/hostcheck/check_codebook.cpp -- TEST TOOLING: the error exits of the entropy stage's alphabet construction
// (nhwcodec_b200/csrc/enc_pack.cuh: pack_alphabet), which stand for the reference's exit(-1) at
// encoder/compress_pixel.c:234,270-271.  No natural input we know of reaches them on the GPU, so the branches are driven
// here with hand-made histograms: the same __host__ __device__ functions the kernels call, compiled with g++.
#include <cstdio>
#include <cstring>
#include "../../nhwcodec_b200/csrc/enc_pack.cuh"

static int run(PackState &st, int part, int &select, int &k, int &b)
{
	select = part ? 3 : 4;
	k = b = 0;
	return pack_alphabet(st, part, select, k, b);
}

int main()
{
	int fails = 0, select, k, b;
	static PackState st;
	auto clear = [&]() { memset(&st, 0, sizeof st); };
	auto expect = [&](const char *what, bool ok) { if (!ok) { printf("FAIL %s (select %d, k %d, b %d)\n", what, select, k, b); fails++; } };

	// 1. a small alphabet: no error, the zero-run marker ranks first (b = 1)
	clear();
	st.rle_buf[128] = 100000;
	for (int i = 0; i < 40; i += 2) st.rle_buf[i] = 1000 - i;
	for (int j = 4; j < 30; j++) st.rle_128[j] = 50;
	expect("small alphabet", run(st, 0, select, k, b) == 0 && select == 4 && b == 1 && k == 20 + 26 + 1);

	// 2. every symbol of the candidate alphabet and every run length present: 105 + 252 > 354 entries, so the minimum coded run
	//    length is raised until they fit (compress_pixel.c:129-229); luma with the marker first may use all 354 (zone escape)
	clear();
	for_each_symbol([&](int i) { st.rle_buf[i] = 10; });
	st.rle_buf[128] = 1 << 20;
	for (int j = 2; j < 256; j++) st.rle_128[j] = 3;
	{
		const int rc = run(st, 0, select, k, b);
		expect("full alphabet raises select", select > 4 && k <= 354);
		expect("full alphabet, luma: k > 290 is an error once select left 4", rc == NHW_ERR_CODEBOOK_DEV);
	}

	// 3. luma whose most frequent entry is NOT the zero-run marker (b = 0) with more than 290 entries: the reference exits
	clear();
	for_each_symbol([&](int i) { st.rle_buf[i] = 5; });
	st.rle_buf[2] = 1 << 22;
	st.rle_buf[128] = 7;
	for (int j = 4; j < 200; j++) st.rle_128[j] = 2;
	expect("luma, b == 0, k > 290", run(st, 0, select, k, b) == NHW_ERR_CODEBOOK_DEV && b == 0 && k > 290);

	// 4. the same histograms with few enough entries pass
	clear();
	for_each_symbol([&](int i) { st.rle_buf[i] = 5; });
	st.rle_buf[2] = 1 << 22;
	st.rle_buf[128] = 7;
	for (int j = 4; j < 150; j++) st.rle_128[j] = 2;
	expect("luma, b == 0, k <= 290", run(st, 0, select, k, b) == 0 && k <= 290);

	// 5. chroma (part 1, select starts at 3): more than 290 entries is an error there too
	clear();
	for_each_symbol([&](int i) { st.rle_buf[i] = 5; });
	st.rle_buf[128] = 1 << 20;
	for (int j = 3; j < 200; j++) st.rle_128[j] = 2;
	expect("chroma, k > 290", run(st, 1, select, k, b) == NHW_ERR_CODEBOOK_DEV);

	// 6. the stable sort keeps the enumeration order among equal weights (the reference's bubble sort does)
	clear();
	st.rle_buf[128] = 1000;
	st.rle_buf[4] = st.rle_buf[8] = st.rle_buf[12] = 77;
	st.rle_128[9] = 77;
	expect("no error", run(st, 0, select, k, b) == 0);
	expect("stable order", st.sym[1] == ((9 << 8) | 128) && st.sym[2] == ((1 << 8) | 4) && st.sym[3] == ((1 << 8) | 8) && st.sym[4] == ((1 << 8) | 12));
	printf("codebook checks failed: %d\n", fails);
	return fails ? 1 : 0;
}
