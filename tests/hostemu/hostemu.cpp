// tests/hostemu/hostemu.cpp -- TEST TOOLING ONLY, never part of the product library.
//
// Compiles the encoder's stage functions (nhwcodec_b200/csrc/enc_*.cuh, written as
// __host__ __device__ code) with g++ and runs them sequentially on the CPU, so stage parity
// against the reference taps can be iterated on a machine without a GPU.  It proves nothing
// about the CUDA path by itself: the parity claims come from tests/ running the CUDA library
// on a B200 against oracle/_ref.  Rows/strips are deliberately visited in REVERSE order where
// the CUDA path runs them as independent threads, to catch hidden cross-row dependencies.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../nhwcodec_b200/csrc/dec_par.cuh"
#include "../../nhwcodec_b200/csrc/dec_parse.h"
#include "../../nhwcodec_b200/csrc/enc_ll2_masks.cuh"
#include "../../nhwcodec_b200/csrc/enc_lowq.cuh"
#include "../../nhwcodec_b200/csrc/pre_lowq.cuh"
#include "../../nhwcodec_b200/csrc/enc_hq.cuh"
#include "../../nhwcodec_b200/csrc/dec_stages.cuh"
#include "serial_forms.cuh"

namespace {

struct Work {
	std::vector<uint8_t> mem;
	EncImg im;
	EncHdr hdr;
};

template <typename T>
T *carve(std::vector<uint8_t> &mem, size_t &off, size_t count, size_t guard_bytes)
{
	off = (off + 63) & ~size_t(63);
	off += guard_bytes;
	T *p = reinterpret_cast<T *>(mem.data() + off);
	off += count * sizeof(T) + guard_bytes;
	return p;
}

Work *make_work()
{
	Work *w = new Work();
	w->mem.assign(16u << 20, 0);
	size_t off = 0;
	EncImg &im = w->im;
	const size_t G = NHW_GUARD_S * 2;
	im.proc = carve<int16_t>(w->mem, off, 512 * 512, G);
	im.jpeg = carve<int16_t>(w->mem, off, 512 * 512, G);
	im.aux = carve<int16_t>(w->mem, off, 512 * 512, G);
	im.ll1 = carve<int16_t>(w->mem, off, 256 * 256, G);
	im.ll2s = carve<int16_t>(w->mem, off, 256 * 256, G);
	im.cproc = carve<int16_t>(w->mem, off, 256 * 256, G);
	im.cjpeg = carve<int16_t>(w->mem, off, 256 * 256, G);
	im.caux = carve<int16_t>(w->mem, off, 256 * 256, G);
	im.cll1 = carve<int16_t>(w->mem, off, 128 * 128, G);
	im.cll2s = carve<int16_t>(w->mem, off, 128 * 128, G);
	im.scan = carve<uint8_t>(w->mem, off, NHW_SCAN_BYTES, NHW_GUARD_B);
	im.tree1 = carve<uint8_t>(w->mem, off, NHW_CAP_TREE1, NHW_GUARD_B);
	im.ch_res = carve<uint8_t>(w->mem, off, 16384, NHW_GUARD_B);
	im.llcode = carve<uint8_t>(w->mem, off, 49152, NHW_GUARD_B);
	im.exw = carve<uint8_t>(w->mem, off, 3 * 16384, NHW_GUARD_B);
	im.exw_uv = carve<uint8_t>(w->mem, off, 2 * 16384, NHW_GUARD_B);
	im.res1 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.res1_bit = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.res1_word = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.res3 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.res3_bit = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.res3_word = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 4 + 16, NHW_GUARD_B);
	im.res4 = carve<uint8_t>(w->mem, off, 8192, NHW_GUARD_B);
	im.res5 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.res5_bit = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.res5_word = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.res6 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.res6_bit = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.res6_word = carve<uint8_t>(w->mem, off, NHW_CAP_LIST / 8 + 16, NHW_GUARD_B);
	im.char_res1 = carve<uint16_t>(w->mem, off, NHW_CAP_CHAR_RES1, NHW_GUARD_B);
	im.qsetting3 = carve<uint32_t>(w->mem, off, NHW_CAP_QSETTING3, NHW_GUARD_B);
	im.hq_qs = carve<int16_t>(w->mem, off, 256 * 512, NHW_GUARD_B);
	im.hq_fo = carve<int16_t>(w->mem, off, 256 * 256, NHW_GUARD_B);
	im.hq_band = carve<int16_t>(w->mem, off, 256 * 256, NHW_GUARD_B);
	im.hq_tag = carve<uint8_t>(w->mem, off, 256 * 512, NHW_GUARD_B);
	im.tmp1 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.tmp2 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.tmp3 = carve<uint8_t>(w->mem, off, NHW_CAP_LIST, NHW_GUARD_B);
	im.highres_mem = carve<uint16_t>(w->mem, off, 16384, NHW_GUARD_B);
	im.highres_word = carve<uint8_t>(w->mem, off, 16384, NHW_GUARD_B);
	im.res_uv64 = carve<uint8_t>(w->mem, off, 1024, NHW_GUARD_B);
	im.sel1 = carve<uint8_t>(w->mem, off, 32768 + 16, NHW_GUARD_B);
	im.sel2 = carve<uint8_t>(w->mem, off, 32768 + 16, NHW_GUARD_B);
	im.codebook1 = carve<uint8_t>(w->mem, off, 1024, NHW_GUARD_B);
	im.codebook2 = carve<uint8_t>(w->mem, off, 1024, NHW_GUARD_B);
	im.words = carve<uint32_t>(w->mem, off, 131072, NHW_GUARD_B);
	im.pack_scratch = carve<PackState>(w->mem, off, 1, NHW_GUARD_B);
	im.hdr = &w->hdr;
	if (off > w->mem.size()) abort();
	memset(&w->hdr, 0, sizeof w->hdr);
	return w;
}

// ---- plain loops over the shared 1-D filter primitives ----
// forward level: in(y,x) natural, N x N, row stride si -> out P(k,m), row stride so; tmp N*N
void fwd_level(const int16_t *in, int si, bool in_transposed, int16_t *out, int so, int N, std::vector<int16_t> &tmp)
{
	tmp.resize((size_t)N * N);
	for (int y = 0; y < N; y++) {
		auto ld = [&](int x) { return (int)(in_transposed ? in[x * si + y] : in[y * si + x]); };
		for (int e = 0; e < N / 2; e++) {
			tmp[y * N + e] = (int16_t)tap_low(ld, e, N);
			tmp[y * N + N / 2 + e] = (int16_t)first_pass_high(ld, e, N);
		}
	}
	for (int k = 0; k < N; k++) {
		auto ld = [&](int y) { return (int)tmp[y * N + k]; };
		const bool fine = k < N / 2;
		for (int e = 0; e < N / 2; e++) {
			out[k * so + e] = (int16_t)second_pass_low(ld, e, N, fine);
			out[k * so + N / 2 + e] = (int16_t)second_pass_high(ld, e, N, fine);
		}
	}
}

// inverse level: bands J(k,m) in `in` (stride s) -> natural image in `out` (stride s)
void inv_level(const int16_t *in, int16_t *out, int s, int N, std::vector<int16_t> &tmp)
{
	const int M = N / 2;
	tmp.resize((size_t)N * N);
	for (int k = 0; k < N; k++) {
		auto l = [&](int t) { return (int)in[k * s + t]; };
		auto h = [&](int t) { return (int)in[k * s + M + t]; };
		for (int t = 0; t < M; t++) {
			int ev, od;
			inverse_pair(l, h, t, M, false, ev, od);
			tmp[k * N + 2 * t] = (int16_t)ev;
			tmp[k * N + 2 * t + 1] = (int16_t)od;
		}
	}
	for (int y = 0; y < N; y++) {
		auto l = [&](int t) { return (int)tmp[t * N + y]; };
		auto h = [&](int t) { return (int)tmp[(M + t) * N + y]; };
		for (int t = 0; t < M; t++) {
			int ev, od;
			inverse_pair(l, h, t, M, true, ev, od);
			out[y * s + 2 * t] = (int16_t)ev;
			out[y * s + 2 * t + 1] = (int16_t)od;
		}
	}
}

void copy_region(int16_t *dst, int ds, const int16_t *src, int ss, int N)
{
	for (int r = 0; r < N; r++) memcpy(dst + r * ds, src + r * ss, N * sizeof(int16_t));
}

typedef void (*tap_fn)(const char *, const void *, size_t);

// segment-parallel peephole, run sequentially in the kernel's phase order
void host_peephole(const EncImg &im)
{
	uint8_t *s = im.scan;
	const int N = 262144;
	std::vector<int> heads;
	for (int i = 0; i < N - 4; i++)
		if (peep_pair_candidate(s, i, N) && !peep_pair_candidate(s, i - 4, N)) heads.push_back(i);
	for (size_t k = heads.size(); k-- > 0;) peep_merge_chain(s, heads[k], N);
	s[0] = s[1] = s[2] = s[3] = 128;
	s[N - 4] = s[N - 3] = s[N - 2] = s[N - 1] = 128;
	std::vector<uint8_t> out(N);
	std::vector<uint32_t> bits(N / 32, 0);
	for (int i = 0; i < N; i++) if (s[i] != 128) bits[i >> 5] |= 1u << (i & 31);
	for (int i = 0; i < N; i += 4) {
		uint32_t w; memcpy(&w, s + i, 4);
		if (nz_mask4(w) != ((bits[i >> 5] >> (i & 31)) & 15u)) abort();
	}
	const NzBits nz{bits.data(), 0, N, nullptr};
	int sel1 = 0, sel2 = 0;
	for (int i = N - 1; i >= 0; i--) {
		int a, b;
		out[i] = (uint8_t)peep_select_byte(s, nz, i, N, a, b);
		sel1 += a;
		sel2 += b;
	}
	memcpy(s, out.data(), N);
	im.hdr->select1 = sel1;
	im.hdr->select2 = sel2;
}

// segment-parallel entropy stage for one stream; word0 = first output word, returns status
int host_entropy(const EncImg &im, int part, int &word0)
{
	PackState &st = *static_cast<PackState *>(im.pack_scratch);
	uint8_t *s = im.scan;
	EncHdr *h = im.hdr;
	const int p1 = part ? 262144 : 0, p2 = part ? 393216 : 262144;
	uint8_t saved = 0;
	if (!part) { saved = s[262144]; s[262144] = 3; } else s[393215] = s[393214];
	const int S = (p2 - p1) / SEG_THREADS;
	std::vector<uint32_t> nzb((p2 - p1) / 32, 0);
	for (int i = p1; i < p2; i++) if (s[i] != 128) nzb[(i - p1) >> 5] |= 1u << ((i - p1) & 31);
	std::vector<uint32_t> nzs((p2 - p1) / 1024, 0);
	for (size_t j = 0; j < nzs.size(); j++) for (int k = 0; k < 32; k++) if (nzb[32 * j + k]) nzs[j] |= 1u << k;
	// work-balanced segment cut of k_entropy: segment t starts at the (t*T/256)-th non-zero byte
	std::vector<int> bnd(SEG_THREADS + 1, p2);
	bnd[0] = p1;
	if (!getenv("HE_EQUAL_SEGMENTS")) {
		std::vector<int> nzpos;
		for (int i = p1; i < p2; i++) if (s[i] != 128) nzpos.push_back(i);
		const long long T = (long long)nzpos.size();
		for (int t = 1; t < SEG_THREADS && T > 0; t++) bnd[t] = nzpos[(size_t)(t * T / SEG_THREADS)];
	}
	SegStream ss{s, p1, p2, S, NzBits{nzb.data(), p1, p2 - p1, nzs.data()}, getenv("HE_EQUAL_SEGMENTS") ? nullptr : bnd.data()};
	for (int i = 0; i < 256; i++) { st.rle_buf[i] = 0; st.rle_128[i] = 0; }
	for (int t = SEG_THREADS - 1; t >= 0; t--)
		seg_stats(ss, t, [&](bool run, int idx) { if (run) st.rle_128[idx]++; else st.rle_buf[idx]++; });
	int select = part ? 3 : 4, k = 0, b = 0;
	int rc = pack_alphabet(st, part, select, k, b);
	if (rc) return rc;
	const bool zone = (part == 0 && select == 4 && b == 1);
	std::vector<long> bits(SEG_THREADS + 1, 0), n1(SEG_THREADS + 1, 0), n2(SEG_THREADS + 1, 0);
	int bad = 0;
	for (int t = 0; t < SEG_THREADS; t++) {
		long nb = 0, a1 = 0, a2 = 0;
		bad |= seg_emit(ss, t, st.rle_buf, st.rle_128, select, zone, [&](uint32_t, int len) { nb += len; },
		                [&](int) { a1++; }, [&](int) { a2++; });
		bits[t + 1] = bits[t] + nb; n1[t + 1] = n1[t] + a1; n2[t + 1] = n2[t] + a2;
	}
	if (bad) return NHW_ERR_CODEBOOK_DEV;
	const long total = bits[SEG_THREADS];
	const int nwords = total > 0 ? (int)((total + 31) / 32) : 1;
	if (word0 + nwords >= NHW_WORDS_LIMIT) return NHW_ERR_OVERFLOW_DEV;
	if (!part) {
		memset(im.sel1, 0, (n1[SEG_THREADS] >> 3) + 1);
		memset(im.sel2, 0, (n2[SEG_THREADS] >> 3) + 1);
	}
	for (int t = SEG_THREADS - 1; t >= 0; t--) {
		long off = bits[t], o1 = n1[t], o2 = n2[t];
		seg_emit(ss, t, st.rle_buf, st.rle_128, select, zone,
		         [&](uint32_t code, int len) { seg_put_bits([&](int w, uint32_t v) { im.words[w] |= v; }, word0, off, code, len); off += len; },
		         [&](int bit) { if (!part) im.sel1[o1 >> 3] |= (uint8_t)(bit << (7 - (o1 & 7))); o1++; },
		         [&](int bit) { if (!part) im.sel2[o2 >> 3] |= (uint8_t)(bit << (7 - (o2 & 7))); o2++; });
	}
	if (!part) {
		h->size_data1 = word0 + nwords;
		h->wavelet_type = (select > 4 || b == 0) ? 4 : 0;
		h->select1 = (int)(n1[SEG_THREADS] >> 3) + 1;
		h->select2 = (int)(n2[SEG_THREADS] >> 3) + 1;
	} else h->size_data2 = word0 + nwords;
	word0 += nwords;
	pack_codebook(im, st, part, k);
	if (!part) s[262144] = saved;
	return 0;
}

// the k_groups_inplace schedule: every group computed from the plane as it was before the stage
template <typename F>
static void host_groups_inplace(const EncImg &im, int rows, int gpr, F f)
{
	std::vector<int16_t> before(im.proc - 4096, im.proc + 512 * 512 + 4096);
	const int16_t *B = before.data() + 4096;
	for (int r = rows - 1; r >= 0; r--)
		for (int g = gpr - 1; g >= 0; g--) {
			int o[8];
			if (f(B, r, g, o)) for (int x = 0; x < 8; x++) im.proc[r * 512 + g * 8 + x] = (int16_t)o[x];
		}
}

// the cell form of the dead-zone quantiser: the band is read only, im_jpeg gets the written cells
static void host_recons_quant_cells(const EncImg &im, int m1, int part)
{
	for (int r = 255; r >= 0; r--)
		for (int g = 31; g >= 0; g--) {
			int o[8];
			const int mask = y_recons_quant_cells(im.proc + r * 512, r, g, m1, part, o);
			for (int x = 0; x < 8; x++) if (mask >> x & 1) im.jpeg[r * 512 + g * 8 + x] = (int16_t)o[x];
		}
}

static void host_c_recons_cells(const EncImg &im, int m1, int comp)
{
	for (int r = 127; r >= 0; r--)
		for (int g = 15; g >= 0; g--) {
			int o[8];
			c_recons_cells(im.cproc + r * 256, r, g, m1, comp, o);
			for (int x = 0; x < 8; x++) im.cjpeg[r * 256 + g * 8 + x] = (int16_t)o[x];
		}
}

// the k_patterns schedule (enc_patterns.cuh): class masks, rows solved independently, fix-up walk, parallel apply
static void host_patterns(const EncImg &im, int kind)
{
	static PatMasks m;
	memset(&m, 0, sizeof m);
	for (int r = 0; r < 257; r++)
		for (int g = 0; g < 32; g++) {
			int v[8];
			uint32_t p8, n8;
			ld8(im.proc + r * 512 + g * 8, v);
			pat_class_bits8(v, p8, n8);
			((uint8_t *)m.pos[r])[g] = (uint8_t)p8;
			((uint8_t *)m.neg[r])[g] = (uint8_t)n8;
		}
	const uint64_t zero[4] = {0, 0, 0, 0};
	for (int r = pat_rows(kind) - 1; r >= 0; r--) pat_solve_row(m, kind, r, zero);
	pat_fixup(m, kind);
	if (getenv("HE_STATS")) { int nt = 0, nb = 0, nf = 0; uint64_t cl[4]; for (int r = 0; r < pat_rows(kind); r++) { for (int w = 0; w < 4; w++) { nt += __builtin_popcountll(m.ft[r][w]); nb += __builtin_popcountll(m.fb[r][w]); } if (r && pat_cleared_by(m, r - 1, cl)) nf++; } fprintf(stderr, "patterns kind %d: %d triples, %d blocks, %d rows redone\n", kind, nt, nb, nf); }
	for (int r = pat_rows(kind) - 1; r >= 0; r--)
		for (int j = 255; j >= 0; j--) {
			const bool t = m.ft[r][j >> 6] >> (j & 63) & 1, bk = m.fb[r][j >> 6] >> (j & 63) & 1;
			if (t || bk) pat_apply(im.proc, im.jpeg, kind, r, j, t, m.pos[r][j >> 6] >> (j & 63) & 1);
		}
}

// the CUDA wavefront schedule, run sequentially: step t lets row ri handle column t - skew*ri;
// rows of one step are visited bottom-up so that any same-step dependency shows up as a diff
template <typename Cell>
void host_wavefront(WfGeom g, Cell cell)
{
	std::vector<int> next(g.rows, 0);
	const int steps = g.cols + g.skew * (g.rows - 1);
	for (int t = 0; t < steps; t++)
		for (int ri = (getenv("HE_TOPDOWN") ? 0 : g.rows - 1); ri >= 0 && ri < g.rows; ri += (getenv("HE_TOPDOWN") ? 1 : -1)) {
			const int c = t - g.skew * ri;
			if (c >= 0 && c < g.cols && c == next[ri]) next[ri] = c + cell(g.r0 + ri, g.c0 + c);
		}
}

// parallel form of the LL2 part of offsetY_recons256, on the plane itself (PS = 512)
void host_recons_ll2(const EncImg &im, int q, int part)
{
	int16_t *P = im.proc, *J = im.jpeg;
	if (q > 17) for (int r = 127; r >= 0; r--) y_recons_ll2_tag_row(P, 512, r, part);
	if (!getenv("HE_WAVEFRONT")) {   // mask form (enc_ll2_masks.cuh), the schedule of k_recons_ll2
		static Ll2Masks m;
		if (getenv("HE_LL2_DEBUG")) {
			std::vector<int16_t> cp(P - 4096, P + 512 * 512 + 4096), jj(512 * 512 + 8192);
			int16_t *Q = cp.data() + 4096;
			std::vector<int16_t> before(cp);
			host_wavefront(wf_ll2_geom(), [&](int r, int j) { return y_recons_ll2_cell(Q, 512, jj.data() + 4096, q, part, r, j); });
			memset(&m, 0, sizeof m);
			for (int r = 127; r >= 0; r--) ll2_masks_row(P + r * 512, m.odd[r], m.tag[r], m.d2inc[r]);
			ll2_nudge_solve(m, q, part == 1);
			int shown = 0;
			if (part) for (int r = 0; r < 128 && shown < 6; r++) for (int j = 0; j < 128 && shown < 6; j++) {
				int16_t p2 = before[4096 + r * 512 + j], j2 = 0, t2 = 0;
				ll2_recons_apply_cell(&p2, &j2, &t2, m.d2inc[r][j >> 6] >> (j & 63) & 1, part);
				if (j2 != jj[4096 + r * 512 + j] || p2 != Q[r * 512 + j]) {
					shown++;
					fprintf(stderr, "ll2 part %d (r=%d,j=%d): J want %d got %d, P want %d got %d (before %d)\n", part, r, j, jj[4096 + r * 512 + j], j2, Q[r * 512 + j], p2, before[4096 + r * 512 + j]);
				}
			}
			for (int r = 0; r < 128 && shown < 6; r++) for (int j = 0; j < 128 && shown < 6; j++) {
				const int16_t *B = before.data() + 4096;
				int want = Q[r * 512 + j] - B[r * 512 + j]; if (B[r * 512 + j] > 10000 && part) want += 16000;
				int got = m.d2inc[r][j >> 6] >> (j & 63) & 1;
				if (want != got) {
					shown++;
					fprintf(stderr, "ll2 part %d (r=%d,j=%d): want inc %d got %d\n", part, r, j, want, got);
					for (int dr = -1; dr <= 3; dr++) { for (int dj = -3; dj <= 3; dj++) fprintf(stderr, "%6d", r + dr >= 0 ? B[(r + dr) * 512 + j + dj] : 0); fprintf(stderr, "\n"); }
				}
			}
		}
		memset(&m, 0, sizeof m);
		for (int r = 127; r >= 0; r--) ll2_masks_row(P + r * 512, m.odd[r], m.tag[r], m.d2inc[r]);
		ll2_nudge_solve(m, q, part == 1);
		for (int r = 127; r >= 0; r--)
			for (int j = 127; j >= 0; j--) ll2_recons_apply_cell(P + r * 512 + j, J + r * 512 + j, im.aux + r * 128 + j, m.d2inc[r][j >> 6] >> (j & 63) & 1, part);
	} else {
	host_wavefront(wf_ll2_geom(), [&](int r, int j) { return y_recons_ll2_cell(P, 512, J, q, part, r, j); });
	if (!part) for (int r = 127; r >= 0; r--) y_recons_ll2_tail_row(P, 512, J, im.aux, r);
	}
	if (!part) {
		if (q > 15)
			for (int k = im.hdr->highres_mem_len - 1; k >= 0; k--) {
				const int m = im.highres_mem[k];
				J[((m >> 7) << 9) + (m & 127)] = im.aux[m];
			}
	}
}

// parallel form of LL2 -> bytes + DPCM coding (enc_ll_par.cuh), on the plane itself
void host_ll2_code(const EncImg &im, int q)
{
	int16_t *P = im.proc;
	std::vector<int16_t> V(16384);
	std::vector<uint8_t> r4(128 * 32);
	std::vector<int> n4(128, 0);
	if (q > 17) for (int r = 127; r >= 0; r--) n4[r] = ll2_bytes_tag_row(P, 512, r, &r4[r * 32]);
	if (!getenv("HE_WAVEFRONT")) {
		static Ll2Masks m;
		memset(&m, 0, sizeof m);
		for (int r = 127; r >= 0; r--) ll2_masks_row(P + r * 512, m.odd[r], m.tag[r], m.d2inc[r]);
		ll2_nudge_solve(m, q, false);
		for (int r = 127; r >= 0; r--)
			for (int j = 127; j >= 0; j--) { V[r * 128 + j] = (int16_t)ll2_bytes_value(P[r * 512 + j], m.d2inc[r][j >> 6] >> (j & 63) & 1, q); P[r * 512 + j] = 0; }
	} else
	host_wavefront(wf_ll2_geom(), [&](int r, int j) { return ll2_bytes_cell(P, 512, V.data(), q, r, j); });
	for (int a = 16383; a >= 0; a--)
		if (!ll2_is_escape(V[a], a)) ll2_bytes_store(im, a, V[a]);
	int e = 0;
	for (int a = 0; a < 16384; a++)
		if (ll2_is_escape(V[a], a)) ll2_bytes_escape(im, a, V[a], e);
	im.hdr->exw_y_len = e;
	if (q > 17) {
		int n = 0;
		for (int r = 0; r < 128; r++)
			for (int k = 0; k < n4[r]; k++) im.res4[n++] = r4[r * 32 + k];
		im.hdr->res4_len = n;
	}
	ll_dpcm_luma_steps(im, im.tree1, q);
}

void host_e16(const EncImg &im, int q)
{
	memcpy(im.aux, im.proc, E16_SNAP_P_CELLS * sizeof(int16_t));
	memcpy(im.aux + E16_SNAP_L_OFF, im.ll1, 65536 * sizeof(int16_t));
	memset(im.aux + E16_SNAP_L_OFF + 65536, 0, 1024 * sizeof(int16_t));
	const bool mem = getenv("HE_E16_MEM") != nullptr;   // the memory-resident walk instead of the register window
	std::vector<uint8_t> lutv(E16_LUT_SIZE);
	for (int k = 0; k < E16_LUT_SIZE; k++) lutv[k] = (uint8_t)e16_lut_entry(k, res_setting_of(q));
	const uint8_t *lut = getenv("HE_E16_NOLUT") ? nullptr : lutv.data();
	for (int jj = 254; jj >= 0; jj--) {
		int j = getenv("HE_TOPDOWN") ? 254 - jj : jj;
		if (mem) y_e16_residual_col(im, q, j, im.aux, im.aux + E16_SNAP_L_OFF);
		else y_e16_residual_col_w(im, q, j, im.aux, im.aux + E16_SNAP_L_OFF, lut);
	}
	if (mem) y_e16_residual_col(im, q, 255, im.proc, im.ll1);
	else y_e16_residual_col_w(im, q, 255, im.proc, im.ll1, lut);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// decoder pipeline on the host (same stage functions as the CUDA path)
static void inv_rows(const int16_t *in, int in_stride, int16_t *out, int out_stride, int rows, int M, bool norm)
{
	for (int k = 0; k < rows; k++) {
		auto l = [&](int t) { return (int)in[k * in_stride + t]; };
		auto h = [&](int t) { return (int)in[k * in_stride + M + t]; };
		for (int t = 0; t < M; t++) {
			int ev, od;
			inverse_pair(l, h, t, M, norm, ev, od);
			out[k * out_stride + 2 * t] = (int16_t)ev;
			out[k * out_stride + 2 * t + 1] = (int16_t)od;
		}
	}
}

static void transpose_sq(const int16_t *in, int16_t *out, int stride, int N)
{
	for (int i = 0; i < N; i++)
		for (int j = 0; j < N; j++) out[i * stride + j] = in[j * stride + i];
}

static int host_decode(const uint8_t *blob, size_t len, uint8_t *rgb, uint8_t *yuv_out)
{
	DecDesc d;
	int rc = nhw_parse_header(blob, len, &d);
	if (rc) return rc;
	std::vector<int16_t> proc(262144 + 8192, 0), jpeg(262144 + 8192, 0), aux(262144 + 8192, 0), uv(131072 + 64, 0);
	std::vector<int16_t> cproc(65536 + 8192, 0), cjpeg(65536 + 8192, 0), caux(65536 + 8192, 0), tmp;
	std::vector<uint8_t> res_comp(24577 + 64, 0), yuv(3 * 262144), btmp(2048);
	std::vector<uint16_t> lists(8 * 65536), flags(65536), book(1024, 0), ltmp(65536 + 64);
	int32_t list_len[16] = {0};
	DecImg im;
	im.blob = blob; im.d = &d;
	im.proc = proc.data() + 4096; im.jpeg = jpeg.data() + 4096; im.aux = aux.data() + 4096;
	im.uvcoef = uv.data();
	im.cproc = cproc.data() + 4096; im.cjpeg = cjpeg.data() + 4096; im.caux = caux.data() + 4096;
	im.res_comp = res_comp.data();
	for (int k = 0; k < 8; k++) im.list[k] = lists.data() + k * 65536;
	im.list_len = list_len; im.flags = flags.data(); im.book = book.data(); im.yuv = yuv.data();
	std::vector<uint16_t> lut(NHW_LUT_WORDS);
	dec_build_lut(lut.data());
	im.lut = lut.data();

	rc = dec_ll_dpcm(im);
	if (rc) return rc;
	dec_build_book(blob + d.off_tree1, d.size_tree1, 3, -1, im.book, btmp.data());
	rc = dec_prefix_luma(im, im.proc, dec_build_actions(im.book, true));
	if (rc) return rc;
	for (int s = 127; s >= 0; s--) dec_y_descan_strip(im.proc, im.jpeg, s);
	dec_lists_image(im, ltmp.data());
	std::vector<uint32_t> hq0(NHW_CAP_HQ_LIST), hq1(NHW_CAP_HQ_LIST), hqtmp(NHW_CAP_HQ_LIST + 64);
	im.hq_list[0] = hq0.data(); im.hq_list[1] = hq1.data();
	dec_hq_lists_image(im, hqtmp.data());
	if (getenv("HE_SERIAL")) dec_y_markers_image(im);
	else {   // parallel form (dec_par.cuh), phases in the kernel's order
		int16_t *J = im.jpeg;
		std::vector<int> c1, c2, c3;
		for (int r = 0; r < 256; r++) for (int j = 0; j < 512; j++) if (J[r * 512 + j] > 1000) c1.push_back(r * 512 + j);
		for (int r = 256; r < 512; r++) for (int j = 0; j < 256; j++) if (J[r * 512 + j] > 1000) c2.push_back(r * 512 + j);
		for (int r = 256; r < 512; r++) for (int j = 256; j < 512; j++) if (J[r * 512 + j] > 1000) c3.push_back(r * 512 + j);
		for (int s : c1) dec_marker_apply(J, s, false, nullptr, nullptr);
		for (int s : c2) dec_marker_apply(J, s, true, nullptr, nullptr);
		std::vector<int16_t> S(J, J + 512 * 512);
		std::vector<uint32_t> W(2048, 0), A(2048, 0);
		for (int s : c3) dec_marker_apply(J, s, true, W.data(), A.data());
		int first = 1 << 30;
		if (d.quality < 23)
		for (int r = 511; r >= 256; r--) for (int j = 510; j > 256; j--) {
			const int s = r * 512 + j;
			if (dec_dense_qualifies(S.data(), A.data(), s) && s < first) first = s;
		}
		if (d.quality < 23)
		for (int r = 511; r >= 256; r--) for (int j = 510; j > 256; j--) {
			const int s = r * 512 + j, k = ((r - 256) << 8) + (j - 256);
			if (!dec_dense_qualifies(S.data(), A.data(), s)) continue;
			const int cnt = dec_dense_count(J, S.data(), s) + (s == first ? im.list_len[8] : 0);
			if (cnt >= 2 && !((W[k >> 5] >> (k & 31)) & 1u)) J[s] += S[s] > 0 ? 1 : -1;
		}
	}
	int exw = dec_y_ll_image(im);
	if (getenv("HE_SERIAL")) dec_y_shrink_image(im);
	else if (d.quality <= 16) { for (int r = 1; r < 255; r++) for (int j = 254; j >= 1; j--) dec_shrink_lowq_cell(im.jpeg, r, j); }
	else if (getenv("HE_WAVEFRONT")) host_wavefront(dwf_shrink_geom(), [&](int r, int j) { return dwf_shrink_cell(im.jpeg, r, j); });
	else
		for (int r = 255; r >= 0; r--)
			for (int g = 31; g >= 0; g--) {
				int o[8];
				if (shrink_cells8(im.jpeg, r, g, 9, o)) for (int x = 0; x < 8; x++) im.jpeg[r * 512 + g * 8 + x] = (int16_t)o[x];
			}
	inv_level(im.jpeg, im.proc, 512, 256, tmp);                 // LL1 reconstruction, natural orientation
	dec_y_addbacks_image(im);
	if (getenv("HE_SERIAL")) dec_y_edge_flags_image(im);
	else {
		host_wavefront(dwf_edge_geom(), [&](int r, int p) { return dwf_edge_cell(im.proc, r, p); });
		int n = 0;
		for (int r = 1; r < 255; r++)
			for (int j = 0; j < 256; j++) {
				const int s = r * 512 + j;
				if (im.proc[s] > 10000) { im.flags[n++] = (uint16_t)((r << 8) + j); im.proc[s] -= 16000; }
			}
		im.list_len[9] = n;
	}
	transpose_sq(im.proc, im.jpeg, 512, 256);
	inv_rows(im.jpeg, 512, im.proc, 512, 512, 256, false);      // wavelet_synthesis2: first half
	for (int k = dec_hq_addback_count(im) - 1; k >= 0; k--) {
		int pos, amount;
		if (dec_hq_addback(im, k, pos, amount) && pos >= 0 && pos < 512 * 512) im.proc[pos] = (int16_t)(im.proc[pos] + amount);
	}
	transpose_sq(im.proc, im.jpeg, 512, 512);
	dec_y_smooth_flags_image(im);
	inv_rows(im.jpeg, 512, im.proc, 512, 512, 256, true);       // wavelet_synthesis(...,3): second half
	for (int i = 0; i < 262144; i++) im.yuv[i] = dec_clip8(im.proc[i]);

	std::fill(book.begin(), book.end(), 0);
	dec_build_book(blob + d.off_tree2, d.size_tree2, 128, d.tree_end, im.book, btmp.data());
	rc = dec_prefix_chroma(im, im.uvcoef, dec_build_actions(im.book, false));
	if (rc) return rc;
	for (int v = 0; v < 2; v++) {
		for (int s = 31; s >= 0; s--) dec_c_descan_strip(im.uvcoef, im.cjpeg, s, v);
		exw = dec_c_ll_image(im, v, exw);
		inv_level(im.cjpeg, im.cproc, 256, 128, tmp);
		dec_c_markers_image(im);
		transpose_sq(im.cproc, im.cjpeg, 256, 128);
		inv_level(im.cjpeg, im.cproc, 256, 256, tmp);
		if (getenv("HE_SERIAL")) {
			dec_c_sharpen_image(im);
			for (int y = 511; y >= 0; y--) dec_c_upsample_row(im.cproc, im.yuv + (1 + v) * 262144, y);
		} else {
			const int thr = d.quality <= 14 ? 35 : 60;
			host_wavefront(dwf_sharpen_geom(), [&](int r, int j) { return dwf_sharpen_cell(im.cproc, thr, r, j); });
			for (int i = 65535; i >= 0; i--) dec_c_upsample_cell(im.cproc, im.yuv + (1 + v) * 262144, i >> 8, i & 255);
		}
	}
	const DecColor col = dec_color_of(d.quality);
	for (int i = 0; i < 262144; i++) dec_ycc_to_rgb(im.yuv[i], im.yuv[262144 + i], im.yuv[524288 + i], col, rgb + 3 * i);
	if (yuv_out) memcpy(yuv_out, im.yuv, 3 * 262144);
	return 0;
}

// ---- luma at q <= 16: the stage list of encode_image in its row / image forms (the forms the CUDA path runs for
// these settings), rows visited bottom-up where they are independent units
static int he_luma_lowq(const EncImg &im, int q, int ratio, std::vector<int16_t> &tmp, tap_fn tap)
{
	EncHdr *h = im.hdr;
	auto T = [&](const char *name, const void *p, size_t n) { if (tap) tap(name, p, n); };
	if (q > 6) {   // closed loop (nhw_encoder.c:141-283)
		for (int r = 255; r >= 0; r--) y_e6a_tag_row(im, r);
		T("y_e6a_ll1", im.ll1, 65536 * 2);
		y_recons_ll2_image(im, q, 1);
		for (int r = 255; r >= 0; r--) y_recons_quant_row(im, r, ratio, 1, q);
		T("y_rec1_jpeg", im.jpeg, 512 * 512 * 2);
		inv_level(im.jpeg, im.proc, 512, 256, tmp);
		T("y_syn1_proc", im.proc, 512 * 512 * 2);
		for (int r = 255; r >= 0; r--) y_e6c_apply_row(im, r);
		T("y_e6c_proc", im.proc, 512 * 512 * 2);
		T("y_e6c_ll1", im.ll1, 65536 * 2);
		for (int r = 255; r >= 0; r--) y_e6d_correct_row(im, r);
		T("y_e6d_jpeg", im.jpeg, 512 * 512 * 2);
		fwd_level(im.jpeg, 512, false, im.proc, 512, 256, tmp);
		T("y_dwt2b_proc", im.proc, 512 * 512 * 2);
	}
	if (q <= 11) for (int r = 255; r >= 128; r--) y_e7_kill_row(im, q, ratio, r);
	if (q < 13) {
		if (getenv("HE_SERIAL")) y_e8_smooth_image(im, q);
		else {   // the band staged apart from the plane, as the CUDA kernel runs it
			std::vector<int16_t> band(128 * 128);
			for (int r = 0; r < 128; r++) memcpy(&band[r * 128], im.proc + r * 512, 256);
			y_e8_smooth_band(band.data(), 128, im.proc, q);
			for (int r = 0; r < 128; r++) memcpy(im.proc + r * 512, &band[r * 128], 256);
		}
	}
	T("y_e8_proc", im.proc, 512 * 512 * 2);
	copy_region(im.ll2s, 256, im.proc, 512, 256);
	y_ll2_to_bytes_image(im, q);
	T("y_e11_tree1", im.tree1, 16384);
	T("y_e11_chres", im.ch_res, 16384);
	T("y_e11_exw", im.exw, h->exw_y_len);
	ll_dpcm_luma_image(im, q);
	T("y_e12_chres", im.llcode, h->y_res_comp);
	copy_region(im.proc, 512, im.ll2s, 256, 256);
	if (q > 12) {
		y_recons_ll2_image(im, q, 0);
		for (int r = 255; r >= 0; r--) y_recons_quant_row(im, r, ratio, 0, q);
		if (getenv("HE_SERIAL")) y_recons_shrink_image(im, q);
		else for (int r = 1; r < 255; r++) for (int j = 254; j >= 1; j--) y_recons_shrink_lowq_cell(im.jpeg, r, j);
		T("y_rec0_jpeg", im.jpeg, 512 * 512 * 2);
		inv_level(im.jpeg, im.proc, 512, 256, tmp);
		T("y_syn0_proc", im.proc, 512 * 512 * 2);
	}
	if (q == 16) { for (int r = 511; r >= 256; r--) y_e14_threshold_row(im, q, ratio, r); }
	else y_e14_lowq_image(im, q, ratio);
	T("y_e14_proc", im.proc, 512 * 512 * 2);
	T("y_e15_proc", im.proc, 512 * 512 * 2);
	if (q > 12) {
		host_e16(im, q);
		T("y_e16_proc", im.proc, 512 * 512 * 2);
		T("y_e16_ll1", im.ll1, 65536 * 2);
		y_e16b_classify_image(im, q);
		T("y_e16b_proc", im.proc, 512 * 512 * 2);
		T("y_e16b_ll1", im.ll1, 65536 * 2);
		y_e18_pack_list_image(im, 1);
		T("y_e18_ll1", im.ll1, 65536 * 2);
	}
	for (int r = 255; r >= 0; r--) y_e19_restore_row(im, r);
	y_e20_cleanup_image(im, q, ratio);
	T("y_e20_proc", im.proc, 512 * 512 * 2);
	y_offset_pairs_image(im);
	if (getenv("HE_SERIAL")) y_offset_quant_lowq_image(im, ratio);
	else {   // the CUDA schedule: every row dry for the three incoming states, one chain walk, rows committed bottom-up
		std::vector<int> next0(512), st_out(512 * 3), sp(512 * 3), in_state(512);
		for (int r = 0; r < 512; r++) next0[r] = r < 511 ? im.proc[(r + 1) * 512] : 0;
		for (int r = 511; r >= 0; r--)
			for (int t = 0; t < 3; t++) st_out[r * 3 + t] = y_offset_quant_lowq_row(im.proc + r * 512, nullptr, r, ratio, next0[r], t, sp[r * 3 + t]);
		int state = 0, spilled = 0;
		for (int r = 0; r < 512; r++) { in_state[r] = state; spilled |= sp[r * 3 + state]; state = st_out[r * 3 + state]; }
		if (spilled) y_offset_quant_lowq_image(im, ratio);
		else
			for (int r = 511; r >= 0; r--) { int dummy; y_offset_quant_lowq_row(im.proc + r * 512, im.proc + r * 512, r, ratio, next0[r], in_state[r], dummy); }
	}
	T("y_e21_proc", im.proc, 512 * 512 * 2);
	for (int s = 127; s >= 0; s--) y_scan_strip(im, s);
	T("y_e23_scan", im.scan, 262144);
	y_peephole_image(im);
	T("y_e24_scan", im.scan, 262144);
	return 0;
}

extern "C" {

int he_decode(const uint8_t *blob, long len, uint8_t *rgb, uint8_t *yuv_out) { return host_decode(blob, (size_t)len, rgb, yuv_out); }

// luma pre-sharpening at q <= 16, in place on a 512x512 plane
void he_pre_lowq(int16_t *y, int q)
{
	std::vector<int16_t> o(512 * 512 + 8192, 0), k(512 * 512 + 8192, 0), yy(512 * 512 + 8192, 0);
	std::vector<uint8_t> m(512 * 512 + 8192, 0);
	memcpy(yy.data() + 4096, y, 512 * 512 * 2);
	for (auto &v : k) v = 0x5555;   // the kernel plane is scratch: stale contents must not matter
	pre_low_image(yy.data() + 4096, o.data() + 4096, k.data() + 4096, m.data() + 4096, q);
	memcpy(y, yy.data() + 4096, 512 * 512 * 2);
}

void *he_new() { return make_work(); }
void he_free(void *h) { delete static_cast<Work *>(h); }

// Full encode of one image from the front-end's outputs: the pre-sharpened luma plane and
// the two 4:2:0 chroma byte planes.  tap (optional) receives named intermediate buffers.
// Returns the stream length or a negative status.
int he_encode(void *hw, const int16_t *y_pre, const uint8_t *u8, const uint8_t *v8, int q, uint8_t *out, tap_fn tap)
{
	Work *w = static_cast<Work *>(hw);
	std::fill(w->mem.begin(), w->mem.end(), 0);
	memset(&w->hdr, 0, sizeof w->hdr);
	const EncImg &im = w->im;
	EncHdr *h = im.hdr;
	h->quality = q;
	const int ratio = 8;
	std::vector<int16_t> tmp;
	auto T = [&](const char *name, const void *p, size_t n) { if (tap) tap(name, p, n); };
	if (q < 1 || q > 23) return -3;

	// ---------------- luma ----------------
	memcpy(im.jpeg, y_pre, 512 * 512 * 2);
	fwd_level(im.jpeg, 512, false, im.proc, 512, 512, tmp);
	if (q > 21)   // kept first pass, low half, transposed (tmp = R[y][k] after fwd_level)
		for (int k = 0; k < 256; k++)
			for (int y = 0; y < 512; y++) im.hq_qs[k * 512 + y] = tmp[y * 512 + k];
	for (int m = 0; m < 256; m++)
		for (int k = 0; k < 256; k++) im.ll1[m * 256 + k] = im.proc[k * 512 + m];
	fwd_level(im.proc, 512, true, im.proc, 512, 256, tmp);
	T("y_dwt2_proc", im.proc, 512 * 512 * 2);
	T("y_ll1", im.ll1, 65536 * 2);
	if (q <= 16) {
		int rc = he_luma_lowq(im, q, ratio, tmp, tap);
		if (rc) return rc;
	} else {
	if (getenv("HE_ROWFORM")) { for (int r = 255; r >= 0; r--) y_e6a_tag_row(im, r); }
	else for (int r = 255; r >= 0; r--) for (int g = 31; g >= 0; g--) y_e6a_tag_cells(im.proc, im.ll1 + r * 256, r, g);
	T("y_e6a_ll1", im.ll1, 65536 * 2);
	host_recons_ll2(im, q, 1);
	if (getenv("HE_WAVEFRONT")) {
		for (int reg = 0; reg < 2; reg++)
			host_wavefront(wf_recons_patterns_geom(reg), [&](int r, int j) { return wf_recons_patterns_cell(im, r, j); });
	} else host_patterns(im, 0);
	if (getenv("HE_ROWFORM")) { for (int r = 255; r >= 0; r--) y_recons_quant_row(im, r, ratio, 1); }
	else host_recons_quant_cells(im, ratio, 1);
	T("y_rec1_jpeg", im.jpeg, 512 * 512 * 2);
	inv_level(im.jpeg, im.proc, 512, 256, tmp);
	T("y_syn1_proc", im.proc, 512 * 512 * 2);
	if (getenv("HE_ROWFORM")) { for (int r = 255; r >= 0; r--) y_e6c_apply_row(im, r); }
	else for (int r = 255; r >= 0; r--) for (int g = 31; g >= 0; g--) y_e6c_apply_cells(im.proc, im.ll1 + r * 256, r, g);
	T("y_e6c_proc", im.proc, 512 * 512 * 2);
	T("y_e6c_ll1", im.ll1, 65536 * 2);
	if (getenv("HE_SERIAL")) { for (int r = 255; r >= 0; r--) y_e6d_correct_row(im, r); }
	else {   // cell-parallel form: every cell from the row's original differences
		std::vector<int16_t> sc(264);
		for (int r = 255; r >= 0; r--) {
			int16_t *P = im.proc + r * 512, *J = im.jpeg + r * 512;
			const int16_t *L = im.ll1 + r * 256;
			for (int j = -1; j <= 256; j++) sc[j + 2] = (int16_t)(P[j] - L[j]);
			if (getenv("HE_E6D_CELL")) {
				for (int j = 255; j >= 0; j--) {
					const int d = e6d_delta_at(&sc[2], j);
					J[j] = (int16_t)(L[j] + d);
					P[j] = (int16_t)(P[j] + d);
				}
			} else {
				for (int g = 31; g >= 0; g--) {
					int own[8], d[8];
					for (int x = 0; x < 8; x++) own[x] = sc[2 + g * 8 + x];
					e6d_delta_cells(&sc[2], g * 8, own, d);
					for (int x = 0; x < 8; x++) { J[g * 8 + x] = (int16_t)(L[g * 8 + x] + d[x]); P[g * 8 + x] = (int16_t)(P[g * 8 + x] + d[x]); }
				}
			}
		}
	}
	T("y_e6d_jpeg", im.jpeg, 512 * 512 * 2);
	fwd_level(im.jpeg, 512, false, im.proc, 512, 256, tmp);
	T("y_dwt2b_proc", im.proc, 512 * 512 * 2);
	copy_region(im.ll2s, 256, im.proc, 512, 256);
	if (getenv("HE_SERIAL")) y_ll2_to_bytes_image(im, q); else host_ll2_code(im, q);
	T("y_e11_tree1", im.tree1, 16384);
	T("y_e11_chres", im.ch_res, 16384);
	T("y_e11_exw", im.exw, h->exw_y_len);
	T("y_e11_res4", im.res4, h->res4_len);
	if (getenv("HE_SERIAL")) ll_dpcm_luma_image(im, q);
	T("y_e12_chres", im.llcode, h->y_res_comp);
	copy_region(im.proc, 512, im.ll2s, 256, 256);
	host_recons_ll2(im, q, 0);
	if (getenv("HE_WAVEFRONT")) {
		for (int reg = 0; reg < 2; reg++)
			host_wavefront(wf_recons_patterns_geom(reg), [&](int r, int j) { return wf_recons_patterns_cell(im, r, j); });
	} else host_patterns(im, 0);
	if (getenv("HE_ROWFORM")) {
		for (int r = 255; r >= 0; r--) y_recons_tag57_row(im, r);
		for (int r = 255; r >= 0; r--) y_recons_quant_row(im, r, ratio, 0);
	} else host_recons_quant_cells(im, ratio, 0);
	if (getenv("HE_WAVEFRONT")) host_wavefront(wf_shrink_geom(), [&](int r, int j) { return wf_shrink_cell(im, r, j); });
	else   // order-free form, run in place in the reverse of the reference's order
		for (int r = 255; r >= 0; r--)
			for (int g = 31; g >= 0; g--) {
				int o[8];
				if (shrink_cells8(im.jpeg, r, g, 8, o)) for (int x = 0; x < 8; x++) im.jpeg[r * 512 + g * 8 + x] = (int16_t)o[x];
			}
	T("y_rec0_jpeg", im.jpeg, 512 * 512 * 2);
	inv_level(im.jpeg, im.proc, 512, 256, tmp);
	T("y_syn0_proc", im.proc, 512 * 512 * 2);
	if (q > 21) {
		for (int r = 0; r < 256; r++)
			for (int j = 0; j < 256; j++) im.hq_fo[r * 256 + j] = im.proc[j * 512 + r];   // the reference copies its transposed work plane
		T("y_hq_fo0", im.hq_fo, 65536 * 2);
	}
	if (getenv("HE_ROWFORM")) {
		for (int r = 511; r >= 256; r--) y_e14_threshold_row(im, q, ratio, r);
		T("y_e14_proc", im.proc, 512 * 512 * 2);
		for (int r = 510; r >= 1; r--) y_e15_tags_row(im, r);
	} else host_groups_inplace(im, 512, 64, [&](const int16_t *B, int r, int g, int *o) { return y_e14_e15_cells(B + r * 512, r, g, q, ratio, o); });
	T("y_e15_proc", im.proc, 512 * 512 * 2);
	host_e16(im, q);
	T("y_e16_proc", im.proc, 512 * 512 * 2);
	T("y_e16_ll1", im.ll1, 65536 * 2);
	y_e16b_classify_image(im, q);
	T("y_e16b_proc", im.proc, 512 * 512 * 2);
	T("y_e16b_ll1", im.ll1, 65536 * 2);
	if (q > 21) { hq_e17_image(im); T("y_hq_fo1", im.hq_fo, 65536 * 2); }
	y_e18_pack_list_image(im, 1);
	if (q >= 19) y_e18_pack_list_image(im, 3);
	if (q >= 21) y_e18_pack_list_image(im, 5);
	T("y_e18_ll1", im.ll1, 65536 * 2);
	for (int r = 255; r >= 0; r--) y_e19_restore_row(im, r);
	if (getenv("HE_WAVEFRONT")) {
		for (int pass = 0; pass < 3; pass++)
			host_wavefront(wf_e20_geom(pass), [&](int r, int j) { return wf_e20_cell(im, q, ratio, pass, r, j); });
	} else {   // cell-parallel form: every final value from the plane as it was before the stage
		std::vector<int16_t> before(im.proc - 4096, im.proc + 512 * 512 + 4096);
		const int16_t *B = before.data() + 4096;
		for (int pass = 2; pass >= 0; pass--) {
			const E20Pass g = e20_pass(q, ratio, pass);
			for (int r = g.r1 - 1; r >= g.r0; r--) {
				if (getenv("HE_E20_CELL")) {
					for (int j = g.j1; j >= g.j0; j--) im.proc[r * 512 + j] = (int16_t)e20_final_cell(B + r * 512, 512, g, r, j);
					continue;
				}
				const int cb = pass == 1 ? 0 : 256;
				if (pass == 1) im.proc[r * 512 + 256] = (int16_t)e20_edge_cell(B + r * 512, 512, g, r);
				for (int gi = 31; gi >= 0; gi--) {
					int o[8];
					e20_cells8(B + (r - 1) * 512, B + r * 512, B + (r + 1) * 512, g, r, cb + gi * 8, o);
					for (int x = 0; x < 8; x++) {
						const int j = cb + gi * 8 + x;
						if (j >= g.j0 && j <= g.j1) im.proc[r * 512 + j] = (int16_t)o[x];
					}
				}
			}
		}
	}
	T("y_e20_proc", im.proc, 512 * 512 * 2);
	if (getenv("HE_ROWFORM")) { for (int r = 511; r >= 0; r--) y_offset_mult8_row(im, r); }
	else host_groups_inplace(im, 512, 64, [&](const int16_t *B, int r, int g, int *o) { return y_offset_mult8_cells(B, r, g, o); });
	if (getenv("HE_WAVEFRONT")) host_wavefront(wf_offset_patterns_geom(), [&](int r, int j) { return wf_offset_patterns_cell(im, r, j); });
	else host_patterns(im, 1);
	if (getenv("HE_ROWFORM")) { for (int r = 255; r >= 0; r--) y_offset_pairs57_row(im, r); }
	else host_groups_inplace(im, 256, 32, [&](const int16_t *B, int r, int g, int *o) { return y_offset_pairs57_cells(B, r, g, o); });
	if (getenv("HE_ROWFORM")) {
		std::vector<int> next0(512);
		for (int r = 0; r < 512; r++) next0[r] = r < 511 ? im.proc[(r + 1) * 512] : 0;
		for (int r = 511; r >= 0; r--) y_offset_quant_row(im, ratio, r, next0[r]);
		T("y_e21_proc", im.proc, 512 * 512 * 2);
		for (int s = 127; s >= 0; s--) y_scan_strip(im, s);
	} else {   // pointwise form (enc_point.cuh), as the fused quantise+scan kernel runs it
		const int16_t *P = im.proc;
		for (int i = 512 * 512 - 1; i >= 0; i--) {
			const int row = i >> 9, col = i & 511;
			const int op1 = i + 1 < 512 * 512 ? P[i + 1] : 0;
			im.scan[y_scan_pos(row, col)] = (uint8_t)y_quant_byte(col ? P[i - 1] : 0, P[i], op1, col >= 1, col < 511, ratio);
		}
	}
	T("y_e23_scan", im.scan, 262144);
	if (q > 21) {
		auto byte_at = [&](int c) { return (int)im.scan[y_scan_pos(c >> 8, 256 + (c & 255))]; };
		for (int c = 65535; c >= 0; c--) im.hq_band[c] = (int16_t)hq_band_cell(byte_at, c);
		T("y_hq_band", im.hq_band, 65536 * 2);
		for (int r = 255; r >= 0; r--) hq_tag_row(im, q, r);
		int rc = hq_lists_image(im, q);
		if (rc) return rc;
	}
	if (getenv("HE_SERIAL")) y_peephole_image(im); else host_peephole(im);
	T("y_e24_scan", im.scan, 262144);
	}

	// ---------------- chroma ----------------
	for (int v = 0; v < 2; v++) {
		const uint8_t *src = v ? v8 : u8;
		const char *pfx = v ? "v" : "u";
		char name[64];
		auto TN = [&](const char *sfx, const void *p, size_t n) { snprintf(name, sizeof name, "%s_%s", pfx, sfx); T(name, p, n); };
		if (q <= 14) { for (int i = 65535; i >= 0; i--) im.cjpeg[i] = (int16_t)c_pre_uv_cell(src, q, i >> 8, i & 255); }
		else for (int i = 0; i < 65536; i++) im.cjpeg[i] = src[i];
		fwd_level(im.cjpeg, 256, false, im.cproc, 256, 256, tmp);
		TN("dwt1_proc", im.cproc, 65536 * 2);
		for (int m = 0; m < 128; m++)
			for (int k = 0; k < 128; k++) im.cll1[m * 128 + k] = im.cproc[k * 256 + m];
		if (q <= 16) for (int i = 65535; i >= 0; i--) im.cproc[i] = (int16_t)c_threshold_cell(im.cproc[i], ratio, i >> 8, i & 255);
		fwd_level(im.cproc, 256, true, im.cproc, 256, 128, tmp);
		TN("dwt2_proc", im.cproc, 65536 * 2);
		TN("ll1", im.cll1, 16384 * 2);
		if (getenv("HE_ROWFORM") || q <= 16) {
			for (int r = 63; r >= 0; r--) c_recons_ll_row(im, r, 1, q);
			for (int r = 127; r >= 0; r--) c_recons_quant_row(im, r, ratio, 1);
		} else host_c_recons_cells(im, ratio, 1);
		TN("rec1_jpeg", im.cjpeg, 65536 * 2);
		inv_level(im.cjpeg, im.cproc, 256, 128, tmp);
		TN("syn1_proc", im.cproc, 65536 * 2);
		if (getenv("HE_ROWFORM")) { for (int r = 127; r >= 0; r--) c_correct_row(im, r, v); }
		else for (int r = 127; r >= 0; r--) for (int g = 15; g >= 0; g--) {
			int o[8];
			c_correct_cells(im.cproc + r * 256, im.cll1 + r * 128, g, v, o);
			for (int x = 0; x < 8; x++) im.cjpeg[r * 256 + g * 8 + x] = (int16_t)o[x];
		}
		TN("corr_jpeg", im.cjpeg, 65536 * 2);
		fwd_level(im.cjpeg, 256, false, im.cproc, 256, 128, tmp);
		TN("dwt2b_proc", im.cproc, 65536 * 2);
		copy_region(im.cll2s, 128, im.cproc, 256, 128);
		if (getenv("HE_ROWFORM") || q <= 16) {
			for (int r = 63; r >= 0; r--) c_recons_ll_row(im, r, 0, q);
			for (int r = 127; r >= 0; r--) c_recons_quant_row(im, r, ratio, 0);
		} else host_c_recons_cells(im, ratio, 0);
		inv_level(im.cjpeg, im.cproc, 256, 128, tmp);
		TN("syn0_proc", im.cproc, 65536 * 2);
		if (getenv("HE_ROWFORM")) { for (int r = 127; r >= 0; r--) c_residual_tags_row(im, q, r); }
		else {
			std::vector<int> tg(128 * 128, 0);
			for (int r = 127; r >= 0; r--) for (int g = 15; g >= 0; g--) c_residual_tags_cells(im.cproc, im.cll1, q, r, g, &tg[r * 128 + g * 8]);
			for (int r = 127; r >= 0; r--) for (int j = 127; j >= 0; j--) if (tg[r * 128 + j]) c_drop_tag(im.cproc, r * 256 + j, tg[r * 128 + j]);
		}
		copy_region(im.cproc, 256, im.cll2s, 128, 128);
		if (q <= 11) c_ll_smooth_image(im);
		TN("tags_proc", im.cproc, 65536 * 2);
		int e = c_ll_to_bytes_image(im, v);
		if (v) h->exw_v_len = e; else h->exw_u_len = e;
		if (q > 15) c_ll_bit1_plane(im, v);
		if (getenv("HE_ROWFORM")) {
			std::vector<int> next0(256);
			for (int r = 0; r < 256; r++) next0[r] = r < 255 ? im.cproc[(r + 1) * 256] : 0;
			for (int r = 255; r >= 0; r--) c_offset_quant_row(im, ratio, r, next0[r]);
			TN("quant_proc", im.cproc, 65536 * 2);
			for (int s = 31; s >= 0; s--) c_scan_strip(im, s, v);
		} else {
			const int16_t *P = im.cproc;
			for (int i = 65535; i >= 0; i--) {
				const int row = i >> 8, col = i & 255;
				int run = 0;
				bool pre = false;
				if (c_pairable(P[i])) { while (run < col && c_pairable(P[i - 1 - run])) run++; }
				else if (P[i] == 7) {
					while (run < col && P[i - 1 - run] == 7) run++;
					pre = run < col && c_bumps_next(P[i - 1 - run]);
				}
				const int op1 = i + 1 < 65536 ? P[i + 1] : 0;
				im.scan[c_scan_pos(row, col, v)] = (uint8_t)c_quant_byte(P[i], op1, run, pre, col < 255, ratio);
			}
		}
	}
	T("uv_scan", im.scan + 262144, 131072);
	ll_dpcm_chroma_image(im);
	T("llcode", im.llcode, h->end_ch_res);
	int rc;
	if (getenv("HE_SERIAL")) {
		int a = 0;
		rc = packet_stream_image(im, 0, a);
		if (rc) return rc;
		a++;
		rc = packet_stream_image(im, 1, a);
		if (rc) return rc;
	} else {
		int word0 = 0;
		rc = host_entropy(im, 0, word0);
		if (rc) return rc;
		rc = host_entropy(im, 1, word0);
		if (rc) return rc;
	}
	T("hdr", h, sizeof *h);
	return write_stream_image(im, out);
}

}  // extern "C"
