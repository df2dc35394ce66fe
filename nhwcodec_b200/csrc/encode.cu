// encode.cu -- batch encoder driver: sequences the CUDA kernels of the full .nhw encode for a
// chunk of images resident in device memory.  Replaces encode_image + write_compressed_file
// (encoder/nhw_encoder.c:103-2878, 3100-3220) for a whole batch; every image is independent.
//
// Kernel granularity of this first complete version:
//   * front end (front.cu)            : data-parallel, shared-memory tiled
//   * inverse / forward transforms    : data-parallel
//   * *_row stages                    : one thread per (image,row)
//   * *_image stages                  : one thread per image (raster-order dependencies)
// The per-image serial kernels are latency-bound and are the next thing to parallelise
// (wavefronts / finite-state scans); they are correct and batch-parallel as they stand.
#include <stdlib.h>
#include "nhw_ctx.h"
#include "nhw_dev.cuh"
#include "enc_seg.cuh"
#include "enc_ll2_masks.cuh"
#include "enc_lowq.cuh"
#include "enc_batch.cuh"
#include "../../include/nhw_cuda.h"

namespace {

__device__ __forceinline__ EncImg make_img(const EncBatch &b, int i, int comp)
{
	EncImg im;
	im.proc = b.y_proc + (size_t)i * NHW_Y_SLOT;
	im.jpeg = b.y_jpeg + (size_t)i * NHW_Y_SLOT;
	im.aux = b.y_aux + (size_t)i * NHW_Y_SLOT;
	im.ll1 = b.y_ll1 + (size_t)i * NHW_C_SLOT;
	im.ll2s = b.y_ll2s + (size_t)i * NHW_C_SLOT;
	const size_t p = (size_t)i * 2 + comp;
	im.cproc = b.c_proc + p * NHW_C_SLOT;
	im.cjpeg = b.c_jpeg + p * NHW_C_SLOT;
	im.caux = b.c_aux + p * NHW_C_SLOT;
	im.cll1 = b.c_ll1 + p * NHW_Q_SLOT;
	im.cll2s = b.c_ll2s + p * NHW_Q_SLOT;
	uint8_t *bytes = b.bytes + (size_t)i * ENC_BYTES_SLOT;
	im.scan = bytes + OFF_SCAN;
	im.tree1 = bytes + OFF_TREE1;
	im.ch_res = bytes + OFF_CHRES;
	im.llcode = bytes + OFF_LLCODE;
	im.exw = bytes + OFF_EXW;
	im.exw_uv = bytes + OFF_EXWUV;
	im.res1 = bytes + OFF_RES1;
	im.res1_bit = bytes + OFF_RES1_BIT;
	im.res1_word = bytes + OFF_RES1_WORD;
	im.res3 = bytes + OFF_RES3;
	im.res3_bit = bytes + OFF_RES3_BIT;
	im.res3_word = bytes + OFF_RES3_WORD;
	im.res4 = bytes + OFF_RES4;
	im.res5 = bytes + OFF_RES5;
	im.res5_bit = bytes + OFF_RES5_BIT;
	im.res5_word = bytes + OFF_RES5_WORD;
	im.res6 = bytes + OFF_RES6;
	im.res6_bit = bytes + OFF_RES6_BIT;
	im.res6_word = bytes + OFF_RES6_WORD;
	im.char_res1 = reinterpret_cast<uint16_t *>(bytes + OFF_CHARRES1);
	im.qsetting3 = reinterpret_cast<uint32_t *>(bytes + OFF_QSET3);
	im.hq_qs = b.y_hq + (size_t)i * NHW_Y_SLOT;
	im.hq_fo = im.hq_qs + 131072;
	im.hq_band = im.hq_qs + 196608;
	im.hq_tag = reinterpret_cast<uint8_t *>(im.jpeg);   // im_jpeg is dead once the second reconstruction is done
	im.tmp1 = bytes + OFF_TMP1;
	im.tmp2 = bytes + OFF_TMP2;
	im.tmp3 = bytes + OFF_TMP3;
	im.highres_mem = reinterpret_cast<uint16_t *>(bytes + OFF_HRMEM);
	im.highres_word = bytes + OFF_HRWORD;
	im.res_uv64 = bytes + OFF_UV64;
	im.sel1 = bytes + OFF_SEL1;
	im.sel2 = bytes + OFF_SEL2;
	im.codebook1 = bytes + OFF_BOOK1;
	im.codebook2 = bytes + OFF_BOOK2;
	im.words = reinterpret_cast<uint32_t *>(bytes + OFF_WORDS);
	im.pack_scratch = bytes + OFF_PACK;
	im.hdr = b.hdr + i;
	return im;
}

// ---- generic launch shapes ----
template <typename F>
__global__ void k_image(EncBatch b, int n, F f)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f(make_img(b, i, 0), i);
}

template <typename F>
__global__ void k_plane(EncBatch b, int n2, F f)   // thread per chroma plane (2 per image)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n2) f(make_img(b, p >> 1, p & 1), p & 1);
}

// "thread = row" stages walk rows that lie 512 B or 1 KB apart, so every load of a warp touches 32 lines and the
// rest of each line is only used if it is still in L1 when the thread gets there.  The grid is therefore kept at
// ROWS_CTAS_PER_SM resident CTAs per SM (persistent, looping over the work) so that the lines a warp has in
// flight stay in L1 instead of being evicted by thousands of other row walkers.
#define ROWS_CTAS_PER_SM 24
template <typename F>
__global__ void __launch_bounds__(64) k_rows(EncBatch b, int rows, int n, F f)
{
	const int nblk = (rows + 63) >> 6;
	for (int item = blockIdx.x; item < n * nblk; item += gridDim.x) {
		const int img = item / nblk, r = (item % nblk) * 64 + threadIdx.x;
		if (r < rows) f(make_img(b, img, 0), r);
	}
}

template <typename F>
__global__ void __launch_bounds__(64) k_plane_rows(EncBatch b, int rows, int n2, F f)
{
	const int nblk = (rows + 63) >> 6;
	for (int item = blockIdx.x; item < n2 * nblk; item += gridDim.x) {
		const int pl = item / nblk, r = (item % nblk) * 64 + threadIdx.x;
		if (r < rows) f(make_img(b, pl >> 1, pl & 1), r, pl & 1);
	}
}

// ---- cell-group executors (enc_cells.cuh): one thread per group of 8 cells, 256 threads = whole rows of a CTA.
// k_groups: f does its own loads and stores (stages that never read what they write).
// k_groups_inplace: f only reads and returns the group's final values; the CTA synchronises, then stores, so a
// group never sees another group's result (rows never straddle CTAs, and these stages only look along their row).
// A CTA takes GROUP_ITER consecutive blocks of 256 groups.  Measured: 4 blocks per CTA is slower for the stages with
// data-dependent replays (level-2 quantiser 1.6 -> 2.7 ms) and no faster for the others, so it stays at 1.
#define GROUP_ITER 1
template <typename F>
__global__ void __launch_bounds__(256) k_groups(EncBatch b, int gshift, F f)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	#pragma unroll
	for (int it = 0; it < GROUP_ITER; it++) {
		const int idx = (blockIdx.x * GROUP_ITER + it) * 256 + threadIdx.x;
		f(im, idx >> gshift, idx & ((1 << gshift) - 1));
	}
}
template <typename F>
__global__ void __launch_bounds__(256) k_groups_inplace(EncBatch b, int gshift, F f)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	for (int it = 0; it < GROUP_ITER; it++) {
		const int idx = (blockIdx.x * GROUP_ITER + it) * 256 + threadIdx.x;
		const int r = idx >> gshift, g = idx & ((1 << gshift) - 1);
		int o[8];
		const bool changed = f(im, r, g, o);
		__syncthreads();   // every group of these rows has read; the next iteration works on other rows
		if (changed) st8(im.proc + r * YW + g * 8, o);
	}
}
// dead-zone quantiser of the level-2 detail bands into im_jpeg (enc_cells.cuh: y_recons_quant_cells); the band is only read
__device__ __forceinline__ void recons_quant_group(const EncImg &im, int r, int g, int m1, int part)
{
	int o[8];
	const int mask = y_recons_quant_cells(im.proc + r * YW, r, g, m1, part, o);
	int16_t *J = im.jpeg + r * YW + g * 8;
	if (mask == 0xff) st8(J, o);
	else for (int x = 0; x < 8; x++) if (mask >> x & 1) J[x] = (int16_t)o[x];
}
template <typename F>
void run_groups(nhw_ctx *c, const char *label, const EncBatch &b, int n, int rows, int gshift, F f)
{
	NHW_LAUNCH_L(c, label, k_groups, dim3((rows << gshift) / 256 / GROUP_ITER, n), 256, 0, b, gshift, f);
}
template <typename F>
void run_groups_inplace(nhw_ctx *c, const char *label, const EncBatch &b, int n, int rows, int gshift, F f)
{
	NHW_LAUNCH_L(c, label, k_groups_inplace, dim3((rows << gshift) / 256 / GROUP_ITER, n), 256, 0, b, gshift, f);
}

template <typename F>
__global__ void __launch_bounds__(256) k_plane_groups(EncBatch b, int gshift, F f)
{
	const EncImg im = make_img(b, blockIdx.y >> 1, blockIdx.y & 1);
	#pragma unroll
	for (int it = 0; it < GROUP_ITER; it++) {
		const int idx = (blockIdx.x * GROUP_ITER + it) * 256 + threadIdx.x;
		f(im, idx >> gshift, idx & ((1 << gshift) - 1), (int)(blockIdx.y & 1));
	}
}
template <typename F>
void run_plane_groups(nhw_ctx *c, const char *label, const EncBatch &b, int n, int rows, int gshift, F f)
{
	NHW_LAUNCH_L(c, label, k_plane_groups, dim3((rows << gshift) / 256 / GROUP_ITER, 2 * n), 256, 0, b, gshift, f);
}
template <typename F>
void run_image(nhw_ctx *c, const char *label, const EncBatch &b, int n, F f)
{
	NHW_LAUNCH_L(c, label, k_image, (n + 31) / 32, 32, 0, b, n, f);
}
template <typename F>
void run_plane(nhw_ctx *c, const char *label, const EncBatch &b, int n, F f)
{
	NHW_LAUNCH_L(c, label, k_plane, (2 * n + 31) / 32, 32, 0, b, 2 * n, f);
}
template <typename F>
void run_rows(nhw_ctx *c, const char *label, const EncBatch &b, int n, int rows, F f)
{
	const int work = ((rows + 63) / 64) * n, cap = c->tune.rows_grid_cap;
	NHW_LAUNCH_L(c, label, k_rows, work < cap ? work : cap, 64, 0, b, rows, n, f);
}
template <typename F>
void run_plane_rows(nhw_ctx *c, const char *label, const EncBatch &b, int n, int rows, F f)
{
	const int work = ((rows + 63) / 64) * 2 * n, cap = c->tune.rows_grid_cap;
	NHW_LAUNCH_L(c, label, k_plane_rows, work < cap ? work : cap, 64, 0, b, rows, 2 * n, f);
}

// ---- wavefront executor: one CTA per image, thread = row of the stage's region; see enc_par.cuh
template <typename Cell>
__global__ void __launch_bounds__(256) k_wavefront(EncBatch b, WfGeom g, Cell cell)
{
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int ri = threadIdx.x;
	const int steps = g.cols + g.skew * (g.rows - 1);
	int next = 0;
	for (int t = 0; t < steps; t++) {
		const int c = t - g.skew * ri;
		if (ri < g.rows && c >= 0 && c < g.cols && c == next) next = c + cell(im, g.r0 + ri, g.c0 + c);
		__syncthreads();
	}
}

template <typename Cell>
void run_wavefront(nhw_ctx *c, const char *label, const EncBatch &b, int n, WfGeom g, Cell cell)
{
	NHW_LAUNCH_L(c, label, k_wavefront, n, 256, 0, b, g, cell);
}

// ---- residual coding (E16): columns 0..254 concurrently against a snapshot, then column 255
__global__ void __launch_bounds__(256) k_e16_residual(EncBatch b, int q)
{
	__shared__ uint8_t lut[E16_LUT_SIZE];   // arm of the rule chain per (res, a, b), enc_y2.cuh: e16_arm
	for (int k = threadIdx.x; k < E16_LUT_SIZE; k += 256) lut[k] = (uint8_t)e16_lut_entry(k, res_setting_of(q));
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int4 *ps = reinterpret_cast<const int4 *>(im.proc);
	int4 *pd = reinterpret_cast<int4 *>(im.aux);
	// neighbour columns are read as they were before the stage: left half of rows 0..257 (columns <= 255 are all a
	// column walk looks at next to it) and the LL1 copy
	for (int i = threadIdx.x; i < 258 * 32; i += 256) { const int k = (i >> 5) * 64 + (i & 31); pd[k] = ps[k]; }
	const int4 *ls = reinterpret_cast<const int4 *>(im.ll1);
	int4 *ld = reinterpret_cast<int4 *>(im.aux + E16_SNAP_L_OFF);
	for (int i = threadIdx.x; i < E16_SNAP_L_CELLS / 8; i += 256) ld[i] = i < 65536 / 8 ? ls[i] : make_int4(0, 0, 0, 0);
	__syncthreads();
	const int j = threadIdx.x;
	// all 256 columns in lockstep, one row per barrier; column 255 (the one that reads live data: column 0's codes
	// a few rows further down) runs E16_LAG rows behind instead of in a serial pass of its own
	y_e16_residual_col_t<true>(im, q, j, j < 255 ? im.aux : im.proc, j < 255 ? im.aux + E16_SNAP_L_OFF : im.ll1, lut, j < 255 ? 0 : E16_LAG);
}

__global__ void __launch_bounds__(256) k_e16b_classify(EncBatch b, int q)
{
	const EncImg im = make_img(b, blockIdx.x, 0);
	int w1 = 0, w3 = 0, w5 = 0;   // the reference only uses these as malloc sizes
	y_e16b_classify_col_w(im, q, threadIdx.x, w1, w3, w5);
}

// the LL2 band staged in shared memory: 128 rows x 132 columns (4 columns of look-ahead)
#define LL2_PS 132
#define LL2_SMEM_BYTES (128 * LL2_PS * 2)

// LL2 part of offsetY_recons256 in its parallel form (enc_par.cuh): band in shared memory,
// thread = LL2 row; tagging rows -> wavefront (skew 3) -> second-call tail.
__global__ void __launch_bounds__(128) k_recons_ll2_wave(EncBatch b, int q, int part)
{
	extern __shared__ __align__(16) int16_t sP[];
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int t = threadIdx.x;
	for (int idx = t; idx < 128 * (LL2_PS / 2); idx += 128) {
		const int r = idx / (LL2_PS / 2), c = idx % (LL2_PS / 2);
		reinterpret_cast<uint32_t *>(sP + r * LL2_PS)[c] = reinterpret_cast<const uint32_t *>(im.proc + r * YW)[c];
	}
	__syncthreads();
	// quad tagging (rows independent), masks of the tagged row, one thread solves the parity nudges on the masks
	// (enc_ll2_masks.cuh), then every cell gets its final value and its im_jpeg sample
	__shared__ Ll2Masks m;
	if (q > 17) y_recons_ll2_tag_row(sP, LL2_PS, t, part);
	ll2_masks_row(sP + t * LL2_PS, m.odd[t], m.tag[t], m.d2inc[t]);
	if (t < LL2M_ROWS - 128) {
		m.odd[128 + t][0] = m.odd[128 + t][1] = m.odd[128 + t][2] = 0;
		m.tag[128 + t][0] = m.tag[128 + t][1] = 0;
	}
	__syncthreads();
	if (t == 0) ll2_nudge_solve(m, q, part == 1);
	__syncthreads();
	for (int r = 0; r < 128; r++)
		ll2_recons_apply_cell(sP + r * LL2_PS + t, im.jpeg + r * YW + t, im.aux + r * 128 + t, (int)(m.d2inc[r][t >> 6] >> (t & 63) & 1), part);
	if (!part) {
		__syncthreads();
		if (q > 15) {
			const int n = im.hdr->highres_mem_len;
			for (int k = t; k < n; k += 128) {
				const int m = im.highres_mem[k];
				im.jpeg[((m >> 7) << 9) + (m & 127)] = im.aux[m];
			}
		}
	}
	__syncthreads();
	for (int idx = t; idx < 128 * 64; idx += 128) {
		const int r = idx >> 6, c = idx & 63;
		reinterpret_cast<uint32_t *>(im.proc + r * YW)[c] = reinterpret_cast<const uint32_t *>(sP + r * LL2_PS)[c];
	}
}

// LL2 -> bytes + DPCM coding in parallel form (enc_ll_par.cuh), one CTA of 128 threads per image.
// shared memory: band copy (later: output offsets), sample values (later: step links), res4 rows.
#define LL2_CODE_SMEM (LL2_SMEM_BYTES + 16384 * 2 + 128 * 32 + 128 * 4 + 64)   // band (later: bytes + chain marks), samples (later: steps), res4 rows, counters
// chain marks and step links are walked one 128-position segment per thread: the segments are laid out 33 / 65 words
// apart so that the lanes of a warp hit different shared-memory banks
#define LL2_VI(i) ((i) + (((i) >> 7) << 2))
#define LL2_II(i) ((i) + (((i) >> 7) << 1))
__global__ void __launch_bounds__(128) k_ll2_code(EncBatch b, int q)
{
	extern __shared__ __align__(16) int16_t sP[];
	int16_t *V = sP + 128 * LL2_PS;
	uint8_t *r4 = reinterpret_cast<uint8_t *>(V + 16384);
	int *n4 = reinterpret_cast<int *>(r4 + 128 * 32);
	int *cnt = n4 + 128;            // [0] a8, [1] y16
	const EncImg im = make_img(b, blockIdx.x, 0);
	EncHdr *h = im.hdr;
	const int t = threadIdx.x;
	for (int idx = t; idx < 128 * (LL2_PS / 2); idx += 128) {
		const int r = idx / (LL2_PS / 2), c = idx % (LL2_PS / 2);
		reinterpret_cast<uint32_t *>(sP + r * LL2_PS)[c] = reinterpret_cast<const uint32_t *>(im.proc + r * YW)[c];
	}
	if (t < 2) cnt[t] = 0;
	__syncthreads();
	n4[t] = q > 17 ? ll2_bytes_tag_row(sP, LL2_PS, t, r4 + t * 32) : 0;
	__syncthreads();
	{   // parity nudges on row masks (enc_ll2_masks.cuh); the masks borrow the sample array, which is filled last
		Ll2Masks &m = *reinterpret_cast<Ll2Masks *>(V);
		ll2_masks_row(sP + t * LL2_PS, m.odd[t], m.tag[t], m.d2inc[t]);
		if (t < LL2M_ROWS - 128) {
			m.odd[128 + t][0] = m.odd[128 + t][1] = m.odd[128 + t][2] = 0;
			m.tag[128 + t][0] = m.tag[128 + t][1] = 0;
		}
		__syncthreads();
		if (t == 0) ll2_nudge_solve(m, q, false);
		__syncthreads();
		uint64_t mine[2] = {0, 0};   // bit r = "cell (r, t) gets +1"
		for (int r = 0; r < 128; r++) mine[r >> 6] |= (m.d2inc[r][t >> 6] >> (t & 63) & 1) << (r & 63);
		__syncthreads();
		for (int r = 0; r < 128; r++) V[r * 128 + t] = (int16_t)ll2_bytes_value(sP[r * LL2_PS + t], (int)(mine[r >> 6] >> (r & 63) & 1), q);
		__syncthreads();
	}
	// ---- bytes.  A cell whose value does not fit a byte ("escape") repeats the byte on its left and goes to
	// the exw_Y list; lists are in raster order: each thread owns 128 consecutive cells, counts, CTA scan, writes.
	uint8_t *sx = reinterpret_cast<uint8_t *>(sP);            // the band copy is dead: tree1 bytes [0, 16384 + 128)
	uint8_t *vis = sx + 16384 + 256;                          // chain marks, one byte per position
	int *tot = cnt + 2;                                       // [0] escapes, [1] code bytes, [2] raw samples
	__shared__ int seg_cnt[3][129];
	__shared__ int seg_exit[128], seg_merge[128];
	// pass 1, thread = column: bytes of every cell (coalesced), escapes marked per row by warp ballots
	__shared__ uint32_t escw[128][4];
	for (int k = 0; k < 128; k++) {
		const int a = 128 * k + t;
		const int v = V[a];
		const bool esc = ll2_is_escape(v, a);
		const int cl = v > 255 ? 255 : v < 0 ? 0 : v;
		sx[a] = (uint8_t)(cl & 254);
		if (!esc) { im.tree1[a] = (uint8_t)(cl & 254); im.ch_res[a] = (uint8_t)cl; }
		const uint32_t m = __ballot_sync(0xffffffffu, esc);
		if ((t & 31) == 0) escw[k][t >> 5] = m;
	}
	__syncthreads();
	seg_cnt[0][t] = __popc(escw[t][0]) + __popc(escw[t][1]) + __popc(escw[t][2]) + __popc(escw[t][3]);   // thread = row from here on
	__syncthreads();
	if (t == 0) {
		int run = 0;
		for (int k = 0; k < 128; k++) { const int v = seg_cnt[0][k]; seg_cnt[0][k] = run; run += v; }
		tot[0] = run;
		h->exw_y_len = 3 * run;
		if (q > 17) {
			int n = 0;
			for (int r = 0; r < 128; r++)
				for (int k = 0; k < n4[r]; k++) im.res4[n++] = r4[r * 32 + k];
			h->res4_len = n;
		}
	}
	__syncthreads();
	{   // pass 2: the escapes of row t in raster order
		int e = 3 * seg_cnt[0][t];
		for (int w = 0; w < 4; w++)
			for (uint32_t m = escw[t][w]; m; m &= m - 1) {
				const int a = 128 * t + 32 * w + __ffs(m) - 1;
				const int v = V[a];
				im.exw[e++] = (uint8_t)(a >> 7);
				if (v > 255) { im.exw[e++] = (uint8_t)((a & 127) + 128); const int y = v - 255; im.exw[e++] = (uint8_t)(y > 255 ? 255 : y); }
				else { im.exw[e++] = (uint8_t)(a & 127); im.exw[e++] = (uint8_t)(v < -255 ? 255 : -v); }
				int p = a - 1;
				while (ll2_is_escape(V[p], p)) p--;               // position 0 never is one
				int u = V[p];
				u = u > 255 ? 255 : u < 0 ? 0 : u;
				sx[a] = (uint8_t)(u & 254);                        // (predecessors are looked up in V, not in sx)
				im.tree1[a] = (uint8_t)(u & 254);
				im.ch_res[a] = (uint8_t)(u & 254);
			}
		if (t == 0) for (int k = 0; k < 256; k++) sx[16384 + k] = 0;   // tree1[16384..] is still zero while luma is coded
	}
	__syncthreads();
	// ---- DPCM coder over the 16384 bytes (enc_ll_par.cuh).  Every position gets its step (where the coder would
	// go next from there, how many bytes it emits); the positions the coder really visits are the orbit of 1.
	// Each thread walks the orbit of its segment's first position (speculation), one thread then stitches the
	// segments together: the true chain enters a segment somewhere, runs until it meets the speculative
	// chain (from there on they coincide) or leaves the segment.  Offsets are a CTA scan over the segments.
	const uint8_t *x = sx;
	uint16_t *info = reinterpret_cast<uint16_t *>(V);      // (next - i) << 2 | (nbytes - 1) << 1 | raw
	{
		int a8 = 0, y16 = 0;
		for (int i = t + 1; i < 16384; i += 128)
			if (x[i] == x[i - 1] && (i == 1 || x[i - 1] != x[i - 2])) ll_stats_run(x, i, 16384, a8, y16);
		if (a8) atomicAdd(&cnt[0], a8);
		if (y16) atomicAdd(&cnt[1], y16);
	}
	__syncthreads();
	const int mode = cnt[1] > 299 ? 2 : (cnt[0] + cnt[1] > 179 ? 1 : 0);
	for (int i = t; i < 16384; i += 128) {
		vis[LL2_VI(i)] = 0;
		if (i >= 1) {
			const LlStep s = ll_dpcm_step(x, i, mode, q);
			info[LL2_II(i)] = (uint16_t)(((s.next - i) << 2) | ((s.nbytes - 1) << 1) | s.raw);
		}
	}
	__syncthreads();
	{
		int i = t ? 128 * t : 1;
		const int end = 128 * t + 128;
		while (i < end) { vis[LL2_VI(i)] = 1; i += info[LL2_II(i)] >> 2; }
		seg_exit[t] = i;
	}
	__syncthreads();
	if (t == 0) {
		int p = seg_exit[0];
		seg_merge[0] = 0;
		for (int k = 1; k < 128; k++) {
			const int end = 128 * k + 128;
			int i = p;
			while (i < end && vis[LL2_VI(i)] != 1) { vis[LL2_VI(i)] = 2; i += info[LL2_II(i)] >> 2; }
			if (i < end) { seg_merge[k] = i; p = seg_exit[k]; }     // met the speculative chain: it is the true one from here
			else { seg_merge[k] = end; p = i; }                     // never met it inside this segment
		}
	}
	__syncthreads();
	{
		const int m = seg_merge[t];
		int nb = 0, nr = 0;
		for (int i = 128 * t; i < 128 * t + 128; i++) {
			if (i < m && vis[LL2_VI(i)] == 1) vis[LL2_VI(i)] = 0;
			if (vis[LL2_VI(i)]) { const int inf = info[LL2_II(i)]; nb += 1 + ((inf >> 1) & 1); nr += inf & 1; }
		}
		seg_cnt[1][t] = nb;
		seg_cnt[2][t] = nr;
	}
	__syncthreads();
	if (t < 2) {
		int *v = seg_cnt[1 + t];
		int run = 0;
		for (int k = 0; k < 128; k++) { const int c = v[k]; v[k] = run; run += c; }
		tot[1 + t] = run;
	}
	__syncthreads();
	{
		int off = 1 + seg_cnt[1][t], nm = seg_cnt[2][t];
		for (int i = 128 * t; i < 128 * t + 128; i++) {
			if (!vis[LL2_VI(i)]) continue;
			const LlStep s = ll_dpcm_step(x, i, mode, q);
			im.llcode[off++] = s.b[0];
			if (s.nbytes == 2) im.llcode[off++] = s.b[1];
			if (s.raw) { im.highres_word[nm] = im.ch_res[i]; im.highres_mem[nm++] = (uint16_t)i; }
		}
	}
	if (t == 0) {
		im.llcode[0] = x[0];
		h->highres_comp_len = tot[2];
		h->highres_mem_len = tot[2];
		h->res_low = mode;
		h->y_res_comp = 1 + tot[1];
	}
}

// ---- res1/res3/res5 side-channel lists: rows collected in parallel (count, CTA prefix, write),
// the short list post-processing by one thread
// One coalesced sweep over LL1 collects all three lists (y_e18_classify): warp = row, lane = 8 columns; rows are
// counted, a CTA prefix gives every row its place in each list, a second sweep writes entries and the rewritten
// cells.  The serial tails of the lists (pruning, packing) are 3 x n independent single-thread jobs: they run in
// their own launch, one job per warp, so that thousands of them are resident at once (k_e18_tails).
#define E18_PART 21800        // scratch entries per list; longer lists (never seen) take the one-list-at-a-time path
__global__ void __launch_bounds__(256, 4) k_e18_lists(EncBatch b, int q)
{
	__shared__ int cnt[3][257];
	__shared__ int too_long;
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	auto sweep = [&](bool write) {
		for (int row = warp; row < 256; row += 8) {
			int16_t *L = im.ll1 + row * 256;
			const uint4 raw = reinterpret_cast<const uint4 *>(L)[lane];
			const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
			int v[8], n[3] = {0, 0, 0};
			uint32_t mw[8];   // membership bits (24..26) and the three word values (one byte each) of a cell
#pragma unroll
			for (int t = 0; t < 8; t++) {
				const int orig = (int16_t)(wv[t >> 1] >> ((t & 1) * 16));
				const int j = 8 * lane + t;
				int mem = 0, w[3] = {0, 0, 0};
				if (j < 254) v[t] = y_e18_classify(orig, q, mem, w);
				else v[t] = write ? 0 : orig;                       // the stage clears columns 254, 255
				mw[t] = ((uint32_t)mem << 24) | ((uint32_t)(w[0] & 255)) | ((uint32_t)(w[1] & 255) << 8) | ((uint32_t)(w[2] & 255) << 16);
#pragma unroll
				for (int k = 0; k < 3; k++) n[k] += (mem >> k) & 1;
			}
			int off[3];
			for (int k = 0; k < 3; k++) {
				int inc = n[k];
				for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
				off[k] = inc - n[k];
				if (!write && lane == 31) cnt[k][row] = inc;
			}
			if (write) {
				for (int k = 0; k < 3; k++) {
					const int e0 = cnt[k][row];
					uint8_t *pos = im.tmp1 + k * E18_PART + e0 + row, *wrd = im.tmp3 + k * E18_PART + e0;
					int o = off[k];
#pragma unroll
					for (int t = 0; t < 8; t++)
						if ((mw[t] >> (24 + k)) & 1u) { pos[o] = (uint8_t)(8 * lane + t); wrd[o] = (uint8_t)(mw[t] >> (8 * k)); o++; }
					if (lane == 31) pos[o] = 254;     // end-of-row marker after the row's last entry
				}
				reinterpret_cast<uint4 *>(L)[lane] =
				    make_uint4((uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16), (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16),
				               (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16), (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16));
			}
		}
	};
	sweep(false);
	__syncthreads();
	if (threadIdx.x < 3) {
		int *c = cnt[threadIdx.x];
		int run = 0;
		for (int r = 0; r < 256; r++) { const int x = c[r]; c[r] = run; run += x; }
		c[256] = run;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		too_long = (cnt[0][256] + 300 > E18_PART || cnt[1][256] + 300 > E18_PART || cnt[2][256] + 300 > E18_PART) ? 1 : 0;
		for (int k = 0; k < 3; k++) im.hdr->pad[k] = too_long ? -1 : cnt[k][256];   // entries per list, for k_e18_tails
	}
	__syncthreads();
	if (too_long) {
		// one list at a time with the whole scratch (the original schedule); thread = row
		const int row = threadIdx.x;
		for (int which = 1; which <= 5; which += 2) {
			if ((which == 3 && q < 19) || (which == 5 && q < 21)) continue;
			__syncthreads();
			cnt[0][row] = y_e18_collect_row(im, which, row, nullptr, nullptr);
			__syncthreads();
			if (row == 0) {
				int run = 0;
				for (int r = 0; r < 256; r++) { const int x = cnt[0][r]; cnt[0][r] = run; run += x; }
				cnt[0][256] = run;
			}
			__syncthreads();
			const int e0 = cnt[0][row];
			y_e18_collect_row(im, which, row, im.tmp1 + e0 + row, im.tmp3 + e0);
			__syncthreads();
			if (row == 0) y_e18_finish_list_image(im, which, cnt[0][256] + 256, cnt[0][256]);
		}
		return;
	}
	sweep(true);
}

// list tails: block (k, image), one thread
// The list tail (enc_y2.cuh: y_e18_finish_list_image states the steps) by one warp, 32 entries per step:
//   marker pruning   : pointwise on the collected list -> ballot compaction
//   LSB plane        : compaction of the real positions' low bits, then 8 per byte
//   pair merging     : "merge with the next entry, then skip it" = every other cell of each run of mergeable cursors
//                      (carry trick on the ballot word, the state carried across words is one bit)
//   word plane       : pointwise
__device__ __forceinline__ uint64_t alt64(uint64_t e)   // every other bit of each run of ones, from the run's lowest bit
{
	const uint64_t EVEN = 0x5555555555555555ull;
	const uint64_t es = e & ~(e << 1) & EVEN, x = e + es;
	return (e & ~x & EVEN) | (e & x & ~EVEN);
}
__device__ void e18_finish_list_warp(const EncImg &im, int which, int count, int e, int lane)
{
	EncHdr *h = im.hdr;
	uint8_t *pos = im.tmp1, *A = im.tmp2, *wrd = im.tmp3;
	uint8_t *out, *out_bit, *out_word;
	if (which == 1) { out = im.res1; out_bit = im.res1_bit; out_word = im.res1_word; }
	else if (which == 3) { out = im.res3; out_bit = im.res3_bit; out_word = im.res3_word; }
	else { out = im.res5; out_bit = im.res5_bit; out_word = im.res5_word; }
	const uint32_t lt = (1u << lane) - 1u;
	if (lane < 8) wrd[e + lane] = 0;   // the reference reads up to 7 entries past the end
	// ---- drop end-of-row markers the decoder can infer from a decreasing position: pos[0, count) -> A[0, len)
	int len = 0;
	for (int base = 0; base < count; base += 32) {
		const int i = base + lane;
		bool keep = false;
		int v = 0;
		if (i < count) {
			v = pos[i];
			keep = true;
			if (i >= 1 && i <= count - 2 && v == 254) {
				const int l = pos[i - 1], r = pos[i + 1];
				if (l != 254 && r != 254 && l > r) keep = false;
			}
		}
		const uint32_t m = __ballot_sync(0xffffffffu, keep);
		if (keep) A[len + __popc(m & lt)] = (uint8_t)v;
		len += __popc(m);
	}
	__syncwarp();
	// ---- LSB plane of the real positions: compacted into pos[] (free now), then 8 per byte, MSB first
	int np = 0;
	for (int base = 0; base < len; base += 32) {
		const int i = base + lane;
		const int v = i < len ? (int)A[i] : 254;
		const uint32_t m = __ballot_sync(0xffffffffu, v != 254);
		if (v != 254) pos[np + __popc(m & lt)] = (uint8_t)(v & 1);
		np += __popc(m);
	}
	if (lane < 8) pos[np + lane] = 0;
	__syncwarp();
	const int bit_len = (np >> 3) + 1;
	for (int o = lane; o < bit_len; o += 32) {
		int bb = 0;
		for (int k = 0; k < 8; k++) bb |= (pos[8 * o + k] & 1) << (7 - k);
		out_bit[o] = (uint8_t)bb;
	}
	// ---- halve the positions and merge (small delta, small delta) pairs into one byte >= 128; cursors 1 .. len-2
	int olen = 1;
	if (lane == 0) out[0] = (uint8_t)(A[0] >> 1);
	uint32_t carry = 0;   // the cursor before this word merged (and swallows this word's first cursor)
	for (int base = 1; base < len - 1; base += 32) {
		const int i = base + lane;
		const bool in = i < len - 1;
		int cur = 0, d1 = -1, d2 = -1;
		if (in) { cur = A[i] >> 1; d1 = cur - (A[i - 1] >> 1); d2 = (A[i + 1] >> 1) - cur; }
		const bool mg = in && d1 >= 0 && d1 < 8 && d2 >= 0 && d2 < 16;
		const uint32_t M = __ballot_sync(0xffffffffu, mg);
		const uint64_t f64 = alt64(((uint64_t)M << 1) | carry);
		const uint32_t F = (uint32_t)(f64 >> 1);                  // cursors that merge
		const uint32_t skipped = (F << 1) | carry;                 // cursors swallowed by the merge before them
		const uint32_t inm = __ballot_sync(0xffffffffu, in);
		const uint32_t vis = inm & ~skipped;
		if ((vis >> lane) & 1u) out[olen + __popc(vis & lt)] = (uint8_t)(((F >> lane) & 1u) ? 128 + (d1 << 4) + d2 : cur);
		olen += __popc(vis);
		carry = F >> 31;
	}
	// ---- word plane
	const int groups = (e >> 3) + 1;
	for (int g = lane; g < groups; g += 32) {
		const uint8_t *w = wrd + 8 * g;
		if (which == 3) {
			out_word[2 * g] = (uint8_t)(((w[0] & 3) << 6) | ((w[1] & 3) << 4) | ((w[2] & 3) << 2) | (w[3] & 3));
			out_word[2 * g + 1] = (uint8_t)(((w[4] & 3) << 6) | ((w[5] & 3) << 4) | ((w[6] & 3) << 2) | (w[7] & 3));
		} else {
			int bb = 0;
			for (int k = 0; k < 8; k++) bb |= (w[k] & 1) << (7 - k);
			out_word[g] = (uint8_t)bb;
		}
	}
	if (lane == 0) {
		const int wbytes = which == 3 ? 2 * groups : groups;
		if (which == 1) { h->res1_len = olen; h->res1_bit_len = bit_len; h->res1_word_len = wbytes; }
		else if (which == 3) { h->res3_len = olen; h->res3_bit_len = bit_len; h->res3_word_len = wbytes; }
		else { h->res5_len = olen; h->res5_bit_len = bit_len; h->res5_word_len = wbytes; }
	}
}

// list tails: block (k, image), one warp
__global__ void __launch_bounds__(32) k_e18_tails(EncBatch b, int q)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	const int k = blockIdx.x, which = 1 + 2 * k, e = im.hdr->pad[k];
	if (e < 0 || (which == 3 && q < 19) || (which == 5 && q < 21)) return;
	EncImg lm = im;
	lm.tmp1 = im.tmp1 + k * E18_PART;
	lm.tmp2 = im.tmp2 + k * E18_PART;
	lm.tmp3 = im.tmp3 + k * E18_PART;
	e18_finish_list_warp(lm, which, e + 256, e, threadIdx.x);
}

// ---- pattern substitutions of the level-2 region on 256-bit row masks (enc_patterns.cuh).  One CTA per image:
// class bits of rows 0..256 (a warp per row, a lane per 8 cells), every row solved by its own thread as if the row
// above fired no block, one thread walks down the rows and redoes those below a row with blocks, then the fired
// events are applied, one thread per 32 columns of a row.
__global__ void __launch_bounds__(256) k_patterns(EncBatch b, int kind)
{
	__shared__ PatMasks m;
	const EncImg im = make_img(b, blockIdx.x, 0);
	int16_t *P = im.proc, *J = im.jpeg;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint8_t *posb = reinterpret_cast<uint8_t *>(m.pos), *negb = reinterpret_cast<uint8_t *>(m.neg);
	for (int r = warp; r < 257; r += 8) {
		int v[8];
		uint32_t p8, n8;
		ld8(P + r * YW + lane * 8, v);
		pat_class_bits8(v, p8, n8);
		posb[r * 32 + lane] = (uint8_t)p8;
		negb[r * 32 + lane] = (uint8_t)n8;
	}
	__syncthreads();
	const int rows = pat_rows(kind);
	if (tid < rows) {
		const uint64_t zero[4] = {0, 0, 0, 0};
		pat_solve_row(m, kind, tid, zero);
	}
	__syncthreads();
	if (tid == 0) pat_fixup(m, kind);
	__syncthreads();
	const uint32_t *ft = reinterpret_cast<const uint32_t *>(m.ft), *fb = reinterpret_cast<const uint32_t *>(m.fb);
	const uint32_t *pos = reinterpret_cast<const uint32_t *>(m.pos);
	for (int item = tid; item < rows * 8; item += 256) {
		const int r = item >> 3, w = item & 7;
		const uint32_t p = pos[r * 8 + w];
		for (uint32_t f = ft[r * 8 + w]; f; f &= f - 1) { const int bit = __ffs(f) - 1; pat_apply(P, J, kind, r, w * 32 + bit, true, p >> bit & 1); }
		for (uint32_t f = fb[r * 8 + w]; f; f &= f - 1) { const int bit = __ffs(f) - 1; pat_apply(P, J, kind, r, w * 32 + bit, false, p >> bit & 1); }
	}
}

// ---- chroma residual tags (enc_cells.cuh: c_residual_tags_cells): 16 rows of a plane per CTA, a thread decides for 8
// cells on the plane as it is, the tags are dropped after the barrier
__global__ void __launch_bounds__(256) k_c_residual_tags(EncBatch b, int q)
{
	const EncImg im = make_img(b, blockIdx.y >> 1, blockIdx.y & 1);
	const int idx = blockIdx.x * 256 + threadIdx.x, r = idx >> 4, g = idx & 15;
	int tags[8];
	c_residual_tags_cells(im.cproc, im.cll1, q, r, g, tags);
	__syncthreads();
	for (int k = 0; k < 8; k++) if (tags[k]) c_drop_tag(im.cproc, r * CW + g * 8 + k, tags[k]);
}

// ---- chroma LL (64 x 64) -> tree1 bytes + exw escapes + bit-1 planes (enc_c.cuh: c_ll_to_bytes_image,
// c_ll_bit1_plane), one warp per plane, two planes (U, V) per CTA.  A sample that does not fit a byte repeats the
// byte on its left (through any run of such samples) and goes to the escape list in raster order: a lane owns 128
// consecutive samples, counts its escapes, the warp scans, lanes write.  The band is zeroed on the way.
#define CLV(a) ((a) + (((a) >> 7) << 1))
#define CLB(a) ((a) + (((a) >> 7) << 2))
__global__ void __launch_bounds__(64) k_c_ll_quant(EncBatch b, int q)
{
	// a lane walks 128 consecutive samples: lanes are laid out 65 / 33 words apart (bank-conflict-free)
	__shared__ __align__(16) int16_t sv[2][4096 + 64];
	__shared__ __align__(16) uint8_t sb[2][4096 + 128], sf[2][4096 + 128];
	const int v = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const EncImg im = make_img(b, blockIdx.x, v);
	int16_t *P = im.cproc;
	int16_t *V = sv[v];
	uint8_t *B = sb[v];
	for (int r = 0; r < 64; r++) {   // a row of 64 samples = 32 words
		uint32_t *row = reinterpret_cast<uint32_t *>(P + r * CW);
		{ const uint32_t w = row[lane]; const int a = r * 64 + 2 * lane; V[CLV(a)] = (int16_t)(w & 0xffff); V[CLV(a + 1)] = (int16_t)(w >> 16); }
		row[lane] = 0;
	}
	__syncwarp();
	const int a0 = lane * 128;
	int ne = 0;
	for (int a = a0; a < a0 + 128; a++) {
		int x = V[CLV(a)];
		const bool esc = (x > 255 || x < 0) && a > 0;
		ne += esc ? 1 : 0;
		x = x > 255 ? 255 : x < 0 ? 0 : x;
		B[CLB(a)] = (uint8_t)(x & 254);
	}
	int off = ne;
	for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, off, d); if (lane >= d) off += o; }
	const int total = __shfl_sync(0xffffffffu, off, 31);
	off -= ne;
	__syncwarp();
	uint8_t *exw = im.exw_uv + (v ? 16384 : 0) + 3 * off;
	uint8_t *F = sf[v];
	for (int a = a0; a < a0 + 128; a++) {
		const int x = V[CLV(a)];
		uint8_t byte = B[CLB(a)];
		if ((x > 255 || x < 0) && a > 0) {
			*exw++ = (uint8_t)(a >> 6);
			if (x > 255) { *exw++ = (uint8_t)((a & 63) + 128); const int y = x - 255; *exw++ = (uint8_t)(y > 255 ? 255 : y); }
			else { *exw++ = (uint8_t)(a & 63); *exw++ = (uint8_t)(x < -255 ? 255 : -x); }
			int p = a - 1;
			while (p > 0 && (V[CLV(p)] > 255 || V[CLV(p)] < 0)) p--;     // sample 0 never is an escape
			byte = B[CLB(p)];
		}
		F[CLB(a)] = byte;
	}
	__syncwarp();
	uint32_t *t = reinterpret_cast<uint32_t *>(im.tree1 + (v ? 20480 : 16384));
	for (int k = lane; k < 1024; k += 32) t[k] = *reinterpret_cast<const uint32_t *>(F + CLB(4 * k));
	if (q > 15) {   // bit 1 of 8 consecutive bytes, MSB first (res_U_64 / res_V_64)
		uint8_t *o = im.res_uv64 + (v ? 512 : 0);
		for (int k = lane; k < 512; k += 32) {
			int pk = 0;
			for (int i = 0; i < 8; i++) pk |= ((F[CLB(8 * k + i)] >> 1) & 1) << (7 - i);
			o[k] = (uint8_t)pk;
		}
	}
	if (lane == 0) { if (v) im.hdr->exw_v_len = 3 * total; else im.hdr->exw_u_len = 3 * total; }
}

// ---- E6c (enc_cells.cuh: y_e6c_apply_cells states the rule): un-tag LL1, push +-1 into the trial reconstruction at
// the transposed position.  A CTA takes a 32 x 32 tile of LL1: tags are read (and cleared) row-wise, the +-1 go
// through shared memory and are applied column-wise, so that a warp's 32 targets lie in one 128-byte span.
__global__ void __launch_bounds__(256) k_e6c_apply(EncBatch b)
{
	__shared__ int8_t dt[32][33];
	const EncImg im = make_img(b, blockIdx.y, 0);
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	int16_t *P = im.proc;
	for (int tile = blockIdx.x * 4; tile < blockIdx.x * 4 + 4; tile++) {
		const int r0 = (tile >> 3) * 32, j0 = (tile & 7) * 32;
		#pragma unroll
		for (int i = 0; i < 4; i++) {
			const int rr = ty + 8 * i;
			int16_t *L = im.ll1 + (r0 + rr) * 256 + j0 + tx;
			const int l = *L;
			int d = 0;
			if (l > 14000) { *L = (int16_t)(l - 16000); d = 1; }
			else if (l > 10000) { *L = (int16_t)(l - 12000); d = -1; }
			dt[rr][tx] = (int8_t)d;
		}
		__syncthreads();
		#pragma unroll
		for (int i = 0; i < 4; i++) {
			const int jj = ty + 8 * i, d = dt[tx][jj], r = r0 + tx, j = j0 + jj;
			if (!d) continue;
			if (r < 128 && j >= 128) P[2 * r + ((j - 128) << 10) + YW] += d;
			else if (r >= 128 && j < 128) P[2 * (r - 128) + (j << 10) + 1] += d;
			else if (r >= 128 && j >= 128) P[2 * (r - 128) + ((j - 128) << 10) + YW + 1] += d;
		}
		__syncthreads();
	}
}

// ---- E6d: LL1 correction (enc_cells.cuh: e6d_delta_cells), one thread per 8 cells; 8 rows per CTA
__global__ void __launch_bounds__(256) k_e6d_correct(EncBatch b)
{
	__shared__ __align__(16) int16_t sc[8][272];   // differences of a row, columns -1 .. 256 at index 7 .. 264
	const EncImg im = make_img(b, blockIdx.y, 0);
	const int k = threadIdx.x >> 5, g = threadIdx.x & 31, j0 = g * 8;
	const int r = blockIdx.x * 8 + k;
	int16_t *P = im.proc + r * YW, *J = im.jpeg + r * YW;
	const int16_t *L = im.ll1 + r * 256;
	int p[8], l[8], d[8], own[8];
	ld8(P + j0, p);
	ld8(L + j0, l);
	#pragma unroll
	for (int x = 0; x < 8; x++) own[x] = (int16_t)(p[x] - l[x]);
	st8(&sc[k][8 + j0], own);
	if (g == 0) sc[k][7] = (int16_t)(P[-1] - L[-1]);
	if (g == 31) sc[k][264] = (int16_t)(P[256] - L[256]);
	__syncthreads();
	e6d_delta_cells(&sc[k][8], j0, own, d);
	#pragma unroll
	for (int x = 0; x < 8; x++) { l[x] += d[x]; p[x] += d[x]; }
	st8(J + j0, l);
	st8(P + j0, p);
}

// ---- E20: clean-up of the three level-1 bands (enc_cells.cuh: e20_cells8).  One CTA per (image, pass) walks its
// region top-down in bands of 8 rows staged in shared memory (two buffers in turn); a thread owns 8 cells of a row.
// Every final value is computed from the rows as they were before the stage: the row above a band is copied from
// the previous band's buffer, the row below has not been touched yet; results go straight back to the plane.
// Column 256 of rows 256..510 is written by pass 1 (it receives from column 255) and only read by pass 2, whose
// test on it has the same outcome either way: each pass stores only its own columns.
#define E20_ROWS 8
#define E20_TS 272   // 256 columns, column 256 (pass 1) and the two look-ahead cells behind the last group
__global__ void __launch_bounds__(256) k_e20_bands(EncBatch b, int q, int ratio)
{
	__shared__ __align__(16) int16_t tile[2][E20_ROWS + 2][E20_TS];
	const EncImg im = make_img(b, blockIdx.y, 0);
	const E20Pass g = e20_pass(q, ratio, blockIdx.x);
	const int cb = g.pass == 1 ? 0 : 256, tid = threadIdx.x, k = tid >> 5, c = (tid & 31) * 8;
	int16_t *P = im.proc;
	int buf = 0;
	for (int r0 = g.r0; r0 < g.r1; r0 += E20_ROWS, buf ^= 1) {
		int16_t (*T)[E20_TS] = tile[buf];
		{   // rows r0-1 .. r0+8; the cells right of column 255 of the block only matter for pass 1 (column 256)
			const int r = r0 + k;
			if (r <= 511) *reinterpret_cast<uint4 *>(&T[1 + k][c]) = *reinterpret_cast<const uint4 *>(P + r * YW + cb + c);
			if (k == 0) {
				const int r2 = r0 + E20_ROWS;
				if (r2 <= 511) *reinterpret_cast<uint4 *>(&T[1 + E20_ROWS][c]) = *reinterpret_cast<const uint4 *>(P + r2 * YW + cb + c);
			} else if (k == 1) {
				if (r0 == g.r0) *reinterpret_cast<uint4 *>(&T[0][c]) = *reinterpret_cast<const uint4 *>(P + (r0 - 1) * YW + cb + c);
				else *reinterpret_cast<uint4 *>(&T[0][c]) = *reinterpret_cast<const uint4 *>(&tile[buf ^ 1][E20_ROWS][c]);
			} else if (k == 2 && tid - 64 < E20_ROWS + 2) {
				const int kk = tid - 64, r2 = r0 - 1 + kk;   // tile row kk
				int e0 = 0, e1 = 0;
				if (g.pass == 1 && r2 <= 511) { e0 = P[r2 * YW + 256]; e1 = P[r2 * YW + 257]; }
				T[kk][256] = (int16_t)e0;
				T[kk][257] = (int16_t)e1;
			}
		}
		__syncthreads();
		const int r = r0 + k;
		if (r < g.r1) {
			const int16_t *row = &T[1 + k][0] - cb;
			int o[8];
			e20_cells8(row - E20_TS, row, row + E20_TS, g, r, cb + c, o);
			int16_t *dst = P + r * YW + cb + c;
			if (c == 0 && cb) { for (int x = 1; x < 8; x++) dst[x] = (int16_t)o[x]; }   // column 256 is not this pass's
			else st8(dst, o);
			if (g.pass == 1 && c == 248) P[r * YW + 256] = (int16_t)e20_edge_cell(row, E20_TS, g, r);
		}
		// the next band loads into the other buffer; this one is read again (its last row) during that load
	}
}

// ---- E19: restore the level-2 region from the resIII snapshot (y_e19_restore_row), one thread per 8 cells, 32 rows per CTA
__global__ void __launch_bounds__(256) k_e19_restore(EncBatch b)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	#pragma unroll
	for (int it = 0; it < 4; it++) {
		const int idx = (blockIdx.x * 4 + it) * 256 + threadIdx.x, r = idx >> 5, j = (idx & 31) * 8;
		int v[8];
		ld8(im.ll2s + r * 256 + j, v);
		if (r < 128 && j < 128) { for (int k = 0; k < 8; k++) if (v[k] <= 8000) v[k] = 0; }
		st8(im.proc + r * YW + j, v);
	}
}

// ---- offsetY loop 4 + serpentine scan, pointwise (enc_point.cuh): 16 rows per CTA, 8 cells per thread
// and step; the bytes go through shared memory so that they leave in 64-byte runs of the scan order.
#define YQ_STRIP 68
__global__ void __launch_bounds__(256) k_y_quant_scan(EncBatch b, int m1)
{
	// a strip's 64 bytes are 68 bytes apart: the lanes of a warp store to every other strip, which at 64 bytes apart
	// was one bank for the whole warp (32-way conflict on both stores, 97 % of the store wavefronts)
	__shared__ __align__(16) uint8_t sout[128 * YQ_STRIP];
	const EncImg im = make_img(b, blockIdx.y, 0);
	const int band = blockIdx.x, tid = threadIdx.x;
	const int16_t *P = im.proc;
	for (int q = tid; q < 16 * 64; q += 256) {
		const int rr = q >> 6, c = (q & 63) * 8, row = band * 16 + rr, i = row * YW + c;
		const uint4 w = *reinterpret_cast<const uint4 *>(P + i);
		int o[10];
		o[0] = c ? (int)P[i - 1] : 0;
		o[1] = (int16_t)(w.x & 0xffff); o[2] = (int16_t)(w.x >> 16);
		o[3] = (int16_t)(w.y & 0xffff); o[4] = (int16_t)(w.y >> 16);
		o[5] = (int16_t)(w.z & 0xffff); o[6] = (int16_t)(w.z >> 16);
		o[7] = (int16_t)(w.w & 0xffff); o[8] = (int16_t)(w.w >> 16);
		o[9] = i + 8 < 512 * 512 ? (int)P[i + 8] : 0;
		uint32_t by[8];
		int big = 0;
#pragma unroll
		for (int t = 1; t <= 8; t++) big = max(big, nhw_iabs(o[t]));
		uint32_t lo, hi;
		if (big <= 6 && m1 > 6) { lo = 0x80808080u; hi = 0x80808080u; }   // |o| <= 6 quantises to "zero" whatever its neighbours are
		else {
#pragma unroll
		for (int t = 0; t < 8; t++) by[t] = (uint32_t)y_quant_byte(o[t], o[t + 1], o[t + 2], c + t >= 1, c + t < 511, m1);
		if (rr & 1) { lo = by[3] | (by[2] << 8) | (by[1] << 16) | (by[0] << 24); hi = by[7] | (by[6] << 8) | (by[5] << 16) | (by[4] << 24); }
		else { lo = by[0] | (by[1] << 8) | (by[2] << 16) | (by[3] << 24); hi = by[4] | (by[5] << 8) | (by[6] << 16) | (by[7] << 24); }
		}
		const int strip = c >> 2, off = (rr >> 1) * 8 + (rr & 1) * 4;
		*reinterpret_cast<uint32_t *>(sout + strip * YQ_STRIP + off) = lo;
		*reinterpret_cast<uint32_t *>(sout + (strip + 1) * YQ_STRIP + off) = hi;
	}
	__syncthreads();
	for (int idx = tid; idx < 512; idx += 256) {
		const int strip = idx >> 2, part = idx & 3;
		const uint32_t *src = reinterpret_cast<const uint32_t *>(sout + strip * YQ_STRIP + part * 16);
		*reinterpret_cast<uint4 *>(im.scan + strip * 2048 + band * 64 + part * 16) = make_uint4(src[0], src[1], src[2], src[3]);
	}
}

// ---- offsetUV + interleaved chroma scan, pointwise: both planes of 16 rows per CTA
#define CQ_STRIP 272
__global__ void __launch_bounds__(256) k_c_quant_scan(EncBatch b, int m2)
{
	__shared__ __align__(16) uint8_t sout[32 * CQ_STRIP];   // strips 272 bytes apart: lane = strip, 16-byte stores
	const EncImg imu = make_img(b, blockIdx.y, 0), imv = make_img(b, blockIdx.y, 1);
	const int band = blockIdx.x, tid = threadIdx.x;
	for (int q = tid; q < 16 * 32; q += 256) {
		const int rr = q >> 5, c = (q & 31) * 8, row = band * 16 + rr, i = row * CW + c;
		uint32_t by[2][8];
#pragma unroll
		for (int v = 0; v < 2; v++) {
			const int16_t *P = v ? imv.cproc : imu.cproc;
			const uint4 w = *reinterpret_cast<const uint4 *>(P + i);
			int o[9];
			o[0] = (int16_t)(w.x & 0xffff); o[1] = (int16_t)(w.x >> 16);
			o[2] = (int16_t)(w.y & 0xffff); o[3] = (int16_t)(w.y >> 16);
			o[4] = (int16_t)(w.z & 0xffff); o[5] = (int16_t)(w.z >> 16);
			o[6] = (int16_t)(w.w & 0xffff); o[7] = (int16_t)(w.w >> 16);
			o[8] = i + 8 < 65536 ? (int)P[i + 8] : 0;
			int big = 0;
#pragma unroll
			for (int t = 0; t < 8; t++) big = max(big, nhw_iabs(o[t]));
			if (big <= 6 && m2 > 6) {   // |o| <= 6 quantises to "zero" whatever its neighbours are
#pragma unroll
				for (int t = 0; t < 8; t++) by[v][t] = 128u;
				continue;
			}
#pragma unroll
			for (int t = 0; t < 8; t++) {
				int run = 0;
				bool pre = false;
				const int col = c + t;
				if (c_pairable(o[t])) {
					while (run < col && c_pairable(P[i + t - 1 - run])) run++;
				} else if (o[t] == 7) {
					while (run < col && P[i + t - 1 - run] == 7) run++;
					pre = run < col && c_bumps_next(P[i + t - 1 - run]);
				}
				by[v][t] = (uint32_t)c_quant_byte(o[t], o[t + 1], run, pre, col < 255, m2);
			}
		}
		uint32_t wv[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int t0 = (rr & 1) ? 7 - 2 * k : 2 * k, t1 = (rr & 1) ? 6 - 2 * k : 2 * k + 1;
			wv[k] = by[0][t0] | (by[1][t0] << 8) | (by[0][t1] << 16) | (by[1][t1] << 24);
		}
		const int strip = c >> 3;
		*reinterpret_cast<uint4 *>(sout + strip * CQ_STRIP + (rr >> 1) * 32 + (rr & 1) * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
	}
	__syncthreads();
	for (int idx = tid; idx < 512; idx += 256) {
		const int strip = idx >> 4, part = idx & 15;
		*reinterpret_cast<uint4 *>(imu.scan + 262144 + strip * 4096 + band * 256 + part * 16) =
		    *reinterpret_cast<const uint4 *>(sout + strip * CQ_STRIP + part * 16);
	}
}

// ---- q22/q23 side channel (enc_hq.cuh) -------------------------------------------------------
// LL1 copy: fo[r][j] = reconstruction[j][r] (the reference copies its transposed work plane)
__global__ void __launch_bounds__(256) k_hq_first_order(EncBatch b)
{
	__shared__ int16_t tile[32][33];
	const EncImg im = make_img(b, blockIdx.z, 0);
	const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	for (int r = ty; r < 32; r += 8) tile[r][tx] = im.proc[(y0 + r) * YW + x0 + tx];
	__syncthreads();
	for (int r = ty; r < 32; r += 8) im.hq_fo[(x0 + r) * 256 + y0 + tx] = tile[tx][r];
}

// E17: residual codes of LL1 cell (r, j) adjust up to three consecutive cells of the copy
__global__ void __launch_bounds__(256) k_hq_e17(EncBatch b)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	const int r = blockIdx.x, j = threadIdx.x;
	if (j >= 254) return;
	int d[3];
	if (!hq_e17_delta(im.ll1[r * 256 + j], d)) return;
	for (int k = 0; k < 3; k++)
		if (d[k]) atomic_add_s16(im.hq_fo + j * 256 + r + k, d[k]);
}

// LH1 rebuilt from the scan bytes, then half synthesis + comparison of row blockIdx.x
__global__ void __launch_bounds__(256) k_hq_band(EncBatch b)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	const int c = blockIdx.x * 256 + threadIdx.x;
	auto byte_at = [&](int cc) { return (int)im.scan[y_scan_pos(cc >> 8, 256 + (cc & 255))]; };
	im.hq_band[c] = (int16_t)hq_band_cell(byte_at, c);
}

__global__ void __launch_bounds__(256) k_hq_tags(EncBatch b, int q)
{
	const EncImg im = make_img(b, blockIdx.y, 0);
	hq_tag_pair(im, q, blockIdx.x, threadIdx.x);
}

// lists: rows counted and written in parallel (thread = row), the short tail by one thread
__global__ void __launch_bounds__(256) k_hq_lists(EncBatch b, int q)
{
	__shared__ int np[257], nw[257], nc[257], n3[257];
	__shared__ int bad;
	const EncImg im = make_img(b, blockIdx.x, 0);
	EncHdr *h = im.hdr;
	const int r = threadIdx.x;
	const uint8_t *tag = im.hq_tag + r * 512;
	{
		int w;
		np[r] = hq_collect_row(im, r, nullptr, nullptr, w);
		nw[r] = w;
		int c1 = 0, c3 = 0;
		for (int k = 0; k < 2; k++) c1 += (tag[254 + k] == 1 || tag[254 + k] == 2);
		if (q > 22)
			for (int j = 0; j < 512; j++) c3 += (tag[j] >= 3);
		nc[r] = c1;
		n3[r] = c3;
	}
	__syncthreads();
	if (r < 4) {
		int *v = r == 0 ? np : r == 1 ? nw : r == 2 ? nc : n3;
		int run = 0;
		for (int k = 0; k < 256; k++) { const int x = v[k]; v[k] = run; run += x; }
		v[256] = run;
	}
	__syncthreads();
	if (r == 0) bad = (np[256] + 16 > NHW_CAP_LIST || nc[256] > NHW_CAP_CHAR_RES1 || n3[256] > NHW_CAP_QSETTING3) ? 1 : 0;
	__syncthreads();
	if (bad) { if (r == 0 && h->status == 0) h->status = NHW_ERR_OVERFLOW_DEV; return; }
	{
		int w;
		hq_collect_row(im, r, im.tmp1 + np[r], im.tmp3 + nw[r], w);
		int o = nc[r];
		for (int k = 0; k < 2; k++) {
			const int g = tag[254 + k];
			if (g == 1 || g == 2) im.char_res1[o++] = (uint16_t)(r * 256 + 2 * k + (g - 1));
		}
		if (q > 22) {
			int o3 = n3[r];
			for (int j = 0; j < 512; j++) {
				const int g = tag[j];
				if (g >= 3) im.qsetting3[o3++] = (uint32_t)((r * 512 + j) << 1) + (g == 4 ? 1u : 0u);
			}
		}
	}
	__syncthreads();
	if (r == 0) {
		h->char_res1_len = nc[256];
		h->qsetting3_len = n3[256];
		y_e18_finish_list_image(im, 6, np[256], nw[256]);
	}
}

// ---- peephole passes over the luma scan, one CTA per image (enc_seg.cuh)
#define PEEP_THREADS 512
// non-zero bitmap of `count` stream bytes starting at s (16-byte aligned, count % 32 == 0) into shared memory
// (count % 1024 == 0; `sum` = the second level, count / 1024 words, needs a barrier of its own: callers sync after)
__device__ __forceinline__ void build_nz_bitmap(const uint8_t *s, int count, uint32_t *bits, uint32_t *sum, int tid, int nthreads)
{
	for (int w = tid; w < count / 32; w += nthreads) {
		const uint4 a = reinterpret_cast<const uint4 *>(s)[2 * w], b = reinterpret_cast<const uint4 *>(s)[2 * w + 1];
		bits[w] = nz_mask4(a.x) | (nz_mask4(a.y) << 4) | (nz_mask4(a.z) << 8) | (nz_mask4(a.w) << 12) | (nz_mask4(b.x) << 16) |
		          (nz_mask4(b.y) << 20) | (nz_mask4(b.z) << 24) | (nz_mask4(b.w) << 28);
	}
	__syncthreads();
	// summary word j = ballot of "bits[32 j + lane] != 0" (one conflict-free load per warp and word)
	for (int j = tid >> 5; j < count / 1024; j += nthreads >> 5) {
		const uint32_t m = __ballot_sync(0xffffffffu, bits[32 * j + (tid & 31)] != 0);
		if ((tid & 31) == 0) sum[j] = m;
	}
}

// 4 bits: which of the 4 stream bytes in a word are the +-8 codes (136 / 120), the only bytes the passes look for
__device__ __forceinline__ uint32_t pm8_mask4(uint32_t w)
{
	return (~nz_mask4(w ^ (0x88888888u ^ 0x80808080u)) | ~nz_mask4(w ^ (0x78787878u ^ 0x80808080u))) & 15u;
}
__device__ __forceinline__ uint32_t pm8_mask16(const uint4 &a) { return pm8_mask4(a.x) | (pm8_mask4(a.y) << 4) | (pm8_mask4(a.z) << 8) | (pm8_mask4(a.w) << 12); }

// The stream is read 16 bytes at a time; only the +-8 codes can start a pair (pass A) or change (passes B, C), so a
// vector without one is skipped.  Pass A never turns a zero into a non-zero or back (136/120 -> 132..135 / 201), so
// the non-zero bitmap of the stream after pass A is built from the same loads as the search for chain heads.
__global__ void __launch_bounds__(PEEP_THREADS) k_peephole(EncBatch b)
{
	extern __shared__ __align__(16) uint32_t nzbits[];   // 262144 bits
	__shared__ int n_heads, sel1, sel2;
	__shared__ uint32_t nzsum[256];
	const EncImg im = make_img(b, blockIdx.x, 0);
	uint8_t *s = im.scan;
	const int N = 262144;
	const int tid = threadIdx.x;
	uint4 *s4 = reinterpret_cast<uint4 *>(s);
	uint4 *out4 = reinterpret_cast<uint4 *>(im.aux);                    // changed vectors, at their own index
	int *heads = reinterpret_cast<int *>(im.aux) + N / 4;               // chain heads after them
	if (tid == 0) { n_heads = 0; sel1 = 0; sel2 = 0; }
	__syncthreads();
	// pass A, phase 1: chain heads on the un-edited stream (+ the non-zero bitmap)
	for (int w = tid; w < N / 32; w += PEEP_THREADS) {
		const uint4 a = s4[2 * w], c = s4[2 * w + 1];
		nzbits[w] = nz_mask4(a.x) | (nz_mask4(a.y) << 4) | (nz_mask4(a.z) << 8) | (nz_mask4(a.w) << 12) | (nz_mask4(c.x) << 16) |
		            (nz_mask4(c.y) << 20) | (nz_mask4(c.z) << 24) | (nz_mask4(c.w) << 28);
		for (uint32_t m = pm8_mask16(a) | (pm8_mask16(c) << 16); m; m &= m - 1) {
			const int i = 32 * w + __ffs(m) - 1;
			if (i < N - 4 && peep_pair_candidate(s, i, N) && !peep_pair_candidate(s, i - 4, N)) heads[atomicAdd(&n_heads, 1)] = i;
		}
	}
	__syncthreads();
	// pass A, phase 2: merge along each chain (chains are disjoint)
	for (int k = tid; k < n_heads; k += PEEP_THREADS) peep_merge_chain(s, heads[k], N);
	__syncthreads();
	if (tid < 4) { s[tid] = 128; s[N - 4 + tid] = 128; }
	if (tid == 0) { nzbits[0] &= ~15u; nzbits[N / 32 - 1] &= ~(15u << 28); }
	__syncthreads();
	for (int j = tid >> 5; j < N / 1024; j += PEEP_THREADS >> 5) {   // second level of the bitmap
		const uint32_t m = __ballot_sync(0xffffffffu, nzbits[32 * j + (tid & 31)] != 0);
		if ((tid & 31) == 0) nzsum[j] = m;
	}
	__syncthreads();
	const NzBits nz{nzbits, 0, N, nzsum};
	// passes B + C: every byte that changes, from the pass-A stream; changed vectors are parked and written back
	// once everybody has read
	int a1 = 0, a2 = 0;
	uint32_t changed = 0;   // N / 16 / PEEP_THREADS = 32 vectors per thread
	for (int w = tid, it = 0; w < N / 16; w += PEEP_THREADS, it++) {
		uint4 a = s4[w];
		uint32_t m = pm8_mask16(a);
		if (!m) continue;
		uint32_t v[4] = {a.x, a.y, a.z, a.w};
		bool ch = false;
		for (; m; m &= m - 1) {
			const int k = __ffs(m) - 1, i = 16 * w + k;
			int x1, x2;
			const uint32_t old = (v[k >> 2] >> (8 * (k & 3))) & 255u, nw = (uint32_t)peep_select_byte(s, nz, i, N, x1, x2);
			a1 += x1;
			a2 += x2;
			if (nw != old) { v[k >> 2] ^= (old ^ nw) << (8 * (k & 3)); ch = true; }
		}
		if (ch) { out4[w] = make_uint4(v[0], v[1], v[2], v[3]); changed |= 1u << it; }
	}
	if (a1) atomicAdd(&sel1, a1);
	if (a2) atomicAdd(&sel2, a2);
	__syncthreads();
	for (int w = tid, it = 0; changed; w += PEEP_THREADS, it++)
		if (changed >> it & 1u) { s4[w] = out4[w]; changed &= ~(1u << it); }
	if (tid == 0) { im.hdr->select1 = sel1; im.hdr->select2 = sel2; }
}

// ---- entropy stage, one CTA per image, one thread per stream segment (enc_seg.cuh)
__device__ __forceinline__ void block_excl_scan3(int *a, int *b, int *c, int t)   // 256 entries each, in place
{
	__syncthreads();
	if (t < 3) {
		int *v = t == 0 ? a : t == 1 ? b : c;
		int run = 0;
		for (int k = 0; k < SEG_THREADS; k++) { int x = v[k]; v[k] = run; run += x; }
		v[SEG_THREADS] = run;
	}
	__syncthreads();
}

__global__ void __launch_bounds__(SEG_THREADS) k_entropy(EncBatch b)
{
	extern __shared__ __align__(16) uint32_t nzbits[];   // 262144 bits (luma part), reused for the chroma part
	__shared__ uint32_t nzsum[256];
	__shared__ int hist_sym[256], hist_run[256];
	__shared__ int sbits[SEG_THREADS + 1], sn1[SEG_THREADS + 1], sn2[SEG_THREADS + 1], bnd[SEG_THREADS + 1];
	__shared__ uint32_t s_weight[354];
	__shared__ uint16_t s_sym[354];
	__shared__ int s_select, s_k, s_b, s_rc, s_bad, s_word0;
	const EncImg im = make_img(b, blockIdx.x, 0);
	__shared__ PackState st;                       // histograms / alphabet: the serial sections run on shared memory
	__shared__ uint8_t book_scratch[2048];
	uint8_t *s = im.scan;
	EncHdr *h = im.hdr;
	const int t = threadIdx.x;
	if (t == 0) { s_word0 = 0; s_rc = 0; }
	for (int part = 0; part < 2; part++) {
		const int p1 = part ? 262144 : 0, p2 = part ? 393216 : 262144;
		const int S = (p2 - p1) / SEG_THREADS;
		uint8_t saved = 0;
		if (t == 0) {
			if (!part) { saved = s[262144]; s[262144] = 3; } else s[393215] = s[393214];
			s_bad = 0;
		}
		hist_sym[t] = 0;
		hist_run[t] = 0;
		__syncthreads();
		build_nz_bitmap(s + p1, p2 - p1, nzbits, nzsum, t, SEG_THREADS);
		__syncthreads();
		// Work-balanced segments: thread t gets the stretch between the (t*T/256)-th and the ((t+1)*T/256)-th non-zero
		// byte (T of them in the stream).  With equal-sized segments a warp waited for its busiest lane while the
		// lanes over flat image areas had nothing to do (6 of 32 lanes active on average).
		{
			const int W = (p2 - p1) / 32 / SEG_THREADS;
			int cnt = 0;
			for (int k = 0; k < W; k++) cnt += __popc(nzbits[t * W + k]);
			sbits[t] = cnt;
			__syncthreads();
			if (t == 0) {
				int run = 0;
				for (int k = 0; k < SEG_THREADS; k++) { const int x = sbits[k]; sbits[k] = run; run += x; }
				sbits[SEG_THREADS] = run;
			}
			__syncthreads();
			const int T = sbits[SEG_THREADS];
			int pos = t ? p2 : p1;
			if (t && T > 0) {
				int rem = (int)((long long)t * T / SEG_THREADS);
				int lo = 0, hi = SEG_THREADS - 1;            // last j with sbits[j] <= rem
				while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sbits[mid] <= rem) lo = mid; else hi = mid - 1; }
				rem -= sbits[lo];
				for (int w = lo * W; w < (lo + 1) * W; w++) {
					const uint32_t m = nzbits[w];
					const int pc = __popc(m);
					if (rem < pc) { pos = p1 + 32 * w + (int)__fns(m, 0, rem + 1); break; }
					rem -= pc;
				}
			}
			__syncthreads();
			bnd[t] = pos;
			if (t == 0) bnd[SEG_THREADS] = p2;
			__syncthreads();
		}
		SegStream ss{s, p1, p2, S, NzBits{nzbits, p1, p2 - p1, nzsum}, bnd};
		seg_stats(ss, t, [&](bool run, int idx) { atomicAdd(run ? &hist_run[idx] : &hist_sym[idx], 1); });
		__syncthreads();
		st.rle_buf[t] = hist_sym[t];
		st.rle_128[t] = hist_run[t];
		__syncthreads();
		if (t == 0) {
			int select = part ? 3 : 4, k = 0;
			int rc = pack_enumerate(st, select, k);
			s_select = select; s_k = k;
			if (rc) s_rc = rc;
		}
		__syncthreads();
		if (s_rc) break;
		const int k = s_k;
		// stable rank sort by decreasing weight across the CTA
		for (int i = t; i < k; i += SEG_THREADS) { s_weight[i] = st.weight[i]; s_sym[i] = st.sym[i]; }
		__syncthreads();
		for (int i = t; i < k; i += SEG_THREADS) {
			const uint32_t w = s_weight[i];
			int rank = 0;
			for (int j = 0; j < k; j++) rank += (s_weight[j] > w || (s_weight[j] == w && j < i)) ? 1 : 0;
			st.weight[rank] = w;
			st.sym[rank] = s_sym[i];
		}
		__syncthreads();
		if (t == 0) {
			int bb = 0;
			int rc = pack_finish(st, part, s_select, k, bb);
			s_b = bb;
			if (rc) s_rc = rc;
		}
		__syncthreads();
		if (s_rc) break;
		hist_sym[t] = st.rle_buf[t];     // now: byte -> rank, run length -> rank
		hist_run[t] = st.rle_128[t];
		__syncthreads();
		const int select = s_select;
		const bool zone = (part == 0 && select == 4 && s_b == 1);
		// counting pass
		{
			int nb = 0, a1 = 0, a2 = 0;
			int bad = seg_emit(ss, t, hist_sym, hist_run, select, zone, [&](uint32_t, int len) { nb += len; },
			                   [&](int) { a1++; }, [&](int) { a2++; });
			if (bad) s_bad = 1;
			sbits[t] = nb; sn1[t] = a1; sn2[t] = a2;
		}
		block_excl_scan3(sbits, sn1, sn2, t);
		const int total = sbits[SEG_THREADS];
		const int nwords = total > 0 ? (total + 31) / 32 : 1;
		const int word0 = s_word0;
		if (s_bad || word0 + nwords >= NHW_WORDS_LIMIT) {
			if (t == 0) s_rc = s_bad ? NHW_ERR_CODEBOOK_DEV : NHW_ERR_OVERFLOW_DEV;
			__syncthreads();
			break;
		}
		if (!part) {
			for (int i = t; i < (sn1[SEG_THREADS] >> 3) + 1; i += SEG_THREADS) im.sel1[i] = 0;
			for (int i = t; i < (sn2[SEG_THREADS] >> 3) + 1; i += SEG_THREADS) im.sel2[i] = 0;
		}
		__syncthreads();
		// emission pass: codes are OR-ed into the (pre-zeroed) word array at their bit offsets
		{
			long off = sbits[t];
			int o1 = sn1[t], o2 = sn2[t];
			uint32_t *words = im.words;
			auto or_byte = [](uint8_t *base, int idx, int bit) {
				uint32_t *w = reinterpret_cast<uint32_t *>(base) + (idx >> 5);
				atomicOr(w, (uint32_t)bit << (((idx >> 3) & 3) * 8 + (7 - (idx & 7))));
			};
			seg_emit(ss, t, hist_sym, hist_run, select, zone,
			         [&](uint32_t code, int len) { seg_put_bits([&](int w, uint32_t v) { atomicOr(&words[w], v); }, word0, off, code, len); off += len; },
			         [&](int bit) { if (!part && bit) or_byte(im.sel1, o1, 1); o1++; },
			         [&](int bit) { if (!part && bit) or_byte(im.sel2, o2, 1); o2++; });
		}
		__syncthreads();
		if (t == 0) {
			if (!part) {
				h->size_data1 = word0 + nwords;
				h->wavelet_type = (select > 4 || s_b == 0) ? 4 : 0;
				h->select1 = (sn1[SEG_THREADS] >> 3) + 1;
				h->select2 = (sn2[SEG_THREADS] >> 3) + 1;
				s[262144] = saved;
			} else h->size_data2 = word0 + nwords;
			s_word0 = word0 + nwords;
			pack_codebook(im, st, part, k, book_scratch);
		}
		__syncthreads();
	}
	if (t == 0) h->status = s_rc;
}

// ---- inverse transform of one level (wavelet_synthesis, encoder/wavelet_filterbank.c:305-496)
// pass 1: rows of the band plane J(k,m) -> T(k,y) (no normalisation), natural row layout
template <int N>
__global__ void __launch_bounds__(256) k_idwt_rows(const int16_t *__restrict__ in, int16_t *__restrict__ out,
                                                   size_t in_stride, size_t out_stride, int row_stride)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int k = blockIdx.x * 8 + warp;
	const int16_t *src = in + (size_t)blockIdx.y * in_stride + k * row_stride;
	int16_t *dst = out + (size_t)blockIdx.y * out_stride + k * row_stride;
	constexpr int M = N / 2;
	auto l = [&](int t) { return (int)src[t]; };
	auto h = [&](int t) { return (int)src[M + t]; };
	for (int t = lane; t < M; t += 32) {
		int ev, od;
		inverse_pair(l, h, t, M, false, ev, od);
		*reinterpret_cast<uint32_t *>(dst + 2 * t) = (uint32_t)(uint16_t)ev | ((uint32_t)(uint16_t)od << 16);
	}
}

// pass 2: columns of T -> natural image rows out(y, x), normalised.  CTA = 32 columns y.
template <int N>
__global__ void __launch_bounds__(256) k_idwt_cols_t(const int16_t *__restrict__ in, int16_t *__restrict__ out,
                                                     size_t in_stride, size_t out_stride, int row_stride)
{
	extern __shared__ int16_t tile[];   // [N][33]
	const int y0 = blockIdx.x * 32;
	const int16_t *src = in + (size_t)blockIdx.y * in_stride;
	int16_t *dst = out + (size_t)blockIdx.y * out_stride;
	for (int i = threadIdx.x; i < N * 32; i += 256) {
		int k = i >> 5, yy = i & 31;
		tile[k * 33 + yy] = src[k * row_stride + y0 + yy];
	}
	__syncthreads();
	constexpr int M = N / 2;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int yy = warp; yy < 32; yy += 8) {
		auto l = [&](int t) { return (int)tile[t * 33 + yy]; };
		auto h = [&](int t) { return (int)tile[(M + t) * 33 + yy]; };
		int16_t *row = dst + (y0 + yy) * row_stride;
		for (int t = lane; t < M; t += 32) {
			int ev, od;
			inverse_pair(l, h, t, M, true, ev, od);
			*reinterpret_cast<uint32_t *>(row + 2 * t) = (uint32_t)(uint16_t)ev | ((uint32_t)(uint16_t)od << 16);
		}
	}
}

// Both passes of one synthesis level in ONE kernel, the band plane held in shared memory (the counterpart of k_dwt_level):
// load the N x N band plane J(k, m) -> row pass in place (a warp per band row; a lane takes its inputs into registers
// before any lane writes) -> column pass out of shared memory, normalised, written as natural image rows out(y, x).
// The row pitch of N + 2 int16 (an odd number of 32-bit words) keeps the column pass's lanes on different banks.
// Replaces k_idwt_rows + k_idwt_cols_t and the trip of the intermediate plane T through HBM.
template <int N, int THREADS>
__global__ void __launch_bounds__(THREADS) k_idwt_level(const int16_t *__restrict__ in, int16_t *__restrict__ out,
                                                        size_t in_stride, size_t out_stride, int row_stride)
{
	extern __shared__ __align__(16) int16_t band[];   // [N][N + 2]
	constexpr int P = N + 2, M = N / 2, PER = M / 32;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int16_t *src = in + (size_t)blockIdx.x * in_stride;
	int16_t *dst = out + (size_t)blockIdx.x * out_stride;
	for (int idx = tid; idx < N * N / 8; idx += THREADS) {
		const int k = idx / (N / 8), c = idx % (N / 8);
		const uint4 v = *reinterpret_cast<const uint4 *>(src + (size_t)k * row_stride + c * 8);
		uint32_t *d = reinterpret_cast<uint32_t *>(band + k * P + c * 8);   // rows are 4-byte aligned only
		d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
	}
	__syncthreads();
	for (int k = warp; k < N; k += THREADS / 32) {
		int16_t *row = band + k * P;
		uint32_t o[PER];
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const int t = lane + 32 * u;
			auto l = [&](int i) { return (int)row[i]; };
			auto h = [&](int i) { return (int)row[M + i]; };
			int ev, od;
			inverse_pair(l, h, t, M, false, ev, od);
			o[u] = (uint32_t)(uint16_t)ev | ((uint32_t)(uint16_t)od << 16);
		}
		__syncwarp();
#pragma unroll
		for (int u = 0; u < PER; u++) reinterpret_cast<uint32_t *>(row)[lane + 32 * u] = o[u];
	}
	__syncthreads();
	for (int y = warp; y < N; y += THREADS / 32) {
		auto l = [&](int i) { return (int)band[i * P + y]; };
		auto h = [&](int i) { return (int)band[(M + i) * P + y]; };
		int16_t *row = dst + (size_t)y * row_stride;
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const int t = lane + 32 * u;
			int ev, od;
			inverse_pair(l, h, t, M, true, ev, od);
			*reinterpret_cast<uint32_t *>(row + 2 * t) = (uint32_t)(uint16_t)ev | ((uint32_t)(uint16_t)od << 16);
		}
	}
}

// copy an N x N region between planes of possibly different row strides (one thread per 2 cells)
__global__ void k_copy_region(const int16_t *__restrict__ src, size_t src_slot, int src_stride,
                              int16_t *__restrict__ dst, size_t dst_slot, int dst_stride, int N)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // pair index
	const int per_row = N / 2;
	if (idx >= N * per_row) return;
	const int r = idx / per_row, c = (idx % per_row) * 2;
	const uint32_t v = *reinterpret_cast<const uint32_t *>(src + (size_t)blockIdx.y * src_slot + r * src_stride + c);
	*reinterpret_cast<uint32_t *>(dst + (size_t)blockIdx.y * dst_slot + r * dst_stride + c) = v;
}

__global__ void k_zero_bytes(uint8_t *base, size_t slot, size_t off, size_t count)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // 16-byte units
	if (i * 16 < count) reinterpret_cast<uint4 *>(base + (size_t)blockIdx.y * slot + off)[i] = make_uint4(0, 0, 0, 0);
}

// ---- chroma LL code: a serial coder (each step looks a few bytes ahead and skips a variable distance), run by
// one thread per image out of shared memory; the warp stages the 8192 input bytes and the output.
#define CLL_PAD 32
// Parallel form of the chroma LL coder (enc_c.cuh: c_ll_step is a pure function of the position and emits one byte):
// the step is evaluated at all 8191 positions, every thread walks the orbit of its 32-position segment's first
// position (speculation), one thread stitches the segments (the true chain enters a segment somewhere and runs
// until it meets the speculative one or leaves), a CTA scan of the visit counts gives the output offsets.
// Per-position arrays are laid out 36 bytes per segment so that the lanes of a warp hit different banks.
#define CLL_SW(i) ((i) + (((i) >> 5) << 2))
__global__ void __launch_bounds__(256) k_c_ll_code(EncBatch b)
{
	__shared__ __align__(16) uint8_t sx[8192 + CLL_PAD];
	__shared__ uint8_t snext[8192 + 1024], sbyte[8192 + 1024], vis[8192 + 1024];
	__shared__ int seg_exit[256], seg_merge[256], seg_cnt[257];
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int t = threadIdx.x;
	for (int k = t; k < (8192 + CLL_PAD) / 4; k += 256) {
		uint32_t w = reinterpret_cast<const uint32_t *>(im.tree1 + 16384)[k];
		if (k < 2048) w &= 0xfcfcfcfcu;   // x[i] &= 252 (compress_pixel.c:886)
		reinterpret_cast<uint32_t *>(sx)[k] = w;
	}
	__syncthreads();
	const uint8_t *x = sx - 16384;        // indexed as in tree1
	for (int p = t; p < 8192; p += 256) {
		vis[CLL_SW(p)] = 0;
		if (p >= 1) {
			int byte;
			const int nx = c_ll_step(x, 16384 + p, byte);
			snext[CLL_SW(p)] = (uint8_t)(nx - (16384 + p));
			sbyte[CLL_SW(p)] = (uint8_t)byte;
		}
	}
	__syncthreads();
	{
		int p = t ? 32 * t : 1;
		const int end = 32 * t + 32;
		while (p < end) { vis[CLL_SW(p)] = 1; p += snext[CLL_SW(p)]; }
		seg_exit[t] = p;
	}
	__syncthreads();
	if (t == 0) {
		int p = seg_exit[0];
		seg_merge[0] = 0;
		for (int k = 1; k < 256; k++) {
			const int end = 32 * k + 32;
			int i = p;
			while (i < end && vis[CLL_SW(i)] != 1) { vis[CLL_SW(i)] = 2; i += snext[CLL_SW(i)]; }
			if (i < end) { seg_merge[k] = i; p = seg_exit[k]; }     // met the speculative chain: it is the true one from here
			else { seg_merge[k] = end; p = i; }                     // never met it inside this segment
		}
	}
	__syncthreads();
	{
		const int m = seg_merge[t];
		int n = 0;
		for (int p = 32 * t; p < 32 * t + 32; p++) {
			if (p < m && vis[CLL_SW(p)] == 1) vis[CLL_SW(p)] = 0;
			n += vis[CLL_SW(p)] ? 1 : 0;
		}
		seg_cnt[t] = n;
	}
	__syncthreads();
	if (t == 0) {
		int run = 0;
		for (int k = 0; k < 256; k++) { const int c = seg_cnt[k]; seg_cnt[k] = run; run += c; }
		seg_cnt[256] = run;
	}
	__syncthreads();
	const int j0 = im.hdr->y_res_comp, cnt = 1 + seg_cnt[256];
	{
		uint8_t *o = im.llcode + j0 + 1 + seg_cnt[t];
		for (int p = 32 * t; p < 32 * t + 32; p++)
			if (vis[CLL_SW(p)]) *o++ = sbyte[CLL_SW(p)];
	}
	if (t == 0) { im.llcode[j0] = sx[0]; im.hdr->end_ch_res = j0 + cnt; }
	for (int k = t; k < 2048; k += 256) reinterpret_cast<uint32_t *>(im.tree1 + 16384)[k] = reinterpret_cast<const uint32_t *>(sx)[k];
}

// final: container bytes.  One CTA per image: thread 0 lays the sections out (enc_pack.cuh), all threads copy.
#define WS_MAX_SECTIONS 40
__global__ void __launch_bounds__(256) k_write_stream(EncBatch b, int n, uint8_t *out, uint32_t *len, int32_t *status)
{
	__shared__ const uint8_t *src[WS_MAX_SECTIONS];
	__shared__ int off[WS_MAX_SECTIONS + 1];
	__shared__ uint8_t hdr[64];
	__shared__ int nsec, st;
	const int i = blockIdx.x;
	const EncImg im = make_img(b, i, 0);
	if (threadIdx.x == 0) {
		st = im.hdr->status;
		nsec = 0;
		off[0] = 0;
		if (st == 0)
			stream_layout(im, hdr, [&](const uint8_t *s, int cnt) {
				if (nsec < WS_MAX_SECTIONS) { src[nsec] = s; off[nsec + 1] = off[nsec] + cnt; nsec++; }
			});
		// the output slot is NHW_MAX_STREAM_BYTES: a stream that would not fit (the section caps add up to more) is
		// reported, not written
		if (off[nsec] > (int)NHW_MAX_STREAM_BYTES) { st = NHW_ERR_OVERFLOW_DEV; nsec = 0; }
		if (len) len[i] = (uint32_t)off[nsec];
		if (status) status[i] = st;
	}
	__syncthreads();
	uint8_t *dst = out + (size_t)i * NHW_MAX_STREAM_BYTES;
	for (int k = 0; k < nsec; k++) {
		const uint8_t *s = src[k];
		const int o = off[k], cnt = off[k + 1] - o;
		for (int t = threadIdx.x; t < cnt; t += 256) dst[o + t] = s ? s[t] : (uint8_t)0;
	}
}

// ---- E8 (q <= 12, enc_lowq.cuh: y_e8_smooth_walk): a raster-serial walk over the 128 x 128 LL2 band, one WARP per
// image.  The band is staged in shared memory (coalesced) and lane 0 walks it at shared-memory latency; the cells the
// walk silences (descendants in the level-1 bands, siblings in the level-2 bands) are never read by it, so the walk
// only records the requests -- per LL2 cell the largest threshold asked for its diagonal-band children, per plane index
// a sibling bit -- and the 32 lanes apply them afterwards.  The warp then writes the band back.
struct E8Deferred {
	uint8_t *child;      // [128 * 128] by LL2 cell: 0 none, else the threshold for the diagonal-band children
	uint32_t *sib;       // bitmap over flat plane indices 0 .. 65536 + 63
	__device__ void children(int p, int, int, int c) const
	{
		const int k = ((p >> 9) << 7) + (p & 127);
		if (c > child[k]) child[k] = (uint8_t)c;
	}
	__device__ void siblings(int p) const { sib[p >> 5] |= 1u << (p & 31); }
};
#define E8_SMEM (128 * 128 * 2 + 128 * 128 + (65536 / 32 + 2) * 4)
__global__ void __launch_bounds__(32) k_e8_staged(EncBatch b, int q)
{
	extern __shared__ __align__(16) uint8_t e8mem[];
	int16_t *band = reinterpret_cast<int16_t *>(e8mem);
	uint8_t *child = e8mem + 128 * 128 * 2;
	uint32_t *sib = reinterpret_cast<uint32_t *>(child + 128 * 128);
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int lane = threadIdx.x;
	for (int k = lane; k < 128 * 16; k += 32) reinterpret_cast<uint4 *>(band)[k] = *reinterpret_cast<const uint4 *>(im.proc + (k >> 4) * YW + (k & 15) * 8);
	for (int k = lane; k < 128 * 128 / 16; k += 32) reinterpret_cast<uint4 *>(child)[k] = make_uint4(0, 0, 0, 0);
	for (int k = lane; k < 65536 / 32 + 2; k += 32) sib[k] = 0;
	__syncwarp();
	if (lane == 0) y_e8_smooth_walk(band, 128, E8Deferred{child, sib}, q);
	__syncwarp();
	const E8Thr t = e8_thresholds(q);
	int16_t *P = im.proc;
	for (int k = lane; k < 128 * 128; k += 32)
		if (child[k]) e8_silence_children(P, ((k >> 7) << 9) + (k & 127), t.t6, t.t6 + 6, child[k]);
	for (int w = lane; w < 65536 / 32 + 2; w += 32)
		for (uint32_t m = sib[w]; m; m &= m - 1) e8_silence_siblings(P, 32 * w + __ffs(m) - 1);
	__syncwarp();
	for (int k = lane; k < 128 * 16; k += 32) *reinterpret_cast<uint4 *>(im.proc + (k >> 4) * YW + (k & 15) * 8) = reinterpret_cast<const uint4 *>(band)[k];
}

// ---- isolated-coefficient shrink at q <= 16 (enc_lowq.cuh: y_recons_shrink_lowq_cell): one CTA per image, thread =
// column, rows top to bottom with a barrier per row
__global__ void __launch_bounds__(256) k_recons_shrink_lowq(EncBatch b)
{
	const EncImg im = make_img(b, blockIdx.x, 0);
	const int j = threadIdx.x;
	for (int r = 1; r < 255; r++) {
		if (j >= 1 && j < 255) y_recons_shrink_lowq_cell(im.jpeg, r, j);
		__syncthreads();
	}
}

// exclusive prefix of the stream lengths (one CTA, n <= a few thousand) ...
__global__ void k_stream_offsets(const uint32_t *__restrict__ len, uint64_t *__restrict__ offs, int n)
{
	__shared__ uint64_t part[1024];
	const int t = threadIdx.x, per = (n + 1023) / 1024;
	uint64_t s = 0;
	for (int i = t * per; i < n && i < (t + 1) * per; i++) s += len[i];
	part[t] = s;
	__syncthreads();
	if (t == 0) {
		uint64_t run = 0;
		for (int k = 0; k < 1024; k++) { uint64_t v = part[k]; part[k] = run; run += v; }
		offs[n] = run;
	}
	__syncthreads();
	uint64_t run = part[t];
	for (int i = t * per; i < n && i < (t + 1) * per; i++) { offs[i] = run; run += len[i]; }
}

// ... and the gather of the fixed-size slots into one dense buffer (one CTA per image)
__global__ void k_pack_streams(const uint8_t *__restrict__ slots, const uint32_t *__restrict__ len,
                               const uint64_t *__restrict__ offs, uint8_t *__restrict__ dense)
{
	const int i = blockIdx.x;
	const uint8_t *src = slots + (size_t)i * NHW_MAX_STREAM_BYTES;
	uint8_t *dst = dense + offs[i];
	const uint32_t L = len[i];
	for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) dst[k] = src[k];
}


}  // namespace

namespace nhw {

bool encode_device_init(nhw_ctx *c)
{
	(void)c;
	bool ok = check(cudaFuncSetAttribute(k_ll2_code, cudaFuncAttributeMaxDynamicSharedMemorySize, LL2_CODE_SMEM), "attr k_ll2_code");
	ok = ok && check(cudaFuncSetAttribute(k_e8_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, E8_SMEM), "attr k_e8_staged");
	ok = ok && check(cudaFuncSetAttribute(k_recons_ll2_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, LL2_SMEM_BYTES), "attr k_recons_ll2_wave");
	ok = ok && check(cudaFuncSetAttribute(k_idwt_cols_t<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 33 * 2), "attr k_idwt_cols_t");
	ok = ok && check(cudaFuncSetAttribute(k_idwt_cols_t<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 33 * 2), "attr k_idwt_cols_t");
	ok = ok && check(cudaFuncSetAttribute(k_idwt_level<256, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 258 * 2), "attr k_idwt_level");
	ok = ok && check(cudaFuncSetAttribute(k_idwt_level<128, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 130 * 2), "attr k_idwt_level");
	return ok;
}

EncBatch enc_batch_of(nhw_ctx *c)
{
	EncBatch b;
	b.y_proc = c->y_proc + NHW_GUARD_S;
	b.y_jpeg = c->y_jpeg + NHW_GUARD_S;
	b.y_aux = c->y_aux + NHW_GUARD_S;
	b.y_ll1 = c->y_ll1 + NHW_GUARD_S;
	b.y_ll2s = c->y_ll2save + NHW_GUARD_S;
	b.y_hq = c->y_aux2 + NHW_GUARD_S;
	b.c_proc = c->c_proc + NHW_GUARD_S;
	b.c_jpeg = c->c_jpeg + NHW_GUARD_S;
	b.c_aux = c->c_aux + NHW_GUARD_S;
	b.c_ll1 = c->c_ll1 + NHW_GUARD_S;
	b.c_ll2s = c->c_ll2save + NHW_GUARD_S;
	b.bytes = c->enc_bytes;
	b.hdr = c->enc_hdr;
	return b;
}

// one synthesis level of n_planes N x N band planes (row stride `stride`, planes `slot` apart): in -> out
static void idwt_level(nhw_ctx *c, int n_planes, const int16_t *in, int16_t *out, size_t slot, int N, int stride)
{
	if (N == 256) NHW_LAUNCH_L(c, "k_idwt_level<256>", (k_idwt_level<256, 1024>), n_planes, 1024, 256 * 258 * 2, in, out, slot, slot, stride);
	else NHW_LAUNCH_L(c, "k_idwt_level<128>", (k_idwt_level<128, 256>), n_planes, 256, 128 * 130 * 2, in, out, slot, slot, stride);
}

// luma inverse level-2 transform: jpeg region (256x256) -> proc region, natural orientation
static void idwt_luma256(nhw_ctx *c, const EncBatch &b, int n) { idwt_level(c, n, b.y_jpeg, b.y_proc, (size_t)NHW_Y_SLOT, 256, 512); }

// generic form used by the decoder (tmp: unused since the two passes became one kernel)
void idwt_rows_cols(nhw_ctx *c, int n_planes, const int16_t *in, int16_t *tmp, int16_t *out, size_t slot, int N, int stride)
{
	(void)tmp;
	idwt_level(c, n_planes, in, out, slot, N, stride);
}

static void idwt_chroma128(nhw_ctx *c, const EncBatch &b, int n) { idwt_level(c, 2 * n, b.c_jpeg, b.c_proc, (size_t)NHW_C_SLOT, 128, 256); }

static void zero_bytes(nhw_ctx *c, const EncBatch &b, int n, size_t off, size_t count)
{
	size_t units = (count + 15) / 16;
	NHW_LAUNCH(c, k_zero_bytes, dim3((unsigned)((units + 255) / 256), n), 256, 0, b.bytes, (size_t)ENC_BYTES_SLOT, off, count);
}

// ---- the luma chain at q <= 16 (nhw_encoder.c:141-2252 with the low-quality arms): the kernels that are written for
// every quality are shared with the q17..q23 chain; the stages whose fast forms only cover q17..q23, and the stages
// that exist only down here (enc_lowq.cuh), run in their row / image forms.
static void encode_luma_lowq(nhw_ctx *c, const EncBatch &b, int n, int q, int ratio)
{
	const size_t YS = NHW_Y_SLOT, CS = NHW_C_SLOT;
	if (q > 6) {   // closed loop
		run_groups(c, "y_e6a_tag", b, n, 256, 5, [=] __device__(const EncImg &im, int r, int g) { y_e6a_tag_cells(im.proc, im.ll1 + r * 256, r, g); });
		NHW_LAUNCH_L(c, "y_recons1_ll2", k_recons_ll2_wave, n, 128, LL2_SMEM_BYTES, b, q, 1);
		run_rows(c, "y_recons1_quant_rows", b, n, 256, [=] __device__(const EncImg &im, int r) { y_recons_quant_row(im, r, ratio, 1, q); });
		idwt_luma256(c, b, n);
		NHW_LAUNCH_L(c, "y_e6c_apply", k_e6c_apply, dim3(16, n), 256, 0, b);
		NHW_LAUNCH_L(c, "y_e6d_correct", k_e6d_correct, dim3(32, n), 256, 0, b);
		dwt_level_from_jpeg(c, n, b.y_jpeg, YS, b.y_proc, YS, 256, 512, nullptr, 0);
	}
	if (q <= 11) run_rows(c, "y_e7_kill", b, n, 128, [=] __device__(const EncImg &im, int r) { y_e7_kill_row(im, q, ratio, 128 + r); });
	if (q < 13) NHW_LAUNCH_L(c, "y_e8_smooth", k_e8_staged, n, 32, E8_SMEM, b, q);
	NHW_LAUNCH(c, k_copy_region, dim3(256 * 128 / 256, n), 256, 0, b.y_proc, YS, 512, b.y_ll2s, CS, 256, 256);
	NHW_LAUNCH_L(c, "y_ll2_code", k_ll2_code, n, 128, LL2_CODE_SMEM, b, q);
	if (q > 12) {   // second reconstruction
		NHW_LAUNCH_L(c, "y_recons0_ll2", k_recons_ll2_wave, n, 128, LL2_SMEM_BYTES, b, q, 0);
		run_rows(c, "y_recons0_quant_rows", b, n, 256, [=] __device__(const EncImg &im, int r) { y_recons_quant_row(im, r, ratio, 0, q); });
		NHW_LAUNCH_L(c, "y_recons0_shrink_lowq", k_recons_shrink_lowq, n, 256, 0, b);
		idwt_luma256(c, b, n);
	}
	if (q == 16) run_rows(c, "y_e14_rows", b, n, 256, [=] __device__(const EncImg &im, int r) { y_e14_threshold_row(im, q, ratio, 256 + r); });
	else if (q >= 14) run_rows(c, "y_e14_rows", b, n, 256, [=] __device__(const EncImg &im, int r) { y_e14_q14_row(im, q, ratio, 256 + r); });
	else run_image(c, "y_e14_lowq", b, n, [=] __device__(const EncImg &im, int) { y_e14_lowq_image(im, q, ratio); });
	if (q > 12) {
		NHW_LAUNCH_L(c, "y_e16_residual", k_e16_residual, n, 256, 0, b, q);
		NHW_LAUNCH_L(c, "y_e16b_classify", k_e16b_classify, n, 256, 0, b, q);
		NHW_LAUNCH_L(c, "y_e18_lists", k_e18_lists, n, 256, 0, b, q);
		NHW_LAUNCH_L(c, "y_e18_tails", k_e18_tails, dim3(3, n), 32, 0, b, q);
	}
	NHW_LAUNCH_L(c, "y_e19_restore", k_e19_restore, dim3(8, n), 256, 0, b);
	NHW_LAUNCH_L(c, "y_e20_cleanup", k_e20_bands, dim3(3, n), 256, 0, b, q, ratio);
	run_groups_inplace(c, "y_offset_mult8", b, n, 512, 6, [=] __device__(const EncImg &im, int r, int g, int *o) { return y_offset_mult8_cells(im.proc, r, g, o); });
	// offsetY's byte loop at q <= 16 (enc_lowq.cuh): rows are independent except for a three-state cycle that is never
	// restarted.  Every row is walked dry for each state it could start in, one thread per image chains the rows, then
	// the rows are walked for real.  Scratch (im.tmp1): next0[512] int16 | out-state / spill per (row, state) | in-state per row | flag
	run_rows(c, "y_offq_dry", b, n, 512, [=] __device__(const EncImg &im, int r) {
		int16_t *next0 = reinterpret_cast<int16_t *>(im.tmp1);
		uint8_t *res = im.tmp1 + 1024;
		const int nx = r < 511 ? im.proc[(r + 1) * YW] : 0;
		next0[r] = (int16_t)nx;
		for (int t = 0; t < 3; t++) {
			int spill;
			const int o = y_offset_quant_lowq_row(im.proc + r * YW, nullptr, r, ratio, nx, t, spill);
			res[r * 3 + t] = (uint8_t)(o | (spill ? 4 : 0));
		}
	});
	run_image(c, "y_offq_chain", b, n, [=] __device__(const EncImg &im, int) {
		const uint8_t *res = im.tmp1 + 1024;
		uint8_t *in_state = im.tmp1 + 4096;
		int state = 0, spilled = 0;
		for (int r = 0; r < 512; r++) { in_state[r] = (uint8_t)state; const int v = res[r * 3 + state]; spilled |= v & 4; state = v & 3; }
		in_state[512] = (uint8_t)(spilled ? 1 : 0);   // a row traded with the next row's first cell: serial form for this image
	});
	run_rows(c, "y_offq_commit", b, n, 512, [=] __device__(const EncImg &im, int r) {
		const uint8_t *in_state = im.tmp1 + 4096;
		if (in_state[512]) return;
		int spill;
		y_offset_quant_lowq_row(im.proc + r * YW, im.proc + r * YW, r, ratio, reinterpret_cast<const int16_t *>(im.tmp1)[r], in_state[r], spill);
	});
	run_image(c, "y_offq_serial", b, n, [=] __device__(const EncImg &im, int) { if (im.tmp1[4096 + 512]) y_offset_quant_lowq_image(im, ratio); });
	run_rows(c, "y_scan_strips", b, n, 128, [=] __device__(const EncImg &im, int s) { y_scan_strip(im, s); });
	NHW_LAUNCH_L(c, "y_peephole", k_peephole, n, PEEP_THREADS, 262144 / 8, b);
}

// Encode n <= max_batch images whose pixels are in device memory.  out_dev: n slots of
// NHW_MAX_STREAM_BYTES.  All work is queued on c->stream; the caller synchronises.
void encode_chunk(nhw_ctx *c, const uint8_t *rgb, int n, int q, uint8_t *out_dev, uint32_t *len_dev, int32_t *status_dev)
{
	const EncBatch b = enc_batch_of(c);
	const int ratio = 8;
	const size_t YS = NHW_Y_SLOT, CS = NHW_C_SLOT, QS = NHW_Q_SLOT;

	// ---- per-call state that must start at zero (canonical "fresh calloc" semantics)
	cudaMemsetAsync(c->enc_hdr, 0, (size_t)n * sizeof(EncHdr), c->stream);
	zero_bytes(c, b, n, OFF_TREE1 - 64, NHW_CAP_TREE1 + 64 + 64);
	zero_bytes(c, b, n, OFF_SCAN + 262144, 131072 + 64);
	zero_bytes(c, b, n, OFF_WORDS, ENC_WORDS_BYTES);
	run_image(c, "init_hdr", b, n, [=] __device__(const EncImg &im, int) { im.hdr->quality = q; });

	// ---- front end: colour, 4:2:0, pre-sharpening, two analysis levels (front.cu)
	front_fused(c, rgb, n, q, b.y_proc, YS, b.y_ll1, CS, c->c_u8, b.c_proc, CS, b.c_ll1, QS, b.y_hq, YS);

	// ---- chroma chain: independent of the luma chain from here to the entropy stage (own planes, own part of tree1,
	// of the scan buffer and of the header); on its own stream when the chunk runs alone on the GPU
	// (forked here, next to the short kernels of the luma closed loop: 43.9 -> 43.5 ms; forked before the LL2 coder it
	// competes with the latency-bound kernels and the step gets slower: 45.6 ms)
	nhw_ctx *cs = c;
	const cudaStream_t main_stream = c->stream;
	const bool side = c->chroma_side && c->chroma_stream;
	if (side) {   // the launch macros issue on c->stream: point it at the side stream for the length of the chroma chain
		cudaEventRecord(c->ev_chroma0, main_stream);
		cudaStreamWaitEvent(c->chroma_stream, c->ev_chroma0, 0);
		c->stream = c->chroma_stream;
	}
	// U and V planes side by side (nhw_encoder.c:2255-2868)
	const bool lowq = q <= 16;   // row / image forms of the stages whose cell-group forms are q17..q23 only
	if (lowq)
		run_plane_rows(cs, "c_recons1_rows", b, n, 128, [=] __device__(const EncImg &im, int r, int) {
			if (r < 64) c_recons_ll_row(im, r, 1, q);
			c_recons_quant_row(im, r, ratio, 1);
		});
	else
	run_plane_groups(cs, "c_recons1", b, n, 128, 4, [=] __device__(const EncImg &im, int r, int g, int) {
		int o[8];
		c_recons_cells(im.cproc + r * CW, r, g, ratio, 1, o);
		st8(im.cjpeg + r * CW + g * 8, o);
	});
	idwt_chroma128(cs, b, n);
	run_plane_groups(cs, "c_correct", b, n, 128, 4, [=] __device__(const EncImg &im, int r, int g, int v) {
		int o[8];
		c_correct_cells(im.cproc + r * CW, im.cll1 + r * 128, g, v, o);
		st8(im.cjpeg + r * CW + g * 8, o);
	});
	dwt_level_from_jpeg(cs, 2 * n, b.c_jpeg, CS, b.c_proc, CS, 128, 256, b.c_ll2s, QS);   // (+ the level-2 snapshot)
	if (lowq)
		run_plane_rows(cs, "c_recons0_rows", b, n, 128, [=] __device__(const EncImg &im, int r, int) {
			if (r < 64) c_recons_ll_row(im, r, 0, q);
			c_recons_quant_row(im, r, ratio, 0);
		});
	else
	run_plane_groups(cs, "c_recons0", b, n, 128, 4, [=] __device__(const EncImg &im, int r, int g, int) {
		int o[8];
		c_recons_cells(im.cproc + r * CW, r, g, ratio, 0, o);
		st8(im.cjpeg + r * CW + g * 8, o);
	});
	idwt_chroma128(cs, b, n);
	if (q >= 18) NHW_LAUNCH_L(cs, "c_residual_tags", k_c_residual_tags, dim3(8, 2 * n), 256, 0, b, q);
	NHW_LAUNCH(cs, k_copy_region, dim3(128 * 64 / 256, 2 * n), 256, 0, b.c_ll2s, QS, 128, b.c_proc, CS, 256, 128);
	if (q <= 11) run_plane(cs, "c_ll_smooth", b, n, [=] __device__(const EncImg &im, int) { c_ll_smooth_image(im); });
	NHW_LAUNCH_L(cs, "c_ll_quant", k_c_ll_quant, n, 64, 0, b, q);
	NHW_LAUNCH_L(cs, "c_quant_scan", k_c_quant_scan, dim3(16, n), 256, 0, b, ratio);

	if (side) {
		cudaEventRecord(c->ev_chroma1, c->chroma_stream);
		c->stream = main_stream;
	}

	if (lowq) encode_luma_lowq(c, b, n, q, ratio);
	else {
	// ---- luma closed loop (nhw_encoder.c:141-283)
	run_groups(c, "y_e6a_tag", b, n, 256, 5, [=] __device__(const EncImg &im, int r, int g) { y_e6a_tag_cells(im.proc, im.ll1 + r * 256, r, g); });
	NHW_LAUNCH_L(c, "y_recons1_ll2", k_recons_ll2_wave, n, 128, LL2_SMEM_BYTES, b, q, 1);
	NHW_LAUNCH_L(c, "y_recons_patterns", k_patterns, n, 256, 0, b, 0);
	run_groups(c, "y_recons1_quant", b, n, 256, 5, [=] __device__(const EncImg &im, int r, int g) { recons_quant_group(im, r, g, ratio, 1); });
	idwt_luma256(c, b, n);
	NHW_LAUNCH_L(c, "y_e6c_apply", k_e6c_apply, dim3(16, n), 256, 0, b);
	NHW_LAUNCH_L(c, "y_e6d_correct", k_e6d_correct, dim3(32, n), 256, 0, b);
	dwt_level_from_jpeg(c, n, b.y_jpeg, YS, b.y_proc, YS, 256, 512, nullptr, 0);

	// ---- LL2 coding (nhw_encoder.c:623-757)
	// (the `resIII` snapshot as a copy of its own: folded into the analysis kernel its second set of transposed 16-byte
	// stores costs more than this coalesced copy -- 0.92 vs 0.49 + 0.30 ms; the 128 x 128 chroma snapshot is folded)
	NHW_LAUNCH(c, k_copy_region, dim3(256 * 128 / 256, n), 256, 0, b.y_proc, YS, 512, b.y_ll2s, CS, 256, 256);
	// LL2 -> bytes (wavefront) and the DPCM coder (step links + chain walk), see enc_ll_par.cuh
	NHW_LAUNCH_L(c, "y_ll2_code", k_ll2_code, n, 128, LL2_CODE_SMEM, b, q);
	// (the coder works on a shared-memory copy of the band: the plane still holds what the snapshot holds)

	// ---- second reconstruction = what the decoder will see as LL1 (nhw_encoder.c:759-781)
	NHW_LAUNCH_L(c, "y_recons0_ll2", k_recons_ll2_wave, n, 128, LL2_SMEM_BYTES, b, q, 0);
	NHW_LAUNCH_L(c, "y_recons_patterns", k_patterns, n, 256, 0, b, 0);
	run_groups(c, "y_recons0_quant", b, n, 256, 5, [=] __device__(const EncImg &im, int r, int g) { recons_quant_group(im, r, g, ratio, 0); });
	run_groups(c, "y_recons0_shrink", b, n, 256, 5, [=] __device__(const EncImg &im, int r, int g) {
		int o[8];
		if (shrink_cells8(im.jpeg, r, g, 8, o)) st8(im.jpeg + r * YW + g * 8, o);
	});
	idwt_luma256(c, b, n);
	if (q > 21) NHW_LAUNCH_L(c, "y_hq_first_order", k_hq_first_order, dim3(8, 8, n), 256, 0, b);

	// ---- level-1 thresholds, pattern tags, residual side channels (nhw_encoder.c:783-1887)
	run_groups_inplace(c, "y_e14_e15_tags", b, n, 512, 6, [=] __device__(const EncImg &im, int r, int g, int *o) { return y_e14_e15_cells(im.proc + r * YW, r, g, q, ratio, o); });
	NHW_LAUNCH_L(c, "y_e16_residual", k_e16_residual, n, 256, 0, b, q);
	NHW_LAUNCH_L(c, "y_e16b_classify", k_e16b_classify, n, 256, 0, b, q);
	if (q > 21) NHW_LAUNCH_L(c, "y_hq_e17", k_hq_e17, dim3(256, n), 256, 0, b);
	NHW_LAUNCH_L(c, "y_e18_lists", k_e18_lists, n, 256, 0, b, q);
	NHW_LAUNCH_L(c, "y_e18_tails", k_e18_tails, dim3(3, n), 32, 0, b, q);

	// ---- clean-up, quantisation to bytes, scan, peephole (nhw_encoder.c:1893-2252)
	NHW_LAUNCH_L(c, "y_e19_restore", k_e19_restore, dim3(8, n), 256, 0, b);
	NHW_LAUNCH_L(c, "y_e20_cleanup", k_e20_bands, dim3(3, n), 256, 0, b, q, ratio);
	run_groups_inplace(c, "y_offset_mult8", b, n, 512, 6, [=] __device__(const EncImg &im, int r, int g, int *o) { return y_offset_mult8_cells(im.proc, r, g, o); });
	NHW_LAUNCH_L(c, "y_offset_patterns", k_patterns, n, 256, 0, b, 1);
	run_groups_inplace(c, "y_offset_pairs57", b, n, 256, 5, [=] __device__(const EncImg &im, int r, int g, int *o) { return y_offset_pairs57_cells(im.proc, r, g, o); });
	NHW_LAUNCH_L(c, "y_quant_scan", k_y_quant_scan, dim3(32, n), 256, 0, b, ratio);
	if (q > 21) {   // res6 / char_res1 / high_qsetting3 from the quantised LH1 bytes, before the peephole edits them
		NHW_LAUNCH_L(c, "y_hq_band", k_hq_band, dim3(256, n), 256, 0, b);
		NHW_LAUNCH_L(c, "y_hq_tags", k_hq_tags, dim3(256, n), 256, 0, b, q);
		NHW_LAUNCH_L(c, "y_hq_lists", k_hq_lists, n, 256, 0, b, q);
	}
	NHW_LAUNCH_L(c, "y_peephole", k_peephole, n, PEEP_THREADS, 262144 / 8, b);

	}

	if (side) cudaStreamWaitEvent(c->stream, c->ev_chroma1, 0);   // the chroma chain joins here
	// ---- LL code tail, entropy stage, container (compress_pixel.c:878-1022, 53-469)
	NHW_LAUNCH_L(c, "c_ll_code", k_c_ll_code, n, 256, 0, b);
	NHW_LAUNCH_L(c, "entropy_pack", k_entropy, n, SEG_THREADS, 262144 / 8, b);
	NHW_LAUNCH(c, k_write_stream, n, 256, 0, b, n, out_dev, len_dev, status_dev);
}

// caller-owned buffers (nhw_pack_batch_device): n may exceed max_batch -- the offsets kernel is one CTA over all n
void pack_streams_to(nhw_ctx *c, const uint8_t *slots, const uint32_t *len, int n, uint64_t *offs, uint8_t *dense)
{
	NHW_LAUNCH(c, k_stream_offsets, 1, 1024, 0, len, offs, n);
	NHW_LAUNCH(c, k_pack_streams, n, 256, 0, slots, len, offs, dense);
}

void pack_streams(nhw_ctx *c, int n)
{
	NHW_LAUNCH(c, k_stream_offsets, 1, 1024, 0, c->len_dev, c->offs_dev, n);
	NHW_LAUNCH(c, k_pack_streams, n, 256, 0, c->out_dev, c->len_dev, c->offs_dev, c->pack_dev);
}

}  // namespace nhw
