// decode.cu -- batch decoder driver: .nhw streams (device-resident, described by DecDesc) ->
// 786432 BMP pixel bytes per image.  Replaces decode_image + write_image_bmp's pixel path
// (decoder/nhw_decoder.c:54-1476, decoder/nhw_decoder_cli.c:108-291) for a whole batch.
//
// First complete version: the bit-serial prefix decode and the raster-order inline stages run
// one thread per image (they are the decoder's serial spine); the
// transforms, transposes, chroma upsampling and the colour conversion are data-parallel.
#include <stdlib.h>

#include "nhw_ctx.h"
#include "nhw_dev.cuh"
#include "dec_par.cuh"
#include "dec_parse.h"
#include "../../include/nhw_cuda.h"

namespace {

#define DEC_CMARK_CAP 8191   // chroma markers per plane kept as a list; a plane with more falls back to a sweep

// decode-time carve-up of the per-image byte slot (the encoder's layout is not live during decode)
enum : int {
	DOFF_RESCOMP = 4096,
	DOFF_BOOK = DOFF_RESCOMP + 24640,
	DOFF_BTMP = DOFF_BOOK + 4096,                  // two books (luma, chroma) of 1024 entries
	DOFF_LISTLEN = DOFF_BTMP + 4096,               // and their two scratch areas
	DOFF_FLAGS = DOFF_LISTLEN + 256,
	DOFF_LTMP = DOFF_FLAGS + 131072,
	DOFF_LISTS = DOFF_LTMP + 131072 + 256,        // 8 lists x 65536 entries
	DOFF_HQ = DOFF_LISTS + 8 * 131072,             // q22/q23: two res6 position lists, NHW_CAP_HQ_LIST x u32 each
	DOFF_MBITS = DOFF_HQ + 2 * 4 * NHW_CAP_HQ_LIST,   // luma marker bitmap: 512 rows x 16 words (written by the inverse scan)
	DOFF_CMARK = DOFF_MBITS + 512 * 64,             // chroma marker lists: 2 planes x (count + DEC_CMARK_CAP entries)
	DOFF_END = DOFF_CMARK + 2 * 4 * (1 + DEC_CMARK_CAP),
};
static_assert(DOFF_END <= NHW_DEC_BYTES_SLOT, "decode scratch must fit the per-image decoder byte slot");

struct DecBatch {
	const uint8_t *blobs;         // dense concatenation of the chunk's streams
	const uint64_t *blob_off;     // per image
	const DecDesc *desc;
	int16_t *y_proc, *y_jpeg, *y_aux, *uvcoef;
	int16_t *c_proc, *c_jpeg, *c_aux;
	uint8_t *bytes;
	uint8_t *yuv;                 // n x 3 x 262144
	int32_t *status;
	const uint16_t *lut;          // primary prefix-code table (device memory)
};

__device__ uint16_t g_dec_lut[NHW_LUT_WORDS];

__device__ __forceinline__ void unpack8(const uint4 &w, int *v)
{
	v[0] = (int16_t)(w.x & 0xffff); v[1] = (int16_t)(w.x >> 16); v[2] = (int16_t)(w.y & 0xffff); v[3] = (int16_t)(w.y >> 16);
	v[4] = (int16_t)(w.z & 0xffff); v[5] = (int16_t)(w.z >> 16); v[6] = (int16_t)(w.w & 0xffff); v[7] = (int16_t)(w.w >> 16);
}
__device__ __forceinline__ uint4 pack8(const int *v)
{
	return make_uint4((uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16), (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16),
	                  (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16), (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16));
}

__device__ __forceinline__ DecImg make_dec(const DecBatch &b, int i, int comp)
{
	DecImg im;
	im.blob = b.blobs + b.blob_off[i];
	im.d = b.desc + i;
	im.proc = b.y_proc + (size_t)i * NHW_Y_SLOT;
	im.jpeg = b.y_jpeg + (size_t)i * NHW_Y_SLOT;
	im.aux = b.y_aux + (size_t)i * NHW_Y_SLOT;
	im.uvcoef = b.uvcoef + (size_t)i * NHW_Y_SLOT;
	const size_t p = (size_t)i * 2 + comp;
	im.cproc = b.c_proc + p * NHW_C_SLOT;
	im.cjpeg = b.c_jpeg + p * NHW_C_SLOT;
	im.caux = b.c_aux + p * NHW_C_SLOT;
	uint8_t *bytes = b.bytes + (size_t)i * NHW_DEC_BYTES_SLOT;
	im.res_comp = bytes + DOFF_RESCOMP;
	im.book = reinterpret_cast<uint16_t *>(bytes + DOFF_BOOK);
	im.list_len = reinterpret_cast<int32_t *>(bytes + DOFF_LISTLEN);
	im.flags = reinterpret_cast<uint16_t *>(bytes + DOFF_FLAGS);
	for (int k = 0; k < 8; k++) im.list[k] = reinterpret_cast<uint16_t *>(bytes + DOFF_LISTS) + (size_t)k * 65536;
	im.hq_list[0] = reinterpret_cast<uint32_t *>(bytes + DOFF_HQ);
	im.hq_list[1] = im.hq_list[0] + NHW_CAP_HQ_LIST;
	im.mbits = reinterpret_cast<uint32_t *>(bytes + DOFF_MBITS);
	im.cmark = reinterpret_cast<uint32_t *>(bytes + DOFF_CMARK) + (size_t)comp * (1 + DEC_CMARK_CAP);
	im.yuv = b.yuv + (size_t)i * 786432;
	im.lut = b.lut;
	return im;
}

template <typename F>
__global__ void kd_image(DecBatch b, int n, F f)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && b.status[i] == 0) f(make_dec(b, i, 0), i);
}
// same, with the prefix-code table staged in shared memory (the bit-serial decoders hit it once per code)
template <typename F>
__global__ void __launch_bounds__(32) kd_image_lut(DecBatch b, int n, F f)
{
	__shared__ __align__(16) uint16_t slut[NHW_LUT_WORDS];
	for (int k = threadIdx.x; k < NHW_LUT_WORDS / 8; k += 32)
		reinterpret_cast<uint4 *>(slut)[k] = reinterpret_cast<const uint4 *>(b.lut)[k];
	__syncwarp();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && b.status[i] == 0) {
		DecImg im = make_dec(b, i, 0);
		im.lut = slut;
		f(im, i);
	}
}
// ---- the decoder's serial front: three independent single-thread jobs per image run side by side
// (threadIdx.y = job) instead of one after the other: luma prefix decode, chroma prefix decode,
// side-channel list expansion.  The jobs are branchy bit-serial parsers: lanes of a warp working on
// different streams diverge at almost every step and the warp pays for every path taken, so a warp only
// carries DSF_STREAMS streams (one active lane in every 32 / DSF_STREAMS); the many more warps this makes
// also hide each other's latency.  The prefix-code table is staged in shared memory.
#define DSF_STREAMS 4
__global__ void __launch_bounds__(96) kd_serial_front(DecBatch b, int n, int spw, int job_mask)
{
	__shared__ __align__(16) uint16_t slut[NHW_LUT_WORDS];
	const int tid = threadIdx.y * 32 + threadIdx.x;
	for (int k = tid; k < NHW_LUT_WORDS / 8; k += 96) reinterpret_cast<uint4 *>(slut)[k] = reinterpret_cast<const uint4 *>(b.lut)[k];
	__syncthreads();
	const int group = 32 / spw;
	if (threadIdx.x % group) return;
	const int i = blockIdx.x * spw + threadIdx.x / group, job = threadIdx.y;
	if (i >= n || b.status[i] != 0 || !((job_mask >> job) & 1)) return;   // (job_mask: timing experiments only)
	DecImg im = make_dec(b, i, 0);
	im.lut = slut;
	if (job == 0) {
		uint8_t *btmp = reinterpret_cast<uint8_t *>(im.book) + 4096;
		dec_build_book(im.blob + im.d->off_tree1, im.d->size_tree1, 3, -1, im.book, btmp);
		const int rc = dec_prefix_luma(im, im.proc, dec_build_actions(im.book, true));
		if (rc) b.status[i] = rc;
	} else if (job == 1) {
		im.book += 1024;
		uint8_t *btmp = reinterpret_cast<uint8_t *>(im.book) + 4096;
		for (int k = 0; k < 1024; k++) im.book[k] = 0;
		dec_build_book(im.blob + im.d->off_tree2, im.d->size_tree2, 128, im.d->tree_end, im.book, btmp);
		const int rc = dec_prefix_chroma(im, im.uvcoef, dec_build_actions(im.book, false));
		if (rc) b.status[i] = rc;
	} else {   // (the LL bytes, once the fourth job here, have their parallel form: kd_ll_parallel)
		dec_lists_image(im, reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(im.flags) + 131072));
		dec_hq_lists_image(im, reinterpret_cast<uint32_t *>(im.aux));
	}
}

// ---- LL bytes in parallel form (the serial statement is dec_ll_dpcm, dec_core.cuh; parse_file, decoder/nhw_decoder.c:1663-2026).
// The coder looks sequential -- a cursor over code bytes, a cursor over output bytes, every value relative to the one before --
// but each piece has the shape of a scan:
//   * where codes start: a code is one byte, or two when its first byte lies in [64, 128) (part A, the 16384 luma bytes; part
//     B, the chroma bytes, has one-byte codes only).  So position p starts a code unless p - 1 started a two-byte one: a bit
//     that is reset by every one-byte lead and toggles along a run of two-byte leads -- a map {0, 1} -> {0, 1} per segment;
//   * where a code's output goes, and which bytes of the side list (highres) it takes: sums of per-code counts;
//   * the values: a code either sets the running value or adds its deltas to it (mod 256) -- maps "x -> a" / "x -> x + d".
// One CTA per stream: a thread owns a contiguous run of code positions and walks it once per quantity; the per-thread
// aggregates are combined by one thread (256 entries).  Part A ends at the first code whose output cursor has reached 16384;
// part B starts on the byte after it with an absolute value.  Outputs a code writes beyond its part's end are dropped (the
// serial form overwrites or never reads them).  A stream that ends early leaves the rest of the bytes as they were, as the
// serial form does.
struct LlCode { int R, nd, d[3], nabs, a[2]; };   // R copies of the running value, then nd deltas; or nabs absolute values

__device__ __forceinline__ void ll_triple(int c0, int c1, LlCode &k)
{
	k.nd = 3;
	k.d[0] = (((c0 >> 1) & 31) << 1) - 32;
	k.d[1] = ((((c0 & 1) << 3) | (c1 >> 5)) << 1) - 16;
	k.d[2] = ((c1 & 31) << 1) - 32;
}
// part A (luma LL bytes): code byte c, the byte after it c1 (used by the two-byte codes only), the stream's coder mode
__device__ __forceinline__ LlCode ll_code_a(int c, int c1, int mode, bool side, int hr)
{
	LlCode k = {0, 0, {0, 0, 0}, 0, {0, 0}};
	if (c >= 128) {
		if (side) k.a[k.nabs++] = hr;
		k.a[k.nabs++] = (c - 128) << 1;
	} else if (c >= 64) ll_triple(c - 64, c1, k);
	else if (mode == 0) {
		if (c < 16) {
			k.R = ((c >> 3) & 1) + 2;
			const int t = c & 7;
			if (t == 1) { k.nd = 1; k.d[0] = 2; }
			else if (t == 2) { k.nd = 2; k.d[0] = 2; k.d[1] = -2; }
			else if (t == 3) { k.nd = 2; k.d[0] = 2; k.d[1] = 0; }
			else if (t == 4) { k.nd = 2; k.d[0] = -2; k.d[1] = 2; }
			else if (t == 5) { k.nd = 2; k.d[0] = -2; k.d[1] = 0; }
			else if (t == 6) { k.nd = 1; k.d[0] = -2; }
			else if (t == 7) { k.nd = 1; k.d[0] = 4; }
		} else if (c < 32) { k.nd = 2; k.d[0] = c >= 24 ? 4 : 2; k.d[1] = ((c & 7) << 1) - 8; }
		else { const int x = c - 32; k.nd = 2; k.d[0] = ((x >> 3) << 1) - 6; k.d[1] = ((x & 7) << 1) - 8; }
	} else if (mode == 1) {
		if (c < 32) {
			k.R = ((c >> 2) & 7) + 2;
			const int t = c & 3;
			if (t) { k.nd = 1; k.d[0] = t == 1 ? 2 : t == 2 ? -2 : 0; }
		} else { const int x = c - 32; k.nd = 2; k.d[0] = ((x >> 3) << 1) - 4; k.d[1] = ((x & 7) << 1) - 8; }
	} else k.R = (c & 63) + 2;
	return k;
}
// part B (chroma LL bytes)
__device__ __forceinline__ LlCode ll_code_b(int c)
{
	LlCode k = {0, 0, {0, 0, 0}, 0, {0, 0}};
	if (c >= 192) {
		const int x = c - 192, t = x >> 2, m = x & 3;
		k.nd = 3;
		k.d[0] = t < 2 ? 0 : (t == 2 || t == 4 || t == 5) ? 4 : -4;
		k.d[1] = (t == 0 || t == 4 || t == 6) ? 4 : (t == 1 || t == 5 || t == 7) ? -4 : 0;
		k.d[2] = m == 0 ? 0 : m == 1 ? 4 : m == 2 ? -4 : 8;
	} else if (c >= 128) { k.nabs = 1; k.a[0] = (c - 128) << 2; }
	else if (c >= 64) {
		const int run = (c >> 3) & 7, t = c & 7;
		if (run == 7) k.R = t + 7 + 2;
		else {
			k.R = run + 2;
			if (t == 1) { k.nd = 1; k.d[0] = 4; }
			else if (t == 2) { k.nd = 2; k.d[0] = 4; k.d[1] = -4; }
			else if (t == 3) { k.nd = 3; k.d[0] = 4; k.d[1] = -4; k.d[2] = 0; }
			else if (t == 4) { k.nd = 3; k.d[0] = -4; k.d[1] = 4; k.d[2] = 0; }
			else if (t == 5) { k.nd = 2; k.d[0] = -4; k.d[1] = 4; }
			else if (t == 6) { k.nd = 1; k.d[0] = -4; }
			else if (t == 7) { k.nd = 1; k.d[0] = 8; }
		}
	} else { k.nd = 2; k.d[0] = ((c >> 3) << 2) - 16; k.d[1] = ((c & 7) << 2) - 16; }
	return k;
}
// value maps: bit 8 set = "x -> low byte", clear = "x -> x + low byte"; then(f, g) = g after f
__device__ __forceinline__ int ll_map_of(const LlCode &k)
{
	if (k.nabs) return 256 | (k.a[k.nabs - 1] & 255);
	return (k.d[0] + k.d[1] + k.d[2]) & 255;
}
__device__ __forceinline__ int ll_map_then(int f, int g) { return (g & 256) ? g : ((f & 256) | ((f + g) & 255)); }
__device__ __forceinline__ int ll_map_apply(int f, int x) { return (f & 256) ? (f & 255) : ((x + f) & 255); }

// Per code byte, what the code does, as a table entry built once per CTA (the code tables depend on the stream's coder mode
// and quality only): R | nd << 7 | nabs << 9 | d0 << 11 | d1 << 18 | d2 << 25 (deltas as 7-bit two's complement), and
// the value map next to it.  The two-byte codes of part A take their deltas from both bytes and are worked out on the spot.
#define LLP_THREADS 256
__device__ __forceinline__ uint32_t ll_pack(const LlCode &k)
{
	return (uint32_t)k.R | ((uint32_t)k.nd << 7) | ((uint32_t)k.nabs << 9) | ((uint32_t)(k.d[0] & 127) << 11) |
	       ((uint32_t)(k.d[1] & 127) << 18) | ((uint32_t)(k.d[2] & 127) << 25);
}
__device__ __forceinline__ int ll_d0(uint32_t e) { return (int)(e << 14) >> 25; }
__device__ __forceinline__ int ll_d1(uint32_t e) { return (int)(e << 7) >> 25; }
__device__ __forceinline__ int ll_d2(uint32_t e) { return (int)e >> 25; }
__device__ __forceinline__ int ll_nout(uint32_t e) { return (int)(e & 127u) + (int)((e >> 7) & 3u) + (int)((e >> 9) & 3u); }

__global__ void __launch_bounds__(LLP_THREADS) kd_ll_parallel(DecBatch b)
{
	__shared__ int s_a[LLP_THREADS + 1], s_b[LLP_THREADS + 1], s_c[LLP_THREADS + 1];
	__shared__ uint32_t tabA[256], tabB[256];
	__shared__ uint16_t mapA[256], mapB[256];
	__shared__ int s_iB, s_ok, s_last;
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	const DecDesc *d = im.d;
	const uint8_t *ch = im.blob + d->off_ch_res, *hr = im.blob + d->off_highres;
	const uint8_t *ub = im.blob + d->off_u64, *vb = im.blob + d->off_v64;
	const int ch_len = d->end_ch_res, hr_len = d->highres_comp_len, q = d->quality, mode = d->byte0 & 3;
	const bool side = q > 15;
	uint8_t *o = im.res_comp;
	const int t = threadIdx.x;
	if (ch_len < 3) return;
	{
		const LlCode ka = ll_code_a(t, 0, mode, side, 0), kb = ll_code_b(t);
		tabA[t] = ll_pack(ka);
		mapA[t] = (uint16_t)ll_map_of(ka);
		tabB[t] = ll_pack(kb);
		mapB[t] = (uint16_t)ll_map_of(kb);
	}
	// the code at position p of part A as (entry, map); c1 = the byte after it
	auto code_a = [&](int c, int c1, uint32_t &e, int &m) {
		e = tabA[c];
		m = mapA[c];
		if (c >= 64 && c < 128) {
			LlCode k = {0, 0, {0, 0, 0}, 0, {0, 0}};
			ll_triple(c - 64, c1, k);
			e = ll_pack(k);
			m = ll_map_of(k);
		}
	};
	// ================= part A: code positions 1 .. ch_len - 2 (every code may take a second byte) =================
	const int nposA = ch_len - 2;                         // positions p = 1 + k, k in [0, nposA)
	const int segA = (nposA + LLP_THREADS - 1) / LLP_THREADS;
	const int p0 = 1 + t * segA, p1 = min(1 + (t + 1) * segA, 1 + nposA);
	// ---- 1. which positions start a code: start[p + 1] = !(start[p] && two(p)); per segment, the map of that bit
	int f0 = 0, f1 = 1;                                   // the bit at p1, given the bit at p0 was 0 / 1
	for (int p = p0; p < p1; p++) {
		const int c = ch[p];
		const bool tw = c >= 64 && c < 128;
		f0 = !(f0 && tw);
		f1 = !(f1 && tw);
	}
	s_a[t] = f0 | (f1 << 1);
	__syncthreads();
	if (t == 0) {
		int st = 1;                                        // position 1 starts a code
		for (int k = 0; k < LLP_THREADS; k++) { const int m = s_a[k]; s_a[k] = st; st = (m >> st) & 1; }
		s_last = st;                                       // does position ch_len - 1 start a code?
		s_iB = 0x7fffffff;
		s_ok = 1;
	}
	__syncthreads();
	const int start0 = s_a[t];
	__syncthreads();
	// ---- 2. output and side-list cursors, and the segment's value map (a map does not depend on the side-list byte a code
	// takes: such a code ends on an absolute value of its own).  Maps of segments beyond the end of part A are never used.
	int nout = 0, nhr = 0, fmap = 0;                       // fmap: identity, x -> x + 0
	{
		int st = start0;
		for (int p = p0; p < p1; p++) {
			const int c = ch[p];
			if (st) {
				uint32_t e;
				int m;
				code_a(c, ch[p + 1], e, m);
				nout += ll_nout(e);
				nhr += (c >= 128 && side) ? 1 : 0;
				fmap = ll_map_then(fmap, m);
			}
			st = !(st && c >= 64 && c < 128);
		}
	}
	s_a[t] = nout;
	s_b[t] = nhr;
	s_c[t] = fmap;
	__syncthreads();
	if (t == 0) {
		int j = 1, a = 0;
		for (int k = 0; k < LLP_THREADS; k++) { const int x = s_a[k], y = s_b[k]; s_a[k] = j; s_b[k] = a; j += x; a += y; }
		s_a[LLP_THREADS] = j;
	} else if (t == 32) {
		int x = ch[0];
		for (int k = 0; k < LLP_THREADS; k++) { const int f = s_c[k]; s_c[k] = x; x = ll_map_apply(f, x); }
		o[0] = ch[0];
	}
	__syncthreads();
	const int j0 = s_a[t], a0 = s_b[t];
	// ---- 3. the bytes of part A.  The first code whose output cursor has reached 16384 ends the part (it is not executed).
	{
		int st = start0, j = j0, a = a0, last = s_c[t];
		auto put = [&](int v) { last = v & 255; if (j < 16384) o[j] = (uint8_t)last; j++; };
		for (int p = p0; p < p1; p++) {
			const int c = ch[p];
			if (st) {
				if (j >= 16384) { atomicMin(&s_iB, p); break; }
				uint32_t e;
				int m;
				code_a(c, ch[p + 1], e, m);
				const int nd = (int)((e >> 7) & 3u);
				if (c >= 128) {
					if (side) {
						if (a >= hr_len) { s_ok = 0; break; }          // side list exhausted: the serial form stops here
						put(hr[a++]);
					}
					put((c - 128) << 1);
				} else {
					for (int r = (int)(e & 127u); r > 0; r--) put(last);
					if (nd > 0) put(last + ll_d0(e));
					if (nd > 1) put(last + ll_d1(e));
					if (nd > 2) put(last + ll_d2(e));
				}
			}
			st = !(st && c >= 64 && c < 128);
		}
	}
	__syncthreads();
	// ================= part B: one absolute byte at iB, then one-byte codes =================
	int iB = s_iB;
	if (iB == 0x7fffffff) {
		// no code of the scanned range saw the cursor at 16384: the very last code can still have carried it there, with
		// part B starting at ch_len - 1 if that position starts a code; anything else is a short stream
		if (s_a[LLP_THREADS] >= 16384 && s_last) iB = ch_len - 1;
		else return;
	}
	if (!s_ok || iB >= ch_len) return;
	const int nposB = ch_len - (iB + 1);                   // code positions iB + 1 .. ch_len - 1
	const int segB = (nposB + LLP_THREADS - 1) / LLP_THREADS;
	const int q0 = iB + 1 + t * segB, q1 = min(iB + 1 + (t + 1) * segB, ch_len);
	nout = 0;
	fmap = 0;
	for (int p = q0; p < q1; p++) {
		const int c = ch[p];
		nout += ll_nout(tabB[c]);
		fmap = ll_map_then(fmap, mapB[c]);
	}
	__syncthreads();
	s_a[t] = nout;
	s_c[t] = fmap;
	__syncthreads();
	auto lsb = [&](int j) {   // res_U_64 / res_V_64 LSB planes (nhw_decoder.c:1983-2026)
		if (!side) return 0;
		const int k = (j - 16384) & 4095;
		const uint8_t *pl = j < 20480 ? ub : vb;
		return ((pl[k >> 3] >> (7 - (k & 7))) & 1) << 1;
	};
	if (t == 0) {
		int j = 16385;
		for (int k = 0; k < LLP_THREADS; k++) { const int x = s_a[k]; s_a[k] = j; j += x; }
	} else if (t == 32) {
		int x = ch[iB];
		o[16384] = (uint8_t)(x + lsb(16384));
		for (int k = 0; k < LLP_THREADS; k++) { const int f = s_c[k]; s_c[k] = x; x = ll_map_apply(f, x); }
	}
	__syncthreads();
	{
		int j = s_a[t], last = s_c[t];
		auto put = [&](int v) { last = v & 255; if (j < 24576) o[j] = (uint8_t)(last + lsb(j)); j++; };
		for (int p = q0; p < q1 && j < 24576; p++) {
			const int c = ch[p];
			const uint32_t e = tabB[c];
			const int nd = (int)((e >> 7) & 3u);
			if ((e >> 9) & 3u) put((c - 128) << 2);
			else {
				for (int r = (int)(e & 127u); r > 0; r--) put(last);
				if (nd > 0) put(last + ll_d0(e));
				if (nd > 1) put(last + ll_d1(e));
				if (nd > 2) put(last + ll_d2(e));
			}
		}
	}
}

// ---- chroma markers 5003..5006 (dec_c_markers_image): every marker only ADDS to cells of the reconstructed LL, so
// they commute (atomic adds).  The inverse scan has already collected them (cmark list) and cleared their cells; a
// plane with more markers than the list holds is swept instead.  One CTA per plane.
__device__ __forceinline__ void c_marker_add(int16_t *P, int s, int code)
{
	const int r = s >> 8, j = s & 255;
	int t = s;
	if (r < 128) t -= 128;
	else t -= 32768 + (j < 128 ? 0 : 128);
	if (code == 5) { atomic_add_s16(P + t, -4); atomic_add_s16(P + t + 1, -4); }
	else if (code == 6) { atomic_add_s16(P + t, 4); atomic_add_s16(P + t + 1, 4); }
	else if (code == 3) atomic_add_s16(P + t, -6);
	else if (code == 4) atomic_add_s16(P + t, 6);
}
__global__ void __launch_bounds__(256) kd_c_markers(DecBatch b)
{
	const int img = blockIdx.x >> 1;
	if (b.status[img] != 0) return;
	const DecImg im = make_dec(b, img, blockIdx.x & 1);
	const uint32_t n = im.cmark[0];
	for (uint32_t k = threadIdx.x; k < n && k < DEC_CMARK_CAP; k += 256) {
		const uint32_t e = im.cmark[1 + k];
		c_marker_add(im.cproc, (int)(e & 0xffffffu), (int)(e >> 24));
	}
	if (n <= DEC_CMARK_CAP) return;
	for (int s = threadIdx.x; s < 65536; s += 256) {   // the markers that did not fit stayed in the plane
		if ((s >> 8) < 128 && (s & 255) < 128) continue;
		const int v = im.cjpeg[s];
		if (v > 5000) { c_marker_add(im.cproc, s, v - 5000); im.cjpeg[s] = 0; }
	}
}

// ---- chroma LL fill + exw overrides (dec_c_ll_image), one CTA per image: the fills in parallel, the short override
// lists (U then V, they share a cursor) by one thread
__global__ void __launch_bounds__(256) kd_c_ll(DecBatch b)
{
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im0 = make_dec(b, blockIdx.x, 0), im1 = make_dec(b, blockIdx.x, 1);
	const int bias = im0.d->quality > 15 ? 0 : 1;
	for (int i = threadIdx.x; i < 8192; i += 256) {
		const DecImg &im = i < 4096 ? im0 : im1;
		const int k = i & 4095;
		im.cjpeg[(k >> 6) * CW + (k & 63)] = (int16_t)(im.res_comp[16384 + i] + bias);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		int exw = im0.list_len[10];
		exw = dec_c_ll_overrides(im0, exw);
		dec_c_ll_overrides(im1, exw);
	}
}

template <typename F>
__global__ void kd_plane(DecBatch b, int n2, F f)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n2 && b.status[p >> 1] == 0) f(make_dec(b, p >> 1, p & 1), p & 1);
}
template <typename F>
__global__ void kd_rows(DecBatch b, int rows, F f)   // grid (ceil(rows/64), n)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < rows && b.status[blockIdx.y] == 0) f(make_dec(b, blockIdx.y, 0), r);
}
template <typename F>
__global__ void kd_plane_rows(DecBatch b, int rows, F f)   // grid (ceil(rows/64), 2n)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < rows && b.status[blockIdx.y >> 1] == 0) f(make_dec(b, blockIdx.y >> 1, blockIdx.y & 1), r, blockIdx.y & 1);
}

template <typename F> void d_image(nhw_ctx *c, const char *l, const DecBatch &b, int n, F f) { NHW_LAUNCH_L(c, l, kd_image, (n + 31) / 32, 32, 0, b, n, f); }
template <typename F> void d_image_lut(nhw_ctx *c, const char *l, const DecBatch &b, int n, F f) { NHW_LAUNCH_L(c, l, kd_image_lut, (n + 31) / 32, 32, 0, b, n, f); }
template <typename F> void d_plane(nhw_ctx *c, const char *l, const DecBatch &b, int n, F f) { NHW_LAUNCH_L(c, l, kd_plane, (2 * n + 31) / 32, 32, 0, b, 2 * n, f); }
template <typename F> void d_rows(nhw_ctx *c, const char *l, const DecBatch &b, int n, int rows, F f) { NHW_LAUNCH_L(c, l, kd_rows, dim3((rows + 63) / 64, n), 64, 0, b, rows, f); }
template <typename F> void d_plane_rows(nhw_ctx *c, const char *l, const DecBatch &b, int n, int rows, F f) { NHW_LAUNCH_L(c, l, kd_plane_rows, dim3((rows + 63) / 64, 2 * n), 64, 0, b, rows, f); }

// ---- D8: isolated-coefficient shrink of the level-2 region, order-free form (cells8.cuh: shrink_cells8)
__global__ void __launch_bounds__(256) kd_shrink_y(DecBatch b)
{
	if (b.status[blockIdx.y] != 0) return;
	const DecImg im = make_dec(b, blockIdx.y, 0);
	if (im.d->quality <= 16) return;   // kd_shrink_y_lowq
	const int idx = blockIdx.x * 256 + threadIdx.x, r = idx >> 5, g = idx & 31;
	int o[8];
	if (shrink_cells8(im.jpeg, r, g, 9, o)) st8(im.jpeg + r * YW + g * 8, o);
}

// q <= 16 (nhw_decoder.c:660-684): diagonal neighbours only block the shrink above 16, so a diagonal neighbour of
// exactly +-17 that shrinks first changes the outcome: the rows above must be final, the rows below untouched (the
// four orthogonal neighbours still cannot change while this cell is a candidate).  One CTA per image, thread =
// column, one barrier per row.
__global__ void __launch_bounds__(256) kd_shrink_y_lowq(DecBatch b)
{
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	if (im.d->quality > 16) return;
	const int j = threadIdx.x;
	for (int r = 1; r < 255; r++) {
		if (j >= 1 && j < 255) dec_shrink_lowq_cell(im.jpeg, r, j);
		__syncthreads();
	}
}

// ---- D4: marker expansion + right-half nudges in parallel form (dec_par.cuh).  One CTA per image.
// Sweep rows ("slots"): 0..255 = rows 0..255 (all 512 columns), 256..511 = rows 256..511 left half,
// 512..767 = rows 256..511 right half.  Candidates are compacted in sweep order (count, prefix, write),
// one thread applies them; the snapshot and the candidate list live in the scratch plane (im.aux).
#define MK_CAP 32768
__global__ void __launch_bounds__(256) kd_y_markers(DecBatch b)
{
	__shared__ int cnt[768 + 1];
	__shared__ uint32_t W[2048], A[2048];
	__shared__ int first_q, fallback;
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	int16_t *J = im.jpeg, *S = im.aux;
	int *cand = reinterpret_cast<int *>(im.aux);   // rows 0..127 of the scratch plane (the snapshot uses rows >= 255)
	const uint32_t *mb = im.mbits;                 // marker bitmap from the inverse scan: 16 words per row
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	// slot -> (first bitmap word, words, first cell)
	auto slot_words = [&](int slot, int &w0, int &nw, int &base) {
		if (slot < 256) { w0 = slot * 16; nw = 16; base = slot * YW; }
		else if (slot < 512) { w0 = slot * 16; nw = 8; base = slot * YW; }
		else { w0 = (slot - 256) * 16 + 8; nw = 8; base = (slot - 256) * YW + 256; }
	};
	for (int k = tid; k < 2048; k += 256) { W[k] = 0; A[k] = 0; }
	if (tid == 0) { first_q = 1 << 30; fallback = 0; }
	for (int slot = tid; slot < 768; slot += 256) {
		int w0, nw, base, c = 0;
		slot_words(slot, w0, nw, base);
		for (int k = 0; k < nw; k++) c += __popc(mb[w0 + k]);
		cnt[slot] = c;
	}
	__syncthreads();
	if (warp == 0) {   // exclusive prefix over the 768 slots: 24 per lane
		int run = 0;
		for (int k = 0; k < 24; k++) run += cnt[lane * 24 + k];
		int inc = run;
		for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
		int acc = inc - run;
		for (int k = 0; k < 24; k++) { const int v = cnt[lane * 24 + k]; cnt[lane * 24 + k] = acc; acc += v; }
		if (lane == 31) {
			cnt[768] = inc;
			if (inc > MK_CAP) { fallback = 1; dec_y_markers_image(im); }   // never seen; keeps the stage total
		}
	}
	__syncthreads();
	if (fallback) return;
	for (int slot = tid; slot < 768; slot += 256) {
		int w0, nw, base, o = cnt[slot];
		slot_words(slot, w0, nw, base);
		for (int k = 0; k < nw; k++)
			for (uint32_t m = mb[w0 + k]; m; m &= m - 1) cand[o++] = base + 32 * k + __ffs(m) - 1;
	}
	__syncthreads();
	const int n1 = cnt[256], n12 = cnt[512], n3 = cnt[768];
	// Applying the markers.  A marker writes constants into a footprint of at most 3 columns x 2 rows around itself
	// (dec_marker_apply), so two markers interact only when they sit within two columns and one row of each other; such
	// neighbours are rare.  A marker with no other marker bit in that window is applied by its own thread, whenever; the
	// others are flagged and applied by one thread in sweep order afterwards.  (The three sweeps stay one after the other:
	// a sweep may overwrite marker cells of the next.)
	uint8_t *defer = reinterpret_cast<uint8_t *>(im.aux + 128 * YW);     // one flag per candidate (rows 128.. of the scratch plane)
	auto crowded = [&](int s) {
		const int r = s >> 9, c = s & 511;
		// footprints that wrap around a row end (the plane is addressed flat) or reach column 255 from the right half meet
		// writers far outside the window: those markers always take the ordered path
		if (c <= 1 || c >= 510 || c == 256) return true;
		for (int rr = max(r - 1, 0); rr <= min(r + 1, 511); rr++)
			for (int cc = max(c - 2, 0); cc <= min(c + 2, 511); cc++)
				if ((rr != r || cc != c) && ((mb[rr * 16 + (cc >> 5)] >> (cc & 31)) & 1u)) return true;
		return false;
	};
	auto sweep = [&](int a, int bnd, bool lower, uint32_t *Wm, uint32_t *Am) {
		for (int k = a + tid; k < bnd; k += 256) {
			const int s = cand[k];
			const bool d = crowded(s);
			defer[k] = d ? 1 : 0;
			if (!d) dec_marker_apply<true>(J, s, lower, Wm, Am);
		}
		__syncthreads();
		if (tid == 0)
			for (int k = a; k < bnd; k++)
				if (defer[k]) dec_marker_apply(J, cand[k], lower, Wm, Am);
		__syncthreads();
	};
	sweep(0, n1, false, nullptr, nullptr);
	sweep(n1, n12, true, nullptr, nullptr);
	if (im.d->quality >= 23 && n3 == n12) return;   // no right-half markers and no nudges (the rule is off at q23)
	// snapshot of rows 255..511, columns 256..511, before the right-half markers
	for (int idx = tid; idx < 257 * 32; idx += 256) {
		const int r = 255 + (idx >> 5), c8 = (idx & 31) * 8;
		*reinterpret_cast<uint4 *>(S + r * YW + 256 + c8) = *reinterpret_cast<const uint4 *>(J + r * YW + 256 + c8);
	}
	__syncthreads();
	sweep(n12, n3, true, W, A);
	if (im.d->quality >= 23) return;   // the nudge rule is off at q23 (nhw_decoder.c:588)
	// One pass: every qualifying cell is nudged on its own count.  The reference's count variable is stale at its first
	// use, so the FIRST qualifying cell (raster order) gets that stale value on top: found with an atomic min here and
	// looked at again by one thread afterwards (it can only turn a "no" into a "yes").
	for (int r = 256 + warp; r < 512; r += 8)
		for (int j = 257 + lane; j < 511; j += 32) {
			const int s = r * YW + j, k = ((r - 256) << 8) + (j - 256);
			if (!dec_dense_qualifies(S, A, s)) continue;
			atomicMin(&first_q, s);
			const int c = dec_dense_count(J, S, s);
			if (c >= 2 && !((W[k >> 5] >> (k & 31)) & 1u)) J[s] += S[s] > 0 ? 1 : -1;
		}
	__syncthreads();
	if (tid == 0 && first_q < (1 << 30)) {
		// (a nudged cell keeps |v| >= 9, so the neighbours' "< 8" tests read the same before and after any nudge)
		const int s = first_q, k = (((s >> 9) - 256) << 8) + ((s & 511) - 256);
		const int c = dec_dense_count(J, S, s);
		if (c < 2 && c + im.list_len[8] >= 2 && !((W[k >> 5] >> (k & 31)) & 1u)) J[s] += S[s] > 0 ? 1 : -1;
	}
}

// ---- D5-D7: LL2 fill in parallel, then the two short override lists
__global__ void __launch_bounds__(256) kd_y_ll(DecBatch b)
{
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	for (int i = threadIdx.x; i < 16384; i += 256) im.jpeg[(i >> 7) * YW + (i & 127)] = im.res_comp[i];
	__syncthreads();
	if (threadIdx.x >= 32) return;
	// res4 parity restore (dec_y_ll_overrides' first loop) by one warp: an entry's row is the number of row-ending codes
	// (>= 128) before it -- a ballot count -- and what it does to its four cells ("make it odd") does not depend on the
	// order of the entries.  The serial form spent ~1 us per entry waiting for its own read-modify-writes.
	const DecDesc *d = im.d;
	const int lane = threadIdx.x;
	if (d->quality > 17) {
		const uint8_t *r4 = im.blob + d->off_res4;
		int16_t *J = im.jpeg;
		int base = 0;
		for (int i0 = 0; i0 < d->res4_len && base < 128; i0 += 32) {
			const int i = i0 + lane;
			const int v = i < d->res4_len ? r4[i] : 0;
			const uint32_t ends = __ballot_sync(0xffffffffu, v >= 128);
			const int count = base + __popc(ends & ((1u << lane) - 1u));
			if (i < d->res4_len && count < 128 && v != 128) {
				const int col = v > 128 ? v - 129 : v - 1;
				if (col >= 0 && col <= 124) {
					const int e = (count << 9) + col;
					for (int k = 0; k < 4; k++)
						if (!(J[e + k] & 1)) J[e + k]++;
				}
			}
			base += __popc(ends);
		}
		__syncwarp();
	}
	if (lane == 0) im.list_len[10] = dec_y_ll_overrides(im, true);
}

// ---- D14: conditional 5-tap smoothing at the flagged positions (dec_y_smooth_flags_plane).  The targets sit on even
// rows of the half-synthesised plane, so the only neighbour of a target that can itself be a target is the one on its
// left in the same row: a run of horizontally adjacent flags is a chain, everything else is independent.  One CTA
// per image, a thread per run head.
__global__ void __launch_bounds__(256) kd_smooth_flags(DecBatch b)
{
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	int16_t *J = im.aux;
	const int n = im.list_len[9];
	for (int i = threadIdx.x; i < n; i += 256) {
		const int f = im.flags[i];
		if (i > 0 && im.flags[i - 1] == f - 1 && (f & 255) != 0) continue;   // not a run head
		for (int k = i, g = f;;) {
			const int s = ((g >> 8) << 10) + (g & 255);
			const int res = dec_lap8(J, s, YW);
			if (nhw_iabs(res) < 116) J[s] = (int16_t)(((J[s] << 2) + J[s - 1] + J[s + 1] + J[s - YW] + J[s + YW] + 4) >> 3);
			if (++k >= n || im.flags[k] != g + 1 || ((g + 1) & 255) == 0) break;
			g++;
		}
	}
}

// ---- D10: residual add-backs (nhw_decoder.c:721-787): commutative += / -= at listed positions
__global__ void __launch_bounds__(256) kd_addbacks(DecBatch b)
{
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	int16_t *P = im.proc;
	const int q = im.d->quality, t = threadIdx.x;
	auto at = [](uint16_t v) { return ((v & 65280) << 1) + (v & 255); };
	if (q >= 21) {
		for (int i = t; i < im.list_len[2]; i += 256) atomic_add_s16(P + at(im.list[2][i]), -3);
		for (int i = t; i < im.list_len[3]; i += 256) atomic_add_s16(P + at(im.list[3][i]), 3);
	}
	if (q > 12) {
		const int e = q >= 18 ? 5 : q >= 15 ? 7 : 9;
		for (int i = t; i < im.list_len[0]; i += 256) atomic_add_s16(P + at(im.list[0][i]), -e);
		for (int i = t; i < im.list_len[1]; i += 256) atomic_add_s16(P + at(im.list[1][i]), e);
	}
	if (q >= 19) {
		for (int i = t; i < im.list_len[5]; i += 256) { const int a = at(im.list[5][i]); atomic_add_s16(P + a, -4); atomic_add_s16(P + a + YW, -3); }
		for (int i = t; i < im.list_len[4]; i += 256) { const int a = at(im.list[4][i]); atomic_add_s16(P + a, 4); atomic_add_s16(P + a + YW, 3); }
		for (int i = t; i < im.list_len[6]; i += 256) { const int a = at(im.list[6][i]); atomic_add_s16(P + a, 2); atomic_add_s16(P + a + YW, 2); atomic_add_s16(P + a + 2 * YW, 2); }
		for (int i = t; i < im.list_len[7]; i += 256) { const int a = at(im.list[7][i]); atomic_add_s16(P + a, -2); atomic_add_s16(P + a + YW, -2); atomic_add_s16(P + a + 2 * YW, -2); }
	}
}

// ---- q22/q23 corrections between the two halves of the level-1 synthesis (dec_hq_addback)
// The positions index the half-synthesised plane row-major (row = band row k, column = output sample j); the plane
// they are applied to is kd_inv_rows_t's output, which is that plane transposed.
__global__ void __launch_bounds__(256) kd_hq_addbacks(DecBatch b)
{
	if (b.status[blockIdx.x] != 0) return;
	const DecImg im = make_dec(b, blockIdx.x, 0);
	const int n = dec_hq_addback_count(im);
	for (int k = threadIdx.x; k < n; k += 256) {
		int pos, amount;
		if (dec_hq_addback(im, k, pos, amount) && pos >= 0 && pos < 512 * 512) atomic_add_s16(im.aux + ((pos & 511) << 9) + (pos >> 9), amount);
	}
}

// ---- D13: first half of the level-1 luma synthesis (wavelet_synthesis2, decoder/wavelet_filterbank.c:237-295) with both
// of its transposes folded in.  Band row k of the reference's transposed plane is: low = column k of the reconstructed
// LL1 (k < 256; natural orientation in y_proc) or the band cells themselves (k >= 256), high = the band cells right of
// it (y_jpeg).  upfilter53I + upfilter53III along the row give 512 samples, which the reference then transposes: the
// output goes out transposed, O[j][k].  A CTA takes 32 band rows: the LL1 columns come in through a transposing
// shared-memory tile (64-byte row segments), the results leave through another one (64-byte column segments).
#define IRT_ROWS 32
#define IRT_LS 258      // row strides of the two tiles, in int16: both are 1 (mod 32) in 32-bit words, so the transposing
#define IRT_OS 514      // accesses (8 rows apart per lane group) fall into distinct banks
#define IRT_SMEM ((IRT_ROWS * IRT_LS + IRT_ROWS * IRT_OS) * 2)
__global__ void __launch_bounds__(256) kd_inv_rows_t(DecBatch b)
{
	extern __shared__ __align__(16) int16_t irt[];
	int16_t *tl = irt, *to = irt + IRT_ROWS * IRT_LS;
	const int img = blockIdx.y, k0 = blockIdx.x * IRT_ROWS, tid = threadIdx.x;
	if (b.status[img] != 0) return;
	const int16_t *P = b.y_proc + (size_t)img * NHW_Y_SLOT, *J = b.y_jpeg + (size_t)img * NHW_Y_SLOT;
	int16_t *O = b.y_aux + (size_t)img * NHW_Y_SLOT;
	if (k0 < 256) {
		for (int idx = tid; idx < 256 * (IRT_ROWS / 8); idx += 256) {
			const int t = idx / (IRT_ROWS / 8), c = idx % (IRT_ROWS / 8);
			int v[8];
			ld8(P + t * YW + k0 + 8 * c, v);
#pragma unroll
			for (int i = 0; i < 8; i++) tl[(8 * c + i) * IRT_LS + t] = (int16_t)v[i];
		}
		__syncthreads();
	}
	// Two band rows at a time, 128 threads each; a thread makes pairs 2u and 2u + 1 from ONE 32-bit load per band (its
	// own two cells; the neighbours come from the lanes next to it, the warp's edge cells from a second, scalar load).
	// The first form read every cell it needed with 2-byte loads of its own -- three per pair, one row at a time: 80 % of
	// the kernel's stall samples were waits for them.
	{
		const int half = tid >> 7, u = tid & 127, lane = tid & 31, t0 = 2 * u;
		for (int kk = half; kk < IRT_ROWS; kk += 2) {
			const int16_t *row = J + (k0 + kk) * YW;
			const uint32_t hw = *reinterpret_cast<const uint32_t *>(row + 256 + t0);
			const int h0 = (int16_t)(hw & 0xffffu), h1 = (int16_t)(hw >> 16);
			int hm = __shfl_up_sync(0xffffffffu, h1, 1), hp = __shfl_down_sync(0xffffffffu, h0, 1);
			if (lane == 0) hm = u ? (int)row[256 + t0 - 1] : h0;              // h[-1] = h[0]
			if (lane == 31) hp = u < 127 ? (int)row[256 + t0 + 2] : h1;        // h[256] = h[255]
			int l0, l1, l2;
			if (k0 < 256) {
				const int16_t *low = tl + kk * IRT_LS;
				l0 = low[t0]; l1 = low[t0 + 1]; l2 = u < 127 ? (int)low[t0 + 2] : l1;
			} else {
				const uint32_t lw = *reinterpret_cast<const uint32_t *>(row + t0);
				l0 = (int16_t)(lw & 0xffffu); l1 = (int16_t)(lw >> 16);
				l2 = __shfl_down_sync(0xffffffffu, l0, 1);
				if (lane == 31) l2 = u < 127 ? (int)row[t0 + 2] : l1;
			}
			// pairs t0 and t0 + 1 of upfilter53I + upfilter53III without normalisation (dwt_core.cuh: inverse_pair)
			const int16_t ev0 = (int16_t)((int16_t)(l0 << 3) - ((h0 + hm) << 1));
			const int16_t od0 = (int16_t)((int16_t)((l1 + l0) << 2) + (6 * h0 - h1 - hm));
			const int16_t ev1 = (int16_t)((int16_t)(l1 << 3) - ((h1 + h0) << 1));
			const int16_t od1 = (int16_t)((int16_t)(u < 127 ? ((l2 + l1) << 2) : (l1 << 3)) + (6 * h1 - hp - h0));
			uint32_t *o = reinterpret_cast<uint32_t *>(to + kk * IRT_OS + 4 * u);   // (the tile's rows are 4-byte aligned only)
			o[0] = (uint32_t)(uint16_t)ev0 | ((uint32_t)(uint16_t)od0 << 16);
			o[1] = (uint32_t)(uint16_t)ev1 | ((uint32_t)(uint16_t)od1 << 16);
		}
	}
	__syncthreads();
	for (int item = tid; item < 512 * (IRT_ROWS / 8); item += 256) {
		const int j = item / (IRT_ROWS / 8), c = item % (IRT_ROWS / 8);
		int v[8];
#pragma unroll
		for (int i = 0; i < 8; i++) v[i] = to[(8 * c + i) * IRT_OS + j];
		st8(O + j * YW + k0 + 8 * c, v);
	}
}

// ---- D2: inverse scans as gathers.  Luma: cell (row, col) = coef[y_scan_pos(row, col)] (enc_point.cuh); one
// thread per 4 cells = one 8-byte run of the scan.  Chroma: one thread per (row, 8-column strip) = 16 interleaved
// coefficients of which this plane takes every other one.
__global__ void __launch_bounds__(128) kd_descan_y(DecBatch b)
{
	if (b.status[blockIdx.y] != 0) return;
	const DecImg im = make_dec(b, blockIdx.y, 0);
	// a CTA takes four rows: the four 8-byte runs a strip contributes to them are one 32-byte sector of the scan
	const int quad = blockIdx.x, strip = threadIdx.x, lane = strip & 31;
	const uint4 *src = reinterpret_cast<const uint4 *>(im.proc + strip * 2048 + quad * 16);
	const uint4 a = src[0], c = src[1];
	const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const int row = 4 * quad + i;
		uint2 o = make_uint2(w[2 * i], w[2 * i + 1]);
		if (i & 1) {   // second row of a step is stored reversed
			o.x = (w[2 * i + 1] >> 16) | (w[2 * i + 1] << 16);
			o.y = (w[2 * i] >> 16) | (w[2 * i] << 16);
		}
		*reinterpret_cast<uint2 *>(im.jpeg + row * YW + strip * 4) = o;
		// which cells hold a marker code (> 1000): one bit per cell, 16 words per row, for kd_y_markers
		const uint32_t nib = ((int16_t)(o.x & 0xffff) > 1000 ? 1u : 0u) | ((int16_t)(o.x >> 16) > 1000 ? 2u : 0u) |
		                     ((int16_t)(o.y & 0xffff) > 1000 ? 4u : 0u) | ((int16_t)(o.y >> 16) > 1000 ? 8u : 0u);
		const uint32_t word = __reduce_or_sync(0xffu << (lane & 24), nib << (4 * (lane & 7)));
		if ((lane & 7) == 0) im.mbits[row * 16 + (strip >> 3)] = word;
	}
	if (quad == 0 && strip < 2) im.cmark[strip * (1 + DEC_CMARK_CAP)] = 0;   // the chroma marker lists start empty (kd_descan_uv)
}
// Chroma: the stream interleaves U and V, 8-column strips, two rows per step (the second one reversed): a thread takes
// the 32 coefficients of one (strip, step) -- 64 contiguous bytes -- and writes the four 8-cell runs they hold (U and V,
// two rows).  A warp = the 32 strips of one step: whole sectors in, whole 512-byte rows out.  Marker codes 5003..5006
// outside the LL quadrant are not stored: they go to the component's marker list (applied after the level-2
// synthesis by kd_c_markers) and the cell is cleared, which is what the reference leaves behind.
__global__ void __launch_bounds__(256) kd_descan_uv(DecBatch b)
{
	const int img = blockIdx.y;
	if (b.status[img] != 0) return;
	const int strip = threadIdx.x & 31, step = blockIdx.x * 8 + (threadIdx.x >> 5);   // step = row pair 0..127
	const DecImg im0 = make_dec(b, img, 0), im1 = make_dec(b, img, 1);
	const uint4 *src = reinterpret_cast<const uint4 *>(im0.uvcoef + strip * 4096 + step * 32);
	int x[32];
#pragma unroll
	for (int k = 0; k < 4; k++) unpack8(src[k], x + 8 * k);
#pragma unroll
	for (int v = 0; v < 2; v++) {
		const DecImg &im = v ? im1 : im0;
#pragma unroll
		for (int half = 0; half < 2; half++) {
			const int row = 2 * step + half;
			int o[8];
#pragma unroll
			for (int t = 0; t < 8; t++) o[half ? 7 - t : t] = x[16 * half + 2 * t + v];
			if (row >= 128 || strip >= 16) {
#pragma unroll
				for (int t = 0; t < 8; t++)
					if (o[t] > 5000) {
						const uint32_t at = atomicAdd(im.cmark, 1u);
						if (at < DEC_CMARK_CAP) { im.cmark[1 + at] = ((uint32_t)(o[t] - 5000) << 24) | (uint32_t)(row * CW + strip * 8 + t); o[t] = 0; }
					}
			}
			*reinterpret_cast<uint4 *>(im.cjpeg + row * CW + strip * 8) = pack8(o);
		}
	}
}

// square transpose of the top-left N x N cells of a plane into another plane (32x32 tiles)
__global__ void __launch_bounds__(256) kd_transpose(const int16_t *in, int16_t *out, size_t in_slot, size_t out_slot, int stride)
{
	__shared__ int16_t tile[32][33];
	const int16_t *src = in + (size_t)blockIdx.z * in_slot;
	int16_t *dst = out + (size_t)blockIdx.z * out_slot;
	const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	for (int r = ty; r < 32; r += 8) tile[r][tx] = src[(y0 + r) * stride + x0 + tx];
	__syncthreads();
	for (int r = ty; r < 32; r += 8) dst[(x0 + r) * stride + y0 + tx] = tile[tx][r];
}

__global__ void kd_zero(int16_t *p, size_t slot, size_t count16)   // count16 = number of 16-byte units
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count16) reinterpret_cast<uint4 *>(p + (size_t)blockIdx.y * slot)[i] = make_uint4(0, 0, 0, 0);
}

// =====================================================================================
// Row-sequential in-place stencils as "a row per warp step"
// =====================================================================================
// D11 (edge flags) and D16 (chroma sharpen) edit a plane in place in raster order: a cell's 8-neighbour Laplacian sees
// the row above and its left neighbour AFTER their own turn, the right neighbour and the row below before.  Rows
// therefore have to be final one after the other, but inside a row the only thing that travels from cell to cell is
// what the left neighbour did to itself -- one of 5 values for the sharpen (0, +-2, +-3), one bit for the flags
// (flagged or not).  So a cell is a map {what the left cell did} -> {what this cell does}, maps compose, and a warp
// resolves a whole row with one scan of lane maps: a lane owns 8 consecutive columns, keeps the rows above / below
// in registers (16-byte loads and stores, the plane is read once and written once), and evaluates its cells for
// every incoming possibility only when one of them sits within reach of a threshold.  One warp per plane, no
// barriers: 254 row steps of a few hundred instructions instead of 760 block-wide wavefront steps.

// ---- D16: chroma sharpen (nhw_decoder.c:1085-1109; cell form dwf_sharpen_cell)
__device__ __forceinline__ int sharpen_delta(int res, int thr)
{
	const int a = res < 0 ? -res : res;
	const int d = a > thr ? (a > 160 ? 3 : 2) : 0;
	return res < 0 ? -d : d;
}
__device__ __forceinline__ int sh_state(int d) { return d == 0 ? 2 : d < 0 ? (d == -3 ? 0 : 1) : (d == 2 ? 3 : 4); }   // -3,-2,0,2,3 -> 0..4
__device__ __forceinline__ int sh_value(int st) { return st == 2 ? 0 : st < 2 ? st - 3 : st - 1; }
// maps over the five states, 3 bits per entry; (first then second)[i] = second[first[i]]
__device__ __forceinline__ uint32_t map5_then(uint32_t first, uint32_t second)
{
	uint32_t r = 0;
#pragma unroll
	for (int i = 0; i < 5; i++) r |= ((second >> (3 * ((first >> (3 * i)) & 7u))) & 7u) << (3 * i);
	return r;
}
__global__ void __launch_bounds__(128) kd_sharpen_rows(DecBatch b, int n2)
{
	const int lane = threadIdx.x & 31, pl = blockIdx.x * 4 + (threadIdx.x >> 5);
	if (pl >= n2 || b.status[pl >> 1] != 0) return;
	int16_t *P = b.c_proc + (size_t)pl * NHW_C_SLOT;
	const int thr = b.desc[pl >> 1].quality <= 14 ? 35 : 60;
	// x[1..8] = the lane's columns 8*lane .. 8*lane+7; x[0], x[9] = the neighbours' edge columns
	int up[10], cur[10], dn[10];
	auto halo = [&](int *v) {
		v[0] = __shfl_up_sync(0xffffffffu, v[8], 1);
		v[9] = __shfl_down_sync(0xffffffffu, v[1], 1);
	};
	unpack8(reinterpret_cast<const uint4 *>(P)[lane], up + 1);
	unpack8(reinterpret_cast<const uint4 *>(P + CW)[lane], cur + 1);
	halo(up);
	for (int r = 1; r < 255; r++) {
		unpack8(reinterpret_cast<const uint4 *>(P + (r + 1) * CW)[lane], dn + 1);
		halo(cur);
		halo(dn);
		// the Laplacian of every cell with its left neighbour as it was; what the left neighbour did is subtracted later
		int base[9];
		bool touchy = false;
#pragma unroll
		for (int k = 1; k <= 8; k++) {
			base[k] = 8 * cur[k] - (up[k - 1] + up[k] + up[k + 1]) - cur[k - 1] - cur[k + 1] - (dn[k - 1] + dn[k] + dn[k + 1]);
			const int a = base[k] < 0 ? -base[k] : base[k];
			touchy |= (a >= thr - 2 && a <= thr + 3) || (a >= 158 && a <= 163);   // |left's step| <= 3 can change the outcome
		}
		// columns 0 and 255 are never edited: their "step" is 0 whatever comes in
		const bool first_col = lane == 0, last_col = lane == 31;
		int din = 0;
		if (__any_sync(0xffffffffu, touchy)) {
			uint32_t m;
			if (touchy) {
				m = 0;
#pragma unroll
				for (int h = 0; h < 5; h++) {
					int x = sh_value(h);
#pragma unroll
					for (int k = 1; k <= 8; k++) {
						const bool edit = !(first_col && k == 1) && !(last_col && k == 8);
						x = edit ? sharpen_delta(base[k] - x, thr) : 0;
					}
					m |= (uint32_t)sh_state(x) << (3 * h);
				}
			} else {
				const int x = last_col ? 0 : sharpen_delta(base[8], thr);   // (lane 0 has more than one cell: cell 8 is an edited one)
				m = (uint32_t)sh_state(x) * 0x1249u;                         // the same state for all five entries
			}
#pragma unroll
			for (int dlt = 1; dlt < 32; dlt <<= 1) {
				const uint32_t o = __shfl_up_sync(0xffffffffu, m, dlt);
				if (lane >= dlt) m = map5_then(o, m);
			}
			const uint32_t before = __shfl_up_sync(0xffffffffu, m, 1);
			din = lane ? sh_value((int)((before >> 6) & 7u)) : 0;            // entry for "nothing came in" (state 2) at the row start
		}
		int x = din;
#pragma unroll
		for (int k = 1; k <= 8; k++) {
			const bool edit = !(first_col && k == 1) && !(last_col && k == 8);
			x = edit ? sharpen_delta(base[k] - x, thr) : 0;
			cur[k] += x;
		}
		reinterpret_cast<uint4 *>(P + r * CW)[lane] = pack8(cur + 1);
		halo(cur);
#pragma unroll
		for (int k = 0; k < 10; k++) { up[k] = cur[k]; cur[k] = dn[k]; }
	}
}

// ---- D11 + D12: edge flags on the reconstructed LL1 and the list of flagged positions (nhw_decoder.c:789-839; cell
// form dwf_edge_cell).  Pairs of columns (1+2p, 2+2p); a flagged cell carries +16000 while the pass runs, so the row
// above is kept WITH its flags in registers while the plane itself is written back without them (the reference
// removes them again right after, when it collects the list).  What travels along a row: whether the previous
// pair flagged its second cell (the left neighbour of this pair's first cell).
__global__ void __launch_bounds__(128) kd_edge_rows(DecBatch b, int n)
{
	const int lane = threadIdx.x & 31, img = blockIdx.x * 4 + (threadIdx.x >> 5);
	if (img >= n || b.status[img] != 0) return;
	const DecImg im = make_dec(b, img, 0);
	int16_t *P = im.proc;
	// x[1..8] = columns 8*lane .. 8*lane+7, x[0] = column 8*lane-1, x[9], x[10] = columns 8*lane+8, +9
	int up[11], cur[11], dn[11];
	auto halo = [&](int *v) {
		v[0] = __shfl_up_sync(0xffffffffu, v[8], 1);
		v[9] = __shfl_down_sync(0xffffffffu, v[1], 1);
		v[10] = __shfl_down_sync(0xffffffffu, v[2], 1);
		if (lane == 31) { v[9] = 0; v[10] = 0; }   // columns 256, 257: outside every stencil that is evaluated
	};
	unpack8(reinterpret_cast<const uint4 *>(P)[lane], up + 1);
	unpack8(reinterpret_cast<const uint4 *>(P + YW)[lane], cur + 1);
	halo(up);
	int nflags = 0;
	for (int r = 1; r < 255; r++) {
		unpack8(reinterpret_cast<const uint4 *>(P + (r + 1) * YW)[lane], dn + 1);
		halo(cur);
		halo(dn);
		auto lap = [&](int k) {
			return 8 * cur[k] - cur[k - 1] - cur[k + 1] - (up[k - 1] + up[k] + up[k + 1]) - (dn[k - 1] + dn[k] + dn[k + 1]);
		};
		// the lane's four pairs start at x[2], x[4], x[6], x[8] (columns 8*lane+1, +3, +5, +7); the last pair of lane 31
		// (columns 255, 256) does not exist
		int res[4], cnt[4];
#pragma unroll
		for (int q = 0; q < 4; q++) { res[q] = lap(2 + 2 * q); cnt[q] = lap(3 + 2 * q); }
		const int npairs = lane == 31 ? 3 : 4;
		// outcome of a pair: 0 nothing, 1 first cell flagged, 2 second cell flagged; `in` = the previous pair flagged its
		// second cell, which is this pair's first cell's left neighbour (-16000 in its Laplacian)
		auto outcome = [&](int q, int in) {
			const int rs = res[q] - (in ? 16000 : 0), ct = cnt[q];
			if (rs > 41 && rs < 108 && ct < 16) return 1;
			if (rs < -41 && rs > -108 && ct > -16) return 1;
			if (ct > 41 && ct < 108 && rs < 16) return 2;
			if (ct < -41 && ct > -108 && rs > -16) return 2;
			return 0;
		};
		// lane map over the two incoming possibilities: bit i = "the lane's last pair flags its second cell" given in = i
		uint32_t m = 0;
#pragma unroll
		for (int in = 0; in < 2; in++) {
			int x = in;
			for (int q = 0; q < npairs; q++) x = outcome(q, x) == 2 ? 1 : 0;
			m |= (uint32_t)x << in;
		}
#pragma unroll
		for (int dlt = 1; dlt < 32; dlt <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, m, dlt);
			if (lane >= dlt) m = ((m >> (o & 1u)) & 1u) | (((m >> ((o >> 1) & 1u)) & 1u) << 1);
		}
		const uint32_t before = __shfl_up_sync(0xffffffffu, m, 1);
		int x = lane ? (int)(before & 1u) : 0;     // nothing is flagged left of column 1
		const int carried = x;                      // the previous lane's last pair flagged this lane's column 8*lane
		uint32_t fl = 0;                            // bit k-1 = x[k] flagged (own columns)
		int spill = 0;
		for (int q = 0; q < npairs; q++) {
			const int o = outcome(q, x);
			if (o == 1) fl |= 1u << (1 + 2 * q);
			else if (o == 2) { if (q < 3) fl |= 1u << (2 + 2 * q); else spill = 1; }
			x = o == 2 ? 1 : 0;
		}
		if (carried) fl |= 1u;
		(void)spill;                                // (the next lane learns it through `carried`)
		// D12: positions of the flagged cells in raster order
		const int mine = __popc(fl);
		int off = mine;
#pragma unroll
		for (int dlt = 1; dlt < 32; dlt <<= 1) { const int o = __shfl_up_sync(0xffffffffu, off, dlt); if (lane >= dlt) off += o; }
		const int total = __shfl_sync(0xffffffffu, off, 31);
		off += nflags - mine;
		for (uint32_t f = fl; f; f &= f - 1) im.flags[off++] = (uint16_t)((r << 8) + 8 * lane + __ffs(f) - 1);
		nflags += total;
		// the row above the next row keeps its flags (register copy); the plane itself never shows them
#pragma unroll
		for (int k = 1; k <= 8; k++) up[k] = cur[k] + (((fl >> (k - 1)) & 1u) ? 16000 : 0);
		halo(up);
#pragma unroll
		for (int k = 0; k < 11; k++) cur[k] = dn[k];
	}
	if (lane == 0) im.list_len[9] = nflags;
}

// ---- back end: second half of the level-1 luma synthesis + clip, chroma clip + 2x upsample, YCbCr -> RGB, in one pass
// from the int16 planes to the BMP pixel bytes (the 6 B/pixel unit of SURVEY.md section 8(d): 786432 B of int16
// coefficients in, 786432 B of pixels out per image).  Everything here is local to an output row: row y of the luma
// plane comes from row y of the band plane (upfilter53I + upfilter53VI, decoder/filters.c:143-194), its chroma from
// chroma rows y/2 and y/2 + 1 (decoder/nhw_decoder.c:1137-1181).  A CTA takes BE_ROWS output rows of one image: the
// band rows and the clipped chroma rows are staged in shared memory with 16-byte loads, a thread computes two
// neighbouring pixels of every row, the pixels leave through shared memory as one contiguous 16-byte-per-thread copy.
#define BE_ROWS 16
template <bool WRITE_YUV>
__global__ void __launch_bounds__(256) kd_backend(DecBatch b, uint8_t *__restrict__ rgb)
{
	__shared__ __align__(16) int16_t sy[BE_ROWS][512];
	__shared__ __align__(16) uint8_t sc[2][BE_ROWS / 2 + 1][256];
	__shared__ __align__(16) uint8_t so[BE_ROWS * 1536];
	const int img = blockIdx.y, y0 = blockIdx.x * BE_ROWS, t = threadIdx.x;
	uint4 *dst = reinterpret_cast<uint4 *>(rgb + (size_t)img * 786432 + (size_t)y0 * 1536);
	if (b.status[img] != 0) {   // a refused stream decodes to black
		for (int k = t; k < BE_ROWS * 96; k += 256) dst[k] = make_uint4(0, 0, 0, 0);
		return;
	}
	const int16_t *J = b.y_aux + (size_t)img * NHW_Y_SLOT + y0 * YW;   // kd_inv_rows_t's output, after the flag smoothing
	for (int k = t; k < BE_ROWS * 64; k += 256) reinterpret_cast<uint4 *>(&sy[0][0])[k] = reinterpret_cast<const uint4 *>(J)[k];
	const int r0 = y0 >> 1;
	for (int k = t; k < 2 * (BE_ROWS / 2 + 1) * 32; k += 256) {
		const int v = k / ((BE_ROWS / 2 + 1) * 32), rem = k % ((BE_ROWS / 2 + 1) * 32), rr = rem >> 5, c8 = (rem & 31) * 8;
		const int r = r0 + rr < 256 ? r0 + rr : 255;
		const int16_t *P = b.c_proc + ((size_t)img * 2 + v) * NHW_C_SLOT + r * CW + c8;
		int x[8];
		ld8(P, x);
		uint2 o;
		o.x = (uint32_t)dec_clip8(x[0]) | ((uint32_t)dec_clip8(x[1]) << 8) | ((uint32_t)dec_clip8(x[2]) << 16) | ((uint32_t)dec_clip8(x[3]) << 24);
		o.y = (uint32_t)dec_clip8(x[4]) | ((uint32_t)dec_clip8(x[5]) << 8) | ((uint32_t)dec_clip8(x[6]) << 16) | ((uint32_t)dec_clip8(x[7]) << 24);
		*reinterpret_cast<uint2 *>(&sc[v][rr][c8]) = o;
	}
	__syncthreads();
	const DecColor col = dec_color_of(b.desc[img].quality);
	uint8_t *yuv = WRITE_YUV ? b.yuv + (size_t)img * 786432 : nullptr;
	const int c1 = t < 255 ? t + 1 : 255;
#pragma unroll 2
	for (int rr = 0; rr < BE_ROWS; rr++) {
		const int16_t *row = sy[rr];
		int ev, od;
		inverse_pair([&](int k) { return (int)row[k]; }, [&](int k) { return (int)row[256 + k]; }, t, 256, true, ev, od);
		const int Y0 = dec_clip8(ev), Y1 = dec_clip8(od);
		// chroma of pixels (2t, 2t+1) of output row y: dec_c_upsample_cell's rule on the clipped samples
		const int y = y0 + rr, cr = (y >> 1) - r0;
		int uvs[2][2];
#pragma unroll
		for (int v = 0; v < 2; v++) {
			int a0 = sc[v][cr][t], a1 = sc[v][cr][c1];
			if ((y & 1) && (y >> 1) < 255) {
				a0 = (a0 + sc[v][cr + 1][t] + 1) >> 1;
				a1 = (a1 + sc[v][cr + 1][c1] + 1) >> 1;
			}
			uvs[v][0] = a0;
			uvs[v][1] = t < 255 ? (a0 + a1 + 1) >> 1 : a0;
		}
		uint8_t px[6];
		dec_ycc_to_rgb(Y0, uvs[0][0], uvs[1][0], col, px);
		dec_ycc_to_rgb(Y1, uvs[0][1], uvs[1][1], col, px + 3);
		uint16_t *o = reinterpret_cast<uint16_t *>(so + rr * 1536 + 6 * t);
		o[0] = (uint16_t)(px[0] | (px[1] << 8));
		o[1] = (uint16_t)(px[2] | (px[3] << 8));
		o[2] = (uint16_t)(px[4] | (px[5] << 8));
		if (WRITE_YUV) {
			*reinterpret_cast<uint16_t *>(yuv + y * 512 + 2 * t) = (uint16_t)(Y0 | (Y1 << 8));
			*reinterpret_cast<uint16_t *>(yuv + 262144 + y * 512 + 2 * t) = (uint16_t)(uvs[0][0] | (uvs[0][1] << 8));
			*reinterpret_cast<uint16_t *>(yuv + 524288 + y * 512 + 2 * t) = (uint16_t)(uvs[1][0] | (uvs[1][1] << 8));
		}
	}
	__syncthreads();
	for (int k = t; k < BE_ROWS * 96; k += 256) dst[k] = reinterpret_cast<const uint4 *>(so)[k];
}

// every (Y, U, V) triple through both forms of the q >= 20 decoder colour matrix; counts disagreements
__global__ void kd_color_check(unsigned long long *bad)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = t & 255u, u = (t >> 8) & 255u, v = t >> 16;
	const DecColor col = dec_color_of(20);
	uint8_t a[3], b[3];
	dec_ycc_to_rgb_ieee(y, u, v, col, a);
	dec_ycc_to_rgb(y, u, v, col, b);
	if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) atomicAdd(bad, 1ull);
}

// ---- device-resident streams: header walk on the device (dec_parse.h), one thread per stream
__global__ void kd_parse_headers(const uint8_t *in, size_t stride, const uint32_t *len, const uint64_t *offs, int n, DecDesc *desc,
                                 uint64_t *off_out, int32_t *status)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t o = offs ? offs[i] : (uint64_t)i * stride;
	const uint64_t l = offs ? offs[i + 1] - offs[i] : (uint64_t)len[i];
	off_out[i] = o;
	status[i] = nhw_parse_header(in + o, (size_t)l, desc + i);
}

}  // namespace

namespace nhw {

bool decode_device_init(nhw_ctx *c)
{
	// Built ONCE, by whichever thread gets here first (function-local static: thread-safe initialisation).  It used to be a
	// plain static array that every nhw_create zeroed and refilled: two threads creating contexts at the same time could
	// upload a half-zeroed table, and every context of the device then failed on the rarer codes (NHW_ERR_CODEBOOK on noise
	// images) until the next nhw_create repaired it -- seen once, right after tests/test_contexts_gpu.py.
	struct Lut { uint16_t w[NHW_LUT_WORDS]; };
	static const Lut *const table = [] { Lut *t = new Lut; dec_build_lut(t->w); return t; }();
	void *p = nullptr;
	if (!check(cudaMemcpyToSymbol(g_dec_lut, table->w, sizeof(table->w)), "prefix-code table") || !check(cudaGetSymbolAddress(&p, g_dec_lut), "prefix-code table"))
		return false;
	c->dec_lut = static_cast<const uint16_t *>(p);
	return check(cudaFuncSetAttribute(kd_inv_rows_t, cudaFuncAttributeMaxDynamicSharedMemorySize, IRT_SMEM), "attr kd_inv_rows_t");
}

long dec_color_fast_path_mismatches(nhw_ctx *c)
{
	unsigned long long *bad = nullptr, host = ~0ull;
	if (cudaMalloc(&bad, 8) != cudaSuccess) return -1;
	cudaMemsetAsync(bad, 0, 8, c->stream);
	kd_color_check<<<65536, 256, 0, c->stream>>>(bad);
	cudaMemcpyAsync(&host, bad, 8, cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	cudaFree(bad);
	return (long)host;
}

// from encode.cu (same inverse kernels, natural-orientation output)
void idwt_rows_cols(nhw_ctx *c, int n_planes, const int16_t *in, int16_t *tmp, int16_t *out, size_t slot, int N, int stride);

// Decode n <= max_batch streams.  blobs/offs/desc/status are device arrays for this chunk; rgb_dev
// receives n x 786432 bytes.  All work is queued on c->stream.
void decode_chunk(nhw_ctx *c, const uint8_t *blobs, const uint64_t *offs, const DecDesc *desc, int32_t *status, int n,
                  uint8_t *rgb_dev, bool any_lowq, bool any_hq, bool want_yuv)
{
	DecBatch b;
	b.blobs = blobs; b.blob_off = offs; b.desc = desc; b.status = status;
	b.y_proc = c->y_proc + NHW_GUARD_S; b.y_jpeg = c->y_jpeg + NHW_GUARD_S; b.y_aux = c->y_aux + NHW_GUARD_S;
	b.uvcoef = c->y_aux2 + NHW_GUARD_S;
	b.c_proc = c->c_proc + NHW_GUARD_S; b.c_jpeg = c->c_jpeg + NHW_GUARD_S; b.c_aux = c->c_aux + NHW_GUARD_S;
	b.bytes = c->dec_bytes; b.yuv = c->dec_yuv;
	b.lut = c->dec_lut;
	const size_t YS = NHW_Y_SLOT, CS = NHW_C_SLOT;
	const bool side = c->chroma_side && c->chroma_stream;

	// coefficient planes start at zero: zero runs are skipped, not written (decoder/nhw_decoder.c:2029)
	NHW_LAUNCH(c, kd_zero, dim3(262144 * 2 / 16 / 256, n), 256, 0, b.y_proc, YS, (size_t)(262144 * 2 / 16));
	NHW_LAUNCH(c, kd_zero, dim3(131072 * 2 / 16 / 256, n), 256, 0, b.uvcoef, YS, (size_t)(131072 * 2 / 16));

	// ---- luma
	{
		// streams per warp: with few streams in flight one per warp is fastest (no divergence between the parsers, and there
		// are warps to spare: 512 streams 2.6 ms vs 4.4 ms at four per warp); with thousands the warps themselves become the
		// limit and sharing one pays (4096 streams: 5.7 ms at four per warp, 7.8 at two, 11.2 at one)
		const int spw = c->tune.dsf_streams ? c->tune.dsf_streams : n <= 1536 ? 1 : n <= 3072 ? 2 : 4;
		if (side) cudaEventRecord(c->ev_chroma0, c->stream);   // (the LL kernel depends on what precedes the serial front only)
		NHW_LAUNCH_L(c, "d_serial_front", kd_serial_front, (n + spw - 1) / spw, dim3(32, 3), 0, b, n, spw, c->tune.dsf_job_mask);
	}
	// the LL bytes next to the serial front, which leaves most issue slots idle (launched after it, so that its blocks -- whose
	// longest chain sets the kernel's time -- are resident first): on the side stream when the chunk has the GPU to itself
	// (device-resident calls), joined before their first reader (d_ll_y)
	if (side) {
		const cudaStream_t main_stream = c->stream;
		cudaStreamWaitEvent(c->chroma_stream, c->ev_chroma0, 0);
		c->stream = c->chroma_stream;
		NHW_LAUNCH_L(c, "d_ll_bytes", kd_ll_parallel, n, LLP_THREADS, 0, b);
		cudaEventRecord(c->ev_chroma1, c->chroma_stream);
		c->stream = main_stream;
	} else NHW_LAUNCH_L(c, "d_ll_bytes", kd_ll_parallel, n, LLP_THREADS, 0, b);
	NHW_LAUNCH_L(c, "d_descan_y", kd_descan_y, dim3(128, n), 128, 0, b);
	NHW_LAUNCH_L(c, "d_markers_y", kd_y_markers, n, 256, 0, b);
	if (side) cudaStreamWaitEvent(c->stream, c->ev_chroma1, 0);
	NHW_LAUNCH_L(c, "d_ll_y", kd_y_ll, n, 256, 0, b);
	NHW_LAUNCH_L(c, "d_shrink_y", kd_shrink_y, dim3(32, n), 256, 0, b);
	if (any_lowq) NHW_LAUNCH_L(c, "d_shrink_y_lowq", kd_shrink_y_lowq, n, 256, 0, b);
	idwt_rows_cols(c, n, b.y_jpeg, b.y_aux, b.y_proc, YS, 256, 512);
	NHW_LAUNCH_L(c, "d_addbacks", kd_addbacks, n, 256, 0, b);
	NHW_LAUNCH_L(c, "d_edge_rows", kd_edge_rows, (n + 3) / 4, 128, 0, b, n);   // D11 + D12
	NHW_LAUNCH_L(c, "d_inv_rows_t", kd_inv_rows_t, dim3(512 / IRT_ROWS, n), 256, IRT_SMEM, b);   // -> y_aux
	if (any_hq) NHW_LAUNCH_L(c, "d_hq_addbacks", kd_hq_addbacks, n, 256, 0, b);   // q22 / q23 streams only
	NHW_LAUNCH_L(c, "d_smooth_flags", kd_smooth_flags, n, 256, 0, b);
	// (the second half of the synthesis and the clip are part of the back-end kernel below)

	// ---- chroma
	NHW_LAUNCH_L(c, "d_descan_uv", kd_descan_uv, dim3(16, n), 256, 0, b);
	NHW_LAUNCH_L(c, "d_ll_uv", kd_c_ll, n, 256, 0, b);
	idwt_rows_cols(c, 2 * n, b.c_jpeg, b.c_aux, b.c_proc, CS, 128, 256);
	NHW_LAUNCH_L(c, "d_markers_uv", kd_c_markers, 2 * n, 256, 0, b);
	NHW_LAUNCH(c, kd_transpose, dim3(4, 4, 2 * n), 256, 0, b.c_proc, b.c_jpeg, CS, CS, 256);
	idwt_rows_cols(c, 2 * n, b.c_jpeg, b.c_aux, b.c_proc, CS, 256, 256);
	NHW_LAUNCH_L(c, "d_sharpen_rows", kd_sharpen_rows, (2 * n + 3) / 4, 128, 0, b, 2 * n);

	// ---- back end: luma synthesis tail, chroma upsample, colour -> pixels (and the Y/U/V byte planes when asked for)
	if (want_yuv) NHW_LAUNCH_L(c, "d_backend<yuv>", kd_backend<true>, dim3(512 / BE_ROWS, n), 256, 0, b, rgb_dev);
	else NHW_LAUNCH_L(c, "d_backend", kd_backend<false>, dim3(512 / BE_ROWS, n), 256, 0, b, rgb_dev);
}

// streams resident in device memory: stream i at in + offs[i] (offs != NULL, n + 1 entries) or at in + i * stride with
// length len[i].  Headers are walked on the device; everything is queued on c->stream.
void decode_chunk_device(nhw_ctx *c, const uint8_t *in, size_t stride, const uint32_t *len, const uint64_t *offs, int n,
                         uint8_t *rgb_dev, int32_t *status_dev)
{
	DecDesc *desc = static_cast<DecDesc *>(c->dec_desc_dev);
	NHW_LAUNCH(c, kd_parse_headers, (n + 127) / 128, 128, 0, in, stride, len, offs, n, desc, c->offs_dev, c->status_dev);
	c->chroma_side = (c->tune.chroma_stream && !c->profile && !c->dbg_label[0]) ? 1 : 0;
	decode_chunk(c, in, c->offs_dev, desc, c->status_dev, n, rgb_dev, true, true, false);
	c->chroma_side = 0;
	if (status_dev) cudaMemcpyAsync(status_dev, c->status_dev, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream);
}

}  // namespace nhw
