// front_fused.cu -- the encoder front end as two kernels:
//
//   k_front_luma : RGB -> YCbCr (+4:2:0 chroma bytes) -> luma pre-sharpening -> level-1
//                  wavelet analysis of the luma plane, ONE pass over the pixels.
//   k_dwt_level  : one analysis level of an N x N band held entirely in shared memory
//                  (luma level 2, chroma levels 1 and 2, and the closed-loop re-analysis).
//
// Reference behaviour:
//   downsample_YUV420            encoder/colorspace.c:55-260
//   pre_processing (q17..q21)    encoder/image_processing.c:558-836,1926-1990
//   wavelet_analysis             encoder/wavelet_filterbank.c:52-302
//   downfilter53IV / 53VI / 53   encoder/filters.c:346-386,203-287,55-114
//
// k_front_luma streams one image per CTA top to bottom in strips of 16 rows.  The only
// whole-image dependency of the front end is the 4-bit remainder that pre_processing's first
// loop carries through the image in raster order; it enters the next element as one of five
// classes, so a row is a map {class -> class} and the strip order of the CTA *is* the raster
// order: no second pass over the pixels.  Per strip:
//   C   colour transform of the rows that just arrived (TMA bulk copy, issued one strip
//       ahead) -> luma bytes ring, horizontally filtered chroma ring
//   E   Laplacian energies of 16 rows (packed byte arithmetic), per-lane carry maps, warp scan
//       of the maps -> row maps;  chroma vertical filter + 2:1 decimation -> HBM
//   A   each warp walks the row maps of the rows above it in the strip, replays its row with
//       the true carry, applies the pair nudges and runs the horizontal filter out of
//       registers -> first-pass ring (the sharpened luma plane never exists in memory)
//   V   vertical filter: thread = column, 8 low + 8 high outputs per strip -> HBM, already
//       in the reference's transposed orientation; LL1 goes out in natural orientation as
//       `res256` (encoder/nhw_encoder.c:127-135), which is also the input of level 2.
// HBM traffic per image: 786432 B of pixels in, 524288 B luma coefficients + 131072 B chroma
// bytes out; nothing else leaves the SM.
#include "nhw_ctx.h"
#include "nhw_dev.cuh"
#include "enc_img.cuh"
#include "dwt_core.cuh"
#include "color_core.cuh"
#include "pre_core.cuh"

namespace {

constexpr int FT = 512;          // threads per CTA (16 warps: one per row of a strip)
constexpr int RING = 24;         // rows in each ring buffer
constexpr int RGB_ROWS = 18;     // the first strip carries two extra rows

struct __align__(128) FrontSmem {
	uint8_t rgb[RGB_ROWS * 1536];        // TMA destination, rows of the current strip
	uint8_t yring[RING][512];            // luma bytes, row y at slot y % RING
	uint8_t uvh[2][RING][256];           // horizontally filtered U / V at even pixels
	int16_t rring[RING + 19][512];       // horizontal-pass output R[y][k]; slots 0..18 are mirrored at
	                                     // RING..RING+18 so a 19-row window never wraps
	uint16_t plut[PAIR_CATS * PAIR_CATS + 1];   // pair rule tables (pre_core.cuh)
	uint8_t pcat[512];
	uint32_t rowmap_lo[16], rowmap_hi[16];   // class map of each row of the strip
	uint32_t part_lo[16], part_hi[16];       // class map of the row up to x = 508
	int e509[16], e510[16];                  // signed energies of the row's last pair
	int strip_carry[2];                      // class entering the first row of strip i (i & 1)
	int strip_flag[2];                       // pair flag left behind by the row above strip i
	unsigned long long mbar;
};

__device__ uint8_t g_pair_cat[512];
__device__ uint16_t g_pair_lut[PAIR_CATS * PAIR_CATS];

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

// ---- carry maps ------------------------------------------------------------------------
// A class map is five classes (0..4) as bytes: lo = classes for entry class 0..3, hi = class
// for entry class 4.  compose(first, then)[c] = then[first[c]] is two byte permutes.
struct CMap { uint32_t lo, hi; };
__device__ __forceinline__ int cmap_at(const CMap &m, int c) { return (int)((c < 4 ? (m.lo >> (8 * c)) : m.hi) & 7u); }
__device__ __forceinline__ CMap cmap_then(const CMap &first, const CMap &then)
{
	uint32_t t = first.lo | (first.lo >> 4);
	uint32_t sel = __byte_perm(t, 0u, 0x4420);          // nibbles f0 f1 f2 f3
	CMap r;
	r.lo = __byte_perm(then.lo, then.hi, sel);
	r.hi = __byte_perm(then.lo, then.hi, first.hi & 7u) & 0xffu;
	return r;
}
// five 4-bit states in 6-bit fields -> classes (state + 2) >> 2, still in 6-bit fields
__device__ __forceinline__ uint32_t sv_classes(uint32_t v) { return ((v + 0x02082082u) >> 2) & 0x071C71C7u; }
__device__ __forceinline__ CMap sv_to_cmap(uint32_t v)
{
	const uint32_t c = sv_classes(v);
	CMap m;
	m.lo = (c & 7u) | (((c >> 6) & 7u) << 8) | (((c >> 12) & 7u) << 16) | (((c >> 18) & 7u) << 24);
	m.hi = (c >> 24) & 7u;
	return m;
}
// one element of the raster recurrence applied to all five hypotheses at once
__device__ __forceinline__ uint32_t sv_step(uint32_t v, int e)
{
	if (e == 0) return 0u;
	const uint32_t a = (uint32_t)(e < 0 ? -e : e) & 15u;
	return (sv_classes(v) + a * 0x01041041u) & 0x0F3CF3CFu;
}
constexpr uint32_t SV_INIT = 0u | (2u << 6) | (6u << 12) | (10u << 18) | (14u << 24);   // states of class 0..4

// signed pre-sharpening value and the carry of one element (image_processing.c:622-626)
__device__ __forceinline__ int carry_step(int e, int &cls)
{
	if (e == 0) { cls = 0; return 0; }
	const int v = (e < 0 ? -e : e) + cls;
	cls = ((v & 15) + 2) >> 2;
	const int k = v >> 4;
	return e < 0 ? -k : k;
}

__device__ __forceinline__ uint32_t pack2(int a, int b) { return (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16); }

// horizontal filter of one row held as 16 values per lane (x[0..15] = columns 16*lane..),
// written to R[0..255] (low) and R[256..511] (high).  downfilter53IV, encoder/filters.c:346-386.
__device__ __forceinline__ void row_pass_regs(const int (&x)[16], int lane, int16_t *R, bool mirror)
{
	int xm2 = __shfl_up_sync(0xffffffffu, x[14], 1), xm1 = __shfl_up_sync(0xffffffffu, x[15], 1);
	int xp = __shfl_down_sync(0xffffffffu, x[0], 1);
	if (lane == 0) { xm2 = x[2]; xm1 = x[1]; }
	if (lane == 31) xp = x[14];
	int lo[8], hi[8];
#pragma unroll
	for (int s = 0; s < 8; s++) {
		const int a2 = s ? x[2 * s - 2] : xm2, a1 = s ? x[2 * s - 1] : xm1;
		const int b2 = s < 7 ? x[2 * s + 2] : xp;
		lo[s] = 6 * x[2 * s] + 2 * (a1 + x[2 * s + 1]) - (a2 + b2);
		hi[s] = 2 * x[2 * s + 1] - (x[2 * s] + b2);
	}
	if (lane == 31) hi[7] = (x[15] - x[14]) << 1;
	uint4 L = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
	uint4 H = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
	reinterpret_cast<uint4 *>(R)[lane] = L;
	reinterpret_cast<uint4 *>(R + 256)[lane] = H;
	if (mirror) {
		reinterpret_cast<uint4 *>(R + RING * 512)[lane] = L;
		reinterpret_cast<uint4 *>(R + RING * 512 + 256)[lane] = H;
	}
}

__device__ __forceinline__ void unpack16(const uint4 &w, int (&x)[16])
{
	const uint32_t v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
	for (int j = 0; j < 4; j++) {
		x[4 * j] = v[j] & 255u;
		x[4 * j + 1] = (v[j] >> 8) & 255u;
		x[4 * j + 2] = (v[j] >> 16) & 255u;
		x[4 * j + 3] = v[j] >> 24;
	}
}

// 16 consecutive int16 samples of a plane row (32-byte aligned)
__device__ __forceinline__ void load16(const int16_t *p, int (&x)[16])
{
	const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
	const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
	for (int k = 0; k < 8; k++) {
		x[2 * k] = (int16_t)(w[k] & 0xffffu);
		x[2 * k + 1] = (int16_t)(w[k] >> 16);
	}
}

// 8 low + 8 high outputs of the vertical filter for one column: col[j] = R[2*e0 - 2 + j][k],
// j = 0..18.  downfilter53VI (fine) / downfilter53, encoder/filters.c:203-287,55-114.
__device__ __forceinline__ void col_pass8(const int (&col)[19], int e0, bool fine, bool last_group, int &rem,
                                          int (&lo)[8], int (&hi)[8])
{
#pragma unroll
	for (int s = 0; s < 8; s++) {
		const int c = 2 + 2 * s;
		const int rl = 6 * col[c] + 2 * (col[c - 1] + col[c + 1]) - (col[c - 2] + col[c + 2]);
		if (fine) {
			const int acc = rl + ((e0 + s) > 0 ? rem : 0);
			lo[s] = nhw_sround((int)(int16_t)acc, 32, 6);
			rem = vi_remainder(rl);
		} else {
			lo[s] = nhw_sround(rl, 8, 4);
		}
		if (last_group && s == 7) {
			const int d = col[c + 1] - col[c];
			hi[s] = fine ? (d >> 3) : ((d + 1) >> 1);
		} else {
			int a = col[c] + col[c + 2];
			if ((s & 1) && (a & 1) && ((col[c - 2] + col[c]) & 1)) a++;   // e0 is even: parity of e = parity of s
			const int rh = col[c + 1] - (a >> 1);
			hi[s] = fine ? nhw_sround(rh, 4, 3) : (rh > 0 ? ((rh + 1) >> 1) : (rh >> 1));
		}
	}
}

// =====================================================================================
// k_front_luma
// =====================================================================================
// PLANE_IN (q <= 16): the luma plane already exists in memory as int16 (colour transform and the pre-sharpening state
// machine ran as kernels of their own, front.cu), so stages C and E are skipped, stage A reads its row from `yplane`
// and the chroma bytes are not touched.
template <bool PLANE_IN>
__global__ void __launch_bounds__(FT, 2)
k_front_luma(const uint8_t *__restrict__ rgb, const int16_t *__restrict__ yplane, size_t ystride, int16_t *__restrict__ proc,
             size_t pstride, int16_t *__restrict__ ll1, size_t lstride, uint8_t *__restrict__ uv, size_t uvstride,
             ColorParams cp, int pre, int16_t *__restrict__ kept, size_t kstride)
{
	extern __shared__ __align__(128) uint8_t smem_raw[];
	FrontSmem &S = *reinterpret_cast<FrontSmem *>(smem_raw);
	const int img = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint8_t *src = PLANE_IN ? nullptr : rgb + (size_t)img * NHW_RGB_BYTES;
	const int16_t *YP = PLANE_IN ? yplane + (size_t)img * ystride : nullptr;
	int16_t *P = proc + (size_t)img * pstride;
	int16_t *LL = ll1 + (size_t)img * lstride;
	uint8_t *UV = PLANE_IN ? nullptr : uv + (size_t)img * uvstride;

	if (!PLANE_IN) {
		if (tid == 0) {
			mbar_init(&S.mbar, 1);
			S.strip_carry[0] = 0;
			S.strip_flag[0] = 0;
		}
		S.pcat[tid] = g_pair_cat[tid];
		if (tid < PAIR_CATS * PAIR_CATS) S.plut[tid] = g_pair_lut[tid];
		__syncthreads();
		if (tid == 0) bulk_load(S.rgb, src, RGB_ROWS * 1536, &S.mbar);
	}

	int v_rem = 0;   // downfilter53VI's fed-forward remainder of this thread's column
	// registers that live from stage E to stage A of a strip
	int e[16];
	uint4 ymid = make_uint4(0, 0, 0, 0);
	CMap excl = {0x03020100u, 4u};

	for (int i = 0; i < 32; i++) {
		const int y_lo = i ? 16 * i + 2 : 0;
		const int y_hi = 16 * i + 18 < 512 ? 16 * i + 18 : 512;

		// ------------------------------------------------------------------ C: colour
		if (!PLANE_IN) {
		mbar_wait(&S.mbar, (uint32_t)(i & 1));
		for (int row = y_lo + warp; row < y_hi; row += 16) {
			const uint8_t *line = S.rgb + (row - y_lo) * 1536;
			uint8_t *yrow = S.yring[row % RING];
			uint8_t *urow = S.uvh[0][row % RING], *vrow = S.uvh[1][row % RING];
			uint32_t edge = 0;
#pragma unroll 1
			for (int step = 0; step < 4; step++) {
				const uint32_t *w = reinterpret_cast<const uint32_t *>(line + step * 384 + lane * 12);
				const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
				// four pixels, each as c0 | c1 << 8 | c2 << 16
				const uint32_t px[4] = {w0 & 0xffffffu, __funnelshift_r(w0, w1, 24) & 0xffffffu, __funnelshift_r(w1, w2, 16) & 0xffffffu, w2 >> 8};
				int Y[4];
				uint32_t uv[4];   // U | V << 16
				if (cp.mode == 0) {
					rgb4px_to_ycc_q20(px, Y, uv);
				} else {
#pragma unroll
					for (int k = 0; k < 4; k++) {
						int U, V;
						rgb_to_ycc((int)(px[k] & 255u), (int)((px[k] >> 8) & 255u), (int)(px[k] >> 16), cp, Y[k], U, V);
						uv[k] = (uint32_t)U | ((uint32_t)V << 16);
					}
				}
				reinterpret_cast<uint32_t *>(yrow)[step * 32 + lane] =
				    (uint32_t)Y[0] | ((uint32_t)Y[1] << 8) | ((uint32_t)Y[2] << 16) | ((uint32_t)Y[3] << 24);
				// [1 2 1]/4 on even pixels (colorspace.c:220-234), U and V side by side in 16-bit
				// halves; the pixel left of the lane comes from the neighbour lane, or from lane 31
				// of the previous step
				uint32_t left = __shfl_up_sync(0xffffffffu, uv[3], 1);
				if (lane == 0) left = edge;
				edge = __shfl_sync(0xffffffffu, uv[3], 31);
				uint32_t fa = ((left + 2u * uv[0] + uv[1] + 0x00020002u) >> 2) & 0x00ff00ffu;
				if (step == 0 && lane == 0) fa = ((uv[0] + uv[1] + 0x00010001u) >> 1) & 0x00ff00ffu;
				const uint32_t fb = ((uv[1] + 2u * uv[2] + uv[3] + 0x00020002u) >> 2) & 0x00ff00ffu;
				reinterpret_cast<uint16_t *>(urow)[step * 32 + lane] = (uint16_t)__byte_perm(fa, fb, 0x0040);
				reinterpret_cast<uint16_t *>(vrow)[step * 32 + lane] = (uint16_t)__byte_perm(fa, fb, 0x0062);
			}
		}
		__syncthreads();
		if (tid == 0 && i < 31) {
			const int n_lo = 16 * (i + 1) + 2;
			const int n_hi = n_lo + 16 < 512 ? n_lo + 16 : 512;
			bulk_load(S.rgb, src + (size_t)n_lo * 1536, (uint32_t)(n_hi - n_lo) * 1536u, &S.mbar);
		}
		}

		// ------------------------------------------------------------------ E: energies, carry maps; chroma out
		const int r = 16 * i + 1 + warp;
		const bool sharpen = !PLANE_IN && pre && r <= 510;
		if (sharpen) {
			const uint8_t *up = S.yring[(r - 1) % RING], *mid = S.yring[r % RING], *dn = S.yring[(r + 1) % RING];
			uint32_t wu[6], wm[6], wd[6];
			{
				const uint4 a = reinterpret_cast<const uint4 *>(up)[lane], b = reinterpret_cast<const uint4 *>(mid)[lane],
				            c = reinterpret_cast<const uint4 *>(dn)[lane];
				wu[1] = a.x; wu[2] = a.y; wu[3] = a.z; wu[4] = a.w;
				wm[1] = b.x; wm[2] = b.y; wm[3] = b.z; wm[4] = b.w;
				wd[1] = c.x; wd[2] = c.y; wd[3] = c.z; wd[4] = c.w;
				ymid = b;
				const int lw = lane ? 4 * lane - 1 : 0, rw = lane < 31 ? 4 * lane + 4 : 127;
				wu[0] = reinterpret_cast<const uint32_t *>(up)[lw]; wu[5] = reinterpret_cast<const uint32_t *>(up)[rw];
				wm[0] = reinterpret_cast<const uint32_t *>(mid)[lw]; wm[5] = reinterpret_cast<const uint32_t *>(mid)[rw];
				wd[0] = reinterpret_cast<const uint32_t *>(dn)[lw]; wd[5] = reinterpret_cast<const uint32_t *>(dn)[rw];
			}
			uint32_t sv = SV_INIT, sv508 = SV_INIT;
#pragma unroll
			for (int t = 0; t < 16; t++) {
				// bytes (x-1, x, x+1, x) of each row, x = 16*lane + t, out of the 24-byte window
				const int b0 = 3 + t, j = b0 >> 2, o = b0 & 3;
				const uint32_t sel = (uint32_t)o | ((uint32_t)(o + 1) << 4) | ((uint32_t)(o + 2) << 8) | ((uint32_t)(o + 1) << 12);
				const uint32_t mu = __byte_perm(wu[j], wu[j + 1], sel) & 0x00ffffffu;
				const uint32_t mm = __byte_perm(wm[j], wm[j + 1], sel);
				const uint32_t md = __byte_perm(wd[j], wd[j + 1], sel) & 0x00ffffffu;
				const uint32_t c = (mm >> 8) & 255u;
				const uint32_t c4 = c * 0x01010101u, c3 = c4 & 0x00ffffffu;
				const uint32_t cnt = __vsadu4(mu, c3) + __vsadu4(md, c3) + __vsadu4(mm, c4);
				const uint32_t sum = __dp4a(mu, 0x01010101u, __dp4a(md, 0x01010101u, __dp4a(mm, 0x00010001u, 0u)));
				const int res = (int)(8u * c) - (int)sum;
				const int mag = 15 * (res < 0 ? -res : res) + (int)cnt;
				int ev = res == 0 ? 0 : (res < 0 ? -mag : mag);
				const bool outside = (t == 0 && lane == 0) || (t == 15 && lane == 31);   // x = 0, x = 511: not visited
				if (outside) ev = 0;
				e[t] = ev;
				if (!outside) sv = sv_step(sv, ev);
				if (t == 12) sv508 = sv;
			}
			// inclusive scan of the lane maps; keep the exclusive prefix for stage A
			CMap m = sv_to_cmap(sv);
			const CMap own = m;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				CMap o;
				o.lo = __shfl_up_sync(0xffffffffu, m.lo, d);
				o.hi = __shfl_up_sync(0xffffffffu, m.hi, d);
				if (lane >= d) m = cmap_then(o, m);
			}
			excl.lo = __shfl_up_sync(0xffffffffu, m.lo, 1);
			excl.hi = __shfl_up_sync(0xffffffffu, m.hi, 1);
			if (lane == 0) { excl.lo = 0x03020100u; excl.hi = 4u; }
			(void)own;
			if (lane == 31) {
				S.rowmap_lo[warp] = m.lo;
				S.rowmap_hi[warp] = m.hi;
				const CMap part = cmap_then(excl, sv_to_cmap(sv508));
				S.part_lo[warp] = part.lo;
				S.part_hi[warp] = part.hi;
				S.e509[warp] = e[13];
				S.e510[warp] = e[14];
			}
		}
		// chroma rows whose three source rows are complete: vertical [1 2 1]/4 + 2:1 (colorspace.c:241-256)
		if (!PLANE_IN) {
			const int c_lo = i ? 8 * i + 1 : 0, c_hi = i < 31 ? 8 * i + 9 : 256;
			for (int task = warp; task < 2 * (c_hi - c_lo); task += 16) {
				const int plane = task & 1, cr = c_lo + (task >> 1);
				const uint8_t(*hrows)[256] = S.uvh[plane];
				const uint2 b = reinterpret_cast<const uint2 *>(hrows[(2 * cr) % RING])[lane];
				const uint2 c = reinterpret_cast<const uint2 *>(hrows[(2 * cr + 1) % RING])[lane];
				uint2 o;
				if (cr == 0) {
					o.x = __vavgu4(b.x, c.x);
					o.y = __vavgu4(b.y, c.y);
				} else {
					const uint2 a = reinterpret_cast<const uint2 *>(hrows[(2 * cr - 1) % RING])[lane];
					const uint32_t aw[2] = {a.x, a.y}, bw[2] = {b.x, b.y}, cw[2] = {c.x, c.y};
					uint32_t ow[2];
#pragma unroll
					for (int h = 0; h < 2; h++) {
						const uint32_t ev = (aw[h] & 0x00ff00ffu) + 2u * (bw[h] & 0x00ff00ffu) + (cw[h] & 0x00ff00ffu) + 0x00020002u;
						const uint32_t od = ((aw[h] >> 8) & 0x00ff00ffu) + 2u * ((bw[h] >> 8) & 0x00ff00ffu) +
						                    ((cw[h] >> 8) & 0x00ff00ffu) + 0x00020002u;
						ow[h] = ((ev >> 2) & 0x00ff00ffu) | (((od >> 2) & 0x00ff00ffu) << 8);
					}
					o.x = ow[0];
					o.y = ow[1];
				}
				reinterpret_cast<uint2 *>(UV + (size_t)plane * NHW_CPLANE + cr * 256)[lane] = o;
			}
		}
		__syncthreads();   // (PLANE_IN: the one barrier that keeps the next stage A off the ring rows stage V is still reading)

		// ------------------------------------------------------------------ A: apply carry, nudge, horizontal filter
		if (r <= 511) {
			int x[16];
			if (sharpen) {
				// carry class entering this row: walk the maps of the strip's rows above it
				int cls = S.strip_carry[i & 1], cls_prev = 0;
				for (int j = 0; j < warp; j++) {
					cls_prev = cls;
					CMap rm = {S.rowmap_lo[j], S.rowmap_hi[j]};
					cls = cmap_at(rm, cls);
				}
				// flag left by the last pair (509,510) of the row above
				int flag;
				if (r == 1) flag = 0;
				else if (warp == 0) flag = S.strip_flag[i & 1];
				else {
					CMap pm = {S.part_lo[warp - 1], S.part_hi[warp - 1]};
					int c9 = cmap_at(pm, cls_prev);
					const int k509 = carry_step(S.e509[warp - 1], c9);
					const int k510 = carry_step(S.e510[warp - 1], c9);
					flag = pair_flag(k509, k510);
				}
				int cin = lane ? cmap_at(excl, cls) : cls;
				int K[16];
#pragma unroll
				for (int t = 0; t < 16; t++) {
					const bool outside = (t == 0 && lane == 0) || (t == 15 && lane == 31);
					K[t] = outside ? 0 : carry_step(e[t], cin);
				}
				if (warp == 15 && lane == 31) S.strip_carry[(i + 1) & 1] = cin;
				const int knext = __shfl_down_sync(0xffffffffu, K[0], 1);
				// pair rule through its tables: interval of each value, then one entry per pair
				int cat[17];
#pragma unroll
				for (int t = 1; t < 17; t++) {
					const int kv = t < 16 ? K[t] : knext;
					cat[t] = S.pcat[max(min(kv, 255), -255) + 256];
				}
				uint32_t ent[8];
#pragma unroll
				for (int p = 0; p < 8; p++) ent[p] = S.plut[cat[1 + 2 * p] * PAIR_CATS + cat[2 + 2 * p]];
				if (lane == 31) ent[7] = 0x92u;   // there is no pair (511, 512)
				int a = __shfl_up_sync(0xffffffffu, (int)(ent[7] >> 9), 1);
				if (lane == 0) a = flag;
				if (warp == 15 && lane == 31) S.strip_flag[(i + 1) & 1] = (int)(ent[6] >> 9);
				int d[17];
#pragma unroll
				for (int p = 0; p < 8; p++) {
					d[1 + 2 * p] = (int)((ent[p] >> (3 * a)) & 7u) - 2;
					d[2 + 2 * p] = (int)((ent[p] >> 6) & 7u) - 2;
					a = (int)(ent[p] >> 9);
				}
				const int dprev = __shfl_up_sync(0xffffffffu, d[16], 1);
				d[0] = lane ? dprev : 0;
				unpack16(ymid, x);
#pragma unroll
				for (int t = 0; t < 16; t++) x[t] += d[t];
			} else if (PLANE_IN) {
				load16(YP + r * 512 + 16 * lane, x);
			} else {
				unpack16(reinterpret_cast<const uint4 *>(S.yring[r % RING])[lane], x);
			}
			row_pass_regs(x, lane, S.rring[r % RING], r % RING < 19);
			if (i == 0 && warp == 0) {   // row 0 is outside the sharpening window
				if (PLANE_IN) load16(YP + 16 * lane, x);
				else unpack16(reinterpret_cast<const uint4 *>(S.yring[0])[lane], x);
				row_pass_regs(x, lane, S.rring[0], true);
			}
		}
		__syncthreads();

		// ------------------------------------------------------------------ V: vertical filter, outputs 8i .. 8i+7
		{
			const int k = tid, e0 = 8 * i;
			int col[19];
			const int16_t *win = &S.rring[(16 * i + RING - 2) % RING][k];   // rows 16i-2 .. 16i+16, never wraps
#pragma unroll
			for (int j = 0; j < 19; j++) col[j] = win[j * 512];
			if (i == 0) { col[0] = col[4]; col[1] = col[3]; }   // rows -2, -1 mirror rows 2, 1
			if (i == 31) col[18] = col[16];                     // row 512 mirrors row 510
			int lo[8], hi[8];
			const bool fine = k < 256;
			if (kept && fine) {
				// q22/q23: the low half of the first pass is kept, transposed, for the res6 side channel
				// (im_quality_setting, encoder/wavelet_filterbank.c:107-112): rows 16i .. 16i+15 of column k
				uint4 *dst = reinterpret_cast<uint4 *>(kept + (size_t)img * kstride + k * 512 + 16 * i);
				dst[0] = make_uint4(pack2(col[2], col[3]), pack2(col[4], col[5]), pack2(col[6], col[7]), pack2(col[8], col[9]));
				dst[1] = make_uint4(pack2(col[10], col[11]), pack2(col[12], col[13]), pack2(col[14], col[15]), pack2(col[16], col[17]));
			}
			col_pass8(col, e0, fine, i == 31, v_rem, lo, hi);
			const uint4 H = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
			*reinterpret_cast<uint4 *>(P + k * 512 + 256 + e0) = H;
			if (fine) {
#pragma unroll
				for (int s = 0; s < 8; s++) LL[(e0 + s) * 256 + k] = (int16_t)lo[s];
			} else {
				const uint4 L = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
				*reinterpret_cast<uint4 *>(P + k * 512 + e0) = L;
			}
		}
		// no barrier here: the next strip's C and E stages touch neither rring nor anything V reads,
		// and two barriers separate this V from the next A.
	}
}

// =====================================================================================
// k_dwt_level: one analysis level of an N x N band, whole band in shared memory.
//   in  : band in natural orientation, in[m * in_stride + k]  (m = line, k = position along it)
//   out : coefficient plane in the reference's transposed orientation, out[k2 * out_stride + m2]
//   ll  : if not NULL the LL quadrant goes there instead, in natural orientation [m2][k2]
//         (dense N/2 x N/2): the reference's `res256` copy, which the next level reads.
// =====================================================================================
template <int N, typename InT, int THREADS>
__global__ void __launch_bounds__(THREADS) k_dwt_level(const InT *__restrict__ in, size_t in_slot, int in_stride,
                                                       int16_t *__restrict__ out, size_t out_slot, int out_stride,
                                                       int16_t *__restrict__ ll, size_t ll_slot,
                                                       int16_t *__restrict__ snap = nullptr, size_t snap_slot = 0)
{
	// snap (only with ll == NULL): a second, dense (row stride N) copy of the N x N output -- the closed loop's snapshot
	// of the level-2 region (`resIII`, encoder/nhw_encoder.c:623-631), which used to be a copy kernel of its own
	extern __shared__ __align__(16) int16_t band[];   // [N][N]
	constexpr int H = N / 2;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const InT *src = in + (size_t)blockIdx.x * in_slot;
	int16_t *dst = out + (size_t)blockIdx.x * out_slot;
	// ---- load (8 elements per thread and step)
	for (int idx = tid; idx < N * N / 8; idx += THREADS) {
		const int m = idx / (N / 8), c = idx % (N / 8);
		uint4 v;
		if (sizeof(InT) == 2) {
			v = *reinterpret_cast<const uint4 *>(src + (size_t)m * in_stride + c * 8);
		} else {
			const uint2 b = *reinterpret_cast<const uint2 *>(src + (size_t)m * in_stride + c * 8);
			v.x = (b.x & 255u) | (((b.x >> 8) & 255u) << 16);
			v.y = ((b.x >> 16) & 255u) | ((b.x >> 24) << 16);
			v.z = (b.y & 255u) | (((b.y >> 8) & 255u) << 16);
			v.w = ((b.y >> 16) & 255u) | ((b.y >> 24) << 16);
		}
		reinterpret_cast<uint4 *>(band + m * N)[c] = v;
	}
	__syncthreads();
	// ---- horizontal pass in place, one warp per line: lane holds PER consecutive inputs
	constexpr int PER = N / 32, OUTS = PER / 2;
	for (int m = warp; m < N; m += THREADS / 32) {
		int16_t *row = band + m * N;
		int x[PER];
#pragma unroll
		for (int t = 0; t < PER; t++) x[t] = row[lane * PER + t];
		int xm2 = __shfl_up_sync(0xffffffffu, x[PER - 2], 1), xm1 = __shfl_up_sync(0xffffffffu, x[PER - 1], 1);
		int xp = __shfl_down_sync(0xffffffffu, x[0], 1);
		if (lane == 0) { xm2 = x[2]; xm1 = x[1]; }
		if (lane == 31) xp = x[PER - 2];
		int lo[OUTS], hi[OUTS];
#pragma unroll
		for (int s = 0; s < OUTS; s++) {
			const int a2 = s ? x[2 * s - 2] : xm2, a1 = s ? x[2 * s - 1] : xm1;
			const int b2 = s < OUTS - 1 ? x[2 * s + 2] : xp;
			lo[s] = 6 * x[2 * s] + 2 * (a1 + x[2 * s + 1]) - (a2 + b2);
			hi[s] = 2 * x[2 * s + 1] - (x[2 * s] + b2);
		}
		if (lane == 31) hi[OUTS - 1] = (x[PER - 1] - x[PER - 2]) << 1;
		__syncwarp();
#pragma unroll
		for (int s = 0; s < OUTS; s++) {
			row[lane * OUTS + s] = (int16_t)lo[s];
			row[H + lane * OUTS + s] = (int16_t)hi[s];
		}
	}
	__syncthreads();
	// ---- vertical pass: task = (group of 8 outputs, column k); lanes run along k
	int16_t *llp = ll ? ll + (size_t)blockIdx.x * ll_slot : nullptr;
	int16_t *sn = snap ? snap + (size_t)blockIdx.x * snap_slot : nullptr;
	for (int task = tid; task < N * (H / 8); task += THREADS) {
		const int k = task % N, g = task / N, e0 = 8 * g;
		const bool fine = k < H;
		int col[19];
#pragma unroll
		for (int j = 0; j < 19; j++) {
			int y = 2 * e0 - 2 + j;
			if (y < 0) y = -y;
			if (y > N - 1) y = N - 2;
			col[j] = band[y * N + k];
		}
		int rem = 0;
		if (fine && e0 > 0) {
			const int y4 = band[(2 * e0 - 4) * N + k], y3 = band[(2 * e0 - 3) * N + k];
			rem = vi_remainder(6 * col[0] + 2 * (y3 + col[1]) - (y4 + col[2]));
		}
		int lo[8], hi[8];
		col_pass8(col, e0, fine, e0 + 8 == H, rem, lo, hi);
		const uint4 Hh = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
		*reinterpret_cast<uint4 *>(dst + (size_t)k * out_stride + H + e0) = Hh;
		if (sn) *reinterpret_cast<uint4 *>(sn + k * N + H + e0) = Hh;
		if (fine && llp) {
#pragma unroll
			for (int s = 0; s < 8; s++) llp[(e0 + s) * H + k] = (int16_t)lo[s];
		} else {
			const uint4 L = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
			*reinterpret_cast<uint4 *>(dst + (size_t)k * out_stride + e0) = L;
			if (sn) *reinterpret_cast<uint4 *>(sn + k * N + e0) = L;
		}
	}
}

// every RGB triple through both forms of the q>=20 colour transform; counts disagreements
__global__ void k_color_check(ColorParams p, unsigned long long *bad)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const int c0 = t & 255u, c1 = (t >> 8) & 255u, c2 = t >> 16;
	int Y, U, V, y, u, v;
	rgb_to_ycc(c0, c1, c2, p, Y, U, V);
	rgb_to_ycc_q20(c0, c1, c2, y, u, v);
	if (Y != y || U != u || V != v) atomicAdd(bad, 1ull);
	// the packed dp2a form the fused kernel uses (four pixels per call: this triple and three neighbours)
	const uint32_t px[4] = {t, (t + 1u) & 0xffffffu, (t * 2654435761u) & 0xffffffu, (t ^ 0x5a5a5au) & 0xffffffu};
	int Y4[4];
	uint32_t uv4[4];
	rgb4px_to_ycc_q20(px, Y4, uv4);
	if (Y4[0] != Y || (int)(uv4[0] & 0xffffu) != U || (int)(uv4[0] >> 16) != V) atomicAdd(bad, 1ull);
}

template <int N, typename InT, int THREADS>
void launch_level(nhw_ctx *c, const char *label, int n_planes, const InT *in, size_t in_slot, int in_stride, int16_t *out,
                  size_t out_slot, int out_stride, int16_t *ll, size_t ll_slot, int16_t *snap = nullptr, size_t snap_slot = 0)
{
	NHW_LAUNCH_L(c, label, (k_dwt_level<N, InT, THREADS>), n_planes, THREADS, N * N * 2, in, in_slot, in_stride, out, out_slot,
	             out_stride, ll, ll_slot, snap, snap_slot);
}

}  // namespace

namespace nhw {

ColorParams color_params(int quality);

bool front_device_init(nhw_ctx *c)
{
	bool ok = check(cudaFuncSetAttribute(k_front_luma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontSmem)), "attr k_front_luma");
	ok = ok && check(cudaFuncSetAttribute(k_front_luma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontSmem)), "attr k_front_luma<plane>");
	ok = ok && check(cudaFuncSetAttribute(k_dwt_level<256, int16_t, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256 * 2), "attr k_dwt_level");
	ok = ok && check(cudaFuncSetAttribute(k_dwt_level<256, uint8_t, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256 * 2), "attr k_dwt_level");
	ok = ok && check(cudaFuncSetAttribute(k_dwt_level<128, int16_t, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 2), "attr k_dwt_level");
	uint8_t cat[512];
	uint16_t lut[PAIR_CATS * PAIR_CATS];
	pair_build_tables(cat, lut);
	ok = ok && check(cudaMemcpyToSymbol(g_pair_cat, cat, sizeof(cat)), "pair tables");
	ok = ok && check(cudaMemcpyToSymbol(g_pair_lut, lut, sizeof(lut)), "pair tables");
	(void)c;
	return ok;
}

// The whole front end of n images: rgb -> luma coefficient planes (both levels) + `res256`,
// chroma coefficient planes (both levels) + chroma `res256`.  uv_bytes: scratch, 2*65536 B / image.
void front_fused(nhw_ctx *c, const uint8_t *rgb, int n, int quality, int16_t *y_proc, size_t ypstride, int16_t *y_ll1,
                 size_t ylstride, uint8_t *uv_bytes, int16_t *c_proc, size_t cpstride, int16_t *c_ll1, size_t clstride,
                 int16_t *kept, size_t kstride)
{
	const ColorParams p = color_params(quality);
	if (quality <= 16) {
		// colour (integer form) -> luma plane + 4:2:0 bytes, the pre-sharpening state machine in place on the plane,
		// then the same level-1 analysis fed from the plane.  The plane lives in the context's im_jpeg slots (the
		// walkers use y_proc / y_aux / y_aux2 as scratch, all dead at this point).
		int16_t *yplane = c->y_jpeg + NHW_GUARD_S;
		const size_t ys = NHW_Y_SLOT;
		colorspace(c, rgb, n, quality, yplane, ys, uv_bytes, uv_bytes + NHW_CPLANE, (size_t)2 * NHW_CPLANE);
		pre_processing_lowq(c, n, quality, yplane, ys);
		NHW_LAUNCH_L(c, "k_front_luma<plane>", k_front_luma<true>, n, FT, sizeof(FrontSmem), (const uint8_t *)nullptr, yplane, ys,
		             y_proc, ypstride, y_ll1, ylstride, (uint8_t *)nullptr, (size_t)0, p, 0, (int16_t *)nullptr, kstride);
		launch_level<256, int16_t, 1024>(c, "k_dwt_level<256>", n, y_ll1, ylstride, 256, y_proc, ypstride, 512, nullptr, 0);
		if (quality <= 14) {
			// pre_processing_UV: the filtered samples leave 0..255, so the planes go through int16 (chroma im_jpeg slots)
			int16_t *cj = c->c_jpeg + NHW_GUARD_S;
			chroma_pre_uv(c, 2 * n, quality, uv_bytes, cj, (size_t)NHW_C_SLOT);
			launch_level<256, int16_t, 1024>(c, "k_dwt_level<256>", 2 * n, cj, (size_t)NHW_C_SLOT, 256, c_proc, cpstride, 256, c_ll1, clstride);
		} else {
			launch_level<256, uint8_t, 1024>(c, "k_dwt_level<256,u8>", 2 * n, uv_bytes, (size_t)NHW_CPLANE, 256, c_proc, cpstride, 256,
			                                 c_ll1, clstride);
		}
		chroma_thresholds(c, 2 * n, c_proc, cpstride, 8);
		launch_level<128, int16_t, 256>(c, "k_dwt_level<128>", 2 * n, c_ll1, clstride, 128, c_proc, cpstride, 256, nullptr, 0);
		return;
	}
	NHW_LAUNCH_L(c, "k_front_luma", k_front_luma<false>, n, FT, sizeof(FrontSmem), rgb, (const int16_t *)nullptr, (size_t)0, y_proc,
	             ypstride, y_ll1, ylstride, uv_bytes, (size_t)2 * NHW_CPLANE, p, quality < 22 ? 1 : 0,
	             quality > 21 ? kept : (int16_t *)nullptr, kstride);
	launch_level<256, int16_t, 1024>(c, "k_dwt_level<256>", n, y_ll1, ylstride, 256, y_proc, ypstride, 512, nullptr, 0);
	launch_level<256, uint8_t, 1024>(c, "k_dwt_level<256,u8>", 2 * n, uv_bytes, (size_t)NHW_CPLANE, 256, c_proc, cpstride, 256,
	                                 c_ll1, clstride);
	launch_level<128, int16_t, 256>(c, "k_dwt_level<128>", 2 * n, c_ll1, clstride, 128, c_proc, cpstride, 256, nullptr, 0);
}

// test hook: number of RGB triples (of 2^24) on which the integer colour path differs from the IEEE one
long color_fast_path_mismatches(nhw_ctx *c)
{
	unsigned long long *bad = nullptr, host = ~0ull;
	if (cudaMalloc(&bad, 8) != cudaSuccess) return -1;
	cudaMemsetAsync(bad, 0, 8, c->stream);
	k_color_check<<<65536, 256, 0, c->stream>>>(color_params(20), bad);
	cudaMemcpyAsync(&host, bad, 8, cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	cudaFree(bad);
	return (long)host;
}

// one more analysis level of a band held in `im_jpeg` orientation (closed loop,
// encoder/nhw_encoder.c:281,2339)
void dwt_level_from_jpeg(nhw_ctx *c, int n_planes, const int16_t *jpeg, size_t jstride, int16_t *proc, size_t pstride,
                         int N, int row_stride, int16_t *snap, size_t snap_slot)
{
	if (N == 256)
		launch_level<256, int16_t, 1024>(c, "k_dwt_level<256>", n_planes, jpeg, jstride, row_stride, proc, pstride, row_stride, nullptr, 0, snap, snap_slot);
	else
		launch_level<128, int16_t, 256>(c, "k_dwt_level<128>", n_planes, jpeg, jstride, row_stride, proc, pstride, row_stride, nullptr, 0, snap, snap_slot);
}

}  // namespace nhw
