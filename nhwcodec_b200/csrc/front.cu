// front.cu -- encoder front end, stage by stage: colour transform + 4:2:0 and luma pre-sharpening.
//
// What it computes (reference behaviour, re-designed for a batch on one GPU):
//   colorspace      : downsample_YUV420            encoder/colorspace.c:55-260
//   pre_processing  : pre_processing (q17..q21)    encoder/image_processing.c:558-836,1926-1990
//
// These are the stage-level kernels behind nhw_stage_colorspace_device (parity tests of the
// colour and pre-sharpening stages against the reference's taps).  The encoder itself runs
// the fused front end in front_fused.cu, which shares the per-element arithmetic
// (color_core.cuh, pre_core.cuh, dwt_core.cuh).
#include "nhw_ctx.h"
#include "nhw_dev.cuh"
#include "enc_img.cuh"
#include "dwt_core.cuh"
#include "color_core.cuh"
#include "pre_core.cuh"
#include "pre_lowq.cuh"
#include "enc_lowq.cuh"

namespace {

// =====================================================================================
// colour
// =====================================================================================
#define CS_ROWS 8   // image rows per CTA (-> 4 chroma rows), plus one halo row above

__global__ void __launch_bounds__(256) k_colorspace(const uint8_t *__restrict__ rgb, int16_t *__restrict__ yout,
                                                    uint8_t *__restrict__ uout, uint8_t *__restrict__ vout,
                                                    size_t ystride, size_t cstride, ColorParams p)
{
	__shared__ uint8_t su[CS_ROWS + 1][512];
	__shared__ uint8_t sv[CS_ROWS + 1][512];
	const int img = blockIdx.y;
	const int y0 = blockIdx.x * CS_ROWS;
	const uint8_t *src = rgb + (size_t)img * NHW_RGB_BYTES;
	int16_t *yp = yout ? yout + (size_t)img * ystride : nullptr;

	for (int idx = threadIdx.x; idx < (CS_ROWS + 1) * 512; idx += 256) {
		int row = idx >> 9, x = idx & 511;
		int y = y0 - 1 + row;
		if (y < 0) continue;
		const uint8_t *px = src + ((size_t)y * 512 + x) * 3;
		int Y, U, V;
		rgb_to_ycc(px[0], px[1], px[2], p, Y, U, V);
		if (row > 0 && yp) yp[y * 512 + x] = (int16_t)Y;
		su[row][x] = (uint8_t)U;
		sv[row][x] = (uint8_t)V;
	}
	__syncthreads();
	// [1 2 1]/4 horizontally on even pixels, then [1 2 1]/4 vertically + 2:1 decimation
	// (encoder/colorspace.c:220-256; first pixel / first row use (a+b+1)>>1).
	for (int o = threadIdx.x; o < (CS_ROWS / 2) * 256; o += 256) {
		int rr = o >> 8, cx = o & 255, x = cx * 2;
		int r = (y0 >> 1) + rr;
		int hu[3], hv[3];
#pragma unroll
		for (int k = 0; k < 3; k++) {
			int row = 2 * rr + k;
			if (x == 0) {
				hu[k] = (su[row][0] + su[row][1] + 1) >> 1;
				hv[k] = (sv[row][0] + sv[row][1] + 1) >> 1;
			} else {
				hu[k] = (su[row][x - 1] + 2 * su[row][x] + su[row][x + 1] + 2) >> 2;
				hv[k] = (sv[row][x - 1] + 2 * sv[row][x] + sv[row][x + 1] + 2) >> 2;
			}
		}
		int U, V;
		if (r == 0) {
			U = (hu[1] + hu[2] + 1) >> 1;
			V = (hv[1] + hv[2] + 1) >> 1;
		} else {
			U = (hu[0] + 2 * hu[1] + hu[2] + 2) >> 2;
			V = (hv[0] + 2 * hv[1] + hv[2] + 2) >> 2;
		}
		if (uout) uout[(size_t)img * cstride + r * 256 + cx] = (uint8_t)U;
		if (vout) vout[(size_t)img * cstride + r * 256 + cx] = (uint8_t)V;
	}
}

// =====================================================================================
// luma pre-sharpening, q17..q21
// =====================================================================================
// Loop A of the reference (image_processing.c:601-764) carries `res4` (4 bits) through the
// whole image in raster order.  The carry only enters the next element as (res4+2)>>2, one of
// five classes, so each element is a map {class} -> {4-bit state}; maps compose, which turns
// the raster recurrence into a scan: per-lane maps -> warp scan -> per-row map -> a 510-step
// chain per image -> re-apply.  A map is packed as five nibbles.
__device__ __forceinline__ unsigned cmap_apply(unsigned m, int cls) { return (m >> (4 * cls)) & 15u; }
__device__ __forceinline__ int carry_class(unsigned state) { return (int)((state + 2u) >> 2); }
__device__ __forceinline__ unsigned cmap_compose(unsigned first, unsigned then)
{
	unsigned r = 0;
#pragma unroll
	for (int c = 0; c < 5; c++) r |= cmap_apply(then, carry_class(cmap_apply(first, c))) << (4 * c);
	return r;
}

// signed energy of one element: sign(res) * (15*|res| + count); 0 resets the carry.
__device__ __forceinline__ int lap_energy(const int16_t *up, const int16_t *mid, const int16_t *dn, int x)
{
	int c = mid[x];
	int w1 = c - mid[x - 1], w2 = c - mid[x + 1], w3 = c - up[x], w4 = c - dn[x];
	int w5 = c - up[x + 1], w6 = c - up[x - 1], w7 = c - dn[x - 1], w8 = c - dn[x + 1];
	int res = w1 + w2 + w3 + w4 + w5 + w6 + w7 + w8;
	int cnt = nhw_iabs(w1) + nhw_iabs(w2) + nhw_iabs(w3) + nhw_iabs(w4) + nhw_iabs(w5) + nhw_iabs(w6) +
	          nhw_iabs(w7) + nhw_iabs(w8);
	if (res == 0) return 0;
	int e = 15 * nhw_iabs(res) + cnt;
	return res < 0 ? -e : e;
}

// map of a run of 16 signed energies (positions outside 1..510 are skipped)
__device__ __forceinline__ unsigned lane_map(const int *e, int x0)
{
	unsigned m = 0;
#pragma unroll
	for (int c = 0; c < 5; c++) {
		int cls = c;
		unsigned st = 0;
		bool any = false;
#pragma unroll
		for (int t = 0; t < 16; t++) {
			int x = x0 + t;
			if (x < 1 || x > 510) continue;
			any = true;
			st = e[t] == 0 ? 0u : (unsigned)((nhw_iabs(e[t]) + cls) & 15);
			cls = carry_class(st);
		}
		(void)any;
		m |= st << (4 * c);
	}
	return m;
}

#define PRE_WARPS 4

__global__ void __launch_bounds__(32 * PRE_WARPS) k_pre_energy(const int16_t *__restrict__ y, int16_t *__restrict__ energy,
                                                               uint32_t *__restrict__ rowmap, size_t ystride, size_t astride)
{
	__shared__ __align__(16) int16_t rows[PRE_WARPS][3][512];
	const int img = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int r = 1 + blockIdx.x * PRE_WARPS + warp;
	if (r > 510) return;
	const int16_t *src = y + (size_t)img * ystride;
	for (int k = 0; k < 3; k++) {
		const int4 *s4 = reinterpret_cast<const int4 *>(src + (r - 1 + k) * 512);
		int4 a = s4[lane * 2], b = s4[lane * 2 + 1];
		int4 *d4 = reinterpret_cast<int4 *>(&rows[warp][k][0]);
		d4[lane * 2] = a;
		d4[lane * 2 + 1] = b;
	}
	__syncwarp();
	const int16_t *up = rows[warp][0], *mid = rows[warp][1], *dn = rows[warp][2];
	int e[16];
	const int x0 = lane * 16;
#pragma unroll
	for (int t = 0; t < 16; t++) {
		int x = x0 + t;
		e[t] = (x >= 1 && x <= 510) ? lap_energy(up, mid, dn, x) : 0;
	}
	int16_t *dst = energy + (size_t)img * astride + r * 512 + x0;
#pragma unroll
	for (int t = 0; t < 16; t++) dst[t] = (int16_t)e[t];
	unsigned m = lane_map(e, x0);
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned o = __shfl_up_sync(0xffffffffu, m, d);
		if (lane >= d) m = cmap_compose(o, m);
	}
	if (lane == 31) rowmap[(size_t)img * 512 + r] = m;
}

// one thread per image: chain the 510 row maps, emit the carry class entering each row
__global__ void k_pre_chain(const uint32_t *__restrict__ rowmap, uint8_t *__restrict__ rowcarry, int n)
{
	int img = blockIdx.x * blockDim.x + threadIdx.x;
	if (img >= n) return;
	int cls = 0;
	for (int r = 1; r <= 510; r++) {
		rowcarry[(size_t)img * 512 + r] = (uint8_t)cls;
		cls = carry_class(cmap_apply(rowmap[(size_t)img * 512 + r], cls));
	}
}

// final kernel values nhw_kernel[r][x] = sign * ((15|res| + count + carry) >> 4)
__global__ void __launch_bounds__(32 * PRE_WARPS) k_pre_apply(const int16_t *__restrict__ energy,
                                                              const uint8_t *__restrict__ rowcarry,
                                                              int16_t *__restrict__ kern, size_t astride)
{
	const int img = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int r = 1 + blockIdx.x * PRE_WARPS + warp;
	if (r > 510) return;
	const int x0 = lane * 16;
	const int16_t *src = energy + (size_t)img * astride + r * 512 + x0;
	int e[16];
	{
		const int4 *s4 = reinterpret_cast<const int4 *>(src);
		int4 a = s4[0], b = s4[1];
		int w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
		for (int t = 0; t < 8; t++) {
			e[2 * t] = (int)(int16_t)(w[t] & 0xffff);
			e[2 * t + 1] = w[t] >> 16;
		}
	}
	unsigned m = lane_map(e, x0);
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned o = __shfl_up_sync(0xffffffffu, m, d);
		if (lane >= d) m = cmap_compose(o, m);
	}
	unsigned excl = __shfl_up_sync(0xffffffffu, m, 1);
	int cls = rowcarry[(size_t)img * 512 + r];
	if (lane > 0) cls = carry_class(cmap_apply(excl, cls));
	int16_t out[16];
#pragma unroll
	for (int t = 0; t < 16; t++) {
		int x = x0 + t;
		out[t] = 0;
		if (x < 1 || x > 510) continue;
		if (e[t] == 0) {
			cls = 0;   // state 0 -> class 0
		} else {
			int v = nhw_iabs(e[t]) + cls;
			int k = v >> 4;
			out[t] = (int16_t)(e[t] < 0 ? -k : k);
			cls = carry_class((unsigned)(v & 15));
		}
	}
	int16_t *dst = kern + (size_t)img * astride + r * 512 + x0;
#pragma unroll
	for (int t = 0; t < 16; t++) dst[t] = out[t];
}

__global__ void __launch_bounds__(256) k_pre_nudge(const int16_t *__restrict__ kern, int16_t *__restrict__ y, size_t astride, size_t ystride)
{
	const int img = blockIdx.y, r = 1 + blockIdx.x, p = threadIdx.x;
	if (p >= 255) return;
	const int16_t *k = kern + (size_t)img * astride + r * 512;
	int j = 1 + 2 * p;
	int res = k[j], cnt = k[j + 1];
	int a;
	if (p > 0) a = pair_flag(k[j - 2], k[j - 1]);
	else if (r > 1) a = pair_flag(k[-512 + 509], k[-512 + 510]);
	else a = 0;
	int d0, d1;
	pair_nudge(res, cnt, a, d0, d1);
	int16_t *dst = y + (size_t)img * ystride + r * 512 + j;
	if (d0) dst[0] = (int16_t)(dst[0] + d0);
	if (d1) dst[1] = (int16_t)(dst[1] + d1);
}

// =====================================================================================
// q <= 16
// =====================================================================================
// luma pre-sharpening state machine (pre_lowq.cuh).  Walk A's kernel values are the q > 16 recurrence and come from the
// scan kernels above (k_pre_energy / k_pre_chain / k_pre_apply); what is left is raster-serial per image and runs one
// WARP per image: lane 0 walks, all lanes stage the rows it walks through shared memory (coalesced 16-byte loads and
// stores), so the walker never waits for global memory, and the parts that are parallel inside an image (the q <= 14
// smoothing of a row, walk D's rows) use all lanes.
//   events : bitmap of the pixels walk A's marker rules have to look at (pre_low_is_event), one bit per pixel
//   walk   : marker rules over the event pixels -> walk B row by row -> walk C on a two-row window -> walk D
#define PLW_WARPS 4
__global__ void __launch_bounds__(256) k_pre_low_events(const int16_t *__restrict__ kern, size_t kstride, uint32_t *__restrict__ bits,
                                                        size_t bstride, int s2)
{
	const int s = blockIdx.x * 256 + threadIdx.x, r = s >> 9, j = s & 511;
	const bool ev = r >= 1 && r <= 510 && j >= 1 && j <= 510 && pre_low_is_event(kern[(size_t)blockIdx.y * kstride + s], s2);
	const uint32_t m = __ballot_sync(0xffffffffu, ev);
	if ((threadIdx.x & 31) == 0) bits[(size_t)blockIdx.y * bstride + (s >> 5)] = m;
}

__global__ void __launch_bounds__(32 * PLW_WARPS) k_pre_low_walk(int16_t *__restrict__ y, size_t ystride, const int16_t *__restrict__ copy,
                                                                   int16_t *__restrict__ kern, int16_t *__restrict__ marks, size_t astride,
                                                                   int n, int q, int phases)
{
	__shared__ __align__(16) int16_t sYa[PLW_WARPS][2 * PW], sKa[PLW_WARPS][2 * PW];
	__shared__ __align__(16) uint8_t sMa[PLW_WARPS][2 * PW];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, img = blockIdx.x * PLW_WARPS + warp;
	if (img >= n) return;
	int16_t *Y = y + (size_t)img * ystride, *K = kern + (size_t)img * astride;
	const int16_t *O = copy + (size_t)img * astride;
	uint8_t *M = reinterpret_cast<uint8_t *>(marks + (size_t)img * astride);
	const uint32_t *bits = reinterpret_cast<const uint32_t *>(M + PW * PW);
	int16_t *sY = sYa[warp], *sK = sKa[warp];
	uint8_t *sM = sMa[warp];
	const PreLowParams p = pre_low_params(q);
	// rows move between the plane and the window as 32 lanes x 16 cells (int16: two 16-byte words, marks: one)
	auto load_row16 = [&](int16_t *dst, const int16_t *src) {
		reinterpret_cast<uint4 *>(dst)[2 * lane] = reinterpret_cast<const uint4 *>(src)[2 * lane];
		reinterpret_cast<uint4 *>(dst)[2 * lane + 1] = reinterpret_cast<const uint4 *>(src)[2 * lane + 1];
	};
	auto load_row8 = [&](uint8_t *dst, const uint8_t *src) { reinterpret_cast<uint4 *>(dst)[lane] = reinterpret_cast<const uint4 *>(src)[lane]; };

	// ---- walk A, marker rules: the event pixels in raster order
	if (phases & 1) {
		PreWalkA w = {0, 0, 0, 0, 0, 0, 0, 0, 0};
		for (int w0 = 0; w0 < PW * PW / 32; w0 += 32) {
			const uint32_t mine = bits[w0 + lane];
			uint32_t any = __ballot_sync(0xffffffffu, mine != 0);
			while (any) {
				const int src = __ffs(any) - 1;
				any &= any - 1;
				uint32_t m = __shfl_sync(0xffffffffu, mine, src);
				if (lane == 0)
					for (; m; m &= m - 1) {
						const int s = 32 * (w0 + src) + __ffs(m) - 1;
						pre_low_event(w, p, O, K, s, s & 511);
					}
			}
		}
		__syncwarp();
	}
	// ---- walk B, row by row through the window's second row
	if (phases & 2) {
		PairThrottle t;
		t.init();
		int a = 0;
		for (int r = 1; r < 511; r++) {
			load_row16(sY + PW, Y + r * PW);
			load_row16(sK + PW, K + r * PW);
			reinterpret_cast<uint4 *>(sM + PW)[lane] = make_uint4(0, 0, 0, 0);
			__syncwarp();
			if (p.smooth_on) {
				for (int k = 0; k < 16; k++) {
					const int j = 16 * lane + k;
					int v;
					if (j >= 1 && j <= 510 && pre_low_smooth_cell(O, r * PW + j, sK[PW + j], p, v)) sY[PW + j] = (int16_t)v;
				}
				__syncwarp();
			}
			if (lane == 0) pre_low_walk_b_row(t, a, p, r, sY + PW, sK + PW, sM + PW);
			__syncwarp();
			load_row16(Y + r * PW, sY + PW);
			load_row16(K + r * PW, sK + PW);
			load_row8(M + r * PW, sM + PW);
			__syncwarp();
		}
	}
	// ---- walk C on the two-row window (rows r - 1, r)
	if (phases & 4) {
		PreWalkC w = {0, 0, 0, 0, 0, 0};
		for (int r = 1; r < 511; r++) {
			for (int h = 0; h < 2; h++) {
				load_row16(sY + h * PW, Y + (r - 1 + h) * PW);
				load_row16(sK + h * PW, K + (r - 1 + h) * PW);
				load_row8(sM + h * PW, M + (r - 1 + h) * PW);
			}
			__syncwarp();
			if (lane == 0) pre_low_walk_c_row(w, p, r, sY, sK, sM);
			__syncwarp();
			for (int h = 0; h < 2; h++) {
				load_row16(Y + (r - 1 + h) * PW, sY + h * PW);
				load_row8(M + (r - 1 + h) * PW, sM + h * PW);
			}
			load_row16(K + r * PW, sK + PW);   // the row above is only read
			__syncwarp();
		}
	}
	// ---- walk D: rows are independent
	if (phases & 8) for (int r = 1 + lane; r < 511; r += 32) pre_low_walk_d_row(Y, K, M, p, r);
}

// chroma pre-filter (pre_processing_UV, q <= 14): 4:2:0 bytes -> int16 plane with the +-1 / +-2 nudges applied
__global__ void __launch_bounds__(256) k_c_pre_uv(const uint8_t *__restrict__ uv, int16_t *__restrict__ out, size_t oslot, int q)
{
	const uint8_t *src = uv + (size_t)blockIdx.y * NHW_CPLANE;
	int16_t *dst = out + (size_t)blockIdx.y * oslot;
	const int r = blockIdx.x, j = threadIdx.x;
	dst[r * 256 + j] = (int16_t)c_pre_uv_cell(src, q, r, j);
}

// chroma level-1 band thresholds (q <= 16): in place on the coefficient plane, outside the level-2 region
__global__ void __launch_bounds__(256) k_c_thresholds(int16_t *__restrict__ proc, size_t pslot, int ratio)
{
	int16_t *P = proc + (size_t)blockIdx.y * pslot;
	const int r = blockIdx.x, j = threadIdx.x;
	if (r < 128 && j < 128) return;
	const int v = P[r * 256 + j], w = c_threshold_cell(v, ratio, r, j);
	if (w != v) P[r * 256 + j] = (int16_t)w;
}

}  // namespace

namespace nhw {

void pre_processing_lowq(nhw_ctx *c, int n, int quality, int16_t *y, size_t ystride)
{
	const size_t AS = NHW_Y_SLOT;
	int16_t *copy = c->y_proc + NHW_GUARD_S, *kern = c->y_aux + NHW_GUARD_S, *scratch = c->y_aux2 + NHW_GUARD_S;
	const PreLowParams p = pre_low_params(quality);
	// plain kernel values (the remainder chain as a scan): scratch = signed energies, kern = values.  Rows 0 and 511 of
	// the kernel plane are outside every walk's writes but inside their reads: zero.
	dim3 grid((510 + PRE_WARPS - 1) / PRE_WARPS, n);
	NHW_LAUNCH(c, k_pre_energy, grid, 32 * PRE_WARPS, 0, y, scratch, c->rowmap, ystride, AS);
	NHW_LAUNCH(c, k_pre_chain, (n + 63) / 64, 64, 0, c->rowmap, c->rowcarry, n);
	NHW_LAUNCH(c, k_pre_apply, grid, 32 * PRE_WARPS, 0, scratch, c->rowcarry, kern, AS);
	cudaMemset2DAsync(kern, AS * 2, 0, PW * 2, n, c->stream);
	cudaMemset2DAsync(kern + 511 * PW, AS * 2, 0, PW * 2, n, c->stream);
	// the plane as it was (walk A's Laplacians and the smoothing read it while walk B rewrites the plane), the marks
	cudaMemcpy2DAsync(copy, AS * 2, y, ystride * 2, (size_t)PW * PW * 2, n, cudaMemcpyDeviceToDevice, c->stream);
	cudaMemset2DAsync(scratch, AS * 2, 0, PW * PW, n, c->stream);
	uint32_t *bits = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(scratch) + PW * PW);
	NHW_LAUNCH(c, k_pre_low_events, dim3(PW * PW / 256, n), 256, 0, kern, AS, bits, AS / 2, p.sharp2);
	NHW_LAUNCH_L(c, "k_pre_low_walk", k_pre_low_walk, (n + PLW_WARPS - 1) / PLW_WARPS, 32 * PLW_WARPS, 0, y, ystride, copy, kern, scratch,
	             AS, n, quality, c->tune.plw_phases);
}

void chroma_pre_uv(nhw_ctx *c, int n_planes, int quality, const uint8_t *uv, int16_t *out, size_t oslot)
{
	NHW_LAUNCH_L(c, "k_c_pre_uv", k_c_pre_uv, dim3(256, n_planes), 256, 0, uv, out, oslot, quality);
}

void chroma_thresholds(nhw_ctx *c, int n_planes, int16_t *proc, size_t pslot, int ratio)
{
	NHW_LAUNCH_L(c, "k_c_thresholds", k_c_thresholds, dim3(256, n_planes), 256, 0, proc, pslot, ratio);
}

ColorParams color_params(int quality)
{
	static const int qtz[17] = {0, 15900, 16500, 17100, 18000, 18820, 19670, 20640, 21540, 23540, 25570, 27522, 27830, 27607, 28786, 31262, 32375};
	ColorParams p;
	p.yq = 1.0;
	p.qtz = 0;
	if (quality >= 20) p.mode = 0;
	else if (quality >= 18) { p.mode = 1; p.yq = (double)(quality == 19 ? 0.975f : 0.93f); }
	else if (quality == 17) p.mode = 2;
	else { p.mode = 3; p.qtz = qtz[quality < 0 ? 0 : quality]; }
	return p;
}

void colorspace(nhw_ctx *c, const uint8_t *rgb, int n, int quality, int16_t *y, size_t ystride, uint8_t *u, uint8_t *v,
                size_t cstride)
{
	const ColorParams p = color_params(quality);
	NHW_LAUNCH(c, k_colorspace, dim3(512 / CS_ROWS, n), 256, 0, rgb, y, u, v, ystride, cstride, p);
}

void pre_processing(nhw_ctx *c, int n, int quality, int16_t *y, size_t ystride)
{
	if (quality <= 16) { pre_processing_lowq(c, n, quality, y, ystride); return; }
	// q17..q21 share one rule set; q>=22 never gets here
	dim3 grid((510 + PRE_WARPS - 1) / PRE_WARPS, n);
	int16_t *energy = c->y_aux2 + NHW_GUARD_S, *kern = c->y_aux + NHW_GUARD_S;
	NHW_LAUNCH(c, k_pre_energy, grid, 32 * PRE_WARPS, 0, y, energy, c->rowmap, ystride, (size_t)NHW_Y_SLOT);
	NHW_LAUNCH(c, k_pre_chain, (n + 63) / 64, 64, 0, c->rowmap, c->rowcarry, n);
	NHW_LAUNCH(c, k_pre_apply, grid, 32 * PRE_WARPS, 0, energy, c->rowcarry, kern, (size_t)NHW_Y_SLOT);
	NHW_LAUNCH(c, k_pre_nudge, dim3(510, n), 256, 0, kern, y, (size_t)NHW_Y_SLOT, ystride);
}

}  // namespace nhw
