// front.cu -- encoder front end: colour transform + 4:2:0, luma pre-sharpening and the
// two-level integer wavelet analysis.
//
// What it computes (reference behaviour, re-designed for a batch on one GPU):
//   colorspace      : downsample_YUV420            encoder/colorspace.c:55-260
//   pre_processing  : pre_processing (q17..q21)    encoder/image_processing.c:558-836,1926-1990
//   dwt_*           : wavelet_analysis             encoder/wavelet_filterbank.c:52-302
//                     downfilter53IV / 53VI / 53   encoder/filters.c:346-386,203-287,55-114
//
// Orientation note.  The reference filters rows, transposes, filters rows again, and leaves
// the coefficient plane transposed (row index = horizontal frequency index k, column index =
// vertical index m).  Here a level is: row pass in natural layout R[y][k], then one kernel
// that walks columns of R out of a shared-memory tile and writes P[k][m] directly -- the
// transpose is folded into the column pass instead of being two extra trips through memory.
// Level 2 needs no transpose at all (LL1 is consumed in the orientation level 1 left it in).
#include "nhw_ctx.h"
#include "nhw_dev.cuh"
#include "enc_img.cuh"
#include "dwt_core.cuh"

namespace {

// =====================================================================================
// colour
// =====================================================================================
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
// wavelet analysis
// =====================================================================================
// ---- level 1, row pass: X[y][x] -> R[y][k], k<N/2 low, k>=N/2 high.  One warp per row. ----
template <int N>
__global__ void __launch_bounds__(256) k_dwt_rows(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t in_stride, size_t out_stride)
{
	__shared__ __align__(16) int16_t srow[8][N + 8];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int y = blockIdx.x * 8 + warp;
	const int16_t *src = in + (size_t)blockIdx.y * in_stride + y * N;
	int16_t *dst = out + (size_t)blockIdx.y * out_stride + y * N;
	for (int i = lane; i < N / 8; i += 32)
		reinterpret_cast<int4 *>(srow[warp])[i] = reinterpret_cast<const int4 *>(src)[i];
	__syncwarp();
	const int16_t *s = srow[warp];
	auto ld = [&](int i) { return (int)s[i]; };
	for (int e = lane; e < N / 2; e += 32) {
		dst[e] = (int16_t)tap_low(ld, e, N);
		dst[N / 2 + e] = (int16_t)first_pass_high(ld, e, N);
	}
}

// ---- level 1, column pass + transpose: R[y][k] -> P[k][m].  CTA = 32 columns k. ----
template <int N>
__global__ void __launch_bounds__(256) k_dwt_cols_t(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t in_stride, size_t out_stride)
{
	extern __shared__ int16_t tile[];   // [N][33]
	const int k0 = blockIdx.x * 32;
	const int16_t *src = in + (size_t)blockIdx.y * in_stride;
	int16_t *dst = out + (size_t)blockIdx.y * out_stride;
	for (int i = threadIdx.x; i < N * 32; i += 256) {
		int yy = i >> 5, kk = i & 31;
		tile[yy * 33 + kk] = src[yy * N + k0 + kk];
	}
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int kk = warp; kk < 32; kk += 8) {
		const int k = k0 + kk;
		const bool fine = k < N / 2;
		auto ld = [&](int i) { return (int)tile[i * 33 + kk]; };
		for (int e = lane; e < N / 2; e += 32) {
			dst[k * N + e] = (int16_t)second_pass_low(ld, e, N, fine);
			dst[k * N + N / 2 + e] = (int16_t)second_pass_high(ld, e, N, fine);
		}
	}
}

// ---- level 2 (and any later level): whole LL band in shared memory, one CTA per plane. ----
// in_transposed: read the band as J[m][k] = in[k*stride + m] (level-1 output, P orientation);
// otherwise in[m*stride + k] (an `im_jpeg`-oriented band, as in the encoder's closed loop).
// Writes P2[k2][m2] into out (stride `stride`) and, if ll_copy != NULL, J (the reference's
// `res256`) as a dense NxN array.
template <int N>
__global__ void __launch_bounds__(N) k_dwt_level_smem(const int16_t *in, int16_t *out,
                                                      int16_t *__restrict__ ll_copy, size_t in_stride, size_t out_stride,
                                                      size_t ll_stride, int stride, int in_transposed)
{
	extern __shared__ int16_t sm[];
	constexpr int S = N + 2;            // padded stride: conflict-free along both axes
	int16_t *band = sm;                 // band[k*S + m] = J[m][k]
	int16_t *rowbuf = sm + N * S;       // 2 rows of N
	const int16_t *src = in + (size_t)blockIdx.x * in_stride;
	int16_t *dst = out + (size_t)blockIdx.x * out_stride;
	const int t = threadIdx.x;
	if (in_transposed) {
		for (int k = 0; k < N; k++) band[k * S + t] = src[k * stride + t];
	} else {
		for (int m = 0; m < N; m++) band[t * S + m] = src[m * stride + t];
	}
	__syncthreads();
	if (ll_copy) {
		int16_t *ll = ll_copy + (size_t)blockIdx.x * ll_stride;
		for (int m = 0; m < N; m++) ll[m * N + t] = band[t * S + m];
	}
	// thread t owns column m=t of the row pass (walks k), then output column t of the column pass.
	const int m = t;
	auto ldk = [&](int k) { return (int)band[k * S + m]; };
	for (int e = 0; e < N / 2; e++) {
		int lo = (int16_t)tap_low(ldk, e, N);
		int hi = first_pass_high(ldk, e, N);
		rowbuf[m] = (int16_t)lo;
		rowbuf[N + m] = (int16_t)hi;
		__syncthreads();
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const int16_t *rb = rowbuf + h * N;
			auto ld = [&](int i) { return (int)rb[i]; };
			const int k2 = h ? N / 2 + e : e;
			const bool fine = (h == 0);
			int v = (t < N / 2) ? second_pass_low(ld, t, N, fine) : second_pass_high(ld, t - N / 2, N, fine);
			dst[k2 * stride + t] = (int16_t)v;
		}
		__syncthreads();
	}
}

__global__ void k_u8_to_s16(const uint8_t *__restrict__ in, int16_t *__restrict__ out, size_t in_stride, size_t out_stride)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < NHW_CPLANE) out[(size_t)blockIdx.y * out_stride + i] = in[(size_t)blockIdx.y * in_stride + i];
}

}  // namespace

namespace nhw {

void colorspace(nhw_ctx *c, const uint8_t *rgb, int n, int quality, int16_t *y, size_t ystride, uint8_t *u, uint8_t *v,
                size_t cstride)
{
	static const int qtz[17] = {0, 15900, 16500, 17100, 18000, 18820, 19670, 20640, 21540, 23540, 25570, 27522, 27830, 27607, 28786, 31262, 32375};
	ColorParams p;
	p.yq = 1.0;
	p.qtz = 0;
	if (quality >= 20) p.mode = 0;
	else if (quality >= 18) { p.mode = 1; p.yq = (double)(quality == 19 ? 0.975f : 0.93f); }
	else if (quality == 17) p.mode = 2;
	else { p.mode = 3; p.qtz = qtz[quality < 0 ? 0 : quality]; }
	NHW_LAUNCH(c, k_colorspace, dim3(512 / CS_ROWS, n), 256, 0, rgb, y, u, v, ystride, cstride, p);
}

void pre_processing(nhw_ctx *c, int n, int quality, int16_t *y, size_t ystride)
{
	(void)quality;   // q17..q21 share one rule set; q>=22 never gets here; q<=16 is rejected upstream
	dim3 grid((510 + PRE_WARPS - 1) / PRE_WARPS, n);
	int16_t *energy = c->y_aux2 + NHW_GUARD_S, *kern = c->y_aux + NHW_GUARD_S;
	NHW_LAUNCH(c, k_pre_energy, grid, 32 * PRE_WARPS, 0, y, energy, c->rowmap, ystride, (size_t)NHW_Y_SLOT);
	NHW_LAUNCH(c, k_pre_chain, (n + 63) / 64, 64, 0, c->rowmap, c->rowcarry, n);
	NHW_LAUNCH(c, k_pre_apply, grid, 32 * PRE_WARPS, 0, energy, c->rowcarry, kern, (size_t)NHW_Y_SLOT);
	NHW_LAUNCH(c, k_pre_nudge, dim3(510, n), 256, 0, kern, y, (size_t)NHW_Y_SLOT, ystride);
}

static void dwt_attrs()
{
	static bool done = false;
	if (done) return;
	cudaFuncSetAttribute(k_dwt_level_smem<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (256 * 258 + 512) * 2);
	cudaFuncSetAttribute(k_dwt_level_smem<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 130 + 256) * 2);
	cudaFuncSetAttribute(k_dwt_cols_t<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 33 * 2);
	cudaFuncSetAttribute(k_dwt_cols_t<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 33 * 2);
	done = true;
}

// level 1 (512) + level 2 (256) of the luma plane; jpeg is consumed, proc/ll1 produced
void dwt_luma(nhw_ctx *c, int n, const int16_t *jpeg, size_t jstride, int16_t *proc, size_t pstride, int16_t *ll1,
              size_t lstride)
{
	dwt_attrs();
	int16_t *rows = c->y_aux + NHW_GUARD_S;
	NHW_LAUNCH(c, k_dwt_rows<512>, dim3(512 / 8, n), 256, 0, jpeg, rows, jstride, (size_t)NHW_Y_SLOT);
	NHW_LAUNCH(c, k_dwt_cols_t<512>, dim3(512 / 32, n), 256, 512 * 33 * 2, rows, proc, (size_t)NHW_Y_SLOT, pstride);
	NHW_LAUNCH(c, k_dwt_level_smem<256>, n, 256, (256 * 258 + 512) * 2, proc, proc, ll1, pstride, pstride, lstride, 512, 1);
}

void chroma_to_short(nhw_ctx *c, int n_planes, const uint8_t *u8, size_t in_stride, int16_t *jpeg, size_t out_stride)
{
	NHW_LAUNCH(c, k_u8_to_s16, dim3(NHW_CPLANE / 256, n_planes), 256, 0, u8, jpeg, in_stride, out_stride);
}

// level 1 (256) + level 2 (128) of n_planes chroma planes
void dwt_chroma(nhw_ctx *c, int n_planes, const int16_t *jpeg, size_t jstride, int16_t *proc, size_t pstride,
                int16_t *ll1, size_t lstride)
{
	dwt_attrs();
	int16_t *rows = c->c_aux + NHW_GUARD_S;
	NHW_LAUNCH(c, k_dwt_rows<256>, dim3(256 / 8, n_planes), 256, 0, jpeg, rows, jstride, (size_t)NHW_C_SLOT);
	NHW_LAUNCH(c, k_dwt_cols_t<256>, dim3(256 / 32, n_planes), 256, 256 * 33 * 2, rows, proc, (size_t)NHW_C_SLOT, pstride);
	NHW_LAUNCH(c, k_dwt_level_smem<128>, n_planes, 128, (128 * 130 + 256) * 2, proc, proc, ll1, pstride, pstride, lstride, 256, 1);
}

// one more analysis level on a band held in `im_jpeg` orientation (closed loop, encoder/nhw_encoder.c:281,2339)
void dwt_level_from_jpeg(nhw_ctx *c, int n_planes, const int16_t *jpeg, size_t jstride, int16_t *proc, size_t pstride,
                         int N, int row_stride)
{
	dwt_attrs();
	if (N == 256)
		NHW_LAUNCH(c, k_dwt_level_smem<256>, n_planes, 256, (256 * 258 + 512) * 2, jpeg, proc, (int16_t *)nullptr, jstride, pstride, (size_t)0, row_stride, 0);
	else
		NHW_LAUNCH(c, k_dwt_level_smem<128>, n_planes, 128, (128 * 130 + 256) * 2, jpeg, proc, (int16_t *)nullptr, jstride, pstride, (size_t)0, row_stride, 0);
}

}  // namespace nhw
