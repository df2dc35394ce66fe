// synth.cu -- deterministic synthetic test images generated on the device (bench/test tooling,
// SURVEY.md section 8d).  Integer-only so that nhwcodec_b200/synth.py reproduces every byte
// on the host with numpy.  "natural-like" = 12 sinusoids + 10 rectangles + grain, three
// correlated channels; "noise" = i.i.d. uniform bytes;
// kind 2 = natural-like with 4x grain ("textured").
#include "nhw_ctx.h"
#include "nhw_dev.cuh"

namespace {

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h)
{
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}
__device__ __forceinline__ uint32_t prm(uint32_t seed, uint32_t k)
{
	return fmix32(seed * 0x9E3779B1u + k * 0x632BE5ABu + 0x7F4A7C15u);
}

struct Wave { int fx, fy, ph, amp, gain[3]; };
struct Rect { int x0, y0, x1, y1, d[3]; };

__global__ void __launch_bounds__(256) k_synth(uint8_t *__restrict__ rgb, uint32_t seed0, int kind,
                                               const int16_t *__restrict__ lut)
{
	__shared__ Wave wv[12];
	__shared__ Rect rc[10];
	__shared__ int16_t slut[1024];
	const int img = blockIdx.y;
	const uint32_t seed = seed0 + (uint32_t)img;
	if (kind != 1) {
		for (int i = threadIdx.x; i < 1024; i += 256) slut[i] = lut[i];
		if (threadIdx.x < 12) {
			uint32_t t = threadIdx.x;
			uint32_t p0 = prm(seed, 4 * t), p1 = prm(seed, 4 * t + 1), p2 = prm(seed, 4 * t + 2), p3 = prm(seed, 4 * t + 3);
			Wave w;
			w.fx = (int)(p0 % 25u) - 12;
			w.fy = (int)(p1 % 25u) - 12;
			w.ph = (int)(p2 & 1023u);
			w.amp = 4 + (int)(p3 % 24u);
			for (int ch = 0; ch < 3; ch++) w.gain[ch] = 128 + (int)((p3 >> (8 + 8 * ch)) & 127u);
			wv[t] = w;
		} else if (threadIdx.x >= 32 && threadIdx.x < 42) {
			uint32_t t = threadIdx.x - 32;
			uint32_t q0 = prm(seed, 100 + 5 * t), q1 = prm(seed, 101 + 5 * t), q2 = prm(seed, 102 + 5 * t),
			         q3 = prm(seed, 103 + 5 * t), q4 = prm(seed, 104 + 5 * t);
			Rect r;
			r.x0 = (int)(q0 & 511u);
			r.y0 = (int)(q1 & 511u);
			r.x1 = r.x0 + 16 + (int)(q2 % 200u);
			r.y1 = r.y0 + 16 + (int)(q3 % 200u);
			for (int ch = 0; ch < 3; ch++) r.d[ch] = (int)((q4 >> (8 * ch)) & 127u) - 64;
			rc[t] = r;
		}
		__syncthreads();
	}
	uint8_t *dst = rgb + (size_t)img * NHW_RGB_BYTES;
	for (int pix = blockIdx.x * 256 + threadIdx.x; pix < NHW_YPLANE; pix += gridDim.x * 256) {
		const int x = pix & 511, y = pix >> 9;
		for (int ch = 0; ch < 3; ch++) {
			uint32_t g = fmix32((seed * 0x9E3779B1u) ^ (((uint32_t)pix * 3u + (uint32_t)ch) * 0x85EBCA77u + 0x1B873593u));
			int val;
			if (kind == 1) {
				val = (int)(g & 255u);
			} else {
				int acc = 128 << 8;
				for (int t = 0; t < 12; t++) {
					int phase = ((wv[t].fx * x + wv[t].fy * y) * 2 + wv[t].ph) & 1023;
					acc += (wv[t].amp * wv[t].gain[ch] * (int)slut[phase]) >> 8;
				}
				for (int t = 0; t < 10; t++)
					if (x >= rc[t].x0 && x < rc[t].x1 && y >= rc[t].y0 && y < rc[t].y1) acc += rc[t].d[ch] << 8;
				int s = 0;
				for (int b = 0; b < 8; b++) s += (int)((g >> (2 * b)) & 3u);
				acc += ((s - 12) * (kind == 2 ? 4 : 1)) << 8;
				val = acc >> 8;
				val = val < 0 ? 0 : (val > 255 ? 255 : val);
			}
			dst[(size_t)pix * 3 + ch] = (uint8_t)val;
		}
	}
}

// ---- per-item digest of a batch of byte strings (decoded images, .nhw streams): a position-weighted 64-bit sum,
// sum over k of (byte[k] + 1) * w(k) mod 2^64 with w(k) = odd multiplier of (k + 1), so that any order of summation
// gives the same value and the whole CTA can work on one item.  Used to compare runs (1 GPU vs N GPUs) without
// moving the data; not a cryptographic hash.
__global__ void __launch_bounds__(256) k_digest(const uint8_t *__restrict__ data, size_t stride, const uint32_t *__restrict__ len,
                                                uint32_t fixed_len, unsigned long long *__restrict__ out)
{
	__shared__ unsigned long long part[256];
	const int i = blockIdx.x;
	const uint8_t *p = data + (size_t)i * stride;
	const uint32_t L = len ? len[i] : fixed_len;
	unsigned long long acc = 0;
	const uint32_t words = (((uintptr_t)p & 3) == 0) ? L >> 2 : 0;   // aligned body as 32-bit words, the rest bytewise
	for (uint32_t w = threadIdx.x; w < words; w += 256) {
		const uint32_t v = reinterpret_cast<const uint32_t *>(p)[w];
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const unsigned long long k = 4ull * w + b + 1ull;
			acc += (unsigned long long)(((v >> (8 * b)) & 255u) + 1u) * (k * 0x9E3779B97F4A7C15ull | 1ull);
		}
	}
	for (uint32_t k = 4 * words + threadIdx.x; k < L; k += 256) acc += (unsigned long long)(p[k] + 1u) * (((unsigned long long)k + 1ull) * 0x9E3779B97F4A7C15ull | 1ull);
	part[threadIdx.x] = acc;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[i] = part[0] ^ ((unsigned long long)L << 40);
}

}  // namespace

namespace nhw {
void digest(nhw_ctx *c, const uint8_t *data, size_t stride, const uint32_t *len, uint32_t fixed_len, int n, uint64_t *out)
{
	NHW_LAUNCH(c, k_digest, n, 256, 0, data, stride, len, fixed_len, reinterpret_cast<unsigned long long *>(out));
}

void synth(nhw_ctx *c, uint8_t *rgb, int n, uint32_t seed0, int kind, const int16_t *sin_lut)
{
	NHW_LAUNCH(c, k_synth, dim3(64, n), 256, 0, rgb, seed0, kind, sin_lut);
}
}  // namespace nhw
