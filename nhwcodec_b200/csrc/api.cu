// api.cu -- context management and the C-ABI entry points of libnhw_cuda.so
// (declared in include/nhw_cuda.h).  There is no CPU fallback anywhere in this library:
// without a usable CUDA device nhw_create() fails and every other call needs a context.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <new>

#include "../../include/nhw_cuda.h"
#include "nhw_ctx.h"
#include "nhw_dev.cuh"

namespace nhw {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

bool check(cudaError_t e, const char *what)
{
	if (e == cudaSuccess) return true;
	set_error("%s: %s", what, cudaGetErrorString(e));
	return false;
}

}  // namespace nhw

using nhw::check;

template <typename T>
static bool dev_alloc(T **p, size_t count)
{
	return check(cudaMalloc((void **)p, count * sizeof(T)), "cudaMalloc");
}

extern "C" {

const char *nhw_last_error(void) { return nhw::g_err; }
int nhw_version(void) { return 100; }

int nhw_create(int device, int max_batch, nhw_ctx **out)
{
	if (!out || max_batch <= 0) return NHW_ERR_ARG;
	*out = nullptr;
	int count = 0;
	if (!check(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count <= 0) {
		if (count <= 0 && nhw::g_err[0] == 0) nhw::set_error("no CUDA device");
		return NHW_ERR_CUDA;
	}
	if (device < 0 || device >= count) { nhw::set_error("device %d out of range (%d devices)", device, count); return NHW_ERR_ARG; }
	if (!check(cudaSetDevice(device), "cudaSetDevice")) return NHW_ERR_CUDA;
	nhw_ctx *c = new (std::nothrow) nhw_ctx();
	if (!c) return NHW_ERR_NOMEM;
	memset(c, 0, sizeof *c);
	c->device = device;
	c->max_batch = max_batch;
	const size_t B = (size_t)max_batch;
	bool ok = check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
	ok = ok && dev_alloc(&c->rgb, B * NHW_RGB_BYTES);
	ok = ok && dev_alloc(&c->y_jpeg, B * NHW_YPLANE) && dev_alloc(&c->y_proc, B * NHW_YPLANE);
	ok = ok && dev_alloc(&c->y_aux, B * NHW_YPLANE) && dev_alloc(&c->y_aux2, B * NHW_YPLANE);
	ok = ok && dev_alloc(&c->y_ll1, B * NHW_CPLANE) && dev_alloc(&c->y_ll2save, B * NHW_CPLANE);
	ok = ok && dev_alloc(&c->c_u8, B * 2 * NHW_CPLANE);
	ok = ok && dev_alloc(&c->c_jpeg, B * 2 * NHW_CPLANE) && dev_alloc(&c->c_proc, B * 2 * NHW_CPLANE);
	ok = ok && dev_alloc(&c->c_aux, B * 2 * NHW_CPLANE) && dev_alloc(&c->c_ll1, B * 2 * 128 * 128);
	ok = ok && dev_alloc(&c->rowmap, B * 512) && dev_alloc(&c->rowcarry, B * 512);
	ok = ok && dev_alloc(&c->len_dev, B) && dev_alloc(&c->status_dev, B);
	if (!ok) { nhw_destroy(c); return NHW_ERR_CUDA; }
	// zero once: halos / never-written borders read as 0, matching the canonical oracle
	cudaMemsetAsync(c->y_aux, 0, B * NHW_YPLANE * 2, c->stream);
	cudaMemsetAsync(c->y_aux2, 0, B * NHW_YPLANE * 2, c->stream);
	cudaMemsetAsync(c->rowmap, 0, B * 512 * 4, c->stream);
	cudaMemsetAsync(c->rowcarry, 0, B * 512, c->stream);
	if (!check(cudaStreamSynchronize(c->stream), "nhw_create sync")) { nhw_destroy(c); return NHW_ERR_CUDA; }
	*out = c;
	return NHW_OK;
}

void nhw_destroy(nhw_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	void *ptrs[] = {c->rgb, c->y_jpeg, c->y_proc, c->y_aux, c->y_aux2, c->y_ll1, c->y_ll2save, c->c_u8, c->c_jpeg,
	                c->c_proc, c->c_aux, c->c_ll1, c->rowmap, c->rowcarry, c->out_dev, c->len_dev, c->status_dev};
	for (void *p : ptrs)
		if (p) cudaFree(p);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

uint64_t nhw_launch_count(const nhw_ctx *c) { return c ? c->launches : 0; }

static int finish(nhw_ctx *c, const char *what)
{
	if (!check(cudaGetLastError(), what)) return NHW_ERR_CUDA;
	if (!check(cudaStreamSynchronize(c->stream), what)) return NHW_ERR_CUDA;
	return NHW_OK;
}

static bool quality_supported_frontend(int q) { return q >= 17 && q <= 23; }

int nhw_stage_colorspace_device(nhw_ctx *c, const uint8_t *rgb_dev, int n, int quality, int pre,
                                int16_t *y, uint8_t *u, uint8_t *v)
{
	if (!c || !rgb_dev || n <= 0) return NHW_ERR_ARG;
	if (quality < 1 || quality > 23) return NHW_ERR_QUALITY;
	if (pre && !quality_supported_frontend(quality)) return NHW_ERR_QUALITY;
	cudaSetDevice(c->device);
	for (int i0 = 0; i0 < n; i0 += c->max_batch) {
		int m = n - i0 < c->max_batch ? n - i0 : c->max_batch;
		int16_t *yy = y ? y + (size_t)i0 * NHW_YPLANE : c->y_jpeg;
		nhw::colorspace(c, rgb_dev + (size_t)i0 * NHW_RGB_BYTES, m, quality, yy,
		                u ? u + (size_t)i0 * NHW_CPLANE : nullptr, v ? v + (size_t)i0 * NHW_CPLANE : nullptr);
		if (pre && quality < 22) nhw::pre_processing(c, m, quality, yy);
	}
	return finish(c, "nhw_stage_colorspace_device");
}

int nhw_stage_frontend_device(nhw_ctx *c, const uint8_t *rgb_dev, int n, int quality,
                              int16_t *y_proc, int16_t *y_ll1, int16_t *c_proc, int16_t *c_ll1)
{
	if (!c || !rgb_dev || n <= 0) return NHW_ERR_ARG;
	if (!quality_supported_frontend(quality)) return NHW_ERR_QUALITY;
	cudaSetDevice(c->device);
	for (int i0 = 0; i0 < n; i0 += c->max_batch) {
		int m = n - i0 < c->max_batch ? n - i0 : c->max_batch;
		uint8_t *u8 = c->c_u8, *v8 = c->c_u8 + (size_t)m * NHW_CPLANE;
		nhw::colorspace(c, rgb_dev + (size_t)i0 * NHW_RGB_BYTES, m, quality, c->y_jpeg, u8, v8);
		if (quality < 22) nhw::pre_processing(c, m, quality, c->y_jpeg);
		int16_t *yp = y_proc ? y_proc + (size_t)i0 * NHW_YPLANE : c->y_proc;
		int16_t *yl = y_ll1 ? y_ll1 + (size_t)i0 * NHW_CPLANE : c->y_ll1;
		nhw::dwt_luma(c, m, c->y_jpeg, yp, yl);
		// chroma planes are laid out [U images..., V images...] inside a chunk
		nhw::chroma_to_short(c, m, c->c_u8, c->c_jpeg);
		nhw::dwt_chroma(c, m, c->c_jpeg, c->c_proc, c->c_ll1);
		if (c_proc) {
			// hand back as [image][U,V]
			for (int k = 0; k < 2; k++)
				cudaMemcpy2DAsync(c_proc + ((size_t)i0 * 2 + k) * NHW_CPLANE, 2 * NHW_CPLANE * 2,
				                  c->c_proc + (size_t)k * m * NHW_CPLANE, NHW_CPLANE * 2, NHW_CPLANE * 2, m,
				                  cudaMemcpyDeviceToDevice, c->stream);
		}
		if (c_ll1) {
			for (int k = 0; k < 2; k++)
				cudaMemcpy2DAsync(c_ll1 + ((size_t)i0 * 2 + k) * 16384, 2 * 16384 * 2,
				                  c->c_ll1 + (size_t)k * m * 16384, 16384 * 2, 16384 * 2, m,
				                  cudaMemcpyDeviceToDevice, c->stream);
		}
	}
	return finish(c, "nhw_stage_frontend_device");
}

int nhw_synth_batch_device(nhw_ctx *c, uint8_t *rgb_dev, int n, uint32_t seed0, int kind, const int16_t *sin_lut_dev)
{
	if (!c || !rgb_dev || n <= 0 || (kind != 1 && !sin_lut_dev)) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	nhw::synth(c, rgb_dev, n, seed0, kind, sin_lut_dev);
	return finish(c, "nhw_synth_batch_device");
}

int nhw_encode_batch_device(nhw_ctx *c, const uint8_t *rgb_dev, int n, int quality,
                            uint8_t *out_dev, uint32_t *len_dev, int32_t *status_dev)
{
	(void)c; (void)rgb_dev; (void)n; (void)quality; (void)out_dev; (void)len_dev; (void)status_dev;
	nhw::set_error("nhw_encode_batch_device: full encode path not built yet");
	return NHW_ERR_QUALITY;
}

int nhw_encode_batch(nhw_ctx *c, const uint8_t *rgb, int n, int quality,
                     uint8_t *out, size_t out_cap, uint64_t *offsets, int32_t *status)
{
	(void)c; (void)rgb; (void)n; (void)quality; (void)out; (void)out_cap; (void)offsets; (void)status;
	nhw::set_error("nhw_encode_batch: full encode path not built yet");
	return NHW_ERR_QUALITY;
}

int nhw_decode_batch(nhw_ctx *c, const uint8_t *in, const uint64_t *offsets, int n, uint8_t *rgb, int32_t *status)
{
	(void)c; (void)in; (void)offsets; (void)n; (void)rgb; (void)status;
	nhw::set_error("nhw_decode_batch: decode path not built yet");
	return NHW_ERR_QUALITY;
}

}  // extern "C"
