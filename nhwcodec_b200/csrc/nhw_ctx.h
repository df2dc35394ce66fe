// nhw_ctx.h -- host-side context of libnhw_cuda (internal; the public face is include/nhw_cuda.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

struct nhw_ctx {
	int device;
	int max_batch;
	cudaStream_t stream;
	uint64_t launches;

	// ---- per-image workspace, each sized for max_batch images (device memory) ----
	uint8_t *rgb;        // staged input pixels (host API only)          786432 B / image
	int16_t *y_jpeg;     // luma `im_jpeg`                               512*512 int16
	int16_t *y_proc;     // luma `im_process`
	int16_t *y_aux;      // scratch plane (pre-sharpen kernel values / DWT row-pass output)
	int16_t *y_aux2;     // scratch plane (signed Laplacian energy)
	int16_t *y_ll1;      // `res256` (LL1 copy)                          256*256 int16
	int16_t *y_ll2save;  // `resIII` snapshot                            256*256 int16
	uint8_t *c_u8;       // U then V byte planes                         2 * 256*256 u8
	int16_t *c_jpeg;     // chroma `im_jpeg`, U then V                   2 * 256*256 int16
	int16_t *c_proc;     // chroma `im_process`
	int16_t *c_aux;      // chroma scratch
	int16_t *c_ll1;      // chroma `res256`                              2 * 128*128 int16
	uint32_t *rowmap;    // pre-sharpen carry maps                       512 u32
	uint8_t *rowcarry;   // pre-sharpen carry-in class per row           512 u8

	uint8_t *out_dev;    // staged output streams (host API only)
	uint32_t *len_dev;
	int32_t *status_dev;
};

namespace nhw {

// front.cu
void colorspace(nhw_ctx *c, const uint8_t *rgb, int n, int quality, int16_t *y, uint8_t *u, uint8_t *v);
void pre_processing(nhw_ctx *c, int n, int quality, int16_t *y);
void dwt_luma(nhw_ctx *c, int n, int16_t *y_jpeg, int16_t *y_proc, int16_t *y_ll1);
void chroma_to_short(nhw_ctx *c, int n, const uint8_t *u8, int16_t *c_jpeg);
void dwt_chroma(nhw_ctx *c, int n, int16_t *c_jpeg, int16_t *c_proc, int16_t *c_ll1);
// single-level transforms on planes of arbitrary stride, used by the closed loop
void dwt_level2_from_jpeg(nhw_ctx *c, int n_planes, const int16_t *jpeg, int16_t *proc, int N, int stride);

// synth.cu
void synth(nhw_ctx *c, uint8_t *rgb, int n, uint32_t seed0, int kind, const int16_t *sin_lut);

bool check(cudaError_t e, const char *what);
void set_error(const char *fmt, ...);

}  // namespace nhw
