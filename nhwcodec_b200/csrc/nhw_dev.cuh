// nhw_dev.cuh -- shared device-side definitions for libnhw_cuda (sm_100a only).
//
// Plane vocabulary follows the reference (encoder/codec.h:112-123): a luma working plane is
// 512x512 int16 (row stride 512), a chroma working plane 256x256 int16 (row stride 256).
// "proc" = im_process, "jpeg" = im_jpeg.  All planes of a batch are stored image-major.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NHW_YW 512                 // luma width/height  (2*IM_DIM, encoder/codec.h:61)
#define NHW_CW 256                 // chroma width/height (IM_DIM)
#define NHW_YPLANE (512 * 512)     // elements in a luma plane   (4*IM_SIZE)
#define NHW_CPLANE (256 * 256)     // elements in a chroma plane (IM_SIZE)
#define NHW_RGB_BYTES (512 * 512 * 3)

struct nhw_ctx;

// Launch bookkeeping: every kernel launch goes through this so gpu_launches is a real count.
#define NHW_LAUNCH_L(ctx, label, kernel, grid, block, smem, ...)                          \
	do {                                                                                  \
		if (nhw::dbg_skip((ctx), (label))) break;                                         \
		nhw::prof_begin((ctx), (label));                                                  \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                  \
		(ctx)->launches++;                                                                \
		nhw::prof_end((ctx));                                                             \
	} while (0)
#define NHW_LAUNCH(ctx, kernel, grid, block, smem, ...) NHW_LAUNCH_L(ctx, #kernel, kernel, grid, block, smem, __VA_ARGS__)


// 16-bit add on a cell other threads may be adding to as well (CAS on the containing word)
__device__ __forceinline__ void atomic_add_s16(int16_t *p, int v)
{
	uint32_t *w = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
	const bool hi = (reinterpret_cast<uintptr_t>(p) & 2) != 0;
	uint32_t old = *w, assumed;
	do {
		assumed = old;
		const uint32_t cur = hi ? assumed >> 16 : assumed & 0xffffu;
		const uint32_t nv = (cur + (uint32_t)v) & 0xffffu;
		old = atomicCAS(w, assumed, hi ? (assumed & 0xffffu) | (nv << 16) : (assumed & 0xffff0000u) | nv);
	} while (old != assumed);
}
