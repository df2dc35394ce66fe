// nhw_ctx.h -- host-side context of libnhw_cuda (internal; the public face is include/nhw_cuda.h)
#pragma once
#define NHW_LANES 4
#define NHW_DEC_BYTES_SLOT (1728 * 1024)   // >= decode.cu's DOFF_END (static_assert there)
#define NHW_MAX_SUB 16
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "enc_batch.cuh"

struct DecDesc;

// scheduling knobs, read from the environment ONCE in nhw_create (never on the call path)
struct NhwTuning {
	int lanes_device;      // NHW_LANES_DEVICE   sub-chunks of a device-resident encode call (1)
	int chroma_stream;     // NHW_CHROMA_STREAM  chroma chain on a side stream (1)
	int subs_encode, lanes_encode;   // NHW_SUBS_ENCODE (16), NHW_LANES_ENCODE (4): host-buffer encode waves
	int subs_decode, lanes_decode;   // NHW_SUBS_DECODE (8), NHW_LANES_DECODE (4)
	int dsf_streams;       // NHW_DSF_STREAMS    streams per warp in the decoder's serial front (0 = by batch size: 1, 2 or 4)
	int dsf_job_mask;      // NHW_DSF_JOBS       timing experiments only: which of the three serial jobs run (bit 0 luma, 1 chroma, 2 lists; 15 = all)
	int rows_grid_cap;     // SM count x NHW_ROWS_CTAS_PER_SM (24): grid cap of the "thread = row" kernels
	int plw_phases;        // NHW_PLW_PHASES     timing experiments only: which walks of the q <= 16 pre-sharpening run (15 = all)
	int fetch_kernel;      // NHW_FETCH_KERNEL   decode: stream bytes in pinned host memory are fetched by a kernel, not the copy engine (1)
};

struct nhw_ctx {
	int device;
	int max_batch;
	NhwTuning tune;
	const uint16_t *dec_lut;    // the decoder's prefix-code table in this device's memory (decode.cu)
	cudaStream_t stream;        // the stream kernels are issued on (lanes[0] for the context itself)
	// A batch is cut into up to NHW_LANES sub-chunks that run side by side, each on its own stream and on its
	// own slice of every workspace array (api.cu: lane_view): the many latency-bound stages of one sub-chunk
	// overlap the others', and host<->device copies overlap kernels.
	cudaStream_t lanes[NHW_LANES];
	cudaStream_t copy_stream;   // device->host copies of finished sub-chunks
	cudaStream_t up_stream;     // host->device copies of the pixels, queued back to back ahead of the sub-chunks that use them
	cudaEvent_t ev_fork, ev_join[NHW_LANES], ev_sub[NHW_MAX_SUB], ev_up[NHW_MAX_SUB];
	// the chroma chain of a chunk runs next to the (latency-bound) luma chain on this stream when the chunk is
	// encoded as one sub-chunk (device-resident input); NULL: everything on `stream`
	cudaStream_t chroma_stream;
	cudaEvent_t ev_chroma0, ev_chroma1;
	int chroma_side;            // 1: encode_chunk may use chroma_stream for this call
	uint64_t launches;
	char dbg_label[64];  // debug: stop issuing kernels after the dbg_count-th launch of this label
	int dbg_count, dbg_seen, dbg_stopped;
	int profile;         // per-kernel CUDA-event timing on/off (nhw_profile)
	void *prof;          // nhw::ProfState

	// ---- per-image workspace, each array sized for max_batch images (device memory).
	// Plane arrays are made of slots (plane + zero guard bands, see enc_img.cuh).
	uint8_t *rgb;        // staged input pixels (host API only), 786432 B / image
	int16_t *y_jpeg;     // luma `im_jpeg`                       NHW_Y_SLOT
	int16_t *y_proc;     // luma `im_process`
	int16_t *y_aux;      // scratch plane (pre-sharpen kernel values / transform row passes)
	int16_t *y_aux2;     // scratch plane (signed Laplacian energy)
	int16_t *y_ll1;      // `res256` (LL1 copy)                  NHW_C_SLOT
	int16_t *y_ll2save;  // `resIII` snapshot                    NHW_C_SLOT
	uint8_t *c_u8;       // 4:2:0 byte planes, [image][U,V]      2 * 65536 B
	int16_t *c_jpeg;     // chroma `im_jpeg`, [image][U,V]       2 * NHW_C_SLOT
	int16_t *c_proc;     // chroma `im_process`
	int16_t *c_aux;      // chroma scratch
	int16_t *c_ll1;      // chroma `res256`                      2 * NHW_Q_SLOT
	int16_t *c_ll2save;  // chroma `resIII`                      2 * NHW_Q_SLOT
	uint32_t *rowmap;    // pre-sharpen carry maps               512 u32
	uint8_t *rowcarry;   // pre-sharpen carry-in class per row   512 u8
	uint8_t *enc_bytes;  // byte-sized encoder state             ENC_BYTES_SLOT
	uint8_t *dec_bytes;  // byte-sized decoder state             NHW_DEC_BYTES_SLOT (its own array: the encoder relies on the
	                     // never-written gaps of enc_bytes reading as zero, so the decoder must not scribble there)
	EncHdr *enc_hdr;

	uint8_t *out_dev;    // staged output streams (host API only), NHW_MAX_STREAM_BYTES / image
	uint8_t *pack_dev;   // the chunk's streams packed back to back
	uint32_t *len_dev;
	int32_t *status_dev;
	uint64_t *offs_dev;  // max_batch + 1
	uint64_t *offs_host; // pinned
	uint8_t *dec_yuv;    // decoder: Y,U,V byte planes 512x512, 786432 B / image
	void *dec_desc_dev;  // decoder: DecDesc per image (device) and its pinned host staging
	void *dec_desc_host;
	int32_t *status_host;
};

namespace nhw {

// front.cu
void colorspace(nhw_ctx *c, const uint8_t *rgb, int n, int quality, int16_t *y, size_t ystride, uint8_t *u, uint8_t *v,
                size_t cstride);
void pre_processing(nhw_ctx *c, int n, int quality, int16_t *y, size_t ystride);
void pre_processing_lowq(nhw_ctx *c, int n, int quality, int16_t *y, size_t ystride);   // y must not be y_proc / y_aux / y_aux2
void chroma_pre_uv(nhw_ctx *c, int n_planes, int quality, const uint8_t *uv, int16_t *out, size_t oslot);
void chroma_thresholds(nhw_ctx *c, int n_planes, int16_t *proc, size_t pslot, int ratio);
// front_fused.cu
void front_fused(nhw_ctx *c, const uint8_t *rgb, int n, int quality, int16_t *y_proc, size_t ypstride, int16_t *y_ll1,
                 size_t ylstride, uint8_t *uv_bytes, int16_t *c_proc, size_t cpstride, int16_t *c_ll1, size_t clstride,
                 int16_t *kept, size_t kstride);   // kept: q22/q23 only, 256x512 per image (may be NULL below q22)
long color_fast_path_mismatches(nhw_ctx *c);
long dec_color_fast_path_mismatches(nhw_ctx *c);   // decode.cu
void dwt_level_from_jpeg(nhw_ctx *c, int n_planes, const int16_t *jpeg, size_t jstride, int16_t *proc, size_t pstride,
                         int N, int row_stride, int16_t *snap, size_t snap_slot);

// encode.cu
void encode_chunk(nhw_ctx *c, const uint8_t *rgb, int n, int q, uint8_t *out_dev, uint32_t *len_dev, int32_t *status_dev);
void pack_streams(nhw_ctx *c, int n);   // out_dev slots + len_dev -> offs_dev, pack_dev
void pack_streams_to(nhw_ctx *c, const uint8_t *slots, const uint32_t *len, int n, uint64_t *offs, uint8_t *dense);

// decode.cu
void decode_chunk(nhw_ctx *c, const uint8_t *blobs, const uint64_t *offs, const struct DecDesc *desc, int32_t *status, int n,
                  uint8_t *rgb_dev, bool any_lowq, bool any_hq, bool want_yuv);
void decode_chunk_device(nhw_ctx *c, const uint8_t *in, size_t stride, const uint32_t *len, const uint64_t *offs, int n,
                         uint8_t *rgb_dev, int32_t *status_dev);

// synth.cu
void digest(nhw_ctx *c, const uint8_t *data, size_t stride, const uint32_t *len, uint32_t fixed_len, int n, uint64_t *out);
void synth(nhw_ctx *c, uint8_t *rgb, int n, uint32_t seed0, int kind, const int16_t *sin_lut);

// per-device set-up, called by nhw_create with the device current: opt-in to > 48 KB of dynamic shared memory for the
// kernels that need it and upload of the constant tables.  Both are per-device state in CUDA; nothing here is
// guarded by process-wide flags, so contexts on different GPUs (and from different threads) are independent.
bool front_device_init(nhw_ctx *c);
bool encode_device_init(nhw_ctx *c);
bool decode_device_init(nhw_ctx *c);

// api.cu: per-kernel timing with CUDA events recorded on c->stream around each launch
bool dbg_skip(nhw_ctx *c, const char *label);
void prof_begin(nhw_ctx *c, const char *label);
void prof_end(nhw_ctx *c);

bool check(cudaError_t e, const char *what);
void set_error(const char *fmt, ...);

}  // namespace nhw
