/*
 * include/nhw_cuda.h -- C-ABI of libnhw_cuda.so, the B200 (sm_100a) implementation of the
 * NHW codec hot path.  Plain C, plain pointers and sizes; no torch / C++ types.
 *
 * The reference (rcanut/nhwcodec) has no plugin/FFI interface; its boundary is the set of
 * extern functions its two CLI files call:
 *     encoder/codec.h:184-189   encode_image, read_image_bmp, write_compressed_file,
 *                               downsample_YUV420
 *     decoder/codec.h:198-202   decode_image, parse_file
 *     decoder/nhw_decoder_cli.c:58-59  setup_bmp_header, write_image_bmp
 * Those per-image symbols are provided, with the reference's exact signatures, by
 * libnhw_compat (include/nhw_compat.h) on top of the batch entry points below, so the
 * reference's own nhw_encoder_cli.c / nhw_decoder_cli.c link against this library unchanged.
 *
 * Batch entry points replace the reference's one-process-per-image model: every image is an
 * independent unit (all reference state lives in per-call structs, encoder/codec.h:112-181),
 * errors are per-image status codes instead of exit() (encoder/compress_pixel.c:234,270-271).
 */
#ifndef NHW_CUDA_H
#define NHW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NHW_IMG_W 512
#define NHW_IMG_H 512
#define NHW_PIX_BYTES (512 * 512 * 3)      /* raw BMP pixel bytes per image (encoder/nhw_encoder.c:3086) */
#define NHW_MAX_STREAM_BYTES (1u << 19)    /* per-image upper bound on a .nhw stream: 512 KiB */

/* status / error codes (per image in status[], or as function return value) */
#define NHW_OK 0
#define NHW_ERR_CUDA (-1)                  /* CUDA runtime failure; see nhw_last_error() */
#define NHW_ERR_ARG (-2)
#define NHW_ERR_QUALITY (-3)               /* quality setting not supported by this build */
#define NHW_ERR_CODEBOOK (-4)              /* reference would exit(-1): encoder/compress_pixel.c:234,270-271 */
#define NHW_ERR_OVERFLOW (-5)              /* an internal list outgrew its slot (reference would write out of bounds) */
#define NHW_ERR_NOMEM (-6)
#define NHW_ERR_STREAM (-7)                /* malformed .nhw (decoder/nhw_decoder.c:1497-1500) */

typedef struct nhw_ctx nhw_ctx;

/* Create a codec context on CUDA device `device` with workspace for up to `max_batch`
 * images in flight (larger batches are processed in chunks of max_batch).  Fails loudly
 * (NHW_ERR_CUDA) when there is no usable GPU: there is no CPU fallback. */
int nhw_create(int device, int max_batch, nhw_ctx **out);
void nhw_destroy(nhw_ctx *ctx);
const char *nhw_last_error(void);
int nhw_version(void);

/* ---- encode: replaces read_image_bmp's pixel path + downsample_YUV420 + encode_image +
 * write_compressed_file (encoder/nhw_encoder_cli.c:179-183) for n images at once. ----
 * rgb      : n * NHW_PIX_BYTES bytes, each image = the 786432 bytes that follow the BMP
 *            header, in file order (host memory; pinned memory is faster).
 * quality  : the reference's -q value (encoder/nhw_encoder_cli.c:118).
 * out      : receives the .nhw streams back to back; offsets[i]..offsets[i+1] is image i.
 * out_cap  : capacity of out in bytes.
 * offsets  : n+1 entries.    status : n entries (NHW_OK or an error code).
 * Returns NHW_OK if the call ran (inspect status[] per image), else a negative error. */
int nhw_encode_batch(nhw_ctx *ctx, const uint8_t *rgb, int n, int quality,
                     uint8_t *out, size_t out_cap, uint64_t *offsets, int32_t *status);

/* Same with every buffer resident in device memory of the context's GPU (no host copies).
 * out_dev holds n slots of NHW_MAX_STREAM_BYTES; len_dev[i] receives the stream length. */
int nhw_encode_batch_device(nhw_ctx *ctx, const uint8_t *rgb_dev, int n, int quality,
                            uint8_t *out_dev, uint32_t *len_dev, int32_t *status_dev);

/* Pack the n stream slots nhw_encode_batch_device wrote (stride NHW_MAX_STREAM_BYTES) back to back into dense_dev:
 * offs_dev receives n + 1 offsets (offs_dev[n] = total bytes; dense_dev must hold that many -- n * NHW_MAX_STREAM_BYTES
 * always suffices).  This is the local half of the one exchange step of a multi-GPU encode (the stream gather). */
int nhw_pack_batch_device(nhw_ctx *ctx, const uint8_t *slots_dev, const uint32_t *len_dev, int n, uint64_t *offs_dev,
                          uint8_t *dense_dev);

/* A 64-bit position-weighted checksum per item of a batch of byte strings in device memory (item i = data_dev +
 * i * stride, length len_dev[i], or fixed_len when len_dev is NULL).  Lets sharded results (decoded pixels stay on
 * their GPU) be compared between runs without moving them. */
int nhw_digest_batch_device(nhw_ctx *ctx, const uint8_t *data_dev, size_t stride, const uint32_t *len_dev, uint32_t fixed_len,
                            int n, uint64_t *digest_dev);

/* ---- decode: replaces decode_image + write_image_bmp's pixel path
 * (decoder/nhw_decoder_cli.c:83-90) for n streams. ----
 * in/offsets : concatenated .nhw streams (host).   rgb : n * NHW_PIX_BYTES bytes out. */
int nhw_decode_batch(nhw_ctx *ctx, const uint8_t *in, const uint64_t *offsets, int n,
                     uint8_t *rgb, int32_t *status);

/* Same with every buffer resident in device memory (no host copies, no host synchronisation before the kernels:
 * the .nhw headers are walked on the device).  Stream i occupies in_dev[i * stride, i * stride + len_dev[i]) -- with
 * stride = NHW_MAX_STREAM_BYTES this is exactly what nhw_encode_batch_device leaves behind.  64 bytes past the end of
 * the last stream must be readable (the bit reader looks ahead).  rgb_dev: n * NHW_PIX_BYTES.  status_dev: n entries. */
int nhw_decode_batch_device(nhw_ctx *ctx, const uint8_t *in_dev, size_t stride, const uint32_t *len_dev, int n,
                            uint8_t *rgb_dev, int32_t *status_dev);
/* ... and for streams packed back to back: stream i is in_dev[offs_dev[i], offs_dev[i + 1]) (n + 1 offsets). */
int nhw_decode_batch_packed_device(nhw_ctx *ctx, const uint8_t *in_dev, const uint64_t *offs_dev, int n,
                                   uint8_t *rgb_dev, int32_t *status_dev);

/* Same, but stops before the colour conversion: yuv receives, per image, the three 512x512
 * byte planes Y, U, V (786432 bytes) that decode_image leaves in im_bufferY/U/V for
 * write_image_bmp (decoder/nhw_decoder.c:877-891,1137-1181).  quality (n entries, may be NULL)
 * receives each stream's quality byte. */
int nhw_decode_batch_planes(nhw_ctx *ctx, const uint8_t *in, const uint64_t *offsets, int n,
                            uint8_t *yuv, int32_t *quality, int32_t *status);

/* ---- stage-level entry points (device pointers), used by the parity tests and ncu runs.
 * Layouts follow the reference's working planes (encoder/codec.h:112-123):
 *   y_proc : n * 512*512 int16, the luma coefficient plane (`im_process`) after both DWT
 *            levels, transposed orientation;
 *   y_ll1  : n * 256*256 int16, the LL1 copy (`res256`, encoder/nhw_encoder.c:127-135);
 *   c_proc : n * 2 * 256*256 int16, U then V coefficient planes;
 *   c_ll1  : n * 2 * 128*128 int16, chroma LL1 copies (encoder/nhw_encoder.c:2270-2276).
 * Any output pointer may be NULL. */
int nhw_stage_frontend_device(nhw_ctx *ctx, const uint8_t *rgb_dev, int n, int quality,
                              int16_t *y_proc, int16_t *y_ll1, int16_t *c_proc, int16_t *c_ll1);

/* colour stage only: y (n*512*512 int16), u,v (n*256*256 u8 each) -- downsample_YUV420,
 * encoder/colorspace.c:55.  pre != 0 also applies pre_processing (encoder/image_processing.c:558). */
int nhw_stage_colorspace_device(nhw_ctx *ctx, const uint8_t *rgb_dev, int n, int quality, int pre,
                                int16_t *y, uint8_t *u, uint8_t *v);

/* Deterministic synthetic "natural-like" test images, generated on the device
 * (SURVEY.md section 8d): image i of the call uses seed seed0+i.  kind 0 = natural-like,
 * 1 = uniform noise.  sin_lut = 1024 int16 (device), see nhwcodec_b200/synth.py. */
int nhw_synth_batch_device(nhw_ctx *ctx, uint8_t *rgb_dev, int n, uint32_t seed0, int kind,
                           const int16_t *sin_lut_dev);

/* Page-locked ("pinned") host memory for the buffers handed to the host-buffer entry points: copies from and to pinned
 * memory run at full PCIe speed and asynchronously; with pageable memory the calls still work, slower.  NULL on failure. */
void *nhw_host_alloc(size_t bytes);
void nhw_host_free(void *p);

/* number of kernel launches issued by this context since creation (bench.py gpu_launches) */
uint64_t nhw_launch_count(const nhw_ctx *ctx);

/* the CUDA stream (cudaStream_t) every kernel of this context is launched on, so callers can
 * bracket calls with their own CUDA events */
void *nhw_stream(const nhw_ctx *ctx);

/* Per-kernel timing with CUDA events recorded on that stream around each launch.
 * enable: 0 = off, 1 = on, 2 = on and reset the accumulated table.
 * nhw_profile_read writes one line per kernel label: "label\tmilliseconds\tlaunches\n"
 * (accumulated since the last reset) and returns the number of bytes written. */
int nhw_profile(nhw_ctx *ctx, int enable);
long nhw_profile_read(nhw_ctx *ctx, char *buf, size_t cap);

/* Debug aids for stage-level parity bisection (tests only): stop issuing kernels after the
 * `occurrence`-th launch of the kernel labelled `label` in the next encode call (NULL = off),
 * and read a workspace array of image `img` of the last chunk back to the host.
 * what: proc jpeg aux ll1 ll2s cproc_u cproc_v cjpeg_u cjpeg_v scan tree1 llcode hdr */
int nhw_debug_stop_after(nhw_ctx *ctx, const char *label, int occurrence);
int nhw_debug_read(nhw_ctx *ctx, const char *what, int img, void *host, size_t bytes);
/* Runs all 2^24 RGB triples through the integer form of the q>=20 colour transform used by the
 * fused front end and through the IEEE double/float form of encoder/colorspace.c:71-99;
 * returns the number of triples on which they differ (must be 0), or a negative error. */
long nhw_debug_color_check(nhw_ctx *ctx);
/* The same for the decoder's q >= 20 YCbCr -> RGB matrix (decoder/nhw_decoder_cli.c:148-150): all 2^24 (Y, U, V) triples
 * through the integer form of the back-end kernel and through the IEEE double form; returns the number that differ. */
long nhw_debug_dec_color_check(nhw_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NHW_CUDA_H */
