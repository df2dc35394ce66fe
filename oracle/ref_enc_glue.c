/*
 * oracle/ref_enc_glue.c -- TEST INFRASTRUCTURE ONLY.
 *
 * In-memory entry points around the UNMODIFIED reference encoder objects
 * (/root/reference/encoder/*.c, compiled where they lie by oracle/build_ref.sh into
 * oracle/_ref/libnhwref_enc.so).  Nothing here re-implements codec arithmetic: it only
 * sets up the structs exactly like the reference CLI does
 * (encoder/nhw_encoder_cli.c:175-183, encoder/nhw_encoder.c:3055-3060) and moves bytes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include <unistd.h>
#include "codec.h"   /* the reference's own header, via -I/root/reference/encoder */
#include "tap.h"

extern __thread jmp_buf nhwref_exit_jmp;
extern __thread int nhwref_exit_armed;
extern __thread int nhwref_exit_code;

/* ---------------- tap registry ---------------- */
#define MAX_TAPS 256
static int g_tap_on = 0;
static int g_ntaps = 0;
static struct { char name[48]; void *data; size_t bytes; } g_taps[MAX_TAPS];

void *__real_malloc(size_t);
void __real_free(void *);

void nhw_tap(const char *name, const void *ptr, size_t bytes)
{
	if (!g_tap_on || g_ntaps >= MAX_TAPS) return;
	strncpy(g_taps[g_ntaps].name, name, 47);
	g_taps[g_ntaps].name[47] = 0;
	g_taps[g_ntaps].data = __real_malloc(bytes ? bytes : 1);
	memcpy(g_taps[g_ntaps].data, ptr, bytes);
	g_taps[g_ntaps].bytes = bytes;
	g_ntaps++;
}

void nhwref_tap_enable(int on)
{
	int i;
	for (i = 0; i < g_ntaps; i++) __real_free(g_taps[i].data);
	g_ntaps = 0;
	g_tap_on = on;
}
int nhwref_tap_count(void) { return g_ntaps; }
const char *nhwref_tap_name(int i) { return g_taps[i].name; }
const void *nhwref_tap_data(int i) { return g_taps[i].data; }
size_t nhwref_tap_bytes(int i) { return g_taps[i].bytes; }

/* ---------------- encode ---------------- */

/* The reference also reads never-written STACK memory: wavlts2packet keeps its de-interleaved codebook in a local array
 * (encoder/compress_pixel.c:58) and its run-length loop looks one entry past the end of the list (:412, :446), so what a
 * call finds there is whatever EARLIER calls in the same process left at that stack depth (seen as a one-longer final run
 * in tree1 on about 1 image in 4000, depending on which images were encoded before it).  The zero-guard allocator cannot
 * reach the stack; the canonical oracle defines such reads as 0 like every other never-written read,
 * so the stack below the caller's frame is cleared before every call into the reference. */
#ifndef NHW_NO_TAPS
static void __attribute__((noinline)) scrub_stack(void)
{
	unsigned char pad[768 * 1024];
	memset(pad, 0, sizeof pad);
	__asm__ volatile("" : : "r"(pad) : "memory");
}
#else
static void scrub_stack(void) {}
#endif

/* Runs downsample_YUV420 + encode_image on 786432 raw BMP pixel bytes (file order,
 * i.e. what read_image_bmp fread()s at encoder/nhw_encoder.c:3086).  If out_path is
 * non-NULL the stream is written there by the reference's own write_compressed_file.
 * Returns 0, or 1000+|exit code| if the reference called exit(). */
static int run_encode(const unsigned char *pix, int quality, const char *out_path)
{
	image_buffer im;
	encode_state enc;
	codec_setup setup;
	memset(&im, 0, sizeof im);
	memset(&enc, 0, sizeof enc);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.quality_setting = (unsigned char)quality;
	setup.colorspace = YUV;
	setup.wavelet_type = WVLTS_53;
	setup.RES_HIGH = 0;
	setup.RES_LOW = 3;
	setup.wvlts_order = 2;

	nhwref_exit_armed = 1;
	if (setjmp(nhwref_exit_jmp)) {
		nhwref_exit_armed = 0;
		return 1000 + abs(nhwref_exit_code);
	}
	im.im_buffer4 = (unsigned char *)calloc(4 * 3 * IM_SIZE, 1);
	memcpy(im.im_buffer4, pix, 4 * 3 * IM_SIZE);
	NHW_TAP("in_rgb", im.im_buffer4, 4 * 3 * IM_SIZE);
	scrub_stack();
	downsample_YUV420(&im, 8);
	NHW_TAP("cs_Y", im.im_jpeg, 4 * IM_SIZE * sizeof(short));
	NHW_TAP("cs_U", im.im_bufferU, IM_SIZE);
	NHW_TAP("cs_V", im.im_bufferV, IM_SIZE);
	encode_image(&im, &enc, 8);
	write_compressed_file(&im, &enc, (char *)(out_path ? out_path : "/dev/null"));
	nhwref_exit_armed = 0;
	return 0;
}

static long slurp(const char *path, unsigned char *out, long cap)
{
	FILE *f = fopen(path, "rb");
	long n;
	if (!f) return -1;
	n = (long)fread(out, 1, (size_t)cap, f);
	fclose(f);
	return n;
}

/* Encode to memory.  Returns the .nhw length (>0) or a negative error. */
long nhwref_encode(const unsigned char *pix, int quality, unsigned char *out, long cap)
{
	char path[64];
	int fd, rc;
	long n;
	strcpy(path, "/dev/shm/nhwref_XXXXXX");
	fd = mkstemp(path);
	if (fd < 0) { strcpy(path, "/tmp/nhwref_XXXXXX"); fd = mkstemp(path); }
	if (fd < 0) return -1;
	close(fd);
	rc = run_encode(pix, quality, path);
	if (rc) { unlink(path); return -rc; }
	n = slurp(path, out, cap);
	unlink(path);
	return n;
}

/* Encode and discard (CPU-baseline timing: no temp-file traffic beyond /dev/null). */
int nhwref_encode_discard(const unsigned char *pix, int quality)
{
	return run_encode(pix, quality, NULL);
}

/* Full CLI-equivalent path: read_image_bmp on a real file (header check, flip). */
long nhwref_encode_bmp_file(const char *bmp_path, int quality, const char *out_path)
{
	image_buffer im;
	encode_state enc;
	codec_setup setup;
	memset(&im, 0, sizeof im);
	memset(&enc, 0, sizeof enc);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.quality_setting = (unsigned char)quality;
	nhwref_exit_armed = 1;
	if (setjmp(nhwref_exit_jmp)) {
		nhwref_exit_armed = 0;
		return -(1000 + abs(nhwref_exit_code));
	}
	read_image_bmp((char *)bmp_path, &enc, &im, 8);
	encode_image(&im, &enc, 8);
	write_compressed_file(&im, &enc, (char *)out_path);
	nhwref_exit_armed = 0;
	return 0;
}

/* ---------------- individually callable stages (all extern in encoder/codec.h:184-219) --- */

/* colour stage only: pix -> Y int16[512*512], U u8[256*256], V u8[256*256] */
int nhwref_stage_colorspace(const unsigned char *pix, int quality, short *Y, unsigned char *U, unsigned char *V)
{
	image_buffer im;
	codec_setup setup;
	memset(&im, 0, sizeof im);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.quality_setting = (unsigned char)quality;
	im.im_buffer4 = (unsigned char *)calloc(4 * 3 * IM_SIZE, 1);
	memcpy(im.im_buffer4, pix, 4 * 3 * IM_SIZE);
	downsample_YUV420(&im, 8);
	memcpy(Y, im.im_jpeg, 4 * IM_SIZE * sizeof(short));
	memcpy(U, im.im_bufferU, IM_SIZE);
	memcpy(V, im.im_bufferV, IM_SIZE);
	free(im.im_jpeg); free(im.im_bufferU); free(im.im_bufferV);
	return 0;
}

/* pre_processing on a Y plane in place (encoder/image_processing.c:558) */
int nhwref_stage_pre_processing(short *Y, int quality)
{
	image_buffer im;
	codec_setup setup;
	memset(&im, 0, sizeof im);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.quality_setting = (unsigned char)quality;
	im.im_jpeg = (short *)malloc(4 * IM_SIZE * sizeof(short));
	im.im_process = (short *)malloc(4 * IM_SIZE * sizeof(short));
	memcpy(im.im_jpeg, Y, 4 * IM_SIZE * sizeof(short));
	pre_processing(&im);
	memcpy(Y, im.im_jpeg, 4 * IM_SIZE * sizeof(short));
	free(im.im_jpeg); free(im.im_process);
	return 0;
}

/* Two-level luma analysis exactly as encode_image sequences it
 * (encoder/nhw_encoder.c:125-139): out_proc = im_process after level 2,
 * out_ll1 = the 256x256 LL1 copy ("res256"). */
int nhwref_stage_dwt_y(const short *Y, int quality, short *out_proc, short *out_ll1)
{
	image_buffer im;
	codec_setup setup;
	int i, j, c;
	memset(&im, 0, sizeof im);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.quality_setting = (unsigned char)quality;
	setup.wvlts_order = 2;
	im.im_jpeg = (short *)malloc(4 * IM_SIZE * sizeof(short));
	im.im_process = (short *)malloc(4 * IM_SIZE * sizeof(short));
	memcpy(im.im_jpeg, Y, 4 * IM_SIZE * sizeof(short));
	wavelet_analysis(&im, 2 * IM_DIM, 0, 1);
	for (i = 0, c = 0; i < 2 * IM_SIZE; i += 2 * IM_DIM)
		for (j = 0; j < IM_DIM; j++) out_ll1[c++] = im.im_jpeg[i + j];
	wavelet_analysis(&im, IM_DIM, 1, 1);
	memcpy(out_proc, im.im_process, 4 * IM_SIZE * sizeof(short));
	if (im.im_quality_setting) free(im.im_quality_setting);
	free(im.im_jpeg); free(im.im_process);
	return 0;
}

/* Generic single-level calls on caller-owned planes (stride 512 for Y=1, 256 for Y=0). */
int nhwref_wavelet_analysis(short *jpeg, short *proc, int norder, int last_stage, int Y, int quality)
{
	image_buffer im;
	codec_setup setup;
	memset(&im, 0, sizeof im);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.quality_setting = (unsigned char)quality;
	setup.wvlts_order = 2;
	im.im_jpeg = jpeg;
	im.im_process = proc;
	wavelet_analysis(&im, norder, last_stage, Y);
	if (im.im_quality_setting) free(im.im_quality_setting);
	return 0;
}

int nhwref_wavelet_synthesis(short *jpeg, short *proc, int norder, int last_stage, int Y)
{
	image_buffer im;
	codec_setup setup;
	memset(&im, 0, sizeof im);
	memset(&setup, 0, sizeof setup);
	im.setup = &setup;
	setup.wvlts_order = 2;
	setup.wavelet_type = WVLTS_53;
	im.im_jpeg = jpeg;
	im.im_process = proc;
	wavelet_synthesis(&im, norder, last_stage, Y);
	return 0;
}
