/*
 * oracle/ref_dec_glue.c -- TEST INFRASTRUCTURE ONLY.
 * In-memory entry points around the UNMODIFIED reference decoder objects
 * (/root/reference/decoder/*.c -> oracle/_ref/libnhwref_dec.so).  Mirrors
 * decoder/nhw_decoder_cli.c:67-93 (main) without re-implementing any arithmetic.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include <unistd.h>
#include "codec.h"   /* /root/reference/decoder/codec.h */
#include "tap.h"

extern __thread jmp_buf nhwref_exit_jmp;
extern __thread int nhwref_exit_armed;
extern __thread int nhwref_exit_code;
extern unsigned char bmp_header[54];
int setup_bmp_header(char *bmp_header, nhw_image_s *im_ctx);
int write_image_bmp(image_buffer *im, char *file_name);

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

static int tmp_path(char *path)
{
	int fd;
	strcpy(path, "/dev/shm/nhwref_XXXXXX");
	fd = mkstemp(path);
	if (fd < 0) { strcpy(path, "/tmp/nhwref_XXXXXX"); fd = mkstemp(path); }
	if (fd < 0) return -1;
	close(fd);
	return 0;
}

/* Never-written stack memory reads as 0 in the canonical oracle (see ref_enc_glue.c: scrub_stack). */
static void __attribute__((noinline)) scrub_stack(void)
{
	unsigned char pad[768 * 1024];
	memset(pad, 0, sizeof pad);
	__asm__ volatile("" : : "r"(pad) : "memory");
}

/* Decode a .nhw byte string to 786432 BMP pixel bytes (file order, header stripped).
 * planes (optional, 3*262144 bytes) receives the Y,U,V u8 planes decode_image leaves
 * for the writer (decoder/nhw_decoder.c:877-891,1137-1181). */
long nhwref_decode(const unsigned char *nhw, long len, unsigned char *out_pix, unsigned char *planes, int to_bmp)
{
	char in_path[64], out_path[64];
	image_buffer im;
	decode_state dec;
	nhw_image_s ctx;
	FILE *f;
	long n = 0;
	memset(&im, 0, sizeof im);
	memset(&dec, 0, sizeof dec);
	if (tmp_path(in_path)) return -1;
	f = fopen(in_path, "wb");
	fwrite(nhw, 1, (size_t)len, f);
	fclose(f);
	nhwref_exit_armed = 1;
	if (setjmp(nhwref_exit_jmp)) {
		nhwref_exit_armed = 0;
		unlink(in_path);
		return -(1000 + abs(nhwref_exit_code));
	}
	scrub_stack();
	decode_image(&im, &dec, in_path);
	unlink(in_path);
	if (planes) {
		memcpy(planes, im.im_bufferY, 4 * IM_SIZE);
		memcpy(planes + 4 * IM_SIZE, im.im_bufferU, 4 * IM_SIZE);
		memcpy(planes + 8 * IM_SIZE, im.im_bufferV, 4 * IM_SIZE);
	}
	if (to_bmp) {
		if (tmp_path(out_path)) return -1;
		ctx.width = 512; ctx.height = 512; ctx.bpp = 24;
		setup_bmp_header((char *)bmp_header, &ctx);
		write_image_bmp(&im, out_path);
		f = fopen(out_path, "rb");
		fseek(f, 54, SEEK_SET);
		n = (long)fread(out_pix, 1, 12 * IM_SIZE, f);
		fclose(f);
		unlink(out_path);
	} else {
		free(im.im_bufferY); free(im.im_bufferU); free(im.im_bufferV);
	}
	nhwref_exit_armed = 0;
	return n;
}
