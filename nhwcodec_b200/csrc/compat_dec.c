/*
 * compat_dec.c -- libnhw_compat_dec.so: decode_image / parse_file with the reference's
 * signatures (see include/nhw_compat_dec.h) on top of nhw_decode_batch_planes(n = 1).
 * Separate from the encoder compat library because both reference programs define
 * same-named but different structs and helpers.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nhw_compat_dec.h"
#include "../../include/nhw_cuda.h"

static nhw_ctx *g_ctx = NULL;

static nhw_ctx *ctx(void)
{
	if (!g_ctx) {
		const char *d = getenv("NHW_CUDA_DEVICE");
		int rc = nhw_create(d ? atoi(d) : 0, 1, &g_ctx);
		if (rc != NHW_OK) {
			fprintf(stderr, "nhw: cannot create CUDA codec context (%d): %s\n", rc, nhw_last_error());
			exit(-1);
		}
	}
	return g_ctx;
}

int parse_file(nhw_dec_image_buffer *im, void *decode_state, char *file_name)
{
	(void)im; (void)decode_state; (void)file_name;
	return 2;   /* wavelet order; the container is parsed inside decode_image here */
}

void decode_image(nhw_dec_image_buffer *im, void *decode_state, char *file_name)
{
	static unsigned char blob[NHW_MAX_STREAM_BYTES];
	unsigned char *yuv;
	uint64_t offs[2];
	int32_t status = 0, quality = 0;
	size_t n;
	int rc;
	FILE *f = fopen(file_name, "rb");
	(void)decode_state;
	if (f == NULL) { printf("\nCould not open file\n"); exit(-1); }      /* decoder/nhw_decoder.c:1488-1491 */
	n = fread(blob, 1, sizeof blob, f);
	fclose(f);
	if (n >= 1 && blob[0] > 6) { printf("\nNot an .nhw file"); exit(-1); }  /* decoder/nhw_decoder.c:1497-1500 */
	yuv = (unsigned char *)malloc(NHW_PIX_BYTES);
	offs[0] = 0;
	offs[1] = n;
	rc = nhw_decode_batch_planes(ctx(), blob, offs, 1, yuv, &quality, &status);
	if (rc != NHW_OK || status != NHW_OK) {
		fprintf(stderr, "nhw: decode failed (%d/%d): %s\n", rc, (int)status, nhw_last_error());
		exit(-1);
	}
	im->setup = (nhw_dec_codec_setup *)calloc(1, sizeof(nhw_dec_codec_setup));
	im->setup->colorspace = 1;
	im->setup->wvlts_order = 2;
	im->setup->RES_HIGH = blob[0];
	im->setup->quality_setting = (unsigned char)quality;
	im->im_bufferY = (unsigned char *)malloc(512 * 512);
	im->im_bufferU = (unsigned char *)malloc(512 * 512);
	im->im_bufferV = (unsigned char *)malloc(512 * 512);
	memcpy(im->im_bufferY, yuv, 512 * 512);
	memcpy(im->im_bufferU, yuv + 512 * 512, 512 * 512);
	memcpy(im->im_bufferV, yuv + 2 * 512 * 512, 512 * 512);
	free(yuv);
}
