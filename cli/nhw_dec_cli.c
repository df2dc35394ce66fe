/*
 * cli/nhw_dec_cli.c -- `nhw-dec in.nhw out.bmp`, same behaviour as the reference CLI
 * (decoder/nhw_decoder_cli.c:67-93): no flags, fixed 54-byte BMP header, 786432 pixel bytes.
 * The whole decode, including the colour conversion, runs in nhw_decode_batch().
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/nhw_cuda.h"

int main(int argc, char **argv)
{
	/* the reference's fixed header for a 512x512 24-bit image (decoder/nhw_decoder_cli.c:61-65) */
	static const unsigned char header[54] = {66, 77, 54, 0, 12, 0, 0, 0, 0, 0, 54, 0, 0, 0, 40, 0, 0, 0, 0, 2, 0, 0, 0, 2, 0, 0, 1,
	                                         0, 24, 0, 0, 0, 0, 0, 0, 0, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	static unsigned char blob[NHW_MAX_STREAM_BYTES], rgb[NHW_PIX_BYTES];
	nhw_ctx *ctx;
	uint64_t offs[2];
	int32_t status = 0;
	const char *dev = getenv("NHW_CUDA_DEVICE");
	FILE *f;
	int rc;
	if (argc < 3) {
		printf("usage: nhw-dec <image.nhw> <image.bmp>\n  decodes a .nhw file to a 512x512 24-bit BMP\n");
		return 0;
	}
	if ((f = fopen(argv[1], "rb")) == NULL) { printf("\nCould not open file\n"); exit(-1); }
	offs[0] = 0;
	offs[1] = fread(blob, 1, sizeof blob, f);
	fclose(f);
	if (offs[1] >= 1 && blob[0] > 6) { printf("\nNot an .nhw file"); exit(-1); }
	if ((rc = nhw_create(dev ? atoi(dev) : 0, 1, &ctx)) != NHW_OK) {
		fprintf(stderr, "nhw: cannot create CUDA codec context (%d): %s\n", rc, nhw_last_error());
		return 1;
	}
	rc = nhw_decode_batch(ctx, blob, offs, 1, rgb, &status);
	nhw_destroy(ctx);
	if (rc != NHW_OK || status != NHW_OK) { fprintf(stderr, "nhw: decode failed (%d/%d): %s\n", rc, (int)status, nhw_last_error()); return 1; }
	if ((f = fopen(argv[2], "wb")) == NULL) { printf("Failed to open output decompressed .bmp file %s\n", argv[2]); return 1; }
	fwrite(header, 54, 1, f);
	fwrite(rgb, NHW_PIX_BYTES, 1, f);
	fclose(f);
	return 0;
}
