/*
 * compat.c -- libnhw_compat.so: the reference's per-image encoder entry points on top of the
 * batch C-ABI (see include/nhw_compat.h).  Host code, plain C, no codec arithmetic here: the
 * pixels go to nhw_encode_batch() and the returned .nhw stream is split back into the
 * encode_state fields the reference's writer expects (layout: SURVEY.md Appendix A).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nhw_compat.h"
#include "../../include/nhw_cuda.h"

/* first 34 bytes of the last BMP header seen, like the reference's global (encoder/nhw_encoder.c:73) */
unsigned char bmp_header[54];

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

static unsigned rd16(const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }
static unsigned rd32(const unsigned char *p) { return rd16(p) | (rd16(p + 2) << 16); }

/* BMP validation with the reference's exit codes (encoder/nhw_encoder.c:63-71, 2902-3025):
 * -10 null pointer, -11 seek, -12 short read, -13 not "BM", -14 unknown info header,
 * -15 planes != 1, -16 not 512x512 24-bit BI_RGB. */
static int check_bmp(FILE *f, int *data_offset, int *flipped)
{
	int bih, width, height, planes, bpp, compr;
	if (fseek(f, 0, SEEK_SET) != 0) return -11;
	if (fread(bmp_header, 1, 34, f) < 34) return -12;
	if (bmp_header[0] != 'B' || bmp_header[1] != 'M') return -13;
	*data_offset = (int)rd32(bmp_header + 10);
	bih = (int)rd32(bmp_header + 14);
	if (bih != 12 && bih != 40 && bih != 52 && bih != 56 && bih != 108 && bih != 124) return -14;
	if (bih == 12) {
		width = (int)rd16(bmp_header + 18);
		height = (int)rd16(bmp_header + 20);
		planes = (short)rd16(bmp_header + 22);
		bpp = (short)rd16(bmp_header + 24);
		compr = 0;
	} else {
		width = (int)rd32(bmp_header + 18);
		height = (int)rd32(bmp_header + 22);
		planes = (short)rd16(bmp_header + 26);
		bpp = (short)rd16(bmp_header + 28);
		compr = (int)rd32(bmp_header + 30);
	}
	if (planes != 1) return -15;
	if (!(width == 512 && (height == 512 || height == -512) && bpp == 24 && compr == 0)) return -16;
	*flipped = height < 0;
	return 0;
}

int read_image_bmp(char *file_name, encode_state *os, image_buffer *im, int rate)
{
	FILE *f;
	int rc, data_offset = 0, flipped = 0;
	(void)os; (void)rate;
	im->setup->colorspace = 1;      /* YUV */
	im->setup->wavelet_type = 0;    /* WVLTS_53 */
	im->setup->RES_HIGH = 0;
	im->setup->RES_LOW = 3;
	im->setup->wvlts_order = 2;
	im->im_buffer4 = (unsigned char *)calloc(NHW_PIX_BYTES, 1);
	if ((f = fopen(file_name, "rb")) == NULL) {
		printf("menu(): Could not open file: %s\n", file_name);
		exit(-1);
	}
	if ((rc = check_bmp(f, &data_offset, &flipped)) != 0) {
		printf("invalid image file.\n");
		exit(rc);
	}
	if (fseek(f, data_offset, SEEK_SET) != 0) {
		printf("unable to seek to actual data.\n");
		exit(-2);
	}
	if (fread(im->im_buffer4, NHW_PIX_BYTES, 1, f) != 1) { /* the reference ignores short reads too */ }
	fclose(f);
	if (flipped) {   /* top-down BMP: bring it to the bottom-up row order the codec expects */
		unsigned char *tmp = (unsigned char *)malloc(1536);
		int y;
		for (y = 0; tmp && y < 256; y++) {
			memcpy(tmp, im->im_buffer4 + 1536 * y, 1536);
			memcpy(im->im_buffer4 + 1536 * y, im->im_buffer4 + 1536 * (511 - y), 1536);
			memcpy(im->im_buffer4 + 1536 * (511 - y), tmp, 1536);
		}
		free(tmp);
	}
	return 0;
}

static unsigned char *take(const unsigned char **p, size_t n)
{
	unsigned char *d = (unsigned char *)malloc(n ? n : 1);
	memcpy(d, *p, n);
	*p += n;
	return d;
}

void encode_image(image_buffer *im, encode_state *enc, int ratio)
{
	static unsigned char stream[NHW_MAX_STREAM_BYTES];
	uint64_t offs[2];
	int32_t status = 0;
	const unsigned char *p = stream;
	const int q = im->setup->quality_setting;
	int rc;
	(void)ratio;
	rc = nhw_encode_batch(ctx(), im->im_buffer4, 1, q, stream, sizeof stream, offs, &status);
	free(im->im_buffer4);
	im->im_buffer4 = NULL;
	if (rc != NHW_OK || status != NHW_OK) {
		fprintf(stderr, "nhw: encode failed (%d/%d): %s\n", rc, (int)status, nhw_last_error());
		exit(-1);   /* the reference exits with -1 on codebook overflow (encoder/compress_pixel.c:234,270-271) */
	}
	memset(enc, 0, sizeof *enc);
	/* header */
	im->setup->RES_HIGH = (unsigned char)(p[0] & 3);
	im->setup->wavelet_type = (unsigned char)(p[0] & 4);
	p += 2;
	enc->size_tree1 = (unsigned short)rd16(p); p += 2;
	enc->size_tree2 = (unsigned short)rd16(p); p += 2;
	enc->size_data1 = (int)rd32(p); p += 4;
	enc->size_data2 = (int)rd32(p); p += 4;
	enc->tree_end = (unsigned short)rd16(p); p += 2;
	enc->exw_Y_end = (unsigned short)rd16(p); p += 2;
	if (q > 12) { enc->nhw_res1_len = (unsigned short)rd16(p); p += 2; }
	if (q >= 19) { enc->nhw_res3_len = (unsigned short)rd16(p); p += 2; enc->nhw_res3_bit_len = (unsigned short)rd16(p); p += 2; }
	if (q > 17) { enc->nhw_res4_len = (unsigned short)rd16(p); p += 2; }
	if (q > 12) { enc->nhw_res1_bit_len = (unsigned short)rd16(p); p += 2; }
	if (q >= 21) { enc->nhw_res5_len = (unsigned short)rd16(p); p += 2; enc->nhw_res5_bit_len = (unsigned short)rd16(p); p += 2; }
	if (q > 21) { enc->nhw_res6_len = rd32(p); p += 4; enc->nhw_res6_bit_len = (unsigned short)rd16(p); p += 2; enc->nhw_char_res1_len = (unsigned short)rd16(p); p += 2; }
	if (q > 22) { enc->qsetting3_len = (unsigned short)rd16(p); p += 2; }
	enc->nhw_select1 = (unsigned short)rd16(p); p += 2;
	enc->nhw_select2 = (unsigned short)rd16(p); p += 2;
	if (q > 15) { enc->highres_comp_len = (unsigned short)rd16(p); p += 2; }
	enc->end_ch_res = (unsigned short)rd16(p); p += 2;
	/* sections */
	enc->tree1 = take(&p, enc->size_tree1);
	enc->tree2 = take(&p, enc->size_tree2);
	enc->exw_Y = take(&p, enc->exw_Y_end);
	if (q > 12) {
		enc->nhw_res1 = take(&p, enc->nhw_res1_len);
		enc->nhw_res1_bit = take(&p, enc->nhw_res1_bit_len);
		enc->nhw_res1_word_len = enc->nhw_res1_bit_len;
		enc->nhw_res1_word = take(&p, enc->nhw_res1_word_len);
	}
	if (q > 17) enc->nhw_res4 = take(&p, enc->nhw_res4_len);
	if (q >= 19) {
		enc->nhw_res3 = take(&p, enc->nhw_res3_len);
		enc->nhw_res3_bit = take(&p, enc->nhw_res3_bit_len);
		enc->nhw_res3_word_len = (unsigned short)(2 * enc->nhw_res3_bit_len);
		enc->nhw_res3_word = take(&p, enc->nhw_res3_word_len);
	}
	if (q >= 21) {
		enc->nhw_res5 = take(&p, enc->nhw_res5_len);
		enc->nhw_res5_bit = take(&p, enc->nhw_res5_bit_len);
		enc->nhw_res5_word_len = enc->nhw_res5_bit_len;
		enc->nhw_res5_word = take(&p, enc->nhw_res5_word_len);
	}
	if (q > 21) {
		enc->nhw_res6 = take(&p, enc->nhw_res6_len);
		enc->nhw_res6_bit = take(&p, enc->nhw_res6_bit_len);
		enc->nhw_res6_word_len = enc->nhw_res6_bit_len;
		enc->nhw_res6_word = take(&p, enc->nhw_res6_word_len);
		enc->nhw_char_res1 = (unsigned short *)take(&p, 2u * enc->nhw_char_res1_len);
	}
	if (q > 22) enc->high_qsetting3 = (unsigned int *)take(&p, 4u * enc->qsetting3_len);
	enc->nhw_select_word1 = take(&p, enc->nhw_select1);
	enc->nhw_select_word2 = take(&p, enc->nhw_select2);
	if (q > 15) {
		enc->res_U_64 = take(&p, 512);
		enc->res_V_64 = take(&p, 512);
		enc->highres_word = take(&p, enc->highres_comp_len);
	}
	enc->ch_res = take(&p, enc->end_ch_res);
	enc->encode = (unsigned int *)take(&p, 4u * (size_t)enc->size_data2);
}

int write_compressed_file(image_buffer *im, encode_state *enc, char *file_name)
{
	const int q = im->setup->quality_setting;
	FILE *f = fopen(file_name, "wb");
	if (f == NULL) {
		printf("Failed to create file: %s\n", file_name);
		return -1;
	}
	im->setup->RES_HIGH += im->setup->wavelet_type;
	fwrite(&im->setup->RES_HIGH, 1, 1, f);
	fwrite(&im->setup->quality_setting, 1, 1, f);
	fwrite(&enc->size_tree1, 2, 1, f);
	fwrite(&enc->size_tree2, 2, 1, f);
	fwrite(&enc->size_data1, 4, 1, f);
	fwrite(&enc->size_data2, 4, 1, f);
	fwrite(&enc->tree_end, 2, 1, f);
	fwrite(&enc->exw_Y_end, 2, 1, f);
	if (q > 12) fwrite(&enc->nhw_res1_len, 2, 1, f);
	if (q >= 19) { fwrite(&enc->nhw_res3_len, 2, 1, f); fwrite(&enc->nhw_res3_bit_len, 2, 1, f); }
	if (q > 17) fwrite(&enc->nhw_res4_len, 2, 1, f);
	if (q > 12) fwrite(&enc->nhw_res1_bit_len, 2, 1, f);
	if (q >= 21) { fwrite(&enc->nhw_res5_len, 2, 1, f); fwrite(&enc->nhw_res5_bit_len, 2, 1, f); }
	if (q > 21) { fwrite(&enc->nhw_res6_len, 4, 1, f); fwrite(&enc->nhw_res6_bit_len, 2, 1, f); fwrite(&enc->nhw_char_res1_len, 2, 1, f); }
	if (q > 22) fwrite(&enc->qsetting3_len, 2, 1, f);
	fwrite(&enc->nhw_select1, 2, 1, f);
	fwrite(&enc->nhw_select2, 2, 1, f);
	if (q > 15) fwrite(&enc->highres_comp_len, 2, 1, f);
	fwrite(&enc->end_ch_res, 2, 1, f);
	fwrite(enc->tree1, enc->size_tree1, 1, f);
	fwrite(enc->tree2, enc->size_tree2, 1, f);
	fwrite(enc->exw_Y, enc->exw_Y_end, 1, f);
	if (q > 12) { fwrite(enc->nhw_res1, enc->nhw_res1_len, 1, f); fwrite(enc->nhw_res1_bit, enc->nhw_res1_bit_len, 1, f); fwrite(enc->nhw_res1_word, enc->nhw_res1_word_len, 1, f); }
	if (q > 17) fwrite(enc->nhw_res4, enc->nhw_res4_len, 1, f);
	if (q >= 19) { fwrite(enc->nhw_res3, enc->nhw_res3_len, 1, f); fwrite(enc->nhw_res3_bit, enc->nhw_res3_bit_len, 1, f); fwrite(enc->nhw_res3_word, enc->nhw_res3_word_len, 1, f); }
	if (q >= 21) { fwrite(enc->nhw_res5, enc->nhw_res5_len, 1, f); fwrite(enc->nhw_res5_bit, enc->nhw_res5_bit_len, 1, f); fwrite(enc->nhw_res5_word, enc->nhw_res5_word_len, 1, f); }
	if (q > 21) { fwrite(enc->nhw_res6, enc->nhw_res6_len, 1, f); fwrite(enc->nhw_res6_bit, enc->nhw_res6_bit_len, 1, f); fwrite(enc->nhw_res6_word, enc->nhw_res6_word_len, 1, f); fwrite(enc->nhw_char_res1, enc->nhw_char_res1_len, 2, f); }
	if (q > 22) fwrite(enc->high_qsetting3, enc->qsetting3_len, 4, f);
	fwrite(enc->nhw_select_word1, enc->nhw_select1, 1, f);
	fwrite(enc->nhw_select_word2, enc->nhw_select2, 1, f);
	if (q > 15) { fwrite(enc->res_U_64, 512, 1, f); fwrite(enc->res_V_64, 512, 1, f); fwrite(enc->highres_word, enc->highres_comp_len, 1, f); }
	fwrite(enc->ch_res, enc->end_ch_res, 1, f);
	fwrite(enc->encode, (size_t)enc->size_data2 * 4, 1, f);
	fclose(f);
	free(enc->encode); free(enc->tree1); free(enc->tree2); free(enc->exw_Y);
	free(enc->nhw_res1); free(enc->nhw_res1_bit); free(enc->nhw_res1_word);
	free(enc->nhw_res3); free(enc->nhw_res3_bit); free(enc->nhw_res3_word); free(enc->nhw_res4);
	free(enc->nhw_res5); free(enc->nhw_res5_bit); free(enc->nhw_res5_word);
	free(enc->nhw_res6); free(enc->nhw_res6_bit); free(enc->nhw_res6_word); free(enc->nhw_char_res1);
	free(enc->high_qsetting3); free(enc->nhw_select_word1); free(enc->nhw_select_word2);
	free(enc->res_U_64); free(enc->res_V_64); free(enc->highres_word); free(enc->ch_res);
	memset(enc, 0, sizeof *enc);
	return 0;
}
