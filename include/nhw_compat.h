/*
 * include/nhw_compat.h -- the reference's own per-image entry points, provided on top of
 * libnhw_cuda so that rcanut/nhwcodec's CLI sources link against this repository unchanged.
 *
 * Each declaration replaces the reference symbol of the same name:
 *   read_image_bmp, encode_image, write_compressed_file   encoder/codec.h:184-187
 *   bmp_header (encoder)                                   encoder/nhw_encoder.c:73
 * The three structs mirror encoder/codec.h:103-181 field for field: they are the ABI the
 * reference CLI (encoder/nhw_encoder_cli.c:88-186) is compiled against.  Ownership and error
 * conventions are the reference's: read_image_bmp allocates the pixel
 * buffer and exit()s on bad input with the reference's codes; encode_image consumes it and
 * fills encode_state with malloc'd arrays; write_compressed_file frees them.
 *
 * What differs: read_image_bmp keeps the raw BMP pixel bytes (the colour transform runs on
 * the GPU inside encode_image), and encode_image runs the whole path through
 * nhw_encode_batch() with n = 1 on CUDA device NHW_CUDA_DEVICE (default 0).
 */
#ifndef NHW_COMPAT_H
#define NHW_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	unsigned char colorspace;
	unsigned char wavelet_type;
	unsigned char RES_HIGH;
	unsigned char RES_LOW;
	unsigned char wvlts_order;
	unsigned char quality_setting;
} codec_setup;

typedef struct {
	short *im_process;
	short *im_jpeg;
	unsigned char *im_bufferU;
	unsigned char *im_bufferV;
	unsigned char *im_buffer4;   /* raw BMP pixel bytes between read_image_bmp and encode_image */
	unsigned char *im_nhw;
	short *im_wavelet_first_order;
	short *im_quality_setting;
	short *im_wavelet_band;
	codec_setup *setup;
} image_buffer;

typedef struct {
	unsigned int *encode;
	unsigned char *tree1;
	unsigned char *tree2;
	unsigned short nhw_res1_len, nhw_res3_len, nhw_res4_len, nhw_res5_len;
	unsigned int nhw_res6_len;
	unsigned short nhw_res1_word_len, nhw_res3_word_len, nhw_res5_word_len, nhw_res6_word_len;
	unsigned short nhw_res1_bit_len, nhw_res3_bit_len, nhw_res5_bit_len, nhw_res6_bit_len;
	unsigned char *nhw_res1, *nhw_res3, *nhw_res4, *nhw_res5, *nhw_res6;
	unsigned char *nhw_res1_bit, *nhw_res3_bit, *nhw_res5_bit, *nhw_res6_bit;
	unsigned char *nhw_res1_word, *nhw_res3_word, *nhw_res5_word, *nhw_res6_word;
	unsigned short *nhw_char_res1;
	unsigned short nhw_char_res1_len;
	unsigned short nhw_select1, nhw_select2;
	unsigned char *nhw_select_word1, *nhw_select_word2;
	int size_data1, size_data2;
	unsigned short size_tree1, size_tree2, tree_end, Y_res_comp, exw_Y_end, end_ch_res, qsetting3_len;
	unsigned int *high_qsetting3;
	unsigned short highres_mem_len, highres_comp_len;
	unsigned short *highres_mem;
	unsigned char *highres_comp;
	unsigned char *highres_word;
	unsigned char *res_U_64, *res_V_64;
	unsigned char *exw_Y;
	unsigned char *ch_res;
	unsigned int *high_res;
} encode_state;

extern unsigned char bmp_header[];

int read_image_bmp(char *file_name, encode_state *os, image_buffer *im, int rate);
void encode_image(image_buffer *im, encode_state *enc, int ratio);
int write_compressed_file(image_buffer *im, encode_state *enc, char *file_name);

#ifdef __cplusplus
}
#endif
#endif
