/*
 * include/nhw_compat_dec.h -- the reference DECODER's per-image entry points on top of
 * libnhw_cuda, so that decoder/nhw_decoder_cli.c (which itself contains main, write_image_bmp
 * and setup_bmp_header) links against this repository unchanged.
 *
 *   void decode_image(image_buffer*, decode_state*, char*)   decoder/codec.h:198, decoder/nhw_decoder.c:54
 *   int  parse_file(image_buffer*, decode_state*, char*)     decoder/codec.h:199, decoder/nhw_decoder.c:1478
 *
 * decode_image leaves what the reference leaves for its writer: im->setup (malloc'd, with
 * quality_setting) and the three malloc'd 512x512 byte planes im_bufferY/U/V.  The structs mirror
 * decoder/codec.h:118-137 (only the fields the CLI touches matter; decode_state is opaque here).
 */
#ifndef NHW_COMPAT_DEC_H
#define NHW_COMPAT_DEC_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	unsigned char colorspace, wavelet_type, RES_HIGH, RES_LOW, wvlts_order, quality_setting;
} nhw_dec_codec_setup;

typedef struct {
	short *im_process;
	short *im_jpeg;
	unsigned char *im_bufferY;
	unsigned char *im_bufferU;
	unsigned char *im_bufferV;
	unsigned char *im_buffer4;
	short *im_nhw3;
	unsigned char *scale;
	nhw_dec_codec_setup *setup;
} nhw_dec_image_buffer;

void decode_image(nhw_dec_image_buffer *im, void *decode_state, char *file_name);
int parse_file(nhw_dec_image_buffer *im, void *decode_state, char *file_name);

#ifdef __cplusplus
}
#endif
#endif
