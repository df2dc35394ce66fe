/*
 * include/nhw_batchio.h -- batch I/O front-end of the NHW codec, plain C on top
 * of the batch C-ABI of include/nhw_cuda.h.
 *
 * The reference moves one image per process through one fopen/fread/fwrite each
 * (encoder/nhw_encoder.c:3046-3087 read_image_bmp, :3100-3220 write_compressed_file, decoder/nhw_decoder_cli.c:67-93) and
 * accepts exactly one input format: 512x512, 24 bits per pixel, BI_RGB (header_check, encoder/nhw_encoder.c:2980-2988;
 * a negative height is flipped, :3089-3093).  This layer replaces that for batches:
 *
 *   - image readers for 24 / 32 / 8-bit BMP (bottom-up or top-down) and binary PPM / PGM, of ANY size: an image is cut into
 *     512x512 tiles (edge pixels replicated into the padding), every tile is an ordinary NHW image;
 *   - a multi-image container (".nhwpack": tile streams back to back + an index at the end of the file) so that a batch is
 *     one file, not one file per image; every blob in it is byte for byte the .nhw file the reference encoder writes for
 *     that tile, and can be extracted and fed to the reference's nhw-dec;
 *   - manifest / directory driven encode and decode with pinned, double-buffered host staging: a reader (writer) thread
 *     fills (drains) one buffer while the GPU works on the other.
 *
 * A 512x512 24-bit bottom-up or top-down BMP goes through here to exactly the bytes read_image_bmp + encode_image +
 * write_compressed_file produce, and comes back as exactly the BMP file nhw-dec writes (tests/test_batchio_gpu.py).
 */
#ifndef NHW_BATCHIO_H
#define NHW_BATCHIO_H

#include <stddef.h>
#include <stdint.h>

#include "nhw_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NHW_TILE 512
#define NHW_IO_OK 0
#define NHW_IO_ERR_OPEN (-101)      /* cannot open / create a file */
#define NHW_IO_ERR_FORMAT (-102)    /* not a supported image file / not a .nhwpack */
#define NHW_IO_ERR_READ (-103)      /* short read, truncated file */
#define NHW_IO_ERR_WRITE (-104)
#define NHW_IO_ERR_NOMEM (-105)
#define NHW_IO_ERR_ARG (-106)
#define NHW_IO_ERR_CODEC (-107)     /* a tile failed in the codec: see nhw_batch_stats.first_bad_* */

/* ---- images on the host ----------------------------------------------------------------------------------------- */
/* pixels: height rows of width*3 bytes, B,G,R per pixel, BOTTOM-UP (row 0 = bottom line) and without row padding: for a
 * 512x512 image exactly the 786432 bytes read_image_bmp hands to the codec. */
typedef struct {
	uint32_t width, height;
	uint8_t *pixels;             /* malloc'd; nhw_image_free */
} nhw_image;

/* BMP (BITMAPINFOHEADER and later; 24-bit, 32-bit BI_RGB / BI_BITFIELDS with the usual masks, 8-bit paletted; bottom-up or
 * top-down) and binary PNM (P6, P5; maxval 255), recognised by content.  Returns NHW_IO_OK or an error. */
int nhw_image_load(const char *path, nhw_image *out);
int nhw_image_load_mem(const uint8_t *data, size_t len, nhw_image *out);
void nhw_image_free(nhw_image *im);
/* format: 0 = 24-bit bottom-up BMP (for 512x512 byte-identical to the reference decoder's output file), 1 = binary PPM */
int nhw_image_save(const char *path, const nhw_image *im, int format);

/* tiles: tiles_x * tiles_y tiles of NHW_PIX_BYTES, row-major from the TOP-left corner of the picture; pixels outside the
 * picture replicate the nearest edge pixel.  `tiles` must hold nhw_tile_count() * NHW_PIX_BYTES bytes. */
uint32_t nhw_tiles_x(uint32_t width);
uint32_t nhw_tiles_y(uint32_t height);
void nhw_image_to_tiles(const nhw_image *im, uint8_t *tiles);
/* the inverse (crops the padding); im->pixels must hold width*height*3 bytes */
void nhw_tiles_to_image(const uint8_t *tiles, nhw_image *im);

/* ---- the container ---------------------------------------------------------------------------------------------- */
/*   header  (32 B) : "NHWPACK1", u32 version = 1, u32 quality, 16 reserved bytes
 *   blobs          : the tiles' .nhw streams back to back
 *   index          : nhw_pack_image[n_images] | u64 blob_offset[n_tiles + 1] (absolute file offsets) | names
 *   trailer (32 B) : u64 index_offset, u64 n_images, u64 n_tiles, "NHWPKEND"
 * all integers little endian.  The index sits at the end so that a pack is written in one streaming pass. */
typedef struct {
	uint32_t width, height;      /* picture size */
	uint32_t tiles_x, tiles_y;
	uint64_t first_tile;         /* index of tile (0,0) among the pack's blobs */
	uint32_t name_off, name_len; /* into the names section (not NUL-terminated there) */
} nhw_pack_image;

typedef struct nhw_pack nhw_pack;   /* an open pack, reading */

int nhw_pack_open(const char *path, nhw_pack **out);
void nhw_pack_close(nhw_pack *p);
uint64_t nhw_pack_images(const nhw_pack *p);
uint64_t nhw_pack_tiles(const nhw_pack *p);
int nhw_pack_quality(const nhw_pack *p);
const nhw_pack_image *nhw_pack_image_info(const nhw_pack *p, uint64_t image);
/* name of image i copied into buf (NUL-terminated, truncated to cap); returns its full length */
size_t nhw_pack_image_name(const nhw_pack *p, uint64_t image, char *buf, size_t cap);
/* length of tile blob t; nhw_pack_read_tiles reads blobs [t0, t0+n) back to back into buf and fills offsets[0..n] */
uint64_t nhw_pack_tile_bytes(const nhw_pack *p, uint64_t tile);
int nhw_pack_read_tiles(nhw_pack *p, uint64_t t0, uint64_t n, uint8_t *buf, size_t cap, uint64_t *offsets);

/* ---- batch jobs -------------------------------------------------------------------------------------------------- */
typedef struct {
	uint64_t images, tiles;
	uint64_t bytes_in, bytes_out;         /* file bytes read / written */
	double seconds_total;
	double seconds_read, seconds_codec, seconds_write;   /* busy time of the three stages (they overlap) */
	int64_t first_bad_image;              /* -1: none */
	int32_t first_bad_status;
} nhw_batch_stats;

/* Encode the listed image files into one pack.  group_tiles: tiles per staging buffer (0 = 512; grows to the largest
 * image).  names: what is stored per image (NULL = the paths).  Stops at the first failing image. */
int nhw_batch_encode_files(nhw_ctx *ctx, const char *const *paths, const char *const *names, uint64_t n, int quality,
                           const char *pack_path, uint32_t group_tiles, nhw_batch_stats *stats);
/* one path per line; blank lines and lines starting with '#' are skipped */
int nhw_batch_encode_manifest(nhw_ctx *ctx, const char *manifest_path, int quality, const char *pack_path,
                              uint32_t group_tiles, nhw_batch_stats *stats);
/* every *.bmp / *.ppm / *.pgm of a directory, sorted by name */
int nhw_batch_encode_dir(nhw_ctx *ctx, const char *dir, int quality, const char *pack_path, uint32_t group_tiles,
                         nhw_batch_stats *stats);

/* Decode every image of a pack into out_dir/<basename of the stored name>.<bmp|ppm> (format as nhw_image_save). */
int nhw_batch_decode_pack(nhw_ctx *ctx, const char *pack_path, const char *out_dir, int format, uint32_t group_tiles,
                          nhw_batch_stats *stats);
/* Write every tile of a pack as an individual .nhw file (out_dir/<name>[.tY_X].nhw): what the reference's nhw-dec reads. */
int nhw_batch_extract_pack(const char *pack_path, const char *out_dir, nhw_batch_stats *stats);

const char *nhw_batchio_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NHW_BATCHIO_H */
