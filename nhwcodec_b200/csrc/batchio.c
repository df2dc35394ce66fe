/*
 * batchio.c -- libnhw_batchio.so: image readers / writers, 512x512 tiling, the .nhwpack container and the manifest driven
 * batch jobs declared in include/nhw_batchio.h.  Plain C (host side of the codec stays C); the codec work happens behind
 * nhw_encode_batch / nhw_decode_batch of libnhw_cuda.so.
 *
 * What it replaces in the reference: read_image_bmp + header_check (encoder/nhw_encoder.c:2960-3093: one format, one size,
 * one file per process), write_compressed_file's fopen/fwrite per image (:3100-3220) and the decoder CLI's
 * fopen/fwrite per image (decoder/nhw_decoder_cli.c:67-93).
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "../../include/nhw_batchio.h"

#define EXPORT __attribute__((visibility("default")))

static __thread char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}
EXPORT const char *nhw_batchio_last_error(void) { return g_err; }

static double now_s(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static uint32_t rd16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t rd32(const uint8_t *p) { return rd16(p) | (rd16(p + 2) << 16); }
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
static void wr32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void wr64(uint8_t *p, uint64_t v) { wr32(p, (uint32_t)v); wr32(p + 4, (uint32_t)(v >> 32)); }

/* ================================================================================================================ */
/* image readers                                                                                                    */
/* ================================================================================================================ */
static int load_bmp(const uint8_t *d, size_t len, nhw_image *out)
{
	if (len < 54) return fail(NHW_IO_ERR_FORMAT, "BMP shorter than its headers");
	const uint32_t data_off = rd32(d + 10), dib = rd32(d + 14);
	if (dib < 40) return fail(NHW_IO_ERR_FORMAT, "BMP with a %u-byte info header (OS/2 style) is not supported", dib);
	const int32_t w = (int32_t)rd32(d + 18), hs = (int32_t)rd32(d + 22);
	const uint32_t planes = rd16(d + 26), bpp = rd16(d + 28), compr = rd32(d + 30);
	uint32_t ncol = rd32(d + 46);
	const int top_down = hs < 0;
	const int64_t h = top_down ? -(int64_t)hs : hs;
	if (planes != 1 || w <= 0 || h <= 0 || w > 65536 || h > 65536) return fail(NHW_IO_ERR_FORMAT, "BMP with implausible geometry %dx%d", w, hs);
	if (!(bpp == 24 || bpp == 32 || bpp == 8)) return fail(NHW_IO_ERR_FORMAT, "BMP with %u bits per pixel is not supported", bpp);
	if (!(compr == 0 || (compr == 3 && bpp == 32))) return fail(NHW_IO_ERR_FORMAT, "compressed BMP (method %u) is not supported", compr);
	if (compr == 3) {   /* BI_BITFIELDS: only the layout that equals BI_RGB (B, G, R in bytes 0, 1, 2) */
		const size_t mo = dib >= 52 ? 54 : 14 + (size_t)dib;
		if (mo + 12 > len) return fail(NHW_IO_ERR_READ, "truncated BMP");
		if (rd32(d + mo) != 0x00ff0000u || rd32(d + mo + 4) != 0x0000ff00u || rd32(d + mo + 8) != 0x000000ffu)
			return fail(NHW_IO_ERR_FORMAT, "BMP bit-field masks other than 8-8-8 B,G,R are not supported");
	}
	const uint8_t *pal = NULL;
	if (bpp == 8) {
		if (ncol == 0 || ncol > 256) ncol = 256;
		pal = d + 14 + dib;
		if (14 + (size_t)dib + 4 * (size_t)ncol > len) return fail(NHW_IO_ERR_READ, "truncated BMP palette");
	}
	const size_t row_in = (((size_t)w * bpp + 31) / 32) * 4;
	if (data_off > len || (len - data_off) / row_in < (size_t)h) return fail(NHW_IO_ERR_READ, "truncated BMP: pixel data ends early");
	out->width = (uint32_t)w;
	out->height = (uint32_t)h;
	out->pixels = (uint8_t *)malloc((size_t)w * (size_t)h * 3);
	if (!out->pixels) return fail(NHW_IO_ERR_NOMEM, "out of memory");
	for (int64_t y = 0; y < h; y++) {
		/* file row y is picture row (top_down ? y from the top : y from the bottom); ours is bottom-up */
		const uint8_t *src = d + data_off + (size_t)y * row_in;
		uint8_t *dst = out->pixels + (size_t)(top_down ? h - 1 - y : y) * (size_t)w * 3;
		if (bpp == 24) memcpy(dst, src, (size_t)w * 3);
		else if (bpp == 32)
			for (int32_t x = 0; x < w; x++) { dst[3 * x] = src[4 * x]; dst[3 * x + 1] = src[4 * x + 1]; dst[3 * x + 2] = src[4 * x + 2]; }
		else
			for (int32_t x = 0; x < w; x++) {
				const uint8_t *e = pal + 4 * (size_t)(src[x] < ncol ? src[x] : 0);
				dst[3 * x] = e[0]; dst[3 * x + 1] = e[1]; dst[3 * x + 2] = e[2];
			}
	}
	return NHW_IO_OK;
}

static int pnm_int(const uint8_t *d, size_t len, size_t *pos, uint32_t *v)
{
	size_t p = *pos;
	for (;;) {   /* white space and comments */
		while (p < len && (d[p] == ' ' || d[p] == '\t' || d[p] == '\n' || d[p] == '\r')) p++;
		if (p < len && d[p] == '#') { while (p < len && d[p] != '\n') p++; continue; }
		break;
	}
	if (p >= len || d[p] < '0' || d[p] > '9') return -1;
	uint64_t x = 0;
	while (p < len && d[p] >= '0' && d[p] <= '9') { x = x * 10 + (uint64_t)(d[p] - '0'); if (x > 1u << 20) return -1; p++; }
	*v = (uint32_t)x;
	*pos = p;
	return 0;
}

static int load_pnm(const uint8_t *d, size_t len, nhw_image *out)
{
	const int grey = d[1] == '5';
	size_t pos = 2;
	uint32_t w, h, maxv;
	if (pnm_int(d, len, &pos, &w) || pnm_int(d, len, &pos, &h) || pnm_int(d, len, &pos, &maxv)) return fail(NHW_IO_ERR_FORMAT, "malformed PNM header");
	if (maxv != 255 || w == 0 || h == 0 || w > 65536 || h > 65536) return fail(NHW_IO_ERR_FORMAT, "PNM: only maxval 255 is supported (%ux%u, maxval %u)", w, h, maxv);
	pos++;   /* the single white-space byte that ends the header */
	const size_t bpp = grey ? 1 : 3;
	if (pos > len || (len - pos) / ((size_t)w * bpp) < h) return fail(NHW_IO_ERR_READ, "truncated PNM");
	out->width = w;
	out->height = h;
	out->pixels = (uint8_t *)malloc((size_t)w * h * 3);
	if (!out->pixels) return fail(NHW_IO_ERR_NOMEM, "out of memory");
	for (uint32_t y = 0; y < h; y++) {   /* PNM rows run top-down, R,G,B */
		const uint8_t *src = d + pos + (size_t)y * w * bpp;
		uint8_t *dst = out->pixels + (size_t)(h - 1 - y) * w * 3;
		if (grey) for (uint32_t x = 0; x < w; x++) dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = src[x];
		else for (uint32_t x = 0; x < w; x++) { dst[3 * x] = src[3 * x + 2]; dst[3 * x + 1] = src[3 * x + 1]; dst[3 * x + 2] = src[3 * x]; }
	}
	return NHW_IO_OK;
}

EXPORT int nhw_image_load_mem(const uint8_t *d, size_t len, nhw_image *out)
{
	if (!d || !out) return fail(NHW_IO_ERR_ARG, "null argument");
	memset(out, 0, sizeof *out);
	if (len >= 2 && d[0] == 'B' && d[1] == 'M') return load_bmp(d, len, out);
	if (len >= 2 && d[0] == 'P' && (d[1] == '6' || d[1] == '5')) return load_pnm(d, len, out);
	return fail(NHW_IO_ERR_FORMAT, "neither a BMP nor a binary PPM/PGM file");
}

static int read_whole(const char *path, uint8_t **data, size_t *len)
{
	FILE *f = fopen(path, "rb");
	if (!f) return fail(NHW_IO_ERR_OPEN, "cannot open %s: %s", path, strerror(errno));
	struct stat st;
	if (fstat(fileno(f), &st) != 0 || st.st_size < 0) { fclose(f); return fail(NHW_IO_ERR_READ, "cannot stat %s", path); }
	*len = (size_t)st.st_size;
	*data = (uint8_t *)malloc(*len ? *len : 1);
	if (!*data) { fclose(f); return fail(NHW_IO_ERR_NOMEM, "out of memory"); }
	if (*len && fread(*data, 1, *len, f) != *len) { fclose(f); free(*data); return fail(NHW_IO_ERR_READ, "short read on %s", path); }
	fclose(f);
	return NHW_IO_OK;
}

EXPORT int nhw_image_load(const char *path, nhw_image *out)
{
	uint8_t *d;
	size_t len;
	int rc = read_whole(path, &d, &len);
	if (rc) return rc;
	rc = nhw_image_load_mem(d, len, out);
	free(d);
	if (rc) { char msg[400]; snprintf(msg, sizeof msg, "%s", g_err); return fail(rc, "%s: %s", path, msg); }
	return rc;
}

EXPORT void nhw_image_free(nhw_image *im)
{
	if (im) { free(im->pixels); im->pixels = NULL; }
}

EXPORT int nhw_image_save(const char *path, const nhw_image *im, int format)
{
	if (!path || !im || !im->pixels) return fail(NHW_IO_ERR_ARG, "null argument");
	FILE *f = fopen(path, "wb");
	if (!f) return fail(NHW_IO_ERR_OPEN, "cannot create %s: %s", path, strerror(errno));
	const size_t w = im->width, h = im->height;
	int ok = 1;
	if (format == 0) {
		/* the reference decoder's fixed header generalised to w x h (decoder/nhw_decoder_cli.c:61-65: every field it leaves
		 * zero stays zero), rows bottom-up, padded to 4 bytes */
		const size_t row = (w * 3 + 3) & ~(size_t)3;
		uint8_t hd[54] = {0};
		hd[0] = 'B'; hd[1] = 'M';
		wr32(hd + 2, (uint32_t)(54 + row * h));
		wr32(hd + 10, 54);
		wr32(hd + 14, 40);
		wr32(hd + 18, (uint32_t)w);
		wr32(hd + 22, (uint32_t)h);
		hd[26] = 1;
		hd[28] = 24;
		wr32(hd + 34, (uint32_t)(row * h));
		ok = fwrite(hd, 54, 1, f) == 1;
		if (row == w * 3) ok = ok && fwrite(im->pixels, w * 3, h, f) == h;
		else {
			const uint8_t pad[4] = {0, 0, 0, 0};
			for (size_t y = 0; ok && y < h; y++)
				ok = fwrite(im->pixels + y * w * 3, w * 3, 1, f) == 1 && fwrite(pad, row - w * 3, 1, f) == 1;
		}
	} else {
		uint8_t *line = (uint8_t *)malloc(w * 3);
		ok = line != NULL && fprintf(f, "P6\n%zu %zu\n255\n", w, h) > 0;
		for (size_t y = 0; ok && y < h; y++) {
			const uint8_t *src = im->pixels + (h - 1 - y) * w * 3;
			for (size_t x = 0; x < w; x++) { line[3 * x] = src[3 * x + 2]; line[3 * x + 1] = src[3 * x + 1]; line[3 * x + 2] = src[3 * x]; }
			ok = fwrite(line, w * 3, 1, f) == 1;
		}
		free(line);
	}
	if (fclose(f) != 0) ok = 0;
	return ok ? NHW_IO_OK : fail(NHW_IO_ERR_WRITE, "write error on %s", path);
}

/* ================================================================================================================ */
/* tiling                                                                                                           */
/* ================================================================================================================ */
EXPORT uint32_t nhw_tiles_x(uint32_t width) { return (width + NHW_TILE - 1) / NHW_TILE; }
EXPORT uint32_t nhw_tiles_y(uint32_t height) { return (height + NHW_TILE - 1) / NHW_TILE; }

/* picture row `py` counted from the TOP is stored at bottom-up row height-1-py */
EXPORT void nhw_image_to_tiles(const nhw_image *im, uint8_t *tiles)
{
	const uint32_t w = im->width, h = im->height, tx_n = nhw_tiles_x(w), ty_n = nhw_tiles_y(h);
	for (uint32_t ty = 0; ty < ty_n; ty++)
		for (uint32_t tx = 0; tx < tx_n; tx++) {
			uint8_t *t = tiles + ((size_t)ty * tx_n + tx) * NHW_PIX_BYTES;
			const uint32_t x0 = tx * NHW_TILE, have = w - x0 < NHW_TILE ? w - x0 : NHW_TILE;
			for (uint32_t r = 0; r < NHW_TILE; r++) {           /* r = tile row from the top */
				uint32_t py = ty * NHW_TILE + r;
				if (py >= h) py = h - 1;
				const uint8_t *src = im->pixels + ((size_t)(h - 1 - py) * w + x0) * 3;
				uint8_t *dst = t + (size_t)(NHW_TILE - 1 - r) * NHW_TILE * 3;   /* tiles are bottom-up too */
				memcpy(dst, src, (size_t)have * 3);
				for (uint32_t x = have; x < NHW_TILE; x++) memcpy(dst + 3 * x, src + 3 * (size_t)(have - 1), 3);
			}
		}
}

EXPORT void nhw_tiles_to_image(const uint8_t *tiles, nhw_image *im)
{
	const uint32_t w = im->width, h = im->height, tx_n = nhw_tiles_x(w), ty_n = nhw_tiles_y(h);
	for (uint32_t ty = 0; ty < ty_n; ty++)
		for (uint32_t tx = 0; tx < tx_n; tx++) {
			const uint8_t *t = tiles + ((size_t)ty * tx_n + tx) * NHW_PIX_BYTES;
			const uint32_t x0 = tx * NHW_TILE, have = w - x0 < NHW_TILE ? w - x0 : NHW_TILE;
			for (uint32_t r = 0; r < NHW_TILE; r++) {
				const uint32_t py = ty * NHW_TILE + r;
				if (py >= h) break;
				memcpy(im->pixels + ((size_t)(h - 1 - py) * w + x0) * 3, t + (size_t)(NHW_TILE - 1 - r) * NHW_TILE * 3, (size_t)have * 3);
			}
		}
}

/* ================================================================================================================ */
/* the container                                                                                                    */
/* ================================================================================================================ */
#define PACK_HEADER 32
#define PACK_TRAILER 32
#define PACK_IMAGE_REC 32

struct nhw_pack {
	FILE *f;
	int quality;
	uint64_t n_images, n_tiles;
	nhw_pack_image *images;
	uint64_t *offsets;     /* n_tiles + 1 */
	char *names;
	uint64_t names_bytes;
};

/* writer state: blobs are appended as they come, the index is kept in memory and written by pack_finish */
typedef struct {
	FILE *f;
	uint64_t pos;
	uint64_t n_images, n_tiles, cap_images, off_cap;
	uint64_t planned;            /* tiles promised by the image records so far (blobs follow their records) */
	nhw_pack_image *images;
	uint64_t *offsets;
	char *names;
	uint64_t names_bytes, names_cap;
} pack_writer;

static int pack_begin(pack_writer *w, const char *path, int quality)
{
	memset(w, 0, sizeof *w);
	w->f = fopen(path, "wb");
	if (!w->f) return fail(NHW_IO_ERR_OPEN, "cannot create %s: %s", path, strerror(errno));
	uint8_t hd[PACK_HEADER] = {0};
	memcpy(hd, "NHWPACK1", 8);
	wr32(hd + 8, 1);
	wr32(hd + 12, (uint32_t)quality);
	if (fwrite(hd, PACK_HEADER, 1, w->f) != 1) return fail(NHW_IO_ERR_WRITE, "write error on %s", path);
	w->pos = PACK_HEADER;
	return NHW_IO_OK;
}

static int grow(void **p, uint64_t *cap, uint64_t need, size_t elem)
{
	if (need <= *cap) return 0;
	uint64_t c = *cap ? *cap : 64;
	while (c < need) c *= 2;
	void *q = realloc(*p, (size_t)c * elem);
	if (!q) return -1;
	*p = q;
	*cap = c;
	return 0;
}

static int pack_add_image(pack_writer *w, uint32_t width, uint32_t height, const char *name)
{
	const size_t nl = strlen(name);
	const uint64_t tiles = (uint64_t)nhw_tiles_x(width) * nhw_tiles_y(height);
	if (grow((void **)&w->images, &w->cap_images, w->n_images + 1, sizeof *w->images) ||
	    grow((void **)&w->offsets, &w->off_cap, w->planned + tiles + 1, sizeof *w->offsets) ||
	    grow((void **)&w->names, &w->names_cap, w->names_bytes + nl + 1, 1))
		return fail(NHW_IO_ERR_NOMEM, "out of memory");
	nhw_pack_image *im = &w->images[w->n_images++];
	im->width = width; im->height = height;
	im->tiles_x = nhw_tiles_x(width); im->tiles_y = nhw_tiles_y(height);
	im->first_tile = w->planned;
	w->planned += tiles;
	im->name_off = (uint32_t)w->names_bytes; im->name_len = (uint32_t)nl;
	memcpy(w->names + w->names_bytes, name, nl);
	w->names_bytes += nl;
	return NHW_IO_OK;
}

static int pack_add_blob(pack_writer *w, const uint8_t *blob, uint64_t len)
{
	w->offsets[w->n_tiles++] = w->pos;
	if (len && fwrite(blob, 1, (size_t)len, w->f) != len) return fail(NHW_IO_ERR_WRITE, "write error on the pack");
	w->pos += len;
	return NHW_IO_OK;
}

static int pack_finish(pack_writer *w, uint64_t *file_bytes)
{
	int ok = 1;
	if (w->f) {
		const uint64_t index_off = w->pos;
		uint8_t rec[PACK_IMAGE_REC], tr[PACK_TRAILER], o8[8];
		if (grow((void **)&w->offsets, &w->off_cap, w->n_tiles + 1, sizeof *w->offsets) == 0) w->offsets[w->n_tiles] = w->pos;
		else ok = 0;
		for (uint64_t i = 0; ok && i < w->n_images; i++) {
			const nhw_pack_image *im = &w->images[i];
			wr32(rec, im->width); wr32(rec + 4, im->height); wr32(rec + 8, im->tiles_x); wr32(rec + 12, im->tiles_y);
			wr64(rec + 16, im->first_tile); wr32(rec + 24, im->name_off); wr32(rec + 28, im->name_len);
			ok = fwrite(rec, PACK_IMAGE_REC, 1, w->f) == 1;
		}
		for (uint64_t t = 0; ok && t <= w->n_tiles; t++) { wr64(o8, w->offsets[t]); ok = fwrite(o8, 8, 1, w->f) == 1; }
		if (ok && w->names_bytes) ok = fwrite(w->names, 1, (size_t)w->names_bytes, w->f) == w->names_bytes;
		wr64(tr, index_off); wr64(tr + 8, w->n_images); wr64(tr + 16, w->n_tiles); memcpy(tr + 24, "NHWPKEND", 8);
		ok = ok && fwrite(tr, PACK_TRAILER, 1, w->f) == 1;
		if (file_bytes) *file_bytes = index_off + w->n_images * PACK_IMAGE_REC + (w->n_tiles + 1) * 8 + w->names_bytes + PACK_TRAILER;
		if (fclose(w->f) != 0) ok = 0;
		w->f = NULL;
	}
	free(w->images); free(w->offsets); free(w->names);
	w->images = NULL; w->offsets = NULL; w->names = NULL;
	return ok ? NHW_IO_OK : fail(NHW_IO_ERR_WRITE, "write error on the pack");
}

EXPORT int nhw_pack_open(const char *path, nhw_pack **out)
{
	if (!path || !out) return fail(NHW_IO_ERR_ARG, "null argument");
	*out = NULL;
	FILE *f = fopen(path, "rb");
	if (!f) return fail(NHW_IO_ERR_OPEN, "cannot open %s: %s", path, strerror(errno));
	uint8_t hd[PACK_HEADER], tr[PACK_TRAILER];
	struct stat st;
	if (fstat(fileno(f), &st) != 0 || st.st_size < PACK_HEADER + PACK_TRAILER || fread(hd, PACK_HEADER, 1, f) != 1 ||
	    memcmp(hd, "NHWPACK1", 8) != 0 || rd32(hd + 8) != 1) {
		fclose(f);
		return fail(NHW_IO_ERR_FORMAT, "%s is not a version-1 .nhwpack", path);
	}
	const uint64_t size = (uint64_t)st.st_size;
	if (fseeko(f, (off_t)(size - PACK_TRAILER), SEEK_SET) != 0 || fread(tr, PACK_TRAILER, 1, f) != 1 || memcmp(tr + 24, "NHWPKEND", 8) != 0) {
		fclose(f);
		return fail(NHW_IO_ERR_FORMAT, "%s: no trailer (truncated pack?)", path);
	}
	const uint64_t index_off = rd64(tr), ni = rd64(tr + 8), nt = rd64(tr + 16);
	const uint64_t fixed = ni * PACK_IMAGE_REC + (nt + 1) * 8;
	if (ni > (1ull << 40) || nt > (1ull << 40) || index_off < PACK_HEADER || index_off > size - PACK_TRAILER ||
	    fixed > size - PACK_TRAILER - index_off) {
		fclose(f);
		return fail(NHW_IO_ERR_FORMAT, "%s: index does not fit the file", path);
	}
	nhw_pack *p = (nhw_pack *)calloc(1, sizeof *p);
	const uint64_t nb = size - PACK_TRAILER - index_off - fixed;
	uint8_t *raw = (uint8_t *)malloc((size_t)(fixed + nb) + 1);
	if (!p || !raw) { free(p); free(raw); fclose(f); return fail(NHW_IO_ERR_NOMEM, "out of memory"); }
	p->images = (nhw_pack_image *)malloc((size_t)(ni ? ni : 1) * sizeof *p->images);
	p->offsets = (uint64_t *)malloc((size_t)(nt + 1) * sizeof *p->offsets);
	p->names = (char *)malloc((size_t)nb + 1);
	int ok = p->images && p->offsets && p->names && fseeko(f, (off_t)index_off, SEEK_SET) == 0 &&
	         fread(raw, 1, (size_t)(fixed + nb), f) == fixed + nb;
	if (ok) {
		for (uint64_t i = 0; i < ni; i++) {
			const uint8_t *r = raw + i * PACK_IMAGE_REC;
			nhw_pack_image *im = &p->images[i];
			im->width = rd32(r); im->height = rd32(r + 4); im->tiles_x = rd32(r + 8); im->tiles_y = rd32(r + 12);
			im->first_tile = rd64(r + 16); im->name_off = rd32(r + 24); im->name_len = rd32(r + 28);
			/* a record must describe what its own geometry implies and stay inside the tables */
			if (im->width == 0 || im->height == 0 || im->tiles_x != nhw_tiles_x(im->width) || im->tiles_y != nhw_tiles_y(im->height) ||
			    im->first_tile > nt || (uint64_t)im->tiles_x * im->tiles_y > nt - im->first_tile ||
			    (uint64_t)im->name_off + im->name_len > nb)
				ok = 0;
		}
		for (uint64_t t = 0; t <= nt; t++) {
			p->offsets[t] = rd64(raw + ni * PACK_IMAGE_REC + t * 8);
			if (p->offsets[t] < PACK_HEADER || p->offsets[t] > index_off || (t && p->offsets[t] < p->offsets[t - 1])) ok = 0;
		}
		memcpy(p->names, raw + fixed, (size_t)nb);
	}
	free(raw);
	if (!ok) {
		free(p->images); free(p->offsets); free(p->names); free(p);
		fclose(f);
		return fail(NHW_IO_ERR_FORMAT, "%s: inconsistent index", path);
	}
	p->f = f;
	p->quality = (int)rd32(hd + 12);
	p->n_images = ni;
	p->n_tiles = nt;
	p->names_bytes = nb;
	*out = p;
	return NHW_IO_OK;
}

EXPORT void nhw_pack_close(nhw_pack *p)
{
	if (!p) return;
	if (p->f) fclose(p->f);
	free(p->images); free(p->offsets); free(p->names); free(p);
}
EXPORT uint64_t nhw_pack_images(const nhw_pack *p) { return p ? p->n_images : 0; }
EXPORT uint64_t nhw_pack_tiles(const nhw_pack *p) { return p ? p->n_tiles : 0; }
EXPORT int nhw_pack_quality(const nhw_pack *p) { return p ? p->quality : -1; }
EXPORT const nhw_pack_image *nhw_pack_image_info(const nhw_pack *p, uint64_t i) { return p && i < p->n_images ? &p->images[i] : NULL; }
EXPORT size_t nhw_pack_image_name(const nhw_pack *p, uint64_t i, char *buf, size_t cap)
{
	if (!p || i >= p->n_images) return 0;
	const nhw_pack_image *im = &p->images[i];
	if (buf && cap) {
		const size_t n = im->name_len < cap - 1 ? im->name_len : cap - 1;
		memcpy(buf, p->names + im->name_off, n);
		buf[n] = 0;
	}
	return im->name_len;
}
EXPORT uint64_t nhw_pack_tile_bytes(const nhw_pack *p, uint64_t t) { return p && t < p->n_tiles ? p->offsets[t + 1] - p->offsets[t] : 0; }

EXPORT int nhw_pack_read_tiles(nhw_pack *p, uint64_t t0, uint64_t n, uint8_t *buf, size_t cap, uint64_t *offsets)
{
	if (!p || !buf || !offsets || t0 > p->n_tiles || n > p->n_tiles - t0) return fail(NHW_IO_ERR_ARG, "tile range outside the pack");
	const uint64_t a = p->offsets[t0], total = p->offsets[t0 + n] - a;
	if (total > cap) return fail(NHW_IO_ERR_ARG, "buffer too small for %llu tile bytes", (unsigned long long)total);
	if (fseeko(p->f, (off_t)a, SEEK_SET) != 0 || (total && fread(buf, 1, (size_t)total, p->f) != total)) return fail(NHW_IO_ERR_READ, "short read on the pack");
	for (uint64_t i = 0; i <= n; i++) offsets[i] = p->offsets[t0 + i] - a;
	return NHW_IO_OK;
}

/* ================================================================================================================ */
/* batch jobs: two staging slots, one helper thread                                                                  */
/* ================================================================================================================ */
typedef struct {
	uint32_t width, height;
	uint64_t index;              /* image number in the job */
} group_image;

typedef struct {
	uint8_t *tiles;              /* pinned, cap_tiles * NHW_PIX_BYTES */
	uint64_t n_tiles, n_images;
	group_image *images;         /* cap_tiles entries (an image has at least one tile) */
	int state;                   /* 0 = free (owned by the producer), 1 = full (owned by the consumer) */
	int last;                    /* nothing follows this group */
	int rc;                      /* producer's verdict on this group */
	char err[512];
} slot_t;

typedef struct {
	pthread_mutex_t mu;
	pthread_cond_t cv;
	slot_t slot[2];
	uint32_t cap_tiles;
	int abort;
	double busy;                 /* helper thread's busy seconds */
	uint64_t bytes;              /* helper thread's file bytes */
	/* encode: the file list;  decode: the pack + output directory */
	const char *const *paths;
	uint64_t n_paths;
	nhw_pack *pack;
	const char *out_dir;
	int format;
} job_t;

static void slot_wait(job_t *j, slot_t *s, int want)
{
	pthread_mutex_lock(&j->mu);
	while (s->state != want && !j->abort) pthread_cond_wait(&j->cv, &j->mu);
	pthread_mutex_unlock(&j->mu);
}
static void slot_set(job_t *j, slot_t *s, int state)
{
	pthread_mutex_lock(&j->mu);
	s->state = state;
	pthread_cond_broadcast(&j->cv);
	pthread_mutex_unlock(&j->mu);
}
static void job_abort(job_t *j)
{
	pthread_mutex_lock(&j->mu);
	j->abort = 1;
	pthread_cond_broadcast(&j->cv);
	pthread_mutex_unlock(&j->mu);
}

static int job_init(job_t *j, uint32_t cap_tiles)
{
	memset(j, 0, sizeof *j);
	pthread_mutex_init(&j->mu, NULL);
	pthread_cond_init(&j->cv, NULL);
	j->cap_tiles = cap_tiles;
	for (int s = 0; s < 2; s++) {
		j->slot[s].tiles = (uint8_t *)nhw_host_alloc((size_t)cap_tiles * NHW_PIX_BYTES);
		j->slot[s].images = (group_image *)malloc((size_t)cap_tiles * sizeof(group_image));
		if (!j->slot[s].tiles || !j->slot[s].images) return fail(NHW_IO_ERR_NOMEM, "cannot allocate %u-tile staging buffers", cap_tiles);
	}
	return NHW_IO_OK;
}
static void job_free(job_t *j)
{
	for (int s = 0; s < 2; s++) { nhw_host_free(j->slot[s].tiles); free(j->slot[s].images); }
	pthread_mutex_destroy(&j->mu);
	pthread_cond_destroy(&j->cv);
}

/* ---- encode: the helper thread reads and tiles images into the slots --------------------------------------------- */
static void *encode_reader(void *arg)
{
	job_t *j = (job_t *)arg;
	nhw_image pending;
	int have_pending = 0;
	uint64_t next = 0;
	for (int s = 0;; s ^= 1) {
		slot_t *sl = &j->slot[s];
		slot_wait(j, sl, 0);
		if (j->abort) break;
		const double t0 = now_s();
		sl->n_tiles = sl->n_images = 0;
		sl->rc = NHW_IO_OK;
		sl->last = 0;
		for (;;) {
			if (!have_pending) {
				if (next >= j->n_paths) { sl->last = 1; break; }
				int rc = nhw_image_load(j->paths[next], &pending);
				if (rc == NHW_IO_OK) {
					struct stat st;
					if (stat(j->paths[next], &st) == 0) j->bytes += (uint64_t)st.st_size;
				} else {
					sl->rc = rc;
					snprintf(sl->err, sizeof sl->err, "%s", g_err);
					sl->last = 1;
					break;
				}
				have_pending = 1;
			}
			const uint64_t t = (uint64_t)nhw_tiles_x(pending.width) * nhw_tiles_y(pending.height);
			if (t > j->cap_tiles) {
				sl->rc = NHW_IO_ERR_ARG;
				snprintf(sl->err, sizeof sl->err, "%s needs %llu tiles, the staging buffers hold %u (raise group_tiles)", j->paths[next],
				         (unsigned long long)t, j->cap_tiles);
				nhw_image_free(&pending);
				have_pending = 0;
				sl->last = 1;
				break;
			}
			if (sl->n_tiles + t > j->cap_tiles) break;          /* goes into the next group */
			nhw_image_to_tiles(&pending, sl->tiles + (size_t)sl->n_tiles * NHW_PIX_BYTES);
			sl->images[sl->n_images].width = pending.width;
			sl->images[sl->n_images].height = pending.height;
			sl->images[sl->n_images].index = next;
			sl->n_images++;
			sl->n_tiles += t;
			nhw_image_free(&pending);
			have_pending = 0;
			next++;
		}
		j->busy += now_s() - t0;
		const int last = sl->last;
		slot_set(j, sl, 1);
		if (last) break;
	}
	if (have_pending) nhw_image_free(&pending);
	return NULL;
}

/* encode tiles [a, b) of a slot; on "output buffer too small" split the range (a single tile always fits) */
static int encode_range(nhw_ctx *ctx, const uint8_t *tiles, uint64_t a, uint64_t b, int quality, uint8_t *out, size_t out_cap,
                        uint64_t *offs, int32_t *status, pack_writer *w, int64_t *bad_tile, int32_t *bad_status)
{
	int rc = nhw_encode_batch(ctx, tiles + (size_t)a * NHW_PIX_BYTES, (int)(b - a), quality, out, out_cap, offs, status);
	if (rc == NHW_ERR_ARG && b - a > 1) {
		const uint64_t m = a + (b - a) / 2;
		rc = encode_range(ctx, tiles, a, m, quality, out, out_cap, offs, status, w, bad_tile, bad_status);
		if (rc) return rc;
		return encode_range(ctx, tiles, m, b, quality, out, out_cap, offs, status, w, bad_tile, bad_status);
	}
	if (rc != NHW_OK) return fail(NHW_IO_ERR_CODEC, "nhw_encode_batch: %d (%s)", rc, nhw_last_error());
	for (uint64_t i = 0; i < b - a; i++) {
		if (status[i] != NHW_OK) {
			*bad_tile = (int64_t)(a + i);
			*bad_status = status[i];
			return fail(NHW_IO_ERR_CODEC, "tile %llu of the group: codec status %d", (unsigned long long)(a + i), (int)status[i]);
		}
		if ((rc = pack_add_blob(w, out + offs[i], offs[i + 1] - offs[i])) != NHW_IO_OK) return rc;
	}
	return NHW_IO_OK;
}

EXPORT int nhw_batch_encode_files(nhw_ctx *ctx, const char *const *paths, const char *const *names, uint64_t n, int quality,
                                  const char *pack_path, uint32_t group_tiles, nhw_batch_stats *stats)
{
	if (!ctx || (!paths && n) || !pack_path) return fail(NHW_IO_ERR_ARG, "null argument");
	nhw_batch_stats st;
	memset(&st, 0, sizeof st);
	st.first_bad_image = -1;
	const double t_start = now_s();
	const uint32_t cap = group_tiles ? group_tiles : 512;
	job_t j;
	pack_writer w;
	int rc = job_init(&j, cap);
	if (rc) { job_free(&j); return rc; }
	j.paths = paths;
	j.n_paths = n;
	const size_t out_cap = (size_t)cap * (NHW_MAX_STREAM_BYTES / 2) + NHW_MAX_STREAM_BYTES;
	uint8_t *out = (uint8_t *)nhw_host_alloc(out_cap);
	uint64_t *offs = (uint64_t *)malloc(((size_t)cap + 1) * sizeof *offs);
	int32_t *status = (int32_t *)malloc((size_t)cap * sizeof *status);
	if (!out || !offs || !status) rc = fail(NHW_IO_ERR_NOMEM, "out of memory");
	if (!rc) rc = pack_begin(&w, pack_path, quality);
	else memset(&w, 0, sizeof w);
	pthread_t th;
	int started = 0;
	if (!rc) {
		if (pthread_create(&th, NULL, encode_reader, &j) != 0) rc = fail(NHW_IO_ERR_NOMEM, "cannot start the reader thread");
		else started = 1;
	}
	for (int s = 0; !rc; s ^= 1) {
		slot_t *sl = &j.slot[s];
		slot_wait(&j, sl, 1);
		/* the images that made it into the group are encoded even when the reader stopped on the one after them */
		double t0 = now_s();
		int64_t bad_tile = -1;
		int32_t bad_status = 0;
		uint64_t tile0 = 0;
		for (uint64_t i = 0; !rc && i < sl->n_images; i++) {
			const uint64_t idx = sl->images[i].index;
			rc = pack_add_image(&w, sl->images[i].width, sl->images[i].height, names && names[idx] ? names[idx] : paths[idx]);
		}
		if (!rc && sl->n_tiles) {
			rc = encode_range(ctx, sl->tiles, 0, sl->n_tiles, quality, out, out_cap, offs, status, &w, &bad_tile, &bad_status);
			if (rc == NHW_IO_ERR_CODEC && bad_tile >= 0) {
				for (uint64_t i = 0; i < sl->n_images; i++) {
					const uint64_t t = (uint64_t)nhw_tiles_x(sl->images[i].width) * nhw_tiles_y(sl->images[i].height);
					if ((uint64_t)bad_tile < tile0 + t) { st.first_bad_image = (int64_t)sl->images[i].index; break; }
					tile0 += t;
				}
				st.first_bad_status = bad_status;
			}
		}
		st.seconds_codec += now_s() - t0;
		if (!rc) { st.images += sl->n_images; st.tiles += sl->n_tiles; }
		if (!rc && sl->rc != NHW_IO_OK) {
			rc = fail(sl->rc, "%s", sl->err);
			st.first_bad_image = (int64_t)st.images;
			st.first_bad_status = sl->rc;
		}
		const int last = sl->last;
		if (rc) break;
		slot_set(&j, sl, 0);
		if (last) break;
	}
	if (rc) job_abort(&j);
	if (started) pthread_join(th, NULL);
	char keep[512];
	snprintf(keep, sizeof keep, "%s", g_err);
	const int frc = pack_finish(&w, &st.bytes_out);
	if (rc) snprintf(g_err, sizeof g_err, "%s", keep);
	else rc = frc;
	st.bytes_in = j.bytes;
	st.seconds_read = j.busy;
	st.seconds_total = now_s() - t_start;
	nhw_host_free(out);
	free(offs);
	free(status);
	job_free(&j);
	if (stats) *stats = st;
	return rc;
}

static int cmp_str(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

static void free_list(char **v, uint64_t n)
{
	for (uint64_t i = 0; i < n; i++) free(v[i]);
	free(v);
}

EXPORT int nhw_batch_encode_manifest(nhw_ctx *ctx, const char *manifest_path, int quality, const char *pack_path,
                                     uint32_t group_tiles, nhw_batch_stats *stats)
{
	FILE *f = fopen(manifest_path, "r");
	if (!f) return fail(NHW_IO_ERR_OPEN, "cannot open %s: %s", manifest_path, strerror(errno));
	char **v = NULL, *line = NULL;
	uint64_t n = 0, cap = 0;
	size_t lcap = 0;
	ssize_t len;
	int rc = NHW_IO_OK;
	while ((len = getline(&line, &lcap, f)) >= 0) {
		while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r' || line[len - 1] == ' ' || line[len - 1] == '\t')) line[--len] = 0;
		char *s = line;
		while (*s == ' ' || *s == '\t') s++;
		if (*s == 0 || *s == '#') continue;
		if (grow((void **)&v, &cap, n + 1, sizeof *v) || !(v[n] = strdup(s))) { rc = fail(NHW_IO_ERR_NOMEM, "out of memory"); break; }
		n++;
	}
	free(line);
	fclose(f);
	if (!rc) rc = nhw_batch_encode_files(ctx, (const char *const *)v, NULL, n, quality, pack_path, group_tiles, stats);
	free_list(v, n);
	return rc;
}

static int has_image_ext(const char *name)
{
	const char *dot = strrchr(name, '.');
	return dot && (!strcasecmp(dot, ".bmp") || !strcasecmp(dot, ".ppm") || !strcasecmp(dot, ".pgm"));
}

EXPORT int nhw_batch_encode_dir(nhw_ctx *ctx, const char *dir, int quality, const char *pack_path, uint32_t group_tiles,
                                nhw_batch_stats *stats)
{
	DIR *d = opendir(dir);
	if (!d) return fail(NHW_IO_ERR_OPEN, "cannot open directory %s: %s", dir, strerror(errno));
	char **v = NULL;
	uint64_t n = 0, cap = 0;
	int rc = NHW_IO_OK;
	struct dirent *e;
	while ((e = readdir(d)) != NULL) {
		if (!has_image_ext(e->d_name)) continue;
		char *p = NULL;
		if (asprintf(&p, "%s/%s", dir, e->d_name) < 0 || grow((void **)&v, &cap, n + 1, sizeof *v)) { free(p); rc = fail(NHW_IO_ERR_NOMEM, "out of memory"); break; }
		v[n++] = p;
	}
	closedir(d);
	if (!rc) {
		qsort(v, (size_t)n, sizeof *v, cmp_str);
		rc = nhw_batch_encode_files(ctx, (const char *const *)v, NULL, n, quality, pack_path, group_tiles, stats);
	}
	free_list(v, n);
	return rc;
}

/* ---- decode: the helper thread reassembles and writes the images of a finished group ------------------------------ */
static void out_name(const nhw_pack *p, uint64_t image, const char *dir, const char *suffix, const char *ext, char *buf, size_t cap)
{
	char name[1024];
	nhw_pack_image_name(p, image, name, sizeof name);
	const char *base = strrchr(name, '/');
	base = base ? base + 1 : name;
	char stem[1024];
	snprintf(stem, sizeof stem, "%s", *base ? base : "image");
	char *dot = strrchr(stem, '.');
	if (dot && dot != stem) *dot = 0;
	snprintf(buf, cap, "%s/%s%s.%s", dir, stem, suffix, ext);
}

static void *decode_writer(void *arg)
{
	job_t *j = (job_t *)arg;
	for (int s = 0;; s ^= 1) {
		slot_t *sl = &j->slot[s];
		slot_wait(j, sl, 1);
		if (j->abort) break;
		const double t0 = now_s();
		uint64_t tile = 0;
		sl->rc = NHW_IO_OK;
		for (uint64_t i = 0; i < sl->n_images && sl->rc == NHW_IO_OK; i++) {
			nhw_image im;
			im.width = sl->images[i].width;
			im.height = sl->images[i].height;
			const uint64_t t = (uint64_t)nhw_tiles_x(im.width) * nhw_tiles_y(im.height);
			char path[2048];
			out_name(j->pack, sl->images[i].index, j->out_dir, "", j->format == 0 ? "bmp" : "ppm", path, sizeof path);
			if (t == 1 && im.width == NHW_TILE && im.height == NHW_TILE) {
				im.pixels = sl->tiles + (size_t)tile * NHW_PIX_BYTES;        /* the tile is the image */
				sl->rc = nhw_image_save(path, &im, j->format);
			} else {
				im.pixels = (uint8_t *)malloc((size_t)im.width * im.height * 3);
				if (!im.pixels) sl->rc = fail(NHW_IO_ERR_NOMEM, "out of memory");
				else {
					nhw_tiles_to_image(sl->tiles + (size_t)tile * NHW_PIX_BYTES, &im);
					sl->rc = nhw_image_save(path, &im, j->format);
					free(im.pixels);
				}
			}
			if (sl->rc == NHW_IO_OK) {
				struct stat st;
				if (stat(path, &st) == 0) j->bytes += (uint64_t)st.st_size;
			} else snprintf(sl->err, sizeof sl->err, "%s", g_err);
			tile += t;
		}
		j->busy += now_s() - t0;
		const int last = sl->last, bad = sl->rc != NHW_IO_OK;
		slot_set(j, sl, 0);
		if (last || bad) break;
	}
	return NULL;
}

EXPORT int nhw_batch_decode_pack(nhw_ctx *ctx, const char *pack_path, const char *out_dir, int format, uint32_t group_tiles,
                                 nhw_batch_stats *stats)
{
	if (!ctx || !pack_path || !out_dir) return fail(NHW_IO_ERR_ARG, "null argument");
	nhw_batch_stats st;
	memset(&st, 0, sizeof st);
	st.first_bad_image = -1;
	const double t_start = now_s();
	nhw_pack *p;
	int rc = nhw_pack_open(pack_path, &p);
	if (rc) return rc;
	uint32_t cap = group_tiles ? group_tiles : 512;
	for (uint64_t i = 0; i < p->n_images; i++) {
		const uint64_t t = (uint64_t)p->images[i].tiles_x * p->images[i].tiles_y;
		if (t > cap) cap = (uint32_t)t;
	}
	mkdir(out_dir, 0777);
	job_t j;
	rc = job_init(&j, cap);
	j.pack = p;
	j.out_dir = out_dir;
	j.format = format;
	const size_t in_cap = (size_t)cap * (NHW_MAX_STREAM_BYTES / 2) + NHW_MAX_STREAM_BYTES;
	uint8_t *in = rc ? NULL : (uint8_t *)nhw_host_alloc(in_cap);
	uint64_t *offs = (uint64_t *)malloc(((size_t)cap + 1) * sizeof *offs);
	int32_t *status = (int32_t *)malloc((size_t)cap * sizeof *status);
	if (!rc && (!in || !offs || !status)) rc = fail(NHW_IO_ERR_NOMEM, "out of memory");
	pthread_t th;
	int started = 0;
	if (!rc) {
		if (pthread_create(&th, NULL, decode_writer, &j) != 0) rc = fail(NHW_IO_ERR_NOMEM, "cannot start the writer thread");
		else started = 1;
	}
	uint64_t image = 0;
	for (int s = 0; !rc; s ^= 1) {
		slot_t *sl = &j.slot[s];
		slot_wait(&j, sl, 0);
		if (sl->rc != NHW_IO_OK) { rc = fail(sl->rc, "%s", sl->err); break; }    /* the writer failed on this slot's previous group */
		sl->n_images = sl->n_tiles = 0;
		uint64_t bytes = 0;
		const uint64_t t_first = image < p->n_images ? p->images[image].first_tile : 0;
		while (image < p->n_images) {
			const nhw_pack_image *im = &p->images[image];
			const uint64_t t = (uint64_t)im->tiles_x * im->tiles_y;
			if (im->first_tile != t_first + sl->n_tiles) break;                  /* groups are runs of consecutive tiles */
			const uint64_t b = p->offsets[im->first_tile + t] - p->offsets[im->first_tile];
			if (sl->n_tiles && (sl->n_tiles + t > cap || bytes + b > in_cap)) break;
			sl->images[sl->n_images].width = im->width;
			sl->images[sl->n_images].height = im->height;
			sl->images[sl->n_images].index = image;
			sl->n_images++;
			sl->n_tiles += t;
			bytes += b;
			image++;
		}
		sl->last = image >= p->n_images;
		if (sl->n_tiles) {
			double t0 = now_s();
			rc = bytes > in_cap ? fail(NHW_IO_ERR_ARG, "an image's tile streams exceed the staging buffer")
			                    : nhw_pack_read_tiles(p, t_first, sl->n_tiles, in, in_cap, offs);
			st.seconds_read += now_s() - t0;
			st.bytes_in += bytes;
			if (!rc) {
				t0 = now_s();
				const int crc = nhw_decode_batch(ctx, in, offs, (int)sl->n_tiles, sl->tiles, status);
				st.seconds_codec += now_s() - t0;
				if (crc != NHW_OK) rc = fail(NHW_IO_ERR_CODEC, "nhw_decode_batch: %d (%s)", crc, nhw_last_error());
				for (uint64_t i = 0, tile = 0; !rc && i < sl->n_images; i++) {
					const uint64_t t = (uint64_t)nhw_tiles_x(sl->images[i].width) * nhw_tiles_y(sl->images[i].height);
					for (uint64_t k = 0; k < t; k++)
						if (status[tile + k] != NHW_OK) {
							st.first_bad_image = (int64_t)sl->images[i].index;
							st.first_bad_status = status[tile + k];
							rc = fail(NHW_IO_ERR_CODEC, "image %llu, tile %llu: codec status %d", (unsigned long long)sl->images[i].index,
							          (unsigned long long)k, (int)status[tile + k]);
							break;
						}
					tile += t;
				}
			}
		}
		if (rc) break;
		st.images += sl->n_images;
		st.tiles += sl->n_tiles;
		const int last = sl->last;
		slot_set(&j, sl, 1);
		if (last) break;
	}
	if (rc) job_abort(&j);
	if (started) pthread_join(th, NULL);
	for (int s = 0; !rc && s < 2; s++)
		if (j.slot[s].rc != NHW_IO_OK) rc = fail(j.slot[s].rc, "%s", j.slot[s].err);
	st.bytes_out = j.bytes;
	st.seconds_write = j.busy;
	st.seconds_total = now_s() - t_start;
	nhw_host_free(in);
	free(offs);
	free(status);
	job_free(&j);
	nhw_pack_close(p);
	if (stats) *stats = st;
	return rc;
}

EXPORT int nhw_batch_extract_pack(const char *pack_path, const char *out_dir, nhw_batch_stats *stats)
{
	nhw_batch_stats st;
	memset(&st, 0, sizeof st);
	st.first_bad_image = -1;
	const double t_start = now_s();
	nhw_pack *p;
	int rc = nhw_pack_open(pack_path, &p);
	if (rc) return rc;
	mkdir(out_dir, 0777);
	uint8_t *buf = (uint8_t *)malloc(NHW_MAX_STREAM_BYTES);
	if (!buf) rc = fail(NHW_IO_ERR_NOMEM, "out of memory");
	for (uint64_t i = 0; !rc && i < p->n_images; i++) {
		const nhw_pack_image *im = &p->images[i];
		for (uint32_t ty = 0; !rc && ty < im->tiles_y; ty++)
			for (uint32_t tx = 0; !rc && tx < im->tiles_x; tx++) {
				const uint64_t t = im->first_tile + (uint64_t)ty * im->tiles_x + tx;
				uint64_t offs[2];
				char suffix[64] = "", path[2048];
				if (im->tiles_x * im->tiles_y > 1) snprintf(suffix, sizeof suffix, ".t%u_%u", ty, tx);
				out_name(p, i, out_dir, suffix, "nhw", path, sizeof path);
				if (nhw_pack_tile_bytes(p, t) > NHW_MAX_STREAM_BYTES) { rc = fail(NHW_IO_ERR_FORMAT, "tile %llu is larger than any .nhw stream", (unsigned long long)t); break; }
				if ((rc = nhw_pack_read_tiles(p, t, 1, buf, NHW_MAX_STREAM_BYTES, offs)) != 0) break;
				FILE *f = fopen(path, "wb");
				if (!f || (offs[1] && fwrite(buf, 1, (size_t)offs[1], f) != offs[1])) rc = fail(NHW_IO_ERR_WRITE, "cannot write %s", path);
				if (f) fclose(f);
				st.bytes_in += offs[1];
				st.bytes_out += offs[1];
				st.tiles++;
			}
		st.images++;
	}
	free(buf);
	nhw_pack_close(p);
	st.seconds_total = now_s() - t_start;
	if (stats) *stats = st;
	return rc;
}
