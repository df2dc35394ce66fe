/*
 * cli/nhw_batch_cli.c -- `nhw-batch`: directory / manifest driven batch encode and decode through the multi-image
 * container of include/nhw_batchio.h.  Same quality flag as nhw-enc (-q<1..23>, default 20).
 *
 *   nhw-batch enc [-q<n>] [-g<tiles>] -o out.nhwpack (-m manifest.txt | -d directory | image files ...)
 *   nhw-batch dec [-g<tiles>] [-p] in.nhwpack outdir        decode every image to outdir/<name>.bmp (-p: .ppm)
 *   nhw-batch extract in.nhwpack outdir                     write every tile as an ordinary .nhw file (no GPU needed)
 *   nhw-batch list in.nhwpack
 *
 * Inputs: 24/32/8-bit BMP (bottom-up or top-down) and binary PPM/PGM of any size (cut into 512x512 tiles).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/nhw_batchio.h"

static void usage(void)
{
	printf("usage: nhw-batch enc [-q<quality>] [-g<tiles>] -o out.nhwpack (-m manifest | -d dir | files...)\n"
	       "       nhw-batch dec [-g<tiles>] [-p] in.nhwpack outdir\n"
	       "       nhw-batch extract in.nhwpack outdir\n"
	       "       nhw-batch list in.nhwpack\n");
}

static void report(const char *what, const nhw_batch_stats *s)
{
	const double mpix = (double)s->tiles * 0.262144;
	fprintf(stderr, "%s: %llu images, %llu tiles, %.1f MB in, %.1f MB out, %.3f s (read %.3f, codec %.3f, write %.3f) = %.1f MPix/s\n", what,
	        (unsigned long long)s->images, (unsigned long long)s->tiles, s->bytes_in / 1e6, s->bytes_out / 1e6, s->seconds_total,
	        s->seconds_read, s->seconds_codec, s->seconds_write, s->seconds_total > 0 ? mpix / s->seconds_total : 0.0);
}

static nhw_ctx *make_ctx(int batch)
{
	nhw_ctx *ctx = NULL;
	const char *dev = getenv("NHW_CUDA_DEVICE");
	const int rc = nhw_create(dev ? atoi(dev) : 0, batch, &ctx);
	if (rc != NHW_OK) {
		fprintf(stderr, "nhw: cannot create CUDA codec context (%d): %s\n", rc, nhw_last_error());
		exit(1);
	}
	return ctx;
}

int main(int argc, char **argv)
{
	if (argc < 2) { usage(); return 0; }
	const char *cmd = argv[1];
	int quality = 20, ppm = 0, group = 512, a = 2;
	const char *out = NULL, *manifest = NULL, *dir = NULL;
	nhw_batch_stats st;
	for (; a < argc && argv[a][0] == '-' && argv[a][1]; a++) {
		const char *o = argv[a];
		if (o[1] == 'q') {
			if (o[2] < '0' || o[2] > '9') { printf("invalid quality='%s'\n", o + 2); return 1; }
			quality = atoi(o + 2);
			if (quality < 0 || quality > 23) { printf("quality=%d out of range\n", quality); return 1; }
		} else if (o[1] == 'g') group = atoi(o + 2) > 0 ? atoi(o + 2) : 512;
		else if (o[1] == 'p') ppm = 1;
		else if (o[1] == 'o' && a + 1 < argc) out = argv[++a];
		else if (o[1] == 'm' && a + 1 < argc) manifest = argv[++a];
		else if (o[1] == 'd' && a + 1 < argc) dir = argv[++a];
		else if (o[1] == 'h') { usage(); return 0; }
		else { fprintf(stderr, "Unknown option '%s'\n", o); return 1; }
	}
	int rc;
	if (!strcmp(cmd, "enc")) {
		if (!out || (!manifest && !dir && a >= argc)) { usage(); return 1; }
		nhw_ctx *ctx = make_ctx(group < 4096 ? group : 4096);
		if (manifest) rc = nhw_batch_encode_manifest(ctx, manifest, quality, out, (uint32_t)group, &st);
		else if (dir) rc = nhw_batch_encode_dir(ctx, dir, quality, out, (uint32_t)group, &st);
		else rc = nhw_batch_encode_files(ctx, (const char *const *)(argv + a), NULL, (uint64_t)(argc - a), quality, out, (uint32_t)group, &st);
		nhw_destroy(ctx);
		report("enc", &st);
	} else if (!strcmp(cmd, "dec")) {
		if (a + 2 > argc) { usage(); return 1; }
		nhw_ctx *ctx = make_ctx(group < 4096 ? group : 4096);
		rc = nhw_batch_decode_pack(ctx, argv[a], argv[a + 1], ppm, (uint32_t)group, &st);
		nhw_destroy(ctx);
		report("dec", &st);
	} else if (!strcmp(cmd, "extract")) {
		if (a + 2 > argc) { usage(); return 1; }
		rc = nhw_batch_extract_pack(argv[a], argv[a + 1], &st);
		report("extract", &st);
	} else if (!strcmp(cmd, "list")) {
		if (a + 1 > argc) { usage(); return 1; }
		nhw_pack *p;
		rc = nhw_pack_open(argv[a], &p);
		if (rc == NHW_IO_OK) {
			printf("%llu images, %llu tiles, quality %d\n", (unsigned long long)nhw_pack_images(p), (unsigned long long)nhw_pack_tiles(p), nhw_pack_quality(p));
			for (uint64_t i = 0; i < nhw_pack_images(p); i++) {
				char name[1024];
				const nhw_pack_image *im = nhw_pack_image_info(p, i);
				uint64_t bytes = 0;
				for (uint64_t t = 0; t < (uint64_t)im->tiles_x * im->tiles_y; t++) bytes += nhw_pack_tile_bytes(p, im->first_tile + t);
				nhw_pack_image_name(p, i, name, sizeof name);
				printf("%6llu  %5ux%-5u  %ux%u tiles  %9llu B  %s\n", (unsigned long long)i, im->width, im->height, im->tiles_x, im->tiles_y,
				       (unsigned long long)bytes, name);
			}
			nhw_pack_close(p);
		}
	} else { usage(); return 1; }
	if (rc != NHW_IO_OK) { fprintf(stderr, "nhw-batch: %s (%d)\n", nhw_batchio_last_error(), rc); return 1; }
	return 0;
}
