/*
 * cli/nhw_enc_cli.c -- `nhw-enc`, same flags and exit codes as the reference CLI
 * (encoder/nhw_encoder_cli.c:88-186): nhw-enc [-hV] [-f] [-q<1..23>] in.bmp out.nhw
 * Host code stays C; the work happens behind read_image_bmp / encode_image /
 * write_compressed_file (libnhw_compat -> libnhw_cuda).  The reference's own
 * nhw_encoder_cli.c links against the same two libraries unchanged.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/nhw_compat.h"

#define PROGRAM "nhw-enc"
#define VERSION "0.3.0-b200"

static void usage(void)
{
	printf("usage: %s [-hV] [-f] [-q<quality>] <image.bmp> <image.nhw>\n", PROGRAM);
	printf("  512x512 24-bit uncompressed BMP in, .nhw out; quality 1 (lowest) .. 23 (highest), default 20\n");
	printf("  -f  overwrite an existing output file   -h  this text   -V  version\n");
}

int main(int argc, char **argv)
{
	image_buffer im;
	encode_state enc;
	int quality = 20, overwrite = 0, i;
	char *in, *out;
	while (argc > 1 && argv[1][0] == '-') {
		for (i = 1; argv[1][i] != '\0'; i++) {
			const char c = argv[1][i];
			if (c >= '0' && c <= '9') continue;          /* digits belong to a preceding -q */
			if (c == 'q') {
				const char n = argv[1][i + 1];
				if (n < '0' || n > '9') { printf("invalid quality='%s'\n", &argv[1][i + 1]); exit(1); }
				quality = atoi(&argv[1][i + 1]);
				if (quality < 0 || quality > 23) { printf("quality=%d out of range\n", quality); exit(1); }
			} else if (c == 'f') overwrite = 1;
			else if (c == 'h') { usage(); exit(0); }
			else if (c == 'V') { printf("%s %s\n", PROGRAM, VERSION); exit(0); }
			else { fprintf(stderr, "Unknown option '-%c'\n", c); exit(1); }
		}
		argc--;
		argv++;
	}
	if (argc < 3) { printf("Not enough arguments. Check help.\n"); usage(); return 0; }
	in = argv[1];
	out = argv[2];
	if (strcmp(in, out) == 0) { fprintf(stdout, "Input and output are the same file: '%s'.\n", in); return 1; }
	if (!overwrite) {
		FILE *f = fopen(out, "rb");
		if (f) { fclose(f); fprintf(stderr, "File '%s' already exists. Try `-f' to overwrite.\n", out); return 1; }
	}
	memset(&im, 0, sizeof im);
	memset(&enc, 0, sizeof enc);
	im.setup = (codec_setup *)malloc(sizeof(codec_setup));
	im.setup->quality_setting = (unsigned char)quality;
	read_image_bmp(in, &enc, &im, 8);
	encode_image(&im, &enc, 8);
	return write_compressed_file(&im, &enc, out) == 0 ? 0 : 1;
}
