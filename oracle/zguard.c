/*
 * oracle/zguard.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Canonicalising allocator + exit() trap for the compiled reference.
 * The reference encoder reads a few bytes outside some heap blocks and reads malloc'd
 * memory before writing it; those reads reach the .nhw bytes.  Linking the UNMODIFIED
 * reference objects with
 *     -Wl,--wrap=malloc,--wrap=calloc,--wrap=free,--wrap=exit
 * routes them here: every block is zero-filled and surrounded by NHW_GUARD zero bytes,
 * so "uninitialised" == 0 and "a little out of bounds" == 0.  That is the parity contract
 * the CUDA path implements (explicit zero halos).
 *
 * exit() inside the reference (codebook overflow, bad BMP ...) is turned into a longjmp
 * back to the glue entry point so one bad image cannot kill a batch run.
 */
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>

#define NHW_GUARD 65536

#ifndef NHW_NO_ZGUARD
void *__real_malloc(size_t);
void *__real_calloc(size_t, size_t);
void __real_free(void *);

void *__wrap_malloc(size_t n)
{
	char *p = (char *)__real_calloc(1, n + 2 * NHW_GUARD);
	return p ? p + NHW_GUARD : NULL;
}

void *__wrap_calloc(size_t a, size_t b)
{
	char *p = (char *)__real_calloc(1, a * b + 2 * NHW_GUARD);
	return p ? p + NHW_GUARD : NULL;
}

void __wrap_free(void *p)
{
	if (p) __real_free((char *)p - NHW_GUARD);
}

#else
/* timing build (libnhwref_enc_stock.so): blocks are NOT zero-filled (malloc stays malloc, so the reference runs at
 * its own speed), but every block is still padded by NHW_GUARD bytes on both sides: the reference's out-of-bounds
 * reads then land in mapped memory.  Without the padding a read past a large (mmap'd) block can hit an unmapped
 * page and kill the process -- seen once on a 32-thread baseline run.  Never used for parity. */
void *__real_malloc(size_t);
void *__real_calloc(size_t, size_t);
void __real_free(void *);

void *__wrap_malloc(size_t n)
{
	char *p = (char *)__real_malloc(n + 2 * NHW_GUARD);
	return p ? p + NHW_GUARD : NULL;
}

void *__wrap_calloc(size_t a, size_t b)
{
	char *p = (char *)__real_calloc(1, a * b + 2 * NHW_GUARD);
	return p ? p + NHW_GUARD : NULL;
}

void __wrap_free(void *p)
{
	if (p) __real_free((char *)p - NHW_GUARD);
}
#endif

__thread jmp_buf nhwref_exit_jmp;
__thread int nhwref_exit_armed = 0;
__thread int nhwref_exit_code = 0;

void __real_exit(int);

void __wrap_exit(int code)
{
	if (nhwref_exit_armed) {
		nhwref_exit_code = code;
		longjmp(nhwref_exit_jmp, 1);
	}
	__real_exit(code);
}
