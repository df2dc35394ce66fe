/* oracle/tap.h -- TEST INFRASTRUCTURE ONLY.
 * Stage-tap registry used by the tap-instrumented build of the reference
 * (oracle/make_tapped.py inserts NHW_TAP(...) calls at chosen line numbers of a
 * scratch copy; arithmetic is untouched). */
#ifndef NHW_ORACLE_TAP_H
#define NHW_ORACLE_TAP_H
#include <stddef.h>
void nhw_tap(const char *name, const void *ptr, size_t bytes);
#define NHW_TAP(name, ptr, bytes) nhw_tap((name), (const void *)(ptr), (size_t)(bytes))
#endif
