/*
 * utils.h -- the two helpers of reference utils.h / utils.c that sit on the path:
 * the hex decoder used by the pattern-file parser and the microsecond clock.
 */
#ifndef _UTILS_H_
#define _UTILS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference utils.h:14 */
#define MAX_PAT_SIZE 4096

/*
 * "4d5a90" -> {0x4d,0x5a,0x90}; upper or lower case (reference utils.c:33-54).
 * Returns a malloc'd buffer of strlen/2 bytes, or NULL for an odd-length or
 * non-hex string (the reference exit()s on odd length, utils.c:39-42).
 */
unsigned char *printable_hex_to_bytes(unsigned char *);

/* CLOCK_MONOTONIC in microseconds (reference utils.c:61-68) */
size_t gettime(void);

/*
 * Pattern-file loader behind ocl_worker_ctx_init(): adds every line of `path` to
 * the automaton.  Returns the number of patterns added, or -1 (cannot open /
 * bad categorical id) / -2 (bad hex line).
 */
struct _acsm;
int acsm_load_pattern_file(struct _acsm *, const char *path, int hex_pat, int pat_size_limit);

#ifdef __cplusplus
}
#endif
#endif
