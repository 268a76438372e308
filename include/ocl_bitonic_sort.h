/*
 * ocl_bitonic_sort.h -- drop-in for reference ocl_bitonic_sort.h:13-18.  Same
 * signature; underneath it is a stable LSD radix sort, so any length works (the
 * reference requires a power of two, ocl_bitonic_sort.c:150-153).
 */
#ifndef _OCL_BITONIC_SORT_
#define _OCL_BITONIC_SORT_

#include "ocl_context.h"

#ifdef __cplusplus
extern "C" {
#endif

int ocl_bitonic_sort_init(struct clconf *);
int ocl_bitonic_sort_close(struct clconf *);

/*
 * (conf, dst keys, dst values, src keys, src values, batch, array length, dir):
 * sorts `batch` consecutive arrays of `array length` uint32 (key, value) pairs,
 * dir != 0 ascending, 0 descending (reference BitonicSort.cl sortDir).  Returns 0,
 * or a negative error.
 */
int ocl_bitonic_sort(struct clconf *, cl_mem, cl_mem, cl_mem, cl_mem, unsigned int, unsigned int,
        unsigned int dir);

#ifdef __cplusplus
}
#endif
#endif
