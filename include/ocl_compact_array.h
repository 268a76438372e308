/* ocl_compact_array.h -- drop-in for reference ocl_compact_array.h:12-22 */
#ifndef _OCL_COMPACT_ARRAY_H_
#define _OCL_COMPACT_ARRAY_H_

#include "ocl_context.h"
#include "databuf.h"

#ifdef __cplusplus
extern "C" {
#endif

void ocl_compact_array_init(struct clconf *c);  /* nothing to build */
void ocl_compact_array_close(struct clconf *c);

/*
 * compacts the column-major buckets d_results / d_results2 into d_results_comp /
 * d_results2_comp = [total, values..., tail] using d_prefixsum
 * (reference ocl_compact_array.c:130-172); the size_t is the local work size,
 * ignored.
 */
void ocl_compact_array(struct clconf *, struct databuf *, size_t);

#ifdef __cplusplus
}
#endif
#endif
