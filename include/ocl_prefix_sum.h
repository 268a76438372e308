/* ocl_prefix_sum.h -- drop-in for reference ocl_prefix_sum.h:12-22 */
#ifndef _OCL_PREFIX_SUM_H_
#define _OCL_PREFIX_SUM_H_

#include "ocl_context.h"
#include "databuf.h"

#ifdef __cplusplus
extern "C" {
#endif

void ocl_prefix_sum_init(struct clconf *c);     /* nothing to build */
void ocl_prefix_sum_close(struct clconf *c);

/*
 * exclusive prefix sum of the first n per-chunk counts db->d_results[0..n) into
 * db->d_prefixsum (reference ocl_prefix_sum.c:165,218), one single-pass
 * decoupled-look-back kernel instead of up to 5 launches per level.
 */
void ocl_prefix_sum(struct clconf *, struct databuf *, unsigned int);

#ifdef __cplusplus
}
#endif
#endif
