/*
 * ocl_context.h -- device context of a worker.  Same struct name and field names as
 * reference ocl_context.h:8-28 so caller code (ctx->cl.queue, &ctx->cl ...) compiles
 * unchanged; the handles are the opaque types of acm_compat.h.  `ctx` is the GPU,
 * `queue` wraps the CUDA stream all work of this worker is ordered on.  The
 * program_ / kernel_ fields are unused (kernels are compiled ahead of time for
 * sm_100a, not JIT-built from .cl files in the CWD as reference ocl_aho_match.c:21).
 */
#ifndef _OCL_CONTEXT_H_
#define _OCL_CONTEXT_H_

#include "acm_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

struct clconf {
	cl_platform_id   platform;
	cl_device_id     dev;
	cl_context       ctx;
	cl_command_queue queue;

	cl_program       program_aho_match;
	cl_kernel        kernel_aho_match;

	cl_program       program_prefixsum;
	cl_kernel        kernel_prescan;
	cl_kernel        kernel_prescan_store_sum;
	cl_kernel        kernel_prescan_store_sum_non_power_of_two;
	cl_kernel        kernel_prescan_non_power_of_two;
	cl_kernel        kernel_uniform_add;

	cl_program       program_compact_array;
	cl_kernel        kernel_compact_array;

	cl_device_type   type;
};

/*
 * replaces reference ocl_context.c:19.  (conf, device position, sub position):
 * opens CUDA device `pos` (sub position ignored) and creates the queue.  On
 * failure conf->ctx stays NULL and acm_last_error() says why; the reference
 * exit(1)s instead.
 */
void clinitctx(struct clconf *, int, int);

/* addition: releases what clinitctx created */
void clfreectx(struct clconf *);

#ifdef __cplusplus
}
#endif
#endif /* _OCL_CONTEXT_H_ */
