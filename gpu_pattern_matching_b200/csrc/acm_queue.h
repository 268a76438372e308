/*
 * acm_queue.h -- what the opaque cl_context / cl_command_queue of acm_compat.h
 * really are.  Private to the library.
 */
#ifndef ACM_QUEUE_H
#define ACM_QUEUE_H

#include "../../include/acm.h"
#include "../../include/acm_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

struct acm_queue {
	struct acm_device *dev;
};

/* device behind (ctx, queue); either may be NULL, in which case the default device is used */
static inline struct acm_device *
acm_queue_device(cl_context ctx, cl_command_queue queue)
{
	if (queue && queue->dev)
		return queue->dev;
	if (ctx)
		return (struct acm_device *)ctx;
	return acm_default_device();
}

#ifdef __cplusplus
}
#endif
#endif
