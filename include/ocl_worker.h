/*
 * ocl_worker.h -- per-thread worker context.  Drop-in for reference
 * ocl_worker.h:9-82: same struct (field names and order), same three functions.
 */
#ifndef _OCL_WORKER_H_
#define _OCL_WORKER_H_

#include "ocl_context.h"
#include "acsmx.h"
#include "databuf.h"

#ifdef __cplusplus
extern "C" {
#endif

struct ocl_worker_ctx {
	int            id;
	int            text_mode;
	int            follow;
	int            verbose;
	int            thread_no;
	int            total_files;
	int            *fds;
	char           **filenames;
	size_t         matches_total;
	size_t         matches_reported;
	size_t         bytes;
	size_t         lines;
	size_t         rounds;
	size_t         global_ws;
	size_t         local_ws;
	struct clconf  cl;
	struct databuf *db;
	acsm_t         *acsm;
	acsm_pattern_t *patterns;
	size_t         patterns_size;
};

/* (device position) -> context or NULL.  reference ocl_worker.c:21 */
struct ocl_worker_ctx *ocl_worker_ctx_create(int);

/*
 * (ctx, device position, local ws, global ws = chunks per buffer, mapped, pattern
 * file, hex patterns, pattern size limit or -1, max chunk size, max results,
 * verbose, text mode, follow, thread id, threads, total files, fds, filenames).
 * Parses the pattern file (plain / hex / categorical `ID "pattern"` lines,
 * reference ocl_worker.c:74-145), compiles, uploads, builds the patterns table
 * and the databuf.  0 on success, -1 on failure.  reference ocl_worker.c:48
 */
int ocl_worker_ctx_init(struct ocl_worker_ctx *, int, size_t, size_t, int, char *, int, int, size_t,
        int, int, int, int, int, int, int, int *, char **);

/* reference ocl_worker.c:192 */
void ocl_worker_ctx_free(struct ocl_worker_ctx *);

#ifdef __cplusplus
}
#endif
#endif /* _OCL_WORKER_H_ */
