/*
 * databuf.h -- chunked input buffer + results, B200 edition.  Drop-in for reference
 * databuf.h:15-174 (+ databuf_add_chunk, reference databuf.c:488): same struct
 * name, same field names, same functions and return codes.
 *
 * What differs underneath:
 *   - h_data is pinned host memory; d_data is one device allocation preceded by a
 *     carry area that keeps the last Lmax-1 bytes of the previous buffer, which is
 *     how a match that straddles two buffers is found (the reference hands one
 *     DFA state across, databuf.h:35 / ahomatch.cl:42-43, and loses matches that
 *     straddle chunks, ahomatch.cl:151-155);
 *   - the device produces ONE sorted, compacted list (pattern index, end offset)
 *     per buffer; h_results_comp / h_results2_comp hold it in the reference's
 *     compact format [total, v0..v(total-1), tail] (reference compactarray.cl:49-55),
 *     and the per-chunk bucket arrays h_results / h_results2 (column-major,
 *     reference ahomatch.cl:67-73) are derived from it on the host so that code
 *     reading either keeps working;
 *   - databuf_process_results() reports EVERY match (the reference drops those
 *     beyond max_results-1 per chunk, databuf.c:766-768) in (offset, index) order.
 */
#ifndef _DATABUF_H_
#define _DATABUF_H_

#include <stdio.h>

#include "acm_compat.h"
#include "ocl_context.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference databuf.h:9 */
#define MAX_RESULTS 16

struct databuf {
	unsigned char	*h_data;	 /* host data (pinned)                         */
	int 		*h_indices;	 /* chunk start offsets                        */
	int		*h_sizes;	 /* chunk sizes                                */
	int		*h_results;	 /* buckets: pattern index, column-major       */
	int		*h_results2;	 /* buckets: end offset, column-major          */
	int		*h_prefixsum;	 /* exclusive scan of the per-chunk counts     */
	int		*h_results_comp; /* [total, pattern indices..., tail]          */
	int		*h_results2_comp;/* [total, end offsets..., tail]              */

	size_t		results_comp_size; /* capacity of h_results_comp (ints)       */
	size_t		results2_comp_size;

	int		*file_ids;	 /* file id per chunk                          */
	int		mapped;
	int		max_results;
	long		last_state;	 /* kept for source compatibility; the carry is bytes, see above */
	size_t		max_chunks;
	size_t		max_chunk_size;
	size_t		size;		 /* max_chunks * max_chunk_size                */
	size_t		chunks;
	size_t		bytes;

	cl_mem		d_data;
	cl_mem		d_indices;	 /* unused: the device sees one contiguous stream */
	cl_mem		d_sizes;	 /* unused                                      */
	cl_mem		d_results;	 /* device buckets, allocated on first post-pass use */
	cl_mem		d_results2;
	cl_mem		d_prefixsum;
	cl_mem		d_results_comp;
	cl_mem		d_results2_comp;

	cl_mem		p_data;		 /* unused (h_data itself is pinned)            */
	cl_mem		p_indices;
	cl_mem		p_sizes;
	cl_mem		p_results;
	cl_mem		p_results2;
	cl_mem		p_prefixsum;
	cl_mem		p_results_comp;
	cl_mem		p_results2_comp;

	cl_mem		*ScanPartialSums;	/* unused: the scan is single pass       */
	unsigned int	ScanPartialSums_size;

	struct clconf	*cl;
	void		*priv;
};

/* (max_chunks, max_chunk_size, max_results, mapped, conf); NULL on failure.  reference databuf.c:77 */
struct databuf *databuf_new(size_t, size_t, int, int, struct clconf *);

/*
 * read() as much of fd as fits, in max_chunk_size chunks, zero-padding a short
 * last chunk.  Returns >0 (bytes read, buffer can take more), 0 (EOF), -1 (all
 * chunks used), -2 (all bytes used), -4 (the read failed: acm_last_error(); the reference
 * aborts); *rd_bytes always set.  reference databuf.c:327
 */
int  databuf_add_fd(struct databuf *, int, int, size_t *);

/* text mode: one chunk per line, optionally padded to 16 bytes.  reference databuf.c:413 */
int  databuf_add_fp(struct databuf *, FILE *, int, int, size_t *, size_t *);

/* one chunk from memory; -3 too large, -1 no chunk left, -2 no room.  reference databuf.c:488 */
int  databuf_add_chunk(struct databuf *, char *, size_t, int, char);

void databuf_reset(struct databuf *);   /* reference databuf.c:534 */
void databuf_clear(struct databuf *);   /* reference databuf.c:547; also forgets the carry */

/* H2D of bytes [0, db->bytes) on the queue's stream.  reference databuf.c:575 */
void databuf_copy_host_to_device(struct databuf *, cl_command_queue);

/* D2H of the sorted match list; fills the compact and the bucket arrays.  reference databuf.c:604 */
void databuf_copy_device_to_host(struct databuf *, cl_command_queue);

/*
 * cb(file id, pattern index, chunk index, end offset + 1, uarg) per match, in
 * (offset, index) order; returns the number of matches.  reference databuf.c:788
 * (the "+ 1" is what the reference's default build passes, databuf.c:771).
 */
int  databuf_process_results(struct databuf *db,
         int (*cb)(int file_idx, int patrn_idx, int chunk_idx, int offset, void *uarg), void *uarg);

void databuf_free(struct databuf *, int, cl_command_queue);   /* reference databuf.c:801 */

/* additions */
int    databuf_status(struct databuf *);            /* 0 or the last negative ACM_ERR_* */
size_t databuf_match_count(struct databuf *);       /* matches of the last ocl_aho_match() (after databuf_copy_device_to_host: those that passed the per-file rule) */
/*
 * Per-file semantics (default on; off = the reference's behaviour, also ACM_DATABUF_STREAM_QUIRK=1
 * in the environment at databuf_new): a match is reported only if all its bytes are real bytes of
 * ONE file -- not a chunk's zero padding, not the tail of one file plus the head of the next --
 * and a match may reach back into the previous buffer only where that buffer ended with bytes
 * of the same file.  The reference carries one DFA state across chunks, buffers and files alike
 * (ahomatch.cl:38-45,86-93, databuf.c:610,622).
 */
void   databuf_set_file_semantics(struct databuf *, int on);
/* allocates d_results, d_results2, d_prefixsum, d_results_comp, d_results2_comp (reference shapes) */
int    databuf_alloc_postpass(struct databuf *);
/*
 * What databuf_add_fd() reads with: like read(fd, buf, want), but a regular file is read by several
 * threads with pread() on adjacent segments (one read() of a 128 MiB buffer copies out of the page
 * cache at 3 GB/s on one core; ACM_READ_THREADS, default 4, 1 = plain read) and the file offset is
 * advanced by what was read.  Pipes, FIFOs and small requests take the plain read().  Returns the
 * number of contiguous bytes read, 0 at end of file, -1 on error.
 */
long   databuf_read_fd(int fd, void *buf, size_t want);

#ifdef __cplusplus
}
#endif
#endif /* _DATABUF_H_ */
