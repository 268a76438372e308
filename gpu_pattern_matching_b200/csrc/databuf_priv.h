/* databuf_priv.h -- what hangs off struct databuf.priv.  Private to the library. */
#ifndef DATABUF_PRIV_H
#define DATABUF_PRIV_H

#include <pthread.h>
#include <stdint.h>
#include <sys/types.h>

#include "../../include/acm.h"

/* bytes reserved in front of d_data for the cross-buffer carry (>= Lmax - 1 symbols) */
#define DATABUF_CARRY_CAP 65536

/*
 * Read-ahead for databuf_add_fd(): when one read fills a whole buffer from a regular file, the
 * worker is about to spend ~2.5 ms on H2D + scan + D2H + callbacks and then ask for the next
 * buffer of the same file; a helper thread reads that next buffer into a second pinned buffer
 * meanwhile, and the next databuf_add_fd() just swaps the two.
 */
struct databuf_readahead {
	pthread_t       thread;
	pthread_mutex_t lock;
	pthread_cond_t  cond;
	int             started;         /* thread exists                                */
	int             quit;
	int             busy;            /* a request is being served                     */
	int             fd;              /* request / result: file, offset, bytes          */
	off_t           off;
	size_t          want;
	long            got;             /* result (< 0: error), valid when !busy && have */
	int             have;
	unsigned char  *buf;             /* the second pinned buffer (h_alt)              */
	/* ACM_DATABUF_STATS=1: printed by databuf_free */
	double          t_direct, t_wait, t_ra_read;   /* seconds in foreground reads, waiting for the read-ahead, reading ahead */
	long            n_direct, n_taken, n_missed;
};

struct databuf_priv {
	struct acm_device    *dev;
	struct acm_scanner   *scanner;
	struct acm_automaton *scanner_aut;   /* automaton the scanner was built for */
	unsigned char        *d_base;        /* allocation: [carry area | data]     */
	unsigned char        *d_carry_tmp;
	uint64_t              carry_len;     /* symbols currently held in the carry */
	uint64_t              n_matches;     /* of the last ocl_aho_match(): on the device, before the file filter */
	uint64_t              n_valid;       /* after it (what databuf_process_results reports) */
	uint64_t             *h_off;
	uint32_t             *h_pat;
	uint64_t              h_cap;
	int                   fetched;
	int                   status;
	int                   sym_size;      /* 1 bytes, 2 ushort symbols           */
	/* per-file semantics (bytes only): a match must lie inside the real bytes of ONE file */
	int                   file_semantics;/* 1 (default) / 0 = the reference's one-stream quirk */
	int                   carry_file;    /* file id the carried bytes belong to, valid when carry_real > 0 */
	uint64_t              carry_real;    /* trailing symbols of the carry that are real, contiguous bytes of carry_file */
	uint64_t              filt_carry_real;  /* the two above as they were when the last match STARTED: what */
	int                   filt_carry_file;  /* databuf_copy_device_to_host filters that match's list against */
	int                  *run_id;        /* [max_chunks] scratch: chunks of one contiguous stretch of one file share an id */
	int                   have_match;    /* ocl_aho_match has run on this databuf */
	int                   buckets_on_device; /* d_results / d_results2 hold the bucket view of the last match */
	struct databuf_readahead ra;
};

struct acm_automaton;
/* compat_ocl.c: scanner creation + first launches ahead of the scan loop (ocl_worker_ctx_init) */
int databuf_prepare(struct databuf *db, struct acm_automaton *aut, int sym_size);

#endif
