/*
 * worker_loop.c -- the reference's steady-state loop, written against the headers in include/ exactly
 * the way reference ocl_aho_grep.c:36-144 (cpu_worker) and :272-308 (callback_match) use
 * the API: ocl_worker_ctx_create / _init, databuf_add_fd, the five-call sequence, the
 * counters in struct ocl_worker_ctx, ctx->patterns[idx].{iid,pattern,n}, ctx->db->file_ids.
 *
 *   gcc -Iinclude examples/worker_loop.c -Lgpu_pattern_matching_b200 -lacmatch_b200 \
 *       -Wl,-rpath,$PWD/gpu_pattern_matching_b200 -o worker_loop
 *   ./worker_loop patterns.txt input.bin [-x]
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "acm.h"
#include "ocl_aho_match.h"
#include "ocl_worker.h"
#include "utils.h"

static int verbose = 1;

static int
callback_match(int f_id, int p_idx, int c_id, int off, void *uarg)
{
	struct ocl_worker_ctx *ctx = uarg;

	ctx->matches_reported += 1;
	if (verbose)
		printf("Pattern %d ('%s') found in file '%s' at offset %d [relative: %d]\n",
		    ctx->patterns[p_idx].iid, ctx->patterns[p_idx].pattern,
		    ctx->filenames[ctx->db->file_ids[c_id]], off, off - ctx->db->h_indices[c_id]);
	(void)f_id;
	return 0;
}

int
main(int argc, char **argv)
{
	struct ocl_worker_ctx *ctx;
	char *filenames[1];
	int fds[1], e, done = 0;
	size_t rd_bytes = 0, t0;

	if (argc < 3) {
		fprintf(stderr, "usage: %s patterns input [-x] [-q]\n", argv[0]);
		return 2;
	}
	for (int i = 3; i < argc; i++)
		if (!strcmp(argv[i], "-q"))
			verbose = 0;
	filenames[0] = argv[2];
	fds[0] = open(argv[2], O_RDONLY);
	if (fds[0] < 0) {
		perror(argv[2]);
		return 1;
	}
	ctx = ocl_worker_ctx_create(0);
	if (!ctx) {
		fprintf(stderr, "no device: %s\n", acm_last_error());
		return 1;
	}
	/* README shape: -G 32768 -B 4096 -L 1024 -R 16 */
	e = ocl_worker_ctx_init(ctx, 0, 1024, 32768, 0, argv[1], argc > 3 && !strcmp(argv[3], "-x"), -1,
	    4096, MAX_RESULTS, verbose, 0, 0, 0, 1, 1, fds, filenames);
	if (e != 0) {
		fprintf(stderr, "init failed: %s\n", acm_last_error());
		return 1;
	}
	t0 = gettime();
	while (!done) {
		e = databuf_add_fd(ctx->db, ctx->fds[0], 0, &rd_bytes);
		ctx->bytes += rd_bytes;
		if (rd_bytes == 0)
			done = 1;
		else if (e != -1 && e != -2)
			continue;
		if (ctx->db->chunks > 0) {
			databuf_copy_host_to_device(ctx->db, ctx->cl.queue);
			ocl_aho_match(&ctx->cl, ctx->db, ctx->acsm, ctx->local_ws, 1 /* stream */);
			databuf_copy_device_to_host(ctx->db, ctx->cl.queue);
			ctx->matches_total += databuf_process_results(ctx->db, callback_match, ctx);
			databuf_reset(ctx->db);
			ctx->rounds++;
		}
	}
	{
		double sec = (gettime() - t0) / 1e6;
		printf("STATS: matches %zu reported %zu bytes %zu rounds %zu states %d automaton %.1f MB "
		    "%.3f s throughput %.1f Mbps\n", ctx->matches_total, ctx->matches_reported, ctx->bytes,
		    ctx->rounds, acsm_get_states(ctx->acsm), acsm_get_size(ctx->acsm) / 1048576.0, sec,
		    ctx->bytes * 8.0 / 1048576.0 / sec);
	}
	close(fds[0]);
	ocl_worker_ctx_free(ctx);
	return 0;
}
