/*
 * b200_aho_grep -- command-line front end over libacmatch_b200.so with the flags, the -v
 * line format and the STATS block of the reference's ocl_aho_grep
 * (reference ocl_aho_grep.c:150-204 usage, :272-308 callback_match, :354-647 main,
 * README:5-83), so scripts that drive the reference (apps/sentiment_analysis.py:188-198
 * parses the "Pattern %d ('%s') found in file ..." lines) keep working.
 *
 * Written from scratch on the worker API (ocl_worker.h / databuf.h / ocl_aho_match.h):
 * -w worker threads, files striped over them (thread t takes files t, t+w, ...), each thread
 * runs the five-call sequence per buffer.  Differences from the reference, all inherited from
 * the library: every match is reported (no per-chunk cap), in (offset, pattern) order, and
 * matches straddling chunks or buffers are found exactly once.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "acm.h"
#include "ocl_aho_match.h"
#include "ocl_worker.h"
#include "utils.h"

static volatile sig_atomic_t terminate;
static pthread_mutex_t out_lock = PTHREAD_MUTEX_INITIALIZER;

static void
on_sigint(int sig)
{
	(void)sig;
	terminate = 1;
}

static void
usage(void)
{
	printf("\nUsage:\n"
	    "    b200_aho_grep -f file -p file [-B chunk_size] [-D devpos] [-G global_ws]\n"
	    "                  [-L local_ws] [-m max] [-w cpu_threads] [-R max] [-tvxFM]\n"
	    "    b200_aho_grep -h\n\n"
	    "Options (same letters as ocl_aho_grep):\n"
	    "  -f file        input: one directory, one file, or comma-separated files\n"
	    "  -p file        pattern file, one pattern per line (plain, -x hex, or `ID \"pattern\"`)\n"
	    "  -F             follow: keep processing data appended to the files (until SIGINT)\n"
	    "  -B chunk_size  bytes per chunk (default 4096, rounded up to 16)\n"
	    "  -D devpos      CUDA device ordinal, or a comma-separated list: worker t runs on the\n"
	    "                 (t mod n)-th of them (default 0)\n"
	    "  -G global_ws   chunks per buffer; buffer = global_ws * chunk_size (default 32768)\n"
	    "  -L local_ws    accepted for compatibility, ignored\n"
	    "  -m max         limit patterns to max bytes\n"
	    "  -w threads     worker threads feeding the device (default 2)\n"
	    "  -R max         result slots per chunk in the bucket view (default 16)\n"
	    "  -v             print every match\n"
	    "  -t             text mode: one chunk per line\n"
	    "  -x             patterns are printable hex\n"
	    "  -M             accepted for compatibility, ignored\n"
	    "  -h             this message\n");
	exit(EXIT_FAILURE);
}

/* reference ocl_aho_grep.c:272-308 */
static int
callback_match(int f_id, int p_idx, int c_id, int off, void *uarg)
{
	struct ocl_worker_ctx *ctx = uarg;
	const struct databuf *db = ctx->db;
	int i;

	(void)f_id;
	ctx->matches_reported += 1;
	if (!ctx->verbose)
		return 0;
	pthread_mutex_lock(&out_lock);
	printf("Pattern %d ('%s') found in file '%s' at offset %d [relative: %d]\n",
	    ctx->patterns[p_idx].iid, ctx->patterns[p_idx].pattern, ctx->filenames[db->file_ids[c_id]], off,
	    off - db->h_indices[c_id]);
	if (ctx->text_mode) {
		/* the matching line */
		for (i = 0; i < db->h_sizes[c_id]; i++)
			putchar(db->h_data[db->h_indices[c_id] + i]);
		if (db->h_sizes[c_id] == 0 || db->h_data[db->h_indices[c_id] + db->h_sizes[c_id] - 1] != '\n')
			putchar('\n');
	}
	pthread_mutex_unlock(&out_lock);
	return 0;
}

static void
process_buffer(struct ocl_worker_ctx *ctx)
{
	if (ctx->db->chunks == 0)
		return;
	databuf_copy_host_to_device(ctx->db, ctx->cl.queue);
	ocl_aho_match(&ctx->cl, ctx->db, ctx->acsm, ctx->local_ws, 1 /* stream */);
	databuf_copy_device_to_host(ctx->db, ctx->cl.queue);
	if (databuf_status(ctx->db) != ACM_OK) {
		fprintf(stderr, "ERROR: %s\n", acm_last_error());
		exit(1);
	}
	ctx->matches_total += (size_t)databuf_process_results(ctx->db, callback_match, ctx);
	databuf_reset(ctx->db);
	ctx->rounds++;
}

/* reference ocl_aho_grep.c:36-144 */
static void *
cpu_worker(void *arg)
{
	struct ocl_worker_ctx *ctx = arg;
	int cur_file = ctx->id, e;
	size_t rd_bytes, rd_lines;
	FILE **fps = calloc((size_t)ctx->total_files, sizeof(FILE *));

	while (cur_file < ctx->total_files && !terminate) {
		rd_bytes = rd_lines = 0;
		if (ctx->text_mode) {
			if (!fps[cur_file])
				fps[cur_file] = fdopen(ctx->fds[cur_file], "r");
			e = databuf_add_fp(ctx->db, fps[cur_file], cur_file, 1, &rd_bytes, &rd_lines);
		} else {
			e = databuf_add_fd(ctx->db, ctx->fds[cur_file], cur_file, &rd_bytes);
		}
		ctx->bytes += rd_bytes;
		ctx->lines += rd_lines;
		if (e == -4) {                      /* the read failed: the reference aborts here */
			fprintf(stderr, "ERROR: %s: %s\n", ctx->filenames[cur_file], acm_last_error());
			exit(1);
		}
		if (e == -1 || e == -2) {           /* buffer full */
			process_buffer(ctx);
			continue;
		}
		if (rd_bytes != 0)
			continue;
		/* end of this file (for now) */
		if (ctx->follow) {
			process_buffer(ctx);
			if (ctx->text_mode)
				clearerr(fps[cur_file]);
			usleep(20000);
			cur_file += ctx->thread_no;
			if (cur_file >= ctx->total_files)
				cur_file = ctx->id;
			continue;
		}
		if (ctx->text_mode)
			fclose(fps[cur_file]);
		else
			close(ctx->fds[cur_file]);
		cur_file += ctx->thread_no;
	}
	process_buffer(ctx);                    /* whatever is left */
	free(fps);
	return NULL;
}

struct file_list {
	char **names;
	int n, cap;
};

static void
fl_add(struct file_list *fl, const char *path)
{
	if (fl->n == fl->cap) {
		fl->cap = fl->cap ? fl->cap * 2 : 64;
		fl->names = realloc(fl->names, (size_t)fl->cap * sizeof(char *));
	}
	fl->names[fl->n++] = strdup(path);
}

static void
fl_walk(struct file_list *fl, const char *path)
{
	struct stat st;

	if (stat(path, &st) != 0) {
		fprintf(stderr, "ERROR: cannot stat %s\n", path);
		exit(1);
	}
	if (S_ISDIR(st.st_mode)) {
		DIR *d = opendir(path);
		struct dirent *de;
		char sub[4096];
		if (!d)
			return;
		while ((de = readdir(d)) != NULL) {
			if (!strcmp(de->d_name, ".") || !strcmp(de->d_name, ".."))
				continue;
			snprintf(sub, sizeof(sub), "%s/%s", path, de->d_name);
			fl_walk(fl, sub);
		}
		closedir(d);
	} else if (S_ISREG(st.st_mode) || S_ISFIFO(st.st_mode)) {
		fl_add(fl, path);
	}
}

static int
cmp_names(const void *a, const void *b)
{
	return strcmp(*(char *const *)a, *(char *const *)b);
}

int
main(int argc, char **argv)
{
	char *data_path = NULL, *pat_path = NULL, *tok, *save;
	int opt, text_mode = 0, verbose = 0, hex_pat = 0, follow = 0, mapped = 0;
	int devs[64] = {0}, ndev = 1;
	int thread_no = 2, pat_size_limit = -1, max_results = MAX_RESULTS, i;
	size_t max_chunk_size = 4096, global_ws = 32768, local_ws = 1024;
	struct file_list fl = {0};
	struct ocl_worker_ctx **w;
	pthread_t *threads;
	size_t t0, t1, matches = 0, reported = 0, bytes = 0, lines = 0, rounds = 0;
	int *fds;

	while ((opt = getopt(argc, argv, "f:m:p:tw:vxB:D:FG:L:R:Mh")) != -1) {
		switch (opt) {
		case 'f': data_path = strdup(optarg); break;
		case 'm': pat_size_limit = atoi(optarg); break;
		case 'p': pat_path = strdup(optarg); break;
		case 't': text_mode = 1; break;
		case 'w': thread_no = atoi(optarg); break;
		case 'v': verbose = 1; break;
		case 'x': hex_pat = 1; break;
		case 'B': max_chunk_size = (size_t)atol(optarg); break;
		case 'D': {
			char *list = strdup(optarg), *sv, *t;
			ndev = 0;
			for (t = strtok_r(list, ",", &sv); t && ndev < 64; t = strtok_r(NULL, ",", &sv))
				devs[ndev++] = atoi(t);
			free(list);
			if (ndev == 0)
				usage();
			break;
		}
		case 'F': follow = 1; break;
		case 'G': global_ws = (size_t)atol(optarg); break;
		case 'L': local_ws = (size_t)atol(optarg); break;
		case 'R': max_results = atoi(optarg); break;
		case 'M': mapped = 1; break;
		default: usage();
		}
	}
	if (!data_path || !pat_path || thread_no < 1 || max_results < 2 || max_chunk_size == 0 || global_ws == 0)
		usage();
	/* reference align_parameters(): 16-byte multiples (ocl_aho_grep.c:316-352) */
	max_chunk_size = (max_chunk_size + 15) & ~(size_t)15;

	for (tok = strtok_r(data_path, ",", &save); tok; tok = strtok_r(NULL, ",", &save))
		fl_walk(&fl, tok);
	if (fl.n == 0) {
		fprintf(stderr, "ERROR: no input files\n");
		return 1;
	}
	qsort(fl.names, (size_t)fl.n, sizeof(char *), cmp_names);
	fds = calloc((size_t)fl.n, sizeof(int));
	for (i = 0; i < fl.n; i++) {
		fds[i] = open(fl.names[i], O_RDONLY);
		if (fds[i] < 0) {
			perror(fl.names[i]);
			return 1;
		}
	}
	if (thread_no > fl.n)
		thread_no = fl.n;

	w = calloc((size_t)thread_no, sizeof(*w));
	threads = calloc((size_t)thread_no, sizeof(*threads));
	for (i = 0; i < thread_no; i++) {
		/* workers are striped over the devices like files over the workers (reference
		 * ocl_aho_grep.c:48,87,498-502: every worker owns a context on its device) */
		const int dev_pos = devs[i % ndev];
		w[i] = ocl_worker_ctx_create(dev_pos);
		if (!w[i]) {
			fprintf(stderr, "ERROR: %s\n", acm_last_error());
			return 1;
		}
		if (ocl_worker_ctx_init(w[i], dev_pos, local_ws, global_ws, mapped, pat_path, hex_pat, pat_size_limit,
		    max_chunk_size, max_results, verbose, text_mode, follow, i, thread_no, fl.n, fds, fl.names) != 0) {
			fprintf(stderr, "ERROR: %s\n", acm_last_error());
			return 1;
		}
	}
	signal(SIGINT, on_sigint);
	t0 = gettime();
	for (i = 0; i < thread_no; i++)
		pthread_create(&threads[i], NULL, cpu_worker, w[i]);
	for (i = 0; i < thread_no; i++)
		pthread_join(threads[i], NULL);
	t1 = gettime();

	for (i = 0; i < thread_no; i++) {
		matches += w[i]->matches_total;
		reported += w[i]->matches_reported;
		bytes += w[i]->bytes;
		lines += w[i]->lines;
		rounds += w[i]->rounds;
	}
	/* reference ocl_aho_grep.c:615-631 */
	printf("-------------- STATS --------------\n");
	printf("Matches:             %lu\n", (unsigned long)matches);
	printf("Matches reported:    %lu\n", (unsigned long)reported);
	printf("Time (secs):         %.5f\n", (double)(t1 - t0) / 1000000);
	printf("Automaton states:    %d\n", acsm_get_states(w[0]->acsm));
	printf("Automaton size (MB): %.3f\n", (double)acsm_get_size(w[0]->acsm) / 1048576);
	printf("Processed bytes:     %lu\n", (unsigned long)bytes);
	if (lines)
		printf("Processed lines:     %lu\n", (unsigned long)lines);
	printf("Processed files:     %d\n", fl.n);
	printf("Kernel launches:     %d\n", (int)rounds);
	printf("Throughput (Mbps):   %.3f\n", ((double)bytes * 8 / 1048576) / ((double)(t1 - t0) / 1000000));
	printf("-----------------------------------\n\n");

	for (i = 0; i < thread_no; i++)
		ocl_worker_ctx_free(w[i]);
	return 0;
}
