/*
 * worker.c -- worker context and pattern-file loader (host, plain C).
 *
 * Replaces reference ocl_worker.c:21-199 and the two helpers of reference
 * utils.c that the loader needs (hex decoding, utils.c:19-54; clock, utils.c:61-68).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/acm.h"
#include "../../include/ocl_aho_match.h"
#include "../../include/ocl_compact_array.h"
#include "../../include/ocl_prefix_sum.h"
#include "../../include/ocl_worker.h"
#include "../../include/utils.h"
#include "acm_core.h"
#include "databuf_priv.h"

static int
hex_nibble(int ch)
{
	if (ch >= '0' && ch <= '9')
		return ch - '0';
	ch = tolower(ch);
	if (ch >= 'a' && ch <= 'f')
		return ch - 'a' + 10;
	return -1;
}

unsigned char *
printable_hex_to_bytes(unsigned char *input)
{
	const size_t L = strlen((const char *)input);
	unsigned char *out;
	size_t i;

	if (L % 2) {
		acm_set_error("hex pattern has odd length %zu", L);
		return NULL;
	}
	out = calloc(L / 2 + 1, 1);
	if (!out) {
		acm_set_error("printable_hex_to_bytes: out of memory");
		return NULL;
	}
	for (i = 0; i < L; i += 2) {
		const int hi = hex_nibble(input[i]), lo = hex_nibble(input[i + 1]);
		if (hi < 0 || lo < 0) {
			acm_set_error("hex pattern has a non-hex character at %zu", i);
			free(out);
			return NULL;
		}
		out[i / 2] = (unsigned char)(hi * 16 + lo);
	}
	return out;
}

size_t
gettime(void)
{
	struct timespec ts;

	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (size_t)ts.tv_sec * 1000000u + (size_t)ts.tv_nsec / 1000u;
}

/*
 * The grammar (reference ocl_worker.c:74-145): lines of at most MAX_PAT_SIZE-1
 * bytes (longer lines are split by fgets, as there); one trailing newline is
 * stripped; the FIRST line decides the format for the whole file: categorical
 * ("ID pattern", ID = [+-]digits) when the text before its first blank is a
 * number; a pattern wrapped in double quotes loses them; -m truncates; -x decodes
 * hex pairs.  A first line without a blank is plain (the reference's sniff loop
 * has no bound there).
 */
int
acsm_load_pattern_file(acsm_t *acsm, const char *path, int hex_pat, int pat_size_limit)
{
	FILE *fp = fopen(path, "r");
	char line[MAX_PAT_SIZE];
	int added = 0, categ = 0, lineno = 0;

	if (!fp) {
		acm_set_error("cannot open pattern file %s: %s", path, strerror(errno));
		return -1;
	}
	while (fgets(line, sizeof(line), fp)) {
		char *pattern = line;
		size_t len = strlen(line);
		long pat_id = lineno;

		if (len && line[len - 1] == '\n')
			line[--len] = '\0';
		if (lineno == 0) {
			const char *blank = strpbrk(line, " \t");
			if (blank && blank != line) {
				const char *q = line;
				if (*q == '+' || *q == '-')
					q++;
				else if (!isdigit((unsigned char)*q))
					q = NULL;
				if (q) {
					/* a lone sign counts, as in the reference (ocl_worker.c:97-100) */
					while (q < blank && isdigit((unsigned char)*q))
						q++;
					categ = (q == blank);
				}
			}
		}
		if (categ) {
			char *end;
			errno = 0;
			pat_id = strtol(line, &end, 10);
			if (errno == ERANGE) {
				acm_set_error("%s:%d: pattern id out of range", path, lineno + 1);
				fclose(fp);
				return -1;
			}
			while (isspace((unsigned char)*end))
				end++;
			pattern = end;
			len = strlen(pattern);
		}
		if (len >= 1 && pattern[0] == '"' && pattern[len - 1] == '"') {
			pattern[len - 1] = '\0';
			pattern++;
			len = len >= 2 ? len - 2 : 0;
		}
		if (hex_pat) {
			unsigned char *raw;
			if (pat_size_limit != -1 && (size_t)pat_size_limit * 2 < len) {
				pattern[(size_t)pat_size_limit * 2] = '\0';
				len = (size_t)pat_size_limit * 2;
			}
			raw = printable_hex_to_bytes((unsigned char *)pattern);
			if (!raw) {
				fclose(fp);
				return -2;
			}
			acsm_add_pattern(acsm, raw, (int)(len / 2), 0, 0, 0, NULL, (int)pat_id);
			free(raw);              /* acsm_add_pattern copies; the reference leaks this */
		} else {
			if (pat_size_limit != -1 && (size_t)pat_size_limit < len) {
				pattern[pat_size_limit] = '\0';
				len = (size_t)pat_size_limit;
			}
			acsm_add_pattern(acsm, (unsigned char *)pattern, (int)len, 0, 0, 0, NULL, (int)pat_id);
		}
		added++;
		lineno++;
	}
	fclose(fp);
	return added;
}

struct ocl_worker_ctx *
ocl_worker_ctx_create(int dev_pos)
{
	struct ocl_worker_ctx *w = calloc(1, sizeof(*w));

	if (!w)
		return NULL;
	clinitctx(&w->cl, dev_pos, -1);
	if (!w->cl.ctx) {
		free(w);
		return NULL;
	}
	ocl_aho_match_init(&w->cl);
	ocl_prefix_sum_init(&w->cl);
	ocl_compact_array_init(&w->cl);
	return w;
}

int
ocl_worker_ctx_init(struct ocl_worker_ctx *w, int dev_pos, size_t local_ws, size_t global_ws, int mapped,
    char *pat_path, int hex_pat, int pat_size_limit, size_t max_chunk_size, int max_results, int verbose,
    int text_mode, int follow, int id, int thread_no, int total_files, int *fds, char **filenames)
{
	(void)dev_pos;
	if (!w || !pat_path)
		return -1;
	w->acsm = acsm_new();
	if (!w->acsm)
		return -1;
	if (acsm_load_pattern_file(w->acsm, pat_path, hex_pat, pat_size_limit) < 0)
		return -1;
	acsm_compile(w->acsm);
	if (acsm_status(w->acsm) != ACM_OK)
		return -1;
	acsm_gen_state_table(w->acsm, mapped, w->cl.ctx, w->cl.queue);
	if (acsm_status(w->acsm) != ACM_OK)
		return -1;
	w->patterns = acsm_get_patterns_table(w->acsm);
	w->patterns_size = (size_t)w->acsm->num_patterns;
	acsm_cleanup(w->acsm);

	w->db = databuf_new(global_ws, max_chunk_size, max_results, mapped, &w->cl);
	if (!w->db)
		return -1;
	/* scanner and kernel code now, as the reference builds its program here (compat_ocl.c:databuf_prepare) */
	if (databuf_prepare(w->db, acsm_device_automaton(w->acsm), 1) != ACM_OK)
		return -1;

	w->local_ws = local_ws;
	w->global_ws = global_ws;
	w->matches_total = 0;
	w->matches_reported = 0;
	w->bytes = 0;
	w->lines = 0;
	w->rounds = 0;
	w->verbose = verbose;
	w->text_mode = text_mode;
	w->follow = follow;
	w->id = id;
	w->thread_no = thread_no;
	w->total_files = total_files;
	w->fds = fds;
	w->filenames = filenames;
	return 0;
}

void
ocl_worker_ctx_free(struct ocl_worker_ctx *w)
{
	if (!w)
		return;
	if (w->db)
		databuf_free(w->db, w->db->mapped, w->cl.queue);
	if (w->patterns)
		acsm_free_patterns_table(w->patterns, (int)w->patterns_size);
	if (w->acsm)
		acsm_free(w->acsm);
	clfreectx(&w->cl);
	free(w);
}
