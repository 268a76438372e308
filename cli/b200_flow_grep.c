/*
 * b200_flow_grep -- the AC_ushorts application on libacmatch_b200.so: signatures are
 * sequences of packet sizes (unsigned shorts, alphabet 2048), inputs are per-flow files of
 * comma-separated packet sizes named src_sport_dst_dport_proto, output is one alert line per
 * (flow, signature) match.
 *
 * Follows reference AC_ushorts/README:14-38 (signature line: "tokens;length;details", flow
 * file naming and contents), AC_ushorts/ocl_aho_grep.c:259-289 (read_signatures) and :296-345
 * (print_matches: the alert line format).  Written from scratch on iacsmx.h / databuf.h /
 * ocl_aho_match.h.  Every flow is scanned as its own stream: flows are laid out back to back
 * in the buffer, one chunk each, separated by a token outside the alphabet (which resets the
 * automaton), so a signature can never match across two flows -- the reference lets a chunk's
 * state run on into the next chunk (AC_ushorts/ahomatch.cl:100-147).
 *
 *   b200_flow_grep -p signatures -f flow_dir_or_files [-D devpos] [-v]
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "acm.h"
#include "iacsmx.h"
#include "ocl_aho_match.h"

#define SEPARATOR 0xFFFFu          /* >= I_ALPHABET_SIZE: cannot be part of a signature */

struct signature {
	char *tokens;
	int   length;
	char *details;
};

struct flows {
	char **names;
	int n, cap;
};

static struct signature *sigs;
static int nsigs;
static struct flows fl;
static int verbose;
static size_t alerts;

static void
add_flow(const char *path)
{
	if (fl.n == fl.cap) {
		fl.cap = fl.cap ? fl.cap * 2 : 64;
		fl.names = realloc(fl.names, (size_t)fl.cap * sizeof(char *));
	}
	fl.names[fl.n++] = strdup(path);
}

static void
walk(const char *path)
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
		while (d && (de = readdir(d)) != NULL) {
			if (de->d_name[0] == '.')
				continue;
			snprintf(sub, sizeof(sub), "%s/%s", path, de->d_name);
			walk(sub);
		}
		if (d)
			closedir(d);
	} else if (S_ISREG(st.st_mode)) {
		add_flow(path);
	}
}

static int
cmp_names(const void *a, const void *b)
{
	return strcmp(*(char *const *)a, *(char *const *)b);
}

/* "tokens;length;details" per line (reference AC_ushorts/ocl_aho_grep.c:259-289) */
static int
read_signatures(const char *path, iacsm_t *m)
{
	FILE *fp = fopen(path, "r");
	char line[8192], *save, *tok;

	if (!fp)
		return -1;
	while (fgets(line, sizeof(line), fp)) {
		size_t L = strlen(line);
		while (L && (line[L - 1] == '\n' || line[L - 1] == '\r'))
			line[--L] = '\0';
		if (!L)
			continue;
		sigs = realloc(sigs, (size_t)(nsigs + 1) * sizeof(*sigs));
		memset(&sigs[nsigs], 0, sizeof(*sigs));
		tok = strtok_r(line, ";", &save);
		sigs[nsigs].tokens = strdup(tok ? tok : "");
		tok = strtok_r(NULL, ";", &save);
		sigs[nsigs].length = tok ? atoi(tok) : 0;
		tok = strtok_r(NULL, ";", &save);
		while (tok && *tok == ' ')
			tok++;
		sigs[nsigs].details = strdup(tok ? tok : "");
		iacsm_add_fullpattern(m, sigs[nsigs].tokens, nsigs);        /* iid = signature id = line number */
		if (iacsm_status(m) != ACM_OK)
			return -2;
		nsigs++;
	}
	fclose(fp);
	return nsigs;
}

/* one flow file -> ushort tokens appended at dst; returns the count (any of , \n \r space separates) */
static size_t
read_flow(const char *path, unsigned short *dst, size_t room)
{
	FILE *fp = fopen(path, "r");
	size_t n = 0;
	int c, have = 0;
	unsigned long v = 0;

	if (!fp)
		return 0;
	while ((c = fgetc(fp)) != EOF) {
		if (c >= '0' && c <= '9') {
			v = v * 10 + (unsigned long)(c - '0');
			have = 1;
		} else if (have) {
			if (n < room)
				dst[n++] = (unsigned short)(v > 0xFFFE ? 0xFFFE : v);
			v = 0;
			have = 0;
		}
	}
	if (have && n < room)
		dst[n++] = (unsigned short)(v > 0xFFFE ? 0xFFFE : v);
	fclose(fp);
	return n;
}

/* reference AC_ushorts/ocl_aho_grep.c:296-345 */
static int
alert(int file_id, int sig_id, int chunk, int offset, void *uarg)
{
	char tm_buffer[128], *copy, *base, *save;
	const char *src_ip, *src_port, *dst_ip, *dst_port, *proto;
	time_t now = time(NULL);

	(void)chunk;
	(void)offset;
	(void)uarg;
	alerts++;
	if (!verbose)
		return 0;
	strftime(tm_buffer, sizeof(tm_buffer), "date: %Y-%m-%d, time: %H:%M:%S", localtime(&now));
	printf("%s, signature id: %d, signature pattern: '%s', signature length: %d, signature details: '%s', ",
	    tm_buffer, sig_id, sigs[sig_id].tokens, sigs[sig_id].length, sigs[sig_id].details);
	copy = strdup(fl.names[file_id]);
	base = strrchr(copy, '/');
	base = base ? base + 1 : copy;
	src_ip = strtok_r(base, "_", &save);
	src_port = strtok_r(NULL, "_", &save);
	dst_ip = strtok_r(NULL, "_", &save);
	dst_port = strtok_r(NULL, "_", &save);
	proto = strtok_r(NULL, "_", &save);
	printf("source ip: %s, source port: %s, destination ip: %s, destination port: %s, protocol: %s \n\n",
	    src_ip ? src_ip : "?", src_port ? src_port : "?", dst_ip ? dst_ip : "?", dst_port ? dst_port : "?",
	    proto ? proto : "?");
	free(copy);
	return 0;
}

/* the callback receives the pattern INDEX; signatures were added in line order, so index == id */
static void
flush(struct clconf *cl, struct databuf *db, iacsm_t *m)
{
	if (db->chunks == 0)
		return;
	databuf_copy_host_to_device(db, cl->queue);
	ocl_aho_match_ushort(cl, db, m, 1024);
	databuf_copy_device_to_host(db, cl->queue);
	if (databuf_status(db) != ACM_OK) {
		fprintf(stderr, "ERROR: %s\n", acm_last_error());
		exit(1);
	}
	databuf_process_results(db, alert, NULL);
	databuf_reset(db);
}

int
main(int argc, char **argv)
{
	const char *pat_path = NULL, *data_path = NULL;
	size_t max_chunks = 65536, chunk_bytes = 4096, tokens = 0;
	struct clconf cl;
	struct databuf *db;
	iacsm_t *m;
	char *paths, *tok, *save;
	int dev = 0, i;

	for (i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "-p") && i + 1 < argc)
			pat_path = argv[++i];
		else if (!strcmp(argv[i], "-f") && i + 1 < argc)
			data_path = argv[++i];
		else if (!strcmp(argv[i], "-D") && i + 1 < argc)
			dev = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-v"))
			verbose = 1;
		else if ((!strcmp(argv[i], "-B") || !strcmp(argv[i], "-G") || !strcmp(argv[i], "-L") ||
		    !strcmp(argv[i], "-w")) && i + 1 < argc)
			i++;                            /* reference launch knobs: accepted, ignored */
		else {
			fprintf(stderr, "usage: %s -p signatures -f flow_dir_or_files [-D devpos] [-v]\n", argv[0]);
			return 2;
		}
	}
	if (!pat_path || !data_path) {
		fprintf(stderr, "usage: %s -p signatures -f flow_dir_or_files [-D devpos] [-v]\n", argv[0]);
		return 2;
	}
	clinitctx(&cl, dev, -1);
	if (!cl.ctx) {
		fprintf(stderr, "ERROR: %s\n", acm_last_error());
		return 1;
	}
	m = iacsm_new();
	if (read_signatures(pat_path, m) <= 0) {
		fprintf(stderr, "ERROR: cannot read signatures from %s: %s\n", pat_path, acm_last_error());
		return 1;
	}
	iacsm_compile(m);
	iacsm_gen_state_table(m, 0, cl.ctx, cl.queue);
	if (iacsm_status(m) != ACM_OK) {
		fprintf(stderr, "ERROR: %s\n", acm_last_error());
		return 1;
	}
	paths = strdup(data_path);
	for (tok = strtok_r(paths, ",", &save); tok; tok = strtok_r(NULL, ",", &save))
		walk(tok);
	qsort(fl.names, (size_t)fl.n, sizeof(char *), cmp_names);

	db = databuf_new(max_chunks, chunk_bytes, MAX_RESULTS, 0, &cl);
	if (!db) {
		fprintf(stderr, "ERROR: %s\n", acm_last_error());
		return 1;
	}
	/* one chunk per flow, laid out back to back as ushorts, a separator token after each */
	for (i = 0; i < fl.n; i++) {
		unsigned short *dst;
		size_t room, n;

		if (db->chunks >= db->max_chunks || db->bytes + 64 >= db->size)
			flush(&cl, db, m);
		dst = (unsigned short *)(db->h_data + db->bytes);
		room = (db->size - db->bytes) / 2 - 1;
		n = read_flow(fl.names[i], dst, room);
		if (n == room) {                        /* flow did not fit: scan what is buffered, retry alone */
			if (db->chunks) {
				flush(&cl, db, m);
				i--;
				continue;
			}
			fprintf(stderr, "WARNING: flow %s truncated to %zu tokens\n", fl.names[i], n);
		}
		dst[n] = SEPARATOR;
		db->h_indices[db->chunks] = (int)db->bytes;
		db->h_sizes[db->chunks] = (int)(n * 2);
		db->file_ids[db->chunks] = i;
		db->chunks++;
		db->bytes += (n + 1) * 2;
		tokens += n;
	}
	flush(&cl, db, m);
	printf("-------------- STATS --------------\n");
	printf("Signatures:          %d\n", nsigs);
	printf("Automaton states:    %d\n", iacsm_get_states(m));
	printf("Processed flows:     %d\n", fl.n);
	printf("Processed tokens:    %zu\n", tokens);
	printf("Alerts:              %zu\n", alerts);
	printf("-----------------------------------\n");
	databuf_free(db, 0, cl.queue);
	iacsm_free(m);
	clfreectx(&cl);
	return 0;
}
