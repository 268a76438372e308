/*
 * oracle/acsm_oracle.c -- CPU restatement of the reference's Aho-Corasick path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (gpu_pattern_matching_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement, table
 * for table and match for match, against oracle/_ref/libacref.so, which is the
 * reference's own acsmx.c / iacsmx.c compiled unmodified (oracle/ref_build), on
 * every fixture the reference ships, and against the known-answer table of
 * SURVEY.md section 4 (committed under tests/golden/).
 *
 * What is restated, and from where (paths relative to /root/reference):
 *   pattern list, index = add order, list is prepended      acsmx.c:514-546
 *   trie insertion in list order (= reverse add order),
 *     states numbered in creation order from 1              acsmx.c:318-349, 579-580
 *   root self-loops                                         acsmx.c:583-585
 *   BFS fail links + match-list inheritance by prepending
 *     copies of the fail state's list, head to tail         acsmx.c:355-438
 *   BFS NFA->DFA fill                                       acsmx.c:444-486
 *   serialised table: sign marks "target has a match list",
 *     second half of the row holds the list head's index    acsmx.c:640-659
 *   walk: state = T[state][byte]; negative => match          ahomatch.cl:56-65
 *   match semantics: every entry of the target's match list  SURVEY.md A.3
 *   ushort twin: alphabet 2048, head is the iid              AC_ushorts/iacsmx.c:158-520
 *   pattern-file grammar                                    ocl_worker.c:74-145
 *   hex decoding                                            utils.c:19-54
 *
 * Everything here is array based (no per-node malloc) but the observable
 * numbering and list orders are the reference's.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct orc_pat {
	unsigned char  *bytes;     /* byte patterns                          */
	unsigned short *syms;      /* ushort patterns                        */
	int             n;
	int             iid;
};

struct orc {
	int             alpha;     /* 256 or 2048                            */
	struct orc_pat *pats;      /* by index (add order)                   */
	int             npats, cap_pats;
	int             max_pattern_len;
	int             max_states;
	int             num_states; /* count, i.e. highest id + 1            */
	int            *next;      /* [max_states][alpha] dense              */
	int            *fail;
	/* match lists as singly linked entries while building */
	int            *ml_head;   /* per state, -1 = empty                  */
	int            *e_pat;     /* entry -> pattern index                 */
	int            *e_next;    /* entry -> next entry                    */
	int64_t         e_count, e_cap;
	/* frozen CSR */
	int64_t        *ml_begin;
	int32_t        *ml_index;
	int             compiled;
	int            *ref_table; /* lazily built reference-layout table    */
};

struct orc *
orc_new(int alphabet)
{
	struct orc *o = calloc(1, sizeof(*o));
	o->alpha = alphabet;
	return o;
}

static void
orc_push(struct orc *o, struct orc_pat p)
{
	if (o->npats == o->cap_pats) {
		o->cap_pats = o->cap_pats ? o->cap_pats * 2 : 1024;
		o->pats = realloc(o->pats, o->cap_pats * sizeof(*o->pats));
	}
	o->pats[o->npats++] = p;
	if (p.n > o->max_pattern_len)
		o->max_pattern_len = p.n;
}

/* acsmx.c:514-546 -- bytes are copied, index = number of patterns so far */
void
orc_add(struct orc *o, const unsigned char *pat, int n, int iid)
{
	struct orc_pat p = {0};
	p.bytes = malloc(n ? n : 1);
	memcpy(p.bytes, pat, n);
	p.n = n;
	p.iid = iid;
	orc_push(o, p);
}

/* AC_ushorts/iacsmx.c:390-415 (without its under-allocation) */
void
orc_add_syms(struct orc *o, const unsigned short *items, int n, int iid)
{
	struct orc_pat p = {0};
	p.syms = malloc((n ? n : 1) * sizeof(unsigned short));
	memcpy(p.syms, items, n * sizeof(unsigned short));
	p.n = n;
	p.iid = iid;
	orc_push(o, p);
}

/* AC_ushorts/iacsmx.c:418-452 -- "40,32,287" -> ushorts, atoi per field */
void
orc_add_csv(struct orc *o, const char *csv, int iid)
{
	unsigned short items[4096];
	char field[64];
	size_t i, L = strlen(csv);
	int j = 0, len = 0;

	for (i = 0; i < L + 1; i++) {
		char c = csv[i];
		if (c == ',' || c == '\n' || c == '\0' || c == '\r') {
			field[j] = '\0';
			if (len < 4096)
				items[len++] = (unsigned short)atoi(field);
			j = 0;
		} else if (j < 63) {
			field[j++] = c;
		}
	}
	orc_add_syms(o, items, len, iid);
}

static inline int
pat_sym(const struct orc_pat *p, int k)
{
	return p->bytes ? p->bytes[k] : p->syms[k];
}

static void
ml_prepend(struct orc *o, int state, int pat)
{
	if (o->e_count == o->e_cap) {
		o->e_cap = o->e_cap ? o->e_cap * 2 : 4096;
		o->e_pat = realloc(o->e_pat, o->e_cap * sizeof(int));
		o->e_next = realloc(o->e_next, o->e_cap * sizeof(int));
	}
	o->e_pat[o->e_count] = pat;
	o->e_next[o->e_count] = o->ml_head[state];
	o->ml_head[state] = (int)o->e_count;
	o->e_count++;
}

/* acsmx.c:552-594 */
void
orc_compile(struct orc *o)
{
	const int A = o->alpha;
	int *queue, qh = 0, qt = 0;
	int i, k, r, s, fs, nx, last;
	int64_t total;

	o->max_states = 1;
	for (k = 0; k < o->npats; k++)
		o->max_states += o->pats[k].n;
	o->next = malloc((size_t)o->max_states * A * sizeof(int));
	for (size_t z = 0; z < (size_t)o->max_states * A; z++)
		o->next[z] = -1;
	o->fail = calloc(o->max_states, sizeof(int));
	o->ml_head = malloc(o->max_states * sizeof(int));
	for (i = 0; i < o->max_states; i++)
		o->ml_head[i] = -1;

	/*
	 * The pattern list is prepended at add time (acsmx.c:536-538), so the
	 * compile loop (acsmx.c:579-580) meets the LAST added pattern first.
	 */
	last = 0;
	for (k = o->npats - 1; k >= 0; k--) {
		const struct orc_pat *p = &o->pats[k];
		int state = 0, pos = 0;
		/* follow existing edges (acsmx.c:331-336) */
		for (; pos < p->n; pos++) {
			nx = o->next[(size_t)state * A + pat_sym(p, pos)];
			if (nx == -1)
				break;
			state = nx;
		}
		/* one new state per remaining symbol (acsmx.c:339-344) */
		for (; pos < p->n; pos++) {
			last++;
			o->next[(size_t)state * A + pat_sym(p, pos)] = last;
			state = last;
		}
		ml_prepend(o, state, k);              /* acsmx.c:346, 300-312 */
	}
	o->num_states = last + 1;

	for (i = 0; i < A; i++)                       /* acsmx.c:583-585 */
		if (o->next[i] == -1)
			o->next[i] = 0;

	queue = malloc((size_t)o->max_states * sizeof(int));

	/* build_NFA, acsmx.c:355-438 */
	for (i = 0; i < A; i++) {
		s = o->next[i];
		if (s) {
			queue[qt++] = s;
			o->fail[s] = 0;
		}
	}
	while (qh < qt) {
		r = queue[qh++];
		for (i = 0; i < A; i++) {
			s = o->next[(size_t)r * A + i];
			if (s == -1)
				continue;
			queue[qt++] = s;
			fs = o->fail[r];
			while ((nx = o->next[(size_t)fs * A + i]) == -1)
				fs = o->fail[fs];
			o->fail[s] = nx;
			/* copy nx's list onto s, head to tail, each prepended */
			for (int e = o->ml_head[nx]; e != -1; e = o->e_next[e])
				ml_prepend(o, s, o->e_pat[e]);
		}
	}

	/* convert_NFA_to_DFA, acsmx.c:444-486 */
	qh = qt = 0;
	for (i = 0; i < A; i++) {
		s = o->next[i];
		if (s)
			queue[qt++] = s;
	}
	while (qh < qt) {
		r = queue[qh++];
		for (i = 0; i < A; i++) {
			s = o->next[(size_t)r * A + i];
			if (s != -1)
				queue[qt++] = s;
			else
				o->next[(size_t)r * A + i] =
				    o->next[(size_t)o->fail[r] * A + i];
		}
	}
	free(queue);

	/* freeze the lists into CSR, list order preserved */
	total = o->e_count;
	o->ml_begin = calloc(o->num_states + 1, sizeof(int64_t));
	o->ml_index = calloc(total + 1, sizeof(int32_t));
	total = 0;
	for (s = 0; s < o->num_states; s++) {
		o->ml_begin[s] = total;
		for (int e = o->ml_head[s]; e != -1; e = o->e_next[e])
			o->ml_index[total++] = o->e_pat[e];
	}
	o->ml_begin[o->num_states] = total;
	o->compiled = 1;
}

int  orc_num_states(struct orc *o)      { return o->num_states; }
int  orc_num_patterns(struct orc *o)    { return o->npats; }
int  orc_max_pattern_len(struct orc *o) { return o->max_pattern_len; }
const int64_t *orc_ml_begin(struct orc *o) { return o->ml_begin; }
const int32_t *orc_ml_index(struct orc *o) { return o->ml_index; }
int  orc_pat_len(struct orc *o, int idx) { return o->pats[idx].n; }
int  orc_pat_iid(struct orc *o, int idx) { return o->pats[idx].iid; }
/* size of the reference's serialised table, acsmx.c:660 */
size_t orc_table_bytes(struct orc *o)
{
	return (size_t)o->num_states * 2 * o->alpha * sizeof(int);
}

/*
 * Reference-layout table (acsmx.c:640-659; ushort twin iacsmx.c:491-510).
 * Cells of the second half that the reference leaves uninitialised are 0 here;
 * comparisons must mask them with the sign of the first half.
 */
const int *
orc_ref_table(struct orc *o)
{
	const int A = o->alpha;
	int i, j, st;

	if (o->ref_table)
		return o->ref_table;
	o->ref_table = calloc((size_t)o->num_states * 2 * A, sizeof(int));
	for (i = 0; i < o->num_states; i++) {
		for (j = 0; j < A; j++) {
			st = o->next[(size_t)i * A + j];
			if (o->ml_begin[st + 1] > o->ml_begin[st]) {
				int head = o->ml_index[o->ml_begin[st]];
				o->ref_table[(size_t)i * 2 * A + j] = -st;
				o->ref_table[(size_t)i * 2 * A + A + j] =
				    (A == 256) ? head : o->pats[head].iid;
			} else {
				o->ref_table[(size_t)i * 2 * A + j] = st;
			}
		}
	}
	return o->ref_table;
}

/*
 * Serial walk, SURVEY.md A.3.  Same contract as ref_search() in
 * oracle/ref_build/ref_driver.c.  text is bytes (alpha 256) or ushorts.
 */
int64_t
orc_search(struct orc *o, const void *text, int64_t n, int start_state,
    int64_t emit_from, uint64_t base, uint64_t *out_off, uint32_t *out_pat,
    int64_t cap, int64_t *hits, int *final_state)
{
	const int A = o->alpha;
	const unsigned char *tb = text;
	const unsigned short *ts = text;
	int64_t found = 0, nh = 0, k, e;
	int state = start_state, c;

	for (k = 0; k < n; k++) {
		c = (A == 256) ? tb[k] : ts[k];
		if (c >= A) {                 /* symbol outside the alphabet */
			state = 0;
			continue;
		}
		state = o->next[(size_t)state * A + c];
		if (o->ml_begin[state + 1] > o->ml_begin[state] && k >= emit_from) {
			nh++;
			for (e = o->ml_begin[state]; e < o->ml_begin[state + 1]; e++) {
				if (found < cap) {
					out_off[found] = base + (uint64_t)k;
					out_pat[found] = (uint32_t)o->ml_index[e];
				}
				found++;
			}
		}
	}
	if (hits)
		*hits = nh;
	if (final_state)
		*final_state = state;
	return found;
}

/*
 * The timed CPU baseline: the same walk over the REFERENCE-LAYOUT table
 * (int32[num_states][512], sign = match) so the memory footprint and access
 * pattern are the reference's.  Counts matches
 * with the full match-list rule; no records are stored.
 */
int64_t
orc_walk_count(struct orc *o, const unsigned char *text, int64_t n,
    int64_t emit_from)
{
	const int *T = orc_ref_table(o);
	const size_t row = 2 * (size_t)o->alpha;
	int64_t found = 0, k;
	int state = 0, t;

	for (k = 0; k < n; k++) {
		t = T[state * row + text[k]];
		if (t < 0) {
			state = -t;
			if (k >= emit_from)
				found += o->ml_begin[state + 1] - o->ml_begin[state];
		} else {
			state = t;
		}
	}
	return found;
}

struct orc_shard {
	struct orc *o;
	const unsigned char *text;
	int64_t lo, hi, halo, found;
};

static void *
orc_shard_main(void *p)
{
	struct orc_shard *a = p;
	int64_t start = a->lo - a->halo;

	if (start < 0)
		start = 0;
	a->found = orc_walk_count(a->o, a->text + start, a->hi - start,
	    a->lo - start);
	return NULL;
}

/*
 * nproc-way version: contiguous shards, leading halo of Lmax-1 bytes, emit only
 * matches that end inside the shard.
 */
int64_t
orc_walk_count_mt(struct orc *o, const unsigned char *text, int64_t n,
    int threads)
{
	pthread_t *tid = calloc(threads, sizeof(*tid));
	struct orc_shard *args = calloc(threads, sizeof(*args));
	int64_t total = 0, halo = o->max_pattern_len - 1;
	int i;

	(void)orc_ref_table(o);          /* build once, before the threads */
	if (halo < 0)
		halo = 0;
	for (i = 0; i < threads; i++) {
		args[i].o = o;
		args[i].text = text;
		args[i].lo = n * i / threads;
		args[i].hi = n * (i + 1) / threads;
		args[i].halo = halo;
		pthread_create(&tid[i], NULL, orc_shard_main, &args[i]);
	}
	for (i = 0; i < threads; i++) {
		pthread_join(tid[i], NULL);
		total += args[i].found;
	}
	free(tid);
	free(args);
	return total;
}

void
orc_free(struct orc *o)
{
	int k;

	if (!o)
		return;
	for (k = 0; k < o->npats; k++) {
		free(o->pats[k].bytes);
		free(o->pats[k].syms);
	}
	free(o->pats); free(o->next); free(o->fail); free(o->ml_head);
	free(o->e_pat); free(o->e_next); free(o->ml_begin); free(o->ml_index);
	free(o->ref_table);
	free(o);
}

/* ---------------- pattern files (ocl_worker.c:74-145, utils.c:19-54) ------- */

static int
hexval(int c)
{
	if (isdigit(c))
		return c - '0';
	c = tolower(c);
	if (c >= 'a' && c <= 'f')
		return c + 10 - 'a';
	return 0;     /* the reference returns garbage here; 0 is as good */
}

#define ORC_LINE 4096   /* MAX_PAT_SIZE, utils.h:14 */

/*
 * Returns the number of patterns added, or -1 (file cannot be opened, strtol
 * range error) / -2 (odd-length hex line; the reference exit()s there).
 */
int
orc_load_pattern_file(struct orc *o, const char *path, int hex_pat,
    int pat_size_limit)
{
	FILE *fp = fopen(path, "r");
	char line[ORC_LINE];
	unsigned char raw[ORC_LINE];
	char *pattern, *end;
	long pat_id;
	size_t len;
	int i = 0, categ = 0, j, first_blank;

	if (!fp)
		return -1;
	while (fgets(line, sizeof(line), fp)) {
		len = strlen(line);
		if (len && line[len - 1] == '\n')
			line[len - 1] = '\0';
		if (i == 0) {
			/*
			 * Categorical iff the text before the first blank is
			 * [+-]?digits (ocl_worker.c:79-102).  No blank on line 0:
			 * the reference's sniff loop has no bound; treated as
			 * non-categorical here.
			 */
			first_blank = -1;
			for (j = 0; line[j]; j++)
				if (line[j] == ' ' || line[j] == '\t') {
					first_blank = j;
					break;
				}
			categ = 0;
			if (first_blank > 0) {
				for (j = first_blank - 1; j > 0; j--)
					if (!isdigit((unsigned char)line[j]))
						break;
				if (j == 0 && (line[0] == '+' || line[0] == '-' ||
				    isdigit((unsigned char)line[0])))
					categ = 1;
			}
		}
		if (categ) {
			errno = 0;
			pat_id = strtol(line, &end, 10);
			if ((errno == ERANGE &&
			    (pat_id == LONG_MAX || pat_id == LONG_MIN)) ||
			    (errno != 0 && pat_id == 0)) {
				fclose(fp);
				return -1;
			}
			while (isspace((unsigned char)*end))
				end++;
			pattern = end;
		} else {
			pattern = line;
			pat_id = i;
		}
		len = strlen(pattern);
		if (len >= 1 && pattern[0] == '"' && pattern[len - 1] == '"') {
			pattern[len - 1] = '\0';
			pattern++;
			len = (len >= 2) ? len - 2 : 0;
		}
		if (hex_pat) {
			if (pat_size_limit != -1 && (size_t)pat_size_limit * 2 < len)
				pattern[pat_size_limit * 2] = '\0';
			len = strlen(pattern);
			if (len % 2) {
				fclose(fp);
				return -2;
			}
			for (j = 0; j < (int)len; j += 2)
				raw[j / 2] = (unsigned char)(hexval(pattern[j]) * 16 +
				    hexval(pattern[j + 1]));
			orc_add(o, raw, (int)len / 2, (int)pat_id);
		} else {
			if (pat_size_limit != -1 && (size_t)pat_size_limit < len)
				pattern[pat_size_limit] = '\0';
			orc_add(o, (unsigned char *)pattern, (int)strlen(pattern),
			    (int)pat_id);
		}
		i++;
	}
	fclose(fp);
	return i;
}

/* accessors used by the tests to feed the same patterns to oracle/_ref */
const unsigned char *orc_pat_bytes(struct orc *o, int idx) { return o->pats[idx].bytes; }
const unsigned short *orc_pat_syms(struct orc *o, int idx) { return o->pats[idx].syms; }
const int *orc_next_table(struct orc *o) { return o->next; }
const int *orc_fail_table(struct orc *o) { return o->fail; }
