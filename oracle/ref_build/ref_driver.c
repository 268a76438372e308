/*
 * oracle/_ref driver -- TEST INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Links against the UNMODIFIED reference builder (acsmx.c, compiled in place
 * from /root/reference by oracle/ref_build/Makefile) and exposes, over a plain
 * C ABI for ctypes:
 *   - the tables the reference builds (state count, Lmax, h_trans in the
 *     reference layout int32[num_states][512], acsmx.c:640-659),
 *   - every state's match list in list order, snapshotted before
 *     acsm_cleanup() throws it away (acsmx.c:771-803),
 *   - a serial walk over those tables: the reference ships no
 *     CPU search function, so "reference output" means: the
 *     reference's own DFA + match lists, walked one byte at a time with the
 *     transition rule of ahomatch.cl:56-65.
 *
 * The three OpenCL buffer calls acsm_gen_state_table() makes are defined here
 * as host no-ops (mapped == 0 path, so h_trans is a plain MALLOC).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>

#include "acsmx.h"

#define REF_ROW (2 * ALPHABET_SIZE)

/* ---- OpenCL stand-ins (host only) ---- */
cl_mem
clCreateBuffer(cl_context c, cl_mem_flags f, size_t n, void *p, cl_int *e)
{
	(void)c; (void)f; (void)n; (void)p;
	if (e)
		*e = CL_SUCCESS;
	return NULL;
}

void *
clEnqueueMapBuffer(cl_command_queue q, cl_mem m, cl_bool b, cl_map_flags f,
    size_t o, size_t n, cl_uint ne, const cl_event *w, cl_event *ev, cl_int *e)
{
	(void)q; (void)m; (void)b; (void)f; (void)o; (void)ne; (void)w; (void)ev;
	if (e)
		*e = CL_SUCCESS;
	return malloc(n);
}

cl_int
clEnqueueWriteBuffer(cl_command_queue q, cl_mem m, cl_bool b, size_t o,
    size_t n, const void *p, cl_uint ne, const cl_event *w, cl_event *ev)
{
	(void)q; (void)m; (void)b; (void)o; (void)n; (void)p; (void)ne; (void)w;
	(void)ev;
	return CL_SUCCESS;
}

char *
clstrerror(int err)
{
	static char buf[32];
	snprintf(buf, sizeof(buf), "cl error %d", err);
	return buf;
}

/* ---- handle ---- */
struct ref_handle {
	acsm_t   *acsm;
	int       compiled;
	int       num_states;      /* count (after the +1 of acsmx.c:615) */
	int64_t  *ml_begin;        /* CSR over states, num_states + 1      */
	int32_t  *ml_index;        /* pattern index, list order            */
	int32_t  *ml_iid;
	int32_t  *ml_len;
	int32_t  *pat_len;         /* by pattern index                     */
	int32_t  *pat_iid;
};

struct ref_handle *
ref_new(void)
{
	struct ref_handle *h = calloc(1, sizeof(*h));
	h->acsm = acsm_new();
	return h;
}

void
ref_add(struct ref_handle *h, const unsigned char *pat, int n, int iid)
{
	acsm_add_pattern(h->acsm, (unsigned char *)pat, n, 0, 0, 0, NULL, iid);
}

/* compile, snapshot match lists, serialise (no device) */
void
ref_compile(struct ref_handle *h)
{
	acsm_t *a = h->acsm;
	acsm_pattern_t *p;
	int64_t total = 0, k = 0;
	int s, ns;

	acsm_compile(a);
	ns = a->num_states + 1;

	h->pat_len = calloc(a->num_patterns + 1, sizeof(int32_t));
	h->pat_iid = calloc(a->num_patterns + 1, sizeof(int32_t));
	for (p = a->patterns; p; p = p->next) {
		h->pat_len[p->index] = p->n;
		h->pat_iid[p->index] = p->iid;
	}

	for (s = 0; s < ns; s++)
		for (p = a->state_table[s].match_list; p; p = p->next)
			total++;
	h->ml_begin = calloc(ns + 1, sizeof(int64_t));
	h->ml_index = calloc(total + 1, sizeof(int32_t));
	h->ml_iid   = calloc(total + 1, sizeof(int32_t));
	h->ml_len   = calloc(total + 1, sizeof(int32_t));
	for (s = 0; s < ns; s++) {
		h->ml_begin[s] = k;
		for (p = a->state_table[s].match_list; p; p = p->next) {
			h->ml_index[k] = (int32_t)p->index;
			h->ml_iid[k]   = p->iid;
			h->ml_len[k]   = p->n;
			k++;
		}
	}
	h->ml_begin[ns] = k;

	acsm_gen_state_table(a, 0, NULL, NULL);
	h->num_states = a->num_states;
	h->compiled = 1;
}

int  ref_num_states(struct ref_handle *h)      { return h->num_states; }
int  ref_num_patterns(struct ref_handle *h)    { return h->acsm->num_patterns; }
int  ref_max_pattern_len(struct ref_handle *h) { return acsm_get_max_pattern_size(h->acsm); }
size_t ref_table_bytes(struct ref_handle *h)   { return acsm_get_size(h->acsm); }
const int *ref_h_trans(struct ref_handle *h)   { return h->acsm->h_trans; }
const int64_t *ref_ml_begin(struct ref_handle *h) { return h->ml_begin; }
const int32_t *ref_ml_index(struct ref_handle *h) { return h->ml_index; }
const int32_t *ref_ml_iid(struct ref_handle *h)   { return h->ml_iid; }
const int32_t *ref_pat_len(struct ref_handle *h)  { return h->pat_len; }
const int32_t *ref_pat_iid(struct ref_handle *h)  { return h->pat_iid; }

/*
 * head-only view of the transition a walk takes: what ahomatch.cl would store
 * as the pattern index for a final-state hit (acsmx.c:648-650).
 */
int
ref_head_index(struct ref_handle *h, int prev_state, int c)
{
	return h->acsm->h_trans[(size_t)prev_state * REF_ROW + ALPHABET_SIZE + c];
}

/*
 * Serial walk, full match-list semantics.
 *   start_state : state to start from (0 = cold)
 *   emit_from   : only matches with end offset >= emit_from are recorded
 *   base        : added to every recorded offset
 * Records up to cap (offset, pattern index) pairs in walk order; returns the
 * number of matches found (may exceed cap).  *hits counts final-state
 * transitions, *final_state is the state after the last byte.
 */
int64_t
ref_search(struct ref_handle *h, const unsigned char *text, int64_t n,
    int start_state, int64_t emit_from, uint64_t base, uint64_t *out_off,
    uint32_t *out_pat, int64_t cap, int64_t *hits, int *final_state)
{
	const int *T = h->acsm->h_trans;
	int64_t found = 0, nh = 0, k, e;
	int state = start_state, t;

	for (k = 0; k < n; k++) {
		t = T[(size_t)state * REF_ROW + text[k]];
		if (t < 0) {
			state = -t;
			if (k >= emit_from) {
				nh++;
				for (e = h->ml_begin[state];
				    e < h->ml_begin[state + 1]; e++) {
					if (found < cap) {
						out_off[found] = base + (uint64_t)k;
						out_pat[found] = (uint32_t)h->ml_index[e];
					}
					found++;
				}
			}
		} else {
			state = t;
		}
	}
	if (hits)
		*hits = nh;
	if (final_state)
		*final_state = state;
	return found;
}

/* walk only, no emission: the timed CPU baseline inner loop */
int64_t
ref_walk_count(struct ref_handle *h, const unsigned char *text, int64_t n,
    int64_t emit_from)
{
	const int *T = h->acsm->h_trans;
	int64_t found = 0, k;
	int state = 0, t;

	for (k = 0; k < n; k++) {
		t = T[(size_t)state * REF_ROW + text[k]];
		if (t < 0) {
			state = -t;
			if (k >= emit_from)
				found += h->ml_begin[state + 1] - h->ml_begin[state];
		} else {
			state = t;
		}
	}
	return found;
}

/* ---- pthread-sharded walk ---- */
struct shard_arg {
	struct ref_handle *h;
	const unsigned char *text;
	int64_t lo, hi, halo;
	int64_t found;
};

static void *
shard_main(void *p)
{
	struct shard_arg *a = p;
	int64_t start = a->lo - a->halo;

	if (start < 0)
		start = 0;
	a->found = ref_walk_count(a->h, a->text + start, a->hi - start,
	    a->lo - start);
	return NULL;
}

int64_t
ref_walk_count_mt(struct ref_handle *h, const unsigned char *text, int64_t n,
    int threads)
{
	pthread_t *tid = calloc(threads, sizeof(*tid));
	struct shard_arg *args = calloc(threads, sizeof(*args));
	int64_t total = 0, halo = ref_max_pattern_len(h) - 1;
	int i;

	if (halo < 0)
		halo = 0;
	for (i = 0; i < threads; i++) {
		args[i].h = h;
		args[i].text = text;
		args[i].lo = n * i / threads;
		args[i].hi = n * (i + 1) / threads;
		args[i].halo = halo;
		pthread_create(&tid[i], NULL, shard_main, &args[i]);
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
ref_free(struct ref_handle *h)
{
	if (!h)
		return;
	if (h->compiled)
		acsm_cleanup(h->acsm);
	free(h->acsm->h_trans);
	acsm_free(h->acsm);
	free(h->ml_begin); free(h->ml_index); free(h->ml_iid); free(h->ml_len);
	free(h->pat_len); free(h->pat_iid);
	free(h);
}
