/*
 * oracle/_ref driver for the ushort-symbol builder -- TEST INFRASTRUCTURE ONLY.
 *
 * Wraps the UNMODIFIED reference AC_ushorts/iacsmx.c (alphabet 2048,
 * iacsmx.h:43) the same way ref_driver.c wraps acsmx.c: build, snapshot the
 * match lists (iid, not index: iacsmx.c:504), serialise, serial walk.
 *
 * Known reference defect kept in mind by the tests: iacsm_add_pattern()
 * allocates sizeof(int) bytes for the pattern copy (iacsmx.c:398), so only
 * short patterns (the shipped fixtures use <= 3 tokens) are safe to feed it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "iacsmx.h"

#define IREF_ROW (2 * I_ALPHABET_SIZE)

struct iref_handle {
	iacsm_t  *m;
	int       num_states;
	int64_t  *ml_begin;
	int32_t  *ml_iid;
	int32_t  *ml_len;
};

struct iref_handle *
iref_new(void)
{
	struct iref_handle *h = calloc(1, sizeof(*h));
	h->m = iacsm_new();
	return h;
}

void
iref_add(struct iref_handle *h, const unsigned short *items, int len, int iid)
{
	iacsm_add_pattern(h->m, (unsigned short *)items, len, 0, 0, NULL, iid);
}

void
iref_add_csv(struct iref_handle *h, const char *csv, int iid)
{
	iacsm_add_fullpattern(h->m, csv, iid);
}

void
iref_compile(struct iref_handle *h)
{
	iacsm_t *m = h->m;
	iacsm_pattern_t *p;
	int64_t total = 0, k = 0;
	int s, ns;

	iacsm_compile(m);
	ns = m->num_states + 1;
	for (s = 0; s < ns; s++)
		for (p = m->state_table[s].match_list; p; p = p->next)
			total++;
	h->ml_begin = calloc(ns + 1, sizeof(int64_t));
	h->ml_iid   = calloc(total + 1, sizeof(int32_t));
	h->ml_len   = calloc(total + 1, sizeof(int32_t));
	for (s = 0; s < ns; s++) {
		h->ml_begin[s] = k;
		for (p = m->state_table[s].match_list; p; p = p->next) {
			h->ml_iid[k] = p->iid;
			h->ml_len[k] = p->n;
			k++;
		}
	}
	h->ml_begin[ns] = k;
	iacsm_gen_state_table(m, 0, NULL, NULL);
	h->num_states = m->num_states;
}

int  iref_num_states(struct iref_handle *h)      { return h->num_states; }
int  iref_max_pattern_len(struct iref_handle *h) { return iacsm_get_max_pattern_size(h->m); }
size_t iref_table_bytes(struct iref_handle *h)   { return iacsm_get_size(h->m); }
const int *iref_h_trans(struct iref_handle *h)   { return h->m->h_trans; }
const int64_t *iref_ml_begin(struct iref_handle *h) { return h->ml_begin; }
const int32_t *iref_ml_iid(struct iref_handle *h)   { return h->ml_iid; }

/* serial walk over ushort tokens; records (token offset, iid) */
int64_t
iref_search(struct iref_handle *h, const unsigned short *text, int64_t n,
    uint64_t *out_off, int32_t *out_iid, int64_t cap, int *final_state)
{
	const int *T = h->m->h_trans;
	int64_t found = 0, k, e;
	int state = 0, t;

	for (k = 0; k < n; k++) {
		if (text[k] >= I_ALPHABET_SIZE) {       /* outside the alphabet */
			state = 0;
			continue;
		}
		t = T[(size_t)state * IREF_ROW + text[k]];
		if (t < 0) {
			state = -t;
			for (e = h->ml_begin[state]; e < h->ml_begin[state + 1];
			    e++) {
				if (found < cap) {
					out_off[found] = (uint64_t)k;
					out_iid[found] = h->ml_iid[e];
				}
				found++;
			}
		} else {
			state = t;
		}
	}
	if (final_state)
		*final_state = state;
	return found;
}

void
iref_free(struct iref_handle *h)
{
	if (!h)
		return;
	free(h->m->h_trans);
	iacsm_free(h->m);
	free(h->ml_begin); free(h->ml_iid); free(h->ml_len);
	free(h);
}
