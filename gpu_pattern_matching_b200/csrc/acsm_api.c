/*
 * acsm_api.c -- the acsmx.h entry points (reference acsmx.c:489-815) over the
 * array-based builder in acm_core.c and the device upload in acm_cuda.cu.
 */
#define _GNU_SOURCE
#include <malloc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/acm.h"
#include "../../include/acsmx.h"
#include "acm_core.h"
#include "acm_queue.h"

static struct acm_core *
core_of(acsm_t *a)
{
	return a ? (struct acm_core *)a->priv : NULL;
}

acsm_t *
acsm_new(void)
{
	acsm_t *a = calloc(1, sizeof(*a));

	if (!a)
		return NULL;
	a->priv = acm_core_new(ALPHABET_SIZE);
	if (!a->priv) {
		free(a);
		return NULL;
	}
	return a;
}

void
acsm_add_pattern(acsm_t *a, unsigned char *pat, int n, int nocase, int offset, int depth, void *id,
    int iid)
{
	struct acm_core *c = core_of(a);

	if (!c)
		return;
	if (acm_core_add(c, pat, n, nocase, offset, depth, id, iid) == ACM_OK ||
	    c->status == ACM_ERR_EMPTY_PATTERN) {
		a->num_patterns = c->npats;
		a->max_pattern_len = c->max_len;
	}
}

void
acsm_compile(acsm_t *a)
{
	struct acm_core *c = core_of(a);

	if (!c)
		return;
	if (acm_core_compile(c) != ACM_OK)
		return;
	c->status = ACM_OK;
	/* the reference counts one row per pattern byte plus the root (acsmx.c:560-563) */
	a->max_states = 1;
	for (int k = 0; k < c->npats; k++)
		a->max_states += c->pats[k].n;
	/* highest state id; acsm_gen_state_table() turns it into a count (acsmx.c:615) */
	a->num_states = (int)c->tab.num_states - 1;
}

void
acsm_gen_state_table(acsm_t *a, int mapped, cl_context ctx, cl_command_queue queue)
{
	struct acm_core *c = core_of(a);
	struct acm_device *dev = acm_queue_device(ctx, queue);
	int rc;

	(void)mapped;
	if (!c)
		return;
	if (!c->compiled) {
		acm_set_error("acsm_gen_state_table: call acsm_compile first");
		c->status = ACM_ERR_STATE;
		return;
	}
	if (c->dev)
		return;                     /* already uploaded */
	a->num_states = (int)c->tab.num_states;
	if (!dev) {
		c->status = ACM_ERR_NO_DEVICE;
		return;
	}
	rc = acm_automaton_upload(dev, &c->tab, &c->dev);
	c->status = rc;
	if (rc != ACM_OK)
		return;
	a->size = acm_automaton_device_bytes(c->dev);
	a->d_trans = (cl_mem)c->dev;
}

acsm_pattern_t *
acsm_get_patterns_table(acsm_t *a)
{
	struct acm_core *c = core_of(a);
	acsm_pattern_t *tab;

	if (!c || c->npats == 0)
		return NULL;
	tab = calloc((size_t)c->npats, sizeof(*tab));
	if (!tab)
		return NULL;
	for (int k = 0; k < c->npats; k++) {
		const struct acm_pat *p = &c->pats[k];
		if (!p->syms) {             /* after acsm_cleanup() */
			for (int j = 0; j < k; j++) {
				free(tab[j].pattern);
				free(tab[j].casepattern);
			}
			free(tab);
			acm_set_error("acsm_get_patterns_table: call it before acsm_cleanup");
			return NULL;
		}
		tab[k].pattern = malloc((size_t)p->n + 1);
		tab[k].casepattern = malloc((size_t)p->n + 1);
		if (tab[k].pattern) {
			memcpy(tab[k].pattern, p->syms, (size_t)p->n);
			tab[k].pattern[p->n] = '\0';
		}
		if (tab[k].casepattern) {
			memcpy(tab[k].casepattern, p->syms, (size_t)p->n);
			tab[k].casepattern[p->n] = '\0';
		}
		tab[k].n = p->n;
		tab[k].nocase = p->nocase;
		tab[k].offset = p->offset;
		tab[k].depth = p->depth;
		tab[k].id = p->id;
		tab[k].iid = p->iid;
		tab[k].index = (unsigned int)k;
		tab[k].next = NULL;
	}
	return tab;
}

void
acsm_free_patterns_table(acsm_pattern_t *tab, int num_patterns)
{
	if (!tab)
		return;
	for (int k = 0; k < num_patterns; k++) {
		free(tab[k].pattern);
		free(tab[k].casepattern);
	}
	free(tab);
}

int
acsm_get_max_pattern_size(acsm_t *a)
{
	return a->max_pattern_len;
}

int
acsm_get_min_pattern_size(acsm_t *a)
{
	struct acm_core *c = core_of(a);
	return c ? c->min_len : 0;
}

int
acsm_get_states(acsm_t *a)
{
	return a->num_states;
}

size_t
acsm_get_size(acsm_t *a)
{
	return a->size;
}

int
acsm_status(acsm_t *a)
{
	struct acm_core *c = core_of(a);
	return c ? c->status : ACM_ERR_ARG;
}

int
acsm_check_filters(acsm_t *a)
{
	struct acm_core *c = core_of(a);
	return c ? acm_core_check_filters(c) : ACM_ERR_ARG;
}

int
acsm_check_xd(acsm_t *a, unsigned int *slots)
{
	struct acm_core *c = core_of(a);
	return c ? acm_core_check_xd(c, slots) : ACM_ERR_ARG;
}

const struct acm_tables *
acsm_tables(acsm_t *a)
{
	struct acm_core *c = core_of(a);

	return c && c->compiled && c->tab.T ? &c->tab : NULL;
}

int
acsm_check_cdfa(acsm_t *a, unsigned int *slots, unsigned int *dense_rows)
{
	struct acm_core *c = core_of(a);
	return c ? acm_core_check_rd(c, slots, dense_rows) : ACM_ERR_ARG;
}

int
acsm_export_ref_table(acsm_t *a)
{
	struct acm_core *c = core_of(a);
	int *tab = NULL;
	int rc;

	if (!c)
		return ACM_ERR_ARG;
	if (a->h_trans)
		return ACM_OK;
	rc = acm_core_export_ref(c, &tab);
	if (rc == ACM_OK)
		a->h_trans = tab;
	return rc;
}

struct acm_automaton *
acsm_device_automaton(acsm_t *a)
{
	struct acm_core *c = core_of(a);
	return c ? c->dev : NULL;
}

void
acsm_cleanup(acsm_t *a)
{
	struct acm_core *c = core_of(a);

	if (!c)
		return;
	acm_core_cleanup(c);
	a->patterns = NULL;
	a->state_table = NULL;
}

void
acsm_free(acsm_t *a)
{
	struct acm_core *c = core_of(a);

	if (!a)
		return;
	if (c) {
		if (c->dev)
			acm_automaton_free(c->dev);
		acm_core_free(c);
	}
	free(a->h_trans);
	free(a);
}
