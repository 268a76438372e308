/*
 * iacsm_api.c -- the iacsmx.h entry points (reference AC_ushorts/iacsmx.c:158-599)
 * over the alphabet-generic builder.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/acm.h"
#include "../../include/iacsmx.h"
#include "acm_core.h"
#include "acm_queue.h"

#define IACSM_MAX_ITEMS 4096

static struct acm_core *
core_of(iacsm_t *m)
{
	return m ? (struct acm_core *)m->priv : NULL;
}

iacsm_t *
iacsm_new(void)
{
	iacsm_t *m = calloc(1, sizeof(*m));

	if (!m)
		return NULL;
	m->priv = acm_core_new(I_ALPHABET_SIZE);
	if (!m->priv) {
		free(m);
		return NULL;
	}
	return m;
}

void
iacsm_add_pattern(iacsm_t *m, unsigned short *items, int len, int offset, int depth, void *id, int iid)
{
	struct acm_core *c = core_of(m);

	if (!c)
		return;
	if (acm_core_add(c, items, len, 0, offset, depth, id, iid) == ACM_OK)
		m->max_pattern_len = c->max_len;
}

/* comma separated decimal items; also stops at CR / LF like the reference parser */
void
iacsm_add_fullpattern(iacsm_t *m, const char *pattern, int np)
{
	unsigned short items[IACSM_MAX_ITEMS];
	char field[32];
	int len = 0, j = 0;
	size_t i, L = strlen(pattern);

	for (i = 0; i <= L; i++) {
		const char ch = pattern[i];
		if (ch == ',' || ch == '\n' || ch == '\r' || ch == '\0') {
			field[j] = '\0';
			if (len < IACSM_MAX_ITEMS)
				items[len++] = (unsigned short)atoi(field);
			j = 0;
		} else if (j < (int)sizeof(field) - 1) {
			field[j++] = ch;
		}
	}
	iacsm_add_pattern(m, items, len, 0, 0, NULL, np);
}

void
iacsm_compile(iacsm_t *m)
{
	struct acm_core *c = core_of(m);

	if (!c || acm_core_compile(c) != ACM_OK)
		return;
	c->status = ACM_OK;
	m->max_states = 1;
	for (int k = 0; k < c->npats; k++)
		m->max_states += c->pats[k].n;
	m->num_states = (int)c->tab.num_states - 1;
}

void
iacsm_gen_state_table(iacsm_t *m, int mapped, cl_context ctx, cl_command_queue queue)
{
	struct acm_core *c = core_of(m);
	struct acm_device *dev = acm_queue_device(ctx, queue);

	(void)mapped;
	if (!c)
		return;
	if (!c->compiled) {
		acm_set_error("iacsm_gen_state_table: call iacsm_compile first");
		c->status = ACM_ERR_STATE;
		return;
	}
	if (c->dev)
		return;
	m->num_states = (int)c->tab.num_states;
	if (!dev) {
		c->status = ACM_ERR_NO_DEVICE;
		return;
	}
	c->status = acm_automaton_upload(dev, &c->tab, &c->dev);
	if (c->status != ACM_OK)
		return;
	m->size = acm_automaton_device_bytes(c->dev);
	m->d_trans = (cl_mem)c->dev;
}

int    iacsm_get_max_pattern_size(iacsm_t *m) { return m->max_pattern_len; }
int    iacsm_get_states(iacsm_t *m) { return m->num_states; }
size_t iacsm_get_size(iacsm_t *m) { return m->size; }

int
iacsm_status(iacsm_t *m)
{
	struct acm_core *c = core_of(m);
	return c ? c->status : ACM_ERR_ARG;
}

int
iacsm_export_ref_table(iacsm_t *m)
{
	struct acm_core *c = core_of(m);
	int *tab = NULL;
	int rc;

	if (!c)
		return ACM_ERR_ARG;
	if (m->h_trans)
		return ACM_OK;
	rc = acm_core_export_ref(c, &tab);
	if (rc == ACM_OK)
		m->h_trans = tab;
	return rc;
}

int
iacsm_check_xd(iacsm_t *a, unsigned int *slots)
{
	struct acm_core *c = a ? (struct acm_core *)a->priv : NULL;
	return c ? acm_core_check_xd(c, slots) : ACM_ERR_ARG;
}

struct acm_automaton *
iacsm_device_automaton(iacsm_t *m)
{
	struct acm_core *c = core_of(m);
	return c ? c->dev : NULL;
}

void
iacsm_cleanup(iacsm_t *m)
{
	struct acm_core *c = core_of(m);

	if (!m)
		return;
	if (c)
		acm_core_cleanup(c);
	m->patterns = NULL;
	m->state_table = NULL;
}

void
iacsm_free(iacsm_t *m)
{
	struct acm_core *c = core_of(m);

	if (!m)
		return;
	if (c) {
		if (c->dev)
			acm_automaton_free(c->dev);
		acm_core_free(c);
	}
	free(m->h_trans);
	free(m);
}
