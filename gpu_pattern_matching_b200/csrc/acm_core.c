/*
 * acm_core.c -- Aho-Corasick automaton builder (host, plain C).
 *
 * Produces what the reference's acsm_compile()/acsm_gen_state_table() produce
 * (reference acsmx.c:552-671; ushort twin AC_ushorts/iacsmx.c:357-520) -- a dense
 * DFA plus per-state match information -- but built for the sm_100a kernels:
 *
 *   - array based, no per-node or per-queue-entry malloc (the reference mallocs
 *     every queue node, acsmx.c:166-186, and every match-list copy,
 *     acsmx.c:283-294);
 *   - states renumbered breadth-first so that shallow (hot) rows are contiguous
 *     and "is this edge a trie edge" is a compare against level_start[];
 *   - the match list of a state is not materialised: each state keeps the
 *     patterns that end exactly there (own list) and a link to the nearest
 *     proper suffix that has any (olink); the union along olink is exactly the
 *     reference's per-state match_list set (acsmx.c:417-429);
 *   - for byte automata, the entry filters of the scan kernels.
 *
 * The trie is still grown in the reference's order (last added pattern first,
 * one new state per byte, acsmx.c:318-349,579-580) so that node creation order
 * IS the reference's state numbering; acm_core_export_ref() uses that to emit a
 * table that can be compared bit for bit with the reference's h_trans.
 */
#define _GNU_SOURCE
#include <malloc.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acm_core.h"
#include "../../include/acm.h"

static __thread char acm_errbuf[512];

void
acm_set_error(const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(acm_errbuf, sizeof(acm_errbuf), fmt, ap);
	va_end(ap);
}

const char *
acm_last_error(void)
{
	return acm_errbuf;
}

struct acm_core *
acm_core_new(int alpha)
{
	struct acm_core *c = calloc(1, sizeof(*c));

	if (!c) {
		acm_set_error("acm_core_new: out of memory");
		return NULL;
	}
	c->alpha = alpha;
	c->sym_size = (alpha <= 256) ? 1 : 2;
	c->min_len = 0;
	return c;
}

int
acm_core_add(struct acm_core *c, const void *syms, int n, int nocase, int offset,
    int depth, void *id, int iid)
{
	struct acm_pat *p;

	if (c->compiled) {
		acm_set_error("add_pattern after compile");
		return c->status = ACM_ERR_STATE;
	}
	if (n < 0 || (n > 0 && !syms)) {
		acm_set_error("add_pattern: bad length %d", n);
		return c->status = ACM_ERR_ARG;
	}
	if (c->npats >= (int)ACM_KEY_PAT_MASK) {
		acm_set_error("add_pattern: more than %u patterns", ACM_KEY_PAT_MASK);
		return c->status = ACM_ERR_LIMIT;
	}
	if (c->npats == c->cap_pats) {
		int nc = c->cap_pats ? c->cap_pats * 2 : 1024;
		void *np = realloc(c->pats, (size_t)nc * sizeof(*c->pats));
		if (!np) {
			acm_set_error("add_pattern: out of memory");
			return c->status = ACM_ERR_NOMEM;
		}
		c->pats = np;
		c->cap_pats = nc;
	}
	p = &c->pats[c->npats];
	memset(p, 0, sizeof(*p));
	p->syms = malloc((size_t)(n ? n : 1) * c->sym_size);
	if (!p->syms) {
		acm_set_error("add_pattern: out of memory");
		return c->status = ACM_ERR_NOMEM;
	}
	if (n)
		memcpy(p->syms, syms, (size_t)n * c->sym_size);
	if (c->sym_size == 2) {
		const unsigned short *s = syms;
		for (int k = 0; k < n; k++)
			if (s[k] >= c->alpha) {
				free(p->syms);
				acm_set_error("add_pattern: symbol %u outside alphabet %d",
				    s[k], c->alpha);
				return c->status = ACM_ERR_ARG;
			}
	}
	p->n = n;
	p->nocase = nocase;
	p->offset = offset;
	p->depth = depth;
	p->id = id;
	p->iid = iid;
	c->npats++;
	if (n > c->max_len)
		c->max_len = n;
	if (n > 0 && (c->min_len == 0 || n < c->min_len))
		c->min_len = n;
	if (n == 0) {
		/* kept so later indices agree with the reference; can never match */
		acm_set_error("add_pattern: zero-length pattern %d ignored", c->npats - 1);
		return c->status = ACM_ERR_EMPTY_PATTERN;
	}
	return ACM_OK;
}

static inline unsigned
sym_at(const struct acm_core *c, const struct acm_pat *p, int k)
{
	return c->sym_size == 1 ? ((const unsigned char *)p->syms)[k]
	                        : ((const unsigned short *)p->syms)[k];
}

void
acm_tables_free(struct acm_tables *t)
{
	free(t->T); free(t->level_start); free(t->own_begin); free(t->own_pat);
	free(t->olink); free(t->fail); free(t->pat_len); free(t->pat_iid);
	free(t->bfs_to_ref); free(t->f1); free(t->f2); free(t->grams); free(t->b2s); free(t->b3);
	free(t->cand); free(t->pat_blob); free(t->pat_off); free(t->pat_win);
	free(t->cd_cls); free(t->cd_tab); free(t->cd_flat_begin); free(t->cd_flat_pat);
	free(t->rd_tab); free(t->rd_flat4); free(t->cd_flat4); free(t->xd_tab); free(t->xd_sid);
	memset(t, 0, sizeof(*t));
}

size_t
acm_tables_device_bytes(const struct acm_tables *t)
{
	size_t b = (size_t)t->num_states * t->alpha * 4;

	b += (size_t)(t->num_states + 1) * 4 + (size_t)t->own_total * 4;
	b += (size_t)t->num_states * 4 + (size_t)t->num_patterns * 4;
	b += (size_t)(t->max_depth + 2) * 4;
	if (t->f1)
		b += (1u << ACM_F1_BITS_LOG2) / 8 + ACM_F2_WORDS * 4 +
		    (size_t)t->gram_slots * sizeof(struct acm_gram_slot) +
		    ((size_t)t->cand_count + ACM_CAND_PAD) * sizeof(struct acm_cand) +
		    t->pat_blob_bytes + (size_t)t->num_patterns * 4;
	if (t->b3)
		b += ACM_B3_WORDS * 4;
	if (t->b2s)
		b += 65536 / 8;
	if (t->cd_tab)
		b += (size_t)t->num_states * t->cd_classes * 2 + 256 + (size_t)(t->num_states + 1) * 4 +
		    (size_t)t->cd_flat_total * 4 + (size_t)t->num_states * 16;
	if (t->rd_tab)
		b += (size_t)t->rd_len * 20;
	if (t->xd_tab)
		b += (size_t)t->xd_len * 8;
	return b;
}

/* ------------------------------------------------------------------------- */

struct trie {
	uint32_t  n;           /* nodes, creation order == reference state ids */
	uint32_t *parent;
	uint16_t *sym;
	uint32_t *first_child;
	uint32_t *next_sib;
	uint32_t *root_child;  /* [alpha], 0 = none                            */
	int32_t  *own_head;    /* per node: first own entry or -1               */
	int32_t  *own_next;    /* per pattern: next own entry of the same node  */
};

static void
trie_free(struct trie *t)
{
	free(t->parent); free(t->sym); free(t->first_child); free(t->next_sib);
	free(t->root_child); free(t->own_head); free(t->own_next);
}

static uint32_t
trie_child(const struct trie *t, uint32_t u, unsigned a)
{
	uint32_t c;

	if (u == 0)
		return t->root_child[a];
	for (c = t->first_child[u]; c; c = t->next_sib[c])
		if (t->sym[c] == a)
			return c;
	return 0;
}

static int
cmp_gtrip(const void *a, const void *b)
{
	const uint32_t *x = a, *y = b;
	if (x[0] != y[0])
		return x[0] < y[0] ? -1 : 1;
	return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

static int
cmp_u32(const void *a, const void *b)
{
	const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;

	return x < y ? -1 : (x > y);
}

/*
 * Row-displaced form of cd_tab: ONE 4-byte shared-memory lookup per transition in the common case.
 *
 * A state's row differs from the row of a state on its failure chain only in the columns where
 * some state in between has a trie child, so a deep state needs one or two "explicit" entries
 * instead of C.  Rows of the states up to depth d ("dense" states, ids 0 .. nd-1) are stored whole,
 * 33 entries (132 bytes) apart -- 33, not 32, so that column c of different dense rows falls into
 * different shared-memory banks: text is mostly "the same few letters after different prefixes",
 * and with a stride of 32 words every lane that reads 'e' out of a dense row would hit one bank;
 * a deeper state s keeps only the columns where its row differs from
 * the row of D(s), the first dense state on its failure chain.  All sparse rows are overlaid into
 * one array by first-fit row displacement (Tarjan & Yao): s gets a base rd_off[s], unique among
 * all states, such that its explicit column c lives at rd_tab[rd_off[s] + c] and nobody else's
 * entry does.  An entry carries the column it was stored for, so
 *
 *     e = rd_tab[off(s) + c];  if (column(e) != c)  e = rd_tab[33 * drow(s) + c];
 *
 * is the transition: the entry found at off(s) + c with column field c can only belong to the
 * state whose base is exactly off(s), i.e. to s.  The state IS the entry that led to it:
 *
 *     bits  0..4   column the entry is stored for (31 = empty slot)
 *     bits  5..7   |full match list of the target| (0 .. 4)
 *     bits  8..15  drow(target): the target's dense row (its own when the target is dense)
 *     bits 16..31  off(target)
 *
 * so no per-state record is read (the delta encoding this replaces read an 8-byte record and
 * then the entry: two dependent, bank-conflicting loads per byte).  rd_flat4[off] is the full
 * match list of the state with that base.  Needs the pattern bytes in one range (cd_range_lo),
 * C <= 31, at most 256 dense rows and 65 504 slots; picks the dense depth that makes the table
 * smallest; leaves rd_tab NULL when it does not fit the shared-memory budget.
 * fail[] / depth[] are in cd ids.
 */
struct rd_row {
	uint32_t state, n;
};

static int
cmp_rd_row(const void *a, const void *b)
{
	const struct rd_row *x = a, *y = b;

	if (x->n != y->n)
		return x->n > y->n ? -1 : 1;
	return x->state < y->state ? -1 : (x->state > y->state);
}

static int
build_cdfa_rd(struct acm_tables *t, const uint32_t *fail, const uint8_t *depth, int max_depth)
{
	const uint32_t C = t->cd_classes, n = t->num_states, CAP = 65536;
	uint32_t *dflt = NULL, *off = NULL, *tab = NULL;
	uint8_t *occ = NULL, *used_base = NULL;
	struct rd_row *rows = NULL;
	uint32_t s, k, nd = 0, len = 0, lowfree;
	int d, best_d = -1, rc = ACM_OK;
	size_t best = (size_t)-1;

	if (t->cd_range_lo < 0 || C > 31)
		return ACM_OK;
	dflt = malloc((size_t)n * 4);
	off = malloc((size_t)n * 4);
	rows = malloc((size_t)n * sizeof(*rows));
	occ = calloc(CAP + 64, 1);
	used_base = calloc(CAP + 64, 1);
	if (!dflt || !off || !rows || !occ || !used_base) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	/* cd ids keep breadth-first order except that a few deep states sit at the very end:
	 * "depth <= d" must be a prefix of the ids */
	for (d = 0; d <= max_depth && d < 255; d++) {
		uint32_t m = 0;
		size_t entries;
		int prefix = 1;
		while (m < n && depth[m] <= d)
			m++;
		for (s = m; s < n && prefix; s++)
			prefix = depth[s] > d;
		if (!prefix || m > ACM_RD_MAX_DENSE)
			continue;
		entries = (size_t)m * ACM_RD_ROW;
		for (s = m; s < n; s++) {
			const uint16_t *a, *b;
			dflt[s] = depth[fail[s]] <= d ? fail[s] : dflt[fail[s]];
			a = t->cd_tab + (size_t)s * C;
			b = t->cd_tab + (size_t)dflt[s] * C;
			for (k = 0; k < C; k++)
				entries += a[k] != b[k];
		}
		if (entries < best) {
			best = entries;
			best_d = d;
		}
	}
	if (best_d < 0 || best + 64 > CAP || best * 4 > ACM_RD_SMEM_BUDGET)
		goto out;
	d = best_d;
	while (nd < n && depth[nd] <= d)
		nd++;
	for (s = 0; s < n; s++) {
		const uint16_t *a = t->cd_tab + (size_t)s * C, *b;
		rows[s].state = s;
		rows[s].n = 0;
		if (s < nd) {
			dflt[s] = s;
			continue;
		}
		dflt[s] = depth[fail[s]] <= d ? fail[s] : dflt[fail[s]];
		b = t->cd_tab + (size_t)dflt[s] * C;
		for (k = 0; k < C; k++)
			rows[s].n += a[k] != b[k];
	}
	/* dense rows first, whole; their columns C .. 31 stay free for displaced entries */
	for (s = 0; s < nd; s++) {
		off[s] = s * ACM_RD_ROW;
		used_base[off[s]] = 1;
		memset(occ + off[s], 1, C);
	}
	len = nd * ACM_RD_ROW;
	/* first fit, rows with the most explicit entries first */
	qsort(rows + nd, n - nd, sizeof(*rows), cmp_rd_row);
	lowfree = 0;
	for (uint32_t r = nd; r < n; r++) {
		const uint32_t st = rows[r].state;
		const uint16_t *a = t->cd_tab + (size_t)st * C, *b = t->cd_tab + (size_t)dflt[st] * C;
		uint32_t cols[32], nc = 0, base, p;
		int placed = 0;

		for (k = 0; k < C; k++)
			if (a[k] != b[k])
				cols[nc++] = k;
		while (lowfree < CAP && occ[lowfree])
			lowfree++;
		if (nc == 0) {
			/* no entry of its own: any base nobody else uses (identity only) */
			for (base = 0; base + 32 < CAP && used_base[base]; base++)
				;
			if (base + 32 >= CAP)
				goto out;
			off[st] = base;
			used_base[base] = 1;
			if (base + 32 > len)
				len = base + 32;
			continue;
		}
		/* the first explicit column goes to a free slot p >= lowfree: base = p - cols[0] */
		for (p = lowfree > cols[0] ? lowfree : cols[0]; p + 32 < CAP; p++) {
			if (occ[p])
				continue;
			base = p - cols[0];
			if (used_base[base])
				continue;
			for (k = 1; k < nc && !occ[base + cols[k]]; k++)
				;
			if (k < nc)
				continue;
			placed = 1;
			break;
		}
		if (!placed)
			goto out;
		off[st] = base;
		used_base[base] = 1;
		for (k = 0; k < nc; k++)
			occ[base + cols[k]] = 1;
		if (base + 32 > len)
			len = base + 32;
	}
	if ((size_t)len * 4 > ACM_RD_SMEM_BUDGET)
		goto out;
	tab = malloc((size_t)len * 4 + 64);
	t->rd_flat4 = calloc((size_t)len * 4 + 16, 4);
	if (!tab || !t->rd_flat4) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	for (k = 0; k < len; k++)
		tab[k] = ACM_RD_EMPTY;
	for (s = 0; s < n; s++) {
		const uint16_t *a = t->cd_tab + (size_t)s * C, *b = t->cd_tab + (size_t)dflt[s] * C;
		for (k = 0; k < C; k++) {
			if (s < nd || a[k] != b[k]) {
				const uint32_t nx = a[k] & ACM_CD_STATE_MASK, recs = t->cd_flat4[(size_t)nx * 4] >> 24;
				tab[off[s] + k] = k | (recs << 5) | (dflt[nx] << 8) | (off[nx] << 16);
			}
		}
		memcpy(t->rd_flat4 + (size_t)off[s] * 4, t->cd_flat4 + (size_t)s * 4, 16);
	}
	t->rd_tab = tab;
	tab = NULL;
	t->rd_len = len;
	t->rd_dense_rows = nd;
	t->rd_dense_depth = d;
out:
	if (rc != ACM_OK || !t->rd_tab) {
		free(t->rd_flat4);
		t->rd_flat4 = NULL;
	}
	free(dflt); free(off); free(rows); free(occ); free(used_base); free(tab);
	return rc;
}

/*
 * The DFA as ONE small row-displaced array ("xd"): what k_scan_xd walks instead of the dense
 * T[states][alpha] (370 MiB for 10 000 ClamAV signatures, one 4-byte gather per input byte into
 * 2-5 x the L2: reference ahomatch.cl:56-65 does exactly that out of a table twice that size).
 *
 * A DFA row is the failure state's row with the state's own trie edges on top, so it differs from
 * the row of a shallow state on its failure chain only where the states in between have trie
 * children.  Which shallow state?  For bytes, the depth-1 state of the LAST INPUT BYTE: the
 * failure chain of any state ends in (the state of its last byte, or the root when no pattern
 * starts with it) -> root.  So with b the byte that led to state s,
 *
 *     T[s][c] = xd[off(s) + c]               if that slot carries symbol c (then it is s's own:
 *                                             bases are unique, see build_cdfa_rd), else
 *             = xd[off(d1(b)) + c]           d1(b) = T[root][b], if that slot carries c, else
 *             = xd[c]                        the root row, stored whole at base 0.
 *
 * The second and third lookups depend on the INPUT only, not on the walk: the kernel computes
 * them a step ahead, and only the first is on the state's dependency chain.  Explicit entries
 * (row differs from that default): 1.6 M for 10 000 ClamAV signatures = 6.8 MB with everything.
 * The same holds for ushort symbols (alphabet 2048) with "the previous symbol"; stored against
 * the root row alone, the depth-1 states of a few popular symbols would put ~40 inherited entries
 * into every deeper state's row (1.4 M entries for 2 000 packet-size signatures instead of ~0.1 M).
 * entry = symbol | any-match flag << sym_bits | base(target) << (sym_bits + 1); the root's row
 * comes first, then the states in breadth-first order (first fit, lowest free slot; no base is
 * congruent to alpha-1 modulo alpha, which is what lets free slots be marked unreadable), so a prefix
 * of the array -- what k_scan_xd keeps in shared memory -- holds the shallow states.
 * xd_sid[base] = breadth-first id of the state with that base (emission needs olink / own lists).
 */
/* the first free slot at or behind p; nxt[] is a union-find "next" array with path halving */
static uint32_t
xd_next_free(uint32_t *nxt, uint32_t p)
{
	while (nxt[p] != p) {
		nxt[p] = nxt[nxt[p]];
		p = nxt[p];
	}
	return p;
}

static int
build_xd(struct acm_tables *t)
{
	const uint32_t A = (uint32_t)t->alpha, n = t->num_states;
	const uint32_t sym_bits = A == 256 ? 8u : 11u;
	const uint32_t max_slots = 1u << (31 - sym_bits);
	const int levels = 1;
	uint32_t *off = NULL, *tab = NULL, *sid = NULL, *cols = NULL, *nxt = NULL;
	uint8_t *occ = NULL, *used_base = NULL, *depth1 = NULL;
	uint32_t s, k, len, lowfree, cap, hint[17], d1_end = 0;
	uint64_t est = 0;
	int rc = ACM_OK;

	if (!t->T || !t->fail || n == 0)
		return ACM_OK;
	for (k = 0; k < 17; k++)
		hint[k] = A;
	cols = malloc((size_t)A * 4);
	off = malloc((size_t)n * 4);
	depth1 = calloc(n, 1);                         /* 1: depth <= 1 */
	if (!cols || !off || !depth1) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	for (s = 0; s < t->level_start[2 <= t->max_depth + 1 ? 2 : t->max_depth + 1] && s < n; s++)
		depth1[s] = 1;
	/* how many slots at most: the root row plus every explicit entry (+ slack for first fit) */
	for (s = 1; s < n; s++) {
		const uint32_t *row = t->T + (size_t)s * A, *def;
		uint32_t d = 0;
		if (levels && !depth1[s]) {
			d = t->fail[s];
			while (!depth1[d])
				d = t->fail[d];
		}
		def = t->T + (size_t)d * A;
		for (k = 0; k < A; k++)
			est += (row[k] & ACM_T_MASK) != (def[k] & ACM_T_MASK);
	}
	cap = (uint32_t)(2 * est + 8 * A < max_slots ? 2 * est + 8 * A : max_slots);
	if (est + est / 2 + 2 * A >= max_slots)
		goto out;                                   /* does not fit the entry format: no xd table */
	occ = calloc((size_t)cap + A + 8, 1);
	used_base = calloc((size_t)cap + A + 8, 1);
	nxt = malloc(((size_t)cap + A + 8) * 4);        /* nxt[p]: a free slot >= p is at or behind nxt[p] (path-compressed) */
	if (!occ || !used_base || !nxt) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	for (k = 0; k < cap + A + 8; k++)
		nxt[k] = k;
	off[0] = 0;
	used_base[0] = 1;
	memset(occ, 1, A);
	for (k = 0; k < A; k++)
		nxt[k] = A;
	len = A;
	d1_end = A;
	lowfree = A;
	for (s = 1; s < n; s++) {
		const uint32_t *row = t->T + (size_t)s * A, *def;
		uint32_t d = 0, nc = 0, base = 0, p;
		int placed = 0;
		if (levels && !depth1[s]) {
			d = t->fail[s];
			while (!depth1[d])
				d = t->fail[d];
		}
		def = t->T + (size_t)d * A;
		for (k = 0; k < A; k++)
			if ((row[k] & ACM_T_MASK) != (def[k] & ACM_T_MASK))
				cols[nc++] = k;
		lowfree = xd_next_free(nxt, lowfree);
		if (nc == 0) {
			/* identity only: any base nobody else uses, near the front of the free area */
			for (base = lowfree > A ? lowfree - A : 0; base < cap && (used_base[base] || base % A == A - 1); base++)
				;
			if (base >= cap)
				goto out;
			placed = 1;
		} else {
			/* first fit over the free slots from where the last row with as many entries went (rows
			 * with one entry fit any hole; wide rows would test every hole of the crowded front each
			 * time), at most 4096 tries; then behind everything placed so far */
			uint32_t tries = 0;
			uint32_t *hp = &hint[nc < 16 ? nc : 16];
			if (*hp < lowfree)
				*hp = lowfree;
			for (p = xd_next_free(nxt, *hp > cols[0] ? *hp : cols[0]); p < cap && tries < 4096;
			     p = xd_next_free(nxt, p + 1), tries++) {
				base = p - cols[0];
				if (used_base[base] || base % A == A - 1)
					continue;
				for (k = 1; k < nc && !occ[base + cols[k]]; k++)
					;
				if (k < nc)
					continue;
				placed = 1;
				if (nc > 1)
					*hp = p;
				break;
			}
			if (!placed) {
				for (base = len; base + A < cap && (used_base[base] || base % A == A - 1); base++)
					;
				placed = base + A < cap;
				if (placed && nc > 1)
					*hp = base + cols[0];
			}
		}
		if (!placed)
			goto out;                               /* ran out of slots: leave xd unbuilt */
		off[s] = base;
		used_base[base] = 1;
		for (k = 0; k < nc; k++) {
			occ[base + cols[k]] = 1;
			nxt[base + cols[k]] = base + cols[k] + 1;
		}
		if (base + A > len)
			len = base + A;
		if (depth1[s])
			d1_end = len;                           /* every lookup from a state of depth <= 1 stays below this */
	}
	if (len > max_slots)
		goto out;
	tab = malloc((size_t)len * 4 + 64);
	sid = calloc((size_t)len + 16, 4);
	if (!tab || !sid) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	memset(tab, 0, (size_t)len * 4);
	for (s = 0; s < n; s++) {
		const uint32_t *row = t->T + (size_t)s * A, *def;
		uint32_t d = 0;
		if (s && levels && !depth1[s]) {
			d = t->fail[s];
			while (!depth1[d])
				d = t->fail[d];
		}
		def = t->T + (size_t)d * A;
		sid[off[s]] = s;
		for (k = 0; k < A; k++) {
			if (s == 0 || (row[k] & ACM_T_MASK) != (def[k] & ACM_T_MASK)) {
				const uint32_t nx = row[k] & ACM_T_MASK;
				tab[off[s] + k] = k | ((row[k] & ACM_T_ANY) ? 1u << sym_bits : 0u) | (off[nx] << (sym_bits + 1));
			}
		}
	}
	/* A free slot p must fail the check of EVERY state that can read it -- the state with base
	 * p - c reads it for symbol c -- so it claims a symbol whose implied base no state has: bases
	 * congruent to alpha-1 modulo alpha are never handed out (see the placement above), and
	 * c = (p + 1) mod alpha implies exactly such a base (or a negative one). */
	for (k = 0; k < len; k++)
		if (!occ[k])
			tab[k] = (k + 1) % A;                   /* next = root, no flag: never selected anyway */
	t->xd_tab = tab;
	t->xd_sid = sid;
	t->xd_len = len;
	t->xd_d1_end = d1_end;
	t->xd_levels = levels;
	t->xd_sym_bits = sym_bits;
	tab = sid = NULL;
out:
	free(cols); free(off); free(depth1); free(occ); free(used_base); free(nxt); free(tab); free(sid);
	return rc;
}

/*
 * Test support: every transition of the dense table T through the xd lookup rule (with the
 * previous symbol that leads to each state), over all states.  Violations, or -1 if not built.
 */
int
acm_core_check_xd(const struct acm_core *c, uint32_t *slots)
{
	const struct acm_tables *t = &c->tab;
	const uint32_t A = (uint32_t)t->alpha, n = t->num_states;
	uint32_t *off, *last, s, k;
	int bad = 0;

	if (!t->xd_tab || !t->T)
		return -1;
	if (slots)
		*slots = t->xd_len;
	off = malloc((size_t)n * 4);
	last = malloc((size_t)n * 4);
	if (!off || !last) {
		free(off); free(last);
		return ACM_ERR_NOMEM;
	}
	memset(off, 0xFF, (size_t)n * 4);
	for (k = 0; k < t->xd_len; k++)
		if (t->xd_sid[k] || k == 0)
			off[t->xd_sid[k]] = k;
	/* the symbol on the trie edge into s (breadth-first order: parents first) */
	last[0] = 0;
	for (s = 0; s < n; s++) {
		uint32_t d = 0;
		while (d + 1 <= (uint32_t)t->max_depth && s >= t->level_start[d + 1])
			d++;
		for (k = 0; k < A; k++) {
			const uint32_t nx = t->T[(size_t)s * A + k] & ACM_T_MASK;
			if (nx >= t->level_start[d + 1] && nx < t->level_start[d + 2 <= (uint32_t)t->max_depth + 1 ? d + 2 : (uint32_t)t->max_depth + 1])
				last[nx] = k;
		}
	}
	const uint32_t sb = t->xd_sym_bits, smask = (1u << sb) - 1u;
	for (s = 0; s < n; s++) {
		if (off[s] == 0xFFFFFFFFu) {
			bad++;
			continue;
		}
		for (k = 0; k < A; k++) {
			const uint32_t want = t->T[(size_t)s * A + k];
			uint32_t x = t->xd_tab[off[s] + k];
			if ((x & smask) != k) {
				x = t->xd_tab[k];                                       /* the root row */
				if (t->xd_levels && s) {
					const uint32_t r = t->xd_tab[last[s]];              /* root row entry of the previous symbol */
					const uint32_t y = t->xd_tab[(r >> (sb + 1)) + k];   /* the row of its depth-1 state */
					if ((y & smask) == k)
						x = y;
				}
			}
			if ((x & smask) != k || (x >> (sb + 1)) != off[want & ACM_T_MASK] ||
			    (((x >> sb) & 1u) != 0) != ((want & ACM_T_ANY) != 0))
				bad++;
		}
	}
	free(off); free(last);
	return bad;
}

/*
 * Test support: walks cd_tab and rd_tab in lockstep from the root over every (state, column) and
 * counts disagreements in target, match-count code or match list.  Every state is a trie node,
 * hence reachable, so 0 means the row-displaced table IS the automaton.  -1: no rd table.
 * *slots / *dense report the table shape.
 */
int
acm_core_check_rd(const struct acm_core *c, uint32_t *slots, uint32_t *dense)
{
	const struct acm_tables *t = &c->tab;
	const uint32_t C = t->cd_classes, n = t->num_states;
	uint32_t *ent, *queue, head = 0, tail = 0, k;
	int bad = 0;

	if (!t->rd_tab || !t->cd_tab)
		return -1;
	if (slots)
		*slots = t->rd_len;
	if (dense)
		*dense = t->rd_dense_rows;
	ent = malloc((size_t)n * 4);           /* cd id -> the entry (state word) it was reached by */
	queue = malloc((size_t)n * 4);
	if (!ent || !queue) {
		free(ent); free(queue);
		return ACM_ERR_NOMEM;
	}
	memset(ent, 0xFF, (size_t)n * 4);
	ent[0] = 0;                            /* root: base 0, dense row 0 */
	queue[tail++] = 0;
	while (head < tail) {
		const uint32_t s = queue[head++], e = ent[s];
		const uint32_t off = e >> 16, drow = (e >> 8) & 0xFF;
		for (k = 0; k < C; k++) {
			uint32_t x, nx, want = t->cd_tab[(size_t)s * C + k];
			if (off + k >= t->rd_len) {
				bad++;
				continue;
			}
			x = t->rd_tab[off + k];
			if ((x & 31) != k)
				x = t->rd_tab[drow * ACM_RD_ROW + k];
			if ((x & 31) != k || ((x >> 5) & 7) != (t->cd_flat4[(size_t)(want & ACM_CD_STATE_MASK) * 4] >> 24)) {
				bad++;
				continue;
			}
			nx = want & ACM_CD_STATE_MASK;
			if (ent[nx] == 0xFFFFFFFFu) {
				ent[nx] = x;
				queue[tail++] = nx;
			} else if ((ent[nx] >> 8) != (x >> 8)) {
				bad++;                     /* same target must mean same base and dense row */
			}
			if (memcmp(t->rd_flat4 + (size_t)(x >> 16) * 4, t->cd_flat4 + (size_t)nx * 4, 16) != 0)
				bad++;
		}
	}
	if (tail != n)
		bad += (int)(n - tail);
	/* bases are unique */
	{
		uint8_t *seen = calloc(65536, 1);
		uint32_t s;
		if (seen) {
			for (s = 0; s < n; s++) {
				if (ent[s] == 0xFFFFFFFFu)
					continue;
				if (seen[ent[s] >> 16])
					bad++;
				seen[ent[s] >> 16] = 1;
			}
			free(seen);
		}
	}
	free(ent); free(queue);
	return bad;
}

/*
 * Class-compressed DFA for small automata over few distinct bytes (word lists over text):
 * bytes that occur in no pattern all behave alike (every state goes to the root on them), so
 * the 256 columns of T collapse to C = (distinct pattern bytes) + 1, and with <= 2^14 states
 * an entry fits 16 bits together with a 2-bit "how many patterns end here" code (0, 1, 2,
 * 3 = three or four).  States keep their breadth-first ids (a prefix of the table = the shallow,
 * most visited states) except that the states where FOUR patterns end are moved to the very
 * end, so that "entry >= cd_thr4" tells three from four without a lookup; automata with a
 * state where more than four patterns end do not get this form.  Each state's FULL match list
 * (own patterns plus those inherited along the output links -- reference acsmx.c:417-429
 * copies them) is flattened and sorted by pattern index, so a sequential walk emits records
 * already in canonical order.
 */
static int
build_cdfa(struct acm_core *c)
{
	struct acm_tables *t = &c->tab;
	const uint32_t n = t->num_states;
	uint8_t used[256];
	uint8_t col_byte[ACM_CD_MAX_CLASSES];
	uint32_t s, k, C = 0, n4 = 0, next_lo, next_hi;
	int lo = -1, hi = -1, b, other = -1, d, rc = ACM_OK;
	uint64_t total = 0;
	uint32_t *cnt = NULL, *perm = NULL, *fail = NULL;
	uint8_t *depth = NULL;

	if (n > ACM_CD_MAX_STATES || n == 0)
		return ACM_OK;
	memset(used, 0, sizeof(used));
	for (k = 0; k < (uint32_t)c->npats; k++)
		for (b = 0; b < c->pats[k].n; b++)
			used[((const unsigned char *)c->pats[k].syms)[b]] = 1;
	for (b = 0; b < 256; b++) {
		if (used[b]) {
			if (lo < 0)
				lo = b;
			hi = b;
			C++;
		} else if (other < 0) {
			other = b;
		}
	}
	if (C == 0 || C + 1 > ACM_CD_MAX_CLASSES || other < 0)
		return ACM_OK;

	/* full-list lengths (olink[s] < s in BFS order); more than four anywhere: no cdfa */
	cnt = calloc(n + 1, 4);
	perm = malloc((size_t)n * 4);
	fail = malloc((size_t)n * 4);
	depth = malloc(n);
	if (!cnt || !perm || !fail || !depth) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	for (s = 0; s < n; s++) {
		cnt[s] = t->own_begin[s + 1] - t->own_begin[s] + (s ? cnt[t->olink[s]] : 0);
		total += cnt[s];
		if (cnt[s] > 4)
			goto out;
		n4 += cnt[s] == 4;
	}
	/* cd id: BFS order, four-pattern states last */
	next_lo = 0;
	next_hi = n - n4;
	for (s = 0; s < n; s++)
		perm[s] = cnt[s] == 4 ? next_hi++ : next_lo++;
	for (d = 0; d <= t->max_depth; d++)
		for (s = t->level_start[d]; s < t->level_start[d + 1]; s++)
			depth[perm[s]] = (uint8_t)(d > 255 ? 255 : d);
	for (s = 0; s < n; s++)
		fail[perm[s]] = perm[t->fail[s]];

	t->cd_cls = malloc(256);
	t->cd_flat_begin = malloc((size_t)(n + 1) * 4);
	t->cd_flat_pat = malloc((size_t)(total + 1) * 4);
	if (!t->cd_cls || !t->cd_flat_begin || !t->cd_flat_pat) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	if (hi - lo + 2 <= ACM_CD_MAX_CLASSES) {
		/* contiguous range: column = byte - lo (unused bytes inside the range keep a real column) */
		t->cd_range_lo = lo;
		C = (uint32_t)(hi - lo + 2);
		for (b = 0; b < 256; b++)
			t->cd_cls[b] = (uint8_t)((b >= lo && b <= hi) ? b - lo : (int)C - 1);
		for (k = 0; k + 1 < C; k++)
			col_byte[k] = (uint8_t)(lo + (int)k);
	} else {
		t->cd_range_lo = -1;
		C = C + 1;
		k = 0;
		for (b = 0; b < 256; b++) {
			if (used[b]) {
				col_byte[k] = (uint8_t)b;
				t->cd_cls[b] = (uint8_t)k++;
			} else {
				t->cd_cls[b] = (uint8_t)(C - 1);
			}
		}
	}
	col_byte[C - 1] = (uint8_t)other;
	t->cd_tab = malloc((size_t)n * C * 2);
	if (!t->cd_tab) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}

	/* flat lists in cd-id order: inv[] walks the cd ids */
	{
		uint32_t *inv = malloc((size_t)n * 4);
		uint32_t w = 0, id;
		if (!inv) {
			rc = ACM_ERR_NOMEM;
			goto out;
		}
		for (s = 0; s < n; s++)
			inv[perm[s]] = s;
		for (id = 0; id < n; id++) {
			uint32_t v, w0 = w;
			t->cd_flat_begin[id] = w;
			for (v = inv[id]; v; v = t->olink[v])
				for (k = t->own_begin[v]; k < t->own_begin[v + 1]; k++)
					t->cd_flat_pat[w++] = t->own_pat[k];
			if (w - w0 > 1)
				qsort(t->cd_flat_pat + w0, w - w0, 4, cmp_u32);
		}
		t->cd_flat_begin[n] = w;
		t->cd_flat_total = w;
		free(inv);
		/* the same lists inline, one 16-byte load per hit in the expansion kernel */
		t->cd_flat4 = calloc((size_t)n * 4, 4);
		if (!t->cd_flat4) {
			rc = ACM_ERR_NOMEM;
			goto out;
		}
		for (id = 0; id < n; id++) {
			const uint32_t fb = t->cd_flat_begin[id], m = t->cd_flat_begin[id + 1] - fb;
			for (k = 0; k < m; k++)
				t->cd_flat4[(size_t)id * 4 + k] = t->cd_flat_pat[fb + k];
			t->cd_flat4[(size_t)id * 4] |= m << 24;
		}
	}
	for (s = 0; s < n; s++) {
		for (k = 0; k < C; k++) {
			const uint32_t nx = t->T[(size_t)s * 256 + col_byte[k]] & ACM_T_MASK;
			const uint32_t code = cnt[nx] < 3 ? cnt[nx] : 3;
			t->cd_tab[(size_t)perm[s] * C + k] = (uint16_t)(perm[nx] | (code << ACM_CD_STATE_BITS));
		}
	}
	t->cd_classes = C;
	t->cd_thr4 = n4 ? ((3u << ACM_CD_STATE_BITS) | (n - n4)) : 0xFFFFFFFFu;
	rc = build_cdfa_rd(t, fail, depth, t->max_depth);
out:
	free(cnt); free(perm); free(fail); free(depth);
	if (rc == ACM_OK && !t->cd_classes) {
		/* not eligible: leave nothing half built */
		free(t->cd_cls); free(t->cd_tab); free(t->cd_flat_begin); free(t->cd_flat_pat); free(t->cd_flat4);
		t->cd_flat4 = NULL;
		t->cd_cls = NULL;
		t->cd_tab = NULL;
		t->cd_flat_begin = NULL;
		t->cd_flat_pat = NULL;
	}
	return rc;
}

/* first-three-bytes Bloom bitmap (k_scan_start2<false>): three bits per key in one 32-bit word;
 * a pattern of one or two bytes enters every completion */
static uint32_t *
build_b3(const struct acm_core *c)
{
	uint32_t *b3 = calloc(ACM_B3_WORDS, 4);

	if (!b3)
		return NULL;
	for (int k = 0; k < c->npats; k++) {
		const unsigned char *p = c->pats[k].syms;
		const uint32_t n = (uint32_t)c->pats[k].n;
		if (n == 0)
			continue;
		const uint32_t fixed = n < 3 ? n : 3;
		uint32_t g0 = 0;
		for (uint32_t b = 0; b < fixed; b++)
			g0 |= (uint32_t)p[b] << (8 * b);
		for (uint32_t rest = 0; rest < (1u << (8 * (3 - fixed))); rest++) {
			const uint32_t g = g0 | (rest << (8 * fixed));
			const uint32_t h = g * ACM_HASH2_MUL;
			b3[h >> 18] |= (1u << (h & 31)) | (1u << ((h >> 5) & 31)) | (1u << ((h >> 10) & 31));
		}
	}
	return b3;
}

/*
 * How common a 4-byte window is in the data a scanner meets, whatever the signature set says: a
 * window the filter indexes is a window whose every occurrence in the input costs an exact-table
 * probe and a candidate compare.  Added to the gram's popularity among the signatures when the
 * indexed window of a (pattern, alignment) is chosen, so that a signature is found through a
 * window of machine code or binary structure rather than through its run of zeros, its UTF-16
 * string or its English words when it has the choice (10 000 ClamAV signatures: true gram hits on
 * English text 9.5 % -> 0.8 % of the aligned windows, on a corpus of shared libraries 0.87 % ->
 * 0.18 %; tools/density_sweep.py).
 */
static uint32_t
gram_background(uint32_t g)
{
	const unsigned b[4] = {g & 255u, (g >> 8) & 255u, (g >> 16) & 255u, g >> 24};
	unsigned zeros = 0, texty = 0, k;

	if (b[0] == b[1] && b[1] == b[2] && b[2] == b[3])
		return 1000;                      /* runs: zero pages, erased flash, blanks, NOP sleds */
	for (k = 0; k < 4; k++) {
		zeros += b[k] == 0;
		texty += (b[k] >= 'a' && b[k] <= 'z') || (b[k] >= 'A' && b[k] <= 'Z') || (b[k] >= '0' && b[k] <= '9') ||
		    b[k] == ' ' || b[k] == '.' || b[k] == ',' || b[k] == '\n' || b[k] == '\r' || b[k] == '-' || b[k] == '_' ||
		    b[k] == '/';
	}
	if (zeros >= 2)
		return 200;                       /* small little-endian integers, UTF-16 text, padding */
	if (texty == 4)
		return 100;                       /* words */
	return 0;
}

static int
build_filters(struct acm_core *c)
{
	struct acm_tables *t = &c->tab;
	uint32_t s, k;

	t->b3 = build_b3(c);
	if (!t->b3)
		return ACM_ERR_NOMEM;

	/*
	 * Mixed sets: a few patterns shorter than the sampled filter can index (7 bytes at stride
	 * 4, 10 at stride 8) among many long ones.  Instead of giving the whole set to the 2-byte
	 * start filter (15 x slower on signature sets), the patterns shorter than split_len are left
	 * out of the sampled filter and found by a second pass: the start filter over b2s, the start
	 * bitmap of the short patterns alone, walking the trie to depth split_len - 1 at most (a
	 * node at depth d ends patterns of length d only, so that pass reports short patterns
	 * and nothing else).  Taken when at most one pattern in eight is short; ACM_HYBRID=0 turns
	 * it off, ACM_HYBRID=1 takes it whenever both kinds exist (tests).
	 */
	t->split_len = 0;
	if (t->min_pattern_len < 7) {
		const char *hy = getenv("ACM_HYBRID"), *force = getenv("ACM_SAMPLE_STRIDE");
		const int forced = hy && atoi(hy) == 1;
		uint32_t total = 0, n7 = 0, n10 = 0;

		if (hy && atoi(hy) == 0)
			return ACM_OK;
		for (k = 0; k < (uint32_t)c->npats; k++) {
			const uint32_t n = (uint32_t)c->pats[k].n;
			total += n > 0;
			n7 += n > 0 && n < 7;
			n10 += n > 0 && n < 10;
		}
		if (!(force && atoi(force) == 4) && n10 < total && (forced || n10 * 8 <= total))
			t->split_len = 10;
		else if (n7 < total && (forced || n7 * 8 <= total))
			t->split_len = 7;
		else
			return ACM_OK;
		t->b2s = calloc(65536 / 32, 4);
		if (!t->b2s)
			return ACM_ERR_NOMEM;
		for (k = 0; k < (uint32_t)c->npats; k++) {
			const unsigned char *p = c->pats[k].syms;
			const uint32_t n = (uint32_t)c->pats[k].n;
			if (n == 0 || n >= (uint32_t)t->split_len)
				continue;
			for (unsigned b1 = (n == 1 ? 0 : p[1]); b1 < (n == 1 ? 256u : (unsigned)p[1] + 1u); b1++) {
				const uint32_t idx = p[0] | (b1 << 8);
				t->b2s[idx >> 5] |= 0x80000000u >> (idx & 31);
			}
		}
	}
#define IS_SHORT(n_) ((n_) == 0 || (n_) < (uint32_t)t->split_len)

	/*
	 * Sampled entry filter.  stride 4: every occurrence of a pattern >= 7 bytes contains a
	 * 4-byte window at a multiple of 4; stride 8: every occurrence of a pattern >= 10 bytes
	 * contains a 3-byte window at a multiple of 8 (half the bitmap lookups per input byte).
	 * f1 / f2 hash the window bytes (G = 4 or 3); the exact table is keyed by the 4 bytes at
	 * the window position in both cases -- when the pattern ends after the third byte (length
	 * 10, offset 7) the fourth is a wildcard and all 256 keys are entered.
	 */
	t->f1 = calloc((1u << ACM_F1_BITS_LOG2) / 32, 4);
	t->f2 = calloc(ACM_F2_WORDS, 4);
	if (!t->f1 || !t->f2)
		return ACM_ERR_NOMEM;
	{
		struct gtrip { uint32_t gram, cand; } *tr;
		const char *force = getenv("ACM_SAMPLE_STRIDE");
		const uint32_t long_min = t->split_len ? (uint32_t)t->split_len : (uint32_t)t->min_pattern_len;
		const uint32_t S = (long_min >= 10 && !(force && atoi(force) == 4)) ? 8 : 4;
		uint64_t want, ntr_max = 0;
		uint32_t slots = 1024, lg = 10, ntr = 0, blob = 0;

		t->sample_stride = (int)S;
		for (k = 0; k < (uint32_t)c->npats; k++) {
			if (IS_SHORT((uint32_t)c->pats[k].n))
				continue;
			for (uint32_t j = 0; j < S; j++)
				ntr_max += (j + 4 <= (uint32_t)c->pats[k].n) ? 1 : 256;
		}
		want = ntr_max * 2;
		while (slots < want) {
			slots <<= 1;
			lg++;
		}
		t->gram_slots = slots;
		t->grams = calloc(slots, sizeof(*t->grams));
		tr = malloc((ntr_max + 1) * sizeof(*tr));
		t->pat_off = calloc((size_t)c->npats + 1, 4);
		t->pat_win = calloc((size_t)c->npats * 8 + 8, 1);
		if (!t->grams || !tr || !t->pat_off || !t->pat_win) {
			free(tr);
			return ACM_ERR_NOMEM;
		}
		for (k = 0; k < (uint32_t)c->npats; k++) {
			t->pat_off[k] = blob;
			blob += ((uint32_t)c->pats[k].n + 3u + 4u) & ~3u;   /* >= 4 zero bytes after each */
		}
		t->pat_blob_bytes = blob + 16;
		t->pat_blob = calloc(t->pat_blob_bytes, 1);
		if (!t->pat_blob) {
			free(tr);
			return ACM_ERR_NOMEM;
		}
		/*
		 * Which window of a pattern serves alignment j?  Any offset o = j + m * S with the
		 * window inside the pattern will do: an occurrence starting at s with (-s) mod S = j has
		 * an aligned window at s + o, and the kernel tests every aligned window.  The first one
		 * (o = j) is taken unless its 4-byte key is popular -- virus prologues such as e8 00 00
		 * 5d 81 ed head hundreds of signatures, and a text that contains one made the kernel walk
		 * candidate lists for ~35 us in one chunk (per-chunk trace: 44 list rounds), which is
		 * what spread the CTA exit times -- in which case the rarest key among the later windows of
		 * that alignment is used.  Popularity = how many (pattern, alignment) entries a key would
		 * get with first-window indexing.
		 */
		uint32_t *pop_key = NULL, *pop_cnt = NULL, pop_slots = 1024;
		{
			uint64_t first_entries = 0;
			for (k = 0; k < (uint32_t)c->npats; k++)
				if (!IS_SHORT((uint32_t)c->pats[k].n))
					first_entries += S;
			while (pop_slots < first_entries * 2)
				pop_slots <<= 1;
			pop_key = calloc(pop_slots, 4);
			pop_cnt = calloc(pop_slots, 4);
			if (!pop_key || !pop_cnt) {
				free(tr); free(pop_key); free(pop_cnt);
				return ACM_ERR_NOMEM;
			}
			for (k = 0; k < (uint32_t)c->npats; k++) {
				const unsigned char *p = c->pats[k].syms;
				const uint32_t n = (uint32_t)c->pats[k].n;
				for (uint32_t j = 0; !IS_SHORT(n) && j < S && j + 4 <= n; j++) {
					const uint32_t g = (uint32_t)p[j] | ((uint32_t)p[j + 1] << 8) | ((uint32_t)p[j + 2] << 16) |
					    ((uint32_t)p[j + 3] << 24);
					for (s = (g * ACM_HASH3_MUL) & (pop_slots - 1);; s = (s + 1) & (pop_slots - 1)) {
						if (pop_cnt[s] == 0 || pop_key[s] == g) {
							pop_key[s] = g;
							pop_cnt[s]++;
							break;
						}
					}
				}
			}
		}
#define POP_OF(g, out)                                                                  \
	do {                                                                            \
		uint32_t s_ = ((g) * ACM_HASH3_MUL) & (pop_slots - 1);                      \
		(out) = 0;                                                                  \
		while (pop_cnt[s_]) {                                                       \
			if (pop_key[s_] == (g)) {                                               \
				(out) = pop_cnt[s_];                                                \
				break;                                                              \
			}                                                                       \
			s_ = (s_ + 1) & (pop_slots - 1);                                        \
		}                                                                           \
	} while (0)
		for (k = 0; k < (uint32_t)c->npats; k++) {
			const unsigned char *p = c->pats[k].syms;
			const uint32_t n = (uint32_t)c->pats[k].n;
			if (n == 0)
				continue;
			memcpy(t->pat_blob + t->pat_off[k], p, n);
			if (IS_SHORT(n))
				continue;
			for (uint32_t j = 0; j < S; j++) {
				uint32_t o = j, g = 0;
				if (j + 4 <= n) {
					uint32_t best;
					const uint32_t g0 = (uint32_t)p[j] | ((uint32_t)p[j + 1] << 8) | ((uint32_t)p[j + 2] << 16) |
					    ((uint32_t)p[j + 3] << 24);
					uint32_t pop0;
					POP_OF(g0, pop0);
					best = pop0 + gram_background(g0);
					if (best > ACM_CAND_POPULAR) {
						for (uint32_t o2 = j + S; o2 + 4 <= n && o2 <= ACM_CAND_O_MAX; o2 += S) {
							uint32_t pc;
							const uint32_t g2 = (uint32_t)p[o2] | ((uint32_t)p[o2 + 1] << 8) |
							    ((uint32_t)p[o2 + 2] << 16) | ((uint32_t)p[o2 + 3] << 24);
							POP_OF(g2, pc);
							/* never into a longer candidate list than the one it leaves (or a short one) */
							if (pc > pop0 && pc > ACM_CAND_POPULAR)
								continue;
							pc += gram_background(g2);
							if (pc < best) {
								best = pc;
								o = o2;
							}
						}
					}
				}
				/*
				 * Both bitmaps hash the 4 bytes at the window.  A pattern that ends after the third
				 * (length 10 at alignment 7 with stride 8: a dozen of the ClamAV signatures) leaves
				 * the fourth byte free: all 256 completions are entered, as in the exact table.
				 * Hashing only 3 bytes for everybody made 0.45 % of all random windows TRUE gram hits
				 * (75 k grams of 2^24) that only the exact table could reject.
				 */
				t->pat_win[(size_t)k * 8 + j] = (uint8_t)o;
				if (o > t->max_win)
					t->max_win = o;
				const uint32_t fixed = o + 4 <= n ? 4 : 3;
				for (uint32_t b = 0; b < fixed; b++)
					g |= (uint32_t)p[o + b] << (8 * b);
				for (uint32_t b3 = 0; b3 < (fixed == 4 ? 1u : 256u); b3++) {
					const uint32_t gg = fixed == 4 ? g : (g | (b3 << 24));
					const uint32_t h1 = gg * ACM_HASH1_MUL;
					const uint32_t h2 = gg * ACM_HASH2_MUL;
					/* both levels are blocked Bloom filters: all bits of a gram live in the one 32-bit
					 * word the kernel fetches.  Level 1: k = 2 (bit indices from hash bits 0..4 and
					 * 12..16); level 2, tested only for level-1 survivors: k = 3 (+ bits 6..10) */
					t->f1[h1 >> (32 - (ACM_F1_BITS_LOG2 - 5))] |=
					    (0x80000000u >> (h1 & 31)) | (0x80000000u >> ((h1 >> 12) & 31));
					t->f2[(uint32_t)(((uint64_t)h2 * ACM_F2_WORDS) >> 32)] |= (0x80000000u >> (h2 & 31)) |
					    (0x80000000u >> ((h2 >> 12) & 31)) | (0x80000000u >> ((h2 >> 6) & 31));
				}
				if (o + 4 <= n) {
					tr[ntr].gram = (uint32_t)p[o] | ((uint32_t)p[o + 1] << 8) |
					    ((uint32_t)p[o + 2] << 16) | ((uint32_t)p[o + 3] << 24);
					tr[ntr].cand = k | (o << ACM_CAND_O_SHIFT);
					ntr++;
				} else {
					for (uint32_t b3 = 0; b3 < 256; b3++) {
						tr[ntr].gram = g | (b3 << 24);
						tr[ntr].cand = k | (o << ACM_CAND_O_SHIFT);
						ntr++;
					}
				}
			}
		}
#undef POP_OF
		free(pop_key);
		free(pop_cnt);
		/* group by gram (stable order inside a gram is irrelevant: results get sorted) */
		qsort(tr, ntr, sizeof(*tr), cmp_gtrip);
		t->cand = calloc((size_t)ntr + ACM_CAND_PAD, sizeof(*t->cand));
		if (!t->cand) {
			free(tr);
			return ACM_ERR_NOMEM;
		}
		t->cand_count = ntr;
		for (k = 0; k < ntr; k++) {
			const int first = (k == 0) || tr[k - 1].gram != tr[k].gram;
			const int lastc = (k + 1 == ntr) || tr[k + 1].gram != tr[k].gram;
			const uint32_t pid = tr[k].cand & ACM_CAND_ID_MASK;
			const uint32_t o = tr[k].cand >> ACM_CAND_O_SHIFT;
			const unsigned char *pb = t->pat_blob + t->pat_off[pid] + o;   /* zero padded: >= 4 zero bytes + next pattern */
			const uint32_t rem = (uint32_t)c->pats[pid].n - o;           /* pattern bytes from o on, >= 3 */
			uint32_t w[2] = {0, 0};
			for (uint32_t b = 0; b < 8 && b < rem; b++)
				w[b >> 2] |= (uint32_t)pb[b] << (8 * (b & 3));
			t->cand[k].info = tr[k].cand;
			t->cand[k].at0 = w[0];
			t->cand[k].at1 = w[1];
			t->cand[k].len = (uint32_t)c->pats[pid].n | (lastc ? ACM_CAND_LAST : 0);
			t->cand[k].pat_off = t->pat_off[pid];
			{
				const unsigned char *pe = t->pat_blob + t->pat_off[pid] + c->pats[pid].n - 4;   /* n >= 7 here */
				t->cand[k].tail = (uint32_t)pe[0] | ((uint32_t)pe[1] << 8) | ((uint32_t)pe[2] << 16) |
				    ((uint32_t)pe[3] << 24);
			}
			if (first) {
				uint32_t g = tr[k].gram;
				for (s = (g * ACM_HASH3_MUL) >> (32 - lg);; s = (s + 1) & (slots - 1))
					if (t->grams[s].begin1 == 0) {
						t->grams[s].gram = g;
						uint32_t cnt = 1;
						while (k + cnt < ntr && tr[k + cnt].gram == g)
							cnt++;
						t->grams[s].begin1 = k + 1;
						t->grams[s].count = cnt;
						t->gram_count++;
						break;
					}
			}
		}
		free(tr);
	}
#undef IS_SHORT
	return ACM_OK;
}

/*
 * Consistency of the sampled-filter tables with the patterns (host side, no device): for every
 * pattern and every alignment j < stride exactly one window o = j (mod stride) is indexed; its
 * 4-byte key (all 256 completions when the pattern ends after the third byte) has its bits in
 * both bitmaps and a slot in the exact table whose candidate list -- `count` records, the last
 * one flagged -- holds (pattern, o) with the pattern's bytes at o, its length, its last four
 * bytes and its offset in pat_blob.  Returns the number of violations, -1 if no filter was built.
 */
int
acm_core_check_filters(const struct acm_core *c)
{
	const struct acm_tables *t = &c->tab;
	const uint32_t S = (uint32_t)t->sample_stride;
	int bad = 0;

	if (!c->compiled || !t->f1 || !t->grams || S == 0)
		return -1;
	for (int k = 0; k < c->npats; k++) {
		const unsigned char *p = c->pats[k].syms;
		const uint32_t n = (uint32_t)c->pats[k].n;
		if (n == 0)
			continue;
		if (memcmp(t->pat_blob + t->pat_off[k], p, n) != 0)
			bad++;
		if (n < (uint32_t)t->split_len) {
			/* a short pattern: not in the sampled filter, its start is in b2s */
			for (unsigned b1 = 0; b1 < 256; b1++) {
				const uint32_t idx = p[0] | (b1 << 8);
				if ((n == 1 || b1 == p[1]) && !(t->b2s[idx >> 5] & (0x80000000u >> (idx & 31))))
					bad++;
			}
			continue;
		}
		for (uint32_t j = 0; j < S; j++) {
			uint32_t found = 0;
			for (uint32_t o = j; o + 3 <= n; o += S) {
				const uint32_t fixed = o + 4 <= n ? 4 : 3;
				uint32_t g = 0, hit_all = 1;
				for (uint32_t b = 0; b < fixed; b++)
					g |= (uint32_t)p[o + b] << (8 * b);
				for (uint32_t b3 = 0; b3 < (fixed == 4 ? 1u : 256u) && hit_all; b3++) {
					const uint32_t gg = fixed == 4 ? g : (g | (b3 << 24));
					const uint32_t h1 = gg * ACM_HASH1_MUL, h2 = gg * ACM_HASH2_MUL;
					const uint32_t w1 = t->f1[h1 >> (32 - (ACM_F1_BITS_LOG2 - 5))];
					const uint32_t w2 = t->f2[(uint32_t)(((uint64_t)h2 * ACM_F2_WORDS) >> 32)];
					const uint32_t m1 = (0x80000000u >> (h1 & 31)) | (0x80000000u >> ((h1 >> 12) & 31));
					const uint32_t m2 = (0x80000000u >> (h2 & 31)) | (0x80000000u >> ((h2 >> 12) & 31)) |
					    (0x80000000u >> ((h2 >> 6) & 31));
					uint32_t s, in_list = 0;
					if ((w1 & m1) != m1 || (w2 & m2) != m2) {
						hit_all = 0;
						break;
					}
					for (s = (gg * ACM_HASH3_MUL) >> (32 - __builtin_ctz(t->gram_slots));; s = (s + 1) & (t->gram_slots - 1)) {
						if (t->grams[s].begin1 == 0 || t->grams[s].gram == gg)
							break;
					}
					if (t->grams[s].begin1 == 0) {
						hit_all = 0;
						break;
					}
					for (uint32_t i = 0; i < t->grams[s].count; i++) {
						const struct acm_cand *cd = &t->cand[t->grams[s].begin1 - 1 + i];
						const int last = (cd->len & ACM_CAND_LAST) != 0;
						if (last != (i + 1 == t->grams[s].count))
							bad++;
						if ((cd->info & ACM_CAND_ID_MASK) == (uint32_t)k && (cd->info >> ACM_CAND_O_SHIFT) == o) {
							uint32_t w[2] = {0, 0}, tail;
							for (uint32_t b = 0; b < 8 && o + b < n; b++)
								w[b >> 2] |= (uint32_t)p[o + b] << (8 * (b & 3));
							tail = (uint32_t)p[n - 4] | ((uint32_t)p[n - 3] << 8) | ((uint32_t)p[n - 2] << 16) |
							    ((uint32_t)p[n - 1] << 24);
							/* with a free fourth byte the record holds this completion's byte */
							if (fixed == 3)
								w[0] |= b3 << 24;
							if ((cd->len & ~ACM_CAND_LAST) != n || cd->pat_off != t->pat_off[k] || cd->tail != tail ||
							    cd->at1 != w[1] || (fixed == 4 && cd->at0 != w[0]) ||
							    (fixed == 3 && (cd->at0 & 0x00ffffffu) != (w[0] & 0x00ffffffu)))
								bad++;
							in_list++;
						}
					}
					if (in_list != 1)
						hit_all = 0;
				}
				found += hit_all;
			}
			/* alignments past the end of a short pattern cannot occur (n >= stride + 2) */
			if (found != 1)
				bad++;
		}
	}
	return bad;
}

int
acm_core_compile(struct acm_core *c)
{
	struct acm_tables *t = &c->tab;
	struct trie tr;
	const int A = c->alpha;
	uint32_t total = 1, last = 0, i, u, qt;
	uint32_t *order = NULL, *bfs_id = NULL, *depth = NULL;
	uint32_t *cbuf = NULL;
	int k, rc = ACM_OK;

	if (c->compiled)
		return ACM_OK;
	memset(&tr, 0, sizeof(tr));
	for (k = 0; k < c->npats; k++)
		total += (uint32_t)c->pats[k].n;
	if (total >= ACM_T_MASK) {
		acm_set_error("compile: too many states");
		return c->status = ACM_ERR_LIMIT;
	}

	tr.parent      = calloc(total, 4);
	tr.sym         = calloc(total, 2);
	tr.first_child = calloc(total, 4);
	tr.next_sib    = calloc(total, 4);
	tr.root_child  = calloc(A, 4);
	tr.own_head    = malloc((size_t)total * 4);
	tr.own_next    = malloc((size_t)(c->npats + 1) * 4);
	if (!tr.parent || !tr.sym || !tr.first_child || !tr.next_sib ||
	    !tr.root_child || !tr.own_head || !tr.own_next) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	memset(tr.own_head, 0xff, (size_t)total * 4);

	/* 1. trie, reference insertion order: last added pattern first */
	for (k = c->npats - 1; k >= 0; k--) {
		const struct acm_pat *p = &c->pats[k];
		uint32_t state = 0, nx;
		int pos = 0;

		if (p->n == 0) {
			tr.own_next[k] = -1;
			continue;
		}
		for (; pos < p->n; pos++) {
			nx = trie_child(&tr, state, sym_at(c, p, pos));
			if (!nx)
				break;
			state = nx;
		}
		for (; pos < p->n; pos++) {
			unsigned a = sym_at(c, p, pos);
			last++;
			tr.parent[last] = state;
			tr.sym[last] = (uint16_t)a;
			if (state == 0) {
				tr.root_child[a] = last;
			} else {
				tr.next_sib[last] = tr.first_child[state];
				tr.first_child[state] = last;
			}
			state = last;
		}
		/* prepend: the own list ends up in ascending pattern index */
		tr.own_next[k] = tr.own_head[state];
		tr.own_head[state] = k;
	}
	tr.n = last + 1;

	/* 2. breadth-first numbering, children visited in symbol order
	 *    (the reference's BFS loops i = 0..ALPHABET_SIZE-1, acsmx.c:376-395) */
	order  = malloc((size_t)tr.n * 4);
	bfs_id = malloc((size_t)tr.n * 4);
	depth  = malloc((size_t)tr.n * 4);
	cbuf   = malloc((size_t)(A > 256 ? A : 256) * 4);
	t->level_start = calloc((size_t)c->max_len + 3, 4);
	if (!order || !bfs_id || !depth || !cbuf || !t->level_start) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	order[0] = 0;
	bfs_id[0] = 0;
	depth[0] = 0;
	qt = 1;
	for (i = 0; i < qt; i++) {
		uint32_t nc = 0, x, y;

		u = order[i];
		if (u == 0) {
			for (x = 0; x < (uint32_t)A; x++)
				if (tr.root_child[x])
					cbuf[nc++] = tr.root_child[x];
		} else {
			for (x = tr.first_child[u]; x; x = tr.next_sib[x])
				cbuf[nc++] = x;
			for (x = 1; x < nc; x++) {       /* insertion sort by symbol */
				uint32_t v = cbuf[x];
				for (y = x; y > 0 && tr.sym[cbuf[y - 1]] > tr.sym[v]; y--)
					cbuf[y] = cbuf[y - 1];
				cbuf[y] = v;
			}
		}
		for (x = 0; x < nc; x++) {
			uint32_t ch = cbuf[x];
			depth[ch] = depth[u] + 1;
			bfs_id[ch] = qt;
			order[qt++] = ch;
		}
	}
	t->alpha = A;
	t->num_states = tr.n;
	t->num_patterns = (uint32_t)c->npats;
	t->max_pattern_len = c->max_len;
	t->min_pattern_len = c->min_len;
	t->max_depth = c->max_len;
	{
		int d = 0;
		for (i = 0; i < tr.n; i++)
			while ((int)depth[order[i]] >= d)
				t->level_start[d++] = i;
		for (; d <= c->max_len + 1; d++)
			t->level_start[d] = tr.n;
	}

	/* 3. dense DFA rows + fail links, in BFS order */
	t->T          = malloc((size_t)tr.n * A * 4);
	t->fail       = calloc(tr.n, 4);
	t->olink      = calloc(tr.n, 4);
	t->own_begin  = calloc((size_t)tr.n + 1, 4);
	t->own_pat    = malloc((size_t)(c->npats + 1) * 4);
	t->pat_len    = malloc((size_t)(c->npats + 1) * 4);
	t->pat_iid    = malloc((size_t)(c->npats + 1) * 4);
	t->bfs_to_ref = malloc((size_t)tr.n * 4);
	if (!t->T || !t->fail || !t->olink || !t->own_begin || !t->own_pat ||
	    !t->pat_len || !t->pat_iid || !t->bfs_to_ref) {
		rc = ACM_ERR_NOMEM;
		goto out;
	}
	for (k = 0; k < c->npats; k++) {
		t->pat_len[k] = (uint32_t)c->pats[k].n;
		t->pat_iid[k] = c->pats[k].iid;
	}
	memset(t->T, 0, (size_t)A * 4);
	for (i = 0; i < tr.n; i++) {
		uint32_t *row = t->T + (size_t)i * A;
		const uint32_t *frow = t->T + (size_t)t->fail[i] * A;
		uint32_t x;

		u = order[i];
		t->bfs_to_ref[i] = u;
		if (i)
			memcpy(row, frow, (size_t)A * 4);
		if (u == 0) {
			for (x = 0; x < (uint32_t)A; x++)
				if (tr.root_child[x]) {
					uint32_t cid = bfs_id[tr.root_child[x]];
					row[x] = cid;
					t->fail[cid] = 0;
				}
		} else {
			for (x = tr.first_child[u]; x; x = tr.next_sib[x]) {
				uint32_t cid = bfs_id[x];
				t->fail[cid] = frow[tr.sym[x]];
				row[tr.sym[x]] = cid;
			}
		}
	}

	/* 4. own lists (CSR, BFS order) and output links */
	{
		uint32_t w = 0;
		for (i = 0; i < tr.n; i++) {
			int32_t e;
			t->own_begin[i] = w;
			for (e = tr.own_head[order[i]]; e != -1; e = tr.own_next[e])
				t->own_pat[w++] = (uint32_t)e;
		}
		t->own_begin[tr.n] = w;
		t->own_total = w;
		for (i = 1; i < tr.n; i++) {
			uint32_t f = t->fail[i];
			t->olink[i] = (t->own_begin[f + 1] > t->own_begin[f]) ? f
			                                                     : t->olink[f];
		}
	}

	/* 5. flag every edge with what its target reports */
	{
		uint8_t *flag = malloc(tr.n);
		size_t z, cells = (size_t)tr.n * A;

		if (!flag) {
			rc = ACM_ERR_NOMEM;
			goto out;
		}
		for (i = 0; i < tr.n; i++) {
			int own = t->own_begin[i + 1] > t->own_begin[i];
			flag[i] = (uint8_t)((own ? 2 : 0) | ((own || t->olink[i]) ? 1 : 0));
		}
		for (z = 0; z < cells; z++) {
			uint32_t v = t->T[z];
			uint8_t f = flag[v];
			if (f)
				t->T[z] = v | ((f & 2) ? ACM_T_OWN : 0) | ACM_T_ANY;
		}
		free(flag);
	}

	if (A == 256)
		rc = build_filters(c);
	if (A == 256 && rc == ACM_OK)
		rc = build_cdfa(c);
	if (rc == ACM_OK)
		rc = build_xd(t);

out:
	trie_free(&tr);
	free(order); free(bfs_id); free(depth); free(cbuf);
	if (rc != ACM_OK) {
		acm_tables_free(t);
		acm_set_error("compile failed (%d)", rc);
		return c->status = rc;
	}
	c->compiled = 1;
	return ACM_OK;
}

/*
 * Reference-layout table (reference acsmx.c:640-659): row stride 2*alpha ints,
 * first half = +/- next state (negative: target has a match list), second half =
 * index (bytes) or iid (ushorts, iacsmx.c:504) of the HEAD of the target's list.
 *
 * The reference's list for state s is reverse(list(fail s)) followed by the own
 * entries in ascending index (inherited copies are prepended one by one,
 * acsmx.c:417-429; own entries were prepended at insertion, acsmx.c:300-312,
 * in descending index order).  So
 *     head(s) = list(fail s) non-empty ? last(fail s) : min own index
 *     last(s) = own non-empty ? max own index : head(fail s)
 * States use the reference's numbering (trie creation order).
 */
int
acm_core_export_ref(struct acm_core *c, int **out)
{
	const struct acm_tables *t = &c->tab;
	const int A = c->alpha;
	const size_t row = 2 * (size_t)A;
	int32_t *head, *lastp;
	int *tab;
	uint32_t i;

	if (!c->compiled || !t->T) {
		acm_set_error("export_ref_table: automaton not compiled or already cleaned up");
		return ACM_ERR_STATE;
	}
	head = malloc((size_t)t->num_states * 4);
	lastp = malloc((size_t)t->num_states * 4);
	tab = memalign(4096, (size_t)t->num_states * row * sizeof(int));
	if (!head || !lastp || !tab) {
		free(head); free(lastp); free(tab);
		acm_set_error("export_ref_table: out of memory");
		return ACM_ERR_NOMEM;
	}
	head[0] = lastp[0] = -1;
	for (i = 1; i < t->num_states; i++) {
		uint32_t f = t->fail[i];
		int own = t->own_begin[i + 1] > t->own_begin[i];
		int inh = head[f] != -1;

		head[i] = inh ? lastp[f] : (own ? (int32_t)t->own_pat[t->own_begin[i]] : -1);
		lastp[i] = own ? (int32_t)t->own_pat[t->own_begin[i + 1] - 1]
		               : (inh ? head[f] : -1);
	}
	memset(tab, 0, (size_t)t->num_states * row * sizeof(int));
	for (i = 0; i < t->num_states; i++) {
		int *r = tab + (size_t)t->bfs_to_ref[i] * row;
		const uint32_t *src = t->T + (size_t)i * A;
		for (int a = 0; a < A; a++) {
			uint32_t nx = src[a] & ACM_T_MASK;
			int ref = (int)t->bfs_to_ref[nx];
			if (head[nx] != -1) {
				r[a] = -ref;
				r[A + a] = (A == 256) ? head[nx] : t->pat_iid[head[nx]];
			} else {
				r[a] = ref;
			}
		}
	}
	free(head);
	free(lastp);
	*out = tab;
	return ACM_OK;
}

void
acm_core_cleanup(struct acm_core *c)
{
	int k;

	for (k = 0; k < c->npats; k++) {
		free(c->pats[k].syms);
		c->pats[k].syms = NULL;
	}
	acm_tables_free(&c->tab);
}

void
acm_core_free(struct acm_core *c)
{
	if (!c)
		return;
	acm_core_cleanup(c);
	free(c->pats);
	free(c);
}
