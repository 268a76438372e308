/*
 * acm_multi.c -- several GPUs in one process: one host thread per device over the native API of
 * acm.h (plain C, pthreads; no torch, no MPI, no NCCL -- the path has no collective).
 *
 * Replaces, for one stream, what the reference does with `-w` worker threads that each own a
 * context on device `dev_pos` (reference ocl_aho_grep.c:498-502, ocl_worker.c:32).  Partition:
 * SURVEY.md 8(e) -- contiguous ranges, Lmax-1 symbols of leading context, replicated automaton,
 * rank-order concatenation.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/acm.h"
#include "acm_core.h"

enum { JOB_NONE, JOB_SCAN_DEVICE, JOB_SCAN_HOST, JOB_QUIT };

struct acm_multi;

struct multi_worker {
	struct acm_multi     *m;
	int                   g;
	struct acm_device    *dev;
	struct acm_automaton *aut;
	struct acm_scanner   *scanner;
	pthread_t             thread;
	int                   rc;            /* of the last job                         */
	uint64_t              count;         /* matches of the last job                 */
	struct acm_scan_result res;
	uint64_t             *off;           /* JOB_SCAN_HOST: private result arrays    */
	uint32_t             *pat;
	uint64_t              cap;
};

struct acm_multi {
	int                  n;
	int                  n_threads;      /* workers whose thread exists                */
	int                  halo;           /* Lmax - 1                                 */
	struct multi_worker *w;
	pthread_mutex_t      lock;
	pthread_cond_t       cond;
	pthread_barrier_t    bar;            /* the n workers, inside a job              */
	unsigned             generation;     /* bumped when a job is posted              */
	int                  done;           /* workers that finished the current job    */
	int                  job;
	/* job arguments */
	const void *const   *d_data;
	const void          *h_data;
	uint64_t             total, base, cap;
	uint64_t            *h_keys, *h_off;
	uint32_t            *h_pat;
	uint64_t             offsets[65];    /* exclusive scan of the per-device counts   */
};

void
acm_multi_shard(const struct acm_multi *m, uint64_t total, int g, uint64_t *read_lo, uint64_t *lo, uint64_t *hi)
{
	const uint64_t a = g >= m->n ? total : (total * (uint64_t)g / (uint64_t)m->n) & ~(uint64_t)15;
	const uint64_t b = g + 1 >= m->n ? total : (total * (uint64_t)(g + 1) / (uint64_t)m->n) & ~(uint64_t)15;
	const uint64_t h = (uint64_t)m->halo;

	*lo = a;
	*hi = b;
	*read_lo = (a > h ? a - h : 0) & ~(uint64_t)15;
}

static void
job_scan_device(struct multi_worker *w)
{
	struct acm_multi *m = w->m;
	uint64_t read_lo, lo, hi, i;

	acm_multi_shard(m, m->total, w->g, &read_lo, &lo, &hi);
	w->count = 0;
	w->rc = ACM_OK;
	if (hi > lo)
		w->rc = acm_scan_device(w->scanner, m->d_data[w->g], hi - read_lo, lo - read_lo, hi - read_lo, &w->res);
	if (w->rc == ACM_OK && hi > lo)
		w->count = w->res.n_matches;
	/* counts -> offsets: every worker computes the same scan after the barrier */
	pthread_barrier_wait(&m->bar);
	{
		uint64_t off = 0;
		for (int g = 0; g < w->g; g++)
			off += m->w[g].count;
		if (w->rc == ACM_OK && w->count && off < m->cap) {
			const uint64_t take = off + w->count <= m->cap ? w->count : m->cap - off;
			/* this device's slice, over this device's PCIe link */
			w->rc = acm_memcpy_d2h(w->dev, m->h_keys + off, acm_scan_keys(w->scanner), take * 8);
			if (w->rc == ACM_OK)
				w->rc = acm_device_sync(w->dev);
			/* offsets are relative to the shard's buffer: shift them to stream positions */
			for (i = 0; w->rc == ACM_OK && i < take; i++)
				m->h_keys[off + i] += read_lo << 24;
		}
	}
}

static void
job_scan_host(struct multi_worker *w)
{
	struct acm_multi *m = w->m;
	const uint64_t sym = (uint64_t)(acm_automaton_alphabet(w->aut) == 256 ? 1 : 2);
	uint64_t read_lo, lo, hi;
	int64_t got;

	acm_multi_shard(m, m->total, w->g, &read_lo, &lo, &hi);
	w->count = 0;
	w->rc = ACM_OK;
	if (hi > lo) {
		for (;;) {
			got = acm_scan_host_ex(w->scanner, (const unsigned char *)m->h_data + lo * sym, hi - lo, lo - read_lo,
			    m->base + lo, w->off, w->pat, w->cap, &w->res);
			if (got < 0) {
				w->rc = (int)got;
				break;
			}
			if ((uint64_t)got <= w->cap) {
				w->count = (uint64_t)got;
				break;
			}
			/* private arrays too small: grow and scan the shard again */
			free(w->off);
			free(w->pat);
			w->cap = (uint64_t)got + (uint64_t)got / 4 + 1024;
			w->off = malloc(w->cap * sizeof(uint64_t));
			w->pat = malloc(w->cap * sizeof(uint32_t));
			if (!w->off || !w->pat) {
				w->cap = 0;
				w->rc = ACM_ERR_NOMEM;
				break;
			}
		}
	}
	pthread_barrier_wait(&m->bar);
	{
		uint64_t off = 0;
		for (int g = 0; g < w->g; g++)
			off += m->w[g].count;
		if (w->rc == ACM_OK && w->count && off < m->cap) {
			const uint64_t take = off + w->count <= m->cap ? w->count : m->cap - off;
			memcpy(m->h_off + off, w->off, take * sizeof(uint64_t));
			memcpy(m->h_pat + off, w->pat, take * sizeof(uint32_t));
		}
	}
}

static void *
worker_main(void *arg)
{
	struct multi_worker *w = arg;
	struct acm_multi *m = w->m;
	unsigned seen = 0;

	for (;;) {
		int job;
		pthread_mutex_lock(&m->lock);
		while (m->generation == seen)
			pthread_cond_wait(&m->cond, &m->lock);
		seen = m->generation;
		job = m->job;
		pthread_mutex_unlock(&m->lock);
		if (job == JOB_QUIT)
			return NULL;
		if (job == JOB_SCAN_DEVICE)
			job_scan_device(w);
		else if (job == JOB_SCAN_HOST)
			job_scan_host(w);
		pthread_mutex_lock(&m->lock);
		m->done++;
		pthread_cond_broadcast(&m->cond);
		pthread_mutex_unlock(&m->lock);
	}
}

/* post a job to every worker and wait for all of them */
static void
run_job(struct acm_multi *m, int job)
{
	pthread_mutex_lock(&m->lock);
	m->job = job;
	m->done = 0;
	m->generation++;
	pthread_cond_broadcast(&m->cond);
	while (m->done < m->n)
		pthread_cond_wait(&m->cond, &m->lock);
	pthread_mutex_unlock(&m->lock);
}

int
acm_multi_open(const struct acm_tables *t, const int *ordinals, int n, uint64_t max_bytes_per_device,
    const struct acm_scan_params *params, struct acm_multi **out)
{
	struct acm_multi *m;
	int g, rc = ACM_OK;

	*out = NULL;
	if (!t || !ordinals || n < 1 || n > 64) {
		acm_set_error("acm_multi_open: 1 .. 64 devices and a compiled automaton required");
		return ACM_ERR_ARG;
	}
	m = calloc(1, sizeof(*m));
	if (m)
		m->w = calloc((size_t)n, sizeof(*m->w));
	if (!m || !m->w) {
		free(m);
		return ACM_ERR_NOMEM;
	}
	m->n = n;
	m->halo = t->max_pattern_len > 0 ? t->max_pattern_len - 1 : 0;
	pthread_mutex_init(&m->lock, NULL);
	pthread_cond_init(&m->cond, NULL);
	pthread_barrier_init(&m->bar, NULL, (unsigned)n);
	for (g = 0; g < n && rc == ACM_OK; g++) {
		struct multi_worker *w = &m->w[g];
		w->m = m;
		w->g = g;
		if ((rc = acm_device_open(ordinals[g], &w->dev)) != ACM_OK)
			break;
		if ((rc = acm_automaton_upload(w->dev, t, &w->aut)) != ACM_OK)
			break;
		/* the shard plus its leading context, rounded the way acm_multi_shard rounds */
		rc = acm_scanner_create(w->dev, w->aut, max_bytes_per_device + (uint64_t)m->halo + 64, params, &w->scanner);
	}
	if (rc != ACM_OK) {
		acm_multi_close(m);
		return rc;
	}
	for (g = 0; g < n; g++) {
		if (pthread_create(&m->w[g].thread, NULL, worker_main, &m->w[g]) != 0) {
			acm_set_error("acm_multi_open: cannot start the thread of device %d", g);
			acm_multi_close(m);
			return ACM_ERR_NOMEM;
		}
		m->n_threads = g + 1;
	}
	*out = m;
	return ACM_OK;
}

void
acm_multi_close(struct acm_multi *m)
{
	int g;

	if (!m)
		return;
	if (m->n_threads) {
		pthread_mutex_lock(&m->lock);
		m->job = JOB_QUIT;
		m->generation++;
		pthread_cond_broadcast(&m->cond);
		pthread_mutex_unlock(&m->lock);
		for (g = 0; g < m->n_threads; g++)
			pthread_join(m->w[g].thread, NULL);
	}
	for (g = 0; g < m->n; g++) {
		struct multi_worker *w = &m->w[g];
		if (w->scanner)
			acm_scanner_free(w->scanner);
		if (w->aut)
			acm_automaton_free(w->aut);
		if (w->dev)
			acm_device_close(w->dev);
		free(w->off);
		free(w->pat);
	}
	pthread_barrier_destroy(&m->bar);
	pthread_mutex_destroy(&m->lock);
	pthread_cond_destroy(&m->cond);
	free(m->w);
	free(m);
}

int
acm_multi_devices(const struct acm_multi *m)
{
	return m->n;
}

struct acm_device *
acm_multi_device(struct acm_multi *m, int g)
{
	return g >= 0 && g < m->n ? m->w[g].dev : NULL;
}

static int64_t
collect(struct acm_multi *m, uint64_t *counts, struct acm_scan_result *res)
{
	uint64_t total = 0;
	int g;

	if (res)
		memset(res, 0, sizeof(*res));
	for (g = 0; g < m->n; g++) {
		if (m->w[g].rc != ACM_OK)
			return m->w[g].rc;
		if (counts)
			counts[g] = m->w[g].count;
		total += m->w[g].count;
		if (res) {
			res->n_bytes += m->w[g].res.n_bytes;
			res->mode = m->w[g].res.mode;
			res->fallback |= m->w[g].res.fallback;
			res->launches += m->w[g].res.launches;
			if (m->w[g].res.ms_scan > res->ms_scan)
				res->ms_scan = m->w[g].res.ms_scan;
		}
	}
	if (res)
		res->n_matches = total;
	return (int64_t)total;
}

int64_t
acm_multi_scan_device(struct acm_multi *m, const void *const *d_data, uint64_t total, uint64_t *h_keys,
    uint64_t cap, uint64_t *counts, struct acm_scan_result *res)
{
	m->d_data = d_data;
	m->total = total;
	m->h_keys = h_keys;
	m->cap = cap;
	run_job(m, JOB_SCAN_DEVICE);
	return collect(m, counts, res);
}

int64_t
acm_multi_scan_host(struct acm_multi *m, const void *h_data, uint64_t n, uint64_t base, uint64_t *h_off,
    uint32_t *h_pat, uint64_t cap, struct acm_scan_result *res)
{
	int g;

	for (g = 0; g < m->n; g++) {
		struct multi_worker *w = &m->w[g];
		if (!w->cap) {
			w->cap = 1 << 16;
			w->off = malloc(w->cap * sizeof(uint64_t));
			w->pat = malloc(w->cap * sizeof(uint32_t));
			if (!w->off || !w->pat) {
				w->cap = 0;
				return ACM_ERR_NOMEM;
			}
		}
	}
	m->h_data = h_data;
	m->total = n;
	m->base = base;
	m->h_off = h_off;
	m->h_pat = h_pat;
	m->cap = cap;
	run_job(m, JOB_SCAN_HOST);
	return collect(m, NULL, res);
}
