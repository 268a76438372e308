/*
 * compat_ocl.c -- the ocl_* entry points of the reference, over acm.h.
 *
 *   clinitctx            reference ocl_context.c:19
 *   ocl_aho_match*       reference ocl_aho_match.c:13-131, AC_ushorts/ocl_aho_match.c
 *   ocl_prefix_sum*      reference ocl_prefix_sum.c:70-498
 *   ocl_compact_array*   reference ocl_compact_array.c:14-172
 *   ocl_bitonic_sort*    reference ocl_bitonic_sort.c:24-251
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/acm.h"
#include "../../include/ocl_aho_match.h"
#include "../../include/ocl_bitonic_sort.h"
#include "../../include/ocl_compact_array.h"
#include "../../include/ocl_context.h"
#include "../../include/ocl_prefix_sum.h"
#include "acm_core.h"
#include "acm_tables.h"
#include "acm_queue.h"
#include "databuf_priv.h"

/* cudaMemcpy D2D without pulling the CUDA headers into C code */
int acm_memcpy_d2d(struct acm_device *, void *d_dst, const void *d_src, size_t bytes);

void
clinitctx(struct clconf *c, int pos, int subpos)
{
	struct acm_device *dev = NULL;
	struct acm_queue *q;

	(void)subpos;
	memset(c, 0, sizeof(*c));
	if (acm_device_open(pos, &dev) != ACM_OK)
		return;
	q = calloc(1, sizeof(*q));
	if (!q) {
		acm_device_close(dev);
		acm_set_error("clinitctx: out of memory");
		return;
	}
	q->dev = dev;
	c->ctx = (cl_context)dev;
	c->queue = q;
	c->type = CL_DEVICE_TYPE_GPU;
}

void
clfreectx(struct clconf *c)
{
	if (!c)
		return;
	if (c->ctx)
		acm_device_close((struct acm_device *)c->ctx);
	free(c->queue);
	memset(c, 0, sizeof(*c));
}

void ocl_aho_match_init(struct clconf *c) { (void)c; }
void ocl_aho_match_close(struct clconf *c) { (void)c; }
void ocl_prefix_sum_init(struct clconf *c) { (void)c; }
void ocl_prefix_sum_close(struct clconf *c) { (void)c; }
void ocl_compact_array_init(struct clconf *c) { (void)c; }
void ocl_compact_array_close(struct clconf *c) { (void)c; }
int  ocl_bitonic_sort_init(struct clconf *c) { (void)c; return 0; }
int  ocl_bitonic_sort_close(struct clconf *c) { (void)c; return 0; }

/* the databuf's scanner for this automaton (one per databuf, replaced when the automaton changes) */
static int
ensure_scanner(struct databuf *db, struct acm_automaton *aut, int sym_size)
{
	struct databuf_priv *pv = (struct databuf_priv *)db->priv;
	struct acm_scan_params p;
	int rc;

	if (pv->scanner && pv->scanner_aut != aut) {
		acm_scanner_free(pv->scanner);
		pv->scanner = NULL;
		pv->carry_len = 0;
	}
	if (pv->scanner)
		return ACM_OK;
	memset(&p, 0, sizeof(p));
	/* 32 KiB result buckets for the sparse-output kernels; the dense-output kernels (word
	 * lists: a bucket is one thread's chunk there) keep their own shape */
	if (acm_automaton_default_mode(aut) != ACM_MODE_CDFA)
		p.bucket_shift = 15;
	rc = acm_scanner_create(pv->dev, aut, db->size / (uint64_t)sym_size + 1, &p, &pv->scanner);
	if (rc == ACM_OK)
		pv->scanner_aut = aut;
	return rc;
}

/*
 * What the reference does in its init calls -- build the OpenCL program, create the kernels
 * (ocl_aho_match.c:25-60, called from ocl_worker_ctx_init before any file is opened) -- has its
 * counterpart here: create the scanner (result buckets, queues) and run the kernels once over a few
 * KiB of the still empty device buffer, so that the driver loads their code now and not inside the
 * first buffer of the scan loop.  Nothing of the databuf's state changes.
 */
int
databuf_prepare(struct databuf *db, struct acm_automaton *aut, int sym_size)
{
	struct databuf_priv *pv = db ? (struct databuf_priv *)db->priv : NULL;
	const uint64_t carry_cap = DATABUF_CARRY_CAP / (uint64_t)sym_size;
	uint64_t n = 65536;
	struct acm_scan_result res;
	int rc;

	if (!pv || !aut)
		return ACM_ERR_ARG;
	if ((rc = ensure_scanner(db, aut, sym_size)) != ACM_OK)
		return rc;
	if (n > db->size / (uint64_t)sym_size)
		n = db->size / (uint64_t)sym_size;
	if (n == 0)
		return ACM_OK;
	if ((rc = acm_dev_memset(pv->dev, pv->d_base + DATABUF_CARRY_CAP, 0, (size_t)n * sym_size)) != ACM_OK)
		return rc;
	rc = acm_scan_device_ex(pv->scanner, pv->d_base, carry_cap + n, carry_cap, carry_cap, carry_cap + n, &res);
	if (rc == ACM_OK)
		rc = acm_device_sync(pv->dev);
	return rc;
}

static void
match_common(struct databuf *db, struct acm_automaton *aut, int sym_size, int stream)
{
	struct databuf_priv *pv = (struct databuf_priv *)db->priv;
	const uint64_t carry_cap = DATABUF_CARRY_CAP / (uint64_t)sym_size;
	const uint64_t nsym = db->bytes / (uint64_t)sym_size;
	struct acm_scan_result res;
	uint64_t halo, keep;
	int rc;

	pv->n_matches = pv->n_valid = 0;
	pv->fetched = 0;
	pv->buckets_on_device = 0;
	if (!aut) {
		acm_set_error("ocl_aho_match: automaton has no device tables (acsm_gen_state_table not called or failed)");
		pv->status = ACM_ERR_STATE;
		return;
	}
	pv->sym_size = sym_size;
	if ((rc = ensure_scanner(db, aut, sym_size)) != ACM_OK) {
		pv->status = rc;
		return;
	}
	halo = (uint64_t)(acm_automaton_max_pattern_len(aut) > 0 ? acm_automaton_max_pattern_len(aut) - 1 : 0);
	if (halo > carry_cap)
		halo = carry_cap;
	if (!stream) {
		pv->carry_len = 0;
		pv->carry_real = 0;
		pv->carry_file = -1;
	}
	pv->filt_carry_real = pv->carry_real;
	pv->filt_carry_file = pv->carry_file;
	if (nsym == 0)
		return;
	rc = acm_scan_device_ex(pv->scanner, pv->d_base, carry_cap + nsym, carry_cap - pv->carry_len,
	    carry_cap, carry_cap + nsym, &res);
	if (rc != ACM_OK) {
		pv->status = rc;
		return;
	}
	pv->n_matches = res.n_matches;
	pv->have_match = 1;

	/* new carry = last `halo` symbols of (old carry + this buffer), right-aligned before d_data */
	keep = pv->carry_len + nsym;
	if (keep > halo)
		keep = halo;
	if (stream && keep) {
		unsigned char *data = pv->d_base + DATABUF_CARRY_CAP;
		const size_t kb = (size_t)keep * sym_size, nb = (size_t)nsym * sym_size;
		/* source range [data + nb - kb, data + nb) may reach back into the old carry */
		rc = acm_memcpy_d2d(pv->dev, pv->d_carry_tmp, data + nb - kb, kb);
		if (rc == ACM_OK)
			rc = acm_memcpy_d2d(pv->dev, data - kb, pv->d_carry_tmp, kb);
		if (rc != ACM_OK) {
			pv->status = rc;
			return;
		}
	}
	pv->carry_len = stream ? keep : 0;
	/*
	 * Per-file semantics: which of the carried symbols are real, contiguous bytes of the file the
	 * buffer ends with -- the stretch of chunks of that file that runs up to the very end of the
	 * buffer (plus, if it also starts the buffer and continues the previous carry's file, that
	 * carry).  A padded last chunk, or nothing carried, leaves 0: nothing of the next buffer may
	 * reach back.
	 */
	if (stream && keep && db->chunks > 0) {
		size_t c = db->chunks - 1;
		const int fid = db->file_ids[c];
		uint64_t real = 0;
		if ((size_t)db->h_indices[c] + (size_t)db->h_sizes[c] == db->bytes) {
			real = (uint64_t)db->h_sizes[c];
			while (c > 0 && db->file_ids[c - 1] == fid &&
			    db->h_indices[c - 1] + db->h_sizes[c - 1] == db->h_indices[c]) {
				c--;
				real += (uint64_t)db->h_sizes[c];
			}
			if (c == 0 && db->h_indices[0] == 0 && pv->carry_file == fid)
				real += pv->carry_real * (uint64_t)sym_size;
		}
		real /= (uint64_t)sym_size;
		pv->carry_real = real < keep ? real : keep;
		pv->carry_file = fid;
	} else {
		pv->carry_real = 0;
		pv->carry_file = -1;
	}
	/* the reference blocks in clFinish (ocl_aho_match.c:128) */
	pv->status = acm_device_sync(pv->dev);
}

void
ocl_aho_match(struct clconf *cl, struct databuf *db, acsm_t *acsm, size_t local_ws, int stream)
{
	(void)cl;
	(void)local_ws;
	match_common(db, acsm_device_automaton(acsm), 1, stream);
}

void
ocl_aho_match_ushort(struct clconf *cl, struct databuf *db, iacsm_t *iacsm, size_t local_ws)
{
	(void)cl;
	(void)local_ws;
	match_common(db, iacsm_device_automaton(iacsm), 2, 1);
}

/*
 * The reference's COMPACT_RESULTS order (databuf.c:648-651) is ocl_aho_match -> ocl_prefix_sum ->
 * ocl_compact_array -> databuf_copy_device_to_host, with the match kernel itself filling the
 * device buckets.  Here the device holds one sorted list, so the bucket view the two post-passes
 * work on is built from it first (the host does that in databuf_copy_device_to_host) and
 * uploaded: they never see memory nobody wrote.
 */
static int
upload_bucket_view(struct databuf *db)
{
	struct databuf_priv *pv = (struct databuf_priv *)db->priv;
	const size_t nres = ((size_t)db->max_results * db->max_chunks + 1) * sizeof(int);
	int rc;

	/* no match has run on this databuf: the buckets on the device are the caller's own
	 * (the reference's self-test, databuf.c:935-1021, fills them by hand) */
	if (pv->buckets_on_device || !pv->have_match)
		return ACM_OK;
	if (!pv->fetched)
		databuf_copy_device_to_host(db, NULL);
	if (pv->status != ACM_OK)
		return pv->status;
	/* rows 1.. of the host view are only valid below a chunk's count: clear the device copy first */
	if ((rc = acm_dev_memset(pv->dev, db->d_results, 0, nres)) != ACM_OK ||
	    (rc = acm_dev_memset(pv->dev, db->d_results2, 0, nres)) != ACM_OK)
		return pv->status = rc;
	{
		const size_t C = db->chunks, R = (size_t)db->max_results;
		size_t k;
		/* the count row, the rows in use, and the tail word (last state) */
		for (k = 0; k < R && C; k++) {
			if ((rc = acm_memcpy_h2d(pv->dev, (int *)db->d_results + k * C, db->h_results + k * C, C * sizeof(int))) != ACM_OK ||
			    (rc = acm_memcpy_h2d(pv->dev, (int *)db->d_results2 + k * C, db->h_results2 + k * C, C * sizeof(int))) != ACM_OK)
				return pv->status = rc;
		}
		if ((rc = acm_memcpy_h2d(pv->dev, (int *)db->d_results + R * C, db->h_results + R * C, sizeof(int))) != ACM_OK)
			return pv->status = rc;
	}
	if ((rc = acm_device_sync(pv->dev)) != ACM_OK)
		return pv->status = rc;
	pv->buckets_on_device = 1;
	return ACM_OK;
}

void
ocl_prefix_sum(struct clconf *cl, struct databuf *db, unsigned int n)
{
	struct databuf_priv *pv = (struct databuf_priv *)db->priv;

	(void)cl;
	if (databuf_alloc_postpass(db) != ACM_OK || upload_bucket_view(db) != ACM_OK)
		return;
	if (n > db->max_chunks)
		n = (unsigned int)db->max_chunks;
	pv->status = acm_exclusive_scan_u32(pv->dev, (const uint32_t *)db->d_results,
	    (uint32_t *)db->d_prefixsum, n, NULL);
	if (pv->status == ACM_OK)
		pv->status = acm_device_sync(pv->dev);
}

void
ocl_compact_array(struct clconf *cl, struct databuf *db, size_t local)
{
	struct databuf_priv *pv = (struct databuf_priv *)db->priv;

	(void)cl;
	(void)local;
	if (databuf_alloc_postpass(db) != ACM_OK || db->chunks == 0 || upload_bucket_view(db) != ACM_OK)
		return;
	pv->status = acm_compact_columns_i32(pv->dev, (int32_t *)db->d_results_comp,
	    (const int32_t *)db->d_results, (const int32_t *)db->d_prefixsum, (int32_t)db->chunks,
	    db->max_results, (int64_t)db->size + 2);
	if (pv->status == ACM_OK)
		pv->status = acm_compact_columns_i32(pv->dev, (int32_t *)db->d_results2_comp,
		    (const int32_t *)db->d_results2, (const int32_t *)db->d_prefixsum, (int32_t)db->chunks,
		    db->max_results, (int64_t)db->size + 2);
	if (pv->status == ACM_OK)
		pv->status = acm_device_sync(pv->dev);
}

int
ocl_bitonic_sort(struct clconf *cl, cl_mem dst_key, cl_mem dst_val, cl_mem src_key, cl_mem src_val,
    unsigned int batch, unsigned int len, unsigned int dir)
{
	struct acm_device *dev = acm_queue_device(cl ? cl->ctx : NULL, cl ? cl->queue : NULL);
	unsigned int b;
	int rc;

	if (!dev)
		return -1;
	if (len < 2)
		return -2;              /* reference ocl_bitonic_sort.c:146-148 */
	for (b = 0; b < batch; b++) {
		const size_t o = (size_t)b * len;
		rc = acm_sort_pairs_u32(dev, (uint32_t *)dst_key + o, (uint32_t *)dst_val + o,
		    (const uint32_t *)src_key + o, (const uint32_t *)src_val + o, len, dir == 0);
		if (rc != ACM_OK)
			return rc;
	}
	return 0;
}
