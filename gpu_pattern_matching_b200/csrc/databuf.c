/*
 * databuf.c -- chunked input buffer and result decoding (host side, plain C).
 *
 * Replaces reference databuf.c:77-843.  The chunk bookkeeping (fixed-size chunks
 * read straight into the host buffer, zero-padded tail chunk, one chunk per line
 * in text mode, the return codes) follows the reference function by function; the
 * device side is one contiguous stream with a byte carry in front of it, scanned
 * through acm.h.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <pthread.h>
#include <sys/stat.h>
#include <sys/types.h>

#include "../../include/acm.h"
#include "../../include/databuf.h"
#include "acm_core.h"
#include "acm_queue.h"
#include "databuf_priv.h"

#define ROUNDUP16(x) (((x) + 15) & ~(size_t)15)
#define MIN(a, b) ((a) < (b) ? (a) : (b))

static struct databuf_priv *
priv_of(struct databuf *db)
{
	return (struct databuf_priv *)db->priv;
}

#define READ_PAR_MIN   ((size_t)8 << 20)      /* below this one read() is as fast */
#define READ_PAR_MAX   16


struct databuf *
databuf_new(size_t max_chunks, size_t max_chunk_size, int max_results, int mapped, struct clconf *conf)
{
	struct databuf *db;
	struct databuf_priv *pv;
	struct acm_device *dev = acm_queue_device(conf ? conf->ctx : NULL, conf ? conf->queue : NULL);
	size_t i, nres;
	void *p;

	if (!dev || max_chunks == 0 || max_chunk_size == 0 || max_results < 2) {
		if (dev)
			acm_set_error("databuf_new: bad shape (%zu chunks x %zu bytes, %d results)",
			    max_chunks, max_chunk_size, max_results);
		return NULL;
	}
	if (max_chunks * max_chunk_size > (size_t)0x7fffffff) {
		acm_set_error("databuf_new: buffer larger than 2 GiB (int offsets, as in the reference)");
		return NULL;
	}
	db = calloc(1, sizeof(*db));
	pv = calloc(1, sizeof(*pv));
	if (!db || !pv)
		goto fail;
	db->priv = pv;
	pv->dev = dev;
	pv->sym_size = 1;
	{
		/* ACM_DATABUF_STREAM_QUIRK=1: one stream across files and padding, as the reference scans it */
		const char *q = getenv("ACM_DATABUF_STREAM_QUIRK");
		pv->file_semantics = !(q && atoi(q));
	}
	pv->carry_file = -1;
	pthread_mutex_init(&pv->ra.lock, NULL);
	pthread_cond_init(&pv->ra.cond, NULL);
	db->cl = conf;
	db->mapped = mapped;
	db->max_results = max_results;
	db->max_chunks = max_chunks;
	db->max_chunk_size = max_chunk_size;
	db->size = max_chunks * max_chunk_size;
	nres = (size_t)max_results * max_chunks + 1;

	if (acm_host_alloc_pinned_near(dev, db->size + 64, &p) != ACM_OK)
		goto fail;
	db->h_data = p;
	/*
	 * The second pinned buffer of the read-ahead, here and not on first use: pinning 128 MiB takes
	 * ~50 ms, which inside the scan loop cost more than the overlap gains on anything below a few GiB
	 * per worker (cli_bench, 4 x 512 MiB: 25 -> 6 GB/s).  Buffers too small for a parallel read do
	 * not read ahead; ACM_READAHEAD=0 turns it off.
	 */
	{
		const char *env = getenv("ACM_READAHEAD");
		if (!(env && atoi(env) == 0) && db->size >= READ_PAR_MIN &&
		    acm_host_alloc_pinned_near(dev, db->size + 64, &p) == ACM_OK)
			pv->ra.buf = p;
	}
	if (acm_dev_alloc(dev, DATABUF_CARRY_CAP + db->size + 64, &p) != ACM_OK)
		goto fail;
	pv->d_base = p;
	db->d_data = pv->d_base + DATABUF_CARRY_CAP;
	if (acm_dev_alloc(dev, DATABUF_CARRY_CAP, &p) != ACM_OK)
		goto fail;
	pv->d_carry_tmp = p;

	db->h_indices = malloc(max_chunks * sizeof(int));
	db->h_sizes = malloc(max_chunks * sizeof(int));
	db->file_ids = malloc(max_chunks * sizeof(int));
	db->h_results = calloc(nres, sizeof(int));
	db->h_results2 = calloc(nres, sizeof(int));
	db->h_prefixsum = calloc(max_chunks, sizeof(int));
	pv->run_id = malloc(max_chunks * sizeof(int));
	db->results_comp_size = db->results2_comp_size = MIN(db->size + 2, (size_t)1 << 16);
	db->h_results_comp = calloc(db->results_comp_size, sizeof(int));
	db->h_results2_comp = calloc(db->results2_comp_size, sizeof(int));
	if (!db->h_indices || !db->h_sizes || !db->file_ids || !db->h_results || !db->h_results2 ||
	    !db->h_prefixsum || !db->h_results_comp || !db->h_results2_comp || !pv->run_id) {
		acm_set_error("databuf_new: out of memory");
		goto fail;
	}
	/* reference databuf.c:312-316 */
	for (i = 0; i < max_chunks; i++) {
		db->h_sizes[i] = (int)max_chunk_size;
		db->h_indices[i] = (int)(max_chunk_size * i);
		db->file_ids[i] = -1;
	}
	return db;
fail:
	if (db)
		databuf_free(db, mapped, conf ? conf->queue : NULL);
	else
		free(pv);
	return NULL;
}

/* ---- parallel read of a regular file ---- */


struct read_seg {
	int     fd;
	char   *dst;
	off_t   off;
	size_t  want;
	ssize_t got;                              /* bytes read, contiguous from off; -1 on error */
};

static void *
read_seg_main(void *arg)
{
	struct read_seg *s = arg;
	size_t done = 0;

	while (done < s->want) {
		const ssize_t r = pread(s->fd, s->dst + done, s->want - done, s->off + (off_t)done);
		if (r < 0) {
			s->got = done ? (ssize_t)done : -1;
			return NULL;
		}
		if (r == 0)
			break;                            /* end of file */
		done += (size_t)r;
	}
	s->got = (ssize_t)done;
	return NULL;
}

static int
read_threads(void)
{
	const char *env = getenv("ACM_READ_THREADS");
	int nt = env ? atoi(env) : 4;

	return nt > READ_PAR_MAX ? READ_PAR_MAX : nt;
}

/*
 * [off, off + want) of a regular file into buf with nt threads pread()ing adjacent segments; the file
 * offset is not touched.  Returns the contiguous bytes read from `off` on (short at the end of the
 * file), -1 when nothing could be read because of an error.
 */
static long
read_segments(int fd, void *buf, off_t off, size_t want, int nt)
{
	struct read_seg seg[READ_PAR_MAX];
	pthread_t th[READ_PAR_MAX];
	int started = 0, i;
	size_t per, total = 0;

	if (nt < 1)
		nt = 1;
	per = ((want + (size_t)nt - 1) / (size_t)nt + 4095) & ~(size_t)4095;
	for (i = 0; i < nt && (size_t)i * per < want; i++) {
		seg[i].fd = fd;
		seg[i].dst = (char *)buf + (size_t)i * per;
		seg[i].off = off + (off_t)((size_t)i * per);
		seg[i].want = want - (size_t)i * per < per ? want - (size_t)i * per : per;
		seg[i].got = 0;
	}
	nt = i;
	for (i = 1; i < nt; i++) {
		if (pthread_create(&th[i], NULL, read_seg_main, &seg[i]) != 0)
			break;
		started = i;
	}
	read_seg_main(&seg[0]);
	for (i = 1; i <= started; i++)
		pthread_join(th[i], NULL);
	for (i = started + 1; i < nt; i++)        /* threads that could not be created: read here */
		read_seg_main(&seg[i]);
	/* the contiguous prefix: everything up to and including the first short segment */
	for (i = 0; i < nt; i++) {
		if (seg[i].got < 0)
			break;
		total += (size_t)seg[i].got;
		if ((size_t)seg[i].got < seg[i].want)
			break;
	}
	if (total == 0 && nt > 0 && seg[0].got < 0)
		return -1;
	return (long)total;
}

long
databuf_read_fd(int fd, void *buf, size_t want)
{
	struct stat st;
	const int nt = read_threads();
	off_t off;
	long total;

	if (nt < 2 || want < READ_PAR_MIN || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) ||
	    (off = lseek(fd, 0, SEEK_CUR)) == (off_t)-1)
		return (long)read(fd, buf, want);
	if (off >= st.st_size)
		return (long)read(fd, buf, want);     /* at (what was) the end: 0, or freshly appended data */
	if ((size_t)(st.st_size - off) < want)
		want = (size_t)(st.st_size - off);    /* what the file holds now; a later call sees what is appended */
	if (want < READ_PAR_MIN)
		return (long)read(fd, buf, want);
	total = read_segments(fd, buf, off, want, nt);
	if (total < 0)
		return -1;
	if (lseek(fd, off + (off_t)total, SEEK_SET) == (off_t)-1)
		return -1;
	return total;
}

/* ---- read-ahead of the next whole buffer (see databuf_priv.h) ---- */

static double
now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *
readahead_main(void *arg)
{
	struct databuf_readahead *ra = arg;

	pthread_mutex_lock(&ra->lock);
	for (;;) {
		while (!ra->busy && !ra->quit)
			pthread_cond_wait(&ra->cond, &ra->lock);
		if (ra->quit)
			break;
		{
			const int fd = ra->fd;
			const off_t off = ra->off;
			size_t want = ra->want;
			struct stat st;
			long got = 0;
			pthread_mutex_unlock(&ra->lock);
			const double t0 = now_s();
			/* what the file holds now, in parallel segments like a read in the foreground */
			if (fstat(fd, &st) == 0 && off < st.st_size) {
				if ((size_t)(st.st_size - off) < want)
					want = (size_t)(st.st_size - off);
				got = read_segments(fd, ra->buf, off, want, want < READ_PAR_MIN ? 1 : read_threads());
			}
			pthread_mutex_lock(&ra->lock);
			ra->t_ra_read += now_s() - t0;
			ra->got = got;
			ra->have = 1;
			ra->busy = 0;
			pthread_cond_broadcast(&ra->cond);
		}
	}
	pthread_mutex_unlock(&ra->lock);
	return NULL;
}

/* ask for [off, off + want) of fd; needs the second buffer (databuf_new; ACM_READAHEAD=0: there is none) */
static void
readahead_start(struct databuf *db, int fd, off_t off, size_t want)
{
	struct databuf_priv *pv = priv_of(db);
	struct databuf_readahead *ra = &pv->ra;

	if (!ra->buf)
		return;
	pthread_mutex_lock(&ra->lock);
	if (!ra->started) {
		if (pthread_create(&ra->thread, NULL, readahead_main, ra) != 0) {
			pthread_mutex_unlock(&ra->lock);
			return;
		}
		ra->started = 1;
	}
	while (ra->busy)
		pthread_cond_wait(&ra->cond, &ra->lock);
	ra->fd = fd;
	ra->off = off;
	ra->want = want;
	ra->have = 0;
	ra->busy = 1;
	pthread_cond_broadcast(&ra->cond);
	pthread_mutex_unlock(&ra->lock);
}

/* the read-ahead result if it is exactly what is asked for now: swaps the buffers, returns the bytes; -2 = none */
static long
readahead_take(struct databuf *db, int fd, off_t off, size_t want)
{
	struct databuf_priv *pv = priv_of(db);
	struct databuf_readahead *ra = &pv->ra;
	long got = -2;

	if (!ra->started)
		return -2;
	const double t0 = now_s();
	pthread_mutex_lock(&ra->lock);
	while (ra->busy)
		pthread_cond_wait(&ra->cond, &ra->lock);
	ra->t_wait += now_s() - t0;
	if (ra->have && ra->fd == fd && ra->off == off && ra->want == want && ra->got >= 0) {
		unsigned char *t = db->h_data;
		db->h_data = ra->buf;
		ra->buf = t;
		got = ra->got;
		ra->n_taken++;
	} else if (ra->have) {
		ra->n_missed++;
	}
	ra->have = 0;
	pthread_mutex_unlock(&ra->lock);
	return got;
}

static void
readahead_stop(struct databuf_priv *pv)
{
	struct databuf_readahead *ra = &pv->ra;

	if (ra->started) {
		pthread_mutex_lock(&ra->lock);
		while (ra->busy)
			pthread_cond_wait(&ra->cond, &ra->lock);
		ra->quit = 1;
		pthread_cond_broadcast(&ra->cond);
		pthread_mutex_unlock(&ra->lock);
		pthread_join(ra->thread, NULL);
		ra->started = 0;
	}
	acm_host_free_pinned(ra->buf);
	ra->buf = NULL;
}

/*
 * reference databuf.c:327-407.  Returns what the reference returns, and -4 (with the reason in
 * acm_last_error()) when the read itself fails: the reference aborts there; "0 = end of file"
 * would silently truncate the scan.
 */
int
databuf_add_fd(struct databuf *db, int fd, int id, size_t *rd_bytes)
{
	size_t i, cur_chunks, tail;
	const size_t room = (db->max_chunks - db->chunks) * db->max_chunk_size;
	ssize_t got = -2;
	struct stat st;
	off_t pos = (off_t)-1;
	int regular;

	*rd_bytes = 0;
	if (db->chunks >= db->max_chunks)
		return -1;
	regular = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && (pos = lseek(fd, 0, SEEK_CUR)) != (off_t)-1;
	/* chunks are fixed size in this mode: chunk k starts at k * max_chunk_size */
	/* (priv is NULL for a caller-built, host-only struct: no read-ahead there) */
	if (regular && db->chunks == 0 && db->priv) {
		got = readahead_take(db, fd, pos, room);
		if (got >= 0 && lseek(fd, pos + (off_t)got, SEEK_SET) == (off_t)-1)
			got = -1;
	}
	if (got == -2) {
		const double t0 = now_s();
		got = databuf_read_fd(fd, db->h_data + db->chunks * db->max_chunk_size, room);
		if (db->priv) {
			priv_of(db)->ra.t_direct += now_s() - t0;
			priv_of(db)->ra.n_direct++;
		}
	}
	if (got < 0) {
		acm_set_error("databuf_add_fd: read failed: %s", strerror(errno));
		if (db->priv)
			priv_of(db)->status = ACM_ERR_IO;
		return -4;
	}
	if (got == 0)
		return 0;
	*rd_bytes = (size_t)got;
	/* a whole buffer in one go and more of the file behind it: fetch the next one meanwhile */
	if (regular && db->priv && db->chunks == 0 && (size_t)got == db->size && pos + (off_t)got < st.st_size)
		readahead_start(db, fd, pos + (off_t)got, db->size);

	cur_chunks = (size_t)got / db->max_chunk_size;
	for (i = db->chunks; i < db->chunks + cur_chunks; i++) {
		db->h_indices[i] = (int)(i * db->max_chunk_size);
		db->h_sizes[i] = (int)db->max_chunk_size;
		db->file_ids[i] = id;
	}
	db->chunks += cur_chunks;
	tail = (size_t)got % db->max_chunk_size;
	if (tail) {
		db->h_indices[db->chunks] = (int)(db->chunks * db->max_chunk_size);
		db->h_sizes[db->chunks] = (int)tail;
		memset(db->h_data + db->chunks * db->max_chunk_size + tail, 0, db->max_chunk_size - tail);
		db->file_ids[db->chunks] = id;
		db->chunks++;
	}
	db->bytes = db->chunks * db->max_chunk_size;
	if (db->chunks == db->max_chunks)
		return -1;
	if ((size_t)got == db->size)
		return -2;
	return (int)got;
}

/* reference databuf.c:413-481 */
int
databuf_add_fp(struct databuf *db, FILE *fp, int id, int aligned, size_t *rd_bytes, size_t *rd_lines)
{
	char *buf;
	size_t toread, len, adv;

	*rd_bytes = *rd_lines = 0;
	if (db->chunks >= db->max_chunks)
		return -1;
	if (db->bytes >= db->size)
		return -2;
	buf = (char *)db->h_data + db->bytes;
	toread = MIN(db->size - db->bytes, db->max_chunk_size);
	/* fgets needs room for its NUL: a line longer than the chunk is split, as in the reference */
	while (toread >= 2 && fgets(buf, (int)toread, fp) != NULL) {
		len = strnlen(buf, toread);
		*rd_bytes += len;
		if (len && buf[len - 1] == '\n')
			*rd_lines += 1;
		db->h_indices[db->chunks] = (int)db->bytes;
		db->h_sizes[db->chunks] = (int)len;
		db->file_ids[db->chunks] = id;
		db->chunks += 1;
		adv = aligned ? ROUNDUP16(len) : len;
		if (db->bytes + adv > db->size)
			adv = db->size - db->bytes;
		/* zero the padding (and fgets' NUL) so stale bytes cannot match */
		if (adv > len)
			memset(buf + len, 0, adv - len);
		db->bytes += adv;
		if (db->chunks >= db->max_chunks)
			return -1;
		if (db->bytes >= db->size)
			return -2;
		buf = (char *)db->h_data + db->bytes;
		toread = MIN(db->size - db->bytes, db->max_chunk_size);
	}
	if (toread < 2 && !feof(fp))
		return -2;
	return (int)(db->size - db->bytes);
}

/* reference databuf.c:488-528 */
int
databuf_add_chunk(struct databuf *db, char *chunk, size_t len, int id, char aligned)
{
	size_t adv;

	if (len > db->max_chunk_size)
		return -3;
	if (db->chunks >= db->max_chunks)
		return -1;
	if (db->bytes + len >= db->size)
		return -2;
	memcpy(db->h_data + db->bytes, chunk, len);
	db->h_indices[db->chunks] = (int)db->bytes;
	db->h_sizes[db->chunks] = (int)len;
	db->file_ids[db->chunks] = id;
	db->chunks += 1;
	adv = aligned ? ROUNDUP16(len) : len;
	if (db->bytes + adv > db->size)
		adv = db->size - db->bytes;
	if (adv > len)
		memset(db->h_data + db->bytes + len, 0, adv - len);
	db->bytes += adv;
	return (int)(db->size - db->bytes);
}

void
databuf_reset(struct databuf *db)
{
	db->chunks = 0;
	db->bytes = 0;
}

void
databuf_clear(struct databuf *db)
{
	struct databuf_priv *pv = priv_of(db);
	size_t nres = (size_t)db->max_results * db->max_chunks + 1;

	memset(db->h_data, 0, db->size);
	memset(db->h_indices, 0, db->max_chunks * sizeof(int));
	memset(db->h_sizes, 0, db->max_chunks * sizeof(int));
	memset(db->h_results, 0, nres * sizeof(int));
	memset(db->h_results2, 0, nres * sizeof(int));
	memset(db->h_results_comp, 0, db->results_comp_size * sizeof(int));
	memset(db->h_results2_comp, 0, db->results2_comp_size * sizeof(int));
	memset(db->file_ids, 0, db->max_chunks * sizeof(int));
	pv->carry_len = 0;
	pv->carry_real = 0;
	pv->carry_file = -1;
	pv->n_matches = pv->n_valid = 0;
	db->last_state = 0;
	databuf_reset(db);
}

void
databuf_copy_host_to_device(struct databuf *db, cl_command_queue queue)
{
	struct databuf_priv *pv = priv_of(db);

	(void)queue;
	if (db->bytes == 0)
		return;
	pv->status = acm_memcpy_h2d(pv->dev, db->d_data, db->h_data, db->bytes);
}

static int
chunk_of(const struct databuf *db, long off)
{
	size_t lo = 0, hi = db->chunks;

	/* last chunk whose start is <= off */
	while (hi - lo > 1) {
		size_t mid = (lo + hi) / 2;
		if ((long)db->h_indices[mid] <= off)
			lo = mid;
		else
			hi = mid;
	}
	return (int)lo;
}

/*
 * Chunks of one contiguous stretch of real bytes of one file get the same id: consecutive chunks
 * continue a run when the file id is the same and the earlier chunk is filled to where the later
 * one starts (no zero padding in between).
 */
static void
label_runs(const struct databuf *db, int *run_id)
{
	size_t c;
	int run = 0;

	for (c = 0; c < db->chunks; c++) {
		if (c && !(db->file_ids[c] == db->file_ids[c - 1] &&
		    db->h_indices[c - 1] + db->h_sizes[c - 1] == db->h_indices[c]))
			run++;
		run_id[c] = run;
	}
}

void
databuf_copy_device_to_host(struct databuf *db, cl_command_queue queue)
{
	struct databuf_priv *pv = priv_of(db);
	const size_t R = (size_t)db->max_results, C = db->chunks;
	const uint64_t n = pv->n_matches;
	const uint32_t *plen = (pv->file_semantics && pv->sym_size == 1 && pv->scanner_aut) ?
	    acm_automaton_pattern_lengths(pv->scanner_aut) : NULL;
	uint64_t i, kept = 0;
	int64_t got;

	(void)queue;
	/* bucket rows 1.. are only ever read below a chunk's count: clearing the counts is enough */
	memset(db->h_results, 0, db->max_chunks * sizeof(int));
	memset(db->h_results2, 0, db->max_chunks * sizeof(int));
	if (n + 2 > db->results_comp_size) {
		size_t nc = db->results_comp_size;
		int *a, *b;
		while (nc < n + 2)
			nc *= 2;
		a = realloc(db->h_results_comp, nc * sizeof(int));
		b = realloc(db->h_results2_comp, nc * sizeof(int));
		if (a)
			db->h_results_comp = a;
		if (b)
			db->h_results2_comp = b;
		if (!a || !b) {
			acm_set_error("databuf_copy_device_to_host: out of memory");
			pv->status = ACM_ERR_NOMEM;
			return;
		}
		db->results_comp_size = db->results2_comp_size = nc;
	}
	if (n > pv->h_cap) {
		free(pv->h_off);
		free(pv->h_pat);
		pv->h_off = malloc(n * sizeof(uint64_t));
		pv->h_pat = malloc(n * sizeof(uint32_t));
		pv->h_cap = (pv->h_off && pv->h_pat) ? n : 0;
		if (!pv->h_cap) {
			acm_set_error("databuf_copy_device_to_host: out of memory");
			pv->status = ACM_ERR_NOMEM;
			return;
		}
	}
	got = n ? acm_scan_fetch(pv->scanner, 0, pv->h_off, pv->h_pat, n) : 0;
	if (got < 0) {
		pv->status = (int)got;
		return;
	}
	if (plen && n)
		label_runs(db, pv->run_id);
	for (i = 0; i < n; i++) {
		const long off = (long)(pv->h_off[i] - DATABUF_CARRY_CAP / pv->sym_size);
		const int pat = (int)pv->h_pat[i];
		const int c = chunk_of(db, off * pv->sym_size);
		int k;

		if (plen) {
			/*
			 * Per-file semantics: every byte of the match is a real byte of ONE file.  The end
			 * must not lie in a chunk's zero padding; the start must lie in the same contiguous
			 * run of chunks -- or in the carried tail of the previous buffer when that tail is the
			 * same file, ran up to that buffer's end and this run starts the buffer.
			 */
			const long start = off - (long)plen[pat] + 1;
			if (off >= (long)db->h_indices[c] + db->h_sizes[c])
				continue;
			if (start >= 0) {
				if (pv->run_id[chunk_of(db, start)] != pv->run_id[c])
					continue;
			} else if (pv->run_id[c] != 0 || db->h_indices[0] != 0 || db->file_ids[0] != pv->filt_carry_file ||
			    (uint64_t)(-start) > pv->filt_carry_real) {
				continue;
			}
		}
		db->h_results_comp[kept + 1] = pat;
		db->h_results2_comp[kept + 1] = (int)off;
		kept++;
		/* reference bucket layout, ahomatch.cl:67-73: row 0 = counts, row k = k-th match */
		k = ++db->h_results[c];
		db->h_results2[c] = k;
		if ((size_t)k < R) {
			db->h_results[(size_t)k * C + c] = pat;
			db->h_results2[(size_t)k * C + c] = (int)off;
		}
	}
	db->h_results_comp[0] = db->h_results2_comp[0] = (int)kept;
	db->h_results_comp[kept + 1] = db->h_results2_comp[kept + 1] = (int)db->last_state;
	db->h_results[C * R] = (int)db->last_state;
	{
		int run = 0;
		for (i = 0; i < C; i++) {
			db->h_prefixsum[i] = run;
			run += db->h_results[i];
		}
	}
	pv->n_valid = kept;
	pv->fetched = 1;
}

int
databuf_process_results(struct databuf *db,
    int (*cb)(int file_idx, int patrn_idx, int chunk_idx, int offset, void *uarg), void *uarg)
{
	const int n = db->h_results_comp[0];
	struct databuf_priv *pv = priv_of(db);
	int i;

	if (cb) {
		for (i = 0; i < n; i++) {
			const int off = db->h_results2_comp[i + 1];
			const int c = chunk_of(db, (long)off * pv->sym_size);
			cb(db->file_ids[c], db->h_results_comp[i + 1], c, off + 1, uarg);
		}
	}
	return n;
}

void
databuf_free(struct databuf *db, int mapped, cl_command_queue queue)
{
	struct databuf_priv *pv;

	(void)mapped;
	(void)queue;
	if (!db)
		return;
	pv = priv_of(db);
	if (pv) {
		readahead_stop(pv);
		if (getenv("ACM_DATABUF_STATS"))
			fprintf(stderr, "databuf %p: %ld foreground reads %.3f s, %ld buffers read ahead (%.3f s reading, %.3f s waited for), %ld dropped\n",
			    (void *)db, pv->ra.n_direct, pv->ra.t_direct, pv->ra.n_taken, pv->ra.t_ra_read, pv->ra.t_wait, pv->ra.n_missed);
		pthread_mutex_destroy(&pv->ra.lock);
		pthread_cond_destroy(&pv->ra.cond);
		free(pv->run_id);
		if (pv->scanner)
			acm_scanner_free(pv->scanner);
		if (pv->dev) {
			acm_dev_free(pv->dev, pv->d_base);
			acm_dev_free(pv->dev, pv->d_carry_tmp);
			acm_dev_free(pv->dev, db->d_results);
			acm_dev_free(pv->dev, db->d_results2);
			acm_dev_free(pv->dev, db->d_prefixsum);
			acm_dev_free(pv->dev, db->d_results_comp);
			acm_dev_free(pv->dev, db->d_results2_comp);
		}
		free(pv->h_off);
		free(pv->h_pat);
		free(pv);
	}
	acm_host_free_pinned(db->h_data);
	free(db->h_indices); free(db->h_sizes); free(db->file_ids);
	free(db->h_results); free(db->h_results2); free(db->h_prefixsum);
	free(db->h_results_comp); free(db->h_results2_comp);
	free(db);
}

int
databuf_status(struct databuf *db)
{
	return priv_of(db)->status;
}

size_t
databuf_match_count(struct databuf *db)
{
	struct databuf_priv *pv = priv_of(db);

	return (size_t)(pv->fetched ? pv->n_valid : pv->n_matches);
}

void
databuf_set_file_semantics(struct databuf *db, int on)
{
	priv_of(db)->file_semantics = on != 0;
}

int
databuf_alloc_postpass(struct databuf *db)
{
	struct databuf_priv *pv = priv_of(db);
	const size_t nres = ((size_t)db->max_results * db->max_chunks + 1) * sizeof(int);
	const size_t ncomp = (db->size + 2) * sizeof(int);
	void *p;
	int rc;

	if (db->d_results)
		return ACM_OK;
#define PP(field, bytes)                                                            \
	if ((rc = acm_dev_alloc(pv->dev, (bytes), &p)) != ACM_OK)                     \
		return pv->status = rc;                                                     \
	db->field = p
	PP(d_results, nres);
	PP(d_results2, nres);
	PP(d_prefixsum, db->max_chunks * sizeof(int) + 16);
	PP(d_results_comp, ncomp);
	PP(d_results2_comp, ncomp);
#undef PP
	/* nothing on the device has filled these yet: never let a post-pass read raw memory */
	if ((rc = acm_dev_memset(pv->dev, db->d_results, 0, nres)) != ACM_OK ||
	    (rc = acm_dev_memset(pv->dev, db->d_results2, 0, nres)) != ACM_OK ||
	    (rc = acm_dev_memset(pv->dev, db->d_prefixsum, 0, db->max_chunks * sizeof(int) + 16)) != ACM_OK ||
	    (rc = acm_dev_memset(pv->dev, db->d_results_comp, 0, ncomp)) != ACM_OK ||
	    (rc = acm_dev_memset(pv->dev, db->d_results2_comp, 0, ncomp)) != ACM_OK)
		return pv->status = rc;
	return ACM_OK;
}
