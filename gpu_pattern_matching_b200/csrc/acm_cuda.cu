/*
 * acm_cuda.cu -- device objects, kernel launchers and the native C ABI of acm.h.
 *
 * Compiled for sm_100a only.  There is no CPU fallback anywhere in this file: every
 * entry point that needs a GPU returns ACM_ERR_NO_DEVICE / ACM_ERR_CUDA when there
 * is none.
 */
#include <cuda_runtime.h>

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "../../include/acm.h"
#include "acm_core.h"
#include "acm_tables.h"
#include "k1_scan.cuh"
#include "k234_post.cuh"
#include "k1_rd.cuh"

#define CUDA_TRY(expr)                                                              \
	do {                                                                            \
		cudaError_t e_ = (expr);                                                    \
		if (e_ != cudaSuccess) {                                                    \
			acm_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #expr,              \
			    cudaGetErrorString(e_));                                            \
			return ACM_ERR_CUDA;                                                    \
		}                                                                           \
	} while (0)

struct acm_device {
	int          ordinal;
	int          sm_count;
	cudaStream_t own_stream;
	cudaStream_t stream;        /* own_stream or an adopted one */
	cudaStream_t copy_stream;   /* H2D staging for acm_scan_host, side D2H */
	cudaEvent_t  side_ev;
	/* scanners with their own streams (acm_scan_params.own_stream): the scan stage of a step waits
	 * for the scan stage of the step launched before it on ANOTHER scanner, not for that step's
	 * post-pass -- k1_ev is that scanner's event, recorded behind its scan-stage kernels */
	cudaEvent_t  k1_ev;
	struct acm_scanner *k1_owner;
	/* scratch for the stand-alone scan / sort entry points */
	uint64_t    *tile_state;
	uint32_t    *tile_counter;
	uint32_t     tile_cap;
	uint32_t    *hist;
	uint64_t     hist_cap;
};

struct acm_automaton {
	struct acm_device *dev;
	AutDev    d;                /* device pointers */
	uint32_t  num_states, num_patterns, gram_count;
	int       alpha, max_len, min_len;
	size_t    bytes;
	uint32_t *h_pat_len;        /* host copies that outlive acsm_cleanup() */
	int32_t  *h_pat_iid;
	void     *allocs[32];
	int       n_allocs;
};

struct acm_scanner {
	struct acm_device    *dev;
	struct acm_automaton *aut;
	struct acm_scan_params p;
	uint64_t  max_bytes;
	uint32_t  shift, cap, max_buckets;
	uint32_t  cd_hot, cd_tab_bytes;                 /* CDFA, plain table: shared-memory layout (see k_scan_cdfa) */
	int       rd;                                   /* CDFA through the row-displaced table (k_scan_rd + k_rd_expand) */
	uint32_t  rd_warps, rd_tab_bytes, rd_log_cap, rd_regions;
	uint32_t *rd_loglen;
	uint64_t *buckets;
	uint32_t *counts, *offsets;
	uint8_t  *scratch;          /* one allocation, one memset per scan: flags | bucket_tiles | counts */
	uint64_t *bucket_tiles;     /* look-back tile states of the bucket-count scan */
	uint32_t  n_bucket_tiles;
	uint32_t *flags;            /* [0] overflow, [1] total, [2] final state, [3] scan tile counter, [5] output too small, [6,7] K1 work counters */
	uint64_t *tile_state;
	uint32_t  max_tiles;
	uint64_t *out;              /* sorted keys of the last scan */
	uint64_t  out_cap;
	uint64_t *tmp;              /* radix sort ping-pong (fallback only) */
	uint64_t  tmp_cap;
	uint32_t *hist;             /* radix histograms (fallback only) */
	uint64_t  hist_cap;
	uint32_t *h_flags;          /* pinned + mapped, 8 words: written by k_publish_flags at the end of a step */
	uint32_t *h_flags_dev;      /* the same memory as the device sees it */
	uint64_t  last_n;           /* matches of the last scan */
	int       densify;          /* a bucket overflowed: use smaller, deeper buckets from the next scan on */
	int       user_shape;       /* bucket shape was given by the caller: never change it */
	cudaEvent_t ev[4];
	/* acm_scan_host staging */
	uint8_t  *stage[2];
	uint64_t  stage_bytes;
	cudaEvent_t ev_copied[2], ev_free[2];
	uint64_t *trace;            /* ACM_TRACE=1: per-CTA timestamps of the last K1 launch */
	uint4    *vq;               /* sampled mode: verification queue, one region per scanning warp */
	uint32_t *vq_count;
	uint32_t  vq_cap;
	uint32_t *dq;               /* sampled mode: per-warp lists of dense chunks */
	uint32_t *dq_count;
	uint32_t  dq_cap;
	uint64_t *h_keys;           /* pinned bounce buffer for results */
	uint64_t  h_keys_cap;
	/* a scan queued by acm_scan_device_async and not yet finished */
	struct {
		int          active;
		cudaStream_t st;
		const void  *d_data;
		uint64_t     n, emit_lo, emit_hi, out_cap_at_launch;
		EmitCtx      E;
		uint32_t     nb, k3_blocks, launches;
		uint64_t    *push_dst;      /* optional: keys pushed to a gather region by the step itself */
		uint64_t     push_cap, push_add;
	} pend;
	cudaEvent_t ev_done;
	cudaStream_t own_stream;    /* acm_scan_params.own_stream: this scanner's scans run here, not on the device's stream */
	cudaEvent_t  ev_dep;        /* orders a scan on own_stream behind what the device's stream held at launch */
	cudaEvent_t  ev_k1;         /* own_stream: recorded behind the scan-stage kernels of a step (acm_device.k1_ev) */
	uint32_t     reserve_sms;   /* own_stream: SMs the persistent scan kernel leaves to the other scanner's post-pass */
};

/* ------------------------------------------------------------------------- */
/* device                                                                    */
/* ------------------------------------------------------------------------- */

extern "C" int
acm_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

/*
 * Launch as the programmatic dependent of the kernel before it in the stream (k1_scan.cuh:pdl_wait):
 * set-up and CTA residency overlap the predecessor's tail.  ACM_PDL=0: plain launches.
 */
static int g_pdl = -1;

template <typename... KArgs, typename... Args>
static cudaError_t
launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
	cudaLaunchConfig_t cfg;
	cudaLaunchAttribute at[1];

	if (g_pdl < 0)
		g_pdl = !(getenv("ACM_PDL") && !atoi(getenv("ACM_PDL")));
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = g_pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static int g_attr_done[64];

static int
set_kernel_attrs(int ordinal)
{
	if (ordinal < 64 && g_attr_done[ordinal])
		return ACM_OK;
	CUDA_TRY(cudaFuncSetAttribute(k_scan_sampled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    S4_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_sampled<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    S4_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_start2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    S2_SMEM_BYTES(0)));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_start2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    S2_SMEM_BYTES(1)));
	CUDA_TRY(cudaFuncSetAttribute(k_bucket_sort_compact, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    65536));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_cdfa<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    CD_SMEM_MAX));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_cdfa<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    CD_SMEM_MAX));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_rd, cudaFuncAttributeMaxDynamicSharedMemorySize, CD_SMEM_MAX));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_xd<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, XD_SMEM_MAX));
	CUDA_TRY(cudaFuncSetAttribute(k_scan_xd<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, XD_SMEM_MAX));
	CUDA_TRY(cudaFuncSetAttribute(k_dense_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, XD_SMEM_MAX));
	CUDA_TRY(cudaFuncSetAttribute(k_rd_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, RD_K3_SMEM));
	if (ordinal < 64)
		g_attr_done[ordinal] = 1;
	return ACM_OK;
}

extern "C" int
acm_device_open(int ordinal, struct acm_device **out)
{
	int n = acm_device_count();
	struct acm_device *d;
	cudaDeviceProp prop;
	int rc;

	*out = NULL;
	if (n <= 0) {
		acm_set_error("no CUDA device visible; this library has no CPU fallback");
		return ACM_ERR_NO_DEVICE;
	}
	if (ordinal < 0 || ordinal >= n) {
		acm_set_error("device ordinal %d out of range (0..%d)", ordinal, n - 1);
		return ACM_ERR_ARG;
	}
	CUDA_TRY(cudaSetDevice(ordinal));
	CUDA_TRY(cudaGetDeviceProperties(&prop, ordinal));
	if (prop.major < 10) {
		acm_set_error("device %d is sm_%d%d; this build carries sm_100a code only", ordinal,
		    prop.major, prop.minor);
		return ACM_ERR_NO_DEVICE;
	}
	d = (struct acm_device *)calloc(1, sizeof(*d));
	if (!d)
		return ACM_ERR_NOMEM;
	d->ordinal = ordinal;
	d->sm_count = prop.multiProcessorCount;
	CUDA_TRY(cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
	d->stream = d->own_stream;
	if ((rc = set_kernel_attrs(ordinal)) != ACM_OK) {
		free(d);
		return rc;
	}
	*out = d;
	return ACM_OK;
}

extern "C" void
acm_device_close(struct acm_device *d)
{
	if (!d)
		return;
	cudaSetDevice(d->ordinal);
	cudaStreamSynchronize(d->stream);
	cudaFree(d->tile_state);
	cudaFree(d->tile_counter);
	cudaFree(d->hist);
	if (d->side_ev)
		cudaEventDestroy(d->side_ev);
	cudaStreamDestroy(d->own_stream);
	cudaStreamDestroy(d->copy_stream);
	free(d);
}

extern "C" int acm_device_ordinal(struct acm_device *d) { return d->ordinal; }
extern "C" void *acm_device_stream(struct acm_device *d) { return (void *)d->stream; }

extern "C" int
acm_device_set_stream(struct acm_device *d, void *s)
{
	d->stream = s ? (cudaStream_t)s : d->own_stream;
	return ACM_OK;
}

extern "C" int
acm_device_sync(struct acm_device *d)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaStreamSynchronize(d->stream));
	return ACM_OK;
}

static struct acm_device *g_default_dev;

extern "C" struct acm_device *
acm_default_device(void)
{
	if (!g_default_dev) {
		const char *e = getenv("ACM_DEVICE");
		int ord = e ? atoi(e) : 0;
		if (!e && getenv("LOCAL_RANK") && acm_device_count() > 1)
			ord = atoi(getenv("LOCAL_RANK")) % acm_device_count();
		if (acm_device_open(ord, &g_default_dev) != ACM_OK)
			return NULL;
	}
	return g_default_dev;
}

extern "C" int
acm_dev_alloc(struct acm_device *d, size_t bytes, void **p)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
	return ACM_OK;
}

extern "C" void
acm_dev_free(struct acm_device *d, void *p)
{
	if (!p)
		return;
	cudaSetDevice(d->ordinal);
	cudaFree(p);
}

/*
 * NUMA node the GPU hangs off (/sys/bus/pci/devices/<bus id>/numa_node), -1 if unknown.  On a
 * two-socket 8-GPU box a pinned buffer on the other socket makes every H2D copy cross the
 * socket interconnect; eight ranks that all allocate on the node their process happens to run
 * on share that one node's memory controllers and links.
 */
static int
numa_node_of_device(int ordinal)
{
	char bus[32], path[128];
	int node = -1;
	FILE *f;

	if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), ordinal) != cudaSuccess) {
		cudaGetLastError();
		return -1;
	}
	for (char *c = bus; *c; ++c)
		if (*c >= 'A' && *c <= 'Z')
			*c = (char)(*c - 'A' + 'a');
	snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
	f = fopen(path, "r");
	if (!f)
		return -1;
	if (fscanf(f, "%d", &node) != 1)
		node = -1;
	fclose(f);
	return node;
}

/* pinned host memory, preferably on the NUMA node of GPU `ordinal` (ACM_NUMA=0: wherever the
 * calling thread's policy puts it).  The preference is a hint: where the node is not allowed
 * (cpuset) or unknown, the call is a plain cudaHostAlloc. */
static int
alloc_pinned_near(int ordinal, size_t bytes, void **p)
{
	const char *env = getenv("ACM_NUMA");
	const int node = (env && atoi(env) == 0) ? -1 : numa_node_of_device(ordinal);
	int preferred = 0;

	if (node >= 0 && node < 1024) {
		unsigned long mask[1024 / (8 * sizeof(unsigned long))];
		memset(mask, 0, sizeof(mask));
		mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
		preferred = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024ul + 1) == 0;
	}
	const cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocDefault);
	if (preferred)
		syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, NULL, 0ul);
	if (e != cudaSuccess) {
		acm_set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
		cudaGetLastError();
		return ACM_ERR_CUDA;
	}
	return ACM_OK;
}

extern "C" int
acm_host_alloc_pinned(size_t bytes, void **p)
{
	int ordinal = 0;

	if (cudaGetDevice(&ordinal) != cudaSuccess) {
		cudaGetLastError();
		ordinal = 0;
	}
	return alloc_pinned_near(ordinal, bytes, p);
}

extern "C" int
acm_host_alloc_pinned_near(struct acm_device *d, size_t bytes, void **p)
{
	return alloc_pinned_near(d ? d->ordinal : 0, bytes, p);
}

extern "C" void
acm_host_free_pinned(void *p)
{
	if (p)
		cudaFreeHost(p);
}

extern "C" int
acm_memcpy_h2d(struct acm_device *d, void *dst, const void *src, size_t bytes)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d->stream));
	return ACM_OK;
}

extern "C" int
acm_memcpy_d2h(struct acm_device *d, void *dst, const void *src, size_t bytes)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d->stream));
	return ACM_OK;
}

/* D2H on the side (copy) stream, ordered after everything queued so far on the main stream */
extern "C" int
acm_memcpy_d2h_side(struct acm_device *d, void *dst, const void *src, size_t bytes)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	if (!d->side_ev)
		CUDA_TRY(cudaEventCreateWithFlags(&d->side_ev, cudaEventDisableTiming));
	CUDA_TRY(cudaEventRecord(d->side_ev, d->stream));
	CUDA_TRY(cudaStreamWaitEvent(d->copy_stream, d->side_ev, 0));
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d->copy_stream));
	return ACM_OK;
}

/*
 * D2H of nseg device segments, packed back to back into h_dst, on the side stream with NO
 * ordering against the main stream: for data the caller already knows to be complete (gather
 * regions whose writers have finished) while later steps are queued on the main stream.
 * Waits for the copies.
 */
extern "C" int
acm_memcpy_d2h_segments(struct acm_device *d, void *h_dst, const void *const *d_src, const uint64_t *bytes,
    uint32_t nseg)
{
	uint8_t *dst = (uint8_t *)h_dst;
	int queued = 0;

	CUDA_TRY(cudaSetDevice(d->ordinal));
	for (uint32_t i = 0; i < nseg; i++) {
		if (!bytes[i])
			continue;
		CUDA_TRY(cudaMemcpyAsync(dst, d_src[i], bytes[i], cudaMemcpyDeviceToHost, d->copy_stream));
		dst += bytes[i];
		queued = 1;
	}
	if (queued)
		CUDA_TRY(cudaStreamSynchronize(d->copy_stream));
	return ACM_OK;
}

/* the same without the wait: the copies are complete after the next acm_side_sync() */
extern "C" int
acm_memcpy_d2h_segments_async(struct acm_device *d, void *h_dst, const void *const *d_src, const uint64_t *bytes,
    uint32_t nseg)
{
	uint8_t *dst = (uint8_t *)h_dst;

	CUDA_TRY(cudaSetDevice(d->ordinal));
	for (uint32_t i = 0; i < nseg; i++) {
		if (!bytes[i])
			continue;
		CUDA_TRY(cudaMemcpyAsync(dst, d_src[i], bytes[i], cudaMemcpyDeviceToHost, d->copy_stream));
		dst += bytes[i];
	}
	return ACM_OK;
}

extern "C" int
acm_side_sync(struct acm_device *d)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaStreamSynchronize(d->copy_stream));
	return ACM_OK;
}

extern "C" int
acm_dev_memset(struct acm_device *d, void *dst, int value, size_t bytes)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMemsetAsync(dst, value, bytes, d->stream));
	return ACM_OK;
}

extern "C" int
acm_memcpy_d2d(struct acm_device *d, void *dst, const void *src, size_t bytes)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, d->stream));
	return ACM_OK;
}

/* ------------------------------------------------------------------------- */
/* automaton                                                                 */
/* ------------------------------------------------------------------------- */

static int
upload(struct acm_automaton *a, const void *src, size_t bytes, const void **dst)
{
	void *p = NULL;
	size_t padded = (bytes + 255) & ~(size_t)255;

	*dst = NULL;
	if (a->n_allocs >= 32)
		return ACM_ERR_LIMIT;
	CUDA_TRY(cudaMalloc(&p, padded ? padded : 256));
	a->allocs[a->n_allocs++] = p;
	if (bytes)
		CUDA_TRY(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, a->dev->stream));
	a->bytes += padded;
	*dst = p;
	return ACM_OK;
}

extern "C" void
acm_automaton_free(struct acm_automaton *a)
{
	if (!a)
		return;
	cudaSetDevice(a->dev->ordinal);
	cudaStreamSynchronize(a->dev->stream);
	for (int i = 0; i < a->n_allocs; i++)
		cudaFree(a->allocs[i]);
	free(a->h_pat_len);
	free(a->h_pat_iid);
	free(a);
}

extern "C" int
acm_automaton_upload(struct acm_device *dev, const struct acm_tables *t, struct acm_automaton **out)
{
	struct acm_automaton *a;
	int rc = ACM_OK;

	*out = NULL;
	if (!dev) {
		acm_set_error("automaton_upload: no device");
		return ACM_ERR_NO_DEVICE;
	}
	if (!t || !t->T) {
		acm_set_error("automaton_upload: automaton not compiled");
		return ACM_ERR_STATE;
	}
	CUDA_TRY(cudaSetDevice(dev->ordinal));
	a = (struct acm_automaton *)calloc(1, sizeof(*a));
	if (!a)
		return ACM_ERR_NOMEM;
	a->dev = dev;
	a->num_states = t->num_states;
	a->num_patterns = t->num_patterns;
	a->alpha = t->alpha;
	a->max_len = t->max_pattern_len;
	a->min_len = t->min_pattern_len;
	a->gram_count = t->gram_count;
	a->h_pat_len = (uint32_t *)malloc((size_t)(t->num_patterns + 1) * 4);
	a->h_pat_iid = (int32_t *)malloc((size_t)(t->num_patterns + 1) * 4);
	if (!a->h_pat_len || !a->h_pat_iid) {
		acm_automaton_free(a);
		return ACM_ERR_NOMEM;
	}
	memcpy(a->h_pat_len, t->pat_len, (size_t)t->num_patterns * 4);
	memcpy(a->h_pat_iid, t->pat_iid, (size_t)t->num_patterns * 4);

#define UP(field, src, bytes)                                                       \
	if (rc == ACM_OK)                                                               \
		rc = upload(a, (src), (bytes), (const void **)&a->d.field)
	UP(T, t->T, (size_t)t->num_states * t->alpha * 4);
	UP(level_start, t->level_start, (size_t)(t->max_depth + 2) * 4);
	UP(own_begin, t->own_begin, (size_t)(t->num_states + 1) * 4);
	UP(own_pat, t->own_pat, (size_t)t->own_total * 4);
	UP(olink, t->olink, (size_t)t->num_states * 4);
	if (t->b2s)
		UP(b2s, t->b2s, 65536 / 8);
	if (t->b3)
		UP(b3, t->b3, ACM_B3_WORDS * 4);
	if (t->f1) {
		UP(f1, t->f1, (1u << ACM_F1_BITS_LOG2) / 8);
		UP(f2, t->f2, ACM_F2_WORDS * 4);
		UP(grams, t->grams, (size_t)t->gram_slots * sizeof(struct acm_gram_slot));
		UP(cand, t->cand, ((size_t)t->cand_count + ACM_CAND_PAD) * sizeof(struct acm_cand));
		UP(pat_blob, t->pat_blob, t->pat_blob_bytes);
		UP(pat_off, t->pat_off, (size_t)t->num_patterns * 4);
		UP(pat_win, t->pat_win, (size_t)t->num_patterns * 8);
		UP(pat_len, t->pat_len, (size_t)t->num_patterns * 4);
		a->d.gram_mask = t->gram_slots - 1;
		uint32_t lg = 0;
		while ((1u << lg) < t->gram_slots)
			lg++;
		a->d.gram_shift = 32 - lg;
	}
	if (t->cd_tab) {
		UP(cd_tab, t->cd_tab, (size_t)t->num_states * t->cd_classes * 2);
		UP(cd_cls, t->cd_cls, 256);
		UP(cd_flat_begin, t->cd_flat_begin, (size_t)(t->num_states + 1) * 4);
		UP(cd_flat_pat, t->cd_flat_pat, (size_t)(t->cd_flat_total + 1) * 4);
		UP(cd_flat4, t->cd_flat4, (size_t)t->num_states * 16);
		a->d.cd_classes = t->cd_classes;
		a->d.cd_range_lo = t->cd_range_lo;
		a->d.cd_thr4 = t->cd_thr4;
		if (t->rd_tab) {
			UP(rd_tab, t->rd_tab, (size_t)t->rd_len * 4);
			UP(rd_flat4, t->rd_flat4, (size_t)t->rd_len * 16);
			a->d.rd_len = t->rd_len;
		}
	}
	if (t->xd_tab) {
		UP(xd_tab, t->xd_tab, (size_t)t->xd_len * 4);
		UP(xd_sid, t->xd_sid, (size_t)t->xd_len * 4);
		a->d.xd_len = t->xd_len;
		a->d.xd_d1_end = t->xd_d1_end;
		a->d.xd_sym_bits = t->xd_sym_bits;
	}
#undef UP
	a->d.num_states = t->num_states;
	a->d.alpha = t->alpha;
	a->d.max_len = t->max_pattern_len;
	a->d.sample_stride = t->sample_stride;
	a->d.max_win = t->max_win;
	a->d.split_len = t->b2s ? (uint32_t)t->split_len : 0u;
	if (rc == ACM_OK && cudaStreamSynchronize(dev->stream) != cudaSuccess) {
		acm_set_error("automaton_upload: %s", cudaGetErrorString(cudaGetLastError()));
		rc = ACM_ERR_CUDA;
	}
	if (rc != ACM_OK) {
		acm_automaton_free(a);
		return rc;
	}
	*out = a;
	return ACM_OK;
}

extern "C" uint32_t acm_automaton_states(const struct acm_automaton *a) { return a->num_states; }
extern "C" const uint32_t *acm_automaton_pattern_lengths(const struct acm_automaton *a) { return a->h_pat_len; }
extern "C" uint32_t acm_automaton_patterns(const struct acm_automaton *a) { return a->num_patterns; }
extern "C" int acm_automaton_max_pattern_len(const struct acm_automaton *a) { return a->max_len; }
extern "C" int acm_automaton_min_pattern_len(const struct acm_automaton *a) { return a->min_len; }
extern "C" int acm_automaton_alphabet(const struct acm_automaton *a) { return a->alpha; }
extern "C" size_t acm_automaton_device_bytes(const struct acm_automaton *a) { return a->bytes; }
extern "C" uint32_t acm_automaton_gram_count(const struct acm_automaton *a) { return a->gram_count; }
extern "C" int acm_automaton_sample_stride(const struct acm_automaton *a) { return a->d.sample_stride; }
extern "C" int acm_automaton_split_len(const struct acm_automaton *a) { return (int)a->d.split_len; }

extern "C" int
acm_automaton_default_mode(const struct acm_automaton *a)
{
	if (a->alpha != 256)
		return ACM_MODE_DFA;
	if (a->d.f1 && a->min_len >= 7)
		return ACM_MODE_SAMPLED4;
	/* cdfa chunks (>= 4 x (Lmax - 1) bytes) must fit the 18-bit offset of a hit */
	if (a->d.cd_tab && a->max_len <= (1 << (32 - ACM_CD_STATE_BITS - 2)))
		return ACM_MODE_CDFA;
	/* mixed set: sampled filter over the long patterns + start filter over the few short ones */
	if (a->d.f1 && a->d.split_len)
		return ACM_MODE_SAMPLED4;
	return ACM_MODE_START2;
}

/* columns of the class-compressed table (ACM_MODE_CDFA), 0 when the automaton has none */
extern "C" int
acm_automaton_cdfa_classes(const struct acm_automaton *a)
{
	return a->d.cd_tab ? (int)a->d.cd_classes : 0;
}

/* ------------------------------------------------------------------------- */
/* stand-alone post-pass primitives                                          */
/* ------------------------------------------------------------------------- */

static int
launch_exclusive_scan(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n,
    uint64_t *tile_state, uint32_t *tile_counter, uint32_t *total, int pre_zeroed = 0)
{
	const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;

	if (n == 0) {
		if (total)
			CUDA_TRY(cudaMemsetAsync(total, 0, 4, st));
		return ACM_OK;
	}
	if (!pre_zeroed) {
		CUDA_TRY(cudaMemsetAsync(tile_state, 0, (size_t)tiles * 8, st));
		CUDA_TRY(cudaMemsetAsync(tile_counter, 0, 4, st));
	}
	CUDA_TRY(launch_dep(k_scan_lookback, dim3(tiles), dim3(SCAN_THREADS), 0, st, in, out, n, tile_state, tile_counter,
	    total));
	return ACM_OK;
}

static int
dev_scan_scratch(struct acm_device *d, uint32_t tiles)
{
	if (tiles > d->tile_cap) {
		cudaFree(d->tile_state);
		d->tile_state = NULL;
		d->tile_cap = 0;
		CUDA_TRY(cudaMalloc(&d->tile_state, (size_t)tiles * 8));
		d->tile_cap = tiles;
	}
	if (!d->tile_counter)
		CUDA_TRY(cudaMalloc(&d->tile_counter, 16));
	return ACM_OK;
}

extern "C" int
acm_exclusive_scan_u32(struct acm_device *d, const uint32_t *in, uint32_t *out, uint32_t n,
    uint32_t *total)
{
	int rc;

	CUDA_TRY(cudaSetDevice(d->ordinal));
	if ((rc = dev_scan_scratch(d, (n + SCAN_TILE - 1) / SCAN_TILE + 1)) != ACM_OK)
		return rc;
	return launch_exclusive_scan(d->stream, in, out, n, d->tile_state, d->tile_counter, total);
}

extern "C" int
acm_compact_columns_i32(struct acm_device *d, int32_t *dst, const int32_t *src, const int32_t *prefix,
    int32_t len, int32_t max_results, int64_t dst_cap)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	if (len <= 0)
		return ACM_OK;
	k_compact_columns<<<(len + 255) / 256, 256, 0, d->stream>>>(dst, src, prefix, len, max_results, dst_cap);
	CUDA_TRY(cudaGetLastError());
	return ACM_OK;
}

/* sorts keys[0..n) on bits [b0, b1); result ends in `keys`; hist has 256 * nblocks + scratch */
static int
radix_sort_impl(cudaStream_t st, uint64_t *keys, uint64_t *tmp, uint64_t n, int b0, int b1, int desc,
    uint32_t *hist, uint64_t *tile_state, uint32_t *tile_counter, uint32_t *launches)
{
	const uint32_t nblocks = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
	uint64_t *src = keys, *dst = tmp;
	int rc;

	if (n <= 1 || b1 <= b0)
		return ACM_OK;
	for (int shift = b0; shift < b1; shift += 8) {
		const uint64_t flip = desc ? ~0ull : 0ull;
		k_radix_hist<<<nblocks, RS_THREADS, 0, st>>>(src, n, shift, flip, hist, nblocks);
		CUDA_TRY(cudaGetLastError());
		rc = launch_exclusive_scan(st, hist, hist, 256u * nblocks, tile_state, tile_counter, NULL);
		if (rc != ACM_OK)
			return rc;
		k_radix_scatter<<<nblocks, RS_THREADS, 0, st>>>(src, dst, n, shift, flip, hist, nblocks);
		CUDA_TRY(cudaGetLastError());
		if (launches)
			*launches += 3;
		uint64_t *t = src;
		src = dst;
		dst = t;
	}
	if (src != keys)
		CUDA_TRY(cudaMemcpyAsync(keys, src, n * 8, cudaMemcpyDeviceToDevice, st));
	return ACM_OK;
}

static int
dev_sort_scratch(struct acm_device *d, uint64_t n)
{
	const uint64_t nblocks = (n + RS_TILE - 1) / RS_TILE;
	const uint64_t need = 256 * nblocks + 256;
	int rc;

	if (need > d->hist_cap) {
		cudaFree(d->hist);
		d->hist = NULL;
		d->hist_cap = 0;
		CUDA_TRY(cudaMalloc(&d->hist, need * 4));
		d->hist_cap = need;
	}
	if ((rc = dev_scan_scratch(d, (uint32_t)((need + SCAN_TILE - 1) / SCAN_TILE + 1))) != ACM_OK)
		return rc;
	return ACM_OK;
}

extern "C" int
acm_radix_sort_u64(struct acm_device *d, uint64_t *keys, uint64_t *tmp, uint64_t n, int b0, int b1,
    int desc)
{
	int rc;

	if (b0 < 0 || b1 > 64 || n >= (1ull << 32)) {
		acm_set_error("radix_sort: bad arguments");
		return ACM_ERR_ARG;
	}
	CUDA_TRY(cudaSetDevice(d->ordinal));
	if ((rc = dev_sort_scratch(d, n)) != ACM_OK)
		return rc;
	return radix_sort_impl(d->stream, keys, tmp, n, b0, b1, desc, d->hist, d->tile_state,
	    d->tile_counter, NULL);
}

extern "C" int
acm_sort_pairs_u32(struct acm_device *d, uint32_t *dk, uint32_t *dv, const uint32_t *sk,
    const uint32_t *sv, uint32_t n, int desc)
{
	uint64_t *buf = NULL;
	int rc;

	if (n == 0)
		return ACM_OK;
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMalloc(&buf, (size_t)n * 16));
	k_pack_pairs<<<(n + 255) / 256, 256, 0, d->stream>>>(sk, sv, buf, n);
	rc = acm_radix_sort_u64(d, buf, buf + n, n, 0, 64, desc);
	if (rc == ACM_OK) {
		k_unpack_pairs<<<(n + 255) / 256, 256, 0, d->stream>>>(buf, dk, dv, n);
		if (cudaStreamSynchronize(d->stream) != cudaSuccess) {
			acm_set_error("sort_pairs: %s", cudaGetErrorString(cudaGetLastError()));
			rc = ACM_ERR_CUDA;
		}
	}
	cudaFree(buf);
	return rc;
}

/* ------------------------------------------------------------------------- */
/* scanner                                                                   */
/* ------------------------------------------------------------------------- */

static uint32_t
next_pow2(uint32_t v)
{
	uint32_t p = 1;
	while (p < v)
		p <<= 1;
	return p;
}

extern "C" void
acm_scanner_free(struct acm_scanner *s)
{
	if (!s)
		return;
	cudaSetDevice(s->dev->ordinal);
	cudaStreamSynchronize(s->dev->stream);
	cudaStreamSynchronize(s->dev->copy_stream);
	if (s->own_stream) {
		cudaStreamSynchronize(s->own_stream);
		cudaStreamDestroy(s->own_stream);
	}
	if (s->ev_dep)
		cudaEventDestroy(s->ev_dep);
	if (s->ev_k1) {
		if (s->dev->k1_owner == s) {
			s->dev->k1_owner = NULL;
			s->dev->k1_ev = NULL;
		}
		cudaEventDestroy(s->ev_k1);
	}
	cudaFree(s->buckets); cudaFree(s->scratch); cudaFree(s->offsets);
	cudaFree(s->tile_state); cudaFree(s->out); cudaFree(s->tmp); cudaFree(s->hist);
	cudaFree(s->stage[0]); cudaFree(s->stage[1]); cudaFree(s->trace);
	cudaFree(s->vq); cudaFree(s->vq_count); cudaFree(s->dq); cudaFree(s->dq_count);
	if (s->h_flags)
		cudaFreeHost(s->h_flags);
	if (s->h_keys)
		cudaFreeHost(s->h_keys);
	for (int i = 0; i < 4; i++)
		if (s->ev[i])
			cudaEventDestroy(s->ev[i]);
	if (s->ev_done)
		cudaEventDestroy(s->ev_done);
	for (int i = 0; i < 2; i++) {
		if (s->ev_copied[i])
			cudaEventDestroy(s->ev_copied[i]);
		if (s->ev_free[i])
			cudaEventDestroy(s->ev_free[i]);
	}
	free(s);
}

/* (re)allocates buckets, scratch (flags | scan tile states | counts | log lengths)
 * and offsets for s->shift / s->cap */
static int
scanner_alloc_buckets(struct acm_scanner *s)
{
	cudaFree(s->buckets);
	cudaFree(s->scratch);
	cudaFree(s->offsets);
	s->buckets = NULL;
	s->scratch = NULL;
	s->offsets = NULL;
	/* CDFA cuts chunks on absolute multiples of 2^shift: one more partial bucket */
	s->max_buckets = (uint32_t)((s->max_bytes + (1ull << s->shift) - 1) >> s->shift) + 2;
	s->n_bucket_tiles = (s->max_buckets + SCAN_TILE - 1) / SCAN_TILE + 1;
	/* k_scan_rd: a region = 32 chunks; its log takes the 32 chunk rows (cap hits each) */
	s->rd_regions = s->rd ? s->max_buckets / 32 + 2 : 0;
	s->rd_log_cap = 32 * s->cap;
	/* test hook: a tiny log sends every dense region down the exact two-pass path */
	if (s->rd && getenv("ACM_RD_LOG_CAP") && atoi(getenv("ACM_RD_LOG_CAP")) > 32 * RD_CHECK &&
	    (uint32_t)atoi(getenv("ACM_RD_LOG_CAP")) < s->rd_log_cap)
		s->rd_log_cap = (uint32_t)atoi(getenv("ACM_RD_LOG_CAP")) / RD_LOG_ALIGN * RD_LOG_ALIGN;
	/* plain CDFA rows hold 4-byte hits; k_scan_rd logs 8-byte hits; everything else 8-byte record keys */
	const size_t slot = s->p.mode == ACM_MODE_CDFA && !s->rd ? 4 : 8;
	const size_t rows = s->rd ? (size_t)s->rd_regions * 32 : s->max_buckets;
	const size_t scratch_bytes = 64 + (size_t)s->n_bucket_tiles * 8 + (size_t)s->max_buckets * 4 +
	    (size_t)s->rd_regions * 4;
	if (cudaMalloc((void **)&s->buckets, rows * s->cap * slot) != cudaSuccess ||
	    cudaMalloc((void **)&s->scratch, scratch_bytes) != cudaSuccess ||
	    cudaMalloc((void **)&s->offsets, (size_t)s->max_buckets * 4) != cudaSuccess) {
		acm_set_error("scanner: cannot allocate %zu bytes of result buckets: %s",
		    rows * s->cap * slot, cudaGetErrorString(cudaGetLastError()));
		return ACM_ERR_CUDA;
	}
	s->flags = (uint32_t *)s->scratch;
	s->bucket_tiles = (uint64_t *)(s->scratch + 64);
	s->counts = (uint32_t *)(s->scratch + 64 + (size_t)s->n_bucket_tiles * 8);
	s->rd_loglen = s->counts + s->max_buckets;
	return ACM_OK;
}

extern "C" int
acm_scanner_create(struct acm_device *dev, struct acm_automaton *aut, uint64_t max_bytes,
    const struct acm_scan_params *params, struct acm_scanner **out)
{
	struct acm_scanner *s;
	int mode, rc;

	*out = NULL;
	if (!dev || !aut) {
		acm_set_error("scanner_create: device and automaton required");
		return ACM_ERR_ARG;
	}
	if (aut->dev != dev && aut->dev->ordinal != dev->ordinal) {
		acm_set_error("scanner_create: the automaton lives on device %d, the scanner is asked for device %d",
		    aut->dev->ordinal, dev->ordinal);
		return ACM_ERR_ARG;
	}
	if (max_bytes == 0 || max_bytes > (1ull << 40)) {
		acm_set_error("scanner_create: max_bytes must be in 1 .. 2^40");
		return ACM_ERR_LIMIT;
	}
	CUDA_TRY(cudaSetDevice(dev->ordinal));
	s = (struct acm_scanner *)calloc(1, sizeof(*s));
	if (!s)
		return ACM_ERR_NOMEM;
	s->dev = dev;
	s->aut = aut;
	if (params)
		s->p = *params;
	mode = s->p.mode ? s->p.mode : acm_automaton_default_mode(aut);
	if (aut->alpha != 256 && mode != ACM_MODE_DFA) {
		acm_set_error("scanner_create: ushort automata run in DFA mode only");
		free(s);
		return ACM_ERR_ARG;
	}
	if (mode == ACM_MODE_SAMPLED4 && (!aut->d.f1 || (aut->min_len < 7 && !aut->d.split_len))) {
		acm_set_error("scanner_create: sampled mode needs every pattern >= 7 bytes (shortest is %d)",
		    aut->min_len);
		free(s);
		return ACM_ERR_ARG;
	}
	if (mode == ACM_MODE_CDFA && !aut->d.cd_tab) {
		acm_set_error("scanner_create: this automaton has no class-compressed table "
		    "(needs <= %u states, <= %d distinct pattern bytes, <= 4 patterns ending in one state)",
		    ACM_CD_MAX_STATES, ACM_CD_MAX_CLASSES - 1);
		free(s);
		return ACM_ERR_ARG;
	}
	s->p.mode = mode;
	s->max_bytes = max_bytes;
	s->user_shape = s->p.bucket_shift || s->p.bucket_cap;
	if (mode == ACM_MODE_CDFA) {
		/*
		 * A bucket is one thread's chunk: 256 bytes unless the halo (Lmax - 1) asks for more
		 * (chunk >= 4 x halo keeps the cold-start overhead <= 25 %), room for one hit per
		 * 4 bytes before the exact two-pass path takes over.
		 */
		const uint32_t halo = aut->max_len > 0 ? (uint32_t)aut->max_len - 1 : 0;
		const char *plain = getenv("ACM_CD_PLAIN");
		uint32_t sh_rd = 8;
		while ((1u << sh_rd) < 4 * halo)
			sh_rd++;
		/* the row-displaced table in shared memory with at least 8 warps beside it: k_scan_rd */
		if (aut->d.rd_tab && !(plain && atoi(plain)) && sh_rd <= RD_MAX_SHIFT) {
			const uint32_t tab_bytes = (aut->d.rd_len * 4 + 15) & ~15u;
			const uint32_t tabpad = (tab_bytes + 1023) & ~1023u;
			uint32_t w = tabpad + 64 < CD_SMEM_MAX ? (CD_SMEM_MAX - 64 - tabpad) / RD_WARP_SMEM : 0;
			const char *we = getenv("ACM_RD_WARPS");
			if (w > 32)
				w = 32;
			if (we && atoi(we) > 0 && (uint32_t)atoi(we) < w)
				w = (uint32_t)atoi(we);
			if (w >= 8 || (we && w >= 1)) {
				s->rd = 1;
				s->rd_warps = w;
				s->rd_tab_bytes = tab_bytes;
			}
		}
		uint32_t sh = s->p.bucket_shift ? (uint32_t)s->p.bucket_shift : 8u;
		if (s->rd)
			sh = sh_rd;             /* the chunk size is the kernel's: a caller's bucket shape is a hint */
		while (sh < 30 && (1u << sh) < 4 * halo)
			sh++;
		if (sh > 32 - ACM_CD_STATE_BITS) {
			acm_set_error("scanner_create: cdfa chunks are at most 2^%d bytes (a hit is offset-in-chunk | state in "
			    "32 bits); patterns of %d bytes need more", 32 - ACM_CD_STATE_BITS, aut->max_len);
			free(s);
			return ACM_ERR_ARG;
		}
		s->p.bucket_shift = (int)sh;
		if (!s->p.bucket_cap)
			s->p.bucket_cap = (int)((1u << sh) / 4 > 8192 ? 8192 : (1u << sh) / 4);
		if (s->rd && (uint32_t)s->p.bucket_cap > (1u << sh))
			s->p.bucket_cap = (int)(1u << sh);  /* one hit per byte is all a chunk can log */
		const uint32_t lut_bytes = aut->d.cd_range_lo >= 0 ? 0 : CD_LUT_WORDS * 4;
		if (!s->rd) {
			const uint32_t row = aut->d.cd_classes * 2;
			uint32_t budget = CD_SMEM_MAX - 16 - lut_bytes;
			const char *kb = getenv("ACM_CD_HOT_KB");
			if (kb && atoi(kb) >= 0 && (uint32_t)atoi(kb) * 1024 < budget)
				budget = (uint32_t)atoi(kb) * 1024;
			else if (!kb && budget > 96 * 1024)
				budget = 96 * 1024;     /* measured: the rest is worth more as L1 for the cold rows */
			s->cd_hot = budget / row;
			if (s->cd_hot > aut->num_states)
				s->cd_hot = aut->num_states;
			s->cd_tab_bytes = (s->cd_hot * row + 15) & ~15u;
			if (s->cd_tab_bytes > budget) {
				s->cd_hot = s->cd_hot ? s->cd_hot - 1 : 0;
				s->cd_tab_bytes = (s->cd_hot * row + 15) & ~15u;
			}
		}
	}
	/* sparse matches (signature sets): 128 KiB buckets of up to 1024 records keep the
	 * post-passes at ~8k buckets per GiB; dense output is handled by the adaptive re-shape */
	s->shift = s->p.bucket_shift ? (uint32_t)s->p.bucket_shift : (mode == ACM_MODE_SAMPLED4 ? 17u : 15u);
	if (s->shift < 8 || s->shift > 30) {
		acm_set_error("scanner_create: bucket_shift must be in 8..30");
		free(s);
		return ACM_ERR_ARG;
	}
	s->cap = s->p.bucket_cap ? (uint32_t)s->p.bucket_cap : 1024u;
	if (s->cap < 32)
		s->cap = 32;
	if (s->cap > 8192)
		s->cap = 8192;
	s->cap = next_pow2(s->cap);
	if ((rc = scanner_alloc_buckets(s)) != ACM_OK) {
		acm_scanner_free(s);
		return rc;
	}
	s->max_tiles = (s->max_buckets + SCAN_TILE - 1) / SCAN_TILE + 1;

#define SALLOC(ptr, bytes)                                                          \
	do {                                                                            \
		cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                      \
		if (e_ != cudaSuccess) {                                                    \
			acm_set_error("scanner_create: cudaMalloc(%zu): %s", (size_t)(bytes),  \
			    cudaGetErrorString(e_));                                            \
			acm_scanner_free(s);                                                    \
			return ACM_ERR_CUDA;                                                    \
		}                                                                           \
	} while (0)
	SALLOC(s->tile_state, (size_t)s->max_tiles * 8);
	s->out_cap = 1u << 16;
	SALLOC(s->out, s->out_cap * 8);
	if (getenv("ACM_TRACE"))
		SALLOC(s->trace, 1024 * 4 * 8);
	if (s->p.mode == ACM_MODE_SAMPLED4) {
		/* room for one filter survivor per KiB of input (random data with 15 000 signatures: one per
		 * 5 KiB; the bench plants one signature per 10 KiB), at least 64 per scanning warp; beyond
		 * that the scan kernel resolves inline */
		const uint32_t regions = (uint32_t)dev->sm_count * (S4_THREADS / 32);
		const uint64_t want = (max_bytes >> 10) / regions;
		s->vq_cap = 64;
		while (s->vq_cap < want && s->vq_cap < 4096)
			s->vq_cap <<= 1;
		/* test hook: a tiny queue sends most survivors down the inline path of the scan kernel */
		if (getenv("ACM_VQ_CAP") && atoi(getenv("ACM_VQ_CAP")) > 0)
			s->vq_cap = (uint32_t)atoi(getenv("ACM_VQ_CAP"));
		SALLOC(s->vq, (size_t)regions * s->vq_cap * sizeof(uint4));
		SALLOC(s->vq_count, (size_t)regions * 4);
		/* dense chunks (2 KiB) a warp may have to hand on: all of its share and then some (the work
		 * hand-out is dynamic); a full list makes the scanning warp walk the chunk itself */
		s->dq_cap = (uint32_t)(2 * ((max_bytes >> 11) / regions) + 64);
		if (getenv("ACM_DQ_CAP") && atoi(getenv("ACM_DQ_CAP")) >= 0)
			s->dq_cap = (uint32_t)atoi(getenv("ACM_DQ_CAP"));   /* test hook */
		SALLOC(s->dq, ((size_t)regions * s->dq_cap + 1) * 4);
		SALLOC(s->dq_count, (size_t)regions * 4);
	}
#undef SALLOC
	if (cudaHostAlloc((void **)&s->h_flags, 64, cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **)&s->h_flags_dev, s->h_flags, 0) != cudaSuccess) {
		acm_set_error("scanner_create: cudaHostAlloc (mapped) failed");
		acm_scanner_free(s);
		return ACM_ERR_CUDA;
	}
	for (int i = 0; i < 4; i++)
		cudaEventCreate(&s->ev[i]);
	cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming);
	if (s->p.own_stream) {
		/* ACM_RESERVE_SMS (default 8 of 148): the scan kernel is persistent, one CTA per SM, and takes
		 * the whole SM; the few it leaves free are where the prefix sum / compaction / status kernels
		 * of the step before run, under this step's scan.  The scan kernel is bound by HBM, not by
		 * SM count: 140 CTAs stream the GiB in 0.167 ms, 148 in 0.162.  Measured 1 GiB step: 0.2153 ms
		 * on one stream; own streams with 0 / 2 / 4 / 8 / 12 / 16 / 24 SMs left free: 0.2068 / 0.2047 /
		 * 0.2045 / 0.2037 / 0.2056 / 0.2072 / 0.2155 */
		s->reserve_sms = getenv("ACM_RESERVE_SMS") ? (uint32_t)atoi(getenv("ACM_RESERVE_SMS")) : 8u;
		if (s->reserve_sms >= (uint32_t)dev->sm_count)
			s->reserve_sms = 0;
		if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&s->ev_k1, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&s->ev_dep, cudaEventDisableTiming) != cudaSuccess) {
			acm_set_error("scanner_create: cannot create the scanner's stream");
			acm_scanner_free(s);
			return ACM_ERR_CUDA;
		}
	}
	for (int i = 0; i < 2; i++) {
		cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&s->ev_free[i], cudaEventDisableTiming);
	}
	*out = s;
	return ACM_OK;
}

static int
grow(uint64_t **buf, uint64_t *cap, uint64_t need, const char *what)
{
	if (need <= *cap)
		return ACM_OK;
	uint64_t nc = *cap ? *cap : 1024;
	while (nc < need)
		nc *= 2;
	cudaFree(*buf);
	*buf = NULL;
	*cap = 0;
	if (cudaMalloc((void **)buf, nc * 8) != cudaSuccess) {
		acm_set_error("scan: cannot allocate %llu bytes for %s", (unsigned long long)(nc * 8), what);
		cudaGetLastError();
		return ACM_ERR_CUDA;
	}
	*cap = nc;
	return ACM_OK;
}


/*
 * Symbols per thread of k_scan_dfa.  The walk is latency bound, so a scan is cut into one chunk
 * per thread the GPU can hold (DFA_MINB CTAs of 256 per SM) -- but never below twice the halo (the
 * cold-start walk every chunk repeats) or 128 symbols, and no more than 4096.
 */
static uint64_t
scan_dfa_chunk(const struct acm_scanner *s, uint64_t span, uint64_t chains_per_sm)
{
	if (s->p.dfa_chunk > 0)
		return (uint64_t)s->p.dfa_chunk;
	const uint64_t resident = (uint64_t)s->dev->sm_count * chains_per_sm;
	const uint64_t halo = s->aut->max_len > 0 ? (uint64_t)(s->aut->max_len - 1) : 0;
	uint64_t chunk = (span + resident - 1) / resident;
	if (chunk < 2 * halo)
		chunk = 2 * halo;
	if (chunk < 128)
		chunk = 128;
	if (chunk > 4096)
		chunk = 4096;
	return (chunk + 15) & ~(uint64_t)15;
}

static int
launch_k1(struct acm_scanner *s, cudaStream_t st, const void *d_data, uint64_t n, const EmitCtx &E,
    int zero_work_counter, uint32_t *launches)
{
	const struct acm_automaton *a = s->aut;
	const uint64_t limit = E.emit_hi < n ? E.emit_hi : n;
	const uint64_t halo = a->max_len > 0 ? (uint64_t)(a->max_len - 1) : 0;
	uint64_t scan_lo = E.emit_lo > halo ? E.emit_lo - halo : 0;
	if (scan_lo < E.valid_lo)
		scan_lo = E.valid_lo;
	const uint64_t vec_lo = scan_lo >> 4, vec_hi = (limit + 15) >> 4;

	if (limit <= E.emit_lo)
		return ACM_OK;
	if (s->p.mode == ACM_MODE_SAMPLED4) {
		/* persistent: one CTA per SM; CTAs pull 512 KiB blocks from flags[6], their warps 16 KiB runs
		 * from the block; then the queued candidates are compared in full, one CTA per scanning warp */
		const uint64_t unit = 32ull * S4_UNROLL * S4_UNIT_CHUNKS * (S4_THREADS / 32);
		uint64_t blocks = (vec_hi - vec_lo + unit - 1) / unit;
		/* the SMs left to the other scanner's post-pass: that work has a whole scan to hide under, so
		 * the longer the scan, the fewer it needs (8 for 1 GiB, 2 from 4 GiB on) */
		uint32_t reserve = s->reserve_sms;
		const uint64_t span_bytes = (vec_hi - vec_lo) * 16;
		if (reserve > 2 && span_bytes > (1ull << 30)) {
			reserve = (uint32_t)((uint64_t)reserve * (1ull << 30) / span_bytes);
			if (reserve < 2)
				reserve = 2;
		}
		if (blocks > (uint64_t)s->dev->sm_count - reserve)
			blocks = (uint64_t)s->dev->sm_count - reserve;
		if (zero_work_counter)
			CUDA_TRY(cudaMemsetAsync(s->flags + 6, 0, 8, st));
		EmitCtx Eq = E;
		Eq.vq = s->vq;
		Eq.vq_count = s->vq_count;
		Eq.vq_cap = s->vq_cap;
		Eq.dq = s->dq;
		Eq.dq_count = s->dq_count;
		Eq.dq_cap = s->dq_cap;
		if (a->d.sample_stride == 8)
			CUDA_TRY(launch_dep(k_scan_sampled<8>, dim3((unsigned)blocks), dim3(S4_THREADS), S4_SMEM_BYTES, st, a->d, Eq,
			    (const uint8_t *)d_data, n, vec_lo, vec_hi, limit, s->flags + 6));
		else
			CUDA_TRY(launch_dep(k_scan_sampled<4>, dim3((unsigned)blocks), dim3(S4_THREADS), S4_SMEM_BYTES, st, a->d, Eq,
			    (const uint8_t *)d_data, n, vec_lo, vec_hi, limit, s->flags + 6));
		if (s->p.timing == 3)       /* the streaming kernel alone */
			CUDA_TRY(cudaEventRecord(s->ev[1], st));
		/* dense chunks (zero pages, padding, repeated prologues): their own kernel, hot rows of the
		 * row-displaced table in shared memory, launched as the programmatic dependent of
		 * k_resolve_queue so that it costs nothing when no chunk was queued (ACM_DENSE_KERNEL=0, or no
		 * such table: k_resolve_queue walks them, table through L1) */
		const int dense_kernel = a->d.xd_tab && (size_t)a->d.xd_d1_end * 4 + 64 <= XD_SMEM_MAX &&
		    !(getenv("ACM_DENSE_KERNEL") && !atoi(getenv("ACM_DENSE_KERNEL")));
		CUDA_TRY(launch_dep(dense_kernel ? k_resolve_queue<false> : k_resolve_queue<true>,
		    dim3((unsigned)blocks * (S4_THREADS / 32)), dim3(RQ_THREADS), 0, st, a->d, Eq,
		    (const uint8_t *)d_data, n, limit, vec_lo, (uint32_t)a->d.sample_stride));
		*launches += 1;
		if (dense_kernel) {
			uint32_t slots = a->d.xd_len < XD_SMEM_DEFAULT / 4 ? a->d.xd_len : XD_SMEM_DEFAULT / 4;
			slots &= ~3u;
			if (slots < ((a->d.xd_d1_end + 3) & ~3u))
				slots = (a->d.xd_d1_end + 3) & ~3u;
			const uint32_t nreg = (uint32_t)blocks * (S4_THREADS / 32);
			CUDA_TRY(launch_dep(k_dense_walk, dim3((unsigned)blocks), dim3(XD_THREADS),
			    (size_t)slots * 4 + 16 + ((size_t)nreg + 1) * 4 + 64, st, a->d, Eq,
			    (const uint8_t *)d_data, limit, vec_lo, (uint32_t)a->d.sample_stride, nreg, slots));
			*launches += 1;
		}
		if (a->d.split_len) {
			/* mixed set: the patterns shorter than split_len, second pass into the same buckets */
			const uint64_t lead = (uint64_t)a->d.split_len - 2;      /* longest short pattern - 1 */
			uint64_t lo2 = E.emit_lo > lead ? E.emit_lo - lead : 0;
			if (lo2 < E.valid_lo)
				lo2 = E.valid_lo;
			const uint64_t tile = (uint64_t)S2_THREADS * S2_UNROLL, v2 = lo2 >> 4;
			uint64_t b2 = (vec_hi - v2 + tile - 1) / tile;
			if (b2 > (uint64_t)s->dev->sm_count * 2)
				b2 = (uint64_t)s->dev->sm_count * 2;
			k_scan_start2<true><<<(unsigned)b2, S2_THREADS, S2_SMEM_BYTES(1), st>>>(a->d, E,
			    (const uint8_t *)d_data, n, v2, vec_hi, limit, a->d.b2s, a->d.split_len - 1);
			*launches += 1;
		}
	} else if (s->p.mode == ACM_MODE_CDFA && s->rd) {
		/* persistent: one CTA per SM holding the whole table, warps take regions of 32 chunks round-robin */
		const uint64_t regions = (((limit - 1) >> E.shift) - (E.emit_lo >> E.shift) + 32) / 32;
		uint64_t blocks = (regions + s->rd_warps - 1) / s->rd_warps;
		if (blocks > (uint64_t)s->dev->sm_count)
			blocks = s->dev->sm_count;
		RdCtx R;
		R.log = (uint2 *)s->buckets;
		R.loglen = s->rd_loglen;
		R.log_cap = s->rd_log_cap;
		R.tab_bytes = s->rd_tab_bytes;
		R.mode = E.direct ? RD_MODE_DIRECT : RD_MODE_LOG;
		R.lo = (uint32_t)a->d.cd_range_lo;
		R.cmax = a->d.cd_classes - 1;
		R.tabw_out = s->flags + 8;
		const size_t smem = ((s->rd_tab_bytes + 1023) & ~1023u) + (size_t)s->rd_warps * RD_WARP_SMEM + 64;
		k_scan_rd<<<(unsigned)blocks, s->rd_warps * 32, smem, st>>>(a->d, E, R, (const uint8_t *)d_data, limit);
	} else if (s->p.mode == ACM_MODE_CDFA) {
		/* persistent: one CTA per SM holding the hot rows, threads stride over chunk pairs */
		const uint64_t pairs = (((limit - 1) >> E.shift) - (E.emit_lo >> E.shift) + 2) / 2;
		uint64_t blocks = (pairs + CD_THREADS - 1) / CD_THREADS;
		if (blocks > (uint64_t)s->dev->sm_count)
			blocks = s->dev->sm_count;
		const size_t smem = s->cd_tab_bytes + (a->d.cd_range_lo >= 0 ? 0 : CD_LUT_WORDS * 4) + 16;
		if (a->d.cd_range_lo >= 0)
			k_scan_cdfa<true><<<(unsigned)blocks, CD_THREADS, smem, st>>>(a->d, E, (const uint8_t *)d_data, limit,
			    s->cd_hot, s->cd_tab_bytes);
		else
			k_scan_cdfa<false><<<(unsigned)blocks, CD_THREADS, smem, st>>>(a->d, E, (const uint8_t *)d_data, limit,
			    s->cd_hot, s->cd_tab_bytes);
	} else if (s->p.mode == ACM_MODE_START2) {
		const uint64_t tile = (uint64_t)S2_THREADS * S2_UNROLL;
		uint64_t blocks = (vec_hi - vec_lo + tile - 1) / tile;
		if (blocks > (uint64_t)s->dev->sm_count * 2)
			blocks = (uint64_t)s->dev->sm_count * 2;
		k_scan_start2<false><<<(unsigned)blocks, S2_THREADS, S2_SMEM_BYTES(0), st>>>(a->d, E,
		    (const uint8_t *)d_data, n, vec_lo, vec_hi, limit, a->d.b3, 0xFFFFFFFFu);
	} else if (a->d.xd_tab && (size_t)a->d.xd_d1_end * 4 + 64 <= XD_SMEM_MAX &&
	    !(getenv("ACM_DFA_DENSE") && atoi(getenv("ACM_DFA_DENSE")))) {
		/* the row-displaced table, its first ~200 KiB in shared memory: one persistent CTA per SM,
		 * threads take chunks round-robin (ACM_DFA_DENSE=1: the dense-table kernel below instead) */
		const uint64_t chunk = scan_dfa_chunk(s, limit - E.emit_lo, 2 * XD_THREADS);      /* two chunks per thread */
		const uint64_t nthreads = (limit - E.emit_lo + chunk - 1) / chunk;
		uint64_t blocks = ((nthreads + 1) / 2 + 1 + XD_THREADS - 1) / XD_THREADS;
		if (blocks > (uint64_t)s->dev->sm_count)
			blocks = s->dev->sm_count;
		/* half of the SM's 256 KiB for the staged prefix, the other half stays L1 for the slots behind
		 * it (measured: 10 000 ClamAV signatures 720 GB/s with 64-128 KiB staged, 700 with 224; 2 000
		 * packet-size signatures 1 231 GB/s with 128, 951 with 224) */
		uint32_t slots = a->d.xd_len < XD_SMEM_DEFAULT / 4 ? a->d.xd_len : XD_SMEM_DEFAULT / 4;
		slots &= ~3u;
		if (getenv("ACM_XD_SMEM_KB") && (uint32_t)atoi(getenv("ACM_XD_SMEM_KB")) * 256 < slots)
			slots = (uint32_t)atoi(getenv("ACM_XD_SMEM_KB")) * 256;      /* experiments: smaller staged prefix */
		if (slots < ((a->d.xd_d1_end + 3) & ~3u))
			slots = (a->d.xd_d1_end + 3) & ~3u;         /* the rows of depth <= 1 at least: the kernel reads them unconditionally from shared memory */
		const size_t smem = (size_t)slots * 4 + 64;
		if (a->alpha == 256)
			k_scan_xd<uint8_t><<<(unsigned)blocks, XD_THREADS, smem, st>>>(a->d, E, (const uint8_t *)d_data, n,
			    chunk, nthreads, s->flags + 2, slots);
		else
			k_scan_xd<uint16_t><<<(unsigned)blocks, XD_THREADS, smem, st>>>(a->d, E, (const uint16_t *)d_data,
			    n, chunk, nthreads, s->flags + 2, slots);
	} else {
		const uint64_t chunk = scan_dfa_chunk(s, limit - E.emit_lo, 256 * DFA_MINB);
		const uint64_t nthreads = (limit - E.emit_lo + chunk - 1) / chunk;
		const uint64_t blocks = (nthreads + 1 + 255) / 256;
		if (a->alpha == 256)
			k_scan_dfa<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(a->d, E, (const uint8_t *)d_data, n,
			    chunk, nthreads, s->flags + 2);
		else
			k_scan_dfa<uint16_t><<<(unsigned)blocks, 256, 0, st>>>(a->d, E, (const uint16_t *)d_data,
			    n, chunk, nthreads, s->flags + 2);
	}
	CUDA_TRY(cudaGetLastError());
	return ACM_OK;
}

static void
launch_k3(struct acm_scanner *s, cudaStream_t st, uint32_t nb, uint32_t k3_blocks, int with_push)
{
	if (s->p.mode == ACM_MODE_CDFA && s->rd) {
		/* every logged hit expands on its own: grid-stride over the regions, one warp per region */
		const uint32_t nreg = (nb + 31) / 32, w = RD_K3_THREADS / 32;
		uint32_t blocks = (nreg + w - 1) / w;
		if (blocks > (uint32_t)s->dev->sm_count * 8)
			blocks = (uint32_t)s->dev->sm_count * 8;
		launch_dep(k_rd_expand, dim3(blocks), dim3(RD_K3_THREADS), RD_K3_SMEM, st, (const uint2 *)s->buckets, s->rd_loglen,
		    s->rd_log_cap, nreg, s->offsets, nb, s->out, s->out_cap, s->flags, s->aut->d.rd_flat4,
		    s->pend.emit_lo >> s->shift, s->shift);
		return;
	}
	if (s->p.mode == ACM_MODE_CDFA) {
		/* hits are in order already: an expanding copy, one warp per bucket */
		uint32_t blocks = (nb + 8 * K3X_NB - 1) / (8 * K3X_NB);
		if (blocks > (uint32_t)s->dev->sm_count * 8)
			blocks = (uint32_t)s->dev->sm_count * 8;
		launch_dep(k_bucket_expand_compact, dim3(blocks), dim3(256), 0, st, (const uint32_t *)s->buckets, s->counts,
		    s->offsets, s->out, s->cap, nb, s->out_cap, s->flags, s->aut->d.cd_flat4, s->pend.emit_lo >> s->shift, s->shift);
		return;
	}
	/* with_push: the step's own launch also stores the keys into the gather region; the relaunch
	 * after the output buffer grew does not (the host pushes then, scan_complete) */
	launch_dep(k_bucket_sort_compact, dim3(k3_blocks), dim3(K3_THREADS), (size_t)s->cap * 8, st, s->buckets, s->counts,
	    s->offsets, s->out, s->cap, nb, s->out_cap, s->flags, with_push ? s->pend.push_dst : (uint64_t *)NULL,
	    s->pend.push_cap, s->pend.push_add);
}

/*
 * First half of a scan: queue memset -> K1 -> K2 -> K3 (-> push) -> 32-byte flag readback on
 * `st` and record ev_done.  Nothing is waited for.
 */
static int
scan_launch(struct acm_scanner *s, cudaStream_t st, const void *d_data, uint64_t n, uint64_t valid_lo,
    uint64_t emit_lo, uint64_t emit_hi, const struct acm_push_target *push)
{
	EmitCtx E;
	uint32_t nb, launches = 0;
	int rc, timing = s->p.timing;

	if (s->pend.active) {
		acm_set_error("scan: the previous asynchronous scan on this scanner was not finished");
		return ACM_ERR_STATE;
	}
	if (emit_hi > n)
		emit_hi = n;
	s->last_n = 0;
	memset(&s->pend, 0, sizeof(s->pend));
	s->pend.st = st;
	s->pend.emit_lo = emit_lo;
	s->pend.emit_hi = emit_hi;
	if (push) {
		s->pend.push_dst = push->d_dst;
		s->pend.push_cap = push->cap;
		s->pend.push_add = push->key_add;
	}
	if (emit_lo >= emit_hi) {
		s->pend.active = 2;         /* empty window: nothing queued */
		return ACM_OK;
	}
	if (emit_hi - emit_lo > s->max_bytes) {
		acm_set_error("scan: %llu bytes exceed the scanner's max_bytes %llu",
		    (unsigned long long)(emit_hi - emit_lo), (unsigned long long)s->max_bytes);
		return ACM_ERR_ARG;
	}
	if (((uintptr_t)d_data & 15) != 0) {
		acm_set_error("scan: device buffer must be 16-byte aligned");
		return ACM_ERR_ARG;
	}
	if (n >= (1ull << 40)) {
		acm_set_error("scan: more than 2^40 symbols in one call");
		return ACM_ERR_LIMIT;
	}
	/* record counts and offsets are 32-bit: the dense-output kernels can produce four records per
	 * byte, so one call of theirs covers at most 2^30 bytes (the sparse kernels stay far below
	 * one record per 2^8 bytes: their limit is the 2^40 above) */
	if (s->p.mode == ACM_MODE_CDFA && emit_hi - emit_lo > (1ull << 30)) {
		acm_set_error("scan: the dense-output kernel takes at most 2^30 bytes per call (32-bit record offsets); "
		    "split the scan");
		return ACM_ERR_LIMIT;
	}
	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	if (s->densify) {
		/*
		 * The previous scan overflowed its buckets and paid for the exact two-pass path.
		 * Dense output (a lexicon over text: one match per ~9 bytes) wants 4 KiB buckets of
		 * 1024 records -- 2 bytes of bucket per input byte, so only when that fits easily.
		 */
		size_t free_b = 0, total_b = 0;
		const size_t want = (size_t)((s->max_bytes >> 12) + 2) * 1024 * 8;
		s->densify = 0;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
		    want < free_b / 2 + (size_t)s->max_buckets * s->cap * 8) {
			const uint32_t old_shift = s->shift, old_cap = s->cap;
			s->shift = 12;
			s->cap = 1024;
			if (scanner_alloc_buckets(s) != ACM_OK) {
				s->shift = old_shift;
				s->cap = old_cap;
				if ((rc = scanner_alloc_buckets(s)) != ACM_OK)
					return rc;
			}
		}
	}
	if (s->p.mode == ACM_MODE_CDFA)        /* chunks are cut on absolute multiples of 2^shift */
		nb = (uint32_t)(((emit_hi - 1) >> s->shift) - (emit_lo >> s->shift) + 1);
	else
		nb = (uint32_t)((emit_hi - emit_lo + (1ull << s->shift) - 1) >> s->shift);

	memset(&E, 0, sizeof(E));
	E.buckets = s->buckets;
	E.counts = s->counts;
	E.overflow = s->flags;
	E.emit_lo = emit_lo;
	E.emit_hi = emit_hi;
	E.valid_lo = valid_lo;
	E.trace = s->trace;
	if (s->trace)
		CUDA_TRY(cudaMemsetAsync(s->trace, 0, 1024 * 4 * 8, st));
	E.cap = s->cap;
	E.shift = s->shift;

	/* flags (incl. the K1 work counters and the scan's tile counter), tile states, counts (k_scan_rd
	 * writes every chunk's count and every region's log length itself) */
	CUDA_TRY(cudaMemsetAsync(s->scratch, 0, 64 + (size_t)s->n_bucket_tiles * 8 + (s->rd ? 0 : (size_t)nb * 4), st));
	if (s->own_stream && st == s->own_stream && s->dev->k1_owner && s->dev->k1_owner != s)
		CUDA_TRY(cudaStreamWaitEvent(st, s->dev->k1_ev, 0));
	if (timing)
		CUDA_TRY(cudaEventRecord(s->ev[0], st));
	if ((rc = launch_k1(s, st, d_data, n, E, 0, &launches)) != ACM_OK)
		return rc;
	launches++;
	if (s->own_stream && st == s->own_stream) {
		CUDA_TRY(cudaEventRecord(s->ev_k1, st));
		s->dev->k1_ev = s->ev_k1;
		s->dev->k1_owner = s;
	}
	if (timing && !(timing == 3 && s->p.mode == ACM_MODE_SAMPLED4))
		CUDA_TRY(cudaEventRecord(s->ev[1], st));
	const int cdfa = s->p.mode == ACM_MODE_CDFA;
	if ((rc = launch_exclusive_scan(st, s->counts, s->offsets, nb, s->bucket_tiles, s->flags + 3,
	    s->flags + 1, 1)) != ACM_OK)
		return rc;
	launches++;
	if (timing == 1)
		CUDA_TRY(cudaEventRecord(s->ev[2], st));
	/*
	 * K3 goes out right behind K2, before the host knows the total: it guards itself against
	 * an overflowed scan (flags[0]) and against an output buffer that is too small (flags[5]);
	 * one synchronisation per step instead of two.  For the bucketed kernels it is the gather
	 * push as well.
	 */
	const uint32_t k3_warps = K3_THREADS / 32;
	uint32_t k3_blocks = (nb + k3_warps - 1) / k3_warps;
	if (k3_blocks > (uint32_t)s->dev->sm_count * 16)
		k3_blocks = (uint32_t)s->dev->sm_count * 16;
	launch_k3(s, st, nb, k3_blocks, 1);
	CUDA_TRY(cudaGetLastError());
	launches++;
	if (timing == 1)
		CUDA_TRY(cudaEventRecord(s->ev[3], st));
	if (cdfa && s->pend.push_dst) {
		/* the gather push reads the total on the device; it skips itself in the cases the
		 * host repairs in scan_complete (overflow, output or region too small) */
		k_push_keys_dev<<<s->dev->sm_count, 256, 0, st>>>(s->out, s->flags, s->out_cap, s->pend.push_dst,
		    s->pend.push_cap, s->pend.push_add);
		CUDA_TRY(cudaGetLastError());
		launches++;
	}
	/*
	 * The status words go to the host as eight posted stores of a one-warp kernel into mapped
	 * pinned memory: a 32-byte cudaMemcpyAsync here put a copy-engine round trip between every
	 * two steps queued on this stream (1 GiB step 0.2255 -> 0.2190 ms).
	 */
	CUDA_TRY(launch_dep(k_publish_flags, dim3(1), dim3(32), 0, st, s->flags, s->h_flags_dev));
	launches++;
	CUDA_TRY(cudaEventRecord(s->ev_done, st));
	s->pend.active = 1;
	s->pend.d_data = d_data;
	s->pend.n = n;
	s->pend.E = E;
	s->pend.nb = nb;
	s->pend.k3_blocks = k3_blocks;
	s->pend.launches = launches;
	s->pend.out_cap_at_launch = s->out_cap;
	return ACM_OK;
}

/*
 * Second half: wait for the flag readback, repair the rare cases (output buffer too small,
 * bucket overflow -> exact two-pass path) and fill the result.
 */
static int
scan_complete(struct acm_scanner *s, struct acm_scan_result *res)
{
	int rc, timing = s->p.timing, repaired = 0;

	if (res)
		memset(res, 0, sizeof(*res));
	if (!s->pend.active) {
		acm_set_error("scan_finish: no scan is pending on this scanner");
		return ACM_ERR_STATE;
	}
	if (s->pend.active == 2) {
		s->pend.active = 0;
		return ACM_OK;
	}
	s->pend.active = 0;
	cudaStream_t st = s->pend.st;
	EmitCtx E = s->pend.E;
	const uint32_t nb = s->pend.nb, k3_blocks = s->pend.k3_blocks;
	uint32_t launches = s->pend.launches;
	const uint64_t emit_lo = s->pend.emit_lo, emit_hi = s->pend.emit_hi;

	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	CUDA_TRY(cudaEventSynchronize(s->ev_done));

	const uint32_t overflow = s->h_flags[0];
	const uint64_t total = s->h_flags[1];
	if (total > s->out_cap) {
		if ((rc = grow(&s->out, &s->out_cap, total, "the match list")) != ACM_OK)
			return rc;
		repaired = 1;
		if (!overflow) {
			launch_k3(s, st, nb, k3_blocks, 0);
			CUDA_TRY(cudaGetLastError());
			launches++;
			if (timing == 1)
				CUDA_TRY(cudaEventRecord(s->ev[3], st));
		}
	}
	if (total && !overflow) {
		/* the sorted list is in place */
	} else if (total) {
		/* exact two-pass path: counts are exact, refill straight into place, then sort */
		int bits = ACM_KEY_PAT_BITS;
		uint64_t span = emit_hi;
		const uint64_t nblocks = (total + RS_TILE - 1) / RS_TILE;
		uint64_t hist_need = (256 * nblocks + 256 + 1) / 2;   /* in u64 units for grow() */

		repaired = 1;
		while (span) {
			bits++;
			span >>= 1;
		}
		if ((rc = grow(&s->tmp, &s->tmp_cap, total, "the sort buffer")) != ACM_OK)
			return rc;
		if ((rc = grow((uint64_t **)&s->hist, &s->hist_cap, hist_need, "sort histograms")) != ACM_OK)
			return rc;
		const uint32_t scan_tiles = (uint32_t)((256 * nblocks + SCAN_TILE - 1) / SCAN_TILE) + 1;
		if (scan_tiles > s->max_tiles) {
			cudaFree(s->tile_state);
			s->tile_state = NULL;
			CUDA_TRY(cudaMalloc((void **)&s->tile_state, (size_t)scan_tiles * 8));
			s->max_tiles = scan_tiles;
		}
		E.direct = 1;
		E.offsets = s->offsets;
		E.out = s->out;
		CUDA_TRY(cudaMemsetAsync(s->counts, 0, (size_t)nb * 4, st));
		if ((rc = launch_k1(s, st, s->pend.d_data, s->pend.n, E, 1, &launches)) != ACM_OK)
			return rc;
		launches++;
		/* a CDFA walk writes every chunk's records in order at its scanned offset: sorted already */
		if (s->p.mode != ACM_MODE_CDFA &&
		    (rc = radix_sort_impl(st, s->out, s->tmp, total, 0, bits, 0, s->hist, s->tile_state,
		    s->flags + 3, &launches)) != ACM_OK)
			return rc;
		if (timing == 1)
			CUDA_TRY(cudaEventRecord(s->ev[3], st));
	}
	s->last_n = total;
	if (s->pend.push_dst && total) {
		if (total > s->pend.push_cap) {
			acm_set_error("scan: %llu keys exceed the gather region (%llu)", (unsigned long long)total,
			    (unsigned long long)s->pend.push_cap);
			return ACM_ERR_LIMIT;
		}
		if (repaired) {             /* the step's own push skipped itself */
			uint64_t blocks = (total + 255) / 256;
			if (blocks > 592)
				blocks = 592;
			k_push_keys<<<(unsigned)blocks, 256, 0, st>>>(s->out, total, s->pend.push_dst,
			    s->pend.push_add);
			CUDA_TRY(cudaGetLastError());
			launches++;
		}
	}
	if (repaired)
		CUDA_TRY(cudaStreamSynchronize(st));
	if (overflow && !s->user_shape && s->shift > 12)
		s->densify = 1;
	if (res) {
		res->n_matches = total;
		res->n_bytes = emit_hi - emit_lo;
		res->mode = s->p.mode;
		res->fallback = overflow ? 1 : 0;
		res->final_state = s->h_flags[2];
		res->n_buckets = nb;
		res->launches = launches;
		if (timing)
			cudaEventElapsedTime(&res->ms_scan, s->ev[0], s->ev[1]);
		if (timing == 1) {
			cudaEventElapsedTime(&res->ms_prefix, s->ev[1], s->ev[2]);
			cudaEventElapsedTime(&res->ms_compact, s->ev[2], s->ev[3]);
			cudaEventElapsedTime(&res->ms_total, s->ev[0], s->ev[3]);
		}
	}
	return ACM_OK;
}

/* the stream a scan of this scanner is queued on; with an own stream, first ordered behind
 * everything the caller has queued on the device's stream so far (the input, usually) */
static int
scanner_stream(struct acm_scanner *s, cudaStream_t *st)
{
	*st = s->dev->stream;
	if (!s->own_stream)
		return ACM_OK;
	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	CUDA_TRY(cudaEventRecord(s->ev_dep, s->dev->stream));
	CUDA_TRY(cudaStreamWaitEvent(s->own_stream, s->ev_dep, 0));
	*st = s->own_stream;
	return ACM_OK;
}

static int
scan_on_stream(struct acm_scanner *s, cudaStream_t st, const void *d_data, uint64_t n, uint64_t valid_lo,
    uint64_t emit_lo, uint64_t emit_hi, struct acm_scan_result *res)
{
	int rc = scan_launch(s, st, d_data, n, valid_lo, emit_lo, emit_hi, NULL);

	if (rc != ACM_OK) {
		if (res)
			memset(res, 0, sizeof(*res));
		return rc;
	}
	return scan_complete(s, res);
}

extern "C" int
acm_scan_device(struct acm_scanner *s, const void *d_data, uint64_t n, uint64_t emit_lo, uint64_t emit_hi,
    struct acm_scan_result *res)
{
	cudaStream_t st;
	int rc = scanner_stream(s, &st);
	return rc != ACM_OK ? rc : scan_on_stream(s, st, d_data, n, 0, emit_lo, emit_hi, res);
}

extern "C" int
acm_scan_device_ex(struct acm_scanner *s, const void *d_data, uint64_t n, uint64_t valid_lo,
    uint64_t emit_lo, uint64_t emit_hi, struct acm_scan_result *res)
{
	if (valid_lo > emit_lo) {
		acm_set_error("scan: valid_lo must not exceed emit_lo");
		return ACM_ERR_ARG;
	}
	cudaStream_t st;
	int rc = scanner_stream(s, &st);
	return rc != ACM_OK ? rc : scan_on_stream(s, st, d_data, n, valid_lo, emit_lo, emit_hi, res);
}

extern "C" int
acm_scan_device_async(struct acm_scanner *s, const void *d_data, uint64_t n, uint64_t valid_lo,
    uint64_t emit_lo, uint64_t emit_hi, const struct acm_push_target *push)
{
	if (valid_lo > emit_lo) {
		acm_set_error("scan: valid_lo must not exceed emit_lo");
		return ACM_ERR_ARG;
	}
	cudaStream_t st;
	int rc = scanner_stream(s, &st);
	return rc != ACM_OK ? rc : scan_launch(s, st, d_data, n, valid_lo, emit_lo, emit_hi, push);
}

extern "C" int
acm_scan_finish(struct acm_scanner *s, struct acm_scan_result *res)
{
	return scan_complete(s, res);
}

/* ACM_TRACE=1: copies {t_entry, t_ready, t_exit, chunks} x n_ctas (ns, globaltimer) of the last scan kernel */
extern "C" int
acm_scan_trace(struct acm_scanner *s, uint64_t *h_out, uint32_t n_ctas)
{
	if (!s->trace) {
		acm_set_error("scan_trace: set ACM_TRACE=1 before creating the scanner");
		return ACM_ERR_STATE;
	}
	if (n_ctas > 1024)
		n_ctas = 1024;
	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	CUDA_TRY(cudaMemcpy(h_out, s->trace, (size_t)n_ctas * 32, cudaMemcpyDeviceToHost));
	return ACM_OK;
}

extern "C" void *
acm_scanner_stream(struct acm_scanner *s)
{
	return (void *)(s->own_stream ? s->own_stream : s->dev->stream);
}

extern "C" const uint64_t *
acm_scan_keys(struct acm_scanner *s)
{
	return s->out;
}

static int
ensure_h_keys(struct acm_scanner *s, uint64_t n)
{
	if (n <= s->h_keys_cap)
		return ACM_OK;
	uint64_t nc = s->h_keys_cap ? s->h_keys_cap : (1u << 16);
	while (nc < n)
		nc *= 2;
	if (s->h_keys)
		cudaFreeHost(s->h_keys);
	s->h_keys = NULL;
	s->h_keys_cap = 0;
	if (cudaHostAlloc((void **)&s->h_keys, nc * 8, cudaHostAllocDefault) != cudaSuccess) {
		acm_set_error("scan_fetch: cannot pin %llu bytes", (unsigned long long)(nc * 8));
		cudaGetLastError();
		return ACM_ERR_CUDA;
	}
	s->h_keys_cap = nc;
	return ACM_OK;
}

/* D2H of the last scan's keys on stream st, unpacked into h_off/h_pat with a signed shift */
static int64_t
fetch_on_stream(struct acm_scanner *s, cudaStream_t st, uint64_t base, int64_t rel_shift, uint64_t *h_off,
    uint32_t *h_pat, uint64_t cap)
{
	const uint64_t n = s->last_n < cap ? s->last_n : cap;
	int rc;

	if (n == 0)
		return 0;
	if ((rc = ensure_h_keys(s, n)) != ACM_OK)
		return rc;
	CUDA_TRY(cudaMemcpyAsync(s->h_keys, s->out, n * 8, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	for (uint64_t i = 0; i < n; i++) {
		const uint64_t k = s->h_keys[i];
		h_off[i] = base + (uint64_t)((int64_t)(k >> ACM_KEY_PAT_BITS) + rel_shift);
		h_pat[i] = (uint32_t)(k & ACM_KEY_PAT_MASK);
	}
	return (int64_t)n;
}

extern "C" int64_t
acm_scan_fetch(struct acm_scanner *s, uint64_t base, uint64_t *h_off, uint32_t *h_pat, uint64_t cap)
{
	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	return fetch_on_stream(s, s->own_stream ? s->own_stream : s->dev->stream, base, 0, h_off, h_pat, cap);
}

/* ---- peer gather (single node): CUDA IPC mapping + a store kernel ---- */

extern "C" int
acm_ipc_export(struct acm_device *d, void *d_ptr, void *handle64)
{
	cudaIpcMemHandle_t h;

	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaIpcGetMemHandle(&h, d_ptr));
	memcpy(handle64, &h, sizeof(h));
	return ACM_OK;
}

extern "C" int
acm_ipc_open(struct acm_device *d, const void *handle64, void **d_ptr)
{
	cudaIpcMemHandle_t h;

	memcpy(&h, handle64, sizeof(h));
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return ACM_OK;
}

extern "C" int
acm_ipc_close(struct acm_device *d, void *d_ptr)
{
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
	return ACM_OK;
}

extern "C" int
acm_scan_push_keys(struct acm_scanner *s, uint64_t *d_dst, uint64_t dst_index, uint64_t key_add)
{
	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	if (s->last_n == 0)
		return ACM_OK;
	uint64_t blocks = (s->last_n + 255) / 256;
	if (blocks > 592)
		blocks = 592;
	k_push_keys<<<(unsigned)blocks, 256, 0, s->own_stream ? s->own_stream : s->dev->stream>>>(s->out, s->last_n, d_dst + dst_index,
	    key_add);
	CUDA_TRY(cudaGetLastError());
	return ACM_OK;
}

extern "C" int
acm_scan_histogram(struct acm_scanner *s, uint64_t *d_counts)
{
	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	if (s->last_n == 0)
		return ACM_OK;
	uint64_t blocks = (s->last_n + 255) / 256;
	if (blocks > 1184)
		blocks = 1184;
	k_histogram<<<(unsigned)blocks, 256, 0, s->own_stream ? s->own_stream : s->dev->stream>>>(s->out, s->last_n,
	    (unsigned long long *)d_counts);
	CUDA_TRY(cudaGetLastError());
	return ACM_OK;
}

/*
 * Host-buffer scan: double-buffered H2D on the copy stream, scan on the device
 * stream.  Segment i is copied together with the Lmax-1 bytes before it, so each
 * segment is an independent halo scan and the concatenation of
 * the per-segment sorted lists is the sorted list of the whole stream.
 */
extern "C" int64_t
acm_scan_host(struct acm_scanner *s, const void *h_data, uint64_t n, uint64_t base, uint64_t *h_off,
    uint32_t *h_pat, uint64_t cap, struct acm_scan_result *res)
{
	return acm_scan_host_ex(s, h_data, n, 0, base, h_off, h_pat, cap, res);
}

/* the same with lead0 symbols of valid context in front of h_data (a shard of a longer host stream) */
extern "C" int64_t
acm_scan_host_ex(struct acm_scanner *s, const void *h_data, uint64_t n, uint64_t lead0, uint64_t base,
    uint64_t *h_off, uint32_t *h_pat, uint64_t cap, struct acm_scan_result *res)
{
	const uint8_t *src = (const uint8_t *)h_data;
	const uint64_t sym = s->aut->alpha == 256 ? 1 : 2;
	const uint64_t halo = s->aut->max_len > 0 ? (uint64_t)(s->aut->max_len - 1) : 0;
	uint64_t seg = s->max_bytes;
	uint64_t written = 0, found = 0, launches = 0;
	cudaStream_t ks = s->own_stream ? s->own_stream : s->dev->stream, cs = s->dev->copy_stream;
	struct acm_scan_result r1;
	int rc, fallback = 0;

	CUDA_TRY(cudaSetDevice(s->dev->ordinal));
	if (res)
		memset(res, 0, sizeof(*res));
	if (n == 0)
		return 0;
	if (seg > n)
		seg = n;
	const uint64_t nseg = (n + seg - 1) / seg;
	const uint64_t need = (halo + seg) * sym + 64;
	if (need > s->stage_bytes) {
		for (int i = 0; i < 2; i++) {
			cudaFree(s->stage[i]);
			s->stage[i] = NULL;
		}
		s->stage_bytes = 0;
		for (int i = 0; i < 2; i++)
			CUDA_TRY(cudaMalloc((void **)&s->stage[i], need));
		s->stage_bytes = need;
	}

	/* segment i travels with its `lead` symbols of context: [lo - lead, hi) -> stage[i & 1] */
	auto enqueue_copy = [&](uint64_t i) -> int {
		const uint64_t lo = i * seg, hi = (lo + seg < n) ? lo + seg : n;
		const uint64_t lead = lo + lead0 < halo ? lo + lead0 : halo;
		const int b = (int)(i & 1);
		if (i >= 2)
			CUDA_TRY(cudaStreamWaitEvent(cs, s->ev_free[b], 0));
		CUDA_TRY(cudaMemcpyAsync(s->stage[b], src + (lo - lead) * sym, (hi - lo + lead) * sym,
		    cudaMemcpyHostToDevice, cs));
		CUDA_TRY(cudaEventRecord(s->ev_copied[b], cs));
		return ACM_OK;
	};

	if ((rc = enqueue_copy(0)) != ACM_OK)
		return rc;
	for (uint64_t i = 0; i < nseg; i++) {
		const uint64_t lo = i * seg, hi = (lo + seg < n) ? lo + seg : n;
		const uint64_t lead = lo + lead0 < halo ? lo + lead0 : halo;
		const int b = (int)(i & 1);

		if (i + 1 < nseg && (rc = enqueue_copy(i + 1)) != ACM_OK)
			return rc;
		CUDA_TRY(cudaStreamWaitEvent(ks, s->ev_copied[b], 0));
		rc = scan_on_stream(s, ks, s->stage[b], lead + (hi - lo), 0, lead, lead + (hi - lo), &r1);
		if (rc != ACM_OK)
			return rc;
		/* K1 has finished (scan_on_stream synchronised): the staging buffer is free again */
		CUDA_TRY(cudaEventRecord(s->ev_free[b], ks));
		found += r1.n_matches;
		launches += r1.launches;
		fallback |= r1.fallback;
		if (written < cap) {
			int64_t got = fetch_on_stream(s, ks, base + lo, -(int64_t)lead, h_off + written,
			    h_pat + written, cap - written);
			if (got < 0)
				return got;
			written += (uint64_t)got;
		}
	}
	if (res) {
		res->n_matches = found;
		res->n_bytes = n;
		res->mode = s->p.mode;
		res->fallback = fallback;
		res->launches = (uint32_t)launches;
	}
	return (int64_t)found;
}

/* ------------------------------------------------------------------------- */
/* synthetic streams                                                         */
/* ------------------------------------------------------------------------- */

extern "C" void
acm_synth_fill_host(void *dst, uint64_t n, uint64_t seed, uint64_t offset)
{
	uint8_t *p = (uint8_t *)dst;
	uint64_t i = 0;

	while (i < n) {
		const uint64_t g = offset + i;
		const uint64_t w = acm_mix64(seed, g >> 3);
		if ((g & 7) == 0 && n - i >= 8) {
			memcpy(p + i, &w, 8);
			i += 8;
		} else {
			p[i] = (uint8_t)(w >> (8 * (g & 7)));
			i++;
		}
	}
}

extern "C" int
acm_synth_fill_device(struct acm_device *d, void *dst, uint64_t n, uint64_t seed, uint64_t offset)
{
	if (((uintptr_t)dst & 7) || (offset & 7) || (n & 7)) {
		acm_set_error("synth_fill_device: pointer, offset and size must be multiples of 8");
		return ACM_ERR_ARG;
	}
	CUDA_TRY(cudaSetDevice(d->ordinal));
	if (n == 0)
		return ACM_OK;
	k_synth_fill<<<d->sm_count * 8, 256, 0, d->stream>>>((uint64_t *)dst, n >> 3, seed, offset >> 3);
	CUDA_TRY(cudaGetLastError());
	return ACM_OK;
}

extern "C" int
acm_plant_device(struct acm_device *d, void *buf, uint64_t n, uint64_t buf_offset, const uint64_t *h_pos,
    const uint32_t *h_blob_off, const uint32_t *h_len, uint32_t count, const uint8_t *h_blob,
    uint32_t blob_bytes)
{
	uint8_t *mem = NULL;
	const size_t a = (size_t)count * 8, b = (size_t)count * 4;
	int rc = ACM_OK;

	if (count == 0)
		return ACM_OK;
	CUDA_TRY(cudaSetDevice(d->ordinal));
	CUDA_TRY(cudaMalloc((void **)&mem, a + 2 * b + blob_bytes + 64));
	cudaMemcpyAsync(mem, h_pos, a, cudaMemcpyHostToDevice, d->stream);
	cudaMemcpyAsync(mem + a, h_blob_off, b, cudaMemcpyHostToDevice, d->stream);
	cudaMemcpyAsync(mem + a + b, h_len, b, cudaMemcpyHostToDevice, d->stream);
	cudaMemcpyAsync(mem + a + 2 * b, h_blob, blob_bytes, cudaMemcpyHostToDevice, d->stream);
	k_plant<<<(count * 32 + 255) / 256, 256, 0, d->stream>>>((uint8_t *)buf, n, buf_offset,
	    (const uint64_t *)mem, (const uint32_t *)(mem + a), (const uint32_t *)(mem + a + b), count,
	    mem + a + 2 * b);
	if (cudaStreamSynchronize(d->stream) != cudaSuccess) {
		acm_set_error("plant_device: %s", cudaGetErrorString(cudaGetLastError()));
		rc = ACM_ERR_CUDA;
	}
	cudaFree(mem);
	return rc;
}
