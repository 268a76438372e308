/*
 * k1_scan.cuh -- K1: the Aho-Corasick scan kernels for sm_100a.
 *
 * Replaces the reference's `ahomatch` OpenCL kernel (reference ahomatch.cl:1-165
 * and its ushort twin AC_ushorts/ahomatch.cl:1-148).  The reference gives every
 * work-item one chunk and does one dependent, uncoalesced 4-byte global load per
 * input byte into a table of up to 1.29 GiB (ahomatch.cl:56-65).  Here:
 *
 *   k_scan_sampled<STRIDE> + k_resolve_queue  (patterns >= 7 bytes, e.g. the ClamAV sets)
 *       Every occurrence of a pattern of length >= 7 contains a 4-byte window that starts
 *       at a multiple of 4 (>= 10: at a multiple of 8).  The kernel streams the input once
 *       with coalesced 16-byte loads and tests only those aligned windows against f1, a
 *       blocked Bloom bitmap of one indexed window per (pattern, alignment) (128 KiB of
 *       shared memory, staged by TMA bulk copies).  Survivors (a few %) are re-tested
 *       lane-locally against f2 (96 KiB, second hash) and the few that remain are written
 *       to the scanning warp's queue region; k_resolve_queue then probes the exact gram
 *       table, fetches the candidate records and compares the candidate patterns with the
 *       text in CTA-wide phases.  0.125 shared-memory lookups per input byte instead of one
 *       table gather per byte.  Dense (repetitive) chunks fall back to a bounded DFA walk.
 *
 *   k_scan_start2<PAIR>  (any pattern length)
 *       A filter tested at every position ("can an automaton walk started here report
 *       anything"), hits walk the trie edges of T.  PAIR = false: blocked Bloom bitmap of the
 *       first three bytes of every pattern (mode 2); PAIR = true: exact bitmap over (byte,
 *       next byte) of the short patterns of a MIXED set, whose long patterns go through
 *       k_scan_sampled (both passes emit into the same buckets).
 *
 *   k_scan_dfa<SYM>  (bytes or ushort symbols; cross-check and AC_ushorts path)
 *       The textbook form: one thread per chunk, cold start Lmax-1 symbols early
 *       (SURVEY.md A.5), one table lookup per symbol, matches reported through the
 *       output links; chunks sized so that every resident thread has one.
 *
 *   k_scan_rd (k1_rd.cuh) / k_scan_cdfa<RANGE>  (small automata over few distinct bytes: word
 *       lists over text, one match per ~9 bytes)
 *       The same walk out of shared memory: k_scan_rd from a row-displaced table of 4-byte
 *       entries (one lookup per byte); k_scan_cdfa, for automata that table does not fit, from
 *       the first rows of the class-compressed 16-bit table with the rest served by L1/L2.
 *
 * All of them report exactly the set { (end offset, pattern index) } that a serial walk
 * of the reference's automaton reports with full match lists.
 *
 * Match emission: records go to the bucket of their END offset
 * ((end - emit_lo) >> shift); lanes of a warp that emit into the same bucket at
 * the same time are aggregated with match.any/popc into one atomicAdd.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "acm_tables.h"

struct AutDev {
	const uint32_t *T;
	const uint32_t *level_start;
	const uint32_t *own_begin;
	const uint32_t *own_pat;
	const uint32_t *olink;
	const uint32_t *f1;
	const uint32_t *f2;
	const acm_gram_slot *grams;
	const acm_cand *cand;
	const uint8_t  *pat_blob;
	const uint32_t *pat_off;
	const uint32_t *pat_len;
	const uint8_t  *pat_win;    /* [patterns][8]: offset of the indexed window per alignment (sampled filter) */
	const uint32_t *b2s;        /* start bitmap of the patterns shorter than split_len (mixed sets) */
	const uint32_t *b3;         /* start filter of mode 2: first-three-bytes Bloom bitmap of all patterns */
	uint32_t split_len;         /* 0: every pattern is in the sampled filter */
	const uint16_t *cd_tab;        /* class-compressed DFA (k_scan_cdfa), NULL if not built */
	const uint8_t  *cd_cls;
	const uint32_t *cd_flat_begin;
	const uint32_t *cd_flat_pat;
	const uint4    *cd_flat4;      /* [states]: list inline, .x = first | count << 24 */
	const uint32_t *rd_tab;        /* row-displaced form (k_scan_rd, all of it in shared memory), NULL if not built */
	const uint4    *rd_flat4;      /* [rd_len]: match list of the state whose base is the slot */
	uint32_t rd_len;
	const uint32_t *xd_tab;        /* the whole DFA row-displaced (k_scan_xd; acm_core.c:build_xd), NULL if not built */
	const uint32_t *xd_sid;        /* [xd_len]: breadth-first id of the state whose base is the slot */
	uint32_t xd_len;
	uint32_t xd_d1_end;            /* the rows of the states of depth <= 1 end here */
	uint32_t xd_sym_bits;          /* 8 (bytes) or 11 (ushort symbols) */
	uint32_t cd_thr4;              /* entry >= this: four patterns end there (code 3 = three or four) */
	uint32_t cd_classes;
	int      cd_range_lo;
	uint32_t gram_mask;
	uint32_t gram_shift;     /* 32 - log2(slots) */
	uint32_t num_states;
	int      alpha;
	int      max_len;
	int      sample_stride;
	uint32_t max_win;        /* largest indexed window offset of the sampled filter (<= ACM_CAND_O_MAX) */
};

struct EmitCtx {
	uint64_t *buckets;       /* [n_buckets][cap] record keys                         */
	uint32_t *counts;        /* [n_buckets] records produced (exact, may exceed cap) */
	uint32_t *overflow;      /* set to 1 when any bucket exceeded cap               */
	const uint32_t *offsets; /* direct mode: exclusive scan of counts               */
	uint64_t *out;           /* direct mode: dense output                           */
	uint64_t  emit_lo, emit_hi;
	uint64_t  valid_lo;      /* positions before this are never read nor used as walk starts */
	uint64_t *trace;         /* optional: per CTA {t_entry, t_ready, t_exit, chunks} (globaltimer ns) */
	uint4    *vq;            /* sampled kernel: per-warp regions of candidates awaiting the full compare */
	uint32_t *vq_count;      /* [regions] entries written (<= vq_cap)                */
	uint32_t  vq_cap;        /* entries per region                                   */
	uint32_t *dq;            /* sampled kernel: per-warp lists of DENSE chunks (first vector index), left to k_resolve_queue */
	uint32_t *dq_count;      /* [regions]                                            */
	uint32_t  dq_cap;
	uint32_t  cap;
	uint32_t  shift;
	int       direct;        /* 1: second pass of the exact two-pass path; 2: k_scan_rd's counting pass before it */
};

#define F1_WORDS (1u << (ACM_F1_BITS_LOG2 - 5))
#define F2_WORDS ACM_F2_WORDS

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

/*
 * Programmatic dependent launch.  The kernels of a step are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization (acm_cuda.cu:launch_dep): a kernel may be set up
 * and its CTAs made resident while its predecessor in the stream still runs, as soon as every CTA of
 * the predecessor has passed pdl_trigger() (or ended).  EVERY such kernel calls pdl_wait() before it
 * touches anything a predecessor wrote -- it returns when the predecessor grid has completed and its
 * stores are visible -- and on every path before it ends, so that "this grid is done" keeps
 * implying "everything before it in the stream is done" along the chain.  Both are no-ops in a
 * kernel launched the plain way.
 */
__device__ __forceinline__ void pdl_wait()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void pdl_trigger()
{
	asm volatile("griddepcontrol.launch_dependents;");
}

__device__ __forceinline__ uint64_t globaltimer_ns()
{
	uint64_t t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
	uint32_t m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

/* ---- TMA bulk copy + mbarrier (global -> shared) ---- */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
	    ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	    ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* pull [p, p + bytes) into L2; p 16-byte aligned, bytes a multiple of 16 */
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes)
{
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

/* ---- emission ---- */
__device__ __forceinline__ void emit_record(const EmitCtx &E, uint64_t end, uint32_t pat)
{
	if (end < E.emit_lo || end >= E.emit_hi)
		return;
	const uint32_t b = (uint32_t)((end - E.emit_lo) >> E.shift);
	const unsigned act = __activemask();
	const unsigned peers = __match_any_sync(act, b);
	const int leader = __ffs(peers) - 1;
	const int lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (lane == leader)
		base = atomicAdd(&E.counts[b], (uint32_t)__popc(peers));
	base = __shfl_sync(peers, base, leader);
	const uint32_t slot = base + (uint32_t)__popc(peers & lanemask_lt());
	const uint64_t key = (end << ACM_KEY_PAT_BITS) | pat;
	if (E.direct) {
		E.out[(uint64_t)E.offsets[b] + slot] = key;
	} else if (slot < E.cap) {
		E.buckets[(uint64_t)b * E.cap + slot] = key;
	} else {
		*E.overflow = 1u;
	}
}

__device__ __forceinline__ void emit_own(const AutDev &A, const EmitCtx &E, uint32_t state, uint64_t end)
{
	const uint32_t lo = __ldg(&A.own_begin[state]);
	const uint32_t hi = __ldg(&A.own_begin[state + 1]);
	for (uint32_t k = lo; k < hi; ++k)
		emit_record(E, end, __ldg(&A.own_pat[k]));
}

/*
 * Per-warp staging of match records in shared memory.  A walk emits from one lane at a
 * time, and a global atomicAdd per record is a ~500-cycle round trip on the emitting lane's
 * critical path; staged, a record costs one shared-memory atomic, and the whole warp later
 * flushes up to WQ_CAP records with one global atomic per (bucket, 32 records).
 */
#define WQ_CAP   128
#define WQ_FLUSH 64

struct WarpQueue {
	uint64_t *rec;          /* [WQ_CAP] in shared memory */
	uint32_t *count;        /* records appended since the last flush (may exceed WQ_CAP) */
};

__device__ __forceinline__ void stage_record(const WarpQueue &Q, const EmitCtx &E, uint64_t end, uint32_t pat)
{
	if (end < E.emit_lo || end >= E.emit_hi)
		return;
	const uint32_t slot = atomicAdd(Q.count, 1u);
	if (slot < WQ_CAP)
		Q.rec[slot] = (end << ACM_KEY_PAT_BITS) | pat;
	else
		emit_record(E, end, pat);           /* queue full: straight to the bucket */
}

/* whole warp, converged */
__device__ __forceinline__ void flush_queue(const WarpQueue &Q, const EmitCtx &E, int lane)
{
	__syncwarp();
	uint32_t n = *Q.count;
	if (n > WQ_CAP)
		n = WQ_CAP;
	for (uint32_t i = lane; i < n; i += 32) {
		const uint64_t key = Q.rec[i];
		emit_record(E, key >> ACM_KEY_PAT_BITS, (uint32_t)(key & ACM_KEY_PAT_MASK));
	}
	__syncwarp();
	if (lane == 0)
		*Q.count = 0;
	__syncwarp();
}

/*
 * Enter the automaton at position s: follow trie edges while they exist, report the
 * patterns ending at every node passed.  `limit` is exclusive.
 */
template <bool STAGED>
__device__ __forceinline__ void walk_from_t(const AutDev &A, const EmitCtx &E, const WarpQueue &Q,
    const uint8_t *__restrict__ data, uint64_t s, uint64_t limit, uint32_t max_depth)
{
	uint32_t state = 0;
	uint32_t d = 0;
	for (uint64_t pos = s; pos < limit && d < max_depth; ++pos) {
		const uint32_t e = __ldg(&A.T[(size_t)state * 256 + __ldg(&data[pos])]);
		const uint32_t nx = e & ACM_T_MASK;
		if (nx < __ldg(&A.level_start[d + 1]))
			break;
		state = nx;
		++d;
		if (e & ACM_T_OWN) {
			const uint32_t lo = __ldg(&A.own_begin[state]);
			const uint32_t hi = __ldg(&A.own_begin[state + 1]);
			for (uint32_t k = lo; k < hi; ++k) {
				if (STAGED)
					stage_record(Q, E, pos, __ldg(&A.own_pat[k]));
				else
					emit_record(E, pos, __ldg(&A.own_pat[k]));
			}
		}
	}
}

__device__ __noinline__ void walk_from(const AutDev A, const EmitCtx E, const uint8_t *__restrict__ data,
    uint64_t s, uint64_t limit)
{
	WarpQueue none = {nullptr, nullptr};
	walk_from_t<false>(A, E, none, data, s, limit, 0xFFFFFFFFu);
}

/*
 * 16 bytes at vector index idx.  The caller guarantees idx * 16 < n, and device buffers
 * are 16-byte aligned and readable up to the next multiple of 16 (acm.h), so the last,
 * partial vector is a plain load too; whatever lies beyond n can only raise a filter hit
 * that the exact stages then reject.
 */
__device__ __forceinline__ uint4 load_vec(const uint8_t *__restrict__ data, uint64_t idx)
{
	return __ldcs(reinterpret_cast<const uint4 *>(data) + idx);
}

/* ------------------------------------------------------------------------- */
/* sampled 4-gram entry filter                                               */
/* ------------------------------------------------------------------------- */

#ifndef S4_THREADS
#define S4_THREADS 1024
#endif
#define S4_UNROLL  4
#ifndef S4_UNIT_CHUNKS
#define S4_UNIT_CHUNKS 8                 /* chunks per run (16 KiB) while plenty of work is left */
#endif
#ifndef S4_INTERLEAVE
#define S4_INTERLEAVE 0                  /* 1: the 32 runs of a block are interleaved chunk by chunk */
#endif
#define S4_BLOCK_RUNS 32                 /* runs per block = warps per CTA */
#define DQ_ANY_WORD 10                   /* flags[10]: some warp queued a dense chunk this step */
#define S4_SLOTS 4                       /* published blocks per CTA (ring) */
#ifndef S4_PF_DIST
#define S4_PF_DIST 1                     /* L2 prefetch distance in chunks (1 .. S4_UNIT_CHUNKS) */
#endif
#define S4_SMEM_BYTES ((F1_WORDS + F2_WORDS) * 4 + 16 + 16 + S4_SLOTS * 16)
#define FULL_MASK 0xffffffffu

/*
 * Whole-warp verification of one candidate: does pattern `pid` occur at text position s?
 * Lane l compares bytes [4l, 4l+4) of each 128-byte round; all arguments are warp-uniform.
 */
__device__ __forceinline__ void s4_verify(const AutDev &A, const EmitCtx &E, const uint8_t *__restrict__ data,
    uint32_t pid, uint32_t len, uint64_t s, int lane)
{
	const uint32_t *pw = reinterpret_cast<const uint32_t *>(A.pat_blob + __ldg(&A.pat_off[pid]));
	const uint32_t *tw = reinterpret_cast<const uint32_t *>(data + (s & ~3ull));
	const uint32_t sh = (uint32_t)(s & 3) * 8;
	bool ok = true;
	for (uint32_t r = 0; r < len; r += 128) {
		const uint32_t k = r + 4 * (uint32_t)lane;        /* byte offset inside the pattern */
		uint32_t diff = 0;
		if (k < len) {
			/* aligned words around the unaligned text position; both contain valid bytes */
			const uint32_t rem = len - k;
			const uint32_t w0 = __ldg(tw + (k >> 2));
			/* the second word only when this lane's bytes really reach into it */
			const bool need1 = ((uint32_t)(s & 3) + (rem < 4 ? rem : 4u)) > 4u;
			const uint32_t w1 = need1 ? __ldg(tw + (k >> 2) + 1) : 0u;
			const uint32_t t = __funnelshift_r(w0, w1, sh);
			const uint32_t p = __ldg(pw + (k >> 2));
			const uint32_t mask = rem >= 4 ? 0xffffffffu : ((1u << (8 * rem)) - 1u);
			diff = (t ^ p) & mask;
		}
		ok = __all_sync(FULL_MASK, diff == 0);
		if (!ok)
			break;
	}
	if (ok && lane == 0)
		emit_record(E, s + len - 1, pid);
}

/*
 * Dense-hit fallback.  On repetitive input (zero pages, period-4/8 fills, code padding) almost
 * every aligned window is a pattern gram with a long candidate list, and filter + verify would
 * degenerate to (windows x candidates).  A 2 KiB chunk whose level-1 hit count says "this is not
 * random-looking data" is handed to the automaton instead, which is linear whatever the text.
 *
 * WHO REPORTS WHAT.  The filter finds an occurrence through its INDEXED window -- one per
 * (pattern, alignment of the start), not necessarily the first aligned one: the builder moves it
 * to a rarer later window when the first one's gram is popular (zero runs, common prologues:
 * exactly the material dense chunks are made of).  So an occurrence belongs to the chunk its
 * indexed window lies in, and the walk of a dense chunk reports exactly the occurrences whose
 * indexed window lies in that chunk -- whether they start before it (by up to ACM_CAND_O_MAX
 * bytes) or end far behind it.  (Round 1 let the walk own "starts whose FIRST aligned window lies
 * in the chunk": next to a filter-scanned chunk that reported some occurrences twice and lost
 * others; tools/density_sweep.py found it.)
 *
 * The scanning warp only queues the chunk; threads of k_resolve_queue walk it, in 1 to 8 slices
 * (a few dense chunks in a region: many short walks, so that a CTA does not wait for three threads;
 * all chunks dense: one long walk per thread, little overlap): cold start max_win bytes -- the
 * largest window offset in the index -- in front (nothing that starts earlier can have its window inside), the
 * row-displaced table (acm_core.c:build_xd) through L1 -- zero pages keep every lane on the same
 * few entries -- or the dense table when there is none, until no occurrence that began inside the
 * chunk can still be open.
 */
__device__ __noinline__ void s4_dense_chunk(const AutDev *__restrict__ Ap, const EmitCtx *__restrict__ Ep,
    const uint8_t *__restrict__ data, uint64_t chunk_lo, uint64_t chunk_hi, uint64_t limit, uint32_t stride)
{
	/* [chunk_lo, chunk_hi): the chunk, or the slice of it this thread owns the windows of */
	const AutDev &A = *Ap;
	const EmitCtx &E = *Ep;
	uint64_t pos = chunk_lo > A.max_win ? chunk_lo - A.max_win : 0;
	if (pos < E.valid_lo)
		pos = E.valid_lo;
	uint64_t end = chunk_hi + (uint64_t)A.max_len;
	if (end > limit)
		end = limit;
	const bool use_xd = A.xd_tab != nullptr;
	const uint32_t sh = A.xd_sym_bits + 1, smask = (1u << A.xd_sym_bits) - 1u;
	uint32_t os = 0, ob = 0, state = 0;
	/* the text comes 16 bytes per load and is shifted out of a 128-bit register: a 1-byte load per
	 * step from 32 different sectors per warp instruction made the L1 the bound of the whole walk */
	uint4 tv = make_uint4(0, 0, 0, 0);
	uint32_t have = 0;                                /* bytes left in tv */
	for (; pos < end; ++pos) {
		if (have == 0) {
			tv = __ldg(reinterpret_cast<const uint4 *>(data + (pos & ~15ull)));
			have = 16;
			for (uint32_t skip = (uint32_t)(pos & 15); skip; --skip) {      /* only before the first byte */
				tv.x = __funnelshift_r(tv.x, tv.y, 8);
				tv.y = __funnelshift_r(tv.y, tv.z, 8);
				tv.z = __funnelshift_r(tv.z, tv.w, 8);
				tv.w >>= 8;
				--have;
			}
		}
		const uint32_t c = tv.x & 0xFFu;
		tv.x = __funnelshift_r(tv.x, tv.y, 8);
		tv.y = __funnelshift_r(tv.y, tv.z, 8);
		tv.z = __funnelshift_r(tv.z, tv.w, 8);
		tv.w >>= 8;
		--have;
		bool any;
		if (use_xd) {
			const uint32_t r = __ldg(A.xd_tab + c), g = __ldg(A.xd_tab + ob + c);
			uint32_t x = g;
			if (os != ob)
				x = __ldg(A.xd_tab + os + c);
			if ((x & smask) != c)
				x = (g & smask) == c ? g : r;
			os = x >> sh;
			ob = r >> sh;
			any = (x >> (sh - 1)) & 1u;
			if (any || pos >= chunk_hi)
				state = __ldg(A.xd_sid + os);
		} else {
			const uint32_t e = __ldg(&A.T[(size_t)state * 256 + c]);
			state = e & ACM_T_MASK;
			any = (e & ACM_T_ANY) != 0;
		}
		if (any) {
			for (uint32_t v = state; v; v = __ldg(&A.olink[v])) {
				const uint32_t b = __ldg(&A.own_begin[v]), t = __ldg(&A.own_begin[v + 1]);
				for (uint32_t k = b; k < t; ++k) {
					const uint32_t pid = __ldg(&A.own_pat[k]);
					const uint64_t len = __ldg(&A.pat_len[pid]);
					/* mixed sets: the short patterns belong to the start-filter pass */
					if (len < A.split_len || pos + 1 < len)
						continue;
					const uint64_t s = pos + 1 - len;
					const uint64_t w = s + __ldg(&A.pat_win[(size_t)pid * 8 + ((stride - (uint32_t)(s % stride)) % stride)]);
					if (s >= E.valid_lo && w >= chunk_lo && w < chunk_hi)
						emit_record(E, pos, pid);
				}
			}
		}
		/* past the chunk: stop once the longest open prefix began behind it */
		if (pos >= chunk_hi) {
			const uint64_t d = pos - chunk_hi + 1;       /* symbols read beyond the chunk */
			if (state < __ldg(&A.level_start[d + 1 <= (uint64_t)A.max_len ? d + 1 : (uint64_t)A.max_len]))
				break;
		}
	}
}

/*
 * STRIDE 4: 4-byte windows at multiples of 4 (every pattern >= 7 bytes).
 * STRIDE 8: 3-byte windows at multiples of 8 (every pattern >= 10 bytes): half the bitmap
 *           lookups per input byte; the exact stage still keys on the 4 bytes at the window.
 */
template <int STRIDE>
__global__ void __launch_bounds__(S4_THREADS, 1)
k_scan_sampled(const __grid_constant__ AutDev A, const __grid_constant__ EmitCtx E,
    const uint8_t *__restrict__ data, uint64_t n,
    uint64_t vec_lo, uint64_t vec_hi, uint64_t limit, uint32_t *work_counter)
{
	extern __shared__ __align__(128) uint32_t s4_smem[];
	uint32_t *f1 = s4_smem;
	uint32_t *f2 = s4_smem + F1_WORDS;
	uint64_t *bar = reinterpret_cast<uint64_t *>(s4_smem + F1_WORDS + F2_WORDS);
	const int lane = threadIdx.x & 31;
	uint32_t trace_chunks = 0;
	constexpr int WPV = 16 / STRIDE;                 /* windows per 16-byte vector */
	constexpr int NW = S4_UNROLL * WPV;              /* windows per lane per chunk */

	pdl_trigger();            /* k_resolve_queue's CTAs follow on every SM as its scanning CTA ends */
	if (E.trace && threadIdx.x == 0)
		E.trace[blockIdx.x * 4 + 0] = globaltimer_ns();
	/* stage both bitmaps with TMA bulk copies, 32 KiB each */
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		mbar_expect_tx(bar, (F1_WORDS + F2_WORDS) * 4);
		for (uint32_t off = 0; off < F1_WORDS; off += 8192)
			bulk_g2s(f1 + off, A.f1 + off, 32768, bar);
		for (uint32_t off = 0; off < F2_WORDS; off += 8192)
			bulk_g2s(f2 + off, A.f2 + off, 32768, bar);
	}
	/* the bitmaps belong to the automaton; everything below -- work counter, text, queues -- may
	 * have been written by what precedes this kernel in the stream */
	pdl_wait();
	__syncthreads();

	/*
	 * Work distribution.  The range is cut into chunks of 32 lanes x S4_UNROLL vectors
	 * (2 KiB; a lane reads vectors chunk_base + u * 32 + lane, so each load instruction
	 * covers 512 contiguous bytes).  Two levels:
	 *   - the CTA takes BLOCKS of 32 runs x L chunks from the one global counter (L = 8, i.e.
	 *     512 KiB, while plenty is left; 4, 2, 1 towards the end so that all SMs -- they differ
	 *     by +-5 % in speed -- finish together): ~2 300 global atomics per GiB;
	 *   - its warps take RUNS from the block by tickets from a shared-memory counter, one run
	 *     ahead of use.  Ticket t is run t % 32 of the CTA's block t / 32; the warp that draws
	 *     run 0 of block k fetches block k + 2 and publishes it in slot (k + 2) % 4.
	 * The whole grid sweeps the stream front to back, 148 x 512 KiB at a time.  (One global
	 * atomic per 16 KiB run -- the first version -- cost 13 % of the kernel: 65 536
	 * same-address atomics per GiB keep one L2 slice busy a third of the time and every
	 * warp's loads pass through it; loads alone ran at 5.46 TB/s with 16 KiB runs and at
	 * 6.32 TB/s with 64 KiB runs.  Static per-CTA ranges with stealing ran at 3.7 TB/s: 148
	 * separate streams are much worse for DRAM than one front.)
	 */
	const uint64_t chunk_vecs = 32ull * S4_UNROLL;
	const uint32_t n_chunks = (uint32_t)((vec_hi - vec_lo + chunk_vecs - 1) / chunk_vecs);
	uint32_t *s_ticket = reinterpret_cast<uint32_t *>(bar + 2);
	/* slot: {seq = block number + 1 once published, first chunk, L, readers} */
	volatile uint32_t *s_slot = reinterpret_cast<volatile uint32_t *>(bar + 4);
	auto fetch_block = [&](uint32_t &first_chunk, uint32_t &len) {
		const uint32_t seen = *reinterpret_cast<volatile uint32_t *>(work_counter);
		const uint32_t left = seen < n_chunks ? n_chunks - seen : 0;
		const uint32_t full = S4_BLOCK_RUNS * S4_UNIT_CHUNKS * gridDim.x;   /* one full block per CTA */
		len = S4_UNIT_CHUNKS;
		while (len > 1 && left < 2 * (full / S4_UNIT_CHUNKS) * len)
			len >>= 1;
		first_chunk = atomicAdd(work_counter, S4_BLOCK_RUNS * len);
	};
	if (threadIdx.x == 32) {
		*s_ticket = 0;
		for (uint32_t j = 0; j < S4_SLOTS; ++j) {
			uint32_t fc = 0, len = 0;
			if (j < 2)
				fetch_block(fc, len);
			s_slot[4 * j + 1] = fc;
			s_slot[4 * j + 2] = len;
			s_slot[4 * j + 3] = j < 2 ? 0 : S4_BLOCK_RUNS;       /* unused slots count as fully read */
			s_slot[4 * j + 0] = j < 2 ? j + 1 : 0;
		}
	}
	__syncthreads();
	/* lane 0 only: a run is (first chunk, number of chunks); count 0 = nothing left */
	auto grab = [&](uint32_t &start, uint32_t &count) {
		start = 0;
		count = 0;
		if (lane == 0) {
			const uint32_t t = atomicAdd(s_ticket, 1u);
			const uint32_t k = t / S4_BLOCK_RUNS, r = t % S4_BLOCK_RUNS;
			if (r == 0) {
				/* block k + 2 goes where block k - 2 was: wait until its 32 runs have been read */
				const uint32_t j = (k + 2) % S4_SLOTS;
				while (s_slot[4 * j + 3] < S4_BLOCK_RUNS)
					;
				uint32_t fc, len;
				fetch_block(fc, len);
				s_slot[4 * j + 1] = fc;
				s_slot[4 * j + 2] = len;
				s_slot[4 * j + 3] = 0;
				__threadfence_block();
				s_slot[4 * j + 0] = k + 3;
			}
			const uint32_t j = k % S4_SLOTS;
			while (s_slot[4 * j + 0] != k + 1)
				;
			__threadfence_block();
			const uint32_t fc = s_slot[4 * j + 1], len = s_slot[4 * j + 2];
			atomicAdd(const_cast<uint32_t *>(&s_slot[4 * j + 3]), 1u);
			if (S4_INTERLEAVE) {
				start = fc + r;
				if (start < n_chunks) {
					const uint32_t fit = (n_chunks - start + S4_BLOCK_RUNS - 1) / S4_BLOCK_RUNS;
					count = fit < len ? fit : len;
				}
			} else {
				start = fc + r * len;
				if (start < n_chunks)
					count = n_chunks - start < len ? n_chunks - start : len;
			}
		}
	};
	constexpr uint64_t run_step = (S4_INTERLEAVE ? S4_BLOCK_RUNS : 1) * 32ull * S4_UNROLL;   /* vectors between chunks of a run */
	uint32_t run_start, next_start;
	uint32_t run_count, next_count;
	grab(run_start, run_count);
	grab(next_start, next_count);
	run_start = __shfl_sync(FULL_MASK, run_start, 0);
	run_count = __shfl_sync(FULL_MASK, run_count, 0);

	/*
	 * Latency hiding without a register double buffer (the slow path below needs the
	 * registers): while a chunk is processed, the next one is pulled into L2 with one bulk
	 * prefetch per warp, so the loads at the top of the next iteration are L2 hits, and the
	 * seven other warps of the scheduler cover those.
	 */
	uint4 v[S4_UNROLL];
	uint64_t first = vec_lo + run_start * chunk_vecs;        /* first vector of the next chunk */
	/* lane 0: pull the chunk d positions ahead in this warp's sequence into L2 (d = 1 is `first`;
	 * run_count chunks are left in the current run, `first` included, then comes the next run) */
	auto prefetch_ahead = [&](uint32_t d) {
		if (d <= run_count)
			prefetch_l2_bulk(data + (first + (d - 1) * run_step) * 16, (uint32_t)chunk_vecs * 16);
		else if (d - 1 - run_count < next_count)
			prefetch_l2_bulk(data + (vec_lo + next_start * chunk_vecs + (d - 1 - run_count) * run_step) * 16,
			    (uint32_t)chunk_vecs * 16);
	};
	if (run_count && lane == 0) {
		for (uint32_t d = 1; d <= S4_PF_DIST; ++d)
			prefetch_ahead(d);
	}
	mbar_wait(bar, 0);
	if (E.trace && threadIdx.x == 0)
		E.trace[blockIdx.x * 4 + 1] = globaltimer_ns();

	const uint32_t vq_region = blockIdx.x * (S4_THREADS / 32) + (threadIdx.x >> 5);
	uint4 *const vq_mine = E.vq + (size_t)vq_region * E.vq_cap;
	uint32_t vq_n = 0;
	uint32_t *const dq_mine = E.dq + (size_t)vq_region * E.dq_cap;
	uint32_t dq_n = 0;                                /* lane 0 only */
	auto load_chunk = [&](uint4 (&d)[S4_UNROLL], uint64_t f) {
		if (f + chunk_vecs <= vec_hi) {                  /* all but the last chunk */
			const uint4 *p = reinterpret_cast<const uint4 *>(data) + f + lane;
#pragma unroll
			for (int u = 0; u < S4_UNROLL; ++u)
				d[u] = __ldcs(p + u * 32);
		} else {
#pragma unroll
			for (int u = 0; u < S4_UNROLL; ++u) {
				const uint64_t idx = f + (uint64_t)u * 32 + lane;
				d[u] = (idx < vec_hi) ? load_vec(data, idx) : make_uint4(0, 0, 0, 0);
			}
		}
	};
	/* `first` moves on to the chunk after the one just loaded; its successor is prefetched into L2 */
	auto advance = [&]() {
		if (--run_count == 0) {
			run_start = __shfl_sync(FULL_MASK, next_start, 0);
			run_count = __shfl_sync(FULL_MASK, next_count, 0);
			if (run_count)
				grab(next_start, next_count);
			first = vec_lo + run_start * chunk_vecs;
		} else {
			first += run_step;
		}
		if (run_count && lane == 0)
			prefetch_ahead(S4_PF_DIST);
	};
	while (run_count) {
		++trace_chunks;
		const uint64_t cur_first = first;
		load_chunk(v, cur_first);
		advance();
		uint32_t hits = 0;
#pragma unroll
		for (int u = 0; u < S4_UNROLL; ++u) {
			const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
			for (int k = 0; k < WPV; ++k) {
				const uint32_t h = w[k * (4 / WPV)] * ACM_HASH1_MUL;
				const uint32_t word = f1[h >> (32 - (ACM_F1_BITS_LOG2 - 5))];
				/* bitmap words are bit-reversed: a tested bit lands in the MSB.  Two bits per
				 * gram (blocked Bloom filter in the one word fetched): 3 more ALU ops per
				 * window buy a ~5x lower hit rate, i.e. fewer rounds of the slow loop below */
				const uint32_t t = __funnelshift_l(0u, word, h) & __funnelshift_l(0u, word, h >> 12);
				hits = __funnelshift_l(t, hits, 1);
			}
		}
		/* the last chunk may stick out past vec_hi: those vectors were not loaded (zeros) and
		 * must not be tested -- an all-zero window is a real, and popular, pattern gram */
		if (cur_first + chunk_vecs > vec_hi) {
#pragma unroll
			for (int u = 0; u < S4_UNROLL; ++u)
				if (cur_first + (uint64_t)u * 32 + lane >= vec_hi)
					hits &= ~((((1u << WPV) - 1u) << (NW - WPV)) >> (WPV * u));
		}
		if (__reduce_add_sync(FULL_MASK, (uint32_t)__popc(hits)) >= NW * 32 / 4) {
			/* dense: queue the chunk for k_resolve_queue (a full list: walk it here, one lane, slowly) */
			if (lane == 0) {
				if (dq_n < E.dq_cap) {
					if (dq_n == 0)
						E.overflow[DQ_ANY_WORD] = 1u;
					dq_mine[dq_n++] = (uint32_t)(cur_first - vec_lo);
				}
				else
					s4_dense_chunk(&A, &E, data, cur_first * 16, cur_first * 16 + chunk_vecs * 16, limit, STRIDE);
			}
			continue;
		}
		/*
		 * bit (NW - 1 - q) of hits belongs to window q = u * WPV + k.  Level 2 (second hash, own
		 * bitmap of 96 KiB, three bits per gram in the one word fetched) is tested lane-locally first: a
		 * short divergent loop, one iteration per level-1 survivor of the lane (random data:
		 * ~5 per chunk over the whole warp, 1.1 iterations).  It leaves < 0.1 windows per chunk,
		 * so most chunks never enter the warp-uniform loop below and its L2 round trips.
		 */
		{
			uint32_t hh = hits;
			hits = 0;
			while (hh) {
				const uint32_t p = 31u - (uint32_t)__clz(hh);
				const uint32_t bit = 1u << p;
				hh ^= bit;
				const uint32_t q = NW - 1 - p;
				const uint32_t u = q / WPV, k = q % WPV;
				const uint4 x = (u & 2) ? ((u & 1) ? v[3] : v[2]) : ((u & 1) ? v[1] : v[0]);
				uint32_t word4;                          /* the 4 bytes at the window */
				if (STRIDE == 4)
					word4 = (k & 2) ? ((k & 1) ? x.w : x.z) : ((k & 1) ? x.y : x.x);
				else
					word4 = k ? x.z : x.x;
				const uint32_t h2 = word4 * ACM_HASH2_MUL;
				const uint32_t word2 = f2[__umulhi(h2, F2_WORDS)];
				if ((int32_t)(__funnelshift_l(0u, word2, h2) & __funnelshift_l(0u, word2, h2 >> 12) &
				    __funnelshift_l(0u, word2, h2 >> 6)) < 0)
					hits |= bit;
			}
		}
		/* Survivors are rare: from here on control flow is warp-uniform and verification is done
		 * by the whole warp. */
		while (__any_sync(FULL_MASK, hits != 0)) {
			uint32_t cbegin = 0, word4 = 0;
			uint64_t e = 0;
			const bool has = hits != 0;
			if (has) {
				const uint32_t p = 31u - (uint32_t)__clz(hits);
				hits ^= 1u << p;
				const uint32_t q = NW - 1 - p;
				const uint32_t u = q / WPV, k = q % WPV;
				e = (cur_first + (uint64_t)u * 32 + lane) * 16 + (uint64_t)k * STRIDE;
				const uint4 x = (u & 2) ? ((u & 1) ? v[3] : v[2]) : ((u & 1) ? v[1] : v[0]);
				if (STRIDE == 4)
					word4 = (k & 2) ? ((k & 1) ? x.w : x.z) : ((k & 1) ? x.y : x.x);
				else
					word4 = k ? x.z : x.x;
			}
			/*
			 * Deferred resolution.  What a surviving window still needs -- exact-table probe,
			 * candidate list, bytes at the window, full compare, bucket counter -- is a chain of
			 * five to six dependent L2 / DRAM round trips that this warp would sit through (2.3 us
			 * of warp time per real match: 25 % of the whole kernel at one match per 10 KiB; the
			 * stream loads of a stalled warp are not in flight).  So the window is only WRITTEN
			 * here -- a plain store into this warp's own region, no atomic, nothing to wait for
			 * -- and k_resolve_queue does the rest for all windows at once.  A full region falls
			 * back to the inline path below.
			 */
			{
				const uint32_t qm = __ballot_sync(FULL_MASK, has);
				if (vq_n + (uint32_t)__popc(qm) <= E.vq_cap) {
					if (has)
						vq_mine[vq_n + (uint32_t)__popc(qm & lanemask_lt())] =
						    make_uint4((uint32_t)e, (uint32_t)(e >> 32), word4, 0u);
					vq_n += (uint32_t)__popc(qm);
					continue;
				}
			}
			if (has) {
				/* exact table keyed by the 4 bytes at the window (L2 resident) */
				uint32_t sl = (word4 * ACM_HASH3_MUL) >> A.gram_shift;
				for (;;) {
					const uint2 slot = __ldg(reinterpret_cast<const uint2 *>(A.grams) + 2 * (size_t)sl);
					if (slot.y == 0)
						break;
					if (slot.x == word4) {
						cbegin = slot.y;
						break;
					}
					sl = (sl + 1) & A.gram_mask;
				}
			}
			uint32_t pend = __ballot_sync(FULL_MASK, cbegin != 0);
			while (pend) {
				const int src = __ffs(pend) - 1;
				pend &= pend - 1;
				uint32_t ci = __shfl_sync(FULL_MASK, cbegin, src) - 1;
				const uint64_t ew = __shfl_sync(FULL_MASK, e, src);
				/* the 8 bytes of text at the window: every candidate of this gram is compared on its
				 * own bytes o .. o+7 (it starts at ew - o) */
				const uint32_t *wp = reinterpret_cast<const uint32_t *>(data + ew);
				const uint32_t t0 = __ldg(wp), t1 = (ew + 4 < n) ? __ldg(wp + 1) : 0u;
				for (;;) {
					/* one candidate per lane (lists are padded, lanes past LAST are ignored) */
					const uint4 *cp = reinterpret_cast<const uint4 *>(A.cand) + 2 * (size_t)(ci + lane);
					const uint4 c = __ldg(cp);
					const uint32_t lastm = __ballot_sync(FULL_MASK, (c.w & ACM_CAND_LAST) != 0);
					const int nvalid = lastm ? __ffs(lastm) : 32;
					const uint32_t o = c.x >> ACM_CAND_O_SHIFT;
					const uint32_t len = c.w & ~ACM_CAND_LAST;
					const uint32_t rem = len - o;                  /* pattern bytes from the window on, >= 3 */
					const uint32_t m0 = rem >= 4 ? 0xffffffffu : 0x00ffffffu;
					const uint32_t m1 = rem >= 8 ? 0xffffffffu : (rem <= 4 ? 0u : ((1u << (8 * (rem - 4))) - 1u));
					const uint64_t s = ew - o;
					bool ok = lane < nvalid && ew >= o && s >= E.valid_lo && s + len <= limit &&
					    ((t0 ^ c.y) & m0) == 0 && ((t1 ^ c.z) & m1) == 0;
					if (ok) {
						/* the pattern's last 4 bytes against the text: repetitive text makes many
						 * candidates share the bytes at the window, very few also share their end */
						const uint64_t ta = s + len - 4;
						const uint32_t *tp = reinterpret_cast<const uint32_t *>(data + (ta & ~3ull));
						const uint32_t sh = (uint32_t)(ta & 3) * 8;
						const uint32_t a0 = __ldg(tp);
						const uint32_t a1 = sh ? __ldg(tp + 1) : 0u;
						ok = __funnelshift_r(a0, a1, sh) == __ldg(reinterpret_cast<const uint32_t *>(cp + 1));
					}
					uint32_t surv = __ballot_sync(FULL_MASK, ok);
					while (surv) {
						const int l = __ffs(surv) - 1;
						surv &= surv - 1;
						s4_verify(A, E, data, __shfl_sync(FULL_MASK, c.x, l) & ACM_CAND_ID_MASK,
						    __shfl_sync(FULL_MASK, len, l), __shfl_sync(FULL_MASK, s, l), lane);
					}
					if (lastm)
						break;
					ci += 32;
				}
			}
		}
	}
	if (lane == 0) {
		E.vq_count[vq_region] = vq_n;
		E.dq_count[vq_region] = dq_n;
	}
	if (E.trace && lane == 0) {
		atomicMax((unsigned long long *)&E.trace[blockIdx.x * 4 + 2], (unsigned long long)globaltimer_ns());
		atomicAdd((unsigned long long *)&E.trace[blockIdx.x * 4 + 3], (unsigned long long)trace_chunks);
	}
}

/*
 * Everything the filter survivors of k_scan_sampled still need, as a pipeline of CTA-wide
 * phases so that every dependent L2 / DRAM round trip is taken once for all the windows of a
 * region (and, over the grid, for all ~2 x 10^5 queued windows of a GiB) instead of stalling
 * one streaming warp at a time:
 *   1a  one lane per queued window: exact-table probe (L2) with the text bytes 4..7 after the
 *       window fetched alongside (DRAM); a hit posts its candidates as work items;
 *   1b  one lane per candidate: record load (L2), compare on the bytes at the window; the
 *       survivors -- on anything but repetitive text, real occurrences -- go to a list;
 *   2   four lanes per listed candidate: full compare (lane j takes bytes [16 j + 64 r,
 *       16 j + 64 r + 16) in round r, all loads of a round issued together; 90 % of the ClamAV
 *       signatures are one round, 187 bytes are three), then the record is emitted.
 * One CTA per region (= per scanning warp).  Queue entry: {window offset lo, hi, the 4 bytes at
 * the window, -}.
 */
#define RQ_THREADS 128
#ifndef RQ_MINB
#define RQ_MINB 12                          /* the usual instantiation: 40 registers, no spills, 12 CTAs per SM (10 / 11 / 12 / 16:
                                             * 1 GiB step 0.2038 / 0.2011 / 0.2014 / 0.2035 ms -- at 16 the 32 registers spill) */
#endif
#define RQ_MINB_DENSE 10                    /* <= 51 registers: the (rare, out-of-line) dense-chunk walk may spill, the resolve phases must not grow */
#define RQ_WORK    512                      /* candidate work items per CTA round */
#define RQ_LIST    (2 * RQ_THREADS)         /* candidates awaiting the full compare: one batch of 1b + overflow of 1a */

struct RqShared {
	uint4    list[RQ_LIST];                 /* {start lo, start hi | pattern << 8, length, offset in pat_blob} */
	uint2    work[RQ_WORK];                 /* {candidate index, source lane} */
	uint64_t e[RQ_THREADS];                 /* per source lane: window offset */
	uint2    t[RQ_THREADS];                 /*                  text bytes 0..3, 4..7 at the window */
	uint32_t n_work, n_list;
};

/* one lane, whole pattern: only when the shared list is full (the work list overflowed into it) */
__device__ __noinline__ bool rq_compare_lane(const uint8_t *__restrict__ data, const uint8_t *__restrict__ pat,
    uint64_t s, uint32_t len)
{
	const uint32_t *pw = reinterpret_cast<const uint32_t *>(pat);
	const uint32_t *tw = reinterpret_cast<const uint32_t *>(data + (s & ~3ull));
	const uint32_t sh = (uint32_t)(s & 3) * 8;
	for (uint32_t k = 0; k < len; k += 4) {
		const uint32_t rem = len - k;
		const uint32_t w0 = __ldg(tw + (k >> 2));
		const uint32_t w1 = ((uint32_t)(s & 3) + (rem < 4 ? rem : 4u) > 4u) ? __ldg(tw + (k >> 2) + 1) : 0u;
		const uint32_t mask = rem >= 4 ? 0xffffffffu : ((1u << (8 * rem)) - 1u);
		if ((__funnelshift_r(w0, w1, sh) ^ __ldg(pw + (k >> 2))) & mask)
			return false;
	}
	return true;
}

/* phase 1b for one candidate: compare on the bytes at the window, then list it */
__device__ __forceinline__ bool rq_candidate(const AutDev &A, const EmitCtx &E, RqShared &S,
    const uint8_t *__restrict__ data, uint64_t limit, uint32_t ci, uint64_t e, uint32_t t0, uint32_t t1)
{
	const uint4 *cp = reinterpret_cast<const uint4 *>(A.cand) + 2 * (size_t)ci;
	const uint4 c = __ldg(cp);
	const uint4 c2 = __ldg(cp + 1);                      /* .y = offset of the pattern bytes */
	const uint32_t o = c.x >> ACM_CAND_O_SHIFT;
	const uint32_t len = c.w & ~ACM_CAND_LAST;
	const uint32_t rem = len - o;                        /* pattern bytes from the window on, >= 3 */
	const uint32_t m0 = rem >= 4 ? 0xffffffffu : 0x00ffffffu;
	const uint32_t m1 = rem >= 8 ? 0xffffffffu : (rem <= 4 ? 0u : ((1u << (8 * (rem - 4))) - 1u));
	const uint64_t s = e - o;
	if (e >= o && s >= E.valid_lo && s + len <= limit && ((t0 ^ c.y) & m0) == 0 && ((t1 ^ c.z) & m1) == 0) {
		const uint32_t pid = c.x & ACM_CAND_ID_MASK;
		const uint32_t i = atomicAdd(&S.n_list, 1u);
		if (i < RQ_LIST)
			S.list[i] = make_uint4((uint32_t)s, (uint32_t)(s >> 32) | (pid << 8), len, c2.y);
		else if (rq_compare_lane(data, A.pat_blob + c2.y, s, len))
			emit_record(E, s + len - 1, pid);
	}
	return (c.w & ACM_CAND_LAST) != 0;
}

/* DENSE: this kernel also walks the region's dense chunks (automata without a row-displaced table;
 * the usual instantiation carries none of that code and none of its register pressure) */
template <bool DENSE>
__global__ void __launch_bounds__(RQ_THREADS, DENSE ? RQ_MINB_DENSE : RQ_MINB)
k_resolve_queue(const __grid_constant__ AutDev A, const __grid_constant__ EmitCtx E,
    const uint8_t *__restrict__ data, uint64_t n, uint64_t limit, uint64_t vec_lo, uint32_t stride)
{
	pdl_wait();               /* the queues of k_scan_sampled */
	/* k_dense_walk, launched behind this kernel as its programmatic dependent, needs nothing from it,
	 * only k_scan_sampled's lists -- complete once every CTA here is past its wait: its CTAs may take
	 * the SMs this grid's last wave leaves */
	pdl_trigger();
	/* the region's dense chunks first (usually none), when k_dense_walk cannot take them (no
	 * row-displaced table): one thread per slice, table entries through L1 */
	if (DENSE) {
		const uint32_t nd = min(E.dq_count[blockIdx.x], E.dq_cap);
		const uint32_t *dl = E.dq + (size_t)blockIdx.x * E.dq_cap;
		const uint32_t per = nd >= 96 ? 1u : (nd >= 48 ? 2u : (nd >= 24 ? 4u : 8u));   /* slices per chunk */
		const uint32_t slice = 32u * S4_UNROLL * 16u / per;
		for (uint32_t i = threadIdx.x; i < nd * per; i += RQ_THREADS) {
			const uint64_t lo = (vec_lo + dl[i / per]) * 16 + (uint64_t)(i % per) * slice;
			s4_dense_chunk(&A, &E, data, lo, lo + slice, limit, stride);
		}
	}
	__shared__ RqShared S;
	const uint4 *q = E.vq + (size_t)blockIdx.x * E.vq_cap;
	/* the first entry is fetched before the count is known (the region exists either way) */
	uint4 ent = threadIdx.x < E.vq_cap ? q[threadIdx.x] : make_uint4(0, 0, 0, 0);
	const uint32_t count = E.vq_count[blockIdx.x];

	for (uint32_t base = 0; base < count; base += RQ_THREADS) {      /* CTA-uniform */
		if (threadIdx.x == 0) {
			S.n_work = 0;
			S.n_list = 0;
		}
		__syncthreads();
		/* ---- 1a: probe ---- */
		const uint32_t slot = base + threadIdx.x;
		if (slot < count) {
			if (base)
				ent = q[slot];
			const uint64_t e = (uint64_t)ent.x | ((uint64_t)ent.y << 32);
			const uint32_t word4 = ent.z;
			const uint32_t t1 = (e + 4 < n) ? __ldg(reinterpret_cast<const uint32_t *>(data + e) + 1) : 0u;
			uint32_t sl = (word4 * ACM_HASH3_MUL) >> A.gram_shift;
			uint32_t cbegin = 0, cnt = 0;
			for (;;) {
				const uint4 g = __ldg(reinterpret_cast<const uint4 *>(A.grams) + sl);
				if (g.y == 0)
					break;
				if (g.x == word4) {
					cbegin = g.y;
					cnt = g.z;
					break;
				}
				sl = (sl + 1) & A.gram_mask;
			}
			if (cbegin) {
				S.e[threadIdx.x] = e;
				S.t[threadIdx.x] = make_uint2(word4, t1);
				const uint32_t w0 = atomicAdd(&S.n_work, cnt);
				for (uint32_t i = 0; i < cnt; ++i) {
					if (w0 + i < RQ_WORK)
						S.work[w0 + i] = make_uint2(cbegin - 1 + i, threadIdx.x);
					else
						rq_candidate(A, E, S, data, limit, cbegin - 1 + i, e, word4, t1);   /* work list full */
				}
			}
		}
		__syncthreads();
		/* ---- 1b + 2, one batch of RQ_THREADS candidates at a time (the list cannot overflow) ---- */
		const uint32_t nw = S.n_work < RQ_WORK ? S.n_work : RQ_WORK;
		const uint32_t j = threadIdx.x & 3;
		for (uint32_t wb = 0; wb == 0 || wb < nw; wb += RQ_THREADS) {  /* CTA-uniform */
			if (wb + threadIdx.x < nw) {
				const uint2 it = S.work[wb + threadIdx.x];
				const uint2 t = S.t[it.y];
				rq_candidate(A, E, S, data, limit, it.x, S.e[it.y], t.x, t.y);
			}
			__syncthreads();
			const uint32_t m = S.n_list < RQ_LIST ? S.n_list : RQ_LIST;
			for (uint32_t i0 = 0; i0 < m; i0 += RQ_THREADS / 4) {     /* CTA-uniform */
				const uint32_t i = i0 + (threadIdx.x >> 2);
				const bool live = i < m;
				const uint4 c = live ? S.list[i] : make_uint4(0, 0, 0, 0);
				const uint64_t s = (uint64_t)c.x | ((uint64_t)(c.y & 0xffu) << 32);
				const uint32_t pid = c.y >> 8, len = c.z;
				const uint32_t *pw = reinterpret_cast<const uint32_t *>(A.pat_blob + c.w);
				const uint32_t *tw = reinterpret_cast<const uint32_t *>(data + (s & ~3ull));
				const uint32_t sh = (uint32_t)(s & 3) * 8;
				uint32_t diff = 0;
				for (uint32_t r = 0; __any_sync(FULL_MASK, live && r < len && diff == 0); r += 64) {
					const uint32_t k0 = r + 16 * j;                    /* this lane's 16 pattern bytes */
					if (live && k0 < len) {
						const uint32_t nv = len - k0 < 16u ? len - k0 : 16u;       /* valid bytes, 1..16 */
						const uint32_t nwords = (nv + 3) >> 2;
						/* aligned text words: one more than pattern words when the last bytes reach into it */
						const uint32_t need = nwords + (((uint32_t)(s & 3) + (nv - 4 * (nwords - 1)) > 4u) ? 1u : 0u);
						const uint32_t *tp = tw + (k0 >> 2), *pp = pw + (k0 >> 2);
						uint32_t w[5], p[4];
#pragma unroll
						for (int i = 0; i < 5; ++i)
							w[i] = (uint32_t)i < need ? __ldg(tp + i) : 0u;
#pragma unroll
						for (int i = 0; i < 4; ++i)
							p[i] = (uint32_t)i < nwords ? __ldg(pp + i) : 0u;
#pragma unroll
						for (int i = 0; i < 4; ++i) {
							const uint32_t left = nv > 4u * i ? nv - 4u * i : 0u;      /* valid bytes from word i on */
							const uint32_t mask = left >= 4 ? 0xffffffffu : ((1u << (8 * left)) - 1u);
							diff |= (__funnelshift_r(w[i], w[i + 1], sh) ^ p[i]) & mask;
						}
					}
					/* a mismatch anywhere in the quad ends the candidate */
					diff |= __shfl_xor_sync(FULL_MASK, diff, 1);
					diff |= __shfl_xor_sync(FULL_MASK, diff, 2);
				}
				if (live && j == 0 && diff == 0)
					emit_record(E, s + len - 1, pid);
			}
			__syncthreads();
			if (threadIdx.x == 0)
				S.n_list = 0;
			__syncthreads();
		}
		__syncthreads();
	}
}

/* ------------------------------------------------------------------------- */
/* exact 2-byte start filter                                                 */
/* ------------------------------------------------------------------------- */

#define S2_THREADS 512
#ifndef S2_UNROLL
#define S2_UNROLL  4
#endif
#ifndef S2_PREFETCH
#define S2_PREFETCH 1              /* pull the CTA's next tile into L2 while this one is processed */
#endif
#define S2_SMEM_BYTES(pair) (((pair) ? 65536 / 8 : ACM_B3_WORDS * 4) + 16 + (S2_THREADS / 32) * (WQ_CAP * 8 + 16))

/*
 * PAIR = true:  the filter is an exact 2^16-bit bitmap of (byte, next byte) -- 8 KiB; right when few
 *               patterns stand behind it (the short patterns of a mixed set: hardly any position passes);
 * PAIR = false: a 2^19-bit blocked Bloom bitmap (three bits per key, all in one word) of the first
 *               THREE bytes -- 64 KiB; with thousands of patterns the pair bitmap lets 3 .. 15 % of all
 *               positions through and every one of them costs a trie walk of dependent L2 reads.
 */
template <bool PAIR>
__global__ void __launch_bounds__(S2_THREADS, 2)
k_scan_start2(const AutDev A, const EmitCtx E, const uint8_t *__restrict__ data, uint64_t n,
    uint64_t vec_lo, uint64_t vec_hi, uint64_t limit, const uint32_t *__restrict__ start_filter,
    uint32_t max_depth)
{
	constexpr uint32_t FILTER_BYTES = PAIR ? 65536 / 8 : ACM_B3_WORDS * 4;
	extern __shared__ __align__(128) uint32_t s2_smem[];
	uint32_t *b3 = s2_smem;
	uint64_t *bar = reinterpret_cast<uint64_t *>(s2_smem + FILTER_BYTES / 4);
	WarpQueue Q;
	{
		uint8_t *qbase = reinterpret_cast<uint8_t *>(s2_smem) + FILTER_BYTES + 16 +
		    (threadIdx.x >> 5) * (WQ_CAP * 8 + 16);
		Q.rec = reinterpret_cast<uint64_t *>(qbase + 16);
		Q.count = reinterpret_cast<uint32_t *>(qbase);
		if ((threadIdx.x & 31) == 0)
			*Q.count = 0;
	}

	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		mbar_expect_tx(bar, FILTER_BYTES);
		if (PAIR) {
			bulk_g2s(b3, start_filter, 8192, bar);
		} else {
			bulk_g2s(b3, start_filter, 32768, bar);
			bulk_g2s(b3 + 8192, start_filter + 8192, 32768, bar);
		}
	}
	__syncthreads();
	mbar_wait(bar, 0);

	const uint64_t tile_vecs = (uint64_t)S2_THREADS * S2_UNROLL;
	const int lane = threadIdx.x & 31;

	for (uint64_t first = vec_lo + (uint64_t)blockIdx.x * tile_vecs; first < vec_hi;
	     first += (uint64_t)gridDim.x * tile_vecs) {
		/* all loads of the tile first: S2_UNROLL x 16 bytes in flight per thread */
		uint4 vv[S2_UNROLL];
#pragma unroll
		for (int u = 0; u < S2_UNROLL; ++u) {
			const uint64_t idx = first + (uint64_t)u * S2_THREADS + threadIdx.x;
			vv[u] = idx < vec_hi ? load_vec(data, idx) : make_uint4(0, 0, 0, 0);
		}
#if S2_PREFETCH
		if (lane < S2_UNROLL) {
			/* lane u: the 512 bytes this warp reads as part u of the CTA's next tile */
			const uint64_t nidx = first + (uint64_t)gridDim.x * tile_vecs + (uint64_t)lane * S2_THREADS +
			    (threadIdx.x & ~31u);
			if ((nidx + 32) * 16 <= n)
				prefetch_l2_bulk(data + nidx * 16, 512);
		}
#endif
#pragma unroll
		for (int u = 0; u < S2_UNROLL; ++u) {
			const uint64_t idx = first + (uint64_t)u * S2_THREADS + threadIdx.x;
			/* whole warps fall off the end together except in the last tile */
			const bool live = idx < vec_hi;
			const uint4 v = vv[u];
			/* first word of the next vector: the neighbour lane has it */
			uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
			if (lane == 31) {
				nxt = 0;
				if (live) {
					const uint64_t b = (idx + 1) * 16;
					if (b + 4 <= n)
						nxt = __ldg(reinterpret_cast<const uint32_t *>(data + b));
					else
						for (int k = 0; k < 4; ++k)
							if (b + k < n)
								nxt |= (uint32_t)data[b + k] << (8 * k);
				}
			}
			const uint32_t w[5] = {v.x, v.y, v.z, v.w, nxt};
			uint32_t hits = 0;
#pragma unroll
			for (int p = 0; p < 16; ++p) {
				/* low 24 bits: bytes p, p+1, p+2 (whatever lies past the end of the buffer is
				 * harmless: patterns of one or two bytes are in the bitmap with every completion) */
				const uint32_t x = __funnelshift_r(w[p >> 2], w[(p >> 2) + 1], 8 * (p & 3));
				if (PAIR) {
					/* low 16 bits: bytes p, p+1; words are bit-reversed (tested bit -> MSB) */
					const uint32_t word = b3[(x >> 5) & 0x7FF];
					const uint32_t t = __funnelshift_l(0u, word, x);
					hits = __funnelshift_l(t, hits, 1);
				} else {
					const uint32_t h = (x & 0x00FFFFFFu) * ACM_HASH2_MUL;
					const uint32_t word = b3[h >> 18];
					const uint32_t t = (word >> (h & 31)) & (word >> ((h >> 5) & 31)) & (word >> ((h >> 10) & 31)) & 1u;
					hits = (hits << 1) | t;
				}
			}
			if (!live)
				hits = 0;
			/* warp-uniform loop: every lane pops one of its hits per round and walks the
			 * automaton from there; records are staged per warp and flushed together */
			while (__any_sync(FULL_MASK, hits != 0)) {
				if (hits) {
					const int p = 31 - __clz(hits);
					hits &= ~(1u << p);
					const uint64_t s = idx * 16 + (uint64_t)(15 - p);
					if (s >= E.valid_lo && s < limit)
						walk_from_t<true>(A, E, Q, data, s, limit, max_depth);
				}
				__syncwarp();
				if (*Q.count >= WQ_FLUSH)
					flush_queue(Q, E, lane);
			}
		}
	}
	flush_queue(Q, E, lane);
}

/* ------------------------------------------------------------------------- */
/* plain DFA walk, one thread per chunk, leading halo                        */
/* ------------------------------------------------------------------------- */

/*
 * One transition of reference ahomatch.cl:56-65 (AC_ushorts/ahomatch.cl:37-66 for ushort
 * symbols): a symbol outside the alphabet sends the automaton back to the root; the full match
 * list of the target = the own lists along the output links (acsmx.c:417-429).
 */
__device__ __forceinline__ void dfa_step(const AutDev &A, const EmitCtx &E, uint32_t alpha, uint32_t &state,
    uint32_t c, uint64_t pos, uint64_t a)
{
	if (c >= alpha) {
		state = 0;
		return;
	}
	const uint32_t e = __ldg(&A.T[(size_t)state * alpha + c]);
	state = e & ACM_T_MASK;
	if ((e & ACM_T_ANY) && pos >= a) {
		for (uint32_t v = state; v; v = __ldg(&A.olink[v]))
			emit_own(A, E, v, pos);
	}
}

/*
 * One thread per chunk, cold start Lmax-1 symbols early.  The walk is a chain of
 * dependent table reads (L1 for the root and depth-1 rows, L2 beyond), so what the kernel needs is
 * many chains in flight: the host cuts the scan into as many chunks as the GPU holds threads
 * (scan_dfa_chunk), and the symbols arrive 16 bytes at a time (one sector per lane per two loads)
 * instead of one load per symbol.
 */
#ifndef DFA_MINB
#define DFA_MINB 6                 /* resident CTAs of 256 per SM (40 registers) */
#endif
template <typename SYM>
__global__ void __launch_bounds__(256, DFA_MINB)
k_scan_dfa(const AutDev A, const EmitCtx E, const SYM *__restrict__ data, uint64_t n, uint64_t chunk,
    uint64_t nthreads, uint32_t *final_state)
{
	constexpr uint64_t VS = 16 / sizeof(SYM);          /* symbols per 16-byte vector */
	constexpr uint32_t PER_WORD = 4 / sizeof(SYM);
	constexpr uint32_t SYM_BITS = 8 * sizeof(SYM);
	constexpr uint32_t SYM_MASK = (1u << SYM_BITS) - 1u;
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t alpha = (uint32_t)A.alpha;
	const uint64_t limit = E.emit_hi < n ? E.emit_hi : n;

	if (t == nthreads) {
		/* the true state after the last symbol: the longest suffix that is a
		 * pattern prefix fits in the last Lmax symbols */
		const uint64_t back = (uint64_t)A.max_len;
		uint32_t state = 0;
		uint64_t p0 = n > back ? n - back : 0;
		if (p0 < E.valid_lo)
			p0 = E.valid_lo;
		for (uint64_t pos = p0; pos < n; ++pos) {
			const uint32_t c = data[pos];
			state = (c < alpha) ? (__ldg(&A.T[(size_t)state * alpha + c]) & ACM_T_MASK) : 0u;
		}
		*final_state = state;
		return;
	}
	if (t > nthreads)
		return;

	const uint64_t a = E.emit_lo + t * chunk;
	uint64_t b = a + chunk;
	if (b > limit)
		b = limit;
	if (a >= b)
		return;
	const uint64_t halo = A.max_len > 0 ? (uint64_t)(A.max_len - 1) : 0;
	uint32_t state = 0;
	uint64_t pos = a > halo ? a - halo : 0;
	if (pos < E.valid_lo)
		pos = E.valid_lo;

	/* head: symbol by symbol up to the next 16-byte boundary (all of it when the buffer itself
	 * is not 16-byte aligned) */
	uint64_t head_end = (pos + VS - 1) / VS * VS;
	if (head_end > b || ((uintptr_t)data & 15) != 0)
		head_end = b;
	for (; pos < head_end; ++pos)
		dfa_step(A, E, alpha, state, data[pos], pos, a);
	/* body: whole vectors */
	for (; pos + VS <= b; pos += VS) {
		const uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos));
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (uint32_t k = 0; k < (uint32_t)VS; ++k)
			dfa_step(A, E, alpha, state, (w[k / PER_WORD] >> ((k % PER_WORD) * SYM_BITS)) & SYM_MASK,
			    pos + k, a);
	}
	for (; pos < b; ++pos)
		dfa_step(A, E, alpha, state, data[pos], pos, a);
}

/* ------------------------------------------------------------------------- */
/* the DFA walk out of the row-displaced table: hot rows in shared memory     */
/* ------------------------------------------------------------------------- */

/*
 * k_scan_xd -- reference ahomatch.cl:56-65 (AC_ushorts/ahomatch.cl:37-66) in the form north_star
 * names: one thread per chunk, cold start Lmax-1 symbols early, "the hot rows of the transition
 * table in shared memory, cold states served from L2".
 *
 * The table is acm_core.c:build_xd's: the whole DFA as ONE array of 4-byte entries, root row
 * first, then the states in breadth-first order (9.5 MB for 10 000 ClamAV signatures where the
 * dense table is 370 MiB; 0.9 MB where the reference's ushort table is 528 MiB).  Its first
 * `smem_slots` entries -- root, depth 1, depth 2, ... as far as 128 KiB reach -- are staged into
 * shared memory with TMA bulk copies; a lookup whose slot lies beyond comes from L1 / L2.
 *
 * Transition on symbol c from the state with base o, b the previous symbol:
 *     x = xd[o + c]                       is it this state's entry for c?  (carries c, bases unique)
 *     g = xd[base(xd[b]) + c]             else the entry of the depth-1 state of b
 *     r = xd[c]                           else the root row
 * g and r depend on the input only -- r is next step's xd[b] -- so they are off the walk's
 * dependency chain; when the walk sits in that depth-1 state itself (random input: most of the
 * time) x IS g and the lane skips the third lookup.
 * A symbol outside the alphabet sends the automaton to the root (cli/b200_flow_grep separates
 * flows that way).  Matches: the entry's flag says "the target reports something"; the state's
 * breadth-first id (xd_sid) leads to the own lists along the output links as in k_scan_dfa.
 */
#define XD_THREADS 1024
#define XD_SMEM_MAX (224 * 1024)
#define XD_SMEM_DEFAULT (128 * 1024)

struct XdLook {
	uint32_t tab_sa;          /* shared address of the staged prefix                  */
	uint32_t smem_slots;      /* >= xd_d1_end: the rows of all states of depth <= 1 are in it */
	const uint32_t *tab;      /* the whole table in global memory                     */
	uint32_t sym_mask, base_shift;
};

/* the rare part of a step, out of line: every pattern the state with base `os` reports */
__device__ __noinline__ void xd_report(const AutDev *__restrict__ Ap, const EmitCtx *__restrict__ Ep, uint32_t os,
    uint64_t pos)
{
	for (uint32_t v = __ldg(Ap->xd_sid + os); v; v = __ldg(&Ap->olink[v]))
		emit_own(*Ap, *Ep, v, pos);
}

/*
 * The same for a slice of a dense chunk of the sampled filter (k_dense_walk).  The CHUNK owns the
 * occurrences whose INDEXED window lies in it (s4_dense_chunk tells why); among the chunk's slices an
 * occurrence belongs to the one it STARTS in, those that start in front of the chunk to slice 0 -- so
 * only slice 0 needs the cold start in front of it.  lo / len: the slice; chunk_off = lo - chunk start.
 */
__device__ __noinline__ void xd_report_win(const AutDev *__restrict__ Ap, const EmitCtx *__restrict__ Ep, uint32_t os,
    uint64_t pos, uint64_t lo, uint32_t len, uint32_t stride, uint32_t chunk_off)
{
	const AutDev &A = *Ap;
	const uint64_t chunk_lo = lo - chunk_off, chunk_hi = chunk_lo + 32u * S4_UNROLL * 16u;
	for (uint32_t v = __ldg(A.xd_sid + os); v; v = __ldg(&A.olink[v])) {
		const uint32_t b = __ldg(&A.own_begin[v]), t = __ldg(&A.own_begin[v + 1]);
		for (uint32_t k = b; k < t; ++k) {
			const uint32_t pid = __ldg(&A.own_pat[k]);
			const uint64_t plen = __ldg(&A.pat_len[pid]);
			/* mixed sets: the short patterns belong to the start-filter pass */
			if (plen < A.split_len || pos + 1 < plen)
				continue;
			const uint64_t s = pos + 1 - plen;
			if (s >= lo + len || (s < lo && chunk_off != 0))
				continue;
			const uint64_t w = s + __ldg(&A.pat_win[(size_t)pid * 8 + ((stride - (uint32_t)(s % stride)) % stride)]);
			if (s >= Ep->valid_lo && w >= chunk_lo && w < chunk_hi)
				emit_record(*Ep, pos, pid);
		}
	}
}

struct XdState {
	uint32_t os;              /* base of the current state                            */
	uint32_t ob;              /* base of the depth-1 state of the previous symbol     */
};

/*
 * One step, branch-free up to the (rare) report: r and g always come from shared memory, the
 * state's own entry x from shared memory or, beyond the staged prefix, from L1 / L2 -- two
 * predicated loads, of which at most one executes, none when the walk sits in the depth-1 state
 * of the previous symbol (x is g then).
 */
template <bool WIN = false>
__device__ __forceinline__ bool xd_step(const AutDev &A, const EmitCtx &E, const XdLook &L, XdState &S, uint32_t c,
    uint64_t pos, uint64_t a, uint32_t win_len = 0, uint32_t stride = 0, uint32_t chunk_off = 0)
{
	uint32_t r, g, x;
	asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(L.tab_sa + c * 4));
	asm("ld.shared.u32 %0, [%1];" : "=r"(g) : "r"(L.tab_sa + (S.ob + c) * 4));
	const uint32_t slot = S.os + c;
	x = g;
	asm("{\n"
	    ".reg .pred ps, pg;\n"
	    ".reg .b32 sa;\n"
	    ".reg .b64 ga;\n"
	    "setp.ne.u32 ps, %1, %2;\n"                 /* the state is not that depth-1 state     */
	    "setp.lt.and.u32 pg, %3, %4, ps;\n"         /* ... and its slot is in shared memory    */
	    "setp.ge.and.u32 ps, %3, %4, ps;\n"         /* ... or beyond it                        */
	    "mad.lo.u32 sa, %3, 4, %5;\n"
	    "mad.wide.u32 ga, %3, 4, %6;\n"
	    "@pg ld.shared.u32 %0, [sa];\n"
	    "@ps ld.global.nc.u32 %0, [ga];\n"
	    "}\n" : "+r"(x) : "r"(S.os), "r"(S.ob), "r"(slot), "r"(L.smem_slots), "r"(L.tab_sa), "l"(L.tab));
	const bool okx = (x & L.sym_mask) == c, okg = (g & L.sym_mask) == c;
	const uint32_t nx = okx ? x : (okg ? g : r);
	S.os = nx >> L.base_shift;
	S.ob = r >> L.base_shift;
	const bool any = ((nx >> (L.base_shift - 1)) & 1u) != 0;
	if (any && pos >= a) {
		if (WIN)
			xd_report_win(&A, &E, S.os, pos, a, win_len, stride, chunk_off);
		else
			xd_report(&A, &E, S.os, pos);
	}
	return any;
}

template <typename SYM>
__global__ void __launch_bounds__(XD_THREADS, 1)
k_scan_xd(const __grid_constant__ AutDev A, const __grid_constant__ EmitCtx E, const SYM *__restrict__ data,
    uint64_t n, uint64_t chunk, uint64_t nthreads, uint32_t *final_state, uint32_t smem_slots)
{
	extern __shared__ __align__(128) uint32_t xd_smem[];
	constexpr uint64_t VS = 16 / sizeof(SYM);          /* symbols per 16-byte vector */
	constexpr uint32_t PER_WORD = 4 / sizeof(SYM);
	constexpr uint32_t SYM_BITS = 8 * sizeof(SYM);
	constexpr uint32_t SYM_MASK = (1u << SYM_BITS) - 1u;
	uint64_t *bar = reinterpret_cast<uint64_t *>(xd_smem + smem_slots);
	const uint32_t alpha = (uint32_t)A.alpha;
	const uint64_t limit = E.emit_hi < n ? E.emit_hi : n;

	if (threadIdx.x == 0) {
		const uint32_t bytes = smem_slots * 4;
		mbar_init(bar, 1);
		mbar_expect_tx(bar, bytes);
		for (uint32_t off = 0; off < bytes; off += 16384)
			bulk_g2s(reinterpret_cast<uint8_t *>(xd_smem) + off, reinterpret_cast<const uint8_t *>(A.xd_tab) + off,
			    min(bytes - off, 16384u), bar);
	}
	__syncthreads();
	mbar_wait(bar, 0);

	XdLook L;
	L.tab_sa = smem_u32(xd_smem);
	L.smem_slots = smem_slots;
	L.tab = A.xd_tab;
	L.sym_mask = (1u << A.xd_sym_bits) - 1u;
	L.base_shift = A.xd_sym_bits + 1;
	const uint64_t halo = A.max_len > 0 ? (uint64_t)(A.max_len - 1) : 0;

	/* a symbol outside the alphabet (ushort streams only: flow separators) sends the walk to the root */
	auto step = [&](XdState &S, uint32_t c, uint64_t p, uint64_t a) {
		if (sizeof(SYM) > 1 && c >= alpha)
			S.os = S.ob = 0;
		else
			xd_step(A, E, L, S, c, p, a);
	};
	/*
	 * Every thread walks TWO chunks at a time, vector by vector in lockstep: two independent
	 * dependency chains per lane (a deep state's entry may come from L2).  Chunk pairs are taken
	 * round-robin; u == npairs is the extra job "state after the last symbol".
	 */
	const uint64_t npairs = (nthreads + 1) / 2;
	for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u <= npairs;
	     u += (uint64_t)gridDim.x * blockDim.x) {
		if (u == npairs) {
			/* the true state after the last symbol (breadth-first id): the longest suffix that is a
			 * pattern prefix fits in the last Lmax symbols */
			const uint64_t back = (uint64_t)A.max_len;
			uint64_t p0 = n > back ? n - back : 0;
			if (p0 < E.valid_lo)
				p0 = E.valid_lo;
			XdState S = {0u, 0u};
			for (uint64_t pos = p0; pos < n; ++pos)
				step(S, data[pos], pos, ~0ull);
			*final_state = __ldg(A.xd_sid + S.os);
			continue;
		}
		uint64_t a[2], b[2], pos[2];
		XdState S[2] = {{0u, 0u}, {0u, 0u}};
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			a[q] = E.emit_lo + (2 * u + q) * chunk;
			b[q] = a[q] + chunk;
			if (b[q] > limit)
				b[q] = limit;
			if (a[q] > limit)
				a[q] = limit;                           /* empty */
			pos[q] = a[q] > halo ? a[q] - halo : 0;
			if (pos[q] < E.valid_lo)
				pos[q] = E.valid_lo;
			if (a[q] >= b[q])
				pos[q] = b[q];
			/* head: symbol by symbol up to the next 16-byte boundary (all of it when the buffer itself
			 * is not 16-byte aligned) */
			uint64_t head_end = (pos[q] + VS - 1) / VS * VS;
			if (head_end > b[q] || ((uintptr_t)data & 15) != 0)
				head_end = b[q];
			for (; pos[q] < head_end; ++pos[q])
				step(S[q], data[pos[q]], pos[q], a[q]);
		}
		/* body: whole vectors of both chunks in lockstep */
		while (pos[0] + VS <= b[0] && pos[1] + VS <= b[1]) {
			const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(data + pos[0]));
			const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(data + pos[1]));
			const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
			for (uint32_t k = 0; k < (uint32_t)VS; ++k) {
				step(S[0], (w0[k / PER_WORD] >> ((k % PER_WORD) * SYM_BITS)) & SYM_MASK, pos[0] + k, a[0]);
				step(S[1], (w1[k / PER_WORD] >> ((k % PER_WORD) * SYM_BITS)) & SYM_MASK, pos[1] + k, a[1]);
			}
			pos[0] += VS;
			pos[1] += VS;
		}
		/* what is left of either (the other one ended): vectors, then the tail */
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			for (; pos[q] + VS <= b[q]; pos[q] += VS) {
				const uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos[q]));
				const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (uint32_t k = 0; k < (uint32_t)VS; ++k)
					step(S[q], (w[k / PER_WORD] >> ((k % PER_WORD) * SYM_BITS)) & SYM_MASK, pos[q] + k, a[q]);
			}
			for (; pos[q] < b[q]; ++pos[q])
				step(S[q], data[pos[q]], pos[q], a[q]);
		}
	}
}

/*
 * k_dense_walk -- the dense chunks k_scan_sampled queued (per scanning warp: E.dq; see s4_dense_chunk),
 * walked out of the row-displaced table with its hot rows in shared memory, the way k_scan_xd walks a
 * whole buffer.
 *
 * The slices of all queued chunks form one list that the grid shares out thread by thread.  No dense
 * chunk -- random-looking input, the usual case: a CTA is done after reading one word (flags[10], set by
 * the scanning warp that queues its first chunk), before anything is staged.
 * The kernel is launched as the programmatic dependent of k_resolve_queue (it needs k_scan_sampled's
 * lists only, and those are complete before k_resolve_queue starts), so its launch and the empty
 * check hide behind that kernel's last wave; griddepcontrol.wait before the exit keeps "this grid
 * done" meaning "k_resolve_queue done" for what follows in the stream.
 * Otherwise the T chunks are cut into `per` slices each, 1 <= per <= 16, so that every thread has up to
 * two walks in lockstep (a handful of zero pages in a buffer must not become a handful of 2 KiB walks on
 * one lane each: with six warps on an SM every step is a bare latency chain).  The chunk owns the
 * occurrences whose indexed window lies in it; a slice reports those of them that START in it, the
 * chunk's first slice also those that start in front of the chunk.  So a walk starts at its slice --
 * the first one max_win bytes early: nothing that starts earlier can have its window in the chunk --
 * and runs on behind it until no occurrence that began inside can still be open.
 */
__global__ void __launch_bounds__(XD_THREADS, 1)
k_dense_walk(const __grid_constant__ AutDev A, const __grid_constant__ EmitCtx E, const uint8_t *__restrict__ data,
    uint64_t limit, uint64_t vec_lo, uint32_t stride, uint32_t nreg, uint32_t smem_slots)
{
	extern __shared__ __align__(128) uint32_t xd_smem[];
	__shared__ uint32_t s_warp[XD_THREADS / 32];
	/* nothing queued anywhere (random-looking input, the usual case): one word read, done */
	pdl_trigger();
	if (((volatile uint32_t *)E.overflow)[DQ_ANY_WORD] == 0) {
		pdl_wait();
		return;
	}
	uint64_t *bar = reinterpret_cast<uint64_t *>(xd_smem + smem_slots);
	uint32_t *pre = reinterpret_cast<uint32_t *>(bar + 2);     /* [nreg + 1]: chunks queued in the regions before r */
	if (threadIdx.x == 0) {
		const uint32_t bytes = smem_slots * 4;
		mbar_init(bar, 1);
		mbar_expect_tx(bar, bytes);
		for (uint32_t off = 0; off < bytes; off += 16384)
			bulk_g2s(reinterpret_cast<uint8_t *>(xd_smem) + off, reinterpret_cast<const uint8_t *>(A.xd_tab) + off,
			    min(bytes - off, 16384u), bar);
	}
	/* exclusive prefix sum over ALL regions' counts, the same in every CTA: the slices of all dense chunks
	 * are one list that the whole grid shares out, thread by thread (a few dense chunks in a buffer --
	 * real files: padding, a zero page here and there -- used to be walked by the six warps of the CTA
	 * whose scanning twin had met them while the other SMs had nothing to do) */
	{
		const uint32_t per_t = (nreg + XD_THREADS - 1) / XD_THREADS;
		const uint32_t r0 = threadIdx.x * per_t;
		uint32_t sum = 0;
		for (uint32_t k = 0; k < per_t; ++k)
			if (r0 + k < nreg)
				sum += min(E.dq_count[r0 + k], E.dq_cap);
		uint32_t inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(FULL_MASK, inc, d);
			if ((int)(threadIdx.x & 31) >= d)
				inc += t;
		}
		if ((threadIdx.x & 31) == 31)
			s_warp[threadIdx.x >> 5] = inc;
		__syncthreads();
		uint32_t run = inc - sum;
		for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w)
			run += s_warp[w];
		for (uint32_t k = 0; k < per_t; ++k)
			if (r0 + k < nreg) {
				pre[r0 + k] = run;
				run += min(E.dq_count[r0 + k], E.dq_cap);
			}
		if (r0 < nreg && r0 + per_t >= nreg)
			pre[nreg] = run;
	}
	__syncthreads();
	const uint32_t T = pre[nreg];
	mbar_wait(bar, 0);

	XdLook L;
	L.tab_sa = smem_u32(xd_smem);
	L.smem_slots = smem_slots;
	L.tab = A.xd_tab;
	L.sym_mask = (1u << A.xd_sym_bits) - 1u;
	L.base_shift = A.xd_sym_bits + 1;

	constexpr uint32_t CHUNK = 32u * S4_UNROLL * 16u;
	uint32_t per_log = 0;
	const uint32_t G = gridDim.x * XD_THREADS;       /* walkers in the grid, two walks each at a time */
	/* (more, smaller slices than the walkers need were measured and lose: at least 4 / 8 / 16 per chunk
	 * took zero pages from 1 190 to 1 030 / 870 GB/s and repeated signature heads from 320 to 280 / 230) */
	while (per_log < 4 && ((uint64_t)T << per_log) < 2ull * G)
		++per_log;
	const uint32_t slice = CHUNK >> per_log;
	const uint64_t items = (uint64_t)T << per_log;

	/* consecutive threads take consecutive slices (neighbours share lines of text and of the table) */
	for (uint64_t u = (uint64_t)blockIdx.x * XD_THREADS + threadIdx.x; u < items; u += 2ull * G) {
		uint64_t a[2], b[2], pos[2];
		uint32_t co[2] = {0u, 0u};                       /* slice start - chunk start */
		XdState S[2] = {{0u, 0u}, {0u, 0u}};
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const uint64_t item = u + (uint64_t)q * G;
			a[q] = b[q] = pos[q] = 0;
			if (item >= items)
				continue;
			const uint32_t ci = (uint32_t)(item >> per_log);
			/* the region r with pre[r] <= ci < pre[r + 1] */
			uint32_t r = 0, hi = nreg;
			while (hi - r > 1) {
				const uint32_t mid = (r + hi) >> 1;
				if (pre[mid] <= ci)
					r = mid;
				else
					hi = mid;
			}
			const uint32_t first = E.dq[(size_t)r * E.dq_cap + (ci - pre[r])];
			co[q] = ((uint32_t)item & ((1u << per_log) - 1u)) * slice;
			a[q] = (vec_lo + first) * 16 + co[q];
			b[q] = a[q] + slice < limit ? a[q] + slice : limit;
			if (a[q] >= b[q]) {
				a[q] = b[q] = 0;
				continue;
			}
			/* the chunk's first slice also owns what starts in front of the chunk: cold start there */
			pos[q] = a[q];
			if (co[q] == 0)
				pos[q] = a[q] > A.max_win ? a[q] - A.max_win : 0;
			if (pos[q] < E.valid_lo)
				pos[q] = E.valid_lo;
			/* head: byte by byte up to the next 16-byte boundary */
			uint64_t head_end = (pos[q] + 15) & ~15ull;
			if (head_end > b[q])
				head_end = b[q];
			for (; pos[q] < head_end; ++pos[q])
				xd_step<true>(A, E, L, S[q], data[pos[q]], pos[q], a[q], slice, stride, co[q]);
		}
		/*
		 * body: whole vectors of both slices in lockstep.  Sixteen equal bytes (zero pages, padding,
		 * erased flash: most of what makes a chunk dense) whose first step leaves the walk where it
		 * was, in a state that reports nothing: the other fifteen steps would do the same -- the
		 * transition depends on (state, byte) only -- and are skipped.
		 */
		auto constant = [](const uint4 &v) {
			const uint32_t rep = (v.x & 0xFFu) * 0x01010101u;
			return v.x == rep && v.y == rep && v.z == rep && v.w == rep;
		};
		while (pos[0] + 16 <= b[0] && pos[1] + 16 <= b[1]) {
			const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(data + pos[0]));
			const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(data + pos[1]));
			const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w};
			const XdState P0 = S[0], P1 = S[1];
			const bool r0 = xd_step<true>(A, E, L, S[0], w0[0] & 0xFFu, pos[0], a[0], slice, stride, co[0]);
			const bool r1 = xd_step<true>(A, E, L, S[1], w1[0] & 0xFFu, pos[1], a[1], slice, stride, co[1]);
			const bool skip0 = !r0 && S[0].os == P0.os && S[0].ob == P0.ob && constant(v0);
			const bool skip1 = !r1 && S[1].os == P1.os && S[1].ob == P1.ob && constant(v1);
			if (!(skip0 && skip1)) {
#pragma unroll
				for (uint32_t k = 1; k < 16; ++k) {
					if (!skip0)
						xd_step<true>(A, E, L, S[0], (w0[k / 4] >> ((k % 4) * 8)) & 0xFFu, pos[0] + k, a[0], slice, stride, co[0]);
					if (!skip1)
						xd_step<true>(A, E, L, S[1], (w1[k / 4] >> ((k % 4) * 8)) & 0xFFu, pos[1] + k, a[1], slice, stride, co[1]);
				}
			}
			pos[0] += 16;
			pos[1] += 16;
		}
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			for (; pos[q] + 16 <= b[q]; pos[q] += 16) {
				const uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos[q]));
				const uint32_t w[4] = {v.x, v.y, v.z, v.w};
				const XdState P = S[q];
				if (!xd_step<true>(A, E, L, S[q], w[0] & 0xFFu, pos[q], a[q], slice, stride, co[q]) && S[q].os == P.os &&
				    S[q].ob == P.ob && constant(v))
					continue;
#pragma unroll
				for (uint32_t k = 1; k < 16; ++k)
					xd_step<true>(A, E, L, S[q], (w[k / 4] >> ((k % 4) * 8)) & 0xFFu, pos[q] + k, a[q], slice, stride, co[q]);
			}
			for (; pos[q] < b[q]; ++pos[q])
				xd_step<true>(A, E, L, S[q], data[pos[q]], pos[q], a[q], slice, stride, co[q]);
			/* behind the slice: until the longest open prefix began behind it (d bytes read beyond the
			 * slice, state shallower than d + 1).  The test costs two dependent global loads (state id,
			 * level table), so it is made once per 16-byte vector, not per byte: running up to 15 steps
			 * too far reports nothing the slice does not own.  Inside a run of equal bytes the state sits
			 * at a fixed point as deep as the longest signature prefix made of that byte -- a hundred
			 * and more zeros -- and the run is skipped a vector at a time here as well. */
			if (a[q] < b[q]) {
				const uint64_t end = b[q] + (uint64_t)A.max_len < limit ? b[q] + (uint64_t)A.max_len : limit;
				while (pos[q] < end) {
					if ((pos[q] & 15) == 0 && pos[q] + 16 <= end) {
						const uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos[q]));
						const uint32_t w[4] = {v.x, v.y, v.z, v.w};
						const XdState P = S[q];
						const bool r = xd_step<true>(A, E, L, S[q], w[0] & 0xFFu, pos[q], a[q], slice, stride, co[q]);
						if (r || S[q].os != P.os || S[q].ob != P.ob || !constant(v)) {
#pragma unroll
							for (uint32_t k = 1; k < 16; ++k)
								xd_step<true>(A, E, L, S[q], (w[k / 4] >> ((k % 4) * 8)) & 0xFFu, pos[q] + k, a[q], slice, stride, co[q]);
						}
						pos[q] += 16;
					} else {
						xd_step<true>(A, E, L, S[q], data[pos[q]], pos[q], a[q], slice, stride, co[q]);
						++pos[q];
						if ((pos[q] & 15) != 0 && pos[q] < end)
							continue;
					}
					const uint64_t d = pos[q] - b[q];
					if (__ldg(A.xd_sid + S[q].os) < __ldg(&A.level_start[d + 1 <= (uint64_t)A.max_len ? d + 1 : (uint64_t)A.max_len]))
						break;
				}
			}
		}
	}
	pdl_wait();
}

/* ------------------------------------------------------------------------- */
/* class-compressed DFA: 16-bit entries, table in shared memory              */
/* ------------------------------------------------------------------------- */

/*
 * k_scan_cdfa -- dense-output walk for the automata k_scan_rd (k1_rd.cuh) cannot take: more than
 * 31 byte classes, pattern bytes not in one range, or a row-displaced table that does not fit
 * shared memory.
 *
 * The plain DFA walk of reference ahomatch.cl:56-65 -- one transition per byte, one thread per
 * chunk, cold start Lmax-1 bytes early:
 *   - the table is class-compressed to uint16[states][C] (acm_core.c build_cdfa); the first
 *     n_hot rows (breadth-first order: the shallow, most visited states) live in shared memory,
 *     the rest is served by L1/L2;
 *   - byte -> column is arithmetic when the pattern bytes span < 64 values, else one lookup in
 *     a 32-way replicated (bank-conflict-free) table;
 *   - every thread walks TWO adjacent chunks in lockstep, two independent dependency chains;
 *   - a chunk is its own result bucket: the thread owns the slot counter (no atomics), hits
 *     come out in end-offset order because the walk is sequential, and every state's full
 *     match list is stored sorted by pattern index, so the post-pass is an expanding copy,
 *     not a sort.
 * Chunks are cut on absolute multiples of 2^shift in the buffer, so interior chunks are
 * 16-byte aligned and read with 16-byte loads; the first and last chunk of a scan (cut by
 * emit_lo / the end) take a byte-wise path.
 */
#define CD_THREADS   1024
#define CD_LUT_WORDS 2048          /* 64 words x 32 banks */
#define CD_SMEM_MAX  (227 * 1024)

/* predicated 4-byte store: no branch, nothing happens when p is false */
__device__ __forceinline__ void st_pred_u32(uint32_t *addr, uint32_t v, bool p)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "setp.ne.b32 p, %2, 0;\n"
	    "@p st.global.u32 [%0], %1;\n"
	    "}\n" ::"l"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}

/* the tables are read-only once staged: plain (non-volatile) asm so loads can be scheduled freely */
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
	uint16_t v;
	asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
	return v;
}

struct CdLook {
	uint32_t tab_sa;          /* shared address: the hot rows                             */
	uint32_t lut_sa;          /* shared address: replicated class map (!RANGE)            */
	const uint16_t *tab;      /* the whole table in global memory                         */
	uint32_t C, n_hot, rlo, cmax, lane4;
};

template <bool RANGE>
__device__ __forceinline__ uint32_t cd_class(const CdLook &L, uint32_t b)
{
	if (RANGE)
		return min(b - L.rlo, L.cmax);            /* below lo wraps to huge -> cmax */
	uint32_t w;
	asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(L.lut_sa + ((b & 0xFCu) << 5) + L.lane4));
	return (w >> ((b & 3u) * 8u)) & 0xFFu;
}

__device__ __forceinline__ uint32_t cd_next(const CdLook &L, uint32_t state, uint32_t c)
{
	const uint32_t idx = state * L.C + c;
	return state < L.n_hot ? lds_u16(L.tab_sa + idx * 2) : (uint32_t)__ldg(L.tab + idx);
}

/*
 * Emission.  Looking up a state's pattern list inside the walk would put two dependent L2
 * round trips on the warp's critical path at almost every step (with one match per ~9 bytes
 * nearly every warp step has a lane that matched): measured 60 % of all stall samples.  So the
 * walk stores only the HIT, (offset in chunk << 14) | state, a fire-and-forget 4-byte store (an
 * 8-byte key needs a register pair per store, and the pair cannot be rewritten until the store
 * has left the LSU queue: those write-after-read waits were 30 % of all stall samples), and
 * adds the 2-bit list length carried by the transition entry to its record count; the
 * expansion state -> pattern indices happens in the compaction kernel
 * (k_bucket_expand_compact), which is a bandwidth-bound copy anyway.  Bucket row (32-bit
 * words): slot 0 = number of hits, slots 1.. = hits.  Direct mode (second pass after a bucket
 * overflow) expands in place instead.
 */
struct CdOut {
	uint64_t *dst;      /* direct mode: records                                        */
	uint32_t *row;      /* bucket mode: hits                                           */
	uint32_t  k;        /* hits (bucket mode) written so far                            */
	uint32_t  nrec;     /* records this chunk produces                                 */
	uint32_t  cap;      /* 0 = direct mode                                             */
};

__device__ __forceinline__ void cd_emit(const AutDev &A, CdOut &o, uint32_t state, uint32_t code, uint64_t pos,
    uint32_t off_in_chunk)
{
	uint32_t cnt = code;
	if (code == 3 || o.cap == 0) {
		const uint32_t fb = __ldg(&A.cd_flat_begin[state]);
		cnt = __ldg(&A.cd_flat_begin[state + 1]) - fb;
		if (o.cap == 0)
			for (uint32_t j = 0; j < cnt; ++j)
				o.dst[o.nrec + j] = (pos << ACM_KEY_PAT_BITS) | __ldg(&A.cd_flat_pat[fb + j]);
	}
	if (o.cap) {
		++o.k;
		if (o.k < o.cap)
			o.row[o.k] = (off_in_chunk << ACM_CD_STATE_BITS) | state;
	}
	o.nrec += cnt;
}

__device__ __forceinline__ CdOut cd_open(const EmitCtx &E, uint64_t b)
{
	CdOut o;
	o.k = 0;
	o.nrec = 0;
	o.dst = nullptr;
	o.row = nullptr;
	if (E.direct) {
		o.dst = E.out + E.offsets[b];
		o.cap = 0;
	} else {
		o.row = reinterpret_cast<uint32_t *>(E.buckets) + b * E.cap;
		o.cap = E.cap;
	}
	return o;
}

__device__ __forceinline__ void cd_close(const EmitCtx &E, uint64_t b, const CdOut &o)
{
	E.counts[b] = o.nrec;
	if (o.cap) {
		o.row[0] = o.k;
		if (o.k >= o.cap)
			*E.overflow = 1u;
	}
}

/* byte-wise walk of chunk k (absolute chunk index): the first / last chunk of a scan, direct mode */
template <bool RANGE>
__device__ __noinline__ void cd_chunk_bytes(const AutDev *__restrict__ Ap, const EmitCtx *__restrict__ Ep,
    const CdLook *__restrict__ Lp, const uint8_t *__restrict__ data, uint64_t k, uint64_t limit)
{
	const AutDev &A = *Ap;
	const EmitCtx &E = *Ep;
	const CdLook L = *Lp;
	uint64_t lo = k << E.shift, hi = (k + 1) << E.shift;
	if (lo < E.emit_lo)
		lo = E.emit_lo;
	if (hi > limit)
		hi = limit;
	const uint64_t halo = A.max_len > 0 ? (uint64_t)(A.max_len - 1) : 0;
	uint64_t pos = lo > halo ? lo - halo : 0;
	if (pos < E.valid_lo)
		pos = E.valid_lo;
	const uint64_t b = k - (E.emit_lo >> E.shift);
	CdOut o = cd_open(E, b);
	uint32_t state = 0;
	for (; pos < hi; ++pos) {
		const uint32_t e = cd_next(L, state, cd_class<RANGE>(L, __ldg(data + pos)));
		state = e & ACM_CD_STATE_MASK;
		if ((e >> ACM_CD_STATE_BITS) && pos >= lo)
			cd_emit(A, o, state, e >> ACM_CD_STATE_BITS, pos, (uint32_t)(pos - (k << E.shift)));
	}
	cd_close(E, b, o);
}

/*
 * Shared memory: [tab_bytes: hot rows][class map (!RANGE)][mbarrier]; all sizes multiples of 16.
 */
template <bool RANGE>
__global__ void __launch_bounds__(CD_THREADS, 1)
k_scan_cdfa(const __grid_constant__ AutDev A, const __grid_constant__ EmitCtx E,
    const uint8_t *__restrict__ data, uint64_t limit, uint32_t n_hot, uint32_t tab_bytes)
{
	extern __shared__ __align__(128) uint32_t cd_smem[];
	uint8_t *sm = reinterpret_cast<uint8_t *>(cd_smem);
	uint32_t *lut = reinterpret_cast<uint32_t *>(sm + tab_bytes);
	uint64_t *bar = reinterpret_cast<uint64_t *>(sm + tab_bytes + (RANGE ? 0 : CD_LUT_WORDS * 4));

	/* stage the tables: TMA bulk copies of up to 16 KiB on one mbarrier */
	if (threadIdx.x == 0) {
		const uint8_t *src_tab = reinterpret_cast<const uint8_t *>(A.cd_tab);
		mbar_init(bar, 1);
		mbar_expect_tx(bar, tab_bytes);
		for (uint32_t off = 0; off < tab_bytes; off += 16384)
			bulk_g2s(sm + off, src_tab + off, min(tab_bytes - off, 16384u), bar);
	}
	if (!RANGE) {
		/* word (b >> 2) of the class map, once per bank */
		const uint32_t *cls32 = reinterpret_cast<const uint32_t *>(A.cd_cls);
		for (uint32_t i = threadIdx.x; i < CD_LUT_WORDS; i += CD_THREADS)
			lut[i] = __ldg(cls32 + (i >> 5));
	}
	__syncthreads();
	mbar_wait(bar, 0);

	CdLook L;
	L.tab_sa = smem_u32(sm);
	L.lut_sa = smem_u32(lut);
	L.tab = A.cd_tab;
	L.C = A.cd_classes;
	L.n_hot = n_hot;
	L.rlo = (uint32_t)A.cd_range_lo;
	L.cmax = A.cd_classes - 1;
	L.lane4 = (threadIdx.x & 31) * 4;

	const uint32_t shift = E.shift;
	const uint64_t chunk = 1ull << shift;
	const uint64_t halo = A.max_len > 0 ? (uint64_t)(A.max_len - 1) : 0;
	const uint64_t k_first = E.emit_lo >> shift, k_last = (limit - 1) >> shift;
	const uint64_t n_pairs = (k_last - k_first + 2) / 2;

	for (uint64_t g = (uint64_t)blockIdx.x * CD_THREADS + threadIdx.x; g < n_pairs;
	     g += (uint64_t)gridDim.x * CD_THREADS) {
		const uint64_t ka = k_first + 2 * g, kb = ka + 1;
		const uint64_t a0 = ka << shift;
		/* both chunks whole, aligned, with their halo inside the valid range */
		const bool fast = !E.direct && kb <= k_last && a0 >= E.emit_lo && a0 + 2 * chunk <= limit &&
		    a0 >= halo && a0 - halo >= E.valid_lo;
		if (!fast) {
			cd_chunk_bytes<RANGE>(&A, &E, &L, data, ka, limit);
			if (kb <= k_last)
				cd_chunk_bytes<RANGE>(&A, &E, &L, data, kb, limit);
			continue;
		}
		/*
		 * Bucket mode, branch-free: with one match per ~9 bytes nearly every warp step has a
		 * lane that matched, so an `if (hit)` body would run (divergently) at every step.
		 * Instead every lane executes a predicated 8-byte store of its hit and two adds.
		 */
		const uint64_t ba = ka - k_first;
		uint32_t *rowa = reinterpret_cast<uint32_t *>(E.buckets) + ba * E.cap + 1, *rowb = rowa + E.cap;   /* slot 0 = hit count */
		/* keep the row pointers as pointers: the store address is then one IMAD.WIDE of the hit count */
		asm volatile("" : "+l"(rowa), "+l"(rowb));
		const uint32_t capm1 = E.cap - 1, thr4 = A.cd_thr4;
		uint32_t hka = 0, hkb = 0, nra = 0, nrb = 0;       /* hits, records */
		uint32_t sa = 0, sb = 0;
		const uint4 *pa = reinterpret_cast<const uint4 *>(data + a0);
		const uint4 *pb = reinterpret_cast<const uint4 *>(data + a0 + chunk);
		/*
		 * Cold start over the halo: the 16-byte vectors that cover [a0 - halo, a0) (for chunk b:
		 * the end of chunk a).  Bytes in front of a0 - halo are walked too and then forgotten
		 * (state reset at the first halo byte), which is the same as starting there.
		 */
		{
			const uint32_t hv = (uint32_t)((halo + 15) >> 4), skip = hv * 16 - (uint32_t)halo;
			for (uint32_t v = 0; v < hv; ++v) {
				const uint4 xa = __ldg(pa - hv + v), xb = __ldg(pb - hv + v);
				const uint32_t wa[4] = {xa.x, xa.y, xa.z, xa.w}, wb[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
				for (int q = 0; q < 16; ++q) {
					if (v == 0 && (uint32_t)q == skip)
						sa = sb = 0;
					sa = cd_next(L, sa, cd_class<RANGE>(L, (wa[q >> 2] >> (8 * (q & 3))) & 0xFFu)) &
					    ACM_CD_STATE_MASK;
					sb = cd_next(L, sb, cd_class<RANGE>(L, (wb[q >> 2] >> (8 * (q & 3))) & 0xFFu)) &
					    ACM_CD_STATE_MASK;
				}
			}
		}
		/* 32 bytes (one sector) per chain per round: both halves are requested together; the
		 * sectors of the round after next are pulled into L2 meanwhile */
		const uint32_t rounds16 = (uint32_t)(chunk >> 4);
		for (uint32_t i = 0; i < rounds16; i += 2) {
			const uint4 va0 = __ldg(pa + i), va1 = __ldg(pa + i + 1);
			const uint4 vb0 = __ldg(pb + i), vb1 = __ldg(pb + i + 1);
			if (i + 4 < rounds16) {
				asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + i + 4));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + i + 4));
			}
			const uint32_t wa[8] = {va0.x, va0.y, va0.z, va0.w, va1.x, va1.y, va1.z, va1.w};
			const uint32_t wb[8] = {vb0.x, vb0.y, vb0.z, vb0.w, vb1.x, vb1.y, vb1.z, vb1.w};
			/* hit = (offset in chunk) << 14 | state; the offset is 16 i + q */
			const uint32_t hbase = (16u * i) << ACM_CD_STATE_BITS;
#pragma unroll
			for (int q = 0; q < 32; ++q) {
				const uint32_t ca = cd_class<RANGE>(L, (wa[q >> 2] >> (8 * (q & 3))) & 0xFFu);
				const uint32_t cb = cd_class<RANGE>(L, (wb[q >> 2] >> (8 * (q & 3))) & 0xFFu);
				const uint32_t ea = cd_next(L, sa, ca);
				const uint32_t eb = cd_next(L, sb, cb);
				sa = ea & ACM_CD_STATE_MASK;
				sb = eb & ACM_CD_STATE_MASK;
				const uint32_t da = ea >> ACM_CD_STATE_BITS, db = eb >> ACM_CD_STATE_BITS;
				st_pred_u32(rowa + hka, hbase | ((uint32_t)q << ACM_CD_STATE_BITS) | sa, da && hka < capm1);
				st_pred_u32(rowb + hkb, hbase | ((uint32_t)q << ACM_CD_STATE_BITS) | sb, db && hkb < capm1);
				if (da)
					++hka;
				if (db)
					++hkb;
				/* records: the 2-bit code, plus one where four patterns end (those states have the
				 * highest ids, so the raw entry says it) */
				nra += da + (ea >= thr4 ? 1u : 0u);
				nrb += db + (eb >= thr4 ? 1u : 0u);
			}
		}
		E.counts[ba] = nra;
		E.counts[ba + 1] = nrb;
		rowa[-1] = hka;
		rowb[-1] = hkb;
		if (hka > capm1 || hkb > capm1)
			*E.overflow = 1u;
	}
}
