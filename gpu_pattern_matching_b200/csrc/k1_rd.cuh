/*
 * k1_rd.cuh -- the dense-output scan (word lists over text: one match per ~9 bytes; BASELINE
 * configs[3]) and its expanding post-pass.
 *
 *   k_scan_rd     reference ahomatch.cl:50-77 (one transition per byte, one work-item per chunk,
 *                 cold start Lmax-1 bytes early) out of the ROW-DISPLACED table of
 *                 acm_core.c:build_cdfa_rd, all of it in shared memory.
 *   k_rd_expand   reference compactarray.cl:40-68: every logged hit becomes its (end offset,
 *                 pattern index) records at the position the walk already worked out for it.
 *   (between them: k_scan_lookback over the per-chunk record counts, as for every other kernel)
 *
 * What bounds a DFA walk out of shared memory is the L1 pipe (one wavefront per conflict-free
 * 128 bytes per clock) and the issue slots; the predecessor of this kernel (8-byte state record +
 * 16-bit entry per byte, per-lane 32-byte sector loads, one predicated global store per hit) ran
 * at 17 shared-memory wavefronts per 64 input bytes and 2.4 x the input in DRAM reads.  Here:
 *
 *   transition   ONE 4-byte lookup, tab[off(state) + class]; the entry names the class it was
 *                stored for, and only a lane whose entry is somebody else's (its state has no
 *                explicit edge on this byte) does a second lookup in the row of the nearest
 *                shallow state.  The state is the entry itself: no per-state record.  Dense rows
 *                are 33 words apart, so lanes reading the same letter out of different rows hit
 *                different banks.
 *   input        coalesced: 4 lanes fetch 64 contiguous bytes of a chunk with cp.async (16 bytes
 *                each) into a swizzled shared-memory tile one piece ahead of use; every byte of
 *                the stream crosses DRAM -> L2 -> SM once, in whole 64-byte pieces.  A lane then
 *                pulls its chain's 64 bytes into registers with four conflict-free LDS.128.
 *   hits         8 bytes: the entry (base of the state + how many patterns end there) and a word
 *                the lane keeps up to date with one add per byte -- offset in chunk, records its
 *                chunk has produced so far, lane.  The hits of a step (3 of 32 lanes on English
 *                text) are ranked with ballot/popc and stored side by side at the end of the
 *                region's log: consecutive addresses, so a sector is complete a step or two later
 *                and leaves L2 whole.  No per-chunk rows, no atomics, no staging.  Because a
 *                hit knows its rank among the records of its chunk, the post-pass needs no sort:
 *                hits are logged in walk order (step, lane), the canonical order is (chunk, step),
 *                and every hit can be expanded independently of all others.
 *
 * Geometry: chunk = 2^shift bytes (256 unless the halo asks for more), cut on absolute multiples
 * of the chunk size in the buffer; a REGION is 32 consecutive chunks, one per lane of a warp;
 * warps take regions round-robin, so the grid sweeps the stream as one front.
 */
#pragma once

#include "k1_scan.cuh"
#include "k234_post.cuh"

#define RD_PIECE      64                    /* bytes per chain per staging step                 */
#define RD_WARP_SMEM  (32 * RD_PIECE)        /* per warp: the staging tile                       */
#define RD_CHECK      8                     /* steps between two looks at the room left in the log */
#define RD_LOG_ALIGN  32                    /* log_cap is a multiple of this                    */
#define RD_MODE_LOG    0
#define RD_MODE_DIRECT 1                    /* second pass of the exact two-pass path: records straight to out[offsets[chunk] ...] */

/* entry (= state):  column | records << 5 | dense row << 8 | base << 16   (acm_core.c:build_cdfa_rd)
 * hit word 0:       the entry
 * hit word 1:       lane | (records of the chunk up to and including this hit) << 5 | (offset in chunk + 1) << 17 */
#define RD_ENT_RECS(e)  (((e) >> 5) & 7u)
#define RD_W1_LANE(w)   ((w) & 31u)
#define RD_W1_RECS(w)   (((w) >> 5) & 0xFFFu)
#define RD_W1_OFF1(w)   ((w) >> 17)
#define RD_MAX_SHIFT    9                   /* chunks of at most 512 bytes: <= 2048 records, 12 bits */

struct RdCtx {
	uint2    *log;            /* [regions][log_cap] hits in walk order               */
	uint32_t *loglen;         /* [regions]                                           */
	uint32_t  log_cap;        /* entries per region, > 32 * RD_CHECK                 */
	uint32_t  tab_bytes;      /* rd_len * 4 rounded up to 16                         */
	uint32_t  mode;
	uint32_t  lo, cmax;       /* class(b) = min(b - lo, cmax)                        */
	uint32_t *tabw_out;       /* the kernel reports (shared address of the table) >> 2: hits carry absolute bases */
};

__device__ __forceinline__ uint4 lds_v4_volatile(uint32_t saddr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
	return v;
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}

__device__ __forceinline__ void cp_async_commit_wait()
{
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

static_assert(ACM_RD_ROW * 4 == 132, "rd_next hard-codes the dense row stride in its PTX");

/*
 * One transition: the entry for (state e, class c).  Inside the kernel the base field of every
 * entry is ABSOLUTE -- (shared byte address of the slot) >> 2, patched in after staging -- so the
 * common case is shift, add, shift, load; tab_sa (shared address of the table) is only needed for
 * the row of the shallow ancestor, dense row r at tab_sa + 4 * ACM_RD_ROW * r.  Written in PTX so
 * that the second lookup stays a predicated load into the same register (no branch).
 */
__device__ __forceinline__ uint32_t rd_next(uint32_t tab_sa, uint32_t e, uint32_t c)
{
	uint32_t x;
	asm("{\n"
	    ".reg .pred p;\n"
	    ".reg .b32 a, t;\n"
	    "shr.u32 a, %1, 16;\n"
	    "add.u32 a, a, %2;\n"
	    "shl.b32 a, a, 2;\n"
	    "ld.shared.u32 %0, [a];\n"
	    "xor.b32 t, %0, %2;\n"
	    "and.b32 t, t, 31;\n"
	    "setp.ne.u32 p, t, 0;\n"
	    "prmt.b32 t, %1, 0, 0x4441;\n"
	    "mad.lo.u32 a, %2, 4, %3;\n"
	    "mad.lo.u32 t, t, 132, a;\n"
	    "@p ld.shared.u32 %0, [t];\n"
	    "}\n" : "=&r"(x) : "r"(e), "r"(c), "r"(tab_sa));
	return x;
}

/*
 * All lanes, converged: one step of a lane's bookkeeping and the append of the step's hits.
 * t = (entry & 0xE0) = 32 x (patterns ending here); w1 gains t and one offset step; a lane with
 * t != 0 stores {entry, w1} at its rank among the step's hits, right behind the hits of the steps
 * before: qptr is the end of the region's log.  The caller makes sure 32 more entries fit.
 * In PTX so that the address is one wide multiply-add of the rank and the pointer moves by one
 * wide multiply-add of the hit count (ptxas turns each into a shift-and-add pair, LEA + LEA.HI.X;
 * forcing IMAD.WIDE with an opaque multiplier was measured and is slower: 1.16 ms against 1.12).
 */
__device__ __forceinline__ void rd_step_emit(uint2 *&qptr, uint32_t e, uint32_t &w1, uint32_t lt)
{
	const uint32_t t = e & 0xE0u;
	w1 += t + (1u << 17);
	const bool hit = t != 0;
	const uint32_t m = __ballot_sync(FULL_MASK, hit);
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    ".reg .b32 r, n;\n"
	    ".reg .b64 a;\n"
	    "and.b32 r, %1, %2;\n"
	    "popc.b32 r, r;\n"
	    "mad.wide.u32 a, r, 8, %0;\n"
	    "setp.ne.b32 p, %3, 0;\n"
	    "@p st.global.v2.u32 [a], {%4, %5};\n"
	    "popc.b32 n, %1;\n"
	    "mad.wide.u32 %0, n, 8, %0;\n"
	    "}\n" : "+l"(qptr) : "r"(m), "r"(lt), "r"((uint32_t)hit), "r"(e), "r"(w1) : "memory");
}

/*
 * Shared memory: [table: tab_bytes][per warp: 32 x 64-byte staging rows, swizzled][mbarrier]
 * blockDim.x = 32 x (warps that fit beside the table), one CTA per SM.
 */
__global__ void __launch_bounds__(1024, 1)
k_scan_rd(const __grid_constant__ AutDev A, const __grid_constant__ EmitCtx E, const __grid_constant__ RdCtx R,
    const uint8_t *__restrict__ data, uint64_t limit)
{
	extern __shared__ __align__(1024) uint8_t rd_smem[];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
	uint8_t *wsm = rd_smem + ((R.tab_bytes + 1023) & ~1023u) + warp * RD_WARP_SMEM;
	uint64_t *bar = reinterpret_cast<uint64_t *>(rd_smem + ((R.tab_bytes + 1023) & ~1023u) + warps * RD_WARP_SMEM);

	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		mbar_expect_tx(bar, R.tab_bytes);
		for (uint32_t off = 0; off < R.tab_bytes; off += 16384)
			bulk_g2s(rd_smem + off, reinterpret_cast<const uint8_t *>(A.rd_tab) + off,
			    min(R.tab_bytes - off, 16384u), bar);
	}
	__syncthreads();
	mbar_wait(bar, 0);

	/* make the base field of every entry absolute (see rd_next); k_rd_expand and the byte-wise path
	 * subtract tabw again to index the match lists */
	const uint32_t tab_sa = smem_u32(rd_smem);
	const uint32_t tabw = tab_sa >> 2;
	{
		uint32_t *tab = reinterpret_cast<uint32_t *>(rd_smem);
		for (uint32_t i = threadIdx.x; i < R.tab_bytes / 4; i += blockDim.x)
			tab[i] += tabw << 16;
		if (blockIdx.x == 0 && threadIdx.x == 0)
			*R.tabw_out = tabw;
	}
	__syncthreads();
	const uint32_t e_root = tabw << 16;             /* the root: base 0, dense row 0 */
	const uint32_t stage_sa = smem_u32(wsm);
	const uint32_t lt = lanemask_lt();
	const uint32_t shift = E.shift;
	const uint64_t chunk = 1ull << shift;
	const uint64_t halo = A.max_len > 0 ? (uint64_t)(A.max_len - 1) : 0;
	const uint64_t k_first = E.emit_lo >> shift, k_last = (limit - 1) >> shift;
	const uint64_t n_regions = (k_last - k_first + 32) / 32;
	const uint32_t G = (uint32_t)(chunk / RD_PIECE);
	const uint32_t H = ((uint32_t)halo + 3u) & ~3u;     /* fast path: the cold start covers whole words */
	const uint32_t npre = (H + RD_PIECE - 1) / RD_PIECE;
	/* this lane's row in the staging tile; 16-byte units of a row are XOR-swizzled by (row >> 1) & 3
	 * so that 8 consecutive rows read the same unit from 8 different bank groups */
	const uint32_t my_row_sa = stage_sa + lane * RD_PIECE;
	const uint32_t my_swz = (lane >> 1) & 3u;
	const uint32_t q_room = R.log_cap - 32 * RD_CHECK;  /* more hits than this: the next RD_CHECK steps might not fit */

	for (uint64_t r = (uint64_t)blockIdx.x * warps + warp; r < n_regions; r += (uint64_t)gridDim.x * warps) {
		const uint64_t kb = k_first + 32 * r;             /* first chunk of the region */
		const uint64_t a_first = kb << shift;
		const uint64_t k = kb + lane;                     /* this lane's chunk */
		const bool have = k <= k_last;
		const uint64_t bidx = k - k_first;                /* its bucket index in the scan */
		uint2 *const qdst = R.log + r * R.log_cap;
		uint2 *qptr = qdst;                                 /* end of the region's log */
		uint2 *const qfull = qdst + q_room, *const qlast = qdst + (R.log_cap - 32);
		const bool fast = R.mode == RD_MODE_LOG && a_first >= E.emit_lo && ((kb + 32) << shift) <= limit &&
		    a_first >= (uint64_t)npre * RD_PIECE && a_first >= H && a_first - H >= E.valid_lo;
		if (fast) {
			/* piece p of every chunk of the region: bytes [64 p, 64 p + 64) of the chunk, p = -npre .. G-1 */
			auto issue = [&](int p) {
				const uint8_t *src = data + a_first + (int64_t)p * RD_PIECE + (lane & 3u) * 16;
#pragma unroll
				for (uint32_t j = 0; j < 4; ++j) {
					const uint32_t row = 8 * j + (lane >> 2);
					cp_async16(stage_sa + row * RD_PIECE + (((lane & 3u) ^ ((row >> 1) & 3u)) << 4),
					    src + ((uint64_t)row << shift));
				}
			};
			/* half a row (32 bytes) into registers: two conflict-free LDS.128 */
			auto fetch = [&](uint32_t (&w)[8], uint32_t half) {
#pragma unroll
				for (uint32_t q = 0; q < 2; ++q) {
					const uint4 v = lds_v4_volatile(my_row_sa + (((2 * half + q) ^ my_swz) << 4));
					w[4 * q] = v.x;
					w[4 * q + 1] = v.y;
					w[4 * q + 2] = v.z;
					w[4 * q + 3] = v.w;
				}
			};
			uint32_t e = e_root;
			uint32_t w1 = lane;                            /* offset 0, no records yet */
			issue(-(int)npre);
			for (int p = -(int)npre; p < 0; ++p) {
				cp_async_commit_wait();
				__syncwarp();
				/* cold start: the last H bytes in front of the chunk (H = halo rounded up to whole
				 * words: starting a little earlier is as good), nothing reported */
				const uint32_t w0 = (p == -(int)npre) ? (npre * RD_PIECE - H) / 4 : 0u;   /* warp-uniform */
				for (uint32_t half = 0; half < 2; ++half) {
					uint32_t w[8];
					fetch(w, half);
					if (half == 1) {
						__syncwarp();
						issue(p + 1);
					}
#pragma unroll
					for (uint32_t wi = 0; wi < 8; ++wi) {
						if (8 * half + wi < w0)
							continue;
#pragma unroll
						for (uint32_t kk = 0; kk < 4; ++kk)
							e = rd_next(tab_sa, e, min(__byte_perm(w[wi], 0u, 0x4440u + kk) - R.lo, R.cmax));
					}
				}
			}
			for (uint32_t p = 0; p < G; ++p) {
				cp_async_commit_wait();
				__syncwarp();
				for (uint32_t half = 0; half < 2; ++half) {
					uint32_t w[8];
					fetch(w, half);
					if (half == 1) {
						/* every lane has its row: the tile is free for the next piece, which has the
						 * 32 steps below (~8 000 cycles) to arrive */
						__syncwarp();
						if (p + 1 < G)
							issue((int)p + 1);
					}
#pragma unroll
					for (uint32_t i = 0; i < RD_PIECE / 2; ++i) {
						if (i % RD_CHECK == 0 && qptr > qfull) {    /* warp-uniform: the log is full */
							*E.overflow = 1u;
							qptr = qdst;                              /* the scan is repeated exactly; keep going harmlessly */
						}
						const uint32_t b = __byte_perm(w[i >> 2], 0u, 0x4440u + (i & 3));
						e = rd_next(tab_sa, e, min(b - R.lo, R.cmax));
						rd_step_emit(qptr, e, w1, lt);
					}
				}
			}
			E.counts[bidx] = RD_W1_RECS(w1);
		} else {
			/* first / last region of a scan, and the second pass of the exact two-pass path:
			 * byte-wise, every lane over the same number of steps so that the warp stays converged */
			uint64_t lo = k << shift, hi = (k + 1) << shift;
			if (lo < E.emit_lo)
				lo = E.emit_lo;
			if (hi > limit)
				hi = limit;
			uint64_t start = lo > halo ? lo - halo : 0;
			if (start < E.valid_lo)
				start = E.valid_lo;
			uint64_t *dst = (R.mode == RD_MODE_DIRECT && have) ? E.out + E.offsets[bidx] : nullptr;
			uint32_t nrec = 0, e = e_root;
			const int64_t p0 = (int64_t)(k << shift) - (int64_t)halo;
			for (uint32_t i = 0; i < (uint32_t)(halo + chunk); ++i) {
				const int64_t pos = p0 + i;
				const bool live = have && pos >= (int64_t)start && pos < (int64_t)hi;
				bool hit = false;
				if (live) {
					e = rd_next(tab_sa, e, min((uint32_t)__ldg(data + pos) - R.lo, R.cmax));
					hit = (e & 0xE0u) != 0 && pos >= (int64_t)lo;
				}
				if (R.mode == RD_MODE_LOG) {
					/* w1 as the fast path has it BEFORE this step: offset << 17 | records so far << 5 | lane */
					uint32_t w = (((uint32_t)(pos - (int64_t)(k << shift)) & 0x1FFu) << 17) | (nrec << 5) | lane;
					if (qptr > qlast) {                           /* warp-uniform */
						*E.overflow = 1u;
						qptr = qdst;
					}
					rd_step_emit(qptr, hit ? e : 0u, w, lt);
					if (hit)
						nrec += RD_ENT_RECS(e);
				} else if (hit) {
					const uint4 f = __ldg(A.rd_flat4 + ((e >> 16) - tabw));
					const uint32_t cnt = f.x >> 24;
					if (dst) {
						const uint64_t hi64 = (uint64_t)pos << ACM_KEY_PAT_BITS;
						dst[nrec] = hi64 | (f.x & ACM_KEY_PAT_MASK);
						if (cnt > 1)
							dst[nrec + 1] = hi64 | f.y;
						if (cnt > 2)
							dst[nrec + 2] = hi64 | f.z;
						if (cnt > 3)
							dst[nrec + 3] = hi64 | f.w;
					}
					nrec += cnt;
				}
			}
			if (R.mode == RD_MODE_LOG && have)
				E.counts[bidx] = nrec;
		}
		if (R.mode == RD_MODE_LOG && lane == 0)
			R.loglen[r] = (uint32_t)(qptr - qdst);
	}
}

/*
 * Post-pass of k_scan_rd: every hit of every region's log becomes its records.  A hit of lane l of
 * region r belongs to chunk c = 32 r + l; its records go to out[offsets[c] + (records of c before
 * the hit)], and that number is in the hit: no ordering between hits, no sort, no look-back.
 * One warp per region (grid-stride).  The records are put together in shared memory -- a window of
 * RD_K3_WIN records of the region's part of the list at a time, one window for ordinary text --
 * and leave for global memory as whole lines: written one by one where they belong, 8 bytes here,
 * 8 bytes there, every 32-byte sector would reach DRAM half filled and be read back for the merge
 * (measured: 1.5 x the list in DRAM reads on top of its writes).  128 hits per step and lane-group,
 * their match-list loads (one 16-byte load each) in flight together.
 * Guards as in k_bucket_sort_compact: flags[0] (a log overflowed) makes the kernel a no-op, a region
 * that does not fit out_cap is skipped and reported in flags[5].
 */
#define RD_K3_THREADS 256
#define RD_K3_WIN     1024
#define RD_K3_SMEM    ((RD_K3_THREADS / 32) * RD_K3_WIN * 8)

__global__ void __launch_bounds__(RD_K3_THREADS)
k_rd_expand(const uint2 *__restrict__ log, const uint32_t *__restrict__ loglen, uint32_t log_cap,
    uint32_t n_regions, const uint32_t *__restrict__ offsets, uint32_t n_chunks, uint64_t *__restrict__ out,
    uint64_t out_cap, uint32_t *flags, const uint4 *__restrict__ flat4_rel, uint64_t chunk0, uint32_t shift)
{
	extern __shared__ __align__(16) uint64_t rx_smem[];
	const uint32_t lane = threadIdx.x & 31;
	uint64_t *win = rx_smem + (size_t)(threadIdx.x >> 5) * RD_K3_WIN;
	const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;

	pdl_trigger();
	pdl_wait();
	if (*(volatile uint32_t *)flags)
		return;
	/* hits carry the ABSOLUTE base of k_scan_rd's shared-memory table: flags[8] = its word address */
	const uint4 *flat4 = flat4_rel - flags[8];
	const uint32_t grand_total = flags[1];
	for (uint32_t r = gw; r < n_regions; r += nw) {
		const uint32_t n = min(loglen[r], log_cap);
		const uint2 *src = log + (size_t)r * log_cap;
		const uint32_t c_mine = 32 * r + lane;
		const uint32_t off_mine = c_mine < n_chunks ? __ldg(offsets + c_mine) : grand_total;
		const uint32_t base = __shfl_sync(FULL_MASK, off_mine, 0);
		const uint32_t next = 32 * r + 32 < n_chunks ? __ldg(offsets + 32 * r + 32) : grand_total;
		const uint32_t T = next - base;                       /* records of this region */
		const uint64_t region_base = (chunk0 + 32ull * r) << shift;
		if ((uint64_t)base + T > out_cap) {
			if (lane == 0 && T)
				flags[5] = 1u;
			continue;
		}
		for (uint32_t w0 = 0; w0 < T; w0 += RD_K3_WIN) {
			for (uint32_t i0 = 0; i0 < n; i0 += 128) {
				uint2 h[4];
				uint4 f[4];
#pragma unroll
				for (int u = 0; u < 4; ++u)
					h[u] = i0 + 32 * u + lane < n ? __ldg(src + i0 + 32 * u + lane) : make_uint2(0u, 0u);
#pragma unroll
				for (int u = 0; u < 4; ++u)
					f[u] = i0 + 32 * u + lane < n ? __ldg(flat4 + (h[u].x >> 16)) : make_uint4(0, 0, 0, 0);
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const uint32_t l = RD_W1_LANE(h[u].y);
					const uint32_t o = __shfl_sync(FULL_MASK, off_mine, l);
					const uint32_t c = f[u].x >> 24;              /* 0 for the lanes past the end */
					/* position of the hit's first record in the window (may be negative or beyond it) */
					const uint32_t q = o - base + RD_W1_RECS(h[u].y) - c - w0;
					const uint64_t end = region_base + ((uint64_t)l << shift) + RD_W1_OFF1(h[u].y) - 1;
					const uint64_t hi = end << ACM_KEY_PAT_BITS;
					if (c > 0 && q < RD_K3_WIN)
						win[q] = hi | (f[u].x & ACM_KEY_PAT_MASK);
					if (c > 1 && q + 1 < RD_K3_WIN)
						win[q + 1] = hi | f[u].y;
					if (c > 2 && q + 2 < RD_K3_WIN)
						win[q + 2] = hi | f[u].z;
					if (c > 3 && q + 3 < RD_K3_WIN)
						win[q + 3] = hi | f[u].w;
				}
			}
			__syncwarp();
			const uint32_t m = T - w0 < RD_K3_WIN ? T - w0 : RD_K3_WIN;
			uint64_t *dst = out + base + w0;
			for (uint32_t q = lane; q < m; q += 32)
				dst[q] = win[q];
			__syncwarp();
		}
	}
}
