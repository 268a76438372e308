/*
 * k234_post.cuh -- the post-passes that turn per-bucket match records into one
 * canonical, sorted list.
 *
 *   K2  k_scan_lookback      single-pass exclusive prefix sum with decoupled
 *                            look-back and warp-shuffle scans.  Replaces the five
 *                            Blelloch kernels of reference scan_kernel.cl:307-415
 *                            driven by PreScanBufferRecursive
 *                            (reference ocl_prefix_sum.c:389-498).
 *   K3  k_bucket_sort_compact
 *                            stream compaction fused with a per-bucket bitonic
 *                            sort in shared memory.  Replaces reference
 *                            compactarray.cl:40-68 (and, for the normal case, the
 *                            sort).
 *       k_compact_columns    the reference's own column-major bucket format,
 *                            for the ocl_compact_array() entry point.
 *   K4  k_radix_hist / k_radix_scatter
 *                            stable LSD radix sort of 64-bit keys, 8 bits a pass.
 *                            Replaces reference BitonicSort.cl:50-249
 *                            (power-of-two lengths only there).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "acm_tables.h"

/* ------------------------------------------------------------------------- */
/* K2: decoupled look-back exclusive scan                                    */
/* ------------------------------------------------------------------------- */

#define SCAN_THREADS 256
#define SCAN_ITEMS   8
#define SCAN_TILE    (SCAN_THREADS * SCAN_ITEMS)

#define TS_EMPTY     0u
#define TS_AGGREGATE 1u
#define TS_INCLUSIVE 2u

__device__ __forceinline__ uint64_t ts_load(const uint64_t *p)
{
	uint64_t v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void ts_store(uint64_t *p, uint32_t flag, uint32_t value)
{
	const uint64_t v = ((uint64_t)flag << 32) | value;
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/*
 * tile_state[t] = (flag << 32) | value, zeroed before launch; tile_counter hands
 * out tile ids in launch order so a tile only ever waits on tiles already running.
 */
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n,
    uint64_t *tile_state, uint32_t *tile_counter, uint32_t *total)
{
	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_warp[SCAN_THREADS / 32];
	__shared__ uint32_t s_prefix;

	pdl_trigger();
	pdl_wait();
	if (threadIdx.x == 0)
		s_tile = atomicAdd(tile_counter, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;

	uint32_t x[SCAN_ITEMS];
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		x[k] = (base + k < n) ? in[base + k] : 0u;
		sum += x[k];
	}
	/* warp inclusive scan of the per-thread sums */
	uint32_t inc = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (uint32_t)d)
			inc += y;
	}
	if (lane == 31)
		s_warp[warp] = inc;
	__syncthreads();
	uint32_t warp_off = 0, block_sum = 0;
#pragma unroll
	for (int w = 0; w < SCAN_THREADS / 32; ++w) {
		const uint32_t v = s_warp[w];
		if ((uint32_t)w < warp)
			warp_off += v;
		block_sum += v;
	}

	/* publish, then look back (warp 0) */
	if (warp == 0) {
		uint32_t excl = 0;
		if (tile == 0) {
			if (lane == 0)
				ts_store(&tile_state[0], TS_INCLUSIVE, block_sum);
		} else {
			if (lane == 0)
				ts_store(&tile_state[tile], TS_AGGREGATE, block_sum);
			int32_t pred = (int32_t)tile - 1;
			for (;;) {
				const int32_t idx = pred - (int32_t)lane;
				uint64_t st = ((uint64_t)TS_INCLUSIVE << 32);   /* before tile 0: prefix 0 */
				if (idx >= 0) {
					do {
						st = ts_load(&tile_state[idx]);
					} while ((uint32_t)(st >> 32) == TS_EMPTY);
				}
				const uint32_t flag = (uint32_t)(st >> 32);
				const uint32_t val = (uint32_t)st;
				const uint32_t inc_mask = __ballot_sync(0xffffffffu, flag == TS_INCLUSIVE);
				/* nearest inclusive predecessor = lowest lane with the flag */
				const int stop = inc_mask ? (__ffs(inc_mask) - 1) : 32;
				uint32_t contrib = ((int)lane <= stop && (int)lane < 32) ? val : 0u;
				if ((int)lane > stop)
					contrib = 0;
#pragma unroll
				for (int d = 16; d > 0; d >>= 1)
					contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
				excl += contrib;
				if (inc_mask)
					break;
				pred -= 32;
			}
			if (lane == 0)
				ts_store(&tile_state[tile], TS_INCLUSIVE, excl + block_sum);
		}
		if (lane == 0)
			s_prefix = excl;
	}
	__syncthreads();
	uint32_t run = s_prefix + warp_off + (inc - sum);
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		if (base + k < n)
			out[base + k] = run;
		run += x[k];
	}
	if (total && tile == (n - 1) / SCAN_TILE && threadIdx.x == SCAN_THREADS - 1)
		*total = s_prefix + block_sum;
}

/* ------------------------------------------------------------------------- */
/* K3: compaction + per-bucket sort                                          */
/* ------------------------------------------------------------------------- */

#define K3_THREADS 128

/*
 * Launched right after K2 without waiting for the host to learn the total: `flags[0]`
 * (a bucket overflowed -> the exact two-pass path will run instead) makes it a no-op, and a
 * bucket whose output would not fit `out_cap` is skipped and reported in flags[5] (the host
 * then grows the buffer and relaunches).  Persistent over buckets: 4 buckets per block-step.
 * push_dst (may be NULL): every key is also stored, plus key_add, at the same index of a
 * gather region (possibly another GPU's memory, NVLink peer stores) when it fits push_cap --
 * the multi-GPU gather costs no launch of its own.
 * (Folding K2 into this kernel as well -- look-back over tiles of 4 buckets -- was tried and
 * was slower for the sparse case, 25 us against 8 + 10 us, so K2 stays a kernel.)
 */
__global__ void __launch_bounds__(K3_THREADS)
k_bucket_sort_compact(const uint64_t *__restrict__ buckets, const uint32_t *__restrict__ counts,
    const uint32_t *__restrict__ offsets, uint64_t *__restrict__ out, uint32_t cap, uint32_t n_buckets,
    uint64_t out_cap, uint32_t *flags, uint64_t *__restrict__ push_dst, uint64_t push_cap, uint64_t key_add)
{
	extern __shared__ uint64_t k3_keys[];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = K3_THREADS / 32;

	pdl_trigger();
	pdl_wait();
	if (*(volatile uint32_t *)flags)        /* overflow: buckets are incomplete */
		return;
	for (uint32_t blk = blockIdx.x; blk * warps < n_buckets; blk += gridDim.x) {
		/* buckets with at most 32 records: one warp each, sorted in registers */
		const uint32_t b0 = blk * warps + (threadIdx.x >> 5);
		uint32_t cnt = 0;
		uint64_t o0 = 0;
		if (b0 < n_buckets) {
			cnt = min(__ldg(&counts[b0]), cap);
			o0 = __ldg(&offsets[b0]);
			if (cnt && o0 + cnt > out_cap) {
				if (lane == 0)
					flags[5] = 1u;
				cnt = 0;
			}
		}
		const bool push0 = push_dst && o0 + cnt <= push_cap;
		const uint32_t big = __syncthreads_or(cnt > 32);

		if (cnt > 0 && cnt <= 32) {
			uint64_t key = (lane < cnt) ? buckets[(uint64_t)b0 * cap + lane] : ~0ull;
			/* bitonic network across the warp */
#pragma unroll
			for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
				for (int j = k >> 1; j > 0; j >>= 1) {
					const uint64_t other = __shfl_xor_sync(0xffffffffu, key, j);
					const bool up = ((lane & k) == 0);
					const bool lower = ((lane & j) == 0);
					const bool take_min = (up == lower);
					key = take_min ? (key < other ? key : other) : (key > other ? key : other);
				}
			}
			if (lane < cnt) {
				out[o0 + lane] = key;
				if (push0)
					push_dst[o0 + lane] = key + key_add;
			}
		}
		if (!big)
			continue;

		/* the rare larger buckets of this step: whole block, shared memory */
		for (uint32_t w = 0; w < warps; ++w) {
			const uint32_t b = blk * warps + w;
			if (b >= n_buckets)
				break;
			const uint32_t c = min(__ldg(&counts[b]), cap);
			const uint64_t o = __ldg(&offsets[b]);
			if (c <= 32 || o + c > out_cap)
				continue;
			const bool pushw = push_dst && o + c <= push_cap;
			uint32_t P = 64;
			while (P < c)
				P <<= 1;
			for (uint32_t i = threadIdx.x; i < P; i += K3_THREADS)
				k3_keys[i] = (i < c) ? buckets[(uint64_t)b * cap + i] : ~0ull;
			__syncthreads();
			for (uint32_t k = 2; k <= P; k <<= 1) {
				for (uint32_t j = k >> 1; j > 0; j >>= 1) {
					for (uint32_t i = threadIdx.x; i < P; i += K3_THREADS) {
						const uint32_t ixj = i ^ j;
						if (ixj > i) {
							const uint64_t a = k3_keys[i], bb = k3_keys[ixj];
							const bool up = ((i & k) == 0);
							if ((a > bb) == up) {
								k3_keys[i] = bb;
								k3_keys[ixj] = a;
							}
						}
					}
					__syncthreads();
				}
			}
			for (uint32_t i = threadIdx.x; i < c; i += K3_THREADS) {
				out[o + i] = k3_keys[i];
				if (pushw)
					push_dst[o + i] = k3_keys[i] + key_add;
			}
			__syncthreads();
		}
	}
}

/*
 * K3 for k_scan_cdfa.  A bucket row holds one chunk's HITS in end-offset order as 32-bit words:
 * row[0] = number of hits, row[1 + i] = (offset in chunk << 14) | state; counts[] / offsets[]
 * are in RECORDS.  Lane i expands hit i into the state's full match list (flat4: up to four
 * pattern indices inline, ascending, one 16-byte load), placed by a warp prefix sum -- the
 * output is the canonical (end offset, pattern index) order with no sort.  Bucket b is chunk
 * (chunk0 + b) of the buffer.  A warp works on K3X_NB buckets at a time, all loads of one stage
 * issued before any is used: one bucket per warp was bound by four dependent round trips per
 * 30 records.  Same guards as k_bucket_sort_compact.
 */
#define K3X_NB 4

__device__ __forceinline__ void k3x_expand(uint64_t *__restrict__ out, uint64_t &o, uint64_t base, uint32_t hit,
    uint4 f, bool live, uint32_t lane)
{
	const uint32_t cnt = live ? (f.x >> 24) : 0u;
	uint32_t inc = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (uint32_t)d)
			inc += y;
	}
	const uint64_t hi = (base + (hit >> ACM_CD_STATE_BITS)) << ACM_KEY_PAT_BITS;
	uint64_t *dst = out + o + (inc - cnt);
	if (cnt > 0)
		dst[0] = hi | (f.x & ACM_KEY_PAT_MASK);
	if (cnt > 1)
		dst[1] = hi | f.y;
	if (cnt > 2)
		dst[2] = hi | f.z;
	if (cnt > 3)
		dst[3] = hi | f.w;
	o += __shfl_sync(0xffffffffu, inc, 31);
}

__global__ void __launch_bounds__(256)
k_bucket_expand_compact(const uint32_t *__restrict__ buckets, const uint32_t *__restrict__ counts,
    const uint32_t *__restrict__ offsets, uint64_t *__restrict__ out, uint32_t cap, uint32_t n_buckets,
    uint64_t out_cap, uint32_t *flags, const uint4 *__restrict__ flat4, uint64_t chunk0, uint32_t shift)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;

	pdl_trigger();
	pdl_wait();
	if (*(volatile uint32_t *)flags)
		return;
	for (uint32_t g = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * K3X_NB; g < n_buckets;
	     g += warps * K3X_NB) {
		/* stage 1: lane k < K3X_NB reads the header of bucket g + k */
		uint32_t nh_l = 0;
		uint64_t o_l = 0;
		if (lane < K3X_NB && g + lane < n_buckets) {
			const uint32_t nrec = __ldg(&counts[g + lane]);
			o_l = __ldg(&offsets[g + lane]);
			if (nrec) {
				if (o_l + nrec > out_cap)
					flags[5] = 1u;
				else
					nh_l = __ldcs(buckets + (uint64_t)(g + lane) * cap);
			}
		}
		uint32_t nh[K3X_NB], hit[K3X_NB];
		uint64_t o[K3X_NB];
		uint4 f[K3X_NB];
		/* stage 2: the first 32 hits of every bucket */
#pragma unroll
		for (int k = 0; k < K3X_NB; ++k) {
			nh[k] = __shfl_sync(0xffffffffu, nh_l, k);
			o[k] = __shfl_sync(0xffffffffu, o_l, k);
			hit[k] = lane < nh[k] ? __ldcs(buckets + (uint64_t)(g + k) * cap + 1 + lane) : 0u;
		}
		/* stage 3: their match lists */
#pragma unroll
		for (int k = 0; k < K3X_NB; ++k)
			f[k] = lane < nh[k] ? __ldg(flat4 + (hit[k] & ACM_CD_STATE_MASK)) : make_uint4(0, 0, 0, 0);
		/* stage 4: place and store */
#pragma unroll
		for (int k = 0; k < K3X_NB; ++k) {
			if (nh[k] == 0)
				continue;
			const uint64_t base = (chunk0 + g + k) << shift;
			k3x_expand(out, o[k], base, hit[k], f[k], lane < nh[k], lane);
			for (uint32_t i0 = 32; i0 < nh[k]; i0 += 32) {       /* more than 32 hits in the chunk */
				const uint32_t i = i0 + lane;
				const bool live = i < nh[k];
				const uint32_t h = live ? __ldcs(buckets + (uint64_t)(g + k) * cap + 1 + i) : 0u;
				const uint4 ff = live ? __ldg(flat4 + (h & ACM_CD_STATE_MASK)) : make_uint4(0, 0, 0, 0);
				k3x_expand(out, o[k], base, h, ff, live, lane);
			}
		}
	}
}

/* reference bucket format -> [total, values..., tail]  (reference compactarray.cl:40-68).
 * dst holds dst_cap ints: counts that would run past it (garbage in, or a caller's short buffer)
 * are cut off instead of written out of bounds. */
__global__ void k_compact_columns(int32_t *__restrict__ dst, const int32_t *__restrict__ src,
    const int32_t *__restrict__ prefix, int32_t len, int32_t max_results, int64_t dst_cap)
{
	const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid == 0) {
		const int64_t total = (int64_t)prefix[len - 1] + src[len - 1];
		dst[0] = (int32_t)total;
		if (total >= 0 && total + 1 < dst_cap)
			dst[total + 1] = src[(int64_t)max_results * len];
	}
	if (gid >= len)
		return;
	const int64_t off = prefix[gid];
	const int32_t m = src[gid];
	if (off < 0)
		return;
	for (int32_t i = 0; i < m && i < max_results - 1 && off + 1 + i < dst_cap; ++i)
		dst[off + 1 + i] = src[(int64_t)len * (i + 1) + gid];
}

/* ------------------------------------------------------------------------- */
/* K4: LSD radix sort, 8 bits per pass                                       */
/* ------------------------------------------------------------------------- */

#define RS_THREADS 256
#define RS_ROUNDS  16
#define RS_TILE    (RS_THREADS * RS_ROUNDS)

/* hist is digit-major: hist[digit * nblocks + block] */
__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint64_t flip,
    uint32_t *__restrict__ hist, uint32_t nblocks)
{
	__shared__ uint32_t s_h[256];
	s_h[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
	for (int r = 0; r < RS_ROUNDS; ++r) {
		const uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
		if (i < n)
			atomicAdd(&s_h[((keys[i] ^ flip) >> shift) & 0xFF], 1u);
	}
	__syncthreads();
	hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];
}

/* stable scatter: item order within a block is (round, thread) */
__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const uint64_t *__restrict__ keys, uint64_t *__restrict__ out, uint64_t n, int shift,
    uint64_t flip, const uint32_t *__restrict__ hist_scan, uint32_t nblocks)
{
	__shared__ uint32_t s_run[256];                     /* next free slot per digit   */
	__shared__ uint32_t s_wc[RS_THREADS / 32][256];     /* per-warp digit counts      */
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	s_run[threadIdx.x] = hist_scan[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
	const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
	for (int r = 0; r < RS_ROUNDS; ++r) {
		for (int w = 0; w < RS_THREADS / 32; ++w)
			s_wc[w][threadIdx.x] = 0;
		__syncthreads();
		const uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
		const bool live = i < n;
		uint64_t key = 0;
		uint32_t digit = 0, rank = 0;
		if (live) {
			key = keys[i];
			digit = (uint32_t)(((key ^ flip) >> shift) & 0xFF);
		}
		const unsigned act = __ballot_sync(0xffffffffu, live);
		if (live) {
			const unsigned peers = __match_any_sync(act, digit);
			rank = __popc(peers & ((1u << lane) - 1));
			if (rank == 0)
				s_wc[warp][digit] = __popc(peers);
		}
		__syncthreads();
		if (live) {
			uint32_t pos = s_run[digit] + rank;
			for (uint32_t w = 0; w < warp; ++w)
				pos += s_wc[w][digit];
			out[pos] = key;
		}
		__syncthreads();
		{
			uint32_t add = 0;
			for (int w = 0; w < RS_THREADS / 32; ++w)
				add += s_wc[w][threadIdx.x];
			s_run[threadIdx.x] += add;
		}
		__syncthreads();
	}
}

/* (key, value) u32 pairs <-> u64 keys for acm_sort_pairs_u32 */
__global__ void k_pack_pairs(const uint32_t *__restrict__ k, const uint32_t *__restrict__ v,
    uint64_t *__restrict__ out, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = ((uint64_t)k[i] << 32) | v[i];
}

__global__ void k_unpack_pairs(const uint64_t *__restrict__ in, uint32_t *__restrict__ k,
    uint32_t *__restrict__ v, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		k[i] = (uint32_t)(in[i] >> 32);
		v[i] = (uint32_t)in[i];
	}
}

/*
 * Gather step of the multi-GPU path: copy this GPU's sorted keys into the gather buffer of
 * the collecting GPU (a peer mapping over NVLink, or local memory on that GPU itself),
 * shifting the offsets to stream positions on the way.
 */
__global__ void k_push_keys(const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ dst,
    uint64_t key_add)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
	     i += (uint64_t)gridDim.x * blockDim.x)
		dst[i] = keys[i] + key_add;
}

/*
 * Same, launched inside the step before the host knows the count: flags[1] is the total,
 * flags[0] / "does not fit" are the cases the host repairs afterwards (acm_scan_finish pushes
 * then), so this kernel simply does nothing for them.
 */
__global__ void k_push_keys_dev(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ flags,
    uint64_t out_cap, uint64_t *__restrict__ dst, uint64_t dst_cap, uint64_t key_add)
{
	const uint64_t n = flags[1];
	if (flags[0] || n > out_cap || n > dst_cap)
		return;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
	     i += (uint64_t)gridDim.x * blockDim.x)
		dst[i] = keys[i] + key_add;
}

/* per-pattern counts of a key list */
__global__ void k_histogram(const uint64_t *__restrict__ keys, uint64_t n, unsigned long long *counts)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
	     i += (uint64_t)gridDim.x * blockDim.x)
		atomicAdd(&counts[keys[i] & ACM_KEY_PAT_MASK], 1ull);
}

/* ------------------------------------------------------------------------- */
/* synthetic streams                                                         */
/* ------------------------------------------------------------------------- */

__host__ __device__ __forceinline__ uint64_t acm_mix64(uint64_t seed, uint64_t i)
{
	uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

/* dst must be 8-byte aligned and offset a multiple of 8 for the fast path; the host wrapper guarantees it */
__global__ void k_synth_fill(uint64_t *__restrict__ dst, uint64_t nwords, uint64_t seed, uint64_t word0)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords;
	     i += (uint64_t)gridDim.x * blockDim.x)
		dst[i] = acm_mix64(seed, word0 + i);
}

__global__ void k_plant(uint8_t *__restrict__ buf, uint64_t n, uint64_t buf_offset,
    const uint64_t *__restrict__ pos, const uint32_t *__restrict__ blob_off,
    const uint32_t *__restrict__ len, uint32_t count, const uint8_t *__restrict__ blob)
{
	/* one warp per plant; plants are disjoint by construction */
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	if (w >= count)
		return;
	const uint64_t p = pos[w];
	const uint32_t L = len[w];
	const uint8_t *src = blob + blob_off[w];
	for (uint32_t k = lane; k < L; k += 32) {
		const uint64_t g = p + k;
		if (g >= buf_offset && g - buf_offset < n)
			buf[g - buf_offset] = src[k];
	}
}

/* ------------------------------------------------------------------------- */
/* status words -> mapped pinned host memory                                 */
/* ------------------------------------------------------------------------- */

__global__ void
k_publish_flags(const uint32_t *__restrict__ flags, uint32_t *__restrict__ h_flags)
{
	pdl_wait();
	if (threadIdx.x < 8) {
		h_flags[threadIdx.x] = flags[threadIdx.x];
		__threadfence_system();
	}
}
