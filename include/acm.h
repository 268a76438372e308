/*
 * acm.h -- native C ABI of the B200 Aho-Corasick matcher.
 *
 * Everything the reference-compatible entry points (acsmx.h, iacsmx.h, databuf.h,
 * ocl_aho_match.h, ocl_worker.h, ocl_prefix_sum.h, ocl_compact_array.h,
 * ocl_bitonic_sort.h) do on the device goes through the functions declared
 * here; language bindings (ctypes, cgo, JNI ...) can bind either layer.  Plain
 * pointers and sizes only.  Functions return ACM_OK (0) or a negative ACM_ERR_*;
 * acm_last_error() gives the text for the calling thread.  Nothing exit()s.
 *
 * The hot path these functions replace in the reference is the five-call
 * sequence of cpu_worker() (reference ocl_aho_grep.c:116-137):
 *   databuf_copy_host_to_device -> ocl_aho_match -> databuf_copy_device_to_host
 *   -> databuf_process_results -> databuf_reset
 * plus the optional post-passes ocl_prefix_sum / ocl_compact_array /
 * ocl_bitonic_sort (reference databuf.c:633-706, ocl_bitonic_sort.c:140).
 */
#ifndef ACM_H
#define ACM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACM_OK                  0
#define ACM_ERR_CUDA          -10   /* a CUDA runtime call failed              */
#define ACM_ERR_NOMEM         -11
#define ACM_ERR_ARG           -12
#define ACM_ERR_STATE         -13   /* call made in the wrong order            */
#define ACM_ERR_LIMIT         -14   /* > 2^24-1 patterns, > 2^30 states, > 2^40 bytes per scan (2^30 for the dense-output kernels) */
#define ACM_ERR_EMPTY_PATTERN -15   /* zero-length pattern (kept, never matches) */
#define ACM_ERR_IO            -16
#define ACM_ERR_NO_DEVICE     -17   /* no CUDA device: there is no CPU fallback */

const char *acm_last_error(void);

struct acm_device;      /* one GPU: ordinal, stream, events                    */
struct acm_automaton;   /* device-resident automaton + filters                 */
struct acm_scanner;     /* scratch and result buffers for scans up to a size   */
struct acm_tables;      /* host-side compiled automaton (private layout)       */

/* ---- device ---- */
int   acm_device_count(void);
int   acm_device_open(int ordinal, struct acm_device **out);
void  acm_device_close(struct acm_device *);
int   acm_device_ordinal(struct acm_device *);
/* the cudaStream_t all work of this device object is ordered on */
void *acm_device_stream(struct acm_device *);
/* adopt an external stream (e.g. torch.cuda.current_stream().cuda_stream); NULL restores the own stream */
int   acm_device_set_stream(struct acm_device *, void *cuda_stream);
int   acm_device_sync(struct acm_device *);
/* process-wide default device object for the compat layer (ordinal from ACM_DEVICE, default 0) */
struct acm_device *acm_default_device(void);

/* raw memory helpers (bindings that cannot call the CUDA runtime themselves) */
int   acm_dev_alloc(struct acm_device *, size_t bytes, void **d_ptr);
void  acm_dev_free(struct acm_device *, void *d_ptr);
/* pinned host memory, preferably on the NUMA node of the calling thread's current CUDA device (of `dev`
 * in the _near form): on a two-socket multi-GPU host the H2D copies then stay on the GPU's own socket.
 * A hint only (set_mempolicy(MPOL_PREFERRED) around the allocation); ACM_NUMA=0 turns it off. */
int   acm_host_alloc_pinned(size_t bytes, void **h_ptr);
int   acm_host_alloc_pinned_near(struct acm_device *dev, size_t bytes, void **h_ptr);
void  acm_host_free_pinned(void *h_ptr);
int   acm_dev_memset(struct acm_device *, void *d_dst, int value, size_t bytes);         /* async on the stream */
int   acm_memcpy_h2d(struct acm_device *, void *d_dst, const void *h_src, size_t bytes);  /* async on the stream */
int   acm_memcpy_d2h(struct acm_device *, void *h_dst, const void *d_src, size_t bytes);  /* async on the stream */
/* D2H on a side stream (ordered after the work queued so far), so it overlaps later kernels; acm_side_sync waits for it */
int   acm_memcpy_d2h_side(struct acm_device *, void *h_dst, const void *d_src, size_t bytes);
int   acm_side_sync(struct acm_device *);
/* D2H of nseg device segments packed back to back into h_dst, on the side stream and NOT ordered
 * against the main stream (for data already known to be complete); waits for the copies */
int   acm_memcpy_d2h_segments(struct acm_device *, void *h_dst, const void *const *d_src,
          const uint64_t *bytes, uint32_t nseg);
/* the same, not waited for: complete after the next acm_side_sync() */
int   acm_memcpy_d2h_segments_async(struct acm_device *, void *h_dst, const void *const *d_src,
          const uint64_t *bytes, uint32_t nseg);

/* ---- automaton ---- */
int   acm_automaton_upload(struct acm_device *, const struct acm_tables *, struct acm_automaton **out);
void  acm_automaton_free(struct acm_automaton *);
uint32_t acm_automaton_states(const struct acm_automaton *);
uint32_t acm_automaton_patterns(const struct acm_automaton *);
/* host array [patterns]: length of every pattern by index (add order); lives as long as the automaton */
const uint32_t *acm_automaton_pattern_lengths(const struct acm_automaton *);
int      acm_automaton_max_pattern_len(const struct acm_automaton *);
int      acm_automaton_min_pattern_len(const struct acm_automaton *);
int      acm_automaton_alphabet(const struct acm_automaton *);
size_t   acm_automaton_device_bytes(const struct acm_automaton *);
/* which kernel ACM_MODE_AUTO picks: 1 = sampled gram filter, 2 = 2-byte start filter, 3 = DFA,
 * 4 = class-compressed DFA with the hot rows in shared memory (small automata over few distinct bytes) */
int      acm_automaton_default_mode(const struct acm_automaton *);
uint32_t acm_automaton_gram_count(const struct acm_automaton *);
/* columns of the class-compressed table (mode 4), 0 if the automaton has none */
int      acm_automaton_cdfa_classes(const struct acm_automaton *);
/* window stride of the sampled kernel for this automaton: 8 (every pattern >= 10 bytes), 4 (>= 7), 0 (not available).
 * ACM_SAMPLE_STRIDE=4 in the environment at compile (acsm_compile) time forces 4. */
int      acm_automaton_sample_stride(const struct acm_automaton *);
/* mixed sets (a few patterns shorter than 7 bytes among many long ones; at most one in eight, or
 * ACM_HYBRID=1 at compile time): the patterns shorter than this many bytes are left out of the
 * sampled filter and found by a second pass of the 2-byte start filter into the same result
 * buckets; mode 1 then means both passes.  0 = not a split automaton.  ACM_HYBRID=0 disables. */
int      acm_automaton_split_len(const struct acm_automaton *);

/* ---- scan ---- */
struct acm_scan_params {
	int      mode;          /* 0 auto, 1 sampled4, 2 start2, 3 dfa, 4 cdfa                */
	int      bucket_shift;  /* log2 bytes of input per result bucket; 0 = default (17 sampled, 15 otherwise) */
	int      bucket_cap;    /* records per bucket before the exact 2-pass fallback; 0 = default */
	int      timing;        /* 1: CUDA events around each kernel; 2: around the scan stage only (ms_scan);
	                           3: sampled mode: around the streaming kernel alone, without the resolve kernel */
	int      dfa_chunk;     /* symbols per thread in DFA mode; 0 = sized to fill the GPU */
	int      own_stream;    /* 1: the scanner gets a stream of its own; each of its scans is ordered behind
	                           what the device's stream holds when it is queued.  Scans of DIFFERENT
	                           scanners of a device overlap in one way only: the scan stage of a step
	                           waits for the scan stage of the step queued before it, not for that step's
	                           prefix sum / compaction / status kernels, which run beside it on the few
	                           SMs the persistent scan kernel leaves free in this mode (ACM_RESERVE_SMS,
	                           default 8; fewer for scans beyond 1 GiB).  Two scanners used alternately
	                           (acm_scan_device_async / acm_scan_finish): 1 GiB step 0.215 -> 0.204 ms */
	int      reserved[2];
};

struct acm_scan_result {
	uint64_t n_matches;     /* records in the sorted list                                  */
	uint64_t n_bytes;       /* bytes whose matches were kept (emit_hi - emit_lo)           */
	int      mode;          /* kernel that ran                                              */
	int      fallback;      /* 1 if a bucket overflowed and the exact 2-pass path ran       */
	uint32_t final_state;   /* DFA mode only: state after the last byte (BFS numbering)    */
	uint32_t n_buckets;
	float    ms_scan;       /* timing != 0: kernel times, this call                         */
	float    ms_prefix;
	float    ms_compact;
	float    ms_total;
	uint32_t launches;      /* kernels launched by this call                                */
	uint32_t reserved;
};

int  acm_scanner_create(struct acm_device *, struct acm_automaton *, uint64_t max_bytes,
         const struct acm_scan_params *, struct acm_scanner **out);
void acm_scanner_free(struct acm_scanner *);

/*
 * Scan n symbols that are already in device memory (bytes, or ushorts for a
 * 2048-symbol automaton).  Only matches whose END offset e (in symbols, relative
 * to d_data) satisfies emit_lo <= e < emit_hi are kept, so a caller that shards a
 * stream passes Lmax-1 symbols of leading context and emit_lo = that length
 * (SURVEY.md A.5).  d_data must be 16-byte aligned and readable up to the next
 * multiple of 16 bytes past its end (true of any cudaMalloc / torch allocation; the
 * kernels load whole 16-byte vectors).  The sorted result stays on
 * the device (acm_scan_keys) until the next scan on this scanner.
 */
int  acm_scan_device(struct acm_scanner *, const void *d_data, uint64_t n,
         uint64_t emit_lo, uint64_t emit_hi, struct acm_scan_result *res);

/*
 * Same, for a buffer whose first valid_lo symbols are not part of the stream
 * (stale bytes in front of an aligned carry area): nothing before valid_lo is
 * read or used as a match start.  valid_lo <= emit_lo.
 */
int  acm_scan_device_ex(struct acm_scanner *, const void *d_data, uint64_t n, uint64_t valid_lo,
         uint64_t emit_lo, uint64_t emit_hi, struct acm_scan_result *res);

/*
 * The same scan in two halves, so that a caller can keep the GPU busy: _async queues the whole
 * step (memset, scan, prefix sum, compaction + sort, a 32-byte status readback) on the device's
 * stream and returns; acm_scan_finish waits for that readback only, runs the exact two-pass
 * path if a bucket overflowed, and fills res.  With two scanners used alternately
 * (launch i, finish i-1) there is no idle gap between steps.  One scan may be pending per scanner.
 *
 * push (may be NULL): the step also copies its sorted keys, each plus key_add, to d_dst[0 ..
 * total) -- a gather region that may live in another GPU's memory (acm_ipc_open) -- without the
 * host ever seeing the count first.  If the list does not fit (total > cap) acm_scan_finish
 * returns ACM_ERR_LIMIT; the keys stay available through acm_scan_keys.
 */
struct acm_push_target {
	uint64_t *d_dst;
	uint64_t  cap;          /* keys */
	uint64_t  key_add;
};
int  acm_scan_device_async(struct acm_scanner *, const void *d_data, uint64_t n, uint64_t valid_lo,
         uint64_t emit_lo, uint64_t emit_hi, const struct acm_push_target *push);
int  acm_scan_finish(struct acm_scanner *, struct acm_scan_result *res);

/* the cudaStream_t this scanner's scans are queued on (its own, or the device's) */
void *acm_scanner_stream(struct acm_scanner *);

/* device pointer to the sorted u64 keys of the last scan: (end offset relative to d_data) << 24 | pattern index */
const uint64_t *acm_scan_keys(struct acm_scanner *);

/*
 * Copy the last scan's matches to the host in canonical order (end offset, then
 * pattern index): h_off[i] = base + end offset relative to d_data, h_pat[i] =
 * pattern index (add order).  Returns the number of matches written (<= cap) or a
 * negative error.
 */
int64_t acm_scan_fetch(struct acm_scanner *, uint64_t base, uint64_t *h_off,
            uint32_t *h_pat, uint64_t cap);

/* tracing: with ACM_TRACE=1 in the environment at scanner creation, the scan kernel records per-CTA
 * {t_entry, t_ready, t_exit, chunks} (globaltimer ns); this copies n_ctas x 4 u64 to h_out */
int  acm_scan_trace(struct acm_scanner *, uint64_t *h_out, uint32_t n_ctas);

/*
 * Single-node gather without a collective: the collecting rank exports its gather buffer
 * (cudaMalloc'd through acm_dev_alloc) as a 64-byte CUDA IPC handle, the other ranks map it
 * and push their sorted keys straight into it over NVLink.  acm_scan_push_keys copies the last
 * scan's keys to d_dst[dst_index ...], adding key_add to each (shifts end offsets, which sit
 * above bit 24, to stream positions).  Asynchronous on the device's stream.
 */
int  acm_ipc_export(struct acm_device *, void *d_ptr, void *handle64);
int  acm_ipc_open(struct acm_device *, const void *handle64, void **d_ptr);
int  acm_ipc_close(struct acm_device *, void *d_ptr);
int  acm_scan_push_keys(struct acm_scanner *, uint64_t *d_dst, uint64_t dst_index, uint64_t key_add);

/* add the last scan's per-pattern match counts into d_counts[num_patterns] (u64, device) */
int  acm_scan_histogram(struct acm_scanner *, uint64_t *d_counts);

/*
 * End-to-end: scan a HOST buffer.  The stream is cut into segments, each copied
 * with its Lmax-1 bytes of leading context into one of two device staging
 * buffers while the previous segment is being scanned; sorted matches are
 * appended to h_off/h_pat (absolute offsets = base + position in h_data).
 * h_data should be pinned (acm_host_alloc_pinned) for full PCIe rate.
 * Returns the number of matches (may exceed cap; only cap are written).
 */
int64_t acm_scan_host(struct acm_scanner *, const void *h_data, uint64_t n, uint64_t base,
            uint64_t *h_off, uint32_t *h_pat, uint64_t cap, struct acm_scan_result *res);
/* the same for a range in the middle of a longer host stream: lead symbols in front of h_data are
 * valid context (h_data - lead is readable); matches ENDING in [0, n) are reported */
int64_t acm_scan_host_ex(struct acm_scanner *, const void *h_data, uint64_t n, uint64_t lead, uint64_t base,
            uint64_t *h_off, uint32_t *h_pat, uint64_t cap, struct acm_scan_result *res);

/*
 * ---- several GPUs in one process (no torch, no MPI): one host thread per device ----
 * What the reference gets from `-w` worker threads that each own a context on device `dev_pos`
 * (ocl_aho_grep.c:498-502, ocl_worker.c:32), for ONE stream: the automaton is replicated, device g
 * of P takes bytes [g N/P, (g+1) N/P) (cuts multiples of 16) plus Lmax-1 bytes of leading context
 * and keeps the matches that END in its own range, so the global sorted list is the concatenation
 * of the per-device lists in device order.  There is no collective: the devices
 * exchange nothing but their match counts, through host memory; every device copies its own sorted
 * list into its slice of the caller's buffer over its own PCIe link.
 * The same ordinal may be listed more than once (two "devices" on one GPU: separate streams and
 * buffers) -- that is how the logic is tested on a single-GPU box.
 */
struct acm_multi;
int   acm_multi_open(const struct acm_tables *, const int *ordinals, int n_devices, uint64_t max_bytes_per_device,
          const struct acm_scan_params *, struct acm_multi **out);
void  acm_multi_close(struct acm_multi *);
int   acm_multi_devices(const struct acm_multi *);
struct acm_device *acm_multi_device(struct acm_multi *, int g);
/* shard g of a stream of `total` symbols: device g reads [read_lo, hi) and keeps matches ending in [lo, hi) */
void  acm_multi_shard(const struct acm_multi *, uint64_t total, int g, uint64_t *read_lo, uint64_t *lo, uint64_t *hi);
/*
 * Shards already in device memory: d_data[g] holds stream[read_lo_g, hi_g) on device g (16-byte
 * aligned).  All devices scan at once; h_keys (pinned recommended) receives the global sorted
 * list, (stream end offset << 24) | pattern index, each device's part copied by that device;
 * counts[g] (may be NULL) = matches of device g.  Returns the total, or a negative error; more
 * than cap matches: the total is returned, nothing beyond cap is written.
 */
int64_t acm_multi_scan_device(struct acm_multi *, const void *const *d_data, uint64_t total,
            uint64_t *h_keys, uint64_t cap, uint64_t *counts, struct acm_scan_result *res);
/* a host stream (pinned for full rate): every device streams its own shard through its own
 * double-buffered H2D pipeline; results as acm_scan_host */
int64_t acm_multi_scan_host(struct acm_multi *, const void *h_data, uint64_t n, uint64_t base,
            uint64_t *h_off, uint32_t *h_pat, uint64_t cap, struct acm_scan_result *res);

/* ---- post-pass primitives (reference ocl_prefix_sum / ocl_compact_array / ocl_bitonic_sort) ---- */
/* exclusive prefix sum, single-pass decoupled look-back; d_total may be NULL */
int  acm_exclusive_scan_u32(struct acm_device *, const uint32_t *d_in, uint32_t *d_out,
         uint32_t n, uint32_t *d_total);
/* column-major bucket compaction (reference compactarray.cl:40-68): dst = [total, values..., tail];
 * nothing is written at or beyond d_dst[dst_capacity_ints] */
int  acm_compact_columns_i32(struct acm_device *, int32_t *d_dst, const int32_t *d_src,
         const int32_t *d_prefix, int32_t len, int32_t max_results, int64_t dst_capacity_ints);
/* LSD radix sort of u64 keys on bits [begin_bit, end_bit); d_tmp has n entries */
int  acm_radix_sort_u64(struct acm_device *, uint64_t *d_keys, uint64_t *d_tmp, uint64_t n,
         int begin_bit, int end_bit, int descending);
/* key/value u32 sort built on it; in-place when dst == src */
int  acm_sort_pairs_u32(struct acm_device *, uint32_t *d_dst_key, uint32_t *d_dst_val,
         const uint32_t *d_src_key, const uint32_t *d_src_val, uint32_t n, int descending);

/* ---- synthetic streams (bench / tests): counter-based, identical on CPU and GPU ---- */
/* byte i of the stream = byte (i & 7) of splitmix64(seed, i >> 3); fills d_dst[0..n) with stream[offset..offset+n) */
int  acm_synth_fill_device(struct acm_device *, void *d_dst, uint64_t n, uint64_t seed, uint64_t offset);
void acm_synth_fill_host(void *h_dst, uint64_t n, uint64_t seed, uint64_t offset);
/* overwrite d_buf[pos[i] - buf_offset ...] with blob[blob_off[i] .. +len[i]) for every plant that intersects the buffer */
int  acm_plant_device(struct acm_device *, void *d_buf, uint64_t n, uint64_t buf_offset,
         const uint64_t *h_pos, const uint32_t *h_blob_off, const uint32_t *h_len, uint32_t count,
         const uint8_t *h_blob, uint32_t blob_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ACM_H */
