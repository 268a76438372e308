/*
 * acsmx.h -- byte-alphabet Aho-Corasick builder, B200 edition.
 *
 * Drop-in for the reference's acsmx.h (reference acsmx.h:51-196): same type names,
 * same field names in acsm_t / acsm_pattern_t, same functions with the same
 * arity.  What is behind them is new: an array-based builder that emits a
 * breadth-first-numbered automaton plus the filter tables the sm_100a scan
 * kernels use, and an upload through the CUDA runtime instead of OpenCL.
 *
 * Differences a caller can observe (all deliberate):
 *   - acsm_t.state_table is always NULL (the reference's 1040 B/state build
 *     array, acsmx.h:67-73, is never materialised);
 *   - acsm_t.h_trans is NULL unless acsm_export_ref_table() is called;
 *   - acsm_get_size() is the real device footprint, not 2 KiB x states;
 *   - acsm_get_patterns_table() keeps embedded NUL bytes and leaves .next NULL
 *     (the reference builds cyclic chains there, acsmx.c:707-721);
 *   - zero-length patterns are rejected (return silently ignored + error text),
 *     the reference accepts them and then cannot report them (acsmx.c:640-650).
 */
#ifndef _ACSMX_H_
#define _ACSMX_H_

#include <stdio.h>
#include <stdlib.h>

#include "acm_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

/* byte alphabet (reference acsmx.h:44) */
#define ALPHABET_SIZE	256

/* reference acsmx.h:47 */
#define ACSM_FAIL_STATE	-1

/* one pattern; layout follows reference acsmx.h:51-63 */
struct _acsm_pattern {
	struct _acsm_pattern	*next;
	unsigned char		*pattern;
	unsigned char		*casepattern;
	int			n;
	int			nocase;
	int			offset;
	int			depth;
	void			*id;
	int			iid;
	unsigned int		index;
};
typedef struct _acsm_pattern acsm_pattern_t;

/* kept only so that code naming the type still compiles (reference acsmx.h:67-73) */
struct _acsm_state_table {
	int		next_state[ALPHABET_SIZE];
	int		fail_state;
	int		num_finals;
	acsm_pattern_t	*match_list;
};
typedef struct _acsm_state_table acsm_state_table_t;

/* the automaton; first nine fields as in reference acsmx.h:77-88 */
struct _acsm {
	int			max_states;
	int			num_states;
	int			max_pattern_len;
	size_t			size;
	acsm_pattern_t		*patterns;
	int			num_patterns;
	acsm_state_table_t	*state_table;
	int			*h_trans;
	cl_mem			d_trans;
	void			*priv;		/* builder + device state */
};
typedef struct _acsm acsm_t;

/* replaces reference acsmx.c:495 */
acsm_t *acsm_new(void);

/*
 * replaces reference acsmx.c:514.  (acsm, bytes, n, nocase, offset, depth, id,
 * iid); bytes are copied; nocase/offset/depth are stored and ignored, as in the
 * reference (case folding is disabled there, acsmx.c:265-275).
 */
void acsm_add_pattern(acsm_t *, unsigned char *, int, int, int, int, void *, int);

/* replaces reference acsmx.c:552 */
void acsm_compile(acsm_t *);

/*
 * replaces reference acsmx.c:600: builds the device tables and uploads them on
 * the queue's stream.  `mapped` is accepted and ignored (no zero-copy tables).
 * ctx/queue may be NULL: the library's default device 0 context is used.
 */
void acsm_gen_state_table(acsm_t *, int, cl_context, cl_command_queue);

/* replaces reference acsmx.c:677; caller owns the result (free with acsm_free_patterns_table) */
acsm_pattern_t *acsm_get_patterns_table(acsm_t *acsm);
void acsm_free_patterns_table(acsm_pattern_t *, int num_patterns);

/* replace reference acsmx.c:741-765 */
int    acsm_get_max_pattern_size(acsm_t *);
int    acsm_get_states(acsm_t *);
size_t acsm_get_size(acsm_t *);

/* replaces reference acsmx.c:771: drops host build structures, keeps device tables */
void acsm_cleanup(acsm_t *);

/* replaces reference acsmx.c:809; also releases the device tables (the reference leaks them) */
void acsm_free(acsm_t *);

/* ---- additions (not in the reference) ---- */

/* 0 if the last add/compile/upload succeeded, else a negative ACM_ERR_* (acm.h) */
int  acsm_status(acsm_t *);

/* shortest pattern length (drives kernel selection) */
int  acsm_get_min_pattern_size(acsm_t *);

/*
 * Fills acsm->h_trans with the table in the REFERENCE layout and numbering,
 * int32[num_states][512] (reference acsmx.c:640-659), for table-parity tests and
 * for callers that want to read it.  Must be called after acsm_compile() and
 * before acsm_cleanup().  Returns 0 or a negative error.
 */
int  acsm_export_ref_table(acsm_t *);

/*
 * Host-side self-check of the scan-filter tables built by acsm_compile() for byte patterns of
 * length >= 7 (bitmaps, exact table, candidate records, pattern blob) against the patterns:
 * number of violations (0 = consistent), -1 if this automaton has no filter.  Test support.
 */
int  acsm_check_filters(acsm_t *);

/*
 * Word lists (<= 16 384 states, pattern bytes within one range of < 31 values, <= 4 patterns ending
 * in a state) also get a row-displaced transition table that lives in shared memory (k_scan_rd).
 * Checks it against the dense automaton over every (state, byte class): number of violations
 * (0 = identical), -1 if this automaton has none.  *slots / *dense_rows (may be NULL) report its
 * shape.  Test support.
 */
int  acsm_check_cdfa(acsm_t *, unsigned int *slots, unsigned int *dense_rows);

/*
 * Every automaton also gets its DFA as one small row-displaced array (k_scan_xd walks it; 6.8 MB
 * instead of the 370 MiB dense table for 10 000 ClamAV signatures).  Checks it against the dense
 * table over every (state, symbol): violations (0 = identical), -1 if none was built.  Test support.
 */
int  acsm_check_xd(acsm_t *, unsigned int *slots);

/* the compiled host tables (acm_automaton_upload, acm_multi_open); NULL before acsm_compile and after acsm_cleanup */
struct acm_tables;
const struct acm_tables *acsm_tables(acsm_t *);

/* device automaton handle for the native API in acm.h (NULL before upload) */
struct acm_automaton *acsm_device_automaton(acsm_t *);

#ifdef __cplusplus
}
#endif
#endif /* _ACSMX_H_ */
