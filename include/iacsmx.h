/*
 * iacsmx.h -- Aho-Corasick over unsigned-short symbols (packet-size trains),
 * B200 edition.  Drop-in for reference AC_ushorts/iacsmx.h:43-185: alphabet of
 * 2048 symbols, patterns are unsigned short strings, matches report the
 * pattern's iid.  Same builder and kernels as acsmx.h; the scan runs the plain
 * DFA kernel over 16-bit symbols.
 *
 * Deliberate differences: the pattern copy is sized correctly (the reference
 * allocates sizeof(int) bytes, iacsmx.c:398), iacsm_add_fullpattern() accepts up
 * to 4096 items (the reference's stack array holds 16, iacsmx.c:11,423), symbols
 * >= 2048 are rejected instead of corrupting the table.
 */
#ifndef _IACSMX_H_
#define _IACSMX_H_

#include <stdio.h>
#include <stdlib.h>

#include "acm_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

#define I_ALPHABET_SIZE	2048

#ifndef ACSM_FAIL_STATE
#define ACSM_FAIL_STATE	-1
#endif

/* layout follows reference iacsmx.h:50-59 */
struct _iacsm_pattern {
	struct _iacsm_pattern	*next;
	unsigned short		*pattern;
	int			n;
	int			offset;
	int			depth;
	void			*id;
	int			iid;
};
typedef struct _iacsm_pattern iacsm_pattern_t;

struct _iacsm_state_table {
	int		next_state[I_ALPHABET_SIZE];
	int		fail_state;
	int		num_finals;
	iacsm_pattern_t	*match_list;
};
typedef struct _iacsm_state_table iacsm_state_table_t;

/* first eight fields as in reference iacsmx.h:73-83 */
struct _iacsm {
	int			max_states;
	int			num_states;
	int			max_pattern_len;
	size_t			size;
	iacsm_pattern_t		*patterns;
	iacsm_state_table_t	*state_table;
	int			*h_trans;
	cl_mem			d_trans;
	void			*priv;
};
typedef struct _iacsm iacsm_t;

iacsm_t *iacsm_new(void);                                                   /* iacsmx.c:158 */
void     iacsm_add_pattern(iacsm_t *, unsigned short *, int, int, int, void *, int); /* iacsmx.c:390 */
void     iacsm_add_fullpattern(iacsm_t *, const char *, int);               /* iacsmx.c:418: "40,32,287" */
void     iacsm_compile(iacsm_t *);                                          /* iacsmx.c:357 */
void     iacsm_gen_state_table(iacsm_t *, int, cl_context, cl_command_queue); /* iacsmx.c:455 */
int      iacsm_get_max_pattern_size(iacsm_t *);
int      iacsm_get_states(iacsm_t *);
size_t   iacsm_get_size(iacsm_t *);
void     iacsm_cleanup(iacsm_t *);
void     iacsm_free(iacsm_t *);

/* additions */
int      iacsm_status(iacsm_t *);
int      iacsm_export_ref_table(iacsm_t *);      /* int32[num_states][4096], second half = iid */
/* the row-displaced DFA (k_scan_xd) against the dense table: violations, -1 if not built.  Test support. */
int  iacsm_check_xd(iacsm_t *, unsigned int *slots);
struct acm_automaton *iacsm_device_automaton(iacsm_t *);

#ifdef __cplusplus
}
#endif
#endif /* _IACSMX_H_ */
