/*
 * acm_core.h -- alphabet-generic automaton builder behind acsm_t and iacsm_t.
 * Private to the library.
 */
#ifndef ACM_CORE_H
#define ACM_CORE_H

#include "acm_tables.h"

#ifdef __cplusplus
extern "C" {
#endif

struct acm_automaton;
struct acm_device;

struct acm_pat {
	void *syms;           /* unsigned char[n] or unsigned short[n]            */
	int   n;
	int   nocase, offset, depth;
	void *id;
	int   iid;
};

struct acm_core {
	int   alpha;
	int   sym_size;       /* 1 or 2                                            */
	struct acm_pat *pats; /* by index = add order                              */
	int   npats, cap_pats;
	int   max_len, min_len;
	int   status;         /* last error, 0 = ok                                */
	int   compiled;
	struct acm_tables tab;
	struct acm_automaton *dev;   /* device copy, after upload                 */
};

struct acm_core *acm_core_new(int alpha);
int  acm_core_add(struct acm_core *, const void *syms, int n, int nocase,
         int offset, int depth, void *id, int iid);
int  acm_core_compile(struct acm_core *);
/* reference-layout table, reference numbering; *out is malloc'd (memalign 4096) */
int  acm_core_export_ref(struct acm_core *, int **out);
/* sampled-filter tables against the patterns: number of violations, -1 if no filter was built */
int  acm_core_check_filters(const struct acm_core *);
/* row-displaced dense-output table against the class-compressed one: violations, -1 if not built */
int  acm_core_check_rd(const struct acm_core *, uint32_t *slots, uint32_t *dense);
/* the row-displaced DFA (xd) against the dense table over every (state, symbol): violations, -1 if not built */
int  acm_core_check_xd(const struct acm_core *, uint32_t *slots);
/* drop host tables and pattern bytes (device copy stays) */
void acm_core_cleanup(struct acm_core *);
void acm_core_free(struct acm_core *);

void acm_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
