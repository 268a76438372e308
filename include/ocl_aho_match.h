/*
 * ocl_aho_match.h -- the match launch wrapper.  Drop-in for reference
 * ocl_aho_match.h:12-30.
 */
#ifndef _OCL_AHO_MATCH_H_
#define _OCL_AHO_MATCH_H_

#include "acm_compat.h"
#include "ocl_context.h"
#include "databuf.h"
#include "acsmx.h"
#include "iacsmx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference ocl_aho_match.c:13 JIT-compiles ahomatch.cl here; nothing to do, kernels are prebuilt */
void ocl_aho_match_init(struct clconf *c);
void ocl_aho_match_close(struct clconf *c);

/*
 * replaces reference ocl_aho_match.c:83.  (conf, databuf, automaton, local work
 * size, stream mode).  Scans bytes [0, db->bytes) of the databuf's device copy
 * and leaves the sorted match list on the device; blocks until the scan has
 * finished, like the reference's clFinish (ocl_aho_match.c:128).  local_ws is
 * accepted and ignored (launch shapes are fixed by the kernels).  stream != 0:
 * matches that began in the previous buffer are found too.
 */
void ocl_aho_match(struct clconf *, struct databuf *, acsm_t *, size_t, int);

/* ushort-symbol form (reference AC_ushorts/ocl_aho_match.h:24-25); db->h_data holds unsigned shorts */
void ocl_aho_match_ushort(struct clconf *, struct databuf *, iacsm_t *, size_t);

#ifdef __cplusplus
}
#endif
#endif /* _OCL_AHO_MATCH_H_ */
