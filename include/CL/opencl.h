/*
 * CL/opencl.h -- shim so that UNMODIFIED reference sources that include <CL/opencl.h>
 * (reference ocl_aho_grep.c:10, acsmx.h:40, databuf.h:4, ocl_context.h:4) compile against
 * this library's headers with -I include.  There is no OpenCL here: the handle types are the
 * opaque stand-ins of acm_compat.h, nothing else of the OpenCL API exists.
 */
#ifndef ACM_CL_OPENCL_SHIM_H
#define ACM_CL_OPENCL_SHIM_H
#include "../acm_compat.h"
#endif
