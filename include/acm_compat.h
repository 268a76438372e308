/*
 * acm_compat.h -- opaque stand-ins for the OpenCL handle types that appear in the
 * reference's public headers (acsmx.h:40, databuf.h:4, ocl_context.h:4 include
 * <CL/opencl.h>).  This library has no OpenCL in it: a cl_command_queue is a CUDA
 * stream owned by the library, a cl_mem is a device (or pinned host) allocation.
 * Reference-style caller code keeps compiling; nobody may dereference these.
 */
#ifndef ACM_COMPAT_H
#define ACM_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct acm_device    *cl_context;        /* one per GPU (acm.h)          */
typedef struct acm_queue     *cl_command_queue;  /* wraps one CUDA stream        */
typedef void                 *cl_mem;            /* raw device / pinned pointer  */
typedef struct acm_opaque_pl *cl_platform_id;
typedef struct acm_opaque_dv *cl_device_id;
typedef struct acm_opaque_pg *cl_program;
typedef struct acm_opaque_kn *cl_kernel;
typedef uint64_t              cl_device_type;
typedef int32_t               cl_int;
typedef uint32_t              cl_uint;
typedef int64_t               cl_long;
typedef uint64_t              cl_ulong;
typedef uint8_t               cl_uchar;

#define CL_DEVICE_TYPE_GPU (1u << 2)

#ifdef __cplusplus
}
#endif
#endif /* ACM_COMPAT_H */
