/*
 * Minimal stand-in for <CL/opencl.h>, TEST INFRASTRUCTURE ONLY.
 *
 * The reference's DFA builder (acsmx.c, AC_ushorts/iacsmx.c) includes the
 * OpenCL headers only because acsm_gen_state_table() also uploads the table.
 * This image has no OpenCL SDK, so oracle/ref_build compiles the unmodified
 * reference sources against these declarations and ref_driver.c supplies
 * host-only definitions of the three buffer calls.  Nothing in the product
 * library includes this file.
 */
#ifndef ORACLE_STUB_CL_OPENCL_H
#define ORACLE_STUB_CL_OPENCL_H

#include <stddef.h>
#include <stdint.h>

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef int64_t  cl_long;
typedef uint64_t cl_ulong;
typedef uint8_t  cl_uchar;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_map_flags;
typedef cl_bitfield cl_device_type;

typedef struct stub_cl_platform  *cl_platform_id;
typedef struct stub_cl_device    *cl_device_id;
typedef struct stub_cl_context   *cl_context;
typedef struct stub_cl_queue     *cl_command_queue;
typedef struct stub_cl_program   *cl_program;
typedef struct stub_cl_kernel    *cl_kernel;
typedef struct stub_cl_mem       *cl_mem;
typedef struct stub_cl_event     *cl_event;

#define CL_SUCCESS              0
#define CL_TRUE                 1
#define CL_FALSE                0
#define CL_MEM_READ_WRITE       (1 << 0)
#define CL_MEM_WRITE_ONLY       (1 << 1)
#define CL_MEM_READ_ONLY        (1 << 2)
#define CL_MEM_USE_HOST_PTR     (1 << 3)
#define CL_MEM_ALLOC_HOST_PTR   (1 << 4)
#define CL_MAP_READ             (1 << 0)
#define CL_MAP_WRITE            (1 << 1)

cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
void  *clEnqueueMapBuffer(cl_command_queue, cl_mem, cl_bool, cl_map_flags,
           size_t, size_t, cl_uint, const cl_event *, cl_event *, cl_int *);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t,
           const void *, cl_uint, const cl_event *, cl_event *);

#endif
