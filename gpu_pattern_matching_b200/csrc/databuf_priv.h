/* databuf_priv.h -- what hangs off struct databuf.priv.  Private to the library. */
#ifndef DATABUF_PRIV_H
#define DATABUF_PRIV_H

#include <stdint.h>

#include "../../include/acm.h"

/* bytes reserved in front of d_data for the cross-buffer carry (>= Lmax - 1 symbols) */
#define DATABUF_CARRY_CAP 65536

struct databuf_priv {
	struct acm_device    *dev;
	struct acm_scanner   *scanner;
	struct acm_automaton *scanner_aut;   /* automaton the scanner was built for */
	unsigned char        *d_base;        /* allocation: [carry area | data]     */
	unsigned char        *d_carry_tmp;
	uint64_t              carry_len;     /* symbols currently held in the carry */
	uint64_t              n_matches;     /* of the last ocl_aho_match()         */
	uint64_t             *h_off;
	uint32_t             *h_pat;
	uint64_t              h_cap;
	int                   fetched;
	int                   status;
	int                   sym_size;      /* 1 bytes, 2 ushort symbols           */
};

#endif
