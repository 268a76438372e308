/*
 * acm_tables.h -- host-side compiled automaton, shared between the C builder
 * (acsm_build.c) and the CUDA side (acm_cuda.cu).  Private to the library.
 *
 * States are numbered breadth-first, so ids are sorted by depth and
 * level_start[d] is the first id at depth d.  A DFA edge T[s][c] leads to a trie
 * child of s exactly when its target id is >= level_start[depth(s) + 1].
 */
#ifndef ACM_TABLES_H
#define ACM_TABLES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* transition entry: low 30 bits = next state id */
#define ACM_T_OWN   0x80000000u   /* target ends >= 1 pattern exactly here     */
#define ACM_T_ANY   0x40000000u   /* target's full match list is non-empty     */
#define ACM_T_MASK  0x3FFFFFFFu

/* record key: (end offset relative to the scan base) << 24 | pattern index */
#define ACM_KEY_PAT_BITS 24
#define ACM_KEY_PAT_MASK 0xFFFFFFu

/* scan kernels */
enum {
	ACM_MODE_AUTO     = 0,
	ACM_MODE_SAMPLED4 = 1,   /* aligned 4-gram entry filter, needs min len >= 7 */
	ACM_MODE_START2   = 2,   /* 2-byte start filter, any pattern length        */
	ACM_MODE_DFA      = 3,   /* plain per-chunk DFA walk with leading halo     */
	ACM_MODE_CDFA     = 4    /* class-compressed 16-bit DFA, hot rows in shared memory */
};

/* class-compressed DFA entry (uint16): low 14 bits next state (cd id), top 2 bits min(|full match list|, 3) */
#define ACM_CD_STATE_BITS 14
#define ACM_CD_STATE_MASK 0x3FFFu
#define ACM_CD_MAX_STATES (1u << ACM_CD_STATE_BITS)
#define ACM_CD_MAX_CLASSES 64
/* row-displaced table (k_scan_rd): entry = column | records << 5 | dense row << 8 | base << 16 */
#define ACM_RD_EMPTY       31u             /* column field of a free slot: never a real column (C <= 31) */
#define ACM_RD_MAX_DENSE   256u
#define ACM_RD_ROW         33u             /* slots between dense rows (odd: bank rotation)        */
#define ACM_RD_SMEM_BUDGET (176 * 1024)    /* bytes of shared memory the table may take (>= 16 warps still fit beside it) */

#define ACM_F1_BITS_LOG2   20              /* level-1 gram bitmap: 128 KiB smem */
#ifndef ACM_F2_WORDS
#define ACM_F2_WORDS       24576u          /* level-2 gram bitmap:  96 KiB smem, word = mulhi(hash, words) */
#endif
#define ACM_B3_WORDS       16384u          /* level-2 start bitmap:  64 KiB smem, word = hash >> 18 */
#define ACM_HASH1_MUL 0x9E3779B1u
#define ACM_HASH2_MUL 0x85EBCA6Bu
#define ACM_HASH3_MUL 0xC2B2AE35u

struct acm_gram_slot {     /* exact 4-gram table, open addressing in HBM/L2, 16 bytes */
	uint32_t gram;         /* little-endian 4 bytes of the pattern          */
	uint32_t begin1;       /* 1 + index of the gram's first candidate in cand[]; 0 = empty slot */
	uint32_t count;        /* candidates of this gram (cand[begin1 - 1 ...])  */
	uint32_t pad;
};

/*
 * candidate: pattern `id` is indexed under the 4 bytes at its byte offset o (8 bits, o mod stride
 * = the alignment this entry serves; the builder moves o from the first aligned window to a later
 * one of the same alignment when the first one's gram is popular); bit 31 of len marks the end of
 * a gram's list.  at0/at1 (pattern bytes o .. o+7, zero padded past the end) and tail (last 4
 * bytes) give a quick reject before the full compare; 32 bytes per record = two 16-byte loads.
 */
#define ACM_CAND_ID_MASK 0x00FFFFFFu
#define ACM_CAND_O_SHIFT 24
#define ACM_CAND_O_MAX   248u        /* largest indexed offset */
#define ACM_CAND_LAST    0x80000000u /* in len */
#define ACM_CAND_PAD     32          /* zeroed entries after the last list */
#define ACM_CAND_POPULAR 8           /* lists longer than this make the builder look for a rarer window */

struct acm_cand {
	uint32_t info;         /* id | o << 24 */
	uint32_t at0, at1;     /* pattern bytes o .. o+7 */
	uint32_t len;          /* | ACM_CAND_LAST */
	uint32_t tail;         /* pattern bytes len-4 .. len-1 */
	uint32_t pat_off;      /* byte offset of the pattern in pat_blob (what the verification queue stores) */
	uint32_t pad[2];
};

struct acm_tables {
	int       alpha;             /* 256 (bytes) or 2048 (ushort symbols)        */
	uint32_t  num_states;
	uint32_t  num_patterns;
	int       max_pattern_len;
	int       min_pattern_len;
	int       max_depth;         /* == max_pattern_len                           */

	uint32_t *T;                 /* [num_states][alpha]                          */
	uint32_t *level_start;       /* [max_depth + 2]; level_start[max_depth+1] == num_states */
	uint32_t *own_begin;         /* [num_states + 1] CSR into own_pat            */
	uint32_t *own_pat;           /* pattern indices ending exactly at the state  */
	uint32_t  own_total;
	uint32_t *olink;             /* [num_states] nearest proper suffix state with own patterns, 0 = none */
	uint32_t *fail;              /* [num_states] (host only)                     */
	uint32_t *pat_len;           /* [num_patterns]                               */
	int32_t  *pat_iid;           /* [num_patterns]                               */
	uint32_t *bfs_to_ref;        /* [num_states] id in the reference's numbering  */

	/* --- byte alphabet only: scan filters --- */
	int       sample_stride;     /* 0 none, 4: 4-byte grams at offsets 0..3 (min len >= 7), 8: 3-byte grams at offsets 0..7 (min len >= 10) */
	uint32_t *f1;                /* 2^20-bit bitmap of hashed pattern 4-grams at offsets 0..3 (bit-reversed words) */
	uint32_t *f2;                /* 2^19-bit second hash of the same grams        */
	struct acm_gram_slot *grams; /* exact gram table                              */
	uint32_t  gram_slots;        /* power of two                                  */
	uint32_t  gram_count;        /* distinct grams                                */
	struct acm_cand *cand;       /* candidate lists, grouped by gram, + ACM_CAND_PAD */
	uint32_t  cand_count;
	uint8_t  *pat_blob;          /* pattern bytes, each pattern 4-byte aligned, zero padded */
	uint32_t  pat_blob_bytes;
	uint32_t *pat_off;           /* [num_patterns] byte offset into pat_blob      */
	uint32_t  max_win;           /* largest entry of pat_win                       */
	uint8_t  *pat_win;           /* [num_patterns][8]: offset of the indexed window of (pattern, alignment j = (-start) mod stride) */
	int       split_len;         /* > 0: patterns shorter than this are not in the sampled filter (mixed sets) */
	uint32_t *b2s;               /* start bitmap of those short patterns alone; NULL when split_len == 0 */
	uint32_t *b3;                /* start filter of mode 2: 2^19-bit blocked Bloom bitmap (k = 3) of the first THREE
	                                bytes of every pattern (a pattern of 1 or 2 bytes enters all completions) */

	/* --- byte alphabet, <= 2^14 states, <= 63 distinct pattern bytes: class-compressed DFA --- */
	uint32_t  cd_classes;        /* C = columns per row; column C-1 = "a byte that occurs in no pattern"; 0 = not built */
	int       cd_range_lo;       /* >= 0: class(b) = min((unsigned)(b - lo), C-1) (pattern bytes span < 64 values); -1: use cd_cls */
	uint8_t  *cd_cls;            /* [256] byte -> column                          */
	uint16_t *cd_tab;            /* [num_states][C]                               */
	uint32_t *cd_flat_begin;     /* [num_states + 1] CSR into cd_flat_pat         */
	uint32_t *cd_flat_pat;       /* FULL match list of each state, ascending pattern index */
	uint32_t  cd_flat_total;
	uint32_t *cd_flat4;          /* [num_states][4]: the same lists inline (<= 4 entries): word 0 = first | count << 24 */
	uint32_t  cd_thr4;           /* an entry >= this has FOUR patterns ending (code 3 = three or four) */
	/* the same automaton as ONE row-displaced array of 4-byte entries that fits in shared memory
	 * (build_cdfa_rd in acm_core.c has the format); NULL = not built */
	uint32_t *rd_tab;            /* [rd_len] entries                              */
	uint32_t *rd_flat4;          /* [rd_len][4]: full match list of the state whose base is the slot */
	uint32_t  rd_len;            /* slots, a multiple of nothing; reads reach rd_len - 1 at most */
	/* the whole DFA as one row-displaced array of 4-byte entries (build_xd in acm_core.c has the
	 * format); NULL = not built (more slots than the entry format addresses) */
	uint32_t *xd_tab;            /* [xd_len] entries: symbol | any-match << sym_bits | base << (sym_bits + 1) */
	uint32_t *xd_sid;            /* [xd_len]: breadth-first id of the state whose base is the slot */
	uint32_t  xd_len;
	uint32_t  xd_d1_end;         /* slots [0, xd_d1_end) hold every row of the states of depth <= 1 (k_scan_xd keeps at least these in shared memory) */
	int       xd_levels;         /* 1 (bytes): miss -> row of the previous byte's depth-1 state -> root row; 0: miss -> root row */
	uint32_t  xd_sym_bits;       /* 8 or 11 */
	uint32_t  rd_dense_rows;     /* states 0 .. rd_dense_rows-1 keep whole rows, ACM_RD_ROW slots apart */
	int       rd_dense_depth;
};

void acm_tables_free(struct acm_tables *t);
size_t acm_tables_device_bytes(const struct acm_tables *t);

#ifdef __cplusplus
}
#endif
#endif
