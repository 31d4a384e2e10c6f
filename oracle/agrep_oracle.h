/* oracle/agrep_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, 64-bit state words) of the Wu-Manber scan path of
 * Wikinaut/agrep 3.41.5: pattern front-end (preproce.c, maskgen.c), the exact
 * shift-and loop (bitap.c), the k-error loops (asearch.c: asearch/asearch0),
 * the non-unit-cost loop (asearch1.c) and the simple-literal engine
 * (sgrep.c: sgrep()+bm()).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may call into this file.
 * The product (agrep_b200/) never links or imports it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this
 * restatement against the unmodified reference built by oracle/Makefile
 * (oracle/_ref) and against tests/golden/ fixtures generated from it.
 */
#ifndef AGREP_ORACLE_H
#define AGREP_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXPOS 64          /* word width W of the restatement (reference: WORD=32, agrep.h:44) */
#define ORC_MAXERR 8           /* MaxError, agrep.h:45 */
#define ORC_MAXDELIM 8         /* MAXDELIM, agrep.h:35 */

/* options = the subset of agrep's command line that reaches the scan path (agrep.c:2121-2739) */
typedef struct {
	int k;             /* -# : number of errors D                                   */
	int nocase;        /* -i : NOUPPER                                              */
	int wordbound;     /* -w : WORDBOUND                                            */
	int wholeline;     /* -x : WHOLELINE                                            */
	int inverse;       /* -v : INVERSE                                              */
	int linenum;       /* -n : LINENUM (forces the bitap family, checksg.c:132)     */
	int ins_free;      /* -p : I = 0  (Init1 := all ones, bitap.c:123)              */
	int cost_i, cost_s, cost_d;  /* -I# -S# -D#; 0 = not given (JUMP stays off)     */
	int bestmatch;     /* -B : BESTMATCH (forces the bitap family, checksg.c:127)   */
	int width;         /* 0 -> 64.  32 reproduces the reference's "pattern too long" limit (maskgen.c:201) */
	const char *delim; /* -d argument as typed (NULL = default newline records)     */
} orc_opts;

typedef struct {
	uint64_t mask[256];        /* Mask[c]        (maskgen.c:239-266)                */
	uint64_t init0;            /* Init[0]        (maskgen.c:224-225)                */
	uint64_t init1;            /* Init1          (maskgen.c:232; all-ones if -p)    */
	uint64_t noerr;            /* NO_ERR_MASK    (maskgen.c:222-223)                */
	uint64_t endpos;           /* endposition    (maskgen.c:231,234)                */
	uint64_t dendpos;          /* D_endpos       (maskgen.c:233)                    */
	uint64_t dmask;            /* D_Mask         (bitap.c:131-133)                  */
	uint64_t wildmask;         /* wildmask       (maskgen.c:78,220)                 */
	int M;                     /* number of automaton positions                     */
	int L;                     /* strlen(old_D_pat)                                 */
	unsigned char dpat[2 * ORC_MAXDELIM + 2]; /* old_D_pat with ^,$ -> '\n'          */
	int and_mode;              /* AND            (maskgen.c:153)                    */
	int user_delim;            /* DELIMITER                                         */
	int outtail;               /* OUTTAIL (agrep.c:2290,2307)                       */
	/* dispatch decisions (checksg.c:124-144, bitap.c:96-121) */
	int sgrep;                 /* 1: simple-pattern engine sgrep()/bm() (k must be 0 here) */
	int engine;                /* 0 bitap-exact, 1 asearch, 2 asearch0, 3 asearch1, 4 sgrep/bm */
	int k, inverse, jump, ci, cs, cd;
	int lut_fold;              /* bitap exact applies LUT[] (-i: ISO-8859-1 lower_1) before Mask[] (bitap.c:171) */
	/* sgrep/bm literal (sgrep.c:289-320) */
	unsigned char lit[256]; int litlen; int lit_word;
} orc_automaton;

typedef struct {
	int64_t  begin;   /* file offset of lasti   (first byte of the delimiter that closed the previous record; -1 = virtual '\n') */
	int64_t  end;     /* file offset one past print_end (= first byte of this record's closing delimiter)          */
	int64_t  ordinal; /* j at the call of output() (agrep.c:3805); -n prints j-1 (+1 when DELIMITER)                */
	int      level;   /* best-match mode only: smallest error level that matched, else -1                           */
} orc_record;

/* returns 0 ok; -1 error with message in err (pattern too long, unsupported metachar, ...) */
int orc_compile(const char *pattern, const orc_opts *o, orc_automaton *a, char *err, size_t errlen);

/* Scan n bytes exactly as the reference scans a FILE (virtual '\n' in front, delimiter appended at EOF,
 * phantom last record suppressed); appends matched records to recs[0..cap) when recs != NULL.
 * Returns the number of matched records (num_of_matched). */
int64_t orc_scan(const orc_automaton *a, const unsigned char *text, uint64_t n,
                 orc_record *recs, uint64_t cap);

/* one pass with kmax rows that reports for every record the smallest level 0..kmax whose end bit
 * is set (the nesting A_j >= A_{j-1} of asearch.c:98-114); histogram[lvl] counts records. Used to
 * restate the -B sweep of agrep.c:3582-3728 in one pass. */
int64_t orc_scan_levels(const orc_automaton *a, int kmax, const unsigned char *text, uint64_t n,
                        uint64_t histogram[ORC_MAXERR + 1], orc_record *recs, uint64_t cap, int want_level);

/* ISO-8859-1 lower_1 table as selected by -i (agrep.c:2769-2792, codepage.c) */
void orc_lut_lower1(unsigned char lut[256]);

#ifdef __cplusplus
}
#endif
#endif
