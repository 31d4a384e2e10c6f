/* agrep_b200/csrc/dropin.c -- the drop-in layer: the reference's own entry points over the B200 engine.
 *
 * Exports, with the reference's exact (K&R) signatures:
 *     bitap()    bitap.c:78      asearch()  asearch.c:32     asearch0() asearch.c:574
 *     asearch1() asearch1.c:28   sgrep()    sgrep.c:262
 *     fill_buf() bitap.c:450     alloc_buf() bitap.c:484     free_buf() bitap.c:496
 * and reads/writes the reference's globals exactly where those functions do (agrep.c:113-140), so the
 * reference's exec() (agrep.c:3332) and everything above it link against libagrepb200_dropin.so instead of
 * bitap.o asearch.o asearch1.o sgrep.o with no source change (INTEGRATION.md shows the link line).
 *
 * What happens per call: the globals maskgen() left behind become an agb_desc; the file is read into a
 * host buffer and handed to agb_scan_host() (pinned-ring H2D + the two CUDA stages); the ordered record list
 * that comes back is replayed through the reference's own output() (agrep.c:3805), which keeps every
 * formatting switch (-n -b -h -l -c -s ...) byte-identical.  Regular expressions keep going to the
 * reference's re()/re1() (agrep.c:468,1267), as bitap.c:96-111 does.
 */
#include "agrep_b200.h"
#include "agrep_b200_dropin.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <unistd.h>

#define AGREP_ERROR 123        /* agrep.h:173 */
#define SHORTREG 15            /* agrep.h:36  */

/* ---- the reference's globals (defined in agrep.c) ---- */
extern unsigned Mask[], Init1, NO_ERR_MASK, Init[], endposition, D_endpos, wildmask;
extern int AND, INVERSE, DELIMITER, D_length, I, S, DD, JUMP, REGEX, COUNT, FILENAMEONLY, SILENT;
extern int LIMITOUTPUT, LIMITPERFILE, NEW_FILE, POST_FILTER, num_of_matched, prev_num_of_matched;
extern int CurrentByteOffset, TRUNCATE, NOUPPER, WORDBOUND, WHOLELINE, LINENUM, OUTTAIL;
extern int FNAME, BYTECOUNT, PRINTOFFSET, PRINTRECORD;
extern unsigned char LUT[256];
extern char CurrentFileName[], D_pattern[], Progname[];
extern int agrep_inlen, agrep_outlen, agrep_outpointer;
extern unsigned char *agrep_inbuffer, *agrep_outbuffer;
extern FILE *agrep_finalfp;
extern int glimpse_clientdied;
extern int output();           /* agrep.c:3805 */
extern int re(), re1();        /* agrep.c:1267, 468 */

/* ---- fill_buf / alloc_buf / free_buf: same contracts (other reference files, e.g. newmgrep.c and file_out(),
 * keep calling them) ---- */
int fill_buf(int fd, unsigned char *buf, int record_size)
{
	int num_read = 1, total_read = 0;
	if (fd < 0) return 0;                                   /* bitap.c:466 */
	while (total_read < record_size && num_read > 0) {      /* bitap.c:459-463 */
		if (glimpse_clientdied) return 0;
		num_read = (int)read(fd, buf + total_read, (size_t)(record_size - total_read));
		if (num_read > 0) total_read += num_read;
	}
	if (glimpse_clientdied) return 0;
	return total_read;
}
void alloc_buf(int fd, unsigned char **buf, int size) { if (fd != -1) *buf = (unsigned char *)malloc((size_t)size); }
void free_buf(int fd, char *buf) { if (fd != -1) free(buf); }

/* whole file into hb[1..n], hb[0] = the virtual '\n' (bitap.c:140), hb[n+1..] = the delimiter (bitap.c:161-165):
 * the layout output() expects to index with (lasti, print_end) */
static unsigned char *slurp(int fd, const unsigned char *dpat, int L, size_t *n_out)
{
	size_t cap = 1u << 22, n = 0; unsigned char *hb = (unsigned char *)malloc(cap + 64);
	if (!hb) return NULL;
	hb[0] = '\n';
	if (fd == -1) {                                         /* memory mode (agrep.c:3282): the caller's leading '\n' is our virtual one */
		size_t len = agrep_inlen > 0 ? (size_t)agrep_inlen : 0, skip = (len && agrep_inbuffer[0] == '\n') ? 1 : 0;
		free(hb);
		hb = (unsigned char *)malloc(len + 64);
		if (!hb) return NULL;
		hb[0] = '\n';
		memcpy(hb + 1, agrep_inbuffer + skip, len - skip);
		n = len - skip;
	} else for (;;) {
		ssize_t r;
		if (n + 1 + 64 >= cap) { unsigned char *nb; cap *= 2; nb = (unsigned char *)realloc(hb, cap + 64); if (!nb) { free(hb); return NULL; } hb = nb; }
		r = read(fd, hb + 1 + n, cap - n - 1);
		if (r <= 0) break;
		n += (size_t)r;
	}
	memcpy(hb + 1 + n, dpat, (size_t)L);
	hb[1 + n + L] = 0;
	*n_out = n;
	return hb;
}

/* a scan that wants the record list: the list is sized from a guess and, when the scan reports more matching records
 * than fit (agb_result.truncated), once more with exactly n_matched entries -- an empty record is a record too, so no
 * bound short of one entry per text byte is safe up front (the reference prints blank lines under -v) */
static int scan_records(const agb_pattern *p, const unsigned char *text, size_t n, int want, agb_record **recs, agb_result *res)
{
	size_t cap = n / 64 + 65536; int rc;
	for (;;) {
		agb_record *r = (agb_record *)realloc(*recs, cap * sizeof **recs);
		if (!r) return AGB_ERR_NOMEM;
		*recs = r;
		rc = agb_scan_host(p, text, n, want | AGB_WANT_RECORDS, r, cap, res);
		if (rc || !res->truncated) return rc;
		cap = (size_t)res->n_matched;
	}
}

static int fail(const char *what)
{
	fprintf(stderr, "%s: %s: %s\n", Progname, what, agb_last_error());
	errno = AGREP_ERROR;
	return -1;
}

/* The anchor plan (stage 1 / 1.5 of the device scan) from the reference's INTERNAL pattern string, i.e. what
 * preprocess() hands to maskgen() and bitap() (preproce.c:221-341; symbols agrep.h:69-87): positions are counted
 * exactly as maskgen() counts them, literal runs are cut into k+1 disjoint anchors of 4 (3, 2) bytes.  Only for a
 * single pattern part without '#', not under -v or -p -- otherwise every chunk goes to the record stage. */
enum { S_HYPHEN = 129, S_NOCARE = 130, S_NNLINE = 131, S_WORDB = 133, S_LPAREN = 134, S_RPAREN = 135, S_LRANGE = 136,
       S_RRANGE = 137, S_LANGLE = 138, S_RANGLE = 139, S_NOT = 140, S_WILD = 141, S_ORSYM = 142, S_ORPAT = 143,
       S_ANDPAT = 144, S_STAR = 145 };

static void plan_from_internal(const unsigned char *P, int L, int D, agb_desc *d)
{
	int lit[80], n = 0, i, seps = 0, A, plen = (int)strlen((const char *)P);
	d->plan = AGB_PLAN_ALL; d->n_anchors = 0; d->refine = 0;
	if (INVERSE || I == 0 || wildmask) return;
	for (i = 0; i < plen && n < 70; i++) {
		int c = P[i];
		if (c == S_LANGLE || c == S_RANGLE || c == S_LPAREN || c == S_RPAREN || c == S_STAR || c == S_ORSYM) continue;
		if (c == S_WILD) return;
		if (c == S_ORPAT || c == S_ANDPAT) { if (++seps > 1) return; lit[++n] = -1; continue; }
		if (c == S_LRANGE) { while (i < plen && P[i] != S_RRANGE) i++; lit[++n] = -1; continue; }
		if (c == '\n' || c >= 128) { lit[++n] = -1; continue; }                 /* newline, WORDB, NNLINE, NOCARE: a position, not a literal */
		lit[++n] = NOUPPER && c >= 'A' && c <= 'Z' ? c + 32 : c;
		/* (bytes >= 128 were excluded above: under -i the exact engine folds them through LUT[], bitap.c:171) */
	}
	if (n != d->M) return;                                                          /* our count must agree with maskgen's */
	for (A = 4; A >= 2; A--) {
		int got = 0, run = 0, p;
		for (p = L + 2; p <= n && got < D + 1; p++) {
			if (lit[p] < 0) { run = 0; continue; }
			if (++run == A) {
				uint32_t v = 0; int t;
				for (t = 0; t < A; t++) v |= (uint32_t)lit[p - A + 1 + t] << (8 * t);
				d->anchor[got] = v; d->anchor_off[got] = p - A + 1 - (L + 2); got++; run = 0;
			}
		}
		if (got < D + 1) continue;
		d->plan = AGB_PLAN_ANCHORS; d->n_anchors = D + 1; d->anchor_len = A;
		d->anchor_mask = A == 4 ? 0xFFFFFFFFu : (A == 3 ? 0x00FFFFFFu : 0x0000FFFFu);
		d->anchor_fold = NOUPPER ? 0x20202020u : 0;
		for (p = 0; p < d->n_anchors; p++) d->anchor[p] = (d->anchor[p] | d->anchor_fold) & d->anchor_mask;
		d->pat_len = n - L - 1; d->refine = 1;
		return;
	}
}

/* the common tail of bitap()/asearch*(): desc from the globals, scan, replay through output() */
static int scan_and_replay(char old_D_pat[], const unsigned char *Pattern, int fd, int M, int D, int engine)
{
	agb_desc d; agb_pattern *p = NULL; agb_result res; agb_record *recs = NULL;
	unsigned char dpat[2 * AGB_MAXDELIM + 2], *hb; size_t n = 0, i; int L, c, rc, ret = 0;
	char err[256];
	const uint64_t HI = 0xFFFFFFFF00000000ull;

	L = (int)strlen(old_D_pat);
	if (L < 1 || L > AGB_MAXDELIM) { fprintf(stderr, "%s: delimiter pattern too long\n", Progname); errno = AGREP_ERROR; return -1; }
	for (c = 0; c < L; c++) {                               /* bitap.c:92-94 */
		if (old_D_pat[c] == '^' || old_D_pat[c] == '$') old_D_pat[c] = '\n';
		dpat[c] = (unsigned char)old_D_pat[c];
	}
	D_length = L;
	if (I == 0) Init1 = 037777777777u;                      /* bitap.c:123, asearch.c:49, asearch1.c:41 */

	memset(&d, 0, sizeof d);
	for (c = 0; c < 256; c++) d.mask[c] = Mask[engine == AGB_ENGINE_BITAP ? LUT[c] : c];   /* bitap.c:171 vs asearch.c:96 */
	d.init0 = HI | Init[0]; d.init1 = HI | Init1; d.noerr = HI | NO_ERR_MASK;
	d.endpos = endposition; d.dendpos = D_endpos; d.wildmask = wildmask;
	d.dmask = 0;
	for (c = 0; c < L; c++) d.dmask |= (uint64_t)D_endpos << c;             /* bitap.c:131-133 */
	d.dmask = ~d.dmask;
	d.M = M; d.L = L; memcpy(d.delim, dpat, (size_t)L);
	d.k = D; d.engine = engine; d.and_mode = AND; d.inverse = INVERSE; d.user_delim = DELIMITER; d.outtail = OUTTAIL;
	d.cost_i = I > D ? D + 1 : I; d.cost_s = S > D ? D + 1 : S; d.cost_d = DD > D ? D + 1 : DD;   /* asearch1.c:42-44 */
	if (d.cost_i < 1) d.cost_i = 1;
	if (Pattern) plan_from_internal(Pattern, L, D, &d);
	rc = agb_pattern_from_desc(&d, &p, err, sizeof err);
	if (rc) { fprintf(stderr, "%s: %s\n", Progname, err); errno = AGREP_ERROR; return -1; }

	hb = slurp(fd, dpat, L, &n);
	if (!hb) { agb_pattern_free(p); fprintf(stderr, "%s: out of memory\n", Progname); errno = AGREP_ERROR; return -1; }

	if (COUNT && !FILENAMEONLY && fd != -1) {               /* output() would only count (agrep.c:3812-3813) */
		rc = agb_scan_host(p, hb + 1, n, AGB_WANT_COUNT, NULL, 0, &res);
		if (rc) ret = fail("scan");
		else num_of_matched += (int)res.n_matched;
		goto done;
	}
	rc = scan_records(p, hb + 1, n, LINENUM ? AGB_WANT_ORDINALS : 0, &recs, &res);   /* j for output() (-n) on the device */
	if (rc) { ret = fail("scan"); goto done; }
	for (i = 0; i < res.n_records; i++) {
		if (fd == -1 && recs[i].end >= (long long)n) continue;      /* memory mode appends no delimiter (bitap.c:310-314) */
		if (COUNT && !FILENAMEONLY) { num_of_matched++; continue; } /* (memory mode: the count leaves that last record out too) */
		if (FILENAMEONLY && (NEW_FILE || !POST_FILTER)) {       /* bitap.c:184-210 */
			num_of_matched++;
			if (agrep_finalfp != NULL) fprintf(agrep_finalfp, "%s\n", CurrentFileName);
			else {
				size_t fl = strlen(CurrentFileName);
				if (agrep_outpointer + (int)fl + 1 >= agrep_outlen) { ret = -1; break; }
				memcpy(agrep_outbuffer + agrep_outpointer, CurrentFileName, fl);
				agrep_outbuffer[agrep_outpointer + fl] = '\n';
				agrep_outpointer += (int)fl + 1;
			}
			NEW_FILE = 0;
			break;
		}
		/* CurrentByteOffset as the loop leaves it at the output() call (bitap.c:172,179): bytes consumed minus the delimiter */
		CurrentByteOffset = (int)(recs[i].end + 1);
		TRUNCATE = 0;
		if (-1 == output(hb, (int)(recs[i].begin + 1), (int)recs[i].end, (int)recs[i].ordinal)) { ret = -1; break; }
		if ((LIMITOUTPUT > 0 && LIMITOUTPUT <= num_of_matched) ||
		    (LIMITPERFILE > 0 && LIMITPERFILE <= num_of_matched - prev_num_of_matched)) break;     /* bitap.c:215-219 */
	}
done:
	free(recs); free(hb); agb_pattern_free(p);
	return ret;
}

int bitap(char old_D_pat[], char *Pattern, int fd, int M, int D)
{
	if (REGEX) {                                            /* bitap.c:96-111: stays with the reference's NFA code */
		if (D > 4) { fprintf(stderr, "%s: the maximum number of erorrs allowed for full regular expressions is 4\n", Progname); errno = AGREP_ERROR; return -1; }
		D_length = (int)strlen(old_D_pat);
		return M <= SHORTREG ? re(fd, M, D) : re1(fd, M, D);
	}
	if (D > 0 && JUMP == 1) return scan_and_replay(old_D_pat, (const unsigned char *)Pattern, fd, M, D, AGB_ENGINE_ASEARCH1);   /* bitap.c:113-116 */
	if (D > 4) return scan_and_replay(old_D_pat, (const unsigned char *)Pattern, fd, M, D, AGB_ENGINE_ASEARCH0);               /* asearch.c:50-52 */
	if (D > 0) return scan_and_replay(old_D_pat, (const unsigned char *)Pattern, fd, M, D, AGB_ENGINE_ASEARCH);                /* bitap.c:118-121 */
	return scan_and_replay(old_D_pat, (const unsigned char *)Pattern, fd, M, D, AGB_ENGINE_BITAP);
}

/* M is not a parameter of these three in the reference; it is recovered from the always-on bits of Init[0] */
static int positions_from_init0(void)
{
	unsigned v = ~Init[0]; int M = 0;           /* bits >= M of Init[0] are ones (maskgen.c:224) */
	while (v) { M++; v >>= 1; }
	return M;
}
int asearch(unsigned char old_D_pat[], int text, unsigned D)
{ return scan_and_replay((char *)old_D_pat, NULL, text, positions_from_init0(), (int)D, D > 4 ? AGB_ENGINE_ASEARCH0 : AGB_ENGINE_ASEARCH); }
int asearch0(unsigned char old_D_pat[], int text, unsigned D)
{ return scan_and_replay((char *)old_D_pat, NULL, text, positions_from_init0(), (int)D, AGB_ENGINE_ASEARCH0); }
int asearch1(char old_D_pat[], int Text, unsigned D)
{ return scan_and_replay(old_D_pat, NULL, Text, positions_from_init0(), (int)D, AGB_ENGINE_ASEARCH1); }

/* sgrep(): simple patterns (checksg.c:138).  k = 0: bm() semantics (ASCII case folded literal, once per record,
 * -w by isalnum neighbours, sgrep.c:741-755).  k > 0: the reference runs lossy filters here;
 * we run the exact automaton.  Record printing follows bm()/s_output() (sgrep.c:812-933, 1274-1483). */
int sgrep(unsigned char *in_pat, int in_m, int fd, int D, int samepattern)
{
	agb_options o; agb_pattern *p = NULL; agb_result res; agb_record *recs = NULL;
	char err[256], pat[1024], delim[64]; unsigned char *hb; size_t n = 0, i; int rc, ret = 0, L;
	unsigned char dpat[2 * AGB_MAXDELIM + 2];
	(void)samepattern;
	memset(&o, 0, sizeof o);
	if (in_m < 1 || in_m >= (int)sizeof pat) { errno = AGREP_ERROR; return -1; }
	memcpy(pat, in_pat, (size_t)in_m); pat[in_m] = 0;
	o.k = D; o.wordbound = WORDBOUND; o.wholeline = WHOLELINE; o.inverse = (INVERSE && !COUNT) ? 1 : 0; o.nocase = NOUPPER;
	if (DELIMITER) {
		/* D_pattern holds the delimiter bytes themselves here (agrep.c:3182-3185); re-escape for agb_compile */
		int q = 0, t;
		for (t = 0; t < D_length && q + 2 < (int)sizeof delim; t++) { delim[q++] = '\\'; delim[q++] = D_pattern[t]; }
		delim[q] = 0; o.delim = delim;
	}
	rc = agb_compile(pat, &o, &p, err, sizeof err);
	if (rc) { fprintf(stderr, "%s: %s\n", Progname, err); errno = AGREP_ERROR; return -1; }
	L = agb_pattern_desc(p)->L; memcpy(dpat, agb_pattern_desc(p)->delim, (size_t)L);
	hb = slurp(fd, dpat, L, &n);
	if (!hb) { agb_pattern_free(p); errno = AGREP_ERROR; return -1; }
	if (COUNT || SILENT || (FILENAMEONLY && !INVERSE)) {
		/* bm() only counts matching records here, also under -v (sgrep.c:813-815, 968-971) */
		rc = agb_scan_host(p, hb + 1, n, AGB_WANT_COUNT, NULL, 0, &res);
		if (rc) ret = fail("scan");
		else if (FILENAMEONLY) {
			if (res.n_matched) {
				num_of_matched++;                                       /* bm() returns at the first hit (sgrep.c:813-814) */
				if (NEW_FILE || !POST_FILTER) {                         /* sgrep.c:443-463 */
					if (agrep_finalfp != NULL) fprintf(agrep_finalfp, "%s\n", CurrentFileName);
					NEW_FILE = 0;
				}
			}
		} else num_of_matched += (int)res.n_matched;
		goto done;
	}
	rc = scan_records(p, hb + 1, n, 0, &recs, &res);
	if (rc) { ret = fail("scan"); goto done; }
	for (i = 0; i < res.n_records; i++) {
		/* bm() prints [curtextbegin, curtextend): the line and its trailing newline (sgrep.c:775-789, 916); with a user
		 * delimiter the record together with the delimiter in FRONT of it, or behind it under -t
		 * (backward_/forward_delimiter(), delim.c:52-117) */
		const int tail = !DELIMITER || OUTTAIL;
		long long b = recs[i].begin < 0 ? 0 : (tail && !(DELIMITER && i == 0 && recs[i].begin == 0 && !(n >= (size_t)L && memcmp(hb + 1, dpat, (size_t)L) == 0)) ? recs[i].begin + L : recs[i].begin);
		long long e = tail ? recs[i].end + L : recs[i].end;
		if ((size_t)e > n) e = (long long)n;
		if (!INVERSE) num_of_matched++;
		if (agrep_finalfp != NULL) {
			if (FNAME && (NEW_FILE || !POST_FILTER)) { fprintf(agrep_finalfp, "%s: ", CurrentFileName); }
			if (BYTECOUNT) fprintf(agrep_finalfp, "%d= ", (int)recs[i].end);
			if (PRINTRECORD) {
				fwrite(hb + 1 + b, 1, (size_t)(e - b), agrep_finalfp);
				if ((size_t)e == n && n && hb[n] != dpat[L - 1]) fputc('\n', agrep_finalfp);   /* sgrep.c:786-789 */
			} else if (FNAME || BYTECOUNT) fputc('\n', agrep_finalfp);
		} else {
			if (agrep_outpointer + (int)(e - b) + 1 >= agrep_outlen) { ret = -1; break; }
			memcpy(agrep_outbuffer + agrep_outpointer, hb + 1 + b, (size_t)(e - b));
			agrep_outpointer += (int)(e - b);
		}
		if ((LIMITOUTPUT > 0 && LIMITOUTPUT <= num_of_matched) ||
		    (LIMITPERFILE > 0 && LIMITPERFILE <= num_of_matched - prev_num_of_matched)) break;
	}
	if (INVERSE) {
		/* bm() counts the MATCHING records even when it prints the others (sgrep.c:813); keep num_of_matched faithful */
		agb_options o2 = o; agb_pattern *p2 = NULL; agb_result r2;
		o2.inverse = 0;
		if (agb_compile(pat, &o2, &p2, err, sizeof err) == 0) {
			if (agb_scan_host(p2, hb + 1, n, AGB_WANT_COUNT, NULL, 0, &r2) == 0) num_of_matched += (int)r2.n_matched;
			agb_pattern_free(p2);
		}
	}
done:
	free(recs); free(hb); agb_pattern_free(p);
	return ret;
}
