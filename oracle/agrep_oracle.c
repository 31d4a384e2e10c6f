/* oracle/agrep_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see agrep_oracle.h).
 *
 * A sequential, byte-at-a-time CPU restatement of the reference's scan path with 64-bit words.
 * Every function cites the reference file:line it follows.  Deliberately simple and slow.
 */
#include "agrep_oracle.h"
#include <string.h>
#include <stdio.h>
#include <ctype.h>

/* internal symbols of the preprocessed pattern (agrep.h:69-87, the non-EMX branch) */
enum { HYPHEN = 129, NOCARE = 130, NNLINE = 131, WORDB = 133, LPARENT = 134, RPARENT = 135,
       LRANGE = 136, RRANGE = 137, LANGLE = 138, RANGLE = 139, NOTSYM = 140, WILDCD = 141,
       ORSYM = 142, ORPAT = 143, ANDPAT = 144, STAR = 145 };

#define FAIL(...) do { if (err && errlen) snprintf(err, errlen, __VA_ARGS__); return -1; } while (0)

/* -i selects CP[ISO-8859-1].lower_1 as LUT (agrep.c:2769-2792; table codepage.c:399-533) and then puts every
 * byte that serves as a metasymbol back to itself (agrep.c:2835-2848: 0x83, 0x8f, 0x99 lose their lower_1 entry).
 * Restated as "identity except": ASCII A-Z -> a-z plus the irregular high half of the table the reference ends up
 * with (pinned by tests/golden/lut_lower1.json, generated from the reference, and against the binary itself in
 * tests/test_oracle_vs_reference.py). */
void orc_lut_lower1(unsigned char lut[256])
{
	static const unsigned char ex[][2] = {
		{0x80,0x87},{0x8a,0x9a},{0x8c,0x9c},{0x8e,0x9e},{0x90,0x82},
		{0x92,0x91},{0xc1,0xe1},{0xc3,0xe3},{0xc4,0xe4},{0xc5,0xe5},{0xc7,0xe7},
		{0xc8,0xe8},{0xc9,0xe9},{0xca,0xea},{0xcc,0xec},{0xcd,0xed},{0xce,0xee},{0xcf,0xef},
		{0xd1,0xf1},{0xd2,0xf2},{0xd3,0xf3},{0xd4,0xf4},{0xd5,0xf5},{0xd6,0xf6},{0xd8,0xf8},
		{0xda,0xfa},{0xdc,0xfc},{0xdd,0xfd},{0xde,0xfe} };
	int i;
	for (i = 0; i < 256; i++) lut[i] = (unsigned char)((i >= 'A' && i <= 'Z') ? i + 32 : i);
	for (i = 0; i < (int)(sizeof ex / sizeof ex[0]); i++) lut[ex[i][0]] = ex[i][1];
}

static int ascii_upper(int c) { return c >= 'A' && c <= 'Z'; }
static int ascii_alnum(int c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }

/* checksg.c:43-122: is the pattern free of meta characters?  *notsgrep: saw ^ or $ */
static int simple_pattern(const unsigned char *p, int m, int k, int *notsgrep)
{
	int i;
	*notsgrep = 0;
	for (i = 0; i < m; i++) {
		switch (p[i]) {
		case ';': case ',': case '.': case '*': case '-': case '[': case ']': case '(': case ')':
		case '<': case '>': case '|': case '#': case '{': case '}': case '~':
			return 0;
		case '^': case '$':
			*notsgrep = 1;
			return k > 0 ? 0 : 1;     /* checksg.c:80-87: goto outoffor either way */
		case '\\': i++; break;
		default: break;
		}
	}
	return 1;
}

/* preproce.c:137-341 (normal_processing, non-regex): delimiter part + ANDPAT + [-w/-x wrap] pattern,
 * meta characters mapped to the internal symbols.  out must hold 2*strlen+64 bytes. */
static int preprocess(const unsigned char *pat, const orc_opts *o, const char *dpattern /* "<X>; " or "\n; " */,
                      unsigned char *out, int *outlen, unsigned char *old_d, int *dlen, char *err, size_t errlen)
{
	unsigned char temp[1024];
	int t, m, i, j = 0, d_end, in_range = 0, L = 0;
	size_t plen = strlen((const char *)pat);
	if (plen + strlen(dpattern) + 16 > sizeof temp) FAIL("pattern too long");
	for (i = 0; i < (int)plen; i++) {          /* preproce.c:139-142: '|' or '*' means REGEX */
		if (pat[i] == '\\') i++;
		else if (pat[i] == '|' || pat[i] == '*') FAIL("regular expressions (re/re1) are outside the scan path");
	}
	strcpy((char *)temp, dpattern);
	d_end = t = (int)strlen(dpattern);
	if (o->wholeline) {                         /* preproce.c:148-159 */
		temp[t++] = LANGLE; temp[t++] = NNLINE; temp[t++] = RANGLE; temp[t] = 0;
		strcat((char *)temp, (const char *)pat);
		m = (int)strlen((char *)temp);
		temp[m++] = LANGLE; temp[m++] = '\n'; temp[m++] = RANGLE; temp[m] = 0;
	} else {                                    /* preproce.c:160-175 */
		if (o->wordbound) { temp[t++] = LANGLE; temp[t++] = WORDB; temp[t++] = RANGLE; temp[t] = 0; }
		strcat((char *)temp, (const char *)pat);
		m = (int)strlen((char *)temp);
		if (o->wordbound) { temp[m++] = LANGLE; temp[m++] = WORDB; temp[m++] = RANGLE; }
		temp[m] = 0;
	}
	for (i = 0; i < d_end - 2; i++) {           /* preproce.c:181-210: delimiter part */
		switch (temp[i]) {
		case '\\': i++; out[j++] = temp[i]; old_d[L++] = temp[i]; break;
		case '<': out[j++] = LANGLE; break;
		case '>': out[j++] = RANGLE; break;
		case '^': case '$': out[j++] = '\n'; old_d[L++] = temp[i]; break;
		default: out[j++] = temp[i]; old_d[L++] = temp[i]; break;
		}
		if (L > ORC_MAXDELIM) FAIL("delimiter pattern too long (has > %d chars)", ORC_MAXDELIM);
	}
	out[j++] = ANDPAT;                          /* preproce.c:221 */
	old_d[L] = 0;
	for (i = d_end; i < m; i++) {               /* preproce.c:238-332 */
		switch (temp[i]) {
		case '\\': i++; if (temp[i] == 0) { i = m; break; }   /* a lone backslash at the end escapes the terminator: maskgen's strlen() stops there */
			out[j++] = temp[i]; break;
		case '#': out[j++] = WILDCD; break;
		case '(': out[j++] = LPARENT; break;
		case ')': out[j++] = RPARENT; break;
		case '[': out[j++] = LRANGE; in_range = 1; break;
		case ']': out[j++] = RRANGE; in_range = 0; break;
		case '<': out[j++] = LANGLE; break;
		case '>': out[j++] = RANGLE; break;
		case '^': out[j++] = (temp[i - 1] == '[') ? NOTSYM : '\n'; break;
		case '$': out[j++] = '\n'; break;
		case '.': out[j++] = NOCARE; break;
		case '*': out[j++] = STAR; break;
		case '|': out[j++] = ORSYM; break;
		case ',': out[j++] = ORPAT; break;
		case ';': out[j++] = ANDPAT; break;
		case '-': out[j++] = in_range ? HYPHEN : temp[i]; break;
		default: out[j++] = temp[i]; break;
		}
	}
	out[j] = 0;
	*outlen = j;
	/* bitap.c:92-94 / asearch.c: '^' and '$' of old_D_pat become '\n' */
	for (i = 0; i < L; i++) if (old_d[i] == '^' || old_d[i] == '$') old_d[i] = '\n';
	*dlen = L;
	return 0;
}

/* maskgen.c:26-269 with W-bit words, LSB aligned: position p (1-based) lives at bit M-p. */
static int maskgen(unsigned char *P, int plen, int L, const orc_opts *o, int W, orc_automaton *a, char *err, size_t errlen)
{
	struct { int compl_; unsigned char cls[2 * 32 + 2]; int ncls; } pos[ORC_MAXPOS + 12];
	uint64_t wild = 0, prot = 0, sep = 0;   /* as sets of positions: bit (p) for position p, resolved after M is known */
	int i, j = 1, no_error = 0, even = 0, orflag = 0, M, k, c;
	memset(pos, 0, sizeof pos);
	a->and_mode = 0;
	if (o->nocase)                              /* maskgen.c:52-59 (C locale: ASCII only) */
		for (i = 0; i < plen; i++) if (ascii_upper(P[i])) P[i] = (unsigned char)(P[i] + 32);
	for (i = 0; i < plen; i++) {                /* maskgen.c:68-209 */
		unsigned char pp = P[i];
		if (pp == WILDCD) { if (j - 1 >= 1) wild |= 1ull << ((j - 1) & 63); }          /* :72-79 */
		else if (pp == LANGLE) { no_error = 1; even++; }                         /* :80-83 */
		else if (pp == RANGLE) { no_error = 0; even--; if (even < 0) FAIL("unmatched '<', '>'"); }
		else if (pp == LRANGE) {                                                 /* :96-127 */
			int kk = 0;
			if (no_error) prot |= 1ull << (j & 63);
			i++;
			if (P[i] == NOTSYM) { pos[j].compl_ = 1; i++; }
			while (P[i] != RRANGE && i < plen) {
				if (P[i] == HYPHEN) { if (kk > 0) pos[j].cls[kk - 1] = P[i + 1]; i += 2; }
				else { if (kk + 2 > 64) FAIL("character class too long"); pos[j].cls[kk] = pos[j].cls[kk + 1] = P[i]; kk += 2; i++; }
			}
			if (i >= plen) FAIL("unmatched '[', ']'");
			pos[j].ncls = kk;
			j++;
		}
		else if (pp == RRANGE) FAIL("unmatched '[', ']'");
		else if (pp == ORPAT) {                                                  /* :136-149 */
			if (a->and_mode) FAIL("cannot handle OR (',') and AND (';') simultaneously");
			orflag = 1; sep |= 1ull << (j & 63); j++;
		}
		else if (pp == ANDPAT) {                                                 /* :150-163 */
			if (j > L + 1) a->and_mode = 1;        /* D_length(global) == L+1 here (preproce.c:224) */
			if (orflag) FAIL("cannot handle AND (';') and OR (',') simultaneously");
			sep |= 1ull << (j & 63); j++;
		}
		else if (pp == '\n') { prot |= 1ull << (j & 63); pos[j].cls[0] = pos[j].cls[1] = '\n'; pos[j].ncls = 2; j++; }  /* :171-175 */
		else if (pp == WORDB) {                                                  /* :176-187 */
			static const unsigned char wb[8] = { 1, 47, 58, 64, 91, 96, 123, 127 };
			prot |= 1ull << (j & 63); memcpy(pos[j].cls, wb, 8); pos[j].ncls = 8; j++;
		}
		else if (pp == NNLINE) {                                                 /* :188-193 */
			prot |= 1ull << (j & 63); pos[j].cls[0] = pos[j].cls[1] = '\n'; pos[j].cls[2] = pos[j].cls[3] = NNLINE; pos[j].ncls = 4; j++;
		}
		else if (pp != STAR && pp != ORSYM && pp != LPARENT && pp != RPARENT) {  /* :194-199 */
			if (no_error) prot |= 1ull << (j & 63);
			pos[j].cls[0] = pos[j].cls[1] = pp; pos[j].ncls = 2; j++;
		}
		if (j > W) FAIL("pattern too long (has > %d chars)", W);                 /* :201-208 */
	}
	if (even != 0) FAIL("unmatched '<', '>'");
	M = j - 1;                                                                   /* :218 */
#define BITP(p) (1ull << (M - (p)))
	{
		uint64_t high = (M >= 64) ? 0 : (~0ull << M);       /* bits above the field: Init[0] |= Bit[1..W-M] (:224) */
		uint64_t wmask = (W >= 64) ? ~0ull : ((1ull << W) - 1);
		uint64_t sepbits = 0, wildbits = 0, protbits = 0, endp;
		int p;
		for (p = 1; p <= M; p++) {
			if (sep >> p & 1) sepbits |= BITP(p);
			if (wild >> p & 1) wildbits |= BITP(p);
			if (prot >> p & 1) protbits |= BITP(p);
		}
		a->wildmask = wildbits;
		a->noerr = ~protbits & wmask;                       /* :222-223 */
		a->init0 = (high | sepbits) & wmask;                /* :224-225 */
		endp = (sepbits << 1) + 1;                          /* :231 */
		a->init1 = (a->init0 | wildbits | endp) & wmask;    /* :232 */
		a->dendpos = (L >= 1 && L <= M) ? BITP(L) & endp : 0;   /* :233 keeps only the delimiter's end bit */
		a->endpos = endp ^ a->dendpos;                      /* :234 */
		a->dmask = 0;                                       /* bitap.c:131-133 */
		for (p = 0; p < L; p++) a->dmask |= a->dendpos << p;
		a->dmask = ~a->dmask & wmask;
	}
	memset(a->mask, 0, sizeof a->mask);
	for (c = 0; c < 256; c++) {                              /* :239-257 */
		for (k = 1; k <= M; k++) {
			int l, hit = 0;
			for (l = 0; l < pos[k].ncls; l += 2) {
				if (pos[k].cls[l] == NOCARE && c != '\n') { hit = 1; break; }
				if (c >= pos[k].cls[l] && c <= pos[k].cls[l + 1]) { hit = 1; break; }
			}
			if (pos[k].compl_) hit = !hit;
			if (hit) a->mask[c] |= BITP(k);
		}
	}
	if (o->nocase) for (c = 'A'; c <= 'Z'; c++) a->mask[c] = a->mask[c + 32];   /* :259-266 */
	a->M = M;
	return 0;
}

int orc_compile(const char *pattern, const orc_opts *o, orc_automaton *a, char *err, size_t errlen)
{
	unsigned char internal[1200], pat[600];
	char dpattern[64];
	int m, plen, notsgrep = 0, simple, W = o->width ? o->width : 64, jump;
	memset(a, 0, sizeof *a);
	m = (int)strlen(pattern);
	if (m < 1) FAIL("pattern length too small");
	if (m >= 256) FAIL("pattern too long");                                /* agrep.c:3057 */
	if (m <= o->k) FAIL("size of pattern must be > #of errors %d", o->k);  /* checksg.c:34 */
	if (o->wordbound && o->wholeline) FAIL("illegal option combination (-x and -w)");
	if (o->delim && o->wholeline) FAIL("-d and -x are not compatible");
	memcpy(pat, pattern, (size_t)m + 1);
	jump = (o->cost_i || o->cost_s || o->cost_d);
	a->k = o->k; a->inverse = o->inverse; a->jump = jump;
	a->ci = o->cost_i ? o->cost_i : 1; a->cs = o->cost_s ? o->cost_s : 1; a->cd = o->cost_d ? o->cost_d : 1;
	a->user_delim = o->delim != NULL;
	if (o->delim) {                                                         /* agrep.c:2272-2314 */
		size_t dl = strlen(o->delim);
		if (dl < 1 || dl > 16) FAIL("delimiter pattern too long");
		snprintf(dpattern, sizeof dpattern, "<%s>; ", o->delim);
		if (dl == 1 && (o->delim[0] == '\n' || o->delim[0] == '$' || o->delim[0] == '^')) a->outtail = 1;
	} else strcpy(dpattern, "\n; ");
	/* checksg.c:124-144 */
	simple = simple_pattern(pat, m, o->k, &notsgrep);
	a->sgrep = simple && !o->bestmatch && !(o->nocase && o->k > 0) && !jump && !o->ins_free && !o->linenum
	           && !(o->wordbound && o->k > 0) && !(o->wholeline && o->k > 0) && !notsgrep;
	if (a->sgrep && o->k > 0)
		FAIL("k>0 simple patterns use sgrep's lossy filters in the reference (SURVEY 8c); force the automaton with linenum=1");
	if (a->sgrep) {                                                         /* sgrep.c:289-320 */
		int i, n = 0;
		if (o->delim) {
			/* sgrep keeps its engine under -d (checksg.c:124-138): bm() cuts records with backward_/forward_delimiter()
			 * (sgrep.c:775-795, delim.c:52-117: plain byte search for the delimiter).  Restated for delimiters that cannot
			 * overlap themselves or the literal (anything else depends on bm()'s skip order). */
			unsigned char dl[ORC_MAXDELIM + 2]; int L = 0, b; size_t q, dn = strlen(o->delim);
			for (q = 0; q < dn; q++) {                      /* delim.c:7-29 preprocess_delimiter() */
				unsigned char c = (unsigned char)o->delim[q];
				if (c == '\\' && q + 1 < dn) c = (unsigned char)o->delim[++q];
				else if (c == '^' || c == '$') c = '\n';
				if (L >= ORC_MAXDELIM) FAIL("delimiter pattern too long (has > %d chars)", ORC_MAXDELIM);
				dl[L++] = c;
			}
			for (b = 1; b < L; b++) if (memcmp(dl, dl + L - b, (size_t)b) == 0) FAIL("oracle: sgrep -d with a self-overlapping delimiter is not restated");
			memcpy(a->dpat, dl, (size_t)L); a->L = L;
		}
		if (o->inverse) FAIL("sgrep -v: the reference counts MATCHING lines under -c (defect); not restated");
		if (o->wholeline) FAIL("oracle does not restate sgrep -x (reference defect, SURVEY 8c(8))");
		for (i = 0; i < m; i++) { if (pat[i] == '\\') i++; if (i < m) a->lit[n++] = pat[i]; }
		if (n > 20) { /* LONG_EXAC: monkey() instead of bm(); same record semantics */ }
		a->litlen = n; a->lit_word = o->wordbound; a->engine = 4;
		if (!o->delim) { a->L = 1; a->dpat[0] = '\n'; }
		else {
			int q, t;                                       /* the (folded) literal must not hold a delimiter byte */
			for (q = 0; q < n; q++) for (t = 0; t < a->L; t++) {
				int x = a->lit[q], y = a->dpat[t];
				if (ascii_upper(x)) x += 32;
				if (ascii_upper(y)) y += 32;
				if (x == y) FAIL("oracle: sgrep -d with delimiter bytes inside the literal is not restated");
			}
		}
		return 0;
	}
	if (preprocess(pat, o, dpattern, internal, &plen, a->dpat, &a->L, err, errlen)) return -1;
	if (maskgen(internal, plen, a->L, o, W, a, err, errlen)) return -1;
	if (o->ins_free || (jump && a->ci == 0)) a->init1 = (W >= 64) ? ~0ull : ((1ull << W) - 1);   /* bitap.c:123 */
	if (o->k > 0 && jump) {                                                 /* bitap.c:113-116, compat.c */
		if (a->ci <= 0 || a->cs <= 0 || a->cd <= 0) FAIL("the error cost cannot be 0");
		a->engine = 3;
	} else if (o->k > 4) a->engine = 2;                                     /* asearch.c:50-52 */
	else if (o->k > 0) a->engine = 1;
	else { a->engine = 0; a->lut_fold = o->nocase; }
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * record bookkeeping shared by all engines: bitap.c:177-229, asearch.c:119-199, agrep.c:3805-3813
 * Buffer coordinates: b[0] = virtual '\n' (bitap.c:140,149), b[1+x] = text[x], b[1+n+y] = dpat[y]
 * (bitap.c:161-165).  lasti starts at 1 (= Max_record in the reference's buffer).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
	const orc_automaton *a; uint64_t n; int64_t j; uint64_t lasti; int64_t matched;
	orc_record *recs; uint64_t cap;
} recstate;

static void rec_init(recstate *r, const orc_automaton *a, const unsigned char *text, uint64_t n, orc_record *recs, uint64_t cap)
{
	r->a = a; r->n = n; r->j = 0; r->lasti = 1; r->matched = 0; r->recs = recs; r->cap = cap;
	if (a->user_delim && n >= (uint64_t)a->L && memcmp(text, a->dpat, (size_t)a->L) == 0) r->j--;   /* bitap.c:151-156 */
}

/* called when row 0 shows D_endpos after consuming b[i-1]; i = index of the next byte */
static int rec_counts(const recstate *r, uint64_t i)
{
	int64_t print_end = (int64_t)i - r->a->L - 1;                       /* bitap.c:212 */
	return !(r->lasti >= r->n)                                          /* bitap.c:213 (file mode) */
	    && !((int64_t)r->lasti > print_end);                            /* agrep.c:3811 */
}

static void rec_close(recstate *r, uint64_t i, int cond, int level)
{
	int L = r->a->L;
	r->j++;
	if (cond && rec_counts(r, i)) {
		if (r->recs && (uint64_t)r->matched < r->cap) {
			orc_record *q = &r->recs[r->matched];
			q->begin = (int64_t)r->lasti - 1;        /* buffer index -> file offset (-1: the virtual '\n') */
			q->end = (int64_t)(i - (uint64_t)L - 1); /* print_end + 1 (exclusive), as a file offset */
			q->ordinal = r->j; q->level = level;
		}
		r->matched++;
	}
	r->lasti = i - (uint64_t)L;                                         /* bitap.c:221 */
}

static int match_cond(const orc_automaton *a, uint64_t r)
{
	/* bitap.c:182 -- note C precedence:  (AND && full) || ((!AND && any) ^ INVERSE) */
	if (a->and_mode) return ((r & a->endpos) == a->endpos) || (0 ^ (a->inverse != 0));
	return ((r & a->endpos) != 0) ^ (a->inverse != 0);
}

static inline unsigned char ext_byte(const orc_automaton *a, const unsigned char *text, uint64_t n, uint64_t i)
{
	if (i == 0) return '\n';
	if (i <= n) return text[i - 1];
	return a->dpat[i - 1 - n];
}

/* bitap.c:169-229 (k = 0) */
static int64_t scan_exact(const orc_automaton *a, const unsigned char *text, uint64_t n, orc_record *recs, uint64_t cap)
{
	recstate rs; uint64_t r = a->init0, i, end = n + 1 + (uint64_t)a->L;
	unsigned char lut[256]; int c;
	for (c = 0; c < 256; c++) lut[c] = (unsigned char)c;
	if (a->lut_fold) orc_lut_lower1(lut);
	rec_init(&rs, a, text, n, recs, cap);
	for (i = 0; i < end; ) {
		uint64_t cm = a->mask[lut[ext_byte(a, text, n, i++)]];
		r = ((r >> 1) & cm) | (a->init1 & r);                               /* :175-176 */
		if (r & a->dendpos) {                                               /* :177 */
			rec_close(&rs, i, match_cond(a, r), 0);
			r = (((a->init0 >> 1) & cm) | (a->init1 & a->init0)) & a->dmask;   /* :223-225 */
		}
	}
	return rs.matched;
}

/* asearch.c:94-199 (k<=4) and :620-697 (asearch0, k=5..8): same recurrence, array form.
 * levels_out != NULL: best-match mode, rows 0..k, report smallest matching level. */
static int64_t scan_approx(const orc_automaton *a, int k, const unsigned char *text, uint64_t n,
                           orc_record *recs, uint64_t cap, uint64_t *hist, int want_level)
{
	recstate rs; uint64_t A[ORC_MAXERR + 1], B[ORC_MAXERR + 1], i, end = n + 1 + (uint64_t)a->L; int r;
	for (r = 0; r <= k; r++) A[r] = B[r] = a->init0;
	rec_init(&rs, a, text, n, recs, cap);
	for (i = 0; i < end; ) {
		uint64_t cm = a->mask[ext_byte(a, text, n, i++)];                   /* asearch.c:96-97: no LUT */
		A[0] = ((B[0] >> 1) & cm) | (a->init1 & B[0]);                      /* :98-99 */
		for (r = 1; r <= k; r++)                                            /* :100-114 */
			A[r] = ((B[r] >> 1) & cm) | (a->init1 & B[r]) | B[r - 1] | (((A[r - 1] | B[r - 1]) >> 1) & a->noerr);
		if (A[0] & a->dendpos) {                                            /* :119 */
			if (hist) {
				int lvl = -1;
				for (r = 0; r <= k; r++) if (match_cond(a, A[r])) { lvl = r; break; }
				if (lvl >= 0) {
					/* every record enters the histogram; only those within want_level are emitted */
					if (rec_counts(&rs, i)) hist[lvl]++;
					rec_close(&rs, i, (want_level < 0) || (lvl <= want_level), lvl);
				} else rec_close(&rs, i, 0, -1);
			} else rec_close(&rs, i, match_cond(a, A[k]), k);                /* :123-128 */
			for (r = 0; r <= k; r++) B[r] = a->init0;                       /* :177-186 */
			A[0] = (((B[0] >> 1) & cm) | (B[0] & a->init1)) & a->dmask;
			for (r = 1; r <= k; r++)
				A[r] = ((B[r] >> 1) & cm) | (a->init1 & B[r]) | B[r - 1] | (((A[r - 1] | B[r - 1]) >> 1) & a->noerr);
		}
		for (r = 0; r <= k; r++) B[r] = A[r];   /* the reference ping-pongs A/B (:200-305); same thing */
	}
	return rs.matched;
}

/* asearch1.c:86-161: rows D..2D (cost 0..D), rows < D are zero */
static int64_t scan_costs(const orc_automaton *a, const unsigned char *text, uint64_t n, orc_record *recs, uint64_t cap)
{
	recstate rs; uint64_t A[2 * ORC_MAXERR + 1], B[2 * ORC_MAXERR + 1], i, end = n + 1 + (uint64_t)a->L;
	int D = a->k, r, I = a->ci, S = a->cs, DD = a->cd;
	if (DD > D) DD = D + 1;                                                /* asearch1.c:42-44 */
	if (I > D) I = D + 1;
	if (S > D) S = D + 1;
	for (r = 0; r < D; r++) A[r] = B[r] = 0;                                /* :55-56 */
	for (r = D; r <= 2 * D; r++) A[r] = B[r] = a->init0;
	rec_init(&rs, a, text, n, recs, cap);
	for (i = 0; i < end; ) {
		uint64_t cm = a->mask[ext_byte(a, text, n, i++)];
		A[D] = ((B[D] >> 1) & cm) | (a->init1 & B[D]);                      /* :90-91 */
		for (r = D + 1; r <= 2 * D; r++)                                    /* :92-97 */
			A[r] = ((B[r] >> 1) & cm) | B[r - I] | (((A[r - DD] | B[r - S]) >> 1) & a->noerr) | (a->init1 & B[r]);
		if (A[D] & a->dendpos) {                                            /* :98 */
			rec_close(&rs, i, match_cond(a, A[2 * D]), D);
			for (r = D; r <= 2 * D; r++) A[r] = B[r] = a->init0;            /* :150-158 */
			A[D] = (((B[D] >> 1) & cm) | (a->init1 & B[D])) & a->dmask;
			for (r = D + 1; r <= 2 * D; r++)
				A[r] = ((B[r] >> 1) & cm) | B[r - I] | (((A[r - DD] | B[r - S]) >> 1) & a->noerr) | (a->init1 & B[r]);
		}
		for (r = D; r <= 2 * D; r++) B[r] = A[r];
	}
	return rs.matched;
}

/* sgrep.c:262-477 + bm() :694-1016 with prep_bm :1485-1534 / char_tr :215-260, newline records:
 * a line is reported once if it contains the literal, compared under TR[] (ASCII upper->lower,
 * unconditional: sgrep.c:226-236); -w: neither neighbour isalnum (:750-755).  Every line counts at
 * most once (jump to end of record, :812,889-891). */
static int64_t scan_bm(const orc_automaton *a, const unsigned char *text, uint64_t n, orc_record *recs, uint64_t cap)
{
	uint64_t rs = 0, matched = 0, recno = 0; int m = a->litlen, L = a->L;
	int64_t begin = a->user_delim ? 0 : -1;      /* same convention as the automaton path: the delimiter that closed the record before */
	unsigned char pat[256]; int i;
	for (i = 0; i < m; i++) pat[i] = (unsigned char)(ascii_upper(a->lit[i]) ? a->lit[i] + 32 : a->lit[i]);
	while (rs < n) {
		uint64_t re = rs, p; int hit = 0;
		while (re < n && !(re + (uint64_t)L <= n && memcmp(text + re, a->dpat, (size_t)L) == 0)) re++;   /* forward_delimiter(), delim.c:52-76 */
		recno++;
		for (p = rs; !hit && p + (uint64_t)m <= re; p++) {
			for (i = 0; i < m; i++) { int c = text[p + i]; if (ascii_upper(c)) c += 32; if (c != pat[i]) break; }
			if (i < m) continue;
			if (a->lit_word) {
				int before = (p == 0) ? '\n' : text[p - 1];                     /* text[start-1]='\n' sgrep.c:393 */
				int after = (p + m < n) ? text[p + m] : '\n';                    /* sgrep.c:480 */
				if (ascii_alnum(before) || ascii_alnum(after)) continue;
			}
			hit = 1;
		}
		if (hit) {
			if (recs && matched < cap) {
				recs[matched].begin = begin; recs[matched].end = (int64_t)re; recs[matched].ordinal = (int64_t)recno; recs[matched].level = 0;
			}
			matched++;
		}
		begin = (int64_t)re;
		rs = re + (uint64_t)L;
	}
	return (int64_t)matched;
}

int64_t orc_scan(const orc_automaton *a, const unsigned char *text, uint64_t n, orc_record *recs, uint64_t cap)
{
	switch (a->engine) {
	case 0: return scan_exact(a, text, n, recs, cap);
	case 1: case 2: return scan_approx(a, a->k, text, n, recs, cap, NULL, -1);
	case 3: return scan_costs(a, text, n, recs, cap);
	case 4: return scan_bm(a, text, n, recs, cap);
	}
	return -1;
}

int64_t orc_scan_levels(const orc_automaton *a, int kmax, const unsigned char *text, uint64_t n,
                        uint64_t histogram[ORC_MAXERR + 1], orc_record *recs, uint64_t cap, int want_level)
{
	int r;
	if (a->engine > 2 || kmax < 0 || kmax > ORC_MAXERR) return -1;
	for (r = 0; r <= ORC_MAXERR; r++) histogram[r] = 0;
	return scan_approx(a, kmax, text, n, recs, cap, histogram, want_level);
}
