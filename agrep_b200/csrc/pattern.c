/* agrep_b200/csrc/pattern.c -- host-side pattern front-end of libagrepb200 (C, no CUDA).
 *
 * Turns (pattern, options) into the scan descriptor agb_desc the kernels consume.  It mirrors what the
 * reference does on the host before a scan -- checksg() (checksg.c:19-165: which engine), preprocess()
 * (preproce.c:137-341: delimiter + separator + -w/-x wrap + meta characters) and maskgen()
 * (maskgen.c:26-269: Mask[], Init[0], Init1, NO_ERR_MASK, endposition, D_endpos, wildmask) -- but is
 * organised as one pass over the user's pattern that emits automaton positions with 256-bit classes,
 * in 64-bit words (the reference stops at 32 positions, maskgen.c:201-208).
 *
 * It also derives what only the device path needs: the constant post-delimiter rows (asearch.c:175-186),
 * the delimiter kind, and the pigeonhole anchor plan for the front-end kernel.
 */
#include "agrep_b200.h"
#include "pattern_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WIDTH 64
#define FAIL(...) do { if (err && errlen) snprintf(err, errlen, __VA_ARGS__); return AGB_ERR_PATTERN; } while (0)

/* internal symbol codes of the reference (agrep.h:69-87); raw pattern bytes in this range are refused */
enum { S_HYPHEN = 129, S_NOCARE = 130, S_NNLINE = 131, S_WORDB = 133, S_LPAREN = 134, S_RPAREN = 135,
       S_LRANGE = 136, S_RRANGE = 137, S_LANGLE = 138, S_RANGLE = 139, S_NOT = 140, S_WILD = 141,
       S_ORSYM = 142, S_ORPAT = 143, S_ANDPAT = 144, S_STAR = 145 };

typedef struct {
	uint64_t cls[4];     /* 256-bit character class */
	int is_sep;          /* ORPAT / ANDPAT slot: empty class, always-on start bit (maskgen.c:136-163) */
	int prot;            /* NO_ERR_MASK position (maskgen.c:97,172,177,189,195) */
	int wild;            /* '#' after this position: sticky self loop (maskgen.c:72-79) */
	int lit;             /* single literal byte (anchor eligible), else -1 */
	int part;            /* index of the ','/';' sub-pattern this position belongs to */
} pos_t;

typedef struct {
	pos_t p[WIDTH + 4];
	int n;               /* positions so far (1-based: p[1..n]) */
	int no_error, even;
	int or_seen, and_mode, nparts;
} build_t;

static void cls_set(pos_t *p, int c) { p->cls[c >> 6] |= 1ull << (c & 63); }
static int  cls_has(const pos_t *p, int c) { return (int)(p->cls[c >> 6] >> (c & 63) & 1); }
static void cls_range(pos_t *p, int lo, int hi)
{
	int c;
	if (lo == S_NOCARE) for (c = 0; c < 256; c++) if (c != '\n') cls_set(p, c);   /* maskgen.c:243-247 */
	for (c = lo; c <= hi && c < 256; c++) cls_set(p, c);                         /* maskgen.c:248-252 */
}
static int is_upper(int c) { return c >= 'A' && c <= 'Z'; }
static int is_alpha(int c) { return (c | 32) >= 'a' && (c | 32) <= 'z'; }
static int is_alnum(int c) { return is_alpha(c) || (c >= '0' && c <= '9'); }

static pos_t *new_pos(build_t *b)
{
	pos_t *p;
	if (b->n + 1 > WIDTH - 1) return NULL;     /* M <= W-1: one always-on feed bit above the field */
	p = &b->p[++b->n];
	memset(p, 0, sizeof *p);
	p->lit = -1;
	p->part = b->nparts;
	return p;
}

/* one user-pattern character -> the reference's internal symbol (preproce.c:238-332) */
static int map_sym(const unsigned char *s, int *i, int n, int in_range, int *escaped)
{
	int c = s[*i];
	*escaped = 0;
	if (c == '\\') { (*i)++; *escaped = 1; return (*i < n) ? s[*i] : 0; }
	switch (c) {
	case '#': return S_WILD;   case '(': return S_LPAREN; case ')': return S_RPAREN;
	case '[': return S_LRANGE; case ']': return S_RRANGE; case '<': return S_LANGLE; case '>': return S_RANGLE;
	case '^': return (*i > 0 && s[*i - 1] == '[') ? S_NOT : '\n';
	case '$': return '\n';     case '.': return S_NOCARE; case '*': return S_STAR;   case '|': return S_ORSYM;
	case ',': return S_ORPAT;  case ';': return S_ANDPAT; case '-': return in_range ? S_HYPHEN : '-';
	default: return c;
	}
}

static int add_literal(build_t *b, int c, int nocase, char *err, size_t errlen)
{
	pos_t *p = new_pos(b);
	if (!p) FAIL("pattern too long (has > %d chars)", WIDTH);
	if (c == '\n') { p->prot = 1; cls_set(p, '\n'); return 0; }               /* maskgen.c:171-175 */
	if (b->no_error) p->prot = 1;                                              /* maskgen.c:195 */
	if (nocase && is_upper(c)) c += 32;                                        /* maskgen.c:52-59 */
	cls_set(p, c);
	if (nocase && is_alpha(c)) cls_set(p, c - 32);                             /* maskgen.c:259-266 */
	p->lit = c;
	return 0;
}

static int add_wordb(build_t *b, char *err, size_t errlen)
{
	pos_t *p = new_pos(b);
	if (!p) FAIL("pattern too long (has > %d chars)", WIDTH);
	p->prot = 1;                                                               /* maskgen.c:176-187 */
	cls_range(p, 1, 47); cls_range(p, 58, 64); cls_range(p, 91, 96); cls_range(p, 123, 127);
	return 0;
}

static int add_sep(build_t *b, int is_and, int L, char *err, size_t errlen)
{
	pos_t *p;
	if (is_and) {                                                              /* maskgen.c:150-163 */
		if (b->n + 1 > L + 1) b->and_mode = 1;
		if (b->or_seen) FAIL("illegal pattern: cannot handle AND (';') and OR (',') simultaneously");
	} else {                                                                   /* maskgen.c:136-149 */
		if (b->and_mode) FAIL("illegal pattern: cannot handle OR (',') and AND (';') simultaneously");
		b->or_seen = 1;
	}
	p = new_pos(b);
	if (!p) FAIL("pattern too long (has > %d chars)", WIDTH);
	p->is_sep = 1;
	b->nparts++;
	return 0;
}

/* the user's pattern (after the delimiter part and the optional -w/-x opener) */
static int add_pattern(build_t *b, const unsigned char *s, int n, const agb_options *o, int L, char *err, size_t errlen)
{
	int i, esc;
	for (i = 0; i < n; i++) {
		if (s[i] == '\\') {                                                    /* preproce.c:139-142 */
			/* a lone backslash at the very end escapes whatever preprocess() appended behind the pattern (the string
			 * terminator, or the '<' of the -w/-x wrapper, preproce.c:148-175): not restated, refused */
			if (++i >= n) {
				if (o->wordbound || o->wholeline) FAIL("the pattern ends in a lone backslash");
				n--;                                                           /* it escapes the terminator: nothing (the reference's strlen() stops there) */
			}
		}
		else if (s[i] == '|' || s[i] == '*')
			FAIL("regular expressions (re()/re1(), agrep.c:468-1917) are outside the accelerated scan path");
	}
	for (i = 0; i < n; i++) {
		int c = map_sym(s, &i, n, 0, &esc);
		if (!esc && s[i] >= 129 && s[i] <= 145) FAIL("byte %d in the pattern collides with an internal symbol (agrep.h:69-87)", s[i]);
		if (esc) { if (add_literal(b, c, o->nocase, err, errlen)) return AGB_ERR_PATTERN; continue; }
		switch (c) {
		case S_WILD: if (b->n >= 1) b->p[b->n].wild = 1; break;
		case S_LANGLE: b->no_error = 1; b->even++; break;
		case S_RANGLE: b->no_error = 0; if (--b->even < 0) FAIL("unmatched '<', '>' (use \\<, \\> to search for <, >)"); break;
		case S_LPAREN: case S_RPAREN: break;                                   /* maskgen.c:194: no position */
		case S_RRANGE: FAIL("unmatched '[', ']' (use \\[, \\] to search for [, ])");
		case S_ORPAT: if (add_sep(b, 0, L, err, errlen)) return AGB_ERR_PATTERN; break;
		case S_ANDPAT: if (add_sep(b, 1, L, err, errlen)) return AGB_ERR_PATTERN; break;
		case S_NOCARE: {
			pos_t *p = new_pos(b);
			if (!p) FAIL("pattern too long (has > %d chars)", WIDTH);
			if (b->no_error) p->prot = 1;
			cls_range(p, S_NOCARE, S_NOCARE);
			break; }
		case S_LRANGE: {                                                       /* maskgen.c:96-127 */
			pos_t *p = new_pos(b); int compl_ = 0, closed = 0, lo;
			uint64_t keep[4];
			if (!p) FAIL("pattern too long (has > %d chars)", WIDTH);
			if (b->no_error) p->prot = 1;
			i++;
			if (i < n && s[i] == '^') { compl_ = 1; i++; }
			lo = -1;
			{   /* maskgen.c:104-116 keeps the class as (low, high) pairs: a symbol opens the pair (c, c), "-x" replaces the high
			     * end of the last pair -- so a descending range like z-a matches nothing, not even z */
				int plo[2 * WIDTH], phi[2 * WIDTH], np = 0, q;
				for (; i < n; i++) {
					int cc = map_sym(s, &i, n, 1, &esc);
					if (!esc && cc == S_RRANGE) { closed = 1; break; }
					if (!esc && cc == S_HYPHEN) {                              /* class[k-1] = next symbol */
						int hi;
						i++;
						if (i >= n) break;
						hi = map_sym(s, &i, n, 1, &esc);
						if (o->nocase && is_upper(hi)) hi += 32;
						if (np > 0) phi[np - 1] = hi;
						continue;
					}
					if (o->nocase && is_upper(cc)) cc += 32;                   /* Pattern[] is lower-cased as a whole */
					if (np < 2 * WIDTH) { plo[np] = phi[np] = cc; np++; }
				}
				for (q = 0; q < np; q++) {
					if (plo[q] == S_NOCARE) cls_range(p, S_NOCARE, S_NOCARE);   /* maskgen.c:242-246 looks at the low end first: '.' = any */
					else if (plo[q] <= phi[q]) cls_range(p, plo[q], phi[q]);
				}
			}
			(void)lo;
			if (!closed) FAIL("unmatched '[', ']' (use \\[, \\] to search for [, ])");
			if (compl_) { p->cls[0] = ~p->cls[0]; p->cls[1] = ~p->cls[1]; p->cls[2] = ~p->cls[2]; p->cls[3] = ~p->cls[3]; }
			if (o->nocase) {                                                   /* maskgen.c:259-266: Mask[U] = Mask[u] */
				int u;
				memcpy(keep, p->cls, sizeof keep);
				for (u = 'A'; u <= 'Z'; u++) {
					p->cls[u >> 6] &= ~(1ull << (u & 63));
					if (keep[(u + 32) >> 6] >> ((u + 32) & 63) & 1) cls_set(p, u);
				}
			}
			break; }
		default:
			if (add_literal(b, c, o->nocase, err, errlen)) return AGB_ERR_PATTERN;
		}
	}
	if (b->even != 0) FAIL("unmatched '<', '>' (use \\<, \\> to search for <, >)");
	return 0;
}

/* the -i table of bitap.c:171 as the reference leaves it: CP[ISO-8859-1].lower_1 (agrep.c:2769-2792,
 * codepage.c:399-533) with every byte that serves as a metasymbol put back to itself (agrep.c:2835-2848; in this
 * codepage that undoes the lower_1 entries of 0x83, 0x8f and 0x99), stated as identity + ASCII folding + the 29
 * high-half entries that remain */
void agbi_lut_lower1(unsigned char lut[256])
{
	static const unsigned char hi[] = {
		0x80,0x87, 0x8a,0x9a, 0x8c,0x9c, 0x8e,0x9e, 0x90,0x82, 0x92,0x91,
		0xc1,0xe1, 0xc3,0xe3, 0xc4,0xe4, 0xc5,0xe5, 0xc7,0xe7, 0xc8,0xe8, 0xc9,0xe9, 0xca,0xea, 0xcc,0xec,
		0xcd,0xed, 0xce,0xee, 0xcf,0xef, 0xd1,0xf1, 0xd2,0xf2, 0xd3,0xf3, 0xd4,0xf4, 0xd5,0xf5, 0xd6,0xf6,
		0xd8,0xf8, 0xda,0xfa, 0xdc,0xfc, 0xdd,0xfd, 0xde,0xfe };
	size_t i;
	for (i = 0; i < 256; i++) lut[i] = (unsigned char)(is_upper((int)i) ? i + 32 : i);
	for (i = 0; i + 1 < sizeof hi; i += 2) lut[hi[i]] = hi[i + 1];
}

/* one automaton step on all rows: asearch.c:96-115 (unit costs) / asearch1.c:88-97 (costs) / bitap.c:175-176 */
void agbi_step(const agb_desc *d, const uint64_t *B, uint64_t *A, uint64_t cm)
{
	int r, n = d->k;
	A[0] = ((B[0] >> 1) & cm) | (d->init1 & B[0]);
	if (d->engine == AGB_ENGINE_ASEARCH1) {
		int I = d->cost_i, S = d->cost_s, DD = d->cost_d;
		for (r = 1; r <= n; r++) {
			uint64_t bi = (r - I >= 0) ? B[r - I] : 0, ad = (r - DD >= 0) ? A[r - DD] : 0, bs = (r - S >= 0) ? B[r - S] : 0;
			A[r] = ((B[r] >> 1) & cm) | bi | (((ad | bs) >> 1) & d->noerr) | (d->init1 & B[r]);
		}
	} else {
		for (r = 1; r <= n; r++)
			A[r] = ((B[r] >> 1) & cm) | (d->init1 & B[r]) | B[r - 1] | (((A[r - 1] | B[r - 1]) >> 1) & d->noerr);
	}
}

static int has_border(const unsigned char *d, int L)
{
	int b;
	for (b = 1; b < L; b++) if (memcmp(d, d + L - b, (size_t)b) == 0) return 1;
	return 0;
}
/* the delimiter as the device compares it: letters that accept both cases in lower case */
static void folded_delim(const agb_desc *d, unsigned char *out) { int p; for (p = 0; p < d->L; p++) out[p] = (unsigned char)(d->delim[p] | d->delim_fold[p]); }

/* derive the words from the positions (maskgen.c:218-257 with WORD = 64, LSB aligned) and the device-only constants */
static int finish(build_t *b, agb_desc *d, const agb_options *o, const unsigned char *lut, char *err, size_t errlen)
{
	int M = b->n, p, c, L = d->L;
	uint64_t sep = 0, endp;
#define BITP(q) (1ull << (M - (q)))
	d->M = M;
	d->wildmask = 0; d->noerr = ~0ull; d->init0 = ~0ull << M;
	memset(d->mask, 0, sizeof d->mask);
	for (p = 1; p <= M; p++) {
		pos_t *q = &b->p[p];
		if (q->is_sep) sep |= BITP(p);
		if (q->wild) d->wildmask |= BITP(p);
		if (q->prot) d->noerr &= ~BITP(p);
		for (c = 0; c < 256; c++) if (cls_has(q, lut ? lut[c] : c)) d->mask[c] |= BITP(p);
	}
	d->init0 |= sep;
	endp = (sep << 1) + 1;
	d->init1 = d->init0 | d->wildmask | endp;
	d->dendpos = endp & BITP(L);
	d->endpos = endp ^ d->dendpos;
	d->dmask = 0;
	for (p = 1; p <= L; p++) d->dmask |= BITP(p);
	d->dmask = ~d->dmask;
	d->and_mode = b->and_mode;
	if (o->ins_free) d->init1 = ~0ull;                                         /* bitap.c:123, asearch.c:49 */
	return agbi_derive(d, err, errlen);
}

/* post-delimiter rows: asearch.c:175-186 / bitap.c:223-225 / asearch1.c:150-158.  Row 0 is masked with
 * D_Mask BEFORE the upper rows read it, exactly as the reference orders the statements. */
static void reset_rows(const agb_desc *d, uint64_t cm, uint64_t *A)
{
	uint64_t B[2 * AGB_MAXERR + 1]; int r;
	for (r = 0; r <= d->k; r++) B[r] = d->init0;
	A[0] = (((B[0] >> 1) & cm) | (d->init1 & B[0])) & d->dmask;
	if (d->engine == AGB_ENGINE_ASEARCH1) {
		int I = d->cost_i, S = d->cost_s, DD = d->cost_d;
		for (r = 1; r <= d->k; r++) {
			uint64_t bi = (r - I >= 0) ? B[r - I] : 0, ad = (r - DD >= 0) ? A[r - DD] : 0, bs = (r - S >= 0) ? B[r - S] : 0;
			A[r] = ((B[r] >> 1) & cm) | bi | (((ad | bs) >> 1) & d->noerr) | (d->init1 & B[r]);
		}
	} else {
		for (r = 1; r <= d->k; r++)
			A[r] = ((B[r] >> 1) & cm) | (d->init1 & B[r]) | B[r - 1] | (((A[r - 1] | B[r - 1]) >> 1) & d->noerr);
	}
}

/* everything the device path needs beyond the reference's words */
int agbi_derive(agb_desc *d, char *err, size_t errlen)
{
	int L = d->L, p, r;
	uint64_t B[2 * AGB_MAXERR + 1], A[2 * AGB_MAXERR + 1];
	if (L < 1 || L > AGB_MAXDELIM || d->M < L + 1 || d->M > WIDTH - 1) FAIL("bad descriptor (M=%d, L=%d)", d->M, L);
	if (d->k < 0 || d->k > AGB_MAXERR) FAIL("bad descriptor (k=%d)", d->k);
	if (!d->dendpos) FAIL("internal: delimiter end bit missing");
	/* the device also recognises delimiters away from the automaton (record starts, ordinals), by their bytes: position p
	 * of the delimiter must accept delim[p-1] and nothing else (-i with letters in the delimiter makes it accept both cases) */
	/* -p (Init1 all ones, bitap.c:123) makes every position sticky, the delimiter's too: with a delimiter of two or more
	 * bytes "a ... b" then closes a record like "ab" does.  One byte is fine (its only position is D_endpos itself). */
	if (d->init1 == ~0ull && L > 1)
		FAIL("-p with a delimiter of more than one byte is not supported (insertions inside the delimiter would be free too)");
	memset(d->delim_fold, 0, sizeof d->delim_fold);
	for (p = 1; p <= L; p++) {
		const uint64_t bit = 1ull << (d->M - p); int c, cnt = 0, lo = d->delim[p - 1] | 0x20;
		for (c = 0; c < 256; c++) if (d->mask[c] & bit) cnt++;
		if (cnt == 1 && (d->mask[d->delim[p - 1]] & bit)) continue;
		/* -i with a letter in the delimiter: both cases end the record (maskgen.c:52-58, 259-266) */
		if (cnt == 2 && lo >= 'a' && lo <= 'z' && (d->mask[lo] & bit) && (d->mask[lo - 32] & bit)) { d->delim_fold[p - 1] = 0x20; continue; }
		FAIL("the delimiter matches more than its own bytes here: not supported by the device record search");
	}
	/* delimiter recognition away from the automaton (record-start search on the device) */
	unsigned char fd[2 * AGB_MAXDELIM + 2];
	folded_delim(d, fd);
	if (L == 1 || !has_border(fd, L)) d->delim_kind = 0;
	else {
		for (p = 1; p < L; p++) if (fd[p] != fd[0]) break;
		d->delim_kind = p < L ? 2 : 1;          /* a run such as $$, or any other self-overlap ("aba"): automaton.cuh delim_ends_at */
	}
	reset_rows(d, d->mask[d->delim[L - 1]], d->reset);
	/* the virtual '\n' in front of the text (bitap.c:140,148-149) */
	for (r = 0; r <= d->k; r++) B[r] = d->init0;
	agbi_step(d, B, A, d->mask['\n']);
	if (A[0] & d->dendpos) { d->start_closes = 1; memcpy(d->start, d->reset, sizeof(uint64_t) * (size_t)(d->k + 1)); }
	else { d->start_closes = 0; memcpy(d->start, A, sizeof(uint64_t) * (size_t)(d->k + 1)); }
	d->nrows = d->k + 1;
	return 0;
}

/* pigeonhole anchor plan: k errors can damage at most k of k+1 disjoint runs of consecutive literal
 * positions, so a matching record contains one run verbatim (the idea of sgrep.c:1053-1154, made exact). */
/* the bytes a position accepts, when they are at most two (a literal, or a class such as [ea]); 0: not usable in an anchor */
static int pos_values(const pos_t *q, int ascii_only, int literal_only, int *vals)
{
	int c, n = 0;
	if (q->is_sep || (literal_only && q->lit < 0)) return 0;
	if (q->lit >= 0) {
		/* ascii_only (-i): the exact engine folds bytes >= 0x80 through the ISO-8859-1 LUT (bitap.c:171), which the
		 * anchors' plain 0x20 fold cannot express -- such bytes never sit inside an anchor */
		if (q->lit == '\n' || (ascii_only && q->lit >= 0x80)) return 0;
		vals[0] = q->lit; return 1;
	}
	for (c = 0; c < 256; c++) if ((q->cls[c >> 6] >> (c & 63)) & 1) {
		/* (under -i the class holds both cases and the anchors are compared with 0x20 OR-ed in: one value per letter) */
		const int v = (ascii_only && is_alpha(c)) ? (c | 0x20) : c;
		if (c == '\n' || (ascii_only && c >= 0x80)) return 0;
		if (n && (vals[0] == v || (n == 2 && vals[1] == v))) continue;
		if (n == 2) return 0;
		vals[n++] = v;
	}
	return n;
}

#define PIECE_VARIANTS 4
typedef struct { int where, nvar; uint32_t v[PIECE_VARIANTS]; } piece_t;

/* disjoint runs of A consecutive positions that each accept one byte -- or two, as long as a run spells at most
 * PIECE_VARIANTS strings: "b[ea]c" is the two anchors "bec" and "bac" at the same place */
static int collect_runs(const build_t *b, int part, int A, int ascii_only, int literal_only, piece_t *out, int cap)
{
	int p, run = 0, n = 0;
	for (p = 1; p <= b->n; p++) {
		const pos_t *q = &b->p[p];
		int vals[2];
		if (q->part != part || !pos_values(q, ascii_only, literal_only, vals)) { run = 0; continue; }
		if (++run == A) {
			piece_t pc; int t, nv = 1, i;
			pc.where = p - A + 1; pc.v[0] = 0;
			for (t = 0; t < A && nv; t++) {
				int vv[2], m = pos_values(&b->p[p - A + 1 + t], ascii_only, literal_only, vv), old = nv;
				if (nv * m > PIECE_VARIANTS) { nv = 0; break; }
				for (i = 0; i < old; i++) {
					const uint32_t base = pc.v[i];
					pc.v[i] = base | (uint32_t)(vv[0] & 0xFF) << (8 * t);
					if (m == 2) pc.v[nv++] = base | (uint32_t)(vv[1] & 0xFF) << (8 * t);
				}
			}
			pc.nvar = nv;
			if (nv && n < cap) out[n++] = pc;
			run = 0;
		}
		if (q->wild) run = 0;     /* '#' behind q: free insertions there, a verbatim run cannot continue through it */
	}
	return n;
}

static void plan_anchors(const build_t *b, agb_desc *d, const agb_options *o, int fold_all)
{
	int A, p, part, step;
	d->plan = AGB_PLAN_ALL; d->n_anchors = 0; d->n_anchors3 = 0; d->adaptive = 1;
	/* -p makes insertions free: no piece need survive.  -v reports the NON-matching records: the anchors cannot point at
	 * them, so the plan stays AGB_PLAN_ALL -- but the pieces are worked out all the same (see the returns below): a count of
	 * non-matching records is the number of records minus the matching ones (scan.cu, the complement count) */
	if (o->ins_free) return;
	/* per anchor length: literal runs first, then runs that may hold two-valued classes (more anchors for the same pieces) */
	for (step = 0; step < 6; step++) {
		const int lit_only = !(step & 1);
		A = 4 - step / 2;
		uint32_t got[AGB_MAXANCHOR]; int pos[AGB_MAXANCHOR], ngot = 0, ok = 1, classes = 0;
		piece_t pcs[AGB_MAXANCHOR]; int i, v;
#define TAKE_PIECES(arr, cnt) do { \
			for (i = 0; i < (cnt) && ok; i++) for (v = 0; v < (arr)[i].nvar; v++) { \
				if (ngot >= AGB_MAXANCHOR) { ok = 0; break; } \
				got[ngot] = (arr)[i].v[v]; pos[ngot++] = (arr)[i].where; if ((arr)[i].nvar > 1) classes = 1; \
			} } while (0)
		if (b->or_seen) {                    /* a,b : any alternative may match -> k+1 runs from each */
			for (part = 1; part <= b->nparts && ok; part++) {
				int nt = collect_runs(b, part, A, o->nocase || fold_all, lit_only, pcs, AGB_MAXANCHOR);
				if (nt < d->k + 1) ok = 0; else TAKE_PIECES(pcs, d->k + 1);
			}
		} else {                             /* single pattern or a;b (all must match): the part richest in runs */
			int best = -1, bestpart = 0;
			for (part = 1; part <= b->nparts; part++) {
				int nt = collect_runs(b, part, A, o->nocase || fold_all, lit_only, pcs, AGB_MAXANCHOR);
				if (nt > best) { best = nt; bestpart = part; }
			}
			if (best < d->k + 1) ok = 0;
			else { collect_runs(b, bestpart, A, o->nocase || fold_all, lit_only, pcs, AGB_MAXANCHOR); TAKE_PIECES(pcs, d->k + 1); }
		}
#undef TAKE_PIECES
		if (!ok) continue;
		d->plan = AGB_PLAN_ANCHORS; d->anchor_len = A;
		d->anchor_mask = (A == 4) ? 0xFFFFFFFFu : (A == 3 ? 0x00FFFFFFu : 0x0000FFFFu);
		/* case folding: the SAME 0x20 is OR-ed into every byte of the text words and of the anchors (a window
		 * is cut from two words at any byte offset, so the fold must not depend on the byte lane).  Both
		 * sides are folded alike, so this only widens the filter ('@' and '`' fall together, etc.). */
		d->anchor_fold = 0;
		if (o->nocase || fold_all)
			for (p = 0; p < ngot; p++) {
				int t;
				for (t = 0; t < A; t++) if (is_alpha((int)(got[p] >> (8 * t) & 0xFF))) d->anchor_fold = 0x20202020u;
			}
		/* stage 1.5: a hit of anchor i at text offset t can only belong to a match inside
		 * [t - off_i - k, t + pat_len - off_i + k) when the pattern is a single part without '#' */
		d->pat_len = d->M - d->L - 1;
		d->n_anchors = 0;
		for (p = 0; p < ngot; p++) {
			const uint32_t val = (got[p] | d->anchor_fold) & d->anchor_mask; int dup = 0;
			for (i = 0; i < d->n_anchors; i++) if (d->anchor[i] == val && d->anchor_off[i] == pos[p] - (d->L + 2)) dup = 1;   /* [eE] under -i */
			if (dup) continue;
			d->anchor[d->n_anchors] = val; d->anchor_off[d->n_anchors++] = pos[p] - (d->L + 2);
		}
		d->refine = (b->nparts == 1 && !b->and_mode && !b->or_seen && d->wildmask == 0) ? 1 : 0;
		/* the planner in scan.cu re-derives plans from the literal positions alone: not for a plan that leans on classes */
		if (classes) d->adaptive = 0;
		if (d->inverse) d->plan = AGB_PLAN_ALL;       /* (the anchors stay in the descriptor for the complement count) */
		return;
	}
}

/* checksg.c:43-122 */
static int simple_pattern(const unsigned char *s, int m, int k, int *notsgrep)
{
	int i;
	*notsgrep = 0;
	for (i = 0; i < m; i++) {
		if (strchr(";,.*-[]()<>|#{}~", s[i])) return 0;
		if (s[i] == '^' || s[i] == '$') { *notsgrep = 1; return k > 0 ? 0 : 1; }
		if (s[i] == '\\') i++;
	}
	return 1;
}

static int parse_delim(const agb_options *o, build_t *b, agb_desc *d, char *err, size_t errlen)
{
	/* agrep.c:2272-2314 builds "<X>; "; preproce.c:181-210 walks it; bitap.c:92-94 maps ^,$ to '\n' */
	d->L = 0; d->user_delim = 0; d->outtail = 0;
	if (!o->delim) {
		pos_t *p = new_pos(b);
		p->prot = 1; cls_set(p, '\n');
		d->delim[d->L++] = '\n';
	} else {
		const unsigned char *s = (const unsigned char *)o->delim; size_t n = strlen(o->delim), i;
		if (n < 1) FAIL("the -d option must have a delimiter argument");
		if (n > 16) FAIL("delimiter pattern too long (has > %d chars)", 16);
		if (n == 1 && (s[0] == '\n' || s[0] == '$' || s[0] == '^')) d->outtail = 1;
		d->user_delim = 1;
		b->no_error = 1; b->even = 1;                  /* the '<' agrep.c:2287 puts in front */
		for (i = 0; i < n; i++) {
			int c = s[i]; pos_t *p;
			if (c == '\\') { if (++i >= n) break; c = s[i]; }
			else if (c == '<') { b->no_error = 1; b->even++; continue; }
			else if (c == '>') { b->no_error = 0; b->even--; continue; }
			else if (c == '^' || c == '$') c = '\n';
			if (c >= 129 && c <= 145) FAIL("byte %d in the delimiter collides with an internal symbol", c);
			if (d->L >= AGB_MAXDELIM) FAIL("delimiter pattern too long (has > %d chars)", AGB_MAXDELIM);
			p = new_pos(b);
			if (c == '\n' || b->no_error) p->prot = 1;
			cls_set(p, c);
			/* -i lower-cases the WHOLE internal pattern, the delimiter included, and copies every lower-case mask to its
			 * upper-case byte (maskgen.c:52-58, 259-266): 'x' and 'X' both end a record then.  Stated here as it is;
			 * agbi_derive() refuses it, because the device finds delimiters by their bytes. */
			if (o->nocase && is_alpha(c)) { cls_set(p, c | 32); cls_set(p, (c | 32) - 32); }
			d->delim[d->L++] = (unsigned char)c;
		}
		b->no_error = 0; b->even--;                    /* the closing '>' */
		if (b->even != 0) FAIL("unmatched '<', '>' in the delimiter");
		if (d->L < 1) FAIL("empty delimiter");
	}
	return 0;
}

int agbi_build(const char *pattern, const agb_options *o, agb_desc *d, char *err, size_t errlen)
{
	build_t *b; int m, rc, notsgrep = 0, simple, jump, sg; unsigned char lut[256];
	const unsigned char *s = (const unsigned char *)pattern;
	memset(d, 0, sizeof *d);
	if (!pattern || !o) FAIL("null argument");
	m = (int)strlen(pattern);
	if (m < 1) FAIL("pattern length %d too small", m);                         /* agrep.c:3052 */
	if (m >= 256) FAIL("pattern '%s' too long", pattern);                      /* agrep.c:3057 */
	if (o->k < 0 || o->k > AGB_MAXERR) FAIL("the maximum number of errors is %d", AGB_MAXERR);   /* agrep.c:2713 */
	if (m <= o->k) FAIL("size of pattern '%s' must be > #of errors %d", pattern, o->k);          /* checksg.c:34 */
	if (o->wordbound && o->wholeline) FAIL("illegal option combination (-x and -w)");            /* agrep.c:2194 */
	if (o->delim && o->wholeline) FAIL("-d and -x are not compatible");                          /* compat.c */
	jump = (o->cost_i || o->cost_s || o->cost_d);
	if (jump && (o->cost_i < 0 || o->cost_s < 0 || o->cost_d < 0)) FAIL("the error cost cannot be 0");
	d->k = o->k; d->inverse = o->inverse != 0;
	d->cost_i = o->cost_i ? o->cost_i : 1; d->cost_s = o->cost_s ? o->cost_s : 1; d->cost_d = o->cost_d ? o->cost_d : 1;
	if (d->cost_i > d->k) d->cost_i = d->k + 1;                                /* asearch1.c:42-44 */
	if (d->cost_s > d->k) d->cost_s = d->k + 1;
	if (d->cost_d > d->k) d->cost_d = d->k + 1;

	/* engine choice: checksg.c:124-144 then bitap.c:96-121, asearch.c:50-52 */
	simple = simple_pattern(s, m, o->k, &notsgrep);
	sg = simple && !o->bestmatch && !(o->nocase && o->k > 0) && !jump && !o->ins_free && !o->linenum
	     && !(o->wordbound && o->k > 0) && !(o->wholeline && o->k > 0) && !notsgrep;
	if (sg && o->k == 0 && !o->wholeline) d->engine = AGB_ENGINE_SGREP_BM;   /* also under -d: checksg() does not look at the delimiter */
	else if (o->k > 0 && jump) d->engine = AGB_ENGINE_ASEARCH1;
	else if (o->k > 4) d->engine = AGB_ENGINE_ASEARCH0;
	else if (o->k > 0) d->engine = AGB_ENGINE_ASEARCH;     /* also simple k>0 literals: the reference's sgrep filters are lossy */
	else d->engine = AGB_ENGINE_BITAP;

	b = (build_t *)calloc(1, sizeof *b);
	if (!b) FAIL("out of memory");
	rc = parse_delim(o, b, d, err, errlen);
	if (!rc) rc = add_sep(b, 1, d->L, err, errlen);                            /* preproce.c:221: ANDPAT after the delimiter */
	b->and_mode = 0;
	if (!rc && d->engine == AGB_ENGINE_SGREP_BM) {
		/* sgrep.c:289-320 + bm() :741-755: literal compared under TR[] (ASCII case folded, unconditional,
		 * sgrep.c:226-236); -w = neither neighbour isalnum().  Stated as an exact automaton. */
		int i;
		if (o->wordbound) { pos_t *p = new_pos(b); int c; if (p) { p->prot = 1; for (c = 0; c < 256; c++) if (!is_alnum(c)) cls_set(p, c); } else rc = AGB_ERR_PATTERN; }
		for (i = 0; i < m && !rc; i++) {
			int c = s[i]; pos_t *p;
			if (c == '\\') { if (++i >= m) break; c = s[i]; }
			p = new_pos(b);
			if (!p) { rc = AGB_ERR_PATTERN; if (err) snprintf(err, errlen, "pattern too long (has > %d chars)", WIDTH); break; }
			if (is_upper(c)) c += 32;
			cls_set(p, c); if (is_alpha(c)) cls_set(p, c - 32);
			p->lit = c;
			if (c == '\n') p->prot = 1;
		}
		if (!rc && o->wordbound) { pos_t *p = new_pos(b); int c; if (p) { p->prot = 1; for (c = 0; c < 256; c++) if (!is_alnum(c)) cls_set(p, c); } else rc = AGB_ERR_PATTERN; }
	} else if (!rc) {
		if (o->wholeline) {                                                    /* preproce.c:148-159, maskgen.c:188-193 */
			pos_t *p = new_pos(b); if (p) { p->prot = 1; cls_set(p, '\n'); cls_set(p, S_NNLINE); } else rc = AGB_ERR_PATTERN;
		} else if (o->wordbound) rc = add_wordb(b, err, errlen);               /* preproce.c:161-166 */
		if (!rc) rc = add_pattern(b, s, m, o, d->L, err, errlen);
		if (!rc && o->wholeline) { pos_t *p = new_pos(b); if (p) { p->prot = 1; cls_set(p, '\n'); } else rc = AGB_ERR_PATTERN; }
		else if (!rc && o->wordbound) rc = add_wordb(b, err, errlen);          /* preproce.c:169-173 */
	}
	if (rc) { if (err && errlen && !err[0]) snprintf(err, errlen, "pattern too long (has > %d chars)", WIDTH); free(b); return AGB_ERR_PATTERN; }
	if (d->engine == AGB_ENGINE_BITAP && o->nocase) { agbi_lut_lower1(lut); rc = finish(b, d, o, lut, err, errlen); }
	else rc = finish(b, d, o, NULL, err, errlen);
	if (!rc) plan_anchors(b, d, o, d->engine == AGB_ENGINE_SGREP_BM);
	free(b);
	return rc;
}

/* ---- public wrappers ---- */
int agb_compile(const char *pattern, const agb_options *opt, agb_pattern **out, char *err, size_t errlen)
{
	agb_pattern *p; int rc;
	if (err && errlen) err[0] = 0;
	if (!out) return AGB_ERR_ARG;
	p = (agb_pattern *)calloc(1, sizeof *p);
	if (!p) return AGB_ERR_NOMEM;
	rc = agbi_build(pattern, opt, &p->d, err, errlen);
	if (rc) { free(p); *out = NULL; return rc; }
	*out = p;
	return AGB_OK;
}

int agb_pattern_from_desc(const agb_desc *d, agb_pattern **out, char *err, size_t errlen)
{
	agb_pattern *p; int rc;
	if (err && errlen) err[0] = 0;
	if (!d || !out) return AGB_ERR_ARG;
	p = (agb_pattern *)calloc(1, sizeof *p);
	if (!p) return AGB_ERR_NOMEM;
	p->d = *d;
	/* the caller may bring its own anchor plan (the drop-in layer derives one from the reference's internal pattern) */
	if (p->d.plan != AGB_PLAN_ANCHORS || p->d.n_anchors < 1 || p->d.n_anchors > AGB_MAXANCHOR || p->d.anchor_len < 2 || p->d.anchor_len > 4) {
		p->d.plan = AGB_PLAN_ALL; p->d.n_anchors = 0; p->d.refine = 0;
	}
	if (p->d.n_anchors3 < 0 || p->d.n_anchors3 > 2 || p->d.plan != AGB_PLAN_ANCHORS) p->d.n_anchors3 = 0;
	rc = agbi_derive(&p->d, err, errlen);
	if (rc) { free(p); *out = NULL; return rc; }
	*out = p;
	return AGB_OK;
}

void agb_pattern_free(agb_pattern *p) { free(p); }
const agb_desc *agb_pattern_desc(const agb_pattern *p) { return p ? &p->d : NULL; }

/* j of the reference's loops (bitap.c:178, asearch.c:120): incremented at every record close, the virtual '\n'
 * included; pre-decremented when the text starts with the user's delimiter (bitap.c:151-156; asearch0() has no
 * such correction, asearch.c:609-612).  Same greedy delimiter rule as the device (scan.cu delim_ends_at). */
void agb_fill_ordinals(const agb_pattern *p, const void *h_text, uint64_t n, agb_record *records, uint64_t n_records)
{
	const agb_desc *d = &p->d; const unsigned char *t = (const unsigned char *)h_text;
	const int L = d->L; uint64_t i = 0; long long j = 0, run = 0, q, taken = -(1ll << 60); unsigned char fd[2 * AGB_MAXDELIM + 2]; int z, head = 1;
	if (!n_records) return;
	folded_delim(d, fd);
#define DEQ(c, p) ((((c) | d->delim_fold[p]) & 0x1FF) == fd[p])
	(void)z; (void)head;
	/* (byte for byte against the delimiter as typed, also under -i: bitap.c:151-154 compares old_D_pat) */
	if (d->user_delim && d->engine != AGB_ENGINE_ASEARCH0 && n >= (uint64_t)L && memcmp(t, d->delim, (size_t)L) == 0) j = -1;
	/* position q = -1 is the virtual '\n'; positions n .. n+L-1 are the delimiter appended at EOF */
	for (q = -1; q < (long long)n + L && i < n_records; q++) {
		int c = q < 0 ? '\n' : (q < (long long)n ? t[q] : d->delim[q - (long long)n]), e;
		if (L == 1) e = DEQ(c, 0);
		else if (d->delim_kind == 1) { run = DEQ(c, 0) ? run + 1 : 0; e = run > 0 && run % L == 0; }
		else if (d->delim_kind == 2) {
			/* occurrences are taken from the left; one that shares a byte with the one taken before it is dropped */
			int m = 1, u;
			for (u = 0; u < L && m; u++) {
				long long at = q - u; int cc = at < -1 ? 256 : (at < 0 ? '\n' : (at < (long long)n ? t[at] : d->delim[at - (long long)n]));
				if (!DEQ(cc, L - 1 - u)) m = 0;
			}
			e = m && q - L + 1 > taken;
			if (e) taken = q;
		}
		else {
			int m = 1, u;
			for (u = 0; u < L && m; u++) {
				long long at = q - u; int cc = at < -1 ? 256 : (at < 0 ? '\n' : (at < (long long)n ? t[at] : d->delim[at - (long long)n]));
				if (!DEQ(cc, L - 1 - u)) m = 0;
			}
			e = m;
		}
		if (!e) continue;
		j++;
		while (i < n_records && records[i].end + L - 1 == q) records[i++].ordinal = j;
	}
}
