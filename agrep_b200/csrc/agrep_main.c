/* agrep_b200/csrc/agrep_main.c -- `agrep-b200`: a small stand-alone command line over libagrepb200.
 *
 * It covers the switches of agrep that reach the scan path (reference agrep.c:2121-2739) and prints what the
 * reference's exec()/output() print for them (agrep.c:3332-3752, 3805-3956): -# -c -i -w -x -v -n -p -I# -S# -D#
 * -d delim -B -y -l -h -s -b -t -V# -e pat, one or more files.  It is NOT the drop-in (that is the reference's
 * own main() linked against libagrepb200_dropin.so); it exists so the engine can be used
 * where the reference's sources are not around.  No regex, no -f/-m multi-pattern, no -r (out of scope).
 */
#include "agrep_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>

static const char *prog = "agrep-b200";
static int COUNT, SILENT, FILENAMEONLY, NOFILENAME, LINENUM, BYTECOUNT, BESTMATCH, NOPROMPT, VERBOSE = 1, OUTTAIL;
static int FNAME, num_of_matched, FIRSTOUTPUT = 1, EATFIRST;

static void usage(void)
{
	fprintf(stderr, "usage: %s [-#cdehilnpstvwxyBDIS] [-d delim] [-e] pattern [files]\n", prog);
	exit(2);
}

static unsigned char *slurp(const char *path, size_t *n, int L, const unsigned char *dpat)
{
	int fd = path ? open(path, O_RDONLY) : 0; struct stat sb; size_t cap, len = 0; unsigned char *b;
	if (fd < 0) { fprintf(stderr, "%s: can't open file for reading: %s\n", prog, path); return NULL; }
	cap = (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode)) ? (size_t)sb.st_size + 1 : (1u << 20);
	b = (unsigned char *)malloc(cap + 64);
	if (!b) { if (path) close(fd); return NULL; }
	b[0] = '\n';                                                  /* the virtual newline (bitap.c:140) */
	for (;;) {
		ssize_t r;
		if (len + 1 >= cap) { cap *= 2; b = (unsigned char *)realloc(b, cap + 64); if (!b) return NULL; }
		r = read(fd, b + 1 + len, cap - len - 1);
		if (r <= 0) break;
		len += (size_t)r;
	}
	if (path) close(fd);
	memcpy(b + 1 + len, dpat, (size_t)L);                         /* bitap.c:161-165 */
	*n = len;
	return b;
}

/* output() of the reference, agrep.c:3805-3956, for the switches we carry */
static void print_record(const unsigned char *hb, const agb_desc *d, const agb_record *rec, const char *fname)
{
	long long i1 = rec->begin + 1, i2 = rec->end, j = rec->ordinal; int L = d->L;   /* buffer indexes (lasti, print_end) */
	if (i1 > i2) return;                                                             /* agrep.c:3811 */
	num_of_matched++;
	if (COUNT || SILENT) return;
	if (OUTTAIL || (!d->user_delim && L == 1 && d->delim[0] == '\n')) { if (j > 1) i1 += L; i2 += L; }   /* agrep.c:3815-3818 */
	if (d->user_delim) j++;                                                          /* agrep.c:3819 */
	if (FIRSTOUTPUT) { if (hb[i1] == '\n') { i1++; EATFIRST = 1; } FIRSTOUTPUT = 0; }  /* agrep.c:3820-3826 */
	while (hb[i1] == '\n' && i1 <= i2) { fputc('\n', stdout); i1++; }                /* agrep.c:3832-3843 */
	if (FNAME) printf("%s: ", fname);
	if (LINENUM) printf("%lld: ", j - 1);
	if (BYTECOUNT) printf("%lld= ", (long long)rec->end);
	if (i1 <= i2) fwrite(hb + i1, 1, (size_t)(i2 - i1 + 1), stdout);
}

/* one pass of exec() over the files (agrep.c:3411-3576); counting = the COUNT=ON passes of the -B sweep */
static int scan_files(const agb_pattern *p, char **files, int nfiles, int counting)
{
	int fi;
	for (fi = 0; fi < (nfiles ? nfiles : 1); fi++) {
		const char *fname = nfiles ? files[fi] : NULL;
		const agb_desc *d = agb_pattern_desc(p);
		size_t n = 0, cap, i; unsigned char *hb = slurp(fname, &n, d->L, d->delim);
		agb_result res; agb_record *recs = NULL; int before = num_of_matched, rc;
		const int count_only = counting || COUNT || SILENT || FILENAMEONLY;
		if (!hb) continue;
		/* the list is sized from a guess; a scan that reports more matching records than fit (an empty record is a record:
		 * -v on blank lines) is run again with exactly n_matched entries */
		cap = count_only ? 0 : n / 64 + 65536;
		for (;;) {
			if (cap) { recs = (agb_record *)realloc(recs, cap * sizeof *recs); if (!recs) { fprintf(stderr, "%s: out of memory\n", prog); exit(255); } }
			rc = agb_scan_host(p, hb + 1, n, count_only ? AGB_WANT_COUNT : (AGB_WANT_RECORDS | AGB_WANT_ORDINALS)   /* output() needs j even without -n (agrep.c:3815) */, recs, cap, &res);
			if (rc) { fprintf(stderr, "%s: scan failed: %s\n", prog, agb_last_error()); exit(255); }   /* no CPU fallback */
			if (!res.truncated) break;
			cap = (size_t)res.n_matched;
		}
		if (FILENAMEONLY && !counting) num_of_matched += res.n_matched ? 1 : 0;   /* the scan stops at the first hit (bitap.c:184-210, sgrep.c:813-814) */
		else if (count_only) num_of_matched += (int)res.n_matched;
		else {
			for (i = 0; i < res.n_records; i++) print_record(hb, d, &recs[i], fname ? fname : "");
		}
		if (!counting) {
			if (COUNT && !FILENAMEONLY) { if (FNAME) printf("%s: %d\n", fname, num_of_matched - before); else printf("%d\n", num_of_matched - before); }   /* agrep.c:3501-3557 */
			if (FILENAMEONLY && num_of_matched > before) printf("%s\n", fname ? fname : "(standard input)");
		}
		free(recs); free(hb);
	}
	return num_of_matched;
}

int main(int argc, char **argv)
{
	agb_options o; agb_pattern *p = NULL; char err[256]; const char *pattern = NULL; int ai, nfiles, rc;
	memset(&o, 0, sizeof o);
	if (argc > 0 && argv[0]) { const char *s = strrchr(argv[0], '/'); prog = s ? s + 1 : argv[0]; }
	for (ai = 1; ai < argc && argv[ai][0] == '-' && argv[ai][1]; ai++) {
		const char *q = argv[ai] + 1; int stop = 0;
		for (; *q && !stop; q++) {
			switch (*q) {
			case 'c': COUNT = 1; break;            case 's': SILENT = 1; break;
			case 'l': FILENAMEONLY = 1; break;     case 'h': NOFILENAME = 1; break;
			case 'n': LINENUM = 1; o.linenum = 1; break;
			case 'b': BYTECOUNT = 1; break;        case 'i': o.nocase = 1; if (q[1] == '0') { o.nocase = 0; q++; } break;
			case 'w': o.wordbound = 1; break;      case 'x': o.wholeline = 1; break;
			case 'v': o.inverse = 1; break;        case 'p': o.ins_free = 1; break;
			case 'B': BESTMATCH = 1; o.bestmatch = 1; break;
			case 'y': NOPROMPT = 1; break;         case 't': OUTTAIL = 1; break;
			case 'I': o.cost_i = atoi(q + 1); stop = 1; break;
			case 'S': o.cost_s = atoi(q + 1); stop = 1; break;
			case 'D': o.cost_d = atoi(q + 1); stop = 1; break;
			case 'V': VERBOSE = isdigit((unsigned char)q[1]) ? atoi(q + 1) : 1; stop = 1; break;
			case 'd': if (q[1]) o.delim = q + 1; else if (ai + 1 < argc) o.delim = argv[++ai]; else usage(); stop = 1; break;
			case 'e': if (ai + 1 < argc) pattern = argv[++ai]; else usage(); stop = 1; break;
			default:
				if (isdigit((unsigned char)*q)) { o.k = atoi(q); if (o.k > AGB_MAXERR) { fprintf(stderr, "%s: the maximum number of errors is %d\n", prog, AGB_MAXERR); return 2; } stop = 1; }
				else { fprintf(stderr, "%s: illegal option  -%c\n", prog, *q); usage(); }
			}
		}
	}
	if (!pattern) { if (ai >= argc) usage(); pattern = argv[ai++]; }
	nfiles = argc - ai;
	if (BESTMATCH && (COUNT || FILENAMEONLY || o.k)) { BESTMATCH = 0; o.bestmatch = 0; fprintf(stderr, "%s: -B option ignored when -c, -l, -f, or -# is on\n", prog); }   /* compat.c:26-29 */
	if (COUNT && LINENUM) { LINENUM = 0; fprintf(stderr, "%s: -n option ignored with -c\n", prog); }   /* compat.c:30-33 (the engine choice stays) */
	FNAME = nfiles > 1 && !NOFILENAME;
	if (o.delim && strlen(o.delim) == 1 && (o.delim[0] == '\n' || o.delim[0] == '$' || o.delim[0] == '^')) OUTTAIL = 1;   /* agrep.c:2290 */
	rc = agb_compile(pattern, &o, &p, err, sizeof err);
	if (rc) { fprintf(stderr, "%s: %s\n", prog, err); return 255; }

	scan_files(p, argv + ai, nfiles, 0);
	if (BESTMATCH && num_of_matched == 0 && nfiles > 0) {
		/* agrep.c:3582-3728: nothing matched -> counting passes at D = 1, 2, ... < M, <= 8 until something matches,
		 * report, ask (unless -y), then one printing pass at that D */
		const int M = agb_pattern_desc(p)->M; int k, best = -1;
		for (k = 1; k < M && k <= AGB_MAXERR && best < 0; k++) {
			agb_pattern *pk; o.k = k;
			if (agb_compile(pattern, &o, &pk, err, sizeof err)) break;
			num_of_matched = 0;
			if (scan_files(pk, argv + ai, nfiles, 1) > 0) best = k;
			agb_pattern_free(pk);
		}
		if (best > 0) {
			int go = 1;
			if (num_of_matched == 1) fprintf(stderr, "%s: 1 word matches within ", prog); else fprintf(stderr, "%s: %d words match within ", prog, num_of_matched);
			if (best == 1) fprintf(stderr, "1 error"); else fprintf(stderr, "%d errors", best);
			if (NOPROMPT) fprintf(stderr, "\n");
			else {
				char c[8] = "y";
				fprintf(stderr, num_of_matched == 1 ? "; search for it? (y/n)" : "; search for them? (y/n)");
				if (!fgets(c, 4, stdin) || c[0] != 'y') go = 0;
			}
			if (go) {
				agb_pattern_free(p); p = NULL; o.k = best;
				if (agb_compile(pattern, &o, &p, err, sizeof err)) return 255;
				num_of_matched = 0;
				scan_files(p, argv + ai, nfiles, 0);
			}
		} else num_of_matched = 0;
	}
	if (EATFIRST) printf("\n");                                         /* agrep.c:3731-3741 */
	if (VERBOSE > 0) printf("Grand Total: %d match(es) found.\n", num_of_matched);   /* agrep.c:3229-3231 */
	if (p) agb_pattern_free(p);
	return num_of_matched;                                              /* main.c:96: exit status = number of matches */
}
