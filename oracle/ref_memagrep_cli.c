/* oracle/ref_memagrep_cli.c -- TEST INFRASTRUCTURE.
 * 30-line driver around the UNMODIFIED reference's in-memory entry point
 * memagrep() (reference agrep.c:3282).  Links against the reference objects
 * built by oracle/Makefile; used to pin oracle/agrep_oracle.c and to generate
 * tests/golden/.  Contract of memagrep (agrep.c:3275-3280): the buffer starts
 * with '\n' and has slack after its end.
 *
 * usage: memagrep_cli [-dump] FILE agrep-args...   (pattern is among the args)
 * prints: "ret=<N>" on stderr-free stdout tail after the matched records;
 * with -dump also prints the automaton words maskgen left in the globals.   */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
extern int memagrep();
extern unsigned Mask[], Init1, NO_ERR_MASK, Init[], endposition, D_endpos, wildmask;
extern int M, D, AND, SGREP, D_length, INVERSE, DELIMITER, I, S, DD, JUMP, NOUPPER, WORDBOUND, REGEX;
extern char Pattern[], D_pattern[], old_D_pat[];
int main(int argc, char **argv)
{
	int dump = 0, a = 1, i, n, ret;
	FILE *f; char *buf; char *av[64]; int ac = 0;
	if (argc > 1 && !strcmp(argv[1], "-dump")) { dump = 1; a = 2; }
	if (argc - a < 2) { fprintf(stderr, "usage: memagrep_cli [-dump] FILE args...\n"); return 2; }
	f = fopen(argv[a], "rb"); if (!f) { perror(argv[a]); return 2; }
	fseek(f, 0, SEEK_END); n = (int)ftell(f); fseek(f, 0, SEEK_SET);
	buf = calloc(1, (size_t)n + 1 + 8192);
	buf[0] = '\n';
	if (n && fread(buf + 1, 1, n, f) != (size_t)n) { perror("read"); return 2; }
	fclose(f);
	av[ac++] = "agrep";
	for (i = a + 1; i < argc && ac < 62; i++) av[ac++] = argv[i];
	av[ac++] = argv[a];           /* this fork insists on one existing file name (agrep.c:2922-2935) */
	ret = memagrep(ac, av, n + 1, buf, 0, stdout);
	fflush(stdout);
	printf("ret=%d\n", ret);
	if (dump) {
		printf("M=%d D=%d AND=%d SGREP=%d D_length=%d INVERSE=%d DELIMITER=%d I=%d S=%d DD=%d JUMP=%d NOUPPER=%d WORDBOUND=%d REGEX=%d\n",
		       M, D, AND, SGREP, D_length, INVERSE, DELIMITER, I, S, DD, JUMP, NOUPPER, WORDBOUND, REGEX);
		printf("Init0=%08x Init1=%08x NO_ERR_MASK=%08x endposition=%08x D_endpos=%08x wildmask=%08x\n",
		       Init[0], Init1, NO_ERR_MASK, endposition, D_endpos, wildmask);
		for (i = 0; i < 256; i++) if (Mask[i]) printf("Mask[%d]=%08x\n", i, Mask[i]);
	}
	return 0;
}
