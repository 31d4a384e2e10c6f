/* agrep_b200/csrc/corpus.h -- deterministic synthetic text, identical on host (gcc) and device (nvcc).
 *
 * newline-delimited lower-case ASCII words from a fixed 256-word vocabulary, 6..14 words
 * per line, 10 % of the lines capitalised, no NUL, no byte >= 128, every page ends with '\n'.  The corpus
 * is a sequence of independent 4096-byte pages keyed by (seed, page index) with a counter-based
 * generator (SplitMix64), so any shard can be produced anywhere without generating what precedes it,
 * and a record never crosses a page boundary.  Optional paragraph mode (blank line every 3..8 lines) and
 * planted needle lines carrying the pattern with 0..maxedits substitutions.
 */
#ifndef AGB_CORPUS_H
#define AGB_CORPUS_H
#include <stdint.h>

#ifdef __CUDACC__
#define AGB_HD __host__ __device__ __forceinline__
#else
#define AGB_HD static inline
#endif

#define AGB_PAGE 4096

/* 256 words, packed: offsets into one string.  Lengths 1..12. */
#define AGB_VOCAB_STR \
 "the of and to in that is was he for it with as his on be at by had not are but from or have an they " \
 "which one you were her all she there would their we him been has when who will more no if out so said " \
 "what up its about into than them can only other new some could time these two may then do first any my " \
 "now such like our over man me even most made after also did many before must through back years where " \
 "much your way well down should because each just those people how too little state good very make world " \
 "still own see men work long get here between both life being under never day same another know while " \
 "last might us great old year off come since against go came right used take three states himself few " \
 "house use during without again place american around however home small found mrs thought went say part " \
 "once general high upon school every does got united left number course war until always away something " \
 "fact though water less public put think almost hand enough far took head yet government system better " \
 "set told nothing night end why called eyes find going look asked later knew point next program city " \
 "business give group toward young days let room president side social given present several order national " \
 "possible rather second face per among form important often things " \
 "looked early white case john become large big need four within felt along children saw best church ever least " \
 "power development light thing seemed family interest want members mind country area others done turned "

typedef struct { uint64_t s; } agb_rng;
AGB_HD uint64_t agb_rng_next(agb_rng *r)
{
	uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

/* vocab: pointer to AGB_VOCAB_STR bytes; woff[257]: word start offsets (word i = [woff[i], woff[i+1]-1)) */
AGB_HD void agb_corpus_page(uint8_t *out, uint64_t seed, uint64_t page, const char *vocab, const uint16_t *woff,
                            int paragraphs, int needle_every, const char *needle, int needle_len, int needle_maxedits)
{
	agb_rng rng; int pos = 0, lines_to_gap, first = 1;
	rng.s = seed * 0xD6E8FEB86659FD93ull + page * 0xA0761D6478BD642Full + 0x1234567;
	lines_to_gap = 3 + (int)(agb_rng_next(&rng) % 6);
	for (;;) {
		uint64_t r = agb_rng_next(&rng);
		int nwords = 6 + (int)(r % 9), cap = ((r >> 8) % 10) == 0, w, start = pos;
		int plant = first && needle_every > 0 && needle_len > 0 && (page % (uint64_t)needle_every) == 0;
		int budget = nwords * 13 + (plant ? needle_len + 1 : 0) + 2;
		first = 0;
		if (pos + budget > AGB_PAGE - 1) break;
		for (w = 0; w < nwords; w++) {
			uint64_t q = agb_rng_next(&rng);
			int wi = (int)(q & 255), a = woff[wi], b = woff[wi + 1] - 1, t;
			if (w) out[pos++] = ' ';
			if (plant && w == nwords / 2) {
				/* the pattern with e = (page/needle_every) mod (maxedits+1) substitutions at seeded places */
				int e = (int)((page / (uint64_t)needle_every) % (uint64_t)(needle_maxedits + 1)), base = pos;
				for (t = 0; t < needle_len; t++) out[pos++] = (uint8_t)needle[t];
				for (t = 0; t < e; t++) {
					uint64_t z = agb_rng_next(&rng);
					int at = (int)(z % (uint64_t)needle_len);
					out[base + at] = (uint8_t)('a' + (int)((z >> 16) % 26));
				}
				out[pos++] = ' ';
			}
			for (t = a; t < b; t++) out[pos++] = (uint8_t)vocab[t];
		}
		if (cap && out[start] >= 'a' && out[start] <= 'z') out[start] = (uint8_t)(out[start] - 32);
		out[pos++] = '\n';
		if (paragraphs && --lines_to_gap == 0) {
			out[pos++] = '\n';
			lines_to_gap = 3 + (int)(agb_rng_next(&rng) % 6);
		}
	}
	/* filler line up to the page end: "zq zq zq ...\n" keeps pages independent and NUL-free */
	while (pos < AGB_PAGE - 1) { out[pos] = (uint8_t)((pos % 3) == 2 ? ' ' : ((pos % 3) ? 'q' : 'z')); pos++; }
	out[AGB_PAGE - 1] = '\n';
}

/* builds woff[257] from the vocabulary string (host side; the device gets the table by copy) */
static inline int agb_vocab_offsets(const char *vocab, uint16_t *woff)
{
	int n = 0, i = 0;
	while (vocab[i] && n < 256) {
		woff[n++] = (uint16_t)i;
		while (vocab[i] && vocab[i] != ' ') i++;
		if (vocab[i] == ' ') i++;
	}
	woff[n] = (uint16_t)i;       /* one past the trailing space of the last word */
	return n;
}
#endif
