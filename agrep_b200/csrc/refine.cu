/* agrep_b200/csrc/refine.cu -- stage 1.5: local verification of anchor hits */
#include "scan_internal.cuh"

/* ================================================================================================
 * stage 1.5: local verification of anchor hits
 *
 * Stage 1 passes every chunk in which an anchor starts; for a pattern made of common words that is a few
 * percent of all chunks (4.5 % for `because each`), 99 % of which belong to no match.  A match that uses the
 * anchor occurrence at text offset t aligns the pat_len pattern positions to text inside
 * [t - off - k, t + pat_len - off + k): only that window has to be looked at.  Three steps, each a necessary
 * condition for the next, each run by full warps (32 candidates at a time, whatever their density in the text):
 *
 *   detect  which windows of the flagged chunk start an anchor -> the pattern starts p0 = t - off (stage 1's
 *           polynomial again: the bitmap only says "somewhere in these 16 bytes")
 *   count   (T1) with at most k errors at most k pattern positions stay unmatched, and a matched position sits on a
 *           diagonal within +-k of its own: the literal positions whose byte shows up nowhere in its 2k+1 band
 *           are counted with byte-parallel compares on the window held in registers (no table lookups, no
 *           recurrence); more than k of them -> no match can use this anchor hit.  Removes 87 % of the hits of
 *           the benchmark pattern at a quarter of the cost of the recurrence
 *   walk    the SAME recurrence as the record stage over just the window (all rows started at Init[0], whose
 *           separator bit is the always-on start state; no record logic, which can only remove bits)
 *
 * A chunk one of whose windows passes the walk gets its bit in the survivor bitmap (a second bitmap: stage 1's
 * is only read); the record stage then runs the reference's loop on the records those chunks meet -- this
 * stage only ever removes work, it decides nothing.
 *
 * A warp owns a contiguous range of bitmap words.  Flagged chunks go through a ring in shared memory, 32 at a
 * time; the text around a batch (3..8 sixteen-byte groups per lane) is loaded one batch ahead into registers and
 * stored to the lane's strip of shared memory (odd stride: no bank conflicts) while the next batch's loads are
 * in flight.  detect and count run back to back on the first pattern start of every chunk; the few chunks with a
 * second, different start and the starts that passed the count wait in two more rings until 32 of them are
 * there (partial batches only at the very end of the range).  The warp also counts the survivor bits it sets:
 * the per-warp counts are what the compaction into the ordered candidate list scans (k_compact_ranges).
 * ============================================================================================== */

#define REFINE_MAXG  8
/* refine_u32a.cu, refine_u32b.cu, refine_u64.cu, refine_costs.cu: the kernel's instantiations */
int refine_launch_u32a(int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st);
int refine_launch_u32b(int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st);
int refine_launch_u64(int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st);
int refine_launch_costs(bool narrow, int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st);

/* survivor bitmap -> ordered candidate list, by the same warp ranges stage 1.5 used: range_offsets is the
 * exclusive scan of the per-warp survivor counts */
__global__ void __launch_bounds__(REFINE_THREADS)
k_compact_ranges(const uint32_t *bitmap, uint64_t n_words, const uint64_t *range_offsets, uint64_t *list, uint64_t cap)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * REFINE_THREADS + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * REFINE_THREADS) >> 5;
	const uint64_t n_groups = (n_words + 31) / 32, per = (n_groups + nwarps - 1) / nwarps;
	const uint64_t g_begin = warp * per < n_groups ? warp * per : n_groups, g_end = (g_begin + per < n_groups) ? g_begin + per : n_groups;
	uint64_t at = range_offsets[warp];
	for (uint64_t g0 = g_begin; g0 < g_end; g0 += 8) {      /* eight loads in flight: the survivors are few, the loop is all latency */
		uint32_t wd[8];
#pragma unroll
		for (int u = 0; u < 8; u++) { const uint64_t w = (g0 + u) * 32 + lane; wd[u] = (g0 + u < g_end && w < n_words) ? bitmap[w] : 0u; }
#pragma unroll
		for (int u = 0; u < 8; u++) {
			const uint32_t word = wd[u];
			if (!__ballot_sync(0xffffffffu, word != 0)) continue;
			const uint64_t w = (g0 + u) * 32 + lane;
			uint32_t c = __popc(word), pre = c;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= (uint32_t)o) pre += v; }
			const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
			uint64_t mine = at + (pre - c);
			for (uint32_t b = word; b; b &= b - 1) { if (mine < cap) list[mine] = w * 32 + (uint64_t)(__ffs(b) - 1); mine++; }
			at += total;
		}
	}
}

static bool refine_geometry(const agb_desc &d, RefineParams &P)
{
	if (!front_usable(d) || !d.refine) return false;
	int max_off = 0, min_off = 1 << 30;
	for (int i = 0; i < d.n_anchors; i++) { max_off = std::max(max_off, (int)d.anchor_off[i]); min_off = std::min(min_off, (int)d.anchor_off[i]); }
	for (int i = 0; i < d.n_anchors3; i++) { max_off = std::max(max_off, (int)d.anchor3_off[i]); min_off = std::min(min_off, (int)d.anchor3_off[i]); }
	P.lo_off = max_off + d.k;
	P.hi_off = 15 + d.pat_len - min_off + d.k;
	P.gb = (P.lo_off + 15) / 16;
	P.ng = P.gb + (P.hi_off + 15) / 16;
	if (P.ng < P.gb + 2) P.ng = P.gb + 2;              /* the chunk itself and the word that follows it */
	const int pw = (d.pat_len + 3) / 4, ngc = (d.engine != AGB_ENGINE_ASEARCH1 && d.M <= 31 && pw <= 4) ? 4 : REFINE_MAXG;
	if (!(P.ng <= ngc && max_off <= 31 && d.pat_len >= 1 && d.pat_len <= 64)) return false;
	/* which anchor a hit window holds is looked up through a multiplicative hash of its bytes: 32 slots, a multiplier
	 * that keeps the anchors apart.  Equal anchors at different offsets ("abab" in "abababab"), or a three-byte anchor
	 * that is the head of a four-byte one, have no such thing: those patterns go to stage 2 unthinned. */
	const int na = d.n_anchors + d.n_anchors3;
	if (na > 16) return false;
	uint32_t val[16], msk[16]; int off[16];
	for (int i = 0; i < d.n_anchors; i++) { val[i] = d.anchor[i]; msk[i] = d.anchor_mask; off[i] = d.anchor_off[i]; }
	for (int i = 0; i < d.n_anchors3; i++) { val[d.n_anchors + i] = d.anchor3[i]; msk[d.n_anchors + i] = 0x00FFFFFFu; off[d.n_anchors + i] = d.anchor3_off[i]; }
	for (int i = 0; i < na; i++) for (int j = 0; j < i; j++) if ((val[i] & msk[j]) == val[j] || (val[j] & msk[i]) == val[i]) return false;
	uint32_t mul = 0x9E3779B1u;
	for (int attempt = 0; attempt < 4096; attempt++, mul = mul * 0x2C1B3C6Du + 0x297A2D39u) {
		int8_t slot[32]; bool ok = true;
		memset(slot, 0, sizeof slot);
		mul |= 1u;
		for (int i = 0; i < na && ok; i++) {
			const uint32_t h = (val[i] * mul) >> 27;
			if (slot[h]) ok = false; else slot[h] = (int8_t)(i + 1);
		}
		if (!ok) continue;
		P.hmul = mul;
		memcpy(P.hidx64, slot, sizeof slot);
		int8_t offs8[16];
		for (int i = 0; i < 16; i++) { P.hval[i] = i < na ? val[i] : 0; P.hmask[i] = i < na ? msk[i] : 0; offs8[i] = i < na ? (int8_t)off[i] : 0; }
		memcpy(P.hoffs64, offs8, sizeof offs8);
		return true;
	}
	return false;
}

/* the count step's view of the pattern, from the Mask[] words alone (so descriptors that come from the drop-in layer
 * work too): position j of the pattern proper is literal when exactly one byte value below 0x80 reaches it, or two
 * that differ in the 0x20 bit (then every compare is made under that fold, as stage 1 does); everything else --
 * classes, '.', the -w / -x wrappers, bytes >= 0x80 -- counts as matched */
static void refine_t1_setup(const agb_desc &d, RefineParams &P)
{
	uint8_t lit[64]; bool is_lit[64]; bool fold = false; int n_lit = 0;
	for (int j = 0; j < d.pat_len; j++) {
		const uint64_t bit = 1ull << (d.M - (d.L + 2 + j));
		int cnt = 0, c0 = -1, c1 = -1;
		for (int c = 0; c < 256; c++) if (d.mask[c] & bit) { if (cnt == 0) c0 = c; else if (cnt == 1) c1 = c; cnt++; }
		is_lit[j] = false; lit[j] = 0;
		if (cnt == 1 && c0 < 0x80) { is_lit[j] = true; lit[j] = (uint8_t)c0; }
		else if (cnt == 2 && (c0 ^ c1) == 0x20 && c1 < 0x80) { is_lit[j] = true; lit[j] = (uint8_t)(c0 | 0x20); fold = true; }
		if (is_lit[j]) n_lit++;
	}
	P.t1 = 0;
	if (n_lit <= d.k + 1) return;                        /* the count could never fail */
	P.t1 = 1; P.t1_fold = fold ? 0x20202020u : 0u;
	for (int w = 0; w < 16; w++) { P.t1_pat[w] = 0; P.t1_care[w] = 0; }
	for (int j = 0; j < d.pat_len; j++) if (is_lit[j]) {
		/* stored as pattern byte XOR'ed against the (folded) text; the compare adds 0x7f to the difference */
		P.t1_pat[j >> 2] |= (uint32_t)(fold ? (lit[j] | 0x20) : lit[j]) << (8 * (j & 3));
		P.t1_care[j >> 2] |= 0x80u << (8 * (j & 3));
	}
	/* positions that are not literal: their byte of the difference is arbitrary and not looked at (care = 0); make the
	 * pattern byte 0 there so that the 7-bit sums never carry into the next byte: a 7-bit value + 0x7f stays below 0x100 */
}

/* stage 1.5 over the whole bitmap: W.bitmap -> W.bitmap2, per-warp survivor counts in W.range_counts.
 * *ran = false when the pattern has no local window (then stage 2 works from W.bitmap) */
int refine_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, cudaStream_t st, bool *ran)
{
	RefineParams P; memset(&P, 0, sizeof P);
	*ran = false;
	if (n == 0 || !refine_geometry(d, P)) return AGB_OK;
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32;
	P.text = (const uint8_t *)d_text; P.bitmap = W.bitmap; P.out = W.bitmap2; P.warp_counts = W.range_counts;
	P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.fold = d.anchor_fold; P.amask = d.anchor_mask; P.na = d.n_anchors; P.k = d.k; P.pat_len = d.pat_len;
	for (int i = 0; i < d.n_anchors; i++) { P.anchor[i] = d.anchor[i]; P.off[i] = d.anchor_off[i]; }
	{   /* stage 1's polynomial over the anchors, usable when they are pairwise distinct and pass its false-positive guard */
		bool distinct = true;
		for (int i = 0; i < d.n_anchors; i++) for (int j = 0; j < i; j++) if (d.anchor[i] == d.anchor[j]) distinct = false;
		P.one = 1; P.scale = 1;
		for (int i = d.anchor_len; i < 4; i++) P.scale <<= 8;
		P.poly = (distinct && poly_setup(d.anchor, d.n_anchors, 8 * d.anchor_len, P.coef)) ? 1 : 0;
		P.n3 = 0;
		if (d.n_anchors3) {
			uint32_t c3[AGB_MAXANCHOR];
			if (!P.poly || !poly_setup(d.anchor3, d.n_anchors3, 24, c3)) return AGB_OK;      /* (the planner only makes plans that have both) */
			P.n3 = d.n_anchors3;
			for (int i = 0; i < d.n_anchors3; i++) P.coef3[i] = c3[i];
		}
	}
	refine_t1_setup(d, P);
	unsigned grid = refine_grid(W, n);
	P.sm_count = W.sm_count;
	const bool costs = d.engine == AGB_ENGINE_ASEARCH1, narrow = d.M <= 31;
	int rc = costs ? refine_launch_costs(narrow, d.nrows, P, grid, st)
	       : !narrow ? refine_launch_u64(d.nrows, P, grid, st)
	       : d.nrows <= 4 ? refine_launch_u32a(d.nrows, P, grid, st) : refine_launch_u32b(d.nrows, P, grid, st);
	if (rc) return AGB_ERR_ARG;
	g_launches++;
	CUDA_TRY(cudaGetLastError());
	W.refine_ctas = grid;
	*ran = true;
	return AGB_OK;
}

/* upper bound of the grid of stage 1.5 (the launch trims it to one wave and leaves the result in W.refine_ctas: the
 * warp ranges of the survivor counts and of k_compact_ranges) */
unsigned refine_grid(const Workspace &W, uint64_t n)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, groups = (n_words + 31) / 32;
	unsigned grid = (unsigned)std::min<uint64_t>((groups + 3) / 4, (uint64_t)W.sm_count * 16);
	return grid ? grid : 1;
}

int compact_ranges_launch(Workspace &W, uint64_t n, cudaStream_t st)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32;
	const unsigned grid = W.refine_ctas;
	k_compact_ranges<<<grid, REFINE_THREADS, 0, st>>>(W.bitmap2, n_words, W.range_offsets, W.cand, W.cand_cap); g_launches++;
	CUDA_TRY(cudaGetLastError());
	return AGB_OK;
}
