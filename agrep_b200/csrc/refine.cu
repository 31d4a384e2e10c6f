/* agrep_b200/csrc/refine.cu -- stage 1.5: local verification of anchor hits */
#include "automaton.cuh"

/* ================================================================================================
 * stage 1.5: local verification of anchor hits
 *
 * Stage 1 passes every chunk in which an anchor starts; for a pattern made of common words that is a few
 * percent of all chunks, almost none of which belong to a match.  A match that uses the anchor occurrence at
 * text offset t aligns the pat_len pattern positions to text inside [t - off - k, t + pat_len - off + k), so
 * running the SAME recurrence over just that window (all rows started at Init[0], whose separator bit is the
 * always-on start state; no record logic, which can only remove bits) decides whether the hit can matter.
 * Chunks none of whose hits survive lose their bitmap bit.  A warp first compacts the flagged chunks of its
 * 32 bitmap words into a queue, so all lanes verify; each lane stages the few 16-byte groups around its chunk
 * in shared memory and the lanes walk their windows in lockstep (same length for everyone).
 * ============================================================================================== */

/* the recurrence over one window: rows started at Init[0]; the end bits of the last row are sticky (Init1 holds
 * them, maskgen.c:232), so looking at it after the walk is enough.  Called by all lanes together. */
template <typename T, int NR, bool COSTS>
__device__ __forceinline__ bool window_passes(const uint8_t *bytes, const bool run, const int wlen, const T init0,
                                              const T *mask, const DevConsts<T> &C)
{
	T S[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) S[r] = init0;
	for (int q = 0; q < wlen; q++) {
		const int c = run ? bytes[q] : 0;
		rows_step<T, NR, COSTS>(S, mask[c], C);
	}
	return run && (S[NR - 1] & C.endpos) != 0;
}

/* which of the 16 windows of a chunk start an anchor: bit (32 + s - off_a) per hit, i.e. the distinct WINDOW STARTS
 * of the pattern around this chunk (two anchors of one occurrence, "beca" and "use " inside "because ", coincide).
 * POLY: stage 1's polynomial finds the (rare) windows worth comparing with IMADs on the otherwise idle FMA pipe. */
template <int NA, bool POLY>
__device__ __forceinline__ uint64_t window_starts(const uint32_t *cw, const RefineParams &P)
{
	const uint32_t x0 = cw[0] | P.fold, x1 = cw[1] | P.fold, x2 = cw[2] | P.fold, x3 = cw[3] | P.fold, x4 = cw[4] | P.fold;
	uint32_t wv[16];
	wv[0] = x0; wv[1] = __funnelshift_r(x0, x1, 8); wv[2] = __funnelshift_r(x0, x1, 16); wv[3] = __funnelshift_r(x0, x1, 24);
	wv[4] = x1; wv[5] = __funnelshift_r(x1, x2, 8); wv[6] = __funnelshift_r(x1, x2, 16); wv[7] = __funnelshift_r(x1, x2, 24);
	wv[8] = x2; wv[9] = __funnelshift_r(x2, x3, 8); wv[10] = __funnelshift_r(x2, x3, 16); wv[11] = __funnelshift_r(x2, x3, 24);
	wv[12] = x3; wv[13] = __funnelshift_r(x3, x4, 8); wv[14] = __funnelshift_r(x3, x4, 16); wv[15] = __funnelshift_r(x3, x4, 24);
	uint64_t starts = 0;
	if (POLY) {
		uint32_t zm = 0;
#pragma unroll
		for (int s16 = 0; s16 < 16; s16++) {
			uint32_t r = wv[s16] * P.one + P.coef[NA - 1];
#pragma unroll
			for (int i = NA - 2; i >= 0; i--) r = r * wv[s16] + P.coef[i];
			if (r * P.scale == 0) zm |= 1u << s16;
		}
		for (; zm; zm &= zm - 1) {                             /* usually one bit */
			const int s16 = __ffs(zm) - 1;
			const uint32_t lo = cw[s16 >> 2] | P.fold, hi = cw[(s16 >> 2) + 1] | P.fold;     /* rebuilt: wv[] stays in registers */
			const uint32_t wsel = __funnelshift_r(lo, hi, (s16 & 3) * 8) & P.amask;
#pragma unroll
			for (int a = 0; a < NA; a++) if (wsel == P.anchor[a]) starts |= 1ull << (32 + s16 - P.off[a]);
		}
	} else {
#pragma unroll
		for (int a = 0; a < NA; a++) {
			const uint32_t A = P.anchor[a];
			uint32_t m = 0;
#pragma unroll
			for (int s16 = 0; s16 < 16; s16++) if ((wv[s16] & P.amask) == A) m |= 1u << s16;
			starts |= (uint64_t)m << (32 - P.off[a]);              /* off <= 31: refine_geometry() */
		}
	}
	return starts;
}

template <bool POLY>
__device__ __forceinline__ uint64_t window_starts_na(const uint32_t *cw, const RefineParams &P)
{
	switch (P.na) {
	case 1: return window_starts<1, POLY>(cw, P);  case 2: return window_starts<2, POLY>(cw, P);
	case 3: return window_starts<3, POLY>(cw, P);  case 4: return window_starts<4, POLY>(cw, P);
	case 5: return window_starts<5, POLY>(cw, P);  case 6: return window_starts<6, POLY>(cw, P);
	case 7: return window_starts<7, POLY>(cw, P);  case 8: return window_starts<8, POLY>(cw, P);
	default: return window_starts<9, POLY>(cw, P);
	}
}

/* Streaming form: every warp owns a contiguous range of bitmap words.  It appends the flagged chunks of 32 words at
 * a time to a ring and, whenever 32 are waiting, judges them together: stage the 16-byte groups around the chunk
 * in shared memory, find the window starts, walk the first window; a chunk whose first window fails loses its bit
 * at once (atomicAnd on the bitmap), its other windows (3 % of the chunks have any) go to a second ring and are
 * judged 32 at a time later -- a pass sets the bit again (atomicOr; same warp, program order).  Rings are only
 * flushed partially at the very end of the warp's range, so the lanes stay full. */
#define REFINE_RING  1088         /* >= 31 left over + 1024 new per refill */

/* take up to 32 chunks off the ring and start loading the text around them (NGC x 16 bytes per lane, in registers) */
template <int NGC>
__device__ __forceinline__ void refine_pop(const RefineParams &P, const uint32_t *ring, uint32_t &head, uint32_t &count, uint32_t lane,
                                           uint64_t chunk0, uint64_t &chunk, bool &keep, uint4 (&nx)[NGC])
{
	const uint32_t m = count < 32 ? count : 32;
	const bool active = lane < m;
	chunk = chunk0 + (active ? ring[head + lane] : 0u);
	head += m; count -= m;
	const int64_t base = (int64_t)chunk * 16;
	/* windows that touch the virtual '\n', the appended delimiter or the end of the buffer are not judged here */
	keep = !active || (base - P.lo_off < 0 || (uint64_t)(base + P.hi_off + 16) > P.n || chunk + 2 >= P.n_chunks);
	if (!keep) {
		const uint4 *src = reinterpret_cast<const uint4 *>(P.text) + ((int64_t)chunk - P.gb);
#pragma unroll
		for (int gi = 0; gi < NGC; gi++) if (gi < P.ng) nx[gi] = __ldg(src + gi);
	}
}

/* append the flagged chunks of the next 32 bitmap words (one per lane, loaded one group ahead) to the ring */
__device__ __forceinline__ void refine_refill(const RefineParams &P, uint32_t *ring, uint32_t &head, uint32_t &count, uint32_t lane,
                                              uint64_t &g, uint64_t g_begin, uint64_t g_end, uint32_t &next_word)
{
	const uint32_t word = next_word;
	if (g + 1 < g_end) { const uint64_t w = (g + 1) * 32 + lane; next_word = (w < P.n_words) ? P.bitmap[w] : 0u; }
	uint32_t c = __popc(word), pre = c;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= (uint32_t)o) pre += v; }
	const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
	pre -= c;
	const uint32_t rel0 = (uint32_t)((g - g_begin) * 1024) + lane * 32;
	if (head) {   /* the (< 32) entries left over move to the front: the ring is used linearly */
		const uint32_t v = lane < count ? ring[head + lane] : 0u;
		__syncwarp();
		if (lane < count) ring[lane] = v;
		head = 0;
	}
	for (uint32_t b = word; b; b &= b - 1) { ring[count + pre] = rel0 + (uint32_t)(__ffs(b) - 1); pre++; }
	count += total;
	g++;
	__syncwarp();
}

template <typename T, int NR, bool COSTS, int NGC>
__global__ void __launch_bounds__(REFINE_THREADS)
k_refine(const RefineParams P)
{
	extern __shared__ __align__(16) uint32_t s_stage[];     /* REFINE_THREADS x (ng*4 + 1) words */
	__shared__ RecShared<T, NR> SH;
	__shared__ uint32_t s_ring[REFINE_THREADS / 32][REFINE_RING];
	__shared__ unsigned long long s_defer[REFINE_THREADS / 32][REFINE_DEFER];
	DevConsts<T> C;
	shared_init<T, NR>(SH, C, P.desc, REFINE_THREADS);
	const T init0 = mirror<T>((T)P.desc->init0);
	const int pat_len = P.desc->pat_len, k = C.k, wlen = pat_len + 2 * k;
	const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lt_mask = (1u << lane) - 1u;
	const int stride_w = P.ng * 4 + 1;                      /* odd number of words: lanes hit different banks */
	uint32_t *my_stage = s_stage + threadIdx.x * stride_w;
	const uint8_t *my_bytes = reinterpret_cast<const uint8_t *>(my_stage);
	const int ws0 = P.gb * 16 - 32 - k;                     /* window offset in the staged bytes = ws0 + start bit */
	uint32_t *ring = s_ring[wib];
	unsigned long long *defer = s_defer[wib];
#define REFINE_WINDOW(ptr, run) window_passes<T, NR, COSTS>((ptr), (run), wlen, init0, SH.mask, C)

	/* this warp's groups of 32 bitmap words: [g_begin, g_end) */
	const uint64_t warp = ((uint64_t)blockIdx.x * REFINE_THREADS + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * REFINE_THREADS) >> 5;
	const uint64_t n_groups = (P.n_words + 31) / 32, per = (n_groups + nwarps - 1) / nwarps;
	const uint64_t g_begin = warp * per, g_end = (g_begin + per < n_groups) ? g_begin + per : n_groups;
	if (g_begin >= g_end) return;
	const uint64_t chunk0 = g_begin * 1024;                 /* ring entries are chunk numbers relative to this */

	uint32_t head = 0, count = 0, ndefer = 0;               /* warp-uniform */
	uint64_t g = g_begin;
	uint32_t next_word = (g * 32 + lane < P.n_words) ? P.bitmap[g * 32 + lane] : 0u;     /* one group ahead */
	bool have = false;                                      /* a batch is popped and its text on the way in nx[] */
	uint64_t chunk = 0; bool keep = true;
	uint4 nx[NGC];
	for (;;) {
		/* ---- 32 deferred windows (or what is left of them at the very end) ---- */
		if (ndefer >= 32 || (ndefer && !have && count == 0 && g >= g_end)) {
			const uint32_t m = ndefer < 32 ? ndefer : 32;
			const bool run = lane < m;
			const unsigned long long e = run ? defer[ndefer - m + lane] : 0ull;
			const uint64_t dchunk = chunk0 + (uint32_t)(e >> 6);
			if (run) {
				const uint4 *src = reinterpret_cast<const uint4 *>(P.text) + ((int64_t)dchunk - P.gb);
				for (int gi = 0; gi < P.ng; gi++) {
					const uint4 v = __ldg(src + gi);
					my_stage[gi * 4 + 0] = v.x; my_stage[gi * 4 + 1] = v.y; my_stage[gi * 4 + 2] = v.z; my_stage[gi * 4 + 3] = v.w;
				}
			}
			if (REFINE_WINDOW(my_bytes + ws0 + (int)(e & 63ull), run))
				atomicOr(&P.bitmap[dchunk >> 5], 1u << (dchunk & 31));
			ndefer -= m;
			__syncwarp();
			continue;
		}
		if (!have) {
			while (count < 32 && g < g_end) refine_refill(P, ring, head, count, lane, g, g_begin, g_end, next_word);
			if (count == 0) break;                              /* range done, rings empty */
			refine_pop<NGC>(P, ring, head, count, lane, chunk0, chunk, keep, nx);      /* (a partial batch only at the very end) */
		}
		/* ---- the batch in nx[]: into shared memory; the next one starts loading while this one is judged ---- */
		const uint64_t cchunk = chunk; const bool ckeep = keep;
		if (!ckeep) {
#pragma unroll
			for (int gi = 0; gi < NGC; gi++) if (gi < P.ng) {
				my_stage[gi * 4 + 0] = nx[gi].x; my_stage[gi * 4 + 1] = nx[gi].y; my_stage[gi * 4 + 2] = nx[gi].z; my_stage[gi * 4 + 3] = nx[gi].w;
			}
		}
		while (count < 32 && g < g_end) refine_refill(P, ring, head, count, lane, g, g_begin, g_end, next_word);
		have = count != 0;                                      /* fewer than 32 only when the range is exhausted */
		if (have) refine_pop<NGC>(P, ring, head, count, lane, chunk0, chunk, keep, nx);
		uint64_t starts = 0;
		if (!ckeep) starts = P.poly ? window_starts_na<true>(my_stage + P.gb * 4, P) : window_starts_na<false>(my_stage + P.gb * 4, P);
		bool pass = ckeep;
		{
			const bool run = starts != 0;
			const bool ok = REFINE_WINDOW(my_bytes + ws0 + (run ? __ffsll((long long)starts) - 1 : 32), run);
			if (!ckeep && !ok) atomicAnd(&P.bitmap[cchunk >> 5], ~(1u << (cchunk & 31)));      /* undecided chunks lose the bit now ... */
			if (ok) pass = true;
		}
		/* ... and get it back if one of their other windows passes later */
		uint64_t rest = (starts && !pass) ? (starts & (starts - 1)) : 0ull;
		for (;;) {
			const uint32_t pend = __ballot_sync(0xffffffffu, rest != 0);
			if (!pend) break;
			if (ndefer + __popc(pend) > REFINE_DEFER) break;          /* ring full: see below */
			if (rest != 0) {
				defer[ndefer + __popc(pend & lt_mask)] = ((unsigned long long)(uint32_t)(cchunk - chunk0) << 6) | (unsigned long long)(__ffsll((long long)rest) - 1);
				rest &= rest - 1;
			}
			ndefer += __popc(pend);
		}
		if (rest != 0) atomicOr(&P.bitmap[cchunk >> 5], 1u << (cchunk & 31));     /* could not be queued: keep (stage 2 is exact) */
		__syncwarp();
	}
#undef REFINE_WINDOW
}

template <typename T, int NR, bool COSTS>
static void launch_refine_one(const RefineParams &P, unsigned grid, cudaStream_t st)
{
	const size_t smem = (size_t)REFINE_THREADS * (P.ng * 4 + 1) * sizeof(uint32_t);
	if (P.ng <= 4) k_refine<T, NR, COSTS, 4><<<grid, REFINE_THREADS, smem, st>>>(P);
	else k_refine<T, NR, COSTS, REFINE_MAXG><<<grid, REFINE_THREADS, smem, st>>>(P);
}
template <typename T, bool COSTS>
static int launch_refine_t(int nrows, const RefineParams &P, unsigned grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: launch_refine_one<T, 1, COSTS>(P, grid, st); break;
	case 2: launch_refine_one<T, 2, COSTS>(P, grid, st); break;
	case 3: launch_refine_one<T, 3, COSTS>(P, grid, st); break;
	case 4: launch_refine_one<T, 4, COSTS>(P, grid, st); break;
	case 5: launch_refine_one<T, 5, COSTS>(P, grid, st); break;
	case 6: launch_refine_one<T, 6, COSTS>(P, grid, st); break;
	case 7: launch_refine_one<T, 7, COSTS>(P, grid, st); break;
	case 8: launch_refine_one<T, 8, COSTS>(P, grid, st); break;
	case 9: launch_refine_one<T, 9, COSTS>(P, grid, st); break;
	default: return -1;
	}
	g_launches++;
	return 0;
}

static bool refine_geometry(const agb_desc &d, RefineParams &P)
{
	if (!front_usable(d) || !d.refine) return false;
	int max_off = 0, min_off = 1 << 30;
	for (int i = 0; i < d.n_anchors; i++) { max_off = std::max(max_off, (int)d.anchor_off[i]); min_off = std::min(min_off, (int)d.anchor_off[i]); }
	P.lo_off = max_off + d.k;
	P.hi_off = 15 + d.pat_len - min_off + d.k;
	P.gb = (P.lo_off + 15) / 16;
	P.ng = P.gb + (P.hi_off + 15) / 16;
	if (P.ng < P.gb + 2) P.ng = P.gb + 2;              /* the chunk itself and the word that follows it */
	return P.ng <= REFINE_MAXG && max_off <= 31;
}

/* stage 1.5 over the whole bitmap */
int refine_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, cudaStream_t st)
{
	RefineParams P; memset(&P, 0, sizeof P);
	if (n == 0 || !refine_geometry(d, P)) return AGB_OK;
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32;
	P.text = (const uint8_t *)d_text; P.bitmap = W.bitmap; P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.fold = d.anchor_fold; P.amask = d.anchor_mask; P.na = d.n_anchors;
	for (int i = 0; i < d.n_anchors; i++) { P.anchor[i] = d.anchor[i]; P.off[i] = d.anchor_off[i]; }
	{   /* stage 1's polynomial over the anchors, usable when they are pairwise distinct and pass its false-positive guard */
		bool distinct = true;
		for (int i = 0; i < d.n_anchors; i++) for (int j = 0; j < i; j++) if (d.anchor[i] == d.anchor[j]) distinct = false;
		P.one = 1; P.scale = 1;
		for (int i = d.anchor_len; i < 4; i++) P.scale <<= 8;
		P.poly = (distinct && poly_setup(d.anchor, d.n_anchors, 8 * d.anchor_len, P.coef)) ? 1 : 0;
	}
	const uint64_t groups = (n_words + 31) / 32;
	unsigned grid = (unsigned)std::min<uint64_t>((groups + 3) / 4, (uint64_t)W.sm_count * 16);
	if (!grid) grid = 1;
	const bool costs = d.engine == AGB_ENGINE_ASEARCH1, narrow = d.M <= 31;
	int rc = costs ? (narrow ? launch_refine_t<uint32_t, true>(d.nrows, P, grid, st) : launch_refine_t<uint64_t, true>(d.nrows, P, grid, st))
	               : (narrow ? launch_refine_t<uint32_t, false>(d.nrows, P, grid, st) : launch_refine_t<uint64_t, false>(d.nrows, P, grid, st));
	if (rc) return AGB_ERR_ARG;
	CUDA_TRY(cudaGetLastError());
	return AGB_OK;
}

