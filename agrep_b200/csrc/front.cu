/* agrep_b200/csrc/front.cu -- stage 1: the anchor front-end, the HBM-bound kernel */
#include "scan_internal.cuh"
#include "tma.cuh"

/* ================================================================================================
 * stage 1: anchor front-end
 * ============================================================================================== */

/* The 4 windows that start in word `lo` (bytes 0..3; `hi` = the following word) against the NA anchors.
 * Result: acc stays non-zero unless some window equals some anchor.
 *
 * POLY: f(w) = prod_i (w - A_i) mod 2^32, evaluated by Horner -- NA IMADs on the FMA pipe per window and
 * half a VIMNMX3 on the ALU pipe, instead of NA compare-class ALU ops.  w == A_i  =>  f(w) == 0 exactly
 * (ring identity), so the filter never loses a match; f(w) == 0 without an equal factor needs the 2-adic
 * valuations of the NA differences to add up to 32, which front_launch() rules out up front (it falls back
 * to the compare form when anchors share low-order bytes).  Anchors shorter than 4 bytes: f is scaled by
 * 256^(4-len), which zeroes exactly when the low len bytes agree.
 * !POLY: unsigned min of the differences (VIADDMNMX), one ALU op per window and anchor. */
template <int NA, bool MASKED, bool POLY, int N3>
__device__ __forceinline__ uint32_t windows_test(uint32_t lo, uint32_t hi, const FrontParams &P, uint32_t acc)
{
	uint32_t w[4] = { lo, __funnelshift_r(lo, hi, 8), __funnelshift_r(lo, hi, 16), __funnelshift_r(lo, hi, 24) };
	if (POLY && N3 > 0) {
		/* mixed plan: NA anchors of four bytes and N3 of three (a piece of the pattern that is only three bytes long, or
		 * whose rare gram is).  Two polynomials -- one product would cut every anchor to its low three bytes -- the second
		 * scaled by 256 so that it vanishes exactly when the low three bytes agree; one VIMNMX3 per window takes both. */
#pragma unroll
		for (int t = 0; t < 4; t++) {
			uint32_t r = (NA >= 2) ? __viaddmin_u32(w[t], P.coef[NA - 1], 0xFFFFFFFFu) : w[t] * P.one + P.coef[NA - 1];
#pragma unroll
			for (int i = NA - 2; i >= 0; i--) r = r * w[t] + P.coef[i];
			uint32_t q;
			if (N3 == 1) q = w[t] * P.s256 + P.coef3[0];             /* 256 (w - B): coef3[0] = -256 B; s256: a run-time 256, so that this stays an IMAD (FMA pipe) and does not become a shift-add on the ALU pipe, which the min and funnel shifts already fill */
			else {
				q = w[t] * P.one + P.coef3[N3 - 1];
#pragma unroll
				for (int i = N3 - 2; i >= 0; i--) q = q * w[t] + P.coef3[i];
				q *= P.s256;
			}
			acc = __vimin3_u32(acc, r, q);
		}
		return acc;
	}
	if (POLY) {
		uint32_t f[4];
#pragma unroll
		for (int t = 0; t < 4; t++) {
			/* Horner.  IMAD (FMA pipe, 64 lanes/clk/SM) and the ALU pipe (64 lanes/clk/SM) both count: with three
			 * or more anchors the first step, w + c, goes to the ALU pipe as VIADDMNMX (min(w + c, ~0)), which
			 * leaves NA-1 IMADs per window; `one` is a runtime 1 that keeps the step an IMAD otherwise */
			uint32_t r = (NA >= 3) ? __viaddmin_u32(w[t], P.coef[NA - 1], 0xFFFFFFFFu) : w[t] * P.one + P.coef[NA - 1];
#pragma unroll
			for (int i = NA - 2; i >= 0; i--) r = r * w[t] + P.coef[i];
			f[t] = MASKED ? r * P.scale : r;
		}
		acc = __vimin3_u32(acc, f[0], f[1]);
		acc = __vimin3_u32(acc, f[2], f[3]);
	} else {
		if (MASKED) { w[0] &= P.amask; w[1] &= P.amask; w[2] &= P.amask; w[3] &= P.amask; }
#pragma unroll
		for (int a = 0; a < NA; a++) {
			uint32_t A = P.anchor[a];
			acc = __vimin3_u32(acc, w[0] - A, w[1] - A);
			acc = __vimin3_u32(acc, w[2] - A, w[3] - A);
		}
	}
	return acc;
}

/* the FRONT_CH chunks a thread takes from one stage; FULL = no chunk of the stage is near the end of the text */
template <int NA, bool MASKED, bool FOLD, bool POLY, bool FULL, bool COUNT, int N3>
__device__ __forceinline__ void front_chunks(const FrontParams &P, const uint8_t *st, uint32_t tid, uint32_t lane, uint32_t rem, uint32_t *bm, uint16_t *nlb)
{
#pragma unroll
	for (int c = 0; c < FRONT_CH; c++) {
		const uint32_t idx = c * FRONT_THREADS + tid;
		uint4 v = *reinterpret_cast<const uint4 *>(st + idx * 16);
		/* the first word of the next chunk (a 4-way bank conflict, measured cheaper than SHFL + a predicated LDS:
		 * 4905 vs 4787 GB/s, profiles/round1_front_variants.md) */
		uint32_t x4 = *reinterpret_cast<const uint32_t *>(st + idx * 16 + 16);
		if (COUNT) {
			/* -n: the delimiter bytes of this chunk (SWAR: 0x80 where a byte equals the delimiter), summed over the warp =
			 * one 512-byte block of the ordinals pass (aux.cu), which then need not read the text again */
			const uint32_t xs[4] = { v.x, v.y, v.z, v.w };
			uint32_t cn = 0;
#pragma unroll
			for (int w = 0; w < 4; w++) {
				const uint32_t t = xs[w] ^ P.delim4;
				uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);
				if (!FULL) {                                     /* only bytes of the text */
					/* rem chunks are left from the start of the stage, the last one padded: bytes of this word inside the text */
					const int64_t valid = ((int64_t)rem - (int64_t)idx) * 16 - (int64_t)(P.n_chunks * 16 - P.n) - 4 * w;
					if (valid <= 0) z = 0; else if (valid < 4) z &= (1u << (8 * (uint32_t)valid)) - 1u;
				}
				cn += __popc(z);
			}
			const uint32_t blk = __reduce_add_sync(0xffffffffu, cn);
			if (lane == 0 && (FULL || idx < rem)) nlb[c * (FRONT_THREADS / 32)] = (uint16_t)blk;
		}
		if (FOLD) { v.x |= P.fold; v.y |= P.fold; v.z |= P.fold; v.w |= P.fold; x4 |= P.fold; }
		uint32_t acc = 0xffffffffu;
		acc = windows_test<NA, MASKED, POLY, N3>(v.x, v.y, P, acc);
		acc = windows_test<NA, MASKED, POLY, N3>(v.y, v.z, P, acc);
		acc = windows_test<NA, MASKED, POLY, N3>(v.z, v.w, P, acc);
		acc = windows_test<NA, MASKED, POLY, N3>(v.w, x4, P, acc);
		if (FULL) {
			const uint32_t word = __ballot_sync(0xffffffffu, acc == 0);
			if (lane == 0) bm[c * (FRONT_THREADS / 32)] = word;
		} else {
			/* the last chunks are always passed on: a match may run into the delimiter appended at EOF
			 * (bitap.c:161-165), and their look-ahead bytes may not exist */
			const bool flag = (idx < rem) && (acc == 0 || idx + 2 >= rem);
			const uint32_t word = __ballot_sync(0xffffffffu, flag);
			if (lane == 0 && idx < rem) bm[c * (FRONT_THREADS / 32)] = word;
		}
	}
}

/* Persistent CTAs.  Thread 0 keeps FRONT_NST bulk copies of 16 KiB (+16 B) in flight into a shared-memory
 * ring, each completing on its own mbarrier; all 256 threads take 4 chunks per stage from shared memory
 * (LDS.128, conflict-free: a warp reads 512 consecutive bytes), test the 16 windows of each chunk and ballot
 * the 32 verdicts of a warp into one bitmap word.  Every text byte crosses HBM->SM once. */
template <int NA, bool MASKED, bool FOLD, bool POLY, bool COUNT, int N3>
__global__ void __launch_bounds__(FRONT_THREADS, FRONT_CTAS_PER_SM)
k_front(const FrontParams P)
{
	extern __shared__ __align__(128) uint8_t s_ring[];
	__shared__ uint64_t s_bar[FRONT_NST];
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	if (tid == 0) {
		for (int i = 0; i < FRONT_NST; i++) mbar_init(&s_bar[i], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto issue = [&](uint64_t it) {
		const uint64_t sg = P.stage_begin + blockIdx.x + it * gridDim.x;
		if (sg >= P.stage_end) return;
		const uint32_t slot = (uint32_t)(it % FRONT_NST);
		const uint64_t off = sg * FRONT_STAGE_BYTES, avail = P.readable - off;
		const uint32_t bytes = (uint32_t)(avail < FRONT_SLOT_BYTES ? (avail & ~15ull) : FRONT_SLOT_BYTES);
		mbar_expect_tx(&s_bar[slot], bytes);
		bulk_g2s(s_ring + slot * FRONT_SLOT_BYTES, P.text + off, bytes, &s_bar[slot]);
	};
	if (tid == 0) for (int i = 0; i < FRONT_NST; i++) issue(i);
	const uint32_t warp_in_cta = tid >> 5;
	for (uint32_t it = 0;; it++) {
		const uint64_t sg = P.stage_begin + blockIdx.x + (uint64_t)it * gridDim.x;
		if (sg >= P.stage_end) break;
		const uint32_t slot = it % FRONT_NST;
		mbar_wait(&s_bar[slot], (it / FRONT_NST) & 1u);
		const uint8_t *st = s_ring + slot * FRONT_SLOT_BYTES;
		/* per-stage scalars, so that the per-chunk bookkeeping below is 32-bit */
		const uint64_t left = P.n_chunks - sg * FRONT_STAGE_CHUNKS;                    /* chunks from the start of this stage to EOF */
		const uint32_t rem = left > 0xFFFF0000ull ? 0xFFFF0000u : (uint32_t)left;
		uint32_t *bm = P.bitmap + sg * FRONT_WORDS_PER_STAGE + warp_in_cta;
		/* full = every chunk of the stage exists and none is among the last two of the text: no per-chunk EOF logic */
		const bool full = left >= FRONT_STAGE_CHUNKS + 2;
		uint16_t *nlb = COUNT ? P.nl_blocks + sg * FRONT_WORDS_PER_STAGE + warp_in_cta : nullptr;
		if (full) front_chunks<NA, MASKED, FOLD, POLY, true, COUNT, N3>(P, st, tid, lane, rem, bm, nlb);
		else front_chunks<NA, MASKED, FOLD, POLY, false, COUNT, N3>(P, st, tid, lane, rem, bm, nlb);
		__syncthreads();                       /* everyone is done reading this slot */
		if (tid == 0) issue((uint64_t)it + FRONT_NST);   /* refill it with the stage FRONT_NST iterations ahead */
	}
}

template <int NA, bool MASKED, bool FOLD, bool POLY, bool COUNT, int N3>
static void launch_front_cnt(const FrontParams &P, unsigned grid, cudaStream_t st)
{
	static bool configured[64] = {false};
	int dev = 0; cudaGetDevice(&dev);
	if (!configured[dev & 63]) {
		cudaFuncSetAttribute(k_front<NA, MASKED, FOLD, POLY, COUNT, N3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FRONT_SMEM);
		configured[dev & 63] = true;
	}
	k_front<NA, MASKED, FOLD, POLY, COUNT, N3><<<grid, FRONT_THREADS, FRONT_SMEM, st>>>(P);
}
template <int NA, bool MASKED, bool FOLD, bool POLY, int N3>
static void launch_front_one(const FrontParams &P, unsigned grid, cudaStream_t st)
{
	if (P.nl_blocks) launch_front_cnt<NA, MASKED, FOLD, POLY, true, N3>(P, grid, st);
	else launch_front_cnt<NA, MASKED, FOLD, POLY, false, N3>(P, grid, st);
}
template <int NA, bool POLY>
static void launch_front_na(const FrontParams &P, bool masked, bool fold, unsigned grid, cudaStream_t st)
{
	if (masked) { if (fold) launch_front_one<NA, true, true, POLY, 0>(P, grid, st); else launch_front_one<NA, true, false, POLY, 0>(P, grid, st); }
	else        { if (fold) launch_front_one<NA, false, true, POLY, 0>(P, grid, st); else launch_front_one<NA, false, false, POLY, 0>(P, grid, st); }
}
/* mixed plans: four-byte anchors by the polynomial + one or two three-byte anchors */
template <int NA>
static void launch_front_mixed(const FrontParams &P, bool fold, unsigned grid, cudaStream_t st)
{
	if (P.n3 == 1) { if (fold) launch_front_one<NA, false, true, true, 1>(P, grid, st); else launch_front_one<NA, false, false, true, 1>(P, grid, st); }
	else           { if (fold) launch_front_one<NA, false, true, true, 2>(P, grid, st); else launch_front_one<NA, false, false, true, 2>(P, grid, st); }
}

/* coefficients of prod_i (x - a_i) mod 2^32 and the false-positive guard of the polynomial form:
 * a zero product without a zero factor needs sum_i v2(w - a_i) >= bits; with t = the largest v2(a_i - a_j)
 * at most one factor can exceed t, so w must agree with an anchor in its low bits - (na-1)*t bits.  We ask
 * for at least 20 agreeing bits (a 2.5-byte accidental match) or use the compare form instead. */
bool poly_setup(const uint32_t *a, int na, int bits, uint32_t *coef)
{
	uint32_t c[AGB_MAXANCHOR + 1]; int deg = 0, t = 0;
	memset(c, 0, sizeof c); c[0] = 1;
	for (int i = 0; i < na; i++) {
		uint32_t m = 0u - a[i];
		for (int j = deg + 1; j >= 1; j--) c[j] = c[j - 1] + c[j] * m;
		c[0] = c[0] * m; deg++;
		for (int j = 0; j < i; j++) { uint32_t dd = a[i] - a[j]; int v = dd ? __builtin_ctz(dd) : 32; if (v > t) t = v; }
	}
	for (int i = 0; i < na; i++) coef[i] = c[i];
	return bits - (na - 1) * t >= 20;
}

bool front_usable(const agb_desc &d)
{
	if (d.plan != AGB_PLAN_ANCHORS || d.n_anchors < 1 || d.n_anchors > 9) return false;
	return d.n_anchors3 == 0 || (d.n_anchors3 <= 2 && d.anchor_len == 4 && d.n_anchors <= 7);
}

/* stage 1 over bitmap words [word_begin, word_end) of a text of n bytes; word_begin must be a multiple of 32
 * (a stage is 32 words).  slack16: 16 more bytes after the last chunk are readable (true for our own buffers). */
int front_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n,
                        uint64_t word_begin, uint64_t word_end, bool slack16, cudaStream_t st, bool count_delims)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32;
	if (word_end > n_words) word_end = n_words;
	if (word_begin >= word_end) return AGB_OK;
	FrontParams F; memset(&F, 0, sizeof F);
	F.text = (const uint8_t *)d_text; F.bitmap = W.bitmap; F.n = n; F.n_chunks = n_chunks;
	F.readable = n_chunks * 16 + (slack16 ? 16 : 0);
	F.stage_begin = word_begin / FRONT_WORDS_PER_STAGE;
	F.stage_end = (word_end + FRONT_WORDS_PER_STAGE - 1) / FRONT_WORDS_PER_STAGE;
	F.fold = d.anchor_fold; F.amask = d.anchor_mask;
	F.nl_blocks = count_delims ? W.ord_blocks : nullptr; F.delim4 = d.delim[0] * 0x01010101u;
	const uint64_t stages = F.stage_end - F.stage_begin;
	unsigned grid = (unsigned)std::min<uint64_t>(stages, (uint64_t)W.sm_count * FRONT_CTAS_PER_SM);
	if (!grid) grid = 1;
	bool masked = d.anchor_mask != 0xFFFFFFFFu, fold = d.anchor_fold != 0;
	/* identical anchors (e.g. from "abababab") are tested once */
	int na = 0;
	for (int i = 0; i < d.n_anchors; i++) {
		bool dup = false;
		for (int j = 0; j < na; j++) if (F.anchor[j] == d.anchor[i]) dup = true;
		if (!dup) F.anchor[na++] = d.anchor[i];
	}
	F.one = 1; F.scale = 1; F.s256 = 256;
	for (int i = d.anchor_len; i < 4; i++) F.scale <<= 8;
	bool poly = poly_setup(F.anchor, na, 8 * d.anchor_len, F.coef);
	if (d.n_anchors3 > 0) {
		/* mixed plan: the second polynomial over the three-byte anchors; if either guard fails the three-byte anchors join
		 * the compare form as four-byte... no such form: fall back to flagging on the four-byte group alone is NOT allowed
		 * (it would lose matches), so the compare form below gets the 3-byte anchors through their own mask */
		int n3 = 0; uint32_t a3[4];
		for (int i = 0; i < d.n_anchors3; i++) {
			bool dup = false;
			for (int j = 0; j < n3; j++) if (a3[j] == d.anchor3[i]) dup = true;
			if (!dup) a3[n3++] = d.anchor3[i];
		}
		uint32_t c3[AGB_MAXANCHOR];
		const bool poly3 = poly_setup(a3, n3, 24, c3);
		if (!(poly && poly3)) { snprintf(g_err, sizeof g_err, "internal: mixed anchor plan without a polynomial form"); return AGB_ERR_ARG; }
		F.n3 = n3;
		for (int i = 0; i < n3; i++) F.coef3[i] = c3[i];
		if (n3 == 1) F.coef3[0] = 0u - 256u * a3[0];
		switch (na) {
		case 1: launch_front_mixed<1>(F, fold, grid, st); break;  case 2: launch_front_mixed<2>(F, fold, grid, st); break;
		case 3: launch_front_mixed<3>(F, fold, grid, st); break;  case 4: launch_front_mixed<4>(F, fold, grid, st); break;
		case 5: launch_front_mixed<5>(F, fold, grid, st); break;  case 6: launch_front_mixed<6>(F, fold, grid, st); break;
		case 7: launch_front_mixed<7>(F, fold, grid, st); break;
		default: return AGB_ERR_ARG;
		}
		g_launches++;
		CUDA_TRY(cudaGetLastError());
		return AGB_OK;
	}
#define FRONT_CASE(N) case N: if (poly) launch_front_na<N, true>(F, masked, fold, grid, st); else launch_front_na<N, false>(F, masked, fold, grid, st); break;
	switch (na) {
	FRONT_CASE(1) FRONT_CASE(2) FRONT_CASE(3) FRONT_CASE(4) FRONT_CASE(5) FRONT_CASE(6) FRONT_CASE(7) FRONT_CASE(8) FRONT_CASE(9)
	default: return AGB_ERR_ARG;
	}
#undef FRONT_CASE
	g_launches++;
	CUDA_TRY(cudaGetLastError());
	return AGB_OK;
}

