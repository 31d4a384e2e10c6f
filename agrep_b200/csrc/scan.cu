/* agrep_b200/csrc/scan.cu -- the sm_100a scan path of libagrepb200 and its C ABI (include/agrep_b200.h).
 *
 * What the reference does in bitap()/asearch()/asearch0()/asearch1()/sgrep()+bm() (one byte at a time,
 * one file block at a time, bitap.c:169-284, asearch.c:94-306, :620-774, asearch1.c:86-235,
 * sgrep.c:694-1016) is done here in two device stages over text that is resident in HBM:
 *
 *   stage 1  k_front   "which 16-byte chunks can matter"  -- the HBM-bound kernel.
 *            Coalesced 16-byte loads, every byte read once.  For each chunk it decides whether one of
 *            the pattern's k+1 disjoint literal anchors (agb_desc.anchor[], pigeonhole argument in
 *            pattern.c:plan_anchors) starts inside it; warp ballot packs 32 decisions into one word of
 *            a chunk bitmap (1 bit per 16 bytes of text = 0.8 % write traffic).
 *   stage 2  k_records "which records match"               -- exact, the Wu-Manber recurrence itself.
 *            For every flagged chunk the owning thread finds the record(s) that meet the chunk, runs the
 *            automaton from the record start in the constant post-delimiter state (asearch.c:175-186)
 *            until the record's closing delimiter, applies the reference's match test and record
 *            bookkeeping (bitap.c:177-229, agrep.c:3811), and counts / emits (lasti, print_end).
 *            A record is owned by the first flagged chunk that meets it, so it is reported exactly once
 *            and the output is ordered by construction (count pass -> tile scan -> emit pass).
 *
 * Patterns for which no anchor plan exists (classes everywhere, -v, -p ...) run stage 2 with every
 * chunk flagged.  There is no CPU path in this file.
 */
#include "agrep_b200.h"
#include "pattern_internal.h"
#include "corpus.h"
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <unistd.h>
#include <sys/stat.h>
#include <mutex>
#include <atomic>
#include <algorithm>
#include <thread>

/* ------------------------------------------------------------------------------------------------ */
static thread_local char g_err[512];
static std::atomic<uint64_t> g_launches{0};

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
	snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
	return AGB_ERR_CUDA; } } while (0)

extern "C" const char *agb_last_error(void) { return g_err; }
extern "C" const char *agb_version(void) { return "agrep-b200 0.1 (sm_100a)"; }
extern "C" uint64_t agb_kernel_launches(void) { return g_launches.load(); }
extern "C" int agb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
extern "C" int agb_set_device(int dev) { CUDA_TRY(cudaSetDevice(dev)); return AGB_OK; }

/* ================================================================================================
 * stage 1: anchor front-end
 * ============================================================================================== */
#define FRONT_THREADS 256
#define FRONT_CH      4                                   /* 16-byte chunks per thread and stage              */
#define FRONT_STAGE_CHUNKS (FRONT_THREADS * FRONT_CH)     /* 1024 chunks = 16 KiB = 32 bitmap words per stage */
#define FRONT_STAGE_BYTES  (FRONT_STAGE_CHUNKS * 16)
#define FRONT_SLOT_BYTES   (FRONT_STAGE_BYTES + 16)       /* + the 16 bytes that follow: the last chunk's windows look 3 bytes ahead */
#define FRONT_NST     2                                   /* stages in flight per CTA (32 KiB); 6 CTAs = 48 warps per SM: measured best */
#define FRONT_CTAS_PER_SM 6
#define FRONT_WORDS_PER_STAGE (FRONT_STAGE_CHUNKS / 32)

struct FrontParams {
	const uint8_t *text;         /* 16-byte aligned */
	uint32_t    *bitmap;         /* one word per 32 chunks */
	uint64_t     n;              /* bytes */
	uint64_t     n_chunks;       /* ceil(n/16) */
	uint64_t     readable;       /* bytes that may be read from text: 16 * n_chunks (+16 when the caller's slack allows) */
	uint64_t     stage_begin, stage_end;   /* this launch covers stages [stage_begin, stage_end) of 1024 chunks each */
	uint32_t     fold, amask;
	uint32_t     one, scale;     /* 1 (kept opaque so the first Horner step stays an IMAD) and 256^(4-anchor_len) */
	uint32_t     anchor[AGB_MAXANCHOR];
	uint32_t     coef[AGB_MAXANCHOR];   /* prod_i (x - anchor[i]) mod 2^32, low order first, leading 1 implied */
};

/* ---- bulk-async copy (TMA, SASS UBLKCP) + mbarrier plumbing, shared::cta addressing ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}"
	             :: "r"(smem_u32(b)), "r"(parity) : "memory");
}

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
template <int NA, bool MASKED, bool POLY>
__device__ __forceinline__ uint32_t windows_test(uint32_t lo, uint32_t hi, const FrontParams &P, uint32_t acc)
{
	uint32_t w[4] = { lo, __funnelshift_r(lo, hi, 8), __funnelshift_r(lo, hi, 16), __funnelshift_r(lo, hi, 24) };
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
template <int NA, bool MASKED, bool FOLD, bool POLY, bool FULL>
__device__ __forceinline__ void front_chunks(const FrontParams &P, const uint8_t *st, uint32_t tid, uint32_t lane, uint32_t rem, uint32_t *bm)
{
#pragma unroll
	for (int c = 0; c < FRONT_CH; c++) {
		const uint32_t idx = c * FRONT_THREADS + tid;
		uint4 v = *reinterpret_cast<const uint4 *>(st + idx * 16);
		/* the first word of the next chunk (a 4-way bank conflict, measured cheaper than SHFL + a predicated LDS:
		 * 4905 vs 4787 GB/s, profiles/round1_front_variants.md) */
		uint32_t x4 = *reinterpret_cast<const uint32_t *>(st + idx * 16 + 16);
		if (FOLD) { v.x |= P.fold; v.y |= P.fold; v.z |= P.fold; v.w |= P.fold; x4 |= P.fold; }
		uint32_t acc = 0xffffffffu;
		acc = windows_test<NA, MASKED, POLY>(v.x, v.y, P, acc);
		acc = windows_test<NA, MASKED, POLY>(v.y, v.z, P, acc);
		acc = windows_test<NA, MASKED, POLY>(v.z, v.w, P, acc);
		acc = windows_test<NA, MASKED, POLY>(v.w, x4, P, acc);
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
template <int NA, bool MASKED, bool FOLD, bool POLY>
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
		if (full) front_chunks<NA, MASKED, FOLD, POLY, true>(P, st, tid, lane, rem, bm);
		else front_chunks<NA, MASKED, FOLD, POLY, false>(P, st, tid, lane, rem, bm);
		__syncthreads();                       /* everyone is done reading this slot */
		if (tid == 0) issue((uint64_t)it + FRONT_NST);   /* refill it with the stage FRONT_NST iterations ahead */
	}
}

/* ================================================================================================
 * shared device pieces of stages 1.5 and 2: the recurrence, the match test, the text reader
 * ============================================================================================== */
/* 32-bit rows run on the MIRRORED automaton: every word that holds pattern positions -- character masks, Init0/1,
 * NO_ERR_MASK, endposition, D_endpos, the reset and start rows -- is bit-reversed (__brev) when it is loaded, which
 * turns the recurrence's `>> 1` into `<< 1` and changes nothing else (the kernels only ever AND/OR/compare these
 * words).  A left shift by one is a multiply by two, and IMAD runs on the FMA pipe, which these kernels leave idle,
 * instead of the ALU pipe that bounds them: 5 ALU + 2 FMA operations per row and byte instead of 7 ALU.
 * 64-bit rows (M > 31) stay as the reference has them. */
template <typename T> __device__ __forceinline__ T mirror(T x) { return x; }
template <> __device__ __forceinline__ uint32_t mirror<uint32_t>(uint32_t x) { return __brev(x); }
template <typename T> __device__ __forceinline__ T shift1(T x) { return x >> 1; }
template <> __device__ __forceinline__ uint32_t shift1<uint32_t>(uint32_t x)
{
	uint32_t r;
	asm("mad.lo.u32 %0, %1, 2, 0;" : "=r"(r) : "r"(x));
	return r;
}

template <typename T> struct DevConsts {
	T init1, noerr, endpos, dendpos;
	int L, k, and_mode, inverse, kind, ci, cs, cd;
};

template <typename T, int NR> struct RecShared {
	T mask[257];                 /* mask[256] = 0: "a byte outside the text" */
	T reset[NR], start[NR];
	uint8_t delim[2 * AGB_MAXDELIM + 2];
	unsigned long long hist[AGB_MAXERR + 1];
	int start_closes;
};

template <typename T, int NR>
__device__ __forceinline__ void shared_init(RecShared<T, NR> &S, DevConsts<T> &C, const agb_desc *D, int nthreads)
{
	for (int i = threadIdx.x; i < 256; i += nthreads) S.mask[i] = mirror<T>((T)D->mask[i]);
	if (threadIdx.x == 0) { S.mask[256] = 0; S.start_closes = D->start_closes; }
	if (threadIdx.x < NR) { S.reset[threadIdx.x] = mirror<T>((T)D->reset[threadIdx.x]); S.start[threadIdx.x] = mirror<T>((T)D->start[threadIdx.x]); }
	if (threadIdx.x < 2 * AGB_MAXDELIM + 2) S.delim[threadIdx.x] = D->delim[threadIdx.x];
	if (threadIdx.x <= AGB_MAXERR) S.hist[threadIdx.x] = 0;
	C.init1 = mirror<T>((T)D->init1); C.noerr = mirror<T>((T)D->noerr); C.endpos = mirror<T>((T)D->endpos); C.dendpos = mirror<T>((T)D->dendpos);
	C.L = D->L; C.k = D->k; C.and_mode = D->and_mode; C.inverse = D->inverse; C.kind = D->delim_kind;
	C.ci = D->cost_i; C.cs = D->cost_s; C.cd = D->cost_d;
	__syncthreads();
}

/* text reader with a one-group (16 byte) register cache; positions are file offsets.
 * -1 is the virtual '\n' (bitap.c:140), n..n+L-1 the delimiter appended at EOF (bitap.c:161-165). */
struct Reader {
	const uint8_t *text; uint64_t n; const uint8_t *delim; int L;
	uint4 v; int64_t grp;
	__device__ __forceinline__ void init(const uint8_t *t, uint64_t n_, const uint8_t *d, int L_) { text = t; n = n_; delim = d; L = L_; grp = -1; v = make_uint4(0, 0, 0, 0); }
	__device__ __forceinline__ int get(int64_t p)
	{
		if (p < 0) return p == -1 ? '\n' : 256;
		if ((uint64_t)p >= n) { int64_t o = p - (int64_t)n; return o < L ? delim[o] : 256; }
		int64_t g = p >> 4;
		if (g != grp) { v = __ldg(reinterpret_cast<const uint4 *>(text) + g); grp = g; }
		uint32_t sel = (uint32_t)(p >> 2) & 3u;
		uint32_t w = sel == 0 ? v.x : (sel == 1 ? v.y : (sel == 2 ? v.z : v.w));
		return (int)((w >> (((uint32_t)p & 3u) * 8u)) & 0xFFu);
	}
};

/* is q (file offset, < n) the last byte of a delimiter that closes a record?  kind 0: every occurrence
 * does (no self overlap); kind 1 (c^L, e.g. $$): greedy, non-overlapping from the start of the run of c,
 * the virtual '\n' counting as part of the run (asearch.c:55-57 D_Mask + the reset at :181). */
__device__ __forceinline__ bool delim_ends_at(Reader &R, int64_t q, const uint8_t *delim, int L, int kind)
{
	if (L == 1) return R.get(q) == delim[0];
	if (kind == 0) {
		for (int t = 0; t < L; t++) if (R.get(q - t) != delim[L - 1 - t]) return false;
		return true;
	}
	int c = delim[0];
	if (R.get(q) != c) return false;
	int64_t len = 1, p = q - 1;
	while (p >= -1 && R.get(p) == c) { len++; p--; }
	return (len % L) == 0;
}

/* one text byte through all rows: asearch.c:96-115 (unit costs), asearch1.c:88-97 (COSTS), bitap.c:175-176 (NR = 1) */
template <typename T, int NR, bool COSTS>
__device__ __forceinline__ void rows_step(T (&S)[NR], T cm, const DevConsts<T> &C)
{
	if (!COSTS) {
		T prevB = S[0];
		T prevA = (shift1<T>(prevB) & cm) | (C.init1 & prevB);
#pragma unroll
		for (int r = 1; r < NR; r++) {
			T b = S[r];
			T a = (shift1<T>(b) & cm) | (C.init1 & b) | prevB | (shift1<T>(prevA | prevB) & C.noerr);
			S[r - 1] = prevA; prevA = a; prevB = b;
		}
		S[NR - 1] = prevA;
	} else {
		T A[NR];
		A[0] = (shift1<T>(S[0]) & cm) | (C.init1 & S[0]);
#pragma unroll
		for (int r = 1; r < NR; r++) {
			T bi = (r - C.ci >= 0) ? S[r - C.ci] : (T)0, ad = (r - C.cd >= 0) ? A[r - C.cd] : (T)0, bs = (r - C.cs >= 0) ? S[r - C.cs] : (T)0;
			A[r] = (shift1<T>(S[r]) & cm) | bi | (shift1<T>(ad | bs) & C.noerr) | (C.init1 & S[r]);
		}
#pragma unroll
		for (int r = 0; r < NR; r++) S[r] = A[r];
	}
}

template <typename T>
__device__ __forceinline__ bool match_cond(T r, const DevConsts<T> &C)
{
	/* bitap.c:182, asearch.c:128 -- C precedence: (AND && all) || ((!AND && any) ^ INVERSE) */
	if (C.and_mode) return ((r & C.endpos) == C.endpos) || (C.inverse != 0);
	return ((r & C.endpos) != 0) != (C.inverse != 0);
}

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
#define REFINE_THREADS 128
#define REFINE_MAXG 8
#define REFINE_DEFER 96          /* deferred windows per warp */
struct RefineParams {
	const uint8_t *text; uint32_t *bitmap; uint64_t n, n_chunks, n_words;
	const agb_desc *desc;
	uint32_t fold, amask; int na;
	int gb, ng;                  /* groups staged before the chunk, groups staged in total (<= REFINE_MAXG) */
	int lo_off, hi_off;          /* the windows of a chunk at byte `base` lie inside [base - lo_off, base + hi_off) */
	uint32_t anchor[AGB_MAXANCHOR]; int32_t off[AGB_MAXANCHOR];
	uint32_t coef[AGB_MAXANCHOR]; uint32_t one, scale; int poly;   /* stage 1's polynomial, to spot the candidate windows cheaply */
};

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

/* ================================================================================================
 * stage 2: records
 * ============================================================================================== */
#define REC_THREADS 128          /* dense form: one thread per bitmap word, a block covers 128*512 B = 64 KiB of text */

struct RecParams {
	const uint8_t  *text;
	const uint32_t *bitmap;      /* NULL: every chunk flagged */
	uint64_t n, n_chunks, n_words;
	const agb_desc *desc;        /* device copy */
	uint32_t *tile_counts;       /* dense: per block; list: per candidate */
	const uint64_t *tile_offsets;/* exclusive scan of tile_counts (emit pass) */
	agb_record *records; uint64_t capacity;
	unsigned long long *totals;  /* [0] matched, [1] flagged chunks, [2..10] level histogram, [12] candidates in the list */
	const uint64_t *cand; uint64_t cand_cap;   /* list form: ordered flagged chunk numbers, totals[12] of them */
	agb_record *cand_first;      /* list form: the first record each candidate reported in the count launch (most report 0 or 1) */
	int emit;                    /* 0: count pass, 1: emit pass */
	int levels;                  /* 1: best-match bookkeeping (smallest matching row) */
	int want_level;              /* levels: report records whose smallest level <= want_level (-1: all matching) */
	int warm;                    /* slices form: bytes of warm-up before a slice (>= positions + rows) */
};

/* The records chunk c owns: a record [s-1, close) belongs to the FIRST flagged chunk that meets it, so
 *   (a) the record that contains byte 16c is ours iff its re-fed byte s-1 lies after the previous flagged chunk
 *       (search backwards, stop at a delimiter end -> ours, or at a flagged chunk -> theirs);
 *   (b) every record whose re-fed byte lies inside the chunk is ours.
 * done_until (dense form) remembers how far this thread's previous chunk already got.
 * Returns the number of reported records; writes them at out_pos.. when write is set. */
template <typename T, int NR, bool COSTS>
__device__ __forceinline__ uint32_t chunk_records(const RecParams &P, const DevConsts<T> &C, RecShared<T, NR> &SH, Reader &R,
                                                 const int64_t c, int64_t &done_until, const bool write, const uint64_t out_pos, const bool hist,
                                                 agb_record *first_out = nullptr)
{
	const int L = C.L;
	const int64_t n = (int64_t)P.n, lo = c * 16, hi = lo + 15;
	uint32_t cnt = 0;
	int64_t s = -2;                /* record start to run from; -2: none */
	if (done_until > lo) {
		/* the record this thread closed last reaches into this chunk; what starts here starts at done_until */
		if (done_until - 1 <= hi) s = done_until; else return 0;
	} else {
		bool found = false;
		if (c == 0) { s = 0; found = true; }
		for (int64_t cc = c - 1; !found; cc--) {
			if (cc < 0) { s = 0; found = true; break; }
			uint32_t pw = P.bitmap ? P.bitmap[cc >> 5] : 0xffffffffu;
			if (pw >> (cc & 31) & 1u) break;                  /* an earlier flagged chunk meets that record: not ours */
			for (int64_t q = cc * 16 + 15; q >= cc * 16; q--)
				if (delim_ends_at(R, q, SH.delim, L, C.kind)) { s = q + 1; found = true; break; }
		}
		if (!found) {
			for (int64_t q = lo; q <= hi && q < n; q++)
				if (delim_ends_at(R, q, SH.delim, L, C.kind)) { s = q + 1; break; }
		}
	}
	/* run records while their re-fed byte (s-1) is at or before the end of this chunk */
	while (s >= 0 && s - 1 <= hi && s <= n) {
		T S[NR];
		int64_t begin;
		if (s == 0) {
#pragma unroll
			for (int r = 0; r < NR; r++) S[r] = SH.start[r];
			begin = SH.start_closes ? -(int64_t)L : 0;
		} else {
#pragma unroll
			for (int r = 0; r < NR; r++) S[r] = SH.reset[r];
			begin = s - L;
		}
		int64_t p = s, close_at = -1;
		const int64_t limit = n + L;
		for (; p < limit; p++) {
			rows_step<T, NR, COSTS>(S, SH.mask[R.get(p)], C);
			if (S[0] & C.dendpos) { close_at = p; break; }
		}
		if (close_at < 0) { done_until = limit + 1; break; }           /* never closed: dropped, as the reference does */
		const int64_t end = close_at + 1 - L;
		const bool counts = (begin + 1 < n) && (begin + 1 <= end);       /* bitap.c:213 + agrep.c:3811 */
		int level = C.k;
		bool cond;
		if (P.levels) {
			level = -1;
#pragma unroll
			for (int r = 0; r < NR; r++) if (level < 0 && match_cond<T>(S[r], C)) level = r;
			cond = level >= 0;
			if (cond && counts && hist) atomicAdd(&SH.hist[level], 1ull);
			if (cond && P.want_level >= 0 && level > P.want_level) cond = false;
		} else cond = match_cond<T>(S[NR - 1], C);
		if (cond && counts) {
			if (write) {
				const uint64_t at = out_pos + cnt;
				if (at < P.capacity) {
					agb_record rec; rec.begin = begin; rec.end = end; rec.ordinal = 0; rec.level = level; rec.pad = 0;
					P.records[at] = rec;
				}
			}
			if (first_out && cnt == 0) { first_out->begin = begin; first_out->end = end; first_out->ordinal = 0; first_out->level = level; first_out->pad = 0; }
			cnt++;
		}
		s = close_at + 1;
		done_until = s;
	}
	return cnt;
}

/* dense form: every thread owns one bitmap word (32 chunks); used when the plan flags everything or stage 1.5
 * cannot thin the bitmap.  Count pass -> per-tile counts; emit pass recounts, scans inside the block, writes. */
template <typename T, int NR, bool COSTS>
__global__ void __launch_bounds__(REC_THREADS)
k_records(const RecParams P)
{
	__shared__ RecShared<T, NR> SH;
	__shared__ uint32_t s_scan[REC_THREADS];
	DevConsts<T> C;
	shared_init<T, NR>(SH, C, P.desc, REC_THREADS);
	const uint64_t gw = (uint64_t)blockIdx.x * REC_THREADS + threadIdx.x;     /* bitmap word of this thread */
	uint32_t word = 0;
	if (gw < P.n_words) {
		word = P.bitmap ? P.bitmap[gw] : 0xffffffffu;
		uint64_t rem = P.n_chunks - gw * 32;
		if (rem < 32) word &= (1u << rem) - 1u;
	}
	Reader R; R.init(P.text, P.n, SH.delim, C.L);
	uint32_t my_count = 0;
	uint64_t out_pos = 0;
	for (int pass = 0; pass < (P.emit ? 2 : 1); pass++) {
		uint32_t bits = word, cnt = 0;
		int64_t done_until = INT64_MIN;
		while (bits) {
			const int b = __ffs(bits) - 1; bits &= bits - 1;
			cnt += chunk_records<T, NR, COSTS>(P, C, SH, R, (int64_t)(gw * 32 + b), done_until, pass == 1, out_pos + cnt, pass == 0 && !P.emit);
		}
		if (pass == 0) {
			my_count = cnt;
			s_scan[threadIdx.x] = cnt;
			__syncthreads();
			for (int off = 1; off < REC_THREADS; off <<= 1) {
				uint32_t v = (threadIdx.x >= (unsigned)off) ? s_scan[threadIdx.x - off] : 0;
				__syncthreads();
				s_scan[threadIdx.x] += v;
				__syncthreads();
			}
			if (!P.emit) {
				if (threadIdx.x == REC_THREADS - 1) {
					P.tile_counts[blockIdx.x] = s_scan[REC_THREADS - 1];
					if (s_scan[REC_THREADS - 1]) atomicAdd(&P.totals[0], (unsigned long long)s_scan[REC_THREADS - 1]);
				}
				uint32_t fl = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(word));
				if ((threadIdx.x & 31) == 0 && fl) atomicAdd(&P.totals[1], (unsigned long long)fl);
				__syncthreads();
				if (P.levels && threadIdx.x <= AGB_MAXERR && SH.hist[threadIdx.x]) atomicAdd(&P.totals[2 + threadIdx.x], SH.hist[threadIdx.x]);
			} else {
				out_pos = P.tile_offsets[blockIdx.x] + (s_scan[threadIdx.x] - my_count);
			}
		}
	}
}

/* dense tile form: the automaton over EVERYTHING (no anchor plan: classes, -v, -p, '#', short patterns ...).
 * One CTA per 32 KiB tile, brought into shared memory (+2 KiB that follow it) by one bulk-async copy.  Thread t
 * owns the records whose opening delimiter ends inside its 128-byte slice: it starts at the first of them in the
 * constant post-delimiter state and simply keeps walking -- a record that closes at a delimiter inside the slice
 * hands over to the next one at the following byte -- until the last of its records has closed (on average half a
 * record past the slice; the neighbour skips that head).  So every lane walks about the same number of bytes in
 * lockstep, bytes and the Mask[] table come from shared memory (global memory only for a record that outruns the
 * staged bytes), and nothing is carried between threads or tiles.  Same loop as asearch.c:94-199. */
#define DENSE_THREADS 256
#define DENSE_TILE    32768
#define DENSE_TAIL    2048
#define DENSE_PER     (DENSE_TILE / DENSE_THREADS)          /* 128 bytes per thread */

template <typename T, int NR, bool COSTS>
__global__ void __launch_bounds__(DENSE_THREADS)
k_records_dense(const RecParams P)
{
	extern __shared__ __align__(128) uint8_t s_text[];                 /* DENSE_TILE + DENSE_TAIL */
	__shared__ RecShared<T, NR> SH;
	__shared__ uint64_t s_bar;
	__shared__ uint32_t s_scan[DENSE_THREADS];
	const uint32_t tid = threadIdx.x;
	const int64_t n = (int64_t)P.n, tile0 = (int64_t)blockIdx.x * DENSE_TILE;
	const uint64_t readable = P.n_chunks * 16;
	const uint64_t avail = (readable - (uint64_t)tile0) & ~15ull;
	const uint32_t loaded = (uint32_t)(avail < (uint64_t)(DENSE_TILE + DENSE_TAIL) ? avail : (uint64_t)(DENSE_TILE + DENSE_TAIL));
	if (tid == 0) {
		mbar_init(&s_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(&s_bar, loaded);
		bulk_g2s(s_text, P.text + tile0, loaded, &s_bar);
	}
	DevConsts<T> C;
	shared_init<T, NR>(SH, C, P.desc, DENSE_THREADS);                  /* ends with __syncthreads(): the barrier init is visible */
	mbar_wait(&s_bar, 0);
	const int L = C.L;
	/* bytes [tile0, tile0 + in_smem) come from shared memory: staged AND inside the text */
	const uint32_t in_smem = (uint32_t)((int64_t)loaded < n - tile0 ? (int64_t)loaded : n - tile0);
	const uint32_t tile_len = (uint32_t)((int64_t)DENSE_TILE < n - tile0 ? (int64_t)DENSE_TILE : n - tile0);
	Reader R; R.init(P.text, P.n, SH.delim, L);

	/* ---- which delimiters end in my slice (bit j: at byte 128 t + j)?  The record that follows each is mine. ---- */
	uint64_t bits[DENSE_PER / 64];
#pragma unroll
	for (int w = 0; w < DENSE_PER / 64; w++) bits[w] = 0;
	if (L == 1) {
		/* 16 bytes at a time: exact per-byte equality by SWAR, 4 flags gathered by one multiply */
		const uint32_t d4 = SH.delim[0] * 0x01010101u;
#pragma unroll
		for (int v = 0; v < DENSE_PER / 16; v++) {
			const uint4 x = *reinterpret_cast<const uint4 *>(s_text + tid * DENSE_PER + v * 16);
			const uint32_t xs[4] = { x.x, x.y, x.z, x.w };
			uint32_t m16 = 0;
#pragma unroll
			for (int w = 0; w < 4; w++) {
				const uint32_t t = xs[w] ^ d4;
				const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);    /* 0x80 where the byte equals the delimiter */
				m16 |= ((((z >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * w);
			}
			bits[v >> 2] |= (uint64_t)m16 << (16 * (v & 3));
		}
	} else {
		for (uint32_t j = 0; j < DENSE_PER; j++) {
			const int64_t q = tile0 + (int64_t)tid * DENSE_PER + j;
			if (q < n && delim_ends_at(R, q, SH.delim, L, C.kind)) bits[j >> 6] |= 1ull << (j & 63);
		}
	}
	{   /* only delimiters inside the text (q < n) */
		const int64_t last_q = (int64_t)tile_len - 1 - (int64_t)tid * DENSE_PER;
#pragma unroll
		for (int w = 0; w < DENSE_PER / 64; w++) {
			const int64_t hi = last_q - 64 * w;
			if (hi < 0) bits[w] = 0; else if (hi < 63) bits[w] &= (2ull << hi) - 1;
		}
	}
	/* the very first record of the text has no delimiter in front of it: thread 0 of tile 0 */
	const bool first = (tile0 == 0 && tid == 0);
	uint32_t owned = first ? 1u : 0u;
#pragma unroll
	for (int w = 0; w < DENSE_PER / 64; w++) owned += __popcll(bits[w]);

	const int64_t limit = n + L;
	/* rows after a delimiter, kept in registers: a close is a handful of moves, not shared-memory traffic */
	T RS[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) RS[r] = SH.reset[r];
	/* away from the end of the text every record counts unless it is empty (bitap.c:213, agrep.c:3811) */
	const bool easy = tile0 + (int64_t)DENSE_TILE + DENSE_TAIL + L + 2 < n;
	uint32_t my_count = 0;
	uint64_t out_pos = 0;
	for (int pass = 0; pass < (P.emit ? 2 : 1); pass++) {
		uint32_t cnt = 0, left = owned;
		if (left) {
			uint32_t rel;                                               /* position relative to tile0 while inside the staged bytes */
			if (first) rel = 0;
			else rel = tid * DENSE_PER + 1 + (bits[0] ? __ffsll((long long)bits[0]) - 1 : 64 + __ffsll((long long)bits[1]) - 1);
			T S[NR];
			int64_t begin;
			if (tile0 == 0 && rel == 0) {
#pragma unroll
				for (int r = 0; r < NR; r++) S[r] = SH.start[r];
				begin = SH.start_closes ? -(int64_t)L : 0;
			} else {
#pragma unroll
				for (int r = 0; r < NR; r++) S[r] = RS[r];
				begin = tile0 + rel - L;
			}
			uint32_t begin_rel = (uint32_t)(begin - tile0);              /* begin - tile0 (mod 2^32; -1 for the virtual newline) */
			/* ---- fast part: text bytes from shared memory, 32-bit bookkeeping.  The plain counting pass (no -B levels,
			 * no list) gets its own loop so that a close is a dozen instructions ---- */
			if (!P.levels && !P.emit) {
				for (; rel < in_smem; rel++) {
					rows_step<T, NR, COSTS>(S, SH.mask[s_text[rel]], C);
					if (S[0] & C.dendpos) {
						const uint32_t end_rel = rel + 1 - L;
						bool counts = (int32_t)(end_rel - begin_rel) >= 1;
						if (!easy) counts = counts && (tile0 + (int64_t)(int32_t)begin_rel + 1 < n);
						cnt += (match_cond<T>(S[NR - 1], C) && counts) ? 1u : 0u;
						if (--left == 0) break;
#pragma unroll
						for (int r = 0; r < NR; r++) S[r] = RS[r];
						begin_rel = end_rel;
					}
				}
			} else
			for (; rel < in_smem; rel++) {
				rows_step<T, NR, COSTS>(S, SH.mask[s_text[rel]], C);
				if (S[0] & C.dendpos) {
					const uint32_t end_rel = rel + 1 - L;
					bool counts = (int32_t)(end_rel - begin_rel) >= 1;  /* begin + 1 <= end (agrep.c:3811) */
					if (!easy) counts = counts && (tile0 + (int64_t)(int32_t)begin_rel + 1 < n);
					int level = C.k;
					bool cond;
					if (P.levels) {
						level = -1;
#pragma unroll
						for (int r = 0; r < NR; r++) if (level < 0 && match_cond<T>(S[r], C)) level = r;
						cond = level >= 0;
						if (cond && counts && pass == 0 && !P.emit) atomicAdd(&SH.hist[level], 1ull);
						if (cond && P.want_level >= 0 && level > P.want_level) cond = false;
					} else cond = match_cond<T>(S[NR - 1], C);
					if (cond && counts) {
						if (pass == 1) {
							const uint64_t at = out_pos + cnt;
							if (at < P.capacity) {
								agb_record rec; rec.begin = tile0 + (int64_t)(int32_t)begin_rel; rec.end = tile0 + end_rel;
								rec.ordinal = 0; rec.level = level; rec.pad = 0;
								P.records[at] = rec;
							}
						}
						cnt++;
					}
					if (--left == 0) break;                               /* the record that starts at the next byte is somebody else's */
#pragma unroll
					for (int r = 0; r < NR; r++) S[r] = RS[r];
					begin_rel = end_rel;
				}
			}
			/* ---- slow part: the record outruns the staged bytes or the text ends (appended delimiter) ---- */
			if (left) {
				begin = tile0 + (int64_t)(int32_t)begin_rel;
				for (int64_t p = tile0 + rel; p < limit; p++) {
					rows_step<T, NR, COSTS>(S, SH.mask[R.get(p)], C);
					if (S[0] & C.dendpos) {
						const int64_t end = p + 1 - L;
						const bool counts = (begin + 1 < n) && (begin + 1 <= end);
						int level = C.k;
						bool cond;
						if (P.levels) {
							level = -1;
#pragma unroll
							for (int r = 0; r < NR; r++) if (level < 0 && match_cond<T>(S[r], C)) level = r;
							cond = level >= 0;
							if (cond && counts && pass == 0 && !P.emit) atomicAdd(&SH.hist[level], 1ull);
							if (cond && P.want_level >= 0 && level > P.want_level) cond = false;
						} else cond = match_cond<T>(S[NR - 1], C);
						if (cond && counts) {
							if (pass == 1) {
								const uint64_t at = out_pos + cnt;
								if (at < P.capacity) {
									agb_record rec; rec.begin = begin; rec.end = end; rec.ordinal = 0; rec.level = level; rec.pad = 0;
									P.records[at] = rec;
								}
							}
							cnt++;
						}
						if (--left == 0) break;
#pragma unroll
						for (int r = 0; r < NR; r++) S[r] = RS[r];
						begin = end;
					}
				}
			}
		}
		if (pass == 0) {
			my_count = cnt;
			s_scan[tid] = cnt;
			__syncthreads();
			for (int off = 1; off < DENSE_THREADS; off <<= 1) {
				uint32_t v = (tid >= (unsigned)off) ? s_scan[tid - off] : 0;
				__syncthreads();
				s_scan[tid] += v;
				__syncthreads();
			}
			if (!P.emit) {
				if (tid == DENSE_THREADS - 1) {
					P.tile_counts[blockIdx.x] = s_scan[DENSE_THREADS - 1];
					if (s_scan[DENSE_THREADS - 1]) atomicAdd(&P.totals[0], (unsigned long long)s_scan[DENSE_THREADS - 1]);
					atomicAdd(&P.totals[1], (unsigned long long)((tile_len + 15) / 16));
				}
				__syncthreads();
				if (P.levels && tid <= AGB_MAXERR && SH.hist[tid]) atomicAdd(&P.totals[2 + tid], SH.hist[tid]);
			} else out_pos = P.tile_offsets[blockIdx.x] + (s_scan[tid] - my_count);
		}
	}
}

/* how dense are the flags?  popcount of every `stride`-th bitmap word (an estimate is all the host needs to pick the
 * record stage's form before it spends time on stage 1.5) */
__global__ void __launch_bounds__(256) k_bitmap_sample(const uint32_t *bitmap, uint64_t n_words, uint32_t stride, unsigned long long *out)
{
	unsigned long long c = 0;
	for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * stride; i < n_words; i += (uint64_t)gridDim.x * blockDim.x * stride)
		c += __popc(bitmap[i]);
	c = __reduce_add_sync(0xffffffffu, (uint32_t)c);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

/* slices form: the automaton over EVERYTHING in lockstep.  The dense tile form above gives every thread whole
 * records, so a warp waits for its longest lane (16 of 32 lanes busy on text lines).  Here every thread walks a
 * fixed slice of SL_PER bytes, whatever the records do:
 *   - the state of the automaton at a slice start depends only on the last M + rows bytes (a state bit moves on or
 *     costs an error every byte) unless the pattern has positions that hold for ever ('#', -p): the thread starts
 *     `warm` bytes early from the post-delimiter rows, throws the results of that stretch away and clears the
 *     sticky end bits -- from there on its rows are exactly the reference's;
 *   - what a record has matched so far lives in the sticky end bits (Init1 keeps them, maskgen.c:232), so for a
 *     record that spans slices the verdict is the OR of the end bits the slices saw: each thread publishes the
 *     end bits left after its last close, and a thread whose first close ends a record it did not see open ORs
 *     the tails of the threads before it back to the one that saw the opening delimiter;
 *   - the record that opens in the tile and closes after it is finished by the tile's last thread, which simply
 *     keeps walking (global memory); the next tile ignores that close.  Nothing is carried between tiles.
 * Text is staged in shared memory in strips of SL_PER + 4 bytes per thread (an odd number of words), so the 32
 * lanes of a warp, which all read the same offset of their strips, hit 32 different banks.
 * Not for: patterns with '#' or -p (unbounded memory), run delimiters ($$: the pairing depends on the start of
 * the run): those keep the dense tile form. */
#define SL_THREADS 128
#define SL_PER     256
#define SL_TILE    (SL_THREADS * SL_PER)        /* 32 KiB: six CTAs per SM, so that the staging of one overlaps the walk of others */
#define SL_APRON   128                          /* bytes staged before the tile: the warm-up of thread 0 */
#define SL_STRIDE  (SL_PER + 4)
#define SL_SMEM    ((SL_THREADS + 1) * SL_STRIDE + 12)

template <typename T, int NR, bool COSTS>
__global__ void __launch_bounds__(SL_THREADS)
k_records_slices(const RecParams P)
{
	extern __shared__ __align__(16) uint8_t s_text[];       /* strip 0: the apron; strip t + 1: thread t's slice */
	__shared__ RecShared<T, NR> SH;
	__shared__ T s_tail[NR][SL_THREADS];                     /* end bits seen since the thread's last close (or its slice start) */
	__shared__ long long s_last[SL_THREADS];                 /* where the thread's last close ended = the begin of the open record */
	__shared__ uint8_t s_has[SL_THREADS];                    /* the thread knows where its open record begins */
	__shared__ uint32_t s_scan[SL_THREADS];
	const uint32_t tid = threadIdx.x;
	const int64_t n = (int64_t)P.n, tile0 = (int64_t)blockIdx.x * SL_TILE, tile_end = tile0 + SL_TILE;
	const int64_t readable = (int64_t)(P.n_chunks * 16);
	DevConsts<T> C;
	shared_init<T, NR>(SH, C, P.desc, SL_THREADS);
	const int L = C.L, warm = P.warm;
	const int64_t limit = n + L;
	/* ---- stage [tile0 - SL_APRON, tile_end): coalesced 16-byte loads, stored into the padded strips ---- */
	for (uint32_t u = tid; u < (SL_APRON + SL_TILE) / 16; u += SL_THREADS) {
		const int64_t g = tile0 - SL_APRON + (int64_t)u * 16;
		if (g >= 0 && g < readable) {
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(P.text + g));
			const uint32_t x = (uint32_t)(g - tile0 + SL_PER);      /* byte number counted from the start of strip 0 */
			uint32_t *dst = reinterpret_cast<uint32_t *>(s_text + (x >> 8) * SL_STRIDE + (x & (SL_PER - 1)));
			dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
		}
	}
	__syncthreads();
	if (tid < (uint32_t)L) {                                 /* the delimiter appended at EOF (bitap.c:161-165) */
		const int64_t g = n + tid;
		if (g >= tile0 - SL_APRON && g < tile_end) {
			const uint32_t x = (uint32_t)(g - tile0 + SL_PER);
			s_text[(x >> 8) * SL_STRIDE + (x & (SL_PER - 1))] = SH.delim[tid];
		}
	}
	__syncthreads();

	const int64_t a = tile0 + (int64_t)tid * SL_PER;         /* my slice: [a, a + SL_PER) */
	const uint8_t *mine = s_text + (tid + 1) * SL_STRIDE, *before = s_text + tid * SL_STRIDE + SL_PER;
	const bool text_start = (tile0 == 0 && tid == 0);
	const bool active = a < limit;
	const uint32_t steps = !active ? 0u : (uint32_t)((limit - a) < (int64_t)SL_PER ? (limit - a) : (int64_t)SL_PER);
	const bool overrun = active && tid == SL_THREADS - 1 && tile_end < limit;    /* finish the record that is open at the end of the tile */
	const bool easy = tile0 > 0 && tile_end + L + 2 < n;
	T RS[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) RS[r] = SH.reset[r];
	Reader R; R.init(P.text, P.n, SH.delim, L);

	/* what the first walk leaves behind */
	bool has_first = false, first_ok = false; int first_level = 0;
	int64_t first_end = 0, first_begin = 0;
	T first_bits[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) first_bits[r] = 0;
	uint32_t my_count = 0;
	uint64_t out_pos = 0;

	for (int pass = 0; pass < (P.emit ? 2 : 1); pass++) {
		const bool writing = pass == 1, tally = (pass == 0 && !P.emit);
		uint32_t cnt = 0;
		T S[NR];
		int64_t begin = 0; bool have_begin = false;
		if (active) {
			if (text_start) {
#pragma unroll
				for (int r = 0; r < NR; r++) S[r] = SH.start[r];
				begin = SH.start_closes ? -(int64_t)L : 0; have_begin = true;
			} else {
#pragma unroll
				for (int r = 0; r < NR; r++) S[r] = RS[r];
				/* four bytes per shared-memory word, their four Mask[] words fetched together: the loads of a group do
				 * not wait for the rows of the group before */
				for (int j = -warm; j < 0; j += 4) {
					const uint32_t w = *reinterpret_cast<const uint32_t *>(before + j);
					T m[4];
#pragma unroll
					for (int i = 0; i < 4; i++) m[i] = SH.mask[(w >> (8 * i)) & 0xFFu];
#pragma unroll
					for (int i = 0; i < 4; i++) {
						rows_step<T, NR, COSTS>(S, m[i], C);
						const bool cl = (S[0] & C.dendpos) != 0;             /* selects, not a branch: see below */
#pragma unroll
						for (int r = 0; r < NR; r++) S[r] = cl ? RS[r] : S[r];
					}
				}
#pragma unroll
				for (int r = 0; r < NR; r++) S[r] &= ~C.endpos;      /* whatever matched before the slice is somebody else's business */
			}
		}
		/* one close: the record [begin, end) is complete */
#define SL_CLOSE(endv) do { \
			const int64_t end_ = (endv); \
			if (!have_begin) { \
				if (pass == 0) { has_first = true; first_end = end_; _Pragma("unroll") for (int r = 0; r < NR; r++) first_bits[r] = S[r] & C.endpos; } \
			} else { \
				const bool counts = (begin + 1 < n) && (begin + 1 <= end_); \
				int level = C.k; bool cond; \
				if (P.levels) { \
					level = -1; \
					_Pragma("unroll") for (int r = 0; r < NR; r++) if (level < 0 && match_cond<T>(S[r], C)) level = r; \
					cond = level >= 0; \
					if (cond && counts && tally) atomicAdd(&SH.hist[level], 1ull); \
					if (cond && P.want_level >= 0 && level > P.want_level) cond = false; \
				} else cond = match_cond<T>(S[NR - 1], C); \
				if (cond && counts) { \
					if (writing) { \
						const uint64_t at = out_pos + cnt; \
						if (at < P.capacity) { agb_record rec; rec.begin = begin; rec.end = end_; rec.ordinal = 0; rec.level = level; rec.pad = 0; P.records[at] = rec; } \
					} \
					cnt++; \
				} \
			} \
			begin = end_; have_begin = true; \
			_Pragma("unroll") for (int r = 0; r < NR; r++) S[r] = RS[r]; \
		} while (0)

		if (easy && !P.levels && !C.and_mode) {
			/* Plain counting away from both ends of the text: every record counts (agrep.c:3811 only bites at the ends).
			 * No branch on a close: one would be taken by one or two lanes in almost every other step of a warp (a line
			 * ends every ~60 bytes) and the divergence costs far more than it skips (measured: 260 cycles per step and
			 * warp).  The loop only resets the rows with selects and shifts two flags per step into a pair of 32-bit
			 * histories -- "a record closed here", "and an end bit was up" (bitap.c:182 without -v; `;` patterns take
			 * the general loop) -- which are counted and located with popc/clz/ffs once per 32 bytes. */
			int first_j = -1, last_j = -1; bool first_found = false;
			const bool inv = C.inverse != 0;
			uint32_t w = *reinterpret_cast<const uint32_t *>(mine);
			for (uint32_t jb = 0; jb < SL_PER; jb += 32) {              /* easy: the whole slice is text */
				uint32_t cw = 0, fw = 0;                                   /* step jb + s  <->  bit 31 - s */
#pragma unroll (NR <= 3 ? 8 : 1)                                    /* many rows: the body is long enough, keep it in the instruction cache */
				for (int g = 0; g < 8; g++) {
					const uint32_t wn = *reinterpret_cast<const uint32_t *>(mine + jb + 4 * g + 4);   /* (the last one reads the strip's padding) */
					T m[4];
#pragma unroll
					for (int i = 0; i < 4; i++) m[i] = SH.mask[(w >> (8 * i)) & 0xFFu];
#pragma unroll
					for (int i = 0; i < 4; i++) {
						rows_step<T, NR, COSTS>(S, m[i], C);
						const bool cl = (S[0] & C.dendpos) != 0;
						cw = cw * 2u + (cl ? 1u : 0u);
						fw = fw * 2u + ((S[NR - 1] & C.endpos) ? 1u : 0u);         /* sticky: only looked at where cw has a bit */
#pragma unroll
						for (int r = 0; r < NR; r++) S[r] = cl ? RS[r] : S[r];
					}
					w = wn;
				}
				if (cw) {
					uint32_t hits = cw & (inv ? ~fw : fw);
					if (last_j < 0) {                                        /* the slice's first close: that record opened before my slice */
						const int sft = __clz(cw);
						first_j = (int)jb + sft; first_found = ((fw >> (31 - sft)) & 1u) != 0;
						hits &= ~(0x80000000u >> sft);
					}
					if (writing) {
						/* the emit pass walks the closes of this word in order: every hit is a record [previous close, this close) */
						int64_t bg = last_j >= 0 ? a + last_j + 1 - L : 0;       /* (the first close of the slice is never a hit here) */
						for (uint32_t c = cw; c; ) {
							const int sft = __clz(c); const uint32_t bit = 0x80000000u >> sft;
							const int64_t en = a + (int64_t)jb + sft + 1 - L;
							if (hits & bit) {
								const uint64_t at = out_pos + cnt;
								if (at < P.capacity) { agb_record rec; rec.begin = bg; rec.end = en; rec.ordinal = 0; rec.level = C.k; rec.pad = 0; P.records[at] = rec; }
								cnt++;
							}
							bg = en; c &= ~bit;
						}
					} else cnt += __popc(hits);
					last_j = (int)jb + 32 - __ffs(cw);
				}
			}
			const T fb = first_found ? C.endpos : (T)0;
			if (first_j >= 0 && pass == 0) { has_first = true; first_end = a + first_j + 1 - L; first_bits[NR - 1] = fb; }
			if (last_j >= 0) { begin = a + last_j + 1 - L; have_begin = true; }
		} else {
			uint32_t w = *reinterpret_cast<const uint32_t *>(mine);
			for (uint32_t j = 0; j < steps; j += 4) {
				const uint32_t wn = *reinterpret_cast<const uint32_t *>(mine + j + 4);
				T m[4];
#pragma unroll
				for (int i = 0; i < 4; i++) m[i] = SH.mask[(w >> (8 * i)) & 0xFFu];
#pragma unroll
				for (int i = 0; i < 4; i++) if (j + i < steps) {
					rows_step<T, NR, COSTS>(S, m[i], C);
					if (S[0] & C.dendpos) SL_CLOSE(a + (int64_t)(j + i) + 1 - L);
				}
				w = wn;
			}
		}
		if (overrun) {
			for (int64_t p = tile_end; p < limit; p++) {
				rows_step<T, NR, COSTS>(S, SH.mask[R.get(p)], C);
				if (S[0] & C.dendpos) { SL_CLOSE(p + 1 - L); break; }
			}
		}
#undef SL_CLOSE
		if (pass == 0) {
			/* ---- the records that span slices ---- */
#pragma unroll
			for (int r = 0; r < NR; r++) s_tail[r][tid] = active ? (T)(S[r] & C.endpos) : (T)0;
			s_last[tid] = begin; s_has[tid] = have_begin ? 1 : 0;
			__syncthreads();
			if (has_first) {
				bool found = false;
				for (int t = (int)tid - 1; t >= 0; t--) {
#pragma unroll
					for (int r = 0; r < NR; r++) first_bits[r] |= s_tail[r][t];
					if (s_has[t]) { first_begin = s_last[t]; found = true; break; }
				}
				if (found) {                                         /* else: it opened in an earlier tile, whose last thread reports it */
					const bool counts = (first_begin + 1 < n) && (first_begin + 1 <= first_end);
					int level = C.k; bool cond;
					if (P.levels) {
						level = -1;
#pragma unroll
						for (int r = 0; r < NR; r++) if (level < 0 && match_cond<T>(first_bits[r], C)) level = r;
						cond = level >= 0;
						if (cond && counts && tally) atomicAdd(&SH.hist[level], 1ull);
						if (cond && P.want_level >= 0 && level > P.want_level) cond = false;
					} else cond = match_cond<T>(first_bits[NR - 1], C);
					first_ok = cond && counts; first_level = level;
				}
			}
			my_count = cnt + (first_ok ? 1u : 0u);
			{   /* inclusive scan of the counts: shuffles inside a warp, the warp totals through shared memory */
				uint32_t inc = my_count;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= (uint32_t)o) inc += v; }
				if ((tid & 31) == 31) s_scan[tid >> 5] = inc;
				__syncthreads();
				uint32_t base = 0;
#pragma unroll
				for (int w = 0; w < SL_THREADS / 32; w++) if ((uint32_t)w < (tid >> 5)) base += s_scan[w];
				__syncthreads();
				s_scan[tid] = base + inc;
				__syncthreads();
			}
			if (!P.emit) {
				if (tid == SL_THREADS - 1) {
					P.tile_counts[blockIdx.x] = s_scan[SL_THREADS - 1];
					if (s_scan[SL_THREADS - 1]) atomicAdd(&P.totals[0], (unsigned long long)s_scan[SL_THREADS - 1]);
					const int64_t tile_len = (n - tile0) < (int64_t)SL_TILE ? (n - tile0) : (int64_t)SL_TILE;
					atomicAdd(&P.totals[1], (unsigned long long)((tile_len + 15) / 16));
				}
				__syncthreads();
				if (P.levels && tid <= AGB_MAXERR && SH.hist[tid]) atomicAdd(&P.totals[2 + tid], SH.hist[tid]);
			} else {
				out_pos = P.tile_offsets[blockIdx.x] + (s_scan[tid] - my_count);
				if (first_ok) {                                      /* the spanning record comes before the thread's own */
					if (out_pos < P.capacity) { agb_record rec; rec.begin = first_begin; rec.end = first_end; rec.ordinal = 0; rec.level = first_level; rec.pad = 0; P.records[out_pos] = rec; }
					out_pos++;
				}
			}
		}
	}
}

/* list form: one thread per surviving chunk of the ordered candidate list (all lanes busy however sparse the
 * survivors are).  Count launch: per-candidate counts; emit launch: writes at the scanned offsets. */
template <typename T, int NR, bool COSTS>
__global__ void __launch_bounds__(REC_THREADS)
k_records_list(const RecParams P)
{
	__shared__ RecShared<T, NR> SH;
	DevConsts<T> C;
	shared_init<T, NR>(SH, C, P.desc, REC_THREADS);
	unsigned long long ncand = P.totals[12];
	if (ncand > P.cand_cap) ncand = P.cand_cap;
	const uint64_t i = (uint64_t)blockIdx.x * REC_THREADS + threadIdx.x;
	uint32_t cnt = 0;
	if (i < ncand) {
		if (P.emit) {
			/* emit launch: 0 records -> nothing; exactly 1 -> the count launch kept it; more (rare) -> walk again */
			const uint32_t c0 = P.tile_counts[i];
			if (c0 == 1) { const uint64_t at = P.tile_offsets[i]; if (at < P.capacity) P.records[at] = P.cand_first[i]; }
			else if (c0 > 1) {
				Reader R; R.init(P.text, P.n, SH.delim, C.L);
				int64_t done_until = INT64_MIN;
				chunk_records<T, NR, COSTS>(P, C, SH, R, (int64_t)P.cand[i], done_until, true, P.tile_offsets[i], false);
			}
		} else {
			Reader R; R.init(P.text, P.n, SH.delim, C.L);
			int64_t done_until = INT64_MIN;
			cnt = chunk_records<T, NR, COSTS>(P, C, SH, R, (int64_t)P.cand[i], done_until, false, 0, true, P.cand_first ? &P.cand_first[i] : nullptr);
			P.tile_counts[i] = cnt;
		}
	}
	if (!P.emit) {
		uint32_t sum = __reduce_add_sync(0xffffffffu, cnt);
		if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&P.totals[0], (unsigned long long)sum);
		__syncthreads();
		if (P.levels && threadIdx.x <= AGB_MAXERR && SH.hist[threadIdx.x]) atomicAdd(&P.totals[2 + threadIdx.x], SH.hist[threadIdx.x]);
	}
}

/* bitmap -> ordered list of flagged chunk numbers: per-block popcounts, scan (k_scan_tiles), scatter */
#define COMPACT_THREADS 256
#define COMPACT_WPT 4            /* words per thread: a block covers 1024 words */
__global__ void __launch_bounds__(COMPACT_THREADS) k_compact_count(const uint32_t *bitmap, uint64_t n_words, uint32_t *block_counts, unsigned long long *totals)
{
	const uint64_t w0 = ((uint64_t)blockIdx.x * COMPACT_THREADS + threadIdx.x) * COMPACT_WPT;
	uint32_t c = 0;
#pragma unroll
	for (int j = 0; j < COMPACT_WPT; j++) if (w0 + j < n_words) c += __popc(bitmap[w0 + j]);
	__shared__ uint32_t s_part[COMPACT_THREADS / 32];
	uint32_t sum = __reduce_add_sync(0xffffffffu, c);
	if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int j = 0; j < COMPACT_THREADS / 32; j++) t += s_part[j];
		block_counts[blockIdx.x] = t;
		if (t) atomicAdd(&totals[1], (unsigned long long)t);
	}
}

__global__ void __launch_bounds__(COMPACT_THREADS) k_compact_write(const uint32_t *bitmap, uint64_t n_words, const uint64_t *block_offsets,
                                                                   uint64_t *list, uint64_t cap)
{
	const uint64_t w0 = ((uint64_t)blockIdx.x * COMPACT_THREADS + threadIdx.x) * COMPACT_WPT;
	uint32_t wd[COMPACT_WPT], c = 0;
#pragma unroll
	for (int j = 0; j < COMPACT_WPT; j++) { wd[j] = (w0 + j < n_words) ? bitmap[w0 + j] : 0u; c += __popc(wd[j]); }
	__shared__ uint32_t s_scan[COMPACT_THREADS];
	s_scan[threadIdx.x] = c;
	__syncthreads();
	for (int off = 1; off < COMPACT_THREADS; off <<= 1) {
		uint32_t v = (threadIdx.x >= (unsigned)off) ? s_scan[threadIdx.x - off] : 0;
		__syncthreads();
		s_scan[threadIdx.x] += v;
		__syncthreads();
	}
	uint64_t at = block_offsets[blockIdx.x] + (s_scan[threadIdx.x] - c);
#pragma unroll
	for (int j = 0; j < COMPACT_WPT; j++)
		for (uint32_t b = wd[j]; b; b &= b - 1) { if (at < cap) list[at] = (w0 + j) * 32 + (uint64_t)(__ffs(b) - 1); at++; }
}

/* exclusive scan of 32-bit counts into 64-bit offsets (one block, coalesced tiles of 4096 with a running carry);
 * the grand total goes to *total when given */
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t *counts, uint64_t *offsets, uint64_t n_tiles, unsigned long long *total)
{
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_carry;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (uint64_t base = 0; base < n_tiles; base += 4096) {
		uint32_t v[4]; unsigned long long sum = 0;
#pragma unroll
		for (int j = 0; j < 4; j++) { const uint64_t i = base + (uint64_t)tid * 4 + j; v[j] = i < n_tiles ? counts[i] : 0u; sum += v[j]; }
		unsigned long long inc = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
		if (lane == 31) s_warp[wid] = inc;
		__syncthreads();
		if (wid == 0) {
			unsigned long long w = s_warp[lane], winc = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= (uint32_t)o) winc += t; }
			s_warp[lane] = winc - w;                 /* exclusive prefix of the warp sums */
		}
		__syncthreads();
		unsigned long long run = s_carry + s_warp[wid] + (inc - sum);
#pragma unroll
		for (int j = 0; j < 4; j++) { const uint64_t i = base + (uint64_t)tid * 4 + j; if (i < n_tiles) offsets[i] = run; run += v[j]; }
		__syncthreads();
		if (tid == 1023) s_carry = run;
		__syncthreads();
	}
	if (total && tid == 0) *total = s_carry;
}

/* two-level exclusive scan for long count arrays (the per-candidate counts of the list form):
 * k_scan_partial sums blocks of 16384 counts, k_scan_tiles scans those sums, k_scan_apply finishes each block */
#define SCAN_BLOCK 16384
__global__ void __launch_bounds__(1024) k_scan_partial(const uint32_t *counts, uint64_t n, uint32_t *block_sums)
{
	__shared__ uint32_t s_w[32];
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK;
	uint32_t sum = 0;
#pragma unroll
	for (int j = 0; j < SCAN_BLOCK / 1024; j++) { const uint64_t i = base + (uint64_t)j * 1024 + threadIdx.x; if (i < n) sum += counts[i]; }
	sum = __reduce_add_sync(0xffffffffu, sum);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
	__syncthreads();
	if (threadIdx.x < 32) { uint32_t v = __reduce_add_sync(0xffffffffu, s_w[threadIdx.x]); if (threadIdx.x == 0) block_sums[blockIdx.x] = v; }
}

__global__ void __launch_bounds__(1024) k_scan_apply(const uint32_t *counts, uint64_t n, const uint64_t *block_offsets, uint64_t *offsets)
{
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_carry;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint64_t base0 = (uint64_t)blockIdx.x * SCAN_BLOCK;
	if (tid == 0) s_carry = block_offsets[blockIdx.x];
	__syncthreads();
	for (uint64_t base = base0; base < base0 + SCAN_BLOCK && base < n; base += 4096) {
		uint32_t v[4]; unsigned long long sum = 0;
#pragma unroll
		for (int j = 0; j < 4; j++) { const uint64_t i = base + (uint64_t)tid * 4 + j; v[j] = i < n ? counts[i] : 0u; sum += v[j]; }
		unsigned long long inc = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
		if (lane == 31) s_warp[wid] = inc;
		__syncthreads();
		if (wid == 0) {
			unsigned long long w = s_warp[lane], winc = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= (uint32_t)o) winc += t; }
			s_warp[lane] = winc - w;
		}
		__syncthreads();
		unsigned long long run = s_carry + s_warp[wid] + (inc - sum);
#pragma unroll
		for (int j = 0; j < 4; j++) { const uint64_t i = base + (uint64_t)tid * 4 + j; if (i < n) offsets[i] = run; run += v[j]; }
		__syncthreads();
		if (tid == 1023) s_carry = run;
		__syncthreads();
	}
}

/* ================================================================================================
 * synthetic corpus
 * ============================================================================================== */
__constant__ char     c_vocab[sizeof(AGB_VOCAB_STR)];
__constant__ uint16_t c_woff[257];
static const char h_vocab[] = AGB_VOCAB_STR;

struct CorpusParams { agb_corpus_spec s; int needle_len; };

#define CORPUS_THREADS 32
#define CORPUS_STRIDE  (AGB_PAGE + 4)     /* +1 word: the 32 generator threads hit different banks */
__global__ void __launch_bounds__(CORPUS_THREADS) k_corpus(uint8_t *out, uint64_t n_pages, const CorpusParams P)
{
	/* one thread generates one 4 KiB page (the generator is inherently sequential) into shared memory,
	 * then the warp writes the 32 pages out with coalesced 128-byte stores */
	extern __shared__ __align__(16) uint8_t s_pages[];
	const uint64_t pg0 = (uint64_t)blockIdx.x * CORPUS_THREADS, pg = pg0 + threadIdx.x;
	if (pg < n_pages)
		agb_corpus_page(s_pages + threadIdx.x * CORPUS_STRIDE, P.s.seed, P.s.first_page + pg, c_vocab, c_woff,
		                P.s.paragraphs, P.s.needle_every, P.s.needle, P.needle_len, P.s.needle_maxedits);
	__syncwarp();
	for (int q = 0; q < CORPUS_THREADS && pg0 + q < n_pages; q++) {
		const uint32_t *src = reinterpret_cast<const uint32_t *>(s_pages + q * CORPUS_STRIDE);
		uint32_t *dst = reinterpret_cast<uint32_t *>(out + (pg0 + q) * AGB_PAGE);
		for (int j = threadIdx.x; j < AGB_PAGE / 4; j += CORPUS_THREADS) dst[j] = src[j];
	}
}

static int corpus_check(const agb_corpus_spec *s, uint16_t *woff)
{
	if (!s || (s->n_bytes % AGB_PAGE) != 0) { snprintf(g_err, sizeof g_err, "corpus size must be a multiple of %d", AGB_PAGE); return AGB_ERR_ARG; }
	if (agb_vocab_offsets(h_vocab, woff) != 256) { snprintf(g_err, sizeof g_err, "vocabulary must hold 256 words"); return AGB_ERR_ARG; }
	return AGB_OK;
}

extern "C" int agb_corpus_fill_device(const agb_corpus_spec *s, void *d_text, void *stream)
{
	uint16_t woff[257]; int rc = corpus_check(s, woff); if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	CUDA_TRY(cudaMemcpyToSymbolAsync(c_vocab, h_vocab, sizeof h_vocab, 0, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemcpyToSymbolAsync(c_woff, woff, sizeof woff, 0, cudaMemcpyHostToDevice, st));
	CorpusParams P; P.s = *s; P.s.needle[63] = 0; P.needle_len = (int)strlen(P.s.needle);
	uint64_t n_pages = s->n_bytes / AGB_PAGE;
	if (n_pages) {
		const int smem = CORPUS_THREADS * CORPUS_STRIDE;
		CUDA_TRY(cudaFuncSetAttribute(k_corpus, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k_corpus<<<(unsigned)((n_pages + CORPUS_THREADS - 1) / CORPUS_THREADS), CORPUS_THREADS, smem, st>>>((uint8_t *)d_text, n_pages, P);
		g_launches++;
		CUDA_TRY(cudaGetLastError());
	}
	return AGB_OK;
}

extern "C" int agb_corpus_fill_host(const agb_corpus_spec *s, void *h_text)
{
	uint16_t woff[257]; int rc = corpus_check(s, woff); if (rc) return rc;
	agb_corpus_spec t = *s; t.needle[63] = 0;
	int nl = (int)strlen(t.needle);
	for (uint64_t pg = 0; pg < s->n_bytes / AGB_PAGE; pg++)
		agb_corpus_page((uint8_t *)h_text + pg * AGB_PAGE, t.seed, t.first_page + pg, h_vocab, woff,
		                t.paragraphs, t.needle_every, t.needle, nl, t.needle_maxedits);
	return AGB_OK;
}

/* ================================================================================================
 * host side of the scan
 * ============================================================================================== */
#define H2D_SLICE   (64ull << 20)      /* bytes per H2D slice of agb_scan_host; a multiple of the 16 KiB stage */
#define STAGE_BUFS  3

/* ================================================================================================
 * ordinals: j of the reference's loops (bitap.c:178, asearch.c:120), what -n prints minus one.
 *
 * j is incremented at every record close, the virtual '\n' included, so the ordinal of a record is the number of
 * delimiter ends at or before the delimiter that closes it -- a property of the text alone.  k_delim_count counts
 * the delimiter ends of every 512-byte block (16-bit) and every 32 KiB tile (one more HBM-bound pass, only when
 * ordinals are asked for); the tile counts are scanned; k_ordinals gives every record tile prefix + the blocks
 * of its tile before its own + the delimiter ends of its own block up to its close.  Same delimiter rule as
 * delim_ends_at() / agb_fill_ordinals(): every occurrence of a border-free delimiter, greedy pairing from the
 * start of the run for c^L ("$$"), the virtual '\n' and the delimiter appended at EOF included.
 * ============================================================================================== */
#define ORD_THREADS 256
#define ORD_TILE    32768
#define ORD_PER     (ORD_TILE / ORD_THREADS)       /* 128 bytes per thread */
#define ORD_BLOCK   512

struct OrdParams {
	const uint8_t *text; uint64_t n;
	uint16_t *blocks; uint32_t *tiles; const uint64_t *tile_off;
	agb_record *records; const unsigned long long *totals; uint64_t capacity;
	uint8_t delim[AGB_MAXDELIM + 2]; int L, kind;
	long long j0;                /* 0, or -1 when the text starts with the user's delimiter (bitap.c:151-156) */
};

/* delimiter ends in [from, to) (file offsets; to <= n + L), sequentially; run: the length of the run of delim[0]
 * that ends at from - 1 (kind 1) */
__device__ __forceinline__ uint32_t ord_count_seq(Reader &R, const OrdParams &P, int64_t from, int64_t to)
{
	uint32_t cnt = 0;
	if (P.L == 1) { for (int64_t q = from; q < to; q++) cnt += R.get(q) == P.delim[0]; return cnt; }
	if (P.kind == 0) {
		for (int64_t q = from; q < to; q++) {
			bool m = true;
			for (int u = 0; u < P.L && m; u++) m = R.get(q - u) == P.delim[P.L - 1 - u];
			cnt += m ? 1u : 0u;
		}
		return cnt;
	}
	const int c = P.delim[0];
	int64_t run = 0;
	for (int64_t q = from - 1; q >= -1 && R.get(q) == c; q--) run++;       /* (-1 is the virtual '\n') */
	for (int64_t q = from; q < to; q++) {
		run = R.get(q) == c ? run + 1 : 0;
		cnt += (run > 0 && run % P.L == 0) ? 1u : 0u;
	}
	return cnt;
}

__global__ void __launch_bounds__(ORD_THREADS) k_delim_count(const OrdParams P)
{
	__shared__ uint32_t s_warp[ORD_THREADS / 32];
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int64_t n = (int64_t)P.n, limit = n + P.L, tile0 = (int64_t)blockIdx.x * ORD_TILE;
	uint32_t cnt = 0;                                           /* this thread's share of the tile */
	if (P.L == 1 && tile0 + ORD_TILE <= n) {
		/* a warp takes a 512-byte block per iteration, 16 bytes per lane (coalesced): exact per-byte equality by
		 * SWAR (0x80 where the byte equals the delimiter), one warp reduction per block */
		const uint32_t d4 = P.delim[0] * 0x01010101u;
#pragma unroll
		for (int it = 0; it < ORD_TILE / ORD_BLOCK / (ORD_THREADS / 32); it++) {
			const uint32_t blk = wid * (ORD_TILE / ORD_BLOCK / (ORD_THREADS / 32)) + it;
			const uint4 x = __ldg(reinterpret_cast<const uint4 *>(P.text + tile0 + (int64_t)blk * ORD_BLOCK) + lane);
			const uint32_t xs[4] = { x.x, x.y, x.z, x.w };
			uint32_t c = 0;
#pragma unroll
			for (int w = 0; w < 4; w++) {
				const uint32_t t = xs[w] ^ d4;
				c += __popc(~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu));
			}
			const uint32_t b = __reduce_add_sync(0xffffffffu, c);
			if (lane == 0) P.blocks[(uint64_t)blockIdx.x * (ORD_TILE / ORD_BLOCK) + blk] = (uint16_t)b;
			cnt += c;
		}
	} else {
		/* other delimiters and the last tile: every thread walks its 128 bytes; 4 threads = one block */
		const int64_t s0 = tile0 + (int64_t)tid * ORD_PER, s1 = s0 + ORD_PER < limit ? s0 + ORD_PER : limit;
		if (s0 < limit) {
			Reader R; R.init(P.text, P.n, P.delim, P.L);
			cnt = ord_count_seq(R, P, s0, s1);
		}
		uint32_t b = cnt;
		b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
		if ((tid & 3) == 0) P.blocks[(uint64_t)blockIdx.x * (ORD_TILE / ORD_BLOCK) + (tid >> 2)] = (uint16_t)b;
	}
	const uint32_t w = __reduce_add_sync(0xffffffffu, cnt);
	if (lane == 0) s_warp[wid] = w;
	__syncthreads();
	if (tid == 0) { uint32_t t = 0; for (int i = 0; i < ORD_THREADS / 32; i++) t += s_warp[i]; P.tiles[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(256) k_ordinals(const OrdParams P)
{
	unsigned long long nrec = P.totals[0];
	if (nrec > P.capacity) nrec = P.capacity;
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nrec) return;
	const int64_t q = P.records[i].end + P.L - 1;                          /* the last byte of the closing delimiter */
	const uint64_t tile = (uint64_t)q / ORD_TILE, blk = (uint64_t)q / ORD_BLOCK;
	unsigned long long j = P.tile_off[tile];
	for (uint64_t b = tile * (ORD_TILE / ORD_BLOCK); b < blk; b++) j += P.blocks[b];
	Reader R; R.init(P.text, P.n, P.delim, P.L);
	j += ord_count_seq(R, P, (int64_t)(blk * ORD_BLOCK), q + 1);
	/* the virtual '\n' closes a record of its own when it completes a delimiter: only a 1-byte '\n' can */
	const long long virt = (P.L == 1 && P.delim[0] == '\n') ? 1 : 0;
	P.records[i].ordinal = (long long)j + virt + P.j0;
}

struct Workspace {               /* grow-only device scratch, one per device */
	uint32_t *bitmap = nullptr; size_t bitmap_bytes = 0;
	uint32_t *tile_counts = nullptr; uint64_t *tile_offsets = nullptr; size_t tiles = 0;
	uint64_t *cand = nullptr; uint32_t *cand_counts = nullptr; uint64_t *cand_offsets = nullptr; agb_record *cand_first = nullptr; size_t cand_cap = 0;
	uint32_t *scan_sums = nullptr; uint64_t *scan_offs = nullptr; size_t scan_cap = 0;
	uint16_t *ord_blocks = nullptr; size_t ord_blocks_cap = 0;     /* delimiter ends per 512-byte block (AGB_WANT_ORDINALS) */
	int ord_virt = 0;                                              /* 1: the virtual '\n' closes a record of its own (1-byte '\n' delimiter) */
	unsigned long long *totals = nullptr;          /* 16 counters */
	unsigned long long *h_totals = nullptr;        /* pinned */
	agb_desc *d_desc = nullptr; agb_desc h_desc_copy; bool desc_valid = false;
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
	int sm_count = 0;
	/* agb_scan_host: device copy of the text, record buffer, copy stream, pinned staging for pageable sources */
	uint8_t *h2d_text = nullptr; size_t h2d_cap = 0;
	agb_record *h2d_rec = nullptr; size_t h2d_rec_cap = 0;
	cudaStream_t s_copy = nullptr, s_comp = nullptr;
	cudaEvent_t ev_copy[STAGE_BUFS] = {nullptr, nullptr, nullptr};
	uint8_t *stage[STAGE_BUFS] = {nullptr, nullptr, nullptr};
};
static Workspace g_ws[64];
static std::mutex g_ws_mu;

static int ws_prepare(Workspace &W, uint64_t n)
{
	if (!W.totals) {
		CUDA_TRY(cudaMalloc(&W.totals, 16 * sizeof(unsigned long long)));
		CUDA_TRY(cudaMallocHost(&W.h_totals, 16 * sizeof(unsigned long long)));
		CUDA_TRY(cudaMalloc(&W.d_desc, sizeof(agb_desc)));
		CUDA_TRY(cudaEventCreate(&W.e0)); CUDA_TRY(cudaEventCreate(&W.e1)); CUDA_TRY(cudaEventCreate(&W.e2));
		int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
		CUDA_TRY(cudaDeviceGetAttribute(&W.sm_count, cudaDevAttrMultiProcessorCount, dev));
	}
	uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n + DENSE_TILE - 1) / DENSE_TILE + 1;
	size_t bb = (size_t)(n_words + FRONT_WORDS_PER_STAGE) * 4;
	if (bb > W.bitmap_bytes) {
		if (W.bitmap) cudaFree(W.bitmap);
		W.bitmap = nullptr; W.bitmap_bytes = 0;
		CUDA_TRY(cudaMalloc(&W.bitmap, bb)); W.bitmap_bytes = bb;
	}
	if (tiles + 1 > W.tiles) {
		if (W.tile_counts) cudaFree(W.tile_counts);
		if (W.tile_offsets) cudaFree(W.tile_offsets);
		W.tile_counts = nullptr; W.tile_offsets = nullptr; W.tiles = 0;
		CUDA_TRY(cudaMalloc(&W.tile_counts, (tiles + 1) * sizeof(uint32_t)));
		CUDA_TRY(cudaMalloc(&W.tile_offsets, (tiles + 1) * sizeof(uint64_t)));
		W.tiles = tiles + 1;
	}
	return AGB_OK;
}

/* the candidate list of the list form is sized by what stage 1.5 actually left (known on the host by then) */
static int ws_cand_reserve(Workspace &W, size_t want_cand)
{
	want_cand = std::max<size_t>(want_cand, (size_t)1 << 20);
	if (want_cand > W.cand_cap) {
		want_cand += want_cand / 4;
		if (W.cand) cudaFree(W.cand);
		if (W.cand_counts) cudaFree(W.cand_counts);
		if (W.cand_offsets) cudaFree(W.cand_offsets);
		if (W.cand_first) cudaFree(W.cand_first);
		if (W.scan_sums) cudaFree(W.scan_sums);
		if (W.scan_offs) cudaFree(W.scan_offs);
		W.cand = nullptr; W.cand_counts = nullptr; W.cand_offsets = nullptr; W.cand_first = nullptr; W.cand_cap = 0;
		W.scan_sums = nullptr; W.scan_offs = nullptr; W.scan_cap = 0;
		CUDA_TRY(cudaMalloc(&W.cand_first, want_cand * sizeof(agb_record)));
		W.scan_cap = want_cand / SCAN_BLOCK + 2;
		CUDA_TRY(cudaMalloc(&W.scan_sums, W.scan_cap * sizeof(uint32_t)));
		CUDA_TRY(cudaMalloc(&W.scan_offs, W.scan_cap * sizeof(uint64_t)));
		CUDA_TRY(cudaMalloc(&W.cand, want_cand * sizeof(uint64_t)));
		CUDA_TRY(cudaMalloc(&W.cand_counts, want_cand * sizeof(uint32_t)));
		CUDA_TRY(cudaMalloc(&W.cand_offsets, want_cand * sizeof(uint64_t)));
		W.cand_cap = want_cand;
	}
	return AGB_OK;
}

static int ws_upload_desc(Workspace &W, const agb_desc &d, cudaStream_t st)
{
	if (!W.desc_valid || memcmp(&W.h_desc_copy, &d, sizeof d) != 0) {
		CUDA_TRY(cudaMemcpyAsync(W.d_desc, &d, sizeof d, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaStreamSynchronize(st));     /* &d may be on the caller's stack */
		W.h_desc_copy = d; W.desc_valid = true;
	}
	return AGB_OK;
}

#define FRONT_SMEM (FRONT_NST * FRONT_SLOT_BYTES)
template <int NA, bool MASKED, bool FOLD, bool POLY>
static void launch_front_one(const FrontParams &P, unsigned grid, cudaStream_t st)
{
	static bool configured[64] = {false};
	int dev = 0; cudaGetDevice(&dev);
	if (!configured[dev & 63]) {
		cudaFuncSetAttribute(k_front<NA, MASKED, FOLD, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, FRONT_SMEM);
		configured[dev & 63] = true;
	}
	k_front<NA, MASKED, FOLD, POLY><<<grid, FRONT_THREADS, FRONT_SMEM, st>>>(P);
}
template <int NA, bool POLY>
static void launch_front_na(const FrontParams &P, bool masked, bool fold, unsigned grid, cudaStream_t st)
{
	if (masked) { if (fold) launch_front_one<NA, true, true, POLY>(P, grid, st); else launch_front_one<NA, true, false, POLY>(P, grid, st); }
	else        { if (fold) launch_front_one<NA, false, true, POLY>(P, grid, st); else launch_front_one<NA, false, false, POLY>(P, grid, st); }
}

/* coefficients of prod_i (x - a_i) mod 2^32 and the false-positive guard of the polynomial form:
 * a zero product without a zero factor needs sum_i v2(w - a_i) >= bits; with t = the largest v2(a_i - a_j)
 * at most one factor can exceed t, so w must agree with an anchor in its low bits - (na-1)*t bits.  We ask
 * for at least 20 agreeing bits (a 2.5-byte accidental match) or use the compare form instead. */
static bool poly_setup(const uint32_t *a, int na, int bits, uint32_t *coef)
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

static bool front_usable(const agb_desc &d) { return d.plan == AGB_PLAN_ANCHORS && d.n_anchors >= 1 && d.n_anchors <= 9; }

/* stage 1 over bitmap words [word_begin, word_end) of a text of n bytes; word_begin must be a multiple of 32
 * (a stage is 32 words).  slack16: 16 more bytes after the last chunk are readable (true for our own buffers). */
static int front_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n,
                        uint64_t word_begin, uint64_t word_end, bool slack16, cudaStream_t st)
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
	F.one = 1; F.scale = 1;
	for (int i = d.anchor_len; i < 4; i++) F.scale <<= 8;
	const bool poly = poly_setup(F.anchor, na, 8 * d.anchor_len, F.coef);
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

template <typename T, bool COSTS>
static int launch_records_t(int nrows, const RecParams &P, unsigned grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: k_records<T, 1, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 2: k_records<T, 2, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 3: k_records<T, 3, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 4: k_records<T, 4, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 5: k_records<T, 5, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 6: k_records<T, 6, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 7: k_records<T, 7, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 8: k_records<T, 8, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 9: k_records<T, 9, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	default: return -1;
	}
	g_launches++;
	return 0;
}

static int launch_records(const agb_desc &d, const RecParams &P, unsigned grid, cudaStream_t st)
{
	bool costs = d.engine == AGB_ENGINE_ASEARCH1;
	bool narrow = d.M <= 31;        /* the reference's own word width; wider patterns use 64-bit rows */
	if (costs) return narrow ? launch_records_t<uint32_t, true>(d.nrows, P, grid, st) : launch_records_t<uint64_t, true>(d.nrows, P, grid, st);
	return narrow ? launch_records_t<uint32_t, false>(d.nrows, P, grid, st) : launch_records_t<uint64_t, false>(d.nrows, P, grid, st);
}

#define DENSE_SMEM (DENSE_TILE + DENSE_TAIL)
template <typename T, int NR, bool COSTS>
static void launch_dense_one(const RecParams &P, unsigned grid, cudaStream_t st)
{
	static bool configured[64] = {false};
	int dev = 0; cudaGetDevice(&dev);
	if (!configured[dev & 63]) {
		cudaFuncSetAttribute(k_records_dense<T, NR, COSTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM);
		configured[dev & 63] = true;
	}
	k_records_dense<T, NR, COSTS><<<grid, DENSE_THREADS, DENSE_SMEM, st>>>(P);
}
template <typename T, bool COSTS>
static int launch_dense_t(int nrows, const RecParams &P, unsigned grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: launch_dense_one<T, 1, COSTS>(P, grid, st); break;
	case 2: launch_dense_one<T, 2, COSTS>(P, grid, st); break;
	case 3: launch_dense_one<T, 3, COSTS>(P, grid, st); break;
	case 4: launch_dense_one<T, 4, COSTS>(P, grid, st); break;
	case 5: launch_dense_one<T, 5, COSTS>(P, grid, st); break;
	case 6: launch_dense_one<T, 6, COSTS>(P, grid, st); break;
	case 7: launch_dense_one<T, 7, COSTS>(P, grid, st); break;
	case 8: launch_dense_one<T, 8, COSTS>(P, grid, st); break;
	case 9: launch_dense_one<T, 9, COSTS>(P, grid, st); break;
	default: return -1;
	}
	g_launches++;
	return 0;
}
static int launch_dense(const agb_desc &d, const RecParams &P, unsigned grid, cudaStream_t st)
{
	const bool costs = d.engine == AGB_ENGINE_ASEARCH1, narrow = d.M <= 31;
	if (costs) return narrow ? launch_dense_t<uint32_t, true>(d.nrows, P, grid, st) : launch_dense_t<uint64_t, true>(d.nrows, P, grid, st);
	return narrow ? launch_dense_t<uint32_t, false>(d.nrows, P, grid, st) : launch_dense_t<uint64_t, false>(d.nrows, P, grid, st);
}

template <typename T, int NR, bool COSTS>
static void launch_slices_one(const RecParams &P, unsigned grid, cudaStream_t st)
{
	static bool configured[64] = {false};
	int dev = 0; cudaGetDevice(&dev);
	if (!configured[dev & 63]) {
		cudaFuncSetAttribute(k_records_slices<T, NR, COSTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM);
		configured[dev & 63] = true;
	}
	k_records_slices<T, NR, COSTS><<<grid, SL_THREADS, SL_SMEM, st>>>(P);
}
template <typename T, bool COSTS>
static int launch_slices_t(int nrows, const RecParams &P, unsigned grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: launch_slices_one<T, 1, COSTS>(P, grid, st); break;
	case 2: launch_slices_one<T, 2, COSTS>(P, grid, st); break;
	case 3: launch_slices_one<T, 3, COSTS>(P, grid, st); break;
	case 4: launch_slices_one<T, 4, COSTS>(P, grid, st); break;
	case 5: launch_slices_one<T, 5, COSTS>(P, grid, st); break;
	case 6: launch_slices_one<T, 6, COSTS>(P, grid, st); break;
	case 7: launch_slices_one<T, 7, COSTS>(P, grid, st); break;
	case 8: launch_slices_one<T, 8, COSTS>(P, grid, st); break;
	case 9: launch_slices_one<T, 9, COSTS>(P, grid, st); break;
	default: return -1;
	}
	g_launches++;
	return 0;
}
static int launch_slices(const agb_desc &d, const RecParams &P, unsigned grid, cudaStream_t st)
{
	const bool costs = d.engine == AGB_ENGINE_ASEARCH1, narrow = d.M <= 31;
	if (costs) return narrow ? launch_slices_t<uint32_t, true>(d.nrows, P, grid, st) : launch_slices_t<uint64_t, true>(d.nrows, P, grid, st);
	return narrow ? launch_slices_t<uint32_t, false>(d.nrows, P, grid, st) : launch_slices_t<uint64_t, false>(d.nrows, P, grid, st);
}
/* the slices form needs a bounded memory: no position that holds for ever ('#': wildmask; -p: Init1 = ~0) and a
 * delimiter whose occurrences do not depend on where a run of it started */
static bool slices_usable(const agb_desc &d)
{
	return d.wildmask == 0 && d.init1 != ~0ull && (d.L == 1 || d.delim_kind == 0) && d.M + d.nrows + 2 <= SL_APRON;
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
static int refine_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, cudaStream_t st)
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

template <typename T, bool COSTS>
static int launch_records_list_t(int nrows, const RecParams &P, unsigned grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: k_records_list<T, 1, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 2: k_records_list<T, 2, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 3: k_records_list<T, 3, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 4: k_records_list<T, 4, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 5: k_records_list<T, 5, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 6: k_records_list<T, 6, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 7: k_records_list<T, 7, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 8: k_records_list<T, 8, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	case 9: k_records_list<T, 9, COSTS><<<grid, REC_THREADS, 0, st>>>(P); break;
	default: return -1;
	}
	g_launches++;
	return 0;
}

static int launch_records_list(const agb_desc &d, const RecParams &P, unsigned grid, cudaStream_t st)
{
	const bool costs = d.engine == AGB_ENGINE_ASEARCH1, narrow = d.M <= 31;
	if (costs) return narrow ? launch_records_list_t<uint32_t, true>(d.nrows, P, grid, st) : launch_records_list_t<uint64_t, true>(d.nrows, P, grid, st);
	return narrow ? launch_records_list_t<uint32_t, false>(d.nrows, P, grid, st) : launch_records_list_t<uint64_t, false>(d.nrows, P, grid, st);
}

/* stage 2 over the whole text.  After stage 1.5 the survivors are few: they are compacted into an ordered list
 * and each gets its own thread (count launch -> scan -> emit launch).  Otherwise (or if the list would not fit)
 * the dense form walks the bitmap, one thread per word. */
static int records_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, bool use_front, int want,
                          int want_level, agb_record *d_records, uint64_t capacity, cudaStream_t st)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n_words + REC_THREADS - 1) / REC_THREADS;
	RecParams P; memset(&P, 0, sizeof P);
	P.text = (const uint8_t *)d_text; P.bitmap = use_front ? W.bitmap : nullptr;
	P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.records = d_records; P.capacity = capacity;
	P.totals = W.totals; P.emit = 0; P.levels = (want & AGB_WANT_LEVELS) ? 1 : 0; P.want_level = want_level;
	if (!tiles) return AGB_OK;
	const bool want_list = (want & AGB_WANT_RECORDS) && capacity;
	if (use_front) {
		const uint64_t blocks = (n_words + COMPACT_THREADS * COMPACT_WPT - 1) / (COMPACT_THREADS * COMPACT_WPT);
		k_compact_count<<<(unsigned)blocks, COMPACT_THREADS, 0, st>>>(W.bitmap, n_words, W.tile_counts, W.totals); g_launches++;
		k_scan_tiles<<<1, 1024, 0, st>>>(W.tile_counts, W.tile_offsets, blocks, W.totals + 12); g_launches++;
		CUDA_TRY(cudaMemcpyAsync(W.h_totals + 12, W.totals + 12, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		const unsigned long long ncand = W.h_totals[12];
		/* list form while the survivors are sparse (20 B of scratch each; measured cross-over against the dense tile
		 * kernel at about 5 % of the chunks: 'the' flags 11 % and runs 13.5 ms per 4 GiB as a list, 'government' 1.1 % and 1.7 ms) */
		const bool sparse = ncand <= n_chunks / 20 + 1024;
		if (!sparse) {
			CUDA_TRY(cudaMemsetAsync(W.totals + 1, 0, sizeof(unsigned long long), st));
			use_front = false; P.bitmap = nullptr;
		} else if (ws_cand_reserve(W, (size_t)ncand) == AGB_OK) {
			if (ncand == 0) return AGB_OK;
			k_compact_write<<<(unsigned)blocks, COMPACT_THREADS, 0, st>>>(W.bitmap, n_words, W.tile_offsets, W.cand, W.cand_cap); g_launches++;
			P.cand = W.cand; P.cand_cap = W.cand_cap; P.tile_counts = W.cand_counts; P.tile_offsets = W.cand_offsets;
			P.cand_first = want_list ? W.cand_first : nullptr;
			const unsigned grid = (unsigned)((ncand + REC_THREADS - 1) / REC_THREADS);
			if (launch_records_list(d, P, grid, st)) return AGB_ERR_ARG;
			CUDA_TRY(cudaGetLastError());
			if (want_list) {
				if (ncand <= 4 * SCAN_BLOCK) { k_scan_tiles<<<1, 1024, 0, st>>>(W.cand_counts, W.cand_offsets, ncand, nullptr); g_launches++; }
				else {
					const unsigned nb = (unsigned)((ncand + SCAN_BLOCK - 1) / SCAN_BLOCK);
					k_scan_partial<<<nb, 1024, 0, st>>>(W.cand_counts, ncand, W.scan_sums);
					k_scan_tiles<<<1, 1024, 0, st>>>(W.scan_sums, W.scan_offs, nb, nullptr);
					k_scan_apply<<<nb, 1024, 0, st>>>(W.cand_counts, ncand, W.scan_offs, W.cand_offsets);
					g_launches += 3;
				}
				P.emit = 1;
				if (launch_records_list(d, P, grid, st)) return AGB_ERR_ARG;
				CUDA_TRY(cudaGetLastError());
			}
			return AGB_OK;
		}
		else CUDA_TRY(cudaMemsetAsync(W.totals + 1, 0, sizeof(unsigned long long), st));   /* no scratch for a list: bitmap form below recounts totals[1] */
	}
	P.tile_counts = W.tile_counts; P.tile_offsets = W.tile_offsets;
	if (!use_front) {
		/* no bitmap at all: the dense tile kernel, one CTA per 32 KiB (tile_counts has n/64KiB... entries: 2 per REC tile) */
		const bool slices = slices_usable(d);
		const uint64_t dtiles = slices ? (n + SL_TILE - 1) / SL_TILE : (n + DENSE_TILE - 1) / DENSE_TILE;
		P.warm = (d.M + d.nrows + 2 + 3) & ~3;
		if (slices ? launch_slices(d, P, (unsigned)dtiles, st) : launch_dense(d, P, (unsigned)dtiles, st)) return AGB_ERR_ARG;
		CUDA_TRY(cudaGetLastError());
		if (want_list) {
			k_scan_tiles<<<1, 1024, 0, st>>>(W.tile_counts, W.tile_offsets, dtiles, nullptr); g_launches++;
			P.emit = 1;
			if (slices ? launch_slices(d, P, (unsigned)dtiles, st) : launch_dense(d, P, (unsigned)dtiles, st)) return AGB_ERR_ARG;
			CUDA_TRY(cudaGetLastError());
		}
		return AGB_OK;
	}
	if (launch_records(d, P, (unsigned)tiles, st)) return AGB_ERR_ARG;
	CUDA_TRY(cudaGetLastError());
	if (want_list) {
		k_scan_tiles<<<1, 1024, 0, st>>>(W.tile_counts, W.tile_offsets, tiles, nullptr); g_launches++;
		P.emit = 1;
		if (launch_records(d, P, (unsigned)tiles, st)) return AGB_ERR_ARG;
		CUDA_TRY(cudaGetLastError());
	}
	return AGB_OK;
}

/* after stage 1: is the bitmap so full that thinning it (stage 1.5) and walking a candidate list cannot pay?  Then the
 * record stage walks every byte anyway (slices / dense tile form) and stage 1.5 is skipped.  Estimated from every
 * 61st bitmap word; same 5 % threshold as the list/dense switch in records_launch(). */
static int front_is_dense(Workspace &W, uint64_t n, cudaStream_t st, bool *dense)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32;
	const uint32_t stride = n_words > (1u << 16) ? 61u : 1u;
	CUDA_TRY(cudaMemsetAsync(W.totals + 14, 0, sizeof(unsigned long long), st));
	const uint64_t samples = (n_words + stride - 1) / stride;
	const unsigned grid = (unsigned)std::min<uint64_t>((samples + 255) / 256, (uint64_t)W.sm_count * 8);
	k_bitmap_sample<<<grid ? grid : 1, 256, 0, st>>>(W.bitmap, n_words, stride, W.totals + 14); g_launches++;
	CUDA_TRY(cudaMemcpyAsync(W.h_totals + 14, W.totals + 14, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	*dense = W.h_totals[14] * stride > n_chunks / 20 + 1024;
	return AGB_OK;
}

/* AGB_WANT_ORDINALS: fill agb_record.ordinal of the list just written and leave the number of record closes of the
 * whole text (j after the last record, the basis of the next shard's ordinals) in totals[13].  Runs after the
 * record stage, whose tile scratch it reuses. */
static int ordinals_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, agb_record *d_records,
                           uint64_t capacity, cudaStream_t st)
{
	uint8_t *h_head = reinterpret_cast<uint8_t *>(W.h_totals + 14);      /* pinned scratch: the first bytes of the text */
	if (n >= (uint64_t)d.L && d.user_delim) {
		CUDA_TRY(cudaMemcpyAsync(h_head, d_text, (size_t)d.L, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	const uint64_t limit = n + (uint64_t)d.L, tiles = (limit + ORD_TILE - 1) / ORD_TILE;
	if (tiles + 1 > W.tiles) return AGB_ERR_NOMEM;                       /* (ws_prepare sized them for n + one tile) */
	const size_t nb = (size_t)tiles * (ORD_TILE / ORD_BLOCK);
	if (nb > W.ord_blocks_cap) {
		if (W.ord_blocks) cudaFree(W.ord_blocks);
		W.ord_blocks = nullptr; W.ord_blocks_cap = 0;
		CUDA_TRY(cudaMalloc(&W.ord_blocks, nb * sizeof(uint16_t))); W.ord_blocks_cap = nb;
	}
	OrdParams P; memset(&P, 0, sizeof P);
	P.text = (const uint8_t *)d_text; P.n = n; P.blocks = W.ord_blocks; P.tiles = W.tile_counts; P.tile_off = W.tile_offsets;
	P.records = d_records; P.totals = W.totals; P.capacity = capacity;
	memcpy(P.delim, d.delim, AGB_MAXDELIM + 2 < sizeof d.delim ? AGB_MAXDELIM + 2 : sizeof d.delim);
	P.L = d.L; P.kind = d.delim_kind;
	W.ord_virt = (d.L == 1 && d.delim[0] == '\n') ? 1 : 0;
	/* bitap.c:151-156: j starts at -1 when the text begins with the user's delimiter (asearch0() has no such correction) */
	P.j0 = (d.user_delim && d.engine != AGB_ENGINE_ASEARCH0 && n >= (uint64_t)d.L && memcmp(h_head, d.delim, (size_t)d.L) == 0) ? -1 : 0;
	k_delim_count<<<(unsigned)tiles, ORD_THREADS, 0, st>>>(P); g_launches++;
	k_scan_tiles<<<1, 1024, 0, st>>>(W.tile_counts, W.tile_offsets, tiles, W.totals + 13); g_launches++;
	if (d_records && capacity) {
		/* the list length is on the device (totals[0]); one thread per possible entry, bounded by the capacity */
		CUDA_TRY(cudaMemcpyAsync(W.h_totals, W.totals, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		const uint64_t nrec = std::min<uint64_t>(W.h_totals[0], capacity);
		if (nrec) { k_ordinals<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(P); g_launches++; }
	}
	CUDA_TRY(cudaGetLastError());
	return AGB_OK;
}

static int fetch_result(Workspace &W, int want, uint64_t capacity, cudaStream_t st, agb_result *res)
{
	CUDA_TRY(cudaMemcpyAsync(W.h_totals, W.totals, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	res->n_matched = W.h_totals[0];
	res->n_flagged = W.h_totals[1];
	for (int i = 0; i <= AGB_MAXERR; i++) res->level_hist[i] = W.h_totals[2 + i];
	res->n_records = (want & AGB_WANT_RECORDS) ? std::min<uint64_t>(res->n_matched, capacity) : 0;
	res->n_closes = (want & AGB_WANT_ORDINALS) ? W.h_totals[13] + (uint64_t)W.ord_virt : 0;
	return AGB_OK;
}

static int scan_device_impl(const agb_desc &d, const void *d_text, uint64_t n, int want, int want_level,
                            agb_record *d_records, uint64_t capacity, cudaStream_t st, agb_result *res)
{
	if (!res) return AGB_ERR_ARG;
	memset(res, 0, sizeof *res);
	if (((uintptr_t)d_text & 15) != 0) { snprintf(g_err, sizeof g_err, "text pointer must be 16-byte aligned"); return AGB_ERR_ARG; }
	if ((want & AGB_WANT_RECORDS) && capacity && !d_records) return AGB_ERR_ARG;
	int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) return AGB_ERR_ARG;
	std::lock_guard<std::mutex> lk(g_ws_mu);
	Workspace &W = g_ws[dev];
	int rc = ws_prepare(W, n); if (rc) return rc;
	rc = ws_upload_desc(W, d, st); if (rc) return rc;
	CUDA_TRY(cudaMemsetAsync(W.totals, 0, 16 * sizeof(unsigned long long), st));
	CUDA_TRY(cudaEventRecord(W.e0, st));
	bool use_front = front_usable(d) && n > 0;
	if (use_front) { rc = front_launch(d, W, d_text, n, 0, ~0ull, false, st); if (rc) return rc; }
	CUDA_TRY(cudaEventRecord(W.e1, st));
	if (use_front) { bool dense = false; rc = front_is_dense(W, n, st, &dense); if (rc) return rc; if (dense) use_front = false; }
	if (use_front) { rc = refine_launch(d, W, d_text, n, st); if (rc) return rc; }
	rc = records_launch(d, W, d_text, n, use_front, want, want_level, d_records, capacity, st); if (rc) return rc;
	if (want & AGB_WANT_ORDINALS) { rc = ordinals_launch(d, W, d_text, n, (want & AGB_WANT_RECORDS) ? d_records : nullptr, capacity, st); if (rc) return rc; }
	CUDA_TRY(cudaEventRecord(W.e2, st));
	rc = fetch_result(W, want, capacity, st, res); if (rc) return rc;
	CUDA_TRY(cudaEventElapsedTime(&res->ms_front, W.e0, W.e1));
	CUDA_TRY(cudaEventElapsedTime(&res->ms_records, W.e1, W.e2));
	return AGB_OK;
}

extern "C" int agb_scan_device(const agb_pattern *p, const void *d_text, uint64_t n, int want,
                               agb_record *d_records, uint64_t capacity, void *stream, agb_result *res)
{
	if (!p) return AGB_ERR_ARG;
	return scan_device_impl(p->d, d_text, n, want, -1, d_records, capacity, (cudaStream_t)stream, res);
}

static void par_memcpy(uint8_t *dst, const uint8_t *src, size_t len)
{
	const int T = 4; const size_t part = ((len + T - 1) / T + 4095) & ~(size_t)4095;
	std::thread th[T]; int used = 0;
	for (int t = 0; t < T; t++) {
		size_t a = (size_t)t * part; if (a >= len) break;
		size_t l = std::min(part, len - a);
		th[used++] = std::thread([=] { memcpy(dst + a, src + a, l); });
	}
	for (int t = 0; t < used; t++) th[t].join();
}

/* Host text -> HBM -> scan: the replacement of the fill_buf()/read(2) loop (bitap.c:143,450-477).  The text is
 * moved in 64 MiB slices on a copy stream -- straight from the caller's memory when it is page-locked; through a
 * pinned ring filled by 4 host threads when it is pageable; read(2) straight into the pinned ring when the
 * source is a file descriptor -- while stage 1 runs on the slice that arrived before (its last chunk looks 4
 * bytes into the next one), so the scan hides behind PCIe; stages 1.5 and 2 run once over the whole bitmap. */
struct SliceSource {
	const uint8_t *mem;      /* host memory source, or NULL */
	bool pinned;             /* mem is page-locked: copy from it directly */
	int fd;                  /* file descriptor source when mem == NULL */
};

static int scan_stream_impl(const agb_desc &d, uint64_t n, const SliceSource &src, int want,
                            agb_record *records, uint64_t capacity, agb_result *res)
{
	memset(res, 0, sizeof *res);
	int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) return AGB_ERR_ARG;
	std::lock_guard<std::mutex> lk(g_ws_mu);
	Workspace &W = g_ws[dev];
	int rc = ws_prepare(W, n); if (rc) return rc;
	if (!W.s_copy) {
		CUDA_TRY(cudaStreamCreateWithFlags(&W.s_copy, cudaStreamNonBlocking));
		CUDA_TRY(cudaStreamCreateWithFlags(&W.s_comp, cudaStreamNonBlocking));
		for (int i = 0; i < STAGE_BUFS; i++) CUDA_TRY(cudaEventCreateWithFlags(&W.ev_copy[i], cudaEventDisableTiming));
	}
	const size_t need = (size_t)((n + 15) / 16 * 16 + 4096);
	if (need > W.h2d_cap) {
		if (W.h2d_text) cudaFree(W.h2d_text);
		W.h2d_text = nullptr; W.h2d_cap = 0;
		CUDA_TRY(cudaMalloc(&W.h2d_text, need)); W.h2d_cap = need;
	}
	if ((want & AGB_WANT_RECORDS) && capacity > W.h2d_rec_cap) {
		if (W.h2d_rec) cudaFree(W.h2d_rec);
		W.h2d_rec = nullptr; W.h2d_rec_cap = 0;
		CUDA_TRY(cudaMalloc(&W.h2d_rec, capacity * sizeof(agb_record))); W.h2d_rec_cap = capacity;
	}
	rc = ws_upload_desc(W, d, W.s_comp); if (rc) return rc;
	const bool direct = src.mem && src.pinned;
	if (!direct && n && !W.stage[0]) for (int i = 0; i < STAGE_BUFS; i++) CUDA_TRY(cudaMallocHost(&W.stage[i], H2D_SLICE));
	const bool use_front = front_usable(d) && n > 0;
	const uint64_t words_per_slice = H2D_SLICE / 512, n_slices = (n + H2D_SLICE - 1) / H2D_SLICE;
	CUDA_TRY(cudaMemsetAsync(W.totals, 0, 16 * sizeof(unsigned long long), W.s_comp));
	CUDA_TRY(cudaEventRecord(W.e0, W.s_comp));
	/* zero the slack after the text once (stage 1 reads whole 16-byte chunks, stage 2 whole groups) */
	CUDA_TRY(cudaMemsetAsync(W.h2d_text + (n & ~(uint64_t)15), 0, need - (n & ~(uint64_t)15), W.s_copy));
	for (uint64_t i = 0; i < n_slices; i++) {
		const uint64_t off = i * H2D_SLICE, len = std::min<uint64_t>(H2D_SLICE, n - off);
		const int sb = (int)(i % STAGE_BUFS);
		if (direct) {
			CUDA_TRY(cudaMemcpyAsync(W.h2d_text + off, src.mem + off, len, cudaMemcpyHostToDevice, W.s_copy));
		} else {
			if (i >= STAGE_BUFS) CUDA_TRY(cudaEventSynchronize(W.ev_copy[sb]));     /* that staging buffer has been consumed */
			if (src.mem) par_memcpy(W.stage[sb], src.mem + off, len);
			else {
				uint64_t got = 0;                                                   /* fill_buf(): read(2) until the slice is full */
				while (got < len) {
					ssize_t r = read(src.fd, W.stage[sb] + got, (size_t)(len - got));
					if (r <= 0) { snprintf(g_err, sizeof g_err, "read(2) returned %zd at offset %llu of %llu", r, (unsigned long long)(off + got), (unsigned long long)n); return AGB_ERR_ARG; }
					got += (uint64_t)r;
				}
			}
			CUDA_TRY(cudaMemcpyAsync(W.h2d_text + off, W.stage[sb], len, cudaMemcpyHostToDevice, W.s_copy));
		}
		CUDA_TRY(cudaEventRecord(W.ev_copy[sb], W.s_copy));
		/* stage 1 on the previous slice: its last chunk looks 4 bytes into this one, which is now on its way */
		if (use_front) {
			CUDA_TRY(cudaStreamWaitEvent(W.s_comp, W.ev_copy[sb], 0));
			if (i > 0) { rc = front_launch(d, W, W.h2d_text, n, (i - 1) * words_per_slice, i * words_per_slice, true, W.s_comp); if (rc) return rc; }
		}
	}
	if (n_slices) {
		CUDA_TRY(cudaStreamWaitEvent(W.s_comp, W.ev_copy[(n_slices - 1) % STAGE_BUFS], 0));
		if (use_front) { rc = front_launch(d, W, W.h2d_text, n, (n_slices - 1) * words_per_slice, ~0ull, true, W.s_comp); if (rc) return rc; }
	}
	CUDA_TRY(cudaEventRecord(W.e1, W.s_comp));
	bool use_bitmap = use_front;
	if (use_bitmap) { bool dense = false; rc = front_is_dense(W, n, W.s_comp, &dense); if (rc) return rc; if (dense) use_bitmap = false; }
	if (use_bitmap) { rc = refine_launch(d, W, W.h2d_text, n, W.s_comp); if (rc) return rc; }
	rc = records_launch(d, W, W.h2d_text, n, use_bitmap, want, -1, W.h2d_rec, capacity, W.s_comp); if (rc) return rc;
	if (want & AGB_WANT_ORDINALS) { rc = ordinals_launch(d, W, W.h2d_text, n, (want & AGB_WANT_RECORDS) ? W.h2d_rec : nullptr, capacity, W.s_comp); if (rc) return rc; }
	CUDA_TRY(cudaEventRecord(W.e2, W.s_comp));
	rc = fetch_result(W, want, capacity, W.s_comp, res); if (rc) return rc;
	if (res->n_records) {
		CUDA_TRY(cudaMemcpyAsync(records, W.h2d_rec, res->n_records * sizeof(agb_record), cudaMemcpyDeviceToHost, W.s_comp));
		CUDA_TRY(cudaStreamSynchronize(W.s_comp));
	}
	CUDA_TRY(cudaStreamSynchronize(W.s_copy));
	CUDA_TRY(cudaEventElapsedTime(&res->ms_front, W.e0, W.e1));
	CUDA_TRY(cudaEventElapsedTime(&res->ms_records, W.e1, W.e2));
	return AGB_OK;
}

extern "C" int agb_scan_host(const agb_pattern *p, const void *h_text, uint64_t n, int want,
                             agb_record *records, uint64_t capacity, agb_result *res)
{
	if (!p || !res || (!h_text && n)) return AGB_ERR_ARG;
	if ((want & AGB_WANT_RECORDS) && capacity && !records) return AGB_ERR_ARG;
	SliceSource src; src.mem = (const uint8_t *)h_text; src.fd = -1; src.pinned = false;
	if (n) {
		cudaPointerAttributes attr; memset(&attr, 0, sizeof attr);
		src.pinned = cudaPointerGetAttributes(&attr, h_text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
		cudaGetLastError();
	}
	return scan_stream_impl(p->d, n, src, want, records, capacity, res);
}

extern "C" int agb_scan_fd(const agb_pattern *p, int fd, int want, agb_record *records, uint64_t capacity, agb_result *res)
{
	if (!p || !res) return AGB_ERR_ARG;
	if ((want & AGB_WANT_RECORDS) && capacity && !records) return AGB_ERR_ARG;
	struct stat sb;
	if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode)) {
		/* regular file: the size is known, read(2) goes straight into the pinned ring, slice by slice */
		off_t cur = lseek(fd, 0, SEEK_CUR);
		uint64_t n = (cur >= 0 && sb.st_size > cur) ? (uint64_t)(sb.st_size - cur) : 0;
		SliceSource src; src.mem = nullptr; src.pinned = false; src.fd = fd;
		return scan_stream_impl(p->d, n, src, want, records, capacity, res);
	}
	/* pipes, ttys: fill_buf() semantics -- read until EOF into a growing buffer, then as host memory */
	size_t cap = 1 << 20, len = 0; uint8_t *buf = (uint8_t *)malloc(cap);
	if (!buf) return AGB_ERR_NOMEM;
	for (;;) {
		if (len == cap) { cap *= 2; uint8_t *nb = (uint8_t *)realloc(buf, cap); if (!nb) { free(buf); return AGB_ERR_NOMEM; } buf = nb; }
		ssize_t r = read(fd, buf + len, cap - len);
		if (r < 0) { free(buf); snprintf(g_err, sizeof g_err, "read failed"); return AGB_ERR_ARG; }
		if (r == 0) break;
		len += (size_t)r;
	}
	int rc = agb_scan_host(p, buf, len, want, records, capacity, res);
	free(buf);
	return rc;
}

/* agrep.c:3582-3728: when the exact pass finds nothing, -B looks for the smallest D in 1..min(M-1,8) with a
 * match, rescanning every file once per D.  The rows are nested (A_j contains A_{j-1}, asearch.c:98-114), so
 * ONE pass at the largest D yields every record's smallest level; level_hist tells the best D. */
extern "C" int agb_bestmatch_device(const char *pattern, const agb_options *opt, const void *d_text, uint64_t n,
                                    void *stream, int *best_k, agb_result *res, char *err, size_t errlen)
{
	if (!pattern || !opt || !best_k || !res) return AGB_ERR_ARG;
	agb_options o = *opt; agb_desc d; int m = (int)strlen(pattern);
	o.bestmatch = 1;
	*best_k = -1;
	/* D < M of the exact pattern (agrep.c:3594); M there counts the delimiter and separator too */
	o.k = 0;
	int rc = agbi_build(pattern, &o, &d, err, errlen); if (rc) return rc;
	int kmax = d.M - 1; if (kmax > AGB_MAXERR) kmax = AGB_MAXERR; if (kmax > m - 1) kmax = m - 1;
	/* staged doubling keeps the anchor filter selective: k = 0, then 2, 4, 8 */
	int stages[5] = { 0, 2, 4, 8, 8 }, prev = -1;
	for (int si = 0; si < 4; si++) {
		int k = stages[si] < kmax ? stages[si] : kmax;
		if (k <= prev) break;
		o.k = k;
		rc = agbi_build(pattern, &o, &d, err, errlen); if (rc) return rc;
		rc = scan_device_impl(d, d_text, n, AGB_WANT_COUNT | AGB_WANT_LEVELS, -1, nullptr, 0, (cudaStream_t)stream, res);
		if (rc) return rc;
		for (int l = prev + 1; l <= k; l++) if (res->level_hist[l]) { *best_k = l; res->n_matched = res->level_hist[l]; return AGB_OK; }
		/* levels <= prev were already known to be empty */
		prev = k;
	}
	res->n_matched = 0;
	return AGB_OK;
}
