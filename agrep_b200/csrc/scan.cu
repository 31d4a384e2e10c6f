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
#define FRONT_NST     4                                   /* stages in flight per CTA (64 KiB), 3 CTAs per SM  */
#define FRONT_CTAS_PER_SM 3
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
			uint32_t r = w[t] * P.one + P.coef[NA - 1];
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
	for (uint64_t it = 0;; it++) {
		const uint64_t sg = P.stage_begin + blockIdx.x + it * gridDim.x;
		if (sg >= P.stage_end) break;
		const uint32_t slot = (uint32_t)(it % FRONT_NST);
		mbar_wait(&s_bar[slot], (uint32_t)((it / FRONT_NST) & 1));
		const uint8_t *st = s_ring + slot * FRONT_SLOT_BYTES;
#pragma unroll
		for (int c = 0; c < FRONT_CH; c++) {
			const uint32_t idx = c * FRONT_THREADS + tid;
			uint4 v = *reinterpret_cast<const uint4 *>(st + idx * 16);
			uint32_t x4 = *reinterpret_cast<const uint32_t *>(st + idx * 16 + 16);
			if (FOLD) { v.x |= P.fold; v.y |= P.fold; v.z |= P.fold; v.w |= P.fold; x4 |= P.fold; }
			uint32_t acc = 0xffffffffu;
			acc = windows_test<NA, MASKED, POLY>(v.x, v.y, P, acc);
			acc = windows_test<NA, MASKED, POLY>(v.y, v.z, P, acc);
			acc = windows_test<NA, MASKED, POLY>(v.z, v.w, P, acc);
			acc = windows_test<NA, MASKED, POLY>(v.w, x4, P, acc);
			const uint64_t chunk = sg * FRONT_STAGE_CHUNKS + idx;
			/* the last chunks are always passed on: a match may run into the delimiter appended at EOF (bitap.c:161-165),
			 * and their look-ahead bytes may not exist */
			const bool flag = (chunk < P.n_chunks) && (acc == 0 || chunk + 2 >= P.n_chunks);
			const uint32_t word = __ballot_sync(0xffffffffu, flag);
			if (lane == 0 && chunk < P.n_chunks) P.bitmap[chunk >> 5] = word;
		}
		__syncthreads();                       /* everyone is done reading this slot */
		if (tid == 0) issue(it + FRONT_NST);   /* refill it with the stage FRONT_NST iterations ahead */
	}
}

/* ================================================================================================
 * stage 2: records
 * ============================================================================================== */
#define REC_THREADS 128          /* one thread per bitmap word: a block covers 128*512 B = 64 KiB of text */

struct RecParams {
	const uint8_t  *text;
	const uint32_t *bitmap;      /* NULL: every chunk flagged */
	uint64_t n, n_chunks, n_words;
	const agb_desc *desc;        /* device copy */
	uint32_t *tile_counts;       /* per block */
	const uint64_t *tile_offsets;/* exclusive scan of tile_counts (emit pass) */
	agb_record *records; uint64_t capacity;
	unsigned long long *totals;  /* [0] matched, [1] flagged chunks, [2..10] level histogram, [11] emitted */
	int emit;                    /* 0: count pass, 1: emit pass */
	int levels;                  /* 1: best-match bookkeeping (smallest matching row) */
	int want_level;              /* levels: report records whose smallest level <= want_level (-1: all matching) */
};

template <typename T> struct DevConsts {
	T init1, noerr, endpos, dendpos;
	int L, k, and_mode, inverse, kind, ci, cs, cd;
};

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

template <typename T, int NR, bool COSTS>
__device__ __forceinline__ void rows_step(T (&S)[NR], T cm, const DevConsts<T> &C)
{
	if (!COSTS) {
		T prevB = S[0];
		T prevA = ((prevB >> 1) & cm) | (C.init1 & prevB);
#pragma unroll
		for (int r = 1; r < NR; r++) {
			T b = S[r];
			T a = ((b >> 1) & cm) | (C.init1 & b) | prevB | (((prevA | prevB) >> 1) & C.noerr);
			S[r - 1] = prevA; prevA = a; prevB = b;
		}
		S[NR - 1] = prevA;
	} else {
		T A[NR];
		A[0] = ((S[0] >> 1) & cm) | (C.init1 & S[0]);
#pragma unroll
		for (int r = 1; r < NR; r++) {
			T bi = (r - C.ci >= 0) ? S[r - C.ci] : (T)0, ad = (r - C.cd >= 0) ? A[r - C.cd] : (T)0, bs = (r - C.cs >= 0) ? S[r - C.cs] : (T)0;
			A[r] = ((S[r] >> 1) & cm) | bi | (((ad | bs) >> 1) & C.noerr) | (C.init1 & S[r]);
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

template <typename T, int NR, bool COSTS>
__global__ void __launch_bounds__(REC_THREADS)
k_records(const RecParams P)
{
	__shared__ T s_mask[256];
	__shared__ T s_reset[NR], s_start[NR];
	__shared__ uint8_t s_delim[2 * AGB_MAXDELIM + 2];
	__shared__ uint32_t s_scan[REC_THREADS];
	__shared__ unsigned long long s_hist[AGB_MAXERR + 1];
	const agb_desc *D = P.desc;
	for (int i = threadIdx.x; i < 256; i += REC_THREADS) s_mask[i] = (T)D->mask[i];
	if (threadIdx.x < NR) { s_reset[threadIdx.x] = (T)D->reset[threadIdx.x]; s_start[threadIdx.x] = (T)D->start[threadIdx.x]; }
	if (threadIdx.x < 2 * AGB_MAXDELIM + 2) s_delim[threadIdx.x] = D->delim[threadIdx.x];
	if (threadIdx.x <= AGB_MAXERR) s_hist[threadIdx.x] = 0;
	DevConsts<T> C;
	C.init1 = (T)D->init1; C.noerr = (T)D->noerr; C.endpos = (T)D->endpos; C.dendpos = (T)D->dendpos;
	C.L = D->L; C.k = D->k; C.and_mode = D->and_mode; C.inverse = D->inverse; C.kind = D->delim_kind;
	C.ci = D->cost_i; C.cs = D->cost_s; C.cd = D->cost_d;
	__syncthreads();

	const uint64_t gw = (uint64_t)blockIdx.x * REC_THREADS + threadIdx.x;     /* bitmap word of this thread */
	const int L = C.L;
	const int64_t n = (int64_t)P.n;
	uint32_t word = 0;
	if (gw < P.n_words) {
		word = P.bitmap ? P.bitmap[gw] : 0xffffffffu;
		uint64_t rem = P.n_chunks - gw * 32;
		if (rem < 32) word &= (1u << rem) - 1u;
	}
	Reader R; R.init(P.text, P.n, s_delim, L);

	uint32_t my_count = 0;
	uint64_t out_pos = 0;
	/* pass 0 counts; in emit mode pass 1 repeats the walk and writes at the scanned offsets */
	for (int pass = 0; pass < (P.emit ? 2 : 1); pass++) {
		uint32_t bits = word;
		int64_t done_until = INT64_MIN;    /* everything before this offset belongs to records this thread already closed */
		uint32_t cnt = 0;
		while (bits) {
			int b = __ffs(bits) - 1; bits &= bits - 1;
			const int64_t c = (int64_t)(gw * 32 + b), lo = c * 16, hi = lo + 15;
			int64_t s = -2;                /* record start to run from; -2: none */
			if (done_until > lo) {
				/* the record this thread closed last reaches into this chunk; what starts here starts at done_until */
				if (done_until - 1 <= hi) s = done_until; else continue;
			} else {
				/* (a) the record that contains byte lo: ours iff its re-fed byte s-1 lies after the previous flagged chunk */
				bool found = false;
				if (c == 0) { s = 0; found = true; }
				for (int64_t cc = c - 1; !found; cc--) {
					if (cc < 0) { s = 0; found = true; break; }
					uint32_t pw = P.bitmap ? P.bitmap[cc >> 5] : 0xffffffffu;
					if (pw >> (cc & 31) & 1u) break;                  /* an earlier flagged chunk meets that record: not ours */
					for (int64_t q = cc * 16 + 15; q >= cc * 16; q--)
						if (delim_ends_at(R, q, s_delim, L, C.kind)) { s = q + 1; found = true; break; }
				}
				if (!found) {
					/* (b) the first record whose re-fed byte lies inside this chunk */
					for (int64_t q = lo; q <= hi && q < n; q++)
						if (delim_ends_at(R, q, s_delim, L, C.kind)) { s = q + 1; break; }
				}
			}
			/* run records while their re-fed byte (s-1) is at or before the end of this chunk */
			while (s >= 0 && s - 1 <= hi && s <= n) {
				T S[NR];
				int64_t begin;
				if (s == 0) {
#pragma unroll
					for (int r = 0; r < NR; r++) S[r] = s_start[r];
					begin = D->start_closes ? -(int64_t)L : 0;
				} else {
#pragma unroll
					for (int r = 0; r < NR; r++) S[r] = s_reset[r];
					begin = s - L;
				}
				int64_t p = s, close_at = -1;
				const int64_t limit = n + L;
				for (; p < limit; p++) {
					T cm = s_mask[R.get(p)];
					rows_step<T, NR, COSTS>(S, cm, C);
					if (S[0] & C.dendpos) { close_at = p; break; }
				}
				if (close_at < 0) { s = -2; done_until = limit + 1; break; }   /* never closed: dropped, as the reference does */
				const int64_t end = close_at + 1 - L;
				bool counts = (begin + 1 < n) && (begin + 1 <= end);           /* bitap.c:213 + agrep.c:3811 */
				int level = C.k;
				bool cond;
				if (P.levels) {
					level = -1;
#pragma unroll
					for (int r = 0; r < NR; r++) if (level < 0 && match_cond<T>(S[r], C)) level = r;
					cond = level >= 0;
					if (cond && counts && pass == 0) atomicAdd(&s_hist[level], 1ull);
					if (cond && P.want_level >= 0 && level > P.want_level) cond = false;
				} else cond = match_cond<T>(S[NR - 1], C);
				if (cond && counts) {
					if (pass == 1) {
						uint64_t at = out_pos + cnt;
						if (at < P.capacity) {
							agb_record rec; rec.begin = begin; rec.end = end; rec.ordinal = 0; rec.level = level; rec.pad = 0;
							P.records[at] = rec;
						}
					}
					cnt++;
				}
				s = close_at + 1;
				done_until = s;
			}
		}
		if (pass == 0) {
			my_count = cnt;
			/* block scan of the per-thread counts */
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
				uint32_t fl = __popc(word);
				fl = __reduce_add_sync(0xffffffffu, fl);
				if ((threadIdx.x & 31) == 0 && fl) atomicAdd(&P.totals[1], (unsigned long long)fl);
				__syncthreads();
				if (P.levels && threadIdx.x <= AGB_MAXERR && s_hist[threadIdx.x]) atomicAdd(&P.totals[2 + threadIdx.x], s_hist[threadIdx.x]);
			} else {
				out_pos = P.tile_offsets[blockIdx.x] + (s_scan[threadIdx.x] - my_count);
			}
		}
	}
}

/* ================================================================================================
 * stage 1.5: local verification of anchor hits
 *
 * Stage 1 passes every chunk in which an anchor starts; for a pattern made of common words that is a few
 * percent of all chunks, almost none of which belong to a match.  A match that uses the anchor occurrence at
 * text offset t aligns the pat_len pattern positions to text inside [t - off - k, t + pat_len - off + k), so
 * running the SAME recurrence over just that window (all rows started at Init[0], whose separator bit is the
 * always-on start state; no record logic, which can only remove bits) decides whether the hit can matter.
 * Chunks none of whose hits survive lose their bitmap bit.  Warps compact their 1024 chunks into a queue first,
 * so all 32 lanes verify candidates.
 * ============================================================================================== */
#define REFINE_THREADS 128
struct RefineParams {
	const uint8_t *text; uint32_t *bitmap; uint64_t n, n_chunks, n_words;
	const agb_desc *desc;
	uint32_t fold, amask; int na;
	uint32_t anchor[AGB_MAXANCHOR]; int32_t off[AGB_MAXANCHOR];
};

template <typename T, int NR, bool COSTS>
__global__ void __launch_bounds__(REFINE_THREADS)
k_refine(const RefineParams P)
{
	__shared__ T s_mask[257];
	__shared__ uint8_t s_delim[2 * AGB_MAXDELIM + 2];
	__shared__ uint16_t s_queue[REFINE_THREADS / 32][1024];
	__shared__ uint32_t s_keep[REFINE_THREADS / 32][32];
	const agb_desc *D = P.desc;
	for (int i = threadIdx.x; i < 256; i += REFINE_THREADS) s_mask[i] = (T)D->mask[i];
	if (threadIdx.x == 0) s_mask[256] = 0;                   /* byte "256": outside the text, matches nothing */
	if (threadIdx.x < 2 * AGB_MAXDELIM + 2) s_delim[threadIdx.x] = D->delim[threadIdx.x];
	DevConsts<T> C;
	C.init1 = (T)D->init1; C.noerr = (T)D->noerr; C.endpos = (T)D->endpos; C.dendpos = (T)D->dendpos;
	C.L = D->L; C.k = D->k; C.and_mode = D->and_mode; C.inverse = D->inverse; C.kind = D->delim_kind;
	C.ci = D->cost_i; C.cs = D->cost_s; C.cd = D->cost_d;
	const T init0 = (T)D->init0;
	const int pat_len = D->pat_len, k = D->k;
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint64_t warp = ((uint64_t)blockIdx.x * REFINE_THREADS + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * REFINE_THREADS) >> 5;
	const uint64_t n_groups = (P.n_words + 31) / 32;
	Reader R; R.init(P.text, P.n, s_delim, C.L);
	for (uint64_t g = warp; g < n_groups; g += nwarps) {
		const uint64_t w = g * 32 + lane;
		const uint32_t word = (w < P.n_words) ? P.bitmap[w] : 0u;
		if (__ballot_sync(0xffffffffu, word != 0) == 0) continue;
		/* compact the flagged chunks of these 32 words into the warp's queue */
		uint32_t cnt = __popc(word), pre = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= (uint32_t)o) pre += v; }
		const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
		pre -= cnt;
		for (uint32_t b = word; b; b &= b - 1) s_queue[wib][pre++] = (uint16_t)(lane * 32 + (__ffs(b) - 1));
		s_keep[wib][lane] = 0;
		__syncwarp();
		for (uint32_t qi = lane; qi < total; qi += 32) {
			const uint32_t cidx = s_queue[wib][qi];
			const int64_t chunk = (int64_t)(g * 1024 + cidx), base = chunk * 16;
			bool keep = (uint64_t)chunk + 2 >= P.n_chunks;      /* the EOF chunks stay (appended delimiter, bitap.c:161-165) */
			if (!keep) {
				/* which windows of this chunk hit which anchor?  (exactly stage 1's test, exact compare) */
				const uint4 v = __ldg(reinterpret_cast<const uint4 *>(P.text) + chunk);
				const uint32_t nx = __ldg(reinterpret_cast<const uint32_t *>(P.text) + (chunk + 1) * 4);
				uint32_t x[5] = { v.x | P.fold, v.y | P.fold, v.z | P.fold, v.w | P.fold, nx | P.fold };
				for (int s = 0; s < 16 && !keep; s++) {
					const uint32_t wv = __funnelshift_r(x[s >> 2], x[(s >> 2) + 1], (s & 3) * 8) & P.amask;
					for (int a = 0; a < P.na && !keep; a++) {
						if (wv != P.anchor[a]) continue;
						const int64_t t = base + s, ws = t - P.off[a] - k, we = t + pat_len - P.off[a] + k;
						T S[NR];
#pragma unroll
						for (int r = 0; r < NR; r++) S[r] = init0;
						T seen = 0;
						for (int64_t q = ws; q < we; q++) {
							rows_step<T, NR, COSTS>(S, s_mask[R.get(q)], C);
							seen |= S[NR - 1];
						}
						keep = (seen & C.endpos) != 0;
					}
				}
			}
			if (keep) atomicOr(&s_keep[wib][cidx >> 5], 1u << (cidx & 31));
		}
		__syncwarp();
		const uint32_t nw = s_keep[wib][lane];
		if (w < P.n_words && nw != word) P.bitmap[w] = nw;
		__syncwarp();
	}
}

/* exclusive scan of the per-tile counts (one block; the array has n/64KiB entries) */
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t *counts, uint64_t *offsets, uint64_t n_tiles)
{
	__shared__ unsigned long long part[1024];
	const uint64_t per = (n_tiles + 1023) / 1024, a = threadIdx.x * per, b = (a + per < n_tiles) ? a + per : n_tiles;
	unsigned long long s = 0;
	for (uint64_t i = a; i < b; i++) s += counts[i];
	part[threadIdx.x] = s;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1) {
		unsigned long long v = (threadIdx.x >= (unsigned)off) ? part[threadIdx.x - off] : 0;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	unsigned long long run = part[threadIdx.x] - s;
	for (uint64_t i = a; i < b; i++) { offsets[i] = run; run += counts[i]; }
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

struct Workspace {               /* grow-only device scratch, one per device */
	uint32_t *bitmap = nullptr; size_t bitmap_bytes = 0;
	uint32_t *tile_counts = nullptr; uint64_t *tile_offsets = nullptr; size_t tiles = 0;
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
	uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n_words + REC_THREADS - 1) / REC_THREADS;
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

template <typename T, bool COSTS>
static int launch_refine_t(int nrows, const RefineParams &P, unsigned grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: k_refine<T, 1, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 2: k_refine<T, 2, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 3: k_refine<T, 3, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 4: k_refine<T, 4, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 5: k_refine<T, 5, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 6: k_refine<T, 6, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 7: k_refine<T, 7, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 8: k_refine<T, 8, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	case 9: k_refine<T, 9, COSTS><<<grid, REFINE_THREADS, 0, st>>>(P); break;
	default: return -1;
	}
	g_launches++;
	return 0;
}

/* stage 1.5 over the whole bitmap (only when the plan allows a purely local check) */
static int refine_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, cudaStream_t st)
{
	if (!front_usable(d) || !d.refine || n == 0) return AGB_OK;
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32;
	RefineParams P; memset(&P, 0, sizeof P);
	P.text = (const uint8_t *)d_text; P.bitmap = W.bitmap; P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.fold = d.anchor_fold; P.amask = d.anchor_mask; P.na = d.n_anchors;
	for (int i = 0; i < d.n_anchors; i++) { P.anchor[i] = d.anchor[i]; P.off[i] = d.anchor_off[i]; }
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

/* stage 2 (+ tile scan + emit pass when a list is wanted) over the whole text */
static int records_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, bool use_front, int want,
                          int want_level, agb_record *d_records, uint64_t capacity, cudaStream_t st)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n_words + REC_THREADS - 1) / REC_THREADS;
	RecParams P; memset(&P, 0, sizeof P);
	P.text = (const uint8_t *)d_text; P.bitmap = use_front ? W.bitmap : nullptr;
	P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.tile_counts = W.tile_counts; P.tile_offsets = W.tile_offsets; P.records = d_records; P.capacity = capacity;
	P.totals = W.totals; P.emit = 0; P.levels = (want & AGB_WANT_LEVELS) ? 1 : 0; P.want_level = want_level;
	if (!tiles) return AGB_OK;
	if (launch_records(d, P, (unsigned)tiles, st)) return AGB_ERR_ARG;
	CUDA_TRY(cudaGetLastError());
	if ((want & AGB_WANT_RECORDS) && capacity) {
		k_scan_tiles<<<1, 1024, 0, st>>>(W.tile_counts, W.tile_offsets, tiles); g_launches++;
		P.emit = 1;
		if (launch_records(d, P, (unsigned)tiles, st)) return AGB_ERR_ARG;
		CUDA_TRY(cudaGetLastError());
	}
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
	const bool use_front = front_usable(d) && n > 0;
	if (use_front) { rc = front_launch(d, W, d_text, n, 0, ~0ull, false, st); if (rc) return rc; }
	CUDA_TRY(cudaEventRecord(W.e1, st));
	if (use_front) { rc = refine_launch(d, W, d_text, n, st); if (rc) return rc; }
	rc = records_launch(d, W, d_text, n, use_front, want, want_level, d_records, capacity, st); if (rc) return rc;
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
 * moved in 64 MiB slices on a copy stream (straight from the caller's memory when it is page-locked, else
 * through a pinned ring filled by 4 host threads) while stage 1 runs on the slice that arrived before, so
 * the scan hides behind PCIe; stage 2 runs once over the whole bitmap. */
extern "C" int agb_scan_host(const agb_pattern *p, const void *h_text, uint64_t n, int want,
                             agb_record *records, uint64_t capacity, agb_result *res)
{
	if (!p || !res || (!h_text && n)) return AGB_ERR_ARG;
	if ((want & AGB_WANT_RECORDS) && capacity && !records) return AGB_ERR_ARG;
	memset(res, 0, sizeof *res);
	const agb_desc &d = p->d;
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
	cudaPointerAttributes attr; memset(&attr, 0, sizeof attr);
	bool pinned = n && cudaPointerGetAttributes(&attr, h_text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!pinned && n && !W.stage[0]) for (int i = 0; i < STAGE_BUFS; i++) CUDA_TRY(cudaMallocHost(&W.stage[i], H2D_SLICE));
	const bool use_front = front_usable(d) && n > 0;
	const uint64_t words_per_slice = H2D_SLICE / 512, n_slices = (n + H2D_SLICE - 1) / H2D_SLICE;
	CUDA_TRY(cudaMemsetAsync(W.totals, 0, 16 * sizeof(unsigned long long), W.s_comp));
	CUDA_TRY(cudaEventRecord(W.e0, W.s_comp));
	/* zero the slack after the text once (stage 1 reads whole 16-byte chunks, stage 2 whole groups) */
	CUDA_TRY(cudaMemsetAsync(W.h2d_text + (n & ~(uint64_t)15), 0, need - (n & ~(uint64_t)15), W.s_copy));
	for (uint64_t i = 0; i < n_slices; i++) {
		const uint64_t off = i * H2D_SLICE, len = std::min<uint64_t>(H2D_SLICE, n - off);
		const int sb = (int)(i % STAGE_BUFS);
		if (pinned) {
			CUDA_TRY(cudaMemcpyAsync(W.h2d_text + off, (const uint8_t *)h_text + off, len, cudaMemcpyHostToDevice, W.s_copy));
		} else {
			if (i >= STAGE_BUFS) CUDA_TRY(cudaEventSynchronize(W.ev_copy[sb]));     /* that staging buffer has been consumed */
			par_memcpy(W.stage[sb], (const uint8_t *)h_text + off, len);
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
	if (use_front) { rc = refine_launch(d, W, W.h2d_text, n, W.s_comp); if (rc) return rc; }
	rc = records_launch(d, W, W.h2d_text, n, use_front, want, -1, W.h2d_rec, capacity, W.s_comp); if (rc) return rc;
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

extern "C" int agb_scan_fd(const agb_pattern *p, int fd, int want, agb_record *records, uint64_t capacity, agb_result *res)
{
	/* fill_buf() (bitap.c:450-477): read(2) until EOF; here into one growing host buffer, then agb_scan_host */
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
