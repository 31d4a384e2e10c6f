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
#define FRONT_UNROLL  4          /* 16-byte loads in flight per thread */

struct FrontParams {
	const uint4 *text;           /* 16-byte aligned */
	uint32_t    *bitmap;         /* one word per 32 chunks */
	uint64_t     n;              /* bytes */
	uint64_t     n_chunks;       /* ceil(n/16) */
	uint64_t     n_words;        /* ceil(n_chunks/32) */
	uint32_t     fold, amask;
	uint32_t     anchor[AGB_MAXANCHOR];
};

__device__ __forceinline__ uint4 ld_stream16(const uint4 *p)
{
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

/* does any window of the word pair (lo = bytes 0..3, hi = following word) equal an anchor?
 * windows start at byte 0,1,2,3 of lo.  NA anchors live in the constant bank (kernel parameters). */
template <int NA, bool MASKED>
__device__ __forceinline__ uint32_t windows_min(uint32_t lo, uint32_t hi, const FrontParams &P, uint32_t acc)
{
	uint32_t w0 = lo, w1 = __funnelshift_r(lo, hi, 8), w2 = __funnelshift_r(lo, hi, 16), w3 = __funnelshift_r(lo, hi, 24);
	if (MASKED) { w0 &= P.amask; w1 &= P.amask; w2 &= P.amask; w3 &= P.amask; }
#pragma unroll
	for (int a = 0; a < NA; a++) {
		/* unsigned min of differences is 0 iff some window equals some anchor (DPX min3: VIMNMX3) */
		uint32_t A = P.anchor[a];
		acc = __vimin3_u32(acc, w0 - A, w1 - A);
		acc = __vimin3_u32(acc, w2 - A, w3 - A);
	}
	return acc;
}

template <int NA, bool MASKED, bool FOLD>
__global__ void __launch_bounds__(FRONT_THREADS)
k_front(const FrontParams P)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * FRONT_THREADS + threadIdx.x) >> 5;
	const uint64_t nwarps = ((uint64_t)gridDim.x * FRONT_THREADS) >> 5;
	/* each warp takes groups of FRONT_UNROLL consecutive bitmap words (= FRONT_UNROLL * 512 bytes) */
	const uint64_t n_groups = (P.n_words + FRONT_UNROLL - 1) / FRONT_UNROLL;
	for (uint64_t g = warp; g < n_groups; g += nwarps) {
		const uint64_t w0 = g * FRONT_UNROLL;
		uint4 v[FRONT_UNROLL + 1];
#pragma unroll
		for (int u = 0; u < FRONT_UNROLL; u++) {
			uint64_t c = (w0 + u) * 32 + lane;
			v[u] = (c < P.n_chunks) ? ld_stream16(P.text + c) : make_uint4(0, 0, 0, 0);
		}
		{   /* the word that follows this group: only its first 4 bytes are needed, by lane 31 of the last word */
			uint64_t c = (w0 + FRONT_UNROLL) * 32;
			uint32_t nx = 0;
			if (lane == 0 && c < P.n_chunks) nx = __ldg(reinterpret_cast<const uint32_t *>(P.text + c));
			v[FRONT_UNROLL] = make_uint4(nx, 0, 0, 0);
		}
#pragma unroll
		for (int u = 0; u < FRONT_UNROLL; u++) {
			if (w0 + u >= P.n_words) break;
			/* lane L needs the first word of the next chunk: lane L+1 of this vector, or lane 0 of the next one */
			uint32_t give = (lane == 0) ? v[u + 1].x : v[u].x;
			uint32_t x4 = __shfl_sync(0xffffffffu, give, (lane + 1) & 31);
			uint32_t x0 = v[u].x, x1 = v[u].y, x2 = v[u].z, x3 = v[u].w;
			if (FOLD) { x0 |= P.fold; x1 |= P.fold; x2 |= P.fold; x3 |= P.fold; x4 |= P.fold; }
			uint32_t acc = 0xffffffffu;
			acc = windows_min<NA, MASKED>(x0, x1, P, acc);
			acc = windows_min<NA, MASKED>(x1, x2, P, acc);
			acc = windows_min<NA, MASKED>(x2, x3, P, acc);
			acc = windows_min<NA, MASKED>(x3, x4, P, acc);
			uint64_t c = (w0 + u) * 32 + lane;
			/* the last chunks are always passed on: a match may run into the delimiter appended at EOF (bitap.c:161-165) */
			bool flag = (c < P.n_chunks) && (acc == 0 || c + 2 >= P.n_chunks);
			uint32_t word = __ballot_sync(0xffffffffu, flag);
			if (lane == 0) P.bitmap[w0 + u] = word;
		}
	}
}

/* with FOLD the fold mask must not be applied twice to the shifted-in word: windows are built from
 * already folded words, so x4 is folded above and funnel shifts commute with the byte-wise OR. */

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

__global__ void __launch_bounds__(128) k_corpus(uint8_t *out, uint64_t n_pages, const CorpusParams P)
{
	/* one thread per page keeps the generator identical to the host loop; pages are staged in
	 * shared memory? no: 4 KiB per thread is too much -- write straight to HBM (generation is not timed) */
	uint64_t pg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (pg >= n_pages) return;
	agb_corpus_page(out + pg * AGB_PAGE, P.s.seed, P.s.first_page + pg, c_vocab, c_woff,
	                P.s.paragraphs, P.s.needle_every, P.s.needle, P.needle_len, P.s.needle_maxedits);
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
		k_corpus<<<(unsigned)((n_pages + 127) / 128), 128, 0, st>>>((uint8_t *)d_text, n_pages, P);
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
struct Workspace {               /* grow-only device scratch, one per device */
	uint32_t *bitmap = nullptr; size_t bitmap_bytes = 0;
	uint32_t *tile_counts = nullptr; uint64_t *tile_offsets = nullptr; size_t tiles = 0;
	unsigned long long *totals = nullptr;          /* 12 counters */
	unsigned long long *h_totals = nullptr;        /* pinned */
	agb_desc *d_desc = nullptr; agb_desc h_desc_copy; bool desc_valid = false;
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
	int sm_count = 0;
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
	size_t bb = (size_t)(n_words + FRONT_UNROLL) * 4;
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

template <int NA>
static void launch_front_na(const FrontParams &P, bool masked, bool fold, unsigned grid, cudaStream_t st)
{
	if (masked) { if (fold) k_front<NA, true, true><<<grid, FRONT_THREADS, 0, st>>>(P); else k_front<NA, true, false><<<grid, FRONT_THREADS, 0, st>>>(P); }
	else        { if (fold) k_front<NA, false, true><<<grid, FRONT_THREADS, 0, st>>>(P); else k_front<NA, false, false><<<grid, FRONT_THREADS, 0, st>>>(P); }
}

static int launch_front(const agb_desc &d, const FrontParams &P, unsigned grid, cudaStream_t st)
{
	bool masked = d.anchor_mask != 0xFFFFFFFFu, fold = d.anchor_fold != 0;
	switch (d.n_anchors) {
	case 1: launch_front_na<1>(P, masked, fold, grid, st); break;
	case 2: launch_front_na<2>(P, masked, fold, grid, st); break;
	case 3: launch_front_na<3>(P, masked, fold, grid, st); break;
	case 4: launch_front_na<4>(P, masked, fold, grid, st); break;
	case 5: launch_front_na<5>(P, masked, fold, grid, st); break;
	case 6: launch_front_na<6>(P, masked, fold, grid, st); break;
	case 7: launch_front_na<7>(P, masked, fold, grid, st); break;
	case 8: launch_front_na<8>(P, masked, fold, grid, st); break;
	case 9: launch_front_na<9>(P, masked, fold, grid, st); break;
	default: return -1;
	}
	g_launches++;
	return 0;
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
	if (!W.desc_valid || memcmp(&W.h_desc_copy, &d, sizeof d) != 0) {
		CUDA_TRY(cudaMemcpyAsync(W.d_desc, &d, sizeof d, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaStreamSynchronize(st));     /* &d may be on the caller's stack */
		W.h_desc_copy = d; W.desc_valid = true;
	}
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n_words + REC_THREADS - 1) / REC_THREADS;
	CUDA_TRY(cudaMemsetAsync(W.totals, 0, 16 * sizeof(unsigned long long), st));
	CUDA_TRY(cudaEventRecord(W.e0, st));
	const bool use_front = d.plan == AGB_PLAN_ANCHORS && d.n_anchors >= 1 && d.n_anchors <= 9 && n_words > 0;
	if (use_front) {
		FrontParams F; memset(&F, 0, sizeof F);
		F.text = (const uint4 *)d_text; F.bitmap = W.bitmap; F.n = n; F.n_chunks = n_chunks; F.n_words = n_words;
		F.fold = d.anchor_fold; F.amask = d.anchor_mask;
		for (int i = 0; i < d.n_anchors; i++) F.anchor[i] = d.anchor[i];
		uint64_t groups = (n_words + FRONT_UNROLL - 1) / FRONT_UNROLL;
		uint64_t want_blocks = (groups * 32 + FRONT_THREADS - 1) / FRONT_THREADS;
		unsigned grid = (unsigned)std::min<uint64_t>(want_blocks, (uint64_t)W.sm_count * 8);
		if (launch_front(d, F, grid ? grid : 1, st)) return AGB_ERR_ARG;
		CUDA_TRY(cudaGetLastError());
	}
	CUDA_TRY(cudaEventRecord(W.e1, st));
	RecParams P; memset(&P, 0, sizeof P);
	P.text = (const uint8_t *)d_text; P.bitmap = use_front ? W.bitmap : nullptr;
	P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.tile_counts = W.tile_counts; P.tile_offsets = W.tile_offsets; P.records = d_records; P.capacity = capacity;
	P.totals = W.totals; P.emit = 0; P.levels = (want & AGB_WANT_LEVELS) ? 1 : 0; P.want_level = want_level;
	if (tiles) {
		if (launch_records(d, P, (unsigned)tiles, st)) return AGB_ERR_ARG;
		CUDA_TRY(cudaGetLastError());
		if ((want & AGB_WANT_RECORDS) && capacity) {
			k_scan_tiles<<<1, 1024, 0, st>>>(W.tile_counts, W.tile_offsets, tiles); g_launches++;
			P.emit = 1;
			if (launch_records(d, P, (unsigned)tiles, st)) return AGB_ERR_ARG;
			CUDA_TRY(cudaGetLastError());
		}
	}
	CUDA_TRY(cudaEventRecord(W.e2, st));
	CUDA_TRY(cudaMemcpyAsync(W.h_totals, W.totals, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	res->n_matched = W.h_totals[0];
	res->n_flagged = W.h_totals[1];
	for (int i = 0; i <= AGB_MAXERR; i++) res->level_hist[i] = W.h_totals[2 + i];
	res->n_records = (want & AGB_WANT_RECORDS) ? std::min<uint64_t>(res->n_matched, capacity) : 0;
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

/* host text: for now one staged copy + one scan per slice of at most SLICE bytes cut at record boundaries */
extern "C" int agb_scan_host(const agb_pattern *p, const void *h_text, uint64_t n, int want,
                             agb_record *records, uint64_t capacity, agb_result *res)
{
	if (!p || !res || (!h_text && n)) return AGB_ERR_ARG;
	uint8_t *d_text = nullptr; agb_record *d_rec = nullptr;
	size_t alloc = (size_t)((n + 15) / 16 * 16 + 16);
	CUDA_TRY(cudaMalloc(&d_text, alloc));
	cudaError_t e = cudaMemcpy(d_text, h_text, n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemset(d_text + n, 0, alloc - n);
	if (e != cudaSuccess) { cudaFree(d_text); snprintf(g_err, sizeof g_err, "H2D copy failed: %s", cudaGetErrorString(e)); return AGB_ERR_CUDA; }
	if ((want & AGB_WANT_RECORDS) && capacity) {
		e = cudaMalloc(&d_rec, capacity * sizeof(agb_record));
		if (e != cudaSuccess) { cudaFree(d_text); snprintf(g_err, sizeof g_err, "cudaMalloc(records) failed: %s", cudaGetErrorString(e)); return AGB_ERR_CUDA; }
	}
	int rc = scan_device_impl(p->d, d_text, n, want, -1, d_rec, capacity, 0, res);
	if (rc == AGB_OK && d_rec && res->n_records) {
		e = cudaMemcpy(records, d_rec, res->n_records * sizeof(agb_record), cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "D2H copy failed: %s", cudaGetErrorString(e)); rc = AGB_ERR_CUDA; }
	}
	cudaFree(d_text); if (d_rec) cudaFree(d_rec);
	return rc;
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
