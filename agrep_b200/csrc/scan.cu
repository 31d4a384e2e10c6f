/* agrep_b200/csrc/scan.cu -- the host side of libagrepb200's scan path and its C ABI (include/agrep_b200.h).
 *
 * What the reference does in bitap()/asearch()/asearch0()/asearch1()/sgrep()+bm() (one byte at a time, one file
 * block at a time, bitap.c:169-284, asearch.c:94-306, :620-774, asearch1.c:86-235, sgrep.c:694-1016) is done in
 * device stages over text that is resident in HBM:
 *
 *   stage 1    k_front (front.cu)    which 16-byte chunks can matter: one of the pattern's k+1 disjoint literal
 *              anchors starts there (pigeonhole, pattern.c:plan_anchors); one bit per chunk.  HBM-bound.
 *   stage 1.5  k_refine (refine.cu)  the same recurrence over just the window around an anchor hit; chunks whose
 *              hits cannot belong to a match lose their bit.
 *   stage 2    records (records.cu, slices.cu)  exact: the recurrence from the record start in the constant
 *              post-delimiter state until the closing delimiter, the reference's match test and bookkeeping,
 *              ordered (lasti, print_end) lists by count pass -> scan -> emit pass.  List form for sparse
 *              survivors, slices / dense tile form when every byte has to be walked.
 *   ordinals   (aux.cu)  the j that -n prints, from delimiter counts.
 *
 * This file: the per-device workspace, which form runs when (records_launch), the streaming host entry points
 * (the fill_buf replacement) and the exported functions.  There is no CPU path anywhere in the library.
 */
#include "scan_internal.cuh"
#include "corpus.h"
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <mutex>
#include <thread>

/* ------------------------------------------------------------------------------------------------ */
thread_local char g_err[512];
std::atomic<uint64_t> g_launches{0};

extern "C" const char *agb_last_error(void) { return g_err; }
extern "C" const char *agb_version(void) { return "agrep-b200 0.1 (sm_100a)"; }
extern "C" uint64_t agb_kernel_launches(void) { return g_launches.load(); }
/* frees the per-device scratch of this process (bitmaps, candidate lists, pinned rings, streams, events); the next
 * scan allocates again */
extern "C" void agb_shutdown(void)
{
	int cur = 0; cudaGetDevice(&cur);
	for (int dev = 0; dev < 64; dev++) {
		std::lock_guard<std::mutex> lk(g_ws_mu[dev]);
		Workspace &W = g_ws[dev];
		if (!W.totals && !W.bitmap && !W.h2d_text) continue;
		if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); continue; }
		cudaDeviceSynchronize();
		cudaFree(W.bitmap); cudaFree(W.bitmap2); cudaFree(W.range_counts); cudaFree(W.range_offsets);
		cudaFree(W.tile_counts); cudaFree(W.tile_offsets); cudaFree(W.cand); cudaFree(W.cand_counts); cudaFree(W.cand_offsets);
		cudaFree(W.cand_first); cudaFree(W.scan_sums); cudaFree(W.scan_offs); cudaFree(W.ord_blocks); cudaFree(W.totals);
		cudaFreeHost(W.h_totals); cudaFree(W.d_desc); cudaFree(W.h2d_text); cudaFree(W.h2d_rec); cudaFree(W.d_gram); cudaFreeHost(W.h_gram);
		if (W.e0) cudaEventDestroy(W.e0); if (W.e1) cudaEventDestroy(W.e1); if (W.e2) cudaEventDestroy(W.e2);
		for (int i = 0; i < STAGE_BUFS; i++) { if (W.ev_copy[i]) cudaEventDestroy(W.ev_copy[i]); if (W.stage[i]) cudaFreeHost(W.stage[i]); }
		if (W.s_copy) cudaStreamDestroy(W.s_copy); if (W.s_comp) cudaStreamDestroy(W.s_comp);
		W = Workspace();
	}
	cudaSetDevice(cur);
}
extern "C" int agb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
extern "C" int agb_set_device(int dev) { CUDA_TRY(cudaSetDevice(dev)); return AGB_OK; }

/* ================================================================================================
 * host side of the scan
 * ============================================================================================== */

Workspace g_ws[64];
std::mutex g_ws_mu[64];          /* one scan at a time per DEVICE (its workspace is shared scratch); different devices run side by side */

static int ws_prepare(Workspace &W, uint64_t n)
{
	if (!W.totals) {
		CUDA_TRY(cudaMalloc(&W.totals, 24 * sizeof(unsigned long long)));
		CUDA_TRY(cudaMallocHost(&W.h_totals, 24 * sizeof(unsigned long long)));
		CUDA_TRY(cudaMalloc(&W.d_desc, sizeof(agb_desc)));
		CUDA_TRY(cudaMalloc(&W.range_counts, REFINE_MAX_RANGES * sizeof(uint32_t)));
		CUDA_TRY(cudaMalloc(&W.range_offsets, REFINE_MAX_RANGES * sizeof(uint64_t)));
		CUDA_TRY(cudaEventCreate(&W.e0)); CUDA_TRY(cudaEventCreate(&W.e1)); CUDA_TRY(cudaEventCreate(&W.e2));
		int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
		CUDA_TRY(cudaDeviceGetAttribute(&W.sm_count, cudaDevAttrMultiProcessorCount, dev));
	}
	uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n + std::min<uint64_t>(DENSE_TILE, SL_TILE) - 1) / std::min<uint64_t>(DENSE_TILE, SL_TILE) + 1;
	size_t bb = (size_t)(n_words + FRONT_WORDS_PER_STAGE) * 4;
	if (bb > W.bitmap_bytes) {
		if (W.bitmap) cudaFree(W.bitmap);
		if (W.bitmap2) cudaFree(W.bitmap2);
		W.bitmap = nullptr; W.bitmap2 = nullptr; W.bitmap_bytes = 0;
		CUDA_TRY(cudaMalloc(&W.bitmap, bb));
		CUDA_TRY(cudaMalloc(&W.bitmap2, bb)); W.bitmap_bytes = bb;
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

/* the ordered candidate list -> records: count launch (per-candidate counts, the first record of each kept), scan,
 * emit launch.  The list length lives on the device (totals[12]); every grid here is sized by the list's capacity. */
static int list_stage(const agb_desc &d, Workspace &W, RecParams &P, bool want_list, cudaStream_t st)
{
	P.cand = W.cand; P.cand_cap = W.cand_cap; P.tile_counts = W.cand_counts; P.tile_offsets = W.cand_offsets;
	P.cand_first = want_list ? W.cand_first : nullptr;
	const unsigned grid = (unsigned)std::min<uint64_t>((W.cand_cap + REC_THREADS - 1) / REC_THREADS, (uint64_t)W.sm_count * 16);
	if (launch_records_list(d, P, grid, st)) return AGB_ERR_ARG;
	CUDA_TRY(cudaGetLastError());
	if (want_list) {
		const unsigned nb = (unsigned)((W.cand_cap + SCAN_BLOCK - 1) / SCAN_BLOCK);
		k_scan_partial<<<nb, 1024, 0, st>>>(W.cand_counts, W.cand_cap, W.scan_sums, W.totals + 12);
		k_scan_tiles<<<1, 1024, 0, st>>>(W.scan_sums, W.scan_offs, nb, nullptr);
		k_scan_apply<<<nb, 1024, 0, st>>>(W.cand_counts, W.cand_cap, W.scan_offs, W.cand_offsets, W.totals + 12);
		g_launches += 3;
		P.emit = 1;
		if (launch_records_list(d, P, grid, st)) return AGB_ERR_ARG;
		CUDA_TRY(cudaGetLastError());
	}
	return AGB_OK;
}

/* stage 2 over the whole text.  After stage 1.5 the survivors are few: they are compacted into an ordered list
 * and each gets its own thread (count launch -> scan -> emit launch).  Otherwise (or if the list would not fit)
 * the dense form walks the bitmap, one thread per word.
 * refined: stage 1.5 ran -- the survivors are in W.bitmap2, their per-range counts in W.range_counts, and nothing
 * here needs the host to know how many there are (the list is sized by W.cand_hint / a fraction of the chunks; the
 * caller checks totals[12] against W.cand_cap afterwards and comes back with refined_retry set if it was too small). */
static int records_launch(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, bool use_front, bool refined, int want,
                          int want_level, agb_record *d_records, uint64_t capacity, cudaStream_t st, const ShardInfo *sh)
{
	const uint64_t n_chunks = (n + 15) / 16, n_words = (n_chunks + 31) / 32, tiles = (n_words + REC_THREADS - 1) / REC_THREADS;
	RecParams P; memset(&P, 0, sizeof P);
	P.text = (const uint8_t *)d_text; P.bitmap = use_front ? (refined ? W.bitmap2 : W.bitmap) : nullptr;
	P.n = n; P.n_chunks = n_chunks; P.n_words = n_words; P.desc = W.d_desc;
	P.records = d_records; P.capacity = capacity;
	P.totals = W.totals; P.emit = 0; P.levels = (want & AGB_WANT_LEVELS) ? 1 : 0; P.want_level = want_level;
	P.own_lo = sh ? sh->own_lo : INT64_MIN; P.own_hi = sh ? sh->own_hi : INT64_MAX; P.shard_last = sh ? sh->last : 1;
	if (!tiles) return AGB_OK;
	const bool want_list = (want & AGB_WANT_RECORDS) && capacity;
	if (use_front && refined) {
		int rc = ws_cand_reserve(W, std::max<size_t>(W.cand_hint + W.cand_hint / 4, (size_t)(n_chunks / 512) + 65536)); if (rc) return rc;
		const unsigned ranges = W.refine_ctas * (REFINE_THREADS / 32);
		k_scan_tiles<<<1, 1024, 0, st>>>(W.range_counts, W.range_offsets, ranges, W.totals + 12); g_launches++;
		rc = compact_ranges_launch(W, n, st); if (rc) return rc;
		return list_stage(d, W, P, want_list, st);
	}
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
			return list_stage(d, W, P, want_list, st);
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

/* an exact pattern that is no longer than its anchor ('the'): every chunk stage 1 flags holds a real occurrence, so
 * stage 1.5 has nothing to remove -- if such flags are dense (front_is_dense), the record stage walks every byte
 * anyway and stages 1.5 and the compaction are skipped.  For every other pattern only stage 1.5 can tell (a k = 4
 * pattern of common words flags 8 % of the chunks and keeps none), so the decision waits for the list length. */
static bool refine_cannot_thin(const agb_desc &d) { return d.k == 0 && d.n_anchors == 1 && d.pat_len <= d.anchor_len; }

static int fetch_result(Workspace &W, int want, uint64_t capacity, bool refined, cudaStream_t st, agb_result *res)
{
	CUDA_TRY(cudaMemcpyAsync(W.h_totals, W.totals, 24 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));   /* [16..18]: shard_aux_enqueue */
	CUDA_TRY(cudaStreamSynchronize(st));
	res->n_matched = W.h_totals[0];
	res->n_flagged = refined ? W.h_totals[12] : W.h_totals[1];
	for (int i = 0; i <= AGB_MAXERR; i++) res->level_hist[i] = W.h_totals[2 + i];
	res->n_records = (want & AGB_WANT_RECORDS) ? std::min<uint64_t>(res->n_matched, capacity) : 0;
	res->truncated = ((want & AGB_WANT_RECORDS) && res->n_matched > capacity) ? 1 : 0;
	res->n_closes = (want & AGB_WANT_ORDINALS) ? W.h_totals[13] + (uint64_t)W.ord_virt : 0;
	return AGB_OK;
}

/* the block counters stage 1 fills when it counts delimiters: allocated, and zeroed behind the last bitmap word
 * (the tile sums read whole tiles; the block of the delimiter appended at EOF may lie there) */
static int ordinals_prepare_blocks(const agb_desc &d, Workspace &W, uint64_t n, cudaStream_t st)
{
	int rc = ordinals_reserve(d, W, n); if (rc) return rc;
	const uint64_t n_words = ((n + 15) / 16 + 31) / 32;
	if (W.ord_blocks_cap > n_words) CUDA_TRY(cudaMemsetAsync(W.ord_blocks + n_words, 0, (W.ord_blocks_cap - n_words) * sizeof(uint16_t), st));
	return AGB_OK;
}

/* ------------------------------------------------------------------------------------------------
 * the anchor planner.  Any k+1 disjoint literal grams of the pattern make a valid pigeonhole filter; which ones
 * decides how many chunks stage 1 flags -- 4.5 % of the benchmark text for beca|use |each, 2.4 % for beca|se e|ach --
 * and stage 1.5 pays per flagged chunk.  The static plan (pattern.c) takes the first k+1 runs; here, for texts large
 * enough to care, the candidate grams (every literal 4-gram and 3-gram of the pattern) are counted on a 4 MiB sample
 * of the text and the cheapest set is chosen by a small dynamic program: four-byte grams, plus up to two three-byte
 * grams (which cost stage 1 one more operation per window: they only pay when they save enough flags).
 * Works from the Mask[] words alone, so the drop-in layer's descriptors are re-planned too.
 * ---------------------------------------------------------------------------------------------- */
#define PLAN_MIN_BYTES   (256ull << 20)
#define PLAN_SAMPLE_BLK  64               /* stretches */
#define PLAN_BLK_CHUNKS  4096             /* of 64 KiB */
static int adaptive_plan(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, cudaStream_t st, agb_desc *out)
{
	*out = d;
	if (!d.adaptive || !front_usable(d) || !d.refine || d.n_anchors3 || n < PLAN_MIN_BYTES || d.pat_len < 4 || d.pat_len > 60) return AGB_OK;
	const int need = d.k + 1;
	if (need > 9) return AGB_OK;
	uint64_t key = 1469598103934665603ull;
	{
		const unsigned char *b = (const unsigned char *)&d;
		for (size_t i = 0; i < sizeof d; i++) { key ^= b[i]; key *= 1099511628211ull; }
		key ^= (uint64_t)(uintptr_t)d_text; key *= 1099511628211ull; key ^= n; key *= 1099511628211ull;
	}
	if (W.plan_valid && W.plan_key == key) { *out = W.plan_desc; return AGB_OK; }
	/* literal bytes of the pattern proper, from the masks */
	int lit[64]; bool pair_any = false;
	for (int j = 0; j < d.pat_len; j++) {
		const uint64_t bit = 1ull << (d.M - (d.L + 2 + j));
		int cnt = 0, c0 = -1, c1 = -1;
		for (int c = 0; c < 256; c++) if (d.mask[c] & bit) { if (cnt == 0) c0 = c; else if (cnt == 1) c1 = c; cnt++; }
		lit[j] = -1;
		if (cnt == 1 && c0 != '\n' && c0 < 0x80) lit[j] = c0;
		else if (cnt == 2 && (c0 ^ c1) == 0x20 && c1 < 0x80) { lit[j] = c0 | 0x20; pair_any = true; }
	}
	const uint32_t fold = pair_any ? 0x20202020u : 0u;
	/* candidate grams */
	struct Gram { int s, len; uint32_t v, m; double cost; };
	Gram g[128]; int ng = 0;
	for (int len = 4; len >= 3; len--)
		for (int s = 0; s + len <= d.pat_len && ng < 120; s++) {
			bool ok = true; uint32_t v = 0;
			for (int t = 0; t < len; t++) { if (lit[s + t] < 0) ok = false; else v |= (uint32_t)(lit[s + t] | (fold & 0x20)) << (8 * t); }
			if (!ok) continue;
			g[ng].s = s; g[ng].len = len; g[ng].v = v; g[ng].m = len == 4 ? 0xFFFFFFFFu : 0x00FFFFFFu; g[ng].cost = 0; ng++;
		}
	if (ng < need) return AGB_OK;
	if (!W.d_gram) { CUDA_TRY(cudaMalloc(&W.d_gram, 3 * 128 * sizeof(uint32_t))); CUDA_TRY(cudaMallocHost(&W.h_gram, 3 * 128 * sizeof(uint32_t))); }
	for (int i = 0; i < 128; i++) { W.h_gram[i] = i < ng ? g[i].v : 0; W.h_gram[128 + i] = i < ng ? g[i].m : 0; W.h_gram[256 + i] = 0; }
	CUDA_TRY(cudaMemcpyAsync(W.d_gram, W.h_gram, 3 * 128 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	const uint64_t n_chunks = (n + 15) / 16;
	const uint64_t threads = (uint64_t)PLAN_SAMPLE_BLK * PLAN_BLK_CHUNKS;
	k_gram_sample<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const uint8_t *)d_text, n_chunks, PLAN_SAMPLE_BLK, PLAN_BLK_CHUNKS,
	                                                                 ng, W.d_gram, W.d_gram + 128, fold, W.d_gram + 256);
	g_launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(W.h_gram + 256, W.d_gram + 256, 128 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	for (int i = 0; i < ng; i++) g[i].cost = (double)W.h_gram[256 + i] / (double)threads;
	/* dynamic program over the pattern positions: f[p][j][t] = least total rate of j disjoint grams inside [0, p), t of them
	 * three bytes long */
	/* what a three-byte group costs, in units of flag rate: measured on the benchmark pattern (beca|se e|ach flags 2.4 %
	 * of the chunks instead of 4.5 %), stage 1 ran 22 % longer (14.5 instead of 11.8 ms per 64 GiB: the second polynomial
	 * and one VIMNMX3 per window instead of half of one) and stage 1.5 did not get cheaper in proportion -- a mixed plan
	 * only pays when the four-byte grams of a piece are really common */
	const char *mp = getenv("AGB_PLAN_MIXED");           /* (tests force mixed plans with AGB_PLAN_MIXED=0) */
	const double INF = 1e30, MIXED = mp ? atof(mp) : 0.03;
	static double f[66][10][3]; static int from[66][10][3];   /* gram taken to get here, -1: position skipped */
	for (int p = 0; p <= d.pat_len; p++) for (int j = 0; j <= need; j++) for (int t = 0; t < 3; t++) { f[p][j][t] = INF; from[p][j][t] = -2; }
	f[0][0][0] = 0;
	for (int p = 0; p < d.pat_len; p++)
		for (int j = 0; j <= need; j++) for (int t = 0; t < 3; t++) {
			if (f[p][j][t] >= INF) continue;
			if (f[p][j][t] < f[p + 1][j][t]) { f[p + 1][j][t] = f[p][j][t]; from[p + 1][j][t] = -1; }
			if (j == need) continue;
			for (int i = 0; i < ng; i++) if (g[i].s == p) {
				const int t2 = t + (g[i].len == 3), p2 = p + g[i].len;
				if (t2 > 2) continue;
				if (f[p][j][t] + g[i].cost < f[p2][j + 1][t2]) { f[p2][j + 1][t2] = f[p][j][t] + g[i].cost; from[p2][j + 1][t2] = i; }
			}
		}
	int best_t = -1; double best = INF;
	for (int t = 0; t < 3; t++) {
		if (need - t < 1 || need - t > (t ? 7 : 9)) continue;
		const double c = f[d.pat_len][need][t] + (t ? MIXED : 0);
		if (c < best) { best = c; best_t = t; }
	}
	double cur = 0;                                     /* the static plan's rate, from the same sample where it can be read off */
	for (int a = 0; a < d.n_anchors; a++) { double r = 1.0; for (int i = 0; i < ng; i++) if (g[i].len == d.anchor_len && g[i].s == d.anchor_off[a]) r = g[i].cost; cur += r; }
	if (best_t < 0 || !(best < 0.85 * cur)) { W.plan_key = key; W.plan_valid = true; W.plan_desc = d; return AGB_OK; }
	agb_desc nd = d;
	nd.n_anchors = 0; nd.n_anchors3 = 0; nd.anchor_len = 4; nd.anchor_mask = 0xFFFFFFFFu; nd.anchor_fold = fold;
	{
		int p = d.pat_len, j = need, t = best_t;
		while (p > 0 && j >= 0) {
			const int fr = from[p][j][t];
			if (fr == -1) { p--; continue; }
			if (fr < 0) break;
			if (g[fr].len == 4) { nd.anchor[nd.n_anchors] = g[fr].v; nd.anchor_off[nd.n_anchors] = g[fr].s; nd.n_anchors++; }
			else { nd.anchor3[nd.n_anchors3] = g[fr].v; nd.anchor3_off[nd.n_anchors3] = g[fr].s; nd.n_anchors3++; }
			p -= g[fr].len; j--; t -= (g[fr].len == 3);
		}
	}
	/* the kernels want their polynomials: both groups must pass the false-positive guard, and the anchors must be
	 * pairwise distinct (stage 1.5 tells them apart by their bytes) */
	bool ok = nd.n_anchors + nd.n_anchors3 == need && nd.n_anchors >= 1;
	uint32_t tmp[AGB_MAXANCHOR];
	if (ok) ok = poly_setup(nd.anchor, nd.n_anchors, 32, tmp);
	if (ok && nd.n_anchors3) ok = poly_setup(nd.anchor3, nd.n_anchors3, 24, tmp);
	for (int a = 0; a < nd.n_anchors && ok; a++) {
		for (int b = 0; b < a; b++) if (nd.anchor[a] == nd.anchor[b]) ok = false;
		for (int b = 0; b < nd.n_anchors3; b++) if ((nd.anchor[a] & 0x00FFFFFFu) == nd.anchor3[b]) ok = false;
	}
	for (int a = 0; a < nd.n_anchors3 && ok; a++) for (int b = 0; b < a; b++) if (nd.anchor3[a] == nd.anchor3[b]) ok = false;
	W.plan_key = key; W.plan_valid = true; W.plan_desc = ok ? nd : d;
	*out = W.plan_desc;
	if (getenv("AGB_DEBUG_PLAN")) {
		fprintf(stderr, "agb plan: static rate %.4f -> %s rate %.4f:", cur, ok ? "chosen" : "rejected", best);
		for (int a = 0; a < nd.n_anchors; a++) fprintf(stderr, " [%.4s]@%d", (const char *)&nd.anchor[a], nd.anchor_off[a]);
		for (int a = 0; a < nd.n_anchors3; a++) fprintf(stderr, " [%.3s]@%d", (const char *)&nd.anchor3[a], nd.anchor3_off[a]);
		fprintf(stderr, "\n");
	}
	return AGB_OK;
}

/* everything after stage 1, on one stream: stage 1.5, the record stage, the ordinals, the result read-back (the one
 * host synchronisation of a scan).  The candidate list of the list form is sized without asking the device how many
 * survivors there are; should it turn out too small (totals[12] > capacity, seen in the read-back) the record stage
 * alone is run again with the right size -- or in its every-byte form when the survivors are dense. */
/* shard.cu: the delimiter counts of the halos and the run check of the left halo, into totals[16..18] (read back with the rest) */
int shard_aux_enqueue(const agb_desc &d, Workspace &W, const uint8_t *text, uint64_t n, const ShardInfo *sh, bool ordinals, cudaStream_t st);

static int stages_after_front(const agb_desc &d, Workspace &W, const void *d_text, uint64_t n, bool use_front, bool count_in_front,
                              int want, int want_level, agb_record *d_records, uint64_t capacity, cudaStream_t st, agb_result *res,
                              const ShardInfo *sh = nullptr)
{
	int rc;
	const uint64_t n_chunks = (n + 15) / 16;
	if (use_front && refine_cannot_thin(d)) { bool dense = false; rc = front_is_dense(W, n, st, &dense); if (rc) return rc; if (dense) use_front = false; }
	bool refined = false;
	if (use_front) { rc = refine_launch(d, W, d_text, n, st, &refined); if (rc) return rc; }
	for (int attempt = 0; ; attempt++) {
		rc = records_launch(d, W, d_text, n, use_front, refined, want, want_level, d_records, capacity, st, sh); if (rc) return rc;
		if (want & AGB_WANT_ORDINALS) { rc = ordinals_launch(d, W, d_text, n, (want & AGB_WANT_RECORDS) ? d_records : nullptr, capacity, st, count_in_front); if (rc) return rc; }
		CUDA_TRY(cudaEventRecord(W.e2, st));
		if (sh) { rc = shard_aux_enqueue(d, W, (const uint8_t *)d_text, n, sh, (want & AGB_WANT_ORDINALS) != 0, st); if (rc) return rc; }
		rc = fetch_result(W, want, capacity, use_front && refined, st, res); if (rc) return rc;
		if (!(use_front && refined)) break;
		const uint64_t ncand = W.h_totals[12];
		W.cand_hint = (size_t)ncand;
		if (ncand <= W.cand_cap || attempt) break;
		/* the list was too small: again, with the size known now (sparse) or over every byte (dense) */
		if (ncand > n_chunks / 20 + 1024) { use_front = false; refined = false; }
		CUDA_TRY(cudaMemsetAsync(W.totals, 0, 12 * sizeof(unsigned long long), st));
	}
	return AGB_OK;
}

/* -v, count only: can the answer be had as (records) - (matching records)?  Newline records, one part, no wildcards, an
 * anchor plan that stage 1.5 can verify (pattern.c keeps the anchors of a -v pattern in the descriptor) */
static bool complement_usable(const agb_desc &d)
{
	if (!d.inverse || d.plan != AGB_PLAN_ALL || d.n_anchors < 1 || d.n_anchors > 9 || d.n_anchors3 || !d.refine) return false;
	if (d.L != 1 || d.delim[0] != '\n' || d.user_delim || d.delim_fold[0] || d.and_mode || d.wildmask || d.init1 == ~0ull) return false;
	return true;
}

int scan_device_impl(const agb_desc &d_in, const void *d_text, uint64_t n, int want, int want_level,
                     agb_record *d_records, uint64_t capacity, cudaStream_t st, agb_result *res, const ShardInfo *sh)
{
	if (!res) return AGB_ERR_ARG;
	memset(res, 0, sizeof *res);
	if (((uintptr_t)d_text & 15) != 0) { snprintf(g_err, sizeof g_err, "text pointer must be 16-byte aligned"); return AGB_ERR_ARG; }
	if ((want & AGB_WANT_RECORDS) && capacity && !d_records) return AGB_ERR_ARG;
	int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) return AGB_ERR_ARG;
	if (want == AGB_WANT_COUNT && !sh && n >= (1u << 20) && complement_usable(d_in)) {
		/* `agrep -c -v pattern`, newline records: every record either matches or does not (the same test at the same close,
		 * bitap.c:182 with INVERSE flipped), so the count of the non-matching ones is the number of records minus the count of
		 * the matching ones -- and those the anchors find.  Records: one per newline of the text, one more for an unterminated
		 * last line (the delimiter appended at EOF closes it; after a final newline it would close the phantom record that
		 * agrep.c:3811 drops).  The newlines are counted by the same pass (j at EOF = newlines + appended + virtual). */
		agb_desc pos = d_in; pos.inverse = 0; pos.plan = AGB_PLAN_ANCHORS;
		agb_result r;
		int rc = scan_device_impl(pos, d_text, n, AGB_WANT_COUNT | AGB_WANT_ORDINALS, -1, nullptr, 0, st, &r, nullptr); if (rc) return rc;
		unsigned char last = 0;
		CUDA_TRY(cudaMemcpyAsync(&last, (const uint8_t *)d_text + n - 1, 1, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		const uint64_t records = r.n_closes - 2 + (last != '\n' ? 1 : 0);
		if (r.n_closes < 2 || r.n_matched > records) { snprintf(g_err, sizeof g_err, "internal: complement count out of range"); return AGB_ERR_CUDA; }
		*res = r;
		res->n_matched = records - r.n_matched; res->n_closes = 0; res->n_records = 0;
		return AGB_OK;
	}
	std::lock_guard<std::mutex> lk(g_ws_mu[dev]);
	Workspace &W = g_ws[dev];
	int rc = ws_prepare(W, n); if (rc) return rc;
	agb_desc planned;
	rc = adaptive_plan(d_in, W, d_text, n, st, &planned); if (rc) return rc;
	const agb_desc &d = planned;
	if (want == AGB_WANT_COUNT && !sh && n >= (1u << 20) && exact_count_usable(d)) {
		/* `agrep -c the`: an exact literal no longer than its anchor needs no automaton (front.cu, exact_count_launch) */
		CUDA_TRY(cudaMemsetAsync(W.totals, 0, 16 * sizeof(unsigned long long), st));
		CUDA_TRY(cudaEventRecord(W.e0, st));
		rc = exact_count_launch(d, W, d_text, n, st); if (rc) return rc;
		CUDA_TRY(cudaEventRecord(W.e1, st));
		CUDA_TRY(cudaEventRecord(W.e2, st));
		rc = fetch_result(W, want, capacity, false, st, res); if (rc) return rc;
		res->n_flagged = 0;
		CUDA_TRY(cudaEventElapsedTime(&res->ms_front, W.e0, W.e1));
		res->ms_records = 0;
		return AGB_OK;
	}
	rc = ws_upload_desc(W, d, st); if (rc) return rc;
	CUDA_TRY(cudaMemsetAsync(W.totals, 0, 16 * sizeof(unsigned long long), st));
	CUDA_TRY(cudaEventRecord(W.e0, st));
	bool use_front = front_usable(d) && n > 0;
	/* -n with a 1-byte delimiter: stage 1 reads every byte anyway and counts the delimiters of each 512-byte block */
	const bool count_in_front = use_front && (want & AGB_WANT_ORDINALS) && d.L == 1;
	if (count_in_front) { rc = ordinals_prepare_blocks(d, W, n, st); if (rc) return rc; }
	if (use_front) { rc = front_launch(d, W, d_text, n, 0, ~0ull, false, st, count_in_front); if (rc) return rc; }
	CUDA_TRY(cudaEventRecord(W.e1, st));
	rc = stages_after_front(d, W, d_text, n, use_front, count_in_front, want, want_level, d_records, capacity, st, res, sh); if (rc) return rc;
	if (sh && W.h_totals[11]) { snprintf(g_err, sizeof g_err, "a record of this shard runs past its halo (%d bytes behind the shard)", AGB_HALO_RIGHT); return AGB_ERR_ARG; }
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
 * pinned ring filled by 4 host threads when it is pageable; pread(2) by 4 threads straight into the pinned ring when the
 * source is a regular file -- while stage 1 runs on the slice that arrived before (its last chunk looks 4
 * bytes into the next one), so the scan hides behind PCIe; stages 1.5 and 2 run once over the whole bitmap. */
struct SliceSource {
	const uint8_t *mem;      /* host memory source, or NULL */
	bool pinned;             /* mem is page-locked: copy from it directly */
	int fd;                  /* file descriptor source when mem == NULL (a regular file) */
	off_t fd_off = 0;        /* where the text starts in it */
};

/* a slice of a regular file into the pinned ring: 4 host threads pread(2) a quarter each (one thread's read(2) from the page
 * cache is a third of what PCIe takes).  direct: fd was opened with O_DIRECT -- whole 4 KiB blocks are asked for (the ring's
 * buffers are page aligned and a multiple of 4 KiB long; the file's last block comes back short) */
static bool par_pread(int fd, off_t at, uint8_t *dst, size_t len, bool direct)
{
	const int T = 4; const size_t part = ((len + T - 1) / T + 4095) & ~(size_t)4095;
	std::thread th[T]; int used = 0; std::atomic<int> bad{0};
	for (int t = 0; t < T; t++) {
		size_t a = (size_t)t * part; if (a >= len) break;
		size_t l = std::min(part, len - a);
		th[used++] = std::thread([=, &bad] {
			size_t got = 0;
			while (got < l) {
				const size_t ask = direct ? ((l - got + 4095) & ~(size_t)4095) : l - got;
				ssize_t r = pread(fd, dst + a + got, ask, at + (off_t)(a + got));
				if (r <= 0 || (direct && (size_t)r < l - got && ((size_t)r & 4095))) { bad = 1; return; }
				got += std::min((size_t)r, l - got);
			}
		});
	}
	for (int t = 0; t < used; t++) th[t].join();
	return bad == 0;
}

/* AGB_ODIRECT=1: read regular files past the page cache (a second descriptor on the same file, through /proc/self/fd);
 * -1 when not asked for, when the text does not start on a block boundary, or when the file system refuses */
static int open_direct(int fd, off_t fd_off)
{
	const char *e = getenv("AGB_ODIRECT");
	if (!e || !*e || *e == '0' || (fd_off & 4095)) return -1;
	char path[64]; snprintf(path, sizeof path, "/proc/self/fd/%d", fd);
	return open(path, O_RDONLY | O_DIRECT);
}
struct FdGuard { int fd = -1; ~FdGuard() { if (fd >= 0) close(fd); } };
/* one slice of the file into dst; falls back to the caller's own descriptor for good if the direct one fails */
static bool read_slice(const SliceSource &src, int *dfd, uint64_t off, uint8_t *dst, size_t len)
{
	if (*dfd >= 0) {
		if (par_pread(*dfd, src.fd_off + (off_t)off, dst, len, true)) return true;
		close(*dfd); *dfd = -1;
	}
	return par_pread(src.fd, src.fd_off + (off_t)off, dst, len, false);
}

static int scan_stream_impl(const agb_desc &d, uint64_t n, const SliceSource &src, int want,
                            agb_record *records, uint64_t capacity, agb_result *res)
{
	memset(res, 0, sizeof *res);
	int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) return AGB_ERR_ARG;
	std::lock_guard<std::mutex> lk(g_ws_mu[dev]);
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
	FdGuard dg; if (!src.mem && src.fd >= 0) dg.fd = open_direct(src.fd, src.fd_off);
	int &dfd = dg.fd;
	if (!direct && n && !W.stage[0]) for (int i = 0; i < STAGE_BUFS; i++) CUDA_TRY(cudaMallocHost(&W.stage[i], H2D_SLICE));
	const bool use_front = front_usable(d) && n > 0;
	const bool count_in_front = use_front && (want & AGB_WANT_ORDINALS) && d.L == 1;
	if (count_in_front) { rc = ordinals_prepare_blocks(d, W, n, W.s_comp); if (rc) return rc; }
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
				/* fill_buf(): the slice from the file */
				if (!read_slice(src, &dfd, off, W.stage[sb], (size_t)len)) { snprintf(g_err, sizeof g_err, "pread(2) failed or hit the end of the file in [%llu, %llu) of %llu", (unsigned long long)off, (unsigned long long)(off + len), (unsigned long long)n); return AGB_ERR_ARG; }
			}
			CUDA_TRY(cudaMemcpyAsync(W.h2d_text + off, W.stage[sb], len, cudaMemcpyHostToDevice, W.s_copy));
		}
		CUDA_TRY(cudaEventRecord(W.ev_copy[sb], W.s_copy));
		/* stage 1 on the previous slice: its last chunk looks 4 bytes into this one, which is now on its way */
		if (use_front) {
			CUDA_TRY(cudaStreamWaitEvent(W.s_comp, W.ev_copy[sb], 0));
			if (i > 0) { rc = front_launch(d, W, W.h2d_text, n, (i - 1) * words_per_slice, i * words_per_slice, true, W.s_comp, count_in_front); if (rc) return rc; }
		}
	}
	if (n_slices) {
		CUDA_TRY(cudaStreamWaitEvent(W.s_comp, W.ev_copy[(n_slices - 1) % STAGE_BUFS], 0));
		if (use_front) { rc = front_launch(d, W, W.h2d_text, n, (n_slices - 1) * words_per_slice, ~0ull, true, W.s_comp, count_in_front); if (rc) return rc; }
	}
	CUDA_TRY(cudaEventRecord(W.e1, W.s_comp));
	rc = stages_after_front(d, W, W.h2d_text, n, use_front, count_in_front, want, -1, W.h2d_rec, capacity, W.s_comp, res); if (rc) return rc;
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
		SliceSource src; src.mem = nullptr; src.pinned = false; src.fd = fd; src.fd_off = cur >= 0 ? cur : 0;
		int rc = scan_stream_impl(p->d, n, src, want, records, capacity, res);
		if (cur >= 0) lseek(fd, cur + (off_t)n, SEEK_SET);            /* as read(2) would have left it */
		return rc;
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

/* ---- a text kept in HBM across scans (the drop-in layer's exec() scans the same file K + 2 times under -B,
 * agrep.c:3582-3728: one upload instead of K + 2) ---- */
struct agb_text { uint8_t *d; uint64_t n; int dev; };

static int text_upload(const SliceSource &src, uint64_t n, agb_text **out)
{
	int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) return AGB_ERR_ARG;
	agb_text *t = new agb_text; t->d = nullptr; t->n = n; t->dev = dev;
	const size_t need = (size_t)((n + 15) / 16 * 16 + 4096);
	if (cudaMalloc(&t->d, need) != cudaSuccess) { delete t; snprintf(g_err, sizeof g_err, "cudaMalloc of %zu bytes for the text failed", need); cudaGetLastError(); return AGB_ERR_NOMEM; }
	std::lock_guard<std::mutex> lk(g_ws_mu[dev]);
	Workspace &W = g_ws[dev];
	int rc = ws_prepare(W, 0); if (rc) { cudaFree(t->d); delete t; return rc; }
	if (!W.s_copy) {
		CUDA_TRY(cudaStreamCreateWithFlags(&W.s_copy, cudaStreamNonBlocking));
		CUDA_TRY(cudaStreamCreateWithFlags(&W.s_comp, cudaStreamNonBlocking));
		for (int i = 0; i < STAGE_BUFS; i++) CUDA_TRY(cudaEventCreateWithFlags(&W.ev_copy[i], cudaEventDisableTiming));
	}
	const bool direct = src.mem && src.pinned;
	FdGuard dg; if (!src.mem && src.fd >= 0) dg.fd = open_direct(src.fd, src.fd_off);
	int &dfd = dg.fd;
	if (!direct && n && !W.stage[0]) for (int i = 0; i < STAGE_BUFS; i++) CUDA_TRY(cudaMallocHost(&W.stage[i], H2D_SLICE));
	CUDA_TRY(cudaMemsetAsync(t->d + (n & ~(uint64_t)15), 0, need - (n & ~(uint64_t)15), W.s_copy));
	const uint64_t n_slices = (n + H2D_SLICE - 1) / H2D_SLICE;
	for (uint64_t i = 0; i < n_slices; i++) {
		const uint64_t off = i * H2D_SLICE, len = std::min<uint64_t>(H2D_SLICE, n - off);
		const int sb = (int)(i % STAGE_BUFS);
		if (direct) CUDA_TRY(cudaMemcpyAsync(t->d + off, src.mem + off, len, cudaMemcpyHostToDevice, W.s_copy));
		else {
			if (i >= STAGE_BUFS) CUDA_TRY(cudaEventSynchronize(W.ev_copy[sb]));
			if (src.mem) par_memcpy(W.stage[sb], src.mem + off, len);
			else {
				if (!read_slice(src, &dfd, off, W.stage[sb], (size_t)len)) { snprintf(g_err, sizeof g_err, "pread(2) failed or hit the end of the file in [%llu, %llu) of %llu", (unsigned long long)off, (unsigned long long)(off + len), (unsigned long long)n); cudaStreamSynchronize(W.s_copy); cudaFree(t->d); delete t; return AGB_ERR_ARG; }
			}
			CUDA_TRY(cudaMemcpyAsync(t->d + off, W.stage[sb], len, cudaMemcpyHostToDevice, W.s_copy));
		}
		CUDA_TRY(cudaEventRecord(W.ev_copy[sb], W.s_copy));
	}
	CUDA_TRY(cudaStreamSynchronize(W.s_copy));
	*out = t;
	return AGB_OK;
}

extern "C" int agb_text_from_host(const void *h_text, uint64_t n, agb_text **out)
{
	if (!out || (!h_text && n)) return AGB_ERR_ARG;
	SliceSource src; src.mem = (const uint8_t *)h_text; src.fd = -1; src.pinned = false;
	if (n) {
		cudaPointerAttributes attr; memset(&attr, 0, sizeof attr);
		src.pinned = cudaPointerGetAttributes(&attr, h_text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
		cudaGetLastError();
	}
	return text_upload(src, n, out);
}

extern "C" int agb_text_from_fd(int fd, agb_text **out)
{
	if (!out) return AGB_ERR_ARG;
	struct stat sb;
	if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { snprintf(g_err, sizeof g_err, "agb_text_from_fd needs a regular file"); return AGB_ERR_ARG; }
	off_t cur = lseek(fd, 0, SEEK_CUR);
	const uint64_t n = (cur >= 0 && sb.st_size > cur) ? (uint64_t)(sb.st_size - cur) : 0;
	SliceSource src; src.mem = nullptr; src.pinned = false; src.fd = fd; src.fd_off = cur >= 0 ? cur : 0;
	int rc = text_upload(src, n, out);
	if (cur >= 0) lseek(fd, cur + (off_t)n, SEEK_SET);                /* as read(2) would have left it */
	return rc;
}

extern "C" void agb_text_free(agb_text *t) { if (t) { cudaFree(t->d); delete t; } }
extern "C" uint64_t agb_text_size(const agb_text *t) { return t ? t->n : 0; }
extern "C" const void *agb_text_device(const agb_text *t) { return t ? t->d : nullptr; }

/* device scan of a resident text with the record list delivered to host memory */
static int scan_text_impl(const agb_desc &d, const agb_text *t, int want, int want_level, agb_record *records, uint64_t capacity, agb_result *res)
{
	if (!t || !res) return AGB_ERR_ARG;
	if ((want & AGB_WANT_RECORDS) && capacity && !records) return AGB_ERR_ARG;
	CUDA_TRY(cudaSetDevice(t->dev));
	agb_record *d_rec = nullptr;
	{
		std::lock_guard<std::mutex> lk(g_ws_mu[t->dev]);
		Workspace &W = g_ws[t->dev];
		if ((want & AGB_WANT_RECORDS) && capacity > W.h2d_rec_cap) {
			if (W.h2d_rec) cudaFree(W.h2d_rec);
			W.h2d_rec = nullptr; W.h2d_rec_cap = 0;
			CUDA_TRY(cudaMalloc(&W.h2d_rec, capacity * sizeof(agb_record))); W.h2d_rec_cap = capacity;
		}
		d_rec = W.h2d_rec;
	}
	int rc = scan_device_impl(d, t->d, t->n, want, want_level, d_rec, capacity, nullptr, res); if (rc) return rc;
	if (res->n_records) CUDA_TRY(cudaMemcpy(records, d_rec, res->n_records * sizeof(agb_record), cudaMemcpyDeviceToHost));
	return AGB_OK;
}

extern "C" int agb_scan_text(const agb_pattern *p, const agb_text *t, int want, agb_record *records, uint64_t capacity, agb_result *res)
{
	if (!p) return AGB_ERR_ARG;
	return scan_text_impl(p->d, t, want, -1, records, capacity, res);
}

/* keep the records of one level (stable, in place: the output index never overtakes the input index) */
__global__ void __launch_bounds__(1024) k_filter_level(agb_record *recs, uint64_t n, int level, unsigned long long *n_out)
{
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_base;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) s_base = 0;
	__syncthreads();
	for (uint64_t t0 = 0; t0 < n; t0 += 1024) {
		const uint64_t i = t0 + tid;
		agb_record r; r.level = -1;
		if (i < n) r = recs[i];
		const bool keep = i < n && r.level == level;
		const uint32_t m = __ballot_sync(0xffffffffu, keep);
		if (lane == 0) s_warp[wid] = __popc(m);
		__syncthreads();                                    /* every read of this tile is done */
		unsigned long long before = s_base;
		for (uint32_t w = 0; w < wid; w++) before += s_warp[w];
		if (keep) recs[before + __popc(m & ((1u << lane) - 1u))] = r;
		__syncthreads();
		if (tid == 0) { unsigned long long t = 0; for (int w = 0; w < 32; w++) t += s_warp[w]; s_base += t; }
		__syncthreads();
	}
	if (tid == 0) *n_out = s_base;
}

/* agrep.c:3582-3728: when the exact pass finds nothing, -B looks for the smallest D in 1..min(M-1,8) with a match,
 * rescanning every file once per D, and then once more at that D to print.  The rows are nested (A_j contains A_{j-1},
 * asearch.c:98-114), so ONE pass at a level k yields every record's smallest level <= k: the histogram tells the best
 * level, the list -- filtered to that level on the device -- is what the printing pass would print.  The pass runs at
 * k = 2 first (the anchor filter is still selective there; it also answers the exact question), then 4, then 8: one
 * pass for every best level up to 2, at most three.
 * best_k: smallest level with a match (-1: none up to min(M-1, 8)); res->n_matched: records at that level (what
 * the reference reports as "N words match within K errors"); d_records/capacity: their ordered list (level filled). */
extern "C" int agb_bestmatch_device(const char *pattern, const agb_options *opt, const void *d_text, uint64_t n,
                                    void *stream, agb_record *d_records, uint64_t capacity, int *best_k, agb_result *res,
                                    char *err, size_t errlen)
{
	if (!pattern || !opt || !best_k || !res) return AGB_ERR_ARG;
	if (capacity && !d_records) return AGB_ERR_ARG;
	agb_options o = *opt; agb_desc d; int m = (int)strlen(pattern);
	cudaStream_t st = (cudaStream_t)stream;
	o.bestmatch = 1;
	*best_k = -1;
	/* D < M of the exact pattern (agrep.c:3594); M there counts the delimiter and separator too */
	o.k = 0;
	int rc = agbi_build(pattern, &o, &d, err, errlen); if (rc) return rc;
	int kmax = d.M - 1; if (kmax > AGB_MAXERR) kmax = AGB_MAXERR; if (kmax > m - 1) kmax = m - 1;
	const int want = AGB_WANT_LEVELS | (capacity ? AGB_WANT_RECORDS : AGB_WANT_COUNT);
	int stages[3] = { 2, 4, 8 }, prev = -1;
	for (int si = 0; si < 3; si++) {
		int k = stages[si] < kmax ? stages[si] : kmax;
		if (k <= prev) break;
		o.k = k;
		rc = agbi_build(pattern, &o, &d, err, errlen); if (rc) return rc;
		rc = scan_device_impl(d, d_text, n, want, -1, d_records, capacity, st, res);
		if (rc) return rc;
		int best = -1;
		for (int l = prev + 1; l <= k; l++) if (res->level_hist[l]) { best = l; break; }    /* levels <= prev were already known to be empty */
		prev = k;
		if (best < 0) continue;
		*best_k = best;
		const uint64_t n_best = res->level_hist[best];
		if (capacity) {
			if (res->truncated) {
				/* the list of all levels up to k did not fit: once more, at the best level only */
				o.k = best;
				rc = agbi_build(pattern, &o, &d, err, errlen); if (rc) return rc;
				agb_result r2;
				rc = scan_device_impl(d, d_text, n, want, best, d_records, capacity, st, &r2); if (rc) return rc;
				res->n_records = r2.n_records; res->truncated = r2.truncated;
			} else if (best < k && res->n_records) {
				int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
				std::lock_guard<std::mutex> lk(g_ws_mu[dev]);
				Workspace &W = g_ws[dev];
				k_filter_level<<<1, 1024, 0, st>>>(d_records, res->n_records, best, W.totals + 15); g_launches++;
				CUDA_TRY(cudaGetLastError());
				CUDA_TRY(cudaStreamSynchronize(st));
				res->n_records = std::min<uint64_t>(n_best, capacity);
			}
		}
		res->n_matched = n_best;
		return AGB_OK;
	}
	res->n_matched = 0; res->n_records = 0;
	return AGB_OK;
}
