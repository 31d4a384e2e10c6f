// tools/front_bench.cu -- development microbenchmark (not product): inner-compare variants of the stage-1
// anchor kernel on synthetic text, to pick the instruction mix with measurements instead of guesses.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/front_bench tools/front_bench.cu
// run:   tools/front_bench [GiB]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct P { const uint4 *text; uint32_t *bitmap; uint64_t n_chunks, n_words; uint32_t anchor[8]; };

__device__ __forceinline__ uint4 ld16(const uint4 *p) {
	uint4 v; asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }

template <int V, int NA> __device__ __forceinline__ uint32_t win(uint32_t lo, uint32_t hi, const P &p, uint32_t acc)
{
	uint32_t w0 = lo, w1 = __funnelshift_r(lo, hi, 8), w2 = __funnelshift_r(lo, hi, 16), w3 = __funnelshift_r(lo, hi, 24);
#pragma unroll
	for (int a = 0; a < NA; a++) {
		uint32_t A = p.anchor[a];
		if (V == 0) { acc = __vimin3_u32(acc, w0 - A, w1 - A); acc = __vimin3_u32(acc, w2 - A, w3 - A); }
		if (V == 1) { acc |= (w0 == A) | (w1 == A) | (w2 == A) | (w3 == A); }
		if (V == 2) { acc = __vimin3_u32(acc, w0 ^ A, w1 ^ A); acc = __vimin3_u32(acc, w2 ^ A, w3 ^ A); }
		if (V == 3) {
			uint32_t d0, d1, d2, d3, nA = 0u - A;
			asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d0) : "r"(w0), "r"(nA));
			asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d1) : "r"(w1), "r"(nA));
			asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d2) : "r"(w2), "r"(nA));
			asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d3) : "r"(w3), "r"(nA));
			acc = __vimin3_u32(acc, d0, d1); acc = __vimin3_u32(acc, d2, d3);
		}
		if (V == 4) { acc = min(acc, min(min(w0 - A, w1 - A), min(w2 - A, w3 - A))); }
	}
	return acc;
}

// V==9: no compare at all (pure streaming read + ballot): the load-path ceiling of this kernel shape
template <int V, int NA, int U, int T>
__global__ void __launch_bounds__(T) k(const P p)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * T + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * T) >> 5;
	const uint64_t n_groups = (p.n_words + U - 1) / U;
	for (uint64_t g = warp; g < n_groups; g += nwarps) {
		const uint64_t w0 = g * U;
		uint4 v[U + 1];
#pragma unroll
		for (int u = 0; u < U; u++) { uint64_t c = (w0 + u) * 32 + lane; v[u] = (c < p.n_chunks) ? ld16(p.text + c) : make_uint4(0, 0, 0, 0); }
		{ uint64_t c = (w0 + U) * 32; uint32_t nx = 0; if (lane == 0 && c < p.n_chunks) nx = __ldg((const uint32_t *)(p.text + c)); v[U] = make_uint4(nx, 0, 0, 0); }
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (w0 + u >= p.n_words) break;
			uint32_t give = (lane == 0) ? v[u + 1].x : v[u].x;
			uint32_t x4 = __shfl_sync(0xffffffffu, give, (lane + 1) & 31);
			uint32_t x0 = v[u].x, x1 = v[u].y, x2 = v[u].z, x3 = v[u].w;
			bool flag;
			if (V == 9) flag = ((x0 ^ x1 ^ x2 ^ x3 ^ x4) == p.anchor[0]);
			else {
				uint32_t acc = (V == 1) ? 0u : 0xffffffffu;
				acc = win<V, NA>(x0, x1, p, acc); acc = win<V, NA>(x1, x2, p, acc); acc = win<V, NA>(x2, x3, p, acc); acc = win<V, NA>(x3, x4, p, acc);
				flag = (V == 1) ? (acc != 0) : (acc == 0);
			}
			uint32_t word = __ballot_sync(0xffffffffu, flag);
			if (lane == 0) p.bitmap[w0 + u] = word;
		}
	}
}

__global__ void fill(uint32_t *t, uint64_t nwords, uint32_t seed)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
	for (; i < nwords; i += stride) {
		uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
		uint32_t w = 0;
		for (int b = 0; b < 4; b++) { uint32_t r = (z >> (b * 8)) & 31; w |= (r < 26 ? 'a' + r : (r < 31 ? ' ' : '\n')) << (8 * b); }
		t[i] = w;
	}
}

template <int V, int NA, int U, int T>
static void run(const char *name, P p, uint64_t bytes, int sms, int bps)
{
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	int grid = sms * bps;
	std::vector<float> ms;
	for (int it = 0; it < 6; it++) {
		CK(cudaEventRecord(e0));
		k<V, NA, U, T><<<grid, T>>>(p);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float t; CK(cudaEventElapsedTime(&t, e0, e1)); if (it >= 2) ms.push_back(t);
	}
	CK(cudaGetLastError());
	std::sort(ms.begin(), ms.end());
	printf("%-28s NA=%d U=%d T=%d grid=%dx%d  %8.3f ms  %8.1f GB/s\n", name, NA, U, T, sms, bps, ms[0], bytes / ms[0] / 1e6);
	fflush(stdout);
}

int main(int argc, char **argv)
{
	double gib = argc > 1 ? atof(argv[1]) : 4.0;
	uint64_t bytes = (uint64_t)(gib * (1ull << 30)) & ~511ull;
	uint32_t *text, *bitmap; CK(cudaMalloc(&text, bytes + 64)); CK(cudaMalloc(&bitmap, bytes / 128 + 64));
	fill<<<148 * 8, 256>>>(text, bytes / 4, 1); CK(cudaDeviceSynchronize());
	cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
	int sms = pr.multiProcessorCount;
	printf("device %s, %d SMs, %.2f GiB text\n", pr.name, sms, gib);
	P p; p.text = (const uint4 *)text; p.bitmap = bitmap; p.n_chunks = bytes / 16; p.n_words = p.n_chunks / 32;
	const char *a[8] = { "beca", "use ", "each", "gove", "rnme", "ntal", "xyzw", "qqqq" };
	for (int i = 0; i < 8; i++) p.anchor[i] = *(const uint32_t *)a[i];
	run<9, 1, 4, 256>("stream-only", p, bytes, sms, 8);
	run<9, 1, 4, 256>("stream-only", p, bytes, sms, 4);
	run<9, 1, 8, 256>("stream-only", p, bytes, sms, 4);
	run<9, 1, 2, 256>("stream-only", p, bytes, sms, 8);
	run<9, 1, 4, 512>("stream-only", p, bytes, sms, 4);
	run<0, 1, 4, 256>("viaddmnmx", p, bytes, sms, 8);
	run<0, 2, 4, 256>("viaddmnmx", p, bytes, sms, 8);
	run<0, 3, 4, 256>("viaddmnmx", p, bytes, sms, 8);
	run<0, 4, 4, 256>("viaddmnmx", p, bytes, sms, 8);
	run<0, 5, 4, 256>("viaddmnmx", p, bytes, sms, 8);
	run<1, 1, 4, 256>("isetp-or", p, bytes, sms, 8);
	run<1, 2, 4, 256>("isetp-or", p, bytes, sms, 8);
	run<1, 3, 4, 256>("isetp-or", p, bytes, sms, 8);
	run<1, 4, 4, 256>("isetp-or", p, bytes, sms, 8);
	run<1, 5, 4, 256>("isetp-or", p, bytes, sms, 8);
	run<2, 3, 4, 256>("xor+vimin3", p, bytes, sms, 8);
	run<3, 1, 4, 256>("imad-sub+vimin3", p, bytes, sms, 8);
	run<3, 3, 4, 256>("imad-sub+vimin3", p, bytes, sms, 8);
	run<3, 5, 4, 256>("imad-sub+vimin3", p, bytes, sms, 8);
	run<4, 3, 4, 256>("sub+min tree", p, bytes, sms, 8);
	run<0, 3, 2, 256>("viaddmnmx", p, bytes, sms, 8);
	run<0, 3, 8, 256>("viaddmnmx", p, bytes, sms, 4);
	run<0, 3, 4, 128>("viaddmnmx", p, bytes, sms, 16);
	run<3, 3, 8, 256>("imad-sub+vimin3", p, bytes, sms, 4);
	run<3, 3, 2, 256>("imad-sub+vimin3", p, bytes, sms, 8);
	return 0;
}
