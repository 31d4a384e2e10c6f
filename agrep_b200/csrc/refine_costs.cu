/* agrep_b200/csrc/refine_costs.cu -- instantiations of stage 1.5 (refine_kernel.cuh) */
#include "refine_kernel.cuh"

int refine_launch_costs(bool narrow, int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: if (narrow) launch_refine_one<uint32_t, 1, true>(P, grid, st); else launch_refine_one<uint64_t, 1, true>(P, grid, st); break;
	case 2: if (narrow) launch_refine_one<uint32_t, 2, true>(P, grid, st); else launch_refine_one<uint64_t, 2, true>(P, grid, st); break;
	case 3: if (narrow) launch_refine_one<uint32_t, 3, true>(P, grid, st); else launch_refine_one<uint64_t, 3, true>(P, grid, st); break;
	case 4: if (narrow) launch_refine_one<uint32_t, 4, true>(P, grid, st); else launch_refine_one<uint64_t, 4, true>(P, grid, st); break;
	case 5: if (narrow) launch_refine_one<uint32_t, 5, true>(P, grid, st); else launch_refine_one<uint64_t, 5, true>(P, grid, st); break;
	case 6: if (narrow) launch_refine_one<uint32_t, 6, true>(P, grid, st); else launch_refine_one<uint64_t, 6, true>(P, grid, st); break;
	case 7: if (narrow) launch_refine_one<uint32_t, 7, true>(P, grid, st); else launch_refine_one<uint64_t, 7, true>(P, grid, st); break;
	case 8: if (narrow) launch_refine_one<uint32_t, 8, true>(P, grid, st); else launch_refine_one<uint64_t, 8, true>(P, grid, st); break;
	case 9: if (narrow) launch_refine_one<uint32_t, 9, true>(P, grid, st); else launch_refine_one<uint64_t, 9, true>(P, grid, st); break;
	default: return -1;
	}
	return 0;
}
