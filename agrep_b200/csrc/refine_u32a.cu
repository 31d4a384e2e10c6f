/* agrep_b200/csrc/refine_u32a.cu -- instantiations of stage 1.5 (refine_kernel.cuh) */
#include "refine_kernel.cuh"

int refine_launch_u32a(int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st)
{
	switch (nrows) {
	case 1: launch_refine_one<uint32_t, 1, false>(P, grid, st); break;
	case 2: launch_refine_one<uint32_t, 2, false>(P, grid, st); break;
	case 3: launch_refine_one<uint32_t, 3, false>(P, grid, st); break;
	case 4: launch_refine_one<uint32_t, 4, false>(P, grid, st); break;
	default: return -1;
	}
	return 0;
}
