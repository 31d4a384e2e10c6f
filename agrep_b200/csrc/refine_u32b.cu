/* agrep_b200/csrc/refine_u32b.cu -- instantiations of stage 1.5 (refine_kernel.cuh) */
#include "refine_kernel.cuh"

int refine_launch_u32b(int nrows, const RefineParams &P, unsigned &grid, cudaStream_t st)
{
	switch (nrows) {
	case 5: launch_refine_one<uint32_t, 5, false>(P, grid, st); break;
	case 6: launch_refine_one<uint32_t, 6, false>(P, grid, st); break;
	case 7: launch_refine_one<uint32_t, 7, false>(P, grid, st); break;
	case 8: launch_refine_one<uint32_t, 8, false>(P, grid, st); break;
	case 9: launch_refine_one<uint32_t, 9, false>(P, grid, st); break;
	default: return -1;
	}
	return 0;
}
