/* all_dbg.cu -- debug build (make dbg): the three translation units as one, so that the stamp buffer of ric_fast.cuh
 * (HBF_TIMING) is a single set of device globals */
#include "ric_kernels.cu"
#include "ipm_kernels.cu"
#include "blk_kernels.cu"
#include "tree_ipm_kernels.cu"
