// The back-jumping instance of the general search kernel, k_search<false, true> with BJ = true
// (csolve_solve_options.backjump; conflict_backtrack, src/csolve.c:350-364): kernels.cu compiled a second time, up to the
// end of k_search, in a namespace of its own. A unit of its own because an additional instantiation in kernels.cu moves
// the inliner's choices in the instances that are already there, and those keep the code they were measured with.
#define CSOLVE_BJ_UNIT 1
#define CSOLVE_BJ 1
#define csolve_dev csolve_dev_bj
#include "kernels.cu"
