/*
 * csolve_accessors.c -- getters for the two options the reference keeps in statics of src/csolve.c
 * (`_time_max`, set by timeout_init() for -t, src/csolve.c:191-193; `_workers_max`, set by shared_init() for -j,
 * src/csolve.c:86-88). This translation unit INCLUDES the unmodified csolve.c (build it with -I<reference>/src and
 * -Dsolve=solve_cpu, exactly like the plain csolve.c of INTEGRATION.md, and link it INSTEAD of csolve.o): nothing of
 * the reference is edited, its solve() stays available as solve_cpu(), and the drop-in solve() of
 * csolve_gpu_shim.c can honour `-t` and map `-j N` to N GPUs.
 */
#include "csolve.c"

uint32_t csolve_shim_time_max(void) { return _time_max; }
uint32_t csolve_shim_workers_max(void) { return _workers_max; }
