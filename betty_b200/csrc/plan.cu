// placeholder until the second-order tape executor lands (next commit)
#include "bb_common.cuh"
#include "../../include/betty_b200.h"
struct bb_plan { int dummy; };
extern "C" {
int bb_node_bytes(void) { return 0; }
int bb_plan_create(const struct bb_node*, int, bb_plan**) { return BB_ERR_UNSUPPORTED; }
int bb_plan_destroy(bb_plan*) { return BB_ERR_UNSUPPORTED; }
int bb_plan_set_zero_regions(bb_plan*, int, void* const*, const int64_t*, int) { return BB_ERR_UNSUPPORTED; }
int bb_plan_run(bb_plan*, int, void*) { return BB_ERR_UNSUPPORTED; }
int bb_plan_launch_count(const bb_plan*, int) { return BB_ERR_UNSUPPORTED; }
int bb_plan_hvp(bb_plan*, void*) { return BB_ERR_UNSUPPORTED; }
int bb_plan_neumann_loop(bb_plan*, int, float, float*, float*, const float*, int64_t, int, void*) { return BB_ERR_UNSUPPORTED; }
int bb_plan_cg_loop(bb_plan*, int, float, float*, float*, float*, const float*, int64_t, void*, int, void*) { return BB_ERR_UNSUPPORTED; }
const char* bb_version(void) { return "betty_b200 0.1.0 (sm_100a)"; }
}
