// plan_scan_kernel<RRTK_STAR, K, T> instantiations (see plan_scan.cuh)
#define RRTK_SCAN_KIND RRTK_STAR
#define RRTK_SCAN_FN scan_launch_star
#define RRTK_SCAN_OCC_FN scan_occupancy_star
#include "plan_scan_inst.cuh"
