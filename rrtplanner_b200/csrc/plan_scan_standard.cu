// plan_scan_kernel<RRTK_STANDARD, K, T> instantiations (see plan_scan.cuh)
#define RRTK_SCAN_KIND RRTK_STANDARD
#define RRTK_SCAN_FN scan_launch_standard
#define RRTK_SCAN_OCC_FN scan_occupancy_standard
#include "plan_scan_inst.cuh"
