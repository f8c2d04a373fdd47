// plan_scan_kernel<RRTK_INFORMED, K, T> instantiations (see plan_scan.cuh)
#define RRTK_SCAN_KIND RRTK_INFORMED
#define RRTK_SCAN_FN scan_launch_informed
#define RRTK_SCAN_OCC_FN scan_occupancy_informed
#include "plan_scan_inst.cuh"
