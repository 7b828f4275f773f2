/* Compiled as plain C99 by tests/test_capi_host.py: the C ABI header must be usable from C, and the struct layouts
 * the ctypes binding mirrors (fast_go_icp_b200/capi.py) must match what a C compiler lays out. */
#include <fgoicp_c.h>
#include <stddef.h>
#include <stdio.h>

int main(void)
{
    printf("fgoicp_info %zu %zu %zu %zu\n", sizeof(fgoicp_info), offsetof(fgoicp_info, dims), offsetof(fgoicp_info, grid_bytes),
           offsetof(fgoicp_info, build_ms));
    printf("fgoicp_level_stats %zu %zu %zu %zu\n", sizeof(fgoicp_level_stats), offsetof(fgoicp_level_stats, n_icp),
           offsetof(fgoicp_level_stats, ms_bnb_ub), offsetof(fgoicp_level_stats, best_icp_index));
    printf("fgoicp_normalisation %zu %zu %zu %zu\n", sizeof(fgoicp_normalisation), offsetof(fgoicp_normalisation, scale),
           offsetof(fgoicp_normalisation, bbox_min), offsetof(fgoicp_normalisation, device_ms));
    return 0;
}
