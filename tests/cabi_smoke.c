/* Plain-C consumer of the ABI: proves the header is C, the library links without torch/python, and the
 * host-side planning entry points work without a GPU.  Built and run by tests/test_cabi_host.py. */
#include <stdio.h>
#include <string.h>
#include "pvnet_vote_b200.h"

int main(void)
{
    pvb_desc d;
    pvb_layout L;
    memset(&d, 0, sizeof d);
    d.B = 16; d.H = 480; d.W = 640; d.K = 9; d.hn = 512;
    d.inlier_thresh = 0.99f; d.min_num = 5; d.max_num = 30000;
    d.mask_dtype = PVB_MASK_I64; d.select_mode = PVB_SELECT_BYTE;
    if (pvb_version() != PVB_VERSION) return 1;
    if (pvb_workspace_layout(&d, &L) != PVB_OK) { printf("%s\n", pvb_last_error()); return 2; }
    if (pvb_workspace_bytes(&d) != L.total || L.total == 0) return 3;
    d.K = 0;
    if (pvb_workspace_layout(&d, &L) != PVB_ERR_INVALID || strlen(pvb_last_error()) == 0) return 4;
    /* NULL tensors are rejected before any CUDA call */
    d.K = 9;
    if (pvb_ransac_voting_v3(&d, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) != PVB_ERR_INVALID) return 5;
    printf("pvb %d: cfg2 workspace %zu bytes, capacity %d, sizeof(pvb_desc)=%zu\n", pvb_version(), pvb_workspace_bytes(&d),
           (int)L.capacity, sizeof(pvb_desc));
    return 0;
}
