// Host build of clean_pvnet_b200/csrc/p3p_core.cuh for the CPU test-suite (tests/test_p3p_host_core.py).
// Test infrastructure only -- nothing in the product links or loads this.
#include "../clean_pvnet_b200/csrc/p3p_core.cuh"

extern "C" int p3p_host_solve4(const double *pts3d /*[4][3]*/, const double *pts2d /*[4][2]*/, const double *K /*[3][3]*/,
                               double *rt /*[6]*/)
{
    double X[4][3], x2[4][2];
    for (int i = 0; i < 4; ++i) {
        for (int r = 0; r < 3; ++r) X[i][r] = pts3d[i * 3 + r];
        x2[i][0] = pts2d[i * 2]; x2[i][1] = pts2d[i * 2 + 1];
    }
    const double cam[4] = { K[0], K[4], K[2], K[5] };
    return pvb::p3p_solve4(X, x2, cam, rt);
}

extern "C" int p3p_host_quartic(const double *k, double *roots) { return pvb::p3p_quartic_real_roots(k, roots); }

extern "C" void p3p_host_select4(const double *w, int pn, int *idx) { pvb::p3p_select4(w, pn, idx); }
