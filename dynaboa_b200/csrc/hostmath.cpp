// Host-side instantiation of the per-element math in rotmath.cuh, for CPU-only validation of the
// hand-derived adjoints against torch autograd (tests/test_hostmath.py).  Test infrastructure for the
// build container (no GPU there); the product never loads this library.
#include "rotmath.cuh"

extern "C" {

void hm_rot6d_fwd(const float* x, float* R, int n) { for (int i = 0; i < n; ++i) dboa::rot6d_fwd(x + 6 * i, R + 9 * i); }
void hm_rot6d_bwd(const float* x, const float* dR, float* dx, int n) {
    for (int i = 0; i < n; ++i) dboa::rot6d_bwd(x + 6 * i, dR + 9 * i, dx + 6 * i);
}
void hm_quat_rodrigues(const float* th, float* R, int n) { for (int i = 0; i < n; ++i) dboa::quat_rodrigues(th + 3 * i, R + 9 * i); }
void hm_smplx_rodrigues(const float* th, float* R, int n) { for (int i = 0; i < n; ++i) dboa::smplx_rodrigues(th + 3 * i, R + 9 * i); }
void hm_r2aa_fwd(const float* R, float* aa, int n) { for (int i = 0; i < n; ++i) dboa::r2aa_fwd(R + 9 * i, aa + 3 * i); }
void hm_r2aa_bwd(const float* R, const float* daa, float* dR, int n) {
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 9; ++k) dR[9 * i + k] = 0.f;
        dboa::r2aa_bwd(R + 9 * i, daa + 3 * i, dR + 9 * i);
    }
}
void hm_project_fwd(const float* cam, const float* X, float* p, int nb, int nj) {
    for (int b = 0; b < nb; ++b)
        for (int j = 0; j < nj; ++j) dboa::project_fwd(cam + 3 * b, X + (b * nj + j) * 3, p + (b * nj + j) * 2);
}
void hm_project_bwd(const float* cam, const float* X, const float* dp, float* dX, float* dcam, int nb, int nj) {
    for (int b = 0; b < nb; ++b) {
        for (int k = 0; k < 3; ++k) dcam[3 * b + k] = 0.f;
        for (int j = 0; j < nj; ++j) {
            float* dx = dX + (b * nj + j) * 3;
            dx[0] = dx[1] = dx[2] = 0.f;
            dboa::project_bwd(cam + 3 * b, X + (b * nj + j) * 3, dp + (b * nj + j) * 2, dx, dcam + 3 * b);
        }
    }
}
void hm_chain_fwd(const float* R, const float* J, const int* parents, float* Gr, float* Gt, float* A, int nb) {
    for (int b = 0; b < nb; ++b) dboa::chain_fwd(R + b * 216, J + b * 72, parents, Gr + b * 216, Gt + b * 72, A + b * 288);
}
void hm_chain_bwd(const float* R, const float* J, const int* parents, const float* Gr, const float* dA, const float* dJtr,
                  float* dR, float* dJ, int nb) {
    float dGr[216], dGt[72];
    for (int b = 0; b < nb; ++b)
        dboa::chain_bwd(R + b * 216, J + b * 72, parents, Gr + b * 216, dA + b * 288, dJtr + b * 72, dGr, dGt, dR + b * 216, dJ + b * 72);
}

}  // extern "C"
