// Per-element rotation / camera / kinematic-chain math shared by the device kernels and by the
// host-side math checker (csrc/hostmath.cu).  Everything is __host__ __device__ so the hand-derived
// adjoints can be validated against torch autograd on a CPU-only machine.
//
// Reference semantics restated here (all citations into the reference repo):
//   rot6d_*        utils/geometry.py:47-61   (rot6d_to_rotmat, F.normalize eps 1e-12)
//   quat_rodrigues utils/geometry.py:9-45    (batch_rodrigues via quaternion, angle = ||theta+1e-8||)
//   smplx_rodrigues  smplx/lbs.py::batch_rodrigues (third-party)
//   r2aa_*         utils/geometry.py:184-306 (rotation_matrix_to_angle_axis, 4-branch masked quaternion)
//   project_*      base_adaptor.py:160-170 + utils/geometry.py:63-91
//   chain_*        smplx/lbs.py::batch_rigid_transform (third-party)
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define DBOA_HD __host__ __device__ __forceinline__
#else
#define DBOA_HD inline
#endif

namespace dboa {

// ------------------------------------------------------------------------------------------------
// small 3-vector helpers
// ------------------------------------------------------------------------------------------------
DBOA_HD float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DBOA_HD void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ------------------------------------------------------------------------------------------------
// 6D -> rotation matrix (row-major R[3][3]; columns are b1, b2, b3)
// x is the 6-vector viewed as (3,2): a1 = (x0,x2,x4), a2 = (x1,x3,x5).
// ------------------------------------------------------------------------------------------------
DBOA_HD void rot6d_fwd(const float* x, float* R) {
    const float eps = 1e-12f;
    float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
    float n1 = fmaxf(sqrtf(dot3(a1, a1)), eps);
    float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    float d = dot3(b1, a2);
    float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    float n2 = fmaxf(sqrtf(dot3(u, u)), eps);
    float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    float b3[3];
    cross3(b1, b2, b3);
    for (int i = 0; i < 3; ++i) { R[i * 3 + 0] = b1[i]; R[i * 3 + 1] = b2[i]; R[i * 3 + 2] = b3[i]; }
}

DBOA_HD void rot6d_bwd(const float* x, const float* dR, float* dx) {
    const float eps = 1e-12f;
    float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
    float r1 = sqrtf(dot3(a1, a1)), n1 = fmaxf(r1, eps);
    float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    float d = dot3(b1, a2);
    float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    float r2 = sqrtf(dot3(u, u)), n2 = fmaxf(r2, eps);
    float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    float db1[3], db2[3], db3[3], t[3];
    for (int i = 0; i < 3; ++i) { db1[i] = dR[i * 3 + 0]; db2[i] = dR[i * 3 + 1]; db3[i] = dR[i * 3 + 2]; }
    // b3 = b1 x b2
    cross3(b2, db3, t); for (int i = 0; i < 3; ++i) db1[i] += t[i];
    cross3(db3, b1, t); for (int i = 0; i < 3; ++i) db2[i] += t[i];
    // b2 = u / max(|u|, eps)
    float du[3];
    if (r2 > eps) { float s = dot3(b2, db2); for (int i = 0; i < 3; ++i) du[i] = (db2[i] - b2[i] * s) / n2; }
    else          { for (int i = 0; i < 3; ++i) du[i] = db2[i] / n2; }
    // u = a2 - d b1 ; d = b1 . a2
    float da2[3] = {du[0], du[1], du[2]};
    float dd = -dot3(du, b1);
    for (int i = 0; i < 3; ++i) { db1[i] += -d * du[i] + dd * a2[i]; da2[i] += dd * b1[i]; }
    // b1 = a1 / max(|a1|, eps)
    float da1[3];
    if (r1 > eps) { float s = dot3(b1, db1); for (int i = 0; i < 3; ++i) da1[i] = (db1[i] - b1[i] * s) / n1; }
    else          { for (int i = 0; i < 3; ++i) da1[i] = db1[i] / n1; }
    dx[0] = da1[0]; dx[2] = da1[1]; dx[4] = da1[2];
    dx[1] = da2[0]; dx[3] = da2[1]; dx[5] = da2[2];
}

// ------------------------------------------------------------------------------------------------
// axis-angle -> R, the reference's quaternion route
// ------------------------------------------------------------------------------------------------
DBOA_HD void quat_rodrigues(const float* th, float* R) {
    float e0 = th[0] + 1e-8f, e1 = th[1] + 1e-8f, e2 = th[2] + 1e-8f;
    float angle = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    float nx = th[0] / angle, ny = th[1] / angle, nz = th[2] / angle;
    float half = angle * 0.5f;
    float c = cosf(half), s = sinf(half);
    float q[4] = {c, s * nx, s * ny, s * nz};
    float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] / qn, x = q[1] / qn, y = q[2] / qn, z = q[3] / qn;
    float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;     R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;     R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;
}

// axis-angle -> R, smplx's Rodrigues formula  R = I + sin(a) K + (1 - cos(a)) K^2
DBOA_HD void smplx_rodrigues(const float* th, float* R) {
    float e0 = th[0] + 1e-8f, e1 = th[1] + 1e-8f, e2 = th[2] + 1e-8f;
    float angle = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    float kx = th[0] / angle, ky = th[1] / angle, kz = th[2] / angle;
    float s = sinf(angle), c1 = 1.0f - cosf(angle);
    float K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float kk = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
            R[i * 3 + j] = (i == j ? 1.0f : 0.0f) + s * K[i * 3 + j] + c1 * kk;
        }
}

// ------------------------------------------------------------------------------------------------
// rotation matrix -> axis-angle (reference's kornia-derived 4-branch formulation)
// ------------------------------------------------------------------------------------------------
struct R2AA {       // intermediate values shared by forward and backward
    int branch;     // 0..3
    float t;        // selected trace-like term
    float qraw[4];  // selected un-normalised quaternion
    float q[4];     // 0.5 * qraw / sqrt(t)
};

DBOA_HD void r2aa_quat(const float* R, R2AA& s) {
    const float r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7], r22 = R[8];
    bool d2 = r22 < 1e-6f, d0_d1 = r00 > r11, d0_nd1 = r00 < -r11;
    if (d2 && d0_d1) {
        s.branch = 0; s.t = 1 + r00 - r11 - r22;
        s.qraw[0] = r21 - r12; s.qraw[1] = s.t; s.qraw[2] = r10 + r01; s.qraw[3] = r02 + r20;
    } else if (d2) {
        s.branch = 1; s.t = 1 - r00 + r11 - r22;
        s.qraw[0] = r02 - r20; s.qraw[1] = r10 + r01; s.qraw[2] = s.t; s.qraw[3] = r21 + r12;
    } else if (d0_nd1) {
        s.branch = 2; s.t = 1 - r00 - r11 + r22;
        s.qraw[0] = r10 - r01; s.qraw[1] = r02 + r20; s.qraw[2] = r21 + r12; s.qraw[3] = s.t;
    } else {
        s.branch = 3; s.t = 1 + r00 + r11 + r22;
        s.qraw[0] = s.t; s.qraw[1] = r21 - r12; s.qraw[2] = r02 - r20; s.qraw[3] = r10 - r01;
    }
    float rs = sqrtf(s.t);
    for (int i = 0; i < 4; ++i) s.q[i] = s.qraw[i] / rs * 0.5f;
}

DBOA_HD void r2aa_fwd(const float* R, float* aa) {
    R2AA s;
    r2aa_quat(R, s);
    float q1 = s.q[1], q2 = s.q[2], q3 = s.q[3];
    float s2 = q1 * q1 + q2 * q2 + q3 * q3;
    float sn = sqrtf(s2), cs = s.q[0];
    float two_theta = 2.0f * (cs < 0.0f ? atan2f(-sn, -cs) : atan2f(sn, cs));
    float k = s2 > 0.0f ? two_theta / sn : 2.0f;
    aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
    for (int i = 0; i < 3; ++i) if (aa[i] != aa[i]) aa[i] = 0.0f;      // geometry.py:212
}

// dR += J^T daa.  At sin^2 == 0 torch autograd yields NaN (0 * inf through sqrt); we return the
// gradient of the selected k = 2 branch instead (documented deviation, DESIGN.md "quirks").
DBOA_HD void r2aa_bwd(const float* R, const float* daa_in, float* dR) {
    R2AA s;
    r2aa_quat(R, s);
    float q1 = s.q[1], q2 = s.q[2], q3 = s.q[3];
    float s2 = q1 * q1 + q2 * q2 + q3 * q3;
    float sn = sqrtf(s2), cs = s.q[0];
    float two_theta = 2.0f * (cs < 0.0f ? atan2f(-sn, -cs) : atan2f(sn, cs));
    float k = s2 > 0.0f ? two_theta / sn : 2.0f;
    float daa[3] = {daa_in[0], daa_in[1], daa_in[2]};
    float aa[3] = {q1 * k, q2 * k, q3 * k};
    for (int i = 0; i < 3; ++i) if (aa[i] != aa[i]) daa[i] = 0.0f;
    float dq[4] = {0.f, daa[0] * k, daa[1] * k, daa[2] * k};
    if (s2 > 0.0f) {
        float dk = daa[0] * q1 + daa[1] * q2 + daa[2] * q3;
        float dtt = dk / sn;
        float dsn = -dk * two_theta / (sn * sn);
        float den = sn * sn + cs * cs;
        dsn += dtt * (2.0f * cs / den);
        float dcs = dtt * (-2.0f * sn / den);
        float ds2 = dsn / (2.0f * sn);
        dq[1] += 2.0f * q1 * ds2; dq[2] += 2.0f * q2 * ds2; dq[3] += 2.0f * q3 * ds2;
        dq[0] += dcs;
    }
    // q = 0.5 * qraw / sqrt(t)
    float rs = sqrtf(s.t);
    float dqraw[4], dt = 0.f;
    for (int i = 0; i < 4; ++i) {
        dqraw[i] = 0.5f * dq[i] / rs;
        dt += dq[i] * (-0.25f * s.qraw[i] / (s.t * rs));
    }
    // entries: d(rAB) accumulators
    float g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, g20 = 0, g21 = 0, g22 = 0;
    switch (s.branch) {
    case 0:
        dt += dqraw[1];
        g21 += dqraw[0]; g12 -= dqraw[0]; g10 += dqraw[2]; g01 += dqraw[2]; g02 += dqraw[3]; g20 += dqraw[3];
        g00 += dt; g11 -= dt; g22 -= dt; break;
    case 1:
        dt += dqraw[2];
        g02 += dqraw[0]; g20 -= dqraw[0]; g10 += dqraw[1]; g01 += dqraw[1]; g21 += dqraw[3]; g12 += dqraw[3];
        g00 -= dt; g11 += dt; g22 -= dt; break;
    case 2:
        dt += dqraw[3];
        g10 += dqraw[0]; g01 -= dqraw[0]; g02 += dqraw[1]; g20 += dqraw[1]; g21 += dqraw[2]; g12 += dqraw[2];
        g00 -= dt; g11 -= dt; g22 += dt; break;
    default:
        dt += dqraw[0];
        g21 += dqraw[1]; g12 -= dqraw[1]; g02 += dqraw[2]; g20 -= dqraw[2]; g10 += dqraw[3]; g01 -= dqraw[3];
        g00 += dt; g11 += dt; g22 += dt; break;
    }
    dR[0] += g00; dR[1] += g01; dR[2] += g02; dR[3] += g10; dR[4] += g11; dR[5] += g12;
    dR[6] += g20; dR[7] += g21; dR[8] += g22;
}

// ------------------------------------------------------------------------------------------------
// weak-perspective projection of one joint: cam = (s, tx, ty) -> t = (tx, ty, 2 f / (res s + 1e-9))
// p = f * (X + t).xy / (X + t).z / (res / 2)
// ------------------------------------------------------------------------------------------------
#define DBOA_FOCAL 5000.0f
#define DBOA_RES 224.0f

DBOA_HD void project_fwd(const float* cam, const float* X, float* p) {
    float tz = 2.0f * DBOA_FOCAL / (DBOA_RES * cam[0] + 1e-9f);
    float px = X[0] + cam[1], py = X[1] + cam[2], pz = X[2] + tz;
    // reference divides all three coordinates by z, multiplies by K, then scales by 1/(res/2)
    p[0] = (DBOA_FOCAL * (px / pz)) / (DBOA_RES * 0.5f);
    p[1] = (DBOA_FOCAL * (py / pz)) / (DBOA_RES * 0.5f);
}

// accumulates dX[3] and dcam[3]
DBOA_HD void project_bwd(const float* cam, const float* X, const float* dp, float* dX, float* dcam) {
    float den = DBOA_RES * cam[0] + 1e-9f;
    float tz = 2.0f * DBOA_FOCAL / den;
    float px = X[0] + cam[1], py = X[1] + cam[2], pz = X[2] + tz;
    const float sc = DBOA_FOCAL / (DBOA_RES * 0.5f);
    float gx = dp[0] * sc / pz, gy = dp[1] * sc / pz;
    float gz = -(dp[0] * sc * px + dp[1] * sc * py) / (pz * pz);
    dX[0] += gx; dX[1] += gy; dX[2] += gz;
    dcam[1] += gx; dcam[2] += gy;
    dcam[0] += gz * (-2.0f * DBOA_FOCAL * DBOA_RES / (den * den));
}

// ------------------------------------------------------------------------------------------------
// SMPL kinematic chain for one body (24 joints, parents[j] < j).
//   inputs : R[24][9] local rotations, J[24][3] rest joints
//   outputs: Gr[24][9], Gt[24][3] world transforms (Gt == posed joints), A[24][12] skinning
//            transforms with the rest joint removed: A_rot = Gr, A_t = Gt - Gr J
// ------------------------------------------------------------------------------------------------
DBOA_HD void mat3_mul(const float* a, const float* b, float* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
}
DBOA_HD void mat3_vec(const float* a, const float* v, float* o) {
    for (int i = 0; i < 3; ++i) o[i] = a[i * 3 + 0] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
}
DBOA_HD void mat3T_vec(const float* a, const float* v, float* o) {
    for (int i = 0; i < 3; ++i) o[i] = a[0 * 3 + i] * v[0] + a[1 * 3 + i] * v[1] + a[2 * 3 + i] * v[2];
}

DBOA_HD void chain_fwd(const float* R, const float* J, const int* parents, float* Gr, float* Gt, float* A) {
    for (int j = 0; j < 24; ++j) {
        if (j == 0) {
            for (int i = 0; i < 9; ++i) Gr[i] = R[i];
            for (int i = 0; i < 3; ++i) Gt[i] = J[i];
        } else {
            int p = parents[j];
            float rel[3] = {J[j * 3 + 0] - J[p * 3 + 0], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
            mat3_mul(Gr + p * 9, R + j * 9, Gr + j * 9);
            float t[3];
            mat3_vec(Gr + p * 9, rel, t);
            for (int i = 0; i < 3; ++i) Gt[j * 3 + i] = t[i] + Gt[p * 3 + i];
        }
        float gj[3];
        mat3_vec(Gr + j * 9, J + j * 3, gj);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) A[j * 12 + r * 4 + c] = Gr[j * 9 + r * 3 + c];
            A[j * 12 + r * 4 + 3] = Gt[j * 3 + r] - gj[r];
        }
    }
}

// Inputs: dA[24][12] (grad of skinning transforms), dJtr[24][3] (grad of posed joints).
// Outputs (overwritten): dR[24][9], dJ[24][3].   dGr/dGt are caller-provided scratch [24][9]/[24][3].
DBOA_HD void chain_bwd(const float* R, const float* J, const int* parents, const float* Gr,
                       const float* dA, const float* dJtr, float* dGr, float* dGt, float* dR, float* dJ) {
    for (int j = 0; j < 24; ++j) {
        float dAt[3] = {dA[j * 12 + 3], dA[j * 12 + 7], dA[j * 12 + 11]};
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) dGr[j * 9 + r * 3 + c] = dA[j * 12 + r * 4 + c] - dAt[r] * J[j * 3 + c];
            dGt[j * 3 + r] = dAt[r] + dJtr[j * 3 + r];
        }
        float t[3];
        mat3T_vec(Gr + j * 9, dAt, t);
        for (int i = 0; i < 3; ++i) dJ[j * 3 + i] = -t[i];
    }
    for (int j = 23; j >= 1; --j) {
        int p = parents[j];
        const float* Grp = Gr + p * 9;
        // dR_j = Grp^T dGr_j
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                dR[j * 9 + r * 3 + c] = Grp[0 * 3 + r] * dGr[j * 9 + 0 * 3 + c] + Grp[1 * 3 + r] * dGr[j * 9 + 1 * 3 + c] +
                                        Grp[2 * 3 + r] * dGr[j * 9 + 2 * 3 + c];
        float rel[3] = {J[j * 3 + 0] - J[p * 3 + 0], J[j * 3 + 1] - J[p * 3 + 1], J[j * 3 + 2] - J[p * 3 + 2]};
        // dGr_p += dGr_j R_j^T + dGt_j (x) rel
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                dGr[p * 9 + r * 3 + c] += dGr[j * 9 + r * 3 + 0] * R[j * 9 + c * 3 + 0] + dGr[j * 9 + r * 3 + 1] * R[j * 9 + c * 3 + 1] +
                                          dGr[j * 9 + r * 3 + 2] * R[j * 9 + c * 3 + 2] + dGt[j * 3 + r] * rel[c];
        float drel[3];
        mat3T_vec(Grp, dGt + j * 3, drel);
        for (int i = 0; i < 3; ++i) {
            dGt[p * 3 + i] += dGt[j * 3 + i];
            dJ[j * 3 + i] += drel[i];
            dJ[p * 3 + i] -= drel[i];
        }
    }
    for (int i = 0; i < 9; ++i) dR[i] = dGr[i];
    for (int i = 0; i < 3; ++i) dJ[i] += dGt[i];
}

}  // namespace dboa
