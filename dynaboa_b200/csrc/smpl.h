// Internal SMPL workspace carving (device pointers) shared by smpl.cu and the C-ABI wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/dynaboa_b200.h"

namespace dboa {

constexpr size_t SMPL_NV3 = 6890 * 3;
constexpr size_t SMPL_SKIN_CTAS = (6890 + 127) / 128;   // 54
constexpr size_t SMPL_NSPLIT = 7;                       // row splits of the blend-shape product (7 x 31 = 217 rows)

// saved by the forward for the backward: per body (SMPL_NSPLIT + 1) * 20670 + 648 floats
struct SmplTape {
    float *partial, *vposed, *A, *Gr, *J, *Jtr;
    SmplTape(float* base, int B) {
        partial = base;                       // [SMPL_NSPLIT][B][20670]
        vposed = partial + SMPL_NSPLIT * (size_t)B * SMPL_NV3;
        A = vposed + (size_t)B * SMPL_NV3;    // [B][24][12]
        Gr = A + (size_t)B * 288;             // [B][24][9]
        J = Gr + (size_t)B * 216;             // [B][24][3] rest joints
        Jtr = J + (size_t)B * 72;             // [B][24][3] posed joints
    }
    static size_t floats(int B) { return (size_t)B * ((SMPL_NSPLIT + 1) * SMPL_NV3 + 288 + 216 + 72 + 72); }
};

struct SmplScratch {
    float *dverts, *dvposed, *dA_part, *dJtr, *dc;
    SmplScratch(float* base, int B) {
        dverts = base;
        dvposed = dverts + (size_t)B * SMPL_NV3;
        dA_part = dvposed + (size_t)B * SMPL_NV3;          // [B][54][288]
        dJtr = dA_part + (size_t)B * SMPL_SKIN_CTAS * 288;
        dc = dJtr + (size_t)B * 72;                        // [B][217]
    }
    static size_t floats(int B) { return (size_t)B * (2 * SMPL_NV3 + SMPL_SKIN_CTAS * 288 + 72 + 217 + 7); }
};

int smpl_forward(const dboa_smpl_model& m, const float* betas, const float* rot, int B, float* verts, float* joints, float* tape,
                 cudaStream_t st);
int smpl_backward(const dboa_smpl_model& m, const float* rot, int B, const float* tape, const float* djoints, float* scratch,
                  float* drot, float* dbetas, int accumulate, cudaStream_t st);
int rodrigues_launch(const float* aa, float* R, int n, int kind, cudaStream_t st);

}  // namespace dboa
