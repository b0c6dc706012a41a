// SMPL body model: shape/pose blend, kinematic chain, linear-blend skinning, 49-joint extraction,
// and the hand-written backward of all of it (gradients w.r.t. betas and the 24 rotation matrices).
//
// Replaces smplx.lbs / SMPL.forward / VertexJointSelector as used through reference
// model/smpl.py:25-37.  Pure HBM-streaming fp32 work with
// warp/block reductions -- no tensor cores (not a dense contraction at batch <= 8).
//
// Data layout (device, fp32):
//   blend_dirs  [217][20670]  rows 0..9 = shapedirs transposed to (l, v*3+k), rows 10..216 = posedirs
//   J_template  [24][3], J_shapedirs [24][3][10]   (J_regressor folded through v_template / shapedirs)
//   lbs_weights [6890][24], J_extra [9][6890], joint_map[49] (into the 54-joint set), vertex_ids[21]
#include "common.cuh"
#include "kernels.h"
#include "rotmath.cuh"
#include "smpl.h"

namespace dboa {

constexpr int NV = 6890, NV3 = NV * 3, NROW = 217, NSPLIT = (int)SMPL_NSPLIT, ROWS_PER_SPLIT = (NROW + NSPLIT - 1) / NSPLIT;

// coefficient vector c[b] = [betas(10); (R[1:] - I)(207)]
__device__ __forceinline__ float blend_coeff(const float* betas, const float* rot, int b, int row) {
    if (row < 10) return betas[b * 10 + row];
    int p = row - 10, j = p / 9 + 1, rc = p - (j - 1) * 9;
    float v = rot[(size_t)b * 216 + j * 9 + rc];
    return (rc == 0 || rc == 4 || rc == 8) ? v - 1.0f : v;
}

// ---------------------------------------------------------------------------------------------
// blend forward: partial[s][b][idx] = (s==0 ? v_template[idx] : 0) + sum_{row in split s} c[b][row] D[row][idx]
// grid (ceil(20670/256), NSPLIT)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smpl_blend_fwd_kernel(const float* __restrict__ D, const float* __restrict__ vt,
                                                             const float* __restrict__ betas, const float* __restrict__ rot,
                                                             float* __restrict__ partial, int b0, int nb, int B) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sc[8][ROWS_PER_SPLIT];
    const int s = blockIdx.y, rbeg = s * ROWS_PER_SPLIT, rend = min(rbeg + ROWS_PER_SPLIT, NROW);
    for (int i = threadIdx.x; i < 8 * ROWS_PER_SPLIT; i += 256) {
        int b = i / ROWS_PER_SPLIT, r = i - b * ROWS_PER_SPLIT;
        sc[b][r] = (b < nb && rbeg + r < rend) ? blend_coeff(betas, rot, b0 + b, rbeg + r) : 0.f;
    }
    __syncthreads();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= NV3) return;
    float acc[8];
    const float base = (s == 0) ? vt[idx] : 0.f;
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = base;
#pragma unroll 8
    for (int r = rbeg; r < rend; ++r) {
        float d = __ldg(D + (size_t)r * NV3 + idx);
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[b] = fmaf(sc[b][r - rbeg], d, acc[b]);
    }
    for (int b = 0; b < nb; ++b) partial[((size_t)s * B + b0 + b) * NV3 + idx] = acc[b];
}

// ---------------------------------------------------------------------------------------------
// rest joints + kinematic chain: one 288-thread block per body (all loads of the set-up in flight at once; the chain
// itself is serial in thread 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288) smpl_chain_fwd_kernel(const float* __restrict__ Jt, const float* __restrict__ Js,
                                                            const int* __restrict__ parents, const float* __restrict__ betas,
                                                            const float* __restrict__ rot, float* __restrict__ A_out,
                                                            float* __restrict__ Gr_out, float* __restrict__ J_out,
                                                            float* __restrict__ Jtr_out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sR[216], sJ[72], sGr[216], sGt[72], sA[288];
    __shared__ int sp[24];
    const int b = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < 216; i += 288) sR[i] = rot[(size_t)b * 216 + i];
    for (int i = t; i < 72; i += 288) {
        float v = Jt[i];
#pragma unroll
        for (int l = 0; l < 10; ++l) v = fmaf(Js[i * 10 + l], betas[b * 10 + l], v);
        sJ[i] = v;
    }
    if (t < 24) sp[t] = parents[t];
    __syncthreads();
    if (t == 0) chain_fwd(sR, sJ, sp, sGr, sGt, sA);
    __syncthreads();
    for (int i = t; i < 288; i += 288) A_out[(size_t)b * 288 + i] = sA[i];
    for (int i = t; i < 216; i += 288) Gr_out[(size_t)b * 216 + i] = sGr[i];
    for (int i = t; i < 72; i += 288) { J_out[(size_t)b * 72 + i] = sJ[i]; Jtr_out[(size_t)b * 72 + i] = sGt[i]; }
}

// ---------------------------------------------------------------------------------------------
// skinning: thread per vertex; grid (ceil(6890/128), B)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) smpl_skin_fwd_kernel(const float* __restrict__ partial, const float* __restrict__ A,
                                                            const float* __restrict__ W, float* __restrict__ vposed,
                                                            float* __restrict__ verts, int B) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sA[288];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 288; i += 128) sA[i] = A[(size_t)b * 288 + i];
    __syncthreads();
    const int v = blockIdx.x * 128 + threadIdx.x;
    if (v >= NV) return;
    float vp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float s = 0.f;
        for (int z = 0; z < NSPLIT; ++z) s += partial[((size_t)z * B + b) * NV3 + v * 3 + k];
        vp[k] = s;
        vposed[(size_t)b * NV3 + v * 3 + k] = s;
    }
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        float4 w4 = ldg4(W + (size_t)v * 24 + q * 4);
        float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float* Aj = sA + (q * 4 + e) * 12;
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = fmaf(wv[e], Aj[i], T[i]);
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        verts[(size_t)b * NV3 + v * 3 + r] = T[r * 4 + 0] * vp[0] + T[r * 4 + 1] * vp[1] + T[r * 4 + 2] * vp[2] + T[r * 4 + 3];
}

// ---------------------------------------------------------------------------------------------
// 49 joints: grid (49, B), 512 threads
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) smpl_joints_fwd_kernel(const float* __restrict__ verts, const float* __restrict__ Jtr,
                                                              const float* __restrict__ Jx, const int* __restrict__ joint_map,
                                                              const int* __restrict__ vertex_ids, float* __restrict__ joints) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[32];
    const int i = blockIdx.x, b = blockIdx.y, src = joint_map[i];
    float* o = joints + ((size_t)b * 49 + i) * 3;
    if (src < 24) {
        if (threadIdx.x < 3) o[threadIdx.x] = Jtr[(size_t)b * 72 + src * 3 + threadIdx.x];
    } else if (src < 45) {
        int v = vertex_ids[src - 24];
        if (threadIdx.x < 3) o[threadIdx.x] = verts[(size_t)b * NV3 + v * 3 + threadIdx.x];
    } else {
        const float* jr = Jx + (size_t)(src - 45) * NV;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int v = threadIdx.x; v < NV; v += 512) {         // 14 steps, 4 in flight: the dense regressor rows are latency-bound
            float w = __ldg(jr + v);
            const float* p = verts + (size_t)b * NV3 + v * 3;
            s0 = fmaf(w, p[0], s0); s1 = fmaf(w, p[1], s1); s2 = fmaf(w, p[2], s2);
        }
        s0 = block_sum(s0, red); s1 = block_sum(s1, red); s2 = block_sum(s2, red);
        if (threadIdx.x == 0) { o[0] = s0; o[1] = s1; o[2] = s2; }
    }
}

int smpl_forward(const dboa_smpl_model& m, const float* betas, const float* rot, int B, float* verts, float* joints, float* tape,
                 cudaStream_t st) {
    SmplTape t(tape, B);
    for (int b0 = 0; b0 < B; b0 += 8) {
        int nb = B - b0 < 8 ? B - b0 : 8;
        dim3 g(ceil_div(NV3, 256), NSPLIT);
        DBOA_TRY(launch_ex(smpl_blend_fwd_kernel, dim3(g), dim3(256), 0, st, dim3(1, 1, 1), true, m.blend_dirs, m.v_template, betas, rot, t.partial, b0, nb, B));
    }
    DBOA_TRY(launch_ex(smpl_chain_fwd_kernel, dim3(B), dim3(288), 0, st, dim3(1, 1, 1), true, m.J_template, m.J_shapedirs, m.parents, betas, rot, t.A, t.Gr, t.J, t.Jtr));
    DBOA_TRY(launch_ex(smpl_skin_fwd_kernel, dim3(ceil_div(NV, 128), B), dim3(128), 0, st, dim3(1, 1, 1), true, t.partial, t.A, m.lbs_weights, t.vposed, verts, B));
    return launch_ex(smpl_joints_fwd_kernel, dim3(49, B), dim3(512), 0, st, dim3(1, 1, 1), true, verts, t.Jtr, m.J_extra, m.joint_map, m.vertex_ids, joints);
}

// =============================================================================================
// backward
// =============================================================================================
// d(joints49) -> dverts (dense through J_extra + vertex picks) and dJtr; grid (ceil(6890/256), B)
__global__ void __launch_bounds__(256) smpl_joints_bwd_kernel(const float* __restrict__ dj, const float* __restrict__ Jx,
                                                              const int* __restrict__ joint_map, const int* __restrict__ vertex_ids,
                                                              float* __restrict__ dverts, float* __restrict__ dJtr) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sde[9][3], sdp[21][3], sdk[24][3];
    __shared__ int svid[21];
    const int b = blockIdx.y;
    if (threadIdx.x < 54) {
        // gather the gradient of each of the 54 source joints from the (up to 2) outputs mapped to it
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        for (int i = 0; i < 49; ++i)
            if (joint_map[i] == (int)threadIdx.x) {
                const float* p = dj + ((size_t)b * 49 + i) * 3;
                g0 += p[0]; g1 += p[1]; g2 += p[2];
            }
        float* dst = threadIdx.x < 24 ? sdk[threadIdx.x] : (threadIdx.x < 45 ? sdp[threadIdx.x - 24] : sde[threadIdx.x - 45]);
        dst[0] = g0; dst[1] = g1; dst[2] = g2;
    }
    if (threadIdx.x >= 64 && threadIdx.x < 85) svid[threadIdx.x - 64] = vertex_ids[threadIdx.x - 64];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < 72) dJtr[(size_t)b * 72 + threadIdx.x] = sdk[threadIdx.x / 3][threadIdx.x % 3];
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= NV) return;
    float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        float w = __ldg(Jx + (size_t)e * NV + v);
        g[0] = fmaf(w, sde[e][0], g[0]); g[1] = fmaf(w, sde[e][1], g[1]); g[2] = fmaf(w, sde[e][2], g[2]);
    }
    for (int p = 0; p < 21; ++p)
        if (svid[p] == v) { g[0] += sdp[p][0]; g[1] += sdp[p][1]; g[2] += sdp[p][2]; }
    float* o = dverts + (size_t)b * NV3 + v * 3;
    o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
}

// skin backward: dvposed = T_rot^T dv; dA partial per CTA; grid (ceil(6890/128), B)
__global__ void __launch_bounds__(128) smpl_skin_bwd_kernel(const float* __restrict__ dverts, const float* __restrict__ vposed,
                                                            const float* __restrict__ A, const float* __restrict__ W,
                                                            float* __restrict__ dvposed, float* __restrict__ dA_part) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sA[288];
    __shared__ float sw[128][25];
    __shared__ float sdT[128][13];
    const int b = blockIdx.y, t = threadIdx.x;
    for (int i = t; i < 288; i += 128) sA[i] = A[(size_t)b * 288 + i];
    __syncthreads();
    const int v = blockIdx.x * 128 + t;
    const bool valid = v < NV;
    float wv[24];
    float dv[3] = {0.f, 0.f, 0.f}, vp[3] = {0.f, 0.f, 0.f};
    if (valid) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            float4 w4 = ldg4(W + (size_t)v * 24 + q * 4);
            wv[q * 4 + 0] = w4.x; wv[q * 4 + 1] = w4.y; wv[q * 4 + 2] = w4.z; wv[q * 4 + 3] = w4.w;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { dv[k] = dverts[(size_t)b * NV3 + v * 3 + k]; vp[k] = vposed[(size_t)b * NV3 + v * 3 + k]; }
    } else {
#pragma unroll
        for (int j = 0; j < 24; ++j) wv[j] = 0.f;
    }
    // T_rot for dvposed
    float Tr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Tr[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 24; ++j)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Tr[r * 3 + c] = fmaf(wv[j], sA[j * 12 + r * 4 + c], Tr[r * 3 + c]);
    if (valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            dvposed[(size_t)b * NV3 + v * 3 + c] = Tr[0 * 3 + c] * dv[0] + Tr[1 * 3 + c] * dv[1] + Tr[2 * 3 + c] * dv[2];
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) sw[t][j] = wv[j];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        sdT[t][r * 4 + 0] = dv[r] * vp[0]; sdT[t][r * 4 + 1] = dv[r] * vp[1]; sdT[t][r * 4 + 2] = dv[r] * vp[2];
        sdT[t][r * 4 + 3] = dv[r];
    }
    __syncthreads();
    for (int e = t; e < 288; e += 128) {
        const int j = e / 12, c = e - j * 12;
        float s = 0.f;
        for (int u = 0; u < 128; ++u) s = fmaf(sw[u][j], sdT[u][c], s);
        dA_part[((size_t)b * gridDim.x + blockIdx.x) * 288 + e] = s;
    }
}

// blend backward: dc[b][row] = sum_idx D[row][idx] dvposed[b][idx]; grid (217), 1024 threads
__global__ void __launch_bounds__(1024) smpl_blend_bwd_kernel(const float* __restrict__ D, const float* __restrict__ dvposed,
                                                             float* __restrict__ dc, int b0, int nb) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[32];
    const int row = blockIdx.x;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int idx = threadIdx.x; idx < NV3; idx += 1024) {     // 21 steps, 4 in flight (was 81 dependent steps of 256 threads)
        float d = __ldg(D + (size_t)row * NV3 + idx);
#pragma unroll
        for (int b = 0; b < 8; ++b)
            if (b < nb) acc[b] = fmaf(d, dvposed[(size_t)(b0 + b) * NV3 + idx], acc[b]);
    }
    for (int b = 0; b < nb; ++b) {
        float s = block_sum(acc[b], red);
        if (threadIdx.x == 0) dc[(size_t)(b0 + b) * NROW + row] = s;
    }
}

// chain backward + assembly of the final gradients; one 288-thread block per body (one thread per entry of the 24 x 12
// joint-transform gradient while the 54 skinning partials are summed; the chain itself is serial in thread 0)
__global__ void __launch_bounds__(288) smpl_chain_bwd_kernel(const float* __restrict__ Js, const int* __restrict__ parents,
                                                            const float* __restrict__ rot, const float* __restrict__ J,
                                                            const float* __restrict__ Gr, const float* __restrict__ dA_part, int nparts,
                                                            const float* __restrict__ dJtr, const float* __restrict__ dc,
                                                            float* __restrict__ drot, float* __restrict__ dbetas, int accumulate) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sR[216], sJ[72], sGr[216], sdA[288], sdJt[72], sdGr[216], sdGt[72], sdR[216], sdJ[72];
    __shared__ int sp[24];
    const int b = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < 216; i += 288) { sR[i] = rot[(size_t)b * 216 + i]; sGr[i] = Gr[(size_t)b * 216 + i]; }
    for (int i = t; i < 72; i += 288) { sJ[i] = J[(size_t)b * 72 + i]; sdJt[i] = dJtr[(size_t)b * 72 + i]; }
    for (int i = t; i < 288; i += 288) {
        float s = 0.f;
#pragma unroll 6
        for (int p = 0; p < nparts; ++p) s += dA_part[((size_t)b * nparts + p) * 288 + i];
        sdA[i] = s;
    }
    if (t < 24) sp[t] = parents[t];
    __syncthreads();
    if (t == 0) chain_bwd(sR, sJ, sp, sGr, sdA, sdJt, sdGr, sdGt, sdR, sdJ);
    __syncthreads();
    for (int i = t; i < 216; i += 288) {
        float g = sdR[i];
        if (i >= 9) g += dc[(size_t)b * NROW + 10 + (i - 9)];       // pose-blend feature gradient
        size_t o = (size_t)b * 216 + i;
        drot[o] = accumulate ? drot[o] + g : g;
    }
    if (t < 10) {
        float g = dc[(size_t)b * NROW + t];
        for (int i = 0; i < 72; ++i) g = fmaf(Js[i * 10 + t], sdJ[i], g);
        size_t o = (size_t)b * 10 + t;
        dbetas[o] = accumulate ? dbetas[o] + g : g;
    }
}

int smpl_backward(const dboa_smpl_model& m, const float* rot, int B, const float* tape, const float* djoints, float* scratch,
                  float* drot, float* dbetas, int accumulate, cudaStream_t st) {
    SmplTape t(const_cast<float*>(tape), B);
    SmplScratch s(scratch, B);
    const int nparts = ceil_div(NV, 128);
    DBOA_TRY(launch_ex(smpl_joints_bwd_kernel, dim3(ceil_div(NV, 256), B), dim3(256), 0, st, dim3(1, 1, 1), true, djoints, m.J_extra, m.joint_map, m.vertex_ids, s.dverts, s.dJtr));
    DBOA_TRY(launch_ex(smpl_skin_bwd_kernel, dim3(nparts, B), dim3(128), 0, st, dim3(1, 1, 1), true, s.dverts, t.vposed, t.A, m.lbs_weights, s.dvposed, s.dA_part));
    for (int b0 = 0; b0 < B; b0 += 8) {
        int nb = B - b0 < 8 ? B - b0 : 8;
        DBOA_TRY(launch_ex(smpl_blend_bwd_kernel, dim3(NROW), dim3(1024), 0, st, dim3(1, 1, 1), true, m.blend_dirs, s.dvposed, s.dc, b0, nb));
    }
    return launch_ex(smpl_chain_bwd_kernel, dim3(B), dim3(288), 0, st, dim3(1, 1, 1), true, m.J_shapedirs, m.parents, rot, t.J, t.Gr, s.dA_part, nparts, s.dJtr, s.dc, drot, dbetas, accumulate);
}

// axis-angle -> rotation matrix, kind 0 = reference quaternion route, 1 = smplx Rodrigues formula
__global__ void rodrigues_kernel(const float* __restrict__ aa, float* __restrict__ R, int n, int kind) {
    pdl_wait();
    pdl_trigger();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float th[3] = {aa[(size_t)i * 3], aa[(size_t)i * 3 + 1], aa[(size_t)i * 3 + 2]}, Ri[9];
    if (kind == 0) quat_rodrigues(th, Ri); else smplx_rodrigues(th, Ri);
    for (int k = 0; k < 9; ++k) R[(size_t)i * 9 + k] = Ri[k];
}
int rodrigues_launch(const float* aa, float* R, int n, int kind, cudaStream_t st) {
    return launch_ex(rodrigues_kernel, dim3(ceil_div(n, 128)), dim3(128), 0, st, dim3(1, 1, 1), true, aa, R, n, kind);
}

}  // namespace dboa
