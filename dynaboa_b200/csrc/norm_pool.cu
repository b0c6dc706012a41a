// GroupNorm(4 groups) forward/backward, ReLU masks, max/avg pooling and the NCHW->NHWC input
// transpose, all on NHWC fp32 activations.
//
// Replaces the ATen kernels behind reference model/hmr.py:14-18 (gn_helper -> nn.GroupNorm(4, C),
// eps 1e-5, affine), :36/57-58 (ReLU, residual add), :75 (MaxPool2d(3,2,1)), :80 (AvgPool2d(7))
// -- SURVEY.md §2.1 K2/K3.  Statistics use a chunked two-pass (count, mean, M2) scheme merged
// with Chan's formula in a fixed order, so results are deterministic and robust to large means.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace dboa {

constexpr int GN_G = 4;
constexpr float GN_EPS = 1e-5f;
constexpr int GN_BUDGET = 8192;          // floats per statistics CTA (8 float4 per thread)
constexpr int GN_NT = 256;

static inline int gn_rows(int C) { int cg = C / GN_G; int r = GN_BUDGET / cg; return r < 1 ? 1 : r; }
int gn_chunks(int HW, int C) { return ceil_div(HW, gn_rows(C)); }
size_t gn_partial_floats(int B, int HW, int C) { return (size_t)B * GN_G * gn_chunks(HW, C) * 3; }
size_t gn_bwd_partial_floats(int B, int HW, int C) {
    size_t ch = gn_chunks(HW, C);            // legacy 3-pass layout; also >= the cluster plan's 2*B*16*C? no: take the max
    size_t legacy = (size_t)B * GN_G * ch * 2 + 2 * (size_t)B * ch * C;
    size_t cluster = 2 * (size_t)B * 16 * C;
    return legacy > cluster ? legacy : cluster;
}

// ---------------------------------------------------------------------------------------------
// statistics: grid (chunks, 4, B); each CTA reduces rows [chunk*R, ..) x channels of one group
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GN_NT) gn_stats_kernel(const float* __restrict__ y, float* __restrict__ partial, int HW, int C, int R) {
    __shared__ float red[32];
    const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z, chunks = gridDim.x;
    const int cg = C / GN_G, cg4 = cg / 4;
    const int r0 = chunk * R, rows = min(R, HW - r0);
    const int nvec = rows * cg4;
    const float* base = y + ((size_t)b * HW + r0) * C + g * cg;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            int row = idx / cg4, cv = idx - row * cg4;
            v[i] = ldg4(base + (size_t)row * C + cv * 4);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float cnt = (float)(rows * cg);
    const float mean = block_sum(s, red) / cnt;
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            float a = v[i].x - mean, c = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
            m2 += (a * a + c * c) + (e * e + f * f);
        }
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) {
        float* p = partial + ((size_t)(b * GN_G + g) * chunks + chunk) * 3;
        p[0] = cnt; p[1] = mean; p[2] = m2;
    }
}

__device__ __forceinline__ void gn_combine(const float* partial, int chunks, float& mean, float& rstd) {
    float n = 0.f, mu = 0.f, M2 = 0.f;
    for (int c = 0; c < chunks; ++c) {
        float nb = partial[c * 3 + 0], mb = partial[c * 3 + 1], Mb = partial[c * 3 + 2];
        float tot = n + nb, delta = mb - mu;
        mu += delta * (nb / tot);
        M2 += Mb + delta * delta * (n * nb / tot);
        n = tot;
    }
    mean = mu;
    rstd = 1.0f / sqrtf(M2 / n + GN_EPS);
}

// ---------------------------------------------------------------------------------------------
// apply: out = relu?( (y-mean)*rstd*gamma+beta [+ res] [+ second normalised tensor] )
// grid (ceil(HW*C/4 / 256 / 4), B): 4 float4 per thread
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GN_NT) gn_apply_kernel(const float* __restrict__ y, const float* __restrict__ partial,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ stats, const float* __restrict__ res,
                                                         const float* __restrict__ y2, const float* __restrict__ partial2,
                                                         const float* __restrict__ gamma2, const float* __restrict__ beta2,
                                                         float* __restrict__ stats2, float* __restrict__ out, int HW, int C,
                                                         int chunks, int relu) {
    __shared__ float sm[16];
    const int b = blockIdx.y;
    if (threadIdx.x < GN_G) {
        float m, r;
        gn_combine(partial + (size_t)(b * GN_G + threadIdx.x) * chunks * 3, chunks, m, r);
        sm[threadIdx.x * 2] = m; sm[threadIdx.x * 2 + 1] = r;
        if (blockIdx.x == 0) { stats[(b * GN_G + threadIdx.x) * 2] = m; stats[(b * GN_G + threadIdx.x) * 2 + 1] = r; }
    } else if (y2 != nullptr && threadIdx.x < 2 * GN_G) {
        int g = threadIdx.x - GN_G;
        float m, r;
        gn_combine(partial2 + (size_t)(b * GN_G + g) * chunks * 3, chunks, m, r);
        sm[8 + g * 2] = m; sm[8 + g * 2 + 1] = r;
        if (blockIdx.x == 0) { stats2[(b * GN_G + g) * 2] = m; stats2[(b * GN_G + g) * 2 + 1] = r; }
    }
    __syncthreads();
    const int cg = C / GN_G;
    const size_t n4 = (size_t)HW * C / 4;
    const size_t boff = (size_t)b * HW * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        size_t i4 = ((size_t)blockIdx.x * 4 + i) * GN_NT + threadIdx.x;
        if (i4 >= n4) break;
        size_t e = i4 * 4;
        int c = (int)(e % C), g = c / cg;
        float mean = sm[g * 2], rstd = sm[g * 2 + 1];
        float4 v = ldg4(y + boff + e), ga = ldg4(gamma + c), be = ldg4(beta + c);
        float4 o;
        o.x = (v.x - mean) * rstd * ga.x + be.x; o.y = (v.y - mean) * rstd * ga.y + be.y;
        o.z = (v.z - mean) * rstd * ga.z + be.z; o.w = (v.w - mean) * rstd * ga.w + be.w;
        if (res != nullptr) { float4 r = ldg4(res + boff + e); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
        if (y2 != nullptr) {
            float m2 = sm[8 + g * 2], r2 = sm[8 + g * 2 + 1];
            float4 w = ldg4(y2 + boff + e), g2 = ldg4(gamma2 + c), b2 = ldg4(beta2 + c);
            o.x += (w.x - m2) * r2 * g2.x + b2.x; o.y += (w.y - m2) * r2 * g2.y + b2.y;
            o.z += (w.z - m2) * r2 * g2.z + b2.z; o.w += (w.w - m2) * r2 * g2.w + b2.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(out + boff + e) = o;
    }
}

int gn_stats(const float* y, int B, int HW, int C, float* partial, cudaStream_t st) {
    if (C % 16 != 0 || 256 % (C / 16) != 0 || C / GN_G > GN_BUDGET) return DBOA_ERR_SHAPE;
    int R = gn_rows(C);
    dim3 grid(gn_chunks(HW, C), GN_G, B);
    gn_stats_kernel<<<grid, GN_NT, 0, st>>>(y, partial, HW, C, R);
    return check_launch();
}

int gn_apply(const float* y, const float* partial, const float* gamma, const float* beta, float* stats, const float* res,
             const float* y2, const float* partial2, const float* gamma2, const float* beta2, float* stats2, float* out, int B,
             int HW, int C, int relu, cudaStream_t st) {
    size_t n4 = (size_t)HW * C / 4;
    dim3 grid(ceil_div(n4, GN_NT * 4), B);
    gn_apply_kernel<<<grid, GN_NT, 0, st>>>(y, partial, gamma, beta, stats, res, y2, partial2, gamma2, beta2, stats2, out, HW, C,
                                            gn_chunks(HW, C), relu);
    return check_launch();
}

// ---------------------------------------------------------------------------------------------
// backward.  With g = dz*gamma, xh = (y-mean)*rstd and group size N:
//   dy = rstd * (g - (sum g)/N - xh * (sum g*xh)/N);  dgamma_c = sum dz*xh;  dbeta_c = sum dz
// pass1: per (chunk, group, sample) partial sums; pass2: elementwise; pass3: per-channel reduce.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GN_NT) gn_bwd_pass1_kernel(const float* __restrict__ dout, const float* __restrict__ mask_src,
                                                             const float* __restrict__ y, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, float* __restrict__ spart,
                                                             float* __restrict__ dgpart, float* __restrict__ dbpart, int HW,
                                                             int C, int R) {
    __shared__ float red[32];
    __shared__ __align__(16) float smg[GN_NT * 4], smb[GN_NT * 4];
    const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z, chunks = gridDim.x;
    const int cg = C / GN_G, cg4 = cg / 4;
    const int r0 = chunk * R, rows = min(R, HW - r0);
    const int nvec = rows * cg4;
    const size_t off = ((size_t)b * HW + r0) * C + g * cg;
    const float mean = stats[(b * GN_G + g) * 2], rstd = stats[(b * GN_G + g) * 2 + 1];
    const int cv = threadIdx.x % cg4;                 // constant per thread since 256 % cg4 == 0
    const float4 ga = ldg4(gamma + g * cg + cv * 4);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            int row = idx / cg4;
            size_t e = off + (size_t)row * C + cv * 4;
            float4 d = ldg4(dout + e), v = ldg4(y + e);
            if (mask_src != nullptr) {
                float4 m = ldg4(mask_src + e);
                d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f; d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
            }
            float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
            dg.x += d.x * xh.x; dg.y += d.y * xh.y; dg.z += d.z * xh.z; dg.w += d.w * xh.w;
            db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
            float gx = d.x * ga.x, gy = d.y * ga.y, gz = d.z * ga.z, gw = d.w * ga.w;
            s1 += (gx + gy) + (gz + gw);
            s2 += (gx * xh.x + gy * xh.y) + (gz * xh.z + gw * xh.w);
        }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        float* p = spart + ((size_t)(b * GN_G + g) * chunks + chunk) * 2;
        p[0] = s1; p[1] = s2;
    }
    *reinterpret_cast<float4*>(&smg[threadIdx.x * 4]) = dg;
    *reinterpret_cast<float4*>(&smb[threadIdx.x * 4]) = db;
    __syncthreads();
    for (int c = threadIdx.x; c < cg; c += GN_NT) {
        const int ccv = c >> 2, comp = c & 3, rp = GN_NT / cg4;
        float a = 0.f, bsum = 0.f;
        for (int j = 0; j < rp; ++j) { a += smg[(j * cg4 + ccv) * 4 + comp]; bsum += smb[(j * cg4 + ccv) * 4 + comp]; }
        size_t o = ((size_t)b * chunks + chunk) * C + g * cg + c;
        dgpart[o] = a; dbpart[o] = bsum;
    }
}

__global__ void __launch_bounds__(GN_NT) gn_bwd_pass2_kernel(const float* __restrict__ dout, const float* __restrict__ mask_src,
                                                             const float* __restrict__ y, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, const float* __restrict__ spart,
                                                             float* __restrict__ dy, int HW, int C, int chunks) {
    __shared__ float sm[8];
    const int b = blockIdx.y;
    if (threadIdx.x < GN_G) {
        const float* p = spart + (size_t)(b * GN_G + threadIdx.x) * chunks * 2;
        float a = 0.f, c = 0.f;
        for (int k = 0; k < chunks; ++k) { a += p[k * 2]; c += p[k * 2 + 1]; }
        sm[threadIdx.x * 2] = a; sm[threadIdx.x * 2 + 1] = c;
    }
    __syncthreads();
    const int cg = C / GN_G;
    const float invN = 1.0f / ((float)HW * (float)cg);
    const size_t n4 = (size_t)HW * C / 4, boff = (size_t)b * HW * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        size_t i4 = ((size_t)blockIdx.x * 4 + i) * GN_NT + threadIdx.x;
        if (i4 >= n4) break;
        size_t e = i4 * 4;
        int c = (int)(e % C), g = c / cg;
        float mean = stats[(b * GN_G + g) * 2], rstd = stats[(b * GN_G + g) * 2 + 1];
        float m1 = sm[g * 2] * invN, m2 = sm[g * 2 + 1] * invN;
        float4 d = ldg4(dout + boff + e), v = ldg4(y + boff + e), ga = ldg4(gamma + c);
        if (mask_src != nullptr) {
            float4 m = ldg4(mask_src + boff + e);
            d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f; d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
        }
        float4 o;
        o.x = rstd * (d.x * ga.x - m1 - (v.x - mean) * rstd * m2);
        o.y = rstd * (d.y * ga.y - m1 - (v.y - mean) * rstd * m2);
        o.z = rstd * (d.z * ga.z - m1 - (v.z - mean) * rstd * m2);
        o.w = rstd * (d.w * ga.w - m1 - (v.w - mean) * rstd * m2);
        *reinterpret_cast<float4*>(dy + boff + e) = o;
    }
}

__global__ void gn_bwd_param_kernel(const float* __restrict__ dgpart, const float* __restrict__ dbpart, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int nrows, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < nrows; ++r) { a += dgpart[(size_t)r * C + c]; b += dbpart[(size_t)r * C + c]; }
    dgamma[c] += a; dbeta[c] += b;
}

int gn_bwd(const float* dout, const float* mask_src, const float* y, const float* stats, const float* gamma, float* dy,
           float* dgamma, float* dbeta, float* partial, int B, int HW, int C, cudaStream_t st) {
    if (C % 16 != 0 || 256 % (C / 16) != 0 || C / GN_G > GN_BUDGET) return DBOA_ERR_SHAPE;
    const int chunks = gn_chunks(HW, C), R = gn_rows(C);
    float* spart = partial;
    float* dgpart = partial + (size_t)B * GN_G * chunks * 2;
    float* dbpart = dgpart + (size_t)B * chunks * C;
    dim3 g1(chunks, GN_G, B);
    gn_bwd_pass1_kernel<<<g1, GN_NT, 0, st>>>(dout, mask_src, y, stats, gamma, spart, dgpart, dbpart, HW, C, R);
    DBOA_TRY(check_launch());
    size_t n4 = (size_t)HW * C / 4;
    dim3 g2(ceil_div(n4, GN_NT * 4), B);
    gn_bwd_pass2_kernel<<<g2, GN_NT, 0, st>>>(dout, mask_src, y, stats, gamma, spart, dy, HW, C, chunks);
    DBOA_TRY(check_launch());
    gn_bwd_param_kernel<<<ceil_div(C, 256), 256, 0, st>>>(dgpart, dbpart, dgamma, dbeta, B * chunks, C);
    return check_launch();
}

// =============================================================================================
// Single-launch GroupNorm forward / backward on thread-block clusters.
//
// At batch 1 every layer is a few-microsecond problem, so the dependent phases inside and between
// kernels -- not bytes -- set the time (profiles/r01_summary.md).  One (sample, group) is handled by
// ONE cluster of <= 16 CTAs: each CTA keeps its slab of the group in registers (<= 13 float4 per
// thread), publishes its partial sums in its own shared memory, the cluster synchronises in
// hardware (barrier.cluster), every CTA reads all partials through distributed shared memory and
// combines them in chunk order (deterministic), then normalises its registers.  Every tensor is
// read exactly once; there are no global partials, atomics or second launches on the data path.
// =============================================================================================
constexpr int GN_VMAX = 13;                      // float4 per thread held in registers
constexpr int GN_MAXCL = 16;                     // CTAs per cluster (non-portable limit)

struct GnPlan { int chunks, rows; };
static GnPlan gn_plan(int HW, int C) {
    const int cg4 = C / GN_G / 4, cap = GN_NT * GN_VMAX;
    int chunks = ceil_div((long long)HW * cg4, cap);
    if (chunks < 1) chunks = 1;
    int rows = ceil_div(HW, chunks);
    while (rows * cg4 > cap) { ++chunks; rows = ceil_div(HW, chunks); }
    chunks = ceil_div(HW, rows);
    return {chunks, rows};
}

static unsigned* g_sync_base = nullptr;
enum { SYNC_GNB_PCNT = 0, SYNC_REGIONS };
unsigned* sync_words(int which) {
    (void)which;
    if (g_sync_base == nullptr) {
        void* p = nullptr;
        if (cudaMalloc(&p, 256 * sizeof(unsigned)) != cudaSuccess) return nullptr;
        if (cudaMemset(p, 0, 256 * sizeof(unsigned)) != cudaSuccess) return nullptr;
        g_sync_base = static_cast<unsigned*>(p);
    }
    return g_sync_base;
}

template <typename K, typename... Args>
static int launch_x_cluster(K kernel, dim3 grid, cudaStream_t st, Args... args) {
    if (grid.x > 8) {
        // non-portable cluster sizes (9..16) must be enabled per kernel FUNCTION (several kernels share this template's type)
        static const void* enabled[16];
        static int n_enabled = 0;
        bool seen = false;
        for (int i = 0; i < n_enabled; ++i) seen = seen || (enabled[i] == (const void*)kernel);
        if (!seen) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) { g_last_cuda_error = (int)e; return DBOA_ERR_CUDA; }
            if (n_enabled < 16) enabled[n_enabled++] = (const void*)kernel;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(GN_NT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = grid.x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    ++g_launch_count;
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return DBOA_ERR_CUDA; }
    return DBOA_OK;
}

// grid (chunks, 4, B), cluster (chunks, 1, 1)
__global__ void __launch_bounds__(GN_NT, 2) gn_fwd_fused_kernel(const float* __restrict__ y, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ res,
                                                             float* __restrict__ out, float* __restrict__ stats, int HW, int C, int R,
                                                             int relu) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[32];
    __shared__ float part[4];                 // this CTA's (count, mean, M2)
    __shared__ float all[GN_MAXCL][3];
    __shared__ float sm[2];
    const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z, chunks = gridDim.x, slot = b * GN_G + g;
    const int cg = C / GN_G, cg4 = cg / 4;
    const int r0 = chunk * R, rows = min(R, HW - r0);
    const int nvec = rows * cg4;
    const size_t base = ((size_t)b * HW + r0) * C + g * cg;
    // every independent global load is issued up front (input slab, residual slab, affine parameters) so that the
    // kernel pays ONE memory round trip before its reductions
    const int cvf = threadIdx.x % cg4;                     // 256 % cg4 == 0: the channel vector is fixed per thread
    const float4 ga = ldg4(gamma + g * cg + cvf * 4), be = ldg4(beta + g * cg + cvf * 4);
    float4 v[GN_VMAX], rr[GN_VMAX];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < GN_VMAX; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            int row = idx / cg4;
            const size_t e = base + (size_t)row * C + cvf * 4;
            v[i] = ldg4(y + e);
            if (res != nullptr) rr[i] = ldg4(res + e);
        }
    }
#pragma unroll
    for (int i = 0; i < GN_VMAX; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float cnt = (float)(rows * cg);
    const float cmean = block_sum(s, red) / cnt;
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < GN_VMAX; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            float a = v[i].x - cmean, c = v[i].y - cmean, e = v[i].z - cmean, f = v[i].w - cmean;
            m2 += (a * a + c * c) + (e * e + f * f);
        }
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) { part[0] = cnt; part[1] = cmean; part[2] = m2; }
    cluster.sync();
    if (threadIdx.x < chunks * 3) {
        const int c = threadIdx.x / 3, k = threadIdx.x - c * 3;
        all[c][k] = cluster.map_shared_rank(part, c)[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float n = 0.f, mu = 0.f, M2 = 0.f;
        for (int c = 0; c < chunks; ++c) {          // Chan's merge in chunk order
            float nb = all[c][0], mb = all[c][1], Mb = all[c][2];
            float tot = n + nb, delta = mb - mu;
            mu += delta * (nb / tot);
            M2 += Mb + delta * delta * (n * nb / tot);
            n = tot;
        }
        sm[0] = mu; sm[1] = 1.0f / sqrtf(M2 / n + GN_EPS);
        if (chunk == 0) { stats[slot * 2] = sm[0]; stats[slot * 2 + 1] = sm[1]; }
    }
    __syncthreads();
    const float mean = sm[0], rstd = sm[1];
#pragma unroll
    for (int i = 0; i < GN_VMAX; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            int row = idx / cg4;
            const size_t e = base + (size_t)row * C + cvf * 4;
            float4 o;
            o.x = (v[i].x - mean) * rstd * ga.x + be.x; o.y = (v[i].y - mean) * rstd * ga.y + be.y;
            o.z = (v[i].z - mean) * rstd * ga.z + be.z; o.w = (v[i].w - mean) * rstd * ga.w + be.w;
            if (res != nullptr) { o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w; }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(out + e) = o;
        }
    }
    cluster.sync();                                        // peers may still be reading `part`
}

int gn_fwd_fused(const float* y, const float* gamma, const float* beta, const float* res, float* out, float* stats, float* partial,
                 int B, int HW, int C, int relu, cudaStream_t st) {
    (void)partial;
    if (C % 16 != 0 || 256 % (C / 16) != 0) return DBOA_ERR_SHAPE;
    const GnPlan pl = gn_plan(HW, C);
    if (pl.chunks > GN_MAXCL) return DBOA_ERR_SHAPE;
    dim3 grid(pl.chunks, GN_G, B);
    return launch_x_cluster(gn_fwd_fused_kernel, grid, st, y, gamma, beta, res, out, stats, HW, C, pl.rows, relu);
}

// backward: grid (chunks, 4, B), cluster (chunks, 1, 1)
__global__ void __launch_bounds__(GN_NT) gn_bwd_fused_kernel(const float* __restrict__ dout, const float* __restrict__ mask_src,
                                                             const float* __restrict__ y, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, float* __restrict__ dy,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             float* __restrict__ dgpart, float* __restrict__ dbpart, unsigned* pcounters,
                                                             int HW, int C, int R) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[32];
    __shared__ float part[2];
    __shared__ float all[GN_MAXCL][2];
    __shared__ float sm[2];
    __shared__ int s_plast;
    __shared__ __align__(16) float smg[GN_NT * 4], smb[GN_NT * 4];
    const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z, chunks = gridDim.x, B = gridDim.z, slot = b * GN_G + g;
    const int cg = C / GN_G, cg4 = cg / 4;
    const int r0 = chunk * R, rows = min(R, HW - r0);
    const int nvec = rows * cg4;
    const size_t base = ((size_t)b * HW + r0) * C + g * cg;
    const float mean = stats[slot * 2], rstd = stats[slot * 2 + 1];
    const int cv = threadIdx.x % cg4;
    const float4 ga = ldg4(gamma + g * cg + cv * 4);
    float4 gq[GN_VMAX], xh[GN_VMAX];
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < GN_VMAX; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            int row = idx / cg4;
            const size_t e = base + (size_t)row * C + cv * 4;
            float4 d = ldg4(dout + e), v = ldg4(y + e);
            if (mask_src != nullptr) {
                float4 m = ldg4(mask_src + e);
                d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f; d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
            }
            xh[i] = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
            gq[i] = make_float4(d.x * ga.x, d.y * ga.y, d.z * ga.z, d.w * ga.w);
            dg.x += d.x * xh[i].x; dg.y += d.y * xh[i].y; dg.z += d.z * xh[i].z; dg.w += d.w * xh[i].w;
            db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
            s1 += (gq[i].x + gq[i].y) + (gq[i].z + gq[i].w);
            s2 += (gq[i].x * xh[i].x + gq[i].y * xh[i].y) + (gq[i].z * xh[i].z + gq[i].w * xh[i].w);
        }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { part[0] = s1; part[1] = s2; }
    *reinterpret_cast<float4*>(&smg[threadIdx.x * 4]) = dg;
    *reinterpret_cast<float4*>(&smb[threadIdx.x * 4]) = db;
    cluster.sync();
    if (threadIdx.x < chunks * 2) {
        const int c = threadIdx.x >> 1, k = threadIdx.x & 1;
        all[c][k] = cluster.map_shared_rank(part, c)[k];
    }
    // per-channel partial sums of this chunk (fixed-order smem reduction) -> global rows for the affine gradients
    for (int c = threadIdx.x; c < cg; c += GN_NT) {
        const int ccv = c >> 2, comp = c & 3, rp = GN_NT / cg4;
        float a = 0.f, bsum = 0.f;
        for (int j = 0; j < rp; ++j) { a += smg[(j * cg4 + ccv) * 4 + comp]; bsum += smb[(j * cg4 + ccv) * 4 + comp]; }
        const size_t o = ((size_t)b * chunks + chunk) * C + g * cg + c;
        dgpart[o] = a; dbpart[o] = bsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int k = 0; k < chunks; ++k) { a += all[k][0]; c += all[k][1]; }
        sm[0] = a; sm[1] = c;
    }
    __syncthreads();
    const float invN = 1.0f / ((float)HW * (float)cg);
    const float m1 = sm[0] * invN, m2 = sm[1] * invN;
#pragma unroll
    for (int i = 0; i < GN_VMAX; ++i) {
        int idx = threadIdx.x + i * GN_NT;
        if (idx < nvec) {
            int row = idx / cg4;
            const size_t e = base + (size_t)row * C + cv * 4;
            float4 o;
            o.x = rstd * (gq[i].x - m1 - xh[i].x * m2); o.y = rstd * (gq[i].y - m1 - xh[i].y * m2);
            o.z = rstd * (gq[i].z - m1 - xh[i].z * m2); o.w = rstd * (gq[i].w - m1 - xh[i].w * m2);
            *reinterpret_cast<float4*>(dy + e) = o;
        }
    }
    // affine-parameter gradients: the last CTA of group g (over all samples and chunks) sums the rows in order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&pcounters[g], 1u);
        s_plast = (t == (unsigned)(B * chunks) - 1);
        if (s_plast) pcounters[g] = 0;
    }
    __syncthreads();
    if (s_plast) {
        __threadfence();
        const int nrows = B * chunks;
        for (int c = threadIdx.x; c < cg; c += GN_NT) {
            float a = 0.f, bsum = 0.f;
#pragma unroll 4
            for (int r = 0; r < nrows; ++r) { a += __ldcg(dgpart + (size_t)r * C + g * cg + c); bsum += __ldcg(dbpart + (size_t)r * C + g * cg + c); }
            dgamma[g * cg + c] += a; dbeta[g * cg + c] += bsum;
        }
    }
    cluster.sync();                                        // peers may still be reading `part`
}

int gn_bwd_fused(const float* dout, const float* mask_src, const float* y, const float* stats, const float* gamma, float* dy,
                 float* dgamma, float* dbeta, float* partial, int B, int HW, int C, cudaStream_t st) {
    if (C % 16 != 0 || 256 % (C / 16) != 0) return DBOA_ERR_SHAPE;
    const GnPlan pl = gn_plan(HW, C);
    if (pl.chunks > GN_MAXCL) return DBOA_ERR_SHAPE;
    float* dgpart = partial;
    float* dbpart = partial + (size_t)B * pl.chunks * C;
    unsigned* cnt = sync_words(SYNC_GNB_PCNT);
    if (!cnt) return DBOA_ERR_CUDA;
    dim3 grid(pl.chunks, GN_G, B);
    return launch_x_cluster(gn_bwd_fused_kernel, grid, st, dout, mask_src, y, stats, gamma, dy, dgamma, dbeta, dgpart, dbpart, cnt, HW, C,
                            pl.rows);
}

// ---------------------------------------------------------------------------------------------
__global__ void relu_mask_kernel(const float* __restrict__ dout, const float* __restrict__ mask_src, float* __restrict__ dz, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 d = ldg4(dout + i * 4), m = ldg4(mask_src + i * 4);
    d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f; d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
    *reinterpret_cast<float4*>(dz + i * 4) = d;
}
int relu_mask(const float* dout, const float* mask_src, float* dz, size_t n, cudaStream_t st) {
    relu_mask_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(dout, mask_src, dz, n / 4);
    return check_launch();
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over B*HW pixels
    if (i >= total) return;
    size_t b = i / HW, p = i - b * HW;
    for (int c = 0; c < C; ++c) y[i * C + c] = __ldg(x + (b * C + c) * HW + p);
}
int nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, cudaStream_t st) {
    size_t total = (size_t)B * H * W;
    nchw_to_nhwc_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, y, C, H * W, total);
    return check_launch();
}

// MaxPool2d(kernel 3, stride 2, pad 1); first maximum in (row, col) scan order wins, as ATen does.
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ idx, int B, int H,
                                   int W, int C) {
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * Ho * Wo * C4;
    if (i >= total) return;
    int cv = (int)(i % C4);
    size_t p = i / C4;
    int wo = (int)(p % Wo); p /= Wo;
    int ho = (int)(p % Ho);
    int b = (int)(p / Ho);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned char bi[4] = {0, 0, 0, 0};
    for (int r = 0; r < 3; ++r) {
        int hi = ho * 2 - 1 + r;
        if ((unsigned)hi >= (unsigned)H) continue;
        for (int s = 0; s < 3; ++s) {
            int wi = wo * 2 - 1 + s;
            if ((unsigned)wi >= (unsigned)W) continue;
            float4 v = ldg4(x + (((size_t)b * H + hi) * W + wi) * C + cv * 4);
            float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (vv[e] > best[e] || vv[e] != vv[e]) { best[e] = vv[e]; bi[e] = (unsigned char)(r * 3 + s); }
        }
    }
    size_t o = (((size_t)b * Ho + ho) * Wo + wo) * C + cv * 4;
    *reinterpret_cast<float4*>(y + o) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<uchar4*>(idx + o) = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
}
int maxpool3x3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C, cudaStream_t st) {
    size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 4);
    maxpool_fwd_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, y, idx, B, H, W, C);
    return check_launch();
}

__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx, float* __restrict__ dx, int B,
                                   int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * H * W * C4;
    if (i >= total) return;
    int cv = (int)(i % C4);
    size_t p = i / C4;
    int wi = (int)(p % W); p /= W;
    int hi = (int)(p % H);
    int b = (int)(p / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ho = hi / 2; ho <= (hi + 1) / 2; ++ho) {
        if (ho >= Ho) continue;
        int r = hi - (ho * 2 - 1);
        if (r < 0 || r > 2) continue;
        for (int wo = wi / 2; wo <= (wi + 1) / 2; ++wo) {
            if (wo >= Wo) continue;
            int s = wi - (wo * 2 - 1);
            if (s < 0 || s > 2) continue;
            size_t o = (((size_t)b * Ho + ho) * Wo + wo) * C + cv * 4;
            uchar4 k = *reinterpret_cast<const uchar4*>(idx + o);
            float4 g = ldg4(dy + o);
            unsigned char me = (unsigned char)(r * 3 + s);
            if (k.x == me) acc[0] += g.x;
            if (k.y == me) acc[1] += g.y;
            if (k.z == me) acc[2] += g.z;
            if (k.w == me) acc[3] += g.w;
        }
    }
    *reinterpret_cast<float4*>(dx + (((size_t)b * H + hi) * W + wi) * C + cv * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}
int maxpool3x3s2_bwd(const float* dy, const unsigned char* idx, float* dx, int B, int H, int W, int C, cudaStream_t st) {
    size_t total = (size_t)B * H * W * (C / 4);
    maxpool_bwd_kernel<<<ceil_div(total, 256), 256, 0, st>>>(dy, idx, dx, B, H, W, C);
    return check_launch();
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int HW, int C, int ld, int ncopy,
                                   size_t copy_stride, int total) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int b = i / C, c = i - b * C;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __ldg(x + ((size_t)b * HW + p) * C + c);
    s = s / (float)HW;
    for (int k = 0; k < ncopy; ++k) out[k * copy_stride + (size_t)b * ld + c] = s;
}
int avgpool_fwd(const float* x, float* out, int B, int HW, int C, int ld, int ncopy, size_t copy_stride, cudaStream_t st) {
    avgpool_fwd_kernel<<<ceil_div(B * C, 256), 256, 0, st>>>(x, out, HW, C, ld, ncopy, copy_stride, B * C);
    return check_launch();
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dxf, int ld, float* __restrict__ dx, int HW, int C, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % C);
    size_t b = i / ((size_t)HW * C);
    dx[i] = dxf[b * ld + c] / (float)HW;
}
int avgpool_bwd(const float* dxf, int ld, float* dx, int B, int HW, int C, cudaStream_t st) {
    size_t total = (size_t)B * HW * C;
    avgpool_bwd_kernel<<<ceil_div(total, 256), 256, 0, st>>>(dxf, ld, dx, HW, C, total);
    return check_launch();
}

}  // namespace dboa
