// Fused data gradient of a stride-1 backbone convolution on tcgen05 tensor cores (the backward mirror of conv_wide.cu):
//
//      dy   = GroupNorm_c backward of dz          (on load: dy = rstd (dz gamma - m1 - x^ m2), x^ = (y - mean) rstd)
//      dX   = conv_c^T(dy)  [+ addend]            (implicit GEMM: rows = input pixels, K = (tap, co), flipped taps)
//      dz_p = dX * (a_p > 0)                      (epilogue: ReLU mask of the producing layer p)
//      sums of GroupNorm_p's backward            (epilogue: sum q, sum q x^ per (sample, group); d gamma, d beta per channel)
//
// Replaces, per layer, the launches  gn_bwd_fused -> conv dgrad (-> relu_mask)  of round 1 (reference: the autograd backward of
// nn.GroupNorm / ReLU / nn.Conv2d / the residual add in Bottleneck.forward, model/hmr.py:40-60, under MAML.adapt / loss.backward(),
// dynaboa_benchmark.py:140,150): GroupNorm backward needs two group-wide sums before it can produce dy, so round 1 ran it as its
// own cluster kernel (53 launches of ~10 us on the critical chain of every backward).  Here those sums are produced by the
// epilogue of the kernel that creates dz (this kernel, one layer later in the chain), as 64-bit fixed-point atomics (exact, order
// independent -> deterministic), and GroupNorm backward itself happens while the operand tile is in shared memory.
//   * A operand (dz and y of layer c): two 4-D TMA boxes per k-block, K-major 128B swizzle; one designated tap also writes dy to
//     memory, because the weight-gradient kernel (side stream) reads it;
//   * B operand (W_c^T): W[co][tap][ci] is ci-contiguous, i.e. MN-major for this GEMM; TMA boxes of 32 ci x 32 co with the
//     128B / 32-byte-atom swizzle feed tcgen05 directly (instruction descriptor b_major = MN) -- no transposing store;
//   * TF32x3 split and GroupNorm-backward transform: one elementwise pass over the landed tiles (16 warps);
//   * split-K over a thread-block cluster, DSMEM reduction (as conv_wide.cu).
// Stride-2 layers stay on conv_tc.cu + groupnorm.cu (hmr_plan.cu stitches the two worlds with gn_bwd_prep below).
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace dboa {
const void* tma_act_map(const float* x, int B, int H, int W, int C, int bw, int bh, bool atom32, int stride);      // conv_wide.cu
const void* tma_weight_map_mn(const float* w, int K, int Cout);                                         // conv_wide.cu

namespace dz {

constexpr int BM = 128, BN = 64, BK = 32;
constexpr int NTW = 16, NTT = NTW * 32, W_MMA = 16, W_TMA = 17, NT = 576;
constexpr int DMAX = 3;
constexpr int NACC = 4;                                  // TMEM accumulators a reduction chain rotates over (see conv_wide.cu)
constexpr uint32_t A_TILE = BM * BK * 4, B_TILE = BN * BK * 4;          // 16 KB, 8 KB (2 atoms of 32 ci x 32 co)
constexpr uint32_t SLOT = 2 * A_TILE + B_TILE;                          // dz tile, y tile, weight tile
constexpr int RED_LD = BN + 4;
constexpr double FIX = 268435456.0;                                     // 2^28 fixed point of the backward sums

struct Prep {                      // GroupNorm whose backward the epilogue prepares (layer p = producer of this conv's input)
    const float* y;                // raw output of layer p [B][H][W][C]
    const float* stats;            // (mean, rstd) [B][4][2]
    const float* gamma;
    unsigned long long* sums;      // [B][4][2]: sum q, sum q x^
    unsigned long long* dgb;       // [C][2]: d gamma, d beta
};
struct Launch {
    // GroupNorm_c backward on load
    const float* stats_c; const long long* sums_c; const float* gamma_c;
    float* dy_out;                 // materialised dy_c [B][H][W][Cout] (weight gradient operand) or NULL
    // output side
    const float* addend;           // [B][H][W][Cin] or NULL
    float* out;                    // plain mode: dX; prep mode: dz_p
    const float* mask;             // prep mode: a_p (post-activation output of layer p); NULL = plain mode
    Prep p[2];
    int nprep;
    int H, W, Cin, Cout, k, pad, B;
    int bh, tps, ntiles, nz, per, D, tabc, accumulate;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {       // K-major, 128B swizzle (activations)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((1024 >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major TF32 (weights): SWIZZLE_128B_BASE32B, LBO = 4096 B between the two 32-channel atoms, SBO = 512 B between 4-row groups
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((4096 >> 4) & 0x3FFF) << 16) | ((uint64_t)((512 >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           (1ull << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, addr = smem_u32(bar);
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000ll) __trap();       // a protocol error fails the launch instead of hanging the device
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ long long to_fix(float v) { return __double2ll_rn((double)v * FIX); }

__global__ void __launch_bounds__(NT, 1) dgrad_wide_kernel(const __grid_constant__ Launch L, const __grid_constant__ CUtensorMap tmdz,
                                                           const __grid_constant__ CUtensorMap tmy, const __grid_constant__ CUtensorMap tmw) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nz = L.nz, D = L.D;
    const int cidx = blockIdx.x / nz, rank = blockIdx.x - cidx * nz;
    const int H = L.H, W = L.W, Cin = L.Cin, Cout = L.Cout, ks = L.k, pad = L.pad, bh = L.bh;
    const int nt = cidx % L.ntiles, bm = cidx / L.ntiles, mt = bm % L.tps, b = bm / L.tps;
    const int h0 = mt * bh, n0 = nt * BN;
    const int rows_valid = min(bh, H - h0) * W, m0 = h0 * W;
    const int nkb_total = (ks * ks * Cout) / BK;
    const int kb_begin = rank * L.per;
    const int nkb = max(0, min(L.per, nkb_total - kb_begin));
    const uint32_t a_bytes = (uint32_t)(bh * W) * 128u;

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* slots = base;                                  // D x {dz 16 KB, y 16 KB, W^T 8 KB}
    uint8_t* lo_a = slots + (size_t)D * SLOT;               // 2 x 16 KB
    uint8_t* lo_b = lo_a + 2 * A_TILE;                      // 2 x 8 KB
    float* tab = reinterpret_cast<float*>(lo_b + 2 * B_TILE);          // gamma_c of this CTA's channel range
    uint64_t* bars = reinterpret_cast<uint64_t*>(tab + L.tabc);
    uint64_t* s_full = bars;               // [DMAX]
    uint64_t* s_empty = bars + DMAX;       // [DMAX]
    uint64_t* l_full = bars + 2 * DMAX;    // [2]
    uint64_t* l_empty = l_full + 2;        // [2]
    uint64_t* done = l_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    float* sstat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);                     // [16]: mean, rstd, m1, m2 per group
    unsigned long long* sacc = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(bars) + 256);    // [2 GN][4 groups][2]
    unsigned long long* sgb = sacc + 16;                                                                  // [2 GN][64 channels][2]
    float* red = reinterpret_cast<float*>(lo_a);

    if (tid == 0) {
        for (int s = 0; s < DMAX; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&l_full[s], NTW); mbar_init(&l_empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN * NACC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // gamma of layer c for the output channels this CTA reduces over (parameters: no dependency on the previous kernel)
    const int tc0 = ks == 1 ? kb_begin * BK : 0, tcn = ks == 1 ? nkb * BK : Cout;
    for (int i = tid; i < tcn; i += NT) tab[i] = __ldg(L.gamma_c + tc0 + i);
    for (int i = tid; i < 16 + 2 * 128; i += NT) sacc[i] = 0ull;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    // k-block kb: filter tap (r, s) and output-channel offset c
    auto tap_of = [&](int kb, int& r, int& s, int& c) {
        const int k0 = kb * BK, tap = k0 / Cout;
        c = k0 - tap * Cout; r = tap / ks; s = tap - r * ks;
    };

    if (warp == W_TMA) {
        if (lane == 0 && nkb > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmw)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmdz)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmy)) : "memory");
            auto load_w = [&](int it, int sl) {             // W^T tile: 2 boxes of 32 ci x 32 co at (tap * Cin + n0 + 32 j, co)
                int r, s, c;
                tap_of(kb_begin + it, r, s, c);
                uint8_t* wt = slots + (size_t)sl * SLOT + 2 * A_TILE;
                tma_load_2d(smem_u32(wt), &tmw, (r * ks + s) * Cin + n0, c, &s_full[sl]);
                tma_load_2d(smem_u32(wt + 4096), &tmw, (r * ks + s) * Cin + n0 + 32, c, &s_full[sl]);
            };
            const int npre = min(nkb, D);
            for (int it = 0; it < npre; ++it) {
                mbar_expect_tx(&s_full[it], 2 * a_bytes + B_TILE);
                load_w(it, it);
            }
            pdl_wait();
            pdl_trigger();
            for (int it = 0; it < nkb; ++it) {
                const int sl = it % D;
                uint8_t* slot = slots + (size_t)sl * SLOT;
                if (it >= D) {
                    mbar_wait(&s_empty[sl], (uint32_t)(((it / D) - 1) & 1));
                    mbar_expect_tx(&s_full[sl], 2 * a_bytes + B_TILE);
                    load_w(it, sl);
                }
                int r, s, c;
                tap_of(kb_begin + it, r, s, c);
                // dX[hi][wi] += dy[hi + pad - r][wi + pad - s] W[co][r][s][ci]: the box starts at (pad - s, h0 + pad - r)
                tma_load_4d(smem_u32(slot), &tmdz, c, pad - s, h0 + pad - r, b, &s_full[sl]);
                tma_load_4d(smem_u32(slot + A_TILE), &tmy, c, pad - s, h0 + pad - r, b, &s_full[sl]);
            }
        } else {
            pdl_wait();
            pdl_trigger();
        }
    } else if (warp == W_MMA) {
        if (lane == 0 && nkb > 0) {
            // D = F32, A = B = TF32, A K-major, B MN-major (bit 16), N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint64_t dslot = desc_sw128(smem_u32(slots)), dloa = desc_sw128(smem_u32(lo_a));
            const uint64_t dwslot = desc_mn(smem_u32(slots + 2 * A_TILE)), dlob = desc_mn(smem_u32(lo_b));
            constexpr uint64_t KSTEP_A = 32 >> 4;            // 8 channels = 32 bytes inside the 128-byte row
            constexpr uint64_t KSTEP_B = 1024 >> 4;          // 8 output channels = 8 rows of the MN-major tile
#pragma unroll 1
            for (int it = 0; it < nkb; ++it) {
                const int sl = it % D, ls = it & 1;
                mbar_wait(&l_full[ls], (uint32_t)((it >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t so = (uint64_t)((sl * SLOT) >> 4);
                const uint64_t dah = dslot + so, dbh = dwslot + so;
                const uint64_t dal = dloa + (uint64_t)((ls * A_TILE) >> 4), dbl = dlob + (uint64_t)((ls * B_TILE) >> 4);
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    const uint32_t dacc = tmem_d + (uint32_t)((it & (NACC - 1)) * BN);      // truncating accumulation: short chains
                    mma_tf32(dacc, dah + kk * KSTEP_A, dbh + kk * KSTEP_B, idesc, (it >= NACC || kk > 0) ? 1u : 0u);
                    mma_tf32(dacc, dah + kk * KSTEP_A, dbl + kk * KSTEP_B, idesc, 1u);
                    mma_tf32(dacc, dal + kk * KSTEP_A, dbh + kk * KSTEP_B, idesc, 1u);
                }
                umma_commit(&l_empty[ls]);
                umma_commit(&s_empty[sl]);
            }
            umma_commit(done);
        }
        pdl_wait();
        pdl_trigger();
    } else {
        // ---- transform warps: GroupNorm_c backward + TF32 split of (dz, y) -> dy hi / lo; split of the weight tile
        const int r0 = tid >> 3, pc = tid & 7, lc = pc ^ (r0 & 7);
        int oh[2], ow[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) { const int i = r0 + 64 * q; oh[q] = i / W; ow[q] = i - oh[q] * W; }
        float* dyb = (L.dy_out != nullptr && nt == 0) ? L.dy_out + (size_t)b * H * W * Cout : nullptr;
        int lgw = 0;
        while ((4 << lgw) < Cout) ++lgw;
        pdl_wait();
        pdl_trigger();
        if (tid < 4) {                                      // (mean, rstd, m1, m2) of group tid of sample b
            const float* st = L.stats_c + ((size_t)b * 4 + tid) * 2;
            const long long* sm = L.sums_c + ((size_t)b * 4 + tid) * 2;
            const double N = (double)H * W * (Cout >> 2);
            sstat[tid] = __ldcg(st); sstat[4 + tid] = __ldcg(st + 1);
            sstat[8 + tid] = (float)((double)__ldcg(sm) / FIX / N);
            sstat[12 + tid] = (float)((double)__ldcg(sm + 1) / FIX / N);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NTT) : "memory");
        float gmean[4], grstd[4], gm1[4], gm2[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) { gmean[g] = sstat[g]; grstd[g] = sstat[4 + g]; gm1[g] = sstat[8 + g]; gm2[g] = sstat[12 + g]; }
#pragma unroll 1
        for (int it = 0; it < nkb; ++it) {
            const int sl = it % D, ls = it & 1;
            uint8_t* slot = slots + (size_t)sl * SLOT;
            int r, s, c;
            tap_of(kb_begin + it, r, s, c);
            const int cch = c + lc * 4, g = cch >> lgw;
            const float4 ga = *reinterpret_cast<const float4*>(tab + (cch - tc0));
            const float mu = g == 0 ? gmean[0] : (g == 1 ? gmean[1] : (g == 2 ? gmean[2] : gmean[3]));
            const float rs = g == 0 ? grstd[0] : (g == 1 ? grstd[1] : (g == 2 ? grstd[2] : grstd[3]));
            const float m1 = g == 0 ? gm1[0] : (g == 1 ? gm1[1] : (g == 2 ? gm1[2] : gm1[3]));
            const float m2 = g == 0 ? gm2[0] : (g == 1 ? gm2[1] : (g == 2 ? gm2[2] : gm2[3]));
            const bool desig = ks == 1 || (r == 1 && s == 1);
            mbar_wait(&s_full[sl], (uint32_t)((it / D) & 1));
            if (it >= 2) mbar_wait(&l_empty[ls], (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t off = (uint32_t)(tid + q * NTT) * 16u;
                const float4 d = *reinterpret_cast<const float4*>(slot + off), yv = *reinterpret_cast<const float4*>(slot + A_TILE + off);
                const int ho = h0 + oh[q] + pad - r, wo = ow[q] + pad - s;           // pixel of layer c's output this row reads
                const bool inb = (r0 + 64 * q) < rows_valid && (unsigned)ho < (unsigned)H && (unsigned)wo < (unsigned)W;
                float4 o;
                o.x = rs * (d.x * ga.x - m1 - ((yv.x - mu) * rs) * m2); o.y = rs * (d.y * ga.y - m1 - ((yv.y - mu) * rs) * m2);
                o.z = rs * (d.z * ga.z - m1 - ((yv.z - mu) * rs) * m2); o.w = rs * (d.w * ga.w - m1 - ((yv.w - mu) * rs) * m2);
                if (!inb) o = make_float4(0.f, 0.f, 0.f, 0.f);
                else if (dyb != nullptr && desig) *reinterpret_cast<float4*>(dyb + ((size_t)ho * W + wo) * Cout + cch) = o;
                const float4 h = make_float4(tf32_hi(o.x), tf32_hi(o.y), tf32_hi(o.z), tf32_hi(o.w));
                *reinterpret_cast<float4*>(slot + off) = h;
                *reinterpret_cast<float4*>(lo_a + ls * A_TILE + off) = make_float4(o.x - h.x, o.y - h.y, o.z - h.z, o.w - h.w);
            }
            {
                const uint32_t off = (uint32_t)tid * 16u;
                uint8_t* wraw = slot + 2 * A_TILE;
                const float4 v = *reinterpret_cast<const float4*>(wraw + off);
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                *reinterpret_cast<float4*>(wraw + off) = h;
                *reinterpret_cast<float4*>(lo_b + ls * B_TILE + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&l_full[ls]);
        }
    }
    if (nkb > 0) mbar_wait(done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM -> shared memory, cluster reduction, [+ addend], ReLU mask, stores, GroupNorm_p backward sums
    if (warp < NTW) {
        const int q4 = warp & 3, cgp = warp >> 2;
        float facc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) facc[q] = 0.f;
        const int nacc = nkb < NACC ? nkb : NACC;
#pragma unroll 1
        for (int a = 0; a < nacc; ++a) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * BN + cgp * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; ++q) facc[q] += __uint_as_float(v[q]);
        }
        float* dstrow = red + (q4 * 32 + lane) * RED_LD + cgp * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(dstrow + q * 4) = make_float4(facc[q * 4], facc[q * 4 + 1], facc[q * 4 + 2], facc[q * 4 + 3]);
    }
    cg::cluster_group cluster = cg::this_cluster();
    if (nz == 1) __syncthreads(); else cluster.sync();

    const int rows_per = BM / nz, items = rows_per * (BN / 4);
    const int gw = Cin >> 2, gpt = gw >= BN ? 1 : BN / gw, lpg = 16 / gpt;
    const bool prep = L.mask != nullptr;
    if (warp < NTW) {
        const size_t img = ((size_t)b * H * W + m0) * Cin;
        const int c4 = (tid & 15) * 4, cabs = n0 + c4, g = cabs / gw;           // this thread's 4 channels: fixed over the loop
        float4 sdg[2], sdb = make_float4(0.f, 0.f, 0.f, 0.f);
        sdg[0] = sdg[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        float sq[2] = {0.f, 0.f}, sqx[2] = {0.f, 0.f};
        float4 gam[2];
        float pmu[2] = {0.f, 0.f}, prs[2] = {1.f, 1.f};
        if (prep) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (j < L.nprep) {
                    gam[j] = ldg4(L.p[j].gamma + cabs);
                    pmu[j] = __ldg(L.p[j].stats + ((size_t)b * 4 + g) * 2); prs[j] = __ldg(L.p[j].stats + ((size_t)b * 4 + g) * 2 + 1);
                }
        }
#pragma unroll 1
        for (int v = tid; v < items; v += NTT) {
            const int lr = rank * rows_per + (v >> 4);
            if (lr < rows_valid) {
                float4 acc;
                if (nz == 1) {
                    acc = *reinterpret_cast<const float4*>(red + lr * RED_LD + c4);
                } else {
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
                    for (int zb = 0; zb < nz; zb += 8) {
                        float4 q[8];
#pragma unroll
                        for (int z = 0; z < 8; ++z)
                            if (zb + z < nz) q[z] = *reinterpret_cast<const float4*>(cluster.map_shared_rank(red, zb + z) + lr * RED_LD + c4);
#pragma unroll
                        for (int z = 0; z < 8; ++z)
                            if (zb + z < nz) { acc.x += q[z].x; acc.y += q[z].y; acc.z += q[z].z; acc.w += q[z].w; }
                    }
                }
                const size_t e = img + (size_t)lr * Cin + cabs;
                if (L.addend != nullptr) { const float4 a = __ldcg(reinterpret_cast<const float4*>(L.addend + e)); acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w; }
                if (prep) {
                    const float4 m = __ldcg(reinterpret_cast<const float4*>(L.mask + e));
                    acc.x = m.x > 0.f ? acc.x : 0.f; acc.y = m.y > 0.f ? acc.y : 0.f; acc.z = m.z > 0.f ? acc.z : 0.f; acc.w = m.w > 0.f ? acc.w : 0.f;
                    sdb.x += acc.x; sdb.y += acc.y; sdb.z += acc.z; sdb.w += acc.w;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (j < L.nprep) {
                            const float4 yv = __ldcg(reinterpret_cast<const float4*>(L.p[j].y + e));
                            const float x0 = (yv.x - pmu[j]) * prs[j], x1 = (yv.y - pmu[j]) * prs[j], x2 = (yv.z - pmu[j]) * prs[j], x3 = (yv.w - pmu[j]) * prs[j];
                            const float q0 = acc.x * gam[j].x, q1 = acc.y * gam[j].y, q2 = acc.z * gam[j].z, q3 = acc.w * gam[j].w;
                            sdg[j].x += acc.x * x0; sdg[j].y += acc.y * x1; sdg[j].z += acc.z * x2; sdg[j].w += acc.w * x3;
                            sq[j] += (q0 + q1) + (q2 + q3);
                            sqx[j] += (q0 * x0 + q1 * x1) + (q2 * x2 + q3 * x3);
                        }
                } else if (L.accumulate) {
                    const float4 a = *reinterpret_cast<const float4*>(L.out + e);
                    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
                }
                *reinterpret_cast<float4*>(L.out + e) = acc;
            }
        }
        if (prep) {
            // everything below is integer: per-thread sums -> 2^28 fixed point -> warp shuffles -> shared-memory atomics
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (j < L.nprep) {
                    long long a = to_fix(sq[j]), c = to_fix(sqx[j]);
#pragma unroll 1
                    for (int o = 16; o >= 1; o >>= 1)
                        if (o == 16 || o < lpg) { a += __shfl_down_sync(0xffffffffu, a, o); c += __shfl_down_sync(0xffffffffu, c, o); }
                    if (lane < 16 && (lane % lpg) == 0) {
                        atomicAdd(&sacc[j * 8 + (lane / lpg) * 2], (unsigned long long)a);
                        atomicAdd(&sacc[j * 8 + (lane / lpg) * 2 + 1], (unsigned long long)c);
                    }
                    // per-channel d gamma (this GroupNorm) and d beta (the same for both): rows of lanes l and l + 16
                    long long dg[4] = {to_fix(sdg[j].x), to_fix(sdg[j].y), to_fix(sdg[j].z), to_fix(sdg[j].w)};
                    long long db[4] = {to_fix(sdb.x), to_fix(sdb.y), to_fix(sdb.z), to_fix(sdb.w)};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { dg[e] += __shfl_down_sync(0xffffffffu, dg[e], 16); db[e] += __shfl_down_sync(0xffffffffu, db[e], 16); }
                    if (lane < 16) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            atomicAdd(&sgb[(j * 64 + c4 + e) * 2], (unsigned long long)dg[e]);
                            atomicAdd(&sgb[(j * 64 + c4 + e) * 2 + 1], (unsigned long long)db[e]);
                        }
                    }
                }
        }
    }
    __syncthreads();
    if (prep) {
        for (int j = 0; j < L.nprep; ++j) {
            if (tid < 2 * gpt) {
                const int gi = tid >> 1, g = gw >= BN ? (nt * BN) / gw : nt * gpt + gi;
                atomicAdd(L.p[j].sums + ((size_t)b * 4 + g) * 2 + (tid & 1), sacc[j * 8 + tid]);
            }
            if (tid >= 64 && tid < 64 + 128) {
                const int i = tid - 64;
                atomicAdd(L.p[j].dgb + (size_t)(n0 + (i >> 1)) * 2 + (i & 1), sgb[j * 128 + i]);
            }
        }
    }
    if (nz > 1) cluster.sync();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN * NACC) : "memory");
}

// Stand-alone version of the epilogue above for the seams with the unfused kernels (after avgpool_bwd; after a stride-2 block):
// dz = dA * (a > 0) -> out; sums / d gamma / d beta of GroupNorm_p.  grid (ceil(HW * C / 4 / 256 / 8), B): one thread = one
// float4 column, 8 rows apart... simple elementwise layout: thread t owns channel vector cv = t % (C/4) of rows t / (C/4) + k * rstep.
__global__ void __launch_bounds__(256) gn_bwd_prep_kernel(const float* __restrict__ dA, const float* __restrict__ mask, float* __restrict__ out,
                                                          Prep p, int HW, int C, int rows_per_cta) {
    __shared__ unsigned long long sacc[8];
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.y, C4 = C >> 2, gw = C >> 2;
    if (threadIdx.x < 8) sacc[threadIdx.x] = 0ull;
    __syncthreads();
    const int r_begin = blockIdx.x * rows_per_cta, r_end = min(HW, r_begin + rows_per_cta);
    // threads stride over channel vectors; every thread keeps its channel vector over all rows of the CTA (C4 <= 512: up to 2 per thread)
    for (int cv = threadIdx.x; cv < C4; cv += 256) {
        const int c = cv * 4, g = c / gw;
        const float4 ga = ldg4(p.gamma + c);
        const float mu = __ldg(p.stats + ((size_t)b * 4 + g) * 2), rs = __ldg(p.stats + ((size_t)b * 4 + g) * 2 + 1);
        float4 sdg = make_float4(0.f, 0.f, 0.f, 0.f), sdb = sdg;
        float sq = 0.f, sqx = 0.f;
        for (int r = r_begin; r < r_end; ++r) {
            const size_t e = ((size_t)b * HW + r) * C + c;
            float4 d = ldg4(dA + e);
            const float4 m = ldg4(mask + e), yv = ldg4(p.y + e);
            d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f; d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
            *reinterpret_cast<float4*>(out + e) = d;
            const float x0 = (yv.x - mu) * rs, x1 = (yv.y - mu) * rs, x2 = (yv.z - mu) * rs, x3 = (yv.w - mu) * rs;
            const float q0 = d.x * ga.x, q1 = d.y * ga.y, q2 = d.z * ga.z, q3 = d.w * ga.w;
            sdg.x += d.x * x0; sdg.y += d.y * x1; sdg.z += d.z * x2; sdg.w += d.w * x3;
            sdb.x += d.x; sdb.y += d.y; sdb.z += d.z; sdb.w += d.w;
            sq += (q0 + q1) + (q2 + q3);
            sqx += (q0 * x0 + q1 * x1) + (q2 * x2 + q3 * x3);
        }
        atomicAdd(&sacc[g * 2], (unsigned long long)to_fix(sq));
        atomicAdd(&sacc[g * 2 + 1], (unsigned long long)to_fix(sqx));
        const float dgv[4] = {sdg.x, sdg.y, sdg.z, sdg.w}, dbv[4] = {sdb.x, sdb.y, sdb.z, sdb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            atomicAdd(p.dgb + (size_t)(c + e) * 2, (unsigned long long)to_fix(dgv[e]));
            atomicAdd(p.dgb + (size_t)(c + e) * 2 + 1, (unsigned long long)to_fix(dbv[e]));
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) atomicAdd(p.sums + (size_t)b * 8 + threadIdx.x, sacc[threadIdx.x]);
}

// G[g_off + c] += d gamma, G[b_off + c] += d beta from the fixed-point accumulators of every GroupNorm (one launch per backward)
__global__ void __launch_bounds__(256) gn_dgb_finish_kernel(const GnFinishItem* __restrict__ items, const long long* __restrict__ dgb, float* __restrict__ G) {
    pdl_wait();
    pdl_trigger();
    const GnFinishItem it = items[blockIdx.x];
    const long long* a = dgb + 2 * it.cum_channels;
    for (int c = threadIdx.x; c < it.C; c += 256) {
        const long long vg = a[2 * c], vb = a[2 * c + 1];
        if (vg != 0) G[it.g_off + c] += (float)((double)vg / FIX);
        if (vb != 0) G[it.b_off + c] += (float)((double)vb / FIX);
    }
}

}  // namespace dz

bool dgrad_wide_ok(const ConvDims& d) {
    return d.stride == 1 && d.Ho == d.Hi && d.Wo == d.Wi && d.Hi == d.Wi && d.Cin % 64 == 0 && d.Cout % 64 == 0 && (d.kh == 1 || d.kh == 3) && d.kh == d.kw &&
           d.pad == d.kh / 2 && d.Kpitch == d.kh * d.kw * d.Cin && d.Hi <= 128;
}

int dgrad_wide(const DgradFused& f, const ConvDims& d, cudaStream_t st, bool pdl) {
    if (!dgrad_wide_ok(d)) return DBOA_ERR_UNSUPPORTED;
    dz::Launch L;
    memset(&L, 0, sizeof L);
    L.stats_c = f.stats_c; L.sums_c = reinterpret_cast<const long long*>(f.sums_c); L.gamma_c = f.gamma_c; L.dy_out = f.dy_out;
    L.addend = f.addend; L.out = f.out; L.mask = f.mask; L.nprep = f.mask ? f.nprep : 0; L.accumulate = f.accumulate;
    for (int j = 0; j < 2; ++j) {
        L.p[j].y = f.prep[j].y; L.p[j].stats = f.prep[j].stats; L.p[j].gamma = f.prep[j].gamma;
        L.p[j].sums = reinterpret_cast<unsigned long long*>(f.prep[j].sums); L.p[j].dgb = reinterpret_cast<unsigned long long*>(f.prep[j].dgb);
    }
    L.H = d.Hi; L.W = d.Wi; L.Cin = d.Cin; L.Cout = d.Cout; L.k = d.kh; L.pad = d.pad; L.B = d.B;
    L.bh = d.Hi * d.Hi <= dz::BM ? d.Hi : dz::BM / d.Hi;
    L.tps = ceil_div(d.Hi, L.bh); L.ntiles = d.Cin / dz::BN;
    const int tiles = d.B * L.tps * L.ntiles, nkb = d.kh * d.kw * d.Cout / dz::BK;
    static const int budget = [] { const char* e = getenv("DBOA_DGRAD_MAX_CTAS"); int v = e ? atoi(e) : 128; return v; }();
    int nz = 1;
    while (nz < 16 && tiles * nz * 2 <= (budget > 2 * tiles ? budget : (2 * tiles < 128 ? 2 * tiles : 128)) && nkb / (nz * 2) >= 2) nz *= 2;
    while (nz > 1 && (nz - 1) * ceil_div(nkb, nz) >= nkb) nz >>= 1;
    L.nz = nz; L.per = ceil_div(nkb, nz);
    L.tabc = d.kh == 1 ? L.per * dz::BK : d.Cout;
    const size_t fixed = 2 * (size_t)(dz::A_TILE + dz::B_TILE) + (size_t)L.tabc * sizeof(float) + 4096 + 1024;
    int D = L.per < dz::DMAX ? L.per : dz::DMAX;
    while (D > 1 && fixed + (size_t)D * dz::SLOT > 227 * 1024) --D;
    L.D = D;
    const size_t smem = fixed + (size_t)D * dz::SLOT;
    const CUtensorMap* tmdz = static_cast<const CUtensorMap*>(tma_act_map(f.dz, d.B, d.Hi, d.Wi, d.Cout, d.Wi, L.bh, false, 1));
    const CUtensorMap* tmy = static_cast<const CUtensorMap*>(tma_act_map(f.y_c, d.B, d.Hi, d.Wi, d.Cout, d.Wi, L.bh, false, 1));
    const CUtensorMap* tmw = static_cast<const CUtensorMap*>(tma_weight_map_mn(f.w, d.kh * d.kw * d.Cin, d.Cout));
    if (!tmdz || !tmy || !tmw) return DBOA_ERR_CUDA;
    return launch_ex(dz::dgrad_wide_kernel, dim3(tiles * nz), dim3(dz::NT), smem, st, dim3(nz, 1, 1), pdl, L, *tmdz, *tmy, *tmw);
}

int gn_bwd_prep(const float* dA, const float* mask, float* out, const DgradPrep& p, int B, int HW, int C, cudaStream_t st) {
    if (C % 4 != 0 || C / 4 > 512) return DBOA_ERR_SHAPE;
    dz::Prep pp;
    pp.y = p.y; pp.stats = p.stats; pp.gamma = p.gamma; pp.sums = reinterpret_cast<unsigned long long*>(p.sums); pp.dgb = reinterpret_cast<unsigned long long*>(p.dgb);
    const int rows_per = HW >= 784 ? 28 : (HW >= 196 ? 14 : 7);
    return launch_ex(dz::gn_bwd_prep_kernel, dim3(ceil_div(HW, rows_per), B), dim3(256), 0, st, dim3(1, 1, 1), true, dA, mask, out, pp, HW, C, rows_per);
}

int gn_dgb_finish(const GnFinishItem* items_dev, int n_items, const float* dgb, float* G, cudaStream_t st) {
    return launch_ex(dz::gn_dgb_finish_kernel, dim3(n_items), dim3(256), 0, st, dim3(1, 1, 1), true, items_dev, reinterpret_cast<const long long*>(dgb), G);
}

}  // namespace dboa
