// Fused forward convolution of the HMR backbone on tcgen05 tensor cores:
//
//      y = conv( T(x) , W )          T = on-load transform of the operand (GroupNorm apply / residual / ReLU)
//      + GroupNorm statistics of y   (per (sample, group) partial (count, mean, M2) of every output tile, in the epilogue)
//
// Replaces, per backbone layer, the pair  `nn.Conv2d` -> `nn.GroupNorm(4, C)` (+ ReLU, + residual add) of reference
// model/hmr.py:29-60 (Bottleneck.forward) / :14-18 (gn_helper): round 1 ran it as two launches per layer (conv_tc.cu, then
// groupnorm.cu: 121 launches per forward, the GroupNorm launches 22 % of the kernel time and one activation round trip each).
// Here the normalisation never gets its own launch:
//   * the PRODUCING convolution leaves, next to its raw output y, one (count, mean, M2) triple per (tile, sample, group);
//   * the CONSUMING convolution merges those triples in its prologue (every producer warp, fixed order: deterministic),
//     and applies   a = relu( (y - mean) * (rstd * gamma) + beta  [+ residual] )   to the operand while it is in registers on
//     its way from global memory to the shared-memory stage.  One designated CTA column also writes `a` (and the
//     (mean, rstd) pair) to the tape, because the hand-written backward reads both; that store is off the critical path.
// Transform modes (template parameter):
//   0  x is an activation that already exists in memory (max-pool output)
//   1  a = relu(gn(y))                                 input of conv2 / conv3 of a bottleneck
//   2  a = relu(gn(y3) + res)                          input of the next block's conv1 (identity shortcut; `res` materialised)
//   3  a = relu(gn(y3) + gn_d(yd))                     same after a block with a down-sampling shortcut (two GroupNorms)
//
// Operand feed:
//   * WEIGHTS come through TMA (`cp.async.bulk.tensor.2d`, 128B-swizzled 64 x 32 fp32 boxes) into a ring of up to 14
//     shared-memory slots.  They do not depend on the previous kernel, so with programmatic dependent launch the whole ring
//     is in flight BEFORE `griddepcontrol.wait`: at batch 1 weights are 55 % of the bytes of a forward and they now stream
//     during the predecessor's tail.  A dedicated warp re-arms ring slots as `tcgen05.commit` releases them.
//   * TF32x3: every fp32 value v is used as hi = v with the low 13 mantissa bits cleared and lo = v - hi; the tensor
//     core accumulates Ah*Bh + Ah*Bl + Al*Bh in fp32 (TMEM).  For the weights the producer warps turn a landed raw tile
//     into (hi in place, lo in a second tile) with one elementwise shared-memory pass -- layout agnostic, so the tile can
//     stay in whatever swizzle TMA wrote.
//   * ACTIVATIONS keep the register path (they need the transform): 8 producer warps, two k-blocks in flight.
// One CTA per SM (<= 227 KB of shared memory), 128 (pixels) x 64 (channels) output tile, split-K across a thread-block
// cluster with a DSMEM reduction (as conv_tc.cu).  Tiles never straddle samples (tile -> (sample, pixel tile)), so the
// sample of a CTA is fixed and the statistics need no segmented reductions.
// A launch can carry TWO problems with the same reduction length (conv1 and the down-sampling 1x1 conv of a block read
// the same transformed input): the cluster index picks the problem.
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace dboa {
namespace fz {

constexpr int BM = 128, BN = 64, BK = 32, SA = 2;
constexpr int NPW = 8, NPROD = NPW * 32, W_MMA = 8, W_TMA = 9, NT = 320;
constexpr int NB_MAX = 14;                                // weight ring slots (8 KB each)
constexpr uint32_t CORE_BYTES = 128, GROUP_BYTES = (BK / 4) * CORE_BYTES;
constexpr uint32_t A_TILE = BM * BK * 4, B_TILE = BN * BK * 4;
constexpr int RED_LD = BN + 4;
constexpr float GN_EPS = 1e-5f;

struct Problem {
    const float* x;          // operand source: raw conv output (modes 1-3) or activation (mode 0), NHWC [B][Hi][Wi][Cin]
    const float* res;        // mode 2: residual activation; mode 3: raw output of the down-sampling conv (same shape as x)
    float* a_out;            // materialised transformed operand (same shape as x) or NULL
    float* stats_out;        // (mean, rstd) [B][4][2] of x's GroupNorm or NULL
    float* stats2_out;       // mode 3: same for the second GroupNorm
    const float4* part_in;   // modes 1-3: partial statistics of x   [B][4][S_in]
    const float4* part2_in;  // mode 3: partial statistics of res
    const float* gamma; const float* beta; const float* gamma2; const float* beta2;
    float* y;                // output [B][Ho][Wo][Cout]
    float4* part_out;        // [B][4][S_out]
    int S_in, S2_in;
    int Hi, Wi, Cin, Ho, Wo, Cout, k, stride, pad;
    int tps, ntiles, nclusters;
};
struct Launch {
    Problem p[2];
    int nprob, nz, per, NB, tabc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle (activations: written by the producer threads); see conv_tc.cu
__device__ __forceinline__ uint64_t desc_ns(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((CORE_BYTES >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((GROUP_BYTES >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// K-major, 128-byte swizzle (weights: written by TMA): rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1),
// version 1 at [46,48), layout_type SWIZZLE_128B = 2 at [61,64)  (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((1024 >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
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
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the device.
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
        else if (clock64() - t0 > 4000000000ll) __trap();
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

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void split_store4(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 v) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// (count, mean, M2) merge (Chan et al.); the right-hand side may be empty
__device__ __forceinline__ void chan_merge(float& n, float& m, float& M2, float nb, float mb, float Mb) {
    if (nb > 0.f) {
        const float nn = n + nb, d = mb - m, f = nb / nn;
        M2 = M2 + Mb + d * d * n * f;
        m = m + d * f;
        n = nn;
    }
}
__device__ __forceinline__ float sel4(const float (&v)[4], int g) { return g == 0 ? v[0] : (g == 1 ? v[1] : (g == 2 ? v[2] : v[3])); }

// Every lane of the warp ends up with (mean, rstd) of the 4 groups of sample b: lanes stride over the S partial slots,
// merge sequentially, then a shuffle-down tree into lane 0 and a broadcast (identical in every warp of every CTA).
__device__ __forceinline__ void merge_stats(const float4* __restrict__ part, int b, int S, int lane, float (&mean)[4], float (&rstd)[4]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4* src = part + (size_t)(b * 4 + g) * S;
        float n = 0.f, m = 0.f, M2 = 0.f;
        for (int s = lane; s < S; s += 32) {
            const float4 q = __ldcg(src + s);
            chan_merge(n, m, M2, q.x, q.y, q.z);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float nb = __shfl_down_sync(0xffffffffu, n, o), mb = __shfl_down_sync(0xffffffffu, m, o), Mb = __shfl_down_sync(0xffffffffu, M2, o);
            chan_merge(n, m, M2, nb, mb, Mb);
        }
        n = __shfl_sync(0xffffffffu, n, 0); m = __shfl_sync(0xffffffffu, m, 0); M2 = __shfl_sync(0xffffffffu, M2, 0);
        mean[g] = m;
        rstd[g] = 1.0f / sqrtf(M2 / n + GN_EPS);
    }
}

#define PF(field) (second ? L.p[1].field : L.p[0].field)

template <int MODE>
__global__ void __launch_bounds__(NT, 1) conv_fused_kernel(const __grid_constant__ Launch L, const __grid_constant__ CUtensorMap tm0,
                                                           const __grid_constant__ CUtensorMap tm1) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nz = L.nz, NB = L.NB;
    const int cidx = blockIdx.x / nz, rank = blockIdx.x - cidx * nz;
    const bool second = L.nprob > 1 && cidx >= L.p[0].nclusters;
    const int tix = second ? cidx - L.p[0].nclusters : cidx;
    const CUtensorMap* tm = second ? &tm1 : &tm0;
    const int Hi = PF(Hi), Wi = PF(Wi), Cin = PF(Cin), Ho = PF(Ho), Wo = PF(Wo), Cout = PF(Cout), ks = PF(k), stride = PF(stride), pad = PF(pad);
    const int tps = PF(tps), ntiles = PF(ntiles);
    const int nt = tix % ntiles, bm = tix / ntiles, mt = bm % tps, b = bm / tps;
    const int HWo = Ho * Wo, m0 = mt * BM, n0 = nt * BN;
    const int rows_valid = min(BM, HWo - m0);
    const int nkb_total = (ks * ks * Cin) / BK;
    const int kb_begin = rank * L.per;
    const int nkb = max(0, min(L.per, nkb_total - kb_begin));

    // ---- shared memory carve-up (1024-byte aligned: the swizzled tiles need it)
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = base;                                           // NB x 8 KB raw weight tiles (TMA), hi in place after the split pass
    uint8_t* a_hi = ring + (size_t)NB * B_TILE;                     // SA x 16 KB
    uint8_t* a_lo = a_hi + SA * A_TILE;                             // SA x 16 KB
    uint8_t* b_lo = a_lo + SA * A_TILE;                             // SA x 8 KB
    float* tab_g = reinterpret_cast<float*>(b_lo + SA * B_TILE);    // gamma / beta (/ second GroupNorm) of this CTA's channel range
    float* tab_b = tab_g + L.tabc;
    float* tab_g2 = tab_b + L.tabc;
    float* tab_b2 = tab_g2 + L.tabc;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tab_g + 4 * (size_t)L.tabc);
    uint64_t* full = bars;                 // [SA]   producers -> MMA issuer
    uint64_t* empty = bars + SA;           // [SA]   tcgen05.commit -> producers
    uint64_t* wfull = bars + 2 * SA;       // [NB_MAX] TMA -> producers
    uint64_t* wempty = wfull + NB_MAX;     // [NB_MAX] tcgen05.commit -> TMA warp
    uint64_t* done = wempty + NB_MAX;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    float4* wpart = reinterpret_cast<float4*>(tmem_slot + 4);       // [NPW][4] per-warp statistics of the epilogue
    float* red = reinterpret_cast<float*>(a_hi);                    // 128 x RED_LD fp32 partial tile, over the A stages after the last MMA

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(&full[s], NPW); mbar_init(&empty[s], 1); }
        for (int s = 0; s < NB_MAX; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // GroupNorm affine parameters of the operand (parameters: safe to read before the dependency wait)
    const int tc0 = ks == 1 ? kb_begin * BK : 0, tcn = ks == 1 ? nkb * BK : Cin;
    if (MODE >= 1) {
        const float* ga = PF(gamma); const float* be = PF(beta);
        for (int i = tid; i < tcn; i += NT) { tab_g[i] = __ldg(ga + tc0 + i); tab_b[i] = __ldg(be + tc0 + i); }
        if (MODE == 3) {
            const float* ga2 = PF(gamma2); const float* be2 = PF(beta2);
            for (int i = tid; i < tcn; i += NT) { tab_g2[i] = __ldg(ga2 + tc0 + i); tab_b2[i] = __ldg(be2 + tc0 + i); }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == W_TMA) {
        // =====================================================================================
        // weight feed: one thread, TMA boxes of 64 output channels x 32 reduction elements
        // =====================================================================================
        if (lane == 0 && nkb > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
            for (int it = 0; it < nkb; ++it) {
                const int r = it % NB;
                if (it >= NB) mbar_wait(&wempty[r], (uint32_t)(((it / NB) - 1) & 1));      // the MMAs that read the slot are done
                mbar_expect_tx(&wfull[r], B_TILE);
                tma_load_2d(smem_u32(ring + (size_t)r * B_TILE), tm, (kb_begin + it) * BK, n0, &wfull[r]);
            }
        }
        pdl_wait();
        pdl_trigger();
    } else if (warp == W_MMA) {
        // =====================================================================================
        // MMA issuer: one thread, 12 x tcgen05.mma per k-block (4 k-steps of 8 x {Ah*Bh, Ah*Bl, Al*Bh})
        // =====================================================================================
        if (lane == 0 && nkb > 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint64_t dah[SA], dal[SA], dbl[SA];
#pragma unroll
            for (int s = 0; s < SA; ++s) {
                dah[s] = desc_ns(smem_u32(a_hi + s * A_TILE)); dal[s] = desc_ns(smem_u32(a_lo + s * A_TILE));
                dbl[s] = desc_sw128(smem_u32(b_lo + s * B_TILE));
            }
            const uint64_t dring = desc_sw128(smem_u32(ring));
            constexpr uint64_t KSTEP_A = (2 * CORE_BYTES) >> 4;     // 8 tf32 = two 16-byte k-chunks (core matrices 128 B apart)
            constexpr uint64_t KSTEP_B = 32 >> 4;                   // 8 tf32 = 32 bytes inside the 128-byte swizzle row
            for (int it0 = 0; it0 < nkb; it0 += SA) {
#pragma unroll
                for (int s = 0; s < SA; ++s) {
                    const int it = it0 + s;
                    if (it < nkb) {
                        const int r = it % NB;
                        mbar_wait(&full[s], (uint32_t)((it / SA) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t dbh = dring + (uint64_t)(r * (int)(B_TILE >> 4));
#pragma unroll
                        for (int kk = 0; kk < BK / 8; ++kk) {
                            mma_tf32(tmem_d, dah[s] + kk * KSTEP_A, dbh + kk * KSTEP_B, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                            mma_tf32(tmem_d, dah[s] + kk * KSTEP_A, dbl[s] + kk * KSTEP_B, idesc, 1u);
                            mma_tf32(tmem_d, dal[s] + kk * KSTEP_A, dbh + kk * KSTEP_B, idesc, 1u);
                        }
                        umma_commit(&empty[s]);
                        umma_commit(&wempty[r]);
                    }
                }
            }
            umma_commit(done);
        }
        pdl_wait();
        pdl_trigger();
    } else {
        // =====================================================================================
        // producers: global -> registers (two k-blocks ahead) -> transform -> hi/lo -> shared memory; weight split pass
        // Warp w owns the 8-row groups 2w, 2w+1 of the pixel tile; lane = (row lr8, 16-byte chunk pair cpair); register
        // slot q*2 + h holds row group 2w + q, k-chunk h*4 + cpair (see conv_tc.cu).
        // =====================================================================================
        const int lr8 = lane & 7, cpair = lane >> 3;
        const float* xb = PF(x) + (size_t)b * Hi * Wi * Cin;
        const float* rb = (MODE >= 2) ? PF(res) + (size_t)b * Hi * Wi * Cin : nullptr;
        float* ab = PF(a_out) != nullptr && nt == 0 ? PF(a_out) + (size_t)b * Hi * Wi * Cin : nullptr;
        bool avalid[2];
        int ph[2], pw[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int p = m0 + (warp * 2 + q) * 8 + lr8;
            avalid[q] = p < HWo;
            const int h = p / Wo, w_ = p - h * Wo;
            ph[q] = h * stride - pad; pw[q] = w_ * stride - pad;
        }
        const uint32_t a_off = (uint32_t)(warp * 2) * GROUP_BYTES + (uint32_t)cpair * CORE_BYTES + (uint32_t)lr8 * 16;
        struct Cursor { int r, s, c; };
        Cursor cur;
        {
            const int k0 = kb_begin * BK, tap = k0 / Cin;
            cur.c = k0 - tap * Cin; cur.r = tap / ks; cur.s = tap - cur.r * ks;
        }
        auto advance = [&](Cursor& c) { c.c += BK; if (c.c >= Cin) { c.c = 0; if (++c.s == ks) { c.s = 0; ++c.r; } } };
        int lgw = 0;
        while ((4 << lgw) < Cin) ++lgw;                         // channels per group = Cin / 4 = 1 << lgw

        float4 ra[SA][4], rr[SA][4];
        int mc[SA];                                             // channel base of the k-block held in the slot
        unsigned mflag[SA];                                     // bits 0..1: row q in bounds; bit 2: tap designated for the a_out store
        size_t moff[SA][2];                                     // element offset of (row q, channel base) in x / res / a_out
        auto fetch = [&](int f) {
            unsigned fl = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int hi = ph[q] + cur.r, wi = pw[q] + cur.s;
                const bool inb = avalid[q] && (unsigned)hi < (unsigned)Hi && (unsigned)wi < (unsigned)Wi;
                const size_t off = ((size_t)hi * Wi + wi) * Cin + cur.c + cpair * 4;
                moff[f][q] = off;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                ra[f][q * 2] = inb ? ldg4(xb + off) : z;
                ra[f][q * 2 + 1] = inb ? ldg4(xb + off + 16) : z;
                if (MODE >= 2) {
                    rr[f][q * 2] = inb ? ldg4(rb + off) : z;
                    rr[f][q * 2 + 1] = inb ? ldg4(rb + off + 16) : z;
                }
                fl |= inb ? (1u << q) : 0u;
            }
            const bool desig = ks == 1 ? true : (stride == 1 ? (cur.r == 1 && cur.s == 1) : (cur.r >= 1 && cur.s >= 1));
            mflag[f] = fl | (desig ? 4u : 0u);
            mc[f] = cur.c;
            advance(cur);
        };

        pdl_wait();          // activations and statistics of the previous layer exist from here on
        pdl_trigger();
#pragma unroll
        for (int f = 0; f < SA; ++f)
            if (f < nkb) fetch(f);
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f}, mean2[4] = {0.f, 0.f, 0.f, 0.f}, rstd2[4] = {1.f, 1.f, 1.f, 1.f};
        if (MODE >= 1) {
            merge_stats(PF(part_in), b, PF(S_in), lane, mean, rstd);
            if (MODE == 3) merge_stats(PF(part2_in), b, PF(S2_in), lane, mean2, rstd2);
            if (mt == 0 && nt == 0 && rank == 0 && warp == 0 && lane == 0) {
                float* so = PF(stats_out);
                if (so != nullptr)
                    for (int g = 0; g < 4; ++g) { so[(b * 4 + g) * 2] = mean[g]; so[(b * 4 + g) * 2 + 1] = rstd[g]; }
                if (MODE == 3) {
                    float* so2 = PF(stats2_out);
                    if (so2 != nullptr)
                        for (int g = 0; g < 4; ++g) { so2[(b * 4 + g) * 2] = mean2[g]; so2[(b * 4 + g) * 2 + 1] = rstd2[g]; }
                }
            }
        }
        auto transform = [&](float4 v, float4 r, int c_abs, bool inb) {
            if (MODE == 0) return v;
            const int g = c_abs >> lgw, ti = c_abs - tc0;
            const float mu = sel4(mean, g), rs = sel4(rstd, g);
            const float4 ga = *reinterpret_cast<const float4*>(tab_g + ti), be = *reinterpret_cast<const float4*>(tab_b + ti);
            float4 o;
            o.x = (v.x - mu) * (rs * ga.x) + be.x; o.y = (v.y - mu) * (rs * ga.y) + be.y;
            o.z = (v.z - mu) * (rs * ga.z) + be.z; o.w = (v.w - mu) * (rs * ga.w) + be.w;
            if (MODE == 2) { o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
            if (MODE == 3) {
                const float mu2 = sel4(mean2, g), rs2 = sel4(rstd2, g);
                const float4 ga2 = *reinterpret_cast<const float4*>(tab_g2 + ti), be2 = *reinterpret_cast<const float4*>(tab_b2 + ti);
                o.x += (r.x - mu2) * (rs2 * ga2.x) + be2.x; o.y += (r.y - mu2) * (rs2 * ga2.y) + be2.y;
                o.z += (r.z - mu2) * (rs2 * ga2.z) + be2.z; o.w += (r.w - mu2) * (rs2 * ga2.w) + be2.w;
            }
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            if (!inb) o = make_float4(0.f, 0.f, 0.f, 0.f);      // zero padding lives in the ACTIVATION domain
            return o;
        };

        for (int it0 = 0; it0 < nkb; it0 += SA) {
#pragma unroll
            for (int f = 0; f < SA; ++f) {
                const int it = it0 + f;
                if (it < nkb) {
                    const int s = f, r = it % NB;
                    if (it >= SA) mbar_wait(&empty[s], (uint32_t)(((it / SA) - 1) & 1));
                    // ---- activations: transform, optional tape store, hi/lo split into the stage
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int q = j >> 1, h = j & 1;
                        const bool inb = (mflag[f] >> q) & 1u;
                        const float4 o = transform(ra[f][j], MODE >= 2 ? rr[f][j] : ra[f][j], mc[f] + cpair * 4 + h * 16, inb);
                        if (MODE >= 1 && ab != nullptr && inb && (mflag[f] & 4u)) *reinterpret_cast<float4*>(ab + moff[f][q] + h * 16) = o;
                        split_store4(a_hi + s * A_TILE, a_lo + s * A_TILE, a_off + q * GROUP_BYTES + h * 4 * CORE_BYTES, o);
                    }
                    // ---- weights: raw tile (TMA) -> hi in place, lo into the stage (same byte offsets: swizzle agnostic)
                    mbar_wait(&wfull[r], (uint32_t)((it / NB) & 1));
                    {
                        float4* raw = reinterpret_cast<float4*>(ring + (size_t)r * B_TILE);
                        float4* lo = reinterpret_cast<float4*>(b_lo + s * B_TILE);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int idx = tid + j * NPROD;
                            const float4 v = raw[idx];
                            const float4 hh = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                            raw[idx] = hh;
                            lo[idx] = make_float4(v.x - hh.x, v.y - hh.y, v.z - hh.z, v.w - hh.w);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async-proxy (UMMA) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    if (it + SA < nkb) fetch(f);
                }
            }
        }
    }
    if (nkb > 0) mbar_wait(done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM -> shared memory (thread = row), split-K reduction over the cluster, output + statistics
    if (warp < NPW) {
        const int q4 = warp & 3, cgp = warp >> 2;
        uint32_t r[32];
        if (nkb > 0) {
            const uint32_t taddr = tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(cgp * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
        float* dstrow = red + (q4 * 32 + lane) * RED_LD + cgp * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dstrow + q * 4) = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]),
                                                                     __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
    }
    cg::cluster_group cluster = cg::this_cluster();
    if (nz == 1) __syncthreads(); else cluster.sync();

    // every CTA of the cluster owns the band of 128/nz rows `rank`; producer thread t holds up to 8 float4 of it
    const int rows_per = BM / nz, items = rows_per * (BN / 4);
    const int gw = Cout >> 2;                                   // channels per GroupNorm group of the OUTPUT
    const int gpt = gw >= BN ? 1 : BN / gw;                     // groups per 64-channel tile (1, 2 or 4)
    const int lpg = 16 / gpt;                                   // lanes (float4 columns) per group
    float sn = 0.f, smean = 0.f, sM2 = 0.f;
    if (warp < NPW) {
        float* Y = PF(y) + (size_t)b * HWo * Cout;
        float4 vals[8];
        int cnt = 0;
        float ssum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int v = tid + i * NPROD;
            vals[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < items) {
                const int lr = rank * rows_per + (v >> 4), c4 = (v & 15) * 4;
                if (lr < rows_valid) {
                    float4 acc;
                    if (nz == 1) {
                        acc = *reinterpret_cast<const float4*>(red + lr * RED_LD + c4);
                    } else {
                        acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int zb = 0; zb < 16; zb += 8) {          // 8 remote loads in flight, added in the order z = 0..nz-1
                            if (zb < nz) {
                                float4 q[8];
#pragma unroll
                                for (int z = 0; z < 8; ++z)
                                    if (zb + z < nz) q[z] = *reinterpret_cast<const float4*>(cluster.map_shared_rank(red, zb + z) + lr * RED_LD + c4);
#pragma unroll
                                for (int z = 0; z < 8; ++z)
                                    if (zb + z < nz) { acc.x += q[z].x; acc.y += q[z].y; acc.z += q[z].z; acc.w += q[z].w; }
                            }
                        }
                    }
                    *reinterpret_cast<float4*>(Y + (size_t)(m0 + lr) * Cout + n0 + c4) = acc;
                    vals[i] = acc;
                    ssum += (acc.x + acc.y) + (acc.z + acc.w);
                    cnt |= 1 << i;
                }
            }
        }
        // thread-local (count, mean, M2), then a fixed shuffle-down tree over the lanes of the same group
        sn = 4.f * (float)__popc(cnt);
        smean = cnt ? ssum / sn : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((cnt >> i) & 1) {
                const float a = vals[i].x - smean, c = vals[i].y - smean, e = vals[i].z - smean, f = vals[i].w - smean;
                sM2 += (a * a + c * c) + (e * e + f * f);
            }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            if (o == 16 || o < lpg) {
                const float nb = __shfl_down_sync(0xffffffffu, sn, o), mb = __shfl_down_sync(0xffffffffu, smean, o),
                            Mb = __shfl_down_sync(0xffffffffu, sM2, o);
                chan_merge(sn, smean, sM2, nb, mb, Mb);
            }
        }
        if (lane < 16 && (lane % lpg) == 0) wpart[warp * 4 + lane / lpg] = make_float4(sn, smean, sM2, 0.f);
    }
    __syncthreads();
    if (warp == 0 && lane < gpt) {
        float n = 0.f, m = 0.f, M2 = 0.f;
        for (int w = 0; w < NPW; ++w) { const float4 q = wpart[w * 4 + lane]; chan_merge(n, m, M2, q.x, q.y, q.z); }
        const int ntg = gw >= BN ? gw / BN : 1;
        const int g = gw >= BN ? (nt * BN) / gw : nt * gpt + lane;
        const int nig = gw >= BN ? nt % ntg : 0;
        const int S_out = tps * ntg * nz;
        const int slot = (mt * ntg + nig) * nz + rank;
        PF(part_out)[(size_t)(b * 4 + g) * S_out + slot] = make_float4(n, m, M2, 0.f);
    }
    if (nz > 1) cluster.sync();                                  // peers may still be reading this CTA's partial tile
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// tensor map of a weight matrix [Cout][K] (K contiguous): boxes of 64 rows x 32 floats, 128-byte swizzle
static const CUtensorMap* weight_map(const float* w, int K, int Cout) {
    static std::map<std::tuple<const float*, int, int>, CUtensorMap*> cache;
    auto key = std::make_tuple(w, K, Cout);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return nullptr;
    CUtensorMap* tm = static_cast<CUtensorMap*>(aligned_alloc(64, sizeof(CUtensorMap)));
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { free(tm); return nullptr; }
    cache[key] = tm;
    return tm;
}

}  // namespace fz

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        g_num_sms = n > 0 ? n : 148;
    }
    return g_num_sms;
}

int conv_fused_slots(const FusedConv& d, int nz) {
    const int tps = ceil_div(d.Ho * d.Ho, fz::BM), gw = d.Cout / 4;
    return tps * (gw >= fz::BN ? gw / fz::BN : 1) * nz;
}

// Cluster size (K-slices) of a launch: the largest power of two <= 16 that keeps the grid within one wave of one CTA per SM
// and leaves every rank at least 2 k-blocks.
int conv_fused_plan(const FusedConv* d, int nprob, int B) {
    int tiles = 0;
    for (int i = 0; i < nprob; ++i) tiles += B * ceil_div(d[i].Ho * d[i].Ho, fz::BM) * (d[i].Cout / fz::BN);
    const int nkb = d[0].k * d[0].k * d[0].Cin / fz::BK;
    int nz = 1;
    while (nz < 16 && tiles * nz * 2 <= num_sms() && nkb / (nz * 2) >= 2) nz *= 2;
    while (nz > 1 && (nz - 1) * ceil_div(nkb, nz) >= nkb) nz >>= 1;
    return nz;
}

bool conv_fused_ok(const FusedConv& d) {
    return d.Cin % 64 == 0 && d.Cout % 64 == 0 && (d.k == 1 || d.k == 3) && (d.mode >= 0 && d.mode <= 3);
}

int conv_fused_launch(const FusedConv* d, int nprob, int B, int nz, cudaStream_t st, bool pdl) {
    if (nprob < 1 || nprob > 2 || B < 1) return DBOA_ERR_ARG;
    fz::Launch L;
    memset(&L, 0, sizeof L);
    const CUtensorMap* tms[2] = {nullptr, nullptr};
    const int K0 = d[0].k * d[0].k * d[0].Cin;
    int total = 0, tabc = 0;
    const int nkb = K0 / fz::BK, per = ceil_div(nkb, nz);
    if (nz < 1 || nz > 16 || (nz & (nz - 1)) || (nz - 1) * per >= nkb) return DBOA_ERR_ARG;
    for (int i = 0; i < nprob; ++i) {
        const FusedConv& c = d[i];
        if (!conv_fused_ok(c) || c.mode != d[0].mode || c.k * c.k * c.Cin != K0) return DBOA_ERR_UNSUPPORTED;
        if (c.a_out != nullptr && c.k == 1 && c.stride != 1) return DBOA_ERR_ARG;      // a strided 1x1 does not visit every input pixel
        fz::Problem& p = L.p[i];
        p.x = c.x; p.res = c.res; p.a_out = c.a_out; p.stats_out = c.stats_out; p.stats2_out = c.stats2_out;
        p.part_in = reinterpret_cast<const float4*>(c.part_in); p.part2_in = reinterpret_cast<const float4*>(c.part2_in);
        p.gamma = c.gamma; p.beta = c.beta; p.gamma2 = c.gamma2; p.beta2 = c.beta2;
        p.y = c.y; p.part_out = reinterpret_cast<float4*>(c.part_out);
        p.S_in = c.S_in; p.S2_in = c.S2_in;
        p.Hi = c.Hi; p.Wi = c.Hi; p.Cin = c.Cin; p.Ho = c.Ho; p.Wo = c.Ho; p.Cout = c.Cout; p.k = c.k; p.stride = c.stride; p.pad = c.pad;
        p.tps = ceil_div(c.Ho * c.Ho, fz::BM); p.ntiles = c.Cout / fz::BN; p.nclusters = B * p.tps * p.ntiles;
        total += p.nclusters;
        if (c.mode >= 1) tabc = tabc > (c.k == 1 ? per * fz::BK : c.Cin) ? tabc : (c.k == 1 ? per * fz::BK : c.Cin);
        tms[i] = fz::weight_map(c.w, K0, c.Cout);
        if (tms[i] == nullptr) return DBOA_ERR_CUDA;
    }
    if (nprob == 1) tms[1] = tms[0];
    L.nprob = nprob; L.nz = nz; L.per = per; L.tabc = tabc;
    const size_t fixed = (size_t)fz::SA * 2 * fz::A_TILE + (size_t)fz::SA * fz::B_TILE + 4 * (size_t)tabc * sizeof(float) + 2048 + 1024;
    int NB = per < fz::NB_MAX ? per : fz::NB_MAX;
    while (NB > 2 && fixed + (size_t)NB * fz::B_TILE > 227 * 1024) --NB;
    L.NB = NB;
    const size_t smem = fixed + (size_t)NB * fz::B_TILE;
    const dim3 grid(total * nz), block(fz::NT), cl(nz, 1, 1);
    switch (d[0].mode) {
        case 0: return launch_ex(fz::conv_fused_kernel<0>, grid, block, smem, st, cl, pdl, L, *tms[0], *tms[1]);
        case 1: return launch_ex(fz::conv_fused_kernel<1>, grid, block, smem, st, cl, pdl, L, *tms[0], *tms[1]);
        case 2: return launch_ex(fz::conv_fused_kernel<2>, grid, block, smem, st, cl, pdl, L, *tms[0], *tms[1]);
        default: return launch_ex(fz::conv_fused_kernel<3>, grid, block, smem, st, cl, pdl, L, *tms[0], *tms[1]);
    }
}

// -------------------------------------------------------------------------------------------------
// Last GroupNorm of the backbone fused with the 7x7 average pool (reference model/hmr.py:57-60 of layer4.2 and :156-157):
//   a = relu(gn(y3) + res)  -> tape;  xf[b][c] = mean_p a[b][p][c]  -> the three regressor input rows.
// grid (C / 128, B), 128 threads = one channel each... 4 channels per thread, 32 threads x 4 pixel lanes.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_res_avgpool_kernel(const float* __restrict__ y, const float* __restrict__ res, const float4* __restrict__ part,
                                                              int S, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float* __restrict__ a_out, float* __restrict__ stats_out, float* __restrict__ out,
                                                              int HW, int C, int ld, int ncopy, size_t copy_stride) {
    __shared__ float4 acc[8][32];
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mean[4], rstd[4];
    fz::merge_stats(part, b, S, lane, mean, rstd);
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats_out != nullptr)
        for (int g = 0; g < 4; ++g) { stats_out[(b * 4 + g) * 2] = mean[g]; stats_out[(b * 4 + g) * 2 + 1] = rstd[g]; }
    const int c = blockIdx.x * 128 + lane * 4, g = c / (C >> 2);
    const float mu = fz::sel4(mean, g), rs = fz::sel4(rstd, g);
    const float4 ga = ldg4(gamma + c), be = ldg4(beta + c);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = warp; p < HW; p += 8) {
        const size_t off = ((size_t)b * HW + p) * C + c;
        const float4 v = ldg4(y + off), r = ldg4(res + off);
        float4 o;
        o.x = (v.x - mu) * (rs * ga.x) + be.x; o.y = (v.y - mu) * (rs * ga.y) + be.y;
        o.z = (v.z - mu) * (rs * ga.z) + be.z; o.w = (v.w - mu) * (rs * ga.w) + be.w;
        o.x = fmaxf(o.x + r.x, 0.f); o.y = fmaxf(o.y + r.y, 0.f); o.z = fmaxf(o.z + r.z, 0.f); o.w = fmaxf(o.w + r.w, 0.f);
        *reinterpret_cast<float4*>(a_out + off) = o;
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    acc[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        float4 t = acc[0][lane];
        for (int w = 1; w < 8; ++w) { const float4 q = acc[w][lane]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
        const float hw = (float)HW;
        t.x /= hw; t.y /= hw; t.z /= hw; t.w /= hw;
        for (int k = 0; k < ncopy; ++k) *reinterpret_cast<float4*>(out + k * copy_stride + (size_t)b * ld + c) = t;
    }
}

int gn_res_avgpool(const float* y, const float* res, const float* part, int S, const float* gamma, const float* beta, float* a_out,
                   float* stats_out, float* out, int B, int HW, int C, int ld, int ncopy, size_t copy_stride, cudaStream_t st) {
    if (C % 128 != 0 || ld % 4 != 0 || copy_stride % 4 != 0) return DBOA_ERR_SHAPE;
    return launch_ex(gn_res_avgpool_kernel, dim3(C / 128, B), dim3(256), 0, st, dim3(1, 1, 1), true, y, res, reinterpret_cast<const float4*>(part), S,
                     gamma, beta, a_out, stats_out, out, HW, C, ld, ncopy, copy_stride);
}

}  // namespace dboa
