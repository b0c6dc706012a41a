// tcgen05 (5th-generation tensor core) implicit-GEMM convolution: forward, data gradient and weight gradient of
// every backbone convolution with Cin % 64 == 0 and Cout % 64 == 0 (all but the 7x7 stem).
//
//   FWD   : Y [m=(b,ho,wo)][n=co]      = sum_{k=(tap,ci)} Xcol[m][k]  * W[co][k]
//   DGRAD : dX[m=(b,hi,wi)][n=ci]   (+)= sum_{k=(tap,co)} dYcol[m][k] * W[co][tap][ci]
//   WGRAD : dW[m=co][n=(tap,ci)]      += sum_{k=pixel}    dY[pix][co] * Xcol[pix][(tap,ci)]
// activations NHWC, weights [Cout][kh][kw][Cin] (row pitch == kh*kw*Cin).
//
// Precision: `tcgen05.mma.kind::tf32` with a 3-term split.  Every operand value x is split by the loader threads
// into hi = x with the low 13 mantissa bits cleared (exactly a TF32 number, so the tensor core's own fp32->tf32
// conversion is the identity) and lo = x - hi (exact in fp32); the accumulator receives
//   Ah*Bh + Ah*Bl + Al*Bh      (the dropped Al*Bl term is ~2^-22 relative)
// in fp32 inside TMEM, which keeps the result within ~2e-6 of the exact-fp32 CUDA-core path (conv.cu), against
// which this kernel is validated on the device (tests/test_gpu_tc.py).  SURVEY.md §7 "hard part 1".
//
// Structure (one CTA = 512 threads, one 128 x 64 output tile, 32 reduction elements per k-block).  At batch 1 a CTA
// has 2..9 k-blocks and the kernel is a chain of dependent phases; the first version (128 threads, every thread 12
// loads + their index arithmetic per k-block) spent ~1.5 us per k-block with ONE warp per scheduler and nothing to hide
// instruction latency behind (scripts/kernel_timeline.py).  Hence 16 warps and <= 3 loads per thread and k-block:
//   loaders  : per k-block 2 float4 of the 128-row operand and 1 of the 64-row operand per thread; im2col /
//              transposed-filter / pixel-major gathers computed on the fly with zero fill, issued THREE k-blocks
//              ahead in registers; a warp request covers 8 rows x 64 contiguous bytes.  hi/lo are split in registers
//              and stored into the UMMA canonical K-major no-swizzle layout (8-row x 16-byte core matrices) of one of
//              3 shared-memory stages.  Operands whose memory order is reduction-minor (K-major) are stored with
//              float4 (conflict free); the others are transposed by the store.
//   thread 0 : 12 x tcgen05.mma (4 k-steps of 8 x 3 products) per stage, tcgen05.commit -> mbarrier of the stage.
//   epilogue : tcgen05.ld (32 lanes x 16 columns per warp) -> shared memory -> row-contiguous global stores.  Split-K
//              runs across a thread-block cluster (1,1,nz): each CTA sums a band of 128/nz rows over its peers through
//              distributed shared memory (all remote loads in flight, then added in the fixed order z = 0..nz-1).
// Accumulators live in TMEM (64 columns x 128 lanes), never in registers.
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace dboa {

// 0 = fp32 CUDA cores; 1 = forward on tensor cores; 2 = forward, data and weight gradient on tensor cores;
// 3 = forward and data gradient on tensor cores, weight gradient on CUDA cores (default: fastest per scripts/conv_microbench.py)
static int g_tc_mode = 3;
bool conv_tc_enabled() { return g_tc_mode != 0; }
bool conv_tc_bwd_enabled() { return g_tc_mode >= 2; }
bool conv_tc_wgrad_enabled() { return g_tc_mode == 2; }
void conv_tc_set_enabled(bool on) { g_tc_mode = on ? 2 : 0; }
void conv_tc_set_mode(int mode) { g_tc_mode = mode; }

namespace tc {

// Diagnostic build (-DDBOA_TIMELINE, scripts/kernel_timeline.py): thread 0 of every CTA stamps %globaltimer at the phase
// boundaries of the kernel into a device buffer [cta][16]; compiled out of the product library.
#ifdef DBOA_TIMELINE
__device__ unsigned long long* g_timeline = nullptr;
#define DBOA_TL(i)                                                                                        \
    do {                                                                                                  \
        if (threadIdx.x == 0 && g_timeline != nullptr) {                                                  \
            unsigned long long t_;                                                                        \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                        \
            g_timeline[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (i)] = t_; \
        }                                                                                                 \
    } while (0)
// per-iteration cycle stamps of CTA (0,0,0), thread 0: slot 40000 + it * 8 + j
#define DBOA_TLC(it, j)                                                                                   \
    do {                                                                                                  \
        if (threadIdx.x == 0 && g_timeline != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (it) < 12)   \
            g_timeline[40000 + (it) * 8 + (j)] = (unsigned long long)clock64();                           \
    } while (0)
// same for the MMA-issuer thread (slots 5..7 of the iteration)
#define DBOA_TLM(it, j)                                                                                   \
    do {                                                                                                  \
        if (g_timeline != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (it) < 12)            \
            g_timeline[40000 + (it) * 8 + (j)] = (unsigned long long)clock64();                           \
    } while (0)
#else
#define DBOA_TL(i)
#define DBOA_TLC(it, j)
#define DBOA_TLM(it, j)
#endif

enum { FWD = 0, DGRAD = 1, WGRAD = 2 };
constexpr int BM = 128, BN = 64, BK = 32, STAGES = 2;
constexpr int NPW = 8, NPROD = NPW * 32, NT = NPROD + 32;     // 8 producer warps + 1 MMA-issuer warp
constexpr uint32_t CORE_BYTES = 128;                     // one 8 x 16B core matrix
constexpr uint32_t GROUP_BYTES = (BK / 4) * CORE_BYTES;  // one 8-row group of a stage tile (1024 B)
constexpr uint32_t A_TILE = BM * BK * 4, B_TILE = BN * BK * 4;
constexpr int RED_LD = BN + 4;                           // padded row pitch of the split-K partial tile (bank-conflict free)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout_type=0 (no swizzle)
    // K-major: LBO = byte step between the two 16-byte k-chunks of one MMA, SBO = byte step between 8-row groups
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((CORE_BYTES >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((GROUP_BYTES >> 4) & 0x3FFF) << 32) | (1ull << 46);
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// K-major source: 4 consecutive reduction elements of one row -> one float4 slot of the canonical layout
__device__ __forceinline__ void split_store4(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 v) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}
// row-minor source: 4 consecutive ROWS (row0 % 4 == 0) at one reduction index k -> 4 scalar slots (transposing store)
__device__ __forceinline__ void split_store_t(uint8_t* hi_tile, uint8_t* lo_tile, int row0, int k, float4 v) {
    const uint32_t off = (uint32_t)(row0 >> 3) * GROUP_BYTES + (uint32_t)(k >> 2) * CORE_BYTES + (uint32_t)(row0 & 7) * 16 + (uint32_t)(k & 3) * 4;
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float h = tf32_hi(x[e]);
        *reinterpret_cast<float*>(hi_tile + off + e * 16) = h;
        *reinterpret_cast<float*>(lo_tile + off + e * 16) = x[e] - h;
    }
}

// same, with the 4 element stores issued in the order (e + rot) & 3: when the lanes of a warp are (k & 3, (row0 >> 2) & 1, rot)
// the 32 scalar stores of one instruction fall into 32 different banks (the plain version is an 8-way conflict)
__device__ __forceinline__ void split_store_t_rot(uint8_t* hi_tile, uint8_t* lo_tile, int row0, int k, float4 v, int rot) {
    const uint32_t off = (uint32_t)(row0 >> 3) * GROUP_BYTES + (uint32_t)(k >> 2) * CORE_BYTES + (uint32_t)(row0 & 7) * 16 + (uint32_t)(k & 3) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ee = (e + rot) & 3;
        const float x = ee == 0 ? v.x : (ee == 1 ? v.y : (ee == 2 ? v.z : v.w));
        const float h = tf32_hi(x);
        *reinterpret_cast<float*>(hi_tile + off + ee * 16) = h;
        *reinterpret_cast<float*>(lo_tile + off + ee * 16) = x - h;
    }
}

struct Smem {
    // stage tiles in UMMA canonical layout: [row_group][k_chunk][8 rows][16 B]; after the last MMA the A tiles are
    // re-used as the 128 x RED_LD fp32 partial tile of the cluster split-K reduction
    alignas(128) uint8_t a_hi[STAGES][A_TILE];
    alignas(128) uint8_t a_lo[STAGES][A_TILE];
    alignas(128) uint8_t b_hi[STAGES][B_TILE];
    alignas(128) uint8_t b_lo[STAGES][B_TILE];
    alignas(8) uint64_t full[STAGES];      // producers -> MMA issuer: stage filled (one arrival per producer warp)
    alignas(8) uint64_t empty[STAGES];     // tcgen05.commit -> producers: the MMAs that read the stage have completed
    alignas(8) uint64_t done;              // tcgen05.commit after the last MMA -> epilogue
    uint32_t tmem_base;
};

// P: the operand that pairs with the weights in FWD/DGRAD (x or dy); Q: weights (FWD/DGRAD) or x (WGRAD, with P = dy)
template <int MODE>
__global__ void __launch_bounds__(NT, 2) conv_tf32x3_kernel(const float* __restrict__ P, const float* __restrict__ Q, float* __restrict__ O,
                                                            ConvDims d, int kb_per_split, int accumulate) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    DBOA_TL(0);
    const int Ktaps = d.kh * d.kw, Kfull = Ktaps * d.Cin;
    // GEMM extents of this mode
    // Stride-2 data gradient: an input pixel (hi, wi) only receives the filter taps r = hi + pad (mod 2), s = wi + pad (mod 2).
    // The rows are therefore enumerated per PARITY CLASS (hi & 1, wi & 1): blockIdx.x = class * tiles_per_class + tile, every
    // tile holds rows of one class and reduces over that class's taps only (3x3: 1, 2, 2, 4 of 9 taps; 1x1: one class, the
    // other three receive nothing).  Without this 3/4 of the rows of every 128-row tile were zero-filled.
    const bool par = MODE == DGRAD && d.stride == 2;
    int py = 0, px = 0, Hh = d.Hi, Wh = d.Wi, r0 = 0, s0 = 0, nr = d.kh, nsx = d.kw, tstep = 1, mtile = blockIdx.x;
    if (par) {
        // a 1-tap filter axis has ONE non-empty parity (the epilogue zero-fills the sibling pixels); classes = npy * npx
        const int npy = d.kh >= 2 ? 2 : 1, npx = d.kw >= 2 ? 2 : 1;
        const int tpc = gridDim.x / (npy * npx), cls = blockIdx.x / tpc;
        mtile = blockIdx.x - cls * tpc;
        py = npy == 2 ? cls / npx : (d.pad & 1); px = npx == 2 ? cls % npx : (d.pad & 1); Hh = d.Hi >> 1; Wh = d.Wi >> 1;
        r0 = (py + d.pad) & 1; s0 = (px + d.pad) & 1;
        nr = (d.kh - r0 + 1) >> 1; nsx = (d.kw - s0 + 1) >> 1; tstep = 2;
    }
    const int Mrows = MODE == FWD ? d.B * d.Ho * d.Wo : (MODE == DGRAD ? d.B * Hh * Wh : d.Cout);
    const int ldo = MODE == FWD ? d.Cout : (MODE == DGRAD ? d.Cin : Kfull);
    const int Kred = MODE == FWD ? Kfull : (MODE == DGRAD ? nr * nsx * d.Cout : d.B * d.Ho * d.Wo);
    const int m0 = mtile * BM, n0 = blockIdx.y * BN;
    const int nkb_total = (Kred + BK - 1) / BK;
    if (par) kb_per_split = (nkb_total + (int)gridDim.z - 1) / (int)gridDim.z;      // the classes have different reduction lengths
    const int kb_begin = blockIdx.z * kb_per_split;
    const int nkb = max(0, min(kb_begin + kb_per_split, nkb_total) - kb_begin);
    // output row of tile row `row` (identity except for the parity classes)
    auto out_row = [&](int row) {
        if (!par) return row;
        const int b = row / (Hh * Wh), rem = row - b * (Hh * Wh);
        const int hh = rem / Wh, wh = rem - hh * Wh;
        return (b * d.Hi + 2 * hh + py) * d.Wi + 2 * wh + px;
    };
    // input pixels of the parities no tap reaches get zeros (non-accumulating calls): written by the thread that stores
    // the sibling pixel of the same 2x2 cell
    const bool fill_y = par && d.kh < 2 && !accumulate, fill_x = par && d.kw < 2 && !accumulate;
    auto zero_siblings = [&](int orow, int col) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const int dyr = (py ? -1 : 1) * d.Wi, dxr = px ? -1 : 1;
        if (fill_x) *reinterpret_cast<float4*>(O + (size_t)(orow + dxr) * ldo + col) = z;
        if (fill_y) *reinterpret_cast<float4*>(O + (size_t)(orow + dyr) * ldo + col) = z;
        if (fill_x && fill_y) *reinterpret_cast<float4*>(O + (size_t)(orow + dyr + dxr) * ldo + col) = z;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], NPW); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = sm.tmem_base;
    DBOA_TL(1);

    if (warp == NPW) {
        // =====================================================================================
        // MMA issuer: one thread.  Per filled stage 12 x tcgen05.mma (4 k-steps of 8 x {Ah*Bh, Ah*Bl, Al*Bh}); the commit
        // releases the stage to the producers.  Descriptors are precomputed per stage -- a k-step only advances the
        // 14-bit start-address field -- because this thread's issue rate bounds the k-loop (scripts/kernel_timeline.py).
        // =====================================================================================
        if (lane == 0 && nkb > 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint64_t dah[STAGES], dal[STAGES], dbh[STAGES], dbl[STAGES];
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                dah[s] = make_desc(smem_u32(sm.a_hi[s])); dal[s] = make_desc(smem_u32(sm.a_lo[s]));
                dbh[s] = make_desc(smem_u32(sm.b_hi[s])); dbl[s] = make_desc(smem_u32(sm.b_lo[s]));
            }
            constexpr uint64_t KSTEP = (2 * CORE_BYTES) >> 4;      // 8 tf32 = two 16-byte chunks along K, in descriptor units
            for (int it0 = 0, round = 0; it0 < nkb; it0 += STAGES, ++round) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    const int it = it0 + s;
                    if (it < nkb) {
                        DBOA_TLM(it, 5);
                        mbar_wait(&sm.full[s], (uint32_t)(round & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        DBOA_TLM(it, 6);
#pragma unroll
                        for (int kk = 0; kk < BK / 8; ++kk) {
                            mma_tf32(tmem_d, dah[s] + kk * KSTEP, dbh[s] + kk * KSTEP, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                            mma_tf32(tmem_d, dah[s] + kk * KSTEP, dbl[s] + kk * KSTEP, idesc, 1u);
                            mma_tf32(tmem_d, dal[s] + kk * KSTEP, dbh[s] + kk * KSTEP, idesc, 1u);
                        }
                        umma_commit(&sm.empty[s]);       // arrives when every MMA issued so far has completed
                        DBOA_TLM(it, 7);
                    }
                }
            }
            umma_commit(&sm.done);
        }
        pdl_wait();          // every thread of a programmatically launched grid passes the dependency wait before it exits
        pdl_trigger();
    } else {
        // =====================================================================================
        // producers (8 warps): global -> registers (two k-blocks ahead) -> hi/lo split -> shared memory stage
        // FWD / DGRAD: the 128-row operand is pixel-major.  Warp w owns the 8-row groups 2w, 2w+1; lane = (row lr8, chunk
        // pair cpair): one LDG.128 of a warp covers 8 rows x 64 contiguous bytes (whole sectors), and the 8 lanes of one
        // shared-memory store phase fill the 8 rows of one core matrix (conflict free).  Register slot q*2 + h holds
        // row group 2w + q, 16-byte k-chunk h*4 + cpair.
        // =====================================================================================
        const int lr8 = lane & 7, cpair = lane >> 3;
        bool avalid[2] = {false, false};
        int ph[2] = {0, 0}, pw[2] = {0, 0};                // FWD: hi0, wi0 (top-left of the window); DGRAD: hi, wi
        const float* pb[2] = {P, P};
        if (MODE != WGRAD) {
            const int HW = MODE == FWD ? d.Ho * d.Wo : Hh * Wh, Wd = MODE == FWD ? d.Wo : Wh;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int arow = m0 + (warp * 2 + q) * 8 + lr8;
                avalid[q] = arow < Mrows;
                if (avalid[q]) {
                    const int b = arow / HW, rem = arow - b * HW;
                    const int h = rem / Wd, w_ = rem - h * Wd;
                    if (MODE == FWD) { ph[q] = h * d.stride - d.pad; pw[q] = w_ * d.stride - d.pad; pb[q] = P + (size_t)b * d.Hi * d.Wi * d.Cin; }
                    else { ph[q] = tstep * h + py; pw[q] = tstep * w_ + px; pb[q] = P + (size_t)b * d.Ho * d.Wo * d.Cout; }
                }
            }
        }
        // byte offset of (row group 2w + q, k-chunk h*4 + cpair, row lr8) in a stage tile: + q * GROUP_BYTES + h * 4 * CORE_BYTES
        const uint32_t a_off = (uint32_t)(warp * 2) * GROUP_BYTES + (uint32_t)cpair * CORE_BYTES + (uint32_t)lr8 * 16;
        // FWD: the weight tile is K-major too: row group w, chunks h*4 + cpair  ->  register slot h
        const uint32_t b_off = (uint32_t)warp * GROUP_BYTES + (uint32_t)cpair * CORE_BYTES + (uint32_t)lr8 * 16;
        // WGRAD: column tile -> (tap, ci0)
        const int wtap = MODE == WGRAD ? n0 / d.Cin : 0, wci0 = MODE == WGRAD ? n0 - wtap * d.Cin : 0;
        const int wr = wtap / d.kw, wsx = wtap - wr * d.kw;
        // reduction cursor of the two operand streams (tap row, tap column, channel offset): k-blocks are fetched in order,
        // so the im2col decode is an increment, not a division, per k-block
        const int Cred = MODE == DGRAD ? d.Cout : d.Cin;
        struct Cursor { int r, s, c; };
        // (r, s) count the taps of this CTA's tap set: filter tap = (r0 + tstep * r, s0 + tstep * s), nsx taps per filter row
        auto make_cursor = [&](int kb) { Cursor c; const int k0 = kb * BK, tap = k0 / Cred; c.c = k0 - tap * Cred; c.r = tap / max(nsx, 1); c.s = tap - c.r * max(nsx, 1); return c; };
        auto advance = [&](Cursor& c) { c.c += BK; if (c.c >= Cred) { c.c = 0; if (++c.s == nsx) { c.s = 0; ++c.r; } } };
        Cursor ca = make_cursor(kb_begin), cb = ca;
        int kb_a = kb_begin, kb_b = kb_begin;              // next k-block of each stream

        auto fetch_a = [&](float4 (&ra)[4]) {
            if (MODE == FWD) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int hi = ph[q] + ca.r, wi = pw[q] + ca.s;
                    const bool inb = avalid[q] && (unsigned)hi < (unsigned)d.Hi && (unsigned)wi < (unsigned)d.Wi;
                    const float* src = pb[q] + ((size_t)hi * d.Wi + wi) * d.Cin + ca.c + cpair * 4;
                    ra[q * 2] = inb ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ra[q * 2 + 1] = inb ? ldg4(src + 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else if (MODE == DGRAD) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int th = ph[q] + d.pad - (r0 + tstep * ca.r), tw = pw[q] + d.pad - (s0 + tstep * ca.s);
                    bool inb = avalid[q] && th >= 0 && tw >= 0;
                    int ho = th, wo = tw;
                    if (d.stride != 1) {
                        ho = th / d.stride; wo = tw / d.stride;
                        inb = inb && (ho * d.stride == th) && (wo * d.stride == tw);
                    }
                    inb = inb && ho < d.Ho && wo < d.Wo;
                    const float* src = pb[q] + ((size_t)ho * d.Wo + wo) * d.Cout + ca.c + cpair * 4;
                    ra[q * 2] = inb ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ra[q * 2 + 1] = inb ? ldg4(src + 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                // A'[m=co][k=pix] = dY[pix][co]: rows of 128 consecutive co, one row per pixel (transposed by the store)
                const int k0 = kb_a * BK;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = tid + NPROD * j, k = idx >> 5, cv = idx & 31;
                    const int pix = k0 + k, co = m0 + cv * 4;
                    ra[j] = (pix < Kred && co < d.Cout) ? ldg4(P + (size_t)pix * d.Cout + co) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            ++kb_a;
            if (MODE != WGRAD) advance(ca);
        };
        auto fetch_b = [&](float4 (&rb)[2]) {
            const int k0 = kb_b * BK;
            if (MODE == FWD) {
                const float* wrow = Q + (size_t)(n0 + warp * 8 + lr8) * Kfull + k0 + cpair * 4;
                rb[0] = ldg4(wrow);
                rb[1] = ldg4(wrow + 16);
            } else if (MODE == DGRAD) {
                // B[n=ci][k=co] = W[co][tap][ci]: rows of 64 consecutive ci, one row per co (transposed by the store)
                // thread = (k = warp*4 + (lane & 3), nv = j*8 + (lane >> 3)*2 + ((lane >> 2) & 1)): bank-conflict-free transposing store
                const int tap = (r0 + tstep * cb.r) * d.kw + (s0 + tstep * cb.s);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = warp * 4 + (lane & 3), nv = j * 8 + (lane >> 3) * 2 + ((lane >> 2) & 1);
                    rb[j] = ldg4(Q + (size_t)(cb.c + k) * Kfull + (size_t)tap * d.Cin + n0 + nv * 4);
                }
            } else {
                // B'[n=ci][k=pix] = X[pixel shifted by the tap][ci0 + n]
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int idx = tid + NPROD * j, k = idx >> 4, nv = idx & 15;
                    const int pix = k0 + k;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pix < Kred) {
                        const int b = pix / (d.Ho * d.Wo), rem = pix - b * d.Ho * d.Wo;
                        const int ho = rem / d.Wo, wo = rem - ho * d.Wo;
                        const int hi = ho * d.stride - d.pad + wr, wi = wo * d.stride - d.pad + wsx;
                        if ((unsigned)hi < (unsigned)d.Hi && (unsigned)wi < (unsigned)d.Wi)
                            v = ldg4(Q + (((size_t)b * d.Hi + hi) * d.Wi + wi) * d.Cin + wci0 + nv * 4);
                    }
                    rb[j] = v;
                }
            }
            ++kb_b;
            if (MODE == DGRAD) advance(cb);
        };
        auto stash = [&](int s, const float4 (&ra)[4], const float4 (&rb)[2]) {
            if (MODE != WGRAD) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    split_store4(sm.a_hi[s], sm.a_lo[s], a_off + (j >> 1) * GROUP_BYTES + (j & 1) * 4 * CORE_BYTES, ra[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = tid + NPROD * j;
                    split_store_t(sm.a_hi[s], sm.a_lo[s], (idx & 31) * 4, idx >> 5, ra[j]);
                }
            }
            if (MODE == FWD) {
                split_store4(sm.b_hi[s], sm.b_lo[s], b_off, rb[0]);
                split_store4(sm.b_hi[s], sm.b_lo[s], b_off + 4 * CORE_BYTES, rb[1]);
            } else if (MODE == DGRAD) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = warp * 4 + (lane & 3), nv = j * 8 + (lane >> 3) * 2 + ((lane >> 2) & 1);
                    split_store_t_rot(sm.b_hi[s], sm.b_lo[s], nv * 4, k, rb[j], lane >> 3);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int idx = tid + NPROD * j;
                    split_store_t(sm.b_hi[s], sm.b_lo[s], (idx & 15) * 4, idx >> 4, rb[j]);
                }
            }
        };

        // Programmatic dependent launch: everything above (barrier init, TMEM allocation, index set-up) and -- for the
        // forward and data-gradient products -- the first WEIGHT tiles do not depend on the previous kernel in the stream
        // and overlap its tail; activations are touched only after the wait.  (Weights are never written by the kernel
        // that immediately precedes a convolution: optimizer updates are followed by a normally serialized launch.)
        float4 ra[STAGES][4], rb[STAGES][2];
        if (MODE != WGRAD) {
#pragma unroll
            for (int f = 0; f < STAGES; ++f)
                if (f < nkb) fetch_b(rb[f]);
        }
        DBOA_TL(2);
        pdl_wait();
        pdl_trigger();
        DBOA_TL(3);
#pragma unroll
        for (int f = 0; f < STAGES; ++f)
            if (f < nkb) { fetch_a(ra[f]); if (MODE == WGRAD) fetch_b(rb[f]); }
        for (int it0 = 0, round = 0; it0 < nkb; it0 += STAGES, ++round) {
#pragma unroll
            for (int f = 0; f < STAGES; ++f) {
                const int it = it0 + f;
                if (it < nkb) {
                    const int s = f;                                                              // stage == register set
                    DBOA_TLC(it, 0);
                    if (round > 0) mbar_wait(&sm.empty[s], (uint32_t)((round - 1) & 1));         // the MMAs that read this stage are done
                    DBOA_TLC(it, 1);
                    stash(s, ra[f], rb[f]);
                    DBOA_TLC(it, 2);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy (UMMA) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.full[s]);
                    DBOA_TLC(it, 3);
                    if (it == 0) DBOA_TL(4);
                    if (it + STAGES < nkb) { fetch_a(ra[f]); fetch_b(rb[f]); }                    // in flight behind the next stage
                    DBOA_TLC(it, 4);
                }
            }
        }
    }
    DBOA_TL(5);
    if (nkb > 0) mbar_wait(&sm.done, 0u);                 // all MMAs of this CTA have completed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DBOA_TL(6);

    // ---- epilogue: producer warp w reads TMEM lanes (w & 3) * 32 .. + 31 (= output rows), columns (w >> 2) * 32 .. + 31
    const int nz = gridDim.z;
    float* red = reinterpret_cast<float*>(sm.a_hi[0]);    // 128 rows x RED_LD floats (34 KB) over the A tiles, free after the last MMA
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
        // the fp32 tile goes through shared memory (thread = row, conflict free) so that global stores are row-contiguous
        float* dstrow = red + (q4 * 32 + lane) * RED_LD + cgp * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dstrow + q * 4) = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]),
                                                                     __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
    }
    DBOA_TL(7);
    if (nz == 1) {
        __syncthreads();
        for (int v = tid; v < BM * (BN / 4); v += NT) {               // 16 lanes write one 256-byte output row
            const int lr = v >> 4, c4 = (v & 15) * 4;
            const int row = m0 + lr;
            if (row < Mrows) {
                float4 q = *reinterpret_cast<const float4*>(red + lr * RED_LD + c4);
                const int orow = out_row(row);
                float4* dst = reinterpret_cast<float4*>(O + (size_t)orow * ldo + n0 + c4);
                if (accumulate) { const float4 c = *dst; q.x += c.x; q.y += c.y; q.z += c.z; q.w += c.w; }
                *dst = q;
                if (fill_x || fill_y) zero_siblings(orow, n0 + c4);
            }
        }
    } else {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        DBOA_TL(8);
        const int rank = (int)cluster.block_rank();
        const int rows_per = BM / nz;                                 // nz is a power of two <= 16
        for (int v = tid; v < rows_per * (BN / 4); v += NT) {
            const int lr = rank * rows_per + (v >> 4), c4 = (v & 15) * 4;
            const int row = m0 + lr;
            if (row >= Mrows) continue;
            const int orow = out_row(row);
            float4* dst = reinterpret_cast<float4*>(O + (size_t)orow * ldo + n0 + c4);
            if (fill_x || fill_y) zero_siblings(orow, n0 + c4);
            float4 sacc = accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int zb = 0; zb < 16; zb += 8) {                      // 8 remote loads in flight, then added in the order z = 0..nz-1
                if (zb < nz) {
                    float4 q[8];
#pragma unroll
                    for (int z = 0; z < 8; ++z)
                        if (zb + z < nz) q[z] = *reinterpret_cast<const float4*>(cluster.map_shared_rank(red, zb + z) + lr * RED_LD + c4);
#pragma unroll
                    for (int z = 0; z < 8; ++z)
                        if (zb + z < nz) { sacc.x += q[z].x; sacc.y += q[z].y; sacc.z += q[z].z; sacc.w += q[z].w; }
                }
            }
            *dst = sacc;
        }
        DBOA_TL(9);
        cluster.sync();                                               // peers may still be reading this CTA's tile
    }
    DBOA_TL(10);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
}

template <int MODE>
static int launch(const float* p, const float* q, float* o, const ConvDims& d, int rows, int cols, int kred, int accumulate, cudaStream_t st,
                  bool pdl) {
    int mtiles = ceil_div(rows, BM);
    if (MODE == DGRAD && d.stride == 2) {                // parity classes of B*Hi/2*Wi/2 rows, longest tap set decides the split
        const int ncls = (d.kh >= 2 ? 2 : 1) * (d.kw >= 2 ? 2 : 1);
        mtiles = ncls * ceil_div(d.B * (d.Hi / 2) * (d.Wi / 2), BM);
        kred = ((d.kh + 1) / 2) * ((d.kw + 1) / 2) * d.Cout;
    }
    const int nkb = (kred + BK - 1) / BK;
    const int tiles = mtiles * (cols / BN);
    // K-slices per tile (cluster size): a power of two <= 16.  A k-block costs ~0.4 us, the cluster barrier + DSMEM reduction
    // of a split tile ~3 us (profiles/r01b_summary.md section 6), so a split must leave at least `min_kb` k-blocks per CTA.
    static const int min_kb = [] { const char* e = getenv("DBOA_TC_MINKB"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    int ns = 1;
    while (ns < 16 && tiles * ns * 2 <= 296 + tiles && nkb / (ns * 2) >= min_kb) ns *= 2;
    const int per = (nkb + ns - 1) / ns;
    const size_t smem = sizeof(Smem) + 128;
    return launch_ex(conv_tf32x3_kernel<MODE>, dim3(mtiles, cols / BN, ns), dim3(NT), smem, st, dim3(1, 1, ns), pdl, p, q, o, d, per,
                     accumulate);
}

static bool shape_ok(const ConvDims& d) {
    return d.Cin % 64 == 0 && d.Cout % 64 == 0 && d.Kpitch == d.kh * d.kw * d.Cin;
}

}  // namespace tc

#ifdef DBOA_TIMELINE
extern "C" int dboa_debug_set_timeline(unsigned long long* buf) {
    return cudaMemcpyToSymbol(tc::g_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : -3;
}
#endif

int conv_tc_fwd(const float* x, const float* w, float* y, const ConvDims& d, cudaStream_t st, bool pdl) {
    if (g_tc_mode == 0 || !tc::shape_ok(d)) return DBOA_ERR_UNSUPPORTED;
    return tc::launch<tc::FWD>(x, w, y, d, d.B * d.Ho * d.Wo, d.Cout, d.kh * d.kw * d.Cin, 0, st, pdl);
}
int conv_tc_dgrad(const float* dy, const float* w, float* dx, const ConvDims& d, int accumulate, cudaStream_t st, bool pdl) {
    if (g_tc_mode < 2 || !tc::shape_ok(d)) return DBOA_ERR_UNSUPPORTED;
    if (d.stride == 2 && ((d.Hi | d.Wi) & 1)) return DBOA_ERR_UNSUPPORTED;       // the parity-class enumeration needs even extents
    return tc::launch<tc::DGRAD>(dy, w, dx, d, d.B * d.Hi * d.Wi, d.Cin, d.kh * d.kw * d.Cout, accumulate, st, pdl);
}
int conv_tc_wgrad(const float* dy, const float* x, float* dw, const ConvDims& d, cudaStream_t st, bool pdl) {
    if (g_tc_mode < 2 || !tc::shape_ok(d)) return DBOA_ERR_UNSUPPORTED;
    return tc::launch<tc::WGRAD>(dy, x, dw, d, d.Cout, d.kh * d.kw * d.Cin, d.B * d.Ho * d.Wo, 1, st, pdl);
}

int conv1x1_tc_fwd(const float* x, const float* w, float* y, int M, int Cin, int Cout, cudaStream_t st, bool pdl) {
    ConvDims d;
    d.B = 1; d.Hi = M; d.Wi = 1; d.Cin = Cin; d.Ho = M; d.Wo = 1; d.Cout = Cout; d.kh = 1; d.kw = 1; d.stride = 1; d.pad = 0; d.Kpitch = Cin;
    return conv_tc_fwd(x, w, y, d, st, pdl);
}

}  // namespace dboa
