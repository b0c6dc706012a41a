// tcgen05 (5th-generation tensor core) path for GEMM-shaped convolutions: 1x1 / stride-1 convolutions of the
// backbone, i.e.  Y[M][N] = X[M][K] * W[N][K]^T  with M = B*H*W pixels, K = Cin, N = Cout, all fp32 K-major.
//
// Precision: `tcgen05.mma.kind::tf32` with a 3-term split.  Every operand value x is split by the loader
// warps into hi = x with the low 13 mantissa bits cleared (exactly a TF32 number, so the tensor core's own
// fp32->tf32 conversion is the identity) and lo = x - hi (exact in fp32); the accumulator receives
//   Ah*Bh + Ah*Bl + Al*Bh      (the dropped Al*Bl term is ~2^-22 relative)
// in fp32 inside TMEM, which keeps the result within ~1e-6 of the exact-fp32 CUDA-core path (conv.cu), against
// which this kernel is validated on the device (tests/test_gpu_tc.py).  SURVEY.md §7 "hard part 1".
//
// Structure (one CTA = 128 threads, one 128 x BN output tile, BK = 32 per stage, 2 stages):
//   all threads : coalesced float4 loads of the X / W sub-tiles -> hi/lo split in registers -> st.shared into the
//                 UMMA canonical K-major no-swizzle layout (8-row x 16-byte core matrices)
//   thread 0    : 12 x tcgen05.mma (4 k-steps of 8 x 3 products) per stage, tcgen05.commit -> mbarrier of the stage
//   epilogue    : tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> global (or split-K partial) stores
// Accumulators live in TMEM (BN columns x 128 lanes), never in registers.
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace dboa {

static int g_tc_mode = 0;          // 0 = off, 1 = on, 2 = on with LBO/SBO roles swapped (bring-up aid)
bool conv_tc_enabled() { return g_tc_mode != 0; }
void conv_tc_set_enabled(bool on) { g_tc_mode = on ? 1 : 0; }
void conv_tc_set_mode(int mode) { g_tc_mode = mode; }

namespace tc {

constexpr int BM = 128, BK = 32, NT = 128, STAGES = 2;
constexpr uint32_t CORE_BYTES = 128;                 // one 8 x 16B core matrix
constexpr uint32_t KCHUNKS = BK / 4;                 // 16-byte chunks along K per stage (8)
constexpr uint32_t GROUP_BYTES = KCHUNKS * CORE_BYTES;   // one 8-row group of a stage tile (1024 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout_type=0 (no swizzle)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
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

// split a float4 into (hi, lo) and store both at byte offset `off` of the hi / lo stage tiles
__device__ __forceinline__ void split_store(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 v) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
}

template <int BN>
struct Smem {
    // stage tiles in UMMA canonical layout: [row_group][k_chunk][8 rows][16 B]
    alignas(128) uint8_t a_hi[STAGES][BM * BK * 4];
    alignas(128) uint8_t a_lo[STAGES][BM * BK * 4];
    alignas(128) uint8_t b_hi[STAGES][BN * BK * 4];
    alignas(128) uint8_t b_lo[STAGES][BN * BK * 4];
    alignas(8) uint64_t mma_done[STAGES];
    uint32_t tmem_base;
};

template <int BN>
__global__ void __launch_bounds__(NT) gemm_tf32x3_kernel(const float* __restrict__ X, const float* __restrict__ W, float* __restrict__ Y,
                                                         int M, int N, int K, int kb_per_split, int swap_lbo_sbo) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem<BN>& sm = *reinterpret_cast<Smem<BN>*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nkb_total = K / BK;
    const int kb_begin = blockIdx.z * kb_per_split;
    const int kb_end = min(kb_begin + kb_per_split, nkb_total);
    const int nkb = kb_end - kb_begin;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&sm.mma_done[s], 1);
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

    // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t lbo = swap_lbo_sbo ? GROUP_BYTES : CORE_BYTES;
    const uint32_t sbo = swap_lbo_sbo ? CORE_BYTES : GROUP_BYTES;

    for (int it = 0; it < nkb; ++it) {
        const int s = it & 1;
        if (it >= STAGES) mbar_wait(&sm.mma_done[s], (uint32_t)(((it >> 1) - 1) & 1));   // stage buffers free again
        const int k0 = (kb_begin + it) * BK;
        // ---- A (activations): 128 x 32 floats = 1024 float4, 8 per thread
#pragma unroll
        for (int j = 0; j < (BM * BK / 4) / NT; ++j) {
            const int idx = tid + NT * j;
            const int i = idx & 7, c = (idx >> 3) & 7, g = idx >> 6;
            const int row = m0 + g * 8 + i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < M) v = ldg4(X + (size_t)row * K + k0 + c * 4);
            split_store(sm.a_hi[s], sm.a_lo[s], g * GROUP_BYTES + c * CORE_BYTES + i * 16, v);
        }
        // ---- B (weights): BN x 32 floats
#pragma unroll
        for (int j = 0; j < (BN * BK / 4) / NT; ++j) {
            const int idx = tid + NT * j;
            const int i = idx & 7, c = (idx >> 3) & 7, g = idx >> 6;
            float4 v = ldg4(W + (size_t)(n0 + g * 8 + i) * K + k0 + c * 4);
            split_store(sm.b_hi[s], sm.b_lo[s], g * GROUP_BYTES + c * CORE_BYTES + i * 16, v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async-proxy (UMMA) reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(sm.a_hi[s]), al = smem_u32(sm.a_lo[s]);
            const uint32_t bh = smem_u32(sm.b_hi[s]), bl = smem_u32(sm.b_lo[s]);
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {
                const uint32_t koff = kk * 2 * CORE_BYTES;                 // 8 tf32 = two 16-byte chunks along K
                const uint64_t dah = make_desc(ah + koff, lbo, sbo), dal = make_desc(al + koff, lbo, sbo);
                const uint64_t dbh = make_desc(bh + koff, lbo, sbo), dbl = make_desc(bl + koff, lbo, sbo);
                mma_tf32(tmem_d, dah, dbh, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                mma_tf32(tmem_d, dal, dbh, idesc, 1u);
            }
            umma_commit(&sm.mma_done[s]);        // arrives when every MMA issued so far has completed
        }
    }
    // the last commit covers all MMAs of this CTA
    if (nkb > 0) {
        const int last = nkb - 1;
        mbar_wait(&sm.mma_done[last & 1], (uint32_t)((last >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: thread t owns accumulator lane (= output row) t; 32 columns per tcgen05.ld
    const int row = m0 + tid;
    float* out = Y + (size_t)blockIdx.z * M * N;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
        if (row < M && nkb > 0) {
            float4* dst = reinterpret_cast<float4*>(out + (size_t)row * N + n0 + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                                     __uint_as_float(r[q * 4 + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
}

}  // namespace tc

static float* g_tc_ws = nullptr;           // split-K workspace handed over by the plan
static size_t g_tc_ws_floats = 0;
void conv_tc_set_workspace(float* ws, size_t floats) { g_tc_ws = ws; g_tc_ws_floats = floats; }

int conv1x1_tc_fwd(const float* x, const float* w, float* y, int M, int Cin, int Cout, cudaStream_t st) {
    using namespace tc;
    constexpr int BN = 64;
    if (g_tc_mode == 0 || Cin % BK != 0 || Cout % BN != 0 || M < 1) return DBOA_ERR_UNSUPPORTED;
    const int nkb = Cin / BK;
    const int tiles = ceil_div(M, BM) * (Cout / BN);
    int ns = 1;
    if (tiles < 148 && nkb >= 8 && g_tc_ws != nullptr) {
        ns = (296 + tiles - 1) / tiles;
        if (ns > nkb / 4) ns = nkb / 4;
        while (ns > 1 && (size_t)ns * M * Cout > g_tc_ws_floats) --ns;
        if (ns < 1) ns = 1;
    }
    int per = (nkb + ns - 1) / ns;
    ns = (nkb + per - 1) / per;
    static bool attr_set = false;
    const size_t smem = sizeof(Smem<BN>) + 128;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return DBOA_ERR_CUDA; }
        attr_set = true;
    }
    dim3 grid(ceil_div(M, BM), Cout / BN, ns);
    gemm_tf32x3_kernel<BN><<<grid, NT, smem, st>>>(x, w, ns > 1 ? g_tc_ws : y, M, Cout, Cin, per, g_tc_mode == 2 ? 1 : 0);
    DBOA_TRY(check_launch());
    if (ns > 1) {
        DBOA_TRY(splitk_reduce(g_tc_ws, y, (size_t)M * Cout, ns, 0, st));
    }
    return DBOA_OK;
}

}  // namespace dboa
