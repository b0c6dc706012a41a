// Experiment: what does the FIRST global load of a kernel cost, compared with later loads?  (nvcc -arch=sm_100a)
// Each CTA's thread 0..255 loads one float4 from an L2-resident buffer at kernel start (phase A), then a second,
// different line (phase B), then a third (phase C); thread 0 of each CTA records clock64 deltas.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void writer(float4* buf, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void probe(const float4* __restrict__ buf, size_t stride, long long* out, float* sink, int use_pdl) {
    if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    const size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long t0 = clock64();
    float4 a = __ldg(buf + base);
    float s = a.x + a.y;
    asm volatile("" ::"f"(s));
    long long t1 = clock64();
    float4 b = __ldg(buf + base + stride);
    s += b.x + b.y;
    asm volatile("" ::"f"(s));
    long long t2 = clock64();
    float4 c = __ldg(buf + base + 2 * stride);
    s += c.x + c.y;
    asm volatile("" ::"f"(s));
    long long t3 = clock64();
    __syncthreads();
    long long t4 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 4 + 0] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t1; out[blockIdx.x * 4 + 2] = t3 - t2; out[blockIdx.x * 4 + 3] = t4 - t3; }
    if (s == 12345.f) sink[0] = s;
}
int main() {
    const int ctas = 128, nt = 256;
    const size_t stride = (size_t)ctas * nt, n = stride * 3;
    float4* buf; long long* out; float* sink;
    cudaMalloc(&buf, n * sizeof(float4)); cudaMalloc(&out, ctas * 4 * sizeof(long long)); cudaMalloc(&sink, 4);
    long long h[ctas * 4];
    for (int mode = 0; mode < 3; ++mode) {            // 0: cold-ish after writer kernel, 1: back-to-back probes, 2: probes with PDL attribute
        for (int rep = 0; rep < 3; ++rep) {
            if (mode == 0) writer<<<(n + 255) / 256, 256>>>(buf, n);
            if (mode == 2) {
                cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(nt);
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                for (int k = 0; k < 4; ++k) cudaLaunchKernelEx(&cfg, probe, (const float4*)buf, stride, out, sink, 1);
            } else {
                for (int k = 0; k < 4; ++k) probe<<<ctas, nt>>>(buf, stride, out, sink, 0);
            }
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double m[4] = {0, 0, 0, 0}; long long mx[4] = {0, 0, 0, 0};
            for (int c = 0; c < ctas; ++c) for (int j = 0; j < 4; ++j) { m[j] += h[c * 4 + j]; if (h[c * 4 + j] > mx[j]) mx[j] = h[c * 4 + j]; }
            printf("mode %d rep %d: first load %.0f (max %lld), second %.0f (max %lld), third %.0f (max %lld), syncthreads %.0f cycles\n", mode, rep,
                   m[0] / ctas, mx[0], m[1] / ctas, mx[1], m[2] / ctas, mx[2], m[3] / ctas);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
