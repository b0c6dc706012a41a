#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace dboa {

struct CosinePairs {
    const float* a[16];
    const float* b[16];
    long long n[16];
    int blk_off[17];
    int npairs;
};

int sgd_update(const float* p, const float* g, float* out, float lr, size_t n, cudaStream_t st);
int adam_ema(float* p, const float* g, float* m, float* v, float* teacher, size_t n, float lr, float beta1, float beta2, float eps, int step,
             float alpha, float gscale, cudaStream_t st);
int ema_update(float* teacher, const float* p, size_t n, float alpha, cudaStream_t st);
// out[i] = cos(a_i, b_i) (NULL to skip); terms[i] = (a.b, |a|^2, |b|^2) in double (NULL to skip): the data-parallel feature test all-reduces them
int cosine_pairs(const CosinePairs& cp, float* partial, size_t partial_floats, float* out, double* terms, float eps, cudaStream_t st);
long long cosine_partial_floats(const long long* n, int npairs);
int retrieval_nearest(const float* feat, const float* centers, int K, int D, int* best, float* dists, cudaStream_t st);

}  // namespace dboa
