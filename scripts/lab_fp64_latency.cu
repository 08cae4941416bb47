#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double a, double b, double* out, long long* t)
{
    double x = a, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = x / y + 1e-9;
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = sqrt(x) + 1.5;
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = 1.0 / x + 0.7;
    long long t3 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = fma(x, 0.999, 0.001);
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { double s, c; sincos(x * 1e-3, &s, &c); x = s + c; }
    long long t5 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = sin(x * 1e-3) + 0.5;
    long long t6 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = __shfl_sync(0xffffffffu, x, 3) + 0.5;
    long long t7 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { double q0 = x * y; double r0 = fma(-q0, b, x); double q1 = fma(r0, y, q0); double r1 = fma(-q1, b, x); x = fma(r1, y, q1) + 1e-9; }
    long long t8 = clock64();
    if (threadIdx.x == 0) { out[0] = x; t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; t[7] = t8 - t7; }
}
int main()
{
    double* o; long long* t; cudaMalloc(&o, 8); cudaMalloc(&t, 64);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(3.7, 1.3, o, t);
    long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
    const char* n[8] = {"div+add", "sqrt+add", "rcp+add", "fma", "sincos+add", "sin+add", "shfl64+add", "markstein5+add"};
    for (int i = 0; i < 8; ++i) printf("%s: %.1f cycles\n", n[i], h[i] / 1000.0);
    return 0;
}
