// Micro-benchmark: issue / pipe throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and
// mixed with integer-ALU, MUFU and shared-memory instructions (the instruction classes of the compositing kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_f32x2 tools/ubench_f32x2.cu && tools/ubench_f32x2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float lo(unsigned long long v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}

constexpr int ITERS = 4096;

// MODE 0: 16 scalar FFMA chains; 1: 8 FFMA2 chains; 2: 0 + 8 integer ops/iter; 3: 1 + 8 integer ops/iter;
// 4: 0 + 4 MUFU.EX2/iter; 5: 1 + 4 MUFU.EX2/iter; 6: 0 + 4 LDS.128/iter; 7: 1 + 4 LDS.128/iter
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float x, float y, int seed)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(x, y, x, y);
    __syncthreads();
    float a[16];
    unsigned long long p[8];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3f + i;
    for (int i = 0; i < 8; i++) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const unsigned long long xx = pk(x, x), yy = pk(y, y);
    unsigned v0 = seed + threadIdx.x, v1 = v0 * 3, v2 = v0 * 5, v3 = v0 * 7;
    float e0 = x, e1 = y, e2 = x + y, e3 = x - y;
    float4 l0 = make_float4(0, 0, 0, 0), l1 = l0, l2 = l0, l3 = l0;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        if (MODE & 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], xx, yy);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmaf_rn(a[i], x, y);
        }
        if ((MODE >> 1) == 1) {
            v0 = (v0 ^ v1) + it; v1 = (v1 & v2) + v0; v2 = (v2 | v3) + v1; v3 = (v3 ^ v0) + v2;
        }
        if ((MODE >> 1) == 2) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e3));
        }
        if ((MODE >> 1) == 3) {
            const unsigned ad = sb + ((it & 15) << 6);
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l0.x), "=f"(l0.y), "=f"(l0.z), "=f"(l0.w) : "r"(ad));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l1.x), "=f"(l1.y), "=f"(l1.z), "=f"(l1.w) : "r"(ad + 16));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l2.x), "=f"(l2.y), "=f"(l2.z), "=f"(l2.w) : "r"(ad + 32));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l3.x), "=f"(l3.y), "=f"(l3.z), "=f"(l3.w) : "r"(ad + 48));
            e0 += l0.x + l1.y + l2.z + l3.w;
        }
    }
    float s = 0.f;
    if (MODE & 1) for (int i = 0; i < 8; i++) s += lo(p[i]);
    else for (int i = 0; i < 16; i++) s += a[i];
    s += __uint_as_float((v0 ^ v1 ^ v2 ^ v3) & 0x3fffffu) + e0 + e1 + e2 + e3;
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, float* d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    k<MODE><<<grid, 256>>>(d, 1.0001f, 1e-7f, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<grid, 256>>>(d, 1.0001f, 1e-7f, r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double fma = (double)grid * 256 * ITERS * 16;
    printf("%-34s %8.3f ms   %7.2f TFMA/s   %6.1f FMA/clk/SM @1.965GHz\n", name, ms, fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    float* d;
    cudaMalloc(&d, 4);
    run<0>("16 FFMA", d);
    run<1>("8 FFMA2", d);
    run<2>("16 FFMA + 8 int", d);
    run<3>("8 FFMA2 + 8 int", d);
    run<4>("16 FFMA + 4 MUFU", d);
    run<5>("8 FFMA2 + 4 MUFU", d);
    run<6>("16 FFMA + 4 LDS.128", d);
    run<7>("8 FFMA2 + 4 LDS.128", d);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
