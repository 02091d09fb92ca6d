// Probe: does `cp.async.bulk.tensor.2d...tile::gather4` (SASS UTMALDG.2D.GATHER4) gather four 64-byte (or 48-byte) per-Gaussian
// records by row index from a [P,16] float array, and how must the tensor map be encoded?  Prints one line per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_gather4_probe tools/tma_gather4_probe.cu && tools/tma_gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, const int* idx, float* out, int* status, int floats_per_row, int bytes)
{
    __shared__ __align__(128) float s[4 * 16];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned sbar = (unsigned)__cvta_generic_to_shared(&bar), ss = (unsigned)__cvta_generic_to_shared(s);
    if (threadIdx.x < 64) s[threadIdx.x] = -1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(bytes));
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(ss), "l"(&tm), "r"(0), "r"(idx[0]), "r"(idx[1]), "r"(idx[2]), "r"(idx[3]), "r"(sbar) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < 2000000 && !ok; spin++) {
        unsigned p;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(p) : "r"(sbar));
        ok = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) *status = ok;
    if (threadIdx.x < 64) out[threadIdx.x] = s[threadIdx.x];
}

int main()
{
    const int P = 1000;
    std::vector<float> h(P * 16);
    for (int i = 0; i < P * 16; i++) h[i] = (float)i;
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int hidx[4] = {7, 901, 33, 512};
    int* didx; cudaMalloc(&didx, 16); cudaMemcpy(didx, hidx, 16, cudaMemcpyHostToDevice);
    float* dout; cudaMalloc(&dout, 64 * 4);
    int* dst; cudaMalloc(&dst, 4);
    PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    for (int width : {16, 12}) {
        for (int boxrows : {1, 4}) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {16, (cuuint64_t)P};
            cuuint64_t strides[1] = {64};
            cuuint32_t box[2] = {(cuuint32_t)width, (cuuint32_t)boxrows};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("width %d boxrows %d: encode failed %d\n", width, boxrows, (int)r); continue; }
            cudaMemset(dst, 0, 4);
            probe<<<1, 64>>>(tm, didx, dout, dst, width, 4 * width * 4);
            cudaError_t e = cudaDeviceSynchronize();
            float o[64]; int st = -1;
            cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
            cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
            int good = 1;
            for (int k = 0; k < 4; k++)
                for (int c = 0; c < width; c++) good &= (o[k * width + c] == (float)(hidx[k] * 16 + c));
            printf("width %2d boxrows %d: %s barrier %s rows %s  first of each row: %.0f %.0f %.0f %.0f\n", width, boxrows,
                   cudaGetErrorString(e), st == 1 ? "completed" : "TIMEOUT", good ? "CORRECT (dense, row k at k*width)" : "mismatch",
                   o[0], o[width], o[2 * width], o[3 * width]);
            if (e != cudaSuccess) return 2;
        }
    }
    return 0;
}
