// Row N4 of SURVEY.md §8f, second half: the tensor surgery of densification.  Every 100 iterations the reference
// rebuilds all 15 parameter tensors, both RAdam moments of each and ~18 per-Gaussian statistics tensors
//   * by boolean mask   - scene/c_gaussian_model.py:693-713 `_prune_optimizer` (`x[mask]` per tensor and per moment),
//                         :715-763 `prune_points` (the statistics),
//   * by concatenation  - :765-787 `cat_tensors_to_optimizer` (`torch.cat((x, extension))`, zero moments appended),
//                         :966-1017 `densify_and_clone` (the extension rows are `x[selected]`),
//                         :874-964 `densify_and_split` (`x[selected].repeat(N, 1)`):
// ~45 masked gathers (each a `nonzero` with a host wait) and ~45 concatenations per call.
//
// Here all of it is ONE kernel over a table of row-gather jobs that travels in the kernel parameters:
//     dst[r] = r < n_a ? a[index ? index[r] : r] : (b ? b[r - n_a] : 0)        for r in [0, n_out)
// - pruning is `index = nonzero(keep)`, n_a = n_out; cloning / splitting is `index = [0..n) ++ selected (repeated)`
// with `n_a = n` for the moments (their new rows are zero) and n_a = n_out for the parameters; a plain concatenation
// has no index, `b` = the extension.  Rows are 4-byte words.  Jobs without an index are two contiguous copies (or a copy
// and a fill): they run flat, 128-bit accesses over the whole array whatever the row length.  Indexed jobs: rows of at
// least 16 words are copied by one warp each (128-bit accesses when the row length and all base addresses allow),
// shorter rows one word per thread.  HBM-bound: every array is read once and written once.
#include "common.cuh"

namespace {

constexpr int kGatherThreads = 256;
constexpr unsigned kRowsPerBlock = 64;          // long rows: 8 warps x 8 rows
constexpr unsigned kWordsPerBlock = 2048;       // short rows: 256 threads x 8 words
constexpr unsigned kFlatPerBlock = 4096;        // jobs without an index: 256 threads x 4 x 128 bits

struct GatherKernelParams {
    int n;
    unsigned block_end[EX_GATHER_MAX_JOBS];     // exclusive prefix of blocks per job
    GatherJob j[EX_GATHER_MAX_JOBS];
};

__device__ __forceinline__ const uint32_t* source_row(const GatherJob& d, unsigned r)
{
    if (r < d.n_a) {
        const long long s = d.index ? __ldg(d.index + r) : (long long)r;
        return d.a + (size_t)s * d.words;
    }
    return d.b ? d.b + (size_t)(r - d.n_a) * d.words : nullptr;
}

__global__ void __launch_bounds__(kGatherThreads) gather_rows_kernel(const __grid_constant__ GatherKernelParams k, unsigned total_blocks)
{
    for (unsigned blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
        int ji = 0;
#pragma unroll 1
        while (ji + 1 < k.n && blk >= k.block_end[ji]) ji++;
        const GatherJob& d = k.j[ji];
        const unsigned local = blk - (ji ? k.block_end[ji - 1] : 0u);
        if (d.index == nullptr) {
            // contiguous: dst[0, split) = a[0, split), dst[split, total) = b[0, total - split) or zero
            const size_t total = (size_t)d.n_out * d.words, split = (size_t)d.n_a * d.words;
            if (d.vec4) {
                const size_t t4 = total >> 2, s4 = split >> 2;
#pragma unroll
                for (unsigned i = 0; i < kFlatPerBlock / (4 * kGatherThreads); i++) {
                    const size_t o4 = (size_t)local * (kFlatPerBlock / 4) + i * kGatherThreads + threadIdx.x;
                    if (o4 < t4) {
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (o4 < s4) v = __ldg(reinterpret_cast<const uint4*>(d.a) + o4);
                        else if (d.b) v = __ldg(reinterpret_cast<const uint4*>(d.b) + (o4 - s4));
                        reinterpret_cast<uint4*>(d.dst)[o4] = v;
                    }
                }
                if (local == 0 && threadIdx.x < (total & 3)) {          // the last 1-3 words (they lie behind `split`)
                    const size_t o = (t4 << 2) + threadIdx.x;
                    d.dst[o] = d.b ? __ldg(d.b + (o - split)) : 0u;
                }
            } else {
                for (unsigned i = 0; i < kFlatPerBlock / kGatherThreads; i++) {
                    const size_t o = (size_t)local * kFlatPerBlock + i * kGatherThreads + threadIdx.x;
                    if (o >= total) break;
                    d.dst[o] = o < split ? __ldg(d.a + o) : (d.b ? __ldg(d.b + (o - split)) : 0u);
                }
            }
        } else if (d.words >= 16) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (unsigned i = 0; i < kRowsPerBlock / 8; i++) {
                const unsigned r = local * kRowsPerBlock + i * 8 + warp;
                if (r >= d.n_out) break;
                const uint32_t* src = source_row(d, r);
                uint32_t* dst = d.dst + (size_t)r * d.words;
                if (d.vec4) {
                    const unsigned w4 = d.words >> 2;
                    for (unsigned c = lane; c < w4; c += 32)
                        reinterpret_cast<uint4*>(dst)[c] = src ? __ldg(reinterpret_cast<const uint4*>(src) + c) : make_uint4(0u, 0u, 0u, 0u);
                } else {
                    for (unsigned c = lane; c < d.words; c += 32) dst[c] = src ? __ldg(src + c) : 0u;
                }
            }
        } else {
            const size_t total = (size_t)d.n_out * d.words;
            for (unsigned i = 0; i < kWordsPerBlock / kGatherThreads; i++) {
                const size_t o = (size_t)local * kWordsPerBlock + i * kGatherThreads + threadIdx.x;
                if (o >= total) break;
                const unsigned r = (unsigned)(o / d.words), c = (unsigned)(o - (size_t)r * d.words);
                const uint32_t* src = source_row(d, r);
                d.dst[o] = src ? __ldg(src + c) : 0u;
            }
        }
    }
}

}  // namespace

cudaError_t launch_gather_rows(const GatherJob* jobs, int n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    GatherKernelParams k;
    k.n = n;
    unsigned long long blocks = 0;
    for (int i = 0; i < n; i++) {
        k.j[i] = jobs[i];
        const GatherJob& d = jobs[i];
        const uintptr_t al = (uintptr_t)d.a | (uintptr_t)d.b | (uintptr_t)d.dst;
        if (d.index == nullptr) {
            const unsigned long long total = (unsigned long long)d.n_out * d.words, split = (unsigned long long)d.n_a * d.words;
            k.j[i].vec4 = ((al & 15) == 0 && split % 4 == 0) ? 1 : 0;
            blocks += (total + kFlatPerBlock - 1) / kFlatPerBlock;
        } else {
            k.j[i].vec4 = (d.words % 4 == 0 && (al & 15) == 0) ? 1 : 0;
            if (d.words >= 16) blocks += ((unsigned long long)d.n_out + kRowsPerBlock - 1) / kRowsPerBlock;
            else blocks += ((unsigned long long)d.n_out * d.words + kWordsPerBlock - 1) / kWordsPerBlock;
        }
        if (blocks > 0xFFFFFFFFull) return cudaErrorInvalidValue;
        k.block_end[i] = (unsigned)blocks;
    }
    for (int i = n; i < EX_GATHER_MAX_JOBS; i++) k.block_end[i] = (unsigned)blocks;
    if (blocks == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long want = (unsigned long long)sms * 8;
    gather_rows_kernel<<<(unsigned)(blocks < want ? blocks : want), kGatherThreads, 0, s>>>(k, (unsigned)blocks);
    return cudaGetLastError();
}
