// TMA feed probe (sm_100a): what does one SM ingest from L2 / HBM through cp.async.bulk.tensor when nothing else runs?
// Every CTA (one per SM) keeps `slots` box loads in flight (one mbarrier each) and re-issues a slot as soon as it has landed:
// bytes per clock and SM as a function of the box height (rows of 128 bytes), the ring depth and the source (a 2.5 MB matrix
// that lives in L2 = the weights; a 0.7 GB matrix streamed from HBM = the residual stream).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_feed tma_feed.cu -lcuda && ./tma_feed
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) feed_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int slots, int iters,
                                                      int tensor_rows, int n_colblk, int same_lines, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* data = smem + 1024;
    const int slot_bytes = ((box_rows * 128 + 1023) / 1024) * 1024;
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots * 4; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // `issuers` warps (lane 0 of each) run the loop on their own slots: is the cost per box a property of the issuing thread
    // or of the SM's TMA unit?
    // gridDim.y = issuers; gridDim.z = 1: the issuers are lane 0 of `issuers` warps, 2: lanes 0 .. issuers - 1 of warp 0
    const int issuers = gridDim.y;
    const bool lanes_mode = gridDim.z == 2;
    const int w = lanes_mode ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
    if (blockIdx.y != 0 || blockIdx.z != 0) return;
    if (lanes_mode ? threadIdx.x >= issuers : ((threadIdx.x & 31) != 0 || w >= issuers)) return;
    bars += w * slots;
    data += (size_t)w * slots * slot_bytes;
    const unsigned long long t0 = clock64();
    // every SM walks its own part of the tensor (or, same_lines = 1, all SMs the same lines at the same time)
    int row = same_lines ? 0 : (int)((long long)blockIdx.x * 4096 % tensor_rows), col = 0;
    auto issue = [&](int s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(box_rows * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(data + s * slot_bytes)), "l"((uint64_t)&tm), "r"(smem_u32(&bars[s])), "r"(col * 64), "r"(row) : "memory");
        col += 1;
        if (col == n_colblk) { col = 0; row += box_rows; if (row + box_rows > tensor_rows) row = 0; }
    };
    for (int s = 0; s < slots; ++s) issue(s);
    for (int it = 0; it < iters; ++it) {
        const int s = it % slots;
        const uint32_t parity = (it / slots) & 1;
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(parity) : "memory");
        if (it + slots < iters) issue(s);
    }
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    cudaSetDevice(0);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    unsigned long long* cyc;
    cudaMalloc(&cyc, sms * 8);
    struct Src { const char* name; long long rows; int cols; } srcs[2] = {{"weights (L2, 640 x 1920 x 2 planes)", 640, 3840},
                                                                           {"residual stream (HBM, 513280 x 640)", 513280, 640}};
    for (const Src& src : srcs) {
        void* buf;
        cudaMalloc(&buf, (size_t)src.rows * src.cols * 2);
        cudaMemset(buf, 1, (size_t)src.rows * src.cols * 2);
        for (int box_rows : {112}) {
            if (box_rows > src.rows) continue;
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)src.cols, (cuuint64_t)src.rows}, strides[1] = {(cuuint64_t)src.cols * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int slots : {2, 4}) {
                for (int issuers : {1, 2, 4, -2, -4}) {
                    const int same = 0;
                    const int zmode = issuers < 0 ? 2 : 1;
                    if (issuers < 0) issuers = -issuers;
                    if (src.rows > 100000 && issuers > 1) continue;
                    const int iters = 4000;
                    const int slot_bytes = ((box_rows * 128 + 1023) / 1024) * 1024;
                    const size_t smem = 1024 + (size_t)issuers * slots * slot_bytes;
                    if (smem > 220 * 1024) continue;
                    std::vector<unsigned long long> h(sms);
                    for (int rep = 0; rep < 2; ++rep) {
                        feed_kernel<<<dim3(sms, issuers, zmode), 128, smem>>>(tm, box_rows, slots, iters, (int)src.rows, src.cols / 64, same, cyc);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
                    double mx = 0, sum = 0;
                    for (auto c : h) { sum += c; if (c > mx) mx = c; }
                    const double bytes = (double)iters * box_rows * 128 * issuers;
                    printf("%-40s box %3d rows  slots %2d  issuers %d %s  %7.1f B/clk/SM, %6.0f cyc per box and warp (%5.0f per box overall)\n",
                           src.name, box_rows, slots, issuers, zmode == 2 ? "lanes of one warp" : "warps            ", bytes / (sum / sms), (sum / sms) / iters, (sum / sms) / iters / issuers);
                }
            }
        }
        cudaFree(buf);
    }
    return 0;
}
