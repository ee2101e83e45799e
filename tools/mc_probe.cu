// TMA multicast probe (B200, sm_100a): does sharing an operand slab between the CTAs of a cluster by
// cp.async.bulk.tensor ... .multicast::cluster relieve the L2 -> shared-memory path that bounds the CTA-pair GEMM kernels
// of this repo?  Developer tool, not part of the library:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o gpurun_out/mc_probe tools/mc_probe.cu -lcuda
//   gpurun_out/mc_probe            (prints one line per mode; run the same binary under ncu for the L2 sector counts)
//
// Every CTA keeps a ring of 32 KB stages (the B slab of one K block of a 256-wide tile: 256 rows x 64 bf16) filled by TMA
// from an L2-resident matrix and frees a stage as soon as it has landed -- no math, so the kernel measures how fast the memory
// system delivers slabs into shared memory.  Modes, cluster size CS in {1, 2, 4}:
//   distinct   every CTA loads its own slab                                (no sharing: the unicast cap)
//   shared-uc  the CS CTAs of a cluster load the SAME slab, each by itself (what neighbouring tiles of a GEMM do today)
//   shared-mc  each CTA loads 1/CS of the slab and multicasts it to all CS (one L2 read per slab and cluster)
// "delivered" counts the bytes that land in shared memory (CS x 32 KB per slab in the shared modes).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../strotss_tensorflow_b200/csrc/common.cuh"

using namespace sb;

constexpr int kStages = 4;
constexpr int kSlabRows = 256, kSlabCols = 64;
constexpr int kSlabBytes = kSlabRows * kSlabCols * 2;

__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r;
}
__device__ __forceinline__ void arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}

// mode 0 distinct, 1 shared-uc, 2 shared-mc.  tm_full: box 256 x 64; tm_part: box (256 / CS) x 64.
template <int CS>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_part,
                                                       int mode, int iters, int slabs_k, int slabs_r, unsigned long long* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[kStages], empty[kStages];
    const uint32_t rank = CS > 1 ? cta_rank() : 0;
    const int cluster = blockIdx.x / CS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode == 2 ? CS : 1); }
        fence_barrier_init();
    }
    __syncthreads();
    if (CS > 1) cluster_sync();
    const int who = (mode == 0) ? blockIdx.x : cluster;             // which slab sequence this CTA walks
    if (threadIdx.x == 0) {                                          // producer
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            const int idx = (who * 7 + it) % (slabs_k * slabs_r);
            const int c0 = (idx % slabs_k) * kSlabCols, c1 = (idx / slabs_k) * kSlabRows;
            mbar_arrive_expect_tx(&full[stage], kSlabBytes);
            uint8_t* dst = smem + stage * kSlabBytes;
            if (mode == 2) {
                constexpr int part = kSlabRows / CS;
                tma_load_2d_mc(dst + rank * part * kSlabCols * 2, &tm_part, &full[stage], c0, c1 + static_cast<int>(rank) * part,
                               static_cast<uint16_t>((1u << CS) - 1));
            } else {
                tma_load_2d(dst, &tm_full, &full[stage], c0, c1);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {                                  // consumer: frees a stage as soon as it has landed
        int stage = 0; uint32_t phase = 0;
        unsigned long long acc = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&full[stage], phase);
            acc += *reinterpret_cast<volatile unsigned int*>(smem + stage * kSlabBytes + (it & 255) * 4);
            if (mode == 2) {
                for (int r = 0; r < CS; ++r) arrive_remote(mapa(smem_u32(&empty[stage]), r));
            } else {
                mbar_arrive(&empty[stage]);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x1234567887654321ull) *sink = acc;
    }
    __syncthreads();
    if (CS > 1) cluster_sync();                                      // nobody exits while a peer may still multicast into it
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <int CS>
int run(const CUtensorMap& tf, const CUtensorMap& tp, int mode, int iters, int slabs_k, int slabs_r, unsigned long long* sink, int sms, const char* name) {
    auto kern = probe_kernel<CS>;
    const int smem = kStages * kSlabBytes + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = sms / CS * CS;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, kern, tf, tp, mode, iters, slabs_k, slabs_r, sink));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) CK(cudaLaunchKernelEx(&cfg, kern, tf, tp, mode, iters, slabs_k, slabs_r, sink));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    ms /= reps;
    const double delivered = static_cast<double>(grid) * iters * kSlabBytes;
    const double from_l2 = (mode == 2) ? delivered / CS : delivered;      // what the loads ask the L2 for
    printf("{\"mode\": \"%s\", \"cluster\": %d, \"ctas\": %d, \"ms\": %.4f, \"delivered_TBps\": %.3f, \"requested_from_l2_TBps\": %.3f}\n", name, CS, grid, ms,
           delivered / ms * 1e-9, from_l2 / ms * 1e-9);
    return 0;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encode encode = reinterpret_cast<PFN_encode>(fn);
    // L2-resident matrix: 8192 rows x 2240 bf16 (one operand of the self-similarity at half height, 36.7 MB)
    const int rows = 8192, cols = 2240, slabs_k = cols / kSlabCols, slabs_r = rows / kSlabRows;
    __nv_bfloat16* buf;
    CK(cudaMalloc(&buf, static_cast<size_t>(rows) * cols * 2));
    CK(cudaMemset(buf, 0, static_cast<size_t>(rows) * cols * 2));
    unsigned long long* sink;
    CK(cudaMalloc(&sink, 8));
    auto make = [&](CUtensorMap* tm, int box_rows) {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(kSlabCols), static_cast<cuuint32_t>(box_rows)};
        cuuint32_t estr[2] = {1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap tf, tp1, tp2, tp4;
    if (make(&tf, 256) || make(&tp1, 256) || make(&tp2, 128) || make(&tp4, 64)) { printf("tensor map encode failed\n"); return 1; }
    const int iters = 4000, sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"slab_bytes\": %d, \"stages\": %d, \"iters_per_cta\": %d}\n", prop.name, sms, kSlabBytes, kStages, iters);
    if (run<1>(tf, tp1, 0, iters, slabs_k, slabs_r, sink, sms, "distinct")) return 1;
    if (run<2>(tf, tp2, 0, iters, slabs_k, slabs_r, sink, sms, "distinct")) return 1;
    if (run<2>(tf, tp2, 1, iters, slabs_k, slabs_r, sink, sms, "shared-uc")) return 1;
    if (run<2>(tf, tp2, 2, iters, slabs_k, slabs_r, sink, sms, "shared-mc")) return 1;
    if (run<4>(tf, tp4, 0, iters, slabs_k, slabs_r, sink, sms, "distinct")) return 1;
    if (run<4>(tf, tp4, 1, iters, slabs_k, slabs_r, sink, sms, "shared-uc")) return 1;
    if (run<4>(tf, tp4, 2, iters, slabs_k, slabs_r, sink, sms, "shared-mc")) return 1;
    return 0;
}
