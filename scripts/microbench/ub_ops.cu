// ub_ops.cu — micro-benchmarks behind the count-stage design: what one k-mer insertion may cost
// on a B200 SM.  Every test runs 148*R CTAs of 512 threads (16 warps), each lane doing ITER operations on random
// addresses inside a warp-private region of shared memory (or a global region that fits L2), and prints the chip-wide
// rate in G lane-operations per second and in cycles per warp instruction per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub_ops ub_ops.cu && ./ub_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

constexpr int kThreads = 512, kWarps = 16, kIter = 4096;
constexpr int kWarpWords = 2048;                       // 8 KB of u32 per warp region

template <int OP>
__global__ void __launch_bounds__(kThreads) k_smem(unsigned long long* sink) {
    extern __shared__ __align__(16) uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* reg = sm + warp * kWarpWords;
    for (int i = lane; i < kWarpWords; i += 32) reg[i] = 0;
    __syncwarp();
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
    unsigned long long acc = 0;
    unsigned long long* reg64 = reinterpret_cast<unsigned long long*>(reg);
#pragma unroll 4
    for (int it = 0; it < kIter; it++) {
        const uint32_t r = lcg(s);
        if (OP == 0) { atomicAdd(&reg[r & (kWarpWords - 1)], (r >> 12) | 1u); }                       // ATOMS.ADD, variable addend, no return
        else if (OP == 1) { acc += atomicAdd(&reg[r & (kWarpWords - 1)], 1u); }                       // ATOMS.ADD with return
        else if (OP == 2) { acc += atomicCAS(&reg64[r & (kWarpWords / 2 - 1)], 0ull, (unsigned long long)r); }   // ATOMS.CAS.64
        else if (OP == 3) { acc += __match_any_sync(0xFFFFFFFFu, (unsigned long long)(r & 63u) | ((unsigned long long)r << 40)); }   // MATCH.ANY.U64, ~40 % duplicates
        else if (OP == 4) { acc += reg64[r & (kWarpWords / 2 - 1)]; }                                 // LDS.64 random
        else if (OP == 5) { reg64[r & (kWarpWords / 2 - 1)] = acc + r; }                              // STS.64 random
        else if (OP == 6) { uint32_t v = reg[r & (kWarpWords - 1)]; reg[r & (kWarpWords - 1)] = v + 1u; acc += v; }   // LDS.32 + STS.32 (plain read-modify-write)
        else if (OP == 7) { acc += __shfl_sync(0xFFFFFFFFu, r, (r >> 3) & 31); }                      // SHFL.IDX
        else if (OP == 8) { acc += __match_any_sync(0xFFFFFFFFu, r & 63u); }                          // MATCH.ANY.U32
        else if (OP == 9) { acc += __reduce_or_sync(0xFFFFFFFFu, r); }                                // REDUX.OR
        else if (OP == 10) { const uint4 v = reinterpret_cast<const uint4*>(reg)[r & (kWarpWords / 4 - 1)]; acc += v.x + v.w; }   // LDS.128 random
        else if (OP == 11) { acc += __ballot_sync(0xFFFFFFFFu, r & 1u); }                             // VOTE
        else if (OP == 12) { atomicAdd(&reg[r & (kWarpWords - 1)], 1u); }                             // ATOMS.POPC.INC
    }
    if (acc == 0x123456789ull) sink[0] = acc;
    __syncwarp();
    if (lane == 0 && reg[0] == 0xFFFFFFFFu) sink[1] = 1;
}

// global-memory operations on a region of `words` u64 (L2-resident when it fits)
template <int OP>
__global__ void __launch_bounds__(kThreads) k_gmem(unsigned long long* tbl, uint32_t mask, unsigned long long* sink) {
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 777u;
    unsigned long long acc = 0;
#pragma unroll 4
    for (int it = 0; it < kIter / 4; it++) {
        const uint32_t r = lcg(s) * 4099u + (lcg(s) << 12);
        unsigned long long* p = tbl + (r & mask);
        if (OP == 0) { atomicAdd(reinterpret_cast<unsigned int*>(p), 1u); }                           // RED.ADD.32
        else if (OP == 1) { acc += atomicCAS(p, 0x1234ull, (unsigned long long)r); }                  // ATOMG.CAS.64 (never succeeds: table stays zero)
        else if (OP == 2) { acc += __ldcg(p); }                                                       // LDG.64 random, independent
        else if (OP == 3) { const unsigned long long v = __ldcg(p); if (v == 0ull) atomicAdd(reinterpret_cast<unsigned int*>(p) + 1, 1u); acc += v; }   // read, then RED on the same sector
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

template <typename F>
static float time_ms(F&& launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms / reps;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int n_sm = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock attr %.0f MHz\n", p.name, n_sm, khz / 1e3);
    unsigned long long* sink; CK(cudaMalloc(&sink, 64));
    const size_t smem = (size_t)kWarps * kWarpWords * 4;
    const char* names[] = {"ATOMS.ADD.32 (red)", "ATOMS.ADD.32 (ret)", "ATOMS.CAS.64", "MATCH.ANY.U64", "LDS.64 random", "STS.64 random",
                           "LDS.32+STS.32 rmw", "SHFL.IDX", "MATCH.ANY.U32", "REDUX.OR", "LDS.128 random", "VOTE.BALLOT", "ATOMS.POPC.INC"};
    void (*ks[])(unsigned long long*) = {k_smem<0>, k_smem<1>, k_smem<2>, k_smem<3>, k_smem<4>, k_smem<5>, k_smem<6>, k_smem<7>, k_smem<8>, k_smem<9>, k_smem<10>, k_smem<11>, k_smem<12>};
    for (int op = 0; op < 13; op++) {
        CK(cudaFuncSetAttribute(ks[op], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = n_sm;                                     // one 16-warp CTA per SM
        const float ms = time_ms([&]() { ks[op]<<<grid, kThreads, smem>>>(sink); }, 5);
        CK(cudaGetLastError());
        const double ops = (double)grid * kThreads * kIter;
        const double warp_instr_per_sm = (double)kWarps * kIter;
        printf("smem %-22s %8.3f ms  %8.1f G lane-ops/s   %6.2f cyc per warp-instr per SM (at 1.9 GHz, incl. loop overhead)\n",
               names[op], ms, ops / ms / 1e6, ms * 1e-3 * 1.9e9 / warp_instr_per_sm);
    }
    const char* gn[] = {"RED.ADD.32", "ATOMG.CAS.64", "LDG.64", "LDG.64 + RED same sector"};
    void (*kg[])(unsigned long long*, uint32_t, unsigned long long*) = {k_gmem<0>, k_gmem<1>, k_gmem<2>, k_gmem<3>};
    for (int sz = 0; sz < 3; sz++) {
        const size_t words = sz == 0 ? (1u << 22) : sz == 1 ? (1u << 24) : (1u << 28);     // 32 MB, 128 MB, 2 GB
        unsigned long long* tbl; CK(cudaMalloc(&tbl, words * 8)); CK(cudaMemset(tbl, 0, words * 8));
        for (int op = 0; op < 4; op++) {
            const int grid = n_sm * 4;
            const float ms = time_ms([&]() { kg[op]<<<grid, kThreads>>>(tbl, (uint32_t)(words - 1), sink); }, 3);
            CK(cudaGetLastError());
            const double ops = (double)grid * kThreads * (kIter / 4);
            printf("gmem %5zu MB %-26s %8.3f ms  %8.1f G lane-ops/s\n", words * 8 >> 20, gn[op], ms, ops / ms / 1e6);
        }
        CK(cudaFree(tbl));
    }
    return 0;
}
