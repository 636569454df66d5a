// gather_probe.cu — ceiling of RANDOM ROW GATHERS on this GPU, the access pattern of the graph traversal (K1): every
// 8-lane group (row_bytes <= 1024) or warp (longer rows) reads one whole row of a [n][row_bytes] array at a pseudo-random
// index with LDG.128 (L1 no-allocate, as K1 does), U rows in flight per group, nothing else. What it reports is the rate at
// which HBM + L2 deliver scattered rows of that length with every latency hidden — the roofline K1's "algorithmic GB/s" can be
// held against for short rows, where the copy bandwidth (sequential pages) is not reachable.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/gather_probe benchmarks/gather_probe.cu
//   benchmarks/gather_probe [total_GB=4.8]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// LPV lanes per row, VPL float4 per lane (row = LPV * VPL * 16 bytes), U rows in flight per group.
template <int LPV, int VPL, int U>
__global__ void gather_kernel(const float4* __restrict__ base, uint32_t n_rows, uint32_t rows_per_group, float* __restrict__ sink) {
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPV;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / LPV;
    float acc = 0.f;
    for (uint32_t it = 0; it < rows_per_group; it += U) {
        float4 v[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = mix(group * 0x9E3779B1u + it + u) % n_rows;
            const float4* row = base + (size_t)r * (LPV * VPL);
#pragma unroll
            for (int i = 0; i < VPL; ++i) v[u][i] = ldg_stream(row + i * LPV + sub);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc += v[u][i].x + v[u][i].y + v[u][i].z + v[u][i].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int LPV, int VPL, int U>
double run(const float4* base, size_t total_bytes, int warps_per_sm, int sms, float* sink) {
    const size_t row_bytes = (size_t)LPV * VPL * 16;
    const uint32_t n_rows = (uint32_t)(total_bytes / row_bytes);
    const int threads = 128, blocks = sms * warps_per_sm / 4;
    const size_t groups = (size_t)blocks * threads / LPV;
    const size_t target = (size_t)24 << 30;   // ~24 GB gathered per launch
    uint32_t rpg = (uint32_t)(target / row_bytes / groups);
    rpg = rpg / U * U;
    if (rpg < (uint32_t)U) rpg = U;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    gather_kernel<LPV, VPL, U><<<blocks, threads>>>(base, n_rows, rpg, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        gather_kernel<LPV, VPL, U><<<blocks, threads>>>(base, n_rows, rpg, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    const double bytes = (double)groups * rpg * row_bytes;
    return bytes / best / 1e6;   // GB/s
}

__global__ void fill_kernel(float4* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int main(int argc, char** argv) {
    const double gb = argc > 1 ? atof(argv[1]) : 4.8;
    const size_t total = (size_t)(gb * 1e9) / 4096 * 4096;
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float4* base; float* sink;
    CK(cudaMalloc(&base, total)); CK(cudaMalloc(&sink, 4));
    fill_kernel<<<(unsigned)((total / 16 + 255) / 256), 256>>>(base, total / 16);
    CK(cudaDeviceSynchronize());
    // sequential copy-like read for reference: rows of 4096 B are close to streaming
    printf("{\"array_gb\": %.2f, \"sms\": %d, \"results\": [\n", total / 1e9, sms);
    bool first = true;
    auto emit = [&](int row_bytes, int u, int wps, double gbs) {
        printf("%s {\"row_bytes\": %d, \"rows_in_flight_per_group\": %d, \"warps_per_sm\": %d, \"gbs\": %.1f}", first ? "" : ",\n", row_bytes, u, wps, gbs);
        first = false;
        fflush(stdout);
    };
    for (int wps : {16, 24, 32, 48, 64}) {
        emit(128, 4, wps, run<8, 1, 4>(base, total, wps, sms, sink));
        emit(128, 8, wps, run<8, 1, 8>(base, total, wps, sms, sink));
        emit(256, 4, wps, run<8, 2, 4>(base, total, wps, sms, sink));
        emit(256, 8, wps, run<8, 2, 8>(base, total, wps, sms, sink));
        emit(384, 2, wps, run<8, 3, 2>(base, total, wps, sms, sink));
        emit(384, 4, wps, run<8, 3, 4>(base, total, wps, sms, sink));
        emit(384, 8, wps, run<8, 3, 8>(base, total, wps, sms, sink));
        emit(512, 4, wps, run<8, 4, 4>(base, total, wps, sms, sink));
        emit(1024, 4, wps, run<8, 8, 4>(base, total, wps, sms, sink));
        emit(1536, 2, wps, run<32, 3, 2>(base, total, wps, sms, sink));
        emit(3072, 2, wps, run<32, 6, 2>(base, total, wps, sms, sink));
        emit(3072, 4, wps, run<32, 6, 4>(base, total, wps, sms, sink));
    }
    printf("\n]}\n");
    return 0;
}
