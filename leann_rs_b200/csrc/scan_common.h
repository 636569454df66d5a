// scan_common.h — shared between exact_scan.cu (f32 tiles, select, merge) and exact_scan_tc.cu (tcgen05 pass).
#pragma once
#include "internal.h"

namespace leann {

__host__ __device__ __forceinline__ uint32_t scan_order_f32_bits(uint32_t u) { return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t scan_order_f32(float f) { return scan_order_f32_bits(__float_as_uint(f)); }
__device__ __forceinline__ float scan_unorder_f32(uint32_t u) {
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}
#endif

struct ScanScratch {
    unsigned long long* cand;      // [nq][cap] packed rank keys
    uint32_t* cand_cnt;            // [nq]
    unsigned long long* best;      // [nq][kpad]
    uint32_t* best_cnt;            // [nq]
    unsigned long long* thr;       // [nq] packed key of the current k-th best (~0 = none yet)
    uint32_t* overflow;            // [1]
    float4* qpad;                  // [nq][d4]
};

struct TcIndexView {               // per-index tensor-path data (built once)
    const void* x_bf16;            // [n][dp8] bf16
    const uint32_t* xmax_bits;     // [0] max row norm |x|, [1] max bf16 rounding residual |x^ - x| (f32 bits)
    uint32_t dp8;
};
struct TcScratch {                 // per-call tensor-path scratch
    void* q_bf16;                  // [nq][dp8] bf16
    float* qnorm;                  // [nq] |q|
    float* qres;                   // [nq] |q^ - q| (bf16 rounding residual)
    float* thr_dot;                // [nq]
    float* cut_slack;              // [nq] slack of the re-rank's approximate cut (dot_threshold_kernel)
    uint32_t* cand_ids;            // [nq][cap]
    float* cand_sc;                // [nq][cap] tensor-core score of each survivor
};

bool exact_scan_tc_supported(const FlatView& f, uint32_t nq);
uint32_t exact_scan_tc_dp8(uint32_t d, uint32_t d4, int metric);   // columns of the bf16 copies (L2: + the |x|^2/2 group)
void exact_scan_tc_prepare(const float4* vecs, size_t n, uint32_t d4, uint32_t dp8, void* bf16_rows, float* norms,
                           uint32_t* xmax_bits, int metric, cudaStream_t s);
void exact_scan_tc_queries(const float4* qpad, uint32_t nq, uint32_t d4, uint32_t dp8, int metric, const uint32_t* xmax_bits,
                           const TcScratch& ts, cudaStream_t stream);
void exact_scan_tc_canonical_first(const FlatView& f, const ScanScratch& s, const TcScratch& ts, uint32_t nq, uint32_t kpad, uint32_t cap,
                                   cudaStream_t stream);
void exact_scan_tc_round(const FlatView& f, const TcIndexView& tv, const ScanScratch& s, const TcScratch& ts, uint32_t nq, uint32_t k,
                         uint32_t r0, uint32_t r1, const uint64_t* d_mask, uint32_t cap, int sms, cudaStream_t stream);

}  // namespace leann
