// compat.h — the behaviours of usearch 2.23.0 / diskann-rs 0.3.4 / anndists 0.1.3 that are RECALLED (their sources are
// not under the reference tree), one named constant each. The kernels read these constants
// and nothing else decides these points; leann_cuda_compat_flags() reports them with the oracle's bit layout
// (oracle/graph_oracle.cpp CompatBits) and tests/test_oracle_graph.py asserts both sides agree. If a golden file made with
// the real crates (oracle/pin_graph_golden.py -> tests/golden/graph_golden.json) disagrees, flip the constant here and
// the default in the oracle: a one-line change on each side.
#pragma once
#include <cstdint>

namespace leann {
namespace compat {
constexpr bool USEARCH_STOP_STRICT = true;          // search_to_find_in_base_: stop when cand.d > radius (false: >=)
constexpr bool DISKANN_STOP_STRICT = false;         // search_with_dists: stop when full && best.d >= worst (true: >)
constexpr bool TOP_NEWCOMER_BEFORE_EQUALS = true;   // sorted_buffer_gt::insert is a lower_bound insert
constexpr bool NEXT_FIFO_AMONG_EQUALS = true;       // equal distances leave the candidate queue in arrival order
constexpr bool DISTDOT_CLAMP_AT_ZERO = true;        // DistDot::eval = max(0, 1 - dot). anndists also ASSERTS 1 - dot >= -2e-6
                                                    // (panics on non-unit vectors); the library clamps instead.
constexpr uint32_t FLAGS = (USEARCH_STOP_STRICT ? 1u : 0u) | (DISKANN_STOP_STRICT ? 2u : 0u) | (TOP_NEWCOMER_BEFORE_EQUALS ? 4u : 0u) |
                           (NEXT_FIFO_AMONG_EQUALS ? 8u : 0u) | (DISTDOT_CLAMP_AT_ZERO ? 16u : 0u);
}  // namespace compat
}  // namespace leann
