// Native driver for the C ABI (what a Rust/C++ host does): T threads issue single-query
// leann_cuda_search calls on ONE shared handle — the access pattern of `leann serve`
// (src/cli/serve.rs:84,260-311) — with and without request coalescing, and every answer is checked
// against the batched call. Usage: coalesce_driver <base_path> <dims> <threads> <per_thread>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/leann_cuda.h"

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage\n"); return 2; }
    const char* base = argv[1];
    size_t d = (size_t)atol(argv[2]);
    int T = atoi(argv[3]), per = atoi(argv[4]);
    char err[1024];
    leann_cuda_index* ix = nullptr;
    if (leann_cuda_open(base, LEANN_BACKEND_HNSW, d, LEANN_METRIC_DEFAULT, 0, &ix, err, sizeof err) != LEANN_OK) { fprintf(stderr, "open: %s\n", err); return 1; }
    const size_t nq = (size_t)T * per, k = 10, ef = 64;
    std::vector<float> q(nq * d);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < nq; ++i) {
        double nrm = 0;
        for (size_t j = 0; j < d; ++j) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            float v = (float)((double)(s >> 11) / 9007199254740992.0 - 0.5);
            q[i * d + j] = v; nrm += (double)v * v;
        }
        for (size_t j = 0; j < d; ++j) q[i * d + j] = (float)(q[i * d + j] / std::sqrt(nrm));
    }
    std::vector<uint64_t> want_k(nq * k), got_k(nq * k);
    std::vector<float> want_d(nq * k), got_d(nq * k);
    std::vector<uint32_t> want_c(nq), got_c(nq);
    if (leann_cuda_search(ix, q.data(), nq, k, ef, nullptr, 0, want_k.data(), want_d.data(), want_c.data(), err, sizeof err) != LEANN_OK) { fprintf(stderr, "batch: %s\n", err); return 1; }
    double qps[2] = {0, 0};
    uint64_t batches = 0, requests = 0;
    int bad = 0;
    for (int mode = 0; mode < 2; ++mode) {
        leann_cuda_set_coalescing(ix, mode ? 512 : 0, mode ? 200 : 0);
        std::atomic<int> failures{0};
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                char e2[256];
                for (int i = 0; i < per; ++i) {
                    size_t qi = (size_t)t * per + i;
                    if (leann_cuda_search(ix, &q[qi * d], 1, k, ef, nullptr, 0, &got_k[qi * k], &got_d[qi * k], &got_c[qi], e2, sizeof e2) != LEANN_OK) failures++;
                }
            });
        for (auto& x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        qps[mode] = nq / dt;
        bad += failures.load();
        if (memcmp(want_k.data(), got_k.data(), nq * k * 8) || memcmp(want_d.data(), got_d.data(), nq * k * 4) || memcmp(want_c.data(), got_c.data(), nq * 4)) bad += 1000000;
        std::fill(got_k.begin(), got_k.end(), 0);
    }
    leann_cuda_coalescing_stats(ix, &batches, &requests);
    printf("{\"threads\": %d, \"requests\": %zu, \"mismatch_or_fail\": %d, \"qps_uncoalesced\": %.0f, \"qps_coalesced\": %.0f, \"batches\": %llu, \"coalesced_requests\": %llu}\n",
           T, nq, bad, qps[0], qps[1], (unsigned long long)batches, (unsigned long long)requests);
    leann_cuda_close(ix);
    return bad ? 1 : 0;
}
