// Multi-threaded `yams add` shape: T host threads, each chunk+hash-ing its own files through the C ABI.
//   g++ -std=c++17 -O2 -pthread tools/mt_ingest_bench.cpp -I include -L yams_b200 -lyams_b200 -Wl,-rpath,$PWD/yams_b200 -o /tmp/mt_ingest
//   /tmp/mt_ingest <file_bytes> <threads> <calls_per_thread>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "yams_b200.h"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    size_t bytes = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1u << 20);
    int threads = argc > 2 ? atoi(argv[2]) : 4;
    int calls = argc > 3 ? atoi(argv[3]) : 50;
    if (yams_plugin_init("{}", nullptr) != YAMS_PLUGIN_OK) { fprintf(stderr, "init failed: %s\n", yams_b200_last_error()); return 2; }
    yams_cdc_config cfg;
    yams_b200_cdc_default_config(&cfg);
    std::vector<std::vector<uint8_t>> bufs(threads, std::vector<uint8_t>(bytes));
    for (int t = 0; t < threads; ++t) {
        uint64_t x = 0x9E3779B97F4A7C15ull * (t + 1);
        for (size_t i = 0; i < bytes; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; bufs[t][i] = (uint8_t)x; }
    }
    {   // warm-up on the main thread
        yams_chunk_desc* d = nullptr; size_t n = 0;
        for (int i = 0; i < 3; ++i) { yams_b200_chunk_and_hash(nullptr, bufs[0].data(), bytes, &cfg, &d, &n); yams_b200_free_chunks(nullptr, d, n); }
    }
    if (argc > 4 && !strcmp(argv[4], "batch")) {
        // one chunk_and_hash_batch call over threads * calls files (the same buffers reused as distinct files)
        size_t nfiles = (size_t)threads * calls;
        std::vector<const uint8_t*> ptrs(nfiles);
        std::vector<size_t> lens(nfiles, bytes);
        for (size_t i = 0; i < nfiles; ++i) ptrs[i] = bufs[i % threads].data();
        std::vector<uint64_t> first(nfiles + 1);
        double best = 1e30;
        size_t nchunks = 0;
        for (int rep = 0; rep < 3; ++rep) {
            yams_chunk_desc* d = nullptr; size_t n = 0;
            double a = now_ms();
            if (yams_b200_chunk_and_hash_batch(nullptr, ptrs.data(), lens.data(), nfiles, &cfg, &d, &n, first.data()) != YAMS_OK) {
                fprintf(stderr, "batch failed: %s\n", yams_b200_last_error());
                return 3;
            }
            double dt = now_ms() - a;
            if (dt < best) best = dt;
            nchunks = n;
            yams_b200_free_chunks(nullptr, d, n);
        }
        float ms[8];
        yams_b200_ingest_last_timings(nullptr, ms);
        printf("| %zu KiB | batch of %zu files | %.3f ms per call | %.0f files/s | %.2f GB/s | %zu chunks | scan %.3f select %.3f sha %.3f dev-total %.3f |\n",
               bytes >> 10, nfiles, best, nfiles / (best / 1e3), (double)bytes * nfiles / best / 1e6, nchunks, ms[0], ms[1], ms[2], ms[3]);
        return 0;
    }
    std::vector<double> tmax(threads, 0), tsum(threads, 0);
    std::vector<std::vector<float>> last(threads, std::vector<float>(8, 0.f));
    std::atomic<int> errors{0};
    double t0 = now_ms();
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t)
        ts.emplace_back([&, t] {
            for (int i = 0; i < calls; ++i) {
                yams_chunk_desc* d = nullptr; size_t n = 0;
                double a = now_ms();
                if (yams_b200_chunk_and_hash(nullptr, bufs[t].data(), bytes, &cfg, &d, &n) != YAMS_OK) ++errors;
                double dt = now_ms() - a;
                tsum[t] += dt;
                if (dt > tmax[t]) tmax[t] = dt;
                yams_b200_free_chunks(nullptr, d, n);
            }
            yams_b200_ingest_last_timings(nullptr, last[t].data());
        });
    for (auto& th : ts) th.join();
    double wall = now_ms() - t0;
    double avg = 0, mx = 0;
    for (int t = 0; t < threads; ++t) { avg += tsum[t] / calls / threads; if (tmax[t] > mx) mx = tmax[t]; }
    printf("| %zu KiB | %d | %d | %.3f | %.3f | %.0f | %.2f | scan %.3f select %.3f sha %.3f dev-total %.3f sync1 %.3f sync2 %.3f process %.3f | err %d |\n",
           bytes >> 10, threads, calls * threads, avg, mx, calls * threads / (wall / 1e3), (double)bytes * calls * threads / wall / 1e6, last[0][0], last[0][1],
           last[0][2], last[0][3], last[0][4], last[0][5], last[0][7], errors.load());
    return 0;
}
