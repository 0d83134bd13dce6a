/**
 * @file HostExpandTest.cpp
 * Host-only checks of csrc/host_expand.{h,cpp}: the expansion of the compact control-matrix
 * download into dense iDynTree::Matrix6x6 blocks (every ISA form against a plain loop, bit for bit,
 * sign of zero included, guard bands untouched) and the worker pool (jobs after a resize must never
 * re-run an earlier job).  `--bandwidth` prints the expansion rate for 1..N threads.
 */
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../csrc/host_expand.h"

using namespace blfccm;

static int failures = 0;
#define CHECK(cond)                                                       \
    do {                                                                  \
        if (!(cond)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                   \
        }                                                                 \
    } while (0)

static void naive(const double* s, double* d, long long n)
{
    for (long long i = 0; i < n; ++i, s += 8, d += 36) {
        for (int e = 0; e < 36; ++e) d[e] = 0.0;
        d[0] = d[7] = d[14] = s[0];
        d[21] = s[1]; d[22] = s[2]; d[23] = s[3];
        d[27] = s[2]; d[28] = s[4]; d[29] = s[5];
        d[33] = s[3]; d[34] = s[5]; d[35] = s[6];
    }
}

static void test_expand()
{
    std::mt19937_64 gen(7);
    std::uniform_real_distribution<double> u(-1e3, 1e3);
    for (long long n : {0LL, 1LL, 2LL, 3LL, 31LL, 32LL, 1000LL, 4097LL}) {
        std::vector<double> src(n * 8 + 8);
        for (auto& x : src) x = u(gen);
        for (long long i = 0; i < n; ++i) src[i * 8 + 7] = 0.0;
        if (n > 2) src[2 * 8 + 0] = -0.0;   // a negative zero VALUE must survive; structural zeros are +0.0
        std::vector<double> want(n * 36 + 1);
        naive(src.data(), want.data(), n);
        const int guard = 16;
        for (int shift : {0, 1, 2, 4}) {   // destination alignment: 64, 8, 16, 32 bytes
            void* raw = nullptr;
            CHECK(posix_memalign(&raw, 64, (n * 36 + 2 * guard + 8) * sizeof(double)) == 0);
            double* base = static_cast<double*>(raw);
            for (int variant = 0; variant < 2; ++variant) {
                for (long long i = 0; i < n * 36 + 2 * guard + 8; ++i) base[i] = -7.25;
                double* dst = base + guard + shift;
                if (variant == 0) expand_ctrl(src.data(), dst, n);
                else expand_ctrl_sse2(src.data(), dst, n);
                CHECK(std::memcmp(dst, want.data(), n * 36 * sizeof(double)) == 0);
                for (int gidx = 0; gidx < guard + shift; ++gidx) CHECK(base[gidx] == -7.25);
                for (int gidx = 0; gidx < guard; ++gidx) CHECK(dst[n * 36 + gidx] == -7.25);
            }
            std::free(raw);
        }
    }
}

static void test_pool()
{
    HostPool pool;
    for (int round = 0; round < 50; ++round) {
        const int threads = 1 + (round * 3) % 7;
        pool.resize(threads);
        {
            // the job's captures live on this block's stack only: a worker that re-ran it later
            // (the bug a resize after a finished job would trigger) would scribble over dead memory
            std::atomic<int> hits{0};
            std::vector<int> seen(threads, 0);
            pool.start([&](int j) {
                seen[j] += 1;
                hits.fetch_add(1);
            });
            pool.wait();
            CHECK(hits.load() == threads);
            for (int j = 0; j < threads; ++j) CHECK(seen[j] == 1);
        }
    }
    pool.resize(0);
    pool.resize(3);
    std::atomic<int> hits{0};
    pool.start([&](int) { hits.fetch_add(1); });
    pool.wait();
    CHECK(hits.load() == 3);
}

static void bandwidth(int max_threads)
{
    const long long n = 1LL << 21;
    void *s = nullptr, *d = nullptr;
    if (posix_memalign(&s, 64, n * 8 * sizeof(double)) || posix_memalign(&d, 64, n * 36 * sizeof(double))) return;
    double* src = static_cast<double*>(s);
    double* dst = static_cast<double*>(d);
    for (long long i = 0; i < n * 8; ++i) src[i] = 1.0 + i % 7;
    std::memset(dst, 0, n * 36 * sizeof(double));
    std::printf("expand_ctrl isa: %s; %lld contacts (%.0f MB dense)\n", expand_ctrl_isa(), n, n * 288 / 1e6);
    HostPool pool;
    for (int threads = 1; threads <= max_threads; threads *= 2) {
        pool.resize(threads);
        for (int form = 0; form < 2; ++form) {
            double best = 1e30;
            for (int rep = 0; rep < 5; ++rep) {
                const auto t0 = std::chrono::steady_clock::now();
                pool.start([&](int j) {
                    const long long lo = (n * j / threads) & ~1LL, hi = j == threads - 1 ? n : ((n * (j + 1) / threads) & ~1LL);
                    if (form == 0) expand_ctrl(src + lo * 8, dst + lo * 36, hi - lo);
                    else expand_ctrl_sse2(src + lo * 8, dst + lo * 36, hi - lo);
                });
                pool.wait();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            std::printf("  %2d threads, %s: %.2f ms = %.1f GB/s written, %.0f M contacts/s\n", threads,
                        form == 0 ? expand_ctrl_isa() : "sse2 (forced)", best * 1e3, n * 288 / best / 1e9, n / best / 1e6);
        }
    }
    std::free(s);
    std::free(d);
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::strcmp(argv[1], "--bandwidth") == 0) {
        bandwidth(argc > 2 ? std::atoi(argv[2]) : 8);
        return 0;
    }
    test_expand();
    test_pool();
    std::printf("%s (expand_ctrl isa: %s)\n", failures ? "FAILED" : "All tests passed", expand_ctrl_isa());
    return failures ? 1 : 0;
}
