// host_expand.h -- host-side helpers of blf_ccm_eval_batch_host: a small worker pool and the
// expansion of the compact control-matrix download into dense iDynTree::Matrix6x6 blocks.
// Plain C++17 + SSE2, no CUDA: also compiled on its own by tests/test_host_expand.py.
#pragma once

#include <emmintrin.h>

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace blfccm {

// Worker threads of the host-buffer entry point: they expand the compact control-matrix download
// (8 doubles per contact) into the caller's dense Matrix6x6 array while later chunks are still on
// the PCIe link.  Data-format work only; no contact-model arithmetic runs on the host.
class HostPool {
public:
    ~HostPool() { stop(); }
    int size() const { return static_cast<int>(th_.size()); }
    void resize(int n)
    {
        if (n == size()) return;
        stop();
        quit_ = false;
        // a new worker must not mistake the job of an earlier call (whose captures are gone) for a
        // fresh one: it starts from the current generation
        const unsigned long long now = gen_;
        for (int j = 0; j < n; ++j) th_.emplace_back([this, j, now] { loop(j, now); });
    }
    // run job(j) on every worker; returns at once, wait() blocks until all are done
    void start(std::function<void(int)> job)
    {
        std::lock_guard<std::mutex> lk(m_);
        job_ = std::move(job);
        active_ = size();
        ++gen_;
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return active_ == 0; });
    }

private:
    void stop()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_ = true;
            cv_.notify_all();
        }
        for (auto& t : th_) t.join();
        th_.clear();
    }
    void loop(int j, unsigned long long seen)
    {
        for (;;) {
            std::function<void(int)> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_;
                job = job_;
            }
            job(j);
            job = nullptr;
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) {
                job_ = nullptr;   // its captures die with the call that started it
                done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(int)> job_;
    unsigned long long gen_ = 0;
    int active_ = 0;
    bool quit_ = false;
};

// compact control matrix {gd, gs_xx, gs_xy, gs_xz, gs_yy, gs_yz, gs_zz, pad} -> dense row-major 6x6
// (iDynTree::Matrix6x6; ContinuousContactModel.cpp:165-170: the top-left diagonal and the symmetric
// bottom-right block, every other entry the +0.0 left by the constructor's zero(), :18).  Data
// movement only, implemented in host_expand.cpp: AVX-512 full-line non-temporal stores when the CPU
// has them and the destination is 64-byte aligned, SSE2 otherwise (bit-identical results).
void expand_ctrl(const double* src, double* dst, long long cnt);
// the portable form on its own (tests compare the two)
void expand_ctrl_sse2(const double* src, double* dst, long long cnt);
// which form expand_ctrl picks for a 64-byte aligned destination: "avx512" or "sse2"
const char* expand_ctrl_isa();

}  // namespace blfccm
