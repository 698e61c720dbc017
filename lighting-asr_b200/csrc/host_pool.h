// Host-side staging helpers of the C ABI (include/b200fe.h, "host staging"): a small persistent thread pool that
//   * packs a LIST of utterances (what the reference's collate loop receives one by one, R/lasr/data/dataset.py:190-206;
//     float64 from soundfile.read, R/lasr/data/reader.py:24) into one pinned staging buffer, converting float64 -> float32
//     on the way (the `.float()` of WavToKaldiFbank, R/lasr/data/datatrans.py:73) with non-temporal stores, and
//   * zero-fills the padding rows of the host feature batch (pad_audio = 0, R/lasr/data/dataset.py:18)
// while the GPU works.  Plain C++11 threads; no CUDA calls in here (the SIMD loops live in host_simd.cpp).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b200fe_host {

struct Job {
    std::atomic<long long> remaining{0};
    std::atomic<bool> done{false};
    std::atomic<int> flag{0};           // set by a task that could not do what was asked (kind 3: a sample that is not a PCM16 value)
    std::function<void()> on_done;      // runs on the thread that finishes the job's last task, before waiters are released
    std::mutex m;
    std::condition_variable cv;
};

struct Task {
    int kind;                 // 0 memcpy, 1 float64 -> float32, 2 zero fill, 3 float64 holding PCM16 values -> int16 (checked)
    const void* src;
    void* dst;
    long long n;              // bytes (kinds 0, 2) or elements (kind 1)
    long long tail_zero;      // bytes to clear right after the destination range (alignment gap of the packed layout)
    std::shared_ptr<Job> job;
};

// Streaming primitives (host_simd.cpp: SSE2 baseline and an AVX-512 variant with software prefetch, picked at load time).
//   cvt_f64_f32: float64 -> float32 with non-temporal stores (the staging buffer is read next by the DMA engine, not by a core)
//   zero_stream: zero fill of (pinned) memory nobody reads from a core before the GPU / trainer does: no read-for-ownership
//                traffic (a 1 MB memset stays below glibc's non-temporal threshold and costs twice the DRAM traffic)
//   copy_stream: copy with non-temporal stores.  glibc's memcpy switches to them only far above a task's 256 kB, so a plain
//                memcpy leaves the staged waveforms dirty in the cores' caches (and reads the destination lines first): the DMA
//                engine then has to snoop them out -- measured on the int16 plug-in call: H2D chunks at 27-39 GB/s instead of 52.
void cvt_f64_f32(const double* s, float* d, long long n);
bool cvt_f64_pcm16(const double* s, short* d, long long n);     // false: a sample is not k / 32768 with an int16 k; d may be null (probe)
void zero_stream(void* dst, long long n);
void copy_stream(const void* src, void* dst, long long n);
int host_isa();          // 0 SSE2, 1 AVX-512F

inline void run_task(const Task& t)
{
    switch (t.kind) {
        case 0: copy_stream(t.src, t.dst, t.n); if (t.tail_zero > 0) memset(static_cast<char*>(t.dst) + t.n, 0, (size_t)t.tail_zero); break;
        case 1:
            cvt_f64_f32(static_cast<const double*>(t.src), static_cast<float*>(t.dst), t.n);
            if (t.tail_zero > 0) memset(static_cast<float*>(t.dst) + t.n, 0, (size_t)t.tail_zero);
            break;
        case 3:
            if (!cvt_f64_pcm16(static_cast<const double*>(t.src), static_cast<short*>(t.dst), t.n)) t.job->flag.store(1);
            if (t.tail_zero > 0) memset(static_cast<short*>(t.dst) + t.n, 0, (size_t)t.tail_zero);
            break;
        default: zero_stream(t.dst, t.n); break;
    }
}

}  // namespace b200fe_host

struct b200fe_host_pool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv;
    std::deque<b200fe_host::Task> q;
    bool stop = false;
    // CUDA device the completion hooks (uploads) use: every worker binds it ONCE, before its first task after it was set -- the
    // first runtime call of a thread attaches the primary context, which costs milliseconds and must not sit inside a batch
    std::atomic<int> device{-1};
    void (*bind_device)(int) = nullptr;
    long long next_ticket = 1;
    std::map<long long, std::shared_ptr<b200fe_host::Job>> jobs;

    explicit b200fe_host_pool(int n)
    {
        for (int i = 0; i < n; ++i) threads.emplace_back([this] { worker(); });
    }
    ~b200fe_host_pool()
    {
        { std::lock_guard<std::mutex> g(m); stop = true; }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
    static void finish(const b200fe_host::Task& t)
    {
        if (t.job->remaining.fetch_sub(1) == 1) {
            if (t.job->on_done) t.job->on_done();
            std::lock_guard<std::mutex> g(t.job->m);
            t.job->done = true;
            t.job->cv.notify_all();
        }
    }
    void worker()
    {
        int bound = -1;
        for (;;) {
            b200fe_host::Task t;
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [this] { return stop || !q.empty(); });
                if (q.empty()) return;
                t = q.front(); q.pop_front();
            }
            const int want = device.load();
            if (want >= 0 && want != bound && bind_device) { bind_device(want); bound = want; }
            b200fe_host::run_task(t);
            finish(t);
        }
    }
    long long submit(std::vector<b200fe_host::Task>& tasks, std::function<void()> on_done = nullptr)
    {
        auto job = std::make_shared<b200fe_host::Job>();
        job->remaining = (long long)tasks.size();
        job->on_done = std::move(on_done);
        for (auto& t : tasks) t.job = job;
        long long ticket;
        {
            std::lock_guard<std::mutex> g(m);
            ticket = next_ticket++;
            jobs[ticket] = job;
            for (auto& t : tasks) q.push_back(t);
        }
        cv.notify_all();
        return ticket;
    }
    // The waiting thread helps: it drains queued tasks (of any job, FIFO) until its own job is complete.
    int wait(long long ticket, int* flag = nullptr)
    {
        std::shared_ptr<b200fe_host::Job> job;
        {
            std::lock_guard<std::mutex> g(m);
            auto it = jobs.find(ticket);
            if (it == jobs.end()) return -1;
            job = it->second;
            jobs.erase(it);
        }
        while (!job->done.load()) {
            b200fe_host::Task t;
            bool have = false;
            {
                std::lock_guard<std::mutex> g(m);
                if (!q.empty()) { t = q.front(); q.pop_front(); have = true; }
            }
            if (have) { b200fe_host::run_task(t); finish(t); continue; }
            std::unique_lock<std::mutex> g(job->m);
            job->cv.wait(g, [&] { return job->done.load(); });
        }
        if (flag) *flag = job->flag.load();
        return 0;
    }
};
