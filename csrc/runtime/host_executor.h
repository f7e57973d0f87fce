// HostExecutor core: one OS worker thread per device slot, FIFO queue per worker, error capture, join-all `sync`.
// Header-only and free of CUDA / Python so that the same code is (a) wrapped by runtime.cpp, where the tasks are
// cudaGraphLaunch / cudaStreamSynchronize calls issued without the GIL, and (b) compiled stand-alone with
// -fsanitize=thread by tests/native/host_executor_tsan.cpp (SURVEY §5 "TSAN build of the C++ host runtime").
#pragma once

#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace pa {
namespace rt {

class HostExecutorCore {
 public:
  // `on_thread_start(slot)` runs once on every worker thread before it takes work (runtime.cpp: cudaSetDevice)
  explicit HostExecutorCore(int n_slots, std::function<void(int)> on_thread_start = nullptr) {
    for (int i = 0; i < n_slots; ++i) {
      workers_.emplace_back(new Worker());
      Worker* w = workers_.back().get();
      w->thread = std::thread([w, i, on_thread_start] {
        if (on_thread_start) on_thread_start(i);
        std::unique_lock<std::mutex> lk(w->mu);
        for (;;) {
          w->cv.wait(lk, [w] { return w->stop || !w->q.empty(); });
          if (w->q.empty()) return;                       // stop requested and nothing left to run
          auto fn = std::move(w->q.front());
          w->q.pop();
          lk.unlock();
          std::string err;
          try {
            fn();
          } catch (const std::exception& e) {
            err = e.what();
          } catch (...) {
            err = "unknown exception in a HostExecutor task";
          }
          lk.lock();
          if (!err.empty() && w->error.empty()) w->error = std::move(err);
          if (--w->pending == 0) w->done.notify_all();
        }
      });
    }
  }
  ~HostExecutorCore() { shutdown(); }
  HostExecutorCore(const HostExecutorCore&) = delete;
  HostExecutorCore& operator=(const HostExecutorCore&) = delete;

  void submit(int slot, std::function<void()> fn) {
    if (slot < 0 || slot >= static_cast<int>(workers_.size())) throw std::runtime_error("[pa.rt] bad executor slot");
    Worker* w = workers_[slot].get();
    {
      std::lock_guard<std::mutex> g(w->mu);
      if (w->stop) throw std::runtime_error("[pa.rt] executor already shut down");
      w->q.push(std::move(fn));
      ++w->pending;
    }
    w->cv.notify_one();
  }

  // Wait until every queue has drained; rethrows (and clears) the first recorded task error.
  void sync() {
    std::string err;
    for (auto& w : workers_) {
      std::unique_lock<std::mutex> lk(w->mu);
      w->done.wait(lk, [&] { return w->pending == 0; });
      if (!w->error.empty()) {
        if (err.empty()) err = w->error;
        w->error.clear();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
  }

  void shutdown() {
    for (auto& w : workers_) {
      {
        std::lock_guard<std::mutex> g(w->mu);
        w->stop = true;
      }
      w->cv.notify_all();
    }
    for (auto& w : workers_)
      if (w->thread.joinable()) w->thread.join();
    workers_.clear();
  }

  int size() const { return static_cast<int>(workers_.size()); }

 private:
  struct Worker {
    std::thread thread;
    std::mutex mu;
    std::condition_variable cv, done;
    std::queue<std::function<void()>> q;
    int pending = 0;
    bool stop = false;
    std::string error;
  };
  std::vector<std::unique_ptr<Worker>> workers_;
};

}  // namespace rt
}  // namespace pa
