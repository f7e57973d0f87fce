// ThreadSanitizer stress of the native host runtime's executor (csrc/runtime/host_executor.h): several producer threads
// submit to several worker slots concurrently, interleaved with sync() calls, task errors and a shutdown while work is
// still queued.  Built and run by tests/test_native_tsan.py with  g++ -fsanitize=thread.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../csrc/runtime/host_executor.h"

int main() {
  using pa::rt::HostExecutorCore;
  constexpr int kSlots = 4, kProducers = 6, kRounds = 400;
  std::atomic<int> started{0};
  std::vector<long long> counters(kSlots, 0);          // each only touched by its own worker thread (FIFO per slot)
  std::atomic<long long> total{0};
  int failures = 0;
  {
    HostExecutorCore ex(kSlots, [&](int) { started.fetch_add(1); });
    std::vector<std::thread> producers;
    for (int p = 0; p < kProducers; ++p) {
      producers.emplace_back([&, p] {
        for (int r = 0; r < kRounds; ++r) {
          const int slot = (p + r) % kSlots;
          ex.submit(slot, [&, slot] {
            counters[slot] += 1;
            total.fetch_add(1, std::memory_order_relaxed);
          });
          if (r % 97 == 0) ex.submit(slot, [] { throw std::runtime_error("injected task failure"); });
          if (r % 50 == 0) {
            try {
              ex.sync();
            } catch (const std::exception&) {          // an injected failure surfaced in this sync
            }
          }
        }
      });
    }
    for (auto& t : producers) t.join();
    try {
      ex.sync();
    } catch (const std::exception&) {
    }
    long long sum = 0;
    for (long long c : counters) sum += c;
    if (sum != static_cast<long long>(kProducers) * kRounds || total.load() != sum) {
      std::fprintf(stderr, "lost work: %lld / %lld\n", sum, static_cast<long long>(kProducers) * kRounds);
      failures++;
    }
    if (started.load() != kSlots) failures++;
    // errors are reported exactly once and then cleared
    ex.submit(0, [] { throw std::runtime_error("boom"); });
    bool threw = false;
    try {
      ex.sync();
    } catch (const std::exception&) {
      threw = true;
    }
    if (!threw) failures++;
    try {
      ex.sync();
    } catch (const std::exception&) {
      failures++;
    }
    // shutdown with work still queued: everything queued before shutdown still runs, later submits are rejected
    std::atomic<int> late{0};
    for (int i = 0; i < 64; ++i) ex.submit(i % kSlots, [&] { late.fetch_add(1); });
    ex.shutdown();
    if (late.load() != 64) failures++;
  }
  std::printf("host_executor_tsan: %s\n", failures ? "FAILED" : "ok");
  return failures ? 1 : 0;
}
