// gpu_ai_b200/csrc/hostpool.h -- persistent host worker threads shared by the C ABI (api.cu: packing reference
// States, widening results) and the search tree (tree.cu: parallel leaf selection and statistics update).
// Created on first use, parked on a condition variable between calls.
#pragma once

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace b2p {

class Pool {
 public:
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  // runs f() on `workers` threads in total (the caller is one of them); returns when all have returned
  void run(size_t workers, const std::function<void()> &f) {
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t cap = std::min<size_t>(hw ? hw : 1, 64);
    workers = std::min(workers, cap);
    if (workers <= 1) {
      f();
      return;
    }
    const size_t helpers = workers - 1;
    {
      std::lock_guard<std::mutex> l(mu_);
      while (th_.size() < helpers) th_.emplace_back([this, id = th_.size()] { loop(id); });
      job_ = &f;
      want_ = helpers;
      pending_ = helpers;
      gen_++;
    }
    cv_.notify_all();
    f();
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop(size_t id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void()> *job = nullptr;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || (gen_ != seen && id < want_); });
        if (stop_) return;
        seen = gen_;
        job = job_;
      }
      (*job)();
      std::lock_guard<std::mutex> l(mu_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void()> *job_ = nullptr;
  size_t want_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

}  // namespace b2p
