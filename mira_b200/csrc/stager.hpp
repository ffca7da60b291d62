// stager.hpp — staging of PAGEABLE host scalars through a small ring of page-locked slots.
//
// A Rust `Vec<C::Scalar>` handed to CommitmentKey::commit (/root/reference/src/commitment.rs:78) is pageable:
// cudaMemcpyAsync from it is synchronous and driver-staged (~6 GB/s measured, 90 ms end to end at 2^24 scalars against
// 42 ms from page-locked memory).  The stager copies 8 MiB chunks into page-locked slots with a few worker threads
// (host memcpy runs at several times the driver's staging rate) and issues the H2D of each chunk from its slot, so
// the copies stay asynchronous and keep overlapping with the accumulation of the previous slice.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace mira_host {

class CopyPool {       // fixed pool of workers that split one memcpy between them
 public:
  explicit CopyPool(int n_workers) {
    for (int i = 0; i < n_workers; i++) workers_.emplace_back([this, i, n_workers] { run(i, n_workers); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      generation_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void copy(void* dst, const void* src, size_t bytes) {
    std::unique_lock<std::mutex> lk(mu_);
    dst_ = static_cast<char*>(dst);
    src_ = static_cast<const char*>(src);
    bytes_ = bytes;
    pending_ = (int)workers_.size();
    generation_++;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run(int index, int n) {
    uint64_t seen = 0;
    for (;;) {
      char* dst;
      const char* src;
      size_t bytes;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        dst = dst_; src = src_; bytes = bytes_;
      }
      size_t per = ((bytes + n - 1) / n + 63) & ~(size_t)63, lo = per * index;
      if (lo < bytes) std::memcpy(dst + lo, src + lo, lo + per <= bytes ? per : bytes - lo);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0;
  int pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

struct Stager {
  static constexpr int SLOTS = 4;
  static constexpr size_t SLOT_BYTES = (size_t)8 << 20;
  void* slot[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t slot_free[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  CopyPool* pool = nullptr;
  size_t next = 0;

  cudaError_t init() {
    if (pool) return cudaSuccess;
    for (int i = 0; i < SLOTS; i++) {
      cudaError_t e = slot[i] ? cudaSuccess : cudaMallocHost(&slot[i], SLOT_BYTES);
      if (e == cudaSuccess && !slot_free[i]) e = cudaEventCreateWithFlags(&slot_free[i], cudaEventDisableTiming);
      if (e != cudaSuccess) {          // leave nothing half-made behind: the next commit starts from scratch
        release();
        return e;
      }
    }
    unsigned hw = std::thread::hardware_concurrency();
    pool = new CopyPool(hw >= 16 ? 6 : (hw >= 8 ? 4 : 2));
    return cudaSuccess;
  }
  // host (pageable) -> device, asynchronous on `copy_stream` from the device's point of view
  cudaError_t copy(void* dev, const void* host, size_t bytes, cudaStream_t copy_stream) {
    const char* h = static_cast<const char*>(host);
    char* d = static_cast<char*>(dev);
    for (size_t off = 0; off < bytes; off += SLOT_BYTES, next++) {
      int s = (int)(next % SLOTS);
      size_t len = bytes - off < SLOT_BYTES ? bytes - off : SLOT_BYTES;
      cudaError_t e = cudaEventSynchronize(slot_free[s]);      // the slot's previous H2D has drained
      if (e != cudaSuccess) return e;
      pool->copy(slot[s], h + off, len);
      e = cudaMemcpyAsync(d + off, slot[s], len, cudaMemcpyHostToDevice, copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(slot_free[s], copy_stream);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  void release() {
    delete pool;
    pool = nullptr;
    for (int i = 0; i < SLOTS; i++) {
      if (slot[i]) cudaFreeHost(slot[i]);
      if (slot_free[i]) cudaEventDestroy(slot_free[i]);
      slot[i] = nullptr;
      slot_free[i] = nullptr;
    }
  }
};

}  // namespace mira_host
