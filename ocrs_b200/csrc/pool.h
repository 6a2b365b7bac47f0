// Engine pool: pipelining and multi-GPU fan-out INSIDE the library.
//
// The reference runs `detect_words -> find_text_lines -> recognize_text` page by page on the calling
// thread (ocrs/src/lib.rs:193-300, ocrs-cli/src/main.rs:438-446).  One batch alternates between GPU
// phases and host phases (layout analysis, result assembly), so a single in-flight batch leaves the GPU
// idle part of the time.  The pool owns, per device, `in_flight` worker threads with one Engine each
// (own streams, own scratch); callers submit batches of pages and collect results by ticket, so the host
// phases of one batch overlap the kernels of another and a multi-GPU box is fed from one process.
// Workers are pinned to the CPUs of their GPU's NUMA node.
#pragma once
#include <condition_variable>
#include <deque>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "engine.h"

namespace ocrs {

struct PoolPage {  // ImageSource (preprocess.rs:61-124): borrowed until the ticket has been waited for
  const void* pixels = nullptr;
  int dtype = 0, order = 0;  // 0 = u8 / HWC, 1 = f32 / CHW
  int H = 0, W = 0, C = 0;
  bool on_device = false;    // `pixels` lives in device memory (of exactly one of the pool's devices)
};

struct PoolParams {
  EngineParams engine;       // `device` is ignored
  std::vector<int> devices;  // empty = every visible device
  int in_flight = 2;         // worker threads (engines) per device
  bool pin_numa = true;
  int layout_threads = 4;    // host threads per worker for layout analysis (pages of a batch in parallel)
};

class Pool {
 public:
  explicit Pool(const PoolParams& p);
  ~Pool();
  Pool(const Pool&) = delete;
  Pool& operator=(const Pool&) = delete;

  // Enqueues one batch; returns its ticket.  Never blocks on the GPU.
  uint64_t submit(const PoolPage* pages, size_t n_pages);
  // Blocks until the batch is done; rethrows the batch's error.  A ticket can be waited for once.
  std::vector<std::vector<TextLine>> wait(uint64_t ticket);
  bool done(uint64_t ticket);

  int n_devices() const { return (int)devices_.size(); }
  int in_flight() const { return in_flight_; }
  int device_id(int dev_index) const { return devices_.at((size_t)dev_index); }
  // worker engine (for the profiling / statistics hooks); dev_index in [0, n_devices), k in [0, in_flight)
  std::shared_ptr<Engine> engine(int dev_index, int k) const;
  std::string numa_report() const;  // one line per worker: device, NUMA node, CPUs pinned to (diagnostics)

 private:
  struct Job {
    uint64_t ticket = 0;
    std::vector<PoolPage> pages;
    int dev_index = -1;  // -1 = any device
    bool finished = false;
    std::vector<std::vector<TextLine>> result;
    std::exception_ptr error;
  };
  struct Worker {
    int dev_index = 0, slot = 0;
    std::shared_ptr<Engine> engine;
    std::thread thread;
    std::string numa;
  };
  void run_worker(Worker* w);

  std::vector<int> devices_;
  int in_flight_ = 2;
  bool pin_numa_ = true;
  std::vector<std::unique_ptr<Worker>> workers_;
  mutable std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  std::deque<std::shared_ptr<Job>> q_any_;
  std::vector<std::deque<std::shared_ptr<Job>>> q_dev_;
  std::map<uint64_t, std::shared_ptr<Job>> jobs_;
  uint64_t next_ticket_ = 1;
  bool stop_ = false;
};

}  // namespace ocrs
