// Engine pool (see pool.h).
#include "pool.h"

#include <cstdlib>

#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <fstream>
#include <sstream>

namespace ocrs {

namespace {

// CPUs of the NUMA node the GPU hangs off, intersected with the CPUs this process may use.
// Returns false (and leaves the thread alone) when the topology cannot be read.
bool pin_thread_to_gpu_numa(int device, std::string* report) {
  char bus[64] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  std::string b(bus);
  for (auto& c : b) c = (char)std::tolower((unsigned char)c);
  int node = -1;
  {
    std::ifstream f("/sys/bus/pci/devices/" + b + "/numa_node");
    if (!(f >> node) || node < 0) {
      *report = "device " + std::to_string(device) + " (" + b + "): NUMA node unknown, not pinned";
      return false;
    }
  }
  std::string list;
  {
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    if (!std::getline(f, list) || list.empty()) return false;
  }
  cpu_set_t allowed, want;
  CPU_ZERO(&allowed);
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return false;
  std::stringstream ss(list);
  std::string tok;
  int n = 0;
  while (std::getline(ss, tok, ',')) {
    int lo = 0, hi = 0;
    if (sscanf(tok.c_str(), "%d-%d", &lo, &hi) == 2) {
    } else if (sscanf(tok.c_str(), "%d", &lo) == 1) {
      hi = lo;
    } else {
      continue;
    }
    for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &allowed)) {
        CPU_SET(c, &want);
        ++n;
      }
  }
  if (n == 0) {
    *report = "device " + std::to_string(device) + ": NUMA node " + std::to_string(node) + " has no CPU this process may use";
    return false;
  }
  if (pthread_setaffinity_np(pthread_self(), sizeof(want), &want) != 0) return false;
  *report = "device " + std::to_string(device) + " (" + b + "): NUMA node " + std::to_string(node) + ", " + std::to_string(n) +
            " CPUs (" + list + ")";
  return true;
}

}  // namespace

Pool::Pool(const PoolParams& p) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  OCRS_CHECK(e == cudaSuccess && ndev > 0, kNoDevice,
             "no CUDA device available: ocrs_b200 has no CPU fallback (cudaGetDeviceCount: " +
                 std::string(cudaGetErrorString(e)) + ")");
  devices_ = p.devices;
  if (devices_.empty())
    for (int d = 0; d < ndev; ++d) devices_.push_back(d);
  for (size_t i = 0; i < devices_.size(); ++i) {
    OCRS_CHECK(devices_[i] >= 0 && devices_[i] < ndev, kInvalidArg, "device index out of range");
    for (size_t j = 0; j < i; ++j) OCRS_CHECK(devices_[i] != devices_[j], kInvalidArg, "device listed twice");
  }
  OCRS_CHECK(p.in_flight >= 1 && p.in_flight <= 16, kInvalidArg, "in_flight must be in [1, 16]");
  in_flight_ = p.in_flight;
  pin_numa_ = p.pin_numa;
  q_dev_.resize(devices_.size());
  // engines are created here, on the caller's thread: the model buffers are only borrowed for this call
  for (size_t d = 0; d < devices_.size(); ++d)
    for (int k = 0; k < in_flight_; ++k) {
      auto w = std::make_unique<Worker>();
      w->dev_index = (int)d;
      w->slot = k;
      EngineParams ep = p.engine;
      ep.device = devices_[d];
      w->engine = std::make_shared<Engine>(ep);
      w->engine->set_layout_threads(p.layout_threads);
      {
        // workers wait for their batch without holding a core; OCRS_B200_SPIN_SYNC=1 keeps the spinning wait
        static const bool spin = [] { const char* e = std::getenv("OCRS_B200_SPIN_SYNC"); return e != nullptr && e[0] == '1'; }();
        w->engine->set_blocking_sync(!spin);
      }
      workers_.push_back(std::move(w));
    }
  for (auto& w : workers_) w->thread = std::thread([this, wp = w.get()] { run_worker(wp); });
}

Pool::~Pool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  for (auto& w : workers_)
    if (w->thread.joinable()) w->thread.join();
}

std::shared_ptr<Engine> Pool::engine(int dev_index, int k) const {
  OCRS_CHECK(dev_index >= 0 && dev_index < (int)devices_.size() && k >= 0 && k < in_flight_, kInvalidArg,
             "pool engine index out of range");
  return workers_[(size_t)dev_index * in_flight_ + k]->engine;
}

std::string Pool::numa_report() const {
  std::lock_guard<std::mutex> lk(mu_);
  std::string s;
  for (const auto& w : workers_) s += "worker " + std::to_string(w->dev_index) + "." + std::to_string(w->slot) + ": " + w->numa + "\n";
  return s;
}

uint64_t Pool::submit(const PoolPage* pages, size_t n_pages) {
  OCRS_CHECK(pages != nullptr || n_pages == 0, kInvalidArg, "pages is null");
  auto job = std::make_shared<Job>();
  job->pages.assign(pages, pages + n_pages);
  for (const PoolPage& pg : job->pages) {
    OCRS_CHECK(pg.pixels != nullptr, kInvalidArg, "pixels is null");
    if (pg.on_device) {
      // a device-resident page pins the batch to the GPU that holds it
      cudaPointerAttributes at{};
      OCRS_CHECK(cudaPointerGetAttributes(&at, pg.pixels) == cudaSuccess && at.type == cudaMemoryTypeDevice, kInvalidArg,
                 "on_device page does not point to device memory");
      int idx = -1;
      for (size_t d = 0; d < devices_.size(); ++d)
        if (devices_[d] == at.device) idx = (int)d;
      OCRS_CHECK(idx >= 0, kInvalidArg, "on_device page lives on a device this pool does not own");
      OCRS_CHECK(job->dev_index < 0 || job->dev_index == idx, kInvalidArg, "pages of one batch live on different devices");
      job->dev_index = idx;
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu_);
    OCRS_CHECK(!stop_, kInvalidArg, "pool is shutting down");
    job->ticket = next_ticket_++;
    jobs_[job->ticket] = job;
    if (job->dev_index >= 0) q_dev_[(size_t)job->dev_index].push_back(job);
    else q_any_.push_back(job);
  }
  cv_work_.notify_all();
  return job->ticket;
}

bool Pool::done(uint64_t ticket) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = jobs_.find(ticket);
  OCRS_CHECK(it != jobs_.end(), kInvalidArg, "unknown ticket");
  return it->second->finished;
}

std::vector<std::vector<TextLine>> Pool::wait(uint64_t ticket) {
  std::shared_ptr<Job> job;
  {
    std::unique_lock<std::mutex> lk(mu_);
    auto it = jobs_.find(ticket);
    OCRS_CHECK(it != jobs_.end(), kInvalidArg, "unknown ticket (already waited for?)");
    job = it->second;
    cv_done_.wait(lk, [&] { return job->finished; });
    jobs_.erase(ticket);
  }
  if (job->error) std::rethrow_exception(job->error);
  return std::move(job->result);
}

void Pool::run_worker(Worker* w) {
  const int device = devices_[(size_t)w->dev_index];
  cudaSetDevice(device);
  std::string rep = "not pinned";
  if (pin_numa_) pin_thread_to_gpu_numa(device, &rep);
  {
    std::lock_guard<std::mutex> lk(mu_);
    w->numa = rep;
  }
  for (;;) {
    std::shared_ptr<Job> job;
    {
      std::unique_lock<std::mutex> lk(mu_);
      auto& mine = q_dev_[(size_t)w->dev_index];
      cv_work_.wait(lk, [&] { return stop_ || !mine.empty() || !q_any_.empty(); });
      if (!mine.empty()) {
        job = mine.front();
        mine.pop_front();
      } else if (!q_any_.empty()) {
        job = q_any_.front();
        q_any_.pop_front();
      } else {
        return;  // stop_ and nothing left
      }
    }
    try {
      std::vector<Engine::PageSpec> specs;
      specs.reserve(job->pages.size());
      for (const PoolPage& pg : job->pages) specs.push_back(Engine::PageSpec{pg.pixels, pg.dtype, pg.order, pg.H, pg.W, pg.C, pg.on_device});
      std::vector<std::unique_ptr<OcrInput>> inputs = w->engine->prepare_inputs(specs);
      std::vector<const OcrInput*> ptrs;
      for (const auto& in : inputs) ptrs.push_back(in.get());
      job->result = w->engine->ocr_pages(ptrs);
      inputs.clear();
    } catch (...) {
      job->error = std::current_exception();
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      job->finished = true;
    }
    cv_done_.notify_all();
  }
}

}  // namespace ocrs
