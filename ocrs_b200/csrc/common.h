// Shared error handling / small utilities for the ocrs_b200 library.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ocrs {

// Status codes of the C ABI (include/ocrs_b200.h mirrors these).
enum Status : int {
  kOk = 0,
  kInvalidArg = -1,
  kUnsupportedChannelCount = -2,  // ImageSourceError::UnsupportedChannelCount (preprocess.rs:41)
  kInvalidDataLength = -3,        // ImageSourceError::InvalidDataLength (preprocess.rs:44)
  kModelNotLoaded = -4,           // lib.rs:197,211,254,274
  kModelLoad = -5,
  kRunFailed = -6,                // ModelRunError::RunFailed (errors.rs:8)
  kWrongOutput = -7,              // ModelRunError::WrongOutput (errors.rs:11)
  kCuda = -8,
  kNoDevice = -9,
  kInternal = -10,
};

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

#define OCRS_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      throw ::ocrs::Error(::ocrs::kCuda, std::string("CUDA error ") + cudaGetErrorString(_e) + \
                                             " at " + __FILE__ + ":" + std::to_string(__LINE__) + \
                                             " (" #expr ")");                                   \
    }                                                                                           \
  } while (0)

#define OCRS_CHECK(cond, code, msg)                  \
  do {                                               \
    if (!(cond)) throw ::ocrs::Error((code), (msg)); \
  } while (0)

// Number of kernels launched by this library (all engines) since process start.
extern std::atomic<int64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// Optional CUDA-event profiler: brackets named regions on a stream, resolved after a sync.
struct OpProfile {
  double ms = 0, flops = 0, bytes = 0;
  int64_t calls = 0, launches = 0;
};
class Profiler {
 public:
  ~Profiler();
  bool enabled = false;
  // returns a token (index) or -1 when disabled
  int begin(const std::string& name, cudaStream_t st);
  void end(int token, cudaStream_t st, double flops = 0, double bytes = 0);
  void collect();  // call after the stream is synchronized
  void reset() { acc_.clear(); }
  const std::map<std::string, OpProfile>& results() const { return acc_; }

 private:
  struct Pending { std::string name; cudaEvent_t a, b; double flops, bytes; int64_t launches0, launches; };
  std::vector<Pending> pending_;
  std::vector<cudaEvent_t> free_events_;
  std::map<std::string, OpProfile> acc_;
  cudaEvent_t get_event();
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Simple owning device buffer (grow-only).
struct DeviceBuffer {
  void* ptr = nullptr;
  size_t cap = 0;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : ptr(o.ptr), cap(o.cap) { o.ptr = nullptr; o.cap = 0; }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    if (this != &o) { release(); ptr = o.ptr; cap = o.cap; o.ptr = nullptr; o.cap = 0; }
    return *this;
  }
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  // Ensures capacity; contents are NOT preserved on growth.
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    release();
    size_t want = bytes + bytes / 4 + 256;
    OCRS_CUDA_CHECK(cudaMalloc(&ptr, want));
    cap = want;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// Pinned host buffer (grow-only).
struct PinnedBuffer {
  void* ptr = nullptr;
  size_t cap = 0;
  PinnedBuffer() = default;
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  ~PinnedBuffer() { if (ptr) cudaFreeHost(ptr); }
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    OCRS_CUDA_CHECK(cudaMallocHost(&ptr, want));
    cap = want;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

}  // namespace ocrs
