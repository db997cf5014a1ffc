// Shared host/device helpers for the ntk_b200 library (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ntk_b200.h"

namespace ntk {

// ---- thread-local error string ------------------------------------------------
inline std::string& last_error() {
  thread_local std::string e;
  return e;
}

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

#define NTK_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::ntk::fail(NTK_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,           \
                         cudaGetErrorString(_e));                                          \
  } while (0)

#define NTK_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != NTK_OK) return _s; \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (device, function): a host thread that
// moves to another GPU must set it again there.  One process-wide table keyed by (device, function).
inline int ensure_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;  // largest size configured so far
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(NTK_ECUDA, "cudaGetDevice -> %s", cudaGetErrorString(e));
  // never below the 48 KB every launch may use without the attribute: a small first request must not make a later
  // unconfigured launch of the same kernel (<= 48 KB) fail with "invalid argument"
  if (bytes < (size_t)48 * 1024) bytes = (size_t)48 * 1024;
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find({dev, func});
  if (it != done.end() && it->second >= bytes) return NTK_OK;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess)
    return fail(NTK_ECUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %zu) -> %s", bytes, cudaGetErrorString(e));
  done[{dev, func}] = bytes;
  return NTK_OK;
}

// ---- padding arithmetic (lax.padtype_to_pads) ----------
struct AxisGeom {
  int out;  // output size
  int lo;   // low-side padding
};

inline AxisGeom axis_geom(int n, int k, int s, int padding) {
  AxisGeom g;
  if (padding == NTK_PAD_VALID) {
    g.out = n >= k ? (n - k) / s + 1 : 0;
    g.lo = 0;
  } else {  // SAME and CIRCULAR share the SAME geometry (linear.py:3158-3165)
    g.out = (n + s - 1) / s;
    int tot = (g.out - 1) * s + k - n;
    if (tot < 0) tot = 0;
    g.lo = tot / 2;
  }
  return g;
}

// ---- stream-ordered bump/free-list arena over one device workspace -------------
// All kernels of a context run on one stream, so a block can be handed out
// again as soon as the host frees it.  In `dry` mode nothing is backed by memory
// and `peak` reports the workspace a real run with identical calls would need.
class Arena {
 public:
  void reset(char* base, size_t cap, bool dry) {
    base_ = dry ? reinterpret_cast<char*>(uintptr_t(1) << 40) : base;
    cap_ = dry ? (size_t(1) << 60) : cap;
    dry_ = dry;
    peak_ = 0;
    free_.clear();
    used_.clear();
    free_[0] = cap_;
  }
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        size_t off = it->first, sz = it->second;
        free_.erase(it);
        if (sz > bytes) free_[off + bytes] = sz - bytes;
        used_[off] = bytes;
        if (off + bytes > peak_) peak_ = off + bytes;
        return base_ + off;
      }
    }
    return nullptr;
  }
  void release(void* p) {
    if (!p) return;
    size_t off = static_cast<char*>(p) - base_;
    auto u = used_.find(off);
    if (u == used_.end()) return;
    size_t sz = u->second;
    used_.erase(u);
    auto nx = free_.lower_bound(off);
    if (nx != free_.end() && off + sz == nx->first) {
      sz += nx->second;
      nx = free_.erase(nx);
    }
    if (nx != free_.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == off) {
        pv->second += sz;
        return;
      }
    }
    free_[off] = sz;
  }
  size_t largest_free() const {
    size_t m = 0;
    for (const auto& f : free_) m = f.second > m ? f.second : m;
    return m;
  }
  size_t peak() const { return peak_; }
  bool dry() const { return dry_; }

 private:
  char* base_ = nullptr;
  size_t cap_ = 0, peak_ = 0;
  bool dry_ = false;
  std::map<size_t, size_t> free_, used_;
};

constexpr int kNumSMs = 148;  // B200

// Optional per-stage device timing (ntk_context_set_profiling).
struct StageProfile {
  static constexpr int kMaxStages = 16;
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[kMaxStages];
  std::vector<long long> pending_pairs[kMaxStages];
  double total_ms[kMaxStages] = {0};
  long long launches[kMaxStages] = {0};
  long long pairs[kMaxStages] = {0};
};

}  // namespace ntk
