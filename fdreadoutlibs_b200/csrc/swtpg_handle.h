// Internal layout of swtpg_handle, shared by the translation units behind include/swtpg.h:
//   swtpg_capi.cu    create / destroy / start / stop, the batch entry points and the kernel launch table
//   swtpg_stream.cu  the streaming path (swtpg_submit ... swtpg_poll): per-link rings, dispatcher and completion threads,
//                    the gather kernel that pulls frames out of pinned host memory
// Not installed; nothing outside csrc/ includes it.
#pragma once

#include "../../include/swtpg.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace swtpg_internal {
struct StreamEngine;
struct TpSorter;
}

struct swtpg_handle
{
  swtpg_config cfg{};
  uint32_t unit_bytes = 0, channels = 0, ticks = 0, groups_per_link = 0, n_groups = 0;
  uint32_t tp_capacity = 0;
  bool fast_simple = false, fast_fir = false, fast_rs = false, fast_rs_wib2 = false, fast_fir_any = false;
  bool fir_force_exact = false; // packed FIR / WIB2-AbsRS trackers with the exact-threshold tier on every group (see swtpg_create)
  std::atomic<bool> started{ false };

  cudaStream_t stream = nullptr; // batch path + all kernels (state is carried batch to batch: kernels are ordered)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  cudaStream_t last_stream = nullptr;

  uint32_t* d_state = nullptr;
  uint32_t* d_flags = nullptr;
  uint32_t* d_link_cursor = nullptr; // {claimed, finished}: dynamic link hand-out of the kernels, self-resetting; followed by
                                     // [n_links] slices done per link (wibeth_kernel's sliced hand-out, self-resetting too)
  swtpg_tp* d_tps = nullptr;
  unsigned* d_count = nullptr;
  unsigned* h_count = nullptr;
  uint32_t* d_nunits = nullptr;
  // ragged lengths of the batch entry points: two pinned buffers used alternately, each guarded by the event of the copy that
  // last read it (back-to-back swtpg_process_device calls must not overwrite lengths an earlier async copy still reads)
  uint32_t* h_nunits[2] = { nullptr, nullptr };
  cudaEvent_t ev_nunits[2] = { nullptr, nullptr };
  uint32_t nunits_turn = 0;
  uint8_t* d_frames = nullptr;
  size_t d_frames_bytes = 0;
  int16_t* d_ped = nullptr;
  int16_t* d_wav = nullptr;
  size_t d_dump_elems = 0;
  uint16_t* h_rs_factor = nullptr; // [n_links][channels] or null
  swtpg_internal::TpSorter* sorter = nullptr; // SWTPG_FLAG_SORTED_TPS: device-side ordering of TP lists (swtpg_sort.cu)

  // streaming path (swtpg_stream.cu), created on the first swtpg_submit / swtpg_register_buffer
  std::mutex engine_mu;
  std::atomic<swtpg_internal::StreamEngine*> engine{ nullptr };

  // bounce pipeline of swtpg_process_host for pageable sources: per worker two pinned buffers, a stream and events
  struct Bounce
  {
    uint8_t* buf[2] = { nullptr, nullptr };
    cudaEvent_t free_ev[2] = { nullptr, nullptr };
    cudaEvent_t done = nullptr;
    cudaStream_t stream = nullptr;
  };
  std::vector<Bounce> bounce;

  // counters behind swtpg_get_counters: written by the caller's threads and by the engine's threads
  struct Counters
  {
    std::atomic<uint64_t> units_processed{ 0 }, samples_processed{ 0 }, tps_emitted{ 0 }, tps_dropped_overflow{ 0 }, batches{ 0 },
      submit_busy{ 0 }, h2d_bytes{ 0 }, d2h_bytes{ 0 };
    void reset()
    {
      units_processed = samples_processed = tps_emitted = tps_dropped_overflow = batches = submit_busy = h2d_bytes = d2h_bytes = 0;
    }
  } counters;

  // last failure text: any thread may set it, swtpg_last_error hands out a per-thread copy
  std::mutex err_mu;
  std::string last_error;
  void set_error(const char* msg)
  {
    std::lock_guard<std::mutex> lk(err_mu);
    last_error = msg;
  }
};

namespace swtpg_internal {

void set_create_error(const char* msg);

inline swtpg_status
fail(swtpg_handle* h, swtpg_status s, const char* msg)
{
  if (h)
    h->set_error(msg);
  else
    set_create_error(msg);
  return s;
}

#define SW_CUDA(h, call)                                                                                                          \
  do {                                                                                                                            \
    cudaError_t e_ = (call);                                                                                                      \
    if (e_ != cudaSuccess) {                                                                                                      \
      char buf_[512];                                                                                                             \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);                    \
      swtpg_internal::fail((h), SWTPG_ERR_CUDA, buf_);                                                                            \
      return SWTPG_ERR_CUDA;                                                                                                      \
    }                                                                                                                             \
  } while (0)

// swtpg_capi.cu: launches the handle's fused kernel over one batch laid out link-major in device memory
// ([n_links][units_stride][unit_bytes], d_nunits = per-link valid units or nullptr) on stream `s`. TPs go to d_tps / d_count.
cudaError_t launch_batch_kernel(swtpg_handle* h, const void* d_frames, const uint32_t* d_nunits, uint32_t units_stride, swtpg_tp* d_tps,
                                unsigned* d_count, cudaStream_t s);

// swtpg_sort.cu: device-side ordering of a batch's TP list (SWTPG_FLAG_SORTED_TPS)
TpSorter* sorter_create();
void sorter_destroy(TpSorter* s);
void sorter_stats(const TpSorter* s, double* last_ms, double* total_ms, uint64_t* calls, uint64_t* host_fallbacks);
swtpg_status sort_tps_device(swtpg_handle* h, TpSorter* st, const swtpg_tp* d_tps, size_t n, cudaStream_t s, const swtpg_tp** out,
                             bool* finish_on_host, std::unique_lock<std::mutex>* lock);

// swtpg_stream.cu
void engine_destroy(swtpg_handle* h);          // stops the threads, frees everything (swtpg_destroy)
swtpg_status engine_reset(swtpg_handle* h);    // swtpg_start: idle engine, empty rings
swtpg_status engine_quiesce(swtpg_handle* h);  // swtpg_sync: every dispatched batch has completed

} // namespace swtpg_internal
