/*
 * swtpg.h — C ABI of the B200-native software trigger-primitive generator (libswtpg_b200.so).
 *
 * This is the drop-in boundary for ONE path of DUNE-DAQ/fdreadoutlibs: what happens inside
 * WIBEthFrameProcessor::find_hits / WIB2FrameProcessor::find_hits, i.e.
 *   14-bit unpack -> frugal-streaming pedestal -> [running sum | FIR] -> threshold hit finding -> TriggerPrimitive fields.
 * Every entry point names the reference interface it replaces (paths relative to the reference repository root).
 * Plain C: POD structs, pointers and sizes, integer status codes; no exceptions cross this boundary and no
 * torch / CUDA types appear in any signature (streams and device pointers travel as void*).
 *
 * There is no CPU fallback: every compute entry point fails with SWTPG_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SWTPG_H_
#define SWTPG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWTPG_ABI_VERSION 2u

/* Frame geometry restated from the reference (sizes pinned by static_asserts there). */
#define SWTPG_WIBETH_FRAME_BYTES 7200u /* include/fdreadoutlibs/DUNEWIBEthTypeAdapter.hpp:20-22,98 */
#define SWTPG_WIBETH_CHANNELS 64u
#define SWTPG_WIBETH_TICKS 64u        /* wibeth/tpg/TPGConstants_wibeth.hpp:22 (FRAMES_PER_MSG) */
#define SWTPG_WIBETH_TS_PER_FRAME 2048u /* DUNEWIBEthTypeAdapter.hpp:93 */
#define SWTPG_WIB2_FRAME_BYTES 472u
#define SWTPG_WIB2_SUPERCHUNK_BYTES 5664u /* include/fdreadoutlibs/DUNEWIBSuperChunkTypeAdapter.hpp:18-22,100 */
#define SWTPG_WIB2_CHANNELS 256u
#define SWTPG_WIB2_TICKS 12u          /* wib2/tpg/TPGConstants_wib2.hpp:30 (FRAMES_PER_MSG) */
#define SWTPG_TS_PER_TICK 32u         /* DUNEWIBEthTypeAdapter.hpp:95, DUNEWIBSuperChunkTypeAdapter.hpp:97 */
#define SWTPG_MAX_TAPS 8u

typedef enum swtpg_status
{
  SWTPG_OK = 0,
  SWTPG_ERR_INVALID_ARG = 1,
  SWTPG_ERR_CUDA = 2,       /* no usable device / CUDA runtime error; see swtpg_last_error() */
  SWTPG_ERR_BUSY = 3,       /* back-pressure: staging ring full (reference: FailedToSendTP-style drop, never blocks) */
  SWTPG_ERR_OVERFLOW = 4,   /* more TPs than the caller's / the device buffer's capacity; count is still reported */
  SWTPG_ERR_STATE = 5,      /* call sequence violated (e.g. submit before start) */
  SWTPG_ERR_UNSUPPORTED = 6 /* reference: TPGAlgorithmInexistent (include/fdreadoutlibs/FDReadoutIssues.hpp:27-31) */
} swtpg_status;

typedef enum swtpg_format
{
  SWTPG_FORMAT_WIBETH = 0, /* unit = one 7200-B WIBEthFrame: 64 channels x 64 ticks */
  SWTPG_FORMAT_WIB2 = 1    /* unit = one 5664-B superchunk of 12 WIB2 frames: 256 channels x 12 ticks */
} swtpg_format;

/* tpg_algorithm strings of the reference (src/wibeth/WIBEthFrameProcessor.cpp:180-197) plus the FIR+IQR finder that
 * the reference ships for WIB2/ProtoWIB (include/fdreadoutlibs/wib2/tpg/ProcessAVX2FIR.hpp). */
typedef enum swtpg_algorithm
{
  SWTPG_ALGO_SIMPLE_THRESHOLD = 0, /* "SimpleThreshold": wibeth/tpg/ProcessAVX2.hpp, wib2/tpg/ProcessAVX2.hpp */
  SWTPG_ALGO_ABS_RS = 1,           /* "AbsRS":      wibeth/tpg/ProcessAbsRSAVX2.hpp, wib2/tpg/ProcessRSAVX2.hpp */
  SWTPG_ALGO_STANDARD_RS = 2,      /* "StandardRS": wibeth/tpg/ProcessStandardRSAVX2.hpp */
  SWTPG_ALGO_FIR_IQR = 3           /* FIR matched filter + IQR threshold: wib2/tpg/ProcessAVX2FIR.hpp */
} swtpg_algorithm;

/* One trigger primitive as emitted by the device: the fields process_swtpg_hits derives per hit
 * (src/wibeth/WIBEthFrameProcessor.cpp:523-545, src/wib2/WIB2FrameProcessor.cpp:431-455) BEFORE the
 * register->offline-channel LUT, channel mask and tp_timeout filters, which stay on the host shim.
 * `channel` is the frame channel (WIBEth 0..63, WIB2 0..255); see DESIGN.md "H2". 32 bytes. */
typedef struct swtpg_tp
{
  uint64_t time_start;          /* 62.5 MHz ticks: ts(frame of hit end) + 32*(t_end - tover) */
  uint64_t time_peak;
  uint32_t time_over_threshold; /* 32 * samples over threshold */
  uint32_t adc_integral;
  uint16_t adc_peak;
  uint16_t channel;
  uint32_t link;                /* index of the link inside this handle (0..n_links-1) */
} swtpg_tp;

/* Mirrors the fields of readoutlibs' RawDataProcessorConf that the hot path consumes
 * (src/wibeth/WIBEthFrameProcessor.cpp:175-230) and the handler constants
 * (include/fdreadoutlibs/wibeth/WIBEthFrameProcessor.hpp:69, wib2/WIB2FrameProcessor.hpp:68-69). */
typedef struct swtpg_config
{
  uint32_t struct_size;      /* = sizeof(swtpg_config); lets the ABI grow */
  int32_t device;            /* CUDA device ordinal */
  int32_t format;            /* swtpg_format */
  int32_t algorithm;         /* swtpg_algorithm */
  uint32_t n_links;          /* independent links owned by this handle (one reference FrameProcessor each) */
  uint32_t max_units;        /* superchunk length: units (frames / WIB2 superchunks) per link per batch, < 2^18 */
  uint32_t tp_capacity;      /* device TP buffer, records per batch; 0 = sized for the worst case */
  uint32_t n_slots;          /* staging-ring depth of the streaming path (>= 2); 0 = 4 */
  uint16_t threshold;        /* tpg_threshold (ADC; sigma units for FIR_IQR) */
  int16_t frugal_acc_limit;  /* tpg_frugal_streaming_accumulator_limit (WIB2 and FIR paths hard-wire 10) */
  uint16_t rs_memory_factor; /* already x10, as conf() scales it (WIBEthFrameProcessor.cpp:202) */
  uint16_t rs_scale_factor;  /* already 10/x (WIBEthFrameProcessor.cpp:206) */
  int16_t fir_taps[SWTPG_MAX_TAPS]; /* FIR_IQR only; all-zero = firwin_int(7,0.1,64)+{0} = {1,6,15,20,15,6,1,0} */
  uint8_t tap_exponent;      /* m_tpg_tap_exponent = 6 */
  uint8_t reserved0[3];
  uint32_t wib2_adc_offset;  /* byte offset of adc_words inside a WIB2 frame; 0 = 20 */
  uint32_t flags;            /* SWTPG_FLAG_* */
  uint32_t dispatch_timeout_us; /* streaming path: units that have waited this long are dispatched even if no link has a full
                                   superchunk yet (a stalled or dead link never holds the others back); 0 = 5000 */
} swtpg_config;

#define SWTPG_FLAG_NONE 0u
/* The TP list of every batch (swtpg_process_host, swtpg_fetch_tps, swtpg_poll*) comes back ordered by (time_start, link,
 * channel) — the order TriggerPrimitiveTypeAdapter::operator< imposes downstream
 * (include/fdreadoutlibs/TriggerPrimitiveTypeAdapter.hpp:26-29) — ordered ON THE DEVICE before it crosses the host link (an LSD
 * radix sort of packed keys in HBM), identical to what swtpg_sort_tps makes of the unordered list. Lists of one handle can
 * then be merged across GPUs with swtpg_merge_sorted without a host sort. */
#define SWTPG_FLAG_SORTED_TPS 1u

/* Per-channel carried state, by frame channel: ChanState of wibeth/tpg/ProcessingInfo.hpp:20-66 and
 * wib2/tpg/ProcessingInfo.hpp:20-68. Used for parity dumps only. */
typedef struct swtpg_channel_state
{
  int16_t pedestal, accum;
  int16_t quantile25, quantile75, accum25, accum75;
  int16_t rs, pedestal_rs, accum_rs;
  uint16_t rs_memory_factor;
  uint16_t prev_was_over, hit_charge, hit_tover, hit_peak_adc, hit_peak_time;
  uint16_t initialized; /* 0 until the first unit seeded the pedestal (setState) */
  int16_t prev_samp[SWTPG_MAX_TAPS];
} swtpg_channel_state;

/* Counters behind the opmon fields of get_info (src/wibeth/WIBEthFrameProcessor.cpp:237-292). */
typedef struct swtpg_counters
{
  uint64_t units_processed;  /* frames / superchunks */
  uint64_t samples_processed;
  uint64_t tps_emitted;
  uint64_t tps_dropped_overflow;
  uint64_t batches;
  uint64_t submit_busy;      /* swtpg_submit calls refused with SWTPG_ERR_BUSY */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t units_zero_copy;  /* streaming path: units read where they lay in a registered latency buffer ... */
  uint64_t units_staged;     /* ... and units copied into the pinned staging ring by swtpg_submit */
} swtpg_counters;

typedef struct swtpg_handle swtpg_handle;

uint32_t swtpg_abi_version(void);
const char* swtpg_status_string(swtpg_status s);
/* Thread-local text of the last failure on this handle (NULL handle: last create failure). */
const char* swtpg_last_error(const swtpg_handle* h);
/* 1 if a CUDA device of compute capability 10.x is visible, else 0. Never throws, never initialises a context. */
int swtpg_device_available(void);

/* Replaces WIBEthFrameHandler::initialize / WIB2FrameHandler::initialize (buffers, taps, ProcessingInfo):
 * src/wibeth/WIBEthFrameProcessor.cpp:74-91, src/wib2/WIB2FrameProcessor.cpp:90-120. */
swtpg_status swtpg_create(const swtpg_config* cfg, swtpg_handle** out);
void swtpg_destroy(swtpg_handle* h);

/* Replaces the TPG part of FrameProcessor::start / ::stop: fresh zeroed ChanState, first_hit re-armed
 * (src/wibeth/WIBEthFrameProcessor.cpp:111-154, 67-72). */
swtpg_status swtpg_start(swtpg_handle* h);
swtpg_status swtpg_stop(swtpg_handle* h);

/* Per-position RS memory factor of AbsRS / StandardRS, by frame channel: [n_links][channels]
 * (src/wibeth/WIBEthFrameProcessor.cpp:437-456). NULL = cfg.rs_memory_factor everywhere. Call before the first unit. */
swtpg_status swtpg_set_rs_memory_factor(swtpg_handle* h, const uint16_t* by_link_channel);
/* The same for ONE link: by_channel[channels]. This is what a frame processor calls from find_hits on its first frame
 * (WIBEthFrameProcessor.cpp:437-456 runs per link): one small asynchronous copy ordered on the compute stream, thread-safe
 * against the other links' calls and against batches in flight. */
swtpg_status swtpg_set_link_rs_memory_factor(swtpg_handle* h, uint32_t link, const uint16_t* by_channel);

/*
 * Batch entry points. `frames` is link-major: unit u of link l starts at ((l * units_stride) + u) * unit_bytes.
 * n_units[l] <= units_stride <= cfg.max_units valid units for link l (ragged batches allowed; NULL = all links
 * have units_stride units). State is carried from the previous batch per link, exactly as ProcessingInfo carries
 * it from frame to frame. TPs come back unordered; sort with swtpg_sort_tps for (time_start, link, channel) order.
 *
 * Together these replace the body of find_hits: expand_wibeth_adcs + setState on the first frame +
 * m_assigned_tpg_algorithm_function + process_swtpg_hits' field arithmetic
 * (src/wibeth/WIBEthFrameProcessor.cpp:410-476,478-549; src/wib2/WIB2FrameProcessor.cpp:345-396,398-458).
 */
/* Host buffers (pageable or pinned): H2D copy, kernel, D2H of the TP list, all inside the call. Pinned or registered sources
 * are read by the copy engine directly; pageable sources of 16 MB and more go through a pipeline of pinned bounce buffers
 * filled by a few worker threads (3.6x the rate of a plain cudaMemcpy from pageable memory on the bench box). */
swtpg_status swtpg_process_host(swtpg_handle* h, const void* frames, const uint32_t* n_units, uint32_t units_stride,
                                swtpg_tp* out, size_t cap, size_t* n_out);
/* Frames already resident in HBM. `stream` is a cudaStream_t; NULL = the handle's own (non-blocking) stream — to run on
 * the legacy default stream pass cudaStreamLegacy, i.e. (void*)1. Asynchronous: the caller orders the producer of
 * d_frames before this call on that stream. TPs stay on the device until swtpg_fetch_tps. */
swtpg_status swtpg_process_device(swtpg_handle* h, const void* d_frames, const uint32_t* n_units, uint32_t units_stride,
                                  void* stream);
/* Waits for the last swtpg_process_device, copies its TPs out. *n_out = TPs found (may exceed cap: OVERFLOW). */
swtpg_status swtpg_fetch_tps(swtpg_handle* h, swtpg_tp* out, size_t cap, size_t* n_out);
/* Device time of the last batch kernel in milliseconds (CUDA events on the launching stream); < 0 if none. */
double swtpg_last_kernel_ms(swtpg_handle* h);

/* Parity dump of the intermediate waveforms the reference can print (docs/README.md:31-36 "save-adc"):
 * same as swtpg_process_host, plus per sample the pedestal AFTER its frugal update and the waveform the
 * threshold is applied to (pedestal-subtracted ADC, running sum, or FIR output), both int16,
 * laid out [link][unit][tick][channel] over units_stride units per link. Either may be NULL. */
swtpg_status swtpg_process_host_debug(swtpg_handle* h, const void* frames, const uint32_t* n_units, uint32_t units_stride,
                                      swtpg_tp* out, size_t cap, size_t* n_out, int16_t* pedestal_out, int16_t* waveform_out);

/*
 * Streaming entry points: what a FrameProcessor's post-processing thread calls once per payload.
 *
 * Every link owns a ring of n_slots * max_units pending units. swtpg_submit appends one unit of `link` to that link's ring —
 * the payload's address if it lies in a latency buffer registered with swtpg_register_buffer (zero-copy), else a copy of it
 * in the link's pinned staging ring (the reference's constframeptr is only borrowed for the duration of find_hits) — and
 * returns; it takes no lock and never blocks: SWTPG_ERR_BUSY when that link's ring is full (the reference's failed try_send).
 * One producer thread per link at a time, any number of links concurrently.
 *
 * A dispatcher thread inside the library turns pending units into batches: as soon as ANY link has a full superchunk
 * (cfg.max_units units), or pending units have waited cfg.dispatch_timeout_us, or on swtpg_flush, every link contributes what
 * it has delivered so far (0 .. max_units units: batches are ragged, per-link state is carried from batch to batch exactly as
 * ProcessingInfo carries it from frame to frame). So links advance independently; a silent link does not hold back the TPs
 * of the others. A batch is: ONE gather kernel that pulls the units out of pinned host memory through a per-unit pointer
 * table (no per-link copy calls) -> the fused TPG kernel -> the TP list back to pinned host memory. A completion thread frees
 * the ring space of a batch as soon as its gather has finished (which also ends the borrow of zero-copy units) and queues
 * the batch's TPs; swtpg_poll hands them out, oldest batch first. It replaces the per-hit m_tp_sink->try_send loop's source
 * (src/wibeth/WIBEthFrameProcessor.cpp:555). At most n_slots batches exist at a time: un-polled TPs back-pressure the ring.
 */
swtpg_status swtpg_submit(swtpg_handle* h, uint32_t link, const void* unit, size_t bytes);
/* Like swtpg_submit, but waits (sleeping, not spinning) up to timeout_us for room in the link's ring; SWTPG_ERR_BUSY on
 * time-out. For file replay and emulators, where stalling the source is right and dropping frames is not. */
swtpg_status swtpg_submit_wait(swtpg_handle* h, uint32_t link, const void* unit, size_t bytes, uint64_t timeout_us);
/*
 * Zero-copy ingest. The constframeptr a post-processing task receives points INTO the link's latency buffer
 * (readoutlibs IterableQueueModel / FixedRateQueueModel: one contiguous array of payloads; SURVEY.md 8b "Ownership").
 * Registering that array here (cudaHostRegister, mapped) makes swtpg_submit record the payload's address instead of copying
 * it: the batch's gather kernel reads the unit where it lies, over the host link. Units outside every registered range (or
 * not 16-byte aligned) are still copied, so both kinds may be mixed. Contract: a submitted unit must stay unmodified until
 * its batch's gather has completed — at most n_slots superchunks after the submit, milliseconds against a latency buffer's
 * seconds of retention; swtpg_flush + swtpg_sync (or stop) end every such borrow. Any number of ranges (one per link is
 * typical). swtpg_unregister_buffer flushes and waits for the handle to drain first.
 */
swtpg_status swtpg_register_buffer(swtpg_handle* h, void* base, size_t bytes);
swtpg_status swtpg_unregister_buffer(swtpg_handle* h, void* base);
/* Dispatches everything submitted so far (ragged batch), returns when it has been handed to the device. Safe against
 * concurrent swtpg_submit calls (units submitted meanwhile may or may not be included). SWTPG_ERR_BUSY if all n_slots
 * batches hold TPs nobody has polled yet: poll, then flush again. */
swtpg_status swtpg_flush(swtpg_handle* h);
/* TPs of completed batches, oldest first, without blocking. *n_out = records written (<= cap); call again for more. */
swtpg_status swtpg_poll(swtpg_handle* h, swtpg_tp* out, size_t cap, size_t* n_out);
/* The same, but sleeps up to timeout_us until at least one completed batch is available. */
swtpg_status swtpg_poll_wait(swtpg_handle* h, swtpg_tp* out, size_t cap, size_t* n_out, uint64_t timeout_us);
/* Blocks until every dispatched batch has completed (used by stop and by tests). */
swtpg_status swtpg_sync(swtpg_handle* h);
/* Streaming path bookkeeping: units submitted but not yet dispatched, batches dispatched but not yet completed, completed
 * batches waiting for swtpg_poll. Any pointer may be NULL. */
swtpg_status swtpg_stream_status(swtpg_handle* h, uint64_t* units_pending, uint32_t* batches_in_flight, uint32_t* batches_ready);
/* Device time (CUDA events) the completed batches of the streaming path spent in the gather kernel — the host-link transfer —
 * and in the fused TPG kernel since swtpg_start, and how many batches that covers. Monitoring / bench aid. */
swtpg_status swtpg_stream_timing(swtpg_handle* h, double* gather_ms, double* kernel_ms, uint64_t* batches);

/* Carried state of one link, by frame channel (ChanState parity). out has SWTPG_*_CHANNELS entries. */
swtpg_status swtpg_dump_state(swtpg_handle* h, uint32_t link, swtpg_channel_state* out);
swtpg_status swtpg_get_counters(swtpg_handle* h, swtpg_counters* out);
/* SWTPG_FLAG_SORTED_TPS bookkeeping: device time (CUDA events, includes the 16-byte read-back of the key range) of the last
 * ordered list and of all lists since swtpg_create, how many lists were ordered on the device, and how many of them the host
 * had to finish (equal keys, or keys wider than 64 bits: links whose timestamps are unrelated). Any pointer may be NULL. */
swtpg_status swtpg_sort_stats(swtpg_handle* h, double* last_ms, double* total_ms, uint64_t* lists, uint64_t* finished_on_host);

/* Page-locked host memory for frame buffers handed to swtpg_process_host (or used as a latency buffer without a later
 * swtpg_register_buffer). write_combined != 0 asks for write-combined pages: the CPU only ever WRITES frames there (reads are
 * very slow) and the device's reads do not snoop the CPU caches, which matters when several GPUs of one host ingest at once.
 * Returns NULL on failure. */
void* swtpg_alloc_pinned(size_t bytes, int write_combined);
void swtpg_free_pinned(void* p);

/* Host-side ordering of a TP list by (time_start, link, channel): the order TriggerPrimitiveTypeAdapter::operator<
 * imposes downstream (include/fdreadoutlibs/TriggerPrimitiveTypeAdapter.hpp:26-29). In place. */
void swtpg_sort_tps(swtpg_tp* tps, size_t n);
/* k-way merge of already sorted per-GPU TP lists into `out` (capacity sum of n[i]): the host-side time-ordered
 * merge that follows link sharding across GPUs. */
void swtpg_merge_sorted(const swtpg_tp* const* lists, const size_t* n, size_t k, swtpg_tp* out);

/* firwin_int of src/wib2/tpg/DesignFIR.cpp:57-68 (host, double precision): taps[n]. Returns n. */
int swtpg_firwin_int(int n, double cutoff, int multiplier, int16_t* taps);

#ifdef __cplusplus
}
#endif
#endif /* SWTPG_H_ */
