// wibeth_tpg_algorithms_emulator — file-replay front-end of the B200 SWTPG, with the options of the reference's emulator of the
// same name (docs/README.md:20-48; the application itself is absent from the reference snapshot, CMakeLists.txt:77-79).
//
// Reads a raw binary file of concatenated 7200-byte WIBEth frames (docs/README.md:74-82: the wire / on-disk format), pushes
// every frame through the drop-in frame processor exactly as a readout application would — pre-process tasks
// (sequence_check, timestamp_check), then the post-process task find_hits — and collects the TriggerPrimitives the
// processor sends to its "tp_out" sink. One WIBEthFrameProcessor per link on one TpgEngine (GPU); the file is one link's
// frame stream, replicated on --links links if asked (each with its own stream id, hence its own offline channels).
//
//   -f,--frame-file-path TEXT    input frame file
//   -a,--algorithm TEXT          SimpleThreshold | AbsRS | StandardRS
//   -i,--implementation TEXT     CUDA (default). AVX is accepted as an alias: the CUDA path reproduces the AVX2 processor bit
//                                for bit. NAIVE is refused: there is no CPU implementation in this library.
//   -d,--duration-test INT       keep replaying the file for this many seconds (0 = one pass, the default; when looping,
//                                emulator mode keeps the timestamps running, as the reference's emulated links do)
//   -n,--num-frames-to-read INT  frames to read from the file (default: all)
//   -t,--tpg-threshold INT       threshold
//   --save-adc-data              write the ADC values after the 14 -> 16 bit expansion to a text file (one row per tick)
//   --save-trigprim              write the TriggerPrimitives to a text file: channel,time_start,time_over_threshold,time_peak,
//                                adc_integral,adc_peak,type (docs/README.md:84-88)
//   --links INT, --superchunk INT, --device INT, --out-prefix TEXT   (ours)
#include "../fdreadoutlibs_b200/host/swtpg_host.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using namespace swtpg::host;

namespace {

struct Options
{
  std::string file, algorithm = "SimpleThreshold", implementation = "CUDA", prefix = "wibeth_tpg";
  int duration = 0, device = 0;
  long n_frames = -1;
  int threshold = 100;
  unsigned links = 1, superchunk = 64;
  bool save_adc = false, save_tp = false;
};

void
usage()
{
  std::puts("Test TPG algorithms (B200)\n"
            "Usage: wibeth_tpg_algorithms_emulator [OPTIONS]\n\n"
            "Options:\n"
            "  -h,--help                   Print this help message and exit\n"
            "  -f,--frame-file-path TEXT   Path to the input frame file\n"
            "  -a,--algorithm TEXT         TPG Algorithm (SimpleThreshold / AbsRS / StandardRS)\n"
            "  -i,--implementation TEXT    TPG implementation (CUDA; AVX is an alias, NAIVE does not exist here)\n"
            "  -d,--duration-test INT      Duration (in seconds) to run the test (0: one pass over the file)\n"
            "  -n,--num-frames-to-read INT Number of frames to read. Default: select all frames.\n"
            "  -t,--tpg-threshold INT      Value of the TPG threshold\n"
            "  --save-adc-data             Save ADC data\n"
            "  --save-trigprim             Save trigger primitive data\n"
            "  --links INT                 Replay the file on this many links at once (default 1)\n"
            "  --superchunk INT            Frames per link and GPU batch (default 64)\n"
            "  --device INT                CUDA device ordinal (default 0)\n"
            "  --out-prefix TEXT           Prefix of the output text files (default wibeth_tpg)");
}

bool
parse(int argc, char** argv, Options& o)
{
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    std::replace(a.begin(), a.end(), '_', '-'); // the reference's examples write --frame_file_path as well
    auto val = [&](const char* what) -> std::string {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "%s needs a value\n", what);
        std::exit(2);
      }
      return argv[++i];
    };
    if (a == "-h" || a == "--help") {
      usage();
      std::exit(0);
    } else if (a == "-f" || a == "--frame-file-path")
      o.file = val("-f");
    else if (a == "-a" || a == "--algorithm")
      o.algorithm = val("-a");
    else if (a == "-i" || a == "--implementation")
      o.implementation = val("-i");
    else if (a == "-d" || a == "--duration-test")
      o.duration = std::atoi(val("-d").c_str());
    else if (a == "-n" || a == "--num-frames-to-read")
      o.n_frames = std::atol(val("-n").c_str());
    else if (a == "-t" || a == "--tpg-threshold")
      o.threshold = std::atoi(val("-t").c_str());
    else if (a == "--save-adc-data")
      o.save_adc = true;
    else if (a == "--save-trigprim")
      o.save_tp = true;
    else if (a == "--links")
      o.links = unsigned(std::atoi(val("--links").c_str()));
    else if (a == "--superchunk")
      o.superchunk = unsigned(std::atoi(val("--superchunk").c_str()));
    else if (a == "--device")
      o.device = std::atoi(val("--device").c_str());
    else if (a == "--out-prefix")
      o.prefix = val("--out-prefix");
    else {
      std::fprintf(stderr, "unknown option %s\n", argv[i]);
      return false;
    }
  }
  return true;
}

// ADC value of channel c at tick t of a frame: 14 bits at bit 14 c of the tick's 112-byte row (fddetdataformats get_adc)
inline uint16_t
get_adc(const DUNEWIBEthTypeAdapter& f, unsigned c, unsigned t)
{
  const unsigned char* row = reinterpret_cast<const unsigned char*>(f.data) + 32 + 112 * t;
  const unsigned bit = 14 * c;
  uint32_t w = 0;
  std::memcpy(&w, row + bit / 8, bit / 8 + 4 <= 112 ? 4 : 112 - bit / 8);
  return uint16_t((w >> (bit % 8)) & 0x3FFFu);
}

} // namespace

int
main(int argc, char** argv)
{
  Options o;
  if (!parse(argc, argv, o))
    return 2;
  if (o.file.empty()) {
    usage();
    return 2;
  }
  if (o.implementation == "NAIVE") {
    std::fprintf(stderr, "implementation NAIVE: this library has no CPU implementation (CUDA only; AVX is accepted as an alias)\n");
    return 2;
  }
  if (o.implementation != "CUDA" && o.implementation != "AVX") {
    std::fprintf(stderr, "unknown implementation %s\n", o.implementation.c_str());
    return 2;
  }
  if (o.links == 0 || o.superchunk == 0)
    return 2;

  // ---- the frame file: concatenated 7200-byte frames ----
  std::ifstream in(o.file, std::ios::binary | std::ios::ate);
  if (!in) {
    std::fprintf(stderr, "cannot open %s\n", o.file.c_str());
    return 1;
  }
  const size_t bytes = size_t(in.tellg());
  size_t n_frames = bytes / sizeof(DUNEWIBEthTypeAdapter);
  if (o.n_frames >= 0)
    n_frames = std::min(n_frames, size_t(o.n_frames));
  if (n_frames == 0) {
    std::fprintf(stderr, "%s holds no complete frame\n", o.file.c_str());
    return 1;
  }
  std::vector<DUNEWIBEthTypeAdapter> file_frames(n_frames);
  in.seekg(0);
  in.read(reinterpret_cast<char*>(file_frames.data()), std::streamsize(n_frames * sizeof(DUNEWIBEthTypeAdapter)));
  std::printf("Read %zu frames (%zu bytes) from %s\n", n_frames, n_frames * sizeof(DUNEWIBEthTypeAdapter), o.file.c_str());

  if (o.save_adc) { // the raw ADC values after the 14 -> 16 bit expansion, one row per tick, channels in frame order
    std::ofstream adc(o.prefix + "_adc_data.txt");
    for (size_t f = 0; f < n_frames; ++f)
      for (unsigned t = 0; t < 64; ++t) {
        for (unsigned c = 0; c < 64; ++c)
          adc << get_adc(file_frames[f], c, t) << (c == 63 ? '\n' : ',');
      }
  }

  // ---- one frame processor per link on one engine ----
  const bool looping = o.duration > 0;
  std::vector<std::vector<TriggerPrimitive>> tps(o.links);
  std::vector<std::unique_ptr<FrameErrorRegistry>> regs(o.links);
  std::vector<std::unique_ptr<WIBEthFrameProcessor>> procs;
  // each link replays its own copy of the file: the pre-process tasks may rewrite headers (emulator mode)
  std::vector<std::vector<DUNEWIBEthTypeAdapter>> frames(o.links, file_frames);
  try {
    auto engine = std::make_shared<TpgEngine>(o.device, SWTPG_FORMAT_WIBETH, o.links, o.superchunk);
    const DAQEthHeader* h0 = file_frames[0].header();
    for (unsigned l = 0; l < o.links; ++l) {
      regs[l] = std::make_unique<FrameErrorRegistry>();
      procs.push_back(std::make_unique<WIBEthFrameProcessor>(regs[l], engine));
      auto* sink = &tps[l];
      procs[l]->init([sink](TriggerPrimitiveTypeAdapter&& tp) {
        sink->push_back(tp.tp);
        return true;
      });
      RawDataProcessorConf c;
      c.tpg_algorithm = o.algorithm;
      c.tpg_threshold = uint16_t(o.threshold);
      c.crate_id = uint16_t(h0->crate_id);
      c.slot_id = uint16_t(h0->slot_id);
      c.link_id = uint16_t((h0->stream_id + l) & 0xFF);
      if (l != 0) // a replicated link is the same data arriving on another stream of the same crate and slot
        for (auto& f : frames[l])
          f.header()->stream_id = c.link_id;
      c.emulator_mode = looping; // when the file is replayed over and over the timestamps keep running, as on emulated links
      c.block_on_backpressure = true;      // a replay stalls its source instead of dropping frames
      c.tp_timeout = ~uint64_t(0);
      procs[l]->conf(c);
    }
    for (auto& p : procs)
      p->start();

    const auto t0 = std::chrono::steady_clock::now();
    size_t passes = 0;
    do {
      for (size_t f = 0; f < n_frames; ++f)
        for (unsigned l = 0; l < o.links; ++l) {
          procs[l]->preprocess_item(&frames[l][f]);
          procs[l]->postprocess_item(&frames[l][f]);
        }
      ++passes;
    } while (looping && std::chrono::steady_clock::now() - t0 < std::chrono::seconds(o.duration));
    for (auto& p : procs)
      p->stop(); // flushes the ragged tail and delivers every remaining TriggerPrimitive
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    size_t n_tp = 0;
    for (auto& v : tps)
      n_tp += v.size();
    const double frames_done = double(passes) * double(n_frames) * o.links;
    std::printf("Algorithm %s (CUDA), threshold %d: %zu pass(es) of %zu frames on %u link(s) in %.3f s = %.1f frames/s = %.3f Gsamples/s\n",
                o.algorithm.c_str(), o.threshold, passes, n_frames, o.links, secs, frames_done / secs, frames_done * 4096 / secs / 1e9);
    std::printf("Found %zu hits\n", n_tp);
    RawDataProcessorInfo info;
    procs[0]->get_info(info);
    std::printf("link 0: seq id errors %llu, timestamp errors %llu, frames dropped %llu\n", (unsigned long long)info.num_seq_id_errors,
                (unsigned long long)info.num_ts_errors, (unsigned long long)info.num_frames_dropped_busy);

    if (o.save_tp) {
      std::ofstream out(o.prefix + "_trigprim.txt");
      out << "channel,time_start,time_over_threshold,time_peak,adc_integral,adc_peak,type\n";
      for (auto& v : tps) {
        std::stable_sort(v.begin(), v.end(), [](const TriggerPrimitive& a, const TriggerPrimitive& b) {
          return a.time_start != b.time_start ? a.time_start < b.time_start : a.channel < b.channel;
        });
        for (const auto& t : v)
          out << t.channel << ',' << t.time_start << ',' << t.time_over_threshold << ',' << t.time_peak << ',' << t.adc_integral << ','
              << t.adc_peak << ',' << uint32_t(t.type) << '\n';
      }
    }
  } catch (const TPGAlgorithmInexistent& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
