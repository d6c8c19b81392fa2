// Unpacker: the reference's footage unpacker (source/camera_isp/Unpacker.cpp:30-222) on the B200 path. Same flags:
//   --bin_list (comma-separated .bin footage files) --isp_dir (<serial>.json per camera) --output_dir --output_raw_dir (optional)
//   --start_frame --frame_count (0 = all).
// Per frame and camera: the packed 8 / 12-bit samples go to the device as they lie in the file, are widened and run through the ISP
// there (s360_isp_process_footage: RawConverter + loadImage + the pipeline, tone map on, 16 output bits), and come back as the
// 16-bit colour image the reference writes to <output_dir>/<serial>/<frame:06>.png; with --output_raw_dir the widened samples are
// written to <output_raw_dir>/<serial>/<frame:06>.tiff (16-bit grey, uncompressed, as the reference's tiffParams ask). A camera
// without <isp_dir>/<serial>.json is unpacked to raw only, as in the reference. Afterwards the serial-number directories are renamed
// cam0, cam1, ... in ascending order of serial number. The ISP arithmetic is the reference's C++ CameraIsp (bit-identical), not its
// Halide-generated variant ("Soft ISP").
#include <sys/stat.h>
#include <sys/types.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/s360.h"
#include "png_io.h"

namespace {
struct VrCamException : std::runtime_error { using std::runtime_error::runtime_error; };
void check(int rc) { if (rc) throw VrCamException(s360_last_error()); }

void makeCameraDir(const std::string& outDir, uint32_t serial) { mkdir((outDir + "/" + std::to_string(serial)).c_str(), 0755); }   // Unpacker.cpp:44-48
std::string createFilename(const std::string& outDir, uint32_t serial, size_t frameIndex, const std::string& extension) {       // :50-63
  std::ostringstream s;
  s << outDir << "/" << serial << "/" << std::setfill('0') << std::setw(6) << frameIndex << extension;
  return s.str();
}

// 16-bit grey baseline TIFF, little endian, one uncompressed strip (imwrite with {TIFFTAG_COMPRESSION, COMPRESSION_NONE})
void writeTiff16(const std::string& path, const uint16_t* data, int w, int h) {
  std::vector<uint8_t> out;
  auto u16 = [&](unsigned v) { out.push_back((uint8_t)v); out.push_back((uint8_t)(v >> 8)); };
  auto u32 = [&](unsigned v) { u16(v & 0xffff); u16(v >> 16); };
  const unsigned dataBytes = (unsigned)((size_t)w * h * 2), ifdOff = 8 + dataBytes;
  out.push_back('I'); out.push_back('I'); u16(42); u32(ifdOff);
  out.insert(out.end(), reinterpret_cast<const uint8_t*>(data), reinterpret_cast<const uint8_t*>(data) + dataBytes);
  auto entry = [&](unsigned tag, unsigned type, unsigned count, unsigned value) { u16(tag); u16(type); u32(count); if (type == 3 && count == 1) { u16(value); u16(0); } else u32(value); };
  u16(9);
  entry(256, 4, 1, (unsigned)w);    // ImageWidth
  entry(257, 4, 1, (unsigned)h);    // ImageLength
  entry(258, 3, 1, 16);             // BitsPerSample
  entry(259, 3, 1, 1);              // Compression: none
  entry(262, 3, 1, 1);              // Photometric: black is zero
  entry(273, 4, 1, 8);              // StripOffsets
  entry(277, 3, 1, 1);              // SamplesPerPixel
  entry(278, 4, 1, (unsigned)h);    // RowsPerStrip
  entry(279, 4, 1, dataBytes);      // StripByteCounts
  u32(0);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw VrCamException("failed to write image: " + path);
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  if (!ok) throw VrCamException("failed to write image: " + path);
}

int run(int argc, char** argv) {
  std::map<std::string, std::string> v = {{"isp_dir", ""}, {"output_dir", ""}, {"output_raw_dir", ""}, {"bin_list", ""}, {"start_frame", "0"},
                                          {"frame_count", "0"}, {"device", "0"}, {"logbuflevel", "0"}, {"log_dir", ""}, {"stderrthreshold", "2"},
                                          {"v", "0"}, {"logtostderr", "false"}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.size() < 2 || a[0] != '-') throw VrCamException("unexpected argument: " + a);
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, val;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); }
    else if (name == "logtostderr") val = "true";
    else { if (i + 1 >= argc) throw VrCamException("flag '" + name + "' is missing its argument"); val = argv[++i]; }
    if (!v.count(name)) throw VrCamException("unknown command line flag '" + name + "'");
    v[name] = val;
  }
  auto req = [&](const char* n) { if (v[n].empty()) throw VrCamException(std::string("missing required command line argument: ") + n); };
  req("isp_dir"); req("output_dir"); req("bin_list");   // Unpacker.cpp:67-69
  const bool verbose = std::atoi(v["v"].c_str()) > 0 || v["logtostderr"] == "true";
  const std::string outDir = v["output_dir"], rawDir = v["output_raw_dir"];
  const int startFrame = std::atoi(v["start_frame"].c_str()), frameCountFlag = std::atoi(v["frame_count"].c_str());

  s360_ctx* ctx = nullptr;
  check(s360_ctx_create(std::atoi(v["device"].c_str()), &ctx));
  std::set<uint32_t> serialNumbers;   // merged over the files (Unpacker.cpp:196-201)
  std::map<uint32_t, s360_isp*> isps; // one configured ISP per camera serial; nullptr = no <serial>.json
  std::istringstream binList(v["bin_list"]);
  std::string binFile;
  while (std::getline(binList, binFile, ',')) {
    if (verbose) std::cerr << "Reading " << binFile << "...\n";
    s360_footage* file = nullptr;
    check(s360_footage_open(binFile.c_str(), &file));
    s360_footage_header hd;
    uint64_t frames = 0, frameBytes = 0;
    check(s360_footage_info(file, &hd, &frames, &frameBytes));
    // BinaryFootageFile::readMetadataHeader prints the header (BinaryFootageFile.cpp:121-133)
    std::cout << "Metadata:\nmagic = " << std::hex << hd.magic << std::dec << "\ntimestamp = " << hd.timestamp << "\nfileIndex = " << hd.file_index
              << "\nfileCount = " << hd.file_count << "\nwidth = " << hd.width << "\nheight = " << hd.height << "\nbpp = " << hd.bits_per_pixel
              << "\nnumberOfCameras = " << hd.number_of_cameras << std::endl;
    const int numCameras = (int)hd.number_of_cameras;
    if (numCameras == 0) { if (verbose) std::cerr << "No cameras found...\n"; s360_footage_close(file); continue; }
    if (hd.bits_per_pixel != 8 && hd.bits_per_pixel != 12) throw VrCamException("footage samples are 8 or 12 bits per pixel, not " + std::to_string(hd.bits_per_pixel));
    const int endFrameMax = (int)frames - 1;
    int endFrame = frameCountFlag == 0 ? endFrameMax : startFrame + frameCountFlag - 1;
    if (startFrame > endFrameMax)   // Unpacker.cpp:103-108
      throw VrCamException("Start frame (" + std::to_string(startFrame) + ") larger than total number of frames (" + std::to_string(endFrameMax) + ")");
    if (endFrame > endFrameMax) {
      std::cerr << "End frame (" << endFrame << ") larger than total number of frames (" << endFrameMax << ")\n";
      endFrame = endFrameMax;
    }
    const int frameCount = endFrame - startFrame + 1;
    const int w = (int)hd.width, h = (int)hd.height, bits = (int)hd.bits_per_pixel;
    const size_t px = (size_t)w * h;
    // per camera: the 16-bit colour image and (only with --output_raw_dir) the widened samples
    std::vector<std::vector<uint16_t>> hOut(numCameras), hRaw(numCameras);
    for (int c = 0; c < numCameras; ++c) { hOut[c].resize(px * 3); if (!rawDir.empty()) hRaw[c].resize(px); }
    int percentDonePrev = 0;
    for (int frameIndex = startFrame; frameIndex <= endFrame; ++frameIndex) {
      std::vector<uint32_t> serials(numCameras);
      std::vector<char> hasIsp(numCameras, 0);
      std::vector<std::future<void>> tasks;
      for (int c = 0; c < numCameras; ++c) {
        const uint8_t* frame = nullptr;
        check(s360_footage_frame(file, (uint32_t)frameIndex, (uint32_t)c, &frame));
        uint32_t serial;
        std::memcpy(&serial, frame + 4, 4);   // reinterpret_cast<const uint32_t*>(frame)[1], Unpacker.cpp:125
        serials[c] = serial;
        if (verbose) std::cerr << "Looking for serial number " << serial << "\n";
        if (!serialNumbers.count(serial)) {
          serialNumbers.insert(serial);
          makeCameraDir(outDir, serial);
          if (!rawDir.empty()) makeCameraDir(rawDir, serial);
        }
        if (!isps.count(serial)) {
          const std::string fname = v["isp_dir"] + "/" + std::to_string(serial) + ".json";
          std::ifstream ifs(fname, std::ios::in);
          if (!ifs) { isps[serial] = nullptr; }   // "we can still unpack raws", Unpacker.cpp:165-168
          else {
            const std::string json((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
            s360_isp* isp = nullptr;
            check(s360_isp_create(json.c_str(), 16, &isp));   // kOutputBpp, Unpacker.cpp:175
            check(s360_isp_set_bits_per_pixel(isp, bits));     // :177
            check(s360_isp_set_tone_map(isp, 1));              // :178
            isps[serial] = isp;
          }
        }
        s360_isp* isp = isps[serial];
        if (!isp && verbose) std::cerr << "Cannot convert to RGB, file not found: " << v["isp_dir"] << "/" << serial << ".json\n";
        if (isp || !rawDir.empty())
          check(s360_isp_process_footage(ctx, isp, frame, bits, w, h, hOut[c].data(), rawDir.empty() ? nullptr : hRaw[c].data()));
        hasIsp[c] = isp != nullptr;
        // encoding is host work: one task per camera (the reference runs a task per camera, Unpacker.cpp:117-119), overlapping
        // the next camera's GPU pass
        tasks.push_back(std::async(std::launch::async, [&, c, frameIndex] {
          if (!rawDir.empty()) writeTiff16(createFilename(rawDir, serials[c], (size_t)frameIndex, ".tiff"), hRaw[c].data(), w, h);
          if (hasIsp[c]) pngio::write16(createFilename(outDir, serials[c], (size_t)frameIndex, ".png"), hOut[c].data(), w, h);
        }));
      }
      for (auto& t : tasks) t.get();
      const int percentDoneCurr = (frameIndex - startFrame + 1) * 100 / frameCount;
      if (verbose && percentDoneCurr != percentDonePrev) std::cerr << "Percent done " << percentDoneCurr << "%\n";
      percentDonePrev = percentDoneCurr;
    }
    s360_footage_close(file);
  }
  // Rename output directories from serial number to camN, sorted by serial number (Unpacker.cpp:203-221)
  size_t ordinal = 0;
  for (uint32_t serial : serialNumbers) {
    const std::string oldDir = outDir + "/" + std::to_string(serial), newDir = outDir + "/cam" + std::to_string(ordinal);
    std::rename(oldDir.c_str(), newDir.c_str());
    ++ordinal;
  }
  for (auto& kv : isps) if (kv.second) s360_isp_destroy(kv.second);
  s360_ctx_destroy(ctx);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << "Unpacker: " << e.what() << "\n";
    return 1;
  }
}
