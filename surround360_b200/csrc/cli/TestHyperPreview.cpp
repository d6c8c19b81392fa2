// TestHyperPreview: the reference's fast mono preview straight from the capture's .bin footage (source/test/TestHyperPreview.cpp:40-55,
// 187-424) on the B200 path. Same flags: --binary_prefix (directory holding 0.bin, 1.bin, ...) --file_count --rig_json_file
// --preview_dest --eqr_width --eqr_height --gamma --softmax_coef --top_cam_index --bottom_cam_index --bottom_cam2_index (positions in
// the ascending order of the cameras' serial numbers) --start_frame --frame_count (0 = all) --image_width --image_height
// --enable_pole_removal (accepted; the reference defines but never reads it).
// Footage layout as the reference reads it: camera c of a frame lives in file c % file_count, every file starts with a 4096-byte
// header (magic, timestamp, file index, file count, width, height, bits, cameras in the file) and holds its cameras' packed 8 / 12-bit
// frames one after the other; bytes 4..7 of a frame carry the camera's serial number. Per frame the three selected cameras' samples
// are widened on the GPU (s360_raw_convert = RawConverter's arithmetic, which the reference repeats inline at :371-395) and rendered by
// s360_hyper_preview (PreviewRenderer::render). The reference writes <preview_dest>/<frame:06>.jpg; there is no JPEG encoder here, so
// the same 8-bit image (saturate_cast<uchar> of equirect * 255 / 65535, as cv::imwrite converts the float image) is written losslessly
// as <frame:06>.png.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/s360.h"
#include "png_io.h"

namespace {
struct VrCamException : std::runtime_error { using std::runtime_error::runtime_error; };
void check(int rc) { if (rc) throw VrCamException(s360_last_error()); }

int run(int argc, char** argv) {
  std::map<std::string, std::string> v = {{"image_width", "2048"}, {"image_height", "2048"}, {"binary_prefix", ""}, {"file_count", "2"}, {"start_frame", "0"},
                                          {"frame_count", "0"}, {"rig_json_file", ""}, {"eqr_width", "2048"}, {"eqr_height", "1024"}, {"preview_dest", ""},
                                          {"gamma", "1.0"}, {"top_cam_index", "0"}, {"bottom_cam_index", "15"}, {"bottom_cam2_index", "16"},
                                          {"enable_pole_removal", "true"}, {"softmax_coef", "30.0"}, {"device", "0"}, {"logbuflevel", "0"}, {"log_dir", ""},
                                          {"stderrthreshold", "2"}, {"v", "0"}, {"logtostderr", "false"}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.size() < 2 || a[0] != '-') throw VrCamException("unexpected argument: " + a);
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, val;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); }
    else if (name == "logtostderr" || name == "enable_pole_removal") val = "true";
    else { if (i + 1 >= argc) throw VrCamException("flag '" + name + "' is missing its argument"); val = argv[++i]; }
    if (!v.count(name)) throw VrCamException("unknown command line flag '" + name + "'");
    v[name] = val;
  }
  auto req = [&](const char* n) { if (v[n].empty()) throw VrCamException(std::string("missing required command line argument: ") + n); };
  req("binary_prefix"); req("rig_json_file"); req("preview_dest");   // TestHyperPreview.cpp:189-191
  const bool verbose = std::atoi(v["v"].c_str()) > 0 || v["logtostderr"] == "true";
  const int fileCount = std::atoi(v["file_count"].c_str()), startFrame = std::atoi(v["start_frame"].c_str()), frameCountFlag = std::atoi(v["frame_count"].c_str());
  const int eqrW = std::atoi(v["eqr_width"].c_str()), eqrH = std::atoi(v["eqr_height"].c_str());
  const size_t idxTop = (size_t)std::atoi(v["top_cam_index"].c_str()), idxBottom = (size_t)std::atoi(v["bottom_cam_index"].c_str()),
               idxBottom2 = (size_t)std::atoi(v["bottom_cam2_index"].c_str());
  if (fileCount < 1) throw VrCamException("file_count must be positive");

  std::vector<int> fd(fileCount);
  std::vector<std::string> names;
  for (int i = 0; i < fileCount; ++i) {   // :200-210
    names.push_back(v["binary_prefix"] + "/" + std::to_string(i) + ".bin");
    fd[i] = open(names[i].c_str(), O_RDONLY);
    if (fd[i] < 0) throw VrCamException("error opening binary file. err: " + std::string(strerror(errno)) + " filename: " + names[i]);
  }
  uint32_t cameraCount = 0, imageWidth = 0, imageHeight = 0, nBits = 0;
  for (int i = 0; i < fileCount; ++i) {   // :212-253 (the reference also compares time stamps it never reads: left out)
    uint32_t tag[8];
    if (pread(fd[i], tag, sizeof tag, 0) != (ssize_t)sizeof tag) throw VrCamException("empty binary file: " + names[i]);
    if (tag[0] != 0xfaceb00c) throw VrCamException("binary file not tagged: " + names[i]);
    if (tag[2] != (uint32_t)i) throw VrCamException("binary file name in metadata != loaded .bin file");
    if (tag[3] != (uint32_t)fileCount) throw VrCamException("binary file count in metadata != num .bin files");
    if (i == 0) { imageWidth = tag[4]; imageHeight = tag[5]; nBits = tag[6]; }
    else if (tag[4] != imageWidth || tag[5] != imageHeight || tag[6] != nBits) throw VrCamException("metadata not consistent across binary files!");
    cameraCount += tag[7];
  }
  if (nBits != 8 && nBits != 12) throw VrCamException("footage samples are 8 or 12 bits per pixel");
  if (cameraCount == 0) throw VrCamException("no cameras in the binary files");
  const size_t imageSize = (size_t)imageWidth * imageHeight * nBits / 8;
  std::vector<off_t> pos(fileCount), posInit(fileCount);
  long totalFrameCount = (long)frameCountFlag * cameraCount;
  for (int i = 0; i < fileCount; ++i) {   // :272-300
    if (frameCountFlag == 0) {
      struct stat st;
      fstat(fd[i], &st);
      totalFrameCount += (long)(st.st_size / imageSize);
    }
    posInit[i] = 4096 + (off_t)imageSize * startFrame * ((cameraCount / fileCount) + ((uint32_t)i < cameraCount % fileCount ? 1 : 0));
    pos[i] = posInit[i];
  }
  // camera serial numbers in footage order, from the first frame (:305-321); the flags index their ascending (string) order
  std::vector<unsigned char> imgbuf(imageSize);
  std::vector<std::string> serialNumbers;
  for (uint32_t cameraNumber = 0; cameraNumber < cameraCount; ++cameraNumber) {
    const int idx = (int)(cameraNumber % fileCount);
    const ssize_t got = pread(fd[idx], imgbuf.data(), imageSize, pos[idx]);
    if (got != (ssize_t)imageSize) throw VrCamException("binary file too short for one frame of every camera: " + names[idx]);
    uint32_t serial;
    std::memcpy(&serial, imgbuf.data() + 4, 4);
    serialNumbers.push_back(std::to_string(serial));
    pos[idx] += got;
  }
  std::vector<std::string> sorted(serialNumbers);
  std::sort(sorted.begin(), sorted.end());
  if (idxTop >= sorted.size() || idxBottom >= sorted.size() || idxBottom2 >= sorted.size()) throw VrCamException("camera index beyond the number of cameras in the footage");
  for (int i = 0; i < fileCount; ++i) pos[i] = posInit[i];

  s360_ctx* ctx = nullptr;
  check(s360_ctx_create(std::atoi(v["device"].c_str()), &ctx));
  check(s360_rig_load_file(ctx, v["rig_json_file"].c_str()));
  const size_t px = (size_t)imageWidth * imageHeight;
  std::vector<uint16_t> raw[3];   // initRaws: zero images until the first frame delivers them (:330)
  for (auto& r : raw) r.assign(px, 0);
  std::vector<float> eqr((size_t)eqrW * eqrH * 3);
  std::vector<uint8_t> out8(eqr.size());
  const float kNorm16to8 = 255.0f / 65535.0f;
  bool isDone = false;
  const long frames = totalFrameCount / cameraCount;
  if (verbose) std::cerr << "Generating previews...\n";
  for (long frameNumber = startFrame; frameNumber < startFrame + frames; ++frameNumber) {
    for (uint32_t cameraNumber = 0; cameraNumber < cameraCount; ++cameraNumber) {
      const int idx = (int)(cameraNumber % fileCount);
      const std::string& sn = serialNumbers[cameraNumber];
      const int which = sn == sorted[idxTop] ? 0 : sn == sorted[idxBottom] ? 1 : sn == sorted[idxBottom2] ? 2 : -1;
      if (which < 0) { pos[idx] += (off_t)imageSize; continue; }
      const ssize_t got = pread(fd[idx], imgbuf.data(), imageSize, pos[idx]);
      if (got <= 0) { isDone = true; if (verbose) std::cerr << "Reached EOF\n"; break; }   // :356-366
      if (got != (ssize_t)imageSize) throw VrCamException("truncated frame in " + names[idx]);
      pos[idx] += got;
      check(s360_raw_convert(ctx, imgbuf.data(), (int)nBits, (int)imageWidth, (int)imageHeight, raw[which].data()));   // :371-395
    }
    if (verbose) std::cerr << "Rendering frame " << frameNumber << "...\n";
    check(s360_hyper_preview(ctx, raw[0].data(), raw[1].data(), raw[2].data(), (int)imageWidth, (int)imageHeight, eqrW, eqrH, std::atof(v["gamma"].c_str()),
                             std::atof(v["softmax_coef"].c_str()), eqr.data()));
    for (size_t i = 0; i < eqr.size(); ++i) {   // imwrite(eqrImage * kNorm16to8): float -> uchar by saturate_cast (round half to even)
      const float f = eqr[i] * kNorm16to8;
      const long r = std::lrintf(f);
      out8[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
    std::ostringstream ss;
    ss << std::setw(6) << std::setfill('0') << frameNumber;
    pngio::write(v["preview_dest"] + "/" + ss.str() + ".png", out8.data(), eqrW, eqrH, 3);
    if (isDone) break;
  }
  if (verbose) std::cerr << "Closing binary files...\n";
  for (int i = 0; i < fileCount; ++i) close(fd[i]);
  s360_ctx_destroy(ctx);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << "TestHyperPreview: " << e.what() << "\n";
    return 1;
  }
}
