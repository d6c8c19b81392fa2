// The capture software's footage container (".bin"), as the reference's Unpacker reads it (camera_isp/BinaryFootageFile.h:16-27,
// BinaryFootageFile.cpp:30-219): a 32-byte header at offset 0, frames from offset 4096 on, frame f of camera c at
// 4096 + (numberOfCameras * f + c) * frameSize, frameSize = width * height * bitsPerPixel / 8. Read-only mapping; host I/O in front
// of the GPU path (the frames are copied to the device packed and widened there, isp_kernels.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace s360 {

struct FootageHeader {   // BinaryFootageFile::MetadataHeader, BinaryFootageFile.h:18-27 (eight little-endian 32-bit words)
  uint32_t magic, timestamp, fileIndex, fileCount, width, height, bitsPerPixel, numberOfCameras;
};
constexpr uint32_t kFootageMagic = 0xfaceb00c;   // isMetadataValid, BinaryFootageFile.cpp:135-137
constexpr size_t kFootageMetadataSize = 4096;    // calculateFrameAddress, BinaryFootageFile.cpp:180

class FootageFile {
 public:
  explicit FootageFile(const std::string& path);   // BinaryFootageFile(path) + open()
  ~FootageFile();
  FootageFile(const FootageFile&) = delete;
  FootageFile& operator=(const FootageFile&) = delete;
  const FootageHeader& header() const { return header_; }
  bool metadataValid() const { return header_.magic == kFootageMagic; }
  size_t frameSize() const;        // getFrameSize, :170-172 (64-bit arithmetic here; the reference's product is 32-bit)
  size_t numberOfFrames() const;   // getNumberOfFrames, :160-168
  // getFrame, :143-154: pointer into the mapping; throws when the camera or the frame lies outside the file. Like the reference, asking
  // for camera 0 of frame f > 0 releases the pages of camera 0 of frame f - 1 (sequential access pattern).
  const uint8_t* frame(uint32_t frameNumber, uint32_t cameraNumber) const;

 private:
  std::string path_;
  int fd_ = -1;
  void* base_ = nullptr;
  size_t size_ = 0;
  FootageHeader header_{};
};

}  // namespace s360
