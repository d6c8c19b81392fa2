// Footage container reader (see footage.h). Reference: camera_isp/BinaryFootageFile.cpp.
#include "footage.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <stdexcept>

namespace s360 {

FootageFile::FootageFile(const std::string& path) : path_(path) {
  fd_ = ::open(path.c_str(), O_RDONLY);
  if (fd_ == -1) throw std::runtime_error("Error opening file " + path + ": " + std::strerror(errno));   // openFile, :38-52
  struct stat info;
  if (fstat(fd_, &info) < 0) {
    const std::string why = std::strerror(errno);
    ::close(fd_);
    fd_ = -1;
    throw std::runtime_error("Error retrieving stat() information for file " + path + ": " + why);       // getFileSize, :54-67
  }
  size_ = (size_t)info.st_size;
  if (size_ < sizeof(FootageHeader)) {
    ::close(fd_);
    fd_ = -1;
    throw std::runtime_error("Error mmap'ing() file " + path + ": shorter than the metadata header");
  }
  void* addr = mmap(nullptr, size_, PROT_READ, MAP_FILE | MAP_PRIVATE, fd_, 0);                          // mapFile, :69-85
  if (addr == MAP_FAILED) {
    const std::string why = std::strerror(errno);
    ::close(fd_);
    fd_ = -1;
    throw std::runtime_error("Error mmap'ing() file " + path + ": " + why);
  }
  base_ = addr;
  std::memcpy(&header_, base_, sizeof header_);                                                          // readMetadataHeader, :119-121
}

FootageFile::~FootageFile() {
  if (base_) munmap(base_, size_);
  if (fd_ != -1) ::close(fd_);
}

size_t FootageFile::frameSize() const { return (size_t)header_.width * header_.height * header_.bitsPerPixel / 8; }

size_t FootageFile::numberOfFrames() const {
  const size_t cams = header_.numberOfCameras, fs = frameSize();
  if (cams == 0 || fs == 0 || size_ < kFootageMetadataSize) return 0;
  return (size_ - kFootageMetadataSize) / fs / cams;
}

const uint8_t* FootageFile::frame(uint32_t frameNumber, uint32_t cameraNumber) const {
  if (cameraNumber >= header_.numberOfCameras) throw std::runtime_error("Camera number out of range");   // :183-185
  const size_t fs = frameSize();
  const size_t off = kFootageMetadataSize + ((size_t)header_.numberOfCameras * frameNumber + cameraNumber) * fs;
  // the reference compares the frame's START with framesBase + mappingSize (4096 bytes past the mapping) and would fault on a frame
  // that straddles the end; here the whole frame must lie inside the file
  if (fs == 0 || off + fs > size_)
    throw std::runtime_error("Accessing frame " + std::to_string(frameNumber) + " is out of range for this file.");
  const uint8_t* addr = static_cast<const uint8_t*>(base_) + off;
  if (cameraNumber == 0 && frameNumber > 0) {   // discardFrameMemory(frameNumber - 1, 0), :146-148, 203-219 (page-aligned part)
    const size_t prev = off - (size_t)header_.numberOfCameras * fs;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE), a = (prev + page - 1) / page * page, b = (prev + fs) / page * page;
    if (b > a) madvise(static_cast<uint8_t*>(base_) + a, b - a, MADV_DONTNEED);
  }
  return addr;
}

}  // namespace s360
