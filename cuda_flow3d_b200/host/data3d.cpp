// data3d.cpp -- host volume container with the reference's Data3D interface
// (behaviour: src/data_types/data3d.cpp:95-264; implementation is new: whole-volume buffered I/O,
// page-locked storage when a device is present, no dangling pointer after a failed read).
#include "flow3d/data3d.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <utility>
#include <vector>

#include "flow3d_c.h"

Data3D::Data3D() {}

Data3D::Data3D(size_t width, size_t height, size_t depth) { Allocate(width, height, depth); }

Data3D::~Data3D() { Release(); }

bool Data3D::Allocate(size_t width, size_t height, size_t depth) {
  Release();
  const size_t n = width * height * depth;
  if (n == 0) return false;
  void* p = nullptr;
  if (flow3d_host_alloc(&p, n * sizeof(float)) == FLOW3D_OK && p) {
    pinned_ = true;
  } else {
    p = new (std::nothrow) float[n];
    pinned_ = false;
    if (!p) {
      std::printf("Error. Cannot allocate memory on the host.\n");
      return false;
    }
  }
  data_ = static_cast<float*>(p);
  width_ = width;
  height_ = height;
  depth_ = depth;
  return true;
}

void Data3D::Release() {
  if (data_) {
    if (pinned_) flow3d_host_free(data_);
    else delete[] data_;
  }
  data_ = nullptr;
  width_ = height_ = depth_ = 0;
  pinned_ = false;
}

void Data3D::Swap(Data3D& other) {
  if (width_ == other.width_ && height_ == other.height_ && depth_ == other.depth_) {
    std::swap(data_, other.data_);
    std::swap(pinned_, other.pinned_);
  } else {
    std::printf("Error. Cannot swap two Data3D objects (wrong dimensions).\n");
  }
}

void Data3D::ZeroData() {
  if (data_) std::memset(data_, 0, width_ * height_ * depth_ * sizeof(float));
}

namespace {
struct FileCloser {
  void operator()(std::FILE* f) const { if (f) std::fclose(f); }
};
using File = std::unique_ptr<std::FILE, FileCloser>;

// true when exactly `bytes` bytes remain in the file (no more, no less)
bool exact_size(std::FILE* f, unsigned long long bytes) {
  if (std::fseek(f, 0, SEEK_END) != 0) return false;
  const long long size = std::ftell(f);
  std::rewind(f);
  return size >= 0 && (unsigned long long)size == bytes;
}
}  // namespace

bool Data3D::ReadRAWFromFileU8(const char* filename, size_t width, size_t height, size_t depth) {
  File file(std::fopen(filename, "rb"));
  if (!file) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  const size_t n = width * height * depth;
  if (!exact_size(file.get(), n) || !Allocate(width, height, depth)) {
    std::printf("Error reading RAW data from file '%s': wrong dimensions.", filename);
    Release();
    return false;
  }
  const size_t chunk = std::min<size_t>(n, size_t(1) << 22);
  std::vector<unsigned char> buf(chunk);
  size_t done = 0;
  while (done < n) {
    const size_t want = std::min(chunk, n - done);
    if (std::fread(buf.data(), 1, want, file.get()) != want) {
      std::printf("Error reading RAW data from file '%s': wrong dimensions.", filename);
      Release();
      return false;
    }
    for (size_t i = 0; i < want; ++i) data_[done + i] = static_cast<float>(buf[i]);
    done += want;
  }
  return true;
}

bool Data3D::ReadRAWFromFileF32(const char* filename, size_t width, size_t height, size_t depth) {
  File file(std::fopen(filename, "rb"));
  if (!file) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  const size_t n = width * height * depth;
  if (!exact_size(file.get(), (unsigned long long)n * sizeof(float)) || !Allocate(width, height, depth) ||
      std::fread(data_, sizeof(float), n, file.get()) != n) {
    std::printf("Error reading RAW data from file '%s': wrong dimensions.", filename);
    Release();
    return false;
  }
  return true;
}

bool Data3D::WriteRAWToFileU8(const char* filename) const {
  File file(std::fopen(filename, "wb"));
  if (!file) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  const size_t n = width_ * height_ * depth_;
  const size_t chunk = std::min<size_t>(std::max<size_t>(n, 1), size_t(1) << 22);
  std::vector<unsigned char> buf(chunk);
  size_t done = 0;
  while (done < n) {
    const size_t want = std::min(chunk, n - done);
    for (size_t i = 0; i < want; ++i)
      buf[i] = static_cast<unsigned char>(std::min(255.f, std::max(0.f, data_[done + i])));
    if (std::fwrite(buf.data(), 1, want, file.get()) != want) {
      std::printf("Error writing RAW data to file '%s'.", filename);
      return false;
    }
    done += want;
  }
  return true;
}

bool Data3D::WriteRAWToFileF32(const char* filename) const {
  File file(std::fopen(filename, "wb"));
  if (!file) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  const size_t n = width_ * height_ * depth_;
  if (n && std::fwrite(data_, sizeof(float), n, file.get()) != n) {
    std::printf("Error writing RAW data to file '%s'.", filename);
    return false;
  }
  return true;
}

bool Data3D::WriteFlowToFileVTK(const char* filename, const Data3D& flow_u, const Data3D& flow_v,
                                const Data3D& flow_w) {
  if (flow_u.width_ != flow_v.width_ || flow_u.width_ != flow_w.width_ || flow_u.height_ != flow_v.height_ ||
      flow_u.height_ != flow_w.height_ || flow_u.depth_ != flow_v.depth_ || flow_u.depth_ != flow_w.depth_) {
    std::printf("Error. Flow components have different dimensions.\n");
    return false;
  }
  File file(std::fopen(filename, "wb"));
  if (!file) {
    std::printf("Cannot open file '%s'.\n", filename);
    return false;
  }
  const size_t n = flow_u.width_ * flow_u.height_ * flow_u.depth_;
  std::fprintf(file.get(), "# vtk DataFile Version 2.0\n");
  std::fprintf(file.get(), "3D Vector field computed by GpuFlow3D\n");
  std::fprintf(file.get(), "BINARY\n");
  std::fprintf(file.get(), "DATASET STRUCTURED_POINTS\n");
  std::fprintf(file.get(), "DIMENSIONS %zu %zu %zu\n", flow_u.width_, flow_u.height_, flow_u.depth_);
  std::fprintf(file.get(), "ORIGIN 0 0 0\n");
  std::fprintf(file.get(), "SPACING 1 1 1\n");
  std::fprintf(file.get(), "POINT_DATA %zu\n", n);
  std::fprintf(file.get(), "VECTORS vectors float\n");
  const size_t chunk = size_t(1) << 20;
  std::vector<float> buf(3 * std::min(chunk, std::max<size_t>(n, 1)));
  for (size_t done = 0; done < n;) {
    const size_t want = std::min(chunk, n - done);
    for (size_t i = 0; i < want; ++i) {
      buf[3 * i + 0] = flow_u.data_[done + i];
      buf[3 * i + 1] = flow_v.data_[done + i];
      buf[3 * i + 2] = flow_w.data_[done + i];
    }
    if (std::fwrite(buf.data(), sizeof(float), 3 * want, file.get()) != 3 * want) return false;
    done += want;
  }
  return true;
}
