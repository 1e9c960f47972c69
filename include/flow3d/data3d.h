// data3d.h -- host float volume, x fastest, with headerless RAW I/O.
// Keeps the public interface of the reference's Data3D (src/data_types/data3d.h:22-62) so
// reference-style driver code compiles unchanged; the storage is page-locked when a CUDA device is
// present (fast, asynchronous H2D/D2H) and falls back to ordinary heap memory for I/O-only use.
#ifndef FLOW3D_DATA3D_H_
#define FLOW3D_DATA3D_H_

#include <cstddef>

class Data3D {
 public:
  Data3D();
  Data3D(size_t width, size_t height, size_t depth);
  Data3D(const Data3D&) = delete;
  Data3D& operator=(const Data3D&) = delete;
  ~Data3D();

  inline size_t Width() const { return width_; }
  inline size_t Height() const { return height_; }
  inline size_t Depth() const { return depth_; }
  inline float* DataPtr() { return data_; }
  inline const float* DataPtr() const { return data_; }
  // element (x,y,z) lives at (z*H + y)*W + x  (reference data3d.h:30-32)
  inline float& Data(size_t x, size_t y, size_t z) { return data_[(z * height_ + y) * width_ + x]; }

  void Swap(Data3D& other);  // same-shape volumes only (prints an error otherwise, like the reference)
  void ZeroData();

  // RAW readers: the file must hold exactly width*height*depth samples (reference data3d.cpp:95-178)
  bool ReadRAWFromFileU8(const char* filename, size_t width, size_t height, size_t depth);
  bool ReadRAWFromFileF32(const char* filename, size_t width, size_t height, size_t depth);
  // U8 writer clamps to [0,255] and truncates (reference data3d.cpp:180-208)
  bool WriteRAWToFileU8(const char* filename) const;
  bool WriteRAWToFileF32(const char* filename) const;
  // legacy-VTK STRUCTURED_POINTS vector field, byte-compatible with the reference's writer
  // (data3d.cpp:234-264; it stores native-endian floats under a BINARY header)
  static bool WriteFlowToFileVTK(const char* filename, const Data3D& flow_u, const Data3D& flow_v,
                                 const Data3D& flow_w);

  bool IsPinned() const { return pinned_; }

 private:
  float* data_ = nullptr;
  size_t width_ = 0, height_ = 0, depth_ = 0;
  bool pinned_ = false;
  bool Allocate(size_t width, size_t height, size_t depth);
  void Release();
};

#endif  // FLOW3D_DATA3D_H_
