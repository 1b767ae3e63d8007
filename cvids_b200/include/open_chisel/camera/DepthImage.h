// open_chisel/camera/DepthImage.h -- facade; cf. OC/include/open_chisel/camera/DepthImage.h:33-103. Caller-owned host image, in
// page-locked memory (b200/PinnedBuffer.h) so that it uploads at full PCIe speed.
#ifndef CHISEL_B200_DEPTHIMAGE_H_
#define CHISEL_B200_DEPTHIMAGE_H_
#include <memory>
#include <vector>
#include <open_chisel/b200/PinnedBuffer.h>

namespace chisel
{
template <class DataType = float>
class DepthImage
{
  public:
    DepthImage() : width(-1), height(-1) {}
    DepthImage(int w, int h) : store(static_cast<size_t>(w) * h), width(w), height(h) {}
    int Index(int row, int col) const { return col + row * width; }
    float DepthAt(int row, int col) const { return static_cast<float>(store[Index(row, col)]); }
    const DataType &At(int row, int col) const { return store[Index(row, col)]; }
    DataType &AtMutable(int row, int col) { return store[Index(row, col)]; }
    const DataType *GetData() const { return store.data(); }
    DataType *GetMutableData() { return store.data(); }
    int GetWidth() const { return width; }
    int GetHeight() const { return height; }

  protected:
    b200::PinnedBuffer<DataType> store;
    int width, height;
};
} // namespace chisel
#endif
