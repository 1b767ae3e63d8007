// open_chisel/camera/ColorImage.h -- facade; cf. OC/include/open_chisel/camera/ColorImage.h:30-134. Channel order as the
// reference reads it: 1 = mono, 3 = BGR, 4 = BGRA.
#ifndef CHISEL_B200_COLORIMAGE_H_
#define CHISEL_B200_COLORIMAGE_H_
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
#include <open_chisel/b200/PinnedBuffer.h>

namespace chisel
{
template <class DataType = uint8_t>
struct Color
{
    DataType red, green, blue, alpha;
};

template <class DataType = uint8_t>
class ColorImage
{
  public:
    ColorImage() : width(-1), height(-1), numChannels(0) {}
    ColorImage(int w, int h, size_t channels) : store(static_cast<size_t>(w) * h * channels), width(w), height(h), numChannels(channels) {}
    int Index(int row, int col, int channel) const { return (col + row * width) * static_cast<int>(numChannels) + channel; }
    const DataType &At(int row, int col, int channel) const { return store[Index(row, col, channel)]; }
    DataType &AtMutable(int row, int col, int channel) { return store[Index(row, col, channel)]; }
    const DataType *GetData() const { return store.data(); }
    DataType *GetMutableData() { return store.data(); }
    int GetWidth() const { return width; }
    int GetHeight() const { return height; }
    size_t GetNumChannels() const { return numChannels; }

  protected:
    b200::PinnedBuffer<DataType> store;
    int width, height;
    size_t numChannels;
};
} // namespace chisel
#endif
