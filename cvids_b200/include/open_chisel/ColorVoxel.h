// open_chisel/ColorVoxel.h -- host mirror; cf. OC/include/open_chisel/ColorVoxel.h:33-100.
#ifndef CHISEL_B200_COLORVOXEL_H_
#define CHISEL_B200_COLORVOXEL_H_
#include <stdint.h>
namespace chisel
{
class ColorVoxel
{
  public:
    ColorVoxel() : red(0), green(0), blue(0), weight(0) {}
    ColorVoxel(uint8_t r, uint8_t g, uint8_t b, uint8_t w) : red(r), green(g), blue(b), weight(w) {}
    uint8_t GetRed() const { return red; }
    uint8_t GetGreen() const { return green; }
    uint8_t GetBlue() const { return blue; }
    uint8_t GetWeight() const { return weight; }

  protected:
    uint8_t red, green, blue, weight;
};
} // namespace chisel
#endif
