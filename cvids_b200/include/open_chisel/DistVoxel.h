// open_chisel/DistVoxel.h -- host mirror of one voxel of the device pool; cf. OC/include/open_chisel/DistVoxel.h:33-77.
#ifndef CHISEL_B200_DISTVOXEL_H_
#define CHISEL_B200_DISTVOXEL_H_
namespace chisel
{
class DistVoxel
{
  public:
    DistVoxel() : sdf(99999), weight(0) {}
    DistVoxel(float s, float w) : sdf(s), weight(w) {}
    float GetSDF() const { return sdf; }
    void SetSDF(const float &d) { sdf = d; }
    float GetWeight() const { return weight; }
    void SetWeight(const float &w) { weight = w; }

  protected:
    float sdf, weight;
};
} // namespace chisel
#endif
