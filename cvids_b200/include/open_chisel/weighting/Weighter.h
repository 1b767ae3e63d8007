// cf. OC/include/open_chisel/weighting/Weighter.h:27-41
#ifndef CHISEL_B200_WEIGHTER_H_
#define CHISEL_B200_WEIGHTER_H_
#include <memory>
namespace chisel
{
class Weighter
{
  public:
    Weighter() = default;
    virtual ~Weighter() {}
    virtual float GetWeight(float surfaceDist, float truncationDist) const = 0;
    // The device evaluates weight / (5 * truncation) (ConstantWeighter); other weighters are not supported by the C ABI.
    virtual bool b200_constant(float *weight) const { (void)weight; return false; }
};
typedef std::shared_ptr<Weighter> WeighterPtr;
typedef std::shared_ptr<const Weighter> WeighterConstPtr;
} // namespace chisel
#endif
