// cf. OC/include/open_chisel/weighting/ConstantWeighter.h:30-52: weight / (5 * truncationDist), not a constant
#ifndef CHISEL_B200_CONSTANTWEIGHTER_H_
#define CHISEL_B200_CONSTANTWEIGHTER_H_
#include "Weighter.h"
namespace chisel
{
class ConstantWeighter : public Weighter
{
  public:
    ConstantWeighter() : weight(1.0f) {}
    ConstantWeighter(float w) : weight(w) {}
    float GetWeight(float, float truncationDist) const override { return weight / (5 * truncationDist); }
    bool b200_constant(float *w) const override { *w = weight; return true; }

  protected:
    float weight;
};
typedef std::shared_ptr<ConstantWeighter> ConstantWeighterPtr;
} // namespace chisel
#endif
