// cf. OC/include/open_chisel/truncation/ConstantTruncator.h:30-57
#ifndef CHISEL_B200_CONSTANTTRUNCATOR_H_
#define CHISEL_B200_CONSTANTTRUNCATOR_H_
#include "Truncator.h"
namespace chisel
{
class ConstantTruncator : public Truncator
{
  public:
    ConstantTruncator() : truncationDistance(0) {}
    ConstantTruncator(float value) : truncationDistance(value) {}
    void SetTruncationDistance(float value) { truncationDistance = value; }
    float GetTruncationDistance(float) const override { return truncationDistance; }
    int b200_kind() const override { return CHS_TRUNC_CONSTANT; }
    float b200_param() const override { return truncationDistance; }

  protected:
    float truncationDistance;
};
typedef std::shared_ptr<ConstantTruncator> ConstantTruncatorPtr;
} // namespace chisel
#endif
