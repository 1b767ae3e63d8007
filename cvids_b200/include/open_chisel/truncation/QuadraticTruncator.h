// cf. OC/include/open_chisel/truncation/QuadraticTruncator.h:30-70; the arithmetic itself is chs_truncation (bit-identical).
#ifndef CHISEL_B200_QUADRATICTRUNCATOR_H_
#define CHISEL_B200_QUADRATICTRUNCATOR_H_
#include "Truncator.h"
namespace chisel
{
class QuadraticTruncator : public Truncator
{
  public:
    QuadraticTruncator() = delete;
    QuadraticTruncator(float scale) : scalingFactor(scale) {}
    float GetTruncationDistance(float reading) const override { return chs_truncation(CHS_TRUNC_QUADRATIC, scalingFactor, reading); }
    float GetScalingFactor() const { return scalingFactor; }
    int b200_kind() const override { return CHS_TRUNC_QUADRATIC; }
    float b200_param() const override { return scalingFactor; }

  protected:
    const float scalingFactor;
};
typedef std::shared_ptr<QuadraticTruncator> QuadraticTruncatorPtr;
} // namespace chisel
#endif
