// cf. OC/include/open_chisel/truncation/InverseTruncator.h:27-55 (the truncator ChiselNode instantiates, CR/src/ChiselNode.cpp:98)
#ifndef CHISEL_B200_INVERSETRUNCATOR_H_
#define CHISEL_B200_INVERSETRUNCATOR_H_
#include "Truncator.h"
namespace chisel
{
class InverseTruncator : public Truncator
{
  public:
    InverseTruncator() : scalingFactor(1.0f) {}
    InverseTruncator(float scale) : scalingFactor(scale) {}
    float GetTruncationDistance(float reading) const override { return chs_truncation(CHS_TRUNC_INVERSE, scalingFactor, reading); }
    int b200_kind() const override { return CHS_TRUNC_INVERSE; }
    float b200_param() const override { return scalingFactor; }

  protected:
    const float scalingFactor;
};
typedef std::shared_ptr<InverseTruncator> InverseTruncatorPtr;
} // namespace chisel
#endif
