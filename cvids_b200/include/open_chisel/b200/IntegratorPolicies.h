// open_chisel/b200/IntegratorPolicies.h -- the integrator's policy objects in one place: the Truncator and Weighter interfaces
// with the subclasses the reference ships (OC/include/open_chisel/truncation/*.h, OC/include/open_chisel/weighting/*.h). On the
// device none of them is a virtual call per voxel: ProjectionIntegrator::ToC folds them into a chs_integrator (kind + parameter,
// or a per-pixel truncation image for foreign subclasses), which is why they live together here. The reference's header names
// (truncation/Truncator.h, truncation/ConstantTruncator.h, ... weighting/ConstantWeighter.h) include this file.
#ifndef CHISEL_B200_INTEGRATOR_POLICIES_H_
#define CHISEL_B200_INTEGRATOR_POLICIES_H_
#include <memory>
#include <chisel_b200.h>

namespace chisel
{
// ---------------------------------------------------------------------------------------------- truncation distance
// open_chisel/truncation/Truncator.h -- facade; cf. OC/include/open_chisel/truncation/Truncator.h:27-41.
// b200_kind()/b200_param() let the facade hand the three shipped truncators to the device by (kind, parameter); any other
// subclass is evaluated on the host once per pixel (CHS_TRUNC_PER_PIXEL).

class Truncator
{
  public:
    Truncator() = default;
    virtual ~Truncator() {}
    virtual float GetTruncationDistance(float depthReading) const = 0;
    virtual int b200_kind() const { return CHS_TRUNC_PER_PIXEL; }
    virtual float b200_param() const { return 0.0f; }
};
typedef std::shared_ptr<Truncator> TruncatorPtr;
typedef std::shared_ptr<const Truncator> TruncatorConstPtr;

// cf. OC/include/open_chisel/truncation/ConstantTruncator.h:30-57
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

// cf. OC/include/open_chisel/truncation/QuadraticTruncator.h:30-70; the arithmetic itself is chs_truncation (bit-identical).
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

// cf. OC/include/open_chisel/truncation/InverseTruncator.h:27-55 (the truncator ChiselNode instantiates, CR/src/ChiselNode.cpp:98)
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

// ---------------------------------------------------------------------------------------------- update weight
// cf. OC/include/open_chisel/weighting/Weighter.h:27-41
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

// cf. OC/include/open_chisel/weighting/ConstantWeighter.h:30-52: weight / (5 * truncationDist), not a constant
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
