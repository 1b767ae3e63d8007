// open_chisel/truncation/Truncator.h -- facade; cf. OC/include/open_chisel/truncation/Truncator.h:27-41.
// b200_kind()/b200_param() let the facade hand the three shipped truncators to the device by (kind, parameter); any other
// subclass is evaluated on the host once per pixel (CHS_TRUNC_PER_PIXEL).
#ifndef CHISEL_B200_TRUNCATOR_H_
#define CHISEL_B200_TRUNCATOR_H_
#include <memory>
#include <chisel_b200.h>

namespace chisel
{
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
} // namespace chisel
#endif
