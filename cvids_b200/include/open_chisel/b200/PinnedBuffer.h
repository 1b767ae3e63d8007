// open_chisel/b200/PinnedBuffer.h -- page-locked host storage for the facade's images (chs_host_alloc): the frames a caller hands
// to Chisel::IntegrateDepthScan[Color] then cross PCIe at full speed straight from the caller's buffer, without a staging copy.
// Falls back to ordinary memory when no page-locked memory can be had (e.g. no CUDA device: the library then fails later, loudly).
#ifndef CHISEL_B200_PINNEDBUFFER_H_
#define CHISEL_B200_PINNEDBUFFER_H_
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <chisel_b200.h>

namespace chisel
{
namespace b200
{
template <class T>
class PinnedBuffer
{
  public:
    PinnedBuffer() : data_(nullptr), n_(0), pinned_(false) {}
    explicit PinnedBuffer(size_t n) : data_(nullptr), n_(0), pinned_(false) { Allocate(n); }
    PinnedBuffer(const PinnedBuffer &o) : data_(nullptr), n_(0), pinned_(false)
    {
        Allocate(o.n_);
        if (n_)
            std::memcpy(data_, o.data_, n_ * sizeof(T));
    }
    PinnedBuffer &operator=(const PinnedBuffer &o)
    {
        if (this != &o)
        {
            Release();
            Allocate(o.n_);
            if (n_)
                std::memcpy(data_, o.data_, n_ * sizeof(T));
        }
        return *this;
    }
    ~PinnedBuffer() { Release(); }
    T *data() { return data_; }
    const T *data() const { return data_; }
    size_t size() const { return n_; }
    T &operator[](size_t i) { return data_[i]; }
    const T &operator[](size_t i) const { return data_[i]; }

  private:
    void Allocate(size_t n)
    {
        n_ = n;
        if (!n)
            return;
        data_ = static_cast<T *>(chs_host_alloc(n * sizeof(T)));
        pinned_ = data_ != nullptr;
        if (!data_)
            data_ = static_cast<T *>(std::malloc(n * sizeof(T)));
        std::memset(static_cast<void *>(data_), 0, n * sizeof(T));
    }
    void Release()
    {
        if (data_)
        {
            if (pinned_)
                chs_host_free(data_);
            else
                std::free(data_);
        }
        data_ = nullptr;
        n_ = 0;
    }
    T *data_;
    size_t n_;
    bool pinned_;
};
} // namespace b200
} // namespace chisel
#endif
