// TEST INFRASTRUCTURE ONLY -- minimal thrust::counting_iterator for the emulator build (see ../../cuda_runtime.h).
#pragma once
namespace thrust {
template <class T>
struct counting_iterator {
  T base;
  explicit counting_iterator(T b = T()) : base(b) {}
  T operator[](long i) const { return (T)(base + i); }
  T operator*() const { return base; }
  counting_iterator operator+(long i) const { return counting_iterator((T)(base + i)); }
};
}  // namespace thrust
