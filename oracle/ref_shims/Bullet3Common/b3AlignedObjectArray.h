// Test-only stand-in for Bullet3Common/b3AlignedObjectArray.h (see btVector3.h shim header).
#ifndef ORACLE_SHIM_B3_ALIGNED_OBJECT_ARRAY_H
#define ORACLE_SHIM_B3_ALIGNED_OBJECT_ARRAY_H
#include <vector>
template <typename T>
class b3AlignedObjectArray
{
    std::vector<T> v_;
public:
    int size() const { return (int)v_.size(); }
    void push_back(const T& t) { v_.push_back(t); }
    T& at(int i) { return v_[i]; }
    const T& at(int i) const { return v_[i]; }
    T& operator[](int i) { return v_[i]; }
    const T& operator[](int i) const { return v_[i]; }
};
#endif
