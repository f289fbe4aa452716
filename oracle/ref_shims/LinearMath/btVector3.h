// Test-only stand-in for Bullet's LinearMath/btVector3.h (Bullet is NOT vendored in the reference,
// SURVEY.md section 0 fact 2).  It restates the scalar (non-SSE) code path bullet3 uses on
// Linux/GCC and MinGW for the handful of btVector3 members the reference's ray.cpp / transducer.h
// touch, so those reference files can be compiled *where they lie* under /root/reference into
// oracle/_ref/ and used as known-answer generators (oracle/Makefile, target _ref).
// btScalar.h in real Bullet includes <math.h> and <stdlib.h>; ray.cpp:188's unqualified abs()
// depends on that, so this shim includes them too.
#ifndef ORACLE_SHIM_BT_VECTOR3_H
#define ORACLE_SHIM_BT_VECTOR3_H
#include <math.h>
#include <stdlib.h>
#include <cmath>

typedef float btScalar;
inline btScalar btSqrt(btScalar x) { return sqrtf(x); }
inline btScalar btSin(btScalar x) { return sinf(x); }
inline btScalar btCos(btScalar x) { return cosf(x); }

class btVector3
{
public:
    btScalar m_floats[4];
    btVector3() {}
    btVector3(const btScalar& x_, const btScalar& y_, const btScalar& z_)
    {
        m_floats[0] = x_; m_floats[1] = y_; m_floats[2] = z_; m_floats[3] = btScalar(0.f);
    }
    const btScalar& getX() const { return m_floats[0]; }
    const btScalar& getY() const { return m_floats[1]; }
    const btScalar& getZ() const { return m_floats[2]; }
    const btScalar& x() const { return m_floats[0]; }
    const btScalar& y() const { return m_floats[1]; }
    const btScalar& z() const { return m_floats[2]; }
    void setValue(const btScalar& x_, const btScalar& y_, const btScalar& z_)
    {
        m_floats[0] = x_; m_floats[1] = y_; m_floats[2] = z_; m_floats[3] = btScalar(0.f);
    }
    btScalar& operator[](int i) { return m_floats[i]; }
    const btScalar& operator[](int i) const { return m_floats[i]; }

    btVector3& operator+=(const btVector3& v)
    {
        m_floats[0] += v.m_floats[0]; m_floats[1] += v.m_floats[1]; m_floats[2] += v.m_floats[2];
        return *this;
    }
    btVector3& operator-=(const btVector3& v)
    {
        m_floats[0] -= v.m_floats[0]; m_floats[1] -= v.m_floats[1]; m_floats[2] -= v.m_floats[2];
        return *this;
    }
    btVector3& operator*=(const btScalar& s)
    {
        m_floats[0] *= s; m_floats[1] *= s; m_floats[2] *= s;
        return *this;
    }
    btVector3& operator/=(const btScalar& s) { return *this *= btScalar(1.0) / s; }
    btScalar dot(const btVector3& v) const
    {
        return m_floats[0] * v.m_floats[0] + m_floats[1] * v.m_floats[1] + m_floats[2] * v.m_floats[2];
    }
    btScalar length2() const { return dot(*this); }
    btScalar length() const { return btSqrt(length2()); }
    btScalar distance(const btVector3& v) const;
    btVector3& normalize() { return *this /= length(); }
    btVector3 normalized() const;
    btVector3 cross(const btVector3& v) const
    {
        return btVector3(m_floats[1] * v.m_floats[2] - m_floats[2] * v.m_floats[1],
                         m_floats[2] * v.m_floats[0] - m_floats[0] * v.m_floats[2],
                         m_floats[0] * v.m_floats[1] - m_floats[1] * v.m_floats[0]);
    }
    btVector3 rotate(const btVector3& wAxis, const btScalar angle) const;
};

inline btVector3 operator+(const btVector3& a, const btVector3& b)
{
    return btVector3(a.m_floats[0] + b.m_floats[0], a.m_floats[1] + b.m_floats[1], a.m_floats[2] + b.m_floats[2]);
}
inline btVector3 operator-(const btVector3& a, const btVector3& b)
{
    return btVector3(a.m_floats[0] - b.m_floats[0], a.m_floats[1] - b.m_floats[1], a.m_floats[2] - b.m_floats[2]);
}
inline btVector3 operator-(const btVector3& v) { return btVector3(-v.m_floats[0], -v.m_floats[1], -v.m_floats[2]); }
inline btVector3 operator*(const btVector3& v, const btScalar& s)
{
    return btVector3(v.m_floats[0] * s, v.m_floats[1] * s, v.m_floats[2] * s);
}
inline btVector3 operator*(const btScalar& s, const btVector3& v) { return v * s; }
inline btVector3 operator/(const btVector3& v, const btScalar& s) { return v * (btScalar(1.0) / s); }

inline btScalar btVector3::distance(const btVector3& v) const { return (v - *this).length(); }
inline btVector3 btVector3::normalized() const { btVector3 nrm = *this; return nrm.normalize(); }
inline btVector3 btVector3::rotate(const btVector3& wAxis, const btScalar angle) const
{
    // wAxis must be a unit length vector
    btVector3 o = wAxis * wAxis.dot(*this);
    btVector3 x_ = *this - o;
    btVector3 y_ = wAxis.cross(*this);
    return (o + x_ * btCos(angle) + y_ * btSin(angle));
}
#endif
