// Test-only stand-in for Bullet3Common/b3MinMax.h (see btVector3.h shim header).
#ifndef ORACLE_SHIM_B3_MINMAX_H
#define ORACLE_SHIM_B3_MINMAX_H
template <class T> inline const T& b3Min(const T& a, const T& b) { return a < b ? a : b; }
template <class T> inline const T& b3Max(const T& a, const T& b) { return a > b ? a : b; }
#endif
