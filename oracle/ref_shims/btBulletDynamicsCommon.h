// Test-only stand-in: objloader.h only needs btVector3, btScalar and SIMD_EPSILON from this header.
#ifndef ORACLE_SHIM_BT_BULLET_DYNAMICS_COMMON_H
#define ORACLE_SHIM_BT_BULLET_DYNAMICS_COMMON_H
#include <LinearMath/btVector3.h>
#include <float.h>
#define SIMD_EPSILON FLT_EPSILON
#endif
