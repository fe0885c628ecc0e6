// v.h -- float3 algebra of the host side (optx/v.h:24-53, 80-81), namespace V.
#ifndef V_H
#define V_H

#include <cmath>

#include <vector_functions.h>
#include <vector_types.h>

#include "util.h"

namespace V {

inline float3 operator + ( const float3& a, const float3& b ) { return make_float3( a.x+b.x, a.y+b.y, a.z+b.z ) ; }
inline float3 operator - ( const float3& a, const float3& b ) { return make_float3( a.x-b.x, a.y-b.y, a.z-b.z ) ; }
inline float3 operator - ( const float3& a )                  { return make_float3( -a.x, -a.y, -a.z ) ; }
inline float3 operator * ( const float a, const float3& b )   { return make_float3( a*b.x, a*b.y, a*b.z ) ; }
inline float3 operator * ( const float3& a, const float b )   { return make_float3( a.x*b, a.y*b, a.z*b ) ; }
inline float3 operator * ( const float3& a, const float3& b ) { return make_float3( a.x*b.x, a.y*b.y, a.z*b.z ) ; }
inline float3 operator / ( const float3& a, const float b )   { return 1.f/b*a ; }

inline float  dot  ( const float3& a, const float3& b ) { return a.x*b.x+a.y*b.y+a.z*b.z ; }
inline float  len  ( const float3& v )                  { return sqrtf( dot( v, v ) ) ; }
inline float3 cross( const float3& a, const float3& b ) { return make_float3( a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x ) ; }
inline float3 unitV( const float3& v )                  { return 1.f/len( v )*v ; }
inline bool   near0( const float3& v )                  { return fabsf( v.x )<util::kNear0 && fabsf( v.y )<util::kNear0 && fabsf( v.z )<util::kNear0 ; }

// host-side random vectors for scene recipes; members are drawn x, y, z (brace order)
inline float3 rnd()                                   { const float x = util::rnd(), y = util::rnd(), z = util::rnd() ; return make_float3( x, y, z ) ; }
inline float3 rnd( const float min, const float max ) { const float x = util::rnd( min, max ), y = util::rnd( min, max ), z = util::rnd( min, max ) ; return make_float3( x, y, z ) ; }

}

#endif // V_H
