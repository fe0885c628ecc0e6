// util.h -- host utilities of the rtwo driver: constants, random helpers and the error
// convention of the reference (optx/util.h:17-30, optx/util_cpu.h:18-53, 88-95): a
// failing device call throws std::runtime_error, caught once in main().
#ifndef UTIL_H
#define UTIL_H

#include <cstdlib>
#include <limits>
#include <sstream>
#include <stdexcept>

#include <cuda_runtime.h>

#include "../../include/rtx.h"

namespace util {

const float kInfinity = std::numeric_limits<float>::infinity() ;
const float kNear0    = 1e-8f ;
const float kAcne0    = 1e-3f ;
const float kPi       = 3.14159265358979323846f ;

// the scene recipe draws from the process-global libc stream, in float (optx/util_cpu.h:90-91)
inline float rnd()                                   { return static_cast<float>( rand() )/( static_cast<float>( RAND_MAX )+1.f ) ; }
inline float rnd( const float min, const float max ) { return min+rnd()*( max-min ) ; }

inline float deg( const float rad ) { return rad*180.f/kPi ; }
inline float rad( const float deg ) { return deg*kPi/180.f ; }

inline float clamp( const float x, const float min, const float max ) { return min>x ? min : x>max ? max : x ; }

}

// librtx calls report through rtx_last_error(); the shims turn that into the exception
// the reference's CUDA_CHECK / OPTX_CHECK would have thrown
#define RTX_CHECK( ctx, api )                                                  \
	do {                                                                       \
		if ( ( api ) != 0 ) {                                                  \
			std::ostringstream comment ;                                       \
			comment << "RTX error: " << #api << " : "                          \
				<< rtx_last_error( ctx ) << std::endl ;                        \
			throw std::runtime_error( comment.str() ) ;                        \
		}                                                                      \
	} while ( false )

#define CUDA_CHECK( api )                                                      \
	do {                                                                       \
		cudaError_t e = api ;                                                  \
		if ( e != cudaSuccess ) {                                              \
			std::ostringstream comment ;                                       \
			comment << "CUDA error: " << #api << " : "                         \
				<< cudaGetErrorString( e ) << std::endl ;                      \
			throw std::runtime_error( comment.str() ) ;                        \
		}                                                                      \
	} while ( false )

#endif // UTIL_H
