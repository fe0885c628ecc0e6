// rtwo.h -- launch parameter block of the scene API (optx/rtwo.h:26-52) and the few OptiX
// type names user code mentions, mapped onto librtx so that a recipe such as
// `class RTWO : public Scene { RTWO( const OptixDeviceContext& c ) : Scene( c ) {} ... }`
// (optx/rtwo.cxx:133-245) compiles unchanged.
#ifndef RTWO_H
#define RTWO_H

#include <cuda_runtime.h>
#include <vector_types.h>

#include "../../include/rtx.h"
#include "camera.h"
#include "thing.h"

#ifndef OPTIX_VERSION
typedef rtx_ctx*           OptixDeviceContext ;      // optixDeviceContextCreate -> rtx_init
typedef unsigned long long OptixTraversableHandle ;  // top-level structure lives inside the context
struct OptixPipeline_t {} ;                          // programs are compiled into librtx
typedef OptixPipeline_t*   OptixPipeline ;
struct OptixShaderBindingTable {} ;                  // per-thing records live inside the context
typedef cudaStream_t       CUstream ;
#endif

struct LpGeneral { // launch parameter
	uchar4*                image ;
	unsigned int           image_w ;
	unsigned int           image_h ;

	float3*                rawRGB ;  // device, owned by the launcher

	float3*                normals ; // denoiser guide layers (device)
	float3*                albedos ;

	Camera                 camera ;

	unsigned int           spp ;     // samples per pixel
	unsigned int           depth ;   // scatter events per path

	OptixTraversableHandle is_handle ;

	unsigned int*          rpp ;     // AOV rays per pixel (device)

	bool                   picker ;  // thing picker for scene editing
	unsigned int           pick_x ;
	unsigned int           pick_y ;
	unsigned int*          pick_id ; // device
} ;

#endif // RTWO_H
