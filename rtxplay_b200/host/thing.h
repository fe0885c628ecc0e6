// thing.h -- Thing / Optics records of the scene API (optx/thing.h:17-52).  Field names
// and the type enumeration are the reference's, so scene recipes compile unchanged; the
// device pointers a Thing carried for OptiX's SBT stay for source compatibility and are
// filled with the mesh's device buffers by Scene::add.
#ifndef THING_H
#define THING_H

#include <vector_types.h>

struct Diffuse { float3 albedo ; } ;               // wavefront MTL: Kd
struct Reflect { float3 albedo ; float fuzz ; } ;  // Kd, sharpness
struct Refract { float  index ; } ;                // Ni

struct Optics {
	enum { TYPE_DIFFUSE, TYPE_REFLECT, TYPE_REFRACT, TYPE_NUM } ;

	int type ;

	union {
		Diffuse diffuse ;
		Reflect reflect ;
		Refract refract ;
	} ;
} ;

struct Thing {
	float3* vces ;
	uint3*  ices ;

	Optics  optics ;
} ;

#endif // THING_H
