// sphere.h -- sphere approximation by subdivided tetrahedron (optx/sphere.h:20-38).
#ifndef SPHERE_H
#define SPHERE_H

#include <tuple>
#include <vector>

#include <vector_types.h>

typedef std::tuple<float3*, unsigned int, uint3*, unsigned int> Mesh ;

class Sphere {
	public:
		Sphere( const float radius = 1.f, const unsigned int ndiv = 6 ) ;

		const Mesh mesh() ;

	private:
		std::vector<float3> vces_ ; // unique vertices ...
		std::vector<uint3>  ices_ ; // ... as indexed triangles
} ;

#endif // SPHERE_H
