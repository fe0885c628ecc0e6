// object.h -- triangle meshes from Wavefront OBJ scene files (optx/object.h:20-33).
#ifndef OBJECT_H
#define OBJECT_H

#include <string>
#include <tuple>
#include <vector>

#include <vector_types.h>

typedef std::tuple<float3*, unsigned int, uint3*, unsigned int> Mesh ;

class Object {
	public:
		const Mesh operator[] ( unsigned int m ) ;

		Object( const std::string& wavefront ) ;
		// one submesh taken from memory (e.g. Sphere::mesh()), no file involved
		Object( const Mesh& mesh ) ;

		size_t size() const { return vces_.size() ; } ;

	private:
		std::vector<std::vector<float3>> vces_ ; // per submesh: unique vertices ...
		std::vector<std::vector<uint3>>  ices_ ; // ... as indexed triangles

		void procWavefrontObj( const std::string& wavefront ) ;
} ;

#endif // OBJECT_H
