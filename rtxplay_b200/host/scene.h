// scene.h -- the scene API of the reference (optx/scene.h:19-63) over librtx: meshes become
// bottom-level LBVHs (instead of OptiX GASes), things become instances of the top-level
// LBVH (instead of an IAS).  A user subclasses Scene and implements load().
#ifndef SCENE_H
#define SCENE_H

#include <vector>

#include "object.h"
#include "rtwo.h"
#include "thing.h"

class Scene {
	public:
		const Thing& operator[] ( unsigned int i ) { return things_[i] ; } ;

		Scene( const OptixDeviceContext& optx_context ) ;
		virtual ~Scene() noexcept ( false ) ;

		virtual unsigned int load() = 0 ;

		unsigned int add( Object& object ) ;                     // build the LBVH of the object's (concatenated) submeshes
		unsigned int add( Thing& thing, unsigned int object ) ;  // create thing and connect with that mesh

		// additive: a thing that is the analytic unit sphere of the CPU path (sphere.h:20-48)
		unsigned int addAnalyticSphere() ;

		bool set( unsigned int thing, const float* transform ) ; // set thing's 3x4 transform
		bool get( unsigned int thing, float* transform ) ;       // get thing's transform

		void build( OptixTraversableHandle* is_handle ) ;        // top-level build
		void update( OptixTraversableHandle is_handle ) ;        // top-level refit after set()

	private:
		rtx_ctx*                  ctx_ ;
		std::vector<unsigned int> meshes_ ; // per object: librtx mesh id
		std::vector<Thing>        things_ ;
		bool                      built_ ;
} ;

#endif // SCENE_H
