// scene.cxx -- Scene over the C ABI.  Error convention as in the reference: out-of-range
// ids give false, device failures throw std::runtime_error (optx/scene.cxx:198-225,
// optx/util_cpu.h:18-53).
#include "scene.h"
#include "util.h"

Scene::Scene( const OptixDeviceContext& optx_context ) : ctx_( optx_context ), built_( false ) {
}

Scene::~Scene() noexcept ( false ) {
	// geometry and acceleration structures belong to the context (rtx_shutdown frees them)
}

unsigned int Scene::add( Object& object ) {
	// submeshes are concatenated into one mesh, later indices shifted past earlier vertices
	std::vector<float3> vces ;
	std::vector<uint3>  ices ;
	for ( unsigned int o = 0 ; object.size()>o ; o++ ) {
		float3* v ; unsigned int nv ; uint3* i ; unsigned int ni ;
		std::tie( v, nv, i, ni ) = object[o] ;
		const unsigned int base = static_cast<unsigned int>( vces.size() ) ;
		vces.insert( vces.end(), v, v+nv ) ;
		for ( unsigned int k = 0 ; k<ni ; k++ )
			ices.push_back( { i[k].x+base, i[k].y+base, i[k].z+base } ) ;
	}
	uint32_t mesh = 0 ;
	RTX_CHECK( ctx_, rtx_mesh_create( ctx_, &vces[0].x, static_cast<uint32_t>( vces.size() ), &ices[0].x, static_cast<uint32_t>( ices.size() ), &mesh ) ) ;
	meshes_.push_back( mesh ) ;
	return static_cast<unsigned int>( meshes_.size()-1 ) ;
}

unsigned int Scene::addAnalyticSphere() {
	uint32_t mesh = 0 ;
	RTX_CHECK( ctx_, rtx_sphere_create( ctx_, &mesh ) ) ;
	meshes_.push_back( mesh ) ;
	return static_cast<unsigned int>( meshes_.size()-1 ) ;
}

static rtx_optics optics_of( const Thing& thing ) {
	rtx_optics o = {} ;
	o.type = thing.optics.type ;
	switch ( thing.optics.type ) {
		case Optics::TYPE_DIFFUSE:
			o.albedo[0] = thing.optics.diffuse.albedo.x ; o.albedo[1] = thing.optics.diffuse.albedo.y ; o.albedo[2] = thing.optics.diffuse.albedo.z ;
			break ;
		case Optics::TYPE_REFLECT:
			o.albedo[0] = thing.optics.reflect.albedo.x ; o.albedo[1] = thing.optics.reflect.albedo.y ; o.albedo[2] = thing.optics.reflect.albedo.z ;
			o.fuzz = thing.optics.reflect.fuzz ;
			break ;
		case Optics::TYPE_REFRACT:
			o.index = thing.optics.refract.index ;
			break ;
	}
	return o ;
}

unsigned int Scene::add( Thing& thing, unsigned int object ) {
	if ( object>=meshes_.size() )
		throw std::runtime_error( "Scene::add: unknown object\n" ) ;
	const rtx_optics o = optics_of( thing ) ;
	uint32_t id = 0 ;
	RTX_CHECK( ctx_, rtx_thing_add( ctx_, meshes_[object], &o, &id ) ) ;
	// the reference stores the GAS's device vertex/index buffers in the caller's Thing for
	// its SBT record (optx/scene.cxx:191-192); the hit programs are inside librtx now
	thing.vces = nullptr ;
	thing.ices = nullptr ;
	things_.push_back( thing ) ;
	return static_cast<unsigned int>( id ) ;
}

bool Scene::set( unsigned int thing, const float* transform ) {
	if ( things_.size()>thing ) {
		RTX_CHECK( ctx_, rtx_thing_set_xf( ctx_, thing, transform ) ) ;
		return true ;
	} else
		return false ;
}

bool Scene::get( unsigned int thing, float* transform ) {
	if ( things_.size()>thing ) {
		RTX_CHECK( ctx_, rtx_thing_get_xf( ctx_, thing, transform ) ) ;
		return true ;
	} else
		return false ;
}

void Scene::build( OptixTraversableHandle* is_handle ) {
	RTX_CHECK( ctx_, rtx_accel_build( ctx_ ) ) ;
	built_ = true ;
	if ( is_handle )
		*is_handle = reinterpret_cast<OptixTraversableHandle>( ctx_ ) ;
}

void Scene::update( OptixTraversableHandle /*is_handle*/ ) {
	RTX_CHECK( ctx_, rtx_accel_refit( ctx_ ) ) ;
}
