// launcher.cxx -- Launcher over the C ABI (optx/launcher.cxx:24-85).  ignite() re-reads
// the global launch parameters every frame, renders, and blocks until the frame is
// complete; device failures throw.
#include <cstring>

#include "launcher.h"
#include "util.h"

namespace cg {
	extern LpGeneral lp_general ;
	extern rtx_ctx*  rtx_context ;  // postproc.cxx: the context pp_none / pp_sRGB run on
}
using namespace cg ;

Launcher::Launcher( const OptixPipeline& /*pipeline*/, const OptixShaderBindingTable& /*sbt*/ ) : ctx_( nullptr ), seed_( 4711 ) {
	// the context arrives with the scene handle (Scene::build stores it in lp_general.is_handle)
	ctx_ = reinterpret_cast<rtx_ctx*>( lp_general.is_handle ) ;
	if ( ! ctx_ )
		throw std::runtime_error( "Launcher: build the scene before creating the launcher\n" ) ;
	rtx_context = ctx_ ;
	resize( lp_general.image_w, lp_general.image_h ) ;
}

Launcher::Launcher( const OptixDeviceContext& optx_context ) : ctx_( optx_context ), seed_( 4711 ) {
	rtx_context = ctx_ ;
	resize( lp_general.image_w, lp_general.image_h ) ;
}

Launcher::~Launcher() noexcept ( false ) {
	// frame buffers belong to the context
}

void Launcher::bind() {
	void* p = nullptr ;
	RTX_CHECK( ctx_, rtx_device_ptr( ctx_, RTX_BUF_RAWRGB,  &p, nullptr ) ) ; lp_general.rawRGB  = static_cast<float3*>( p ) ;
	RTX_CHECK( ctx_, rtx_device_ptr( ctx_, RTX_BUF_RPP,     &p, nullptr ) ) ; lp_general.rpp     = static_cast<unsigned int*>( p ) ;
	RTX_CHECK( ctx_, rtx_device_ptr( ctx_, RTX_BUF_NORMALS, &p, nullptr ) ) ; lp_general.normals = static_cast<float3*>( p ) ;
	RTX_CHECK( ctx_, rtx_device_ptr( ctx_, RTX_BUF_ALBEDOS, &p, nullptr ) ) ; lp_general.albedos = static_cast<float3*>( p ) ;
	RTX_CHECK( ctx_, rtx_device_ptr( ctx_, RTX_BUF_PICK_ID, &p, nullptr ) ) ; lp_general.pick_id = static_cast<unsigned int*>( p ) ;
}

void Launcher::resize( const unsigned int w, const unsigned int h ) {
	RTX_CHECK( ctx_, rtx_resize( ctx_, w, h ) ) ;
	bind() ;
}

void Launcher::ignite( const CUstream& /*cuda_stream*/, bool once ) {
	rtx_params p ;
	memset( &p, 0, sizeof( p ) ) ;
	p.image_w = lp_general.image_w ; p.image_h = lp_general.image_h ;
	p.spp = lp_general.spp ; p.depth = lp_general.depth ;
	p.camera = lp_general.camera.derived() ;
	p.seed = seed_ ; p.sample0 = 0 ; p.sample_stride = 1 ; p.accumulate = 0 ; p.guides = guides_ ? 1u : 0u ; p.variant = variant_ ;

	if ( once || lp_general.picker ) {
		// scene editing: one primary ray, the thing id lands in *pick_id (device)
		uint32_t id = 0 ;
		RTX_CHECK( ctx_, rtx_pick( ctx_, &p, lp_general.pick_x, lp_general.pick_y, &id ) ) ;
		return ;
	}
	RTX_CHECK( ctx_, rtx_render( ctx_, &p ) ) ;
}
