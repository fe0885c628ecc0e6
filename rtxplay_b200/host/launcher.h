// launcher.h -- launch boundary of the reference (optx/launcher.h:17-31): owns the frame
// buffers behind cg::lp_general and renders one frame per ignite().
#ifndef LAUNCHER_H
#define LAUNCHER_H

#include "rtwo.h"

class Launcher {
	public:
		// reference signature; pipeline and SBT have no counterpart (programs and records
		// live inside librtx), the context is taken from the scene handle at ignite()
		Launcher( const OptixPipeline& pipeline, const OptixShaderBindingTable& sbt ) ;
		explicit Launcher( const OptixDeviceContext& optx_context ) ;
		~Launcher() noexcept ( false ) ;

		void resize( const unsigned int w, const unsigned int h ) ;
		void ignite( const CUstream& cuda_stream, bool once = false ) ;

		// additive: stream key and sample window of the frame (defaults: 4711, all samples)
		void seed( unsigned long long seed ) { seed_ = seed ; }
		// additive: fill lp_general.normals / albedos (the reference always does; here on request)
		void guides( bool on ) { guides_ = on ; }
		// additive: which of the reference's programs the paths follow where they disagree
		// (include/rtx.h RTX_VARIANT_*): rtow.cxx by default, the iterative OptiX programs, or the
		// recursive ones the reference selects at build time with -DRECURSIVE
		void variant( unsigned int v ) { variant_ = v ; }

	private:
		rtx_ctx*           ctx_ ;
		unsigned long long seed_ ;
		bool               guides_ = false ;
		unsigned int       variant_ = 0 ;

		void bind() ;
} ;

#endif // LAUNCHER_H
