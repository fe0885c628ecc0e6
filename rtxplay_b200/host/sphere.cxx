// sphere.cxx -- Sphere over librtx's tessellator (rtx_sphere_mesh restates
// optx/sphere.cxx:28-106); with -DMAIN the scene-file writer of optx/sphere.cxx:108-142
// (vertices with printf's %f, six decimals).
#include <cstdio>
#include <iostream>
#include <stdexcept>

#include "../../include/rtx.h"
#include "sphere.h"

Sphere::Sphere( const float radius, const unsigned int ndiv ) {
	uint32_t nv = 0, nt = 0 ;
	if ( rtx_sphere_mesh( radius, ndiv, nullptr, &nv, nullptr, &nt ) != 0 )
		throw std::runtime_error( "Sphere: unsupported subdivision count" ) ;
	vces_.resize( nv ) ;
	ices_.resize( nt ) ;
	if ( rtx_sphere_mesh( radius, ndiv, &vces_[0].x, &nv, &ices_[0].x, &nt ) != 0 )
		throw std::runtime_error( "Sphere: tessellation failed" ) ;
	vces_.resize( nv ) ;
}

const Mesh Sphere::mesh() {
	return Mesh( vces_.data(), static_cast<unsigned int>( vces_.size() ), ices_.data(), static_cast<unsigned int>( ices_.size() ) ) ;
}

#ifdef MAIN

int main( const int argc, const char** argv ) {
	float radius = 1.f ;
	unsigned int ndiv = 6 ;
	if ( argc>1 ) sscanf( argv[1], "%f", &radius ) ;
	if ( argc>2 ) sscanf( argv[2], "%u", &ndiv ) ;
	Sphere sphere( radius, ndiv ) ;

	float3* vces ; unsigned int nv ; uint3* ices ; unsigned int nt ;
	std::tie( vces, nv, ices, nt ) = sphere.mesh() ;

	printf( "# sphere approximation by `inflated' tetrahedron\n" ) ;
	printf( "# obtained by %u-fold triangular area subdivision\n", ndiv ) ;
	printf( "o sphere_%u\n", ndiv ) ;
	for ( unsigned int v = 0 ; v<nv ; v++ )
		printf( "v %f %f %f\n", vces[v].x, vces[v].y, vces[v].z ) ;
	printf( "# %u vertices\n", nv ) ;
	for ( unsigned int i = 0 ; i<nt ; i++ )
		printf( "f %d %d %d\n", ices[i].x+1, ices[i].y+1, ices[i].z+1 ) ;
	printf( "# %u triangles\n\n", nt ) ;

	return 0 ;
}

#endif // MAIN
