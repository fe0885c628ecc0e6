// rtx_host.cpp -- host-only helpers of the C ABI (no device work): the float camera of
// optx/camera.h:30-48 and the sphere tessellator of optx/sphere.cxx:28-106.  They live
// in librtx.so so that the C++ shims, the rtwo driver and the Python tests all obtain the
// same bits.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/rtx.h"

namespace {

struct P3 { float x, y, z ; } ;

inline P3 add( const P3& a, const P3& b ) { return { a.x+b.x, a.y+b.y, a.z+b.z } ; }
inline P3 sub( const P3& a, const P3& b ) { return { a.x-b.x, a.y-b.y, a.z-b.z } ; }
inline P3 mul( float t, const P3& a )     { return { t*a.x, t*a.y, t*a.z } ; }
inline float dot( const P3& a, const P3& b ) { return a.x*b.x+a.y*b.y+a.z*b.z ; }
inline P3 cross( const P3& a, const P3& b ) { return { a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x } ; }
// optx/v.h:46,51: unitV(v) = 1.f/len(v)*v
inline P3 unit( const P3& a ) { return mul( 1.f/sqrtf( dot( a, a ) ), a ) ; }

// 4-way split of a spherical triangle, midpoints pushed back onto the unit sphere
// (optx/sphere.cxx:51-66); leaves are emitted as a triangle soup scaled by the radius
void subdivide( const P3& a, const P3& b, const P3& c, unsigned level, float radius, std::vector<P3>& soup ) {
	if ( level == 0 ) {
		soup.push_back( mul( radius, a ) ) ;
		soup.push_back( mul( radius, b ) ) ;
		soup.push_back( mul( radius, c ) ) ;
		return ;
	}
	const P3 ab = unit( mul( .5f, add( a, b ) ) ) ;
	const P3 bc = unit( mul( .5f, add( b, c ) ) ) ;
	const P3 ca = unit( mul( .5f, add( c, a ) ) ) ;
	subdivide(  a, ab, ca, level-1, radius, soup ) ;
	subdivide( ab,  b, bc, level-1, radius, soup ) ;
	subdivide( ca, bc,  c, level-1, radius, soup ) ;
	subdivide( ab, bc, ca, level-1, radius, soup ) ;
}

struct Key {
	uint32_t a, b, c ;
	bool operator == ( const Key& o ) const { return a == o.a && b == o.b && c == o.c ; }
} ;
struct KeyHash {
	size_t operator () ( const Key& k ) const {
		uint64_t h = 0x9E3779B97F4A7C15ull*k.a ;
		h ^= ( h>>29 ) ; h += 0xBF58476D1CE4E5B9ull*k.b ;
		h ^= ( h>>31 ) ; h += 0x94D049BB133111EBull*k.c ;
		return size_t( h^( h>>32 ) ) ;
	}
} ;
inline uint32_t fbits( float f ) { if ( f == 0.f ) f = 0.f ; uint32_t u ; memcpy( &u, &f, 4 ) ; return u ; }

} // namespace

extern "C" {

void rtx_camera_set( rtx_camera* cam, const float eye[3], const float pat[3], const float vup[3], float fov, float aspratio, float aperture, float fostance ) {
	const P3 e = { eye[0], eye[1], eye[2] }, p = { pat[0], pat[1], pat[2] }, up = { vup[0], vup[1], vup[2] } ;
	const float kPi = 3.14159265358979323846f ;          // optx/util.h:22
	const P3 w = unit( sub( e, p ) ) ;
	const P3 u = unit( cross( up, w ) ) ;
	const P3 v = cross( w, u ) ;
	const float h  = 2.f*tanf( .5f*( fov*kPi/180.f ) ) ; // focus plane height
	const float wd = h*aspratio ;
	const P3 hvec = mul( fostance*h/2.f, v ) ;
	const P3 wvec = mul( fostance*wd/2.f, u ) ;
	const P3 dvec = mul( fostance, w ) ;
	const P3 src[6] = { e, u, v, hvec, wvec, dvec } ;
	float* dst[6] = { cam->eye, cam->u, cam->v, cam->hvec, cam->wvec, cam->dvec } ;
	for ( int k = 0 ; k<6 ; k++ ) { dst[k][0] = src[k].x ; dst[k][1] = src[k].y ; dst[k][2] = src[k].z ; }
	cam->aperture = aperture ;
}

int rtx_sphere_mesh( float radius, uint32_t ndiv, float* xyz, uint32_t* n_vertices, uint32_t* idx, uint32_t* n_triangles ) {
	if ( ndiv>12 )
		return 1 ;
	const uint32_t nt = 4u<<( 2*ndiv ) ;        // 4 * 4^n
	const uint32_t nv = ( 2u<<( 2*ndiv ) )+2u ; // 2 * 4^n + 2
	if ( ! xyz || ! idx ) {
		if ( n_vertices ) *n_vertices = nv ;
		if ( n_triangles ) *n_triangles = nt ;
		return 0 ;
	}
	// optx/sphere.cxx:28-49: unit tetrahedron, midsphere radius m
	const float m = .57735026919f ;
	const P3 v0 = {  m,  m,  m }, v1 = {  m, -m, -m }, v2 = { -m, -m,  m }, v3 = { -m,  m, -m } ;
	std::vector<P3> soup ;
	soup.reserve( 3*size_t( nt ) ) ;
	subdivide( v0, v1, v2, ndiv, radius, soup ) ;
	subdivide( v0, v2, v3, ndiv, radius, soup ) ;
	subdivide( v0, v3, v1, ndiv, radius, soup ) ;
	subdivide( v3, v2, v1, ndiv, radius, soup ) ;
	// optx/sphere.cxx:68-106: soup -> indexed, exact compare, first-appearance order
	std::unordered_map<Key, uint32_t, KeyHash> seen ;
	seen.reserve( nv*2 ) ;
	uint32_t next = 0 ;
	for ( size_t k = 0 ; k<soup.size() ; k++ ) {
		const Key key = { fbits( soup[k].x ), fbits( soup[k].y ), fbits( soup[k].z ) } ;
		auto it = seen.find( key ) ;
		uint32_t id ;
		if ( it == seen.end() ) {
			id = next++ ;
			seen.emplace( key, id ) ;
			if ( id<nv ) { xyz[3*size_t( id )] = soup[k].x ; xyz[3*size_t( id )+1] = soup[k].y ; xyz[3*size_t( id )+2] = soup[k].z ; }
		} else
			id = it->second ;
		idx[k] = id ;
	}
	if ( n_vertices ) *n_vertices = next ;
	if ( n_triangles ) *n_triangles = uint32_t( soup.size()/3 ) ;
	return next<=nv ? 0 : 2 ;   // more unique vertices than the closed form: rounding split a shared vertex
}

} // extern "C"
