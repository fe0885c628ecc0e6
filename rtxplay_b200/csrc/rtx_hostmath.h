// rtx_hostmath.h -- host-side arithmetic that is part of the float contract.
#pragma once

namespace rtx {

// world->object inverse of a row-major 3x4 affine map: cofactors in double (the float
// contract fixes this expression; the oracle states the same one independently)
inline void affine_inverse( const float xf[12], double inv[12] ) {
	const double a = xf[0], b = xf[1], c = xf[2],  tx = xf[3] ;
	const double d = xf[4], e = xf[5], f = xf[6],  ty = xf[7] ;
	const double g = xf[8], h = xf[9], i = xf[10], tz = xf[11] ;
	const double A =  ( e*i-f*h ), B = -( d*i-f*g ), C =  ( d*h-e*g ) ;
	const double D = -( b*i-c*h ), E =  ( a*i-c*g ), F = -( a*h-b*g ) ;
	const double G =  ( b*f-c*e ), H = -( a*f-c*d ), I =  ( a*e-b*d ) ;
	const double det = a*A+b*B+c*C ;
	const double s = 1./det ;
	inv[0] = s*A ; inv[1] = s*D ; inv[2]  = s*G ;
	inv[4] = s*B ; inv[5] = s*E ; inv[6]  = s*H ;
	inv[8] = s*C ; inv[9] = s*F ; inv[10] = s*I ;
	inv[3]  = -( inv[0]*tx+inv[1]*ty+inv[2]*tz ) ;
	inv[7]  = -( inv[4]*tx+inv[5]*ty+inv[6]*tz ) ;
	inv[11] = -( inv[8]*tx+inv[9]*ty+inv[10]*tz ) ;
}


} // namespace rtx

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace rtx {

// Object-space bounding sphere of a mesh: centre of its vertex box, radius to the farthest
// vertex (double).  out = cx, cy, cz, r.
inline void mesh_bsphere( const float* xyz, uint32_t nv, double out[4] ) {
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 } ;
	for ( uint32_t v = 0 ; v<nv ; v++ )
		for ( int a = 0 ; a<3 ; a++ ) { lo[a] = std::min( lo[a], double( xyz[3*size_t( v )+a] ) ) ; hi[a] = std::max( hi[a], double( xyz[3*size_t( v )+a] ) ) ; }
	for ( int a = 0 ; a<3 ; a++ ) out[a] = .5*( lo[a]+hi[a] ) ;
	double r2 = 0. ;
	for ( uint32_t v = 0 ; v<nv ; v++ ) {
		double d2 = 0. ;
		for ( int a = 0 ; a<3 ; a++ ) { const double d = double( xyz[3*size_t( v )+a] )-out[a] ; d2 += d*d ; }
		r2 = std::max( r2, d2 ) ;
	}
	out[3] = std::sqrt( r2 ) ;
}

// World-space bounding sphere of an instance (float, padded): centre through the transform,
// radius times the largest singular value of its 3x3 part (power iteration on M^T M).
// The padding (2^-12 relative) covers the float evaluation of the ray test below, including
// the cancellation in |o-c|^2 - r^2 for huge spheres.
inline void world_bsphere( const float xf[12], const double bs[4], float out[4] ) {
	const double m[9] = { xf[0], xf[1], xf[2], xf[4], xf[5], xf[6], xf[8], xf[9], xf[10] } ;
	double g[9] ;
	for ( int i = 0 ; i<3 ; i++ ) for ( int j = 0 ; j<3 ; j++ ) g[3*i+j] = m[i]*m[j]+m[3+i]*m[3+j]+m[6+i]*m[6+j] ;
	double v[3] = { 1., .7, .4 }, lam = 0. ;
	if ( g[1] == 0. && g[2] == 0. && g[5] == 0. )
		lam = std::max( g[0], std::max( g[4], g[8] ) ) ;   // orthogonal columns (scale + translation, the usual case): exact
	else
		for ( int it = 0 ; it<64 ; it++ ) {
			const double w[3] = { g[0]*v[0]+g[1]*v[1]+g[2]*v[2], g[3]*v[0]+g[4]*v[1]+g[5]*v[2], g[6]*v[0]+g[7]*v[1]+g[8]*v[2] } ;
			const double was = lam ;
			lam = std::sqrt( w[0]*w[0]+w[1]*w[1]+w[2]*w[2] ) ;
			if ( lam == 0. ) break ;
			v[0] = w[0]/lam ; v[1] = w[1]/lam ; v[2] = w[2]/lam ;
			if ( it>=4 && std::fabs( lam-was )<=1e-9*lam ) break ;   // (the 1.0001 below covers far more)
		}
	// power iteration approaches the largest eigenvalue from below: bound it by the trace too
	const double sig = std::sqrt( std::min( std::max( lam*1.0001, 0. ), g[0]+g[4]+g[8] ) ) ;
	const double c[3] = { bs[0]*xf[0]+bs[1]*xf[1]+bs[2]*xf[2]+xf[3], bs[0]*xf[4]+bs[1]*xf[5]+bs[2]*xf[6]+xf[7], bs[0]*xf[8]+bs[1]*xf[9]+bs[2]*xf[10]+xf[11] } ;
	out[0] = float( c[0] ) ; out[1] = float( c[1] ) ; out[2] = float( c[2] ) ;
	const double cmax = std::max( std::fabs( c[0] ), std::max( std::fabs( c[1] ), std::fabs( c[2] ) ) ) ;
	out[3] = float( ( bs[3]*sig+cmax*1e-6 )*( 1.+1./4096. )+1e-30 ) ;
}

} // namespace rtx


namespace rtx {

// Certifies that an indexed triangle mesh is the boundary of a convex body: closed,
// consistently oriented, connected, Euler characteristic 2, and locally convex at every
// edge (van Heijenoort: a closed connected surface that is locally convex everywhere bounds
// a convex body).  Returns +1 when the faces wind counter-clockwise seen from outside
// (cross(e1,e2) points outward), -1 when clockwise, 0 when the mesh is not certified.
// A ray that leaves such a thing outward cannot hit it again, which lets the traversal
// skip the thing the ray just left (rtx_pool.cuh step_thing).
inline int mesh_convexity( const float* xyz, uint32_t nv, const uint32_t* idx, uint32_t nt ) {
	if ( nt<4 || nv<4 )
		return 0 ;
	struct Half { uint32_t a, b, face, opp ; bool fwd ; } ;
	std::vector<Half> h ;
	h.reserve( 3*size_t( nt ) ) ;
	for ( uint32_t f = 0 ; f<nt ; f++ ) {
		const uint32_t v[3] = { idx[3*size_t( f )], idx[3*size_t( f )+1], idx[3*size_t( f )+2] } ;
		if ( v[0] == v[1] || v[1] == v[2] || v[2] == v[0] )
			return 0 ;
		for ( int e = 0 ; e<3 ; e++ ) {
			const uint32_t a = v[e], b = v[( e+1 )%3], o = v[( e+2 )%3] ;
			h.push_back( { std::min( a, b ), std::max( a, b ), f, o, a<b } ) ;
		}
	}
	std::sort( h.begin(), h.end(), []( const Half& x, const Half& y ) { return x.a != y.a ? x.a<y.a : x.b<y.b ; } ) ;
	if ( h.size()%2 )
		return 0 ;
	auto P = [&]( uint32_t v, double p[3] ) { p[0] = xyz[3*size_t( v )] ; p[1] = xyz[3*size_t( v )+1] ; p[2] = xyz[3*size_t( v )+2] ; } ;
	auto normal = [&]( uint32_t f, double n[3], double a[3] ) {
		double b[3], c[3] ;
		P( idx[3*size_t( f )], a ) ; P( idx[3*size_t( f )+1], b ) ; P( idx[3*size_t( f )+2], c ) ;
		const double u[3] = { b[0]-a[0], b[1]-a[1], b[2]-a[2] }, w[3] = { c[0]-a[0], c[1]-a[1], c[2]-a[2] } ;
		n[0] = u[1]*w[2]-u[2]*w[1] ; n[1] = u[2]*w[0]-u[0]*w[2] ; n[2] = u[0]*w[1]-u[1]*w[0] ;
	} ;
	// edges: exactly two half-edges of opposite direction; local convexity, both signs tracked
	std::vector<uint32_t> parent( nt ) ;
	for ( uint32_t f = 0 ; f<nt ; f++ ) parent[f] = f ;
	auto find = [&]( uint32_t x ) { while ( parent[x] != x ) { parent[x] = parent[parent[x]] ; x = parent[x] ; } return x ; } ;
	bool may_out = true, may_in = true ;   // all neighbours on/below (faces wind CCW from outside) / on/above (CW)
	for ( size_t k = 0 ; k<h.size() ; k += 2 ) {
		const Half& x = h[k] ; const Half& y = h[k+1] ;
		if ( x.a != y.a || x.b != y.b || x.fwd == y.fwd )
			return 0 ;
		if ( k+2<h.size() && h[k+2].a == x.a && h[k+2].b == x.b )
			return 0 ;   // an edge with more than two faces
		parent[find( x.face )] = find( y.face ) ;
		double n[3], a[3], w[3] ;
		normal( x.face, n, a ) ;
		P( y.opp, w ) ;
		const double nn = std::sqrt( n[0]*n[0]+n[1]*n[1]+n[2]*n[2] ) ;
		const double dw[3] = { w[0]-a[0], w[1]-a[1], w[2]-a[2] } ;
		const double dl = std::sqrt( dw[0]*dw[0]+dw[1]*dw[1]+dw[2]*dw[2] ) ;
		if ( nn == 0. )
			return 0 ;
		const double s = ( n[0]*dw[0]+n[1]*dw[1]+n[2]*dw[2] )/( nn*( dl>0. ? dl : 1. ) ) ;   // sine of the fold
		if ( s> 1e-9 ) may_out = false ;
		if ( s<-1e-9 ) may_in = false ;
	}
	if ( may_out == may_in )   // neither, or completely flat
		return 0 ;
	const uint32_t root = find( 0 ) ;
	for ( uint32_t f = 1 ; f<nt ; f++ ) if ( find( f ) != root ) return 0 ;
	// Euler characteristic of a sphere; count the vertices actually used
	std::vector<char> used( nv, 0 ) ;
	size_t nu = 0 ;
	for ( size_t k = 0 ; k<3*size_t( nt ) ; k++ ) if ( ! used[idx[k]] ) { used[idx[k]] = 1 ; nu++ ; }
	if ( long( nu )-long( h.size()/2 )+long( nt ) != 2 )
		return 0 ;
	return may_out ? 1 : -1 ;
}

} // namespace rtx
