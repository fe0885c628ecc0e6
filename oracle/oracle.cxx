// oracle.cxx -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the algorithm of RTXplay's path-tracing hot path, used by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER.
// Nothing in rtxplay_b200/ (the product) may include, link or call this file.
//
// What it restates (file:line under /root/reference):
//   rtow.cxx:34-49   trace()            -> Tracer<R,Rng>::path()
//   rtow.cxx:51-80   scene()            -> orc_rtow_scene()
//   rtow.cxx:82-122  main() pixel loop  -> render_rows()
//   rtow.cxx:6-21    sRGB() (gamma 2)   -> orc_ppm_rtow()
//   camera.h:10-31   Camera::set/ray    -> orc_camera_set_f64(), Tracer::camray()
//   sphere.h:20-48   Sphere::hit        -> hit_sphere()
//   things.h:22-36   Things::hit        -> closest()
//   optics.h:11-75   Diffuse/Reflect/Refract::spray, schlick -> scatter()
//   v.h:42-62        V ops, rndVin1sphere, rndVon1sphere, rndVin1disk, reflect, refract
//   util.h:7-15      kAcne0, kNear0, rnd()
// Triangle mode additionally follows
//   optx/optics_i.cu:40-82 (indexed vertices -> instance transform -> barycentric hit
//   point -> flat normal flipped against the ray) and optx/optics_i.cu:259-267 (the
//   dielectric inside/outside decision from dot(d, hit-centre)); termination and
//   material semantics stay those of rtow.cxx.  optx/postproc.cu:2-16 is restated in
//   orc_srgb8().  optx/camera.h:30-48 (float camera) in orc_camera_set_f32().
//
// Three instantiations of the SAME template code:
//   <double, RngLibc>  reproduces the unmodified reference binary bit for bit
//                      (the pin: md5 of its 1280x720x10spp PPM, see tests/golden/).
//   <double, RngPcg>   the semantic reference at any resolution/spp ("converged
//                      radiance" target), consuming the product's random streams.
//   <float,  RngPcg>   the float mirror: every operation individually rounded, in
//                      the order written here (build with -ffp-contract=off), which
//                      is the arithmetic contract the CUDA kernels implement.
//
// Argument-evaluation order: rtow.cxx has several calls whose operand order the C++
// standard leaves unspecified (rtow.cxx:59, :62; v.h:30-31, :59).  g++ on x86-64
// evaluates them right to left; this file sequences the draws explicitly in that
// order, and the md5 pin proves it.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <limits>
#include <type_traits>

namespace {

// ---------------------------------------------------------------- vectors (v.h:9-50)
template <class R> struct V3 { R x, y, z ; } ;

template <class R> inline V3<R> mk( R x, R y, R z ) { V3<R> v = { x, y, z } ; return v ; }
template <class R> inline V3<R> operator + ( const V3<R>& u, const V3<R>& v ) { return mk<R>( u.x+v.x, u.y+v.y, u.z+v.z ) ; }
template <class R> inline V3<R> operator - ( const V3<R>& u, const V3<R>& v ) { return mk<R>( u.x-v.x, u.y-v.y, u.z-v.z ) ; }
template <class R> inline V3<R> operator - ( const V3<R>& u )                 { return mk<R>( -u.x, -u.y, -u.z ) ; }
template <class R> inline V3<R> operator * ( const V3<R>& u, const V3<R>& v ) { return mk<R>( u.x*v.x, u.y*v.y, u.z*v.z ) ; }
template <class R> inline V3<R> operator * ( R t, const V3<R>& v )            { return mk<R>( t*v.x, t*v.y, t*v.z ) ; }
// v.h:46: division is multiplication by the reciprocal
template <class R> inline V3<R> operator / ( const V3<R>& v, R t )            { return ( R( 1 )/t )*v ; }
template <class R> inline R dot( const V3<R>& u, const V3<R>& v )             { return u.x*v.x+u.y*v.y+u.z*v.z ; }
template <class R> inline R len( const V3<R>& v )                             { return std::sqrt( dot( v, v ) ) ; }
template <class R> inline V3<R> cross( const V3<R>& u, const V3<R>& v )       { return mk<R>( u.y*v.z-u.z*v.y, u.z*v.x-u.x*v.z, u.x*v.y-u.y*v.x ) ; }
template <class R> inline V3<R> unitV( const V3<R>& v )                       { return v/len( v ) ; }

// ---------------------------------------------------------------- random sources
// util.h:12  rnd() = rand()/(RAND_MAX+1.)  -- process-global libc stream
struct RngLibc {
	static std::atomic<uint64_t> calls ;
	void seed( uint64_t, uint32_t, uint32_t ) {}
	double next() { calls.fetch_add( 1, std::memory_order_relaxed ) ; return rand()/( RAND_MAX+1. ) ; }
} ;
std::atomic<uint64_t> RngLibc::calls( 0 ) ;

// The product's counter-based generator: the stream of path (pixel, sample) is a pure
// function of (seed, pixel, sample) -- a PCG32 (XSH-RR 64/32) whose state is keyed by
// a splitmix64 finaliser.  A draw is the top 24 bits scaled by 2^-24: exactly
// representable in float and double alike, so both precisions see the same numbers.
struct RngPcg {
	uint64_t state ;
	static uint64_t mix( uint64_t z ) {
		z = ( z^( z>>30 ) )*0xBF58476D1CE4E5B9ull ;
		z = ( z^( z>>27 ) )*0x94D049BB133111EBull ;
		return z^( z>>31 ) ;
	}
	void seed( uint64_t seed, uint32_t pixel, uint32_t sample ) {
		state = mix( ( ( uint64_t( pixel )<<32 )|uint64_t( sample ) )^( seed*0x9E3779B97F4A7C15ull ) ) ;
	}
	uint32_t bits() {
		const uint64_t old = state ;
		state = old*6364136223846793005ull+1442695040888963407ull ;
		const uint32_t xs  = uint32_t( ( ( old>>18 )^old )>>27 ) ;
		const uint32_t rot = uint32_t( old>>59 ) ;
		return ( xs>>rot )|( xs<<( ( 32-rot )&31 ) ) ;
	}
	double next() { return double( bits()>>8 )*( 1./16777216. ) ; }
} ;

// ---------------------------------------------------------------- scene tables
// One row per Thing, shared layout with the product's host API (see tests/):
//  [0] geometry kind: 0 analytic sphere, 1 triangle mesh instance
//  [1] mesh id (kind 1)
//  [2..13] row-major 3x4 object->world transform (optx/scene.cxx:183-188);
//          analytic sphere: centre = translation column, radius = [2] (scale)
//  [14] optics type 0 diffuse / 1 reflect / 2 refract (optx/thing.h:29-34)
//  [15..17] albedo  [18] fuzz  [19] refraction index
enum { TH_KIND = 0, TH_MESH = 1, TH_XF = 2, TH_TYPE = 14, TH_ALB = 15, TH_FUZZ = 18, TH_INDEX = 19, TH_STRIDE = 20 } ;
// camera block: eye, u, v, hvec, wvec, dvec (3 each), aperture
enum { CAM_EYE = 0, CAM_U = 3, CAM_V = 6, CAM_HVEC = 9, CAM_WVEC = 12, CAM_DVEC = 15, CAM_APERTURE = 18, CAM_STRIDE = 19 } ;

struct MeshRef {
	const float*    vces ; uint32_t nv ;
	const uint32_t* ices ; uint32_t nt ;
} ;

template <class R> struct ThingT {
	int   kind, mesh, type ;
	V3<R> center ; R radius ;     // analytic
	R     xf[12] ;                // object->world
	R     inv[12] ;               // world->object (mesh instances)
	V3<R> albedo ; R fuzz, index ;
	// double copies of the geometry, used by the float contract where single
	// precision is not enough (see "float contract" below)
	double xf_d[12], inv_d[12] ;
} ;

template <class R> struct TriT { V3<R> v0, e1, e2 ; } ;

template <class R> struct Hit {
	R     t ;
	V3<R> p, normal ;
	bool  facing ;
	int   thing ;
	int   prim ;  // -1 for analytic
} ;

// world->object inverse of a 3x4 affine map, cofactors in double, written order is
// part of the float-mirror contract (the product computes the same expression).
inline void invert_affine( const double m[12], double inv[12] ) {
	const double a = m[0], b = m[1], c = m[2],  tx = m[3] ;
	const double d = m[4], e = m[5], f = m[6],  ty = m[7] ;
	const double g = m[8], h = m[9], i = m[10], tz = m[11] ;
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

template <class R> struct SceneT {
	std::vector<ThingT<R>>            things ;
	std::vector<std::vector<TriT<R>>> tris ;   // per mesh, object space
	std::vector<MeshRef>              meshes ;

	void load( const double* tab, int n, const MeshRef* m, int nm ) {
		things.resize( n ) ;
		for ( int k = 0 ; k<n ; k++ ) {
			const double* row = tab+size_t( k )*TH_STRIDE ;
			ThingT<R>& th = things[k] ;
			th.kind = int( row[TH_KIND] ) ; th.mesh = int( row[TH_MESH] ) ; th.type = int( row[TH_TYPE] ) ;
			double inv[12] ;
			invert_affine( row+TH_XF, inv ) ;
			for ( int j = 0 ; j<12 ; j++ ) { th.xf[j] = R( row[TH_XF+j] ) ; th.inv[j] = R( inv[j] ) ; th.xf_d[j] = row[TH_XF+j] ; th.inv_d[j] = inv[j] ; }
			th.center = mk<R>( R( row[TH_XF+3] ), R( row[TH_XF+7] ), R( row[TH_XF+11] ) ) ;
			th.radius = R( row[TH_XF+0] ) ;
			th.albedo = mk<R>( R( row[TH_ALB] ), R( row[TH_ALB+1] ), R( row[TH_ALB+2] ) ) ;
			th.fuzz   = R( row[TH_FUZZ] ) ; th.index = R( row[TH_INDEX] ) ;
		}
		meshes.assign( m, m+nm ) ;
		tris.resize( nm ) ;
		for ( int q = 0 ; q<nm ; q++ ) {
			tris[q].resize( m[q].nt ) ;
			for ( uint32_t f = 0 ; f<m[q].nt ; f++ ) {
				const uint32_t* ix = m[q].ices+3*size_t( f ) ;
				// edges are formed in float first (that is what the product stores),
				// then widened: both precisions intersect the same stored triangle
				const float* a = m[q].vces+3*size_t( ix[0] ) ;
				const float* b = m[q].vces+3*size_t( ix[1] ) ;
				const float* c = m[q].vces+3*size_t( ix[2] ) ;
				TriT<R>& T = tris[q][f] ;
				T.v0 = mk<R>( R( a[0] ), R( a[1] ), R( a[2] ) ) ;
				T.e1 = mk<R>( R( float( b[0]-a[0] ) ), R( float( b[1]-a[1] ) ), R( float( b[2]-a[2] ) ) ) ;
				T.e2 = mk<R>( R( float( c[0]-a[0] ) ), R( float( c[1]-a[1] ) ), R( float( c[2]-a[2] ) ) ) ;
			}
		}
	}
} ;

template <class R> struct Consts ;
template <> struct Consts<double> {
	static double acne()  { return 1e-3f ; }   // util.h:9  (float literal widened)
	static double near0() { return 1e-8f ; }   // util.h:8
	static double tmax()  { return std::numeric_limits<double>::infinity() ; }
} ;
template <> struct Consts<float> {
	static float acne()  { return 1e-3f ; }
	static float near0() { return 1e-8f ; }
	static float tmax()  { return std::numeric_limits<float>::infinity() ; }
} ;

// ---------------------------------------------------------------- the tracer
template <class R, class Rng> struct Tracer {
	const SceneT<R>* scene ;
	Rng   rng ;
	V3<R> eye, cu, cv, hvec, wvec, dvec ; R aperture ;

	R rnd() { return R( rng.next() ) ; }
	R rnd( R min, R max ) { return min+rnd()*( max-min ) ; }   // util.h:13

	// v.h:30-31 V::rnd(min,max) -- g++ draws z, y, x
	V3<R> rndV( R min, R max ) { const R z = rnd( min, max ) ; const R y = rnd( min, max ) ; const R x = rnd( min, max ) ; return mk<R>( x, y, z ) ; }
	// v.h:53
	V3<R> rndVin1sphere() { while ( true ) { const V3<R> v = rndV( R( -1 ), R( 1 ) ) ; if ( R( 1 )>dot( v, v ) ) return v ; } }
	// v.h:55
	V3<R> rndVon1sphere() { return unitV( rndVin1sphere() ) ; }
	// v.h:59 -- g++ draws y, then x
	V3<R> rndVin1disk() { while ( true ) { const R y = rnd( R( -1 ), R( 1 ) ) ; const R x = rnd( R( -1 ), R( 1 ) ) ; const V3<R> v = mk<R>( x, y, R( 0 ) ) ; if ( R( 1 )>dot( v, v ) ) return v ; } }

	void setcam( const double* cam ) {
		eye  = mk<R>( R( cam[CAM_EYE] ),  R( cam[CAM_EYE+1] ),  R( cam[CAM_EYE+2] ) ) ;
		cu   = mk<R>( R( cam[CAM_U] ),    R( cam[CAM_U+1] ),    R( cam[CAM_U+2] ) ) ;
		cv   = mk<R>( R( cam[CAM_V] ),    R( cam[CAM_V+1] ),    R( cam[CAM_V+2] ) ) ;
		hvec = mk<R>( R( cam[CAM_HVEC] ), R( cam[CAM_HVEC+1] ), R( cam[CAM_HVEC+2] ) ) ;
		wvec = mk<R>( R( cam[CAM_WVEC] ), R( cam[CAM_WVEC+1] ), R( cam[CAM_WVEC+2] ) ) ;
		dvec = mk<R>( R( cam[CAM_DVEC] ), R( cam[CAM_DVEC+1] ), R( cam[CAM_DVEC+2] ) ) ;
		aperture = R( cam[CAM_APERTURE] ) ;
	}

	// camera.h:25-31
	void camray( R s, R t, V3<R>& ori, V3<R>& dir ) {
		const V3<R> r = ( aperture/R( 2 ) )*rndVin1disk() ;
		const V3<R> o = r.x*cu+r.y*cv ;
		ori = eye+o ;
		dir = s*wvec+t*hvec-dvec-o ;
	}

	// ------------------------------------------------------------------------------
	// FLOAT CONTRACT.  Colour, random numbers, camera, scattering: float, one rounding
	// per operation.  Geometry is where single precision visibly breaks the image (a
	// radius-1000 ground sphere: float cannot place a point on it better than 6e-5,
	// which turns ~0.2 % of grazing bounces into false self-hits that then rattle
	// around INSIDE the sphere for 50 segments).  So the contract evaluates
	//   * the analytic sphere roots with the reference's own double formula
	//     (sphere.h:21-38 verbatim, on the widened float ray),
	//   * the world->object ray origin in double, carried as a float pair (hi, lo);
	//     Moeller-Trumbore itself stays float and forms s = (hi - v0) + lo,
	//   * the shading frame of a hit (hit point, normal, facing) in double, rounded to
	//     float once,
	// and compares/accumulates t as a float.  Everything else is plain float.
	// ------------------------------------------------------------------------------

	// sphere.h:21-38 verbatim: the smallest root not below tmin (things.h:27-33 then
	// keeps it when it is not above the best t so far; t == best is accepted, so of two
	// things at exactly equal t the later-listed wins).
	template <class Q> static bool sphere_root( const V3<Q>& center, Q radius, const V3<Q>& ori, const V3<Q>& dir, Q tmin, Q& t ) {
		const V3<Q> o = ori-center ;
		const Q a = dot( dir, dir ) ;
		const Q b = dot( dir, o ) ;
		const Q c = dot( o, o )-radius*radius ;
		const Q discriminant = b*b-a*c ;
		if ( Q( 0 )>discriminant )
			return false ;
		const Q x = std::sqrt( discriminant ) ;
		t = ( -b-x )/a ;
		if ( tmin>t ) {
			t = ( -b+x )/a ;
			if ( tmin>t )
				return false ;
		}
		return true ;
	}
	static V3<double> wide( const V3<R>& v ) { return mk<double>( double( v.x ), double( v.y ), double( v.z ) ) ; }
	static V3<R> narrow( const V3<double>& v ) { return mk<R>( R( v.x ), R( v.y ), R( v.z ) ) ; }

	// sphere.h:20-48
	bool hit_sphere( const ThingT<R>& th, const V3<R>& ori, const V3<R>& dir, R tmin, R tmax, Hit<R>& h ) const {
		const V3<double> c = mk<double>( th.xf_d[3], th.xf_d[7], th.xf_d[11] ) ;
		const double r = th.xf_d[0] ;
		const V3<double> o = wide( ori ), d = wide( dir ) ;
		double td ;
		if ( ! sphere_root<double>( c, r, o, d, double( tmin ), td ) )
			return false ;
		const R t = R( td ) ;
		if ( t>tmax )
			return false ;
		h.t = t ;
		const V3<double> p = o+td*d ;
		const V3<double> outward = ( p-c )/r ;
		h.p = narrow( p ) ;
		h.facing = 0.>dot( d, outward ) ;
		h.normal = narrow( h.facing ? outward : -outward ) ;
		h.prim   = -1 ;
		return true ;
	}

	template <class Q> static V3<Q> xfpoint( const Q* m, const V3<Q>& p ) {
		return mk<Q>( p.x*m[0]+p.y*m[1]+p.z*m[2]+m[3], p.x*m[4]+p.y*m[5]+p.z*m[6]+m[7], p.x*m[8]+p.y*m[9]+p.z*m[10]+m[11] ) ;
	}
	template <class Q> static V3<Q> xfvec( const Q* m, const V3<Q>& p ) {
		return mk<Q>( p.x*m[0]+p.y*m[1]+p.z*m[2], p.x*m[4]+p.y*m[5]+p.z*m[6], p.x*m[8]+p.y*m[9]+p.z*m[10] ) ;
	}

	// Moeller-Trumbore on the stored (v0,e1,e2), two-sided, in the instance's object
	// space; t is the world-space ray parameter because the direction is not
	// re-normalised.  Replaces OptiX's built-in triangle test (optx/scene.cxx:63).
	// The origin arrives as hi+lo (lo = 0 in the double instantiation).
	static bool hit_tri( const TriT<R>& T, const V3<R>& ohi, const V3<R>& olo, const V3<R>& d, R tmin, R tmax, R& t, R& u, R& v ) {
		const V3<R> p = cross( d, T.e2 ) ;
		const R det = dot( T.e1, p ) ;
		if ( det == R( 0 ) )
			return false ;
		const R inv = R( 1 )/det ;
		const V3<R> s = ( ohi-T.v0 )+olo ;
		u = dot( s, p )*inv ;
		if ( ! ( u>=R( 0 ) && u<=R( 1 ) ) )       // written so that a NaN rejects
			return false ;
		const V3<R> q = cross( s, T.e1 ) ;
		v = dot( d, q )*inv ;
		if ( ! ( v>=R( 0 ) && u+v<=R( 1 ) ) )
			return false ;
		t = dot( T.e2, q )*inv ;
		if ( ! ( t>=tmin && t<=tmax ) )
			return false ;
		return true ;
	}

	// optx/optics_i.cu:40-82: indexed vertices -> world, barycentric hit point, flat
	// normal against the ray; :259-267 inside/outside from the instance centre.
	void finish_tri( const ThingT<R>& th, int prim, R u, R v, const V3<R>& dir, Hit<R>& h ) const {
		const MeshRef& m = scene->meshes[th.mesh] ;
		const uint32_t* ix = m.ices+3*size_t( prim ) ;
		const float* fa = m.vces+3*size_t( ix[0] ) ;
		const float* fb = m.vces+3*size_t( ix[1] ) ;
		const float* fc = m.vces+3*size_t( ix[2] ) ;
		const V3<double> A = xfpoint<double>( th.xf_d, mk<double>( fa[0], fa[1], fa[2] ) ) ;
		const V3<double> B = xfpoint<double>( th.xf_d, mk<double>( fb[0], fb[1], fb[2] ) ) ;
		const V3<double> C = xfpoint<double>( th.xf_d, mk<double>( fc[0], fc[1], fc[2] ) ) ;
		const R w = R( 1 )-u-v ;
		const V3<double> p = double( w )*A+double( u )*B+double( v )*C ;
		const V3<double> d = wide( dir ) ;
		V3<double> N = unitV( cross( B-A, C-A ) ) ;
		if ( dot( d, N )>0. )
			N = -N ;
		h.p = narrow( p ) ;
		h.normal = narrow( N ) ;
		h.facing = 0.>dot( d, p-mk<double>( th.xf_d[3], th.xf_d[7], th.xf_d[11] ) ) ;
	}

	// things.h:22-36 -- linear scan, shrinking tmax, later thing wins exact ties
	bool closest( const V3<R>& ori, const V3<R>& dir, R tmin, R tmax, Hit<R>& best ) const {
		bool shot = false ;
		R tact = tmax ;
		const int n = int( scene->things.size() ) ;
		for ( int k = 0 ; k<n ; k++ ) {
			const ThingT<R>& th = scene->things[k] ;
			if ( th.kind == 0 ) {
				Hit<R> h ;
				if ( hit_sphere( th, ori, dir, tmin, tact, h ) ) {
					shot = true ; tact = h.t ; best = h ; best.thing = k ;
				}
			} else {
				const V3<double> od = xfpoint<double>( th.inv_d, wide( ori ) ) ;
				const V3<R> ohi = narrow( od ) ;
				const V3<R> olo = narrow( od-wide( ohi ) ) ;
				const V3<R> d   = narrow( xfvec<double>( th.inv_d, wide( dir ) ) ) ;
				const std::vector<TriT<R>>& tr = scene->tris[th.mesh] ;
				int prim = -1 ; R bu = 0, bv = 0 ;
				for ( size_t f = 0 ; f<tr.size() ; f++ ) {
					R t, u, v ;
					if ( hit_tri( tr[f], ohi, olo, d, tmin, tact, t, u, v ) ) {
						tact = t ; prim = int( f ) ; bu = u ; bv = v ;
					}
				}
				if ( prim>=0 ) {
					shot = true ;
					best.t = tact ; best.thing = k ; best.prim = prim ;
					best.p.x = bu ; best.p.y = bv ;   // parked; finished once the scan is over
				}
			}
		}
		if ( shot && best.prim>=0 ) {
			const R u = best.p.x, v = best.p.y ;
			finish_tri( scene->things[best.thing], best.prim, u, v, dir, best ) ;
		}
		return shot ;
	}

	static V3<R> reflect( const V3<R>& v, const V3<R>& n ) { return v-( R( 2 )*dot( v, n ) )*n ; }                 // v.h:61
	static V3<R> refract( const V3<R>& v, const V3<R>& n, R ratio ) {                                              // v.h:62
		const R theta = std::fmin( dot( -v, n ), R( 1 ) ) ;
		const V3<R> perpen = ratio*( v+theta*n ) ;
		const V3<R> parall = ( -std::sqrt( std::fabs( R( 1 )-dot( perpen, perpen ) ) ) )*n ;
		return perpen+parall ;
	}
	// optics.h:74
	static R schlick( R cos_theta, R ratio ) {
		R r0 = ( R( 1 )-ratio )/( R( 1 )+ratio ) ; r0 = r0*r0 ;
		const R m = R( 1 )-cos_theta ;
		return r0+( R( 1 )-r0 )*pow5( m ) ;
	}
	// pow(x,5): libm's pow is correctly rounded for these arguments in double; the
	// float contract is the explicit product ((x*x)*(x*x))*x
	static double pow5( double m ) { return std::pow( m, 5 ) ; }
	static float  pow5( float m )  { const float m2 = m*m ; return ( m2*m2 )*m ; }

	// optics.h:15-24, 34-40, 51-68
	// variant: which of the reference's programs the path follows where they disagree:
	//   0 rtow.cxx (the parity target)   1 optx iterative programs (camera_i.cu, optics_i.cu)
	//   2 optx recursive programs (camera_r.cu, optics_r.cu)
	int variant = 0 ;
	bool scatter( const ThingT<R>& th, const V3<R>& dir, const Hit<R>& h, V3<R>& attened, V3<R>& out ) {
		if ( th.type == 0 ) {
			V3<R> d = h.normal+rndVon1sphere() ;
			// optics.h:17-18; optx/optics_i.cu:86 and optx/optics_r.cu:108 have no guard
			if ( variant == 0 && std::fabs( d.x )<Consts<R>::near0() && std::fabs( d.y )<Consts<R>::near0() && std::fabs( d.z )<Consts<R>::near0() )
				d = h.normal ;
			out = d ; attened = th.albedo ;
			return true ;
		}
		if ( th.type == 1 ) {
			const V3<R> r = reflect( unitV( dir ), h.normal ) ;
			out = r+th.fuzz*rndVin1sphere() ;
			attened = th.albedo ;
			return dot( out, h.normal )>R( 0 ) ;
		}
		const V3<R> d1V = unitV( dir ) ;
		const R cos_theta = std::fmin( dot( -d1V, h.normal ), R( 1 ) ) ;
		const R sin_theta = std::sqrt( R( 1 )-cos_theta*cos_theta ) ;
		const R ratio = h.facing ? R( 1 )/th.index : th.index ;
		const bool cannot = ratio*sin_theta>R( 1 ) ;
		if ( cannot || schlick( cos_theta, ratio )>rnd() )
			out = reflect( d1V, h.normal ) ;
		else
			out = refract( d1V, h.normal, ratio ) ;
		attened = mk<R>( R( 1 ), R( 1 ), R( 1 ) ) ;
		return true ;
	}

	// rtow.cxx:34-49.  The reference recurses and multiplies attenuations on the way
	// back (a1*(a2*(...*sky))); back_to_front=true reproduces that product order, false
	// is the float-mirror contract: throughput front to back, then times the sky.
	// guide layers (optx/optics_i.cu:97-101, 185-189): normal and albedo of the first diffuse
	// or reflecting hit of the path that is shaded
	bool  guide_set_ ;
	V3<R> guide_n_, guide_a_ ;

	V3<R> path( V3<R> ori, V3<R> dir, int depth, bool back_to_front, unsigned& segments, int* first_thing, int* first_prim, R* first_t ) {
		guide_set_ = false ;
		std::vector<V3<R>>& att = att_ ;
		att.clear() ;
		// rtow.cxx:39-42 allows `depth` scatter events and tests before it intersects; the OptiX
		// programs count rays: the depth-th hit is the last (optx/camera_i.cu:75-92, optx/optics_r.cu:30-42)
		if ( variant != 0 ) depth = depth>1 ? depth-1 : 0 ;
		V3<R> tail ;
		bool first = true ;
		while ( true ) {
			Hit<R> h ;
			segments++ ;
			const bool shot = closest( ori, dir, Consts<R>::acne(), Consts<R>::tmax(), h ) ;
			if ( log_ ) {
				const double rec[6] = { shot ? double( h.thing ) : -1., shot ? double( h.prim ) : -1., shot ? double( h.t ) : -1.,
					shot ? double( h.p.x ) : 0., shot ? double( h.p.y ) : 0., shot ? double( h.p.z ) : 0. } ;
				log_->insert( log_->end(), rec, rec+6 ) ;
			}
			if ( first ) {
				first = false ;
				if ( first_thing ) *first_thing = shot ? h.thing : -1 ;
				if ( first_prim )  *first_prim  = shot ? h.prim : -1 ;
				if ( first_t )     *first_t     = shot ? h.t : R( -1 ) ;
			}
			if ( shot ) {
				V3<R> attened, out ;
				// (the iterative programs run the hit program of the last ray too)
				const bool shaded = depth>0 || variant == 1 ;
				const bool go = shaded && scatter( scene->things[h.thing], dir, h, attened, out ) ;
				if ( shaded && ! guide_set_ && scene->things[h.thing].type != 2 ) {
					guide_set_ = true ; guide_n_ = h.normal ; guide_a_ = scene->things[h.thing].albedo ;
				}
				if ( go && depth == 0 ) {
					// optx/camera_i.cu:92-95: the loop ends with stat CONT, the throughput product is the colour
					att.push_back( attened ) ;
					tail = mk<R>( R( 1 ), R( 1 ), R( 1 ) ) ;
					break ;
				}
				if ( go ) {
					att.push_back( attened ) ;
					ori = h.p ; dir = out ; depth-- ;
					continue ;
				}
				tail = mk<R>( R( 0 ), R( 0 ), R( 0 ) ) ;
				break ;
			}
			const V3<R> unit = unitV( dir ) ;
			const R t = R( .5 )*( unit.y+R( 1 ) ) ;
			tail = ( R( 1 )-t )*mk<R>( R( 1 ), R( 1 ), R( 1 ) )+t*mk<R>( R( .5 ), R( .7 ), R( 1 ) ) ;
			break ;
		}
		if ( back_to_front ) {
			for ( size_t k = att.size() ; k>0 ; k-- )
				tail = att[k-1]*tail ;
			return tail ;
		}
		V3<R> thr = mk<R>( R( 1 ), R( 1 ), R( 1 ) ) ;
		for ( size_t k = 0 ; k<att.size() ; k++ )
			thr = thr*att[k] ;
		return thr*tail ;
	}

	std::vector<V3<R>> att_ ;
	// optional per-segment log: (thing, prim, t, px, py, pz) per closest-hit query
	std::vector<double>* log_ = nullptr ;
} ;

// fixed-point radiance contract of the product: a path's colour c (each channel in
// [0,1]) is accumulated as the integer trunc(c * 2^32); integer sums are associative,
// so any partition of the samples over lanes / GPUs gives the same bits.
inline uint64_t tofix( float c )  { return uint64_t( c*4294967296.f ) ; }
inline uint64_t tofix( double c ) { return uint64_t( c*4294967296. ) ; }

struct RenderArgs {
	int w, h, spp, depth ;
	uint64_t seed ;
	int sample0, sample_stride ;     // this call renders samples sample0 + k*stride, k < spp
	int y0, y1 ;                     // rows [y0,y1)
	int threads ;
	int back_to_front ;
	int variant ;                    // see Tracer::variant
	double*   sum ;                  // [h*w*3] double sums (f64 kinds) or nullptr
	uint64_t* fix ;                  // [h*w*3] fixed-point sums or nullptr
	uint32_t* rpp ;                  // [h*w] segments per pixel or nullptr
	int64_t*  first_id ;             // [h*w] (thing<<32|prim+1) of sample0's primary ray, -1 miss
	double*   first_t ;              // [h*w]
	int64_t*  guide ;                // [h*w*6] fixed-point (2^-30) sums of guide normal xyz, albedo rgb, or nullptr
} ;

// rtow.cxx:105-117: the pixel loop (rows h-1..0 for the libc stream, which is order
// sensitive; any order for the keyed stream)
template <class R, class Rng> void render_rows( const SceneT<R>& scene, const double* cam, const RenderArgs& a, int ya, int yb ) {
	Tracer<R, Rng> tr ;
	tr.scene = &scene ;
	tr.setcam( cam ) ;
	tr.variant = a.variant ;
	// rtow.cxx:112-113 divides by w-1, h-1; optx/camera_i.cu:61-62 and optx/camera_r.cu:69-70 by w, h
	const int wdiv = a.variant == 0 ? a.w-1 : a.w, hdiv = a.variant == 0 ? a.h-1 : a.h ;
	for ( int y = yb-1 ; y>=ya ; --y ) {
		for ( int x = 0 ; x<a.w ; ++x ) {
			const size_t pix = size_t( a.w )*y+x ;
			V3<R> color = mk<R>( R( 0 ), R( 0 ), R( 0 ) ) ;
			uint64_t fx[3] = { 0, 0, 0 } ;
			int64_t gd[6] = { 0, 0, 0, 0, 0, 0 } ;
			unsigned segments = 0 ;
			for ( int k = 0 ; k<a.spp ; ++k ) {
				const uint32_t sample = uint32_t( a.sample0+k*a.sample_stride ) ;
				tr.rng.seed( a.seed, uint32_t( pix ), sample ) ;
				// rtow.cxx:112-113
				const R s = R( 2 )*( R( x )+tr.rnd() )/R( wdiv )-R( 1 ) ;
				const R t = R( 2 )*( R( y )+tr.rnd() )/R( hdiv )-R( 1 ) ;
				V3<R> ori, dir ;
				tr.camray( s, t, ori, dir ) ;
				int ft = -1, fp = -1 ; R ftt = R( -1 ) ;
				const V3<R> c = tr.path( ori, dir, a.depth, a.back_to_front != 0, segments, k == 0 ? &ft : nullptr, k == 0 ? &fp : nullptr, k == 0 ? &ftt : nullptr ) ;
				if ( k == 0 ) {
					if ( a.first_id ) a.first_id[pix] = ft<0 ? int64_t( -1 ) : ( ( int64_t( ft )<<32 )|int64_t( uint32_t( fp+1 ) ) ) ;
					if ( a.first_t )  a.first_t[pix]  = double( ftt ) ;
				}
				color = color+c ;
				fx[0] += tofix( c.x ) ; fx[1] += tofix( c.y ) ; fx[2] += tofix( c.z ) ;
				if ( tr.guide_set_ ) {
					const R v[6] = { tr.guide_n_.x, tr.guide_n_.y, tr.guide_n_.z, tr.guide_a_.x, tr.guide_a_.y, tr.guide_a_.z } ;
					for ( int q = 0 ; q<6 ; q++ ) gd[q] += int64_t( v[q]*R( 1073741824 ) ) ;
				}
			}
			if ( a.guide ) for ( int q = 0 ; q<6 ; q++ ) a.guide[6*pix+q] = gd[q] ;
			if ( a.sum ) { a.sum[3*pix] = double( color.x ) ; a.sum[3*pix+1] = double( color.y ) ; a.sum[3*pix+2] = double( color.z ) ; }
			if ( a.fix ) { a.fix[3*pix] = fx[0] ; a.fix[3*pix+1] = fx[1] ; a.fix[3*pix+2] = fx[2] ; }
			if ( a.rpp ) a.rpp[pix] = segments ;
		}
	}
}

template <class R, class Rng> void render( const double* things, int n, const MeshRef* m, int nm, const double* cam, const RenderArgs& a, bool parallel ) {
	SceneT<R> scene ;
	scene.load( things, n, m, nm ) ;
	int nth = parallel ? ( a.threads>0 ? a.threads : int( std::thread::hardware_concurrency() ) ) : 1 ;
	if ( nth<1 ) nth = 1 ;
	if ( nth == 1 ) { render_rows<R, Rng>( scene, cam, a, a.y0, a.y1 ) ; return ; }
	// rows are dealt dynamically, one at a time: row costs differ a lot (sky vs glass)
	std::atomic<int> next( a.y0 ) ;
	std::vector<std::thread> pool ;
	for ( int q = 0 ; q<nth ; q++ )
		pool.emplace_back( [&]() {
			while ( true ) {
				const int y = next.fetch_add( 1 ) ;
				if ( y>=a.y1 ) break ;
				render_rows<R, Rng>( scene, cam, a, y, y+1 ) ;
			}
		} ) ;
	for ( auto& t : pool ) t.join() ;
}

} // namespace

extern "C" {

// kinds for orc_render
enum { ORC_F64_LIBC = 0, ORC_F64_PCG = 1, ORC_F32_PCG = 2 } ;

int orc_thing_stride()  { return TH_STRIDE ; }
int orc_camera_stride() { return CAM_STRIDE ; }

// rtow.cxx:51-80 with the process-global libc stream; g++ operand order made explicit
int orc_rtow_scene( double* things, int max_things ) {
	RngLibc r ;
	std::vector<double> rows ;
	auto add = [&]( double cx, double cy, double cz, double rad, int type, double ar, double ag, double ab, double fuzz, double index ) {
		double row[TH_STRIDE] = { 0 } ;
		row[TH_KIND] = 0 ; row[TH_MESH] = -1 ;
		row[TH_XF+0] = rad ; row[TH_XF+5] = rad ; row[TH_XF+10] = rad ;
		row[TH_XF+3] = cx ; row[TH_XF+7] = cy ; row[TH_XF+11] = cz ;
		row[TH_TYPE] = type ; row[TH_ALB] = ar ; row[TH_ALB+1] = ag ; row[TH_ALB+2] = ab ;
		row[TH_FUZZ] = fuzz ; row[TH_INDEX] = index ;
		rows.insert( rows.end(), row, row+TH_STRIDE ) ;
	} ;
	add( 0, -1000, 0, 1000., 0, .5, .5, .5, 0, 0 ) ;
	for ( int a = -11 ; a<11 ; a++ ) {
		for ( int b = -11 ; b<11 ; b++ ) {
			const double select = r.next() ;
			// rtow.cxx:59  P center( a+.9*rnd(), .2, b+.9*rnd() ) : z is drawn first
			const double cz = b+.9*r.next() ;
			const double cx = a+.9*r.next() ;
			const double cy = .2 ;
			const double dx = cx-4, dy = cy-.2, dz = cz-0 ;
			if ( std::sqrt( dx*dx+dy*dy+dz*dz )>.9 ) {
				if ( select<.8 ) {
					// rtow.cxx:62  C::rnd()*C::rnd() : each operand draws z,y,x; the
					// product is commutative so operand order does not matter
					const double z1 = r.next(), y1 = r.next(), x1 = r.next() ;
					const double z2 = r.next(), y2 = r.next(), x2 = r.next() ;
					add( cx, cy, cz, .2, 0, x2*x1, y2*y1, z2*z1, 0, 0 ) ;
				} else if ( select<.95 ) {
					const double z = .5+r.next()*( 1.-.5 ), y = .5+r.next()*( 1.-.5 ), x = .5+r.next()*( 1.-.5 ) ;
					const double fuzz = 0+r.next()*( .5-0 ) ;
					add( cx, cy, cz, .2, 1, x, y, z, fuzz, 0 ) ;
				} else
					add( cx, cy, cz, .2, 2, 0, 0, 0, 0, 1.5 ) ;
			}
		}
	}
	add(  0, 1, 0, 1., 2, 0, 0, 0, 0, 1.5 ) ;
	add( -4, 1, 0, 1., 0, .4, .2, .1, 0, 0 ) ;
	add(  4, 1, 0, 1., 1, .7, .6, .5, 0, 0 ) ;
	const int n = int( rows.size()/TH_STRIDE ) ;
	if ( things && n<=max_things )
		std::memcpy( things, rows.data(), rows.size()*sizeof( double ) ) ;
	return n ;
}

// camera.h:10-23 (double)
void orc_camera_set_f64( const double* eye, const double* pat, const double* vup, double fov, double aspratio, double aperture, double fostance, double* cam ) {
	typedef V3<double> V ;
	const double kPi = 3.141592653589793238 ;
	const V e = mk<double>( eye[0], eye[1], eye[2] ), p = mk<double>( pat[0], pat[1], pat[2] ), up = mk<double>( vup[0], vup[1], vup[2] ) ;
	const V w = unitV( e-p ) ;
	const V u = unitV( cross( up, w ) ) ;
	const V v = cross( w, u ) ;
	const double h = 2.*std::tan( .5*fov*kPi/180. ) ;
	const double wd = h*aspratio ;
	const V hvec = ( fostance*h/2. )*v ;
	const V wvec = ( fostance*wd/2. )*u ;
	const V dvec = fostance*w ;
	const V out[6] = { e, u, v, hvec, wvec, dvec } ;
	for ( int k = 0 ; k<6 ; k++ ) { cam[3*k] = out[k].x ; cam[3*k+1] = out[k].y ; cam[3*k+2] = out[k].z ; }
	cam[CAM_APERTURE] = aperture ;
}

// optx/camera.h:30-48 (float arithmetic, widened on output)
void orc_camera_set_f32( const float* eye, const float* pat, const float* vup, float fov, float aspratio, float aperture, float fostance, double* cam ) {
	typedef V3<float> V ;
	const float kPi = 3.14159265358979323846f ;                  // optx/util.h:22
	const V e = mk<float>( eye[0], eye[1], eye[2] ), p = mk<float>( pat[0], pat[1], pat[2] ), up = mk<float>( vup[0], vup[1], vup[2] ) ;
	const V w = unitV( e-p ) ;
	const V u = unitV( cross( up, w ) ) ;
	const V v = cross( w, u ) ;
	const float h  = 2.f*tanf( .5f*( fov*kPi/180.f ) ) ;
	const float wd = h*aspratio ;
	const V hvec = ( fostance*h/2.f )*v ;
	const V wvec = ( fostance*wd/2.f )*u ;
	const V dvec = fostance*w ;
	const V out[6] = { e, u, v, hvec, wvec, dvec } ;
	for ( int k = 0 ; k<6 ; k++ ) { cam[3*k] = out[k].x ; cam[3*k+1] = out[k].y ; cam[3*k+2] = out[k].z ; }
	cam[CAM_APERTURE] = aperture ;
}

uint64_t orc_libc_calls()            { return RngLibc::calls.load() ; }
void     orc_libc_reset( uint64_t skip ) {
	srand( 1 ) ;                       // the unseeded state rtow starts from
	for ( uint64_t k = 0 ; k<skip ; k++ ) rand() ;
	RngLibc::calls.store( skip ) ;
}

// meshes: nm entries; vces[q] -> float[3*nv[q]], ices[q] -> uint32[3*nt[q]]
// variant of the following orc_render calls (see Tracer::variant); 0 = rtow.cxx
static int g_variant = 0 ;
void orc_set_variant( int v ) { g_variant = v ; }

int orc_render( int kind, const double* things, int n_things,
		int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int spp, int depth, uint64_t seed, int sample0, int sample_stride,
		int y0, int y1, int threads,
		double* sum, uint64_t* fix, uint32_t* rpp, int64_t* first_id, double* first_t, int64_t* guide ) {
	std::vector<MeshRef> m( n_meshes ) ;
	for ( int q = 0 ; q<n_meshes ; q++ ) { m[q].vces = vces[q] ; m[q].nv = nv[q] ; m[q].ices = ices[q] ; m[q].nt = nt[q] ; }
	RenderArgs a ;
	a.w = w ; a.h = h ; a.spp = spp ; a.depth = depth ; a.seed = seed ; a.sample0 = sample0 ; a.sample_stride = sample_stride ;
	a.y0 = y0 ; a.y1 = y1 ; a.threads = threads ; a.variant = g_variant ;
	a.sum = sum ; a.fix = fix ; a.rpp = rpp ; a.first_id = first_id ; a.first_t = first_t ; a.guide = guide ;
	switch ( kind ) {
		case ORC_F64_LIBC: a.back_to_front = 1 ; render<double, RngLibc>( things, n_things, m.data(), n_meshes, cam, a, false ) ; break ;
		case ORC_F64_PCG:  a.back_to_front = 1 ; render<double, RngPcg>( things, n_things, m.data(), n_meshes, cam, a, true ) ; break ;
		case ORC_F32_PCG:  a.back_to_front = 0 ; render<float, RngPcg>( things, n_things, m.data(), n_meshes, cam, a, true ) ; break ;
		default: return 1 ;
	}
	return 0 ;
}

} // extern "C"

// one path, segment by segment (debugging aid for parity hunts): returns the number of
// segments; log receives 6 doubles per segment (thing, prim, t, hit point)
template <class R> static int path_log( const double* things, int n_things, const std::vector<MeshRef>& m, const double* cam,
		int w, int h, int depth, uint64_t seed, int x, int y, int sample, double* log, int max_segments, double* rgb ) {
	SceneT<R> scene ;
	scene.load( things, n_things, m.data(), int( m.size() ) ) ;
	Tracer<R, RngPcg> tr ;
	tr.scene = &scene ;
	tr.setcam( cam ) ;
	std::vector<double> rec ;
	tr.log_ = &rec ;
	tr.rng.seed( seed, uint32_t( size_t( w )*y+x ), uint32_t( sample ) ) ;
	const R s = R( 2 )*( R( x )+tr.rnd() )/R( w-1 )-R( 1 ) ;
	const R t = R( 2 )*( R( y )+tr.rnd() )/R( h-1 )-R( 1 ) ;
	V3<R> ori, dir ;
	tr.camray( s, t, ori, dir ) ;
	unsigned segments = 0 ;
	const V3<R> c = tr.path( ori, dir, depth, false, segments, nullptr, nullptr, nullptr ) ;
	if ( rgb ) { rgb[0] = double( c.x ) ; rgb[1] = double( c.y ) ; rgb[2] = double( c.z ) ; }
	const int n = int( rec.size()/6 ) ;
	for ( int k = 0 ; k<n && k<max_segments ; k++ )
		std::memcpy( log+6*k, rec.data()+6*k, 6*sizeof( double ) ) ;
	return n ;
}

extern "C" int orc_path_log( int kind, const double* things, int n_things,
		int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int depth, uint64_t seed, int x, int y, int sample, double* log, int max_segments, double* rgb ) {
	std::vector<MeshRef> m( n_meshes ) ;
	for ( int q = 0 ; q<n_meshes ; q++ ) { m[q].vces = vces[q] ; m[q].nv = nv[q] ; m[q].ices = ices[q] ; m[q].nt = nt[q] ; }
	if ( kind == 2 )
		return path_log<float>( things, n_things, m, cam, w, h, depth, seed, x, y, sample, log, max_segments, rgb ) ;
	return path_log<double>( things, n_things, m, cam, w, h, depth, seed, x, y, sample, log, max_segments, rgb ) ;
}

extern "C" {

// closest hit of arbitrary rays in the float mirror ("identical ray set" checks)
int orc_trace_rays_f32( const double* things, int n_things,
		int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		int n_rays, const float* ori, const float* dir, float tmin, int threads, int64_t* id, float* t_out ) {
	std::vector<MeshRef> m( n_meshes ) ;
	for ( int q = 0 ; q<n_meshes ; q++ ) { m[q].vces = vces[q] ; m[q].nv = nv[q] ; m[q].ices = ices[q] ; m[q].nt = nt[q] ; }
	SceneT<float> scene ;
	scene.load( things, n_things, m.data(), n_meshes ) ;
	int nth = threads>0 ? threads : int( std::thread::hardware_concurrency() ) ;
	if ( nth<1 ) nth = 1 ;
	std::atomic<int> next( 0 ) ;
	std::vector<std::thread> pool ;
	for ( int q = 0 ; q<nth ; q++ )
		pool.emplace_back( [&]() {
			Tracer<float, RngPcg> tr ;
			tr.scene = &scene ;
			while ( true ) {
				const int r0 = next.fetch_add( 256 ) ;
				if ( r0>=n_rays ) break ;
				for ( int r = r0 ; r<r0+256 && r<n_rays ; r++ ) {
					Hit<float> h ;
					const V3<float> o = mk<float>( ori[3*r], ori[3*r+1], ori[3*r+2] ) ;
					const V3<float> d = mk<float>( dir[3*r], dir[3*r+1], dir[3*r+2] ) ;
					if ( tr.closest( o, d, tmin, Consts<float>::tmax(), h ) ) {
						id[r] = ( int64_t( h.thing )<<32 )|int64_t( uint32_t( h.prim+1 ) ) ;
						if ( t_out ) t_out[r] = h.t ;
					} else {
						id[r] = -1 ;
						if ( t_out ) t_out[r] = -1.f ;
					}
				}
			}
		} ) ;
	for ( auto& t : pool ) t.join() ;
	return 0 ;
}

// rtow.cxx:6-21: gamma 2, 256*clamp(.,0,.999); rows are emitted h-1..0 by the caller
void orc_ppm_rtow( const double* sum, int spp, size_t npix, uint8_t* rgb ) {
	for ( size_t p = 0 ; p<npix ; p++ )
		for ( int c = 0 ; c<3 ; c++ ) {
			// color/spp is (1/spp)*color (v.h:46)
			double v = ( 1/double( spp ) )*sum[3*p+c] ;
			v = std::sqrt( v ) ;
			v = 0>v ? 0 : v>.999 ? .999 : v ;
			rgb[3*p+c] = uint8_t( int( 256*v ) ) ;
		}
}

// the product's resolve contract: mean of the fixed-point sums, clamped like
// optx/camera_i.cu:105
void orc_resolve_fix( const uint64_t* fix, uint64_t total_spp, size_t npix, float* raw ) {
	for ( size_t k = 0 ; k<3*npix ; k++ ) {
		const float v = float( double( fix[k] )*( 1./4294967296. )/double( total_spp ) ) ;
		raw[k] = 0.f>v ? 0.f : v>1.f ? 1.f : v ;
	}
}

// optx/postproc.cu:2-16, 36-47 (sRGB) and :18-34 (none)
// x^(1/2.4) of the sRGB transfer (optx/postproc.cu:2-16 calls powf, which no two libraries round
// alike): here, as in the CUDA path, one stated sequence of IEEE double operations -- the cube
// root by 14 Newton steps from 1 (x in [0.0031308, 1]: converged after 10), then
// x^(5/12) = sqrt(sqrt(c^5)) -- rounded to float once.  Every step is a correctly rounded basic
// operation, so host and device agree bit for bit.
static float srgb_pow( float x ) {
	const double v = double( x ) ;
	double c = 1. ;
	for ( int i = 0 ; i<14 ; i++ )
		c = ( 2.*c+v/( c*c ) )*( 1./3. ) ;
	const double c2 = c*c ;
	return float( std::sqrt( std::sqrt( ( c2*c2 )*c ) ) ) ;
}

void orc_srgb8( const float* raw, size_t npix, int srgb, uint8_t* rgba ) {
	for ( size_t p = 0 ; p<npix ; p++ ) {
		for ( int c = 0 ; c<3 ; c++ ) {
			float v = raw[3*p+c] ;
			if ( srgb )
				v = v<.0031308f ? 12.92f*v : 1.055f*srgb_pow( v )-.055f ;
			rgba[4*p+c] = static_cast<unsigned char>( v*255 ) ;
		}
		rgba[4*p+3] = 255u ;
	}
}

} // extern "C"
