// rtx_core.cuh -- device-side core of the B200 path tracer: vector math, the keyed
// random stream, two-level LBVH traversal with exact primitive tests, and the three
// scattering models.  Everything here is __host__ __device__ so that the SAME source can
// be driven serially on the CPU by the development harness in tests/hostemu (no GPU in
// the build container); the product library only ever runs it in kernels.
//
// Arithmetic contract ("float contract"): this translation unit is
// compiled with -fmad=false, so every + - * / sqrt below is one IEEE rounding, in the
// order written; the only fused operations are the explicit fmaf() calls of the
// box test, which is conservative and never decides a result.  Reference semantics
// followed (file:line under /root/reference): rtow.cxx:34-49 (trace), sphere.h:20-48,
// things.h:22-36, optics.h:11-75, v.h:42-62, camera.h:25-31, rtow.cxx:112-113;
// triangle mode: optx/optics_i.cu:40-82, :259-267.
#pragma once

#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>

#if defined( __CUDACC__ )
#define RTX_HD __host__ __device__ __forceinline__
// big, rarely executed pieces (shading frame, scattering, ray generation) are real calls:
// the render kernel is ~50 KB of code run by ~20 independent warps per SM, and instruction
// fetch was its top stall (ncu: stall_no_inst 28 %)
#if defined( RTX_INLINE_ALL )
#define RTX_HD_CALL __host__ __device__ __forceinline__
#else
#define RTX_HD_CALL __host__ __device__ __noinline__
#endif
#else
#define RTX_HD inline
#define RTX_HD_CALL inline
#endif

#if defined( __CUDA_ARCH__ )
#define RTX_LDG( p ) __ldg( p )
// warp vote: the tracing loops are written warp-synchronously (all 32 lanes stay in the
// loop until none has work) so that the lanes of a warp reconverge every iteration
#define RTX_ANY( pred ) __any_sync( 0xffffffffu, ( pred ) )
#else
#define RTX_LDG( p ) ( *( p ) )
#define RTX_ANY( pred ) ( pred )
#endif

#if defined( RTX_DEVICE_COUNTERS ) && defined( __CUDACC__ )
namespace rtx { enum { DC_rays, DC_nodes, DC_leaves, DC_tris, DC_things, DC_spheres, DC_enters, DC_N } ; __device__ unsigned long long g_dev_counts[DC_N] ;   // (librtx has one CUDA translation unit)
	// scheduling statistics of the render kernels: warp iterations and lane-steps per step kind, paths alive per bounce
	__device__ unsigned long long g_dev_steps[8], g_dev_lanes[8], g_dev_live[64] ; }
#endif
// traversal statistics for the host harness (compiled out of the product)
#if defined( RTX_STATS )
namespace rtx { struct Stats { unsigned long long rays, nodes, leaves, tris, things, spheres, enters, pushes, maxsp ; } ; extern Stats g_stats ; }
#endif
#if defined( RTX_STATS ) && ! defined( __CUDA_ARCH__ )
namespace rtx { void trace_event( char c ) ; }
#define RTX_EVENT( c ) rtx::trace_event( c )
#define RTX_COUNT( f ) ( rtx::g_stats.f++ )
#define RTX_COUNT_MAX( f, v ) do { if ( ( unsigned long long )( v )>rtx::g_stats.f ) rtx::g_stats.f = ( v ) ; } while ( 0 )
#elif defined( RTX_DEVICE_COUNTERS ) && defined( __CUDA_ARCH__ )
// instrumented build (librtx_count.so, bench.py "counted"): the same events counted on the
// device with global atomics -- a measuring instrument, far too slow for the product library
#define RTX_EVENT( c ) ( ( void ) 0 )
#define RTX_COUNT( f ) ( ( void ) atomicAdd( &rtx::g_dev_counts[rtx::DC_##f], 1ull ) )
#define RTX_COUNT_MAX( f, v ) ( ( void ) 0 )
#else
#define RTX_EVENT( c ) ( ( void ) 0 )
#define RTX_COUNT( f ) ( ( void ) 0 )
#define RTX_COUNT_MAX( f, v ) ( ( void ) 0 )
#endif
#if defined( RTX_DEVICE_COUNTERS ) && defined( __CUDA_ARCH__ )
// one warp iteration of step kind k with `mask` = the lanes that take part (call warp-uniformly)
#define RTX_COUNT_STEP( k, mask ) do { const unsigned m_ = ( mask ) ; if ( ( threadIdx.x&31u ) == 0u ) { atomicAdd( &rtx::g_dev_steps[( k )&7], 1ull ) ; atomicAdd( &rtx::g_dev_lanes[( k )&7], ( unsigned long long ) __popc( m_ ) ) ; } } while ( 0 )   // (the mask is a warp vote: every lane evaluates it)
// a ray of a path's bounce b (0 = primary) has been traced
#define RTX_COUNT_LIVE( b ) ( ( void ) atomicAdd( &rtx::g_dev_live[( b )<63u ? ( b ) : 63u], 1ull ) )
#else
#define RTX_COUNT_STEP( k, mask ) ( ( void ) 0 )
#define RTX_COUNT_LIVE( b ) ( ( void ) 0 )
#endif

namespace rtx {

// ----------------------------------------------------------------------------- vectors
struct f3 { float x, y, z ; } ;
struct d3 { double x, y, z ; } ;

RTX_HD f3 mk3( float x, float y, float z ) { f3 v ; v.x = x ; v.y = y ; v.z = z ; return v ; }
RTX_HD d3 mk3( double x, double y, double z ) { d3 v ; v.x = x ; v.y = y ; v.z = z ; return v ; }

#define RTX_VEC_OPS( V, S )                                                                                  \
RTX_HD V operator + ( const V& u, const V& v ) { return mk3( u.x+v.x, u.y+v.y, u.z+v.z ) ; }                 \
RTX_HD V operator - ( const V& u, const V& v ) { return mk3( u.x-v.x, u.y-v.y, u.z-v.z ) ; }                 \
RTX_HD V operator - ( const V& u )             { return mk3( -u.x, -u.y, -u.z ) ; }                          \
RTX_HD V operator * ( const V& u, const V& v ) { return mk3( u.x*v.x, u.y*v.y, u.z*v.z ) ; }                 \
RTX_HD V operator * ( S t, const V& v )        { return mk3( t*v.x, t*v.y, t*v.z ) ; }                       \
RTX_HD S dot( const V& u, const V& v )         { return u.x*v.x+u.y*v.y+u.z*v.z ; }                          \
RTX_HD V cross( const V& u, const V& v )       { return mk3( u.y*v.z-u.z*v.y, u.z*v.x-u.x*v.z, u.x*v.y-u.y*v.x ) ; }
RTX_VEC_OPS( f3, float )
RTX_VEC_OPS( d3, double )
#undef RTX_VEC_OPS

// v.h:46,51: division is multiplication by the reciprocal; unitV(v) = v/len(v)
RTX_HD f3 unitV( const f3& v ) { return ( 1.f/sqrtf( dot( v, v ) ) )*v ; }
RTX_HD d3 unitV( const d3& v ) { return ( 1./sqrt( dot( v, v ) ) )*v ; }
RTX_HD d3 wide( const f3& v )   { return mk3( double( v.x ), double( v.y ), double( v.z ) ) ; }
RTX_HD f3 narrow( const d3& v ) { return mk3( float( v.x ), float( v.y ), float( v.z ) ) ; }

// row-major 3x4 affine map, written order = contract
RTX_HD d3 xfpoint( const double* m, const d3& p ) {
	return mk3( p.x*m[0]+p.y*m[1]+p.z*m[2]+m[3], p.x*m[4]+p.y*m[5]+p.z*m[6]+m[7], p.x*m[8]+p.y*m[9]+p.z*m[10]+m[11] ) ;
}
RTX_HD d3 xfvec( const double* m, const d3& p ) {
	return mk3( p.x*m[0]+p.y*m[1]+p.z*m[2], p.x*m[4]+p.y*m[5]+p.z*m[6], p.x*m[8]+p.y*m[9]+p.z*m[10] ) ;
}
// the same maps when the off-diagonal terms are zero: the dropped products are +-0 and adding
// them changes no value, so these give the bits of the general form with a third of the work
RTX_HD d3 xfpoint_diag( double m0, double m3, double m5, double m7, double m10, double m11, const d3& p ) {
	return mk3( p.x*m0+m3, p.y*m5+m7, p.z*m10+m11 ) ;
}
RTX_HD d3 xfvec_diag( double m0, double m5, double m10, const d3& p ) {
	return mk3( p.x*m0, p.y*m5, p.z*m10 ) ;
}

// ----------------------------------------------------------------------------- random stream
// Stream of path (pixel, sample) = PCG32 (XSH-RR 64/32) started at a splitmix64 hash of
// (seed, pixel, sample): any partition of the samples over lanes or GPUs sees the same
// numbers.  Replaces Frand48 (optx/frand48.h:17-35), whose stream is per pixel only.
struct Pcg {
	uint64_t state ;
	RTX_HD static uint64_t mix( uint64_t z ) {
		z = ( z^( z>>30 ) )*0xBF58476D1CE4E5B9ull ;
		z = ( z^( z>>27 ) )*0x94D049BB133111EBull ;
		return z^( z>>31 ) ;
	}
	RTX_HD void seed( uint64_t seed, uint32_t pixel, uint32_t sample ) {
		state = mix( ( ( uint64_t( pixel )<<32 )|uint64_t( sample ) )^( seed*0x9E3779B97F4A7C15ull ) ) ;
	}
	RTX_HD uint32_t bits() {
		const uint64_t old = state ;
		state = old*6364136223846793005ull+1442695040888963407ull ;
		const uint32_t xs  = uint32_t( ( ( old>>18 )^old )>>27 ) ;
		const uint32_t rot = uint32_t( old>>59 ) ;
		return ( xs>>rot )|( xs<<( ( 32-rot )&31 ) ) ;
	}
	// util.h:12 rnd(): here 24 random bits, exact in float
	RTX_HD float rnd() { return float( bits()>>8 )*( 1.f/16777216.f ) ; }
	// util.h:13
	RTX_HD float rnd( float min, float max ) { return min+rnd()*( max-min ) ; }
	// v.h:30-31 as g++ sequences it: z, y, x
	RTX_HD f3 rndV( float min, float max ) { const float z = rnd( min, max ) ; const float y = rnd( min, max ) ; const float x = rnd( min, max ) ; return mk3( x, y, z ) ; }
	// v.h:53
	RTX_HD f3 rndVin1sphere() { while ( true ) { const f3 v = rndV( -1.f, 1.f ) ; if ( 1.f>dot( v, v ) ) return v ; } }
	// v.h:55
	RTX_HD f3 rndVon1sphere() { return unitV( rndVin1sphere() ) ; }
	// v.h:59 as g++ sequences it: y, x
	RTX_HD f3 rndVin1disk() { while ( true ) { const float y = rnd( -1.f, 1.f ) ; const float x = rnd( -1.f, 1.f ) ; const f3 v = mk3( x, y, 0.f ) ; if ( 1.f>dot( v, v ) ) return v ; } }
} ;

// ----------------------------------------------------------------------------- scene (device view)
struct q4 { float x, y, z, w ; } ;   // 16-byte record, bit-compatible with float4

// Wide (4-ary) BVH node, 128 bytes = 8 x 16-byte records = one cache line, child boxes
// in SoA form so that each record is one vector load feeding four slab tests:
//   r0 = lo.x[0..3]  r1 = lo.y[0..3]  r2 = lo.z[0..3]
//   r3 = hi.x[0..3]  r4 = hi.y[0..3]  r5 = hi.z[0..3]
//   r6 = child refs[0..3] (bit patterns)   r7 = unused
// child ref: 0 <= ref < RTX_REF_EMPTY inner node index; ref < 0 leaf, ~ref = first<<3 |
// (count-1); RTX_REF_EMPTY marks an unused slot, whose box is lo = hi = +inf (never entered).  The two values above it are stack
// sentinels.
#ifndef RTX_WIDTH
#define RTX_WIDTH      4            // children per node: 4, or 8 (experimental: two such 128-byte blocks per node, children 0-3 and 4-7)
#endif
#define RTX_NODE_RECS  ( 2*RTX_WIDTH )
// Child boxes of a 4-wide node are stored in centre / half-extent form (r0..r2 = centre x, y, z of the
// four children, r3..r5 = half extents): the slab test then needs no min / max per plane pair --
// t_near = (c.d' - o.d') - h|d'|, t_far = ... + h|d'| are three multiply-adds per axis -- 18 packed FMAs
// instead of 12 packed FMAs + 24 min/max per node step (the FMA pipe has room, the ALU pipe does not).
// [c-h, c+h] contains the padded [lo, hi] (box_ch rounds h up); an unused slot is c = +inf, h = 0.
// (The experimental 8-wide layout keeps lo / hi.)
#ifndef RTX_NODE_CH
#define RTX_NODE_CH    ( RTX_WIDTH == 4 )
#endif
#define RTX_TRI_RECS   4            // a triangle: (a, prim id) (e1, b.x) (e2, b.y) (b.z, c) = 64 bytes, two 256-bit loads
#ifndef RTX_LEAF_MAX
#define RTX_LEAF_MAX   3            // triangles per mesh leaf, at most 8 (top level: 1 thing per leaf); measured 2/3/4/6/8: 708/684/698/690/697 ms
#endif
#define RTX_REF_EMPTY  0x7ffffffd
#define RTX_STK_DONE   0x7ffffffe   // bottom of the stack
#define RTX_STK_RETURN 0x7fffffff   // leave the mesh, back to the top level

// per-thing record read by traversal (128 bytes)
struct ThingTrav {
	double     inv[12] ;   // world->object (mesh); analytic sphere: inv[0..3] = cx, cy, cz, r
	const q4*  nodes ;     // mesh LBVH
	const q4*  tris ;      // mesh triangles in leaf order: (v0, asfloat(prim)), (e1,-), (e2,-), pad
	int32_t    kind ;      // 0 analytic sphere, 1 mesh instance
	uint32_t   n_tris ;
	int32_t    diag ;      // 1: the transform is a per-axis scale + translation (the off-diagonal terms are zero)
	int32_t    pad ;
} ;

// per-thing record read by shading (160 bytes = 5 x 32: the pool kernel fetches it with 256-bit loads)
struct ThingShade {
	double          xf[12] ;  // object->world
	float           albedo[3], fuzz ;
	float           index ;
	int32_t         type ;    // rtx_optics type
	int32_t         kind ;
	int32_t         diag ;    // as in ThingTrav
	const float*    vces ;    // mesh vertices (xyz) and indices as uploaded
	const uint32_t* ices ;
	double          pad_[2] ;
} ;

// Which of the reference's own programs the path semantics follow (SURVEY.md 8a "divergences"):
//   RTOW    rtow.cxx, the CPU path and the parity target: pixel -> viewport over w-1 / h-1, Lambert with
//           the near-zero guard, `depth` scatter events and black when they are used up
//   RTWO_I  the iterative OptiX programs (optx/camera_i.cu:59-95, optx/optics_i.cu:86): over w / h, no
//           guard, at most `depth` rays per path, and a path whose last ray still scattered keeps its
//           throughput product as colour
//   RTWO_R  the recursive ones (optx/camera_r.cu:67-70, optx/optics_r.cu:30-42, :108): over w / h, no
//           guard, black at the `depth`-th hit
enum { RTX_SEM_RTOW = 0, RTX_SEM_RTWO_I = 1, RTX_SEM_RTWO_R = 2 } ;

struct SceneDev {
	const q4*         tlas_nodes ;
	const uint32_t*   tlas_order ;   // leaf slot -> thing id
	const q4*         bsphere ;      // per thing: world bounding sphere (cx,cy,cz,r) of a mesh instance; r < 0: none
	const ThingTrav*  trav ;
	const ThingShade* shade ;
	uint32_t          n_things ;
	uint32_t          variant ;      // RTX_SEM_* of the frame being rendered
	const q4*         arena ;        // lowest address of all node / triangle arrays: the pool kernel keeps them as 32-bit offsets (16-byte units) from here
	uint32_t*         fault ;        // device word: bit 0 set by a traversal whose stack ran out of room (the launch is then reported as failed)
} ;

// a traversal stack ran out of room: the ray can no longer be finished correctly -- say so
// (rtx_render* / rtx_primary_hits / rtx_trace_rays / rtx_pick fail with "traversal stack overflow")
RTX_HD void stack_fault( uint32_t* fault ) {
#if defined( __CUDA_ARCH__ )
	if ( fault ) atomicOr( fault, 1u ) ;
#else
	if ( fault ) *fault |= 1u ;
#endif
}

struct CameraDev { f3 eye, u, v, hvec, wvec, dvec ; float aperture ; } ;

struct HitRec {
	float   t ;
	int32_t thing ;   // -1: miss
	int32_t prim ;    // -1: analytic
	float   u, v ;    // barycentrics (mesh)
	int32_t slot ;    // triangle record of the hit in the mesh's leaf-ordered array
} ;

// shading frame of a hit
struct Frame { f3 p, normal ; bool facing ; } ;

RTX_HD q4 ldq( const q4* p ) {
#if defined( __CUDA_ARCH__ )
	const float4 v = __ldg( reinterpret_cast<const float4*>( p ) ) ;
	q4 r ; r.x = v.x ; r.y = v.y ; r.z = v.z ; r.w = v.w ; return r ;
#else
	return *p ;
#endif
}
// two adjacent records with ONE 256-bit load (sm_100: LDG.E.256): a lane that fetches a
// whole 128-byte node touches its cache line 4 times instead of 7 -- the L1 tag stage was a
// co-bottleneck of the traversal (ncu: l1tex throughput 47 %)
struct o8 { q4 a, b ; } ;
// L1 policy of the node / triangle fetches (tuning knobs; default: the plain non-coherent load)
#ifndef RTX_NODE_LD
#define RTX_NODE_LD "ld.global.nc.v8.f32"
#endif
#ifndef RTX_TRI_LD
#define RTX_TRI_LD "ld.global.nc.v8.f32"
#endif
#define RTX_LDO_BODY( INSN )                                                                                   \
	o8 r ;                                                                                                     \
	asm( INSN " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                              \
		: "=f"( r.a.x ), "=f"( r.a.y ), "=f"( r.a.z ), "=f"( r.a.w ), "=f"( r.b.x ), "=f"( r.b.y ), "=f"( r.b.z ), "=f"( r.b.w ) : "l"( p ) ) ; \
	return r ;
RTX_HD o8 ldo( const q4* p ) {
#if defined( __CUDA_ARCH__ )
	RTX_LDO_BODY( RTX_NODE_LD )
#else
	o8 r ; r.a = p[0] ; r.b = p[1] ; return r ;
#endif
}
RTX_HD o8 ldo_tri( const q4* p ) {
#if defined( __CUDA_ARCH__ )
	RTX_LDO_BODY( RTX_TRI_LD )
#else
	o8 r ; r.a = p[0] ; r.b = p[1] ; return r ;
#endif
}
// __ldg has no pointer overload: read the 8 bytes as an integer
template <class T> RTX_HD const T* ldptr( const T* const* pp ) {
#if defined( __CUDA_ARCH__ )
	return reinterpret_cast<const T*>( __ldg( reinterpret_cast<const unsigned long long*>( pp ) ) ) ;
#else
	return *pp ;
#endif
}
RTX_HD int32_t asint( float f ) {
#if defined( __CUDA_ARCH__ )
	return __float_as_int( f ) ;
#else
	union { float f ; int32_t i ; } c ; c.f = f ; return c.i ;
#endif
}
RTX_HD float asfloat( int32_t i ) {
#if defined( __CUDA_ARCH__ )
	return __int_as_float( i ) ;
#else
	union { float f ; int32_t i ; } c ; c.i = i ; return c.f ;
#endif
}

// ----------------------------------------------------------------------------- primitive tests
// sphere.h:21-38 verbatim, in double on the widened float ray: the smallest root not
// below tmin.  (things.h:27-33 keeps it when it does not exceed the best t so far.)
RTX_HD bool sphere_root( const d3& center, double radius, const d3& ori, const d3& dir, double tmin, double& t ) {
	const d3 o = ori-center ;
	const double a = dot( dir, dir ) ;
	const double b = dot( dir, o ) ;
	const double c = dot( o, o )-radius*radius ;
	const double discriminant = b*b-a*c ;
	if ( 0.>discriminant )
		return false ;
	const double x = sqrt( discriminant ) ;
	t = ( -b-x )/a ;
	if ( tmin>t ) {
		t = ( -b+x )/a ;
		if ( tmin>t )
			return false ;
	}
	return true ;
}

// Moeller-Trumbore, two-sided, on the stored (v0, e1, e2) in object space; the origin is
// the float pair hi+lo of the double-precision object-space origin.
RTX_HD bool tri_test( const f3& v0, const f3& e1, const f3& e2, const f3& ohi, const f3& olo, const f3& d, float tmin, float& t, float& u, float& v ) {
	const f3 p = cross( d, e2 ) ;
	const float det = dot( e1, p ) ;
	if ( det == 0.f )
		return false ;
	const float inv = 1.f/det ;
	const f3 s = ( ohi-v0 )+olo ;
	u = dot( s, p )*inv ;
	if ( ! ( u>=0.f && u<=1.f ) )
		return false ;
	const f3 q = cross( s, e1 ) ;
	v = dot( d, q )*inv ;
	if ( ! ( v>=0.f && u+v<=1.f ) )
		return false ;
	t = dot( e2, q )*inv ;
	return t>=tmin ;
}

// Can the ray segment [tmin, tbest] touch the bounding sphere?  A thing's box is a poor fit
// for round or distant-origin cases: most visits of a mesh found nothing (measured in the
// host harness: 2/3 of them, half of all node steps) -- above all the ray that has just left
// a small sphere and starts inside its box.  Conservative: the sphere is padded at build
// time, the comparisons carry a margin; never decides a result.
// Evaluated without the cancellation of |o-c|^2 - r^2 (whose float error grows with
// (distance/radius)^2 and exceeded the pad beyond distance/radius ~ 30 -- round-1 advice): the
// closest approach l = f + tc*d of the line to the centre carries an absolute error of a few
// eps*|f| (f = o-c is exact to half an ulp, tc = -(f.d)/(d.d) to ~4 eps*|f|/|d|), which the
// radius margin below covers five times over; the chord interval tc -+ h gets a relative slack.
RTX_HD float approx_rcp( float x ) {
#if defined( __CUDA_ARCH__ )
	float r ;
	asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ) ;   // 1 ulp; inside the margins
	return r ;
#else
	return 1.f/x ;
#endif
}
RTX_HD bool bsphere_miss( const q4& bs, const f3& o, const f3& d, float tmin, float tbest ) {
	if ( bs.w<0.f )
		return false ;
	const float fx = o.x-bs.x, fy = o.y-bs.y, fz = o.z-bs.z ;
	const float a = d.x*d.x+d.y*d.y+d.z*d.z ;
	const float b = fx*d.x+fy*d.y+fz*d.z ;
	const float ia = approx_rcp( a ) ;
	const float tc = -b*ia ;
	const float lx = fx+tc*d.x, ly = fy+tc*d.y, lz = fz+tc*d.z ;
	const float l2 = lx*lx+ly*ly+lz*lz ;
	const float m = 4e-6f*( fabsf( fx )+fabsf( fy )+fabsf( fz ) ) ;
	const float r = bs.w+m ;
	const float h2 = r*r-l2 ;
	if ( h2<0.f )
		return true ;                                   // the line passes the sphere
	// half chord in t, widened by the absolute error m/|d| of tc: (sqrt(h2)+m)^2 <= h2+m(2r+m)
	const float h = sqrtf( ( h2+m*( r+r+m ) )*ia )*1.00001f ;
	const float slack = 1e-5f*fabsf( tc ) ;
	return tc+h+slack<tmin || tc-h-slack>tbest ;        // (a NaN from a degenerate ray compares false: no cull)
}

// order-independent closest-hit rule = things.h:27-33 scanned in thing/primitive order:
// smaller t wins, at exactly equal t the later-listed (thing, prim) wins
RTX_HD bool better( float t, int32_t thing, int32_t prim, const HitRec& best ) {
	return t<best.t || ( t == best.t && ( thing>best.thing || ( thing == best.thing && prim>best.prim ) ) ) ;
}

// ----------------------------------------------------------------------------- box test
// Conservative slab test (boxes are padded at build time, the far bound carries a
// relative slack): may visit too much, never too little.  This is the one place where
// FMA and an approximate reciprocal are used on purpose -- it never decides a result.
#define RTX_SLACK 1.0000038f
// one axis of a child box for the node record: (lo, hi) -> what the record stores
RTX_HD void box_ch( float lo, float hi, float& a, float& b ) {
#if RTX_NODE_CH
	if ( lo == INFINITY ) { a = INFINITY ; b = 0.f ; return ; }   // unused slot: entered by no ray
	const float c = .5f*lo+.5f*hi ;
	a = c ;
	b = fmaxf( hi-c, c-lo )*1.000001f+1e-37f ;                      // rounded up: [c-b, c+b] contains [lo, hi]
#else
	a = lo ; b = hi ;
#endif
}
RTX_HD float safe_rcp( float d ) {
	// 1/d clamped to +-1e30: a zero component gives a huge finite slope instead of inf, so
	// that o*idir and the slab products stay finite (no inf-inf, no 0*inf)
#if defined( __CUDA_ARCH__ )
	// one MUFU.RCP: flushing a denormal d to zero gives +-inf, which the clamp turns into the
	// same +-1e30 as its huge reciprocal (the non-flushing form costs four more instructions)
	float r ;
	asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( d ) ) ;
	return fminf( fmaxf( r, -1e30f ), 1e30f ) ;
#else
	return fminf( fmaxf( 1.f/d, -1e30f ), 1e30f ) ;
#endif
}
// entry distance of the ray into box (lo,hi), or +inf when it misses [tmin, tbest]
RTX_HD float slab( float lox, float loy, float loz, float hix, float hiy, float hiz, const f3& idir, const f3& ood, float tmin, float tbest_s ) {
#if RTX_NODE_CH
	// (lox.. are the centre, hix.. the half extents)
	const float tx = fmaf( lox, idir.x, -ood.x ), ty = fmaf( loy, idir.y, -ood.y ), tz = fmaf( loz, idir.z, -ood.z ) ;
	const float ax = fabsf( idir.x ), ay = fabsf( idir.y ), az = fabsf( idir.z ) ;
	const float x0 = fmaf( hix, -ax, tx ), x1 = fmaf( hix, ax, tx ) ;
	const float y0 = fmaf( hiy, -ay, ty ), y1 = fmaf( hiy, ay, ty ) ;
	const float z0 = fmaf( hiz, -az, tz ), z1 = fmaf( hiz, az, tz ) ;
	const float tn = fmaxf( fmaxf( x0, y0 ), fmaxf( z0, tmin ) ) ;
	const float tf = fminf( fminf( x1, y1 ), fminf( z1, tbest_s ) ) ;
	return tn<=tf*RTX_SLACK ? tn : INFINITY ;
#else
	const float x0 = fmaf( lox, idir.x, -ood.x ), x1 = fmaf( hix, idir.x, -ood.x ) ;
	const float y0 = fmaf( loy, idir.y, -ood.y ), y1 = fmaf( hiy, idir.y, -ood.y ) ;
	const float z0 = fmaf( loz, idir.z, -ood.z ), z1 = fmaf( hiz, idir.z, -ood.z ) ;
	const float tn = fmaxf( fmaxf( fminf( x0, x1 ), fminf( y0, y1 ) ), fmaxf( fminf( z0, z1 ), tmin ) ) ;
	const float tf = fminf( fminf( fmaxf( x0, x1 ), fmaxf( y0, y1 ) ), fminf( fmaxf( z0, z1 ), tbest_s ) ) ;
	return tn<=tf*RTX_SLACK ? tn : INFINITY ;
#endif
}

#ifndef RTX_W8_ORDER
#define RTX_W8_ORDER 0
#endif
#if RTX_WIDTH == 8
// the eight (entry distance, child) pairs of a node, nearest first (misses: +inf, last)
RTX_HD void node_children8( const q4* n, const f3& idir, const f3& ood, float tmin, float tbest_s, float t[8], int32_t c[8] ) {
#pragma unroll
	for ( int h = 0 ; h<2 ; h++ ) {
		const o8 n01 = ldo( n+8*h ), n23 = ldo( n+8*h+2 ), n45 = ldo( n+8*h+4 ), n67 = ldo( n+8*h+6 ) ;
		const q4 lx = n01.a, ly = n01.b, lz = n23.a, hx = n23.b, hy = n45.a, hz = n45.b, rf = n67.a ;
		c[4*h] = asint( rf.x ) ; c[4*h+1] = asint( rf.y ) ; c[4*h+2] = asint( rf.z ) ; c[4*h+3] = asint( rf.w ) ;
		t[4*h]   = slab( lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, idir, ood, tmin, tbest_s ) ;
		t[4*h+1] = slab( lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, idir, ood, tmin, tbest_s ) ;
		t[4*h+2] = slab( lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, idir, ood, tmin, tbest_s ) ;
		t[4*h+3] = slab( lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, idir, ood, tmin, tbest_s ) ;
	}
#define RTX_CS( i, j ) if ( t[j]<t[i] ) { const float tt = t[i] ; t[i] = t[j] ; t[j] = tt ; const int32_t cc = c[i] ; c[i] = c[j] ; c[j] = cc ; }
#if RTX_W8_ORDER == 1
	// cheaper: each half sorted (5 comparators), then the half with the nearer first child in front
	RTX_CS( 0, 1 ) RTX_CS( 2, 3 ) RTX_CS( 0, 2 ) RTX_CS( 1, 3 ) RTX_CS( 1, 2 )
	RTX_CS( 4, 5 ) RTX_CS( 6, 7 ) RTX_CS( 4, 6 ) RTX_CS( 5, 7 ) RTX_CS( 5, 6 )
	if ( t[4]<t[0] ) {
#pragma unroll
		for ( int k = 0 ; k<4 ; k++ ) { const float tt = t[k] ; t[k] = t[4+k] ; t[4+k] = tt ; const int32_t cc = c[k] ; c[k] = c[4+k] ; c[4+k] = cc ; }
	}
	// (callers push 7..1: a miss inside the front half must not hide the hits behind it)
#else
	// 19-comparator network for eight keys
	RTX_CS( 0, 1 ) RTX_CS( 2, 3 ) RTX_CS( 4, 5 ) RTX_CS( 6, 7 )
	RTX_CS( 0, 2 ) RTX_CS( 1, 3 ) RTX_CS( 4, 6 ) RTX_CS( 5, 7 )
	RTX_CS( 1, 2 ) RTX_CS( 5, 6 ) RTX_CS( 0, 4 ) RTX_CS( 3, 7 )
	RTX_CS( 1, 5 ) RTX_CS( 2, 6 )
	RTX_CS( 1, 4 ) RTX_CS( 3, 6 )
	RTX_CS( 2, 4 ) RTX_CS( 3, 5 )
	RTX_CS( 3, 4 )
#endif
#undef RTX_CS
}
#endif

#if defined( __CUDA_ARCH__ )
// The same test for the four children of a node with packed single-precision FMAs (sm_100:
// FFMA2, two lanes of one 64-bit register pair per instruction, scalar operands broadcast):
// 12 instead of 24 multiply-adds per node step, identical values (each half is an IEEE fma).
__device__ __forceinline__ void fma2( float a0, float a1, float s, float c, float& r0, float& r1 ) {
	float2 av = make_float2( a0, a1 ), sv = make_float2( s, s ), cv = make_float2( -c, -c ) ;
	unsigned long long r ;
	asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( r ) : "l"( *reinterpret_cast<unsigned long long*>( &av ) ), "l"( *reinterpret_cast<unsigned long long*>( &sv ) ), "l"( *reinterpret_cast<unsigned long long*>( &cv ) ) ) ;
	const float2 rv = *reinterpret_cast<float2*>( &r ) ;
	r0 = rv.x ; r1 = rv.y ;
}
// (a0,a1)*s + (c0,c1)
__device__ __forceinline__ void fma2p( float a0, float a1, float s, float c0, float c1, float& r0, float& r1 ) {
	float2 av = make_float2( a0, a1 ), sv = make_float2( s, s ), cv = make_float2( c0, c1 ) ;
	unsigned long long r ;
	asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( r ) : "l"( *reinterpret_cast<unsigned long long*>( &av ) ), "l"( *reinterpret_cast<unsigned long long*>( &sv ) ), "l"( *reinterpret_cast<unsigned long long*>( &cv ) ) ) ;
	const float2 rv = *reinterpret_cast<float2*>( &r ) ;
	r0 = rv.x ; r1 = rv.y ;
}
__device__ __forceinline__ void slab4( const q4& lx, const q4& ly, const q4& lz, const q4& hx, const q4& hy, const q4& hz, const f3& idir, const f3& ood, float tmin, float tbest_s,
		float& t0, float& t1, float& t2, float& t3 ) {
	float x0[4], x1[4], y0[4], y1[4], z0[4], z1[4] ;
#if RTX_NODE_CH
	// centre / half-extent records: per axis t = c d' - o d', near = t - h|d'|, far = t + h|d'|
	float tx[4], ty[4], tz[4] ;
	fma2( lx.x, lx.y, idir.x, ood.x, tx[0], tx[1] ) ; fma2( lx.z, lx.w, idir.x, ood.x, tx[2], tx[3] ) ;
	fma2( ly.x, ly.y, idir.y, ood.y, ty[0], ty[1] ) ; fma2( ly.z, ly.w, idir.y, ood.y, ty[2], ty[3] ) ;
	fma2( lz.x, lz.y, idir.z, ood.z, tz[0], tz[1] ) ; fma2( lz.z, lz.w, idir.z, ood.z, tz[2], tz[3] ) ;
	const float ax = fabsf( idir.x ), ay = fabsf( idir.y ), az = fabsf( idir.z ) ;
	fma2p( hx.x, hx.y, -ax, tx[0], tx[1], x0[0], x0[1] ) ; fma2p( hx.z, hx.w, -ax, tx[2], tx[3], x0[2], x0[3] ) ;
	fma2p( hx.x, hx.y,  ax, tx[0], tx[1], x1[0], x1[1] ) ; fma2p( hx.z, hx.w,  ax, tx[2], tx[3], x1[2], x1[3] ) ;
	fma2p( hy.x, hy.y, -ay, ty[0], ty[1], y0[0], y0[1] ) ; fma2p( hy.z, hy.w, -ay, ty[2], ty[3], y0[2], y0[3] ) ;
	fma2p( hy.x, hy.y,  ay, ty[0], ty[1], y1[0], y1[1] ) ; fma2p( hy.z, hy.w,  ay, ty[2], ty[3], y1[2], y1[3] ) ;
	fma2p( hz.x, hz.y, -az, tz[0], tz[1], z0[0], z0[1] ) ; fma2p( hz.z, hz.w, -az, tz[2], tz[3], z0[2], z0[3] ) ;
	fma2p( hz.x, hz.y,  az, tz[0], tz[1], z1[0], z1[1] ) ; fma2p( hz.z, hz.w,  az, tz[2], tz[3], z1[2], z1[3] ) ;
	float t[4] ;
#pragma unroll
	for ( int k = 0 ; k<4 ; k++ ) {
		const float tn = fmaxf( fmaxf( x0[k], y0[k] ), fmaxf( z0[k], tmin ) ) ;
		const float tf = fminf( fminf( x1[k], y1[k] ), fminf( z1[k], tbest_s ) ) ;
		t[k] = tn<=tf*RTX_SLACK ? tn : INFINITY ;
	}
#else
	fma2( lx.x, lx.y, idir.x, ood.x, x0[0], x0[1] ) ; fma2( lx.z, lx.w, idir.x, ood.x, x0[2], x0[3] ) ;
	fma2( hx.x, hx.y, idir.x, ood.x, x1[0], x1[1] ) ; fma2( hx.z, hx.w, idir.x, ood.x, x1[2], x1[3] ) ;
	fma2( ly.x, ly.y, idir.y, ood.y, y0[0], y0[1] ) ; fma2( ly.z, ly.w, idir.y, ood.y, y0[2], y0[3] ) ;
	fma2( hy.x, hy.y, idir.y, ood.y, y1[0], y1[1] ) ; fma2( hy.z, hy.w, idir.y, ood.y, y1[2], y1[3] ) ;
	fma2( lz.x, lz.y, idir.z, ood.z, z0[0], z0[1] ) ; fma2( lz.z, lz.w, idir.z, ood.z, z0[2], z0[3] ) ;
	fma2( hz.x, hz.y, idir.z, ood.z, z1[0], z1[1] ) ; fma2( hz.z, hz.w, idir.z, ood.z, z1[2], z1[3] ) ;
	float t[4] ;
#pragma unroll
	for ( int k = 0 ; k<4 ; k++ ) {
		const float tn = fmaxf( fmaxf( fminf( x0[k], x1[k] ), fminf( y0[k], y1[k] ) ), fmaxf( fminf( z0[k], z1[k] ), tmin ) ) ;
		const float tf = fminf( fminf( fmaxf( x0[k], x1[k] ), fmaxf( y0[k], y1[k] ) ), fminf( fmaxf( z0[k], z1[k] ), tbest_s ) ) ;
		t[k] = tn<=tf*RTX_SLACK ? tn : INFINITY ;
	}
#endif
	t0 = t[0] ; t1 = t[1] ; t2 = t[2] ; t3 = t[3] ;
}
#endif

// ----------------------------------------------------------------------------- traversal
// "while-while" over the 4-wide trees, warp-synchronous: all lanes of the warp first walk
// inner nodes until none of them holds one (lanes that reached a leaf wait), then every lane
// with a leaf processes it.  The loops are driven by warp votes instead of per-lane exits:
// with independent thread scheduling, lanes that leave a loop at different times are not
// brought back together, and the warp degenerates into small groups executing the same
// code one after the other (measured: 6.9 of 32 lanes active).  `active` = this lane has
// a ray; idle lanes tag along.  One stack serves both levels: entering a mesh pushes
// RTX_STK_RETURN, popping it restores the world-space ray; RTX_STK_DONE sits at the bottom.
template <class Stack>
RTX_HD void closest( const SceneDev& S, const f3& o, const f3& d, float tmin, Stack& st, HitRec& best, bool active = true ) {
	best.t = INFINITY ; best.thing = -1 ; best.prim = -1 ; best.u = 0.f ; best.v = 0.f ; best.slot = 0 ;
	if ( S.n_things == 0 )   // uniform over the launch
		return ;
	float tbest_s = INFINITY ;

	// current-level ray
	f3 idir = mk3( safe_rcp( d.x ), safe_rcp( d.y ), safe_rcp( d.z ) ) ;
	f3 ood  = mk3( o.x*idir.x, o.y*idir.y, o.z*idir.z ) ;
	const q4* nodes = S.tlas_nodes ;
	// mesh-level state
	f3 ohi = o, olo = mk3( 0.f, 0.f, 0.f ), dd = d ;
	const q4* tris = nullptr ;
	int32_t thing = -1 ;

	st.reset() ;
	st.push( RTX_STK_DONE ) ;
	int32_t cur = active ? 0 : RTX_STK_DONE ;   // 0 = root
	RTX_COUNT( rays ) ;
	while ( true ) {
		// ---- phase 1: inner nodes
		while ( RTX_ANY( uint32_t( cur )<uint32_t( RTX_REF_EMPTY ) ) ) {
			if ( uint32_t( cur )<uint32_t( RTX_REF_EMPTY ) ) {
				const q4* n = nodes+size_t( cur )*RTX_NODE_RECS ;
				RTX_COUNT( nodes ) ; RTX_EVENT( 'N' ) ;
#if RTX_WIDTH == 8
				float t8[8] ; int32_t c8[8] ;
				node_children8( n, idir, ood, tmin, tbest_s, t8, c8 ) ;
				if ( t8[0] == INFINITY )
					cur = st.pop() ;
				else {
#pragma unroll
					for ( int k = 7 ; k>0 ; k-- ) if ( t8[k]<INFINITY ) st.push( c8[k] ) ;
					cur = c8[0] ;
				}
				continue ;
#endif
				const o8 n01 = ldo( n ), n23 = ldo( n+2 ), n45 = ldo( n+4 ), n67 = ldo( n+6 ) ;
				const q4 lx = n01.a, ly = n01.b, lz = n23.a, hx = n23.b, hy = n45.a, hz = n45.b, rf = n67.a ;
				int32_t c0 = asint( rf.x ), c1 = asint( rf.y ), c2 = asint( rf.z ), c3 = asint( rf.w ) ;
				float t0 = slab( lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, idir, ood, tmin, tbest_s ) ;
				float t1 = slab( lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, idir, ood, tmin, tbest_s ) ;
				float t2 = slab( lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, idir, ood, tmin, tbest_s ) ;
				float t3 = slab( lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, idir, ood, tmin, tbest_s ) ;
				// sort the four (entry distance, child) pairs, nearest first
#define RTX_CSWAP( ta, ca, tb, cb ) if ( tb<ta ) { const float tt = ta ; ta = tb ; tb = tt ; const int32_t cc = ca ; ca = cb ; cb = cc ; }
				RTX_CSWAP( t0, c0, t1, c1 ) RTX_CSWAP( t2, c2, t3, c3 ) RTX_CSWAP( t0, c0, t2, c2 ) RTX_CSWAP( t1, c1, t3, c3 ) RTX_CSWAP( t1, c1, t2, c2 )
#undef RTX_CSWAP
				if ( t0 == INFINITY )
					cur = st.pop() ;
				else {
					if ( t3<INFINITY ) st.push( c3 ) ;
					if ( t2<INFINITY ) st.push( c2 ) ;
					if ( t1<INFINITY ) st.push( c1 ) ;
					cur = c0 ;
				}
			}
		}
		if ( ! RTX_ANY( cur != RTX_STK_DONE ) )
			break ;
		// ---- phase 2: a leaf, or the return to the top level
		if ( cur == RTX_STK_RETURN ) {
			idir = mk3( safe_rcp( d.x ), safe_rcp( d.y ), safe_rcp( d.z ) ) ;
			ood  = mk3( o.x*idir.x, o.y*idir.y, o.z*idir.z ) ;
			nodes = S.tlas_nodes ; tris = nullptr ; thing = -1 ;
			cur = st.pop() ;
			RTX_EVENT( 'R' ) ;
		} else if ( cur != RTX_STK_DONE ) {
			const uint32_t ref = uint32_t( ~cur ) ;
			const uint32_t first = ref>>3, count = ( ref&7u )+1u ;
			cur = RTX_REF_EMPTY ;   // "pop next" unless a mesh is entered
			if ( tris ) {
				RTX_COUNT( leaves ) ; RTX_EVENT( char( '0'+count ) ) ;
				for ( uint32_t k = 0 ; k<count ; k++ ) {
					RTX_COUNT( tris ) ;
					const q4* T = tris+size_t( first+k )*RTX_TRI_RECS ;
					const o8 t01 = ldo_tri( T ), t23 = ldo_tri( T+2 ) ;
					const q4 a = t01.a, b = t01.b, c = t23.a ;
					float t, u, v ;
					if ( tri_test( mk3( a.x, a.y, a.z ), mk3( b.x, b.y, b.z ), mk3( c.x, c.y, c.z ), ohi, olo, dd, tmin, t, u, v ) ) {
						const int32_t prim = asint( a.w ) ;
						if ( better( t, thing, prim, best ) ) {
							best.t = t ; best.thing = thing ; best.prim = prim ; best.u = u ; best.v = v ; best.slot = int32_t( first+k ) ;
							tbest_s = t*RTX_SLACK ;
							RTX_EVENT( 'H' ) ;
						}
					}
				}
			} else {
				// top level: one thing per leaf
				const int32_t k = int32_t( RTX_LDG( S.tlas_order+first ) ) ;
				const ThingTrav* tt = S.trav+k ;
				RTX_COUNT( things ) ;
				if ( bsphere_miss( ldq( S.bsphere+k ), o, d, tmin, best.t ) ) {
					cur = st.pop() ;
					continue ;
				}
				const double m0 = RTX_LDG( tt->inv+0 ), m1 = RTX_LDG( tt->inv+1 ), m2 = RTX_LDG( tt->inv+2 ), m3 = RTX_LDG( tt->inv+3 ) ;
				if ( RTX_LDG( &tt->kind ) == 0 ) {
					double td ;
					RTX_COUNT( spheres ) ; RTX_EVENT( 'P' ) ;
					if ( sphere_root( mk3( m0, m1, m2 ), m3, wide( o ), wide( d ), double( tmin ), td ) ) {
						const float t = float( td ) ;
						if ( better( t, k, -1, best ) ) {
							best.t = t ; best.thing = k ; best.prim = -1 ;
							tbest_s = t*RTX_SLACK ;
						}
					}
				} else {
					// enter the mesh: object-space ray, origin in double carried as hi+lo
					d3 od, ddd ;
					if ( RTX_LDG( &tt->diag ) ) {
						const double m5 = RTX_LDG( tt->inv+5 ), m7 = RTX_LDG( tt->inv+7 ), m10 = RTX_LDG( tt->inv+10 ), m11 = RTX_LDG( tt->inv+11 ) ;
						od = xfpoint_diag( m0, m3, m5, m7, m10, m11, wide( o ) ) ;
						ddd = xfvec_diag( m0, m5, m10, wide( d ) ) ;
					} else {
						double m[12] ;
						m[0] = m0 ; m[1] = m1 ; m[2] = m2 ; m[3] = m3 ;
						for ( int j = 4 ; j<12 ; j++ ) m[j] = RTX_LDG( tt->inv+j ) ;
						od = xfpoint( m, wide( o ) ) ;
						ddd = xfvec( m, wide( d ) ) ;
					}
					ohi = narrow( od ) ;
					olo = narrow( od-wide( ohi ) ) ;
					dd  = narrow( ddd ) ;
					idir = mk3( safe_rcp( dd.x ), safe_rcp( dd.y ), safe_rcp( dd.z ) ) ;
					ood  = mk3( ohi.x*idir.x, ohi.y*idir.y, ohi.z*idir.z ) ;
					nodes = ldptr( &tt->nodes ) ;
					tris  = ldptr( &tt->tris ) ;
					thing = k ;
					st.push( RTX_STK_RETURN ) ;
					cur = 0 ;
					RTX_COUNT( enters ) ; RTX_EVENT( k == 0 ? 'G' : 'E' ) ;
				}
			}
			if ( cur == RTX_REF_EMPTY )
				cur = st.pop() ;
		}
	}
}

// exhaustive scan (validation instrument; same tests, same tie rule, no boxes)
RTX_HD void closest_brute( const SceneDev& S, const f3& o, const f3& d, float tmin, HitRec& best ) {
	best.t = INFINITY ; best.thing = -1 ; best.prim = -1 ; best.u = 0.f ; best.v = 0.f ; best.slot = 0 ;
	for ( uint32_t k = 0 ; k<S.n_things ; k++ ) {
		const ThingTrav* tt = S.trav+k ;
		if ( tt->kind == 0 ) {
			double td ;
			if ( sphere_root( mk3( tt->inv[0], tt->inv[1], tt->inv[2] ), tt->inv[3], wide( o ), wide( d ), double( tmin ), td ) ) {
				const float t = float( td ) ;
				if ( better( t, int32_t( k ), -1, best ) ) { best.t = t ; best.thing = int32_t( k ) ; best.prim = -1 ; }
			}
		} else {
			const d3 od = xfpoint( tt->inv, wide( o ) ) ;
			const f3 ohi = narrow( od ) ;
			const f3 olo = narrow( od-wide( ohi ) ) ;
			const f3 dd  = narrow( xfvec( tt->inv, wide( d ) ) ) ;
			for ( uint32_t f = 0 ; f<tt->n_tris ; f++ ) {
				const q4* T = tt->tris+size_t( f )*RTX_TRI_RECS ;
				const q4 a = ldq( T ), b = ldq( T+1 ), c = ldq( T+2 ) ;
				float t, u, v ;
				if ( tri_test( mk3( a.x, a.y, a.z ), mk3( b.x, b.y, b.z ), mk3( c.x, c.y, c.z ), ohi, olo, dd, tmin, t, u, v ) ) {
					const int32_t prim = asint( a.w ) ;
					if ( better( t, int32_t( k ), prim, best ) ) { best.t = t ; best.thing = int32_t( k ) ; best.prim = prim ; best.u = u ; best.v = v ; best.slot = int32_t( f ) ; }
				}
			}
		}
	}
}

// ----------------------------------------------------------------------------- shading frame
// sphere.h:40-45 (analytic) / optx/optics_i.cu:40-82, :259-267 (mesh), in double,
// rounded once
RTX_HD_CALL void frame_of( const SceneDev& S, const HitRec& h, const f3& o, const f3& d, float tmin, Frame& fr ) {
	const ThingShade* ts = S.shade+h.thing ;
	double m[12] ;
	for ( int j = 0 ; j<12 ; j++ ) m[j] = RTX_LDG( ts->xf+j ) ;
	const d3 dw = wide( d ) ;
	const d3 center = mk3( m[3], m[7], m[11] ) ;
	if ( h.prim<0 ) {
		const double r = m[0] ;
		double td = 0. ;
		sphere_root( center, r, wide( o ), dw, double( tmin ), td ) ;
		const d3 p = wide( o )+td*dw ;
		const d3 outward = ( 1./r )*( p-center ) ;
		fr.p = narrow( p ) ;
		fr.facing = 0.>dot( dw, outward ) ;
		fr.normal = narrow( fr.facing ? outward : -outward ) ;
		return ;
	}
	// the three vertices as uploaded ride in the triangle record the traversal just tested
	const ThingTrav* tt = S.trav+h.thing ;
	const q4* T = ldptr( &tt->tris )+size_t( h.slot )*RTX_TRI_RECS ;
	const q4 t0 = ldq( T ), t1 = ldq( T+1 ), t2 = ldq( T+2 ), t3 = ldq( T+3 ) ;
	const d3 a = mk3( double( t0.x ), double( t0.y ), double( t0.z ) ) ;
	const d3 b = mk3( double( t1.w ), double( t2.w ), double( t3.x ) ) ;
	const d3 c = mk3( double( t3.y ), double( t3.z ), double( t3.w ) ) ;
	d3 A, B, C ;
	if ( RTX_LDG( &ts->diag ) ) {
		A = xfpoint_diag( m[0], m[3], m[5], m[7], m[10], m[11], a ) ;
		B = xfpoint_diag( m[0], m[3], m[5], m[7], m[10], m[11], b ) ;
		C = xfpoint_diag( m[0], m[3], m[5], m[7], m[10], m[11], c ) ;
	} else {
		A = xfpoint( m, a ) ; B = xfpoint( m, b ) ; C = xfpoint( m, c ) ;
	}
	const float w = 1.f-h.u-h.v ;
	const d3 p = double( w )*A+double( h.u )*B+double( h.v )*C ;
	d3 N = unitV( cross( B-A, C-A ) ) ;
	if ( dot( dw, N )>0. )
		N = -N ;
	fr.p = narrow( p ) ;
	fr.normal = narrow( N ) ;
	fr.facing = 0.>dot( dw, p-center ) ;
}

// ----------------------------------------------------------------------------- scattering
RTX_HD f3 reflect( const f3& v, const f3& n ) { return v-( 2.f*dot( v, n ) )*n ; }                                   // v.h:61
RTX_HD f3 refract( const f3& v, const f3& n, float ratio ) {                                                        // v.h:62
	const float theta = fminf( dot( -v, n ), 1.f ) ;
	const f3 perpen = ratio*( v+theta*n ) ;
	const f3 parall = ( -sqrtf( fabsf( 1.f-dot( perpen, perpen ) ) ) )*n ;
	return perpen+parall ;
}
RTX_HD float schlick( float cos_theta, float ratio ) {                                                              // optics.h:74
	float r0 = ( 1.f-ratio )/( 1.f+ratio ) ; r0 = r0*r0 ;
	const float m = 1.f-cos_theta ;
	const float m2 = m*m ;
	return r0+( 1.f-r0 )*( ( m2*m2 )*m ) ;
}

// optics.h:15-24 (Diffuse), :34-40 (Reflect), :51-68 (Refract).  Returns false when
// the path is absorbed.
RTX_HD_CALL bool scatter( const ThingShade* ts, const f3& dir, const Frame& fr, Pcg& rng, f3& attened, f3& out, bool guard = true ) {
	const int32_t type = RTX_LDG( &ts->type ) ;
	if ( type != 2 ) {
		// Diffuse and Reflect both start with one rndVin1sphere(): the lanes of a warp run that
		// rejection loop -- the longest piece of shading, and as long as its slowest lane -- together
		// instead of once per material branch.  Per path the draws and their order are unchanged.
		const f3 s = rng.rndVin1sphere() ;
		attened = mk3( RTX_LDG( ts->albedo ), RTX_LDG( ts->albedo+1 ), RTX_LDG( ts->albedo+2 ) ) ;
		if ( type == 0 ) {
			f3 dnew = fr.normal+unitV( s ) ;
			if ( guard && fabsf( dnew.x )<1e-8f && fabsf( dnew.y )<1e-8f && fabsf( dnew.z )<1e-8f )   // util.h:8 kNear0 (optx/optics_i.cu:86 has no guard)
				dnew = fr.normal ;
			out = dnew ;
			return true ;
		}
		const f3 r = reflect( unitV( dir ), fr.normal ) ;
		out = r+RTX_LDG( &ts->fuzz )*s ;
		return dot( out, fr.normal )>0.f ;
	}
	const f3 d1V = unitV( dir ) ;
	const float cos_theta = fminf( dot( -d1V, fr.normal ), 1.f ) ;
	const float sin_theta = sqrtf( 1.f-cos_theta*cos_theta ) ;
	const float index = RTX_LDG( &ts->index ) ;
	const float ratio = fr.facing ? 1.f/index : index ;
	const bool cannot = ratio*sin_theta>1.f ;
	if ( cannot || schlick( cos_theta, ratio )>rng.rnd() )
		out = reflect( d1V, fr.normal ) ;
	else
		out = refract( d1V, fr.normal, ratio ) ;
	attened = mk3( 1.f, 1.f, 1.f ) ;
	return true ;
}

RTX_HD q4 mkq( float x, float y, float z, float w ) { q4 r ; r.x = x ; r.y = y ; r.z = z ; r.w = w ; return r ; }
RTX_HD float ibits( int32_t i ) { return asfloat( i ) ; }
RTX_HD float ubits( uint32_t u ) { return asfloat( int32_t( u ) ) ; }
RTX_HD uint32_t bitsu( float f ) { return uint32_t( asint( f ) ) ; }
RTX_HD double dbl( float lo, float hi ) {
#if defined( __CUDA_ARCH__ )
	return __hiloint2double( __float_as_int( hi ), __float_as_int( lo ) ) ;
#else
	const uint64_t b = uint64_t( bitsu( lo ) )|( uint64_t( bitsu( hi ) )<<32 ) ;
	double d ; memcpy( &d, &b, 8 ) ; return d ;
#endif
}

// 32 bytes of a thing record with one 256-bit load (the records are 16-byte aligned arrays of
// 128 / 144 bytes whose 32-byte pieces never straddle the alignment a 256-bit load needs: the
// arrays come from cudaMalloc, 256-byte aligned, and both record sizes are multiples of 16 --
// so pieces at offsets that are multiples of 32 within a 32-byte aligned record are fine for
// ThingTrav (128 bytes); ThingShade (144 bytes) is read with 128-bit loads)
RTX_HD o8 ldo_rec( const void* p ) {
#if defined( __CUDA_ARCH__ )
	o8 r ;
	asm( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"( r.a.x ), "=f"( r.a.y ), "=f"( r.a.z ), "=f"( r.a.w ), "=f"( r.b.x ), "=f"( r.b.y ), "=f"( r.b.z ), "=f"( r.b.w ) : "l"( p ) ) ;
	return r ;
#else
	o8 r ; memcpy( &r, p, 32 ) ; return r ;
#endif
}
RTX_HD q4 ldq_rec( const void* p ) {
#if defined( __CUDA_ARCH__ )
	const float4 v = __ldg( reinterpret_cast<const float4*>( p ) ) ;
	return mkq( v.x, v.y, v.z, v.w ) ;
#else
	q4 r ; memcpy( &r, p, 16 ) ; return r ;
#endif
}

// ---- shading without calls and without local memory (the __noinline__ frame_of / scatter of
// rtx_core.cuh pass their structures through the stack): the same expressions in the same order
// -- sphere.h:40-45 / optx/optics_i.cu:40-82, :259-267 and optics.h:15-24, :34-40, :51-68 -- on
// thing records fetched with 256-bit loads.  `mat` returns albedo.rgb, fuzz | index, type, kind, diag.
RTX_HD void qframe_of( const SceneDev& S, int32_t thing, int32_t tslot, float u, float v, const f3& o, const f3& d, float tmin, Frame& fr, o8& mat ) {
	const char* ts = reinterpret_cast<const char*>( S.shade+thing ) ;
	const o8 X0 = ldo_rec( ts ), X1 = ldo_rec( ts+32 ), X2 = ldo_rec( ts+64 ) ;
	mat = ldo_rec( ts+96 ) ;
	const double m0 = dbl( X0.a.x, X0.a.y ), m3 = dbl( X0.b.z, X0.b.w ), m5 = dbl( X1.a.z, X1.a.w ), m7 = dbl( X1.b.z, X1.b.w ), m10 = dbl( X2.b.x, X2.b.y ), m11 = dbl( X2.b.z, X2.b.w ) ;
	const d3 dw = wide( d ) ;
	const d3 center = mk3( m3, m7, m11 ) ;
	if ( tslot<0 ) {
		const double r = m0 ;
		double td = 0. ;
		sphere_root( center, r, wide( o ), dw, double( tmin ), td ) ;
		const d3 pp = wide( o )+td*dw ;
		const d3 outward = ( 1./r )*( pp-center ) ;
		fr.p = narrow( pp ) ;
		fr.facing = 0.>dot( dw, outward ) ;
		fr.normal = narrow( fr.facing ? outward : -outward ) ;
		return ;
	}
	// the three vertices as uploaded ride in the triangle record the traversal tested
#if defined( __CUDA_ARCH__ )
	const o8 D = ldo_rec( reinterpret_cast<const char*>( S.trav+thing )+96 ) ;
	const q4* tris = reinterpret_cast<const q4*>( ( unsigned long long )__float_as_uint( D.a.z )|( ( unsigned long long )__float_as_uint( D.a.w )<<32 ) ) ;
#else
	const q4* tris = S.trav[thing].tris ;
#endif
	const q4* T = tris+size_t( tslot )*RTX_TRI_RECS ;
	const o8 t01 = ldo_tri( T ), t23 = ldo_tri( T+2 ) ;
	const q4 t0 = t01.a, t1 = t01.b, t2 = t23.a, t3 = t23.b ;
	const d3 a = mk3( double( t0.x ), double( t0.y ), double( t0.z ) ) ;
	const d3 b = mk3( double( t1.w ), double( t2.w ), double( t3.x ) ) ;
	const d3 c = mk3( double( t3.y ), double( t3.z ), double( t3.w ) ) ;
	d3 A, B, C ;
	if ( asint( mat.b.w ) ) {
		A = xfpoint_diag( m0, m3, m5, m7, m10, m11, a ) ;
		B = xfpoint_diag( m0, m3, m5, m7, m10, m11, b ) ;
		C = xfpoint_diag( m0, m3, m5, m7, m10, m11, c ) ;
	} else {
		double m[12] ;
		m[0] = m0 ; m[1] = dbl( X0.a.z, X0.a.w ) ; m[2] = dbl( X0.b.x, X0.b.y ) ; m[3] = m3 ;
		m[4] = dbl( X1.a.x, X1.a.y ) ; m[5] = m5 ; m[6] = dbl( X1.b.x, X1.b.y ) ; m[7] = m7 ;
		m[8] = dbl( X2.a.x, X2.a.y ) ; m[9] = dbl( X2.a.z, X2.a.w ) ; m[10] = m10 ; m[11] = m11 ;
		A = xfpoint( m, a ) ; B = xfpoint( m, b ) ; C = xfpoint( m, c ) ;
	}
	const float w = 1.f-u-v ;
	const d3 pp = double( w )*A+double( u )*B+double( v )*C ;
	d3 N = unitV( cross( B-A, C-A ) ) ;
	if ( dot( dw, N )>0. )
		N = -N ;
	fr.p = narrow( pp ) ;
	fr.normal = narrow( N ) ;
	fr.facing = 0.>dot( dw, pp-center ) ;
}

RTX_HD bool qscatter( const o8& mat, const f3& dir, const Frame& fr, Pcg& rng, f3& attened, f3& out, bool guard ) {
	const int32_t type = asint( mat.b.y ) ;
	if ( type != 2 ) {
		const f3 s = rng.rndVin1sphere() ;
		attened = mk3( mat.a.x, mat.a.y, mat.a.z ) ;
		if ( type == 0 ) {
			f3 dnew = fr.normal+unitV( s ) ;
			if ( guard && fabsf( dnew.x )<1e-8f && fabsf( dnew.y )<1e-8f && fabsf( dnew.z )<1e-8f )   // util.h:8 kNear0 (optx/optics_i.cu:86 has no guard)
				dnew = fr.normal ;
			out = dnew ;
			return true ;
		}
		const f3 r = reflect( unitV( dir ), fr.normal ) ;
		out = r+mat.a.w*s ;
		return dot( out, fr.normal )>0.f ;
	}
	const f3 d1V = unitV( dir ) ;
	const float cos_theta = fminf( dot( -d1V, fr.normal ), 1.f ) ;
	const float sin_theta = sqrtf( 1.f-cos_theta*cos_theta ) ;
	const float index = mat.b.x ;
	const float ratio = fr.facing ? 1.f/index : index ;
	const bool cannot = ratio*sin_theta>1.f ;
	if ( cannot || schlick( cos_theta, ratio )>rng.rnd() )
		out = reflect( d1V, fr.normal ) ;
	else
		out = refract( d1V, fr.normal, ratio ) ;
	attened = mk3( 1.f, 1.f, 1.f ) ;
	return true ;
}


// rtow.cxx:45-48
RTX_HD f3 sky( const f3& dir ) {
	const f3 unit = unitV( dir ) ;
	const float t = .5f*( unit.y+1.f ) ;
	return ( 1.f-t )*mk3( 1.f, 1.f, 1.f )+t*mk3( .5f, .7f, 1.f ) ;
}

// rtow.cxx:112-113 + camera.h:25-31
// (optx/camera_i.cu:61-62 divides by w and h: `whole`)
RTX_HD_CALL void primary_ray( const CameraDev& cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h, Pcg& rng, f3& ori, f3& dir, bool whole = false ) {
	const float s = 2.f*( float( x )+rng.rnd() )/float( whole ? w : w-1 )-1.f ;
	const float t = 2.f*( float( y )+rng.rnd() )/float( whole ? h : h-1 )-1.f ;
	const f3 r = ( cam.aperture/2.f )*rng.rndVin1disk() ;
	const f3 o = r.x*cam.u+r.y*cam.v ;
	ori = cam.eye+o ;
	dir = s*cam.wvec+t*cam.hvec-cam.dvec-o ;
}

// a path colour channel in [0,1] -> 2^-32 fixed point (integer sums are associative)
RTX_HD uint64_t tofix( float c ) { return uint64_t( c*4294967296.f ) ; }

// scatter events a path may take before its next hit ends it: `depth` for rtow.cxx (rtow.cxx:39-42
// tests before it intersects); the OptiX programs count rays, the depth-th hit is the last
RTX_HD uint32_t sem_depth( uint32_t variant, uint32_t depth ) {
	return variant == RTX_SEM_RTOW ? depth : ( depth>1u ? depth-1u : 0u ) ;
}

// One whole path (rtow.cxx:34-49 unrolled into a loop, throughput front to back).
// Used by the host harness and the picker; the render kernel inlines the same steps
// in its regenerating loop.
template <class Stack>
RTX_HD f3 path_radiance( const SceneDev& S, f3 ori, f3 dir, uint32_t depth, Pcg& rng, Stack& st, uint32_t& segments ) {
	f3 thr = mk3( 1.f, 1.f, 1.f ) ;
	depth = sem_depth( S.variant, depth ) ;
	while ( true ) {
		HitRec h ;
		segments++ ;
		closest( S, ori, dir, 1e-3f, st, h ) ;                       // util.h:9 kAcne0
		if ( h.thing<0 )
			return thr*sky( dir ) ;
		if ( depth == 0 && S.variant != RTX_SEM_RTWO_I )
			return mk3( 0.f, 0.f, 0.f ) ;
		Frame fr ;
		frame_of( S, h, ori, dir, 1e-3f, fr ) ;
		f3 att, out ;
		if ( ! scatter( S.shade+h.thing, dir, fr, rng, att, out, S.variant == RTX_SEM_RTOW ) )
			return mk3( 0.f, 0.f, 0.f ) ;
		thr = thr*att ;
		if ( depth == 0 )
			return thr ;                                             // optx/camera_i.cu:92-95
		ori = fr.p ; dir = out ; depth-- ;
	}
}

} // namespace rtx
