// rtx_kernels.cuh -- the frame kernels: path tracing (replaces __raygen__camera,
// __miss__ambient and the three __closesthit__ programs, optx/camera_i.cu:24-141 and
// optx/optics_i.cu:23-288, plus OptiX's traversal), the resolve (mean + clamp of
// optx/camera_i.cu:105 applied to the fixed-point sums), the post-processing pair of
// optx/postproc.cu, and the parity instruments.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rtx_core.cuh"

namespace rtx {

#define RTX_BLOCK      128   // threads per CTA of the tracing kernels (4 warps)
#define RTX_SM_STACK   24    // traversal stack entries per thread kept in shared memory
#define RTX_OVF_STACK  72    // further entries in local memory (LBVH worst case: 63 key bits + ties, two levels)

// Traversal stack: slot i of thread t lives at smem[i*RTX_BLOCK+t] (bank = t mod 32, so
// a warp's pushes and pops never conflict); deeper entries spill to local memory.
struct DevStack {
	int32_t* base ;
	int32_t  sp ;
	int32_t  ovf[RTX_OVF_STACK] ;
	__device__ __forceinline__ void reset() { sp = 0 ; }
	__device__ __forceinline__ bool empty() const { return sp == 0 ; }
	__device__ __forceinline__ void push( int32_t v ) {
		if ( sp<RTX_SM_STACK ) base[sp*RTX_BLOCK] = v ;
		else if ( sp<RTX_SM_STACK+RTX_OVF_STACK ) ovf[sp-RTX_SM_STACK] = v ;
		sp++ ;
	}
	__device__ __forceinline__ int32_t pop() {
		sp-- ;
		if ( sp<RTX_SM_STACK ) return base[sp*RTX_BLOCK] ;
		return sp<RTX_SM_STACK+RTX_OVF_STACK ? ovf[sp-RTX_SM_STACK] : RTX_STK_RETURN ;
	}
} ;

struct FrameArgs {
	SceneDev  S ;
	CameraDev cam ;
	uint32_t  w, h ;
	uint32_t  spp, depth ;
	uint64_t  seed ;
	uint32_t  sample0, sample_stride ;
	uint32_t  accumulate ;
	uint64_t* accum ;    // [4*w*h]  r, g, b fixed-point sums, segments
	int64_t*  hit_id ;   // [w*h]
	float*    hit_t ;    // [w*h]
} ;

// lane -> pixel: each warp owns an 8x4 tile (primary rays of a warp stay close)
__device__ __forceinline__ bool tile_pixel( uint32_t w, uint32_t h, uint32_t& x, uint32_t& y ) {
	const uint32_t warp = ( blockIdx.x*RTX_BLOCK+threadIdx.x )>>5 ;
	const uint32_t lane = threadIdx.x&31u ;
	const uint32_t tiles_x = ( w+7u )>>3 ;
	x = ( warp%tiles_x )*8u+( lane&7u ) ;
	y = ( warp/tiles_x )*4u+( lane>>3 ) ;
	return x<w && y<h ;
}
inline uint32_t tile_grid( uint32_t w, uint32_t h ) {
	const uint32_t warps = ( ( w+7u )>>3 )*( ( h+3u )>>2 ) ;
	return ( warps*32u+RTX_BLOCK-1u )/RTX_BLOCK ;
}

// The path tracer.  One lane = one pixel; the lane walks through its samples and starts
// the next path the moment the current one ends (no lane waits for the longest path of
// its warp, only for the warp's slowest pixel).  Colour is summed in 2^-32 fixed point.
__global__ void __launch_bounds__( RTX_BLOCK ) k_render( const FrameArgs a ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	uint32_t x, y ;
	if ( ! tile_pixel( a.w, a.h, x, y ) )
		return ;
	const uint32_t pix = a.w*y+x ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ;

	uint64_t acc0 = 0, acc1 = 0, acc2 = 0 ;
	uint32_t segments = 0 ;
	uint32_t k = 0, depth_left = 0 ;
	bool alive = false ;
	Pcg rng ; rng.state = 0 ;
	f3 ori = mk3( 0.f, 0.f, 0.f ), dir = mk3( 0.f, 0.f, 1.f ), thr = mk3( 1.f, 1.f, 1.f ) ;

	while ( true ) {
		if ( ! alive ) {
			if ( k>=a.spp )
				break ;
			rng.seed( a.seed, pix, a.sample0+k*a.sample_stride ) ;
			primary_ray( a.cam, x, y, a.w, a.h, rng, ori, dir ) ;
			thr = mk3( 1.f, 1.f, 1.f ) ;
			depth_left = a.depth ;
			alive = true ;
			k++ ;
		}
		HitRec hit ;
		closest( a.S, ori, dir, 1e-3f, st, hit ) ;
		segments++ ;
		f3 c = mk3( 0.f, 0.f, 0.f ) ;
		bool done = true ;
		if ( hit.thing<0 )
			c = thr*sky( dir ) ;
		else if ( depth_left>0 ) {
			Frame fr ;
			frame_of( a.S, hit, ori, dir, 1e-3f, fr ) ;
			f3 att, out ;
			if ( scatter( a.S.shade+hit.thing, dir, fr, rng, att, out ) ) {
				thr = thr*att ;
				ori = fr.p ; dir = out ; depth_left-- ;
				done = false ;
			}
		}
		if ( done ) {
			acc0 += tofix( c.x ) ; acc1 += tofix( c.y ) ; acc2 += tofix( c.z ) ;
			alive = false ;
		}
	}

	ulonglong2* out = reinterpret_cast<ulonglong2*>( a.accum+4*size_t( pix ) ) ;
	ulonglong2 lo = make_ulonglong2( acc0, acc1 ), hi = make_ulonglong2( acc2, uint64_t( segments ) ) ;
	if ( a.accumulate ) {
		const ulonglong2 plo = out[0], phi = out[1] ;
		lo.x += plo.x ; lo.y += plo.y ; hi.x += phi.x ; hi.y += phi.y ;
	}
	out[0] = lo ; out[1] = hi ;
}

// first hit of the primary ray of sample `sample0` of every pixel
__global__ void __launch_bounds__( RTX_BLOCK ) k_primary_hits( const FrameArgs a ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	uint32_t x, y ;
	if ( ! tile_pixel( a.w, a.h, x, y ) )
		return ;
	const uint32_t pix = a.w*y+x ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ;
	Pcg rng ;
	rng.seed( a.seed, pix, a.sample0 ) ;
	f3 ori, dir ;
	primary_ray( a.cam, x, y, a.w, a.h, rng, ori, dir ) ;
	HitRec hit ;
	closest( a.S, ori, dir, 1e-3f, st, hit ) ;
	a.hit_id[pix] = hit.thing<0 ? int64_t( -1 ) : ( ( int64_t( hit.thing )<<32 )|int64_t( uint32_t( hit.prim+1 ) ) ) ;
	a.hit_t[pix]  = hit.thing<0 ? -1.f : hit.t ;
}

// picker (optx/camera_i.cu:27-29, optx/optics_i.cu:25-29): one primary ray through
// (px,py), the thing id or UINT_MAX
__global__ void k_pick( const FrameArgs a, uint32_t px, uint32_t py, uint32_t* pick_id ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	if ( threadIdx.x != 0 || blockIdx.x != 0 )
		return ;
	DevStack st ;
	st.base = stack_mem ;
	Pcg rng ;
	rng.seed( a.seed, a.w*py+px, a.sample0 ) ;
	f3 ori, dir ;
	primary_ray( a.cam, px, py, a.w, a.h, rng, ori, dir ) ;
	HitRec hit ;
	closest( a.S, ori, dir, 1e-3f, st, hit ) ;
	*pick_id = hit.thing<0 ? 0xffffffffu : uint32_t( hit.thing ) ;
}

// closest hits of caller-supplied rays: through the LBVH, or by exhaustive scan
__global__ void __launch_bounds__( RTX_BLOCK ) k_trace_rays( const SceneDev S, uint32_t n, const float* ori, const float* dir, float tmin, int brute, int64_t* id, float* t ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	const uint32_t r = blockIdx.x*RTX_BLOCK+threadIdx.x ;
	if ( r>=n )
		return ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ;
	const f3 o = mk3( ori[3*size_t( r )], ori[3*size_t( r )+1], ori[3*size_t( r )+2] ) ;
	const f3 d = mk3( dir[3*size_t( r )], dir[3*size_t( r )+1], dir[3*size_t( r )+2] ) ;
	HitRec hit ;
	if ( brute ) closest_brute( S, o, d, tmin, hit ) ;
	else         closest( S, o, d, tmin, st, hit ) ;
	id[r] = hit.thing<0 ? int64_t( -1 ) : ( ( int64_t( hit.thing )<<32 )|int64_t( uint32_t( hit.prim+1 ) ) ) ;
	t[r]  = hit.thing<0 ? -1.f : hit.t ;
}

// mean of the fixed-point sums -> clamp(0,1) (optx/camera_i.cu:105) -> rawRGB; rpp
__global__ void __launch_bounds__( 256 ) k_resolve( const uint64_t* accum, uint32_t npix, uint64_t total_spp, float* raw, uint32_t* rpp ) {
	const uint32_t p = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( p>=npix )
		return ;
	const ulonglong2 lo = reinterpret_cast<const ulonglong2*>( accum )[2*size_t( p )] ;
	const ulonglong2 hi = reinterpret_cast<const ulonglong2*>( accum )[2*size_t( p )+1] ;
	const double n = double( total_spp ) ;
	const uint64_t s[3] = { lo.x, lo.y, hi.x } ;
	for ( int c = 0 ; c<3 ; c++ ) {
		const float v = float( double( s[c] )*( 1./4294967296. )/n ) ;
		raw[3*size_t( p )+c] = 0.f>v ? 0.f : v>1.f ? 1.f : v ;
	}
	rpp[p] = uint32_t( hi.y ) ;
}

// optx/postproc.cu:18-34 (none) and :2-16, 36-47 (sRGB): float3 -> uchar4, truncating
__global__ void __launch_bounds__( 256 ) k_postproc( const float* raw, uchar4* dst, uint32_t npix, int srgb ) {
	const uint32_t p = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( p>=npix )
		return ;
	float c[3] = { raw[3*size_t( p )], raw[3*size_t( p )+1], raw[3*size_t( p )+2] } ;
	if ( srgb )
		for ( int k = 0 ; k<3 ; k++ )
			c[k] = c[k]<.0031308f ? 12.92f*c[k] : 1.055f*powf( c[k], 1.f/2.4f )-.055f ;
	dst[p] = make_uchar4( static_cast<unsigned char>( c[0]*255 ), static_cast<unsigned char>( c[1]*255 ), static_cast<unsigned char>( c[2]*255 ), 255u ) ;
}

} // namespace rtx
