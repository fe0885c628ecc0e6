// rtx_kernels.cuh -- the frame kernels: path tracing (replaces __raygen__camera,
// __miss__ambient and the three __closesthit__ programs, optx/camera_i.cu:24-141 and
// optx/optics_i.cu:23-288, plus OptiX's traversal), the resolve (mean + clamp of
// optx/camera_i.cu:105 applied to the fixed-point sums), the post-processing pair of
// optx/postproc.cu, and the parity instruments.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rtx_core.cuh"
#include "rtx_pool.cuh"

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
	uint32_t  guides ;   // fill the guide layers too
	long long* guide_acc ; // [6*w*h] fixed-point (2^-30) sums: normal xyz, albedo rgb
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

// The path tracer (see rtx_pool.cuh): one warp per CTA, persistent; a warp takes an 8x4 pixel
// tile, traces all its paths -- 32*spp of them, RTX_K*32 in flight -- and writes the tile's
// fixed-point sums.  Tiles are handed out through a global counter.
#define RTX_POOL_R ( 32*RTX_K )
#ifndef RTX_STICKY
#define RTX_STICKY 14          // keep doing node steps while at least this many lanes have one
#endif
#ifndef RTX_MIN_CTAS
#define RTX_MIN_CTAS 20         // resident render warps per SM the register budget is set for (96 registers)
#endif
template <bool GUIDES>
__global__ void __launch_bounds__( 32, RTX_MIN_CTAS ) k_render( const FrameArgs a, uint32_t* tile_counter, int32_t* ovf_all ) {
	__shared__ unsigned long long acc[32*3] ;
	__shared__ unsigned long long gacc[GUIDES ? 32*6 : 1] ;
	__shared__ uint32_t segs[32] ;
	const uint32_t lane = threadIdx.x ;
	const uint32_t lt = ( 1u<<lane )-1u ;
#if defined( RTX_REGPOOL )
	__shared__ uint32_t stack_words[2*RTX_POOL_STACK*32] ;
	RegPool p ;
	p.stk = stack_words+lane ;
	p.ovf = ovf_all+( size_t( blockIdx.x )*RTX_POOL_R+lane )*RTX_POOL_OVF*2 ;
#else
	__shared__ uint32_t words[F_WORDS*RTX_POOL_R] ;
	DevPool p ;
	p.w = words ;
	p.ovf = ovf_all+size_t( blockIdx.x )*RTX_POOL_R*RTX_POOL_OVF*2 ;
#endif
	const uint32_t tiles_x = ( a.w+7u )>>3, tiles_y = ( a.h+3u )>>2, n_tiles = tiles_x*tiles_y ;

	while ( true ) {
		uint32_t tile = 0 ;
		if ( lane == 0 ) tile = atomicAdd( tile_counter, 1u ) ;
		tile = __shfl_sync( 0xffffffffu, tile, 0 ) ;
		if ( tile>=n_tiles )
			break ;
		const uint32_t x0 = ( tile%tiles_x )*8u, y0 = ( tile/tiles_x )*4u ;
		acc[lane] = 0 ; acc[lane+32] = 0 ; acc[lane+64] = 0 ; segs[lane] = 0 ;
		if ( GUIDES )
			for ( int g = 0 ; g<6 ; g++ ) gacc[lane+32*g] = 0 ;
		const uint32_t vmask = __ballot_sync( 0xffffffffu, x0+( lane&7u )<a.w && y0+( lane>>3 )<a.h ) ;
		const uint32_t n_valid = __popc( vmask ) ;
		const uint32_t total = n_valid*a.spp ;   // paths of this tile: index = sample*n_valid + (rank of pixel)
		uint32_t next = 0 ;
		int kinds[RTX_K] ;
#pragma unroll
		for ( int j = 0 ; j<RTX_K ; j++ ) kinds[j] = K_REGEN ;
		__syncwarp() ;

		while ( true ) {
			// vote: which step kind can most lanes take?
#if RTX_K == 1
			// one ray per lane: lanes of equal kind find each other (match), the largest group
			// wins (warp-wide max of size<<3|kind)
			const uint32_t peers = __match_any_sync( 0xffffffffu, kinds[0] ) ;
			const int kind = int( __reduce_max_sync( 0xffffffffu, kinds[0] == K_DONE ? 0u : ( uint32_t( __popc( peers ) )<<3 )|uint32_t( kinds[0] ) )&7u ) ;
#else
			uint32_t mine = 0 ;
#pragma unroll
			for ( int j = 0 ; j<RTX_K ; j++ ) mine |= 1u<<kinds[j] ;
			int kind = K_DONE ; int most = 0 ;
#pragma unroll
			for ( int k = 1 ; k<K_KINDS ; k++ ) {
				const int n = __popc( __ballot_sync( 0xffffffffu, ( mine>>k )&1u ) ) ;
				if ( n>most ) { most = n ; kind = k ; }
			}
#endif
			if ( kind == K_DONE )
				break ;
			int j = -1 ;
#pragma unroll
			for ( int jj = RTX_K-1 ; jj>=0 ; jj-- ) if ( kinds[jj] == kind ) j = jj ;
			const int slot = j*32+int( lane ) ;
			int nk = kind ;
			switch ( kind ) {
				case K_NODE: {
					// node steps dominate: stay with them (one ballot per step instead of a full
					// vote) while enough lanes still have one
					int jn = j ;
					while ( true ) {
						if ( jn>=0 ) {
							const int kn = step_node( p, jn*32+int( lane ), a.S ) ;
#pragma unroll
							for ( int jj = 0 ; jj<RTX_K ; jj++ ) if ( jj == jn ) kinds[jj] = kn ;
						}
						jn = -1 ;
#pragma unroll
						for ( int jj = RTX_K-1 ; jj>=0 ; jj-- ) if ( kinds[jj] == K_NODE ) jn = jj ;
						if ( __popc( __ballot_sync( 0xffffffffu, jn>=0 ) )<RTX_STICKY )
							break ;
					}
					j = -1 ;   // kinds[] already updated
					break ;
				}
				case K_LEAF:
					if ( j>=0 ) nk = step_leaf( p, slot, a.S ) ;
					break ;
				case K_THING:
					if ( j>=0 ) nk = step_thing( p, slot, a.S ) ;
					break ;
				case K_SHADE:
					if ( j>=0 ) {
						f3 c, gn, ga ; bool g ;
						nk = step_shade( p, slot, a.S, c, g, gn, ga ) ;
						const uint32_t px = uint32_t( p.i( F_PIX, slot ) ) ;
						atomicAdd( segs+px, 1u ) ;
						if ( GUIDES && g ) {
							const float v[6] = { gn.x, gn.y, gn.z, ga.x, ga.y, ga.z } ;
							for ( int q = 0 ; q<6 ; q++ )
								atomicAdd( gacc+32*q+px, ( unsigned long long )( long long )( v[q]*1073741824.f ) ) ;
						}
						if ( nk == K_REGEN ) {
							atomicAdd( acc+px, ( unsigned long long ) tofix( c.x ) ) ;
							atomicAdd( acc+32+px, ( unsigned long long ) tofix( c.y ) ) ;
							atomicAdd( acc+64+px, ( unsigned long long ) tofix( c.z ) ) ;
						}
					}
					break ;
				default: {   // K_REGEN
					const uint32_t want = __ballot_sync( 0xffffffffu, j>=0 ) ;
					const uint32_t idx = next+__popc( want&lt ) ;
					next += __popc( want ) ;
					if ( j>=0 ) {
						if ( idx<total ) {
							// path idx -> (sample, pixel of the tile); tiles cut by the image border take the slow way
							uint32_t px, smp ;
							if ( n_valid == 32u ) { px = idx&31u ; smp = idx>>5 ; }
							else { px = __fns( vmask, 0, int( idx%n_valid )+1 ) ; smp = idx/n_valid ; }
							nk = step_regen( p, slot, a.S, a.cam, x0+( px&7u ), y0+( px>>3 ), a.w, a.h, px, a.seed, a.sample0+smp*a.sample_stride, a.depth ) ;
						} else
							nk = K_DONE ;
					}
				}
			}
#pragma unroll
			for ( int jj = 0 ; jj<RTX_K ; jj++ ) if ( jj == j ) kinds[jj] = nk ;
		}
		__syncwarp() ;

		if ( ( vmask>>lane )&1u ) {
			const uint32_t pix = a.w*( y0+( lane>>3 ) )+x0+( lane&7u ) ;
			ulonglong2* out = reinterpret_cast<ulonglong2*>( a.accum+4*size_t( pix ) ) ;
			ulonglong2 lo = make_ulonglong2( acc[lane], acc[32+lane] ), hi = make_ulonglong2( acc[64+lane], ( unsigned long long ) segs[lane] ) ;
			if ( a.accumulate ) {
				const ulonglong2 plo = out[0], phi = out[1] ;
				lo.x += plo.x ; lo.y += plo.y ; hi.x += phi.x ; hi.y += phi.y ;
			}
			out[0] = lo ; out[1] = hi ;
			if ( GUIDES )
				for ( int g = 0 ; g<6 ; g++ ) {
					long long* o = a.guide_acc+6*size_t( pix )+g ;
					*o = ( a.accumulate ? *o : 0ll )+( long long ) gacc[32*g+lane] ;
				}
		}
		__syncwarp() ;
	}
}

// first hit of the primary ray of sample `sample0` of every pixel
__global__ void __launch_bounds__( RTX_BLOCK ) k_primary_hits( const FrameArgs a ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	uint32_t x, y ;
	const bool valid = tile_pixel( a.w, a.h, x, y ) ;
	const uint32_t pix = a.w*y+x ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ;
	Pcg rng ;
	rng.seed( a.seed, pix, a.sample0 ) ;
	f3 ori, dir ;
	primary_ray( a.cam, x, y, a.w, a.h, rng, ori, dir ) ;
	HitRec hit ;
	closest( a.S, ori, dir, 1e-3f, st, hit, valid ) ;
	if ( ! valid )
		return ;
	a.hit_id[pix] = hit.thing<0 ? int64_t( -1 ) : ( ( int64_t( hit.thing )<<32 )|int64_t( uint32_t( hit.prim+1 ) ) ) ;
	a.hit_t[pix]  = hit.thing<0 ? -1.f : hit.t ;
}

// picker (optx/camera_i.cu:27-29, optx/optics_i.cu:25-29): one primary ray through
// (px,py), the thing id or UINT_MAX.  Launched as one warp; lane 0 carries the ray.
__global__ void k_pick( const FrameArgs a, uint32_t px, uint32_t py, uint32_t* pick_id ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ;
	Pcg rng ;
	rng.seed( a.seed, a.w*py+px, a.sample0 ) ;
	f3 ori, dir ;
	primary_ray( a.cam, px, py, a.w, a.h, rng, ori, dir ) ;
	HitRec hit ;
	closest( a.S, ori, dir, 1e-3f, st, hit, threadIdx.x == 0 ) ;
	if ( threadIdx.x == 0 )
		*pick_id = hit.thing<0 ? 0xffffffffu : uint32_t( hit.thing ) ;
}

// closest hits of caller-supplied rays: through the LBVH, or by exhaustive scan
__global__ void __launch_bounds__( RTX_BLOCK ) k_trace_rays( const SceneDev S, uint32_t n, const float* ori, const float* dir, float tmin, int brute, int64_t* id, float* t ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	const uint32_t r = blockIdx.x*RTX_BLOCK+threadIdx.x ;
	const bool valid = r<n ;
	const uint32_t q = valid ? r : 0u ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ;
	const f3 o = mk3( ori[3*size_t( q )], ori[3*size_t( q )+1], ori[3*size_t( q )+2] ) ;
	const f3 d = mk3( dir[3*size_t( q )], dir[3*size_t( q )+1], dir[3*size_t( q )+2] ) ;
	HitRec hit ;
	if ( brute ) { if ( valid ) closest_brute( S, o, d, tmin, hit ) ; }
	else         closest( S, o, d, tmin, st, hit, valid ) ;
	if ( ! valid )
		return ;
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

// guide layers: mean of the fixed-point sums (optx/camera_i.cu:109-113)
__global__ void __launch_bounds__( 256 ) k_resolve_guides( const long long* gacc, uint32_t npix, uint64_t total_spp, float* normals, float* albedos ) {
	const uint32_t p = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( p>=npix )
		return ;
	const double n = double( total_spp ) ;
	for ( int c = 0 ; c<3 ; c++ ) {
		normals[3*size_t( p )+c] = float( double( gacc[6*size_t( p )+c] )*( 1./1073741824. )/n ) ;
		albedos[3*size_t( p )+c] = float( double( gacc[6*size_t( p )+3+c] )*( 1./1073741824. )/n ) ;
	}
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
