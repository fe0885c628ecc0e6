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
	uint32_t* fault ;   // SceneDev::fault
	int32_t  ovf[RTX_OVF_STACK] ;
	__device__ __forceinline__ void reset() { sp = 0 ; }
	__device__ __forceinline__ bool empty() const { return sp == 0 ; }
	__device__ __forceinline__ void push( int32_t v ) {
		if ( sp<RTX_SM_STACK ) base[sp*RTX_BLOCK] = v ;
		else if ( sp<RTX_SM_STACK+RTX_OVF_STACK ) ovf[sp-RTX_SM_STACK] = v ;
		else stack_fault( fault ) ;
		sp++ ;
	}
	__device__ __forceinline__ int32_t pop() {
		sp-- ;
		if ( sp<RTX_SM_STACK ) return base[sp*RTX_BLOCK] ;
		return sp<RTX_SM_STACK+RTX_OVF_STACK ? ovf[sp-RTX_SM_STACK] : RTX_STK_RETURN ;
	}
} ;

#define RTX_TAPER_MAX 39
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
	// work units of k_render (unit_plan): `chunks_full` sample chunks of RTX_UNIT_SPP, then
	// `chunks_taper` shrinking ones starting at taper_s0[k] (relative to the end of the full ones)
	uint32_t  chunks_full, chunks_taper ;
	uint16_t  taper_s0[RTX_TAPER_MAX+1] ;
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

// fire-and-forget 64-bit add to GLOBAL memory (RED.E.ADD.64): atomicAdd() on a pointer the
// compiler only knows as generic emits an ATOM plus a shared-memory CAS spin path
__device__ __forceinline__ void red_add_u64( unsigned long long* addr, unsigned long long v ) {
	asm volatile( "red.global.add.u64 [%0], %1;" :: "l"( __cvta_generic_to_global( addr ) ), "l"( v ) : "memory" ) ;
}

// The path tracer (see rtx_pool.cuh): one warp per CTA, persistent.  Work is handed out in
// small units -- an 8x4 pixel tile x RTX_UNIT_SPP samples = up to 2048 paths, fewer towards the end -- through a
// global counter; a warp streams from one unit straight into the next (a lane whose path
// ends takes the next path of the warp's current unit, whichever tile that is), so no lane
// idles at tile ends and the kernel's tail is one unit, not one tile of 16 000 paths.  A
// finished path adds its fixed-point colour and its segment count to the pixel with global
// integer atomics: order free, bit-reproducible, and no per-tile buffers.
#define RTX_POOL_R ( 32*RTX_K )
#ifndef RTX_STICKY
#define RTX_STICKY 6           // keep doing node steps while at least this many lanes have one (launch bound 18: 4 / 5 / 6 / 8 -> 586.7 / 582.9 / 580.8 / 581.3 ms per frame; launch bound 17: 4 / 5 / 6 / 7 / 8 -> 578.4 / 573.7 / 571.7 / 571.4 / 571.5)
#endif
#ifndef RTX_MIN_CTAS
#define RTX_MIN_CTAS 17         // launch bound: 17-20 all compile to 96 registers = 5 warps per scheduler, 20 per SM (24 at 80
                                // registers measure 9 % slower); which of them schedules the code best changes with the code --
                                // 17 / 18 / 19 / 20: 574.2 / 582.3 / 582.2 / 586.4 ms per frame with the final step functions
#endif
#ifndef RTX_NODE_BIAS
#define RTX_NODE_BIAS 0         // votes added to the node kind
#endif
#ifndef RTX_LONE
#define RTX_LONE 0             // (experiment) lanes left in a warp below which the tail of a launch runs without votes.  Measured with 2 / 4:
                              // the 500-spp frame 594.0 / 594.3 against 581.8 ms (the second copy of the step functions costs the main loop 2 %),
                              // the 63-spp frame 79.4 / 79.3 against 77.4, the 1-spp frame 5.4 / 5.2 against 5.1 ms: the votes are not what the tail waits for
#endif
#ifndef RTX_UNIT_SPP
#define RTX_UNIT_SPP 64u        // samples per pixel of a full work unit (16-128 measured alike; swept with the taper in place)
#endif
#ifndef RTX_SM_SUPER
#define RTX_SM_SUPER 1          // the warps of an SM draw their units from a shared block of adjacent tiles
#endif
#ifndef RTX_SUPER_BWLOG
#define RTX_SUPER_BWLOG 1       // an SM's shared block: 2^BWLOG x 2^BHLOG tiles x 32/(tiles) consecutive sample chunks
#endif
#ifndef RTX_SUPER_BHLOG
#define RTX_SUPER_BHLOG 1
#endif
#define RTX_SUPER_CG ( 32u>>( RTX_SUPER_BWLOG+RTX_SUPER_BHLOG ) )
#ifndef RTX_TILE_WLOG
#define RTX_TILE_WLOG 3         // a unit's pixels: a tile of 2^WLOG x 2^HLOG (at most 32) ...
#endif
#ifndef RTX_TILE_HLOG
#define RTX_TILE_HLOG 2
#endif
#define RTX_TILE_W ( 1u<<RTX_TILE_WLOG )
#define RTX_TILE_H ( 1u<<RTX_TILE_HLOG )
#define RTX_TILE_PLOG ( RTX_TILE_WLOG+RTX_TILE_HLOG )
#define RTX_TILE_P ( 1u<<RTX_TILE_PLOG )
#ifndef RTX_UNIT_TAIL
#define RTX_UNIT_TAIL 2u        // ... of the smallest one
#endif
// Work units taper off.  A frame's samples are cut into chunks of RTX_UNIT_SPP, except for the
// last RTX_UNIT_SPP..2*RTX_UNIT_SPP-1 of them (all of them in a short frame), which go out in
// chunks that shrink by a quarter of what is left each time, down to RTX_UNIT_TAIL.  Units are
// handed out chunk by chunk, so the kernel ends on small ones.  Why: per-warp finish times
// (RTX_DEBUG_TIMES) showed warps working up to 12 ms on their last 512-path unit after the
// counter had run dry -- a 15 ms tail on every launch, whatever its length; units of 64 paths
// throughout, on the other hand, trace 17 % slower (a warp that stays on one tile for hundreds
// of paths keeps its lanes in the same part of the hierarchy).
inline void unit_plan( FrameArgs& a ) {
	a.chunks_full = a.spp>RTX_UNIT_SPP ? ( a.spp-RTX_UNIT_SPP )/RTX_UNIT_SPP : 0u ;
	uint32_t rem = a.spp-a.chunks_full*RTX_UNIT_SPP, s = 0, n = 0 ;
	while ( rem>0 && n<RTX_TAPER_MAX ) {
		uint32_t len = ( rem+3u )/4u ;
		if ( len<RTX_UNIT_TAIL ) len = RTX_UNIT_TAIL ;
		if ( len>rem || n == RTX_TAPER_MAX-1 ) len = rem ;
		a.taper_s0[n++] = uint16_t( s ) ;
		s += len ; rem -= len ;
	}
	a.taper_s0[n] = uint16_t( s ) ;
	a.chunks_taper = n ;
}
// SHADE step of one lane + what it adds to the frame: the guide sums of a first diffuse / reflecting hit, and
// the radiance and segment count of a path that has ended
template <bool GUIDES, class P>
__device__ __forceinline__ int shade_and_accumulate( P& p, int slot, const SceneDev& S, unsigned long long* accum, unsigned long long* guide ) {
	f3 c, gn, ga ; bool g ;
	uint32_t segments ;
	const int nk = step_shade( p, slot, S, c, g, gn, ga, segments ) ;
	const size_t pix = size_t( uint32_t( p.i( F_PIX, slot ) ) ) ;
	if ( GUIDES && g ) {
		const float v[6] = { gn.x, gn.y, gn.z, ga.x, ga.y, ga.z } ;
		for ( int q = 0 ; q<6 ; q++ )
			red_add_u64( guide+6*pix+q, ( unsigned long long )( long long )( v[q]*1073741824.f ) ) ;
	}
	if ( nk == K_REGEN ) {
		// (0 contributions -- absorbed paths -- need no atomic)
		const unsigned long long r = tofix( c.x ), gg = tofix( c.y ), bb = tofix( c.z ) ;
#if ! defined( RTX_EXPERIMENT_NO_ACCUM )
		if ( r )  red_add_u64( accum+4*pix, r ) ;
		if ( gg ) red_add_u64( accum+4*pix+1, gg ) ;
		if ( bb ) red_add_u64( accum+4*pix+2, bb ) ;
		red_add_u64( accum+4*pix+3, ( unsigned long long ) segments ) ;
#else
		if ( r+gg+bb == 1 ) atomicAdd( accum+4*pix+3, ( unsigned long long ) segments ) ;
#endif
	}
	return nk ;
}
#if defined( RTX_MAXNREG )
#define RTX_RENDER_BOUNDS __maxnreg__( RTX_MAXNREG )                 // (tuning: an exact register budget instead of a CTA count)
#else
#define RTX_RENDER_BOUNDS __launch_bounds__( 32, RTX_MIN_CTAS )
#endif
template <bool GUIDES>
__global__ void RTX_RENDER_BOUNDS k_render( const __grid_constant__ FrameArgs a, uint32_t* unit_counter, int32_t* ovf_all ) {
	const uint32_t lane = threadIdx.x ;
	const uint32_t lt = ( 1u<<lane )-1u ;
#if defined( RTX_REGPOOL )
	__shared__ __align__( 8 ) uint32_t stack_words[2*RTX_POOL_STACK*32] ;
	RegPool p ;
	p.stk = uint32_t( __cvta_generic_to_shared( stack_words+2*lane ) ) ;   // 8 bytes per lane and entry: a row of the stack is 256 bytes, one 64-bit access per lane
	__shared__ uint32_t cold_words[( RTX_COLD_WORDS>0 ? RTX_COLD_WORDS : 1 )*32] ;
	p.cold = uint32_t( __cvta_generic_to_shared( cold_words+lane ) ) ;
	p.ovf_all = ovf_all ;
#if ! RTX_OVF_LAZY
	p.ovf_ = ovf_all+( size_t( blockIdx.x )*RTX_POOL_R+lane )*RTX_POOL_OVF*2 ;
#endif
	p.fault = a.S.fault ;
#else
	__shared__ uint32_t words[F_WORDS*RTX_POOL_R] ;
	DevPool p ;
	p.w = words ;
	p.ovf = ovf_all+size_t( blockIdx.x )*RTX_POOL_R*RTX_POOL_OVF*2 ;
	p.fault = a.S.fault ;
#endif
	const uint32_t tiles_x = ( a.w+RTX_TILE_W-1u )>>RTX_TILE_WLOG, tiles_y = ( a.h+RTX_TILE_H-1u )>>RTX_TILE_HLOG, n_tiles = tiles_x*tiles_y ;
	const uint32_t n_chunks = a.chunks_full+a.chunks_taper ;
	const uint32_t n_units = n_tiles*n_chunks ;   // unit u: chunk u/n_tiles of tile u%n_tiles
#if RTX_SM_SUPER
	const uint32_t supers_x = ( tiles_x+( 1u<<RTX_SUPER_BWLOG )-1u )>>RTX_SUPER_BWLOG, n_supers = supers_x*( ( tiles_y+( 1u<<RTX_SUPER_BHLOG )-1u )>>RTX_SUPER_BHLOG ) ;
	uint32_t smid ;
	asm( "mov.u32 %0, %smid;" : "=r"( smid ) ) ;
	smid &= 255u ;   // (256 per-SM words are allocated)
#endif
	unsigned long long* accum = reinterpret_cast<unsigned long long*>( a.accum ) ;
	unsigned long long* guide = reinterpret_cast<unsigned long long*>( a.guide_acc ) ;

	// the warp's current unit (uniform over the warp)
	uint32_t unit_x0 = 0, unit_y0 = 0, unit_s0 = 0, unit_pos = 0, unit_left = 0 ;
	bool exhausted = false ;

	int kinds[RTX_K] ;
#pragma unroll
	for ( int j = 0 ; j<RTX_K ; j++ ) kinds[j] = K_REGEN ;
#if defined( RTX_DEBUG_TIMES )
	// development instrument: when this warp started, ran out of units, ended (ns) -> hit_id[3*warp..]
	unsigned long long dbg_t0, dbg_tx = 0 ;
	asm volatile( "mov.u64 %0, %globaltimer;" : "=l"( dbg_t0 ) ) ;
#endif

	while ( true ) {
#if RTX_LONE && RTX_K == 1 && ! defined( RTX_DEVICE_COUNTERS )
		// the end of a launch: no paths left to hand out and at most RTX_LONE lanes of the warp still on a path (the few that
		// bounce dozens of times inside a glass sphere).  There is nothing to schedule any more: each lane runs its path to
		// the end on its own, one step after the other, without the vote and the dispatch of every iteration
		if ( exhausted && __popc( __ballot_sync( 0xffffffffu, kinds[0] != K_DONE ) )<=RTX_LONE ) {
			int k = kinds[0] ;
			while ( k != K_DONE ) {
				if ( k == K_NODE ) k = step_node( p, int( lane ), a.S ) ;
				else if ( k == K_LEAF ) k = step_leaf( p, int( lane ), a.S ) ;
				else if ( k == K_THING ) k = step_thing( p, int( lane ), a.S ) ;
				else if ( k == K_SHADE ) k = shade_and_accumulate<GUIDES>( p, int( lane ), a.S, accum, guide ) ;
				else k = K_DONE ;   // (a new path: there is none)
			}
			break ;
		}
#endif
		// vote: which step kind can most lanes take?
#if RTX_K == 1
		// one ray per lane: lanes of equal kind find each other (match), the largest group
		// wins (warp-wide max of size<<3|kind)
		const uint32_t peers = __match_any_sync( 0xffffffffu, kinds[0] ) ;
		const int kind = int( __reduce_max_sync( 0xffffffffu, kinds[0] == K_DONE ? 0u : ( uint32_t( __popc( peers )+( kinds[0] == K_NODE ? RTX_NODE_BIAS : 0 ) )<<3 )|uint32_t( kinds[0] ) )&7u ) ;
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
#if RTX_K == 1 && ! defined( RTX_DEVICE_COUNTERS ) && ! defined( RTX_STICKY_GENERIC )
				// (one ray per lane: the loop tests the lane's kind, no slot index)
				do {
					if ( kinds[0] == K_NODE )
						kinds[0] = step_node( p, int( lane ), a.S ) ;
				} while ( __popc( __ballot_sync( 0xffffffffu, kinds[0] == K_NODE ) )>=RTX_STICKY ) ;
				j = -1 ;
				break ;
#endif
				int jn = j ;
				while ( true ) {
					RTX_COUNT_STEP( K_NODE, __ballot_sync( 0xffffffffu, jn>=0 ) ) ;
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
				RTX_COUNT_STEP( K_LEAF, __ballot_sync( 0xffffffffu, j>=0 ) ) ;
				if ( j>=0 ) nk = step_leaf( p, slot, a.S ) ;
				break ;
			case K_THING:
				RTX_COUNT_STEP( K_THING, __ballot_sync( 0xffffffffu, j>=0 ) ) ;
				if ( j>=0 ) nk = step_thing( p, slot, a.S ) ;
				break ;
			case K_SHADE:
				RTX_COUNT_STEP( K_SHADE, __ballot_sync( 0xffffffffu, j>=0 ) ) ;
				if ( j>=0 ) nk = shade_and_accumulate<GUIDES>( p, slot, a.S, accum, guide ) ;
				break ;
			default: {   // K_REGEN: hand the next paths of the warp's unit(s) to the lanes that ask
				uint32_t want = __ballot_sync( 0xffffffffu, j>=0 ) ;
				RTX_COUNT_STEP( K_REGEN, want ) ;
				while ( want ) {
					if ( unit_left == 0 && ! exhausted ) {
						uint32_t u = 0 ;
#if RTX_SM_SUPER
						// the warps of an SM share a block of 2x2 adjacent tiles x 8 sample chunks: what one
						// warp pulls into L1 the others use.  Per-SM word: (block+1)<<32 | next tile of the block;
						// the warp that draws the first index past the block fetches the SM's next block from the
						// global counter, the others retry until it is installed.
						if ( lane == 0 ) {
							unsigned long long* state = reinterpret_cast<unsigned long long*>( unit_counter+2 )+smid ;
							while ( true ) {
								const unsigned long long old = atomicAdd( state, 1ull ) ;
								const uint32_t hi = uint32_t( old>>32 ), idx = uint32_t( old ) ;
								if ( hi == 0xffffffffu ) { u = 0xffffffffu ; break ; }
								if ( hi != 0u && idx<32u ) {
									const uint32_t g = hi-1u, cg = g/n_supers, st = g%n_supers ;
									const uint32_t ti = idx&( ( 1u<<( RTX_SUPER_BWLOG+RTX_SUPER_BHLOG ) )-1u ) ;
									const uint32_t chunk = cg*RTX_SUPER_CG+( idx>>( RTX_SUPER_BWLOG+RTX_SUPER_BHLOG ) ) ;
									const uint32_t tx = ( ( st%supers_x )<<RTX_SUPER_BWLOG )+( ti&( ( 1u<<RTX_SUPER_BWLOG )-1u ) ), ty = ( ( st/supers_x )<<RTX_SUPER_BHLOG )+( ti>>RTX_SUPER_BWLOG ) ;
									u = ( tx<tiles_x && ty<tiles_y && chunk<n_chunks ) ? chunk*n_tiles+ty*tiles_x+tx : 0xfffffffeu ;   // (..fe: beyond the image edge / the last chunk)
									break ;
								}
								if ( ( hi == 0u && idx == 0u ) || ( hi != 0u && idx == 32u ) ) {
									const uint32_t g = atomicAdd( unit_counter, 1u ) ;
									atomicExch( state, g<n_supers*( ( n_chunks+RTX_SUPER_CG-1u )/RTX_SUPER_CG ) ? ( ( unsigned long long )( g+1u )<<32 ) : 0xffffffff00000000ull ) ;
								} else
									__nanosleep( 100 ) ;
							}
						}
						u = __shfl_sync( 0xffffffffu, u, 0 ) ;
						if ( u == 0xfffffffeu )
							continue ;
#else
						if ( lane == 0 ) u = atomicAdd( unit_counter, 1u ) ;
						u = __shfl_sync( 0xffffffffu, u, 0 ) ;
#endif
						if ( u>=n_units ) {
							exhausted = true ;
#if defined( RTX_DEBUG_TIMES )
							asm volatile( "mov.u64 %0, %globaltimer;" : "=l"( dbg_tx ) ) ;
#endif
						}
						else {
							const uint32_t tile = u%n_tiles, chunk = u/n_tiles ;
							unit_x0 = ( tile%tiles_x )<<RTX_TILE_WLOG ; unit_y0 = ( tile/tiles_x )<<RTX_TILE_HLOG ;
							uint32_t len = RTX_UNIT_SPP ;
							unit_s0 = chunk*RTX_UNIT_SPP ;
							if ( chunk>=a.chunks_full ) {
								const uint32_t k = chunk-a.chunks_full ;
								unit_s0 = a.chunks_full*RTX_UNIT_SPP+a.taper_s0[k] ;
								len = uint32_t( a.taper_s0[k+1] )-uint32_t( a.taper_s0[k] ) ;
							}
							unit_pos = 0 ;
							unit_left = RTX_TILE_P*len ;
						}
					}
					if ( exhausted ) {
						if ( ( want>>lane )&1u ) nk = K_DONE ;
						break ;
					}
					const uint32_t take = min( uint32_t( __popc( want ) ), unit_left ) ;
					const uint32_t rank = __popc( want&lt ) ;
					const bool served = ( ( want>>lane )&1u ) && rank<take ;
					if ( served ) {
						// path idx of the unit -> (sample, pixel of the tile); pixels beyond the image
						// border are skipped (the lane asks again)
						const uint32_t idx = unit_pos+rank ;
						const uint32_t px = idx&( RTX_TILE_P-1u ), smp = unit_s0+( idx>>RTX_TILE_PLOG ) ;
						const uint32_t x = unit_x0+( px&( RTX_TILE_W-1u ) ), y = unit_y0+( px>>RTX_TILE_WLOG ) ;
						if ( x<a.w && y<a.h )
							nk = step_regen( p, slot, a.S, a.cam, x, y, a.w, a.h, a.w*y+x, a.seed, a.sample0+smp*a.sample_stride, a.depth ) ;
					}
					unit_pos += take ; unit_left -= take ;
					want &= ~__ballot_sync( 0xffffffffu, served ) ;
				}
			}
		}
#pragma unroll
		for ( int jj = 0 ; jj<RTX_K ; jj++ ) if ( jj == j ) kinds[jj] = nk ;
	}
#if defined( RTX_DEBUG_TIMES )
	if ( lane == 0 && a.hit_id ) {
		unsigned long long t1 ;
		asm volatile( "mov.u64 %0, %globaltimer;" : "=l"( t1 ) ) ;
		a.hit_id[3*blockIdx.x] = int64_t( dbg_t0 ) ; a.hit_id[3*blockIdx.x+1] = int64_t( dbg_tx ) ; a.hit_id[3*blockIdx.x+2] = int64_t( t1 ) ;
	}
#endif
}

// first hit of the primary ray of sample `sample0` of every pixel
__global__ void __launch_bounds__( RTX_BLOCK ) k_primary_hits( const FrameArgs a ) {
	__shared__ int32_t stack_mem[RTX_SM_STACK*RTX_BLOCK] ;
	uint32_t x, y ;
	const bool valid = tile_pixel( a.w, a.h, x, y ) ;
	const uint32_t pix = a.w*y+x ;
	DevStack st ;
	st.base = stack_mem+threadIdx.x ; st.fault = a.S.fault ;
	Pcg rng ;
	rng.seed( a.seed, pix, a.sample0 ) ;
	f3 ori, dir ;
	primary_ray( a.cam, x, y, a.w, a.h, rng, ori, dir, a.S.variant != RTX_SEM_RTOW ) ;
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
	st.base = stack_mem+threadIdx.x ; st.fault = a.S.fault ;
	Pcg rng ;
	rng.seed( a.seed, a.w*py+px, a.sample0 ) ;
	f3 ori, dir ;
	primary_ray( a.cam, px, py, a.w, a.h, rng, ori, dir, a.S.variant != RTX_SEM_RTOW ) ;
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
	st.base = stack_mem+threadIdx.x ; st.fault = S.fault ;
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

// Multi-GPU frame (rtx_init_multi): the devices' fixed-point accumulation buffers summed by ONE
// kernel on the root that reads the replicas' buffers through peer memory (NVLink / NVSwitch),
// writes the total back as the root's buffer and, when asked, resolves it in the same pass --
// the reduce of SURVEY.md 8(e) fused with the mean + clamp that has to follow it (optx/camera_i.cu:105).
// Integer sums: the result does not depend on the number of devices.
#define RTX_MAX_DEVICES 8
struct PeerBufs { const uint64_t* accum[RTX_MAX_DEVICES] ; const long long* guide[RTX_MAX_DEVICES] ; int n ; } ;
__global__ void __launch_bounds__( 256 ) k_reduce_resolve( const PeerBufs pb, uint64_t* accum0, uint32_t npix, uint64_t total_spp, int resolve, float* raw, uint32_t* rpp,
		long long* guide0, float* normals, float* albedos ) {
	const uint32_t p = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( p>=npix )
		return ;
	ulonglong2 lo = reinterpret_cast<const ulonglong2*>( pb.accum[0] )[2*size_t( p )] ;
	ulonglong2 hi = reinterpret_cast<const ulonglong2*>( pb.accum[0] )[2*size_t( p )+1] ;
	for ( int r = 1 ; r<pb.n ; r++ ) {
		const ulonglong2 l2 = reinterpret_cast<const ulonglong2*>( pb.accum[r] )[2*size_t( p )] ;
		const ulonglong2 h2 = reinterpret_cast<const ulonglong2*>( pb.accum[r] )[2*size_t( p )+1] ;
		lo.x += l2.x ; lo.y += l2.y ; hi.x += h2.x ; hi.y += h2.y ;
	}
	reinterpret_cast<ulonglong2*>( accum0 )[2*size_t( p )] = lo ;
	reinterpret_cast<ulonglong2*>( accum0 )[2*size_t( p )+1] = hi ;
	const double n = double( total_spp ) ;
	if ( resolve ) {
		const uint64_t s[3] = { lo.x, lo.y, hi.x } ;
		for ( int c = 0 ; c<3 ; c++ ) {
			const float v = float( double( s[c] )*( 1./4294967296. )/n ) ;
			raw[3*size_t( p )+c] = 0.f>v ? 0.f : v>1.f ? 1.f : v ;
		}
		rpp[p] = uint32_t( hi.y ) ;
	}
	if ( guide0 ) {
		long long g[6] ;
		for ( int c = 0 ; c<6 ; c++ ) g[c] = pb.guide[0][6*size_t( p )+c] ;
		for ( int r = 1 ; r<pb.n ; r++ )
			for ( int c = 0 ; c<6 ; c++ ) g[c] += pb.guide[r][6*size_t( p )+c] ;
		for ( int c = 0 ; c<6 ; c++ ) guide0[6*size_t( p )+c] = g[c] ;
		if ( resolve )
			for ( int c = 0 ; c<3 ; c++ ) {
				normals[3*size_t( p )+c] = float( double( g[c] )*( 1./1073741824. )/n ) ;
				albedos[3*size_t( p )+c] = float( double( g[3+c] )*( 1./1073741824. )/n ) ;
			}
	}
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

// x^(1/2.4) of the sRGB transfer as one stated sequence of IEEE double operations (the reference's
// powf, optx/postproc.cu:9, is rounded differently by every math library): cube root by 14 Newton
// steps from 1 (converged after 10 for x in [0.0031308, 1]), then x^(5/12) = sqrt(sqrt(c^5)),
// rounded to float once.  The oracle (orc_srgb8) evaluates the same sequence: 8-bit codes are bit-exact.
__device__ __forceinline__ float srgb_pow( float x ) {
	const double v = double( x ) ;
	double c = 1. ;
	for ( int i = 0 ; i<14 ; i++ )
		c = ( 2.*c+v/( c*c ) )*( 1./3. ) ;
	const double c2 = c*c ;
	return float( sqrt( sqrt( ( c2*c2 )*c ) ) ) ;
}

// optx/postproc.cu:18-34 (none) and :2-16, 36-47 (sRGB): float3 -> uchar4, truncating
__global__ void __launch_bounds__( 256 ) k_postproc( const float* raw, uchar4* dst, uint32_t npix, int srgb ) {
	const uint32_t p = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( p>=npix )
		return ;
	float c[3] = { raw[3*size_t( p )], raw[3*size_t( p )+1], raw[3*size_t( p )+2] } ;
	if ( srgb )
		for ( int k = 0 ; k<3 ; k++ )
			c[k] = c[k]<.0031308f ? 12.92f*c[k] : 1.055f*srgb_pow( c[k] )-.055f ;
	dst[p] = make_uchar4( static_cast<unsigned char>( c[0]*255 ), static_cast<unsigned char>( c[1]*255 ), static_cast<unsigned char>( c[2]*255 ), 255u ) ;
}

// Measurement instrument (bench.py): every CTA reads the whole buffer `repeats`
// times with 128-bit loads that bypass L1 (ld.global.cg) -- with a buffer that fits L2 this is
// the L2 read bandwidth the node fetches of k_render compete for.
__global__ void __launch_bounds__( 256 ) k_probe_read( const uint4* buf, size_t n_vec, uint32_t repeats, uint32_t* sink ) {
	uint32_t acc = 0 ;
	const size_t stride = size_t( gridDim.x )*blockDim.x ;
	for ( uint32_t r = 0 ; r<repeats ; r++ )
#pragma unroll 8
		for ( size_t i = size_t( blockIdx.x )*blockDim.x+threadIdx.x ; i<n_vec ; i += stride ) {
			uint4 v ;
			asm volatile( "ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"( v.x ), "=r"( v.y ), "=r"( v.z ), "=r"( v.w ) : "l"( buf+i ) ) ;
			acc ^= v.x^v.y^v.z^v.w ;
		}
	if ( acc == 0x9e3779b9u ) *sink = acc ;   // (keeps the loads alive)
}

} // namespace rtx
