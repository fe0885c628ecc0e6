// rtx_qkernel.cuh -- k_render_q: the path tracer over a compacting ray pool (rtx_qpool.cuh).
// Replaces, like k_render, __raygen__camera / __miss__ambient / the three __closesthit__
// programs (optx/camera_i.cu:24-141, optx/optics_i.cu:23-288) and OptiX's traversal.
//
// One warp per CTA, persistent.  The warp owns RTX_QR ray slots in shared memory and one queue
// of slot numbers per step kind (a ring of bytes: rays leave at the head, join at the tail, so a
// ray that was queued for a node step waits a few iterations -- long enough for the prefetch of
// its node to land in L1).  Every iteration: the kind that fills the most lanes (at most 32)
// wins, lane i takes the i-th slot of that queue, advances the ray one step, and queues it under
// its new kind.  Work units, the per-SM blocks of adjacent tiles, the tapered tail and the
// fixed-point accumulation with global reductions are those of k_render (rtx_kernels.cuh).
#pragma once

#include "rtx_kernels.cuh"
#include "rtx_qpool.cuh"

namespace rtx {

#define RTX_QRING 128u   // entries per queue ring (a power of two, at least RTX_QR)
static_assert( RTX_QR<=RTX_QRING && RTX_QR<=255, "queue rings hold slot numbers as bytes" ) ;
static_assert( ( RTX_QS&1 ) == 1 && RTX_QS>=7, "slot records: an odd number of quads (bank spread), at least one stack quad" ) ;
#define RTX_Q_SMEM_BYTES ( RTX_QR*RTX_QS*16u+5u*RTX_QRING )
#ifndef RTX_Q_MIN_CTAS
#define RTX_Q_MIN_CTAS 8
#endif

template <bool GUIDES>
__global__ void __launch_bounds__( 32, RTX_Q_MIN_CTAS ) k_render_q( const __grid_constant__ FrameArgs a, uint32_t* unit_counter, int32_t* ovf_all, q4* cold_all ) {
	extern __shared__ __align__( 16 ) unsigned char q_smem[] ;
	const uint32_t lane = threadIdx.x ;
	const uint32_t lt = ( 1u<<lane )-1u ;
	QDev p ;
	p.base = uint32_t( __cvta_generic_to_shared( q_smem ) ) ;
	p.cold = cold_all+size_t( blockIdx.x )*RTX_QR*4 ;
	p.ovf = ovf_all+size_t( blockIdx.x )*RTX_QR*RTX_QOVF*2 ;
	p.fault = a.S.fault ;
	p.arena = a.S.arena ;
	unsigned char* ring = q_smem+RTX_QR*RTX_QS*16u ;   // [5][RTX_QRING]: queue of kind k at ring+(k-1)*RTX_QRING

	const uint32_t tiles_x = ( a.w+RTX_TILE_W-1u )>>RTX_TILE_WLOG, tiles_y = ( a.h+RTX_TILE_H-1u )>>RTX_TILE_HLOG, n_tiles = tiles_x*tiles_y ;
	const uint32_t n_chunks = a.chunks_full+a.chunks_taper ;
	const uint32_t n_units = n_tiles*n_chunks ;   // unit u: chunk u/n_tiles of tile u%n_tiles
	const uint32_t supers_x = ( tiles_x+( 1u<<RTX_SUPER_BWLOG )-1u )>>RTX_SUPER_BWLOG, n_supers = supers_x*( ( tiles_y+( 1u<<RTX_SUPER_BHLOG )-1u )>>RTX_SUPER_BHLOG ) ;
	uint32_t smid ;
	asm( "mov.u32 %0, %smid;" : "=r"( smid ) ) ;
	smid &= 255u ;   // (256 per-SM words are allocated)
	unsigned long long* accum = reinterpret_cast<unsigned long long*>( a.accum ) ;
	unsigned long long* guide = reinterpret_cast<unsigned long long*>( a.guide_acc ) ;

	// the warp's current unit (uniform over the warp)
	uint32_t unit_x0 = 0, unit_y0 = 0, unit_s0 = 0, unit_pos = 0, unit_left = 0 ;
	bool exhausted = false ;

	// queues (uniform state): heads H and lengths N of the kinds 1..4 packed one byte each (ring
	// positions are taken modulo RTX_QRING = 128, lengths stay below 256), the path queue apart.
	// At the start every slot asks for a path.
	uint32_t H = 0, N = 0, h5 = 0, n5 = RTX_QR ;
	for ( uint32_t s = lane ; s<RTX_QR ; s += 32u ) ring[4u*RTX_QRING+s] = ( unsigned char ) s ;

	while ( true ) {
		__syncwarp() ;   // what the lanes wrote to slots and queues in the last iteration is visible to all
		// which kind fills the most lanes (at most 32)?  ties: the later kind -- drain towards the end of the path
		const uint32_t n1 = N&255u, n2 = ( N>>8 )&255u, n3 = ( N>>16 )&255u, n4 = N>>24 ;
		const uint32_t mo = max( max( n2, n3 ), max( n4, n5 ) ) ;
		uint32_t kind, take ;
		if ( n1>=32u ? mo<32u : n1>mo ) { kind = K_NODE ; take = min( n1, 32u ) ; }
		else {
			if ( mo == 0u )
				break ;   // (n1 is 0 too)
			uint32_t best = 0 ;
#define RTX_QVOTE( k, n ) { const uint32_t v = ( min( n, 32u )<<3 )|uint32_t( k ) ; best = max( best, v ) ; }
			RTX_QVOTE( K_LEAF, n2 ) RTX_QVOTE( K_THING, n3 ) RTX_QVOTE( K_SHADE, n4 ) RTX_QVOTE( K_REGEN, n5 )
#undef RTX_QVOTE
			kind = best&7u ; take = best>>3 ;
		}
		const bool active = lane<take ;
		int slot = 0 ;
		int nk = K_DONE ;
		bool may_regen = false ;
		RTX_COUNT_STEP( kind, take>=32u ? 0xffffffffu : ( 1u<<take )-1u ) ;
		// take the first `take` slots of queue k (k = 1..4)
#define RTX_QTAKE( k ) { \
			const uint32_t sh = 8u*( k-1u ), hd = ( H>>sh )&127u ; \
			if ( active ) slot = ring[( k-1u )*RTX_QRING+( ( hd+lane )&( RTX_QRING-1u ) )] ; \
			H = ( H&~( 255u<<sh ) )|( ( ( hd+take )&127u )<<sh ) ; \
			N -= take<<sh ; }
		switch ( kind ) {
			case K_NODE:
				RTX_QTAKE( 1u )
				if ( active ) nk = qstep_node( p, slot, a.S ) ;
				break ;
			case K_LEAF:
				RTX_QTAKE( 2u )
				if ( active ) nk = qstep_leaf( p, slot, a.S ) ;
				break ;
			case K_THING:
				RTX_QTAKE( 3u )
				if ( active ) nk = qstep_thing( p, slot, a.S ) ;
				break ;
			case K_SHADE:
				RTX_QTAKE( 4u )
				may_regen = true ;
				if ( active ) {
					f3 c, gn, ga ; bool g ;
					uint32_t segments, pixw ;
					nk = qstep_shade( p, slot, a.S, c, pixw, g, gn, ga, segments ) ;
					const size_t pix = size_t( pixw ) ;
					if ( GUIDES && g ) {
						const float v[6] = { gn.x, gn.y, gn.z, ga.x, ga.y, ga.z } ;
						for ( int q = 0 ; q<6 ; q++ )
							red_add_u64( guide+6*pix+q, ( unsigned long long )( long long )( v[q]*1073741824.f ) ) ;
					}
					if ( nk == K_REGEN ) {
						// (0 contributions -- absorbed paths -- need no atomic)
						const unsigned long long r = tofix( c.x ), gg = tofix( c.y ), bb = tofix( c.z ) ;
						if ( r )  red_add_u64( accum+4*pix, r ) ;
						if ( gg ) red_add_u64( accum+4*pix+1, gg ) ;
						if ( bb ) red_add_u64( accum+4*pix+2, bb ) ;
						red_add_u64( accum+4*pix+3, ( unsigned long long ) segments ) ;
					}
				}
				break ;
			default: {   // K_REGEN: hand the next paths of the warp's unit(s) to the slots that ask
				may_regen = true ;
				if ( active ) {
					slot = ring[4u*RTX_QRING+( ( h5+lane )&( RTX_QRING-1u ) )] ;
					nk = K_REGEN ;
				}
				h5 += take ; n5 -= take ;
				uint32_t want = __ballot_sync( 0xffffffffu, active ) ;
				while ( want ) {
					if ( unit_left == 0 && ! exhausted ) {
						uint32_t u = 0 ;
						// the warps of an SM share a block of 2x2 adjacent tiles x 8 sample chunks (see k_render)
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
						if ( u>=n_units )
							exhausted = true ;
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
						if ( ( want>>lane )&1u ) nk = K_DONE ;   // the slot retires
						break ;
					}
					const uint32_t tk = min( uint32_t( __popc( want ) ), unit_left ) ;
					const uint32_t rank = __popc( want&lt ) ;
					const bool served = ( ( want>>lane )&1u ) && rank<tk ;
					if ( served ) {
						// path idx of the unit -> (sample, pixel of the tile); a pixel beyond the image
						// border is skipped (the slot stays in the queue and asks again)
						const uint32_t idx = unit_pos+rank ;
						const uint32_t px = idx&( RTX_TILE_P-1u ), smp = unit_s0+( idx>>RTX_TILE_PLOG ) ;
						const uint32_t x = unit_x0+( px&( RTX_TILE_W-1u ) ), y = unit_y0+( px>>RTX_TILE_WLOG ) ;
						if ( x<a.w && y<a.h )
							nk = qstep_regen( p, slot, a.S, a.cam, x, y, a.w, a.h, a.w*y+x, a.seed, a.sample0+smp*a.sample_stride, a.depth ) ;
					}
					unit_pos += tk ; unit_left -= tk ;
					want &= ~__ballot_sync( 0xffffffffu, served ) ;
				}
			}
		}
#undef RTX_QTAKE
		// queue every advanced ray under its new kind (finished slots, K_DONE, drop out): the lanes
		// with equal new kinds find each other (match), each writes its slot behind the tail of that
		// queue at its rank, and one lane per group adds the group's size to the packed lengths
		{
			const uint32_t peers = __match_any_sync( 0xffffffffu, nk ) ;
			const uint32_t rank = uint32_t( __popc( peers&lt ) ) ;
			const bool in4 = nk>=K_NODE && nk<=K_SHADE ;
			const uint32_t sh = in4 ? 8u*uint32_t( nk-1 ) : 0u ;
			if ( in4 ) ring[uint32_t( nk-1 )*RTX_QRING+( ( ( ( H+N )>>sh )+rank )&( RTX_QRING-1u ) )] = ( unsigned char ) slot ;
			N += __reduce_add_sync( 0xffffffffu, ( in4 && rank == 0u ) ? ( uint32_t( __popc( peers ) )<<sh ) : 0u ) ;
			if ( may_regen ) {
				const uint32_t m5 = __ballot_sync( 0xffffffffu, nk == K_REGEN ) ;
				if ( m5 ) {
					if ( nk == K_REGEN ) ring[4u*RTX_QRING+( ( h5+n5+uint32_t( __popc( m5&lt ) ) )&( RTX_QRING-1u ) )] = ( unsigned char ) slot ;
					n5 += uint32_t( __popc( m5 ) ) ;
				}
			}
		}
	}
}

} // namespace rtx
