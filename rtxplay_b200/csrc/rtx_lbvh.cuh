// rtx_lbvh.cuh -- GPU LBVH builder (replaces optixAccelBuild, optx/scene.cxx:122-138 for
// a mesh and :258-270 / :281-293 for the instance level).  All hand-written, no CUB:
//   k_bounds_reduce   centroid bounds of the primitives (ordered-int atomics)
//   k_morton          63-bit Morton key (21 bits/axis) of each primitive centroid
//   k_radix_hist8 / k_radix_bases / k_radix_onesweep
//                     stable LSD radix sort, 8-bit digits, one kernel per digit: tile prefixes
//                     by decoupled look-back, ranks from __match_any_sync, coalesced bucket runs
//                     (k_radix_hist / k_radix_scan / k_radix_scatter: the count / scan /
//                     scatter passes it replaced, -DRTX_SORT_ONESWEEP=0)
//   k_karras          Karras 2012 radix-tree hierarchy over the sorted keys (ties broken
//                     by position)
//   k_refit           bottom-up AABB refit with per-node arrival counters
//   k_wide_level      collapse of the binary tree into 4-wide, 128-byte traversal nodes with
//                     multi-primitive leaves, one launch per level of the wide tree
// Builder inputs are primitive AABBs, so the same code builds the per-mesh trees (over
// triangles) and the top level (over thing bounds); refit alone serves Scene::update.
#pragma once

#include <cuda_runtime.h>
#if defined( __CUDACC__ )
#include <cooperative_groups.h>
#endif
#include <stdint.h>

#include "rtx_core.cuh"

namespace rtx {

// ---- helpers usable on both sides (the host harness reuses them) ---------------------
RTX_HD uint64_t spread21( uint32_t v ) {
	uint64_t x = v&0x1fffffu ;
	x = ( x|( x<<32 ) )&0x1f00000000ffffull ;
	x = ( x|( x<<16 ) )&0x1f0000ff0000ffull ;
	x = ( x|( x<<8 ) )&0x100f00f00f00f00full ;
	x = ( x|( x<<4 ) )&0x10c30c30c30c30c3ull ;
	x = ( x|( x<<2 ) )&0x1249249249249249ull ;
	return x ;
}
RTX_HD uint64_t morton63( float x, float y, float z ) {   // x,y,z in [0,1]
	const float s = 2097151.f ;
	const uint32_t ix = uint32_t( fminf( fmaxf( x*s, 0.f ), s ) ) ;
	const uint32_t iy = uint32_t( fminf( fmaxf( y*s, 0.f ), s ) ) ;
	const uint32_t iz = uint32_t( fminf( fmaxf( z*s, 0.f ), s ) ) ;
	return ( spread21( ix )<<2 )|( spread21( iy )<<1 )|spread21( iz ) ;
}
RTX_HD int clz64( uint64_t v ) {
#if defined( __CUDA_ARCH__ )
	return __clzll( ( long long ) v ) ;
#else
	return v ? __builtin_clzll( v ) : 64 ;
#endif
}
// common-prefix length of sorted keys i and j, position breaks ties (Karras 2012, sec. 4)
RTX_HD int delta( const uint64_t* keys, int n, int i, int j ) {
	if ( j<0 || j>=n ) return -1 ;
	const uint64_t a = keys[i], b = keys[j] ;
	if ( a == b ) return 64+clz64( uint64_t( uint32_t( i )^uint32_t( j ) ) )-32 ;
	return clz64( a^b ) ;
}
// inner node i of the radix tree: children and covered range
RTX_HD void karras_node( const uint64_t* keys, int n, int i, int& left, int& right, bool& left_leaf, bool& right_leaf, int& range_lo, int& range_hi ) {
	const int d = ( delta( keys, n, i, i+1 )-delta( keys, n, i, i-1 ) )>=0 ? 1 : -1 ;
	const int dmin = delta( keys, n, i, i-d ) ;
	int lmax = 2 ;
	while ( delta( keys, n, i, i+lmax*d )>dmin ) lmax *= 2 ;
	int l = 0 ;
	for ( int t = lmax/2 ; t>=1 ; t /= 2 )
		if ( delta( keys, n, i, i+( l+t )*d )>dmin ) l += t ;
	const int j = i+l*d ;
	const int dnode = delta( keys, n, i, j ) ;
	int s = 0 ;
	int t = l ;
	do {
		t = ( t+1 )/2 ;
		if ( delta( keys, n, i, i+( s+t )*d )>dnode ) s += t ;
	} while ( t>1 ) ;
	const int gamma = i+s*d+( d<0 ? -1 : 0 ) ;
	const int lo = i<j ? i : j, hi = i<j ? j : i ;
	left = gamma ; right = gamma+1 ;
	left_leaf = ( lo == gamma ) ; right_leaf = ( hi == gamma+1 ) ;
	range_lo = lo ; range_hi = hi ;
}
// pad a box so that rounding in the (float) slab test can never exclude a primitive hit
RTX_HD void pad_box( f3& lo, f3& hi ) {
	const float m = fmaxf( fmaxf( fmaxf( fabsf( lo.x ), fabsf( hi.x ) ), fmaxf( fabsf( lo.y ), fabsf( hi.y ) ) ), fmaxf( fabsf( lo.z ), fabsf( hi.z ) ) ) ;
	const float e = m*( 1.f/8192.f )+1e-30f ;
	lo = mk3( lo.x-e, lo.y-e, lo.z-e ) ; hi = mk3( hi.x+e, hi.y+e, hi.z+e ) ;
}

// Collapse step: the up to RTX_WIDTH children of the wide node that replaces binary inner
// node `bin`.  Starts from its two children and keeps opening the slot with the largest
// box area, as long as that slot is an inner node covering more than leaf_max primitives.
// Slots: >= 0 binary inner node, < 0 binary leaf slot ~s.  Binary boxes: [0,n-1) inner,
// [n-1,2n-1) leaves.  (Inner node i covers the sorted range child-range[i].)
// a 16-byte record of a builder array with one 128-bit load (the arrays are 256-byte aligned pool allocations; q4
// itself only promises 4-byte alignment, so plain member accesses compile to scalar loads)
RTX_HD q4 ldbox( const q4* p ) {
#if defined( __CUDA_ARCH__ )
	const float4 v = *reinterpret_cast<const float4*>( p ) ;
	q4 r ; r.x = v.x ; r.y = v.y ; r.z = v.z ; r.w = v.w ; return r ;
#else
	return *p ;
#endif
}
RTX_HD float box_area( const q4& lo, const q4& hi ) {
	const float dx = hi.x-lo.x, dy = hi.y-lo.y, dz = hi.z-lo.z ;
	return dx*dy+dy*dz+dz*dx ;
}
template <class I2>
RTX_HD int wide_gather( int bin, const I2* child, const I2* range, const q4* blo, const q4* bhi, int n, int leaf_max, int slots[RTX_WIDTH] ) {
	int ns = 2 ;
	slots[0] = child[bin].x ; slots[1] = child[bin].y ;
	while ( ns<RTX_WIDTH ) {
		int pick = -1 ; float amax = -1.f ;
		for ( int k = 0 ; k<ns ; k++ ) {
			const int s = slots[k] ;
			if ( s<0 || range[s].y-range[s].x+1<=leaf_max )
				continue ;
			const float a = box_area( ldbox( blo+s ), ldbox( bhi+s ) ) ;
			if ( a>amax ) { amax = a ; pick = k ; }
		}
		if ( pick<0 )
			break ;
		const int s = slots[pick] ;
		slots[pick] = child[s].x ;
		slots[ns++] = child[s].y ;
	}
	return ns ;
}
// leaf ref of a slot that ends the wide tree: a binary leaf, or an inner node small enough
template <class I2>
RTX_HD bool wide_leaf_ref( int s, const I2* range, int leaf_max, int& ref ) {
	if ( s<0 ) { ref = ~( ( ( ~s )<<3 )|0 ) ; return true ; }
	const int cnt = range[s].y-range[s].x+1 ;
	if ( cnt<=leaf_max ) { ref = ~( ( range[s].x<<3 )|( cnt-1 ) ) ; return true ; }
	return false ;
}

#if defined( __CUDACC__ )

// ---- ordered-int float atomics ---------------------------------------------------------
__device__ __forceinline__ int f2ord( float f ) { const int i = __float_as_int( f ) ; return i>=0 ? i : i^0x7fffffff ; }
__device__ __forceinline__ float ord2f( int i ) { return __int_as_float( i>=0 ? i : i^0x7fffffff ) ; }

// bounds[0..2] = min centroid, [3..5] = max centroid (ordered ints; init +inf / -inf)
__global__ void k_bounds_init( int* bounds ) {
	if ( threadIdx.x<3 ) bounds[threadIdx.x] = f2ord( INFINITY ) ;
	else if ( threadIdx.x<6 ) bounds[threadIdx.x] = f2ord( -INFINITY ) ;
}
__global__ void __launch_bounds__( 256 ) k_bounds_reduce( const q4* plo, const q4* phi, uint32_t n, int* bounds ) {
	float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY } ;
	for ( uint32_t i = blockIdx.x*blockDim.x+threadIdx.x ; i<n ; i += gridDim.x*blockDim.x ) {
		const q4 lo = plo[i], hi = phi[i] ;
		const float c[3] = { .5f*( lo.x+hi.x ), .5f*( lo.y+hi.y ), .5f*( lo.z+hi.z ) } ;
		for ( int a = 0 ; a<3 ; a++ ) { mn[a] = fminf( mn[a], c[a] ) ; mx[a] = fmaxf( mx[a], c[a] ) ; }
	}
	for ( int a = 0 ; a<3 ; a++ )
		for ( int o = 16 ; o>0 ; o >>= 1 ) {
			mn[a] = fminf( mn[a], __shfl_xor_sync( 0xffffffffu, mn[a], o ) ) ;
			mx[a] = fmaxf( mx[a], __shfl_xor_sync( 0xffffffffu, mx[a], o ) ) ;
		}
	if ( ( threadIdx.x&31 ) == 0 )
		for ( int a = 0 ; a<3 ; a++ ) { atomicMin( bounds+a, f2ord( mn[a] ) ) ; atomicMax( bounds+3+a, f2ord( mx[a] ) ) ; }
}
__global__ void __launch_bounds__( 256 ) k_morton( const q4* plo, const q4* phi, uint32_t n, const int* bounds, uint64_t* keys, uint32_t* vals ) {
	const uint32_t i = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( i>=n ) return ;
	const float bx = ord2f( bounds[0] ), by = ord2f( bounds[1] ), bz = ord2f( bounds[2] ) ;
	const float ex = ord2f( bounds[3] )-bx, ey = ord2f( bounds[4] )-by, ez = ord2f( bounds[5] )-bz ;
	const q4 lo = plo[i], hi = phi[i] ;
	const float cx = .5f*( lo.x+hi.x ), cy = .5f*( lo.y+hi.y ), cz = .5f*( lo.z+hi.z ) ;
	keys[i] = morton63( ex>0.f ? ( cx-bx )/ex : 0.f, ey>0.f ? ( cy-by )/ey : 0.f, ez>0.f ? ( cz-bz )/ez : 0.f ) ;
	vals[i] = i ;
}

// ---- radix sort ----------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits.  A CTA of RTX_RS_WARPS warps owns a tile of
// RTX_RS_TILE consecutive keys; warp w owns the w-th run of RTX_RS_TILE/RTX_RS_WARPS of them
// and keeps them in registers (all loads of a tile are in flight at once).  k_radix_hist
// counts digits per tile, k_radix_scan turns the digit-major count table into global
// offsets, k_radix_scatter re-counts per warp, offsets each warp behind the warps before it
// and ranks the keys of a 32-key round with __match_any_sync (stable: rounds in order, lanes
// in order).
#ifndef RTX_SORT_ONESWEEP
#define RTX_SORT_ONESWEEP 1   // 0: this three-kernel pass (count, scan, scatter); 1: the one-sweep pass further down
#endif
#define RTX_RS_TILE   2048   // keys per CTA
#define RTX_RS_WARPS  4
#define RTX_RS_ROUNDS ( RTX_RS_TILE/RTX_RS_WARPS/32 )   // 32-key rounds per warp

__global__ void __launch_bounds__( 32*RTX_RS_WARPS ) k_radix_hist( const uint64_t* keys, uint32_t n, int shift, uint32_t* counts, uint32_t nblocks ) {
	__shared__ uint32_t hist[256] ;
	for ( int d = threadIdx.x ; d<256 ; d += 32*RTX_RS_WARPS ) hist[d] = 0 ;
	__syncthreads() ;
	const uint32_t base = blockIdx.x*RTX_RS_TILE ;
	uint64_t key[RTX_RS_TILE/( 32*RTX_RS_WARPS )] ;
#pragma unroll
	for ( int r = 0 ; r<RTX_RS_TILE/( 32*RTX_RS_WARPS ) ; r++ ) {
		const uint32_t i = base+uint32_t( r )*32u*RTX_RS_WARPS+threadIdx.x ;
		key[r] = i<n ? keys[i] : 0ull ;
	}
#pragma unroll
	for ( int r = 0 ; r<RTX_RS_TILE/( 32*RTX_RS_WARPS ) ; r++ ) {
		const uint32_t i = base+uint32_t( r )*32u*RTX_RS_WARPS+threadIdx.x ;
		if ( i<n ) atomicAdd( hist+uint32_t( ( key[r]>>shift )&255u ), 1u ) ;
	}
	__syncthreads() ;
	for ( int d = threadIdx.x ; d<256 ; d += 32*RTX_RS_WARPS ) counts[size_t( d )*nblocks+blockIdx.x] = hist[d] ;
}
// exclusive scan of counts[256*nblocks] (digit-major) in one block
__global__ void __launch_bounds__( 1024 ) k_radix_scan( uint32_t* counts, uint32_t m ) {
	__shared__ uint32_t part[1024] ;
	const uint32_t per = ( m+1023u )/1024u ;
	const uint32_t a = threadIdx.x*per, b = min( a+per, m ) ;
	uint32_t s = 0 ;
	for ( uint32_t i = a ; i<b ; i++ ) s += counts[i] ;
	part[threadIdx.x] = s ;
	__syncthreads() ;
	for ( int o = 1 ; o<1024 ; o <<= 1 ) {
		const uint32_t v = threadIdx.x>=o ? part[threadIdx.x-o] : 0u ;
		__syncthreads() ;
		part[threadIdx.x] += v ;
		__syncthreads() ;
	}
	uint32_t run = threadIdx.x ? part[threadIdx.x-1] : 0u ;
	for ( uint32_t i = a ; i<b ; i++ ) { const uint32_t c = counts[i] ; counts[i] = run ; run += c ; }
}
// the same over many blocks, for the count tables of big builds (256 counters per 2048 keys: 12.5 M
// counters for 100 M keys): every block scans a chunk of RTX_SCAN_CHUNK counters in place and leaves
// its total in sums[]; k_radix_scan then scans the (few hundred) totals; k_scan_add adds them back
#define RTX_SCAN_CHUNK 16384u   // counters per block: 1024 threads x 16
__global__ void __launch_bounds__( 1024 ) k_scan_chunks( uint32_t* counts, uint32_t m, uint32_t* sums ) {
	__shared__ uint32_t part[1024] ;
	const uint32_t base = blockIdx.x*RTX_SCAN_CHUNK+threadIdx.x*16u ;
	uint32_t v[16], s = 0 ;
	// (16 consecutive counters per thread: four 128-bit loads when the chunk is full)
#pragma unroll
	for ( int k = 0 ; k<16 ; k++ ) { v[k] = base+k<m ? counts[base+k] : 0u ; s += v[k] ; }
	part[threadIdx.x] = s ;
	__syncthreads() ;
	for ( int o = 1 ; o<1024 ; o <<= 1 ) {
		const uint32_t w = threadIdx.x>=o ? part[threadIdx.x-o] : 0u ;
		__syncthreads() ;
		part[threadIdx.x] += w ;
		__syncthreads() ;
	}
	uint32_t run = threadIdx.x ? part[threadIdx.x-1] : 0u ;
#pragma unroll
	for ( int k = 0 ; k<16 ; k++ ) { if ( base+k<m ) counts[base+k] = run ; run += v[k] ; }
	if ( threadIdx.x == 1023 ) sums[blockIdx.x] = part[1023] ;
}
__global__ void __launch_bounds__( 1024 ) k_scan_add( uint32_t* counts, uint32_t m, const uint32_t* sums ) {
	const uint32_t add = sums[blockIdx.x] ;
	const uint32_t base = blockIdx.x*RTX_SCAN_CHUNK+threadIdx.x*16u ;
#pragma unroll
	for ( int k = 0 ; k<16 ; k++ ) if ( base+k<m ) counts[base+k] += add ;
}
__global__ void __launch_bounds__( 32*RTX_RS_WARPS ) k_radix_scatter( const uint64_t* keys, const uint32_t* vals, uint32_t n, int shift, const uint32_t* offsets, uint32_t nblocks, uint64_t* keys_out, uint32_t* vals_out ) {
	__shared__ uint32_t off[RTX_RS_WARPS][256] ;
	const uint32_t lane = threadIdx.x&31u, warp = threadIdx.x>>5 ;
	for ( int d = threadIdx.x ; d<256*RTX_RS_WARPS ; d += 32*RTX_RS_WARPS ) ( &off[0][0] )[d] = 0 ;
	__syncthreads() ;
	// this warp's run of the tile, in registers
	const uint32_t base = blockIdx.x*RTX_RS_TILE+warp*( RTX_RS_TILE/RTX_RS_WARPS ) ;
	uint64_t key[RTX_RS_ROUNDS] ;
	uint32_t val[RTX_RS_ROUNDS] ;
#pragma unroll
	for ( int r = 0 ; r<RTX_RS_ROUNDS ; r++ ) {
		const uint32_t i = base+uint32_t( r )*32u+lane ;
		key[r] = i<n ? keys[i] : 0ull ;
		val[r] = i<n ? vals[i] : 0u ;
	}
#pragma unroll
	for ( int r = 0 ; r<RTX_RS_ROUNDS ; r++ )
		if ( base+uint32_t( r )*32u+lane<n ) atomicAdd( &off[warp][uint32_t( ( key[r]>>shift )&255u )], 1u ) ;
	__syncthreads() ;
	// digit d: tile offset, then the counts of the warps in front
	for ( int d = threadIdx.x ; d<256 ; d += 32*RTX_RS_WARPS ) {
		uint32_t run = offsets[size_t( d )*nblocks+blockIdx.x] ;
#pragma unroll
		for ( int w = 0 ; w<RTX_RS_WARPS ; w++ ) { const uint32_t c = off[w][d] ; off[w][d] = run ; run += c ; }
	}
	__syncthreads() ;
#pragma unroll
	for ( int r = 0 ; r<RTX_RS_ROUNDS ; r++ ) {
		const bool act = base+uint32_t( r )*32u+lane<n ;
		const uint32_t mask = __ballot_sync( 0xffffffffu, act ) ;
		if ( act ) {
			const uint32_t d = uint32_t( ( key[r]>>shift )&255u ) ;
			const uint32_t peers = __match_any_sync( mask, d ) ;
			const uint32_t rank = __popc( peers&( ( 1u<<lane )-1u ) ) ;
			const uint32_t pos = off[warp][d]+rank ;
			__syncwarp( mask ) ;
			if ( rank == 0 ) off[warp][d] += __popc( peers ) ;
			__syncwarp( mask ) ;
			keys_out[pos] = key[r] ;
			vals_out[pos] = val[r] ;
		}
	}
}

// ---- radix sort, one sweep per digit ---------------------------------------------------
// The same stable LSD sort with every key read once per pass (the sort above reads it twice and
// writes element by element).  k_radix_hist8 counts all eight digits of every key in one read;
// k_radix_bases turns each digit's 256 counts into the global start of its bucket; one launch of
// k_radix_onesweep per digit then does the rest: a CTA draws the next tile of RTX_OS_TILE keys
// (ticket: a tile's predecessors are always resident or done), ranks its keys per warp with
// __match_any_sync (rounds in order, lanes in order: stable), learns where its tile starts in
// every bucket by looking back over the tiles before it (decoupled look-back: a status word per
// tile and digit holds the tile's own count first, the inclusive count once it is known; thread d
// handles digit d; the walk comes after the keys and values have been moved into bucket order in
// shared memory, which needs nothing from other tiles) and writes each bucket's run with
// consecutive threads on consecutive addresses.
#define RTX_OS_THREADS 256
#ifndef RTX_OS_ITEMS
#define RTX_OS_ITEMS   16
#endif
#define RTX_OS_TILE    ( RTX_OS_THREADS*RTX_OS_ITEMS )   // 4096 keys per tile
#define RTX_OS_WARPS   ( RTX_OS_THREADS/32 )
#define RTX_OS_AGG     0x40000000u    // status: the tile's own count ...
#define RTX_OS_INCL    0x80000000u    // ... the count of all tiles up to and including it
#define RTX_OS_COUNT   0x3fffffffu
#ifndef RTX_OS_BALLOT_RANK
#define RTX_OS_BALLOT_RANK 1          // groups of equal digits from eight ballots (1) or from __match_any_sync (0).  106 M keys, eight
                                     // passes: ballots 9.6 (2 CTAs per SM) / 10.2 ms (3), match 10.8 / 10.1 ms
#endif
#ifndef RTX_OS_MIN_CTAS
#define RTX_OS_MIN_CTAS 2             // resident CTAs per SM the register budget is cut for (3: 80 registers and spills)
#endif
#ifndef RTX_OS_WINDOW
#define RTX_OS_WINDOW  4              // status words per look-back round trip (1 / 4 / 8 / 16: 10.0 / 10.1 / 10.2 / 10.9 ms: the walk is not what a pass waits for)
#endif

__global__ void __launch_bounds__( 256 ) k_radix_hist8( const uint64_t* keys, uint32_t n, uint32_t* hist ) {
	__shared__ uint32_t h[8*256] ;
	for ( int i = threadIdx.x ; i<8*256 ; i += 256 ) h[i] = 0 ;
	__syncthreads() ;
	const uint32_t lane = threadIdx.x&31u ;
	// (whole warps stay in the loop so that the warp votes below are complete)
	for ( uint32_t i0 = ( blockIdx.x*256u+( threadIdx.x&~31u ) ) ; i0<n ; i0 += gridDim.x*256u ) {
		const uint32_t i = i0+lane ;
		const bool act = i<n ;
		const uint64_t key = act ? keys[i] : 0ull ;
		const uint32_t mask = __ballot_sync( 0xffffffffu, act ) ;
		if ( act ) {
#pragma unroll
			for ( int p = 0 ; p<8 ; p++ ) {
				const uint32_t d = uint32_t( key>>( 8*p ) )&255u ;
				// neighbouring primitives share their high digits: one add for the warp then
				int same ;
				__match_all_sync( mask, d, &same ) ;
				if ( same ) { if ( lane == uint32_t( __ffs( mask )-1 ) ) atomicAdd( h+256*p+d, uint32_t( __popc( mask ) ) ) ; }
				else atomicAdd( h+256*p+d, 1u ) ;
			}
		}
	}
	__syncthreads() ;
	for ( int i = threadIdx.x ; i<8*256 ; i += 256 ) if ( h[i] ) atomicAdd( hist+i, h[i] ) ;
}
// hist[8][256] -> exclusive scan of every row (block p: digit p), in place
__global__ void __launch_bounds__( 256 ) k_radix_bases( uint32_t* hist ) {
	__shared__ uint32_t part[256] ;
	uint32_t* row = hist+256*blockIdx.x ;
	const uint32_t c = row[threadIdx.x] ;
	part[threadIdx.x] = c ;
	__syncthreads() ;
	for ( int o = 1 ; o<256 ; o <<= 1 ) {
		const uint32_t v = threadIdx.x>=o ? part[threadIdx.x-o] : 0u ;
		__syncthreads() ;
		part[threadIdx.x] += v ;
		__syncthreads() ;
	}
	row[threadIdx.x] = part[threadIdx.x]-c ;
}
// shared memory of a CTA (dynamic: more than the 48 KB a kernel gets without asking)
#define RTX_OS_SMEM_BYTES ( RTX_OS_TILE*12u+RTX_OS_WARPS*1024u+2u*1024u+64u )
__global__ void __launch_bounds__( RTX_OS_THREADS, RTX_OS_MIN_CTAS ) k_radix_onesweep( const uint64_t* keys, const uint32_t* vals, uint32_t n, int shift, const uint32_t* bases, uint32_t* status, uint32_t* ticket, uint64_t* keys_out, uint32_t* vals_out ) {
	extern __shared__ __align__( 16 ) unsigned char os_smem[] ;
	uint64_t* stage = reinterpret_cast<uint64_t*>( os_smem ) ;                              // [RTX_OS_TILE] keys in bucket order
	uint32_t* sval = reinterpret_cast<uint32_t*>( os_smem+RTX_OS_TILE*8u ) ;                // [RTX_OS_TILE] their indices
	uint32_t ( *wh )[256] = reinterpret_cast<uint32_t ( * )[256]>( os_smem+RTX_OS_TILE*12u ) ;   // per warp and digit: count, then offset inside the tile's run of the digit
	uint32_t* tile_start = reinterpret_cast<uint32_t*>( os_smem+RTX_OS_TILE*12u+RTX_OS_WARPS*1024u ) ;   // first position of the digit's run in the tile's bucket order
	uint32_t* gbase = tile_start+256 ;                                                      // where that run goes in the output
	uint32_t* wsum = gbase+256 ;                                                            // [RTX_OS_WARPS]
	uint32_t* s_tile = wsum+RTX_OS_WARPS ;
	const uint32_t tid = threadIdx.x, lane = tid&31u, warp = tid>>5 ;
	if ( tid == 0 ) *s_tile = atomicAdd( ticket, 1u ) ;
	for ( int d = tid ; d<256*RTX_OS_WARPS ; d += RTX_OS_THREADS ) ( &wh[0][0] )[d] = 0 ;
	__syncthreads() ;
	const uint32_t tile = *s_tile ;
	const uint32_t tile0 = tile*RTX_OS_TILE ;
	const uint32_t base = tile0+warp*( RTX_OS_TILE/RTX_OS_WARPS ) ;
	uint64_t key[RTX_OS_ITEMS] ;
	uint32_t val[RTX_OS_ITEMS], pos[RTX_OS_ITEMS] ;
#pragma unroll
	for ( int r = 0 ; r<RTX_OS_ITEMS ; r++ ) {
		const uint32_t i = base+uint32_t( r )*32u+lane ;
		key[r] = i<n ? keys[i] : 0ull ;
		val[r] = i<n ? vals[i] : 0u ;
	}
	// ranks inside the warp's run of the tile: the lanes of a round that share a digit find each other
	// (all rounds first: they are independent), the first of them bumps the warp's counter of
	// the digit and hands the old value to the others (the atomics of successive rounds queue up in
	// order -- no round waits for the one before it)
#pragma unroll
	for ( int r = 0 ; r<RTX_OS_ITEMS ; r++ ) {
		const bool act = base+uint32_t( r )*32u+lane<n ;
		const uint32_t mask = __ballot_sync( 0xffffffffu, act ) ;
#if RTX_OS_BALLOT_RANK
		// the lanes with the same digit, from eight votes (one per digit bit) instead of one __match_any_sync per round,
		// whose results the warps of a tile queue up for (a quarter of the stall samples of a pass)
		const uint32_t d = uint32_t( key[r]>>shift )&255u ;
		uint32_t peers = mask ;
#pragma unroll
		for ( int b = 0 ; b<8 ; b++ ) {
			const bool bit = ( d>>b )&1u ;
			const uint32_t bal = __ballot_sync( 0xffffffffu, act && bit ) ;
			peers &= bit ? bal : ~bal ;
		}
		pos[r] = act ? peers : 0u ;
#else
		pos[r] = act ? __match_any_sync( mask, uint32_t( key[r]>>shift )&255u ) : 0u ;
#endif
	}
#pragma unroll
	for ( int r = 0 ; r<RTX_OS_ITEMS ; r++ ) {
		const bool act = base+uint32_t( r )*32u+lane<n ;
		const uint32_t mask = __ballot_sync( 0xffffffffu, act ) ;
		if ( act ) {
			const uint32_t peers = pos[r] ;
			const uint32_t rank = __popc( peers&( ( 1u<<lane )-1u ) ) ;
			uint32_t old = 0 ;
			if ( rank == 0 ) old = atomicAdd( &wh[warp][uint32_t( key[r]>>shift )&255u], uint32_t( __popc( peers ) ) ) ;
			old = __shfl_sync( mask, old, __ffs( peers )-1 ) ;
			pos[r] = old+rank ;
		}
	}
	__syncthreads() ;
	// digit tid: the warps' counts -> offsets, the tile's count
	uint32_t total = 0 ;
#pragma unroll
	for ( int w = 0 ; w<RTX_OS_WARPS ; w++ ) { const uint32_t c = wh[w][tid] ; wh[w][tid] = total ; total += c ; }
	// publish it for the tiles behind
	volatile uint32_t* st = status ;
	if ( tile == 0 ) st[tid] = total|RTX_OS_INCL ;
	else st[size_t( tile )*256u+tid] = total|RTX_OS_AGG ;
	// exclusive scan of the counts over the digits: the tile's bucket order
	uint32_t incl = total ;
#pragma unroll
	for ( int o = 1 ; o<32 ; o <<= 1 ) { const uint32_t v = __shfl_up_sync( 0xffffffffu, incl, o ) ; if ( lane>=uint32_t( o ) ) incl += v ; }
	if ( lane == 31 ) wsum[warp] = incl ;
	__syncthreads() ;
	uint32_t before = 0 ;
#pragma unroll
	for ( int w = 0 ; w<RTX_OS_WARPS ; w++ ) if ( uint32_t( w )<warp ) before += wsum[w] ;
	tile_start[tid] = before+incl-total ;
	__syncthreads() ;
	// keys and indices into bucket order (shared memory) -- this needs nothing from the tiles in front, whose counts
	// have time to arrive meanwhile
#pragma unroll
	for ( int r = 0 ; r<RTX_OS_ITEMS ; r++ )
		if ( base+uint32_t( r )*32u+lane<n ) {
			const uint32_t d = uint32_t( key[r]>>shift )&255u ;
			const uint32_t q = pos[r]+tile_start[d]+wh[warp][d] ;
			stage[q] = key[r] ;
			sval[q] = val[r] ;
		}
	// how many keys of digit tid the tiles in front hold: walk back, RTX_OS_WINDOW status words per round trip
	// (independent loads), adding own counts until a tile with an inclusive count turns up (tile 0 publishes
	// one: the walk ends there at the latest); a word that is not published yet ends the round, the next one starts there
	uint32_t front = 0 ;
	if ( tile>0 ) {
		int look = int( tile )-1 ;
		bool found = false ;
		while ( ! found ) {
			uint32_t v[RTX_OS_WINDOW] ;
#pragma unroll
			for ( int j = 0 ; j<RTX_OS_WINDOW ; j++ ) v[j] = look-j>=0 ? st[size_t( look-j )*256u+tid] : RTX_OS_INCL ;
			int used = 0 ;
#pragma unroll
			for ( int j = 0 ; j<RTX_OS_WINDOW ; j++ )
				if ( ! found && used == j && ( v[j]&( RTX_OS_INCL|RTX_OS_AGG ) ) ) {
					front += v[j]&RTX_OS_COUNT ;
					found = ( v[j]&RTX_OS_INCL ) != 0u ;
					used = j+1 ;
				}
			look -= used ;
			if ( ! found && used == 0 ) __nanosleep( 64 ) ;
		}
		st[size_t( tile )*256u+tid] = ( front+total )|RTX_OS_INCL ;
	}
	gbase[tid] = bases[tid]+front ;
	__syncthreads() ;
	// out in runs: consecutive threads write consecutive elements of a bucket
	const uint32_t n_tile = min( uint32_t( RTX_OS_TILE ), n-tile0 ) ;
#pragma unroll
	for ( int k = 0 ; k<RTX_OS_ITEMS ; k++ ) {
		const uint32_t i = uint32_t( k )*RTX_OS_THREADS+tid ;
		if ( i<n_tile ) {
			const uint64_t kk = stage[i] ;
			const uint32_t d = uint32_t( kk>>shift )&255u ;
			const uint32_t dst = gbase[d]+( i-tile_start[d] ) ;
			keys_out[dst] = kk ;
			vals_out[dst] = sval[i] ;
		}
	}
}

// ---- hierarchy -----------------------------------------------------------------------
// child encoding inside the builder: >=0 inner node, <0 leaf slot ~c
__global__ void __launch_bounds__( 256 ) k_karras( const uint64_t* keys, int n, int2* child, int2* range, int* parent_inner, int* parent_leaf ) {
	const int i = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( i>=n-1 ) return ;
	int l, r, lo, hi ; bool ll, rl ;
	karras_node( keys, n, i, l, r, ll, rl, lo, hi ) ;
	child[i] = make_int2( ll ? ~l : l, rl ? ~r : r ) ;
	range[i] = make_int2( lo, hi ) ;
	if ( ll ) parent_leaf[l] = i ; else parent_inner[l] = i ;
	if ( rl ) parent_leaf[r] = i ; else parent_inner[r] = i ;
	if ( i == 0 ) parent_inner[0] = -1 ;
}

// boxes: [0, n-1) inner nodes, [n-1, 2n-1) leaf slots.  flags zeroed before launch.
__global__ void __launch_bounds__( 256 ) k_refit( const q4* plo, const q4* phi, const uint32_t* vals, int n, const int2* child, const int* parent_inner, const int* parent_leaf, q4* blo, q4* bhi, uint32_t* flags ) {
	const int j = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( j>=n ) return ;
	const uint32_t prim = vals[j] ;
	const float4 pl = __ldg( reinterpret_cast<const float4*>( plo+prim ) ), ph = __ldg( reinterpret_cast<const float4*>( phi+prim ) ) ;
	f3 lo = mk3( pl.x, pl.y, pl.z ), hi = mk3( ph.x, ph.y, ph.z ) ;
	pad_box( lo, hi ) ;
	q4 qlo = { lo.x, lo.y, lo.z, 0.f }, qhi = { hi.x, hi.y, hi.z, 0.f } ;
	*reinterpret_cast<float4*>( blo+( n-1+j ) ) = make_float4( qlo.x, qlo.y, qlo.z, 0.f ) ;
	*reinterpret_cast<float4*>( bhi+( n-1+j ) ) = make_float4( qhi.x, qhi.y, qhi.z, 0.f ) ;
	if ( n == 1 ) return ;
	int p = parent_leaf[j] ;
	int me = n-1+j ;                       // the node whose box (qlo, qhi) this thread carries upward
	while ( p>=0 ) {
		__threadfence() ;
		if ( atomicAdd( flags+p, 1u ) == 0u )
			return ;                       // the sibling subtree is not finished: its thread carries on
		// the second to arrive: own box from registers, the sibling's with two 128-bit loads from L2
		// (ld.global.cg: written by another thread, in front of its fence and its arrival)
		const int2 c = child[p] ;
		const int a = c.x<0 ? n-1+( ~c.x ) : c.x, b = c.y<0 ? n-1+( ~c.y ) : c.y ;
		const int sib = a == me ? b : a ;
		const float4 slo = __ldcg( reinterpret_cast<const float4*>( blo+sib ) ), shi = __ldcg( reinterpret_cast<const float4*>( bhi+sib ) ) ;
		qlo = { fminf( qlo.x, slo.x ), fminf( qlo.y, slo.y ), fminf( qlo.z, slo.z ), 0.f } ;
		qhi = { fmaxf( qhi.x, shi.x ), fmaxf( qhi.y, shi.y ), fmaxf( qhi.z, shi.z ), 0.f } ;
		*reinterpret_cast<float4*>( blo+p ) = make_float4( qlo.x, qlo.y, qlo.z, 0.f ) ;   // (one 128-bit store each: q4 itself only promises 4-byte alignment)
		*reinterpret_cast<float4*>( bhi+p ) = make_float4( qhi.x, qhi.y, qhi.z, 0.f ) ;
		me = p ;
		p = parent_inner[p] ;
	}
}

// One level of the wide tree: work item = (binary inner node, wide node index).  Children
// that stay inner get a fresh wide index and become work items of the next level.
// counters[0] = wide nodes allocated so far, counters[1] = size of the next frontier.
__device__ __forceinline__ void wide_item( const int2 item, int n, int leaf_max, const int2* child, const int2* range, const q4* blo, const q4* bhi, q4* nodes, int2* next, uint32_t* counters ) {
	float lo[3][RTX_WIDTH], hi[3][RTX_WIDTH] ; int ref[RTX_WIDTH] ;
	for ( int k = 0 ; k<RTX_WIDTH ; k++ ) { ref[k] = RTX_REF_EMPTY ; for ( int a = 0 ; a<3 ; a++ ) { lo[a][k] = INFINITY ; hi[a][k] = INFINITY ; } }   // an unused slot: a box no ray enters
	if ( n == 1 ) {
		ref[0] = ~0 ;   // the single primitive: slot 0, count 1
		lo[0][0] = blo[0].x ; lo[1][0] = blo[0].y ; lo[2][0] = blo[0].z ; hi[0][0] = bhi[0].x ; hi[1][0] = bhi[0].y ; hi[2][0] = bhi[0].z ;
	} else {
		int slots[RTX_WIDTH] ;
		const int ns = wide_gather( item.x, child, range, blo, bhi, n, leaf_max, slots ) ;
		for ( int k = 0 ; k<ns ; k++ ) {
			const int s = slots[k] ;
			const int b = s<0 ? n-1+( ~s ) : s ;
			const q4 bl = ldbox( blo+b ), bh = ldbox( bhi+b ) ;
			lo[0][k] = bl.x ; lo[1][k] = bl.y ; lo[2][k] = bl.z ; hi[0][k] = bh.x ; hi[1][k] = bh.y ; hi[2][k] = bh.z ;
			if ( ! wide_leaf_ref( s, range, leaf_max, ref[k] ) ) {
				const uint32_t idx = atomicAdd( counters, 1u ) ;
				next[atomicAdd( counters+1, 1u )] = make_int2( s, int( idx ) ) ;
				ref[k] = int( idx ) ;
			}
		}
	}
	for ( int h = 0 ; h<RTX_WIDTH/4 ; h++ ) {   // one 128-byte block per four children
		q4* o = nodes+size_t( item.y )*RTX_NODE_RECS+8*h ;
		for ( int a = 0 ; a<3 ; a++ ) {
			for ( int k = 4*h ; k<4*h+4 ; k++ ) box_ch( lo[a][k], hi[a][k], lo[a][k], hi[a][k] ) ;   // (centre / half extent for 4-wide nodes)
			// (128-bit stores: the node array is 128-byte aligned)
			reinterpret_cast<float4*>( o )[a]   = make_float4( lo[a][4*h], lo[a][4*h+1], lo[a][4*h+2], lo[a][4*h+3] ) ;
			reinterpret_cast<float4*>( o )[3+a] = make_float4( hi[a][4*h], hi[a][4*h+1], hi[a][4*h+2], hi[a][4*h+3] ) ;
		}
		reinterpret_cast<float4*>( o )[6] = make_float4( __int_as_float( ref[4*h] ), __int_as_float( ref[4*h+1] ), __int_as_float( ref[4*h+2] ), __int_as_float( ref[4*h+3] ) ) ;
		reinterpret_cast<float4*>( o )[7] = make_float4( 0.f, 0.f, 0.f, 0.f ) ;
	}
}
__global__ void __launch_bounds__( 128 ) k_wide_level( const int2* frontier, uint32_t n_front, int n, int leaf_max, const int2* child, const int2* range, const q4* blo, const q4* bhi, q4* nodes, int2* next, uint32_t* counters ) {
	const uint32_t w = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( w>=n_front ) return ;
	wide_item( frontier[w], n, leaf_max, child, range, blo, bhi, nodes, next, counters ) ;
}
// All levels in one cooperative launch: the frontier loop runs on the device, levels separated by
// grid barriers -- no host round trip per level (a 1 M-triangle mesh has ~12 levels, a flattened
// 100 M-triangle one ~16).  counters[0] = wide nodes allocated (1: the root), counters[1] = size of
// the next frontier (0); front0[0] = (binary root 0, wide node 0).
__global__ void __launch_bounds__( 128 ) k_wide_all( int2* front0, int2* front1, int n, int leaf_max, const int2* child, const int2* range, const q4* blo, const q4* bhi, q4* nodes, uint32_t* counters ) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid() ;
	const uint32_t tid = blockIdx.x*blockDim.x+threadIdx.x, stride = gridDim.x*blockDim.x ;
	int2* fin = front0 ; int2* fout = front1 ;
	uint32_t n_front = 1 ;
	while ( n_front ) {
		for ( uint32_t w = tid ; w<n_front ; w += stride )
			wide_item( fin[w], n, leaf_max, child, range, blo, bhi, nodes, fout, counters ) ;
		grid.sync() ;
		n_front = *reinterpret_cast<volatile uint32_t*>( counters+1 ) ;
		grid.sync() ;
		if ( tid == 0 ) counters[1] = 0 ;
		grid.sync() ;
		int2* t = fin ; fin = fout ; fout = t ;
	}
}

// ---- primitive boxes -------------------------------------------------------------------
__global__ void __launch_bounds__( 256 ) k_tri_bounds( const float* vces, const uint32_t* ices, uint32_t nt, q4* plo, q4* phi ) {
	const uint32_t f = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( f>=nt ) return ;
	const uint32_t i0 = ices[3*size_t( f )], i1 = ices[3*size_t( f )+1], i2 = ices[3*size_t( f )+2] ;
	const float* a = vces+3*size_t( i0 ) ; const float* b = vces+3*size_t( i1 ) ; const float* c = vces+3*size_t( i2 ) ;
	*reinterpret_cast<float4*>( plo+f ) = make_float4( fminf( a[0], fminf( b[0], c[0] ) ), fminf( a[1], fminf( b[1], c[1] ) ), fminf( a[2], fminf( b[2], c[2] ) ), 0.f ) ;
	*reinterpret_cast<float4*>( phi+f ) = make_float4( fmaxf( a[0], fmaxf( b[0], c[0] ) ), fmaxf( a[1], fmaxf( b[1], c[1] ) ), fmaxf( a[2], fmaxf( b[2], c[2] ) ), 0.f ) ;
}
// triangles in leaf order: (a, prim) (e1, b.x) (e2, b.y) (b.z, c) -- the edges are float differences (contract)
__global__ void __launch_bounds__( 256 ) k_pack_tris( const float* vces, const uint32_t* ices, const uint32_t* vals, uint32_t nt, q4* tris ) {
	const uint32_t j = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( j>=nt ) return ;
	const uint32_t f = vals[j] ;
	const uint32_t i0 = ices[3*size_t( f )], i1 = ices[3*size_t( f )+1], i2 = ices[3*size_t( f )+2] ;
	const float* a = vces+3*size_t( i0 ) ; const float* b = vces+3*size_t( i1 ) ; const float* c = vces+3*size_t( i2 ) ;
	float4* T = reinterpret_cast<float4*>( tris+size_t( j )*RTX_TRI_RECS ) ;   // (128-bit stores: the record array is 64-byte aligned)
	const float ax = a[0], ay = a[1], az = a[2], bx = b[0], by = b[1], bz = b[2], cx = c[0], cy = c[1], cz = c[2] ;
	T[0] = make_float4( ax, ay, az, __int_as_float( int( f ) ) ) ;
	T[1] = make_float4( bx-ax, by-ay, bz-az, bx ) ;
	T[2] = make_float4( cx-ax, cy-ay, cz-az, by ) ;
	T[3] = make_float4( bz, cx, cy, cz ) ;   // the vertices as uploaded, for the shading frame
}
// world bounds of every thing: analytic sphere c +- r; mesh = its root box corners mapped
// through the double transform.  root boxes are [lo,hi] of each thing's mesh.
__global__ void __launch_bounds__( 128 ) k_thing_bounds( const ThingTrav* trav, const ThingShade* shade, const q4* mesh_lo, const q4* mesh_hi, uint32_t n, q4* plo, q4* phi ) {
	const uint32_t k = blockIdx.x*blockDim.x+threadIdx.x ;
	if ( k>=n ) return ;
	if ( trav[k].kind == 0 ) {
		const double cx = trav[k].inv[0], cy = trav[k].inv[1], cz = trav[k].inv[2], r = fabs( trav[k].inv[3] ) ;
		plo[k] = { __double2float_rd( cx-r ), __double2float_rd( cy-r ), __double2float_rd( cz-r ), 0.f } ;
		phi[k] = { __double2float_ru( cx+r ), __double2float_ru( cy+r ), __double2float_ru( cz+r ), 0.f } ;
		return ;
	}
	const q4 lo = mesh_lo[k], hi = mesh_hi[k] ;
	double mn[3] = { 1e300, 1e300, 1e300 }, mx[3] = { -1e300, -1e300, -1e300 } ;
	for ( int c = 0 ; c<8 ; c++ ) {
		const d3 p = mk3( double( c&1 ? hi.x : lo.x ), double( c&2 ? hi.y : lo.y ), double( c&4 ? hi.z : lo.z ) ) ;
		const d3 w = xfpoint( shade[k].xf, p ) ;
		mn[0] = fmin( mn[0], w.x ) ; mn[1] = fmin( mn[1], w.y ) ; mn[2] = fmin( mn[2], w.z ) ;
		mx[0] = fmax( mx[0], w.x ) ; mx[1] = fmax( mx[1], w.y ) ; mx[2] = fmax( mx[2], w.z ) ;
	}
	plo[k] = { __double2float_rd( mn[0] ), __double2float_rd( mn[1] ), __double2float_rd( mn[2] ), 0.f } ;
	phi[k] = { __double2float_ru( mx[0] ), __double2float_ru( mx[1] ), __double2float_ru( mx[2] ), 0.f } ;
}

#endif // __CUDACC__

} // namespace rtx
