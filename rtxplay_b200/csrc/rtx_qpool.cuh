// rtx_qpool.cuh -- the path tracer as a COMPACTING ray pool (k_render_q, rtx_kernels.cuh).
//
// Why (round 2): with one ray bound to one lane (rtx_pool.cuh, RegPool) the warp votes for the
// step kind most lanes are in and the others wait -- measured on B200: 12.5 of 32 lanes active
// per instruction, the kernel bound by instruction issue x SIMD efficiency.  Here a warp owns a
// pool of RTX_QR rays (64) whose state lives in shared memory as 16-byte records any lane can
// pick up; per step kind there is a queue of ray slots, every iteration takes up to 32 rays of
// the kind that fills the most lanes, advances each by one step and queues it under its new
// kind.  The scheduling simulator (tests/hostemu emu_warpsim_pool) puts 29 lanes on a node step
// and 26-28 on the others, against 13.7 / 15-17 for the vote with one ray per lane.
//
// Record of a ray slot in shared memory: RTX_QS quads (16 bytes each; RTX_QS odd, so that the
// records of consecutive slots start in different bank groups), moved with 128-bit accesses:
//   Q0  cur, sp, tbest, nodes      next work item; stack size; best hit distance; node array of the current level
//   Q1  idir.xyz, tris             1/d of the current level; triangle array of the mesh being traversed (0: top level)
//   Q2  ood.xyz,  dd.x             o/d;                 object-space direction ...
//   Q3  ohi.xyz,  dd.y             object-space origin, high part
//   Q4  olo.xyz,  dd.z             ... low part
//   Q5  hit thing, hit triangle record, u, v
//   Q6.. traversal stack, (work item, entry distance) pairs, two per quad
// What a ray touches once or twice per segment lives in a 64-byte record in GLOBAL memory
// (L1/L2 resident: 6 KB per warp), so that the pool fits beside ~100 KB of L1:
//   C0  o.xyz, pixel      C1  d.xyz, meta      C2  throughput, stream low      C3  stream high, thing being traversed
// The step functions are templated on the slot store so that the host harness runs the very
// same code serially (tests/hostemu); results do not depend on the schedule (rtx_core.cuh:
// order-independent closest hit, one random stream per path, fixed-point radiance sums).
// Reference semantics: as rtx_pool.cuh (rtow.cxx:34-49; optx/camera_i.cu:24-114, optx/optics_i.cu:23-288).
#pragma once

#include "rtx_core.cuh"
#include "rtx_pool.cuh"

namespace rtx {

#ifndef RTX_QS
#define RTX_QS 11               // quads per slot record: 6 of state + ( RTX_QS-6 ) of stack
#endif
#ifndef RTX_QR
#define RTX_QR 64               // ray slots per warp (simulator: 64 / 96 / 128 -> 158 / 135 / 131 warp instructions per ray; measured on B200 with
                                // 11-quad records: 56 / 64 / 72 / 96 slots -> 669 / 655 / 660 / 797 ms per frame -- the slots compete with L1 for the 228 KB)
#endif
#define RTX_QSTACK ( 2*( RTX_QS-6 ) )   // stack entries per slot in shared memory (10: holds 99.6 % of all pushes of the bench scene)
#define RTX_QOVF   ( 96-RTX_QSTACK )    // further entries per slot in a global overflow area

#if defined( __CUDACC__ )
// ---- slot store on the device
struct QDev {
	uint32_t  base ;    // shared-space byte address of this warp's slot 0
	q4*       cold ;    // this warp's cold records, [RTX_QR][4]
	int32_t*  ovf ;     // this warp's overflow stack entries, [RTX_QR][RTX_QOVF][2]
	uint32_t* fault ;   // SceneDev::fault
	const q4* arena ;   // SceneDev::arena: node / triangle arrays are addressed as 16-byte offsets from it
	__device__ __forceinline__ uint32_t at( int slot, int q ) const { return base+uint32_t( slot )*( RTX_QS*16u )+uint32_t( q )*16u ; }
	__device__ __forceinline__ q4 ldq( int slot, int q ) const {
		q4 r ;
		asm volatile( "ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( r.x ), "=f"( r.y ), "=f"( r.z ), "=f"( r.w ) : "r"( at( slot, q ) ) : "memory" ) ;
		return r ;
	}
	__device__ __forceinline__ void stq( int slot, int q, const q4& v ) {
		asm volatile( "st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"( at( slot, q ) ), "f"( v.x ), "f"( v.y ), "f"( v.z ), "f"( v.w ) : "memory" ) ;
	}
	__device__ __forceinline__ void stw( int slot, int q, int w, float v ) {
		asm volatile( "st.shared.f32 [%0], %1;" :: "r"( at( slot, q )+uint32_t( w )*4u ), "f"( v ) : "memory" ) ;
	}
	__device__ __forceinline__ q4 ldc( int slot, int q ) const {
		q4 r ;
		asm volatile( "ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( r.x ), "=f"( r.y ), "=f"( r.z ), "=f"( r.w ) : "l"( cold+slot*4+q ) : "memory" ) ;
		return r ;
	}
	__device__ __forceinline__ void stc( int slot, int q, const q4& v ) {
		asm volatile( "st.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"( cold+slot*4+q ), "f"( v.x ), "f"( v.y ), "f"( v.z ), "f"( v.w ) : "memory" ) ;
	}
	__device__ __forceinline__ void stcw( int slot, int q, int w, float v ) {
		asm volatile( "st.global.f32 [%0], %1;" :: "l"( reinterpret_cast<float*>( cold+slot*4+q )+w ), "f"( v ) : "memory" ) ;
	}
	// stack: entries below RTX_QSTACK live in the slot record; the rare deeper ones go through
	// out-of-line code into a global overflow area (99.6 % of the pushes of the bench scene stay below 10)
	__device__ __noinline__ void push_deep( int slot, int32_t sp, int32_t v, float t ) {
		if ( sp<RTX_QSTACK+RTX_QOVF ) { int32_t* e = ovf+( size_t( slot )*RTX_QOVF+( sp-RTX_QSTACK ) )*2 ; e[0] = v ; e[1] = __float_as_int( t ) ; }
		else stack_fault( fault ) ;
	}
	__device__ __noinline__ int32_t pop_deep( int slot, int32_t sp, float& t ) {
		if ( sp>=RTX_QSTACK+RTX_QOVF ) { t = 0.f ; return RTX_STK_DONE ; }   // (beyond the overflow area: the fault flag is already set)
		const int32_t* e = ovf+( size_t( slot )*RTX_QOVF+( sp-RTX_QSTACK ) )*2 ;
		t = __int_as_float( e[1] ) ;
		return e[0] ;
	}
	__device__ __forceinline__ void push( int slot, int32_t& sp, int32_t v, float t ) {
		if ( sp<RTX_QSTACK )
			asm volatile( "st.shared.v2.b32 [%0], {%1,%2};" :: "r"( at( slot, 6 )+uint32_t( sp )*8u ), "r"( v ), "r"( __float_as_int( t ) ) : "memory" ) ;
		else
			push_deep( slot, sp, v, t ) ;
		sp++ ;
	}
	// the up to three pushes of a node step (far to near, misses carry +inf) as straight-line
	// predicated stores when all three fit the record
	__device__ __forceinline__ void push3( int slot, int32_t& sp, int32_t c1, float t1, int32_t c2, float t2, int32_t c3, float t3 ) {
		if ( sp<=RTX_QSTACK-3 ) {
			const uint32_t a = at( slot, 6 )+uint32_t( sp )*8u ;
			uint32_t n ;
			asm volatile( "{\n\t.reg .pred q3, q2, q1;\n\t.reg .u32 a2, a1, k;\n\t"
				"setp.lt.s32 q3, %7, 0x7F800000;\n\t"     // (entry distances are positive floats or +inf: compare the bit patterns)
				"setp.lt.s32 q2, %5, 0x7F800000;\n\t"
				"setp.lt.s32 q1, %3, 0x7F800000;\n\t"
				"@q3 st.shared.v2.b32 [%1], {%6,%7};\n\t"
				"selp.u32 k, 8, 0, q3;\n\tadd.u32 a2, %1, k;\n\t"
				"@q2 st.shared.v2.b32 [a2], {%4,%5};\n\t"
				"selp.u32 k, 8, 0, q2;\n\tadd.u32 a1, a2, k;\n\t"
				"@q1 st.shared.v2.b32 [a1], {%2,%3};\n\t"
				"selp.u32 k, 8, 0, q1;\n\tadd.u32 a1, a1, k;\n\t"
				"sub.u32 %0, a1, %1;\n\t}"
				: "=r"( n ) : "r"( a ), "r"( c1 ), "r"( __float_as_int( t1 ) ), "r"( c2 ), "r"( __float_as_int( t2 ) ), "r"( c3 ), "r"( __float_as_int( t3 ) ) : "memory" ) ;
			sp += int32_t( n>>3 ) ;
		} else {
			if ( t3<INFINITY ) push( slot, sp, c3, t3 ) ;
			if ( t2<INFINITY ) push( slot, sp, c2, t2 ) ;
			if ( t1<INFINITY ) push( slot, sp, c1, t1 ) ;
		}
	}
	__device__ __forceinline__ int32_t pop( int slot, int32_t& sp, float& t ) {
		sp-- ;
		if ( sp<RTX_QSTACK ) {
			int32_t v, tb ;
			asm volatile( "ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"( v ), "=r"( tb ) : "r"( at( slot, 6 )+uint32_t( sp )*8u ) : "memory" ) ;
			t = __int_as_float( tb ) ;
			return v ;
		}
		return pop_deep( slot, sp, t ) ;
	}
	// node / triangle arrays as one word: 16-byte offset from the arena (0: none)
	__device__ __forceinline__ float enc( int, const q4* p ) const { return __uint_as_float( p ? uint32_t( p-arena ) : 0u ) ; }
	__device__ __forceinline__ const q4* dec( int, float w ) const { return arena+__float_as_uint( w ) ; }
} ;
#endif

template <class P> RTX_HD void qpush_far_children( P& p, int slot, int32_t& sp, int32_t c1, float t1, int32_t c2, float t2, int32_t c3, float t3 ) {
	if ( t3<INFINITY ) p.push( slot, sp, c3, t3 ) ;
	if ( t2<INFINITY ) p.push( slot, sp, c2, t2 ) ;
	if ( t1<INFINITY ) p.push( slot, sp, c1, t1 ) ;
}
#if defined( __CUDACC__ )
__device__ __forceinline__ void qpush_far_children( QDev& p, int slot, int32_t& sp, int32_t c1, float t1, int32_t c2, float t2, int32_t c3, float t3 ) {
	p.push3( slot, sp, c1, t1, c2, t2, c3, t3 ) ;
}
#endif

// ---- NODE: four slab tests, nearest child next, the other hits pushed far to near
RTX_HD int qkind_of( int32_t cur, bool top ) {
	if ( uint32_t( cur )<uint32_t( RTX_REF_EMPTY ) ) return K_NODE ;
	if ( cur == RTX_STK_DONE ) return K_SHADE ;
	return top ? K_THING : K_LEAF ;
}

// world-space traversal state of a fresh ray (o, d and the path words go to the cold record)
template <class P> RTX_HD int qbegin_ray( P& p, int slot, const SceneDev& S, const f3& o, const f3& d, float pix_w, float meta_w ) {
	p.stc( slot, 0, mkq( o.x, o.y, o.z, pix_w ) ) ;
	p.stc( slot, 1, mkq( d.x, d.y, d.z, meta_w ) ) ;
	const f3 idir = mk3( safe_rcp( d.x ), safe_rcp( d.y ), safe_rcp( d.z ) ) ;
	p.stq( slot, 1, mkq( idir.x, idir.y, idir.z, p.enc( slot, nullptr ) ) ) ;
	p.stq( slot, 2, mkq( o.x*idir.x, o.y*idir.y, o.z*idir.z, 0.f ) ) ;
	p.stq( slot, 5, mkq( ibits( -1 ), ibits( -1 ), 0.f, 0.f ) ) ;
	int32_t sp = 0 ;
	p.push( slot, sp, RTX_STK_DONE, 0.f ) ;
	const int32_t cur = S.n_things ? 0 : RTX_STK_DONE ;   // 0 = root of the top level
	p.stq( slot, 0, mkq( ibits( cur ), ibits( sp ), INFINITY, p.enc( slot, S.tlas_nodes ) ) ) ;
	if ( S.n_things ) prefetch_line( S.tlas_nodes ) ;
	return qkind_of( cur, true ) ;
}

// pop the next work item (see pop_next in rtx_pool.cuh); leaving a mesh restores the world-space ray
template <class P> RTX_HD int32_t qpop_next( P& p, int slot, const SceneDev& S, int32_t& sp, float tbest_s, float& nodes_w, float& tris_w ) {
	while ( true ) {
		float t ;
		int32_t cur ;
		do cur = p.pop( slot, sp, t ) ; while ( t>tbest_s ) ;   // (the sentinels carry distance 0)
		if ( cur != RTX_STK_RETURN )
			return cur ;
		const q4 c0 = p.ldc( slot, 0 ), c1 = p.ldc( slot, 1 ) ;
		const f3 idir = mk3( safe_rcp( c1.x ), safe_rcp( c1.y ), safe_rcp( c1.z ) ) ;
		tris_w = p.enc( slot, nullptr ) ;
		nodes_w = p.enc( slot, S.tlas_nodes ) ;
		p.stq( slot, 1, mkq( idir.x, idir.y, idir.z, tris_w ) ) ;
		p.stq( slot, 2, mkq( c0.x*idir.x, c0.y*idir.y, c0.z*idir.z, 0.f ) ) ;
	}
}

// store the next work item, prefetch what it will read (the ray waits in a queue meanwhile)
template <class P> RTX_HD int qfinish_step( P& p, int slot, int32_t cur, int32_t sp, float tbest, float nodes_w, float tris_w ) {
	p.stq( slot, 0, mkq( ibits( cur ), ibits( sp ), tbest, nodes_w ) ) ;
	const bool top = bitsu( tris_w ) == 0u ;
	const int kind = qkind_of( cur, top ) ;
#if defined( RTX_Q_PREFETCH )   // (measured: 683 ms per frame with these prefetches, 655 without -- the L1 left beside the slots is too small to hold them)
	if ( kind == K_NODE )
		prefetch_line( p.dec( slot, nodes_w )+size_t( cur )*RTX_NODE_RECS ) ;
	else if ( kind == K_LEAF )
		prefetch_line( p.dec( slot, tris_w )+size_t( uint32_t( ~cur )>>3 )*RTX_TRI_RECS ) ;
#endif
	return kind ;
}

template <class P> RTX_HD int qstep_node( P& p, int slot, const SceneDev& S ) {
	const q4 q0 = p.ldq( slot, 0 ), q1 = p.ldq( slot, 1 ), q2 = p.ldq( slot, 2 ) ;
	int32_t cur = asint( q0.x ), sp = asint( q0.y ) ;
	const float tbest = q0.z ;
	float nodes_w = q0.w, tris_w = q1.w ;
	const f3 idir = mk3( q1.x, q1.y, q1.z ), ood = mk3( q2.x, q2.y, q2.z ) ;
	const float tbest_s = tbest*RTX_SLACK, tmin = 1e-3f ;
	const q4* n = p.dec( slot, nodes_w )+size_t( cur )*RTX_NODE_RECS ;
	RTX_COUNT( nodes ) ;
	const o8 n01 = ldo( n ), n23 = ldo( n+2 ), n45 = ldo( n+4 ), n67 = ldo( n+6 ) ;
	const q4 lx = n01.a, ly = n01.b, lz = n23.a, hx = n23.b, hy = n45.a, hz = n45.b, rf = n67.a ;
	int32_t c0 = asint( rf.x ), c1 = asint( rf.y ), c2 = asint( rf.z ), c3 = asint( rf.w ) ;
#if defined( __CUDA_ARCH__ )
	float t0, t1, t2, t3 ;
	slab4( lx, ly, lz, hx, hy, hz, idir, ood, tmin, tbest_s, t0, t1, t2, t3 ) ;
#else
	float t0 = slab( lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, idir, ood, tmin, tbest_s ) ;
	float t1 = slab( lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, idir, ood, tmin, tbest_s ) ;
	float t2 = slab( lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, idir, ood, tmin, tbest_s ) ;
	float t3 = slab( lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, idir, ood, tmin, tbest_s ) ;
#endif
#if ! defined( RTX_SORT_MINMAX )
#define RTX_CSWAP( ta, ca, tb, cb ) if ( tb<ta ) { const float tt = ta ; ta = tb ; tb = tt ; const int32_t cc = ca ; ca = cb ; cb = cc ; }
#else
	// (measured alternative: compare-exchange as min / max on the distances and two selects on the references --
	// 599.9 against 594.4 ms per frame for the predicated swaps)
#define RTX_CSWAP( ta, ca, tb, cb ) { const bool sw_ = tb<ta ; const float lo_ = fminf( ta, tb ), hi_ = fmaxf( ta, tb ) ; const int32_t cl_ = sw_ ? cb : ca, ch_ = sw_ ? ca : cb ; ta = lo_ ; tb = hi_ ; ca = cl_ ; cb = ch_ ; }
#endif
	RTX_CSWAP( t0, c0, t1, c1 ) RTX_CSWAP( t2, c2, t3, c3 ) RTX_CSWAP( t0, c0, t2, c2 ) RTX_CSWAP( t1, c1, t3, c3 ) RTX_CSWAP( t1, c1, t2, c2 )
#undef RTX_CSWAP
	if ( t0 == INFINITY )
		cur = qpop_next( p, slot, S, sp, tbest_s, nodes_w, tris_w ) ;
	else {
		qpush_far_children( p, slot, sp, c1, t1, c2, t2, c3, t3 ) ;
		cur = c0 ;
	}
	return qfinish_step( p, slot, cur, sp, tbest, nodes_w, tris_w ) ;
}

// ---- LEAF: the triangles of a mesh leaf
template <class P> RTX_HD int qstep_leaf( P& p, int slot, const SceneDev& S ) {
	const q4 q0 = p.ldq( slot, 0 ), q1 = p.ldq( slot, 1 ), q2 = p.ldq( slot, 2 ), q3 = p.ldq( slot, 3 ), q4_ = p.ldq( slot, 4 ), q5 = p.ldq( slot, 5 ) ;
	int32_t cur = asint( q0.x ), sp = asint( q0.y ) ;
	float tbest = q0.z ;
	float nodes_w = q0.w, tris_w = q1.w ;
	const uint32_t ref = uint32_t( ~cur ) ;
	const uint32_t first = ref>>3, count = ( ref&7u )+1u ;
	const f3 ohi = mk3( q3.x, q3.y, q3.z ), olo = mk3( q4_.x, q4_.y, q4_.z ), dd = mk3( q2.w, q3.w, q4_.w ) ;
	const q4* tris = p.dec( slot, tris_w ) ;
	int32_t b_thing = asint( q5.x ), b_slot = asint( q5.y ) ;
	float b_u = q5.z, b_v = q5.w ;
	int32_t level = -2 ;   // the thing being traversed: fetched when a triangle is hit
	bool changed = false ;
	RTX_COUNT( leaves ) ;
	for ( uint32_t k = 0 ; k<count ; k++ ) {
		RTX_COUNT( tris ) ;
		const q4* T = tris+size_t( first+k )*RTX_TRI_RECS ;
		const o8 t01 = ldo_tri( T ), t23 = ldo_tri( T+2 ) ;
		const q4 a = t01.a, b = t01.b, c = t23.a ;
		float t, u, v ;
		if ( tri_test( mk3( a.x, a.y, a.z ), mk3( b.x, b.y, b.z ), mk3( c.x, c.y, c.z ), ohi, olo, dd, 1e-3f, t, u, v ) ) {
			bool take = t<tbest ;
			if ( ! take && t == tbest ) {
				// exact tie: the later-listed (thing, primitive) wins (rtx_core.cuh better())
				if ( level == -2 ) level = asint( p.ldc( slot, 3 ).y ) ;
				if ( level>b_thing ) take = true ;
				else if ( level == b_thing ) {
					const int32_t b_prim = b_slot<0 ? -1 : asint( ldq( tris+size_t( b_slot )*RTX_TRI_RECS ).w ) ;
					take = asint( a.w )>b_prim ;
				}
			}
			if ( take ) {
				if ( level == -2 ) level = asint( p.ldc( slot, 3 ).y ) ;
				tbest = t ; b_thing = level ; b_slot = int32_t( first+k ) ; b_u = u ; b_v = v ;
				changed = true ;
			}
		}
	}
	if ( changed )
		p.stq( slot, 5, mkq( ibits( b_thing ), ibits( b_slot ), b_u, b_v ) ) ;
	cur = qpop_next( p, slot, S, sp, tbest*RTX_SLACK, nodes_w, tris_w ) ;
	return qfinish_step( p, slot, cur, sp, tbest, nodes_w, tris_w ) ;
}

// ---- THING: a top-level leaf (one thing): analytic sphere, or enter its mesh
template <class P> RTX_HD int qstep_thing( P& p, int slot, const SceneDev& S ) {
	const q4 q0 = p.ldq( slot, 0 ) ;
	int32_t cur = asint( q0.x ), sp = asint( q0.y ) ;
	float tbest = q0.z ;
	float nodes_w = q0.w, tris_w = p.enc( slot, nullptr ) ;
	const uint32_t first = uint32_t( ~cur )>>3 ;
	const int32_t k = int32_t( RTX_LDG( S.tlas_order+first ) ) ;
	RTX_COUNT( things ) ;
	const q4 c0 = p.ldc( slot, 0 ), c1 = p.ldc( slot, 1 ) ;
	const f3 o = mk3( c0.x, c0.y, c0.z ), d = mk3( c1.x, c1.y, c1.z ) ;
	if ( bsphere_miss( ldq( S.bsphere+k ), o, d, 1e-3f, tbest ) ) {
		RTX_COUNT( spheres ) ;   // (harness: counts culled visits)
		cur = qpop_next( p, slot, S, sp, tbest*RTX_SLACK, nodes_w, tris_w ) ;
		return qfinish_step( p, slot, cur, sp, tbest, nodes_w, tris_w ) ;
	}
	const char* tt = reinterpret_cast<const char*>( S.trav+k ) ;
	const o8 A = ldo_rec( tt ), D = ldo_rec( tt+96 ) ;   // inv[0..3]; nodes, tris, kind, n_tris, diag, pad
	const double m0 = dbl( A.a.x, A.a.y ), m1 = dbl( A.a.z, A.a.w ), m2 = dbl( A.b.x, A.b.y ), m3 = dbl( A.b.z, A.b.w ) ;
	if ( asint( D.b.x ) == 0 ) {
		double td ;
		if ( sphere_root( mk3( m0, m1, m2 ), m3, wide( o ), wide( d ), double( 1e-3f ), td ) ) {
			const float t = float( td ) ;
			const int32_t b_thing = asint( p.ldq( slot, 5 ).x ) ;
			if ( t<tbest || ( t == tbest && k>b_thing ) ) {   // better(): a thing is visited once per ray, so equal things never meet here
				tbest = t ;
				p.stq( slot, 5, mkq( ibits( k ), ibits( -1 ), 0.f, 0.f ) ) ;
			}
		}
		cur = qpop_next( p, slot, S, sp, tbest*RTX_SLACK, nodes_w, tris_w ) ;
		return qfinish_step( p, slot, cur, sp, tbest, nodes_w, tris_w ) ;
	}
	// enter the mesh: object-space ray, origin in double carried as hi+lo
	const o8 B = ldo_rec( tt+32 ), C = ldo_rec( tt+64 ) ;
	d3 od, ddd ;
	if ( asint( D.b.z ) ) {
		const double m5 = dbl( B.a.z, B.a.w ), m7 = dbl( B.b.z, B.b.w ), m10 = dbl( C.b.x, C.b.y ), m11 = dbl( C.b.z, C.b.w ) ;
		od = xfpoint_diag( m0, m3, m5, m7, m10, m11, wide( o ) ) ;
		ddd = xfvec_diag( m0, m5, m10, wide( d ) ) ;
	} else {
		double m[12] ;
		m[0] = m0 ; m[1] = m1 ; m[2] = m2 ; m[3] = m3 ;
		m[4] = dbl( B.a.x, B.a.y ) ; m[5] = dbl( B.a.z, B.a.w ) ; m[6] = dbl( B.b.x, B.b.y ) ; m[7] = dbl( B.b.z, B.b.w ) ;
		m[8] = dbl( C.a.x, C.a.y ) ; m[9] = dbl( C.a.z, C.a.w ) ; m[10] = dbl( C.b.x, C.b.y ) ; m[11] = dbl( C.b.z, C.b.w ) ;
		od = xfpoint( m, wide( o ) ) ;
		ddd = xfvec( m, wide( d ) ) ;
	}
	const f3 ohi = narrow( od ) ;
	const f3 olo = narrow( od-wide( ohi ) ) ;
	const f3 dd  = narrow( ddd ) ;
	const f3 idir = mk3( safe_rcp( dd.x ), safe_rcp( dd.y ), safe_rcp( dd.z ) ) ;
#if defined( __CUDA_ARCH__ )
	const q4* nodes = reinterpret_cast<const q4*>( ( unsigned long long )__float_as_uint( D.a.x )|( ( unsigned long long )__float_as_uint( D.a.y )<<32 ) ) ;
	const q4* tris  = reinterpret_cast<const q4*>( ( unsigned long long )__float_as_uint( D.a.z )|( ( unsigned long long )__float_as_uint( D.a.w )<<32 ) ) ;
#else
	const q4* nodes = S.trav[k].nodes ; const q4* tris = S.trav[k].tris ;
#endif
	nodes_w = p.enc( slot, nodes ) ; tris_w = p.enc( slot, tris ) ;
	p.stq( slot, 1, mkq( idir.x, idir.y, idir.z, tris_w ) ) ;
	p.stq( slot, 2, mkq( ohi.x*idir.x, ohi.y*idir.y, ohi.z*idir.z, dd.x ) ) ;
	p.stq( slot, 3, mkq( ohi.x, ohi.y, ohi.z, dd.y ) ) ;
	p.stq( slot, 4, mkq( olo.x, olo.y, olo.z, dd.z ) ) ;
	p.stcw( slot, 3, 1, ibits( k ) ) ;
	p.push( slot, sp, RTX_STK_RETURN, 0.f ) ;
	return qfinish_step( p, slot, 0, sp, tbest, nodes_w, tris_w ) ;
}

// ---- SHADE: the ray is finished (see step_shade in rtx_pool.cuh).  Returns the kind of the
// path's next ray, or K_REGEN (path ended; `c` is its colour, `pix` its pixel).
template <class P> RTX_HD int qstep_shade( P& p, int slot, const SceneDev& S, f3& c, uint32_t& pix, bool& guide, f3& gnormal, f3& galbedo, uint32_t& segments ) {
	guide = false ;
	RTX_COUNT( rays ) ;
	const q4 q5 = p.ldq( slot, 5 ) ;
	const q4 c0 = p.ldc( slot, 0 ), c1 = p.ldc( slot, 1 ), c2 = p.ldc( slot, 2 ) ;
	const int32_t h_thing = asint( q5.x ), h_slot = asint( q5.y ) ;
	const f3 ori = mk3( c0.x, c0.y, c0.z ), dir = mk3( c1.x, c1.y, c1.z ) ;
	f3 thr = mk3( c2.x, c2.y, c2.z ) ;
	pix = bitsu( c0.w ) ;
	const uint32_t meta = bitsu( c1.w )+512u ;   // bits 0-7 depth left, 8 guide taken, 9-15 segments of the path
	const uint32_t depth_left = meta&255u ;
	segments = meta>>9 ;
	RTX_COUNT_LIVE( segments-1u ) ;
	c = mk3( 0.f, 0.f, 0.f ) ;
	if ( h_thing<0 ) {
		c = thr*sky( dir ) ;
		return K_REGEN ;
	}
	if ( depth_left == 0 && S.variant != RTX_SEM_RTWO_I )
		return K_REGEN ;
	Frame fr ;
	o8 mat ;
	qframe_of( S, h_thing, h_slot, q5.z, q5.w, ori, dir, 1e-3f, fr, mat ) ;
	Pcg rng ;
	rng.state = uint64_t( bitsu( c2.w ) )|( uint64_t( bitsu( p.ldc( slot, 3 ).x ) )<<32 ) ;
	f3 att, out ;
	const bool go = qscatter( mat, dir, fr, rng, att, out, S.variant == RTX_SEM_RTOW ) ;
	uint32_t meta2 = meta ;
	if ( ! ( meta&256u ) && asint( mat.b.y ) != 2 ) {
		guide = true ; gnormal = fr.normal ; galbedo = att ;
		meta2 |= 256u ;
	}
	if ( ! go )
		return K_REGEN ;
	thr = thr*att ;
	if ( depth_left == 0 ) {   // RTX_SEM_RTWO_I: the last ray scattered, its throughput is the colour (optx/camera_i.cu:92-95)
		c = thr ;
		return K_REGEN ;
	}
	p.stc( slot, 2, mkq( thr.x, thr.y, thr.z, ubits( uint32_t( rng.state ) ) ) ) ;
	p.stcw( slot, 3, 0, ubits( uint32_t( rng.state>>32 ) ) ) ;
	return qbegin_ray( p, slot, S, fr.p, out, c0.w, ubits( meta2-1u ) ) ;
}

// ---- REGEN: a new path
template <class P> RTX_HD int qstep_regen( P& p, int slot, const SceneDev& S, const CameraDev& cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h,
		uint32_t tile_pixel, uint64_t seed, uint32_t sample, uint32_t depth ) {
	Pcg rng ;
	rng.seed( seed, w*y+x, sample ) ;
	f3 ori, dir ;
	primary_ray( cam, x, y, w, h, rng, ori, dir, S.variant != RTX_SEM_RTOW ) ;
	depth = sem_depth( S.variant, depth ) ;
	p.stc( slot, 2, mkq( 1.f, 1.f, 1.f, ubits( uint32_t( rng.state ) ) ) ) ;
	p.stc( slot, 3, mkq( ubits( uint32_t( rng.state>>32 ) ), ibits( -1 ), 0.f, 0.f ) ) ;
	return qbegin_ray( p, slot, S, ori, dir, ubits( tile_pixel ), ubits( depth>255u ? 255u : depth ) ) ;
}

} // namespace rtx
