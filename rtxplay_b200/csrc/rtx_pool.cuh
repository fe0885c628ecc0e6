// rtx_pool.cuh -- the path tracer as a warp-scheduled state machine.
//
// Why: traced rays need very different work -- inner nodes, triangle leaves, instance
// entries, shading, new paths -- and a warp that binds one ray to one lane spends most of
// its issue slots on a handful of lanes (measured on B200, ncu: 6.9 of 32 lanes active,
// profiles/r01_v0_*).  Here a ray is always in one of five states, its whole state sits in
// the registers of its lane (RegPool; only the traversal stack is a lane-private column of
// shared memory), and each iteration the warp votes for the step kind most lanes can take;
// the lanes in that state advance their ray by one step:
//     NODE   test the four children of a wide BVH node, push / pop
//     LEAF   test the 1..4 triangles of a mesh leaf
//     THING  top-level leaf: analytic sphere test, or transform the ray into a mesh
//     SHADE  ray finished: sky or scatter, start the next ray of the path
//     REGEN  path finished: fetch the next (pixel, sample) of the warp's tile
// Rays of other kinds wait; while a ray waits for a leaf step, its triangles are being
// prefetched.  Results do not depend on the schedule: closest hits are order-independent
// (rtx_core.cuh better()), every path owns its random stream, and radiance is summed in
// fixed point.  (RTX_K > 1 selects the measured alternative: RTX_K lane-private ray slots per
// lane in shared memory, DevPool -- more candidates per vote, but 14-16 warps per SM and no L1:
// 890-950 ms per frame against 678)
//
// The step functions are templated on the slot store so that the host harness can run the
// very same code serially (tests/hostemu).
#pragma once

#include "rtx_core.cuh"

namespace rtx {

#ifndef RTX_K
#define RTX_K 1                 // ray slots per lane (1: state in registers; >1: in shared memory)
#endif
#if RTX_K == 1 && ! defined( RTX_REGPOOL )
#define RTX_REGPOOL 1
#endif
#ifndef RTX_LEAN_POP
#define RTX_LEAN_POP 1          // pop_next: a bare pop-and-cull inner loop, the sentinels looked at outside of it (r2: 681.1 -> 666.9 ms)
#endif
#ifndef RTX_FFMA2
#define RTX_FFMA2 1             // node step: the 24 multiply-adds of the four slab tests as 12 packed FFMA2 (rtx_core.cuh slab4; r2: 681.1 -> 674.4 ms)
#endif
#ifndef RTX_FAST_PUSH
#define RTX_FAST_PUSH 1         // node step: the three pushes as predicated straight-line stores (RegPool::push3; r2: 681.1 -> 666.6 ms; all three: 648.7)
#endif
#ifndef RTX_SHADE_INLINE
#define RTX_SHADE_INLINE 1      // shading frame and scattering inline on thing records fetched with 256-bit loads (rtx_core.cuh qframe_of / qscatter) instead of the
                                // __noinline__ frame_of / scatter, whose structures travel through local memory (54 LDL/STL + 29 narrow LDG per segment in the r1 profile)
#endif
#ifndef RTX_OVF_LAZY
#define RTX_OVF_LAZY 0          // 1: the lane's overflow-stack address computed in (out-of-line) deep push / pop functions instead of at the top of the loop -- measured slower, 591 against 582 ms per frame
#endif
#ifndef RTX_PREFETCH
#define RTX_PREFETCH 0          // L1 prefetches beyond the first line of the next leaf (bits: 1 its second line, 2 / 4 the second-nearest child when a leaf / a node)
#endif
#ifndef RTX_POOL_STACK
#define RTX_POOL_STACK 16       // stack entries (work item + its entry distance) per slot kept in shared memory
#endif
#define RTX_POOL_OVF   80       // further entries per slot in a global overflow area

// slot fields (32-bit words)
enum {
	F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ,                      // world ray
	F_HX, F_HY, F_HZ, F_LX, F_LY, F_LZ, F_EX, F_EY, F_EZ,    // object ray: origin hi, lo; direction
	F_IX, F_IY, F_IZ, F_QX, F_QY, F_QZ,                      // current level: 1/d, o/d
	F_T, F_THING, F_PRIM, F_U, F_V, F_TRIJ,                  // best hit so far (F_TRIJ: its triangle record)
	F_CUR, F_LEVEL, F_SP,                                    // next work item; thing being traversed (-1: top); stack size
	F_NODES0, F_NODES1, F_TRIS0, F_TRIS1,                    // node / triangle arrays of the current level
	F_THRX, F_THRY, F_THRZ, F_RNG0, F_RNG1, F_PIX, F_META,   // path: throughput, stream, pixel index, counters (see step_shade)
	F_STACK,
	F_WORDS = F_STACK+2*RTX_POOL_STACK
} ;

enum { K_DONE = 0, K_NODE = 1, K_LEAF = 2, K_THING = 3, K_SHADE = 4, K_REGEN = 5, K_KINDS = 6 } ;

// ---- slot store on the device: shared memory, SoA, plus the global overflow stack
#if defined( __CUDACC__ )
struct DevPool {
	uint32_t* w ;      // this warp's words, [F_WORDS][R]
	int32_t*  ovf ;    // this warp's overflow area, [R][RTX_POOL_OVF]
	uint32_t* fault ;  // SceneDev::fault
	__device__ __forceinline__ float    f( int fld, int slot ) const { return __uint_as_float( w[fld*( 32*RTX_K )+slot] ) ; }
	__device__ __forceinline__ int32_t  i( int fld, int slot ) const { return int32_t( w[fld*( 32*RTX_K )+slot] ) ; }
	__device__ __forceinline__ void     sf( int fld, int slot, float v ) { w[fld*( 32*RTX_K )+slot] = __float_as_uint( v ) ; }
	__device__ __forceinline__ void     si( int fld, int slot, int32_t v ) { w[fld*( 32*RTX_K )+slot] = uint32_t( v ) ; }
	__device__ __forceinline__ void     push( int slot, int32_t& sp, int32_t v, float t ) {
		if ( sp<RTX_POOL_STACK ) {
			w[( F_STACK+2*sp )*( 32*RTX_K )+slot] = uint32_t( v ) ;
			w[( F_STACK+2*sp+1 )*( 32*RTX_K )+slot] = __float_as_uint( t ) ;
		} else if ( sp<RTX_POOL_STACK+RTX_POOL_OVF ) {
			ovf[( slot*RTX_POOL_OVF+( sp-RTX_POOL_STACK ) )*2] = v ;
			ovf[( slot*RTX_POOL_OVF+( sp-RTX_POOL_STACK ) )*2+1] = __float_as_int( t ) ;
		} else
			stack_fault( fault ) ;
		sp++ ;
	}
	__device__ __forceinline__ int32_t  pop( int slot, int32_t& sp, float& t ) {
		sp-- ;
		if ( sp<RTX_POOL_STACK ) {
			t = __uint_as_float( w[( F_STACK+2*sp+1 )*( 32*RTX_K )+slot] ) ;
			return int32_t( w[( F_STACK+2*sp )*( 32*RTX_K )+slot] ) ;
		}
		if ( sp>=RTX_POOL_STACK+RTX_POOL_OVF ) { t = 0.f ; return RTX_STK_DONE ; }
		t = __int_as_float( ovf[( slot*RTX_POOL_OVF+( sp-RTX_POOL_STACK ) )*2+1] ) ;
		return ovf[( slot*RTX_POOL_OVF+( sp-RTX_POOL_STACK ) )*2] ;
	}
} ;
#endif

#if defined( __CUDACC__ )
// one ray per lane, state in registers (every field index is a compile-time constant), only
// the stack in shared memory: no shared-memory traffic for the state, no limit on resident
// warps from it, at the price of fewer candidate rays per vote (RTX_K must be 1)
// Fields that are touched a few times per ray only can live in shared memory instead (a
// lane-private column next to the stack, conflict free): RTX_COLD selects how many groups --
//   1: path state (throughput, stream, pixel, counters) and the hit details (u, v, record, prim)
//   2: + the world-space ray (read when a thing is entered or left, and by shading)
//   3: + the low part of the object-space origin (read by leaf steps)
// Every field index is a compile-time constant where it is used, so the choice folds away.
#ifndef RTX_COLD
#define RTX_COLD 0
#endif
__device__ __forceinline__ constexpr int cold_slot( int fld ) {
	int n = 0 ;
#if RTX_COLD >= 1
	if ( fld>=F_THRX && fld<=F_META ) return n+fld-F_THRX ;
	n += F_META-F_THRX+1 ;
	if ( fld>=F_PRIM && fld<=F_TRIJ ) return n+fld-F_PRIM ;
	n += F_TRIJ-F_PRIM+1 ;
#endif
#if RTX_COLD >= 2
	if ( fld>=F_OX && fld<=F_DZ ) return n+fld-F_OX ;
	n += F_DZ-F_OX+1 ;
#endif
#if RTX_COLD >= 3
	if ( fld>=F_LX && fld<=F_LZ ) return n+fld-F_LX ;
	n += F_LZ-F_LX+1 ;
#endif
	return fld<0 ? n : -1 ;   // cold_slot( -1 ) = number of cold fields
}
#define RTX_COLD_WORDS ( cold_slot( -1 ) )
struct RegPool {
	uint32_t  r[F_STACK] ;
	uint32_t  stk ;    // shared-space byte address of this lane's stack column (entry i = the 8 bytes (ref, distance) at stk + i*256)
	uint32_t  cold ;   // shared-space byte address of this lane's column of cold fields (field s at cold + s*128)
	int32_t*  ovf_all ; // overflow entries (pairs) of all lanes of the grid; this lane's: ovf() (computed where it is needed -- the rare deep stack -- instead of living in two registers)
#if RTX_OVF_LAZY
	__device__ __forceinline__ int32_t* ovf() const { return ovf_all+( size_t( blockIdx.x )*32u+threadIdx.x )*RTX_POOL_OVF*2 ; }   // (one ray per lane: 32 columns per CTA)
#define RTX_DEEP_FN __noinline__
#else
	int32_t*  ovf_ ;
	__device__ __forceinline__ int32_t* ovf() const { return ovf_ ; }
#define RTX_DEEP_FN __forceinline__
#endif
	uint32_t* fault ;  // SceneDev::fault
	__device__ __forceinline__ uint32_t ldc( int s ) const { uint32_t v ; asm volatile( "ld.shared.u32 %0, [%1];" : "=r"( v ) : "r"( cold+uint32_t( s )*128u ) : "memory" ) ; return v ; }
	__device__ __forceinline__ void     stc( int s, uint32_t v ) { asm volatile( "st.shared.u32 [%0], %1;" :: "r"( cold+uint32_t( s )*128u ), "r"( v ) : "memory" ) ; }
	__device__ __forceinline__ float    f( int fld, int ) const { return __uint_as_float( cold_slot( fld )>=0 ? ldc( cold_slot( fld ) ) : r[fld] ) ; }
	__device__ __forceinline__ int32_t  i( int fld, int ) const { return int32_t( cold_slot( fld )>=0 ? ldc( cold_slot( fld ) ) : r[fld] ) ; }
	__device__ __forceinline__ void     sf( int fld, int, float v ) { if ( cold_slot( fld )>=0 ) stc( cold_slot( fld ), __float_as_uint( v ) ) ; else r[fld] = __float_as_uint( v ) ; }
	__device__ __forceinline__ void     si( int fld, int, int32_t v ) { if ( cold_slot( fld )>=0 ) stc( cold_slot( fld ), uint32_t( v ) ) ; else r[fld] = uint32_t( v ) ; }
	__device__ __forceinline__ void     push( int, int32_t& sp, int32_t v, float t ) {
		if ( sp<RTX_POOL_STACK ) {
			const uint32_t a = stk+uint32_t( sp )*256u ;
			asm volatile( "st.shared.v2.b32 [%0], {%1,%2};" :: "r"( a ), "r"( v ), "r"( __float_as_int( t ) ) : "memory" ) ;
		} else push_deep( sp, v, t ) ;
		sp++ ;
	}
	// (the rare deep part of the stack, out of line: inlined, its address arithmetic is hoisted to the top of the kernel's loop)
	__device__ RTX_DEEP_FN void push_deep( int32_t sp, int32_t v, float t ) {
		if ( sp<RTX_POOL_STACK+RTX_POOL_OVF ) { int32_t* o = ovf() ; o[2*( sp-RTX_POOL_STACK )] = v ; o[2*( sp-RTX_POOL_STACK )+1] = __float_as_int( t ) ; }
		else stack_fault( fault ) ;
	}
	__device__ RTX_DEEP_FN int32_t pop_deep( int32_t sp, float& t ) {
		if ( sp>=RTX_POOL_STACK+RTX_POOL_OVF ) { t = 0.f ; return RTX_STK_DONE ; }
		const int32_t* o = ovf() ;
		t = __int_as_float( o[2*( sp-RTX_POOL_STACK )+1] ) ;
		return o[2*( sp-RTX_POOL_STACK )] ;
	}
#if RTX_FAST_PUSH
	// the up to three pushes of a node step (children sorted by distance, misses = +inf last) as
	// straight-line predicated stores when all three fit the shared-memory part of the stack: 19
	// instructions instead of three branchy pushes of 12-16 each (cuobjdump), in the hottest loop
	__device__ __forceinline__ void     push3( int slot, int32_t& sp, int32_t c1, float t1, int32_t c2, float t2, int32_t c3, float t3 ) {
		if ( sp<=RTX_POOL_STACK-3 ) {
			const uint32_t a = stk+uint32_t( sp )*256u ;
			uint32_t n ;
			asm volatile( "{\n\t.reg .pred q3, q2, q1;\n\t.reg .u32 a2, a1, k;\n\t"
				"setp.lt.s32 q3, %7, 0x7F800000;\n\t"     // (entry distances are positive floats or +inf: compare the bit patterns)
				"setp.lt.s32 q2, %5, 0x7F800000;\n\t"
				"setp.lt.s32 q1, %3, 0x7F800000;\n\t"
				"@q3 st.shared.v2.b32 [%1], {%6,%7};\n\t"
				"selp.u32 k, 256, 0, q3;\n\tadd.u32 a2, %1, k;\n\t"
				"@q2 st.shared.v2.b32 [a2], {%4,%5};\n\t"
				"selp.u32 k, 256, 0, q2;\n\tadd.u32 a1, a2, k;\n\t"
				"@q1 st.shared.v2.b32 [a1], {%2,%3};\n\t"
				"selp.u32 k, 256, 0, q1;\n\tadd.u32 a1, a1, k;\n\t"
				"sub.u32 %0, a1, %1;\n\t}"
				: "=r"( n ) : "r"( a ), "r"( c1 ), "r"( __float_as_int( t1 ) ), "r"( c2 ), "r"( __float_as_int( t2 ) ), "r"( c3 ), "r"( __float_as_int( t3 ) ) : "memory" ) ;
			sp += int32_t( n>>8 ) ;
		} else {
			if ( t3<INFINITY ) push( slot, sp, c3, t3 ) ;
			if ( t2<INFINITY ) push( slot, sp, c2, t2 ) ;
			if ( t1<INFINITY ) push( slot, sp, c1, t1 ) ;
		}
	}
#endif
	__device__ __forceinline__ int32_t  pop( int, int32_t& sp, float& t ) {
		sp-- ;
		if ( sp<RTX_POOL_STACK ) {
			const uint32_t a = stk+uint32_t( sp )*256u ;
			int32_t v, tb ;
			asm volatile( "ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"( v ), "=r"( tb ) : "r"( a ) : "memory" ) ;
			t = __int_as_float( tb ) ;
			return v ;
		}
		return pop_deep( sp, t ) ;
	}
} ;
#endif

// the hits of a node step behind the nearest one, far to near (misses carry +inf)
template <class P> RTX_HD void push_far_children( P& p, int slot, int32_t& sp, int32_t c1, float t1, int32_t c2, float t2, int32_t c3, float t3 ) {
	if ( t3<INFINITY ) p.push( slot, sp, c3, t3 ) ;
	if ( t2<INFINITY ) p.push( slot, sp, c2, t2 ) ;
	if ( t1<INFINITY ) p.push( slot, sp, c1, t1 ) ;
}
#if RTX_FAST_PUSH && defined( __CUDACC__ )
__device__ __forceinline__ void push_far_children( RegPool& p, int slot, int32_t& sp, int32_t c1, float t1, int32_t c2, float t2, int32_t c3, float t3 ) {
	p.push3( slot, sp, c1, t1, c2, t2, c3, t3 ) ;
}
#endif

template <class P> RTX_HD f3 ld3( const P& p, int fld, int slot ) { return mk3( p.f( fld, slot ), p.f( fld+1, slot ), p.f( fld+2, slot ) ) ; }
template <class P> RTX_HD void st3( P& p, int fld, int slot, const f3& v ) { p.sf( fld, slot, v.x ) ; p.sf( fld+1, slot, v.y ) ; p.sf( fld+2, slot, v.z ) ; }
template <class P, class T> RTX_HD const T* ldp( const P& p, int fld, int slot ) {
	const uint64_t a = uint64_t( uint32_t( p.i( fld, slot ) ) )|( uint64_t( uint32_t( p.i( fld+1, slot ) ) )<<32 ) ;
	return reinterpret_cast<const T*>( a ) ;
}
template <class P> RTX_HD void stp( P& p, int fld, int slot, const void* ptr ) {
	const uint64_t a = reinterpret_cast<uint64_t>( ptr ) ;
	p.si( fld, slot, int32_t( uint32_t( a ) ) ) ; p.si( fld+1, slot, int32_t( uint32_t( a>>32 ) ) ) ;
}

RTX_HD void prefetch_line( const void* a ) {
#if defined( __CUDA_ARCH__ )
	asm volatile( "prefetch.global.L1 [%0];" :: "l"( a ) ) ;
#else
	( void ) a ;
#endif
}

// kind of the work item `cur` for a ray at the top level (level < 0) or inside a mesh
RTX_HD int kind_of( int32_t cur, int32_t level ) {
#if defined( __CUDA_ARCH__ ) && ! defined( RTX_KIND_BRANCHY )
	// three selects (the compiler turns the chain of returns below into three divergent branches at the end of every step)
	int k ;
	asm( "{\n\t.reg .pred pn, pd;\n\t.reg .s32 l;\n\t"
		"setp.lt.u32 pn, %1, %3;\n\t"
		"setp.eq.s32 pd, %1, %4;\n\t"
		"shr.u32 l, %2, 31;\n\tadd.s32 l, l, 2;\n\t"      // K_LEAF = 2, K_THING = 3 (level < 0)
		"selp.s32 l, 4, l, pd;\n\t"                        // K_SHADE
		"selp.s32 %0, 1, l, pn;\n\t}"                      // K_NODE
		: "=r"( k ) : "r"( cur ), "r"( level ), "n"( RTX_REF_EMPTY ), "n"( RTX_STK_DONE ) ) ;
	return k ;
#else
	if ( uint32_t( cur )<uint32_t( RTX_REF_EMPTY ) ) return K_NODE ;
	if ( cur == RTX_STK_DONE ) return K_SHADE ;
	return level<0 ? K_THING : K_LEAF ;
#endif
}
static_assert( K_NODE == 1 && K_LEAF == 2 && K_THING == 3 && K_SHADE == 4, "kind_of's selects are written for these values" ) ;

// world-space traversal state of a fresh ray
template <class P> RTX_HD void begin_ray( P& p, int slot, const SceneDev& S, const f3& o, const f3& d ) {
	st3( p, F_OX, slot, o ) ; st3( p, F_DX, slot, d ) ;
	const f3 idir = mk3( safe_rcp( d.x ), safe_rcp( d.y ), safe_rcp( d.z ) ) ;
	st3( p, F_IX, slot, idir ) ; st3( p, F_QX, slot, mk3( o.x*idir.x, o.y*idir.y, o.z*idir.z ) ) ;
	p.sf( F_T, slot, INFINITY ) ; p.si( F_THING, slot, -1 ) ; p.si( F_PRIM, slot, -1 ) ;
	p.si( F_LEVEL, slot, -1 ) ;
	stp( p, F_NODES0, slot, S.tlas_nodes ) ; stp( p, F_TRIS0, slot, nullptr ) ;
	int32_t sp = 0 ;
	p.push( slot, sp, RTX_STK_DONE, 0.f ) ;
	p.si( F_SP, slot, sp ) ;
	p.si( F_CUR, slot, S.n_things ? 0 : RTX_STK_DONE ) ;   // 0 = root of the top level
	if ( S.n_things ) prefetch_line( S.tlas_nodes ) ;
}

// pop the next work item.  Entries whose box the ray enters beyond the best hit found since
// they were pushed are dropped without touching their node; leaving a mesh (RTX_STK_RETURN)
// is handled on the way.
template <class P> RTX_HD int32_t pop_next( P& p, int slot, const SceneDev& S, int32_t& sp, int32_t& level ) {
	const float tbest_s = p.f( F_T, slot )*RTX_SLACK ;
#if RTX_LEAN_POP
	// the two sentinels sit on the stack with distance 0, so the distance cull alone ends the
	// inner loop for them too; what was popped is looked at once, outside of it
	while ( true ) {
		float t ;
		int32_t cur ;
#if defined( RTX_NO_POPCULL )
		cur = p.pop( slot, sp, t ) ;   // (experiment: no distance cull of popped entries)
#else
		do cur = p.pop( slot, sp, t ) ; while ( t>tbest_s ) ;
#endif
		if ( cur != RTX_STK_RETURN )
			return cur ;
		const f3 o = ld3( p, F_OX, slot ), d = ld3( p, F_DX, slot ) ;
		const f3 idir = mk3( safe_rcp( d.x ), safe_rcp( d.y ), safe_rcp( d.z ) ) ;
		st3( p, F_IX, slot, idir ) ; st3( p, F_QX, slot, mk3( o.x*idir.x, o.y*idir.y, o.z*idir.z ) ) ;
		level = -1 ;
		p.si( F_LEVEL, slot, -1 ) ;
		stp( p, F_NODES0, slot, S.tlas_nodes ) ; stp( p, F_TRIS0, slot, nullptr ) ;
	}
#else
	while ( true ) {
		float t ;
		const int32_t cur = p.pop( slot, sp, t ) ;
		if ( cur == RTX_STK_RETURN ) {
			const f3 o = ld3( p, F_OX, slot ), d = ld3( p, F_DX, slot ) ;
			const f3 idir = mk3( safe_rcp( d.x ), safe_rcp( d.y ), safe_rcp( d.z ) ) ;
			st3( p, F_IX, slot, idir ) ; st3( p, F_QX, slot, mk3( o.x*idir.x, o.y*idir.y, o.z*idir.z ) ) ;
			level = -1 ;
			p.si( F_LEVEL, slot, -1 ) ;
			stp( p, F_NODES0, slot, S.tlas_nodes ) ; stp( p, F_TRIS0, slot, nullptr ) ;
			continue ;
		}
		if ( cur == RTX_STK_DONE || t<=tbest_s )
			return cur ;
	}
#endif
}

// store the next work item and prefetch what it will read
template <class P> RTX_HD int finish_step( P& p, int slot, int32_t cur, int32_t sp, int32_t level ) {
	p.si( F_CUR, slot, cur ) ; p.si( F_SP, slot, sp ) ;
	const int kind = kind_of( cur, level ) ;
#if RTX_K > 1
	// (with one ray per lane the node is needed by the very next step: a prefetch buys nothing)
	if ( kind == K_NODE )
		prefetch_line( ldp<P, q4>( p, F_NODES0, slot )+size_t( cur )*RTX_NODE_RECS ) ;
	else
#endif
#if defined( __CUDA_ARCH__ ) && ! defined( RTX_KIND_BRANCHY ) && ! ( RTX_PREFETCH & 1 )
	{
		// the first triangle of a leaf: one predicated prefetch, the address computed by every lane (no branch)
		const q4* T = ldp<P, q4>( p, F_TRIS0, slot )+size_t( uint32_t( ~cur )>>3 )*RTX_TRI_RECS ;
		asm volatile( "{\n\t.reg .pred pl;\n\tsetp.eq.s32 pl, %1, 2;\n\t@pl prefetch.global.L1 [%0];\n\t}" :: "l"( T ), "r"( kind ) ) ;
		return kind ;
	}
#endif
	if ( kind == K_LEAF ) {
		const q4* T = ldp<P, q4>( p, F_TRIS0, slot )+size_t( uint32_t( ~cur )>>3 )*RTX_TRI_RECS ;
		prefetch_line( T ) ;
#if RTX_PREFETCH & 1
		// a leaf of three 64-byte records spans two or three 128-byte lines
		if ( ( uint32_t( ~cur )&7u )>=1u ) prefetch_line( T+2*RTX_TRI_RECS-1 ) ;
#endif
	}
	return kind ;
}

// ---- NODE: four slab tests, nearest child next, the other hits pushed far to near
template <class P> RTX_HD int step_node( P& p, int slot, const SceneDev& S ) {
	int32_t cur = p.i( F_CUR, slot ), sp = p.i( F_SP, slot ), level = p.i( F_LEVEL, slot ) ;
	const f3 idir = ld3( p, F_IX, slot ), ood = ld3( p, F_QX, slot ) ;
	const float tbest_s = p.f( F_T, slot )*RTX_SLACK ;
	const float tmin = 1e-3f ;
	const q4* n = ldp<P, q4>( p, F_NODES0, slot )+size_t( cur )*RTX_NODE_RECS ;
	RTX_COUNT( nodes ) ;
#if RTX_WIDTH == 8
	{
		float t8[8] ; int32_t c8[8] ;
		node_children8( n, idir, ood, tmin, tbest_s, t8, c8 ) ;
		if ( t8[0] == INFINITY )
			cur = pop_next( p, slot, S, sp, level ) ;
		else {
#pragma unroll
			for ( int k = 7 ; k>0 ; k-- ) if ( t8[k]<INFINITY ) p.push( slot, sp, c8[k], t8[k] ) ;
			cur = c8[0] ;
		}
		return finish_step( p, slot, cur, sp, level ) ;
	}
#endif
	const o8 n01 = ldo( n ), n23 = ldo( n+2 ), n45 = ldo( n+4 ), n67 = ldo( n+6 ) ;
	const q4 lx = n01.a, ly = n01.b, lz = n23.a, hx = n23.b, hy = n45.a, hz = n45.b, rf = n67.a ;
#if defined( RTX_EXPERIMENT_EXTRA_LDG ) && defined( __CUDA_ARCH__ )
	// (timing experiment: 2 x RTX_EXPERIMENT_EXTRA_LDG more loads from the node just read -- L1 hits by construction -- on top of
	// the four of a node step: 591.5 -> 641.8 (+2 loads) -> 652.8 ms (+4 loads) per frame)
#pragma unroll
	for ( int e_ = 0 ; e_<RTX_EXPERIMENT_EXTRA_LDG ; e_++ ) {
		float x0, x1, x2, x3 ;   // (a plain load: ptxas folds a second .nc load of the same address into the first)
		asm volatile( "ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( x0 ), "=f"( x1 ), "=f"( x2 ), "=f"( x3 ) : "l"( n+2*( e_&3 ) ) : "memory" ) ;
		if ( x0 == 1.234567e-30f && x3 == 7.654321e-31f ) stack_fault( p.fault ) ;   // (keeps the load alive; never true)
	}
#endif
	int32_t c0 = asint( rf.x ), c1 = asint( rf.y ), c2 = asint( rf.z ), c3 = asint( rf.w ) ;
#if RTX_FFMA2 && defined( __CUDA_ARCH__ )
	float t0, t1, t2, t3 ;
	slab4( lx, ly, lz, hx, hy, hz, idir, ood, tmin, tbest_s, t0, t1, t2, t3 ) ;
#else
	float t0 = slab( lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, idir, ood, tmin, tbest_s ) ;
	float t1 = slab( lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, idir, ood, tmin, tbest_s ) ;
	float t2 = slab( lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, idir, ood, tmin, tbest_s ) ;
	float t3 = slab( lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, idir, ood, tmin, tbest_s ) ;
#endif
	// (unused child slots hold the box lo = hi = +inf, which no ray enters: no test needed)
	// nearest child next, the others pushed far to near (a cheaper "nearest only" ordering was
	// measured: 902 ms instead of 814 ms per frame -- the order of the pushed children matters)
#if defined( RTX_SORT_SELP ) && defined( __CUDA_ARCH__ )
	// (measured alternative: compare-exchange as one comparison and four selects -- 26 instead of 36 instructions for the
	// network, all of them on the ALU pipe: 596.4 against 591.6 ms per frame for the predicated moves the compiler spreads over
	// the ALU and FMA pipes)
#define RTX_CSWAP( ta, ca, tb, cb ) asm( "{\n\t.reg .pred p;\n\t.reg .f32 x;\n\t.reg .b32 y;\n\t" \
	"setp.lt.f32 p, %2, %0;\n\t" \
	"selp.f32 x, %2, %0, p;\n\tselp.f32 %2, %0, %2, p;\n\tmov.f32 %0, x;\n\t" \
	"selp.b32 y, %3, %1, p;\n\tselp.b32 %3, %1, %3, p;\n\tmov.b32 %1, y;\n\t}" : "+f"( ta ), "+r"( ca ), "+f"( tb ), "+r"( cb ) ) ;
#elif ! defined( RTX_SORT_MINMAX )
#define RTX_CSWAP( ta, ca, tb, cb ) if ( tb<ta ) { const float tt = ta ; ta = tb ; tb = tt ; const int32_t cc = ca ; ca = cb ; cb = cc ; }
#else
	// (measured alternative: compare-exchange as min / max on the distances and two selects on the references --
	// 599.9 against 594.4 ms per frame for the predicated swaps)
#define RTX_CSWAP( ta, ca, tb, cb ) { const bool sw_ = tb<ta ; const float lo_ = fminf( ta, tb ), hi_ = fmaxf( ta, tb ) ; const int32_t cl_ = sw_ ? cb : ca, ch_ = sw_ ? ca : cb ; ta = lo_ ; tb = hi_ ; ca = cl_ ; cb = ch_ ; }
#endif
	RTX_CSWAP( t0, c0, t1, c1 ) RTX_CSWAP( t2, c2, t3, c3 ) RTX_CSWAP( t0, c0, t2, c2 ) RTX_CSWAP( t1, c1, t3, c3 ) RTX_CSWAP( t1, c1, t2, c2 )
#undef RTX_CSWAP
	if ( t0 == INFINITY )
		cur = pop_next( p, slot, S, sp, level ) ;
	else {
#if RTX_PREFETCH & 16
		// the nearest child is the next step's node: ask for its line now, ahead of the pushes and the vote (bit 32: all four
		// 32-byte sectors).  Measured: 586.5 (one prefetch) and 651.3 ms (four) against 582.9 ms per frame without
		if ( uint32_t( c0 )<uint32_t( RTX_REF_EMPTY ) ) {
			const q4* nn = ldp<P, q4>( p, F_NODES0, slot )+size_t( c0 )*RTX_NODE_RECS ;
			prefetch_line( nn ) ;
#if RTX_PREFETCH & 32
			prefetch_line( nn+2 ) ; prefetch_line( nn+4 ) ; prefetch_line( nn+6 ) ;
#endif
		}
#endif
		push_far_children( p, slot, sp, c1, t1, c2, t2, c3, t3 ) ;
#if RTX_PREFETCH & 2
		// the second-nearest child, when it is a mesh leaf, is usually next but one: fetch its triangles now
		if ( t1<INFINITY && c1<0 && level>=0 ) prefetch_line( ldp<P, q4>( p, F_TRIS0, slot )+size_t( uint32_t( ~c1 )>>3 )*RTX_TRI_RECS ) ;
#endif
#if RTX_PREFETCH & 8
		// ... every pushed inner node into L2
		if ( t1<INFINITY && c1>=0 ) asm volatile( "prefetch.global.L2 [%0];" :: "l"( ldp<P, q4>( p, F_NODES0, slot )+size_t( c1 )*RTX_NODE_RECS ) ) ;
		if ( t2<INFINITY && c2>=0 ) asm volatile( "prefetch.global.L2 [%0];" :: "l"( ldp<P, q4>( p, F_NODES0, slot )+size_t( c2 )*RTX_NODE_RECS ) ) ;
#endif
#if RTX_PREFETCH & 4
		// ... or an inner node: its line
		if ( t1<INFINITY && c1>=0 ) prefetch_line( ldp<P, q4>( p, F_NODES0, slot )+size_t( c1 )*RTX_NODE_RECS ) ;
#endif
		cur = c0 ;
	}
	return finish_step( p, slot, cur, sp, level ) ;
}

// ---- LEAF: the triangles of a mesh leaf
template <class P> RTX_HD int step_leaf( P& p, int slot, const SceneDev& S ) {
	int32_t cur = p.i( F_CUR, slot ), sp = p.i( F_SP, slot ), level = p.i( F_LEVEL, slot ) ;
	const uint32_t ref = uint32_t( ~cur ) ;
	const uint32_t first = ref>>3, count = ( ref&7u )+1u ;
	const f3 ohi = ld3( p, F_HX, slot ), olo = ld3( p, F_LX, slot ), dd = ld3( p, F_EX, slot ) ;
	HitRec best ;
	best.t = p.f( F_T, slot ) ; best.thing = p.i( F_THING, slot ) ; best.prim = p.i( F_PRIM, slot ) ; best.u = 0.f ; best.v = 0.f ; best.slot = 0 ;
	const q4* tris = ldp<P, q4>( p, F_TRIS0, slot ) ;
	bool changed = false ;
	RTX_COUNT( leaves ) ;
	for ( uint32_t k = 0 ; k<count ; k++ ) {
		RTX_COUNT( tris ) ;
		const q4* T = tris+size_t( first+k )*RTX_TRI_RECS ;
		const o8 t01 = ldo_tri( T ), t23 = ldo_tri( T+2 ) ;
		const q4 a = t01.a, b = t01.b, c = t23.a ;
		float t, u, v ;
		if ( tri_test( mk3( a.x, a.y, a.z ), mk3( b.x, b.y, b.z ), mk3( c.x, c.y, c.z ), ohi, olo, dd, 1e-3f, t, u, v ) ) {
			const int32_t prim = asint( a.w ) ;
			if ( better( t, level, prim, best ) ) {
				best.t = t ; best.thing = level ; best.prim = prim ; best.u = u ; best.v = v ; best.slot = int32_t( first+k ) ;
				changed = true ;
			}
		}
	}
	if ( changed ) {
		p.sf( F_T, slot, best.t ) ; p.si( F_THING, slot, best.thing ) ; p.si( F_PRIM, slot, best.prim ) ;
		p.sf( F_U, slot, best.u ) ; p.sf( F_V, slot, best.v ) ; p.si( F_TRIJ, slot, best.slot ) ;
	}
	cur = pop_next( p, slot, S, sp, level ) ;
	return finish_step( p, slot, cur, sp, level ) ;
}

// ---- THING: a top-level leaf (one thing): analytic sphere, or enter its mesh
template <class P> RTX_HD int step_thing( P& p, int slot, const SceneDev& S ) {
	int32_t cur = p.i( F_CUR, slot ), sp = p.i( F_SP, slot ), level = -1 ;
	const uint32_t first = uint32_t( ~cur )>>3 ;
	const int32_t k = int32_t( RTX_LDG( S.tlas_order+first ) ) ;
	const ThingTrav* tt = S.trav+k ;
	RTX_COUNT( things ) ;
	const f3 o = ld3( p, F_OX, slot ), d = ld3( p, F_DX, slot ) ;
	if ( bsphere_miss( ldq( S.bsphere+k ), o, d, 1e-3f, p.f( F_T, slot ) ) ) {
		RTX_COUNT( spheres ) ;   // (harness: counts culled visits)
		cur = pop_next( p, slot, S, sp, level ) ;
		return finish_step( p, slot, cur, sp, level ) ;
	}
#if RTX_SHADE_INLINE
	const char* ttb = reinterpret_cast<const char*>( tt ) ;
	const o8 A = ldo_rec( ttb ), D = ldo_rec( ttb+96 ) ;   // inv[0..3]; nodes, tris, kind, n_tris, diag, pad
	const double m0 = dbl( A.a.x, A.a.y ), m1 = dbl( A.a.z, A.a.w ), m2 = dbl( A.b.x, A.b.y ), m3 = dbl( A.b.z, A.b.w ) ;
	const int32_t tt_kind = asint( D.b.x ), tt_diag = asint( D.b.z ) ;
#else
	const double m0 = RTX_LDG( tt->inv+0 ), m1 = RTX_LDG( tt->inv+1 ), m2 = RTX_LDG( tt->inv+2 ), m3 = RTX_LDG( tt->inv+3 ) ;
	const int32_t tt_kind = RTX_LDG( &tt->kind ) ;
#endif
	if ( tt_kind == 0 ) {
		double td ;
		if ( sphere_root( mk3( m0, m1, m2 ), m3, wide( o ), wide( d ), double( 1e-3f ), td ) ) {
			const float t = float( td ) ;
			HitRec best ;
			best.t = p.f( F_T, slot ) ; best.thing = p.i( F_THING, slot ) ; best.prim = p.i( F_PRIM, slot ) ;
			if ( better( t, k, -1, best ) ) {
				p.sf( F_T, slot, t ) ; p.si( F_THING, slot, k ) ; p.si( F_PRIM, slot, -1 ) ;
			}
		}
		cur = pop_next( p, slot, S, sp, level ) ;
		return finish_step( p, slot, cur, sp, level ) ;
	}
	// enter the mesh: object-space ray, origin in double carried as hi+lo
	d3 od, ddd ;
#if RTX_SHADE_INLINE
	const o8 B = ldo_rec( ttb+32 ), C = ldo_rec( ttb+64 ) ;
	if ( tt_diag ) {
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
#else
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
#endif
	const f3 ohi = narrow( od ) ;
	const f3 olo = narrow( od-wide( ohi ) ) ;
	const f3 dd  = narrow( ddd ) ;
	const f3 idir = mk3( safe_rcp( dd.x ), safe_rcp( dd.y ), safe_rcp( dd.z ) ) ;
	st3( p, F_HX, slot, ohi ) ; st3( p, F_LX, slot, olo ) ; st3( p, F_EX, slot, dd ) ;
	st3( p, F_IX, slot, idir ) ; st3( p, F_QX, slot, mk3( ohi.x*idir.x, ohi.y*idir.y, ohi.z*idir.z ) ) ;
#if RTX_SHADE_INLINE && defined( __CUDA_ARCH__ )
	p.si( F_NODES0, slot, asint( D.a.x ) ) ; p.si( F_NODES0+1, slot, asint( D.a.y ) ) ; p.si( F_TRIS0, slot, asint( D.a.z ) ) ; p.si( F_TRIS0+1, slot, asint( D.a.w ) ) ;
#else
	stp( p, F_NODES0, slot, ldptr( &tt->nodes ) ) ; stp( p, F_TRIS0, slot, ldptr( &tt->tris ) ) ;
#endif
	p.si( F_LEVEL, slot, k ) ;
	p.push( slot, sp, RTX_STK_RETURN, 0.f ) ;
	return finish_step( p, slot, 0, sp, k ) ;
}

// ---- SHADE: the ray is finished.  Returns K_NODE/K_SHADE (path continues with a new ray)
// or K_REGEN (path ended; `c` is its colour).  rtow.cxx:34-49.
// Guide layers: the first diffuse or reflecting hit of a path contributes its normal and
// albedo (optx/optics_i.cu:97-101, 185-189; refracting hits do not); bit 8 of F_META marks
// a path that has contributed.  `guide` is set when this call captured them.
template <class P> RTX_HD int step_shade( P& p, int slot, const SceneDev& S, f3& c, bool& guide, f3& gnormal, f3& galbedo, uint32_t& segments ) {
	guide = false ;
	HitRec h ;
	RTX_COUNT( rays ) ;
	h.t = p.f( F_T, slot ) ; h.thing = p.i( F_THING, slot ) ; h.prim = p.i( F_PRIM, slot ) ; h.u = p.f( F_U, slot ) ; h.v = p.f( F_V, slot ) ; h.slot = p.i( F_TRIJ, slot ) ;
	const f3 ori = ld3( p, F_OX, slot ), dir = ld3( p, F_DX, slot ) ;
	f3 thr = ld3( p, F_THRX, slot ) ;
	const uint32_t meta = uint32_t( p.i( F_META, slot ) )+512u ;   // bits 0-7 depth left, 8 guide taken, 9-15 segments of the path
	const uint32_t depth_left = meta&255u ;
	segments = meta>>9 ;
	RTX_COUNT_LIVE( segments-1u ) ;
	c = mk3( 0.f, 0.f, 0.f ) ;
	if ( h.thing<0 ) {
		c = thr*sky( dir ) ;
		return K_REGEN ;
	}
	if ( depth_left == 0 && S.variant != RTX_SEM_RTWO_I )
		return K_REGEN ;
	Frame fr ;
	Pcg rng ;
	rng.state = uint64_t( uint32_t( p.i( F_RNG0, slot ) ) )|( uint64_t( uint32_t( p.i( F_RNG1, slot ) ) )<<32 ) ;
	f3 att, out ;
#if RTX_SHADE_INLINE
	o8 mat ;
	qframe_of( S, h.thing, h.prim<0 ? -1 : h.slot, h.u, h.v, ori, dir, 1e-3f, fr, mat ) ;
	const bool go = qscatter( mat, dir, fr, rng, att, out, S.variant == RTX_SEM_RTOW ) ;
	const int32_t h_type = asint( mat.b.y ) ;
#else
	frame_of( S, h, ori, dir, 1e-3f, fr ) ;
	const bool go = scatter( S.shade+h.thing, dir, fr, rng, att, out, S.variant == RTX_SEM_RTOW ) ;
	const int32_t h_type = RTX_LDG( &( S.shade+h.thing )->type ) ;
#endif
	uint32_t meta2 = meta ;
	if ( ! ( meta&256u ) && h_type != 2 ) {
		guide = true ; gnormal = fr.normal ; galbedo = att ;   // att = the thing's albedo for diffuse / reflect
		meta2 |= 256u ;
	}
	if ( ! go )
		return K_REGEN ;
	thr = thr*att ;
	if ( depth_left == 0 ) {   // RTX_SEM_RTWO_I: the last ray scattered, its throughput is the colour (optx/camera_i.cu:92-95)
		c = thr ;
		return K_REGEN ;
	}
	st3( p, F_THRX, slot, thr ) ;
	p.si( F_RNG0, slot, int32_t( uint32_t( rng.state ) ) ) ; p.si( F_RNG1, slot, int32_t( uint32_t( rng.state>>32 ) ) ) ;
	p.si( F_META, slot, int32_t( meta2-1u ) ) ;
	begin_ray( p, slot, S, fr.p, out ) ;
	return kind_of( p.i( F_CUR, slot ), -1 ) ;
}

// ---- REGEN: a new path (pixel x,y of the image, tile-local pixel slot, global sample index)
template <class P> RTX_HD int step_regen( P& p, int slot, const SceneDev& S, const CameraDev& cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h,
		uint32_t tile_pixel, uint64_t seed, uint32_t sample, uint32_t depth ) {
	Pcg rng ;
	rng.seed( seed, w*y+x, sample ) ;
	f3 ori, dir ;
	primary_ray( cam, x, y, w, h, rng, ori, dir, S.variant != RTX_SEM_RTOW ) ;
	depth = sem_depth( S.variant, depth ) ;
	st3( p, F_THRX, slot, mk3( 1.f, 1.f, 1.f ) ) ;
	p.si( F_RNG0, slot, int32_t( uint32_t( rng.state ) ) ) ; p.si( F_RNG1, slot, int32_t( uint32_t( rng.state>>32 ) ) ) ;
	p.si( F_PIX, slot, int32_t( tile_pixel ) ) ;
	p.si( F_META, slot, int32_t( depth>255u ? 255u : depth ) ) ;
	begin_ray( p, slot, S, ori, dir ) ;
	return kind_of( p.i( F_CUR, slot ), -1 ) ;
}

} // namespace rtx
