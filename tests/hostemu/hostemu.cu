// hostemu.cu -- DEVELOPMENT HARNESS (test infrastructure, not product): drives the
// __host__ __device__ core of the CUDA path tracer (rtx_core.cuh: traversal, primitive
// tests, shading, random stream) and the LBVH helper functions (rtx_lbvh.cuh: Morton
// keys, Karras node, box padding) serially on the CPU, so that their logic and their
// arithmetic contract can be checked against the oracle in a container without a GPU.
// Never linked into librtx.so; the product has no CPU path.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#define RTX_STATS 1
#include "../../rtxplay_b200/csrc/rtx_core.cuh"
#include "../../rtxplay_b200/csrc/rtx_lbvh.cuh"
#include "../../rtxplay_b200/csrc/rtx_pool.cuh"
#include "../../rtxplay_b200/csrc/rtx_qpool.cuh"
#include "../../rtxplay_b200/csrc/rtx_hostmath.h"

using namespace rtx ;
#include <string>
namespace rtx {
Stats g_stats = { 0, 0, 0, 0, 0, 0, 0, 0, 0 } ;
std::string* g_trace = nullptr ;
void trace_event( char c ) { if ( g_trace ) g_trace->push_back( c ) ; }
}

namespace {

uint32_t g_variant = 0 ;   // RTX_SEM_* of the following calls (emu_set_variant)

struct HostStack {
	int32_t v[256] ; int sp ;
	RTX_HD void reset() { sp = 0 ; }
	RTX_HD void push( int32_t x ) { v[sp++] = x ; RTX_COUNT( pushes ) ; RTX_COUNT_MAX( maxsp, sp ) ; }
	RTX_HD int32_t pop() { return v[--sp] ; }
} ;

// one ray slot, array backed: the store the step functions of rtx_pool.cuh run on here
unsigned long long g_sp_hist[128] ;   // stack size seen by pushes (development statistic)
struct HostPool {
	uint32_t w[F_WORDS] ; int32_t ovf[512] ;
	RTX_HD float   f( int fld, int ) const { float v ; memcpy( &v, &w[fld], 4 ) ; return v ; }
	RTX_HD int32_t i( int fld, int ) const { return int32_t( w[fld] ) ; }
	RTX_HD void    sf( int fld, int, float v ) { memcpy( &w[fld], &v, 4 ) ; }
	RTX_HD void    si( int fld, int, int32_t v ) { w[fld] = uint32_t( v ) ; }
	RTX_HD void    push( int, int32_t& sp, int32_t v, float t ) {
		uint32_t tb ; memcpy( &tb, &t, 4 ) ;
#if ! defined( __CUDA_ARCH__ )
		g_sp_hist[sp<127 ? sp : 127]++ ;
#endif
		if ( sp<RTX_POOL_STACK ) { w[F_STACK+2*sp] = uint32_t( v ) ; w[F_STACK+2*sp+1] = tb ; }
		else { ovf[2*( sp-RTX_POOL_STACK )] = v ; ovf[2*( sp-RTX_POOL_STACK )+1] = int32_t( tb ) ; }
		sp++ ;
	}
	RTX_HD int32_t pop( int, int32_t& sp, float& t ) {
		sp-- ;
		uint32_t tb ; int32_t v ;
		if ( sp<RTX_POOL_STACK ) { v = int32_t( w[F_STACK+2*sp] ) ; tb = w[F_STACK+2*sp+1] ; }
		else { v = ovf[2*( sp-RTX_POOL_STACK )] ; tb = uint32_t( ovf[2*( sp-RTX_POOL_STACK )+1] ) ; }
		memcpy( &t, &tb, 4 ) ;
		return v ;
	}
} ;

// a whole path through the state machine of the render kernel
f3 pool_path( const SceneDev& S, const CameraDev& cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint64_t seed, uint32_t sample, uint32_t depth, uint32_t& segments ) {
	static HostPool P ;
	int kind = step_regen( P, 0, S, cam, x, y, w, h, 0u, seed, sample, depth ) ;
	f3 c = mk3( 0.f, 0.f, 0.f ) ;
	while ( true ) {
		switch ( kind ) {
			case K_NODE:  kind = step_node( P, 0, S ) ; break ;
			case K_LEAF:  kind = step_leaf( P, 0, S ) ; break ;
			case K_THING: kind = step_thing( P, 0, S ) ; break ;
			case K_SHADE: { segments++ ; bool g ; f3 gn, ga ; uint32_t sg ; kind = step_shade( P, 0, S, c, g, gn, ga, sg ) ; break ; }
			default: return c ;
		}
	}
}


// one slot of the compacting pool (rtx_qpool.cuh), array backed; node / triangle arrays are
// "encoded" as indices into a small pointer table (the device uses 16-byte offsets from an arena)
struct QHost {
	q4 rec[RTX_QS] ; q4 cold[4] ; int32_t ovf[2*512] ;
	std::vector<const q4*> ptrs ;
	RTX_HD q4   ldq( int, int q ) const { return rec[q] ; }
	RTX_HD void stq( int, int q, const q4& v ) { rec[q] = v ; }
	RTX_HD void stw( int, int q, int w, float v ) { ( &rec[q].x )[w] = v ; }
	RTX_HD q4   ldc( int, int q ) const { return cold[q] ; }
	RTX_HD void stc( int, int q, const q4& v ) { cold[q] = v ; }
	RTX_HD void stcw( int, int q, int w, float v ) { ( &cold[q].x )[w] = v ; }
	RTX_HD void push( int, int32_t& sp, int32_t v, float t ) {
#if ! defined( __CUDA_ARCH__ )
		g_sp_hist[sp<127 ? sp : 127]++ ;
#endif
		if ( sp<RTX_QSTACK ) { float* e = &rec[6].x+2*sp ; e[0] = asfloat( v ) ; e[1] = t ; }
		else { ovf[2*( sp-RTX_QSTACK )] = v ; ovf[2*( sp-RTX_QSTACK )+1] = asint( t ) ; }
		sp++ ;
	}
	RTX_HD int32_t pop( int, int32_t& sp, float& t ) {
		sp-- ;
		if ( sp<RTX_QSTACK ) { const float* e = &rec[6].x+2*sp ; t = e[1] ; return asint( e[0] ) ; }
		t = asfloat( ovf[2*( sp-RTX_QSTACK )+1] ) ;
		return ovf[2*( sp-RTX_QSTACK )] ;
	}
	RTX_HD float enc( int, const q4* p ) {
#if ! defined( __CUDA_ARCH__ )
		if ( ! p ) return asfloat( 0 ) ;
		for ( size_t k = 0 ; k<ptrs.size() ; k++ ) if ( ptrs[k] == p ) return asfloat( int32_t( k+1 ) ) ;
		ptrs.push_back( p ) ;
		return asfloat( int32_t( ptrs.size() ) ) ;
#else
		return 0.f ;
#endif
	}
	RTX_HD const q4* dec( int, float w ) const {
#if ! defined( __CUDA_ARCH__ )
		return ptrs[size_t( asint( w )-1 )] ;
#else
		return nullptr ;
#endif
	}
} ;

// a whole path through the step functions of the compacting pool
f3 qpool_path( const SceneDev& S, const CameraDev& cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint64_t seed, uint32_t sample, uint32_t depth, uint32_t& segments ) {
	static QHost P ;
	int kind = qstep_regen( P, 0, S, cam, x, y, w, h, 0u, seed, sample, depth ) ;
	f3 c = mk3( 0.f, 0.f, 0.f ) ;
	while ( true ) {
		switch ( kind ) {
			case K_NODE:  kind = qstep_node( P, 0, S ) ; break ;
			case K_LEAF:  kind = qstep_leaf( P, 0, S ) ; break ;
			case K_THING: kind = qstep_thing( P, 0, S ) ; break ;
			case K_SHADE: { segments++ ; bool g ; f3 gn, ga ; uint32_t sg, pix ; kind = qstep_shade( P, 0, S, c, pix, g, gn, ga, sg ) ; break ; }
			default: return c ;
		}
	}
}

struct Tree {
	std::vector<q4>       nodes ;
	std::vector<uint32_t> order ;
	q4 root_lo, root_hi ;
} ;

// the same pipeline as lbvh_build(), serial: keys -> stable sort -> karras -> refit -> collapse
struct I2 { int x, y ; } ;
void build_tree( const std::vector<q4>& plo, const std::vector<q4>& phi, Tree& T, int leaf_max ) {
	const int n = int( plo.size() ) ;
	float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY } ;
	for ( int i = 0 ; i<n ; i++ ) {
		const float c[3] = { .5f*( plo[i].x+phi[i].x ), .5f*( plo[i].y+phi[i].y ), .5f*( plo[i].z+phi[i].z ) } ;
		for ( int a = 0 ; a<3 ; a++ ) { mn[a] = fminf( mn[a], c[a] ) ; mx[a] = fmaxf( mx[a], c[a] ) ; }
	}
	std::vector<uint64_t> keys( n ) ;
	T.order.resize( n ) ;
	for ( int i = 0 ; i<n ; i++ ) {
		const float ex = mx[0]-mn[0], ey = mx[1]-mn[1], ez = mx[2]-mn[2] ;
		const float cx = .5f*( plo[i].x+phi[i].x ), cy = .5f*( plo[i].y+phi[i].y ), cz = .5f*( plo[i].z+phi[i].z ) ;
		keys[i] = morton63( ex>0.f ? ( cx-mn[0] )/ex : 0.f, ey>0.f ? ( cy-mn[1] )/ey : 0.f, ez>0.f ? ( cz-mn[2] )/ez : 0.f ) ;
		T.order[i] = uint32_t( i ) ;
	}
	std::stable_sort( T.order.begin(), T.order.end(), [&]( uint32_t a, uint32_t b ) { return keys[a]<keys[b] ; } ) ;
	std::vector<uint64_t> sk( n ) ;
	for ( int i = 0 ; i<n ; i++ ) sk[i] = keys[T.order[i]] ;

	std::vector<q4> blo( 2*size_t( n ) ), bhi( 2*size_t( n ) ) ;
	for ( int j = 0 ; j<n ; j++ ) {
		const uint32_t p = T.order[j] ;
		f3 lo = mk3( plo[p].x, plo[p].y, plo[p].z ), hi = mk3( phi[p].x, phi[p].y, phi[p].z ) ;
		pad_box( lo, hi ) ;
		blo[n-1+j] = { lo.x, lo.y, lo.z, 0.f } ; bhi[n-1+j] = { hi.x, hi.y, hi.z, 0.f } ;
	}
	std::vector<I2> child( n>1 ? n-1 : 1 ), range( n>1 ? n-1 : 1 ) ;
	for ( int i = 0 ; i<n-1 ; i++ ) {
		int l, r, lo, hi ; bool ll, rl ;
		karras_node( sk.data(), n, i, l, r, ll, rl, lo, hi ) ;
		child[i].x = ll ? ~l : l ; child[i].y = rl ? ~r : r ;
		range[i].x = lo ; range[i].y = hi ;
	}
	if ( n>1 ) {
		// post-order refit without recursion
		std::vector<int> todo ; std::vector<char> seen( n-1, 0 ) ;
		todo.push_back( 0 ) ;
		while ( ! todo.empty() ) {
			const int i = todo.back() ;
			if ( ! seen[i] ) {
				seen[i] = 1 ;
				if ( child[i].x>=0 ) todo.push_back( child[i].x ) ;
				if ( child[i].y>=0 ) todo.push_back( child[i].y ) ;
				continue ;
			}
			todo.pop_back() ;
			const int a = child[i].x<0 ? n-1+( ~child[i].x ) : child[i].x, b = child[i].y<0 ? n-1+( ~child[i].y ) : child[i].y ;
			blo[i] = { fminf( blo[a].x, blo[b].x ), fminf( blo[a].y, blo[b].y ), fminf( blo[a].z, blo[b].z ), 0.f } ;
			bhi[i] = { fmaxf( bhi[a].x, bhi[b].x ), fmaxf( bhi[a].y, bhi[b].y ), fmaxf( bhi[a].z, bhi[b].z ), 0.f } ;
		}
	}
	// collapse, breadth first (k_wide_level)
	std::vector<I2> front( 1, I2{ 0, 0 } ) ;
	int n_nodes = 1 ;
	T.nodes.assign( RTX_NODE_RECS, q4{ 0, 0, 0, 0 } ) ;
	while ( ! front.empty() ) {
		std::vector<I2> next ;
		for ( const I2& item : front ) {
			float lo[3][RTX_WIDTH], hi[3][RTX_WIDTH] ; int ref[RTX_WIDTH] ;
			for ( int k = 0 ; k<RTX_WIDTH ; k++ ) { ref[k] = RTX_REF_EMPTY ; for ( int a = 0 ; a<3 ; a++ ) { lo[a][k] = INFINITY ; hi[a][k] = INFINITY ; } }
			if ( n == 1 ) {
				ref[0] = ~0 ;
				lo[0][0] = blo[0].x ; lo[1][0] = blo[0].y ; lo[2][0] = blo[0].z ; hi[0][0] = bhi[0].x ; hi[1][0] = bhi[0].y ; hi[2][0] = bhi[0].z ;
			} else {
				int slots[RTX_WIDTH] ;
				const int ns = wide_gather( item.x, child.data(), range.data(), blo.data(), bhi.data(), n, leaf_max, slots ) ;
				for ( int k = 0 ; k<ns ; k++ ) {
					const int s = slots[k] ;
					const int b = s<0 ? n-1+( ~s ) : s ;
					lo[0][k] = blo[b].x ; lo[1][k] = blo[b].y ; lo[2][k] = blo[b].z ; hi[0][k] = bhi[b].x ; hi[1][k] = bhi[b].y ; hi[2][k] = bhi[b].z ;
					if ( ! wide_leaf_ref( s, range.data(), leaf_max, ref[k] ) ) {
						ref[k] = n_nodes++ ;
						next.push_back( I2{ s, ref[k] } ) ;
					}
				}
			}
			if ( T.nodes.size()<size_t( n_nodes )*RTX_NODE_RECS ) T.nodes.resize( size_t( n_nodes )*RTX_NODE_RECS, q4{ 0, 0, 0, 0 } ) ;
			for ( int hh = 0 ; hh<RTX_WIDTH/4 ; hh++ ) {
				q4* o = T.nodes.data()+size_t( item.y )*RTX_NODE_RECS+8*hh ;
				for ( int a = 0 ; a<3 ; a++ ) {
					for ( int k = 4*hh ; k<4*hh+4 ; k++ ) box_ch( lo[a][k], hi[a][k], lo[a][k], hi[a][k] ) ;
					o[a]   = { lo[a][4*hh], lo[a][4*hh+1], lo[a][4*hh+2], lo[a][4*hh+3] } ;
					o[3+a] = { hi[a][4*hh], hi[a][4*hh+1], hi[a][4*hh+2], hi[a][4*hh+3] } ;
				}
				o[6] = { asfloat( ref[4*hh] ), asfloat( ref[4*hh+1] ), asfloat( ref[4*hh+2] ), asfloat( ref[4*hh+3] ) } ;
			}
		}
		front.swap( next ) ;
	}
	T.root_lo = blo[0] ; T.root_hi = bhi[0] ;
}

struct EmuMesh { std::vector<float> vces ; std::vector<uint32_t> ices ; std::vector<q4> tris ; Tree tree ; double bsphere[4] ; } ;

struct EmuScene {
	std::vector<EmuMesh>    meshes ;
	std::vector<ThingTrav>  trav ;
	std::vector<ThingShade> shade ;
	std::vector<q4>         bsphere ;
	Tree                    tlas ;
	SceneDev                S ;
} ;

enum { TH_KIND = 0, TH_MESH = 1, TH_XF = 2, TH_TYPE = 14, TH_ALB = 15, TH_FUZZ = 18, TH_INDEX = 19, TH_STRIDE = 20 } ;

void build_scene( EmuScene& E, const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt ) {
	E.meshes.resize( n_meshes ) ;
	for ( int q = 0 ; q<n_meshes ; q++ ) {
		EmuMesh& m = E.meshes[q] ;
		m.vces.assign( vces[q], vces[q]+3*size_t( nv[q] ) ) ;
		m.ices.assign( ices[q], ices[q]+3*size_t( nt[q] ) ) ;
		mesh_bsphere( vces[q], nv[q], m.bsphere ) ;
		std::vector<q4> plo( nt[q] ), phi( nt[q] ) ;
		for ( uint32_t f = 0 ; f<nt[q] ; f++ ) {
			const float* a = &m.vces[3*size_t( m.ices[3*f] )] ; const float* b = &m.vces[3*size_t( m.ices[3*f+1] )] ; const float* c = &m.vces[3*size_t( m.ices[3*f+2] )] ;
			plo[f] = { fminf( a[0], fminf( b[0], c[0] ) ), fminf( a[1], fminf( b[1], c[1] ) ), fminf( a[2], fminf( b[2], c[2] ) ), 0.f } ;
			phi[f] = { fmaxf( a[0], fmaxf( b[0], c[0] ) ), fmaxf( a[1], fmaxf( b[1], c[1] ) ), fmaxf( a[2], fmaxf( b[2], c[2] ) ), 0.f } ;
		}
		build_tree( plo, phi, m.tree, RTX_LEAF_MAX ) ;
		m.tris.assign( RTX_TRI_RECS*size_t( nt[q] ), q4{ 0, 0, 0, 0 } ) ;
		for ( uint32_t j = 0 ; j<nt[q] ; j++ ) {
			const uint32_t f = m.tree.order[j] ;
			const float* a = &m.vces[3*size_t( m.ices[3*f] )] ; const float* b = &m.vces[3*size_t( m.ices[3*f+1] )] ; const float* c = &m.vces[3*size_t( m.ices[3*f+2] )] ;
			m.tris[RTX_TRI_RECS*size_t( j )]   = { a[0], a[1], a[2], asfloat( int( f ) ) } ;
			m.tris[RTX_TRI_RECS*size_t( j )+1] = { b[0]-a[0], b[1]-a[1], b[2]-a[2], b[0] } ;
			m.tris[RTX_TRI_RECS*size_t( j )+2] = { c[0]-a[0], c[1]-a[1], c[2]-a[2], b[1] } ;
			m.tris[RTX_TRI_RECS*size_t( j )+3] = { b[2], c[0], c[1], c[2] } ;
		}
	}
	E.trav.resize( n_things ) ; E.shade.resize( n_things ) ; E.bsphere.resize( n_things ) ;
	std::vector<q4> plo( n_things ), phi( n_things ) ;
	for ( int k = 0 ; k<n_things ; k++ ) {
		const double* row = things+size_t( k )*TH_STRIDE ;
		ThingTrav& t = E.trav[k] ; ThingShade& s = E.shade[k] ;
		memset( &t, 0, sizeof( t ) ) ; memset( &s, 0, sizeof( s ) ) ;
		float xf[12] ;
		for ( int j = 0 ; j<12 ; j++ ) { xf[j] = float( row[TH_XF+j] ) ; s.xf[j] = double( xf[j] ) ; }
		s.albedo[0] = float( row[TH_ALB] ) ; s.albedo[1] = float( row[TH_ALB+1] ) ; s.albedo[2] = float( row[TH_ALB+2] ) ;
		s.fuzz = float( row[TH_FUZZ] ) ; s.index = float( row[TH_INDEX] ) ; s.type = int( row[TH_TYPE] ) ;
		if ( int( row[TH_KIND] ) == 0 ) {
			t.kind = 0 ; s.kind = 0 ;
			{
				const double unit[4] = { 0., 0., 0., 1. } ;
				const float sx[12] = { std::fabs( xf[0] ), 0, 0, xf[3], 0, std::fabs( xf[0] ), 0, xf[7], 0, 0, std::fabs( xf[0] ), xf[11] } ;
				world_bsphere( sx, unit, &E.bsphere[k].x ) ;
			}
			t.inv[0] = double( xf[3] ) ; t.inv[1] = double( xf[7] ) ; t.inv[2] = double( xf[11] ) ; t.inv[3] = double( xf[0] ) ;
			const double r = fabs( t.inv[3] ) ;
			plo[k] = { float( t.inv[0]-r )-1e-3f, float( t.inv[1]-r )-1e-3f, float( t.inv[2]-r )-1e-3f, 0.f } ;
			phi[k] = { float( t.inv[0]+r )+1e-3f, float( t.inv[1]+r )+1e-3f, float( t.inv[2]+r )+1e-3f, 0.f } ;
		} else {
			const EmuMesh& m = E.meshes[int( row[TH_MESH] )] ;
			t.kind = 1 ; s.kind = 1 ;
			t.diag = s.diag = ( xf[1] == 0.f && xf[2] == 0.f && xf[4] == 0.f && xf[6] == 0.f && xf[8] == 0.f && xf[9] == 0.f ) ? 1 : 0 ;
			world_bsphere( xf, m.bsphere, &E.bsphere[k].x ) ;
			affine_inverse( xf, t.inv ) ;
			t.nodes = m.tree.nodes.data() ; t.tris = m.tris.data() ; t.n_tris = uint32_t( m.ices.size()/3 ) ;
			s.vces = m.vces.data() ; s.ices = m.ices.data() ;
			double mn[3] = { 1e300, 1e300, 1e300 }, mx[3] = { -1e300, -1e300, -1e300 } ;
			for ( int c = 0 ; c<8 ; c++ ) {
				const d3 p = mk3( double( c&1 ? m.tree.root_hi.x : m.tree.root_lo.x ), double( c&2 ? m.tree.root_hi.y : m.tree.root_lo.y ), double( c&4 ? m.tree.root_hi.z : m.tree.root_lo.z ) ) ;
				const d3 w = xfpoint( s.xf, p ) ;
				mn[0] = fmin( mn[0], w.x ) ; mn[1] = fmin( mn[1], w.y ) ; mn[2] = fmin( mn[2], w.z ) ;
				mx[0] = fmax( mx[0], w.x ) ; mx[1] = fmax( mx[1], w.y ) ; mx[2] = fmax( mx[2], w.z ) ;
			}
			plo[k] = { float( mn[0] )-1e-3f, float( mn[1] )-1e-3f, float( mn[2] )-1e-3f, 0.f } ;
			phi[k] = { float( mx[0] )+1e-3f, float( mx[1] )+1e-3f, float( mx[2] )+1e-3f, 0.f } ;
		}
	}
	if ( n_things ) build_tree( plo, phi, E.tlas, 1 ) ;
	E.S.tlas_nodes = E.tlas.nodes.data() ; E.S.tlas_order = E.tlas.order.data() ;
	E.S.trav = E.trav.data() ; E.S.shade = E.shade.data() ; E.S.bsphere = E.bsphere.data() ; E.S.n_things = uint32_t( n_things ) ; E.S.variant = g_variant ; E.S.fault = nullptr ; E.S.arena = nullptr ;
}

} // namespace

extern "C" {

void emu_set_variant( int v ) { g_variant = uint32_t( v ) ; }

// cam: 19 doubles (eye,u,v,hvec,wvec,dvec,aperture) as in the oracle tables
int emu_render( const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int spp, int depth, uint64_t seed, int sample0, int sample_stride,
		uint64_t* fix, uint32_t* rpp, int64_t* first_id, float* first_t, int brute, int use_pool ) {
	EmuScene E ;
	build_scene( E, things, n_things, n_meshes, vces, nv, ices, nt ) ;
	CameraDev c ;
	c.eye = mk3( float( cam[0] ), float( cam[1] ), float( cam[2] ) ) ; c.u = mk3( float( cam[3] ), float( cam[4] ), float( cam[5] ) ) ;
	c.v = mk3( float( cam[6] ), float( cam[7] ), float( cam[8] ) ) ; c.hvec = mk3( float( cam[9] ), float( cam[10] ), float( cam[11] ) ) ;
	c.wvec = mk3( float( cam[12] ), float( cam[13] ), float( cam[14] ) ) ; c.dvec = mk3( float( cam[15] ), float( cam[16] ), float( cam[17] ) ) ;
	c.aperture = float( cam[18] ) ;
	HostStack st ;
	for ( int y = 0 ; y<h ; y++ )
		for ( int x = 0 ; x<w ; x++ ) {
			const uint32_t pix = uint32_t( w )*y+x ;
			uint64_t acc[3] = { 0, 0, 0 } ; uint32_t segments = 0 ;
			for ( int k = 0 ; k<spp ; k++ ) {
				Pcg rng ;
				rng.seed( seed, pix, uint32_t( sample0+k*sample_stride ) ) ;
				f3 ori, dir ;
				primary_ray( c, uint32_t( x ), uint32_t( y ), uint32_t( w ), uint32_t( h ), rng, ori, dir, g_variant != RTX_SEM_RTOW ) ;
				if ( k == 0 && first_id ) {
					HitRec hr ;
					if ( brute ) closest_brute( E.S, ori, dir, 1e-3f, hr ) ;
					else         closest( E.S, ori, dir, 1e-3f, st, hr ) ;
					first_id[pix] = hr.thing<0 ? int64_t( -1 ) : ( ( int64_t( hr.thing )<<32 )|int64_t( uint32_t( hr.prim+1 ) ) ) ;
					if ( first_t ) first_t[pix] = hr.thing<0 ? -1.f : hr.t ;
				}
				const f3 col = use_pool == 2 ? qpool_path( E.S, c, uint32_t( x ), uint32_t( y ), uint32_t( w ), uint32_t( h ), seed, uint32_t( sample0+k*sample_stride ), uint32_t( depth ), segments )
				                : use_pool ? pool_path( E.S, c, uint32_t( x ), uint32_t( y ), uint32_t( w ), uint32_t( h ), seed, uint32_t( sample0+k*sample_stride ), uint32_t( depth ), segments )
				                        : path_radiance( E.S, ori, dir, uint32_t( depth ), rng, st, segments ) ;
				acc[0] += tofix( col.x ) ; acc[1] += tofix( col.y ) ; acc[2] += tofix( col.z ) ;
			}
			if ( fix ) { fix[3*size_t( pix )] = acc[0] ; fix[3*size_t( pix )+1] = acc[1] ; fix[3*size_t( pix )+2] = acc[2] ; }
			if ( rpp ) rpp[pix] = segments ;
		}
	return 0 ;
}

// event trace of every path of an image: per pixel (row-major) and sample, the traversal
// events of its rays ('N' node, '1'..'4' leaf with that many triangles, 'P' sphere test,
// 'E' enter mesh, 'R' return) with '|' after each ray and ';' after each path
long long emu_trace( const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int spp, int depth, uint64_t seed, char* out, long long cap ) {
	EmuScene E ;
	build_scene( E, things, n_things, n_meshes, vces, nv, ices, nt ) ;
	CameraDev c ;
	c.eye = mk3( float( cam[0] ), float( cam[1] ), float( cam[2] ) ) ; c.u = mk3( float( cam[3] ), float( cam[4] ), float( cam[5] ) ) ;
	c.v = mk3( float( cam[6] ), float( cam[7] ), float( cam[8] ) ) ; c.hvec = mk3( float( cam[9] ), float( cam[10] ), float( cam[11] ) ) ;
	c.wvec = mk3( float( cam[12] ), float( cam[13] ), float( cam[14] ) ) ; c.dvec = mk3( float( cam[15] ), float( cam[16] ), float( cam[17] ) ) ;
	c.aperture = float( cam[18] ) ;
	std::string tr ;
	g_trace = &tr ;
	HostStack st ;
	for ( int y = 0 ; y<h ; y++ )
		for ( int x = 0 ; x<w ; x++ ) {
			const uint32_t pix = uint32_t( w )*y+x ;
			for ( int k = 0 ; k<spp ; k++ ) {
				Pcg rng ;
				rng.seed( seed, pix, uint32_t( k ) ) ;
				f3 ori, dir ;
				primary_ray( c, uint32_t( x ), uint32_t( y ), uint32_t( w ), uint32_t( h ), rng, ori, dir, g_variant != RTX_SEM_RTOW ) ;
				// path_radiance, with a ray separator
				f3 thr = mk3( 1.f, 1.f, 1.f ) ; uint32_t dl = uint32_t( depth ) ;
				while ( true ) {
					HitRec hr ;
					closest( E.S, ori, dir, 1e-3f, st, hr ) ;
					tr.push_back( '|' ) ;
					if ( hr.thing<0 || dl == 0 ) break ;
					Frame fr ; frame_of( E.S, hr, ori, dir, 1e-3f, fr ) ;
					f3 att, out2 ;
					if ( ! scatter( E.S.shade+hr.thing, dir, fr, rng, att, out2 ) ) break ;
					thr = thr*att ; ori = fr.p ; dir = out2 ; dl-- ;
				}
				tr.push_back( ';' ) ;
			}
		}
	g_trace = nullptr ;
	const long long n = ( long long ) tr.size() ;
	if ( out && n<=cap ) memcpy( out, tr.data(), size_t( n ) ) ;
	return n ;
}

// ---- warp scheduling simulator (development instrument): 32 lanes run the step functions in
// lockstep under the vote policy of k_render (policy 0) or the parked-shading policy
// (policy 1, rtx_kernels.cuh RTX_PARK), counting iterations, lane-steps and a rough
// instruction cost per step kind.  Results are schedule independent, so only the counts matter.
//   out[0..7]   iterations per kind (1 node, 2 leaf, 3 thing, 4 shade, 5 regen, 6 swap, 7 batch)
//   out[8..15]  lane-steps per kind
//   out[16]     rays, out[17] estimated warp instructions, out[18] paths
namespace {
struct SimLane { HostPool act, park ; int kind ; int pstate ; } ;   // pstate: 0 empty, 1 finished, 2 ready
enum { S_SWAP = 6, S_BATCH = 7, S_WAIT = 8, S_NONE = 9 } ;
struct SimCost { double vote, node, leaf0, leaf1, thing0, thing1, sky, shade, regen, swap ; } ;
}
long long emu_warpsim( const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int unit_spp, int units_per_warp, int tile_step, int depth, uint64_t seed,
		int policy, int sticky, int t1, int t2, const double* costs, double* out ) {
	EmuScene E ;
	build_scene( E, things, n_things, n_meshes, vces, nv, ices, nt ) ;
	CameraDev c ;
	c.eye = mk3( float( cam[0] ), float( cam[1] ), float( cam[2] ) ) ; c.u = mk3( float( cam[3] ), float( cam[4] ), float( cam[5] ) ) ;
	c.v = mk3( float( cam[6] ), float( cam[7] ), float( cam[8] ) ) ; c.hvec = mk3( float( cam[9] ), float( cam[10] ), float( cam[11] ) ) ;
	c.wvec = mk3( float( cam[12] ), float( cam[13] ), float( cam[14] ) ) ; c.dvec = mk3( float( cam[15] ), float( cam[16] ), float( cam[17] ) ) ;
	c.aperture = float( cam[18] ) ;
	SimCost K ; memcpy( &K, costs, sizeof( K ) ) ;
	for ( int k = 0 ; k<24 ; k++ ) out[k] = 0. ;
	const int tiles_x = ( w+7 )/8, tiles_y = ( h+3 )/4, n_tiles = tiles_x*tiles_y ;
	std::vector<SimLane> L( 32 ) ;
	double cost = 0. ;
	for ( int tile0 = 0 ; tile0<n_tiles ; tile0 += tile_step ) {
		// this warp's stream of units: unit q = chunk q/4 of tile tile0 + q%4 (a 4-tile strip, like an SM's shared block)
		int unit_next = 0 ; uint32_t unit_pos = 0, unit_left = 0, ux0 = 0, uy0 = 0, us0 = 0 ; bool exhausted = false ;
		auto serve = [&]( HostPool& P, int& kind_out ) -> bool {   // next path of the stream into P; false: none left
			while ( true ) {
				if ( unit_left == 0 ) {
					if ( unit_next>=units_per_warp ) { exhausted = true ; return false ; }
					const int tile = ( tile0+unit_next%4 )%n_tiles, chunk = unit_next/4 ;
					unit_next++ ;
					ux0 = uint32_t( tile%tiles_x )*8u ; uy0 = uint32_t( tile/tiles_x )*4u ; us0 = uint32_t( chunk*unit_spp ) ;
					unit_pos = 0 ; unit_left = 32u*uint32_t( unit_spp ) ;
				}
				const uint32_t idx = unit_pos++ ; unit_left-- ;
				const uint32_t px = idx&31u, smp = us0+( idx>>5 ) ;
				const uint32_t x = ux0+( px&7u ), y = uy0+( px>>3 ) ;
				if ( x<uint32_t( w ) && y<uint32_t( h ) ) {
					kind_out = step_regen( P, 0, E.S, c, x, y, uint32_t( w ), uint32_t( h ), uint32_t( w )*y+x, seed, smp, uint32_t( depth ) ) ;
					out[18] += 1. ;
					return true ;
				}
			}
		} ;
		for ( int l = 0 ; l<32 ; l++ ) { L[l].kind = policy != 1 ? int( K_REGEN ) : int( S_NONE ) ; L[l].pstate = 0 ; }
		int drain_k = 0 ;
		while ( true ) {
			int cnt[10] = { 0 } ; int vk[32] ;
			int n_fin = 0 ;
			for ( int l = 0 ; l<32 ; l++ ) {
				int k = L[l].kind ;
				if ( policy == 1 ) {
					if ( L[l].pstate == 1 || ( L[l].pstate == 0 && ! exhausted ) ) n_fin++ ;
					if ( k == K_SHADE ) k = L[l].pstate != 1 ? int( S_SWAP ) : int( S_WAIT ) ;                    // active ray finished
					else if ( k == S_NONE ) k = L[l].pstate == 2 ? int( S_SWAP ) : ( L[l].pstate == 1 || ! exhausted ) ? int( S_WAIT ) : int( K_DONE ) ;
				}
				vk[l] = k ; cnt[k]++ ;
			}
			int kind = K_DONE, best = 0 ;
			for ( int k = 1 ; k<=S_SWAP ; k++ ) if ( cnt[k] && ( ( cnt[k]<<3 )|k )>best ) { best = ( cnt[k]<<3 )|k ; kind = k ; }
			if ( policy == 1 && ( n_fin>=t1 || cnt[S_WAIT]>=t2 || ( kind == K_DONE && cnt[S_WAIT] ) ) ) kind = S_BATCH ;
			if ( kind == K_DONE ) break ;
			if ( policy == 3 ) {
				// drain variant: after a run of node steps, every other kind with enough waiting lanes
				// (t1: leaf / thing, t2: shade / regen) takes one step before the vote returns to nodes
				if ( drain_k == 0 && kind != K_NODE ) drain_k = K_LEAF ;
				if ( drain_k ) {
					int pick = 0 ;
					for ( int k = drain_k ; k<=K_REGEN ; k++ ) if ( cnt[k]>=( k<=K_THING ? t1 : t2 ) ) { pick = k ; break ; }
					if ( pick ) { kind = pick ; drain_k = pick+1 ; } else drain_k = 0 ;
					if ( drain_k>K_REGEN ) drain_k = 0 ;
				}
			}
			cost += K.vote ;
			switch ( kind ) {
				case K_NODE:
					while ( true ) {
						int n = 0 ;
						{ std::vector<const void*> seen ; for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_NODE ) { const void* a = ldp<HostPool, q4>( L[l].act, F_NODES0, 0 )+size_t( L[l].act.i( F_CUR, 0 ) )*RTX_NODE_RECS ; if ( std::find( seen.begin(), seen.end(), a ) == seen.end() ) seen.push_back( a ) ; } out[19] += double( seen.size() ) ; }
						for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_NODE ) { L[l].kind = step_node( L[l].act, 0, E.S ) ; n++ ; }
						out[K_NODE] += 1. ; out[8+K_NODE] += n ; cost += K.node ;
						int m = 0 ;
						for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_NODE ) m++ ;
						if ( m<sticky ) break ;
					}
					break ;
				case K_LEAF: {
					int n = 0, mx = 0 ;
					for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_LEAF ) {
						const int cn = int( ( uint32_t( ~L[l].act.i( F_CUR, 0 ) )&7u )+1u ) ; if ( cn>mx ) mx = cn ;
						L[l].kind = step_leaf( L[l].act, 0, E.S ) ; n++ ;
					}
					out[K_LEAF] += 1. ; out[8+K_LEAF] += n ; cost += K.leaf0+K.leaf1*mx ;
					break ;
				}
				case K_THING: {
					int n = 0 ; bool entered = false ;
					for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_THING ) {
						const unsigned long long before = g_stats.spheres ;
						L[l].kind = step_thing( L[l].act, 0, E.S ) ; n++ ;
						if ( g_stats.spheres == before ) entered = true ;
					}
					out[K_THING] += 1. ; out[8+K_THING] += n ; cost += entered ? K.thing1 : K.thing0 ;
					break ;
				}
				case K_SHADE: {   // policy 0 only
					int n = 0 ; bool hit = false ;
					for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_SHADE ) {
						f3 col, gn, ga ; bool g ; uint32_t sg ;
						if ( L[l].act.i( F_THING, 0 )>=0 ) hit = true ;
						L[l].kind = step_shade( L[l].act, 0, E.S, col, g, gn, ga, sg ) ; n++ ; out[16] += 1. ;
					}
					out[K_SHADE] += 1. ; out[8+K_SHADE] += n ; cost += hit ? K.shade : K.sky ;
					break ;
				}
				case K_REGEN: {   // policy 0 only
					int n = 0 ;
					for ( int l = 0 ; l<32 ; l++ ) if ( L[l].kind == K_REGEN ) {
						int k2 ;
						if ( serve( L[l].act, k2 ) ) L[l].kind = k2 ; else L[l].kind = K_DONE ;
						n++ ;
					}
					out[K_REGEN] += 1. ; out[8+K_REGEN] += n ; cost += K.regen ;
					break ;
				}
				case S_SWAP: {
					int n = 0 ;
					for ( int l = 0 ; l<32 ; l++ ) if ( vk[l] == S_SWAP ) {
						n++ ;
						const bool had = L[l].kind == K_SHADE ;
						const int ps = L[l].pstate ;
						std::swap( L[l].act, L[l].park ) ;
						L[l].pstate = had ? 1 : 0 ;
						L[l].kind = ps == 2 ? kind_of( L[l].act.i( F_CUR, 0 ), -1 ) : S_NONE ;
					}
					out[S_SWAP] += 1. ; out[8+S_SWAP] += n ; cost += K.swap ;
					break ;
				}
				case S_BATCH: {
					int n = 0 ; bool hit = false, regen = false ;
					for ( int l = 0 ; l<32 ; l++ ) {
						bool need = L[l].pstate == 0 ;
						if ( L[l].pstate == 1 ) {
							f3 col, gn, ga ; bool g ; uint32_t sg ;
							if ( L[l].park.i( F_THING, 0 )>=0 ) hit = true ;
							const int k2 = step_shade( L[l].park, 0, E.S, col, g, gn, ga, sg ) ; n++ ; out[16] += 1. ;
							if ( k2 == K_REGEN ) { L[l].pstate = 0 ; need = true ; } else L[l].pstate = 2 ;
						}
						if ( need && ! exhausted ) {
							int k2 ;
							if ( serve( L[l].park, k2 ) ) { L[l].pstate = 2 ; regen = true ; }
						}
					}
					out[S_BATCH] += 1. ; out[8+S_BATCH] += n ; cost += ( hit ? K.shade : K.sky )+( regen ? K.regen : 0. ) ;
					break ;
				}
			}
		}
	}
	out[17] = cost ;
	return 0 ;
}

// policy 2 of the simulator: R rays per warp whose state any lane can pick up (state in shared
// memory, compacted by kind before every step): each iteration takes up to 32 rays of the kind
// most rays are in.  out[] as emu_warpsim.
int g_sim_columns = 0 ;   // >0: slot r belongs to column r % columns, a step serves at most 32/columns rays per column
void emu_sim_columns( int c ) { g_sim_columns = c ; }
long long emu_warpsim_pool( const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int unit_spp, int units_per_warp, int tile_step, int depth, uint64_t seed,
		int R, int sticky, const double* costs, double* out ) {
	EmuScene E ;
	build_scene( E, things, n_things, n_meshes, vces, nv, ices, nt ) ;
	CameraDev c ;
	c.eye = mk3( float( cam[0] ), float( cam[1] ), float( cam[2] ) ) ; c.u = mk3( float( cam[3] ), float( cam[4] ), float( cam[5] ) ) ;
	c.v = mk3( float( cam[6] ), float( cam[7] ), float( cam[8] ) ) ; c.hvec = mk3( float( cam[9] ), float( cam[10] ), float( cam[11] ) ) ;
	c.wvec = mk3( float( cam[12] ), float( cam[13] ), float( cam[14] ) ) ; c.dvec = mk3( float( cam[15] ), float( cam[16] ), float( cam[17] ) ) ;
	c.aperture = float( cam[18] ) ;
	SimCost K ; memcpy( &K, costs, sizeof( K ) ) ;
	for ( int k = 0 ; k<24 ; k++ ) out[k] = 0. ;
	const int tiles_x = ( w+7 )/8, tiles_y = ( h+3 )/4, n_tiles = tiles_x*tiles_y ;
	std::vector<HostPool> P( R ) ; std::vector<int> kd( R ) ;
	double cost = 0. ;
	for ( int tile0 = 0 ; tile0<n_tiles ; tile0 += tile_step ) {
		int unit_next = 0 ; uint32_t unit_pos = 0, unit_left = 0, ux0 = 0, uy0 = 0, us0 = 0 ;
		auto serve = [&]( HostPool& Q, int& kind_out ) -> bool {
			while ( true ) {
				if ( unit_left == 0 ) {
					if ( unit_next>=units_per_warp ) return false ;
					const int tile = ( tile0+unit_next%4 )%n_tiles, chunk = unit_next/4 ;
					unit_next++ ;
					ux0 = uint32_t( tile%tiles_x )*8u ; uy0 = uint32_t( tile/tiles_x )*4u ; us0 = uint32_t( chunk*unit_spp ) ;
					unit_pos = 0 ; unit_left = 32u*uint32_t( unit_spp ) ;
				}
				const uint32_t idx = unit_pos++ ; unit_left-- ;
				const uint32_t px = idx&31u, smp = us0+( idx>>5 ) ;
				const uint32_t x = ux0+( px&7u ), y = uy0+( px>>3 ) ;
				if ( x<uint32_t( w ) && y<uint32_t( h ) ) {
					kind_out = step_regen( Q, 0, E.S, c, x, y, uint32_t( w ), uint32_t( h ), uint32_t( w )*y+x, seed, smp, uint32_t( depth ) ) ;
					out[18] += 1. ;
					return true ;
				}
			}
		} ;
		for ( int r = 0 ; r<R ; r++ ) kd[r] = K_REGEN ;
		int sel[32] ;
		auto pick = [&]( int kind ) {
			int n = 0 ;
			if ( g_sim_columns>0 ) {
				int per[64] = { 0 } ; const int cap = 32/g_sim_columns ;
				for ( int r = 0 ; r<R && n<32 ; r++ ) if ( kd[r] == kind && per[r%g_sim_columns]<cap ) { per[r%g_sim_columns]++ ; sel[n++] = r ; }
				return n ;
			}
			for ( int r = 0 ; r<R && n<32 ; r++ ) if ( kd[r] == kind ) sel[n++] = r ;
			return n ;
		} ;
		while ( true ) {
			int cnt[8] = { 0 } ;
			if ( g_sim_columns>0 ) {
				int per[8][64] = { { 0 } } ; const int cap = 32/g_sim_columns ;
				for ( int r = 0 ; r<R ; r++ ) if ( per[kd[r]][r%g_sim_columns]<cap ) { per[kd[r]][r%g_sim_columns]++ ; cnt[kd[r]]++ ; }
			} else
			for ( int r = 0 ; r<R ; r++ ) cnt[kd[r]]++ ;
			int kind = K_DONE, best = 0 ;
			for ( int k = 1 ; k<=K_REGEN ; k++ ) { const int cc = cnt[k]>32 ? 32 : cnt[k] ; if ( cc && ( ( cc<<3 )|k )>best ) { best = ( cc<<3 )|k ; kind = k ; } }
			if ( kind == K_DONE ) break ;
			cost += K.vote ;
			switch ( kind ) {
				case K_NODE:
					while ( true ) {
						const int n = pick( K_NODE ) ;
						{ std::vector<const void*> seen ; for ( int q = 0 ; q<n ; q++ ) { const void* a = ldp<HostPool, q4>( P[sel[q]], F_NODES0, 0 )+size_t( P[sel[q]].i( F_CUR, 0 ) )*RTX_NODE_RECS ; if ( std::find( seen.begin(), seen.end(), a ) == seen.end() ) seen.push_back( a ) ; } out[19] += double( seen.size() ) ; }
						for ( int q = 0 ; q<n ; q++ ) kd[sel[q]] = step_node( P[sel[q]], 0, E.S ) ;
						out[K_NODE] += 1. ; out[8+K_NODE] += n ; cost += K.node ;
						int m = 0 ;
						for ( int r = 0 ; r<R ; r++ ) if ( kd[r] == K_NODE ) m++ ;
						if ( m<sticky ) break ;
					}
					break ;
				case K_LEAF: {
					const int n = pick( K_LEAF ) ; int mx = 0 ;
					for ( int q = 0 ; q<n ; q++ ) {
						const int cn = int( ( uint32_t( ~P[sel[q]].i( F_CUR, 0 ) )&7u )+1u ) ; if ( cn>mx ) mx = cn ;
						kd[sel[q]] = step_leaf( P[sel[q]], 0, E.S ) ;
					}
					out[K_LEAF] += 1. ; out[8+K_LEAF] += n ; cost += K.leaf0+K.leaf1*mx ;
					break ;
				}
				case K_THING: {
					const int n = pick( K_THING ) ; bool entered = false ;
					for ( int q = 0 ; q<n ; q++ ) {
						const unsigned long long before = g_stats.spheres ;
						kd[sel[q]] = step_thing( P[sel[q]], 0, E.S ) ;
						if ( g_stats.spheres == before ) entered = true ;
					}
					out[K_THING] += 1. ; out[8+K_THING] += n ; cost += entered ? K.thing1 : K.thing0 ;
					break ;
				}
				case K_SHADE: {
					const int n = pick( K_SHADE ) ; bool hit = false ;
					for ( int q = 0 ; q<n ; q++ ) {
						f3 col, gn, ga ; bool g ; uint32_t sg ;
						if ( P[sel[q]].i( F_THING, 0 )>=0 ) hit = true ;
						kd[sel[q]] = step_shade( P[sel[q]], 0, E.S, col, g, gn, ga, sg ) ; out[16] += 1. ;
					}
					out[K_SHADE] += 1. ; out[8+K_SHADE] += n ; cost += hit ? K.shade : K.sky ;
					break ;
				}
				default: {
					const int n = pick( K_REGEN ) ;
					for ( int q = 0 ; q<n ; q++ ) { int k2 ; kd[sel[q]] = serve( P[sel[q]], k2 ) ? k2 : K_DONE ; }
					out[K_REGEN] += 1. ; out[8+K_REGEN] += n ; cost += K.regen ;
				}
			}
		}
	}
	out[17] = cost ;
	return 0 ;
}

// closest hits of caller-supplied rays through the two-level hierarchy (or by exhaustive scan)
int emu_trace_rays( const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		int n_rays, const float* ori, const float* dir, float tmin, int brute, int use_pool, int64_t* id, float* t_out ) {
	EmuScene E ;
	build_scene( E, things, n_things, n_meshes, vces, nv, ices, nt ) ;
	HostStack st ;
	static HostPool P ;
	for ( int r = 0 ; r<n_rays ; r++ ) {
		const f3 o = mk3( ori[3*r], ori[3*r+1], ori[3*r+2] ), d = mk3( dir[3*r], dir[3*r+1], dir[3*r+2] ) ;
		HitRec hr ;
		if ( brute ) closest_brute( E.S, o, d, tmin, hr ) ;
		else if ( ! use_pool ) closest( E.S, o, d, tmin, st, hr ) ;
		else {
			// the step functions of the render kernel (they use tmin = 1e-3)
			begin_ray( P, 0, E.S, o, d ) ;
			int kind = kind_of( P.i( F_CUR, 0 ), -1 ) ;
			while ( kind != K_SHADE )
				kind = kind == K_NODE ? step_node( P, 0, E.S ) : kind == K_LEAF ? step_leaf( P, 0, E.S ) : step_thing( P, 0, E.S ) ;
			hr.t = P.f( F_T, 0 ) ; hr.thing = P.i( F_THING, 0 ) ; hr.prim = P.i( F_PRIM, 0 ) ;
		}
		id[r] = hr.thing<0 ? int64_t( -1 ) : ( ( int64_t( hr.thing )<<32 )|int64_t( uint32_t( hr.prim+1 ) ) ) ;
		t_out[r] = hr.thing<0 ? -1.f : hr.t ;
	}
	return 0 ;
}

// the bounding-sphere pre-test alone (rtx_core.cuh bsphere_miss), for the conservativeness test
void emu_bsphere_miss( int n, const float* bs, const float* ori, const float* dir, const float* tmin, const float* tbest, uint8_t* out ) {
	for ( int r = 0 ; r<n ; r++ ) {
		const q4 b = { bs[4*r], bs[4*r+1], bs[4*r+2], bs[4*r+3] } ;
		out[r] = bsphere_miss( b, mk3( ori[3*r], ori[3*r+1], ori[3*r+2] ), mk3( dir[3*r], dir[3*r+1], dir[3*r+2] ), tmin[r], tbest[r] ) ? 1 : 0 ;
	}
}
// world_bsphere of rtx_hostmath.h (the padded sphere the pre-test is given)
void emu_world_bsphere( const float* xf, const double* bs, float* out ) { world_bsphere( xf, bs, out ) ; }

void emu_sp_hist( unsigned long long* out, int reset ) { for ( int k = 0 ; k<128 ; k++ ) { out[k] = g_sp_hist[k] ; if ( reset ) g_sp_hist[k] = 0 ; } }

// traversal counters since the last reset: rays, nodes, leaves, tris, things, spheres, enters, pushes, max stack
void emu_stats( unsigned long long* out, int reset ) {
	const unsigned long long v[9] = { g_stats.rays, g_stats.nodes, g_stats.leaves, g_stats.tris, g_stats.things, g_stats.spheres, g_stats.enters, g_stats.pushes, g_stats.maxsp } ;
	for ( int k = 0 ; k<9 ; k++ ) out[k] = v[k] ;
	if ( reset ) g_stats = Stats{ 0, 0, 0, 0, 0, 0, 0, 0, 0 } ;
}

} // extern "C"
