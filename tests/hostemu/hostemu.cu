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

#include "../../rtxplay_b200/csrc/rtx_core.cuh"
#include "../../rtxplay_b200/csrc/rtx_lbvh.cuh"
#include "../../rtxplay_b200/csrc/rtx_hostmath.h"

using namespace rtx ;

namespace {

struct HostStack {
	int32_t v[256] ; int sp ;
	RTX_HD void reset() { sp = 0 ; }
	RTX_HD bool empty() const { return sp == 0 ; }
	RTX_HD void push( int32_t x ) { v[sp++] = x ; }
	RTX_HD int32_t pop() { return v[--sp] ; }
} ;

struct Tree {
	std::vector<q4>       nodes ;
	std::vector<uint32_t> order ;
	q4 root_lo, root_hi ;
} ;

// the same pipeline as lbvh_build(), serial: keys -> stable sort -> karras -> refit -> emit
void build_tree( const std::vector<q4>& plo, const std::vector<q4>& phi, Tree& T ) {
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
	T.nodes.assign( size_t( n>1 ? n-1 : 1 )*RTX_NODE_RECS, q4{ 0, 0, 0, 0 } ) ;
	if ( n == 1 ) {
		const int ref = ~0 ;
		T.nodes[0] = { blo[0].x, blo[0].y, blo[0].z, asfloat( ref ) } ;
		T.nodes[1] = { bhi[0].x, bhi[0].y, bhi[0].z, asfloat( ref ) } ;
		T.nodes[2] = T.nodes[0] ; T.nodes[3] = T.nodes[1] ;
		T.root_lo = blo[0] ; T.root_hi = bhi[0] ;
		return ;
	}
	std::vector<int> cl( n-1 ), cr( n-1 ) ;
	for ( int i = 0 ; i<n-1 ; i++ ) {
		int l, r ; bool ll, rl ;
		karras_node( sk.data(), n, i, l, r, ll, rl ) ;
		cl[i] = ll ? ~l : l ; cr[i] = rl ? ~r : r ;
	}
	// post-order refit without recursion
	std::vector<int> todo ; std::vector<char> seen( n-1, 0 ) ;
	todo.push_back( 0 ) ;
	while ( ! todo.empty() ) {
		const int i = todo.back() ;
		if ( ! seen[i] ) {
			seen[i] = 1 ;
			if ( cl[i]>=0 ) todo.push_back( cl[i] ) ;
			if ( cr[i]>=0 ) todo.push_back( cr[i] ) ;
			continue ;
		}
		todo.pop_back() ;
		const int a = cl[i]<0 ? n-1+( ~cl[i] ) : cl[i], b = cr[i]<0 ? n-1+( ~cr[i] ) : cr[i] ;
		blo[i] = { fminf( blo[a].x, blo[b].x ), fminf( blo[a].y, blo[b].y ), fminf( blo[a].z, blo[b].z ), 0.f } ;
		bhi[i] = { fmaxf( bhi[a].x, bhi[b].x ), fmaxf( bhi[a].y, bhi[b].y ), fmaxf( bhi[a].z, bhi[b].z ), 0.f } ;
	}
	for ( int i = 0 ; i<n-1 ; i++ ) {
		const int a = cl[i]<0 ? n-1+( ~cl[i] ) : cl[i], b = cr[i]<0 ? n-1+( ~cr[i] ) : cr[i] ;
		const int ra = cl[i]<0 ? ~( ( ~cl[i] )<<2 ) : cl[i], rb = cr[i]<0 ? ~( ( ~cr[i] )<<2 ) : cr[i] ;
		q4* o = T.nodes.data()+size_t( i )*RTX_NODE_RECS ;
		o[0] = { blo[a].x, blo[a].y, blo[a].z, asfloat( ra ) } ;
		o[1] = { bhi[a].x, bhi[a].y, bhi[a].z, asfloat( rb ) } ;
		o[2] = { blo[b].x, blo[b].y, blo[b].z, 0.f } ;
		o[3] = { bhi[b].x, bhi[b].y, bhi[b].z, 0.f } ;
	}
	T.root_lo = blo[0] ; T.root_hi = bhi[0] ;
}

struct EmuMesh { std::vector<float> vces ; std::vector<uint32_t> ices ; std::vector<q4> tris ; Tree tree ; } ;

struct EmuScene {
	std::vector<EmuMesh>    meshes ;
	std::vector<ThingTrav>  trav ;
	std::vector<ThingShade> shade ;
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
		std::vector<q4> plo( nt[q] ), phi( nt[q] ) ;
		for ( uint32_t f = 0 ; f<nt[q] ; f++ ) {
			const float* a = &m.vces[3*size_t( m.ices[3*f] )] ; const float* b = &m.vces[3*size_t( m.ices[3*f+1] )] ; const float* c = &m.vces[3*size_t( m.ices[3*f+2] )] ;
			plo[f] = { fminf( a[0], fminf( b[0], c[0] ) ), fminf( a[1], fminf( b[1], c[1] ) ), fminf( a[2], fminf( b[2], c[2] ) ), 0.f } ;
			phi[f] = { fmaxf( a[0], fmaxf( b[0], c[0] ) ), fmaxf( a[1], fmaxf( b[1], c[1] ) ), fmaxf( a[2], fmaxf( b[2], c[2] ) ), 0.f } ;
		}
		build_tree( plo, phi, m.tree ) ;
		m.tris.resize( 3*size_t( nt[q] ) ) ;
		for ( uint32_t j = 0 ; j<nt[q] ; j++ ) {
			const uint32_t f = m.tree.order[j] ;
			const float* a = &m.vces[3*size_t( m.ices[3*f] )] ; const float* b = &m.vces[3*size_t( m.ices[3*f+1] )] ; const float* c = &m.vces[3*size_t( m.ices[3*f+2] )] ;
			m.tris[3*size_t( j )]   = { a[0], a[1], a[2], asfloat( int( f ) ) } ;
			m.tris[3*size_t( j )+1] = { b[0]-a[0], b[1]-a[1], b[2]-a[2], 0.f } ;
			m.tris[3*size_t( j )+2] = { c[0]-a[0], c[1]-a[1], c[2]-a[2], 0.f } ;
		}
	}
	E.trav.resize( n_things ) ; E.shade.resize( n_things ) ;
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
			t.inv[0] = double( xf[3] ) ; t.inv[1] = double( xf[7] ) ; t.inv[2] = double( xf[11] ) ; t.inv[3] = double( xf[0] ) ;
			const double r = fabs( t.inv[3] ) ;
			plo[k] = { float( t.inv[0]-r )-1e-3f, float( t.inv[1]-r )-1e-3f, float( t.inv[2]-r )-1e-3f, 0.f } ;
			phi[k] = { float( t.inv[0]+r )+1e-3f, float( t.inv[1]+r )+1e-3f, float( t.inv[2]+r )+1e-3f, 0.f } ;
		} else {
			const EmuMesh& m = E.meshes[int( row[TH_MESH] )] ;
			t.kind = 1 ; s.kind = 1 ;
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
	if ( n_things ) build_tree( plo, phi, E.tlas ) ;
	E.S.tlas_nodes = E.tlas.nodes.data() ; E.S.tlas_order = E.tlas.order.data() ;
	E.S.trav = E.trav.data() ; E.S.shade = E.shade.data() ; E.S.n_things = uint32_t( n_things ) ;
}

} // namespace

extern "C" {

// cam: 19 doubles (eye,u,v,hvec,wvec,dvec,aperture) as in the oracle tables
int emu_render( const double* things, int n_things, int n_meshes, const float* const* vces, const uint32_t* nv, const uint32_t* const* ices, const uint32_t* nt,
		const double* cam, int w, int h, int spp, int depth, uint64_t seed, int sample0, int sample_stride,
		uint64_t* fix, uint32_t* rpp, int64_t* first_id, float* first_t, int brute ) {
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
				primary_ray( c, uint32_t( x ), uint32_t( y ), uint32_t( w ), uint32_t( h ), rng, ori, dir ) ;
				if ( k == 0 && first_id ) {
					HitRec hr ;
					if ( brute ) closest_brute( E.S, ori, dir, 1e-3f, hr ) ;
					else         closest( E.S, ori, dir, 1e-3f, st, hr ) ;
					first_id[pix] = hr.thing<0 ? int64_t( -1 ) : ( ( int64_t( hr.thing )<<32 )|int64_t( uint32_t( hr.prim+1 ) ) ) ;
					if ( first_t ) first_t[pix] = hr.thing<0 ? -1.f : hr.t ;
				}
				const f3 col = path_radiance( E.S, ori, dir, uint32_t( depth ), rng, st, segments ) ;
				acc[0] += tofix( col.x ) ; acc[1] += tofix( col.y ) ; acc[2] += tofix( col.z ) ;
			}
			if ( fix ) { fix[3*size_t( pix )] = acc[0] ; fix[3*size_t( pix )+1] = acc[1] ; fix[3*size_t( pix )+2] = acc[2] ; }
			if ( rpp ) rpp[pix] = segments ;
		}
	return 0 ;
}

} // extern "C"
