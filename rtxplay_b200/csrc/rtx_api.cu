// rtx_api.cu -- the C ABI of include/rtx.h: context, geometry upload, LBVH builds,
// frame buffers and kernel launches.  One translation unit, compiled for sm_100a with
// -fmad=false (see rtx_core.cuh for why).  There is no CPU path in here: every entry
// point that computes launches CUDA kernels, and rtx_init fails without a device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtx.h"
#include "rtx_core.cuh"
#include "rtx_lbvh.cuh"
#include "rtx_kernels.cuh"
#include "rtx_qkernel.cuh"
#include "rtx_hostmath.h"

using namespace rtx ;

#ifndef RTX_DEFAULT_KERNEL
#define RTX_DEFAULT_KERNEL 0            // 0: k_render (one ray per lane, state in registers), 1: k_render_q (compacting ray pool); RTX_KERNEL=reg|q overrides
#endif
#ifndef RTX_Q_DEFAULT_CARVEOUT
#define RTX_Q_DEFAULT_CARVEOUT 75       // k_render_q: 15 CTAs x 12.9 KB of ray slots (64 slots of 176 bytes + queues) take 196 KB of the 228 KB; 60 %: 12 CTAs, 743 ms against 655
#endif
#ifndef RTX_DEFAULT_CARVEOUT
// 20 render CTAs x (4 KB stack + 1 KB the driver reserves) = 100 KB of shared memory; the rest
// of the 228 KB is L1 for the hierarchy.  Measured: 30-35 % best, 40-44 % 1 % slower, 50 % 3-5 %
// slower, 25 % loses resident warps
#define RTX_DEFAULT_CARVEOUT 35
#endif

static_assert( sizeof( ThingTrav ) == 128, "ThingTrav layout" ) ;
static_assert( sizeof( ThingShade ) == 160, "ThingShade layout" ) ;
static_assert( sizeof( q4 ) == 16, "q4 layout" ) ;

namespace {

thread_local std::string g_init_error ;

#define CK( call ) do { cudaError_t e_ = ( call ) ; if ( e_ != cudaSuccess ) { \
	char b_[512] ; snprintf( b_, sizeof( b_ ), "CUDA call (%s) failed with error '%s' (%s:%d)", #call, cudaGetErrorString( e_ ), __FILE__, __LINE__ ) ; \
	throw std::runtime_error( b_ ) ; } } while ( 0 )

struct Lbvh {           // one built tree: wide traversal nodes + (when kept) the binary tree a refit needs
	uint32_t  n = 0 ;
	uint32_t  n_nodes = 0 ;       // wide nodes in use
	uint32_t  cap_nodes = 0 ;     // wide nodes allocated
	q4*       nodes = nullptr ;
	uint32_t* order = nullptr ;   // leaf slot -> primitive
	int2*     child = nullptr ;   // binary tree (Karras)
	int2*     range = nullptr ;
	int*      parent_inner = nullptr ;
	int*      parent_leaf = nullptr ;
	q4*       blo = nullptr ;
	q4*       bhi = nullptr ;
	uint32_t* flags = nullptr ;
	int2*     front0 = nullptr ;  // collapse frontiers
	int2*     front1 = nullptr ;
	uint32_t* counters = nullptr ;
	q4        root_lo = { 0, 0, 0, 0 }, root_hi = { 0, 0, 0, 0 } ;
} ;

struct Mesh {
	bool      analytic = false ;
	double    bsphere[4] = { 0, 0, 0, 0 } ;   // object-space bounding sphere
	uint32_t  nv = 0, nt = 0 ;
	float*    vces = nullptr ;
	uint32_t* ices = nullptr ;
	q4*       tris = nullptr ;
	Lbvh      bvh ;
} ;

struct ThingHost {
	uint32_t   mesh ;
	rtx_optics optics ;
	float      xf[12] ;
} ;

} // namespace

struct rtx_ctx {
	int          device = 0 ;
	std::string  err ;
	cudaStream_t stream = nullptr ;
	cudaMemPool_t pool = nullptr ;   // build workspace and hierarchies (talloc)
	cudaEvent_t  ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev_in = nullptr ;   // frame: before the kernel, behind it, behind the resolve; start of the call
	float        ms_frame = 0.f, ms_resolve = 0.f, ms_postproc = 0.f ;

	std::vector<Mesh>      meshes ;
	std::vector<ThingHost> things ;

	// device scene
	ThingTrav*  d_trav = nullptr ;
	ThingShade* d_shade = nullptr ;
	q4*         d_bsphere = nullptr ;
	q4*         d_tb_lo = nullptr ;   // per-thing mesh root boxes, then world boxes
	q4*         d_tb_hi = nullptr ;
	q4*         d_tp_lo = nullptr ;
	q4*         d_tp_hi = nullptr ;
	uint32_t    n_things_dev = 0 ;
	Lbvh        tlas ;
	bool        built = false ;

	// frame
	uint32_t  w = 0, h = 0 ;
	uint64_t* d_accum = nullptr ;
	float*    d_raw = nullptr ;
	uint32_t* d_rpp = nullptr ;
	uchar4*   d_image = nullptr ;
	int64_t*  d_hit_id = nullptr ;
	float*    d_hit_t = nullptr ;
	float*    d_normals = nullptr ;
	float*    d_albedos = nullptr ;
	long long* d_guide_acc = nullptr ;   // allocated by the first frame that asks for guide layers
	bool      guides_valid = false ;
	uint32_t* d_pick = nullptr ;
	uint32_t* d_fault = nullptr ;    // SceneDev::fault: set by a traversal whose stack overflowed
	uint32_t* h_fault = nullptr ;    // pinned copy read back behind every tracing launch
	uint64_t  stack_faults = 0 ;     // launches that reported one
	unsigned long long* d_counter = nullptr ;
	uint32_t* d_tile_counter = nullptr ;
	int32_t*  d_ovf = nullptr ;      // overflow stacks of the resident render warps
	uint32_t  render_grid = 0 ;
	// multi-GPU (rtx_init_multi): the context the caller holds is the root; its replicas, one per
	// further device, receive every scene call, render their share of the samples, and the root
	// sums their accumulation buffers through peer memory (k_reduce_resolve)
	std::vector<rtx_ctx*> replicas ;
	cudaEvent_t ev_done = nullptr ;           // a replica's frame kernel has finished (the root's stream waits on it)
	std::vector<uint64_t*>  stage_accum ;     // root-side copies of replica buffers where peer access is not available
	std::vector<long long*> stage_guide ;
	std::vector<char>       peer_ok ;         // per replica: the root reads its memory directly
	int       kernel = RTX_DEFAULT_KERNEL ;   // which path-tracing kernel do_render launches
	uint32_t  wide_grid = 0 ;        // resident CTAs of the cooperative k_wide_all (0: no cooperative launch on this device)
	uint32_t  q_grid = 0 ;           // k_render_q: resident warps, their cold ray records and overflow stacks
	q4*       d_qcold = nullptr ;
	int32_t*  d_qovf = nullptr ;

	// statistics
	uint64_t bytes = 0 ;
	uint32_t launches = 0 ;
	float    ms_render = 0.f, ms_blas = 0.f, ms_tlas = 0.f ;
	cudaEvent_t stage_ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr } ;   // marks between the stages of the last lbvh_build
	bool     stage_full = false ;                      // ... a full build (all six marks) or a refit (marks 3-5)
	float    ms_stage_blas[5] = { 0.f, 0.f, 0.f, 0.f, 0.f }, ms_stage_tlas[5] = { 0.f, 0.f, 0.f, 0.f, 0.f } ;   // keys, sort, hierarchy, boxes, wide nodes
	uint64_t paths = 0 ;
} ;

namespace {

template <class T> T* dalloc( rtx_ctx* c, size_t n ) {
	T* p = nullptr ;
	if ( n == 0 ) n = 1 ;
	CK( cudaMalloc( reinterpret_cast<void**>( &p ), n*sizeof( T ) ) ) ;
	c->bytes += n*sizeof( T ) ;
	return p ;
}
template <class T> void dfree( rtx_ctx* c, T*& p, size_t n ) {
	if ( p ) { cudaFree( p ) ; c->bytes -= ( n ? n : 1 )*sizeof( T ) ; p = nullptr ; }
}

// build workspace and hierarchy arrays: stream-ordered allocations from the context's own
// pool (kept, not returned to the driver, between builds) -- a cudaMalloc/cudaFree pair costs
// more than the kernels of a small build and cudaFree synchronises the device
template <class T> T* talloc( rtx_ctx* c, size_t n ) {
	T* p = nullptr ;
	if ( n == 0 ) n = 1 ;
	CK( cudaMallocFromPoolAsync( reinterpret_cast<void**>( &p ), n*sizeof( T ), c->pool, c->stream ) ) ;
	c->bytes += n*sizeof( T ) ;
	return p ;
}
template <class T> void tfree( rtx_ctx* c, T*& p, size_t n ) {
	if ( p ) { cudaFreeAsync( p, c->stream ) ; c->bytes -= ( n ? n : 1 )*sizeof( T ) ; p = nullptr ; }
}

void lbvh_free_binary( rtx_ctx* c, Lbvh& b ) {
	const size_t n = b.n ;
	tfree( c, b.child, n>1 ? n-1 : 1 ) ;
	tfree( c, b.range, n>1 ? n-1 : 1 ) ;
	tfree( c, b.parent_inner, n>1 ? n-1 : 1 ) ;
	tfree( c, b.parent_leaf, n ) ;
	tfree( c, b.blo, 2*n ) ;
	tfree( c, b.bhi, 2*n ) ;
	tfree( c, b.flags, n>1 ? n-1 : 1 ) ;
	tfree( c, b.front0, n ) ;
	tfree( c, b.front1, n ) ;
	tfree( c, b.counters, 2 ) ;
}

void lbvh_free( rtx_ctx* c, Lbvh& b ) {
	lbvh_free_binary( c, b ) ;
	tfree( c, b.nodes, size_t( b.cap_nodes )*RTX_NODE_RECS ) ;
	tfree( c, b.order, b.n ) ;
	b.n = 0 ; b.n_nodes = 0 ; b.cap_nodes = 0 ;
}

// bottom-up boxes of the binary tree, then its collapse into wide traversal nodes, level
// by level (the host reads back the size of each next frontier)
void lbvh_refit( rtx_ctx* c, Lbvh& b, const q4* plo, const q4* phi, int leaf_max ) {
	const int n = int( b.n ) ;
	CK( cudaEventRecord( c->stage_ev[3], c->stream ) ) ;
	CK( cudaMemsetAsync( b.flags, 0, sizeof( uint32_t )*( n>1 ? n-1 : 1 ), c->stream ) ) ;
	k_refit<<<( n+255 )/256, 256, 0, c->stream>>>( plo, phi, b.order, n, b.child, b.parent_inner, b.parent_leaf, b.blo, b.bhi, b.flags ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	CK( cudaEventRecord( c->stage_ev[4], c->stream ) ) ;
	const int2 root = make_int2( 0, 0 ) ;
	uint32_t counters[2] = { 1u, 0u } ;   // wide node 0 is the root
	CK( cudaMemcpyAsync( b.front0, &root, sizeof( int2 ), cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaMemcpyAsync( b.counters, counters, sizeof( counters ), cudaMemcpyHostToDevice, c->stream ) ) ;
	if ( c->wide_grid ) {
		// every level in one cooperative launch (frontier loop and level barriers on the device)
		int n_arg = n, lm = leaf_max ;
		void* args[] = { &b.front0, &b.front1, &n_arg, &lm, &b.child, &b.range, &b.blo, &b.bhi, &b.nodes, &b.counters } ;
		CK( cudaLaunchCooperativeKernel( reinterpret_cast<const void*>( k_wide_all ), dim3( c->wide_grid ), dim3( 128 ), args, 0, c->stream ) ) ;
		c->launches += 1 ;
		CK( cudaMemcpyAsync( counters, b.counters, sizeof( counters ), cudaMemcpyDeviceToHost, c->stream ) ) ;
	} else {
		// (devices without cooperative launch: one launch per level, the host reads each frontier size)
		uint32_t n_front = 1 ;
		int2* fin = b.front0 ; int2* fout = b.front1 ;
		while ( n_front ) {
			k_wide_level<<<( n_front+127 )/128, 128, 0, c->stream>>>( fin, n_front, n, leaf_max, b.child, b.range, b.blo, b.bhi, b.nodes, fout, b.counters ) ;
			c->launches += 1 ;
			CK( cudaGetLastError() ) ;
			CK( cudaMemcpyAsync( counters, b.counters, sizeof( counters ), cudaMemcpyDeviceToHost, c->stream ) ) ;
			CK( cudaStreamSynchronize( c->stream ) ) ;
			n_front = counters[1] ;
			counters[1] = 0 ;
			CK( cudaMemcpyAsync( b.counters, counters, sizeof( counters ), cudaMemcpyHostToDevice, c->stream ) ) ;
			std::swap( fin, fout ) ;
		}
	}
	CK( cudaEventRecord( c->stage_ev[5], c->stream ) ) ;
	c->stage_full = false ;
	CK( cudaMemcpyAsync( &b.root_lo, b.blo, sizeof( q4 ), cudaMemcpyDeviceToHost, c->stream ) ) ;
	CK( cudaMemcpyAsync( &b.root_hi, b.bhi, sizeof( q4 ), cudaMemcpyDeviceToHost, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	b.n_nodes = counters[0] ;
}

// Morton keys -> radix sort -> Karras hierarchy -> refit -> wide nodes
void lbvh_build( rtx_ctx* c, Lbvh& b, const q4* plo, const q4* phi, uint32_t n, int leaf_max, bool keep_binary ) {
	lbvh_free( c, b ) ;
	b.n = n ;
	// a wide node replaces a binary inner node covering more than leaf_max primitives (or the
	// root): there are fewer than n/leaf_max+1 ... n-1 of those; allocate the safe bound,
	// shrink after the build
	b.cap_nodes    = n>1 ? n-1 : 1 ;
	b.nodes        = talloc<q4>( c, size_t( b.cap_nodes )*RTX_NODE_RECS ) ;
	b.order        = talloc<uint32_t>( c, n ) ;
	b.child        = talloc<int2>( c, n>1 ? n-1 : 1 ) ;
	b.range        = talloc<int2>( c, n>1 ? n-1 : 1 ) ;
	b.parent_inner = talloc<int>( c, n>1 ? n-1 : 1 ) ;
	b.parent_leaf  = talloc<int>( c, n ) ;
	b.blo          = talloc<q4>( c, 2*size_t( n ) ) ;
	b.bhi          = talloc<q4>( c, 2*size_t( n ) ) ;
	b.flags        = talloc<uint32_t>( c, n>1 ? n-1 : 1 ) ;
	b.front0       = talloc<int2>( c, n ) ;
	b.front1       = talloc<int2>( c, n ) ;
	b.counters     = talloc<uint32_t>( c, 2 ) ;

	const uint32_t nblocks = ( n+RTX_RS_TILE-1 )/RTX_RS_TILE ;
	int* bounds = talloc<int>( c, 6 ) ;
	uint64_t* keys0 = talloc<uint64_t>( c, n ) ; uint64_t* keys1 = talloc<uint64_t>( c, n ) ;
	uint32_t* vals1 = talloc<uint32_t>( c, n ) ;
#if ! RTX_SORT_ONESWEEP
	uint32_t* counts = talloc<uint32_t>( c, size_t( 256 )*nblocks ) ;
	const uint32_t n_chunks = ( 256u*nblocks+RTX_SCAN_CHUNK-1u )/RTX_SCAN_CHUNK ;   // count-table scan: one block up to 16 K counters, else chunked
	uint32_t* chunk_sums = talloc<uint32_t>( c, n_chunks ) ;
#else
	( void ) nblocks ;
#endif

	CK( cudaEventRecord( c->stage_ev[0], c->stream ) ) ;
	k_bounds_init<<<1, 32, 0, c->stream>>>( bounds ) ;
	k_bounds_reduce<<<min( 1184u, ( n+255u )/256u ), 256, 0, c->stream>>>( plo, phi, n, bounds ) ;
	k_morton<<<( n+255 )/256, 256, 0, c->stream>>>( plo, phi, n, bounds, keys0, b.order ) ;
	c->launches += 3 ;
	CK( cudaEventRecord( c->stage_ev[1], c->stream ) ) ;
	uint64_t* kin = keys0 ; uint64_t* kout = keys1 ;
	uint32_t* vin = b.order ; uint32_t* vout = vals1 ;
#if RTX_SORT_ONESWEEP
	// one read and one write of every key per digit: all eight histograms first, then one
	// look-back pass per digit (rtx_lbvh.cuh)
	const uint32_t ntiles = ( n+RTX_OS_TILE-1u )/RTX_OS_TILE ;
	uint32_t* os_hist = talloc<uint32_t>( c, 8*256+8 ) ;      // [8][256] counts -> bucket starts, then the eight tile tickets
	uint32_t* os_status = talloc<uint32_t>( c, size_t( 256 )*ntiles ) ;
	CK( cudaMemsetAsync( os_hist, 0, sizeof( uint32_t )*( 8*256+8 ), c->stream ) ) ;
	k_radix_hist8<<<std::min( 1184u, ( n+255u )/256u ), 256, 0, c->stream>>>( keys0, n, os_hist ) ;
	k_radix_bases<<<8, 256, 0, c->stream>>>( os_hist ) ;
	c->launches += 2 ;
	for ( int pass = 0 ; pass<8 ; pass++ ) {
		CK( cudaMemsetAsync( os_status, 0, sizeof( uint32_t )*size_t( 256 )*ntiles, c->stream ) ) ;
		k_radix_onesweep<<<ntiles, RTX_OS_THREADS, RTX_OS_SMEM_BYTES, c->stream>>>( kin, vin, n, 8*pass, os_hist+256*pass, os_status, os_hist+8*256+pass, kout, vout ) ;
		c->launches += 1 ;
		std::swap( kin, kout ) ; std::swap( vin, vout ) ;
	}
	tfree( c, os_hist, 8*256+8 ) ; tfree( c, os_status, size_t( 256 )*ntiles ) ;
#else
	for ( int pass = 0 ; pass<8 ; pass++ ) {
		const int shift = 8*pass ;
		k_radix_hist<<<nblocks, 32*RTX_RS_WARPS, 0, c->stream>>>( kin, n, shift, counts, nblocks ) ;
		if ( n_chunks<=1u ) k_radix_scan<<<1, 1024, 0, c->stream>>>( counts, 256u*nblocks ) ;
		else {
			k_scan_chunks<<<n_chunks, 1024, 0, c->stream>>>( counts, 256u*nblocks, chunk_sums ) ;
			k_radix_scan<<<1, 1024, 0, c->stream>>>( chunk_sums, n_chunks ) ;
			k_scan_add<<<n_chunks, 1024, 0, c->stream>>>( counts, 256u*nblocks, chunk_sums ) ;
			c->launches += 2 ;
		}
		k_radix_scatter<<<nblocks, 32*RTX_RS_WARPS, 0, c->stream>>>( kin, vin, n, shift, counts, nblocks, kout, vout ) ;
		c->launches += 3 ;
		std::swap( kin, kout ) ; std::swap( vin, vout ) ;
	}
#endif
	// 8 passes: the sorted data is back in keys0 / b.order
	CK( cudaEventRecord( c->stage_ev[2], c->stream ) ) ;
	if ( n>1 ) {
		k_karras<<<( n-1+255 )/256, 256, 0, c->stream>>>( keys0, int( n ), b.child, b.range, b.parent_inner, b.parent_leaf ) ;
		c->launches += 1 ;
	}
	CK( cudaGetLastError() ) ;
	lbvh_refit( c, b, plo, phi, leaf_max ) ;
	c->stage_full = true ;

	tfree( c, bounds, 6 ) ; tfree( c, keys0, n ) ; tfree( c, keys1, n ) ; tfree( c, vals1, n ) ;
#if ! RTX_SORT_ONESWEEP
	tfree( c, counts, size_t( 256 )*nblocks ) ; tfree( c, chunk_sums, n_chunks ) ;
#endif
	if ( ! keep_binary ) {
		// a mesh is never refitted: drop the binary tree and trim the node array
		lbvh_free_binary( c, b ) ;
		if ( b.n_nodes<b.cap_nodes ) {
			q4* trimmed = talloc<q4>( c, size_t( b.n_nodes )*RTX_NODE_RECS ) ;
			CK( cudaMemcpyAsync( trimmed, b.nodes, sizeof( q4 )*size_t( b.n_nodes )*RTX_NODE_RECS, cudaMemcpyDeviceToDevice, c->stream ) ) ;
			tfree( c, b.nodes, size_t( b.cap_nodes )*RTX_NODE_RECS ) ;
			b.nodes = trimmed ; b.cap_nodes = b.n_nodes ;
		}
	}
}

// stage times of the last lbvh_build / lbvh_refit (call after the stream is synchronised)
void stage_collect( rtx_ctx* c, float* dst, bool add ) {
	for ( int k = 0 ; k<5 ; k++ ) {
		float ms = 0.f ;
		if ( c->stage_full || k>=3 ) CK( cudaEventElapsedTime( &ms, c->stage_ev[k], c->stage_ev[k+1] ) ) ;
		dst[k] = add ? dst[k]+ms : ms ;
	}
}

// host thing list -> device records + world bounds
void upload_things( rtx_ctx* c ) {
	const uint32_t n = uint32_t( c->things.size() ) ;
	if ( n != c->n_things_dev ) {
		dfree( c, c->d_trav, c->n_things_dev ) ; dfree( c, c->d_shade, c->n_things_dev ) ; dfree( c, c->d_bsphere, c->n_things_dev ) ;
		dfree( c, c->d_tb_lo, c->n_things_dev ) ; dfree( c, c->d_tb_hi, c->n_things_dev ) ;
		dfree( c, c->d_tp_lo, c->n_things_dev ) ; dfree( c, c->d_tp_hi, c->n_things_dev ) ;
		c->d_trav = dalloc<ThingTrav>( c, n ) ; c->d_shade = dalloc<ThingShade>( c, n ) ; c->d_bsphere = dalloc<q4>( c, n ) ;
		c->d_tb_lo = dalloc<q4>( c, n ) ; c->d_tb_hi = dalloc<q4>( c, n ) ;
		c->d_tp_lo = dalloc<q4>( c, n ) ; c->d_tp_hi = dalloc<q4>( c, n ) ;
		c->n_things_dev = n ;
	}
	std::vector<ThingTrav> trav( n ) ; std::vector<ThingShade> shade( n ) ; std::vector<q4> lo( n ), hi( n ), bs( n ) ;
	for ( uint32_t k = 0 ; k<n ; k++ ) {
		const ThingHost& th = c->things[k] ;
		const Mesh& m = c->meshes[th.mesh] ;
		ThingTrav& t = trav[k] ; ThingShade& s = shade[k] ;
		memset( &t, 0, sizeof( t ) ) ; memset( &s, 0, sizeof( s ) ) ;
		for ( int j = 0 ; j<12 ; j++ ) s.xf[j] = double( th.xf[j] ) ;
		s.albedo[0] = th.optics.albedo[0] ; s.albedo[1] = th.optics.albedo[1] ; s.albedo[2] = th.optics.albedo[2] ;
		s.fuzz = th.optics.fuzz ; s.index = th.optics.index ; s.type = th.optics.type ;
		if ( m.analytic ) {
			t.kind = 0 ; s.kind = 0 ;
			t.inv[0] = double( th.xf[3] ) ; t.inv[1] = double( th.xf[7] ) ; t.inv[2] = double( th.xf[11] ) ; t.inv[3] = double( th.xf[0] ) ;
			lo[k] = { 0, 0, 0, 0 } ; hi[k] = { 0, 0, 0, 0 } ;
			{	// the padded sphere itself: a cheap float pre-test in front of the double-precision roots
				const double unit[4] = { 0., 0., 0., 1. } ;
				const float sx[12] = { std::fabs( th.xf[0] ), 0, 0, th.xf[3], 0, std::fabs( th.xf[0] ), 0, th.xf[7], 0, 0, std::fabs( th.xf[0] ), th.xf[11] } ;
				world_bsphere( sx, unit, &bs[k].x ) ;
			}
		} else {
			world_bsphere( th.xf, m.bsphere, &bs[k].x ) ;
			t.kind = 1 ; s.kind = 1 ;
			t.diag = s.diag = ( th.xf[1] == 0.f && th.xf[2] == 0.f && th.xf[4] == 0.f && th.xf[6] == 0.f && th.xf[8] == 0.f && th.xf[9] == 0.f ) ? 1 : 0 ;
			affine_inverse( th.xf, t.inv ) ;
			t.nodes = m.bvh.nodes ; t.tris = m.tris ; t.n_tris = m.nt ;
			s.vces = m.vces ; s.ices = m.ices ;
			lo[k] = m.bvh.root_lo ; hi[k] = m.bvh.root_hi ;
		}
	}
	// (the device time of a top-level build or refit starts here: the copies of the thing records, their world bounds,
	// the hierarchy -- not the host loop above, during which the device idles)
	CK( cudaEventRecord( c->ev0, c->stream ) ) ;
	CK( cudaMemcpyAsync( c->d_trav, trav.data(), sizeof( ThingTrav )*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaMemcpyAsync( c->d_shade, shade.data(), sizeof( ThingShade )*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaMemcpyAsync( c->d_bsphere, bs.data(), sizeof( q4 )*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaMemcpyAsync( c->d_tb_lo, lo.data(), sizeof( q4 )*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaMemcpyAsync( c->d_tb_hi, hi.data(), sizeof( q4 )*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;   // the host vectors go out of scope
	if ( n ) {
		k_thing_bounds<<<( n+127 )/128, 128, 0, c->stream>>>( c->d_trav, c->d_shade, c->d_tb_lo, c->d_tb_hi, n, c->d_tp_lo, c->d_tp_hi ) ;
		c->launches += 1 ;
		CK( cudaGetLastError() ) ;
	}
}

SceneDev scene_dev( const rtx_ctx* c ) {
	SceneDev S ;
	S.tlas_nodes = c->tlas.nodes ; S.tlas_order = c->tlas.order ;
	S.trav = c->d_trav ; S.shade = c->d_shade ; S.bsphere = c->d_bsphere ; S.n_things = c->n_things_dev ;
	S.variant = 0 ; S.fault = c->d_fault ;
	// arena: the lowest node / triangle array (rtx_qpool.cuh addresses them as 16-byte offsets in 32 bits)
	uintptr_t lo = ~uintptr_t( 0 ), hi = 0 ;
	auto span = [&]( const void* ptr, size_t bytes ) { if ( ptr ) { lo = std::min( lo, uintptr_t( ptr ) ) ; hi = std::max( hi, uintptr_t( ptr )+bytes ) ; } } ;
	span( c->tlas.nodes, size_t( c->tlas.cap_nodes )*RTX_NODE_RECS*sizeof( q4 ) ) ;
	for ( const Mesh& m : c->meshes ) { span( m.bvh.nodes, size_t( m.bvh.cap_nodes )*RTX_NODE_RECS*sizeof( q4 ) ) ; span( m.tris, size_t( m.nt )*RTX_TRI_RECS*sizeof( q4 ) ) ; }
	S.arena = reinterpret_cast<const q4*>( lo == ~uintptr_t( 0 ) ? uintptr_t( 0 ) : lo-16 ) ;   // (-16: offset 0 means "none")
	if ( getenv( "RTX_VERBOSE" ) ) fprintf( stderr, "rtx: hierarchy arena %.1f MB at %p\n", hi>lo ? double( hi-lo )/1048576. : 0., ( const void* ) lo ) ;
	if ( hi>lo && ( hi-lo )/16>=0xfffffff0ull )
		S.arena = nullptr ;   // (more than 64 GB apart: do_render falls back to k_render)
	return S ;
}

// behind a tracing launch (stream already synchronised by the caller when `synced`): a traversal
// that ran out of stack cannot have finished its ray correctly -- fail the call instead of
// returning a plausible wrong frame
void fault_fetch( rtx_ctx* c ) {
	CK( cudaMemcpyAsync( c->h_fault, c->d_fault, sizeof( uint32_t ), cudaMemcpyDeviceToHost, c->stream ) ) ;
}
void fault_check( rtx_ctx* c, const char* what ) {
	if ( *c->h_fault ) {
		*c->h_fault = 0 ;
		c->stack_faults++ ;
		cudaMemsetAsync( c->d_fault, 0, sizeof( uint32_t ), c->stream ) ;
		throw std::runtime_error( std::string( what )+": traversal stack overflow (hierarchy deeper than the 96-entry stack; result discarded)" ) ;
	}
}

void free_frame( rtx_ctx* c ) {
	const size_t np = size_t( c->w )*c->h ;
	dfree( c, c->d_accum, 4*np ) ; dfree( c, c->d_raw, 3*np ) ; dfree( c, c->d_rpp, np ) ; dfree( c, c->d_image, np ) ;
	dfree( c, c->d_hit_id, np ) ; dfree( c, c->d_hit_t, np ) ; dfree( c, c->d_normals, 3*np ) ; dfree( c, c->d_albedos, 3*np ) ;
	dfree( c, c->d_guide_acc, 6*np ) ; c->guides_valid = false ;
}

CameraDev camera_dev( const rtx_camera& k ) {
	CameraDev d ;
	d.eye = mk3( k.eye[0], k.eye[1], k.eye[2] ) ; d.u = mk3( k.u[0], k.u[1], k.u[2] ) ; d.v = mk3( k.v[0], k.v[1], k.v[2] ) ;
	d.hvec = mk3( k.hvec[0], k.hvec[1], k.hvec[2] ) ; d.wvec = mk3( k.wvec[0], k.wvec[1], k.wvec[2] ) ; d.dvec = mk3( k.dvec[0], k.dvec[1], k.dvec[2] ) ;
	d.aperture = k.aperture ;
	return d ;
}

static_assert( int( RTX_VARIANT_RTOW ) == int( RTX_SEM_RTOW ) && int( RTX_VARIANT_RTWO_I ) == int( RTX_SEM_RTWO_I ) && int( RTX_VARIANT_RTWO_R ) == int( RTX_SEM_RTWO_R ), "variant codes" ) ;
static_assert( sizeof( rtx_params ) == 128, "rtx_params layout (tests/test_abi_and_host.py)" ) ;

FrameArgs frame_args( rtx_ctx* c, const rtx_params* p ) {
	if ( ! c->built ) throw std::runtime_error( "rtx: acceleration structure not built (call rtx_accel_build)" ) ;
	if ( p->image_w != c->w || p->image_h != c->h || c->w == 0 ) throw std::runtime_error( "rtx: image size differs from the last rtx_resize" ) ;
	if ( p->image_w<2 || p->image_h<2 ) throw std::runtime_error( "rtx: image must be at least 2x2" ) ;
	FrameArgs a ;
	memset( &a, 0, sizeof( a ) ) ;
	if ( p->variant>RTX_VARIANT_RTWO_R ) throw std::runtime_error( "rtx: unknown variant" ) ;
	a.S = scene_dev( c ) ; a.S.variant = p->variant ; a.cam = camera_dev( p->camera ) ;
	a.w = p->image_w ; a.h = p->image_h ; a.spp = p->spp ; a.depth = p->depth ; a.seed = p->seed ;
	a.sample0 = p->sample0 ; a.sample_stride = p->sample_stride ? p->sample_stride : 1u ; a.accumulate = p->accumulate ;
	a.accum = c->d_accum ; a.hit_id = c->d_hit_id ; a.hit_t = c->d_hit_t ;
	a.guides = p->guides ? 1u : 0u ;
	if ( a.guides && ! c->d_guide_acc ) {
		c->d_guide_acc = dalloc<long long>( c, 6*size_t( c->w )*c->h ) ;
		CK( cudaMemsetAsync( c->d_guide_acc, 0, sizeof( long long )*6*size_t( c->w )*c->h, c->stream ) ) ;
	}
	a.guide_acc = c->d_guide_acc ;
	return a ;
}

void do_resolve( rtx_ctx* c, uint64_t total_spp ) {
	const uint32_t np = c->w*c->h ;
	k_resolve<<<( np+255 )/256, 256, 0, c->stream>>>( c->d_accum, np, total_spp ? total_spp : 1, c->d_raw, c->d_rpp ) ;
	c->launches += 1 ;
	if ( c->guides_valid ) {
		k_resolve_guides<<<( np+255 )/256, 256, 0, c->stream>>>( c->d_guide_acc, np, total_spp ? total_spp : 1, c->d_normals, c->d_albedos ) ;
		c->launches += 1 ;
	}
	CK( cudaGetLastError() ) ;
}

__global__ void k_sum_segments( const uint64_t* accum, uint32_t npix, unsigned long long* out ) ;

// the asynchronous part of a frame on one context: buffers cleared, the path-tracing kernel
// launched on the context's stream, its fault word fetched behind it
void launch_render( rtx_ctx* c, const rtx_params* p ) {
	CK( cudaSetDevice( c->device ) ) ;
	FrameArgs a = frame_args( c, p ) ;
	unit_plan( a ) ;
	c->guides_valid = a.guides != 0 ;
	if ( a.depth>255u ) throw std::runtime_error( "rtx: depth above 255 is not supported" ) ;
	if ( a.spp == 0 ) throw std::runtime_error( "rtx: spp must be at least 1" ) ;
	const uint32_t n_tiles = ( ( a.w+RTX_TILE_W-1u )>>RTX_TILE_WLOG )*( ( a.h+RTX_TILE_H-1u )>>RTX_TILE_HLOG )*( a.chunks_full+a.chunks_taper ) ;   // work units
	CK( cudaMemsetAsync( c->d_tile_counter, 0, sizeof( uint32_t )*( 2+2*256 ), c->stream ) ) ;   // unit counter + per-SM words
	if ( ! a.accumulate ) {   // paths add into the buffers with atomics: start from zero (optx/camera_i.cu:52)
		CK( cudaMemsetAsync( c->d_accum, 0, sizeof( uint64_t )*4*size_t( a.w )*a.h, c->stream ) ) ;
		if ( a.guides ) CK( cudaMemsetAsync( c->d_guide_acc, 0, sizeof( long long )*6*size_t( a.w )*a.h, c->stream ) ) ;
	}
	if ( c->kernel == 1 && a.S.arena == nullptr && a.S.n_things && getenv( "RTX_VERBOSE" ) ) fprintf( stderr, "rtx: hierarchy arrays too far apart for k_render_q, using k_render\n" ) ;
	if ( c->kernel == 1 && ( a.S.arena != nullptr || a.S.n_things == 0 ) ) {
		// one warp of the pool kernel holds RTX_QR paths: fewer warps for a small frame
		const uint32_t grid = uint32_t( std::min<uint64_t>( c->q_grid, std::max<uint64_t>( 1, ( uint64_t( a.w )*a.h*a.spp+RTX_QR-1 )/RTX_QR ) ) ) ;
		if ( a.guides ) k_render_q<true><<<grid, 32, RTX_Q_SMEM_BYTES, c->stream>>>( a, c->d_tile_counter, c->d_qovf, c->d_qcold ) ;
		else            k_render_q<false><<<grid, 32, RTX_Q_SMEM_BYTES, c->stream>>>( a, c->d_tile_counter, c->d_qovf, c->d_qcold ) ;
	}
	else if ( a.guides ) k_render<true><<<min( c->render_grid, n_tiles ), 32, 0, c->stream>>>( a, c->d_tile_counter, c->d_ovf ) ;
	else            k_render<false><<<min( c->render_grid, n_tiles ), 32, 0, c->stream>>>( a, c->d_tile_counter, c->d_ovf ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	fault_fetch( c ) ;
}

void do_render( rtx_ctx* c, const rtx_params* p, bool resolve ) {
	const uint32_t G = 1u+uint32_t( c->replicas.size() ) ;
	CK( cudaSetDevice( c->device ) ) ;
	CK( cudaEventRecord( c->ev0, c->stream ) ) ;
	if ( G == 1u ) {
		launch_render( c, p ) ;
		CK( cudaEventRecord( c->ev1, c->stream ) ) ;
		if ( resolve )
			do_resolve( c, p->spp ) ;
		CK( cudaEventRecord( c->ev2, c->stream ) ) ;
		CK( cudaStreamSynchronize( c->stream ) ) ;
		CK( cudaEventElapsedTime( &c->ms_render, c->ev0, c->ev1 ) ) ;
		CK( cudaEventElapsedTime( &c->ms_resolve, c->ev1, c->ev2 ) ) ;
		CK( cudaEventElapsedTime( &c->ms_frame, c->ev0, c->ev2 ) ) ;
		c->paths = uint64_t( p->image_w )*p->image_h*p->spp ;
		fault_check( c, "rtx_render" ) ;
		return ;
	}
	// samples split over the devices: device r traces the global samples sample0 + (r + k G) stride
	// (BASELINE.json north_star).  The sums are integers, so the frame does not depend on G.
	const uint32_t stride = p->sample_stride ? p->sample_stride : 1u ;
	PeerBufs pb ;
	memset( &pb, 0, sizeof( pb ) ) ;
	pb.n = int( G ) ;
	pb.accum[0] = c->d_accum ;
	for ( uint32_t r = 0 ; r<G ; r++ ) {
		rtx_ctx* d = r ? c->replicas[r-1] : c ;
		rtx_params q = *p ;
		q.spp = p->spp/G+( r<p->spp%G ? 1u : 0u ) ;
		q.sample0 = p->sample0+r*stride ;
		q.sample_stride = stride*G ;
		if ( r ) q.accumulate = 0 ;   // (the root's buffer carries what was accumulated before; replicas hold this call's share only)
		if ( q.spp ) launch_render( d, &q ) ;
		else if ( r ) {
			CK( cudaSetDevice( d->device ) ) ;
			CK( cudaMemsetAsync( d->d_accum, 0, sizeof( uint64_t )*4*size_t( d->w )*d->h, d->stream ) ) ;
			if ( p->guides && d->d_guide_acc ) CK( cudaMemsetAsync( d->d_guide_acc, 0, sizeof( long long )*6*size_t( d->w )*d->h, d->stream ) ) ;
		}
		if ( r == 0 ) { CK( cudaSetDevice( c->device ) ) ; CK( cudaEventRecord( c->ev_in, c->stream ) ) ; }   // behind the root's own kernel
		if ( r ) {
			CK( cudaSetDevice( d->device ) ) ;
			const size_t np = size_t( d->w )*d->h ;
			if ( c->peer_ok[r-1] ) {
				pb.accum[r] = d->d_accum ; pb.guide[r] = p->guides ? d->d_guide_acc : nullptr ;
			} else {
				// no peer access between the two devices: the replica's sums travel through a copy on the root
				CK( cudaMemcpyPeerAsync( c->stage_accum[r-1], c->device, d->d_accum, d->device, sizeof( uint64_t )*4*np, d->stream ) ) ;
				pb.accum[r] = c->stage_accum[r-1] ;
				if ( p->guides ) {
					if ( ! c->stage_guide[r-1] ) { CK( cudaSetDevice( c->device ) ) ; c->stage_guide[r-1] = dalloc<long long>( c, 6*np ) ; CK( cudaSetDevice( d->device ) ) ; }
					CK( cudaMemcpyPeerAsync( c->stage_guide[r-1], c->device, d->d_guide_acc, d->device, sizeof( long long )*6*np, d->stream ) ) ;
					pb.guide[r] = c->stage_guide[r-1] ;
				}
			}
			CK( cudaEventRecord( d->ev_done, d->stream ) ) ;
		}
	}
	CK( cudaSetDevice( c->device ) ) ;
	pb.guide[0] = p->guides ? c->d_guide_acc : nullptr ;
	c->guides_valid = p->guides != 0 ;
	for ( rtx_ctx* d : c->replicas ) CK( cudaStreamWaitEvent( c->stream, d->ev_done, 0 ) ) ;
	if ( getenv( "RTX_DEBUG_MULTI" ) ) {
		// development aid: the segment sums of every device's buffer before the reduce
		for ( uint32_t r = 0 ; r<G ; r++ ) {
			rtx_ctx* d = r ? c->replicas[r-1] : c ;
			CK( cudaSetDevice( d->device ) ) ;
			CK( cudaStreamSynchronize( d->stream ) ) ;
			CK( cudaMemsetAsync( d->d_counter, 0, sizeof( unsigned long long ), d->stream ) ) ;
			k_sum_segments<<<64, 256, 0, d->stream>>>( d->d_accum, d->w*d->h, d->d_counter ) ;
			unsigned long long sg = 0 ;
			CK( cudaMemcpyAsync( &sg, d->d_counter, sizeof( sg ), cudaMemcpyDeviceToHost, d->stream ) ) ;
			CK( cudaStreamSynchronize( d->stream ) ) ;
			fprintf( stderr, "rtx multi: device slot %u (device %d): %llu segments in %p, pb.accum %p; %ux%u, grid %u, %u things, built %d, fault %u\n", r, d->device, sg, ( void* ) d->d_accum, ( const void* ) pb.accum[r],
				d->w, d->h, d->render_grid, d->n_things_dev, int( d->built ), *d->h_fault ) ;
		}
		CK( cudaSetDevice( c->device ) ) ;
	}
	// one kernel on the root: sum of the devices' buffers read over NVLink peer memory, written back as
	// the root's accumulation buffer, and (rtx_render) the mean + clamp of optx/camera_i.cu:105 behind it
	const uint32_t np = c->w*c->h ;
	CK( cudaEventRecord( c->ev2, c->stream ) ) ;
	k_reduce_resolve<<<( np+255 )/256, 256, 0, c->stream>>>( pb, c->d_accum, np, p->spp ? p->spp : 1, resolve ? 1 : 0, c->d_raw, c->d_rpp,
		p->guides ? c->d_guide_acc : nullptr, c->d_normals, c->d_albedos ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	CK( cudaEventRecord( c->ev1, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	for ( rtx_ctx* d : c->replicas ) { CK( cudaSetDevice( d->device ) ) ; CK( cudaStreamSynchronize( d->stream ) ) ; }
	CK( cudaSetDevice( c->device ) ) ;
	CK( cudaEventElapsedTime( &c->ms_render, c->ev0, c->ev_in ) ) ;
	CK( cudaEventElapsedTime( &c->ms_resolve, c->ev2, c->ev1 ) ) ;
	CK( cudaEventElapsedTime( &c->ms_frame, c->ev0, c->ev1 ) ) ;
	c->paths = uint64_t( p->image_w )*p->image_h*p->spp ;
	fault_check( c, "rtx_render" ) ;
	for ( rtx_ctx* d : c->replicas ) {
		try { fault_check( d, "rtx_render" ) ; }
		catch ( ... ) { CK( cudaSetDevice( c->device ) ) ; throw ; }
	}
}

struct BufInfo { void* ptr ; size_t bytes ; } ;
BufInfo buffer_of( rtx_ctx* c, int buffer ) {
	const size_t np = size_t( c->w )*c->h ;
	switch ( buffer ) {
		case RTX_BUF_ACCUM:   return { c->d_accum,   np*4*sizeof( uint64_t ) } ;
		case RTX_BUF_RAWRGB:  return { c->d_raw,     np*3*sizeof( float ) } ;
		case RTX_BUF_RPP:     return { c->d_rpp,     np*sizeof( uint32_t ) } ;
		case RTX_BUF_IMAGE:   return { c->d_image,   np*4 } ;
		case RTX_BUF_HIT_ID:  return { c->d_hit_id,  np*sizeof( int64_t ) } ;
		case RTX_BUF_HIT_T:   return { c->d_hit_t,   np*sizeof( float ) } ;
		case RTX_BUF_NORMALS: return { c->d_normals, np*3*sizeof( float ) } ;
		case RTX_BUF_ALBEDOS: return { c->d_albedos, np*3*sizeof( float ) } ;
		case RTX_BUF_PICK_ID: return { c->d_pick,    sizeof( uint32_t ) } ;
		case RTX_BUF_GUIDE_ACC: return { c->d_guide_acc, np*6*sizeof( long long ) } ;
	}
	throw std::runtime_error( "rtx: unknown buffer id" ) ;
}

__global__ void __launch_bounds__( 256 ) k_sum_segments( const uint64_t* accum, uint32_t npix, unsigned long long* out ) {
	unsigned long long s = 0 ;
	for ( uint32_t p = blockIdx.x*blockDim.x+threadIdx.x ; p<npix ; p += gridDim.x*blockDim.x )
		s += accum[4*size_t( p )+3] ;
	for ( int o = 16 ; o>0 ; o >>= 1 )
		s += __shfl_xor_sync( 0xffffffffu, s, o ) ;
	if ( ( threadIdx.x&31 ) == 0 )
		atomicAdd( out, s ) ;
}

} // namespace

// multi-GPU: a scene call made on the root is repeated on every replica (same arguments, same ids)
#define RTX_FANOUT( c, call ) do { for ( rtx_ctx* r_ : ( c )->replicas ) { if ( call ) { ( c )->err = "replica on device "+std::to_string( r_->device )+": "+r_->err ; cudaSetDevice( ( c )->device ) ; return 1 ; } } if ( ! ( c )->replicas.empty() ) cudaSetDevice( ( c )->device ) ; } while ( 0 )
#define RTX_TRY( c ) try {
#define RTX_END( c ) return 0 ; } catch ( const std::exception& e_ ) { if ( c ) ( c )->err = e_.what() ; else g_init_error = e_.what() ; cudaGetLastError() ; return 1 ; }

extern "C" {

int rtx_init( int device, rtx_ctx** out ) {
	rtx_ctx* c = nullptr ;
	try {
		int n = 0 ;
		if ( cudaGetDeviceCount( &n ) != cudaSuccess || n<1 )
			throw std::runtime_error( "rtx_init: no CUDA device (this library has no CPU path)" ) ;
		if ( device<0 || device>=n )
			throw std::runtime_error( "rtx_init: device index out of range" ) ;
		CK( cudaSetDevice( device ) ) ;
		CK( cudaFree( 0 ) ) ;
		cudaDeviceProp prop ;
		CK( cudaGetDeviceProperties( &prop, device ) ) ;
		if ( prop.major<10 )
			throw std::runtime_error( "rtx_init: needs an sm_100a (Blackwell B200) device" ) ;
		c = new rtx_ctx ;
		c->device = device ;
		CK( cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) ) ;
		CK( cudaEventCreate( &c->ev0 ) ) ; CK( cudaEventCreate( &c->ev1 ) ) ; CK( cudaEventCreate( &c->ev2 ) ) ; CK( cudaEventCreate( &c->ev_in ) ) ;
		for ( cudaEvent_t& e : c->stage_ev ) CK( cudaEventCreate( &e ) ) ;
		{
			cudaMemPoolProps pp ;
			memset( &pp, 0, sizeof( pp ) ) ;
			pp.allocType = cudaMemAllocationTypePinned ;
			pp.location.type = cudaMemLocationTypeDevice ; pp.location.id = device ;
			CK( cudaMemPoolCreate( &c->pool, &pp ) ) ;
			uint64_t keep = ~0ull ;
			CK( cudaMemPoolSetAttribute( c->pool, cudaMemPoolAttrReleaseThreshold, &keep ) ) ;
			// reserve the workspace of a 1 M-triangle build now: growing the pool maps new device
			// memory, which costs more than the build itself
			void* warm = nullptr ;
			CK( cudaMallocFromPoolAsync( &warm, size_t( 256 )<<20, c->pool, c->stream ) ) ;
			CK( cudaFreeAsync( warm, c->stream ) ) ;
			CK( cudaStreamSynchronize( c->stream ) ) ;
		}
		c->d_pick = dalloc<uint32_t>( c, 1 ) ;
		c->d_fault = dalloc<uint32_t>( c, 1 ) ;
		CK( cudaMemsetAsync( c->d_fault, 0, sizeof( uint32_t ), c->stream ) ) ;
		CK( cudaMallocHost( reinterpret_cast<void**>( &c->h_fault ), sizeof( uint32_t ) ) ) ;
		*c->h_fault = 0 ;
		c->d_counter = dalloc<unsigned long long>( c, 1 ) ;
		c->d_tile_counter = dalloc<uint32_t>( c, 2+2*256 ) ;
		{	// CUDA loads kernels lazily on first launch; do it here so that build and frame
			// timings measure the kernels, not the loader
			cudaFuncAttributes fa ;
			const void* kernels[] = { ( const void* ) k_render<false>, ( const void* ) k_render<true>, ( const void* ) k_primary_hits, ( const void* ) k_trace_rays, ( const void* ) k_pick,
				( const void* ) k_resolve, ( const void* ) k_resolve_guides, ( const void* ) k_postproc, ( const void* ) k_sum_segments, ( const void* ) k_tri_bounds, ( const void* ) k_thing_bounds,
				( const void* ) k_bounds_init, ( const void* ) k_bounds_reduce, ( const void* ) k_morton, ( const void* ) k_radix_hist, ( const void* ) k_radix_scan,
				( const void* ) k_radix_scatter, ( const void* ) k_karras, ( const void* ) k_refit, ( const void* ) k_wide_level, ( const void* ) k_wide_all, ( const void* ) k_scan_chunks, ( const void* ) k_scan_add, ( const void* ) k_pack_tris,
				( const void* ) k_radix_hist8, ( const void* ) k_radix_bases, ( const void* ) k_radix_onesweep } ;
			for ( const void* k : kernels ) CK( cudaFuncGetAttributes( &fa, k ) ) ;
			CK( cudaFuncSetAttribute( k_radix_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, int( RTX_OS_SMEM_BYTES ) ) ) ;
		}
		// the render kernel is persistent: one warp per CTA, as many CTAs as fit
		// shared memory holds the ray slots, L1 caches the BVH: the carveout decides how many
		// render warps an SM hosts and how much L1 is left (tunable: RTX_CARVEOUT, percent)
		int carve = RTX_DEFAULT_CARVEOUT ;
		if ( const char* e = getenv( "RTX_CARVEOUT" ) ) carve = atoi( e ) ;
		if ( carve>0 && carve<=100 ) {
			CK( cudaFuncSetAttribute( k_render<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve ) ) ;
			CK( cudaFuncSetAttribute( k_render<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve ) ) ;
		}
		int per_sm = 0 ;
		CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, k_render<false>, 32, 0 ) ) ;
		if ( per_sm<1 ) per_sm = 1 ;
		if ( const char* e = getenv( "RTX_CTAS_PER_SM" ) ) { const int v = atoi( e ) ; if ( v>=1 && v<per_sm ) per_sm = v ; }   // (tuning)
		if ( getenv( "RTX_VERBOSE" ) ) fprintf( stderr, "rtx_init: %d render warps per SM, carveout %d %%\n", per_sm, carve ) ;
		c->render_grid = uint32_t( per_sm )*uint32_t( prop.multiProcessorCount ) ;
		c->d_ovf = dalloc<int32_t>( c, size_t( c->render_grid )*RTX_POOL_R*RTX_POOL_OVF*2 ) ;
		{	// the collapse into wide nodes runs all its levels in one cooperative launch
			int coop = 0, wide_per_sm = 0 ;
			CK( cudaDeviceGetAttribute( &coop, cudaDevAttrCooperativeLaunch, device ) ) ;
			CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &wide_per_sm, k_wide_all, 128, 0 ) ) ;
			if ( coop && wide_per_sm>0 && ! getenv( "RTX_NO_COOP" ) ) c->wide_grid = uint32_t( std::min( wide_per_sm, 8 ) )*uint32_t( prop.multiProcessorCount ) ;
		}
		{	// the compacting-pool kernel: ray slots in dynamic shared memory
			if ( const char* e = getenv( "RTX_KERNEL" ) ) c->kernel = ( e[0] == 'q' || e[0] == '1' ) ? 1 : 0 ;
			int qcarve = RTX_Q_DEFAULT_CARVEOUT ;
			if ( const char* e = getenv( "RTX_Q_CARVEOUT" ) ) qcarve = atoi( e ) ;
			if ( qcarve>0 && qcarve<=100 ) {
				CK( cudaFuncSetAttribute( k_render_q<false>, cudaFuncAttributePreferredSharedMemoryCarveout, qcarve ) ) ;
				CK( cudaFuncSetAttribute( k_render_q<true>, cudaFuncAttributePreferredSharedMemoryCarveout, qcarve ) ) ;
			}
			int q_per_sm = 0 ;
			CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &q_per_sm, k_render_q<false>, 32, RTX_Q_SMEM_BYTES ) ) ;
			if ( q_per_sm<1 ) q_per_sm = 1 ;
			if ( const char* e = getenv( "RTX_Q_CTAS_PER_SM" ) ) { const int v = atoi( e ) ; if ( v>=1 && v<q_per_sm ) q_per_sm = v ; }   // (tuning)
			if ( getenv( "RTX_VERBOSE" ) ) fprintf( stderr, "rtx_init: kernel %s; pool kernel: %d warps per SM x %d ray slots (%u bytes of shared memory each), carveout %d %%\n",
				c->kernel ? "k_render_q" : "k_render", q_per_sm, int( RTX_QR ), unsigned( RTX_Q_SMEM_BYTES ), qcarve ) ;
			c->q_grid = uint32_t( q_per_sm )*uint32_t( prop.multiProcessorCount ) ;
			c->d_qcold = dalloc<q4>( c, size_t( c->q_grid )*RTX_QR*4 ) ;
			c->d_qovf = dalloc<int32_t>( c, size_t( c->q_grid )*RTX_QR*RTX_QOVF*2 ) ;
		}
		*out = c ;
		return 0 ;
	} catch ( const std::exception& e ) {
		g_init_error = e.what() ;
		delete c ;
		cudaGetLastError() ;
		return 1 ;
	}
}

int rtx_init_multi( int n_devices, const int* device_ids, rtx_ctx** out ) {
	if ( n_devices<1 || n_devices>RTX_MAX_DEVICES || ! device_ids ) { g_init_error = "rtx_init_multi: 1 to 8 devices" ; return 1 ; }
	rtx_ctx* root = nullptr ;
	if ( rtx_init( device_ids[0], &root ) )
		return 1 ;
	try {
		for ( int k = 1 ; k<n_devices ; k++ ) {
			rtx_ctx* r = nullptr ;
			if ( rtx_init( device_ids[k], &r ) )
				throw std::runtime_error( g_init_error ) ;
			r->kernel = root->kernel ;
			CK( cudaEventCreateWithFlags( &r->ev_done, cudaEventDisableTiming ) ) ;
			root->replicas.push_back( r ) ;
			// the root reads the replica's buffers directly where the devices are peers (NVLink / NVSwitch
			// on a B200 board; trivially so for a second context on the same device)
			int ok = r->device == root->device ? 1 : 0 ;
			if ( ! ok ) {
				CK( cudaDeviceCanAccessPeer( &ok, root->device, r->device ) ) ;
				if ( ok ) {
					CK( cudaSetDevice( root->device ) ) ;
					const cudaError_t e = cudaDeviceEnablePeerAccess( r->device, 0 ) ;
					if ( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled ) ok = 0 ;
					cudaGetLastError() ;
				}
			}
			root->peer_ok.push_back( char( ok ) ) ;
			root->stage_accum.push_back( nullptr ) ; root->stage_guide.push_back( nullptr ) ;
			if ( getenv( "RTX_VERBOSE" ) ) fprintf( stderr, "rtx_init_multi: replica %d on device %d, %s\n", k, r->device, ok ? "peer access" : "staged copies" ) ;
		}
		CK( cudaSetDevice( root->device ) ) ;
	} catch ( const std::exception& e ) {
		g_init_error = e.what() ;
		rtx_shutdown( root ) ;
		cudaGetLastError() ;
		return 1 ;
	}
	*out = root ;
	return 0 ;
}

int rtx_device_count( const rtx_ctx* c ) { return c ? 1+int( c->replicas.size() ) : 0 ; }

void rtx_shutdown( rtx_ctx* c ) {
	if ( ! c ) return ;
	for ( rtx_ctx* r : c->replicas ) rtx_shutdown( r ) ;
	c->replicas.clear() ;
	cudaSetDevice( c->device ) ;
	for ( size_t k = 0 ; k<c->stage_accum.size() ; k++ ) {
		const size_t np = size_t( c->w )*c->h ;
		dfree( c, c->stage_accum[k], 4*np ) ; dfree( c, c->stage_guide[k], 6*np ) ;
	}
	if ( c->ev_done ) cudaEventDestroy( c->ev_done ) ;
	cudaSetDevice( c->device ) ;
	cudaStreamSynchronize( c->stream ) ;
	for ( Mesh& m : c->meshes ) {
		dfree( c, m.vces, 3*size_t( m.nv ) ) ; dfree( c, m.ices, 3*size_t( m.nt ) ) ; tfree( c, m.tris, RTX_TRI_RECS*size_t( m.nt ) ) ;
		lbvh_free( c, m.bvh ) ;
	}
	lbvh_free( c, c->tlas ) ;
	dfree( c, c->d_trav, c->n_things_dev ) ; dfree( c, c->d_shade, c->n_things_dev ) ; dfree( c, c->d_bsphere, c->n_things_dev ) ;
	dfree( c, c->d_tb_lo, c->n_things_dev ) ; dfree( c, c->d_tb_hi, c->n_things_dev ) ;
	dfree( c, c->d_tp_lo, c->n_things_dev ) ; dfree( c, c->d_tp_hi, c->n_things_dev ) ;
	free_frame( c ) ;
	dfree( c, c->d_pick, 1 ) ; dfree( c, c->d_fault, 1 ) ; if ( c->h_fault ) cudaFreeHost( c->h_fault ) ;
	dfree( c, c->d_counter, 1 ) ; dfree( c, c->d_tile_counter, 2+2*256 ) ;
	dfree( c, c->d_ovf, size_t( c->render_grid )*RTX_POOL_R*RTX_POOL_OVF*2 ) ;
	dfree( c, c->d_qcold, size_t( c->q_grid )*RTX_QR*4 ) ; dfree( c, c->d_qovf, size_t( c->q_grid )*RTX_QR*RTX_QOVF*2 ) ;
	cudaEventDestroy( c->ev0 ) ; cudaEventDestroy( c->ev1 ) ; if ( c->ev2 ) cudaEventDestroy( c->ev2 ) ; if ( c->ev_in ) cudaEventDestroy( c->ev_in ) ;
	for ( cudaEvent_t e : c->stage_ev ) if ( e ) cudaEventDestroy( e ) ;
	cudaStreamSynchronize( c->stream ) ;   // the stream-ordered frees above
	if ( c->pool ) cudaMemPoolDestroy( c->pool ) ;
	cudaStreamDestroy( c->stream ) ;
	delete c ;
}

const char* rtx_last_error( const rtx_ctx* c ) { return c ? c->err.c_str() : g_init_error.c_str() ; }

int rtx_mesh_create( rtx_ctx* c, const float* xyz, uint32_t nv, const uint32_t* idx, uint32_t nt, uint32_t* mesh_id ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( ! xyz || ! idx || nv == 0 || nt == 0 ) throw std::runtime_error( "rtx_mesh_create: empty mesh" ) ;
	if ( nt>=( 1u<<28 ) ) throw std::runtime_error( "rtx_mesh_create: 2^28 or more triangles in one mesh (leaf references hold first<<3|count-1 in 31 bits)" ) ;
	for ( size_t k = 0 ; k<3*size_t( nt ) ; k++ )
		if ( idx[k]>=nv ) throw std::runtime_error( "rtx_mesh_create: index out of bounds" ) ;   // optx/object.cxx:62-67
	Mesh m ;
	m.nv = nv ; m.nt = nt ;
	mesh_bsphere( xyz, nv, m.bsphere ) ;
	q4* plo = nullptr ; q4* phi = nullptr ;
	try {
		m.vces = dalloc<float>( c, 3*size_t( nv ) ) ; m.ices = dalloc<uint32_t>( c, 3*size_t( nt ) ) ; m.tris = talloc<q4>( c, RTX_TRI_RECS*size_t( nt ) ) ;   // (triangles beside the node arrays: k_render_q addresses both as offsets from one base)
		CK( cudaMemcpyAsync( m.vces, xyz, sizeof( float )*3*nv, cudaMemcpyHostToDevice, c->stream ) ) ;
		CK( cudaMemcpyAsync( m.ices, idx, sizeof( uint32_t )*3*size_t( nt ), cudaMemcpyHostToDevice, c->stream ) ) ;
		plo = talloc<q4>( c, nt ) ; phi = talloc<q4>( c, nt ) ;
		CK( cudaEventRecord( c->ev0, c->stream ) ) ;
		k_tri_bounds<<<( nt+255 )/256, 256, 0, c->stream>>>( m.vces, m.ices, nt, plo, phi ) ;
		c->launches += 1 ;
		lbvh_build( c, m.bvh, plo, phi, nt, RTX_LEAF_MAX, false ) ;
		k_pack_tris<<<( nt+255 )/256, 256, 0, c->stream>>>( m.vces, m.ices, m.bvh.order, nt, m.tris ) ;
		c->launches += 1 ;
		CK( cudaGetLastError() ) ;
		CK( cudaEventRecord( c->ev1, c->stream ) ) ;
		CK( cudaStreamSynchronize( c->stream ) ) ;
	} catch ( ... ) {
		// a build that failed half way leaves nothing behind
		cudaStreamSynchronize( c->stream ) ;
		tfree( c, plo, nt ) ; tfree( c, phi, nt ) ;
		lbvh_free( c, m.bvh ) ;
		dfree( c, m.vces, 3*size_t( nv ) ) ; dfree( c, m.ices, 3*size_t( nt ) ) ; tfree( c, m.tris, RTX_TRI_RECS*size_t( nt ) ) ;
		throw ;
	}
	float ms = 0.f ;
	CK( cudaEventElapsedTime( &ms, c->ev0, c->ev1 ) ) ;
	c->ms_blas += ms ;
	stage_collect( c, c->ms_stage_blas, true ) ;
	tfree( c, plo, nt ) ; tfree( c, phi, nt ) ;
	*mesh_id = uint32_t( c->meshes.size() ) ;
	c->meshes.push_back( m ) ;
	RTX_FANOUT( c, [&]{ uint32_t id_ ; return rtx_mesh_create( r_, xyz, nv, idx, nt, &id_ ) ; }() ) ;
	RTX_END( c )
}

int rtx_sphere_create( rtx_ctx* c, uint32_t* mesh_id ) {
	RTX_TRY( c )
	Mesh m ;
	m.analytic = true ;
	*mesh_id = uint32_t( c->meshes.size() ) ;
	c->meshes.push_back( m ) ;
	RTX_FANOUT( c, [&]{ uint32_t id_ ; return rtx_sphere_create( r_, &id_ ) ; }() ) ;
	RTX_END( c )
}

int rtx_thing_add( rtx_ctx* c, uint32_t mesh_id, const rtx_optics* optics, uint32_t* thing_id ) {
	RTX_TRY( c )
	if ( mesh_id>=c->meshes.size() ) throw std::runtime_error( "rtx_thing_add: unknown mesh" ) ;
	if ( optics->type<0 || optics->type>2 ) throw std::runtime_error( "rtx_thing_add: unknown optics type" ) ;
	if ( c->things.size()>=( size_t( 1 )<<28 ) ) throw std::runtime_error( "rtx_thing_add: 2^28 things at most (leaf references)" ) ;
	ThingHost th ;
	th.mesh = mesh_id ; th.optics = *optics ;
	const float ident[12] = { 1, 0, 0, 0,  0, 1, 0, 0,  0, 0, 1, 0 } ;   // optx/scene.cxx:183-188
	memcpy( th.xf, ident, sizeof( ident ) ) ;
	*thing_id = uint32_t( c->things.size() ) ;
	c->things.push_back( th ) ;
	RTX_FANOUT( c, [&]{ uint32_t id_ ; return rtx_thing_add( r_, mesh_id, optics, &id_ ) ; }() ) ;
	RTX_END( c )
}

int rtx_thing_set_xf( rtx_ctx* c, uint32_t thing_id, const float xf[12] ) {
	RTX_TRY( c )
	if ( thing_id>=c->things.size() ) throw std::runtime_error( "rtx_thing_set_xf: unknown thing" ) ;
	memcpy( c->things[thing_id].xf, xf, sizeof( float )*12 ) ;
	RTX_FANOUT( c, rtx_thing_set_xf( r_, thing_id, xf ) ) ;
	RTX_END( c )
}

int rtx_thing_get_xf( rtx_ctx* c, uint32_t thing_id, float xf[12] ) {
	RTX_TRY( c )
	if ( thing_id>=c->things.size() ) throw std::runtime_error( "rtx_thing_get_xf: unknown thing" ) ;
	memcpy( xf, c->things[thing_id].xf, sizeof( float )*12 ) ;
	RTX_END( c )
}

int rtx_thing_set_optics( rtx_ctx* c, uint32_t thing_id, const rtx_optics* optics ) {
	RTX_TRY( c )
	if ( thing_id>=c->things.size() ) throw std::runtime_error( "rtx_thing_set_optics: unknown thing" ) ;
	if ( optics->type<0 || optics->type>2 ) throw std::runtime_error( "rtx_thing_set_optics: unknown optics type" ) ;
	c->things[thing_id].optics = *optics ;
	RTX_FANOUT( c, rtx_thing_set_optics( r_, thing_id, optics ) ) ;
	RTX_END( c )
}

int rtx_accel_build( rtx_ctx* c ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	upload_things( c ) ;   // (records ev0)
	const uint32_t n = c->n_things_dev ;
	if ( n ) lbvh_build( c, c->tlas, c->d_tp_lo, c->d_tp_hi, n, 1, true ) ;
	else     lbvh_free( c, c->tlas ) ;
	CK( cudaEventRecord( c->ev1, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	CK( cudaEventElapsedTime( &c->ms_tlas, c->ev0, c->ev1 ) ) ;
	if ( n ) stage_collect( c, c->ms_stage_tlas, false ) ;
	c->built = true ;
	RTX_FANOUT( c, rtx_accel_build( r_ ) ) ;
	RTX_END( c )
}

int rtx_accel_refit( rtx_ctx* c ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( ! c->built || c->things.size() != c->tlas.n )
		throw std::runtime_error( "rtx_accel_refit: thing count changed since rtx_accel_build" ) ;
	upload_things( c ) ;   // (records ev0)
	if ( c->tlas.n ) lbvh_refit( c, c->tlas, c->d_tp_lo, c->d_tp_hi, 1 ) ;
	CK( cudaEventRecord( c->ev1, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	CK( cudaEventElapsedTime( &c->ms_tlas, c->ev0, c->ev1 ) ) ;
	if ( c->tlas.n ) stage_collect( c, c->ms_stage_tlas, false ) ;
	RTX_FANOUT( c, rtx_accel_refit( r_ ) ) ;
	RTX_END( c )
}

int rtx_resize( rtx_ctx* c, uint32_t w, uint32_t h ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( w == 0 || h == 0 || uint64_t( w )*h>=( 1ull<<31 ) ) throw std::runtime_error( "rtx_resize: bad image size" ) ;
	for ( size_t k = 0 ; k<c->stage_accum.size() ; k++ ) { dfree( c, c->stage_accum[k], 4*size_t( c->w )*c->h ) ; dfree( c, c->stage_guide[k], 6*size_t( c->w )*c->h ) ; }
	free_frame( c ) ;
	c->w = w ; c->h = h ;
	const size_t np = size_t( w )*h ;
	c->d_accum = dalloc<uint64_t>( c, 4*np ) ; c->d_raw = dalloc<float>( c, 3*np ) ; c->d_rpp = dalloc<uint32_t>( c, np ) ;
	c->d_image = dalloc<uchar4>( c, np ) ; c->d_hit_id = dalloc<int64_t>( c, np ) ; c->d_hit_t = dalloc<float>( c, np ) ;
	c->d_normals = dalloc<float>( c, 3*np ) ; c->d_albedos = dalloc<float>( c, 3*np ) ;
	CK( cudaMemsetAsync( c->d_accum, 0, 4*np*sizeof( uint64_t ), c->stream ) ) ;
	CK( cudaMemsetAsync( c->d_raw, 0, 3*np*sizeof( float ), c->stream ) ) ;
	CK( cudaMemsetAsync( c->d_rpp, 0, np*sizeof( uint32_t ), c->stream ) ) ;
	CK( cudaMemsetAsync( c->d_normals, 0, 3*np*sizeof( float ), c->stream ) ) ;
	CK( cudaMemsetAsync( c->d_albedos, 0, 3*np*sizeof( float ), c->stream ) ) ;
	for ( size_t k = 0 ; k<c->replicas.size() ; k++ ) if ( ! c->peer_ok[k] ) c->stage_accum[k] = dalloc<uint64_t>( c, 4*np ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	RTX_FANOUT( c, rtx_resize( r_, w, h ) ) ;
	RTX_END( c )
}

int rtx_render( rtx_ctx* c, const rtx_params* p ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( p->accumulate ) throw std::runtime_error( "rtx_render: accumulate needs rtx_render_accumulate + rtx_resolve" ) ;
	do_render( c, p, true ) ;
	RTX_END( c )
}

int rtx_render_accumulate( rtx_ctx* c, const rtx_params* p ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	do_render( c, p, false ) ;
	RTX_END( c )
}

int rtx_resolve( rtx_ctx* c, uint64_t total_spp ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( c->w == 0 ) throw std::runtime_error( "rtx_resolve: no frame (call rtx_resize)" ) ;
	do_resolve( c, total_spp ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	RTX_END( c )
}

int rtx_pick( rtx_ctx* c, const rtx_params* p, uint32_t x, uint32_t y, uint32_t* thing_id ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	const FrameArgs a = frame_args( c, p ) ;
	if ( x>=a.w || y>=a.h ) throw std::runtime_error( "rtx_pick: pixel outside the image" ) ;
	k_pick<<<1, 32, 0, c->stream>>>( a, x, y, c->d_pick ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	CK( cudaMemcpyAsync( thing_id, c->d_pick, sizeof( uint32_t ), cudaMemcpyDeviceToHost, c->stream ) ) ;
	fault_fetch( c ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	fault_check( c, "rtx_pick" ) ;
	RTX_END( c )
}

int rtx_postproc( rtx_ctx* c, int kind ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( c->w == 0 ) throw std::runtime_error( "rtx_postproc: no frame (call rtx_resize)" ) ;
	if ( kind != RTX_PP_NONE && kind != RTX_PP_SRGB ) throw std::runtime_error( "rtx_postproc: unknown kind" ) ;
	const uint32_t np = c->w*c->h ;
	CK( cudaEventRecord( c->ev2, c->stream ) ) ;
	k_postproc<<<( np+255 )/256, 256, 0, c->stream>>>( c->d_raw, c->d_image, np, kind == RTX_PP_SRGB ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	CK( cudaEventRecord( c->ev_in, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	CK( cudaEventElapsedTime( &c->ms_postproc, c->ev2, c->ev_in ) ) ;
	RTX_END( c )
}

int rtx_postproc_dev( rtx_ctx* c, int kind, const void* src, void* dst, int w, int h ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( kind != RTX_PP_NONE && kind != RTX_PP_SRGB ) throw std::runtime_error( "rtx_postproc_dev: unknown kind" ) ;
	if ( w<1 || h<1 ) throw std::runtime_error( "rtx_postproc_dev: bad size" ) ;
	const uint32_t np = uint32_t( w )*uint32_t( h ) ;
	k_postproc<<<( np+255 )/256, 256, 0, c->stream>>>( static_cast<const float*>( src ), static_cast<uchar4*>( dst ), np, kind == RTX_PP_SRGB ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	RTX_END( c )
}

int rtx_primary_hits( rtx_ctx* c, const rtx_params* p ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	const FrameArgs a = frame_args( c, p ) ;
	k_primary_hits<<<tile_grid( a.w, a.h ), RTX_BLOCK, 0, c->stream>>>( a ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	fault_fetch( c ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	fault_check( c, "rtx_primary_hits" ) ;
	RTX_END( c )
}

int rtx_trace_rays( rtx_ctx* c, uint32_t n, const float* ori, const float* dir, float tmin, int brute, int64_t* id_out, float* t_out ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( ! c->built ) throw std::runtime_error( "rtx_trace_rays: acceleration structure not built" ) ;
	if ( n == 0 ) return 0 ;
	float* d_o = dalloc<float>( c, 3*size_t( n ) ) ; float* d_d = dalloc<float>( c, 3*size_t( n ) ) ;
	int64_t* d_id = dalloc<int64_t>( c, n ) ; float* d_t = dalloc<float>( c, n ) ;
	CK( cudaMemcpyAsync( d_o, ori, sizeof( float )*3*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaMemcpyAsync( d_d, dir, sizeof( float )*3*n, cudaMemcpyHostToDevice, c->stream ) ) ;
	k_trace_rays<<<( n+RTX_BLOCK-1 )/RTX_BLOCK, RTX_BLOCK, 0, c->stream>>>( scene_dev( c ), n, d_o, d_d, tmin, brute, d_id, d_t ) ;
	c->launches += 1 ;
	CK( cudaGetLastError() ) ;
	CK( cudaMemcpyAsync( id_out, d_id, sizeof( int64_t )*n, cudaMemcpyDeviceToHost, c->stream ) ) ;
	if ( t_out ) CK( cudaMemcpyAsync( t_out, d_t, sizeof( float )*n, cudaMemcpyDeviceToHost, c->stream ) ) ;
	fault_fetch( c ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	dfree( c, d_o, 3*size_t( n ) ) ; dfree( c, d_d, 3*size_t( n ) ) ; dfree( c, d_id, n ) ; dfree( c, d_t, n ) ;
	fault_check( c, "rtx_trace_rays" ) ;
	RTX_END( c )
}

int rtx_read( rtx_ctx* c, int buffer, void* host_dst, size_t bytes ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	const BufInfo b = buffer_of( c, buffer ) ;
	if ( ! b.ptr || bytes>b.bytes ) throw std::runtime_error( "rtx_read: buffer missing or request too large" ) ;
	CK( cudaMemcpyAsync( host_dst, b.ptr, bytes, cudaMemcpyDeviceToHost, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	RTX_END( c )
}

int rtx_write( rtx_ctx* c, int buffer, const void* host_src, size_t bytes ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	const BufInfo b = buffer_of( c, buffer ) ;
	if ( ! b.ptr || bytes>b.bytes ) throw std::runtime_error( "rtx_write: buffer missing or request too large" ) ;
	CK( cudaMemcpyAsync( b.ptr, host_src, bytes, cudaMemcpyHostToDevice, c->stream ) ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	RTX_END( c )
}

int rtx_device_ptr( rtx_ctx* c, int buffer, void** dev_ptr, size_t* bytes ) {
	RTX_TRY( c )
	const BufInfo b = buffer_of( c, buffer ) ;
	if ( ! b.ptr ) throw std::runtime_error( "rtx_device_ptr: no frame (call rtx_resize)" ) ;
	*dev_ptr = b.ptr ;
	if ( bytes ) *bytes = b.bytes ;
	RTX_END( c )
}

int rtx_stats_get( rtx_ctx* c, rtx_stats* out ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	memset( out, 0, sizeof( *out ) ) ;
	if ( c->d_accum ) {
		const uint32_t np = c->w*c->h ;
		CK( cudaMemsetAsync( c->d_counter, 0, sizeof( unsigned long long ), c->stream ) ) ;
		k_sum_segments<<<min( 1184u, ( np+255u )/256u ), 256, 0, c->stream>>>( c->d_accum, np, c->d_counter ) ;
		c->launches += 1 ;
		unsigned long long s = 0 ;
		CK( cudaMemcpyAsync( &s, c->d_counter, sizeof( s ), cudaMemcpyDeviceToHost, c->stream ) ) ;
		CK( cudaStreamSynchronize( c->stream ) ) ;
		out->segments = s ;
	}
	out->paths = c->paths ;
	out->ms_render = c->ms_render ; out->ms_build_blas = c->ms_blas ; out->ms_build_tlas = c->ms_tlas ;
	out->launches = c->launches ;
	out->n_things = uint32_t( c->things.size() ) ; out->n_meshes = uint32_t( c->meshes.size() ) ;
	for ( const Mesh& m : c->meshes ) out->n_triangles += m.nt ;
	for ( const ThingHost& t : c->things ) out->n_triangles_instanced += c->meshes[t.mesh].nt ;
	out->bytes_device = c->bytes ;
	RTX_END( c )
}

int rtx_frame_stats_get( rtx_ctx* c, rtx_frame_stats* out, int reset ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	memset( out, 0, sizeof( *out ) ) ;
	out->ms_frame = c->ms_frame ; out->ms_trace = c->ms_render ; out->ms_reduce_resolve = c->ms_resolve ; out->ms_postproc = c->ms_postproc ;
	out->n_devices = 1u+uint32_t( c->replicas.size() ) ;
	out->kernel = uint32_t( c->kernel ) ;
#if defined( RTX_DEVICE_COUNTERS )
	// (the counters of every device replica live in that device's copy of the symbols)
	out->counted = 1 ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	for ( uint32_t r = 0 ; r<out->n_devices ; r++ ) {
		rtx_ctx* d = r ? c->replicas[r-1] : c ;
		if ( r && d->device == c->device ) continue ;   // same device, same symbols
		CK( cudaSetDevice( d->device ) ) ;
		unsigned long long st[8], ln[8], lv[64] ;
		CK( cudaMemcpyFromSymbol( st, g_dev_steps, sizeof( st ) ) ) ; CK( cudaMemcpyFromSymbol( ln, g_dev_lanes, sizeof( ln ) ) ) ; CK( cudaMemcpyFromSymbol( lv, g_dev_live, sizeof( lv ) ) ) ;
		for ( int k = 0 ; k<8 ; k++ ) { out->steps[k] += st[k] ; out->lanes[k] += ln[k] ; }
		for ( int k = 0 ; k<64 ; k++ ) out->live_paths[k] += lv[k] ;
		if ( reset ) {
			memset( st, 0, sizeof( st ) ) ; memset( lv, 0, sizeof( lv ) ) ;
			CK( cudaMemcpyToSymbol( g_dev_steps, st, sizeof( st ) ) ) ; CK( cudaMemcpyToSymbol( g_dev_lanes, st, sizeof( st ) ) ) ; CK( cudaMemcpyToSymbol( g_dev_live, lv, sizeof( lv ) ) ) ;
		}
	}
	CK( cudaSetDevice( c->device ) ) ;
#else
	( void ) reset ;
#endif
	RTX_END( c )
}

int rtx_probe_read( rtx_ctx* c, size_t bytes, uint32_t repeats, float* gb_per_s ) {
	RTX_TRY( c )
	CK( cudaSetDevice( c->device ) ) ;
	if ( bytes<4096 || repeats == 0 || ! gb_per_s ) throw std::runtime_error( "rtx_probe_read: bad arguments" ) ;
	const size_t n_vec = bytes/16 ;
	uint4* buf = talloc<uint4>( c, n_vec ) ;
	CK( cudaMemsetAsync( buf, 0x5a, n_vec*16, c->stream ) ) ;
	int sms = 0 ;
	CK( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, c->device ) ) ;
	const uint32_t grid = uint32_t( sms )*8u ;
	k_probe_read<<<grid, 256, 0, c->stream>>>( buf, n_vec, 1u, c->d_tile_counter ) ;   // warm the cache
	CK( cudaEventRecord( c->ev0, c->stream ) ) ;
	k_probe_read<<<grid, 256, 0, c->stream>>>( buf, n_vec, repeats, c->d_tile_counter ) ;
	CK( cudaEventRecord( c->ev1, c->stream ) ) ;
	c->launches += 2 ;
	CK( cudaGetLastError() ) ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	float ms = 0.f ;
	CK( cudaEventElapsedTime( &ms, c->ev0, c->ev1 ) ) ;
	tfree( c, buf, n_vec ) ;
	*gb_per_s = float( double( n_vec )*16.*double( repeats )/( double( ms )*1e-3 )/1e9 ) ;
	RTX_END( c )
}

int rtx_build_stages( rtx_ctx* c, float blas_ms[5], float tlas_ms[5] ) {
	RTX_TRY( c )
	if ( blas_ms ) memcpy( blas_ms, c->ms_stage_blas, sizeof( c->ms_stage_blas ) ) ;
	if ( tlas_ms ) memcpy( tlas_ms, c->ms_stage_tlas, sizeof( c->ms_stage_tlas ) ) ;
	RTX_END( c )
}

int rtx_last_render_ms( rtx_ctx* c, float* ms ) {
	RTX_TRY( c )
	*ms = c->ms_render ;
	RTX_END( c )
}

int rtx_counters_get( rtx_ctx* c, uint64_t out[8], int reset ) {
	RTX_TRY( c )
#if defined( RTX_DEVICE_COUNTERS )
	unsigned long long v[DC_N] ;
	CK( cudaStreamSynchronize( c->stream ) ) ;
	CK( cudaMemcpyFromSymbol( v, g_dev_counts, sizeof( v ) ) ) ;
	for ( int k = 0 ; k<8 ; k++ ) out[k] = k<DC_N ? uint64_t( v[k] ) : 0u ;
	if ( reset ) {
		memset( v, 0, sizeof( v ) ) ;
		CK( cudaMemcpyToSymbol( g_dev_counts, v, sizeof( v ) ) ) ;
	}
#else
	( void ) out ; ( void ) reset ;
	throw std::runtime_error( "this library was built without RTX_DEVICE_COUNTERS (use librtx_count.so)" ) ;
#endif
	RTX_END( c )
}

} // extern "C"
