// l1probe.cu -- development microbenchmark (not product): how many cycles does the SM's L1
// data path spend per load instruction as a function of how the 32 lanes' addresses are
// spread over 128-byte lines?  Decides whether k_render's node fetches (one 128-byte node =
// 4 x LDG.256 per lane) are charged per lane or per distinct line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1probe l1probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct o8 { float v[8] ; } ;
__device__ __forceinline__ o8 ld256( const void* p ) {
	o8 r ;
	asm volatile( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"( r.v[0] ), "=f"( r.v[1] ), "=f"( r.v[2] ), "=f"( r.v[3] ), "=f"( r.v[4] ), "=f"( r.v[5] ), "=f"( r.v[6] ), "=f"( r.v[7] ) : "l"( p ) ) ;
	return r ;
}
__device__ __forceinline__ float sum8( const o8& a ) { return ( ( a.v[0]+a.v[1] )+( a.v[2]+a.v[3] ) )+( ( a.v[4]+a.v[5] )+( a.v[6]+a.v[7] ) ) ; }
__device__ __forceinline__ float4 ld128( const void* p ) {
	float4 r ;
	asm volatile( "ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( r.x ), "=f"( r.y ), "=f"( r.z ), "=f"( r.w ) : "l"( p ) ) ;
	return r ;
}

// mode: lanes per distinct line (1: all lanes different lines ... 32: all lanes the same line)
// width: 32 -> one 256-bit load of chunk 0 ; 128 -> four 256-bit loads (a whole line) ; 16 -> one 128-bit load
template <int WIDTH>
__global__ void k_probe( const char* buf, uint32_t lines_mask, int share, int iters, float* sink, unsigned long long* cycles ) {
	const uint32_t lane = threadIdx.x&31u, warp = ( blockIdx.x*blockDim.x+threadIdx.x )>>5 ;
	uint32_t state = warp*2654435761u+12345u ;
	float acc = 0.f ;
	const unsigned long long t0 = clock64() ;
	for ( int it = 0 ; it<iters ; it++ ) {
		state = state*1664525u+1013904223u ;
		// the group's line: lanes lane/share share one
		const uint32_t g = lane/uint32_t( share ) ;
		const uint32_t line = ( ( state>>8 )+g*7919u )&lines_mask ;
		const char* p = buf+size_t( line )*128u ;
		if ( WIDTH == 16 ) { const float4 a = ld128( p ) ; acc += ( a.x+a.y )+( a.z+a.w ) ; }
		else if ( WIDTH == 32 ) { const o8 a = ld256( p ) ; acc += sum8( a ) ; }
		else { const o8 a = ld256( p ), b = ld256( p+32 ), c = ld256( p+64 ), d = ld256( p+96 ) ; acc += ( sum8( a )+sum8( b ) )+( sum8( c )+sum8( d ) ) ; }
	}
	const unsigned long long t1 = clock64() ;
	if ( acc == 12345.678f ) *sink = acc ;
	if ( lane == 0 ) atomicMax( cycles, t1-t0 ) ;
}

// shared memory: 128-bit loads of records at stride `stride_words`, slot chosen per lane:
// pattern 0: slot = lane (conflict free with an odd quad stride), 1: random slots out of n_slots
__global__ void k_probe_smem( int stride_words, int n_slots, int pattern, int iters, float* sink, unsigned long long* cycles ) {
	extern __shared__ float sm[] ;
	const uint32_t lane = threadIdx.x&31u, warp = ( blockIdx.x*blockDim.x+threadIdx.x )>>5 ;
	for ( int i = threadIdx.x ; i<stride_words*n_slots ; i += blockDim.x ) sm[i] = float( i ) ;
	__syncthreads() ;
	uint32_t state = ( warp*32u+lane )*2654435761u+12345u ;
	float acc = 0.f ;
	const unsigned long long t0 = clock64() ;
	for ( int it = 0 ; it<iters ; it++ ) {
		state = state*1664525u+1013904223u ;
		const uint32_t slot = pattern == 0 ? ( lane+uint32_t( it ) )%uint32_t( n_slots ) : ( state>>10 )%uint32_t( n_slots ) ;
		const float4 a = *reinterpret_cast<const float4*>( sm+size_t( slot )*stride_words ) ;
		acc += ( a.x+a.y )+( a.z+a.w ) ;
	}
	const unsigned long long t1 = clock64() ;
	if ( acc == 12345.678f ) *sink = acc ;
	if ( lane == 0 ) atomicMax( cycles, t1-t0 ) ;
}

int main() {
	char* buf ; float* sink ; unsigned long long* cyc ;
	const size_t bytes = size_t( 64 )<<10 ;             // 64 KB: stays L1 resident
	cudaMalloc( &buf, bytes ) ; cudaMemset( buf, 0, bytes ) ;
	cudaMalloc( &sink, 4 ) ; cudaMalloc( &cyc, 8 ) ;
	int sms = 0 ; cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, 0 ) ;
	const int iters = 20000 ;
	for ( int warps = 8 ; warps<=32 ; warps *= 2 )
	for ( int width : { 16, 32, 128 } )
		for ( int share : { 1, 2, 4, 8, 32 } ) {
			unsigned long long h = 0 ;
			for ( int rep = 0 ; rep<2 ; rep++ ) {
				cudaMemset( cyc, 0, 8 ) ;
				const uint32_t mask = uint32_t( bytes/128 )-1u ;
				if ( width == 16 ) k_probe<16><<<sms, 32*warps>>>( buf, mask, share, iters, sink, cyc ) ;
				else if ( width == 32 ) k_probe<32><<<sms, 32*warps>>>( buf, mask, share, iters, sink, cyc ) ;
				else k_probe<128><<<sms, 32*warps>>>( buf, mask, share, iters, sink, cyc ) ;
				cudaDeviceSynchronize() ;
				cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ) ;
			}
			const double per_warp_iter = double( h )/iters/warps ;   // SM cycles per warp-iteration
			printf( "L1 warps/SM %2d  bytes/lane %3d  lanes/line %2d : %.2f cycles per warp-iteration (%.3f per lane)\n", warps, width, share, per_warp_iter, per_warp_iter/32. ) ;
		}
	cudaFuncSetAttribute( k_probe_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100<<10 ) ;
	for ( int stride : { 36, 44, 32 } )
		for ( int pattern : { 0, 1 } ) {
			unsigned long long h = 0 ;
			const int warps = 16, n_slots = 96 ;
			for ( int rep = 0 ; rep<2 ; rep++ ) {
				cudaMemset( cyc, 0, 8 ) ;
				k_probe_smem<<<sms, 32*warps, size_t( stride )*n_slots*4>>>( stride, n_slots, pattern, iters, sink, cyc ) ;
				cudaDeviceSynchronize() ;
				cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost ) ;
			}
			printf( "SMEM LDS.128 stride %d words, %s slots: %.2f cycles per warp-iteration\n", stride, pattern ? "random" : "lane-ordered", double( h )/iters/warps ) ;
		}
	printf( "%s\n", cudaGetErrorString( cudaGetLastError() ) ) ;
	return 0 ;
}
