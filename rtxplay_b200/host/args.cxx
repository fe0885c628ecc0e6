// args.cxx -- option parsing of rtwo (behaviour of optx/args.cxx:19-151): getopt_long with
// the same short/long names; unknown options throw std::invalid_argument; unknown -A / -D
// arguments are reported and ignored; -a needs a preceding -g.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include <getopt.h>

#include "args.h"

Args::Args( const int argc, char* const* argv ) noexcept( false ) {
	static const char* shorts = "g:a:s:d:vtqhA:GSD:" ;
	const struct option longs[] = {
		{ "geometry",          required_argument, 0,   'g' },
		{ "aspect-ratio",      required_argument, 0,   'a' },
		{ "samples-per-pixel", required_argument, 0,   's' },
		{ "trace-depth",       required_argument, 0,   'd' },
		{ "verbose",           no_argument,       &v_, 1   },
		{ "trace-sm",          no_argument,       &t_, 1   },
		{ "quiet",             no_argument,       &q_, 1   },
		{ "silent",            no_argument,       &q_, 1   },
		{ "help",              no_argument,       &h_, 1   },
		{ "usage",             no_argument,       &h_, 1   },
		{ "print-aov",         required_argument, 0,   'A' },
		{ "print-guides",      no_argument,       &G_, 1   },
		{ "print-statistics",  no_argument,       &S_, 1   },
		{ "apply-denoiser",    required_argument, 0,   'D' },
		{ "analytic",          no_argument,       &analytic_, 1 },
		{ "device",            required_argument, 0,   1000 },
		{ "gpus",              required_argument, 0,   1001 },
		{ 0, 0, 0, 0 }
	} ;
	optind = 1 ;
	for ( int n = 0 ; n<MAXOPT ; n++ ) {
		const int c = getopt_long( argc, argv, shorts, longs, 0 ) ;
		if ( c<0 )
			break ;
		switch ( c ) {
			case 'g': {
				const auto named = res_map.find( optarg ) ;
				if ( named != res_map.end() ) {
					g_w_ = named->second.w ; g_h_ = named->second.h ;
				} else {
					sscanf( optarg, "%dx%d", &g_w_, &g_h_ ) ;
					if ( g_w_<1 || g_h_<1 ) { g_w_ = -1 ; g_h_ = -1 ; }
				}
				break ;
			}
			case 'a': {
				float w = 0.f, h = 0.f ;
				sscanf( optarg, "%f:%f", &w, &h ) ;
				if ( g_w_>0 && w>0.f && h>0.f )
					g_h_ = static_cast<int>( static_cast<float>( g_w_ )*h/w+.5f ) ;
				break ;
			}
			case 's': s_ = abs( atoi( optarg ) ) ; break ;
			case 'd': d_ = abs( atoi( optarg ) ) ; break ;
			case 'v': v_ = 1 ; break ;
			case 't': t_ = 1 ; break ;
			case 'q': q_ = 1 ; break ;
			case 'h': h_ = 1 ; break ;
			case 'G': G_ = 1 ; break ;
			case 'S': S_ = 1 ; break ;
			case 'A': {
				// comma separated list
				char* item = strtok( optarg, "," ) ;
				while ( item ) {
					const auto known = aov_map.find( item ) ;
					if ( known != aov_map.end() ) {
						if ( known->second == Aov::RPP ) A_rpp_ = Aov::RPP ;
					} else
						std::cerr << "rtwo: unknown argument for option A ignored -- " << item << std::endl ;
					item = strtok( nullptr, "," ) ;
				}
				break ;
			}
			case 'D': {
				const auto known = dns_map.find( optarg ) ;
				if ( known != dns_map.end() )
					D_typ_ = known->second ;
				else
					std::cerr << "rtwo: unknown argument for option D ignored -- " << optarg << std::endl ;
				break ;
			}
			case 1000: device_ = abs( atoi( optarg ) ) ; break ;
			case 1001: gpus_ = abs( atoi( optarg ) ) ; break ;
			case '?':
				throw std::invalid_argument( "try 'rtwo --help' for more information." ) ;
			default: // 0: a flag was stored by getopt_long
				break ;
		}
	}
}

int  Args::param_w( const int dEfault ) const { return 0>g_w_ ? dEfault : g_w_ ; }
int  Args::param_h( const int dEfault ) const { return 0>g_h_ ? dEfault : g_h_ ; }
int  Args::param_s( const int dEfault ) const { return 0>s_ ? dEfault : s_ ; }
int  Args::param_d( const int dEfault ) const { return 0>d_ ? dEfault : d_ ; }
Dns  Args::param_D( const Dns dEfault ) const { return D_typ_ == Dns::NONE ? dEfault : D_typ_ ; }

bool Args::flag_v() const { return v_>0 ; }
bool Args::flag_h() const { return h_>0 ; }
bool Args::flag_q() const { return q_>0 ; }
bool Args::flag_t() const { return t_>0 ; }
bool Args::flag_G() const { return G_>0 ; }
bool Args::flag_S() const { return S_>0 ; }
bool Args::flag_A( const Aov select ) const { return A_rpp_ == select ; }

void Args::usage() {
	std::cerr <<
"Usage: rtwo [OPTION...]\n"
"  rtwo renders the final image of Pete Shirley's book Ray Tracing in One Weekend on an\n"
"  NVIDIA B200 with hand-written CUDA kernels (LBVH build, traversal, shading) and pipes\n"
"  the result (PPM) to stdout, e.g.  rtwo -S | magick ppm:- rtwo.png\n"
"\n"
"Options (those of RTXplay's rtwo):\n"
"  -g, --geometry {<width>x<height>|RES}   image size; RES one of CGA HVGA VGA WVGA SVGA XGA\n"
"                                          HD (default) SXGA UXGA FullHD 2K QXGA UWHD WQHD\n"
"                                          WQXGA UWQHD UHD-1 4K 5K-UW 5K UHD-2\n"
"  -a, --aspect-ratio <width>:<height>     derive the height from -g's width\n"
"  -s, --samples-per-pixel N               default 50\n"
"  -d, --trace-depth N                     scatter events per path, default and maximum 50\n"
"  -v, --verbose                           print processing details on stderr\n"
"  -t, --trace-sm                          (interactive mode only; ignored)\n"
"  -q, --quiet, --silent                   no output on stdout\n"
"  -h, --help, --usage                     this text\n"
"  -A, --print-aov <AOV>[,...]             after the image; AOVs: RPP (rays per pixel, PGM)\n"
"  -G, --print-guides                      print guide layers before the image (with -D)\n"
"  -S, --print-statistics                  pixels, rays, milliseconds, fps on stderr\n"
"  -D, --apply-denoiser <TYP>              SMP NRM ALB NAA AOV: guide layers only, the OptiX AI\n"
"                                          denoiser is not part of this build\n"
"Additional:\n"
"      --analytic                          analytic spheres (the CPU path's geometry) instead of meshes\n"
"      --device N                          CUDA device index\n"
"      --gpus N                            render on N GPUs of this node (devices N0 .. N0+N-1, N0 = --device):\n"
"                                          the samples of the frame are split, the sums reduced on the first\n"
"\n" ;
}

#ifdef MAIN

int main( int argc, char* argv[] ) {
	try {
		Args args( argc, argv ) ;
		if ( args.flag_h() ) { Args::usage() ; return 0 ; }
		// same report as the reference's harness (optx/args.cxx:294-325)
		std::cout << "geometry   : " << args.param_w( 4711 ) << "x" << args.param_h( 4711 ) << std::endl ;
		std::cout << "spp        : " << args.param_s( 4711 ) << std::endl ;
		std::cout << "depth      : " << args.param_d( 4711 ) << std::endl ;
		std::cout << "denoiser   : " << static_cast<int>( args.param_D( Dns::NONE ) ) << std::endl ;
		std::cout << "verbose    : " << ( args.flag_v() ? "set" : "not set" ) << std::endl ;
		std::cout << "quiet      : " << ( args.flag_q() ? "set" : "not set" ) << std::endl ;
		std::cout << "silent     : " << ( args.flag_q() ? "set" : "not set" ) << std::endl ;
		std::cout << "trace-sm   : " << ( args.flag_t() ? "set" : "not set" ) << std::endl ;
		std::cout << "guides     : " << ( args.flag_G() ? "set" : "not set" ) << std::endl ;
		std::cout << "statistics : " << ( args.flag_S() ? "set" : "not set" ) << std::endl ;
		std::cout << "aov RPP    : " << ( args.flag_A( Aov::RPP ) ? "set" : "not set" ) << std::endl ;
	} catch ( const std::invalid_argument& e ) {
		std::cerr << e.what() << std::endl ;
		return 1 ;
	}
	return 0 ;
}

#endif // MAIN
