// args.h -- command line of rtwo (optx/args.h:20-96): same options, same accessors.
#ifndef ARGS_H
#define ARGS_H

#include <map>
#include <string>

#define MAXOPT 32

typedef struct { int w ; int h ; } res ;
// -g accepts <w>x<h> or one of the reference's 21 names (optx/args.h:23-45)
static const std::map<std::string, res> res_map = {
	{ "CGA",   {  320,  200 } }, { "HVGA",  {  480,  320 } }, { "VGA",    {  640,  480 } },
	{ "WVGA",  {  800,  480 } }, { "SVGA",  {  800,  600 } }, { "XGA",    { 1024,  768 } },
	{ "HD",    { 1280,  720 } }, { "SXGA",  { 1280, 1024 } }, { "UXGA",   { 1600, 1200 } },
	{ "FULLHD",{ 1920, 1080 } }, { "2K",    { 2048, 1080 } }, { "QXGA",   { 2048, 1536 } },
	{ "UWHD",  { 2560, 1080 } }, { "WQHD",  { 2560, 1440 } }, { "WQXGA",  { 2560, 1600 } },
	{ "UWQHD", { 3440, 1440 } }, { "UHD-1", { 3840, 2160 } }, { "4K",     { 4096, 2160 } },
	{ "5K-UW", { 5120, 2160 } }, { "5K",    { 5120, 2880 } }, { "UHD-2",  { 7680, 4320 } }
} ;

enum class Aov { NONE, RPP } ;
static const std::string aov_name[] = { "none", "RPP" } ;
static const std::map<std::string, Aov> aov_map = { { aov_name[static_cast<int>( Aov::RPP )], Aov::RPP } } ;

// the denoiser types are parsed for compatibility; the OptiX AI denoiser itself is out of scope
enum class Dns { NONE, SMP, NRM, ALB, NAA, AOV } ;
static const std::string dns_name[] = { "none", "SMP", "NRM", "ALB", "NAA", "AOV" } ;
static const std::map<std::string, Dns> dns_map = {
	{ "SMP", Dns::SMP }, { "NRM", Dns::NRM }, { "ALB", Dns::ALB }, { "NAA", Dns::NAA }, { "AOV", Dns::AOV }
} ;

class Args {
	public:
		Args( const int argc, char* const* argv ) noexcept( false ) ;

		int  param_w( const int dEfault ) const ; // -g, --geometry <w>x<h>
		int  param_h( const int dEfault ) const ;
		int  param_s( const int dEfault ) const ; // -s, --samples-per-pixel
		int  param_d( const int dEfault ) const ; // -d, --trace-depth

		Dns  param_D( const Dns dEfault ) const ; // -D, --apply-denoiser

		bool flag_v()                     const ; // -v, --verbose
		bool flag_h()                     const ; // -h, --help
		bool flag_q()                     const ; // -q, --quiet
		bool flag_t()                     const ; // -t, --trace-sm
		bool flag_G()                     const ; // -G, --print-guides
		bool flag_S()                     const ; // -S, --print-statistics

		bool flag_A( const Aov select )   const ; // -A, --print-aov

		// additive (not in the reference): --analytic renders the CPU path's analytic spheres
		// instead of the tessellated meshes; --device N selects the GPU
		bool flag_analytic()              const { return analytic_>0 ; }
		int  param_device( const int dEfault ) const { return 0>device_ ? dEfault : device_ ; }
		// additive: --gpus N spreads a frame over N devices (rtx_init_multi)
		int  param_gpus( const int dEfault ) const { return 1>gpus_ ? dEfault : gpus_ ; }

		static void usage() ;

	private:
		int g_w_ = -1, g_h_ = -1 ;
		int s_ = -1, d_ = -1 ;
		Dns D_typ_ = Dns::NONE ;
		int v_ = 0, h_ = 0, q_ = 0, t_ = 0, G_ = 0, S_ = 0 ;
		Aov A_rpp_ = Aov::NONE ;
		int analytic_ = 0, device_ = -1, gpus_ = -1 ;
} ;

#endif // ARGS_H
