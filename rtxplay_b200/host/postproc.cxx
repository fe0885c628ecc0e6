// postproc.cxx -- the post-processing entry points of the reference with their original
// C signatures (optx/postproc.cu:49-61: extern "C" pp_none / pp_sRGB on device buffers),
// forwarding to librtx's kernel on the context the launcher last used.
#include <vector_types.h>

#include "util.h"

namespace cg { rtx_ctx* rtx_context = nullptr ; }

extern "C" void pp_none( const float3* src, uchar4* dst, const int w, const int h ) {
	RTX_CHECK( cg::rtx_context, rtx_postproc_dev( cg::rtx_context, RTX_PP_NONE, src, dst, w, h ) ) ;
}

extern "C" void pp_sRGB( const float3* src, uchar4* dst, const int w, const int h ) {
	RTX_CHECK( cg::rtx_context, rtx_postproc_dev( cg::rtx_context, RTX_PP_SRGB, src, dst, w, h ) ) ;
}
