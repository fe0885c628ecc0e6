/* rtx.h -- C ABI of librtx.so, the B200-native replacement of RTXplay/RTWO's
 * OptiX hot path.  Plain pointers and sizes only; every call is synchronous w.r.t. the
 * host on return (the reference blocks in Launcher::ignite, optx/launcher.cxx:70).
 *
 * Each entry point names the reference interface it stands in for (file:line under
 * /root/reference).  The C++ shims in rtxplay_b200/host/ (Scene, Launcher, pp_sRGB ...)
 * sit on top of this header and keep the reference's class API source-compatible;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: 0 = success, non-zero = failure with the text in rtx_last_error();
 * the library owns all device memory; host pointers passed in are copied before the
 * call returns; one calling thread per context.  There is NO CPU fallback: without a
 * CUDA device rtx_init fails.
 */
#ifndef RTX_H
#define RTX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtx_ctx rtx_ctx ;

/* optx/thing.h:17-45 -- Optics { int type ; union { Diffuse, Reflect, Refract } } */
enum { RTX_OPTICS_DIFFUSE = 0, RTX_OPTICS_REFLECT = 1, RTX_OPTICS_REFRACT = 2 } ;
typedef struct {
	int32_t type ;
	float   albedo[3] ;  /* diffuse / reflect (wavefront Kd) */
	float   fuzz ;       /* reflect */
	float   index ;      /* refract (wavefront Ni) */
} rtx_optics ;

/* The derived camera of optx/camera.h:30-48 (eye, u, v, hvec, wvec, dvec, aperture):
 * what LpGeneral carries by value to the device (optx/rtwo.h:37). */
typedef struct {
	float eye[3], u[3], v[3], hvec[3], wvec[3], dvec[3] ;
	float aperture ;
} rtx_camera ;

/* optx/rtwo.h:26-52 LpGeneral minus device pointers and the OptiX handle. */
typedef struct {
	uint32_t   image_w, image_h ;
	uint32_t   spp ;            /* samples THIS call renders per pixel */
	uint32_t   depth ;          /* max scatter events per path (rtow.cxx:99) */
	rtx_camera camera ;
	uint64_t   seed ;           /* stream key; the reference's constant is 4711 (optx/frand48.h:19) */
	uint32_t   sample0 ;        /* global index of this call's first sample ... */
	uint32_t   sample_stride ;  /* ... and the step between its samples (multi-GPU spp split) */
	uint32_t   accumulate ;     /* 0: start from zero (optx/camera_i.cu:52); 1: add to the buffers */
	uint32_t   guides ;         /* 1: also fill the denoiser guide layers normals / albedos (optx/camera_i.cu:56-57,
	                               99-101, 109-113; optx/optics_i.cu:97-101, 185-189) */
	uint32_t   variant ;        /* RTX_VARIANT_*: which of the reference's programs the path semantics follow */
} rtx_params ;

/* The reference's own variants disagree in four places (SURVEY.md 8a "divergences").  RTOW is the
 * CPU path rtow.cxx, the parity target and the default: pixel -> viewport over w-1 / h-1
 * (rtow.cxx:112-113), Lambert with the near-zero guard (optics.h:17-18), `depth` scatter events and
 * black when they are used up (rtow.cxx:39-42).  RTWO_I is the default OptiX build (iterative
 * programs): over w / h (optx/camera_i.cu:61-62), no guard (optx/optics_i.cu:86), at most `depth`
 * rays per path, and a path whose last ray still scattered keeps its throughput product as colour
 * (optx/camera_i.cu:75-95).  RTWO_R is the -DRECURSIVE build: over w / h (optx/camera_r.cu:69-70),
 * no guard (optx/optics_r.cu:108), black at the `depth`-th hit (optx/optics_r.cu:30-42). */
enum { RTX_VARIANT_RTOW = 0, RTX_VARIANT_RTWO_I = 1, RTX_VARIANT_RTWO_R = 2 } ;

typedef struct {
	uint64_t segments ;         /* closest-hit queries of the last render (sum of RPP, optx/rtwo.cxx:588) */
	uint64_t paths ;
	float    ms_render ;        /* device time of the path-tracing kernel(s) of the last rtx_render */
	float    ms_build_blas ;    /* device time of all mesh LBVH builds so far */
	float    ms_build_tlas ;    /* device time of the last top-level build or refit (from the upload of the thing records on) */
	uint32_t launches ;         /* kernels launched by this context so far */
	uint32_t n_things, n_meshes ;
	uint64_t n_triangles ;      /* unique triangles stored */
	uint64_t n_triangles_instanced ; /* sum over things of their mesh's triangles */
	uint64_t bytes_device ;     /* device memory held */
} rtx_stats ;

/* Per-frame figures of the last rtx_render* (the reference's recipe for these is an ncu run,
 * optx/README.md:294-316; SURVEY.md 5 asks for them from the library).  Stage times are device
 * times from CUDA events on the context's stream.  The scheduling figures come from the kernels'
 * own counters and are filled by the instrumented build only (librtx_count.so; `counted` = 1):
 * per step kind k (1 node, 2 leaf, 3 top-level leaf, 4 shading, 5 new path) the warp iterations that
 * ran it and the lanes that took part -- lanes[k] / steps[k] is the SIMD utilisation ncu reports as
 * smsp__thread_inst_executed_per_inst_executed, per kind -- and live_paths[b] = rays traced at
 * bounce b (0 = primary; 63 = 63 and deeper), i.e. the paths still alive there. */
typedef struct {
	float    ms_frame ;          /* clears + path tracing + (multi-GPU) reduce + resolve */
	float    ms_trace ;          /* the path-tracing kernel on the first device */
	float    ms_reduce_resolve ; /* k_resolve, or the fused peer-memory reduce + resolve */
	float    ms_postproc ;       /* the last rtx_postproc */
	uint32_t n_devices ;
	uint32_t kernel ;            /* 0: k_render (one ray per lane), 1: k_render_q (compacting ray pool) */
	uint32_t counted ;           /* 1: the fields below are filled */
	uint32_t pad ;
	uint64_t steps[8], lanes[8] ;
	uint64_t live_paths[64] ;
} rtx_frame_stats ;

/* buffers of rtx_read / rtx_device_ptr */
enum {
	RTX_BUF_ACCUM   = 0,  /* uint64[4*w*h]: fixed-point (2^-32) radiance sums r,g,b and segment count */
	RTX_BUF_RAWRGB  = 1,  /* float[3*w*h]  LpGeneral.rawRGB  (optx/launcher.cxx:39) */
	RTX_BUF_RPP     = 2,  /* uint32[w*h]   LpGeneral.rpp     (optx/launcher.cxx:41) */
	RTX_BUF_IMAGE   = 3,  /* uint8[4*w*h]  LpGeneral.image   (optx/rtwo.cxx:564)   */
	RTX_BUF_HIT_ID  = 4,  /* int64[w*h]    first-hit id of rtx_primary_hits: thing<<32 | prim+1, -1 = miss */
	RTX_BUF_HIT_T   = 5,  /* float[w*h]    its ray parameter */
	RTX_BUF_NORMALS = 6,  /* float[3*w*h]  LpGeneral.normals (optx/launcher.cxx:43) */
	RTX_BUF_ALBEDOS = 7,  /* float[3*w*h]  LpGeneral.albedos (optx/launcher.cxx:44) */
	RTX_BUF_PICK_ID = 8,  /* uint32[1]     LpGeneral.pick_id (optx/launcher.cxx:46), written by rtx_pick */
	RTX_BUF_GUIDE_ACC = 9 /* int64[6*w*h]  fixed-point (2^-30) sums of the guide layers: normal xyz, albedo rgb */
} ;

enum { RTX_PP_NONE = 0, RTX_PP_SRGB = 1 } ;

/* cudaFree(0) + optixInit + optixDeviceContextCreate, optx/rtwo.cxx:115-128 */
int  rtx_init( int device, rtx_ctx** out ) ;
/* The same on several GPUs of one node (the reference is single-context, optx/rtwo.cxx:126-127; this
 * is where a caller reaches the other devices).  The context returned stands for all of them: scene
 * calls are repeated on every device (scene and hierarchies are replicated), rtx_render* splits the
 * samples of a frame over the devices -- device r traces the global samples r, r+n, ... -- and one
 * kernel on the first device sums the fixed-point accumulation buffers through peer memory (NVLink /
 * NVSwitch; staged copies where two devices are not peers) and resolves the frame.  Integer sums:
 * the frame equals the one-device frame bit for bit.  1 <= n_devices <= 8; a device may be named more
 * than once (two replicas on one GPU: used by the tests on a single-GPU box). */
int  rtx_init_multi( int n_devices, const int* device_ids, rtx_ctx** out ) ;
int  rtx_device_count( const rtx_ctx* ctx ) ;
void rtx_shutdown( rtx_ctx* ctx ) ;
/* util_cpu.h:18-53 CUDA_CHECK/OPTX_CHECK text; ctx may be NULL for rtx_init failures */
const char* rtx_last_error( const rtx_ctx* ctx ) ;

/* Scene::add( Object& ), optx/scene.cxx:38-167: upload float3 vertices + uint3 indices and
 * build the mesh's bottom-level LBVH (replaces optixAccelBuild on a GAS). */
int rtx_mesh_create( rtx_ctx* ctx, const float* xyz, uint32_t n_vertices, const uint32_t* idx, uint32_t n_triangles, uint32_t* mesh_id ) ;
/* The analytic unit sphere of the CPU path (sphere.h:20-48) as a pseudo mesh: a thing
 * instancing it is the sphere centre = translation, radius = transform[0]. */
int rtx_sphere_create( rtx_ctx* ctx, uint32_t* mesh_id ) ;
/* Scene::add( Thing&, unsigned object ), optx/scene.cxx:169-196 (identity transform) */
int rtx_thing_add( rtx_ctx* ctx, uint32_t mesh_id, const rtx_optics* optics, uint32_t* thing_id ) ;
/* Scene::set / Scene::get, optx/scene.cxx:198-225: row-major 3x4 object->world */
int rtx_thing_set_xf( rtx_ctx* ctx, uint32_t thing_id, const float xf[12] ) ;
int rtx_thing_get_xf( rtx_ctx* ctx, uint32_t thing_id, float xf[12] ) ;
int rtx_thing_set_optics( rtx_ctx* ctx, uint32_t thing_id, const rtx_optics* optics ) ;
/* Scene::build, optx/scene.cxx:227-273 (IAS build) and Scene::update, :275-294 (IAS refit) */
int rtx_accel_build( rtx_ctx* ctx ) ;
int rtx_accel_refit( rtx_ctx* ctx ) ;

/* Launcher::resize, optx/launcher.cxx:36-47 */
int rtx_resize( rtx_ctx* ctx, uint32_t w, uint32_t h ) ;
/* Launcher::ignite, optx/launcher.cxx:49-77 + __raygen__camera's per-pixel mean/clamp
 * (optx/camera_i.cu:105): path-trace p->spp samples per pixel into the accumulation
 * buffer and resolve rawRGB / rpp.  Blocking. */
int rtx_render( rtx_ctx* ctx, const rtx_params* p ) ;
/* The same without the resolve: used when partial sums of several GPUs are reduced first. */
int rtx_render_accumulate( rtx_ctx* ctx, const rtx_params* p ) ;
/* mean -> clamp(0,1) -> rawRGB, rpp; total_spp = samples summed into the buffer */
int rtx_resolve( rtx_ctx* ctx, uint64_t total_spp ) ;
/* Launcher::ignite( stream, once=true ) + picker (optx/camera_i.cu:27-29, optx/simplesm.cxx:1014-1034):
 * one primary ray through the pixel; UINT32_MAX = miss */
int rtx_pick( rtx_ctx* ctx, const rtx_params* p, uint32_t x, uint32_t y, uint32_t* thing_id ) ;
/* pp_none / pp_sRGB, optx/postproc.cu:49-61: rawRGB -> uchar4 image buffer */
int rtx_postproc( rtx_ctx* ctx, int kind ) ;
/* the same on caller-owned DEVICE buffers, for callers that keep the reference signature */
int rtx_postproc_dev( rtx_ctx* ctx, int kind, const void* src_float3_dev, void* dst_uchar4_dev, int w, int h ) ;

/* parity instruments: first hit of the primary rays of one sample, and closest hits of
 * caller-supplied rays through the LBVH (brute=0) or by exhaustive scan on the GPU (brute=1) */
int rtx_primary_hits( rtx_ctx* ctx, const rtx_params* p ) ;
int rtx_trace_rays( rtx_ctx* ctx, uint32_t n, const float* ori_xyz, const float* dir_xyz, float tmin, int brute, int64_t* id_out, float* t_out ) ;

/* imgtopnm's device->host copy, optx/rtwo.cxx:61-74 */
int rtx_read( rtx_ctx* ctx, int buffer, void* host_dst, size_t bytes ) ;
/* device address of a buffer (e.g. to hand the accumulation buffer to NCCL) */
int rtx_device_ptr( rtx_ctx* ctx, int buffer, void** dev_ptr, size_t* bytes ) ;
/* overwrite a buffer from host memory (tests; loading reduced sums back) */
int rtx_write( rtx_ctx* ctx, int buffer, const void* host_src, size_t bytes ) ;

/* the -S line of optx/rtwo.cxx:579-591 and more */
int rtx_stats_get( rtx_ctx* ctx, rtx_stats* out ) ;
/* stage times of the last frame; scheduling counters (instrumented build) since the last reset */
int rtx_frame_stats_get( rtx_ctx* ctx, rtx_frame_stats* out, int reset ) ;
/* measurement instrument: read bandwidth of a `bytes` buffer read `repeats` times by all SMs,
 * bypassing L1 -- the L2 read peak when the buffer fits L2 (SURVEY.md 8(d) asks for it measured
 * on the box), the HBM read peak when it is far larger.  No counterpart in the reference. */
int rtx_probe_read( rtx_ctx* ctx, size_t bytes, uint32_t repeats, float* gb_per_s ) ;
/* device time of the stages of the hierarchy builds (stress configuration):
 * [0] centroid bounds + Morton keys, [1] radix sort, [2] Karras hierarchy, [3] bottom-up
 * boxes, [4] collapse into wide nodes.  blas_ms: summed over all rtx_mesh_create calls so far;
 * tlas_ms: the last rtx_accel_build (a refit fills [3] and [4] only).  Either may be NULL.
 * The reference has no counterpart (optixAccelBuild is opaque, optx/scene.cxx:122, 258). */
int rtx_build_stages( rtx_ctx* ctx, float blas_ms[5], float tlas_ms[5] ) ;
/* device time (CUDA events on the launching stream) of the path-tracing kernel of the last
 * rtx_render / rtx_render_accumulate -- the window optx/rtwo.cxx:542-546 times; no device work */
int rtx_last_render_ms( rtx_ctx* ctx, float* ms ) ;
/* measurement instrument (SURVEY.md 8(d) "counted" work): traversal events since the last reset --
 * [0] rays (segments), [1] node steps (4 box tests each), [2] leaf steps, [3] triangle tests,
 * [4] top-level leaf visits, [5] of those culled by the bounding-sphere pre-test (analytic scenes:
 * sphere tests), [6] mesh entries of the parity instruments.  Only the instrumented build
 * (librtx_count.so, -DRTX_DEVICE_COUNTERS: one global atomic per event) counts; in librtx.so the
 * call fails with a message.  The reference has no counterpart. */
int rtx_counters_get( rtx_ctx* ctx, uint64_t out[8], int reset ) ;

/* host helpers shared by the C++ shims and the tests (no device work) */
/* Camera::set, optx/camera.h:30-48 */
void rtx_camera_set( rtx_camera* cam, const float eye[3], const float pat[3], const float vup[3], float fov, float aspratio, float aperture, float fostance ) ;
/* Sphere tessellation by subdivided tetrahedron, optx/sphere.cxx:28-106.  Call with
 * xyz = idx = NULL to get the counts (2*4^n+2 vertices, 4*4^n triangles). */
int rtx_sphere_mesh( float radius, uint32_t ndiv, float* xyz, uint32_t* n_vertices, uint32_t* idx, uint32_t* n_triangles ) ;

#ifdef __cplusplus
}
#endif

#endif /* RTX_H */
