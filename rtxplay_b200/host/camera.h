// camera.h -- thin-lens camera of the scene API (optx/camera.h:28-94).  The derived
// vectors are computed by librtx's rtx_camera_set, so that every client (this class, the
// Python mirror, the tests) gets the same bits; the per-ray part (optx/camera.h:64-79)
// lives in the device code (rtx_core.cuh primary_ray).
#ifndef CAMERA_H
#define CAMERA_H

#include <vector_types.h>

#include "../../include/rtx.h"

class Camera {
	public:
		void set( const float3& eye, const float3& pat, const float3& vup, const float fov, const float aspratio, const float aperture, const float fostance ) {
			eye_ = eye ; pat_ = pat ; vup_ = vup ;
			fov_ = fov ; aspratio_ = aspratio ; aperture_ = aperture ; fostance_ = fostance ;
			const float e[3] = { eye.x, eye.y, eye.z }, p[3] = { pat.x, pat.y, pat.z }, u[3] = { vup.x, vup.y, vup.z } ;
			rtx_camera_set( &derived_, e, p, u, fov, aspratio, aperture, fostance ) ;
		}

		float3 eye()      const { return eye_ ; }
		float3 pat()      const { return pat_ ; }
		float3 vup()      const { return vup_ ; }
		float  fov()      const { return fov_ ; }
		float  aperture() const { return aperture_ ; }
		float  fostance() const { return fostance_ ; }
		void eye( const float3& eye )         { set( eye,  pat_, vup_, fov_, aspratio_, aperture_, fostance_ ) ; }
		void pat( const float3& pat )         { set( eye_, pat,  vup_, fov_, aspratio_, aperture_, fostance_ ) ; }
		void vup( const float3& vup )         { set( eye_, pat_, vup,  fov_, aspratio_, aperture_, fostance_ ) ; }
		void fov( const float fov )           { set( eye_, pat_, vup_, fov,  aspratio_, aperture_, fostance_ ) ; }
		void aspratio( const float aspratio ) { set( eye_, pat_, vup_, fov_, aspratio,  aperture_, fostance_ ) ; }
		void aperture( const float aperture ) { set( eye_, pat_, vup_, fov_, aspratio_, aperture,  fostance_ ) ; }
		void fostance( const float fostance ) { set( eye_, pat_, vup_, fov_, aspratio_, aperture_, fostance ) ; }

		// what the launch boundary hands to the device
		const rtx_camera& derived() const { return derived_ ; }

	private:
		float3 eye_, pat_, vup_ ;
		float  fov_, aspratio_, aperture_, fostance_ ;
		rtx_camera derived_ ;
} ;

#endif // CAMERA_H
