// rtx_hostmath.h -- host-side arithmetic that is part of the float contract.
#pragma once

namespace rtx {

// world->object inverse of a row-major 3x4 affine map: cofactors in double (the float
// contract fixes this expression; the oracle states the same one independently)
inline void affine_inverse( const float xf[12], double inv[12] ) {
	const double a = xf[0], b = xf[1], c = xf[2],  tx = xf[3] ;
	const double d = xf[4], e = xf[5], f = xf[6],  ty = xf[7] ;
	const double g = xf[8], h = xf[9], i = xf[10], tz = xf[11] ;
	const double A =  ( e*i-f*h ), B = -( d*i-f*g ), C =  ( d*h-e*g ) ;
	const double D = -( b*i-c*h ), E =  ( a*i-c*g ), F = -( a*h-b*g ) ;
	const double G =  ( b*f-c*e ), H = -( a*f-c*d ), I =  ( a*e-b*d ) ;
	const double det = a*A+b*B+c*C ;
	const double s = 1./det ;
	inv[0] = s*A ; inv[1] = s*D ; inv[2]  = s*G ;
	inv[4] = s*B ; inv[5] = s*E ; inv[6]  = s*H ;
	inv[8] = s*C ; inv[9] = s*F ; inv[10] = s*I ;
	inv[3]  = -( inv[0]*tx+inv[1]*ty+inv[2]*tz ) ;
	inv[7]  = -( inv[4]*tx+inv[5]*ty+inv[6]*tz ) ;
	inv[11] = -( inv[8]*tx+inv[9]*ty+inv[10]*tz ) ;
}


} // namespace rtx
