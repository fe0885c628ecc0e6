// object.cxx -- in-house Wavefront OBJ reader for scene files (`v` and triangular `f`
// records, `o`/`g` start a submesh).  Stands in for optx/object.cxx:34-93, which wraps an
// un-vendored tinyobjloader: per submesh the vertex list holds every vertex once, in order
// of first use, and the triangles index that list.  (For the reference's own scene files,
// whose vertices are already in first-use order, this equals its global file indices; for
// other files the reference's mix of global indices and a compacted list is not reproduced.)
// Faces may use the v, v/vt, v//vn, v/vt/vn forms and negative (relative) indices.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "object.h"

Object::Object( const std::string& wavefront ) {
	procWavefrontObj( wavefront ) ;
}

Object::Object( const Mesh& mesh ) {
	float3* v ; unsigned int nv ; uint3* i ; unsigned int ni ;
	std::tie( v, nv, i, ni ) = mesh ;
	vces_.push_back( std::vector<float3>( v, v+nv ) ) ;
	ices_.push_back( std::vector<uint3>( i, i+ni ) ) ;
}

const Mesh Object::operator[] ( unsigned int m ) {
	return Mesh( vces_[m].data(), static_cast<unsigned int>( vces_[m].size() ), ices_[m].data(), static_cast<unsigned int>( ices_[m].size() ) ) ;
}

void Object::procWavefrontObj( const std::string& wavefront ) {
	std::ifstream in( wavefront ) ;
	if ( ! in )
		throw std::runtime_error( wavefront+": cannot open scene file\n" ) ;

	std::vector<float3> all ;                   // every `v` of the file
	std::map<unsigned int, unsigned int> recall ; // file index -> index in this submesh
	std::vector<float3> vces ;
	std::vector<uint3>  ices ;
	bool open = false ;

	auto flush = [&]() {
		if ( open && ! ices.empty() ) { vces_.push_back( vces ) ; ices_.push_back( ices ) ; }
		vces.clear() ; ices.clear() ; recall.clear() ; open = false ;
	} ;

	std::string line ;
	while ( std::getline( in, line ) ) {
		std::istringstream rec( line ) ;
		std::string tag ;
		if ( ! ( rec >> tag ) || tag[0] == '#' )
			continue ;
		if ( tag == "v" ) {
			float3 p ;
			if ( ! ( rec >> p.x >> p.y >> p.z ) )
				throw std::runtime_error( wavefront+": malformed vertex\n" ) ;
			all.push_back( p ) ;
		} else if ( tag == "o" || tag == "g" ) {
			flush() ;
		} else if ( tag == "f" ) {
			unsigned int idx[3] ; int n = 0 ;
			std::string ref ;
			while ( rec >> ref ) {
				if ( n == 3 )
					throw std::runtime_error( wavefront+": triangle faces expected\n" ) ;
				long k = strtol( ref.c_str(), nullptr, 10 ) ;          // leading integer of v[/vt][/vn]
				if ( k<0 ) k = static_cast<long>( all.size() )+k+1 ;   // relative index
				if ( k<1 || static_cast<size_t>( k )>all.size() )
					throw std::runtime_error( wavefront+": index out of bounds\n" ) ;
				idx[n++] = static_cast<unsigned int>( k-1 ) ;
			}
			if ( n != 3 )
				throw std::runtime_error( wavefront+": triangle faces expected\n" ) ;
			open = true ;
			unsigned int loc[3] ;
			for ( int c = 0 ; c<3 ; c++ ) {
				auto seen = recall.find( idx[c] ) ;
				if ( seen == recall.end() ) {
					loc[c] = static_cast<unsigned int>( vces.size() ) ;
					recall[idx[c]] = loc[c] ;
					vces.push_back( all[idx[c]] ) ;
				} else
					loc[c] = seen->second ;
			}
			ices.push_back( { loc[0], loc[1], loc[2] } ) ;
		}
	}
	flush() ;
	if ( vces_.empty() )
		throw std::runtime_error( wavefront+": no triangles\n" ) ;
}

#ifdef MAIN
#include <cstdio>
#include <iostream>

int main( const int argc, const char** argv ) {
	if ( argc<2 ) return 1 ;
	Object object( argv[1] ) ;
	std::cout << "o " << argv[1] << std::endl ;
	float3* vces ; unsigned int nv ; uint3* ices ; unsigned int ni ;
	unsigned int vsum = 0 ;
	for ( unsigned int o = 0 ; object.size()>o ; o++ ) {
		std::tie( vces, nv, ices, ni ) = object[o] ;
		for ( unsigned int v = 0 ; v<nv ; v++ )
			printf( "v %f %f %f\n", vces[v].x, vces[v].y, vces[v].z ) ;
		vsum += nv ;
	}
	std::cout << "# " << vsum << " vertices" << std::endl ;
	for ( unsigned int o = 0 ; object.size()>o ; o++ ) {
		std::tie( vces, nv, ices, ni ) = object[o] ;
		for ( unsigned int i = 0 ; i<ni ; i++ )
			printf( "f %d %d %d\n", ices[i].x+1, ices[i].y+1, ices[i].z+1 ) ;
		std::cout << "# " << ni << " triangles" << std::endl ;
	}
	return 0 ;
}
#endif // MAIN
