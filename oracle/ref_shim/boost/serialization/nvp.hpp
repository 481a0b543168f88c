// Shim: the minimum of boost::serialization the reference headers name. Nothing is ever archived.
#pragma once
#include <cstddef>
namespace boost { namespace serialization {
class access {};
template<class T> struct nvp_stub { T* p; };
template<class T> inline nvp_stub<T> make_nvp(const char*, T& t){ return nvp_stub<T>{&t}; }
template<class T> struct array_stub { T* p; std::size_t n; };
template<class T> inline array_stub<T> make_array(T* p, std::size_t n){ return array_stub<T>{p,n}; }
}}
