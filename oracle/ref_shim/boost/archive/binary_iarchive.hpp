// Shim: archive classes that swallow everything.
#pragma once
#include <istream>
namespace boost { namespace archive {
class binary_iarchive { public: explicit binary_iarchive(std::istream&){} template<class T> binary_iarchive& operator&(const T&){ return *this; } };
}}
