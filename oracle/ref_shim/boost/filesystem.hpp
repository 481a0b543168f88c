// Shim: the three boost::filesystem names line3D.cc uses (line3D.cc:60-61, 302-303), on top of <sys/stat.h>.
// Test infrastructure only (oracle/Makefile).
#pragma once
#include <string>
#include <sys/stat.h>
namespace boost { namespace filesystem {
class path { std::string s_; public: path() {} path(const std::string& s) : s_(s) {} const std::string& string() const { return s_; } };
inline bool exists(const path& p) { struct stat st; return ::stat(p.string().c_str(), &st) == 0; }
inline bool create_directory(const path& p) { return ::mkdir(p.string().c_str(), 0777) == 0; }
}}
