// Shim: boost::mutex on top of std::mutex (Boost is not installed in this image).
#pragma once
#include <mutex>
namespace boost { class mutex { public: void lock(){m_.lock();} void unlock(){m_.unlock();} private: std::mutex m_; }; }
