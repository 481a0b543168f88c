// Shim (see nvp.hpp)
#pragma once
#include <boost/serialization/nvp.hpp>
