// Shim: boost/filesystem is included by serialization.h but unused on the oracle path.
#pragma once
