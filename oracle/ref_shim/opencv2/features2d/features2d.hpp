// TEST INFRASTRUCTURE (oracle/_ref build only): forwards to the minimal OpenCV stand-in, see opencv2/core/core.hpp.
#pragma once
#include <opencv2/core/core.hpp>
