// Stand-in for <opencv2/imgproc/imgproc.hpp> (test infrastructure): the shim needs nothing from it.
#pragma once
#include "opencv2/core/core.hpp"
