#include "../imgproc/imgproc.hpp"
