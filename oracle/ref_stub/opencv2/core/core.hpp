// TEST INFRASTRUCTURE ONLY -- part of the OpenCV API stand-in used to compile the unmodified
// reference sources into oracle/_ref (see opencv2/opencv.hpp in this directory).
#include <opencv2/opencv.hpp>
