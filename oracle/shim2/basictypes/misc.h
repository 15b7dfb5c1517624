// TEST INFRASTRUCTURE ONLY — stand-in for src/basictypes/misc.h: the declarations the matcher code uses; the definitions of the
// match filters are the REFERENCE's own (misc.cpp:105-185, cut out into gen/misc_filters.inc and compiled by the wrappers),
// epipolarLineSqDist is its own inline (misc.h:72-81).  computeF12 (misc.cpp:893-920, OpenCV matrix algebra) is supplied by the
// wrapper from the caller's matrix: the fundamental matrix is an INPUT of the stage under test.
#pragma once
#include "map_types/frame.h"
#include <opencv2/features2d/features2d.hpp>
#include <vector>
#include "debug.h"
namespace ucoslam {
void filter_ambiguous_train(std::vector<cv::DMatch>& matches_io);
void filter_ambiguous_query(std::vector<cv::DMatch>& matches_io);
void remove_unused_matches(std::vector<cv::DMatch>& matches_io);
void remove_bad_matches(std::vector<cv::DMatch>& matches_io, const vector<bool>& vBadMatches);
cv::Mat computeF12(const cv::Mat& RT1, const cv::Mat& CameraMatrix1, const cv::Mat& RT2, const cv::Mat& CameraMatrix2 = cv::Mat());
#include "gen/misc_epipolar.inc"
}
