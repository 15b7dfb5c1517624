// TEST INFRASTRUCTURE ONLY — stand-in for src/ucoslamtypes.h: the descriptor-type tags (ucoslamtypes.h:39-43, values are part of the
// stream format) without the cv::FileStorage-based Params class.
#pragma once
#include <cstdint>
#include <string>
#include <opencv2/core/core.hpp>
#include "ucoslam_exports.h"
namespace ucoslam {
class DescriptorTypes {
public:
    enum Type : std::int8_t { DESC_NONE = 0, DESC_ORB = 1, DESC_AKAZE = 2, DESC_BRISK = 3, DESC_FREAK = 4, DESC_SURF = 5 };
};
}
