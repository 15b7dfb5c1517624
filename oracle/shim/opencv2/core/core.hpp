// TEST INFRASTRUCTURE ONLY — minimal stand-in for <opencv2/core/core.hpp>, just enough of cv::Mat for the reference's
// 3rdparty/fbow/fbow/fbow.{h,cpp} to compile unchanged in a container without OpenCV C++ headers.
#pragma once
#include <cassert>
#include <cstddef>
#include <cmath>
#include <math.h>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_32F 5
#include <memory>
#include <algorithm>   // the real OpenCV headers pull these in; the reference relies on that (keyframedatabase.cpp:263-266)
typedef unsigned char uchar;
namespace cv {
struct Point2f {
    float x, y;
    Point2f(float a = 0, float b = 0) : x(a), y(b) {}
};
struct Point3f {
    float x, y, z;
    Point3f(float a = 0, float b = 0, float c = 0) : x(a), y(b), z(c) {}
};
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) : rows(r), cols(c), _type(type) {   // owning (typesg2o.h: cv::Mat cvMat(4,4,CV_32F))
        _step = (size_t)c * elemSize();
        _own = std::shared_ptr<unsigned char>(new unsigned char[_step * r](), std::default_delete<unsigned char[]>());
        _data = _own.get();
    }
    Mat clone() const {
        Mat m(rows, cols, _type);
        for (int r = 0; r < rows; r++) memcpy(m._data + r * m._step, _data + r * _step, (size_t)cols * elemSize());
        return m;
    }
    template <typename T> T& at(int r, int c) { return *(T*)(_data + (size_t)r * _step + (size_t)c * sizeof(T)); }
    Mat(int r, int c, int type, void* data, size_t step = 0) : rows(r), cols(c), _type(type), _data((unsigned char*)data) {
        _step = step ? step : (size_t)c * elemSize();
    }
    int type() const { return _type; }
    size_t elemSize() const { return _type == CV_32FC1 ? 4 : 1; }
    size_t elemSize1() const { return elemSize(); }
    template <typename T> T* ptr(int r = 0) { return (T*)(_data + (size_t)r * _step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(_data + (size_t)r * _step); }
    bool empty() const { return rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return _step == (size_t)cols * elemSize(); }
private:
    int _type = 0;
    unsigned char* _data = nullptr;
    size_t _step = 0;
    std::shared_ptr<unsigned char> _own;
};
}  // namespace cv
