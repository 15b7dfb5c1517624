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
struct KeyPoint {                 // opencv2/core/types.hpp: 28 bytes
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
struct DMatch {                   // 16 bytes
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0;
};
struct Size {
    int width = 0, height = 0;
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};
struct MatStep {
    size_t p[2] = {0, 0};
    size_t operator[](int i) const { return p[i]; }
};
struct Point3f {
    float x, y, z;
    Point3f(float a = 0, float b = 0, float c = 0) : x(a), y(b), z(c) {}
};
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    MatStep step;
    Mat(int r, int c, int type) : rows(r), cols(c), _type(type) {   // owning (typesg2o.h: cv::Mat cvMat(4,4,CV_32F))
        _step = (size_t)c * elemSize();
        step.p[0] = _step; step.p[1] = elemSize();
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
        this->step.p[0] = _step; this->step.p[1] = elemSize();
    }
    Size size() const { Size s; s.width = cols; s.height = rows; return s; }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(_data + (size_t)r * _step + (size_t)c * sizeof(T)); }
    void convertTo(Mat& m, int type) const {                       // 8U / 32F only, enough for the adapters' syntax check
        Mat o(rows, cols, type);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                const double v = _type == CV_32FC1 ? (double)*(const float*)(_data + r * _step + c * 4) : (double)_data[r * _step + c];
                if (type == CV_32FC1) *(float*)(o._data + r * o._step + c * 4) = (float)v; else o._data[r * o._step + c] = (unsigned char)v;
            }
        m = o;
    }
    Mat reshape(int /*cn*/, int new_rows) const {                   // continuous matrices only
        Mat o = *this;
        if (new_rows > 0 && total()) { o.cols = (int)(total() / new_rows); o.rows = new_rows; o._step = (size_t)o.cols * elemSize(); o.step.p[0] = o._step; }
        return o;
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
