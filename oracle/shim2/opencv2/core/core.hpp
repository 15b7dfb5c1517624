// TEST INFRASTRUCTURE ONLY — container-level stand-in for the OpenCV core types the reference's matcher / tracker code touches
// (cv::Mat as a strided byte matrix, Point_/Point3_, KeyPoint, DMatch, InputArray/OutputArray), so that the reference's OWN
// statements (src/utils/framematcher.cpp, the functions oracle/gen_ref_extract.py cuts out of map.cpp / system.cpp / misc.cpp /
// frame.h / mappoint.h) compile in a container without OpenCV C++ headers.  The few arithmetic members follow OpenCV's published
// definitions (opencv2/core/types.hpp): Point3_::dot and operator*= in the element type via saturate_cast, cv::norm(Point3_) in double.
#pragma once
#include <algorithm>
#include <bitset>
#include <cassert>
#include <cmath>
#include <math.h>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_MAJOR_VERSION 2   /* keeps the optional OpenCV-3 feature extractors of the reference out of the build */
typedef unsigned char uchar;
typedef unsigned int uint;
namespace cv {
template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return (int)lrint(v); }
template <typename T> struct Point_ {
    T x, y;
    Point_(T a = 0, T b = 0) : x(a), y(b) {}
    template <typename U> Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_(T a = 0, T b = 0, T c = 0) : x(a), y(b), z(c) {}
    T dot(const Point3_& p) const { return saturate_cast<T>(x * p.x + y * p.y + z * p.z); }
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
template <typename T> static inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) {
    return Point3_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y), saturate_cast<T>(a.z - b.z));
}
template <typename T> static inline Point3_<T>& operator*=(Point3_<T>& a, double b) {
    a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); a.z = saturate_cast<T>(a.z * b);
    return a;
}
template <typename T> static inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }
struct KeyPoint {                 // 28 bytes
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
struct DMatch {                   // 16 bytes
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = std::numeric_limits<float>::max();
};
struct Size {
    int width = 0, height = 0;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
    int area() const { return width * height; }
};
struct MatStep {
    size_t p[2] = {0, 0};
    size_t operator[](int i) const { return p[i]; }
};
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    MatStep step;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* d, size_t st = 0) : rows(r), cols(c), data((unsigned char*)d), _type(type) {
        step.p[0] = st ? st : (size_t)c * elemSize(); step.p[1] = elemSize();
    }
    void create(int r, int c, int type) {
        if (rows == r && cols == c && _type == type && data) return;
        rows = r; cols = c; _type = type;
        step.p[0] = (size_t)c * elemSize(); step.p[1] = elemSize();
        _own = std::shared_ptr<unsigned char>(new unsigned char[std::max<size_t>(1, step.p[0] * r)](), std::default_delete<unsigned char[]>());
        data = _own.get();
    }
    void release() { rows = cols = 0; data = nullptr; _own.reset(); }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.at<float>(i, i) = 1; return m; }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& m) const {            // into a same-sized view (a row of another matrix) or a fresh matrix
        if (!(m.rows == rows && m.cols == cols && m._type == _type && m.data)) m.create(rows, cols, _type);
        for (int r = 0; r < rows; r++) memcpy(m.data + r * m.step.p[0], data + r * step.p[0], (size_t)cols * elemSize());
    }
    void copyTo(Mat&& m) const { Mat& ref = m; copyTo(ref); }
    void resize(size_t n) { if ((int)n < rows) rows = (int)n; else if ((int)n > rows) { Mat m(n, cols, _type); for (int r = 0; r < rows; r++) memcpy(m.data + r * m.step.p[0], data + r * step.p[0], (size_t)cols * elemSize()); *this = m; } }
    Mat row(int r) const { Mat m(1, cols, _type, data + (size_t)r * step.p[0], step.p[0]); m._own = _own; return m; }
    void convertTo(Mat& m, int type) const {
        Mat o(rows, cols, type);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                double v = _type == CV_32F ? (double)at<float>(r, c) : _type == CV_64F ? at<double>(r, c) : _type == CV_32S ? (double)at<int>(r, c) : (double)at<uchar>(r, c);
                if (type == CV_32F) o.at<float>(r, c) = (float)v; else if (type == CV_64F) o.at<double>(r, c) = v; else if (type == CV_32S) o.at<int>(r, c) = (int)lrint(v); else o.at<uchar>(r, c) = (uchar)v;
            }
        m = o;
    }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.p[0] + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step.p[0] + (size_t)c * sizeof(T)); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.p[0]); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.p[0]); }
    int type() const { return _type; }
    size_t elemSize() const { return _type == CV_32F || _type == CV_32S ? 4 : _type == CV_64F ? 8 : 1; }
    size_t elemSize1() const { return elemSize(); }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step.p[0] == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }
private:
    int _type = 0;
    std::shared_ptr<unsigned char> _own;
};
// CV_32F matrix product (double accumulator, like cv::gemm's small-matrix path), as far as the adapters use it (4x4 pose products)
static inline Mat operator*(const Mat& a, const Mat& b) {
    Mat o(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            double v = 0;
            for (int k = 0; k < a.cols; k++) v += (double)a.at<float>(r, k) * (double)b.at<float>(k, c);
            o.at<float>(r, c) = (float)v;
        }
    return o;
}
// InputArray / OutputArray: thin handles on a Mat, as far as the extractor seam uses them (getMat / create / release)
class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat& m) : _m(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return _m ? *_m : Mat(); }
protected:
    Mat* _m = nullptr;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) { _m = &m; }
    void create(int r, int c, int type) const { _m->create(r, c, type); }
    void release() const { _m->release(); }
    Mat& getMatRef() const { return *_m; }
};
inline Mat operator-(const Mat& a, const Mat& b) {   // float matrices only (descriptor rows of float-descriptor extractors)
    Mat o(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c);
    return o;
}
inline double norm(const Mat& m) {
    double s = 0;
    for (int r = 0; r < m.rows; r++) for (int c = 0; c < m.cols; c++) s += (double)m.at<float>(r, c) * m.at<float>(r, c);
    return std::sqrt(s);
}
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline InputArray noArray() { static _InputArray none; return none; }
}  // namespace cv
