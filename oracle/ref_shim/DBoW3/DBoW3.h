// DBoW3/DBoW3.h — stand-in (TEST INFRASTRUCTURE, see ../Eigen/Core).  The reference's GlobalTypes.h pulls DBoW3, which pulls
// OpenCV; the hot path uses neither, but the headers mention a few of their types.  Only what those mentions need.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
template <class T, int N> struct Vec { T v[N]; Vec() {} Vec(T a, T b, T c) { v[0] = a; v[1] = b; v[2] = c; } T& operator[](int i) { return v[i]; } };
typedef Vec<unsigned char, 3> Vec3b;
struct Point2f { float x = 0, y = 0; };
struct Point { int x = 0, y = 0; Point() {} Point(int a, int b) : x(a), y(b) {} };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
typedef unsigned char uchar;
template <class T, class U> inline T saturate_cast(U v) { return (T)v; }
enum { BORDER_REFLECT_101 = 4, INTER_LINEAR = 1 };
struct WindowFlags { enum { WINDOW_KEEPRATIO = 0 }; };
struct KeyPoint { Point2f pt; float size = 0, angle = 0, response = 0; int octave = 0, class_id = 0; };
struct Mat {
  int rows = 0, cols = 0;
  std::vector<unsigned char> buf;
  Mat() {}
  Mat(int r, int c, int, void*) : rows(r), cols(c) {}
  Mat(Size sz, int) : rows(sz.height), cols(sz.width) {}
  Mat operator()(const Rect& r) const { Mat m; m.rows = r.height; m.cols = r.width; return m; }   // (no pixels: the image pyramids of
  int type() const { return 0; }                                                                   //  the feature front-end are not on the path)
  void convertTo(Mat&, int) const {}
  bool empty() const { return rows == 0 || cols == 0; }
  void copyTo(Mat& o) const { o = *this; }
  template <class T> T& at(int, int) { static T t; return t; }
  template <class T> const T* ptr() const { return reinterpret_cast<const T*>(buf.data()); }
};
inline void resize(const Mat&, Mat&, Size, double, double, int) {}
inline void copyMakeBorder(const Mat&, Mat&, int, int, int, int, int) {}
inline void namedWindow(const char*, int) {}
inline void imshow(const char*, const Mat&) {}
inline int waitKey(int) { return 0; }
}  // namespace cv
using cv::uchar;
using cv::copyMakeBorder;
#define CV_32F 5
#define CV_8U 0
inline int cvRound(double v) { return (int)(v + (v >= 0 ? 0.5 : -0.5)); }
namespace DBoW3 {
class Vocabulary {};
class BowVector {};
class FeatureVector {};
}  // namespace DBoW3
