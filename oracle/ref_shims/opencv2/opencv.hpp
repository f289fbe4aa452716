// Test-only stand-in for the slice of OpenCV's C++ API that the reference's rfimage.h uses
// (OpenCV C++ is not installed here).  It exists so rfimage.h can be
// compiled where it lies under /root/reference into oracle/_ref/ and its add_echo / convolve /
// envelope / create_mapping used as known-answer generators.  cv::remap itself is NOT restated
// here (it records its arguments only); scan conversion is pinned against the real cv2.remap from
// opencv-python-headless by tests/golden/make_golden.py instead.
#ifndef ORACLE_SHIM_OPENCV_HPP
#define ORACLE_SHIM_OPENCV_HPP
#include <vector>
#include <string>
#include <cstring>

#define CV_32FC1 5
#define CV_8U 0
#define CV_INTER_LINEAR 1

namespace cv {
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0) { v[0] = a; v[1] = v[2] = v[3] = 0; } };
enum { BORDER_CONSTANT = 0, WINDOW_AUTOSIZE = 1 };

class Mat
{
public:
    int rows, cols;
    std::vector<float> data;
    Mat() : rows(0), cols(0) {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), data((size_t)r * c, 0.0f) {}
    void create(Size s, int /*type*/) { rows = s.height; cols = s.width; data.assign((size_t)rows * cols, 0.0f); }
    Size size() const { return Size(cols, rows); }
    template <typename T> T& at(int r, int c) { return data[(size_t)r * cols + c]; }
    template <typename T> const T& at(int r, int c) const { return data[(size_t)r * cols + c]; }
    template <typename T> T& at(int i) { return data[i]; }
    void setTo(float v) { for (auto& x : data) x = v; }
    void convertTo(Mat& out, int, double scale) const { out = *this; for (auto& x : out.data) x = (float)(x * scale); }
};

inline void minMaxLoc(const Mat& m, double* mn, double* mx)
{
    double a = 1e300, b = -1e300;
    for (float x : m.data) { if (x < a) a = x; if (x > b) b = x; }
    if (mn) *mn = a; if (mx) *mx = b;
}
// records nothing, computes nothing: see header comment
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int, int, const Scalar&) {}
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void namedWindow(const std::string&, int) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int) { return 0; }
}  // namespace cv
#endif
