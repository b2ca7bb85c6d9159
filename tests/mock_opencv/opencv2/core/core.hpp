// Minimal stand-in for <opencv2/core/core.hpp> -- TEST INFRASTRUCTURE ONLY.
// This image has no OpenCV C++ SDK, so the cv::Mat-facing shim (prlib_b200/shim) is compiled and
// exercised against this header: just enough of cv::Mat / cv::Exception for the shim's needs
// (8-bit matrices, reference-counted storage, create / clone / assignment).  With a real OpenCV
// installation the shim compiles unchanged against the real headers.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#define CV_EXPORTS
#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

// same constructor as the real cv::Exception (core.hpp): code, error text, function, file, line
class Exception : public std::runtime_error {
public:
    Exception(int _code, const std::string& _err, const std::string& _func, const std::string& _file, int _line)
        : std::runtime_error(_err), code(_code), err(_err), func(_func), file(_file), line(_line) {}
    int code; std::string err, func, file; int line;
};

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type)
    {
        if (r == rows && c == cols && type == type_ && data) return;
        rows = r; cols = c; type_ = type;
        step = (size_t)c * channels() * elemSize1();
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[step * (r > 0 ? r : 1) + 1], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    size_t elemSize1() const { return depth() < 2 ? 1 : (depth() < 4 ? 2 : (depth() < 6 ? 4 : 8)); }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    bool isContinuous() const { return step == (size_t)cols * channels() * elemSize1(); }
    unsigned char* ptr(int y = 0) { return data + (size_t)y * step; }
    const unsigned char* ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.ptr(y), ptr(y), (size_t)cols * channels() * elemSize1());
        return m;
    }

private:
    int type_ = CV_8UC1;
    std::shared_ptr<unsigned char> buf_;
};

}  // namespace cv
