/*
 * cv:: facade -- TEST INFRASTRUCTURE (oracle/_ref), never linked into the product.
 *
 * The image has no OpenCV C++ SDK, only the cv2 Python wheel (OpenCV 4.13.0).  This header gives the reference's
 * UNMODIFIED sources (/root/reference/src/binarizations/binarize*.cpp, src/removeLines.cpp, src/imageLibCommon.cpp)
 * the slice of the cv:: API they use.  A cv::Mat here is a handle on a numpy array; every function that does
 * arithmetic is forwarded (facade.cpp -> oracle/cvfacade/cvcalls.py) to the real OpenCV inside the wheel.  The
 * facade itself computes nothing: it only keeps headers (rows/cols/type/ROI offsets) and reproduces the *lowering*
 * of cv::MatExpr to OpenCV calls, which in real OpenCV lives in modules/core/src/matop.cpp:
 *
 *   A + B                 MatOp_AddEx(a=A, b=B, 1, 1)      -> cv::add(A, B)
 *   A - B                 MatOp_AddEx(A, B, 1, -1)          -> cv::subtract(A, B)
 *   A + k*B               MatOp::add folds the scaled term: MatOp_AddEx(a=B, b=A, alpha=k, beta=1)
 *                                                           -> cv::scaleAdd(B, k, A)            (binarizeNiblack.cpp:108)
 *   k*A                   MatOp_AddEx(A, -, k, 0, s=0)      -> A.convertTo(dst, type, k, 0)    (binarizeFeng.cpp:130-131)
 *   A + s,  A - s         MatOp_AddEx(A, -, 1, 0, s=+-s)    -> cv::add(A, Scalar(+-s))         (binarizeFeng.cpp:142, WolfJolion.cpp:129)
 *   s - A                 MatOp_AddEx(A, -, -1, 0, s)       -> cv::subtract(Scalar(s), A)
 *   A.mul(B [,scale])     MatOp_Bin('*', A, B, scale)       -> cv::multiply(A, B, dst, scale)
 *   A.mul(expr)           MatOp::multiply: a scaled expr (k*B) folds into `scale`, anything else is evaluated first
 *   A > B                 MatOp_Cmp(A, B, CMP_GT)           -> cv::compare(A, B, dst, CMP_GT)
 *   A ^ s, A | B, ~A      MatOp_Bin('^' / '|' / '~')        -> cv::bitwise_xor / bitwise_or / bitwise_not
 *   A -= B, A += B        cv::subtract(A, B, A) / cv::add(A, B, A)   (operations.hpp)
 *   A *= s                A.convertTo(A, -1, s)
 *
 * Output arguments follow Mat::create(): when the destination already has the result's size and type the result is
 * written into its buffer (so headers that share the buffer see it); otherwise the destination is re-bound.
 */
#ifndef PRL_CVFACADE_CORE_HPP
#define PRL_CVFACADE_CORE_HPP

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <exception>
#include <memory>
#include <string>
#include <vector>

struct _object;
typedef struct _object PyObject;

#define CV_EXPORTS
#define CV_EXPORTS_W

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

static inline int cvRound(double v) { return static_cast<int>(std::lrint(v)); }

namespace cv {

typedef unsigned char uchar;

class Exception : public std::exception {
public:
    Exception() : code(0), line(0) {}
    Exception(int code_, const std::string& err_, const std::string& func_, const std::string& file_, int line_)
        : msg(err_), code(code_), err(err_), func(func_), file(file_), line(line_) {}
    virtual ~Exception() throw() {}
    virtual const char* what() const throw() { return msg.c_str(); }
    std::string msg;
    int code;
    std::string err, func, file;
    int line;
};
[[noreturn]] void error(int code, const std::string& err, const char* func, const char* file, int line);
#define CV_Assert(expr) do { if (!(expr)) cv::error(-215, #expr, __func__, __FILE__, __LINE__); } while (0)
#define CV_Error(code, msg) cv::error(code, msg, __func__, __FILE__, __LINE__)

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x(static_cast<T>(p.x)), y(static_cast<T>(p.y)) {}
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point_& o) const { return !(*this == o); }
    Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
    Size_ operator-(const Size_& o) const { return Size_(width - o.width, height - o.height); }
    Size_ operator+(const Size_& o) const { return Size_(width + o.width, height + o.height); }
};
typedef Size_<int> Size;
typedef Size_<float> Size2f;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Size_<T> size() const { return Size_<T>(width, height); }
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    T area() const { return width * height; }
    bool contains(const Point_<T>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
    bool operator==(const Rect_& o) const { return x == o.x && y == o.y && width == o.width && height == o.height; }
};
typedef Rect_<int> Rect;
template <typename T> static inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) {
    T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    T x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
    if (x2 <= x1 || y2 <= y1) return Rect_<T>();
    return Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b) { val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<int, 4> Vec4i;
typedef Vec<float, 2> Vec2f;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
    bool isReal() const { return val[1] == 0 && val[2] == 0 && val[3] == 0; }
    bool operator==(const Scalar& o) const { return val[0] == o.val[0] && val[1] == o.val[1] && val[2] == o.val[2] && val[3] == o.val[3]; }
    Scalar operator-() const { return Scalar(-val[0], -val[1], -val[2], -val[3]); }
};

struct RotatedRect {
    Point2f center;
    Size2f size;
    float angle;
    RotatedRect() : angle(0) {}
    void points(Point2f pts[]) const;
};

template <typename T> using Ptr = std::shared_ptr<T>;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };

class MatExpr;

class Mat {
public:
    /* public data members of cv::Mat the reference reads */
    int flags, dims, rows, cols;
    uchar* data;
    size_t step;

    Mat();
    Mat(const Mat& m);
    Mat(int rows, int cols, int type);
    Mat(Size size, int type);
    Mat(int rows, int cols, int type, const Scalar& s);
    Mat(Size size, int type, const Scalar& s);
    Mat(const Mat& m, const Rect& roi);
    template <typename T> explicit Mat(const std::vector<Point_<T> >&) : Mat() { not_forwarded("Mat(std::vector<Point_<T>>)"); }
    ~Mat();
    Mat& operator=(const Mat& m);
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s);

    Mat operator()(const Rect& roi) const;
    Mat clone() const;
    void copyTo(Mat& dst) const;
    void copyTo(Mat& dst, const Mat& mask) const;
    void release();
    void create(int rows, int cols, int type);
    void create(Size size, int type);
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    int type() const { return flags & 0xFFF; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return static_cast<size_t>(rows) * cols; }
    size_t elemSize() const;
    bool isContinuous() const;
    bool isSubmatrix() const;

    MatExpr mul(const Mat& m, double scale = 1) const;
    MatExpr mul(const MatExpr& e, double scale = 1) const;
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s, const Mat& mask = Mat());

    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + static_cast<size_t>(r) * step + static_cast<size_t>(c) * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + static_cast<size_t>(r) * step + static_cast<size_t>(c) * sizeof(T)); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + static_cast<size_t>(r) * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + static_cast<size_t>(r) * step); }

    static Mat zeros(int rows, int cols, int type);
    static Mat zeros(Size size, int type);
    static Mat ones(int rows, int cols, int type);
    static Mat ones(Size size, int type);

    /* ---- facade internals (not part of cv::Mat) ---- */
    [[noreturn]] static void not_forwarded(const char* what);
    PyObject* arr;   /* the numpy array (a view when this is a sub-matrix); owned reference or NULL */
    PyObject* root;  /* the array `arr` is a view of (== arr when not a sub-matrix); owned reference or NULL */
    int ox, oy;      /* position of this header inside `root` */
    void bind(PyObject* stolen_arr, PyObject* stolen_root = 0, int ox = 0, int oy = 0);
    void store(PyObject* stolen_result);  /* Mat::create() semantics for an output argument */
};

/* the lazy expression OpenCV builds for operators on Mat; only the node kinds the reference produces */
class MatExpr {
public:
    enum Kind { IDENTITY, ADDEX, BIN, CMP };
    Kind kind;
    int op;  /* BIN: '*', '^', '|', '&', '~';  CMP: cv::CMP_*  */
    Mat a, b;
    double alpha, beta;
    Scalar s;
    MatExpr() : kind(IDENTITY), op(0), alpha(1), beta(1) {}
    explicit MatExpr(const Mat& m) : kind(IDENTITY), op(0), a(m), alpha(1), beta(1) {}
    operator Mat() const;
    void assign(Mat& dst) const;
    bool scaled() const { return kind == ADDEX && (b.data == 0 || beta == 0) && s == Scalar(); }
    MatExpr mul(const Mat& m, double scale = 1) const;
    MatExpr mul(const MatExpr& e, double scale = 1) const;
    Size size() const { return a.size(); }
};

MatExpr operator+(const Mat& a, const Mat& b);
MatExpr operator+(const Mat& a, const MatExpr& e);
MatExpr operator+(const MatExpr& e, const Mat& b);
MatExpr operator+(const MatExpr& e1, const MatExpr& e2);
MatExpr operator+(const Mat& a, const Scalar& s);
MatExpr operator+(const Scalar& s, const Mat& a);
MatExpr operator-(const Mat& a, const Mat& b);
MatExpr operator-(const Mat& a, const Scalar& s);
MatExpr operator-(const Scalar& s, const Mat& a);
MatExpr operator-(const MatExpr& e, const Mat& b);
MatExpr operator*(const Mat& a, double s);
MatExpr operator*(double s, const Mat& a);
MatExpr operator>(const Mat& a, const Mat& b);
MatExpr operator<(const Mat& a, const Mat& b);
MatExpr operator^(const Mat& a, const Scalar& s);
MatExpr operator|(const Mat& a, const Mat& b);
MatExpr operator~(const Mat& a);
Mat& operator-=(Mat& a, const Mat& b);
Mat& operator+=(Mat& a, const Mat& b);
Mat& operator+=(Mat& a, const MatExpr& e);
Mat& operator*=(Mat& a, double s);
Mat& operator/=(Mat& a, double s);

/* ---- core ---- */
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
Scalar mean(const Mat& src);
void sqrt(const Mat& src, Mat& dst);
void pow(const Mat& src, double power, Mat& dst);
void minMaxLoc(const Mat& src, double* minVal, double* maxVal = 0, Point* minLoc = 0, Point* maxLoc = 0,
               const Mat& mask = Mat());
void add(const Mat& a, const Mat& b, Mat& dst);
void subtract(const Mat& a, const Mat& b, Mat& dst);
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1);
void divide(const Mat& a, const Mat& b, Mat& dst, double scale = 1);
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst);
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst);
void compare(const Mat& a, const Mat& b, Mat& dst, int cmpop);
void bitwise_not(const Mat& src, Mat& dst);
void bitwise_or(const Mat& a, const Mat& b, Mat& dst);
void inRange(const Mat& src, const Scalar& lo, const Scalar& hi, Mat& dst);
void split(const Mat& src, std::vector<Mat>& mv);
void merge(const std::vector<Mat>& mv, Mat& dst);

}  // namespace cv

#endif
