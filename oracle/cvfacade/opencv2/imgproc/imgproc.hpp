/* cv:: facade, imgproc slice -- TEST INFRASTRUCTURE (see opencv2/core/core.hpp).  Every function forwards to the
 * cv2 wheel through oracle/cvfacade/cvcalls.py. */
#ifndef PRL_CVFACADE_IMGPROC_HPP
#define PRL_CVFACADE_IMGPROC_HPP

#include "../core/core.hpp"

namespace cv {

enum { COLOR_BGR2BGRA = 0, COLOR_BGRA2BGR = 1, COLOR_BGR2RGB = 4, COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7,
       COLOR_GRAY2BGR = 8, COLOR_GRAY2RGB = 8, COLOR_BGRA2GRAY = 10, COLOR_RGBA2GRAY = 11,
       COLOR_BGR2HSV = 40, COLOR_HSV2BGR = 54, COLOR_BGR2Lab = 44, COLOR_Lab2BGR = 56,
       COLOR_BGR2YCrCb = 36, COLOR_YCrCb2BGR = 38, COLOR_BGR2HLS = 52, COLOR_HLS2BGR = 60 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3, THRESH_TOZERO_INV = 4,
       THRESH_MASK = 7, THRESH_OTSU = 8, THRESH_TRIANGLE = 16 };
enum { ADAPTIVE_THRESH_MEAN_C = 0, ADAPTIVE_THRESH_GAUSSIAN_C = 1 };
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { RETR_EXTERNAL = 0, RETR_LIST = 1, RETR_CCOMP = 2, RETR_TREE = 3 };
enum { CHAIN_APPROX_NONE = 1, CHAIN_APPROX_SIMPLE = 2 };
enum { FILLED = -1 };

void cvtColor(const Mat& src, Mat& dst, int code, int dstCn = 0);
void integral(const Mat& src, Mat& sum, Mat& sqsum, int sdepth = -1, int sqdepth = -1);
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0,
              int borderType = BORDER_DEFAULT);
void dilate(const Mat& src, Mat& dst, const Mat& kernel, Point anchor = Point(-1, -1), int iterations = 1);
void erode(const Mat& src, Mat& dst, const Mat& kernel, Point anchor = Point(-1, -1), int iterations = 1);
double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type);
void adaptiveThreshold(const Mat& src, Mat& dst, double maxValue, int adaptiveMethod, int thresholdType, int blockSize,
                       double C);
Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1));
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void medianBlur(const Mat& src, Mat& dst, int ksize);
void bilateralFilter(const Mat& src, Mat& dst, int d, double sigmaColor, double sigmaSpace, int borderType = BORDER_DEFAULT);
void Canny(const Mat& image, Mat& edges, double threshold1, double threshold2, int apertureSize = 3, bool L2gradient = false);
void equalizeHist(const Mat& src, Mat& dst);
void findContours(const Mat& image, std::vector<std::vector<Point> >& contours, std::vector<Vec4i>& hierarchy, int mode,
                  int method, Point offset = Point());
Rect boundingRect(const std::vector<Point>& points);
double contourArea(const std::vector<Point>& contour, bool oriented = false);
/* declared so that src/imageLibCommon.cpp compiles unmodified; not on the binarization path, never forwarded */
void convexHull(const std::vector<Point>& points, std::vector<Point>& hull, bool clockwise = false, bool returnPoints = true);
void convexHull(const Mat& points, std::vector<Point2f>& hull, bool clockwise = false, bool returnPoints = true);
void convexHull(const std::vector<Point>& points, std::vector<int>& hull, bool clockwise = false, bool returnPoints = true);
RotatedRect minAreaRect(const std::vector<Point>& points);
void calcHist(const Mat* images, int nimages, const int* channels, const Mat& mask, Mat& hist, int dims, const int* histSize,
              const float** ranges, bool uniform = true, bool accumulate = false);

class CLAHE {
public:
    CLAHE() : clip(40.0), tiles(8, 8) {}
    void setClipLimit(double c) { clip = c; }
    double getClipLimit() const { return clip; }
    void setTilesGridSize(Size s) { tiles = s; }
    void apply(const Mat& src, Mat& dst);
private:
    double clip;
    Size tiles;
};
Ptr<CLAHE> createCLAHE(double clipLimit = 40.0, Size tileGridSize = Size(8, 8));

}  // namespace cv

#endif
