// prl_binarize_cuda.h -- the reference's cv::Mat-in / cv::Mat-out binarize* entry points, bodies
// replaced by libprlib_cuda (include/prlib_cuda.h).  Signatures, default arguments and thrown
// exception types are those of the reference headers:
//   binarizeSauvola.h:43-47, binarizeNiblack.h:43-47, binarizeWolfJolion.h:43-47,
//   binarizeNICK.h:43-47, binarizeFeng.h:46-53, binarizeLocalOtsu.h:50-57 (rect-loop core only).
// A caller that includes the reference's own headers links against this translation unit instead of
// src/binarizations/binarize{Sauvola,Niblack,WolfJolion,NICK,Feng}.cpp and sees no difference:
//   * std::invalid_argument for an empty image or a window that is not (>1 and odd);
//   * cv::Exception when the processing rectangle is empty (WJ/NICK/Feng on min(rows,cols) <= window);
//   * the INPUT Mat is left holding the replicate-padded grayscale image, as the reference leaves it
//     (binarizeSauvola.cpp:49-52, :65) -- define PRL_CUDA_NO_INPUT_SIDE_EFFECT to skip that copy.
// There is no CPU fallback: any CUDA failure surfaces as std::runtime_error.
#ifndef PRL_BINARIZE_CUDA_H
#define PRL_BINARIZE_CUDA_H

#include <opencv2/core/core.hpp>
#include <vector>

namespace prl
{
CV_EXPORTS void binarizeSauvola(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize = 101,
                                double thresholdCoefficient = 0.01, int morphIterationCount = 2);
CV_EXPORTS void binarizeNiblack(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize = 101,
                                double thresholdCoefficient = 0.01, int morphIterationCount = 2);
CV_EXPORTS void binarizeWolfJolion(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize = 101,
                                   double thresholdCoefficient = 0.01, int morphIterationCount = 2);
CV_EXPORTS void binarizeNICK(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize = 21,
                             double thresholdCoefficient = -0.01, int morphIterationCount = 0);
CV_EXPORTS void binarizeFeng(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize = 21,
                             double thresholdCoefficient_alpha1 = 0.75, double thresholdCoefficient_k1 = 0.2,
                             double thresholdCoefficient_k2 = 0.03, double thresholdCoefficient_gamma = 2.0,
                             int morphIterationCount = 2);

// Batch forms (no counterpart in the reference, whose API is one cv::Mat per call): the same five functions over a
// vector of pages through the pinned-memory batch loader + page dispatcher of libprlib_cuda (prl_cuda_binarize_batch:
// every visible GPU, H2D / kernels / D2H overlapped).  outputs[i] is what binarizeX(inputs[i], ...) returns; the inputs
// are NOT modified (no padded-image side effect).  Runs of consecutive single-channel pages of one size go through the
// batch loader together, anything else (3/4 channels) through the single-image path; exceptions as above.
CV_EXPORTS void binarizeSauvolaBatch(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize = 101,
                                     double thresholdCoefficient = 0.01, int morphIterationCount = 2);
CV_EXPORTS void binarizeNiblackBatch(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize = 101,
                                     double thresholdCoefficient = 0.01, int morphIterationCount = 2);
CV_EXPORTS void binarizeWolfJolionBatch(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize = 101,
                                        double thresholdCoefficient = 0.01, int morphIterationCount = 2);
CV_EXPORTS void binarizeNICKBatch(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize = 21,
                                  double thresholdCoefficient = -0.01, int morphIterationCount = 0);
CV_EXPORTS void binarizeFengBatch(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize = 21,
                                  double thresholdCoefficient_alpha1 = 0.75, double thresholdCoefficient_k1 = 0.2,
                                  double thresholdCoefficient_k2 = 0.03, double thresholdCoefficient_gamma = 2.0,
                                  int morphIterationCount = 2);

// The per-contour-rectangle Otsu loop of prl::binarizeLocalOtsu (binarizeLocalOtsu.cpp:138-162):
// `gray` is imageToProc, xywh holds cv::boundingRect(contour) as x, y, width, height quadruples.
CV_EXPORTS void binarizeLocalOtsuRects(const cv::Mat& gray, const std::vector<int>& xywh, cv::Mat& binarized,
                                       double maxValue = 255);
// cv::threshold(src, dst, 128, maxValue, THRESH_BINARY | THRESH_OTSU) (deskew.cpp:224, removeLines.cpp:45);
// returns the threshold like cv::threshold does.
CV_EXPORTS double thresholdOtsu(const cv::Mat& src, cv::Mat& dst, double maxValue = 255);
// prl::binarizeLocalOtsu with the reference's signature and defaults (binarizeLocalOtsu.h:50-57); blur, Otsu value, Canny,
// morphology, top-level contour rectangles and the rectangle loop all run on the device, and so does the optional
// CLAHE + equalizeHist pre-step (CLAHEClipLimit > 0, binarizeLocalOtsu.cpp:79-82).
CV_EXPORTS void binarizeLocalOtsu(cv::Mat& inputImage, cv::Mat& outputImage, double maxValue = 255.0,
                                  double CLAHEClipLimit = 0.0, int GaussianBlurKernelSize = 19,
                                  double CannyUpperThresholdCoeff = 0.15, double CannyLowerThresholdCoeff = 0.01,
                                  int CannyMorphIters = 1);
// prl::removeLines (removeLines.h, removeLines.cpp:30-77), 1- or 3-channel (BGR) input.
CV_EXPORTS void removeLines(const cv::Mat& inputImage, cv::Mat& outputImage);
// The edge map prl::binarizeLocalOtsu feeds to cv::findContours: CannyEdgeDetection(imageToProc, resultCanny, ...)
// (imageLibCommon.cpp:244-324) followed by cv::dilate(resultCanny, ..., postDilate = 3) (binarizeLocalOtsu.cpp:88-92),
// single-channel input, same std::invalid_argument checks as the reference.
CV_EXPORTS void localOtsuEdges(const cv::Mat& imageToProc, cv::Mat& resultCanny, int GaussianBlurKernelSize = 19,
                               double CannyUpperThresholdCoeff = 0.15, double CannyLowerThresholdCoeff = 0.01,
                               int CannyMorphIters = 1, int postDilate = 3);

// ---- the adaptive-mean family (binarizeNativeAdaptive.h:63-75, binarizeAT.h:33-34, binarizeAGT.h:32-33, binarizeGAT.h:33-35,
// binarizePureAdaptive.h:33-34, binarizePureAdaptiveGaussian.h:33-34): same signatures and defaults.  The reference's own
// behaviour is kept, quirks included (tests/test_adaptive.py runs the reference's own object code and shows the same):
//   * binarizeAT / binarizeAGT / binarizePureAdaptiveGaussian only work on 3/4-channel input; a 1-channel image ends in the
//     cv::Exception of cv::adaptiveThreshold on the empty Mat the reference passes it;
//   * binarizeGAT and binarizePureAdaptive end in that cv::Exception for every non-empty input;
//   * binarizeNativeAdaptive converts its INPUT Mat to gray in place (binarizeNativeAdaptive.cpp:58-61); its optional
//     bilateral filter of the mask (bilateralFilterBlockSize >= 3, off by default) runs on the device too.
CV_EXPORTS void binarizeNativeAdaptive(cv::Mat& inputImage, cv::Mat& outputImage, bool isGaussianBlurReqiured = 0,
                                       int medianBlurKernelSize = 5, int GaussianBlurKernelSize = 7, double GaussianBlurSigma = 150.0,
                                       bool isAdaptiveThresholdCalculatedByGaussian = true, double adaptiveThresholdingMaxValue = 255.0,
                                       int adaptiveThresholdingBlockSize = 19, double adaptiveThresholdingShift = 9,
                                       int bilateralFilterBlockSize = 0, double bilateralFilterColorSigma = 150.0,
                                       double bilateralFilterSpaceSigma = 150.0);
CV_EXPORTS void binarizeAT(const cv::Mat& inputImage, cv::Mat& outputImage, const int medianKernelSize, const double maxValue,
                           const int blockSize, const int shift);
CV_EXPORTS void binarizeAGT(const cv::Mat& inputImage, cv::Mat& outputImage, const int medianKernelSize, const double maxValue,
                            const int blockSize, const int shift);
CV_EXPORTS void binarizeGAT(const cv::Mat& inputImage, cv::Mat& outputImage, const int gaussianKernelSize, const double sigmaX,
                            const double sigmaY, const double maxValue, const int blockSize, const int shift);
CV_EXPORTS void binarizePureAdaptive(const cv::Mat& inputImage, cv::Mat& outputImage, const double maxValue, const int blockSize,
                                     const int shift);
CV_EXPORTS void binarizePureAdaptiveGaussian(const cv::Mat& inputImage, cv::Mat& outputImage, const double maxValue,
                                             const int blockSize, const int shift);
}  // namespace prl

#endif  // PRL_BINARIZE_CUDA_H
