// prl_binarize_cuda.cpp -- host shim: prl::binarize*(cv::Mat&, cv::Mat&, ...) over the C-ABI of
// libprlib_cuda.  Everything numerical happens on the GPU; this file only validates, converts the
// Mat headers to pointers/strides and reproduces the reference's observable side effects.
#include "prl_binarize_cuda.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/prlib_cuda.h"

namespace
{
// one context per host thread (the reference functions are re-entrant; so are these)
struct ThreadCtx {
    prl_cuda_ctx* ctx = nullptr;
    ~ThreadCtx() { prl_cuda_destroy(ctx); }
};

prl_cuda_ctx* context()
{
    static thread_local ThreadCtx t;
    if (!t.ctx) {
        int rc = prl_cuda_create(0, &t.ctx);
        if (rc != PRL_OK)
            throw std::runtime_error(std::string("libprlib_cuda: ") + prl_cuda_last_error(nullptr));
    }
    return t.ctx;
}

// cv::Exception as OpenCV itself would raise it (CV_Assert -> code -215)
cv::Exception cvError(const std::string& msg) { return cv::Exception(-215, msg, "prl (libprlib_cuda)", __FILE__, __LINE__); }

void check(prl_cuda_ctx* c, int rc)
{
    if (rc == PRL_OK) return;
    const std::string msg = prl_cuda_last_error(c);
    if (rc == PRL_E_INVALID) throw std::invalid_argument(msg);
    if (rc == PRL_E_EMPTY_ROI) throw cvError(msg);            // cv::Mat::operator()(Rect) would assert
    throw std::runtime_error("libprlib_cuda: " + msg);
}

// what the reference leaves in `imageInput`: copyMakeBorder(gray, h, h, h, h, BORDER_REPLICATE)
cv::Mat padReplicate(const cv::Mat& g, int h)
{
    cv::Mat p(g.rows + 2 * h, g.cols + 2 * h, CV_8UC1);
    for (int Y = 0; Y < p.rows; ++Y) {
        int y = Y - h; y = y < 0 ? 0 : (y >= g.rows ? g.rows - 1 : y);
        unsigned char* d = p.ptr(Y);
        const unsigned char* s = g.ptr(y);
        std::memset(d, s[0], (size_t)h);
        std::memcpy(d + h, s, (size_t)g.cols);
        std::memset(d + h + g.cols, s[g.cols - 1], (size_t)h);
    }
    return p;
}

void runLocal(int method, cv::Mat& imageInput, cv::Mat& outputImage, int windowSize, const double params[4], int morph)
{
    if (imageInput.empty())
        throw std::invalid_argument("Input image for binarization is empty");
    if (!((windowSize > 1) && ((windowSize % 2) == 1)))
        throw std::invalid_argument("Window size must satisfy the following condition: "
                                    "( (windowSize > 1) && ((windowSize % 2) == 1) ) ");
    prl_cuda_ctx* c = context();
    const int ch = imageInput.channels();
    if (ch != 1 && ch != 3 && ch != 4) throw cvError("cvtColor: unsupported number of channels");
    int orows = 0, ocols = 0;
    check(c, prl_cuda_output_shape(method, imageInput.rows, imageInput.cols, windowSize, &orows, &ocols));
    cv::Mat out(orows, ocols, CV_8UC1);
    cv::Mat gray = imageInput;                       // 1 channel: the input itself
#ifndef PRL_CUDA_NO_INPUT_SIDE_EFFECT
    if (ch != 1) gray = cv::Mat(imageInput.rows, imageInput.cols, CV_8UC1);
    uint8_t* gray_out = ch != 1 ? gray.data : nullptr;
#else
    uint8_t* gray_out = nullptr;
#endif
    // one call: the (possibly 3/4-channel) image goes up once, cvtColor + both kernels run on the device
    check(c, prl_cuda_binarize_local_image(c, method, imageInput.data, imageInput.rows, imageInput.cols, imageInput.step, ch,
                                           windowSize, params, morph, out.data, out.step, &orows, &ocols, gray_out,
                                           gray_out ? gray.step : 0));
#ifndef PRL_CUDA_NO_INPUT_SIDE_EFFECT
    const int w = windowSize < gray.rows ? (windowSize < gray.cols ? windowSize : gray.cols)
                                         : (gray.rows < gray.cols ? gray.rows : gray.cols);
    imageInput = padReplicate(gray, w / 2);
#endif
    outputImage = out;
}
}  // namespace

void prl::binarizeSauvola(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize, double thresholdCoefficient,
                          int morphIterationCount)
{
    const double p[4] = {thresholdCoefficient, 0, 0, 0};
    runLocal(PRL_SAUVOLA, imageInput, outputImage, windowSize, p, morphIterationCount);
}

void prl::binarizeNiblack(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize, double thresholdCoefficient,
                          int morphIterationCount)
{
    const double p[4] = {thresholdCoefficient, 0, 0, 0};
    runLocal(PRL_NIBLACK, imageInput, outputImage, windowSize, p, morphIterationCount);
}

void prl::binarizeWolfJolion(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize, double thresholdCoefficient,
                             int morphIterationCount)
{
    const double p[4] = {thresholdCoefficient, 0, 0, 0};
    runLocal(PRL_WOLFJOLION, imageInput, outputImage, windowSize, p, morphIterationCount);
}

void prl::binarizeNICK(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize, double thresholdCoefficient,
                       int morphIterationCount)
{
    const double p[4] = {thresholdCoefficient, 0, 0, 0};
    runLocal(PRL_NICK, imageInput, outputImage, windowSize, p, morphIterationCount);
}

void prl::binarizeFeng(cv::Mat& imageInput, cv::Mat& outputImage, int windowSize, double thresholdCoefficient_alpha1,
                       double thresholdCoefficient_k1, double thresholdCoefficient_k2, double thresholdCoefficient_gamma,
                       int morphIterationCount)
{
    const double p[4] = {thresholdCoefficient_alpha1, thresholdCoefficient_k1, thresholdCoefficient_k2,
                         thresholdCoefficient_gamma};
    runLocal(PRL_FENG, imageInput, outputImage, windowSize, p, morphIterationCount);
}

void prl::binarizeLocalOtsuRects(const cv::Mat& gray, const std::vector<int>& xywh, cv::Mat& binarized, double maxValue)
{
    if (gray.empty()) throw std::invalid_argument("Input image for binarization is empty");
    if (!(maxValue >= 0 && maxValue <= 255)) throw std::invalid_argument("Max value must be in range [0; 255]");
    if (gray.channels() != 1 || xywh.size() % 4 != 0) throw std::invalid_argument("expected a gray image and x,y,w,h quadruples");
    prl_cuda_ctx* c = context();
    cv::Mat out(gray.rows, gray.cols, CV_8UC1);
    check(c, prl_cuda_otsu_rects(c, gray.data, gray.rows, gray.cols, gray.step, xywh.empty() ? nullptr : xywh.data(),
                                 (int)(xywh.size() / 4), maxValue, out.data, out.step, nullptr));
    binarized = out;
}

double prl::thresholdOtsu(const cv::Mat& src, cv::Mat& dst, double maxValue)
{
    if (src.empty() || src.channels() != 1) throw std::invalid_argument("expected a non-empty gray image");
    prl_cuda_ctx* c = context();
    cv::Mat out(src.rows, src.cols, CV_8UC1);
    int thr = 0;
    check(c, prl_cuda_otsu_global(c, src.data, src.rows, src.cols, src.step, maxValue, out.data, out.step, &thr));
    dst = out;
    return (double)thr;
}

void prl::localOtsuEdges(const cv::Mat& imageToProc, cv::Mat& resultCanny, int GaussianBlurKernelSize,
                         double CannyUpperThresholdCoeff, double CannyLowerThresholdCoeff, int CannyMorphIters, int postDilate)
{
    if (imageToProc.empty()) throw std::invalid_argument("Image for histograms extraction is empty");          // :248-251
    if (GaussianBlurKernelSize < 3) throw std::invalid_argument("Gaussian blur kernel size is lesser than 3");   // :253-256
    if (CannyUpperThresholdCoeff < 0 || CannyUpperThresholdCoeff > 1)
        throw std::invalid_argument("Canny upper threshold coefficient isn't in range [0;1]");                   // :258-261
    if (CannyLowerThresholdCoeff < 0 || CannyLowerThresholdCoeff > 1)
        throw std::invalid_argument("Canny lower threshold coefficient isn't in range [0;1]");                   // :263-266
    if (CannyLowerThresholdCoeff > CannyUpperThresholdCoeff)
        throw std::invalid_argument("Canny lower threshold coefficient is greater than Canny upper threshold coefficient");
    if (imageToProc.channels() != 1) throw std::invalid_argument("expected a single-channel image");
    prl_cuda_ctx* c = context();
    cv::Mat out(imageToProc.rows, imageToProc.cols, CV_8UC1);
    check(c, prl_cuda_canny_edge_detection(c, imageToProc.data, imageToProc.rows, imageToProc.cols, imageToProc.step,
                                           GaussianBlurKernelSize, CannyUpperThresholdCoeff, CannyLowerThresholdCoeff,
                                           CannyMorphIters, postDilate, out.data, out.step));
    resultCanny = out;
}

void prl::binarizeLocalOtsu(cv::Mat& inputImage, cv::Mat& outputImage, double maxValue, double CLAHEClipLimit,
                            int GaussianBlurKernelSize, double CannyUpperThresholdCoeff, double CannyLowerThresholdCoeff,
                            int CannyMorphIters)
{
    if (inputImage.empty()) throw std::invalid_argument("Input image for binarization is empty");                // :47-50
    if (!(maxValue >= 0 && maxValue <= 255)) throw std::invalid_argument("Max value must be in range [0; 255]");  // :52-55
    if (GaussianBlurKernelSize < 3) throw std::invalid_argument("Gaussian blur kernel size is lesser than 3");
    if (CannyUpperThresholdCoeff < 0 || CannyUpperThresholdCoeff > 1 || CannyLowerThresholdCoeff < 0 ||
        CannyLowerThresholdCoeff > 1 || CannyLowerThresholdCoeff > CannyUpperThresholdCoeff)
        throw std::invalid_argument("Canny threshold coefficients must satisfy 0 <= lower <= upper <= 1");
    prl_cuda_ctx* c = context();
    cv::Mat out(inputImage.rows, inputImage.cols, CV_8UC1);
    int n = 0;
    check(c, prl_cuda_binarize_local_otsu(c, inputImage.data, inputImage.rows, inputImage.cols, inputImage.step,
                                          inputImage.channels(), maxValue, CLAHEClipLimit > 0 ? CLAHEClipLimit : 0.0,
                                          GaussianBlurKernelSize, CannyUpperThresholdCoeff, CannyLowerThresholdCoeff,
                                          CannyMorphIters, out.data, out.step, &n, nullptr, 0));
    outputImage = out;
}

void prl::removeLines(const cv::Mat& inputImage, cv::Mat& outputImage)
{
    if (inputImage.empty()) throw cvError("removeLines: empty image");     // (the reference would fail inside cv::threshold)
    prl_cuda_ctx* c = context();
    cv::Mat out(inputImage.rows, inputImage.cols, CV_8UC1);
    check(c, prl_cuda_remove_lines(c, inputImage.data, inputImage.rows, inputImage.cols, inputImage.step, inputImage.channels(),
                                   out.data, out.step));
    outputImage = out;
}
