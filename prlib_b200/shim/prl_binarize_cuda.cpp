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

// The reference ends in a cv::Exception for anything but 8-bit data (type mismatch at `gray > thresholds`, the
// CV_8UC1 assertion of THRESH_OTSU, cvtColor's depth/channel checks); the device path reads bytes, so refuse here.
void require8U(const cv::Mat& m, bool allow134 = true)
{
    if (m.depth() != CV_8U) throw cvError("expected an 8-bit image (depth() == CV_8U)");
    const int ch = m.channels();
    if (!(ch == 1 || (allow134 && (ch == 3 || ch == 4)))) throw cvError("cvtColor: unsupported number of channels");
}

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
    require8U(imageInput);
    prl_cuda_ctx* c = context();
    const int ch = imageInput.channels();
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

// thread-local page-locked staging of the batch forms (prl_cuda_host_alloc), grown on demand
struct PinnedBuf {
    void* p = nullptr; size_t bytes = 0;
    ~PinnedBuf() { prl_cuda_host_free(p); }
    unsigned char* get(size_t need)
    {
        if (need > bytes) {
            prl_cuda_host_free(p); p = nullptr; bytes = 0;
            if (prl_cuda_host_alloc(need, &p) != PRL_OK) throw std::runtime_error(std::string("libprlib_cuda: ") + prl_cuda_last_error(nullptr));
            bytes = need;
        }
        return static_cast<unsigned char*>(p);
    }
};

void runLocalBatch(int method, const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize,
                   const double params[4], int morph)
{
    static thread_local PinnedBuf pin_in, pin_out;
    std::vector<cv::Mat> result(inputs.size());
    size_t i = 0;
    while (i < inputs.size()) {
        const cv::Mat& first = inputs[i];
        if (first.empty()) throw std::invalid_argument("Input image for binarization is empty");
        if (!((windowSize > 1) && ((windowSize % 2) == 1)))
            throw std::invalid_argument("Window size must satisfy the following condition: "
                                        "( (windowSize > 1) && ((windowSize % 2) == 1) ) ");
        require8U(first);
        if (first.channels() != 1) {                 // colour pages: single-image path (cvtColor on the device)
            cv::Mat in = first, out;                 // header copy: runLocal re-seats `in`, the caller's Mat is untouched
            runLocal(method, in, out, windowSize, params, morph);
            result[i++] = out;
            continue;
        }
        size_t j = i + 1;
        while (j < inputs.size() && !inputs[j].empty() && inputs[j].depth() == CV_8U && inputs[j].channels() == 1 &&
               inputs[j].rows == first.rows && inputs[j].cols == first.cols) ++j;
        const int n = (int)(j - i), rows = first.rows, cols = first.cols;
        int orows = 0, ocols = 0;
        const int rc = prl_cuda_output_shape(method, rows, cols, windowSize, &orows, &ocols);
        if (rc == PRL_E_EMPTY_ROI) throw cvError("empty processing rectangle: min(rows, cols) <= windowSize");
        if (rc != PRL_OK) throw std::invalid_argument("bad geometry");
        const size_t in_page = (size_t)rows * cols, out_page = (size_t)orows * ocols;
        unsigned char* hin = pin_in.get(in_page * n);
        unsigned char* hout = pin_out.get(out_page * n);
        for (int p = 0; p < n; ++p)
            for (int y = 0; y < rows; ++y) std::memcpy(hin + p * in_page + (size_t)y * cols, inputs[i + p].ptr(y), (size_t)cols);
        check(nullptr, prl_cuda_binarize_batch(nullptr, 0, method, hin, n, rows, cols, windowSize, params, morph, hout));
        for (int p = 0; p < n; ++p) {
            cv::Mat out(orows, ocols, CV_8UC1);
            for (int y = 0; y < orows; ++y) std::memcpy(out.ptr(y), hout + p * out_page + (size_t)y * ocols, (size_t)ocols);
            result[i + p] = out;
        }
        i = j;
    }
    outputs.swap(result);
}
}  // namespace

#define PRL_BATCH1(NAME, METHOD)                                                                                       \
    void prl::NAME(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize,                 \
                   double thresholdCoefficient, int morphIterationCount)                                              \
    {                                                                                                                  \
        const double p[4] = {thresholdCoefficient, 0, 0, 0};                                                          \
        runLocalBatch(METHOD, inputs, outputs, windowSize, p, morphIterationCount);                                   \
    }
PRL_BATCH1(binarizeSauvolaBatch, PRL_SAUVOLA)
PRL_BATCH1(binarizeNiblackBatch, PRL_NIBLACK)
PRL_BATCH1(binarizeWolfJolionBatch, PRL_WOLFJOLION)
PRL_BATCH1(binarizeNICKBatch, PRL_NICK)
#undef PRL_BATCH1

void prl::binarizeFengBatch(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, int windowSize,
                            double thresholdCoefficient_alpha1, double thresholdCoefficient_k1, double thresholdCoefficient_k2,
                            double thresholdCoefficient_gamma, int morphIterationCount)
{
    const double p[4] = {thresholdCoefficient_alpha1, thresholdCoefficient_k1, thresholdCoefficient_k2, thresholdCoefficient_gamma};
    runLocalBatch(PRL_FENG, inputs, outputs, windowSize, p, morphIterationCount);
}

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
    require8U(gray, false);
    if (xywh.size() % 4 != 0) throw std::invalid_argument("expected x,y,w,h quadruples");
    prl_cuda_ctx* c = context();
    cv::Mat out(gray.rows, gray.cols, CV_8UC1);
    check(c, prl_cuda_otsu_rects(c, gray.data, gray.rows, gray.cols, gray.step, xywh.empty() ? nullptr : xywh.data(),
                                 (int)(xywh.size() / 4), maxValue, out.data, out.step, nullptr));
    binarized = out;
}

double prl::thresholdOtsu(const cv::Mat& src, cv::Mat& dst, double maxValue)
{
    if (src.empty()) throw cvError("threshold: empty image");
    require8U(src, false);                                                  // THRESH_OTSU asserts CV_8UC1
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
    require8U(imageToProc, false);
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
    require8U(inputImage);
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
    if (inputImage.depth() != CV_8U || (inputImage.channels() != 1 && inputImage.channels() != 3))
        throw cvError("removeLines: expected an 8-bit image with 1 or 3 channels");   // cv::threshold's OTSU assertion
    prl_cuda_ctx* c = context();
    cv::Mat out(inputImage.rows, inputImage.cols, CV_8UC1);
    check(c, prl_cuda_remove_lines(c, inputImage.data, inputImage.rows, inputImage.cols, inputImage.step, inputImage.channels(),
                                   out.data, out.step));
    outputImage = out;
}

// ---- the adaptive-mean family ---------------------------------------------------------------------------------------
namespace
{
void runAdaptive(const cv::Mat& in, cv::Mat& out, const prl_adaptive_params& p)
{
    require8U(in);
    prl_cuda_ctx* c = context();
    cv::Mat res(in.rows, in.cols, CV_8UC1);
    check(c, prl_cuda_binarize_adaptive(c, in.data, in.rows, in.cols, in.step, in.channels(), &p, res.data, res.step));
    out = res;
}

// binarizeAT / binarizeAGT / binarizePureAdaptiveGaussian assign the Mat cv::adaptiveThreshold reads only inside
// `if (inputImageMat.channels() != 1)` (binarizeAT.cpp:55-58): a 1-channel image reaches adaptiveThreshold empty
void colourOnly(const cv::Mat& in)
{
    if (in.empty()) throw std::invalid_argument("Input image for binarization is empty");
    if (in.channels() == 1) throw cvError("adaptiveThreshold: src.type() == CV_8UC1 (the reference passes an empty Mat)");
}
}  // namespace

void prl::binarizeAT(const cv::Mat& inputImage, cv::Mat& outputImage, const int medianKernelSize, const double maxValue,
                     const int blockSize, const int shift)
{
    colourOnly(inputImage);
    prl_adaptive_params p = prl_adaptive_params();
    p.gray_first = 0; p.blur = 1; p.blur_ksize = medianKernelSize; p.method = 0; p.type = 0;
    p.maxval = maxValue; p.block_size = blockSize; p.delta = shift;
    runAdaptive(inputImage, outputImage, p);
}

void prl::binarizeAGT(const cv::Mat& inputImage, cv::Mat& outputImage, const int medianKernelSize, const double maxValue,
                      const int blockSize, const int shift)
{
    colourOnly(inputImage);
    prl_adaptive_params p = prl_adaptive_params();
    p.gray_first = 0; p.blur = 1; p.blur_ksize = medianKernelSize; p.method = 1; p.type = 0;
    p.maxval = maxValue; p.block_size = blockSize; p.delta = shift;
    runAdaptive(inputImage, outputImage, p);
}

void prl::binarizePureAdaptiveGaussian(const cv::Mat& inputImage, cv::Mat& outputImage, const double maxValue, const int blockSize,
                                       const int shift)
{
    colourOnly(inputImage);
    prl_adaptive_params p = prl_adaptive_params();
    p.gray_first = 1; p.blur = 0; p.method = 1; p.type = 0; p.maxval = maxValue; p.block_size = blockSize; p.delta = shift;
    runAdaptive(inputImage, outputImage, p);
}

// binarizeGAT.cpp:37-64 and binarizePureAdaptive.cpp:38-60 convert to gray first, so their `channels() != 1` branch is
// never taken and cv::adaptiveThreshold always sees an empty Mat
void prl::binarizeGAT(const cv::Mat& inputImage, cv::Mat&, const int, const double, const double, const double, const int, const int)
{
    if (inputImage.empty()) throw std::invalid_argument("Input image for binarization is empty");
    throw cvError("adaptiveThreshold: src.type() == CV_8UC1 (the reference passes an empty Mat)");
}

void prl::binarizePureAdaptive(const cv::Mat& inputImage, cv::Mat&, const double, const int, const int)
{
    if (inputImage.empty()) throw std::invalid_argument("Input image for binarization is empty");
    throw cvError("adaptiveThreshold: src.type() == CV_8UC1 (the reference passes an empty Mat)");
}

void prl::binarizeNativeAdaptive(cv::Mat& inputImage, cv::Mat& outputImage, bool isGaussianBlurReqiured, int medianBlurKernelSize,
                                 int GaussianBlurKernelSize, double GaussianBlurSigma, bool isAdaptiveThresholdCalculatedByGaussian,
                                 double adaptiveThresholdingMaxValue, int adaptiveThresholdingBlockSize, double adaptiveThresholdingShift,
                                 int bilateralFilterBlockSize, double bilateralFilterColorSigma, double bilateralFilterSpaceSigma)
{
    if (inputImage.empty()) throw std::invalid_argument("Input image for binarization is empty");
    if (!(adaptiveThresholdingMaxValue >= 0 && adaptiveThresholdingMaxValue <= 255))
        throw std::invalid_argument("Max value must be in range [0; 255]");
    require8U(inputImage);
    prl_cuda_ctx* c = context();
    if (inputImage.channels() > 1) {                       // cv::cvtColor(inputImage, inputImage, COLOR_BGR2GRAY): observable
        cv::Mat gray(inputImage.rows, inputImage.cols, CV_8UC1);
        check(c, prl_cuda_bgr2gray(c, inputImage.data, inputImage.rows, inputImage.cols, inputImage.step, inputImage.channels(),
                                   gray.data, gray.step));
        inputImage = gray;
    }
    prl_adaptive_params p = prl_adaptive_params();
    p.gray_first = 1; p.blur = isGaussianBlurReqiured ? 2 : 1;
    p.blur_ksize = isGaussianBlurReqiured ? GaussianBlurKernelSize : medianBlurKernelSize;
    p.blur_sigma = GaussianBlurSigma; p.assert_ksize = 1;
    p.method = isAdaptiveThresholdCalculatedByGaussian ? 1 : 0; p.type = 1;
    p.maxval = adaptiveThresholdingMaxValue; p.check_maxval = 1;
    p.block_size = adaptiveThresholdingBlockSize; p.auto_block = 1; p.delta = adaptiveThresholdingShift; p.invert_if_dark = 1;
    // cv::bilateralFilter of the result (binarizeNativeAdaptive.cpp:116-134); a sigma <= 0 comes back as PRL_E_INVALID after the
    // threshold's own checks, i.e. as std::invalid_argument in the reference's order
    p.bilateral_d = bilateralFilterBlockSize >= 3 ? bilateralFilterBlockSize : 0;
    p.bilateral_sigma_color = bilateralFilterColorSigma; p.bilateral_sigma_space = bilateralFilterSpaceSigma;
    cv::Mat res;
    runAdaptive(inputImage, res, p);
    outputImage = res;
}
