// Drives the cv::Mat shim on a GPU and compares with the plain-C oracle (test infrastructure).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>
#include "prl_binarize_cuda.h"

extern "C" {
void oracle_synth_page(uint8_t*, size_t, int, int, uint32_t, uint32_t);
int oracle_binarize_local(const uint8_t*, int, int, size_t, int, int, const double*, uint8_t*, uint8_t*, double*);
void oracle_morph(uint8_t*, int, int, int);
int oracle_output_shape(int, int, int, int, int*, int*);
int oracle_otsu_global(const uint8_t*, int, int, size_t, double, uint8_t*, size_t);
}

static int fails = 0;
#define EXPECT(cond, what) do { if (!(cond)) { std::printf("FAIL: %s\n", what); ++fails; } } while (0)

static bool same(const cv::Mat& m, const std::vector<uint8_t>& ref, int rows, int cols)
{
    if (m.rows != rows || m.cols != cols) return false;
    for (int y = 0; y < rows; ++y)
        if (std::memcmp(m.ptr(y), ref.data() + (size_t)y * cols, (size_t)cols) != 0) return false;
    return true;
}

int main()
{
    const int rows = 700, cols = 900;
    cv::Mat page(rows, cols, CV_8UC1);
    oracle_synth_page(page.data, page.step, rows, cols, 2024, 3);

    struct Case { int method; int window; double p[4]; int morph; };
    const Case cases[] = {{0, 15, {0.2, 0, 0, 0}, 0}, {0, 101, {0.01, 0, 0, 0}, 2}, {1, 15, {-0.2, 0, 0, 0}, 0},
                          {2, 15, {0.5, 0, 0, 0}, 0}, {3, 21, {-0.01, 0, 0, 0}, 0}, {4, 21, {0.75, 0.2, 0.03, 2.0}, 2}};
    for (const Case& c : cases) {
        int orow, ocol;
        oracle_output_shape(c.method, rows, cols, c.window, &orow, &ocol);
        std::vector<uint8_t> want((size_t)orow * ocol);
        oracle_binarize_local(page.data, rows, cols, page.step, c.method, c.window, c.p, nullptr, want.data(), nullptr);
        oracle_morph(want.data(), orow, ocol, c.morph);
        cv::Mat in = page.clone(), out;
        switch (c.method) {
        case 0: prl::binarizeSauvola(in, out, c.window, c.p[0], c.morph); break;
        case 1: prl::binarizeNiblack(in, out, c.window, c.p[0], c.morph); break;
        case 2: prl::binarizeWolfJolion(in, out, c.window, c.p[0], c.morph); break;
        case 3: prl::binarizeNICK(in, out, c.window, c.p[0], c.morph); break;
        default: prl::binarizeFeng(in, out, c.window, c.p[0], c.p[1], c.p[2], c.p[3], c.morph); break;
        }
        EXPECT(same(out, want, orow, ocol), "mask equals the oracle");
        const int h = c.window / 2;      // the reference leaves the padded gray image in the input Mat
        EXPECT(in.rows == rows + 2 * h && in.cols == cols + 2 * h, "input Mat replaced by the padded image");
        EXPECT(in.ptr(0)[0] == page.ptr(0)[0] && in.ptr(h)[h] == page.ptr(0)[0] &&
               in.ptr(in.rows - 1)[in.cols - 1] == page.ptr(rows - 1)[cols - 1], "replicate border");
    }
    {   // 3-channel input (what cv::imread hands to the samples): cvtColor on the device, same masks as the gray path
        cv::Mat bgr(rows, cols, CV_8UC3), gray(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const unsigned char b = page.ptr(y)[x], g = (unsigned char)(b ^ 0x5a), r = (unsigned char)(255 - b);
                unsigned char* p = bgr.ptr(y) + 3 * x;
                p[0] = b; p[1] = g; p[2] = r;
                gray.ptr(y)[x] = (unsigned char)((b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15);
            }
        cv::Mat in3 = bgr.clone(), in1 = gray.clone(), out3, out1;
        prl::binarizeSauvola(in3, out3, 15, 0.2, 0);
        prl::binarizeSauvola(in1, out1, 15, 0.2, 0);
        bool same_mask = out3.rows == out1.rows && out3.cols == out1.cols;
        for (int y = 0; same_mask && y < out1.rows; ++y) same_mask = std::memcmp(out3.ptr(y), out1.ptr(y), (size_t)out1.cols) == 0;
        EXPECT(same_mask, "BGR input gives the mask of its gray conversion");
        EXPECT(in3.channels() == 1 && in3.rows == rows + 14 && in3.ptr(7)[7] == gray.ptr(0)[0], "BGR input replaced by the padded GRAY image");
    }
    {   // header defaults
        cv::Mat in = page.clone(), out;
        prl::binarizeNICK(in, out);
        EXPECT(out.rows == rows - 21 && out.cols == cols - 21, "NICK default window 21");
    }
    {   // error behaviour of the reference
        cv::Mat empty, out, in = page.clone();
        bool t1 = false, t2 = false, t3 = false;
        try { prl::binarizeSauvola(empty, out); } catch (const std::invalid_argument&) { t1 = true; }
        try { prl::binarizeSauvola(in, out, 14); } catch (const std::invalid_argument&) { t2 = true; }
        cv::Mat small(30, 40, CV_8UC1); std::memset(small.data, 100, 1200);
        try { prl::binarizeNICK(small, out, 101); } catch (const cv::Exception&) { t3 = true; }
        EXPECT(t1 && t2 && t3, "invalid_argument / cv::Exception like the reference");
    }
    {   // 16-bit / float input: the reference ends in a cv::Exception, never in a mask of reinterpreted bytes
        cv::Mat u16(64, 64, CV_16UC1), f32(64, 64, CV_32FC1), out;
        std::memset(u16.data, 7, 64 * 64 * 2); std::memset(f32.data, 0, 64 * 64 * 4);
        int thrown = 0;
        try { prl::binarizeSauvola(u16, out, 15, 0.2, 0); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizeFeng(f32, out); } catch (const cv::Exception&) { ++thrown; }
        try { prl::thresholdOtsu(u16, out); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizeLocalOtsu(f32, out); } catch (const cv::Exception&) { ++thrown; }
        try { prl::removeLines(u16, out); } catch (const cv::Exception&) { ++thrown; }
        cv::Mat c4(64, 64, CV_8UC4); std::memset(c4.data, 9, 64 * 64 * 4);
        try { prl::removeLines(c4, out); } catch (const cv::Exception&) { ++thrown; }
        EXPECT(thrown == 6 && u16.rows == 64 && u16.depth() == CV_16U, "non-8-bit input -> cv::Exception, input untouched");
    }
    {   // batch forms: pages of two sizes and one BGR page in one vector; outputs equal the single-image calls
        std::vector<cv::Mat> in, out;
        for (int p = 0; p < 5; ++p) { cv::Mat m(300, 421, CV_8UC1); oracle_synth_page(m.data, m.step, 300, 421, 2024, p); in.push_back(m); }
        cv::Mat bgr(200, 300, CV_8UC3);
        for (int y = 0; y < 200; ++y) for (int x = 0; x < 900; ++x) bgr.ptr(y)[x] = (unsigned char)((x * 7 + y * 13) & 255);
        in.push_back(bgr);
        for (int p = 0; p < 3; ++p) { cv::Mat m(220, 180, CV_8UC1); oracle_synth_page(m.data, m.step, 220, 180, 2024, 10 + p); in.push_back(m); }
        const cv::Mat keep0 = in[0].clone();
        prl::binarizeSauvolaBatch(in, out, 15, 0.2, 1);
        bool ok = out.size() == in.size() && in[0].rows == 300 && std::memcmp(in[0].data, keep0.data, 300 * 421) == 0;
        for (size_t i = 0; ok && i < in.size(); ++i) {
            cv::Mat a = in[i].clone(), single;
            prl::binarizeSauvola(a, single, 15, 0.2, 1);
            ok = single.rows == out[i].rows && single.cols == out[i].cols;
            for (int y = 0; ok && y < single.rows; ++y) ok = std::memcmp(single.ptr(y), out[i].ptr(y), (size_t)single.cols) == 0;
        }
        EXPECT(ok, "binarizeSauvolaBatch equals per-image binarizeSauvola, inputs untouched");
        std::vector<cv::Mat> outw, outf;
        prl::binarizeWolfJolionBatch(in, outw, 21, 0.5, 0);
        prl::binarizeFengBatch(in, outf);
        cv::Mat a = in[7].clone(), single;
        prl::binarizeWolfJolion(a, single, 21, 0.5, 0);
        bool okw = outw.size() == in.size() && single.rows == outw[7].rows && std::memcmp(single.data, outw[7].data, (size_t)single.rows * single.cols) == 0;
        EXPECT(okw && outf.size() == in.size() && outf[0].rows == 300 - 21, "Wolf-Jolion / Feng batch forms");
        bool thrown = false;
        std::vector<cv::Mat> bad(2); bad[0] = in[0];
        try { prl::binarizeNICKBatch(bad, out); } catch (const std::invalid_argument&) { thrown = true; }
        EXPECT(thrown, "empty page inside a batch -> invalid_argument");
    }
    {   // the adaptive-mean family: reference signatures, the reference's quirks, properties of the result
        cv::Mat bgr(200, 300, CV_8UC3), out, gout;
        for (int y = 0; y < 200; ++y) for (int x = 0; x < 900; ++x) bgr.ptr(y)[x] = (unsigned char)((x * 7 + y * 13 + (x * y) % 31) & 255);
        prl::binarizeAT(bgr, out, 5, 255, 19, 9);
        bool ok = out.rows == 200 && out.cols == 300 && out.channels() == 1;
        for (int y = 0; ok && y < 200; ++y) for (int x = 0; x < 300; ++x) ok = ok && (out.ptr(y)[x] == 0 || out.ptr(y)[x] == 255);
        EXPECT(ok, "binarizeAT on a BGR image: 0/255 mask of the image size");
        prl::binarizeAGT(bgr, out, 5, 255, 19, 9);
        prl::binarizePureAdaptiveGaussian(bgr, out, 200, 15, 4);
        EXPECT(out.rows == 200 && out.cols == 300, "binarizeAGT / binarizePureAdaptiveGaussian run on BGR");
        int thrown = 0;
        cv::Mat g = page.clone();
        try { prl::binarizeAT(g, out, 5, 255, 19, 9); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizeAGT(g, out, 5, 255, 19, 9); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizePureAdaptiveGaussian(g, out, 255, 19, 9); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizeGAT(bgr, out, 7, 1.0, 1.0, 255, 19, 9); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizePureAdaptive(bgr, out, 255, 19, 9); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizeAT(cv::Mat(), out, 5, 255, 19, 9); } catch (const std::invalid_argument&) { ++thrown; }
        EXPECT(thrown == 6, "the family's exceptions are the reference's (1-channel AT/AGT/PAG, GAT, PureAdaptive, empty input)");
        cv::Mat in = bgr.clone();
        prl::binarizeNativeAdaptive(in, out);
        EXPECT(in.channels() == 1 && in.rows == 200 && out.rows == 200 && out.cols == 300, "binarizeNativeAdaptive: input converted to gray in place");
        long white = 0;
        for (int y = 0; y < 200; ++y) for (int x = 0; x < 300; ++x) white += out.ptr(y)[x] == 255;
        EXPECT(white * 2 >= 200L * 300, "binarizeNativeAdaptive: mean >= 128 after the inversion rule");
        thrown = 0;
        try { prl::binarizeNativeAdaptive(in, out, false, 5, 7, 150.0, true, 300.0); } catch (const std::invalid_argument&) { ++thrown; }
        try { prl::binarizeNativeAdaptive(in, out, false, 2); } catch (const cv::Exception&) { ++thrown; }
        try { prl::binarizeNativeAdaptive(in, out, false, 5, 7, 150.0, true, 255.0, 20); } catch (const cv::Exception&) { ++thrown; }
        EXPECT(thrown == 3, "binarizeNativeAdaptive argument errors");
        // the optional bilateral filter of the mask (binarizeNativeAdaptive.cpp:116-134): a mask in, a gray-level image out;
        // a sigma <= 0 is std::invalid_argument, but only after the threshold's own cv::Exception
        cv::Mat plain, smoothed;
        prl::binarizeNativeAdaptive(in, plain);
        prl::binarizeNativeAdaptive(in, smoothed, false, 5, 7, 150.0, true, 255.0, 19, 9, 5);
        long levels = 0, moved = 0;
        for (int y = 0; y < 200; ++y) for (int x = 0; x < 300; ++x) {
            const int v = smoothed.ptr(y)[x];
            levels += v != 0 && v != 255;
            moved += v != plain.ptr(y)[x];
        }
        EXPECT(smoothed.rows == 200 && smoothed.cols == 300 && levels > 0 && moved > 0, "binarizeNativeAdaptive: bilateral filter applied to the mask");
        thrown = 0;
        try { prl::binarizeNativeAdaptive(in, out, false, 5, 7, 150.0, true, 255.0, 19, 9, 7, 0.0); } catch (const std::invalid_argument&) { ++thrown; }
        try { prl::binarizeNativeAdaptive(in, out, false, 5, 7, 150.0, true, 255.0, 19, 9, 7, 150.0, -1.0); } catch (const std::invalid_argument&) { ++thrown; }
        try { prl::binarizeNativeAdaptive(in, out, false, 5, 7, 150.0, true, 255.0, 20, 9, 7, 0.0); } catch (const cv::Exception&) { ++thrown; }
        EXPECT(thrown == 3, "binarizeNativeAdaptive bilateral argument errors in the reference's order");
    }
    {   // Otsu
        std::vector<uint8_t> want((size_t)rows * cols);
        int thr = oracle_otsu_global(page.data, rows, cols, page.step, 255, want.data(), cols);
        cv::Mat dst;
        double t = prl::thresholdOtsu(page, dst, 255);
        EXPECT((int)t == thr && same(dst, want, rows, cols), "global Otsu equals the oracle");
    }
    {   // prl::binarizeLocalOtsu with the header defaults: every pixel is 0 or 255, black only inside the edge map's reach,
        // flat image -> no contour -> std::invalid_argument like RemoveChildrenContours (imageLibCommon.cpp:643-646)
        cv::Mat in = page.clone(), out, edges;
        prl::binarizeLocalOtsu(in, out);
        prl::localOtsuEdges(page, edges);
        bool ok = out.rows == rows && out.cols == cols && edges.rows == rows && edges.cols == cols;
        long black = 0;
        for (int y = 0; ok && y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const unsigned char v = out.ptr(y)[x];
                ok = ok && (v == 0 || v == 255);
                black += v == 0;
            }
        EXPECT(ok && black > 0, "binarizeLocalOtsu runs with the reference defaults");
        cv::Mat flat(64, 64, CV_8UC1); std::memset(flat.data, 180, 64 * 64);
        bool thrown = false;
        try { prl::binarizeLocalOtsu(flat, out); } catch (const std::invalid_argument&) { thrown = true; }
        EXPECT(thrown, "no contours -> invalid_argument");
    }
    std::printf(fails ? "shim_test: %d failure(s)\n" : "shim_test: all ok\n", fails);
    return fails ? 1 : 0;
}
