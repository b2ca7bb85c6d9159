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
