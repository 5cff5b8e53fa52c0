// imread_check.cpp -- reads a depth PNG (CV_LOAD_IMAGE_ANYDEPTH) and a colour PNG (CV_LOAD_IMAGE_COLOR) with the headless highgui
// stand-in (include/cvcompat) exactly as the reference's apps/demo.cpp does (:91-92) and prints size, type and pixel sums.
#include <opencv2/highgui/highgui.hpp>
#include <cstdio>

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    std::vector<cv::String> files;
    cv::glob(argv[1], files);
    cv::Mat depth = cv::imread(files.at(0), CV_LOAD_IMAGE_ANYDEPTH);
    cv::Mat image = cv::imread(argv[2], CV_LOAD_IMAGE_COLOR);
    if (depth.empty() || image.empty()) return 3;
    unsigned long long ds = 0, b = 0, g = 0, r = 0;
    for (int y = 0; y < depth.rows; ++y) for (int x = 0; x < depth.cols; ++x) ds += depth.ptr<unsigned short>(y)[x];
    for (int y = 0; y < image.rows; ++y) for (int x = 0; x < image.cols; ++x) { const unsigned char *p = image.ptr<unsigned char>(y) + 3 * x; b += p[0]; g += p[1]; r += p[2]; }
    cv::Mat disp;
    depth.convertTo(disp, CV_8U, 255.0 / 4000);
    unsigned long long cs = 0;
    for (int y = 0; y < disp.rows; ++y) for (int x = 0; x < disp.cols; ++x) cs += disp.ptr<unsigned char>(y)[x];
    std::printf("%d %d %d %llu %d %d %d %llu %llu %llu %llu %zu\n", depth.rows, depth.cols, depth.type(), ds, image.rows, image.cols, image.type(), b, g, r, cs, files.size());
    return 0;
}
