// cvcompat/opencv2/highgui/highgui.hpp -- HEADLESS stand-in for the highgui calls of the reference's apps/demo.cpp
// (apps/demo.cpp:35-41,55-59,79-96): imread of 8-bit colour / 16-bit depth PNGs (a small decoder over zlib: link with -lz), glob of
// a directory, imshow / waitKey that show nothing.  DF_CVCOMPAT_VERBOSE=1 makes imshow report what it was given on stderr.
#pragma once
#include <opencv2/core/core.hpp>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_LOAD_IMAGE_COLOR 1
#define CV_LOAD_IMAGE_ANYDEPTH 2

namespace cv {

inline void imshow(const String &name, const Mat &m)
{
    static const bool verbose = std::getenv("DF_CVCOMPAT_VERBOSE") != 0;
    if (verbose) std::fprintf(stderr, "imshow %s %dx%d type %d\n", name.c_str(), m.rows, m.cols, m.type());
}
inline int waitKey(int = 0) { return -1; }
inline bool imwrite(const String &, const Mat &) { return false; }

// files of a directory (or the directory part of a pattern), full paths, like cv::glob(dir, out) for the demo
inline void glob(String pattern, std::vector<String> &result, bool = false)
{
    result.clear();
    struct stat st;
    String dir = pattern;
    if (!(::stat(dir.c_str(), &st) == 0 && S_ISDIR(st.st_mode))) {
        const size_t s = pattern.find_last_of('/');
        dir = s == String::npos ? String(".") : pattern.substr(0, s);
    }
    DIR *d = ::opendir(dir.c_str());
    if (!d) return;
    while (struct dirent *e = ::readdir(d)) {
        if (e->d_name[0] == '.') continue;
        const String path = dir + "/" + e->d_name;
        if (::stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode)) result.push_back(path);
    }
    ::closedir(d);
    std::sort(result.begin(), result.end());
}

namespace cvcompat_detail {
inline unsigned be32(const unsigned char *p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }

// non-interlaced PNG, grey / grey+alpha / RGB / RGBA, 8 or 16 bit -> rows of `channels` samples of `depth` bits (big-endian)
inline bool png_decode(const String &file, int &w, int &h, int &channels, int &depth, std::vector<unsigned char> &pix)
{
    FILE *f = std::fopen(file.c_str(), "rb");
    if (!f) return false;
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 33 || std::memcmp(&buf[0], sig, 8) != 0) return false;
    std::vector<unsigned char> idat;
    int ctype = -1;
    for (size_t pos = 8; pos + 12 <= buf.size();) {
        const unsigned len = be32(&buf[pos]);
        const unsigned char *type = &buf[pos + 4], *data = &buf[pos + 8];
        if (pos + 12 + len > buf.size()) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            w = (int)be32(data); h = (int)be32(data + 4); depth = data[8]; ctype = data[9];
            if (data[12] != 0 || (depth != 8 && depth != 16)) return false;          // interlaced / packed depths: not needed here
        } else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    channels = ctype == 0 ? 1 : ctype == 4 ? 2 : ctype == 2 ? 3 : ctype == 6 ? 4 : 0;
    if (!channels || w <= 0 || h <= 0) return false;
    const size_t bpp = (size_t)channels * depth / 8, stride = (size_t)w * bpp;
    std::vector<unsigned char> raw((stride + 1) * (size_t)h);
    uLongf outlen = (uLongf)raw.size();
    if (uncompress(&raw[0], &outlen, &idat[0], (uLong)idat.size()) != Z_OK || outlen != raw.size()) return false;
    pix.assign(stride * (size_t)h, 0);
    for (int y = 0; y < h; ++y) {
        const unsigned char *in = &raw[(stride + 1) * (size_t)y];
        unsigned char *out = &pix[stride * (size_t)y];
        const unsigned char *up = y ? out - stride : 0;
        const int ft = in[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? out[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            const int x = in[1 + i];
            out[i] = (unsigned char)(ft == 0 ? x : ft == 1 ? x + a : ft == 2 ? x + b : ft == 3 ? x + ((a + b) >> 1) : x + paeth(a, b, c));
        }
    }
    return true;
}
}  // namespace cvcompat_detail

#ifndef CV_8UC3
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#endif

// CV_LOAD_IMAGE_ANYDEPTH: single channel at the file's depth (16-bit depth maps stay u16); CV_LOAD_IMAGE_COLOR: 8-bit BGR
inline Mat imread(const String &file, int flags = CV_LOAD_IMAGE_COLOR)
{
    int w = 0, h = 0, ch = 0, depth = 0;
    std::vector<unsigned char> pix;
    Mat out;
    if (!cvcompat_detail::png_decode(file, w, h, ch, depth, pix)) return out;
    const size_t bps = (size_t)depth / 8, bpp = bps * ch;
    if (flags == CV_LOAD_IMAGE_ANYDEPTH || flags == CV_LOAD_IMAGE_GRAYSCALE) {
        const bool keep16 = flags == CV_LOAD_IMAGE_ANYDEPTH && depth == 16;
        out.create(h, w, keep16 ? CV_16U : CV_8U);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const unsigned char *p = &pix[((size_t)y * w + x) * bpp];          // first channel (grey, or red) -- big-endian samples
                if (keep16) out.ptr<unsigned short>(y)[x] = (unsigned short)((p[0] << 8) | p[1]);
                else out.ptr<unsigned char>(y)[x] = p[0];
            }
        return out;
    }
    out.create(h, w, CV_8UC3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const unsigned char *p = &pix[((size_t)y * w + x) * bpp];
            unsigned char r = p[0], g = p[0], b = p[0];
            if (ch >= 3) { g = p[bps]; b = p[2 * bps]; }
            unsigned char *o = out.ptr<unsigned char>(y) + 3 * x;
            o[0] = b; o[1] = g; o[2] = r;
        }
    return out;
}

}  // namespace cv

inline int cvWaitKey(int delay = 0) { return cv::waitKey(delay); }
